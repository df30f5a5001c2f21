"""ct_mul wall time vs batch size (fixed host overhead vs per-pair device time). Tuning aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pvac_hfhe_cppbyv_b200 import api
eng = api.Engine(0, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX); eng.keygen(1)
rng = np.random.default_rng(1)
N = 4096
va, vb = rng.integers(0, 2**64, N, dtype=np.uint64), rng.integers(0, 2**64, N, dtype=np.uint64)
A, B = eng.enc_value(va, 1), eng.enc_value(vb, 2)
for n in (1, 64, 256, 512, 1024, 4096):
    X, Y = eng.slice(A, 0, n), eng.slice(B, 0, n)
    ts = []
    for r in range(6):
        eng.sync(); t = time.perf_counter(); P = eng.ct_mul(X, Y, 100 + r); eng.sync(); ts.append(time.perf_counter() - t); P.free()
    eng.profile_enable(True); eng.profile_collect()
    P = eng.ct_mul(X, Y, 99); pr = eng.profile_collect(); eng.profile_enable(False); P.free()
    print(f"n={n:5d}: wall ms " + " ".join(f"{x*1e3:7.2f}" for x in ts) + f" | sigma kernel {pr['sigma'][0]:.2f} ms")
    for nm, f in (("add", lambda: eng.ct_add(X, Y)), ("dec", None)):
        if f is None:
            eng.sync(); t = time.perf_counter(); eng.dec_value(X); dt = time.perf_counter() - t
        else:
            f().free(); eng.sync(); t = time.perf_counter(); f().free(); eng.sync(); dt = time.perf_counter() - t
        print(f"          {nm} wall {dt*1e3:.3f} ms")
