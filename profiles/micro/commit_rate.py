"""commit_ct rate: 65 536 fresh ciphertexts (throughput regime) and 4 096 products (few long serial chains)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pvac_hfhe_cppbyv_b200 import api
eng = api.Engine(0, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX); eng.keygen(1)
v = np.arange(65536, dtype=np.uint64)
F = eng.enc_value(v, 1)
A, B = eng.slice(F, 0, 4096), eng.slice(F, 4096, 4096)
P = eng.ct_mul(A, B, 2)
for name, X in (("fresh_65536", F), ("products_4096", P)):
    nl, ne = X.totals()
    eng.commit_ct(X)
    t = time.perf_counter(); d = eng.commit_ct(X); dt = time.perf_counter() - t
    comp = (ne * 1057 + nl * 25 + len(X) * 64) / 64
    print(f"{name}: {dt*1e3:.2f} ms, {len(X)/dt:.0f} commit_ct/s, {comp/dt/1e9:.2f} G compressions/s")
