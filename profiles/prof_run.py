"""Small fixed workload for ncu (one GPU): a few launches of every hot kernel. Usage: python profiles/prof_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pvac_hfhe_cppbyv_b200 import api

eng = api.Engine(0, prf_mode=api.PRF_FAITHFUL, tape=api.TAPE_SPLITMIX)
eng.keygen(1)
rng = np.random.default_rng(1)
v = rng.integers(0, 2**64, 128, dtype=np.uint64)
A = eng.enc_value(v, 1)                 # faithful PRF: prf_lpn_kernel over 128*30 cores
eng.set_prf_mode(api.PRF_LIVE)
n = 1024
va, vb = rng.integers(0, 2**64, n, dtype=np.uint64), rng.integers(0, 2**64, n, dtype=np.uint64)
X, Y = eng.enc_value(va, 2), eng.enc_value(vb, 3)
for k in range(2):
    P = eng.ct_mul(X, Y, 10 + k)        # sigma_fused_kernel over ~1.2M edges
d = eng.dec_value(P)
assert all((int(d[i][0]) | (int(d[i][1]) << 64)) == int(va[i]) * int(vb[i]) % ((1 << 127) - 1) for i in range(n))
SA, SB = eng.synthetic(1 << 14, 20, 5), eng.synthetic(1 << 14, 20, 6)
for k in range(2):
    eng.ct_add(SA, SB).free()           # concat_kernel, 2.7 GB per launch
print("prof_run ok", eng.stats(), "edges of the last product batch:", P.totals()[1])
eng.close()
