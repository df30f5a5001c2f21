"""One launch of prf_lpn_kernel worth ~10 ms for ncu: 32768 PRF seeds in live-row mode (98 304 cores x 4 161 AES blocks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pvac_hfhe_cppbyv_b200 import api
eng = api.Engine(0, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX)
eng.keygen(1)
rng = np.random.default_rng(1)
n = 32768
z, lo, hi = (rng.integers(0, 2**64, n, dtype=np.uint64) for _ in range(3))
for _ in range(3):
    out = eng.prf(z, lo, hi, 0)
print("prof_prf ok", eng.stats())
eng.close()
