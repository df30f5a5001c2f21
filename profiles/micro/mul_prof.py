import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from pvac_hfhe_cppbyv_b200 import api
eng = api.Engine(0, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX); eng.keygen(1)
rng = np.random.default_rng(1)
n = 1024
X, Y = eng.enc_value(rng.integers(0, 2**64, n, dtype=np.uint64), 2), eng.enc_value(rng.integers(0, 2**64, n, dtype=np.uint64), 3)
for k in range(2): eng.ct_mul(X, Y, 10 + k).free()
print("ok")
