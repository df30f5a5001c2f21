"""ncu workload for the kernels that only matter on DEEP ciphertexts: the dense path of mul_pairs_kernel
(products of products), dec_edges_kernel on 172 544-edge ciphertexts, commit_kernel on products.

    ncu --set full --clock-control none --import-source on -k regex:'mul_pairs_kernel|dec_edges_kernel|commit_kernel' -c 8 -o gpurun_out/r02_deep python profiles/prof_deep.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pvac_hfhe_cppbyv_b200 import api

P127 = (1 << 127) - 1
eng = api.Engine(0, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX)
eng.keygen(1)
n = 64
c = eng.enc_value(np.full(n, 2, np.uint64), 1)
want = 2
for step, keep in ((1, 64), (2, 32), (3, 16)):          # test_depth's chain c <- c*c; step 3 multiplies 10 784-edge operands (dense layers)
    s = eng.slice(c, 0, keep)
    c = eng.ct_mul(s, s, 10 + step)
    want = want * want % P127
    d = eng.dec_value(c)                                  # step 3: 16 x 172 544 edges through dec_edges_kernel
    assert all((int(x[0]) | (int(x[1]) << 64)) == want for x in d)
print("chain ok; step-3 batch:", len(c), "ciphertexts,", c.totals())
X, Y = eng.enc_value(np.arange(2048, dtype=np.uint64), 2), eng.enc_value(np.arange(2048, dtype=np.uint64) + 9, 3)
Pm = eng.ct_mul(X, Y, 4)
dg = eng.commit_ct(Pm)                                    # commit_kernel: 2048 chains of ~20 000 compressions
print("prof_deep ok", eng.stats())
eng.close()
