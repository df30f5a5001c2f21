"""Small workload touching every kernel once, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pvac_hfhe_cppbyv_b200 import api
eng = api.Engine(0, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX)
eng.keygen(1)
v = np.array([3, 5, 2**64 - 1, 0], np.uint64)
A, B = eng.enc_value(v, 1), eng.enc_value(v[::-1].copy(), 2)
S, D, P = eng.ct_add(A, B), eng.ct_sub(A, B), eng.ct_mul(A, B, 3)
Q = eng.ct_mul(P, S, 4)
d = eng.dec_value(Q)
exp = [int(a) * int(b) % ((1 << 127) - 1) * ((int(a) + int(b)) % ((1 << 127) - 1)) % ((1 << 127) - 1) for a, b in zip(v, v[::-1])]
assert [int(x[0]) | (int(x[1]) << 64) for x in d] == exp
eng.commit_ct(Q); eng.compact_edges(Q); eng.checksum(Q); eng.ct_neg(A); eng.ct_div_const(A, [7, 0])
eng.enc_value_depth(v, 3, 5); eng.enc_fp_depth(np.array([[5, 0], [7, 1]], np.uint64), 2, 6)
w = eng.export_wire(S); R = eng.import_wire(w); assert eng.export_wire(R) == w
eng.set_prf_mode(api.PRF_FAITHFUL)
eng.dec_value(eng.slice(A, 0, 1))
eng.synthetic(8, 20, 1)
print("sanitizer_run ok")
eng.close()
