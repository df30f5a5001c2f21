"""Quick device-timed micro-benchmarks of the hot kernels (used while tuning). python profiles/quick_bench.py [what ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pvac_hfhe_cppbyv_b200 import api

what = set(sys.argv[1:]) or {"prf", "sigma", "add", "check"}
eng = api.Engine(0, prf_mode=api.PRF_FAITHFUL, tape=api.TAPE_SPLITMIX)
eng.keygen(1)
rng = np.random.default_rng(1)
eng.profile_enable(True)

if "check" in what:   # bit-exactness smoke against golden digests before timing anything
    import json, hashlib
    chain = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "chain.json")))
    a = eng.enc_value([42], tape_states=[1000]); b = eng.enc_value([2**64 - 1], tape_states=[2000])
    p = eng.ct_mul(a, b, tape_states=[3000]); sq = eng.ct_mul(p, p, tape_states=[6000])
    it = api.split_items(eng.export_soa(sq))[0]
    h = hashlib.sha256()
    for k in ("rule", "ztag", "nlo", "nhi", "pa", "pb", "lid", "idx", "ch", "w", "sigma"):
        h.update(np.ascontiguousarray(it[k]).tobytes())
    print("golden chain digest ok:", h.hexdigest() == chain["sq"], flush=True)
    eng.profile_collect()

if "prf" in what:
    for mode, n in ((api.PRF_FAITHFUL, 4096), (api.PRF_LIVE, 1 << 19)):
        eng.set_prf_mode(mode)
        z, lo, hi = (rng.integers(0, 2**64, n, dtype=np.uint64) for _ in range(3))
        for rep in range(3):
            eng.stats_reset(); eng.profile_collect()
            eng.prf(z, lo, hi, 0)
            ms, l = eng.profile_collect()["prf_lpn"]
            blocks = eng.stats()["aes_blocks"]
        print(f"prf mode={mode} n={n}: lpn kernel {ms:.2f} ms, {blocks/ms/1e6:.2f} G AES blocks/s", flush=True)

if "sigma" in what:
    n = 1 << 20
    z, lo, hi, salt = (rng.integers(0, 2**64, n, dtype=np.uint64) for _ in range(4))
    idx = rng.integers(0, 337, n).astype(np.uint16); ch = rng.integers(0, 2, n).astype(np.uint8)
    # time through ct_mul-like volume: use pvacb_sigma_from_H but only keep kernel times
    for rep in range(3):
        eng.profile_collect()
        t0 = time.perf_counter()
        out = eng.sigma_from_H(z, lo, hi, idx, ch, salt)
        wall = time.perf_counter() - t0
        pr = eng.profile_collect()
    ms = pr['sigma'][0]
    print(f"sigma cfg={os.environ.get('PVACB_SIGMA_CFG', '0')} n={n}: fused {ms:.2f} ms ({ms*1e6/n:.2f} ns/edge, {n*131072/ms/1e9:.2f} TB/s L2 gather, "
          f"{n*68/ms/1e6:.2f} G SHA-256 compressions/s), wall {wall*1e3:.0f} ms", flush=True)
    import hashlib
    print("sigma rows digest", hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()[:16], flush=True)
    print("l2 probe GB/s", eng.l2_gather_probe(3), flush=True)

if "add" in what:
    n = 1 << 15
    A, B = eng.synthetic(n, 20, 1), eng.synthetic(n, 20, 2)
    for rep in range(4):
        eng.profile_collect()
        eng.ct_add(A, B).free()
        ms, l = eng.profile_collect()["concat"]
    byts = n * (2 * (40 * 1052 + 58) + 80 * 1052 + 108)
    print(f"add n={n}: concat {ms:.3f} ms, {byts/ms/1e6:.1f} GB/s", flush=True)
eng.close()
