"""Where does an end-to-end ct_mul step (host buffers in, host buffers out) spend its time? Tuning aid for bench.py's e2e leg."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pvac_hfhe_cppbyv_b200 import api, shard
eng = api.Engine(0, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX); eng.keygen(1)
Me = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(1)
va, vb = rng.integers(0, 2**64, Me, dtype=np.uint64), rng.integers(0, 2**64, Me, dtype=np.uint64)
A, B = eng.enc_value(va, 1), eng.enc_value(vb, 2)
def pin(d):
    o = {}
    for k, v in d.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory(); o[k] = t.numpy(); o["_t" + k] = t
    return o
ha, hb = pin(eng.export_soa(A)), pin(eng.export_soa(B))
ia = {k: v for k, v in ha.items() if not k.startswith("_t")}; ib = {k: v for k, v in hb.items() if not k.startswith("_t")}
capE, capL = Me * 1400, Me * 8
spec = dict(loff=(Me + 1, np.uint32), eoff=(Me + 1, np.uint32), rule=(capL, np.uint8), ztag=(capL, np.uint64), nlo=(capL, np.uint64), nhi=(capL, np.uint64),
            pa=(capL, np.uint32), pb=(capL, np.uint32), lid=(capE, np.uint32), idx=(capE, np.uint16), ch=(capE, np.uint8), w=((capE, 2), np.uint64), sigma=((capE, 128), np.uint64))
keep = []
def pset():
    o = {}
    for k, (shape, dt) in spec.items():
        t = torch.empty(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=torch.uint8).pin_memory(); keep.append(t); o[k] = t.numpy().view(dt).reshape(shape)
    return o
bufs = [pset(), pset()]
T = {}
def tick(name, t0):
    T.setdefault(name, []).append(time.perf_counter() - t0)
pending = []
for k in range(8):
    t = time.perf_counter(); X = eng.import_soa(ia); Y = eng.import_soa(ib); tick("import_call", t)
    t = time.perf_counter(); eng.sync(); tick("import_sync", t)
    t = time.perf_counter(); P = eng.ct_mul(X, Y, tape_states=shard.item_tape_states(3000 + k, 0, Me)); tick("ct_mul_call", t)
    t = time.perf_counter(); eng.sync(); tick("ct_mul_sync", t)
    t = time.perf_counter()
    if pending:
        eng.export_wait(); P0 = pending.pop(0); P0.free()
    tick("export_wait", t)
    t = time.perf_counter(); pending.append(P); eng.export_soa_async(P, bufs[k & 1]); tick("export_async_call", t)
    X.free(); Y.free()
eng.export_wait()
for k, v in T.items(): print(f"{k:20s} " + " ".join(f"{x*1e3:7.2f}" for x in v))
# the same without overlap: blocking export
t = time.perf_counter(); d = eng.export_soa_async(P, bufs[0]); eng.export_wait(); dt = time.perf_counter() - t
nb = sum(v.nbytes for v in d.values())
print(f"export alone: {dt*1e3:.2f} ms for {nb/1e6:.1f} MB = {nb/dt/1e9:.1f} GB/s")
