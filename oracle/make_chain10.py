"""TEST INFRASTRUCTURE. Generates tests/golden/chain10.json: the reference's "perf 10 muls" scenario
(examples/basic_usage.cpp:257-264: prod = enc(1); 10 x prod = ct_mul(prod, enc(2))) under a fixed tape, whose tenth product
has 1.38 M edges, passes Params::edge_budget and is therefore rebuilt by guard_budget -> compact_edges
(ops/encrypt.hpp:106-111). Run with the unmodified reference when oracle/_ref is built (default), else with the oracle port.
Takes several minutes of CPU (2.7 M sigma_from_H evaluations).   python oracle/make_chain10.py [port|ref]"""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import port, ref

which = sys.argv[1] if len(sys.argv) > 1 else ("ref" if ref.available() else "port")
M = ref if which == "ref" else port
K = M.Keys.keygen(1)
K.set_lpn_t(127)          # bit-identical outputs, 129x less PRF work
t0 = time.time()
prod = K.enc_value(5000, 1)
out = {"impl": which, "steps": []}
for k in range(10):
    two = K.enc_value(5100 + k, 2)
    prod = K.ct_mul(5200 + k, prod, two)
    d = M.ct_export(prod)
    h = hashlib.sha256()
    for f in ("rule", "ztag", "nlo", "nhi", "pa", "pb", "lid", "idx", "ch", "w", "sigma"):
        h.update(np.ascontiguousarray(d[f]).tobytes())
    out["steps"].append({"layers": int(len(d["rule"])), "edges": int(len(d["lid"])), "sha256": h.hexdigest()})
    print(k, out["steps"][-1], f"{time.time() - t0:.0f}s", flush=True)
dec = K.dec_value(prod)
out["dec"] = [f"{int(x):016x}" for x in dec]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "chain10.json"), "w") as f:
    json.dump(out, f, indent=1)
print("done", out["dec"])
