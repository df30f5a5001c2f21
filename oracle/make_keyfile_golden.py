"""TEST INFRASTRUCTURE ONLY. Key-file goldens from the UNMODIFIED reference (oracle/_ref): keygen under SplitMix64 tape state 1
(the keys of tests/golden/keys_seed1.npz), written in the reference's own pk / sk file format (tests/bounty2_test.cpp:145-192)
by oracle/ref_shim.cpp:ref_keys_save over the reference's PubKey / SecKey objects. The pk file is 16.8 MB, so its SHA-256 is
committed instead of the bytes, together with omega_B (which keygen computes and nothing but the file ever stores).

    python oracle/make_keyfile_golden.py      ->  tests/golden/keyfiles_seed1.json
"""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

ref.set_tape(0)
K = ref.Keys.keygen(1)
with tempfile.TemporaryDirectory() as d:
    pk, sk = os.path.join(d, "pk.bin"), os.path.join(d, "sk.bin")
    K.save(pk, sk)
    out = {
        "generator": "oracle/make_keyfile_golden.py: unmodified reference keygen (tape state 1), files written by ref_keys_save",
        "pk_bytes": os.path.getsize(pk), "pk_sha256": hashlib.sha256(open(pk, "rb").read()).hexdigest(),
        "sk_bytes": os.path.getsize(sk), "sk_hex": open(sk, "rb").read().hex(),
        "omega_B": [f"{int(x):016x}" for x in K.omega_B()],
    }
with open(os.path.join(ROOT, "tests", "golden", "keyfiles_seed1.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out["pk_bytes"], out["pk_sha256"], out["omega_B"])
