import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def chain():
    with open(os.path.join(GOLDEN, "chain.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def port():
    from oracle import port as p
    p.lib()
    return p


@pytest.fixture(scope="session")
def port_keys(port):
    """keygen(tape state 1) in the oracle, evaluating only the 127 live LPN rows (bit-identical outputs, ~130x faster)."""
    k = port.Keys.keygen(1)
    k.set_lpn_t(127)
    return k


@pytest.fixture(scope="session")
def synth_keys_raw():
    hd = np.arange(32, dtype=np.uint8)
    lpn_s = np.arange(1, 65, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    return dict(canon_tag=0x0123456789ABCDEF, H_digest=hd, prf_k=np.array([1, 2, 3, 4], np.uint64), lpn_s=lpn_s)


def ct_equal(a, b, with_sigma=True):
    keys = ["rule", "ztag", "nlo", "nhi", "pa", "pb", "lid", "idx", "ch", "w"] + (["sigma"] if with_sigma else [])
    for k in keys:
        if not np.array_equal(np.asarray(a[k]), np.asarray(b[k])):
            return False, k
    return True, None


def hexwords(a):
    return [f"{int(x):016x}" for x in np.asarray(a).ravel()]


def ct_digest(d):
    import hashlib
    h = hashlib.sha256()
    for k in ("rule", "ztag", "nlo", "nhi", "pa", "pb", "lid", "idx", "ch", "w", "sigma"):
        h.update(np.ascontiguousarray(d[k]).tobytes())
    return h.hexdigest()


P127 = (1 << 127) - 1


def with_duplicates(d):
    """a ciphertext with runs of equal (layer, idx, sign): every edge twice; in the copies, edge 0 cancels exactly
    (w -> p - w, same sigma: the merged edge is all-zero and must be dropped), edge 1 cancels in weight only (kept: sigma != 0)"""
    e = {k: np.concatenate([d[k], d[k]]) for k in ("lid", "idx", "ch", "w", "sigma")}
    n = len(d["lid"])
    for j in (0, 1):
        w = (int(d["w"][j][0]) | (int(d["w"][j][1]) << 64))
        e["w"][n + j] = [((P127 - w) % P127) & (2**64 - 1), ((P127 - w) % P127) >> 64]
    e["sigma"][n + 1] ^= np.uint64(0x5555)
    e["sigma"][n + 2:] ^= np.uint64(0xF0F0)
    out = {k: d[k] for k in ("rule", "ztag", "nlo", "nhi", "pa", "pb")}
    out.update(e)
    return out


@pytest.fixture(scope="session")
def api():
    from pvac_hfhe_cppbyv_b200 import api as a
    return a


# ---- GPU engine (only constructed by -m gpu tests)
@pytest.fixture(scope="session")
def engine():
    from pvac_hfhe_cppbyv_b200 import api
    # the SplitMix64 tape: what the committed golden vectors and the oracle's default streams are defined on
    eng = api.Engine(device=0, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX)
    eng.keygen(1)
    dbg = os.environ.get("PVACB_TEST_DEBUG", "")          # bit-exact test shapes of the library (pvacb_debug_set), chosen by the tests that re-run others
    if dbg == "sigma_test_shape":
        eng.debug_set(2, 1)
    elif dbg == "mul_device_sort":
        eng.debug_set(0, 1, 0)
    elif dbg == "mul_global_table":
        eng.debug_set(0, 0, 1)
    yield eng
    eng.close()
