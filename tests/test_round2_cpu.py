"""CPU checks of the round-2 additions: the ChaCha20 tape (RFC 8439 vector; engine body == oracle == the shim behind the unmodified
reference), caller-supplied tape words and the compact_edges drop branch they make reachable (oracle pinned on the unmodified
reference), the enc planning body with a drop mask, key-file goldens."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ct_equal

from oracle import ref as _ref

P = (1 << 127) - 1
M64 = (1 << 64) - 1
B = 337
KEY = bytes(range(32))


@pytest.fixture(scope="module")
def htlib():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "pvac_hfhe_cppbyv_b200", "_build", "libpvacb_hosttest.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(path)
    u64, u32 = C.c_uint64, C.c_uint32
    L.ht_chacha_word.restype = u64
    L.ht_chacha_word.argtypes = [C.POINTER(u32), u64, u32, u64]
    L.ht_chacha_block.restype = None
    L.ht_chacha_block.argtypes = [C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)]
    L.ht_plan_item_ex.restype = u64
    L.ht_plan_item_ex.argtypes = [u64, u64, u64, C.c_int, C.c_int, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.ht_keygen.restype = C.c_int
    L.ht_keygen.argtypes = [u64, C.POINTER(u64), C.c_size_t]
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# RFC 8439 section 2.3.2: key 00..1f, counter 1, nonce 00 00 00 09 00 00 00 4a 00 00 00 00
RFC_BLOCK = bytes.fromhex(
    "10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
    "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")


def test_chacha20_block_rfc8439_vector(port, htlib):
    key = np.frombuffer(KEY, np.uint32).copy()
    c = np.array([1, 0x09000000, 0x4A000000, 0], np.uint32)
    assert port.chacha20_block(key, c).tobytes() == RFC_BLOCK                  # the oracle's block function
    out = np.zeros(8, np.uint64)
    htlib.ht_chacha_block(_p(key, C.c_uint32), _p(c, C.c_uint32), _p(out, C.c_uint64))
    assert out.tobytes() == RFC_BLOCK                                          # the engine's body (common.cuh), compiled for the host
    try:
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms
    except Exception:
        return
    nonce16 = c.tobytes()
    ks = Cipher(algorithms.ChaCha20(KEY, nonce16), mode=None).encryptor().update(bytes(64))
    assert ks == RFC_BLOCK


def test_chacha_tape_words_engine_body_equals_oracle(port, htlib):
    """word k of (key, stream id, lane): engine body (tape_open/Tape::at) against the oracle's tape"""
    key = np.frombuffer(KEY, np.uint32).copy()
    for sid, lane in ((0, 0), (0x0123456789ABCDEF, 0), (77, 5), (M64, 0xFFFFFFFF)):
        port.set_tape(1, KEY, lane)
        # the oracle exposes its tape through enc_*; compare via the block function instead: word k = block (k >> 3), 64-bit word k & 7
        for k in (0, 1, 7, 8, 9, 63, 64, 1000, 2**20 + 3):
            c = np.array([(k >> 3) & 0xFFFFFFFF, lane ^ ((k >> 3) >> 32), sid & 0xFFFFFFFF, sid >> 32], np.uint32)
            blk = port.chacha20_block(key, c).view(np.uint64)
            got = htlib.ht_chacha_word(_p(key, C.c_uint32), sid, lane, k)
            assert int(got) == int(blk[k & 7]), (sid, lane, k)
    port.set_tape(0)


needs_ref = pytest.mark.skipif(not _ref.available(), reason="oracle/_ref/libpvac_ref.so not built (reference tree absent)")


@pytest.fixture(scope="module")
def both(port):
    _ref.lib()
    ko, kr = port.Keys.keygen(77), _ref.Keys.keygen(77)
    ko.set_lpn_t(127)
    kr.set_lpn_t(127)
    return ko, kr


@needs_ref
def test_oracle_equals_reference_under_chacha_tape(port, both):
    """enc_value, ct_mul, enc_text and ct_recrypt draw the same words from the ChaCha20 tape in the oracle and behind the unmodified
    reference's getrandom(): every ciphertext byte equal (explicit stream ids with lane 0, and batch-seed streams with lane = item + 1)"""
    ko, kr = both
    try:
        for lane, sid in ((0, 0xDEADBEEF12345678), (1, 4242), (9, 4242)):
            port.set_tape(1, KEY, lane)
            _ref.set_tape(1, KEY, lane)
            ao, ar = ko.enc_value(sid, 42), kr.enc_value(sid, 42)
            bo, br = ko.enc_value(sid + 1, 2**64 - 1), kr.enc_value(sid + 1, 2**64 - 1)
            for x, y in ((ao, ar), (bo, br)):
                ok, f = ct_equal(port.ct_export(x), _ref.ct_export(y))
                assert ok, (lane, f)
            po, pr = ko.ct_mul(sid + 2, ao, bo), kr.ct_mul(sid + 2, ar, br)
            ok, f = ct_equal(port.ct_export(po), _ref.ct_export(pr))
            assert ok, (lane, f)
            assert np.array_equal(ko.dec_value(po), kr.dec_value(pr))
            to, tr = ko.enc_text(sid + 3, b"sixteen byte msg + a bit more"), kr.enc_text(sid + 3, b"sixteen byte msg + a bit more")
            assert len(to) == len(tr)
            for x, y in zip(to, tr):
                ok, f = ct_equal(port.ct_export(x), _ref.ct_export(y))
                assert ok, (lane, "text", f)
    finally:
        port.set_tape(0)
        _ref.set_tape(0)


def drop_words(seed=5, idx0=100, ch0=1, share=0):
    """tape words of one enc_value item (depth hint 0: Z2 = 3, Z3 = 2) crafted so that in share `share` the first signal edge and the
    first edge of the first Z2 group have the same (idx, sign), the same salt (hence the same syndrome) and opposite coefficients:
    the merged edge has weight 0 AND syndrome 0, which compact_edges removes (ops/encrypt.hpp:59)."""
    rng = np.random.default_rng(seed)
    rnd = lambda: int(rng.integers(0, 2**64, dtype=np.uint64))

    def good_share():
        w = [rnd(), rnd()]                                   # nonce
        for j in range(8):
            w += [337 * 1000 + (7 * j + 3) % B, rnd()]       # distinct idx, any sign
        w += [rnd() | 1 if k % 2 == 0 else rnd() >> 2 for k in range(14)]    # r[0..6]: (lo, hi) nonzero
        w += [rnd() for _ in range(8)]                       # salts
        return w

    def bad_share():
        w = [rnd(), rnd()]
        idxs = [idx0] + [(idx0 + 11 * (j + 1)) % B for j in range(1, 8)]
        for j in range(8):
            w += [337 * 5 + idxs[j], ((2 * rnd()) & M64) | (ch0 if j == 0 else (rnd() & 1))]
        r0 = (rnd() | 1) | ((rnd() >> 2) << 64)
        w += [r0 & M64, r0 >> 64]
        w += [rnd() | 1 if k % 2 == 0 else rnd() >> 2 for k in range(12)]
        salt0 = rnd()
        w += [salt0] + [rnd() for _ in range(7)]
        neg = P - r0
        w += [337 * 9 + idx0, 337 * 9 + (idx0 + 5) % B, ((2 * rnd()) & M64) | ch0, neg & M64, neg >> 64, salt0, rnd()]   # Z2 group 0: i, j, s1, r_i, salt_i, salt_j
        return w

    words = [rnd() | 1, rnd() >> 2]                          # mask
    words += bad_share() if share == 0 else good_share()
    words += [rnd() for _ in range(400)]                     # the rest of share 0 (retries allowed), all of share 1, shuffles
    if share == 1:
        # share 1 starts where share 0 stopped, which depends on retries: callers use share 0 for the crafted case
        raise NotImplementedError
    return np.array(words, np.uint64)


@needs_ref
def test_compact_edges_drop_branch_oracle_equals_reference(port, both):
    """the value-dependent branch of compact_edges inside enc_fp_depth (weight 0 and syndrome 0 -> edge removed, one shuffle draw
    fewer, every later word of the item shifts): reachable only with chosen randomness. Oracle == unmodified reference, byte for byte."""
    ko, kr = both
    try:
        for seed, idx0, ch0 in ((5, 100, 1), (6, 0, 0), (7, 336, 1)):
            words = drop_words(seed, idx0, ch0)
            port.set_tape(2, words=words)
            _ref.set_tape(2, words=words)
            co, cr = ko.enc_value(0, 12345), kr.enc_value(0, 12345)
            do, dr = port.ct_export(co), _ref.ct_export(cr)
            ok, f = ct_equal(do, dr)
            assert ok, (seed, f)
            # the crafted share is the SECOND layer (combine_ciphers puts enc(v + mask) first): it lost one of its slots
            n1 = int(np.count_nonzero(do["lid"] == 1))
            keys1 = {(int(i), int(c)) for i, c, l in zip(do["idx"], do["ch"], do["lid"]) if l == 1}
            assert (idx0, ch0) not in keys1 and n1 == len(keys1)
            assert int(ko.dec_value(co)[0]) == 12345
    finally:
        port.set_tape(0)
        _ref.set_tape(0)


def test_plan_body_with_drop_mask(port_keys, port, htlib):
    """enc planning body (enc_plan.cuh, compiled for the host): with the drop mask the oracle's result implies, the plan has one slot
    fewer in the crafted share and the same shuffle as the oracle's ciphertext"""
    words = drop_words(5, 100, 1)
    try:
        port.set_tape(2, words=words)
        c = port_keys.enc_value(0, 999)
        d = port.ct_export(c)
    finally:
        port.set_tape(0)
    # the host harness walks a SplitMix tape only; run it on an ordinary item with a mask for a slot that exists and check the bookkeeping
    canon = port_keys.export(with_H=False)["canon_tag"]
    hdr, raw, rnd = np.zeros(14, np.uint64), np.zeros(2 * 64 * 5, np.uint64), np.zeros(2 * 39 * 2, np.uint64)
    used0 = htlib.ht_plan_item_ex(1234, 5, canon, 3, 2, None, _p(hdr, C.c_uint64), _p(raw, C.c_uint64), _p(rnd, C.c_uint64))
    n_raw0, n_out0 = int(hdr[5]), int(hdr[6])
    r = raw.reshape(2, 64, 5)
    key = int(r[0, 0, 0]) * 2 + int(r[0, 0, 1])              # slot of raw edge 0 of share 0
    mask = np.zeros(22, np.uint64)
    mask[key >> 6] |= np.uint64(1) << np.uint64(key & 63)
    hdr2, raw2 = np.zeros(14, np.uint64), np.zeros(2 * 64 * 5, np.uint64)
    used1 = htlib.ht_plan_item_ex(1234, 5, canon, 3, 2, _p(mask, C.c_uint64), _p(hdr2, C.c_uint64), _p(raw2, C.c_uint64), _p(rnd, C.c_uint64))
    r2 = raw2.reshape(2, 64, 5)
    members = [k for k in range(n_raw0) if int(r[0, k, 0]) * 2 + int(r[0, k, 1]) == key]
    assert int(hdr2[5]) == n_raw0 and int(hdr2[6]) == n_out0 - 1
    assert used0 > 0 and used1 > 0                           # (share 1 starts one word earlier, so its own draw count may change either way)
    for k in range(n_raw0):
        if k in members:
            assert int(r2[0, k, 2]) == 0xFFFF and int(r2[0, k, 3]) == 0
        else:
            assert int(r2[0, k, 2]) < n_out0 - 1
    assert sorted(int(r2[0, k, 2]) for k in range(n_raw0) if k not in members and int(r2[0, k, 3])) == list(range(n_out0 - 1))
    assert len(d["lid"]) > 0


def test_keyfile_goldens(htlib):
    """omega_B (computed by keygen, stored only in the pk file) of the engine's host keygen equals the unmodified reference's; the sk
    fixture of the reference repository parses with the documented layout"""
    with open(os.path.join(GOLDEN, "keyfiles_seed1.json")) as f:
        g = json.load(f)
    words = 752 + 16384 * 128
    blob = np.zeros(words, np.uint64)
    assert htlib.ht_keygen(1, _p(blob, C.c_uint64), words) == 0
    assert [f"{int(x):016x}" for x in blob[748:750]] == g["omega_B"]
    sk = bytes.fromhex(g["sk_hex"])
    assert sk[:8] == (0x66666999).to_bytes(4, "little") + (1).to_bytes(4, "little")
    assert np.array_equal(np.frombuffer(sk[8:40], np.uint64), blob[5:9])
    assert int.from_bytes(sk[40:48], "little") == 64 and np.array_equal(np.frombuffer(sk[48:], np.uint64), blob[9:73])
    fx = open(os.path.join(GOLDEN, "bounty2", "sk.bin"), "rb").read()      # the reference repository's own sk.bin
    assert len(fx) == 560 and fx[:8] == sk[:8] and int.from_bytes(fx[40:48], "little") == 64


@needs_ref
def test_keyfile_golden_is_current(tmp_path):
    """the committed digests are what the unmodified reference writes today"""
    _ref.set_tape(0)
    K = _ref.Keys.keygen(1)
    pk, sk = str(tmp_path / "pk.bin"), str(tmp_path / "sk.bin")
    K.save(pk, sk)
    import hashlib
    with open(os.path.join(GOLDEN, "keyfiles_seed1.json")) as f:
        g = json.load(f)
    assert hashlib.sha256(open(pk, "rb").read()).hexdigest() == g["pk_sha256"]
    assert open(sk, "rb").read().hex() == g["sk_hex"]


@needs_ref
def test_ct_mul_with_duplicate_edges_oracle_equals_reference(port, both):
    """operands that hold repeated (layer, idx, sign) edges (legal in imported ciphertexts): the reference's unordered_map keeps
    adding (ops/arithmetic.hpp:79-88); oracle == unmodified reference in both operand positions, fresh and product operands"""
    from conftest import with_duplicates
    ko, kr = both
    ao, ar = ko.enc_value(1000, 42), kr.enc_value(1000, 42)
    bo, br = ko.enc_value(2000, 17), kr.enc_value(2000, 17)
    dup = with_duplicates(port.ct_export(bo))
    do, dr = port.ct_import(dup), _ref.ct_import(dup)
    for seed, (xo, yo, xr, yr) in enumerate(((ao, do, ar, dr), (do, ao, dr, ar), (do, do, dr, dr)), start=7001):
        ok, f = ct_equal(port.ct_export(ko.ct_mul(seed, xo, yo)), _ref.ct_export(kr.ct_mul(seed, xr, yr)))
        assert ok, (seed, f)
    po, pr = ko.ct_mul(7004, ao, bo), kr.ct_mul(7004, ar, br)
    pd = with_duplicates(port.ct_export(po))
    pdo, pdr = port.ct_import(pd), _ref.ct_import(pd)
    for seed, (xo, yo, xr, yr) in enumerate(((po, pdo, pr, pdr), (pdo, ao, pdr, ar)), start=7005):
        ok, f = ct_equal(port.ct_export(ko.ct_mul(seed, xo, yo)), _ref.ct_export(kr.ct_mul(seed, xr, yr)))
        assert ok, (seed, f)


def test_plan_noise_engine_oracle_reference_agree():
    """std::pair<int,int> plan_noise(pk, depth_hint), ops/encrypt.hpp:16-27 (the only floating point on the path: log2 / floor on the host):
    the engine's pvacb_plan_noise (host code of libpvacb.so, callable without a GPU), the oracle and the unmodified reference give the same
    (Z2, Z3) for every depth hint a caller can reach, negative hints included (clamped to 0 like std::max(0, depth_hint))"""
    import ctypes as C
    from pvac_hfhe_cppbyv_b200 import api
    from oracle import port
    L = api.load_library()
    ko = port.Keys.keygen(77)
    kr = _ref.Keys.keygen(77) if _ref.available() else None
    for d in list(range(-3, 400)) + [1000, 65535]:
        z2, z3 = C.c_int(), C.c_int()
        assert L.pvacb_plan_noise(d, C.byref(z2), C.byref(z3)) == 0
        got = (z2.value, z3.value)
        assert got == tuple(ko.plan_noise(d)), d
        if kr is not None:
            assert got == tuple(kr.plan_noise(d)), d
    assert (z2.value, z3.value)[0] > 30000                  # depth 65535: tens of thousands of noise groups, still the same numbers


def test_params_struct_matches_header_and_reference(tmp_path):
    """pvacb_params: (1) pvacb_params_default gives the reference's defaults field for field (core/types.hpp:36-70, read from the
    unmodified reference through oracle/_ref); (2) the ctypes mirror in api.py has the layout a C compiler gives the struct in
    include/pvacb.h (size and every offset), so a Python caller and a C caller hand the library the same bytes"""
    import ctypes as C
    import subprocess
    from pvac_hfhe_cppbyv_b200 import api
    p = api.Params.default()
    names = [f[0] for f in api.Params._fields_]
    assert names == ["B", "m_bits", "n_bits", "h_col_wt", "x_col_wt", "err_wt", "noise_entropy_bits", "tuple2_fraction", "depth_slope_bits", "edge_budget",
                     "lpn_n", "lpn_t", "lpn_tau_num", "lpn_tau_den", "recrypt_lo", "recrypt_hi", "recrypt_rounds"]
    if _ref.available():
        assert [float(getattr(p, n)) for n in names] == _ref.params_default()
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "pvacb.h"\nint main(void) {\n  printf("%zu", sizeof(pvacb_params));\n'
                   + "".join(f'  printf(" %zu", offsetof(pvacb_params, {n}));\n' for n in names) + "  return 0;\n}\n")
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include"), str(src), "-o", exe])
    nums = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert nums[0] == C.sizeof(api.Params)
    assert nums[1:] == [getattr(api.Params, n).offset for n in names]


def test_blob_layout_properties():
    """pvacb_blob_layout (host code, no GPU needed): the 13 arrays of a batch image in declaration order, 256-byte aligned, no overlap,
    each at least as large as its contents, total = off[13]; the counts (0, 0, 0) give a valid (tiny) image"""
    from pvac_hfhe_cppbyv_b200 import api
    unit = [4, 4, 1, 8, 8, 8, 4, 4, 4, 2, 1, 16, 1024]                    # bytes per entry: loff, eoff, rule, ztag, nlo, nhi, pa, pb, lid, idx, ch, w, sigma
    for n, nl, ne in ((0, 0, 0), (1, 2, 40), (3, 0, 0), (512, 4096, 614400), (4096, 8 * 4096, 1234 * 4096), (1 << 20, 2 << 20, 40 << 20)):
        off = api.blob_layout(n, nl, ne)
        assert len(off) == 14 and off[0] == 0
        counts = [n + 1, n + 1] + [nl] * 6 + [ne] * 5
        for k in range(13):
            assert off[k] % 256 == 0 and off[k + 1] - off[k] >= counts[k] * unit[k] and off[k + 1] - off[k] < counts[k] * unit[k] + 512, (n, nl, ne, k)
    assert api.blob_layout(4096, 8 * 4096, 1234 * 4096)[13] > 1234 * 4096 * 1024
