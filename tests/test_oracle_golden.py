"""The oracle (oracle/pvac_oracle.c) against the committed golden fixtures, which were generated from the unmodified
reference (oracle/make_golden.py) and against the reference repository's own golden file bounty2_data/{a,b,sum}.ct."""
import hashlib
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, ct_digest, ct_equal, hexwords, load_npz

SEED = (0x1111, 0x2222, 0x3333)
DOMS = ["pvac.prf.r.1", "pvac.prf.r.2", "pvac.prf.r.3", "pvac.prf.noise.1", "pvac.prf.noise.2", "pvac.prf.noise.3", "pvac.dom.toeplitz"]


@pytest.fixture(scope="module")
def synth(port, synth_keys_raw):
    r = synth_keys_raw
    return port.Keys.from_raw(r["canon_tag"], r["H_digest"], None, None, r["prf_k"], r["lpn_s"])


def test_sha256_abc(port, kat):
    # the reference's only hard known-answer test (tests/test_prf.cpp:11-25)
    assert port.sha256(b"abc").hex() == kat["sha256_abc"] == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"


def test_fnv_and_aes(port, kat):
    for d in DOMS:
        assert f"{port.lib().orc_fnv1a(d.encode()):016x}" == kat["fnv1a"][d]
    key = bytes.fromhex(kat["aes_key"])
    assert hexwords(port.aes_ctr_words(key, 0, 6)) == kat["aes_ctr0_words"]
    assert hexwords(port.aes_ctr_words(key, 2**64 - 2, 8)) == kat["aes_ctr_wrap_words"]   # counter wraps without carry


def test_fp(port, kat):
    for c in kat["fp"]:
        a = [int(x, 16) for x in c["a"]]
        b = [int(x, 16) for x in c["b"]]
        assert hexwords(port.fp_add(a, b)) == c["add"]
        assert hexwords(port.fp_sub(a, b)) == c["sub"]
        assert hexwords(port.fp_mul(a, b)) == c["mul"]
        assert hexwords(port.fp_neg(a)) == c["neg"]
        if c["inv"]:
            assert hexwords(port.fp_inv(a)) == c["inv"]


def test_prf_pieces(port, synth, kat):
    key, nonce = synth.derive_aes_key(*SEED, "pvac.prf.r.1")
    assert key.hex() == kat["derive_key_r1"] and f"{nonce:016x}" == kat["derive_nonce_r1"]
    y = synth.lpn_make_ybits(*SEED, "pvac.prf.r.1", lpn_t=16384)       # all 16384 rows, as the reference
    assert hashlib.sha256(y.tobytes()).hexdigest() == kat["ybits_r1_sha256"]
    assert hexwords(y[:4]) == kat["ybits_r1_first4"]
    assert int(sum(bin(int(v)).count("1") for v in y)) == kat["ybits_r1_popcount"]
    assert f"{port.lib().orc_prg_layer_ztag(0x0123456789ABCDEF, 0x2222, 0x3333):016x}" == kat["ztag"]


def test_prf_faithful_and_live_rows(port, synth, kat):
    for rows in (16384, 127):
        synth.set_lpn_t(rows)
        for d in DOMS[:6]:
            assert hexwords(synth.prf_R_core(*SEED, d)) == kat["prf_R_core"][d], (rows, d)
    synth.set_lpn_t(127)
    assert hexwords(synth.prf_R(*SEED)) == kat["prf_R"]
    assert hexwords(synth.prf_R_noise(*SEED)) == kat["prf_R_noise"]
    for key, val in kat["prf_noise_delta"].items():
        g, kd = (int(x) for x in key.split(","))
        assert hexwords(synth.prf_noise_delta(*SEED, g, kd)) == val
    assert hexwords(port.fp_inv(synth.prf_R(*SEED))) == kat["fp_inv_prf_R"]


def test_prg_choose_k_and_plan(port, synth, kat):
    words = [0x0123456789ABCDEF, 0x1111, 0x2222, 0x3333, 5, 1, 0x4444]
    assert list(port.prg_choose_k(128, 16384, "pvac.dom.x_seed", words)) == kat["choose_x"]
    assert list(port.prg_choose_k(128, 8192, "pvac.dom.noise", words)) == kat["choose_noise"]
    assert [list(synth.plan_noise(d)) for d in range(4)] == kat["plan_noise"]
    for n, nb in kat["buckets"].items():
        assert port.lib().orc_next_bkt(int(n)) == nb


def test_keygen(port, port_keys, kat):
    g = load_npz("keys_seed1.npz")
    e = port_keys.export()
    assert e["canon_tag"] == int(g["canon_tag"])
    for k in ("H_digest", "prf_k", "lpn_s", "powg"):
        assert np.array_equal(e[k], g[k]), k
    for c, h in kat["keygen1_H_col_sha256"].items():
        assert hashlib.sha256(e["H"][int(c)].tobytes()).hexdigest() == h
    sg = port_keys.sigma_from_H(0x1111, 0x2222, 0x3333, 5, 1, 0x4444)
    assert hashlib.sha256(sg.tobytes()).hexdigest() == kat["keygen1_sigma_sha256"]


def test_enc_mul_dec_golden(port, port_keys, kat):
    K = port_keys
    ca = K.enc_value(1000, 42)
    assert port.tape_draws() == kat["enc1000_draws"]
    cb = K.enc_value(2000, 2**64 - 1)
    assert port.tape_draws() == kat["enc2000_draws"]
    ok, k = ct_equal(port.ct_export(ca), load_npz("enc_seed1000.npz"))
    assert ok, k
    ok, k = ct_equal(port.ct_export(cb), load_npz("enc_seed2000.npz"))
    assert ok, k
    cp = K.ct_mul(3000, ca, cb)
    assert port.tape_draws() == kat["mul3000_draws"]
    dp, g = port.ct_export(cp), load_npz("mul_seed3000.npz")
    ok, k = ct_equal(dp, g, with_sigma=False)
    assert ok, k
    hashes = np.frombuffer(b"".join(hashlib.sha256(r.tobytes()).digest() for r in dp["sigma"]), np.uint8).reshape(-1, 32)
    assert np.array_equal(hashes, g["sigma_sha256"])
    assert hexwords(K.dec_value(ca)) == kat["dec_enc1000"]
    assert hexwords(K.dec_value(cp)) == kat["dec_mul3000"] == ["ffffffffffffffd6", "0000000000000029"]   # 42*(2^64-1)


def test_chain_golden(port, port_keys, chain):
    K = port_keys
    ca, cb = K.enc_value(1000, 42), K.enc_value(2000, 2**64 - 1)
    cp = K.ct_mul(3000, ca, cb)
    cs, cd = K.ct_add(ca, cb), K.ct_sub(ca, cb)
    assert ct_digest(port.ct_export(cs)) == chain["add"]
    assert ct_digest(port.ct_export(cd)) == chain["sub"]
    assert hexwords(K.dec_value(cs)) == chain["dec_add"] and hexwords(K.dec_value(cd)) == chain["dec_sub"]
    p2 = K.ct_mul(4000, cp, ca)
    d2 = port.ct_export(p2)
    assert ct_digest(d2) == chain["mul_pa"] and [len(d2["rule"]), len(d2["lid"])] == chain["mul_pa_counts"]
    p3 = K.ct_mul(5000, cs, cp)
    assert ct_digest(port.ct_export(p3)) == chain["mul_sp"] and hexwords(K.dec_value(p3)) == chain["dec_mul_sp"]
    sq = K.ct_mul(6000, cp, cp)
    dsq = port.ct_export(sq)
    assert ct_digest(dsq) == chain["sq"] and [len(dsq["rule"]), len(dsq["lid"])] == chain["sq_counts"]
    assert hexwords(K.dec_value(sq)) == chain["dec_sq"]
    assert ct_digest(port.ct_export(K.ct_scale(ca, [12345, 0]))) == chain["scale"]


# ---- the reference repository's own golden file: sum.ct = ct_add(a.ct, b.ct)  (tests/add.cpp:220-228)
def parse_wire(data):
    magic, ver, cnt = struct.unpack_from("<IIQ", data, 0)
    assert magic == 0x66699666 and ver == 1
    off, out = 16, []
    for _ in range(cnt):
        nL, nE = struct.unpack_from("<II", data, off)
        off += 8
        L = dict(rule=[], ztag=[], nlo=[], nhi=[], pa=[], pb=[])
        for _ in range(nL):
            r = data[off]
            off += 1
            if r == 1:
                pa, pb = struct.unpack_from("<II", data, off)
                off += 8
                z = lo = hi = 0
            else:
                z, lo, hi = struct.unpack_from("<QQQ", data, off)
                off += 24
                pa = pb = 0
            for k, v in zip(("rule", "ztag", "nlo", "nhi", "pa", "pb"), (r, z, lo, hi, pa, pb)):
                L[k].append(v)
        E = dict(lid=[], idx=[], ch=[], w=[], sigma=[])
        for _ in range(nE):
            lid, idx, ch, _pad, wlo, whi, nbits = struct.unpack_from("<IHBBQQI", data, off)
            off += 28
            assert nbits == 8192
            sg = np.frombuffer(data, np.uint64, 128, off)
            off += 1024
            E["lid"].append(lid); E["idx"].append(idx); E["ch"].append(ch); E["w"].append([wlo, whi]); E["sigma"].append(sg)
        out.append(dict(rule=np.array(L["rule"], np.uint8), ztag=np.array(L["ztag"], np.uint64), nlo=np.array(L["nlo"], np.uint64),
                        nhi=np.array(L["nhi"], np.uint64), pa=np.array(L["pa"], np.uint32), pb=np.array(L["pb"], np.uint32),
                        lid=np.array(E["lid"], np.uint32), idx=np.array(E["idx"], np.uint16), ch=np.array(E["ch"], np.uint8),
                        w=np.array(E["w"], np.uint64).reshape(-1, 2), sigma=np.array(E["sigma"], np.uint64).reshape(-1, 128)))
    assert off == len(data)
    return out


def read_bounty2():
    r = {}
    for n in ("a", "b", "sum"):
        with open(os.path.join(GOLDEN, "bounty2", n + ".ct"), "rb") as f:
            r[n] = parse_wire(f.read())
    return r


def test_bounty2_golden_add(port, port_keys):
    f = read_bounty2()
    assert len(f["a"]) == len(f["b"]) == len(f["sum"]) == 1
    a, b = port.ct_import(f["a"][0]), port.ct_import(f["b"][0])
    got = port.ct_export(port_keys.ct_add(a, b))
    ok, k = ct_equal(got, f["sum"][0])
    assert ok, k
