"""BASELINE.json configs 2-5 at their stated sizes, through size-independent properties (the bit-exact comparisons against
the oracle are in test_gpu_parity.py; here a sample of every run is also compared bit for bit).

Sizes: configs 2 and 3 run at the full 2^20; config 4 runs step 1 of the chain over the full 2^18 ciphertexts, step 2 over 2^14
and step 3 over 2^7 of them (PVACB_FULL=1 runs steps 2 and 3 over all 2^18 / 2^10: 1.9 G edges of sigma, several GPU-minutes);
config 5 runs 2^18 items on the one GPU the test has (PVACB_FULL=1: 2^24; the multi-GPU run of the same pipeline is
profiles/mixed_pipeline.py under torchrun). Everything is streamed in tiles that fit HBM."""
import os

import numpy as np
import pytest

from conftest import ct_equal

pytestmark = pytest.mark.gpu

P127 = (1 << 127) - 1
FULL = os.environ.get("PVACB_FULL", "0") == "1"
M64 = (1 << 64) - 1


def fpv(a):
    return int(a[0]) | (int(a[1]) << 64)


from pvac_hfhe_cppbyv_b200.pipeline import mulmod127, run_mixed_pipeline  # noqa: E402


def test_mulmod127_helper():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2**64, 2000, dtype=np.uint64)
    b = rng.integers(0, 2**64, 2000, dtype=np.uint64)
    a[:3] = [M64, M64, 0]
    b[:3] = [M64, 1, 5]
    lo, hi = mulmod127(a, b)
    for i in range(2000):
        assert (int(lo[i]) | (int(hi[i]) << 64)) == int(a[i]) * int(b[i]) % P127


# ------------------------------------------------------------------------------------------------ config 2
def test_config2_add_sub_2p20_pairs(engine, api, port, port_keys):
    """ct_add / ct_sub of 2^20 synthetic fresh-shaped pairs (2 layers, 40 edges, 42 KB each; 88 GB in, 88 GB out per op):
    tiles of 2^16. Every tile: content checksums (sigma XOR, weight / id sums) of the result follow from those of the
    operands; one tile: a 64-pair sample bit for bit against the oracle."""
    total, tile = 1 << 20, 1 << 16
    done = 0
    for t in range(total // tile):
        A, B = engine.synthetic(tile, 20, 1000 + 2 * t), engine.synthetic(tile, 20, 1001 + 2 * t)
        ca, cb = engine.checksum(A), engine.checksum(B)
        S = engine.ct_add(A, B)
        cs = engine.checksum(S)
        assert cs["edges"] == ca["edges"] + cb["edges"] == tile * 80
        assert cs["xor_sigma"] == ca["xor_sigma"] ^ cb["xor_sigma"]
        assert cs["sum_wlo"] == (ca["sum_wlo"] + cb["sum_wlo"]) & M64 and cs["sum_whi"] == (ca["sum_whi"] + cb["sum_whi"]) & M64
        assert cs["sum_lid"] == (ca["sum_lid"] + cb["sum_lid"] + 2 * cb["edges"]) & M64          # B's layer ids shifted by |A.L| = 2
        assert cs["sum_idx"] == ca["sum_idx"] + cb["sum_idx"] and cs["sum_ch"] == ca["sum_ch"] + cb["sum_ch"]
        assert cs["layers"] == (ca["layers"] + cb["layers"]) & M64
        S.free()
        D = engine.ct_sub(A, B)
        cd = engine.checksum(D)
        # w -> p - w for every edge of B (synthetic weights are non-zero; p.lo = 2^64-1 >= w.lo, so no borrow between limbs)
        assert cd["xor_sigma"] == cs["xor_sigma"] and cd["sum_lid"] == cs["sum_lid"] and cd["edges"] == cs["edges"]
        assert cd["sum_wlo"] == (ca["sum_wlo"] + cb["edges"] * (P127 & M64) - cb["sum_wlo"]) & M64
        assert cd["sum_whi"] == (ca["sum_whi"] + cb["edges"] * (P127 >> 64) - cb["sum_whi"]) & M64
        if t == 0:
            k = 64
            ea, eb = api.split_items(engine.export_soa(engine.slice(A, 0, k))), api.split_items(engine.export_soa(engine.slice(B, 0, k)))
            es = api.split_items(engine.export_soa(engine.slice(D, 0, k)))
            Ssmall = engine.ct_add(engine.slice(A, 0, k), engine.slice(B, 0, k))
            ess = api.split_items(engine.export_soa(Ssmall))
            for i in range(0, k, 7):
                oa, ob = port.ct_import(ea[i]), port.ct_import(eb[i])
                ok, f = ct_equal(ess[i], port.ct_export(port_keys.ct_add(oa, ob)))
                assert ok, (i, f)
                ok, f = ct_equal(es[i], port.ct_export(port_keys.ct_sub(oa, ob)))
                assert ok, (i, f)
        D.free(); A.free(); B.free()
        done += tile
    assert done == total


# ------------------------------------------------------------------------------------------------ config 3
def test_config3_enc_2p20_values(engine, api, port, port_keys):
    """enc_value of 2^20 random u64 plaintexts (PRF + LPN noise + hypergraph syndromes), tiles of 2^16: every ciphertext
    decrypts to its plaintext, has 2 layers and 39-40 edges; 48 sampled ciphertexts are bit-identical to the oracle's."""
    total, tile = 1 << 20, 1 << 16
    rng = np.random.default_rng(2020)
    for t in range(total // tile):
        v = rng.integers(0, 2**64, tile, dtype=np.uint64)
        A = engine.enc_value(v, 40000 + t)
        nL, nE = A.totals()
        assert nL == 2 * tile and 38 * tile < nE <= 40 * tile
        d = engine.dec_value(A)
        assert np.array_equal(d[:, 0], v) and not d[:, 1].any()
        if t in (0, 7, 15):
            picks = [0, 1, 2, 3, 1000, 1001, 20000, 20001, 33333, 44444, 50000, 60000, 65533, 65534, 65535, 12345]
            for i in picks:
                got = api.split_items(engine.export_soa(engine.slice(A, i, 1)))[0]
                want = port.ct_export(port_keys.enc_value(port.item_stream_state(40000 + t, i), int(v[i])))
                ok, f = ct_equal(got, want)
                assert ok, (t, i, f)
        A.free()


# ------------------------------------------------------------------------------------------------ config 4
def test_config4_depth_chain_2p18(engine, api, port, port_keys):
    """tests/test_depth.cpp: c0 = enc(2), c <- c*c; 2^18 independent chains. Step 1 over all of them (tiles of 2^12 products,
    4.3 GB each), step 2 over 2^14 (2^18 with PVACB_FULL=1), step 3 over 2^7 (2^10): decrypts 4, 16, 256; layer counts 8, 32, 320
    like the reference; chain 0 of the first tile is compared bit for bit with the oracle at steps 1 and 2."""
    total = 1 << 18
    n2 = total if FULL else 1 << 14
    n3 = (1 << 10) if FULL else 1 << 7
    tile1, tile2, tile3 = 1 << 12, 1 << 10, 1 << 6
    for t in range(total // tile1):
        base = t * tile1
        from pvac_hfhe_cppbyv_b200 import shard
        c0 = engine.enc_value(np.full(tile1, 2, np.uint64), tape_states=shard.item_tape_states(51000, base, tile1))
        c1 = engine.ct_mul(c0, c0, tape_states=shard.item_tape_states(52000, base, tile1))
        d = engine.dec_value(c1)
        assert np.all(d[:, 0] == 4) and not d[:, 1].any()
        assert c1.totals()[0] == 8 * tile1
        if t == 0:
            o0 = port_keys.enc_value(port.item_stream_state(51000, 0), 2)
            o1 = port_keys.ct_mul(port.item_stream_state(52000, 0), o0, o0)
            ok, f = ct_equal(api.split_items(engine.export_soa(engine.slice(c1, 0, 1)))[0], port.ct_export(o1))
            assert ok, f
        if base < n2:
            for s in range(0, tile1, tile2):
                x1 = engine.slice(c1, s, tile2)
                c2 = engine.ct_mul(x1, x1, tape_states=shard.item_tape_states(53000, base + s, tile2))
                d2 = engine.dec_value(c2)
                assert np.all(d2[:, 0] == 16) and not d2[:, 1].any()
                assert c2.totals()[0] == 32 * tile2
                if base + s == 0:
                    o2 = port_keys.ct_mul(port.item_stream_state(53000, 0), o1, o1)
                    ok, f = ct_equal(api.split_items(engine.export_soa(engine.slice(c2, 0, 1), with_sigma=False))[0], port.ct_export(o2, with_sigma=False), with_sigma=False)
                    assert ok, f
                if base + s < n3:
                    for u in range(0, min(tile2, n3 - (base + s)), tile3):
                        x2 = engine.slice(c2, u, tile3)
                        c3 = engine.ct_mul(x2, x2, tape_states=shard.item_tape_states(54000, base + s + u, tile3))
                        d3 = engine.dec_value(c3)
                        assert np.all(d3[:, 0] == 256) and not d3[:, 1].any()
                        assert c3.totals()[0] == 320 * tile3
                        c3.free(); x2.free()
                c2.free(); x1.free()
        c1.free(); c0.free()


# ------------------------------------------------------------------------------------------------ config 5
def test_config5_mixed_pipeline(engine, api, port, port_keys):
    """2^24 mixed enc / ct_mul / dec (2^23 pairs), sharded by index: here the shard of ONE GPU (2^18 items; PVACB_FULL=1: all 2^24).
    Every product is checked against a*b mod p; pair 0 is checked bit for bit against the oracle."""
    from pvac_hfhe_cppbyv_b200 import shard
    items = (1 << 24) if FULL else (1 << 18)
    checked, bad = run_mixed_pipeline(engine, 0, items, 1 << 12)
    assert checked == items // 2 and bad == 0
    va, vb = shard.mix64(np.uint64(0x1234)), shard.mix64(np.uint64(0x1235))
    st = shard.item_tape_states(9000, 0, 2)
    oa, ob = port_keys.enc_value(int(st[0]), int(va)), port_keys.enc_value(int(st[1]), int(vb))
    op = port_keys.ct_mul(int(shard.item_tape_states(9001, 0, 1)[0]), oa, ob)
    A = engine.enc_value(np.array([va], np.uint64), tape_states=st[0:1])
    B = engine.enc_value(np.array([vb], np.uint64), tape_states=st[1:2])
    Pm = engine.ct_mul(A, B, tape_states=shard.item_tape_states(9001, 0, 1))
    ok, f = ct_equal(api.split_items(engine.export_soa(Pm))[0], port.ct_export(op))
    assert ok, f
