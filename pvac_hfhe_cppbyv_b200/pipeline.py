"""BASELINE.json config 5: the mixed enc / ct_mul / dec job over a range of the GLOBAL item index (host-side driver).

Items come in pairs (2j, 2j+1): plaintexts derived from the item index, both encrypted, multiplied, decrypted; only the
16-byte decrypts leave the device, everything else is streamed through HBM in tiles. A rank of a multi-GPU run calls
run_mixed_pipeline on its own range (shard.partition); tape states are those of the global indices, so the results do not
depend on how many GPUs share the job."""
import numpy as np

from . import shard


def mulmod127(a, b):
    """a*b mod 2^127-1 for uint64 arrays -> (lo, hi) uint64 arrays (vectorised; the product is < 2^128)"""
    a, b = np.asarray(a, np.uint64), np.asarray(b, np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    a0, a1, b0, b1 = a & m32, a >> np.uint64(32), b & m32, b >> np.uint64(32)
    with np.errstate(over="ignore"):
        p00, p01, p10, p11 = a0 * b0, a0 * b1, a1 * b0, a1 * b1
        mid = (p00 >> np.uint64(32)) + (p01 & m32) + (p10 & m32)
        lo = (p00 & m32) | ((mid & m32) << np.uint64(32))
        hi = p11 + (p01 >> np.uint64(32)) + (p10 >> np.uint64(32)) + (mid >> np.uint64(32))
        # fold bit 127: value = lo + 2^64*hi ; 2^127 = 1 (mod p)
        top = hi >> np.uint64(63)
        hi = hi & np.uint64(0x7FFFFFFFFFFFFFFF)
        lo2 = lo + top
        hi = hi + (lo2 < top).astype(np.uint64)
    return lo2, hi            # cannot equal p for 64-bit operands (a*b <= (2^64-1)^2 < 2p)


def run_mixed_pipeline(engine, first, count, tile, seed=9000):
    """items [first, first+count) of the mixed job: pairs (2j, 2j+1) of plaintexts derived from the GLOBAL item index ->
    enc, enc, ct_mul, dec. Returns (#pairs checked, #mismatches). Only the 16-byte decrypts leave the device."""
    assert first % 2 == 0 and count % 2 == 0
    bad = checked = 0
    for f0, c in shard.tiles(first // 2, count // 2, tile):          # in pairs
        j = np.arange(f0, f0 + c, dtype=np.uint64)
        va = shard.mix64(j * np.uint64(2) + np.uint64(0x1234))
        vb = shard.mix64(j * np.uint64(2) + np.uint64(0x1235))
        A = engine.enc_value(va, tape_states=shard.item_tape_states(seed, 2 * f0, 2 * c)[0::2])
        B = engine.enc_value(vb, tape_states=shard.item_tape_states(seed, 2 * f0, 2 * c)[1::2])
        Pm = engine.ct_mul(A, B, tape_states=shard.item_tape_states(seed + 1, f0, c))
        d = engine.dec_value(Pm)
        lo, hi = mulmod127(va, vb)
        bad += int(np.count_nonzero((d[:, 0] != lo) | (d[:, 1] != hi)))
        checked += c
        Pm.free(); A.free(); B.free()
    return checked, bad
