"""TEST INFRASTRUCTURE. Regenerates tests/golden/*.json|*.npz from the UNMODIFIED reference compiled in oracle/_ref
(only possible where /root/reference is mounted). Run:  python -m oracle.make_golden

Fixtures (all under the deterministic SplitMix64 tape of oracle/ref_shim.cpp):
  kat.json          known answers of the primitives on a synthetic key and on keygen(1)
  keys_seed1.npz    keygen(tape state 1): canon_tag, H_digest, prf_k, lpn_s, powg_B (H is pinned by H_digest)
  enc_seed1000.npz  enc_value(42) under tape state 1000, full ciphertext
  enc_seed2000.npz  enc_value(2^64-1) under tape state 2000, full ciphertext
  mul_seed3000.npz  ct_mul of the two under tape state 3000: layers, edge fields, SHA-256 of every sigma row, decrypt
  chain.json        digests of a deeper ct_mul / ct_add / ct_sub chain
  bounty2/*.ct      copied from the reference repository (its own golden file for ct_add / combine_ciphers)
"""
import hashlib
import json
import os

import numpy as np

from . import ref

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def hx(a):
    if isinstance(a, (list, tuple)):
        return [f"{int(x):016x}" for x in a]
    return [f"{int(x):016x}" for x in np.asarray(a).ravel()]


def ct_digest(d):
    h = hashlib.sha256()
    for k in ("rule", "ztag", "nlo", "nhi", "pa", "pb", "lid", "idx", "ch", "w", "sigma"):
        h.update(np.ascontiguousarray(d[k]).tobytes())
    return h.hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    L = ref.lib()
    kat = {}
    hd = np.arange(32, dtype=np.uint8)
    lpn_s = np.arange(1, 65, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    k = ref.Keys.from_raw(0x0123456789ABCDEF, hd, None, None, [1, 2, 3, 4], lpn_s)
    seed = (0x1111, 0x2222, 0x3333)
    doms = ["pvac.prf.r.1", "pvac.prf.r.2", "pvac.prf.r.3", "pvac.prf.noise.1", "pvac.prf.noise.2", "pvac.prf.noise.3", "pvac.dom.toeplitz"]
    kat["sha256_abc"] = ref.sha256(b"abc").hex()
    kat["fnv1a"] = {d: f"{L.ref_fnv1a(d.encode()):016x}" for d in doms}
    key = bytes.fromhex("603deb1015ca71be2b73aef0857d77811f352c073b6108d72d9810a30914dff4")
    kat["aes_key"] = key.hex()
    kat["aes_ctr0_words"] = hx(ref.aes_ctr_words(key, 0, 6))
    kat["aes_ctr_wrap_words"] = hx(ref.aes_ctr_words(key, 2**64 - 2, 8))
    kk, nn = k.derive_aes_key(*seed, "pvac.prf.r.1")
    kat["derive_key_r1"] = kk.hex()
    kat["derive_nonce_r1"] = f"{nn:016x}"
    y = k.lpn_make_ybits(*seed, "pvac.prf.r.1")
    kat["ybits_r1_sha256"] = hashlib.sha256(y.tobytes()).hexdigest()
    kat["ybits_r1_first4"] = hx(y[:4])
    kat["ybits_r1_popcount"] = int(sum(bin(int(v)).count("1") for v in y))
    kat["prf_R_core"] = {d: hx(k.prf_R_core(*seed, d)) for d in doms[:6]}
    kat["prf_R"] = hx(k.prf_R(*seed))
    kat["prf_R_noise"] = hx(k.prf_R_noise(*seed))
    kat["prf_noise_delta"] = {f"{g},{kd}": hx(k.prf_noise_delta(*seed, g, kd)) for g in range(4) for kd in range(2)}
    kat["fp_inv_prf_R"] = hx(ref.fp_inv(k.prf_R(*seed)))
    kat["ztag"] = f"{L.ref_prg_layer_ztag(0x0123456789ABCDEF, 0x2222, 0x3333):016x}"
    words = [0x0123456789ABCDEF, 0x1111, 0x2222, 0x3333, 5, 1, 0x4444]
    kat["choose_x"] = [int(v) for v in ref.prg_choose_k(128, 16384, "pvac.dom.x_seed", words)]
    kat["choose_noise"] = [int(v) for v in ref.prg_choose_k(128, 8192, "pvac.dom.noise", words)]
    kat["plan_noise"] = [list(k.plan_noise(d)) for d in range(4)]
    kat["buckets"] = {str(n): int(L.ref_unordered_buckets(n)) for n in (1, 2, 13, 14, 1521, 1560, 1600, 3160, 48080, 865000, 1444804)}
    rng = np.random.default_rng(7)
    fp_cases = []
    P = (1 << 127) - 1
    for i in range(64):
        a = int.from_bytes(rng.bytes(16), "little") % P
        b = int.from_bytes(rng.bytes(16), "little") % P
        if i == 0: a, b = P - 1, P - 1
        if i == 1: a, b = 0, P - 1
        if i == 2: a, b = 1, 1
        A = [a & (2**64 - 1), a >> 64]
        Bv = [b & (2**64 - 1), b >> 64]
        fp_cases.append(dict(a=hx(A), b=hx(Bv), add=hx(ref.fp_add(A, Bv)), sub=hx(ref.fp_sub(A, Bv)), mul=hx(ref.fp_mul(A, Bv)), neg=hx(ref.fp_neg(A)),
                             inv=hx(ref.fp_inv(A)) if a else None))
    kat["fp"] = fp_cases

    # keygen(1)
    K = ref.Keys.keygen(1)
    e = K.export()
    np.savez_compressed(os.path.join(OUT, "keys_seed1.npz"), canon_tag=np.uint64(e["canon_tag"]), H_digest=e["H_digest"], prf_k=e["prf_k"], lpn_s=e["lpn_s"], powg=e["powg"])
    kat["keygen1_H_col_sha256"] = {str(c): hashlib.sha256(e["H"][c].tobytes()).hexdigest() for c in (0, 1, 777, 16383)}
    sg = K.sigma_from_H(0x1111, 0x2222, 0x3333, 5, 1, 0x4444)
    kat["keygen1_sigma_sha256"] = hashlib.sha256(sg.tobytes()).hexdigest()

    ca = K.enc_value(1000, 42)
    kat["enc1000_draws"] = int(L.ref_tape_draws())
    cb = K.enc_value(2000, 2**64 - 1)
    kat["enc2000_draws"] = int(L.ref_tape_draws())
    da, db = ref.ct_export(ca), ref.ct_export(cb)
    np.savez_compressed(os.path.join(OUT, "enc_seed1000.npz"), **da)
    np.savez_compressed(os.path.join(OUT, "enc_seed2000.npz"), **db)
    cp = K.ct_mul(3000, ca, cb)
    kat["mul3000_draws"] = int(L.ref_tape_draws())
    dp = ref.ct_export(cp)
    sig_hash = np.frombuffer(b"".join(hashlib.sha256(r.tobytes()).digest() for r in dp["sigma"]), np.uint8).reshape(-1, 32)
    dp_small = {k_: v for k_, v in dp.items() if k_ != "sigma"}
    np.savez_compressed(os.path.join(OUT, "mul_seed3000.npz"), sigma_sha256=sig_hash, dec=K.dec_value(cp), **dp_small)
    kat["dec_enc1000"] = hx(K.dec_value(ca))
    kat["dec_mul3000"] = hx(K.dec_value(cp))

    chain = {}
    cs = K.ct_add(ca, cb)
    cd = K.ct_sub(ca, cb)
    chain["add"] = ct_digest(ref.ct_export(cs))
    chain["sub"] = ct_digest(ref.ct_export(cd))
    chain["dec_add"] = hx(K.dec_value(cs))
    chain["dec_sub"] = hx(K.dec_value(cd))
    p2 = K.ct_mul(4000, cp, ca)           # (a*b)*a : 8 x 2 layers
    chain["mul_pa"] = ct_digest(ref.ct_export(p2))
    chain["mul_pa_counts"] = [len(ref.ct_export(p2, False)["rule"]), len(ref.ct_export(p2, False)["lid"])]
    p3 = K.ct_mul(5000, cs, cp)           # (a+b)*(a*b)
    chain["mul_sp"] = ct_digest(ref.ct_export(p3))
    chain["dec_mul_sp"] = hx(K.dec_value(p3))
    sq = K.ct_mul(6000, cp, cp)           # test_depth step shape: product squared
    dsq = ref.ct_export(sq)
    chain["sq"] = ct_digest(dsq)
    chain["sq_counts"] = [len(dsq["rule"]), len(dsq["lid"])]
    chain["dec_sq"] = hx(K.dec_value(sq))
    sc = K.ct_scale(ca, [12345, 0])
    chain["scale"] = ct_digest(ref.ct_export(sc))
    with open(os.path.join(OUT, "chain.json"), "w") as f:
        json.dump(chain, f, indent=1)
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("golden fixtures written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
