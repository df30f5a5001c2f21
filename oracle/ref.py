"""TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/_ref/libpvac_ref.so, the UNMODIFIED reference
headers compiled behind a deterministic word tape (see oracle/ref_shim.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
Ciphertexts cross the boundary as dicts of numpy arrays (struct-of-arrays):
  rule u8[nL], ztag/nlo/nhi u64[nL], pa/pb u32[nL], lid u32[nE], idx u16[nE], ch u8[nE],
  w u64[nE,2] (lo,hi), sigma u64[nE,128].
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpvac_ref.so")

M_WORDS = 128
N_COLS = 16384
B = 337
LPN_WORDS = 64

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, u16, u8, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8, C.c_int
    P = C.POINTER
    sig = {
        "ref_init": (None, []),
        "ref_seed": (None, [u64]),
        "ref_tape_draws": (u64, []),
        "ref_tape_word": (u64, []),
        "ref_set_tape": (None, [i32, P(u8)]),
        "ref_set_tape_lane": (None, [u32]),
        "ref_set_tape_words": (None, [P(u64), u64]),
        "ref_keys_omega": (None, [vp, P(u64)]),
        "ref_keys_save": (i32, [vp, C.c_char_p, C.c_char_p]),
        "ref_item_stream_state": (u64, [u64, u64]),
        "ref_keygen": (vp, [u64]),
        "ref_keys_from_raw": (vp, [u64, P(u8), P(u64), P(u64), P(u64), P(u64)]),
        "ref_keys_free": (None, [vp]),
        "ref_keys_set_lpn_t": (None, [vp, i32]),
        "ref_keys_export": (None, [vp, P(u64), P(u8), P(u64), P(u64), P(u64), P(u64)]),
        "ref_fp_mul": (None, [P(u64), P(u64), P(u64)]),
        "ref_fp_add": (None, [P(u64), P(u64), P(u64)]),
        "ref_fp_sub": (None, [P(u64), P(u64), P(u64)]),
        "ref_fp_neg": (None, [P(u64), P(u64)]),
        "ref_fp_inv": (None, [P(u64), P(u64)]),
        "ref_fp_from_words": (None, [u64, u64, P(u64)]),
        "ref_hash_to_fp_nonzero": (None, [u64, u64, P(u64)]),
        "ref_sha256": (None, [P(u8), C.c_size_t, P(u8)]),
        "ref_fnv1a": (u64, [C.c_char_p]),
        "ref_aes_ctr_words": (None, [P(u8), u64, P(u64), C.c_size_t]),
        "ref_aes_ctr_draws": (None, [P(u8), u64, P(u64), P(u64), C.c_size_t]),
        "ref_params_default": (None, [P(C.c_double)]),
        "ref_derive_aes_key": (None, [vp, u64, u64, u64, C.c_char_p, P(u8), P(u64)]),
        "ref_lpn_make_ybits": (None, [vp, u64, u64, u64, C.c_char_p, P(u64)]),
        "ref_toep_127": (None, [P(u64), C.c_size_t, P(u64), C.c_size_t, P(u64)]),
        "ref_prf_R_core": (None, [vp, u64, u64, u64, C.c_char_p, P(u64)]),
        "ref_prf_R": (None, [vp, u64, u64, u64, P(u64)]),
        "ref_prf_R_noise": (None, [vp, u64, u64, u64, P(u64)]),
        "ref_prf_noise_delta": (None, [vp, u64, u64, u64, u32, u8, P(u64)]),
        "ref_prg_layer_ztag": (u64, [u64, u64, u64]),
        "ref_prg_choose_k": (None, [i32, i32, C.c_char_p, P(u64), C.c_size_t, P(C.c_int32)]),
        "ref_sigma_from_H": (None, [vp, u64, u64, u64, u16, u8, u64, P(u64)]),
        "ref_plan_noise": (None, [vp, i32, P(i32), P(i32)]),
        "ref_unordered_buckets": (u64, [u64]),
        "ref_enc_value": (vp, [vp, u64, u64]),
        "ref_enc_value_explicit": (vp, [vp, u64, u64, i32]),
        "ref_enc_fp_depth": (vp, [vp, u64, P(u64), i32]),
        "ref_ct_add": (vp, [vp, vp, vp]),
        "ref_ct_sub": (vp, [vp, vp, vp]),
        "ref_ct_scale": (vp, [vp, vp, P(u64)]),
        "ref_enc_value_depth": (vp, [vp, u64, u64, i32]),
        "ref_enc_zero_depth": (vp, [vp, u64, i32]),
        "ref_ct_neg": (vp, [vp, vp]),
        "ref_ct_div_const": (vp, [vp, vp, P(u64)]),
        "ref_commit_ct": (None, [vp, vp, P(u8)]),
        "ref_enc_text": (i32, [vp, u64, P(u8), u64, P(vp), i32]),
        "ref_dec_text": (i32, [vp, P(vp), i32, P(u8), i32]),
        "ref_ubk_perm": (None, [vp, P(C.c_int32)]),
        "ref_ubk_apply": (vp, [vp, vp]),
        "ref_sigma_density": (C.c_double, [vp, vp]),
        "ref_ct_recrypt": (vp, [vp, u64, vp, P(vp), i32]),
        "ref_compact_edges": (vp, [vp, vp]),
        "ref_ct_mul": (vp, [vp, u64, vp, vp]),
        "ref_dec_value": (None, [vp, vp, P(u64)]),
        "ref_ct_free": (None, [vp]),
        "ref_ct_counts": (None, [vp, P(u32), P(u32)]),
        "ref_ct_export": (None, [vp, P(u8), P(u64), P(u64), P(u64), P(u32), P(u32), P(u32), P(u16), P(u8), P(u64), P(u64)]),
        "ref_ct_import": (vp, [u32, u32, u32, P(u8), P(u64), P(u64), P(u64), P(u32), P(u32), P(u32), P(u16), P(u8), P(u64), P(u64)]),
        "ref_bench": (C.c_double, [vp, i32, i32, i32, u64, P(u64)]),
        "ref_bench_chain": (None, [vp, i32, i32, u64, P(C.c_double), P(C.c_double), P(u64)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    L.ref_init()
    _lib = L
    return L


_tape_words_keep = None


def set_tape(kind, key=None, lane=0, words=None):
    """tape kind behind the reference's getrandom(): 0 SplitMix64 (default), 1 ChaCha20 under key with `lane`, 2 explicit words"""
    global _tape_words_keep
    k = np.frombuffer(bytes(key), np.uint8).copy() if key is not None else None
    lib().ref_set_tape(kind, _p(k, C.c_uint8) if k is not None else None)
    lib().ref_set_tape_lane(lane)
    if words is not None:
        _tape_words_keep = np.ascontiguousarray(words, np.uint64)
        lib().ref_set_tape_words(_p(_tape_words_keep, C.c_uint64), len(_tape_words_keep))


def _fp2(fn, *vals):
    out = np.zeros(2, np.uint64)
    args = [_p(np.asarray(v, np.uint64), C.c_uint64) for v in vals]
    fn(*args, _p(out, C.c_uint64))
    return out


def fp_mul(a, b):
    return _fp2(lib().ref_fp_mul, a, b)


def fp_add(a, b):
    return _fp2(lib().ref_fp_add, a, b)


def fp_sub(a, b):
    return _fp2(lib().ref_fp_sub, a, b)


def fp_neg(a):
    return _fp2(lib().ref_fp_neg, a)


def fp_inv(a):
    return _fp2(lib().ref_fp_inv, a)


def sha256(data: bytes) -> bytes:
    buf = np.frombuffer(data, np.uint8).copy() if data else np.zeros(1, np.uint8)
    out = np.zeros(32, np.uint8)
    lib().ref_sha256(_p(buf, C.c_uint8), len(data), _p(out, C.c_uint8))
    return out.tobytes()


def aes_ctr_words(key: bytes, nonce: int, n: int):
    k = np.frombuffer(key, np.uint8).copy()
    out = np.zeros(n, np.uint64)
    lib().ref_aes_ctr_words(_p(k, C.c_uint8), nonce, _p(out, C.c_uint64), n)
    return out


def params_default():
    """the reference's default Params (core/types.hpp:36-70) in declaration order"""
    out = (C.c_double * 17)()
    lib().ref_params_default(out)
    return [float(x) for x in out]


def aes_ctr_draws(key: bytes, nonce: int, moduli):
    """one stream; moduli[i] == 0 -> next_u64(), else bounded(moduli[i]) (crypto/lpn.hpp:108-148)"""
    k = np.frombuffer(key, np.uint8).copy()
    m = np.ascontiguousarray(moduli, np.uint64)
    out = np.zeros(len(m), np.uint64)
    lib().ref_aes_ctr_draws(_p(k, C.c_uint8), nonce, _p(m, C.c_uint64), _p(out, C.c_uint64), len(m))
    return out


def prg_choose_k(k, N, label: str, words):
    w = np.asarray(words, np.uint64)
    out = np.zeros(k, np.int32)
    lib().ref_prg_choose_k(k, N, label.encode(), _p(w, C.c_uint64), len(w), _p(out, C.c_int32))
    return out


class Keys:
    """Owns a reference PubKey/SecKey pair."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def keygen(cls, tape_state: int):
        return cls(lib().ref_keygen(tape_state))

    @classmethod
    def from_raw(cls, canon_tag, h_digest, H, powg, prf_k, lpn_s):
        hd = np.asarray(h_digest, np.uint8)
        pk = np.asarray(prf_k, np.uint64)
        ls = np.asarray(lpn_s, np.uint64)
        Hp = _p(np.ascontiguousarray(H, np.uint64), C.c_uint64) if H is not None else None
        gp = _p(np.ascontiguousarray(powg, np.uint64), C.c_uint64) if powg is not None else None
        return cls(lib().ref_keys_from_raw(int(canon_tag), _p(hd, C.c_uint8), Hp, gp, _p(pk, C.c_uint64), _p(ls, C.c_uint64)))

    def set_lpn_t(self, t):
        lib().ref_keys_set_lpn_t(self.h, t)

    def omega_B(self):
        o = np.zeros(2, np.uint64)
        lib().ref_keys_omega(self.h, _p(o, C.c_uint64))
        return o

    def save(self, pk_path=None, sk_path=None):
        """the reference's key files (layout of savePk / saveSk, tests/bounty2_test.cpp:145-192) written from the reference's own objects"""
        rc = lib().ref_keys_save(self.h, pk_path.encode() if pk_path else None, sk_path.encode() if sk_path else None)
        assert rc == 0

    def export(self, with_H=True):
        ct = C.c_uint64()
        hd = np.zeros(32, np.uint8)
        H = np.zeros((N_COLS, M_WORDS), np.uint64) if with_H else None
        powg = np.zeros((B, 2), np.uint64)
        prf_k = np.zeros(4, np.uint64)
        lpn_s = np.zeros(LPN_WORDS, np.uint64)
        lib().ref_keys_export(self.h, C.byref(ct), _p(hd, C.c_uint8), _p(H, C.c_uint64) if with_H else None,
                              _p(powg, C.c_uint64), _p(prf_k, C.c_uint64), _p(lpn_s, C.c_uint64))
        return dict(canon_tag=int(ct.value), H_digest=hd, H=H, powg=powg, prf_k=prf_k, lpn_s=lpn_s)

    # ---- primitives bound to the key
    def derive_aes_key(self, ztag, nlo, nhi, dom: str):
        key = np.zeros(32, np.uint8)
        nonce = C.c_uint64()
        lib().ref_derive_aes_key(self.h, ztag, nlo, nhi, dom.encode(), _p(key, C.c_uint8), C.byref(nonce))
        return key.tobytes(), int(nonce.value)

    def lpn_make_ybits(self, ztag, nlo, nhi, dom: str, lpn_t=16384):
        y = np.zeros((lpn_t + 63) // 64, np.uint64)
        lib().ref_lpn_make_ybits(self.h, ztag, nlo, nhi, dom.encode(), _p(y, C.c_uint64))
        return y

    def prf_R_core(self, ztag, nlo, nhi, dom: str):
        o = np.zeros(2, np.uint64)
        lib().ref_prf_R_core(self.h, ztag, nlo, nhi, dom.encode(), _p(o, C.c_uint64))
        return o

    def prf_R(self, ztag, nlo, nhi):
        o = np.zeros(2, np.uint64)
        lib().ref_prf_R(self.h, ztag, nlo, nhi, _p(o, C.c_uint64))
        return o

    def prf_R_noise(self, ztag, nlo, nhi):
        o = np.zeros(2, np.uint64)
        lib().ref_prf_R_noise(self.h, ztag, nlo, nhi, _p(o, C.c_uint64))
        return o

    def prf_noise_delta(self, ztag, nlo, nhi, gid, kind):
        o = np.zeros(2, np.uint64)
        lib().ref_prf_noise_delta(self.h, ztag, nlo, nhi, gid, kind, _p(o, C.c_uint64))
        return o

    def sigma_from_H(self, ztag, nlo, nhi, idx, ch, salt):
        o = np.zeros(M_WORDS, np.uint64)
        lib().ref_sigma_from_H(self.h, ztag, nlo, nhi, idx, ch, salt, _p(o, C.c_uint64))
        return o

    def plan_noise(self, depth):
        a, b = C.c_int(), C.c_int()
        lib().ref_plan_noise(self.h, depth, C.byref(a), C.byref(b))
        return a.value, b.value

    # ---- ciphertext ops (handles are reference Cipher*)
    def enc_value(self, tape_state, v):
        return lib().ref_enc_value(self.h, tape_state, v)

    def enc_value_explicit(self, tape_state, v, second_first):
        return lib().ref_enc_value_explicit(self.h, tape_state, v, int(second_first))

    def enc_value_depth(self, tape_state, v, depth):
        return lib().ref_enc_value_depth(self.h, tape_state, v, depth)

    def enc_zero_depth(self, tape_state, depth):
        return lib().ref_enc_zero_depth(self.h, tape_state, depth)

    def ct_neg(self, a):
        return lib().ref_ct_neg(self.h, a)

    def ct_div_const(self, a, k):
        kk = np.asarray(k, np.uint64)
        return lib().ref_ct_div_const(self.h, a, _p(kk, C.c_uint64))

    def enc_fp_depth(self, tape_state, v, depth=0):
        vv = np.asarray(v, np.uint64)
        return lib().ref_enc_fp_depth(self.h, tape_state, _p(vv, C.c_uint64), depth)

    def ct_add(self, a, b):
        return lib().ref_ct_add(self.h, a, b)

    def ct_sub(self, a, b):
        return lib().ref_ct_sub(self.h, a, b)

    def ct_scale(self, a, s):
        ss = np.asarray(s, np.uint64)
        return lib().ref_ct_scale(self.h, a, _p(ss, C.c_uint64))

    def commit_ct(self, c):
        o = np.zeros(32, np.uint8)
        lib().ref_commit_ct(self.h, c, _p(o, C.c_uint8))
        return o.tobytes()

    def enc_text(self, tape_state, msg: bytes):
        cap = 2 + len(msg) // 15 + 1
        arr = (C.c_void_p * cap)()
        m = np.frombuffer(msg or b"\0", np.uint8).copy()
        n = lib().ref_enc_text(self.h, tape_state, _p(m, C.c_uint8), len(msg), arr, cap)
        return [arr[i] for i in range(n)]

    def dec_text(self, cts):
        arr = (C.c_void_p * len(cts))(*cts)
        buf = np.zeros(15 * len(cts) + 16, np.uint8)
        n = lib().ref_dec_text(self.h, arr, len(cts), _p(buf, C.c_uint8), len(buf))
        return buf[:n].tobytes()

    def ubk_perm(self):
        o = np.zeros(8192, np.int32)
        lib().ref_ubk_perm(self.h, _p(o, C.c_int32))
        return o

    def ubk_apply(self, c):
        return lib().ref_ubk_apply(self.h, c)

    def sigma_density(self, c):
        return float(lib().ref_sigma_density(self.h, c))

    def ct_recrypt(self, tape_state, c, pool):
        arr = (C.c_void_p * max(len(pool), 1))(*pool)
        return lib().ref_ct_recrypt(self.h, tape_state, c, arr, len(pool))

    def compact_edges(self, a):
        return lib().ref_compact_edges(self.h, a)

    def ct_mul(self, tape_state, a, b):
        return lib().ref_ct_mul(self.h, tape_state, a, b)

    def dec_value(self, c):
        o = np.zeros(2, np.uint64)
        lib().ref_dec_value(self.h, c, _p(o, C.c_uint64))
        return o

    def bench_chain(self, threads, steps, seed=1):
        """test_depth's chain c <- c*c on every thread: -> (mul seconds per step, dec seconds per step, edges per step), slowest thread"""
        m, d, e = np.zeros(steps, np.float64), np.zeros(steps, np.float64), np.zeros(steps, np.uint64)
        lib().ref_bench_chain(self.h, threads, steps, seed, _p(m, C.c_double), _p(d, C.c_double), _p(e, C.c_uint64))
        return m, d, e

    def bench(self, op, threads, iters, seed=1):
        done = C.c_uint64()
        secs = lib().ref_bench(self.h, op, threads, iters, seed, C.byref(done))
        return secs, int(done.value)


def ct_free(c):
    lib().ref_ct_free(c)


def ct_export(c, with_sigma=True):
    nL, nE = C.c_uint32(), C.c_uint32()
    lib().ref_ct_counts(c, C.byref(nL), C.byref(nE))
    nL, nE = nL.value, nE.value
    d = dict(
        rule=np.zeros(nL, np.uint8), ztag=np.zeros(nL, np.uint64), nlo=np.zeros(nL, np.uint64), nhi=np.zeros(nL, np.uint64),
        pa=np.zeros(nL, np.uint32), pb=np.zeros(nL, np.uint32),
        lid=np.zeros(nE, np.uint32), idx=np.zeros(nE, np.uint16), ch=np.zeros(nE, np.uint8),
        w=np.zeros((nE, 2), np.uint64), sigma=np.zeros((nE, M_WORDS), np.uint64) if with_sigma else None,
    )
    lib().ref_ct_export(c, _p(d["rule"], C.c_uint8), _p(d["ztag"], C.c_uint64), _p(d["nlo"], C.c_uint64), _p(d["nhi"], C.c_uint64),
                        _p(d["pa"], C.c_uint32), _p(d["pb"], C.c_uint32), _p(d["lid"], C.c_uint32), _p(d["idx"], C.c_uint16),
                        _p(d["ch"], C.c_uint8), _p(d["w"], C.c_uint64), _p(d["sigma"], C.c_uint64) if with_sigma else None)
    return d


def ct_import(d):
    nL, nE = len(d["rule"]), len(d["lid"])
    sg = d.get("sigma")
    return lib().ref_ct_import(
        nL, nE, M_WORDS * 64, _p(d["rule"], C.c_uint8), _p(d["ztag"], C.c_uint64), _p(d["nlo"], C.c_uint64), _p(d["nhi"], C.c_uint64),
        _p(d["pa"], C.c_uint32), _p(d["pb"], C.c_uint32), _p(d["lid"], C.c_uint32), _p(d["idx"], C.c_uint16), _p(d["ch"], C.c_uint8),
        _p(np.ascontiguousarray(d["w"]), C.c_uint64), _p(np.ascontiguousarray(sg), C.c_uint64) if sg is not None else None)
