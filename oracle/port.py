"""TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/libpvac_oracle.so (oracle/pvac_oracle.c), the CPU
restatement of the reference's hot path. Same Python surface as oracle/ref.py so tests can run the same
checks against both. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpvac_oracle.so")

M_WORDS = 128
P127 = (1 << 127) - 1
N_COLS = 16384
B = 337
LPN_WORDS = 64

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libpvac_oracle.so"])


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, u16, u8, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8, C.c_int
    P = C.POINTER
    sig = {
        "orc_item_stream_state": (u64, [u64, u64]),
        "orc_tape_word": (u64, [P(u64)]),
        "orc_set_tape": (None, [i32, P(u8)]),
        "orc_set_tape_lane": (None, [u32]),
        "orc_set_tape_words": (None, [P(u64), u64]),
        "orc_set_prf_patch": (None, [u64, u64]),
        "orc_chacha20_block": (None, [P(u32), P(u32), P(u32)]),
        "orc_keygen": (vp, [u64]),
        "orc_keys_from_raw": (vp, [u64, P(u8), P(u64), P(u64), P(u64), P(u64)]),
        "orc_keys_free": (None, [vp]),
        "orc_keys_set_lpn_rows": (None, [vp, i32]),
        "orc_keys_export": (None, [vp, P(u64), P(u8), P(u64), P(u64), P(u64), P(u64)]),
        "orc_fp_mul": (None, [P(u64), P(u64), P(u64)]),
        "orc_fp_add": (None, [P(u64), P(u64), P(u64)]),
        "orc_fp_sub": (None, [P(u64), P(u64), P(u64)]),
        "orc_fp_neg": (None, [P(u64), P(u64)]),
        "orc_fp_inv": (None, [P(u64), P(u64)]),
        "orc_fp_from_words": (None, [u64, u64, P(u64)]),
        "orc_hash_to_fp_nonzero": (None, [u64, u64, P(u64)]),
        "orc_sha256": (None, [P(u8), C.c_size_t, P(u8)]),
        "orc_fnv1a": (u64, [C.c_char_p]),
        "orc_aes_ctr_words": (None, [P(u8), u64, P(u64), C.c_size_t]),
        "orc_aes_ctr_draws": (None, [P(u8), u64, P(u64), P(u64), C.c_size_t]),
        "orc_derive_aes_key": (None, [vp, u64, u64, u64, C.c_char_p, P(u8), P(u64)]),
        "orc_lpn_make_ybits": (None, [vp, u64, u64, u64, C.c_char_p, i32, P(u64)]),
        "orc_toep_127": (None, [P(u64), C.c_size_t, P(u64), C.c_size_t, P(u64)]),
        "orc_prf_R_core": (None, [vp, u64, u64, u64, C.c_char_p, P(u64)]),
        "orc_prf_R": (None, [vp, u64, u64, u64, P(u64)]),
        "orc_prf_R_noise": (None, [vp, u64, u64, u64, P(u64)]),
        "orc_prf_noise_delta": (None, [vp, u64, u64, u64, u32, u8, P(u64)]),
        "orc_prg_layer_ztag": (u64, [u64, u64, u64]),
        "orc_prg_choose_k": (None, [i32, i32, C.c_char_p, P(u64), C.c_size_t, P(C.c_int32)]),
        "orc_sigma_from_H": (None, [vp, u64, u64, u64, u16, u8, u64, P(u64)]),
        "orc_plan_noise": (None, [i32, P(i32), P(i32)]),
        "orc_next_bkt": (u64, [u64]),
        "orc_unordered_buckets_real": (u64, [u64]),
        "orc_enc_value": (vp, [vp, u64, u64, P(u64)]),
        "orc_enc_value_depth": (vp, [vp, u64, u64, i32, P(u64)]),
        "orc_enc_zero_depth": (vp, [vp, u64, i32, P(u64)]),
        "orc_enc_fp_depth": (vp, [vp, u64, P(u64), i32, P(u64)]),
        "orc_ct_add": (vp, [vp, vp]),
        "orc_ct_sub": (vp, [vp, vp]),
        "orc_ct_scale": (vp, [vp, P(u64)]),
        "orc_commit_ct": (None, [vp, vp, P(u8)]),
        "orc_enc_text": (i32, [vp, u64, P(u8), u64, P(vp), i32, P(u64)]),
        "orc_ubk_perm": (None, [u64, P(C.c_int32)]),
        "orc_ubk_apply": (vp, [vp, vp]),
        "orc_sigma_density": (C.c_double, [vp]),
        "orc_ct_recrypt": (vp, [vp, u64, vp, P(vp), i32, P(u64)]),
        "orc_compact_edges": (vp, [vp]),
        "orc_ct_mul": (vp, [vp, u64, vp, vp, P(u64)]),
        "orc_dec_value": (i32, [vp, vp, P(u64)]),
        "orc_ct_free": (None, [vp]),
        "orc_ct_counts": (None, [vp, P(u32), P(u32)]),
        "orc_ct_export": (None, [vp, P(u8), P(u64), P(u64), P(u64), P(u32), P(u32), P(u32), P(u16), P(u8), P(u64), P(u64)]),
        "orc_ct_import": (vp, [u32, u32, P(u8), P(u64), P(u64), P(u64), P(u32), P(u32), P(u32), P(u16), P(u8), P(u64), P(u64)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


_last_draws = 0
_tape_words_keep = None


def set_tape(kind, key=None, lane=0, words=None):
    """tape kind of the oracle (process-wide): 0 SplitMix64 (default), 1 ChaCha20 under key (32 bytes) with `lane`, 2 explicit words.
    The tape_state argument of enc_* / ct_mul / ... is the SplitMix state resp. the ChaCha20 stream id."""
    global _tape_words_keep
    k = np.frombuffer(bytes(key), np.uint8).copy() if key is not None else None
    lib().orc_set_tape(kind, _p(k, C.c_uint8) if k is not None else None)
    lib().orc_set_tape_lane(lane)
    if words is not None:
        _tape_words_keep = np.ascontiguousarray(words, np.uint64)
        lib().orc_set_tape_words(_p(_tape_words_keep, C.c_uint64), len(_tape_words_keep))


def set_prf_patch(word=2**64 - 1, or_mask=0):
    """test hook: keystream word `word` of every LPN stream is OR-ed with or_mask (word = 2^64-1: off)"""
    lib().orc_set_prf_patch(word, or_mask)


def chacha20_block(key_words, c4):
    k, c, o = np.asarray(key_words, np.uint32), np.asarray(c4, np.uint32), np.zeros(16, np.uint32)
    lib().orc_chacha20_block(_p(k, C.c_uint32), _p(c, C.c_uint32), _p(o, C.c_uint32))
    return o


def tape_draws():
    return _last_draws


def item_stream_state(batch_seed, item):
    return int(lib().orc_item_stream_state(batch_seed, item))


def _fp2(fn, *vals):
    out = np.zeros(2, np.uint64)
    args = [_p(np.asarray(v, np.uint64), C.c_uint64) for v in vals]
    fn(*args, _p(out, C.c_uint64))
    return out


def fp_mul(a, b):
    return _fp2(lib().orc_fp_mul, a, b)


def fp_add(a, b):
    return _fp2(lib().orc_fp_add, a, b)


def fp_sub(a, b):
    return _fp2(lib().orc_fp_sub, a, b)


def fp_neg(a):
    return _fp2(lib().orc_fp_neg, a)


def fp_inv(a):
    return _fp2(lib().orc_fp_inv, a)


def sha256(data: bytes) -> bytes:
    buf = np.frombuffer(data, np.uint8).copy() if data else np.zeros(1, np.uint8)
    out = np.zeros(32, np.uint8)
    lib().orc_sha256(_p(buf, C.c_uint8), len(data), _p(out, C.c_uint8))
    return out.tobytes()


def aes_ctr_words(key: bytes, nonce: int, n: int):
    k = np.frombuffer(key, np.uint8).copy()
    out = np.zeros(n, np.uint64)
    lib().orc_aes_ctr_words(_p(k, C.c_uint8), nonce, _p(out, C.c_uint64), n)
    return out


def aes_ctr_draws(key: bytes, nonce: int, moduli):
    """one stream; moduli[i] == 0 -> next_u64(), else bounded(moduli[i]) (crypto/lpn.hpp:108-148)"""
    k = np.frombuffer(key, np.uint8).copy()
    m = np.ascontiguousarray(moduli, np.uint64)
    out = np.zeros(len(m), np.uint64)
    lib().orc_aes_ctr_draws(_p(k, C.c_uint8), nonce, _p(m, C.c_uint64), _p(out, C.c_uint64), len(m))
    return out


def prg_choose_k(k, N, label: str, words):
    w = np.asarray(words, np.uint64)
    out = np.zeros(k, np.int32)
    lib().orc_prg_choose_k(k, N, label.encode(), _p(w, C.c_uint64), len(w), _p(out, C.c_int32))
    return out


def toep_127(top, y):
    t = np.asarray(top, np.uint64)
    yy = np.asarray(y, np.uint64)
    o = np.zeros(2, np.uint64)
    lib().orc_toep_127(_p(t, C.c_uint64), len(t), _p(yy, C.c_uint64), len(yy), _p(o, C.c_uint64))
    return o


class Keys:
    def __init__(self, handle):
        self.h = handle
        self.lpn_rows = 16384

    @classmethod
    def keygen(cls, tape_state: int):
        return cls(lib().orc_keygen(tape_state))

    @classmethod
    def from_raw(cls, canon_tag, h_digest, H, powg, prf_k, lpn_s):
        hd = np.asarray(h_digest, np.uint8)
        pk = np.asarray(prf_k, np.uint64)
        ls = np.asarray(lpn_s, np.uint64)
        Hp = _p(np.ascontiguousarray(H, np.uint64), C.c_uint64) if H is not None else None
        gp = _p(np.ascontiguousarray(powg, np.uint64), C.c_uint64) if powg is not None else None
        return cls(lib().orc_keys_from_raw(int(canon_tag), _p(hd, C.c_uint8), Hp, gp, _p(pk, C.c_uint64), _p(ls, C.c_uint64)))

    def set_lpn_t(self, rows):
        """rows of the LPN sample evaluated (16384 = like the reference; >=127 gives identical outputs)."""
        self.lpn_rows = rows
        lib().orc_keys_set_lpn_rows(self.h, rows)

    def export(self, with_H=True):
        ct = C.c_uint64()
        hd = np.zeros(32, np.uint8)
        H = np.zeros((N_COLS, M_WORDS), np.uint64) if with_H else None
        powg = np.zeros((B, 2), np.uint64)
        prf_k = np.zeros(4, np.uint64)
        lpn_s = np.zeros(LPN_WORDS, np.uint64)
        lib().orc_keys_export(self.h, C.byref(ct), _p(hd, C.c_uint8), _p(H, C.c_uint64) if with_H else None,
                              _p(powg, C.c_uint64), _p(prf_k, C.c_uint64), _p(lpn_s, C.c_uint64))
        return dict(canon_tag=int(ct.value), H_digest=hd, H=H, powg=powg, prf_k=prf_k, lpn_s=lpn_s)

    def derive_aes_key(self, ztag, nlo, nhi, dom: str):
        key = np.zeros(32, np.uint8)
        nonce = C.c_uint64()
        lib().orc_derive_aes_key(self.h, ztag, nlo, nhi, dom.encode(), _p(key, C.c_uint8), C.byref(nonce))
        return key.tobytes(), int(nonce.value)

    def lpn_make_ybits(self, ztag, nlo, nhi, dom: str, lpn_t=None):
        rows = self.lpn_rows if lpn_t is None else lpn_t
        y = np.zeros((rows + 63) // 64, np.uint64)
        lib().orc_lpn_make_ybits(self.h, ztag, nlo, nhi, dom.encode(), rows, _p(y, C.c_uint64))
        return y

    def prf_R_core(self, ztag, nlo, nhi, dom: str):
        o = np.zeros(2, np.uint64)
        lib().orc_prf_R_core(self.h, ztag, nlo, nhi, dom.encode(), _p(o, C.c_uint64))
        return o

    def prf_R(self, ztag, nlo, nhi):
        o = np.zeros(2, np.uint64)
        lib().orc_prf_R(self.h, ztag, nlo, nhi, _p(o, C.c_uint64))
        return o

    def prf_R_noise(self, ztag, nlo, nhi):
        o = np.zeros(2, np.uint64)
        lib().orc_prf_R_noise(self.h, ztag, nlo, nhi, _p(o, C.c_uint64))
        return o

    def prf_noise_delta(self, ztag, nlo, nhi, gid, kind):
        o = np.zeros(2, np.uint64)
        lib().orc_prf_noise_delta(self.h, ztag, nlo, nhi, gid, kind, _p(o, C.c_uint64))
        return o

    def sigma_from_H(self, ztag, nlo, nhi, idx, ch, salt):
        o = np.zeros(M_WORDS, np.uint64)
        lib().orc_sigma_from_H(self.h, ztag, nlo, nhi, idx, ch, salt, _p(o, C.c_uint64))
        return o

    def plan_noise(self, depth):
        a, b = C.c_int(), C.c_int()
        lib().orc_plan_noise(depth, C.byref(a), C.byref(b))
        return a.value, b.value

    def enc_value(self, tape_state, v):
        global _last_draws
        d = C.c_uint64()
        h = lib().orc_enc_value(self.h, tape_state, v, C.byref(d))
        _last_draws = int(d.value)
        return h

    def enc_value_depth(self, tape_state, v, depth):
        global _last_draws
        d = C.c_uint64()
        c = lib().orc_enc_value_depth(self.h, tape_state, v, depth, C.byref(d))
        _last_draws = int(d.value)
        return c

    def enc_zero_depth(self, tape_state, depth):
        global _last_draws
        d = C.c_uint64()
        c = lib().orc_enc_zero_depth(self.h, tape_state, depth, C.byref(d))
        _last_draws = int(d.value)
        return c

    def ct_neg(self, a):
        return self.ct_scale(a, [P127 - 1 & (2**64 - 1), (P127 - 1) >> 64])

    def ct_div_const(self, a, k):
        kv = int(k[0]) | (int(k[1]) << 64)
        inv = pow(kv, P127 - 2, P127)
        return self.ct_scale(a, [inv & (2**64 - 1), inv >> 64])

    def enc_fp_depth(self, tape_state, v, depth=0):
        global _last_draws
        d = C.c_uint64()
        vv = np.asarray(v, np.uint64)
        h = lib().orc_enc_fp_depth(self.h, tape_state, _p(vv, C.c_uint64), depth, C.byref(d))
        _last_draws = int(d.value)
        return h

    def ct_add(self, a, b):
        return lib().orc_ct_add(a, b)

    def ct_sub(self, a, b):
        return lib().orc_ct_sub(a, b)

    def ct_scale(self, a, s):
        ss = np.asarray(s, np.uint64)
        return lib().orc_ct_scale(a, _p(ss, C.c_uint64))

    def commit_ct(self, c):
        o = np.zeros(32, np.uint8)
        lib().orc_commit_ct(self.h, c, _p(o, C.c_uint8))
        return o.tobytes()

    def enc_text(self, tape_state, msg: bytes):
        global _last_draws
        cap = 2 + len(msg) // 15 + 1
        arr = (C.c_void_p * cap)()
        m = np.frombuffer(msg or b"\0", np.uint8).copy()
        d = C.c_uint64()
        n = lib().orc_enc_text(self.h, tape_state, _p(m, C.c_uint8), len(msg), arr, cap, C.byref(d))
        _last_draws = int(d.value)
        return [arr[i] for i in range(n)]

    def ubk_perm(self):
        o = np.zeros(8192, np.int32)
        lib().orc_ubk_perm(self.export(with_H=False)["canon_tag"], _p(o, C.c_int32))
        return o

    def ubk_apply(self, c):
        return lib().orc_ubk_apply(self.h, c)

    def sigma_density(self, c):
        return float(lib().orc_sigma_density(c))

    def ct_recrypt(self, tape_state, c, pool):
        global _last_draws
        arr = (C.c_void_p * max(len(pool), 1))(*pool)
        d = C.c_uint64()
        r = lib().orc_ct_recrypt(self.h, tape_state, c, arr, len(pool), C.byref(d))
        _last_draws = int(d.value)
        return r

    def compact_edges(self, a):
        return lib().orc_compact_edges(a)

    def ct_mul(self, tape_state, a, b):
        global _last_draws
        d = C.c_uint64()
        h = lib().orc_ct_mul(self.h, tape_state, a, b, C.byref(d))
        _last_draws = int(d.value)
        return h

    def dec_value(self, c):
        o = np.zeros(2, np.uint64)
        rc = lib().orc_dec_value(self.h, c, _p(o, C.c_uint64))
        if rc:
            raise ValueError("malformed layer graph (the reference aborts here, ops/decrypt.hpp:23,36)")
        return o


def ct_free(c):
    lib().orc_ct_free(c)


def ct_export(c, with_sigma=True):
    nL, nE = C.c_uint32(), C.c_uint32()
    lib().orc_ct_counts(c, C.byref(nL), C.byref(nE))
    nL, nE = nL.value, nE.value
    d = dict(
        rule=np.zeros(nL, np.uint8), ztag=np.zeros(nL, np.uint64), nlo=np.zeros(nL, np.uint64), nhi=np.zeros(nL, np.uint64),
        pa=np.zeros(nL, np.uint32), pb=np.zeros(nL, np.uint32),
        lid=np.zeros(nE, np.uint32), idx=np.zeros(nE, np.uint16), ch=np.zeros(nE, np.uint8),
        w=np.zeros((nE, 2), np.uint64), sigma=np.zeros((nE, M_WORDS), np.uint64) if with_sigma else None,
    )
    lib().orc_ct_export(c, _p(d["rule"], C.c_uint8), _p(d["ztag"], C.c_uint64), _p(d["nlo"], C.c_uint64), _p(d["nhi"], C.c_uint64),
                        _p(d["pa"], C.c_uint32), _p(d["pb"], C.c_uint32), _p(d["lid"], C.c_uint32), _p(d["idx"], C.c_uint16),
                        _p(d["ch"], C.c_uint8), _p(d["w"], C.c_uint64), _p(d["sigma"], C.c_uint64) if with_sigma else None)
    return d


def ct_import(d):
    nL, nE = len(d["rule"]), len(d["lid"])
    sg = d.get("sigma")
    return lib().orc_ct_import(
        nL, nE, _p(d["rule"], C.c_uint8), _p(d["ztag"], C.c_uint64), _p(d["nlo"], C.c_uint64), _p(d["nhi"], C.c_uint64),
        _p(d["pa"], C.c_uint32), _p(d["pb"], C.c_uint32), _p(d["lid"], C.c_uint32), _p(d["idx"], C.c_uint16), _p(d["ch"], C.c_uint8),
        _p(np.ascontiguousarray(d["w"]), C.c_uint64), _p(np.ascontiguousarray(sg), C.c_uint64) if sg is not None else None)
