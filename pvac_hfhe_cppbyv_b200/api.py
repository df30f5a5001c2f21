"""Python host-side mirror of the reference API over libpvacb.so (include/pvacb.h).

The reference is a header-only C++ library; its boundary for this path is
    keygen / enc_value / ct_add / ct_sub / ct_mul / dec_value          (pvac/ops/*.hpp, pvac/crypto/keygen.hpp)
Here each of them takes an ARRAY of independent items and runs as CUDA kernels on one B200. There is no CPU fallback:
importing this module loads the CUDA library, and Engine() raises if no sm_100 device is present.

    eng = Engine(device=0)
    eng.keygen(tape_state=1)                       # or eng.import_keys(...)
    a = eng.enc_value(np.array([42, 17], np.uint64), batch_seed=1000)
    b = eng.enc_value(np.array([5, 6], np.uint64), batch_seed=2000)
    s, p = eng.ct_add(a, b), eng.ct_mul(a, b, batch_seed=3000)
    eng.dec_value(p)                                # -> uint64 array [n, 2] (lo, hi) of Fp values
"""
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpvacb.so")

B = 337
M_WORDS = 128
N_COLS = 16384
LPN_WORDS = 64
KEY_BLOB_BYTES = (752 + 16384 * 128) * 8
PRF_FAITHFUL, PRF_LIVE = 0, 1

TAPE_SPLITMIX, TAPE_CHACHA20, TAPE_WORDS = 0, 1, 2

STATUS = {0: "OK", 1: "E_ARG", 2: "E_CUDA", 3: "E_OOM", 4: "E_NOKEYS", 5: "E_EDGE_BUDGET", 6: "E_LAYER_GRAPH", 7: "E_RESERVED7",
          8: "E_RESERVED8", 9: "E_FORMAT", 10: "E_SHAPE"}


class Params(C.Structure):
    """struct pvacb_params = pvac::Params (core/types.hpp:36-70)"""
    _fields_ = [("B", C.c_int32), ("m_bits", C.c_int32), ("n_bits", C.c_int32), ("h_col_wt", C.c_int32), ("x_col_wt", C.c_int32), ("err_wt", C.c_int32),
                ("noise_entropy_bits", C.c_double), ("tuple2_fraction", C.c_double), ("depth_slope_bits", C.c_double), ("edge_budget", C.c_uint64),
                ("lpn_n", C.c_int32), ("lpn_t", C.c_int32), ("lpn_tau_num", C.c_int32), ("lpn_tau_den", C.c_int32),
                ("recrypt_lo", C.c_double), ("recrypt_hi", C.c_double), ("recrypt_rounds", C.c_int32)]

    @classmethod
    def default(cls):
        p = cls()
        load_library().pvacb_params_default(C.byref(p))
        return p

# every symbol include/pvacb.h declares
SYMBOLS = [
    "pvacb_ctx_create", "pvacb_ctx_destroy", "pvacb_last_error", "pvacb_set_prf_mode", "pvacb_get_prf_mode", "pvacb_stream", "pvacb_sync",
    "pvacb_stats", "pvacb_stats_reset", "pvacb_keygen", "pvacb_keys_import_raw", "pvacb_keys_export_raw", "pvacb_keys_device_blob",
    "pvacb_keys_alloc_blob", "pvacb_keys_adopt_blob", "pvacb_enc_value", "pvacb_enc_value_ex", "pvacb_ct_mul_ex", "pvacb_ct_add", "pvacb_ct_sub", "pvacb_ct_scale", "pvacb_ct_mul",
    "pvacb_dec_value", "pvacb_batch_free", "pvacb_batch_count", "pvacb_batch_totals", "pvacb_batch_device_bytes", "pvacb_batch_offsets",
    "pvacb_batch_slice", "pvacb_batch_export_soa", "pvacb_batch_import_soa", "pvacb_batch_wire_size", "pvacb_batch_export_wire",
    "pvacb_batch_import_wire", "pvacb_batch_synthetic", "pvacb_prf", "pvacb_sigma_from_H", "pvacb_fp_op",
    "pvacb_profile_enable", "pvacb_profile_collect", "pvacb_keys_copy_blob_to", "pvacb_keys_adopt_blob_from", "pvacb_l2_gather_probe",
    "pvacb_batch_export_soa_async", "pvacb_export_wait", "pvacb_export_wait_one", "pvacb_compact_edges", "pvacb_batch_checksum", "pvacb_commit_ct",
    "pvacb_enc_value_depth", "pvacb_enc_zero_depth", "pvacb_plan_noise", "pvacb_ct_neg", "pvacb_ct_div_const", "pvacb_enc_fp_depth",
    "pvacb_enc_text", "pvacb_dec_text", "pvacb_batch_concat",
    "pvacb_ct_recrypt", "pvacb_sigma_density", "pvacb_ubk_apply", "pvacb_ubk_perm", "pvacb_batch_select",
    "pvacb_params_default", "pvacb_keygen_params", "pvacb_set_params", "pvacb_get_params", "pvacb_keys_export_file", "pvacb_keys_import_file",
    "pvacb_set_tape", "pvacb_get_tape", "pvacb_set_tape_words", "pvacb_set_item_base", "pvacb_fresh_seed", "pvacb_debug_set",
    "pvacb_blob_layout", "pvacb_batch_blob_info", "pvacb_batch_export_blob_async", "pvacb_batch_import_blob", "pvacb_set_export_relay",
    "pvacb_group_create", "pvacb_group_destroy", "pvacb_group_size", "pvacb_group_ctx", "pvacb_group_last_error", "pvacb_group_keygen_params",
    "pvacb_group_keys_import_file", "pvacb_group_replicate_keys", "pvacb_group_set_tape", "pvacb_group_enc_value", "pvacb_group_ct_add", "pvacb_group_ct_sub",
    "pvacb_group_ct_mul", "pvacb_group_dec_value", "pvacb_group_commit_ct", "pvacb_group_batch_free", "pvacb_group_batch_count", "pvacb_group_batch_part",
    "pvacb_group_tune_export", "pvacb_group_export_blobs",
]


class PvacbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pvacb status {code} ({STATUS.get(code, '?')}): {msg}")
        self.code = code


_lib = None


def load_library():
    """Loads libpvacb.so (built by pvac_hfhe_cppbyv_b200.build). Fails loudly if it is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m pvac_hfhe_cppbyv_b200.build` (needs nvcc)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, u16, u8, i32, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8, C.c_int, C.c_size_t
    P = C.POINTER
    sig = {
        "pvacb_ctx_create": (i32, [i32, P(vp)]),
        "pvacb_ctx_destroy": (None, [vp]),
        "pvacb_last_error": (C.c_char_p, [vp]),
        "pvacb_set_prf_mode": (i32, [vp, i32]),
        "pvacb_get_prf_mode": (i32, [vp]),
        "pvacb_stream": (vp, [vp]),
        "pvacb_sync": (i32, [vp]),
        "pvacb_stats": (None, [vp, P(u64), P(u64), P(u64)]),
        "pvacb_stats_reset": (None, [vp]),
        "pvacb_keygen": (i32, [vp, u64]),
        "pvacb_keys_import_raw": (i32, [vp, u64, P(u8), P(u64), P(u64), P(u64), P(u64)]),
        "pvacb_keys_export_raw": (i32, [vp, P(u64), P(u8), P(u64), P(u64), P(u64), P(u64)]),
        "pvacb_keys_device_blob": (i32, [vp, P(vp), P(sz)]),
        "pvacb_keys_alloc_blob": (i32, [vp, P(vp)]),
        "pvacb_keys_adopt_blob": (i32, [vp]),
        "pvacb_enc_value": (i32, [vp, P(u64), sz, u64, P(vp)]),
        "pvacb_enc_value_ex": (i32, [vp, P(u64), sz, u64, P(u64), P(vp)]),
        "pvacb_ct_mul_ex": (i32, [vp, vp, vp, u64, P(u64), P(vp)]),
        "pvacb_ct_add": (i32, [vp, vp, vp, P(vp)]),
        "pvacb_ct_sub": (i32, [vp, vp, vp, P(vp)]),
        "pvacb_ct_scale": (i32, [vp, vp, P(u64), P(vp)]),
        "pvacb_ct_mul": (i32, [vp, vp, vp, u64, P(vp)]),
        "pvacb_dec_value": (i32, [vp, vp, P(u64)]),
        "pvacb_batch_free": (None, [vp]),
        "pvacb_batch_count": (sz, [vp]),
        "pvacb_batch_totals": (i32, [vp, P(u64), P(u64)]),
        "pvacb_batch_device_bytes": (sz, [vp]),
        "pvacb_batch_offsets": (i32, [vp, vp, P(u32), P(u32)]),
        "pvacb_batch_slice": (i32, [vp, vp, sz, sz, P(vp)]),
        "pvacb_batch_export_soa": (i32, [vp, vp, P(u32), P(u32), P(u8), P(u64), P(u64), P(u64), P(u32), P(u32), P(u32), P(u16), P(u8), P(u64), P(u64)]),
        "pvacb_batch_import_soa": (i32, [vp, sz, P(u32), P(u32), P(u8), P(u64), P(u64), P(u64), P(u32), P(u32), P(u32), P(u16), P(u8), P(u64), P(u64), P(vp)]),
        "pvacb_batch_wire_size": (i32, [vp, vp, P(sz)]),
        "pvacb_batch_export_wire": (i32, [vp, vp, vp, sz, P(sz)]),
        "pvacb_batch_import_wire": (i32, [vp, vp, sz, P(vp)]),
        "pvacb_batch_synthetic": (i32, [vp, sz, i32, u64, P(vp)]),
        "pvacb_prf": (i32, [vp, sz, P(u64), P(u64), P(u64), i32, P(u64), P(u64)]),
        "pvacb_sigma_from_H": (i32, [vp, sz, P(u64), P(u64), P(u64), P(u16), P(u8), P(u64), P(u64)]),
        "pvacb_fp_op": (i32, [vp, i32, sz, P(u64), P(u64), P(u64)]),
        "pvacb_profile_enable": (i32, [vp, i32]),
        "pvacb_profile_collect": (i32, [vp, P(C.c_float), P(u32)]),
        "pvacb_keys_copy_blob_to": (i32, [vp, vp]),
        "pvacb_keys_adopt_blob_from": (i32, [vp, vp]),
        "pvacb_l2_gather_probe": (i32, [vp, i32, P(C.c_double)]),
        "pvacb_batch_export_soa_async": (i32, [vp, vp, P(u32), P(u32), P(u8), P(u64), P(u64), P(u64), P(u32), P(u32), P(u32), P(u16), P(u8), P(u64), P(u64)]),
        "pvacb_export_wait": (i32, [vp]),
        "pvacb_export_wait_one": (i32, [vp]),
        "pvacb_compact_edges": (i32, [vp, vp, P(vp)]),
        "pvacb_batch_checksum": (i32, [vp, vp, P(u64)]),
        "pvacb_commit_ct": (i32, [vp, vp, P(u8)]),
        "pvacb_enc_value_depth": (i32, [vp, P(u64), sz, i32, u64, P(u64), P(vp)]),
        "pvacb_enc_zero_depth": (i32, [vp, sz, i32, u64, P(u64), P(vp)]),
        "pvacb_enc_fp_depth": (i32, [vp, P(u64), sz, i32, u64, P(u64), P(vp)]),
        "pvacb_enc_text": (i32, [vp, P(u8), P(u64), sz, u64, P(u64), P(vp)]),
        "pvacb_dec_text": (i32, [vp, vp, sz, P(u8), sz, P(u64)]),
        "pvacb_batch_concat": (i32, [vp, P(vp), sz, P(vp)]),
        "pvacb_ct_recrypt": (i32, [vp, vp, vp, u64, P(u64), P(vp)]),
        "pvacb_sigma_density": (i32, [vp, vp, P(C.c_double)]),
        "pvacb_ubk_apply": (i32, [vp, vp, P(vp)]),
        "pvacb_ubk_perm": (i32, [vp, P(u16)]),
        "pvacb_batch_select": (i32, [vp, P(vp), i32, P(u32), P(u32), sz, P(vp)]),
        "pvacb_plan_noise": (i32, [i32, P(i32), P(i32)]),
        "pvacb_ct_neg": (i32, [vp, vp, P(vp)]),
        "pvacb_ct_div_const": (i32, [vp, vp, P(u64), P(vp)]),
        "pvacb_params_default": (None, [P(Params)]),
        "pvacb_keygen_params": (i32, [vp, P(Params), P(u8)]),
        "pvacb_set_params": (i32, [vp, P(Params)]),
        "pvacb_get_params": (i32, [vp, P(Params)]),
        "pvacb_keys_export_file": (i32, [vp, C.c_char_p, C.c_char_p]),
        "pvacb_keys_import_file": (i32, [vp, C.c_char_p, C.c_char_p]),
        "pvacb_set_tape": (i32, [vp, i32, P(u8)]),
        "pvacb_get_tape": (i32, [vp]),
        "pvacb_set_tape_words": (i32, [vp, P(u64), sz, sz]),
        "pvacb_set_item_base": (i32, [vp, u64]),
        "pvacb_fresh_seed": (u64, [vp]),
        "pvacb_debug_set": (i32, [vp, i32, u64, u64]),
        "pvacb_blob_layout": (i32, [u64, u64, u64, P(u64)]),
        "pvacb_batch_blob_info": (i32, [vp, P(u64), P(u64), P(u64), P(u64)]),
        "pvacb_batch_export_blob_async": (i32, [vp, vp, vp, sz]),
        "pvacb_batch_import_blob": (i32, [vp, sz, u64, u64, vp, sz, P(vp)]),
        "pvacb_set_export_relay": (i32, [vp, i32]),
        "pvacb_group_create": (i32, [P(i32), i32, P(vp)]),
        "pvacb_group_destroy": (None, [vp]),
        "pvacb_group_size": (i32, [vp]),
        "pvacb_group_ctx": (vp, [vp, i32]),
        "pvacb_group_last_error": (C.c_char_p, [vp]),
        "pvacb_group_keygen_params": (i32, [vp, P(Params), P(u8)]),
        "pvacb_group_keys_import_file": (i32, [vp, C.c_char_p, C.c_char_p]),
        "pvacb_group_replicate_keys": (i32, [vp]),
        "pvacb_group_set_tape": (i32, [vp, i32, P(u8)]),
        "pvacb_group_enc_value": (i32, [vp, P(u64), sz, u64, P(vp)]),
        "pvacb_group_ct_add": (i32, [vp, vp, vp, P(vp)]),
        "pvacb_group_ct_sub": (i32, [vp, vp, vp, P(vp)]),
        "pvacb_group_ct_mul": (i32, [vp, vp, vp, u64, P(vp)]),
        "pvacb_group_dec_value": (i32, [vp, vp, P(u64)]),
        "pvacb_group_commit_ct": (i32, [vp, vp, P(u8)]),
        "pvacb_group_batch_free": (None, [vp]),
        "pvacb_group_batch_count": (sz, [vp]),
        "pvacb_group_batch_part": (vp, [vp, i32]),
        "pvacb_group_tune_export": (i32, [vp, P(C.c_double), P(C.c_double)]),
        "pvacb_group_export_blobs": (i32, [vp, vp, P(vp), P(sz)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class Batch:
    """Device-resident array of ciphertexts (the batched pvac::Cipher). Freed on garbage collection or .free()."""

    def __init__(self, eng, handle):
        self.eng = eng
        self.h = handle
        eng._batches.add(self)

    def __len__(self):
        return int(load_library().pvacb_batch_count(self.h))

    def totals(self):
        nl, ne = C.c_uint64(), C.c_uint64()
        load_library().pvacb_batch_totals(self.h, C.byref(nl), C.byref(ne))
        return int(nl.value), int(ne.value)

    def device_bytes(self):
        return int(load_library().pvacb_batch_device_bytes(self.h))

    def free(self):
        if self.h:
            if self.eng.h:          # the context owns the memory pool: never touch a batch after its engine was closed
                load_library().pvacb_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One B200: context + replicated keys. Mirrors the reference's free functions as methods over batches."""

    def __init__(self, device=0, prf_mode=PRF_FAITHFUL, tape=TAPE_CHACHA20, tape_key=None):
        """tape: RNG tape kind of the context (include/pvacb.h). The default draws every random word from ChaCha20 under a key taken
        from the OS (or tape_key, 32 bytes); TAPE_SPLITMIX is for parity tests against the committed golden vectors only."""
        L = load_library()
        h = C.c_void_p()
        rc = L.pvacb_ctx_create(device, C.byref(h))
        if rc:
            raise PvacbError(rc, f"cannot create a context on CUDA device {device} (an sm_100-class GPU is required; there is no CPU fallback)")
        self.L, self.h, self.device = L, h, device
        self._batches = weakref.WeakSet()
        self.set_prf_mode(prf_mode)
        if tape != TAPE_CHACHA20 or tape_key is not None:
            self.set_tape(tape, tape_key)

    # ---- RNG tape
    def set_tape(self, kind, key=None):
        k = np.frombuffer(bytes(key), np.uint8).copy() if key is not None else None
        if k is not None and len(k) != 32:
            raise PvacbError(1, "tape key must be 32 bytes")
        self._ck(self.L.pvacb_set_tape(self.h, kind, _p(k, C.c_uint8)))

    def set_tape_words(self, words):
        """words: (n_items, words_per_item) uint64 -- every random word of every item (tape kind TAPE_WORDS)"""
        w = np.ascontiguousarray(words, np.uint64)
        self._ck(self.L.pvacb_set_tape_words(self.h, _p(w, C.c_uint64), w.shape[0], w.shape[1]))

    def set_item_base(self, base):
        self._ck(self.L.pvacb_set_item_base(self.h, int(base)))

    def fresh_seed(self):
        return int(self.L.pvacb_fresh_seed(self.h))

    def _seed(self, batch_seed, tape_states):
        """a repeated seed repeats nonces, masks and salts: with neither a seed nor explicit states the context hands out a fresh one"""
        if batch_seed is None:
            return 0 if tape_states is not None else self.fresh_seed()
        return int(batch_seed)

    def debug_set(self, what, a=0, b=0):
        self._ck(self.L.pvacb_debug_set(self.h, what, int(a) & (2**64 - 1), int(b) & (2**64 - 1)))

    def close(self):
        if self.h:
            for b in list(self._batches):
                b.free()
            self.L.pvacb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise PvacbError(rc, self.L.pvacb_last_error(self.h).decode(errors="replace"))

    # ---- context
    def set_prf_mode(self, mode):
        self._ck(self.L.pvacb_set_prf_mode(self.h, mode))

    @property
    def stream(self):
        return self.L.pvacb_stream(self.h)

    def sync(self):
        self._ck(self.L.pvacb_sync(self.h))

    def stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.L.pvacb_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(kernel_launches=int(a.value), aes_blocks=int(b.value), sigma_edges=int(c.value))

    def stats_reset(self):
        self.L.pvacb_stats_reset(self.h)

    # ---- keys (crypto/keygen.hpp:35)
    def keygen(self, tape_state):
        """PARITY / TEST ONLY (SplitMix64 from a 64-bit state, the committed golden keys); use keygen_params for real keys"""
        self._ck(self.L.pvacb_keygen(self.h, tape_state))

    def keygen_params(self, params=None, seed=None):
        """keygen(prm, pk, sk) of the reference (crypto/keygen.hpp:35); seed: 32 bytes or None = OS CSPRNG"""
        sd = np.frombuffer(bytes(seed), np.uint8).copy() if seed is not None else None
        self._ck(self.L.pvacb_keygen_params(self.h, C.byref(params) if params is not None else None, _p(sd, C.c_uint8)))

    def set_params(self, params):
        self._ck(self.L.pvacb_set_params(self.h, C.byref(params)))

    def get_params(self):
        p = Params()
        self._ck(self.L.pvacb_get_params(self.h, C.byref(p)))
        return p

    def export_key_files(self, pk_path=None, sk_path=None):
        """savePk / saveSk of the reference's programs (tests/bounty2_test.cpp:145-192), byte for byte"""
        self._ck(self.L.pvacb_keys_export_file(self.h, pk_path.encode() if pk_path else None, sk_path.encode() if sk_path else None))

    def import_key_files(self, pk_path, sk_path=None):
        self._ck(self.L.pvacb_keys_import_file(self.h, pk_path.encode(), sk_path.encode() if sk_path else None))

    def import_keys(self, canon_tag, H_digest, H, powg, prf_k, lpn_s):
        hd = np.ascontiguousarray(H_digest, np.uint8)
        Hc = _u64(H) if H is not None else None
        gc = _u64(powg) if powg is not None else None
        pk, ls = _u64(prf_k), _u64(lpn_s)
        self._ck(self.L.pvacb_keys_import_raw(self.h, int(canon_tag), _p(hd, C.c_uint8), _p(Hc, C.c_uint64), _p(gc, C.c_uint64), _p(pk, C.c_uint64), _p(ls, C.c_uint64)))

    def export_keys(self, with_H=True):
        ct = C.c_uint64()
        hd = np.zeros(32, np.uint8)
        H = np.zeros((N_COLS, M_WORDS), np.uint64) if with_H else None
        powg = np.zeros((B, 2), np.uint64)
        prf_k = np.zeros(4, np.uint64)
        lpn_s = np.zeros(LPN_WORDS, np.uint64)
        self._ck(self.L.pvacb_keys_export_raw(self.h, C.byref(ct), _p(hd, C.c_uint8), _p(H, C.c_uint64), _p(powg, C.c_uint64), _p(prf_k, C.c_uint64), _p(lpn_s, C.c_uint64)))
        return dict(canon_tag=int(ct.value), H_digest=hd, H=H, powg=powg, prf_k=prf_k, lpn_s=lpn_s)

    def key_blob_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.pvacb_keys_device_blob(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def alloc_key_blob(self):
        p = C.c_void_p()
        self._ck(self.L.pvacb_keys_alloc_blob(self.h, C.byref(p)))
        return int(p.value), KEY_BLOB_BYTES

    def adopt_key_blob(self):
        self._ck(self.L.pvacb_keys_adopt_blob(self.h))

    def copy_key_blob_to(self, device_ptr):
        self._ck(self.L.pvacb_keys_copy_blob_to(self.h, device_ptr))

    def adopt_key_blob_from(self, device_ptr):
        self._ck(self.L.pvacb_keys_adopt_blob_from(self.h, device_ptr))

    PROF_TAGS = ("prf_lpn", "mul_pairs", "sigma", "concat", "dec_edges", "commit", "compact", "t7")

    def l2_gather_probe(self, reps=5):
        g = C.c_double()
        self._ck(self.L.pvacb_l2_gather_probe(self.h, reps, C.byref(g)))
        return float(g.value)

    def profile_enable(self, on=True):
        self._ck(self.L.pvacb_profile_enable(self.h, int(on)))

    def profile_collect(self):
        """-> {tag: (total ms, launches)} of the bracketed kernels since the last collect (CUDA events on the engine stream)."""
        ms = (C.c_float * 8)()
        cnt = (C.c_uint32 * 8)()
        self._ck(self.L.pvacb_profile_collect(self.h, ms, cnt))
        return {t: (float(ms[i]), int(cnt[i])) for i, t in enumerate(self.PROF_TAGS)}

    # ---- hot path
    def enc_value(self, values, batch_seed=None, tape_states=None):
        v = _u64(values)
        out = C.c_void_p()
        st = _u64(tape_states) if tape_states is not None else None
        self._ck(self.L.pvacb_enc_value_ex(self.h, _p(v, C.c_uint64), len(v), self._seed(batch_seed, tape_states), _p(st, C.c_uint64), C.byref(out)))
        return Batch(self, out)

    def ct_add(self, a, b):
        out = C.c_void_p()
        self._ck(self.L.pvacb_ct_add(self.h, a.h, b.h, C.byref(out)))
        return Batch(self, out)

    def ct_sub(self, a, b):
        out = C.c_void_p()
        self._ck(self.L.pvacb_ct_sub(self.h, a.h, b.h, C.byref(out)))
        return Batch(self, out)

    def ct_scale(self, a, s):
        ss = _u64(s)
        out = C.c_void_p()
        self._ck(self.L.pvacb_ct_scale(self.h, a.h, _p(ss, C.c_uint64), C.byref(out)))
        return Batch(self, out)

    def enc_value_depth(self, values, depth_hint, batch_seed=None, tape_states=None):
        v = _u64(values)
        st = _u64(tape_states) if tape_states is not None else None
        out = C.c_void_p()
        self._ck(self.L.pvacb_enc_value_depth(self.h, _p(v, C.c_uint64), len(v), depth_hint, self._seed(batch_seed, tape_states), _p(st, C.c_uint64) if st is not None else None, C.byref(out)))
        return Batch(self, out)

    def enc_fp_depth(self, fp_values, depth_hint=0, batch_seed=None, tape_states=None):
        """fp_values: (n, 2) uint64 canonical field elements -> one-share ciphertexts (enc_fp_depth, ops/encrypt.hpp:162)"""
        v = np.ascontiguousarray(fp_values, np.uint64).reshape(-1, 2)
        st = _u64(tape_states) if tape_states is not None else None
        out = C.c_void_p()
        self._ck(self.L.pvacb_enc_fp_depth(self.h, _p(v, C.c_uint64), len(v), depth_hint, self._seed(batch_seed, tape_states), _p(st, C.c_uint64) if st is not None else None, C.byref(out)))
        return Batch(self, out)

    def enc_zero_depth(self, n, depth_hint, batch_seed=None, tape_states=None):
        st = _u64(tape_states) if tape_states is not None else None
        out = C.c_void_p()
        self._ck(self.L.pvacb_enc_zero_depth(self.h, n, depth_hint, self._seed(batch_seed, tape_states), _p(st, C.c_uint64) if st is not None else None, C.byref(out)))
        return Batch(self, out)

    def plan_noise(self, depth_hint):
        a, b = C.c_int(), C.c_int()
        self.L.pvacb_plan_noise(depth_hint, C.byref(a), C.byref(b))
        return a.value, b.value

    def ct_neg(self, a):
        out = C.c_void_p()
        self._ck(self.L.pvacb_ct_neg(self.h, a.h, C.byref(out)))
        return Batch(self, out)

    def ct_div_const(self, a, k):
        kk = _u64(k)
        out = C.c_void_p()
        self._ck(self.L.pvacb_ct_div_const(self.h, a.h, _p(kk, C.c_uint64), C.byref(out)))
        return Batch(self, out)

    def enc_text(self, msgs, batch_seed=None, tape_states=None):
        """msgs: list of bytes -> one WAVE-MAJOR batch (lengths, then block 0 of every message, then block 1, ...)"""
        off = np.zeros(len(msgs) + 1, np.uint64)
        for i, m in enumerate(msgs):
            off[i + 1] = off[i] + len(m)
        flat = np.frombuffer(b"".join(msgs) or b"\0", np.uint8).copy()
        st = _u64(tape_states) if tape_states is not None else None
        out = C.c_void_p()
        self._ck(self.L.pvacb_enc_text(self.h, _p(flat, C.c_uint8), _p(off, C.c_uint64), len(msgs), self._seed(batch_seed, tape_states), _p(st, C.c_uint64) if st is not None else None, C.byref(out)))
        return Batch(self, out)

    def dec_text(self, c, n_msgs):
        cap = 15 * max(len(c) - n_msgs, 0) + 16
        buf = np.zeros(cap, np.uint8)
        off = np.zeros(n_msgs + 1, np.uint64)
        self._ck(self.L.pvacb_dec_text(self.h, c.h, n_msgs, _p(buf, C.c_uint8), cap, _p(off, C.c_uint64)))
        return [buf[int(off[i]):int(off[i + 1])].tobytes() for i in range(n_msgs)]

    def concat(self, parts):
        arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
        out = C.c_void_p()
        self._ck(self.L.pvacb_batch_concat(self.h, arr, len(parts), C.byref(out)))
        return Batch(self, out)

    def ct_recrypt(self, c, zero_pool, batch_seed=None, tape_states=None):
        st = _u64(tape_states) if tape_states is not None else None
        out = C.c_void_p()
        self._ck(self.L.pvacb_ct_recrypt(self.h, c.h, zero_pool.h, self._seed(batch_seed, tape_states), _p(st, C.c_uint64) if st is not None else None, C.byref(out)))
        return Batch(self, out)

    def make_evalkey(self, pool_size, depth_hint, batch_seed=None, one_seed=None):
        """EvalKey of the reference (ops/recrypt.hpp:12): (zero_pool batch, enc_one batch); every entry on its own tape stream.
        The two seeds must differ (and differ from every other seed used under this key); None = fresh seeds from the context."""
        return self.enc_zero_depth(pool_size, depth_hint, batch_seed), self.enc_value(np.array([1], np.uint64), one_seed)

    def sigma_density(self, c):
        out = np.zeros(len(c), np.float64)
        self._ck(self.L.pvacb_sigma_density(self.h, c.h, _p(out, C.c_double)))
        return out

    def ubk_apply(self, c):
        out = C.c_void_p()
        self._ck(self.L.pvacb_ubk_apply(self.h, c.h, C.byref(out)))
        return Batch(self, out)

    def ubk_perm(self):
        o = np.zeros(8192, np.uint16)
        self._ck(self.L.pvacb_ubk_perm(self.h, _p(o, C.c_uint16)))
        return o

    def select(self, srcs, which, index):
        arr = (C.c_void_p * len(srcs))(*[s.h for s in srcs])
        w, ix = np.ascontiguousarray(which, np.uint32), np.ascontiguousarray(index, np.uint32)
        out = C.c_void_p()
        self._ck(self.L.pvacb_batch_select(self.h, arr, len(srcs), _p(w, C.c_uint32), _p(ix, C.c_uint32), len(w), C.byref(out)))
        return Batch(self, out)

    def compact_edges(self, a):
        out = C.c_void_p()
        self._ck(self.L.pvacb_compact_edges(self.h, a.h, C.byref(out)))
        return Batch(self, out)

    def ct_mul(self, a, b, batch_seed=None, tape_states=None):
        out = C.c_void_p()
        st = _u64(tape_states) if tape_states is not None else None
        self._ck(self.L.pvacb_ct_mul_ex(self.h, a.h, b.h, self._seed(batch_seed, tape_states), _p(st, C.c_uint64), C.byref(out)))
        return Batch(self, out)

    def dec_value(self, c):
        out = np.zeros((len(c), 2), np.uint64)
        self._ck(self.L.pvacb_dec_value(self.h, c.h, _p(out, C.c_uint64)))
        return out

    # ---- batches
    def offsets(self, b):
        n = len(b)
        lo, eo = np.zeros(n + 1, np.uint32), np.zeros(n + 1, np.uint32)
        self._ck(self.L.pvacb_batch_offsets(self.h, b.h, _p(lo, C.c_uint32), _p(eo, C.c_uint32)))
        return lo, eo

    def slice(self, b, first, count):
        out = C.c_void_p()
        self._ck(self.L.pvacb_batch_slice(self.h, b.h, first, count, C.byref(out)))
        return Batch(self, out)

    def export_soa(self, b, with_sigma=True, pinned=None):
        """-> dict of numpy arrays for the whole batch (offsets + concatenated layer / edge arrays)."""
        n = len(b)
        nL, nE = b.totals()
        d = dict(
            loff=np.zeros(n + 1, np.uint32), eoff=np.zeros(n + 1, np.uint32),
            rule=np.zeros(nL, np.uint8), ztag=np.zeros(nL, np.uint64), nlo=np.zeros(nL, np.uint64), nhi=np.zeros(nL, np.uint64),
            pa=np.zeros(nL, np.uint32), pb=np.zeros(nL, np.uint32),
            lid=np.zeros(nE, np.uint32), idx=np.zeros(nE, np.uint16), ch=np.zeros(nE, np.uint8), w=np.zeros((nE, 2), np.uint64),
            sigma=(pinned if pinned is not None else np.zeros((nE, M_WORDS), np.uint64)) if with_sigma else None,
        )
        self._ck(self.L.pvacb_batch_export_soa(
            self.h, b.h, _p(d["loff"], C.c_uint32), _p(d["eoff"], C.c_uint32), _p(d["rule"], C.c_uint8), _p(d["ztag"], C.c_uint64),
            _p(d["nlo"], C.c_uint64), _p(d["nhi"], C.c_uint64), _p(d["pa"], C.c_uint32), _p(d["pb"], C.c_uint32), _p(d["lid"], C.c_uint32),
            _p(d["idx"], C.c_uint16), _p(d["ch"], C.c_uint8), _p(d["w"], C.c_uint64), _p(d["sigma"], C.c_uint64) if with_sigma else None))
        return d

    def commit_ct(self, c):
        """-> (n, 32) uint8: commit_ct digest of every ciphertext"""
        out = np.zeros((len(c), 32), np.uint8)
        self._ck(self.L.pvacb_commit_ct(self.h, c.h, _p(out, C.c_uint8)))
        return out

    def checksum(self, b):
        """order-independent device checksums of the whole batch -> dict (see pvacb_batch_checksum)"""
        o = np.zeros(8, np.uint64)
        self._ck(self.L.pvacb_batch_checksum(self.h, b.h, _p(o, C.c_uint64)))
        return dict(xor_sigma=int(o[0]), sum_wlo=int(o[1]), sum_whi=int(o[2]), sum_lid=int(o[3]), sum_idx=int(o[4]), sum_ch=int(o[5]), layers=int(o[6]), edges=int(o[7]))

    def export_soa_async(self, b, bufs):
        """queue the device->host copies of batch b into the (pinned, large enough) arrays of `bufs` and return at once;
        call export_wait() before reading them or freeing b. -> dict of views trimmed to the batch's sizes."""
        n = len(b)
        nL, nE = b.totals()
        sizes = dict(loff=n + 1, eoff=n + 1, rule=nL, ztag=nL, nlo=nL, nhi=nL, pa=nL, pb=nL, lid=nE, idx=nE, ch=nE, w=nE, sigma=nE)
        d = {}
        for k, cnt in sizes.items():
            a = bufs.get(k)
            if a is None:
                d[k] = None
                continue
            if len(a) < cnt:
                raise PvacbError(1, f"export buffer {k} too small: {len(a)} < {cnt}")
            d[k] = a[:cnt]
        ptr = lambda k, t: _p(d[k], t) if d[k] is not None else None
        self._ck(self.L.pvacb_batch_export_soa_async(
            self.h, b.h, ptr("loff", C.c_uint32), ptr("eoff", C.c_uint32), ptr("rule", C.c_uint8), ptr("ztag", C.c_uint64), ptr("nlo", C.c_uint64),
            ptr("nhi", C.c_uint64), ptr("pa", C.c_uint32), ptr("pb", C.c_uint32), ptr("lid", C.c_uint32), ptr("idx", C.c_uint16), ptr("ch", C.c_uint8),
            ptr("w", C.c_uint64), ptr("sigma", C.c_uint64)))
        return d

    # ---- whole-batch image: one copy per batch (and, optionally, out through another GPU's host link)
    def blob_info(self, b):
        """-> (n, layout_layers, layout_edges, bytes) of the batch's device image"""
        n, nl, ne, by = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self.L.pvacb_batch_blob_info(b.h, C.byref(n), C.byref(nl), C.byref(ne), C.byref(by)))
        return int(n.value), int(nl.value), int(ne.value), int(by.value)

    def export_blob_async(self, b, host_u8):
        """queue ONE device->host copy of the batch image into host_u8 (pinned uint8 array, large enough); wait with export_wait[_one];
        blob_views(host_u8, *blob_info(b)[:3]) then gives the arrays"""
        self._ck(self.L.pvacb_batch_export_blob_async(self.h, b.h, host_u8.ctypes.data_as(C.c_void_p), host_u8.nbytes))

    def import_blob(self, host_u8, n, layout_layers, layout_edges):
        out = C.c_void_p()
        off = blob_layout(n, layout_layers, layout_edges)
        self._ck(self.L.pvacb_batch_import_blob(self.h, n, layout_layers, layout_edges, host_u8.ctypes.data_as(C.c_void_p), off[13], C.byref(out)))
        return Batch(self, out)

    def set_export_relay(self, device):
        self._ck(self.L.pvacb_set_export_relay(self.h, device))

    def export_wait(self):
        self._ck(self.L.pvacb_export_wait(self.h))

    def export_wait_one(self):
        """wait for the oldest outstanding export_soa_async only (later ones keep running)"""
        self._ck(self.L.pvacb_export_wait_one(self.h))

    def import_soa(self, d):
        n = len(d["loff"]) - 1
        out = C.c_void_p()
        sg = d.get("sigma")
        arr = {k: np.ascontiguousarray(v) for k, v in d.items() if v is not None}
        self._ck(self.L.pvacb_batch_import_soa(
            self.h, n, _p(arr["loff"], C.c_uint32), _p(arr["eoff"], C.c_uint32), _p(arr["rule"], C.c_uint8), _p(arr["ztag"], C.c_uint64),
            _p(arr["nlo"], C.c_uint64), _p(arr["nhi"], C.c_uint64), _p(arr["pa"], C.c_uint32), _p(arr["pb"], C.c_uint32), _p(arr["lid"], C.c_uint32),
            _p(arr["idx"], C.c_uint16), _p(arr["ch"], C.c_uint8), _p(arr["w"], C.c_uint64), _p(arr["sigma"], C.c_uint64) if sg is not None else None,
            C.byref(out)))
        return Batch(self, out)

    def export_wire(self, b) -> bytes:
        n = C.c_size_t()
        self._ck(self.L.pvacb_batch_wire_size(self.h, b.h, C.byref(n)))
        buf = (C.c_char * n.value)()
        w = C.c_size_t()
        self._ck(self.L.pvacb_batch_export_wire(self.h, b.h, buf, n.value, C.byref(w)))
        return bytes(buf[: w.value])

    def import_wire(self, data: bytes):
        out = C.c_void_p()
        buf = C.create_string_buffer(data, len(data))
        self._ck(self.L.pvacb_batch_import_wire(self.h, buf, len(data), C.byref(out)))
        return Batch(self, out)

    def synthetic(self, n, edges_per_layer=20, batch_seed=1):
        out = C.c_void_p()
        self._ck(self.L.pvacb_batch_synthetic(self.h, n, edges_per_layer, batch_seed, C.byref(out)))
        return Batch(self, out)

    # ---- building blocks (device kernels) for parity tests
    def prf(self, ztag, nlo, nhi, family=0, want_ybits=False):
        z, lo, hi = _u64(ztag), _u64(nlo), _u64(nhi)
        n = len(z)
        out = np.zeros((n, 2), np.uint64)
        wpc = 2 if self.L.pvacb_get_prf_mode(self.h) == PRF_LIVE else 256
        yb = np.zeros((n * 3, wpc), np.uint64) if want_ybits else None
        self._ck(self.L.pvacb_prf(self.h, n, _p(z, C.c_uint64), _p(lo, C.c_uint64), _p(hi, C.c_uint64), family, _p(out, C.c_uint64), _p(yb, C.c_uint64)))
        return (out, yb) if want_ybits else out

    def sigma_from_H(self, ztag, nlo, nhi, idx, ch, salt):
        z, lo, hi, s = _u64(ztag), _u64(nlo), _u64(nhi), _u64(salt)
        i = np.ascontiguousarray(idx, np.uint16)
        c = np.ascontiguousarray(ch, np.uint8)
        out = np.zeros((len(z), M_WORDS), np.uint64)
        self._ck(self.L.pvacb_sigma_from_H(self.h, len(z), _p(z, C.c_uint64), _p(lo, C.c_uint64), _p(hi, C.c_uint64), _p(i, C.c_uint16), _p(c, C.c_uint8),
                                           _p(s, C.c_uint64), _p(out, C.c_uint64)))
        return out

    def fp_op(self, op, a, b=None):
        aa = _u64(a).reshape(-1, 2)
        bb = _u64(b).reshape(-1, 2) if b is not None else None
        out = np.zeros_like(aa)
        self._ck(self.L.pvacb_fp_op(self.h, op, len(aa), _p(aa, C.c_uint64), _p(bb, C.c_uint64), _p(out, C.c_uint64)))
        return out


def blob_layout(n, n_layers, n_edges):
    off = (C.c_uint64 * 14)()
    load_library().pvacb_blob_layout(n, n_layers, n_edges, off)
    return [int(x) for x in off]


_BLOB_FIELDS = (("loff", np.uint32, 1), ("eoff", np.uint32, 1), ("rule", np.uint8, 1), ("ztag", np.uint64, 1), ("nlo", np.uint64, 1), ("nhi", np.uint64, 1),
                ("pa", np.uint32, 1), ("pb", np.uint32, 1), ("lid", np.uint32, 1), ("idx", np.uint16, 1), ("ch", np.uint8, 1), ("w", np.uint64, 2), ("sigma", np.uint64, M_WORDS))


def blob_views(host_u8, n, layout_layers, layout_edges):
    """numpy views (no copy) of the 13 arrays inside a batch image; the layer arrays are trimmed to the logical layer count loff[n]"""
    off = blob_layout(n, layout_layers, layout_edges)
    d = {}
    for k, (name, dt, width) in enumerate(_BLOB_FIELDS):
        cnt = n + 1 if k < 2 else layout_layers if k < 8 else layout_edges
        a = host_u8[off[k]: off[k] + cnt * width * np.dtype(dt).itemsize].view(dt)
        d[name] = a.reshape(cnt, width) if width > 1 else a
    nl = int(d["loff"][n])
    for name in ("rule", "ztag", "nlo", "nhi", "pa", "pb"):
        d[name] = d[name][:nl]
    return d


def pack_blob(d, out_u8=None):
    """the image of an SoA dict (export_soa layout) -> (uint8 array, n, layers, edges), for import_blob"""
    n, nl, ne = len(d["loff"]) - 1, len(d["rule"]), len(d["lid"])
    off = blob_layout(n, nl, ne)
    buf = out_u8 if out_u8 is not None else np.zeros(off[13], np.uint8)
    for k, (name, dt, width) in enumerate(_BLOB_FIELDS):
        v = d.get(name)
        if v is None:
            continue
        raw = np.ascontiguousarray(v, dt).reshape(-1).view(np.uint8)
        buf[off[k]: off[k] + raw.size] = raw
    return buf, n, nl, ne


def checksum_of_soa(d):
    """the same checksums as Engine.checksum, from an exported SoA dict (numpy, on the host)"""
    M = (1 << 64) - 1
    x = int(np.bitwise_xor.reduce(d["sigma"].reshape(-1))) if d["sigma"].size else 0
    w = np.asarray(d["w"], np.uint64).reshape(-1, 2)
    lay = 0
    for r, z, lo, hi, pa, pb in zip(d["rule"], d["ztag"], d["nlo"], d["nhi"], d["pa"], d["pb"]):
        lay = (lay + ((int(z) ^ int(lo) ^ int(hi)) if r == 0 else (int(pa) + (int(pb) << 32)))) & M
    return dict(xor_sigma=x, sum_wlo=int(sum(int(v) for v in w[:, 0]) & M), sum_whi=int(sum(int(v) for v in w[:, 1]) & M),
                sum_lid=int(np.sum(d["lid"], dtype=np.uint64)), sum_idx=int(np.sum(d["idx"], dtype=np.uint64)), sum_ch=int(np.sum(d["ch"], dtype=np.uint64)),
                layers=lay, edges=len(d["lid"]))


def split_items(d):
    """Splits an export_soa() dict into one dict per ciphertext (same keys as the oracle's export)."""
    out = []
    for i in range(len(d["loff"]) - 1):
        l0, l1, e0, e1 = int(d["loff"][i]), int(d["loff"][i + 1]), int(d["eoff"][i]), int(d["eoff"][i + 1])
        out.append(dict(
            rule=d["rule"][l0:l1], ztag=d["ztag"][l0:l1], nlo=d["nlo"][l0:l1], nhi=d["nhi"][l0:l1], pa=d["pa"][l0:l1], pb=d["pb"][l0:l1],
            lid=d["lid"][e0:e1], idx=d["idx"][e0:e1], ch=d["ch"][e0:e1], w=d["w"][e0:e1],
            sigma=d["sigma"][e0:e1] if d.get("sigma") is not None else None))
    return out


def join_items(items):
    """Inverse of split_items."""
    loff = np.zeros(len(items) + 1, np.uint32)
    eoff = np.zeros(len(items) + 1, np.uint32)
    for i, it in enumerate(items):
        loff[i + 1] = loff[i] + len(it["rule"])
        eoff[i + 1] = eoff[i] + len(it["lid"])

    def cat(k, dt, shape=()):
        parts = [np.asarray(it[k], dt).reshape((-1,) + shape) for it in items]
        return np.concatenate(parts) if parts else np.zeros((0,) + shape, dt)

    d = dict(loff=loff, eoff=eoff, rule=cat("rule", np.uint8), ztag=cat("ztag", np.uint64), nlo=cat("nlo", np.uint64), nhi=cat("nhi", np.uint64),
             pa=cat("pa", np.uint32), pb=cat("pb", np.uint32), lid=cat("lid", np.uint32), idx=cat("idx", np.uint16), ch=cat("ch", np.uint8),
             w=cat("w", np.uint64, (2,)))
    d["sigma"] = cat("sigma", np.uint64, (M_WORDS,)) if all(it.get("sigma") is not None for it in items) else None
    return d
