"""CPU-only checks of the engine's own code (no GPU needed):

  * the C-ABI library libpvacb.so loads and exports every function include/pvacb.h declares (no compute calls);
  * the __host__ __device__ per-thread bodies of the kernels (Fp, tape, SHA-256 message layout, AES-256, one PRF core,
    enc_value planning and weights, keygen, the libstdc++ bucket table), compiled for the CPU by csrc/hosttest.cpp, against
    the oracle. These are the same source lines the kernels execute; the -m gpu tests check the kernels themselves.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, hexwords, load_npz

PKG = os.path.join(ROOT, "pvac_hfhe_cppbyv_b200")
P = (1 << 127) - 1
u64, u8, i32 = C.c_uint64, C.c_uint8, C.c_int


def _p(a, t=u64):
    return a.ctypes.data_as(C.POINTER(t))


# ------------------------------------------------------------------------------------------------ the C ABI
def test_header_symbols_exported():
    lib = os.path.join(PKG, "libpvacb.so")
    if not os.path.exists(lib):
        from pvac_hfhe_cppbyv_b200 import build
        build.build()
    hdr = open(os.path.join(ROOT, "include", "pvacb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pvacb_[A-Za-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 40
    L = C.CDLL(lib)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    from pvac_hfhe_cppbyv_b200 import api
    assert sorted(api.SYMBOLS) == declared          # the Python mirror binds exactly the header
    # no C++ or CUDA type crosses the boundary: the dynamic symbols are unmangled
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared) <= exported
    # INTEGRATION.md maps every entry point to the reference interface it replaces (or says what it is for)
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    undocumented = [s for s in declared if s not in integ]
    assert not undocumented, undocumented


def test_product_has_no_oracle_or_cpu_fallback():
    """the product package never imports / links the oracle; api.load_library() raises when the .so is missing"""
    for dirpath, _, files in os.walk(PKG):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pvac_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f
    from pvac_hfhe_cppbyv_b200 import api
    saved, api._lib, api.LIB_PATH = (api._lib, api.LIB_PATH), None, os.path.join(PKG, "does_not_exist.so")
    try:
        with pytest.raises(ImportError):
            api.load_library()
    finally:
        api._lib, api.LIB_PATH = saved


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pvac_hfhe_cppbyv_b200 import api
    with pytest.raises(api.PvacbError) as e:
        api.Engine(device=0)
    assert e.value.code == 2      # PVACB_E_CUDA


# ------------------------------------------------------------------------------------------------ kernel bodies on the CPU
@pytest.fixture(scope="module")
def ht():
    so = os.path.join(PKG, "_build", "libpvacb_hosttest.so")
    csrc = os.path.join(PKG, "csrc")
    srcs = [os.path.join(csrc, "hosttest.cpp"), os.path.join(csrc, "keygen.cpp")]
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h", ".cpp"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-o", so, "-x", "c++"] + srcs)
    L = C.CDLL(so)
    L.ht_tape_word.restype = u64
    L.ht_tape_word.argtypes = [u64, u64]
    L.ht_item_stream_state.restype = u64
    L.ht_item_stream_state.argtypes = [u64, u64]
    L.ht_fnv.restype = u64
    L.ht_ztag.restype = u64
    L.ht_ztag.argtypes = [u64, u64, u64]
    L.ht_fp_from_words.argtypes = [u64, u64, C.POINTER(u64)]
    L.ht_fp_mac_chain.argtypes = [C.c_size_t, C.POINTER(u64), C.POINTER(u64), u64, C.POINTER(u64), C.POINTER(u64)]
    L.ht_cand_words.argtypes = [i32, C.POINTER(u64), u64, C.POINTER(u64)]
    L.ht_aes_ctr_words.argtypes = [C.POINTER(u8), u64, C.POINTER(u64), C.c_size_t]
    L.ht_prf_core.argtypes = [C.POINTER(u64), u64, C.POINTER(u8), C.POINTER(u64), u64, u64, u64, i32, i32, i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)]
    L.ht_plan_item.restype = u64
    L.ht_plan_item.argtypes = [u64, u64, u64, i32, i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.ht_item_weights.argtypes = [u64, u64, u64, i32, i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.ht_keygen.argtypes = [u64, C.POINTER(u64), C.c_size_t]
    L.ht_next_bkt.restype = u64
    L.ht_next_bkt.argtypes = [u64]
    L.ht_unordered_buckets_real.restype = u64
    L.ht_unordered_buckets_real.argtypes = [u64]
    return L


def test_fp_bodies(ht, port, kat):
    def op(k, a, b=None):
        o = np.zeros(2, np.uint64)
        aa = np.array(a, np.uint64)
        bb = np.array(b, np.uint64) if b is not None else None
        ht.ht_fp_op(k, _p(aa), _p(bb) if bb is not None else None, _p(o))
        return o
    for c in kat["fp"]:
        a, b = [int(x, 16) for x in c["a"]], [int(x, 16) for x in c["b"]]
        assert hexwords(op(0, a, b)) == c["add"] and hexwords(op(1, a, b)) == c["sub"] and hexwords(op(2, a, b)) == c["mul"]
        assert hexwords(op(3, a)) == c["neg"]
        if c["inv"]:
            assert hexwords(op(4, a)) == c["inv"]
    rng = np.random.default_rng(3)
    for _ in range(500):
        a = int.from_bytes(rng.bytes(16), "little") % P
        b = int.from_bytes(rng.bytes(16), "little") % P
        aw, bw = [a & (2**64 - 1), a >> 64], [b & (2**64 - 1), b >> 64]
        m = op(2, aw, bw)
        assert int(m[0]) | (int(m[1]) << 64) == a * b % P
    for lo, hi in ((2**64 - 1, 2**63 - 1), (2**64 - 1, 2**64 - 1), (0, 2**63), (5, 2**63 + 9)):   # p -> 0, folds of bit 127
        o = np.zeros(2, np.uint64)
        ht.ht_fp_from_words(lo, hi, _p(o))
        assert int(o[0]) | (int(o[1]) << 64) == (lo | (hi << 64)) % P


def test_fp_extremes_and_wide_accumulator(ht):
    """fp_mul on the corner operands of the __int128 form (p - 1, 2^126, limbs of all ones, 0, 1) and the unreduced 320-bit
    multiply-accumulate that dec_value's edge stage and ct_mul's dense weight products use (fp_mac_wide + one fp_wide_reduce):
    equal to Python integers, including after enough accumulations of (p - 1)^2 to carry into the fifth limb"""
    M64 = 2**64 - 1
    corners = [0, 1, 2, P - 1, P - 2, 2**126, 2**126 - 1, 2**126 + 1, 2**64, 2**64 - 1, 2**64 + 1, (2**63 - 1) << 64, M64, (P - 1) ^ M64, 2**127 - 2**64,
               0x5555555555555555_5555555555555555, 0x2AAAAAAAAAAAAAAA_AAAAAAAAAAAAAAAA, 3 << 125, (1 << 63) | 1]
    corners = [c % P for c in corners]

    def words(v):
        return [v & M64, v >> 64]
    for a in corners:
        for b in corners:
            o = np.zeros(2, np.uint64)
            ht.ht_fp_op(2, _p(np.array(words(a), np.uint64)), _p(np.array(words(b), np.uint64)), _p(o))
            assert int(o[0]) | (int(o[1]) << 64) == a * b % P, (hex(a), hex(b))
    rng = np.random.default_rng(11)
    for trial in range(40):
        n = int(rng.integers(1, 700))
        xs = [int.from_bytes(rng.bytes(16), "little") % P for _ in range(n)]
        ys = [int.from_bytes(rng.bytes(16), "little") % P for _ in range(n)]
        if trial % 4 == 0:                                   # salt with corner values
            for k in range(0, n, 3):
                xs[k] = corners[(k + trial) % len(corners)]
                ys[k] = corners[(2 * k + trial) % len(corners)]
        a = np.array([w for v in xs for w in words(v)], np.uint64)
        b = np.array([w for v in ys for w in words(v)], np.uint64)
        o, wide = np.zeros(2, np.uint64), np.zeros(5, np.uint64)
        ht.ht_fp_mac_chain(n, _p(a), _p(b), 1, _p(o), _p(wide))
        exact = sum(x * y for x, y in zip(xs, ys))
        assert sum(int(wide[i]) << (64 * i) for i in range(5)) == exact          # the accumulator is the plain integer sum
        assert int(o[0]) | (int(o[1]) << 64) == exact % P
    # worst case: (p - 1)^2 accumulated 2^20 times (a 274-bit sum: the fifth limb is in use); a depth-3 ciphertext has 172 544 edges
    a = np.array(words(P - 1), np.uint64)
    o, wide = np.zeros(2, np.uint64), np.zeros(5, np.uint64)
    ht.ht_fp_mac_chain(1, _p(a), _p(a), 1 << 20, _p(o), _p(wide))
    exact = ((P - 1) ** 2) << 20
    assert sum(int(wide[i]) << (64 * i) for i in range(5)) == exact and int(wide[4]) != 0
    assert int(o[0]) | (int(o[1]) << 64) == exact % P


def test_tape_and_hash_layout(ht, port, kat):
    assert ht.ht_item_stream_state(1000, 3) == port.item_stream_state(1000, 3)
    st = np.array([12345], np.uint64)
    for k in range(5):                                                    # counter-based word k == sequential SplitMix64
        assert ht.ht_tape_word(12345, k) == port.lib().orc_tape_word(_p(st))
    doms = ["pvac.prf.r.1", "pvac.prf.r.2", "pvac.prf.r.3", "pvac.prf.noise.1", "pvac.prf.noise.2", "pvac.prf.noise.3"]
    for fam in (0, 1):
        for t in range(3):
            assert f"{ht.ht_fnv(fam, t):016x}" == kat["fnv1a"][doms[3 * fam + t]]
    assert f"{ht.ht_fnv(0, 3):016x}" == kat["fnv1a"]["pvac.dom.toeplitz"]
    assert f"{ht.ht_ztag(0x0123456789ABCDEF, 0x2222, 0x3333):016x}" == kat["ztag"]
    # sigma candidate stream: midstate + one compression per counter == SHA-256(label || words || ctr) of the oracle
    import hashlib
    import struct
    words = np.array([0x0123456789ABCDEF, 0x1111, 0x2222, 0x3333, 5, 1, 0xFEDCBA9876543210], np.uint64)
    for label, name in ((0, b"pvac.dom.x_seed"), (1, b"pvac.dom.noise")):
        for ctr in (0, 1, 33, 255, 256, 70000):
            o = np.zeros(4, np.uint64)
            ht.ht_cand_words(label, _p(words), ctr, _p(o))
            d = hashlib.sha256(name + words.tobytes() + struct.pack("<Q", ctr)).digest()
            assert o.tobytes() == d


def test_aes_and_prf_core(ht, port, kat, synth_keys_raw):
    key = np.frombuffer(bytes.fromhex(kat["aes_key"]), np.uint8).copy()
    o = np.zeros(6, np.uint64)
    ht.ht_aes_ctr_words(_p(key, u8), 0, _p(o), 3)
    assert hexwords(o) == kat["aes_ctr0_words"]
    o = np.zeros(8, np.uint64)
    ht.ht_aes_ctr_words(_p(key, u8), 2**64 - 2, _p(o), 4)
    assert hexwords(o) == kat["aes_ctr_wrap_words"]
    r = synth_keys_raw
    doms = ["pvac.prf.r.1", "pvac.prf.r.2", "pvac.prf.r.3", "pvac.prf.noise.1", "pvac.prf.noise.2", "pvac.prf.noise.3"]
    for fam in (0, 1):
        for t in range(3):
            out, y, rare = np.zeros(2, np.uint64), np.zeros(2, np.uint64), i32()
            ht.ht_prf_core(_p(r["prf_k"]), r["canon_tag"], _p(r["H_digest"], u8), _p(r["lpn_s"]), 0x1111, 0x2222, 0x3333, fam, t, 128, _p(y), _p(out), C.byref(rare))
            assert hexwords(out) == kat["prf_R_core"][doms[3 * fam + t]] and rare.value == 0
            if fam == 0 and t == 0:
                assert hexwords(y) == kat["ybits_r1_first4"][:2]
    # all 16384 rows through the same row-pair body (what PVACB_PRF_FAITHFUL executes)
    import hashlib
    y = np.zeros(256, np.uint64)
    out = np.zeros(2, np.uint64)
    ht.ht_prf_core(_p(r["prf_k"]), r["canon_tag"], _p(r["H_digest"], u8), _p(r["lpn_s"]), 0x1111, 0x2222, 0x3333, 0, 0, 16384, _p(y), _p(out), None)
    assert hashlib.sha256(y.tobytes()).hexdigest() == kat["ybits_r1_sha256"]


def test_keygen_host(ht):
    g = load_npz("keys_seed1.npz")
    words = 752 + 16384 * 128
    blob = np.zeros(words, np.uint64)
    assert ht.ht_keygen(1, _p(blob), words) == 0
    assert int(blob[0]) == int(g["canon_tag"])
    assert blob[1:5].tobytes() == g["H_digest"].tobytes()
    assert np.array_equal(blob[5:9], g["prf_k"]) and np.array_equal(blob[9:73], g["lpn_s"])
    assert np.array_equal(blob[73:747].reshape(337, 2), g["powg"])


def test_enc_plan_and_weights(ht, port, port_keys):
    """plan_item + share_weights (csrc/enc_plan.cuh) rebuild every (layer seed, idx, sign, weight) of the oracle's enc_value."""
    K = port_keys
    e = K.export(with_H=False)
    canon, powg = e["canon_tag"], np.ascontiguousarray(e["powg"])
    mr, mn = ht.ht_max_raw(), ht.ht_max_rnd()
    Z2, Z3 = K.plan_noise(0)
    for seed, v in ((1000, 42), (2000, 2**64 - 1), (31337, 0), (5, 1 << 63)):
        c = K.enc_value(seed, v)
        draws = port.tape_draws()
        d = port.ct_export(c, with_sigma=False)
        hdr, raw, rnd = np.zeros(14, np.uint64), np.zeros(2 * mr * 5, np.uint64), np.zeros(2 * mn * 2, np.uint64)
        used = ht.ht_plan_item(seed, v, canon, Z2, Z3, _p(hdr), _p(raw), _p(rnd))
        assert used == draws
        hdr, raw = hdr.reshape(2, 7), raw.reshape(2, mr, 5)
        # PRF values of each share from the oracle: prf_R, then prf_noise_delta(gid, kind) for all but the last group
        G = Z2 + Z3
        prf = np.zeros((2, G, 2), np.uint64)
        for s in range(2):
            z, lo, hi = int(hdr[s, 4]), int(hdr[s, 2]), int(hdr[s, 3])
            prf[s, 0] = K.prf_R(z, lo, hi)
            for gid in range(G - 1):
                prf[s, 1 + gid] = K.prf_noise_delta(z, lo, hi, gid, 0 if gid < Z2 else 1)
        wout = np.zeros((2, mr, 2), np.uint64)
        assert ht.ht_item_weights(seed, v, canon, Z2, Z3, _p(prf), _p(powg), _p(wout)) == 1
        e0 = 0
        for layer in range(2):          # combine_ciphers(enc(v+mask), enc(-mask)): layer 0 is the share planned SECOND
            s = 1 - layer
            n_raw, n_out = int(hdr[s, 5]), int(hdr[s, 6])
            assert (int(d["ztag"][layer]), int(d["nlo"][layer]), int(d["nhi"][layer])) == (int(hdr[s, 4]), int(hdr[s, 2]), int(hdr[s, 3]))
            for r in range(n_raw):
                idx, ch, pos = int(raw[s, r, 0]), int(raw[s, r, 1]), int(raw[s, r, 2])
                assert int(d["lid"][e0 + pos]) == layer and int(d["idx"][e0 + pos]) == idx and int(d["ch"][e0 + pos]) == ch
            assert np.array_equal(wout[s, :n_out], d["w"][e0:e0 + n_out])
            e0 += n_out
        assert e0 == len(d["lid"])


def test_bucket_table(ht, kat):
    for n, nb in kat["buckets"].items():
        assert ht.ht_next_bkt(int(n)) == nb == ht.ht_unordered_buckets_real(int(n))
    for n in (1, 2, 13, 1521, 1560, 1600, 3160, 48080, 100000, 1444804):
        assert ht.ht_next_bkt(n) == ht.ht_unordered_buckets_real(n)


def test_c_abi_program_compiles_as_c99(tmp_path):
    """include/pvacb.h is plain C: tests/c/abi_smoke.c builds with gcc -std=c99 -pedantic and links against libpvacb.so;
    without a GPU its first call fails with PVACB_E_CUDA (2) instead of falling back to anything."""
    import torch
    exe = str(tmp_path / "abi_smoke")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                        "-L", PKG, "-lpvacb", f"-Wl,-rpath,{PKG}", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if not torch.cuda.is_available():
        run = subprocess.run([exe], capture_output=True, text=True)
        assert run.returncode == 1 and "pvacb_ctx_create(0, &ctx) -> 2" in run.stderr


def test_cpp_programs_compile_and_link(tmp_path):
    """the C++ programs that the -m gpu tests build on the GPU box (tests/cpp/*.cpp over include/pvacb.hpp) compile and link here too,
    with warnings on: a change to the headers cannot break them unnoticed. Also instantiates the parts of the shim those programs do
    not use (host copies, Params, item base). Running them needs a GPU; here they must fail loudly instead."""
    pkg = PKG
    inc = os.path.join(os.path.dirname(PKG), "include")
    extra = tmp_path / "shim_all.cpp"
    extra.write_text("""
#include "pvacb.hpp"
int main() {
    try {
        pvacb::Engine eng(0);
        eng.keygen();
        auto a = eng.enc_value({1, 2, 3});
        pvacb::HostCiphers h = eng.to_host(a);
        auto b = eng.from_host(h);
        pvacb_params p = eng.get_params();
        eng.set_params(p);
        eng.set_item_base(0);
        eng.sync();
        pvacb::Group g({0});
        g.keygen();
        auto s = g.enc_value({4, 5}, g.fresh_seed());
        return (int)b.size() + (int)g.dec_value(s).size() - 5;
    } catch (const pvacb::Error& e) { return e.code == PVACB_E_CUDA ? 42 : 1; }
}
""")
    for src in (os.path.join(os.path.dirname(PKG), "tests", "cpp", "basic_usage_batched.cpp"),
                os.path.join(os.path.dirname(PKG), "tests", "cpp", "group_pipeline.cpp"), str(extra)):
        exe = str(tmp_path / (os.path.basename(src)[:-4]))
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", inc, src, "-L", pkg, "-lpvacb", f"-Wl,-rpath,{pkg}", "-o", exe],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([str(tmp_path / "shim_all")], capture_output=True, text=True)
        assert r.returncode == 42, (r.returncode, r.stderr[-500:])          # PVACB_E_CUDA from pvacb_ctx_create: no CPU path
