// TEST INFRASTRUCTURE (compiled only where the reference tree is present): the reference-shaped binding of INTEGRATION.md section 3,
// for real. Included AFTER <pvac/pvac.hpp>, it defines -- on the reference's own types (Params, PubKey, SecKey, Cipher, Fp) -- functions
// with the reference's exact signatures that run on the GPU through libpvacb.so (include/pvacb.h), plus std::vector overloads for
// batches. gpu_prelude.hpp then renames the reference's entry points to these, so that the reference's UNMODIFIED programs
// (examples/basic_usage.cpp) compile against the engine:
//     g++ -I/root/reference/include -Iinclude -include tests/cpp/ref_binding/gpu_prelude.hpp /root/reference/examples/basic_usage.cpp -lpvacb
#pragma once
#include <pvac/pvac.hpp>
#include <pvac/utils/text.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pvacb.h"

namespace pvac {
namespace gpu {

[[noreturn]] inline void die(pvacb_ctx* ctx, const char* what, int rc) {      // the reference's failure mode is std::abort() too
    std::fprintf(stderr, "[pvac::gpu] %s failed: status %d (%s)\n", what, rc, ctx ? pvacb_last_error(ctx) : "");
    std::abort();
}

// one context on GPU 0 for the process; keys are (re)loaded whenever a call arrives with another PubKey
struct Device {
    pvacb_ctx* ctx = nullptr;
    uint64_t loaded_tag = 0;
    bool have = false, have_sk = false;
    Device() {
        int rc = pvacb_ctx_create(0, &ctx);
        if (rc) die(nullptr, "pvacb_ctx_create (an sm_100-class GPU is required; there is no CPU path)", rc);
        if (const char* m = std::getenv("PVAC_GPU_PRF_LIVE")) { if (m[0] == '1') pvacb_set_prf_mode(ctx, PVACB_PRF_LIVE); }
    }
    ~Device() { pvacb_ctx_destroy(ctx); }
    static Device& get() { static Device d; return d; }

    void use(const PubKey& pk, const SecKey* sk) {
        if (have && loaded_tag == pk.canon_tag && (have_sk || !sk)) return;
        std::vector<uint64_t> H((size_t)pk.prm.n_bits * (pk.prm.m_bits / 64)), g(2 * (size_t)pk.prm.B);
        for (size_t c = 0; c < pk.H.size(); c++) std::memcpy(&H[c * (pk.prm.m_bits / 64)], pk.H[c].w.data(), pk.prm.m_bits / 8);
        for (size_t i = 0; i < pk.powg_B.size(); i++) { g[2 * i] = pk.powg_B[i].lo; g[2 * i + 1] = pk.powg_B[i].hi; }
        pvacb_params p;
        pvacb_params_default(&p);
        p.B = pk.prm.B; p.m_bits = pk.prm.m_bits; p.n_bits = pk.prm.n_bits; p.h_col_wt = pk.prm.h_col_wt; p.x_col_wt = pk.prm.x_col_wt; p.err_wt = pk.prm.err_wt;
        p.noise_entropy_bits = pk.prm.noise_entropy_bits; p.tuple2_fraction = pk.prm.tuple2_fraction; p.depth_slope_bits = pk.prm.depth_slope_bits;
        p.edge_budget = pk.prm.edge_budget; p.lpn_n = pk.prm.lpn_n; p.lpn_t = pk.prm.lpn_t; p.lpn_tau_num = pk.prm.lpn_tau_num; p.lpn_tau_den = pk.prm.lpn_tau_den;
        static const uint64_t zero_k[4] = {0, 0, 0, 0};
        static const std::vector<uint64_t> zero_s(64, 0);
        int rc = pvacb_keys_import_raw(ctx, pk.canon_tag, pk.H_digest.data(), H.data(), g.data(), sk ? sk->prf_k.data() : zero_k, sk ? sk->lpn_s_bits.data() : zero_s.data());
        if (rc) die(ctx, "pvacb_keys_import_raw", rc);
        const int live = pvacb_get_prf_mode(ctx);
        if ((rc = pvacb_set_params(ctx, &p))) die(ctx, "pvacb_set_params", rc);
        if (live == PVACB_PRF_LIVE) pvacb_set_prf_mode(ctx, PVACB_PRF_LIVE);
        loaded_tag = pk.canon_tag; have = true; have_sk = sk != nullptr;
    }
};

// std::vector<Cipher> (array of structs, one heap BitVec per edge) -> device batch (structure of arrays)
inline pvacb_batch* to_soa(Device& d, const std::vector<const Cipher*>& cs) {
    const size_t n = cs.size();
    std::vector<uint32_t> loff(n + 1, 0), eoff(n + 1, 0);
    for (size_t i = 0; i < n; i++) { loff[i + 1] = loff[i] + (uint32_t)cs[i]->L.size(); eoff[i + 1] = eoff[i] + (uint32_t)cs[i]->E.size(); }
    const size_t nL = loff[n], nE = eoff[n];
    std::vector<uint8_t> rule(nL + 1), ch(nE + 1);
    std::vector<uint64_t> ztag(nL + 1), nlo(nL + 1), nhi(nL + 1), w(2 * nE + 2), sigma(128 * nE + 1);
    std::vector<uint32_t> pa(nL + 1), pb(nL + 1), lid(nE + 1);
    std::vector<uint16_t> idx(nE + 1);
    for (size_t i = 0; i < n; i++) {
        for (size_t k = 0; k < cs[i]->L.size(); k++) {
            const Layer& L = cs[i]->L[k];
            const size_t o = loff[i] + k;
            rule[o] = (uint8_t)L.rule; ztag[o] = L.seed.ztag; nlo[o] = L.seed.nonce.lo; nhi[o] = L.seed.nonce.hi;
            pa[o] = L.rule == RRule::PROD ? L.pa : 0; pb[o] = L.rule == RRule::PROD ? L.pb : 0;
        }
        for (size_t k = 0; k < cs[i]->E.size(); k++) {
            const Edge& e = cs[i]->E[k];
            const size_t o = eoff[i] + k;
            lid[o] = e.layer_id; idx[o] = e.idx; ch[o] = e.ch; w[2 * o] = e.w.lo; w[2 * o + 1] = e.w.hi;
            std::memcpy(&sigma[128 * o], e.s.w.data(), 1024);
        }
    }
    pvacb_batch* b = nullptr;
    int rc = pvacb_batch_import_soa(d.ctx, n, loff.data(), eoff.data(), rule.data(), ztag.data(), nlo.data(), nhi.data(), pa.data(), pb.data(), lid.data(), idx.data(),
                                    ch.data(), w.data(), sigma.data(), &b);
    if (rc) die(d.ctx, "pvacb_batch_import_soa", rc);
    return b;
}

inline std::vector<Cipher> from_soa(Device& d, pvacb_batch* b) {
    const size_t n = pvacb_batch_count(b);
    uint64_t nL = 0, nE = 0;
    pvacb_batch_totals(b, &nL, &nE);
    std::vector<uint32_t> loff(n + 1), eoff(n + 1), pa(nL + 1), pb(nL + 1), lid(nE + 1);
    std::vector<uint8_t> rule(nL + 1), ch(nE + 1);
    std::vector<uint64_t> ztag(nL + 1), nlo(nL + 1), nhi(nL + 1), w(2 * nE + 2), sigma(128 * nE + 1);
    std::vector<uint16_t> idx(nE + 1);
    int rc = pvacb_batch_export_soa(d.ctx, b, loff.data(), eoff.data(), rule.data(), ztag.data(), nlo.data(), nhi.data(), pa.data(), pb.data(), lid.data(), idx.data(),
                                    ch.data(), w.data(), sigma.data());
    if (rc) die(d.ctx, "pvacb_batch_export_soa", rc);
    std::vector<Cipher> out(n);
    for (size_t i = 0; i < n; i++) {
        Cipher& C = out[i];
        C.L.resize(loff[i + 1] - loff[i]);
        C.E.resize(eoff[i + 1] - eoff[i]);
        for (size_t k = 0; k < C.L.size(); k++) {
            const size_t o = loff[i] + k;
            Layer L{};
            L.rule = (RRule)rule[o]; L.seed.ztag = ztag[o]; L.seed.nonce.lo = nlo[o]; L.seed.nonce.hi = nhi[o]; L.pa = pa[o]; L.pb = pb[o];
            C.L[k] = L;
        }
        for (size_t k = 0; k < C.E.size(); k++) {
            const size_t o = eoff[i] + k;
            Edge e{};
            e.layer_id = lid[o]; e.idx = idx[o]; e.ch = ch[o]; e.w = Fp{w[2 * o], w[2 * o + 1]};
            e.s = BitVec::make(8192);
            std::memcpy(e.s.w.data(), &sigma[128 * o], 1024);
            C.E[k] = std::move(e);
        }
    }
    return out;
}

inline std::vector<const Cipher*> ptrs(const std::vector<Cipher>& v) {
    std::vector<const Cipher*> p;
    for (const Cipher& c : v) p.push_back(&c);
    return p;
}

// ---- batched forms: an array of independent items per call
inline std::vector<Cipher> enc_value(const PubKey& pk, const SecKey& sk, const std::vector<uint64_t>& v) {
    Device& d = Device::get();
    d.use(pk, &sk);
    pvacb_batch* b = nullptr;
    int rc = pvacb_enc_value(d.ctx, v.data(), v.size(), pvacb_fresh_seed(d.ctx), &b);       // fresh ChaCha20 streams: no seed in the reference's signature
    if (rc) die(d.ctx, "pvacb_enc_value", rc);
    auto out = from_soa(d, b);
    pvacb_batch_free(b);
    return out;
}
inline std::vector<Cipher> binop(const PubKey& pk, const std::vector<const Cipher*>& A, const std::vector<const Cipher*>& B, int op) {
    Device& d = Device::get();
    d.use(pk, nullptr);
    pvacb_batch *a = to_soa(d, A), *b = to_soa(d, B), *o = nullptr;
    int rc = op == 0 ? pvacb_ct_add(d.ctx, a, b, &o) : op == 1 ? pvacb_ct_sub(d.ctx, a, b, &o) : pvacb_ct_mul(d.ctx, a, b, pvacb_fresh_seed(d.ctx), &o);
    pvacb_batch_free(a); pvacb_batch_free(b);
    if (rc == PVACB_E_SHAPE && std::getenv("PVAC_GPU_SOFT_SHAPE")) {       // test harness: report, hand back empty ciphertexts, let the program go on
        std::fprintf(stderr, "[pvac::gpu] operand too large for the batched path: %s\n", pvacb_last_error(d.ctx));
        return std::vector<Cipher>(A.size());
    }
    if (rc) die(d.ctx, op == 0 ? "pvacb_ct_add" : op == 1 ? "pvacb_ct_sub" : "pvacb_ct_mul", rc);
    auto out = from_soa(d, o);
    pvacb_batch_free(o);
    return out;
}
inline std::vector<Cipher> ct_add(const PubKey& pk, const std::vector<Cipher>& A, const std::vector<Cipher>& B) { return binop(pk, ptrs(A), ptrs(B), 0); }
inline std::vector<Cipher> ct_sub(const PubKey& pk, const std::vector<Cipher>& A, const std::vector<Cipher>& B) { return binop(pk, ptrs(A), ptrs(B), 1); }
inline std::vector<Cipher> ct_mul(const PubKey& pk, const std::vector<Cipher>& A, const std::vector<Cipher>& B) { return binop(pk, ptrs(A), ptrs(B), 2); }
inline std::vector<Fp> dec_value(const PubKey& pk, const SecKey& sk, const std::vector<const Cipher*>& C) {
    Device& d = Device::get();
    d.use(pk, &sk);
    pvacb_batch* c = to_soa(d, C);
    std::vector<uint64_t> o(2 * C.size() + 2);
    int rc = pvacb_dec_value(d.ctx, c, o.data());
    pvacb_batch_free(c);
    if (rc) die(d.ctx, "pvacb_dec_value", rc);
    std::vector<Fp> out(C.size());
    for (size_t i = 0; i < C.size(); i++) out[i] = Fp{o[2 * i], o[2 * i + 1]};
    return out;
}

}  // namespace gpu

// ---- the reference's own signatures (crypto/keygen.hpp:35, ops/encrypt.hpp:289, ops/arithmetic.hpp:12,43,47, ops/decrypt.hpp:62,
// ops/commit.hpp:12, utils/text.hpp:39,63), executed on the GPU
inline void gpu_keygen(const Params& prm, PubKey& pk, SecKey& sk) {
    gpu::Device& d = gpu::Device::get();
    pvacb_params p;
    pvacb_params_default(&p);
    p.B = prm.B; p.m_bits = prm.m_bits; p.n_bits = prm.n_bits; p.h_col_wt = prm.h_col_wt; p.x_col_wt = prm.x_col_wt; p.err_wt = prm.err_wt;
    p.noise_entropy_bits = prm.noise_entropy_bits; p.tuple2_fraction = prm.tuple2_fraction; p.depth_slope_bits = prm.depth_slope_bits; p.edge_budget = prm.edge_budget;
    p.lpn_n = prm.lpn_n; p.lpn_t = prm.lpn_t; p.lpn_tau_num = prm.lpn_tau_num; p.lpn_tau_den = prm.lpn_tau_den;
    p.recrypt_lo = prm.recrypt_lo; p.recrypt_hi = prm.recrypt_hi; p.recrypt_rounds = prm.recrypt_rounds;
    const int mode = pvacb_get_prf_mode(d.ctx);
    int rc = pvacb_keygen_params(d.ctx, &p, nullptr);                    // seed from the OS CSPRNG, like the reference's csprng_u64()
    if (rc) gpu::die(d.ctx, "pvacb_keygen_params", rc);
    if (mode == PVACB_PRF_LIVE) pvacb_set_prf_mode(d.ctx, PVACB_PRF_LIVE);
    pk.prm = prm;
    std::vector<uint64_t> H((size_t)prm.n_bits * (prm.m_bits / 64)), g(2 * (size_t)prm.B), s(64);
    sk.lpn_s_bits.assign(64, 0);
    if ((rc = pvacb_keys_export_raw(d.ctx, &pk.canon_tag, pk.H_digest.data(), H.data(), g.data(), sk.prf_k.data(), sk.lpn_s_bits.data()))) gpu::die(d.ctx, "pvacb_keys_export_raw", rc);
    pk.H.assign(prm.n_bits, BitVec::make(prm.m_bits));
    for (int c = 0; c < prm.n_bits; c++) std::memcpy(pk.H[c].w.data(), &H[(size_t)c * (prm.m_bits / 64)], prm.m_bits / 8);
    pk.powg_B.resize(prm.B);
    for (int i = 0; i < prm.B; i++) pk.powg_B[i] = Fp{g[2 * i], g[2 * i + 1]};
    pk.ubk = gen_ubk_public(pk.canon_tag, prm.m_bits);
    pk.omega_B = fp_from_u64(1);          // never read by any operation (the engine keeps the real value for pk files)
    d.loaded_tag = pk.canon_tag; d.have = true; d.have_sk = true;
}
inline Cipher gpu_enc_value(const PubKey& pk, const SecKey& sk, uint64_t v) { return std::move(gpu::enc_value(pk, sk, std::vector<uint64_t>{v})[0]); }
inline Cipher gpu_ct_add(const PubKey& pk, const Cipher& a, const Cipher& b) { return std::move(gpu::binop(pk, {&a}, {&b}, 0)[0]); }
inline Cipher gpu_ct_sub(const PubKey& pk, const Cipher& a, const Cipher& b) { return std::move(gpu::binop(pk, {&a}, {&b}, 1)[0]); }
inline Cipher gpu_ct_mul(const PubKey& pk, const Cipher& a, const Cipher& b) { return std::move(gpu::binop(pk, {&a}, {&b}, 2)[0]); }
inline Fp gpu_dec_value(const PubKey& pk, const SecKey& sk, const Cipher& c) { return gpu::dec_value(pk, sk, {&c})[0]; }
inline std::array<uint8_t, 32> gpu_commit_ct(const PubKey& pk, const Cipher& c) {
    gpu::Device& d = gpu::Device::get();
    d.use(pk, nullptr);
    pvacb_batch* b = gpu::to_soa(d, {&c});
    std::array<uint8_t, 32> o{};
    int rc = pvacb_commit_ct(d.ctx, b, o.data());
    pvacb_batch_free(b);
    if (rc) gpu::die(d.ctx, "pvacb_commit_ct", rc);
    return o;
}
// ---- riders: ops/arithmetic.hpp:33,39,108, ops/encrypt.hpp:29,39,162,281,293, ops/recrypt.hpp:12,26, crypto/matrix.hpp:306
namespace gpu {
// one ciphertext in, one out, through a batch -> batch entry point of the C ABI
template <class F>
inline Cipher unop(const PubKey& pk, const Cipher& a, const char* what, F&& call) {
    Device& d = Device::get();
    d.use(pk, nullptr);
    pvacb_batch *x = to_soa(d, {&a}), *o = nullptr;
    int rc = call(d.ctx, x, &o);
    pvacb_batch_free(x);
    if (rc) die(d.ctx, what, rc);
    auto out = from_soa(d, o);
    pvacb_batch_free(o);
    return std::move(out[0]);
}
}  // namespace gpu
inline Cipher gpu_ct_scale(const PubKey& pk, const Cipher& a, const Fp& s) {
    const uint64_t k[2] = {s.lo, s.hi};
    return gpu::unop(pk, a, "pvacb_ct_scale", [&](pvacb_ctx* c, pvacb_batch* x, pvacb_batch** o) { return pvacb_ct_scale(c, x, k, o); });
}
inline Cipher gpu_ct_neg(const PubKey& pk, const Cipher& a) {
    return gpu::unop(pk, a, "pvacb_ct_neg", [&](pvacb_ctx* c, pvacb_batch* x, pvacb_batch** o) { return pvacb_ct_neg(c, x, o); });
}
inline Cipher gpu_ct_div_const(const PubKey& pk, const Cipher& a, const Fp& k) {
    const uint64_t kk[2] = {k.lo, k.hi};
    return gpu::unop(pk, a, "pvacb_ct_div_const", [&](pvacb_ctx* c, pvacb_batch* x, pvacb_batch** o) { return pvacb_ct_div_const(c, x, kk, o); });
}
inline void gpu_compact_edges(const PubKey& pk, Cipher& C) {
    C = gpu::unop(pk, C, "pvacb_compact_edges", [&](pvacb_ctx* c, pvacb_batch* x, pvacb_batch** o) { return pvacb_compact_edges(c, x, o); });
}
inline void gpu_ubk_apply(const PubKey& pk, Cipher& C) {
    C = gpu::unop(pk, C, "pvacb_ubk_apply", [&](pvacb_ctx* c, pvacb_batch* x, pvacb_batch** o) { return pvacb_ubk_apply(c, x, o); });
}
inline double gpu_sigma_density(const PubKey& pk, const Cipher& C) {
    gpu::Device& d = gpu::Device::get();
    d.use(pk, nullptr);
    pvacb_batch* b = gpu::to_soa(d, {&C});
    double out = 0.0;
    int rc = pvacb_sigma_density(d.ctx, b, &out);
    pvacb_batch_free(b);
    if (rc) gpu::die(d.ctx, "pvacb_sigma_density", rc);
    return out;
}
inline Cipher gpu_enc_value_depth(const PubKey& pk, const SecKey& sk, uint64_t v, int depth_hint) {
    gpu::Device& d = gpu::Device::get();
    d.use(pk, &sk);
    pvacb_batch* b = nullptr;
    int rc = pvacb_enc_value_depth(d.ctx, &v, 1, depth_hint, pvacb_fresh_seed(d.ctx), nullptr, &b);
    if (rc) gpu::die(d.ctx, "pvacb_enc_value_depth", rc);
    auto out = gpu::from_soa(d, b);
    pvacb_batch_free(b);
    return std::move(out[0]);
}
inline Cipher gpu_enc_fp_depth(const PubKey& pk, const SecKey& sk, const Fp& v, int depth_hint) {
    gpu::Device& d = gpu::Device::get();
    d.use(pk, &sk);
    const uint64_t fv[2] = {v.lo, v.hi};
    pvacb_batch* b = nullptr;
    int rc = pvacb_enc_fp_depth(d.ctx, fv, 1, depth_hint, pvacb_fresh_seed(d.ctx), nullptr, &b);
    if (rc) gpu::die(d.ctx, "pvacb_enc_fp_depth", rc);
    auto out = gpu::from_soa(d, b);
    pvacb_batch_free(b);
    return std::move(out[0]);
}
// n encryptions of zero in ONE call (what make_evalkey's loop asks for)
inline std::vector<Cipher> gpu_enc_zero_depth_n(const PubKey& pk, const SecKey& sk, size_t n, int depth_hint) {
    gpu::Device& d = gpu::Device::get();
    d.use(pk, &sk);
    pvacb_batch* b = nullptr;
    int rc = pvacb_enc_zero_depth(d.ctx, n, depth_hint, pvacb_fresh_seed(d.ctx), nullptr, &b);
    if (rc) gpu::die(d.ctx, "pvacb_enc_zero_depth", rc);
    auto out = gpu::from_soa(d, b);
    pvacb_batch_free(b);
    return out;
}
inline Cipher gpu_enc_zero_depth(const PubKey& pk, const SecKey& sk, int depth_hint) { return std::move(gpu_enc_zero_depth_n(pk, sk, 1, depth_hint)[0]); }
inline EvalKey gpu_make_evalkey(const PubKey& pk, const SecKey& sk, size_t pool_size, int depth_hint) {
    EvalKey ek;
    ek.zero_pool = gpu_enc_zero_depth_n(pk, sk, pool_size, depth_hint);
    ek.enc_one = std::move(gpu::enc_value(pk, sk, std::vector<uint64_t>{1})[0]);
    return ek;
}
inline Cipher gpu_ct_recrypt(const PubKey& pk, const EvalKey& ek, const Cipher& in) {
    if (ek.zero_pool.empty() || in.E.empty()) return in;      // ops/recrypt.hpp:27
    gpu::Device& d = gpu::Device::get();
    d.use(pk, nullptr);
    pvacb_batch *c = gpu::to_soa(d, {&in}), *pool = gpu::to_soa(d, gpu::ptrs(ek.zero_pool)), *o = nullptr;
    int rc = pvacb_ct_recrypt(d.ctx, c, pool, pvacb_fresh_seed(d.ctx), nullptr, &o);
    pvacb_batch_free(c); pvacb_batch_free(pool);
    if (rc) gpu::die(d.ctx, "pvacb_ct_recrypt", rc);
    auto out = gpu::from_soa(d, o);
    pvacb_batch_free(o);
    return std::move(out[0]);
}

inline std::vector<Cipher> gpu_enc_text(const PubKey& pk, const SecKey& sk, const std::string& msg) {
    gpu::Device& d = gpu::Device::get();
    d.use(pk, &sk);
    const uint64_t off[2] = {0, msg.size()};
    pvacb_batch* b = nullptr;
    int rc = pvacb_enc_text(d.ctx, reinterpret_cast<const uint8_t*>(msg.data()), off, 1, pvacb_fresh_seed(d.ctx), nullptr, &b);
    if (rc) gpu::die(d.ctx, "pvacb_enc_text", rc);
    auto out = gpu::from_soa(d, b);      // one message: wave-major order = the reference's order (length, block 0, block 1, ...)
    pvacb_batch_free(b);
    return out;
}
inline std::string gpu_dec_text(const PubKey& pk, const SecKey& sk, const std::vector<Cipher>& cts) {
    gpu::Device& d = gpu::Device::get();
    d.use(pk, &sk);
    if (cts.empty()) return std::string();
    pvacb_batch* b = gpu::to_soa(d, gpu::ptrs(cts));
    std::vector<uint8_t> buf(15 * cts.size() + 16);
    uint64_t off[2] = {0, 0};
    int rc = pvacb_dec_text(d.ctx, b, 1, buf.data(), buf.size(), off);
    pvacb_batch_free(b);
    if (rc) gpu::die(d.ctx, "pvacb_dec_text", rc);
    return std::string(reinterpret_cast<const char*>(buf.data()), (size_t)off[1]);
}

}  // namespace pvac
