// pvacb.hpp -- header-only C++17 convenience layer over the C ABI (pvacb.h), shaped like the reference's API so that a
// program written against include/pvac/pvac.hpp reads the same with batches in place of single ciphertexts:
//
//   reference (one ciphertext, CPU)                     batched (arrays of independent ciphertexts, B200)
//   -----------------------------------------------     ------------------------------------------------------------
//   Params prm; PubKey pk; SecKey sk;                   pvacb::Engine eng(/*device*/ 0);
//   keygen(prm, pk, sk);              keygen.hpp:35     eng.keygen(tape_state);
//   Cipher a = enc_value(pk, sk, 42); encrypt.hpp:289   pvacb::Ciphers a = eng.enc_value({42, 7, ...}, seed);
//   Cipher s = ct_add(pk, a, b);      arithmetic.hpp:12 pvacb::Ciphers s = eng.ct_add(a, b);
//   Cipher d = ct_sub(pk, a, b);      arithmetic.hpp:43 pvacb::Ciphers d = eng.ct_sub(a, b);
//   Cipher p = ct_mul(pk, a, b);      arithmetic.hpp:47 pvacb::Ciphers p = eng.ct_mul(a, b, seed);
//   Fp v = dec_value(pk, sk, p);      decrypt.hpp:62    std::vector<pvacb::Fp> v = eng.dec_value(p);
//
// Where the reference calls std::abort() these throw pvacb::Error carrying the PVACB_E_* status. Nothing here computes on
// the CPU: every call forwards to libpvacb.so (sm_100a kernels).
#ifndef PVACB_HPP
#define PVACB_HPP

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pvacb.h"

namespace pvacb {

struct Fp {            // core/field.hpp:17-20
    uint64_t lo, hi;
    bool operator==(const Fp& o) const { return lo == o.lo && hi == o.hi; }
};

class Error : public std::runtime_error {
public:
    Error(int code, const std::string& what) : std::runtime_error("pvacb status " + std::to_string(code) + ": " + what), code(code) {}
    int code;
};

class Engine;

// A device-resident array of ciphertexts (move-only; frees the batch on destruction).
class Ciphers {
public:
    Ciphers() = default;
    explicit Ciphers(pvacb_batch* h) : h_(h) {}
    Ciphers(Ciphers&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Ciphers& operator=(Ciphers&& o) noexcept { if (this != &o) { reset(); h_ = o.h_; o.h_ = nullptr; } return *this; }
    Ciphers(const Ciphers&) = delete;
    Ciphers& operator=(const Ciphers&) = delete;
    ~Ciphers() { reset(); }
    void reset() { if (h_) pvacb_batch_free(h_); h_ = nullptr; }
    size_t size() const { return h_ ? pvacb_batch_count(h_) : 0; }
    size_t device_bytes() const { return h_ ? pvacb_batch_device_bytes(h_) : 0; }
    std::pair<uint64_t, uint64_t> totals() const { uint64_t l = 0, e = 0; if (h_) pvacb_batch_totals(h_, &l, &e); return {l, e}; }
    pvacb_batch* handle() const { return h_; }
private:
    pvacb_batch* h_ = nullptr;
};

// A batch copied to host memory, structure of arrays: the fields of pvac::Cipher (core/types.hpp:96-119) for n ciphertexts, ciphertext i
// owning layers [layer_off[i], layer_off[i+1]) and edges [edge_off[i], edge_off[i+1]); layer_id is relative to its ciphertext.
struct HostCiphers {
    std::vector<uint32_t> layer_off, edge_off;                                   // n + 1 entries each
    std::vector<uint8_t> rule;                                                   // layers: 0 BASE, 1 PROD
    std::vector<uint64_t> ztag, nonce_lo, nonce_hi;                              //   RSeed of a BASE layer (and of a PROD layer: its sigma seed)
    std::vector<uint32_t> pa, pb;                                                //   parents of a PROD layer
    std::vector<uint32_t> layer_id;                                              // edges
    std::vector<uint16_t> idx;
    std::vector<uint8_t> ch;
    std::vector<uint64_t> w;                                                     //   2 words per edge (lo, hi)
    std::vector<uint64_t> sigma;                                                 //   128 words per edge (empty if exported without sigma)
    size_t size() const { return layer_off.empty() ? 0 : layer_off.size() - 1; }
};

class Engine {
public:
    explicit Engine(int device = 0, int prf_mode = PVACB_PRF_FAITHFUL) {
        int rc = pvacb_ctx_create(device, &ctx_);
        if (rc) throw Error(rc, "pvacb_ctx_create failed (no sm_100-class GPU? there is no CPU fallback)");
        pvacb_set_prf_mode(ctx_, prf_mode);
    }
    ~Engine() { if (ctx_) pvacb_ctx_destroy(ctx_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    // keygen(prm, pk, sk), crypto/keygen.hpp:35: default Params unless given, seed = 32 bytes or nullptr for the OS CSPRNG
    void keygen(const pvacb_params* prm = nullptr, const uint8_t* seed32 = nullptr) { ck(pvacb_keygen_params(ctx_, prm, seed32)); }
    void keygen(uint64_t tape_state) { ck(pvacb_keygen(ctx_, tape_state)); }       // parity vectors only (SplitMix64 from 64 bits)
    void set_prf_mode(int mode) { ck(pvacb_set_prf_mode(ctx_, mode)); }
    // RNG tape of the context (pvacb.h): ChaCha20 under an OS key by default
    void set_tape(int kind, const uint8_t* key32 = nullptr) { ck(pvacb_set_tape(ctx_, kind, key32)); }
    uint64_t fresh_seed() { return pvacb_fresh_seed(ctx_); }
    // savePk / loadPk / saveSk / loadSk of the reference's programs (tests/bounty2_test.cpp:145-236)
    void save_keys(const char* pk_path, const char* sk_path) { ck(pvacb_keys_export_file(ctx_, pk_path, sk_path)); }
    void load_keys(const char* pk_path, const char* sk_path) { ck(pvacb_keys_import_file(ctx_, pk_path, sk_path)); }
    // the reference's signatures carry no seed: these overloads draw a fresh one from the context
    Ciphers enc_value(const std::vector<uint64_t>& v) { return enc_value(v, fresh_seed()); }
    Ciphers ct_mul(const Ciphers& a, const Ciphers& b) { return ct_mul(a, b, fresh_seed()); }

    Ciphers enc_value(const std::vector<uint64_t>& v, uint64_t batch_seed) {
        pvacb_batch* o = nullptr;
        ck(pvacb_enc_value(ctx_, v.data(), v.size(), batch_seed, &o));
        return Ciphers(o);
    }
    Ciphers ct_add(const Ciphers& a, const Ciphers& b) { pvacb_batch* o = nullptr; ck(pvacb_ct_add(ctx_, a.handle(), b.handle(), &o)); return Ciphers(o); }
    Ciphers ct_sub(const Ciphers& a, const Ciphers& b) { pvacb_batch* o = nullptr; ck(pvacb_ct_sub(ctx_, a.handle(), b.handle(), &o)); return Ciphers(o); }
    Ciphers ct_scale(const Ciphers& a, Fp s) { pvacb_batch* o = nullptr; uint64_t w[2] = {s.lo, s.hi}; ck(pvacb_ct_scale(ctx_, a.handle(), w, &o)); return Ciphers(o); }
    Ciphers ct_mul(const Ciphers& a, const Ciphers& b, uint64_t batch_seed) {
        pvacb_batch* o = nullptr;
        ck(pvacb_ct_mul(ctx_, a.handle(), b.handle(), batch_seed, &o));
        return Ciphers(o);
    }
    std::vector<Fp> dec_value(const Ciphers& c) {
        std::vector<Fp> out(c.size());
        static_assert(sizeof(Fp) == 16, "Fp is two u64 limbs");
        ck(pvacb_dec_value(ctx_, c.handle(), reinterpret_cast<uint64_t*>(out.data())));
        return out;
    }
    Ciphers ct_neg(const Ciphers& a) { pvacb_batch* o = nullptr; ck(pvacb_ct_neg(ctx_, a.handle(), &o)); return Ciphers(o); }
    Ciphers ct_div_const(const Ciphers& a, Fp k) { pvacb_batch* o = nullptr; uint64_t w[2] = {k.lo, k.hi}; ck(pvacb_ct_div_const(ctx_, a.handle(), w, &o)); return Ciphers(o); }
    Ciphers compact_edges(const Ciphers& a) { pvacb_batch* o = nullptr; ck(pvacb_compact_edges(ctx_, a.handle(), &o)); return Ciphers(o); }
    Ciphers enc_value_depth(const std::vector<uint64_t>& v, int depth_hint, uint64_t batch_seed) {
        pvacb_batch* o = nullptr;
        ck(pvacb_enc_value_depth(ctx_, v.data(), v.size(), depth_hint, batch_seed, nullptr, &o));
        return Ciphers(o);
    }
    // commit_ct (ops/commit.hpp:12): one 32-byte digest per ciphertext
    std::vector<std::array<uint8_t, 32>> commit_ct(const Ciphers& c) {
        std::vector<std::array<uint8_t, 32>> out(c.size());
        ck(pvacb_commit_ct(ctx_, c.handle(), reinterpret_cast<uint8_t*>(out.data())));
        return out;
    }
    // enc_text / dec_text (utils/text.hpp:39,63) for several messages at once; the batch is wave-major (see pvacb.h)
    Ciphers enc_text(const std::vector<std::string>& msgs, uint64_t batch_seed) {
        std::vector<uint64_t> off(msgs.size() + 1, 0);
        std::string flat;
        for (size_t i = 0; i < msgs.size(); i++) { flat += msgs[i]; off[i + 1] = flat.size(); }
        pvacb_batch* o = nullptr;
        ck(pvacb_enc_text(ctx_, reinterpret_cast<const uint8_t*>(flat.data()), off.data(), msgs.size(), batch_seed, nullptr, &o));
        return Ciphers(o);
    }
    std::vector<std::string> dec_text(const Ciphers& c, size_t n_msgs) {
        std::vector<uint8_t> buf(15 * c.size() + 16);
        std::vector<uint64_t> off(n_msgs + 1, 0);
        ck(pvacb_dec_text(ctx_, c.handle(), n_msgs, buf.data(), buf.size(), off.data()));
        std::vector<std::string> out(n_msgs);
        for (size_t i = 0; i < n_msgs; i++) out[i].assign(reinterpret_cast<const char*>(buf.data()) + off[i], (size_t)(off[i + 1] - off[i]));
        return out;
    }
    // riders (ops/encrypt.hpp:162,293): one-share encryptions of field elements, encryptions of zero
    Ciphers enc_fp_depth(const std::vector<Fp>& v, int depth_hint, uint64_t batch_seed) {
        pvacb_batch* o = nullptr;
        ck(pvacb_enc_fp_depth(ctx_, reinterpret_cast<const uint64_t*>(v.data()), v.size(), depth_hint, batch_seed, nullptr, &o));
        return Ciphers(o);
    }
    Ciphers enc_zero_depth(size_t n, int depth_hint, uint64_t batch_seed) {
        pvacb_batch* o = nullptr;
        ck(pvacb_enc_zero_depth(ctx_, n, depth_hint, batch_seed, nullptr, &o));
        return Ciphers(o);
    }
    // ct_recrypt (ops/recrypt.hpp:26) with EvalKey::zero_pool held as a batch; sigma_density (ops/encrypt.hpp:29); ubk_apply (crypto/matrix.hpp:306)
    Ciphers ct_recrypt(const Ciphers& c, const Ciphers& zero_pool, uint64_t batch_seed) {
        pvacb_batch* o = nullptr;
        ck(pvacb_ct_recrypt(ctx_, c.handle(), zero_pool.handle(), batch_seed, nullptr, &o));
        return Ciphers(o);
    }
    std::vector<double> sigma_density(const Ciphers& c) {
        std::vector<double> d(c.size());
        ck(pvacb_sigma_density(ctx_, c.handle(), d.data()));
        return d;
    }
    Ciphers ubk_apply(const Ciphers& c) { pvacb_batch* o = nullptr; ck(pvacb_ubk_apply(ctx_, c.handle(), &o)); return Ciphers(o); }
    // batch plumbing
    Ciphers slice(const Ciphers& c, size_t first, size_t count) { pvacb_batch* o = nullptr; ck(pvacb_batch_slice(ctx_, c.handle(), first, count, &o)); return Ciphers(o); }
    Ciphers concat(const std::vector<const Ciphers*>& parts) {
        std::vector<const pvacb_batch*> h;
        for (const Ciphers* p : parts) h.push_back(p->handle());
        pvacb_batch* o = nullptr;
        ck(pvacb_batch_concat(ctx_, h.data(), h.size(), &o));
        return Ciphers(o);
    }
    // the reference's on-disk format (tests/bounty2_test.cpp:63-143)
    std::vector<uint8_t> to_wire(const Ciphers& c) {
        size_t n = 0, w = 0;
        ck(pvacb_batch_wire_size(ctx_, c.handle(), &n));
        std::vector<uint8_t> buf(n);
        ck(pvacb_batch_export_wire(ctx_, c.handle(), buf.data(), buf.size(), &w));
        buf.resize(w);
        return buf;
    }
    Ciphers from_wire(const std::vector<uint8_t>& buf) { pvacb_batch* o = nullptr; ck(pvacb_batch_import_wire(ctx_, buf.data(), buf.size(), &o)); return Ciphers(o); }

    // Cipher <-> host memory (pvacb_batch_export_soa / import_soa): what a program does where the reference hands it a Cipher by value
    HostCiphers to_host(const Ciphers& c, bool with_sigma = true) {
        HostCiphers h;
        const size_t n = c.size();
        const auto t = c.totals();
        h.layer_off.assign(n + 1, 0); h.edge_off.assign(n + 1, 0);
        if (!c.handle()) return h;
        h.rule.resize(t.first); h.ztag.resize(t.first); h.nonce_lo.resize(t.first); h.nonce_hi.resize(t.first); h.pa.resize(t.first); h.pb.resize(t.first);
        h.layer_id.resize(t.second); h.idx.resize(t.second); h.ch.resize(t.second); h.w.resize(2 * t.second);
        if (with_sigma) h.sigma.resize((size_t)PVACB_M_WORDS * t.second);
        ck(pvacb_batch_export_soa(ctx_, c.handle(), h.layer_off.data(), h.edge_off.data(), h.rule.data(), h.ztag.data(), h.nonce_lo.data(), h.nonce_hi.data(),
                                  h.pa.data(), h.pb.data(), h.layer_id.data(), h.idx.data(), h.ch.data(), h.w.data(), with_sigma ? h.sigma.data() : nullptr));
        return h;
    }
    Ciphers from_host(const HostCiphers& h) {
        pvacb_batch* o = nullptr;
        ck(pvacb_batch_import_soa(ctx_, h.size(), h.layer_off.data(), h.edge_off.data(), h.rule.data(), h.ztag.data(), h.nonce_lo.data(), h.nonce_hi.data(),
                                  h.pa.data(), h.pb.data(), h.layer_id.data(), h.idx.data(), h.ch.data(), h.w.data(), h.sigma.empty() ? nullptr : h.sigma.data(), &o));
        return Ciphers(o);
    }
    // run-time Params of the context (core/types.hpp:36-70), the global index of item 0 of the next calls (a shard of a larger batch
    // keeps the RNG streams of the whole batch), and a barrier on the context's stream
    void set_params(const pvacb_params& prm) { ck(pvacb_set_params(ctx_, &prm)); }
    pvacb_params get_params() const { pvacb_params p; pvacb_get_params(ctx_, &p); return p; }
    void set_item_base(uint64_t base) { ck(pvacb_set_item_base(ctx_, base)); }
    void sync() { ck(pvacb_sync(ctx_)); }

    pvacb_ctx* handle() const { return ctx_; }

private:
    void ck(int rc) { if (rc) throw Error(rc, pvacb_last_error(ctx_)); }
    pvacb_ctx* ctx_ = nullptr;
};

// Several GPUs of one box behind one object (pvacb_group_*): batches are split into contiguous index ranges, every item keeps the RNG
// stream of its global index (the bytes do not depend on the number of GPUs), keys are replicated once over NVLink.
class ShardedCiphers {
public:
    ShardedCiphers() = default;
    explicit ShardedCiphers(pvacb_gbatch* h) : h_(h) {}
    ShardedCiphers(ShardedCiphers&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    ShardedCiphers& operator=(ShardedCiphers&& o) noexcept { if (this != &o) { reset(); h_ = o.h_; o.h_ = nullptr; } return *this; }
    ShardedCiphers(const ShardedCiphers&) = delete;
    ShardedCiphers& operator=(const ShardedCiphers&) = delete;
    ~ShardedCiphers() { reset(); }
    void reset() { if (h_) pvacb_group_batch_free(h_); h_ = nullptr; }
    size_t size() const { return h_ ? pvacb_group_batch_count(h_) : 0; }
    pvacb_gbatch* handle() const { return h_; }
private:
    pvacb_gbatch* h_ = nullptr;
};

class Group {
public:
    explicit Group(const std::vector<int>& devices, int prf_mode = PVACB_PRF_FAITHFUL) {
        int rc = pvacb_group_create(devices.data(), (int)devices.size(), &g_);
        if (rc) throw Error(rc, "pvacb_group_create failed");
        for (int k = 0; k < size(); k++) pvacb_set_prf_mode(pvacb_group_ctx(g_, k), prf_mode);
    }
    ~Group() { if (g_) pvacb_group_destroy(g_); }
    Group(const Group&) = delete;
    Group& operator=(const Group&) = delete;
    int size() const { return pvacb_group_size(g_); }
    void keygen(const pvacb_params* prm = nullptr, const uint8_t* seed32 = nullptr) { ck(pvacb_group_keygen_params(g_, prm, seed32)); }
    void load_keys(const char* pk_path, const char* sk_path) { ck(pvacb_group_keys_import_file(g_, pk_path, sk_path)); }
    void set_tape(int kind, const uint8_t* key32 = nullptr) { ck(pvacb_group_set_tape(g_, kind, key32)); }
    uint64_t fresh_seed() { return pvacb_fresh_seed(pvacb_group_ctx(g_, 0)); }
    ShardedCiphers enc_value(const std::vector<uint64_t>& v, uint64_t batch_seed) { pvacb_gbatch* o = nullptr; ck(pvacb_group_enc_value(g_, v.data(), v.size(), batch_seed, &o)); return ShardedCiphers(o); }
    ShardedCiphers ct_add(const ShardedCiphers& a, const ShardedCiphers& b) { pvacb_gbatch* o = nullptr; ck(pvacb_group_ct_add(g_, a.handle(), b.handle(), &o)); return ShardedCiphers(o); }
    ShardedCiphers ct_sub(const ShardedCiphers& a, const ShardedCiphers& b) { pvacb_gbatch* o = nullptr; ck(pvacb_group_ct_sub(g_, a.handle(), b.handle(), &o)); return ShardedCiphers(o); }
    ShardedCiphers ct_mul(const ShardedCiphers& a, const ShardedCiphers& b, uint64_t batch_seed) { pvacb_gbatch* o = nullptr; ck(pvacb_group_ct_mul(g_, a.handle(), b.handle(), batch_seed, &o)); return ShardedCiphers(o); }
    std::vector<Fp> dec_value(const ShardedCiphers& c) {
        std::vector<Fp> out(c.size());
        ck(pvacb_group_dec_value(g_, c.handle(), reinterpret_cast<uint64_t*>(out.data())));
        return out;
    }
    std::vector<std::array<uint8_t, 32>> commit_ct(const ShardedCiphers& c) {
        std::vector<std::array<uint8_t, 32>> out(c.size());
        ck(pvacb_group_commit_ct(g_, c.handle(), reinterpret_cast<uint8_t*>(out.data())));
        return out;
    }
    // measures the members' simultaneous export bandwidth and keeps the better routing (direct / relayed over NVLink); GB/s of both
    std::pair<double, double> tune_export() { double a = 0, b = 0; ck(pvacb_group_tune_export(g_, &a, &b)); return {a, b}; }
    pvacb_group* handle() const { return g_; }
private:
    void ck(int rc) { if (rc) throw Error(rc, pvacb_group_last_error(g_)); }
    pvacb_group* g_ = nullptr;
};

}  // namespace pvacb
#endif  // PVACB_HPP
