// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// Thin C-ABI harness around the UNMODIFIED reference headers (included from where they lie under
// /root/reference/include at build time; no reference source is copied into this repository).
// Built by oracle/Makefile into oracle/_ref/libpvac_ref.so (git-ignored, travels to the GPU box).
//
// What it adds on top of the reference:
//   * a deterministic word tape in place of the OS CSPRNG: the reference draws every random word
//     with one 8-byte getrandom() call (core/random.hpp:48,106-110); the macro below renames that
//     call to a SplitMix64 stream whose state the caller sets (thread-local, so the multi-threaded
//     CPU baseline can run one stream per thread);
//   * select_toeplitz() warm-up, because the first toep_127 call otherwise burns 266 tape words
//     (crypto/toeplitz.hpp:202-267);
//   * struct-of-arrays import/export of pvac::Cipher so Python can compare bytes;
//   * a multi-threaded timing loop used as bench.py's "reference" CPU arm.
#include <sys/types.h>
#include <sys/random.h>
#include <cstdint>
#include <cstring>
#include <cstdio>

static thread_local uint64_t g_tape_state = 0;
static thread_local uint64_t g_tape_draws = 0;
// tape kinds as in pvac_hfhe_cppbyv_b200/csrc/common.cuh: 0 SplitMix64, 1 ChaCha20 (stream id = the seeded state, plus a lane), 2 explicit words
static int g_tape_kind = 0;
static uint32_t g_tape_key[8];
static uint32_t g_tape_lane = 0;
static const uint64_t* g_tape_words = nullptr;
static uint64_t g_tape_nwords = 0;

static inline uint32_t rotl32c(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static void chacha_block(const uint32_t key[8], const uint32_t c[4], uint32_t out[16]) {      // RFC 8439 section 2.3, 20 rounds
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7], c[0], c[1], c[2], c[3]};
    uint32_t x[16];
    std::memcpy(x, in, sizeof x);
    auto qr = [&](int a, int b, int cc, int d) {
        x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32c(x[d], 16); x[cc] += x[d]; x[b] ^= x[cc]; x[b] = rotl32c(x[b], 12);
        x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32c(x[d], 8); x[cc] += x[d]; x[b] ^= x[cc]; x[b] = rotl32c(x[b], 7);
    };
    for (int r = 0; r < 10; r++) { qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15); qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14); }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

static inline uint64_t tape_next() {
    const uint64_t k = g_tape_draws++;
    if (g_tape_kind == 1) {
        uint32_t c[4] = {(uint32_t)(k >> 3), g_tape_lane ^ (uint32_t)((k >> 3) >> 32), (uint32_t)g_tape_state, (uint32_t)(g_tape_state >> 32)}, o[16];
        chacha_block(g_tape_key, c, o);
        return (uint64_t)o[2 * (k & 7)] | ((uint64_t)o[2 * (k & 7) + 1] << 32);
    }
    if (g_tape_kind == 2) return k < g_tape_nwords ? g_tape_words[k] : 0;
    uint64_t z = g_tape_state + (k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

extern "C" ssize_t pvac_ref_tape_getrandom(void* buf, size_t n, unsigned) {
    for (size_t off = 0; off < n; off += 8) {
        uint64_t x = tape_next();
        size_t take = n - off < 8 ? n - off : 8;
        std::memcpy((char*)buf + off, &x, take);
    }
    return (ssize_t)n;
}

#define getrandom pvac_ref_tape_getrandom
#include <pvac/pvac.hpp>
#undef getrandom

#include <thread>
#include <vector>
#include <chrono>
#include <atomic>
#include <unordered_map>

using namespace pvac;

namespace {

struct Keys {
    PubKey pk;
    SecKey sk;
};

// stream seed of item i inside a batch; must equal pvacb's definition (include/pvacb.h, "RNG tape")
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t item_stream_state(uint64_t batch_seed, uint64_t item) {
    return mix64(batch_seed + 0xD1342543DE82EF95ull * (item + 1));
}

}  // namespace

extern "C" {

void ref_init() {
    set_debug_level(0);
    uint64_t save = g_tape_state, save_draws = g_tape_draws;
    select_toeplitz();
    g_tape_state = save; g_tape_draws = save_draws;
}

void ref_seed(uint64_t state) { g_tape_state = state; g_tape_draws = 0; }
void ref_set_tape(int kind, const uint8_t* key32) {
    g_tape_kind = kind;
    if (key32) for (int i = 0; i < 8; i++) g_tape_key[i] = (uint32_t)key32[4 * i] | ((uint32_t)key32[4 * i + 1] << 8) | ((uint32_t)key32[4 * i + 2] << 16) | ((uint32_t)key32[4 * i + 3] << 24);
}
void ref_set_tape_lane(uint32_t lane) { g_tape_lane = lane; }
void ref_set_tape_words(const uint64_t* words, uint64_t n) { g_tape_words = words; g_tape_nwords = n; }
uint64_t ref_tape_draws() { return g_tape_draws; }
uint64_t ref_tape_word() { return tape_next(); }
uint64_t ref_item_stream_state(uint64_t batch_seed, uint64_t item) { return item_stream_state(batch_seed, item); }

// ---------------------------------------------------------------- keys
void* ref_keygen(uint64_t tape_state) {
    Keys* k = new Keys();
    Params prm;
    ref_seed(tape_state);
    keygen(prm, k->pk, k->sk);
    return k;
}

// synthetic keys from raw arrays (Params default). H may be null -> all-zero matrix of the right shape.
void* ref_keys_from_raw(uint64_t canon_tag, const uint8_t* h_digest, const uint64_t* H,
                        const uint64_t* powg /*B x (lo,hi)*/, const uint64_t* prf_k, const uint64_t* lpn_s) {
    Keys* k = new Keys();
    Params prm;
    k->pk.prm = prm;
    k->pk.canon_tag = canon_tag;
    std::memcpy(k->pk.H_digest.data(), h_digest, 32);
    size_t words = (size_t)prm.m_bits / 64;
    k->pk.H.resize(prm.n_bits, BitVec::make(prm.m_bits));
    if (H) {
        for (int c = 0; c < prm.n_bits; c++)
            std::memcpy(k->pk.H[c].w.data(), H + (size_t)c * words, words * 8);
    }
    k->pk.powg_B.resize(prm.B);
    for (int i = 0; i < prm.B; i++) k->pk.powg_B[i] = powg ? Fp{powg[2 * i], powg[2 * i + 1]} : fp_from_u64(1);
    k->pk.omega_B = fp_from_u64(1);
    for (int i = 0; i < 4; i++) k->sk.prf_k[i] = prf_k[i];
    k->sk.lpn_s_bits.assign(lpn_s, lpn_s + (prm.lpn_n + 63) / 64);
    return k;
}

void ref_keys_free(void* h) { delete (Keys*)h; }

void ref_keys_omega(void* h, uint64_t* out2) { Keys* k = (Keys*)h; out2[0] = k->pk.omega_B.lo; out2[1] = k->pk.omega_B.hi; }

// The reference's key files. Its writers are lambdas inside tests/bounty2_test.cpp (:145-192), not part of the headers, so the byte
// layout is restated here field by field OVER THE REFERENCE'S OWN PubKey / SecKey objects (ubk, omega_B and the Params it generated).
int ref_keys_save(void* h, const char* pk_path, const char* sk_path) {
    Keys* k = (Keys*)h;
    auto put32 = [](FILE* f, uint32_t x) { fwrite(&x, 4, 1, f); };
    auto put64 = [](FILE* f, uint64_t x) { fwrite(&x, 8, 1, f); };
    if (sk_path) {
        FILE* f = fopen(sk_path, "wb");
        if (!f) return 1;
        put32(f, 0x66666999u); put32(f, 1);
        for (int j = 0; j < 4; j++) put64(f, k->sk.prf_k[j]);
        put64(f, k->sk.lpn_s_bits.size());
        for (auto w : k->sk.lpn_s_bits) put64(f, w);
        fclose(f);
    }
    if (pk_path) {
        FILE* f = fopen(pk_path, "wb");
        if (!f) return 1;
        const PubKey& pk = k->pk;
        put32(f, 0x06660666u); put32(f, 1);
        put32(f, pk.prm.m_bits); put32(f, pk.prm.B); put32(f, pk.prm.lpn_t); put32(f, pk.prm.lpn_n); put32(f, pk.prm.lpn_tau_num); put32(f, pk.prm.lpn_tau_den);
        put32(f, (uint32_t)pk.prm.noise_entropy_bits); put32(f, (uint32_t)pk.prm.depth_slope_bits);
        uint64_t t2; std::memcpy(&t2, &pk.prm.tuple2_fraction, 8);
        put64(f, t2);
        put32(f, (uint32_t)pk.prm.edge_budget);
        put64(f, pk.canon_tag);
        fwrite(pk.H_digest.data(), 1, 32, f);
        put64(f, pk.H.size());
        for (const auto& b : pk.H) { put32(f, (uint32_t)b.nbits); for (size_t i = 0; i < (b.nbits + 63) / 64; ++i) put64(f, b.w[i]); }
        put64(f, pk.ubk.perm.size());
        for (auto v : pk.ubk.perm) put32(f, v);
        put64(f, pk.ubk.inv.size());
        for (auto v : pk.ubk.inv) put32(f, v);
        put64(f, pk.omega_B.lo); put64(f, pk.omega_B.hi);
        put64(f, pk.powg_B.size());
        for (const auto& x : pk.powg_B) { put64(f, x.lo); put64(f, x.hi); }
        fclose(f);
    }
    return 0;
}

void ref_keys_set_lpn_t(void* h, int t) { ((Keys*)h)->pk.prm.lpn_t = t; }

void ref_keys_export(void* h, uint64_t* canon_tag, uint8_t* h_digest, uint64_t* H, uint64_t* powg,
                     uint64_t* prf_k, uint64_t* lpn_s) {
    Keys* k = (Keys*)h;
    *canon_tag = k->pk.canon_tag;
    std::memcpy(h_digest, k->pk.H_digest.data(), 32);
    size_t words = (size_t)k->pk.prm.m_bits / 64;
    if (H)
        for (size_t c = 0; c < k->pk.H.size(); c++) std::memcpy(H + c * words, k->pk.H[c].w.data(), words * 8);
    for (size_t i = 0; i < k->pk.powg_B.size(); i++) { powg[2 * i] = k->pk.powg_B[i].lo; powg[2 * i + 1] = k->pk.powg_B[i].hi; }
    for (int i = 0; i < 4; i++) prf_k[i] = k->sk.prf_k[i];
    std::memcpy(lpn_s, k->sk.lpn_s_bits.data(), k->sk.lpn_s_bits.size() * 8);
}

// ---------------------------------------------------------------- primitives
void ref_fp_mul(const uint64_t* a, const uint64_t* b, uint64_t* o) { Fp r = fp_mul(Fp{a[0], a[1]}, Fp{b[0], b[1]}); o[0] = r.lo; o[1] = r.hi; }
void ref_fp_add(const uint64_t* a, const uint64_t* b, uint64_t* o) { Fp r = fp_add(Fp{a[0], a[1]}, Fp{b[0], b[1]}); o[0] = r.lo; o[1] = r.hi; }
void ref_fp_sub(const uint64_t* a, const uint64_t* b, uint64_t* o) { Fp r = fp_sub(Fp{a[0], a[1]}, Fp{b[0], b[1]}); o[0] = r.lo; o[1] = r.hi; }
void ref_fp_neg(const uint64_t* a, uint64_t* o) { Fp r = fp_neg(Fp{a[0], a[1]}); o[0] = r.lo; o[1] = r.hi; }
void ref_fp_inv(const uint64_t* a, uint64_t* o) { Fp r = fp_inv(Fp{a[0], a[1]}); o[0] = r.lo; o[1] = r.hi; }
void ref_fp_from_words(uint64_t lo, uint64_t hi, uint64_t* o) { Fp r = fp_from_words(lo, hi); o[0] = r.lo; o[1] = r.hi; }
void ref_hash_to_fp_nonzero(uint64_t lo, uint64_t hi, uint64_t* o) { Fp r = hash_to_fp_nonzero(lo, hi); o[0] = r.lo; o[1] = r.hi; }

void ref_sha256(const uint8_t* p, size_t n, uint8_t* out) { sha256_bytes(p, n, out); }
uint64_t ref_fnv1a(const char* dom) { return fnv1a_domain(dom); }

void ref_aes_ctr_words(const uint8_t* key, uint64_t nonce, uint64_t* out, size_t n) {
    AesCtr256 prg;
    prg.init(key, nonce);
    prg.fill_u64(out, n);
}

// the reference's default Params (core/types.hpp:36-70), field by field in declaration order, as doubles
void ref_params_default(double* out /* 17 */) {
    Params p;
    const double v[17] = {(double)p.B, (double)p.m_bits, (double)p.n_bits, (double)p.h_col_wt, (double)p.x_col_wt, (double)p.err_wt, p.noise_entropy_bits,
                          p.tuple2_fraction, p.depth_slope_bits, (double)p.edge_budget, (double)p.lpn_n, (double)p.lpn_t, (double)p.lpn_tau_num,
                          (double)p.lpn_tau_den, p.recrypt_lo, p.recrypt_hi, (double)p.recrypt_rounds};
    for (int i = 0; i < 17; i++) out[i] = v[i];
}

// a mixed sequence of draws from ONE stream of the reference's AesCtr256: moduli[i] == 0 -> next_u64(), else bounded(moduli[i])
void ref_aes_ctr_draws(const uint8_t* key, uint64_t nonce, const uint64_t* moduli, uint64_t* out, size_t n) {
    AesCtr256 prg;
    prg.init(key, nonce);
    for (size_t i = 0; i < n; i++) out[i] = moduli[i] ? prg.bounded(moduli[i]) : prg.next_u64();
}

void ref_derive_aes_key(void* h, uint64_t ztag, uint64_t nlo, uint64_t nhi, const char* dom, uint8_t* key, uint64_t* nonce) {
    Keys* k = (Keys*)h;
    RSeed s{ztag, Nonce128{nlo, nhi}};
    derive_aes_key(k->pk, k->sk, s, dom, key, *nonce);
}

// ybits must hold lpn_t/64 words
void ref_lpn_make_ybits(void* h, uint64_t ztag, uint64_t nlo, uint64_t nhi, const char* dom, uint64_t* ybits) {
    Keys* k = (Keys*)h;
    RSeed s{ztag, Nonce128{nlo, nhi}};
    std::vector<uint64_t> y;
    lpn_make_ybits(k->pk, k->sk, s, dom, y);
    std::memcpy(ybits, y.data(), y.size() * 8);
}

void ref_toep_127(const uint64_t* top, size_t ntop, const uint64_t* y, size_t ny, uint64_t* out) {
    std::vector<uint64_t> t(top, top + ntop), yy(y, y + ny);
    toep_127(t, yy, out[0], out[1]);
}

void ref_prf_R_core(void* h, uint64_t ztag, uint64_t nlo, uint64_t nhi, const char* dom, uint64_t* o) {
    Keys* k = (Keys*)h;
    Fp r = prf_R_core(k->pk, k->sk, RSeed{ztag, Nonce128{nlo, nhi}}, dom);
    o[0] = r.lo; o[1] = r.hi;
}
void ref_prf_R(void* h, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint64_t* o) {
    Keys* k = (Keys*)h;
    Fp r = prf_R(k->pk, k->sk, RSeed{ztag, Nonce128{nlo, nhi}});
    o[0] = r.lo; o[1] = r.hi;
}
void ref_prf_R_noise(void* h, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint64_t* o) {
    Keys* k = (Keys*)h;
    Fp r = prf_R_noise(k->pk, k->sk, RSeed{ztag, Nonce128{nlo, nhi}});
    o[0] = r.lo; o[1] = r.hi;
}
void ref_prf_noise_delta(void* h, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint32_t gid, uint8_t kind, uint64_t* o) {
    Keys* k = (Keys*)h;
    Fp r = prf_noise_delta(k->pk, k->sk, RSeed{ztag, Nonce128{nlo, nhi}}, gid, kind);
    o[0] = r.lo; o[1] = r.hi;
}
uint64_t ref_prg_layer_ztag(uint64_t canon_tag, uint64_t nlo, uint64_t nhi) { return prg_layer_ztag(canon_tag, Nonce128{nlo, nhi}); }

void ref_prg_choose_k(int kk, int N, const char* label, const uint64_t* words, size_t nwords, int32_t* out) {
    std::vector<uint64_t> w(words, words + nwords);
    auto v = prg_choose_k(kk, N, label, w);
    for (int i = 0; i < kk; i++) out[i] = v[i];
}

void ref_sigma_from_H(void* h, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint16_t idx, uint8_t ch, uint64_t salt, uint64_t* out) {
    Keys* k = (Keys*)h;
    BitVec s = sigma_from_H(k->pk, ztag, Nonce128{nlo, nhi}, idx, ch, salt);
    std::memcpy(out, s.w.data(), s.w.size() * 8);
}

void ref_plan_noise(void* h, int depth, int* z2, int* z3) {
    auto p = plan_noise(((Keys*)h)->pk, depth);
    *z2 = p.first; *z3 = p.second;
}

uint64_t ref_unordered_buckets(uint64_t n) {
    std::unordered_map<uint64_t, int> m;
    m.reserve(n);
    return m.bucket_count();
}

// ---------------------------------------------------------------- ciphertexts
void* ref_enc_value(void* h, uint64_t tape_state, uint64_t v) {
    Keys* k = (Keys*)h;
    ref_seed(tape_state);
    return new Cipher(enc_value(k->pk, k->sk, v));
}
int ref_enc_text(void* h, uint64_t tape_state, const uint8_t* msg, uint64_t len, void** out, int cap) {
    Keys* k = (Keys*)h;
    ref_seed(tape_state);
    std::vector<Cipher> cts = enc_text(k->pk, k->sk, std::string((const char*)msg, (size_t)len));
    if ((int)cts.size() > cap) return -1;
    for (size_t i = 0; i < cts.size(); i++) out[i] = new Cipher(cts[i]);
    return (int)cts.size();
}
int ref_dec_text(void* h, void** cts, int n, uint8_t* out, int cap) {
    Keys* k = (Keys*)h;
    std::vector<Cipher> v;
    for (int i = 0; i < n; i++) v.push_back(*(Cipher*)cts[i]);
    std::string s = dec_text(k->pk, k->sk, v);
    if ((int)s.size() > cap) return -1;
    memcpy(out, s.data(), s.size());
    return (int)s.size();
}
void ref_ubk_perm(void* h, int32_t* perm) {
    Keys* k = (Keys*)h;
    if (k->pk.ubk.perm.empty()) k->pk.ubk = gen_ubk_public(k->pk.canon_tag, k->pk.prm.m_bits);
    for (size_t i = 0; i < k->pk.ubk.perm.size(); i++) perm[i] = k->pk.ubk.perm[i];
}
void* ref_ubk_apply(void* h, void* c) {
    Keys* k = (Keys*)h;
    if (k->pk.ubk.perm.empty()) k->pk.ubk = gen_ubk_public(k->pk.canon_tag, k->pk.prm.m_bits);
    Cipher* r = new Cipher(*(Cipher*)c);
    ubk_apply(k->pk, *r);
    return r;
}
double ref_sigma_density(void* h, void* c) { return sigma_density(((Keys*)h)->pk, *(Cipher*)c); }
void* ref_ct_recrypt(void* h, uint64_t tape_state, void* c, void** pool, int npool) {
    Keys* k = (Keys*)h;
    if (k->pk.ubk.perm.empty()) k->pk.ubk = gen_ubk_public(k->pk.canon_tag, k->pk.prm.m_bits);
    EvalKey ek;
    for (int i = 0; i < npool; i++) ek.zero_pool.push_back(*(Cipher*)pool[i]);
    ref_seed(tape_state);
    return new Cipher(ct_recrypt(k->pk, ek, *(Cipher*)c));
}
void* ref_enc_value_depth(void* h, uint64_t tape_state, uint64_t v, int depth) {
    Keys* k = (Keys*)h;
    ref_seed(tape_state);
    return new Cipher(enc_value_depth(k->pk, k->sk, v, depth));
}
void* ref_enc_zero_depth(void* h, uint64_t tape_state, int depth) {
    Keys* k = (Keys*)h;
    ref_seed(tape_state);
    return new Cipher(enc_zero_depth(k->pk, k->sk, depth));
}
void* ref_ct_neg(void* h, void* a) { return new Cipher(ct_neg(((Keys*)h)->pk, *(Cipher*)a)); }
void* ref_ct_div_const(void* h, void* a, const uint64_t* kk) { return new Cipher(ct_div_const(((Keys*)h)->pk, *(Cipher*)a, Fp{kk[0], kk[1]})); }
// explicit order used by g++ 13.3 for enc_value_depth's two argument calls (ops/encrypt.hpp:284-286)
void* ref_enc_value_explicit(void* h, uint64_t tape_state, uint64_t v, int second_first) {
    Keys* k = (Keys*)h;
    ref_seed(tape_state);
    Fp val = fp_from_u64(v);
    Fp mask = rand_fp_nonzero();
    Cipher a, b;
    if (second_first) {
        b = enc_fp_depth(k->pk, k->sk, fp_neg(mask), 0);
        a = enc_fp_depth(k->pk, k->sk, fp_add(val, mask), 0);
    } else {
        a = enc_fp_depth(k->pk, k->sk, fp_add(val, mask), 0);
        b = enc_fp_depth(k->pk, k->sk, fp_neg(mask), 0);
    }
    return new Cipher(combine_ciphers(k->pk, a, b));
}
void* ref_enc_fp_depth(void* h, uint64_t tape_state, const uint64_t* v, int depth) {
    Keys* k = (Keys*)h;
    ref_seed(tape_state);
    return new Cipher(enc_fp_depth(k->pk, k->sk, Fp{v[0], v[1]}, depth));
}
void* ref_ct_add(void* h, void* a, void* b) { return new Cipher(ct_add(((Keys*)h)->pk, *(Cipher*)a, *(Cipher*)b)); }
void* ref_ct_sub(void* h, void* a, void* b) { return new Cipher(ct_sub(((Keys*)h)->pk, *(Cipher*)a, *(Cipher*)b)); }
void* ref_ct_scale(void* h, void* a, const uint64_t* s) { return new Cipher(ct_scale(((Keys*)h)->pk, *(Cipher*)a, Fp{s[0], s[1]})); }
void ref_commit_ct(void* h, void* c, uint8_t* out32) { auto d = commit_ct(((Keys*)h)->pk, *(Cipher*)c); memcpy(out32, d.data(), 32); }
void* ref_compact_edges(void* h, void* a) { Cipher* c = new Cipher(*(Cipher*)a); compact_edges(((Keys*)h)->pk, *c); return c; }
void* ref_ct_mul(void* h, uint64_t tape_state, void* a, void* b) {
    ref_seed(tape_state);
    return new Cipher(ct_mul(((Keys*)h)->pk, *(Cipher*)a, *(Cipher*)b));
}
void ref_dec_value(void* h, void* c, uint64_t* o) {
    Keys* k = (Keys*)h;
    Fp r = dec_value(k->pk, k->sk, *(Cipher*)c);
    o[0] = r.lo; o[1] = r.hi;
}
void ref_ct_free(void* c) { delete (Cipher*)c; }
void ref_ct_counts(void* c, uint32_t* nL, uint32_t* nE) { *nL = (uint32_t)((Cipher*)c)->L.size(); *nE = (uint32_t)((Cipher*)c)->E.size(); }

// struct-of-arrays export. BASE layers export pa=pb=0 (the reference leaves them uninitialised,
// ops/encrypt.hpp:165-169). sigma: nE x (m_bits/64) words.
void ref_ct_export(void* c, uint8_t* rule, uint64_t* ztag, uint64_t* nlo, uint64_t* nhi, uint32_t* pa, uint32_t* pb,
                   uint32_t* lid, uint16_t* idx, uint8_t* ch, uint64_t* w, uint64_t* sigma) {
    Cipher* C = (Cipher*)c;
    for (size_t i = 0; i < C->L.size(); i++) {
        const Layer& L = C->L[i];
        rule[i] = (uint8_t)L.rule;
        ztag[i] = L.seed.ztag; nlo[i] = L.seed.nonce.lo; nhi[i] = L.seed.nonce.hi;
        pa[i] = L.rule == RRule::PROD ? L.pa : 0;
        pb[i] = L.rule == RRule::PROD ? L.pb : 0;
    }
    for (size_t i = 0; i < C->E.size(); i++) {
        const Edge& e = C->E[i];
        lid[i] = e.layer_id; idx[i] = e.idx; ch[i] = e.ch;
        w[2 * i] = e.w.lo; w[2 * i + 1] = e.w.hi;
        if (sigma) std::memcpy(sigma + i * e.s.w.size(), e.s.w.data(), e.s.w.size() * 8);
    }
}

void* ref_ct_import(uint32_t nL, uint32_t nE, uint32_t m_bits, const uint8_t* rule, const uint64_t* ztag, const uint64_t* nlo,
                    const uint64_t* nhi, const uint32_t* pa, const uint32_t* pb, const uint32_t* lid, const uint16_t* idx,
                    const uint8_t* ch, const uint64_t* w, const uint64_t* sigma) {
    Cipher* C = new Cipher();
    C->L.resize(nL);
    C->E.resize(nE);
    size_t words = m_bits / 64;
    for (uint32_t i = 0; i < nL; i++) {
        Layer L{};
        L.rule = (RRule)rule[i];
        L.seed.ztag = ztag[i]; L.seed.nonce.lo = nlo[i]; L.seed.nonce.hi = nhi[i];
        L.pa = pa[i]; L.pb = pb[i];
        C->L[i] = L;
    }
    for (uint32_t i = 0; i < nE; i++) {
        Edge e{};
        e.layer_id = lid[i]; e.idx = idx[i]; e.ch = ch[i];
        e.w = Fp{w[2 * i], w[2 * i + 1]};
        e.s = BitVec::make(m_bits);
        if (sigma) std::memcpy(e.s.w.data(), sigma + i * words, words * 8);
        C->E[i] = std::move(e);
    }
    return C;
}

// ---------------------------------------------------------------- CPU baseline (bench.py --impl reference / cpu_baseline)
// op: 0 enc_value, 1 ct_add, 2 ct_sub, 3 ct_mul (fresh x fresh), 4 dec_value (fresh), 5 dec_value (fresh x fresh product),
//     11 commit_ct (fresh), 12 commit_ct (product), 13 compact_edges (product), 14 ct_recrypt (fresh, pool of 4), 15 enc_text (100 bytes)
// Runs `iters` operations on each of `threads` std::threads (independent items, thread-local tape) and
// returns elapsed seconds of the slowest thread; *ops_done = threads*iters.
double ref_bench(void* h, int op, int threads, int iters, uint64_t seed, uint64_t* ops_done) {
    Keys* k = (Keys*)h;
    ref_init();
    std::vector<double> secs(threads, 0.0);
    std::vector<std::thread> th;
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    for (int t = 0; t < threads; t++) {
        th.emplace_back([&, t]() {
            g_tape_state = item_stream_state(seed, (uint64_t)t);
            Cipher a = enc_value(k->pk, k->sk, 1000 + t);
            Cipher b = enc_value(k->pk, k->sk, 77 + t);
            Cipher p;
            if (op == 5 || op == 12 || op == 13) p = ct_mul(k->pk, a, b);
            EvalKey ek;
            if (op == 14) ek = make_evalkey(k->pk, k->sk, 4, 0);
            const std::string msg(100, 'x');
            ready++;
            while (!go.load()) std::this_thread::yield();
            auto t0 = std::chrono::steady_clock::now();
            uint64_t sink = 0;
            for (int i = 0; i < iters; i++) {
                switch (op) {
                    case 0: { Cipher c = enc_value(k->pk, k->sk, (uint64_t)i * 2654435761u + t); sink += c.E.size(); break; }
                    case 1: { Cipher c = ct_add(k->pk, a, b); sink += c.E.size(); break; }
                    case 2: { Cipher c = ct_sub(k->pk, a, b); sink += c.E.size(); break; }
                    case 3: { Cipher c = ct_mul(k->pk, a, b); sink += c.E.size(); break; }
                    case 4: { Fp r = dec_value(k->pk, k->sk, a); sink += r.lo; break; }
                    case 5: { Fp r = dec_value(k->pk, k->sk, p); sink += r.lo; break; }
                    case 11: { auto d = commit_ct(k->pk, a); sink += d[0]; break; }
                    case 12: { auto d = commit_ct(k->pk, p); sink += d[0]; break; }
                    case 13: { Cipher c = p; compact_edges(k->pk, c); sink += c.E.size(); break; }
                    case 14: { Cipher c = ct_recrypt(k->pk, ek, a); sink += c.E.size(); break; }
                    case 15: { auto v = enc_text(k->pk, k->sk, msg); sink += v.size(); break; }
                }
            }
            auto t1 = std::chrono::steady_clock::now();
            secs[t] = std::chrono::duration<double>(t1 - t0).count() + (sink == 0xFFFFFFFFFFFFFFFFull ? 1e-12 : 0.0);
        });
    }
    while (ready.load() < threads) std::this_thread::yield();
    go.store(true);
    for (auto& x : th) x.join();
    double mx = 0;
    for (double s : secs) mx = s > mx ? s : mx;
    *ops_done = (uint64_t)threads * (uint64_t)iters;
    return mx;
}


// tests/test_depth.cpp:44-72 on every thread at once: c0 = enc(2), then c <- c*c for `steps` steps, timing each ct_mul and each
// dec_value separately. mul_s[s] / dec_s[s] = elapsed seconds of the slowest thread for step s + 1; edges[s] = edges of that product.
void ref_bench_chain(void* h, int threads, int steps, uint64_t seed, double* mul_s, double* dec_s, uint64_t* edges) {
    Keys* k = (Keys*)h;
    ref_init();
    std::vector<std::vector<double>> tm(threads, std::vector<double>(steps, 0.0)), td(threads, std::vector<double>(steps, 0.0));
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        th.emplace_back([&, t]() {
            g_tape_state = item_stream_state(seed, (uint64_t)t);
            g_tape_draws = 0;
            Cipher c = enc_value(k->pk, k->sk, 2);
            Fp expect = fp_from_u64(2);
            for (int s = 0; s < steps; s++) {
                auto t0 = std::chrono::steady_clock::now();
                c = ct_mul(k->pk, c, c);
                auto t1 = std::chrono::steady_clock::now();
                Fp d = dec_value(k->pk, k->sk, c);
                auto t2 = std::chrono::steady_clock::now();
                expect = fp_mul(expect, expect);
                if (d.lo != expect.lo || d.hi != expect.hi) std::abort();
                tm[t][s] = std::chrono::duration<double>(t1 - t0).count();
                td[t][s] = std::chrono::duration<double>(t2 - t1).count();
                if (t == 0) edges[s] = c.E.size();
            }
        });
    }
    for (auto& x : th) x.join();
    for (int s = 0; s < steps; s++) {
        mul_s[s] = dec_s[s] = 0;
        for (int t = 0; t < threads; t++) { mul_s[s] = std::max(mul_s[s], tm[t][s]); dec_s[s] = std::max(dec_s[s], td[t][s]); }
    }
}

}  // extern "C"
