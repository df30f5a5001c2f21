// CPU-side unit-test harness: compiles the engine's __host__ __device__ per-thread bodies with g++ so that the
// CPU-only test-suite (pytest -m "not gpu") can check them against the oracle. Test infrastructure: it is NOT part of
// libpvacb.so and no product code path calls it.
#include "prf_core.cuh"
#include "enc_plan.cuh"
#include <cstring>
#include <vector>
#include <unordered_map>

namespace pvacb { int keygen_host(uint64_t tape_state, std::vector<uint64_t>& blob); }
using namespace pvacb;

static AesTables g_tab;
static bool g_tab_ok = false;
static const AesTables& tab() { if (!g_tab_ok) { aes_make_tables(g_tab); g_tab_ok = true; } return g_tab; }

extern "C" {

void ht_fp_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* o) {
    Fp x = fp_make(a[0], a[1]), y = b ? fp_make(b[0], b[1]) : fp_zero(), r;
    switch (op) { case 0: r = fp_add(x, y); break; case 1: r = fp_sub(x, y); break; case 2: r = fp_mul(x, y); break; case 3: r = fp_neg(x); break; default: r = fp_inv(x); }
    o[0] = r.lo; o[1] = r.hi;
}
// sum_k a[k] * b[k] through the unreduced 320-bit accumulator of dec_value's edge stage and ct_mul's dense weight products
// (fp_mac_wide, then ONE fp_wide_reduce); every pair is accumulated `repeat` times; wide_out (5 words, optional) = the raw accumulator
void ht_fp_mac_chain(size_t n, const uint64_t* a, const uint64_t* b, uint64_t repeat, uint64_t* o, uint64_t* wide_out) {
    uint64_t acc[5] = {0, 0, 0, 0, 0};
    for (uint64_t r = 0; r < repeat; r++)
        for (size_t k = 0; k < n; k++) fp_mac_wide(acc, fp_make(a[2 * k], a[2 * k + 1]), fp_make(b[2 * k], b[2 * k + 1]));
    Fp v = fp_wide_reduce(acc);
    o[0] = v.lo; o[1] = v.hi;
    if (wide_out) for (int i = 0; i < 5; i++) wide_out[i] = acc[i];
}
void ht_fp_from_words(uint64_t lo, uint64_t hi, uint64_t* o) { Fp r = fp_from_words(lo, hi); o[0] = r.lo; o[1] = r.hi; }
uint64_t ht_tape_word(uint64_t s0, uint64_t k) { return splitmix_word(s0, k); }
// word k of the ChaCha20 tape: key (8 x u32), stream id, lane
uint64_t ht_chacha_word(const uint32_t* key, uint64_t sid, uint32_t lane, uint64_t k) {
    TapeSpec ts; ts.kind = TAPE_CHACHA20; for (int i = 0; i < 8; i++) ts.key[i] = key[i];
    ts.states = &sid;
    Tape t = tape_open(ts, 0); t.lane = lane;
    return t.at(k);
}
// the raw block function, for the RFC 8439 vector: key, words 12..15 -> 64 bytes
void ht_chacha_block(const uint32_t* key, const uint32_t* c, uint64_t* out8) { chacha20_block(key, c[0], c[1], c[2], c[3], out8); }
uint64_t ht_item_stream_state(uint64_t seed, uint64_t item) { return item_stream_state(seed, item); }
uint64_t ht_fnv(int family, int t) { return t == 3 ? kFnvToep : fnv_prf_dom(family, t); }
uint64_t ht_ztag(uint64_t canon, uint64_t nlo, uint64_t nhi) { return prg_layer_ztag(canon, nlo, nhi); }

// SHA-256(label || x[0..nx-1]) for the four labels: 0 x_seed, 1 noise, 2 h_gen, 3 ztag
void ht_sha_label(int label, const uint64_t* x, int nx, uint8_t* out32) {
    LabelStream ls = label == 0 ? label_xseed() : label == 1 ? label_noise() : label == 2 ? label_hgen() : label_ztag();
    ShaState st;
    sha_label_words(ls, x, nx, st);
    for (int i = 0; i < 8; i++) { out32[4 * i] = st.h[i] >> 24; out32[4 * i + 1] = st.h[i] >> 16; out32[4 * i + 2] = st.h[i] >> 8; out32[4 * i + 3] = st.h[i]; }
}

// the fast path of sigma_cand_kernel: midstate over block 0, then block 1 built from (salt, ctr); 4 candidate words
void ht_cand_words(int label, const uint64_t* x7, uint64_t ctr, uint64_t* out4) {
    LabelStream ls = label ? label_noise() : label_xseed();
    uint64_t x[8];
    for (int i = 0; i < 7; i++) x[i] = x7[i];
    x[7] = 0;
    uint64_t q[8];
    for (int j = 0; j < 8; j++) q[j] = stream_word(ls, x, 8, j, 0x80ull);
    uint32_t w[16];
    sha_block_from_le64(q, w);
    ShaState st;
    sha_init(st);
    sha_compress(st, w);
    const int sh = 8 * ls.r;
    uint64_t salt = x[6];
    uint64_t q8 = (salt >> (64 - sh)) | (ctr << sh);
    uint64_t q9 = (ctr >> (64 - sh)) | (0x80ull << sh);
    w[0] = sha_bswap((uint32_t)q8); w[1] = sha_bswap((uint32_t)(q8 >> 32));
    w[2] = sha_bswap((uint32_t)q9); w[3] = sha_bswap((uint32_t)(q9 >> 32));
    for (int i = 4; i < 15; i++) w[i] = 0;
    w[15] = (uint32_t)(8 + ls.r + 64) * 8;
    sha_compress(st, w);
    for (int k = 0; k < 4; k++) out4[k] = sha_digest_le64(st, k);
}

void ht_aes_ctr_words(const uint8_t* key32, uint64_t nonce, uint64_t* out, size_t nblocks) {
    uint32_t key[8], rk[60];
    memcpy(key, key32, 32);
    aes256_expand(tab().sbox, key, rk);
    for (size_t i = 0; i < nblocks; i++) aes256_ctr_block(tab().t0, tab().sbox, rk, nonce + i, out[2 * i], out[2 * i + 1]);
}

static void key_mid(const uint64_t* prf_k, uint64_t canon, const uint8_t* digest, uint32_t mid[8], uint64_t& d3) {
    uint64_t d[4];
    memcpy(d, digest, 32);
    uint64_t q[8] = {prf_k[0], prf_k[1], prf_k[2], prf_k[3], canon, d[0], d[1], d[2]};
    uint32_t w[16];
    sha_block_from_le64(q, w);
    ShaState st;
    sha_init(st);
    sha_compress(st, w);
    for (int i = 0; i < 8; i++) mid[i] = st.h[i];
    d3 = d[3];
}

// one PRF core exactly as prf_setup_kernel + prf_lpn_kernel + prf_finalize_kernel compute it, `rows` = 16384 or 128.
// ybits_out (rows/64 words) optional.
void ht_prf_core(const uint64_t* prf_k, uint64_t canon, const uint8_t* digest, const uint64_t* lpn_s, uint64_t ztag, uint64_t nlo, uint64_t nhi,
                 int family, int t, int rows, uint64_t* ybits_out, uint64_t* out_fp, int* rare_out) {
    uint32_t mid[8];
    uint64_t d3;
    key_mid(prf_k, canon, digest, mid, d3);
    uint32_t rk[60];
    uint64_t ctr0, top0, top1;
    prf_core_setup(mid, d3, tab().t0, tab().sbox, ztag, nlo, nhi, fnv_prf_dom(family, t), rk, ctr0, top0, top1);
    std::vector<uint64_t> y(rows / 64, 0);
    bool rare = false;
    LpnMasks msk;
    lpn_masks_from_secret(lpn_s, msk);
    for (int base = 0; base < rows / 2; base += 32) {      // one "warp" = 32 row pairs
        uint32_t be = 0, bo = 0;
        for (int lane = 0; lane < 32; lane++) {
            uint32_t ye, yo;
            uint32_t rp = base + lane;
            lpn_row_pair([&](uint64_t c, uint64_t& w0, uint64_t& w1) { aes256_ctr_block(tab().t0, tab().sbox, rk, c, w0, w1); }, ctr0 + 65ull * rp, msk, ye, yo, rare);
            be |= ye << lane; bo |= yo << lane;
        }
        y[base / 32] = spread_bits32(be) | (spread_bits32(bo) << 1);
    }
    if (ybits_out) memcpy(ybits_out, y.data(), y.size() * 8);
    Fp r = toep127_to_fp(y[0], y[1], top0, top1);
    out_fp[0] = r.lo; out_fp[1] = r.hi;
    if (rare_out) *rare_out = rare ? 1 : 0;
}

// plan of one enc_value item; flat dump per share: [value lo,hi, nlo, nhi, ztag, n_raw, n_out, then per raw edge: idx, ch, pos, first, salt]
// and rnd values. Returns the number of tape words consumed. drop: optional 2 x kKeyWords words (slots compact_edges removes)
constexpr int kHtMaxRaw = 64, kHtMaxRnd = 39;      // dump layout of these test entry points (depth hints up to 23)
struct HtPlans {
    SharePlan P[2];
    std::vector<uint8_t> slab;
    HtPlans(int Z2, int Z3) {
        const int RAW = plan_raw_edges(Z2, Z3), RND = plan_rnd_values(Z2, Z3);
        slab.assign(2 * plan_slab_bytes(RAW, RND), 0);
        memset(P, 0, sizeof P);
        for (int s = 0; s < 2; s++) plan_bind(P[s], slab.data() + s * plan_slab_bytes(RAW, RND), RAW, RND);
    }
};
uint64_t ht_plan_item_ex(uint64_t s0, uint64_t v, uint64_t canon, int Z2, int Z3, const uint64_t* drop, uint64_t* hdr /*2 x 7*/, uint64_t* raw /*2 x kHtMaxRaw x 5*/,
                         uint64_t* rnd /*2 x kHtMaxRnd x 2*/) {
    if (plan_raw_edges(Z2, Z3) > kHtMaxRaw) return 0;
    HtPlans H(Z2, Z3);
    SharePlan* P = H.P;
    Tape t = tape_splitmix(s0);
    uint64_t used = plan_item(t, v, canon, Z2, Z3, P[0], P[1], drop);
    for (int s = 0; s < 2; s++) {
        uint64_t* h = hdr + 7 * s;
        h[0] = P[s].value.lo; h[1] = P[s].value.hi; h[2] = P[s].nlo; h[3] = P[s].nhi; h[4] = P[s].ztag; h[5] = P[s].n_raw; h[6] = P[s].n_out;
        for (uint32_t r = 0; r < P[s].n_raw; r++) {
            uint64_t* e = raw + ((size_t)s * kHtMaxRaw + r) * 5;
            e[0] = P[s].idx[r]; e[1] = P[s].ch[r]; e[2] = P[s].pos[r]; e[3] = P[s].first[r]; e[4] = P[s].salt[r];
        }
        for (int r = 0; r < plan_rnd_values(Z2, Z3); r++) { rnd[((size_t)s * kHtMaxRnd + r) * 2] = P[s].rnd[r].lo; rnd[((size_t)s * kHtMaxRnd + r) * 2 + 1] = P[s].rnd[r].hi; }
    }
    return used;
}
uint64_t ht_plan_item(uint64_t s0, uint64_t v, uint64_t canon, int Z2, int Z3, uint64_t* hdr, uint64_t* raw, uint64_t* rnd) {
    return ht_plan_item_ex(s0, v, canon, Z2, Z3, nullptr, hdr, raw, rnd);
}
int ht_max_raw() { return kHtMaxRaw; }
int ht_max_rnd() { return kHtMaxRnd; }

// weights of both shares of one item given the PRF values (prf: 2 x G x (lo,hi)); out: 2 x kHtMaxRaw x (lo,hi) per slot
int ht_item_weights(uint64_t s0, uint64_t v, uint64_t canon, int Z2, int Z3, const uint64_t* prf, const uint64_t* powg, uint64_t* out) {
    HtPlans H(Z2, Z3);
    SharePlan* P = H.P;
    Tape t = tape_splitmix(s0);
    plan_item(t, v, canon, Z2, Z3, P[0], P[1]);
    int G = Z2 + Z3, ok = 1;
    for (int s = 0; s < 2; s++) {
        ok &= share_weights(P[s], reinterpret_cast<const Fp*>(prf) + (size_t)s * G, reinterpret_cast<const Fp*>(powg), Z2, Z3) == 0 ? 1 : 0;
        for (uint32_t p = 0; p < P[s].n_out; p++) { out[((size_t)s * kHtMaxRaw + p) * 2] = P[s].wsum[p].lo; out[((size_t)s * kHtMaxRaw + p) * 2 + 1] = P[s].wsum[p].hi; }
    }
    return ok;
}

// host keygen -> blob (PVACB_KEY_BLOB_BYTES / 8 words)
int ht_keygen(uint64_t tape_state, uint64_t* blob_out, size_t words) {
    std::vector<uint64_t> blob;
    int rc = keygen_host(tape_state, blob);
    if (rc) return rc;
    if (words < blob.size()) return 1;
    memcpy(blob_out, blob.data(), blob.size() * 8);
    return 0;
}

// the table mul_count_kernel searches, against the real container
uint64_t ht_next_bkt(uint64_t n) {
    static std::vector<uint64_t> t;
    if (t.empty()) {
        std::__detail::_Prime_rehash_policy pol;
        uint64_t q = 1;
        for (;;) {
            uint64_t p = (uint64_t)pol._M_next_bkt((std::size_t)q);
            if (!t.empty() && p <= t.back()) break;
            t.push_back(p);
            if (p >= (1ull << 40)) break;
            q = p + 1;
        }
    }
    size_t lo = 0, hi = t.size() - 1;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (t[mid] < n) lo = mid + 1; else hi = mid; }
    return t[lo];
}
uint64_t ht_unordered_buckets_real(uint64_t n) {
    std::unordered_map<uint64_t, int> m;
    m.reserve(n);
    return m.bucket_count();
}

}  // extern "C"
