// keygen (crypto/keygen.hpp:35-136, gen_H crypto/matrix.hpp:191-251) restated for the host: it runs once at setup, so it is
// plain C++ over the same Fp / SHA-256 helpers the kernels use. Output = the flat key blob of engine.h.
// Tape order: canon_tag; prf_k[0..3]; rand_fp tries for g (hi word drawn before lo: g++ evaluates fp_from_words' second
// argument first, keygen.hpp:59); rand_fp tries for omega_B (value unused on the hot path but it consumes words);
// lpn_s_bits[0..63]. gen_ubk_public draws nothing and is not needed by enc/add/sub/mul/dec.
#include "engine.h"
#include "sha256.cuh"

#include <cstring>

namespace pvacb {

namespace {

struct HostSha {   // byte-streaming wrapper (core/hash.hpp:136-177) around sha_compress
    ShaState st;
    uint8_t buf[64];
    size_t ptr = 0;
    uint64_t len = 0;
    HostSha() { sha_init(st); }
    void block() {
        uint32_t w[16];
        for (int i = 0; i < 16; i++) w[i] = ((uint32_t)buf[4 * i] << 24) | ((uint32_t)buf[4 * i + 1] << 16) | ((uint32_t)buf[4 * i + 2] << 8) | buf[4 * i + 3];
        sha_compress(st, w);
        ptr = 0;
    }
    void update(const void* data, size_t n) {
        const uint8_t* p = (const uint8_t*)data;
        len += n;
        while (n) {
            size_t take = 64 - ptr < n ? 64 - ptr : n;
            memcpy(buf + ptr, p, take);
            ptr += take; p += take; n -= take;
            if (ptr == 64) block();
        }
    }
    void u64le(uint64_t x) { update(&x, 8); }   // little-endian host
    void final(uint8_t out[32]) {
        uint64_t bits = len * 8;
        uint8_t pad = 0x80, z = 0;
        update(&pad, 1);
        while (ptr != 56) update(&z, 1);
        uint8_t be[8];
        for (int i = 0; i < 8; i++) be[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(be, 8);
        for (int i = 0; i < 8; i++) { out[4 * i] = st.h[i] >> 24; out[4 * i + 1] = st.h[i] >> 16; out[4 * i + 2] = st.h[i] >> 8; out[4 * i + 3] = st.h[i]; }
    }
};

// crypto/matrix.hpp:15-92
void prg_choose_k_host(int k, int N, const LabelStream& ls, const uint64_t* words, int nwords, int* out) {
    std::vector<uint8_t> used((size_t)N, 0);
    uint64_t x[8];
    for (int i = 0; i < nwords; i++) x[i] = words[i];
    uint64_t ctr = 0;
    int n = 0;
    const uint64_t lim = ~0ull - (~0ull % (uint64_t)N);
    while (n < k) {
        x[nwords] = ctr++;
        ShaState st;
        sha_label_words(ls, x, nwords + 1, st);
        for (int j = 0; j < 4 && n < k; j++) {
            uint64_t v = sha_digest_le64(st, j);
            if (v > lim) continue;
            int r = (int)(v % (uint64_t)N);
            if (!used[r]) { used[r] = 1; out[n++] = r; }
        }
    }
}

Fp fp_pow_u128(Fp a, unsigned __int128 e) {
    Fp r = fp_one();
    while (e) {
        if (e & 1) r = fp_mul(r, a);
        a = fp_mul(a, a);
        e >>= 1;
    }
    return r;
}

}  // namespace

// the tape is either SplitMix64 from a 64-bit state (reference-parity vectors) or ChaCha20 keyed by a 256-bit seed (pvacb_keygen_params)
int keygen_host(Tape& t, std::vector<uint64_t>& blob) {
    blob.assign(kBlobWords, 0);
    const uint64_t canon = t.next();
    blob[0] = canon;
    uint64_t* H = &blob[kBlobHdrWords];
    std::vector<int> rows(kHColWt);
    for (int c = 0; c < kNBits; c++) {
        uint64_t words[5] = {(uint64_t)kMBits, (uint64_t)kNBits, (uint64_t)kHColWt, (uint64_t)c, canon};
        prg_choose_k_host(kHColWt, kMBits, label_hgen(), words, 5, rows.data());
        uint64_t* col = H + (size_t)c * kMWords;
        for (int r : rows) col[r >> 6] |= 1ull << (r & 63);
    }
    HostSha hs;
    hs.update("H|v2", 4);
    hs.u64le(kMBits); hs.u64le(kNBits); hs.u64le(kHColWt);
    hs.update(H, (size_t)kNBits * kMWords * 8);
    uint8_t dg[32];
    hs.final(dg);
    memcpy(&blob[1], dg, 32);
    for (int i = 0; i < 4; i++) blob[5 + i] = t.next();
    const unsigned __int128 pm1 = ((((unsigned __int128)1) << 127) - 2);
    const unsigned __int128 E = pm1 / (unsigned)kB;
    auto rand_fp = [&]() {
        for (;;) {
            uint64_t hi = t.next() & kMask63;
            uint64_t lo = t.next();
            Fp x = fp_from_words(lo, hi);
            if (!fp_is_zero(x)) return x;
        }
    };
    Fp g;
    for (;;) {
        Fp acc = fp_pow_u128(rand_fp(), E);
        if (!fp_eq(acc, fp_one())) { g = acc; break; }
    }
    Fp pw = fp_one();
    for (int i = 0; i < kB; i++) {
        blob[73 + 2 * i] = pw.lo;
        blob[73 + 2 * i + 1] = pw.hi;
        pw = fp_mul(pw, g);
    }
    for (;;) {   // omega_B, keygen.hpp:99-122: exponent truncated to 64 bits; B = 337 is prime so the first w != 1 is accepted
        Fp w = fp_pow_u128(rand_fp(), (unsigned __int128)(uint64_t)E);
        if (!fp_eq(w, fp_one())) { blob[748] = w.lo; blob[749] = w.hi; break; }
    }
    for (int i = 0; i < kLpnWords; i++) blob[9 + i] = t.next();
    return PV_OK;
}
int keygen_host(uint64_t tape_state, std::vector<uint64_t>& blob) {
    Tape t = tape_splitmix(tape_state);
    return keygen_host(t, blob);
}

}  // namespace pvacb
