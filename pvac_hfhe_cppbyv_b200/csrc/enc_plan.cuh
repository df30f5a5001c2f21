// Per-thread bodies of enc_value's planning and weight stages, shared between device code (enc.cu) and the host unit
// tests (hosttest.cpp). See enc.cu for the pipeline.
#pragma once
#include "common.cuh"
#include "fp127.cuh"
#include "sha256.cuh"

namespace pvacb {

constexpr int kKeySlots = 2 * kB;                     // distinct (idx, sign) keys of one layer: key = 2 idx + sign
constexpr int kKeyWords = (kKeySlots + 63) / 64;      // a key set as a bitmap (11 words)
constexpr uint16_t kNoPos = 0xFFFF;                   // raw edge whose slot compact_edges dropped

// noise groups come from plan_noise(depth_hint) (ops/encrypt.hpp:16-27) and are not bounded: enc_text raises depth_hint by one per
// 15-byte block (utils/text.hpp:49-58). The plan of a share therefore lives in a slab sized by the call:
PV_HD int plan_raw_edges(int Z2, int Z3) { return kSignal + 2 * Z2 + 3 * Z3; }
PV_HD int plan_rnd_values(int Z2, int Z3) { return (kSignal - 1) + Z2 + 2 * Z3; }

struct SharePlan {
    Fp value;                 // the share being encrypted (v+mask or -mask)
    uint64_t nlo, nhi, ztag;
    uint32_t n_raw, n_out;
    // slab arrays, RAW = plan_raw_edges entries each (rnd: plan_rnd_values)
    Fp* rnd;                  // r[0..6], then r_i per Z2 group, then (a,b) per Z3 group
    Fp* coef;                 // scratch of share_weights: coefficient of every raw edge
    Fp* wsum;                 // weight of every output slot
    uint64_t* salt;
    uint16_t* idx;
    uint16_t* pos;            // slot of the raw edge inside the share after compact_edges + shuffle_edges (kNoPos: dropped)
    uint16_t* ord;            // scratch of the shuffle
    uint16_t* whr;
    uint8_t* ch;
    uint8_t* first;           // 1 for the first raw edge of its slot
};
PV_HD size_t plan_slab_bytes(int RAW, int RND) {
    size_t b = (size_t)(RND + 2 * RAW) * 16 + (size_t)RAW * 8 + (size_t)RAW * 2 * 4 + (size_t)RAW * 2;
    return (b + 15) & ~(size_t)15;
}
PV_HD void plan_bind(SharePlan& P, uint8_t* slab, int RAW, int RND) {
    P.rnd = reinterpret_cast<Fp*>(slab); slab += (size_t)RND * 16;
    P.coef = reinterpret_cast<Fp*>(slab); slab += (size_t)RAW * 16;
    P.wsum = reinterpret_cast<Fp*>(slab); slab += (size_t)RAW * 16;
    P.salt = reinterpret_cast<uint64_t*>(slab); slab += (size_t)RAW * 8;
    P.idx = reinterpret_cast<uint16_t*>(slab); slab += (size_t)RAW * 2;
    P.pos = reinterpret_cast<uint16_t*>(slab); slab += (size_t)RAW * 2;
    P.ord = reinterpret_cast<uint16_t*>(slab); slab += (size_t)RAW * 2;
    P.whr = reinterpret_cast<uint16_t*>(slab); slab += (size_t)RAW * 2;
    P.ch = slab; slab += RAW;
    P.first = slab;
}

PV_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// ops/encrypt.hpp:162-258, tape order as listed in SURVEY Appendix A. `drop` (kKeyWords words, or null): keys whose merged edge
// compact_edges removes (weight 0 AND syndrome 0, :59) -- value dependent, so it is only known after the weights and syndromes of
// a first pass exist (enc.cu); it changes n_out and with it the number of shuffle draws.
PV_HD void plan_share(Tape& t, SharePlan& P, uint64_t canon_tag, int Z2, int Z3, const uint64_t* drop = nullptr) {
    P.nlo = t.next();
    P.nhi = t.next();
    P.ztag = prg_layer_ztag(canon_tag, P.nlo, P.nhi);
    int n = 0;
    for (int j = 0; j < kSignal; j++) {           // pick_unique_idx + sign, :179-182
        int x;
        for (;;) {
            x = (int)(t.next() % (uint64_t)kB);
            bool dup = false;
            for (int q = 0; q < j; q++) dup |= (P.idx[q] == x);
            if (!dup) break;
        }
        P.idx[j] = (uint16_t)x;
        P.ch[j] = (uint8_t)(t.next() & 1);
    }
    n = kSignal;
    int nr = 0;
    for (int j = 0; j < kSignal - 1; j++) P.rnd[nr++] = rand_fp_nonzero(t);   // :185-189
    for (int j = 0; j < kSignal; j++) P.salt[j] = t.next();                    // make_edge x8, :197-198
    for (int g = 0; g < Z2; g++) {                                             // :212-228
        int i = (int)(t.next() % (uint64_t)kB), j;
        do { j = (int)(t.next() % (uint64_t)kB); } while (j == i);
        uint8_t s1 = (uint8_t)(t.next() & 1);
        P.rnd[nr++] = rand_fp_nonzero(t);
        P.idx[n] = (uint16_t)i; P.ch[n] = s1; P.salt[n] = t.next(); n++;
        P.idx[n] = (uint16_t)j; P.ch[n] = s1 ^ 1; P.salt[n] = t.next(); n++;
    }
    for (int g = 0; g < Z3; g++) {                                             // :230-252
        int i = (int)(t.next() % (uint64_t)kB), j, k;
        do { j = (int)(t.next() % (uint64_t)kB); } while (j == i);
        do { k = (int)(t.next() % (uint64_t)kB); } while (k == i || k == j);
        uint8_t s1 = (uint8_t)(t.next() & 1), s2 = (uint8_t)(t.next() & 1), s3 = (uint8_t)(t.next() & 1);
        P.rnd[nr++] = rand_fp_nonzero(t);
        P.rnd[nr++] = rand_fp_nonzero(t);
        P.idx[n] = (uint16_t)i; P.ch[n] = s1; P.salt[n] = t.next(); n++;
        P.idx[n] = (uint16_t)j; P.ch[n] = s2; P.salt[n] = t.next(); n++;
        P.idx[n] = (uint16_t)k; P.ch[n] = s3; P.salt[n] = t.next(); n++;
    }
    P.n_raw = (uint32_t)n;
    // compact_edges (:39-71): one slot per distinct (idx, sign), slots ordered by idx then P before M = ascending key
    uint64_t have[kKeyWords];
    for (int w = 0; w < kKeyWords; w++) have[w] = 0;
    for (int a = 0; a < n; a++) {
        const int key = P.idx[a] * 2 + P.ch[a];
        const uint64_t bit = 1ull << (key & 63);
        P.first[a] = (have[key >> 6] & bit) ? 0 : 1;
        have[key >> 6] |= bit;
    }
    if (drop)
        for (int w = 0; w < kKeyWords; w++) have[w] &= ~drop[w];
    int n_out = 0;
    for (int w = 0; w < kKeyWords; w++) n_out += popc64(have[w]);
    P.n_out = (uint32_t)n_out;
    // shuffle_edges (:155-160): for i = n-1..1 swap(E[i], E[word % (i+1)])
    for (int p = 0; p < n_out; p++) P.ord[p] = (uint16_t)p;
    for (int i = n_out - 1; i > 0; i--) {
        int j = (int)(t.next() % (uint64_t)(i + 1));
        uint16_t tmp = P.ord[i]; P.ord[i] = P.ord[j]; P.ord[j] = tmp;
    }
    for (int p = 0; p < n_out; p++) P.whr[P.ord[p]] = (uint16_t)p;
    for (int a = 0; a < n; a++) {
        const int key = P.idx[a] * 2 + P.ch[a];
        if (!(have[key >> 6] >> (key & 63) & 1ull)) { P.pos[a] = kNoPos; P.first[a] = 0; continue; }
        int rank = popc64(have[key >> 6] & ((1ull << (key & 63)) - 1));
        for (int w = 0; w < (key >> 6); w++) rank += popc64(have[w]);
        P.pos[a] = P.whr[rank];
    }
}

// weights of one share (ops/encrypt.hpp:184-252): solves the signal / Z2 / Z3 relations, multiplies by R, sums merged slots into
// P.wsum[0..n_out). prf[0] = prf_R(seed), prf[1 + gid] = prf_noise_delta(seed, gid, kind) for all but the last group. Returns the
// number of slots whose weight came out zero (compact_edges drops such an edge if its merged syndrome is zero as well).
PV_HD int share_weights(const SharePlan& P, const Fp* prf, const Fp* __restrict__ powg, int Z2, int Z3) {
    const int G = Z2 + Z3;
    const Fp R = prf[0];
    Fp* coef = P.coef;
    Fp sumg = fp_zero();
    for (int j = 0; j < kSignal - 1; j++) {        // signal edges: sum_j +-r_j g^idx_j = v
        coef[j] = P.rnd[j];
        Fp term = fp_mul(P.rnd[j], powg[P.idx[j]]);
        sumg = P.ch[j] == 0 ? fp_add(sumg, term) : fp_sub(sumg, term);
    }
    {
        int last = kSignal - 1;
        Fp ginv = powg[(kB - P.idx[last]) % kB];    // g has order B: g^-j = g^(B-j) = fp_inv(g^j)
        Fp rl = fp_mul(fp_sub(P.value, sumg), ginv);
        coef[last] = P.ch[last] ? fp_neg(rl) : rl;
    }
    int e = kSignal, nr = kSignal - 1, gid = 0;
    Fp delta_acc = fp_zero();
    for (int g = 0; g < Z2; g++, gid++) {
        Fp Delta;
        if (G - gid <= 1) Delta = fp_neg(delta_acc);
        else { Delta = prf[1 + gid]; delta_acc = fp_add(delta_acc, Delta); }
        Fp Dp = P.ch[e] == 0 ? Delta : fp_neg(Delta);
        Fp ri = P.rnd[nr++];
        Fp rj = fp_mul(fp_sub(fp_mul(ri, powg[P.idx[e]]), Dp), powg[(kB - P.idx[e + 1]) % kB]);
        coef[e] = ri; coef[e + 1] = rj;
        e += 2;
    }
    for (int g = 0; g < Z3; g++, gid++) {
        Fp Delta;
        if (G - gid <= 1) Delta = fp_neg(delta_acc);
        else { Delta = prf[1 + gid]; delta_acc = fp_add(delta_acc, Delta); }
        Fp a = P.rnd[nr++], b = P.rnd[nr++];
        Fp t1 = fp_mul(a, powg[P.idx[e]]), t2 = fp_mul(b, powg[P.idx[e + 1]]);
        if (P.ch[e]) t1 = fp_neg(t1);
        if (P.ch[e + 1]) t2 = fp_neg(t2);
        Fp gkinv = powg[(kB - P.idx[e + 2]) % kB];
        if (P.ch[e + 2]) gkinv = fp_neg(gkinv);      // 1/(-g^k) = -(1/g^k)
        Fp c = fp_mul(fp_sub(Delta, fp_add(t1, t2)), gkinv);
        coef[e] = a; coef[e + 1] = b; coef[e + 2] = c;
        e += 3;
    }
    for (uint32_t p = 0; p < P.n_out; p++) P.wsum[p] = fp_zero();
    for (uint32_t r = 0; r < P.n_raw; r++)
        if (P.pos[r] != kNoPos) P.wsum[P.pos[r]] = fp_add(P.wsum[P.pos[r]], fp_mul(coef[r], R));
    int zeros = 0;
    for (uint32_t p = 0; p < P.n_out; p++) zeros += fp_is_zero(P.wsum[p]) ? 1 : 0;
    return zeros;
}

// the whole tape walk of one enc_value item (ops/encrypt.hpp:281-291): mask, then share 0 = enc_fp_depth(-mask), then
// share 1 = enc_fp_depth(v+mask). `drop`: 2 x kKeyWords words (share 0, share 1) or null. Returns the index of the next unused tape word.
PV_HD uint64_t plan_item(Tape& t, uint64_t v, uint64_t canon_tag, int Z2, int Z3, SharePlan& P0, SharePlan& P1, const uint64_t* drop = nullptr) {
    Fp mask = rand_fp_nonzero(t);
    P0.value = fp_neg(mask);
    plan_share(t, P0, canon_tag, Z2, Z3, drop);
    P1.value = fp_add(fp_from_words(v, 0), mask);
    plan_share(t, P1, canon_tag, Z2, Z3, drop ? drop + kKeyWords : nullptr);
    return t.k;
}

// enc_fp_depth alone (ops/encrypt.hpp:162-258): one share, no mask; the tape starts at the nonce
PV_HD uint64_t plan_single(Tape& t, Fp v, uint64_t canon_tag, int Z2, int Z3, SharePlan& P, const uint64_t* drop = nullptr) {
    P.value = v;
    plan_share(t, P, canon_tag, Z2, Z3, drop);
    return t.k;
}

}  // namespace pvacb
