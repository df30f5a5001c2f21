// Per-thread bodies of enc_value's planning and weight stages, shared between device code (enc.cu) and the host unit
// tests (hosttest.cpp). See enc.cu for the pipeline.
#pragma once
#include "common.cuh"
#include "fp127.cuh"
#include "sha256.cuh"

namespace pvacb {

constexpr int kMaxZ2 = 16, kMaxZ3 = 8;      // plan_noise(depth_hint) for depth_hint <= 23 (enc_text uses 2 + block index)
constexpr int kMaxRaw = kSignal + 2 * kMaxZ2 + 3 * kMaxZ3;   // 64
constexpr int kMaxRnd = (kSignal - 1) + kMaxZ2 + 2 * kMaxZ3; // 39

struct SharePlan {
    Fp value;                 // the share being encrypted (v+mask or -mask)
    uint64_t nlo, nhi, ztag;
    uint64_t salt[kMaxRaw];
    Fp rnd[kMaxRnd];          // r[0..6], then r_i per Z2 group, then (a,b) per Z3 group
    uint16_t idx[kMaxRaw];
    uint8_t ch[kMaxRaw];
    uint8_t pos[kMaxRaw];     // slot of the raw edge inside the share after compact_edges + shuffle_edges
    uint8_t first[kMaxRaw];   // 1 for the first raw edge of its slot
    uint8_t n_raw, n_out;
};

// ops/encrypt.hpp:162-258, tape order as listed in SURVEY Appendix A
PV_HD void plan_share(Tape& t, SharePlan& P, uint64_t canon_tag, int Z2, int Z3) {
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
    P.n_raw = (uint8_t)n;
    // compact_edges (:39-71): one slot per distinct (idx, sign), slots ordered by idx then P before M.
    // (The value-dependent drop "w == 0 and sigma == 0" has probability ~2^-8319 and is reported by enc_weights_kernel.)
    uint8_t cpos[kMaxRaw];
    int n_out = 0;
    for (int a = 0; a < n; a++) {
        int key = P.idx[a] * 2 + P.ch[a];
        int rank = 0;
        bool first = true;
        for (int b = 0; b < n; b++) {
            int kb = P.idx[b] * 2 + P.ch[b];
            if (kb < key) {
                bool seen = false;   // count distinct smaller keys
                for (int c = 0; c < b; c++) seen |= (P.idx[c] * 2 + P.ch[c] == kb);
                if (!seen) rank++;
            }
            if (b < a && kb == key) first = false;
        }
        cpos[a] = (uint8_t)rank;
        P.first[a] = first ? 1 : 0;
        if (first) n_out++;
    }
    P.n_out = (uint8_t)n_out;
    // shuffle_edges (:155-160): for i = n-1..1 swap(E[i], E[word % (i+1)])
    uint8_t order[kMaxRaw], where[kMaxRaw];
    for (int p = 0; p < n_out; p++) order[p] = (uint8_t)p;
    for (int i = n_out - 1; i > 0; i--) {
        int j = (int)(t.next() % (uint64_t)(i + 1));
        uint8_t tmp = order[i]; order[i] = order[j]; order[j] = tmp;
    }
    for (int p = 0; p < n_out; p++) where[order[p]] = (uint8_t)p;
    for (int a = 0; a < n; a++) P.pos[a] = where[cpos[a]];
}

// weights of one share (ops/encrypt.hpp:184-252): solves the signal / Z2 / Z3 relations, multiplies by R, sums merged slots.
// prf[0] = prf_R(seed), prf[1 + gid] = prf_noise_delta(seed, gid, kind) for all but the last group. Returns false if a slot
// weight came out zero (then compact_edges might drop the edge; probability ~2^-127).
PV_HD bool share_weights(const SharePlan& P, const Fp* prf, const Fp* __restrict__ powg, int Z2, int Z3, Fp wsum[kMaxRaw]) {
    const int G = Z2 + Z3;
    const Fp R = prf[0];
    Fp coef[kMaxRaw];
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
    for (int p = 0; p < P.n_out; p++) wsum[p] = fp_zero();
    for (int r = 0; r < P.n_raw; r++) wsum[P.pos[r]] = fp_add(wsum[P.pos[r]], fp_mul(coef[r], R));
    bool ok = true;
    for (int p = 0; p < P.n_out; p++) ok = ok && !fp_is_zero(wsum[p]);
    return ok;
}

// the whole tape walk of one enc_value item (ops/encrypt.hpp:281-291): mask, then share 0 = enc_fp_depth(-mask), then
// share 1 = enc_fp_depth(v+mask). Returns the number of tape words consumed.
PV_HD uint64_t plan_item(uint64_t s0, uint64_t v, uint64_t canon_tag, int Z2, int Z3, SharePlan& P0, SharePlan& P1) {
    Tape t{s0, 0};
    Fp mask = rand_fp_nonzero(t);
    P0.value = fp_neg(mask);
    plan_share(t, P0, canon_tag, Z2, Z3);
    P1.value = fp_add(fp_from_words(v, 0), mask);
    plan_share(t, P1, canon_tag, Z2, Z3);
    return t.k;
}

// enc_fp_depth alone (ops/encrypt.hpp:162-258): one share, no mask; the tape starts at the nonce
PV_HD uint64_t plan_single(uint64_t s0, Fp v, uint64_t canon_tag, int Z2, int Z3, SharePlan& P) {
    Tape t{s0, 0};
    P.value = v;
    plan_share(t, P, canon_tag, Z2, Z3);
    return t.k;
}

}  // namespace pvacb
