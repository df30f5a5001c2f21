// ct_mul (ops/arithmetic.hpp:47-106) for a batch of ciphertext pairs.
//
// Reference per pair: new PROD layers (nonce from the tape + ztag hash) for every (la, lb); all |A.E|*|B.E| weight products
// accumulated in a std::unordered_map keyed by (layer pair, (idxa+idxb) mod B) with separate sums for equal / unequal
// signs; one output edge per non-zero sum IN THE MAP'S ITERATION ORDER, each drawing a salt and a fresh sigma_from_H;
// compact_layers. Batched restatement:
//   mul_count_kernel    per pair: layer / key counts, libstdc++ bucket count for reserve(|A.E|*|B.E|)
//   mul_layers_kernel   pre-compaction layer table (A, B shifted, PROD with tape nonces and SHA-256 ztag)
//   mul_bylayer_kernel  edges of A and B grouped by layer (counting sort per ciphertext)
//   mul_pairs_kernel    one CTA per (pair, la, lb): B's layer as a dense (idx, sign) table in shared memory, one thread per
//                       output residue s accumulates sum_{ea} w_a * w_b[(s - idx_a) mod B] and the first insertion time
//   mul_bucket_*        first-occupation time of every hash bucket (open-addressing table, atomicMin); for pairs with at most
//                       4096 buckets the table lives in shared memory inside the sort kernel instead
//   radix sort          keys of each pair ordered by (bucket time desc, insertion time desc) = libstdc++ iteration order
//                       (one CTA per pair in shared memory when a pair has <= 2048 keys, else one device-wide sort)
//   mul_emit_*          P / M edges, salts from the tape in emission order
//   sigma_run           one sigma_from_H per output edge (sigma.cu) -- > 99% of the time
//   compact_layers_batch
#include "engine.h"
#include "sha256.cuh"

#include <cstdlib>
#include <cstring>
#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace pvacb {

constexpr uint32_t kNone = 0xFFFFFFFFu;

struct MulCounts {          // per pair
    uint32_t nLpre;         // LA + LB + LA*LB
    uint32_t nLP;           // LA*LB
    uint32_t nKeys;         // LA*LB*B
    uint32_t tblSize;       // bucket-time table slots (power of two)
};

__device__ __forceinline__ uint64_t next_bkt(const uint64_t* __restrict__ primes, int np, uint64_t n) {
    int lo = 0, hi = np - 1;            // smallest table entry >= n (std::lower_bound, as _Prime_rehash_policy::_M_next_bkt)
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (primes[mid] < n) lo = mid + 1; else hi = mid;
    }
    return primes[lo];
}

struct MulShape { uint32_t lpre, lp, keys, tbl; uint64_t nb, pairs; bool bad; };
__device__ __forceinline__ MulShape mul_shape(uint64_t i, const uint32_t* __restrict__ la, const uint32_t* __restrict__ lb, const uint32_t* __restrict__ ea,
                                              const uint32_t* __restrict__ eb, const uint64_t* __restrict__ primes, int np) {
    MulShape m;
    uint64_t LA = la[i + 1] - la[i], LB = lb[i + 1] - lb[i], EA = ea[i + 1] - ea[i], EB = eb[i + 1] - eb[i];
    uint64_t lp = LA * LB, keys = lp * kB, pairs = EA * EB;
    m.bad = lp > 0xFFFFFFull || keys >= (1ull << 31) || pairs >= (1ull << 31);
    if (m.bad) lp = keys = pairs = 0;
    m.lpre = (uint32_t)(LA + LB + lp); m.lp = (uint32_t)lp; m.keys = (uint32_t)keys; m.pairs = pairs;
    uint64_t present_max = keys < pairs ? keys : pairs;
    uint32_t t = 2;
    while (t < 2 * present_max) t <<= 1;
    m.tbl = present_max ? t : 0;
    m.nb = pairs ? next_bkt(primes, np, pairs) : 1;
    return m;
}

// Shapes of every pair, the four offset tables (exclusive scans) and the batch totals in ONE launch of one CTA: the totals go
// straight to the host mailbox (mapped pinned memory), so the host needs a single stream synchronisation before it can size the
// result. mail: [0] layers (pre-compaction), [1] layer pairs, [2] keys, [3] table slots, [4..5] max edge pairs, [6..7] sum keys,
// [8..9] sum table slots, [10..11] max keys, [12..13] max buckets, [14] error bits.
__global__ void __launch_bounds__(1024)
mul_count_scan_kernel(uint64_t n, const uint32_t* __restrict__ la, const uint32_t* __restrict__ lb, const uint32_t* __restrict__ ea,
                      const uint32_t* __restrict__ eb, const uint64_t* __restrict__ primes, int np, uint32_t* __restrict__ loP, uint32_t* __restrict__ lpoff,
                      uint32_t* __restrict__ koff, uint32_t* __restrict__ toff, uint64_t* __restrict__ nb, volatile uint32_t* __restrict__ mail) {
    __shared__ uint32_t p_lpre[1024], p_lp[1024], p_keys[1024], p_tbl[1024];
    __shared__ unsigned long long s_stat[5];
    __shared__ unsigned int s_err;
    const int t = threadIdx.x;
    if (t < 5) s_stat[t] = 0;
    if (t == 0) s_err = 0;
    __syncthreads();
    const uint64_t per = (n + 1023) / 1024, b = (uint64_t)t * per, e = b + per < n ? b + per : n;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    unsigned long long mxp = 0, sk = 0, st = 0, mxk = 0, mxb = 0;
    bool bad = false;
    for (uint64_t i = b; i < e; i++) {
        const MulShape m = mul_shape(i, la, lb, ea, eb, primes, np);
        a0 += m.lpre; a1 += m.lp; a2 += m.keys; a3 += m.tbl;
        bad |= m.bad;
        mxp = max(mxp, (unsigned long long)m.pairs); sk += m.keys; st += m.tbl; mxk = max(mxk, (unsigned long long)m.keys); mxb = max(mxb, (unsigned long long)m.nb);
    }
    p_lpre[t] = a0; p_lp[t] = a1; p_keys[t] = a2; p_tbl[t] = a3;
    if (b < e) {
        atomicMax(&s_stat[0], mxp); atomicAdd(&s_stat[1], sk); atomicAdd(&s_stat[2], st); atomicMax(&s_stat[3], mxk); atomicMax(&s_stat[4], mxb);
        if (bad) atomicOr(&s_err, 1u);
    }
    __syncthreads();
    if (t < 4) {                       // four serial scans of 1024 partials, one per thread
        uint32_t* part = t == 0 ? p_lpre : t == 1 ? p_lp : t == 2 ? p_keys : p_tbl;
        uint32_t run = 0;
        for (int k = 0; k < 1024; k++) { uint32_t v = part[k]; part[k] = run; run += v; }
        (t == 0 ? loP : t == 1 ? lpoff : t == 2 ? koff : toff)[n] = run;
        mail[t] = run;
    }
    __syncthreads();
    if (t < 5) { mail[4 + 2 * t] = (uint32_t)s_stat[t]; mail[5 + 2 * t] = (uint32_t)(s_stat[t] >> 32); }
    if (t == 5) mail[14] = s_err;
    uint32_t r0 = p_lpre[t], r1 = p_lp[t], r2 = p_keys[t], r3 = p_tbl[t];
    for (uint64_t i = b; i < e; i++) {
        const MulShape m = mul_shape(i, la, lb, ea, eb, primes, np);
        loP[i] = r0; lpoff[i] = r1; koff[i] = r2; toff[i] = r3; nb[i] = m.nb;
        r0 += m.lpre; r1 += m.lp; r2 += m.keys; r3 += m.tbl;
    }
    __threadfence_system();
}

// one CTA per pair
__global__ void __launch_bounds__(128)
mul_layers_kernel(uint64_t n, const __grid_constant__ TapeSpec ts, uint64_t canon_tag, const uint32_t* __restrict__ loA, const uint32_t* __restrict__ loB,
                  const uint8_t* __restrict__ ruleA, const uint64_t* __restrict__ ztA, const uint64_t* __restrict__ nlA, const uint64_t* __restrict__ nhA,
                  const uint32_t* __restrict__ paA, const uint32_t* __restrict__ pbA, const uint8_t* __restrict__ ruleB, const uint64_t* __restrict__ ztB,
                  const uint64_t* __restrict__ nlB, const uint64_t* __restrict__ nhB, const uint32_t* __restrict__ paB, const uint32_t* __restrict__ pbB,
                  const uint32_t* __restrict__ loP, const uint32_t* __restrict__ lpoff, uint8_t* __restrict__ rule, uint64_t* __restrict__ zt,
                  uint64_t* __restrict__ nl, uint64_t* __restrict__ nh, uint32_t* __restrict__ pa, uint32_t* __restrict__ pb,
                  uint32_t* __restrict__ lp_item) {
    const uint64_t i = blockIdx.x;
    const uint32_t a0 = loA[i], LA = loA[i + 1] - a0, b0 = loB[i], LB = loB[i + 1] - b0, o0 = loP[i];
    for (uint32_t k = threadIdx.x; k < LA; k += blockDim.x) {
        rule[o0 + k] = ruleA[a0 + k]; zt[o0 + k] = ztA[a0 + k]; nl[o0 + k] = nlA[a0 + k]; nh[o0 + k] = nhA[a0 + k];
        pa[o0 + k] = paA[a0 + k]; pb[o0 + k] = pbA[a0 + k];
    }
    for (uint32_t k = threadIdx.x; k < LB; k += blockDim.x) {
        uint8_t r = ruleB[b0 + k];
        uint32_t off = r == 1 ? LA : 0;
        rule[o0 + LA + k] = r; zt[o0 + LA + k] = ztB[b0 + k]; nl[o0 + LA + k] = nlB[b0 + k]; nh[o0 + LA + k] = nhB[b0 + k];
        pa[o0 + LA + k] = paB[b0 + k] + off; pb[o0 + LA + k] = pbB[b0 + k] + off;
    }
    if (o0 + LA + LB + LA * LB != loP[i + 1]) return;   // shape overflow was flagged by mul_count_kernel
    Tape tp = tape_open(ts, i);
    const uint64_t w0 = tp.k;
    const uint32_t base = o0 + LA + LB, lp0 = lpoff[i];
    for (uint32_t lp = threadIdx.x; lp < LA * LB; lp += blockDim.x) {   // ops/arithmetic.hpp:59-70, row-major (la, lb)
        uint64_t nlo = tp.at(w0 + 2ull * lp), nhi = tp.at(w0 + 2ull * lp + 1);
        rule[base + lp] = 1;
        nl[base + lp] = nlo; nh[base + lp] = nhi;
        zt[base + lp] = prg_layer_ztag(canon_tag, nlo, nhi);
        pa[base + lp] = lp / LB;
        pb[base + lp] = LA + lp % LB;
        lp_item[lp0 + lp] = (uint32_t)i;
    }
}

// edges of one batch grouped by layer: order[eoff[i] + k] = local edge ids sorted by layer; lstart[loff[i] + l] = first k
__global__ void __launch_bounds__(128)
mul_bylayer_kernel(const uint32_t* __restrict__ loff, const uint32_t* __restrict__ eoff, const uint32_t* __restrict__ lid, uint32_t* __restrict__ lstart,
                   uint32_t* __restrict__ lcount, uint32_t* __restrict__ cursor, uint32_t* __restrict__ order, unsigned int* __restrict__ err) {
    const uint64_t i = blockIdx.x;
    const uint32_t l0 = loff[i], L = loff[i + 1] - l0, e0 = eoff[i], E = eoff[i + 1] - e0;
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) lcount[l0 + l] = 0;
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < E; e += blockDim.x) {
        uint32_t l = lid[e0 + e];
        if (l >= L) { atomicOr(err, 2u); continue; }
        atomicAdd(&lcount[l0 + l], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t l = 0; l < L; l++) { lstart[l0 + l] = run; cursor[l0 + l] = run; run += lcount[l0 + l]; }
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < E; e += blockDim.x) {
        uint32_t l = lid[e0 + e];
        if (l >= L) continue;
        uint32_t k = atomicAdd(&cursor[l0 + l], 1u);
        order[e0 + k] = e;
    }
}

struct EdgeView {
    const uint32_t *loff, *eoff;
    const uint16_t* idx;
    const uint8_t* ch;
    const Fp* w;
    const uint32_t *lstart, *lcount, *order;
};

constexpr int kPairThreads = 352;   // >= B = 337 residues
constexpr uint32_t kSparsePairs = 4096;   // layer pairs with at most this many products take the product-parallel path

// one CTA per (pair, la, lb)
__global__ void __launch_bounds__(kPairThreads)
mul_pairs_kernel(EdgeView A, EdgeView Bv, const uint32_t* __restrict__ lp_item, const uint32_t* __restrict__ lpoff, const uint32_t* __restrict__ koff,
                 Fp* __restrict__ k_wp, Fp* __restrict__ k_wm, uint8_t* __restrict__ k_flags, uint32_t* __restrict__ k_tins,
                 unsigned int* __restrict__ err) {
    __shared__ Fp s_w[kB * 2];
    __shared__ uint32_t s_ib[kB * 2];
    __shared__ Fp s_aw[kB * 2];          // a layer holds at most one edge per (idx, sign)
    __shared__ uint32_t s_ai[kB * 2];    // idx << 1 | sign
    __shared__ uint32_t s_aia[kB * 2];   // edge index inside the ciphertext
    __shared__ uint32_t s_ndup;
    const uint32_t g = blockIdx.x;
    const uint32_t i = lp_item[g];
    const uint32_t lp = g - lpoff[i];
    const uint32_t LB = Bv.loff[i + 1] - Bv.loff[i];
    const uint32_t la = lp / LB, lb = lp % LB;
    const uint32_t EB = Bv.eoff[i + 1] - Bv.eoff[i];
    const uint32_t a_l = A.loff[i] + la, b_l = Bv.loff[i] + lb;
    const uint32_t nA = A.lcount[a_l], nB = Bv.lcount[b_l];
    const uint32_t kbase = koff[i] + lp * kB;
    const int s = threadIdx.x;
    if (nA == 0 || nB == 0) {
        if (s < kB) { k_tins[kbase + s] = kNone; k_flags[kbase + s] = 0; }
        return;
    }
    for (int k = s; k < kB * 2; k += kPairThreads) s_ib[k] = kNone;
    if (s == 0) s_ndup = 0;
    __syncthreads();
    const uint32_t ea0 = A.eoff[i], eb0 = Bv.eoff[i];
    for (uint32_t k = s; k < nB; k += kPairThreads) {
        uint32_t ib = Bv.order[eb0 + Bv.lstart[b_l] + k];
        uint32_t slot = (uint32_t)Bv.idx[eb0 + ib] * 2 + Bv.ch[eb0 + ib];
        uint32_t old = atomicMin(&s_ib[slot], ib);    // the earliest edge of the slot decides its insertion times
        if (old != kNone) atomicAdd(&s_ndup, 1u);     // two edges with equal (layer, idx, sign): legal in an imported ciphertext
    }
    __syncthreads();
    for (uint32_t k = s; k < nB; k += kPairThreads) {
        uint32_t ib = Bv.order[eb0 + Bv.lstart[b_l] + k];
        uint32_t slot = (uint32_t)Bv.idx[eb0 + ib] * 2 + Bv.ch[eb0 + ib];
        if (s_ib[slot] == ib) s_w[slot] = Bv.w[eb0 + ib];
    }
    if (s_ndup) {                                     // rare: the reference's map simply keeps adding (ops/arithmetic.hpp:79-88); sums are exact
        __syncthreads();
        if (s == 0)
            for (uint32_t k = 0; k < nB; k++) {
                uint32_t ib = Bv.order[eb0 + Bv.lstart[b_l] + k];
                uint32_t slot = (uint32_t)Bv.idx[eb0 + ib] * 2 + Bv.ch[eb0 + ib];
                if (s_ib[slot] != ib) s_w[slot] = fp_add(s_w[slot], Bv.w[eb0 + ib]);
            }
    }
    const uint32_t npairs = nA * nB;
    if (nA <= kB * 2 && npairs <= kSparsePairs) {
        // ---- sparse layers (fresh x fresh: 20 x 20 products into 2 x 337 slots): one thread per PRODUCT, summed into per-slot
        // accumulators in shared memory under a per-slot lock. The residue-parallel form below visits all 337 x nA x 2
        // candidate slots and finds 3 % of them occupied -- 30 x more issue slots than products (ncu r01: 264 us per 1 024 pairs).
        // Field addition is exact and commutative, so the order of the additions does not matter.
        Fp* acc = s_aw;                        // [2][kB]: sums of equal-sign and unequal-sign products
        uint32_t* lock = s_ai;                 // [2 * kB]
        uint32_t* tfirst = s_aia;              // [kB]  first insertion time of the key   (kB <= 2 kB entries)
        for (int k = s; k < kB * 2; k += kPairThreads) { acc[k] = fp_zero(); lock[k] = 0; if (k < kB) tfirst[k] = kNone; }
        __syncthreads();
        for (uint32_t t = s; t < npairs; t += kPairThreads) {
            const uint32_t ka = t / nB, kb = t - ka * nB;
            const uint32_t ia = A.order[ea0 + A.lstart[a_l] + ka], ib = Bv.order[eb0 + Bv.lstart[b_l] + kb];
            const uint32_t res = ((uint32_t)A.idx[ea0 + ia] + (uint32_t)Bv.idx[eb0 + ib]) % kB;
            const uint32_t slot = (A.ch[ea0 + ia] == Bv.ch[eb0 + ib] ? 0u : (uint32_t)kB) + res;
            const Fp ww = fp_mul(A.w[ea0 + ia], Bv.w[eb0 + ib]);
            atomicMin(&tfirst[res], ia * EB + ib);       // position of the pair in the reference's double loop
            bool done = false;
            while (!done) {
                if (atomicCAS(&lock[slot], 0u, 1u) == 0u) {
                    __threadfence_block();
                    volatile Fp* p = acc + slot;
                    Fp cur = fp_make(p->lo, p->hi);
                    cur = fp_add(cur, ww);
                    p->lo = cur.lo; p->hi = cur.hi;
                    __threadfence_block();
                    atomicExch(&lock[slot], 2u);          // 2 = free and touched
                    done = true;
                } else if (lock[slot] == 2u) {
                    if (atomicCAS(&lock[slot], 2u, 1u) == 2u) {
                        __threadfence_block();
                        volatile Fp* p = acc + slot;
                        Fp cur = fp_make(p->lo, p->hi);
                        cur = fp_add(cur, ww);
                        p->lo = cur.lo; p->hi = cur.hi;
                        __threadfence_block();
                        atomicExch(&lock[slot], 2u);
                        done = true;
                    }
                }
            }
        }
        __syncthreads();
        if (s >= kB) return;
        k_wp[kbase + s] = acc[s];
        k_wm[kbase + s] = acc[kB + s];
        k_flags[kbase + s] = (uint8_t)((lock[s] == 2u ? 1 : 0) | (lock[kB + s] == 2u ? 2 : 0));
        k_tins[kbase + s] = tfirst[s];
        return;
    }
    // ---- dense layers: one thread per output residue. A's layer is staged in shared memory too (in chunks of 2B edges; one
    // chunk unless A carries duplicate (idx, sign) edges). The products are summed UNREDUCED in two 320-bit accumulators per thread
    // (fp_mac_wide: a third of the instructions of fp_mul + fp_add) and reduced once at the end -- r02: this loop was 37 % of a depth-3 ct_mul.
    uint64_t accp[5] = {0, 0, 0, 0, 0}, accm[5] = {0, 0, 0, 0, 0};
    uint32_t tmin = kNone;
    uint8_t fl = 0;
    for (uint32_t c0 = 0; c0 < nA; c0 += kB * 2) {
        const uint32_t cn = min(nA - c0, (uint32_t)(kB * 2));
        __syncthreads();
        for (uint32_t k = s; k < cn; k += kPairThreads) {
            uint32_t ia = A.order[ea0 + A.lstart[a_l] + c0 + k];
            s_aw[k] = A.w[ea0 + ia];
            s_aia[k] = ia;
            s_ai[k] = ((uint32_t)A.idx[ea0 + ia] << 1) | A.ch[ea0 + ia];
        }
        __syncthreads();
        if (s < kB)
            for (uint32_t k = 0; k < cn; k++) {
                const uint32_t pk = s_ai[k], ia = s_aia[k];      // same address for the whole CTA: broadcast
                const uint32_t ida = pk >> 1, cha = pk & 1;
                int j = s - (int)ida;
                if (j < 0) j += kB;
#pragma unroll
                for (uint32_t sb = 0; sb < 2; sb++) {
                    uint32_t ib = s_ib[j * 2 + sb];
                    if (ib == kNone) continue;
                    uint32_t t = ia * EB + ib;                   // position of the pair in the reference's double loop
                    tmin = min(tmin, t);
                    if (cha == sb) { fp_mac_wide(accp, s_aw[k], s_w[j * 2 + sb]); fl |= 1; }
                    else { fp_mac_wide(accm, s_aw[k], s_w[j * 2 + sb]); fl |= 2; }
                }
            }
    }
    const Fp wp = fp_wide_reduce(accp), wm = fp_wide_reduce(accm);
    if (s >= kB) return;
    k_wp[kbase + s] = wp;
    k_wm[kbase + s] = wm;
    k_flags[kbase + s] = fl;
    k_tins[kbase + s] = tmin;
}

__device__ __forceinline__ uint32_t find_item(const uint32_t* __restrict__ off, uint32_t n, uint32_t x) {   // off[i] <= x < off[i+1]
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// thread per key: bucket = (k * phi mod 2^64) % nb  (struct H, ops/arithmetic.hpp:73); table: bucket -> min insertion time
__global__ void mul_bucket_insert_kernel(uint32_t nkeys, uint32_t n, const uint32_t* __restrict__ koff, const uint32_t* __restrict__ toff,
                                         const uint64_t* __restrict__ nb, const uint32_t* __restrict__ k_tins, uint32_t* __restrict__ k_item,
                                         unsigned long long* __restrict__ t_key, uint32_t* __restrict__ t_val) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    uint32_t i = find_item(koff, n, k);
    k_item[k] = i;
    uint32_t t = k_tins[k];
    if (t == kNone) return;
    uint32_t rel = k - koff[i];
    uint64_t key = ((uint64_t)(rel / kB) << 32) | (rel % kB);
    uint64_t bkt = (key * 0x9E3779B97F4A7C15ull) % nb[i];
    uint32_t size = toff[i + 1] - toff[i];
    uint32_t h = (uint32_t)(mix64(bkt) & (size - 1));
    for (;;) {
        unsigned long long cur = atomicCAS(&t_key[toff[i] + h], ~0ull, (unsigned long long)bkt);
        if (cur == ~0ull || cur == bkt) break;
        h = (h + 1) & (size - 1);
    }
    atomicMin(&t_val[toff[i] + h], t);
}

// sort key: (item, PMAX - (t_bkt+1), PMAX - (t_ins+1)) ascending = (bucket time desc, insertion time desc); absent keys last
__global__ void mul_sortkey_kernel(uint32_t nkeys, const uint32_t* __restrict__ koff, const uint32_t* __restrict__ toff, const uint64_t* __restrict__ nb,
                                   const uint32_t* __restrict__ k_tins, const uint32_t* __restrict__ k_item, const unsigned long long* __restrict__ t_key,
                                   const uint32_t* __restrict__ t_val, int pbits, uint64_t* __restrict__ skey, uint32_t* __restrict__ sval) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    uint32_t i = k_item[k];
    uint32_t t = k_tins[k];
    const uint64_t pmax = (1ull << pbits) - 1;
    uint64_t lowpart;
    if (t == kNone) lowpart = (pmax << pbits) | pmax;
    else {
        uint32_t rel = k - koff[i];
        uint64_t key = ((uint64_t)(rel / kB) << 32) | (rel % kB);
        uint64_t bkt = (key * 0x9E3779B97F4A7C15ull) % nb[i];
        uint32_t size = toff[i + 1] - toff[i];
        uint32_t h = (uint32_t)(mix64(bkt) & (size - 1));
        while (t_key[toff[i] + h] != bkt) h = (h + 1) & (size - 1);
        uint64_t tb = t_val[toff[i] + h];
        lowpart = ((pmax - (tb + 1)) << pbits) | (pmax - ((uint64_t)t + 1));
    }
    skey[k] = ((uint64_t)i << (2 * pbits)) | lowpart;
    sval[k] = k;
}

// The same order for batches whose ciphertext pairs have at most kBlockSortKeys keys each (fresh x fresh: 4 * 337 = 1 348): one
// CTA per pair computes the keys and sorts them in shared memory (cub::BlockRadixSort, stable like the device-wide sort), so
// the result is identical to sorting (item, key) globally -- without the 4-5 passes over all keys of the batch.
constexpr int kBlockSortThreads = 256, kBlockSortItems = 8, kBlockSortKeys = kBlockSortThreads * kBlockSortItems;
constexpr uint32_t kSmemBuckets = 4096;     // SMEM_TABLE: the pair's bucket table (first-occupation time per bucket) lives in shared memory
template <bool SMEM_TABLE>
__global__ void __launch_bounds__(kBlockSortThreads)
mul_sort_block_kernel(const uint32_t* __restrict__ koff, const uint32_t* __restrict__ toff, const uint64_t* __restrict__ nb, const uint32_t* __restrict__ k_tins,
                      const unsigned long long* __restrict__ t_key, const uint32_t* __restrict__ t_val, int pbits, uint32_t* __restrict__ sval,
                      uint32_t* __restrict__ k_item) {
    using Sort = cub::BlockRadixSort<uint32_t, kBlockSortThreads, kBlockSortItems, uint32_t>;
    __shared__ typename Sort::TempStorage tmp;
    __shared__ uint32_t s_tb[SMEM_TABLE ? kSmemBuckets : 1];
    const uint32_t i = blockIdx.x;
    const uint32_t k0 = koff[i], nk = koff[i + 1] - k0;
    const uint32_t pmax = (1u << pbits) - 1;
    const uint64_t nbi = nb[i];
    uint32_t key[kBlockSortItems], val[kBlockSortItems], tin[kBlockSortItems], bk[kBlockSortItems];
    if (SMEM_TABLE) {
        for (uint32_t b = threadIdx.x; b < (uint32_t)nbi; b += kBlockSortThreads) s_tb[b] = kNone;
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < kBlockSortItems; j++) {
        const uint32_t rel = threadIdx.x * kBlockSortItems + j;          // blocked arrangement keeps the original order of equal keys
        val[j] = k0 + rel;
        tin[j] = kNone;
        bk[j] = 0;
        if (rel >= nk) continue;
        if (SMEM_TABLE) k_item[k0 + rel] = i;
        tin[j] = k_tins[k0 + rel];
        if (tin[j] == kNone) continue;
        const uint64_t hk = ((uint64_t)(rel / kB) << 32) | (rel % kB);
        const uint64_t bkt = (hk * 0x9E3779B97F4A7C15ull) % nbi;         // struct H, ops/arithmetic.hpp:73
        if (SMEM_TABLE) {
            bk[j] = (uint32_t)bkt;
            atomicMin(&s_tb[bk[j]], tin[j]);
        } else {
            const uint32_t size = toff[i + 1] - toff[i];
            uint32_t h = (uint32_t)(mix64(bkt) & (size - 1));
            while (t_key[toff[i] + h] != bkt) h = (h + 1) & (size - 1);
            bk[j] = t_val[toff[i] + h];                                  // the bucket's first-occupation time itself
        }
    }
    if (SMEM_TABLE) __syncthreads();
#pragma unroll
    for (int j = 0; j < kBlockSortItems; j++) {
        const uint32_t rel = threadIdx.x * kBlockSortItems + j;
        if (rel >= nk) key[j] = (1u << (2 * pbits)) | (pmax << pbits) | pmax;          // padding: after every real key
        else if (tin[j] == kNone) key[j] = (pmax << pbits) | pmax;
        else {
            const uint32_t tb = SMEM_TABLE ? s_tb[bk[j]] : bk[j];
            key[j] = ((pmax - (tb + 1)) << pbits) | (pmax - (tin[j] + 1));
        }
    }
    Sort(tmp).Sort(key, val, 0, 2 * pbits + 1);
#pragma unroll
    for (int j = 0; j < kBlockSortItems; j++) {
        const uint32_t r = threadIdx.x * kBlockSortItems + j;
        if (r < nk) sval[k0 + r] = val[j];
    }
}

// per sorted position: number of edges emitted (P if ip && wp != 0, M if im && wm != 0; ops/arithmetic.hpp:96-101)
__global__ void mul_emit_count_kernel(uint32_t nkeys, const uint32_t* __restrict__ sval, const uint8_t* __restrict__ k_flags, const Fp* __restrict__ k_wp,
                                      const Fp* __restrict__ k_wm, const uint32_t* __restrict__ k_tins, uint32_t* __restrict__ cnt) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nkeys) return;
    uint32_t k = sval[q];
    uint32_t c = 0;
    if (k_tins[k] != kNone) {
        uint8_t fl = k_flags[k];
        if ((fl & 1) && !fp_is_zero(k_wp[k])) c++;
        if ((fl & 2) && !fp_is_zero(k_wm[k])) c++;
    }
    cnt[q] = c;
}

__global__ void mul_eoff_kernel(uint64_t n, const uint32_t* __restrict__ koff, const uint32_t* __restrict__ epos, uint32_t nkeys, uint32_t total,
                                uint32_t* __restrict__ eoff, uint32_t budget, unsigned int* __restrict__ err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t k = koff[i];
    eoff[i] = (i == n || k >= nkeys) ? total : epos[k];
    if (i < n) {
        uint32_t k1 = koff[i + 1];
        uint32_t e1 = (k1 >= nkeys) ? total : epos[k1];
        if (e1 - eoff[i] > budget) atomicOr(err, 8u);
    }
}

__global__ void mul_emit_kernel(uint32_t nkeys, const __grid_constant__ TapeSpec ts, const uint32_t* __restrict__ sval, const uint32_t* __restrict__ epos,
                                const uint32_t* __restrict__ k_item, const uint32_t* __restrict__ koff, const uint32_t* __restrict__ eoff,
                                const uint32_t* __restrict__ loA, const uint32_t* __restrict__ loB, const uint32_t* __restrict__ loP,
                                const uint8_t* __restrict__ k_flags, const Fp* __restrict__ k_wp, const Fp* __restrict__ k_wm,
                                const uint32_t* __restrict__ k_tins, uint32_t* __restrict__ o_lid, uint16_t* __restrict__ o_idx, uint8_t* __restrict__ o_ch,
                                Fp* __restrict__ o_w, uint64_t* __restrict__ salt, uint32_t* __restrict__ seed_idx, unsigned int* __restrict__ err) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nkeys) return;
    uint32_t k = sval[q];
    if (k_tins[k] == kNone) return;
    uint32_t i = k_item[k];
    uint32_t rel = k - koff[i];
    uint32_t lp = rel / kB, s = rel % kB;
    uint32_t LA = loA[i + 1] - loA[i], LB = loB[i + 1] - loB[i];
    uint32_t lid = LA + LB + lp;
    Tape tp = tape_open(ts, i);
    const uint64_t w0 = tp.k;
    uint32_t e = epos[q];
    uint8_t fl = k_flags[k];
    for (int sg = 0; sg < 2; sg++) {
        Fp w = sg ? k_wm[k] : k_wp[k];
        if (!(fl & (1 << sg)) || fp_is_zero(w)) continue;
        o_lid[e] = lid; o_idx[e] = (uint16_t)s; o_ch[e] = (uint8_t)sg; o_w[e] = w;
        salt[e] = tp.at(w0 + 2ull * LA * LB + (e - eoff[i]));      // salts follow the 2*LA*LB nonce words, in emission order
        seed_idx[e] = loP[i] + lid;
        e++;
    }
    if (tp.overrun) atomicOr(err, 16u);
}

#define MUL_ALLOC(ptr, bytes)                                  \
    do {                                                       \
        if ((rc = scratch.alloc(ptr, (bytes)))) return rc;     \
    } while (0)

int op_ct_mul(Ctx* ctx, const Batch* A, const Batch* B, uint64_t batch_seed, const uint64_t* h_states, Batch** out) {
    const uint64_t n = A->n;
    int rc;
    if (n == 0) return batch_alloc(ctx, 0, 0, 0, out);
    Scratch scratch(ctx);                 // every temporary is freed (stream-ordered) on every return path

    uint32_t *loP, *lpoff, *koff, *toff;
    uint64_t* nb;
    unsigned int* err;
    MUL_ALLOC(loP, (n + 1) * 4); MUL_ALLOC(lpoff, (n + 1) * 4); MUL_ALLOC(koff, (n + 1) * 4); MUL_ALLOC(toff, (n + 1) * 4);
    MUL_ALLOC(nb, n * 8); MUL_ALLOC(err, 4);
    TapeSpec ts;
    if ((rc = tape_spec(ctx, scratch, n, batch_seed, h_states, nullptr, nullptr, ts))) return rc;
    PV_CUDA(cudaMemsetAsync(err, 0, 4, ctx->stream));
    mul_count_scan_kernel<<<1, 1024, 0, ctx->stream>>>(n, A->loff, B->loff, A->eoff, B->eoff, ctx->d_primes, ctx->n_primes, loP, lpoff, koff, toff, nb, ctx->d_mail);
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    uint32_t mail[15];
    memcpy(mail, ctx->h_mail, sizeof mail);
    unsigned long long h_stats[5];
    for (int k = 0; k < 5; k++) h_stats[k] = (unsigned long long)mail[4 + 2 * k] | ((unsigned long long)mail[5 + 2 * k] << 32);
    unsigned int h_err = mail[14];
    if (h_err) { ctx->last_error = "ct_mul: operand too large for the batched path (layer pairs / keys / edge pairs overflow)"; return PV_E_SHAPE; }
    const uint32_t nLpre = mail[0], nLP = mail[1], nKeys = mail[2], nTbl = mail[3];
    const unsigned long long h_maxpairs = h_stats[0];
    if (h_stats[1] >= (1ull << 31) || h_stats[2] >= (1ull << 32)) {   // 32-bit offsets index the key space
        ctx->last_error = "ct_mul: batch too large (key space needs more than 2^31 slots); split it into smaller tiles";
        return PV_E_SHAPE;
    }

    // ---- pre-compaction layers
    uint8_t* p_rule; uint64_t *p_zt, *p_nl, *p_nh; uint32_t *p_pa, *p_pb, *lp_item;
    MUL_ALLOC(p_rule, nLpre); MUL_ALLOC(p_zt, (size_t)nLpre * 8); MUL_ALLOC(p_nl, (size_t)nLpre * 8); MUL_ALLOC(p_nh, (size_t)nLpre * 8);
    MUL_ALLOC(p_pa, (size_t)nLpre * 4); MUL_ALLOC(p_pb, (size_t)nLpre * 4); MUL_ALLOC(lp_item, (size_t)(nLP ? nLP : 1) * 4);
    mul_layers_kernel<<<(unsigned)n, 128, 0, ctx->stream>>>(n, ts, ctx->kv.canon_tag, A->loff, B->loff, A->rule, A->ztag, A->nlo, A->nhi, A->pa, A->pb,
                                                            B->rule, B->ztag, B->nlo, B->nhi, B->pa, B->pb, loP, lpoff, p_rule, p_zt, p_nl, p_nh, p_pa, p_pb, lp_item);
    // ---- edges by layer
    uint32_t *lsA, *lcA, *curA, *ordA, *lsB, *lcB, *curB, *ordB;
    MUL_ALLOC(lsA, (A->nL + 1) * 4); MUL_ALLOC(lcA, (A->nL + 1) * 4); MUL_ALLOC(curA, (A->nL + 1) * 4); MUL_ALLOC(ordA, (A->nE + 1) * 4);
    MUL_ALLOC(lsB, (B->nL + 1) * 4); MUL_ALLOC(lcB, (B->nL + 1) * 4); MUL_ALLOC(curB, (B->nL + 1) * 4); MUL_ALLOC(ordB, (B->nE + 1) * 4);
    mul_bylayer_kernel<<<(unsigned)n, 128, 0, ctx->stream>>>(A->loff, A->eoff, A->lid, lsA, lcA, curA, ordA, err);
    mul_bylayer_kernel<<<(unsigned)n, 128, 0, ctx->stream>>>(B->loff, B->eoff, B->lid, lsB, lcB, curB, ordB, err);
    ctx->stat_kernel_launches += 3;

    Batch* o = nullptr;
    uint32_t nEout = 0;
    uint64_t* salt = nullptr;
    uint32_t* seed_idx = nullptr;
    if (nKeys) {
        // ---- pair products per key
        Fp *k_wp, *k_wm; uint8_t* k_flags; uint32_t *k_tins, *k_item, *t_val; unsigned long long* t_key;
        MUL_ALLOC(k_wp, (size_t)nKeys * 16); MUL_ALLOC(k_wm, (size_t)nKeys * 16); MUL_ALLOC(k_flags, nKeys); MUL_ALLOC(k_tins, (size_t)nKeys * 4);
        MUL_ALLOC(k_item, (size_t)nKeys * 4); MUL_ALLOC(t_key, (size_t)(nTbl ? nTbl : 1) * 8); MUL_ALLOC(t_val, (size_t)(nTbl ? nTbl : 1) * 4);
        EdgeView VA{A->loff, A->eoff, A->idx, A->ch, A->w, lsA, lcA, ordA};
        EdgeView VB{B->loff, B->eoff, B->idx, B->ch, B->w, lsB, lcB, ordB};
        {
            ProfScope ps(ctx, PROF_MUL_PAIRS);
            mul_pairs_kernel<<<nLP, kPairThreads, 0, ctx->stream>>>(VA, VB, lp_item, lpoff, koff, k_wp, k_wm, k_flags, k_tins, err);
        }
        const unsigned kb = (nKeys + 255) / 256;
        // ---- libstdc++ iteration order by radix sort
        int pbits = 1;
        while ((1ull << pbits) - 1 < h_maxpairs + 1) pbits++;
        int ibits = 1;
        while ((1ull << ibits) < n) ibits++;
        if (2 * pbits + ibits > 64) { ctx->last_error = "ct_mul: sort key does not fit 64 bits (batch too large for these operand sizes)"; return PV_E_SHAPE; }
        uint64_t *skey = nullptr, *skey2 = nullptr; uint32_t *sval = nullptr, *sval2, *ecnt, *epos;
        MUL_ALLOC(sval2, (size_t)nKeys * 4); MUL_ALLOC(ecnt, (size_t)nKeys * 4); MUL_ALLOC(epos, (size_t)nKeys * 4);
        size_t tmp_bytes = 0, tmp2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp2, ecnt, epos, (int)nKeys, ctx->stream);
        const bool block_sort = !ctx->mul_force_device_sort && h_stats[3] <= (unsigned long long)kBlockSortKeys && 2 * pbits + 1 <= 32;
        const bool smem_table = block_sort && !ctx->mul_force_global_table && h_stats[4] <= (unsigned long long)kSmemBuckets;
        if (!smem_table) {
            // first-occupation time of every hash bucket in an open-addressing table in global memory
            PV_CUDA(cudaMemsetAsync(t_key, 0xFF, (size_t)(nTbl ? nTbl : 1) * 8, ctx->stream));
            PV_CUDA(cudaMemsetAsync(t_val, 0xFF, (size_t)(nTbl ? nTbl : 1) * 4, ctx->stream));
            mul_bucket_insert_kernel<<<kb, 256, 0, ctx->stream>>>(nKeys, (uint32_t)n, koff, toff, nb, k_tins, k_item, t_key, t_val);
        }
        void* tmp;
        if (block_sort) {
            tmp_bytes = tmp2;
            MUL_ALLOC(tmp, tmp_bytes);
            if (smem_table) mul_sort_block_kernel<true><<<(unsigned)n, kBlockSortThreads, 0, ctx->stream>>>(koff, toff, nb, k_tins, t_key, t_val, pbits, sval2, k_item);
            else mul_sort_block_kernel<false><<<(unsigned)n, kBlockSortThreads, 0, ctx->stream>>>(koff, toff, nb, k_tins, t_key, t_val, pbits, sval2, k_item);
        } else {
            MUL_ALLOC(skey, (size_t)nKeys * 8); MUL_ALLOC(skey2, (size_t)nKeys * 8); MUL_ALLOC(sval, (size_t)nKeys * 4);
            mul_sortkey_kernel<<<kb, 256, 0, ctx->stream>>>(nKeys, koff, toff, nb, k_tins, k_item, t_key, t_val, pbits, skey, sval);
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, skey, skey2, sval, sval2, (int)nKeys, 0, 2 * pbits + ibits, ctx->stream);
            if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
            MUL_ALLOC(tmp, tmp_bytes);
            PV_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, skey, skey2, sval, sval2, (int)nKeys, 0, 2 * pbits + ibits, ctx->stream));
        }
        mul_emit_count_kernel<<<kb, 256, 0, ctx->stream>>>(nKeys, sval2, k_flags, k_wp, k_wm, k_tins, ecnt);
        PV_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, ecnt, epos, (int)nKeys, ctx->stream));
        uint32_t last[2] = {0, 0};
        {
            SmallRead sr;
            sr.add(&last[0], epos + (nKeys - 1), 4); sr.add(&last[1], ecnt + (nKeys - 1), 4); sr.add(&h_err, err, 4);
            if ((rc = read_small_sync(ctx, sr))) return rc;
        }
        // pairs, [bucket_insert], sort (block kernel, or sortkey + cub histogram, exclusive sum, one onesweep pass per 8 key bits), emit_count, cub scan (init + scan)
        ctx->stat_kernel_launches += 1 /* pairs */ + (smem_table ? 0 : 1) /* bucket_insert */ + 1 /* emit_count */ + 2 /* cub scan */ +
                                     (block_sort ? 1 : 1 + 2 + (uint64_t)((2 * pbits + ibits + 7) / 8)) /* block sort | sortkey + cub radix sort */;
        if (h_err & 2) { ctx->last_error = "ct_mul: edge layer id out of range"; return PV_E_FORMAT; }
        nEout = last[0] + last[1];
        if ((rc = batch_alloc(ctx, n, nLpre, nEout, &o))) return rc;
        if ((rc = scratch.alloc(salt, (size_t)(nEout ? nEout : 1) * 8)) || (rc = scratch.alloc(seed_idx, (size_t)(nEout ? nEout : 1) * 4))) { batch_free(o); return rc; }
        mul_eoff_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(n, koff, epos, nKeys, nEout, o->eoff, ctx->edge_budget, err);
        mul_emit_kernel<<<kb, 256, 0, ctx->stream>>>(nKeys, ts, sval2, epos, k_item, koff, o->eoff, A->loff, B->loff, loP, k_flags, k_wp, k_wm, k_tins,
                                                     o->lid, o->idx, o->ch, o->w, salt, seed_idx, err);
        ctx->stat_kernel_launches += 2;
    } else {
        if ((rc = batch_alloc(ctx, n, nLpre, 0, &o))) return rc;
        PV_CUDA(cudaMemsetAsync(o->eoff, 0, (n + 1) * 4, ctx->stream));
    }
    auto fail = [&](int code) { batch_free(o); return code; };
    // layers of the result (pre-compaction)
    cudaError_t ce = cudaSuccess;
    auto d2d = [&](void* dst, const void* src, size_t bytes) { if (ce == cudaSuccess && bytes) ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream); };
    d2d(o->loff, loP, (n + 1) * 4); d2d(o->rule, p_rule, nLpre); d2d(o->ztag, p_zt, (size_t)nLpre * 8); d2d(o->nlo, p_nl, (size_t)nLpre * 8);
    d2d(o->nhi, p_nh, (size_t)nLpre * 8); d2d(o->pa, p_pa, (size_t)nLpre * 4); d2d(o->pb, p_pb, (size_t)nLpre * 4);
    if (ce != cudaSuccess) { ctx->last_error = std::string("ct_mul: ") + cudaGetErrorString(ce); return fail(PV_E_CUDA); }
    // ---- sigma of every output edge, seeded by its PROD layer (pre-compaction tables)
    if (nEout) {
        SigmaJobs J;
        J.n = nEout; J.ztag = p_zt; J.nlo = p_nl; J.nhi = p_nh; J.seed_idx = seed_idx; J.idx = o->idx; J.ch = o->ch; J.salt = salt; J.out = o->sigma;
        if ((rc = sigma_run(ctx, J))) return fail(rc);
    }
    // guard_budget(pk, C, "mul") then compact_layers(C), ops/arithmetic.hpp:103-104. Only a batch with more than edge_budget edges
    // in total can hold a ciphertext over the budget; otherwise the error bits of the kernels above ride along with the mailbox
    // read compact_layers needs anyway (one synchronisation instead of two).
    if (nEout > ctx->edge_budget) {
        { SmallRead sr; sr.add(&h_err, err, 4); if ((rc = read_small_sync(ctx, sr))) return fail(rc); }
        if ((h_err & 8) && (rc = guard_budget_batch(ctx, &o, ctx->edge_budget))) return fail(rc);
        if ((rc = compact_layers_batch(ctx, o))) return fail(rc);
    } else if ((rc = compact_layers_batch(ctx, o, err, &h_err))) return fail(rc);
    if (h_err & 16) { ctx->last_error = "ct_mul: the supplied tape words (pvacb_set_tape_words) ran out"; return fail(PV_E_ARG); }
    *out = o;
    return PV_OK;
}

}  // namespace pvacb
