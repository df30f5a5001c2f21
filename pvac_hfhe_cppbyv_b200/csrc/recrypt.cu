// ct_recrypt, ubk_apply, sigma_density, make_evalkey (ops/recrypt.hpp:12-41, crypto/matrix.hpp:95-188,306-310,
// ops/encrypt.hpp:29-37) for device batches, plus batch_select, the generic "pick items from several batches" primitive.
//
// ct_recrypt(pk, ek, C): while sigma_density(C) is outside [0.495, 0.505] (at most 8 times): add a random enc-of-zero from the
// pool, permute every sigma by the public permutation UBK; then compact_edges + compact_layers. With real ciphertexts the
// density is 0.499-0.5 and the loop never runs, so the common path is compact_edges + compact_layers; the
// loop is still restated exactly (per-ciphertext iteration counts, pool indices from the item's tape) and tested with
// crafted all-zero syndromes. Items the reference returns unchanged (no edges, or an empty pool) are returned unchanged.
#include "engine.h"
#include "sha256.cuh"
#include "../../include/pvacb.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace pvacb {

// ------------------------------------------------------------------ UBK: public permutation of the 8192 sigma bits
// crypto/matrix.hpp:95-164: Fisher-Yates driven by SHA-256("UBK" || LE64(canon_tag) || LE64(ctr)), 4 words per hash,
// bounded(M): accept x <= 2^64-1 - ((2^64-1) % M). Runs once per key set on the host (2 048 hashes).
void gen_ubk_perm_host(uint64_t canon_tag, uint16_t perm[kMBits]) {
    for (int i = 0; i < kMBits; i++) perm[i] = (uint16_t)i;
    uint64_t ctr = 0, words[4];
    int have = 4;
    auto next_word = [&]() -> uint64_t {
        if (have >= 4) {
            uint8_t msg[64];
            memset(msg, 0, sizeof msg);
            memcpy(msg, "UBK", 3);
            for (int i = 0; i < 8; i++) { msg[3 + i] = (uint8_t)(canon_tag >> (8 * i)); msg[11 + i] = (uint8_t)(ctr >> (8 * i)); }
            msg[19] = 0x80;
            const uint64_t bits = 19 * 8;
            for (int i = 0; i < 8; i++) msg[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
            uint32_t w[16];
            for (int i = 0; i < 16; i++) w[i] = ((uint32_t)msg[4 * i] << 24) | ((uint32_t)msg[4 * i + 1] << 16) | ((uint32_t)msg[4 * i + 2] << 8) | msg[4 * i + 3];
            ShaState st;
            sha_init(st);
            sha_compress(st, w);
            for (int k = 0; k < 4; k++) words[k] = sha_digest_le64(st, k);
            ctr++;
            have = 0;
        }
        return words[have++];
    };
    for (int i = kMBits - 1; i > 0; --i) {
        const uint64_t M = (uint64_t)i + 1;
        const uint64_t lim = ~0ull - (~0ull % M);
        uint64_t x;
        do { x = next_word(); } while (x > lim);
        const int j = (int)(x % M);
        std::swap(perm[i], perm[j]);
    }
}

// one warp per edge: out bit j = in bit perm[j]  (apply_perm_sigma sets out[inv[src]] for every set bit src)
__global__ void __launch_bounds__(256) ubk_apply_kernel(uint64_t nE, const uint16_t* __restrict__ perm, const uint64_t* __restrict__ in, uint64_t* __restrict__ out) {
    __shared__ uint64_t rows[8][kMWords];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t e = (uint64_t)blockIdx.x * 8 + wid;
    if (e >= nE) return;
    for (int k = lane; k < kMWords; k += 32) rows[wid][k] = in[e * kMWords + k];
    __syncwarp();
    for (int wq = 0; wq < 4; wq++) {              // lane produces output words lane + 32 * wq
        const int ow = lane + 32 * wq;
        uint64_t v = 0;
        for (int b = 0; b < 64; b++) {
            const uint32_t src = perm[ow * 64 + b];
            v |= ((rows[wid][src >> 6] >> (src & 63)) & 1ull) << b;
        }
        out[e * kMWords + ow] = v;
    }
}

// one CTA per ciphertext: number of set sigma bits (sigma_density's numerator)
__global__ void __launch_bounds__(256) sigma_ones_kernel(const uint32_t* __restrict__ eoff, const uint64_t* __restrict__ sigma, unsigned long long* __restrict__ ones) {
    const uint64_t i = blockIdx.x;
    const uint64_t w0 = (uint64_t)eoff[i] * kMWords, w1 = (uint64_t)eoff[i + 1] * kMWords;
    unsigned long long s = 0;
    for (uint64_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) s += __popcll(sigma[w]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < 8; k++) t += part[k];
        ones[i] = t;
    }
}

int sigma_ones(Ctx* ctx, const Batch* b, std::vector<uint64_t>& ones, std::vector<uint32_t>& nedges) {
    ones.assign(b->n, 0);
    nedges.assign(b->n, 0);
    if (b->n == 0) return PV_OK;
    Scratch scratch(ctx);
    unsigned long long* d = nullptr;
    int rc;
    if ((rc = scratch.alloc(d, b->n * 8))) return rc;
    sigma_ones_kernel<<<(unsigned)b->n, 256, 0, ctx->stream>>>(b->eoff, b->sigma, d);
    PV_CUDA(cudaGetLastError());
    std::vector<uint32_t> eo(b->n + 1);
    PV_CUDA(cudaMemcpyAsync(ones.data(), d, b->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(eo.data(), b->eoff, (b->n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint64_t i = 0; i < b->n; i++) nedges[i] = eo[i + 1] - eo[i];
    ctx->stat_kernel_launches += 1;
    return PV_OK;
}

// ops/encrypt.hpp:29-37 with the same (long double) arithmetic
static double density_of(uint64_t ones, uint32_t nedges) {
    if (nedges == 0) return 0.0;
    long double o = (long double)ones, t = (long double)nedges * (long double)kMBits;
    return (double)(o / t);
}

// ------------------------------------------------------------------ batch_select
struct SelView {
    const uint32_t *loff, *eoff;
    const uint8_t* rule;
    const uint64_t *ztag, *nlo, *nhi;
    const uint32_t *pa, *pb, *lid;
    const uint16_t* idx;
    const uint8_t* ch;
    const Fp* w;
    const uint64_t* sigma;
};
constexpr int kSelMaxSrc = 4;
struct SelSources { SelView v[kSelMaxSrc]; };

__global__ void sel_count_kernel(uint64_t n, SelSources S, const uint32_t* __restrict__ which, const uint32_t* __restrict__ index, uint32_t* __restrict__ cl,
                                 uint32_t* __restrict__ ce) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SelView& v = S.v[which[i]];
    cl[i] = v.loff[index[i] + 1] - v.loff[index[i]];
    ce[i] = v.eoff[index[i] + 1] - v.eoff[index[i]];
}
__global__ void __launch_bounds__(256)
sel_copy_kernel(SelSources S, const uint32_t* __restrict__ which, const uint32_t* __restrict__ index, const uint32_t* __restrict__ oloff, const uint32_t* __restrict__ oeoff,
                uint8_t* __restrict__ o_rule, uint64_t* __restrict__ o_ztag, uint64_t* __restrict__ o_nlo, uint64_t* __restrict__ o_nhi, uint32_t* __restrict__ o_pa,
                uint32_t* __restrict__ o_pb, uint32_t* __restrict__ o_lid, uint16_t* __restrict__ o_idx, uint8_t* __restrict__ o_ch, Fp* __restrict__ o_w,
                uint64_t* __restrict__ o_sigma) {
    const uint64_t i = blockIdx.x;
    const SelView& v = S.v[which[i]];
    const uint32_t l0 = v.loff[index[i]], L = v.loff[index[i] + 1] - l0, e0 = v.eoff[index[i]], E = v.eoff[index[i] + 1] - e0;
    const uint32_t ol = oloff[i], oe = oeoff[i];
    if (blockIdx.y == 0)
        for (uint32_t k = threadIdx.x; k < L; k += blockDim.x) {
            o_rule[ol + k] = v.rule[l0 + k]; o_ztag[ol + k] = v.ztag[l0 + k]; o_nlo[ol + k] = v.nlo[l0 + k]; o_nhi[ol + k] = v.nhi[l0 + k];
            o_pa[ol + k] = v.pa[l0 + k]; o_pb[ol + k] = v.pb[l0 + k];
        }
    for (uint32_t k = blockIdx.y * blockDim.x + threadIdx.x; k < E; k += gridDim.y * blockDim.x) {
        o_lid[oe + k] = v.lid[e0 + k]; o_idx[oe + k] = v.idx[e0 + k]; o_ch[oe + k] = v.ch[e0 + k]; o_w[oe + k] = v.w[e0 + k];
    }
    const uint4* s = reinterpret_cast<const uint4*>(v.sigma + (size_t)e0 * kMWords);
    uint4* d = reinterpret_cast<uint4*>(o_sigma + (size_t)oe * kMWords);
    const uint64_t nvec = (uint64_t)E * 64;
    for (uint64_t q = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x; q < nvec; q += (uint64_t)gridDim.y * blockDim.x) d[q] = s[q];
}

// out item i = item index[i] of srcs[which[i]]  (host arrays). Up to kSelMaxSrc sources.
int batch_select(Ctx* ctx, const Batch* const* srcs, int nsrc, const std::vector<uint32_t>& which, const std::vector<uint32_t>& index, Batch** out) {
    const uint64_t n = which.size();
    if (nsrc > kSelMaxSrc || index.size() != n) return PV_E_ARG;
    if (n == 0) return batch_alloc(ctx, 0, 0, 0, out);
    SelSources S;
    memset(&S, 0, sizeof S);
    for (int k = 0; k < nsrc; k++) {
        const Batch* b = srcs[k];
        S.v[k] = SelView{b->loff, b->eoff, b->rule, b->ztag, b->nlo, b->nhi, b->pa, b->pb, b->lid, b->idx, b->ch, b->w, b->sigma};
    }
    for (uint64_t i = 0; i < n; i++)
        if ((int)which[i] >= nsrc || index[i] >= srcs[which[i]]->n) return PV_E_ARG;
    Scratch scratch(ctx);
    uint32_t *d_which, *d_index, *cl, *ce, *ol, *oe;
    int rc;
    if ((rc = scratch.alloc(d_which, n * 4)) || (rc = scratch.alloc(d_index, n * 4)) || (rc = scratch.alloc(cl, n * 4)) || (rc = scratch.alloc(ce, n * 4)) ||
        (rc = scratch.alloc(ol, (n + 1) * 4)) || (rc = scratch.alloc(oe, (n + 1) * 4)))
        return rc;
    PV_CUDA(cudaMemcpyAsync(d_which, which.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(d_index, index.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
    sel_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, S, d_which, d_index, cl, ce);
    if ((rc = scan_u32(ctx, n, cl, ol)) || (rc = scan_u32(ctx, n, ce, oe))) return rc;
    uint32_t tot[2] = {0, 0};
    { SmallRead sr; sr.add(&tot[0], ol + n, 4); sr.add(&tot[1], oe + n, 4); if ((rc = read_small_sync(ctx, sr))) return rc; }
    Batch* o = nullptr;
    if ((rc = batch_alloc(ctx, n, tot[0], tot[1], &o))) return rc;
    PV_CUDA(cudaMemcpyAsync(o->loff, ol, (n + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(o->eoff, oe, (n + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    const uint64_t avg = tot[1] / n + 1;
    const unsigned ych = (unsigned)std::min<uint64_t>(std::max<uint64_t>(avg / 64, 1), 1024);
    sel_copy_kernel<<<dim3((unsigned)n, ych), 256, 0, ctx->stream>>>(S, d_which, d_index, o->loff, o->eoff, o->rule, o->ztag, o->nlo, o->nhi, o->pa, o->pb, o->lid,
                                                                      o->idx, o->ch, o->w, o->sigma);
    PV_CUDA(cudaGetLastError());
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stat_kernel_launches += 2;
    *out = o;
    return PV_OK;
}

// ------------------------------------------------------------------ ubk_apply on a whole batch (new batch)
int op_ubk_apply(Ctx* ctx, const Batch* b, Batch** out) {
    Batch* o = nullptr;
    int rc = batch_clone(ctx, b, &o);
    if (rc) return rc;
    if (b->nE) {
        ubk_apply_kernel<<<(unsigned)((b->nE + 7) / 8), 256, 0, ctx->stream>>>(b->nE, ctx->d_ubk_perm, b->sigma, o->sigma);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { batch_free(o); ctx->last_error = cudaGetErrorString(e); return PV_E_CUDA; }
        ctx->stat_kernel_launches += 1;
    }
    *out = o;
    return PV_OK;
}

// ------------------------------------------------------------------ ct_recrypt
int op_ct_recrypt(Ctx* ctx, const Batch* in, const Batch* pool, uint64_t batch_seed, const uint64_t* h_states, Batch** out) {
    const uint64_t n = in->n;
    int rc;
    std::vector<uint64_t> ones;
    std::vector<uint32_t> ne;
    if ((rc = sigma_ones(ctx, in, ones, ne))) return rc;
    // items the reference returns as they are (ops/recrypt.hpp:27)
    std::vector<uint8_t> untouched(n, 0);
    for (uint64_t i = 0; i < n; i++) untouched[i] = (pool->n == 0 || ne[i] == 0) ? 1 : 0;
    Batch* cur = nullptr;
    if ((rc = batch_clone(ctx, in, &cur))) return rc;
    std::vector<uint64_t> draws(n, 0);
    TapeSpec ts;                            // the pool indices are drawn on the host, item i from word draws[i] of its stream
    ts.kind = ctx->tape_kind;
    for (int k = 0; k < 8; k++) ts.key[k] = ctx->tape_key[k];
    ts.batch_seed = batch_seed; ts.item_base = ctx->item_base; ts.states = h_states;
    std::vector<uint64_t> h_words;
    if (ts.kind == TAPE_WORDS) {
        if (!ctx->d_tape_words || ctx->item_base + n > ctx->tape_words_items) { ctx->last_error = "tape kind WORDS: no words supplied for these items"; return PV_E_ARG; }
        h_words.resize(ctx->tape_words_items * ctx->tape_words_per_item);
        PV_CUDA(cudaMemcpy(h_words.data(), ctx->d_tape_words, h_words.size() * 8, cudaMemcpyDeviceToHost));
        ts.words = h_words.data(); ts.words_per_item = ctx->tape_words_per_item;
    }
    auto needs = [&](uint64_t i) { double d = density_of(ones[i], ne[i]); return d < 0.495 || d > 0.505; };
    for (int it = 0; it < 8; it++) {
        std::vector<uint32_t> act;
        for (uint64_t i = 0; i < n; i++)
            if (!untouched[i] && needs(i)) act.push_back((uint32_t)i);
        if (act.empty()) break;
        // result = ct_add(result, zero_pool[csprng % size]); ubk_apply; guard_budget   -- for the active items only
        std::vector<uint32_t> w0(act.size(), 0), w1(act.size(), 1), zi(act.size());
        for (size_t q = 0; q < act.size(); q++) {
            const uint64_t i = act[q];
            Tape tp = tape_open(ts, i);
            zi[q] = (uint32_t)(tp.at(draws[i]++) % pool->n);
            if (tp.overrun) { batch_free(cur); ctx->last_error = "ct_recrypt: the supplied tape words ran out"; return PV_E_ARG; }
        }
        const Batch* srcs[2] = {cur, pool};
        Batch *ra = nullptr, *za = nullptr, *sum = nullptr, *perm = nullptr;
        if ((rc = batch_select(ctx, srcs, 2, w0, act, &ra))) { batch_free(cur); return rc; }
        if ((rc = batch_select(ctx, srcs, 2, w1, zi, &za))) { batch_free(cur); batch_free(ra); return rc; }
        rc = op_ct_add(ctx, ra, za, 0, &sum);
        batch_free(ra); batch_free(za);
        if (rc) { batch_free(cur); return rc; }
        rc = op_ubk_apply(ctx, sum, &perm);
        batch_free(sum);
        if (rc) { batch_free(cur); return rc; }
        if ((rc = guard_budget_batch(ctx, &perm, ctx->edge_budget))) { batch_free(cur); batch_free(perm); return rc; }
        // put them back in place
        std::vector<uint32_t> which(n, 0), index(n);
        for (uint64_t i = 0; i < n; i++) index[i] = (uint32_t)i;
        for (size_t q = 0; q < act.size(); q++) { which[act[q]] = 1; index[act[q]] = (uint32_t)q; }
        const Batch* srcs2[2] = {cur, perm};
        Batch* next = nullptr;
        rc = batch_select(ctx, srcs2, 2, which, index, &next);
        batch_free(cur); batch_free(perm);
        if (rc) return rc;
        cur = next;
        if ((rc = sigma_ones(ctx, cur, ones, ne))) { batch_free(cur); return rc; }
    }
    // compact_edges + compact_layers on everything that was not returned early
    bool any_untouched = false, any_touched = false;
    for (uint64_t i = 0; i < n; i++) (untouched[i] ? any_untouched : any_touched) = true;
    if (any_touched) {
        if (!any_untouched) {
            if ((rc = guard_budget_batch(ctx, &cur, 0)) || (rc = compact_layers_batch(ctx, cur))) { batch_free(cur); return rc; }
        } else {
            std::vector<uint32_t> t_idx, which(n, 0), index(n);
            for (uint64_t i = 0; i < n; i++) { index[i] = (uint32_t)i; if (!untouched[i]) { which[i] = 1; index[i] = (uint32_t)t_idx.size(); t_idx.push_back((uint32_t)i); } }
            std::vector<uint32_t> w0(t_idx.size(), 0);
            const Batch* s1[1] = {cur};
            Batch* part = nullptr;
            if ((rc = batch_select(ctx, s1, 1, w0, t_idx, &part))) { batch_free(cur); return rc; }
            if ((rc = guard_budget_batch(ctx, &part, 0)) || (rc = compact_layers_batch(ctx, part))) { batch_free(cur); batch_free(part); return rc; }
            const Batch* s2[2] = {cur, part};
            Batch* merged = nullptr;
            rc = batch_select(ctx, s2, 2, which, index, &merged);
            batch_free(cur); batch_free(part);
            if (rc) return rc;
            cur = merged;
        }
    }
    *out = cur;
    return PV_OK;
}

}  // namespace pvacb

using namespace pvacb;
static inline Ctx* C(pvacb_ctx* x) { return reinterpret_cast<Ctx*>(x); }
static inline const Batch* Bt(const pvacb_batch* x) { return reinterpret_cast<const Batch*>(x); }

extern "C" {

int pvacb_sigma_density(pvacb_ctx* x, const pvacb_batch* pb, double* out) {
    Ctx* ctx = C(x);
    if (!out) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    std::vector<uint64_t> ones;
    std::vector<uint32_t> ne;
    int rc = sigma_ones(ctx, Bt(pb), ones, ne);
    if (rc) return rc;
    for (size_t i = 0; i < ones.size(); i++) out[i] = density_of(ones[i], ne[i]);
    return PV_OK;
}

int pvacb_ubk_apply(pvacb_ctx* x, const pvacb_batch* pb, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    cudaSetDevice(ctx->device);
    Batch* o = nullptr;
    int rc = op_ubk_apply(ctx, Bt(pb), &o);
    *out = reinterpret_cast<pvacb_batch*>(o);
    return rc;
}

int pvacb_ubk_perm(pvacb_ctx* x, uint16_t* perm_out) {
    Ctx* ctx = C(x);
    if (!perm_out) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    memcpy(perm_out, ctx->h_ubk_perm.data(), kMBits * 2);
    return PV_OK;
}

int pvacb_ct_recrypt(pvacb_ctx* x, const pvacb_batch* in, const pvacb_batch* zero_pool, uint64_t batch_seed, const uint64_t* tape_states, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out || !in || !zero_pool) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    cudaSetDevice(ctx->device);
    Batch* o = nullptr;
    int rc = op_ct_recrypt(ctx, Bt(in), Bt(zero_pool), batch_seed, tape_states, &o);
    *out = reinterpret_cast<pvacb_batch*>(o);
    return rc;
}

int pvacb_batch_select(pvacb_ctx* x, const pvacb_batch* const* srcs, int nsrc, const uint32_t* which, const uint32_t* index, size_t n, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out || !srcs || (n && (!which || !index)) || nsrc < 1 || nsrc > kSelMaxSrc) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    const Batch* s[kSelMaxSrc];
    for (int k = 0; k < nsrc; k++) { if (!srcs[k]) return PV_E_ARG; s[k] = Bt(srcs[k]); }
    Batch* o = nullptr;
    int rc = batch_select(ctx, s, nsrc, std::vector<uint32_t>(which, which + n), std::vector<uint32_t>(index, index + n), &o);
    *out = reinterpret_cast<pvacb_batch*>(o);
    return rc;
}

}  // extern "C"
