// enc_value (ops/encrypt.hpp:162-291) for a batch of plaintexts.
//
// The reference interleaves tape draws, field arithmetic, PRF calls and sigma generation per edge. No draw depends on a
// PRF output, so the batched engine splits it into data-parallel stages:
//   enc_plan_kernel     one thread per ciphertext walks the item's RNG tape in the reference's exact order (mask; then
//                       enc_fp_depth(-mask), then enc_fp_depth(v+mask): g++ evaluates combine_ciphers' second argument first)
//                       and records indices, signs, random coefficients, salts, the compact_edges grouping and the
//                       Fisher-Yates permutation;
//   prf_run             prf_R of both shares and prf_noise_delta of all but the last noise group (prf.cu);
//   enc_emit_kernel     layers, edge index/sign fields, sigma jobs (one per raw edge) and merge fix-ups;
//   enc_weights_kernel  solves the signal / Z2 / Z3 relations and multiplies by R (g^-j = g^(B-j): no inversion needed);
//   sigma_run           sigma_from_H of every raw edge, written straight to its final (shuffled) row (sigma.cu).
// compact_edges' value-dependent branch (an edge whose merged weight AND merged syndrome are both zero is removed, :59) is
// detected after the fact (enc_dropcheck_kernel); the items it hits are planned again with those slots removed (flagged slow
// path in op_enc_value) -- a removed slot shortens the shuffle and shifts every later tape word of the item.
#include "engine.h"
#include "enc_plan.cuh"

#include <algorithm>
#include <cmath>
#include <map>
#include <array>

namespace pvacb {

// share s = 0 is enc_fp_depth(-mask) (drawn first, becomes layer 1 / the trailing edges), s = 1 is enc_fp_depth(v+mask)
__global__ void enc_plan_kernel(uint64_t n, const uint64_t* __restrict__ values, const __grid_constant__ TapeSpec ts,
                                uint64_t canon_tag, int Z2, int Z3, int S, const uint64_t* __restrict__ drop,
                                SharePlan* __restrict__ plans, uint8_t* __restrict__ slabs, uint32_t* __restrict__ n_edges, uint32_t* __restrict__ n_extra,
                                uint64_t* __restrict__ j_ztag, uint64_t* __restrict__ j_nlo, uint64_t* __restrict__ j_nhi, uint8_t* __restrict__ j_flags,
                                uint64_t* __restrict__ next_word /* optional: first unused tape word per item */, unsigned int* __restrict__ err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int RAW = plan_raw_edges(Z2, Z3), RND = plan_rnd_values(Z2, Z3);
    const size_t slab = plan_slab_bytes(RAW, RND);
    for (int s = 0; s < S; s++) plan_bind(plans[S * i + s], slabs + (S * i + s) * slab, RAW, RND);
    Tape t = tape_open(ts, i);
    uint64_t used;
    const uint64_t* dm = drop ? drop + i * (uint64_t)(S * kKeyWords) : nullptr;
    if (S == 2) used = plan_item(t, values[i], canon_tag, Z2, Z3, plans[2 * i], plans[2 * i + 1], dm);
    else used = plan_single(t, fp_make(values[2 * i], values[2 * i + 1]), canon_tag, Z2, Z3, plans[i], dm);     // enc_fp_depth: values are Fp (lo, hi)
    if (t.overrun) atomicOr(err, 2u);
    if (next_word) next_word[i] = used;
    const int G = Z2 + Z3;
    uint32_t edges = 0, extra = 0;
    for (int s = 0; s < S; s++) {
        const SharePlan& P = plans[S * i + s];
        edges += P.n_out;
        extra += P.n_raw - P.n_out;
        // PRF jobs of the share: prf_R(seed), then prf_noise_delta(seed, gid, kind) for all but the last group
        uint64_t jb = ((uint64_t)S * i + s) * (uint64_t)G;
        j_ztag[jb] = P.ztag; j_nlo[jb] = P.nlo; j_nhi[jb] = P.nhi; j_flags[jb] = 2;
        for (int gid = 0; gid + 1 < G; gid++) {
            uint64_t g = (uint64_t)gid + 1, k = (gid < Z2 ? 0ull : 1ull) + 1;   // ops/encrypt.hpp:114-129
            j_nlo[jb + 1 + gid] = P.nlo ^ (0x9e3779b97f4a7c15ull * g) ^ k;
            j_nhi[jb + 1 + gid] = P.nhi ^ (0x94d049bb133111ebull * g) ^ (k << 32);
            j_ztag[jb + 1 + gid] = P.ztag ^ (0x517cc1b727220a95ull * g) ^ (k << 48);
            j_flags[jb + 1 + gid] = 3;
        }
    }
    n_edges[i] = edges;
    n_extra[i] = extra;
}

constexpr uint32_t kNoFix = 0xFFFFFFFFu;

// thread per share: layers, final edge fields, sigma jobs
__global__ void enc_emit_kernel(uint64_t n, int S, const SharePlan* __restrict__ plans, const uint32_t* __restrict__ eoff, const uint32_t* __restrict__ xoff,
                                uint64_t nE_total, uint32_t* __restrict__ b_loff, uint8_t* __restrict__ b_rule, uint64_t* __restrict__ b_ztag,
                                uint64_t* __restrict__ b_nlo, uint64_t* __restrict__ b_nhi, uint32_t* __restrict__ b_pa, uint32_t* __restrict__ b_pb,
                                uint32_t* __restrict__ b_lid, uint16_t* __restrict__ b_idx, uint8_t* __restrict__ b_ch, int RAW,
                                uint32_t* __restrict__ s_seed, uint16_t* __restrict__ s_idx, uint8_t* __restrict__ s_ch, uint64_t* __restrict__ s_salt,
                                uint32_t* __restrict__ s_row, uint2* __restrict__ fix_pairs) {
    uint64_t sh = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sh >= (uint64_t)S * n) return;
    uint64_t i = sh / S;
    int s = (int)(sh % S);
    const SharePlan& P = plans[sh];
    const SharePlan& Pa = plans[S * i + (S - 1)];            // the share that becomes layer 0
    uint32_t layer = (S == 1 || s == 1) ? 0u : 1u;           // combine_ciphers(enc(v+mask), enc(-mask)), ops/encrypt.hpp:284-286
    uint64_t L = (uint64_t)S * i + layer;
    b_rule[L] = 0; b_ztag[L] = P.ztag; b_nlo[L] = P.nlo; b_nhi[L] = P.nhi; b_pa[L] = 0; b_pb[L] = 0;
    if (s == 0) { b_loff[i] = (uint32_t)(S * i); if (i == n - 1) b_loff[n] = (uint32_t)(S * n); }
    uint32_t e0 = eoff[i] + (layer == 0 ? 0u : Pa.n_out);
    uint32_t x = xoff[i] + (layer == 0 ? 0u : (uint32_t)(Pa.n_raw - Pa.n_out));
    for (uint32_t r = 0; r < P.n_raw; r++) {
        uint64_t j = sh * (uint64_t)RAW + r;
        s_seed[j] = (uint32_t)L; s_idx[j] = P.idx[r]; s_ch[j] = P.ch[r]; s_salt[j] = P.salt[r];
        if (P.first[r]) {
            uint32_t row = e0 + P.pos[r];
            s_row[j] = row;
            b_lid[row] = layer; b_idx[row] = P.idx[r]; b_ch[row] = P.ch[r];
        } else {                                             // merged into an earlier edge of its slot, or its slot was dropped: scratch row
            s_row[j] = (uint32_t)(nE_total + x);
            fix_pairs[x] = make_uint2(P.pos[r] == kNoPos ? kNoFix : e0 + P.pos[r], (uint32_t)(nE_total + x));
            x++;
        }
    }
}

struct DropCand { uint32_t share, key, row, confirmed; };
constexpr uint32_t kMaxDropCand = 4096;

// thread per share: weights (ops/encrypt.hpp:184-252), merged per slot; slots with weight zero become drop candidates
__global__ void enc_weights_kernel(uint64_t n, int S, const SharePlan* __restrict__ plans, const uint32_t* __restrict__ eoff, const Fp* __restrict__ prf,
                                   const Fp* __restrict__ powg, int Z2, int Z3, Fp* __restrict__ b_w, unsigned int* __restrict__ ncand, DropCand* __restrict__ cand) {
    uint64_t sh = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sh >= (uint64_t)S * n) return;
    uint64_t i = sh / S;
    int s = (int)(sh % S);
    const SharePlan& P = plans[sh];
    const int zeros = share_weights(P, prf + sh * (uint64_t)(Z2 + Z3), powg, Z2, Z3);
    uint32_t e0 = eoff[i] + ((S == 1 || s == 1) ? 0u : plans[2 * i + 1].n_out);
    for (uint32_t p = 0; p < P.n_out; p++) b_w[e0 + p] = P.wsum[p];
    if (zeros)
        for (uint32_t r = 0; r < P.n_raw; r++)
            if (P.first[r] && fp_is_zero(P.wsum[P.pos[r]])) {
                unsigned int q = atomicAdd(ncand, 1u);
                if (q < kMaxDropCand) cand[q] = DropCand{(uint32_t)sh, (uint32_t)P.idx[r] * 2u + P.ch[r], e0 + P.pos[r], 0u};
            }
}

// one warp per candidate: is the merged syndrome of that edge zero as well? (then compact_edges removes the edge)
__global__ void enc_dropcheck_kernel(uint32_t ncand, DropCand* __restrict__ cand, const uint64_t* __restrict__ sigma) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= ncand) return;
    const uint64_t* row = sigma + (uint64_t)cand[q].row * kMWords;
    uint64_t acc = 0;
    for (int k = lane; k < kMWords; k += 32) acc |= row[k];
    const unsigned nz = __ballot_sync(0xffffffffu, acc != 0);
    if (lane == 0) cand[q].confirmed = nz ? 0u : 1u;
}

// plan_noise (ops/encrypt.hpp:16-27), B = 337: the only floating point on the path, evaluated on the host with
// the same double expressions as the reference (noise_entropy_bits 120, depth_slope_bits 16, tuple2_fraction 0.55, B 337)
void plan_noise_host(const Ctx* ctx, int depth_hint, int& z2, int& z3) {
    const double neb = ctx ? ctx->noise_entropy_bits : 120.0, slope = ctx ? ctx->depth_slope_bits : 16.0, t2 = ctx ? ctx->tuple2_fraction : 0.55;
    double budget = neb + slope * (double)std::max(0, depth_hint);
    double per2 = 2.0 * std::log2((double)kB), per3 = 3.0 * std::log2((double)kB);
    z2 = std::max(0, (int)std::floor((budget * t2) / std::max(1e-6, per2)));
    z3 = std::max(0, (int)std::floor((budget * (1.0 - t2)) / std::max(1e-6, per3)));
    if (z2 + z3 == 1) { z3 > 0 ? ++z3 : ++z2; }
}

// The tape of a call as the kernels see it: kind and key of the context, per-item stream ids / first words / global item numbers
// uploaded from the host arrays the caller gave (any of them may be null).
int tape_spec(Ctx* ctx, Scratch& scratch, uint64_t n, uint64_t batch_seed, const uint64_t* h_states, const uint64_t* h_k0, const uint64_t* h_ids, TapeSpec& ts) {
    int rc;
    ts = TapeSpec();
    ts.kind = ctx->tape_kind;
    for (int i = 0; i < 8; i++) ts.key[i] = ctx->tape_key[i];
    ts.batch_seed = batch_seed;
    ts.item_base = ctx->item_base;
    auto up = [&](const uint64_t* h, const uint64_t*& d) -> int {
        if (!h || !n) return PV_OK;
        uint64_t* p = nullptr;
        if ((rc = scratch.alloc(p, n * 8))) return rc;
        PV_CUDA(cudaMemcpyAsync(p, h, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        d = p;
        return PV_OK;
    };
    if ((rc = up(h_states, ts.states)) || (rc = up(h_k0, ts.k0)) || (rc = up(h_ids, ts.ids))) return rc;
    if (ts.kind == TAPE_WORDS) {
        uint64_t need = 0;
        if (h_ids) { for (uint64_t i = 0; i < n; i++) need = std::max(need, h_ids[i] + 1); }
        else need = ctx->item_base + n;
        if (!ctx->d_tape_words || need > ctx->tape_words_items) {
            ctx->last_error = "tape kind WORDS: pvacb_set_tape_words has not supplied words for every item of this call";
            return PV_E_ARG;
        }
        ts.words = ctx->d_tape_words;
        ts.words_per_item = ctx->tape_words_per_item;
    }
    return PV_OK;
}

struct EncDrop { uint64_t item; uint32_t share, key; };

// one pass of the pipeline over n items. drops: edges that compact_edges would have removed (confirmed on the device), by item.
static int enc_pass(Ctx* ctx, const uint64_t* values, bool on_device, uint64_t n, uint64_t batch_seed, const uint64_t* h_states, const uint64_t* h_k0,
                    const uint64_t* h_ids, const uint64_t* h_drop, int depth_hint, int shares, uint64_t* h_next, Batch** out, std::vector<EncDrop>& drops) {
    const uint64_t S = (uint64_t)shares;
    Scratch scratch(ctx);
    int Z2, Z3;
    plan_noise_host(ctx, depth_hint, Z2, Z3);
    const int G = Z2 + Z3, RAW = plan_raw_edges(Z2, Z3), RND = plan_rnd_values(Z2, Z3);
    if (G < 1) { ctx->last_error = "enc_value: plan_noise gave no noise group"; return PV_E_ARG; }
    int rc;
    uint64_t *d_vals = nullptr, *d_drop = nullptr;
    SharePlan* plans = nullptr;
    uint8_t* slabs = nullptr;
    uint32_t *cnt = nullptr, *xcnt = nullptr, *eoff = nullptr, *xoff = nullptr;
    uint64_t *j_ztag = nullptr, *j_nlo = nullptr, *j_nhi = nullptr;
    uint8_t* j_flags = nullptr;
    Fp* prf = nullptr;
    unsigned int* err = nullptr;          // [0] error bits, [1] drop candidates
    DropCand* cand = nullptr;
    const uint64_t njobs = S * n * G;
    TapeSpec ts;
    if ((rc = tape_spec(ctx, scratch, n, batch_seed, h_states, h_k0, h_ids, ts))) return rc;
    if (!on_device) {
        const size_t vbytes = n * (shares == 2 ? 8 : 16);
        if ((rc = scratch.alloc(d_vals, vbytes))) return rc;
        PV_CUDA(cudaMemcpyAsync(d_vals, values, vbytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (h_drop) {
        if ((rc = scratch.alloc(d_drop, S * n * kKeyWords * 8))) return rc;
        PV_CUDA(cudaMemcpyAsync(d_drop, h_drop, S * n * kKeyWords * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = scratch.alloc(plans, S * n * sizeof(SharePlan)))) return rc;
    if ((rc = scratch.alloc(slabs, S * n * plan_slab_bytes(RAW, RND)))) return rc;
    if ((rc = scratch.alloc(cnt, n * 4))) return rc;
    if ((rc = scratch.alloc(xcnt, n * 4))) return rc;
    if ((rc = scratch.alloc(eoff, (n + 1) * 4))) return rc;
    if ((rc = scratch.alloc(xoff, (n + 1) * 4))) return rc;
    if ((rc = scratch.alloc(j_ztag, njobs * 8))) return rc;
    if ((rc = scratch.alloc(j_nlo, njobs * 8))) return rc;
    if ((rc = scratch.alloc(j_nhi, njobs * 8))) return rc;
    if ((rc = scratch.alloc(j_flags, njobs))) return rc;
    if ((rc = scratch.alloc(prf, njobs * 16))) return rc;
    if ((rc = scratch.alloc(err, 8))) return rc;
    if ((rc = scratch.alloc(cand, kMaxDropCand * sizeof(DropCand)))) return rc;
    PV_CUDA(cudaMemsetAsync(err, 0, 8, ctx->stream));
    uint64_t* d_next = nullptr;
    if (h_next && (rc = scratch.alloc(d_next, n * 8))) return rc;
    enc_plan_kernel<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(n, on_device ? values : d_vals, ts, ctx->kv.canon_tag, Z2, Z3, shares, d_drop, plans, slabs, cnt, xcnt,
                                                                        j_ztag, j_nlo, j_nhi, j_flags, d_next, err);
    if (h_next) PV_CUDA(cudaMemcpyAsync(h_next, d_next, n * 8, cudaMemcpyDeviceToHost, ctx->stream));   // complete at the next stream sync below
    if ((rc = scan_u32(ctx, n, cnt, eoff))) return rc;
    if ((rc = scan_u32(ctx, n, xcnt, xoff))) return rc;
    uint32_t tot[2] = {0, 0};
    unsigned int h_err[2] = {0, 0};
    { SmallRead sr; sr.add(&tot[0], eoff + n, 4); sr.add(&tot[1], xoff + n, 4); sr.add(h_err, err, 4); if ((rc = read_small_sync(ctx, sr))) return rc; }
    ctx->stat_kernel_launches += 1;
    if (h_err[0] & 2u) { ctx->last_error = "enc_value: the supplied tape words (pvacb_set_tape_words) ran out"; return PV_E_ARG; }
    const uint64_t nE = tot[0], nX = tot[1];

    Batch* b = nullptr;
    if ((rc = batch_alloc(ctx, n, S * n, nE, &b))) return rc;
    uint64_t nS = S * n * (uint64_t)RAW;
    uint32_t *s_seed = nullptr, *s_row = nullptr;
    uint16_t* s_idx = nullptr;
    uint8_t* s_ch = nullptr;
    uint64_t *s_salt = nullptr, *tmp_rows = nullptr;
    uint2* fix = nullptr;
    if ((rc = scratch.alloc(s_seed, nS * 4))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(s_row, nS * 4))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(s_idx, nS * 2))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(s_ch, nS))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(s_salt, nS * 8))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(fix, (nX ? nX : 1) * sizeof(uint2)))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(tmp_rows, (nX ? nX : 1) * (size_t)kMWords * 8))) { batch_free(b); return rc; }
    PV_CUDA(cudaMemcpyAsync(b->eoff, eoff, (n + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    enc_emit_kernel<<<(unsigned)((S * n + 127) / 128), 128, 0, ctx->stream>>>(n, shares, plans, eoff, xoff, nE, b->loff, b->rule, b->ztag, b->nlo, b->nhi, b->pa, b->pb,
                                                                               b->lid, b->idx, b->ch, RAW, s_seed, s_idx, s_ch, s_salt, s_row, fix);
    ctx->stat_kernel_launches += 1;
    // PRF (R of each share, noise deltas)
    if ((rc = prf_run(ctx, njobs, j_ztag, j_nlo, j_nhi, j_flags, prf, nullptr))) { batch_free(b); return rc; }
    enc_weights_kernel<<<(unsigned)((S * n + 63) / 64), 64, 0, ctx->stream>>>(n, shares, plans, eoff, prf, ctx->kv.powg, Z2, Z3, b->w, err + 1, cand);
    ctx->stat_kernel_launches += 1;
    // sigma of every raw edge
    SigmaJobs J;
    J.n = nS; J.ztag = b->ztag; J.nlo = b->nlo; J.nhi = b->nhi; J.seed_idx = s_seed; J.idx = s_idx; J.ch = s_ch; J.salt = s_salt;
    J.out_row = s_row; J.out = b->sigma; J.out_split = nE; J.out2 = tmp_rows;
    if ((rc = sigma_run(ctx, J))) { batch_free(b); return rc; }
    if ((rc = sigma_xor_rows(ctx, nX, fix, b->sigma, nE, tmp_rows))) { batch_free(b); return rc; }
    { SmallRead sr; sr.add(h_err, err, 8); if ((rc = read_small_sync(ctx, sr))) { batch_free(b); return rc; } }
    if (h_err[1]) {
        if (h_err[1] > kMaxDropCand) { batch_free(b); ctx->last_error = "enc_value: more zero-weight edges than the drop path tracks"; return PV_E_SHAPE; }
        enc_dropcheck_kernel<<<(h_err[1] * 32 + 255) / 256, 256, 0, ctx->stream>>>(h_err[1], cand, b->sigma);
        ctx->stat_kernel_launches += 1;
        std::vector<DropCand> hc(h_err[1]);
        PV_CUDA(cudaMemcpyAsync(hc.data(), cand, hc.size() * sizeof(DropCand), cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA(cudaStreamSynchronize(ctx->stream));
        for (const DropCand& c : hc)
            if (c.confirmed) drops.push_back(EncDrop{c.share / S, (uint32_t)(c.share % S), c.key});
    }
    *out = b;
    return PV_OK;
}

// shares = 2: enc_value_depth (values: n u64 plaintexts); shares = 1: enc_fp_depth (values: n x (lo, hi) field elements)
int op_enc_value(Ctx* ctx, const uint64_t* values, bool on_device, uint64_t n, uint64_t batch_seed, const uint64_t* h_states, Batch** out, int depth_hint, int shares,
                 uint64_t* h_next, const uint64_t* h_k0, const uint64_t* h_ids) {
    if (shares != 1 && shares != 2) return PV_E_ARG;
    if (!ctx->have_sk) { ctx->last_error = "this context holds a public key only"; return PV_E_NOKEYS; }
    if (n == 0) return batch_alloc(ctx, 0, 0, 0, out);
    const uint64_t S = (uint64_t)shares;
    std::vector<EncDrop> drops;
    Batch* b = nullptr;
    int rc = enc_pass(ctx, values, on_device, n, batch_seed, h_states, h_k0, h_ids, nullptr, depth_hint, shares, h_next, &b, drops);
    if (rc || drops.empty()) { *out = b; return rc; }
    // ---- flagged slow path: compact_edges removed an edge (ops/encrypt.hpp:59). Plan the affected items again with those slots
    // dropped, until a pass confirms no further removal; then splice them into the batch.
    std::vector<uint64_t> h_vals(n * (shares == 2 ? 1 : 2));
    if (on_device) {
        PV_CUDA(cudaMemcpyAsync(h_vals.data(), values, h_vals.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA(cudaStreamSynchronize(ctx->stream));
    } else memcpy(h_vals.data(), values, h_vals.size() * 8);
    std::map<uint64_t, std::vector<uint64_t>> masks;      // item -> S x kKeyWords
    Batch* sub = nullptr;
    std::vector<uint64_t> items;
    for (int round = 0; round < 16 && !drops.empty(); round++) {
        for (const EncDrop& d : drops) {
            auto& m = masks[d.item];
            if (m.empty()) m.assign(S * kKeyWords, 0);
            if (S == 2 && d.share == 0)                    // share 0 is drawn first: removing one of its slots shifts every word of share 1
                for (int w = 0; w < kKeyWords; w++) m[kKeyWords + w] = 0;
            m[d.share * kKeyWords + (d.key >> 6)] |= 1ull << (d.key & 63);
        }
        drops.clear();
        if (sub) { batch_free(sub); sub = nullptr; }
        items.clear();
        std::vector<uint64_t> v, st, k0, ids, dm, nx;
        for (auto& kv : masks) {
            const uint64_t i = kv.first;
            items.push_back(i);
            if (shares == 2) v.push_back(h_vals[i]); else { v.push_back(h_vals[2 * i]); v.push_back(h_vals[2 * i + 1]); }
            if (h_states) st.push_back(h_states[i]);
            if (h_k0) k0.push_back(h_k0[i]);
            ids.push_back(h_ids ? h_ids[i] : ctx->item_base + i);
            dm.insert(dm.end(), kv.second.begin(), kv.second.end());
        }
        nx.assign(items.size(), 0);
        std::vector<EncDrop> d2;                           // the sub-batch names its items by global number (ids)
        rc = enc_pass(ctx, v.data(), false, items.size(), batch_seed, h_states ? st.data() : nullptr, h_k0 ? k0.data() : nullptr, ids.data(), dm.data(), depth_hint,
                      shares, nx.data(), &sub, d2);
        if (rc) { batch_free(b); return rc; }
        if (h_next) for (size_t q = 0; q < items.size(); q++) h_next[items[q]] = nx[q];
        for (EncDrop d : d2) { d.item = items[d.item]; drops.push_back(d); }
    }
    if (!drops.empty()) { batch_free(b); batch_free(sub); ctx->last_error = "enc_value: compact_edges drop path did not settle"; return PV_E_SHAPE; }
    std::vector<uint32_t> which(n, 0), index(n);
    for (uint64_t i = 0; i < n; i++) index[i] = (uint32_t)i;
    for (size_t q = 0; q < items.size(); q++) { which[items[q]] = 1; index[items[q]] = (uint32_t)q; }
    const Batch* srcs[2] = {b, sub};
    Batch* merged = nullptr;
    rc = batch_select(ctx, srcs, 2, which, index, &merged);
    batch_free(b); batch_free(sub);
    if (rc) return rc;
    rc = compact_layers_batch(ctx, merged);              // a share that lost every edge loses its layer (combine_ciphers -> compact_layers)
    if (rc) { batch_free(merged); return rc; }
    *out = merged;
    return PV_OK;
}

}  // namespace pvacb
