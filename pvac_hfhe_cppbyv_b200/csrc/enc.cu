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
#include "engine.h"
#include "enc_plan.cuh"

#include <algorithm>
#include <cmath>

namespace pvacb {

// share s = 0 is enc_fp_depth(-mask) (drawn first, becomes layer 1 / the trailing edges), s = 1 is enc_fp_depth(v+mask)
__global__ void enc_plan_kernel(uint64_t n, const uint64_t* __restrict__ values, uint64_t batch_seed, const uint64_t* __restrict__ states,
                                uint64_t canon_tag, int Z2, int Z3, int S,
                                SharePlan* __restrict__ plans, uint32_t* __restrict__ n_edges, uint32_t* __restrict__ n_extra,
                                uint64_t* __restrict__ j_ztag, uint64_t* __restrict__ j_nlo, uint64_t* __restrict__ j_nhi, uint8_t* __restrict__ j_flags,
                                uint64_t* __restrict__ draws /* optional: tape words consumed per item */) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t s0 = states ? states[i] : item_stream_state(batch_seed, i);
    uint64_t used;
    if (S == 2) used = plan_item(s0, values[i], canon_tag, Z2, Z3, plans[2 * i], plans[2 * i + 1]);
    else used = plan_single(s0, fp_make(values[2 * i], values[2 * i + 1]), canon_tag, Z2, Z3, plans[i]);     // enc_fp_depth: values are Fp (lo, hi)
    if (draws) draws[i] = used;
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
    for (int r = 0; r < P.n_raw; r++) {
        uint64_t j = sh * (uint64_t)RAW + r;
        uint32_t row = e0 + P.pos[r];
        s_seed[j] = (uint32_t)L; s_idx[j] = P.idx[r]; s_ch[j] = P.ch[r]; s_salt[j] = P.salt[r];
        if (P.first[r]) {
            s_row[j] = row;
            b_lid[row] = layer; b_idx[row] = P.idx[r]; b_ch[row] = P.ch[r];
        } else {
            s_row[j] = (uint32_t)(nE_total + x);
            fix_pairs[x] = make_uint2(row, (uint32_t)(nE_total + x));
            x++;
        }
    }
}

// thread per share: weights (ops/encrypt.hpp:184-252), merged per slot
__global__ void enc_weights_kernel(uint64_t n, int S, const SharePlan* __restrict__ plans, const uint32_t* __restrict__ eoff, const Fp* __restrict__ prf,
                                   const Fp* __restrict__ powg, int Z2, int Z3, Fp* __restrict__ b_w, unsigned int* __restrict__ err) {
    uint64_t sh = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sh >= (uint64_t)S * n) return;
    uint64_t i = sh / S;
    int s = (int)(sh % S);
    const SharePlan& P = plans[sh];
    Fp wsum[kMaxRaw];
    if (!share_weights(P, prf + sh * (uint64_t)(Z2 + Z3), powg, Z2, Z3, wsum)) atomicOr(err, 1u);
    uint32_t e0 = eoff[i] + ((S == 1 || s == 1) ? 0u : plans[2 * i + 1].n_out);
    for (int p = 0; p < P.n_out; p++) b_w[e0 + p] = wsum[p];
}

// plan_noise (ops/encrypt.hpp:16-27) for the default Params: the only floating point on the path, evaluated on the host with
// the same double expressions as the reference (noise_entropy_bits 120, depth_slope_bits 16, tuple2_fraction 0.55, B 337)
void plan_noise_host(int depth_hint, int& z2, int& z3) {
    double budget = 120.0 + 16.0 * (double)std::max(0, depth_hint);
    double per2 = 2.0 * std::log2((double)kB), per3 = 3.0 * std::log2((double)kB);
    z2 = std::max(0, (int)std::floor((budget * 0.55) / std::max(1e-6, per2)));
    z3 = std::max(0, (int)std::floor((budget * (1.0 - 0.55)) / std::max(1e-6, per3)));
    if (z2 + z3 == 1) { z3 > 0 ? ++z3 : ++z2; }
}

// shares = 2: enc_value_depth (values: n u64 plaintexts); shares = 1: enc_fp_depth (values: n x (lo, hi) field elements)
int op_enc_value(Ctx* ctx, const uint64_t* values, bool on_device, uint64_t n, uint64_t batch_seed, const uint64_t* h_states, Batch** out, int depth_hint, int shares, uint64_t* h_draws) {
    const uint64_t S = (uint64_t)shares;
    if (shares != 1 && shares != 2) return PV_E_ARG;
    Scratch scratch(ctx);
    int Z2, Z3;
    plan_noise_host(depth_hint, Z2, Z3);
    const int G = Z2 + Z3, RAW = kSignal + 2 * Z2 + 3 * Z3;
    if (Z2 > kMaxZ2 || Z3 > kMaxZ3 || G < 1) { ctx->last_error = "enc_value: depth_hint outside the supported range 0..23"; return PV_E_ARG; }
    int rc;
    if (n == 0) return batch_alloc(ctx, 0, 0, 0, out);
    uint64_t *d_vals = nullptr, *d_states = nullptr;
    SharePlan* plans = nullptr;
    uint32_t *cnt = nullptr, *xcnt = nullptr, *eoff = nullptr, *xoff = nullptr;
    uint64_t *j_ztag = nullptr, *j_nlo = nullptr, *j_nhi = nullptr;
    uint8_t* j_flags = nullptr;
    Fp* prf = nullptr;
    const uint64_t njobs = S * n * G;
    if (!on_device) {
        const size_t vbytes = n * (shares == 2 ? 8 : 16);
        if ((rc = scratch.alloc(d_vals, vbytes))) return rc;
        PV_CUDA(cudaMemcpyAsync(d_vals, values, vbytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (h_states) {
        if ((rc = scratch.alloc(d_states, n * 8))) return rc;
        PV_CUDA(cudaMemcpyAsync(d_states, h_states, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = scratch.alloc(plans, S * n * sizeof(SharePlan)))) return rc;
    if ((rc = scratch.alloc(cnt, n * 4))) return rc;
    if ((rc = scratch.alloc(xcnt, n * 4))) return rc;
    if ((rc = scratch.alloc(eoff, (n + 1) * 4))) return rc;
    if ((rc = scratch.alloc(xoff, (n + 1) * 4))) return rc;
    if ((rc = scratch.alloc(j_ztag, njobs * 8))) return rc;
    if ((rc = scratch.alloc(j_nlo, njobs * 8))) return rc;
    if ((rc = scratch.alloc(j_nhi, njobs * 8))) return rc;
    if ((rc = scratch.alloc(j_flags, njobs))) return rc;
    if ((rc = scratch.alloc(prf, njobs * 16))) return rc;
    uint64_t* d_draws = nullptr;
    if (h_draws && (rc = scratch.alloc(d_draws, n * 8))) return rc;
    enc_plan_kernel<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(n, on_device ? values : d_vals, batch_seed, d_states, ctx->kv.canon_tag, Z2, Z3, shares, plans, cnt, xcnt,
                                                                        j_ztag, j_nlo, j_nhi, j_flags, d_draws);
    if (h_draws) PV_CUDA(cudaMemcpyAsync(h_draws, d_draws, n * 8, cudaMemcpyDeviceToHost, ctx->stream));   // complete at the next stream sync below
    if ((rc = scan_u32(ctx, n, cnt, eoff))) return rc;
    if ((rc = scan_u32(ctx, n, xcnt, xoff))) return rc;
    uint32_t tot[2] = {0, 0};
    { SmallRead sr; sr.add(&tot[0], eoff + n, 4); sr.add(&tot[1], xoff + n, 4); if ((rc = read_small_sync(ctx, sr))) return rc; }
    ctx->stat_kernel_launches += 1;
    const uint64_t nE = tot[0], nX = tot[1];

    Batch* b = nullptr;
    if ((rc = batch_alloc(ctx, n, S * n, nE, &b))) return rc;
    uint64_t nS = S * n * (uint64_t)RAW;
    uint32_t *s_seed = nullptr, *s_row = nullptr;
    uint16_t* s_idx = nullptr;
    uint8_t* s_ch = nullptr;
    uint64_t *s_salt = nullptr, *tmp_rows = nullptr;
    uint2* fix = nullptr;
    unsigned int* err = nullptr;
    if ((rc = scratch.alloc(s_seed, nS * 4))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(s_row, nS * 4))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(s_idx, nS * 2))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(s_ch, nS))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(s_salt, nS * 8))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(fix, (nX ? nX : 1) * sizeof(uint2)))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(tmp_rows, (nX ? nX : 1) * (size_t)kMWords * 8))) { batch_free(b); return rc; }
    if ((rc = scratch.alloc(err, 4))) { batch_free(b); return rc; }
    PV_CUDA(cudaMemsetAsync(err, 0, 4, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(b->eoff, eoff, (n + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    enc_emit_kernel<<<(unsigned)((S * n + 127) / 128), 128, 0, ctx->stream>>>(n, shares, plans, eoff, xoff, nE, b->loff, b->rule, b->ztag, b->nlo, b->nhi, b->pa, b->pb,
                                                                               b->lid, b->idx, b->ch, RAW, s_seed, s_idx, s_ch, s_salt, s_row, fix);
    ctx->stat_kernel_launches += 1;
    // PRF (R of each share, noise deltas)
    if ((rc = prf_run(ctx, njobs, j_ztag, j_nlo, j_nhi, j_flags, prf, nullptr))) { batch_free(b); return rc; }
    enc_weights_kernel<<<(unsigned)((S * n + 63) / 64), 64, 0, ctx->stream>>>(n, shares, plans, eoff, prf, ctx->kv.powg, Z2, Z3, b->w, err);
    ctx->stat_kernel_launches += 1;
    // sigma of every raw edge
    SigmaJobs J;
    J.n = nS; J.ztag = b->ztag; J.nlo = b->nlo; J.nhi = b->nhi; J.seed_idx = s_seed; J.idx = s_idx; J.ch = s_ch; J.salt = s_salt;
    J.out_row = s_row; J.out = b->sigma; J.out_split = nE; J.out2 = tmp_rows;
    if ((rc = sigma_run(ctx, J))) { batch_free(b); return rc; }
    if ((rc = sigma_xor_rows(ctx, nX, fix, b->sigma, nE, tmp_rows))) { batch_free(b); return rc; }
    unsigned int h_err = 0;
    { SmallRead sr; sr.add(&h_err, err, 4); if ((rc = read_small_sync(ctx, sr))) { batch_free(b); return rc; } }
    if (h_err) {
        batch_free(b);
        ctx->last_error = "enc_value: a merged edge weight is zero (compact_edges drop branch, p ~ 2^-127)";
        return PV_E_RARE_PATH;
    }
    *out = b;
    return PV_OK;
}

}  // namespace pvacb
