// dec_value (ops/decrypt.hpp:12-89) for a batch:
//   1. prf_R of every DISTINCT BASE-layer seed (prf.cu; the expensive part) -- equal seeds share one evaluation,
//   2. PROD layers: R = R[pa]*R[pb] resolved per ciphertext (any DAG order; cycles / bad parents are reported where the
//      reference aborts),
//   3. one Fermat inversion per layer,
//   4. signed edge sum  acc = sum_e +- w_e * g^idx_e * Rinv[layer_e]: reads 23 B per edge (w, layer_id, idx, ch), never
//      sigma -- in the structure-of-arrays layout this stage is a pure HBM stream.
#include "engine.h"

#include <algorithm>

namespace pvacb {

// thread per ciphertext; R[] holds prf_R for BASE layers and 0 elsewhere on entry
__global__ void dec_resolve_kernel(uint64_t n, const uint32_t* __restrict__ loff, const uint8_t* __restrict__ rule, const uint32_t* __restrict__ pa,
                                   const uint32_t* __restrict__ pb, Fp* __restrict__ R, unsigned int* __restrict__ err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t l0 = loff[i], L = loff[i + 1] - l0;
    uint32_t pending = 0;
    for (uint32_t l = 0; l < L; l++)
        if (rule[l0 + l] == 1) {
            if (pa[l0 + l] >= L || pb[l0 + l] >= L) { atomicOr(err, 1u); return; }   // ops/decrypt.hpp:21-24
            pending++;
        }
    while (pending) {
        uint32_t progressed = 0;
        for (uint32_t l = 0; l < L; l++) {
            if (rule[l0 + l] != 1 || !fp_is_zero(R[l0 + l])) continue;
            Fp a = R[l0 + pa[l0 + l]], b = R[l0 + pb[l0 + l]];
            if (fp_is_zero(a) || fp_is_zero(b)) continue;
            R[l0 + l] = fp_mul(a, b);
            progressed++;
        }
        if (!progressed) { atomicOr(err, 2u); return; }   // cycle, ops/decrypt.hpp:31-37
        pending -= progressed;
    }
}

__global__ void dec_inv_kernel(uint64_t nL, const Fp* __restrict__ R, Fp* __restrict__ Rinv) {
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nL) Rinv[l] = fp_inv(R[l]);
}

__device__ __forceinline__ Fp warp_sum_fp(Fp v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Fp t;
        t.lo = __shfl_down_sync(0xffffffffu, v.lo, o);
        t.hi = __shfl_down_sync(0xffffffffu, v.hi, o);
        v = fp_add(v, t);
    }
    return v;
}

// Edge stage: acc = sum_e +- w_e g^idx_e Rinv[layer_e]. Integer-pipe bound, not HBM bound (ncu r02: 23 B per edge against two
// 128-bit modular products): so the work per edge is ONE reduced product (w g^idx, then its sign) and one UNREDUCED 256-bit
// multiply-accumulate with Rinv into a 320-bit accumulator per thread (fp_mac_wide), reduced once at the end.
// kBlockPerItem == false: one warp per ciphertext (small ciphertexts). kBlockPerItem == true: grid (ciphertext, slice); a CTA walks
// chunks slice, slice + gridDim.y, ... of 256 x kDecEdgesPerThread edges of its ciphertext and leaves one partial sum per (ciphertext, slice),
// so that a few very large ciphertexts (depth-3 products: 172 544 edges) still fill the machine; dec_partials_kernel adds them.
constexpr int kDecChunk = 256 * 8;
template <bool kBlockPerItem>
__global__ void __launch_bounds__(256)
dec_edges_kernel(uint64_t n, const uint32_t* __restrict__ loff, const uint32_t* __restrict__ eoff, const uint32_t* __restrict__ lid,
                 const uint16_t* __restrict__ idx, const uint8_t* __restrict__ ch, const Fp* __restrict__ w, const Fp* __restrict__ powg,
                 const Fp* __restrict__ Rinv, Fp* __restrict__ out) {
    __shared__ Fp s_g[kB];
    __shared__ Fp s_part[8];
    for (int k = threadIdx.x; k < kB; k += blockDim.x) s_g[k] = powg[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t item = kBlockPerItem ? (uint64_t)blockIdx.x : (uint64_t)blockIdx.x * 8 + wid;
    uint64_t wide[5] = {0, 0, 0, 0, 0};
    if (item < n) {
        const uint32_t l0 = loff[item], e0 = eoff[item], E = eoff[item + 1] - e0;
        auto edge = [&](uint32_t e) {
            Fp t = fp_mul(w[e0 + e], s_g[idx[e0 + e]]);
            if (ch[e0 + e]) t = fp_neg(t);
            fp_mac_wide(wide, t, Rinv[l0 + lid[e0 + e]]);
        };
        if (kBlockPerItem) {
            for (uint32_t c = blockIdx.y * kDecChunk; c < E; c += gridDim.y * kDecChunk)
                for (uint32_t e = c + threadIdx.x; e < min(E, c + (uint32_t)kDecChunk); e += 256) edge(e);
        } else {
            for (uint32_t e = lane; e < E; e += 32) edge(e);
        }
    }
    Fp acc = warp_sum_fp(fp_wide_reduce(wide));
    if (!kBlockPerItem) {
        if (lane == 0 && item < n) out[item] = acc;
    } else {
        if (lane == 0) s_part[wid] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            Fp t = s_part[0];
            for (int k = 1; k < 8; k++) t = fp_add(t, s_part[k]);
            out[item * gridDim.y + blockIdx.y] = t;
        }
    }
}
__global__ void dec_partials_kernel(uint64_t n, uint32_t slices, const Fp* __restrict__ part, Fp* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp t = part[i * slices];
    for (uint32_t k = 1; k < slices; k++) t = fp_add(t, part[i * slices + k]);
    out[i] = t;
}

// ---- BASE layers with equal seeds share one PRF evaluation. The reference memoises by layer id, so c*c (every seed twice),
// the depth chains (step 3: 320 layers, 2 distinct seeds) or a sum that reuses an operand recompute prf_R for every copy
// (ops/decrypt.hpp:12-60); R depends on the seed only, so evaluating it once per distinct seed gives the same values.
// Open-addressing table keyed by a 64-bit hash of the seed; the representative is the smallest layer index with that hash and
// a layer only aliases it after comparing the full seed (a hash collision just means "evaluate your own").
__device__ __forceinline__ uint64_t seed_hash(uint64_t z, uint64_t lo, uint64_t hi) {
    uint64_t h = mix64(hi + 0x9E3779B97F4A7C15ull);
    h = mix64(lo ^ h);
    h = mix64(z ^ h);
    return h | 1ull;            // 0 marks an empty slot
}
__global__ void dec_seed_insert_kernel(uint64_t nL, const uint8_t* __restrict__ rule, const uint64_t* __restrict__ ztag, const uint64_t* __restrict__ nlo,
                                       const uint64_t* __restrict__ nhi, unsigned long long* __restrict__ tkey, uint32_t* __restrict__ trep, uint64_t mask) {
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nL || rule[l] != 0) return;
    const unsigned long long h = seed_hash(ztag[l], nlo[l], nhi[l]);
    for (uint64_t slot = h & mask;; slot = (slot + 1) & mask) {
        unsigned long long old = atomicCAS(&tkey[slot], 0ull, h);
        if (old == 0ull || old == h) { atomicMin(&trep[slot], (uint32_t)l); return; }
    }
}
__global__ void dec_seed_lookup_kernel(uint64_t nL, const uint8_t* __restrict__ rule, const uint64_t* __restrict__ ztag, const uint64_t* __restrict__ nlo,
                                       const uint64_t* __restrict__ nhi, const unsigned long long* __restrict__ tkey, const uint32_t* __restrict__ trep,
                                       uint64_t mask, uint8_t* __restrict__ flags, uint32_t* __restrict__ alias) {
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nL) return;
    alias[l] = (uint32_t)l;
    if (rule[l] != 0) { flags[l] = 0; return; }
    const unsigned long long h = seed_hash(ztag[l], nlo[l], nhi[l]);
    uint64_t slot = h & mask;
    while (tkey[slot] != h) slot = (slot + 1) & mask;
    const uint32_t rep = trep[slot];
    if (rep != (uint32_t)l && ztag[rep] == ztag[l] && nlo[rep] == nlo[l] && nhi[rep] == nhi[l]) { flags[l] = 0; alias[l] = rep; }
    else flags[l] = 2;          // active, family 0 (prf_R)
}
__global__ void dec_alias_copy_kernel(uint64_t nL, const uint32_t* __restrict__ alias, Fp* __restrict__ R) {
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nL) return;
    const uint32_t a = alias[l];
    if (a != (uint32_t)l) R[l] = R[a];      // representatives are never aliased themselves: no ordering hazard
}

int op_dec_value(Ctx* ctx, const Batch* Cb, uint64_t* h_out) {
    if (Cb->n == 0) return PV_OK;
    if (!ctx->have_sk) { ctx->last_error = "this context holds a public key only"; return PV_E_NOKEYS; }
    Scratch scratch(ctx);
    int rc;
    uint8_t* flags = nullptr;
    Fp *R = nullptr, *Rinv = nullptr, *res = nullptr;
    unsigned int* err = nullptr;
    uint64_t nLa = Cb->nL ? Cb->nL : 1;
    if ((rc = scratch.alloc(flags, nLa))) return rc;
    if ((rc = scratch.alloc(R, nLa * 16))) return rc;
    if ((rc = scratch.alloc(Rinv, nLa * 16))) return rc;
    if ((rc = scratch.alloc(res, Cb->n * 16))) return rc;
    if ((rc = scratch.alloc(err, 4))) return rc;
    PV_CUDA(cudaMemsetAsync(err, 0, 4, ctx->stream));
    if (Cb->nL) {
        uint64_t tsize = 1024;
        while (tsize < 2 * Cb->nL) tsize <<= 1;
        unsigned long long* tkey = nullptr;
        uint32_t *trep = nullptr, *alias = nullptr;
        if ((rc = scratch.alloc(tkey, tsize * 8)) || (rc = scratch.alloc(trep, tsize * 4)) || (rc = scratch.alloc(alias, nLa * 4))) return rc;
        PV_CUDA(cudaMemsetAsync(tkey, 0, tsize * 8, ctx->stream));
        PV_CUDA(cudaMemsetAsync(trep, 0xFF, tsize * 4, ctx->stream));
        const unsigned lb = (unsigned)((Cb->nL + 255) / 256);
        dec_seed_insert_kernel<<<lb, 256, 0, ctx->stream>>>(Cb->nL, Cb->rule, Cb->ztag, Cb->nlo, Cb->nhi, tkey, trep, tsize - 1);
        dec_seed_lookup_kernel<<<lb, 256, 0, ctx->stream>>>(Cb->nL, Cb->rule, Cb->ztag, Cb->nlo, Cb->nhi, tkey, trep, tsize - 1, flags, alias);
        ctx->stat_kernel_launches += 2;
        if ((rc = prf_run(ctx, Cb->nL, Cb->ztag, Cb->nlo, Cb->nhi, flags, R, nullptr))) return rc;
        dec_alias_copy_kernel<<<lb, 256, 0, ctx->stream>>>(Cb->nL, alias, R);
        ctx->stat_kernel_launches += 1;
        dec_resolve_kernel<<<(unsigned)((Cb->n + 127) / 128), 128, 0, ctx->stream>>>(Cb->n, Cb->loff, Cb->rule, Cb->pa, Cb->pb, R, err);
        dec_inv_kernel<<<(unsigned)((Cb->nL + 127) / 128), 128, 0, ctx->stream>>>(Cb->nL, R, Rinv);
        ctx->stat_kernel_launches += 2;
    }
    const bool block_per_item = Cb->nE / Cb->n > 256;
    uint32_t slices = 1;
    Fp* part = res;
    if (block_per_item) {
        // enough CTAs for the whole machine even when the batch holds a handful of huge ciphertexts; a ciphertext with more chunks than
        // slices is walked in a loop
        const uint64_t avg_chunks = (Cb->nE / Cb->n + kDecChunk - 1) / kDecChunk;
        const uint64_t want = ((uint64_t)ctx->sm_count * 8 + Cb->n - 1) / Cb->n;
        slices = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(avg_chunks, want), 1024));
        if (slices > 1 && (rc = scratch.alloc(part, Cb->n * (uint64_t)slices * 16))) return rc;
    }
    {
    ProfScope ps(ctx, PROF_DEC_EDGES);
    if (block_per_item)
        dec_edges_kernel<true><<<dim3((unsigned)Cb->n, slices), 256, 0, ctx->stream>>>(Cb->n, Cb->loff, Cb->eoff, Cb->lid, Cb->idx, Cb->ch, Cb->w, ctx->kv.powg, Rinv, part);
    else
        dec_edges_kernel<false><<<(unsigned)((Cb->n + 7) / 8), 256, 0, ctx->stream>>>(Cb->n, Cb->loff, Cb->eoff, Cb->lid, Cb->idx, Cb->ch, Cb->w, ctx->kv.powg, Rinv, res);
    }
    if (slices > 1) {
        dec_partials_kernel<<<(unsigned)((Cb->n + 127) / 128), 128, 0, ctx->stream>>>(Cb->n, slices, part, res);
        ctx->stat_kernel_launches += 1;
    }
    ctx->stat_kernel_launches += 1;
    PV_CUDA(cudaGetLastError());
    unsigned int h_err = 0;
    PV_CUDA(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(h_out, res, Cb->n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_err) {
        ctx->last_error = (h_err & 1) ? "dec_value: layer parent out of range" : "dec_value: cycle in the layer graph";
        return PV_E_LAYER_GRAPH;
    }
    return PV_OK;
}

}  // namespace pvacb
