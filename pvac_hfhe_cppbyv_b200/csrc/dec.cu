// dec_value (ops/decrypt.hpp:12-89) for a batch:
//   1. prf_R of every BASE layer (prf.cu; the expensive part),
//   2. PROD layers: R = R[pa]*R[pb] resolved per ciphertext (any DAG order; cycles / bad parents are reported where the
//      reference aborts),
//   3. one Fermat inversion per layer,
//   4. signed edge sum  acc = sum_e +- w_e * g^idx_e * Rinv[layer_e]: reads 23 B per edge (w, layer_id, idx, ch), never
//      sigma -- in the structure-of-arrays layout this stage is a pure HBM stream.
#include "engine.h"

namespace pvacb {

__global__ void dec_flags_kernel(uint64_t nL, const uint8_t* __restrict__ rule, uint8_t* __restrict__ flags) {
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nL) flags[l] = rule[l] == 0 ? 2 : 0;   // active, family 0 (prf_R)
}

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

// WARPS_PER_ITEM == 0: one warp per ciphertext (small ciphertexts); otherwise one CTA of 256 threads per ciphertext
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
    Fp acc = fp_zero();
    if (item < n) {
        uint32_t l0 = loff[item], e0 = eoff[item], E = eoff[item + 1] - e0;
        uint32_t start = kBlockPerItem ? threadIdx.x : lane, step = kBlockPerItem ? 256 : 32;
        for (uint32_t e = start; e < E; e += step) {
            Fp term = fp_mul(fp_mul(w[e0 + e], s_g[idx[e0 + e]]), Rinv[l0 + lid[e0 + e]]);
            acc = ch[e0 + e] == 0 ? fp_add(acc, term) : fp_sub(acc, term);
        }
    }
    acc = warp_sum_fp(acc);
    if (!kBlockPerItem) {
        if (lane == 0 && item < n) out[item] = acc;
    } else {
        if (lane == 0) s_part[wid] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            Fp t = s_part[0];
            for (int k = 1; k < 8; k++) t = fp_add(t, s_part[k]);
            out[item] = t;
        }
    }
}

int op_dec_value(Ctx* ctx, const Batch* Cb, uint64_t* h_out) {
    if (Cb->n == 0) return PV_OK;
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
        dec_flags_kernel<<<(unsigned)((Cb->nL + 255) / 256), 256, 0, ctx->stream>>>(Cb->nL, Cb->rule, flags);
        ctx->stat_kernel_launches += 1;
        if ((rc = prf_run(ctx, Cb->nL, Cb->ztag, Cb->nlo, Cb->nhi, flags, R, nullptr))) return rc;
        dec_resolve_kernel<<<(unsigned)((Cb->n + 127) / 128), 128, 0, ctx->stream>>>(Cb->n, Cb->loff, Cb->rule, Cb->pa, Cb->pb, R, err);
        dec_inv_kernel<<<(unsigned)((Cb->nL + 127) / 128), 128, 0, ctx->stream>>>(Cb->nL, R, Rinv);
        ctx->stat_kernel_launches += 2;
    }
    bool block_per_item = Cb->nE / Cb->n > 256;
    {
    ProfScope ps(ctx, PROF_DEC_EDGES);
    if (block_per_item)
        dec_edges_kernel<true><<<(unsigned)Cb->n, 256, 0, ctx->stream>>>(Cb->n, Cb->loff, Cb->eoff, Cb->lid, Cb->idx, Cb->ch, Cb->w, ctx->kv.powg, Rinv, res);
    else
        dec_edges_kernel<false><<<(unsigned)((Cb->n + 7) / 8), 256, 0, ctx->stream>>>(Cb->n, Cb->loff, Cb->eoff, Cb->lid, Cb->idx, Cb->ch, Cb->w, ctx->kv.powg, Rinv, res);
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
