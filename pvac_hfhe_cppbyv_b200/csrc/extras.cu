// Entry points around the hot path: batch slicing, synthetic benchmark batches, and the device building blocks exposed
// for parity tests (pvacb_prf, pvacb_sigma_from_H, pvacb_fp_op). All of them run CUDA kernels; none has a CPU path.
#include "engine.h"
#include <cstdlib>
#include <cstdio>
#include "../../include/pvacb.h"

#include <vector>

using namespace pvacb;

namespace pvacb {

__global__ void fp_op_kernel(int op, uint64_t n, const Fp* __restrict__ a, const Fp* __restrict__ b, Fp* __restrict__ o) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x = a[i], y = b ? b[i] : fp_zero(), r;
    switch (op) {
        case 0: r = fp_add(x, y); break;
        case 1: r = fp_sub(x, y); break;
        case 2: r = fp_mul(x, y); break;
        case 3: r = fp_neg(x); break;
        default: r = fp_inv(x); break;
    }
    o[i] = r;
}

__global__ void fill_flags_kernel(uint64_t n, uint8_t v, uint8_t* f) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = v;
}

// synthetic fresh-shaped ciphertexts: L = 2 BASE layers, epl edges per layer, uniform fields
__global__ void synth_meta_kernel(uint64_t n, int epl, uint64_t seed, uint32_t* loff, uint32_t* eoff, uint8_t* rule, uint64_t* ztag, uint64_t* nlo,
                                  uint64_t* nhi, uint32_t* pa, uint32_t* pb, uint32_t* lid, uint16_t* idx, uint8_t* ch, Fp* w) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    loff[i] = (uint32_t)(2 * i);
    eoff[i] = (uint32_t)(2 * epl * i);
    if (i == n) return;
    Tape t = tape_splitmix(item_stream_state(seed, i));     // benchmark data, not secret: always SplitMix64
    for (int l = 0; l < 2; l++) {
        uint64_t L = 2 * i + l;
        rule[L] = 0; nlo[L] = t.next(); nhi[L] = t.next(); ztag[L] = t.next(); pa[L] = 0; pb[L] = 0;
        for (int k = 0; k < epl; k++) {
            uint64_t e = (2 * i + l) * (uint64_t)epl + k;
            lid[e] = l;
            idx[e] = (uint16_t)(t.next() % kB);
            ch[e] = (uint8_t)(t.next() & 1);
            w[e] = fp_from_words(t.next(), t.next() & kMask63);
        }
    }
}
__global__ void synth_sigma_kernel(uint64_t nwords, uint64_t seed, uint64_t* sigma) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nwords; i += stride) sigma[i] = mix64(seed ^ (i * 0x9E3779B97F4A7C15ull));
}

// L2 gather probe: the access pattern of sigma_gather_kernel (a warp XOR-reads pseudo-random 1 KiB columns of the 16 MiB
// matrix H with 128-bit loads) with nothing else around it. U = columns in flight per warp. The best bandwidth over a
// few (U, CTAs/SM) shapes is used as the practical ceiling of the sigma kernel.
// LD: 0 ld.global.cg (what the sigma kernel uses), 1 plain ld.global, 2 ld.global.nc, 3 ld.global.cs, 4 one 256-bit ld.global.cg per
// lane and column instead of two 128-bit ones  (PVACB_PROBE_LD; 5 = the 256-bit form with 4 columns in flight; 6 / 7 / 8 = the TMA
// form below with 4 / 8 / 16 KiB in flight per warp)
template <int LD>
__device__ __forceinline__ uint4 probe_load(const uint4* p) {
    if (LD == 1) return *p;
    if (LD == 2) return __ldg(p);
    if (LD == 3) return __ldcs(p);
    return __ldcg(p);
}
// PIPE > 1 unrolls the column loop so that ptxas issues the next batch of loads before the last XORs of the current one, like
// the gather loop of sigma_fused_kernel (16 columns per trip).
template <int U, int LD = 0, int PIPE = 1>
__global__ void __launch_bounds__(256) l2_gather_probe_kernel(const uint4* __restrict__ H4, uint32_t cols_per_warp, uint4* __restrict__ sink) {
    const int lane = threadIdx.x & 31;
    uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t state = w * 2654435761u + 12345u;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
#pragma unroll PIPE
    for (uint32_t i = 0; i < cols_per_warp; i += U) {
        uint4 v[2 * U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            state = state * 1664525u + 1013904223u;
            uint32_t col = (state >> 10) & (kNBits - 1);
            if (LD == 4) {
                const uint4* p = H4 + (size_t)col * 64 + 2 * lane;
                asm volatile("ld.global.cg.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(v[2 * k].x), "=r"(v[2 * k].y), "=r"(v[2 * k].z), "=r"(v[2 * k].w), "=r"(v[2 * k + 1].x), "=r"(v[2 * k + 1].y),
                               "=r"(v[2 * k + 1].z), "=r"(v[2 * k + 1].w)
                             : "l"(p));
                continue;
            }
            const uint4* p = H4 + (size_t)col * 64 + lane;
            v[2 * k] = probe_load<LD>(p);
            v[2 * k + 1] = probe_load<LD>(p + 32);
        }
#pragma unroll
        for (int k = 0; k < U; k++) {
            a0.x ^= v[2 * k].x; a0.y ^= v[2 * k].y; a0.z ^= v[2 * k].z; a0.w ^= v[2 * k].w;
            a1.x ^= v[2 * k + 1].x; a1.y ^= v[2 * k + 1].y; a1.z ^= v[2 * k + 1].z; a1.w ^= v[2 * k + 1].w;
        }
    }
    a0.x ^= a1.x; a0.y ^= a1.y; a0.z ^= a1.z; a0.w ^= a1.w;
    if ((a0.x ^ a0.y ^ a0.z ^ a0.w) == 0x9E3779B9u && cols_per_warp == 0xFFFFFFFFu) sink[w] = a0;   // keeps the loads alive
}

// The same gather through the TMA engine: every column is one 1 KiB bulk copy (cp.async.bulk, global -> shared memory, completion on an
// mbarrier) into a per-warp ring of R slots, then two 128-bit shared-memory loads per lane and the XORs. Measures whether staging
// the columns through shared memory moves the ceiling (it does not: profiles/r01_notes.md).
template <int R>
__global__ void __launch_bounds__(256) l2_gather_probe_tma_kernel(const uint4* __restrict__ H4, uint32_t cols_per_warp, uint4* __restrict__ sink) {
    extern __shared__ __align__(128) unsigned char probe_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* ring = reinterpret_cast<uint4*>(probe_smem) + (size_t)warp * R * 64;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(probe_smem + (size_t)8 * R * 1024) + warp * R;
    auto s32 = [](const void* q) { return (uint32_t)__cvta_generic_to_shared(q); };
    if (lane == 0)
        for (int r = 0; r < R; r++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[r])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t state = w * 2654435761u + 12345u;
    auto issue = [&](int slot) {
        state = state * 1664525u + 1013904223u;
        const uint32_t col = (state >> 10) & (kNBits - 1);
        const uint32_t bar = s32(&bars[slot]), dst = s32(ring + slot * 64);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];" ::"r"(dst),
                     "l"(H4 + (size_t)col * 64), "r"(bar)
                     : "memory");
    };
    if (lane == 0)
        for (int r = 0; r < R; r++) issue(r);
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    for (uint32_t i = 0; i < cols_per_warp; i++) {
        const int slot = i % R;
        const uint32_t parity = (i / R) & 1, bar = s32(&bars[slot]);
        uint32_t ok = 0;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        } while (!ok);
        const uint4 v0 = ring[slot * 64 + lane], v1 = ring[slot * 64 + 32 + lane];
        a0.x ^= v0.x; a0.y ^= v0.y; a0.z ^= v0.z; a0.w ^= v0.w;
        a1.x ^= v1.x; a1.y ^= v1.y; a1.z ^= v1.z; a1.w ^= v1.w;
        __syncwarp();
        if (lane == 0 && i + R < cols_per_warp) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the slot was read through the generic proxy
            issue(slot);
        }
    }
    a0.x ^= a1.x; a0.y ^= a1.y; a0.z ^= a1.z; a0.w ^= a1.w;
    if ((a0.x ^ a0.y ^ a0.z ^ a0.w) == 0x9E3779B9u && cols_per_warp == 0xFFFFFFFFu) sink[w] = a0;
}

// order-independent checksums of a whole batch (parity checks at sizes that cannot be exported):
//   out[0] XOR of all sigma words   out[1], out[2] sums of w.lo, w.hi   out[3] sum of layer ids   out[4] sum of idx
//   out[5] sum of ch   out[6] sum over layers of (ztag ^ nonce.lo ^ nonce.hi) for BASE, pa + (pb << 32) for PROD   (all mod 2^64)
__global__ void __launch_bounds__(256) checksum_kernel(uint64_t nE, uint64_t nL, const uint64_t* __restrict__ sigma, const Fp* __restrict__ w,
                                                       const uint32_t* __restrict__ lid, const uint16_t* __restrict__ idx, const uint8_t* __restrict__ ch,
                                                       const uint8_t* __restrict__ rule, const uint64_t* __restrict__ ztag, const uint64_t* __restrict__ nlo,
                                                       const uint64_t* __restrict__ nhi, const uint32_t* __restrict__ pa, const uint32_t* __restrict__ pb,
                                                       unsigned long long* __restrict__ out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    uint64_t x = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0;
    const ulonglong2* s2p = reinterpret_cast<const ulonglong2*>(sigma);
    for (uint64_t i = tid; i < nE * (kMWords / 2); i += nth) { ulonglong2 v = s2p[i]; x ^= v.x ^ v.y; }
    for (uint64_t e = tid; e < nE; e += nth) { s1 += w[e].lo; s2 += w[e].hi; s3 += lid[e]; s4 += idx[e]; s5 += ch[e]; }
    for (uint64_t l = tid; l < nL; l += nth) {
        if (rule[l] == 0) s6 += ztag[l] ^ nlo[l] ^ nhi[l];
        else s6 += (uint64_t)pa[l] + ((uint64_t)pb[l] << 32);
    }
    for (int o = 16; o > 0; o >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
        s4 += __shfl_xor_sync(0xffffffffu, s4, o); s5 += __shfl_xor_sync(0xffffffffu, s5, o); s6 += __shfl_xor_sync(0xffffffffu, s6, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicXor(out + 0, (unsigned long long)x);
        atomicAdd(out + 1, (unsigned long long)s1); atomicAdd(out + 2, (unsigned long long)s2); atomicAdd(out + 3, (unsigned long long)s3);
        atomicAdd(out + 4, (unsigned long long)s4); atomicAdd(out + 5, (unsigned long long)s5); atomicAdd(out + 6, (unsigned long long)s6);
    }
}

__global__ void slice_fix_offsets_kernel(uint64_t cnt, const uint32_t* src_l, const uint32_t* src_e, uint32_t* dst_l, uint32_t* dst_e) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > cnt) return;
    dst_l[i] = src_l[i] - src_l[0];
    dst_e[i] = src_e[i] - src_e[0];
}

}  // namespace pvacb

static inline Ctx* C(pvacb_ctx* x) { return reinterpret_cast<Ctx*>(x); }
static inline const Batch* Bt(const pvacb_batch* x) { return reinterpret_cast<const Batch*>(x); }

extern "C" {

int pvacb_fp_op(pvacb_ctx* x, int op, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    Ctx* ctx = C(x);
    if (!a || !out || op < 0 || op > 4) return PV_E_ARG;
    if (n == 0) return PV_OK;
    cudaSetDevice(ctx->device);
    Fp *da = nullptr, *db = nullptr, *dout = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, (void**)&da, n * 16))) return rc;
    if ((rc = dev_alloc(ctx, (void**)&dout, n * 16))) return rc;
    PV_CUDA(cudaMemcpyAsync(da, a, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    if (b) {
        if ((rc = dev_alloc(ctx, (void**)&db, n * 16))) return rc;
        PV_CUDA(cudaMemcpyAsync(db, b, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    }
    fp_op_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(op, n, da, db, dout);
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    PV_CUDA(cudaMemcpyAsync(out, dout, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, da); dev_free(ctx, db); dev_free(ctx, dout);
    return PV_OK;
}

int pvacb_prf(pvacb_ctx* x, size_t n, const uint64_t* ztag, const uint64_t* nlo, const uint64_t* nhi, int family, uint64_t* out, uint64_t* ybits) {
    Ctx* ctx = C(x);
    if (!ctx->have_keys) return PV_E_NOKEYS;
    if (!ztag || !nlo || !nhi || !out) return PV_E_ARG;
    if (n == 0) return PV_OK;
    cudaSetDevice(ctx->device);
    uint64_t *dz, *dl, *dh, *dy = nullptr;
    uint8_t* fl;
    Fp* dout;
    int rc;
    const size_t wpc = ctx->prf_mode == PRF_LIVE ? 2 : kLpnT / 64;
    if ((rc = dev_alloc(ctx, (void**)&dz, n * 8)) || (rc = dev_alloc(ctx, (void**)&dl, n * 8)) || (rc = dev_alloc(ctx, (void**)&dh, n * 8)) ||
        (rc = dev_alloc(ctx, (void**)&fl, n)) || (rc = dev_alloc(ctx, (void**)&dout, n * 16)))
        return rc;
    if (ybits && (rc = dev_alloc(ctx, (void**)&dy, n * 3 * wpc * 8))) return rc;
    PV_CUDA(cudaMemcpyAsync(dz, ztag, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(dl, nlo, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(dh, nhi, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    fill_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, (uint8_t)(2 | (family & 1)), fl);
    ctx->stat_kernel_launches += 1;
    rc = prf_run(ctx, n, dz, dl, dh, fl, dout, dy);
    if (!rc) {
        PV_CUDA(cudaMemcpyAsync(out, dout, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
        if (ybits) PV_CUDA(cudaMemcpyAsync(ybits, dy, n * 3 * wpc * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    dev_free(ctx, dz); dev_free(ctx, dl); dev_free(ctx, dh); dev_free(ctx, fl); dev_free(ctx, dout); dev_free(ctx, dy);
    return rc;
}

int pvacb_sigma_from_H(pvacb_ctx* x, size_t n, const uint64_t* ztag, const uint64_t* nlo, const uint64_t* nhi, const uint16_t* idx, const uint8_t* ch,
                       const uint64_t* salt, uint64_t* out) {
    Ctx* ctx = C(x);
    if (!ctx->have_keys) return PV_E_NOKEYS;
    if (!ztag || !nlo || !nhi || !idx || !ch || !salt || !out) return PV_E_ARG;
    if (n == 0) return PV_OK;
    cudaSetDevice(ctx->device);
    uint64_t *dz, *dl, *dh, *ds, *dout;
    uint16_t* di;
    uint8_t* dc;
    int rc;
    if ((rc = dev_alloc(ctx, (void**)&dz, n * 8)) || (rc = dev_alloc(ctx, (void**)&dl, n * 8)) || (rc = dev_alloc(ctx, (void**)&dh, n * 8)) ||
        (rc = dev_alloc(ctx, (void**)&ds, n * 8)) || (rc = dev_alloc(ctx, (void**)&di, n * 2)) || (rc = dev_alloc(ctx, (void**)&dc, n)) ||
        (rc = dev_alloc(ctx, (void**)&dout, n * (size_t)kMWords * 8)))
        return rc;
    PV_CUDA(cudaMemcpyAsync(dz, ztag, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(dl, nlo, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(dh, nhi, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(ds, salt, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(di, idx, n * 2, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(dc, ch, n, cudaMemcpyHostToDevice, ctx->stream));
    SigmaJobs J;
    J.n = n; J.ztag = dz; J.nlo = dl; J.nhi = dh; J.idx = di; J.ch = dc; J.salt = ds; J.out = dout;
    rc = sigma_run(ctx, J);
    if (!rc) {
        PV_CUDA(cudaMemcpyAsync(out, dout, n * (size_t)kMWords * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    dev_free(ctx, dz); dev_free(ctx, dl); dev_free(ctx, dh); dev_free(ctx, ds); dev_free(ctx, di); dev_free(ctx, dc); dev_free(ctx, dout);
    return rc;
}

int pvacb_batch_synthetic(pvacb_ctx* x, size_t n, int epl, uint64_t seed, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out || epl < 0 || epl > 4096) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    if ((uint64_t)n * 2 * epl >= (1ull << 32)) return PV_E_SHAPE;
    Batch* b = nullptr;
    int rc = batch_alloc(ctx, n, 2 * n, (uint64_t)n * 2 * epl, &b);
    if (rc) return rc;
    synth_meta_kernel<<<(unsigned)((n + 1 + 127) / 128), 128, 0, ctx->stream>>>(n, epl, seed, b->loff, b->eoff, b->rule, b->ztag, b->nlo, b->nhi, b->pa, b->pb,
                                                                                 b->lid, b->idx, b->ch, b->w);
    uint64_t nwords = b->nE * (uint64_t)kMWords;
    if (nwords) synth_sigma_kernel<<<(unsigned)ctx->sm_count * 16, 256, 0, ctx->stream>>>(nwords, seed, b->sigma);
    PV_CUDA(cudaGetLastError());
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stat_kernel_launches += 2;
    *out = reinterpret_cast<pvacb_batch*>(b);
    return PV_OK;
}

// measured L2 gather bandwidth (GB/s) of this device for the H matrix access pattern; best over shapes and `reps`
int pvacb_l2_gather_probe(pvacb_ctx* x, int reps, double* gbps_out) {
    Ctx* ctx = C(x);
    if (!ctx->have_keys || !gbps_out) return PV_E_NOKEYS;
    cudaSetDevice(ctx->device);
    const uint32_t cols_per_warp = 4096;
    uint4* sink = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, (void**)&sink, (size_t)ctx->sm_count * 16 * 8 * 16))) return rc;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0;
    const uint4* H4 = reinterpret_cast<const uint4*>(ctx->kv.H);
    static const int kCtas[5] = {4, 8, 1, 2, 3};
    // PVACB_PROBE_SMEM=<bytes per CTA>: run the probe with that much dynamic shared memory (shrinks L1), to see whether the
    // gather ceiling depends on the L1 / shared-memory split the sigma kernel runs with
#ifdef PVACB_TUNING
    auto env = [](const char* k) -> const char* { return getenv(k); };
#else
    auto env = [](const char*) -> const char* { return nullptr; };       // the probe's tuning switches exist in a -DPVACB_TUNING build only
#endif
    const int probe_smem = env("PVACB_PROBE_SMEM") ? atoi(env("PVACB_PROBE_SMEM")) : 0;
    if (env("PVACB_PROBE_CARVEOUT")) {
        const int pct = atoi(env("PVACB_PROBE_CARVEOUT"));
        cudaFuncSetAttribute(l2_gather_probe_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(l2_gather_probe_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(l2_gather_probe_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    cudaFuncSetAttribute(l2_gather_probe_tma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8 * 1024 + 8 * 8 * 8);
    cudaFuncSetAttribute(l2_gather_probe_tma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16 * 1024 + 8 * 16 * 8);
    if (probe_smem > 48 * 1024) {
        cudaFuncSetAttribute(l2_gather_probe_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, probe_smem);
        cudaFuncSetAttribute(l2_gather_probe_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, probe_smem);
        cudaFuncSetAttribute(l2_gather_probe_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, probe_smem);
    }
    const int nshapes = env("PVACB_PROBE_VERBOSE") ? 15 : 6;         // the first six shapes hold the maximum; the rest map the curve
    for (int sh = 0; sh < nshapes; sh++) {
        const int shape = sh < 6 ? sh : 2 * ((sh - 6) / 3) + 0;            // columns-in-flight index in shape / 2
        const int ctas_per_sm = sh < 6 ? ((sh & 1) ? 8 : 4) : kCtas[2 + (sh - 6) % 3];
        double shape_best = 0;
        const unsigned grid = (unsigned)ctx->sm_count * ctas_per_sm;
        for (int r = 0; r < reps + 2; r++) {
            cudaEventRecord(a, ctx->stream);
            const int ld = env("PVACB_PROBE_LD") ? atoi(env("PVACB_PROBE_LD")) : 0;
            if (ld == 1) l2_gather_probe_kernel<8, 1><<<grid, 256, probe_smem, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (ld == 2) l2_gather_probe_kernel<8, 2><<<grid, 256, probe_smem, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (ld == 3) l2_gather_probe_kernel<8, 3><<<grid, 256, probe_smem, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (ld == 4) l2_gather_probe_kernel<8, 4><<<grid, 256, probe_smem, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (ld == 5) l2_gather_probe_kernel<4, 4><<<grid, 256, probe_smem, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (ld == 6) l2_gather_probe_tma_kernel<4><<<grid, 256, 8 * 4 * 1024 + 8 * 4 * 8, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (ld == 7) l2_gather_probe_tma_kernel<8><<<grid, 256, 8 * 8 * 1024 + 8 * 8 * 8, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (ld == 8) l2_gather_probe_tma_kernel<16><<<grid, 256, 8 * 16 * 1024 + 8 * 16 * 8, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (shape / 2 == 0 && (sh & 1) == 0 && sh < 6) l2_gather_probe_kernel<4, 0, 4><<<grid, 256, probe_smem, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (shape / 2 == 0) l2_gather_probe_kernel<4><<<grid, 256, probe_smem, ctx->stream>>>(H4, cols_per_warp, sink);
            else if (shape / 2 == 1) l2_gather_probe_kernel<8><<<grid, 256, probe_smem, ctx->stream>>>(H4, cols_per_warp, sink);
            else l2_gather_probe_kernel<16><<<grid, 256, probe_smem, ctx->stream>>>(H4, cols_per_warp, sink);
            cudaEventRecord(b, ctx->stream);
            PV_CUDA(cudaEventSynchronize(b));
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            double bytes = (double)grid * 8 * cols_per_warp * 1024.0;
            double g = bytes / (ms * 1e-3) / 1e9;
            if (r > 1 && g > best) best = g;
            if (r > 1 && g > shape_best) shape_best = g;
        }
        if (env("PVACB_PROBE_VERBOSE")) fprintf(stderr, "l2 gather probe: %d columns in flight per warp%s, %d warps/SM: %.0f GB/s\n", 4 << (shape / 2), sh == 0 ? " (loop unrolled x4)" : "", ctas_per_sm * 8, shape_best);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    dev_free(ctx, sink);
    *gbps_out = best;
    return PV_OK;
}

// compact_edges (ops/encrypt.hpp:39-71) of every ciphertext of the batch: merge equal (layer, idx, sign), drop all-zero
// results, order by (layer, idx, P before M). The input batch is left untouched.
int pvacb_compact_edges(pvacb_ctx* x, const pvacb_batch* pb, pvacb_batch** out) {
    Ctx* ctx = C(x);
    const Batch* s = Bt(pb);
    if (!out) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    Batch* cpy = nullptr;
    int rc = batch_clone(ctx, s, &cpy);
    if (rc) return rc;
    if ((rc = guard_budget_batch(ctx, &cpy, 0))) { batch_free(cpy); return rc; }
    *out = reinterpret_cast<pvacb_batch*>(cpy);
    return PV_OK;
}

int pvacb_commit_ct(pvacb_ctx* x, const pvacb_batch* pb, uint8_t* out) {
    Ctx* ctx = C(x);
    if (!ctx->have_keys) return PV_E_NOKEYS;
    if (!out) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    return op_commit_ct(ctx, Bt(pb), out);
}

int pvacb_batch_checksum(pvacb_ctx* x, const pvacb_batch* pb, uint64_t out[8]) {
    Ctx* ctx = C(x);
    const Batch* b = Bt(pb);
    if (!out) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    unsigned long long* d = nullptr;
    int rc = dev_alloc(ctx, (void**)&d, 64);
    if (rc) return rc;
    PV_CUDA(cudaMemsetAsync(d, 0, 64, ctx->stream));
    checksum_kernel<<<(unsigned)ctx->sm_count * 8, 256, 0, ctx->stream>>>(b->nE, b->nL, b->sigma, b->w, b->lid, b->idx, b->ch, b->rule, b->ztag, b->nlo, b->nhi,
                                                                            b->pa, b->pb, d);
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    SmallRead sr;
    sr.add(out, d, 56);
    rc = read_small_sync(ctx, sr);
    dev_free(ctx, d);
    out[7] = b->nE;
    return rc;
}

int pvacb_batch_slice(pvacb_ctx* x, const pvacb_batch* pb, size_t first, size_t count, pvacb_batch** out) {
    Ctx* ctx = C(x);
    const Batch* s = Bt(pb);
    if (!out || first + count > s->n) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    uint32_t h[4];
    {
        SmallRead sr;
        sr.add(&h[0], s->loff + first, 4); sr.add(&h[1], s->loff + first + count, 4);
        sr.add(&h[2], s->eoff + first, 4); sr.add(&h[3], s->eoff + first + count, 4);
        int rc0 = read_small_sync(ctx, sr);
        if (rc0) return rc0;
    }
    uint64_t l0 = h[0], nL = h[1] - h[0], e0 = h[2], nE = h[3] - h[2];
    Batch* b = nullptr;
    int rc = batch_alloc(ctx, count, nL, nE, &b);
    if (rc) return rc;
    slice_fix_offsets_kernel<<<(unsigned)((count + 1 + 255) / 256), 256, 0, ctx->stream>>>(count, s->loff + first, s->eoff + first, b->loff, b->eoff);
    auto cp = [&](void* d, const void* sp, size_t bytes) { return bytes ? cudaMemcpyAsync(d, sp, bytes, cudaMemcpyDeviceToDevice, ctx->stream) : cudaSuccess; };
    PV_CUDA(cp(b->rule, s->rule + l0, nL));
    PV_CUDA(cp(b->ztag, s->ztag + l0, nL * 8));
    PV_CUDA(cp(b->nlo, s->nlo + l0, nL * 8));
    PV_CUDA(cp(b->nhi, s->nhi + l0, nL * 8));
    PV_CUDA(cp(b->pa, s->pa + l0, nL * 4));
    PV_CUDA(cp(b->pb, s->pb + l0, nL * 4));
    PV_CUDA(cp(b->lid, s->lid + e0, nE * 4));
    PV_CUDA(cp(b->idx, s->idx + e0, nE * 2));
    PV_CUDA(cp(b->ch, s->ch + e0, nE));
    PV_CUDA(cp(b->w, s->w + e0, nE * 16));
    PV_CUDA(cp(b->sigma, s->sigma + e0 * kMWords, nE * (size_t)kMWords * 8));
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    *out = reinterpret_cast<pvacb_batch*>(b);
    return PV_OK;
}

}  // extern "C"
