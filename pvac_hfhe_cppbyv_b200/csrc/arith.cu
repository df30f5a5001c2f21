// ct_add / ct_sub / ct_scale (ops/arithmetic.hpp:12-45) and compact_layers (ops/encrypt.hpp:73-104) on device batches.
//
// ct_add is a segmented concatenation: per ciphertext, layers of A then layers of B (PROD parents and B's edge layer ids
// shifted by |A.L|), edges of A then edges of B. 97% of the bytes are the 1 KiB sigma rows, which are contiguous per
// ciphertext in the structure-of-arrays layout, so the kernel is two streaming copies per ciphertext with 128-bit
// loads/stores: HBM-bound, algorithmic traffic = read(A)+read(B)+write(C).
#include "engine.h"

namespace pvacb {

__global__ void concat_offsets_kernel(uint64_t n, const uint32_t* __restrict__ la, const uint32_t* __restrict__ lb, const uint32_t* __restrict__ ea,
                                      const uint32_t* __restrict__ eb, uint32_t* __restrict__ lo, uint32_t* __restrict__ eo, uint32_t budget, unsigned int* __restrict__ err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    lo[i] = la[i] + lb[i];
    eo[i] = ea[i] + eb[i];
    if (i < n) {
        uint32_t ne = (ea[i + 1] - ea[i]) + (eb[i + 1] - eb[i]);
        if (ne > budget) atomicOr(err, 1u);   // guard_budget (ops/encrypt.hpp:106-111)
    }
}

constexpr int kConcatThreads = 256;
constexpr int kConcatChunk = 64;      // edges per CTA step (64 KiB of sigma)
constexpr int kConcatUnroll = 4;      // 128-bit loads in flight per thread (8 measured 2 % slower)
// plain ld/st, not the streaming (evict-first) __ldcs/__stcs hints: with the hints the copy ran at 6.07 TB/s, without at
// 6.56 TB/s (measured, 2^15 fresh pairs) -- the same rate as the MEASURED_PEAKS copy.
constexpr bool kConcatPlain = true;

struct BatchView {
    const uint32_t *loff, *eoff;
    const uint8_t* rule;
    const uint64_t *ztag, *nlo, *nhi;
    const uint32_t *pa, *pb, *lid;
    const uint16_t* idx;
    const uint8_t* ch;
    const Fp* w;
    const uint64_t* sigma;
};
static BatchView view_of(const Batch* b) {
    return BatchView{b->loff, b->eoff, b->rule, b->ztag, b->nlo, b->nhi, b->pa, b->pb, b->lid, b->idx, b->ch, b->w, b->sigma};
}

// grid = (n, ychunks). mode: 0 add, 1 sub (B's weights multiplied by p-1, ops/arithmetic.hpp:33-45)
__global__ void __launch_bounds__(kConcatThreads)
concat_kernel(BatchView A, BatchView B, int mode, uint32_t* __restrict__ o_loff, uint32_t* __restrict__ o_eoff, uint8_t* __restrict__ o_rule,
              uint64_t* __restrict__ o_ztag, uint64_t* __restrict__ o_nlo, uint64_t* __restrict__ o_nhi, uint32_t* __restrict__ o_pa,
              uint32_t* __restrict__ o_pb, uint32_t* __restrict__ o_lid, uint16_t* __restrict__ o_idx, uint8_t* __restrict__ o_ch,
              Fp* __restrict__ o_w, uint64_t* __restrict__ o_sigma) {
    const uint64_t i = blockIdx.x;
    const uint32_t la0 = A.loff[i], LA = A.loff[i + 1] - la0, lb0 = B.loff[i], LB = B.loff[i + 1] - lb0;
    const uint32_t ea0 = A.eoff[i], EA = A.eoff[i + 1] - ea0, eb0 = B.eoff[i], EB = B.eoff[i + 1] - eb0;
    const uint32_t lo0 = la0 + lb0, eo0 = ea0 + eb0;
    const int tid = threadIdx.x;
    if (blockIdx.y == 0) {
        for (uint32_t k = tid; k < LA + LB; k += kConcatThreads) {
            bool fromA = k < LA;
            uint32_t s = fromA ? la0 + k : lb0 + (k - LA);
            const BatchView& S = fromA ? A : B;
            uint8_t r = S.rule[s];
            uint32_t off = (fromA || r != 1) ? 0u : LA;
            o_rule[lo0 + k] = r;
            o_ztag[lo0 + k] = S.ztag[s];
            o_nlo[lo0 + k] = S.nlo[s];
            o_nhi[lo0 + k] = S.nhi[s];
            o_pa[lo0 + k] = S.pa[s] + off;
            o_pb[lo0 + k] = S.pb[s] + off;
        }
    }
    const Fp pm1 = fp_make(~0ull - 1, kMask63);   // p - 1
    for (uint32_t c0 = blockIdx.y * kConcatChunk; c0 < EA + EB; c0 += gridDim.y * kConcatChunk) {
        uint32_t c1 = min(c0 + kConcatChunk, EA + EB);
        // small per-edge fields
        for (uint32_t e = c0 + tid; e < c1; e += kConcatThreads) {
            bool fromA = e < EA;
            uint32_t s = fromA ? ea0 + e : eb0 + (e - EA);
            const BatchView& S = fromA ? A : B;
            o_lid[eo0 + e] = S.lid[s] + (fromA ? 0u : LA);
            o_idx[eo0 + e] = S.idx[s];
            o_ch[eo0 + e] = S.ch[s];
            Fp w = S.w[s];
            if (!fromA && mode == 1) w = fp_mul(w, pm1);
            o_w[eo0 + e] = w;
        }
        // sigma rows: 64 uint4 per edge; the chunk may straddle the A/B boundary
        uint4* dst = reinterpret_cast<uint4*>(o_sigma + (size_t)(eo0 + c0) * kMWords);
        const uint32_t nvec = (c1 - c0) * 64;
        const uint32_t splitv = c0 >= EA ? 0u : (min(c1, EA) - c0) * 64;   // vectors coming from A
        const uint4* srcA = reinterpret_cast<const uint4*>(A.sigma + (size_t)(ea0 + c0) * kMWords);
        const uint4* srcB = reinterpret_cast<const uint4*>(B.sigma + (size_t)(eb0 + (c0 >= EA ? c0 - EA : 0)) * kMWords) - splitv;
        for (uint32_t v = tid; v < nvec; v += kConcatThreads * kConcatUnroll) {
            uint4 r[kConcatUnroll];
#pragma unroll
            for (int k = 0; k < kConcatUnroll; k++) {
                uint32_t vv = v + k * kConcatThreads;
                if (vv < nvec) r[k] = kConcatPlain ? (vv < splitv ? srcA : srcB)[vv] : __ldcs((vv < splitv ? srcA : srcB) + vv);
            }
#pragma unroll
            for (int k = 0; k < kConcatUnroll; k++) {
                uint32_t vv = v + k * kConcatThreads;
                if (vv < nvec) { if (kConcatPlain) dst[vv] = r[k]; else __stcs(dst + vv, r[k]); }
            }
        }
    }
    (void)o_loff; (void)o_eoff;
}

int op_ct_add(Ctx* ctx, const Batch* A, const Batch* B, int mode, Batch** out) {
    Scratch scratch(ctx);
    Batch* o = nullptr;
    int rc = batch_alloc(ctx, A->n, A->nL + B->nL, A->nE + B->nE, &o);
    if (rc) return rc;
    unsigned int* err = nullptr;
    if ((rc = scratch.alloc(err, 4))) { batch_free(o); return rc; }
    PV_CUDA(cudaMemsetAsync(err, 0, 4, ctx->stream));
    uint64_t n = A->n;
    concat_offsets_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(n, A->loff, B->loff, A->eoff, B->eoff, o->loff, o->eoff, ctx->edge_budget, err);
    if (n) {
        uint64_t avg = (o->nE + n - 1) / n;
        unsigned ych = (unsigned)std::min<uint64_t>(std::max<uint64_t>((avg + kConcatChunk - 1) / kConcatChunk, 1), 4096);
        // CUDA grids allow 2^31-1 blocks in x and 65535 in y
        dim3 grid((unsigned)n, ych);
        ProfScope ps(ctx, PROF_CONCAT);
        concat_kernel<<<grid, kConcatThreads, 0, ctx->stream>>>(view_of(A), view_of(B), mode, o->loff, o->eoff, o->rule, o->ztag, o->nlo, o->nhi, o->pa,
                                                                 o->pb, o->lid, o->idx, o->ch, o->w, o->sigma);
    }
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += n ? 2 : 1;
    unsigned int h_err = 0;
    { SmallRead sr; sr.add(&h_err, err, 4); if ((rc = read_small_sync(ctx, sr))) { batch_free(o); return rc; } }
    if (h_err) {       // guard_budget(pk, C, "add") -> compact_edges, ops/arithmetic.hpp:28
        if ((rc = guard_budget_batch(ctx, &o, ctx->edge_budget))) { batch_free(o); return rc; }
    }
    rc = compact_layers_batch(ctx, o);
    if (rc) { batch_free(o); return rc; }
    *out = o;
    return PV_OK;
}

// ---- ct_scale (ops/arithmetic.hpp:33-37): plain copy with w <- w*s ; no compact_layers
__global__ void scale_w_kernel(uint64_t nE, const Fp* __restrict__ in, Fp s, Fp* __restrict__ out) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nE) out[e] = fp_mul(in[e], s);
}

int op_ct_scale(Ctx* ctx, const Batch* A, Fp s, Batch** out) {
    Batch* o = nullptr;
    int rc = batch_clone(ctx, A, &o);
    if (rc) return rc;
    if (A->nE) scale_w_kernel<<<(unsigned)((A->nE + 255) / 256), 256, 0, ctx->stream>>>(A->nE, A->w, s, o->w);
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    *out = o;
    return PV_OK;
}

// ------------------------------------------------------------------ compact_layers
// One CTA per ciphertext: mark layers that own an edge, close under PROD parents (fixpoint, as the reference), count.
__global__ void __launch_bounds__(128)
layers_mark_kernel(uint64_t n, const uint32_t* __restrict__ loff, const uint32_t* __restrict__ eoff, const uint8_t* __restrict__ rule,
                   const uint32_t* __restrict__ pa, const uint32_t* __restrict__ pb, const uint32_t* __restrict__ lid, uint8_t* __restrict__ used,
                   uint32_t* __restrict__ kept_count, unsigned int* __restrict__ err) {
    const uint64_t i = blockIdx.x;
    const uint32_t l0 = loff[i], L = loff[i + 1] - l0, e0 = eoff[i], E = eoff[i + 1] - e0;
    for (uint32_t e = threadIdx.x; e < E; e += blockDim.x) {
        uint32_t l = lid[e0 + e];
        if (l < L) used[l0 + l] = 1;
    }
    __syncthreads();
    __shared__ int changed;
    __shared__ uint32_t cnt;
    for (;;) {
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) {
            if (!used[l0 + l] || rule[l0 + l] != 1) continue;
            uint32_t a = pa[l0 + l], b = pb[l0 + l];
            if (a < L && !used[l0 + a]) { used[l0 + a] = 1; changed = 1; }
            if (b < L && !used[l0 + b]) { used[l0 + b] = 1; changed = 1; }
        }
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    uint32_t c = 0;
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) c += used[l0 + l];
    if (c) atomicAdd(&cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) kept_count[i] = cnt;
    (void)err;
}

// exclusive scan of per-item counts -> offsets (single CTA; n is at most a few million)
__global__ void __launch_bounds__(1024) scan_u32_kernel(uint64_t n, const uint32_t* __restrict__ in, uint32_t* __restrict__ out /*n+1*/) {
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    uint64_t per = (n + 1023) / 1024;
    uint64_t b = (uint64_t)t * per, e = b + per < n ? b + per : n;
    uint32_t s = 0;
    for (uint64_t i = b; i < e; i++) s += in[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        uint32_t run = 0;
        for (int k = 0; k < 1024; k++) { uint32_t v = part[k]; part[k] = run; run += v; }
        out[n] = run;
    }
    __syncthreads();
    uint32_t run = part[t];
    for (uint64_t i = b; i < e; i++) { out[i] = run; run += in[i]; }
}

int scan_u32(Ctx* ctx, uint64_t n, const uint32_t* in, uint32_t* out) {
    scan_u32_kernel<<<1, 1024, 0, ctx->stream>>>(n, in, out);
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    return PV_OK;
}

// one CTA per ciphertext: remap table, compacted layer rows into temporaries, edge layer ids rewritten in place
__global__ void __launch_bounds__(128)
layers_remap_kernel(const uint32_t* __restrict__ loff, const uint32_t* __restrict__ eoff, const uint32_t* __restrict__ new_loff,
                    const uint8_t* __restrict__ used, uint32_t* __restrict__ remap, const uint8_t* __restrict__ rule, const uint64_t* __restrict__ ztag,
                    const uint64_t* __restrict__ nlo, const uint64_t* __restrict__ nhi, const uint32_t* __restrict__ pa, const uint32_t* __restrict__ pb,
                    uint8_t* __restrict__ t_rule, uint64_t* __restrict__ t_ztag, uint64_t* __restrict__ t_nlo, uint64_t* __restrict__ t_nhi,
                    uint32_t* __restrict__ t_pa, uint32_t* __restrict__ t_pb, uint32_t* __restrict__ lid) {
    const uint64_t i = blockIdx.x;
    const uint32_t l0 = loff[i], L = loff[i + 1] - l0, e0 = eoff[i], E = eoff[i + 1] - e0, n0 = new_loff[i];
    if (threadIdx.x == 0) {
        uint32_t k = 0;
        for (uint32_t l = 0; l < L; l++) remap[l0 + l] = used[l0 + l] ? k++ : 0xFFFFFFFFu;
    }
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) {
        uint32_t r = remap[l0 + l];
        if (r == 0xFFFFFFFFu) continue;
        uint8_t ru = rule[l0 + l];
        t_rule[n0 + r] = ru;
        t_ztag[n0 + r] = ztag[l0 + l];
        t_nlo[n0 + r] = nlo[l0 + l];
        t_nhi[n0 + r] = nhi[l0 + l];
        uint32_t a = pa[l0 + l], b = pb[l0 + l];
        if (ru == 1) { a = a < L ? remap[l0 + a] : a; b = b < L ? remap[l0 + b] : b; }
        t_pa[n0 + r] = a;
        t_pb[n0 + r] = b;
    }
    for (uint32_t e = threadIdx.x; e < E; e += blockDim.x) {
        uint32_t l = lid[e0 + e];
        if (l < L) lid[e0 + e] = remap[l0 + l];
    }
}

// compact_layers of every ciphertext of b, in place (layer arrays shrink; edge arrays keep their place).
int compact_layers_batch(Ctx* ctx, Batch* b, const unsigned int* d_err, unsigned int* h_err) {
    if (b->n == 0 || b->nL == 0) {
        if (d_err) { SmallRead sr; sr.add(h_err, d_err, 4); return read_small_sync(ctx, sr); }
        return PV_OK;
    }
    Scratch scratch(ctx);
    uint8_t* used = nullptr;
    uint32_t *cnt = nullptr, *noff = nullptr;
    int rc;
    if ((rc = scratch.alloc(used, b->nL))) return rc;
    if ((rc = scratch.alloc(cnt, b->n * 4))) return rc;
    if ((rc = scratch.alloc(noff, (b->n + 1) * 4))) return rc;
    PV_CUDA(cudaMemsetAsync(used, 0, b->nL, ctx->stream));
    layers_mark_kernel<<<(unsigned)b->n, 128, 0, ctx->stream>>>(b->n, b->loff, b->eoff, b->rule, b->pa, b->pb, b->lid, used, cnt, nullptr);
    if ((rc = scan_u32(ctx, b->n, cnt, noff))) return rc;
    uint32_t total = 0;
    { SmallRead sr; sr.add(&total, noff + b->n, 4); if (d_err) sr.add(h_err, d_err, 4); if ((rc = read_small_sync(ctx, sr))) return rc; }
    ctx->stat_kernel_launches += 1;
    if (total != b->nL) {
        uint32_t* remap = nullptr;
        uint8_t* t_rule = nullptr;
        uint64_t *t_ztag = nullptr, *t_nlo = nullptr, *t_nhi = nullptr;
        uint32_t *t_pa = nullptr, *t_pb = nullptr;
        if ((rc = scratch.alloc(remap, b->nL * 4))) return rc;
        if ((rc = scratch.alloc(t_rule, total))) return rc;
        if ((rc = scratch.alloc(t_ztag, (size_t)total * 8))) return rc;
        if ((rc = scratch.alloc(t_nlo, (size_t)total * 8))) return rc;
        if ((rc = scratch.alloc(t_nhi, (size_t)total * 8))) return rc;
        if ((rc = scratch.alloc(t_pa, (size_t)total * 4))) return rc;
        if ((rc = scratch.alloc(t_pb, (size_t)total * 4))) return rc;
        layers_remap_kernel<<<(unsigned)b->n, 128, 0, ctx->stream>>>(b->loff, b->eoff, noff, used, remap, b->rule, b->ztag, b->nlo, b->nhi, b->pa, b->pb,
                                                                      t_rule, t_ztag, t_nlo, t_nhi, t_pa, t_pb, b->lid);
        PV_CUDA(cudaMemcpyAsync(b->rule, t_rule, total, cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA(cudaMemcpyAsync(b->ztag, t_ztag, (size_t)total * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA(cudaMemcpyAsync(b->nlo, t_nlo, (size_t)total * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA(cudaMemcpyAsync(b->nhi, t_nhi, (size_t)total * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA(cudaMemcpyAsync(b->pa, t_pa, (size_t)total * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA(cudaMemcpyAsync(b->pb, t_pb, (size_t)total * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA(cudaMemcpyAsync(b->loff, noff, (b->n + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA(cudaGetLastError());
        ctx->stat_kernel_launches += 1;
        b->nL = total;
    }
    return PV_OK;
}

}  // namespace pvacb
