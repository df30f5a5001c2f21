// guard_budget + compact_edges (ops/encrypt.hpp:39-71, 106-111) for device batches.
//
// The reference calls guard_budget at the end of ct_add / ct_mul: a ciphertext with more than Params::edge_budget
// (1.2 M) edges is rebuilt by compact_edges -- edges with equal (layer, idx, sign) are merged (weights added, sigmas
// XORed), merged edges with w == 0 and sigma == 0 are dropped, and the survivors come out ordered by (layer, idx, P before
// M). The tenth multiplication of examples/basic_usage.cpp's "perf 10 muls" is the first place that happens (1.38 M edges).
// Batched restatement, applied only to the ciphertexts of a batch that are over budget:
//   keys  = (item, layer, idx, sign) of their edges  -> one cub radix sort (stable)
//   heads = first edge of every run of equal keys; the run's weight sum; runs whose sum is zero get their sigma XOR
//           checked by a warp (the drop rule);
//   the batch is rebuilt: untouched ciphertexts are copied, compacted ones are written run by run (a warp XORs the run's
//   1 KiB rows).
#include "engine.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace pvacb {

// flag[i] = 1 if ciphertext i is over budget; fcnt[i] = its edge count (else 0)
__global__ void cmp_flag_kernel(uint64_t n, const uint32_t* __restrict__ eoff, uint32_t budget, uint8_t* __restrict__ flag, uint32_t* __restrict__ fcnt) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = eoff[i + 1] - eoff[i];
    bool f = c > budget;
    flag[i] = f ? 1 : 0;
    fcnt[i] = f ? c : 0;
}

// one CTA per flagged ciphertext: sort keys and source indices of its edges
__global__ void cmp_keys_kernel(const uint32_t* __restrict__ eoff, const uint8_t* __restrict__ flag, const uint32_t* __restrict__ foff,
                                const uint32_t* __restrict__ lid, const uint16_t* __restrict__ idx, const uint8_t* __restrict__ ch,
                                uint64_t* __restrict__ key, uint32_t* __restrict__ src) {
    const uint64_t i = blockIdx.x;
    if (!flag[i]) return;
    const uint32_t e0 = eoff[i], E = eoff[i + 1] - e0, f0 = foff[i];
    for (uint32_t k = blockIdx.y * blockDim.x + threadIdx.x; k < E; k += gridDim.y * blockDim.x) {
        uint32_t e = e0 + k;
        key[f0 + k] = (i << 40) | ((uint64_t)lid[e] << 10) | ((uint64_t)idx[e] << 1) | (uint64_t)(ch[e] & 1);
        src[f0 + k] = e;
    }
}

// per sorted position: head of a run? runs are short (length 1 unless the operand carried duplicates)
__global__ void cmp_heads_kernel(uint32_t M, const uint64_t* __restrict__ key, const uint32_t* __restrict__ src, const Fp* __restrict__ w,
                                 uint32_t* __restrict__ keep, Fp* __restrict__ wsum, uint32_t* __restrict__ zero_list, uint32_t* __restrict__ zero_cnt) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    uint64_t k = key[j];
    if (j > 0 && key[j - 1] == k) { keep[j] = 0; return; }
    Fp s = fp_add(fp_zero(), w[src[j]]);
    for (uint32_t q = j + 1; q < M && key[q] == k; q++) s = fp_add(s, w[src[q]]);
    wsum[j] = s;
    keep[j] = 1;
    if (fp_is_zero(s)) zero_list[atomicAdd(zero_cnt, 1u)] = j;     // keep only if the XOR of the run's sigmas is non-zero
}

// one warp per zero-weight run: keep[j] = (XOR of the run's sigma rows != 0)   (ops/encrypt.hpp:59)
__global__ void cmp_zero_runs_kernel(const uint32_t* __restrict__ zero_cnt, const uint32_t* __restrict__ zero_list, uint32_t M, const uint64_t* __restrict__ key,
                                     const uint32_t* __restrict__ src, const uint64_t* __restrict__ sigma, uint32_t* __restrict__ keep) {
    const int lane = threadIdx.x & 31;
    uint32_t wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wi >= *zero_cnt) return;
    uint32_t j = zero_list[wi];
    uint64_t k = key[j];
    uint64_t acc[4] = {0, 0, 0, 0};
    for (uint32_t q = j; q < M && key[q] == k; q++) {
        const uint64_t* row = sigma + (size_t)src[q] * kMWords;
        for (int t = 0; t < 4; t++) acc[t] ^= row[lane + 32 * t];
    }
    bool nz = (acc[0] | acc[1] | acc[2] | acc[3]) != 0;
    if (__ballot_sync(0xffffffffu, nz) == 0 && lane == 0) keep[j] = 0;
}

// new edge count per ciphertext: untouched ones keep theirs, compacted ones count their kept runs (kpos = exclusive scan of keep)
__global__ void cmp_counts_kernel(uint64_t n, const uint32_t* __restrict__ eoff, const uint8_t* __restrict__ flag, const uint32_t* __restrict__ foff,
                                  const uint32_t* __restrict__ kpos, uint32_t M, uint32_t kept_total, uint32_t* __restrict__ cnt) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!flag[i]) { cnt[i] = eoff[i + 1] - eoff[i]; return; }
    uint32_t a = foff[i], b = foff[i + 1];
    uint32_t ka = kpos[a], kb = b >= M ? kept_total : kpos[b];
    cnt[i] = kb - ka;
}

// one CTA per ciphertext writes its edges into the rebuilt batch
__global__ void __launch_bounds__(256)
cmp_write_kernel(const uint32_t* __restrict__ eoff, const uint32_t* __restrict__ neoff, const uint8_t* __restrict__ flag, const uint32_t* __restrict__ foff,
                 const uint32_t* __restrict__ kpos, const uint32_t* __restrict__ keep, uint32_t M, const uint64_t* __restrict__ key,
                 const uint32_t* __restrict__ src, const Fp* __restrict__ wsum,
                 const uint32_t* __restrict__ lid, const uint16_t* __restrict__ idx, const uint8_t* __restrict__ ch, const Fp* __restrict__ w,
                 const uint64_t* __restrict__ sigma, uint32_t* __restrict__ o_lid, uint16_t* __restrict__ o_idx, uint8_t* __restrict__ o_ch,
                 Fp* __restrict__ o_w, uint64_t* __restrict__ o_sigma) {
    const uint64_t i = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t e0 = eoff[i], E = eoff[i + 1] - e0, n0 = neoff[i];
    const uint32_t wstep = nw * gridDim.y, w0 = blockIdx.y * nw + wid;
    if (!flag[i]) {
        for (uint32_t k = w0; k < E; k += wstep) {
            const uint32_t e = e0 + k, o = n0 + k;
            if (lane == 0) { o_lid[o] = lid[e]; o_idx[o] = idx[e]; o_ch[o] = ch[e]; o_w[o] = w[e]; }
            const uint4* s = reinterpret_cast<const uint4*>(sigma + (size_t)e * kMWords);
            uint4* d = reinterpret_cast<uint4*>(o_sigma + (size_t)o * kMWords);
            d[lane] = s[lane];
            d[lane + 32] = s[lane + 32];
        }
        return;
    }
    const uint32_t f0 = foff[i], kbase = kpos[f0];
    for (uint32_t k = w0; k < E; k += wstep) {
        const uint32_t j = f0 + k;
        if (!keep[j]) continue;
        const uint32_t o = n0 + (kpos[j] - kbase);
        const uint64_t kk = key[j];
        uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
        for (uint32_t q = j; q < M && key[q] == kk; q++) {
            const uint4* s = reinterpret_cast<const uint4*>(sigma + (size_t)src[q] * kMWords);
            uint4 v0 = s[lane], v1 = s[lane + 32];
            a0.x ^= v0.x; a0.y ^= v0.y; a0.z ^= v0.z; a0.w ^= v0.w;
            a1.x ^= v1.x; a1.y ^= v1.y; a1.z ^= v1.z; a1.w ^= v1.w;
        }
        uint4* d = reinterpret_cast<uint4*>(o_sigma + (size_t)o * kMWords);
        d[lane] = a0;
        d[lane + 32] = a1;
        if (lane == 0) {
            o_lid[o] = (uint32_t)((kk >> 10) & 0x3FFFFFFFu);
            o_idx[o] = (uint16_t)((kk >> 1) & 0x1FF);
            o_ch[o] = (uint8_t)(kk & 1);
            o_w[o] = wsum[j];
        }
    }
}

// Applies guard_budget to every ciphertext of *pb (pre- or post-compact_layers: the order does not matter for what this
// does). Replaces *pb by the rebuilt batch if any ciphertext was over budget.
int guard_budget_batch(Ctx* ctx, Batch** pb, uint32_t budget) {
    Batch* b = *pb;
    const uint64_t n = b->n;
    if (n == 0) return PV_OK;
    if (n >= (1ull << 24)) { ctx->last_error = "guard_budget: batch too large"; return PV_E_SHAPE; }
    if (b->nL >= (1ull << 30)) { ctx->last_error = "compact_edges: layer ids need more than the 30 bits of the sort key"; return PV_E_SHAPE; }
    int rc;
    std::vector<void*> scratch;
    auto cleanup = [&]() { for (void* p : scratch) dev_free(ctx, p); scratch.clear(); };
#define CMP_ALLOC(ptr, bytes)                                                      \
    do {                                                                           \
        if ((rc = dev_alloc(ctx, (void**)&(ptr), (bytes)))) { cleanup(); return rc; } \
        scratch.push_back((void*)(ptr));                                           \
    } while (0)
    uint8_t* flag; uint32_t *fcnt, *foff;
    CMP_ALLOC(flag, n); CMP_ALLOC(fcnt, n * 4); CMP_ALLOC(foff, (n + 1) * 4);
    cmp_flag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, b->eoff, budget, flag, fcnt);
    if ((rc = scan_u32(ctx, n, fcnt, foff))) { cleanup(); return rc; }
    uint32_t M = 0;
    { SmallRead sr; sr.add(&M, foff + n, 4); if ((rc = read_small_sync(ctx, sr))) { cleanup(); return rc; } }
    ctx->stat_kernel_launches += 1;
    if (M == 0) { cleanup(); return PV_OK; }

    uint64_t *key, *key2; uint32_t *src, *src2, *keep, *kpos, *zero_list, *zero_cnt, *cnt, *neoff; Fp* wsum;
    CMP_ALLOC(key, (size_t)M * 8); CMP_ALLOC(key2, (size_t)M * 8); CMP_ALLOC(src, (size_t)M * 4); CMP_ALLOC(src2, (size_t)M * 4);
    CMP_ALLOC(keep, (size_t)M * 4); CMP_ALLOC(kpos, (size_t)M * 4); CMP_ALLOC(zero_list, (size_t)M * 4); CMP_ALLOC(zero_cnt, 4);
    CMP_ALLOC(wsum, (size_t)M * 16); CMP_ALLOC(cnt, n * 4); CMP_ALLOC(neoff, (n + 1) * 4);
    const dim3 wide((unsigned)n, 128);     // (ciphertext, chunk): an over-budget ciphertext is > 1.2 GB of sigma
    cmp_keys_kernel<<<wide, 256, 0, ctx->stream>>>(b->eoff, flag, foff, b->lid, b->idx, b->ch, key, src);
    int ibits = 1;
    while ((1ull << ibits) < n) ibits++;
    size_t tmp_bytes = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key2, src, src2, (int)M, 0, 40 + ibits, ctx->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, keep, kpos, (int)M, ctx->stream);
    if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
    void* tmp;
    CMP_ALLOC(tmp, tmp_bytes);
    PV_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key2, src, src2, (int)M, 0, 40 + ibits, ctx->stream));
    PV_CUDA(cudaMemsetAsync(zero_cnt, 0, 4, ctx->stream));
    const unsigned mb = (M + 255) / 256;
    cmp_heads_kernel<<<mb, 256, 0, ctx->stream>>>(M, key2, src2, b->w, keep, wsum, zero_list, zero_cnt);
    // zero-weight runs are all but impossible (p = 2^-127 each); the grid covers the worst case and exits on the counter
    cmp_zero_runs_kernel<<<(unsigned)(((uint64_t)M * 32 + 255) / 256), 256, 0, ctx->stream>>>(zero_cnt, zero_list, M, key2, src2, b->sigma, keep);
    PV_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, keep, kpos, (int)M, ctx->stream));
    uint32_t last[2] = {0, 0};
    { SmallRead sr; sr.add(&last[0], kpos + (M - 1), 4); sr.add(&last[1], keep + (M - 1), 4); if ((rc = read_small_sync(ctx, sr))) { cleanup(); return rc; } }
    const uint32_t kept_total = last[0] + last[1];
    cmp_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, b->eoff, flag, foff, kpos, M, kept_total, cnt);
    if ((rc = scan_u32(ctx, n, cnt, neoff))) { cleanup(); return rc; }
    const uint64_t nE_new = b->nE - M + kept_total;
    Batch* o = nullptr;
    if ((rc = batch_alloc(ctx, n, b->nL, nE_new, &o))) { cleanup(); return rc; }
    PV_CUDA(cudaMemcpyAsync(o->loff, b->loff, (n + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(o->eoff, neoff, (n + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(o->rule, b->rule, b->nL, cudaMemcpyDeviceToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(o->ztag, b->ztag, b->nL * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(o->nlo, b->nlo, b->nL * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(o->nhi, b->nhi, b->nL * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(o->pa, b->pa, b->nL * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    PV_CUDA(cudaMemcpyAsync(o->pb, b->pb, b->nL * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    cmp_write_kernel<<<wide, 256, 0, ctx->stream>>>(b->eoff, neoff, flag, foff, kpos, keep, M, key2, src2, wsum, b->lid, b->idx, b->ch, b->w, b->sigma,
                                                           o->lid, o->idx, o->ch, o->w, o->sigma);
    PV_CUDA(cudaGetLastError());
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stat_kernel_launches += 8;
    cleanup();
    batch_free(b);
    *pb = o;
    return PV_OK;
#undef CMP_ALLOC
}

}  // namespace pvacb
