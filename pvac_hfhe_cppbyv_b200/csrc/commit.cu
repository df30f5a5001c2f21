// commit_ct (ops/commit.hpp:12-87) for a batch: one SHA-256 per ciphertext over
//   "pvac.dom.commit" || H_digest || LE64(canon_tag) || layers (rule byte + 3 or 2 u64) || edges (LE64(layer_id), LE64(idx), ch byte,
//   w as 16 bytes with bit 127 cleared, sigma as 1 024 bytes).
// The only public consumer of sigma: 1 057 bytes hashed per edge (16.5 compressions), so it is bound by the SHA-256 ALU rate
// (15 G compressions/s = 0.97 TB/s of input), not by HBM. A hash chain is sequential, so the parallelism is across
// ciphertexts: one thread per ciphertext. A chain is also a long run of DEPENDENT instructions, so a scheduler needs several warps
// to stay busy, while every warp instruction costs an issue slot however few lanes it carries: the launcher picks the smallest
// lanes_per_warp that keeps the warp count within what the schedulers can issue (commit.cu: op_commit_ct). The message is a byte stream with 1-byte fields in it, so u64 items are appended
// through a byte-granular shift register; the 64-byte block being filled lives in shared memory (dynamic index), the
// sigma rows are read 32 bytes (one sector) at a time.
#include "engine.h"
#include "sha256.cuh"

namespace pvacb {

constexpr int kCommitThreads = 128;

struct CommitStream {
    uint32_t h[8];
    uint64_t acc;      // pending bytes (little-endian), nacc of them
    uint32_t nacc;     // 0..7
    uint32_t nw;       // stream words already in the current block, 0..7
    uint64_t* blk;     // &block[0][tid], stride kCommitThreads
    uint32_t one;      // 1 at run time: the round additions go to the FMA pipe as a * 1 + b (sha256.cuh)

    // ONE copy of the compression in the kernel (noinline): inlined at every put_u64 / put_u8 site it made 285 KB of SASS
    __device__ __noinline__ void flush() {
        {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint64_t q = blk[j * kCommitThreads];
                w[2 * j] = sha_bswap((uint32_t)q);
                w[2 * j + 1] = sha_bswap((uint32_t)(q >> 32));
            }
            // the rolled form (4 trips of 16 rounds): the fully unrolled compression is 40 KB of SASS, more than the 32 KB
            // instruction cache -- ncu r02: 6.4 no_instruction stalls per issue, 18 % issue slots used
            sha_compress_from_rolled4<false>(h, w, h, one);
            nw = 0;
        }
    }
    __device__ __forceinline__ void emit(uint64_t word) {
        blk[nw * kCommitThreads] = word;
        if (++nw == 8) flush();
    }
    __device__ __forceinline__ void put_u64(uint64_t x) {
        if (nacc == 0) { emit(x); return; }
        const uint32_t s = 8 * nacc;
        emit(acc | (x << s));
        acc = x >> (64 - s);
    }
    __device__ __forceinline__ void put_u8(uint8_t b) {
        acc |= (uint64_t)b << (8 * nacc);
        if (++nacc == 8) { emit(acc); acc = 0; nacc = 0; }
    }
};

__global__ void __launch_bounds__(kCommitThreads)
commit_kernel(uint64_t n, uint64_t canon_tag, const uint64_t* __restrict__ hdig /*4 words*/, const uint32_t* __restrict__ loff,
              const uint32_t* __restrict__ eoff, const uint8_t* __restrict__ rule, const uint64_t* __restrict__ ztag, const uint64_t* __restrict__ nlo,
              const uint64_t* __restrict__ nhi, const uint32_t* __restrict__ pa, const uint32_t* __restrict__ pb, const uint32_t* __restrict__ lid,
              const uint16_t* __restrict__ idx, const uint8_t* __restrict__ ch, const Fp* __restrict__ w, const uint64_t* __restrict__ sigma,
              uint32_t* __restrict__ out /* n x 8 words = the digest bytes */, int lanes_per_warp, uint32_t one) {
    __shared__ uint64_t block[8][kCommitThreads];
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * kCommitThreads + threadIdx.x) >> 5;
    const uint64_t i = warp * (uint64_t)lanes_per_warp + lane;
    if (lane >= (uint32_t)lanes_per_warp || i >= n) return;
    CommitStream S;
    ShaState iv;
    sha_init(iv);
#pragma unroll
    for (int k = 0; k < 8; k++) S.h[k] = iv.h[k];
    S.blk = &block[0][threadIdx.x];
    S.one = one;
    S.nw = 0;
    // "pvac.dom.commit": 15 bytes = one full word + 7 pending bytes
    S.acc = 0; S.nacc = 0;
    S.emit(pack_label("pvac.dom.commit", 0, 8));
    S.acc = pack_label("pvac.dom.commit", 8, 15);
    S.nacc = 7;
    for (int k = 0; k < 4; k++) S.put_u64(hdig[k]);
    S.put_u64(canon_tag);
    uint64_t bytes = 15 + 32 + 8;
    const uint32_t l0 = loff[i], l1 = loff[i + 1], e0 = eoff[i], e1 = eoff[i + 1];
    for (uint32_t l = l0; l < l1; l++) {
        const uint8_t r = rule[l];
        S.put_u8(r);
        if (r == 0) { S.put_u64(ztag[l]); S.put_u64(nlo[l]); S.put_u64(nhi[l]); bytes += 25; }
        else { S.put_u64(pa[l]); S.put_u64(pb[l]); bytes += 17; }
    }
    for (uint32_t e = e0; e < e1; e++) {
        S.put_u64(lid[e]);
        S.put_u64(idx[e]);
        S.put_u8(ch[e]);
        const Fp we = w[e];
        S.put_u64(we.lo);
        S.put_u64(we.hi & kMask63);
        const ulonglong2* row = reinterpret_cast<const ulonglong2*>(sigma + (size_t)e * kMWords);
#pragma unroll 1
        for (int k = 0; k < kMWords / 4; k++) {
            ulonglong2 a = __ldcs(row + 2 * k), b = __ldcs(row + 2 * k + 1);   // one 32-byte sector, read once
            S.put_u64(a.x); S.put_u64(a.y); S.put_u64(b.x); S.put_u64(b.y);
        }
    }
    bytes += (uint64_t)(e1 - e0) * (8 + 8 + 1 + 16 + kMWords * 8);
    // FIPS 180-4 padding: 0x80, zeros up to 56 mod 64, bit length big-endian
    S.put_u8(0x80);
    while ((S.nw * 8 + S.nacc) != 56) S.put_u8(0);
    const uint64_t bits = bytes * 8;
    S.put_u64(((uint64_t)sha_bswap((uint32_t)bits) << 32) | sha_bswap((uint32_t)(bits >> 32)));   // LE bytes of bswap64 = BE bytes of bits
#pragma unroll
    for (int k = 0; k < 8; k++) out[i * 8 + k] = sha_bswap(S.h[k]);      // word k of the output holds digest bytes 4k..4k+3 in memory order
}

int op_commit_ct(Ctx* ctx, const Batch* b, uint8_t* h_out /* n x 32 */) {
    if (b->n == 0) return PV_OK;
    uint32_t* d_out = nullptr;
    uint64_t* d_dig = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, (void**)&d_out, b->n * 32))) return rc;
    if ((rc = dev_alloc(ctx, (void**)&d_dig, 32))) { dev_free(ctx, d_out); return rc; }
    PV_CUDA(cudaMemcpyAsync(d_dig, ctx->d_blob + 1, 32, cudaMemcpyDeviceToDevice, ctx->stream));   // H_digest = blob words 1..4
    // a chain issues about 0.4 instructions per clock (dependent rounds): up to ~2.5 warps keep one scheduler busy. Fewer ciphertexts
    // than that many full warps -> fewer lanes per warp, so that no chain waits for an issue slot; more -> full warps (throughput).
    const uint64_t busy_warps = (uint64_t)ctx->sm_count * 4 * 5 / 2;
    int lpw = 1;
    while (lpw < 32 && (b->n + lpw - 1) / lpw > busy_warps) lpw <<= 1;
    const uint64_t warps = (b->n + lpw - 1) / lpw;
    {
        ProfScope ps(ctx, PROF_COMMIT);
        commit_kernel<<<(unsigned)((warps * 32 + kCommitThreads - 1) / kCommitThreads), kCommitThreads, 0, ctx->stream>>>(
            b->n, ctx->kv.canon_tag, d_dig, b->loff, b->eoff, b->rule, b->ztag, b->nlo, b->nhi, b->pa, b->pb, b->lid, b->idx, b->ch, b->w, b->sigma, d_out, lpw, 1u);
    }
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    PV_CUDA(cudaMemcpyAsync(h_out, d_out, b->n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, d_out);
    dev_free(ctx, d_dig);
    return PV_OK;
}

}  // namespace pvacb
