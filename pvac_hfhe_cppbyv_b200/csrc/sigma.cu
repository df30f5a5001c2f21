// sigma_from_H (crypto/matrix.hpp:267-303) for arrays of edges: the hot kernel of ct_mul and 40 calls per enc_value.
//
//   sigma_cand_kernel    SHA-256 counter PRG of prg_choose_k (matrix.hpp:15-92): one thread per (source, label, ctr),
//                        block 0 of the 78/79-byte message is constant per (source, label) -> midstate in shared memory,
//                        one compression per thread, 4 candidates per hash. 34 hashes (136 candidates) per label cover
//                        the 128 picks plus up to 8 duplicates; the gather kernel continues the stream if that is short.
//   sigma_gather_kernel  one warp per destination edge: ordered de-duplication with a bitmap in shared memory
//                        (first 128 distinct values in stream order = what the unordered_set loop accepts), XOR-gather of
//                        the 128 chosen 1 KiB columns of H (16 MiB, L2 resident) with 128-bit loads, noise flips taken
//                        straight from the noise bitmap, 1 KiB coalesced store.
#include "engine.h"
#include "sha256.cuh"

namespace pvacb {

constexpr int kCandHashes = 34;
constexpr int kCandPerLabel = kCandHashes * 4;  // 136
constexpr int kCandSrcPerCta = 4;
constexpr int kCandThreads = kCandSrcPerCta * 2 * kCandHashes;  // 272

// candidate value of one PRG word: accept iff x <= 2^64-1 - ((2^64-1) % N) (matrix.hpp:64-75; N is a power of two here)
PV_HD uint16_t cand_from_word(uint64_t x, uint32_t N) {
    uint64_t lim = 0ull - (uint64_t)N;  // 2^64 - N
    return x <= lim ? (uint16_t)(x & (N - 1)) : (uint16_t)0xFFFF;
}

// the 7 words hashed by both prg_choose_k calls of sigma_from_H (crypto/matrix.hpp:278-286)
__device__ __forceinline__ void load_words(const SigmaJobs& J, uint64_t canon, uint64_t job, uint64_t x[8]) {
    uint32_t s = J.seed_idx ? J.seed_idx[job] : (uint32_t)job;
    x[0] = canon;
    x[1] = J.ztag[s];
    x[2] = J.nlo[s];
    x[3] = J.nhi[s];
    x[4] = J.idx[job];
    x[5] = J.ch[job];
    x[6] = J.salt[job];
    x[7] = 0;
}

__global__ void __launch_bounds__(kCandThreads)
sigma_cand_kernel(SigmaJobs J, uint64_t canon, uint16_t* __restrict__ cand) {
    __shared__ uint32_t mid[kCandSrcPerCta * 2][8];
    __shared__ uint64_t tailw[kCandSrcPerCta * 2];   // x[6] (salt), needed for block 1
    const int tid = threadIdx.x;
    const int sl = tid / kCandHashes;        // 0..7 : (source local, label)
    const int ctr = tid % kCandHashes;
    const int label = sl & 1;
    const uint64_t src = (uint64_t)blockIdx.x * kCandSrcPerCta + (sl >> 1);
    const bool valid = src < J.n;
    const LabelStream ls = label ? label_noise() : label_xseed();
    if (ctr == 0 && valid) {
        uint64_t x[8];
        load_words(J, canon, src, x);
        uint64_t q[8];
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = stream_word(ls, x, 8, j, 0x80ull);
        uint32_t w[16];
        sha_block_from_le64(q, w);
        ShaState st;
        sha_init(st);
        sha_compress(st, w);
#pragma unroll
        for (int i = 0; i < 8; i++) mid[sl][i] = st.h[i];
        tailw[sl] = x[6];
    }
    __syncthreads();
    if (!valid) return;
    // block 1: remaining bytes of the salt, LE64(ctr), 0x80, zeros, bit length
    const int sh = 8 * ls.r;
    uint64_t salt = tailw[sl];
    uint64_t c = (uint64_t)ctr;
    uint64_t q8 = (salt >> (64 - sh)) | (c << sh);
    uint64_t q9 = (c >> (64 - sh)) | (0x80ull << sh);
    uint32_t w[16];
    w[0] = sha_bswap((uint32_t)q8); w[1] = sha_bswap((uint32_t)(q8 >> 32));
    w[2] = sha_bswap((uint32_t)q9); w[3] = sha_bswap((uint32_t)(q9 >> 32));
#pragma unroll
    for (int i = 4; i < 15; i++) w[i] = 0;
    w[15] = (uint32_t)(8 + ls.r + 64) * 8;  // message bytes = label + 8 words
    ShaState st;
#pragma unroll
    for (int i = 0; i < 8; i++) st.h[i] = mid[sl][i];
    sha_compress(st, w);
    const uint32_t N = label ? (uint32_t)kMBits : (uint32_t)kNBits;
    uint16_t v0 = cand_from_word(sha_digest_le64(st, 0), N), v1 = cand_from_word(sha_digest_le64(st, 1), N);
    uint16_t v2 = cand_from_word(sha_digest_le64(st, 2), N), v3 = cand_from_word(sha_digest_le64(st, 3), N);
    uint2 pk;
    pk.x = (uint32_t)v0 | ((uint32_t)v1 << 16);
    pk.y = (uint32_t)v2 | ((uint32_t)v3 << 16);
    reinterpret_cast<uint2*>(cand + (src * 2 + label) * kCandPerLabel)[ctr] = pk;
}

// slow path: hash number `ctr` of the stream, both blocks (used only when 136 candidates were not enough)
__device__ __noinline__ void prg_hash_words(const LabelStream ls, const uint64_t xin[8], uint64_t ctr, uint64_t out[4]) {
    uint64_t x[8];
    for (int i = 0; i < 7; i++) x[i] = xin[i];
    x[7] = ctr;
    ShaState st;
    sha_label_words(ls, x, 8, st);
    for (int k = 0; k < 4; k++) out[k] = sha_digest_le64(st, k);
}

constexpr int kGatherWarps = 8;
struct GatherSmem {
    uint32_t bmX[kNBits / 32];     // 2 KiB: seen columns
    uint32_t bmN[kMBits / 32];     // 1 KiB: noise bits (doubles as the flip mask)
    uint16_t cols[kXColWt];        // chosen columns
    uint16_t more[128];            // continuation candidates
};

// ordered de-duplication of one label. Lanes hold candidates 4*lane..4*lane+3 (c[]), positions 128..135 are in gc[128..].
// Returns with exactly `want` distinct values marked in bm; winners are appended to cols (if cols != nullptr).
__device__ __forceinline__ void dedupe_label(uint32_t* bm, uint16_t* cols, uint16_t* more, const uint16_t* __restrict__ gc, const LabelStream ls,
                                             const uint64_t x[8], uint32_t N, int lane) {
    uint2 pk = reinterpret_cast<const uint2*>(gc)[lane];
    uint16_t c[4] = {(uint16_t)(pk.x & 0xffff), (uint16_t)(pk.x >> 16), (uint16_t)(pk.y & 0xffff), (uint16_t)(pk.y >> 16)};
    int have = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        bool win = false;
        if (c[k] != 0xFFFF) {
            uint32_t bit = 1u << (c[k] & 31);
            uint32_t old = atomicOr(&bm[c[k] >> 5], bit);
            win = !(old & bit);
        }
        uint32_t b = __ballot_sync(0xffffffffu, win);
        if (win && cols) cols[have + __popc(b & ((1u << lane) - 1))] = c[k];
        have += __popc(b);
    }
    if (have < kXColWt) {   // both labels want 128 picks (x_col_wt = err_wt = 128)
        // sequential tail, exactly like the reference's one-word-at-a-time loop
        int pos = 128;
        uint64_t next_ctr = kCandHashes;
        const uint16_t* cur = gc;
        int cur_base = 0, cur_end = kCandPerLabel;
        while (have < kXColWt) {
            if (pos >= cur_end) {   // continue the PRG stream: 32 more hashes, one per lane
                uint64_t o[4];
                prg_hash_words(ls, x, next_ctr + lane, o);
                for (int k = 0; k < 4; k++) more[4 * lane + k] = cand_from_word(o[k], N);
                __syncwarp();
                cur = more; cur_base = pos; cur_end = pos + 128; next_ctr += 32;
            }
            if (lane == 0) {
                while (pos < cur_end && have < kXColWt) {
                    uint16_t v = cur[pos - cur_base];
                    pos++;
                    if (v == 0xFFFF) continue;
                    uint32_t bit = 1u << (v & 31);
                    uint32_t old = bm[v >> 5];
                    if (!(old & bit)) {
                        bm[v >> 5] = old | bit;
                        if (cols) cols[have] = v;
                        have++;
                    }
                }
            }
            have = __shfl_sync(0xffffffffu, have, 0);
            pos = __shfl_sync(0xffffffffu, pos, 0);
            __syncwarp();
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kGatherWarps * 32)
sigma_gather_kernel(SigmaJobs J, uint64_t canon, const uint16_t* __restrict__ cand, const uint4* __restrict__ H4) {
    __shared__ GatherSmem sm[kGatherWarps];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    GatherSmem& S = sm[wid];
    for (int i = lane; i < kNBits / 32; i += 32) S.bmX[i] = 0;
    for (int i = lane; i < kMBits / 32; i += 32) S.bmN[i] = 0;
    __syncwarp();
    const uint64_t nwarps = (uint64_t)gridDim.x * kGatherWarps;
    for (uint64_t job = (uint64_t)blockIdx.x * kGatherWarps + wid; job < J.n; job += nwarps) {
        uint4 a0 = make_uint4(0, 0, 0, 0), a1 = make_uint4(0, 0, 0, 0);
        uint64_t x[8];
        load_words(J, canon, job, x);
        const uint16_t* gc = cand + job * 2 * kCandPerLabel;
        // ---- columns of H
        dedupe_label(S.bmX, S.cols, S.more, gc, label_xseed(), x, (uint32_t)kNBits, lane);
#pragma unroll 1
        for (int i = 0; i < kXColWt; i += 8) {
            uint4 cv = *reinterpret_cast<const uint4*>(&S.cols[i]);   // 8 column ids, broadcast
            uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w};
            uint4 v[16];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t col = (cw[k >> 1] >> ((k & 1) * 16)) & 0xffff;
                const uint4* p = H4 + (size_t)col * 64 + lane;
                v[2 * k] = __ldcg(p);
                v[2 * k + 1] = __ldcg(p + 32);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                a0.x ^= v[2 * k].x; a0.y ^= v[2 * k].y; a0.z ^= v[2 * k].z; a0.w ^= v[2 * k].w;
                a1.x ^= v[2 * k + 1].x; a1.y ^= v[2 * k + 1].y; a1.z ^= v[2 * k + 1].z; a1.w ^= v[2 * k + 1].w;
            }
        }
        // clear the column bitmap again: every word that got a bit belongs to one of the chosen columns
        for (int i = lane; i < kXColWt; i += 32) S.bmX[S.cols[i] >> 5] = 0;
        __syncwarp();
        // ---- noise bits: the de-dup bitmap is the flip mask
        dedupe_label(S.bmN, nullptr, S.more, gc + kCandPerLabel, label_noise(), x, (uint32_t)kMBits, lane);
        uint4* bn = reinterpret_cast<uint4*>(S.bmN);
        uint4 n0 = bn[lane], n1 = bn[lane + 32];
        a0.x ^= n0.x; a0.y ^= n0.y; a0.z ^= n0.z; a0.w ^= n0.w;
        a1.x ^= n1.x; a1.y ^= n1.y; a1.z ^= n1.z; a1.w ^= n1.w;
        bn[lane] = make_uint4(0, 0, 0, 0);
        bn[lane + 32] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        uint64_t row = J.out_row ? J.out_row[job] : job;
        uint4* o = reinterpret_cast<uint4*>(row < J.out_split ? J.out + row * kMWords : J.out2 + (row - J.out_split) * kMWords);
        o[lane] = a0;
        o[lane + 32] = a1;
    }
}

// out[dst] ^= scratch row, for merged edges of enc_value (compact_edges XORs the sigmas, ops/encrypt.hpp:39-71).
// Several pairs may share a destination, hence atomics.
__global__ void sigma_xor_rows_kernel(uint64_t npairs, const uint2* __restrict__ pairs, uint64_t* __restrict__ out, uint64_t split,
                                      const uint64_t* __restrict__ out2) {
    const int lane = threadIdx.x & 31;
    uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= npairs) return;
    uint2 pr = pairs[p];
    unsigned long long* d = reinterpret_cast<unsigned long long*>(out + (uint64_t)pr.x * kMWords);
    const uint64_t* s = (uint64_t)pr.y < split ? out + (uint64_t)pr.y * kMWords : out2 + ((uint64_t)pr.y - split) * kMWords;
    for (int k = lane; k < kMWords; k += 32) atomicXor(d + k, (unsigned long long)s[k]);
}

int sigma_xor_rows(Ctx* ctx, uint64_t npairs, const uint2* d_pairs, uint64_t* out, uint64_t split, const uint64_t* out2) {
    if (!npairs) return PV_OK;
    sigma_xor_rows_kernel<<<(unsigned)((npairs * 32 + 255) / 256), 256, 0, ctx->stream>>>(npairs, d_pairs, out, split, out2);
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    return PV_OK;
}

int sigma_run(Ctx* ctx, const SigmaJobs& J) {
    if (J.n == 0) return PV_OK;
    uint16_t* cand = nullptr;
    int rc;
    // candidates are produced in chunks to bound scratch memory (544 B per job)
    const uint64_t chunk = 1ull << 22;
    uint64_t max_n = J.n < chunk ? J.n : chunk;
    if ((rc = dev_alloc(ctx, (void**)&cand, max_n * 2 * kCandPerLabel * 2))) return rc;
    for (uint64_t s0 = 0; s0 < J.n; s0 += chunk) {
        uint64_t ns = J.n - s0 < chunk ? J.n - s0 : chunk;
        SigmaJobs S = J;
        S.n = ns;
        if (S.seed_idx) S.seed_idx += s0; else { S.ztag += s0; S.nlo += s0; S.nhi += s0; }
        S.idx += s0; S.ch += s0; S.salt += s0;
        if (S.out_row) S.out_row += s0; else S.out += s0 * kMWords;
        {
            ProfScope ps(ctx, PROF_SIGMA_CAND);
            sigma_cand_kernel<<<(unsigned)((ns + kCandSrcPerCta - 1) / kCandSrcPerCta), kCandThreads, 0, ctx->stream>>>(S, ctx->kv.canon_tag, cand);
        }
        unsigned grid = (unsigned)((ns + kGatherWarps - 1) / kGatherWarps);
        unsigned cap = (unsigned)ctx->sm_count * 8;
        if (grid > cap) grid = cap;
        {
            ProfScope ps(ctx, PROF_SIGMA_GATHER);
            sigma_gather_kernel<<<grid, kGatherWarps * 32, 0, ctx->stream>>>(S, ctx->kv.canon_tag, cand, reinterpret_cast<const uint4*>(ctx->kv.H));
        }
        PV_CUDA(cudaGetLastError());
        ctx->stat_kernel_launches += 2;
    }
    ctx->stat_sigma_edges += J.n;
    dev_free(ctx, cand);
    return PV_OK;
}

}  // namespace pvacb
