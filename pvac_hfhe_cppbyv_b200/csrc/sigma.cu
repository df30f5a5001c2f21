// sigma_from_H (crypto/matrix.hpp:267-303) for arrays of edges: the hot kernel of ct_mul and 40 calls per enc_value.
//
// ONE fused kernel, one warp per group of G edges, no CTA-wide barrier. Each warp runs three phases per group:
//   A  midstates   prg_choose_k (matrix.hpp:15-92) hashes  label || 7 words || LE64(ctr): block 0 of that 78/79-byte message is
//                  constant per (edge, label), so 2G lanes compress it once and park the SHA-256 midstate in shared memory;
//   B  candidates  the G*68 counter hashes (34 per label = 136 candidates: the 128 picks plus up to 8 duplicates) are dealt
//                  round-robin to the 32 lanes, one compression each (block 1 = tail of the salt, ctr, padding); the accepted
//                  16-bit candidates go to shared memory. G = 8 makes that exactly 17 full rounds;
//   C  gather      per edge: ordered de-duplication with a bitmap in shared memory (first 128 distinct values in stream order
//                  = what the reference's unordered_set loop accepts, continuing the stream in the rare short case), XOR-
//                  gather of the 128 chosen 1 KiB columns of H (16 MiB, L2 resident) with 128-bit loads, noise flips taken
//                  straight from the noise bitmap, one coalesced 1 KiB store.
// Phases A/B are ALU-pipe work (SHF/LOP3/IADD3), phase C is L2->SM bandwidth; warps of one SM drift apart, so the hashing
// of some warps hides under the gathers of the others. (r01: as two kernels, candidates through HBM, the same work took
// 47.7 ms + 31.9 ms per 4.96 M edges; half of the first was a one-lane-per-warp midstate phase.)
#include "engine.h"
#include "sha256.cuh"

#include <cstdlib>

namespace pvacb {

constexpr int kCandHashes = 34;
constexpr int kCandPerLabel = kCandHashes * 4;  // 136

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

// slow path: hash number `ctr` of the stream, both blocks (used only when 136 candidates were not enough)
__device__ __noinline__ void prg_hash_words(const LabelStream ls, const uint64_t xin[8], uint64_t ctr, uint64_t out[4]) {
    uint64_t x[8];
    for (int i = 0; i < 7; i++) x[i] = xin[i];
    x[7] = ctr;
    ShaState st;
    sha_label_words(ls, x, 8, st);
    for (int k = 0; k < 4; k++) out[k] = sha_digest_le64(st, k);
}

template <int G, int NH = kCandHashes>
struct __align__(16) SigmaWarpSmem {
    uint32_t bm[kNBits / 32];               // 2 KiB: de-dup bitmap of the columns, then of the noise bits (= the flip mask)
    uint32_t cols[kXColWt + 4];             // chosen columns in draw order, as byte offsets into H (col * 1024); 4 entries of padding
    uint16_t cand[G * 2 * NH * 4];          // phase B output: [edge][label][4 NH]  (NH = 34: 136 candidates)
    union {
        uint32_t mid[G * 2][8];             // phase A output: SHA-256 state after block 0, per (edge, label)
        uint16_t more[128];                 // phase C, rare: continuation candidates
    };
    uint64_t salt[G < 4 ? 4 : G];           // word 6 of the hashed words: its tail opens block 1
};

// ordered de-duplication of one label. Lanes hold candidates 4*lane..4*lane+3, positions 128..135 follow in gc[128..].
// Returns with exactly 128 distinct values marked in bm (x_col_wt = err_wt = 128); winners are appended to cols if given.
__device__ __forceinline__ void dedupe_label(uint32_t* bm, uint32_t* cols, uint16_t* more, const uint16_t* gc, const LabelStream ls,
                                             const SigmaJobs& J, uint64_t canon, uint64_t job, uint32_t N, int lane, int n_hashes = kCandHashes) {
    uint2 pk = reinterpret_cast<const uint2*>(gc)[lane];
    uint16_t c[4] = {(uint16_t)(pk.x & 0xffff), (uint16_t)(pk.x >> 16), (uint16_t)(pk.y & 0xffff), (uint16_t)(pk.y >> 16)};
    // the four atomics go out back to back (their results are only needed by the ballots below): one shared-memory round
    // trip per label instead of four. Which of two equal candidates "wins" does not matter, only the set of values does.
    uint32_t bit[4], old[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        bit[k] = 1u << (c[k] & 31);
        old[k] = 0xFFFFFFFFu;
        if (c[k] != 0xFFFF) old[k] = atomicOr(&bm[c[k] >> 5], bit[k]);
    }
    int have = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool win = !(old[k] & bit[k]);
        uint32_t b = __ballot_sync(0xffffffffu, win);
        if (win && cols) cols[have + __popc(b & ((1u << lane) - 1))] = (uint32_t)c[k] * 1024u;
        have += __popc(b);
    }
    if (have < kXColWt) {
        // sequential tail, exactly like the reference's one-word-at-a-time loop
        int pos = 128;
        uint64_t next_ctr = (uint64_t)n_hashes;
        const uint16_t* cur = gc;
        int cur_base = 0, cur_end = 4 * n_hashes;
        uint64_t x[8];
        bool have_x = false;
        while (have < kXColWt) {
            if (pos >= cur_end) {   // continue the PRG stream: 32 more hashes, one per lane
                if (!have_x) { load_words(J, canon, job, x); have_x = true; }
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
                        if (cols) cols[have] = (uint32_t)v * 1024u;
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

// address of this lane's 16 bytes of a column = Hl + byte offset, formed on the FMA pipe: the low word as off * one + lo(Hl)
// (`one` is a run-time 1), the high word is that of Hl -- the key blob never straddles a 4 GiB boundary (engine.cu: ensure_blob).
// Written as pointer arithmetic it costs IADD3 + LEA + LEA.HI.X per column on the ALU pipe, which is what bounds the kernel.
__device__ __forceinline__ const uint4* col_ptr(const uint4* Hl, uint32_t byte_off, uint32_t one) {
    const uint64_t b = reinterpret_cast<uint64_t>(Hl);
    uint32_t lo;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(lo) : "r"(byte_off), "r"(one), "r"((uint32_t)b));
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"((uint32_t)(b >> 32)));
    return reinterpret_cast<const uint4*>(r);
}

// phase C for one edge: ordered de-duplication of both labels, XOR-gather of the 128 chosen columns, noise flips, store.
// gc = the edge's candidates [2][136] in shared memory; bm / cols / more = the warp's scratch.
template <int EXP, int CPS, int NH>
__device__ __forceinline__ void sigma_edge(uint32_t* bm, uint32_t* cols, uint16_t* more, const uint16_t* gc, const SigmaJobs& J, uint64_t canon, uint64_t job,
                                           const uint4* Hl, int lane, uint32_t one) {
    dedupe_label(bm, cols, more, gc, label_xseed(), J, canon, job, (uint32_t)kNBits, lane, NH);
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = make_uint4(0, 0, 0, 0);
    // 8 column offsets per step, broadcast reads. The NEXT step's offsets are fetched right after this step's loads
    // have been issued: shared-memory round trips are slow while the gathers saturate L1TEX (ncu: short-scoreboard
    // and mio-throttle stalls), so they must not sit between two batches of global loads.
    if (CPS == 8) {
        uint4 c0 = *reinterpret_cast<const uint4*>(&cols[0]);
        uint4 c1 = *reinterpret_cast<const uint4*>(&cols[4]);
#pragma unroll 1
        for (int i = 0; i < kXColWt; i += 8) {
            const uint32_t co[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            uint4 v[16];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint4* p = col_ptr(Hl, co[k], one);
                if (EXP == 1) { v[2 * k] = make_uint4(co[k], i, k, lane); v[2 * k + 1] = make_uint4(k, co[k], lane, i); continue; }
                v[2 * k] = __ldcg(p);
                v[2 * k + 1] = __ldcg(p + 32);
            }
            if (i == 0) {
                // clear the bitmap again (every word that got a bit belongs to one of the chosen columns) -- under the first loads
                for (int q = lane; q < kXColWt; q += 32) bm[cols[q] >> 15] = 0;
            }
            if (i + 8 < kXColWt) {
                c0 = *reinterpret_cast<const uint4*>(&cols[i + 8]);
                c1 = *reinterpret_cast<const uint4*>(&cols[i + 12]);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                a0.x ^= v[2 * k].x; a0.y ^= v[2 * k].y; a0.z ^= v[2 * k].z; a0.w ^= v[2 * k].w;
                a1.x ^= v[2 * k + 1].x; a1.y ^= v[2 * k + 1].y; a1.z ^= v[2 * k + 1].z; a1.w ^= v[2 * k + 1].w;
            }
        }
    } else {
        // clear the bitmap again (every word that got a bit belongs to one of the chosen columns). The loop below is kept free of
        // everything but loads and XORs: the kernel is bound by the ALU pipe, and a compare per step is 32 ALU instructions per edge.
        for (int q = lane; q < kXColWt; q += 32) bm[cols[q] >> 15] = 0;
        uint4 c0 = *reinterpret_cast<const uint4*>(&cols[0]);
#pragma unroll 4
        for (int i = 0; i < kXColWt; i += 4) {
            const uint32_t co[4] = {c0.x, c0.y, c0.z, c0.w};
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint4* p = col_ptr(Hl, co[k], one);
                if (EXP == 1) { v[2 * k] = make_uint4(co[k], i, k, lane); v[2 * k + 1] = make_uint4(k, co[k], lane, i); continue; }
                v[2 * k] = __ldcg(p);
                v[2 * k + 1] = __ldcg(p + 32);
            }
            c0 = *reinterpret_cast<const uint4*>(&cols[i + 4]);       // the last step reads the 4 padding entries
#pragma unroll
            for (int k = 0; k < 4; k++) {
                a0.x ^= v[2 * k].x; a0.y ^= v[2 * k].y; a0.z ^= v[2 * k].z; a0.w ^= v[2 * k].w;
                a1.x ^= v[2 * k + 1].x; a1.y ^= v[2 * k + 1].y; a1.z ^= v[2 * k + 1].z; a1.w ^= v[2 * k + 1].w;
            }
        }
    }
    __syncwarp();
    // noise bits: the de-dup bitmap is the flip mask (values < 8192: the first KiB of bm)
    dedupe_label(bm, nullptr, more, gc + 4 * NH, label_noise(), J, canon, job, (uint32_t)kMBits, lane, NH);
    uint4* bn = reinterpret_cast<uint4*>(bm);
    uint4 n0 = bn[lane], n1 = bn[lane + 32];
    a0.x ^= n0.x; a0.y ^= n0.y; a0.z ^= n0.z; a0.w ^= n0.w;
    a1.x ^= n1.x; a1.y ^= n1.y; a1.z ^= n1.z; a1.w ^= n1.w;
    bn[lane] = make_uint4(0, 0, 0, 0);
    bn[lane + 32] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    uint64_t row = J.out_row ? J.out_row[job] : job;
    uint4* o = reinterpret_cast<uint4*>(row < J.out_split ? J.out + row * kMWords : J.out2 + (row - J.out_split) * kMWords);
    __stcs(o + lane, a0);        // streaming store: the 1 KiB rows are write-once, keep L2 for H
    __stcs(o + lane + 32, a1);
}

// EXP (tuning experiments only, wrong results): 1 = no gather loads (hash + de-dup only), 2 = no hashing (synthetic candidates)
// NH = counter hashes per label computed up front (34 = 136 candidates: the 128 picks plus 8 spare). NH = 32 leaves no spare, so
// every label with a duplicate takes the in-kernel continuation of the PRG stream: a test shape for that rare path.
// ROLLED: 0 = fully unrolled SHA-256 compressions (45 KB of code in phases A and B: the hashing warps stalled 13-19 % of the time on
// instruction fetch); 1 = rounds 0..15 unrolled, 16..63 as a 3-trip loop (3.8 % faster); 2 = four trips of (16 rounds, next
// schedule): one copy of the round code, the hot code of the kernel is 26 KB, below the 32 KB L1.5 instruction cache (another
// 2 %); 3 = form 2 for the midstates only.
template <int G, int WARPS, int MINB, bool FMA, int EXP = 0, int CPS = 8, int NH = kCandHashes, int ROLLED = 2>
__global__ void __launch_bounds__(WARPS * 32, MINB)
sigma_fused_kernel(SigmaJobs J, uint64_t canon, const uint4* __restrict__ H4, unsigned long long* __restrict__ work, uint32_t one) {
    extern __shared__ __align__(16) uint8_t sigma_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    SigmaWarpSmem<G, NH>& S = reinterpret_cast<SigmaWarpSmem<G, NH>*>(sigma_smem)[wid];
    for (int i = lane; i < kNBits / 32; i += 32) S.bm[i] = 0;
    __syncwarp();
    const uint64_t ngroups = (J.n + G - 1) / G;
    const uint4* const Hl = H4 + lane;
    // persistent warps pull groups from a global counter: gather time varies per SM (L2 distance, neighbours), a static
    // split left a quarter of the warp-time idle at the tail (ncu r01: sm__warps_active 33% of a possible 44%)
    for (;;) {
        unsigned long long grp = 0;
        if (lane == 0) grp = atomicAdd(work, 1ull);
        grp = __shfl_sync(0xffffffffu, grp, 0);
        if (grp >= ngroups) break;
        const uint64_t job0 = grp * G;
        const int ng = (int)(J.n - job0 < (uint64_t)G ? J.n - job0 : (uint64_t)G);
        // ---- phase A: midstate of block 0 for (edge, label) = (lane >> 1, lane & 1)
        if (lane < 2 * ng) {
            const LabelStream ls = (lane & 1) ? label_noise() : label_xseed();
            uint64_t x[8];
            load_words(J, canon, job0 + (lane >> 1), x);
            uint64_t q[8];
#pragma unroll
            for (int j = 0; j < 8; j++) q[j] = stream_word(ls, x, 8, j, 0x80ull);
            uint32_t w[16];
            sha_block_from_le64(q, w);
            if (ROLLED) {
                uint32_t d[8];
                if (ROLLED >= 2) sha_compress_from_rolled4<false>(kShaIv, w, d, one);
                else sha_compress_from_rolled(kShaIv, w, d, one);
#pragma unroll
                for (int i = 0; i < 8; i++) S.mid[lane][i] = d[i];
            } else {
                ShaState st;
                sha_init(st);
                sha_compress(st, w);
#pragma unroll
                for (int i = 0; i < 8; i++) S.mid[lane][i] = st.h[i];
            }
            if (!(lane & 1)) S.salt[lane >> 1] = x[6];
        }
        __syncwarp();
        // ---- phase B: hash h = (edge, label, ctr), one compression: block 1 = tail of the salt, LE64(ctr), 0x80, bit length
        const int nh = ng * 2 * NH;
        for (int h = lane; h < nh; h += 32) {
            const int sl = h / NH;                   // edge * 2 + label
            const uint32_t ctr = (uint32_t)(h - sl * NH);
            const int label = sl & 1;
            const int sh = label ? 48 : 56;          // 8 * (label length - 8)
            const uint64_t q8 = (S.salt[sl >> 1] >> (64 - sh)) | ((uint64_t)ctr << sh);
            uint32_t w[16];
            w[0] = sha_bswap((uint32_t)q8);
            w[1] = sha_bswap((uint32_t)(q8 >> 32));
            w[2] = 0;                                 // ctr < 2^8: the upper counter bytes are zero
            w[3] = label ? 0x00008000u : 0x00000080u; // the 0x80 terminator right after the counter
#pragma unroll
            for (int i = 4; i < 15; i++) w[i] = 0;
            w[15] = label ? 78u * 8u : 79u * 8u;
            uint32_t d[8];
            if (EXP == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) d[i] = (uint32_t)(h * 8 + i) * 2654435761u + w[0];
            } else if (ROLLED >= 2) sha_compress_from_rolled4<true>(S.mid[sl], w, d, one);
            else if (ROLLED) sha_compress_from_rolled(S.mid[sl], w, d, one);
            else if (FMA) sha_compress_from_fma(S.mid[sl], w, d, one);
            else sha_compress_from(S.mid[sl], w, d);
            const uint32_t N = label ? (uint32_t)kMBits : (uint32_t)kNBits;
            // word k of the hash = LE64 of digest bytes 8k..8k+7, i.e. bswap(d[2k]) | bswap(d[2k+1]) << 32. N is a power of two below
            // 2^16: the candidate is the low 16 bits masked (digest bytes 8k, 8k+1 = the top two bytes of d[2k]), and the word is only
            // rejected when it exceeds 2^64 - N, which needs d[2k+1] == 0xFFFFFFFF: that (2^-30 per hash) takes the exact path.
            uint2 pk;
            if (ROLLED != 3 && NH != 32 && max(max(d[1], d[3]), max(d[5], d[7])) != 0xFFFFFFFFu) {     // (the NH = 32 test shape always takes the exact path)
                const uint32_t M = (N - 1) * 0x00010001u;
                pk.x = __byte_perm(d[0], d[2], 0x6723) & M;
                pk.y = __byte_perm(d[4], d[6], 0x6723) & M;
            } else {
                pk.x = (uint32_t)cand_from_word(sha_le64_of(d[0], d[1]), N) | ((uint32_t)cand_from_word(sha_le64_of(d[2], d[3]), N) << 16);
                pk.y = (uint32_t)cand_from_word(sha_le64_of(d[4], d[5]), N) | ((uint32_t)cand_from_word(sha_le64_of(d[6], d[7]), N) << 16);
            }
            reinterpret_cast<uint2*>(S.cand + sl * (4 * NH))[ctr] = pk;
        }
        __syncwarp();
        // ---- phase C: de-duplicate, gather, flip, store
#pragma unroll 1
        for (int e = 0; e < ng; e++) sigma_edge<EXP, CPS, NH>(S.bm, S.cols, S.more, S.cand + e * 2 * (4 * NH), J, canon, job0 + e, Hl, lane, one);
        __syncwarp();
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
    if (pr.x == 0xFFFFFFFFu) return;          // the raw edge's slot was dropped by compact_edges (enc.cu): its row is not merged anywhere
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

template <int G, int WARPS, int MINB, bool FMA = false, int EXP = 0, int CPS = 8, int NH = kCandHashes, int ROLLED = 2>
static int sigma_launch(Ctx* ctx, const SigmaJobs& J) {
    auto kern = sigma_fused_kernel<G, WARPS, MINB, FMA, EXP, CPS, NH, ROLLED>;
    constexpr int smem = (int)sizeof(SigmaWarpSmem<G, NH>) * WARPS;
    const void* kid = reinterpret_cast<const void*>(kern);
    bool attr_done = false;                        // function attributes are per device: remember them per context, not per process
    for (const void* k : ctx->configured_kernels) attr_done |= (k == kid);
    if (!attr_done) {
        PV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        // ask for exactly the shared memory MINB CTAs need, not the maximum: with the 228 KB carve-out (28 KB of L1 left) the L2
        // gather ceiling drops from 19.9 to 16.3 TB/s (pvacb_l2_gather_probe under PVACB_PROBE_CARVEOUT, profiles/r01_notes.md)
        int carve = (int)(((size_t)MINB * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
#ifdef PVACB_TUNING
        if (getenv("PVACB_SIGMA_CARVEOUT")) carve = atoi(getenv("PVACB_SIGMA_CARVEOUT"));
#endif
        PV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve > 100 ? 100 : carve));
        ctx->configured_kernels.push_back(kid);
    }
    const uint64_t ngroups = (J.n + G - 1) / G;
    uint64_t grid = (ngroups + WARPS - 1) / WARPS;
    const uint64_t cap = (uint64_t)ctx->sm_count * MINB;     // persistent: every resident warp walks groups round-robin
    if (grid > cap) grid = cap;
    PV_CUDA(cudaMemsetAsync(ctx->d_work, 0, 8, ctx->stream));
    ProfScope ps(ctx, PROF_SIGMA);
    kern<<<(unsigned)grid, WARPS * 32, smem, ctx->stream>>>(J, ctx->kv.canon_tag, reinterpret_cast<const uint4*>(ctx->kv.H), ctx->d_work, 1u);
    return PV_OK;
}

int sigma_run(Ctx* ctx, const SigmaJobs& J) {
    if (J.n == 0) return PV_OK;
    int rc;
    // shape (G edges per warp-group, warps per CTA, CTAs per SM). G*68 hashes should fill whole 32-lane rounds (G = 8: 17
    // rounds exactly, G = 5: 10.6); shared memory per warp = 2576 + 616 G bytes bounds the resident warps, and the total must
    // stay under ~196 KB per SM or the L2 gather ceiling drops by 18 % (see sigma_launch). Since the ALU trimming of round 1 the
    // kernel fits 64 registers without spills and 32 resident warps beat 28 (7.02 vs 7.15 ms per 2^20 edges). Small batches use groups of 2 edges so that more warps (and SMs) take part.
    // The other compiled shapes (the tuning history of profiles/r01_notes.md, two of which skip half the work and return WRONG
    // syndromes) exist only in a -DPVACB_TUNING build (python -m pvac_hfhe_cppbyv_b200.build --tuning -> libpvacb_tuning.so).
    int cfg = 0;
#ifdef PVACB_TUNING
    static int env_cfg = -1;
    if (env_cfg < 0) {
        const char* e = getenv("PVACB_SIGMA_CFG");
        env_cfg = e ? atoi(e) : 0;
    }
    cfg = env_cfg;
#endif
    if (ctx->sigma_test_shape) rc = sigma_launch<5, 4, 8, true, 0, 4, 32>(ctx, J);   // test shape (pvacb_debug_set): no spare candidates, the PRG continuation runs for ~3 of 4 edges
    else if (J.n < (uint64_t)ctx->sm_count * 8 * 4 * 5) rc = sigma_launch<2, 4, 8, true, 0, 4>(ctx, J);
    else switch (cfg) {
        default: rc = sigma_launch<5, 4, 8, true, 0, 4>(ctx, J); break;   // 32 warps/SM at 64 registers, 183 KB of shared memory: L1 keeps 45 KB
#ifdef PVACB_TUNING
        case 1: rc = sigma_launch<8, 4, 6, true, 0, 4>(ctx, J); break;    // 24 warps/SM, exact 17-round groups
        case 2: rc = sigma_launch<6, 4, 7, true, 0, 4>(ctx, J); break;    // 28 warps/SM at 72 registers (the default before the ALU trimming)
        case 3: rc = sigma_launch<8, 4, 7, true, 0, 4>(ctx, J); break;    // 28 warps/SM but 217 KB shared: the slow L1 split
        case 4: rc = sigma_launch<5, 4, 8, true, 0, 4, kCandHashes, 0>(ctx, J); break;   // the default shape with fully unrolled compressions
        case 5: rc = sigma_launch<5, 4, 8, true, 0, 4, kCandHashes, 1>(ctx, J); break;   // rounds 0..15 unrolled + 3-trip loop (two copies of the round code)
        case 6: rc = sigma_launch<5, 4, 8, true, 0, 4, kCandHashes, 3>(ctx, J); break;   // A/B: exact candidate packing only
        case 8: rc = sigma_launch<8, 4, 5, true, 0, 8>(ctx, J); break;    // 20 warps/SM, 16 loads in flight per lane
        case 9: rc = sigma_launch<8, 4, 6, true, 0, 8>(ctx, J); break;    // 24 warps/SM, 16 loads in flight per lane
        case 21: rc = sigma_launch<5, 4, 8, true, 1, 4>(ctx, J); break;   // experiment: no gather loads (wrong results)
        case 22: rc = sigma_launch<5, 4, 8, true, 2, 4>(ctx, J); break;   // experiment: no hashing (wrong results)
#endif
    }
    if (rc) return rc;
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    ctx->stat_sigma_edges += J.n;
    return PV_OK;
}

}  // namespace pvacb
