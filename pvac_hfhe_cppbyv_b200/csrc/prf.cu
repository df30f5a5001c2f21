// PRF path: prf_R / prf_R_noise of the reference (crypto/lpn.hpp:194-275, crypto/toeplitz.hpp:121-162).
//
//   prf_setup_kernel    one thread per PRF core: SHA-256 key derivation (midstate reuse), AES-256 key schedule,
//                       first block of the Toeplitz key stream ("top")
//   prf_lpn_kernel      THE HOT KERNEL of enc_value / dec_value: AES-256-CTR keystream fused with the LPN parity.
//                       One thread per ROW PAIR (rows 2p, 2p+1 = 130 stream words = exactly 65 AES blocks), so no
//                       cross-lane reduction is needed; T-tables replicated per lane in 128 KiB of shared memory
//                       (bank-conflict free); the secret s sits in the kernel parameter bank; a warp's 64 y-bits are
//                       assembled with two __ballot_sync and one 64-bit store. The keystream is never stored.
//   prf_finalize_kernel one thread per job: 127-bit truncated carry-less product, hash_to_fp_nonzero, r1*r2*r3
//
// Row layout (crypto/lpn.hpp:219-232, word FIFO of :108-139): stream word w = half (w&1) of block (w>>1); row r uses
// words 65r..65r+63 for the dot product with s and word 65r+64 for the Bernoulli(1/8) noise bit.
#include "engine.h"
#include "prf_core.cuh"

namespace pvacb {

// ------------------------------------------------------------------ setup
__global__ void prf_setup_kernel(KeyView kv, uint64_t ncores, const uint64_t* __restrict__ ztag, const uint64_t* __restrict__ nlo,
                                 const uint64_t* __restrict__ nhi, const uint8_t* __restrict__ flags, uint32_t* __restrict__ rk_out,
                                 uint64_t* __restrict__ ctr0_out, uint64_t* __restrict__ top_out, unsigned int* __restrict__ active_cores) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = c < ncores;
    uint64_t job = in_range ? c / 3 : 0;
    int t = (int)(c % 3);
    uint8_t fl = in_range ? flags[job] : 0;
    const bool active = in_range && (fl & 2);
    const unsigned m = __ballot_sync(0xffffffffu, active);          // cores really evaluated (inactive jobs cost nothing): for the statistics
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(active_cores, (unsigned)__popc(m));
    if (!active) return;
    uint32_t rk[60];
    uint64_t ctr0, top0, top1;
    prf_core_setup(kv.kd_mid, kv.digest3, kv.T0, kv.sbox, ztag[job], nlo[job], nhi[job], fnv_prf_dom(fl & 1, t), rk, ctr0, top0, top1);
    for (int i = 0; i < 60; i++) rk_out[c * 60 + i] = rk[i];
    ctr0_out[c] = ctr0;
    top_out[2 * c] = top0;
    top_out[2 * c + 1] = top1;
}

// ------------------------------------------------------------------ LPN rows
constexpr int kLpnThreads = 512;

// rpc_log2: log2(row pairs evaluated per core): 13 (all 16384 rows, as the reference) or 6 (rows 0..127, the only ones
// toep_127 can see). Persistent grid: one CTA per SM, static round-robin over units of kLpnThreads row pairs.
// A noise word >= 2^64 - 8 makes AesCtr256::bounded draw again (crypto/lpn.hpp:141-148), which shifts every later word of that
// core's stream by one: such a core is only FLAGGED here (rare_core) and recomputed serially by prf_lpn_serial_kernel.
// PATCH: test hook, keystream word patch_word of every core is OR-ed with patch_or (the hot instantiation has none of it).
template <bool PATCH>
__global__ void __launch_bounds__(kLpnThreads, 1)
prf_lpn_kernel(const uint32_t* __restrict__ gT0, const __grid_constant__ LpnMasks msk, uint64_t ncores, int rpc_log2,
               const uint32_t* __restrict__ rk_all, const uint64_t* __restrict__ ctr0_all, const uint8_t* __restrict__ flags,
               uint64_t* __restrict__ ybits, unsigned int* __restrict__ rare_flag, uint8_t* __restrict__ rare_core, uint64_t patch_word, uint64_t patch_or) {
    extern __shared__ __align__(16) uint8_t sT[];
    aes_fill_rep_tables(sT, gT0);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t lane4 = (uint32_t)lane * 4;
    const uint64_t slots = ncores << rpc_log2;
    const uint64_t units = (slots + kLpnThreads - 1) / kLpnThreads;
    const uint32_t rp_mask = (1u << rpc_log2) - 1;

    for (uint64_t u = blockIdx.x; u < units; u += gridDim.x) {
        uint64_t slot = u * kLpnThreads + threadIdx.x;
        bool active = slot < slots;
        uint64_t core = active ? (slot >> rpc_log2) : 0;
        uint32_t rp = (uint32_t)slot & rp_mask;
        if (active && !(flags[core / 3] & 2)) active = false;   // warp-uniform: a core spans whole warps
        uint32_t ye = 0, yo = 0;
        if (active) {
            AesCtrThread aes;
            aes.load_keys(rk_all + core * 60);
            const uint64_t c0 = __ldg(ctr0_all + core);
            uint64_t ctr = c0 + 65ull * rp;
            aes.prime(sT, lane4, ctr);
            bool rare = false;
            lpn_row_pair([&](uint64_t c, uint64_t& w0, uint64_t& w1) {
                aes.block(sT, lane4, c, w0, w1);
                if (PATCH && c - c0 == (patch_word >> 1)) { if (patch_word & 1) w1 |= patch_or; else w0 |= patch_or; }
            }, ctr, msk, ye, yo, rare);
            if (rare) { atomicOr(rare_flag, 1u); rare_core[core] = 1; }
        }
        uint32_t be = __ballot_sync(0xffffffffu, ye);
        uint32_t bo = __ballot_sync(0xffffffffu, yo);
        if (lane == 0 && active) ybits[slot >> 5] = spread_bits32(be) | (spread_bits32(bo) << 1);
    }
}

// The flagged slow path: one thread per flagged core walks the keystream as the reference's word FIFO does (crypto/lpn.hpp:108-148,
// 219-232) -- 64 words for the dot product, then bounded(8): words until one is < 2^64 - 8 -- and rewrites all of the core's y bits.
__global__ void prf_lpn_serial_kernel(KeyView kv, const __grid_constant__ LpnMasks msk, uint64_t ncores, int rows, const uint8_t* __restrict__ rare_core,
                                      const uint32_t* __restrict__ rk_all, const uint64_t* __restrict__ ctr0_all, uint64_t* __restrict__ ybits, int words_per_core,
                                      uint64_t patch_word, uint64_t patch_or) {
    const uint64_t core = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (core >= ncores || !rare_core[core]) return;
    const uint32_t* rk = rk_all + core * 60;
    uint64_t ctr = ctr0_all[core], nout = 0, buf1 = 0;
    bool has = false;
    auto next = [&]() -> uint64_t {
        uint64_t x;
        if (has) { has = false; x = buf1; }
        else { uint64_t w0, w1; aes256_ctr_block(kv.T0, kv.sbox, rk, ctr++, w0, w1); buf1 = w1; has = true; x = w0; }
        if (nout == patch_word) x |= patch_or;
        nout++;
        return x;
    };
    uint64_t* y = ybits + core * (uint64_t)words_per_core;
    uint64_t word = 0;
    for (int r = 0; r < rows; r++) {
        uint64_t acc = 0;
        for (int wi = 0; wi < kLpnWords; wi++) acc ^= next() & msk.e[wi];      // e[w] = s[w] for w < 64
        uint64_t x;
        do { x = next(); } while (x >= 0xFFFFFFFFFFFFFFF8ull);                 // lim = UINT64_MAX - UINT64_MAX % 8
        const uint64_t bit = (uint64_t)((__popcll(acc) & 1) ^ ((x & 7ull) == 0ull ? 1 : 0));
        word |= bit << (r & 63);
        if ((r & 63) == 63 || r == rows - 1) { y[r >> 6] = word; word = 0; }
    }
}

// ------------------------------------------------------------------ finalize
__global__ void prf_finalize_kernel(uint64_t njobs, const uint8_t* __restrict__ flags, const uint64_t* __restrict__ ybits,
                                    int words_per_core, const uint64_t* __restrict__ top, Fp* __restrict__ out) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    if (!(flags[j] & 2)) { out[j] = fp_zero(); return; }
    Fp r[3];
    for (int t = 0; t < 3; t++) {
        uint64_t c = 3 * j + t;
        const uint64_t* y = ybits + c * (uint64_t)words_per_core;
        r[t] = toep127_to_fp(y[0], y[1], top[2 * c], top[2 * c + 1]);
    }
    out[j] = fp_mul(fp_mul(r[0], r[1]), r[2]);   // crypto/lpn.hpp:263-275
}

// ------------------------------------------------------------------ host launcher
int prf_run(Ctx* ctx, uint64_t njobs, const uint64_t* d_ztag, const uint64_t* d_nlo, const uint64_t* d_nhi, const uint8_t* d_flags,
            Fp* d_out, uint64_t* d_ybits_out) {
    if (njobs == 0) return PV_OK;
    Scratch scratch(ctx);
    const uint64_t ncores = njobs * 3;
    const int rpc_log2 = ctx->prf_mode == PRF_LIVE ? 6 : 13;
    const int wpc = (1 << rpc_log2) / 32;  // ybits words per core
    uint32_t* rk = nullptr;
    uint64_t *ctr0 = nullptr, *top = nullptr, *ybits = nullptr;
    unsigned int* rare = nullptr;
    uint8_t* rare_core = nullptr;
    int rc;
    if ((rc = scratch.alloc(rk, ncores * 60 * 4))) return rc;
    if ((rc = scratch.alloc(ctr0, ncores * 8))) return rc;
    if ((rc = scratch.alloc(top, ncores * 16))) return rc;
    if (d_ybits_out) ybits = d_ybits_out;
    else if ((rc = scratch.alloc(ybits, ncores * wpc * 8))) return rc;
    if ((rc = scratch.alloc(rare, 8))) return rc;                    // [0] rare-path flag, [1] active cores
    if ((rc = scratch.alloc(rare_core, ncores))) return rc;
    PV_CUDA(cudaMemsetAsync(rare, 0, 8, ctx->stream));
    PV_CUDA(cudaMemsetAsync(rare_core, 0, ncores, ctx->stream));
    const bool patch = ctx->prf_patch_word != ~0ull;

    prf_setup_kernel<<<(unsigned)((ncores + 127) / 128), 128, 0, ctx->stream>>>(ctx->kv, ncores, d_ztag, d_nlo, d_nhi, d_flags, rk, ctr0, top, rare + 1);
    if (!ctx->lpn_attr_set) {
        PV_CUDA(cudaFuncSetAttribute(prf_lpn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesRepBytes));
        PV_CUDA(cudaFuncSetAttribute(prf_lpn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesRepBytes));
        ctx->lpn_attr_set = true;
    }
    uint64_t units = ((ncores << rpc_log2) + kLpnThreads - 1) / kLpnThreads;
    unsigned grid = (unsigned)(units < (uint64_t)ctx->sm_count ? units : (uint64_t)ctx->sm_count);
    {
        ProfScope ps(ctx, PROF_PRF_LPN);
        if (patch) prf_lpn_kernel<true><<<grid, kLpnThreads, kAesRepBytes, ctx->stream>>>(ctx->kv.T0, ctx->lpn_m, ncores, rpc_log2, rk, ctr0, d_flags, ybits, rare, rare_core,
                                                                                          ctx->prf_patch_word, ctx->prf_patch_or);
        else prf_lpn_kernel<false><<<grid, kLpnThreads, kAesRepBytes, ctx->stream>>>(ctx->kv.T0, ctx->lpn_m, ncores, rpc_log2, rk, ctr0, d_flags, ybits, rare, rare_core, ~0ull, 0ull);
    }
    prf_finalize_kernel<<<(unsigned)((njobs + 127) / 128), 128, 0, ctx->stream>>>(njobs, d_flags, ybits, wpc, top, d_out);
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 3;
    unsigned int h_rare = 0, h_active = 0;
    { SmallRead sr; sr.add(&h_rare, rare, 4); sr.add(&h_active, rare + 1, 4); if ((rc = read_small_sync(ctx, sr))) return rc; }
    ctx->stat_aes_blocks += (uint64_t)h_active * ((65ull << rpc_log2) + 1);
    if (h_rare) {
        // AesCtr256::bounded rejected a noise word somewhere (p = 2^-61 per row): the flagged cores are walked again serially with
        // the reference's word FIFO, then the (cheap) finalize step runs once more over everything
        prf_lpn_serial_kernel<<<(unsigned)((ncores + 63) / 64), 64, 0, ctx->stream>>>(ctx->kv, ctx->lpn_m, ncores, 2 << rpc_log2, rare_core, rk, ctr0, ybits, wpc,
                                                                                         ctx->prf_patch_word, ctx->prf_patch_or);
        prf_finalize_kernel<<<(unsigned)((njobs + 127) / 128), 128, 0, ctx->stream>>>(njobs, d_flags, ybits, wpc, top, d_out);
        PV_CUDA(cudaGetLastError());
        ctx->stat_kernel_launches += 2;
        ctx->stat_rare_prf_cores += 1;
    }
    return PV_OK;
}

}  // namespace pvacb
