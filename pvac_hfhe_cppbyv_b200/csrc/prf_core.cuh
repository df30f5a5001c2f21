// Per-thread bodies of the PRF kernels, shared between device code (prf.cu) and the host unit tests
// (hosttest.cpp, which runs them on the CPU against the oracle; the product never runs them on the CPU).
#pragma once
#include "common.cuh"
#include "fp127.cuh"
#include "sha256.cuh"
#include "aes256.cuh"

namespace pvacb {

// derive_aes_key (crypto/lpn.hpp:166-192) from the per-key SHA midstate: key words (little-endian) and the CTR start
PV_HD void prf_derive_key(const uint32_t kd_mid[8], uint64_t digest3, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint64_t domhash, uint32_t key[8]) {
    ShaState st;
#pragma unroll
    for (int i = 0; i < 8; i++) st.h[i] = kd_mid[i];
    // block 1 of the 104-byte message: H_digest[24..31], ztag, nonce.lo, nonce.hi, fnv1a(dom), 0x80 pad, bit length
    uint64_t q[8] = {digest3, ztag, nlo, nhi, domhash, 0x80ull, 0ull, 0ull};
    uint32_t w[16];
    sha_block_from_le64(q, w);
    w[14] = 0;
    w[15] = 104 * 8;
    sha_compress(st, w);
#pragma unroll
    for (int i = 0; i < 8; i++) key[i] = sha_bswap(st.h[i]);   // key bytes = big-endian digest bytes, loaded little-endian
}

// one PRF core (crypto/lpn.hpp:235-254): LPN round keys + counter start, and the first block of the Toeplitz key stream
PV_HD void prf_core_setup(const uint32_t kd_mid[8], uint64_t digest3, const uint32_t* T0, const uint8_t* sbox, uint64_t ztag, uint64_t nlo,
                          uint64_t nhi, uint64_t domhash, uint32_t rk[60], uint64_t& ctr0, uint64_t& top0, uint64_t& top1) {
    uint32_t key[8];
    uint32_t rkt[60];
    prf_derive_key(kd_mid, digest3, ztag, nlo, nhi, kFnvToep, key);
    aes256_expand(sbox, key, rkt);
    // only the first 127 bits of the 258-word Toeplitz stream are ever read (crypto/toeplitz.hpp:153-162)
    aes256_ctr_block(T0, sbox, rkt, kFnvToep ^ nlo ^ domhash, top0, top1);
    prf_derive_key(kd_mid, digest3, ztag, nlo, nhi, domhash, key);
    aes256_expand(sbox, key, rk);
    ctr0 = domhash ^ nlo;
}

// LPN rows 2p and 2p+1 (crypto/lpn.hpp:219-232): 130 stream words = AES blocks ctr..ctr+64. Word w of that span belongs to the
// even row's dot product (w < 64, with s[w]), is its noise word (w = 64), belongs to the odd row's dot product
// (65 <= w < 129, with s[w-65]) or is the odd row's noise word (w = 129). The two dot products are written as ONE loop over
// the 65 blocks with per-word masks (e[w] = s[w] or 0, o[w] = s[w-65] or 0): a single small loop body instead of four
// specialised ones keeps the kernel inside the instruction cache (ncu r01: 0.59 no_instruction stalls per issue with the
// unrolled form). The masks sit in the kernel parameter bank and are read with a warp-uniform index.
struct LpnMasks {
    uint64_t e[130];
    uint64_t o[130];
};
inline void lpn_masks_from_secret(const uint64_t s[kLpnWords], LpnMasks& m) {
    for (int w = 0; w < 130; w++) {
        m.e[w] = w < 64 ? s[w] : 0ull;
        m.o[w] = (w >= 65 && w < 129) ? s[w - 65] : 0ull;
    }
}

// blk(counter, w0, w1) yields one keystream block. Returns the two y bits; `rare` is set if bounded(8) would have rejected.
template <class BlockFn>
PV_HD void lpn_row_pair(BlockFn&& blk, uint64_t ctr, const LpnMasks& m, uint32_t& ye, uint32_t& yo, bool& rare) {
    uint64_t accE = 0, accO = 0, nwE = 0, nwO = 0;
#pragma unroll 1
    for (int q = 0; q < 65; q++) {
        uint64_t w0, w1;
        blk(ctr + q, w0, w1);
        accE ^= (w0 & m.e[2 * q]) ^ (w1 & m.e[2 * q + 1]);
        accO ^= (w0 & m.o[2 * q]) ^ (w1 & m.o[2 * q + 1]);
        if (q == 32) nwE = w0;      // stream word 64: noise word of the even row
        if (q == 64) nwO = w1;      // stream word 129: noise word of the odd row
    }
    const uint32_t nE = ((uint32_t)nwE & 7u) == 0u;        // bounded(8) < 1  (crypto/lpn.hpp:141-148,228)
    const uint32_t nO = ((uint32_t)nwO & 7u) == 0u;
    rare |= (nwE >= 0xFFFFFFFFFFFFFFF8ull) | (nwO >= 0xFFFFFFFFFFFFFFF8ull);   // rejection branch of bounded(): would shift the whole stream
#if defined(__CUDA_ARCH__)
    ye = (__popcll(accE) & 1) ^ nE;
    yo = (__popcll(accO) & 1) ^ nO;
#else
    ye = (uint32_t)(__builtin_popcountll(accE) & 1) ^ nE;
    yo = (uint32_t)(__builtin_popcountll(accO) & 1) ^ nO;
#endif
}

PV_HD uint64_t spread_bits32(uint32_t x) {  // bit i -> bit 2i
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// out_j = XOR_{i<=j} y_i top_{j-i}, j < 127 (crypto/toeplitz.hpp:121-141), then hash_to_fp_nonzero (lpn.hpp:25-37)
PV_HD Fp toep127_to_fp(uint64_t y0, uint64_t y1, uint64_t t0, uint64_t t1) {
    uint64_t lo = 0, hi = 0;
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        uint64_t m = 0ull - ((y0 >> i) & 1ull);
        lo ^= m & (t0 << i);
        hi ^= m & (i ? ((t1 << i) | (t0 >> (64 - i))) : t1);
    }
#pragma unroll 1
    for (int i = 0; i < 63; i++) {
        uint64_t m = 0ull - ((y1 >> i) & 1ull);
        hi ^= m & (t0 << i);
    }
    return hash_to_fp_nonzero(lo, hi & kMask63);
}

}  // namespace pvacb
