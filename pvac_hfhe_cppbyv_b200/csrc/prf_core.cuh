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

// LPN rows 2p and 2p+1 (crypto/lpn.hpp:219-232): 130 stream words = AES blocks ctr..ctr+64. blk(counter, w0, w1) yields
// one keystream block. s = the 64 secret words. Returns the two y bits; `rare` is set if bounded(8) would have rejected.
template <class BlockFn>
PV_HD void lpn_row_pair(BlockFn&& blk, uint64_t ctr, const uint64_t* __restrict__ s, uint32_t& ye, uint32_t& yo, bool& rare) {
    uint64_t accE = 0, accO = 0, w0, w1;
#pragma unroll 2
    for (int q = 0; q < 32; q++) {                // even row: its 64 words are blocks 0..31
        blk(ctr + q, w0, w1);
        accE ^= (w0 & s[2 * q]) ^ (w1 & s[2 * q + 1]);
    }
    blk(ctr + 32, w0, w1);                         // block 32: noise word of the even row | word 0 of the odd row
    uint32_t nE = ((uint32_t)w0 & 7u) == 0u;       // bounded(8) < 1  (crypto/lpn.hpp:141-148,228)
    rare |= w0 >= 0xFFFFFFFFFFFFFFF8ull;            // rejection branch of bounded(): would shift the whole stream
    accO ^= w1 & s[0];
#pragma unroll 2
    for (int j = 0; j < 31; j++) {                 // blocks 33..63: odd-row words 1+2j, 2+2j
        blk(ctr + 33 + j, w0, w1);
        accO ^= (w0 & s[1 + 2 * j]) ^ (w1 & s[2 + 2 * j]);
    }
    blk(ctr + 64, w0, w1);                         // block 64: odd-row word 63 | noise word of the odd row
    accO ^= w0 & s[63];
    uint32_t nO = ((uint32_t)w1 & 7u) == 0u;
    rare |= w1 >= 0xFFFFFFFFFFFFFFF8ull;
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
