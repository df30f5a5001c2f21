// AES-256 for the PRF keystream (restates AesCtr256, crypto/lpn.hpp:41-149, which is AES-NI on the CPU).
// sm_100a has no AES instruction: the hot path uses four T-tables replicated 32x in shared memory so that lane L only
// ever touches bank L (conflict-free for any data), see aes_block_rep(). State columns are little-endian words
// (byte 4c+i of the block = bits 8i.. of word c), T0[a] = (2S[a], S[a], S[a], 3S[a]) from the low byte up.
#pragma once
#include "common.cuh"

namespace pvacb {

struct AesTables {
    uint8_t sbox[256];
    uint32_t t0[256];
};

inline uint8_t aes_gmul_host(uint8_t a, uint8_t b) {
    uint8_t r = 0;
    while (b) {
        if (b & 1) r ^= a;
        a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0));
        b >>= 1;
    }
    return r;
}
// FIPS 197 section 5.1.1: multiplicative inverse in GF(2^8) then the affine map
inline void aes_make_tables(AesTables& t) {
    for (int x = 0; x < 256; x++) {
        uint8_t inv = 0;
        if (x)
            for (int y = 1; y < 256; y++)
                if (aes_gmul_host((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        uint8_t s = inv, r = inv;
        for (int k = 0; k < 4; k++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        s ^= 0x63;
        t.sbox[x] = s;
        t.t0[x] = (uint32_t)aes_gmul_host(s, 2) | ((uint32_t)s << 8) | ((uint32_t)s << 16) | ((uint32_t)aes_gmul_host(s, 3) << 24);
    }
}

PV_HD uint32_t aes_rotl(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(x, x, n);
#else
    return (x << n) | (x >> (32 - n));
#endif
}
PV_HD uint32_t aes_subword(const uint8_t* sbox, uint32_t w) {
    return (uint32_t)sbox[w & 0xff] | ((uint32_t)sbox[(w >> 8) & 0xff] << 8) | ((uint32_t)sbox[(w >> 16) & 0xff] << 16) |
           ((uint32_t)sbox[w >> 24] << 24);
}
// key schedule (crypto/lpn.hpp:47-86); key words = little-endian loads of the 32 key bytes
PV_HD void aes256_expand(const uint8_t* sbox, const uint32_t key[8], uint32_t rk[60]) {
    for (int i = 0; i < 8; i++) rk[i] = key[i];
    uint32_t rcon = 1;
    for (int i = 8; i < 60; i++) {
        uint32_t t = rk[i - 1];
        if ((i & 7) == 0) {
            t = aes_subword(sbox, (t >> 8) | (t << 24)) ^ rcon;
            rcon = (rcon << 1) ^ ((rcon & 0x80) ? 0x11b : 0);
        } else if ((i & 7) == 4) {
            t = aes_subword(sbox, t);
        }
        rk[i] = rk[i - 8] ^ t;
    }
}

// generic single-table block (setup/finalize kernels and host tests). in = LE64(ctr) || 0^8 (crypto/lpn.hpp:84,104)
PV_HD void aes256_ctr_block(const uint32_t* T0, const uint8_t* sbox, const uint32_t* rk, uint64_t ctr, uint64_t& w0, uint64_t& w1) {
    uint32_t s0 = (uint32_t)ctr ^ rk[0], s1 = (uint32_t)(ctr >> 32) ^ rk[1], s2 = rk[2], s3 = rk[3];
    for (int r = 1; r < 14; r++) {
        const uint32_t* k = rk + 4 * r;
        uint32_t t0 = T0[s0 & 0xff] ^ aes_rotl(T0[(s1 >> 8) & 0xff], 8) ^ aes_rotl(T0[(s2 >> 16) & 0xff], 16) ^ aes_rotl(T0[s3 >> 24], 24) ^ k[0];
        uint32_t t1 = T0[s1 & 0xff] ^ aes_rotl(T0[(s2 >> 8) & 0xff], 8) ^ aes_rotl(T0[(s3 >> 16) & 0xff], 16) ^ aes_rotl(T0[s0 >> 24], 24) ^ k[1];
        uint32_t t2 = T0[s2 & 0xff] ^ aes_rotl(T0[(s3 >> 8) & 0xff], 8) ^ aes_rotl(T0[(s0 >> 16) & 0xff], 16) ^ aes_rotl(T0[s1 >> 24], 24) ^ k[2];
        uint32_t t3 = T0[s3 & 0xff] ^ aes_rotl(T0[(s0 >> 8) & 0xff], 8) ^ aes_rotl(T0[(s1 >> 16) & 0xff], 16) ^ aes_rotl(T0[s2 >> 24], 24) ^ k[3];
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    const uint32_t* k = rk + 56;
    uint32_t o0 = ((uint32_t)sbox[s0 & 0xff] | ((uint32_t)sbox[(s1 >> 8) & 0xff] << 8) | ((uint32_t)sbox[(s2 >> 16) & 0xff] << 16) | ((uint32_t)sbox[s3 >> 24] << 24)) ^ k[0];
    uint32_t o1 = ((uint32_t)sbox[s1 & 0xff] | ((uint32_t)sbox[(s2 >> 8) & 0xff] << 8) | ((uint32_t)sbox[(s3 >> 16) & 0xff] << 16) | ((uint32_t)sbox[s0 >> 24] << 24)) ^ k[1];
    uint32_t o2 = ((uint32_t)sbox[s2 & 0xff] | ((uint32_t)sbox[(s3 >> 8) & 0xff] << 8) | ((uint32_t)sbox[(s0 >> 16) & 0xff] << 16) | ((uint32_t)sbox[s1 >> 24] << 24)) ^ k[2];
    uint32_t o3 = ((uint32_t)sbox[s3 & 0xff] | ((uint32_t)sbox[(s0 >> 8) & 0xff] << 8) | ((uint32_t)sbox[(s1 >> 16) & 0xff] << 16) | ((uint32_t)sbox[s2 >> 24] << 24)) ^ k[3];
    w0 = (uint64_t)o0 | ((uint64_t)o1 << 32);
    w1 = (uint64_t)o2 | ((uint64_t)o3 << 32);
}

#if defined(__CUDACC__)
// ---- hot path: lane-private-bank replicated tables in shared memory.
// Layout: word index ((t*256 + a) * 32 + lane), t = 0..3 (T_t = T0 rotated left by 8t bits): 128 KiB.
constexpr int kAesRepWords = 4 * 256 * 32;
constexpr int kAesRepBytes = kAesRepWords * 4;

__device__ __forceinline__ void aes_fill_rep_tables(uint32_t* sT, const uint32_t* __restrict__ gT0) {
    for (int i = threadIdx.x; i < kAesRepWords; i += blockDim.x) {
        int a = (i >> 5) & 255, t = i >> 13;
        uint32_t v = __ldg(gT0 + a);
        sT[i] = __funnelshift_l(v, v, 8 * t);
    }
}

// Tl = sT + lane. One AES-256 block of the counter stream; rk[60] in registers.
#define PVACB_TL(t, a) Tl[(((t) << 8) + (a)) << 5]
__device__ __forceinline__ void aes_block_rep(const uint32_t* __restrict__ Tl, const uint32_t (&rk)[60], uint64_t ctr, uint64_t& w0, uint64_t& w1) {
    uint32_t s0 = (uint32_t)ctr ^ rk[0], s1 = (uint32_t)(ctr >> 32) ^ rk[1], s2 = rk[2], s3 = rk[3];
#pragma unroll
    for (int r = 1; r < 14; r++) {
        uint32_t t0 = PVACB_TL(0, s0 & 0xff) ^ PVACB_TL(1, (s1 >> 8) & 0xff) ^ PVACB_TL(2, (s2 >> 16) & 0xff) ^ PVACB_TL(3, s3 >> 24) ^ rk[4 * r + 0];
        uint32_t t1 = PVACB_TL(0, s1 & 0xff) ^ PVACB_TL(1, (s2 >> 8) & 0xff) ^ PVACB_TL(2, (s3 >> 16) & 0xff) ^ PVACB_TL(3, s0 >> 24) ^ rk[4 * r + 1];
        uint32_t t2 = PVACB_TL(0, s2 & 0xff) ^ PVACB_TL(1, (s3 >> 8) & 0xff) ^ PVACB_TL(2, (s0 >> 16) & 0xff) ^ PVACB_TL(3, s1 >> 24) ^ rk[4 * r + 2];
        uint32_t t3 = PVACB_TL(0, s3 & 0xff) ^ PVACB_TL(1, (s0 >> 8) & 0xff) ^ PVACB_TL(2, (s1 >> 16) & 0xff) ^ PVACB_TL(3, s2 >> 24) ^ rk[4 * r + 3];
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    // last round: S-box bytes picked out of the T-tables (T2 has S in byte 0, T3 in byte 1, T0 in byte 2, T1 in byte 3)
    uint32_t o0 = (PVACB_TL(2, s0 & 0xff) & 0x000000ffu) | (PVACB_TL(3, (s1 >> 8) & 0xff) & 0x0000ff00u) | (PVACB_TL(0, (s2 >> 16) & 0xff) & 0x00ff0000u) | (PVACB_TL(1, s3 >> 24) & 0xff000000u);
    uint32_t o1 = (PVACB_TL(2, s1 & 0xff) & 0x000000ffu) | (PVACB_TL(3, (s2 >> 8) & 0xff) & 0x0000ff00u) | (PVACB_TL(0, (s3 >> 16) & 0xff) & 0x00ff0000u) | (PVACB_TL(1, s0 >> 24) & 0xff000000u);
    uint32_t o2 = (PVACB_TL(2, s2 & 0xff) & 0x000000ffu) | (PVACB_TL(3, (s3 >> 8) & 0xff) & 0x0000ff00u) | (PVACB_TL(0, (s0 >> 16) & 0xff) & 0x00ff0000u) | (PVACB_TL(1, s1 >> 24) & 0xff000000u);
    uint32_t o3 = (PVACB_TL(2, s3 & 0xff) & 0x000000ffu) | (PVACB_TL(3, (s0 >> 8) & 0xff) & 0x0000ff00u) | (PVACB_TL(0, (s1 >> 16) & 0xff) & 0x00ff0000u) | (PVACB_TL(1, s2 >> 24) & 0xff000000u);
    o0 ^= rk[56]; o1 ^= rk[57]; o2 ^= rk[58]; o3 ^= rk[59];
    w0 = (uint64_t)o0 | ((uint64_t)o1 << 32);
    w1 = (uint64_t)o2 | ((uint64_t)o3 << 32);
}
#endif

}  // namespace pvacb
