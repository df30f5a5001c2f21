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
// ---- hot path: lane-private-bank replicated tables in shared memory (128 KiB).
// Byte offset of (table t, entry a, lane L) = (t >> 1) * 65536 + a * 256 + (t & 1) * 128 + L * 4   (T_t = T0 rotated left 8t bits).
// Lane L only ever touches bank L, so every lookup is one conflict-free wavefront whatever the data. The 256-byte entry
// stride lets ONE PRMT build the whole shared-memory offset of a lookup: bytes (lane*4, state byte k, 0, 0); the table
// select rides in the LDS immediate. (ncu r01: the first version spent 48 ALU ops per round on shift/mask/add address
// arithmetic and was ALU-pipe bound at 88%; this form needs 16 PRMT + 8 LOP3.)
// The tables sit at ABSOLUTE shared-window addresses 0x10000 (T0/T1) and 0x20000 (T2/T3), so the PRMT result is the
// complete address and the table select is an LDS immediate -- no per-lookup add of the dynamic-smem base. The kernel
// therefore asks for 0x30000 bytes of dynamic shared memory (the first 64 KiB minus the base stay unused).
constexpr int kAesRepBytes = 0x30000;
constexpr uint32_t kAesTabAbs = 0x10000;

__device__ __forceinline__ void aes_fill_rep_tables(uint8_t* sT, const uint32_t* __restrict__ gT0) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sT);
    for (int i = threadIdx.x; i < 4 * 256 * 32; i += blockDim.x) {
        int odd = (i >> 5) & 1, a = (i >> 6) & 255, tp = i >> 14;
        uint32_t v = __ldg(gT0 + a);
        uint32_t abs_addr = kAesTabAbs + (uint32_t)i * 4;       // = 0x10000*(1+tp) + a*256 + odd*128 + lane*4
        *reinterpret_cast<uint32_t*>(sT + (abs_addr - base)) = __funnelshift_l(v, v, 8 * (2 * tp + odd));
    }
}

template <int kOff>
__device__ __forceinline__ uint32_t aes_lds_abs(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(kOff));
    return v;
}
// address of the lookup indexed by byte k of s: (lane4 | byte_k << 8); table t selected by the immediate
#define PVACB_IDX(s, k) __byte_perm((s), lane4, 0x5504 | ((k) << 4))
#define PVACB_LD(t, idx) aes_lds_abs<(int)kAesTabAbs + ((t) >> 1) * 65536 + ((t) & 1) * 128>(idx)

// AES-256 counter-mode block generator for one thread: round keys in registers. The input block is LE64(ctr) || 0^8
// (crypto/lpn.hpp:84,104) and a thread walks 65 CONSECUTIVE counters, so between two blocks only byte 0 of the state
// changes (until the low byte wraps, once per 256 blocks). Everything of rounds 1 and 2 that does not depend on that byte
// is hoisted: round 1 keeps one lookup (T0 of the changing byte; the other three columns u1..u3 are constant), round 2
// keeps the four lookups indexed by the bytes of the one changing column; the constant parts are c0 and d0..d3.
// 1 + 4 + 11*16 + 16 = 197 lookups per block instead of 224.
struct AesCtrThread {
    uint32_t rk0;              // rk[0]
    uint32_t c0;               // round 1, column 0 without the T0[x.b0] term
    uint32_t d0, d1, d2, d3;   // round 2 without the terms of column 0
    uint64_t tag;              // ctr >> 8 the constants were computed for
    uint32_t rk1, rk2, rk3, rk4, rk5, rk6, rk7;
    uint32_t rk[52];           // rk[8..59]

    __device__ __forceinline__ void prime(const uint8_t* __restrict__ sT, uint32_t lane4, uint64_t ctr) {
        tag = ctr >> 8;
        const uint32_t x0 = (uint32_t)ctr ^ rk0, x1 = (uint32_t)(ctr >> 32) ^ rk1, x2 = rk2, x3 = rk3;
        c0 = PVACB_LD(1, PVACB_IDX(x1, 1)) ^ PVACB_LD(2, PVACB_IDX(x2, 2)) ^ PVACB_LD(3, PVACB_IDX(x3, 3)) ^ rk4;
        const uint32_t u1 = PVACB_LD(0, PVACB_IDX(x1, 0)) ^ PVACB_LD(1, PVACB_IDX(x2, 1)) ^ PVACB_LD(2, PVACB_IDX(x3, 2)) ^ PVACB_LD(3, PVACB_IDX(x0, 3)) ^ rk5;
        const uint32_t u2 = PVACB_LD(0, PVACB_IDX(x2, 0)) ^ PVACB_LD(1, PVACB_IDX(x3, 1)) ^ PVACB_LD(2, PVACB_IDX(x0, 2)) ^ PVACB_LD(3, PVACB_IDX(x1, 3)) ^ rk6;
        const uint32_t u3 = PVACB_LD(0, PVACB_IDX(x3, 0)) ^ PVACB_LD(1, PVACB_IDX(x0, 1)) ^ PVACB_LD(2, PVACB_IDX(x1, 2)) ^ PVACB_LD(3, PVACB_IDX(x2, 3)) ^ rk7;
        d0 = PVACB_LD(1, PVACB_IDX(u1, 1)) ^ PVACB_LD(2, PVACB_IDX(u2, 2)) ^ PVACB_LD(3, PVACB_IDX(u3, 3)) ^ rk[0];
        d1 = PVACB_LD(0, PVACB_IDX(u1, 0)) ^ PVACB_LD(1, PVACB_IDX(u2, 1)) ^ PVACB_LD(2, PVACB_IDX(u3, 2)) ^ rk[1];
        d2 = PVACB_LD(0, PVACB_IDX(u2, 0)) ^ PVACB_LD(1, PVACB_IDX(u3, 1)) ^ PVACB_LD(3, PVACB_IDX(u1, 3)) ^ rk[2];
        d3 = PVACB_LD(0, PVACB_IDX(u3, 0)) ^ PVACB_LD(2, PVACB_IDX(u1, 2)) ^ PVACB_LD(3, PVACB_IDX(u2, 3)) ^ rk[3];
    }

    __device__ __forceinline__ void load_keys(const uint32_t* __restrict__ rk_core) {
        const uint4* p = reinterpret_cast<const uint4*>(rk_core);
        uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
        rk0 = v0.x; rk1 = v0.y; rk2 = v0.z; rk3 = v0.w;
        rk4 = v1.x; rk5 = v1.y; rk6 = v1.z; rk7 = v1.w;
#pragma unroll
        for (int i = 0; i < 13; i++) {
            uint4 v = __ldg(p + 2 + i);
            rk[4 * i] = v.x; rk[4 * i + 1] = v.y; rk[4 * i + 2] = v.z; rk[4 * i + 3] = v.w;
        }
    }

    __device__ __forceinline__ void block(const uint8_t* __restrict__ sT, uint32_t lane4, uint64_t ctr, uint64_t& w0, uint64_t& w1) {
        if ((ctr >> 8) != tag) prime(sT, lane4, ctr);   // the low counter byte wrapped: at most twice per thread
        const uint32_t x = (uint32_t)ctr ^ rk0;
        const uint32_t t0 = c0 ^ PVACB_LD(0, PVACB_IDX(x, 0));                 // round 1: the one column that changes
        uint32_t s0 = d0 ^ PVACB_LD(0, PVACB_IDX(t0, 0));                      // round 2: its four bytes
        uint32_t s1 = d1 ^ PVACB_LD(3, PVACB_IDX(t0, 3));
        uint32_t s2 = d2 ^ PVACB_LD(2, PVACB_IDX(t0, 2));
        uint32_t s3 = d3 ^ PVACB_LD(1, PVACB_IDX(t0, 1));
#pragma unroll
        for (int r = 3; r < 14; r++) {
            const int kb = 4 * r - 8;
            uint32_t t0 = PVACB_LD(0, PVACB_IDX(s0, 0)) ^ PVACB_LD(1, PVACB_IDX(s1, 1)) ^ PVACB_LD(2, PVACB_IDX(s2, 2)) ^ PVACB_LD(3, PVACB_IDX(s3, 3)) ^ rk[kb + 0];
            uint32_t t1 = PVACB_LD(0, PVACB_IDX(s1, 0)) ^ PVACB_LD(1, PVACB_IDX(s2, 1)) ^ PVACB_LD(2, PVACB_IDX(s3, 2)) ^ PVACB_LD(3, PVACB_IDX(s0, 3)) ^ rk[kb + 1];
            uint32_t t2 = PVACB_LD(0, PVACB_IDX(s2, 0)) ^ PVACB_LD(1, PVACB_IDX(s3, 1)) ^ PVACB_LD(2, PVACB_IDX(s0, 2)) ^ PVACB_LD(3, PVACB_IDX(s1, 3)) ^ rk[kb + 2];
            uint32_t t3 = PVACB_LD(0, PVACB_IDX(s3, 0)) ^ PVACB_LD(1, PVACB_IDX(s0, 1)) ^ PVACB_LD(2, PVACB_IDX(s1, 2)) ^ PVACB_LD(3, PVACB_IDX(s2, 3)) ^ rk[kb + 3];
            s0 = t0; s1 = t1; s2 = t2; s3 = t3;
        }
        // last round: S-box bytes picked out of the T-tables (T2 has S in byte 0, T3 in byte 1, T0 in byte 2, T1 in byte 3)
        uint32_t o0 = (PVACB_LD(2, PVACB_IDX(s0, 0)) & 0x000000ffu) | (PVACB_LD(3, PVACB_IDX(s1, 1)) & 0x0000ff00u) | (PVACB_LD(0, PVACB_IDX(s2, 2)) & 0x00ff0000u) | (PVACB_LD(1, PVACB_IDX(s3, 3)) & 0xff000000u);
        uint32_t o1 = (PVACB_LD(2, PVACB_IDX(s1, 0)) & 0x000000ffu) | (PVACB_LD(3, PVACB_IDX(s2, 1)) & 0x0000ff00u) | (PVACB_LD(0, PVACB_IDX(s3, 2)) & 0x00ff0000u) | (PVACB_LD(1, PVACB_IDX(s0, 3)) & 0xff000000u);
        uint32_t o2 = (PVACB_LD(2, PVACB_IDX(s2, 0)) & 0x000000ffu) | (PVACB_LD(3, PVACB_IDX(s3, 1)) & 0x0000ff00u) | (PVACB_LD(0, PVACB_IDX(s0, 2)) & 0x00ff0000u) | (PVACB_LD(1, PVACB_IDX(s1, 3)) & 0xff000000u);
        uint32_t o3 = (PVACB_LD(2, PVACB_IDX(s3, 0)) & 0x000000ffu) | (PVACB_LD(3, PVACB_IDX(s0, 1)) & 0x0000ff00u) | (PVACB_LD(0, PVACB_IDX(s1, 2)) & 0x00ff0000u) | (PVACB_LD(1, PVACB_IDX(s2, 3)) & 0xff000000u);
        o0 ^= rk[48]; o1 ^= rk[49]; o2 ^= rk[50]; o3 ^= rk[51];
        w0 = (uint64_t)o0 | ((uint64_t)o1 << 32);
        w1 = (uint64_t)o2 | ((uint64_t)o3 << 32);
    }
};
#endif

}  // namespace pvacb
