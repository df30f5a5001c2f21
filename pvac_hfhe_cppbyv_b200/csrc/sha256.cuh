// SHA-256 (FIPS 180-4) compression on 32-bit registers, plus the little-endian "u64 stream" message layout the
// reference hashes everywhere (core/hash.hpp:24-191: byte-oriented update, u64s fed little-endian by sha256_acc_u64).
//
// Every message on the hot path is  label-bytes || LE64(x0) || LE64(x1) || ...  so a message is handled as a stream of
// little-endian 64-bit words q[] that is the x-stream shifted by (label_len mod 8) bytes; a 64-byte SHA block is 8 such
// words, and big-endian schedule word W[j] is the byte-swap of the j-th 32-bit half.
#pragma once
#include "common.cuh"

namespace pvacb {

#define PVACB_SHA_K_LIST \
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, \
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, \
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, \
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, \
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, \
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, \
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, \
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2
// round constants as compile-time immediates: the 64 rounds are fully unrolled, so K[i] folds into the IADD3 operand
// (a __constant__ table made ptxas hoist 64 LDCs out of the hashing loop and spill them to local memory)
#define PVACB_SHA_K_DECL constexpr uint32_t kShaK[64] = {PVACB_SHA_K_LIST}
#define PVACB_SHA_K(i) kShaK[i]

struct ShaState {
    uint32_t h[8];
};

PV_HD void sha_init(ShaState& s) {
    s.h[0] = 0x6a09e667; s.h[1] = 0xbb67ae85; s.h[2] = 0x3c6ef372; s.h[3] = 0xa54ff53a;
    s.h[4] = 0x510e527f; s.h[5] = 0x9b05688c; s.h[6] = 0x1f83d9ab; s.h[7] = 0x5be0cd19;
}

PV_HD uint32_t sha_rotr(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}
PV_HD uint32_t sha_bswap(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

// one compression; w[16] = big-endian schedule words of the block (destroyed)
PV_HD void sha_compress(ShaState& s, uint32_t w[16]) {
    PVACB_SHA_K_DECL;
    uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
        }
        uint32_t t1 = h + (sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25)) + ((e & f) ^ (~e & g)) + PVACB_SHA_K(i) + w[i & 15];
        uint32_t t2 = (sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}

// one compression starting from a parked state (e.g. a midstate in shared memory): out = compress(from, w). `from` is read
// again for the final addition instead of being kept in registers across the 64 rounds.
PV_HD void sha_compress_from(const uint32_t* from, uint32_t w[16], uint32_t out[8]) {
    PVACB_SHA_K_DECL;
    uint32_t a = from[0], b = from[1], c = from[2], d = from[3], e = from[4], f = from[5], g = from[6], h = from[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
        }
        uint32_t t1 = h + (sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25)) + ((e & f) ^ (~e & g)) + PVACB_SHA_K(i) + w[i & 15];
        uint32_t t2 = (sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    out[0] = from[0] + a; out[1] = from[1] + b; out[2] = from[2] + c; out[3] = from[3] + d;
    out[4] = from[4] + e; out[5] = from[5] + f; out[6] = from[6] + g; out[7] = from[7] + h;
}
#if defined(__CUDACC__)
// Same compression with the two-input additions issued as IMAD (x * one + y, `one` a run-time 1 so that ptxas keeps the
// multiply-add): they run on the FMA pipe, which the hashing leaves idle, instead of the ALU pipe that SHF/LOP3 and the
// gather's XORs already fill. Used by the fused sigma kernel (PVACB_SIGMA_FMA=1), see profiles/r01_notes.md.
__device__ __forceinline__ uint32_t sha_add_fma(uint32_t a, uint32_t b, uint32_t one) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}
__device__ __forceinline__ void sha_compress_from_fma(const uint32_t* from, uint32_t w[16], uint32_t out[8], uint32_t one) {
    PVACB_SHA_K_DECL;
    uint32_t a = from[0], b = from[1], c = from[2], d = from[3], e = from[4], f = from[5], g = from[6], h = from[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] = sha_add_fma(sha_add_fma(w[i & 15], s0, one), sha_add_fma(w[(i - 7) & 15], s1, one), one);
        }
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t ch = (e & f) ^ (~e & g), maj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t1 = sha_add_fma(sha_add_fma(h, S1, one), sha_add_fma(ch, PVACB_SHA_K(i) + w[i & 15], one), one);
        uint32_t t2 = sha_add_fma(S0, maj, one);
        h = g; g = f; f = e; e = sha_add_fma(d, t1, one); d = c; c = b; b = a; a = sha_add_fma(t1, t2, one);
    }
    out[0] = from[0] + a; out[1] = from[1] + b; out[2] = from[2] + c; out[3] = from[3] + d;
    out[4] = from[4] + e; out[5] = from[5] + f; out[6] = from[6] + g; out[7] = from[7] + h;
}
// The same compression with rounds 16..63 as a 3-trip loop of 16 rounds: ~650 instead of ~1 400 instructions of code. The fully
// unrolled form does not fit the instruction cache next to the gather code (ncu r01: 13-19 % no_instruction stalls in the
// hashing phase of the sigma kernel); the round constants of the looped part come from the constant bank.
static __device__ __constant__ uint32_t c_shaK[64] = {PVACB_SHA_K_LIST};
static __device__ __constant__ uint32_t kShaIv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
__device__ __forceinline__ void sha_compress_from_rolled(const uint32_t* from, uint32_t w[16], uint32_t out[8], uint32_t one) {
    PVACB_SHA_K_DECL;
    uint32_t a = from[0], b = from[1], c = from[2], d = from[3], e = from[4], f = from[5], g = from[6], h = from[7];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t ch = (e & f) ^ (~e & g), maj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t1 = sha_add_fma(sha_add_fma(h, S1, one), sha_add_fma(ch, PVACB_SHA_K(i) + w[i], one), one);
        uint32_t t2 = sha_add_fma(S0, maj, one);
        h = g; g = f; f = e; e = sha_add_fma(d, t1, one); d = c; c = b; b = a; a = sha_add_fma(t1, t2, one);
    }
#pragma unroll 1
    for (int it = 16; it < 64; it += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            uint32_t w15 = w[(j + 1) & 15], w2 = w[(j + 14) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[j] = sha_add_fma(sha_add_fma(w[j], s0, one), sha_add_fma(w[(j + 9) & 15], s1, one), one);
            uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
            uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
            uint32_t ch = (e & f) ^ (~e & g), maj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t1 = sha_add_fma(sha_add_fma(h, S1, one), sha_add_fma(ch, c_shaK[it + j] + w[j], one), one);
            uint32_t t2 = sha_add_fma(S0, maj, one);
            h = g; g = f; f = e; e = sha_add_fma(d, t1, one); d = c; c = b; b = a; a = sha_add_fma(t1, t2, one);
        }
    }
    out[0] = from[0] + a; out[1] = from[1] + b; out[2] = from[2] + c; out[3] = from[3] + d;
    out[4] = from[4] + e; out[5] = from[5] + f; out[6] = from[6] + g; out[7] = from[7] + h;
}
// Fully rolled: four trips of (16 rounds, then the schedule of the next 16 words unless this was the last trip) -- ONE copy of
// the round code (~10 KB of SASS instead of ~14 KB), same dynamic instruction count.
// SPARSE: the caller guarantees w[2] = 0 and w[4..14] = 0 (block 1 of the counter hashes of prg_choose_k: tail of the salt, the
// counter, 0x80, zeros, the bit length). Words 16..31 of the schedule then need 15 small sigmas instead of 32:
//   W16 = s0(w1) + w0            W17 = s1(w15) + w1          W18 = s1(W16) + s0(w3)      W19 = s1(W17) + w3
//   W20 = s1(W18)                W21 = s1(W19)               W22 = s1(W20) + w15         W23..W29 = s1(W[t-2]) + W[t-7]
//   W30 = s1(W28) + W23 + s0(w15)                            W31 = s1(W29) + W24 + s0(W16) + w15
__device__ __forceinline__ uint32_t sha_s0(uint32_t x) { return sha_rotr(x, 7) ^ sha_rotr(x, 18) ^ (x >> 3); }
__device__ __forceinline__ uint32_t sha_s1(uint32_t x) { return sha_rotr(x, 17) ^ sha_rotr(x, 19) ^ (x >> 10); }
template <bool SPARSE = false>
__device__ __forceinline__ void sha_compress_from_rolled4(const uint32_t* from, uint32_t w[16], uint32_t out[8], uint32_t one) {
    uint32_t a = from[0], b = from[1], c = from[2], d = from[3], e = from[4], f = from[5], g = from[6], h = from[7];
#pragma unroll 1
    for (int it = 0; it < 64; it += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
            uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
            uint32_t ch = (e & f) ^ (~e & g), maj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t1 = sha_add_fma(sha_add_fma(h, S1, one), sha_add_fma(ch, c_shaK[it + j] + w[j], one), one);
            uint32_t t2 = sha_add_fma(S0, maj, one);
            h = g; g = f; f = e; e = sha_add_fma(d, t1, one); d = c; c = b; b = a; a = sha_add_fma(t1, t2, one);
        }
        if (SPARSE && it == 0) {
            const uint32_t w3 = w[3], w15 = w[15];
            w[0] = sha_add_fma(sha_s0(w[1]), w[0], one);                                   // W16
            w[1] = sha_add_fma(sha_s1(w15), w[1], one);                                    // W17
            w[2] = sha_add_fma(sha_s1(w[0]), sha_s0(w3), one);                             // W18
            w[3] = sha_add_fma(sha_s1(w[1]), w3, one);                                     // W19
            w[4] = sha_s1(w[2]);                                                           // W20
            w[5] = sha_s1(w[3]);                                                           // W21
            w[6] = sha_add_fma(sha_s1(w[4]), w15, one);                                    // W22
#pragma unroll
            for (int j = 7; j < 14; j++) w[j] = sha_add_fma(sha_s1(w[j - 2]), w[j - 7], one);   // W23..W29
            w[14] = sha_add_fma(sha_add_fma(sha_s1(w[12]), w[7], one), sha_s0(w15), one);  // W30
            w[15] = sha_add_fma(sha_add_fma(sha_s1(w[13]), w[8], one), sha_add_fma(sha_s0(w[0]), w15, one), one);   // W31
        } else if (it < 48) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint32_t w15 = w[(j + 1) & 15], w2 = w[(j + 14) & 15];
                uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
                w[j] = sha_add_fma(sha_add_fma(w[j], s0, one), sha_add_fma(w[(j + 9) & 15], s1, one), one);
            }
        }
    }
    out[0] = from[0] + a; out[1] = from[1] + b; out[2] = from[2] + c; out[3] = from[3] + d;
    out[4] = from[4] + e; out[5] = from[5] + f; out[6] = from[6] + g; out[7] = from[7] + h;
}
#endif

// LE64 of 8 big-endian digest bytes held as two state words (what load_le64(digest + 8k) returns)
PV_HD uint64_t sha_le64_of(uint32_t h_even, uint32_t h_odd) { return (uint64_t)sha_bswap(h_even) | ((uint64_t)sha_bswap(h_odd) << 32); }

// 8 little-endian stream words -> 16 big-endian schedule words
PV_HD void sha_block_from_le64(const uint64_t q[8], uint32_t w[16]) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        w[2 * j] = sha_bswap((uint32_t)q[j]);
        w[2 * j + 1] = sha_bswap((uint32_t)(q[j] >> 32));
    }
}

// digest word k (k = 0..3) as the reference reads it: load_le64(out + 8k) of the big-endian digest bytes
PV_HD uint64_t sha_digest_le64(const ShaState& s, int k) {
    return (uint64_t)sha_bswap(s.h[2 * k]) | ((uint64_t)sha_bswap(s.h[2 * k + 1]) << 32);
}

// A label of L bytes (8 <= L < 16) followed by u64s: stream word j of the message.
//   q[0] = label bytes 0..7 ; q[1] = label bytes 8..L-1 | x0 << 8r ; q[j] = x[j-2] >> (64-8r) | x[j-1] << 8r  (r = L-8)
struct LabelStream {
    uint64_t l0, l1;  // label bytes 0..7 and 8..L-1 (little-endian packed)
    int r;            // L - 8, in 1..7
};
PV_HD uint64_t stream_word(const LabelStream& ls, const uint64_t* x, int nx, int j, uint64_t tail_pad) {
    // tail_pad: value appended after x[nx-1] (0x80 then zeros), as a virtual x[nx]
    if (j == 0) return ls.l0;
    int sh = 8 * ls.r;
    uint64_t prev = (j == 1) ? ls.l1 : ((j - 2 < nx ? x[j - 2] : (j - 2 == nx ? tail_pad : 0ull)) >> (64 - sh));
    uint64_t cur = (j - 1 < nx) ? x[j - 1] : (j - 1 == nx ? tail_pad : 0ull);
    return prev | (cur << sh);
}

constexpr uint64_t pack_label(const char* s, int from, int to) {
    uint64_t v = 0;
    for (int i = from; i < to; i++) v |= (uint64_t)(uint8_t)s[i] << (8 * (i - from));
    return v;
}
// "pvac.dom.x_seed" (15), "pvac.dom.noise" (14), "pvac.dom.h_gen" (14), "pvac.dom.ztag" (13)  (core/types.hpp:14-32)
constexpr uint64_t kLXSeed0 = pack_label("pvac.dom.x_seed", 0, 8), kLXSeed1 = pack_label("pvac.dom.x_seed", 8, 15);
constexpr uint64_t kLNoise0 = pack_label("pvac.dom.noise", 0, 8), kLNoise1 = pack_label("pvac.dom.noise", 8, 14);
constexpr uint64_t kLHGen0 = pack_label("pvac.dom.h_gen", 0, 8), kLHGen1 = pack_label("pvac.dom.h_gen", 8, 14);
constexpr uint64_t kLZtag0 = pack_label("pvac.dom.ztag", 0, 8), kLZtag1 = pack_label("pvac.dom.ztag", 8, 13);
PV_HD LabelStream label_xseed() { LabelStream l; l.l0 = kLXSeed0; l.l1 = kLXSeed1; l.r = 7; return l; }
PV_HD LabelStream label_noise() { LabelStream l; l.l0 = kLNoise0; l.l1 = kLNoise1; l.r = 6; return l; }
PV_HD LabelStream label_hgen() { LabelStream l; l.l0 = kLHGen0; l.l1 = kLHGen1; l.r = 6; return l; }
PV_HD LabelStream label_ztag() { LabelStream l; l.l0 = kLZtag0; l.l1 = kLZtag1; l.r = 5; return l; }

// Generic (not hot) one-shot: SHA-256(label || LE64(x[0..nx-1])) for messages of at most 2 blocks (<= 119 bytes).
PV_HD void sha_label_words(const LabelStream& ls, const uint64_t* x, int nx, ShaState& st) {
    int len = 8 + ls.r + 8 * nx;  // bytes
    sha_init(st);
    uint64_t q[16];
    // the 0x80 terminator sits right after the last x word: as a virtual extra word 0x80
    for (int j = 0; j < 16; j++) q[j] = stream_word(ls, x, nx, j, 0x80ull);
    int nblocks = (len + 1 + 8 + 63) / 64;
    // length field: big-endian 64-bit bit count in the last 8 bytes = last stream word, byte-swapped
    uint64_t bits = (uint64_t)len * 8;
    uint64_t be = 0;
    for (int i = 0; i < 8; i++) be |= ((bits >> (8 * i)) & 0xff) << (8 * (7 - i));
    q[nblocks * 8 - 1] = be;
    uint32_t w[16];
    for (int b = 0; b < nblocks; b++) {
        sha_block_from_le64(q + 8 * b, w);
        sha_compress(st, w);
    }
}

// crypto/matrix.hpp:254-264 : LE64 of the first 8 digest bytes of SHA-256("pvac.dom.ztag" || canon || nonce.lo || nonce.hi)
PV_HD uint64_t prg_layer_ztag(uint64_t canon_tag, uint64_t nlo, uint64_t nhi) {
    uint64_t x[3] = {canon_tag, nlo, nhi};
    ShaState st;
    sha_label_words(label_ztag(), x, 3, st);
    return sha_digest_le64(st, 0);
}

}  // namespace pvacb
