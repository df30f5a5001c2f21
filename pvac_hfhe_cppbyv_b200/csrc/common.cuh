// Shared definitions for the pvacb engine (host + device).
#pragma once
#include <cstdint>
#include <cstddef>

#if defined(__CUDACC__)
#define PV_HD __host__ __device__ __forceinline__
#define PV_D __device__ __forceinline__
#else
#define PV_HD inline
#define PV_D inline
#endif

namespace pvacb {

// default Params of the reference (core/types.hpp:36-70); the engine is specialised to these shapes
constexpr int kB = 337;            // carrier group order
constexpr int kMBits = 8192;       // sigma length
constexpr int kMWords = 128;       // sigma words
constexpr int kNBits = 16384;      // columns of H
constexpr int kHColWt = 192;
constexpr int kXColWt = 128;
constexpr int kErrWt = 128;
constexpr int kLpnN = 4096;
constexpr int kLpnWords = 64;
constexpr int kLpnT = 16384;
constexpr int kSignal = 8;         // "S" signal edges per share (ops/encrypt.hpp:172)
constexpr uint32_t kEdgeBudget = 1200000;

constexpr uint64_t kMask63 = 0x7FFFFFFFFFFFFFFFull;

// ---- RNG tape: replaces csprng_u64 (core/random.hpp:106-110). Word k (k = 0,1,..) of an item's stream, random access.
//   TAPE_CHACHA20 (the default of a context): word k = 64-bit little-endian word (k mod 8) of ChaCha20 block (k div 8) under the
//       context's 256-bit tape key; block input words 12..15 = (block index, lane, stream id lo, stream id hi). A stream derived
//       from a batch_seed has stream id = batch_seed and lane = global item index + 1; an explicit per-item stream id has lane 0.
//       Nothing but keystream ever reaches a ciphertext, and two items never share a stream unless the caller repeats a seed.
//   TAPE_SPLITMIX: counter-based SplitMix64 of one 64-bit state. NOT secret-keeping (mix64 is a bijection and nonces are tape
//       words): parity tests and golden vectors only.
//   TAPE_WORDS: the caller supplies every word (e.g. from its own CSPRNG, or a test that needs particular draws).
enum : int { TAPE_SPLITMIX = 0, TAPE_CHACHA20 = 1, TAPE_WORDS = 2 };

PV_HD uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
PV_HD uint64_t splitmix_word(uint64_t s0, uint64_t k) { return mix64(s0 + (k + 1) * 0x9E3779B97F4A7C15ull); }
PV_HD uint64_t item_stream_state(uint64_t batch_seed, uint64_t item) {
    return mix64(batch_seed + 0xD1342543DE82EF95ull * (item + 1));
}

PV_HD uint32_t rotl32(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(x, x, n);
#else
    return (x << n) | (x >> (32 - n));
#endif
}
#define PV_CHACHA_QR(a, b, c, d) \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); a += b; d ^= a; d = rotl32(d, 8); c += d; b ^= c; b = rotl32(b, 7);
// the ChaCha20 block function (20 rounds): key words k[8] and input words 12..15 = c[4], all little-endian; out = 8 64-bit words
PV_HD void chacha20_block(const uint32_t k[8], uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t out[8]) {
    const uint32_t i0 = 0x61707865u, i1 = 0x3320646eu, i2 = 0x79622d32u, i3 = 0x6b206574u;
    uint32_t x0 = i0, x1 = i1, x2 = i2, x3 = i3, x4 = k[0], x5 = k[1], x6 = k[2], x7 = k[3], x8 = k[4], x9 = k[5], x10 = k[6], x11 = k[7],
             x12 = c0, x13 = c1, x14 = c2, x15 = c3;
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        PV_CHACHA_QR(x0, x4, x8, x12) PV_CHACHA_QR(x1, x5, x9, x13) PV_CHACHA_QR(x2, x6, x10, x14) PV_CHACHA_QR(x3, x7, x11, x15)
        PV_CHACHA_QR(x0, x5, x10, x15) PV_CHACHA_QR(x1, x6, x11, x12) PV_CHACHA_QR(x2, x7, x8, x13) PV_CHACHA_QR(x3, x4, x9, x14)
    }
    out[0] = (uint64_t)(x0 + i0) | ((uint64_t)(x1 + i1) << 32);
    out[1] = (uint64_t)(x2 + i2) | ((uint64_t)(x3 + i3) << 32);
    out[2] = (uint64_t)(x4 + k[0]) | ((uint64_t)(x5 + k[1]) << 32);
    out[3] = (uint64_t)(x6 + k[2]) | ((uint64_t)(x7 + k[3]) << 32);
    out[4] = (uint64_t)(x8 + k[4]) | ((uint64_t)(x9 + k[5]) << 32);
    out[5] = (uint64_t)(x10 + k[6]) | ((uint64_t)(x11 + k[7]) << 32);
    out[6] = (uint64_t)(x12 + c0) | ((uint64_t)(x13 + c1) << 32);
    out[7] = (uint64_t)(x14 + c2) | ((uint64_t)(x15 + c3) << 32);
}

// how the items of one call find their streams; passed by value to kernels
struct TapeSpec {
    int kind = TAPE_SPLITMIX;
    uint32_t key[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // TAPE_CHACHA20
    uint64_t batch_seed = 0;
    uint64_t item_base = 0;            // global index of item 0 of this call (a shard of a larger batch keeps the global streams)
    const uint64_t* states = nullptr;  // explicit 64-bit stream id per item (device or host pointer, as the code using it runs), or null
    const uint64_t* k0 = nullptr;      // first word of each item's stream to use (continuing a stream across calls), or null = 0
    const uint64_t* ids = nullptr;     // global number of each item, or null = item_base + i (a sub-batch re-planned out of a larger call)
    const uint64_t* words = nullptr;   // TAPE_WORDS: [items][words_per_item]
    uint64_t words_per_item = 0;
};

struct Tape {
    int kind;
    const uint32_t* key;
    uint64_t s0;          // SPLITMIX: state; CHACHA20: stream id
    uint32_t lane;
    const uint64_t* words;
    uint64_t nwords;
    uint64_t k;           // next word
    uint64_t blk;         // CHACHA20: index of the cached block (~0 = none)
    uint64_t buf[8];
    bool overrun;         // TAPE_WORDS: a word beyond the supplied ones was asked for
    PV_HD uint64_t at(uint64_t kk) {
        if (kind == TAPE_SPLITMIX) return splitmix_word(s0, kk);
        if (kind == TAPE_WORDS) {
            // past the supplied words: remember it (the call then fails) and hand out varying filler, so that the rejection loops of the
            // walk (distinct indices, non-zero field elements) still terminate
            if (kk >= nwords) { overrun = true; return splitmix_word(0x6F76657272756E21ull, kk); }
            return words[kk];
        }
        const uint64_t b = kk >> 3;
        if (b != blk) {
            chacha20_block(key, (uint32_t)b, lane ^ (uint32_t)(b >> 32), (uint32_t)s0, (uint32_t)(s0 >> 32), buf);
            blk = b;
        }
        return buf[kk & 7];
    }
    PV_HD uint64_t next() { return at(k++); }
};

// the stream of item i of a call
PV_HD Tape tape_open(const TapeSpec& ts, uint64_t i) {
    Tape t;
    t.kind = ts.kind; t.key = ts.key; t.words = nullptr; t.nwords = 0; t.lane = 0; t.blk = ~0ull; t.overrun = false;
    t.k = ts.k0 ? ts.k0[i] : 0;
    const uint64_t g = ts.ids ? ts.ids[i] : ts.item_base + i;
    if (ts.kind == TAPE_WORDS) { t.words = ts.words + g * ts.words_per_item; t.nwords = ts.words_per_item; t.s0 = 0; }
    else if (ts.states) t.s0 = ts.states[i];
    else if (ts.kind == TAPE_SPLITMIX) t.s0 = item_stream_state(ts.batch_seed, g);
    else { t.s0 = ts.batch_seed; t.lane = (uint32_t)(g + 1); }
    return t;
}
// a bare SplitMix64 stream (key generation under the legacy 64-bit seed, synthetic benchmark data)
PV_HD Tape tape_splitmix(uint64_t s0) {
    Tape t;
    t.kind = TAPE_SPLITMIX; t.key = nullptr; t.words = nullptr; t.nwords = 0; t.lane = 0; t.blk = ~0ull; t.overrun = false; t.k = 0; t.s0 = s0;
    return t;
}

// FNV-1a of the PRF domain strings (crypto/lpn.hpp:157-164), precomputed; checked in tests against the oracle
constexpr uint64_t fnv1a_const(const char* s) {
    uint64_t h = 0xcbf29ce484222325ull;
    while (*s) { h ^= (uint64_t)(uint8_t)*s++; h *= 0x100000001b3ull; }
    return h;
}
constexpr uint64_t kFnvR1 = fnv1a_const("pvac.prf.r.1"), kFnvR2 = fnv1a_const("pvac.prf.r.2"), kFnvR3 = fnv1a_const("pvac.prf.r.3");
constexpr uint64_t kFnvN1 = fnv1a_const("pvac.prf.noise.1"), kFnvN2 = fnv1a_const("pvac.prf.noise.2"), kFnvN3 = fnv1a_const("pvac.prf.noise.3");
constexpr uint64_t kFnvToep = fnv1a_const("pvac.dom.toeplitz");
// family 0: Dom::PRF_R1..3 (prf_R), family 1: Dom::PRF_NOISE1..3 (prf_R_noise); crypto/lpn.hpp:263-275
PV_HD uint64_t fnv_prf_dom(int family, int t) {
    if (family) return t == 0 ? kFnvN1 : (t == 1 ? kFnvN2 : kFnvN3);
    return t == 0 ? kFnvR1 : (t == 1 ? kFnvR2 : kFnvR3);
}

}  // namespace pvacb
