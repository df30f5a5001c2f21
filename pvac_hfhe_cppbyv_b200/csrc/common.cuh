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

// ---- RNG tape: SplitMix64 as a counter-based generator (random access), replacing csprng_u64
// (core/random.hpp:106-110). word k (k = 0,1,..) of the stream with initial state s0.
PV_HD uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
PV_HD uint64_t tape_word(uint64_t s0, uint64_t k) { return mix64(s0 + (k + 1) * 0x9E3779B97F4A7C15ull); }
PV_HD uint64_t item_stream_state(uint64_t batch_seed, uint64_t item) {
    return mix64(batch_seed + 0xD1342543DE82EF95ull * (item + 1));
}
struct Tape {
    uint64_t s0;
    uint64_t k;
    PV_HD uint64_t next() { return tape_word(s0, k++); }
};

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
