// Micro-benchmark: SHA-256 compression variants on the B200 integer pipes (tuning aid for csrc/sigma.cu phase B).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../pvac_hfhe_cppbyv_b200/csrc sha_variants.cu -o sha_variants
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sha256.cuh"
using namespace pvacb;

__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one) {   // a + b on the FMA pipe (IMAD)
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}
__device__ __forceinline__ void mulwide(uint32_t x, uint32_t m, uint32_t& lo, uint32_t& hi) {
    uint64_t r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x), "r"(m));
    lo = (uint32_t)r; hi = (uint32_t)(r >> 32);
}

template <int V>
__device__ __forceinline__ void compress(const uint32_t* from, uint32_t w[16], uint32_t out[8], uint32_t one) {
    PVACB_SHA_K_DECL;
    uint32_t a = from[0], b = from[1], c = from[2], d = from[3], e = from[4], f = from[5], g = from[6], h = from[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0, s1;
            if (V & 2) {   // rotations through 32x32->64 multiplies (FMA pipe) + 5-input XOR (2 LOP3)
                uint32_t l1, h1, l2, h2, l3, h3;
                mulwide(w15, 1u << 25, l1, h1); mulwide(w15, 1u << 14, l2, h2); mulwide(w15, 1u << 29, l3, h3);
                s0 = (l1 ^ h1 ^ l2) ^ (h2 ^ h3);
                mulwide(w2, 1u << 15, l1, h1); mulwide(w2, 1u << 13, l2, h2); mulwide(w2, 1u << 22, l3, h3);
                s1 = (l1 ^ h1 ^ l2) ^ (h2 ^ h3);
            } else {
                s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            }
            if (V & 1) w[i & 15] = add_fma(add_fma(w[i & 15], s0, one), add_fma(w[(i - 7) & 15], s1, one), one);
            else w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
        }
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t ch = (e & f) ^ (~e & g), maj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t1, t2;
        if (V & 1) {
            t1 = add_fma(add_fma(h, S1, one), add_fma(ch, PVACB_SHA_K(i) + w[i & 15], one), one);
            t2 = add_fma(S0, maj, one);
            h = g; g = f; f = e; e = add_fma(d, t1, one); d = c; c = b; b = a; a = add_fma(t1, t2, one);
        } else {
            t1 = h + S1 + ch + PVACB_SHA_K(i) + w[i & 15];
            t2 = S0 + maj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
    }
    out[0] = from[0] + a; out[1] = from[1] + b; out[2] = from[2] + c; out[3] = from[3] + d;
    out[4] = from[4] + e; out[5] = from[5] + f; out[6] = from[6] + g; out[7] = from[7] + h;
}

template <int V>
__global__ void __launch_bounds__(128, 8) k(const uint32_t* mid, uint32_t* sink, int iters, uint32_t one) {
    __shared__ uint32_t m[8];
    if (threadIdx.x < 8) m[threadIdx.x] = mid[threadIdx.x];
    __syncthreads();
    uint32_t acc = 0;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; it++) {
        uint32_t w[16];
        w[0] = t; w[1] = 0x12340000u | it; w[2] = 0; w[3] = 0x80;
#pragma unroll
        for (int i = 4; i < 15; i++) w[i] = 0;
        w[15] = 632;
        uint32_t d[8];
        compress<V>(m, w, d, one);
        acc ^= d[0] ^ d[1] ^ d[2] ^ d[3] ^ d[4] ^ d[5] ^ d[6] ^ d[7];
    }
    if (acc == 0x12345678u) sink[t] = acc;
}

template <int V>
void run(const char* name, const uint32_t* mid, uint32_t* sink) {
    const int iters = 256, grid = 148 * 8, block = 128;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        k<V><<<grid, block>>>(mid, sink, iters, 1u);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    double n = (double)grid * block * iters;
    printf("%-28s %.3f ms  %.2f G compressions/s  (%s)\n", name, best, n / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t *mid, *sink;
    cudaMalloc(&mid, 32); cudaMemset(mid, 0x5a, 32);
    cudaMalloc(&sink, 148 * 8 * 128 * 4);
    run<0>("baseline (SHF, IADD3)", mid, sink);
    run<1>("adds as IMAD", mid, sink);
    run<2>("sigma rot via mul.wide", mid, sink);
    run<3>("both", mid, sink);
    return 0;
}
