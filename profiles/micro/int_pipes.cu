// Integer issue-rate probe for the B200 SM ("measure the peak integer issue rate once"): dependent-free chains of
// LOP3, SHF, PRMT (ALU pipe), IMAD (FMA pipe) and a 1:1 mix, 8 independent accumulators per thread, full occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 int_pipes.cu -o int_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 2654435761u + i * a;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == 1) asm volatile("shf.l.wrap.b32 %0, %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(x[i]) : "r"(b));
                if (OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == 5) {
                    if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[i];
    if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, uint32_t* out) {
    const int iters = 2000, grid = 148 * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<OP><<<grid, block>>>(out, 0x9E3779B9u, 0x85EBCA6Bu, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)grid * block * iters * 16 * 8;
    double per_clk_sm = ops / (best * 1e-3) / 148 / 1.965e9;
    printf("%-22s %.3f ms  %.2f T lane-ops/s  %.1f lanes/clk/SM (at 1965 MHz)\n", name, best, ops / best / 1e9, per_clk_sm);
}

int main() {
    uint32_t* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("LOP3", out); run<1>("SHF", out); run<3>("PRMT", out); run<4>("IMAD", out); run<5>("LOP3 + IMAD 1:1", out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
