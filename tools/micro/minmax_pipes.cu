// Micro-benchmark: can the 16x2 min/max network of fast_kernel be split between the integer ALU pipe
// (VIMNMX / VIMNMX3 .S16x2) and the FP16 pipe (HMNMX2)?  Pixel values stored as 0x6400 | v are ordered identically as
// int16 and as fp16 (1024 + v), so either instruction can process the same words.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o minmax_pipes minmax_pipes.cu && ./minmax_pipes
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t hmax2u(uint32_t a, uint32_t b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t hmin2u(uint32_t a, uint32_t b) {
    __half2 r = __hmin2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = 0x64006400u | ((threadIdx.x * 7 + i * 13 + seed) & 0xFF) | (((threadIdx.x * 3 + i * 5) & 0xFF) << 16);
        b[i] = 0x64006400u | ((threadIdx.x * 11 + i * 17 + seed) & 0xFF) | (((threadIdx.x * 5 + i * 3) & 0xFF) << 16);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {  // 2 integer ops
                a[i] = __vmaxs2(a[i], b[(i + 1) & 7]);
                b[i] = __vmins2(b[i], a[(i + 3) & 7]);
            } else if (MODE == 1) {  // 2 fp16 ops
                a[i] = hmax2u(a[i], b[(i + 1) & 7]);
                b[i] = hmin2u(b[i], a[(i + 3) & 7]);
            } else if (MODE == 2) {  // 1 + 1
                a[i] = __vmaxs2(a[i], b[(i + 1) & 7]);
                b[i] = hmin2u(b[i], a[(i + 3) & 7]);
            } else if (MODE == 3) {  // 3-input integer ops
                a[i] = __vimax3_s16x2(a[i], b[(i + 1) & 7], b[(i + 2) & 7]);
                b[i] = __vimin3_s16x2(b[i], a[(i + 3) & 7], a[(i + 5) & 7]);
            } else {  // 3-input integer + fp16
                a[i] = __vimax3_s16x2(a[i], b[(i + 1) & 7], b[(i + 2) & 7]);
                b[i] = hmin2u(b[i], a[(i + 3) & 7]);
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, uint32_t* d) {
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 64, 1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * 16;  // thread-level min/max instructions
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %.3f ms  %.1f warp-instr/clk/SM (at %d MHz)\n", name, ms, ops / 32 / (ms * 1e-3) / (clk * 1e3) / 148, clk / 1000);
}

int main() {
    uint32_t* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("VIMNMX.S16x2 only", d);
    run<1>("HMNMX2 only", d);
    run<2>("VIMNMX + HMNMX2 1:1", d);
    run<3>("VIMNMX3.S16x2 only", d);
    run<4>("VIMNMX3 + HMNMX2 1:1", d);
    uint32_t h[4]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("check %08x\n", h[0]);
    return 0;
}
