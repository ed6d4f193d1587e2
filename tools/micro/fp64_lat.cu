// fp64 latency / throughput probe: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_div(double* o, double x, int n, long long* cyc) {
    double a = x + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = 1.0 / a + 1.5;
    long long t1 = clock64();
    o[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_fma(double* o, double x, int n, long long* cyc) {
    double a = x + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = a * 1.0000001 + 0.5;
    long long t1 = clock64();
    o[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_fma_tp(double* o, double x, int n, long long* cyc) {  // 8 independent chains per thread, 256 threads
    double a[8];
    for (int j = 0; j < 8; ++j) a[j] = x + threadIdx.x + j;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = a[j] * 1.0000001 + 0.5;
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 8; ++j) s += a[j];
    o[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sqrt(double* o, double x, int n, long long* cyc) {
    double a = x + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = sqrt(a) + 2.0;
    long long t1 = clock64();
    o[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_bar(double* o, int n, long long* cyc) {
    __shared__ double s[256];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    double a = 0;
    for (int i = 0; i < n; ++i) {
        a += s[(threadIdx.x + i) & 255];
        __syncthreads();
    }
    long long t1 = clock64();
    o[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(double* o, double x, int n, long long* cyc) {
    double a = x + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = __shfl_sync(0xffffffffu, a, (i * 7) & 31) * 0.999 + 1.0;
    long long t1 = clock64();
    o[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 4096); cudaMallocManaged(&c, 8);
    const int n = 1000;
    k_div<<<1, 32>>>(o, 3.0, n, c); cudaDeviceSynchronize(); printf("1/x + add chain      : %.1f cycles per step\n", (double)*c / n);
    k_fma<<<1, 32>>>(o, 3.0, n, c); cudaDeviceSynchronize(); printf("DFMA chain           : %.1f cycles per step\n", (double)*c / n);
    k_fma_tp<<<1, 256>>>(o, 3.0, n, c); cudaDeviceSynchronize(); printf("DFMA 8 warps x 8 ILP : %.2f cycles per warp-DFMA per SM (=> %.1f DFMA lanes/clk/SM)\n", (double)*c / (n * 64.0), 32.0 * n * 64.0 / (double)*c);
    k_sqrt<<<1, 32>>>(o, 3.0, n, c); cudaDeviceSynchronize(); printf("sqrt + add chain     : %.1f cycles per step\n", (double)*c / n);
    k_bar<<<1, 256>>>(o, n, c); cudaDeviceSynchronize(); printf("LDS + __syncthreads  : %.1f cycles per step (256 threads)\n", (double)*c / n);
    k_shfl<<<1, 32>>>(o, 3.0, n, c); cudaDeviceSynchronize(); printf("shfl(double) + DFMA  : %.1f cycles per step\n", (double)*c / n);
    return 0;
}
