// SM clock seen by a short kernel: cycles / globaltimer ns, cold (after a sleep) and right after a heavy kernel
#include <cstdio>
#include <unistd.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void probe(double* o, int n, double* mhz) {
    double a = threadIdx.x;
    unsigned long long g0 = gt(); long long c0 = clock64();
    for (int i = 0; i < n; ++i) a = a * 1.0000001 + 0.5;
    long long c1 = clock64(); unsigned long long g1 = gt();
    o[threadIdx.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *mhz = (double)(c1 - c0) / (double)(g1 - g0) * 1e3;
}
__global__ void heavy(float* o, int n) {
    float a = threadIdx.x;
    for (int i = 0; i < n; ++i) a = a * 1.0001f + 0.5f;
    o[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
int main() {
    double* o; double* m; float* f; cudaMalloc(&o, 1 << 20); cudaMallocManaged(&m, 8); cudaMalloc(&f, 148 * 8 * 1024 * 4);
    for (int rep = 0; rep < 3; ++rep) {
        usleep(200000);
        probe<<<1, 256>>>(o, 50000, m); cudaDeviceSynchronize(); printf("after 200 ms idle, 1 CTA  : %.0f MHz\n", *m);
        probe<<<148, 256>>>(o, 50000, m); cudaDeviceSynchronize(); printf("back to back, 148 CTAs    : %.0f MHz\n", *m);
        heavy<<<148 * 8, 1024>>>(f, 2000000); probe<<<148, 256>>>(o, 50000, m); cudaDeviceSynchronize(); printf("after a heavy kernel      : %.0f MHz\n", *m);
    }
    return 0;
}
