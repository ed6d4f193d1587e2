// micro-benchmark: dependent-chain latency of one SGM step (one warp, clock64)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define SG_BIG2 0x75307530u
__device__ __forceinline__ uint32_t bcast16(int v) { return (uint32_t)(v & 0xffff) * 0x00010001u; }
template <int MODE>
__device__ __forceinline__ void step(uint32_t& a0, uint32_t& a1, uint32_t& mm, uint32_t c0, uint32_t c1, uint32_t P1b, uint32_t P2b, int src_up) {
    const uint32_t up = __shfl_sync(0xffffffffu, a1, src_up);
    const uint32_t dn = __shfl_down_sync(0xffffffffu, a0, 1);
    const uint32_t lm0 = __byte_perm(up, a0, 0x5432), mid = __byte_perm(a0, a1, 0x5432), lp1 = __byte_perm(a1, dn, 0x5432);
    const uint32_t mp2 = mm + P2b;
    const uint32_t t0 = __vmins2(__vmins2(__vadd2(__vmins2(lm0, mid), P1b), mp2), a0);
    const uint32_t t1 = __vmins2(__vmins2(__vadd2(__vmins2(mid, lp1), P1b), mp2), a1);
    a0 = c0 + t0 - mm; a1 = c1 + t1 - mm;
    const uint32_t w = __vmins2(a0, a1);
    uint32_t v = __vmins2(w, __byte_perm(w, w, 0x1032));
    if (MODE == 0) mm = (uint32_t)__reduce_min_sync(0xffffffffu, (int)v);
    else if (MODE == 1) { for (int o = 16; o; o >>= 1) v = __vmins2(v, __shfl_xor_sync(0xffffffffu, v, o)); mm = v; }
    else if (MODE == 2) mm = v & 0x00010001u;  // no reduction (chain through ALU only)
    else { // MODE 3: 3-level: 8-lane groups via 2 xor-shuffles? (24 lanes) -> use redux on half
        v = __vmins2(v, __shfl_xor_sync(0xffffffffu, v, 16)); mm = (uint32_t)__reduce_min_sync(0xffffffffu, (int)v); }
}
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int n) {
    const int lane = threadIdx.x & 31; const int src_up = (lane + 31) & 31;
    uint32_t a0 = lane < 24 ? 0u : SG_BIG2, a1 = a0, mm = 0; const uint32_t P1b = bcast16(648), P2b = bcast16(2592);
    uint32_t c0 = (lane * 37 + 11) & 0x03ff03ff, c1 = (lane * 91 + 5) & 0x03ff03ff;
    if (lane >= 24) c0 = c1 = 0x6d606d60u;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { step<MODE>(a0, a1, mm, c0, c1, P1b, P2b, src_up); c0 ^= 0x00010001u; }
    long long t1 = clock64();
    out[threadIdx.x] = a0 + a1 + mm; if (threadIdx.x == 0) cyc[MODE] = t1 - t0;
}
__global__ void kshfl(uint32_t* out, long long* cyc, int n) {
    uint32_t v = threadIdx.x; long long t0 = clock64();
    for (int i = 0; i < n; ++i) v = __shfl_down_sync(0xffffffffu, v, 1) + 1;
    long long t1 = clock64(); out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[4] = t1 - t0;
}
__global__ void kredux(uint32_t* out, long long* cyc, int n) {
    uint32_t v = threadIdx.x; long long t0 = clock64();
    for (int i = 0; i < n; ++i) v = (uint32_t)__reduce_min_sync(0xffffffffu, (int)v) + threadIdx.x;
    long long t1 = clock64(); out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[5] = t1 - t0;
}
int main() {
    uint32_t* o; long long* c; cudaMalloc(&o, 4096); cudaMallocManaged(&c, 64); const int n = 4096;
    for (int r = 0; r < 2; ++r) { k<0><<<1, 32>>>(o, c, n); k<1><<<1, 32>>>(o, c, n); k<2><<<1, 32>>>(o, c, n); k<3><<<1, 32>>>(o, c, n); kshfl<<<1, 32>>>(o, c, n); kredux<<<1, 32>>>(o, c, n); cudaDeviceSynchronize(); }
    printf("cycles/step: credux %.1f  shfl-butterfly %.1f  no-reduce %.1f  shfl+credux %.1f | shfl chain %.1f  credux chain %.1f\n",
           c[0] / (double)n, c[1] / (double)n, c[2] / (double)n, c[3] / (double)n, c[4] / (double)n, c[5] / (double)n);
    return 0;
}
