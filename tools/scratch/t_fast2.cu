#include <cstdio>
#include <cstdint>
__global__ void k(const int* din, int* out) {
    int d[16];
    for (int i = 0; i < 16; ++i) d[i] = din[threadIdx.x * 16 + i];
    int mn[16], mx[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn[k] = min(d[k], d[(k + 1) & 15]); mx[k] = max(d[k], d[(k + 1) & 15]); }
    int mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn2[k] = min(mn[k], mn[(k + 2) & 15]); mx2[k] = max(mx[k], mx[(k + 2) & 15]); }
    int besta = -1000, bestb = -1000, best = -1000;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int a = min(min(mn2[k], mn2[(k + 4) & 15]), d[(k + 8) & 15]);
        const int b = max(max(mx2[k], mx2[(k + 4) & 15]), d[(k + 8) & 15]);
        besta = max(besta, a); bestb = max(bestb, -b);
        best = max(best, max(a, -b));
        out[threadIdx.x * 64 + 3 + k] = a; out[threadIdx.x * 64 + 19 + k] = b;
    }
    out[threadIdx.x * 64] = best; out[threadIdx.x * 64 + 1] = besta; out[threadIdx.x * 64 + 2] = bestb;
}
int main() {
    int h[32] = {-55,-55,-55,-55,51,31,34,41,34,37,39,27,22,-55,-55,-55, -10,-8,-20,-18,-3,2,-5,-41,-61,-84,-84,-84,-84,-84,-84,-84};
    int *d, *o; cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, 512); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    k<<<1, 2>>>(d, o); int r[128]; cudaMemcpy(r, o, 512, cudaMemcpyDeviceToHost);
    for (int t = 0; t < 2; ++t) { printf("best=%d besta=%d bestb=%d\n a:", r[t*64], r[t*64+1], r[t*64+2]); for (int k=0;k<16;++k) printf(" %d", r[t*64+3+k]); printf("\n b:"); for (int k=0;k<16;++k) printf(" %d", r[t*64+19+k]); printf("\n"); }
}
