// Stand-alone probe of the small LDL^T solver (register-resident, 2-D cyclic): one CTA, timing with clock64.
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define BA_SMALL_N 96
template <int SB_A, int SB_B>
__device__ bool solve_t(int n, const double* S, const double* rhs, double* M, double* xs, long long* cyc) {
    const int ld = n + 1, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    __shared__ int s_fail;
    __shared__ double s_rd[BA_SMALL_N];
    if (tid == 0) s_fail = 0;
    long long t0 = clock64();
    double v[SB_A][SB_B];
#pragma unroll
    for (int a = 0; a < SB_A; ++a) {
        const int r = warp + 8 * a;
#pragma unroll
        for (int b = 0; b < SB_B; ++b) {
            const int c = lane + 32 * b;
            double x = 0.0;
            if (r < n && c >= r && c < n) x = __ldcg(S + r * n + c);
            if (r < n && c == n) x = __ldcg(rhs + r);
            v[a][b] = x;
        }
    }
    if (warp == 0) {
#pragma unroll
        for (int b = 0; b < SB_B; ++b) {
            const int c = lane + 32 * b;
            if (c <= n) M[c] = v[0][b];
        }
        if (lane == 0) {
            const double d = v[0][0];
            if (!(d > 0.0) || !isfinite(d)) s_fail = 1;
            s_rd[0] = 1.0 / d;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    for (int k = 0; k < n - 1; ++k) {
        const double* Mk = M + k * ld;
        const double rk = s_rd[k];
        double mk[SB_B], prow[SB_B];
#pragma unroll
        for (int b = 0; b < SB_B; ++b) {
            mk[b] = Mk[min(lane + 32 * b, n)];
            prow[b] = 0.0;
        }
        const int a1 = (k + 1 - warp + 7) >> 3;
#pragma unroll
        for (int a = 0; a < SB_A; ++a) {
            const int r = warp + 8 * a;
#ifdef V_NOUPD
            if (a >= a1 && r < n && r == k + 1) {
#else
            if (a >= a1 && r < n) {
#endif
                const double f = Mk[r] * rk;
#pragma unroll
                for (int b = 0; b < SB_B; ++b) v[a][b] -= f * mk[b];
                if (r == k + 1) {
#pragma unroll
                    for (int b = 0; b < SB_B; ++b) prow[b] = v[a][b];
                }
            }
        }
        if (warp == ((k + 1) & 7)) {
            double d = 0.0;
#pragma unroll
            for (int b = 0; b < SB_B; ++b) {
                const int c = lane + 32 * b;
                if (c > k && c <= n) M[(k + 1) * ld + c] = prow[b];
                if (c == k + 1) d = prow[b];
            }
            if (lane == ((k + 1) & 31)) {
#ifndef V_NOFAIL
                if (!(d > 0.0) || !isfinite(d)) s_fail = 1;
#endif
#ifdef V_NODIV
                s_rd[k + 1] = d * 1e-3;
#else
                s_rd[k + 1] = 1.0 / d;
#endif
            }
        }
        __syncthreads();
#ifndef V_NOFAIL
        if (s_fail) break;
#endif
    }
    long long t2 = clock64();
    const bool fail = s_fail != 0;
    if (!fail && warp == 0) {
        constexpr int Q = (SB_A * 8 + 31) / 32;
        double z[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) z[q] = (lane + 32 * q < n) ? M[(lane + 32 * q) * ld + n] : 0.0;
#pragma unroll
        for (int q = Q - 1; q >= 0; --q) {
#pragma unroll 4
            for (int kl = 31; kl >= 0; --kl) {
                const int k = 32 * q + kl;
                if (k >= n) continue;
                const double xk = __shfl_sync(0xFFFFFFFFu, z[q], kl) * s_rd[k];
                if (lane == kl) z[q] = xk;
#pragma unroll
                for (int p = 0; p <= q; ++p) {
                    const int i = lane + 32 * p;
                    if (i < k) z[p] -= M[i * ld + k] * xk;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q)
            if (lane + 32 * q < n) xs[lane + 32 * q] = z[q];
    }
    __syncthreads();
    long long t3 = clock64();
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    return !fail;
}
__global__ void __launch_bounds__(256) probe(int n, const double* S, const double* rhs, double* x, long long* cyc) {
    extern __shared__ double smem[];
    double* xs = smem + n * (n + 1);
    solve_t<8, 2>(n, S, rhs, smem, xs, cyc);
    if (threadIdx.x < n && blockIdx.x == 0) x[threadIdx.x] = xs[threadIdx.x];
}
int main() {
    const int n = 60;
    std::vector<double> A(n * n), b(n), B(n * n);
    srand(1);
    for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += B[i * n + k] * B[j * n + k];
            A[i * n + j] = s + (i == j ? 1.0 : 0.0);
        }
    for (int i = 0; i < n; ++i) b[i] = i * 0.1 - 1;
    double *dS, *db, *dx; long long* dc;
    cudaMalloc(&dS, n * n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dx, n * 8); cudaMallocManaged(&dc, 24);
    cudaMemcpy(dS, A.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
    const size_t sm = (n * (n + 1) + n) * 8;
    for (int grid : {1, 148}) {
        for (int rep = 0; rep < 3; ++rep) probe<<<grid, 256, sm>>>(n, dS, db, dx, dc);
        cudaDeviceSynchronize();
        printf("grid %3d: load %lld  factor %lld (%.0f per column)  backward %lld cycles\n", grid, dc[0], dc[1], dc[1] / (double)(n - 1), dc[2]);
    }
    std::vector<double> x(n);
    cudaMemcpy(x.data(), dx, n * 8, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < n; ++i) {
        double s = -b[i];
        for (int j = 0; j < n; ++j) s += A[i * n + j] * x[j];
        worst = fmax(worst, fabs(s));
    }
    printf("max residual %.3e\n", worst);
    return 0;
}
