// Stand-alone probe of the 6x6-blocked Cholesky of ba_lm_small_kernel (copy of ba_small_solve, csrc/ba.cu) with
// clock64 at its phase boundaries: where do the 24 us of a 60x60 solve go?
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define BA_SMALL_N 96
__device__ __forceinline__ bool chol6_diag(double* M, int ld, int j0, double* rinv) {
    bool ok = true;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double d = M[(j0 + r) * ld + j0 + r];
#pragma unroll
        for (int p = 0; p < 6; ++p)
            if (p < r) d -= M[(j0 + p) * ld + j0 + r] * M[(j0 + p) * ld + j0 + r];
        if (!(d > 0.0) || !isfinite(d)) ok = false;
        double ri = rsqrt(d);
        ri = ri * (1.5 - 0.5 * d * ri * ri);
        rinv[r] = ri;
        M[(j0 + r) * ld + j0 + r] = d * ri;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            if (c > r) {
                double v = M[(j0 + r) * ld + j0 + c];
#pragma unroll
                for (int p = 0; p < 6; ++p)
                    if (p < r) v -= M[(j0 + p) * ld + j0 + r] * M[(j0 + p) * ld + j0 + c];
                M[(j0 + r) * ld + j0 + c] = v * ri;
            }
        }
    }
    return ok;
}
// Root-free factorisation of the 6x6 block with the square roots off the dependency chain: u'[r][c] (unnormalised rows,
// u'[r][r] = d_r), g[p][r] = u'[p][r] / d_p.  The chain per row is one reciprocal + one multiply + one DFMA; the six
// rsqrt(d_r) that turn the rows into the Cholesky factor U = diag(rsqrt(d)) u' are independent of it.
// branch-free reciprocal and reciprocal square root for well-scaled positive doubles (Hessian diagonals): hardware
// seed (about 20 bits) + Newton steps to full precision; unlike 1.0 / d and rsqrt(d) they carry no special-case branch,
// so independent ones interleave
__device__ __forceinline__ double rcp_pos(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    y = y * fma(-h * y, y, 1.5);
    return y * fma(-h * y, y, 1.5);
}
__device__ __forceinline__ bool ldl6_diag(double* D, double* rs, double* G) {
    bool ok = true;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const double d = D[r * 6 + r];
        if (!(d > 0.0) || !isfinite(d)) ok = false;
        const double inv = rcp_pos(d);
        rs[r] = rsqrt_pos(d);
#pragma unroll
        for (int r2 = 0; r2 < 6; ++r2) {
            if (r2 > r) {
                const double g = D[r * 6 + r2] * inv;
                G[r * 6 + r2] = g;
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    if (c >= r2) D[r2 * 6 + c] -= g * D[r * 6 + c];
            }
        }
    }
    return ok;
}
__global__ void __launch_bounds__(256) probe(int n, const double* S, const double* rhs, double* xout, long long* cyc) {
    extern __shared__ double M[];
    double* xs = M + n * (n + 1);
    const int ld = n + 1, tid = threadIdx.x, nt = blockDim.x;
    __shared__ int s_fail;
    __shared__ double s_rd[BA_SMALL_N];
    long long t_diag = 0, t_trail = 0, t_bar = 0, ta, tb;
    if (tid == 0) s_fail = 0;
    long long t0 = clock64();
#ifdef V_OLD
    for (int i = tid; i < n * n; i += nt) {
        const int r = i / n, c = i - r * n;
        M[r * ld + c] = __ldcg(S + i);
    }
#else
    {   // row per warp, lanes along the columns right of the diagonal: every load of a thread in flight at once
        const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
        double lv[12][3];
#pragma unroll
        for (int a = 0; a < 12; ++a) {
            const int r = warp + nw * a;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int c = lane + 32 * b;
                lv[a][b] = (r < n && c >= r && c < n) ? __ldcg(S + r * n + c) : 0.0;
            }
        }
#pragma unroll
        for (int a = 0; a < 12; ++a) {
            const int r = warp + nw * a;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int c = lane + 32 * b;
                if (r < n && c >= r && c < n) M[r * ld + c] = lv[a][b];
            }
        }
    }
#endif
    for (int i = tid; i < n; i += nt) M[i * ld + n] = __ldcg(rhs + i);
    __syncthreads();
    long long t1 = clock64();
    for (int j0 = 0; j0 < n; j0 += 6) {
        ta = clock64();
        {
            double D[36];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) D[r * 6 + c] = (c >= r) ? M[(j0 + r) * ld + j0 + c] : 0.0;
            double rinv[6];
            const bool ok = chol6_diag(D, 6, 0, rinv);
            for (int c = j0 + 6 + tid; c <= n; c += nt) {
                double v[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) v[r] = M[(j0 + r) * ld + c];
#pragma unroll
                for (int r = 0; r < 6; ++r) {
#pragma unroll
                    for (int p = 0; p < 6; ++p)
                        if (p < r) v[r] -= D[p * 6 + r] * v[p];
                    v[r] *= rinv[r];
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) M[(j0 + r) * ld + c] = v[r];
            }
            __syncthreads();
            if (tid == 0) {
                if (!ok) s_fail = 1;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
#pragma unroll
                    for (int c = 0; c < 6; ++c)
                        if (c >= r) M[(j0 + r) * ld + j0 + c] = D[r * 6 + c];
                    s_rd[j0 + r] = rinv[r];
                }
            }
        }
        tb = clock64();
        t_diag += tb - ta;
#ifdef V_OLD
        const int m = n - j0 - 6;
        for (int e = tid; e < m * (m + 1); e += nt) {
            const int r = j0 + 6 + e / (m + 1), c = j0 + 6 + e % (m + 1);
            if (c < r) continue;
            double v = M[r * ld + c];
#pragma unroll
            for (int p = 0; p < 6; ++p) v -= M[(j0 + p) * ld + r] * M[(j0 + p) * ld + c];
            M[r * ld + c] = v;
        }
#else
        {   // lanes along the columns (their 6 panel entries hoisted into registers), warps along the rows, two rows per
            // step for independent DFMA chains; no integer division, no per-element index arithmetic
            const int c0 = j0 + 6, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
            const double* W = M + j0 * ld;  // panel rows j0 .. j0+5
            double wc[3][6];
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int c = min(c0 + lane + 32 * b, n);
#pragma unroll
                for (int p = 0; p < 6; ++p) wc[b][p] = W[p * ld + c];
            }
            // four rows per step, branch-free inside: every slot computes, stores are predicated
            const int nslot = (n - c0) / 32 + 1;  // column slots that hold any column <= n
            for (int r = c0 + 4 * warp; r < n; r += 4 * nw) {
                double wr[4][6], v[4][3];
                int rr[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    rr[q] = min(r + q, n - 1);
#pragma unroll
                    for (int p = 0; p < 6; ++p) wr[q][p] = W[p * ld + rr[q]];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < 3; ++b) v[q][b] = (b < nslot) ? M[rr[q] * ld + min(c0 + lane + 32 * b, n)] : 0.0;
#pragma unroll
                for (int p = 0; p < 6; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int b = 0; b < 3; ++b) v[q][b] -= wr[q][p] * wc[b][p];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const int c = c0 + lane + 32 * b;
                        if (r + q < n && c >= r + q && c <= n) M[(r + q) * ld + c] = v[q][b];
                    }
            }
        }
#endif
        long long tc = clock64();
        __syncthreads();
        t_trail += tc - tb; t_bar += clock64() - tc;
        if (tid == 0) { cyc[8 + j0 / 6] = tc - tb; cyc[24 + j0 / 6] = tb - ta; }
        if (s_fail) break;
    }
    long long t2 = clock64();
    for (int j0 = n - 6; j0 >= 0; j0 -= 6) {
        double D[36], x[6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) D[r * 6 + c] = (c >= r) ? M[(j0 + r) * ld + j0 + c] : 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) x[r] = M[(j0 + r) * ld + n];
#pragma unroll
        for (int r = 5; r >= 0; --r) {
#pragma unroll
            for (int p = 0; p < 6; ++p)
                if (p > r) x[r] -= D[r * 6 + p] * x[p];
            x[r] *= s_rd[j0 + r];
        }
        __syncthreads();
        for (int r = tid; r < j0; r += nt) {
            double v = M[r * ld + n];
#pragma unroll
            for (int p = 0; p < 6; ++p) v -= M[r * ld + j0 + p] * x[p];
            M[r * ld + n] = v;
        }
        if (tid < 6) M[(j0 + tid) * ld + n] = x[tid];
        __syncthreads();
    }
    for (int i = tid; i < n; i += nt) xs[i] = M[i * ld + n];
    __syncthreads();
    long long t3 = clock64();
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t_diag; cyc[2] = t_trail; cyc[3] = t3 - t2; cyc[4] = t3 - t0; cyc[5] = t_bar; }
    if (tid < n) xout[tid] = xs[tid];
}

template <int SB_A, int SB_B>
__global__ void __launch_bounds__(256) probe3(int n, const double* S, const double* rhs, double* xout, long long* cyc) {
    extern __shared__ double M[];
    double* xs = M + n * (n + 1);
    const int ld = n + 1, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_fail;
    __shared__ double s_rd[BA_SMALL_N];
    long long t_diag = 0, t_trail = 0, ta, tb;
    if (tid == 0) s_fail = 0;
    if (tid == 0) { cyc[40] = cyc[41] = cyc[42] = 0; }
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
    long long t1 = clock64();
    for (int j0 = 0; j0 < n; j0 += 6) {
        ta = clock64();
        // (1) publish the six pivot rows (each lives in one warp)
#pragma unroll
        for (int a = 0; a < SB_A; ++a) {
            const int r = warp + 8 * a;
            if (r >= j0 && r < j0 + 6) {
#pragma unroll
                for (int b = 0; b < SB_B; ++b) {
                    const int c = lane + 32 * b;
                    if (c >= r && c <= n) M[r * ld + c] = v[a][b];
                }
            }
        }
        __syncthreads();
        long long q0 = clock64();
        // (2) diagonal block in registers by every thread, panel columns by one thread each
        {
            double D[36];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) D[r * 6 + c] = (c >= r) ? M[(j0 + r) * ld + j0 + c] : 0.0;
            double rinv[6], G[36];
            const bool ok = ldl6_diag(D, rinv, G);
            long long q1 = clock64();
            for (int c = j0 + 6 + tid; c <= n; c += nt) {
                double w[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) w[r] = M[(j0 + r) * ld + c];
#pragma unroll
                for (int r = 0; r < 6; ++r) {
#pragma unroll
                    for (int p = 0; p < 6; ++p)
                        if (p < r) w[r] -= G[p * 6 + r] * w[p];
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) M[(j0 + r) * ld + c] = w[r] * rinv[r];
            }
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    if (c >= r) D[r * 6 + c] *= rinv[r];
            long long q2 = clock64();
            if (tid == 0) { cyc[40] += q0 - ta; cyc[41] += q1 - q0; cyc[42] += q2 - q1; }
            if (tid == 32) {  // a thread that has no panel column in the late steps... any thread: D is private
                if (!ok) s_fail = 1;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
#pragma unroll
                    for (int c = 0; c < 6; ++c)
                        if (c >= r) M[(j0 + r) * ld + j0 + c] = D[r * 6 + c];
                    s_rd[j0 + r] = rinv[r];
                }
            }
        }
        __syncthreads();
        tb = clock64();
        t_diag += tb - ta;
        // (3) rank-6 update of the register-resident trailing rows
        {
            const double* W = M + j0 * ld;
            double wc[SB_B][6];
#pragma unroll
            for (int b = 0; b < SB_B; ++b) {
                const int c = min(lane + 32 * b, n);
#pragma unroll
                for (int p = 0; p < 6; ++p) wc[b][p] = W[p * ld + c];
            }
            const int a1 = (j0 + 6 - warp + 7) >> 3;  // first row slot of this warp below the panel
#pragma unroll
            for (int a = 0; a < SB_A; ++a) {
                const int r = warp + 8 * a;
                if (a >= a1 && r < n) {
                    double wr[6];
#pragma unroll
                    for (int p = 0; p < 6; ++p) wr[p] = W[p * ld + r];
#pragma unroll
                    for (int p = 0; p < 6; ++p)
#pragma unroll
                        for (int b = 0; b < SB_B; ++b) v[a][b] -= wr[p] * wc[b][p];
                }
            }
        }
        t_trail += clock64() - tb;
        if (tid == 0) { cyc[8 + j0 / 6] = clock64() - tb; cyc[24 + j0 / 6] = tb - ta; }
        if (s_fail) break;
    }
    __syncthreads();
    long long t2 = clock64();
    for (int j0 = n - 6; j0 >= 0; j0 -= 6) {
        double D[36], x[6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) D[r * 6 + c] = (c >= r) ? M[(j0 + r) * ld + j0 + c] : 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) x[r] = M[(j0 + r) * ld + n];
#pragma unroll
        for (int r = 5; r >= 0; --r) {
#pragma unroll
            for (int p = 0; p < 6; ++p)
                if (p > r) x[r] -= D[r * 6 + p] * x[p];
            x[r] *= s_rd[j0 + r];
        }
        __syncthreads();
        for (int r = tid; r < j0; r += nt) {
            double vv = M[r * ld + n];
#pragma unroll
            for (int p = 0; p < 6; ++p) vv -= M[r * ld + j0 + p] * x[p];
            M[r * ld + n] = vv;
        }
        if (tid == 0) {
#pragma unroll
            for (int r = 0; r < 6; ++r) M[(j0 + r) * ld + n] = x[r];
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += nt) xs[i] = M[i * ld + n];
    __syncthreads();
    long long t3 = clock64();
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t_diag; cyc[2] = t_trail; cyc[3] = t3 - t2; cyc[4] = t3 - t0; cyc[5] = 0; }
    if (tid < n) xout[tid] = xs[tid];
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
    cudaMalloc(&dS, n * n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dx, n * 8); cudaMallocManaged(&dc, 8 * 64);
    cudaMemcpy(dS, A.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
    const size_t sm = (n * (n + 1) + n) * 8;
#ifdef V_NEW
    for (int rep = 0; rep < 3; ++rep) probe3<8, 2><<<1, 256, sm>>>(n, dS, db, dx, dc);
#else
    for (int rep = 0; rep < 3; ++rep) probe<<<1, 256, sm>>>(n, dS, db, dx, dc);
#endif
    cudaDeviceSynchronize();
    printf("load %lld  diag+panel %lld  trailing %lld (+barrier wait %lld)  backward %lld  total %lld cycles (%.1f us at 1965 MHz)\n", dc[0], dc[1], dc[2], dc[5], dc[3], dc[4], dc[4] / 1965.0);
    printf("publish+barrier %lld  chol6 %lld  panel %lld\n", dc[40], dc[41], dc[42]);
    printf("trailing per block:"); for (int i = 0; i < 10; ++i) printf(" %lld", dc[8 + i]); printf("\ndiag+panel per block:"); for (int i = 0; i < 10; ++i) printf(" %lld", dc[24 + i]); printf("\n");
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
