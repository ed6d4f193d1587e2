// Device-side minimal solver of K12: the 5-point EPnP kernel that cv::solvePnPRansac runs on every RANSAC sample
// (called from VO::motion_estimation, /root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:277).
//
// OpenCV is an un-vendored dependency of the reference; what is reproduced here is cv2 4.13.0's arithmetic, operation
// for operation, because the result of EPnP on five points depends on it: M^T M has an exactly two-dimensional null
// space, OpenCV's one-sided Jacobi SVD returns a basis of it that is fixed by rounding alone, and the three starting
// points of the beta refinement depend on that basis.  Everything below therefore runs in one thread per sample, in
// IEEE fp64 with no contraction (the file is compiled with -fmad=false), sums in OpenCV's order:
//   epnp::compute_pose and its helpers           (calib3d/src/epnp.cpp)
//   JacobiSVDImpl_<double>, SVBkSb, cv's hypot   (core/src/lapack.cpp; sizes below the LAPACK hand-over)
//   MulTransposedR                               (core/src/matmul.simd.hpp)
//   cv::Rodrigues                                (calib3d/src/calibration.cpp)
// tests/test_pnp_gpu.py compares the per-sample models and the inlier lists with live cv2 index for index.
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>

// The same source is compiled for the device (one warp per sample) and, by tests/native/epnp_host.cpp, for the CPU,
// where the lanes of a wavefront become a loop.
#ifdef EPNP_HOST_BUILD
#define EPNP_SYNCWARP()
#define EPNP_FOR_LANES(p, cnt, lane) for (int p = 0; p < (cnt); ++p)
#define EPNP_ANY(x) (x)
#else
#define EPNP_SYNCWARP() __syncwarp()
#define EPNP_FOR_LANES(p, cnt, lane) \
    const int p = (lane);            \
    if (p < (cnt))
#define EPNP_ANY(x) __any_sync(0xFFFFFFFFu, (x))
#endif

#define EPNP_AS 13  // row stride of the 12x12 matrix in shared memory: rows of different lanes fall into different banks

struct EpnpWork {  // per-sample scratch (shared memory)
    double At[12 * EPNP_AS];  // M^T M (symmetric) -> rows of U^T
    double W12[12], d12[12];
    double M[10 * 12];
    double pws[15], us[10], alphas[20];
    double cws[4][3];
    double L[60], rho[6];
    double uc, vc, fu, fv;
};

__device__ __forceinline__ uint32_t cvrng_next(uint64_t& state) {
    state = (uint64_t)(uint32_t)state * 4164903690ull + (uint32_t)(state >> 32);
    return (uint32_t)state;
}

__device__ __forceinline__ double cv_hypot(double a, double b) {
    a = fabs(a);
    b = fabs(b);
    if (a > b) {
        b /= a;
        return a * sqrt(1 + b * b);
    }
    if (b > 0) {
        a /= b;
        return b * sqrt(1 + a * a);
    }
    return 0;
}

// One rotation step of JacobiSVDImpl_ on rows i, j (length M, compile-time so that the loads and products of a row are
// independent instructions; the three sums stay sequential in k, as in OpenCV).  Returns whether the pair was rotated.
template <int M, int NV>
__device__ __forceinline__ bool jacobi_pair(double* __restrict__ Ai, double* __restrict__ Aj, double* __restrict__ Vi,
                                            double* __restrict__ Vj, double& Wi, double& Wj) {
    const double eps = DBL_EPSILON * 10;
    double ai[M], aj[M];
#pragma unroll
    for (int k = 0; k < M; k++) {
        ai[k] = Ai[k];
        aj[k] = Aj[k];
    }
    double a = Wi, p = 0, b = Wj;
#pragma unroll
    for (int k = 0; k < M; k++) p += ai[k] * aj[k];
    if (fabs(p) <= eps * sqrt(a * b)) return false;
    p *= 2;
    const double beta = a - b, gamma = cv_hypot(p, beta);
    double c, s;
    if (beta < 0) {
        const double delta = (gamma - beta) * 0.5;
        s = sqrt(delta / gamma);
        c = p / (gamma * s * 2);
    } else {
        c = sqrt((gamma + beta) / (gamma * 2));
        s = p / (gamma * c * 2);
    }
    a = b = 0;
#pragma unroll
    for (int k = 0; k < M; k++) {
        const double t0 = c * ai[k] + s * aj[k];
        const double t1 = -s * ai[k] + c * aj[k];
        Ai[k] = t0;
        Aj[k] = t1;
        a += t0 * t0;
        b += t1 * t1;
    }
    Wi = a;
    Wj = b;
    if (NV > 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const double t0 = c * Vi[k] + s * Vj[k];
            const double t1 = -s * Vi[k] + c * Vj[k];
            Vi[k] = t0;
            Vj[k] = t1;
        }
    }
    return true;
}

// Tail of JacobiSVDImpl_: singular values = row norms, selection sort (descending, first maximum wins), rows scaled to
// unit length -- a row whose norm is <= DBL_MIN is replaced by OpenCV's random +-1/m vector orthogonalised against the
// rows before it.  AS = row stride of At.
template <int M, int N, int AS, bool WITH_V>
__device__ void jacobi_finish(double* At, double* W, double* Wout, double* Vt) {
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    for (int i = 0; i < N; i++) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < M; k++) {
            const double t = At[i * AS + k];
            sd += t * t;
        }
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < N - 1; i++) {
        int j = i;
        for (int k = i + 1; k < N; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i];
            W[i] = W[j];
            W[j] = t;
#pragma unroll
            for (int k = 0; k < M; k++) {
                t = At[i * AS + k];
                At[i * AS + k] = At[j * AS + k];
                At[j * AS + k] = t;
            }
            if (WITH_V)
#pragma unroll
                for (int k = 0; k < N; k++) {
                    t = Vt[i * N + k];
                    Vt[i * N + k] = Vt[j * N + k];
                    Vt[j * N + k] = t;
                }
        }
    }
    for (int i = 0; i < N; i++) Wout[i] = W[i];
    uint64_t rng = 0x12345678;
    for (int i = 0; i < N; i++) {
        double sd = W[i];
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            const double val0 = 1. / M;
            for (int k = 0; k < M; k++) At[i * AS + k] = (cvrng_next(rng) & 256) != 0 ? val0 : -val0;
            for (int iter = 0; iter < 2; iter++)
                for (int j = 0; j < i; j++) {
                    sd = 0;
                    for (int k = 0; k < M; k++) sd += At[i * AS + k] * At[j * AS + k];
                    double asum = 0;
                    for (int k = 0; k < M; k++) {
                        const double t = At[i * AS + k] - sd * At[j * AS + k];
                        At[i * AS + k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < M; k++) At[i * AS + k] *= asum;
                }
            sd = 0;
            for (int k = 0; k < M; k++) {
                const double t = At[i * AS + k];
                sd += t * t;
            }
            sd = sqrt(sd);
        }
        const double s = sd > minval ? 1 / sd : 0.;
#pragma unroll
        for (int k = 0; k < M; k++) At[i * AS + k] *= s;
    }
}

// One-sided Jacobi SVD (JacobiSVDImpl_<double>) of the N rows (length M) of At; rows become U^T, Wout the singular
// values (descending), Vt (N x N) the right singular vectors.  One thread, cyclic pair order.
template <int M, int N>
__device__ __noinline__ void jacobi_svd_small(double* At, double* Wout, double* Vt) {
    double W[N];
    const int max_iter = M > 30 ? M : 30;
    for (int i = 0; i < N; i++) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < M; k++) {
            const double t = At[i * M + k];
            sd += t * t;
        }
        W[i] = sd;
#pragma unroll
        for (int k = 0; k < N; k++) Vt[i * N + k] = 0;
        Vt[i * N + i] = 1;
    }
#pragma unroll 1
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
#pragma unroll 1
        for (int i = 0; i < N - 1; i++)
#pragma unroll 1
            for (int j = i + 1; j < N; j++)
                changed |= jacobi_pair<M, N>(At + i * M, At + j * M, Vt + i * N, Vt + j * N, W[i], W[j]);
        if (!changed) break;
    }
    jacobi_finish<M, N, M, true>(At, W, Wout, Vt);
}

// The 12x12 decomposition of M^T M (only U^T is used), rows of At padded to EPNP_AS doubles.  OpenCV walks the 66 row
// pairs (i, j) of a sweep in cyclic order; pair (i, j) only depends on the last earlier pairs that touched row i or
// row j, which puts it on wavefront i + j - 1: the up to six pairs of one wavefront touch disjoint rows and are
// rotated by six lanes at once, 21 wavefronts per sweep instead of 66 sequential pairs -- with every row seeing exactly
// the sequence of rotations (and therefore the bits) of the sequential loop.  Called by a whole warp (lanes >= 6 idle);
// on the host build the lanes of a wavefront are a loop.
__device__ void jacobi_svd12_wave(double* At, double* W /*[12] shared*/, double* Wout, int lane) {
    const int N = 12;
    if (lane == 0)
        for (int i = 0; i < N; i++) {
            double sd = 0;
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const double t = At[i * EPNP_AS + k];
                sd += t * t;
            }
            W[i] = sd;
        }
    EPNP_SYNCWARP();
#pragma unroll 1
    for (int iter = 0; iter < 30; iter++) {
        bool changed = false;
#pragma unroll 1
        for (int t = 0; t <= 2 * N - 4; t++) {
            const int i_lo = t + 1 - (N - 1) > 0 ? t + 1 - (N - 1) : 0, cnt = t / 2 - i_lo + 1;
            EPNP_FOR_LANES(p, cnt, lane) {
                const int i = i_lo + p, j = t + 1 - i;
                changed |= jacobi_pair<12, 0>(At + i * EPNP_AS, At + j * EPNP_AS, nullptr, nullptr, W[i], W[j]);
            }
            EPNP_SYNCWARP();
        }
        if (!EPNP_ANY(changed)) break;
    }
    if (lane == 0) jacobi_finish<12, 12, EPNP_AS, false>(At, W, Wout, nullptr);
    EPNP_SYNCWARP();
}

// cv::SVD::compute of A (M x N row-major, M >= N): Ut = N rows of length M, Vt = N x N
template <int M, int N>
__device__ void svd_small_dev(const double* A, double* w, double* Ut, double* Vt) {
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int k = 0; k < M; k++) Ut[i * M + k] = A[k * N + i];
    jacobi_svd_small<M, N>(Ut, w, Vt);
}

// cv::solve(A (6 x n), b, x, DECOMP_SVD): x = V diag(1/w) U^T b, singular values <= 2 eps sum(w) dropped
template <int N>
__device__ void solve_svd6_dev(const double* A, const double* b, double* x) {
    const int n = N;
    double Ut[6 * N], Vt[N * N], w[N];
    svd_small_dev<6, N>(A, w, Ut, Vt);
    double threshold = 0;
    for (int i = 0; i < n; i++) x[i] = 0;
    for (int i = 0; i < n; i++) threshold += w[i];
    threshold *= DBL_EPSILON * 2;
    for (int i = 0; i < n; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double s = 0;
        for (int j = 0; j < 6; j++) s += Ut[i * 6 + j] * b[j];
        s *= wi;
        for (int j = 0; j < n; j++) x[j] = x[j] + s * Vt[i * n + j];
    }
}

// cv::invert(A 3x3, DECOMP_SVD)
__device__ void invert3_svd_dev(const double* A, double* Ainv) {
    double Ut[9], Vt[9], w[3];
    svd_small_dev<3, 3>(A, w, Ut, Vt);
    const double threshold = (w[0] + w[1] + w[2]) * DBL_EPSILON * 2;
    for (int i = 0; i < 9; i++) Ainv[i] = 0;
    for (int i = 0; i < 3; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) Ainv[j * 3 + k] += Vt[i * 3 + j] * (Ut[i * 3 + k] * wi);
    }
}

// dst = src^T src (rows x cols), every entry summed over the rows in order, upper triangle mirrored
__device__ void mul_transposed_dev(const double* src, int rows, int cols, double* dst, int dst_stride) {
    for (int i = 0; i < cols; i++)
        for (int j = i; j < cols; j++) {
            double s0 = 0;
            for (int k = 0; k < rows; k++) s0 += src[k * cols + i] * src[k * cols + j];
            dst[i * dst_stride + j] = s0;
            dst[j * dst_stride + i] = s0;
        }
}

__device__ __forceinline__ double dot3_dev(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ double dist2_dev(const double* p1, const double* p2) {
    return (p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]) + (p1[2] - p2[2]) * (p1[2] - p2[2]);
}

// R -> rotation vector (the matrix is first replaced by U V^T of its SVD, as cv::Rodrigues does)
__device__ void rodrigues_to_vec_dev(const double* Rin, double* r) {
    double w[3], Ut[9], Vt[9], R[9];
    svd_small_dev<3, 3>(Rin, w, Ut, Vt);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = Ut[i] * Vt[j] + Ut[3 + i] * Vt[3 + j] + Ut[6 + i] * Vt[6 + j];
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) {
            r[0] = r[1] = r[2] = 0;
        } else {
            double t = (R[0] + 1) * 0.5;
            rx = sqrt(t > 0. ? t : 0.);
            t = (R[4] + 1) * 0.5;
            ry = sqrt(t > 0. ? t : 0.) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5;
            rz = sqrt(t > 0. ? t : 0.) * (R[2] < 0 ? -1. : 1.);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            r[0] = rx * theta;
            r[1] = ry * theta;
            r[2] = rz * theta;
        }
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        r[0] = rx * vth;
        r[1] = ry * vth;
        r[2] = rz * vth;
    }
}

// rotation vector -> R
__device__ void rodrigues_to_mat_dev(const double* r, double* R) {
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1 : 0;
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = theta ? 1. / theta : 0.;
    const double rx = r[0] * itheta, ry = r[1] * itheta, rz = r[2] * itheta;
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; k++) R[k] = c * ((k % 4 == 0) ? 1 : 0) + c1 * rrt[k] + s * r_x[k];
}

// ------------------------------------------------------------------------------------------------------------------
__device__ void epnp_choose_control_points(EpnpWork& e, int n) {
    e.cws[0][0] = e.cws[0][1] = e.cws[0][2] = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) e.cws[0][j] += e.pws[3 * i + j];
    for (int j = 0; j < 3; j++) e.cws[0][j] /= n;
    double pw0[15], g[9], dc[3], At[9];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) pw0[3 * i + j] = e.pws[3 * i + j] - e.cws[0][j];
    mul_transposed_dev(pw0, n, 3, g, 3);
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) At[i * 3 + k] = g[k * 3 + i];
    double Vt[9];
    jacobi_svd_small<3, 3>(At, dc, Vt);
    for (int i = 1; i < 4; i++) {
        const double k = sqrt(dc[i - 1] / n);
        for (int j = 0; j < 3; j++) e.cws[i][j] = e.cws[0][j] + k * At[3 * (i - 1) + j];
    }
}

__device__ void epnp_barycentric(EpnpWork& e, int n) {
    double cc[9], ci[9];
    for (int i = 0; i < 3; i++)
        for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = e.cws[j][i] - e.cws[0][i];
    invert3_svd_dev(cc, ci);
    for (int i = 0; i < n; i++) {
        const double* pi = e.pws + 3 * i;
        double* a = e.alphas + 4 * i;
        for (int j = 0; j < 3; j++)
            a[1 + j] = ci[3 * j] * (pi[0] - e.cws[0][0]) + ci[3 * j + 1] * (pi[1] - e.cws[0][1]) +
                       ci[3 * j + 2] * (pi[2] - e.cws[0][2]);
        a[0] = 1.0f - a[1] - a[2] - a[3];
    }
}

__device__ double epnp_R_and_t(const EpnpWork& e, int n, const double* betas, double* R /*[9]*/, double* t) {
    // compute_ccs / compute_pcs / solve_for_sign
    double ccs[4][3], pcs[15];
    for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
    for (int i = 0; i < 4; i++) {
        const double* v = e.At + EPNP_AS * (11 - i);
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 3; k++) ccs[j][k] += betas[i] * v[3 * j + k];
    }
    for (int i = 0; i < n; i++) {
        const double* a = e.alphas + 4 * i;
        for (int j = 0; j < 3; j++)
            pcs[3 * i + j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
    }
    if (pcs[2] < 0.0) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 3; j++) ccs[i][j] = -ccs[i][j];
        for (int i = 0; i < 3 * n; i++) pcs[i] = -pcs[i];
    }
    // estimate_R_and_t
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) {
            pc0[j] += pcs[3 * i + j];
            pw0[j] += e.pws[3 * i + j];
        }
    for (int j = 0; j < 3; j++) {
        pc0[j] /= n;
        pw0[j] /= n;
    }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        const double* pc = pcs + 3 * i;
        const double* pw = e.pws + 3 * i;
        for (int j = 0; j < 3; j++) {
            abt[3 * j] += (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
            abt[3 * j + 1] += (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
            abt[3 * j + 2] += (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
        }
    }
    double w[3], Ut[9], Vt[9];
    svd_small_dev<3, 3>(abt, w, Ut, Vt);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = Ut[i] * Vt[j] + Ut[3 + i] * Vt[3 + j] + Ut[6 + i] * Vt[6 + j];
    const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] - R[1] * R[3] * R[8] -
                       R[0] * R[5] * R[7];
    if (det < 0) {
        R[6] = -R[6];
        R[7] = -R[7];
        R[8] = -R[8];
    }
    t[0] = pc0[0] - dot3_dev(R, pw0);
    t[1] = pc0[1] - dot3_dev(R + 3, pw0);
    t[2] = pc0[2] - dot3_dev(R + 6, pw0);
    // reprojection_error
    double sum2 = 0.0;
    for (int i = 0; i < n; i++) {
        const double* pw = e.pws + 3 * i;
        const double Xc = dot3_dev(R, pw) + t[0];
        const double Yc = dot3_dev(R + 3, pw) + t[1];
        const double inv_Zc = 1.0 / (dot3_dev(R + 6, pw) + t[2]);
        const double ue = e.uc + e.fu * Xc * inv_Zc;
        const double ve = e.vc + e.fv * Yc * inv_Zc;
        const double u = e.us[2 * i], v = e.us[2 * i + 1];
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / n;
}

__device__ void epnp_L_6x10(const double* ut, double* l_6x10) {
    const double* v[4] = {ut + EPNP_AS * 11, ut + EPNP_AS * 10, ut + EPNP_AS * 9, ut + EPNP_AS * 8};
    double dv[4][6][3];
    for (int i = 0; i < 4; i++) {
        int a = 0, b = 1;
        for (int j = 0; j < 6; j++) {
            dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
            dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
            dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
            b++;
            if (b > 3) {
                a++;
                b = a + 1;
            }
        }
    }
    for (int i = 0; i < 6; i++) {
        double* row = l_6x10 + 10 * i;
        row[0] = dot3_dev(dv[0][i], dv[0][i]);
        row[1] = 2.0f * dot3_dev(dv[0][i], dv[1][i]);
        row[2] = dot3_dev(dv[1][i], dv[1][i]);
        row[3] = 2.0f * dot3_dev(dv[0][i], dv[2][i]);
        row[4] = 2.0f * dot3_dev(dv[1][i], dv[2][i]);
        row[5] = dot3_dev(dv[2][i], dv[2][i]);
        row[6] = 2.0f * dot3_dev(dv[0][i], dv[3][i]);
        row[7] = 2.0f * dot3_dev(dv[1][i], dv[3][i]);
        row[8] = 2.0f * dot3_dev(dv[2][i], dv[3][i]);
        row[9] = dot3_dev(dv[3][i], dv[3][i]);
    }
}

// starting points of the beta refinement: betas10 = [B11 B12 B22 B13 B23 B33 B14 B24 B34 B44]
__device__ void epnp_betas_approx(int which, const double* L, const double* rho, double* betas) {
    double l[30], b[5];
    if (which == 1) {  // [B11 B12 B13 B14]
        for (int i = 0; i < 6; i++) {
            l[i * 4] = L[i * 10];
            l[i * 4 + 1] = L[i * 10 + 1];
            l[i * 4 + 2] = L[i * 10 + 3];
            l[i * 4 + 3] = L[i * 10 + 6];
        }
        solve_svd6_dev<4>(l, rho, b);
        if (b[0] < 0) {
            betas[0] = sqrt(-b[0]);
            betas[1] = -b[1] / betas[0];
            betas[2] = -b[2] / betas[0];
            betas[3] = -b[3] / betas[0];
        } else {
            betas[0] = sqrt(b[0]);
            betas[1] = b[1] / betas[0];
            betas[2] = b[2] / betas[0];
            betas[3] = b[3] / betas[0];
        }
        return;
    }
    const int nc = which == 2 ? 3 : 5;  // [B11 B12 B22] or [B11 B12 B22 B13 B23]
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < nc; j++) l[i * nc + j] = L[i * 10 + j];
    if (which == 2) solve_svd6_dev<3>(l, rho, b);
    else solve_svd6_dev<5>(l, rho, b);
    if (b[0] < 0) {
        betas[0] = sqrt(-b[0]);
        betas[1] = (b[2] < 0) ? sqrt(-b[2]) : 0.0;
    } else {
        betas[0] = sqrt(b[0]);
        betas[1] = (b[2] > 0) ? sqrt(b[2]) : 0.0;
    }
    if (b[1] < 0) betas[0] = -betas[0];
    betas[2] = which == 3 ? b[3] / betas[0] : 0.0;
    betas[3] = 0.0;
}

// epnp::qr_solve on the 6x4 Gauss-Newton system (Householder; A is overwritten).  Returns false where OpenCV bails
// out on a singular column (x is then left as it was).
__device__ bool epnp_qr_solve(double* A, double* b, double* X) {
    const int nr = 6, nc = 4;
    double A1[4], A2[4];
    for (int k = 0; k < nc; k++) {
        double eta = fabs(A[k * nc + k]);
        for (int i = k + 1; i < nr; i++) {
            // OpenCV scans |A[k][k]|, |A[k][k]|, |A[k+1][k]| ... (the pointer is advanced after the comparison)
            const double elt = fabs(A[(i - 1) * nc + k]);
            if (eta < elt) eta = elt;
        }
        if (eta == 0) return false;
        double sum2 = 0.0;
        const double inv_eta = 1. / eta;
        for (int i = k; i < nr; i++) {
            A[i * nc + k] *= inv_eta;
            sum2 += A[i * nc + k] * A[i * nc + k];
        }
        double sigma = sqrt(sum2);
        if (A[k * nc + k] < 0) sigma = -sigma;
        A[k * nc + k] += sigma;
        A1[k] = sigma * A[k * nc + k];
        A2[k] = -eta * sigma;
        for (int j = k + 1; j < nc; j++) {
            double sum = 0;
            for (int i = k; i < nr; i++) sum += A[i * nc + k] * A[i * nc + j];
            const double tau = sum / A1[k];
            for (int i = k; i < nr; i++) A[i * nc + j] -= tau * A[i * nc + k];
        }
    }
    for (int j = 0; j < nc; j++) {
        double tau = 0;
        for (int i = j; i < nr; i++) tau += A[i * nc + j] * b[i];
        tau /= A1[j];
        for (int i = j; i < nr; i++) b[i] -= tau * A[i * nc + j];
    }
    X[nc - 1] = b[nc - 1] / A2[nc - 1];
    for (int i = nc - 2; i >= 0; i--) {
        double sum = 0;
        for (int j = i + 1; j < nc; j++) sum += A[i * nc + j] * X[j];
        X[i] = (b[i] - sum) / A2[i];
    }
    return true;
}

__device__ void epnp_gauss_newton(const double* L, const double* rho, double* betas) {
    double A[24], b[6], x[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (int it = 0; it < 5; it++) {
        for (int i = 0; i < 6; i++) {
            const double* rowL = L + i * 10;
            double* rowA = A + i * 4;
            rowA[0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
            rowA[1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
            rowA[2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
            rowA[3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
            b[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] + rowL[2] * betas[1] * betas[1] +
                             rowL[3] * betas[0] * betas[2] + rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                             rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] + rowL[8] * betas[2] * betas[3] +
                             rowL[9] * betas[3] * betas[3]);
        }
        epnp_qr_solve(A, b, x);
        for (int i = 0; i < 4; i++) betas[i] += x[i];
    }
}

// solvePnP(SOLVEPNP_EPNP) on the 5 correspondences idx[0..4]: undistortPoints stores the normalised image points as
// float32, epnp's constructor maps them back through the camera matrix in double.
// Phase 1 (one warp): control points, barycentric coordinates, M, M^T M by lane 0; the 12x12 SVD by the warp; L, rho.
__device__ void epnp5_prepare(EpnpWork& e, const float* xyz, const float* uv, const int* idx, double fx, double fy, double cx,
                              double cy, int lane) {
    const int n = 5;
    if (lane == 0) {
        e.fu = fx;
        e.fv = fy;
        e.uc = cx;
        e.vc = cy;
        const double ifx = 1. / fx, ify = 1. / fy;
        for (int i = 0; i < n; i++) {
            const int q = idx[i];
            e.pws[3 * i] = xyz[3 * q];
            e.pws[3 * i + 1] = xyz[3 * q + 1];
            e.pws[3 * i + 2] = xyz[3 * q + 2];
            const float xn = (float)(((double)uv[2 * q] - cx) * ifx), yn = (float)(((double)uv[2 * q + 1] - cy) * ify);
            e.us[2 * i] = xn * fx + cx;
            e.us[2 * i + 1] = yn * fy + cy;
        }
        epnp_choose_control_points(e, n);
        epnp_barycentric(e, n);
        for (int i = 0; i < n; i++) {  // fill_M
            double* M1 = e.M + 2 * i * 12;
            double* M2 = M1 + 12;
            const double* as = e.alphas + 4 * i;
            const double u = e.us[2 * i], v = e.us[2 * i + 1];
            for (int k = 0; k < 4; k++) {
                M1[3 * k] = as[k] * e.fu;
                M1[3 * k + 1] = 0.0;
                M1[3 * k + 2] = as[k] * (e.uc - u);
                M2[3 * k] = 0.0;
                M2[3 * k + 1] = as[k] * e.fv;
                M2[3 * k + 2] = as[k] * (e.vc - v);
            }
        }
        mul_transposed_dev(e.M, 2 * n, 12, e.At, EPNP_AS);  // symmetric: its transpose is itself
    }
    EPNP_SYNCWARP();
    jacobi_svd12_wave(e.At, e.W12, e.d12, lane);
    if (lane == 0) {
        epnp_L_6x10(e.At, e.L);
        e.rho[0] = dist2_dev(e.cws[0], e.cws[1]);
        e.rho[1] = dist2_dev(e.cws[0], e.cws[2]);
        e.rho[2] = dist2_dev(e.cws[0], e.cws[3]);
        e.rho[3] = dist2_dev(e.cws[1], e.cws[2]);
        e.rho[4] = dist2_dev(e.cws[1], e.cws[3]);
        e.rho[5] = dist2_dev(e.cws[2], e.cws[3]);
    }
    EPNP_SYNCWARP();
}

// Phase 2 (one thread per branch N = 1, 2, 3; the branches only read the work area): starting betas, five Gauss-Newton
// steps, pose and mean reprojection error
__device__ double epnp5_branch(const EpnpWork& e, int N, double* R /*[9]*/, double* t /*[3]*/) {
    double betas[4];
    epnp_betas_approx(N, e.L, e.rho, betas);
    epnp_gauss_newton(e.L, e.rho, betas);
    return epnp_R_and_t(e, 5, betas, R, t);
}

// Phase 3: OpenCV's choice  N = 1; if (err[2] < err[1]) N = 2; if (err[3] < err[N]) N = 3  (a NaN never wins)
__device__ __forceinline__ int epnp5_select(const double* err /*[3]*/) {
    int N = 0;
    if (err[1] < err[0]) N = 1;
    if (err[2] < err[N]) N = 2;
    return N;
}

// all phases in one thread (host build of the tests; lane 0 semantics)
__device__ void epnp5_dev(EpnpWork& e, const float* xyz, const float* uv, const int* idx, double fx, double fy, double cx,
                          double cy, double* R /*[9]*/, double* t /*[3]*/) {
    epnp5_prepare(e, xyz, uv, idx, fx, fy, cx, cy, 0);
    double err[3], Rn[3][9], tn[3][3];
    for (int N = 1; N <= 3; ++N) err[N - 1] = epnp5_branch(e, N, Rn[N - 1], tn[N - 1]);
    const int b = epnp5_select(err);
    for (int k = 0; k < 9; k++) R[k] = Rn[b][k];
    for (int k = 0; k < 3; k++) t[k] = tn[b][k];
}
