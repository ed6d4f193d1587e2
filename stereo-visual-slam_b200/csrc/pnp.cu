// K12: PnP-RANSAC hypothesis scoring + inlier refit, and K7 as a stand-alone call (ANMS on caller keypoints).
//
// Replaces cv::solvePnPRansac(pts3d, pts2d, K, noDist, rvec, tvec, false, 100, 4.0, 0.99, inliers) inside
// VO::motion_estimation (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:253-314, call at :277) and the
// reference's own VO::adaptive_non_maximal_suppresion (visual_odometry.cpp:96-157).
//
// Parity definition (SURVEY.md §7 hard part 3, §A.4): OpenCV's result is the Gauss-Newton optimum of the reprojection
// error over the inlier set selected by the best minimal hypothesis.  On inputs with a clear consensus every good
// hypothesis selects the same set, so this implementation matches cv2 on (inlier indices, pose) without reproducing
// OpenCV's RNG stream or its EPnP minimal solver:
//   pnp_hypothesis_kernel  one CTA per hypothesis: thread 0 solves a 6-point DLT (12x12 one-sided Jacobi SVD, polar
//                          orthogonalisation, 5 Gauss-Newton steps on the sample), all threads score the M points
//   pnp_refine_kernel      one CTA: first best hypothesis (strictly-greater rule like RANSACPointSetRegistrator),
//                          inlier mask at reprojection error <= 4 px, Gauss-Newton refit on the inliers (fp64, 6x6
//                          normal equations by block reduction), Rodrigues vector, ascending inlier index list
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

#define PNP_THREADS 256
#define PNP_MAX_HYP 512

struct PnpState {
    float* d_xyz;
    float* d_uv;
    double* d_hyp;    // [H][12] pose [R|t]
    int* d_cnt;       // [H]
    double* d_out;    // rvec(3) tvec(3) R|t (12)
    int* d_inl;       // [cap + 1]: count, then indices
    int cap;
    // ANMS staging
    vslam_keypoint* d_kp;
    double* d_rad;
    int* d_keep;
};

struct PnpCam {
    double fx, fy, cx, cy;
};

__device__ __forceinline__ uint32_t pnp_rng(uint32_t& s) {
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
}

// R <- nearest rotation (polar decomposition by Newton iteration R <- (R + R^-T)/2), det forced positive by the caller
__device__ void orthonormalize3(double* R) {
    for (int it = 0; it < 30; ++it) {
        const double c0 = R[4] * R[8] - R[5] * R[7], c1 = R[5] * R[6] - R[3] * R[8], c2 = R[3] * R[7] - R[4] * R[6];
        const double det = R[0] * c0 + R[1] * c1 + R[2] * c2;
        const double id = 1.0 / det;
        double T[9];  // inverse transpose = cofactor / det
        T[0] = c0 * id; T[1] = c1 * id; T[2] = c2 * id;
        T[3] = (R[2] * R[7] - R[1] * R[8]) * id; T[4] = (R[0] * R[8] - R[2] * R[6]) * id; T[5] = (R[1] * R[6] - R[0] * R[7]) * id;
        T[6] = (R[1] * R[5] - R[2] * R[4]) * id; T[7] = (R[2] * R[3] - R[0] * R[5]) * id; T[8] = (R[0] * R[4] - R[1] * R[3]) * id;
        double diff = 0;
        for (int i = 0; i < 9; ++i) {
            const double n = 0.5 * (R[i] + T[i]);
            diff += fabs(n - R[i]);
            R[i] = n;
        }
        if (diff < 1e-15) break;
    }
}

// solve the 6x6 SPD system H x = g in place (Cholesky), returns false if not positive definite
__device__ bool solve6(double* H, double* g) {
    for (int j = 0; j < 6; ++j) {
        double d = H[j * 6 + j];
        for (int k = 0; k < j; ++k) d -= H[j * 6 + k] * H[j * 6 + k];
        if (!(d > 0)) return false;
        d = sqrt(d);
        H[j * 6 + j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = H[i * 6 + j];
            for (int k = 0; k < j; ++k) s -= H[i * 6 + k] * H[j * 6 + k];
            H[i * 6 + j] = s / d;
        }
    }
    for (int i = 0; i < 6; ++i) {
        double s = g[i];
        for (int k = 0; k < i; ++k) s -= H[i * 6 + k] * g[k];
        g[i] = s / H[i * 6 + i];
    }
    for (int i = 5; i >= 0; --i) {
        double s = g[i];
        for (int k = i + 1; k < 6; ++k) s -= H[k * 6 + i] * g[k];
        g[i] = s / H[i * 6 + i];
    }
    return true;
}

// left-multiplicative SE3 update T <- exp(xi) T (same manifold as the BA, optimization.cpp:31)
__device__ void pose_oplus(double* T, const double* xi) {
    const double w0 = xi[3], w1 = xi[4], w2 = xi[5];
    const double th2 = w0 * w0 + w1 * w1 + w2 * w2, th = sqrt(th2);
    double A, B, C;
    if (th < 1e-10) {
        A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0;
    } else {
        A = sin(th) / th; B = (1 - cos(th)) / th2; C = (th - sin(th)) / (th2 * th);
    }
    const double W[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    double W2[9], dR[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W2[i * 3 + j] = W[i * 3] * W[j] + W[i * 3 + 1] * W[3 + j] + W[i * 3 + 2] * W[6 + j];
    for (int i = 0; i < 9; ++i) {
        dR[i] = A * W[i] + B * W2[i];
        V[i] = B * W[i] + C * W2[i];
    }
    dR[0] += 1; dR[4] += 1; dR[8] += 1;
    V[0] += 1; V[4] += 1; V[8] += 1;
    double Tn[12];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Tn[r * 4 + c] = dR[r * 3] * T[c] + dR[r * 3 + 1] * T[4 + c] + dR[r * 3 + 2] * T[8 + c];
        Tn[r * 4 + 3] = dR[r * 3] * T[3] + dR[r * 3 + 1] * T[7] + dR[r * 3 + 2] * T[11] + V[r * 3] * xi[0] + V[r * 3 + 1] * xi[1] + V[r * 3 + 2] * xi[2];
    }
    for (int i = 0; i < 12; ++i) T[i] = Tn[i];
}

// accumulate the Gauss-Newton normal equations of one correspondence into H (21 upper entries) and g (6)
__device__ __forceinline__ void gn_accumulate(const double* T, const PnpCam& cam, float X, float Y, float Z, float u,
                                              float v, double* H21, double* g, double& err2) {
    const double px = T[0] * X + T[1] * Y + T[2] * Z + T[3], py = T[4] * X + T[5] * Y + T[6] * Z + T[7],
                 pz = T[8] * X + T[9] * Y + T[10] * Z + T[11];
    const double iz = 1.0 / pz, iz2 = iz * iz;
    const double e0 = (double)u - (cam.fx * px * iz + cam.cx), e1 = (double)v - (cam.fy * py * iz + cam.cy);
    err2 = e0 * e0 + e1 * e1;
    double J0[6], J1[6];  // d e / d xi for T <- exp(xi) T  (optimization.cpp:68-71)
    J0[0] = -cam.fx * iz; J0[1] = 0; J0[2] = cam.fx * px * iz2; J0[3] = cam.fx * px * py * iz2;
    J0[4] = -cam.fx - cam.fx * px * px * iz2; J0[5] = cam.fx * py * iz;
    J1[0] = 0; J1[1] = -cam.fy * iz; J1[2] = cam.fy * py * iz2; J1[3] = cam.fy + cam.fy * py * py * iz2;
    J1[4] = -cam.fy * px * py * iz2; J1[5] = -cam.fy * px * iz;
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        g[a] -= J0[a] * e0 + J1[a] * e1;
#pragma unroll
        for (int b = a; b < 6; ++b) H21[q++] += J0[a] * J0[b] + J1[a] * J1[b];
    }
}

// 6-point DLT.  The 12x12 system's null vector is the eigenvector of the smallest eigenvalue of M = A^T A; it is found
// by inverse iteration on a Cholesky factor of M + mu I (the gap between the noise-level smallest eigenvalue and the
// next one makes a handful of iterations converge to machine precision), a few hundred flops instead of the thousands
// of dependent rotations of a one-sided Jacobi SVD that used to make this kernel the longest stage of a VO frame.
// The pose is then polished by Gauss-Newton on the sample itself, which removes what the normal equations lose.
__device__ void dlt6_rows(const float* xyz, const float* uv, const int* idx, const PnpCam& cam, double* A /*[12][12]*/) {
    for (int i = 0; i < 144; ++i) A[i] = 0;
    for (int s = 0; s < 6; ++s) {
        const int i = idx[s];
        const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
        const double x = ((double)uv[2 * i] - cam.cx) / cam.fx, y = ((double)uv[2 * i + 1] - cam.cy) / cam.fy;  // normalised
        double* r0 = A + 12 * (2 * s);
        double* r1 = A + 12 * (2 * s + 1);
        r0[0] = X; r0[1] = Y; r0[2] = Z; r0[3] = 1; r0[8] = -x * X; r0[9] = -x * Y; r0[10] = -x * Z; r0[11] = -x;
        r1[4] = X; r1[5] = Y; r1[6] = Z; r1[7] = 1; r1[8] = -y * X; r1[9] = -y * Y; r1[10] = -y * Z; r1[11] = -y;
    }
}

// M (12x12 symmetric, row-major, overwritten by its Cholesky factor) -> unit eigenvector of the smallest eigenvalue
__device__ bool smallest_eigvec12(double* M, double* x) {
    double tr = 0;
    for (int i = 0; i < 12; ++i) tr += M[i * 13];
    if (!(tr > 0) || !isfinite(tr)) return false;
    const double mu = 1e-13 * tr;
    for (int j = 0; j < 12; ++j) {  // lower Cholesky, in place
        double d = M[j * 13] + mu;
        for (int k = 0; k < j; ++k) d -= M[j * 12 + k] * M[j * 12 + k];
        if (!(d > 0)) d = mu;  // rank-deficient sample: keep going, the hypothesis will score badly
        const double l = sqrt(d), il = 1.0 / l;
        M[j * 13] = l;
        for (int i = j + 1; i < 12; ++i) {
            double v = M[i * 12 + j];
            for (int k = 0; k < j; ++k) v -= M[i * 12 + k] * M[j * 12 + k];
            M[i * 12 + j] = v * il;
        }
    }
    for (int i = 0; i < 12; ++i) x[i] = 0.28867513459481287 * ((i & 1) ? 1.0 : 0.9) * ((i % 3) ? 1.0 : 1.1);  // generic start
    for (int it = 0; it < 6; ++it) {
        for (int i = 0; i < 12; ++i) {  // L y = x
            double v = x[i];
            for (int k = 0; k < i; ++k) v -= M[i * 12 + k] * x[k];
            x[i] = v / M[i * 13];
        }
        for (int i = 11; i >= 0; --i) {  // L^T z = y
            double v = x[i];
            for (int k = i + 1; k < 12; ++k) v -= M[k * 12 + i] * x[k];
            x[i] = v / M[i * 13];
        }
        double nn = 0;
        for (int i = 0; i < 12; ++i) nn += x[i] * x[i];
        if (!(nn > 0) || !isfinite(nn)) return false;
        const double inv = 1.0 / sqrt(nn);
        for (int i = 0; i < 12; ++i) x[i] *= inv;
    }
    return true;
}

// pose from the DLT null vector P = s [R | t], then Gauss-Newton on the six correspondences
__device__ bool dlt6_pose(const float* xyz, const float* uv, const int* idx, const PnpCam& cam, const double* Pv, double* T) {
    double P[12];
    for (int k = 0; k < 12; ++k) P[k] = Pv[k];
    // P = s [R | t]: fix the sign with det(R) > 0, the scale with the mean row norm
    double R[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (!(fabs(det) > 1e-300)) return false;
    const double sc = cbrt(det);
    for (int i = 0; i < 9; ++i) R[i] /= sc;
    orthonormalize3(R);
    T[0] = R[0]; T[1] = R[1]; T[2] = R[2]; T[3] = P[3] / sc;
    T[4] = R[3]; T[5] = R[4]; T[6] = R[5]; T[7] = P[7] / sc;
    T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = P[11] / sc;
    for (int i = 0; i < 12; ++i)
        if (!isfinite(T[i])) return false;
    // a few Gauss-Newton steps on the sample itself
    for (int it = 0; it < 5; ++it) {
        double H21[21], g[6], H[36], e2;
        for (int i = 0; i < 21; ++i) H21[i] = 0;
        for (int i = 0; i < 6; ++i) g[i] = 0;
        for (int s = 0; s < 6; ++s) {
            const int i = idx[s];
            gn_accumulate(T, cam, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], uv[2 * i], uv[2 * i + 1], H21, g, e2);
        }
        int q = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b) { H[a * 6 + b] = H21[q]; H[b * 6 + a] = H21[q]; ++q; }
        for (int a = 0; a < 6; ++a) H[a * 6 + a] += 1e-9;
        if (!solve6(H, g)) return false;
        pose_oplus(T, g);
    }
    return true;
}

__global__ void __launch_bounds__(PNP_THREADS)
pnp_hypothesis_kernel(const float* __restrict__ xyz, const float* __restrict__ uv, int n, PnpCam cam, float thr2,
                      uint32_t seed, double* __restrict__ hyp, int* __restrict__ cnt) {
    __shared__ double sT[12], sA[144], sM[144];
    __shared__ int s_ok, s_cnt, s_idx[6];
    const int h = blockIdx.x;
    if (threadIdx.x == 0) {
        uint32_t s = seed * 2654435761u + (uint32_t)h * 40503u + 12345u;
        pnp_rng(s);
        int idx[6];
        for (int k = 0; k < 6; ++k) {  // 6 distinct indices
            for (;;) {
                const int c = (int)(pnp_rng(s) % (uint32_t)n);
                bool dup = false;
                for (int q = 0; q < k; ++q) dup |= idx[q] == c;
                if (!dup) { idx[k] = c; break; }
            }
        }
        for (int k = 0; k < 6; ++k) s_idx[k] = idx[k];
        dlt6_rows(xyz, uv, idx, cam, sA);
        s_cnt = 0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 144; e += PNP_THREADS) {  // M = A^T A, one entry per thread
        const int r = e / 12, c = e - r * 12;
        double v = 0;
#pragma unroll
        for (int k = 0; k < 12; ++k) v += sA[k * 12 + r] * sA[k * 12 + c];
        sM[e] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int idx[6];
        for (int k = 0; k < 6; ++k) idx[k] = s_idx[k];
        double x[12], T[12];
        bool ok = smallest_eigvec12(sM, x);
        if (ok) ok = dlt6_pose(xyz, uv, idx, cam, x, T);
        s_ok = ok ? 1 : 0;
        for (int i = 0; i < 12; ++i) sT[i] = ok ? T[i] : 0.0;
    }
    __syncthreads();
    int c = 0;
    if (s_ok) {
        for (int i = threadIdx.x; i < n; i += PNP_THREADS) {
            const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
            const double pz = sT[8] * X + sT[9] * Y + sT[10] * Z + sT[11];
            const double px = sT[0] * X + sT[1] * Y + sT[2] * Z + sT[3], py = sT[4] * X + sT[5] * Y + sT[6] * Z + sT[7];
            const double e0 = (double)uv[2 * i] - (cam.fx * px / pz + cam.cx), e1 = (double)uv[2 * i + 1] - (cam.fy * py / pz + cam.cy);
            c += (pz > 0 && e0 * e0 + e1 * e1 <= (double)thr2) ? 1 : 0;
        }
    }
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) {
        cnt[h] = s_ok ? s_cnt : -1;
        for (int i = 0; i < 12; ++i) hyp[12 * h + i] = sT[i];
    }
}

__global__ void __launch_bounds__(PNP_THREADS)
pnp_refine_kernel(const float* __restrict__ xyz, const float* __restrict__ uv, int n, PnpCam cam, float thr2, int n_hyp,
                  const double* __restrict__ hyp, const int* __restrict__ cnt, int max_iter, double* __restrict__ out,
                  int* __restrict__ inl) {
    __shared__ double sT[12];
    __shared__ double sH[PNP_THREADS / 32][28];
    __shared__ int s_best, s_stop, s_base, s_warp[PNP_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        int best = -1, bc = 4;  // RANSACPointSetRegistrator: goodCount > max(maxGoodCount, modelPoints - 1)
        for (int h = 0; h < n_hyp; ++h)
            if (cnt[h] > bc) { bc = cnt[h]; best = h; }
        s_best = best;
        if (best >= 0)
            for (int i = 0; i < 12; ++i) sT[i] = hyp[12 * best + i];
        s_stop = 0;
        s_base = 0;
    }
    __syncthreads();
    if (s_best < 0) {
        if (tid == 0) inl[0] = 0;
        return;
    }
    // inlier mask of the winning hypothesis (kept as flags in registers, recomputed per pass: n is small)
    const double T0[12] = {sT[0], sT[1], sT[2], sT[3], sT[4], sT[5], sT[6], sT[7], sT[8], sT[9], sT[10], sT[11]};
    auto is_inlier = [&](int i) {
        const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
        const double pz = T0[8] * X + T0[9] * Y + T0[10] * Z + T0[11];
        const double px = T0[0] * X + T0[1] * Y + T0[2] * Z + T0[3], py = T0[4] * X + T0[5] * Y + T0[6] * Z + T0[7];
        const double e0 = (double)uv[2 * i] - (cam.fx * px / pz + cam.cx), e1 = (double)uv[2 * i + 1] - (cam.fy * py / pz + cam.cy);
        return pz > 0 && e0 * e0 + e1 * e1 <= (double)thr2;
    };
    // ascending inlier index list
    for (int i0 = 0; i0 < n; i0 += PNP_THREADS) {
        const int i = i0 + tid;
        const bool k = i < n && is_inlier(i);
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, k);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < PNP_THREADS / 32; ++w) {
            before += w < warp ? s_warp[w] : 0;
            total += s_warp[w];
        }
        if (k) inl[1 + s_base + before + __popc(bal & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    // Gauss-Newton refit on the inliers
    for (int it = 0; it < max_iter; ++it) {
        double H21[21], g[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) H21[i] = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) g[i] = 0;
        double T[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) T[i] = sT[i];
        for (int i = tid; i < n; i += PNP_THREADS) {
            if (!is_inlier(i)) continue;
            double e2;
            gn_accumulate(T, cam, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], uv[2 * i], uv[2 * i + 1], H21, g, e2);
        }
#pragma unroll
        for (int i = 0; i < 21; ++i) {
            double v = H21[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (lane == 0) sH[warp][i] = v;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double v = g[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (lane == 0) sH[warp][21 + i] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double H[36], gg[6], acc[27];
            for (int i = 0; i < 27; ++i) {
                acc[i] = 0;
                for (int w = 0; w < PNP_THREADS / 32; ++w) acc[i] += sH[w][i];
            }
            int q = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b) { H[a * 6 + b] = acc[q]; H[b * 6 + a] = acc[q]; ++q; }
            for (int a = 0; a < 6; ++a) gg[a] = acc[21 + a];
            if (solve6(H, gg)) {
                double T2[12];
                for (int i = 0; i < 12; ++i) T2[i] = sT[i];
                pose_oplus(T2, gg);
                for (int i = 0; i < 12; ++i) sT[i] = T2[i];
                double nx = 0;
                for (int a = 0; a < 6; ++a) nx += gg[a] * gg[a];
                if (nx < 1e-28) s_stop = 1;
            } else {
                s_stop = 1;
            }
        }
        __syncthreads();
        if (s_stop) break;
    }
    if (tid == 0) {
        inl[0] = s_base;
        // Rodrigues vector of R (cv::Rodrigues inverse)
        const double* R = sT;
        const double tr = R[0] + R[5] + R[10];
        const double c = fmin(1.0, fmax(-1.0, 0.5 * (tr - 1.0)));
        const double th = acos(c);
        double w[3] = {R[9] - R[6], R[2] - R[8], R[4] - R[1]};
        double f = th < 1e-10 ? 0.5 : th / (2.0 * sin(th));
        if (M_PI - th < 1e-6) {
            double ax[3] = {sqrt(fmax(0.0, (R[0] - c) / (1 - c))), sqrt(fmax(0.0, (R[5] - c) / (1 - c))), sqrt(fmax(0.0, (R[10] - c) / (1 - c)))};
            for (int i = 0; i < 3; ++i) w[i] = (w[i] < 0 ? -ax[i] : ax[i]) * th;
            f = 1.0;
        }
        out[0] = w[0] * f; out[1] = w[1] * f; out[2] = w[2] * f;
        out[3] = sT[3]; out[4] = sT[7]; out[5] = sT[11];
        for (int i = 0; i < 12; ++i) out[6 + i] = sT[i];
    }
}

// ----------------------------------------------------------------------------------------------------------------
// K7 stand-alone: ANMS over a caller-supplied keypoint list (any order), one CTA.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
anms_points_kernel(const vslam_keypoint* __restrict__ kp, int n, int num, float c_robust, double* __restrict__ rad,
                   int* __restrict__ keep) {
    extern __shared__ unsigned long long s_sort[];
    __shared__ int s_base, s_warp[32];
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += 1024) {
        const float thr = __fmul_rn(kp[i].response, c_robust);
        const float xi = kp[i].x, yi = kp[i].y;
        double best = 1.7976931348623157e308;
        for (int j = 0; j < n; ++j) {
            if (!(kp[j].response > thr)) continue;
            const float dx = __fsub_rn(xi, kp[j].x), dy = __fsub_rn(yi, kp[j].y);
            best = fmin(best, sqrt(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy))));
        }
        rad[i] = best;
    }
    __syncthreads();
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = tid; i < n2; i += 1024)
        s_sort[i] = i < n ? ~(unsigned long long)__double_as_longlong(rad[i]) : 0xFFFFFFFFFFFFFFFFull;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += 1024) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long A = s_sort[i], B = s_sort[p];
                    if ((A > B) == ((i & k) == 0)) { s_sort[i] = B; s_sort[p] = A; }
                }
            }
            __syncthreads();
        }
    const double final_radius = __longlong_as_double((long long)~s_sort[num - 1]);
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        const bool k = i < n && rad[i] >= final_radius;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, k);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            before += w < (tid >> 5) ? s_warp[w] : 0;
            total += s_warp[w];
        }
        if (k) keep[1 + s_base + before + __popc(bal & ((1u << (tid & 31)) - 1))] = i;
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0) keep[0] = s_base;
}

// ----------------------------------------------------------------------------------------------------------------
static PnpState* pnp_state(vslam_ctx* ctx) { return ctx->pnp; }

int vslam_pnp_init(vslam_ctx* ctx) {
    PnpState* p = (PnpState*)calloc(1, sizeof(PnpState));
    if (!p) return VSLAM_E_INVALID;
    ctx->pnp = p;
    p->cap = ctx->cfg.max_keypoints > 0 ? ctx->cfg.max_keypoints : 1;
    const size_t cap = (size_t)p->cap;
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_xyz, cap * 12));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_uv, cap * 8));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_hyp, PNP_MAX_HYP * 12 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_cnt, PNP_MAX_HYP * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_out, 18 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_inl, (cap + 1) * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_kp, cap * sizeof(vslam_keypoint)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_rad, cap * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_keep, (cap + 1) * sizeof(int)));
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(anms_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    return VSLAM_OK;
}

void vslam_pnp_free(vslam_ctx* ctx) {
    PnpState* p = ctx->pnp;
    if (!p) return;
    cudaFree(p->d_xyz); cudaFree(p->d_uv); cudaFree(p->d_hyp); cudaFree(p->d_cnt); cudaFree(p->d_out);
    cudaFree(p->d_inl); cudaFree(p->d_kp); cudaFree(p->d_rad); cudaFree(p->d_keep);
    free(p);
    ctx->pnp = nullptr;
}

extern "C" int vslam_pnp_ransac(vslam_ctx* ctx, const float* xyz, const float* uv, int n, const double* Kmat, int iters,
                                float reproj_err, double confidence, double* rvec, double* tvec, double* T_c_w,
                                int32_t* inliers, int32_t* n_inliers) {
    (void)confidence;  // all `iters` hypotheses are evaluated (no early exit): see the header comment
    if (!ctx || !n_inliers || !Kmat || n < 0) return VSLAM_E_INVALID;
    *n_inliers = 0;
    if (n < 6) return VSLAM_OK;  // not enough correspondences for a hypothesis: no inliers, pose untouched
    if (!xyz || !uv || !rvec || !tvec || !inliers) return VSLAM_E_INVALID;
    PnpState* p = pnp_state(ctx);
    if (!p) return VSLAM_E_CAPACITY;
    if (n > p->cap) return VSLAM_E_CAPACITY;
    if (iters <= 0) iters = 100;
    if (iters > PNP_MAX_HYP) iters = PNP_MAX_HYP;
    PnpCam cam = {Kmat[0], Kmat[4], Kmat[2], Kmat[5]};
    const float thr2 = reproj_err * reproj_err;
    cudaStream_t s = ctx->stream;
    VSLAM_CUDA(ctx, cudaMemcpyAsync(p->d_xyz, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(p->d_uv, uv, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    vslam_time_begin(ctx, VK_PNP);
    pnp_hypothesis_kernel<<<iters, PNP_THREADS, 0, s>>>(p->d_xyz, p->d_uv, n, cam, thr2, 0x9E3779B9u, p->d_hyp, p->d_cnt);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "pnp_hypothesis_kernel");
    vslam_time_begin(ctx, VK_PNP_REFINE);
    pnp_refine_kernel<<<1, PNP_THREADS, 0, s>>>(p->d_xyz, p->d_uv, n, cam, thr2, iters, p->d_hyp, p->d_cnt, 30, p->d_out,
                                                p->d_inl);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "pnp_refine_kernel");
    double out[18];
    int cnt = 0;
    // one round trip: count, pose and the whole index buffer (entries past the count are unspecified, as documented)
    VSLAM_CUDA(ctx, cudaMemcpyAsync(&cnt, p->d_inl, sizeof(int), cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(out, p->d_out, sizeof(out), cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(inliers, p->d_inl + 1, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    *n_inliers = cnt;
    if (cnt <= 0) return VSLAM_OK;
    for (int i = 0; i < 3; ++i) {
        rvec[i] = out[i];
        tvec[i] = out[3 + i];
    }
    if (T_c_w)
        for (int i = 0; i < 12; ++i) T_c_w[i] = out[6 + i];
    return VSLAM_OK;
}

extern "C" int vslam_anms(vslam_ctx* ctx, const vslam_keypoint* keypoints, int n, int num, float c_robust,
                          int32_t* keep_idx, int32_t* n_keep) {
    if (!ctx || !n_keep || n < 0) return VSLAM_E_INVALID;
    *n_keep = 0;
    if (n == 0) return VSLAM_OK;
    if (!keypoints || !keep_idx) return VSLAM_E_INVALID;
    if (n < num || num <= 0) {  // reference: no-op when fewer than num keypoints (visual_odometry.cpp:100)
        for (int i = 0; i < n; ++i) keep_idx[i] = i;
        *n_keep = n;
        return VSLAM_OK;
    }
    PnpState* p = pnp_state(ctx);
    if (!p || n > p->cap) return VSLAM_E_CAPACITY;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    if ((size_t)n2 * 8 > 131072) return VSLAM_E_CAPACITY;
    cudaStream_t s = ctx->stream;
    VSLAM_CUDA(ctx, cudaMemcpyAsync(p->d_kp, keypoints, (size_t)n * sizeof(vslam_keypoint), cudaMemcpyHostToDevice, s));
    vslam_time_begin(ctx, VK_ANMS);
    anms_points_kernel<<<1, 1024, (size_t)n2 * 8, s>>>(p->d_kp, n, num, c_robust, p->d_rad, p->d_keep);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "anms_points_kernel");
    int cnt = 0;
    VSLAM_CUDA(ctx, cudaMemcpyAsync(&cnt, p->d_keep, sizeof(int), cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(keep_idx, p->d_keep + 1, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    *n_keep = cnt;
    return VSLAM_OK;
}
