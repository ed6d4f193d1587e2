// K13-K16: sliding-window bundle adjustment, the whole Levenberg-Marquardt optimisation in ONE persistent
// cooperative kernel (grid.sync between phases, LM control flow evaluated redundantly by every thread).
//
// Replaces optimize_map / optimize_pose_only (/root/reference/src/stereo_visual_slam_main/optimization.cpp:103-288,
// 290-436) together with the g2o machinery they run (OptimizationAlgorithmLevenberg + BlockSolver<6,3> with Schur
// complement + Huber kernel; un-vendored, restated in SURVEY.md §3.4/§A.5 and in oracle/ba_oracle.c) and the reference's
// own vertex/edge callbacks (optimization.cpp:26-101).  fp64 throughout; same control flow, same lambda schedule, same
// accept/reject rule as the oracle -- only summation order differs (documented tolerance 1e-4 relative on poses).
//
// Data layout: observations are sorted by landmark on the host (stable counting sort, O(n_obs) index marshalling), so
// a group of eight lanes owns one landmark: its 3x3 Hll block, bl and its Schur products never need atomics.  Pose blocks
// (Hpp, bp) are accumulated in shared memory per CTA and flushed with fp64 atomics; the reduced camera system S
// (6K x 6K) is accumulated per CTA in shared memory and reduced deterministically through a partial buffer when it
// fits (6K <= 96), with global fp64 atomics otherwise.  The dense SPD solve (K15) runs in one CTA, in shared memory
// when 6K <= 160.  The phase functions are also exposed one by one (vslam_ba_phase_*) for the multi-GPU path, where
// the host inserts one all-reduce of [S, b, chi2] per LM trial between them.
#include "common.cuh"

#include <cooperative_groups.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include <vector>

namespace cg = cooperative_groups;

#define BA_THREADS 256
#define BA_SMEM_CHOL_MAX 160 // Cholesky in shared memory up to this dimension

struct BaScalars {
    double chi_cur;                    // robust chi2 at the linearisation point (accumulated by BUILD)
    double chi_trial[2];               // alternating slots, see header comment of ba_lm_kernel
    double scale[2];
    unsigned long long maxdiag_bits;   // max |diag| as ordered bits (positive doubles order like integers)
    int solve_ok[2];
    int cnt_le[6];                     // relabel: #obs with chi2 <= th * 2^r
    // results
    int iterations, trials, accepted;
    double chi2_initial, chi2_final, lambda_final, chi2_threshold;
    int n_inlier_obs, n_outlier_obs;
    int pad;
    unsigned long long phase_ns[12];  // device-side profile of the persistent kernel (globaltimer, thread 0)
};

struct BaParams {
    int K, L, n_obs, pose_only, num_iterations, max_trials;
    double delta, tau, chi2_th;
    double Kc[9];
    int n;                 // 6K
    int n_cta;             // grid size (partials)
    int shard_L0, shard_L1;  // landmark range owned by this rank (multi-GPU); [0, L) on one GPU
    double* poses;         // [2][K][12]
    double* points;        // [2][L][3]
    const int* obs_pose;   // sorted by landmark
    const int* obs_point;
    const double* obs_uv;  // [n_obs][2]
    const int* obs_orig;   // original (insertion-order) index of sorted observation i
    const int* lm_start;   // [L+1]
    const int* pose_start; // [K+1]   pose-major CSR over the (landmark-sorted) observation indices
    const int* pose_obs;   // [n_obs]
    int* obs_of;           // [K][L]  observation index of (pose, landmark) or -1 (filled by BUILD)
    double* dbl;           // [L][3]  Dinv * bl
    double* Ubuf;          // [n][n] Cholesky factor rows of the grid-wide solver
    int* block_flag;       // [K(K+1)/2] 1 if some landmark joins poses (ki, kj): set once by the first BUILD
    int has_dup;           // some landmark is observed twice by one pose: use the atomic Schur path
    double* err;           // [n_obs][2]  the edges' _error (last computed, trial or not -- as in g2o)
    double* Hpl;           // [n_obs][18]
    double* Hll;           // [L][9]
    double* bl;            // [L][3]
    double* Dinv;          // [L][9]
    double* Hpp;           // [K][36]
    double* bp;            // [6K]
    const double* bp_scale;  // gradient used by computeScale: bp, or the sum over ranks in the multi-device kernel
    double* S;             // [n][n]
    double* bs;            // [n]
    double* x;             // [n + 3L]
    BaScalars* sc;
    double* chi2_out;      // [n_obs] original order
    uint8_t* inlier_out;   // [L]
};

// ---------------------------------------------------------------------------------------------------------------
// small fp64 helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void se3_exp_dev(const double* xi, double* R, double* t) {
    const double u0 = xi[0], u1 = xi[1], u2 = xi[2], w0 = xi[3], w1 = xi[4], w2 = xi[5];
    const double th2 = w0 * w0 + w1 * w1 + w2 * w2;
    const double th = sqrt(th2);
    double imag, real;
    if (th < 1e-10) {
        const double th4 = th2 * th2;
        imag = 0.5 - th2 / 48.0 + th4 / 3840.0;
        real = 1.0 - th2 / 8.0 + th4 / 384.0;
    } else {
        double sh, ch;
        sincos(0.5 * th, &sh, &ch);
        imag = sh / th;
        real = ch;
    }
    const double qx = imag * w0, qy = imag * w1, qz = imag * w2, qw = real;
    R[0] = 1 - 2 * (qy * qy + qz * qz); R[1] = 2 * (qx * qy - qz * qw);     R[2] = 2 * (qx * qz + qy * qw);
    R[3] = 2 * (qx * qy + qz * qw);     R[4] = 1 - 2 * (qx * qx + qz * qz); R[5] = 2 * (qy * qz - qx * qw);
    R[6] = 2 * (qx * qz - qy * qw);     R[7] = 2 * (qy * qz + qx * qw);     R[8] = 1 - 2 * (qx * qx + qy * qy);
    double V[9];
    if (th < 1e-10) {
#pragma unroll
        for (int i = 0; i < 9; ++i) V[i] = R[i];
    } else {
        double s, c;
        sincos(th, &s, &c);
        const double a = (1 - c) / th2, b = (th - s) / (th2 * th);
        // Omega = hat(w), Omega^2 = w w^T - |w|^2 I
        V[0] = 1 + b * (w0 * w0 - th2); V[1] = -a * w2 + b * w0 * w1;   V[2] = a * w1 + b * w0 * w2;
        V[3] = a * w2 + b * w0 * w1;    V[4] = 1 + b * (w1 * w1 - th2); V[5] = -a * w0 + b * w1 * w2;
        V[6] = -a * w1 + b * w0 * w2;   V[7] = a * w0 + b * w1 * w2;    V[8] = 1 + b * (w2 * w2 - th2);
    }
    t[0] = V[0] * u0 + V[1] * u1 + V[2] * u2;
    t[1] = V[3] * u0 + V[4] * u1 + V[5] * u2;
    t[2] = V[6] * u0 + V[7] * u1 + V[8] * u2;
}

__device__ __forceinline__ void huber_dev(double e2, double delta, double& rho0, double& rho1) {
    const double dsqr = delta * delta;
    if (e2 <= dsqr) {
        rho0 = e2;
        rho1 = 1.0;
    } else {
        const double s = sqrt(e2);
        rho0 = 2 * s * delta - dsqr;
        rho1 = delta / s;
    }
}

// e = z - pi(K (T p)); returns camera-frame point
__device__ __forceinline__ void residual_dev(const double* T, const double* p, const double* Kc, double u, double v,
                                             double& e0, double& e1, double* pc) {
    pc[0] = T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[3];
    pc[1] = T[4] * p[0] + T[5] * p[1] + T[6] * p[2] + T[7];
    pc[2] = T[8] * p[0] + T[9] * p[1] + T[10] * p[2] + T[11];
    const double q0 = Kc[0] * pc[0] + Kc[1] * pc[1] + Kc[2] * pc[2];
    const double q1 = Kc[3] * pc[0] + Kc[4] * pc[1] + Kc[5] * pc[2];
    const double q2 = Kc[6] * pc[0] + Kc[7] * pc[1] + Kc[8] * pc[2];
    e0 = u - q0 / q2;
    e1 = v - q1 / q2;
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define BA_TICK(slot)                                        \
    do {                                                     \
        if (gtid == 0) {                                     \
            const unsigned long long now__ = gtimer();       \
            sc->phase_ns[slot] += now__ - t_last;            \
            t_last = now__;                                  \
        }                                                    \
    } while (0)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// phases (grid-wide; called either from the persistent kernel or one kernel per phase)
// ---------------------------------------------------------------------------------------------------------------
// ZERO: clear the accumulators the BUILD phases add into
__device__ void ba_phase_zero(const BaParams& P, int gtid, int gsize) {
    for (int i = gtid; i < P.K * 36; i += gsize) P.Hpp[i] = 0.0;
    for (int i = gtid; i < P.n; i += gsize) P.bp[i] = 0.0;
    if (!P.pose_only) {
        for (int i = P.shard_L0 * 9 + gtid; i < P.shard_L1 * 9; i += gsize) P.Hll[i] = 0.0;
        for (int i = P.shard_L0 * 3 + gtid; i < P.shard_L1 * 3; i += gsize) P.bl[i] = 0.0;
    }
    if (gtid == 0) {
        P.sc->chi_cur = 0.0;
        P.sc->maxdiag_bits = 0ull;
    }
}

// residual, Huber weight and pose Jacobian of one edge (optimization.cpp:41-73 / 75-101)
__device__ __forceinline__ void ba_edge(const BaParams& P, const double* T, const double* p, double u, double v,
                                        double& e0, double& e1, double& r0, double& w, double* A) {
    double pc[3];
    residual_dev(T, p, P.Kc, u, v, e0, e1, pc);
    huber_dev(e0 * e0 + e1 * e1, P.delta, r0, w);
    const double fx = P.Kc[0], fy = P.Kc[4];
    const double X = pc[0], Y = pc[1], Z = pc[2];
    if (!P.pose_only) {  // optimization.cpp:52-73
        const double Zinv = 1.0 / (Z + 1e-18), Zinv2 = Zinv * Zinv;
        A[0] = -fx * Zinv; A[1] = 0; A[2] = fx * X * Zinv2; A[3] = fx * X * Y * Zinv2;
        A[4] = -fx - fx * X * X * Zinv2; A[5] = fx * Y * Zinv;
        A[6] = 0; A[7] = -fy * Zinv; A[8] = fy * Y * Zinv2; A[9] = fy + fy * Y * Y * Zinv2;
        A[10] = -fy * X * Y * Zinv2; A[11] = -fy * X * Zinv;
    } else {             // optimization.cpp:84-101
        const double Z2 = Z * Z;
        A[0] = -fx / Z; A[1] = 0; A[2] = fx * X / Z2; A[3] = fx * X * Y / Z2; A[4] = -fx - fx * X * X / Z2;
        A[5] = fx * Y / Z;
        A[6] = 0; A[7] = -fy / Z; A[8] = fy * Y / (Z * Z); A[9] = fy + fy * Y * Y / Z2; A[10] = -fy * X * Y / Z2;
        A[11] = -fy * X / Z;
    }
}

// BUILD-A (K13, per landmark): computeActiveErrors + activeRobustChi2 + the landmark side of buildSystem.
// Eight lanes own one landmark (observations are sorted by landmark): lane q takes the landmark's edges q, q + 8, ...,
// the 3x3 block and gradient are summed over the group with a fixed xor butterfly and written once -- no atomics, the
// same summation order on every run.  Hpl (6x3) is stored per edge; obs_of records which edge joins (pose, landmark).
#define BA_LM_GROUP 8
__device__ void ba_phase_build_edges(const BaParams& P, int cur, int gtid, int gsize) {
    const double* poses = P.poses + (size_t)cur * P.K * 12;
    const double* points = P.points + (size_t)cur * P.L * 3;
    double chi = 0.0;
    const int sub = gtid & (BA_LM_GROUP - 1);
    const int ngroups = gsize / BA_LM_GROUP;
    const int nl = P.shard_L1 - P.shard_L0;
    // every lane of a warp runs the same number of iterations (the butterflies below are warp-wide)
    const int iters = (nl + ngroups - 1) / ngroups;
    for (int itl = 0; itl < iters; ++itl) {
        const int l = P.shard_L0 + itl * ngroups + gtid / BA_LM_GROUP;
        const bool live = l < P.shard_L1;
        const int o0 = live ? P.lm_start[l] : 0, o1 = live ? P.lm_start[l + 1] : 0;
        double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // Hll upper triangle (6) + bl (3)
        const double p[3] = {live ? points[3 * l] : 0.0, live ? points[3 * l + 1] : 0.0, live ? points[3 * l + 2] : 0.0};
        for (int i = o0 + sub; i < o1; i += BA_LM_GROUP) {
            const int k = P.obs_pose[i];
            const double* T = poses + 12 * k;
            double e0, e1, r0, w, A[12];
            ba_edge(P, T, p, P.obs_uv[2 * i], P.obs_uv[2 * i + 1], e0, e1, r0, w, A);
            P.err[2 * i] = e0;
            P.err[2 * i + 1] = e1;
            chi += r0;
            P.obs_of[(size_t)k * P.L + l] = i;
            if (!P.pose_only) {
                // which (ki <= kj) blocks of the reduced system are non-empty: idempotent stores, structure is fixed
                for (int j = o0; j < o1; ++j) {
                    const int kj = P.obs_pose[j];
                    if (kj >= k) P.block_flag[k * P.K - k * (k - 1) / 2 + (kj - k)] = 1;
                }
                const double om0 = -w * e0, om1 = -w * e1;
                double B[6];
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) B[r * 3 + c] = A[r * 6] * T[c] + A[r * 6 + 1] * T[4 + c] + A[r * 6 + 2] * T[8 + c];
                h[0] += w * (B[0] * B[0] + B[3] * B[3]);
                h[1] += w * (B[0] * B[1] + B[3] * B[4]);
                h[2] += w * (B[0] * B[2] + B[3] * B[5]);
                h[3] += w * (B[1] * B[1] + B[4] * B[4]);
                h[4] += w * (B[1] * B[2] + B[4] * B[5]);
                h[5] += w * (B[2] * B[2] + B[5] * B[5]);
                h[6] += B[0] * om0 + B[3] * om1;
                h[7] += B[1] * om0 + B[4] * om1;
                h[8] += B[2] * om0 + B[5] * om1;
                double* hp = P.Hpl + 18 * (size_t)i;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) hp[a * 3 + c] = w * (A[a] * B[c] + A[6 + a] * B[3 + c]);
            }
        }
        if (!P.pose_only) {
#pragma unroll
            for (int q = 0; q < 9; ++q) {
#pragma unroll
                for (int o = BA_LM_GROUP / 2; o > 0; o >>= 1) h[q] += __shfl_xor_sync(0xFFFFFFFFu, h[q], o);
            }
            if (live && sub == 0 && o1 > o0) {  // ZERO cleared the blocks of landmarks without edges
                double* H = P.Hll + 9 * (size_t)l;
                H[0] = h[0]; H[1] = h[1]; H[2] = h[2]; H[4] = h[3]; H[5] = h[4]; H[8] = h[5];
                double* bb = P.bl + 3 * (size_t)l;
                bb[0] = h[6]; bb[1] = h[7]; bb[2] = h[8];
            }
        }
    }
    chi = warp_sum(chi);
    if ((threadIdx.x & 31) == 0 && chi != 0.0) atomicAdd(&P.sc->chi_cur, chi);
}

// BUILD-B (K13, per pose): Hpp_kk = sum w A^T A, bp_k = sum A^T (-w e) over the edges of pose k.  A CTA owns one
// (pose, slice) work item: every thread accumulates its edges in registers (27 values: upper triangle + gradient), one
// block reduction, and at most n_cta / K partial sums per address reach L2 -- instead of one atomic per edge per value.
__device__ void ba_phase_build_poses(const BaParams& P, int cur, double* smem) {
    const double* poses = P.poses + (size_t)cur * P.K * 12;
    const double* points = P.points + (size_t)cur * P.L * 3;
    const int parts = max(1, (int)gridDim.x / P.K);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int wi = blockIdx.x; wi < P.K * parts; wi += gridDim.x) {
        const int k = wi % P.K, part = wi / P.K;
        const int ps = P.pose_start[k], pe = P.pose_start[k + 1];
        const double* T = poses + 12 * k;
        double acc[27];
#pragma unroll
        for (int q = 0; q < 27; ++q) acc[q] = 0.0;
        for (int t = ps + part * BA_THREADS + tid; t < pe; t += parts * BA_THREADS) {
            const int i = P.pose_obs[t];
            const int l = P.obs_point[i];
            if (l < P.shard_L0 || l >= P.shard_L1) continue;
            const double p[3] = {points[3 * l], points[3 * l + 1], points[3 * l + 2]};
            double e0, e1, r0, w, A[12];
            ba_edge(P, T, p, P.obs_uv[2 * i], P.obs_uv[2 * i + 1], e0, e1, r0, w, A);
            const double om0 = -w * e0, om1 = -w * e1;
            int q = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
#pragma unroll
                for (int b = a; b < 6; ++b) acc[q++] += w * (A[a] * A[b] + A[6 + a] * A[6 + b]);
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[21 + a] += A[a] * om0 + A[6 + a] * om1;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 27; ++q) {
            const double v = warp_sum(acc[q]);
            if (lane == 0) smem[warp * 27 + q] = v;
        }
        __syncthreads();
        if (tid < 27) {
            double v = 0.0;
            for (int w8 = 0; w8 < BA_THREADS / 32; ++w8) v += smem[w8 * 27 + tid];
            if (v != 0.0) {
                if (tid < 21) {
                    int a = 0, rem = tid;  // unpack upper-triangle index -> (a, b)
                    while (rem >= 6 - a) { rem -= 6 - a; ++a; }
                    atomicAdd(&P.Hpp[36 * k + a * 6 + a + rem], v);
                } else {
                    atomicAdd(&P.bp[6 * k + tid - 21], v);
                }
            }
        }
    }
}

// MAXDIAG: max |diagonal| over this rank's landmark blocks (feeds computeLambdaInit); after BUILD-A + grid.sync
__device__ void ba_phase_maxdiag(const BaParams& P, int gtid, int gsize) {
    if (P.pose_only) return;
    double mdiag = 0.0;
    for (int l = P.shard_L0 + gtid; l < P.shard_L1; l += gsize) {
        if (P.lm_start[l + 1] == P.lm_start[l]) continue;
        const double* H = P.Hll + 9 * (size_t)l;
        mdiag = fmax(mdiag, fmax(fabs(H[0]), fmax(fabs(H[4]), fabs(H[8]))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mdiag = fmax(mdiag, __shfl_xor_sync(0xFFFFFFFFu, mdiag, o));
    if ((threadIdx.x & 31) == 0 && mdiag > 0.0)
        atomicMax(&P.sc->maxdiag_bits, (unsigned long long)__double_as_longlong(mdiag));
}

__device__ void ba_phase_schur_init(const BaParams& P, double lambda, int add_hpp, int gtid, int gsize);

// DINV (K14a): per landmark Dinv = (Hll + lambda I)^-1 by the direct 3x3 inverse, and Dinv * bl; clears this trial's
// accumulator slots.
__device__ void ba_phase_dinv(const BaParams& P, double lambda, int slot, int add_hpp, int gtid, int gsize) {
    if (gtid == 0) {
        P.sc->chi_trial[slot] = 0.0;
        P.sc->scale[slot] = 0.0;
        P.sc->solve_ok[slot] = 0;
    }
    ba_phase_schur_init(P, lambda, add_hpp, gtid, gsize);  // S = [Hpp + lambda I], bs = bp; SCHUR subtracts from it
    if (P.pose_only) return;
    for (int l = P.shard_L0 + gtid; l < P.shard_L1; l += gsize) {
        double* Di = P.Dinv + 9 * (size_t)l;
        double* db = P.dbl + 3 * (size_t)l;
        if (P.lm_start[l + 1] == P.lm_start[l]) {
#pragma unroll
            for (int q = 0; q < 9; ++q) Di[q] = 0.0;
            db[0] = db[1] = db[2] = 0.0;
            continue;
        }
        const double* H = P.Hll + 9 * (size_t)l;
        const double m0 = H[0] + lambda, m1 = H[1], m2 = H[2], m4 = H[4] + lambda, m5 = H[5], m8 = H[8] + lambda;
        const double c0 = m4 * m8 - m5 * m5, c1 = m5 * m2 - m1 * m8, c2 = m1 * m5 - m4 * m2;
        const double id = 1.0 / (m0 * c0 + m1 * c1 + m2 * c2);
        double d[9];
        d[0] = c0 * id; d[1] = c1 * id; d[2] = c2 * id;
        d[3] = d[1];    d[4] = (m0 * m8 - m2 * m2) * id; d[5] = (m2 * m1 - m0 * m5) * id;
        d[6] = d[2];    d[7] = d[5];                     d[8] = (m0 * m4 - m1 * m1) * id;
#pragma unroll
        for (int q = 0; q < 9; ++q) Di[q] = d[q];
        const double b0 = P.bl[3 * l], b1 = P.bl[3 * l + 1], b2 = P.bl[3 * l + 2];
        db[0] = d[0] * b0 + d[1] * b1 + d[2] * b2;
        db[1] = d[3] * b0 + d[4] * b1 + d[5] * b2;
        db[2] = d[6] * b0 + d[7] * b1 + d[8] * b2;
    }
}

// SCHUR (K14b): the reduced camera system, one 6x6 block (ki <= kj, upper block triangle as g2o) per CTA work item:
//   S_ij = [i==j] (Hpp_ii + lambda I) - sum over landmarks seen by both  Hpl_i Dinv Hpl_j^T,   bs_i = bp_i - sum Hpl_i Dinv bl
// A thread walks (a slice of) pose ki's edges, finds the partner edge of pose kj on the same landmark through obs_of
// and accumulates the 36 (+6) products in registers; one block reduction per work item, then one fp64 reduction per
// value into the S initialised by DINV (n_cta / n_blocks partial sums per address at most).  Blocks that no landmark
// joins are skipped through block_flag.
__device__ void ba_phase_schur_blocks(const BaParams& P, double* smem) {
    if (P.pose_only) return;
    const int n = P.n, K = P.K;
    const int nb = K * (K + 1) / 2;
    const int parts = max(1, (int)gridDim.x / nb);  // CTAs per block when there are fewer blocks than CTAs
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int wi = blockIdx.x; wi < nb * parts; wi += gridDim.x) {
        const int b = wi % nb, part = wi / nb;
        if (!P.block_flag[b]) continue;  // uniform per CTA
        int ki = 0, rem = b;             // unpack the upper-triangle block index
        while (rem >= K - ki) { rem -= K - ki; ++ki; }
        const int kj = ki + rem;
        double acc[42];
#pragma unroll
        for (int q = 0; q < 42; ++q) acc[q] = 0.0;
        const int ps = P.pose_start[ki], pe = P.pose_start[ki + 1];
        for (int t = ps + part * BA_THREADS + tid; t < pe; t += parts * BA_THREADS) {
            const int i = P.pose_obs[t];
            const int l = P.obs_point[i];
            if (l < P.shard_L0 || l >= P.shard_L1) continue;
            const int j = (ki == kj) ? i : P.obs_of[(size_t)kj * P.L + l];
            if (j < 0) continue;
            const double* Bi = P.Hpl + 18 * (size_t)i;
            const double* d = P.Dinv + 9 * (size_t)l;
            double BD[18];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double x0 = Bi[a * 3], x1 = Bi[a * 3 + 1], x2 = Bi[a * 3 + 2];
                BD[a * 3] = x0 * d[0] + x1 * d[3] + x2 * d[6];
                BD[a * 3 + 1] = x0 * d[1] + x1 * d[4] + x2 * d[7];
                BD[a * 3 + 2] = x0 * d[2] + x1 * d[5] + x2 * d[8];
            }
            const double* Bj = P.Hpl + 18 * (size_t)j;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    acc[a * 6 + c] += BD[a * 3] * Bj[c * 3] + BD[a * 3 + 1] * Bj[c * 3 + 1] + BD[a * 3 + 2] * Bj[c * 3 + 2];
            if (ki == kj) {
                const double* db = P.dbl + 3 * (size_t)l;
#pragma unroll
                for (int a = 0; a < 6; ++a) acc[36 + a] += Bi[a * 3] * db[0] + Bi[a * 3 + 1] * db[1] + Bi[a * 3 + 2] * db[2];
            }
        }
        __syncthreads();
        const int nacc = (ki == kj) ? 42 : 36;
#pragma unroll
        for (int q = 0; q < 42; ++q) {
            if (q < nacc) {
                const double v = warp_sum(acc[q]);
                if (lane == 0) smem[warp * 42 + q] = v;
            }
        }
        __syncthreads();
        if (tid < nacc) {
            double v = 0.0;
            for (int w8 = 0; w8 < BA_THREADS / 32; ++w8) v += smem[w8 * 42 + tid];
            if (v != 0.0) {
                if (tid < 36) atomicAdd(&P.S[(size_t)(6 * ki + tid / 6) * n + 6 * kj + tid % 6], -v);
                else atomicAdd(&P.bs[6 * ki + tid - 36], -v);
            }
        }
    }
}

// SCHUR fallback when a landmark has two edges to the same pose (obs_of cannot hold both): initialise S, then one
// thread per edge adds its products with fp64 reductions at L2.
__device__ void ba_phase_schur_init(const BaParams& P, double lambda, int add_hpp, int gtid, int gsize) {
    const int n = P.n;
    for (int i = gtid; i < n * n; i += gsize) {
        const int r = i / n, c = i - r * n;
        double v = 0.0;
        if (add_hpp && r / 6 == c / 6 && c >= r) v = P.Hpp[(r / 6) * 36 + (r % 6) * 6 + (c % 6)];
        if (add_hpp == 1 && r == c) v += lambda;  // add_hpp == 2: this rank's partial Hpp only, lambda after the exchange
        P.S[i] = v;
    }
    for (int i = gtid; i < n; i += gsize) P.bs[i] = add_hpp ? P.bp[i] : 0.0;
}

__device__ void ba_phase_schur_atomic(const BaParams& P, int gtid, int gsize) {
    if (P.pose_only) return;
    const int n = P.n;
    const int i0 = P.lm_start[P.shard_L0], i1 = P.lm_start[P.shard_L1];
    for (int i = i0 + gtid; i < i1; i += gsize) {
        const int l = P.obs_point[i];
        const int o0 = P.lm_start[l], o1 = P.lm_start[l + 1];
        const double* d = P.Dinv + 9 * (size_t)l;
        const double* db = P.dbl + 3 * (size_t)l;
        const int ki = P.obs_pose[i];
        const double* Bi = P.Hpl + 18 * (size_t)i;
        double BD[18];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double x0 = Bi[a * 3], x1 = Bi[a * 3 + 1], x2 = Bi[a * 3 + 2];
            BD[a * 3] = x0 * d[0] + x1 * d[3] + x2 * d[6];
            BD[a * 3 + 1] = x0 * d[1] + x1 * d[4] + x2 * d[7];
            BD[a * 3 + 2] = x0 * d[2] + x1 * d[5] + x2 * d[8];
            atomicAdd(&P.bs[6 * ki + a], -(x0 * db[0] + x1 * db[1] + x2 * db[2]));
        }
        for (int j = o0; j < o1; ++j) {
            const int kj = P.obs_pose[j];
            if (kj < ki) continue;
            const double* Bj = P.Hpl + 18 * (size_t)j;
            double* dst = P.S + (size_t)(6 * ki) * n + 6 * kj;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    if (ki == kj && c < a) continue;
                    atomicAdd(&dst[a * n + c], -(BD[a * 3] * Bj[c * 3] + BD[a * 3 + 1] * Bj[c * 3 + 1] + BD[a * 3 + 2] * Bj[c * 3 + 2]));
                }
        }
    }
}

// ---- K15: dense SPD solve of the reduced camera system, 6x6-blocked (the natural block size of the pose system) ----
// Upper Cholesky S = U^T U.  Right-looking over block rows: (1) one thread factors the 6x6 diagonal block, (2) one
// thread per trailing column forward-substitutes its 6 panel entries, (3) all threads apply the rank-6 update.

// factor the 6x6 diagonal block at (j0, j0) of the row-major matrix M (leading dimension ld) in place; false if not PD
// rinv receives the reciprocals of the diagonal of U: every later division by U_rr becomes a multiplication
// (fp64 division and sqrt are long multi-instruction sequences; one rsqrt per pivot replaces a sqrt and a division).
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
        ri = ri * (1.5 - 0.5 * d * ri * ri);  // one Newton step: full double accuracy
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

// branch-free reciprocal and reciprocal square root for well-scaled positive doubles (Hessian pivots): hardware seed
// (about 20 bits) + two Newton steps to full precision.  Unlike 1.0 / d and rsqrt(d) they carry no special-case branch,
// so independent ones interleave in the instruction stream.
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

// Root-free factorisation of a 6x6 diagonal block with the square roots OFF the dependency chain: on return D holds the
// unnormalised rows u'[r][c] (u'[r][r] = d_r), G[p][r] = u'[p][r] / d_p and rs[r] = 1 / sqrt(d_r); the Cholesky factor is
// U = diag(rs) u'.  The chain per row is one reciprocal, one multiply and one DFMA; the six rs[] are independent of it.
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

// small systems (6K <= 160): everything in the shared memory of CTA 0.  The right-hand side rides along as column n of
// the augmented matrix [S | bs] (so the forward substitution happens inside the factorisation), every thread factors
// the 6x6 diagonal block redundantly in registers (no serial single-thread stage), and the backward substitution is
// blocked the same way: 3 barriers per block row in total.
__device__ void ba_phase_solve_cta(const BaParams& P, int slot, double* smem) {
    if (blockIdx.x != 0) return;
    const int n = P.n, ld = P.n + 1, tid = threadIdx.x, nt = blockDim.x;
    __shared__ int s_fail;
    __shared__ double s_rd[BA_SMEM_CHOL_MAX];  // reciprocal diagonal of U, for the backward substitution
    double* M = smem;  // n x (n + 1)
    if (tid == 0) s_fail = 0;
    for (int base = 0; base < n * n; base += 8 * nt) {  // eight independent L2 reads in flight per thread
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = base + k * nt + tid;
            v[k] = i < n * n ? P.S[i] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = base + k * nt + tid;
            if (i < n * n) M[(i / n) * ld + i % n] = v[k];
        }
    }
    for (int i = tid; i < n; i += nt) M[i * ld + n] = P.bs[i];
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += 6) {
        {
            double D[36];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) D[r * 6 + c] = (c >= r) ? M[(j0 + r) * ld + j0 + c] : 0.0;
            double rinv[6], G[36];
            const bool ok = ldl6_diag(D, rinv, G);  // root-free, branch-free reciprocals (see ba_small_solve_t)
            for (int c = j0 + 6 + tid; c <= n; c += nt) {  // panel (and the rhs column c == n)
                double v[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) v[r] = M[(j0 + r) * ld + c];
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int p = 0; p < 6; ++p)
                        if (p < r) v[r] -= G[p * 6 + r] * v[p];
#pragma unroll
                for (int r = 0; r < 6; ++r) M[(j0 + r) * ld + c] = v[r] * rinv[r];
            }
            __syncthreads();  // block row done; everyone has read the un-factored diagonal block
            if (tid == 0) {   // nobody reads this diagonal block again before the backward substitution
                if (!ok) s_fail = 1;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
#pragma unroll
                    for (int c = 0; c < 6; ++c)
                        if (c >= r) M[(j0 + r) * ld + j0 + c] = D[r * 6 + c] * rinv[r];
                    s_rd[j0 + r] = rinv[r];
                }
            }
        }
        const int m = n - j0 - 6;  // trailing update of the upper triangle and of the rhs column
        for (int e = tid; e < m * (m + 1); e += nt) {
            const int r = j0 + 6 + e / (m + 1), c = j0 + 6 + e % (m + 1);
            if (c < r) continue;
            double v = M[r * ld + c];
#pragma unroll
            for (int p = 0; p < 6; ++p) v -= M[(j0 + p) * ld + r] * M[(j0 + p) * ld + c];
            M[r * ld + c] = v;
        }
        __syncthreads();
        if (s_fail) break;
    }
    const int fail = s_fail;
    if (!fail) {
        for (int j0 = n - 6; j0 >= 0; j0 -= 6) {  // backward substitution U x = y, y = column n
            double D[36], x[6];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) D[r * 6 + c] = (c >= r) ? M[(j0 + r) * ld + j0 + c] : 0.0;
            double rinv[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                x[r] = M[(j0 + r) * ld + n];
                rinv[r] = s_rd[j0 + r];
            }
#pragma unroll
            for (int r = 5; r >= 0; --r) {
#pragma unroll
                for (int p = 0; p < 6; ++p)
                    if (p > r) x[r] -= D[r * 6 + p] * x[p];
                x[r] *= rinv[r];
            }
            __syncthreads();  // everyone has read y_j
            for (int r = tid; r < j0; r += nt) {
                double v = M[r * ld + n];
#pragma unroll
                for (int p = 0; p < 6; ++p) v -= M[r * ld + j0 + p] * x[p];
                M[r * ld + n] = v;
            }
            if (tid < 6) M[(j0 + tid) * ld + n] = x[tid];
            __syncthreads();
        }
        for (int i = tid; i < n; i += nt) P.x[i] = M[i * ld + n];
    }
    __syncthreads();
    if (tid == 0) P.sc->solve_ok[slot] = fail ? 0 : 1;
}

// large systems (6K > 160): S stays in L2 and is factored in panels of BA_NB = 30 rows (5 pose blocks).  Every CTA loads
// the panel [S | bs](J0 .. J0+30, J0 .. n] into its own shared memory and factors it redundantly (6x6 diagonal blocks
// in registers by every thread, forward substitution of the panel columns and of the right-hand side spread over the
// threads, 3 block barriers per 6 rows); then the whole grid shares the rank-30 trailing update of S and bs.  ONE
// grid-wide barrier per panel: 10 for K = 50 instead of one per block row.  CTA 0 keeps the factor rows in Ubuf and the
// forward-substituted right-hand side in x, and finishes with the blocked backward substitution (diagonal blocks
// staged in shared memory once).
#define BA_NB 30
__device__ void ba_phase_solve_grid(const BaParams& P, int slot, cg::grid_group& grid, double* smem, double* Ubuf,
                                    int gtid, int gsize) {
    const int n = P.n, ld = P.n + 1, tid = threadIdx.x, nt = blockDim.x;
    __shared__ int s_fail;
    __shared__ double s_rinv[BA_NB];  // reciprocal diagonal of the factored nb x nb block
    double* Pn = smem;  // BA_NB x ld, columns J0 .. n valid (column n = right-hand side)
    int fail = 0;
    for (int J0 = 0; J0 < n; J0 += BA_NB) {
        const int nb = min(BA_NB, n - J0), wcols = n - J0;
        if (tid == 0) s_fail = 0;
        // panel load: row by row, unrolled so that a thread has a dozen independent L2 reads in flight
        for (int cc = tid; cc < wcols; cc += nt) {
#pragma unroll 6
            for (int r = 0; r < BA_NB; ++r)
                if (r < nb) Pn[r * ld + J0 + cc] = P.S[(size_t)(J0 + r) * n + J0 + cc];
        }
        if (tid < nb) Pn[tid * ld + n] = P.bs[J0 + tid];
        __syncthreads();
        // (1) the nb x nb diagonal block, 6x6-blocked, inside shared memory (tiny: every step is a few dozen entries)
        const int cend = J0 + nb;  // first column right of the diagonal block
        for (int j = 0; j < nb; j += 6) {
            const int c0 = J0 + j;  // column of this 6x6 block
            double Dr[36], rinv[6], G[36];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) Dr[r * 6 + c] = (c >= r) ? Pn[(j + r) * ld + c0 + c] : 0.0;
            const bool ok = ldl6_diag(Dr, rinv, G);  // every thread, in registers; root-free, branch-free reciprocals
            for (int c = c0 + 6 + tid; c < cend; c += nt) {
                double v[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) v[r] = Pn[(j + r) * ld + c];
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int q = 0; q < 6; ++q)
                        if (q < r) v[r] -= G[q * 6 + r] * v[q];
#pragma unroll
                for (int r = 0; r < 6; ++r) Pn[(j + r) * ld + c] = v[r] * rinv[r];
            }
            __syncthreads();  // block row done; everyone has read the un-factored 6x6 block
            if (tid == 0) {
                if (!ok) s_fail = 1;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
#pragma unroll
                    for (int c = 0; c < 6; ++c)
                        if (c >= r) Pn[(j + r) * ld + c0 + c] = Dr[r * 6 + c] * rinv[r];
                    s_rinv[j + r] = rinv[r];
                }
            }
            const int mr = nb - j - 6;  // rows of the diagonal block still to factor
            for (int e = tid; e < mr * mr; e += nt) {
                const int r2 = j + 6 + e / mr, c = c0 + 6 + e % mr;
                if (c < J0 + r2) continue;
                double v = Pn[r2 * ld + c];
#pragma unroll
                for (int q = 0; q < 6; ++q) v -= Pn[(j + q) * ld + J0 + r2] * Pn[(j + q) * ld + c];
                Pn[r2 * ld + c] = v;
            }
            __syncthreads();
            if (s_fail) break;
        }
        // (2) the panel right of it and the right-hand side: one thread per column runs the whole nb-step forward
        // substitution in registers against the factored block (broadcast reads), no barrier inside.  Each entry first
        // collects the contributions of the finished 6-row groups in four independent accumulators (short dependency
        // chains), then the 6x6 triangle of its own group.
        if (!s_fail) {
            for (int c = cend + tid; c <= n; c += nt) {
                double v[BA_NB];
#pragma unroll
                for (int r = 0; r < BA_NB; ++r) v[r] = r < nb ? Pn[r * ld + c] : 0.0;
#pragma unroll
                for (int g0 = 0; g0 < BA_NB; g0 += 6) {
                    if (g0 < nb) {
#pragma unroll
                        for (int r = g0; r < g0 + 6; ++r) {
                            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                            for (int q = 0; q < g0; ++q) acc[q & 3] += Pn[q * ld + J0 + r] * v[q];
                            v[r] -= (acc[0] + acc[1]) + (acc[2] + acc[3]);
                        }
#pragma unroll
                        for (int r = g0; r < g0 + 6; ++r) {
#pragma unroll
                            for (int q = g0; q < g0 + 6; ++q)
                                if (q < r) v[r] -= Pn[q * ld + J0 + r] * v[q];
                            v[r] *= s_rinv[r];
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < BA_NB; ++r)
                    if (r < nb) Pn[r * ld + c] = v[r];
            }
        }
        __syncthreads();
        if (s_fail) {  // uniform over the grid: every CTA factors the same panel
            fail = 1;
            break;
        }
        if (blockIdx.x == 0) {
            for (int e = tid; e < nb * wcols; e += nt) {
                const int r = e / wcols, cc = e - r * wcols;
                if (cc >= r) Ubuf[(size_t)(J0 + r) * n + J0 + cc] = Pn[r * ld + J0 + cc];
            }
            for (int r = tid; r < nb; r += nt) P.x[J0 + r] = Pn[r * ld + n];  // y = U^-T bs
        }
        const int m = n - J0 - nb;
        for (int e = gtid; e < m * (m + 1); e += gsize) {
            const int r = J0 + nb + e / (m + 1), cc = e % (m + 1);
            const int c = cc < m ? J0 + nb + cc : n;
            if (c < r) continue;
            double v = c < n ? P.S[(size_t)r * n + c] : P.bs[r];
#pragma unroll 6
            for (int q = 0; q < nb; ++q) v -= Pn[q * ld + r] * Pn[q * ld + c];
            if (c < n) P.S[(size_t)r * n + c] = v;
            else P.bs[r] = v;
        }
        grid.sync();
    }
    if (!fail && blockIdx.x == 0) {
        // backward substitution U x = y by CTA 0: y and the 6x6 diagonal blocks live in shared memory
        double* y = smem;
        double* Dg = smem + n;  // [n / 6][36]
        for (int i = tid; i < n; i += nt) y[i] = P.x[i];
        for (int e = tid; e < n * 6; e += nt) {
            const int blk = e / 36, q = e - blk * 36;
            Dg[e] = Ubuf[(size_t)(blk * 6 + q / 6) * n + blk * 6 + q % 6];
        }
        __syncthreads();
        // each thread owns rows tid and tid + nt (n <= 2 nt) and fetches their 6 factor entries one block step ahead
        double un[2][6];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int r = tid + k * nt;
#pragma unroll
            for (int q = 0; q < 6; ++q) un[k][q] = r < n - 6 ? Ubuf[(size_t)r * n + n - 6 + q] : 0.0;
        }
        for (int j0 = n - 6; j0 >= 0; j0 -= 6) {
            const double* D = Dg + (j0 / 6) * 36;
            double x[6], rinv[6], uc[2][6];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int r = tid + k * nt;
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    uc[k][q] = un[k][q];
                    un[k][q] = (j0 >= 6 && r < j0 - 6) ? Ubuf[(size_t)r * n + j0 - 6 + q] : 0.0;
                }
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                x[r] = y[j0 + r];
                rinv[r] = 1.0 / D[r * 6 + r];  // six independent divisions, pipelined
            }
#pragma unroll
            for (int r = 5; r >= 0; --r) {
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    if (q > r) x[r] -= D[r * 6 + q] * x[q];
                x[r] *= rinv[r];
            }
            __syncthreads();  // everyone has read y_j
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int r = tid + k * nt;
                if (r < j0) {
                    double v = y[r];
#pragma unroll
                    for (int q = 0; q < 6; ++q) v -= uc[k][q] * x[q];
                    y[r] = v;
                }
            }
            if (tid < 6) y[j0 + tid] = x[tid];
            __syncthreads();
        }
        for (int i = tid; i < n; i += nt) P.x[i] = y[i];
    }
    if (gtid == 0) P.sc->solve_ok[slot] = fail ? 0 : 1;
}

// UPDATE (K16): landmark back-substitution, trial estimate = oplus(current, x), computeScale partial sums
__device__ void ba_phase_update(const BaParams& P, int cur, double lambda, int slot, int gtid, int gsize) {
    const int n = P.n;
    if (!P.sc->solve_ok[slot]) return;
    const double* poses = P.poses + (size_t)cur * P.K * 12;
    double* tposes = P.poses + (size_t)(1 - cur) * P.K * 12;
    const double* points = P.points + (size_t)cur * P.L * 3;
    double* tpoints = P.points + (size_t)(1 - cur) * P.L * 3;
    double scale = 0.0;
    if (gtid < P.K && P.shard_L0 == 0) {  // poses are replicated; rank 0's shard accounts for their scale term
        const int k = gtid;
        const double* xi = P.x + 6 * k;
#pragma unroll
        for (int a = 0; a < 6; ++a) scale += xi[a] * (lambda * xi[a] + P.bp_scale[6 * k + a]);
    }
    if (gtid < P.K) {
        const int k = gtid;
        double dR[9], dt[3];
        se3_exp_dev(P.x + 6 * k, dR, dt);
        const double* T = poses + 12 * k;
        double* Tn = tposes + 12 * k;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) Tn[r * 4 + c] = dR[r * 3] * T[c] + dR[r * 3 + 1] * T[4 + c] + dR[r * 3 + 2] * T[8 + c];
            Tn[r * 4 + 3] = dR[r * 3] * T[3] + dR[r * 3 + 1] * T[7] + dR[r * 3 + 2] * T[11] + dt[r];
        }
    }
    if (!P.pose_only) {
        for (int l = P.shard_L0 + gtid; l < P.shard_L1; l += gsize) {
            const int o0 = P.lm_start[l], o1 = P.lm_start[l + 1];
            double c0 = P.bl[3 * l], c1 = P.bl[3 * l + 1], c2 = P.bl[3 * l + 2];
            for (int i = o0; i < o1; ++i) {
                const double* Bi = P.Hpl + 18 * (size_t)i;
                const double* xp = P.x + 6 * P.obs_pose[i];
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    c0 -= Bi[a * 3] * xp[a];
                    c1 -= Bi[a * 3 + 1] * xp[a];
                    c2 -= Bi[a * 3 + 2] * xp[a];
                }
            }
            const double* d = P.Dinv + 9 * (size_t)l;
            const double x0 = d[0] * c0 + d[1] * c1 + d[2] * c2, x1 = d[3] * c0 + d[4] * c1 + d[5] * c2,
                         x2 = d[6] * c0 + d[7] * c1 + d[8] * c2;
            P.x[n + 3 * l] = x0; P.x[n + 3 * l + 1] = x1; P.x[n + 3 * l + 2] = x2;
            tpoints[3 * l] = points[3 * l] + x0;
            tpoints[3 * l + 1] = points[3 * l + 1] + x1;
            tpoints[3 * l + 2] = points[3 * l + 2] + x2;
            scale += x0 * (lambda * x0 + P.bl[3 * l]) + x1 * (lambda * x1 + P.bl[3 * l + 1]) + x2 * (lambda * x2 + P.bl[3 * l + 2]);
        }
    }
    scale = warp_sum(scale);
    if ((threadIdx.x & 31) == 0 && scale != 0.0) atomicAdd(&P.sc->scale[slot], scale);
}

// TRIAL_ERR: computeActiveErrors + activeRobustChi2 at the trial estimate (the edges' _error is overwritten, as in g2o)
__device__ void ba_phase_trial_err(const BaParams& P, int cur, int slot, int gtid, int gsize) {
    if (!P.sc->solve_ok[slot]) return;
    const double* tposes = P.poses + (size_t)(1 - cur) * P.K * 12;
    const double* tpoints = P.pose_only ? P.points + (size_t)cur * P.L * 3 : P.points + (size_t)(1 - cur) * P.L * 3;
    double chi = 0.0;
    const int i0 = P.lm_start[P.shard_L0], i1 = P.lm_start[P.shard_L1];
    for (int i = i0 + gtid; i < i1; i += gsize) {
        double pc[3], e0, e1, r0, w;
        residual_dev(tposes + 12 * P.obs_pose[i], tpoints + 3 * P.obs_point[i], P.Kc, P.obs_uv[2 * i], P.obs_uv[2 * i + 1],
                     e0, e1, pc);
        P.err[2 * i] = e0;
        P.err[2 * i + 1] = e1;
        huber_dev(e0 * e0 + e1 * e1, P.delta, r0, w);
        chi += r0;
    }
    chi = warp_sum(chi);
    if ((threadIdx.x & 31) == 0 && chi != 0.0) atomicAdd(&P.sc->chi_trial[slot], chi);
}

// RELABEL counts: #edges with chi2 <= th * 2^r for r = 0..5 in one pass (optimization.cpp:224-252)
__device__ void ba_phase_relabel_count(const BaParams& P, int gtid, int gsize) {
    int c[6] = {0, 0, 0, 0, 0, 0};
    const int i0 = P.lm_start[P.shard_L0], i1 = P.lm_start[P.shard_L1];
    for (int i = i0 + gtid; i < i1; i += gsize) {
        const double chi = P.err[2 * i] * P.err[2 * i] + P.err[2 * i + 1] * P.err[2 * i + 1];
        double th = P.chi2_th;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            c[r] += !(chi > th);
            th *= 2;
        }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        c[r] = __reduce_add_sync(0xFFFFFFFFu, c[r]);
        if ((threadIdx.x & 31) == 0 && c[r]) atomicAdd(&P.sc->cnt_le[r], c[r]);
    }
}

__device__ double ba_relabel_threshold(const BaParams& P, int total_obs, int* n_in) {
    double th = P.chi2_th;
    int r = 0;
    for (; r < 5; ++r) {
        const double ratio = P.sc->cnt_le[r] / (double)total_obs;
        if (ratio > 0.5) break;
        th *= 2;
    }
    *n_in = P.sc->cnt_le[r];
    return th;
}

// RELABEL apply: per-edge chi2 (insertion order) and Landmark::is_inlier = verdict of the landmark's LAST edge
__device__ void ba_phase_relabel_apply(const BaParams& P, double th, int gtid, int gsize) {
    for (int l = P.shard_L0 + gtid; l < P.shard_L1; l += gsize) {
        const int o0 = P.lm_start[l], o1 = P.lm_start[l + 1];
        for (int i = o0; i < o1; ++i) {
            const double chi = P.err[2 * i] * P.err[2 * i] + P.err[2 * i + 1] * P.err[2 * i + 1];
            if (P.chi2_out) P.chi2_out[P.obs_orig[i]] = chi;
            if (i == o1 - 1 && P.inlier_out) P.inlier_out[l] = chi > th ? 0 : 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the persistent single-GPU kernel: optimizer.optimize(num_ite) + relabel in one launch
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BA_THREADS)
ba_lm_kernel(const __grid_constant__ BaParams P) {
    extern __shared__ double smem[];
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    BaScalars* sc = P.sc;

    // every thread carries an identical copy of the LM state (all inputs are read after a grid.sync)
    int cur = 0, trials = 0, accepted = 0, it = 0;
    double lambda = 0.0, ni = 2.0, chi_first = 0.0, chi_last = 0.0;

    unsigned long long t_last = 0;
    if (gtid == 0) {
#pragma unroll
        for (int r = 0; r < 6; ++r) sc->cnt_le[r] = 0;
        for (int r = 0; r < 12; ++r) sc->phase_ns[r] = 0;
        t_last = gtimer();
    }
    for (it = 0; it < P.num_iterations; ++it) {
        ba_phase_zero(P, gtid, gsize);
        grid.sync();
        BA_TICK(0);
        ba_phase_build_edges(P, cur, gtid, gsize);
        ba_phase_build_poses(P, cur, smem);
        if (it == 0) {
            grid.sync();  // Hll complete
            ba_phase_maxdiag(P, gtid, gsize);
        }
        grid.sync();
        BA_TICK(1);
        double currentChi = sc->chi_cur;
        if (it == 0) {
            chi_first = currentChi;
            __shared__ double s_md;
            if (threadIdx.x == 0) {
                double md = __longlong_as_double((long long)sc->maxdiag_bits);
                for (int i = 0; i < P.n; ++i) md = fmax(md, fabs(P.Hpp[(i / 6) * 36 + (i % 6) * 7]));
                s_md = md;
            }
            __syncthreads();
            lambda = P.tau * s_md;  // computeLambdaInit: tau * max |diagonal| over poses and landmarks
            ni = 2.0;
        }
        double rho = 0.0;
        int qmax = 0;
        do {
            const int slot = trials & 1;
            ba_phase_dinv(P, lambda, slot, 1, gtid, gsize);
            grid.sync();
            BA_TICK(2);
            if (P.has_dup) ba_phase_schur_atomic(P, gtid, gsize);
            else ba_phase_schur_blocks(P, smem);
            grid.sync();
            BA_TICK(3);
            BA_TICK(4);
            if (P.n <= BA_SMEM_CHOL_MAX) ba_phase_solve_cta(P, slot, smem);
            else ba_phase_solve_grid(P, slot, grid, smem, P.Ubuf, gtid, gsize);
            grid.sync();
            BA_TICK(5);
            ba_phase_update(P, cur, lambda, slot, gtid, gsize);
            grid.sync();
            BA_TICK(6);
            ba_phase_trial_err(P, cur, slot, gtid, gsize);
            grid.sync();
            BA_TICK(7);
            const int ok2 = sc->solve_ok[slot];
            const double tempChi = ok2 ? sc->chi_trial[slot] : DBL_MAX;
            const double scale = (ok2 ? sc->scale[slot] : 0.0) + 1e-3;
            rho = (currentChi - tempChi) / scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
                cur = 1 - cur;  // discardTop(): the trial estimate becomes the estimate
                if (P.pose_only) {
                    // points are not touched in pose-only mode: both buffers hold the same values
                }
                accepted++;
            } else {
                lambda *= ni;
                ni *= 2;  // pop(): keep `cur`; the edges' errors stay at the trial values, as in g2o
            }
            qmax++;
            trials++;
        } while (rho < 0 && qmax < P.max_trials);
        chi_last = currentChi;
        if (qmax == P.max_trials || rho == 0) {
            it++;
            break;
        }
    }
    if (P.num_iterations <= 0) {
        ba_phase_zero(P, gtid, gsize);
        grid.sync();
        ba_phase_build_edges(P, cur, gtid, gsize);
        grid.sync();
        chi_first = chi_last = sc->chi_cur;
    }
    // relabel
    grid.sync();
    ba_phase_relabel_count(P, gtid, gsize);
    grid.sync();
    int n_in = 0;
    const double th = ba_relabel_threshold(P, P.n_obs, &n_in);
    ba_phase_relabel_apply(P, th, gtid, gsize);
    // the accepted estimate must end up in buffer 0
    if (cur == 1) {
        for (int i = gtid; i < P.K * 12; i += gsize) P.poses[i] = P.poses[P.K * 12 + i];
        if (!P.pose_only)
            for (int i = gtid; i < P.L * 3; i += gsize) P.points[i] = P.points[(size_t)P.L * 3 + i];
    }
    if (gtid == 0) {
        sc->iterations = it;
        sc->trials = trials;
        sc->accepted = accepted;
        sc->chi2_initial = chi_first;
        sc->chi2_final = chi_last;
        sc->lambda_final = lambda;
        sc->chi2_threshold = th;
        sc->n_inlier_obs = n_in;
        sc->n_outlier_obs = P.n_obs - n_in;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Small windows (6K <= BA_SMALL_N, i.e. K <= 16: the reference's own window is 10 keyframes): the same LM loop with TWO
// grid barriers per trial and one per outer iteration instead of five and three.
//   * BUILD writes into one of two Hpp/bp/chi accumulators and clears the other (no separate ZERO phase), and takes the
//     max |diag Hll| for lambda_0 in passing.
//   * SCHUR recomputes Dinv = (Hll + lambda I)^-1 per use (45 flops) instead of a DINV phase, and accumulates
//     -Hpl Dinv Hpl^T into one of two zero-initialised buffers; Hpp + lambda I is added when the system is loaded.
//   * SOLVE + UPDATE + TRIAL_ERR are one phase: EVERY CTA factors the (identical) 6K x 6K system redundantly in its own
//     shared memory -- so the update vector needs no broadcast -- computes all K trial poses into shared memory, then
//     back-substitutes, updates and re-evaluates the edges of the landmarks it owns.
// Same arithmetic per edge / per block as the general kernel; only summation order differs.
// ---------------------------------------------------------------------------------------------------------------
#define BA_SMALL_N 96

__device__ __forceinline__ void inv3_sym_dev(const double* H, double lambda, double* d) {
    const double m0 = H[0] + lambda, m1 = H[1], m2 = H[2], m4 = H[4] + lambda, m5 = H[5], m8 = H[8] + lambda;
    const double c0 = m4 * m8 - m5 * m5, c1 = m5 * m2 - m1 * m8, c2 = m1 * m5 - m4 * m2;
    const double id = 1.0 / (m0 * c0 + m1 * c1 + m2 * c2);
    d[0] = c0 * id; d[1] = c1 * id; d[2] = c2 * id;
    d[3] = d[1];    d[4] = (m0 * m8 - m2 * m2) * id; d[5] = (m2 * m1 - m0 * m5) * id;
    d[6] = d[2];    d[7] = d[5];                     d[8] = (m0 * m4 - m1 * m1) * id;
}

// SCHUR blocks with Dinv recomputed per use; accumulates into S / bs (zero-initialised): S_ij -= sum Hpl_i Dinv Hpl_j^T
__device__ void ba_small_schur(const BaParams& P, double lambda, double* __restrict__ S, double* __restrict__ bs, double* smem) {
    const int n = P.n, K = P.K;
    const int nb = K * (K + 1) / 2;
    const int parts = max(1, (int)gridDim.x / nb);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int wi = blockIdx.x; wi < nb * parts; wi += gridDim.x) {
        const int b = wi % nb, part = wi / nb;
        if (!P.block_flag[b]) continue;
        int ki = 0, rem = b;
        while (rem >= K - ki) { rem -= K - ki; ++ki; }
        const int kj = ki + rem;
        double acc[42];
#pragma unroll
        for (int q = 0; q < 42; ++q) acc[q] = 0.0;
        const int ps = P.pose_start[ki], pe = P.pose_start[ki + 1];
        for (int t = ps + part * BA_THREADS + tid; t < pe; t += parts * BA_THREADS) {
            const int i = P.pose_obs[t];
            const int l = P.obs_point[i];
            const int j = (ki == kj) ? i : P.obs_of[(size_t)kj * P.L + l];
            if (j < 0) continue;
            const double* Bi = P.Hpl + 18 * (size_t)i;
            double d[9];
            inv3_sym_dev(P.Hll + 9 * (size_t)l, lambda, d);
            double BD[18];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double x0 = Bi[a * 3], x1 = Bi[a * 3 + 1], x2 = Bi[a * 3 + 2];
                BD[a * 3] = x0 * d[0] + x1 * d[3] + x2 * d[6];
                BD[a * 3 + 1] = x0 * d[1] + x1 * d[4] + x2 * d[7];
                BD[a * 3 + 2] = x0 * d[2] + x1 * d[5] + x2 * d[8];
            }
            const double* Bj = P.Hpl + 18 * (size_t)j;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    acc[a * 6 + c] += BD[a * 3] * Bj[c * 3] + BD[a * 3 + 1] * Bj[c * 3 + 1] + BD[a * 3 + 2] * Bj[c * 3 + 2];
            if (ki == kj) {
                const double b0 = P.bl[3 * l], b1 = P.bl[3 * l + 1], b2 = P.bl[3 * l + 2];
#pragma unroll
                for (int a = 0; a < 6; ++a) acc[36 + a] += BD[a * 3] * b0 + BD[a * 3 + 1] * b1 + BD[a * 3 + 2] * b2;
            }
        }
        __syncthreads();
        const int nacc = (ki == kj) ? 42 : 36;
#pragma unroll
        for (int q = 0; q < 42; ++q) {
            if (q < nacc) {
                const double v = warp_sum(acc[q]);
                if (lane == 0) smem[warp * 42 + q] = v;
            }
        }
        __syncthreads();
        if (tid < nacc) {
            double v = 0.0;
            for (int w8 = 0; w8 < BA_THREADS / 32; ++w8) v += smem[w8 * 42 + tid];
            if (v != 0.0) {
                if (tid < 36) atomicAdd(&S[(size_t)(6 * ki + tid / 6) * n + 6 * kj + tid % 6], -v);
                else atomicAdd(&bs[6 * ki + tid - 36], -v);
            }
        }
    }
}

// Every CTA: solve [S + Hpp + lambda I] x = bs + bp by a 6x6-blocked Cholesky factorisation; x -> xs (shared).  Returns
// false (uniformly over the grid: identical inputs) if the system is not positive definite.
//
// Measured with tools/micro/solve_probe2.cu (n = 60: 42.3 k -> 31.7 k cycles stand-alone, profiles/r02_v3_fp64_latency.txt):
//   * the rank-6 trailing update is SHARED-MEMORY-BANDWIDTH bound when the trailing matrix lives in shared memory (every
//     element read and written once per block step), so it lives in REGISTERS instead, 2-D cyclic (warp w owns the rows
//     r = w mod 8, lane l the columns c = l mod 32, SB_A x SB_B entries per thread, loaded straight from global memory);
//     shared memory only ever holds finished factor rows.  Entries below the diagonal are updated with whatever they
//     hold and never published: the update carries no per-entry predicates;
//   * per block step: the six pivot rows are published (each lives in one warp), every thread factors the 6x6 diagonal
//     block redundantly in registers (root-free, branch-free reciprocals: the dependency chain of the step), one thread
//     per column forward-substitutes the panel and the right-hand side (column n), then the trailing rows are updated
//     from the panel: two block barriers per block step;
//   * the backward substitution reads the factor rows from shared memory, blocked the same way.
template <int SB_A, int SB_B>
__device__ bool ba_small_solve_t(const BaParams& P, const double* S, const double* bs, const double* Hpp, const double* bp,
                                 double lambda, double* M, double* xs) {
    const int n = P.n, ld = P.n + 1, tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    static_assert(BA_THREADS == 256, "8 warps: row residues mod 8");
    __shared__ int s_fail;
    __shared__ double s_rd[BA_SMALL_N];
    if (tid == 0) s_fail = 0;
    double v[SB_A][SB_B];
#pragma unroll
    for (int a = 0; a < SB_A; ++a) {
        const int r = warp + 8 * a;
        const int hb = (r / 6) * 36 + (r % 6) * 6 - (r / 6) * 6;  // Hpp offset of (r, c) inside r's diagonal block: hb + c
#pragma unroll
        for (int b = 0; b < SB_B; ++b) {
            const int c = lane + 32 * b;
            double x = 0.0;
            if (r < n && c >= r && c < n) {
                x = __ldcg(S + r * n + c);
                if (c < (r / 6) * 6 + 6) x += __ldcg(Hpp + hb + c);
                if (c == r) x += lambda;
            }
            if (r < n && c == n) x = __ldcg(bs + r) + __ldcg(bp + r);
            v[a][b] = x;
        }
    }
    __syncthreads();  // s_fail cleared; M free (the caller's previous phase used it as scratch)
    for (int j0 = 0; j0 < n; j0 += 6) {
        // (1) publish the six pivot rows
#pragma unroll
        for (int a = 0; a < SB_A; ++a) {
            const int r = warp + 8 * a;
            if (r >= j0 && r < j0 + 6) {  // warp-uniform
#pragma unroll
                for (int b = 0; b < SB_B; ++b) {
                    const int c = lane + 32 * b;
                    if (c >= r && c <= n) M[r * ld + c] = v[a][b];
                }
            }
        }
        __syncthreads();
        // (2) diagonal block by every thread, panel columns (and the right-hand side) by one thread each
        {
            double D[36], rs[6], G[36];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) D[r * 6 + c] = (c >= r) ? M[(j0 + r) * ld + j0 + c] : 0.0;
            const bool ok = ldl6_diag(D, rs, G);
            for (int c = j0 + 6 + tid; c <= n; c += nt) {
                double w[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) w[r] = M[(j0 + r) * ld + c];
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int p = 0; p < 6; ++p)
                        if (p < r) w[r] -= G[p * 6 + r] * w[p];
#pragma unroll
                for (int r = 0; r < 6; ++r) M[(j0 + r) * ld + c] = w[r] * rs[r];
            }
            if (tid == 0 && !ok) s_fail = 1;
            __syncthreads();  // panel done; everyone has read the un-factored diagonal block
            if (tid == nt - 1) {  // the factored diagonal block and its reciprocals: only the backward substitution reads them
#pragma unroll
                for (int r = 0; r < 6; ++r) {
#pragma unroll
                    for (int c = 0; c < 6; ++c)
                        if (c >= r) M[(j0 + r) * ld + j0 + c] = D[r * 6 + c] * rs[r];
                    s_rd[j0 + r] = rs[r];
                }
            }
        }
        if (s_fail) break;
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
                if (a >= a1 && r < n) {  // warp-uniform
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
    }
    __syncthreads();
    const bool fail = s_fail != 0;
    if (!fail) {
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
                double y = M[r * ld + n];
#pragma unroll
                for (int p = 0; p < 6; ++p) y -= M[r * ld + j0 + p] * x[p];
                M[r * ld + n] = y;
            }
            if (tid == 0) {
#pragma unroll
                for (int r = 0; r < 6; ++r) M[(j0 + r) * ld + n] = x[r];
            }
            __syncthreads();
        }
        for (int i = tid; i < n; i += nt) xs[i] = M[i * ld + n];
    }
    __syncthreads();
    return !fail;
}

__device__ bool ba_small_solve(const BaParams& P, const double* S, const double* bs, const double* Hpp, const double* bp,
                               double lambda, double* M, double* xs) {
    const int n = P.n, tid = threadIdx.x;
    if (P.pose_only) {
        __shared__ int s_fail;
        if (tid == 0) s_fail = 0;

        // No landmark vertices: the system is block diagonal (K independent 6x6 blocks Hpp_k + lambda I, right-hand
        // side bp_k; S and bs stay zero), one thread per pose.  The same arithmetic as the dense factorisation applied
        // to a block-diagonal matrix -- the off-diagonal zeros only ever contribute exact zeros -- without its 2 K
        // block barriers (24 us -> 1 us per trial of optimize_pose_only).
        __syncthreads();
        if (tid < P.K) {
            double D[36], rinv[6], x[6];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    D[r * 6 + c] = (c >= r) ? __ldcg(Hpp + tid * 36 + r * 6 + c) + __ldcg(S + (6 * tid + r) * n + 6 * tid + c) +
                                                  (r == c ? lambda : 0.0)
                                            : 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) x[r] = __ldcg(bs + 6 * tid + r) + __ldcg(bp + 6 * tid + r);
            if (!chol6_diag(D, 6, 0, rinv)) s_fail = 1;
#pragma unroll
            for (int r = 0; r < 6; ++r) {  // U^T y = b
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    if (q < r) x[r] -= D[q * 6 + r] * x[q];
                x[r] *= rinv[r];
            }
#pragma unroll
            for (int r = 5; r >= 0; --r) {  // U x = y
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    if (q > r) x[r] -= D[r * 6 + q] * x[q];
                x[r] *= rinv[r];
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) xs[6 * tid + r] = x[r];
        }
        __syncthreads();
        return s_fail == 0;
    }
    if (n <= 63) return ba_small_solve_t<8, 2>(P, S, bs, Hpp, bp, lambda, M, xs);
    return ba_small_solve_t<BA_SMALL_N / 8, (BA_SMALL_N + 32) / 32>(P, S, bs, Hpp, bp, lambda, M, xs);
}

__global__ void __launch_bounds__(BA_THREADS)
ba_lm_small_kernel(const __grid_constant__ BaParams P) {
    extern __shared__ double smem[];
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    const int tid = threadIdx.x;
    BaScalars* sc = P.sc;
    const int n = P.n, K = P.K;
    // shared memory: [0, n(n+1)) system / block-reduction scratch, then x [n], trial poses [12 K]
    double* xs = smem + n * (n + 1);
    double* tp = xs + n;
    // double-buffered accumulators carved out of the (much larger) general-kernel buffers
    double* HppB[2] = {P.Hpp, P.Hpp + 36 * K};
    double* bpB[2] = {P.bp, P.bp + n};
    double* SB[2] = {P.S, P.S + (size_t)n * n};
    double* bsB[2] = {P.bs, P.bs + n};

    int cur = 0, trials = 0, accepted = 0, it = 0;
    double lambda = 0.0, ni = 2.0, chi_first = 0.0, chi_last = 0.0;
    unsigned long long t_last = 0;
    // prologue: clear accumulator set 0 and Schur buffer 0
    for (int i = gtid; i < 36 * K; i += gsize) HppB[0][i] = 0.0;
    for (int i = gtid; i < n; i += gsize) { bpB[0][i] = 0.0; bsB[0][i] = 0.0; }
    for (int i = gtid; i < n * n; i += gsize) SB[0][i] = 0.0;
    if (!P.pose_only) {
        for (int i = P.shard_L0 * 9 + gtid; i < P.shard_L1 * 9; i += gsize) P.Hll[i] = 0.0;
        for (int i = P.shard_L0 * 3 + gtid; i < P.shard_L1 * 3; i += gsize) P.bl[i] = 0.0;
    }
    if (gtid == 0) {
        for (int r = 0; r < 6; ++r) sc->cnt_le[r] = 0;
        for (int r = 0; r < 12; ++r) sc->phase_ns[r] = 0;
        sc->chi_cur = 0.0;
        sc->maxdiag_bits = 0ull;
        sc->chi_trial[0] = sc->chi_trial[1] = 0.0;
        sc->scale[0] = sc->scale[1] = 0.0;
        t_last = gtimer();
    }
    grid.sync();
    BA_TICK(0);
    for (it = 0; it < P.num_iterations; ++it) {
        const int hb = it & 1;
        {   // BUILD into accumulator set hb; clear the other set for the next outer iteration
            BaParams Q = P;
            Q.Hpp = HppB[hb];
            Q.bp = bpB[hb];
            ba_phase_build_edges(Q, cur, gtid, gsize);
            ba_phase_build_poses(Q, cur, smem);
            for (int i = gtid; i < 36 * K; i += gsize) HppB[1 - hb][i] = 0.0;
            for (int i = gtid; i < n; i += gsize) bpB[1 - hb][i] = 0.0;
        }
        if (it == 0) {
            grid.sync();  // Hll complete (lambda_0 needs max |diag|); later iterations skip this
            ba_phase_maxdiag(P, gtid, gsize);
        }
        grid.sync();
        BA_TICK(1);
        double currentChi = sc->chi_cur;
        if (it == 0) {
            chi_first = currentChi;
            __shared__ double s_md;
            if (tid == 0) {
                double md = __longlong_as_double((long long)sc->maxdiag_bits);
                for (int i = 0; i < n; ++i) md = fmax(md, fabs(__ldcg(HppB[hb] + (i / 6) * 36 + (i % 6) * 7)));
                s_md = md;
            }
            __syncthreads();
            lambda = P.tau * s_md;
            ni = 2.0;
        }
        double rho = 0.0;
        int qmax = 0;
        do {
            const int tb = trials & 1;
            if (!P.pose_only) {
                ba_small_schur(P, lambda, SB[tb], bsB[tb], smem);
                grid.sync();
            }
            BA_TICK(3);
            // ---- every CTA: solve, then update + trial errors of its own landmarks --------------------------------
            const bool ok = ba_small_solve(P, SB[tb], bsB[tb], HppB[hb], bpB[hb], lambda, smem, xs);
            BA_TICK(5);
            // clear the other Schur buffer and the other trial slot for the next trial; chi_cur for the next BUILD
            for (int i = gtid; i < n * n; i += gsize) SB[1 - tb][i] = 0.0;
            for (int i = gtid; i < n; i += gsize) bsB[1 - tb][i] = 0.0;
            if (gtid == 0) {
                sc->chi_trial[1 - tb] = 0.0;
                sc->scale[1 - tb] = 0.0;
                sc->chi_cur = 0.0;
            }
            double chi = 0.0, scale = 0.0;
            if (ok) {
                const double* poses = P.poses + (size_t)cur * K * 12;
                double* tposes = P.poses + (size_t)(1 - cur) * K * 12;
                if (tid < K) {  // trial poses, redundantly per CTA (CTA 0 publishes them)
                    double dR[9], dt[3];
                    se3_exp_dev(xs + 6 * tid, dR, dt);
                    const double* T = poses + 12 * tid;
                    double* Tn = tp + 12 * tid;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) Tn[r * 4 + c] = dR[r * 3] * T[c] + dR[r * 3 + 1] * T[4 + c] + dR[r * 3 + 2] * T[8 + c];
                        Tn[r * 4 + 3] = dR[r * 3] * T[3] + dR[r * 3 + 1] * T[7] + dR[r * 3 + 2] * T[11] + dt[r];
                    }
                    if (blockIdx.x == 0) {
#pragma unroll
                        for (int q = 0; q < 12; ++q) tposes[12 * tid + q] = Tn[q];
                    }
                }
                if (gtid < K) {  // computeScale, pose part
                    const double* xi = xs + 6 * gtid;
#pragma unroll
                    for (int a = 0; a < 6; ++a) scale += xi[a] * (lambda * xi[a] + __ldcg(bpB[hb] + 6 * gtid + a));
                }
                __syncthreads();
                // landmarks: eight lanes per landmark, as in BUILD
                const double* points = P.points + (size_t)cur * P.L * 3;
                double* tpoints = P.points + (size_t)(1 - cur) * P.L * 3;
                const int sub = gtid & (BA_LM_GROUP - 1), ngroups = gsize / BA_LM_GROUP;
                const int nl = P.pose_only ? 0 : P.shard_L1 - P.shard_L0;
                const int iters = (nl + ngroups - 1) / ngroups;
                for (int itl = 0; itl < iters; ++itl) {
                    const int l = P.shard_L0 + itl * ngroups + gtid / BA_LM_GROUP;
                    const bool live = l < P.shard_L1;
                    const int o0 = live ? P.lm_start[l] : 0, o1 = live ? P.lm_start[l + 1] : 0;
                    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
                    for (int i = o0 + sub; i < o1; i += BA_LM_GROUP) {
                        const double* Bi = P.Hpl + 18 * (size_t)i;
                        const double* xp = xs + 6 * P.obs_pose[i];
#pragma unroll
                        for (int a = 0; a < 6; ++a) {
                            c0 -= Bi[a * 3] * xp[a];
                            c1 -= Bi[a * 3 + 1] * xp[a];
                            c2 -= Bi[a * 3 + 2] * xp[a];
                        }
                    }
#pragma unroll
                    for (int o = BA_LM_GROUP / 2; o > 0; o >>= 1) {
                        c0 += __shfl_xor_sync(0xFFFFFFFFu, c0, o);
                        c1 += __shfl_xor_sync(0xFFFFFFFFu, c1, o);
                        c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, o);
                    }
                    double pt[3] = {0.0, 0.0, 0.0};
                    if (live && o1 > o0) {
                        const double b0 = P.bl[3 * l], b1 = P.bl[3 * l + 1], b2 = P.bl[3 * l + 2];
                        c0 += b0; c1 += b1; c2 += b2;
                        double d[9];
                        inv3_sym_dev(P.Hll + 9 * (size_t)l, lambda, d);
                        const double x0 = d[0] * c0 + d[1] * c1 + d[2] * c2, x1 = d[3] * c0 + d[4] * c1 + d[5] * c2,
                                     x2 = d[6] * c0 + d[7] * c1 + d[8] * c2;
                        pt[0] = points[3 * l] + x0; pt[1] = points[3 * l + 1] + x1; pt[2] = points[3 * l + 2] + x2;
                        if (sub == 0) {
                            P.x[n + 3 * l] = x0; P.x[n + 3 * l + 1] = x1; P.x[n + 3 * l + 2] = x2;
                            tpoints[3 * l] = pt[0]; tpoints[3 * l + 1] = pt[1]; tpoints[3 * l + 2] = pt[2];
                            scale += x0 * (lambda * x0 + b0) + x1 * (lambda * x1 + b1) + x2 * (lambda * x2 + b2);
                        }
                    } else if (live && sub == 0) {  // landmark without edges: carried over unchanged
                        tpoints[3 * l] = points[3 * l]; tpoints[3 * l + 1] = points[3 * l + 1]; tpoints[3 * l + 2] = points[3 * l + 2];
                    }
                    for (int i = o0 + sub; i < o1; i += BA_LM_GROUP) {  // computeActiveErrors at the trial estimate
                        double pc[3], e0, e1, r0, w;
                        residual_dev(tp + 12 * P.obs_pose[i], pt, P.Kc, P.obs_uv[2 * i], P.obs_uv[2 * i + 1], e0, e1, pc);
                        P.err[2 * i] = e0;
                        P.err[2 * i + 1] = e1;
                        huber_dev(e0 * e0 + e1 * e1, P.delta, r0, w);
                        chi += r0;
                    }
                }
                if (P.pose_only) {  // points are fixed: every edge against the trial poses
                    for (int i = gtid; i < P.n_obs; i += gsize) {
                        double pc[3], e0, e1, r0, w;
                        residual_dev(tp + 12 * P.obs_pose[i], points + 3 * P.obs_point[i], P.Kc, P.obs_uv[2 * i], P.obs_uv[2 * i + 1],
                                     e0, e1, pc);
                        P.err[2 * i] = e0;
                        P.err[2 * i + 1] = e1;
                        huber_dev(e0 * e0 + e1 * e1, P.delta, r0, w);
                        chi += r0;
                    }
                }
                chi = warp_sum(chi);
                scale = warp_sum(scale);
                if ((tid & 31) == 0) {
                    if (chi != 0.0) atomicAdd(&sc->chi_trial[tb], chi);
                    if (scale != 0.0) atomicAdd(&sc->scale[tb], scale);
                }
            }
            grid.sync();
            BA_TICK(6);
            const double tempChi = ok ? sc->chi_trial[tb] : DBL_MAX;
            const double scl = (ok ? sc->scale[tb] : 0.0) + 1e-3;
            rho = (currentChi - tempChi) / scl;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
                cur = 1 - cur;
                accepted++;
            } else {
                lambda *= ni;
                ni *= 2;
            }
            qmax++;
            trials++;
        } while (rho < 0 && qmax < P.max_trials);
        chi_last = currentChi;
        if (qmax == P.max_trials || rho == 0) {
            it++;
            break;
        }
    }
    if (P.num_iterations <= 0) {
        ba_phase_build_edges(P, cur, gtid, gsize);
        grid.sync();
        chi_first = chi_last = sc->chi_cur;
    }
    grid.sync();
    ba_phase_relabel_count(P, gtid, gsize);
    grid.sync();
    int n_in = 0;
    const double th = ba_relabel_threshold(P, P.n_obs, &n_in);
    ba_phase_relabel_apply(P, th, gtid, gsize);
    if (cur == 1) {
        for (int i = gtid; i < K * 12; i += gsize) P.poses[i] = P.poses[K * 12 + i];
        if (!P.pose_only)
            for (int i = gtid; i < P.L * 3; i += gsize) P.points[i] = P.points[(size_t)P.L * 3 + i];
    }
    if (gtid == 0) {
        sc->iterations = it;
        sc->trials = trials;
        sc->accepted = accepted;
        sc->chi2_initial = chi_first;
        sc->chi2_final = chi_last;
        sc->lambda_final = lambda;
        sc->chi2_threshold = th;
        sc->n_inlier_obs = n_in;
        sc->n_outlier_obs = P.n_obs - n_in;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct BaState {
    int maxK, maxL, maxObs, n_cta, smem_bytes;
    double *d_poses, *d_points, *d_uv, *d_err, *d_Hpl, *d_Hll, *d_bl, *d_Dinv, *d_Hpp, *d_bp, *d_S, *d_bs, *d_x, *d_dbl;
    int *d_pose_start, *d_pose_obs, *d_obs_of, *d_block_flag;
    double* d_U;
    double* d_chi2;
    int *d_obs_pose, *d_obs_point, *d_obs_orig, *d_lm_start;
    uint8_t* d_inlier;
    BaScalars* d_sc;
    BaScalars* h_sc;  // pinned
    void* h_stage;    // pinned staging arena of the marshalled graph (ba_stage_bytes)
    void* h_out;      // pinned arena for the results of vslam_ba_optimize: poses, points, chi2 per edge, inlier flags
    // landmark-sharded session (multi-GPU): parameters of the open session, current buffer, trial counter
    BaParams* sess;
    int sess_open, sess_cur, sess_trials;
    double *sess_r1, *sess_r2, *sess_r3;
    double* d_dense;  // Y of the dense-SYRK probe (allocated on first use)
    size_t dense_bytes;
    // multi-device exchange areas (allocated on the first vslam_ba_optimize_multi call)
    double *x_big, *x_small, *d_bp_glob;
    unsigned* x_flags;
};

__global__ void ba_phase_kernel(const __grid_constant__ BaParams P, int phase, double lambda, int cur, int slot, double* r1,
                                double* r3);
struct BaMulti;
__global__ void ba_lm_multi_kernel(const __grid_constant__ BaParams P, const __grid_constant__ BaMulti M);

static size_t ba_stage_bytes(size_t K, size_t L, size_t O);
static size_t ba_small_smem_bytes(int K) {
    const size_t n = 6 * (size_t)K;
    size_t need = n * (n + 1) + n + 12 * (size_t)K;
    if (need < (BA_THREADS / 32) * 42) need = (BA_THREADS / 32) * 42;
    return need * sizeof(double);
}

// shared memory of one CTA: max(pose partials K*42, per-CTA S n*n+n, in-shared-memory Cholesky n*n+n) doubles
static int ba_smem_bytes(int K) {
    const size_t n = 6 * (size_t)K;
    size_t need = (BA_THREADS / 32) * 42;  // block-reduction scratch of the BUILD-B / SCHUR phases
    if (n <= BA_SMEM_CHOL_MAX && n * n + n > need) need = n * n + n;     // in-shared-memory augmented system [S | bs]
    if (n > BA_SMEM_CHOL_MAX && 30 * (n + 1) > need) need = 30 * (n + 1);  // grid-wide solver: BA_NB x (n + 1) panel (>= 7n)
    return (int)(need * sizeof(double));
}

int vslam_ba_init(vslam_ctx* ctx) {
    BaState* b = (BaState*)calloc(1, sizeof(BaState));
    if (!b) return VSLAM_E_INVALID;
    ctx->ba = b;
    const vslam_config& c = ctx->cfg;
    b->maxK = c.max_ba_poses;
    b->maxL = c.max_ba_points;
    b->maxObs = c.max_ba_obs;
    if (b->maxK <= 0 || b->maxL <= 0 || b->maxObs <= 0) return VSLAM_OK;  // BA disabled
    const size_t K = b->maxK, L = b->maxL, O = b->maxObs, n = 6 * K;
    b->n_cta = ctx->num_sms;
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_poses, 2 * K * 12 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_points, 2 * L * 3 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_uv, O * 2 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_err, O * 2 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_Hpl, O * 18 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_Hll, L * 9 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_bl, L * 3 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_Dinv, L * 9 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_Hpp, K * 36 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_bp, n * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_S, n * n * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_bs, n * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_x, (2 * n + 3 * L + 16) * sizeof(double)));  // + n + 16: staging of the small exchanges
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_dbl, L * 3 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_U, n * n * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_pose_start, (K + 1) * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_pose_obs, O * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_obs_of, K * L * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_block_flag, (K * (K + 1) / 2) * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_chi2, O * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_obs_pose, O * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_obs_point, O * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_obs_orig, O * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_lm_start, (L + 1) * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_inlier, L));
    VSLAM_CUDA(ctx, cudaMalloc(&b->d_sc, sizeof(BaScalars)));
    VSLAM_CUDA(ctx, cudaMallocHost(&b->h_sc, sizeof(BaScalars)));
    VSLAM_CUDA(ctx, cudaMallocHost(&b->h_stage, ba_stage_bytes(K, L, O)));
    VSLAM_CUDA(ctx, cudaMallocHost(&b->h_out, (12 * K + 3 * L + O + 8) * sizeof(double) + L + 64));
    // dynamic shared memory is sized per call from the actual K (ba_smem_bytes); opt in to the largest case here
    b->smem_bytes = 0;
    for (size_t k = 1; k <= K; ++k) {
        const int need = ba_smem_bytes((int)k);
        if (need > b->smem_bytes) b->smem_bytes = need;
    }
    if (b->smem_bytes > 227 * 1024) return VSLAM_E_CAPACITY;
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(ba_lm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, b->smem_bytes));
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(ba_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, b->smem_bytes));
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(ba_lm_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, b->smem_bytes));
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(ba_lm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ba_small_smem_bytes(BA_SMALL_N / 6)));
    return VSLAM_OK;
}

void vslam_ba_free(vslam_ctx* ctx) {
    BaState* b = ctx->ba;
    if (!b) return;
    cudaFree(b->d_poses); cudaFree(b->d_points); cudaFree(b->d_uv); cudaFree(b->d_err); cudaFree(b->d_Hpl);
    cudaFree(b->d_Hll); cudaFree(b->d_bl); cudaFree(b->d_Dinv); cudaFree(b->d_Hpp); cudaFree(b->d_bp); cudaFree(b->d_S);
    cudaFree(b->d_bs); cudaFree(b->d_x); cudaFree(b->d_dbl); cudaFree(b->d_U); cudaFree(b->d_block_flag); cudaFree(b->d_pose_start); cudaFree(b->d_pose_obs); cudaFree(b->d_obs_of); cudaFree(b->d_chi2); cudaFree(b->d_obs_pose);
    cudaFree(b->d_obs_point); cudaFree(b->d_obs_orig); cudaFree(b->d_lm_start); cudaFree(b->d_inlier); cudaFree(b->d_sc);
    cudaFreeHost(b->h_sc);
    if (b->h_stage) cudaFreeHost(b->h_stage);
    if (b->h_out) cudaFreeHost(b->h_out);
    if (b->d_dense) cudaFree(b->d_dense);
    if (b->x_big) { cudaFree(b->x_big); cudaFree(b->x_small); cudaFree(b->x_flags); cudaFree(b->d_bp_glob); }
    free(b->sess);
    free(b);
    ctx->ba = nullptr;
}

// Host marshalling shared by vslam_ba_optimize, vslam_ba_session_begin and vslam_ba_optimize_multi: stable counting
// sorts of the edge list by landmark (the device order) and by pose (CSR over device indices), duplicate (pose,
// landmark) detection -- once on the host -- then the uploads to one device.  Index bookkeeping only, no arithmetic on
// the measurements.
struct BaHostGraph {  // arrays live in the pinned staging arena of a context (BaState::h_stage)
    int K, n_points, n_obs, has_dup;
    int *lm_start, *op, *ol, *oo, *pose_start, *pose_obs;
    double *uv, *poses, *points;
};

static size_t ba_stage_bytes(size_t K, size_t L, size_t O) {
    return (L + 1 + K + 1 + 4 * O + 16) * sizeof(int) + (2 * O + 12 * K + 3 * L + 8) * sizeof(double);
}

static void ba_graph_bind(BaHostGraph& g, void* mem, int K, int n_points, int n_obs) {
    const size_t L = n_points > 0 ? n_points : 1, O = n_obs > 0 ? n_obs : 1;
    double* d = (double*)mem;  // doubles first: 8-byte alignment for free
    g.uv = d; d += 2 * O;
    g.poses = d; d += 12 * (size_t)K;
    g.points = d; d += 3 * L;
    int* q = (int*)d;
    g.lm_start = q; q += L + 1;
    g.pose_start = q; q += K + 1;
    g.op = q; q += O;
    g.ol = q; q += O;
    g.oo = q; q += O;
    g.pose_obs = q;
    g.K = K; g.n_points = n_points; g.n_obs = n_obs; g.has_dup = 0;
}

static int ba_prepare_host(BaHostGraph& g, const double* poses, const double* points, const int32_t* obs_pose,
                           const int32_t* obs_point, const double* obs_uv) {
    const int K = g.K, n_points = g.n_points, n_obs = g.n_obs;
    const int L = n_points > 0 ? n_points : 1;
    memset(g.lm_start, 0, (size_t)(L + 1) * sizeof(int));
    memset(g.pose_start, 0, (size_t)(K + 1) * sizeof(int));
    for (int i = 0; i < n_obs; ++i) {
        if (obs_point[i] < 0 || obs_point[i] >= n_points || obs_pose[i] < 0 || obs_pose[i] >= K) return VSLAM_E_INVALID;
        g.lm_start[obs_point[i] + 1]++;
        g.pose_start[obs_pose[i] + 1]++;
    }
    for (int l = 0; l < L; ++l) g.lm_start[l + 1] += g.lm_start[l];
    for (int k = 0; k < K; ++k) g.pose_start[k + 1] += g.pose_start[k];
    {
        std::vector<int> fill(g.lm_start, g.lm_start + L);
        for (int i = 0; i < n_obs; ++i) {
            const int d = fill[obs_point[i]]++;
            g.op[d] = obs_pose[i]; g.ol[d] = obs_point[i]; g.oo[d] = i;
            g.uv[2 * d] = obs_uv[2 * i]; g.uv[2 * d + 1] = obs_uv[2 * i + 1];
        }
    }
    {
        std::vector<int> fill(g.pose_start, g.pose_start + K);
        for (int d = 0; d < n_obs; ++d) g.pose_obs[fill[g.op[d]]++] = d;  // ascending device index within each pose
    }
    for (int l = 0; l < n_points && !g.has_dup; ++l)
        for (int i = g.lm_start[l]; i < g.lm_start[l + 1] && !g.has_dup; ++i)
            for (int j = i + 1; j < g.lm_start[l + 1]; ++j)
                if (g.op[i] == g.op[j]) { g.has_dup = 1; break; }
    memcpy(g.poses, poses, (size_t)K * 96);
    if (n_points > 0) memcpy(g.points, points, (size_t)n_points * 24);
    return VSLAM_OK;
}

// uploads (from the pinned arena: truly asynchronous) on the context stream of the CURRENT device
static int ba_upload(vslam_ctx* ctx, BaState* b, const BaHostGraph& g) {
    const int K = g.K, n_points = g.n_points, n_obs = g.n_obs;
    const int L = n_points > 0 ? n_points : 1;
    cudaStream_t s = ctx->stream;
    VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_poses, g.poses, (size_t)K * 96, cudaMemcpyHostToDevice, s));
    if (n_points > 0) {
        VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_points, g.points, (size_t)n_points * 24, cudaMemcpyHostToDevice, s));
        // pose-only mode never writes the trial points: keep both buffers equal
        VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_points + (size_t)L * 3, b->d_points, (size_t)n_points * 24, cudaMemcpyDeviceToDevice, s));
    }
    if (n_obs > 0) {
        VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_obs_pose, g.op, (size_t)n_obs * 4, cudaMemcpyHostToDevice, s));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_obs_point, g.ol, (size_t)n_obs * 4, cudaMemcpyHostToDevice, s));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_obs_orig, g.oo, (size_t)n_obs * 4, cudaMemcpyHostToDevice, s));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_pose_obs, g.pose_obs, (size_t)n_obs * 4, cudaMemcpyHostToDevice, s));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_uv, g.uv, (size_t)n_obs * 16, cudaMemcpyHostToDevice, s));
    }
    VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_lm_start, g.lm_start, (size_t)(L + 1) * 4, cudaMemcpyHostToDevice, s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_pose_start, g.pose_start, (size_t)(K + 1) * 4, cudaMemcpyHostToDevice, s));
    VSLAM_CUDA(ctx, cudaMemsetAsync(b->d_obs_of, 0xFF, (size_t)K * L * sizeof(int), s));
    VSLAM_CUDA(ctx, cudaMemsetAsync(b->d_block_flag, 0, (size_t)(K * (K + 1) / 2) * sizeof(int), s));
    return VSLAM_OK;
}

// the host-synchronous entry points finish (stream synchronised) before they return, so the arena is free again
static int ba_marshal(vslam_ctx* ctx, BaState* b, int K, int n_points, int n_obs, const double* poses,
                      const double* points, const int32_t* obs_pose, const int32_t* obs_point, const double* obs_uv,
                      int* has_dup) {
    BaHostGraph g;
    ba_graph_bind(g, b->h_stage, K, n_points, n_obs);
    int st = ba_prepare_host(g, poses, points, obs_pose, obs_point, obs_uv);
    if (st != VSLAM_OK) return st;
    *has_dup = g.has_dup;
    return ba_upload(ctx, b, g);
}

static void ba_fill_params(BaParams& P, BaState* b, int K, int n_points, int n_obs, const vslam_ba_options* opt,
                           const double* Kmat, int has_dup) {
    memset(&P, 0, sizeof(P));
    P.K = K; P.L = n_points; P.n_obs = n_obs; P.pose_only = opt->pose_only ? 1 : 0;
    P.num_iterations = opt->num_iterations; P.max_trials = opt->max_trials > 0 ? opt->max_trials : 10;
    P.delta = opt->huber_delta; P.tau = opt->tau > 0 ? opt->tau : 1e-5; P.chi2_th = opt->chi2_threshold;
    memcpy(P.Kc, Kmat, 72);
    P.n = 6 * K;
    P.n_cta = b->n_cta;
    P.shard_L0 = 0; P.shard_L1 = n_points;
    P.has_dup = has_dup;
    P.poses = b->d_poses; P.points = b->d_points; P.obs_pose = b->d_obs_pose; P.obs_point = b->d_obs_point;
    P.obs_uv = b->d_uv; P.obs_orig = b->d_obs_orig; P.lm_start = b->d_lm_start; P.pose_start = b->d_pose_start;
    P.pose_obs = b->d_pose_obs; P.obs_of = b->d_obs_of; P.dbl = b->d_dbl; P.Ubuf = b->d_U; P.block_flag = b->d_block_flag; P.err = b->d_err; P.Hpl = b->d_Hpl;
    P.Hll = b->d_Hll; P.bl = b->d_bl; P.Dinv = b->d_Dinv; P.Hpp = b->d_Hpp; P.bp = b->d_bp; P.S = b->d_S; P.bs = b->d_bs;
    P.x = b->d_x; P.sc = b->d_sc; P.chi2_out = b->d_chi2; P.inlier_out = b->d_inlier;
    P.bp_scale = P.bp;
}

extern "C" int vslam_ba_optimize(vslam_ctx* ctx, int n_poses, double* poses, int n_points, double* points, int n_obs,
                                 const int32_t* obs_pose, const int32_t* obs_point, const double* obs_uv,
                                 const double* Kmat, const vslam_ba_options* opt, vslam_ba_result* res,
                                 double* chi2_per_obs, uint8_t* point_inlier) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !poses || !Kmat || !opt || n_poses <= 0 || n_points < 0 || n_obs < 0) return VSLAM_E_INVALID;
    if (n_obs > 0 && (!obs_pose || !obs_point || !obs_uv)) return VSLAM_E_INVALID;
    if (n_points > 0 && !points) return VSLAM_E_INVALID;
    BaState* b = ctx->ba;
    if (!b || !b->d_poses) return VSLAM_E_CAPACITY;
    if (n_poses > b->maxK || n_points > b->maxL || n_obs > b->maxObs) return VSLAM_E_CAPACITY;
    const int K = n_poses;
    int has_dup = 0;
    int mst = ba_marshal(ctx, b, K, n_points, n_obs, poses, points, obs_pose, obs_point, obs_uv, &has_dup);
    if (mst != VSLAM_OK) return mst;
    cudaStream_t s = ctx->stream;
    if (point_inlier && n_points > 0)
        VSLAM_CUDA(ctx, cudaMemcpyAsync(b->d_inlier, point_inlier, (size_t)n_points, cudaMemcpyHostToDevice, s));
    BaParams P;
    ba_fill_params(P, b, K, n_points, n_obs, opt, Kmat, has_dup);

    void* args[] = {(void*)&P};
    static const bool force_general = getenv("VSLAM_BA_GENERAL") != nullptr;  // A/B switch for measurements
    const bool small = P.n <= BA_SMALL_N && !has_dup && !force_general;
    vslam_time_begin(ctx, VK_BA_BUILD);
    if (small) {
        VSLAM_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)ba_lm_small_kernel, dim3(b->n_cta), dim3(BA_THREADS), args,
                                                    ba_small_smem_bytes(K), s));
    } else {
        VSLAM_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)ba_lm_kernel, dim3(b->n_cta), dim3(BA_THREADS), args,
                                                    (size_t)ba_smem_bytes(K), s));
    }
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, small ? "ba_lm_small_kernel" : "ba_lm_kernel");
    // results land in a pinned arena (copies to the caller's pageable arrays would each be staged and waited for by
    // the driver) and are handed over after the one synchronisation
    double* h_poses = (double*)b->h_out;
    double* h_points = h_poses + 12 * (size_t)K;
    double* h_chi2 = h_points + 3 * (size_t)(n_points > 0 ? n_points : 1);
    uint8_t* h_inl = (uint8_t*)(h_chi2 + (n_obs > 0 ? n_obs : 1));
    const bool want_points = n_points > 0 && !P.pose_only;
    VSLAM_CUDA(ctx, cudaMemcpyAsync(b->h_sc, b->d_sc, sizeof(BaScalars), cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(h_poses, b->d_poses, (size_t)K * 96, cudaMemcpyDeviceToHost, s));
    if (want_points) VSLAM_CUDA(ctx, cudaMemcpyAsync(h_points, b->d_points, (size_t)n_points * 24, cudaMemcpyDeviceToHost, s));
    if (chi2_per_obs && n_obs > 0)
        VSLAM_CUDA(ctx, cudaMemcpyAsync(h_chi2, b->d_chi2, (size_t)n_obs * 8, cudaMemcpyDeviceToHost, s));
    if (point_inlier && n_points > 0)
        VSLAM_CUDA(ctx, cudaMemcpyAsync(h_inl, b->d_inlier, (size_t)n_points, cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    memcpy(poses, h_poses, (size_t)K * 96);
    if (want_points) memcpy(points, h_points, (size_t)n_points * 24);
    if (chi2_per_obs && n_obs > 0) memcpy(chi2_per_obs, h_chi2, (size_t)n_obs * 8);
    if (point_inlier && n_points > 0) memcpy(point_inlier, h_inl, (size_t)n_points);
    if (res) {
        res->iterations = b->h_sc->iterations;
        res->trials = b->h_sc->trials;
        res->accepted = b->h_sc->accepted;
        res->chi2_initial = b->h_sc->chi2_initial;
        res->chi2_final = b->h_sc->chi2_final;
        res->lambda_final = b->h_sc->lambda_final;
        res->chi2_threshold = b->h_sc->chi2_threshold;
        res->n_inlier_obs = b->h_sc->n_inlier_obs;
        res->n_outlier_obs = b->h_sc->n_outlier_obs;
    }
    return VSLAM_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Landmark-sharded session: one process per GPU owns the landmarks [shard_begin, shard_end) with all
// their observations; poses are replicated.  The caller runs the LM control flow and performs the collectives on the
// three device reduce buffers between the phases (stereo-visual-slam_b200/sharding.py does it with torch.distributed):
//   r1 = [Hpp (36K) | bp (6K) | chi2 | max|diag Hll|]     all-reduce(sum) once per outer iteration (+ max for r1[42K+1])
//   r2 = [S (6K x 6K) | bs (6K)]                           all-reduce(sum) once per LM trial: the reduced camera system
//   r3 = [chi2_trial | scale | ok | - | cnt_le[0..5]]      all-reduce(sum) once per LM trial (2 doubles) / once at the end
// Every rank then solves the identical 6K x 6K system redundantly and back-substitutes its own landmarks.
// ---------------------------------------------------------------------------------------------------------------
enum { BA_PH_BUILD = 1, BA_PH_IMPORT_BUILD, BA_PH_SCHUR, BA_PH_SOLVE_UPDATE, BA_PH_RELABEL_COUNT, BA_PH_RELABEL_APPLY };

__global__ void __launch_bounds__(BA_THREADS)
ba_phase_kernel(const __grid_constant__ BaParams P, int phase, double lambda, int cur, int slot, double* r1, double* r3) {
    extern __shared__ double smem[];
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    const int n = P.n;
    if (phase == BA_PH_BUILD) {
        ba_phase_zero(P, gtid, gsize);
        grid.sync();
        ba_phase_build_edges(P, cur, gtid, gsize);
        ba_phase_build_poses(P, cur, smem);
        grid.sync();
        ba_phase_maxdiag(P, gtid, gsize);
        grid.sync();
        if (gtid == 0) {
            r1[42 * P.K] = P.sc->chi_cur;
            r1[42 * P.K + 1] = __longlong_as_double((long long)P.sc->maxdiag_bits);
            for (int r = 0; r < 6; ++r) P.sc->cnt_le[r] = 0;
        }
    } else if (phase == BA_PH_IMPORT_BUILD) {
        if (gtid == 0) P.sc->chi_cur = r1[42 * P.K];
    } else if (phase == BA_PH_SCHUR) {
        // partial system only: Hpp + lambda I and bp are added after the all-reduce (SOLVE_UPDATE)
        ba_phase_dinv(P, lambda, slot, 0, gtid, gsize);
        grid.sync();
        if (P.has_dup) ba_phase_schur_atomic(P, gtid, gsize);
        else ba_phase_schur_blocks(P, smem);
    } else if (phase == BA_PH_SOLVE_UPDATE) {
        for (int i = gtid; i < n * n; i += gsize) {  // S (reduced over ranks) += Hpp (reduced) + lambda I
            const int r = i / n, c = i - r * n;
            double v = 0.0;
            if (r / 6 == c / 6 && c >= r) v = P.Hpp[(r / 6) * 36 + (r % 6) * 6 + (c % 6)];
            if (r == c) v += lambda;
            P.S[i] += v;
        }
        for (int i = gtid; i < n; i += gsize) P.bs[i] += P.bp[i];
        grid.sync();
        if (n <= BA_SMEM_CHOL_MAX) ba_phase_solve_cta(P, slot, smem);
        else ba_phase_solve_grid(P, slot, grid, smem, P.Ubuf, gtid, gsize);
        grid.sync();
        ba_phase_update(P, cur, lambda, slot, gtid, gsize);
        grid.sync();
        ba_phase_trial_err(P, cur, slot, gtid, gsize);
        grid.sync();
        if (gtid == 0) {
            r3[0] = P.sc->chi_trial[slot];
            r3[1] = P.sc->scale[slot];
            r3[2] = (double)P.sc->solve_ok[slot];
        }
    } else if (phase == BA_PH_RELABEL_COUNT) {
        ba_phase_relabel_count(P, gtid, gsize);
        grid.sync();
        if (gtid == 0)
            for (int r = 0; r < 6; ++r) r3[4 + r] = (double)P.sc->cnt_le[r];
    } else if (phase == BA_PH_RELABEL_APPLY) {
        ba_phase_relabel_apply(P, lambda /* carries the threshold */, gtid, gsize);
    }
}

extern "C" int vslam_ba_reduce_sizes(int n_poses, int* r1_doubles, int* r2_doubles, int* r3_doubles) {
    if (n_poses <= 0) return VSLAM_E_INVALID;
    const int n = 6 * n_poses;
    if (r1_doubles) *r1_doubles = 42 * n_poses + 2;
    if (r2_doubles) *r2_doubles = n * n + n;
    if (r3_doubles) *r3_doubles = 10;
    return VSLAM_OK;
}

extern "C" int vslam_ba_session_begin(vslam_ctx* ctx, int n_poses, const double* poses, int n_points,
                                      const double* points, int n_obs, const int32_t* obs_pose,
                                      const int32_t* obs_point, const double* obs_uv, const double* Kmat,
                                      const vslam_ba_options* opt, int shard_begin, int shard_end, double* d_r1,
                                      double* d_r2, double* d_r3) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !poses || !points || !Kmat || !opt || !d_r1 || !d_r2 || !d_r3) return VSLAM_E_INVALID;
    if (n_poses <= 0 || n_points <= 0 || n_obs <= 0 || !obs_pose || !obs_point || !obs_uv) return VSLAM_E_INVALID;
    if (shard_begin < 0 || shard_end > n_points || shard_begin > shard_end) return VSLAM_E_INVALID;
    BaState* b = ctx->ba;
    if (!b || !b->d_poses) return VSLAM_E_CAPACITY;
    if (n_poses > b->maxK || n_points > b->maxL || n_obs > b->maxObs) return VSLAM_E_CAPACITY;
    const int K = n_poses, L = n_points;
    int has_dup = 0;
    int mst = ba_marshal(ctx, b, K, n_points, n_obs, poses, points, obs_pose, obs_point, obs_uv, &has_dup);
    if (mst != VSLAM_OK) return mst;
    cudaStream_t s = ctx->stream;
    VSLAM_CUDA(ctx, cudaMemsetAsync(b->d_chi2, 0, (size_t)n_obs * 8, s));
    VSLAM_CUDA(ctx, cudaMemsetAsync(b->d_inlier, 0, (size_t)L, s));
    if (!b->sess) b->sess = (BaParams*)calloc(1, sizeof(BaParams));
    BaParams& P = *b->sess;
    ba_fill_params(P, b, K, n_points, n_obs, opt, Kmat, has_dup);
    P.shard_L0 = shard_begin; P.shard_L1 = shard_end;
    P.Hpp = d_r1; P.bp = d_r1 + 36 * K;          // pose blocks live in the caller's reduce buffer r1
    P.bp_scale = P.bp;
    P.S = d_r2; P.bs = d_r2 + (size_t)P.n * P.n;  // reduced camera system lives in r2
    b->sess_open = 1; b->sess_cur = 0; b->sess_trials = 0;
    b->sess_r1 = d_r1; b->sess_r2 = d_r2; b->sess_r3 = d_r3;
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));  // the pinned staging arena is free again when this returns
    return VSLAM_OK;
}

static int ba_session_launch(vslam_ctx* ctx, int phase, double lambda) {
    BaState* b = ctx->ba;
    if (!b || !b->sess_open) return VSLAM_E_INVALID;
    BaParams P = *b->sess;
    int cur = b->sess_cur, slot = b->sess_trials & 1;
    double* r1 = b->sess_r1;
    double* r3 = b->sess_r3;
    void* args[] = {(void*)&P, (void*)&phase, (void*)&lambda, (void*)&cur, (void*)&slot, (void*)&r1, (void*)&r3};
    vslam_time_begin(ctx, VK_BA_MISC);
    VSLAM_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)ba_phase_kernel, dim3(b->n_cta), dim3(BA_THREADS), args,
                                                (size_t)ba_smem_bytes(P.K), ctx->stream));
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "ba_phase_kernel");
    return VSLAM_OK;
}

// phase: 1 BUILD, 2 IMPORT_BUILD, 3 SCHUR(lambda), 4 SOLVE_UPDATE(lambda), 5 RELABEL_COUNT, 6 RELABEL_APPLY(threshold)
extern "C" int vslam_ba_session_phase(vslam_ctx* ctx, int phase, double value) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || phase < BA_PH_BUILD || phase > BA_PH_RELABEL_APPLY) return VSLAM_E_INVALID;
    return ba_session_launch(ctx, phase, value);
}

// LM verdict for the trial just evaluated: accept != 0 makes the trial estimate current (g2o discardTop), else pop()
extern "C" int vslam_ba_session_trial_done(vslam_ctx* ctx, int accept) {
    if (!ctx || !ctx->ba || !ctx->ba->sess_open) return VSLAM_E_INVALID;
    if (accept) ctx->ba->sess_cur = 1 - ctx->ba->sess_cur;
    ctx->ba->sess_trials++;
    return VSLAM_OK;
}

// download: poses (replicated), and this rank's shard of points / per-edge chi2 (insertion order; zeros elsewhere) /
// inlier flags (zeros elsewhere) so that a sum over ranks assembles the full arrays
extern "C" int vslam_ba_session_end(vslam_ctx* ctx, double* poses, double* points, double* chi2_per_obs,
                                    uint8_t* point_inlier) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !ctx->ba || !ctx->ba->sess_open) return VSLAM_E_INVALID;
    BaState* b = ctx->ba;
    const BaParams& P = *b->sess;
    cudaStream_t s = ctx->stream;
    const int cur = b->sess_cur;
    if (poses) VSLAM_CUDA(ctx, cudaMemcpyAsync(poses, P.poses + (size_t)cur * P.K * 12, (size_t)P.K * 96, cudaMemcpyDeviceToHost, s));
    if (points) {
        memset(points, 0, (size_t)P.L * 24);
        const double* src = P.points + (size_t)(P.pose_only ? 0 : cur) * P.L * 3;
        if (P.shard_L1 > P.shard_L0)
            VSLAM_CUDA(ctx, cudaMemcpyAsync(points + 3 * (size_t)P.shard_L0, src + 3 * (size_t)P.shard_L0,
                                            (size_t)(P.shard_L1 - P.shard_L0) * 24, cudaMemcpyDeviceToHost, s));
    }
    if (chi2_per_obs) VSLAM_CUDA(ctx, cudaMemcpyAsync(chi2_per_obs, b->d_chi2, (size_t)P.n_obs * 8, cudaMemcpyDeviceToHost, s));
    if (point_inlier) VSLAM_CUDA(ctx, cudaMemcpyAsync(point_inlier, b->d_inlier, (size_t)P.L, cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    b->sess_open = 0;
    return VSLAM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// K17b: ONE window on several GPUs of one process, the whole Levenberg-Marquardt loop device-side.
// Rank r (= device r of the call) owns a contiguous landmark range with all its observations, poses are replicated.
// Every rank runs the same persistent cooperative kernel; the ranks meet through peer memory over NVLink:
//   * big exchange, once per LM trial: each rank stores its partial [S (upper block rows) | bs | bp | chi2 at the
//     linearisation point] straight into slot `rank` of EVERY rank's receive area (P2P stores), publishes an epoch flag
//     with st.release.sys, waits for the other ranks' flags with ld.acquire.sys, and sums the slots in rank order --
//     all-gather + local reduction in one hop, no host, no collective library, bit-identical sums on every rank;
//   * small exchange (a few doubles riding on the same flag protocol): lambda initialisation (max |diag|), the trial's
//     [chi2, scale, ok] for the accept/reject decision, the relabel counts.
// Every rank then solves the identical reduced camera system redundantly and back-substitutes its own landmarks, so
// the LM control flow needs no broadcast.  A peer that never shows up trips a bounded wait (trap), not a hang.
// ---------------------------------------------------------------------------------------------------------------
#define BA_MAX_DEV 8
#define BA_XS 512  // doubles per small-exchange slot (>= 6 * 64 + 8)
struct BaMulti {
    int rank, world;
    unsigned epoch0;              // flags only ever grow: base epoch of this launch
    int xlen;                     // doubles per big slot: n*n + 2n + 8
    double* xbig[BA_MAX_DEV];     // receive area of device g: [world][xlen]
    double* xsmall[BA_MAX_DEV];   // [2][world][BA_XS]
    unsigned* flags[BA_MAX_DEV];  // [world]: epoch of the last exchange rank s has published towards device g
    double* bp_glob;              // local [n]: sum over ranks of bp
};

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Called by every thread after its peer stores of this exchange.  Returns when every rank's data of `epoch` is visible.
// The grid barrier orders every thread's peer stores before the publishing threads (gpu scope); ONE fence.sys + release
// store per peer then publishes them at system scope (cumulativity) -- a fence.sys in every thread costs ~100 us here.
__device__ void ba_multi_signal_wait(const BaMulti& M, cg::grid_group& grid, unsigned epoch, unsigned long long* wait_ns) {
    const bool prof = blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long t_a = prof ? gtimer() : 0ull;
    grid.sync();
    const unsigned long long t_pub = gtimer();
    if (prof) wait_ns[1 - 11] += t_pub - t_a;  // phase_ns[1]: grid barrier after the peer stores
    if (blockIdx.x == 0 && (int)threadIdx.x < M.world) {
        __threadfence_system();
        st_release_sys_u32(M.flags[threadIdx.x] + M.rank, epoch);
    }
    if (prof) {
        const unsigned long long t_b = gtimer();
        wait_ns[2 - 11] += t_b - t_pub;        // phase_ns[2]: fence.sys + release stores
    }
    if ((int)threadIdx.x < M.world) {
        const unsigned* f = M.flags[M.rank] + threadIdx.x;
        const unsigned long long t0 = gtimer();
        while ((int)(ld_acquire_sys_u32(f) - epoch) < 0) {
            if (gtimer() - t0 > 4000000000ull) __trap();  // 4 s: a peer kernel is not running
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *wait_ns += gtimer() - t_pub;  // publish + wait for the slowest peer
}

// big exchange + rank-ordered sum: S <- sum_g S_g + lambda I, bs <- sum_g bs_g, bp_glob <- sum_g bp_g; returns sum_g chi_g
__device__ double ba_multi_exchange_system(const BaParams& P, const BaMulti& M, cg::grid_group& grid, unsigned epoch,
                                           double lambda, double chi_local, int gtid, int gsize) {
    const int n = P.n, lane = threadIdx.x & 31, gw = gtid >> 5, nw = gsize >> 5;
    const size_t slot = (size_t)M.rank * M.xlen;
    const unsigned long long t_w0 = gtid == 0 ? gtimer() : 0ull;
    for (int r = gw; r < n; r += nw) {  // one warp per row: the row segment right of (and including) the diagonal block
        const int c0 = 6 * (r / 6);
        for (int c = c0 + lane; c < n; c += 32) {
            const double v = P.S[(size_t)r * n + c];
            for (int g = 0; g < M.world; ++g) M.xbig[g][slot + (size_t)r * n + c] = v;
        }
    }
    for (int i = gtid; i < 2 * n + 1; i += gsize) {
        const double v = i < n ? P.bs[i] : i < 2 * n ? P.bp[i - n] : chi_local;
        for (int g = 0; g < M.world; ++g) M.xbig[g][slot + (size_t)n * n + i] = v;
    }
    if (gtid == 0) P.sc->phase_ns[0] += gtimer() - t_w0;  // this thread's share of the peer-store loops
    ba_multi_signal_wait(M, grid, epoch, &P.sc->phase_ns[11]);
    const unsigned long long t_s0 = gtid == 0 ? gtimer() : 0ull;
    const double* rx = M.xbig[M.rank];
    for (int r = gw; r < n; r += nw) {
        const int c0 = 6 * (r / 6);
        for (int c = c0 + lane; c < n; c += 32) {
            double v = 0.0;
            for (int g = 0; g < M.world; ++g) v += __ldcg(rx + (size_t)g * M.xlen + (size_t)r * n + c);
            P.S[(size_t)r * n + c] = r == c ? v + lambda : v;
        }
    }
    for (int i = gtid; i < 2 * n; i += gsize) {
        double v = 0.0;
        for (int g = 0; g < M.world; ++g) v += __ldcg(rx + (size_t)g * M.xlen + (size_t)n * n + i);
        if (i < n) P.bs[i] = v;
        else M.bp_glob[i - n] = v;
    }
    double chi = 0.0;
    for (int g = 0; g < M.world; ++g) chi += __ldcg(rx + (size_t)g * M.xlen + (size_t)n * n + 2 * n);
    grid.sync();
    if (gtid == 0) P.sc->phase_ns[4] += gtimer() - t_s0;  // rank-ordered sum + barrier
    return chi;
}

// small exchange: thread j < len of every rank publishes its value v (slot j of the rank's area on every device);
// afterwards slot g of the returned area holds rank g's values.  Two parities so that two consecutive small exchanges
// never share a slot.
__device__ const double* ba_multi_exchange_small(const BaMulti& M, cg::grid_group& grid, unsigned epoch, double v, int len,
                                                 int gtid, unsigned long long* wait_ns) {
    const int par = epoch & 1;
    const size_t off = ((size_t)par * M.world + M.rank) * BA_XS;
    if (gtid < len)
        for (int g = 0; g < M.world; ++g) M.xsmall[g][off + gtid] = v;
    ba_multi_signal_wait(M, grid, epoch, wait_ns);
    return M.xsmall[M.rank] + (size_t)par * M.world * BA_XS;
}

__global__ void __launch_bounds__(BA_THREADS)
ba_lm_multi_kernel(const __grid_constant__ BaParams P, const __grid_constant__ BaMulti M) {
    extern __shared__ double smem[];
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    BaScalars* sc = P.sc;
    unsigned epoch = M.epoch0;
    int cur = 0, trials = 0, accepted = 0, it = 0;
    double lambda = 0.0, ni = 2.0, chi_first = 0.0, chi_last = 0.0;
    // device-side profile (thread 0): [8] = start .. first exchange done (includes the launch skew between the GPUs),
    // [9] = first exchange done .. end, [10] = time inside the big exchanges, [11] = publish + wait of all exchanges
    unsigned long long t_start = 0, t_first = 0;
    if (gtid == 0) {
        for (int r = 0; r < 6; ++r) sc->cnt_le[r] = 0;
        for (int r = 0; r < 12; ++r) sc->phase_ns[r] = 0;
        t_start = gtimer();
    }
    for (it = 0; it < P.num_iterations; ++it) {
        ba_phase_zero(P, gtid, gsize);
        grid.sync();
        ba_phase_build_edges(P, cur, gtid, gsize);
        ba_phase_build_poses(P, cur, smem);
        if (it == 0) {
            grid.sync();
            ba_phase_maxdiag(P, gtid, gsize);
        }
        grid.sync();
        const double chi_local = sc->chi_cur;
        if (it == 0) {
            // computeLambdaInit: tau * max |diagonal| over the landmark blocks of all ranks and the SUMMED pose blocks
            const double v0 = gtid < P.n ? P.Hpp[(gtid / 6) * 36 + (gtid % 6) * 7]
                                         : gtid == P.n ? __longlong_as_double((long long)sc->maxdiag_bits) : 0.0;
            const double* rs = ba_multi_exchange_small(M, grid, ++epoch, v0, P.n + 1, gtid, &sc->phase_ns[11]);
            if (gtid == 0) {
                t_first = gtimer();
                sc->phase_ns[11] = 0;  // the first wait is the launch skew, accounted in [8]
            }
            __shared__ double s_md;
            if (threadIdx.x == 0) {
                double md = 0.0;
                for (int g = 0; g < M.world; ++g) md = fmax(md, __ldcg(rs + (size_t)g * BA_XS + P.n));
                for (int i = 0; i < P.n; ++i) {
                    double d = 0.0;
                    for (int g = 0; g < M.world; ++g) d += __ldcg(rs + (size_t)g * BA_XS + i);
                    md = fmax(md, fabs(d));
                }
                s_md = md;
            }
            __syncthreads();
            lambda = P.tau * s_md;
            ni = 2.0;
        }
        double currentChi = 0.0, rho = 0.0;
        int qmax = 0;
        do {
            const int slot = trials & 1;
            ba_phase_dinv(P, lambda, slot, 2, gtid, gsize);  // S = this rank's Hpp (no lambda), bs = its bp
            grid.sync();
            if (P.has_dup) ba_phase_schur_atomic(P, gtid, gsize);
            else ba_phase_schur_blocks(P, smem);
            grid.sync();
            const unsigned long long t_x0 = gtid == 0 ? gtimer() : 0ull;
            const double chi_sum = ba_multi_exchange_system(P, M, grid, ++epoch, lambda, chi_local, gtid, gsize);
            if (gtid == 0) sc->phase_ns[10] += gtimer() - t_x0;
            if (qmax == 0) {
                currentChi = chi_sum;
                if (it == 0) chi_first = currentChi;
            }
            if (P.n <= BA_SMEM_CHOL_MAX) ba_phase_solve_cta(P, slot, smem);
            else ba_phase_solve_grid(P, slot, grid, smem, P.Ubuf, gtid, gsize);
            grid.sync();
            ba_phase_update(P, cur, lambda, slot, gtid, gsize);
            grid.sync();
            ba_phase_trial_err(P, cur, slot, gtid, gsize);
            grid.sync();
            const double v3 = gtid == 0 ? sc->chi_trial[slot] : gtid == 1 ? sc->scale[slot] : gtid == 2 ? (double)sc->solve_ok[slot] : 0.0;
            const double* rs = ba_multi_exchange_small(M, grid, ++epoch, v3, 3, gtid, &sc->phase_ns[11]);
            double tchi = 0.0, tscale = 0.0, tok = 0.0;
            for (int g = 0; g < M.world; ++g) {
                tchi += __ldcg(rs + (size_t)g * BA_XS);
                tscale += __ldcg(rs + (size_t)g * BA_XS + 1);
                tok += __ldcg(rs + (size_t)g * BA_XS + 2);
            }
            const int ok2 = tok > M.world - 0.5;
            const double tempChi = ok2 ? tchi : DBL_MAX;
            const double scale = (ok2 ? tscale : 0.0) + 1e-3;
            rho = (currentChi - tempChi) / scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
                cur = 1 - cur;
                accepted++;
            } else {
                lambda *= ni;
                ni *= 2;
            }
            qmax++;
            trials++;
        } while (rho < 0 && qmax < P.max_trials);
        chi_last = currentChi;
        if (qmax == P.max_trials || rho == 0) {
            it++;
            break;
        }
    }
    if (P.num_iterations <= 0) {
        ba_phase_zero(P, gtid, gsize);
        grid.sync();
        ba_phase_build_edges(P, cur, gtid, gsize);
        grid.sync();
        const double* rs = ba_multi_exchange_small(M, grid, ++epoch, gtid == 0 ? sc->chi_cur : 0.0, 1, gtid, &sc->phase_ns[11]);
        double c = 0.0;
        for (int g = 0; g < M.world; ++g) c += __ldcg(rs + (size_t)g * BA_XS);
        chi_first = chi_last = c;
    }
    // relabel: counts summed over the ranks, identical threshold everywhere
    grid.sync();
    ba_phase_relabel_count(P, gtid, gsize);
    grid.sync();
    const double* rc = ba_multi_exchange_small(M, grid, ++epoch, gtid < 6 ? (double)sc->cnt_le[gtid] : 0.0, 6, gtid, &sc->phase_ns[11]);
    int cnt[6];
    for (int r = 0; r < 6; ++r) {
        double c = 0.0;
        for (int g = 0; g < M.world; ++g) c += __ldcg(rc + (size_t)g * BA_XS + r);
        cnt[r] = (int)(c + 0.5);
    }
    double th = P.chi2_th;
    int rr = 0;
    for (; rr < 5; ++rr) {
        if (cnt[rr] / (double)P.n_obs > 0.5) break;
        th *= 2;
    }
    ba_phase_relabel_apply(P, th, gtid, gsize);
    if (cur == 1) {
        for (int i = gtid; i < P.K * 12; i += gsize) P.poses[i] = P.poses[P.K * 12 + i];
        if (!P.pose_only)
            for (int i = gtid; i < P.L * 3; i += gsize) P.points[i] = P.points[(size_t)P.L * 3 + i];
    }
    if (gtid == 0) {
        sc->iterations = it;
        sc->trials = trials;
        sc->accepted = accepted;
        sc->chi2_initial = chi_first;
        sc->chi2_final = chi_last;
        sc->lambda_final = lambda;
        sc->chi2_threshold = th;
        sc->n_inlier_obs = cnt[rr];
        sc->n_outlier_obs = P.n_obs - cnt[rr];
        sc->pad = (int)(epoch - M.epoch0);  // exchanges executed
        const unsigned long long t_end = gtimer();
        if (t_first == 0) t_first = t_start;
        sc->phase_ns[8] = t_first - t_start;
        sc->phase_ns[9] = t_end - t_first;
    }
}

// contiguous landmark ranges with (nearly) equal numbers of observations
static void ba_landmark_shards(const BaHostGraph& g, int world, int* cuts /*[world + 1]*/) {
    const long long total = g.n_obs;
    cuts[0] = 0;
    int l = 0;
    for (int r = 1; r < world; ++r) {
        const long long target = total * r / world;
        while (l < g.n_points && g.lm_start[l] < target) ++l;
        cuts[r] = l < cuts[r - 1] ? cuts[r - 1] : l;
    }
    cuts[world] = g.n_points;
}

static unsigned g_ba_multi_epoch = 0;  // grows by 65536 per call: flag words are never reset

extern "C" int vslam_ba_optimize_multi(vslam_ctx* const* ctxs, int n_dev, int n_poses, double* poses, int n_points,
                                       double* points, int n_obs, const int32_t* obs_pose, const int32_t* obs_point,
                                       const double* obs_uv, const double* Kmat, const vslam_ba_options* opt,
                                       vslam_ba_result* res, double* chi2_per_obs, uint8_t* point_inlier) {
    if (!ctxs || n_dev < 1 || n_dev > BA_MAX_DEV) return VSLAM_E_INVALID;
    for (int d = 0; d < n_dev; ++d)
        if (!ctxs[d] || !ctxs[d]->ba || !ctxs[d]->ba->d_poses) return VSLAM_E_INVALID;
    if (n_dev == 1)
        return vslam_ba_optimize(ctxs[0], n_poses, poses, n_points, points, n_obs, obs_pose, obs_point, obs_uv, Kmat, opt, res,
                                 chi2_per_obs, point_inlier);
    if (!poses || !points || !Kmat || !opt || n_poses <= 0 || n_points <= 0 || n_obs <= 0 || !obs_pose || !obs_point || !obs_uv)
        return VSLAM_E_INVALID;
    for (int d = 0; d < n_dev; ++d) {
        BaState* b = ctxs[d]->ba;
        if (n_poses > b->maxK || n_points > b->maxL || n_obs > b->maxObs) return VSLAM_E_CAPACITY;
        for (int e = 0; e < d; ++e)
            if (ctxs[e]->cfg.device == ctxs[d]->cfg.device) return VSLAM_E_INVALID;  // one context per device
    }
    int dev0 = 0;
    cudaGetDevice(&dev0);
    const int K = n_poses, n = 6 * K;
    BaHostGraph g;
    ba_graph_bind(g, ctxs[0]->ba->h_stage, K, n_points, n_obs);
    int st = ba_prepare_host(g, poses, points, obs_pose, obs_point, obs_uv);
    if (st != VSLAM_OK) return st;
    int cuts[BA_MAX_DEV + 1];
    ba_landmark_shards(g, n_dev, cuts);
    for (int d = 0; d < n_dev; ++d)
        if (cuts[d + 1] <= cuts[d]) return VSLAM_E_INVALID;  // fewer landmarks than devices: use fewer devices
    const int xlen = n * n + 2 * n + 8;
    vslam_ctx* c0 = ctxs[0];
    std::vector<int> shard_csr((size_t)n_dev * (K + 1 + n_obs));  // stays alive until the final stream synchronisations
#define MULTI_CUDA(d, call)                                                            \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            vslam_set_cuda_error(ctxs[d], e__, #call);                                 \
            if ((d) != 0) vslam_set_cuda_error(c0, e__, #call);                        \
            cudaSetDevice(dev0);                                                       \
            return VSLAM_E_CUDA;                                                       \
        }                                                                              \
    } while (0)
    // peer access + exchange areas (first call per context), uploads
    for (int d = 0; d < n_dev; ++d) {
        vslam_ctx* ctx = ctxs[d];
        BaState* b = ctx->ba;
        MULTI_CUDA(d, cudaSetDevice(ctx->cfg.device));
        for (int e = 0; e < n_dev; ++e) {
            if (e == d) continue;
            int can = 0;
            MULTI_CUDA(d, cudaDeviceCanAccessPeer(&can, ctx->cfg.device, ctxs[e]->cfg.device));
            if (!can) {
                cudaSetDevice(dev0);
                snprintf(c0->err, sizeof(c0->err), "no peer access between devices %d and %d", ctx->cfg.device, ctxs[e]->cfg.device);
                return VSLAM_E_NODEVICE;
            }
            const cudaError_t pe = cudaDeviceEnablePeerAccess(ctxs[e]->cfg.device, 0);
            if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else MULTI_CUDA(d, pe);
        }
        if (!b->x_big) {
            const size_t nmax = 6 * (size_t)b->maxK;
            MULTI_CUDA(d, cudaMalloc(&b->x_big, BA_MAX_DEV * (nmax * nmax + 2 * nmax + 8) * sizeof(double)));
            MULTI_CUDA(d, cudaMalloc(&b->x_small, 2 * BA_MAX_DEV * BA_XS * sizeof(double)));
            MULTI_CUDA(d, cudaMalloc(&b->x_flags, BA_MAX_DEV * sizeof(unsigned)));
            MULTI_CUDA(d, cudaMalloc(&b->d_bp_glob, nmax * sizeof(double)));
            MULTI_CUDA(d, cudaMemset(b->x_flags, 0, BA_MAX_DEV * sizeof(unsigned)));
            // flags written by peers before must not be lost: a fresh context starts at the running epoch
            if (g_ba_multi_epoch) {
                unsigned init[BA_MAX_DEV];
                for (int q = 0; q < BA_MAX_DEV; ++q) init[q] = g_ba_multi_epoch;
                MULTI_CUDA(d, cudaMemcpy(b->x_flags, init, sizeof(init), cudaMemcpyHostToDevice));
            }
            MULTI_CUDA(d, cudaDeviceSynchronize());
        }
        st = ba_upload(ctx, b, g);
        if (st != VSLAM_OK) { cudaSetDevice(dev0); return st; }
        {
            // pose-major CSR restricted to this rank's edges: the per-pose and per-block phases then scan only owned edges
            // instead of filtering the whole list (landmark-sorted edges of the shard are the contiguous range [e0, e1))
            const int e0 = g.lm_start[cuts[d]], e1 = g.lm_start[cuts[d + 1]];
            int* ps = shard_csr.data() + (size_t)d * (K + 1 + n_obs);
            int* po = ps + K + 1;
            for (int k = 0; k <= K; ++k) ps[k] = 0;
            for (int e = e0; e < e1; ++e) ps[g.op[e] + 1]++;
            for (int k = 0; k < K; ++k) ps[k + 1] += ps[k];
            std::vector<int> fill(ps, ps + K);
            for (int e = e0; e < e1; ++e) po[fill[g.op[e]]++] = e;
            MULTI_CUDA(d, cudaMemcpyAsync(b->d_pose_start, ps, (size_t)(K + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
            if (e1 > e0) MULTI_CUDA(d, cudaMemcpyAsync(b->d_pose_obs, po, (size_t)(e1 - e0) * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        MULTI_CUDA(d, cudaMemsetAsync(b->d_chi2, 0, (size_t)n_obs * 8, ctx->stream));
        if (point_inlier) MULTI_CUDA(d, cudaMemcpyAsync(b->d_inlier, point_inlier, (size_t)n_points, cudaMemcpyHostToDevice, ctx->stream));
    }
    const unsigned epoch0 = g_ba_multi_epoch;
    g_ba_multi_epoch += 65536u;
    // launch the same kernel on every device; they meet through peer memory
    for (int d = 0; d < n_dev; ++d) {
        vslam_ctx* ctx = ctxs[d];
        BaState* b = ctx->ba;
        MULTI_CUDA(d, cudaSetDevice(ctx->cfg.device));
        BaParams P;
        ba_fill_params(P, b, K, n_points, n_obs, opt, Kmat, g.has_dup);
        P.shard_L0 = cuts[d]; P.shard_L1 = cuts[d + 1];
        BaMulti M;
        memset(&M, 0, sizeof(M));
        M.rank = d; M.world = n_dev; M.epoch0 = epoch0; M.xlen = xlen;
        for (int e = 0; e < n_dev; ++e) {
            M.xbig[e] = ctxs[e]->ba->x_big;
            M.xsmall[e] = ctxs[e]->ba->x_small;
            M.flags[e] = ctxs[e]->ba->x_flags;
        }
        M.bp_glob = b->d_bp_glob;
        P.bp_scale = b->d_bp_glob;
        void* args[] = {(void*)&P, (void*)&M};
        vslam_time_begin(ctx, VK_BA_BUILD);
        MULTI_CUDA(d, cudaLaunchCooperativeKernel((const void*)ba_lm_multi_kernel, dim3(b->n_cta), dim3(BA_THREADS), args,
                                                  (size_t)ba_smem_bytes(K), ctx->stream));
        vslam_time_end(ctx);
        ctx->launches++;
    }
    // results: poses and scalars from rank 0; points / inlier flags from the owner of each landmark range; per-edge
    // chi2 is zero outside a rank's shard, so the per-rank arrays add up
    std::vector<double> chi_tmp;
    for (int d = 0; d < n_dev; ++d) {
        vslam_ctx* ctx = ctxs[d];
        BaState* b = ctx->ba;
        MULTI_CUDA(d, cudaSetDevice(ctx->cfg.device));
        cudaStream_t s = ctx->stream;
        if (d == 0) {
            MULTI_CUDA(d, cudaMemcpyAsync(b->h_sc, b->d_sc, sizeof(BaScalars), cudaMemcpyDeviceToHost, s));
            MULTI_CUDA(d, cudaMemcpyAsync(poses, b->d_poses, (size_t)K * 96, cudaMemcpyDeviceToHost, s));
        }
        const int l0 = cuts[d], l1 = cuts[d + 1];
        if (l1 > l0) {
            if (!opt->pose_only)
                MULTI_CUDA(d, cudaMemcpyAsync(points + 3 * (size_t)l0, b->d_points + 3 * (size_t)l0, (size_t)(l1 - l0) * 24,
                                              cudaMemcpyDeviceToHost, s));
            if (point_inlier)
                MULTI_CUDA(d, cudaMemcpyAsync(point_inlier + l0, b->d_inlier + l0, (size_t)(l1 - l0), cudaMemcpyDeviceToHost, s));
        }
    }
    if (chi2_per_obs) {
        chi_tmp.resize((size_t)n_obs * n_dev);
        for (int d = 0; d < n_dev; ++d) {
            MULTI_CUDA(d, cudaSetDevice(ctxs[d]->cfg.device));
            MULTI_CUDA(d, cudaMemcpyAsync(chi_tmp.data() + (size_t)d * n_obs, ctxs[d]->ba->d_chi2, (size_t)n_obs * 8,
                                          cudaMemcpyDeviceToHost, ctxs[d]->stream));
        }
    }
    for (int d = 0; d < n_dev; ++d) {
        MULTI_CUDA(d, cudaSetDevice(ctxs[d]->cfg.device));
        MULTI_CUDA(d, cudaStreamSynchronize(ctxs[d]->stream));
    }
    cudaSetDevice(dev0);
#undef MULTI_CUDA
    if (chi2_per_obs) {
        for (int i = 0; i < n_obs; ++i) {
            double v = 0.0;
            for (int d = 0; d < n_dev; ++d) v += chi_tmp[(size_t)d * n_obs + i];
            chi2_per_obs[i] = v;
        }
    }
    if (res) {
        const BaScalars* h = c0->ba->h_sc;
        res->iterations = h->iterations; res->trials = h->trials; res->accepted = h->accepted; res->reserved = h->pad;
        res->chi2_initial = h->chi2_initial; res->chi2_final = h->chi2_final; res->lambda_final = h->lambda_final;
        res->chi2_threshold = h->chi2_threshold; res->n_inlier_obs = h->n_inlier_obs; res->n_outlier_obs = h->n_outlier_obs;
    }
    return VSLAM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// N1 probe: the reduced camera system as ONE dense fp64 SYRK on the tensor cores (DMMA, mma.sync.m8n8k4.f64), to be
// timed beside the sparse per-block formation above.  With Dinv_l = C_l C_l^T (3x3 Cholesky) and Y = [Hpl_l C_l]_l
// (6K x 3L, dense storage, zero where pose k does not see landmark l):  sum_l Hpl_l Dinv_l Hpl_l^T = Y Y^T.
// BlockSolver_6_3 (optimization.cpp:111-120) forms the same product block-sparsely; the dense form spends
// (6K)^2 * 3L * 2 flops (10.8 GFLOP at K=50 / L=20000 against 77 MFLOP sparse).
// ---------------------------------------------------------------------------------------------------------------
#define DS_TM 64   // output tile (rows = cols)
#define DS_TK 32   // k chunk staged in shared memory
#define DS_LD 36   // row stride in doubles: 36 mod 16 == 4 -> the 16 lanes of a half-warp (4 rows x 4 k) hit 16 banks

__global__ void ba_dense_fill_kernel(const __grid_constant__ BaParams P, double* __restrict__ Y, int ldY) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_obs) return;
    const int k = P.obs_pose[i], l = P.obs_point[i];
    const double* d = P.Dinv + 9 * (size_t)l;
    // lower Cholesky factor of the SPD 3x3 block
    const double c00 = sqrt(d[0]), c10 = d[3] / c00, c20 = d[6] / c00;
    const double c11 = sqrt(d[4] - c10 * c10), c21 = (d[7] - c20 * c10) / c11;
    const double c22 = sqrt(d[8] - c20 * c20 - c21 * c21);
    const double* B = P.Hpl + 18 * (size_t)i;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        double* y = Y + (size_t)(6 * k + a) * ldY + 3 * l;
        y[0] = B[a * 3] * c00 + B[a * 3 + 1] * c10 + B[a * 3 + 2] * c20;
        y[1] = B[a * 3 + 1] * c11 + B[a * 3 + 2] * c21;
        y[2] = B[a * 3 + 2] * c22;
    }
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// S(upper tiles) -= Y Y^T.  grid = (tile pairs I <= J, k slices); 4 warps, warp (wm, wn) owns a 32x32 sub-tile = 4x4
// DMMA tiles.  Fragment layout of m8n8k4: A[row = lane/4][k = lane%4], B[k = lane%4][col = lane/4],
// C[row = lane/4][col = 2*(lane%4) + {0,1}].
__global__ void __launch_bounds__(128)
ba_dense_syrk_kernel(const double* __restrict__ Y, int ldY, int n, int tiles, int chunks_per_slice, double* __restrict__ S) {
    __shared__ double As[DS_TM * DS_LD], Bs[DS_TM * DS_LD];
    int I = 0, rem = blockIdx.x;
    while (rem >= tiles - I) { rem -= tiles - I; ++I; }
    const int J = I + rem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 1, wn = warp & 1, g = lane >> 2, t = lane & 3;
    const int n_chunks = ldY / DS_TK;
    const int c_begin = blockIdx.y * chunks_per_slice, c_end = min(n_chunks, c_begin + chunks_per_slice);
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const double* Bt = (I == J) ? As : Bs;
    for (int ch = c_begin; ch < c_end; ++ch) {
        const size_t k0 = (size_t)ch * DS_TK;
        for (int e = tid; e < DS_TM * (DS_TK / 2); e += 128) {  // 16-byte loads, k contiguous
            const int row = e / (DS_TK / 2), c2 = e % (DS_TK / 2);
            const int rI = I * DS_TM + row, rJ = J * DS_TM + row;
            double2 va = make_double2(0.0, 0.0), vb = make_double2(0.0, 0.0);
            if (rI < n) va = *reinterpret_cast<const double2*>(Y + (size_t)rI * ldY + k0 + 2 * c2);
            if (I != J && rJ < n) vb = *reinterpret_cast<const double2*>(Y + (size_t)rJ * ldY + k0 + 2 * c2);
            As[row * DS_LD + 2 * c2] = va.x;
            As[row * DS_LD + 2 * c2 + 1] = va.y;
            if (I != J) {
                Bs[row * DS_LD + 2 * c2] = vb.x;
                Bs[row * DS_LD + 2 * c2 + 1] = vb.y;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < DS_TK; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = As[(wm * 32 + q * 8 + g) * DS_LD + kk + t];
                b[q] = Bt[(wn * 32 + q * 8 + g) * DS_LD + kk + t];
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = I * DS_TM + wm * 32 + mi * 8 + g, c = J * DS_TM + wn * 32 + ni * 8 + 2 * t + h;
                if (r < n && c < n && c >= r && acc[mi][ni][h] != 0.0) atomicAdd(&S[(size_t)r * n + c], -acc[mi][ni][h]);
            }
}

// After vslam_ba_session_phase(SCHUR) of an open session covering ALL landmarks: forms -(Hpl Hll^-1 Hpl^T) densely into
// d_S_dense (n x n device doubles, upper triangle r <= c written, the rest zero) and reports the device time of the
// two kernels.  The session's own (sparse) result stays in its r2 buffer for comparison.
extern "C" int vslam_ba_session_schur_dense(vslam_ctx* ctx, double* d_S_dense, float* ms_fill, float* ms_syrk) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !ctx->ba || !ctx->ba->sess_open || !d_S_dense) return VSLAM_E_INVALID;
    BaState* b = ctx->ba;
    BaParams P = *b->sess;
    if (P.pose_only || P.has_dup || P.shard_L0 != 0 || P.shard_L1 != P.L || P.n_obs <= 0) return VSLAM_E_INVALID;
    const int n = P.n;
    const int ldY = ceil_div(3 * P.L, DS_TK) * DS_TK;
    const size_t bytes = (size_t)n * ldY * sizeof(double);
    if (b->dense_bytes < bytes) {
        if (b->d_dense) cudaFree(b->d_dense);
        b->d_dense = nullptr;
        b->dense_bytes = 0;
        VSLAM_CUDA(ctx, cudaMalloc(&b->d_dense, bytes));
        b->dense_bytes = bytes;
    }
    cudaStream_t s = ctx->stream;
    cudaEvent_t e0, e1, e2;
    VSLAM_CUDA(ctx, cudaEventCreate(&e0));
    VSLAM_CUDA(ctx, cudaEventCreate(&e1));
    VSLAM_CUDA(ctx, cudaEventCreate(&e2));
    VSLAM_CUDA(ctx, cudaMemsetAsync(d_S_dense, 0, (size_t)n * n * sizeof(double), s));
    VSLAM_CUDA(ctx, cudaEventRecord(e0, s));
    VSLAM_CUDA(ctx, cudaMemsetAsync(b->d_dense, 0, bytes, s));
    ba_dense_fill_kernel<<<ceil_div(P.n_obs, 256), 256, 0, s>>>(P, b->d_dense, ldY);
    VSLAM_LAUNCH_CHECK(ctx, "ba_dense_fill_kernel");
    VSLAM_CUDA(ctx, cudaEventRecord(e1, s));
    const int tiles = ceil_div(n, DS_TM), pairs = tiles * (tiles + 1) / 2, n_chunks = ldY / DS_TK;
    int slices = max(1, (4 * ctx->num_sms) / pairs);  // ~4 CTAs per SM
    slices = min(slices, n_chunks);
    const int cps = ceil_div(n_chunks, slices);
    ba_dense_syrk_kernel<<<dim3(pairs, ceil_div(n_chunks, cps)), 128, 0, s>>>(b->d_dense, ldY, n, tiles, cps, d_S_dense);
    VSLAM_LAUNCH_CHECK(ctx, "ba_dense_syrk_kernel");
    VSLAM_CUDA(ctx, cudaEventRecord(e2, s));
    VSLAM_CUDA(ctx, cudaEventSynchronize(e2));
    float a = 0, c = 0;
    cudaEventElapsedTime(&a, e0, e1);
    cudaEventElapsedTime(&c, e1, e2);
    if (ms_fill) *ms_fill = a;
    if (ms_syrk) *ms_syrk = c;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    return VSLAM_OK;
}

extern "C" int vslam_ba_last_phase_ns(vslam_ctx* ctx, uint64_t* ns8) {
    if (!ctx || !ctx->ba || !ctx->ba->h_sc || !ns8) return VSLAM_E_INVALID;
    for (int i = 0; i < 8; ++i) ns8[i] = ctx->ba->h_sc->phase_ns[i];
    return VSLAM_OK;
}

extern "C" int vslam_ba_multi_last_profile_ns(vslam_ctx* ctx, uint64_t* ns4) {
    if (!ctx || !ctx->ba || !ctx->ba->h_sc || !ns4) return VSLAM_E_INVALID;
    for (int i = 0; i < 4; ++i) ns4[i] = ctx->ba->h_sc->phase_ns[8 + i];
    return VSLAM_OK;
}
