// K12: PnP-RANSAC (cv::solvePnPRansac, sample for sample) + inlier refit, and K7 as a stand-alone call (ANMS on caller
// keypoints).
//
// Replaces cv::solvePnPRansac(pts3d, pts2d, K, noDist, rvec, tvec, false, 100, 4.0, 0.99, inliers) inside
// VO::motion_estimation (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:253-314, call at :277) and the
// reference's own VO::adaptive_non_maximal_suppresion (visual_odometry.cpp:96-157).
//
// Parity definition: the inlier list equals cv2 4.13.0's index for index on any input, because the whole RANSAC is
// reproduced (restated and pinned against live cv2 in oracle/pnp_oracle.c):
//   host                   the sample stream: cv::RNG((uint64)-1) multiply-with-carry, getSubset's duplicate rejection
//                          (it depends only on n, so all `iters` samples are drawn up front)
//   pnp_hypothesis_kernel  one CTA per sample: OpenCV's 5-point EPnP in its exact arithmetic (epnp.cuh) -- the 12x12
//                          Jacobi SVD rotates the independent row pairs of a wavefront on six lanes (bit-identical to
//                          the sequential sweep), the three beta initialisations run on three warps -- then R is
//                          turned into the (rvec, tvec) model and back as the RANSAC callback and projectPoints do
//                          and all threads count the inliers with projectPoints' float32 arithmetic
//                          (err = |uv - (float)proj|^2 <= (float)(thr^2), no cheirality test)
//   pnp_refine_kernel      one CTA: RANSACPointSetRegistrator::run's sequential logic over the counts (a model is
//                          kept iff goodCount > max(best, 4); RANSACUpdateNumIters shrinks the iteration budget, later
//                          samples are ignored exactly as if they had never been drawn), inlier mask of the winner,
//                          then OpenCV's final solvePnP(ITERATIVE) on the inliers = the least-squares optimum of the
//                          reprojection error, reached here by Gauss-Newton in fp64 from the RANSAC model
#include "common.cuh"
#include "epnp.cuh"

#include <math.h>
#include <stdlib.h>

#define PNP_THREADS 256
#define PNP_HYP_THREADS 128
#define PNP_MAX_HYP 512
#define PNP_HYP_STRIDE 16  // doubles per hypothesis record: R (9), t (3), pad

struct PnpState {
    float* d_in;      // [cap*3 xyz | cap*2 uv | PNP_MAX_HYP*5 sample indices]
    double* d_hyp;    // [H][PNP_HYP_STRIDE]
    int* d_cnt;       // [H]
    uint8_t* d_mask;  // [cap]
    int* d_res;       // [PNP_RES_HEAD ints: count, best, executed iterations, pad | 18 doubles | cap inlier indices]
    void* h_in;       // pinned mirrors
    void* h_res;
    int cap;
    // ANMS staging
    vslam_keypoint* d_kp;
    double* d_rad;
    int* d_keep;
};
#define PNP_RES_HEAD 4

struct PnpCam {
    double fx, fy, cx, cy;
};

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
                                              float v, double* H21, double* g) {
    const double px = T[0] * X + T[1] * Y + T[2] * Z + T[3], py = T[4] * X + T[5] * Y + T[6] * Z + T[7],
                 pz = T[8] * X + T[9] * Y + T[10] * Z + T[11];
    const double iz = 1.0 / pz, iz2 = iz * iz;
    const double e0 = (double)u - (cam.fx * px * iz + cam.cx), e1 = (double)v - (cam.fy * py * iz + cam.cy);
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

// PnPRansacCallback::computeError for one correspondence: cv::projectPoints in double, stored as float32, squared
// float32 distance to the measured pixel.  R row-major, no distortion.
__device__ __forceinline__ float pnp_reproj_err(const double* R, const double* t, const PnpCam& cam, const float* xyz,
                                                const float* uv, int i) {
    const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z ? 1. / z : 1;
    x *= z;
    y *= z;
    const float pu = (float)(x * cam.fx + cam.cx), pv = (float)(y * cam.fy + cam.cy);
    const float dx = uv[2 * i] - pu, dy = uv[2 * i + 1] - pv;
    return dx * dx + dy * dy;
}

__global__ void __launch_bounds__(PNP_HYP_THREADS)
pnp_hypothesis_kernel(const float* __restrict__ xyz, const float* __restrict__ uv, int n, PnpCam cam, float thr2,
                      const int* __restrict__ samples, double* __restrict__ hyp, int* __restrict__ cnt) {
    __shared__ EpnpWork work;
    __shared__ double sRb[3][9], stb[3][3], s_err[3];
    __shared__ double sR[9], st[3];
    __shared__ int s_cnt, s_idx[5];
    const int h = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 5) s_idx[tid] = samples[5 * h + tid];
    __syncthreads();
    // phase 1: warp 0 (lane 0 sets the 12x12 system up, six lanes rotate the independent row pairs of a wavefront)
    if (warp == 0) epnp5_prepare(work, xyz, uv, s_idx, cam.fx, cam.fy, cam.cx, cam.cy, lane);
    __syncthreads();
    // phase 2: the three beta initialisations are independent: one warp each (a single lane: sequential fp64 chains)
    if (warp < 3 && lane == 0) s_err[warp] = epnp5_branch(work, warp + 1, sRb[warp], stb[warp]);
    __syncthreads();
    if (tid == 0) {
        const int b = epnp5_select(s_err);
        double R[9], rv[3];
        for (int k = 0; k < 9; ++k) R[k] = sRb[b][k];
        rodrigues_to_vec_dev(R, rv);  // the model RANSAC carries is (rvec, tvec) ...
        rodrigues_to_mat_dev(rv, R);  // ... and projectPoints turns it back into a matrix
        for (int k = 0; k < 9; ++k) sR[k] = R[k];
        for (int k = 0; k < 3; ++k) st[k] = stb[b][k];
        s_cnt = 0;
    }
    __syncthreads();
    int c = 0;
    for (int i = tid; i < n; i += PNP_HYP_THREADS) c += pnp_reproj_err(sR, st, cam, xyz, uv, i) <= thr2 ? 1 : 0;
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if (lane == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (tid == 0) {
        cnt[h] = s_cnt;
        for (int k = 0; k < 9; ++k) hyp[PNP_HYP_STRIDE * h + k] = sR[k];
        for (int k = 0; k < 3; ++k) hyp[PNP_HYP_STRIDE * h + 9 + k] = st[k];
    }
}

// cv::RANSACUpdateNumIters
__device__ int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmin(fmax(p, 0.), 1.);
    ep = fmin(fmax(ep, 0.), 1.);
    double num = fmax(1. - p, DBL_MIN);
    double denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : __double2int_rn(num / denom);
}

// direct != 0: n == 5, OpenCV skips RANSAC and returns the EPnP pose of the five points with all of them as inliers
__global__ void __launch_bounds__(PNP_THREADS)
pnp_refine_kernel(const float* __restrict__ xyz, const float* __restrict__ uv, int n, PnpCam cam, float thr2, int n_hyp,
                  double confidence, int direct, const double* __restrict__ hyp, const int* __restrict__ cnt, int max_iter,
                  uint8_t* __restrict__ mask, int* __restrict__ res) {
    __shared__ double sT[12], sR[9], st[3];
    __shared__ double sH[PNP_THREADS / 32][28];
    __shared__ int s_best, s_stop, s_base, s_warp[PNP_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* out = reinterpret_cast<double*>(res + PNP_RES_HEAD);
    int* inl = res + PNP_RES_HEAD + 36;
    if (tid == 0) {
        int best = -1, executed = 0;
        if (direct) {
            best = 0;
            executed = 1;
        } else {  // RANSACPointSetRegistrator::run over the precomputed samples
            int niters = n_hyp > 1 ? n_hyp : 1, max_good = 0;
            for (int it = 0; it < niters && it < n_hyp; ++it) {
                const int good = cnt[it];
                if (good > (max_good > 4 ? max_good : 4)) {
                    best = it;
                    max_good = good;
                    niters = ransac_update_num_iters(confidence, (double)(n - good) / n, 5, niters);
                }
                executed = it + 1;
            }
        }
        s_best = best;
        if (best >= 0) {
            for (int i = 0; i < 9; ++i) sR[i] = hyp[PNP_HYP_STRIDE * best + i];
            for (int i = 0; i < 3; ++i) st[i] = hyp[PNP_HYP_STRIDE * best + 9 + i];
            for (int r = 0; r < 3; ++r) {
                sT[r * 4] = sR[r * 3]; sT[r * 4 + 1] = sR[r * 3 + 1]; sT[r * 4 + 2] = sR[r * 3 + 2]; sT[r * 4 + 3] = st[r];
            }
        }
        res[1] = best;
        res[2] = executed;
        s_stop = 0;
        s_base = 0;
    }
    __syncthreads();
    if (s_best < 0) {
        if (tid == 0) res[0] = 0;
        return;
    }
    // inlier mask of the winning model + ascending index list
    for (int i0 = 0; i0 < n; i0 += PNP_THREADS) {
        const int i = i0 + tid;
        const bool k = i < n && (direct || pnp_reproj_err(sR, st, cam, xyz, uv, i) <= thr2);
        if (i < n) mask[i] = k ? 1 : 0;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, k);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < PNP_THREADS / 32; ++w) {
            before += w < warp ? s_warp[w] : 0;
            total += s_warp[w];
        }
        if (k) inl[s_base + before + __popc(bal & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    // refit on the inliers (skipped when OpenCV returns the EPnP pose itself)
    for (int it = 0; it < max_iter && !direct; ++it) {
        double H21[21], g[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) H21[i] = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) g[i] = 0;
        double T[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) T[i] = sT[i];
        for (int i = tid; i < n; i += PNP_THREADS) {
            if (!mask[i]) continue;
            gn_accumulate(T, cam, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], uv[2 * i], uv[2 * i + 1], H21, g);
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
        res[0] = s_base;
        double R[9] = {sT[0], sT[1], sT[2], sT[4], sT[5], sT[6], sT[8], sT[9], sT[10]}, rv[3];
        rodrigues_to_vec_dev(R, rv);
        out[0] = rv[0]; out[1] = rv[1]; out[2] = rv[2];
        out[3] = sT[3]; out[4] = sT[7]; out[5] = sT[11];
        for (int i = 0; i < 12; ++i) out[6 + i] = sT[i];
    }
}

// ----------------------------------------------------------------------------------------------------------------
// K7 stand-alone: ANMS over a caller-supplied keypoint list (any order), one CTA.
// ----------------------------------------------------------------------------------------------------------------
// radius scan spread over the grid: one warp per keypoint i, lanes stride over all keypoints j with
// response_j > response_i * c (any input order); min of the squared distances, one sqrt at the end (sqrt is monotone
// and correctly rounded, so this equals the reference's min over sqrt bit for bit)
__global__ void __launch_bounds__(256)
anms_points_radius_kernel(const vslam_keypoint* __restrict__ kp, int n, float c_robust, double* __restrict__ rad) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        const float thr = __fmul_rn(kp[i].response, c_robust);
        const float xi = kp[i].x, yi = kp[i].y;
        double best2 = 1.7976931348623157e308;
        for (int j = lane; j < n; j += 32) {
            if (!(kp[j].response > thr)) continue;
            const float dx = __fsub_rn(xi, kp[j].x), dy = __fsub_rn(yi, kp[j].y);
            best2 = fmin(best2, __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best2 = fmin(best2, __shfl_xor_sync(0xFFFFFFFFu, best2, o));
        if (lane == 0) rad[i] = best2 < 1.7976931348623157e308 ? sqrt(best2) : best2;
    }
}

__global__ void __launch_bounds__(1024)
anms_points_kernel(const vslam_keypoint* __restrict__ kp, int n, int num, float c_robust, double* __restrict__ rad,
                   int* __restrict__ keep) {
    extern __shared__ unsigned long long s_sort[];
    __shared__ int s_base, s_warp[32];
    const int tid = threadIdx.x;
    (void)kp;
    (void)c_robust;
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

static size_t pnp_in_bytes(int cap) { return (size_t)cap * 20 + PNP_MAX_HYP * 5 * sizeof(int); }
static size_t pnp_res_bytes(int cap) { return (PNP_RES_HEAD + 36 + (size_t)cap) * sizeof(int); }

int vslam_pnp_init(vslam_ctx* ctx) {
    PnpState* p = (PnpState*)calloc(1, sizeof(PnpState));
    if (!p) return VSLAM_E_INVALID;
    ctx->pnp = p;
    p->cap = ctx->cfg.max_keypoints > 0 ? ctx->cfg.max_keypoints : 1;
    const size_t cap = (size_t)p->cap;
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_in, pnp_in_bytes(p->cap)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_hyp, PNP_MAX_HYP * PNP_HYP_STRIDE * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_cnt, PNP_MAX_HYP * sizeof(int)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_mask, cap));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_res, pnp_res_bytes(p->cap)));
    VSLAM_CUDA(ctx, cudaMallocHost(&p->h_in, pnp_in_bytes(p->cap)));
    VSLAM_CUDA(ctx, cudaMallocHost(&p->h_res, pnp_res_bytes(p->cap)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_kp, cap * sizeof(vslam_keypoint)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_rad, cap * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&p->d_keep, (cap + 1) * sizeof(int)));
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(anms_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    return VSLAM_OK;
}

void vslam_pnp_free(vslam_ctx* ctx) {
    PnpState* p = ctx->pnp;
    if (!p) return;
    cudaFree(p->d_in); cudaFree(p->d_hyp); cudaFree(p->d_cnt); cudaFree(p->d_mask); cudaFree(p->d_res);
    cudaFreeHost(p->h_in); cudaFreeHost(p->h_res);
    cudaFree(p->d_kp); cudaFree(p->d_rad); cudaFree(p->d_keep);
    free(p);
    ctx->pnp = nullptr;
}

// cv::RNG as RANSACPointSetRegistrator::run seeds it, and getSubset's draw of 5 distinct indices per iteration.  The
// stream does not depend on the data, so every sample the loop could reach is drawn before the kernels start.
static void pnp_draw_samples(int n, int iters, int* out) {
    uint64_t state = (uint64_t)-1;
    for (int it = 0; it < iters; ++it) {
        int* idx = out + 5 * it;
        for (int i = 0; i < 5; ++i) {
            for (;;) {
                state = (uint64_t)(uint32_t)state * 4164903690ull + (uint32_t)(state >> 32);
                const int c = (int)((uint32_t)state % (uint32_t)n);
                bool dup = false;
                for (int q = 0; q < i; ++q) dup |= idx[q] == c;
                if (!dup) { idx[i] = c; break; }
            }
        }
    }
}

extern "C" int vslam_pnp_ransac(vslam_ctx* ctx, const float* xyz, const float* uv, int n, const double* Kmat, int iters,
                                float reproj_err, double confidence, double* rvec, double* tvec, double* T_c_w,
                                int32_t* inliers, int32_t* n_inliers) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !n_inliers || !Kmat || n < 0) return VSLAM_E_INVALID;
    *n_inliers = 0;
    // OpenCV asserts n >= 4 and switches to a P3P kernel for exactly 4 points; the reference rejects any frame with
    // fewer than 10 inliers (visual_odometry.cpp:316-346), so both cases end as "no inliers" here
    if (n < 5) return VSLAM_OK;
    if (!xyz || !uv || !rvec || !tvec || !inliers) return VSLAM_E_INVALID;
    if (!(confidence > 0 && confidence < 1)) return VSLAM_E_INVALID;  // CV_Assert in RANSACPointSetRegistrator::run
    PnpState* p = pnp_state(ctx);
    if (!p) return VSLAM_E_CAPACITY;
    if (n > p->cap) return VSLAM_E_CAPACITY;
    if (iters <= 0) iters = 1;  // niters = MAX(maxIters, 1)
    if (iters > PNP_MAX_HYP) return VSLAM_E_CAPACITY;
    const int direct = n == 5;
    const int n_hyp = direct ? 1 : iters;
    PnpCam cam = {Kmat[0], Kmat[4], Kmat[2], Kmat[5]};
    const float thr2 = (float)((double)reproj_err * (double)reproj_err);
    cudaStream_t s = ctx->stream;
    // one upload: [xyz | uv | samples]
    float* h_in = (float*)p->h_in;
    memcpy(h_in, xyz, (size_t)n * 12);
    memcpy(h_in + (size_t)n * 3, uv, (size_t)n * 8);
    int* h_samples = (int*)(h_in + (size_t)n * 5);
    if (direct)
        for (int i = 0; i < 5; ++i) h_samples[i] = i;
    else
        pnp_draw_samples(n, n_hyp, h_samples);
    const size_t in_bytes = (size_t)n * 20 + (size_t)n_hyp * 5 * sizeof(int);
    VSLAM_CUDA(ctx, cudaMemcpyAsync(p->d_in, h_in, in_bytes, cudaMemcpyHostToDevice, s));
    const float* d_xyz = p->d_in;
    const float* d_uv = p->d_in + (size_t)n * 3;
    const int* d_samples = (const int*)(p->d_in + (size_t)n * 5);
    vslam_time_begin(ctx, VK_PNP);
    pnp_hypothesis_kernel<<<n_hyp, PNP_HYP_THREADS, 0, s>>>(d_xyz, d_uv, n, cam, thr2, d_samples, p->d_hyp, p->d_cnt);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "pnp_hypothesis_kernel");
    vslam_time_begin(ctx, VK_PNP_REFINE);
    pnp_refine_kernel<<<1, PNP_THREADS, 0, s>>>(d_xyz, d_uv, n, cam, thr2, n_hyp, confidence, direct, p->d_hyp, p->d_cnt, 30,
                                                p->d_mask, p->d_res);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "pnp_refine_kernel");
    // one download: count, pose and the index list (entries past the count are unspecified, as documented)
    VSLAM_CUDA(ctx, cudaMemcpyAsync(p->h_res, p->d_res, pnp_res_bytes(n), cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    const int* h_res = (const int*)p->h_res;
    const int cnt = h_res[0];
    *n_inliers = cnt;
    if (cnt <= 0) return VSLAM_OK;
    const double* out = (const double*)(h_res + PNP_RES_HEAD);
    memcpy(inliers, h_res + PNP_RES_HEAD + 36, (size_t)cnt * sizeof(int));
    for (int i = 0; i < 3; ++i) {
        rvec[i] = out[i];
        tvec[i] = out[3 + i];
    }
    if (T_c_w)
        for (int i = 0; i < 12; ++i) T_c_w[i] = out[6 + i];
    return VSLAM_OK;
}

/* test tap: model (R row-major 9, t 3) and inlier count of RANSAC sample `it` of the last vslam_pnp_ransac call, and
 * the number of iterations OpenCV's loop would have executed */
extern "C" int vslam_pnp_debug_read(vslam_ctx* ctx, int it, double* R_t12, int32_t* count, int32_t* executed) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !ctx->pnp || it < 0 || it >= PNP_MAX_HYP) return VSLAM_E_INVALID;
    PnpState* p = ctx->pnp;
    VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (R_t12) VSLAM_CUDA(ctx, cudaMemcpy(R_t12, p->d_hyp + (size_t)PNP_HYP_STRIDE * it, 12 * sizeof(double), cudaMemcpyDeviceToHost));
    if (count) VSLAM_CUDA(ctx, cudaMemcpy(count, p->d_cnt + it, sizeof(int), cudaMemcpyDeviceToHost));
    if (executed) *executed = ((const int*)p->h_res)[2];
    return VSLAM_OK;
}

extern "C" int vslam_anms(vslam_ctx* ctx, const vslam_keypoint* keypoints, int n, int num, float c_robust,
                          int32_t* keep_idx, int32_t* n_keep) {
    VslamDeviceGuard device_guard__(ctx);
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
    anms_points_radius_kernel<<<min(ceil_div(n, 8), 4 * ctx->num_sms), 256, 0, s>>>(p->d_kp, n, c_robust, p->d_rad);
    VSLAM_LAUNCH_CHECK(ctx, "anms_points_radius_kernel");
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
