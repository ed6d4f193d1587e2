// K11 (per-match DLT triangulation + depth gates) and the batched stereo frontend
// (ORB on left and right images -> L<->R Hamming matching -> triangulation), all device-resident.
//
// Role in the reference: VO::disparity_map + Frame::find_3d + VO::set_ref_3d_position
// (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:159-217, types_def.cpp:9-18) turn the keypoints of a
// frame into world points with the gates 10 < Z < 400 (usable) and Z < 40 (reliable depth).  The reference gets Z from
// a dense SGBM disparity; the north star replaces that with sparse stereo: ORB on both images, the reference's own
// feature_matching (cross-check + distance gate, visual_odometry.cpp:219-251) between left and right descriptors and a
// per-match DLT whose arithmetic oracle is cv::triangulatePoints (SVD of the 4x4 system; one-sided Jacobi here, the
// scheme OpenCV's JacobiSVD uses).  fp64 throughout; world = T_c_w^-1 * p narrowed to float32 like cv::Point3f.
#include "common.cuh"

#include <stdlib.h>

#define FRONT_CHUNK_PAIRS 16   // pairs per pipeline chunk (two compute streams: 12 -> 17.7k, 16 -> 18.4k, 24 -> 18.2k, 32 -> 18.0k e2e fps)
#define FRONT_MAX_CHUNKS 64

struct FrontState {
    // host-buffer staging for vslam_stereo_frontend_batch / vslam_triangulate
    uint8_t* d_img;  // [2 * pairs][h][pitch] left images first, then right
    int pitch;
    uint8_t* d_raw;  // host-buffer batches whose row pitch is not a multiple of 16 land here (1-D DMAs) and are re-pitched
    size_t raw_cap;  //   into d_img on the device, so that every kernel reads level 0 with aligned 128-bit loads
    vslam_keypoint* d_kp;
    uint8_t* d_desc;
    int32_t* d_nkp;
    vslam_dmatch* d_match;
    int32_t* d_nmatch;
    float* d_xyz;
    uint8_t* d_flags;
    double* d_cam;   // P1 (12) P2 (12)
    double* d_pose;  // [pairs][12]
    float* d_xl;     // vslam_triangulate staging
    float* d_xr;
    int max_pairs, kp_cap;
    // chunked host-buffer pipeline: H2D of chunk c+1 and D2H of chunk c-1 overlap the kernels of chunk c
    cudaStream_t s_in, s_out;
    cudaStream_t s_alt;  // second compute stream: odd chunks run here so that one chunk's launch tails overlap the next
    cudaEvent_t ev_join;
    cudaEvent_t ev_in[FRONT_MAX_CHUNKS], ev_done[FRONT_MAX_CHUNKS], ev_start;
    // a batch enqueued by vslam_stereo_frontend_batch_begin and not yet collected by _end
    int inflight_pairs, inflight_chunks;
    int inflight_c_start[FRONT_MAX_CHUNKS + 1];
};

struct Cam24 {
    double p[24];
};

// One-sided (Hestenes) Jacobi SVD of the 4x4 DLT matrix; returns the right singular vector of the smallest
// singular value.  Fully unrolled so A and V stay in registers.
__device__ __forceinline__ void dlt_null_vector(double A[4][4], double X[4]) {
    double V[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool changed = false;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = i + 1; j < 4; ++j) {
                double a = 0, b = 0, p = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    a += A[k][i] * A[k][i];
                    b += A[k][j] * A[k][j];
                    p += A[k][i] * A[k][j];
                }
                if (fabs(p) <= 2.220446049250313e-16 * sqrt(a * b)) continue;
                changed = true;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double t0 = c * A[k][i] + s * A[k][j], t1 = -s * A[k][i] + c * A[k][j];
                    A[k][i] = t0;
                    A[k][j] = t1;
                    const double v0 = c * V[k][i] + s * V[k][j], v1 = -s * V[k][i] + c * V[k][j];
                    V[k][i] = v0;
                    V[k][j] = v1;
                }
            }
        }
        if (!changed) break;
    }
    double best = 1e300;
    int bi = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double n = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) n += A[k][j] * A[k][j];
        if (n < best) {
            best = n;
            bi = j;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) X[k] = bi == 0 ? V[k][0] : bi == 1 ? V[k][1] : bi == 2 ? V[k][2] : V[k][3];
}

__device__ __forceinline__ void triangulate_one(float xlx, float xly, float xrx, float xry, const double* P1,
                                                const double* P2, const double* T, float* xyz, uint8_t* flags) {
    double A[4][4];
    const double xl = xlx, yl = xly, xr = xrx, yr = xry;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        A[0][k] = xl * P1[8 + k] - P1[k];
        A[1][k] = yl * P1[8 + k] - P1[4 + k];
        A[2][k] = xr * P2[8 + k] - P2[k];
        A[3][k] = yr * P2[8 + k] - P2[4 + k];
    }
    double X[4];
    dlt_null_vector(A, X);
    const double iw = 1.0 / X[3];
    const double px = X[0] * iw, py = X[1] * iw, pz = X[2] * iw;
    // set_ref_3d_position gates (visual_odometry.cpp:194,201)
    const bool usable = pz > 10.0 && pz < 400.0;
    const bool reliable = usable && pz < 40.0;
    // world = T_c_w^-1 * p = R^T (p - t)   (types_def.cpp:17)
    double wx = px, wy = py, wz = pz;
    if (T) {
        const double dx = px - T[3], dy = py - T[7], dz = pz - T[11];
        wx = T[0] * dx + T[4] * dy + T[8] * dz;
        wy = T[1] * dx + T[5] * dy + T[9] * dz;
        wz = T[2] * dx + T[6] * dy + T[10] * dz;
    }
    xyz[0] = (float)wx;
    xyz[1] = (float)wy;
    xyz[2] = (float)wz;
    *flags = (uint8_t)((usable ? 1 : 0) | (reliable ? 2 : 0));
}

__global__ void __launch_bounds__(128)
triangulate_matches_kernel(const vslam_keypoint* __restrict__ kp_left, const vslam_keypoint* __restrict__ kp_right,
                           int kp_stride, const vslam_dmatch* __restrict__ matches,
                           const int32_t* __restrict__ n_matches, int match_stride, const Cam24 cam,
                           const double* __restrict__ poses, float* __restrict__ xyz, uint8_t* __restrict__ flags) {
    const int pair = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_matches[pair]) return;
    const vslam_dmatch m = matches[(size_t)pair * match_stride + i];
    const vslam_keypoint* kl = kp_left + (size_t)pair * kp_stride + m.queryIdx;
    const vslam_keypoint* kr = kp_right + (size_t)pair * kp_stride + m.trainIdx;
    triangulate_one(kl->x, kl->y, kr->x, kr->y, cam.p, cam.p + 12, poses ? poses + 12 * pair : nullptr,
                    xyz + ((size_t)pair * match_stride + i) * 3, flags + (size_t)pair * match_stride + i);
}

__global__ void __launch_bounds__(128)
triangulate_points_kernel(const float* __restrict__ xl, const float* __restrict__ xr, int n, const Cam24 cam,
                          const double* __restrict__ pose, float* __restrict__ xyz, uint8_t* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    triangulate_one(xl[2 * i], xl[2 * i + 1], xr[2 * i], xr[2 * i + 1], cam.p, cam.p + 12, pose, xyz + 3 * i, flags + i);
}

// Copy n_img images of w x h bytes from an arbitrary row pitch to a 16-byte aligned one: 16 output bytes per thread
// from five aligned source words.  HBM-trivial (2 x 467 KB per image) next to the 2-D DMA it replaces.
__global__ void __launch_bounds__(128)
repitch_kernel(const uint8_t* __restrict__ src, int spitch, long long sstride, uint8_t* __restrict__ dst, int dpitch,
               long long dstride) {
    const int x = (blockIdx.x * 128 + threadIdx.x) * 16;
    if (x >= dpitch) return;
    const uint8_t* p = src + blockIdx.z * sstride + (long long)blockIdx.y * spitch + x;
    const uint32_t a = (uint32_t)(uintptr_t)p & 3u;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3), w4 = a ? __ldg(q + 4) : 0u;
    const uint32_t sh = 8 * a;
    *reinterpret_cast<uint4*>(dst + blockIdx.z * dstride + (long long)blockIdx.y * dpitch + x) =
        make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                   __funnelshift_r(w3, w4, sh));
}

int vslam_front_init(vslam_ctx* ctx) {
    FrontState* f = (FrontState*)calloc(1, sizeof(FrontState));
    if (!f) return VSLAM_E_INVALID;
    ctx->front = f;
    const vslam_config& c = ctx->cfg;
    f->max_pairs = c.max_images / 2;
    f->kp_cap = c.max_keypoints;
    if (c.max_images <= 0 || c.max_width <= 0 || c.max_height <= 0 || c.max_keypoints <= 0) return VSLAM_OK;
    const size_t ni = (size_t)c.max_images, np = (size_t)(f->max_pairs > 0 ? f->max_pairs : 1), cap = (size_t)f->kp_cap;
    f->pitch = (c.max_width + 15) & ~15;
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_img, ni * f->pitch * c.max_height));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_kp, ni * cap * sizeof(vslam_keypoint)));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_desc, ni * cap * 32));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_nkp, ni * sizeof(int32_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_match, np * cap * sizeof(vslam_dmatch)));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_nmatch, np * sizeof(int32_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_xyz, np * cap * 3 * sizeof(float)));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_flags, np * cap));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_pose, np * 12 * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_xl, cap * 2 * sizeof(float)));
    VSLAM_CUDA(ctx, cudaMalloc(&f->d_xr, cap * 2 * sizeof(float)));
    VSLAM_CUDA(ctx, cudaStreamCreateWithFlags(&f->s_in, cudaStreamNonBlocking));
    VSLAM_CUDA(ctx, cudaStreamCreateWithFlags(&f->s_out, cudaStreamNonBlocking));
    VSLAM_CUDA(ctx, cudaStreamCreateWithFlags(&f->s_alt, cudaStreamNonBlocking));
    VSLAM_CUDA(ctx, cudaEventCreateWithFlags(&f->ev_join, cudaEventDisableTiming));
    VSLAM_CUDA(ctx, cudaEventCreate(&f->ev_start));
    for (int i = 0; i < FRONT_MAX_CHUNKS; ++i) {
        VSLAM_CUDA(ctx, cudaEventCreate(&f->ev_in[i]));
        VSLAM_CUDA(ctx, cudaEventCreate(&f->ev_done[i]));
    }
    return VSLAM_OK;
}

void vslam_front_free(vslam_ctx* ctx) {
    FrontState* f = ctx->front;
    if (!f) return;
    cudaFree(f->d_img);
    cudaFree(f->d_raw);
    cudaFree(f->d_kp);
    cudaFree(f->d_desc);
    cudaFree(f->d_nkp);
    cudaFree(f->d_match);
    cudaFree(f->d_nmatch);
    cudaFree(f->d_xyz);
    cudaFree(f->d_flags);
    cudaFree(f->d_pose);
    cudaFree(f->d_xl);
    cudaFree(f->d_xr);
    if (f->s_in) cudaStreamDestroy(f->s_in);
    if (f->s_out) cudaStreamDestroy(f->s_out);
    if (f->s_alt) cudaStreamDestroy(f->s_alt);
    if (f->ev_join) cudaEventDestroy(f->ev_join);
    if (f->ev_start) cudaEventDestroy(f->ev_start);
    for (int i = 0; i < FRONT_MAX_CHUNKS; ++i) {
        if (f->ev_in[i]) cudaEventDestroy(f->ev_in[i]);
        if (f->ev_done[i]) cudaEventDestroy(f->ev_done[i]);
    }
    free(f);
    ctx->front = nullptr;
}

extern "C" int vslam_triangulate(vslam_ctx* ctx, const float* xl, const float* xr, int n, const double* P1,
                                 const double* P2, const double* T_c_w, float* xyz_world, uint8_t* flags) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || n < 0 || !P1 || !P2) return VSLAM_E_INVALID;
    if (n == 0) return VSLAM_OK;
    if (!xl || !xr || !xyz_world || !flags) return VSLAM_E_INVALID;
    FrontState* f = ctx->front;
    if (!f || !f->d_xl || n > f->kp_cap) return VSLAM_E_CAPACITY;
    cudaStream_t s = ctx->stream;
    Cam24 cam;
    memcpy(cam.p, P1, 96);
    memcpy(cam.p + 12, P2, 96);
    VSLAM_CUDA(ctx, cudaMemcpyAsync(f->d_xl, xl, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(f->d_xr, xr, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    if (T_c_w) VSLAM_CUDA(ctx, cudaMemcpyAsync(f->d_pose, T_c_w, 96, cudaMemcpyHostToDevice, s));
    vslam_time_begin(ctx, VK_TRIANGULATE);
    triangulate_points_kernel<<<ceil_div(n, 128), 128, 0, s>>>(f->d_xl, f->d_xr, n, cam, T_c_w ? f->d_pose : nullptr,
                                                              f->d_xyz, f->d_flags);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "triangulate_points_kernel");
    VSLAM_CUDA(ctx, cudaMemcpyAsync(xyz_world, f->d_xyz, (size_t)n * 12, cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(flags, f->d_flags, (size_t)n, cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    return VSLAM_OK;
}

extern "C" int vslam_triangulate_matches_batch_dev(vslam_ctx* ctx, const vslam_keypoint* d_kp_left,
                                                   const vslam_keypoint* d_kp_right, int kp_stride,
                                                   const vslam_dmatch* d_matches, const int32_t* d_n_matches,
                                                   int match_stride, int batch, const double* P1, const double* P2,
                                                   const double* d_T_c_w, float* d_xyz, uint8_t* d_flags) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !d_kp_left || !d_kp_right || !d_matches || !d_n_matches || !P1 || !P2 || !d_xyz || !d_flags)
        return VSLAM_E_INVALID;
    if (batch <= 0 || match_stride <= 0) return VSLAM_E_INVALID;
    Cam24 cam;
    memcpy(cam.p, P1, 96);
    memcpy(cam.p + 12, P2, 96);
    dim3 grid(ceil_div(match_stride, 128), batch);
    vslam_time_begin(ctx, VK_TRIANGULATE);
    triangulate_matches_kernel<<<grid, 128, 0, ctx->stream>>>(d_kp_left, d_kp_right, kp_stride, d_matches, d_n_matches,
                                                             match_stride, cam, d_T_c_w, d_xyz, d_flags);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "triangulate_matches_kernel");
    return VSLAM_OK;
}

// ORB(left) + ORB(right) + feature_matching(left -> right) + triangulation for pairs [p0, p0 + n_chunk) of a batch of
// n_pairs, everything enqueued on the context stream.  d_left / d_right point at the chunk's first image; the output
// arrays are the whole batch's (left images occupy slots [0, n_pairs), right images [n_pairs, 2 n_pairs)).
static int front_enqueue_chunk(vslam_ctx* ctx, const uint8_t* d_left, const uint8_t* d_right, int n_pairs, int p0,
                               int n_chunk, int width, int height, int row_pitch, long long image_stride,
                               int nfeatures, int anms_keep, float anms_c, double gate_rel, double gate_abs,
                               const double* P1, const double* P2, const double* d_T_c_w, vslam_keypoint* d_kp,
                               uint8_t* d_desc, int32_t* d_n_kp, vslam_dmatch* d_matches, int32_t* d_n_matches,
                               float* d_xyz, uint8_t* d_flags, int scratch_pair0 = 0) {
    const size_t cap = (size_t)ctx->cfg.max_keypoints;
    ImgSrc src;
    src.base[0] = d_left;
    src.base[1] = d_right;
    src.img_stride = image_stride;
    src.pitch = row_pitch;
    src.per_base = n_chunk;
    src.out_slot[0] = p0;
    src.out_slot[1] = n_pairs + p0;
    int st = vslam_orb_enqueue(ctx, src, 2 * n_chunk, width, height, nfeatures, anms_keep, anms_c, d_kp, d_desc, d_n_kp,
                               2 * scratch_pair0);
    if (st != VSLAM_OK) return st;
    // query = left descriptors (slots p0..), train = right descriptors (slots n_pairs + p0..)
    const size_t lq = (size_t)p0, rq = (size_t)n_pairs + p0;
    st = vslam_match_enqueue(ctx, d_desc + lq * cap * 32, d_n_kp + lq, (int)cap, d_desc + rq * cap * 32, d_n_kp + rq,
                             (int)cap, n_chunk, (int)cap, 1, gate_rel, gate_abs, d_matches + lq * cap, (int)cap,
                             d_n_matches + lq, scratch_pair0);
    if (st != VSLAM_OK) return st;
    return vslam_triangulate_matches_batch_dev(ctx, d_kp + lq * cap, d_kp + rq * cap, (int)cap, d_matches + lq * cap,
                                               d_n_matches + lq, (int)cap, n_chunk, P1, P2,
                                               d_T_c_w ? d_T_c_w + 12 * lq : nullptr, d_xyz + lq * cap * 3,
                                               d_flags + lq * cap);
}

extern "C" int vslam_stereo_frontend_batch_dev(vslam_ctx* ctx, const uint8_t* d_left, const uint8_t* d_right,
                                               int n_pairs, int width, int height, int row_pitch,
                                               long long image_stride, int nfeatures, int anms_keep, float anms_c,
                                               double gate_rel, double gate_abs, const double* P1, const double* P2,
                                               const double* d_T_c_w, vslam_keypoint* d_kp, uint8_t* d_desc,
                                               int32_t* d_n_kp, vslam_dmatch* d_matches, int32_t* d_n_matches,
                                               float* d_xyz, uint8_t* d_flags) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !d_left || !d_right || !P1 || !P2 || !d_kp || !d_desc || !d_n_kp || !d_matches || !d_n_matches ||
        !d_xyz || !d_flags)
        return VSLAM_E_INVALID;
    if (n_pairs <= 0 || width <= 0 || height <= 0 || row_pitch < width) return VSLAM_E_INVALID;
    if (2 * n_pairs > ctx->cfg.max_images) return VSLAM_E_CAPACITY;
    FrontState* f = ctx->front;
    // large batches are cut into chunks that alternate between the context stream and a second stream (disjoint scratch
    // slots): every kernel of the pipeline ends in a tail of a few long CTAs, and the other stream's kernels fill it
    static const int split_env = getenv("VSLAM_FRONT_SPLIT") ? atoi(getenv("VSLAM_FRONT_SPLIT")) : 0;
    int n_chunks = split_env > 0 ? split_env : (n_pairs >= 64 ? 2 : 1);  // measured at 128 pairs: 1 -> 20.99k, 2 -> 21.5k, 4 -> 20.8k fps
    if (!f || !f->s_alt || n_chunks > n_pairs || (ctx->serial && split_env <= 0)) n_chunks = 1;
    if (n_chunks == 1)
        return front_enqueue_chunk(ctx, d_left, d_right, n_pairs, 0, n_pairs, width, height, row_pitch, image_stride,
                                   nfeatures, anms_keep, anms_c, gate_rel, gate_abs, P1, P2, d_T_c_w, d_kp, d_desc,
                                   d_n_kp, d_matches, d_n_matches, d_xyz, d_flags);
    const int chunk = ceil_div(n_pairs, n_chunks);
    int st = vslam_orb_prepare(ctx, width, height);
    if (st != VSLAM_OK) return st;
    cudaStream_t s = ctx->stream;
    VSLAM_CUDA(ctx, cudaEventRecord(f->ev_start, s));
    VSLAM_CUDA(ctx, cudaStreamWaitEvent(f->s_alt, f->ev_start, 0));
    for (int c = 0, p0 = 0; p0 < n_pairs; ++c, p0 += chunk) {
        const int nc = n_pairs - p0 < chunk ? n_pairs - p0 : chunk;
        const bool alt = c & 1;
        ctx->stream = alt ? f->s_alt : s;
        st = front_enqueue_chunk(ctx, d_left + (long long)p0 * image_stride, d_right + (long long)p0 * image_stride, n_pairs, p0,
                                 nc, width, height, row_pitch, image_stride, nfeatures, anms_keep, anms_c, gate_rel,
                                 gate_abs, P1, P2, d_T_c_w, d_kp, d_desc, d_n_kp, d_matches, d_n_matches, d_xyz, d_flags,
                                 alt ? chunk : 0);
        ctx->stream = s;
        if (st != VSLAM_OK) break;
    }
    cudaEventRecord(f->ev_join, f->s_alt);  // join even on error: the context stream must cover everything enqueued
    cudaStreamWaitEvent(s, f->ev_join, 0);
    return st;
}

// Host-buffer form (the call a user of the library makes): images come from host memory (pinned memory makes the
// copies asynchronous DMA), results are copied back, one synchronisation at the end.  Outputs use the same strides as
// the device form (cap = vslam_orb_keypoint_capacity(ctx)): kp [2*n_pairs][cap], desc [2*n_pairs][cap][32],
// n_kp [2*n_pairs], matches [n_pairs][cap], n_matches [n_pairs], xyz [n_pairs][cap][3], flags [n_pairs][cap].
//
// _begin enqueues the whole batch (uploads, kernels, downloads) and returns; _end waits for it and reports the overflow
// flags.  A caller that alternates two contexts keeps two batches in flight: the first upload and the last download of
// one batch hide behind the other batch's kernels (bench.py's e2e loop; inputs and outputs must stay valid, and pinned
// for the copies to be asynchronous, until _end).  vslam_stereo_frontend_batch = _begin + _end.
extern "C" int vslam_stereo_frontend_batch_begin(vslam_ctx* ctx, const uint8_t* left, const uint8_t* right, int n_pairs,
                                                 int width, int height, int row_pitch, long long image_stride,
                                                 int nfeatures, int anms_keep, float anms_c, double gate_rel,
                                                 double gate_abs, const double* P1, const double* P2, const double* T_c_w,
                                                 vslam_keypoint* kp, uint8_t* desc, int32_t* n_kp, vslam_dmatch* matches,
                                                 int32_t* n_matches, float* xyz, uint8_t* flags) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !n_kp || !n_matches) return VSLAM_E_INVALID;
    if (ctx->front && ctx->front->inflight_pairs > 0) return VSLAM_E_INVALID;  // one batch per context at a time
    if (!left || !right) return VSLAM_E_INVALID;  // reference: -1 "Could not open or find the image"
    if (!P1 || !P2 || !kp || !desc || !matches || !xyz || !flags) return VSLAM_E_INVALID;
    FrontState* f = ctx->front;
    if (!f || !f->d_img) return VSLAM_E_CAPACITY;
    if (n_pairs <= 0 || n_pairs > f->max_pairs) return VSLAM_E_CAPACITY;
    if (width <= 0 || height <= 0 || row_pitch < width) return VSLAM_E_INVALID;
    if (width > ctx->cfg.max_width || height > ctx->cfg.max_height) return VSLAM_E_CAPACITY;
    cudaStream_t s = ctx->stream;
    // staging keeps the caller's row pitch when the batch is one contiguous block: 1-D DMAs instead of 2-D copies of
    // 1241-byte rows (the kernels read bytes, so the pitch need not be aligned)
    static const bool align_env = getenv("VSLAM_FRONT_ALIGN") != nullptr;  // experiment: always stage into 16-byte aligned rows
    const bool contiguous = image_stride == (long long)row_pitch * height && row_pitch <= f->pitch && !(align_env && (row_pitch & 15));
    // an unaligned pitch (1241-byte KITTI rows) still goes up as 1-D DMAs, into d_raw, and a device kernel re-pitches
    // each chunk into d_img: level 0 then takes the aligned 128-bit paths of the FAST / blur / resize kernels
    static const bool no_repitch = getenv("VSLAM_FRONT_NO_REPITCH") != nullptr;
    const bool repitch = contiguous && (row_pitch & 15) && !no_repitch;
    if (repitch) {
        const size_t need = 2 * (size_t)n_pairs * image_stride + 64;  // + slack for the last aligned window
        if (need > f->raw_cap) {
            VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
            cudaFree(f->d_raw);
            f->d_raw = nullptr;
            f->raw_cap = 0;
            VSLAM_CUDA(ctx, cudaMalloc(&f->d_raw, need));
            f->raw_cap = need;
        }
    }
    const int dpitch = contiguous && !repitch ? row_pitch : f->pitch;
    const size_t dstride = (size_t)dpitch * height;
    uint8_t* dl = f->d_img;
    uint8_t* dr = f->d_img + (size_t)n_pairs * dstride;
    uint8_t* rawl = f->d_raw;
    uint8_t* rawr = f->d_raw + (size_t)n_pairs * image_stride;
    const size_t cap = (size_t)f->kp_cap, np = (size_t)n_pairs;
    // chunk schedule: equal chunks.  Every chunk costs ~0.35 ms of launch tails on top of ~48 us per pair (measured with
    // VSLAM_FRONT_TRACE=1), so ramping the pipeline up with small chunks loses more than the exposed first upload costs.
    int c_start[FRONT_MAX_CHUNKS + 1], n_chunks = 0;
    {
        static const int chunk_env = getenv("VSLAM_FRONT_CHUNK") ? atoi(getenv("VSLAM_FRONT_CHUNK")) : 0;
        // 256-pair batches (round 2, tools/e2e_sweep.py): 16 -> 19.66 k, 32 -> 19.87 k, 64 -> 19.17 k e2e stereo fps; staging into
        // 16-byte aligned rows with 2-D copies of 1241-byte rows instead of the 1-D DMAs: 7.9 k
        int chunk = chunk_env > 0 ? chunk_env : (n_pairs >= 128 ? 2 * FRONT_CHUNK_PAIRS : FRONT_CHUNK_PAIRS);
        if (ceil_div(n_pairs, chunk) > FRONT_MAX_CHUNKS) chunk = ceil_div(n_pairs, FRONT_MAX_CHUNKS);
        c_start[0] = 0;
        // ramp: half-size first and last chunks shorten the exposed first upload and last download (256 pairs: 10.35 ->
        // 10.12 ms per call; VSLAM_FRONT_NO_RAMP=1 restores equal chunks)
        static const bool ramp_env = getenv("VSLAM_FRONT_NO_RAMP") == nullptr;
        const bool ramp = ramp_env && n_pairs >= 4 * chunk && ceil_div(n_pairs, chunk) + 1 <= FRONT_MAX_CHUNKS;
        for (int done = 0; done < n_pairs;) {
            const int step = ramp && done == 0 ? chunk / 2 : chunk;  // a half first chunk leaves a half last chunk
            done += n_pairs - done < step ? n_pairs - done : step;
            c_start[++n_chunks] = done;
        }
    }
    {
        const int stg = vslam_orb_prepare(ctx, width, height);  // tables go up on s, before the streams fork
        if (stg != VSLAM_OK) return stg;
    }
    // the copy streams start after whatever the caller already queued on the context stream
    VSLAM_CUDA(ctx, cudaEventRecord(f->ev_start, s));
    VSLAM_CUDA(ctx, cudaStreamWaitEvent(f->s_in, f->ev_start, 0));
    VSLAM_CUDA(ctx, cudaStreamWaitEvent(f->s_out, f->ev_start, 0));
    if (T_c_w) VSLAM_CUDA(ctx, cudaMemcpyAsync(f->d_pose, T_c_w, np * 96, cudaMemcpyHostToDevice, f->s_in));
    for (int c = 0; c < n_chunks; ++c) {  // all uploads are queued up front; they run back to back on the H2D engine
        const int p0 = c_start[c], nc = c_start[c + 1] - p0;
        if (repitch) {
            VSLAM_CUDA(ctx, cudaMemcpyAsync(rawl + (size_t)p0 * image_stride, left + (size_t)p0 * image_stride,
                                            (size_t)image_stride * nc, cudaMemcpyHostToDevice, f->s_in));
            VSLAM_CUDA(ctx, cudaMemcpyAsync(rawr + (size_t)p0 * image_stride, right + (size_t)p0 * image_stride,
                                            (size_t)image_stride * nc, cudaMemcpyHostToDevice, f->s_in));
        } else if (contiguous) {
            VSLAM_CUDA(ctx, cudaMemcpyAsync(dl + p0 * dstride, left + (size_t)p0 * image_stride, dstride * nc,
                                            cudaMemcpyHostToDevice, f->s_in));
            VSLAM_CUDA(ctx, cudaMemcpyAsync(dr + p0 * dstride, right + (size_t)p0 * image_stride, dstride * nc,
                                            cudaMemcpyHostToDevice, f->s_in));
        } else {
            for (int i = p0; i < p0 + nc; ++i) {
                VSLAM_CUDA(ctx, cudaMemcpy2DAsync(dl + i * dstride, dpitch, left + (size_t)i * image_stride, row_pitch,
                                                  width, height, cudaMemcpyHostToDevice, f->s_in));
                VSLAM_CUDA(ctx, cudaMemcpy2DAsync(dr + i * dstride, dpitch, right + (size_t)i * image_stride, row_pitch,
                                                  width, height, cudaMemcpyHostToDevice, f->s_in));
            }
        }
        VSLAM_CUDA(ctx, cudaEventRecord(f->ev_in[c], f->s_in));
    }
    // odd chunks run on a second compute stream with their own scratch slots (when the context has room for two chunks):
    // the launch tails of one chunk (a few long CTAs per kernel) are filled by the other chunk's kernels
    int chunk_max = 0;
    for (int c = 0; c < n_chunks; ++c) chunk_max = chunk_max > c_start[c + 1] - c_start[c] ? chunk_max : c_start[c + 1] - c_start[c];
    const bool two_streams = n_chunks > 1 && !ctx->serial && 4 * chunk_max <= ctx->cfg.max_images;
    if (two_streams) VSLAM_CUDA(ctx, cudaStreamWaitEvent(f->s_alt, f->ev_start, 0));
    for (int c = 0; c < n_chunks; ++c) {
        const int p0 = c_start[c], nc = c_start[c + 1] - p0;
        const bool alt = two_streams && (c & 1);
        cudaStream_t sc = alt ? f->s_alt : s;
        VSLAM_CUDA(ctx, cudaStreamWaitEvent(sc, f->ev_in[c], 0));
        ctx->stream = sc;  // every enqueue below launches on the context's current stream
        if (repitch) {
            const dim3 grid(ceil_div(dpitch, 128 * 16), height, nc);
            repitch_kernel<<<grid, 128, 0, sc>>>(rawl + (size_t)p0 * image_stride, row_pitch, image_stride, dl + p0 * dstride,
                                                 dpitch, (long long)dstride);
            repitch_kernel<<<grid, 128, 0, sc>>>(rawr + (size_t)p0 * image_stride, row_pitch, image_stride, dr + p0 * dstride,
                                                 dpitch, (long long)dstride);
            VSLAM_LAUNCH_CHECK(ctx, "repitch_kernel");
        }
        int st = front_enqueue_chunk(ctx, dl + p0 * dstride, dr + p0 * dstride, n_pairs, p0, nc, width, height, dpitch,
                                     (long long)dstride, nfeatures, anms_keep, anms_c, gate_rel, gate_abs, P1, P2,
                                     T_c_w ? f->d_pose : nullptr, f->d_kp, f->d_desc, f->d_nkp, f->d_match,
                                     f->d_nmatch, f->d_xyz, f->d_flags, alt ? chunk_max : 0);
        ctx->stream = s;
        if (st != VSLAM_OK) {
            cudaStreamSynchronize(f->s_in);
            cudaStreamSynchronize(f->s_out);
            cudaStreamSynchronize(f->s_alt);
            cudaStreamSynchronize(s);
            return st;
        }
        VSLAM_CUDA(ctx, cudaEventRecord(f->ev_done[c], sc));
        VSLAM_CUDA(ctx, cudaStreamWaitEvent(f->s_out, f->ev_done[c], 0));
        const size_t l0 = (size_t)p0, r0 = np + p0, n = (size_t)nc;
        cudaStream_t so = f->s_out;
        VSLAM_CUDA(ctx, cudaMemcpyAsync(kp + l0 * cap, f->d_kp + l0 * cap, n * cap * sizeof(vslam_keypoint), cudaMemcpyDeviceToHost, so));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(kp + r0 * cap, f->d_kp + r0 * cap, n * cap * sizeof(vslam_keypoint), cudaMemcpyDeviceToHost, so));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(desc + l0 * cap * 32, f->d_desc + l0 * cap * 32, n * cap * 32, cudaMemcpyDeviceToHost, so));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(desc + r0 * cap * 32, f->d_desc + r0 * cap * 32, n * cap * 32, cudaMemcpyDeviceToHost, so));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(matches + l0 * cap, f->d_match + l0 * cap, n * cap * sizeof(vslam_dmatch), cudaMemcpyDeviceToHost, so));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(xyz + l0 * cap * 3, f->d_xyz + l0 * cap * 3, n * cap * 12, cudaMemcpyDeviceToHost, so));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(flags + l0 * cap, f->d_flags + l0 * cap, n * cap, cudaMemcpyDeviceToHost, so));
    }
    VSLAM_CUDA(ctx, cudaMemcpyAsync(n_kp, f->d_nkp, 2 * np * sizeof(int32_t), cudaMemcpyDeviceToHost, f->s_out));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(n_matches, f->d_nmatch, np * sizeof(int32_t), cudaMemcpyDeviceToHost, f->s_out));
    if (two_streams) {  // join: the context stream (and the flag read below) waits for the second compute stream
        VSLAM_CUDA(ctx, cudaEventRecord(f->ev_join, f->s_alt));
        VSLAM_CUDA(ctx, cudaStreamWaitEvent(s, f->ev_join, 0));
    }
    f->inflight_pairs = n_pairs;
    f->inflight_chunks = n_chunks;
    memcpy(f->inflight_c_start, c_start, sizeof(int) * (n_chunks + 1));
    return VSLAM_OK;
}

extern "C" int vslam_stereo_frontend_batch_end(vslam_ctx* ctx) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !ctx->front || ctx->front->inflight_pairs <= 0) return VSLAM_E_INVALID;
    FrontState* f = ctx->front;
    const int n_pairs = f->inflight_pairs, n_chunks = f->inflight_chunks;
    const int* c_start = f->inflight_c_start;
    f->inflight_pairs = 0;
    const int st = vslam_orb_check_flags(ctx, 2 * n_pairs);  // synchronises the context stream
    VSLAM_CUDA(ctx, cudaStreamSynchronize(f->s_out));
    if (getenv("VSLAM_FRONT_TRACE")) {  // pipeline trace: when each chunk's upload and kernels finished, ms after the call began
        for (int c = 0; c < n_chunks; ++c) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, f->ev_start, f->ev_in[c]);
            cudaEventElapsedTime(&b, f->ev_start, f->ev_done[c]);
            fprintf(stderr, "[front] chunk %d pairs %d..%d  h2d done %.3f ms  kernels done %.3f ms\n", c, c_start[c], c_start[c + 1], a, b);
        }
    }
    return st;  // VSLAM_E_OVERFLOW if a work list overflowed
}

extern "C" int vslam_stereo_frontend_batch(vslam_ctx* ctx, const uint8_t* left, const uint8_t* right, int n_pairs,
                                           int width, int height, int row_pitch, long long image_stride, int nfeatures,
                                           int anms_keep, float anms_c, double gate_rel, double gate_abs,
                                           const double* P1, const double* P2, const double* T_c_w,
                                           vslam_keypoint* kp, uint8_t* desc, int32_t* n_kp, vslam_dmatch* matches,
                                           int32_t* n_matches, float* xyz, uint8_t* flags) {
    const int st = vslam_stereo_frontend_batch_begin(ctx, left, right, n_pairs, width, height, row_pitch, image_stride,
                                                     nfeatures, anms_keep, anms_c, gate_rel, gate_abs, P1, P2, T_c_w, kp,
                                                     desc, n_kp, matches, n_matches, xyz, flags);
    if (st != VSLAM_OK) return st;
    return vslam_stereo_frontend_batch_end(ctx);
}
