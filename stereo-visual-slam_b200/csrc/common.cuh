// Internal definitions shared by the kernels behind include/vslam_b200.h (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vslam_b200.h"

#define VSLAM_NLEVELS 8
#define VSLAM_NUM_SMS 148

struct OrbState;   // orb.cu
struct MatchState; // match.cu
struct BaState;    // ba.cu
struct FrontState; // frontend.cu
struct PnpState;   // pnp.cu
struct SgbmState;  // sgbm.cu

// kernel ids for the optional per-launch CUDA-event timing (vslam_ctx_timing_*)
enum {
    VK_RESIZE = 0, VK_FAST, VK_HARRIS_SELECT, VK_BLUR, VK_ANMS, VK_DESCRIBE, VK_HAMMING_ARGMIN, VK_CROSSCHECK,
    VK_TRIANGULATE, VK_BA_BUILD, VK_BA_SOLVE, VK_BA_UPDATE, VK_BA_MISC, VK_PNP, VK_SGBM_PREFILTER, VK_SGBM_COST,
    VK_SGBM_VERTICAL, VK_SGBM_ROW_FWD, VK_SGBM_ROW_BWD, VK_SGBM_POST, VK_PNP_REFINE, VK_COUNT
};
#define VSLAM_TIMING_CAP 16384
struct TimingRec {
    int id;
    cudaEvent_t a, b;
};

struct vslam_ctx {
    vslam_config cfg;
    cudaStream_t own_stream;
    cudaStream_t stream;
    int64_t launches;
    char err[256];
    int num_sms;
    int timing_on;
    int serial;  // 1 = no internal second stream (vslam_ctx_set_concurrency(ctx, 0))
    int n_trec, n_trec_alloc;
    TimingRec* trec;
    double t_ms[VK_COUNT];
    int64_t t_launches[VK_COUNT];
    OrbState* orb;
    MatchState* match;
    BaState* ba;
    FrontState* front;
    PnpState* pnp;
    SgbmState* sgbm;
};

static inline int vslam_set_cuda_error(vslam_ctx* ctx, cudaError_t e, const char* where) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", where, cudaGetErrorString(e));
    return VSLAM_E_CUDA;
}

#define VSLAM_CUDA(ctx, call)                                                   \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return vslam_set_cuda_error((ctx), e__, #call); \
    } while (0)

#define VSLAM_LAUNCH_CHECK(ctx, name)                                            \
    do {                                                                         \
        (ctx)->launches++;                                                       \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) return vslam_set_cuda_error((ctx), e__, (name)); \
    } while (0)

// bracket a launch with events on the context stream when timing is enabled
static inline void vslam_time_begin(vslam_ctx* ctx, int id) {
    if (!ctx->timing_on || ctx->n_trec >= VSLAM_TIMING_CAP) return;
    TimingRec* r = &ctx->trec[ctx->n_trec];
    if (ctx->n_trec >= ctx->n_trec_alloc) {
        cudaEventCreate(&r->a);
        cudaEventCreate(&r->b);
        ctx->n_trec_alloc = ctx->n_trec + 1;
    }
    r->id = id;
    cudaEventRecord(r->a, ctx->stream);
}
static inline void vslam_time_end(vslam_ctx* ctx) {
    if (!ctx->timing_on || ctx->n_trec >= VSLAM_TIMING_CAP) return;
    cudaEventRecord(ctx->trec[ctx->n_trec].b, ctx->stream);
    ctx->n_trec++;
}

// Every entry point runs on the device of its context, whatever device the calling thread had current, and leaves the
// caller's current device as it found it (a process may hold contexts on several devices: vslam_ba_optimize_multi).
struct VslamDeviceGuard {
    int prev;
    bool switched;
    explicit VslamDeviceGuard(const vslam_ctx* ctx) : prev(-1), switched(false) {
        if (!ctx) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != ctx->cfg.device) switched = cudaSetDevice(ctx->cfg.device) == cudaSuccess;
    }
    ~VslamDeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Level 0 is read in place from the caller's buffers (left images first, then right images).
struct ImgSrc {
    const uint8_t* base[2];
    long long img_stride;  // bytes between consecutive images of one base
    int pitch;             // bytes between rows
    int per_base;          // images per base pointer
    int out_slot[2];       // output image slot of the first image of each base (kp / desc / count arrays)
};

// internal cross-module entry points
// scratch0: first image slot of the context's scratch (pyramids, candidate lists ...) this call may use; calls on
// different streams must use disjoint slot ranges
int vslam_orb_enqueue(vslam_ctx* ctx, const ImgSrc& src, int n_img, int w, int h, int nfeatures, int anms_keep,
                      float anms_c, vslam_keypoint* d_kp, uint8_t* d_desc, int32_t* d_n, int scratch0);
int vslam_orb_prepare(vslam_ctx* ctx, int w, int h);  // upload the per-geometry tables on the context stream
// vslam_match_hamming_batch_dev on the key scratch of pairs [scratch_pair0, scratch_pair0 + batch)
int vslam_match_enqueue(vslam_ctx* ctx, const uint8_t* d_query, const int32_t* d_nq, int q_stride_rows,
                        const uint8_t* d_train, const int32_t* d_nt, int t_stride_rows, int batch, int max_rows,
                        int cross_check, double gate_rel, double gate_abs, vslam_dmatch* d_out, int out_stride,
                        int32_t* d_n_out, int scratch_pair0);
int vslam_orb_check_flags(vslam_ctx* ctx, int n_img);  // synchronises the stream; reads and clears the sticky flags

// sub-module lifetime hooks (each .cu owns its state)
int vslam_match_init(vslam_ctx* ctx);
void vslam_match_free(vslam_ctx* ctx);
int vslam_orb_init(vslam_ctx* ctx);
void vslam_orb_free(vslam_ctx* ctx);
int vslam_ba_init(vslam_ctx* ctx);
void vslam_ba_free(vslam_ctx* ctx);
int vslam_front_init(vslam_ctx* ctx);
void vslam_front_free(vslam_ctx* ctx);
int vslam_pnp_init(vslam_ctx* ctx);
void vslam_pnp_free(vslam_ctx* ctx);
int vslam_sgbm_init(vslam_ctx* ctx);
void vslam_sgbm_free(vslam_ctx* ctx);
