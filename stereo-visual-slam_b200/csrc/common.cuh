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

struct vslam_ctx {
    vslam_config cfg;
    cudaStream_t own_stream;
    cudaStream_t stream;
    int64_t launches;
    char err[256];
    int num_sms;
    OrbState* orb;
    MatchState* match;
    BaState* ba;
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

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// sub-module lifetime hooks (each .cu owns its state)
int vslam_match_init(vslam_ctx* ctx);
void vslam_match_free(vslam_ctx* ctx);
int vslam_orb_init(vslam_ctx* ctx);
void vslam_orb_free(vslam_ctx* ctx);
int vslam_ba_init(vslam_ctx* ctx);
void vslam_ba_free(vslam_ctx* ctx);
