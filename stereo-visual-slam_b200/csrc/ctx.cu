// Context lifetime, status strings and stream plumbing for the C-ABI in include/vslam_b200.h.
#include "common.cuh"

#include <stdlib.h>

extern "C" int vslam_abi_version(void) { return VSLAM_ABI_VERSION; }

// ---- pinned host memory pool ------------------------------------------------------------------------------------
// Image-sized host buffers of the drop-in layer (the cv::Mat behind cv::imread / the disparity image) come from here so
// that the library's uploads and downloads are plain DMAs instead of copies staged through the driver's bounce buffer.
// cudaHostAlloc costs far more than it saves per frame, so freed blocks are kept on per-size free lists and handed out
// again; the pool is process-wide and is left to the OS at exit (the CUDA runtime may already be gone by then).
#include <map>
#include <mutex>
#include <vector>
static std::mutex g_pool_mu;
static std::map<size_t, std::vector<void*>> g_pool_free;  // block size -> free blocks
static std::map<void*, size_t> g_pool_size;              // every block ever allocated -> its size

extern "C" void* vslam_host_alloc(size_t bytes) {
    const size_t sz = (bytes + 65535) & ~(size_t)65535;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto it = g_pool_free.find(sz);
        if (it != g_pool_free.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            return p;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, sz, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool_size[p] = sz;
    return p;
}

extern "C" void vslam_host_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool_size.find(p);
    if (it == g_pool_size.end()) return;  // not ours
    g_pool_free[it->second].push_back(p);
}

extern "C" const char* vslam_status_string(int status) {
    switch (status) {
        case VSLAM_OK: return "ok";
        case VSLAM_E_INVALID: return "invalid argument";
        case VSLAM_E_CAPACITY: return "capacity exceeded";
        case VSLAM_E_CUDA: return "CUDA error";
        case VSLAM_E_NODEVICE: return "no sm_100-class CUDA device";
        case VSLAM_E_NUMERIC: return "numeric failure";
        case VSLAM_E_OVERFLOW: return "device work list overflow";
        default: return "unknown status";
    }
}

extern "C" const char* vslam_last_error(const vslam_ctx* ctx) { return ctx ? ctx->err : ""; }

extern "C" int vslam_ctx_create(const vslam_config* cfg, vslam_ctx** out) {
    if (!cfg || !out) return VSLAM_E_INVALID;
    *out = nullptr;
    if (cfg->max_images < 0 || cfg->max_width < 0 || cfg->max_height < 0 || cfg->max_keypoints < 0 ||
        cfg->max_keypoints > 65535)
        return VSLAM_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return VSLAM_E_NODEVICE;  // no CPU fallback by design
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return VSLAM_E_NODEVICE;
    if (prop.major != 10) return VSLAM_E_NODEVICE;  // the only code in this library is sm_100a SASS
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);  // the caller's current device is restored on return
    if (cudaSetDevice(cfg->device) != cudaSuccess) return VSLAM_E_NODEVICE;

    vslam_ctx* ctx = (vslam_ctx*)calloc(1, sizeof(vslam_ctx));
    if (!ctx) return VSLAM_E_INVALID;
    ctx->cfg = *cfg;
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        free(ctx);
        return VSLAM_E_CUDA;
    }
    ctx->stream = ctx->own_stream;
    ctx->trec = (TimingRec*)calloc(VSLAM_TIMING_CAP, sizeof(TimingRec));
    int st = ctx->trec ? VSLAM_OK : VSLAM_E_INVALID;
    if (st == VSLAM_OK) st = vslam_match_init(ctx);
    if (st == VSLAM_OK) st = vslam_orb_init(ctx);
    if (st == VSLAM_OK) st = vslam_ba_init(ctx);
    if (st == VSLAM_OK) st = vslam_front_init(ctx);
    if (st == VSLAM_OK) st = vslam_pnp_init(ctx);
    if (st == VSLAM_OK) st = vslam_sgbm_init(ctx);
    if (st != VSLAM_OK) {
        vslam_ctx_destroy(ctx);
        if (prev_dev >= 0) cudaSetDevice(prev_dev);
        return st;
    }
    *out = ctx;
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return VSLAM_OK;
}

extern "C" void vslam_ctx_destroy(vslam_ctx* ctx) {
    if (!ctx) return;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->stream);
    vslam_sgbm_free(ctx);
    vslam_pnp_free(ctx);
    vslam_front_free(ctx);
    vslam_ba_free(ctx);
    vslam_orb_free(ctx);
    vslam_match_free(ctx);
    if (ctx->trec) {
        for (int i = 0; i < ctx->n_trec_alloc; ++i) {
            cudaEventDestroy(ctx->trec[i].a);
            cudaEventDestroy(ctx->trec[i].b);
        }
        free(ctx->trec);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    free(ctx);
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
}

extern "C" int vslam_ctx_set_stream(vslam_ctx* ctx, void* cuda_stream) {
    if (!ctx) return VSLAM_E_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return VSLAM_OK;
}

extern "C" int vslam_ctx_set_concurrency(vslam_ctx* ctx, int on) {
    if (!ctx) return VSLAM_E_INVALID;
    ctx->serial = on ? 0 : 1;
    return VSLAM_OK;
}

extern "C" int vslam_ctx_synchronize(vslam_ctx* ctx) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx) return VSLAM_E_INVALID;
    VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VSLAM_OK;
}

extern "C" int64_t vslam_ctx_launch_count(const vslam_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- optional per-launch timing (CUDA events on the launching stream) -------------------------------------
static const char* const k_kernel_names[VK_COUNT] = {
    "resize_level_kernel", "fast_kernel", "harris_select_kernel", "blur_kernel", "anms_kernel", "describe_kernel",
    "hamming_argmin_kernel", "crosscheck_gate_compact_kernel", "triangulate_kernel", "ba_lm_kernel",
    "ba_solve_kernel", "ba_update_kernel", "ba_misc_kernel", "pnp_hypothesis_kernel", "sgbm_prefilter_kernel", "sgbm_cost_kernel",
    "sgbm_vertical_kernel", "sgbm_row_forward_kernel", "sgbm_row_backward_kernel", "sgbm_post_kernels", "pnp_refine_kernel"};

extern "C" int vslam_kernel_count(void) { return VK_COUNT; }
extern "C" const char* vslam_kernel_name(int id) { return id >= 0 && id < VK_COUNT ? k_kernel_names[id] : ""; }

extern "C" int vslam_ctx_timing_enable(vslam_ctx* ctx, int on) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx) return VSLAM_E_INVALID;
    VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->timing_on = on ? 1 : 0;
    ctx->n_trec = 0;
    for (int i = 0; i < VK_COUNT; ++i) {
        ctx->t_ms[i] = 0;
        ctx->t_launches[i] = 0;
    }
    return VSLAM_OK;
}

// synchronise, fold the pending event pairs into per-kernel totals and report kernel `id`
extern "C" int vslam_ctx_timing_read(vslam_ctx* ctx, int id, double* total_ms, int64_t* launches) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || id < 0 || id >= VK_COUNT) return VSLAM_E_INVALID;
    if (ctx->n_trec > 0) {
        VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < ctx->n_trec; ++i) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->trec[i].a, ctx->trec[i].b) == cudaSuccess) {
                ctx->t_ms[ctx->trec[i].id] += ms;
                ctx->t_launches[ctx->trec[i].id]++;
            }
        }
        ctx->n_trec = 0;
    }
    if (total_ms) *total_ms = ctx->t_ms[id];
    if (launches) *launches = ctx->t_launches[id];
    return VSLAM_OK;
}
