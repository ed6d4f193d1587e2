// K10: brute-force Hamming matching with mutual cross-check and the reference's distance gate.
//
// Replaces cv::BFMatcher(NORM_HAMMING, crossCheck=true)::match + the gate loop of VO::feature_matching
// (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:24,33,219-251).
//
// Kernel 1 (hamming_argmin): the Nq x Nt distance matrix is computed ONCE, tile by tile, and reduced both
// ways in the same pass: forward  fw[i] = min_j (D[i][j] << 16 | j)   (per-thread running min, query in registers)
//                        backward bw[j] = min_i (D[i][j] << 16 | i)   (redux.sync.min across the warp + smem atomicMin)
// Packing distance above the index makes "first minimum wins ties" a plain unsigned min.
// Train descriptors are staged into shared memory with cp.async (LDGSTS) 16-byte copies and read back as
// warp-wide broadcasts; query descriptors live in registers (2 per thread, 128-bit loads).
// Kernel 2 (crosscheck_gate_compact): one CTA per pair: mutual test, min distance, gate, ordered compaction.
#include "common.cuh"

#include <cuda_pipeline.h>

#define MT_THREADS 128
#ifndef MT_QPT
#define MT_QPT 4                        // queries per thread (2 -> 4: the per-train overhead -- two broadcast loads, the
#endif                                  // warp minimum, the shared-memory atomic -- is shared by twice the distances)
#define MT_TILE_Q (MT_THREADS * MT_QPT) // queries per CTA
#define MT_TILE_T 128                   // trains per CTA
#define CC_THREADS 1024

struct MatchState {
    uint32_t* d_keys;  // [pairs][2][max_rows]  fw then bw packed keys
    uint8_t* d_q;      // staging for the host-buffer entry point
    uint8_t* d_t;
    vslam_dmatch* d_out;
    int32_t* d_cnt;  // nq, nt, n_out
    int max_pairs;
    int max_rows;
    vslam_dmatch* h_out;  // pinned
    int32_t* h_cnt;       // pinned
};

// 256-bit Hamming distance.  POPC issues on the quarter-rate XU pipe, everything else on the integer ALU pipe, which binds
// (ncu: alu 87 %, xu 62 % with three carry-save adders).  A carry-save adder (2 LOP3) turns three words into a "ones"
// and a "twos" word and saves one POPC; MT_CSA of them balance the two pipes.  Measured per 256 pairs x 2000^2 distances:
// 3 adders / 5 POPC 1.67 ms, 2 adders / 6 POPC 1.55 ms (the default), 1 adder / 7 POPC 1.67 ms, plain 8 POPC 1.88 ms.
#ifndef MT_CSA
#define MT_CSA 2
#endif
__device__ __forceinline__ uint32_t csa_sum(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
__device__ __forceinline__ uint32_t csa_carry(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

__device__ __forceinline__ uint32_t hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
#if MT_CSA == 0
    return __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
#elif MT_CSA == 1
    const uint32_t s1 = csa_sum(x0, x1, x2), c1 = csa_carry(x0, x1, x2);
    return __popc(s1) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7) + 2 * __popc(c1);
#elif MT_CSA == 2
    const uint32_t s1 = csa_sum(x0, x1, x2), c1 = csa_carry(x0, x1, x2);
    const uint32_t s2 = csa_sum(x3, x4, x5), c2 = csa_carry(x3, x4, x5);
    return __popc(s1) + __popc(s2) + __popc(x6) + __popc(x7) + 2 * (__popc(c1) + __popc(c2));
#else
    const uint32_t s1 = csa_sum(x0, x1, x2), c1 = csa_carry(x0, x1, x2);
    const uint32_t s2 = csa_sum(x3, x4, x5), c2 = csa_carry(x3, x4, x5);
    const uint32_t s3 = csa_sum(s1, s2, x6), c3 = csa_carry(s1, s2, x6);
    const uint32_t twos = __popc(c1) + __popc(c2) + __popc(c3);
    return __popc(s3) + __popc(x7) + 2 * twos;
#endif
}

__global__ void __launch_bounds__(MT_THREADS)
hamming_argmin_kernel(const uint8_t* __restrict__ query, const int32_t* __restrict__ d_nq, int q_stride_rows,
                      const uint8_t* __restrict__ train, const int32_t* __restrict__ d_nt, int t_stride_rows,
                      uint32_t* __restrict__ keys, int max_rows) {
    const int pair = blockIdx.z;
    const int nq = d_nq[pair];
    const int nt = d_nt[pair];
    const int q0 = blockIdx.x * MT_TILE_Q;
    const int t0 = blockIdx.y * MT_TILE_T;
    if (q0 >= nq || t0 >= nt) return;

    __shared__ uint4 s_t[MT_TILE_T * 2];
    __shared__ uint32_t s_bw[MT_TILE_T];

    const uint4* tq = reinterpret_cast<const uint4*>(query + (size_t)pair * q_stride_rows * 32);
    const uint4* tt = reinterpret_cast<const uint4*>(train + (size_t)pair * t_stride_rows * 32);
    const int tile_n = min(MT_TILE_T, nt - t0);

    // stage the train tile: tile_n*2 16-byte chunks, coalesced, asynchronous
    for (int c = threadIdx.x; c < tile_n * 2; c += MT_THREADS)
        __pipeline_memcpy_async(&s_t[c], &tt[(size_t)t0 * 2 + c], 16);
    __pipeline_commit();
    for (int j = threadIdx.x; j < MT_TILE_T; j += MT_THREADS) s_bw[j] = 0xFFFFFFFFu;

    // queries of this thread: rows q0 + tid and q0 + tid + MT_THREADS (coalesced 32-byte rows)
    uint4 qa[MT_QPT][2];
    int qi[MT_QPT];
    bool qv[MT_QPT];
#pragma unroll
    for (int r = 0; r < MT_QPT; ++r) {
        qi[r] = q0 + threadIdx.x + r * MT_THREADS;
        qv[r] = qi[r] < nq;
        if (qv[r]) {
            qa[r][0] = __ldg(&tq[(size_t)qi[r] * 2]);
            qa[r][1] = __ldg(&tq[(size_t)qi[r] * 2 + 1]);
        } else {
            qa[r][0] = make_uint4(0, 0, 0, 0);
            qa[r][1] = make_uint4(0, 0, 0, 0);
        }
    }
    uint32_t best[MT_QPT];
#pragma unroll
    for (int r = 0; r < MT_QPT; ++r) best[r] = 0xFFFFFFFFu;

    __pipeline_wait_prior(0);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    if (q0 + MT_TILE_Q <= nq) {
        // every query of the tile exists (3 of 4 tiles at 2000 keypoints): no validity selects in the loop
#pragma unroll 4
        for (int j = 0; j < tile_n; ++j) {
            const uint4 b0 = s_t[2 * j];
            const uint4 b1 = s_t[2 * j + 1];
            uint32_t kb = 0xFFFFFFFFu;
#pragma unroll
            for (int r = 0; r < MT_QPT; ++r) {
                const uint32_t d16 = hamming256(qa[r][0], qa[r][1], b0, b1) << 16;
                best[r] = min(best[r], d16 | (uint32_t)(t0 + j));
                kb = min(kb, d16 | (uint32_t)qi[r]);
            }
            kb = __reduce_min_sync(0xFFFFFFFFu, kb);
            if (lane == 0)
                asm volatile("red.shared.min.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bw[j])), "r"(kb) : "memory");
        }
    } else {
#pragma unroll 4
        for (int j = 0; j < tile_n; ++j) {
            const uint4 b0 = s_t[2 * j];
            const uint4 b1 = s_t[2 * j + 1];
            uint32_t kb = 0xFFFFFFFFu;
#pragma unroll
            for (int r = 0; r < MT_QPT; ++r) {
                const uint32_t d = hamming256(qa[r][0], qa[r][1], b0, b1);
                const uint32_t kf = (d << 16) | (uint32_t)(t0 + j);
                best[r] = min(best[r], qv[r] ? kf : 0xFFFFFFFFu);
                const uint32_t kq = qv[r] ? ((d << 16) | (uint32_t)qi[r]) : 0xFFFFFFFFu;
                kb = min(kb, kq);
            }
            kb = __reduce_min_sync(0xFFFFFFFFu, kb);
            // issued as written: nvcc otherwise wraps a shared-memory atomic in its own warp-aggregation sequence
            if (lane == 0)
                asm volatile("red.shared.min.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bw[j])), "r"(kb) : "memory");
        }
    }
    __syncthreads();

    uint32_t* fw = keys + (size_t)pair * 2 * max_rows;
    uint32_t* bw = fw + max_rows;
#pragma unroll
    for (int r = 0; r < MT_QPT; ++r)
        if (qv[r]) atomicMin(&fw[qi[r]], best[r]);
    for (int j = threadIdx.x; j < tile_n; j += MT_THREADS) atomicMin(&bw[t0 + j], s_bw[j]);
}

__global__ void __launch_bounds__(CC_THREADS)
crosscheck_gate_compact_kernel(const int32_t* __restrict__ d_nq, const int32_t* __restrict__ d_nt,
                               const uint32_t* __restrict__ keys, int max_rows, int cross_check, double gate_rel,
                               double gate_abs, vslam_dmatch* __restrict__ out, int out_stride,
                               int32_t* __restrict__ n_out) {
    const int pair = blockIdx.x;
    const int nq = d_nq[pair];
    const int nt = d_nt[pair];
    const uint32_t* fw = keys + (size_t)pair * 2 * max_rows;
    const uint32_t* bw = fw + max_rows;
    vslam_dmatch* o = out + (size_t)pair * out_stride;

    __shared__ uint32_t s_warp[CC_THREADS / 32];
    __shared__ uint32_t s_min;
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_min = 0xFFFFFFFFu;
        s_base = 0;
    }
    __syncthreads();
    if (nq <= 0 || nt <= 0) {
        if (tid == 0) n_out[pair] = 0;
        return;
    }

    // phase 1: smallest distance among the (mutual) matches
    uint32_t dmin = 0xFFFFFFFFu;
    for (int i = tid; i < nq; i += CC_THREADS) {
        const uint32_t k = fw[i];
        const uint32_t j = k & 0xFFFFu;
        const bool ok = !cross_check || ((bw[j] & 0xFFFFu) == (uint32_t)i);
        if (ok) dmin = min(dmin, k >> 16);
    }
    dmin = __reduce_min_sync(0xFFFFFFFFu, dmin);
    if (lane == 0) atomicMin(&s_min, dmin);
    __syncthreads();
    const uint32_t min_d = s_min;
    double thr = 1e300;
    if (gate_rel >= 0.0 && min_d != 0xFFFFFFFFu) thr = fmax(gate_rel * (double)(float)min_d, gate_abs);

    // phase 2: ordered compaction (ascending queryIdx)
    for (int i0 = 0; i0 < nq; i0 += CC_THREADS) {
        const int i = i0 + tid;
        bool keep = false;
        uint32_t k = 0, j = 0;
        if (i < nq) {
            k = fw[i];
            j = k & 0xFFFFu;
            keep = !cross_check || ((bw[j] & 0xFFFFu) == (uint32_t)i);
            keep = keep && ((double)(float)(k >> 16) <= thr);
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint32_t wsum = 0, total = 0;
#pragma unroll
        for (int w = 0; w < CC_THREADS / 32; ++w) {
            const uint32_t c = s_warp[w];
            wsum += (w < warp) ? c : 0;
            total += c;
        }
        if (keep) {
            const int pos = s_base + (int)wsum + __popc(bal & ((1u << lane) - 1));
            vslam_dmatch m;
            m.queryIdx = i;
            m.trainIdx = (int)j;
            m.imgIdx = 0;
            m.distance = (float)(k >> 16);
            o[pos] = m;
        }
        __syncthreads();
        if (tid == 0) s_base += (int)total;
        __syncthreads();
    }
    if (tid == 0) n_out[pair] = s_base;
}

int vslam_match_init(vslam_ctx* ctx) {
    MatchState* m = (MatchState*)calloc(1, sizeof(MatchState));
    if (!m) return VSLAM_E_INVALID;
    ctx->match = m;
    m->max_pairs = ctx->cfg.max_images > 0 ? ctx->cfg.max_images : 1;
    m->max_rows = ctx->cfg.max_keypoints > 0 ? ctx->cfg.max_keypoints : 1;
    const size_t rows = (size_t)m->max_rows;
    VSLAM_CUDA(ctx, cudaMalloc(&m->d_keys, (size_t)m->max_pairs * 2 * rows * sizeof(uint32_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&m->d_q, rows * 32));
    VSLAM_CUDA(ctx, cudaMalloc(&m->d_t, rows * 32));
    VSLAM_CUDA(ctx, cudaMalloc(&m->d_out, rows * sizeof(vslam_dmatch)));
    VSLAM_CUDA(ctx, cudaMalloc(&m->d_cnt, 4 * sizeof(int32_t)));
    VSLAM_CUDA(ctx, cudaMallocHost(&m->h_out, rows * sizeof(vslam_dmatch)));
    VSLAM_CUDA(ctx, cudaMallocHost(&m->h_cnt, 4 * sizeof(int32_t)));
    return VSLAM_OK;
}

void vslam_match_free(vslam_ctx* ctx) {
    MatchState* m = ctx->match;
    if (!m) return;
    cudaFree(m->d_keys);
    cudaFree(m->d_q);
    cudaFree(m->d_t);
    cudaFree(m->d_out);
    cudaFree(m->d_cnt);
    cudaFreeHost(m->h_out);
    cudaFreeHost(m->h_cnt);
    free(m);
    ctx->match = nullptr;
}

extern "C" int vslam_match_hamming_batch_dev(vslam_ctx* ctx, const uint8_t* d_query, const int32_t* d_nq,
                                             int q_stride_rows, const uint8_t* d_train, const int32_t* d_nt,
                                             int t_stride_rows, int batch, int max_rows, int cross_check,
                                             double gate_rel, double gate_abs, vslam_dmatch* d_out, int out_stride,
                                             int32_t* d_n_out) {
    VslamDeviceGuard device_guard__(ctx);
    return vslam_match_enqueue(ctx, d_query, d_nq, q_stride_rows, d_train, d_nt, t_stride_rows, batch, max_rows,
                               cross_check, gate_rel, gate_abs, d_out, out_stride, d_n_out, 0);
}

int vslam_match_enqueue(vslam_ctx* ctx, const uint8_t* d_query, const int32_t* d_nq, int q_stride_rows,
                        const uint8_t* d_train, const int32_t* d_nt, int t_stride_rows, int batch, int max_rows,
                        int cross_check, double gate_rel, double gate_abs, vslam_dmatch* d_out, int out_stride,
                        int32_t* d_n_out, int scratch_pair0) {
    if (!ctx || !d_query || !d_train || !d_nq || !d_nt || !d_out || !d_n_out) return VSLAM_E_INVALID;
    if (batch <= 0 || max_rows <= 0 || scratch_pair0 < 0) return VSLAM_E_INVALID;
    MatchState* m = ctx->match;
    if (scratch_pair0 + batch > m->max_pairs || max_rows > m->max_rows || max_rows > 65535) return VSLAM_E_CAPACITY;
    if (((uintptr_t)d_query | (uintptr_t)d_train) & 15) return VSLAM_E_INVALID;
    uint32_t* keys = m->d_keys + (size_t)scratch_pair0 * 2 * m->max_rows;
    VSLAM_CUDA(ctx, cudaMemsetAsync(keys, 0xFF, (size_t)batch * 2 * m->max_rows * sizeof(uint32_t), ctx->stream));
    dim3 grid(ceil_div(max_rows, MT_TILE_Q), ceil_div(max_rows, MT_TILE_T), batch);
    vslam_time_begin(ctx, VK_HAMMING_ARGMIN);
    hamming_argmin_kernel<<<grid, MT_THREADS, 0, ctx->stream>>>(d_query, d_nq, q_stride_rows, d_train, d_nt,
                                                                t_stride_rows, keys, m->max_rows);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "hamming_argmin_kernel");
    vslam_time_begin(ctx, VK_CROSSCHECK);
    crosscheck_gate_compact_kernel<<<batch, CC_THREADS, 0, ctx->stream>>>(
        d_nq, d_nt, keys, m->max_rows, cross_check, gate_rel, gate_abs, d_out, out_stride, d_n_out);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "crosscheck_gate_compact_kernel");
    return VSLAM_OK;
}

extern "C" int vslam_match_hamming(vslam_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt,
                                   int cross_check, double gate_rel, double gate_abs, vslam_dmatch* out,
                                   int* n_out) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !n_out || nq < 0 || nt < 0) return VSLAM_E_INVALID;
    *n_out = 0;
    if (nq == 0 || nt == 0) return VSLAM_OK;
    if (!query || !train || !out) return VSLAM_E_INVALID;
    MatchState* m = ctx->match;
    if (nq > m->max_rows || nt > m->max_rows) return VSLAM_E_CAPACITY;
    cudaStream_t s = ctx->stream;
    m->h_cnt[0] = nq;
    m->h_cnt[1] = nt;
    VSLAM_CUDA(ctx, cudaMemcpyAsync(m->d_cnt, m->h_cnt, 2 * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(m->d_q, query, (size_t)nq * 32, cudaMemcpyHostToDevice, s));
    VSLAM_CUDA(ctx, cudaMemcpyAsync(m->d_t, train, (size_t)nt * 32, cudaMemcpyHostToDevice, s));
    int st = vslam_match_hamming_batch_dev(ctx, m->d_q, m->d_cnt, m->max_rows, m->d_t, m->d_cnt + 1, m->max_rows, 1,
                                           nq > nt ? nq : nt, cross_check, gate_rel, gate_abs, m->d_out, m->max_rows,
                                           m->d_cnt + 2);
    if (st != VSLAM_OK) return st;
    VSLAM_CUDA(ctx, cudaMemcpyAsync(m->h_cnt + 2, m->d_cnt + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    const int n = m->h_cnt[2];
    if (n > 0) {
        VSLAM_CUDA(ctx, cudaMemcpyAsync(out, m->d_out, (size_t)n * sizeof(vslam_dmatch), cudaMemcpyDeviceToHost, s));
        VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    }
    *n_out = n;
    return VSLAM_OK;
}
