// K18-K23  dense semi-global stereo matching, the reference's actual depth source.
//
// Replaces cv::StereoSGBM::create(0, 96, 9, 8*9*9, 32*9*9, 1, 63, 10, 100, 32)->compute(left, right, disparity_sgbm)
// in VO::disparity_map (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:159-174).  The arithmetic
// lives in OpenCV (calib3d/src/stereosgbm.cpp, un-vendored); oracle/sgbm_restate.py restates it and is pinned
// bit-exact against live cv2 4.13.0.  Output is bit-exact CV_16S disparity*16 with (minD-1)*16 = -16 as "invalid".
//
// OpenCV walks the image row by row on one core, carrying five path costs per pixel.  The five paths are independent
// of each other, so here every path direction is its own sweep over a materialised cost volume:
//   K18 sgbm_prefilter_kernel   x-Sobel + clip table, raw row, half-pixel min/max (Birchfield-Tomasi operands)
//   K19 sgbm_cost_kernel        BT pixel cost on both planes + 9x9 box sum (replicated borders) -> C[y][x][d] u16
//   K20 sgbm_vertical_kernel    paths from the row above (up-left, up, up-right): two paths per warp (one per
//                               half-warp), diagonal paths wrap around the image so every path walks all H rows with
//                               no inter-warp traffic; rows of C arrive through a cp.async.bulk ring
//   K21 sgbm_row_forward_kernel / sgbm_row_backward_kernel  one warp per row: left->right path, then right->left
//                               path fused with the winner-take-all, uniqueness test, sub-pixel fit, right-image map
//                               and LR consistency check
//   K22 sgbm_median_kernel      cv::medianBlur(3)
//   K23 speckle_*_kernel        cv::filterSpeckles as union-find connected components
// 96 disparities are held as 48 packed s16x2 words; a path step is VIADD.16x2 / VIMNMX.S16x2 on 2 words per lane over 24
// lanes (row sweeps) or 3 words per lane over a half-warp (vertical sweep) plus CREDUX.MIN for min_k L_r(p-r, k).  All
// volumes are HBM-resident u16 and stream through cp.async.bulk (TMA) rings; DESIGN.md §4 has the bounds.
#include "common.cuh"

#include <stdlib.h>

#define SG_D 96
#define SG_NDP 48                    // packed disparity pairs
#define SG_R 4                       // block radius (blockSize 9)
#define SG_BIG2 0x75307530u          // 30000 | 30000 << 16: "no predecessor" cost, > any reachable min + P2
#define SG_MAXCOST 32767
#define SG_INVALID (-16)
#define SG_HW 2                      // rows (warps) per CTA of the row sweeps
#define SG_CHUNK_PAIRS 8             // pairs per scratch chunk (8 x 331 MB)

struct SgParams {
    int P1, P2, disp12, uniq, ftzero, speckle_window, speckle_diff;
};

struct SgbmState {
    uint2* d_pre;       // [2*chunk][H][W] {sobel pack, raw pack}: v | min << 8 | max << 16
    uint32_t* d_C;      // [chunk][H][W1][48] packed u16 pairs
    uint32_t* d_L[3];   // path costs from the row above; d_L[0] is overwritten with S4 = sat(L0+L1+L2+L3)
    int16_t* d_raw;     // [chunk][H][W]
    uint2* d_rec;       // [chunk][H][W1] per-column records of the row sweep
    int16_t* d_med;
    int32_t* d_label;
    int32_t* d_size;
    int32_t* d_runlen;
    uint8_t* d_img;     // host-entry staging: [2*chunk][H][pitch]
    int16_t* d_out;
    float* d_outf;
    int cap_pairs, cap_w, cap_h, pitch;
    int stop_after;     // test tap: 1 = stop after the vertical sweep (keeps L1 intact)
    cudaStream_t s_alt; // second stream: odd chunks of a device-resident batch run here on their own scratch slots
    cudaEvent_t ev_fork, ev_join;
};

// ---------------------------------------------------------------------------------------------------------------
// K18: per pixel the Birchfield-Tomasi operands of both planes
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgbm_prefilter_kernel(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right,
                                                             long long img_stride, int pitch, int n_pairs, int W, int H,
                                                             int ftzero, uint2* __restrict__ pre) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= W) return;
    const uint8_t* base = img < n_pairs ? left + (long long)img * img_stride : right + (long long)(img - n_pairs) * img_stride;
    const uint8_t* r = base + (long long)y * pitch;
    const uint8_t* up = base + (long long)max(y - 1, 0) * pitch;
    const uint8_t* dn = base + (long long)min(y + 1, H - 1) * pitch;
    int s[3], v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int xx = min(max(x - 1 + k, 0), W - 1);
        if (xx <= 0 || xx >= W - 1) {
            s[k] = ftzero;  // tab[0]: zero gradient, on both planes (stereosgbm.cpp calcPixelCostBT border init)
            v[k] = ftzero;
        } else {
            const int g = ((int)r[xx + 1] - (int)r[xx - 1]) * 2 + (int)up[xx + 1] - (int)up[xx - 1] + (int)dn[xx + 1] - (int)dn[xx - 1];
            s[k] = min(max(g, -ftzero), ftzero) + ftzero;
            v[k] = r[xx];
        }
    }
    const int sl = (s[1] + s[0]) >> 1, sr = (s[1] + s[2]) >> 1, vl = (v[1] + v[0]) >> 1, vr = (v[1] + v[2]) >> 1;
    uint2 o;
    o.x = (uint32_t)s[1] | (uint32_t)min(min(sl, sr), s[1]) << 8 | (uint32_t)max(max(sl, sr), s[1]) << 16;
    o.y = (uint32_t)v[1] | (uint32_t)min(min(vl, vr), v[1]) << 8 | (uint32_t)max(max(vl, vr), v[1]) << 16;
    pre[((size_t)img * H + y) * W + x] = o;
}

// ---------------------------------------------------------------------------------------------------------------
// K19: pixel cost + 9x9 box sum.  CTA = CK_CW output columns x one row band, marching down the rows.
// Thread (g, q) owns 4 consecutive columns and the 4 disparities 4q..4q+3 (two s16x2 words).  Walking its columns
// left to right, the right-image operand of disparity d at column x is position x - d: one new position per column
// feeds all four disparities (register window), and the operands of the second word are the first word's operands
// of two columns ago.  Operands stay byte-packed (v | min << 8 | max << 16, as the prefilter wrote them) in shared
// memory and are widened with PRMT; u - v is computed as u + (255 - v) with the +255 bias folded into the max.
// Pixel costs go through a shared exchange buffer once (no halo recomputation inside the CTA), the horizontal
// window slides in registers, the vertical window is a 9-slot ring in shared memory.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bcast16(int v) { return (uint32_t)(v & 0xffff) * 0x00010001u; }

#define CK_Q 24                        // threads per column group
#define CK_GC 4                        // columns per group
#define CK_CW 108                      // output columns per CTA
#define CK_NG (CK_CW / CK_GC + 2)      // 27 output groups + one halo group (= the 4 halo columns) on each side
#define CK_COLS (CK_NG * CK_GC)        // 116 columns whose pixel cost the CTA computes
#define CK_GSTR (CK_GC * SG_NDP + 16)   // words per column group (4 x 48 + 16: neighbouring groups start 16 banks apart)
#define CK_NPOS 232                    // right-image positions staged per row (CK_COLS + 95 + 3, rounded up)
#define CK_THREADS (CK_NG * CK_Q)      // 696
#define CK_WIN (CK_GC + 2 * SG_R)       // exchange-buffer columns under one thread's horizontal windows
#define CK_RSLOT ((CK_NG - 2) * CK_GSTR)  // words per ring slot
#define CK_SMEM ((CK_COLS + CK_NPOS) * 8 + CK_NG * CK_GSTR * 4 + 9 * CK_RSLOT * 4)
#define CK_FF 0x00ff00ffu

struct BtOps {  // widened operands of one (position pair, channel): v, min, 255 - v, 255 - max
    uint32_t v, v0, vc, v1c;
};

__device__ __forceinline__ BtOps bt_expand(uint32_t p_lo, uint32_t p_hi) {  // low half: p_lo's pixel, high half: p_hi's
    BtOps o;
    o.v = __byte_perm(p_lo, p_hi, 0x7430);
    o.v0 = __byte_perm(p_lo, p_hi, 0x7531);
    o.vc = o.v ^ CK_FF;
    o.v1c = __byte_perm(p_lo, p_hi, 0x7632) ^ CK_FF;
    return o;
}

// min(c0, c1) + 255 with c0 = max(0, u - v1, v0 - u), c1 = max(0, v - u1, u0 - v)
__device__ __forceinline__ uint32_t bt_cost255(const BtOps& u, const BtOps& v) {
    const uint32_t c0 = __vimax3_s16x2(u.v + v.v1c, v.v0 + u.vc, CK_FF);
    const uint32_t c1 = __vimax3_s16x2(v.v + u.v1c, u.v0 + v.vc, CK_FF);
    return __vmins2(c0, c1);
}

// sobel-plane cost + (raw-plane cost >> 2)
__device__ __forceinline__ uint32_t bt_pixel_cost(const BtOps& us, const BtOps& ur, const BtOps& vs, const BtOps& vr) {
    const uint32_t cr = ((bt_cost255(ur, vr) - CK_FF) >> 2) & 0x3fff3fffu;
    return bt_cost255(us, vs) + cr - CK_FF;
}

__global__ void __launch_bounds__(CK_THREADS, 1) sgbm_cost_kernel(const uint2* __restrict__ pre, int n_pairs, int W, int H, int W1,
                                                                 int band_rows, uint32_t* __restrict__ Cvol) {
    extern __shared__ __align__(16) unsigned char ck_smem[];
    uint2* sOp = reinterpret_cast<uint2*>(ck_smem);                          // [CK_COLS left | CK_NPOS right]
    uint32_t* sPd = reinterpret_cast<uint32_t*>(sOp + CK_COLS + CK_NPOS);    // [CK_NG groups][CK_GC columns][48] (+16 pad)
    uint32_t* sRing = sPd + CK_NG * CK_GSTR;                                 // [9][CK_NG - 2 groups][CK_GC columns][48] (+16 pad)
    const int t = threadIdx.x, g = t / CK_Q, q = t - g * CK_Q;
    const int x1s = blockIdx.x * CK_CW, pair = blockIdx.z;
    const int y0 = blockIdx.y * band_rows, y1 = min(y0 + band_rows, H);
    if (y0 >= H) return;
    const uint2* preL = pre + (size_t)pair * H * W;
    const uint2* preR = pre + (size_t)(n_pairs + pair) * H * W;
    const int xfirst = x1s - CK_GC;         // W1-coordinate of exchange-buffer column 0
    const int pb = xfirst + SG_D - 95 - 3;  // image position of right-row slot 0
    const int xg = xfirst + g * CK_GC;      // first column of this thread
    const int nvalid = xg < 0 ? 0 : min(CK_GC, W1 - xg);  // columns of this group inside the image (groups are CK_GC-aligned)
    const bool out_group = g >= 1 && g <= CK_NG - 2 && nvalid > 0;
    uint32_t* Cout = Cvol + (size_t)pair * H * W1 * SG_NDP;

    // one operand per thread per row (CK_COLS + CK_NPOS = 348 <= CK_THREADS): fetched early, parked in shared memory late
    const uint2* op_src = nullptr;
    if (t < CK_COLS) op_src = preL + min(max(xfirst + t, 0), W1 - 1) + SG_D;
    else if (t < CK_COLS + CK_NPOS) op_src = preR + min(max(pb + t - CK_COLS, 0), W - 1);
    auto fetch_op = [&](int row) {
        return op_src ? op_src[(size_t)min(max(row, 0), H - 1) * W] : make_uint2(0, 0);
    };

    // exchange-buffer word offsets of the 16 columns under this thread's horizontal windows (borders replicated)
    int woff[CK_WIN];
#pragma unroll
    for (int k = 0; k < CK_WIN; ++k) {
        const int col = min(max(xg - SG_R + k, 0), W1 - 1) - xfirst;
        woff[k] = (col / CK_GC) * CK_GSTR + (col % CK_GC) * SG_NDP + 2 * q;
    }
    uint32_t* ring = sRing + (g - 1) * CK_GSTR + 2 * q;
    if (out_group)
        for (int s = 0; s < 2 * SG_R + 1; ++s)
#pragma unroll
            for (int i = 0; i < CK_GC; ++i) *reinterpret_cast<uint2*>(ring + s * CK_RSLOT + i * SG_NDP) = make_uint2(0, 0);
    uint32_t run0[CK_GC], run1[CK_GC];
#pragma unroll
    for (int i = 0; i < CK_GC; ++i) run0[i] = run1[i] = 0;
    int slot = 0;
    const int r_first = y0 - SG_R, r_last = y1 - 1 + SG_R;
    if (t < CK_COLS + CK_NPOS) sOp[t] = fetch_op(r_first);
    for (int row = r_first; row <= r_last; ++row) {
        __syncthreads();
        // ---- phase B: pixel cost of this thread's columns -> exchange buffer ----
        if (nvalid > 0) {
            const uint2* Lp = sOp + g * CK_GC;
            const uint2* Rp = sOp + CK_COLS + (xg - xfirst) + (SG_D - pb + xfirst) - 4 * q;
            // Rp[i] = position of disparity 4q at column i; Rp[i - e] = disparity 4q + e
            uint2 w1 = Rp[-1], w2 = Rp[-2], w3 = Rp[-3];
            BtOps hs2 = bt_expand(w2.x, w3.x), hr2 = bt_expand(w2.y, w3.y);  // word-0 operands of column -2 = word 1 of column 0
            BtOps hs1 = bt_expand(w1.x, w2.x), hr1 = bt_expand(w1.y, w2.y);  // ... of column -1
            uint32_t* pdst = sPd + g * CK_GSTR + 2 * q;
#pragma unroll
            for (int i = 0; i < CK_GC; ++i) {
                if (i < nvalid) {
                    const uint2 w0 = Rp[i], lp = Lp[i];
                    const BtOps vs = bt_expand(w0.x, w1.x), vr = bt_expand(w0.y, w1.y);
                    const BtOps us = bt_expand(lp.x, lp.x), ur = bt_expand(lp.y, lp.y);
                    uint2 pd;
                    pd.x = bt_pixel_cost(us, ur, vs, vr);
                    pd.y = bt_pixel_cost(us, ur, hs2, hr2);
                    *reinterpret_cast<uint2*>(pdst + i * SG_NDP) = pd;
                    hs2 = hs1;
                    hr2 = hr1;
                    hs1 = vs;
                    hr1 = vr;
                    w1 = w0;
                }
            }
        }
        __syncthreads();
        // ---- phase C: horizontal window in registers, vertical window through the ring, running sum, store ----
        const uint2 next_op = row < r_last ? fetch_op(row + 1) : make_uint2(0, 0);  // in flight during phase C
        if (out_group) {
            const int yo = row - SG_R;
            uint2 win[CK_WIN];
#pragma unroll
            for (int k = 0; k < CK_WIN; ++k) win[k] = *reinterpret_cast<const uint2*>(sPd + woff[k]);
            uint2 hsum = make_uint2(0, 0);
#pragma unroll
            for (int k = 0; k < 2 * SG_R + 1; ++k) {
                hsum.x += win[k].x;
                hsum.y += win[k].y;
            }
#pragma unroll
            for (int i = 0; i < CK_GC; ++i) {
                if (i < nvalid) {
                    uint2* rp = reinterpret_cast<uint2*>(ring + slot * CK_RSLOT + i * SG_NDP);
                    const uint2 old = *rp;
                    *rp = hsum;
                    run0[i] += hsum.x - old.x;
                    run1[i] += hsum.y - old.y;
                    if (yo >= y0) *reinterpret_cast<uint2*>(Cout + ((size_t)yo * W1 + xg + i) * SG_NDP + 2 * q) = make_uint2(run0[i], run1[i]);
                    if (i + 1 < CK_GC) {
                        hsum.x += win[i + 2 * SG_R + 1].x - win[i].x;
                        hsum.y += win[i + 2 * SG_R + 1].y - win[i].y;
                    }
                }
            }
        }
        slot = slot == 2 * SG_R ? 0 : slot + 1;
        if (t < CK_COLS + CK_NPOS) sOp[t] = next_op;  // phase B of this row finished reading before the barrier above
    }
}

// ---- TMA bulk-copy plumbing  ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends for a hardware-defined time slice per call; a copy that never lands (a bad address) traps after
// 2^24 slices instead of hanging the device.  Takes a shared-window address.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "SG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SG_DONE;\n"
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 q, n, 16777216;\n"
        "@q bra SG_WAIT;\n"
        "trap;\n"
        "SG_DONE:\n"
        "}\n" ::"r"(bar_addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_a(smem_u32(bar), parity); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// One SGM step on a warp: lanes 0..23 hold disparities 4l..4l+3 as two s16x2 words (a0 = d, d+1; a1 = d+2, d+3).
//   L(d) = C(d) + min(Lp(d), Lp(d-1) + P1, Lp(d+1) + P1, m + P2) - m,   m = min_k Lp(k)
// `mm` carries m in both halves of a word (the CREDUX of a word whose halves are equal is monotone in that value,
// so the reduction returns the next broadcast word directly).  Lanes 24..31 are padding: they are fed the cost
// SG_CPAD so their values stay in [28000, 28000 + P2] -- above every reachable m + P2, below s16 overflow -- which
// gives lane 23 its "no d+1 neighbour" and, through the rotating shuffle, lane 0 its "no d-1 neighbour" for free.
// ---------------------------------------------------------------------------------------------------------------
#define SG_CPAD 0x6d606d60u  // 28000 | 28000 << 16

__device__ __forceinline__ void sgm_step(uint32_t& a0, uint32_t& a1, uint32_t& mm, uint32_t c0, uint32_t c1, uint32_t P1b,
                                         uint32_t P2b, int src_up) {
    const uint32_t up = __shfl_sync(0xffffffffu, a1, src_up);  // lane - 1 (lane 0 reads padding lane 31)
    const uint32_t dn = __shfl_down_sync(0xffffffffu, a0, 1);
    const uint32_t lm0 = __byte_perm(up, a0, 0x5432);  // (L[4l-1], L[4l])
    const uint32_t mid = __byte_perm(a0, a1, 0x5432);  // (L[4l+1], L[4l+2])
    const uint32_t lp1 = __byte_perm(a1, dn, 0x5432);  // (L[4l+3], L[4l+4])
    const uint32_t mp2 = mm + P2b;
    const uint32_t t0 = __vmins2(__vmins2(__vadd2(__vmins2(lm0, mid), P1b), mp2), a0);
    const uint32_t t1 = __vmins2(__vmins2(__vadd2(__vmins2(mid, lp1), P1b), mp2), a1);
    a0 = c0 + t0 - mm;  // halves stay in [0, 32767]: plain 32-bit arithmetic is exact on the packed pairs
    a1 = c1 + t1 - mm;
    const uint32_t w = __vmins2(a0, a1);
    mm = (uint32_t)__reduce_min_sync(0xffffffffu, (int)__vmins2(w, __byte_perm(w, w, 0x1032)));
}

// Two paths per warp: each half-warp owns one path, lane sl = lane & 15 holds disparities 6 sl .. 6 sl + 5 as three
// s16x2 words -- all 32 lanes carry data (the one-path layout above idles 8), and the per-step bookkeeping of a warp
// (barrier wait, loads, stores, pointer updates) is shared by two paths.  The minimum of each half comes from two
// full-warp CREDUX with the other half neutralised.
__device__ __forceinline__ void sgm_step2(uint32_t (&a)[3], uint32_t& mm, const uint32_t (&c)[3], uint32_t P1b, uint32_t P2b,
                                          int sl, bool upper_half) {
    uint32_t up = __shfl_up_sync(0xffffffffu, a[2], 1, 16);
    uint32_t dn = __shfl_down_sync(0xffffffffu, a[0], 1, 16);
    if (sl == 0) up = SG_BIG2;
    if (sl == 15) dn = SG_BIG2;
    const uint32_t lm0 = __byte_perm(up, a[0], 0x5432);    // (L[6s-1], L[6s])
    const uint32_t m01 = __byte_perm(a[0], a[1], 0x5432);  // (L[6s+1], L[6s+2])
    const uint32_t m12 = __byte_perm(a[1], a[2], 0x5432);  // (L[6s+3], L[6s+4])
    const uint32_t lp2 = __byte_perm(a[2], dn, 0x5432);    // (L[6s+5], L[6s+6])
    const uint32_t mp2 = mm + P2b;
    const uint32_t t0 = __vmins2(__vmins2(__vadd2(__vmins2(lm0, m01), P1b), mp2), a[0]);
    const uint32_t t1 = __vmins2(__vmins2(__vadd2(__vmins2(m01, m12), P1b), mp2), a[1]);
    const uint32_t t2 = __vmins2(__vmins2(__vadd2(__vmins2(m12, lp2), P1b), mp2), a[2]);
    a[0] = c[0] + t0 - mm;
    a[1] = c[1] + t1 - mm;
    a[2] = c[2] + t2 - mm;
    const uint32_t w = __vimin3_s16x2(a[0], a[1], a[2]);
    const uint32_t v = __vmins2(w, __byte_perm(w, w, 0x1032));
    const uint32_t mlo = (uint32_t)__reduce_min_sync(0xffffffffu, (int)(upper_half ? SG_BIG2 : v));
    const uint32_t mhi = (uint32_t)__reduce_min_sync(0xffffffffu, (int)(upper_half ? v : SG_BIG2));
    mm = upper_half ? mhi : mlo;
}

// ---------------------------------------------------------------------------------------------------------------
// K20: the three paths that come from the row above.  blockIdx.y: 0 = from (x-1, y-1), 1 = from (x, y-1),
// 2 = from (x+1, y-1).  Warp k starts at column k of row 0; a diagonal path that leaves the image re-enters on the
// other side with a fresh (zero) predecessor, which is what OpenCV's zero-initialised border columns give.
// The three sweeps of one pair run concurrently and touch row y at about the same time, so C is read with the
// default policy (two of the three reads hit L2) while the path volumes are streamed out.
// ---------------------------------------------------------------------------------------------------------------
#define VT_NST 16   // rows in flight per CTA: 16 x 1536 B
#define VT_PATHS 8  // paths per CTA
#define VT_WARPS 4  // consumer warps per CTA: two paths each (+ 1 producer warp)

template <int DX>
__device__ __forceinline__ void vertical_cta(const unsigned char* __restrict__ gC, unsigned char* __restrict__ gL, int k0, int W1,
                                             int H, uint32_t P1b, uint32_t P2b, unsigned char* ring, uint64_t* full,
                                             uint64_t* empty) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nact = min(VT_PATHS, W1 - k0);  // paths of this CTA
    if (threadIdx.x == 0) {
        for (int s = 0; s < VT_NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], (nact + 1) / 2);  // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (w == VT_WARPS) {
        // producer warp: one lane streams row y of the CTA's nact adjacent columns (1 or 2 pieces when the range wraps)
        if (lane == 0) {
            int xb = k0;
            for (int y = 0; y < H; ++y) {
                const int s = y % VT_NST;
                if (y >= VT_NST) mbar_wait(&empty[s], ((y / VT_NST) - 1) & 1);
                const int n1 = min(nact, W1 - xb);
                mbar_expect_tx(&full[s], (uint32_t)nact * (SG_D * 2));
                bulk_g2s(ring + s * (VT_PATHS * SG_D * 2), gC + ((size_t)y * W1 + xb) * (SG_D * 2), (uint32_t)n1 * (SG_D * 2), &full[s]);
                if (n1 < nact)
                    bulk_g2s(ring + s * (VT_PATHS * SG_D * 2) + n1 * (SG_D * 2), gC + (size_t)y * W1 * (SG_D * 2),
                             (uint32_t)(nact - n1) * (SG_D * 2), &full[s]);
                xb += DX;
                if (DX > 0 && xb == W1) xb = 0;
                if (DX < 0 && xb < 0) xb = W1 - 1;
            }
        }
        return;
    }
    // consumer warp w walks paths k0 + 2w (lower half-warp) and k0 + 2w + 1 (upper half-warp)
    if (2 * w >= nact) return;
    const bool upper = lane >= 16;
    const int sl = lane & 15, pidx = 2 * w + (upper ? 1 : 0);
    const bool pvalid = pidx < nact;  // the last CTA may hold an odd number of paths
    uint32_t a[3] = {0u, 0u, 0u}, mm = 0;
    int x = k0 + pidx;
    // everything the row loop touches is a 32-bit shared address plus an immediate, or a pointer advanced by a constant
    const uint32_t ring_a = smem_u32(ring) + pidx * (SG_D * 2) + sl * 12, full_a = smem_u32(full), empty_a = smem_u32(empty);
    unsigned char* gp = gL + (size_t)x * (SG_D * 2) + sl * 12;
    const long long rowstep = (long long)(W1 + DX) * (SG_D * 2), wrapfix = (long long)W1 * (SG_D * 2);
#pragma unroll 2
    for (int y = 0; y < H; ++y) {
        const uint32_t s = y & (VT_NST - 1), par = (y / VT_NST) & 1;
        mbar_wait_a(full_a + s * 8, par);
        uint32_t c[3] = {0u, 0u, 0u};
        if (pvalid) {
            const uint32_t ra = ring_a + s * (VT_PATHS * SG_D * 2);
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c[0]) : "r"(ra));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c[1]) : "r"(ra + 4));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c[2]) : "r"(ra + 8));
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_a + s * 8) : "memory");
        if (DX != 0 && x == (DX > 0 ? 0 : W1 - 1)) {
            a[0] = a[1] = a[2] = 0u;
            mm = 0;
        }
        sgm_step2(a, mm, c, P1b, P2b, sl, upper);
        if (pvalid) {
            uint32_t* o = reinterpret_cast<uint32_t*>(gp);
            __stcs(o, a[0]);
            __stcs(o + 1, a[1]);
            __stcs(o + 2, a[2]);
        }
        x += DX;
        gp += rowstep;
        if (DX > 0 && x == W1) {
            x = 0;
            gp -= wrapfix;
        }
        if (DX < 0 && x < 0) {
            x = W1 - 1;
            gp += wrapfix;
        }
    }
}

__global__ void __launch_bounds__((VT_WARPS + 1) * 32) sgbm_vertical_kernel(const uint32_t* __restrict__ Cvol, uint32_t* __restrict__ L0v,
                                                                           uint32_t* __restrict__ L1v, uint32_t* __restrict__ L2v,
                                                                           int W1, int H, int P1, int P2) {
    __shared__ __align__(128) unsigned char ring[VT_NST * VT_PATHS * SG_D * 2];
    __shared__ uint64_t full[VT_NST], empty[VT_NST];
    const int k0 = blockIdx.x * VT_PATHS;
    const int dir = blockIdx.y, pair = blockIdx.z;
    const size_t vol = (size_t)pair * H * W1 * (SG_D * 2);
    const unsigned char* gC = reinterpret_cast<const unsigned char*>(Cvol) + vol;
    const uint32_t P1b = bcast16(P1), P2b = bcast16(P2);
    if (dir == 0) vertical_cta<1>(gC, reinterpret_cast<unsigned char*>(L0v) + vol, k0, W1, H, P1b, P2b, ring, full, empty);
    else if (dir == 1) vertical_cta<0>(gC, reinterpret_cast<unsigned char*>(L1v) + vol, k0, W1, H, P1b, P2b, ring, full, empty);
    else vertical_cta<-1>(gC, reinterpret_cast<unsigned char*>(L2v) + vol, k0, W1, H, P1b, P2b, ring, full, empty);
}

// ---------------------------------------------------------------------------------------------------------------
// K21: one warp per image row, two launches (each fits one resident wave).  Pass 1 walks left->right (path 0) and folds the four finished paths into
// S4 = sat16(L0+L1+L2+L3) (written over the first path volume); pass 2 walks right->left (the fifth path of MODE_SGBM's
// single pass), adds it and reduces every column to a record {min S, argmin, not-unique, S[best-1], S[best+1]}.
// The per-column epilogue (right-image map, sub-pixel fit, LR check) then runs lane-parallel over x: OpenCV's
// "first writer in descending x wins a cost tie" becomes an atomicMin on the key (minS << 12 | 4095 - x).
// ---------------------------------------------------------------------------------------------------------------
#define RS_CH 8                         // columns per bulk copy: 8 x 192 B = 1536 contiguous bytes per volume
#define RS_NST 4                        // ring stages per warp
#define RS_ARR (RS_CH * SG_D * 2)       // bytes of one volume's chunk
#define RS_FWD_WARP (RS_NST * 4 * RS_ARR)
#define RS_BWD_STAGES (RS_NST * 2 * RS_ARR)

// The row sweeps stream each image row exactly once, 192 B per column per volume.  Issued as per-lane loads that is a
// trickle of small requests from thousands of concurrent rows (poor DRAM page locality: measured 3.4 TB/s at best);
// here one lane moves 1.5 KB chunks per volume with cp.async.bulk into a per-warp ring, completion on an mbarrier.
__global__ void __launch_bounds__(SG_HW * 32) sgbm_row_forward_kernel(const uint32_t* __restrict__ Cvol, uint32_t* __restrict__ L0v,
                                                                     const uint32_t* __restrict__ L1v,
                                                                     const uint32_t* __restrict__ L2v, int W1, int H,
                                                                     SgParams p) {
    extern __shared__ __align__(128) unsigned char rs_smem[];
    __shared__ uint64_t bars[SG_HW][RS_NST];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int y = blockIdx.x * SG_HW + wp, pair = blockIdx.y;
    if (y >= H) return;
    const size_t rowoff = ((size_t)pair * H + y) * W1 * (SG_D * 2);
    const unsigned char* gC = reinterpret_cast<const unsigned char*>(Cvol) + rowoff;
    unsigned char* gT = reinterpret_cast<unsigned char*>(L0v) + rowoff;
    const unsigned char* gB = reinterpret_cast<const unsigned char*>(L1v) + rowoff;
    const unsigned char* gD = reinterpret_cast<const unsigned char*>(L2v) + rowoff;
    unsigned char* wbase = rs_smem + (size_t)wp * RS_FWD_WARP;
    uint64_t* bar = bars[wp];
    const bool active = lane < 24;
    const int src_up = (lane + 31) & 31;
    const uint32_t P1b = bcast16(p.P1), P2b = bcast16(p.P2);
    const uint32_t SAT = 0x7fff7fffu;
    const int n_chunks = (W1 + RS_CH - 1) / RS_CH;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < RS_NST; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int j) {  // lane 0
        const int s = j % RS_NST, x0 = j * RS_CH;
        const uint32_t bytes = (uint32_t)min(RS_CH, W1 - x0) * (SG_D * 2);
        unsigned char* st = wbase + s * 4 * RS_ARR;
        const size_t go = (size_t)x0 * (SG_D * 2);
        mbar_expect_tx(&bar[s], 4 * bytes);
        bulk_g2s(st, gC + go, bytes, &bar[s]);
        bulk_g2s(st + RS_ARR, gT + go, bytes, &bar[s]);
        bulk_g2s(st + 2 * RS_ARR, gB + go, bytes, &bar[s]);
        bulk_g2s(st + 3 * RS_ARR, gD + go, bytes, &bar[s]);
    };
    if (lane == 0)
        for (int j = 0; j < RS_NST - 1 && j < n_chunks; ++j) issue(j);
    uint32_t a0 = active ? 0u : SG_BIG2, a1 = a0, mm = 0;
    for (int j = 0; j < n_chunks; ++j) {
        const int s = j % RS_NST, x0 = j * RS_CH, n = min(RS_CH, W1 - x0);
        mbar_wait(&bar[s], (j / RS_NST) & 1);
        uint2* sC = reinterpret_cast<uint2*>(wbase + s * 4 * RS_ARR) + lane;
        uint2* sT = sC + RS_ARR / 8;
        const uint2* sB = sT + RS_ARR / 8;
        const uint2* sD = sB + RS_ARR / 8;
#pragma unroll
        for (int i = 0; i < RS_CH; ++i) {
            if (i < n) {
                uint2 c = make_uint2(SG_CPAD, SG_CPAD), l0 = make_uint2(0, 0), l1 = l0, l2 = l0;
                if (active) {
                    c = sC[i * 24];
                    l0 = sT[i * 24];
                    l1 = sB[i * 24];
                    l2 = sD[i * 24];
                }
                sgm_step(a0, a1, mm, c.x, c.y, P1b, P2b, src_up);
                // three paths <= 3 * (15309 + P2) < 65536: exact in u16; then saturate like CostType
                uint2 o;
                o.x = __vminu2(__vminu2(l0.x + l1.x + l2.x, SAT) + a0, SAT);
                o.y = __vminu2(__vminu2(l0.y + l1.y + l2.y, SAT) + a1, SAT);
                if (active) sT[i * 24] = o;
            }
        }
        fence_async_smem();  // the S4 chunk was written through the generic proxy; the bulk store reads it through the async one
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(gT + (size_t)x0 * (SG_D * 2), wbase + s * 4 * RS_ARR + RS_ARR, (uint32_t)n * (SG_D * 2));
            bulk_commit();
            if (j + RS_NST - 1 < n_chunks) {
                bulk_wait_read<1>();  // the store of chunk j-1 has drained its stage: refill it
                issue(j + RS_NST - 1);
            }
        }
    }
    if (lane == 0) bulk_wait_read<0>();
}

__global__ void __launch_bounds__(SG_HW * 32) sgbm_row_backward_kernel(const uint32_t* __restrict__ Cvol,
                                                                      const uint32_t* __restrict__ L0v, int W, int W1, int H,
                                                                      SgParams p, uint2* __restrict__ rec_all,
                                                                      int16_t* __restrict__ disp_raw) {
    extern __shared__ __align__(128) unsigned char rs_smem[];
    __shared__ uint64_t bars[SG_HW][RS_NST];
    __shared__ uint2 s_excl[8];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int y = blockIdx.x * SG_HW + wp, pair = blockIdx.y;
    if (threadIdx.x < 8) {
        // 16-bit lanes of (d, d+1 | d+2, d+3) to ignore when the winner sits at local index e = t - 1
        const uint2 tab[8] = {{0x0000ffffu, 0u}, {0xffffffffu, 0u}, {0xffffffffu, 0x0000ffffu}, {0xffff0000u, 0xffffffffu},
                              {0u, 0xffffffffu}, {0u, 0xffff0000u}, {0u, 0u}, {0u, 0u}};
        s_excl[threadIdx.x] = tab[threadIdx.x];
    }
    __syncthreads();
    if (y >= H) return;
    unsigned char* wbase = rs_smem + (size_t)wp * RS_BWD_STAGES;
    uint32_t* s_key = reinterpret_cast<uint32_t*>(wbase);  // the right-image map reuses the ring once the sweep is over
    uint64_t* bar = bars[wp];
    const size_t rowoff = ((size_t)pair * H + y) * W1 * (SG_D * 2);
    const unsigned char* gC = reinterpret_cast<const unsigned char*>(Cvol) + rowoff;
    const unsigned char* gT = reinterpret_cast<const unsigned char*>(L0v) + rowoff;
    uint2* rec = rec_all + ((size_t)pair * H + y) * W1;
    const bool active = lane < 24;
    const int src_up = (lane + 31) & 31;
    const uint32_t P1b = bcast16(p.P1), P2b = bcast16(p.P2);
    const uint32_t SAT = 0x7fff7fffu;
    const int n_chunks = (W1 + RS_CH - 1) / RS_CH;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < RS_NST; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int j) {  // lane 0; chunk j counts from the right end of the row
        const int s = j % RS_NST, x0 = (n_chunks - 1 - j) * RS_CH;
        const uint32_t bytes = (uint32_t)min(RS_CH, W1 - x0) * (SG_D * 2);
        unsigned char* st = wbase + s * 2 * RS_ARR;
        const size_t go = (size_t)x0 * (SG_D * 2);
        mbar_expect_tx(&bar[s], 2 * bytes);
        bulk_g2s(st, gC + go, bytes, &bar[s]);
        bulk_g2s(st + RS_ARR, gT + go, bytes, &bar[s]);
    };
    if (lane == 0)
        for (int j = 0; j < RS_NST - 1 && j < n_chunks; ++j) issue(j);  // chunk j >= 1 tops the ring up with chunk j + NST - 2
    // ---- right -> left, per-column record ----
    {
        uint32_t a0 = active ? 0u : SG_BIG2, a1 = a0, mm = 0;
        const uint32_t uq = 100 - p.uniq;
        const uint32_t uq_magic = (uint32_t)((1ull << 32) / uq) + 1;  // floor(n / uq) = umulhi(n, magic) while n * uq < 2^32
        const uint32_t d0 = 4 * lane;
        const uint32_t dk0 = d0, dk1 = d0 + 1, dk2 = d0 + 2, dk3 = d0 + 3;  // PRMT source: byte 4 = d, bytes 5..7 = 0
        const uint32_t lane_or = active ? 0u : 0xffffffffu;
        for (int j = 0; j < n_chunks; ++j) {
            const int s = j % RS_NST, x0 = (n_chunks - 1 - j) * RS_CH, n = min(RS_CH, W1 - x0);
            // the stage of chunk j-1 was released by the __syncwarp that closed the previous iteration: refill it
            if (lane == 0 && j >= 1 && j + RS_NST - 2 < n_chunks) issue(j + RS_NST - 2);
            mbar_wait(&bar[s], (j / RS_NST) & 1);
            const uint2* sC = reinterpret_cast<const uint2*>(wbase + s * 2 * RS_ARR) + lane;
            uint2* sT = reinterpret_cast<uint2*>(wbase + s * 2 * RS_ARR + RS_ARR) + lane;
            // phase A: the recurrence for the whole chunk (descending x); S = sat16(S4 + L_r) replaces S4 in the stage
            uint32_t S0[RS_CH], S1[RS_CH];
#pragma unroll
            for (int ii = 0; ii < RS_CH; ++ii) {
                const int i = RS_CH - 1 - ii;
                S0[i] = S1[i] = SAT;
                if (i < n) {
                    uint2 c = make_uint2(SG_CPAD, SG_CPAD), tt = make_uint2(SAT, SAT);  // padding lanes: S = 32767, d >= 96
                    if (active) {
                        c = sC[i * 24];
                        tt = sT[i * 24];
                    }
                    sgm_step(a0, a1, mm, c.x, c.y, P1b, P2b, src_up);
                    S0[i] = __vminu2(tt.x + a0, SAT);
                    S1[i] = __vminu2(tt.y + a1, SAT);
                    if (active) sT[i * 24] = make_uint2(S0[i], S1[i]);  // u16 S[d] of column i at d = 4 lane + k
                }
            }
            // phase B: winner-take-all per column.  The columns are independent, so every warp-wide operation is
            // issued for all of them back to back and their latencies overlap instead of adding up.
            // key = S << 8 | d, one PRMT per element: bytes {d, S.lo, S.hi, 0}
            uint32_t key[RS_CH];
#pragma unroll
            for (int i = 0; i < RS_CH; ++i)
                key[i] = min(min(__byte_perm(S0[i], dk0, 0x5104), __byte_perm(S0[i], dk1, 0x5324)),
                             min(__byte_perm(S1[i], dk2, 0x5104), __byte_perm(S1[i], dk3, 0x5324)));
#pragma unroll
            for (int i = 0; i < RS_CH; ++i) key[i] = __reduce_min_sync(0xffffffffu, key[i]);
            bool nuq[RS_CH];
#pragma unroll
            for (int i = 0; i < RS_CH; ++i) {
                const uint32_t minS = key[i] >> 8, bestd = key[i] & 0xff;
                // uniqueness: some S(d) * (100 - u) < minS * 100 with |d - best| > 1  <=>  S(d) < ceil(minS * 100 / (100 - u))
                const uint32_t q = __umulhi(minS * 100u + uq - 1, uq_magic);
                const uint2 ex = s_excl[min(bestd - d0 + 1u, 6u)];
                const uint32_t z = __vminu2(S0[i] | ex.x | lane_or, S1[i] | ex.y | lane_or);
                nuq[i] = min(z & 0xffffu, z >> 16) < q;
            }
#pragma unroll
            for (int i = 0; i < RS_CH; ++i) nuq[i] = __any_sync(0xffffffffu, nuq[i]);
            __syncwarp();  // the S chunk is complete in shared memory
            {
                // lane i finishes column i: S(best - 1), S(best + 1) come from the chunk in shared memory
                uint32_t k = key[0];
                bool nq = nuq[0];
#pragma unroll
                for (int i = 1; i < RS_CH; ++i)
                    if (lane == i) {
                        k = key[i];
                        nq = nuq[i];
                    }
                if (lane < n) {
                    const uint32_t bestd = k & 0xff;
                    const uint16_t* Sc = reinterpret_cast<const uint16_t*>(wbase + s * 2 * RS_ARR + RS_ARR) + lane * SG_D;
                    const uint32_t Sm = Sc[max((int)bestd - 1, 0)], Sp = Sc[min(bestd + 1, (uint32_t)SG_D - 1)];
                    rec[x0 + lane] = make_uint2((k >> 8) | bestd << 16 | (nq ? 0x80000000u : 0u), Sm | Sp << 16);
                }
            }
            __syncwarp();  // every lane is done reading this stage
        }
    }
    __syncwarp();
    // ---- lane-parallel epilogue: right-image map (every bulk copy has landed and been consumed: the ring is free) ----
    for (int i = lane; i < W; i += 32) s_key[i] = 0xffffffffu;
    __syncwarp();
    for (int xx = lane; xx < W1; xx += 32) {
        const uint32_t r0 = rec[xx].x;
        const uint32_t minS = r0 & 0xffffu, bestd = (r0 >> 16) & 0xffu;
        // OpenCV: if (disp2cost[x2] > minS) with disp2cost initialised to MAX_COST
        if (!(r0 >> 31) && minS < SG_MAXCOST) atomicMin(&s_key[xx + SG_D - bestd], minS << 12 | (4095u - xx));
    }
    __syncwarp();
    // ---- sub-pixel fit, left-right consistency, write-out ----
    int16_t* out = disp_raw + ((size_t)pair * H + y) * W;
    for (int xi = lane; xi < W; xi += 32) {
        int d1 = SG_INVALID;
        if (xi >= SG_D) {
            const uint2 r = rec[xi - SG_D];
            if (!(r.x >> 31)) {
                const int minS = r.x & 0xffffu, bestd = (r.x >> 16) & 0xffu;
                d1 = bestd * 16;
                if (0 < bestd && bestd < SG_D - 1) {
                    const int Sm = r.y & 0xffffu, Sp = r.y >> 16;
                    const int denom2 = max(Sm + Sp - 2 * minS, 1);
                    d1 += ((Sm - Sp) * 16 + denom2) / (denom2 * 2);
                }
                const int dl = d1 >> 4, dh = (d1 + 15) >> 4;
                const int xa = xi - dl, xb = xi - dh;
                if (xa >= 0 && xb >= 0) {  // xa, xb <= xi < W
                    const uint32_t ka = s_key[xa], kb = s_key[xb];
                    if (ka != 0xffffffffu && kb != 0xffffffffu) {
                        // disparity stored in the right-image map = x_left - x_right
                        const int ea = (int)(4095u - (ka & 0xfffu)) + SG_D - xa, eb = (int)(4095u - (kb & 0xfffu)) + SG_D - xb;
                        if (abs(ea - dl) > p.disp12 && abs(eb - dh) > p.disp12) d1 = SG_INVALID;
                    }
                }
            }
        }
        out[xi] = (int16_t)d1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K22: cv::medianBlur(ksize 3) on CV_16S (replicated border)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap(int& a, int& b) {
    const int lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

__global__ void __launch_bounds__(256) sgbm_median_kernel(const int16_t* __restrict__ in, int16_t* __restrict__ out, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int16_t* base = in + (size_t)blockIdx.z * H * W;
    int v[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int16_t* r = base + (size_t)min(max(y - 1 + dy, 0), H - 1) * W;
#pragma unroll
        for (int k = 0; k < 3; ++k) v[dy * 3 + k] = r[min(max(x - 1 + k, 0), W - 1)];
    }
    // median-of-9 exchange network
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
    cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]);
    cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
    cswap(v[4], v[2]);
    out[((size_t)blockIdx.z * H + y) * W + x] = (int16_t)v[4];
}

// ---------------------------------------------------------------------------------------------------------------
// K23: cv::filterSpeckles(disp, newVal = -16, maxSpeckleSize, maxDiff) as union-find connected components.
// Two valid 4-neighbours are connected when |a - b| <= maxDiff; components are equivalence classes, so the result
// does not depend on OpenCV's scan order.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int* L, int a) {
    int r = a;
    while (true) {
        const int pr = L[r];
        if (pr == r) break;
        r = pr;
    }
    return r;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&L[a], b);  // a > b: hang root a under b
        if (old == a) return;
        a = old;
    }
}

// One warp per row: every pixel is labelled with the first pixel of its horizontal run (ballot + clz, carried across
// 32-pixel chunks), so the union-find forest starts with one node per run instead of one per pixel.
__global__ void __launch_bounds__(256) speckle_runs_kernel(const int16_t* __restrict__ d, int* __restrict__ label,
                                                           int* __restrict__ size, int* __restrict__ runlen, int W, int H,
                                                           int n_rows, int max_diff) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);  // row over all images
    if (row >= n_rows) return;
    const int16_t* p = d + (size_t)row * W;
    const int base = (row % H) * W;  // labels are pixel indices inside their own image
    int* L = label + (size_t)row * W;
    int* S = size + (size_t)row * W;
    int* R = runlen + (size_t)row * W;
    int carry_start = -1;       // start of the run that reaches into this chunk from the left
    int prev_last = SG_INVALID; // value of the last pixel of the previous chunk
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        const int v = x < W ? p[x] : SG_INVALID;
        int left = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) left = prev_last;
        const bool valid = v != SG_INVALID;
        const bool conn = valid && left != SG_INVALID && abs(v - left) <= max_diff && x > 0;
        const unsigned starts = __ballot_sync(0xffffffffu, valid && !conn);
        const unsigned below = starts & (0xffffffffu >> (31 - lane));
        const int start = below ? x0 + 31 - __clz(below) : carry_start;
        // is this pixel the last of its run?  (next pixel not connected to it)
        int nv = __shfl_down_sync(0xffffffffu, v, 1);
        int nx_valid_conn;
        if (lane == 31 || x + 1 >= W) {
            const int nxt = x + 1 < W ? p[x + 1] : SG_INVALID;
            nv = nxt;
        }
        nx_valid_conn = valid && nv != SG_INVALID && abs(nv - v) <= max_diff && x + 1 < W;
        if (x < W) {
            L[x] = valid ? base + start : -1;
            S[x] = 0;
            R[x] = 0;
        }
        __syncwarp();
        if (valid && !nx_valid_conn) R[start] = x - start + 1;
        carry_start = __shfl_sync(0xffffffffu, start, 31);
        if (!__shfl_sync(0xffffffffu, (int)valid, 31)) carry_start = -1;
        prev_last = __shfl_sync(0xffffffffu, v, 31);
    }
}

// vertical links between runs; a link is skipped when the pixel to the left already made it
__global__ void __launch_bounds__(256) speckle_merge_kernel(const int16_t* __restrict__ d, int* __restrict__ label, int W, int H,
                                                            int max_diff) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y + 1;
    if (x >= W || y >= H) return;
    const size_t img = (size_t)blockIdx.z * H * W;
    const int16_t* p = d + img;
    int* L = label + img;
    const int i = y * W + x;
    const int v = p[i], u = p[i - W];
    if (v == SG_INVALID || u == SG_INVALID || abs(v - u) > max_diff) return;
    if (x > 0) {
        const int vl = p[i - 1], ul = p[i - W - 1];
        if (vl != SG_INVALID && ul != SG_INVALID && abs(v - vl) <= max_diff && abs(u - ul) <= max_diff &&
            abs(vl - ul) <= max_diff)
            return;
    }
    uf_union(L, L[i], L[i - W]);
}

// every run adds its length to its component's root
__global__ void __launch_bounds__(256) speckle_count_kernel(int* __restrict__ label, int* __restrict__ size,
                                                            const int* __restrict__ runlen, int hw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const size_t img = (size_t)blockIdx.y * hw;
    const int n = runlen[img + i];
    if (n == 0) return;
    int* L = label + img;
    const int r = uf_find(L, i);
    atomicAdd(&size[img + r], n);
    if (r != i) L[i] = r;  // path compression: the apply pass reaches the root in two hops (roots stay roots here)
}

__global__ void __launch_bounds__(256) speckle_apply_kernel(const int16_t* __restrict__ d, const int* __restrict__ label,
                                                            const int* __restrict__ size, int hw, int max_size,
                                                            int16_t* __restrict__ out, float* __restrict__ outf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const size_t img = (size_t)blockIdx.y * hw;
    int v = d[img + i];
    if (max_size > 0 && v != SG_INVALID) {
        int r = label[img + i];
        while (true) {
            const int pr = label[img + r];
            if (pr == r) break;
            r = pr;
        }
        if (size[img + r] <= max_size) v = SG_INVALID;
    }
    if (out) out[img + i] = (int16_t)v;
    if (outf) outf[img + i] = (float)v * 0.0625f;  // convertTo(CV_32F, 1/16): exact
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
int vslam_sgbm_init(vslam_ctx* ctx) {
    ctx->sgbm = (SgbmState*)calloc(1, sizeof(SgbmState));
    if (!ctx->sgbm) return VSLAM_E_INVALID;
    VSLAM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->sgbm->s_alt, cudaStreamNonBlocking));
    VSLAM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->sgbm->ev_fork, cudaEventDisableTiming));
    VSLAM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->sgbm->ev_join, cudaEventDisableTiming));
    return VSLAM_OK;
}

static void sgbm_release(SgbmState* s) {
    cudaFree(s->d_pre);
    cudaFree(s->d_C);
    for (int i = 0; i < 3; ++i) cudaFree(s->d_L[i]);
    cudaFree(s->d_raw);
    cudaFree(s->d_rec);
    cudaFree(s->d_med);
    cudaFree(s->d_label);
    cudaFree(s->d_size);
    cudaFree(s->d_runlen);
    cudaFree(s->d_img);
    cudaFree(s->d_out);
    cudaFree(s->d_outf);
    SgbmState keep = *s;
    memset(s, 0, sizeof(*s));
    s->stop_after = keep.stop_after;
    s->s_alt = keep.s_alt;
    s->ev_fork = keep.ev_fork;
    s->ev_join = keep.ev_join;
}

void vslam_sgbm_free(vslam_ctx* ctx) {
    if (!ctx->sgbm) return;
    sgbm_release(ctx->sgbm);
    if (ctx->sgbm->s_alt) cudaStreamDestroy(ctx->sgbm->s_alt);
    if (ctx->sgbm->ev_fork) cudaEventDestroy(ctx->sgbm->ev_fork);
    if (ctx->sgbm->ev_join) cudaEventDestroy(ctx->sgbm->ev_join);
    free(ctx->sgbm);
    ctx->sgbm = nullptr;
}

// scratch for `pairs` pairs of w x h images (grown on demand; the volumes dominate: 4 x H x W1 x 192 B per pair)
static int sgbm_reserve(vslam_ctx* ctx, int pairs, int w, int h) {
    SgbmState* s = ctx->sgbm;
    if (pairs <= s->cap_pairs && w == s->cap_w && h == s->cap_h) return VSLAM_OK;
    VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int keep = s->cap_w == w && s->cap_h == h ? s->cap_pairs : 0;
    sgbm_release(s);
    pairs = pairs > keep ? pairs : keep;
    const size_t px = (size_t)w * h, vol = (size_t)h * (w - SG_D) * SG_NDP * sizeof(uint32_t);
    s->pitch = (w + 15) & ~15;
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_pre, 2 * pairs * px * sizeof(uint2)));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_C, pairs * vol));
    for (int i = 0; i < 3; ++i) VSLAM_CUDA(ctx, cudaMalloc(&s->d_L[i], pairs * vol));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_raw, pairs * px * sizeof(int16_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_rec, pairs * px * sizeof(uint2)));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_med, pairs * px * sizeof(int16_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_label, pairs * px * sizeof(int32_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_size, pairs * px * sizeof(int32_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_runlen, pairs * px * sizeof(int32_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_img, 2 * (size_t)pairs * h * s->pitch));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_out, pairs * px * sizeof(int16_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&s->d_outf, pairs * px * sizeof(float)));
    s->cap_pairs = pairs;
    s->cap_w = w;
    s->cap_h = h;
    return VSLAM_OK;
}

extern "C" void vslam_sgbm_default_params(vslam_sgbm_params* p) {
    if (!p) return;
    // visual_odometry.cpp:163-164
    p->min_disparity = 0;
    p->num_disparities = 96;
    p->block_size = 9;
    p->P1 = 8 * 9 * 9;
    p->P2 = 32 * 9 * 9;
    p->disp12_max_diff = 1;
    p->pre_filter_cap = 63;
    p->uniqueness_ratio = 10;
    p->speckle_window_size = 100;
    p->speckle_range = 32;
}

static int sgbm_check(const vslam_sgbm_params* in, int w, int h, SgParams* out) {
    vslam_sgbm_params p;
    if (in) p = *in; else vslam_sgbm_default_params(&p);
    // the kernels are specialised for the reference's geometry: 96 disparities from 0, 9x9 block
    if (p.min_disparity != 0 || p.num_disparities != SG_D || p.block_size != 2 * SG_R + 1) return VSLAM_E_INVALID;
    if (h < 1 || w - SG_D <= SG_R) return VSLAM_E_INVALID;  // OpenCV raises for such images (stereosgbm.cpp:511)
    out->P1 = p.P1 > 0 ? p.P1 : 2;
    const int p2 = p.P2 > 0 ? p.P2 : 5;
    out->P2 = p2 > out->P1 + 1 ? p2 : out->P1 + 1;
    out->disp12 = p.disp12_max_diff > 0 ? p.disp12_max_diff : 1;
    out->uniq = p.uniqueness_ratio >= 0 ? p.uniqueness_ratio : 10;
    out->ftzero = (p.pre_filter_cap > 15 ? p.pre_filter_cap : 15) | 1;
    out->speckle_window = p.speckle_window_size;
    out->speckle_diff = 16 * p.speckle_range;
    // packed s16 arithmetic: real costs stay below 81*189 + P2 + P1 < 28000, padding lanes sit at 28000 .. 28000 + P2
    // and must survive + P1 without s16 overflow; ftzero <= 127 (u8 operands); the row sweep packs x into 12 bits
    if (out->P2 + out->P1 + 81 * 189 >= 28000 || 28000 + out->P2 + out->P1 > 32767 || out->ftzero > 127 ||
        out->uniq >= 100 || w - SG_D > 4096 || (size_t)w * 4 > RS_BWD_STAGES)
        return VSLAM_E_INVALID;
    return VSLAM_OK;
}

// enqueue the whole pipeline for n (<= cap_pairs) pairs; images and outputs are device pointers
// scratch of this call: pair slots [slot0, slot0 + n) of the context's volumes
static int sgbm_enqueue(vslam_ctx* ctx, const uint8_t* d_left, const uint8_t* d_right, int n, int w, int h, int pitch,
                        long long img_stride, const SgParams& p, int16_t* d_disp16, float* d_dispf, int slot0 = 0) {
    SgbmState* s0 = ctx->sgbm;
    cudaStream_t st = ctx->stream;
    const int W1 = w - SG_D;
    if (slot0 < 0 || slot0 + n > s0->cap_pairs) return VSLAM_E_CAPACITY;
    const size_t px0 = (size_t)slot0 * w * h, vol0 = (size_t)slot0 * h * W1 * SG_NDP;
    SgbmState view = *s0;  // the same buffers, shifted to this call's slots
    view.d_pre += 2 * px0;
    view.d_C += vol0;
    for (int i = 0; i < 3; ++i) view.d_L[i] += vol0;
    view.d_rec += px0;
    view.d_raw += px0;
    view.d_med += px0;
    view.d_label += px0;
    view.d_size += px0;
    view.d_runlen += px0;
    const SgbmState* s = &view;
    vslam_time_begin(ctx, VK_SGBM_PREFILTER);
    sgbm_prefilter_kernel<<<dim3(ceil_div(w, 256), h, 2 * n), 256, 0, st>>>(d_left, d_right, img_stride, pitch, n, w, h, p.ftzero,
                                                                            s->d_pre);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "sgbm_prefilter_kernel");
    // row bands: every band re-computes 2*SG_R halo rows; pick the band count whose waves x rows is smallest
    const int strips = ceil_div(W1, CK_CW);
    int bands = 1;
    long long best_cost = -1;
    for (int b = 1; b <= 32 && b <= h; ++b) {
        const long long waves = ((long long)strips * b * n + ctx->num_sms - 1) / ctx->num_sms;
        const long long cost = waves * (ceil_div(h, b) + 2 * SG_R);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            bands = b;
        }
    }
    const int band_rows = ceil_div(h, bands);
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(sgbm_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CK_SMEM));
    vslam_time_begin(ctx, VK_SGBM_COST);
    sgbm_cost_kernel<<<dim3(strips, ceil_div(h, band_rows), n), CK_THREADS, CK_SMEM, st>>>(s->d_pre, n, w, h, W1, band_rows, s->d_C);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "sgbm_cost_kernel");
    vslam_time_begin(ctx, VK_SGBM_VERTICAL);
    sgbm_vertical_kernel<<<dim3(ceil_div(W1, VT_PATHS), 3, n), (VT_WARPS + 1) * 32, 0, st>>>(s->d_C, s->d_L[0], s->d_L[1], s->d_L[2], W1, h, p.P1, p.P2);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "sgbm_vertical_kernel");
    if (s->stop_after == 1) return VSLAM_OK;
    const size_t smem_f = (size_t)SG_HW * RS_FWD_WARP;
    const size_t smem = (size_t)SG_HW * RS_BWD_STAGES;  // the right-image map (w x 4 B) aliases the ring afterwards
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(sgbm_row_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(sgbm_row_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vslam_time_begin(ctx, VK_SGBM_ROW_FWD);
    sgbm_row_forward_kernel<<<dim3(ceil_div(h, SG_HW), n), SG_HW * 32, smem_f, st>>>(s->d_C, s->d_L[0], s->d_L[1], s->d_L[2], W1, h,
                                                                                    p);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "sgbm_row_forward_kernel");
    vslam_time_begin(ctx, VK_SGBM_ROW_BWD);
    sgbm_row_backward_kernel<<<dim3(ceil_div(h, SG_HW), n), SG_HW * 32, smem, st>>>(s->d_C, s->d_L[0], w, W1, h, p, s->d_rec,
                                                                                    s->d_raw);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "sgbm_row_backward_kernel");
    const int hw = w * h;
    vslam_time_begin(ctx, VK_SGBM_POST);
    sgbm_median_kernel<<<dim3(ceil_div(w, 256), h, n), 256, 0, st>>>(s->d_raw, s->d_med, w, h);
    ctx->launches++;
    if (p.speckle_window > 0) {
        speckle_runs_kernel<<<ceil_div(n * h, 8), 256, 0, st>>>(s->d_med, s->d_label, s->d_size, s->d_runlen, w, h, n * h,
                                                               p.speckle_diff);
        if (h > 1)
            speckle_merge_kernel<<<dim3(ceil_div(w, 256), h - 1, n), 256, 0, st>>>(s->d_med, s->d_label, w, h, p.speckle_diff);
        speckle_count_kernel<<<dim3(ceil_div(hw, 256), n), 256, 0, st>>>(s->d_label, s->d_size, s->d_runlen, hw);
        ctx->launches += h > 1 ? 3 : 2;
    }
    speckle_apply_kernel<<<dim3(ceil_div(hw, 256), n), 256, 0, st>>>(s->d_med, s->d_label, s->d_size, hw, p.speckle_window, d_disp16,
                                                                    d_dispf);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "sgbm post kernels");
    return VSLAM_OK;
}

extern "C" int vslam_sgbm_compute_dev(vslam_ctx* ctx, const uint8_t* d_left, const uint8_t* d_right, int n_pairs, int width,
                                      int height, int row_pitch, long long image_stride, const vslam_sgbm_params* params,
                                      int16_t* d_disp16, float* d_disp_f32) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !d_left || !d_right || n_pairs < 0 || row_pitch < width || (!d_disp16 && !d_disp_f32)) return VSLAM_E_INVALID;
    SgParams p;
    int st = sgbm_check(params, width, height, &p);
    if (st != VSLAM_OK) return st;
    if (n_pairs == 0) return VSLAM_OK;
    VSLAM_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    // chunks alternate between the context stream and a second stream on disjoint scratch slots: the issue-bound
    // sweeps of one chunk overlap the bandwidth-bound sweeps of the other, and launch tails are filled
    static const int split_env = getenv("VSLAM_SGBM_SPLIT") ? atoi(getenv("VSLAM_SGBM_SPLIT")) : -1;
    const bool two = split_env != 0 && !ctx->serial && n_pairs >= 2;
    int chunk = two ? (n_pairs + 1) / 2 : n_pairs;
    if (chunk > SG_CHUNK_PAIRS / (two ? 2 : 1)) chunk = SG_CHUNK_PAIRS / (two ? 2 : 1);
    st = sgbm_reserve(ctx, two ? 2 * chunk : chunk, width, height);
    if (st != VSLAM_OK) return st;
    SgbmState* sg = ctx->sgbm;
    const size_t px = (size_t)width * height;
    cudaStream_t s = ctx->stream;
    if (two) {
        VSLAM_CUDA(ctx, cudaEventRecord(sg->ev_fork, s));
        VSLAM_CUDA(ctx, cudaStreamWaitEvent(sg->s_alt, sg->ev_fork, 0));
    }
    for (int b = 0, c = 0; b < n_pairs; b += chunk, ++c) {
        const int n = n_pairs - b < chunk ? n_pairs - b : chunk;
        const bool alt = two && (c & 1);
        ctx->stream = alt ? sg->s_alt : s;
        st = sgbm_enqueue(ctx, d_left + (long long)b * image_stride, d_right + (long long)b * image_stride, n, width, height,
                          row_pitch, image_stride, p, d_disp16 ? d_disp16 + b * px : nullptr,
                          d_disp_f32 ? d_disp_f32 + b * px : nullptr, alt ? chunk : 0);
        ctx->stream = s;
        if (st != VSLAM_OK) break;
    }
    if (two) {
        cudaEventRecord(sg->ev_join, sg->s_alt);
        cudaStreamWaitEvent(s, sg->ev_join, 0);
    }
    return st;
}

extern "C" int vslam_sgbm_compute(vslam_ctx* ctx, const uint8_t* left, const uint8_t* right, int n_pairs, int width, int height,
                                  int row_pitch, long long image_stride, const vslam_sgbm_params* params, int16_t* disp16,
                                  float* disp_f32) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !left || !right || n_pairs < 0 || row_pitch < width || (!disp16 && !disp_f32)) return VSLAM_E_INVALID;
    SgParams p;
    int st = sgbm_check(params, width, height, &p);
    if (st != VSLAM_OK) return st;
    if (n_pairs == 0) return VSLAM_OK;
    VSLAM_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const int chunk = n_pairs < SG_CHUNK_PAIRS ? n_pairs : SG_CHUNK_PAIRS;
    st = sgbm_reserve(ctx, chunk, width, height);
    if (st != VSLAM_OK) return st;
    SgbmState* s = ctx->sgbm;
    const size_t px = (size_t)width * height;
    const long long dstride = (long long)s->pitch * height;
    for (int b = 0; b < n_pairs; b += chunk) {
        const int n = n_pairs - b < chunk ? n_pairs - b : chunk;
        for (int side = 0; side < 2; ++side)
            for (int i = 0; i < n; ++i)
                VSLAM_CUDA(ctx, cudaMemcpy2DAsync(s->d_img + (size_t)(side * n + i) * dstride, s->pitch,
                                                  (side ? right : left) + (long long)(b + i) * image_stride, row_pitch, width,
                                                  height, cudaMemcpyHostToDevice, ctx->stream));
        st = sgbm_enqueue(ctx, s->d_img, s->d_img + (size_t)n * dstride, n, width, height, s->pitch, dstride, p,
                          disp16 ? s->d_out : nullptr, disp_f32 ? s->d_outf : nullptr);
        if (st != VSLAM_OK) return st;
        if (s->stop_after == 0) {
            if (disp16)
                VSLAM_CUDA(ctx, cudaMemcpyAsync(disp16 + b * px, s->d_out, n * px * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
            if (disp_f32)
                VSLAM_CUDA(ctx, cudaMemcpyAsync(disp_f32 + b * px, s->d_outf, n * px * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        }
        VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return VSLAM_OK;
}

// test tap: stage 0 = C, 1..3 = path volumes (1 holds S4 after the horizontal sweep), 4 = raw disparity, 5 = median
extern "C" int vslam_sgbm_debug_read(vslam_ctx* ctx, int pair, int stage, void* host_out, size_t bytes) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !ctx->sgbm || !host_out || pair < 0 || pair >= ctx->sgbm->cap_pairs) return VSLAM_E_INVALID;
    SgbmState* s = ctx->sgbm;
    const size_t vol = (size_t)s->cap_h * (s->cap_w - SG_D) * SG_NDP * sizeof(uint32_t);
    const size_t img = (size_t)s->cap_w * s->cap_h * sizeof(int16_t);
    const void* src;
    size_t n;
    if (stage == 0) { src = (const char*)s->d_C + pair * vol; n = vol; }
    else if (stage >= 1 && stage <= 3) { src = (const char*)s->d_L[stage - 1] + pair * vol; n = vol; }
    else if (stage == 4) { src = (const char*)s->d_raw + pair * img; n = img; }
    else if (stage == 5) { src = (const char*)s->d_med + pair * img; n = img; }
    else return VSLAM_E_INVALID;
    if (bytes < n) return VSLAM_E_CAPACITY;
    VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VSLAM_CUDA(ctx, cudaMemcpy(host_out, src, n, cudaMemcpyDeviceToHost));
    return VSLAM_OK;
}

extern "C" int vslam_sgbm_debug_stop_after(vslam_ctx* ctx, int stage) {
    if (!ctx || !ctx->sgbm) return VSLAM_E_INVALID;
    ctx->sgbm->stop_after = stage;
    return VSLAM_OK;
}
