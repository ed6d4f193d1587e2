// K1-K9: bit-exact ORB (oriented FAST + rotated BRIEF) on the GPU, batched over images.
//
// Replaces the arithmetic behind VO::feature_detection
// (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:70-94): detector_->detect (:80, ORB with
// nfeatures = 3000, :22/:31), adaptive_non_maximal_suppresion (:82, :96-157) and descriptor_->compute (:85).
// The arithmetic itself lives in OpenCV (cv::ORB, un-vendored); the bit-exact specification followed here is
// SURVEY.md §A.1 as re-validated by oracle/orb_restate.py against cv2 4.13.0.
//
// Pipeline (every kernel covers ALL images of the batch; nothing returns to the host in between):
//   resize_level_kernel x7   K1  INTER_LINEAR_EXACT chain, 8.8 fixed point, coefficient tables from the host
//   fast_kernel              K2  FAST-9/16 (t=20) score + 3x3 NMS + 31-px border filter, tile = 60x30 in smem,
//                                candidate list + per-(image,level) 256-bin score histogram
//   harris_select_kernel     K3-K5 one CTA per (level,image): histogram cut (retainBest 2N, ties kept), Harris
//                                response (7x7, no FMA), bitonic sort by (response desc, y, x), retainBest N
//   blur_kernel              K8  7x7 sigma-2 float32 separable blur with cv2's exact fused/unfused op order
//   anms_kernel (optional)   K7  the reference's own ANMS
//   describe_kernel          K6+K9 one warp per keypoint: intensity-centroid angle, fastAtan2, 256 rBRIEF tests
// This file is compiled with -fmad=false; fused operations are explicit fmaf() where cv2's AVX2 path fuses.
#include "common.cuh"

#include <cuda_pipeline.h>
#include <math.h>
#include <stdlib.h>

#include <vector>

#define ORB_NL VSLAM_NLEVELS
#define ORB_EDGE 31
#define FAST_T 20
#define FAST_THREADS 256
#define SORT_CAP 8192
#define HS_THREADS 1024
#define DESC_WARPS 4
#define DESC_KPW 8  // keypoints per warp (power of two <= 32)
#ifndef DESC_MIN_CTAS
#define DESC_MIN_CTAS 8
#endif
#define DESC_PH 19   // half height of the staged rBRIEF window: |rotated pattern coordinate| <= 13 sqrt(2) + rounding < 19
#define DESC_PR (2 * DESC_PH + 1)  // rows
#define DESC_PW 64   // bytes per row: the window starts at the 16-byte boundary left of cx - 19, 39 + 15 <= 64

struct OrbLevel {
    int w, h, pitch;
    int off;  // byte offset of the plane inside one image's slab (pyramid and blurred pyramid share the layout)
    int tiles_x, tiles_y, tile_begin;  // blur tiling: BL_W x BL_H over the whole level
    int ft_x, ft_y, ft_begin;          // FAST tiling: FT_W x FT_H over the frame the 31-px border filter keeps
    int cand_off, cand_cap;  // candidate-list region (entries) inside one image's candidate slab
    int xtab_off, ytab_off;  // resize coefficient tables (entries)
    float scale, inv_scale;
};

struct OrbGeom {
    OrbLevel lv[ORB_NL];
    int total_tiles, total_ft;
    int img_slab;   // bytes per image in the pyramid / blurred slabs
    int cand_slab;  // candidate entries per image
    int w, h;
};

struct OrbQuota {
    int n[ORB_NL];
};

struct ImgCounters {
    uint32_t hist[ORB_NL][256];
    uint32_t cand_cnt[ORB_NL];
    uint32_t sel_cnt[ORB_NL];
    uint32_t n_keep;  // ANMS output count
    uint32_t flags;   // bit0: candidate list overflow, bit1: sort overflow, bit2: keypoint capacity overflow
    uint32_t pad[2];
};

struct OrbState {
    OrbGeom geom;
    bool geom_valid;
    uint8_t* d_pyr;
    uint8_t* d_blur;
    uint32_t* d_tab;      // resize tables
    uint32_t* d_ftab;     // FAST tile table: level | tile column << 4 | tile row << 16, one entry per tile of an image
    int ftab_cap;
    uint2* d_cand;        // {x | y << 16, score}
    uint2* d_sel;         // [img][level][SORT_CAP] {x | y << 16, response bits}
    ImgCounters* d_cnt;
    uint32_t* d_sticky;   // OR of every overflow flag raised since the last check (survives chunked batches)
    uint32_t* d_keep;     // ANMS keep list [img][kp_cap]
    double* d_rad;        // ANMS radii [img][kp_cap]
    float* d_pattern;     // 256 x 4 (x0, y0, x1, y1) as float
    // staging for the host-buffer entry point
    uint8_t* d_in;
    int in_pitch;
    vslam_keypoint* d_kp;
    uint8_t* d_desc;
    int32_t* d_n;
    vslam_keypoint* h_kp;
    uint8_t* h_desc;
    int32_t* h_n;
    ImgCounters* h_cnt;
    int max_images, kp_cap;
    size_t slab_cap, cand_cap_total;
    int tab_cap;
};

static const int8_t h_orb_pattern[256 * 4] = {
#include "orb_pattern.inc"
};

// float(pow(double(1.2f), l)), l = 0..7 (cv::ORB getScale; verified against oracle/orb_restate.level_scales)
static const float h_scales[ORB_NL] = {0x1.0p+0f,        0x1.333334p+0f, 0x1.70a3d8p+0f, 0x1.ba5e38p+0f,
                                       0x1.096bbcp+1f, 0x1.3e814ap+1f, 0x1.7e34c0p+1f, 0x1.caa5b4p+1f};

__device__ __forceinline__ const uint8_t* level_ptr(const ImgSrc& src, const uint8_t* slab, const OrbGeom& g, int img,
                                                    int l, int& pitch) {
    if (l == 0) {
        pitch = src.pitch;
        const bool second = img >= src.per_base;
        return (second ? src.base[1] : src.base[0]) + (long long)(second ? img - src.per_base : img) * src.img_stride;
    }
    pitch = g.lv[l].pitch;
    return slab + (size_t)img * g.img_slab + g.lv[l].off;
}

// ------------------------------------------------------------------------------------------------------------
// K1  resize INTER_LINEAR_EXACT: out = (c0y*(c0x*s00 + c1x*s01) + c1y*(c0x*s10 + c1x*s11) + 32768) >> 16
// ------------------------------------------------------------------------------------------------------------
// One thread produces 8 horizontally adjacent output pixels (two quads) and stores them as one 64-bit word.  A quad's
// source bytes span at most 6 columns (scale 1.2), fetched per source row as up to three aligned words and shifted
// into place; the horizontal pass of each pixel is one byte permute + one 16x8-bit dot product (IDP.2A).  Everything
// that depends only on the output column -- first source column of the quad, last needed byte, the four byte selectors
// and the four coefficient pairs -- comes precomputed from the host as 8 words per quad (two 128-bit loads).
__device__ __forceinline__ void load_window(const uint8_t* p, int last_byte, uint32_t& v0, uint32_t& v1) {
    const uint32_t a = (uint32_t)(uintptr_t)p & 3u;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
    const int need = (int)a + last_byte;  // index of the last needed byte inside the aligned window
    const uint32_t w0 = __ldg(q);
    const uint32_t w1 = need >= 4 ? __ldg(q + 1) : 0u;
    const uint32_t w2 = need >= 8 ? __ldg(q + 2) : 0u;
    v0 = __funnelshift_r(w0, w1, 8 * a);
    v1 = __funnelshift_r(w1, w2, 8 * a);
}

__global__ void __launch_bounds__(256)
resize_level_kernel(ImgSrc src, uint8_t* __restrict__ pyr, const uint32_t* __restrict__ tab, const __grid_constant__ OrbGeom g, int l) {
    const int img = blockIdx.z;
    const OrbLevel& L = g.lv[l];
    const int x = (blockIdx.x * 32 + threadIdx.x) * 8;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= L.w || y >= L.h) return;
    int sp;
    const uint8_t* s = level_ptr(src, pyr, g, img, l - 1, sp);
    const int sh = g.lv[l - 1].h;
    const uint32_t ty = __ldg(&tab[L.ytab_off + y]);
    const int oy = ty >> 16, c1y = ty & 0xFFFF, c0y = 256 - c1y;
    const uint8_t* r0 = s + (size_t)oy * sp;
    const uint8_t* r1 = s + (size_t)min(oy + 1, sh - 1) * sp;
    const uint4* qt = reinterpret_cast<const uint4*>(&tab[L.xtab_off]) + (x >> 2) * 2;  // 8 words per quad
    uint32_t out[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (x + 4 * h >= L.w) break;  // the second quad lies in the row padding
        const uint4 qa = __ldg(qt + 2 * h), coef = __ldg(qt + 2 * h + 1);
        const int ox0 = qa.x & 0xFFFF, last = qa.x >> 16;
        uint32_t a0, a1, b0, b1;
        load_window(r0 + ox0, last, a0, a1);
        load_window(r1 + ox0, last, b0, b1);
        const uint32_t cf[4] = {coef.x, coef.y, coef.z, coef.w};
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t sel = (qa.y >> (8 * k)) & 0xFFu;
            const uint32_t h0 = __dp2a_lo(cf[k], __byte_perm(a0, a1, sel), 0u);
            const uint32_t h1 = __dp2a_lo(cf[k], __byte_perm(b0, b1, sel), 0u);
            const uint32_t v = ((uint32_t)c0y * h0 + (uint32_t)c1y * h1 + 32768u) >> 16;
            packed |= v << (8 * k);
        }
        out[h] = packed;
    }
    // x is a multiple of 8 and the plane pitch a multiple of 16: the 8 bytes stay inside the (padded) row
    *reinterpret_cast<uint2*>(&pyr[(size_t)img * g.img_slab + L.off + (size_t)y * L.pitch + x]) = make_uint2(out[0], out[1]);
}

// ------------------------------------------------------------------------------------------------------------
// K2  FAST-9/16 + NMS + border filter
// ------------------------------------------------------------------------------------------------------------
// The tile's pixels are widened to 16 bits in shared memory so that one 32-bit word holds TWO horizontally adjacent
// pixels; every step of the segment test / corner score then runs on a pixel pair with the native packed-16-bit
// instructions of sm_100a (VIADD.16x2, VIMNMX.S16x2, VIMNMX3.S16x2 -- half-rate ALU instructions: 80 of them per pair
// are the corner-score network, the floor of this kernel).
//
// Only pixels that survive ORB's 31-px border filter are ever used, and their 3x3 NMS neighbourhood and 16-px rings
// lie >= 27 px inside the image, so the tile grid covers just the kept frame [30, W-31) x [31, H-31): no border
// cases anywhere, and 20 % (level 0) to 65 % (level 7) fewer pixels than the whole level.
//   output tile : FT_W x FT_H pixels at (tx0, ty0) = (30 + 60 i, 31 + 30 j)   (column 30 is dropped by the filter;
//                 it keeps pixel pairs on even columns, i.e. on the byte pairs of the aligned global loads)
//   score tile  : rows ty0-1 .. ty0+30, 32 pixel pairs from column tx0-2: one warp per row, one lane per pair
//   image tile  : rows ty0-4 .. ty0+33, 96 pixels from gstart = (tx0-6) & ~15 (six aligned 128-bit loads per row)
// Phases: (1) compass pre-test, every warp compacts the surviving pairs of its rows into its own list (ballot +
// popc, no atomics, no block barrier); (2) full segment test + score on the list; (3) 3x3 NMS with a sliding
// 3-row window in registers, border filter, emit into the warp's own output segment (positions from ballots).
// The tile -> (level, column, row) mapping comes from a table built with the geometry (one broadcast load).
#define FT_W 60
#define FT_H 30
#define FT_X0 30
#define FT_Y0 ORB_EDGE
#define FI_ROWS (FT_H + 8)
#define FI_WORDS 48                 // words per image-tile row (96 pixels; 64 measured: same time, the conflicts do not bind)
#define FS_ROWS (FT_H + 2)
#define FS_WORDS 32                 // words (= pixel pairs) per score-tile row
#define FAST_WARPS (FAST_THREADS / 32)
#define FAST_RPW (FS_ROWS / FAST_WARPS)   // score rows per warp
#define FAST_WOUT 64                // candidates one warp can emit: its 4 output rows hold at most 2 x 15 strict 3x3 maxima
#define FAST_TT ((uint32_t)FAST_T | ((uint32_t)FAST_T << 16))
static_assert(FS_ROWS % FAST_WARPS == 0 && FS_WORDS == 32, "one warp per score row, one lane per pixel pair");

__device__ __forceinline__ uint32_t neg16x2(uint32_t a) { return __vadd2(~a, 0x00010001u); }
// two adjacent pixels starting at the odd pixel of word `lo` (upper half of lo, lower half of hi)
__device__ __forceinline__ uint32_t mid16x2(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5432); }
__global__ void __launch_bounds__(FAST_THREADS)
fast_kernel(ImgSrc src, const uint8_t* __restrict__ pyr, const __grid_constant__ OrbGeom g, const uint32_t* __restrict__ tile_tab,
            uint2* __restrict__ cand, ImgCounters* __restrict__ cnt, uint32_t* __restrict__ sticky) {
    __shared__ __align__(16) uint32_t s_img[FI_ROWS * FI_WORDS];
    __shared__ uint32_t s_sc[FS_ROWS * FS_WORDS];
    __shared__ uint16_t s_list[FAST_WARPS][2][FAST_RPW * 32];
    __shared__ uint2 s_out[FAST_WARPS][FAST_WOUT];
    __shared__ uint32_t s_hist[256];
    __shared__ int s_wcnt[FAST_WARPS];
    __shared__ int s_base;

    const int img = blockIdx.y;
    // tile -> (level, tile column, tile row) from a table the host built with the geometry: one broadcast load
    // instead of a level search and an integer division in every thread
    const uint32_t te = __ldg(&tile_tab[blockIdx.x]);
    const int l = te & 15;
    const OrbLevel& L = g.lv[l];
    const int tx0 = FT_X0 + (int)((te >> 4) & 0xFFFu) * FT_W, ty0 = FT_Y0 + (int)(te >> 16) * FT_H;
    int pitch;
    const uint8_t* im = level_ptr(src, pyr, g, img, l, pitch);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int W = L.w, H = L.h;
    const int gstart = (tx0 - 6) & ~15;
    const int shift = (tx0 - 6 - gstart) >> 1;  // word offset of pixel tx0-6 inside an image-tile row (0, 2, 4 or 6)
    // rows / pairs of the score tile that some kept pixel's NMS can read
    const int n_srow = min(FS_ROWS, H - ORB_EDGE - ty0 + 2);       // score rows 0 .. n_srow-1 (gy <= H-31)
    const int n_orow = min(FT_H, H - ORB_EDGE - ty0);              // output rows
    const bool pair_live = tx0 - 2 + 2 * lane <= W - ORB_EDGE;     // this lane's score pair (gx <= W-31)

    // ---- load: one 128-bit load per 16 pixels when the level's rows are 16-byte aligned (every pyramid level, and
    // level 0 whenever the caller's pitch and base allow it); all loads of a thread are issued before any is used.
    // Pixels beyond the image only feed scores of pixels the border filter drops; they read as zero.
    const bool aligned16 = ((pitch & 15) == 0) && (((uintptr_t)im & 15) == 0);
    if (aligned16) {
        constexpr int N = FI_ROWS * 6, ROUNDS = (N + FAST_THREADS - 1) / FAST_THREADS;
        uint4 px[ROUNDS];
#pragma unroll
        for (int k = 0; k < ROUNDS; ++k) {
            const int i = tid + k * FAST_THREADS;
            const int r = i / 6, ch = i - r * 6;
            const int gy = ty0 - 4 + r, gx = gstart + 16 * ch;
            px[k] = make_uint4(0, 0, 0, 0);
            if (i < N && gy < H && gx < pitch) px[k] = __ldg(reinterpret_cast<const uint4*>(im + (size_t)gy * pitch + gx));
        }
#pragma unroll
        for (int k = 0; k < ROUNDS; ++k) {
            const int i = tid + k * FAST_THREADS;
            if (i >= N) continue;
            uint4* dst = reinterpret_cast<uint4*>(&s_img[(i / 6) * FI_WORDS + 8 * (i % 6)]);
            dst[0] = make_uint4(__byte_perm(px[k].x, 0, 0x4140), __byte_perm(px[k].x, 0, 0x4342),
                                __byte_perm(px[k].y, 0, 0x4140), __byte_perm(px[k].y, 0, 0x4342));
            dst[1] = make_uint4(__byte_perm(px[k].z, 0, 0x4140), __byte_perm(px[k].z, 0, 0x4342),
                                __byte_perm(px[k].w, 0, 0x4140), __byte_perm(px[k].w, 0, 0x4342));
        }
    } else {  // any pitch / base: byte loads
        for (int i = tid; i < FI_ROWS * FI_WORDS; i += FAST_THREADS) {
            const int r = i / FI_WORDS, m = i - r * FI_WORDS;
            const int gy = ty0 - 4 + r, gx = gstart + 2 * m;
            uint32_t v = 0;
            if (gy < H) {
                const uint8_t* p = im + (size_t)gy * pitch + gx;
                if (gx < W) v = p[0];
                if (gx + 1 < W) v |= (uint32_t)p[1] << 16;
            }
            s_img[i] = v;
        }
    }
    s_hist[tid] = 0;  // FAST_THREADS == 256
    __syncthreads();

    // ---- phase 1: compass pre-test on pixel pairs (any 9-arc contains two adjacent compass pixels) ---------------
    // The brighter and the darker half are tested separately: a pixel whose brighter (darker) pre-test fails has a
    // brighter- (darker-) arc score <= FAST_T, which the final max with FAST_T hides -- so each half of the score
    // network only runs where its own pre-test passed (a pixel can never be a corner both ways).
    const uint32_t* ctr = &s_img[3 * FI_WORDS + shift + 2 + lane];  // centre word of score pair (row 0, lane)
    uint16_t* listb = s_list[wid][0];
    uint16_t* listd = s_list[wid][1];
    int nb = 0, nd = 0;
#pragma unroll
    for (int k = 0; k < FAST_RPW; ++k) {
        const int r = wid + k * FAST_WARPS;
        const uint32_t* c = ctr + r * FI_WORDS;
        // "two adjacent compass pixels both brighter than c + t" = (b0 | b8) & (b4 | b12) on sign bits: (c + t) - ring is
        // negative exactly where the ring pixel is brighter (values are 9-bit, no 16-bit overflow), so the whole test is
        // eight packed subtractions and four logic ops on the full-rate pipe instead of fourteen half-rate min / max
        const uint32_t cv = c[0];
        const uint32_t r0 = c[3 * FI_WORDS], r8 = c[-3 * FI_WORDS];
        const uint32_t r4 = mid16x2(c[1], c[2]), r12 = mid16x2(c[-2], c[-1]);
        const uint32_t hi = __vadd2(cv, FAST_TT), lo = __vsub2(cv, FAST_TT);
        const uint32_t sb = (__vsub2(hi, r0) | __vsub2(hi, r8)) & (__vsub2(hi, r4) | __vsub2(hi, r12));  // sign: brighter pair
        const uint32_t sd = (__vsub2(r0, lo) | __vsub2(r8, lo)) & (__vsub2(r4, lo) | __vsub2(r12, lo));  // sign: darker pair
        const bool live = pair_live & (r < n_srow);
        const bool pb = live & ((sb & 0x80008000u) != 0), pd = live & ((sd & 0x80008000u) != 0);
        s_sc[r * FS_WORDS + lane] = FAST_TT;
        const uint32_t balb = __ballot_sync(0xFFFFFFFFu, pb), bald = __ballot_sync(0xFFFFFFFFu, pd);
        const uint32_t lt = (1u << lane) - 1;
        if (pb) listb[nb + __popc(balb & lt)] = (uint16_t)(r * FS_WORDS + lane);
        if (pd) listd[nd + __popc(bald & lt)] = (uint16_t)(r * FS_WORDS + lane);
        nb += __popc(balb);
        nd += __popc(bald);
    }
    __syncwarp();

    // ---- phase 2: full segment test + corner score on the surviving pairs ----------------------------------------
    // score = max over the 16 arcs of max(min(d), -max(d)), d = ring - centre; a pixel is a corner iff that exceeds
    // FAST_T.  The network runs on the RAW ring values: min / max commute with subtracting the centre, so
    //   max_k min9_k(ring - c) = max_k min9_k(ring) - c   and   min_k max9_k(ring - c) = min_k max9_k(ring) - c,
    // and a 9-window is three 3-windows: m3[k] = op3(d[k], d[k+1], d[k+2]), m9[k] = op3(m3[k], m3[k+3], m3[k+6])
    // -- 32 three-input instructions per direction instead of 16 subtractions + 48.
    // h = max(score, FAST_T): every pixel scored here has its whole ring inside the image.
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint16_t* list = half ? listd : listb;
        const int n1 = half ? nd : nb;
        for (int e = lane; e < n1; e += 32) {
            const int p = list[e];
            const uint32_t* c = &s_img[3 * FI_WORDS + shift + 2] + (p >> 5) * FI_WORDS + (p & 31);
            const uint32_t cv = c[0];
            uint32_t d[16];
            {
                const uint32_t* q = c + 3 * FI_WORDS;
                const uint32_t a = q[-1], b = q[0], cc = q[1];
                d[15] = mid16x2(a, b); d[0] = b; d[1] = mid16x2(b, cc);
            }
            {
                const uint32_t* q = c - 3 * FI_WORDS;
                const uint32_t a = q[-1], b = q[0], cc = q[1];
                d[9] = mid16x2(a, b); d[8] = b; d[7] = mid16x2(b, cc);
            }
            d[14] = c[2 * FI_WORDS - 1]; d[2] = c[2 * FI_WORDS + 1];
            d[10] = c[-2 * FI_WORDS - 1]; d[6] = c[-2 * FI_WORDS + 1];
            {
                const uint32_t* q = c + FI_WORDS;
                d[13] = mid16x2(q[-2], q[-1]); d[3] = mid16x2(q[1], q[2]);
            }
            {
                const uint32_t* q = c - FI_WORDS;
                d[11] = mid16x2(q[-2], q[-1]); d[5] = mid16x2(q[1], q[2]);
            }
            d[12] = mid16x2(c[-2], c[-1]); d[4] = mid16x2(c[1], c[2]);
            uint32_t m3[16], m9[16];
            if (half == 0) {  // brightest arc: max over the arcs of the arc minimum
#pragma unroll
                for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
#pragma unroll
                for (int k = 0; k < 16; ++k) m9[k] = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                uint32_t best = __vimax3_s16x2(m9[0], m9[1], m9[2]);
#pragma unroll
                for (int k = 3; k < 15; k += 2) best = __vimax3_s16x2(best, m9[k], m9[k + 1]);
                best = __vsub2(__vmaxs2(best, m9[15]), cv);
                s_sc[p] = __vmaxs2(best, FAST_TT);  // phase 1 left FAST_T here
            } else {          // darkest arc: min over the arcs of the arc maximum
#pragma unroll
                for (int k = 0; k < 16; ++k) m3[k] = __vimax3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
#pragma unroll
                for (int k = 0; k < 16; ++k) m9[k] = __vimax3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                uint32_t best = __vimin3_s16x2(m9[0], m9[1], m9[2]);
#pragma unroll
                for (int k = 3; k < 15; k += 2) best = __vimin3_s16x2(best, m9[k], m9[k + 1]);
                best = __vsub2(cv, __vmins2(best, m9[15]));
                s_sc[p] = __vmaxs2(best, s_sc[p]);  // on top of the brighter half's result (>= FAST_T)
            }
        }
        __syncwarp();
    }
    __syncthreads();

    int n_wout = 0;  // candidates this warp has emitted
    // ---- phase 3: 3x3 NMS (strictly greater) on pixel pairs, border filter, emit ------------------------------------
    // Warp w owns output rows 4w .. 4w+3 and slides down them: per score row `side` = max of the two horizontal
    // neighbours, `full` = max(side, centre); the 8-neighbour maximum of row r is max3(full[r-1], side[r], full[r+1]).
    {
        constexpr int OPW = (FT_H + FAST_WARPS - 1) / FAST_WARPS;
        const int ro0 = wid * OPW;
        // lane <-> score word `lane`; output pairs are words 1..30 (pixels tx0 + 2 (lane-1))
        const int gx = tx0 - 2 + 2 * lane;
        const bool col0 = lane >= 1 && lane <= FT_W / 2 && gx >= ORB_EDGE && gx < W - ORB_EDGE;
        const bool col1 = lane >= 1 && lane <= FT_W / 2 && gx + 1 < W - ORB_EDGE;
        const int ll = max(lane - 1, 0), lr = min(lane + 1, 31);
        uint32_t full_prev = 0, c_cur = 0, side_cur = 0, full_cur = 0;
        uint2* wout = s_out[wid];  // this warp's own output segment: positions from ballots, no atomics
        const uint32_t lt = (1u << lane) - 1;
#pragma unroll
        for (int k = 0; k < OPW + 2; ++k) {
            const int sr = ro0 + k;  // score row (clamped: the rows past the tile only feed dead outputs)
            const uint32_t* row = &s_sc[min(sr, FS_ROWS - 1) * FS_WORDS];
            const uint32_t cw = row[lane];
            const uint32_t side = __vmaxs2(mid16x2(row[ll], cw), mid16x2(cw, row[lr]));
            const uint32_t full = __vmaxs2(side, cw);
            if (k >= 2) {
                const int ro = sr - 2;  // output row whose centre is score row sr-1 (= c_cur)
                const uint32_t m = __vimax3_s16x2(full_prev, side_cur, full);
                const uint32_t x = __vmaxs2(c_cur, m) ^ m;  // lane != 0  <=>  centre strictly greater than its 8 neighbours
                const bool live = ro < n_orow;
                const bool k0 = ((x & 0xFFFFu) != 0) & col0 & live, k1 = ((x >> 16) != 0) & col1 & live;
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, k0), b1 = __ballot_sync(0xFFFFFFFFu, k1);
                if (b0 | b1) {
                    int pos = n_wout + __popc(b0 & lt) + __popc(b1 & lt);
                    const uint32_t gy = (uint32_t)(ty0 + ro) << 16;
                    if (k0) {
                        const uint32_t sc = (c_cur & 0xFFFFu) - 1;
                        if (pos < FAST_WOUT) wout[pos] = make_uint2((uint32_t)gx | gy, sc);
                        atomicAdd(&s_hist[sc], 1u);
                        ++pos;
                    }
                    if (k1) {
                        const uint32_t sc = (c_cur >> 16) - 1;
                        if (pos < FAST_WOUT) wout[pos] = make_uint2((uint32_t)(gx + 1) | gy, sc);
                        atomicAdd(&s_hist[sc], 1u);
                    }
                    n_wout += __popc(b0) + __popc(b1);
                }
            }
            full_prev = full_cur;
            c_cur = cw; side_cur = side; full_cur = full;
        }
    }
    if (lane == 0) s_wcnt[wid] = min(n_wout, FAST_WOUT);
    __syncthreads();
    int nout = 0, my_off = 0;
#pragma unroll
    for (int w = 0; w < FAST_WARPS; ++w) {
        const int c = s_wcnt[w];
        if (w < wid) my_off += c;
        nout += c;
    }
    if (nout == 0) return;
    ImgCounters* C = &cnt[img];
    if (tid == 0) s_base = (int)atomicAdd(&C->cand_cnt[l], (uint32_t)nout);
    __syncthreads();
    const int base = s_base + my_off;
    uint2* dst = cand + (size_t)img * g.cand_slab + L.cand_off;
    for (int i = lane; i < s_wcnt[wid]; i += 32) {
        if (base + i < L.cand_cap) dst[base + i] = s_out[wid][i];
    }
    if (tid == 0 && s_base + nout > L.cand_cap) {
        atomicOr(&C->flags, 1u);
        atomicOr(sticky, 1u);
    }
    if (s_hist[tid]) atomicAdd(&C->hist[l][tid], s_hist[tid]);
}

// ------------------------------------------------------------------------------------------------------------
// K3-K5  per (level, image): histogram cut, Harris, sort, retainBest
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_desc_key(float f) {
    uint32_t b = __float_as_uint(f);
    b ^= (b >> 31) ? 0xFFFFFFFFu : 0x80000000u;  // monotone increasing
    return ~b;                                    // ascending key == descending response
}

__global__ void __launch_bounds__(HS_THREADS)
harris_select_kernel(ImgSrc src, const uint8_t* __restrict__ pyr, const __grid_constant__ OrbGeom g, const __grid_constant__ OrbQuota quota,
                     const uint2* __restrict__ cand, ImgCounters* __restrict__ cnt, uint2* __restrict__ sel,
                     uint32_t* __restrict__ sticky) {
    extern __shared__ unsigned long long s_key[];  // SORT_CAP
    __shared__ uint32_t s_h[256];
    __shared__ __align__(16) uint8_t s_patch[HS_THREADS / 32][112];  // 9 rows x 3 words
    __shared__ int s_cut, s_n, s_kept;

    const int l = blockIdx.x, img = blockIdx.y;
    const OrbLevel& L = g.lv[l];
    ImgCounters* C = &cnt[img];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = quota.n[l];
    const int nc = min((int)C->cand_cnt[l], L.cand_cap);
    if (tid < 256) s_h[tid] = C->hist[l][tid];
    if (tid == 0) {
        s_cut = 0;
        s_n = 0;
        s_kept = 0;
    }
    __syncthreads();
    if (N <= 0 || nc == 0) {
        if (tid == 0) C->sel_cnt[l] = 0;
        return;
    }
    // retainBest(2N) on integer FAST scores: cut = value of the 2N-th largest (ties kept)
    if (tid < 256) {
        uint32_t ge = 0;
        for (int v = tid; v < 256; ++v) ge += s_h[v];
        if (ge >= (uint32_t)(2 * N)) atomicMax(&s_cut, tid);
    }
    __syncthreads();
    const uint32_t cut = (uint32_t)s_cut;
    const uint2* cl = cand + (size_t)img * g.cand_slab + L.cand_off;
    for (int i = tid; i < nc; i += HS_THREADS) {
        const uint2 c = cl[i];
        if (c.y >= cut) {
            const int pos = atomicAdd(&s_n, 1);
            if (pos < SORT_CAP) s_key[pos] = c.x;
        }
    }
    __syncthreads();
    int n = s_n;
    if (n > SORT_CAP) {
        if (tid == 0) {
            atomicOr(&C->flags, 2u);
            atomicOr(sticky, 2u);
        }
        n = SORT_CAP;
    }
    int n2 = 1;
    while (n2 < n) n2 <<= 1;

    // Harris response, one warp per candidate (orb.cpp HarrisResponses: blockSize 7, k = 0.04)
    int pitch;
    const uint8_t* im = level_ptr(src, pyr, g, img, l, pitch);
    // The 9 x 9 patch is fetched as 9 rows x 3 aligned words (27 lanes, one word each) and the two 3 x 3 derivative
    // kernels of a position become byte dot products (DP4A) on its three 3-byte row windows:
    //   ix = [-1 0 1; -2 0 2; -1 0 1] . patch,  iy = [-1 -2 -1; 0 0 0; 1 2 1] . patch     (integer sums, order-free)
    for (int e = warp; e < n; e += HS_THREADS / 32) {
        const uint32_t xy = (uint32_t)s_key[e];
        const int x = xy & 0xFFFF, y = xy >> 16;
        uint32_t* P = reinterpret_cast<uint32_t*>(s_patch[warp]);  // [9][3] words; row r starts at the word boundary left of x-4
        const uint8_t* p0 = im + (size_t)(y - 4) * pitch + (x - 4);
        if (lane < 27) {
            const int r = lane / 3, wj = lane - r * 3;
            const uint8_t* pr = p0 + (size_t)r * pitch;
            const uint32_t al = (uint32_t)(uintptr_t)pr & 3u;
            P[lane] = __ldg(reinterpret_cast<const uint32_t*>(pr - al) + wj);
        }
        __syncwarp();
        int a = 0, b = 0, c = 0;
#pragma unroll
        for (int rnd = 0; rnd < 2; ++rnd) {
            const int i = lane + 32 * rnd;
            if (i < 49) {
                const int r = i / 7, q = i - r * 7;
                uint32_t w3[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {  // bytes q .. q+2 of patch row r + k
                    const uint32_t al = (uint32_t)(uintptr_t)(p0 + (size_t)(r + k) * pitch) & 3u;
                    const uint32_t o = al + (uint32_t)q;
                    const uint32_t* pw = P + 3 * (r + k) + (o >> 2);
                    w3[k] = __funnelshift_r(pw[0], (o >> 2) < 2 ? pw[1] : 0u, 8 * (o & 3u));
                }
                int ix = 0, iy = 0;
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(ix) : "r"(w3[0]), "r"(0x0001'00FFu));  // (-1, 0, 1, 0)
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(ix) : "r"(w3[1]), "r"(0x0002'00FEu));  // (-2, 0, 2, 0)
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(ix) : "r"(w3[2]), "r"(0x0001'00FFu));
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(iy) : "r"(w3[0]), "r"(0x00FF'FEFFu));  // (-1, -2, -1, 0)
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(iy) : "r"(w3[2]), "r"(0x0001'0201u));  // ( 1,  2,  1, 0)
                a += ix * ix;
                b += iy * iy;
                c += ix * iy;
            }
        }
        a = __reduce_add_sync(0xFFFFFFFFu, a);
        b = __reduce_add_sync(0xFFFFFFFFu, b);
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane == 0) {
            const float af = (float)a, bf = (float)b, cf = (float)c;
            const float s4 = 0x1.bb9da2p-52f;  // ((s*s)*s)*s, s = 1.f/(4*7*255.f)
            const float kk = 0x1.47ae14p-5f;   // 0.04f
            const float t1 = __fmul_rn(af, bf);
            const float t2 = __fmul_rn(cf, cf);
            const float sm = __fadd_rn(af, bf);
            const float t4 = __fmul_rn(__fmul_rn(kk, sm), sm);
            const float resp = __fmul_rn(__fsub_rn(__fsub_rn(t1, t2), t4), s4);
            s_key[e] = ((unsigned long long)float_desc_key(resp) << 32) | xy;
        }
        __syncwarp();
    }
    for (int i = n + tid; i < n2; i += HS_THREADS) s_key[i] = 0xFFFFFFFFFFFFFFFFull;
    __syncthreads();

    // bitonic sort ascending
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += HS_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long A = s_key[i], B = s_key[p];
                    const bool up = (i & k) == 0;
                    if ((A > B) == up) {
                        s_key[i] = B;
                        s_key[p] = A;
                    }
                }
            }
            __syncthreads();
        }
    }
    // retainBest(N) on the Harris response, ties with the N-th kept
    if (n > N) {
        const uint32_t cut_hi = (uint32_t)(s_key[N - 1] >> 32);
        for (int i = tid; i < n; i += HS_THREADS) {
            const bool in = (uint32_t)(s_key[i] >> 32) <= cut_hi;
            const bool nxt = (i + 1 < n) ? ((uint32_t)(s_key[i + 1] >> 32) <= cut_hi) : false;
            if (in && !nxt) s_kept = i + 1;
        }
    } else if (tid == 0) {
        s_kept = n;
    }
    __syncthreads();
    const int kept = s_kept;
    uint2* so = sel + ((size_t)img * ORB_NL + l) * SORT_CAP;
    for (int i = tid; i < kept; i += HS_THREADS) {
        const unsigned long long K = s_key[i];
        uint32_t b = ~(uint32_t)(K >> 32);
        b ^= (b >> 31) ? 0x80000000u : 0xFFFFFFFFu;  // inverse of the monotone map
        so[i] = make_uint2((uint32_t)K, b);
    }
    if (tid == 0) C->sel_cnt[l] = (uint32_t)kept;
}

// ------------------------------------------------------------------------------------------------------------
// K8  descriptor blur (see oracle/orb_restate.blur7 for how the op order was pinned against cv2)
// ------------------------------------------------------------------------------------------------------------
// Tile = BL_W x BL_H pixels.  Thread (warp = strip of BL_SH rows, lane = quad of 4 columns) marches down its strip:
// per input row one horizontal pass for its 4 columns (10 input bytes -> 4 floats), kept in a 7-row register window,
// and from the 7th row on one vertical pass + rounding + one packed 32-bit store per row.  The float intermediate
// never leaves registers; the horizontal pass is recomputed for the 6 halo rows of a strip (22 / 16).
#define BL_W 128
#define BL_H 128
#define BL_SH 16                         // rows per strip
#define BL_THREADS ((BL_W / 4) * (BL_H / BL_SH))
#define BL_ROWS (BL_H + 6)               // input rows gy = ty0-3 .. ty0+BL_H+2
#define BL_WORDS 40                      // input bytes gx = tx0-16 .. tx0+143 as 40 words per row (10 aligned 16-byte chunks)
static_assert(BL_W == 128 && BL_THREADS == 256, "one lane per quad, one warp per strip");

__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return p;
}

// byte k (0..3) of w as float, exactly: 0x4B0000bb is 8388608 + b
template <int K>
__device__ __forceinline__ float byte_to_float(uint32_t w) {
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 | K)), 8388608.f);
}

__global__ void __launch_bounds__(BL_THREADS)
blur_kernel(ImgSrc src, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, const __grid_constant__ OrbGeom g) {
    __shared__ __align__(16) uint32_t s_in[BL_ROWS * BL_WORDS];
    const float k0 = 0x1.1f5f62p-4f, k1 = 0x1.0c70fcp-3f, k2 = 0x1.869472p-3f, k3 = 0x1.ba95c0p-3f;

    const int img = blockIdx.y;
    int l = 0;
#pragma unroll
    for (int i = 1; i < ORB_NL; ++i)
        if ((int)blockIdx.x >= g.lv[i].tile_begin) l = i;
    const OrbLevel& L = g.lv[l];
    const int t = blockIdx.x - L.tile_begin;
    const int tyi = t / L.tiles_x;
    const int tx0 = (t - tyi * L.tiles_x) * BL_W, ty0 = tyi * BL_H;
    int pitch;
    const uint8_t* im = level_ptr(src, pyr, g, img, l, pitch);
    const int tid = threadIdx.x;
    const int W = L.w, H = L.h;
    const int n_rows = min(BL_H, H - ty0) + 6;  // input rows some output row of this tile reads

    // load: BORDER_REFLECT_101.  Rows reflect through their index; columns are fetched as aligned 16-byte chunks
    // (zero outside the pitch) and the three reflected bytes on either side of the image are patched in afterwards.
    const bool aligned16 = ((pitch & 15) == 0) && (((uintptr_t)im & 15) == 0);
    if (aligned16) {
        for (int i = tid; i < n_rows * 10; i += BL_THREADS) {  // cp.async: the bytes are consumed as they lie
            const int r = i / 10, ch = i - r * 10;
            const int gy = reflect101(min(ty0 - 3 + r, H + 2), H), gx = tx0 - 16 + 16 * ch;
            uint32_t* dst = &s_in[r * BL_WORDS + 4 * ch];
            if (gx >= 0 && gx < pitch) __pipeline_memcpy_async(dst, im + (size_t)gy * pitch + gx, 16);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
        __pipeline_commit();
        __pipeline_wait_prior(0);
        __syncthreads();
        uint8_t* s8 = reinterpret_cast<uint8_t*>(s_in);
        const bool fix_l = tx0 == 0, fix_r = W < tx0 + BL_W + 3;
        for (int r = tid; r < n_rows; r += BL_THREADS) {
            uint8_t* row = s8 + r * (BL_WORDS * 4) + 16 - tx0;  // row[gx]
            if (fix_r) {  // first: the left patch may read a byte this one writes when W is tiny
#pragma unroll
                for (int k = 0; k < 3; ++k) row[W + k] = row[W - 2 - k];
            }
            if (fix_l) {
#pragma unroll
                for (int k = 1; k <= 3; ++k) row[-k] = row[k];
            }
        }
    } else {
        for (int i = tid; i < n_rows * BL_WORDS; i += BL_THREADS) {
            const int r = i / BL_WORDS, wj = i - r * BL_WORDS;
            const uint8_t* row = im + (size_t)reflect101(min(ty0 - 3 + r, H + 2), H) * pitch;
            const int gx = tx0 - 16 + 4 * wj;
            uint32_t px = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) px |= (uint32_t)row[reflect101(min(max(gx + b, -3), W + 2), W)] << (8 * b);
            s_in[i] = px;
        }
    }
    __syncthreads();

    const int q = tid & 31, strip = tid >> 5;
    const int gx = tx0 + 4 * q, y0 = ty0 + strip * BL_SH;
    if (gx >= W || y0 >= H) return;
    // cv2's AVX2 row filter: fused 32-wide body, unfused scalar tail (x >= 32 * (W / 32)); a quad never straddles it
    const bool fused = gx < (W / 32) * 32;
    uint8_t* out = blur + (size_t)img * g.img_slab + L.off + (size_t)y0 * L.pitch + gx;
    const uint32_t* in = &s_in[strip * BL_SH * BL_WORDS + 3 + q];
    float win[7][4];
#pragma unroll
    for (int j = 0; j < BL_SH + 6; ++j) {
        if (y0 + j - 6 >= H) break;  // warp-uniform: the rest of the strip lies below the image
        const uint32_t w0 = in[j * BL_WORDS], w1 = in[j * BL_WORDS + 1], w2 = in[j * BL_WORDS + 2];
        float f[10];
        f[0] = byte_to_float<1>(w0); f[1] = byte_to_float<2>(w0); f[2] = byte_to_float<3>(w0);
        f[3] = byte_to_float<0>(w1); f[4] = byte_to_float<1>(w1); f[5] = byte_to_float<2>(w1); f[6] = byte_to_float<3>(w1);
        f[7] = byte_to_float<0>(w2); f[8] = byte_to_float<1>(w2); f[9] = byte_to_float<2>(w2);
        float o[4];
        if (fused) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float acc = __fmul_rn(k0, f[k]);
                acc = fmaf(k1, f[k + 1], acc);
                acc = fmaf(k2, f[k + 2], acc);
                acc = fmaf(k3, f[k + 3], acc);
                acc = fmaf(k2, f[k + 4], acc);
                acc = fmaf(k1, f[k + 5], acc);
                o[k] = fmaf(k0, f[k + 6], acc);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float acc = __fmul_rn(k0, f[k]);
                acc = __fadd_rn(acc, __fmul_rn(k1, f[k + 1]));
                acc = __fadd_rn(acc, __fmul_rn(k2, f[k + 2]));
                acc = __fadd_rn(acc, __fmul_rn(k3, f[k + 3]));
                acc = __fadd_rn(acc, __fmul_rn(k2, f[k + 4]));
                acc = __fadd_rn(acc, __fmul_rn(k1, f[k + 5]));
                o[k] = __fadd_rn(acc, __fmul_rn(k0, f[k + 6]));
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) win[j % 7][k] = o[k];
        if (j >= 6) {
            // column pass for output row y0 + j - 6: window rows (j-6 .. j) sit in slots (j-6) % 7 .. j % 7; symmetric
            // pairing, fused, rounded once
            uint32_t packed = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float acc = __fmul_rn(k3, win[(j - 3) % 7][c]);
                acc = fmaf(__fadd_rn(win[(j - 2) % 7][c], win[(j - 4) % 7][c]), k2, acc);
                acc = fmaf(__fadd_rn(win[(j - 1) % 7][c], win[(j - 5) % 7][c]), k1, acc);
                acc = fmaf(__fadd_rn(win[j % 7][c], win[(j - 6) % 7][c]), k0, acc);
                int v = __float2int_rn(acc);
                v = max(0, min(255, v));
                packed |= (uint32_t)v << (8 * c);
            }
            // gx is a multiple of 4 below W and the plane pitch is a multiple of 16: the word stays inside the row
            *reinterpret_cast<uint32_t*>(out + (size_t)(j - 6) * L.pitch) = packed;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// helpers shared by ANMS and describe: locate keypoint g of an image in the per-level selections
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int locate_level(const uint32_t* sel_cnt, int gidx, int& j) {
    int l = ORB_NL - 1, base = 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < ORB_NL; ++i) {
        const int c = (int)sel_cnt[i];
        if (!found) {
            if (gidx < base + c) {
                found = true;
                l = i;
            } else if (i < ORB_NL - 1) {
                base += c;
            }
        }
    }
    j = gidx - base;
    return l;
}

// ------------------------------------------------------------------------------------------------------------
// K7  ANMS (the reference's own code, visual_odometry.cpp:96-157), one CTA per image.
//   radius_i = min over {j : response_j > response_i * 1.11f} of sqrt((double)dx*dx + (double)dy*dy), dx,dy float32
//   keep radius_i >= (num-th largest radius); output order = canonical order (what cv::ORB::compute's stable
//   by-octave regrouping produces from the response-sorted list)
// ------------------------------------------------------------------------------------------------------------
#define ANMS_THREADS 1024

// radii: grid (slices, images); every warp owns one keypoint i at a time and scans the stronger ones.  min_j sqrt(d2_j) equals
// sqrt(min_j d2_j) bit for bit (sqrt is monotone and correctly rounded), so the scan keeps the squared distance --
// computed exactly as the reference's operands, float differences widened to double -- and takes one sqrt at the end.
__global__ void __launch_bounds__(256)
anms_radius_kernel(const __grid_constant__ OrbGeom g, const uint2* __restrict__ sel, const ImgCounters* __restrict__ cnt,
                   int kp_cap, int num, float c_robust, double* __restrict__ rad) {
    const int img = blockIdx.y;
    const ImgCounters* C = &cnt[img];
    int n = 0;
#pragma unroll
    for (int i = 0; i < ORB_NL; ++i) n += (int)C->sel_cnt[i];
    if (n > kp_cap) n = kp_cap;
    if (n < num) return;  // ANMS is a no-op for this image (visual_odometry.cpp:100)
    double* R = rad + (size_t)img * kp_cap;
    const uint2* S = sel + (size_t)img * ORB_NL * SORT_CAP;
    // one warp per keypoint i, lanes stride over the stronger keypoints of every level, warp-min at the end
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        int j;
        const int l = locate_level(C->sel_cnt, i, j);
        const uint2 e = S[(size_t)l * SORT_CAP + j];
        const float sc = g.lv[l].scale;
        const float xi = __fmul_rn((float)(e.x & 0xFFFF), sc), yi = __fmul_rn((float)(e.x >> 16), sc);
        const float thr = __fmul_rn(__uint_as_float(e.y), c_robust);
        double best2 = 1.7976931348623157e308;
        for (int l2 = 0; l2 < ORB_NL; ++l2) {
            const int c2 = (int)C->sel_cnt[l2];
            const float sc2 = g.lv[l2].scale;
            const uint2* S2 = S + (size_t)l2 * SORT_CAP;
            for (int q = lane; q < c2; q += 32) {
                const uint2 f = S2[q];
                if (!(__uint_as_float(f.y) > thr)) break;  // per-level lists are response-descending
                const float dx = __fsub_rn(xi, __fmul_rn((float)(f.x & 0xFFFF), sc2));
                const float dy = __fsub_rn(yi, __fmul_rn((float)(f.x >> 16), sc2));
                best2 = fmin(best2, __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best2 = fmin(best2, __shfl_xor_sync(0xFFFFFFFFu, best2, o));
        if (lane == 0) R[i] = best2 < 1.7976931348623157e308 ? sqrt(best2) : best2;
    }
}

__global__ void __launch_bounds__(ANMS_THREADS)
anms_kernel(const __grid_constant__ OrbGeom g, const uint2* __restrict__ sel, ImgCounters* __restrict__ cnt, int kp_cap, int num,
            float c_robust, double* __restrict__ rad, uint32_t* __restrict__ keep) {
    extern __shared__ unsigned long long s_sort[];  // n2 doubles (as ordered bits)
    __shared__ uint32_t s_warp[ANMS_THREADS / 32];
    __shared__ int s_base;
    const int img = blockIdx.x;
    ImgCounters* C = &cnt[img];
    const int tid = threadIdx.x;
    int n = 0;
#pragma unroll
    for (int i = 0; i < ORB_NL; ++i) n += (int)C->sel_cnt[i];
    if (n > kp_cap) n = kp_cap;
    uint32_t* kp_keep = keep + (size_t)img * kp_cap;
    if (n < num) {  // reference: no-op when fewer than num keypoints (visual_odometry.cpp:100)
        for (int i = tid; i < n; i += ANMS_THREADS) kp_keep[i] = i;
        if (tid == 0) C->n_keep = n;
        return;
    }
    const double* R = rad + (size_t)img * kp_cap;  // written by anms_radius_kernel
    __syncthreads();
    // num-th largest radius by bitonic sort (descending) of the (positive) doubles
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = tid; i < n2; i += ANMS_THREADS)
        s_sort[i] = i < n ? ~(unsigned long long)__double_as_longlong(R[i]) : 0xFFFFFFFFFFFFFFFFull;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += ANMS_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long A = s_sort[i], B = s_sort[p];
                    const bool up = (i & k) == 0;
                    if ((A > B) == up) {
                        s_sort[i] = B;
                        s_sort[p] = A;
                    }
                }
            }
            __syncthreads();
        }
    }
    const double final_radius = __longlong_as_double((long long)~s_sort[num - 1]);
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += ANMS_THREADS) {
        const int i = i0 + tid;
        const bool k = i < n && R[i] >= final_radius;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, k);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < ANMS_THREADS / 32; ++w) {
            const uint32_t c = s_warp[w];
            before += (w < (tid >> 5)) ? c : 0;
            total += c;
        }
        if (k) kp_keep[s_base + before + __popc(bal & ((1u << (tid & 31)) - 1))] = i;
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0) C->n_keep = s_base;
}

// ------------------------------------------------------------------------------------------------------------
// K6 + K9  orientation and rotated BRIEF, one warp per keypoint
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float p1 = 0x1.ca44dep+5f, p3 = -0x1.2aaddcp+4f, p5 = 0x1.1d3f7ep+3f, p7 = -0x1.4515b2p+1f;
    const float eps = 0x1.0p-52f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// (A first version of the shared-memory window staged it through registers -- word loads, word stores, 110 registers --
// and was slower than the direct gathers, 1.14 -> 1.94 ms per 512 images; the cp.async form below needs none.)
// One warp describes DESC_KPW keypoints: the intensity-centroid sums and the 256 tests of each keypoint are spread
// over the 32 lanes, while the per-keypoint scalar work (fastAtan2, the double-precision cos/sin OpenCV uses, the
// keypoint record) is done once with lane i owning keypoint i instead of 32 times redundantly.
__global__ void __launch_bounds__(DESC_WARPS * 32, DESC_MIN_CTAS)
describe_kernel(ImgSrc src, const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, const __grid_constant__ OrbGeom g,
                const uint2* __restrict__ sel, ImgCounters* __restrict__ cnt, const uint32_t* __restrict__ keep,
                int use_keep, const float* __restrict__ pattern, int kp_cap, vslam_keypoint* __restrict__ kp_out,
                uint8_t* __restrict__ desc_out, int32_t* __restrict__ n_out, uint32_t* __restrict__ sticky) {
    const int img = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int k0 = (blockIdx.x * DESC_WARPS + (threadIdx.x >> 5)) * DESC_KPW;
    ImgCounters* C = &cnt[img];
    int total = 0;
#pragma unroll
    for (int i = 0; i < ORB_NL; ++i) total += (int)C->sel_cnt[i];
    if (total > kp_cap) {
        if (k0 == 0 && lane == 0) {
            atomicOr(&C->flags, 4u);
            atomicOr(sticky, 4u);
        }
        total = kp_cap;
    }
    const int oslot = img < src.per_base ? src.out_slot[0] + img : src.out_slot[1] + (img - src.per_base);
    const int n = use_keep ? (int)C->n_keep : total;
    if (k0 == 0 && lane == 0) n_out[oslot] = n;
    if (k0 >= n) return;
    const int nk = min(DESC_KPW, n - k0);  // keypoints of this warp

    // lane i (mod DESC_KPW) owns keypoint k0 + i
    const int slot = lane & (DESC_KPW - 1);
    const int my_k = min(k0 + slot, n - 1);
    const int gi = use_keep ? (int)keep[(size_t)img * kp_cap + my_k] : my_k;
    int j;
    const int my_l = locate_level(C->sel_cnt, gi, j);
    const uint2 my_e = sel[((size_t)img * ORB_NL + my_l) * SORT_CAP + j];
    const int my_x = my_e.x & 0xFFFF, my_y = my_e.x >> 16;

    // intensity centroid over the radius-15 disc (orb.cpp ICAngles).  Lane (g, q) = (lane >> 3, lane & 7) reads the four
    // pixels u = -16 + 4q .. -13 + 4q of row v = -15 + 4t + g in step t: eight steps cover the 31 rows, the disc is a byte
    // mask per (step, lane) built once per warp, and both moments are byte dot products (DP4A): m10 += sum u * I,
    // m01 += v * sum I.  Integer sums: the order does not matter.
    __shared__ uint32_t s_icm[DESC_WARPS][8][32];
    const int wq = threadIdx.x >> 5;
    const int ig = lane >> 3, iq = lane & 7, u0 = -16 + 4 * iq;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int v = -15 + 4 * t + ig;
        const int um = v <= 15 ? c_umax[v < 0 ? -v : v] : -1;
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (abs(u0 + k) <= um) m |= 0xFFu << (8 * k);
        s_icm[wq][t][lane] = m;
    }
    __syncwarp();
    const uint32_t uw = (uint32_t)(u0 & 0xFF) | ((uint32_t)((u0 + 1) & 0xFF) << 8) | ((uint32_t)((u0 + 2) & 0xFF) << 16) |
                        ((uint32_t)((u0 + 3) & 0xFF) << 24);
    int my_m10 = 0, my_m01 = 0;
#pragma unroll 2
    for (int i = 0; i < nk; ++i) {
        const int x = __shfl_sync(0xFFFFFFFFu, my_x, i), y = __shfl_sync(0xFFFFFFFFu, my_y, i);
        const int l = __shfl_sync(0xFFFFFFFFu, my_l, i);
        int pitch;
        const uint8_t* im = level_ptr(src, pyr, g, img, l, pitch);
        const uint8_t* rowp = im + (size_t)(y - 15 + ig) * pitch + (x + u0);
        int m10 = 0, m01 = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const uint8_t* p = rowp + (size_t)(4 * t) * pitch;
            const uint32_t a = (uint32_t)(uintptr_t)p & 3u;
            const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
            const uint32_t word = __funnelshift_r(__ldg(q), __ldg(q + 1), 8 * a) & s_icm[wq][t][lane];
            const uint32_t vw = (uint32_t)((-15 + 4 * t + ig) & 0xFF) * 0x01010101u;
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10) : "r"(word), "r"(uw));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m01) : "r"(word), "r"(vw));
        }
        m10 = __reduce_add_sync(0xFFFFFFFFu, m10);
        m01 = __reduce_add_sync(0xFFFFFFFFu, m01);
        if (slot == i) {
            my_m10 = m10;
            my_m01 = m01;
        }
    }
    const OrbLevel& ML = g.lv[my_l];
    const float angle = fast_atan2_deg((float)my_m01, (float)my_m10);
    const float ptx = __fmul_rn((float)my_x, ML.scale), pty = __fmul_rn((float)my_y, ML.scale);
    if (lane < nk) {
        vslam_keypoint kp;
        kp.x = ptx;
        kp.y = pty;
        kp.size = __fmul_rn(31.f, ML.scale);
        kp.angle = angle;
        kp.response = __uint_as_float(my_e.y);
        kp.octave = my_l;
        kp.class_id = -1;
        kp_out[(size_t)oslot * kp_cap + k0 + lane] = kp;
    }
    const int my_cx = __float2int_rn(__fmul_rn(ptx, ML.inv_scale));
    const int my_cy = __float2int_rn(__fmul_rn(pty, ML.inv_scale));
    const float th = __fmul_rn(angle, 0x1.1df46ap-6f);  // (float)(CV_PI/180)
    const float my_a = (float)cos((double)th), my_b = (float)sin((double)th);

    // rotated BRIEF on the blurred level (orb.cpp computeOrbDescriptors, WTA_K = 2); lane = output byte.
    // The 512 samples of a keypoint are byte gathers scattered over a 37 x 37 window: straight from global memory a
    // warp-wide gather touches up to 32 cache lines, and the kernel was bound by exactly those L1 wavefronts.  The window
    // (39 rows x 64 bytes from the 16-byte boundary left of it) is therefore staged into a per-warp shared-memory patch
    // with cp.async -- no registers in between, the next keypoint's window in flight while this one is sampled.
    const float4* pat = reinterpret_cast<const float4*>(pattern) + lane * 8;
    float4 pt[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) pt[t] = __ldg(&pat[t]);  // (x0, y0, x1, y1) of test 8*lane + t
    __shared__ __align__(16) uint8_t s_patch[DESC_WARPS][2][DESC_PR * DESC_PW];
    const int wib = threadIdx.x >> 5;
    const uint8_t* bplane = blur + (size_t)img * g.img_slab;
    auto stage = [&](int i, int buf) {
        const int cx = __shfl_sync(0xFFFFFFFFu, my_cx, i), cy = __shfl_sync(0xFFFFFFFFu, my_cy, i);
        const int l = __shfl_sync(0xFFFFFFFFu, my_l, i);
        const int bp = g.lv[l].pitch;
        const uint8_t* srcp = bplane + g.lv[l].off + (size_t)(cy - DESC_PH) * bp + ((cx - DESC_PH) & ~15);
        uint8_t* dstp = s_patch[wib][buf];
#pragma unroll
        for (int q0 = 0; q0 < DESC_PR * (DESC_PW / 16); q0 += 32) {
            const int q = q0 + lane;
            if (q < DESC_PR * (DESC_PW / 16)) {
                const int row = q / (DESC_PW / 16), ch = q % (DESC_PW / 16);
                __pipeline_memcpy_async(dstp + row * DESC_PW + 16 * ch, srcp + (size_t)row * bp + 16 * ch, 16);
            }
        }
        __pipeline_commit();
    };
    stage(0, 0);
    for (int i = 0; i < nk; ++i) {
        if (i + 1 < nk) stage(i + 1, (i + 1) & 1);  // warp-uniform
        else __pipeline_commit();                    // keep one group per iteration so that wait_prior(1) means "patch i landed"
        __pipeline_wait_prior(1);
        __syncwarp();
        const float a = __shfl_sync(0xFFFFFFFFu, my_a, i), b = __shfl_sync(0xFFFFFFFFu, my_b, i);
        const int cx = __shfl_sync(0xFFFFFFFFu, my_cx, i);
        // sample (ix, iy) lives at row iy + DESC_PH, column ix + (cx - window start)
        const uint8_t* center = s_patch[wib][i & 1] + DESC_PH * DESC_PW + (cx - ((cx - DESC_PH) & ~15));
        uint32_t byte = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const float4 p0 = pt[t];
            // cvRound: adding 1.5 * 2^23 rounds to the nearest integer, ties to even, exactly like F2I.RN (|v| < 2^22)
            const float rx0 = __fadd_rn(__fsub_rn(__fmul_rn(p0.x, a), __fmul_rn(p0.y, b)), 12582912.f);
            const float ry0 = __fadd_rn(__fadd_rn(__fmul_rn(p0.x, b), __fmul_rn(p0.y, a)), 12582912.f);
            const float rx1 = __fadd_rn(__fsub_rn(__fmul_rn(p0.z, a), __fmul_rn(p0.w, b)), 12582912.f);
            const float ry1 = __fadd_rn(__fadd_rn(__fmul_rn(p0.z, b), __fmul_rn(p0.w, a)), 12582912.f);
            const int ix0 = __float_as_int(rx0) - 0x4B400000, iy0 = __float_as_int(ry0) - 0x4B400000;
            const int ix1 = __float_as_int(rx1) - 0x4B400000, iy1 = __float_as_int(ry1) - 0x4B400000;
            const int t0 = center[iy0 * DESC_PW + ix0];
            const int t1 = center[iy1 * DESC_PW + ix1];
            byte |= (uint32_t)(t0 < t1) << t;
        }
        desc_out[((size_t)oslot * kp_cap + k0 + i) * 32 + lane] = (uint8_t)byte;
        __syncwarp();  // everyone has sampled buffer i & 1 before the next iteration's copies overwrite it
    }
    __pipeline_wait_prior(0);
}

// ------------------------------------------------------------------------------------------------------------
// host side: geometry, lifetime, orchestration
// ------------------------------------------------------------------------------------------------------------
static inline int cv_round_f(float v) { return (int)lrintf(v); }  // round-half-even like cvRound

// cv::ORB per-level feature quotas (orb.cpp detectAndCompute -> computeKeyPoints), float32 arithmetic
static void orb_quotas(int nfeatures, OrbQuota* q) {
    const float factor = (float)(1.0 / (double)1.2f);
    float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)ORB_NL));
    int sum = 0;
    for (int l = 0; l < ORB_NL - 1; ++l) {
        q->n[l] = (int)lrint((double)nd);
        sum += q->n[l];
        nd *= factor;
    }
    q->n[ORB_NL - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

// cv::resize(INTER_LINEAR_EXACT) coefficient tables: entry = (src offset << 16) | c1, c0 = 256 - c1
static void resize_table(int dst, int srcn, uint32_t* out) {
    const double scale = 1.0 / ((double)dst / (double)srcn);
    for (int v = 0; v < dst; ++v) {
        const double f = scale * (v + 0.5) - 0.5;
        int i = (int)floor(f);
        int c1 = 0;
        if (i < 0) {
            i = 0;
        } else if (i >= srcn - 1) {
            i = srcn - 1;
        } else {
            c1 = (int)lrint((f - i) * 256.0);
        }
        out[v] = ((uint32_t)i << 16) | (uint32_t)c1;
    }
}

// The x table in the form resize_level_kernel consumes: 8 words per quad of output columns --
//   [0] first source column of the quad | last needed byte (relative) << 16
//   [1] four byte selectors (d0 | d1 << 4: the two source bytes of pixel k relative to the first column), one per byte
//   [2], [3] unused     [4..7] coefficient pairs (256 - c1) | c1 << 16 of the four pixels
// Columns past dst (row padding) repeat the quad's first column.
static void resize_quad_table(int dst, int srcn, uint32_t* out) {
    std::vector<uint32_t> t((size_t)dst);
    resize_table(dst, srcn, t.data());
    for (int q = 0; 4 * q < dst; ++q) {
        uint32_t* o = out + 8 * (size_t)q;
        const int ox0 = (int)(t[4 * q] >> 16);
        int last = 0;
        uint32_t sel = 0;
        for (int k = 0; k < 4; ++k) {
            const uint32_t e = 4 * q + k < dst ? t[4 * q + k] : t[4 * q];
            const int ox = (int)(e >> 16), c1 = (int)(e & 0xFFFF);
            const int d0 = ox - ox0, d1 = (ox + 1 < srcn ? ox + 1 : srcn - 1) - ox0;
            sel |= (uint32_t)(d0 | (d1 << 4)) << (8 * k);
            o[4 + k] = (uint32_t)(256 - c1) | ((uint32_t)c1 << 16);
            if (d1 > last) last = d1;
        }
        o[0] = (uint32_t)ox0 | ((uint32_t)last << 16);
        o[1] = sel;
        o[2] = o[3] = 0;
    }
}

static int orb_set_geometry(vslam_ctx* ctx, int w, int h) {
    OrbState* o = ctx->orb;
    if (o->geom_valid && o->geom.w == w && o->geom.h == h) return VSLAM_OK;
    OrbGeom g;
    memset(&g, 0, sizeof(g));
    g.w = w;
    g.h = h;
    int off = 0, tiles = 0, ft = 0, cand = 0, tab = 0;
    for (int l = 0; l < ORB_NL; ++l) {
        OrbLevel& L = g.lv[l];
        L.scale = h_scales[l];
        L.inv_scale = 1.f / L.scale;
        L.w = cv_round_f((float)w / L.scale);
        L.h = cv_round_f((float)h / L.scale);
        // a level narrower than 2 * 31 px simply yields no keypoints (border filter), as in cv::ORB; only degenerate
        // sizes are refused
        if (L.w < 16 || L.h < 16) return VSLAM_E_INVALID;
        L.pitch = (L.w + 15) & ~15;
        L.off = off;
        off += L.pitch * L.h;
        off = (off + 255) & ~255;
        L.tiles_x = ceil_div(L.w, BL_W);
        L.tiles_y = ceil_div(L.h, BL_H);
        L.tile_begin = tiles;
        tiles += L.tiles_x * L.tiles_y;
        // kept frame: columns [FT_X0, w-31), rows [31, h-31); empty for a level narrower than the border allows
        L.ft_x = L.w - ORB_EDGE > FT_X0 ? ceil_div(L.w - ORB_EDGE - FT_X0, FT_W) : 0;
        L.ft_y = L.h - ORB_EDGE > FT_Y0 ? ceil_div(L.h - ORB_EDGE - FT_Y0, FT_H) : 0;
        if (L.ft_x == 0 || L.ft_y == 0) L.ft_x = L.ft_y = 0;
        L.ft_begin = ft;
        ft += L.ft_x * L.ft_y;
        L.cand_off = cand;
        L.cand_cap = (L.w * L.h) / 12 + 64;
        cand += L.cand_cap;
        L.xtab_off = tab;  // 16-byte aligned regions; x: 8 words per quad of columns (resize_quad_table)
        tab += 8 * ((L.w + 3) / 4);
        L.ytab_off = tab;
        tab += (L.h + 3) & ~3;
    }
    g.total_tiles = tiles;
    g.total_ft = ft;
    g.img_slab = off;
    g.cand_slab = cand;
    if ((size_t)off > o->slab_cap || (size_t)cand > o->cand_cap_total || tab > o->tab_cap) return VSLAM_E_CAPACITY;
    std::vector<uint32_t> t((size_t)tab, 0);
    for (int l = 1; l < ORB_NL; ++l) {
        resize_quad_table(g.lv[l].w, g.lv[l - 1].w, &t[g.lv[l].xtab_off]);
        resize_table(g.lv[l].h, g.lv[l - 1].h, &t[g.lv[l].ytab_off]);
    }
    VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VSLAM_CUDA(ctx, cudaMemcpy(o->d_tab, t.data(), (size_t)tab * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (ft > o->ftab_cap) return VSLAM_E_CAPACITY;
    {
        std::vector<uint32_t> ftab((size_t)(ft > 0 ? ft : 1), 0u);
        for (int l = 0; l < ORB_NL; ++l)
            for (int ty = 0; ty < g.lv[l].ft_y; ++ty)
                for (int tx = 0; tx < g.lv[l].ft_x; ++tx)
                    ftab[(size_t)g.lv[l].ft_begin + ty * g.lv[l].ft_x + tx] = (uint32_t)l | ((uint32_t)tx << 4) | ((uint32_t)ty << 16);
        VSLAM_CUDA(ctx, cudaMemcpy(o->d_ftab, ftab.data(), ftab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    o->geom = g;
    o->geom_valid = true;
    return VSLAM_OK;
}

int vslam_orb_init(vslam_ctx* ctx) {
    OrbState* o = (OrbState*)calloc(1, sizeof(OrbState));
    if (!o) return VSLAM_E_INVALID;
    ctx->orb = o;
    const vslam_config& c = ctx->cfg;
    o->max_images = c.max_images;
    o->kp_cap = c.max_keypoints;
    if (c.max_images <= 0 || c.max_width <= 0 || c.max_height <= 0 || c.max_keypoints <= 0) return VSLAM_OK;  // ORB disabled
    // capacity for the largest geometry: sum over levels of pitch*h < 3.4 * w*h (+ alignment)
    const size_t wh = (size_t)((c.max_width + 15) & ~15) * c.max_height;
    o->slab_cap = (size_t)(wh * 3.4) + 8 * 256 + 4096;
    o->cand_cap_total = (size_t)(wh * 3.4 / 12) + 8 * 64 + 1024;
    o->tab_cap = (int)(c.max_width * 9.6 + c.max_height * 4.8) + 512;  // x: 2 words per column and level, y: 1
    o->in_pitch = (c.max_width + 15) & ~15;
    const size_t ni = (size_t)c.max_images;
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_pyr, ni * o->slab_cap));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_blur, ni * o->slab_cap));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_tab, (size_t)o->tab_cap * sizeof(uint32_t)));
    o->ftab_cap = (int)((c.max_width / FT_W + 2) * (c.max_height / FT_H + 2) * 3.4) + ORB_NL * 4;  // sum over the levels < 3.28 x level 0
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_ftab, (size_t)o->ftab_cap * sizeof(uint32_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_cand, ni * o->cand_cap_total * sizeof(uint2)));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_sel, ni * ORB_NL * SORT_CAP * sizeof(uint2)));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_cnt, ni * sizeof(ImgCounters)));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_sticky, sizeof(uint32_t)));
    VSLAM_CUDA(ctx, cudaMemset(o->d_sticky, 0, sizeof(uint32_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_keep, ni * o->kp_cap * sizeof(uint32_t)));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_rad, ni * o->kp_cap * sizeof(double)));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_pattern, 1024 * sizeof(float)));
    {
        float hp[1024];
        for (int i = 0; i < 1024; ++i) hp[i] = (float)h_orb_pattern[i];
        VSLAM_CUDA(ctx, cudaMemcpy(o->d_pattern, hp, sizeof(hp), cudaMemcpyHostToDevice));
    }
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_in, ni * (size_t)o->in_pitch * c.max_height));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_kp, ni * o->kp_cap * sizeof(vslam_keypoint)));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_desc, ni * o->kp_cap * 32));
    VSLAM_CUDA(ctx, cudaMalloc(&o->d_n, ni * sizeof(int32_t)));
    VSLAM_CUDA(ctx, cudaMallocHost(&o->h_kp, ni * o->kp_cap * sizeof(vslam_keypoint)));
    VSLAM_CUDA(ctx, cudaMallocHost(&o->h_desc, ni * o->kp_cap * 32));
    VSLAM_CUDA(ctx, cudaMallocHost(&o->h_n, ni * sizeof(int32_t)));
    VSLAM_CUDA(ctx, cudaMallocHost(&o->h_cnt, ni * sizeof(ImgCounters)));
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(harris_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SORT_CAP * (int)sizeof(unsigned long long)));
    VSLAM_CUDA(ctx, cudaFuncSetAttribute(anms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    return VSLAM_OK;
}

void vslam_orb_free(vslam_ctx* ctx) {
    OrbState* o = ctx->orb;
    if (!o) return;
    cudaFree(o->d_pyr);
    cudaFree(o->d_blur);
    cudaFree(o->d_tab);
    cudaFree(o->d_ftab);
    cudaFree(o->d_cand);
    cudaFree(o->d_sel);
    cudaFree(o->d_cnt);
    cudaFree(o->d_sticky);
    cudaFree(o->d_keep);
    cudaFree(o->d_rad);
    cudaFree(o->d_pattern);
    cudaFree(o->d_in);
    cudaFree(o->d_kp);
    cudaFree(o->d_desc);
    cudaFree(o->d_n);
    cudaFreeHost(o->h_kp);
    cudaFreeHost(o->h_desc);
    cudaFreeHost(o->h_n);
    cudaFreeHost(o->h_cnt);
    free(o);
    ctx->orb = nullptr;
}

// geometry-dependent tables are uploaded on the context stream; callers that fan out over several streams call this
// first and order their streams after it
int vslam_orb_prepare(vslam_ctx* ctx, int w, int h) {
    if (!ctx->orb) return VSLAM_E_CAPACITY;
    if (w > ctx->cfg.max_width || h > ctx->cfg.max_height || w >= 65536 || h >= 65536) return VSLAM_E_CAPACITY;
    return orb_set_geometry(ctx, w, h);
}

// Enqueue the whole ORB pipeline for n_img images on the context stream (no host synchronisation).
int vslam_orb_enqueue(vslam_ctx* ctx, const ImgSrc& src, int n_img, int w, int h, int nfeatures, int anms_keep,
                      float anms_c, vslam_keypoint* d_kp, uint8_t* d_desc, int32_t* d_n, int scratch0) {
    OrbState* o = ctx->orb;
    if (!o || !o->d_pyr) return VSLAM_E_CAPACITY;
    if (n_img <= 0 || scratch0 < 0 || scratch0 + n_img > o->max_images) return VSLAM_E_CAPACITY;
    if (w > ctx->cfg.max_width || h > ctx->cfg.max_height) return VSLAM_E_CAPACITY;
    if (w >= 65536 || h >= 65536) return VSLAM_E_CAPACITY;
    if (nfeatures <= 0 || nfeatures > o->kp_cap) return VSLAM_E_CAPACITY;
    int st = orb_set_geometry(ctx, w, h);
    if (st != VSLAM_OK) return st;
    const OrbGeom& g = o->geom;
    OrbQuota q;
    orb_quotas(nfeatures, &q);
    cudaStream_t s = ctx->stream;
    // scratch of this call: image slots [scratch0, scratch0 + n_img) -- two chunks on two streams use disjoint slots
    uint8_t* pyr = o->d_pyr + (size_t)scratch0 * g.img_slab;
    uint8_t* blur = o->d_blur + (size_t)scratch0 * g.img_slab;
    uint2* cand = o->d_cand + (size_t)scratch0 * g.cand_slab;
    uint2* sel = o->d_sel + (size_t)scratch0 * ORB_NL * SORT_CAP;
    ImgCounters* cnt = o->d_cnt + scratch0;
    uint32_t* keep = o->d_keep + (size_t)scratch0 * o->kp_cap;
    double* rad = o->d_rad + (size_t)scratch0 * o->kp_cap;
    VSLAM_CUDA(ctx, cudaMemsetAsync(cnt, 0, (size_t)n_img * sizeof(ImgCounters), s));
    for (int l = 1; l < ORB_NL; ++l) {
        dim3 grid(ceil_div(ceil_div(g.lv[l].w, 8), 32), ceil_div(g.lv[l].h, 8), n_img);
        vslam_time_begin(ctx, VK_RESIZE);
        resize_level_kernel<<<grid, dim3(32, 8), 0, s>>>(src, pyr, o->d_tab, g, l);
        vslam_time_end(ctx);
        VSLAM_LAUNCH_CHECK(ctx, "resize_level_kernel");
    }
    vslam_time_begin(ctx, VK_FAST);
    if (g.total_ft > 0) fast_kernel<<<dim3(g.total_ft, n_img), FAST_THREADS, 0, s>>>(src, pyr, g, o->d_ftab, cand, cnt, o->d_sticky);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "fast_kernel");
    vslam_time_begin(ctx, VK_HARRIS_SELECT);
    harris_select_kernel<<<dim3(ORB_NL, n_img), HS_THREADS, SORT_CAP * sizeof(unsigned long long), s>>>(
        src, pyr, g, q, cand, cnt, sel, o->d_sticky);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "harris_select_kernel");
    vslam_time_begin(ctx, VK_BLUR);
    blur_kernel<<<dim3(g.total_tiles, n_img), BL_THREADS, 0, s>>>(src, pyr, blur, g);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "blur_kernel");
    const int use_keep = anms_keep > 0 ? 1 : 0;
    if (use_keep) {
        int n2 = 1;
        while (n2 < o->kp_cap) n2 <<= 1;
        if ((size_t)n2 * 8 > 65536) return VSLAM_E_CAPACITY;
        vslam_time_begin(ctx, VK_ANMS);
        // a single image spreads its radius scan over ~2 CTAs per SM; big batches use one slice per image
        int slices = (4 * ctx->num_sms + n_img - 1) / n_img;
        if (slices > ceil_div(o->kp_cap, 8)) slices = ceil_div(o->kp_cap, 8);  // 8 warps = 8 keypoints per CTA pass
        if (slices < 1) slices = 1;
        anms_radius_kernel<<<dim3(slices, n_img), 256, 0, s>>>(g, sel, cnt, o->kp_cap, anms_keep, anms_c, rad);
        ctx->launches++;
        anms_kernel<<<n_img, ANMS_THREADS, (size_t)n2 * 8, s>>>(g, sel, cnt, o->kp_cap, anms_keep, anms_c, rad, keep);
        vslam_time_end(ctx);
        VSLAM_LAUNCH_CHECK(ctx, "anms_kernel");
    }
    vslam_time_begin(ctx, VK_DESCRIBE);
    describe_kernel<<<dim3(ceil_div(o->kp_cap, DESC_WARPS * DESC_KPW), n_img), DESC_WARPS * 32, 0, s>>>(
        src, pyr, blur, g, sel, cnt, keep, use_keep, o->d_pattern, o->kp_cap, d_kp, d_desc,
        d_n, o->d_sticky);
    vslam_time_end(ctx);
    VSLAM_LAUNCH_CHECK(ctx, "describe_kernel");
    return VSLAM_OK;
}

extern "C" int vslam_orb_keypoint_capacity(const vslam_ctx* ctx) { return ctx && ctx->orb ? ctx->orb->kp_cap : 0; }

extern "C" int vslam_orb_detect_compute_batch_dev(vslam_ctx* ctx, const uint8_t* d_images, int n_images, int width,
                                                  int height, int row_pitch, long long image_stride, int nfeatures,
                                                  int anms_keep, float anms_c, vslam_keypoint* d_kp, uint8_t* d_desc,
                                                  int32_t* d_n) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !d_images || !d_kp || !d_desc || !d_n) return VSLAM_E_INVALID;
    if (width <= 0 || height <= 0 || row_pitch < width) return VSLAM_E_INVALID;
    ImgSrc src;
    src.base[0] = d_images;
    src.base[1] = d_images;
    src.img_stride = image_stride;
    src.pitch = row_pitch;
    src.per_base = n_images > 0 ? n_images : 1;
    src.out_slot[0] = 0;
    src.out_slot[1] = src.per_base;
    return vslam_orb_enqueue(ctx, src, n_images, width, height, nfeatures, anms_keep, anms_c, d_kp, d_desc, d_n, 0);
}

// flags raised by the kernels (bit0 candidate overflow, bit1 sort overflow, bit2 keypoint capacity)
int vslam_orb_check_flags(vslam_ctx* ctx, int n_img) {
    (void)n_img;
    OrbState* o = ctx->orb;
    uint32_t* h = &o->h_cnt[0].flags;  // pinned scratch word
    VSLAM_CUDA(ctx, cudaMemcpyAsync(h, o->d_sticky, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    VSLAM_CUDA(ctx, cudaMemsetAsync(o->d_sticky, 0, sizeof(uint32_t), ctx->stream));
    VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return *h ? VSLAM_E_OVERFLOW : VSLAM_OK;
}

extern "C" int vslam_orb_last_flags(vslam_ctx* ctx, int n_images) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !ctx->orb) return VSLAM_E_INVALID;
    return vslam_orb_check_flags(ctx, n_images);
}

extern "C" int vslam_orb_detect_compute_batch(vslam_ctx* ctx, const uint8_t* images, int n_images, int width,
                                              int height, int row_pitch, long long image_stride, int nfeatures,
                                              int anms_keep, float anms_c, vslam_keypoint* kp_out, uint8_t* desc_out,
                                              int32_t* n_out) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !n_out) return VSLAM_E_INVALID;
    if (!images) return VSLAM_E_INVALID;  // reference: "Could not open or find the image" -> -1 (vo.cpp:73-77)
    if (!kp_out || !desc_out) return VSLAM_E_INVALID;
    OrbState* o = ctx->orb;
    if (!o || !o->d_in) return VSLAM_E_CAPACITY;
    if (n_images <= 0 || n_images > o->max_images) return VSLAM_E_CAPACITY;
    if (width <= 0 || height <= 0 || row_pitch < width) return VSLAM_E_INVALID;
    if (width > ctx->cfg.max_width || height > ctx->cfg.max_height) return VSLAM_E_CAPACITY;
    cudaStream_t s = ctx->stream;
    const size_t dstride = (size_t)o->in_pitch * height;
    for (int i = 0; i < n_images; ++i)
        VSLAM_CUDA(ctx, cudaMemcpy2DAsync(o->d_in + i * dstride, o->in_pitch, images + (size_t)i * image_stride,
                                          row_pitch, width, height, cudaMemcpyHostToDevice, s));
    ImgSrc src;
    src.base[0] = o->d_in;
    src.base[1] = o->d_in;
    src.img_stride = (long long)dstride;
    src.pitch = o->in_pitch;
    src.per_base = n_images;
    src.out_slot[0] = 0;
    src.out_slot[1] = n_images;
    int st = vslam_orb_enqueue(ctx, src, n_images, width, height, nfeatures, anms_keep, anms_c, o->d_kp, o->d_desc, o->d_n, 0);
    if (st != VSLAM_OK) return st;
    VSLAM_CUDA(ctx, cudaMemcpyAsync(o->h_n, o->d_n, n_images * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    st = vslam_orb_check_flags(ctx, n_images);  // synchronises
    if (st != VSLAM_OK) return st;
    for (int i = 0; i < n_images; ++i) {
        const int n = o->h_n[i];
        n_out[i] = n;
        if (n <= 0) continue;
        VSLAM_CUDA(ctx, cudaMemcpyAsync(kp_out + (size_t)i * o->kp_cap, o->d_kp + (size_t)i * o->kp_cap,
                                        (size_t)n * sizeof(vslam_keypoint), cudaMemcpyDeviceToHost, s));
        VSLAM_CUDA(ctx, cudaMemcpyAsync(desc_out + (size_t)i * o->kp_cap * 32, o->d_desc + (size_t)i * o->kp_cap * 32,
                                        (size_t)n * 32, cudaMemcpyDeviceToHost, s));
    }
    VSLAM_CUDA(ctx, cudaStreamSynchronize(s));
    return VSLAM_OK;
}

extern "C" int vslam_orb_detect_compute(vslam_ctx* ctx, const uint8_t* image, int width, int height, int row_pitch,
                                        int nfeatures, int anms_keep, float anms_c, vslam_keypoint* kp_out,
                                        uint8_t* desc_out, int32_t* n_out) {
    VslamDeviceGuard device_guard__(ctx);
    return vslam_orb_detect_compute_batch(ctx, image, 1, width, height, row_pitch, 0, nfeatures, anms_keep, anms_c,
                                          kp_out, desc_out, n_out);
}

// test/debug taps: copy intermediate device state of image `img` to host buffers (any may be NULL)
extern "C" int vslam_orb_debug_read(vslam_ctx* ctx, int img, int level, uint8_t* level_pixels, uint8_t* blurred_pixels,
                                    int* w_out, int* h_out, uint32_t* cand_xy_score, int cand_cap, int* n_cand) {
    VslamDeviceGuard device_guard__(ctx);
    if (!ctx || !ctx->orb || !ctx->orb->geom_valid) return VSLAM_E_INVALID;
    OrbState* o = ctx->orb;
    if (img < 0 || img >= o->max_images || level < 0 || level >= ORB_NL) return VSLAM_E_INVALID;
    const OrbLevel& L = o->geom.lv[level];
    VSLAM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (w_out) *w_out = L.w;
    if (h_out) *h_out = L.h;
    if (level_pixels && level > 0)
        VSLAM_CUDA(ctx, cudaMemcpy2D(level_pixels, L.w, o->d_pyr + (size_t)img * o->geom.img_slab + L.off, L.pitch, L.w,
                                     L.h, cudaMemcpyDeviceToHost));
    if (blurred_pixels)
        VSLAM_CUDA(ctx, cudaMemcpy2D(blurred_pixels, L.w, o->d_blur + (size_t)img * o->geom.img_slab + L.off, L.pitch,
                                     L.w, L.h, cudaMemcpyDeviceToHost));
    if (n_cand) {
        ImgCounters c;
        VSLAM_CUDA(ctx, cudaMemcpy(&c, &o->d_cnt[img], sizeof(c), cudaMemcpyDeviceToHost));
        int n = (int)c.cand_cnt[level];
        if (n > L.cand_cap) n = L.cand_cap;
        *n_cand = n;
        if (cand_xy_score) {
            if (n > cand_cap) n = cand_cap;
            VSLAM_CUDA(ctx, cudaMemcpy(cand_xy_score, o->d_cand + (size_t)img * o->geom.cand_slab + L.cand_off,
                                       (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost));
        }
    }
    return VSLAM_OK;
}
