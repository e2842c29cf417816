// Front end of libcubemap_b200.so: fisheye->cubemap warp + ORB extraction, batched, sm_100a.
//
// Reference path (CPU, one frame at a time):
//   System::CvtFisheyeToCubeMap_reverseQuery_withInterpolation   src/System.cpp:327-355   (5 x cv::remap)
//   ORBextractor::operator()                                      src/ORBExtractor.cpp:838-926
//     ComputePyramid :928-953, ComputeKeyPointsOctTree :739-827 (cv::FAST per 30-px cell, DistributeOctTree
//     :511-737, IC_Angle :48-75), cull :883-904, GaussianBlur :907-908, computeOrbDescriptor :79-118
//
// Kernels (all HBM/L2-bound byte work; no tensor cores on this path):
//   k_warp        canvas face tiles <- bilinear gather of the fisheye frame through a pre-quantised map
//   k_pyramid     level l <- cv::resize(INTER_LINEAR) fixed-point model of level l-1
//   k_fast        per 30-px cell FAST-9/16 arc score, in-cell NMS, 20->7 threshold fallback, candidate append
//   k_distribute  the reference's quadtree, restated as data-parallel rounds over nodes
//   k_describe    IC angle + 7x7 Gaussian of the 43x43 neighbourhood + steered BRIEF, one warp per keypoint
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace cslam {

static const int PATCH_SIZE = 31, HALF_PATCH = 15, EDGE_THRESHOLD = 19;
static const int MAX_LEVELS = 16;
static const int CG = 4;            // cells per CTA (x) in k_fast
static const int CELL_MAX = 48;     // largest supported cell edge
static const int FAST_TS = 208;     // smem tile stride (bytes), >= CG*CELL_MAX+6+3 rounded
static const int FAST_TH = CELL_MAX + 6;
static const int DESC_WARPS = 4;

__constant__ int c_umax[16];
// the BRIEF pattern as k_describe reads it: word k of lane (= output byte) `lane` at [k * 32 + lane] (coalesced; the byte-per-test table above
// read per lane was an 8-way shared-memory bank conflict). Filled by cslam_frontend_create from the same table.
__device__ uint32_t g_patT[8 * 32];

struct LevelGeom {
    int w, h, pitch;          // image size, row pitch in bytes
    int minB, maxBX, maxBY;   // FAST region (EDGE_THRESHOLD-3 inset)
    int wCell, hCell, nColsEff, nRowsEff;
    int quota;                // mnFeaturesPerLevel
    int candCap;
    float scale;              // mvScaleFactor
    float sizeF;              // (int)(PATCH_SIZE*scale)
};

struct DevLevels {
    LevelGeom g[MAX_LEVELS];
    uint8_t* img[MAX_LEVELS];          // batch-major: frame f at img + f*pitch*h
    uint32_t* cand[MAX_LEVELS];        // batch x candCap packed (x | y<<12 | resp<<24), minBorder-relative
    uint16_t* pnode[MAX_LEVELS];       // batch x candCap scratch for k_distribute
    const uint16_t* xofs[MAX_LEVELS];  // resize tables for building level l from l-1
    const uint32_t* xab[MAX_LEVELS];   // a0 | a1<<16
    const uint16_t* yofs[MAX_LEVELS];
    const uint32_t* yab[MAX_LEVELS];
    const uint8_t* skip[MAX_LEVELS];   // per k_fast CTA: 1 = tile provably all-zero when level 0 came from k_warp (black corner tiles)
    int nlevels;
};

// ------------------------------------------------------------------------------------------------- k_warp
// Map entry: sx | sy<<16 with sx=cvRound(u*32), sy=cvRound(v*32) (cv::remap's INTER_BITS=5 fixed point).
// Weights are exact integers: (32-fy)(32-fx)*32 etc. (= cvRound of the fp32 table products * 32768).
// One thread = one canvas pixel of one face row, looped over FPT frames so a map entry is read once per FPT frames.
template <int FPT>
__global__ void __launch_bounds__(256) k_warp(const uint8_t* __restrict__ fisheye, int Iw, int Ih, const uint32_t* __restrict__ map, int W,
                                              uint8_t* __restrict__ canvas, int cpitch, size_t cframe, int batch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    const int face = blockIdx.z % 5, fg = blockIdx.z / 5;
    if (x >= W) return;
    const int tcol = (face == 1) ? 0 : (face == 2) ? 2 : 1;
    const int trow = (face == 3) ? 0 : (face == 4) ? 2 : 1;
    const uint32_t m = __ldg(map + ((size_t)face * W + row) * W + x);
    const int sx = m & 0xffff, sy = m >> 16;
    const int ix = sx >> 5, iy = sy >> 5, fx = sx & 31, fy = sy & 31;
    const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
    // ix,iy >= 0 always (maps are >= 0); each tap outside the source reads 0 (BORDER_CONSTANT)
    const bool x0ok = ix < Iw, x1ok = ix + 1 < Iw, y0ok = iy < Ih, y1ok = iy + 1 < Ih;
    const size_t o00 = (size_t)iy * Iw + ix;
    const size_t dst = (size_t)(trow * W + row) * cpitch + tcol * W + x;
#pragma unroll
    for (int k = 0; k < FPT; k++) {
        const int f = fg * FPT + k;
        if (f >= batch) break;
        const uint8_t* S = fisheye + (size_t)f * Iw * Ih;
        int v = 0;
        if (y0ok) {
            if (x0ok) v += __ldg(S + o00) * w00;
            if (x1ok) v += __ldg(S + o00 + 1) * w01;
        }
        if (y1ok) {
            if (x0ok) v += __ldg(S + o00 + Iw) * w10;
            if (x1ok) v += __ldg(S + o00 + Iw + 1) * w11;
        }
        canvas[(size_t)f * cframe + dst] = (uint8_t)((v + (1 << 14)) >> 15);
    }
}

// ------------------------------------------------------------------------------------------------- k_pyramid
// cv::resize(src,dst,sz,0,0,INTER_LINEAR) 8U model: 11-bit coefficients, horizontal int pass H = S[s]*a0 + S[s+1]*a1,
// vertical (((b0*(H0>>4))>>16)+((b1*(H1>>4))>>16)+2)>>2.  One thread = 4 consecutive dst pixels x PYR_RY dst rows:
// the 8-byte source window of the 4 pixels is fetched with aligned 32-bit loads, the taps are gathered with PRMT and the
// horizontal pass is two-tap IDP.2A (coefficient pair a0|a1<<16 straight from the table); a source row's H is reused by the
// next dst row when it needs it again (1.2 source rows per dst row instead of 2).
static const int PYR_RY = 8;
__global__ void __launch_bounds__(128) k_pyramid(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst, int dw, int dh,
                                                 int dpitch, const uint16_t* __restrict__ xofs, const uint32_t* __restrict__ xab,
                                                 const uint16_t* __restrict__ yofs, const uint32_t* __restrict__ yab) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x4 >= dw) return;
    const uint8_t* S = src + (size_t)blockIdx.z * spitch * sh;
    uint8_t* D = dst + (size_t)blockIdx.z * dpitch * dh;
    int s0[4]; uint32_t ab[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const int x = min(x4 + k, dw - 1); s0[k] = xofs[x]; ab[k] = xab[x]; }
    const int base = s0[0], al = base & ~3, sh8 = (base & 3) * 8;
    // byte selectors into the 8-byte window starting at `base`: pair (left,right) per pixel; right = left when clamped at the last column
    uint32_t sel01 = 0, sel23 = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t dl = (uint32_t)(s0[k] - base), dr = (s0[k] + 1 <= sw - 1) ? dl + 1 : dl;
        const uint32_t pr = dl | (dr << 4);
        if (k < 2) sel01 |= pr << (8 * k); else sel23 |= pr << (8 * (k - 2));
    }
    const bool has1 = al + 4 < spitch, has2 = al + 8 < spitch;
    auto hrow = [&](int sy, uint32_t (&H)[4]) {
        const uint32_t* R = reinterpret_cast<const uint32_t*>(S + (size_t)sy * spitch + al);
        const uint32_t w0 = __ldg(R), w1 = has1 ? __ldg(R + 1) : 0u, w2 = has2 ? __ldg(R + 2) : 0u;
        const uint32_t lo = __funnelshift_r(w0, w1, sh8), hi = __funnelshift_r(w1, w2, sh8);
        const uint32_t p01 = __byte_perm(lo, hi, sel01), p23 = __byte_perm(lo, hi, sel23);
        H[0] = __dp2a_lo(ab[0], p01, 0u) >> 4; H[1] = __dp2a_hi(ab[1], p01, 0u) >> 4;
        H[2] = __dp2a_lo(ab[2], p23, 0u) >> 4; H[3] = __dp2a_hi(ab[3], p23, 0u) >> 4;
    };
    const int y0 = blockIdx.y * PYR_RY, y1 = min(y0 + PYR_RY, dh);
    int prevIdx = -1;
    uint32_t prev[4] = {0, 0, 0, 0};
    for (int y = y0; y < y1; y++) {
        const int a = yofs[y], b = min(a + 1, sh - 1);
        const uint32_t bb = yab[y];
        const uint32_t b0 = bb & 0xffff, b1 = bb >> 16;
        uint32_t r0[4], r1[4];
        if (a == prevIdx) { r0[0] = prev[0]; r0[1] = prev[1]; r0[2] = prev[2]; r0[3] = prev[3]; } else hrow(a, r0);
        if (b == a) { r1[0] = r0[0]; r1[1] = r0[1]; r1[2] = r0[2]; r1[3] = r0[3]; } else hrow(b, r1);
        prevIdx = b; prev[0] = r1[0]; prev[1] = r1[1]; prev[2] = r1[2]; prev[3] = r1[3];
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) out |= ((((b0 * r0[k]) >> 16) + ((b1 * r1[k]) >> 16) + 2) >> 2) << (8 * k);   // <= 255, no clamp needed
        *reinterpret_cast<uint32_t*>(D + (size_t)y * dpitch + x4) = out;   // pitch is a multiple of 128: padding absorbs the tail
    }
}

// ------------------------------------------------------------------------------------------------- k_fast
// Arc score s(p) = max over the 16 arcs of 9 contiguous ring pixels of max(min(d), min(-d)), d = p - ring.
// corner(thr) <=> s > thr; OpenCV response = s-1. The reference runs cv::FAST per cell ROI [ini, ini+cell+6):
// detection areas of neighbouring cells tile the level without overlap, NMS only sees scores of the same cell,
// and the 20->7 fallback is decided per cell on "no keypoint survived NMS" (src/ORBExtractor.cpp:763-803).
// Exact arc score of 2 horizontally adjacent pixels at once in u16x2 lanes (sm_100a VIMNMX3.U16x2):
//   s = max(0, A - c, c - B),  A = max_k min(R[k..k+8]),  B = min_k max(R[k..k+8])   (R = ring intensities, c = centre)
// min/max over 9 contiguous ring pixels = min3 of three min3's. No data-dependent branch, no per-pixel early-out:
// on textured input almost every pixel passes the cheap antipodal tests, so the branch-free full score is cheaper
// than queueing (measured with ncu: profiles/r01_k_fast_*.txt).
__device__ __forceinline__ uint32_t arc_score_x2(const uint32_t (&R)[16], uint32_t c) {
    uint32_t mn[16], mx[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn[k] = __vimin3_u16x2(R[k], R[(k + 1) & 15], R[(k + 2) & 15]);
        mx[k] = __vimax3_u16x2(R[k], R[(k + 1) & 15], R[(k + 2) & 15]);
    }
    uint32_t A = 0u, Bm = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t a0 = __vimin3_u16x2(mn[k], mn[(k + 3) & 15], mn[(k + 6) & 15]);
        const uint32_t a1 = __vimin3_u16x2(mn[k + 1], mn[(k + 4) & 15], mn[(k + 7) & 15]);
        const uint32_t b0 = __vimax3_u16x2(mx[k], mx[(k + 3) & 15], mx[(k + 6) & 15]);
        const uint32_t b1 = __vimax3_u16x2(mx[k + 1], mx[(k + 4) & 15], mx[(k + 7) & 15]);
        A = __vimax3_u16x2(A, a0, a1);
        Bm = __vimin3_u16x2(Bm, b0, b1);
    }
    const uint32_t t1 = __vmaxu2(A, c) - c;      // lanes never borrow: max(A,c) >= c
    const uint32_t t2 = c - __vminu2(Bm, c);
    return __vmaxu2(t1, t2);
}

// q = i / d for 0 <= i < 65536/..: d-uniform fast division (tile loops have i < 8192, d <= 64)
__device__ __forceinline__ int fdiv_small(int i, uint32_t magic) { return (int)(((uint32_t)i * magic) >> 20); }
__device__ __forceinline__ uint32_t fdiv_magic(int d) { return ((1u << 20) + d - 1) / d; }

// Layout of one CTA: CG cells of one cell row. Phases:
//  A  branch-free exact arc score of every detection pixel (4 pixels per work item) -> S (u8, 0 where s <= minTh)
//  B  3x3 NMS at minTh inside each cell -> shared list; a keypoint at iniTh is exactly an NMS survivor with s > iniTh
//     (a suppressor needs s_n >= s), so the per-cell 20->7 fallback is a filter on that list
//  C  compaction of the selected entries into the global candidate list (one global atomic per 256 entries)
static const int FAST_KMAX = CG * (CELL_MAX / 2) * (CELL_MAX / 2);   // NMS survivors are pairwise non-adjacent
static const int FAST_SW = FAST_TS / 2;                              // pair-words per score row

__global__ void __launch_bounds__(256) k_fast(const uint8_t* __restrict__ img, LevelGeom g, int iniTh, int minTh, uint32_t* __restrict__ cand,
                                              uint32_t* __restrict__ candCount, int countStride, int* __restrict__ errFlag,
                                              const uint8_t* __restrict__ skip) {
    __shared__ __align__(16) uint8_t tile[FAST_TH * FAST_TS];
    // score map, one u16 per pixel as u16x2 pair-words: word 2*wi (+1) of row y+1 holds tile columns 4*wi..4*wi+1 (+2..+3)
    __shared__ __align__(16) uint32_t S2[(CELL_MAX + 2) * FAST_SW];
    __shared__ uint32_t klist[FAST_KMAX];                            // x | y<<8 | s<<16 | cell<<24
    __shared__ uint8_t cellInfo[FAST_TS];   // per detection column: cell index | 0x40 first column of its cell | 0x80 last
    __shared__ uint32_t mskV[FAST_SW], mskL[FAST_SW], mskR[FAST_SW];  // per pair-word lane masks: valid column / has left / has right neighbour in its cell
    __shared__ int cnt20[CG];
    __shared__ int anyNonZero, nK, chunkBase;
    __shared__ int warpCnt[8];
    const int frame = blockIdx.z, cellRow = blockIdx.y, j0 = blockIdx.x * CG;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (skip && skip[cellRow * gridDim.x + blockIdx.x]) return;   // tile provably all-zero (black cubemap corner)
    const uint8_t* I = img + (size_t)frame * g.pitch * g.h;
    const int iniY = g.minB + cellRow * g.hCell, maxY = min(iniY + g.hCell + 6, g.maxBY);
    const int ncell = min(CG, g.nColsEff - j0);
    const int tx0 = g.minB + j0 * g.wCell, tx1 = min(tx0 + ncell * g.wCell + 6, g.maxBX);
    const int a0 = tx0 & ~3, nwords = (tx1 - a0 + 3) >> 2, nrows = maxY - iniY;
    const int nx = tx1 - tx0 - 6, ny = nrows - 6;   // detection area
    const int xoff = tx0 - a0 + 3;                   // tile column of detection x = 0
    const int w0 = xoff >> 2, w1 = (xoff + nx + 3) >> 2, nw = w1 - w0;   // S/tile words covering the detection columns
    if (nx <= 0 || ny <= 0) return;   // block-uniform: a ROI with fewer than 7 rows / columns has no detection area (cv::FAST returns nothing)
    if (tid < CG) cnt20[tid] = 0;
    if (tid == 0) { anyNonZero = 0; nK = 0; }
    __syncthreads();
    uint32_t acc = 0;
    {
        const uint32_t mg = fdiv_magic(nwords);
        const uint8_t* base = I + (size_t)iniY * g.pitch + a0;
        for (int i = tid; i < nwords * nrows; i += 256) {
            const int r = fdiv_small(i, mg), wd = i - r * nwords;
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r * g.pitch) + wd);
            reinterpret_cast<uint32_t*>(tile + r * FAST_TS)[wd] = v;
            acc |= v;
        }
    }
    if (acc) anyNonZero = 1;
    // S2 must read 0 around the scored area: rows 0 and ny+1, and the pair-words left/right of [2*w0, 2*w1) in every row
    for (int i = tid; i < 2 * FAST_SW; i += 256) S2[(i >= FAST_SW ? (ny + 1) * FAST_SW : 0) + (i % FAST_SW)] = 0;
    for (int i = tid; i < 2 * ny; i += 256) {
        const int r = (i >> 1) + 1, psel = (i & 1) ? 2 * w1 : 2 * w0 - 1;
        if (psel >= 0 && psel < FAST_SW) S2[r * FAST_SW + psel] = 0;
    }
    for (int pw = tid; pw < FAST_SW; pw += 256) {
        uint32_t v = 0, l = 0, r = 0;
        for (int h = 0; h < 2; h++) {
            const int x = 2 * pw + h - xoff;   // detection x of this lane
            if (x >= 0 && x < nx) {
                const int xl = x % g.wCell;
                v |= 0xffffu << (16 * h);
                if (xl != 0) l |= 0xffffu << (16 * h);
                if (xl != g.wCell - 1 && x != nx - 1) r |= 0xffffu << (16 * h);
            }
        }
        mskV[pw] = v; mskL[pw] = l; mskR[pw] = r;
    }
    for (int x = tid; x < nx; x += 256) {
        const int c = x / g.wCell, xl = x - c * g.wCell;
        cellInfo[x] = (uint8_t)(c | (xl == 0 ? 0x40 : 0) | ((xl == g.wCell - 1 || x == nx - 1) ? 0x80 : 0));
    }
    __syncthreads();
    if (!anyNonZero) return;   // an all-zero tile has no corners at any threshold
    // ---- A
    const uint32_t mgw = fdiv_magic(nw);
    {
        const int nitems = nw * ny;
        for (int it = tid; it < nitems; it += 256) {
            const int y = fdiv_small(it, mgw), wi = w0 + (it - y * nw);
            const uint32_t* r0 = reinterpret_cast<const uint32_t*>(tile + y * FAST_TS) + wi;   // row y-3 relative to the centre row
            const int RS = FAST_TS / 4;
            uint32_t V[16];
            {   // dy=-3: ring 15 (dx-1), 0 (dx 0), 1 (dx+1)
                const uint32_t m = wi > 0 ? r0[-1] : 0u, c = r0[0], p = r0[1];
                V[15] = __funnelshift_r(m, c, 24); V[0] = c; V[1] = __funnelshift_r(c, p, 8);
            }
            {   // dy=-2: ring 14 (dx-2), 2 (dx+2)
                const uint32_t* r = r0 + RS; const uint32_t m = wi > 0 ? r[-1] : 0u, c = r[0], p = r[1];
                V[14] = __funnelshift_r(m, c, 16); V[2] = __funnelshift_r(c, p, 16);
            }
            {   // dy=-1: ring 13 (dx-3), 3 (dx+3)
                const uint32_t* r = r0 + 2 * RS; const uint32_t m = wi > 0 ? r[-1] : 0u, c = r[0], p = r[1];
                V[13] = __funnelshift_r(m, c, 8); V[3] = __funnelshift_r(c, p, 24);
            }
            uint32_t C;
            {   // dy=0: ring 12 (dx-3), 4 (dx+3), centre
                const uint32_t* r = r0 + 3 * RS; const uint32_t m = wi > 0 ? r[-1] : 0u, c = r[0], p = r[1];
                V[12] = __funnelshift_r(m, c, 8); V[4] = __funnelshift_r(c, p, 24); C = c;
            }
            {   // dy=+1: ring 11 (dx-3), 5 (dx+3)
                const uint32_t* r = r0 + 4 * RS; const uint32_t m = wi > 0 ? r[-1] : 0u, c = r[0], p = r[1];
                V[11] = __funnelshift_r(m, c, 8); V[5] = __funnelshift_r(c, p, 24);
            }
            {   // dy=+2: ring 10 (dx-2), 6 (dx+2)
                const uint32_t* r = r0 + 5 * RS; const uint32_t m = wi > 0 ? r[-1] : 0u, c = r[0], p = r[1];
                V[10] = __funnelshift_r(m, c, 16); V[6] = __funnelshift_r(c, p, 16);
            }
            {   // dy=+3: ring 9 (dx-1), 8 (dx 0), 7 (dx+1)
                const uint32_t* r = r0 + 6 * RS; const uint32_t m = wi > 0 ? r[-1] : 0u, c = r[0], p = r[1];
                V[9] = __funnelshift_r(m, c, 24); V[8] = c; V[7] = __funnelshift_r(c, p, 8);
            }
            uint32_t R[16];
#pragma unroll
            for (int k = 0; k < 16; k++) R[k] = __byte_perm(V[k], 0u, 0x4140);
            const uint32_t sA = arc_score_x2(R, __byte_perm(C, 0u, 0x4140));
#pragma unroll
            for (int k = 0; k < 16; k++) R[k] = __byte_perm(V[k], 0u, 0x4342);
            const uint32_t sB = arc_score_x2(R, __byte_perm(C, 0u, 0x4342));
            // scores <= minTh and columns outside the detection area read as 0
            const uint32_t thr = (uint32_t)minTh;
            uint32_t oA = ((sA & 0xffffu) > thr ? (sA & 0xffffu) : 0u) | ((sA >> 16) > thr ? (sA & 0xffff0000u) : 0u);
            uint32_t oB = ((sB & 0xffffu) > thr ? (sB & 0xffffu) : 0u) | ((sB >> 16) > thr ? (sB & 0xffff0000u) : 0u);
            uint32_t* Srow = S2 + (y + 1) * FAST_SW + 2 * wi;
            Srow[0] = oA & mskV[2 * wi]; Srow[1] = oB & mskV[2 * wi + 1];
        }
    }
    __syncthreads();
    // ---- B: NMS at minTh in u16x2 lanes: keep s iff every same-cell neighbour is < s (all stored scores are > minTh)
    {
        const int np = 2 * nw, nitems = np * ny;
        const uint32_t mgp = fdiv_magic(np);
        for (int it = tid; it < nitems; it += 256) {
            const int y = fdiv_small(it, mgp), pw = 2 * w0 + (it - y * np);
            const uint32_t* r1 = S2 + (y + 1) * FAST_SW + pw;
            const uint32_t cur = r1[0];
            if (cur == 0) continue;
            const uint32_t* r0 = r1 - FAST_SW; const uint32_t* r2 = r1 + FAST_SW;
            const uint32_t ml = mskL[pw], mr = mskR[pw];
            const uint32_t a0 = pw > 0 ? r0[-1] : 0u, b0 = r0[0], c0 = r0[1];
            const uint32_t a1 = pw > 0 ? r1[-1] : 0u, c1 = r1[1];
            const uint32_t a2 = pw > 0 ? r2[-1] : 0u, b2 = r2[0], c2 = r2[1];
            const uint32_t upL = __funnelshift_r(a0, b0, 16) & ml, upR = __funnelshift_r(b0, c0, 16) & mr;
            const uint32_t cuL = __funnelshift_r(a1, cur, 16) & ml, cuR = __funnelshift_r(cur, c1, 16) & mr;
            const uint32_t dnL = __funnelshift_r(a2, b2, 16) & ml, dnR = __funnelshift_r(b2, c2, 16) & mr;
            const uint32_t M = __vimax3_u16x2(__vimax3_u16x2(upL, b0, upR), __vimax3_u16x2(cuL, cuR, b2), __vmaxu2(dnL, dnR));
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int sv = (cur >> (16 * h)) & 0xffff, mv = (M >> (16 * h)) & 0xffff;
                if (sv > mv) {
                    const int x = 2 * pw + h - xoff, cell = cellInfo[x] & 0x3f;
                    const int pos = atomicAdd(&nK, 1);
                    if (pos >= FAST_KMAX) *errFlag = CSLAM_E_CAPACITY;
                    else klist[pos] = (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)sv << 16) | ((uint32_t)cell << 24);
                    if (sv > iniTh) atomicAdd(&cnt20[cell], 1);
                }
            }
        }
    }
    __syncthreads();
    // ---- C
    uint32_t* out = cand + (size_t)frame * g.candCap;
    uint32_t* cc = candCount + (size_t)frame * countStride;
    const int nk = min(nK, FAST_KMAX);
    for (int base = 0; base < nk; base += 256) {
        const int i = base + tid;
        uint32_t e = 0; bool sel = false;
        if (i < nk) {
            e = klist[i];
            const int sv = (e >> 16) & 0xff, cell = e >> 24;
            sel = cnt20[cell] > 0 ? sv > iniTh : true;
        }
        const unsigned ball = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) warpCnt[warp] = __popc(ball);
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 8; w++) { const int t = warpCnt[w]; warpCnt[w] = tot; tot += t; }
            chunkBase = tot ? (int)atomicAdd(cc, (uint32_t)tot) : 0;
        }
        __syncthreads();
        if (sel) {
            const uint32_t pos = (uint32_t)chunkBase + warpCnt[warp] + __popc(ball & ((1u << lane) - 1));
            if (pos < (uint32_t)g.candCap) {
                const uint32_t X = tx0 + 3 + (e & 0xff) - g.minB, Y = iniY + 3 + ((e >> 8) & 0xff) - g.minB;
                out[pos] = X | (Y << 12) | ((((e >> 16) & 0xff) - 1) << 24);
            } else {
                *errFlag = CSLAM_E_CAPACITY;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------- k_distribute
// DistributeOctTree (src/ORBExtractor.cpp:511-737) without lists or pointers. Facts used:
//  * children bounds depend only on the parent bounds; a key's child is decided by x<midX / y<midY;
//  * std::list::push_front everywhere => list order == descending creation sequence, so "list order" is a sort key;
//  * the order of keys inside a node only matters for the final "max response, first wins" pick, and the first key
//    of equal response is the one earliest in (cell row, cell col, y, x) order, recoverable from coordinates.
// One CTA per (level, frame). Nodes live in shared memory, keys stay in global/L2 with a per-key node index.
struct QNode { uint16_t ulx, uly, brx, bry; uint32_t cnt; uint32_t seq : 31; uint32_t flagE : 1; };

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t* data, int n, uint32_t* warpSums) {
    // in-place exclusive scan of data[0..n) by the whole block; returns the total. n <= blockDim.x * per
    const int T = blockDim.x, tid = threadIdx.x;
    const int per = (n + T - 1) / T;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    uint32_t s = 0;
    for (int i = lo; i < hi; i++) s += data[i];
    // block scan of s
    const int lane = tid & 31, wid = tid >> 5;
    uint32_t v = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warpSums[wid] = v;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < (T >> 5) ? warpSums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        warpSums[lane] = w;
    }
    __syncthreads();
    uint32_t excl = v - s + (wid ? warpSums[wid - 1] : 0);
    const uint32_t total = warpSums[(T >> 5) - 1];
    for (int i = lo; i < hi; i++) { uint32_t t = data[i]; data[i] = excl; excl += t; }
    __syncthreads();
    return total;
}

struct DistributeArgs {
    DevLevels L;
    const uint32_t* candCount;   // batch x nlevels
    uint32_t* kept;              // batch x nlevels x keptCap : x | y<<12 | resp<<24 in LEVEL coordinates
    uint32_t* keptCount;         // batch x nlevels
    const uint8_t* mask; int maskPitch; int imgW, imgH, faceW, faceH;
    int keptCap, M;              // M: node capacity
    int* errFlag;
};

__global__ void __launch_bounds__(256) k_distribute(DistributeArgs A) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int level = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, T = blockDim.x;
    const LevelGeom g = A.L.g[level];
    const int M = A.M, N = g.quota;
    QNode* listA = reinterpret_cast<QNode*>(smem_raw);
    QNode* listB = listA + M;
    uint32_t* childCnt = reinterpret_cast<uint32_t*>(listB + M);   // M*4 (count, later new position)
    uint32_t* scanA = childCnt + 4 * M;                            // M
    uint32_t* scanB = scanA + M;                                   // M
    uint16_t* ordOf = reinterpret_cast<uint16_t*>(scanB + M);      // M  processing order of a split node
    uint16_t* remap = ordOf + M;                                   // M
    uint8_t* splitF = reinterpret_cast<uint8_t*>(remap + M);       // M
    __shared__ uint32_t warpSums[32];
    __shared__ int sh_nToExpand, sh_J;
    __shared__ uint32_t sh_base;

    const uint32_t nAll = A.candCount[(size_t)frame * A.L.nlevels + level];
    const int n = (int)min(nAll, (uint32_t)g.candCap);
    const uint32_t* pts = A.L.cand[level] + (size_t)frame * g.candCap;
    uint16_t* pnode = A.L.pnode[level] + (size_t)frame * g.candCap;
    uint32_t* keptOut = A.kept + ((size_t)frame * A.L.nlevels + level) * A.keptCap;
    uint32_t* keptCnt = A.keptCount + (size_t)frame * A.L.nlevels + level;
    if (n == 0) { if (tid == 0) *keptCnt = 0; return; }

    QNode* cur = listA; QNode* nxt = listB;
    if (tid == 0) {
        QNode r; r.ulx = 0; r.uly = 0; r.brx = (uint16_t)(g.maxBX - g.minB); r.bry = (uint16_t)(g.maxBY - g.minB);
        r.cnt = (uint32_t)n; r.flagE = 0; r.seq = 0;
        cur[0] = r;
    }
    for (int p = tid; p < n; p += T) pnode[p] = 0;
    __syncthreads();
    int nList = 1; uint32_t seqNext = 1; bool finishMode = false;

    for (int round = 0; round < 64; round++) {
        const int prevSize = nList;
        // ---- which nodes are candidates to split, and in which order
        int nCandSplit = 0;
        if (!finishMode) {
            for (int i = tid; i < nList; i += T) scanA[i] = cur[i].cnt > 1 ? 1u : 0u;
            __syncthreads();
            for (int i = tid; i < nList; i += T) splitF[i] = (uint8_t)scanA[i];
            __syncthreads();
            nCandSplit = (int)block_exclusive_scan(scanA, nList, warpSums);
            for (int i = tid; i < nList; i += T) ordOf[i] = (uint16_t)scanA[i];
        } else {
            // rank expandable nodes by descending (count, seq): the reference sorts ascending and walks from the back
            for (int i = tid; i < nList; i += T) {
                int r = 0; const bool e = cur[i].flagE != 0;
                if (e) {
                    const uint32_t ci = cur[i].cnt, si = cur[i].seq;
                    for (int k = 0; k < nList; k++) {
                        if (!cur[k].flagE) continue;
                        const uint32_t ck = cur[k].cnt, sk = cur[k].seq;
                        r += (ck > ci) || (ck == ci && sk > si);
                    }
                }
                splitF[i] = e; ordOf[i] = (uint16_t)r; scanA[i] = e;
            }
            __syncthreads();
            nCandSplit = (int)block_exclusive_scan(scanA, nList, warpSums);
        }
        if (nCandSplit == 0) break;   // nothing can be divided (the reference would spin forever here)
        // ---- children populations
        for (int i = tid; i < 4 * nList; i += T) childCnt[i] = 0;
        __syncthreads();
        for (int p = tid; p < n; p += T) {
            const int i = pnode[p];
            if (splitF[i]) {
                const QNode nd = cur[i];
                const uint32_t v = pts[p];
                const int x = v & 0xfff, y = (v >> 12) & 0xfff;
                const int midX = nd.ulx + ((nd.brx - nd.ulx + 1) >> 1), midY = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
                atomicAdd(&childCnt[4 * i + (x < midX ? 0 : 1) + (y < midY ? 0 : 2)], 1u);
            }
        }
        __syncthreads();
        // ---- finishing phase: stop after the split that makes the list reach N
        int nSplit = nCandSplit;
        if (finishMode) {
            for (int i = tid; i < nCandSplit; i += T) scanB[i] = 0;
            __syncthreads();
            for (int i = tid; i < nList; i += T) if (splitF[i]) {
                int k = 0;
                for (int q = 0; q < 4; q++) k += childCnt[4 * i + q] > 0;
                scanB[ordOf[i]] = (uint32_t)(k - 1);
            }
            if (tid == 0) sh_J = nCandSplit;
            __syncthreads();
            block_exclusive_scan(scanB, nCandSplit, warpSums);
            // scanB[o] = growth before processing o; size after o = prevSize + scanB[o] + gain(o) = prevSize + scanB[o+1]
            for (int o = tid; o < nCandSplit; o += T) {
                const uint32_t before = prevSize + scanB[o];
                // size after processing o >= N  <=>  the first o whose "before" is < N but after >= N
                // after(o) = before(o+1) for o+1 < nCandSplit; handle the last one by recomputing its gain below
                if (before >= (uint32_t)N) atomicMin(&sh_J, o);   // o-1 was the breaking split => splits = o
            }
            __syncthreads();
            nSplit = sh_J;   // number of nodes actually divided (those with ord < nSplit)
            __syncthreads();
            for (int i = tid; i < nList; i += T) if (splitF[i] && ordOf[i] >= nSplit) splitF[i] = 0;
            __syncthreads();
        }
        // ---- creation ranks of children (reference creation order = processing order, n1..n4)
        for (int i = tid; i < nSplit; i += T) scanB[i] = 0;
        __syncthreads();
        for (int i = tid; i < nList; i += T) {
            if (splitF[i]) {
                int k = 0;
                for (int q = 0; q < 4; q++) k += childCnt[4 * i + q] > 0;
                scanB[ordOf[i]] = (uint32_t)k;
            }
            scanA[i] = splitF[i] ? 0u : 1u;
        }
        __syncthreads();
        const int C = (int)block_exclusive_scan(scanB, nSplit, warpSums);
        const int nUnsplit = (int)block_exclusive_scan(scanA, nList, warpSums);
        const int newSize = C + nUnsplit;
        if (newSize > M) { if (tid == 0) *A.errFlag = CSLAM_E_CAPACITY; break; }
        if (tid == 0) sh_nToExpand = 0;
        __syncthreads();
        int myExp = 0;
        for (int i = tid; i < nList; i += T) {
            const QNode nd = cur[i];
            if (splitF[i]) {
                const int midX = nd.ulx + ((nd.brx - nd.ulx + 1) >> 1), midY = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
                int r = (int)scanB[ordOf[i]];
                for (int q = 0; q < 4; q++) {
                    const uint32_t c = childCnt[4 * i + q];
                    if (c == 0) continue;
                    QNode ch;
                    ch.ulx = (q & 1) ? midX : nd.ulx; ch.brx = (q & 1) ? nd.brx : midX;
                    ch.uly = (q & 2) ? midY : nd.uly; ch.bry = (q & 2) ? nd.bry : midY;
                    ch.cnt = c; ch.flagE = c > 1; ch.seq = seqNext + r;
                    const int pos = C - 1 - r;
                    nxt[pos] = ch;
                    childCnt[4 * i + q] = (uint32_t)pos;
                    myExp += c > 1;
                    r++;
                }
            } else {
                const int pos = C + (int)scanA[i];
                QNode keep = nd; keep.flagE = 0;
                nxt[pos] = keep;
                remap[i] = (uint16_t)pos;
            }
        }
        if (myExp) atomicAdd(&sh_nToExpand, myExp);
        __syncthreads();
        for (int p = tid; p < n; p += T) {
            const int i = pnode[p];
            if (splitF[i]) {
                const QNode nd = cur[i];
                const uint32_t v = pts[p];
                const int x = v & 0xfff, y = (v >> 12) & 0xfff;
                const int midX = nd.ulx + ((nd.brx - nd.ulx + 1) >> 1), midY = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
                pnode[p] = (uint16_t)childCnt[4 * i + (x < midX ? 0 : 1) + (y < midY ? 0 : 2)];
            } else {
                pnode[p] = remap[i];
            }
        }
        __syncthreads();
        const int nToExpand = sh_nToExpand;
        { QNode* t = cur; cur = nxt; nxt = t; }
        nList = newSize; seqNext += C;
        if (!finishMode) {
            if (nList >= N || (nList == prevSize && nList >= N / 100)) break;
            if (nList + 3 * nToExpand > N) finishMode = true;
        } else {
            if (nList >= N || nList == prevSize) break;
        }
        __syncthreads();
    }
    __syncthreads();
    // ---- best key per node: max response, ties -> earliest in (cell row, cell col, y, x) order
    // `best` aliases the inactive list buffer (M*8 bytes <= M*16).
    unsigned long long* best = reinterpret_cast<unsigned long long*>(nxt);
    for (int i = tid; i < nList; i += T) best[i] = 0ull;
    __syncthreads();
    for (int p = tid; p < n; p += T) {
        const uint32_t v = pts[p];
        const uint32_t x = v & 0xfff, y = (v >> 12) & 0xfff, r = v >> 24;
        const uint32_t cy = (y - 3) / g.hCell, cx = (x - 3) / g.wCell;
        const unsigned long long key = ((unsigned long long)cy << 32) | ((unsigned long long)cx << 24) | ((unsigned long long)y << 12) | x;
        const unsigned long long packed = ((unsigned long long)(r + 1) << 40) | (0xffffffffffull - key);
        atomicMax(&best[pnode[p]], packed);
    }
    __syncthreads();
    // ---- cull (src/ORBExtractor.cpp:883-904) in list order, ordered compaction
    for (int base = 0; base < nList; base += T) {
        const int i = base + tid;
        bool keep = false; uint32_t packedOut = 0;
        if (i < nList) {
            const unsigned long long b = best[i];
            const uint32_t key = (uint32_t)(0xffffffffffull - (b & 0xffffffffffull));
            const uint32_t x = key & 0xfff, y = (key >> 12) & 0xfff, r = (uint32_t)(b >> 40) - 1;
            const uint32_t lx = x + g.minB, ly = y + g.minB;
            const float ptx = __fmul_rn((float)lx, g.scale), pty = __fmul_rn((float)ly, g.scale);
            const float fi = __fdiv_rn(ptx, (float)A.faceW), fj = __fdiv_rn(pty, (float)A.faceH);
            const bool face = (fi >= 0 && fi < 1 && fj >= 1 && fj < 2) || (fi >= 1 && fi < 2 && fj >= 0 && fj < 3) || (fi >= 2 && fi < 3 && fj >= 1 && fj < 2);
            const int mxI = (int)__fadd_rn(ptx, 0.5f), myI = (int)__fadd_rn(pty, 0.5f);
            keep = face && !(ptx < 0 || mxI >= A.imgW || pty < 0 || myI >= A.imgH);
            if (keep) keep = A.mask[(size_t)myI * A.maskPitch + mxI] != 0;
            packedOut = lx | (ly << 12) | (r << 24);
        }
        scanA[tid] = keep ? 1u : 0u;
        __syncthreads();
        const uint32_t tot = block_exclusive_scan(scanA, T, warpSums);
        if (tid == 0 && base == 0) sh_base = 0;
        __syncthreads();
        if (keep) {
            const uint32_t pos = sh_base + scanA[tid];
            if (pos < (uint32_t)A.keptCap) keptOut[pos] = packedOut; else *A.errFlag = CSLAM_E_CAPACITY;
        }
        __syncthreads();
        if (tid == 0) sh_base += tot;
        __syncthreads();
        if (base + T >= nList && tid == 0) *keptCnt = min(sh_base, (uint32_t)A.keptCap);
    }
}

// ------------------------------------------------------------------------------------------------- k_describe
struct DescribeArgs {
    DevLevels L;
    const uint32_t* kept; const uint32_t* keptCount; int keptCap;
    cslam_keypoint* kps; uint8_t* desc; int32_t* nOut; int kpCap;
    int* errFlag;
};

static const int DP_STRIDE = 52;                 // patch row stride (bytes): 43 px + up to 3 bytes of word misalignment; 13 words: consecutive rows fall in distinct banks
static const int DH_STRIDE = 46;                 // transposed H: u16 per row index, per column (23 words: the 4-column stride of the transposed stores, 92 words, spreads over 8 banks)
static const int BL_STRIDE = 44;                 // blurred patch row stride in bytes (11 words: 4-row stride 44 words -> 8 banks; 40 gave 4)
__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe(DescribeArgs A) {
    __shared__ __align__(16) uint8_t s_patch[DESC_WARPS][43 * DP_STRIDE + 16];
    __shared__ __align__(16) uint16_t s_ht[DESC_WARPS][37 * DH_STRIDE + 8];
    const int level = blockIdx.y, frame = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t* kc = A.keptCount + (size_t)frame * A.L.nlevels;
    const int slot = blockIdx.x * DESC_WARPS + warp;
    int offset = 0, total = 0;
    for (int l = 0; l < A.L.nlevels; l++) { const int c = (int)kc[l]; if (l < level) offset += c; total += c; }
    if (blockIdx.x == 0 && level == 0 && threadIdx.x == 0) A.nOut[frame] = min(total, A.kpCap);
    if (slot >= (int)kc[level]) return;
    const int outIdx = offset + slot;
    if (outIdx >= A.kpCap) { if (lane == 0) *A.errFlag = CSLAM_E_CAPACITY; return; }
    const LevelGeom g = A.L.g[level];
    const uint8_t* I = A.L.img[level] + (size_t)frame * g.pitch * g.h;
    const uint32_t v = A.kept[((size_t)frame * A.L.nlevels + level) * A.keptCap + slot];
    const int kx = v & 0xfff, ky = (v >> 12) & 0xfff, resp = v >> 24;

    // ---- 43x43 neighbourhood -> P[r][po + c] (REFLECT_101 at the level edges, like the blur of the apron-less clone)
    uint8_t* P = s_patch[warp];
    int po;
    if (kx >= 21 && ky >= 21 && kx + 21 < g.w && ky + 21 < g.h) {
        // interior: aligned 32-bit loads, two rows (16 lanes each, 12 words used) per step
        const int x0 = kx - 21; po = x0 & 3;
        const int wd = lane & 15, rs = lane >> 4;
        const uint8_t* base = I + (size_t)(ky - 21) * g.pitch + (x0 & ~3);
        for (int r = rs; r < 43; r += 2)
            if (wd < 12) reinterpret_cast<uint32_t*>(P + r * DP_STRIDE)[wd] = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r * g.pitch) + wd);
    } else {
        po = 0;
        for (int i = lane; i < 43 * 43; i += 32) {
            const int r = i / 43, c = i - r * 43;
            const int yy = reflect101(ky - 21 + r, g.h), xx = reflect101(kx - 21 + c, g.w);
            P[r * DP_STRIDE + c] = __ldg(I + (size_t)yy * g.pitch + xx);
        }
    }
    __syncwarp();
    // ---- IC_Angle on the unblurred patch: lane = column u in [-15,15]
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - HALF_PATCH;
        int colsum = 0;
        for (int vv = -HALF_PATCH; vv <= HALF_PATCH; vv++) {
            const int av = vv < 0 ? -vv : vv;
            if ((u < 0 ? -u : u) <= c_umax[av]) {
                const int val = P[(21 + vv) * DP_STRIDE + po + 21 + u];
                colsum += val; m01 += vv * val;
            }
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    // ---- 7x7 sigma-2 Gaussian, OpenCV 8.8 fixed point [18,34,48,56,48,34,18], only where BRIEF can sample (+-18).
    // Horizontal: 4 outputs per item from one 12-byte window with IDP.4A; stored transposed (Ht[c][r], u16, exact: <= 65280).
    uint16_t* Ht = s_ht[warp];
    {
        const uint32_t K0123 = 18u | (34u << 8) | (48u << 16) | (56u << 24), K456 = 48u | (34u << 8) | (18u << 16);
        for (int it = lane; it < 43 * 10; it += 32) {   // lanes = consecutive rows of one column group: loads 13 words apart, transposed stores contiguous
            const int cg = it / 43, r = it - cg * 43;
            const int bidx = po + 4 * cg;
            const uint32_t* W = reinterpret_cast<const uint32_t*>(P + r * DP_STRIDE) + (bidx >> 2);
            const int sh = (bidx & 3) * 8;
            const uint32_t w0 = W[0], w1 = W[1], w2 = W[2], w3 = W[3];
            const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh), x2 = __funnelshift_r(w2, w3, sh);
            uint32_t h[4];
            h[0] = __dp4a(x0, K0123, __dp4a(x1, K456, 0u));
            h[1] = __dp4a(__funnelshift_r(x0, x1, 8), K0123, __dp4a(__funnelshift_r(x1, x2, 8), K456, 0u));
            h[2] = __dp4a(__funnelshift_r(x0, x1, 16), K0123, __dp4a(__funnelshift_r(x1, x2, 16), K456, 0u));
            h[3] = __dp4a(__funnelshift_r(x0, x1, 24), K0123, __dp4a(__funnelshift_r(x1, x2, 24), K456, 0u));
#pragma unroll
            for (int j = 0; j < 4; j++) if (4 * cg + j < 37) Ht[(4 * cg + j) * DH_STRIDE + r] = (uint16_t)h[j];
        }
    }
    __syncwarp();
    // Vertical: 4 output rows of one column per item from five u16x2 words with IDP.2A; dst = (V + 32768) >> 16
    uint8_t* Bl = s_patch[warp];   // the raw patch is dead after the horizontal pass
    {
        const uint32_t K01 = 18u | (34u << 8), K23 = 48u | (56u << 8), K45 = 48u | (34u << 8);
        for (int it = lane; it < 37 * 10; it += 32) {   // lanes = consecutive columns of one row group: loads 23 words apart, byte stores contiguous
            const int rg = it / 37, c = it - rg * 37;
            const uint32_t* Hw = reinterpret_cast<const uint32_t*>(Ht + c * DH_STRIDE) + 2 * rg;
            const uint32_t q0 = Hw[0], q1 = Hw[1], q2 = Hw[2], q3 = Hw[3], q4 = Hw[4];
            const uint32_t s01 = __funnelshift_r(q0, q1, 16), s12 = __funnelshift_r(q1, q2, 16), s23 = __funnelshift_r(q2, q3, 16), s34 = __funnelshift_r(q3, q4, 16);
            uint32_t o[4];
            o[0] = __dp2a_lo(q0, K01, __dp2a_lo(q1, K23, __dp2a_lo(q2, K45, 18u * (q3 & 0xffffu))));
            o[1] = __dp2a_lo(s01, K01, __dp2a_lo(s12, K23, __dp2a_lo(s23, K45, 18u * (q3 >> 16))));
            o[2] = __dp2a_lo(q1, K01, __dp2a_lo(q2, K23, __dp2a_lo(q3, K45, 18u * (q4 & 0xffffu))));
            o[3] = __dp2a_lo(s12, K01, __dp2a_lo(s23, K23, __dp2a_lo(s34, K45, 18u * (q4 >> 16))));
#pragma unroll
            for (int j = 0; j < 4; j++) if (4 * rg + j < 37) Bl[(4 * rg + j) * BL_STRIDE + c] = (uint8_t)((o[j] + 32768u) >> 16);
        }
    }
    __syncwarp();
    // ---- steered BRIEF: lane = output byte
    const float factorPI = (float)(3.14159265358979323846 / 180.0);
    float a, b;
    det_sincosf(__fmul_rn(angle, factorPI), &b, &a);
    uint32_t val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t pw = __ldg(g_patT + k * 32 + lane);
        const float x0 = (float)(int)(signed char)(pw & 0xff), y0 = (float)(int)(signed char)((pw >> 8) & 0xff), x1 = (float)(int)(signed char)((pw >> 16) & 0xff),
                    y1 = (float)(int)(signed char)(pw >> 24);
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a))), c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a))), c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = Bl[(r0 + 18) * BL_STRIDE + c0 + 18], t1 = Bl[(r1 + 18) * BL_STRIDE + c1 + 18];
        val |= (uint32_t)(t0 < t1) << k;
    }
    A.desc[((size_t)frame * A.kpCap + outIdx) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        cslam_keypoint kp;
        kp.x = __fmul_rn((float)kx, g.scale); kp.y = __fmul_rn((float)ky, g.scale);
        kp.size = g.sizeF; kp.angle = angle; kp.response = (float)resp; kp.octave = level; kp.class_id = -1;
        A.kps[(size_t)frame * A.kpCap + outIdx] = kp;
    }
}

}  // namespace cslam

// =================================================================================================== host side
using namespace cslam;

struct cslam_frontend {
    int device = 0;
    cslam_cam_params cam;
    cslam_orb_params orb;
    int W = 0, H = 0, CW = 0, CH = 0, maxBatch = 0, kpCap = 0, keptCap = 0, M = 0;
    cudaStream_t stream = nullptr;
    DevLevels L;                     // host copy of the device-pointer table (passed by value to kernels)
    std::vector<float> scale, invScale, sigma2, invSigma2; std::vector<int> perLevel; int umax[16];
    std::vector<float> map1, map2;   // host float maps (3H x 3W), kept for cslam_frontend_get_maps
    uint32_t* d_map = nullptr;       // 5 x W x W
    uint8_t* d_mask = nullptr; int maskPitch = 0;
    uint8_t* d_fisheye = nullptr;    // staging for host entry points: maxBatch x Ih x Iw
    uint32_t* d_candCount = nullptr; // maxBatch x nlevels
    uint32_t* d_kept = nullptr; uint32_t* d_keptCount = nullptr;
    cslam_keypoint* d_kps = nullptr; uint8_t* d_desc = nullptr; int32_t* d_nout = nullptr;
    int* d_err = nullptr;
    uint8_t* h_pin_in = nullptr; cslam_keypoint* h_pin_kps = nullptr; uint8_t* h_pin_desc = nullptr; int32_t* h_pin_n = nullptr; int* h_pin_err = nullptr;
    std::vector<void*> owned;
    static const int MAX_LANES = 4;
    cudaStream_t lane[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr}; cudaEvent_t evLane[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr}; cudaEvent_t evFork = nullptr;
    int devLanes = 2;                // lanes used by cslam_frontend_run_dev (env CSLAM_DEV_LANES)
    int hostChunks = 8;              // chunks a host-buffer batch is cut into (env CSLAM_HOST_CHUNKS)
    int64_t launches = 0;
    int lastBatch = 0;
    // optional per-kernel timing (cslam_frontend_set_timing): events between launches, accumulated per kernel kind
    bool timing = false; std::vector<cudaEvent_t> ev; std::vector<int> evKind; int evUsed = 0;
    double kindMs[8] = {0, 0, 0, 0, 0, 0, 0, 0}; int64_t kindCount[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool cornersDirty = false;   // level-0 buffer holds caller canvases (non-zero corner tiles) from cslam_orb_extract
    size_t distributeSmem = 0;
};

enum { KIND_WARP = 0, KIND_PYRAMID, KIND_FAST, KIND_DISTRIBUTE, KIND_DESCRIBE, KIND_END, KIND_COUNT };
static const char* kKindNames[KIND_COUNT] = {"k_warp", "k_pyramid", "k_fast", "k_distribute", "k_describe", "end"};
static inline void mark(cslam_frontend* fe, cudaStream_t st, int kind) {
    if (kind != KIND_END) fe->launches++;
    if (!fe->timing || fe->evUsed >= (int)fe->ev.size()) return;
    cudaEventRecord(fe->ev[fe->evUsed], st);
    fe->evKind[fe->evUsed++] = kind;
}
static void collect_timing(cslam_frontend* fe) {
    for (int i = 0; i + 1 < fe->evUsed; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, fe->ev[i], fe->ev[i + 1]) == cudaSuccess) { fe->kindMs[fe->evKind[i]] += ms; fe->kindCount[fe->evKind[i]]++; }
    }
    fe->evUsed = 0;
}

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }

// CamModelGeneral::CubemapToFisheye (src/CamModelGeneral.cpp:265-290) + WorldToImg (include/CamModelGeneral.h:359-374):
// start-up code on the host, exactly where the reference runs it (System::CreateUndistortRectifyMap).
static void host_cubemap_to_fisheye(const cslam_cam_params& cp, double up, double vp, double& uf, double& vf) {
    float i = (float)up, j = (float)vp;
    uf = -1; vf = -1;
    const float fi = i / (float)cp.face_w, fj = j / (float)cp.face_h;
    int face = -1;   // 0 front 1 left 2 right 3 upper 4 lower
    if (fi >= 0 && fi < 1 && fj >= 1 && fj < 2) face = 1;
    else if (fi >= 1 && fi < 2 && fj >= 0 && fj < 1) face = 3;
    else if (fi >= 1 && fi < 2 && fj >= 1 && fj < 2) face = 0;
    else if (fi >= 1 && fi < 2 && fj >= 2 && fj < 3) face = 4;
    else if (fi >= 2 && fi < 3 && fj >= 1 && fj < 2) face = 2;
    if (face < 0) return;
    const double fx = cp.face_w / 2.0, fy = cp.face_h / 2.0;
    i = i - static_cast<int>(i / cp.face_w) * cp.face_w;
    j = j - static_cast<int>(j / cp.face_h) * cp.face_h;
    const double lz = 1.0, lx = (i - fx) * lz / fx, ly = (j - fy) * lz / fy;
    double X, Y, Z;
    switch (face) {
        case 0: X = lx; Y = ly; Z = lz; break;
        case 1: X = -lz; Y = ly; Z = lx; break;
        case 2: X = lz; Y = ly; Z = -lx; break;
        case 4: X = lx; Y = lz; Z = -ly; break;
        default: X = lx; Y = -lz; Z = ly; break;
    }
    double norm = std::sqrt(X * X + Y * Y);
    if (norm == 0.0) norm = 1e-14;
    const double theta = std::atan(-Z / norm);
    double rho = 0.0;
    for (int k = 11; k >= 0; k--) rho = rho * theta + cp.invpoly[k];
    const double uu = X / norm * rho, vv = Y / norm * rho;
    uf = uu * cp.c + vv * cp.d + cp.u0;
    vf = uu * cp.e + vv + cp.v0;
    if (uf < 0 || uf >= cp.Iw || vf < 0 || vf >= cp.Ih) { uf = -1; vf = -1; }
}

template <class T>
static int dev_alloc(cslam_frontend* fe, T** p, size_t count, bool zero = true) {
    void* q = nullptr;
    if (count == 0) count = 1;
    CSLAM_CUDA(cudaMalloc(&q, count * sizeof(T)));
    if (zero) CSLAM_CUDA(cudaMemset(q, 0, count * sizeof(T)));
    fe->owned.push_back(q);
    *p = (T*)q;
    return 0;
}
template <class T>
static int dev_upload(cslam_frontend* fe, const T** p, const std::vector<T>& v) {
    T* q = nullptr;
    int rc = dev_alloc(fe, &q, v.size(), false);
    if (rc) return rc;
    CSLAM_CUDA(cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *p = q;
    return 0;
}

static void build_resize_tables(int sn, int dn, std::vector<uint16_t>& ofs, std::vector<uint32_t>& ab) {
    ofs.resize(dn); ab.resize(dn);
    const double scale = (double)sn / dn;
    for (int d = 0; d < dn; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0; }
        if (s >= sn - 1) { s = sn - 1; f = 0; }
        const int a0 = cv_round_f((1.f - f) * 2048.f), a1 = cv_round_f(f * 2048.f);
        ofs[d] = (uint16_t)s; ab[d] = (uint32_t)a0 | ((uint32_t)a1 << 16);
    }
}

extern "C" int cslam_frontend_create(cslam_frontend** out, int device, const cslam_cam_params* cam, const cslam_orb_params* orb,
                                     const uint8_t* mask, int mask_pitch, int max_batch) {
    if (!out || !cam || !orb || !mask || max_batch <= 0) { set_error("cslam_frontend_create: null/invalid argument"); return CSLAM_E_BADARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device (this library has no CPU fallback)"); return CSLAM_E_NODEVICE; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return CSLAM_E_BADARG; }
    if (cam->face_w != cam->face_h || cam->face_w <= 0) { set_error("cube faces must be square (got %dx%d)", cam->face_w, cam->face_h); return CSLAM_E_BADARG; }
    if (3 * cam->face_w > 4095 || cam->Iw * 32 > 65535 || cam->Ih * 32 > 65535) { set_error("image too large for the packed coordinate formats"); return CSLAM_E_BADARG; }
    if (orb->nlevels < 1 || orb->nlevels > MAX_LEVELS || orb->nfeatures < 1 || orb->scale_factor <= 1.f || orb->scale_factor > 2.f) { set_error("bad ORB parameters (need 1 < scaleFactor <= 2)"); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(device));
    cslam_frontend* fe = new cslam_frontend;
    fe->device = device; fe->cam = *cam; fe->orb = *orb; fe->maxBatch = max_batch;
    fe->W = cam->face_w; fe->H = cam->face_h; fe->CW = 3 * fe->W; fe->CH = 3 * fe->H;
    const int nl = orb->nlevels;
    // ---- ORBextractor::ORBextractor (src/ORBExtractor.cpp:381-442): scale tables, per-level quotas, umax
    fe->scale.resize(nl); fe->invScale.resize(nl); fe->sigma2.resize(nl); fe->invSigma2.resize(nl); fe->perLevel.resize(nl);
    fe->scale[0] = 1.f; fe->sigma2[0] = 1.f;
    for (int i = 1; i < nl; i++) { fe->scale[i] = fe->scale[i - 1] * orb->scale_factor; fe->sigma2[i] = fe->scale[i] * fe->scale[i]; }
    for (int i = 0; i < nl; i++) { fe->invScale[i] = 1.0f / fe->scale[i]; fe->invSigma2[i] = 1.0f / fe->sigma2[i]; }
    {
        const float factor = 1.0f / orb->scale_factor;
        float nDesired = orb->nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; l++) { fe->perLevel[l] = cv_round_f(nDesired); sum += fe->perLevel[l]; nDesired *= factor; }
        fe->perLevel[nl - 1] = std::max(orb->nfeatures - sum, 0);
        int v, v0, vmax = (int)std::floor(HALF_PATCH * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(HALF_PATCH * std::sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH * HALF_PATCH;
        for (v = 0; v <= vmax; ++v) fe->umax[v] = cv_round_d(std::sqrt(hp2 - v * v));
        for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) { while (fe->umax[v0] == fe->umax[v0 + 1]) ++v0; fe->umax[v] = v0; ++v0; }
    }
    int rc = 0;
    auto fail = [&](int code) { cslam_frontend_destroy(fe); return code; };
    if (cudaStreamCreateWithFlags(&fe->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(CSLAM_E_CUDA); }
    for (int i = 0; i < cslam_frontend::MAX_LANES; i++)
        if (cudaStreamCreateWithFlags(&fe->lane[i], cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&fe->evLane[i], cudaEventDisableTiming) != cudaSuccess) { set_error("lane stream creation failed"); return fail(CSLAM_E_CUDA); }
    if (cudaEventCreateWithFlags(&fe->evFork, cudaEventDisableTiming) != cudaSuccess) { set_error("event creation failed"); return fail(CSLAM_E_CUDA); }
    if (const char* e = getenv("CSLAM_DEV_LANES")) fe->devLanes = std::max(1, std::min(atoi(e), (int)cslam_frontend::MAX_LANES));
    if (const char* e = getenv("CSLAM_HOST_CHUNKS")) fe->hostChunks = std::max(1, std::min(atoi(e), 64));
    if (cudaMemcpyToSymbol(c_umax, fe->umax, sizeof(fe->umax)) != cudaSuccess) { set_error("cudaMemcpyToSymbol failed"); return fail(CSLAM_E_CUDA); }
    {
        static const signed char hostPattern[1024] = {
#include "brief_pattern.inc"
        };
        uint32_t patT[8 * 32];
        for (int ln = 0; ln < 32; ln++)
            for (int k = 0; k < 8; k++) {
                const unsigned char* q = reinterpret_cast<const unsigned char*>(hostPattern) + ln * 32 + 4 * k;
                patT[k * 32 + ln] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
            }
        if (cudaMemcpyToSymbol(g_patT, patT, sizeof(patT)) != cudaSuccess) { set_error("cudaMemcpyToSymbol failed"); return fail(CSLAM_E_CUDA); }
    }
    // ---- level geometry (ComputePyramid :930-936, ComputeKeyPointsOctTree :747-761)
    std::memset(&fe->L, 0, sizeof(fe->L));
    fe->L.nlevels = nl;
    int maxQuota = 0;
    for (int l = 0; l < nl; l++) {
        LevelGeom& g = fe->L.g[l];
        g.w = cv_round_f((float)fe->CW * fe->invScale[l]); g.h = cv_round_f((float)fe->CH * fe->invScale[l]);
        g.pitch = round_up(g.w, 128);
        g.minB = EDGE_THRESHOLD - 3; g.maxBX = g.w - EDGE_THRESHOLD + 3; g.maxBY = g.h - EDGE_THRESHOLD + 3;
        const float width = (float)(g.maxBX - g.minB), height = (float)(g.maxBY - g.minB);
        const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
        if (nCols < 1 || nRows < 1) { set_error("pyramid level %d (%dx%d) is too small for the 30-px FAST grid", l, g.w, g.h); return fail(CSLAM_E_BADARG); }
        g.wCell = (int)std::ceil(width / nCols); g.hCell = (int)std::ceil(height / nRows);
        if (g.wCell > CELL_MAX || g.hCell > CELL_MAX) { set_error("FAST cell %dx%d exceeds the supported %d px", g.wCell, g.hCell, CELL_MAX); return fail(CSLAM_E_BADARG); }
        g.nColsEff = 0; g.nRowsEff = 0;
        for (int j = 0; j < nCols; j++) if (!((float)(g.minB + j * g.wCell) >= g.maxBX - 6)) g.nColsEff = j + 1;
        for (int i = 0; i < nRows; i++) if (!((float)(g.minB + i * g.hCell) >= g.maxBY - 3)) g.nRowsEff = i + 1;
        g.quota = fe->perLevel[l]; maxQuota = std::max(maxQuota, g.quota);
        g.candCap = std::max(4096, g.nColsEff * g.nRowsEff * 32);
        g.scale = fe->scale[l]; g.sizeF = (float)(int)(PATCH_SIZE * fe->scale[l]);
        if ((rc = dev_alloc(fe, &fe->L.img[l], (size_t)max_batch * g.pitch * g.h + 256))) return fail(rc);   // slack: word loads may run a few bytes past the last row
        if ((rc = dev_alloc(fe, &fe->L.cand[l], (size_t)max_batch * g.candCap, false))) return fail(rc);
        if ((rc = dev_alloc(fe, &fe->L.pnode[l], (size_t)max_batch * g.candCap, false))) return fail(rc);
        if (l > 0) {
            std::vector<uint16_t> xo, yo; std::vector<uint32_t> xa, ya;
            build_resize_tables(fe->L.g[l - 1].w, g.w, xo, xa); build_resize_tables(fe->L.g[l - 1].h, g.h, yo, ya);
            if ((rc = dev_upload(fe, &fe->L.xofs[l], xo)) || (rc = dev_upload(fe, &fe->L.xab[l], xa)) || (rc = dev_upload(fe, &fe->L.yofs[l], yo)) ||
                (rc = dev_upload(fe, &fe->L.yab[l], ya))) return fail(rc);
        }
    }
    // ---- static zero map: Z_0 = corner tiles of the canvas (never written by k_warp, zeroed at allocation); Z_l(x,y) = all four
    // resize taps of (x,y) are in Z_{l-1}. A k_fast tile inside Z_l is all-zero, hence corner-free: its CTA exits at once.
    {
        std::vector<uint8_t> Z((size_t)fe->CW * fe->CH, 0), Zn;
        for (int y = 0; y < fe->CH; y++)
            for (int x = 0; x < fe->CW; x++) {
                const int tc = x / fe->W, tr = y / fe->H;
                Z[(size_t)y * fe->CW + x] = (tc != 1 && tr != 1) ? 1 : 0;
            }
        for (int l = 0; l < nl; l++) {
            const LevelGeom& g = fe->L.g[l];
            if (l > 0) {
                const LevelGeom& sg = fe->L.g[l - 1];
                std::vector<uint16_t> xo, yo; std::vector<uint32_t> xa, ya;
                build_resize_tables(sg.w, g.w, xo, xa); build_resize_tables(sg.h, g.h, yo, ya);
                Zn.assign((size_t)g.w * g.h, 0);
                for (int y = 0; y < g.h; y++) {
                    const int y0 = yo[y], y1 = std::min(y0 + 1, sg.h - 1);
                    for (int x = 0; x < g.w; x++) {
                        const int x0 = xo[x], x1 = std::min(x0 + 1, sg.w - 1);
                        Zn[(size_t)y * g.w + x] = Z[(size_t)y0 * sg.w + x0] & Z[(size_t)y0 * sg.w + x1] & Z[(size_t)y1 * sg.w + x0] & Z[(size_t)y1 * sg.w + x1];
                    }
                }
                Z.swap(Zn);
            }
            const int gx = cdiv(g.nColsEff, CG), gy = g.nRowsEff;
            std::vector<uint8_t> sk((size_t)gx * gy, 0);
            for (int cy = 0; cy < gy; cy++)
                for (int cxg = 0; cxg < gx; cxg++) {
                    const int iniY = g.minB + cy * g.hCell, maxY = std::min(iniY + g.hCell + 6, g.maxBY);
                    const int ncell = std::min(CG, g.nColsEff - cxg * CG);
                    const int tx0 = g.minB + cxg * CG * g.wCell, tx1 = std::min(tx0 + ncell * g.wCell + 6, g.maxBX);
                    uint8_t all = 1;
                    for (int y = iniY; y < maxY && all; y++)
                        for (int x = tx0; x < tx1; x++) if (!Z[(size_t)y * g.w + x]) { all = 0; break; }
                    sk[(size_t)cy * gx + cxg] = all;
                }
            if ((rc = dev_upload(fe, &fe->L.skip[l], sk))) return fail(rc);
        }
    }
    fe->M = round_up(std::max(maxQuota + 4, 16), 32);
    fe->keptCap = fe->M;
    fe->kpCap = orb->nfeatures + 3 * nl;
    fe->distributeSmem = (size_t)fe->M * (2 * sizeof(QNode) + 4 * 4 + 4 + 4 + 2 + 2 + 1);
    fe->distributeSmem = (fe->distributeSmem + 15) & ~(size_t)15;
    if (fe->distributeSmem > 200 * 1024) { set_error("nfeatures too large for the shared-memory quadtree (%zu B)", fe->distributeSmem); return fail(CSLAM_E_BADARG); }
    if (cudaFuncSetAttribute(k_distribute, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe->distributeSmem) != cudaSuccess) { set_error("cudaFuncSetAttribute failed"); return fail(CSLAM_E_CUDA); }
    // ---- System::CreateUndistortRectifyMap (src/System.cpp:301-324) on the host, then quantised like cv::remap does
    // (skipped when no fisheye geometry is given: extraction-only front end, e.g. the ORBextractor facade)
    if (cam->Iw > 0 && cam->Ih > 0) {
        const int CW = fe->CW, CH = fe->CH, W = fe->W;
        fe->map1.assign((size_t)CW * CH, 0.f); fe->map2.assign((size_t)CW * CH, 0.f);
        for (int y = 0; y < CH; y++)
            for (int x = 0; x < CW; x++) {
                double u, v;
                host_cubemap_to_fisheye(*cam, (double)x, (double)y, u, v);
                if (u < 0 || v < 0 || u >= cam->Iw || v >= cam->Ih) continue;
                fe->map1[(size_t)y * CW + x] = (float)u; fe->map2[(size_t)y * CW + x] = (float)v;
            }
        std::vector<uint32_t> q((size_t)5 * W * W);
        const int tc[5] = {1, 0, 2, 1, 1}, tr[5] = {1, 1, 1, 0, 2};
        for (int f = 0; f < 5; f++)
            for (int y = 0; y < W; y++)
                for (int x = 0; x < W; x++) {
                    const size_t src = (size_t)(tr[f] * W + y) * CW + tc[f] * W + x;
                    const int sx = cv_round_f(fe->map1[src] * 32.f), sy = cv_round_f(fe->map2[src] * 32.f);
                    q[((size_t)f * W + y) * W + x] = (uint32_t)sx | ((uint32_t)sy << 16);
                }
        const uint32_t* dq = nullptr;
        if ((rc = dev_upload(fe, &dq, q))) return fail(rc);
        fe->d_map = const_cast<uint32_t*>(dq);
    }
    // ---- mask
    fe->maskPitch = round_up(fe->CW, 128);
    if ((rc = dev_alloc(fe, &fe->d_mask, (size_t)fe->maskPitch * fe->CH))) return fail(rc);
    if (cudaMemcpy2D(fe->d_mask, fe->maskPitch, mask, mask_pitch, fe->CW, fe->CH, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("mask upload failed"); return fail(CSLAM_E_CUDA); }
    // ---- batch buffers
    const size_t B = max_batch;
    if ((rc = dev_alloc(fe, &fe->d_fisheye, B * cam->Iw * cam->Ih)) || (rc = dev_alloc(fe, &fe->d_candCount, B * nl)) ||
        (rc = dev_alloc(fe, &fe->d_kept, B * nl * fe->keptCap)) || (rc = dev_alloc(fe, &fe->d_keptCount, B * nl)) ||
        (rc = dev_alloc(fe, &fe->d_kps, B * fe->kpCap)) || (rc = dev_alloc(fe, &fe->d_desc, B * fe->kpCap * 32)) ||
        (rc = dev_alloc(fe, &fe->d_nout, B)) || (rc = dev_alloc(fe, &fe->d_err, 1))) return fail(rc);
    const size_t inBytes = std::max((size_t)cam->Iw * cam->Ih, (size_t)fe->L.g[0].pitch * fe->CH) * B;
    if (cudaMallocHost(&fe->h_pin_in, inBytes) != cudaSuccess || cudaMallocHost(&fe->h_pin_kps, B * fe->kpCap * sizeof(cslam_keypoint)) != cudaSuccess ||
        cudaMallocHost(&fe->h_pin_desc, B * fe->kpCap * 32) != cudaSuccess || cudaMallocHost(&fe->h_pin_n, B * sizeof(int32_t)) != cudaSuccess ||
        cudaMallocHost(&fe->h_pin_err, sizeof(int)) != cudaSuccess) { set_error("pinned host allocation failed"); return fail(CSLAM_E_CUDA); }
    *out = fe;
    return CSLAM_OK;
}

extern "C" void cslam_frontend_destroy(cslam_frontend* fe) {
    if (!fe) return;
    cudaSetDevice(fe->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < cslam_frontend::MAX_LANES; i++) { if (fe->lane[i]) cudaStreamDestroy(fe->lane[i]); if (fe->evLane[i]) cudaEventDestroy(fe->evLane[i]); }
    if (fe->evFork) cudaEventDestroy(fe->evFork);
    if (fe->stream) cudaStreamDestroy(fe->stream);
    for (void* p : fe->owned) cudaFree(p);
    for (auto& e : fe->ev) cudaEventDestroy(e);
    if (fe->h_pin_in) cudaFreeHost(fe->h_pin_in);
    if (fe->h_pin_kps) cudaFreeHost(fe->h_pin_kps);
    if (fe->h_pin_desc) cudaFreeHost(fe->h_pin_desc);
    if (fe->h_pin_n) cudaFreeHost(fe->h_pin_n);
    if (fe->h_pin_err) cudaFreeHost(fe->h_pin_err);
    delete fe;
}

extern "C" int cslam_frontend_set_timing(cslam_frontend* fe, int enable) {
    if (!fe) return CSLAM_E_BADARG;
    CSLAM_CUDA(cudaSetDevice(fe->device));
    if (enable && fe->ev.empty()) {
        fe->ev.resize(256); fe->evKind.resize(256);
        for (auto& e : fe->ev) CSLAM_CUDA(cudaEventCreate(&e));
    }
    fe->timing = enable != 0; fe->evUsed = 0;
    for (int i = 0; i < 8; i++) { fe->kindMs[i] = 0; fe->kindCount[i] = 0; }
    return CSLAM_OK;
}
extern "C" int cslam_frontend_get_timing(const cslam_frontend* fe, int kind, const char** name, double* ms, int64_t* count) {
    if (!fe || kind < 0 || kind >= KIND_END) return CSLAM_E_BADARG;
    *name = kKindNames[kind]; *ms = fe->kindMs[kind]; *count = fe->kindCount[kind];
    return CSLAM_OK;
}
extern "C" int cslam_frontend_kp_capacity(const cslam_frontend* fe) { return fe ? fe->kpCap : 0; }
extern "C" void* cslam_frontend_stream(const cslam_frontend* fe) { return fe ? (void*)fe->stream : nullptr; }
extern "C" int64_t cslam_frontend_launches(const cslam_frontend* fe) { return fe ? fe->launches : 0; }

extern "C" int cslam_frontend_sync(cslam_frontend* fe) {
    if (!fe) return CSLAM_E_BADARG;
    CSLAM_CUDA(cudaSetDevice(fe->device));
    CSLAM_CUDA(cudaMemcpyAsync(fe->h_pin_err, fe->d_err, sizeof(int), cudaMemcpyDeviceToHost, fe->stream));
    CSLAM_CUDA(cudaStreamSynchronize(fe->stream));
    if (fe->timing) collect_timing(fe);
    if (*fe->h_pin_err != 0) {
        const int code = *fe->h_pin_err;
        cudaMemsetAsync(fe->d_err, 0, sizeof(int), fe->stream);
        set_error("front end: a fixed-capacity buffer overflowed on the device (candidates / nodes / keypoints)");
        return code;
    }
    return CSLAM_OK;
}

// Frame range [f0, f0+cnt) of the per-batch buffers, for one lane
static DevLevels shifted(const cslam_frontend* fe, int f0) {
    DevLevels L = fe->L;
    for (int l = 0; l < L.nlevels; l++) {
        const LevelGeom& g = L.g[l];
        L.img[l] += (size_t)f0 * g.pitch * g.h; L.cand[l] += (size_t)f0 * g.candCap; L.pnode[l] += (size_t)f0 * g.candCap;
    }
    return L;
}

static int launch_warp(cslam_frontend* fe, cudaStream_t st, const uint8_t* d_fisheye, int f0, int cnt) {
    const int FPT = 4;
    if (!fe->d_map) { set_error("this front end was created without fisheye geometry (Iw/Ih = 0): no warp"); return CSLAM_E_BADARG; }
    const size_t cframe = (size_t)fe->L.g[0].pitch * fe->CH;
    dim3 grid(cdiv(fe->W, 256), fe->H, 5 * cdiv(cnt, FPT));
    mark(fe, st, KIND_WARP);
    k_warp<FPT><<<grid, 256, 0, st>>>(d_fisheye, fe->cam.Iw, fe->cam.Ih, fe->d_map, fe->W, fe->L.img[0] + (size_t)f0 * cframe, fe->L.g[0].pitch, cframe, cnt);
    CSLAM_CUDA(cudaGetLastError());
    return 0;
}

static int launch_extract(cslam_frontend* fe, cudaStream_t st, int f0, int cnt, cslam_keypoint* d_kps, uint8_t* d_desc, int32_t* d_nout, bool useSkip) {
    const int nl = fe->L.nlevels;
    const DevLevels L = shifted(fe, f0);
    uint32_t* candCount = fe->d_candCount + (size_t)f0 * nl;
    CSLAM_CUDA(cudaMemsetAsync(candCount, 0, (size_t)cnt * nl * sizeof(uint32_t), st));
    for (int l = 1; l < nl; l++) {
        const LevelGeom& s = L.g[l - 1]; const LevelGeom& d = L.g[l];
        dim3 grid(cdiv(cdiv(d.w, 4), 128), cdiv(d.h, PYR_RY), cnt);
        mark(fe, st, KIND_PYRAMID);
        k_pyramid<<<grid, 128, 0, st>>>(L.img[l - 1], s.w, s.h, s.pitch, L.img[l], d.w, d.h, d.pitch, L.xofs[l], L.xab[l], L.yofs[l], L.yab[l]);
    }
    for (int l = 0; l < nl; l++) {
        const LevelGeom& g = L.g[l];
        dim3 grid(cdiv(g.nColsEff, CG), g.nRowsEff, cnt);
        mark(fe, st, KIND_FAST);
        k_fast<<<grid, 256, 0, st>>>(L.img[l], g, fe->orb.ini_th_fast, fe->orb.min_th_fast, L.cand[l], candCount + l, nl, fe->d_err, useSkip ? L.skip[l] : nullptr);
    }
    CSLAM_CUDA(cudaGetLastError());
    DistributeArgs da;
    da.L = L; da.candCount = candCount; da.kept = fe->d_kept + (size_t)f0 * nl * fe->keptCap; da.keptCount = fe->d_keptCount + (size_t)f0 * nl;
    da.mask = fe->d_mask; da.maskPitch = fe->maskPitch;
    da.imgW = fe->CW; da.imgH = fe->CH; da.faceW = fe->W; da.faceH = fe->H; da.keptCap = fe->keptCap; da.M = fe->M; da.errFlag = fe->d_err;
    mark(fe, st, KIND_DISTRIBUTE);
    k_distribute<<<dim3(nl, cnt), 256, fe->distributeSmem, st>>>(da);
    DescribeArgs ds;
    ds.L = L; ds.kept = da.kept; ds.keptCount = da.keptCount; ds.keptCap = fe->keptCap; ds.kps = d_kps; ds.desc = d_desc; ds.nOut = d_nout; ds.kpCap = fe->kpCap;
    ds.errFlag = fe->d_err;
    mark(fe, st, KIND_DESCRIBE);
    k_describe<<<dim3(cdiv(fe->keptCap, DESC_WARPS), nl, cnt), DESC_WARPS * 32, 0, st>>>(ds);
    mark(fe, st, KIND_END);
    CSLAM_CUDA(cudaGetLastError());
    return 0;
}

static int check_batch(cslam_frontend* fe, int batch, const void* a, const void* b) {
    if (!fe || !a || !b || batch <= 0 || batch > fe->maxBatch) { set_error("bad argument (batch must be in 1..%d)", fe ? fe->maxBatch : 0); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(fe->device));
    return 0;
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// Lanes: the batch is cut into contiguous frame ranges, each enqueued on its own stream (forked from / joined into the
// front end's main stream). Different lanes touch disjoint slices of every buffer, so the copy engines and the
// latency-bound k_distribute of one lane overlap with the ALU-bound kernels of the others.
static int fork_lanes(cslam_frontend* fe, int nl) {
    CSLAM_CUDA(cudaEventRecord(fe->evFork, fe->stream));
    for (int i = 0; i < nl; i++) CSLAM_CUDA(cudaStreamWaitEvent(fe->lane[i], fe->evFork, 0));
    return 0;
}
static int join_lanes(cslam_frontend* fe, int nl) {
    for (int i = 0; i < nl; i++) { CSLAM_CUDA(cudaEventRecord(fe->evLane[i], fe->lane[i])); CSLAM_CUDA(cudaStreamWaitEvent(fe->stream, fe->evLane[i], 0)); }
    return 0;
}
static int prepare_level0(cslam_frontend* fe, bool fromWarp) {
    if (fromWarp && fe->cornersDirty) {   // the warp never writes the 4 corner tiles; they must read as the reference's zeroed canvas
        CSLAM_CUDA(cudaMemsetAsync(fe->L.img[0], 0, (size_t)fe->maxBatch * fe->L.g[0].pitch * fe->CH, fe->stream));
        fe->cornersDirty = false;
    }
    if (!fromWarp) fe->cornersDirty = true;
    return 0;
}
static int lanes_for(const cslam_frontend* fe, int batch, int want) {
    if (fe->timing) return 1;
    return std::max(1, std::min(std::min(want, cslam_frontend::MAX_LANES), batch));
}

extern "C" int cslam_warp(cslam_frontend* fe, const uint8_t* fisheye, int batch, uint8_t* canvas, int canvas_pitch) {
    int rc = check_batch(fe, batch, fisheye, canvas);
    if (rc) return rc;
    if ((rc = prepare_level0(fe, true))) return rc;
    const size_t fb = (size_t)fe->cam.Iw * fe->cam.Ih * batch;
    const uint8_t* src = fisheye;
    if (!is_pinned(fisheye)) { std::memcpy(fe->h_pin_in, fisheye, fb); src = fe->h_pin_in; }
    CSLAM_CUDA(cudaMemcpyAsync(fe->d_fisheye, src, fb, cudaMemcpyHostToDevice, fe->stream));
    if ((rc = launch_warp(fe, fe->stream, fe->d_fisheye, 0, batch))) return rc;
    // the reference writes only the five face ROIs; corner tiles of the caller's canvas are left untouched
    const int W = fe->W, pitch = fe->L.g[0].pitch;
    const int tc[5] = {1, 0, 2, 1, 1}, tr[5] = {1, 1, 1, 0, 2};
    for (int f = 0; f < batch; f++)
        for (int t = 0; t < 5; t++)
            CSLAM_CUDA(cudaMemcpy2DAsync(canvas + (size_t)f * canvas_pitch * fe->CH + (size_t)tr[t] * W * canvas_pitch + tc[t] * W, canvas_pitch,
                                         fe->L.img[0] + (size_t)f * pitch * fe->CH + (size_t)tr[t] * W * pitch + tc[t] * W, pitch, W, W, cudaMemcpyDeviceToHost, fe->stream));
    return cslam_frontend_sync(fe);
}

// host entry points: H2D -> kernels -> D2H per lane; caller buffers are used directly when they are page-locked
static int run_host(cslam_frontend* fe, const uint8_t* in, size_t inFrameBytes, int inPitch, bool fromWarp, int batch, cslam_keypoint* kps, uint8_t* desc, int32_t* n_out) {
    int rc;
    if ((rc = prepare_level0(fe, fromWarp))) return rc;
    const int nlanes = lanes_for(fe, batch, 4);
    // the batch is cut into more chunks than there are streams (chunk c runs on stream c % nlanes): the first upload - the only one nothing overlaps -
    // is then 1/8 of the batch instead of 1/4, and the copy engines stay busy behind the kernels of earlier chunks
    const int nchunks = fe->timing ? 1 : std::max(nlanes, std::min(fe->hostChunks, std::max(1, batch / 8)));
    const bool pinIn = is_pinned(in), pinK = is_pinned(kps), pinD = is_pinned(desc), pinN = is_pinned(n_out);
    const size_t nk = (size_t)fe->kpCap;
    if ((rc = fork_lanes(fe, nlanes))) return rc;
    // errors inside the loop leave through join_lanes: a lane that was forked is always joined before this function returns
    auto chunk = [&](int ck) -> int {
        int rc = 0;
        const int f0 = (int)((long long)batch * ck / nchunks), f1 = (int)((long long)batch * (ck + 1) / nchunks), cnt = f1 - f0;
        if (cnt <= 0) return 0;
        cudaStream_t st = fe->lane[ck % nlanes];
        if (fromWarp) {
            const uint8_t* src = in + (size_t)f0 * inFrameBytes;
            if (!pinIn) { std::memcpy(fe->h_pin_in + (size_t)f0 * inFrameBytes, src, (size_t)cnt * inFrameBytes); src = fe->h_pin_in + (size_t)f0 * inFrameBytes; }
            uint8_t* dfe = fe->d_fisheye + (size_t)f0 * inFrameBytes;
            CSLAM_CUDA(cudaMemcpyAsync(dfe, src, (size_t)cnt * inFrameBytes, cudaMemcpyHostToDevice, st));
            if ((rc = launch_warp(fe, st, dfe, f0, cnt))) return rc;
        } else {
            const int pitch = fe->L.g[0].pitch;
            CSLAM_CUDA(cudaMemcpy2DAsync(fe->L.img[0] + (size_t)f0 * pitch * fe->CH, pitch, in + (size_t)f0 * inPitch * fe->CH, inPitch, fe->CW, (size_t)fe->CH * cnt,
                                         cudaMemcpyHostToDevice, st));
        }
        if ((rc = launch_extract(fe, st, f0, cnt, fe->d_kps + (size_t)f0 * nk, fe->d_desc + (size_t)f0 * nk * 32, fe->d_nout + f0, fromWarp))) return rc;
        CSLAM_CUDA(cudaMemcpyAsync((pinK ? kps : fe->h_pin_kps) + (size_t)f0 * nk, fe->d_kps + (size_t)f0 * nk, (size_t)cnt * nk * sizeof(cslam_keypoint), cudaMemcpyDeviceToHost, st));
        CSLAM_CUDA(cudaMemcpyAsync((pinD ? desc : fe->h_pin_desc) + (size_t)f0 * nk * 32, fe->d_desc + (size_t)f0 * nk * 32, (size_t)cnt * nk * 32, cudaMemcpyDeviceToHost, st));
        CSLAM_CUDA(cudaMemcpyAsync((pinN ? n_out : fe->h_pin_n) + f0, fe->d_nout + f0, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        return 0;
    };
    int erc = 0;
    for (int ck = 0; ck < nchunks && !erc; ck++) erc = chunk(ck);
    rc = join_lanes(fe, nlanes);
    if (erc) { cudaStreamSynchronize(fe->stream); return erc; }
    if (rc) return rc;
    fe->lastBatch = batch;
    if ((rc = cslam_frontend_sync(fe))) return rc;
    const size_t tot = (size_t)batch * nk;
    if (!pinK) std::memcpy(kps, fe->h_pin_kps, tot * sizeof(cslam_keypoint));
    if (!pinD) std::memcpy(desc, fe->h_pin_desc, tot * 32);
    if (!pinN) std::memcpy(n_out, fe->h_pin_n, batch * sizeof(int32_t));
    return CSLAM_OK;
}

extern "C" int cslam_orb_extract(cslam_frontend* fe, const uint8_t* canvas, int canvas_pitch, int batch, cslam_keypoint* kps, uint8_t* desc, int32_t* n_out) {
    int rc = check_batch(fe, batch, canvas, kps);
    if (rc) return rc;
    if (!desc || !n_out || canvas_pitch < fe->CW) { set_error("cslam_orb_extract: bad argument"); return CSLAM_E_BADARG; }
    return run_host(fe, canvas, 0, canvas_pitch, false, batch, kps, desc, n_out);
}

extern "C" int cslam_frontend_run(cslam_frontend* fe, const uint8_t* fisheye, int batch, cslam_keypoint* kps, uint8_t* desc, int32_t* n_out) {
    int rc = check_batch(fe, batch, fisheye, kps);
    if (rc) return rc;
    if (!desc || !n_out) { set_error("cslam_frontend_run: null output"); return CSLAM_E_BADARG; }
    return run_host(fe, fisheye, (size_t)fe->cam.Iw * fe->cam.Ih, 0, true, batch, kps, desc, n_out);
}

extern "C" int cslam_frontend_run_dev(cslam_frontend* fe, const uint8_t* fisheye_dev, int batch, cslam_keypoint* kps_dev, uint8_t* desc_dev, int32_t* n_out_dev) {
    int rc = check_batch(fe, batch, fisheye_dev, kps_dev);
    if (rc) return rc;
    if (!desc_dev || !n_out_dev) { set_error("cslam_frontend_run_dev: null output"); return CSLAM_E_BADARG; }
    if ((rc = prepare_level0(fe, true))) return rc;
    const int nlanes = lanes_for(fe, batch, fe->devLanes);
    const size_t nk = (size_t)fe->kpCap, fsz = (size_t)fe->cam.Iw * fe->cam.Ih;
    if ((rc = fork_lanes(fe, nlanes))) return rc;
    int erc = 0;
    for (int ln = 0; ln < nlanes && !erc; ln++) {
        const int f0 = (int)((long long)batch * ln / nlanes), f1 = (int)((long long)batch * (ln + 1) / nlanes), cnt = f1 - f0;
        if (cnt <= 0) continue;
        if ((erc = launch_warp(fe, fe->lane[ln], fisheye_dev + (size_t)f0 * fsz, f0, cnt))) break;
        erc = launch_extract(fe, fe->lane[ln], f0, cnt, kps_dev + (size_t)f0 * nk, desc_dev + (size_t)f0 * nk * 32, n_out_dev + f0, true);
    }
    fe->lastBatch = batch;
    rc = join_lanes(fe, nlanes);   // forked lanes are joined on the error path too
    return erc ? erc : rc;
}

extern "C" int cslam_frontend_level_size(const cslam_frontend* fe, int level, int* w, int* h) {
    if (!fe || level < 0 || level >= fe->L.nlevels) return CSLAM_E_BADARG;
    *w = fe->L.g[level].w; *h = fe->L.g[level].h;
    return CSLAM_OK;
}
extern "C" int cslam_frontend_get_level(cslam_frontend* fe, int frame, int level, uint8_t* out) {
    if (!fe || level < 0 || level >= fe->L.nlevels || frame < 0 || frame >= fe->maxBatch || !out) return CSLAM_E_BADARG;
    const LevelGeom& g = fe->L.g[level];
    CSLAM_CUDA(cudaSetDevice(fe->device));
    CSLAM_CUDA(cudaStreamSynchronize(fe->stream));
    CSLAM_CUDA(cudaMemcpy2D(out, g.w, fe->L.img[level] + (size_t)frame * g.pitch * g.h, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost));
    return CSLAM_OK;
}
extern "C" int cslam_frontend_get_candidates(cslam_frontend* fe, int frame, int level, int32_t* xyr, int cap, int* n) {
    if (!fe || level < 0 || level >= fe->L.nlevels || frame < 0 || frame >= fe->maxBatch || !xyr || !n) return CSLAM_E_BADARG;
    const LevelGeom& g = fe->L.g[level];
    CSLAM_CUDA(cudaSetDevice(fe->device));
    CSLAM_CUDA(cudaStreamSynchronize(fe->stream));
    uint32_t cnt = 0;
    CSLAM_CUDA(cudaMemcpy(&cnt, fe->d_candCount + (size_t)frame * fe->L.nlevels + level, 4, cudaMemcpyDeviceToHost));
    const int m = (int)std::min<uint32_t>(cnt, (uint32_t)g.candCap);
    std::vector<uint32_t> tmp(m);
    if (m) CSLAM_CUDA(cudaMemcpy(tmp.data(), fe->L.cand[level] + (size_t)frame * g.candCap, (size_t)m * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < m && i < cap; i++) { xyr[3 * i] = tmp[i] & 0xfff; xyr[3 * i + 1] = (tmp[i] >> 12) & 0xfff; xyr[3 * i + 2] = tmp[i] >> 24; }
    *n = m;
    return CSLAM_OK;
}
extern "C" int cslam_frontend_get_maps(const cslam_frontend* fe, float* map1, float* map2) {
    if (!fe || !map1 || !map2) return CSLAM_E_BADARG;
    std::memcpy(map1, fe->map1.data(), fe->map1.size() * 4); std::memcpy(map2, fe->map2.data(), fe->map2.size() * 4);
    return CSLAM_OK;
}
extern "C" int cslam_frontend_tables(const cslam_frontend* fe, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* per, int32_t* umax16) {
    if (!fe) return CSLAM_E_BADARG;
    for (int i = 0; i < fe->L.nlevels; i++) { scale[i] = fe->scale[i]; inv_scale[i] = fe->invScale[i]; sigma2[i] = fe->sigma2[i]; inv_sigma2[i] = fe->invSigma2[i]; per[i] = fe->perLevel[i]; }
    for (int i = 0; i < 16; i++) umax16[i] = fe->umax[i];
    return CSLAM_OK;
}
