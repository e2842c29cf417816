// ORACLE (test infrastructure, not product code).
// CPU restatement of the OpenCV primitives the reference hot path calls. OpenCV is an
// un-vendored, un-pinned dependency of the reference (README.md:59); these models are
// pinned to opencv-python 4.13.0 by tests/test_oracle_cv2.py (bit-exact on every case).
//
// Reference call sites:
//   cv::remap            src/System.cpp:350-354
//   cv::resize           src/ORBExtractor.cpp:941
//   cv::copyMakeBorder   src/ORBExtractor.cpp:943-949   (REFLECT_101; apron never read)
//   cv::FAST             src/ORBExtractor.cpp:783-789   (TYPE_9_16, nonmax suppression)
//   cv::GaussianBlur     src/ORBExtractor.cpp:908       (7x7, sigma 2, REFLECT_101)
//   cv::fastAtan2        src/ORBExtractor.cpp:74
//   cvRound              src/ORBExtractor.cpp:52,86,90-91,414,432,933
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

// cvRound: round-half-to-even (SSE cvtss2si / cvtsd2si in the default rounding mode).
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_round(float v) { return (int)lrintf(v); }
static inline int cv_floor(float v) { int i = (int)v; return i - (i > v); }

static inline int reflect101(int p, int n) {
    // gfedcb|abcdefgh|gfedcba
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * (n - 1) - p; }
    return p;
}

// ---------------------------------------------------------------- remap (bilinear, BORDER_CONSTANT 0)
// Fixed-point model of cv::remap(src,dst,map1(32F),map2(32F),INTER_LINEAR,BORDER_CONSTANT,0):
// INTER_BITS=5 (32x32 sub-pixel table), INTER_REMAP_COEF_BITS=15.
struct RemapTab {
    int w[32 * 32][4];   // int, not short: the (fy=0,fx=0) weight is exactly 32768
    RemapTab() {
        float t[32][2];
        for (int i = 0; i < 32; i++) { float x = i * (1.f / 32); t[i][0] = 1.f - x; t[i][1] = x; }
        for (int fy = 0; fy < 32; fy++)
            for (int fx = 0; fx < 32; fx++) {
                int k = 0;
                for (int k1 = 0; k1 < 2; k1++)
                    for (int k2 = 0; k2 < 2; k2++) {
                        float v = t[fy][k1] * t[fx][k2];
                        w[fy * 32 + fx][k++] = cv_round(v * 32768.f);
                    }
            }
    }
};
static inline const RemapTab& remap_tab() { static RemapTab t; return t; }

static inline uint8_t remap_pixel(const uint8_t* src, int sw, int sh, int sstride, float mx, float my) {
    int sx = cv_round(mx * 32.f), sy = cv_round(my * 32.f);
    int ix = sx >> 5, iy = sy >> 5;
    const int* w = remap_tab().w[(sy & 31) * 32 + (sx & 31)];
    auto px = [&](int x, int y) -> int {
        return ((unsigned)x < (unsigned)sw && (unsigned)y < (unsigned)sh) ? src[y * sstride + x] : 0;
    };
    int v = px(ix, iy) * w[0] + px(ix + 1, iy) * w[1] + px(ix, iy + 1) * w[2] + px(ix + 1, iy + 1) * w[3];
    v = (v + (1 << 14)) >> 15;
    return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
}

static inline void remap_bilinear(const uint8_t* src, int sw, int sh, int sstride,
                                  const float* mapx, const float* mapy, int mstride,
                                  uint8_t* dst, int dw, int dh, int dstride) {
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++)
            dst[y * dstride + x] = remap_pixel(src, sw, sh, sstride, mapx[y * mstride + x], mapy[y * mstride + x]);
}

// ---------------------------------------------------------------- resize (INTER_LINEAR, 8U)
struct ResizeAxis { std::vector<int> ofs; std::vector<short> a0, a1; };
static inline ResizeAxis resize_axis(int sn, int dn) {
    ResizeAxis ax; ax.ofs.resize(dn); ax.a0.resize(dn); ax.a1.resize(dn);
    double scale = (double)sn / dn;
    for (int d = 0; d < dn; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0; }
        if (s >= sn - 1) { s = sn - 1; f = 0; }
        ax.ofs[d] = s;
        ax.a0[d] = (short)cv_round((1.f - f) * 2048.f);
        ax.a1[d] = (short)cv_round(f * 2048.f);
    }
    return ax;
}

static inline void resize_linear(const uint8_t* src, int sw, int sh, int sstride,
                                 uint8_t* dst, int dw, int dh, int dstride) {
    ResizeAxis ax = resize_axis(sw, dw), ay = resize_axis(sh, dh);
    std::vector<int> r0(dw), r1(dw);
    auto hrow = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + sy * sstride;
        for (int x = 0; x < dw; x++) {
            int s = ax.ofs[x], s1 = s + 1 < sw ? s + 1 : sw - 1;
            out[x] = S[s] * ax.a0[x] + S[s1] * ax.a1[x];
        }
    };
    for (int y = 0; y < dh; y++) {
        int s = ay.ofs[y], s1 = s + 1 < sh ? s + 1 : sh - 1;
        hrow(s, r0); hrow(s1, r1);
        int b0 = ay.a0[y], b1 = ay.a1[y];
        for (int x = 0; x < dw; x++) {
            int v = (((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2;
            dst[y * dstride + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
}

// ---------------------------------------------------------------- FAST-9/16 with NMS on a ROI
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};

// Arc score s: largest t such that 9 contiguous ring pixels are all > p+t-... i.e. the pixel is a
// FAST corner for every threshold < s. corner(thr) <=> s > thr ; OpenCV response = s-1.
static inline int fast_arc_score(const uint8_t* p, int stride) {
    // max over the 16 arcs of 9 contiguous ring pixels of max(min(d), min(-d)); sliding min/max by doubling (2,4,8,+1)
    int d[16];
    const int c = p[0];
    for (int k = 0; k < 16; k++) d[k] = c - p[kRingDy[k] * stride + kRingDx[k]];
    int mn2[16], mx2[16], mn4[16], mx4[16];
    for (int k = 0; k < 16; k++) { mn2[k] = std::min(d[k], d[(k + 1) & 15]); mx2[k] = std::max(d[k], d[(k + 1) & 15]); }
    for (int k = 0; k < 16; k++) { mn4[k] = std::min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = std::max(mx2[k], mx2[(k + 2) & 15]); }
    int best = 0;
    for (int k = 0; k < 16; k++) {
        const int a = std::min(std::min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int b = std::max(std::max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
        best = std::max(best, std::max(a, -b));
    }
    return best;
}
// corner test at threshold thr (9 contiguous ring pixels all brighter or all darker by more than thr), bit masks
static inline bool fast_is_corner(const uint8_t* p, int stride, int thr) {
    const int c = p[0];
    unsigned hi = 0, lo = 0;
    for (int k = 0; k < 16; k++) { const int v = p[kRingDy[k] * stride + kRingDx[k]]; hi |= (unsigned)(v > c + thr) << k; lo |= (unsigned)(v < c - thr) << k; }
    auto arc9 = [](unsigned m) { m |= m << 16; unsigned a = m & (m >> 1); a &= a >> 2; a &= a >> 4; a &= m >> 8; return (a & 0xffffu) != 0; };
    return arc9(hi) || arc9(lo);
}

struct FastKp { int x, y, response; };

// cv::FAST(roi, kps, thr, true): scores exist for 3<=x<w-3, 3<=y<h-3 of the ROI; anything else
// counts as 0; keep iff score strictly greater than all 8 neighbours; raster order.
static inline void fast_nms(const uint8_t* img, int w, int h, int stride, int thr, std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    static thread_local std::vector<int> sc;
    sc.assign((size_t)w * h, 0);
    // a 9-arc of the 16-ring contains one pixel of every antipodal pair: if both pixels of a pair are within thr of
    // the centre the pixel cannot be a corner (same early-out OpenCV's FAST uses; results unchanged).
    const int o0 = -3 * stride, o8 = 3 * stride, o2 = -2 * stride + 2, o10 = 2 * stride - 2, o6 = 2 * stride + 2, o14 = -2 * stride - 2;
    for (int y = 3; y < h - 3; y++) {
        const uint8_t* row = img + y * stride;
        for (int x = 3; x < w - 3; x++) {
            const uint8_t* p = row + x;
            const int c = p[0], lo = c - thr, hi = c + thr;
            auto in = [&](int v) { return v >= lo && v <= hi; };
            if (in(p[o0]) && in(p[o8])) continue;
            if (in(p[3]) && in(p[-3])) continue;
            if (in(p[o2]) && in(p[o10])) continue;
            if (in(p[o6]) && in(p[o14])) continue;
            if (!fast_is_corner(p, stride, thr)) continue;
            sc[(size_t)y * w + x] = fast_arc_score(p, stride) - 1;   // a corner at thr has score > thr
        }
    }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = sc[(size_t)y * w + x];
            if (!s) continue;
            const int* r = &sc[(size_t)y * w + x];
            if (s > r[-1] && s > r[1] && s > r[-w - 1] && s > r[-w] && s > r[-w + 1] &&
                s > r[w - 1] && s > r[w] && s > r[w + 1])
                out.push_back({x, y, s});
        }
}

// ---------------------------------------------------------------- GaussianBlur 7x7 sigma=2, 8U, REFLECT_101
// OpenCV 4.x bit-exact 8U path: 8.8 fixed-point kernel [18,34,48,56,48,34,18]/256.
static inline void gaussian7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<uint16_t> H((size_t)w * h);
    std::vector<int> xi(w + 6), yi(h + 6);
    for (int x = -3; x < w + 3; x++) xi[x + 3] = reflect101(x, w);
    for (int y = -3; y < h + 3; y++) yi[y + 3] = reflect101(y, h);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        uint16_t* Hr = &H[(size_t)y * w];
        for (int x = 0; x < w; x++) {
            if (x >= 3 && x < w - 3) {
                const uint8_t* q = S + x - 3;
                Hr[x] = (uint16_t)(18 * (q[0] + q[6]) + 34 * (q[1] + q[5]) + 48 * (q[2] + q[4]) + 56 * q[3]);
            } else {
                int s = 0;
                for (int k = 0; k < 7; k++) s += K[k] * S[xi[x + k]];
                Hr[x] = (uint16_t)s;
            }
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* r[7];
        for (int k = 0; k < 7; k++) r[k] = &H[(size_t)yi[y + k] * w];
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t s = 18u * (r[0][x] + r[6][x]) + 34u * (r[1][x] + r[5][x]) + 48u * (r[2][x] + r[4][x]) + 56u * r[3][x];
            D[x] = (uint8_t)((s + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------- fastAtan2 (scalar, fp32, no FMA)
// This translation unit is compiled with -ffp-contract=off so every * and + rounds separately.
static inline float fast_atan2(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------- deterministic sin/cos
// The reference calls glibc cosf/sinf (src/ORBExtractor.cpp:83-84), which is neither correctly
// rounded nor identical across CPUs (ifunc FMA variants). The oracle therefore DEFINES the CPU path:
// sin/cos evaluated in fp64 from IEEE basic ops + fma only, rounded once to fp32. The CUDA kernel
// implements the same operation sequence (spec in DESIGN.md "det_sincos"), so results are bit-identical.
static inline void det_sincosf(float xf, float* s_out, float* c_out) {
    const double x = (double)xf;
    const double two_over_pi = 0.63661977236758134308;
    const double pio2_hi = 1.57079632679489655800e+00;   // 0x3FF921FB54442D18
    const double pio2_lo = 6.12323399573676603587e-17;   // 0x3C91A62633145C07
    double kd = std::nearbyint(x * two_over_pi);
    int k = (int)kd;
    double r = std::fma(-kd, pio2_hi, x);
    r = std::fma(-kd, pio2_lo, r);
    double r2 = r * r;
    // sin(r), |r| <= pi/4: r + r^3 * P(r^2)   (fdlibm __kernel_sin coefficients)
    double ps = 1.58969099521155010221e-10;
    ps = std::fma(ps, r2, -2.50507602534068634195e-08);
    ps = std::fma(ps, r2, 2.75573137070700676789e-06);
    ps = std::fma(ps, r2, -1.98412698298579493134e-04);
    ps = std::fma(ps, r2, 8.33333333332248946124e-03);
    ps = std::fma(ps, r2, -1.66666666666666324348e-01);
    double sn = std::fma(r * r2, ps, r);
    // cos(r): 1 - r^2/2 + r^4 * Q(r^2)      (fdlibm __kernel_cos coefficients)
    double pc = -1.13596475577881948265e-11;
    pc = std::fma(pc, r2, 2.08757232129817482790e-09);
    pc = std::fma(pc, r2, -2.75573143513906633035e-07);
    pc = std::fma(pc, r2, 2.48015872894767294178e-05);
    pc = std::fma(pc, r2, -1.38888888888741095749e-03);
    pc = std::fma(pc, r2, 4.16666666666666019037e-02);
    double cs = std::fma(r2 * r2, pc, std::fma(-0.5, r2, 1.0));
    double sv, cv;
    switch (k & 3) {
        case 0: sv = sn; cv = cs; break;
        case 1: sv = cs; cv = -sn; break;
        case 2: sv = -sn; cv = -cs; break;
        default: sv = -cs; cv = sn; break;
    }
    *s_out = (float)sv;
    *c_out = (float)cv;
}

}  // namespace orc
