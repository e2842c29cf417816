// Shared helpers for the sm_100a kernels of libcubemap_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "../../include/cubemap_b200.h"

namespace cslam {

void set_error(const char* fmt, ...);   // thread-local message behind cslam_last_error()

#define CSLAM_CUDA(expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            cslam::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return CSLAM_E_CUDA;                                                                \
        }                                                                                       \
    } while (0)

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------------------
// Device math with pinned rounding. The CPU path these must match bit-for-bit is specified without FMA
// contraction (oracle/cvprim.h); nvcc contracts a*b+c by default, so every fp32 step is spelled out.

// cv::fastAtan2 (scalar path), degrees.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = 2.220446049250313e-16f;   // (float)DBL_EPSILON
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
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

// det_sincos: fp64 Cody-Waite reduction by pi/2 + fdlibm kernel polynomials, only IEEE mul/fma,
// rounded once to fp32. fp64 mul/fma are bit-identical on host and device, so this equals the CPU definition.
__device__ __forceinline__ void det_sincosf(float xf, float* s_out, float* c_out) {
    const double x = (double)xf;
    const double kd = rint(__dmul_rn(x, 0.63661977236758134308));
    const int k = (int)kd;
    double r = __fma_rn(-kd, 1.57079632679489655800e+00, x);
    r = __fma_rn(-kd, 6.12323399573676603587e-17, r);
    const double r2 = __dmul_rn(r, r);
    double ps = 1.58969099521155010221e-10;
    ps = __fma_rn(ps, r2, -2.50507602534068634195e-08);
    ps = __fma_rn(ps, r2, 2.75573137070700676789e-06);
    ps = __fma_rn(ps, r2, -1.98412698298579493134e-04);
    ps = __fma_rn(ps, r2, 8.33333333332248946124e-03);
    ps = __fma_rn(ps, r2, -1.66666666666666324348e-01);
    const double sn = __fma_rn(__dmul_rn(r, r2), ps, r);
    double pc = -1.13596475577881948265e-11;
    pc = __fma_rn(pc, r2, 2.08757232129817482790e-09);
    pc = __fma_rn(pc, r2, -2.75573143513906633035e-07);
    pc = __fma_rn(pc, r2, 2.48015872894767294178e-05);
    pc = __fma_rn(pc, r2, -1.38888888888741095749e-03);
    pc = __fma_rn(pc, r2, 4.16666666666666019037e-02);
    const double cs = __fma_rn(__dmul_rn(r2, r2), pc, __fma_rn(-0.5, r2, 1.0));
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

__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

}  // namespace cslam
