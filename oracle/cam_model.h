// ORACLE (test infrastructure, not product code).
// CPU restatement of the reference camera model on the warp / BA path:
//   horner                      include/CamModelGeneral.h:43-50
//   WorldToImg                  include/CamModelGeneral.h:359-374
//   cvtFacesToRig/cvtRigToFaces include/CamModelGeneral.h:388-443
//   FaceInCubemap               include/CamModelGeneral.h:445-470
//   CubemapToFisheye            src/CamModelGeneral.cpp:265-290
//   TransformRaysToTargetFace   src/CamModelGeneral.cpp:228-263
//   CreateUndistortRectifyMap   src/System.cpp:301-324
//   CvtFisheyeToCubeMap_reverseQuery_withInterpolation  src/System.cpp:327-355
#pragma once
#include <cmath>
#include <cstdint>
#include "cvprim.h"

namespace orc {

enum Face { UNKNOWN_FACE = -1, FRONT_FACE = 0, LEFT_FACE = 1, RIGHT_FACE = 2, UPPER_FACE = 3, LOWER_FACE = 4 };

// Plain-C layout shared with the ctypes wrapper (oracle/oracle.py).
struct CamParams {
    double c, d, e, u0, v0;
    double p[5];       // forward polynomial (zero padded, src/System.cpp:67-69)
    double invp[12];   // inverse polynomial (zero padded, src/System.cpp:70-72)
    int Iw, Ih;        // fisheye size
    int faceW, faceH;  // cube face size; fx=fy=cx=cy=W/2 (src/System.cpp:83-84)
    double fov;        // degrees
};

static inline double horner(const double* coeffs, int s, double x) {
    double res = 0.0;
    for (int i = s - 1; i >= 0; i--) res = res * x + coeffs[i];
    return res;
}

static inline void world_to_img(const CamParams& cp, double x, double y, double z, double& u, double& v) {
    double norm = std::sqrt(x * x + y * y);
    if (norm == 0.0) norm = 1e-14;
    const double theta = std::atan(-z / norm);
    const double rho = horner(cp.invp, 12, theta);
    const double uu = x / norm * rho, vv = y / norm * rho;
    u = uu * cp.c + vv * cp.d + cp.u0;
    v = uu * cp.e + vv + cp.v0;
}

template <class T>
static inline void faces_to_rig(T& X, T& Y, T& Z, T x, T y, T z, int face) {
    switch (face) {
        case FRONT_FACE: X = x; Y = y; Z = z; break;
        case LEFT_FACE: X = -z; Y = y; Z = x; break;
        case RIGHT_FACE: X = z; Y = y; Z = -x; break;
        case LOWER_FACE: X = x; Y = z; Z = -y; break;
        case UPPER_FACE: X = x; Y = -z; Z = y; break;
        default: X = 0; Y = 0; Z = 0;
    }
}

template <class T>
static inline void rig_to_faces(T& X, T& Y, T& Z, T x, T y, T z, int face) {
    switch (face) {
        case FRONT_FACE: X = x; Y = y; Z = z; break;
        case LEFT_FACE: X = z; Y = y; Z = -x; break;
        case RIGHT_FACE: X = -z; Y = y; Z = x; break;
        case LOWER_FACE: X = x; Y = -z; Z = y; break;
        case UPPER_FACE: X = x; Y = z; Z = -y; break;
        default: X = 0; Y = 0; Z = 0;
    }
}

// FaceInCubemap(const cv::Point2f&): "double i = pixel.x / mWCubeFace" is a float/int division
// evaluated in fp32 and then widened (include/CamModelGeneral.h:448).
static inline int face_in_cubemap_f(float px, float py, int W, int H) {
    double i = px / (float)W, j = py / (float)H;
    if (i >= 0 && i < 1 && j >= 1 && j < 2) return LEFT_FACE;
    if (i >= 1 && i < 2 && j >= 0 && j < 1) return UPPER_FACE;
    if (i >= 1 && i < 2 && j >= 1 && j < 2) return FRONT_FACE;
    if (i >= 1 && i < 2 && j >= 2 && j < 3) return LOWER_FACE;
    if (i >= 2 && i < 3 && j >= 1 && j < 2) return RIGHT_FACE;
    return UNKNOWN_FACE;
}

static inline void cubemap_to_fisheye(const CamParams& cp, double up, double vp, double& uf, double& vf) {
    float i = (float)up, j = (float)vp;
    uf = -1; vf = -1;
    // FaceInCubemap<float>(i, j): T i = x / mWCubeFace in fp32
    float fi = i / (float)cp.faceW, fj = j / (float)cp.faceH;
    int face = UNKNOWN_FACE;
    if (fi >= 0 && fi < 1 && fj >= 1 && fj < 2) face = LEFT_FACE;
    else if (fi >= 1 && fi < 2 && fj >= 0 && fj < 1) face = UPPER_FACE;
    else if (fi >= 1 && fi < 2 && fj >= 1 && fj < 2) face = FRONT_FACE;
    else if (fi >= 1 && fi < 2 && fj >= 2 && fj < 3) face = LOWER_FACE;
    else if (fi >= 2 && fi < 3 && fj >= 1 && fj < 2) face = RIGHT_FACE;
    if (face == UNKNOWN_FACE) return;
    const double fx = cp.faceW / 2.0, fy = cp.faceH / 2.0, cx = fx, cy = fy;
    i = i - static_cast<int>(i / cp.faceW) * cp.faceW;
    j = j - static_cast<int>(j / cp.faceH) * cp.faceH;
    double z = 1.0, x = (i - cx) * z / fx, y = (j - cy) * z / fy;
    double X, Y, Z;
    faces_to_rig<double>(X, Y, Z, x, y, z, face);
    world_to_img(cp, X, Y, Z, uf, vf);
    if (uf < 0 || uf >= cp.Iw || vf < 0 || vf >= cp.Ih) { uf = -1; vf = -1; }
}

// map1/map2: (3*faceH) x (3*faceW) float32, zero where the canvas pixel has no fisheye preimage.
static inline void build_maps(const CamParams& cp, float* map1, float* map2) {
    const int W3 = cp.faceW * 3, H3 = cp.faceH * 3;
    for (int y = 0; y < H3; y++)
        for (int x = 0; x < W3; x++) {
            double u, v;
            map1[(size_t)y * W3 + x] = 0.f;
            map2[(size_t)y * W3 + x] = 0.f;
            cubemap_to_fisheye(cp, (double)x, (double)y, u, v);
            if (u < 0 || v < 0 || u >= cp.Iw || v >= cp.Ih) continue;
            map1[(size_t)y * W3 + x] = (float)u;
            map2[(size_t)y * W3 + x] = (float)v;
        }
}

// canvas: (3H x 3W) u8; only the 5 face tiles are written (corner tiles keep their content).
static inline void warp_fisheye_to_cubemap(const CamParams& cp, const uint8_t* fisheye, int fstride,
                                           const float* map1, const float* map2, uint8_t* canvas, int cstride) {
    const int W = cp.faceW, H = cp.faceH, W3 = 3 * W;
    const int tiles[5][2] = {{1, 1}, {0, 1}, {2, 1}, {1, 0}, {1, 2}};  // front,left,right,upper,lower (col,row)
    for (int t = 0; t < 5; t++) {
        int x0 = tiles[t][0] * W, y0 = tiles[t][1] * H;
        remap_bilinear(fisheye, cp.Iw, cp.Ih, fstride, map1 + (size_t)y0 * W3 + x0, map2 + (size_t)y0 * W3 + x0, W3,
                       canvas + (size_t)y0 * cstride + x0, W, H, cstride);
    }
}

// TransformRaysToTargetFace (src/CamModelGeneral.cpp:228-263): rigPt is cv::Vec3f; "up = _x * fx / _z + cx"
// is float*double/float+double evaluated in fp64 and stored to float.
static inline void rays_to_target_face(double fx, double fy, double cx, double cy, float rx, float ry, float rz,
                                       int face, float& up, float& vp) {
    float lx, ly, lz;
    if (face < 0 || face > 4) { up = -1.0f; vp = -1.0f; return; }
    rig_to_faces<float>(lx, ly, lz, rx, ry, rz, face);
    up = (float)((double)lx * fx / (double)lz + cx);
    vp = (float)((double)ly * fy / (double)lz + cy);
}

}  // namespace orc
