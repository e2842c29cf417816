// ORACLE (test infrastructure, not product code): CPU restatement of the per-frame indexing the reference runs right after extraction and
// of the windowed feature lookup + projection matcher built on it. Pinned against the reference's own compiled code (oracle/_ref/libref.so)
// in tests/test_oracle_frame_index.py.
//   Frame::ComputeKeyPointRays      src/Frame.cpp:746-760  -> CamModelGeneral::TransformCubemapToRays  include/CamModelGeneral.h:494-513
//   Frame::AssignFeaturesToGrid     src/Frame.cpp:158-176  -> Frame::PosInGrid :728-744   (5 x 50 x 50 cells, include/Frame.h:43-45)
//   Frame::GetFeaturesInArea        src/Frame.cpp:251-716  (cube-face wrap-around cases) + AddCells :37-72
//   ORBMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)   src/ORBMatcher.cpp:51-128
//   ORBMatcher::SearchByProjection(Frame&, const Frame&, th, bMono)        src/ORBMatcher.cpp:130-251
//   CamModelGeneral::TransformRaysToCubemap   src/CamModelGeneral.cpp:95-154 (+ FaceInCubemap(x,y,z) include/CamModelGeneral.h:472-492)
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include "orb_extractor.h"
#include "orb_matcher.h"

namespace orc {

enum { FI_FRONT = 0, FI_LEFT = 1, FI_RIGHT = 2, FI_UPPER = 3, FI_LOWER = 4, FI_UNKNOWN = -1 };
static const int FI_G = 50;   // CUBEFACE_GRID_COLS == CUBEFACE_GRID_ROWS

// FaceInCubemap(const cv::Point2f&): the quotient is a double (float / int promoted)
static inline int fi_face_of_pixel_d(float px, float py, int W, int H) {
    const double i = px / W, j = py / H;   // float / int -> float, then widened: `double i = pixel.x / mWCubeFace`
    if (i >= 0 && i < 1 && j >= 1 && j < 2) return FI_LEFT;
    if (i >= 1 && i < 2 && j >= 0 && j < 1) return FI_UPPER;
    if (i >= 1 && i < 2 && j >= 1 && j < 2) return FI_FRONT;
    if (i >= 1 && i < 2 && j >= 2 && j < 3) return FI_LOWER;
    if (i >= 2 && i < 3 && j >= 1 && j < 2) return FI_RIGHT;
    return FI_UNKNOWN;
}
// FaceInCubemap<float>(x, y)
static inline int fi_face_of_pixel_f(float x, float y, int W, int H) {
    const float i = x / W, j = y / H;
    if (i >= 0 && i < 1 && j >= 1 && j < 2) return FI_LEFT;
    if (i >= 1 && i < 2 && j >= 0 && j < 1) return FI_UPPER;
    if (i >= 1 && i < 2 && j >= 1 && j < 2) return FI_FRONT;
    if (i >= 1 && i < 2 && j >= 2 && j < 3) return FI_LOWER;
    if (i >= 2 && i < 3 && j >= 1 && j < 2) return FI_RIGHT;
    return FI_UNKNOWN;
}

// TransformCubemapToRays: unit bearing vector of a canvas pixel (fx = fy = cx = cy = W/2). Returns the face (or FI_UNKNOWN).
static inline int fi_pixel_to_ray(float px, float py, int W, int H, float ray[3]) {
    const int face = fi_face_of_pixel_d(px, py, W, H);
    if (face == FI_UNKNOWN) return face;
    const double fx = W / 2.0, fy = H / 2.0, cx = W / 2.0, cy = H / 2.0;
    double x, y, z = 1.0;
    double i = px, j = py;
    i = i - static_cast<int>(i / W) * W; j = j - static_cast<int>(j / H) * H;
    x = (i - cx) * z / fx;
    y = (j - cy) * z / fy;
    const float lx = (float)x, ly = (float)y, lz = (float)z;   // cv::Vec3f(x, y, z)
    float p[3];
    switch (face) {   // cvtFacesToRig<float>
        case FI_FRONT: p[0] = lx; p[1] = ly; p[2] = lz; break;
        case FI_LEFT: p[0] = -lz; p[1] = ly; p[2] = lx; break;
        case FI_RIGHT: p[0] = lz; p[1] = ly; p[2] = -lx; break;
        case FI_LOWER: p[0] = lx; p[1] = lz; p[2] = -ly; break;
        default: p[0] = lx; p[1] = -lz; p[2] = ly; break;
    }
    const double normal = std::sqrt((double)p[0] * p[0] + (double)p[1] * p[1] + (double)p[2] * p[2]);   // cv::norm(Vec3f): double accumulation
    const double s = normal > 0 ? 1. / normal : 0.;
    for (int k = 0; k < 3; k++) ray[k] = (float)(p[k] * s);   // Vec3f * double: saturate_cast<float>(double product)
    return face;
}

// the per-frame grid: cells[(face*50 + col)*50 + row] = feature indices in ascending order (AssignFeaturesToGrid pushes i = 0..N-1)
struct FrameGrid {
    int W = 0, H = 0; float inv = 0, len = 0;   // mfGridElementLengthInv / mfGridElementLength for a 3W-wide canvas
    std::vector<std::vector<int> > cells;
    void build(const KeyPoint* kps, int n, int faceW, int faceH) {
        W = faceW; H = faceH;
        const float mnMinX = 0.0f, mnMaxX = (float)(3 * W);
        inv = static_cast<float>(3 * FI_G) / static_cast<float>(mnMaxX - mnMinX);
        len = static_cast<float>(mnMaxX - mnMinX) / static_cast<float>(3 * FI_G);
        cells.assign(5 * FI_G * FI_G, std::vector<int>());
        for (int i = 0; i < n; i++) {
            const int face = fi_face_of_pixel_d(kps[i].x, kps[i].y, W, H);
            if (face == FI_UNKNOWN) continue;
            int posX = static_cast<int>((kps[i].x - mnMinX) * inv), posY = static_cast<int>((kps[i].y - 0.0f) * inv);
            posX %= FI_G; posY %= FI_G;
            cells[(face * FI_G + posX) * FI_G + posY].push_back(i);
        }
    }
};

struct FiRect { int face, x0, x1, y0, y1; };
// the cell rectangles Frame::GetFeaturesInArea visits, in visiting order (src/Frame.cpp:288-714). Returns their number (0..3).
static inline int fi_area_rects(float x, float y, float r, int W, int H, float inv, FiRect out[3]) {
    const int face = fi_face_of_pixel_f(x, y, W, H);
    if (face == FI_UNKNOWN) return 0;
    const int nCornerX = static_cast<int>(x) / W * W, nCornerY = static_cast<int>(y) / H * H;
    const float xIn = x - nCornerX, yIn = y - nCornerY;
    const float xs = xIn - r, xe = xIn + r, ys = yIn - r, ye = yIn + r;
    const bool xU = xs < 0, xO = xe > W - 1, yU = ys < 0, yO = ye > H - 1;
    const bool xIn_ = !xO && !xU, yIn_ = !yO && !yU;
    const int G = FI_G;
    auto F = [&](float v) { return (int)std::floor(v * inv); };
    const int a = F(xs), b = F(xe), c = F(ys), d = F(ye);                      // in-face cell bounds
    const int bW = F(xe - W), dH = F(ye - H), aW = F(xs + W), cH = F(ys + H), aN = F(-xs), cN = F(-ys);
    int n = 0;
    auto add = [&](int f, int x0, int x1, int y0, int y1) { out[n].face = f; out[n].x0 = x0; out[n].x1 = x1; out[n].y0 = y0; out[n].y1 = y1; n++; };
    if (xIn_ && yIn_) { add(face, a, b, c, d); return n; }
    if (xIn_ && !yIn_) {
        switch (face) {
            case FI_FRONT: if (yO) { add(FI_FRONT, a, b, c, G - 1); add(FI_LOWER, a, b, 0, dH); } else { add(FI_UPPER, a, b, cH, G - 1); add(FI_FRONT, a, b, 0, d); } break;
            case FI_LEFT: if (yO) { add(FI_LEFT, a, b, c, G - 1); add(FI_LOWER, 0, dH, G - b - 1, G - a - 1); } else { add(FI_UPPER, 0, cN, a, b); add(FI_LEFT, a, b, 0, d); } break;
            case FI_RIGHT: if (yO) { add(FI_RIGHT, a, b, c, G - 1); add(FI_LOWER, G - dH - 1, G - 1, a, b); } else { add(FI_UPPER, cH, G - 1, G - b - 1, G - a - 1); add(FI_RIGHT, a, b, 0, d); } break;
            case FI_UPPER: if (yO) { add(FI_UPPER, a, b, c, G - 1); add(FI_FRONT, a, b, 0, dH); } else { add(FI_LOWER, a, b, 0, d); } break;
            default: if (yO) { add(FI_LOWER, a, b, c, G - 1); } else { add(FI_FRONT, a, b, cH, G - 1); add(FI_LOWER, a, b, 0, d); } break;
        }
        return n;
    }
    if (!xIn_ && yIn_) {
        switch (face) {
            case FI_FRONT: if (xO) { add(FI_FRONT, a, G - 1, c, d); add(FI_RIGHT, 0, bW, c, d); } else { add(FI_LEFT, aW, G - 1, c, d); add(FI_FRONT, 0, b, c, d); } break;
            case FI_LEFT: if (xO) { add(FI_FRONT, 0, bW, c, d); add(FI_LEFT, a, G - 1, c, d); } else { add(FI_LEFT, 0, b, c, d); } break;
            case FI_RIGHT: if (xO) { add(FI_RIGHT, a, G - 1, c, d); } else { add(FI_FRONT, aW, G - 1, c, d); add(FI_RIGHT, 0, b, c, d); } break;
            case FI_UPPER: if (xO) { add(FI_UPPER, a, G - 1, c, d); add(FI_RIGHT, G - d - 1, G - c - 1, 0, bW); } else { add(FI_LEFT, c, d, 0, aN); add(FI_UPPER, 0, b, c, d); } break;
            default: if (xO) { add(FI_LOWER, a, G - 1, c, d); add(FI_RIGHT, c, d, G - bW - 1, G); } else { add(FI_LEFT, G - d - 1, G - c - 1, aW, G - 1); add(FI_LOWER, 0, b, c, d); } break;
        }
        return n;
    }
    // neither direction stays inside the face: if / else-if chain in the reference's order
    const int sub = (xO && yO) ? 0 : (xU && yO) ? 1 : (xO && yU) ? 2 : (xU && yU) ? 3 : -1;
    if (sub < 0) return 0;
    switch (face) {
        case FI_FRONT:
            if (sub == 0) { add(FI_FRONT, a, G - 1, c, G - 1); add(FI_RIGHT, 0, bW, c, G - 1); add(FI_LOWER, a, G - 1, 0, dH); }
            else if (sub == 1) { add(FI_FRONT, 0, b, c, G - 1); add(FI_LEFT, aW, G - 1, c, G - 1); add(FI_LOWER, 0, b, 0, dH); }
            else if (sub == 2) { add(FI_FRONT, a, G - 1, 0, d); add(FI_RIGHT, 0, bW, 0, d); add(FI_UPPER, a, G - 1, G - cH - 1, G - 1); }
            else { add(FI_FRONT, 0, b, 0, d); add(FI_LEFT, G - aW - 1, G - 1, 0, d); add(FI_UPPER, 0, b, G - cH - 1, G - 1); }
            break;
        case FI_LEFT:
            if (sub == 0) { add(FI_LEFT, a, G - 1, c, G - 1); add(FI_FRONT, 0, bW, c, G - 1); add(FI_LOWER, 0, dH, 0, G - a - 1); }
            else if (sub == 1) { add(FI_LEFT, 0, b, c, G - 1); add(FI_LOWER, 0, dH, G - b - 1, G - 1); }
            else if (sub == 2) { add(FI_LEFT, a, G - 1, 0, d); add(FI_FRONT, 0, bW, 0, d); add(FI_UPPER, 0, cN, a, G - 1); }
            else { add(FI_LEFT, 0, b, 0, d); add(FI_UPPER, 0, cN, 0, d); }
            break;
        case FI_RIGHT:
            if (sub == 0) { add(FI_RIGHT, a, G - 1, c, G - 1); add(FI_LOWER, G - dH - 1, G - 1, a, G - 1); }
            else if (sub == 1) { add(FI_RIGHT, 0, b, c, G - 1); add(FI_FRONT, G - aN - 1, G - 1, c, G - 1); add(FI_LOWER, G - dH - 1, G - 1, 0, b); }
            else if (sub == 2) { add(FI_RIGHT, a, G - 1, 0, d); add(FI_UPPER, G - cN - 1, G - 1, 0, G - a - 1); }
            else { add(FI_RIGHT, 0, b, 0, d); add(FI_FRONT, G - aN - 1, G - 1, 0, d); add(FI_UPPER, G - cN - 1, G - 1, G - b - 1, G - 1); }
            break;
        case FI_UPPER:
            if (sub == 0) { add(FI_UPPER, a, G - 1, c, G - 1); add(FI_RIGHT, 0, G - c - 1, 0, bW); add(FI_FRONT, a, G - 1, 0, dH); }
            else if (sub == 1) { add(FI_UPPER, 0, b, c, G - 1); add(FI_LEFT, c, G - 1, 0, aN); add(FI_FRONT, 0, b, 0, dH); }
            else if (sub == 2) { add(FI_UPPER, a, G - 1, 0, d); add(FI_RIGHT, G - d - 1, G - 1, 0, d); }
            else { add(FI_UPPER, 0, b, 0, d); add(FI_LEFT, 0, d, 0, aN); }
            break;
        default:
            if (sub == 0) { add(FI_LOWER, a, G - 1, c, G - 1); add(FI_RIGHT, a, G - 1, G - bW - 1, G - 1); }
            else if (sub == 1) { add(FI_LOWER, 0, b, c, G - 1); add(FI_LEFT, 0, G - c - 1, G - aW - 1, G - 1); }
            else if (sub == 2) { add(FI_LOWER, a, G - 1, 0, d); add(FI_RIGHT, 0, d, G - bW - 1, G); add(FI_FRONT, a, G - 1, G - cN - 1, G - 1); }
            else { add(FI_LOWER, 0, b, 0, d); add(FI_LEFT, G - aN - 1, G - 1, G - aN - 1, G); add(FI_FRONT, 0, b, G - cN - 1, G - 1); }
            break;
    }
    return n;
}

// Frame::GetFeaturesInArea: indices in the reference's visiting order
static inline void fi_features_in_area(const FrameGrid& g, const KeyPoint* kps, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) {
    out.clear();
    FiRect rc[3];
    const int nr = fi_area_rects(x, y, r, g.W, g.H, g.inv, rc);
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int k = 0; k < nr; k++) {
        const int x0 = std::max(0, rc[k].x0), x1 = std::min(FI_G - 1, rc[k].x1), y0 = std::max(0, rc[k].y0), y1 = std::min(FI_G - 1, rc[k].y1);
        for (int ix = x0; ix <= x1; ix++)
            for (int iy = y0; iy <= y1; iy++) {
                const std::vector<int>& cell = g.cells[(rc[k].face * FI_G + ix) * FI_G + iy];
                for (size_t j = 0; j < cell.size(); j++) {
                    const KeyPoint& kp = kps[cell[j]];
                    if (bCheckLevels) { if (kp.octave < minLevel) continue; if (maxLevel >= 0 && kp.octave > maxLevel) continue; }
                    const float distx = kp.x - x, disty = kp.y - y;
                    if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(cell[j]);
                }
            }
    }
}

// CamModelGeneral::TransformRaysToCubemap(up, vp, x, y, z): ordered face tests, in-face pinhole projection evaluated in double
// (`_x * fx / _z + cx` with fx, cx doubles) and stored as float, in-face bounds check, tile offset added in float.
static inline int fi_ray_to_cubemap(float x, float y, float z, int W, int H, float& up, float& vp) {
    const double fx = W / 2.0, fy = H / 2.0, cx = W / 2.0, cy = H / 2.0;
    int face; float lx, ly, lz;
    if (z > 0 && x / z <= 1 && x / z >= -1 && y / z <= 1 && y / z >= -1) { face = FI_FRONT; lx = x; ly = y; lz = z; }
    else if (x > 0 && y / x <= 1 && y / x >= -1 && z / x <= 1 && z / x >= -1) { face = FI_RIGHT; lx = -z; ly = y; lz = x; }
    else if (x < 0 && y / (-x) <= 1 && y / (-x) >= -1 && z / (-x) <= 1 && z / (-x) >= -1) { face = FI_LEFT; lx = z; ly = y; lz = -x; }
    else if (y > 0 && x / y <= 1 && x / y >= -1 && z / y <= 1 && z / y >= -1) { face = FI_LOWER; lx = x; ly = -z; lz = y; }
    else if (y < 0 && x / (-y) <= 1 && x / (-y) >= -1 && z / (-y) <= 1 && z / (-y) >= -1) { face = FI_UPPER; lx = x; ly = z; lz = -y; }
    else { up = -1; vp = -1; return FI_UNKNOWN; }
    up = (float)(lx * fx / lz + cx);
    vp = (float)(ly * fy / lz + cy);
    if (up < 0 || up >= W || vp < 0 || vp >= H) return FI_UNKNOWN;
    switch (face) {
        case FI_FRONT: up += W; vp += H; break;
        case FI_RIGHT: up += 2 * W; vp += H; break;
        case FI_LEFT: vp += H; break;
        case FI_LOWER: up += W; vp += 2 * H; break;
        default: up += W; break;
    }
    return face;
}

// `Rcw*x3Dw+tcw` on CV_32F Mats: OpenCV's small-matrix gemm accumulates the 3-term inner product in float, left to right, then
// (float)((double)t*1.0 + (double)c*1.0)  (pinned against cv2.gemm: tests/test_oracle_ref.py::test_shim_gemm_matches_cv2)
static inline void fi_transform(const float* Tcw, const float* X, float* out) {
    for (int i = 0; i < 3; i++) {
        float t = Tcw[4 * i] * X[0];
        t = t + Tcw[4 * i + 1] * X[1];
        t = t + Tcw[4 * i + 2] * X[2];
        out[i] = (float)((double)t + (double)Tcw[4 * i + 3]);
    }
}

// ORBMatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono = true)   src/ORBMatcher.cpp:130-251
// last side: per feature i: hasMP (a MapPoint that is not an outlier), its world position, its descriptor, octave / angle of the key point.
// curTaken[i2] != 0: CurrentFrame.mvpMapPoints[i2] already holds a MapPoint with Observations() > 0 before the call.
// mpObs[i]: Observations() of LastFrame's MapPoint i (> 0 makes the slot it is assigned to unavailable for later points).
// matchCur[i2] = index i of the LastFrame feature whose MapPoint was assigned, or -1. Returns nmatches.
static inline int fi_search_by_projection_last(const FrameGrid& g, const KeyPoint* kCur, const uint8_t* dCur, int nCur, const float* TcwCur, const float* scaleFactors,
                                               const KeyPoint* kLast, int nLast, const uint8_t* hasMP, const float* Xw, const uint8_t* dMP, const int* mpObs,
                                               const uint8_t* curTaken, float cosFovTh, float th, bool checkOri, int* matchCur) {
    int nmatches = 0;
    const int nBins = 30;   // ceil(360 / HISTO_LENGTH)
    std::vector<std::vector<int> > rotHist(nBins);
    const float factor = 1.0f / 12;
    std::vector<int> slotObs(nCur, 0);   // Observations() of whatever sits in CurrentFrame.mvpMapPoints[i2]; 0 = empty or obs-less
    for (int i = 0; i < nCur; i++) { matchCur[i] = -1; slotObs[i] = curTaken && curTaken[i] ? 1 : 0; }
    std::vector<int> vIdx;
    for (int i = 0; i < nLast; i++) {
        if (!hasMP[i]) continue;
        float xc[3];
        fi_transform(TcwCur, Xw + 3 * i, xc);
        if (xc[2] < cosFovTh) continue;
        float u, v;
        if (fi_ray_to_cubemap(xc[0], xc[1], xc[2], g.W, g.H, u, v) == FI_UNKNOWN) continue;
        const int oct = kLast[i].octave;
        const float radius = th * scaleFactors[oct];
        fi_features_in_area(g, kCur, u, v, radius, oct - 1, oct + 1, vIdx);
        if (vIdx.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIdx) {
            if (slotObs[i2] > 0) continue;
            const int dist = descriptor_distance(dMP + 32 * (size_t)i, dCur + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= 100) {   // TH_HIGH
            // (an earlier assignment whose MapPoint has no observations is simply overwritten; the reference does not decrement nmatches)
            matchCur[bestIdx2] = i; slotObs[bestIdx2] = mpObs[i];
            nmatches++;
            if (checkOri) {
                float rot = kLast[i].angle - kCur[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == nBins) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist.data(), nBins, ind1, ind2, ind3);
        for (int b = 0; b < nBins; b++)
            if (b != ind1 && b != ind2 && b != ind3)
                for (int j : rotHist[b]) { matchCur[j] = -1; nmatches--; }
    }
    return nmatches;
}

// ORBMatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th)   src/ORBMatcher.cpp:51-128
// per MapPoint: inView (mbTrackInView && !isBad), mTrackProjX/Y, mnTrackScaleLevel, mTrackViewCos, descriptor, Observations().
static inline int fi_search_by_projection_local(const FrameGrid& g, const KeyPoint* kF, const uint8_t* dF, int nF, const float* scaleFactors, int nMP, const uint8_t* inView,
                                                const float* projXY, const int* level, const float* viewCos, const uint8_t* dMP, const int* mpObs, const uint8_t* fTaken,
                                                float th, float nnratio, int* matchF) {
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> slotObs(nF, 0);
    for (int i = 0; i < nF; i++) { matchF[i] = -1; slotObs[i] = fTaken && fTaken[i] ? 1 : 0; }
    std::vector<int> vIdx;
    for (int m = 0; m < nMP; m++) {
        if (!inView[m]) continue;
        const int lvl = level[m];
        float r = viewCos[m] > 0.998 ? 2.5 : 4.0;   // RadiusByViewingCos (:380-386): float compared with a double literal
        if (bFactor) r *= th;
        fi_features_in_area(g, kF, projXY[2 * m], projXY[2 * m + 1], r * scaleFactors[lvl], lvl - 1, lvl, vIdx);
        if (vIdx.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIdx) {
            if (slotObs[idx] > 0) continue;
            const int dist = descriptor_distance(dMP + 32 * (size_t)m, dF + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kF[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = kF[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= 100) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            matchF[bestIdx] = m; slotObs[bestIdx] = mpObs[m];
            nmatches++;
        }
    }
    return nmatches;
}

}  // namespace orc
