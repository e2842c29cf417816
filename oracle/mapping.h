// ORACLE (test infrastructure, not product code): CPU restatement of the feature operations LocalMapping runs either side of LocalBA
// (SURVEY §8(f) rank 4). Pinned against the reference's own compiled code (oracle/_ref/libref.so) in tests/test_oracle_mapping.py.
//   MapPoint::ComputeDistinctiveDescriptors            src/MapPoint.cpp:243-303
//   ORBMatcher::Fuse(KeyFrame*, vector<MapPoint*>&, th) src/ORBMatcher.cpp:1126-1240  (the search; the replace / add bookkeeping stays with the caller)
//   ORBMatcher::SearchForTriangulation                  src/ORBMatcher.cpp:971-1124 + CheckDistEpipolarLine :388-407
//   CamModelGeneral::GetVectorSigma(key, normalRig, sigma) src/CamModelGeneral.cpp:307-335
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "frame_index.h"

namespace orc {

// index of the descriptor with the least median distance to the others (first such index), -1 for an empty list
static inline int distinctive_descriptor(const uint8_t* desc, int N) {
    if (N <= 0) return -1;
    int BestMedian = 0x7fffffff, BestIdx = 0;
    std::vector<int> vDists(N);
    for (int i = 0; i < N; i++) {
        for (int j = 0; j < N; j++) vDists[j] = i == j ? 0 : descriptor_distance(desc + 32 * (size_t)i, desc + 32 * (size_t)j);
        std::sort(vDists.begin(), vDists.end());
        const int median = vDists[(size_t)(0.5 * (N - 1))];
        if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    return BestIdx;
}

// Fuse: for every candidate MapPoint (valid = !isBad && !IsInKeyFrame && depth / viewing-angle tests passed, level = PredictScale, all evaluated by
// the caller on its own objects) the key point of the KeyFrame it would be fused with: bestIdx (-1: none), bestDist (256: none).
static inline void fuse_search(const FrameGrid& g, const KeyPoint* kKF, const uint8_t* dKF, const float* Tcw, const float* scaleFactors, const float* invLevelSigma2,
                               int nMP, const uint8_t* valid, const float* Xw, const int* level, const uint8_t* dMP, float th, int* bestIdxOut, int* bestDistOut) {
    std::vector<int> vIdx;
    for (int m = 0; m < nMP; m++) {
        bestIdxOut[m] = -1; bestDistOut[m] = 256;
        if (!valid[m]) continue;
        float xc[3], u, v;
        fi_transform(Tcw, Xw + 3 * m, xc);
        fi_ray_to_cubemap(xc[0], xc[1], xc[2], g.W, g.H, u, v);   // the face is not looked at: IsInImage decides (a failed in-face test leaves face coordinates)
        if (!(u >= 0.0f && u < (float)(3 * g.W) && v >= 0.0f && v < (float)(3 * g.H))) continue;
        const int nPredictedLevel = level[m];
        const float radius = th * scaleFactors[nPredictedLevel];
        fi_features_in_area(g, kKF, u, v, radius, -1, -1, vIdx);
        int bestDist = 256, bestIdx = -1;
        for (size_t c = 0; c < vIdx.size(); c++) {
            const int idx = vIdx[c];
            const KeyPoint& kp = kKF[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const float ex = u - kp.x, ey = v - kp.y;
            const float e2 = ex * ex + ey * ey;
            if (e2 * invLevelSigma2[kpLevel] > 5.99) continue;
            const int dist = descriptor_distance(dMP + 32 * (size_t)m, dKF + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        bestIdxOut[m] = bestIdx; bestDistOut[m] = bestDist;
    }
}

// cv::norm(Vec3f): double accumulation of the squares, double square root
static inline double mp_norm3(const float* v) { return std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]); }
// cv::Vec3f::dot: float accumulation
static inline float mp_dot3(const float* a, const float* b) { float s = 0; s += a[0] * b[0]; s += a[1] * b[1]; s += a[2] * b[2]; return s; }

// CamModelGeneral::GetVectorSigma(const cv::KeyPoint&, const cv::Vec3f& normalRig, sigmaInPixel = 1)
static inline float vector_sigma(float kx, float ky, const float normalRig[3], int W, int H, float sigmaInPixel = 1.0f) {
    const double fx = W / 2.0, cx = W / 2.0, cy = H / 2.0;
    float nc[3];
    const float x = normalRig[0], y = normalRig[1], z = normalRig[2];
    switch (fi_face_of_pixel_d(kx, ky, W, H)) {   // cvtRigToFaces<float>
        case FI_FRONT: nc[0] = x; nc[1] = y; nc[2] = z; break;
        case FI_LEFT: nc[0] = z; nc[1] = y; nc[2] = -x; break;
        case FI_RIGHT: nc[0] = -z; nc[1] = y; nc[2] = x; break;
        case FI_LOWER: nc[0] = x; nc[1] = -z; nc[2] = y; break;
        case FI_UPPER: nc[0] = x; nc[1] = z; nc[2] = -y; break;
        default: nc[0] = 0; nc[1] = 0; nc[2] = 0; break;
    }
    const float epipolar[3] = {nc[1], -nc[0], 0.0f}, vertical[3] = {nc[0], nc[1], 0.0f};
    const int i = (int)std::floor(kx / W), j = (int)std::floor(ky / H);   // GetPosInFace<float>
    const float u = kx - i * W, v = ky - j * H;
    const float OP[3] = {(float)(u - cx), (float)(v - cy), 0.0f};
    float OO1 = (float)(mp_dot3(OP, epipolar) / mp_norm3(epipolar)); if (OO1 < 0) OO1 = -OO1;
    const float CO1 = (float)std::sqrt(OO1 * OO1 + fx * fx);
    float PO1 = (float)(mp_dot3(OP, vertical) / mp_norm3(vertical)); if (PO1 < 0) PO1 = -PO1;
    const float tan1 = PO1 / CO1;
    const float tan2 = (PO1 + sigmaInPixel) / CO1;
    const float tan3 = (tan2 - tan1) / (1 + tan1 * tan2);
    const float sin_theta = 1.0f / std::sqrt(1.0f / (tan3 * tan3) + 1);
    return sin_theta;
}

// ORBMatcher::CheckDistEpipolarLine
static inline bool check_dist_epipolar_line(const float* ray1, const float* ray2, float k2x, float k2y, int k2oct, const float* E12, const float* levelSigma2, int W, int H) {
    const float a = ray1[0] * E12[0] + ray1[1] * E12[3] + ray1[2] * E12[6];
    const float b = ray1[0] * E12[1] + ray1[1] * E12[4] + ray1[2] * E12[7];
    const float c = ray1[0] * E12[2] + ray1[1] * E12[5] + ray1[2] * E12[8];
    const float num = a * ray2[0] + b * ray2[1] + c * ray2[2];
    const float den = a * a + b * b + c * c;
    if (den == 0) return false;
    const float n3[3] = {a, b, c};
    const float sigma = vector_sigma(k2x, k2y, n3, W, H);
    const float sigmaSquare = sigma * sigma;
    const float dsqr = num * num / (den * sigmaSquare * levelSigma2[k2oct]);
    return dsqr < 3.84;
}

// SearchForTriangulation: node1 / node2 = vocabulary node (levelsup 4) of every feature (the FeatureVectors, a std::map<node, indices ascending>),
// hasMP = the key point already has a MapPoint, Ow1 = camera centre of KF1, Tcw2 = pose of KF2, E12 row-major 3x3.
// match12[i1] = matched feature of KF2 or -1. Returns nmatches. (The reference never sets vbMatched2: several i1 may share an i2.)
static inline int search_for_triangulation(const KeyPoint* k1, const uint8_t* d1, const float* rays1, const uint8_t* hasMP1, const int* node1, int n1, const KeyPoint* k2,
                                           const uint8_t* d2, const float* rays2, const uint8_t* hasMP2, const int* node2, int n2, const float* Ow1, const float* Tcw2,
                                           const float* E12, const float* scaleFactors, const float* levelSigma2, int W, int H, bool checkOri, int* match12) {
    float C2[3], ex, ey;
    fi_transform(Tcw2, Ow1, C2);
    fi_ray_to_cubemap(C2[0], C2[1], C2[2], W, H, ex, ey);
    const int TH_LOW = 50, HISTO_LENGTH = 12;
    const int nBinsAngle = (int)std::ceil(360.0f / HISTO_LENGTH);
    std::vector<std::vector<int> > rotHist(nBinsAngle);
    const float factor = 1.0f / HISTO_LENGTH;
    // the FeatureVectors: node -> ascending indices
    std::vector<std::pair<int, int> > f1(n1), f2(n2);
    for (int i = 0; i < n1; i++) f1[i] = std::make_pair(node1[i], i);
    for (int i = 0; i < n2; i++) f2[i] = std::make_pair(node2[i], i);
    std::sort(f1.begin(), f1.end()); std::sort(f2.begin(), f2.end());
    int nmatches = 0;
    for (int i = 0; i < n1; i++) match12[i] = -1;
    size_t p1 = 0, p2 = 0;
    while (p1 < f1.size() && p2 < f2.size()) {
        const int nd1 = f1[p1].first, nd2 = f2[p2].first;
        if (nd1 < nd2) { while (p1 < f1.size() && f1[p1].first < nd2) p1++; continue; }
        if (nd2 < nd1) { while (p2 < f2.size() && f2[p2].first < nd1) p2++; continue; }
        size_t e1 = p1, e2 = p2;
        while (e1 < f1.size() && f1[e1].first == nd1) e1++;
        while (e2 < f2.size() && f2[e2].first == nd1) e2++;
        for (size_t q1 = p1; q1 < e1; q1++) {
            const int idx1 = f1[q1].second;
            if (hasMP1[idx1]) continue;
            int bestDist = TH_LOW, bestIdx2 = -1;
            for (size_t q2 = p2; q2 < e2; q2++) {
                const int idx2 = f2[q2].second;
                if (hasMP2[idx2]) continue;
                const int dist = descriptor_distance(d1 + 32 * (size_t)idx1, d2 + 32 * (size_t)idx2);
                if (dist > TH_LOW || dist > bestDist) continue;
                const float distex = ex - k2[idx2].x, distey = ey - k2[idx2].y;
                if (distex * distex + distey * distey < 100 * scaleFactors[k2[idx2].octave]) continue;
                if (check_dist_epipolar_line(rays1 + 3 * idx1, rays2 + 3 * idx2, k2[idx2].x, k2[idx2].y, k2[idx2].octave, E12, levelSigma2, W, H)) { bestIdx2 = idx2; bestDist = dist; }
            }
            if (bestIdx2 >= 0) {
                match12[idx1] = bestIdx2; nmatches++;
                if (checkOri) {
                    float rot = k1[idx1].angle - k2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == nBinsAngle) bin = 0;
                    rotHist[bin].push_back(idx1);
                }
            }
        }
        p1 = e1; p2 = e2;
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist.data(), nBinsAngle, ind1, ind2, ind3);
        for (int i = 0; i < nBinsAngle; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { match12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // namespace orc
