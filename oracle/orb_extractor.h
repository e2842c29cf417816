// ORACLE (test infrastructure, not product code).
// CPU restatement of ORBextractor (src/ORBExtractor.cpp, whole file) with the OpenCV primitives
// replaced by the cv2-4.13-pinned models of cvprim.h. Two places where the reference binary is not
// self-deterministic are DEFINED here (and documented in DESIGN.md):
//   * quadtree finishing phase sorts by (count, heap pointer) (src/ORBExtractor.cpp:658): the oracle
//     uses (count, creation sequence number) - "later created = larger address".
//   * cosf/sinf (src/ORBExtractor.cpp:83-84): det_sincosf (cvprim.h).
//   * a level whose node list can no longer grow while it is shorter than N/100 makes the reference loop forever
//     (src/ORBExtractor.cpp:643); here the loop ends (see DistributeOctTree).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <list>
#include <utility>
#include <vector>
#include "cam_model.h"
#include "cvprim.h"

namespace orc {

struct KeyPoint {  // field order of cv::KeyPoint
    float x, y, size, angle, response;
    int octave, class_id;
};

static const int PATCH_SIZE = 31;
static const int HALF_PATCH_SIZE = 15;
static const int EDGE_THRESHOLD = 19;

static const signed char kBriefPattern[1024] = {
#include "brief_pattern.inc"
};

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> px;
    void create(int w_, int h_) { w = w_; h = h_; px.assign((size_t)w * h, 0); }
    uint8_t* row(int y) { return px.data() + (size_t)y * w; }
    const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
};

struct ExtractorNode {
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::vector<KeyPoint> vKeys;
    std::list<ExtractorNode>::iterator lit;
    bool bNoMore = false;
    int seq = 0;  // creation order (the oracle's stand-in for the heap address tie-break)

    // src/ORBExtractor.cpp:453-509
    void DivideNode(ExtractorNode& n1, ExtractorNode& n2, ExtractorNode& n3, ExtractorNode& n4) const {
        const int halfX = (int)std::ceil(static_cast<float>(URx - ULx) / 2);
        const int halfY = (int)std::ceil(static_cast<float>(BRy - ULy) / 2);
        n1.ULx = ULx; n1.ULy = ULy; n1.URx = ULx + halfX; n1.URy = ULy;
        n1.BLx = ULx; n1.BLy = ULy + halfY; n1.BRx = ULx + halfX; n1.BRy = ULy + halfY;
        n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = URx; n2.URy = URy;
        n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = URx; n2.BRy = ULy + halfY;
        n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
        n3.BLx = BLx; n3.BLy = BLy; n3.BRx = n1.BRx; n3.BRy = BLy;
        n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
        n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = BRx; n4.BRy = BRy;
        for (size_t i = 0; i < vKeys.size(); i++) {
            const KeyPoint& kp = vKeys[i];
            if (kp.x < n1.URx) {
                if (kp.y < n1.BRy) n1.vKeys.push_back(kp); else n3.vKeys.push_back(kp);
            } else if (kp.y < n1.BRy) n2.vKeys.push_back(kp);
            else n4.vKeys.push_back(kp);
        }
        if (n1.vKeys.size() == 1) n1.bNoMore = true;
        if (n2.vKeys.size() == 1) n2.bNoMore = true;
        if (n3.vKeys.size() == 1) n3.bNoMore = true;
        if (n4.vKeys.size() == 1) n4.bNoMore = true;
    }
};

class ORBextractor {
public:
    int nfeatures, nlevels, iniThFAST, minThFAST;
    float scaleFactor;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel, umax;
    int faceW, faceH;  // the camera singleton's cube face size (CamModelGeneral::GetCamera())

    // stage outputs kept for stage-wise parity tests of the CUDA path
    std::vector<Image> mvImagePyramid;                    // apron-less levels
    std::vector<std::vector<KeyPoint>> mvCandidates;      // FAST grid output per level (minBorder-relative)
    std::vector<std::vector<KeyPoint>> mvDistributed;     // after DistributeOctTree + border + angle, per level
    std::vector<Image> mvBlurred;                         // blurred levels (only levels that had keypoints)

    // src/ORBExtractor.cpp:381-442
    ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST, int _faceW, int _faceH)
        : nfeatures(_nfeatures), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
          scaleFactor(_scaleFactor), faceW(_faceW), faceH(_faceH) {
        mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) {
            mvScaleFactor[i] = mvScaleFactor[i - 1] * scaleFactor;
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        for (int i = 0; i < nlevels; i++) {
            mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
            mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
        }
        mnFeaturesPerLevel.resize(nlevels);
        float factor = 1.0f / scaleFactor;
        float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
        int sumFeatures = 0;
        for (int level = 0; level < nlevels - 1; level++) {
            mnFeaturesPerLevel[level] = cv_round(nDesired);
            sumFeatures += mnFeaturesPerLevel[level];
            nDesired *= factor;
        }
        mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sumFeatures, 0);

        umax.resize(HALF_PATCH_SIZE + 1);
        int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
        int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    // src/ORBExtractor.cpp:928-953. The 19-px REFLECT_101 apron the reference adds is never read by any
    // later stage (FAST ROI starts at 16, IC_Angle/BRIEF centres are >= 19 from the edge, resize and blur
    // operate on the apron-less ROI / clone), so the oracle stores apron-less levels.
    void ComputePyramid(const uint8_t* image, int cols, int rows, int stride) {
        mvImagePyramid.resize(nlevels);
        for (int level = 0; level < nlevels; ++level) {
            float scale = mvInvScaleFactor[level];
            int sw = cv_round((float)cols * scale), sh = cv_round((float)rows * scale);
            Image& L = mvImagePyramid[level];
            L.create(sw, sh);
            if (level != 0) {
                const Image& P = mvImagePyramid[level - 1];
                resize_linear(P.px.data(), P.w, P.h, P.w, L.px.data(), sw, sh, sw);
            } else {
                for (int y = 0; y < rows; y++) std::memcpy(L.row(y), image + (size_t)y * stride, cols);
            }
        }
    }

    // src/ORBExtractor.cpp:48-75
    float IC_Angle(const Image& image, float ptx, float pty) const {
        int m_01 = 0, m_10 = 0;
        const int step = image.w;
        const uint8_t* center = image.row(cv_round(pty)) + cv_round(ptx);
        for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
        for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
            int v_sum = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int val_plus = center[u + v * step], val_minus = center[u - v * step];
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        return fast_atan2((float)m_01, (float)m_10);
    }

    // src/ORBExtractor.cpp:511-737
    std::vector<KeyPoint> DistributeOctTree(const std::vector<KeyPoint>& vToDistributeKeys, int minX, int maxX,
                                            int minY, int maxY, int N) {
        int seqCounter = 0;
        const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
        const float hX = static_cast<float>(maxX - minX) / nIni;
        std::list<ExtractorNode> lNodes;
        std::vector<ExtractorNode*> vpIniNodes(nIni);
        for (int i = 0; i < nIni; i++) {
            ExtractorNode ni;
            ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
            ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
            ni.BLx = ni.ULx; ni.BLy = maxY - minY;
            ni.BRx = ni.URx; ni.BRy = maxY - minY;
            ni.seq = seqCounter++;
            lNodes.push_back(ni);
            vpIniNodes[i] = &lNodes.back();
        }
        for (size_t i = 0; i < vToDistributeKeys.size(); i++) {
            const KeyPoint& kp = vToDistributeKeys[i];
            vpIniNodes[(int)(kp.x / hX)]->vKeys.push_back(kp);
        }
        auto lit = lNodes.begin();
        while (lit != lNodes.end()) {
            if (lit->vKeys.size() == 1) { lit->bNoMore = true; lit++; }
            else if (lit->vKeys.empty()) lit = lNodes.erase(lit);
            else lit++;
        }
        bool bFinish = false;
        typedef std::pair<int, ExtractorNode*> SizeNode;
        auto bySizeThenSeq = [](const SizeNode& a, const SizeNode& b) {
            if (a.first != b.first) return a.first < b.first;
            return a.second->seq < b.second->seq;
        };
        std::vector<SizeNode> vSizeAndPointerToNode;
        auto pushChild = [&](ExtractorNode& n, std::vector<SizeNode>& rec, int* nToExpand) {
            if (n.vKeys.size() > 0) {
                n.seq = seqCounter++;
                lNodes.push_front(n);
                if (n.vKeys.size() > 1) {
                    if (nToExpand) (*nToExpand)++;
                    rec.push_back(std::make_pair((int)n.vKeys.size(), &lNodes.front()));
                    lNodes.front().lit = lNodes.begin();
                }
            }
        };
        while (!bFinish) {
            int prevSize = (int)lNodes.size();
            lit = lNodes.begin();
            int nToExpand = 0, nDivided = 0;
            vSizeAndPointerToNode.clear();
            while (lit != lNodes.end()) {
                if (lit->bNoMore) { lit++; continue; }
                ExtractorNode n1, n2, n3, n4;
                nDivided++;
                lit->DivideNode(n1, n2, n3, n4);
                pushChild(n1, vSizeAndPointerToNode, &nToExpand);
                pushChild(n2, vSizeAndPointerToNode, &nToExpand);
                pushChild(n3, vSizeAndPointerToNode, &nToExpand);
                pushChild(n4, vSizeAndPointerToNode, &nToExpand);
                lit = lNodes.erase(lit);
            }
            if ((int)lNodes.size() >= N || ((int)lNodes.size() == prevSize && (int)lNodes.size() >= N / 100)) {
                bFinish = true;
            } else if (nDivided == 0) {
                // DEFINED BEHAVIOUR: with fewer than N/100 nodes and nothing left to divide (e.g. a level without any
                // FAST corner) the reference's modified stop rule (src/ORBExtractor.cpp:643) never becomes true and
                // the loop spins forever. The oracle (and the CUDA path) stop here with the current nodes.
                bFinish = true;
            } else if (((int)lNodes.size() + nToExpand * 3) > N) {
                while (!bFinish) {
                    prevSize = (int)lNodes.size();
                    std::vector<SizeNode> vPrev = vSizeAndPointerToNode;
                    vSizeAndPointerToNode.clear();
                    std::sort(vPrev.begin(), vPrev.end(), bySizeThenSeq);
                    for (int j = (int)vPrev.size() - 1; j >= 0; j--) {
                        ExtractorNode n1, n2, n3, n4;
                        vPrev[j].second->DivideNode(n1, n2, n3, n4);
                        pushChild(n1, vSizeAndPointerToNode, nullptr);
                        pushChild(n2, vSizeAndPointerToNode, nullptr);
                        pushChild(n3, vSizeAndPointerToNode, nullptr);
                        pushChild(n4, vSizeAndPointerToNode, nullptr);
                        lNodes.erase(vPrev[j].second->lit);
                        if ((int)lNodes.size() >= N) break;
                    }
                    if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
                }
            }
        }
        std::vector<KeyPoint> vResultKeys;
        vResultKeys.reserve(nfeatures);
        for (auto it = lNodes.begin(); it != lNodes.end(); it++) {
            std::vector<KeyPoint>& vNodeKeys = it->vKeys;
            KeyPoint* pKP = &vNodeKeys[0];
            float maxResponse = pKP->response;
            for (size_t k = 1; k < vNodeKeys.size(); k++)
                if (vNodeKeys[k].response > maxResponse) { pKP = &vNodeKeys[k]; maxResponse = vNodeKeys[k].response; }
            vResultKeys.push_back(*pKP);
        }
        return vResultKeys;
    }

    // src/ORBExtractor.cpp:739-827
    void ComputeKeyPointsOctTree(std::vector<std::vector<KeyPoint>>& allKeypoints) {
        allKeypoints.resize(nlevels);
        mvCandidates.assign(nlevels, {});
        const float W = 30;
        std::vector<FastKp> vKeysCell;
        for (int level = 0; level < nlevels; ++level) {
            const Image& img = mvImagePyramid[level];
            const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
            const int maxBorderX = img.w - EDGE_THRESHOLD + 3, maxBorderY = img.h - EDGE_THRESHOLD + 3;
            std::vector<KeyPoint>& vToDistributeKeys = mvCandidates[level];
            const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
            const int nCols = (int)(width / W), nRows = (int)(height / W);
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minBorderY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBorderY - 3) continue;
                if (maxY > maxBorderY) maxY = (float)maxBorderY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minBorderX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBorderX - 6) continue;
                    if (maxX > maxBorderX) maxX = (float)maxBorderX;
                    const int y0 = (int)iniY, y1 = (int)maxY, x0 = (int)iniX, x1 = (int)maxX;
                    fast_nms(img.row(y0) + x0, x1 - x0, y1 - y0, img.w, iniThFAST, vKeysCell);
                    if (vKeysCell.empty()) fast_nms(img.row(y0) + x0, x1 - x0, y1 - y0, img.w, minThFAST, vKeysCell);
                    for (const FastKp& k : vKeysCell) {
                        KeyPoint kp;
                        kp.x = (float)k.x + j * wCell; kp.y = (float)k.y + i * hCell;
                        kp.size = 7.f; kp.angle = -1.f; kp.response = (float)k.response; kp.octave = 0; kp.class_id = -1;
                        vToDistributeKeys.push_back(kp);
                    }
                }
            }
            std::vector<KeyPoint>& keypoints = allKeypoints[level];
            keypoints = DistributeOctTree(vToDistributeKeys, minBorderX, maxBorderX, minBorderY, maxBorderY,
                                          mnFeaturesPerLevel[level]);
            const int scaledPatchSize = (int)(PATCH_SIZE * mvScaleFactor[level]);
            for (size_t i = 0; i < keypoints.size(); i++) {
                keypoints[i].x += minBorderX; keypoints[i].y += minBorderY;
                keypoints[i].octave = level;
                keypoints[i].size = (float)scaledPatchSize;
            }
        }
        for (int level = 0; level < nlevels; ++level)
            for (KeyPoint& kp : allKeypoints[level]) kp.angle = IC_Angle(mvImagePyramid[level], kp.x, kp.y);
    }

    // src/ORBExtractor.cpp:79-118
    static void computeOrbDescriptor(const KeyPoint& kpt, const Image& img, uint8_t* desc) {
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        float angle = (float)kpt.angle * factorPI;
        float a, b;
        det_sincosf(angle, &b, &a);
        const uint8_t* center = img.row(cv_round(kpt.y)) + cv_round(kpt.x);
        const int step = img.w;
        const signed char* pattern = kBriefPattern;
        auto GET = [&](int idx) -> int {
            float px = (float)pattern[2 * idx], py = (float)pattern[2 * idx + 1];
            return center[cv_round(px * b + py * a) * step + cv_round(px * a - py * b)];
        };
        for (int i = 0; i < 32; ++i, pattern += 32) {
            int val = 0;
            for (int k = 0; k < 8; k++) { int t0 = GET(2 * k), t1 = GET(2 * k + 1); val |= (t0 < t1) << k; }
            desc[i] = (uint8_t)val;
        }
    }

    // src/ORBExtractor.cpp:838-926. mask must be non-empty, same size as image.
    void operator()(const uint8_t* image, int cols, int rows, int stride, const uint8_t* mask, int mstride,
                    std::vector<KeyPoint>& _keypoints, std::vector<uint8_t>& _descriptors) {
        _keypoints.clear(); _descriptors.clear();
        if (!image || cols <= 0 || rows <= 0) return;
        const int width = cols, height = rows;
        ComputePyramid(image, cols, rows, stride);
        std::vector<std::vector<KeyPoint>> allKeypoints;
        ComputeKeyPointsOctTree(allKeypoints);
        mvDistributed = allKeypoints;
        mvBlurred.assign(nlevels, Image());
        for (int level = 0; level < nlevels; ++level) {
            std::vector<KeyPoint>& keypoints = allKeypoints[level];
            if (keypoints.empty()) continue;
            std::vector<std::pair<float, float>> keypoints_new;
            std::vector<KeyPoint> kept;
            const float scale = mvScaleFactor[level];
            for (const KeyPoint& keypoint : keypoints) {
                const float ptx = keypoint.x * scale, pty = keypoint.y * scale;
                if (face_in_cubemap_f(ptx, pty, faceW, faceH) == UNKNOWN_FACE) continue;
                if (ptx < 0 || (int)(ptx + 0.5f) >= width || pty < 0 || (int)(pty + 0.5f) >= height) continue;
                if (mask[(size_t)(int)(pty + 0.5f) * mstride + (int)(ptx + 0.5f)] == 0) continue;
                keypoints_new.push_back({ptx, pty});
                kept.push_back(keypoint);
            }
            const Image& src = mvImagePyramid[level];
            Image& work = mvBlurred[level];
            work.create(src.w, src.h);
            gaussian7(src.px.data(), src.w, src.h, src.w, work.px.data(), src.w);
            size_t off = _descriptors.size();
            _descriptors.resize(off + kept.size() * 32);
            for (size_t i = 0; i < kept.size(); i++) computeOrbDescriptor(kept[i], work, &_descriptors[off + i * 32]);
            for (size_t i = 0; i < kept.size(); i++) { kept[i].x = keypoints_new[i].first; kept[i].y = keypoints_new[i].second; }
            _keypoints.insert(_keypoints.end(), kept.begin(), kept.end());
        }
    }
};

}  // namespace orc
