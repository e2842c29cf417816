// Drop-in DEFINITIONS of the ORBMatcher searches that run on libcubemap_b200.so, compiled INSIDE the CubemapSLAM tree against the
// reference's own, unmodified include/ORBMatcher.h (:42-104). Everything else of the class (constructor, DescriptorDistance - the 8-word
// host popcount of src/ORBMatcher.cpp:951-967 -, ComputeThreeMaxima, the projection / epipolar / Sim3 / Fuse searches not listed here)
// stays the reference's src/ORBMatcher.cpp; the bodies replaced here are renamed away at compile time
// (COMPILE_DEFINITIONS SearchByBoW=SearchByBoW_cpu on that one file), no source edit.
#include "ORBMatcher.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cubemap_b200.h"

namespace {
cslam_matcher* b200_matcher() {   // ORBMatcher objects live on three threads in the reference: one device handle per host thread
    static thread_local cslam_matcher* m = nullptr;
    if (!m && cslam_matcher_create(&m, 0, 1, 4096) != CSLAM_OK) { std::fprintf(stderr, "ORBMatcher (cubemap_b200): %s\n", cslam_last_error()); std::exit(EXIT_FAILURE); }
    return m;
}
void b200_fatal() { std::fprintf(stderr, "ORBMatcher (cubemap_b200): %s\n", cslam_last_error()); std::exit(EXIT_FAILURE); }
// DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned int>>; features absent from it get sentinels that never meet
void node_ids(const DBoW2::FeatureVector& fv, std::vector<int32_t>& node) {
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        for (size_t k = 0; k < it->second.size(); k++) node[it->second[k]] = (int32_t)it->first;
}
void gather(const cv::Mat& desc, const std::vector<cv::KeyPoint>& keys, int n, std::vector<uint8_t>& d, std::vector<float>& a) {
    d.resize((size_t)n * 32); a.resize(n);
    for (int i = 0; i < n; i++) { std::memcpy(&d[(size_t)i * 32], desc.ptr<uchar>(i), 32); a[i] = keys[i].angle; }
}
}  // namespace

// reference src/ORBMatcher.cpp:409-539 (Tracking::TrackReferenceKeyFrame :577, Relocalization :1045)
int ORBMatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    const int nKF = (int)vpMapPointsKF.size(), nF = F.N;
    vpMapPointMatches = std::vector<MapPoint*>(nF, static_cast<MapPoint*>(NULL));
    std::vector<uint8_t> dK, dF, valid(nKF); std::vector<float> aK, aF;
    gather(pKF->mDescriptors, pKF->mvKeys, nKF, dK, aK); gather(F.mDescriptors, F.mvKeys, nF, dF, aF);
    for (int i = 0; i < nKF; i++) valid[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
    std::vector<int32_t> nodeK(nKF, 0x7ffff), nodeF(nF, 0x7fffe), matchF(nF);
    node_ids(pKF->mFeatVec, nodeK); node_ids(F.mFeatVec, nodeF);
    int32_t n = 0;
    if (cslam_search_by_bow(b200_matcher(), dK.data(), aK.data(), valid.data(), nodeK.data(), nKF, dF.data(), aF.data(), nodeF.data(), nF, 1, mfNNratio,
                            mbCheckOrientation, matchF.data(), &n) != CSLAM_OK) b200_fatal();
    for (int j = 0; j < nF; j++) if (matchF[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[matchF[j]];
    return n;
}

// reference src/ORBMatcher.cpp:541-674 (LoopClosing::ComputeSim3 :238)
int ORBMatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
    const int n1 = (int)mp1.size(), n2 = (int)mp2.size();
    vpMatches12 = std::vector<MapPoint*>(n1, static_cast<MapPoint*>(NULL));
    std::vector<uint8_t> d1, d2, v1(n1), v2(n2); std::vector<float> a1, a2;
    gather(pKF1->mDescriptors, pKF1->mvKeys, n1, d1, a1); gather(pKF2->mDescriptors, pKF2->mvKeys, n2, d2, a2);
    for (int i = 0; i < n1; i++) v1[i] = mp1[i] && !mp1[i]->isBad();
    for (int i = 0; i < n2; i++) v2[i] = mp2[i] && !mp2[i]->isBad();
    std::vector<int32_t> nd1(n1, 0x7ffff), nd2(n2, 0x7fffe), m12(n1);
    node_ids(pKF1->mFeatVec, nd1); node_ids(pKF2->mFeatVec, nd2);
    int32_t n = 0;
    if (cslam_search_by_bow_kf(b200_matcher(), d1.data(), a1.data(), v1.data(), nd1.data(), n1, d2.data(), a2.data(), v2.data(), nd2.data(), n2, 1, mfNNratio,
                               mbCheckOrientation, m12.data(), &n) != CSLAM_OK) b200_fatal();
    for (int i = 0; i < n1; i++) if (m12[i] >= 0) vpMatches12[i] = mp2[m12[i]];
    return n;
}
