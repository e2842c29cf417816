/* ORBMatcher.h — facade with the reference's ORBMatcher signature for the Hamming hot path (reference
 * include/ORBMatcher.h:42-104; src/ORBMatcher.cpp:42-45,409-539,905-967) on top of libcubemap_b200.so.
 * SearchByBoW is a template over the reference's KeyFrame / Frame / MapPoint types (only their public members are used:
 * mDescriptors, mvKeys, mFeatVec, N, GetMapPointMatches(), isBad()), so Tracking::TrackReferenceKeyFrame
 * (src/Tracking.cpp:577) and Relocalization (:1045) call it unchanged. The projection / epipolar / Sim3 / Fuse searches
 * are "next" rows of SURVEY.md §8(f) and are not provided yet. */
#ifndef CSLAM_ORBMATCHER_H
#define CSLAM_ORBMATCHER_H
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "cubemap_b200.h"
#include "cv_compat.h"

class ORBMatcher {
public:
    ORBMatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

    // Computes the Hamming distance between two ORB descriptors (reference src/ORBMatcher.cpp:951-967)
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
        int32_t d = 0;
        if (cslam_hamming(handle(), a.ptr<unsigned char>(), b.ptr<unsigned char>(), 1, &d) != CSLAM_OK) fatal();
        return d;
    }

    // Search matches between MapPoints in a KeyFrame and ORB in a Frame, by vocabulary node (reference :409-539)
    template <class KeyFrameT, class FrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches) {
        const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();
        const int nKF = (int)vpMapPointsKF.size(), nF = F.N;
        vpMapPointMatches = std::vector<MapPointT*>(nF, static_cast<MapPointT*>(NULL));
        std::vector<uint8_t> dK((size_t)nKF * 32), dF((size_t)nF * 32), valid(nKF);
        std::vector<float> aK(nKF), aF(nF); std::vector<int32_t> nodeK(nKF, 0x7ffff), nodeF(nF, 0x7fffe), matchF(nF);
        for (int i = 0; i < nKF; i++) { std::memcpy(&dK[(size_t)i * 32], pKF->mDescriptors.template ptr<unsigned char>(i), 32); aK[i] = pKF->mvKeys[i].angle;
            valid[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad(); }
        for (int i = 0; i < nF; i++) { std::memcpy(&dF[(size_t)i * 32], F.mDescriptors.template ptr<unsigned char>(i), 32); aF[i] = F.mvKeys[i].angle; }
        // DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned int>>; features absent from it get distinct sentinels
        for (typename decltype(pKF->mFeatVec)::const_iterator it = pKF->mFeatVec.begin(); it != pKF->mFeatVec.end(); ++it)
            for (size_t k = 0; k < it->second.size(); k++) nodeK[it->second[k]] = (int32_t)it->first;
        for (typename decltype(F.mFeatVec)::const_iterator it = F.mFeatVec.begin(); it != F.mFeatVec.end(); ++it)
            for (size_t k = 0; k < it->second.size(); k++) nodeF[it->second[k]] = (int32_t)it->first;
        int32_t n = 0;
        if (cslam_search_by_bow(handle(), dK.data(), aK.data(), valid.data(), nodeK.data(), nKF, dF.data(), aF.data(), nodeF.data(), nF, 1, mfNNratio,
                                mbCheckOrientation, matchF.data(), &n) != CSLAM_OK) fatal();
        for (int j = 0; j < nF; j++) if (matchF[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[matchF[j]];
        return n;
    }

    // Search matches between the MapPoints of two key frames, by vocabulary node (reference :541-674, used by loop closing)
    template <class KeyFrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12) {
        const std::vector<MapPointT*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
        const int n1 = (int)mp1.size(), n2 = (int)mp2.size();
        vpMatches12 = std::vector<MapPointT*>(n1, static_cast<MapPointT*>(NULL));
        std::vector<uint8_t> d1((size_t)n1 * 32), d2((size_t)n2 * 32), v1(n1), v2(n2);
        std::vector<float> a1(n1), a2(n2); std::vector<int32_t> nd1(n1, 0x7ffff), nd2(n2, 0x7fffe), m12(n1);
        for (int i = 0; i < n1; i++) { std::memcpy(&d1[(size_t)i * 32], pKF1->mDescriptors.template ptr<unsigned char>(i), 32); a1[i] = pKF1->mvKeys[i].angle; v1[i] = mp1[i] && !mp1[i]->isBad(); }
        for (int i = 0; i < n2; i++) { std::memcpy(&d2[(size_t)i * 32], pKF2->mDescriptors.template ptr<unsigned char>(i), 32); a2[i] = pKF2->mvKeys[i].angle; v2[i] = mp2[i] && !mp2[i]->isBad(); }
        for (typename decltype(pKF1->mFeatVec)::const_iterator it = pKF1->mFeatVec.begin(); it != pKF1->mFeatVec.end(); ++it)
            for (size_t k = 0; k < it->second.size(); k++) nd1[it->second[k]] = (int32_t)it->first;
        for (typename decltype(pKF2->mFeatVec)::const_iterator it = pKF2->mFeatVec.begin(); it != pKF2->mFeatVec.end(); ++it)
            for (size_t k = 0; k < it->second.size(); k++) nd2[it->second[k]] = (int32_t)it->first;
        int32_t n = 0;
        if (cslam_search_by_bow_kf(handle(), d1.data(), a1.data(), v1.data(), nd1.data(), n1, d2.data(), a2.data(), v2.data(), nd2.data(), n2, 1, mfNNratio,
                                   mbCheckOrientation, m12.data(), &n) != CSLAM_OK) fatal();
        for (int i = 0; i < n1; i++) if (m12[i] >= 0) vpMatches12[i] = mp2[m12[i]];
        return n;
    }

    static const int TH_LOW = 50;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 12;

protected:
    static cslam_matcher* handle() {   // one matcher per host thread: ORBMatcher objects live on three threads in the reference
        static thread_local cslam_matcher* m = nullptr;
        if (!m && cslam_matcher_create(&m, 0, 1, 4096) != CSLAM_OK) fatal();
        return m;
    }
    static void fatal() { std::fprintf(stderr, "ORBMatcher (cubemap_b200): %s\n", cslam_last_error()); std::exit(EXIT_FAILURE); }
    float mfNNratio; bool mbCheckOrientation;
};
#endif
