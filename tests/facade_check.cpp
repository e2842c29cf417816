// Compile-only check of the facade headers against mock reference types (tests/test_facade_compile.py).
#include <map>
#include <mutex>
#include <set>
#include <type_traits>
#include <vector>
#include "../include/CubemapWarp.h"
#include "../include/ORBExtractor.h"
#include "../include/ORBMatcher.h"
#include "../include/Optimizer.h"
struct Vec3f { float v[3]; float operator()(int i) const { return v[i]; } };
struct KeyFrame;
struct MapPoint {
    unsigned long mnId = 0, mnBALocalForKF = 0; bool isBad() { return false; } cv::Mat GetWorldPos() { cv::Mat m; m.create(3, 1, CV_32F); return m; }
    std::map<KeyFrame*, size_t> GetObservations() { return {}; } void EraseObservation(KeyFrame*) {} void SetWorldPos(const cv::Mat&) {} void UpdateNormalAndDepth() {}
};
struct FeatVec : std::map<unsigned, std::vector<unsigned> > {};
struct Frame {
    int N = 0; cv::Mat mDescriptors, mTcw; std::vector<cv::KeyPoint> mvKeys; FeatVec mFeatVec; std::vector<Vec3f> mvKeyRays; std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier; std::vector<float> mvInvLevelSigma2; void SetPose(const cv::Mat&) {}
};
struct KeyFrame {
    unsigned long mnId = 0, mnBALocalForKF = 0, mnBAFixedForKF = 0; cv::Mat mDescriptors; std::vector<cv::KeyPoint> mvKeys; FeatVec mFeatVec; std::vector<Vec3f> mvKeyRays;
    std::vector<float> mvInvLevelSigma2; std::vector<MapPoint*> GetMapPointMatches() { return {}; } std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return {}; }
    bool isBad() { return false; } cv::Mat GetPose() { cv::Mat m; m.create(4, 4, CV_32F); return m; } void SetPose(const cv::Mat&) {} void EraseMapPointMatch(MapPoint*) {}
};
struct Map { std::mutex mMutexMapUpdate; };
struct Cam { float GetCosFovTh() { return -0.087f; } int GetCubeFaceWidth() { return 650; } int GetCubeFaceHeight() { return 650; } };
int facade_check() {
    ORBextractor ex(2000, 1.2f, 8, 20, 7); cv::Mat im, mask, desc; std::vector<cv::KeyPoint> kps; ex(im, mask, kps, desc);
    ORBMatcher m(0.7f, true); Frame F; KeyFrame K; std::vector<MapPoint*> out; int n = m.SearchByBoW(&K, F, out); KeyFrame K2; n += m.SearchByBoW(&K, &K2, out); n += ORBMatcher::DescriptorDistance(desc, desc);
    Cam cam; Map map; bool stop = false; n += Optimizer::PoseOptimization(&F, &cam); Optimizer::LocalBundleAdjustment(&K, &stop, &map, &cam);
    return n + ex.GetLevels();
}
