/* ORBExtractor.h — drop-in facade with the reference's class signature (reference include/ORBExtractor.h:49-116,
 * implementation src/ORBExtractor.cpp:381-442,838-926) on top of libcubemap_b200.so. Frame::ExtractORB
 * (reference src/Frame.cpp:178-181) calls `(*mpORBextractor)(im, mask, mvKeys, mDescriptors)` unchanged.
 *
 * Differences a maintainer should know: the image must be the 3W x 3H cubemap canvas (it always is in the reference),
 * the first call fixes the canvas size and uploads the mask (re-uploaded if a different mask pointer/size is passed),
 * mvImagePyramid / mvMaskPyramid are not exposed (nothing outside the class reads them in the reference). */
#ifndef CSLAM_ORBEXTRACTOR_H
#define CSLAM_ORBEXTRACTOR_H
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "cubemap_b200.h"
#include "cv_compat.h"

class ORBextractor {
public:
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
        : fe_(nullptr), device_(device), maskData_(nullptr), maskRows_(0), maskCols_(0) {
        orb_.nfeatures = nfeatures; orb_.scale_factor = scaleFactor; orb_.nlevels = nlevels; orb_.ini_th_fast = iniThFAST; orb_.min_th_fast = minThFAST;
        mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) { mvScaleFactor[i] = mvScaleFactor[i - 1] * scaleFactor; mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i]; }
        for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
    }
    ~ORBextractor() { if (fe_) cslam_frontend_destroy(fe_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (reference src/ORBExtractor.cpp:838-926).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors) {
        if (image.empty()) return;                                   // reference :841-842
        if (mask.empty() || mask.rows != image.rows || mask.cols != image.cols) fatal("mask must be a non-empty CV_8UC1 of the image size (reference :848)");
        ensure(image, mask);
        const int cap = cslam_frontend_kp_capacity(fe_);
        kps_.resize(cap); desc_.resize((size_t)cap * 32);
        int32_t n = 0;
        if (cslam_orb_extract(fe_, image.data, (int)image.step, 1, kps_.data(), desc_.data(), &n) != CSLAM_OK) fatal(cslam_last_error());
        keypoints.clear(); keypoints.resize(n);
        static_assert(sizeof(cv::KeyPoint) == sizeof(cslam_keypoint), "cv::KeyPoint layout");
        if (n) std::memcpy(static_cast<void*>(keypoints.data()), kps_.data(), (size_t)n * sizeof(cslam_keypoint));
        if (n == 0) { descriptors.release(); return; }              // reference :863-864
        descriptors.create(n, 32, CV_8U);
        for (int i = 0; i < n; i++) std::memcpy(descriptors.ptr<unsigned char>(i), desc_.data() + (size_t)i * 32, 32);
    }

    int inline GetLevels() { return orb_.nlevels; }
    float inline GetScaleFactor() { return orb_.scale_factor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

protected:
    void ensure(const cv::Mat& image, const cv::Mat& mask) {
        if (fe_ && (image.cols != 3 * cam_.face_w || image.rows != 3 * cam_.face_h)) { cslam_frontend_destroy(fe_); fe_ = nullptr; }
        if (fe_ && mask.data == maskData_ && mask.rows == maskRows_ && mask.cols == maskCols_) return;
        if (fe_) { cslam_frontend_destroy(fe_); fe_ = nullptr; }
        if (image.cols % 3 || image.rows % 3 || image.cols != image.rows) fatal("image must be the square 3W x 3H cubemap canvas");
        std::memset(&cam_, 0, sizeof(cam_));
        cam_.face_w = image.cols / 3; cam_.face_h = image.rows / 3;   // CamModelGeneral::GetCamera()->GetCubeFaceWidth/Height
        cam_.Iw = 0; cam_.Ih = 0;                                     // extraction only: no warp maps
        if (cslam_frontend_create(&fe_, device_, &cam_, &orb_, mask.data, (int)mask.step, 1) != CSLAM_OK) fatal(cslam_last_error());
        maskData_ = mask.data; maskRows_ = mask.rows; maskCols_ = mask.cols;
    }
    static void fatal(const char* msg) { std::fprintf(stderr, "ORBextractor (cubemap_b200): %s\n", msg); std::exit(EXIT_FAILURE); }

    cslam_frontend* fe_; int device_;
    cslam_orb_params orb_; cslam_cam_params cam_;
    const unsigned char* maskData_; int maskRows_, maskCols_;
    std::vector<cslam_keypoint> kps_; std::vector<uint8_t> desc_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};
#endif
