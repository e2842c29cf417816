// Drop-in IMPLEMENTATION of class ORBextractor on libcubemap_b200.so, compiled INSIDE the CubemapSLAM tree against the reference's own,
// unmodified include/ORBExtractor.h (:49-116): it replaces src/ORBExtractor.cpp in the build. Tracking.cpp:95-96
// constructs the two extractors and Frame::ExtractORB (src/Frame.cpp:178-181) calls operator() exactly as before.
// The class declaration cannot grow a member, so the device handle of every extractor lives in a side table keyed by `this`.
#include "ORBExtractor.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "CamModelGeneral.h"
#include "cubemap_b200.h"

namespace {
struct B200State {
    cslam_frontend* fe = nullptr; int rows = 0, cols = 0; const unsigned char* maskData = nullptr; int maskRows = 0, maskCols = 0;
    std::vector<cslam_keypoint> kps; std::vector<uint8_t> desc;
};
std::mutex g_mu;
std::map<const ORBextractor*, B200State*> g_state;
B200State* state_of(const ORBextractor* ex) {
    std::lock_guard<std::mutex> lk(g_mu);
    B200State*& s = g_state[ex];
    if (!s) s = new B200State;
    return s;
}
void fatal(const char* msg) { std::fprintf(stderr, "ORBextractor (cubemap_b200): %s\n", msg); std::exit(EXIT_FAILURE); }
}  // namespace

// reference src/ORBExtractor.cpp:381-442 (the tables the getters expose; the device builds its own copies in cslam_frontend_create)
ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
    mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) { mvScaleFactor[i] = mvScaleFactor[i - 1] * scaleFactor; mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i]; }
    mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
    mvImagePyramid.resize(nlevels); mvMaskPyramid.resize(nlevels);   // members of the reference class; never filled (nothing outside the class reads them)
}

// reference src/ORBExtractor.cpp:838-926
void ORBextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors) {
    if (_image.empty()) return;                                   // :841-842
    cv::Mat image = _image.getMat(), mask = _mask.getMat();
    if (image.type() != CV_8UC1) fatal("image must be CV_8UC1 (reference :845)");
    if (mask.empty() || mask.type() != CV_8UC1 || mask.rows != image.rows || mask.cols != image.cols) fatal("mask must be a non-empty CV_8UC1 of the image size (reference :848)");
    B200State* s = state_of(this);
    if (!s->fe || s->rows != image.rows || s->cols != image.cols || s->maskData != mask.data || s->maskRows != mask.rows || s->maskCols != mask.cols) {
        CamModelGeneral* cam = CamModelGeneral::GetCamera();
        if (image.cols != 3 * cam->GetCubeFaceWidth() || image.rows != 3 * cam->GetCubeFaceHeight()) fatal("image must be the 3W x 3H cubemap canvas of CamModelGeneral::GetCamera()");
        if (s->fe) cslam_frontend_destroy(s->fe);
        cslam_cam_params cp; std::memset(&cp, 0, sizeof(cp));
        cp.face_w = cam->GetCubeFaceWidth(); cp.face_h = cam->GetCubeFaceHeight();   // extraction only: no warp maps (Iw = Ih = 0)
        cslam_orb_params op; op.nfeatures = nfeatures; op.scale_factor = (float)scaleFactor; op.nlevels = nlevels; op.ini_th_fast = iniThFAST; op.min_th_fast = minThFAST;
        if (cslam_frontend_create(&s->fe, 0, &cp, &op, mask.data, (int)mask.step, 1) != CSLAM_OK) fatal(cslam_last_error());
        s->rows = image.rows; s->cols = image.cols; s->maskData = mask.data; s->maskRows = mask.rows; s->maskCols = mask.cols;
        const int cap = cslam_frontend_kp_capacity(s->fe);
        s->kps.resize(cap); s->desc.resize((size_t)cap * 32);
    }
    int32_t n = 0;
    if (cslam_orb_extract(s->fe, image.data, (int)image.step, 1, s->kps.data(), s->desc.data(), &n) != CSLAM_OK) fatal(cslam_last_error());
    static_assert(sizeof(cv::KeyPoint) == sizeof(cslam_keypoint), "cv::KeyPoint layout");
    _keypoints.clear(); _keypoints.resize(n);
    if (n) std::memcpy(static_cast<void*>(_keypoints.data()), s->kps.data(), (size_t)n * sizeof(cslam_keypoint));
    if (n == 0) { _descriptors.release(); return; }               // :863-864
    _descriptors.create(n, 32, CV_8U);
    cv::Mat descriptors = _descriptors.getMat();
    for (int i = 0; i < n; i++) std::memcpy(descriptors.ptr<uchar>(i), s->desc.data() + (size_t)i * 32, 32);
}
