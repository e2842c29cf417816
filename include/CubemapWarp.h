/* CubemapWarp.h — System::CvtFisheyeToCubeMap_reverseQuery_withInterpolation (reference include/System.h:106-107,
 * src/System.cpp:327-355) + the map construction of System::CreateUndistortRectifyMap (src/System.cpp:301-324). */
#ifndef CSLAM_CUBEMAPWARP_H
#define CSLAM_CUBEMAPWARP_H
#include <cstdio>
#include <cstdlib>
#include "cubemap_b200.h"
#include "cv_compat.h"

class CubemapWarp {
public:
    // cam: the values System::System reads from the settings file (src/System.cpp:63-89)
    CubemapWarp(const cslam_cam_params& cam, int device = 0) : fe_(nullptr), cam_(cam) {
        cslam_orb_params orb = {1000, 1.2f, 8, 20, 7};
        std::vector<unsigned char> mask((size_t)9 * cam.face_w * cam.face_h, 255);
        if (cslam_frontend_create(&fe_, device, &cam_, &orb, mask.data(), 3 * cam.face_w, 1) != CSLAM_OK) {
            std::fprintf(stderr, "CubemapWarp: %s\n", cslam_last_error()); std::exit(EXIT_FAILURE);
        }
    }
    ~CubemapWarp() { if (fe_) cslam_frontend_destroy(fe_); }
    // cubemapImg: caller-allocated zeroed 3H x 3W CV_8U canvas (Examples/cubemap_lafida.cpp:111); only the 5 face ROIs are written.
    void CvtFisheyeToCubeMap_reverseQuery_withInterpolation(cv::Mat& cubemapImg, const cv::Mat& fisheyeImg, int /*interpolation = INTER_LINEAR*/) {
        if (fisheyeImg.cols != cam_.Iw || fisheyeImg.rows != cam_.Ih || (int)fisheyeImg.step != cam_.Iw) { std::fprintf(stderr, "CubemapWarp: fisheye image must be continuous %dx%d\n", cam_.Iw, cam_.Ih); std::exit(EXIT_FAILURE); }
        if (cslam_warp(fe_, fisheyeImg.data, 1, cubemapImg.data, (int)cubemapImg.step) != CSLAM_OK) { std::fprintf(stderr, "CubemapWarp: %s\n", cslam_last_error()); std::exit(EXIT_FAILURE); }
    }
private:
    cslam_frontend* fe_; cslam_cam_params cam_;
};
#endif
