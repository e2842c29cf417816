// Drop-in DEFINITION of System::CvtFisheyeToCubeMap_reverseQuery_withInterpolation (reference include/System.h:106-107, src/System.cpp:327-355)
// on libcubemap_b200.so, compiled INSIDE the CubemapSLAM tree against the reference's own, unmodified headers. The replaced body in
// src/System.cpp is renamed away at compile time (COMPILE_DEFINITIONS
// CvtFisheyeToCubeMap_reverseQuery_withInterpolation=CvtFisheyeToCubeMap_reverseQuery_withInterpolation_cpu on that file), no source edit;
// Examples/cubemap_lafida.cpp:143 calls it unchanged. The device front end (maps quantised like cv::remap does) is created on first use
// from the CamModelGeneral singleton System::System configured (src/System.cpp:89).
#include "System.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "CamModelGeneral.h"
#include "cubemap_b200.h"

namespace {
cslam_frontend* g_warp = nullptr; int g_W = 0, g_Iw = 0, g_Ih = 0;
void fatal(const char* m) { std::fprintf(stderr, "System warp (cubemap_b200): %s\n", m); std::exit(EXIT_FAILURE); }
cslam_frontend* warp_frontend() {
    CamModelGeneral* cam = CamModelGeneral::GetCamera();
    if (g_warp && g_W == cam->GetCubeFaceWidth() && g_Iw == cam->GetFisheyeWidth() && g_Ih == cam->GetFisheyeHeight()) return g_warp;
    if (g_warp) cslam_frontend_destroy(g_warp);
    cslam_cam_params cp; std::memset(&cp, 0, sizeof(cp));
    cp.c = cam->Get_c(); cp.d = cam->Get_d(); cp.e = cam->Get_e(); cp.u0 = cam->Get_u0(); cp.v0 = cam->Get_v0();
    const cv::Mat_<double> p = cam->Get_P(), ip = cam->Get_invP();
    for (int i = 0; i < 5 && i < (int)p.total(); i++) cp.poly[i] = p.at<double>(i);
    for (int i = 0; i < 12 && i < (int)ip.total(); i++) cp.invpoly[i] = ip.at<double>(i);
    cp.Iw = cam->GetFisheyeWidth(); cp.Ih = cam->GetFisheyeHeight(); cp.face_w = cam->GetCubeFaceWidth(); cp.face_h = cam->GetCubeFaceHeight(); cp.fov_deg = 0;
    cslam_orb_params orb = {1000, 1.2f, 8, 20, 7};   // the warp-only front end never extracts
    std::vector<unsigned char> mask((size_t)9 * cp.face_w * cp.face_h, 255);
    if (cslam_frontend_create(&g_warp, 0, &cp, &orb, mask.data(), 3 * cp.face_w, 1) != CSLAM_OK) fatal(cslam_last_error());
    g_W = cp.face_w; g_Iw = cp.Iw; g_Ih = cp.Ih;
    return g_warp;
}
}  // namespace

void System::CvtFisheyeToCubeMap_reverseQuery_withInterpolation(cv::Mat& cubemapImg, const cv::Mat& fisheyeImg, int interpolation, int borderType, const cv::Scalar&) {
    if (interpolation != cv::INTER_LINEAR || borderType != cv::BORDER_CONSTANT) fatal("only INTER_LINEAR / BORDER_CONSTANT (what the examples pass)");
    cslam_frontend* fe = warp_frontend();
    if (fisheyeImg.cols != g_Iw || fisheyeImg.rows != g_Ih || fisheyeImg.type() != CV_8UC1) fatal("fisheye image must be the CV_8UC1 Iw x Ih image of the settings file");
    cv::Mat src = fisheyeImg.isContinuous() ? fisheyeImg : fisheyeImg.clone();
    if (cslam_warp(fe, src.data, 1, cubemapImg.data, (int)cubemapImg.step) != CSLAM_OK) fatal(cslam_last_error());
}
