// ORACLE (test infrastructure, not product code): C entry points of oracle/_ref/libref.so.
//
// libref.so = the reference's OWN translation units, compiled UNMODIFIED from where they lie under /root/reference
// (src/ORBExtractor.cpp, CamModelGeneral.cpp, Frame.cpp, ORBMatcher.cpp, KeyFrame.cpp, MapPoint.cpp, Map.cpp, KeyFrameDatabase.cpp and
// ThirdParty/DBoW2/DBoW2/*.cpp) against the cv:: shim in oracle/cvshim (no OpenCV C++ exists in this container), plus this file.
// This file holds only (1) flat C wrappers that call the reference classes the way the reference's own callers do (file:line cited),
// (2) the few caller-side lines of src/System.cpp that cannot be compiled here (System.cpp needs Pangolin): the map loop of
// CreateUndistortRectifyMap (:301-324) and the five cv::remap calls (:327-355), (3) Converter::toDescriptorVector
// (src/Converter.cpp:29-39; Converter.cpp needs Eigen), and (4) two optional PINS of behaviour the reference leaves to its environment:
//   PIN_ALLOC   std::list<ExtractorNode> nodes come from a monotonic arena, so the reference's `sort(pair<int,ExtractorNode*>)`
//               tie-break by heap address (src/ORBExtractor.cpp:658) is "later-created node first" -- the order the oracle defines;
//   PIN_SINCOS  sincosf (what gcc emits for cos/sin at src/ORBExtractor.cpp:83-84) is the fp64 det_sincos of oracle/cvprim.h instead
//               of glibc's, which is not correctly rounded.
// With both pins the reference code is deterministic and the tests require oracle == libref bit for bit; without them the tests
// QUANTIFY how far a stock glibc build drifts (tests/test_oracle_ref.py).
#include <dlfcn.h>
#include <sys/mman.h>
#include <atomic>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "CamModelGeneral.h"
#include "Converter.h"
#include "Frame.h"
#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
#include "ORBExtractor.h"
#include "ORBMatcher.h"

// ------------------------------------------------------------------------------------------------- pins
enum { PIN_ALLOC = 1, PIN_SINCOS = 2 };
static std::atomic<int> g_pins(0);

static const size_t kNodeBytes = sizeof(std::_List_node<ExtractorNode>);
static const size_t kArenaBytes = (size_t)4 << 30;   // virtual reservation only (MAP_NORESERVE)
static char* g_arena = nullptr;
static std::atomic<size_t> g_arenaUsed(0);
static std::once_flag g_arenaOnce;
static void arena_init() {
    void* p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    g_arena = (p == MAP_FAILED) ? nullptr : (char*)p;
}
static inline bool in_arena(const void* p) { return g_arena && (const char*)p >= g_arena && (const char*)p < g_arena + kArenaBytes; }
static inline void* ref_alloc(size_t n) {
    if ((g_pins.load(std::memory_order_relaxed) & PIN_ALLOC) && n == kNodeBytes) {
        std::call_once(g_arenaOnce, arena_init);
        if (g_arena) {
            const size_t off = g_arenaUsed.fetch_add((n + 15) & ~(size_t)15);
            if (off + n <= kArenaBytes) return g_arena + off;
        }
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
static inline void ref_free(void* p) { if (p && !in_arena(p)) std::free(p); }
// replaced inside this shared object only (linked with -Bsymbolic; the library is loaded RTLD_LOCAL)
void* operator new(size_t n) { return ref_alloc(n); }
void* operator new[](size_t n) { return ref_alloc(n); }
void operator delete(void* p) noexcept { ref_free(p); }
void operator delete[](void* p) noexcept { ref_free(p); }
void operator delete(void* p, size_t) noexcept { ref_free(p); }
void operator delete[](void* p, size_t) noexcept { ref_free(p); }

extern "C" void sincosf(float x, float* s, float* c) {
    if (g_pins.load(std::memory_order_relaxed) & PIN_SINCOS) { orc::det_sincosf(x, s, c); return; }
    typedef void (*fn_t)(float, float*, float*);
    static fn_t real = (fn_t)dlsym(RTLD_NEXT, "sincosf");
    real(x, s, c);
}

// src/Converter.cpp:29-39 (Converter.cpp itself needs Eigen / g2o): one Mat per descriptor row
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat& Descriptors) {
    std::vector<cv::Mat> vDesc;
    vDesc.reserve(Descriptors.rows);
    for (int j = 0; j < Descriptors.rows; j++) vDesc.push_back(Descriptors.row(j));
    return vDesc;
}

struct RefCam { double c, d, e, u0, v0, p[5], invp[12]; int Iw, Ih, faceW, faceH; double fov; };   // same layout as orc::CamParams / oracle.CamParams

extern "C" {

void ref_set_pins(int mask) { g_pins.store(mask); }
int ref_get_pins() { return g_pins.load(); }
size_t ref_arena_used() { return g_arenaUsed.load(); }

// System::System, src/System.cpp:63-89
void ref_set_camera(const RefCam* cp) {
    cv::Mat_<double> poly = cv::Mat::zeros(5, 1, CV_64F);
    for (int i = 0; i < 5; ++i) poly.at<double>(i, 0) = cp->p[i];
    cv::Mat_<double> invpoly = cv::Mat::zeros(12, 1, CV_64F);
    for (int i = 0; i < 12; ++i) invpoly.at<double>(i, 0) = cp->invp[i];
    double cdeu0v0[5] = {cp->c, cp->d, cp->e, cp->u0, cp->v0};
    const int nFaceH = cp->faceH, nFaceW = cp->faceW;
    double fx = static_cast<double>(nFaceW) / 2, fy = static_cast<double>(nFaceH) / 2;
    double cx = static_cast<double>(nFaceW) / 2, cy = static_cast<double>(nFaceH) / 2;
    std::streambuf* old = std::cout.rdbuf(nullptr);   // SetCosFovTh prints two lines
    CamModelGeneral::GetCamera()->SetCamParams(cdeu0v0, poly, invpoly, cp->Iw, cp->Ih, fx, fy, cx, cy, nFaceW, nFaceH, cp->fov);
    std::cout.rdbuf(old);
}
float ref_cos_fov_th() { return CamModelGeneral::GetCamera()->GetCosFovTh(); }

// System::CreateUndistortRectifyMap, src/System.cpp:301-324 (the loop around CamModelGeneral::CubemapToFisheye)
void ref_build_maps(float* map1, float* map2) {
    CamModelGeneral* cam = CamModelGeneral::GetCamera();
    const int width3 = cam->GetCubeFaceWidth() * 3, height3 = cam->GetCubeFaceHeight() * 3;
    cv::Mat mMap1(height3, width3, CV_32F, map1), mMap2(height3, width3, CV_32F, map2);
    mMap1.setTo(0); mMap2.setTo(0);
    const int Iw = cam->GetFisheyeWidth(), Ih = cam->GetFisheyeHeight();
    for (int y = 0; y < height3; ++y)
        for (int x = 0; x < width3; ++x) {
            double u, v;
            cam->CubemapToFisheye(u, v, static_cast<double>(x), static_cast<double>(y));
            if (u < 0 || v < 0 || u >= Iw || v >= Ih) continue;
            mMap1.at<float>(y, x) = static_cast<float>(u);
            mMap2.at<float>(y, x) = static_cast<float>(v);
        }
}
void ref_cubemap_to_fisheye(double up, double vp, double* uf, double* vf) { CamModelGeneral::GetCamera()->CubemapToFisheye(*uf, *vf, up, vp); }

// System::CvtFisheyeToCubeMap_reverseQuery_withInterpolation, src/System.cpp:327-355: five cv::remap calls on face ROIs
static void warp_into(cv::Mat& cubemapImg, const cv::Mat& fisheyeImg, const cv::Mat& mMap1, const cv::Mat& mMap2) {
    const int width = CamModelGeneral::GetCamera()->GetCubeFaceWidth(), height = CamModelGeneral::GetCamera()->GetCubeFaceHeight();
    const int r0[5] = {height, height, height, 0, 2 * height}, c0[5] = {width, 0, 2 * width, width, width};   // front, left, right, upper, lower
    for (int f = 0; f < 5; f++) {
        cv::Mat dst = cubemapImg.rowRange(r0[f], r0[f] + height).colRange(c0[f], c0[f] + width);
        cv::remap(fisheyeImg, dst, mMap1.rowRange(r0[f], r0[f] + height).colRange(c0[f], c0[f] + width),
                  mMap2.rowRange(r0[f], r0[f] + height).colRange(c0[f], c0[f] + width), cv::INTER_LINEAR, cv::BORDER_CONSTANT, cv::Scalar());
    }
}
void ref_warp(const uint8_t* fisheye, const float* map1, const float* map2, uint8_t* canvas) {
    CamModelGeneral* cam = CamModelGeneral::GetCamera();
    const int W3 = cam->GetCubeFaceWidth() * 3, H3 = cam->GetCubeFaceHeight() * 3;
    cv::Mat cub(H3, W3, CV_8UC1, canvas), fe(cam->GetFisheyeHeight(), cam->GetFisheyeWidth(), CV_8UC1, (void*)fisheye);
    cv::Mat m1(H3, W3, CV_32F, (void*)map1), m2(H3, W3, CV_32F, (void*)map2);
    warp_into(cub, fe, m1, m2);
}

// ---- ORBextractor (include/ORBExtractor.h:49-116), called like Frame::ExtractORB does (src/Frame.cpp:178-181)
struct RefOrb { ORBextractor* ex; std::vector<cv::KeyPoint> kps; cv::Mat desc; };
void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    RefOrb* h = new RefOrb; h->ex = new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh); return h;
}
void ref_orb_destroy(void* hv) { RefOrb* h = (RefOrb*)hv; delete h->ex; delete h; }
int ref_orb_extract(void* hv, const uint8_t* img, int cols, int rows, const uint8_t* mask) {
    RefOrb* h = (RefOrb*)hv;
    cv::Mat im(rows, cols, CV_8UC1, (void*)img), mk(rows, cols, CV_8UC1, (void*)mask);
    (*h->ex)(im, mk, h->kps, h->desc);
    return (int)h->kps.size();
}
void ref_orb_result(void* hv, cv::KeyPoint* kps, uint8_t* desc) {
    RefOrb* h = (RefOrb*)hv;
    static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
    const size_t n = h->kps.size();
    if (!n) return;
    std::memcpy((void*)kps, h->kps.data(), n * sizeof(cv::KeyPoint));
    for (size_t i = 0; i < n; i++) std::memcpy(desc + 32 * i, h->desc.ptr<uchar>((int)i), 32);
}
void ref_orb_level_size(void* hv, int level, int* w, int* hh) { const cv::Mat& L = ((RefOrb*)hv)->ex->mvImagePyramid[level]; *w = L.cols; *hh = L.rows; }
void ref_orb_level_image(void* hv, int level, uint8_t* out) {
    const cv::Mat& L = ((RefOrb*)hv)->ex->mvImagePyramid[level];
    for (int y = 0; y < L.rows; y++) std::memcpy(out + (size_t)y * L.cols, L.ptr<uchar>(y), (size_t)L.cols);
}

// timed CPU baseline (bench.py cpu_baseline.kind "reference"): warp + ORBextractor::operator() over nframes frames with nthreads
// independent workers, each owning its extractor like one reference process per core. Returns the total keypoint count.
long ref_warp_extract_batch(const uint8_t* fisheyes, int nframes, const float* map1, const float* map2, const uint8_t* mask, int nfeatures, float scaleFactor,
                            int nlevels, int iniTh, int minTh, int nthreads) {
    CamModelGeneral* cam = CamModelGeneral::GetCamera();
    const int W3 = cam->GetCubeFaceWidth() * 3, H3 = cam->GetCubeFaceHeight() * 3, Iw = cam->GetFisheyeWidth(), Ih = cam->GetFisheyeHeight();
    std::vector<long> totals(nthreads, 0);
    auto work = [&](int t) {
        ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
        cv::Mat canvas = cv::Mat::zeros(H3, W3, CV_8UC1), mk(H3, W3, CV_8UC1, (void*)mask), desc;   // Examples/cubemap_lafida.cpp:111
        cv::Mat m1(H3, W3, CV_32F, (void*)map1), m2(H3, W3, CV_32F, (void*)map2);
        std::vector<cv::KeyPoint> kps;
        for (int f = t; f < nframes; f += nthreads) {
            cv::Mat fe(Ih, Iw, CV_8UC1, (void*)(fisheyes + (size_t)f * Iw * Ih));
            warp_into(canvas, fe, m1, m2);
            ex(canvas, mk, kps, desc);
            totals[t] += (long)kps.size();
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    long s = 0;
    for (long v : totals) s += v;
    return s;
}

// ---- Frame (src/Frame.cpp:104-156): extraction + ComputeKeyPointRays (:746-760) + AssignFeaturesToGrid (:158-176), exactly as Tracking builds it
struct RefFrame { Frame* F; };
void* ref_frame_from_image(void* orb, const uint8_t* img, const uint8_t* mask, int rows, int cols) {
    RefOrb* h = (RefOrb*)orb;
    cv::Mat im(rows, cols, CV_8UC1, (void*)img), mk(rows, cols, CV_8UC1, (void*)mask);
    Frame::mbInitialComputations = true;   // grid geometry follows this image size (src/Frame.cpp:141-152)
    RefFrame* f = new RefFrame; f->F = new Frame(im, mk, 0.0, h->ex, static_cast<ORBVocabulary*>(NULL));
    return f;
}
void ref_frame_destroy(void* fv) { RefFrame* f = (RefFrame*)fv; delete f->F; delete f; }
int ref_frame_n(void* fv) { return ((RefFrame*)fv)->F->N; }
void ref_frame_get(void* fv, cv::KeyPoint* kps, uint8_t* desc, float* rays, int32_t* gridCount) {
    Frame* F = ((RefFrame*)fv)->F;
    const int n = F->N;
    if (n && kps) std::memcpy((void*)kps, F->mvKeys.data(), (size_t)n * sizeof(cv::KeyPoint));
    for (int i = 0; i < n; i++) {
        if (desc) std::memcpy(desc + 32 * (size_t)i, F->mDescriptors.ptr<uchar>(i), 32);
        if (rays) for (int c = 0; c < 3; c++) rays[3 * i + c] = F->mvKeyRays[i](c);
    }
    if (gridCount)
        for (int f = 0; f < CUBEMAP_FACES; f++) for (int x = 0; x < CUBEFACE_GRID_COLS; x++) for (int y = 0; y < CUBEFACE_GRID_ROWS; y++)
            gridCount[(f * CUBEFACE_GRID_COLS + x) * CUBEFACE_GRID_ROWS + y] = (int)F->mGrid[f][x][y].size();
}
// grid cell contents in the reference's order: mGrid[face][col][row] (include/Frame.h:129)
int ref_frame_grid_cell(void* fv, int face, int col, int row, int32_t* idx, int cap) {
    const std::vector<size_t>& v = ((RefFrame*)fv)->F->mGrid[face][col][row];
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) idx[i] = (int32_t)v[i];
    return (int)v.size();
}
// Frame::GetFeaturesInArea (src/Frame.cpp:251-716)
int ref_frame_features_in_area(void* fv, float x, float y, float r, int minLevel, int maxLevel, int32_t* idx, int cap) {
    const std::vector<size_t> v = ((RefFrame*)fv)->F->GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) idx[i] = (int32_t)v[i];
    return (int)v.size();
}

// ORBMatcher::DescriptorDistance, src/ORBMatcher.cpp:951-967
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat ma(1, 32, CV_8UC1, (void*)a), mb(1, 32, CV_8UC1, (void*)b);
    return ORBMatcher::DescriptorDistance(ma, mb);
}

}  // extern "C"
