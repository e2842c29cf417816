/* cv_compat.h — the handful of OpenCV core types the reference's hot-path signatures mention, for builds without
 * OpenCV (this container has no OpenCV C++ headers). With -DCSLAM_WITH_OPENCV the real <opencv2/core.hpp> is used and
 * the facade classes compile against it unchanged. Layouts match OpenCV (cv::KeyPoint is 28 bytes, same field order). */
#ifndef CSLAM_CV_COMPAT_H
#define CSLAM_CV_COMPAT_H
#ifdef CSLAM_WITH_OPENCV
#include <opencv2/core.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
namespace cv {
enum { CV_8U_ = 0 };
#ifndef CV_8UC1
#define CV_8UC1 0
#define CV_8U 0
#define CV_32F 5
#endif
struct Point2f { float x, y; Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
/* Minimal owning / non-owning 2-D matrix (continuous or strided rows), enough for image, mask and descriptor I/O. */
class Mat {
public:
    int rows = 0, cols = 0; int flags = 0; size_t step = 0; unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type, void* d, size_t s = 0) : rows(r), cols(c), flags(type), step(s ? s : (size_t)c * elem(type)), data((unsigned char*)d) {}
    void create(int r, int c, int type) {
        rows = r; cols = c; flags = type; step = (size_t)c * elem(type);
        own_.reset(new std::vector<unsigned char>((size_t)r * step)); data = own_->data();
    }
    void release() { own_.reset(); data = nullptr; rows = cols = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return flags; }
    Mat row(int r) const { return Mat(1, cols, flags, data + (size_t)r * step, step); }
    template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    template <class T> T& at(int i) { return ((T*)data)[i]; }
    template <class T> const T& at(int i) const { return ((const T*)data)[i]; }
private:
    static size_t elem(int type) { return type == CV_32F ? 4 : 1; }
    std::shared_ptr<std::vector<unsigned char>> own_;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace cv
#endif
#endif
