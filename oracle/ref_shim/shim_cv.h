// shim_cv.h -- TEST INFRASTRUCTURE.  Stand-in for the OpenCV 2 containers that the reference headers mention, so
// that the reference's SSE translation units compile unmodified (see shim_eigen.h) and that the adapter's
// -DDVO_B200_WITH_EIGEN_OPENCV branch and the loader of benchmark_slam.cpp can be compiled in this image
// (tests/test_boundary_real_types.py).  cv::Mat / cv::Mat_ are dense row-major 2-D buffers with OpenCV's type codes.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <iostream>   // the real opencv2/opencv.hpp pulls these in; the reference headers rely on it
#include <algorithm>
#include <cmath>
#include <limits>
#include <string>
#include <vector>

typedef unsigned char uchar;

// OpenCV's type codes: depth + ((channels - 1) << 3)
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_BGR2GRAY 6

namespace cv {

template <typename T, int N> struct Vec {
  T val[N];
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;

template <typename T> struct DataDepth;
template <> struct DataDepth<uchar> { enum { value = CV_8U }; };
template <> struct DataDepth<unsigned short> { enum { value = CV_16U }; };
template <> struct DataDepth<float> { enum { value = CV_32F }; };
template <> struct DataDepth<double> { enum { value = CV_64F }; };
template <typename T> struct DataType { enum { type = CV_MAKETYPE(DataDepth<T>::value, 1) }; };
template <typename T, int N> struct DataType<Vec<T, N> > { enum { type = CV_MAKETYPE(DataDepth<T>::value, N) }; };

struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };

// dense row-major 2-D buffer with shared ownership: the part of cv::Mat that the reference's headers, its SSE translation
// units, the loader of benchmark_slam.cpp and the adapter touch
class Mat {
 public:
  int rows, cols;
  size_t step;        // bytes per row
  uchar* data;

  Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8UC1) {}
  Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8UC1) { create(r, c, type); }
  Mat(Size s, int type) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8UC1) { create(s.height, s.width, type); }
  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == type_ && data) return;
    rows = r; cols = c; type_ = type; step = elemSize() * (size_t)c;
    void* p = nullptr;
    if (posix_memalign(&p, 64, step * (size_t)r + 64) != 0) p = nullptr;   // OpenCV aligns a fresh Mat to 16 bytes
    buf_.reset((uchar*)p, free);
    data = buf_.get();
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step * (size_t)r); return m; }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize1() const { const int d = depth(); return d == CV_8U ? 1 : d == CV_16U ? 2 : d == CV_32F ? 4 : 8; }
  size_t elemSize() const { return elemSize1() * (size_t)channels(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t total() const { return (size_t)rows * cols; }
  Size size() const { return Size(cols, rows); }
  Mat clone() const { Mat m(rows, cols, type_); if (data) std::memcpy(m.data, data, step * (size_t)rows); return m; }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
  uchar* ptr(int r = 0) { return data + (size_t)r * step; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <typename T> T* ptr(int r, int c) { return (T*)(data + (size_t)r * step + (size_t)c * elemSize()); }
  template <typename T> const T* ptr(int r, int c) const { return (const T*)(data + (size_t)r * step + (size_t)c * elemSize()); }
  template <typename T> T& at(int r, int c) { return *ptr<T>(r, c); }
  template <typename T> const T& at(int r, int c) const { return *ptr<T>(r, c); }
  template <typename T> T& at(int i) { return ((T*)data)[i]; }
  // Mat::convertTo(dst, rtype): element-wise conversion to another depth, same channel count (no scaling)
  void convertTo(Mat& dst, int rtype) const {
    const int ddepth = rtype & 7, cn = channels();
    Mat out(rows, cols, CV_MAKETYPE(ddepth, cn));
    for (int r = 0; r < rows; ++r)
      for (int i = 0; i < cols * cn; ++i) {
        double v = 0;
        switch (depth()) {
          case CV_8U: v = ptr<uchar>(r)[i]; break;
          case CV_16U: v = ptr<unsigned short>(r)[i]; break;
          case CV_32F: v = ptr<float>(r)[i]; break;
          default: v = ptr<double>(r)[i]; break;
        }
        switch (ddepth) {
          case CV_8U: out.ptr<uchar>(r)[i] = (uchar)v; break;
          case CV_16U: out.ptr<unsigned short>(r)[i] = (unsigned short)v; break;
          case CV_32F: out.ptr<float>(r)[i] = (float)v; break;
          default: out.ptr<double>(r)[i] = v; break;
        }
      }
    dst = out;
  }
 private:
  int type_;
  std::shared_ptr<uchar> buf_;
};

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  static Mat_ zeros(int r, int c) { Mat_ m(r, c); std::memset(m.data, 0, m.step * (size_t)r); return m; }
  T& operator()(int r, int c) { return *Mat::ptr<T>(r, c); }
  const T& operator()(int r, int c) const { return *Mat::ptr<T>(r, c); }
};
typedef Mat_<uchar> Mat1b;
typedef Mat_<float> Mat1f;

// image files are out of scope of the stand-in: an empty Mat, like cv::imread on a missing file
inline Mat imread(const std::string&, int = 1) { return Mat(); }
// cv::cvtColor(src, dst, CV_BGR2GRAY) on CV_8UC3: OpenCV's fixed-point weights (B 1868, G 9617, R 4899, >> 14)
inline void cvtColor(const Mat& src, Mat& dst, int code) {
  (void)code;
  Mat out(src.rows, src.cols, CV_8UC1);
  for (int r = 0; r < src.rows; ++r)
    for (int c = 0; c < src.cols; ++c) {
      const uchar* p = src.ptr<uchar>(r) + 3 * c;
      out.ptr<uchar>(r)[c] = (uchar)((1868 * p[0] + 9617 * p[1] + 4899 * p[2] + 8192) >> 14);
    }
  dst = out;
}

inline int64_t getTickCount() { return 0; }
inline double getTickFrequency() { return 1.0; }

}  // namespace cv
