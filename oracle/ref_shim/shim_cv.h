// shim_cv.h -- TEST INFRASTRUCTURE.  Stand-in for the OpenCV 2 containers that the reference headers mention, so
// that the reference's SSE translation units compile unmodified (see shim_eigen.h).  Only cv::Mat / cv::Mat_ as dense
// row-major 2-D buffers with ptr<T>(row, col) are ever exercised.
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

namespace cv {

template <typename T, int N> struct Vec {
  T val[N];
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;

// type code = element size in bytes (all these translation units need of it)
template <typename T> struct DataType { enum { type = sizeof(T) }; };

class Mat {
 public:
  int rows, cols;
  size_t step;        // bytes per row
  size_t elem;        // bytes per element
  uchar* data;

  Mat() : rows(0), cols(0), step(0), elem(0), data(nullptr) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; elem = (size_t)type; step = elem * (size_t)c;
    void* p = nullptr;
    if (posix_memalign(&p, 64, step * (size_t)r + 64) != 0) p = nullptr;   // OpenCV aligns rows of a fresh Mat to 16 bytes
    buf_.reset((uchar*)p, free);
    data = buf_.get();
  }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step * (size_t)r); return m; }
  bool empty() const { return data == nullptr; }
  size_t total() const { return (size_t)rows * cols; }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
  uchar* ptr(int r = 0) { return data + (size_t)r * step; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <typename T> T* ptr(int r, int c) { return (T*)(data + (size_t)r * step + (size_t)c * elem); }
  template <typename T> const T* ptr(int r, int c) const { return (const T*)(data + (size_t)r * step + (size_t)c * elem); }
  template <typename T> T& at(int r, int c) { return *ptr<T>(r, c); }
  template <typename T> const T& at(int r, int c) const { return *ptr<T>(r, c); }
  template <typename T> T& at(int i) { return ((T*)data)[i]; }
 private:
  std::shared_ptr<uchar> buf_;
};

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, (int)sizeof(T)) {}
  static Mat_ zeros(int r, int c) { Mat_ m(r, c); std::memset(m.data, 0, m.step * (size_t)r); return m; }
  T& operator()(int r, int c) { return *Mat::ptr<T>(r, c); }
  const T& operator()(int r, int c) const { return *Mat::ptr<T>(r, c); }
};
typedef Mat_<uchar> Mat1b;
typedef Mat_<float> Mat1f;

inline int64_t getTickCount() { return 0; }
inline double getTickFrequency() { return 1.0; }

}  // namespace cv
