// compat.h -- the third-party types that appear in the reference's public headers
// (cv::Mat, Eigen::Affine3d / Matrix<double,6,6>, boost::shared_ptr).
//
// With -DDVO_B200_WITH_EIGEN_OPENCV the real libraries are used and the adapter headers in
// include/dvo/ are source compatible with dvo_core's (dvo_core/include/dvo/dense_tracking.h:24-35,
// dvo_core/include/dvo/core/rgbd_image.h:25-33).  This image has neither Eigen nor OpenCV nor Boost,
// so by default minimal stand-ins with the same spelling for the members the hot path touches are
// provided; they exist so the adapter can be compiled and tested here, not to replace those libraries.
#ifndef DVO_B200_COMPAT_H_
#define DVO_B200_COMPAT_H_

#ifdef DVO_B200_WITH_EIGEN_OPENCV

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <boost/shared_ptr.hpp>
#include <opencv2/opencv.hpp>
namespace dvo_b200 { namespace compat {
template <typename T> using shared_ptr = boost::shared_ptr<T>;
} }

#else

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#ifndef CV_8UC1
#define CV_8UC1 0
#define CV_16UC1 2
#define CV_32FC1 5
#endif

namespace cv {
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
// Dense single-channel row-major matrix with shared ownership (the subset of cv::Mat used by
// RgbdImage's public fields: rgbd_image.h:161-176).
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int t) { create(r, c, t); }
  void create(int r, int c, int t) {
    if (r == rows && c == cols && t == type_ && buf_) return;
    rows = r; cols = c; type_ = t;
    buf_.reset(new std::vector<uint8_t>(size_t(r) * c * elemSize(), 0));
  }
  void create(Size s, int t) { create(s.height, s.width, t); }
  static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
  int type() const { return type_; }
  bool empty() const { return !buf_ || rows == 0 || cols == 0; }
  size_t total() const { return size_t(rows) * cols; }
  Size size() const { return Size(cols, rows); }
  size_t elemSize() const { return type_ == CV_8UC1 ? 1 : type_ == CV_16UC1 ? 2 : 4; }
  template <typename T> T* ptr(int y = 0, int x = 0) { return reinterpret_cast<T*>(buf_->data()) + size_t(y) * cols + x; }
  template <typename T> const T* ptr(int y = 0, int x = 0) const { return reinterpret_cast<const T*>(buf_->data()) + size_t(y) * cols + x; }
  template <typename T> T& at(int y, int x) { return *ptr<T>(y, x); }
  template <typename T> const T& at(int y, int x) const { return *ptr<T>(y, x); }
  Mat clone() const { Mat m; m.rows = rows; m.cols = cols; m.type_ = type_; if (buf_) m.buf_.reset(new std::vector<uint8_t>(*buf_)); return m; }
 private:
  int type_ = CV_32FC1;
  std::shared_ptr<std::vector<uint8_t>> buf_;
};
}  // namespace cv

namespace Eigen {
// fixed-size row-major double matrices with the handful of members Result / Config users call
template <int R, int C>
struct MatrixRC {
  double v[R * C];
  MatrixRC() { for (double& x : v) x = 0; }
  double& operator()(int i, int j) { return v[i * C + j]; }
  double operator()(int i, int j) const { return v[i * C + j]; }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  void setZero() { for (double& x : v) x = 0; }
  void setConstant(double c) { for (double& x : v) x = c; }
  void setIdentity() { setZero(); for (int i = 0; i < (R < C ? R : C); ++i) v[i * C + i] = 1; }
  double sum() const { double s = 0; for (double x : v) s += x; return s; }
  double determinant() const {   // partial-pivot LU (square matrices only)
    static_assert(R == C, "determinant of a square matrix");
    double a[R * C];
    for (int i = 0; i < R * C; ++i) a[i] = v[i];
    double det = 1.0;
    for (int k = 0; k < R; ++k) {
      int piv = k;
      for (int i = k + 1; i < R; ++i) if (std::fabs(a[i * C + k]) > std::fabs(a[piv * C + k])) piv = i;
      if (a[piv * C + k] == 0.0) return 0.0;
      if (piv != k) { for (int j = 0; j < C; ++j) { double t = a[k * C + j]; a[k * C + j] = a[piv * C + j]; a[piv * C + j] = t; } det = -det; }
      det *= a[k * C + k];
      for (int i = k + 1; i < R; ++i) {
        const double f = a[i * C + k] / a[k * C + k];
        for (int j = k; j < C; ++j) a[i * C + j] -= f * a[k * C + j];
      }
    }
    return det;
  }
  double* data() { return v; }
  const double* data() const { return v; }
};
typedef MatrixRC<4, 4> Matrix4d;
typedef MatrixRC<3, 3> Matrix3d;
typedef MatrixRC<3, 1> Vector3d;
typedef MatrixRC<2, 1> Vector2d;
typedef MatrixRC<2, 2> Matrix2d;
// Affine3d as used by DenseTracker::Result::Transformation (dense_tracking.h:129)
struct Affine3d {
  Matrix4d m;
  Affine3d() { m.setIdentity(); }
  Matrix4d& matrix() { return m; }
  const Matrix4d& matrix() const { return m; }
  void setIdentity() { m.setIdentity(); }
  static Affine3d Identity() { return Affine3d(); }
  Matrix3d rotation() const { Matrix3d r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = m(i, j); return r; }
  Matrix3d linear() const { return rotation(); }
  Vector3d translation() const { Vector3d t; for (int i = 0; i < 3; ++i) t(i) = m(i, 3); return t; }
  Affine3d operator*(const Affine3d& o) const {
    Affine3d r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += m(i, k) * o.m(k, j); r.m(i, j) = s; }
    return r;
  }
  Affine3d inverse() const {   // rigid-body inverse
    Affine3d r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m(i, j) = m(j, i);
    for (int i = 0; i < 3; ++i) { double s = 0; for (int k = 0; k < 3; ++k) s += m(k, i) * m(k, 3); r.m(i, 3) = -s; }
    return r;
  }
};
}  // namespace Eigen

namespace dvo_b200 { namespace compat {
template <typename T> using shared_ptr = std::shared_ptr<T>;
} }

#endif  // DVO_B200_WITH_EIGEN_OPENCV
#endif  // DVO_B200_COMPAT_H_
