// ref_driver.cpp -- TEST INFRASTRUCTURE.  Drives the REFERENCE'S OWN compiled functions (oracle/_ref/, built by
// `make -C oracle ref` from /root/reference/dvo_core/src/{dense_tracking_impl,core/math_sse,core/intrinsic_matrix}.cpp,
// unmodified, against the container shims in oracle/ref_shim/) through one Gauss-Newton linearisation exactly as
// DenseTracker::match() strings them together (dense_tracking.cpp:212-220, 271-343), so that the oracle's FAITHFUL
// mode can be checked bit for bit against reference-produced numbers (tests/test_reference_pin.py).
//
// What comes from the reference's object code: computeResidualsSse, computeWeightsSse, computeScaleSse,
// computeCompleteDataLogLikelihood (dense_tracking_impl.cpp), OptimizedSelfAdjointMatrix6x6f::rankUpdate / toEigen
// (math_sse.cpp), IntrinsicMatrix (intrinsic_matrix.cpp), ValidPointAndGradientThresholdPredicate::isPointOk
// (point_selection.h, inline).  What this file restates because its translation unit needs OpenCV / Sophus proper
// (each block cites its lines): the acceleration image (rgbd_image.cpp:534-543), the point cloud
// (rgbd_image.cpp:186-204, 245-262), the point list (point_selection.cpp:119-152), the residual weight vectors
// (dense_tracking.cpp:215-220), the Jacobians (dense_tracking.cpp:448-476) and the three lines of
// NormalEquationsLeastSquares::update/finish (least_squares.cpp:58-64, 74-80).
#include <dvo/dense_tracking_impl.h>

#include <cstdint>
#include <cstring>
#include <vector>

namespace dvo {
namespace core {
// link stubs for the two constructors the driver needs (rgbd_image.cpp:186-204, 320-332; the template cloud is built below)
RgbdCamera::RgbdCamera(size_t width, size_t height, const IntrinsicMatrix& intrinsics) : width_(width), height_(height), intrinsics_(intrinsics) {}
RgbdCamera::~RgbdCamera() {}
RgbdImage::RgbdImage(const RgbdCamera& camera)
    : width(0), height(0), intensity_requires_calculation_(true), depth_requires_calculation_(true), pointcloud_requires_build_(true), camera_(camera) {}
RgbdImage::~RgbdImage() {}
bool RgbdImage::inImage(const float& x, const float& y) const { return x >= 0 && x < width && y >= 0 && y < height; }   // rgbd_image.cpp (only the scalar twin calls it)
}  // namespace core
}  // namespace dvo

using namespace dvo::core;

extern "C" {

// planes6: I, Z, Ix, Iy, Zx, Zy (h*w floats each).  ti / td: DenseTracker::Config::Intensity/DepthDerivativeThreshold.
// Outputs: n_selected, n (valid constraints), residuals (2 floats per valid point, list order), point index (linear
// pixel index of every valid point), weights, P (row-major 2x2), ll, A (row-major 6x6 float), b (6 floats).
int64_t ref_linearize(const float* ref_planes6, const float* cur_planes6, int w, int h, const float K4[4], const double T[16], float ti,
                      float td, int use_weights, const float prev_precision[4], int64_t* n_selected, float* residuals_out,
                      int32_t* index_out, float* records_out, float* weights_out, float P_out[4], float* ll_out, float A_out[36],
                      float b_out[6]) {
  const size_t N = (size_t)w * h;
  IntrinsicMatrix K = IntrinsicMatrix::create(K4[0], K4[1], K4[2], K4[3]);
  RgbdCamera camera(w, h, K);
  RgbdImage current(camera);
  current.width = w; current.height = h;
  // buildAccelerationStructure (rgbd_image.cpp:534-543): cv::merge of {intensity, depth, idx, idy, zdx, zdy, 0, 0}
  current.acceleration = cv::Mat_<RgbdImage::Vec8f>(h, w);
  for (size_t i = 0; i < N; ++i) {
    float* a = current.acceleration.ptr<float>(0) + 8 * i;
    for (int c = 0; c < 6; ++c) a[c] = cur_planes6[c * N + i];
    a[6] = 0.f; a[7] = 0.f;
  }
  // reference point list: point cloud (rgbd_image.cpp:186-204, 245-262) + selectPointsFromImage (point_selection.cpp:119-152)
  ValidPointAndGradientThresholdPredicate predicate;
  predicate.intensity_threshold = ti; predicate.depth_threshold = td;
  PointWithIntensityAndDepth::VectorType points;
  std::vector<int32_t> point_index;
  {
    size_t idx = 0;
    for (size_t y = 0; y < (size_t)h; ++y)
      for (size_t x = 0; x < (size_t)w; ++x, ++idx) {
        const float tx = (x - K.ox()) / K.fx(), ty = (y - K.oy()) / K.fy();
        const float depth = ref_planes6[1 * N + idx];
        PointWithIntensityAndDepth p;
        p.point.data[0] = tx * depth; p.point.data[1] = ty * depth; p.point.data[2] = 1.0f * depth; p.point.data[3] = 1.0f;
        for (int c = 0; c < 6; ++c) p.intensity_and_depth.data[c] = ref_planes6[c * N + idx];
        p.intensity_and_depth.data[6] = 0.f; p.intensity_and_depth.data[7] = 0.f;
        if (predicate.isPointOk(x, y, p.point.z, p.intensity_and_depth.idx, p.intensity_and_depth.idy, p.intensity_and_depth.zdx, p.intensity_and_depth.zdy)) {
          points.push_back(p);
          point_index.push_back((int32_t)idx);
        }
      }
  }
  *n_selected = (int64_t)points.size();
  // dense_tracking.cpp:215-220
  Vector8f wcur, wref;
  float wcur_id = 0.5f, wref_id = 0.5f, wcur_zd = 1.0f, wref_zd = 0.0f;
  wcur <<  1.0f / 255.0f,  1.0f, wcur_id * K.fx() / 255.0f, wcur_id * K.fy() / 255.0f, wcur_zd * K.fx(), wcur_zd * K.fy(), 0.0f, 0.0f;
  wref << -1.0f / 255.0f, -1.0f, wref_id * K.fx() / 255.0f, wref_id * K.fy() / 255.0f, wref_zd * K.fx(), wref_zd * K.fy(), 0.0f, 0.0f;
  // dense_tracking.cpp:263: transformf = estimate.cast<float>()
  Eigen::Affine3f transformf;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) transformf(i, j) = (float)T[i * 4 + j];

  PointWithIntensityAndDepth::VectorType points_error(points.size() + 2);
  dvo::DenseTracker::ResidualVectorType residuals(points.size() + 2);
  dvo::DenseTracker::WeightVectorType weights(points.size() + 2);
  ComputeResidualsResult rr;
  rr.first_point_error = points_error.begin();
  rr.first_residual = residuals.begin();
  computeResidualsSse(points.begin(), points.end(), current, K, transformf, wref, wcur, rr);      // dense_tracking.cpp:271
  const size_t n = rr.last_residual - rr.first_residual;
  Eigen::Vector2f mean; mean.setZero();
  Eigen::Matrix2f precision;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) precision(i, j) = prev_precision ? prev_precision[i * 2 + j] : 0.f;
  if (!use_weights) std::fill(weights.begin(), weights.begin() + n, 1.0f);                         // dense_tracking.cpp:286-293
  else computeWeightsSse(rr.first_residual, rr.last_residual, weights.begin(), mean, precision);
  precision = computeScaleSse(rr.first_residual, rr.last_residual, weights.begin(), mean).inverse();   // :295
  const float ll = computeCompleteDataLogLikelihood(rr.first_residual, rr.last_residual, weights.begin(), mean, precision);   // :297

  // dense_tracking.cpp:327-343 with NormalEquationsLeastSquares::update / finish (least_squares.cpp:58-64, 74-80)
  OptimizedSelfAdjointMatrix6x6f A_opt;
  A_opt.setZero();
  Vector6 b; b.setZero();
  {
    dvo::DenseTracker::WeightVectorType::iterator w_it = weights.begin();
    for (PointIterator e_it = rr.first_point_error; e_it != rr.last_point_error; ++e_it, ++w_it) {
      const float* p = e_it->point.data;
      Matrix2x6 Jw, J;
      Vector6 Jz;
      // computeJacobianOfProjectionAndTransformation (dense_tracking.cpp:448-466)
      NumType z = 1.0f / p[2];
      NumType z_sqr = 1.0f / (p[2] * p[2]);
      Jw(0, 0) = z; Jw(0, 1) = 0.0f; Jw(0, 2) = -p[0] * z_sqr; Jw(0, 3) = Jw(0, 2) * p[1]; Jw(0, 4) = 1.0f - Jw(0, 2) * p[0]; Jw(0, 5) = -p[1] * z;
      Jw(1, 0) = 0.0f; Jw(1, 1) = z; Jw(1, 2) = -p[1] * z_sqr; Jw(1, 3) = -1.0f + Jw(1, 2) * p[1]; Jw(1, 4) = -Jw(0, 3); Jw(1, 5) = p[0] * z;
      // compute3rdRowOfJacobianOfTransformation (dense_tracking.cpp:468-476)
      Jz(0) = 0.0; Jz(1) = 0.0; Jz(2) = 1.0; Jz(3) = p[1]; Jz(4) = -p[0]; Jz(5) = 0.0;
      // J.row(0) = intensity derivative^T * Jw;  J.row(1) = depth derivative^T * Jw - Jz^T   (dense_tracking.cpp:338-339)
      const float* e = e_it->intensity_and_depth.data;
      for (int c = 0; c < 6; ++c) {
        J(0, c) = e[2] * Jw(0, c) + e[3] * Jw(1, c);
        J(1, c) = (e[4] * Jw(0, c) + e[5] * Jw(1, c)) - Jz(c);
      }
      Eigen::Vector2f r(e[0], e[1]);
      Eigen::Matrix2f W = (*w_it) * precision;
      A_opt.rankUpdate(J, W);                        // least_squares.cpp:60 -> math_sse.cpp:82-178
      b -= J.transpose() * W * r;                    // least_squares.cpp:61
    }
  }
  Matrix6x6 A;
  A_opt.toEigen(A);                                  // least_squares.cpp:76

  for (size_t i = 0; i < n; ++i) {
    if (residuals_out) { residuals_out[2 * i] = residuals[i](0); residuals_out[2 * i + 1] = residuals[i](1); }
    if (weights_out) weights_out[i] = weights[i];
    if (records_out) std::memcpy(records_out + 12 * i, &points_error[i], sizeof(float) * 12);
  }
  if (index_out) {
    // computeResidualsSse copies the reference point (x, y, z, 1) into every valid record: recover its pixel by walking
    // the selected list in order (records keep the list order)
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
      while (k < points.size() && std::memcmp(points[k].point.data, points_error[i].point.data, sizeof(float) * 4) != 0) ++k;
      index_out[i] = k < points.size() ? point_index[k] : -1;
      ++k;
    }
  }
  if (P_out) for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) P_out[i * 2 + j] = precision(i, j);
  if (ll_out) *ll_out = ll;
  if (A_out) for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A_out[i * 6 + j] = A(i, j);
  if (b_out) for (int i = 0; i < 6; ++i) b_out[i] = b(i);
  return (int64_t)n;
}

}  // extern "C"
