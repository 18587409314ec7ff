// dvo/core/intrinsic_matrix.h -- adapter counterpart of dvo_core/include/dvo/core/intrinsic_matrix.h:33-62.
#ifndef DVO_B200_ADAPTER_INTRINSIC_MATRIX_H_
#define DVO_B200_ADAPTER_INTRINSIC_MATRIX_H_
namespace dvo { namespace core {
// fx, fy, ox, oy; scale() multiplies the whole matrix, principal point included
// (dvo_core/src/core/intrinsic_matrix.cpp:90-93).
struct IntrinsicMatrix {
  static IntrinsicMatrix create(float fx, float fy, float ox, float oy) { IntrinsicMatrix k; k.fx_ = fx; k.fy_ = fy; k.ox_ = ox; k.oy_ = oy; return k; }
  IntrinsicMatrix() : fx_(0), fy_(0), ox_(0), oy_(0) {}
  float fx() const { return fx_; }
  float fy() const { return fy_; }
  float ox() const { return ox_; }
  float oy() const { return oy_; }
  void invertOffset() { ox_ *= -1; oy_ *= -1; }
  void scale(float factor) { fx_ *= factor; fy_ *= factor; ox_ *= factor; oy_ *= factor; }
 private:
  float fx_, fy_, ox_, oy_;
};
} }
#endif
