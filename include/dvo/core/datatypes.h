// dvo/core/datatypes.h -- adapter counterpart of dvo_core/include/dvo/core/datatypes.h:27-53.
#ifndef DVO_B200_ADAPTER_DATATYPES_H_
#define DVO_B200_ADAPTER_DATATYPES_H_
#include "../../dvo_b200/compat.h"
namespace dvo { namespace core {
typedef float IntensityType;
typedef float DepthType;
static const float InvalidDepth = std::numeric_limits<float>::quiet_NaN();
typedef Eigen::Affine3d AffineTransformd;
#ifdef DVO_B200_WITH_EIGEN_OPENCV
typedef Eigen::Matrix<double, 6, 6> Matrix6d;
typedef Eigen::Matrix<double, 6, 1> Vector6d;
#else
typedef Eigen::MatrixRC<6, 6> Matrix6d;
typedef Eigen::MatrixRC<6, 1> Vector6d;
#endif
} }
#endif
