// dvo/core/surface_pyramid.h -- adapter counterpart of dvo_core/include/dvo/core/surface_pyramid.h:34-50: the raw depth
// conversion the loaders call before RgbdCameraPyramid::create (benchmark_slam.cpp:77, camera_dense_tracking.cpp:235).
// Host code for callers that hold cv::Mat; the batched device path (dvo_b200_pyramid_create_raw_batch) converts on the GPU.
#ifndef DVO_B200_ADAPTER_SURFACE_PYRAMID_H_
#define DVO_B200_ADAPTER_SURFACE_PYRAMID_H_
#include <limits>
#include "datatypes.h"

namespace dvo { namespace core {

class SurfacePyramid {
 public:
  // CV_16UC1 -> CV_32FC1, every pixel times `scale`, 0 -> NaN (surface_pyramid.cpp:45-63)
  static void convertRawDepthImage(const cv::Mat& input, cv::Mat& output, float scale) {
    output.create(input.rows, input.cols, CV_32FC1);
    for (int y = 0; y < input.rows; ++y) {
      const unsigned short* in = input.ptr<unsigned short>(y);
      float* out = output.ptr<float>(y);
      for (int x = 0; x < input.cols; ++x) out[x] = in[x] == 0 ? std::numeric_limits<float>::quiet_NaN() : ((float)in[x]) * scale;
    }
  }
  // the SSE twin (surface_pyramid.cpp:65-105) computes the same values
  static void convertRawDepthImageSse(const cv::Mat& input, cv::Mat& output, float scale) { convertRawDepthImage(input, output, scale); }
  SurfacePyramid() {}
  virtual ~SurfacePyramid() {}
};

} }
#endif
