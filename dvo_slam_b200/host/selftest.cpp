// selftest.cpp -- exercises the adapter exactly the way dvo_benchmark / LocalTracker use dvo_core:
// RgbdCameraPyramid::create -> two DenseTracker::match calls on a shared current image
// (benchmark_slam.cpp:392,483-488; local_tracker.cpp:172-184).  Reads a raw float32 pair written by
// tests/test_host_adapter.py, prints the resulting pose as JSON.  Exit 3 = no CUDA device.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <chrono>
#include <string>
#include <vector>

#include "dvo/dense_tracking.h"
#include "dvo_slam/batched_alignment.h"
#include "dvo_slam/tracking_result_evaluation.h"

namespace {
// the members of dvo_slam's Keyframe / ConstraintProposal that the validator's tracking loop touches
struct Frame {
  dvo::core::RgbdImagePyramidPtr pyramid;
  dvo::core::RgbdImagePyramidPtr image() const { return pyramid; }
};
struct Proposal {
  Frame *Reference, *Current;
  dvo::core::AffineTransformd InitialTransformation;
  dvo::DenseTracker::Result TrackingResult;
};
bool same_pose(const dvo::DenseTracker::Result& a, const dvo::DenseTracker::Result& b) {
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) if (a.Transformation.matrix()(i, j) != b.Transformation.matrix()(i, j)) return false;
  return a.LogLikelihood == b.LogLikelihood;
}
}  // namespace

static cv::Mat load_plane(std::ifstream& f, int w, int h) {
  cv::Mat m(h, w, CV_32FC1);
  f.read(reinterpret_cast<char*>(m.ptr<float>()), sizeof(float) * size_t(w) * h);
  return m;
}

int main(int argc, char** argv) {
  if (argc < 8) { std::fprintf(stderr, "usage: selftest pair.bin w h fx fy ox oy [first last [error_image.bin [batch]]]\n"); return 2; }
  const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
  dvo::core::IntrinsicMatrix K = dvo::core::IntrinsicMatrix::create(float(std::atof(argv[4])), float(std::atof(argv[5])), float(std::atof(argv[6])), float(std::atof(argv[7])));
  std::ifstream f(argv[1], std::ios::binary);
  if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  cv::Mat Ir = load_plane(f, w, h), Zr = load_plane(f, w, h), Ic = load_plane(f, w, h), Zc = load_plane(f, w, h);

  dvo::core::RgbdCameraPyramid camera(w, h, K);
  dvo::core::RgbdImagePyramidPtr reference = camera.create(Ir, Zr), current = camera.create(Ic, Zc);

  dvo::DenseTracker::Config cfg = dvo::DenseTracker::getDefaultConfig();
  cfg.FirstLevel = argc > 8 ? std::atoi(argv[8]) : 3;
  cfg.LastLevel = argc > 9 ? std::atoi(argv[9]) : 1;
  cfg.MaxIterationsPerLevel = 50;
  cfg.Precision = 1e-4;
  dvo::DenseTracker tracker(cfg);
  tracker.collectIterationStatistics(true);
  dvo::DenseTracker::Result result;
  try {
    tracker.match(*reference, *current, result);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  std::printf("{\"T\": [");
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) std::printf("%s%.17g", (i + j) ? ", " : "", result.Transformation.matrix()(i, j));
  std::printf("], \"ll\": %.17g, \"nan\": %d, \"levels\": [", result.LogLikelihood, int(result.isNaN()));
  for (size_t l = 0; l < result.Statistics.Levels.size(); ++l) {
    const dvo::DenseTracker::LevelStats& s = result.Statistics.Levels[l];
    std::printf("%s{\"id\": %zu, \"tc\": %d, \"valid\": %zu, \"its\": %zu, \"n_last\": %zu}", l ? ", " : "", s.Id, int(s.TerminationCriterion),
                s.ValidPixels, s.Iterations.size(), s.Iterations.empty() ? size_t(0) : s.Iterations.back().ValidConstraints);
  }
  // a second tracker on the same current image (LocalTracker runs two), and the error image (N4)
  dvo::DenseTracker second(tracker);
  dvo::core::AffineTransformd guess;
  second.match(*reference, *current, guess);
  cv::Mat err = tracker.computeIntensityErrorImage(*reference, *current, result.Transformation.inverse(), size_t(cfg.LastLevel));
  double esum = 0;
  for (size_t i = 0; i < err.total(); ++i) esum += err.ptr<float>()[i];
  if (argc > 10) {   // the image itself, for the bit-exact comparison with the oracle's raster walk (tests/test_host_adapter.py)
    std::ofstream ef(argv[10], std::ios::binary);
    for (int y = 0; y < err.rows; ++y) ef.write(reinterpret_cast<const char*>(err.ptr<float>(y)), sizeof(float) * size_t(err.cols));
  }
  // N1: the two fan-outs of dvo_slam as one batched call each; the answers must be those of the sequential calls
  dvo::DenseTracker::Result r_keyframe, r_odometry;
  dvo_slam::matchKeyframeAndOdometry(tracker, *reference, *reference, *current, r_keyframe, r_odometry);
  Frame kf = {reference}, fr = {current};
  Proposal p0 = {&kf, &fr, dvo::core::AffineTransformd(), dvo::DenseTracker::Result()}, p1 = p0, p2 = p0;
  std::vector<Proposal*> proposals;
  proposals.push_back(&p0); proposals.push_back(&p1); proposals.push_back(&p2);
  dvo_slam::matchProposals(tracker, proposals);
  const bool batch_equal = same_pose(r_keyframe, result) && same_pose(r_odometry, result);
  const bool proposals_equal = same_pose(p0.TrackingResult, result) && same_pose(p1.TrackingResult, result) && same_pose(p2.TrackingResult, result);
  // N4: keyframe-selection scores
  dvo_slam::EntropyRatioTrackingResultEvaluation entropy(result);
  dvo_slam::LogLikelihoodTrackingResultEvaluation loglik(result);
  dvo_slam::NormalizedLogLikelihoodTrackingResultEvaluation nloglik(result);
  entropy.add(r_keyframe);
  // ... and on DISTINCT results: a coarser alignment (stops one level earlier) and the reverse alignment
  dvo::DenseTracker::Config coarse_cfg = cfg;
  coarse_cfg.LastLevel = cfg.LastLevel + 1;
  dvo::DenseTracker coarse(coarse_cfg);
  coarse.collectIterationStatistics(true);
  dvo::DenseTracker::Result r_coarse, r_reverse;
  coarse.match(*reference, *current, r_coarse);
  tracker.match(*current, *reference, r_reverse);
  const dvo::DenseTracker::Result* rs[3] = {&result, &r_coarse, &r_reverse};
  std::string eval = "[";
  for (int i = 0; i < 3; ++i) {
    char buf[2048];
    int off = std::snprintf(buf, sizeof(buf), "%s{\"ll\": %.17g, \"n_last\": %zu, \"info\": [", i ? ", " : "", rs[i]->LogLikelihood,
                            rs[i]->Statistics.Levels.back().Iterations.back().ValidConstraints);
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) off += std::snprintf(buf + off, sizeof(buf) - size_t(off), "%s%.17g", (a + b) ? ", " : "", rs[i]->Information(a, b));
    std::snprintf(buf + off, sizeof(buf) - size_t(off), "]}");
    eval += buf;
  }
  eval += "]";
  // keyframe_graph.cpp:370-371: condition number of the last accepted iteration's information matrix
  const dvo::DenseTracker::IterationStats& last_inc = result.Statistics.Levels.back().LastIterationWithIncrement();
  const double kappa = last_inc.InformationConditionNumber();
  std::string kappa_info = "[";
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) { char buf[40]; std::snprintf(buf, sizeof(buf), "%s%.17g", (a + b) ? ", " : "", last_inc.EstimateInformation(a, b)); kappa_info += buf; }
  kappa_info += "]";
  // keyframe_tracker.cpp usage pattern: constructed from the first result, add() the following ones, ratios of a later one
  dvo_slam::EntropyRatioTrackingResultEvaluation e2(result);
  dvo_slam::LogLikelihoodTrackingResultEvaluation l2(result);
  dvo_slam::NormalizedLogLikelihoodTrackingResultEvaluation n2(result);
  e2.add(r_coarse); l2.add(r_coarse); n2.add(r_coarse);
  // the C++ batch path timed: `batch` proposals over 2*batch DISTINCT pyramids (uploads batched inside matchBatch), then
  // the same proposals again with the device mirrors in place
  double batch_first_ms = 0, batch_again_ms = 0, batch_max_dev = 0;
  int batch_bitwise = 1;
  int nbatch = argc > 11 ? std::atoi(argv[11]) : 0;
  if (nbatch > 0) {
    std::vector<Frame> frames(size_t(2 * nbatch));
    std::vector<Proposal> props(size_t(nbatch), p0);
    std::vector<Proposal*> pp;
    for (int i = 0; i < nbatch; ++i) {
      frames[size_t(2 * i)].pyramid = camera.create(Ir, Zr);
      frames[size_t(2 * i + 1)].pyramid = camera.create(Ic, Zc);
      props[size_t(i)].Reference = &frames[size_t(2 * i)];
      props[size_t(i)].Current = &frames[size_t(2 * i + 1)];
      pp.push_back(&props[size_t(i)]);
    }
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    dvo_slam::matchProposals(tracker, pp);
    std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    dvo_slam::matchProposals(tracker, pp);
    std::chrono::steady_clock::time_point t2 = std::chrono::steady_clock::now();
    batch_first_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    batch_again_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    // A batch this large is spread over the grid with other squad sizes than a single alignment; sums are taken in an
    // order fixed by the level geometry, so the answers are still those of the single alignment, bit for bit.
    for (int i = 0; i < nbatch; ++i) {
      const dvo::DenseTracker::Result& r = props[size_t(i)].TrackingResult;
      if (!same_pose(r, result)) batch_bitwise = 0;
      if (!same_pose(r, props[0].TrackingResult)) nbatch = -1;
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 4; ++b)
        batch_max_dev = std::max(batch_max_dev, std::fabs(r.Transformation.matrix()(a, b) - result.Transformation.matrix()(a, b)));
      if (nbatch < 0) break;
    }
  }
  std::printf("], \"second_t\": [%.17g, %.17g, %.17g], \"err_sum\": %.9g, \"level1_w\": %d, \"batch_equal\": %d, \"proposals_equal\": %d, "
              "\"entropy_ratio_first\": %.17g, \"entropy_ratio_avg\": %.17g, \"ll_ratio\": %.17g, \"nll_ratio\": %.17g, \"logdet\": %.17g, "
              "\"eval_results\": %s, \"eval\": {\"entropy_first\": %.17g, \"entropy_avg\": %.17g, \"ll_first\": %.17g, \"ll_avg\": %.17g, "
              "\"nll_first\": %.17g, \"nll_avg\": %.17g}, \"batch\": %d, \"batch_bitwise\": %d, \"batch_max_dev\": %.3g, \"batch_first_ms\": %.3f, \"batch_again_ms\": %.3f, \"kappa\": %.17g, \"kappa_info\": %s}\n",
              guess.matrix()(0, 3), guess.matrix()(1, 3), guess.matrix()(2, 3), esum, reference->level(1).intensity.cols,
              int(batch_equal), int(proposals_equal), entropy.ratioWithFirst(r_odometry), entropy.ratioWithAverage(r_odometry),
              loglik.ratioWithFirst(result), nloglik.ratioWithAverage(result), std::log(result.Information.determinant()),
              eval.c_str(), e2.ratioWithFirst(r_reverse), e2.ratioWithAverage(r_reverse), l2.ratioWithFirst(r_reverse), l2.ratioWithAverage(r_reverse),
              n2.ratioWithFirst(r_reverse), n2.ratioWithAverage(r_reverse), nbatch, batch_bitwise, batch_max_dev, batch_first_ms, batch_again_ms, kappa, kappa_info.c_str());
  std::cerr << result.Statistics;
  return 0;
}
