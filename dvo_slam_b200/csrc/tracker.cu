// tracker.cu -- dvo::DenseTracker::match() (dvo_core/src/dense_tracking.cpp:131-376) for a batch of
// independent frame pairs, all state on the device.
//
// Per Gauss-Newton iteration the reference makes five passes over the points
// (computeResidualsSse, computeWeightsSse, computeScaleSse, computeCompleteDataLogLikelihood and the
// normal-equation loop, dense_tracking.cpp:271-343).  Precision P_k is a global reduction that the
// log-likelihood and J^T W J depend on, so there are exactly two data-parallel stages (stages.cuh):
//   stage A: warp/interpolate/residual/occlusion test, Student-t weight from P_{k-1}, pairwise scale
//            sums, residual record kept for stage B
//   stage B: log-likelihood terms and the 21+6 normal-equation coefficients with W = w*P_k
// each followed by a small per-pair step (pair_mid_warp: P_k; pair_end_cta: accept test, 6x6 LDL^T
// solve, SE(3) update, termination logic).  match() runs them inside ONE persistent cooperative
// kernel per pyramid level (k_level_persistent); the four plain kernels k_residual / k_pair_mid /
// k_normal / k_pair_end launch the same device functions for the test hooks (residual image, linearize).
#include "common.cuh"
#include "stages.cuh"

#include <cstdio>
#include <cstring>
#include <limits>
#include <cmath>
#include <cstdlib>
#include <algorithm>

namespace dvo_b200 {

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct LevelLaunch {
  int w, h, n, ntiles;   // ntiles = CTAs of a stage launch (4 segments each)
  int nseg;              // warp segments of kSegmentPixels pixels
  unsigned wmagic;       // floor(2^32 / w) + 1: idx / w == __umulhi(idx, wmagic) for idx * w < 2^32
  int level_index;   // position in Result.Statistics.Levels
  int level_id;      // pyramid level
  int max_iterations;
  int first_level;   // 1 for the coarsest level of the match
  int use_initial_estimate;
  double precision, mu;
};

// ------------------------------------------------------------------------------------------------
// per-pair helpers
// ------------------------------------------------------------------------------------------------
__device__ void prepare_iteration(PairState& st, const PairLevel& pl) {
  // dense_tracking.cpp:259-263
  st.inc = se3_exp(st.x);
  st.initial_old = st.initial;
  st.initial = se3_mul(se3_inverse(st.inc), st.initial);
  st.estimate_old = st.estimate;
  st.estimate = se3_mul(st.inc, st.estimate);
  double T[16];
  se3_matrix(st.estimate, T);
  // KT = K * float(T)[0:3,:] in float, reference operation order (dense_tracking_impl.cpp:142-152)
  for (int j = 0; j < 4; ++j) {
    float t0 = (float)T[j], t1 = (float)T[4 + j], t2 = (float)T[8 + j];
    st.kt[j] = __fadd_rn(__fmul_rn(pl.cfx, t0), __fmul_rn(pl.cox, t2));
    st.kt[4 + j] = __fadd_rn(__fmul_rn(pl.cfy, t1), __fmul_rn(pl.coy, t2));
    st.kt[8 + j] = t2;
  }
}

__device__ void log_iteration(dvo_b200_iteration_stats* ilog, int max_log, int pair, PairState& st, int level_id,
                              bool with_increment) {
  if (!ilog || st.iter_log_count >= max_log) { st.iter_log_count++; return; }
  dvo_b200_iteration_stats& e = ilog[(size_t)pair * max_log + st.iter_log_count++];
  e.level = level_id;
  e.id = st.iteration;
  e.valid_constraints = st.n;
  e.tdist_log_likelihood = st.nll_cur;
  for (int i = 0; i < 4; ++i) e.tdist_precision[i] = (double)st.precision[i];
  e.prior_log_likelihood = st.prior_cur;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int i = 0; i < 6; ++i) e.increment[i] = with_increment ? st.x[i] : nan;
  for (int i = 0; i < 36; ++i) e.information[i] = with_increment ? st.A_done[i] : nan;
}

// pls_src: this level's pair descriptors, either in device memory already (pls_src == pls) or in pinned host
// memory that the kernel reads over PCIe (the tracker path: keeps the per-level upload off the H2D copy engine,
// where it would queue behind a bulk image upload of another context).  T_init likewise.
__global__ void k_level_begin(PairState* states, const PairLevel* pls_src, PairLevel* pls, const double* T_init, int npairs,
                              LevelLaunch lp) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  PairState& st = states[p];
  const PairLevel pl = pls_src[p];
  if (pls_src != pls) pls[p] = pl;
  if (lp.first_level) {
    // dense_tracking.cpp:137-150: first increment is the given guess
    st.inc = (lp.use_initial_estimate && T_init) ? se3_from_matrix(T_init + (size_t)p * 16) : se3_identity();
    st.initial = st.inc; st.initial_old = st.inc;
    st.estimate = se3_identity(); st.estimate_old = se3_identity();
    st.num_levels = 0; st.num_iterations_total = 0; st.iter_log_count = 0;
  }
  // dense_tracking.cpp:205-210
  st.precision[0] = st.precision[1] = st.precision[2] = st.precision[3] = 0.f;
  st.iteration = 0;
  st.error = 1.7976931348623157e308;
  st.last_error = st.error;
  st.have_done = 0;
  st.termination = -1;
  st.level_active = 1;
  st.phase_ok = 0;
  LevelSummary& ls = st.levels[lp.level_index];
  ls.id = lp.level_id; ls.termination = -1;
  ls.max_valid_pixels = pl.max_valid_pixels;
  ls.valid_pixels = pl.rsel[0];
  ls.num_iterations = 0; ls.has_inc = 0; ls.last_n = 0; ls.last_inc_n = -1;
  ls.last_inc_nll = __longlong_as_double(0x7ff8000000000000LL);
  st.num_levels = lp.level_index + 1;
  se3_log(st.inc, st.x);  // dense_tracking.cpp:238
  prepare_iteration(st, pl);
}

// ------------------------------------------------------------------------------------------------
// stage A kernel: one warp per 256-pixel segment, four segments per CTA
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSegmentsPerTile * 32)
k_residual(const PairState* __restrict__ states, const PairLevel* __restrict__ pls, float* __restrict__ records,
           float* __restrict__ seg_export, LevelLaunch lp) {
  const int pair = blockIdx.y;
  const PairState& st = states[pair];
  if (!st.level_active) return;
  const PairLevel pl = pls[pair];
  StageConsts c;
  load_stage_consts(st, pl, lp.w, lp.h, c);
  const int warp = threadIdx.x >> 5;
  const int seg = blockIdx.x * kSegmentsPerTile + warp;
  const int begin = min(seg * kSegmentPixels, lp.n);
  const int end = min(begin + kSegmentPixels, lp.n);
  const RecordPlanes rec = record_planes(records + (size_t)pair * kRecordFloatsPerPixel * lp.n, lp.n);
  __shared__ float sm_exp[kSegmentsPerTile][kSegExportFloats];
  stage_a_segment(pl, c, lp.w, lp.wmagic, lp.n, begin, end, rec, sm_exp[warp]);
  __syncthreads();
  if (threadIdx.x == 0) cta_export_segments(sm_exp, seg_export + ((size_t)pair * lp.ntiles + blockIdx.x) * kCtaExportFloats);
}

struct PairMidSmem {
  SegT<double> lanes[32];
  long long lane_base[32];
};

// one warp per pair: combine the segment summaries in order -> covariance -> P_k (dense_tracking.cpp:276-295).
// e: this pair's nseg segment exports; seg_base: this pair's exclusive prefix of valid counts (output).
__device__ __noinline__ void pair_mid_warp(PairState& st, int pair, const float* e, int* seg_base, int ntiles, int* active,
                              const LevelLaunch& lp, dvo_b200_iteration_stats* ilog, int max_log, PairMidSmem& sm) {
  const int lane = threadIdx.x & 31;
  SegT<double>* lanes = sm.lanes;
  long long* lane_base = sm.lane_base;
  int chunk = (ntiles + 31) / 32;
  int t0 = lane * chunk, t1 = min(t0 + chunk, ntiles);
  SegT<double> acc;
  acc.n = 0; acc.wf = acc.wl = 0;
  for (int k = 0; k < 3; ++k) acc.S0[k] = acc.S1[k] = acc.ol[k] = 0;
  for (int t = t0; t < t1; ++t) acc = combine_seg<double>(acc, load_seg_export(e + (size_t)t * kCtaExportFloats));
  lanes[lane] = acc;
  {   // exclusive prefix of the lane counts
    long long incl = acc.n;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      long long v = __shfl_up_sync(kFull, incl, off);
      if (lane >= off) incl += v;
    }
    lane_base[lane] = incl - acc.n;
  }
  __syncwarp();
  for (int off = 1; off < 32; off <<= 1) {   // in-order tree combine
    if ((lane & (2 * off - 1)) == 0) lanes[lane] = combine_seg<double>(lanes[lane], lanes[lane + off]);
    __syncwarp();
  }
  if (lane == 0) {
    SegT<double> all = lanes[0];
    long long n = all.n;
    st.n = n;
    st.n_keep = (n / 50) * 50;
    LevelSummary& ls = st.levels[lp.level_index];
    ls.num_iterations += 1;   // level_stats.Iterations.push_back (dense_tracking.cpp:249)
    ls.last_n = n;
    st.num_iterations_total += 1;
    if (n < 6) {
      // dense_tracking.cpp:276-284
      st.initial = st.initial_old; st.estimate = st.estimate_old;
      st.termination = DVO_B200_TERM_TOO_FEW_CONSTRAINTS;
      st.phase_ok = 0;
      st.nll_cur = 0; st.prior_cur = 0;
      log_iteration(ilog, max_log, pair, st, lp.level_id, false);
      // post-loop checks of dense_tracking.cpp:359-363 still apply
      double m = 0; bool nanx = false;
      for (int i = 0; i < 6; ++i) { m = fmax(m, fabs(st.x[i])); nanx |= st.x[i] != st.x[i]; }
      if (!nanx && m <= lp.precision) st.termination = DVO_B200_TERM_INCREMENT_TOO_SMALL;
      if (st.iteration >= lp.max_iterations) st.termination = DVO_B200_TERM_ITERATIONS_EXCEEDED;
      ls.termination = st.termination;
      ls.has_inc = ls.num_iterations >= 2;   // HasIterationWithIncrement (dense_tracking_config.cpp:138-143)
      if (st.termination != DVO_B200_TERM_TOO_FEW_CONSTRAINTS) ls.has_inc = ls.num_iterations >= 1;
      st.have_done = (st.termination == DVO_B200_TERM_TOO_FEW_CONSTRAINTS) ? -1 : st.have_done;
      st.level_active = 0;
      if (active) atomicSub(active, 1);
    } else {
      // tail term for odd n, normaliser 1/(n-3) (dense_tracking_impl.cpp:596), symmetric 2x2
      double c[3];
      bool tail = ((n - 1) & 1) == 0;
      double s = 1.0 / (double)(n - 3);
      for (int k = 0; k < 3; ++k) c[k] = (all.S0[k] + (tail ? all.wl * all.ol[k] : 0.0)) * s;
      float C0 = (float)c[0], C1 = (float)c[1], C3 = (float)c[2];
      // precision = covariance.inverse() (Eigen 2x2 inverse, dense_tracking.cpp:295), float, unfused
      float det = __fsub_rn(__fmul_rn(C0, C3), __fmul_rn(C1, C1));
      float invdet = __fdiv_rn(1.0f, det);
      for (int i = 0; i < 4; ++i) st.precision_prev[i] = st.precision[i];
      st.precision[0] = __fmul_rn(C3, invdet);
      st.precision[1] = __fmul_rn(-C1, invdet);
      st.precision[2] = __fmul_rn(-C1, invdet);
      st.precision[3] = __fmul_rn(C0, invdet);
      st.phase_ok = 1;
    }
  }
  __syncwarp();
  // exclusive prefix of valid counts per tile (rank base for the log-likelihood tail drop)
  long long run = lane_base[lane];
  for (int t = t0; t < t1; ++t) {
    seg_base[t] = (int)run;
    run += __float_as_int(__ldcg(e + (size_t)t * kCtaExportFloats));
  }
  __syncwarp();
}

__global__ void k_pair_mid(PairState* states, const float* __restrict__ scale_export, int* __restrict__ tile_base,
                           int ntiles, int* active, LevelLaunch lp, dvo_b200_iteration_stats* ilog, int max_log) {
  const int pair = blockIdx.x;
  PairState& st = states[pair];
  if (!st.level_active) return;
  __shared__ PairMidSmem sm;
  pair_mid_warp(st, pair, scale_export + (size_t)pair * ntiles * kCtaExportFloats, tile_base + (size_t)pair * ntiles, ntiles,
                active, lp, ilog, max_log, sm);
}

// ------------------------------------------------------------------------------------------------
// stage B kernel: one warp per segment, CTA-level reduction of the 28 values in a fixed order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSegmentsPerTile * 32)
k_normal(const PairState* __restrict__ states, const PairLevel* __restrict__ pls, const float* __restrict__ records,
         const float* __restrict__ seg_export, const int* __restrict__ seg_base, float* __restrict__ partial,
         LevelLaunch lp) {
  const int pair = blockIdx.y, tile = blockIdx.x;
  const PairState& st = states[pair];
  if (!st.level_active || !st.phase_ok) return;
  const PairLevel pl = pls[pair];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  StageBConsts c;
  load_stage_b_consts(st, c);
  const int seg = tile * kSegmentsPerTile + warp;
  StageBAcc acc;
  stage_b_init(acc);
  if (seg < lp.nseg) {
    const int begin = seg * kSegmentPixels;
    const int end = min(begin + kSegmentPixels, lp.n);
    const RecordPlanes rec = record_planes(const_cast<float*>(records) + (size_t)pair * kRecordFloatsPerPixel * lp.n, lp.n);
    const float* ce = seg_export + ((size_t)pair * lp.ntiles + tile) * kCtaExportFloats;
    long long base = seg_base[(size_t)pair * lp.ntiles + tile];
    for (int k = 0; k < warp; ++k) base += __float_as_int(ce[12 + k]);
    const int cnt = __float_as_int(ce[12 + warp]);
    const bool need_rank = base + cnt > st.n_keep;    // only the segments holding the tail of the point list
    stage_b_segment(pl, c, lp.w, lp.wmagic, lp.n, begin, end, rec, base, st.n_keep, need_rank, acc);
  }
  float v[kNormalValues];
  stage_b_values(acc, v);
  __shared__ float red[kSegmentsPerTile][kNormalValues];
#pragma unroll
  for (int i = 0; i < kNormalValues; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(kFull, v[i], off);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kNormalValues; ++i) red[warp][i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < kNormalValues) {
    float s = 0.f;
    for (int k = 0; k < kSegmentsPerTile; ++k) s += red[k][threadIdx.x];
    partial[((size_t)pair * lp.ntiles + tile) * kNormalValues + threadIdx.x] = s;
  }
}

// End of an iteration (dense_tracking.cpp:297-363): reduce the CTA partials, log-likelihood, accept test, solve,
// pose update, termination.  Every thread of the CTA calls; thread 0 does the scalar part in two steps:
//   critical : everything the other CTAs of the squad wait for -- the new K*T and iteration flag, or
//              level_active = 0 -- followed by `release` (the squad barrier of the persistent kernel);
//   deferred : Revertable bookkeeping, statistics, the iteration log.  It finishes before this CTA arrives at
//              the squad's next barrier, so the next P_k / end step (run by whichever CTA arrives last) sees it.
struct PairEndSmem {
  double part[kSegmentsPerTile][32];
};

template <typename Release>
__device__ __noinline__ void pair_end_cta(PairState& st, const PairLevel& pl, int pair, const float* partial, int ntiles,
                                             int* active, const LevelLaunch& lp, dvo_b200_iteration_stats* ilog, int max_log,
                                             PairEndSmem& sm, Release release) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {   // fp64 sum of the partials: warp q takes tiles q, q+4, ... with independent loads in flight
    double v = 0.0;
    if (lane < kNormalValues) {
      const float* p = partial + lane;
      int t = warp;
      for (; t + 3 * kSegmentsPerTile < ntiles; t += 4 * kSegmentsPerTile) {
        const float a0 = __ldcg(p + (size_t)t * kNormalValues);
        const float a1 = __ldcg(p + (size_t)(t + kSegmentsPerTile) * kNormalValues);
        const float a2 = __ldcg(p + (size_t)(t + 2 * kSegmentsPerTile) * kNormalValues);
        const float a3 = __ldcg(p + (size_t)(t + 3 * kSegmentsPerTile) * kNormalValues);
        v += (double)a0; v += (double)a1; v += (double)a2; v += (double)a3;
      }
      for (; t < ntiles; t += kSegmentsPerTile) v += (double)__ldcg(p + (size_t)t * kNormalValues);
    }
    sm.part[warp][lane] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double vals[kNormalValues];
#pragma unroll
  for (int i = 0; i < kNormalValues; ++i) {
    double v = sm.part[0][i];
#pragma unroll
    for (int q = 1; q < kSegmentsPerTile; ++q) v += sm.part[q][i];
    vals[i] = v;
  }

  // ---- critical ----
  const float P0 = st.precision[0], P1 = st.precision[1], P2 = st.precision[2], P3 = st.precision[3];
  // computeCompleteDataLogLikelihood: 0.5 n log det P - 3.5 sum log(1 + 0.2 d), returned as float
  const float det = __fsub_rn(__fmul_rn(P0, P3), __fmul_rn(P1, P2));
  const float logdet = (float)log((double)det);
  const float ll = (float)(0.5 * (double)st.n * (double)logdet - 0.5 * (5.0 + 2.0) * vals[0]);
  double li[6] = {0, 0, 0, 0, 0, 0};
  double sq = 0;
  if (lp.mu != 0.0) {                              // mu == 0: prior term and the mu*log(initial) shift vanish
    se3_log(st.initial, li);
    for (int i = 0; i < 6; ++i) sq += li[i] * li[i];
  }
  const double last_error = st.error;              // dense_tracking.cpp:306-307
  const double error = -(double)ll;
  const bool accept = error < last_error;          // dense_tracking.cpp:312
  double A[36], bvec[6], x[6];
  {
    int k = 1;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) { A[i * 6 + j] = vals[k]; A[j * 6 + i] = vals[k]; ++k; }
    for (int i = 0; i < 6; ++i) bvec[i] = vals[22 + i];
  }
  int iteration = st.iteration;
  if (accept) {
    double As[36], bs[6];
    for (int i = 0; i < 36; ++i) As[i] = A[i];
    for (int i = 0; i < 6; ++i) { As[i * 6 + i] += lp.mu; bs[i] = bvec[i] + lp.mu * li[i]; }   // lines 345-346
    ldlt_solve6(As, bs, x);                                                                    // line 347
    iteration += 1;                                                                            // line 353
  } else {
    for (int i = 0; i < 6; ++i) x[i] = st.x[i];
  }
  double m = 0; bool nanx = false;
  for (int i = 0; i < 6; ++i) { m = fmax(m, fabs(x[i])); nanx |= x[i] != x[i]; }
  const bool big = !nanx && m > lp.precision;
  const bool exceeded = iteration >= lp.max_iterations;
  const bool level_done = !(accept && big && !exceeded);                                       // line 357
  SE3d inc, estimate_new;
  if (!level_done) {
    // dense_tracking.cpp:259-263 for the next iteration: estimate = exp(x) * estimate, then K * float(T)
    inc = se3_exp(x);
    estimate_new = se3_mul(inc, st.estimate);
    double T[16];
    se3_matrix(estimate_new, T);
    for (int j = 0; j < 4; ++j) {   // reference operation order (dense_tracking_impl.cpp:142-152)
      const float t0 = (float)T[j], t1 = (float)T[4 + j], t2 = (float)T[8 + j];
      st.kt[j] = __fadd_rn(__fmul_rn(pl.cfx, t0), __fmul_rn(pl.cox, t2));
      st.kt[4 + j] = __fadd_rn(__fmul_rn(pl.cfy, t1), __fmul_rn(pl.coy, t2));
      st.kt[8 + j] = t2;
    }
    st.iteration = iteration;
  } else {
    st.level_active = 0;
  }
  release();

  // ---- deferred ----
  LevelSummary& ls = st.levels[lp.level_index];
  st.ll = ll;
  st.nll_cur = -(double)ll;
  st.prior_cur = lp.mu * sq;                       // dense_tracking.cpp:302
  st.last_error = last_error;
  st.error = error;
  for (int i = 0; i < 36; ++i) st.A[i] = A[i];
  for (int i = 0; i < 6; ++i) st.b[i] = bvec[i];
  if (!accept) {
    st.initial = st.initial_old; st.estimate = st.estimate_old;   // dense_tracking.cpp:314-321
    st.termination = DVO_B200_TERM_LOG_LIKELIHOOD_DECREASED;
    log_iteration(ilog, max_log, pair, st, lp.level_id, false);
  } else {
    for (int i = 0; i < 6; ++i) st.x[i] = x[i];
    for (int i = 0; i < 36; ++i) st.A_done[i] = A[i];
    for (int i = 0; i < 6; ++i) st.A_done[i * 6 + i] += lp.mu;
    st.nll_done = st.nll_cur; st.prior_done = st.prior_cur; st.have_done = 1;
    ls.last_inc_n = st.n; ls.last_inc_nll = st.nll_cur;
    log_iteration(ilog, max_log, pair, st, lp.level_id, true);
    st.iteration = iteration;
  }
  if (level_done) {
    if (!nanx && m <= lp.precision) st.termination = DVO_B200_TERM_INCREMENT_TOO_SMALL;       // line 359
    if (exceeded) st.termination = DVO_B200_TERM_ITERATIONS_EXCEEDED;                         // line 362
    ls.termination = st.termination;
    int need = (st.termination == DVO_B200_TERM_LOG_LIKELIHOOD_DECREASED ||
                st.termination == DVO_B200_TERM_TOO_FEW_CONSTRAINTS) ? 2 : 1;
    ls.has_inc = ls.num_iterations >= need;
    if (active) atomicSub(active, 1);
  } else {
    st.inc = inc;
    st.initial_old = st.initial;
    st.initial = se3_mul(se3_inverse(inc), st.initial);
    st.estimate_old = st.estimate;
    st.estimate = estimate_new;
  }
}

__global__ void __launch_bounds__(kSegmentsPerTile * 32)
k_pair_end(PairState* states, const PairLevel* pls, const float* __restrict__ partial, int ntiles,
           int* active, LevelLaunch lp, dvo_b200_iteration_stats* ilog, int max_log) {
  const int pair = blockIdx.x;
  PairState& st = states[pair];
  if (!st.level_active || !st.phase_ok) return;
  __shared__ PairEndSmem sm;
  pair_end_cta(st, pls[pair], pair, partial + (size_t)pair * ntiles * kNormalValues, ntiles, active, lp, ilog, max_log, sm, [] {});
}

// ------------------------------------------------------------------------------------------------
// One persistent cooperative kernel per pyramid level.
//
// The grid is num_sms x C CTAs (C = resident CTAs per SM).  CTAs are grouped into squads of g CTAs
// (g chosen per level so that a warp walks >= ~8 rounds of 32 pixels); a squad owns ONE frame pair at
// a time and runs all its Gauss-Newton iterations on this level inside the kernel:
//   stage A over the squad's warp segments -> squad barrier, the last CTA to arrive computes P_k
//   (pair_mid_warp) -> stage B -> squad barrier, the last CTA reduces the partials, tests the
//   log-likelihood, solves the 6x6 system and updates the pose (pair_end_cta) -> next iteration,
// then takes the next pair from a global queue.  The residual records of the pair in flight live in a
// per-squad scratch buffer that is rewritten every iteration and therefore stays in L2; the squads of
// different resident-CTA slots share each SM, so one squad's barrier wait is hidden by the others.
// ------------------------------------------------------------------------------------------------
#ifndef DVO_PERSISTENT_CTAS_PER_SM
#define DVO_PERSISTENT_CTAS_PER_SM 4   // resident 128-thread CTAs per SM the register budget is tuned for (128 regs/thread)
#endif

struct SquadState {
  int pair;
  unsigned arrive;
  unsigned phase;
  int pad_[29];   // one 128-byte line per squad
};

struct PersistentArgs {
  PairState* states;
  const PairLevel* pls;
  float* records;
  float* seg_export;
  int* seg_base;
  float* partial;
  SquadState* squads;
  int* next_pair;
  int* error_flag;
  dvo_b200_iteration_stats* ilog;
  int max_log;
  unsigned long long* dbg;   // optional: ns spent per CTA in {stage A, stage B, wait A, wait B, mid, end, queue, total}
  int npairs, g, squads_per_slot, num_sms, rpw, nseg;
  LevelLaunch lp;
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// returns true in every thread of the CTA that arrived last at barrier episode `episode`
__device__ __forceinline__ bool squad_arrive(SquadState* sq, unsigned episode, int g, int* s_flag) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned old = atomicAdd(&sq->arrive, 1u);
    int last = old == (episode + 1u) * (unsigned)g - 1u;
    if (last) __threadfence();
    s_flag[0] = last;
  }
  __syncthreads();
  return s_flag[0] != 0;
}
__device__ __forceinline__ void squad_release(SquadState* sq, unsigned episode) {
  __threadfence();
  atomicExch(&sq->phase, episode + 1u);
}
__device__ __forceinline__ void squad_wait(SquadState* sq, unsigned episode, int* error_flag) {
  if (threadIdx.x == 0) {
    unsigned spins = 0;
    while (ld_acquire_u32(&sq->phase) < episode + 1u) {
      __nanosleep(200);
      if (((++spins) & 4095u) == 0u) {
        if (*reinterpret_cast<volatile int*>(error_flag)) break;
        if (spins > (1u << 25)) { atomicExch(error_flag, 1); break; }   // ~ seconds: never hang the GPU
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSegmentsPerTile * 32, DVO_PERSISTENT_CTAS_PER_SM)
k_level_persistent(PersistentArgs a) {
  const int slot = blockIdx.x / a.num_sms, smi = blockIdx.x % a.num_sms;
  const int sq_in_slot = smi / a.g;
  if (sq_in_slot >= a.squads_per_slot) return;   // leftover CTAs of this slot
  const int squad = slot * a.squads_per_slot + sq_in_slot;
  const int rank = smi - sq_in_slot * a.g;
  SquadState* sq = a.squads + squad;
  const LevelLaunch& lp = a.lp;
  float* rec_base = a.records + (size_t)squad * kRecordFloatsPerPixel * lp.n;
  float* exports = a.seg_export + (size_t)squad * a.g * kCtaExportFloats;
  int* segbase = a.seg_base + (size_t)squad * a.g;
  float* partial = a.partial + (size_t)squad * a.g * kNormalValues;
  const RecordPlanes rec = record_planes(rec_base, lp.n);

  __shared__ PairMidSmem sm_mid;
  __shared__ PairEndSmem sm_end;
  __shared__ float red[kSegmentsPerTile][kNormalValues];
  __shared__ float sm_exp[kSegmentsPerTile][kSegExportFloats];
  __shared__ int s_flag[2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wk = rank * kSegmentsPerTile + warp;           // this warp's segment index inside the squad
  const int R = (lp.n + 31) >> 5;
  const int r0 = min(wk * a.rpw, R), r1 = min(r0 + a.rpw, R);
  const int begin = r0 * 32, end = min(r1 * 32, lp.n);
  unsigned episode = 0;
  unsigned long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1 = 0;
  const bool timing = a.dbg != nullptr && threadIdx.x == 0;
  const unsigned long long t_start = timing ? global_ns() : 0;
#define DVO_TICK() do { if (timing) t0 = global_ns(); } while (0)
#define DVO_TOCK(slot) do { if (timing) { t1 = global_ns(); t_acc[slot] += t1 - t0; t0 = t1; } } while (0)

  for (;;) {
    DVO_TICK();
    // ---- take the next pair from the queue (the last CTA to arrive does it for the squad) ----
    if (squad_arrive(sq, episode, a.g, s_flag)) {
      if (threadIdx.x == 0) {
        int p = atomicAdd(a.next_pair, 1);
        sq->pair = p < a.npairs ? p : -1;
        squad_release(sq, episode);
      }
    } else {
      squad_wait(sq, episode, a.error_flag);
    }
    __syncthreads();
    ++episode;
    DVO_TOCK(6);
    const int pair = __ldcg(&sq->pair);
    if (pair < 0 || *reinterpret_cast<volatile int*>(a.error_flag)) break;
    PairState& st = a.states[pair];
    const PairLevel pl = a.pls[pair];

    for (;;) {
      // ---- stage A ----
      {
        StageConsts c;
        load_stage_consts(st, pl, lp.w, lp.h, c);
        stage_a_segment(pl, c, lp.w, lp.wmagic, lp.n, begin, end, rec, sm_exp[warp]);
        __syncthreads();
        if (threadIdx.x == 0) cta_export_segments(sm_exp, exports + (size_t)rank * kCtaExportFloats);
      }
      DVO_TOCK(0);
      if (squad_arrive(sq, episode, a.g, s_flag)) {
        DVO_TOCK(2);
        if (warp == 0) {
          pair_mid_warp(st, pair, exports, segbase, a.g, nullptr, lp, a.ilog, a.max_log, sm_mid);
          if (lane == 0) squad_release(sq, episode);
        }
        __syncthreads();
        DVO_TOCK(4);
      } else {
        squad_wait(sq, episode, a.error_flag);
        DVO_TOCK(2);
      }
      __syncthreads();
      ++episode;
      if (!__ldcg(&st.level_active) || *reinterpret_cast<volatile int*>(a.error_flag)) break;   // too few constraints

      // ---- stage B ----
      {
        StageBConsts cb;
        load_stage_b_consts(st, cb);
        StageBAcc acc;
        stage_b_init(acc);
        const long long n_keep = __ldcg(&st.n_keep);
        long long base = __ldcg(&segbase[rank]);
        for (int k = 0; k < warp; ++k) base += __float_as_int(sm_exp[k][0]);
        const int cnt = __float_as_int(sm_exp[warp][0]);
        stage_b_segment(pl, cb, lp.w, lp.wmagic, lp.n, begin, end, rec, base, n_keep, base + cnt > n_keep, acc);
        float v[kNormalValues];
        stage_b_values(acc, v);
#pragma unroll
        for (int i = 0; i < kNormalValues; ++i) {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(kFull, v[i], off);
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < kNormalValues; ++i) red[warp][i] = v[i];
        }
        __syncthreads();
        if (threadIdx.x < kNormalValues) {
          float s = 0.f;
          for (int k = 0; k < kSegmentsPerTile; ++k) s += red[k][threadIdx.x];
          partial[(size_t)rank * kNormalValues + threadIdx.x] = s;
        }
      }
      DVO_TOCK(1);
      if (squad_arrive(sq, episode, a.g, s_flag)) {
        DVO_TOCK(3);
        pair_end_cta(st, pl, pair, partial, a.g, nullptr, lp, a.ilog, a.max_log, sm_end, [&] { squad_release(sq, episode); });
        __syncthreads();
        DVO_TOCK(5);
      } else {
        squad_wait(sq, episode, a.error_flag);
        DVO_TOCK(3);
      }
      __syncthreads();
      ++episode;
      if (!__ldcg(&st.level_active) || *reinterpret_cast<volatile int*>(a.error_flag)) break;
    }
  }
  if (timing) {
    t_acc[7] = global_ns() - t_start;
    for (int i = 0; i < 8; ++i) atomicAdd(a.dbg + i, t_acc[i]);
  }
#undef DVO_TICK
#undef DVO_TOCK
}

// Result assembly (dense_tracking.cpp:368-373)
__global__ void k_finalize(PairState* states, dvo_b200_result* results, int npairs) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  PairState& st = states[p];
  dvo_b200_result& r = results[p];
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  se3_matrix(se3_inverse(st.estimate), r.transformation);
  // last_iteration = Iterations[size-1] unless LogLikelihoodDecreased (then size-2).  With
  // TooFewConstraints on the last level, or no completed iteration, the reference reads an
  // uninitialised / out-of-range element; defined here as NaN so Result::isNaN() fires.
  bool ok = st.have_done == 1 && st.termination != DVO_B200_TERM_TOO_FEW_CONSTRAINTS;
  for (int i = 0; i < 36; ++i) r.information[i] = ok ? st.A_done[i] * 0.008 * 0.008 : nan;
  r.log_likelihood = ok ? st.nll_done + st.prior_done : nan;
  r.num_levels = st.num_levels;
  r.num_iterations_total = st.num_iterations_total;
  for (int l = 0; l < kMaxLevels; ++l) {
    dvo_b200_level_stats& o = r.levels[l];
    if (l < st.num_levels) {
      const LevelSummary& s = st.levels[l];
      o.id = s.id; o.termination = s.termination; o.max_valid_pixels = s.max_valid_pixels;
      o.valid_pixels = s.valid_pixels; o.num_iterations = s.num_iterations;
      o.has_iteration_with_increment = s.has_inc; o.last_valid_constraints = s.last_n;
      o.last_increment_valid_constraints = s.last_inc_n; o.last_increment_log_likelihood = s.last_inc_nll;
    } else {
      o.id = -1; o.termination = -1; o.max_valid_pixels = 0; o.valid_pixels = 0; o.num_iterations = 0;
      o.has_iteration_with_increment = 0; o.last_valid_constraints = 0; o.last_increment_valid_constraints = -1;
      o.last_increment_log_likelihood = nan;
    }
  }
}

// test hook: place a fixed transform / precision / iteration flag into the state (no exp/log chain)
__global__ void k_set_state(PairState* states, const PairLevel* pls, const double* T, const float* prev_precision,
                            int use_weights, LevelLaunch lp) {
  PairState& st = states[0];
  const PairLevel& pl = pls[0];
  st.estimate = se3_from_matrix(T); st.estimate_old = st.estimate;
  st.initial = se3_identity(); st.initial_old = st.initial; st.inc = se3_identity();
  for (int i = 0; i < 6; ++i) st.x[i] = 0;
  st.iteration = use_weights ? 1 : 0;
  for (int i = 0; i < 4; ++i) st.precision[i] = use_weights ? prev_precision[i] : 0.f;
  st.error = 1.7976931348623157e308; st.last_error = st.error;
  st.level_active = 1; st.phase_ok = 0; st.have_done = 0; st.termination = -1;
  st.num_levels = 1; st.num_iterations_total = 0; st.iter_log_count = 0;
  LevelSummary& ls = st.levels[0];
  ls.id = lp.level_id; ls.num_iterations = 0; ls.valid_pixels = pl.rsel[0]; ls.max_valid_pixels = pl.max_valid_pixels;
  double Tm[16];
  se3_matrix(st.estimate, Tm);
  // take the matrix exactly as given (the round trip through the quaternion is not bit exact)
  for (int j = 0; j < 4; ++j) {
    float t0 = (float)T[j], t1 = (float)T[4 + j], t2 = (float)T[8 + j];
    st.kt[j] = __fadd_rn(__fmul_rn(pl.cfx, t0), __fmul_rn(pl.cox, t2));
    st.kt[4 + j] = __fadd_rn(__fmul_rn(pl.cfy, t1), __fmul_rn(pl.coy, t2));
    st.kt[8 + j] = t2;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
inline int level_flag_slot(int li) { return li < 8 ? li : 7; }
template <typename T>
int grow(dvo_b200_ctx* ctx, T*& ptr, size_t& cap, size_t need) {
  if (need <= cap) return 0;
  if (ptr) { cudaStreamSynchronize(ctx->stream); cudaFree(ptr); ptr = nullptr; cap = 0; }
  DVO_CUDA(ctx, cudaMalloc((void**)&ptr, need * sizeof(T)));
  cap = need;
  return 0;
}

struct ScratchNeed {
  size_t record_floats = 0, export_floats = 0, segbase_ints = 0, partial_floats = 0, squads = 0;
};

int ensure_workspace(dvo_b200_ctx* ctx, int npairs, const ScratchNeed& need, int max_log_per_pair) {
  Workspace& ws = ctx->ws;
  if ((size_t)npairs > ws.cap_pairs) {
    if (ws.d_pair_level) { cudaStreamSynchronize(ctx->stream); cudaFree(ws.d_pair_level); cudaFree(ws.d_state); }
    ws.d_pair_level = nullptr; ws.d_state = nullptr; ws.cap_pairs = 0;
    DVO_CUDA(ctx, cudaMalloc((void**)&ws.d_pair_level, sizeof(PairLevel) * npairs));
    DVO_CUDA(ctx, cudaMalloc((void**)&ws.d_state, sizeof(PairState) * npairs));
    ws.cap_pairs = npairs;
  }
  int rc;
  if ((rc = grow(ctx, ws.d_records, ws.cap_records, need.record_floats))) return rc;
  if ((rc = grow(ctx, ws.d_scale_export, ws.cap_export, need.export_floats))) return rc;
  if ((rc = grow(ctx, ws.d_tile_base, ws.cap_segbase, need.segbase_ints))) return rc;
  if ((rc = grow(ctx, ws.d_normal_partial, ws.cap_partial, need.partial_floats))) return rc;
  if ((rc = grow(ctx, ws.d_squads, ws.cap_squads, need.squads * sizeof(SquadState)))) return rc;
  if (!ws.d_active) {
    DVO_CUDA(ctx, cudaMalloc((void**)&ws.d_active, sizeof(int) * 8));
    DVO_CUDA(ctx, cudaMallocHost((void**)&ws.h_active, sizeof(int) * 8));
  }
  if (max_log_per_pair > 0) {
    size_t n = (size_t)npairs * max_log_per_pair;
    if ((rc = grow(ctx, ws.d_iter_log, ws.cap_iter_log, n))) return rc;
  }
  return 0;
}

// How one pyramid level is spread over the persistent grid.
struct LevelPlan {
  int g;                // CTAs per squad
  int squads_per_slot;  // num_sms / g
  int nsquads;          // ctas_per_sm * squads_per_slot
  int rpw;              // rounds of 32 pixels per warp
  int nseg;             // warp segments per pair = g * kSegmentsPerTile
};

LevelPlan plan_level(int n, int num_sms, int ctas_per_sm, int npairs) {
  LevelPlan p;
  const int R = (n + 31) / 32;
  const int warps = kSegmentsPerTile;
  // Squad size.  Small squads keep many pairs in flight and amortise the two barriers and the serial P_k / solve
  // sections of an iteration over long warp segments (`target` rounds of 32 pixels per warp and stage); but the
  // batch is processed in waves of nsquads pairs, and a last wave that is mostly empty wastes more than that.
  // Pairs are handed out from a queue, so a level takes about (pairs per squad + tail) x time per pair, where the
  // tail (pairs that need two or three times the mean number of iterations) is worth a bit more than one pair
  // and the time per pair goes with (rounds per warp + per-iteration overhead in round units).  Pick the number
  // of squads per resident-CTA slot that minimises that among the sizes near the target.
  static const int target = [] { const char* e = getenv("DVO_B200_RPW"); int v = e ? atoi(e) : 0; return v > 0 ? v : 128; }();
  const int overhead = 16;
  const int k_max = std::max(1, (npairs + ctas_per_sm - 1) / ctas_per_sm);   // few pairs: fewer, larger squads (latency)
  int best_k = 1;
  double best_cost = -1.0;
  for (int k = 1; k <= std::min(num_sms, k_max); ++k) {
    const int g = num_sms / k;
    const int rpw = (R + g * warps - 1) / (g * warps);
    if (rpw > target + target / 2 && k > 1) break;    // rpw grows with k
    const double per_squad = (double)npairs / ((double)ctas_per_sm * k);
    const double cost = (std::max(per_squad, 1.0) + 1.2) * (rpw + overhead);
    if (best_cost < 0 || cost < best_cost - 1e-9) { best_cost = cost; best_k = k; }
  }
  const int k = best_k;
  int g = num_sms / k;
  p.g = g;
  p.squads_per_slot = k;
  p.nsquads = ctas_per_sm * p.squads_per_slot;
  p.rpw = (R + g * warps - 1) / (g * warps);
  p.nseg = g * warps;
  return p;
}

int check_batch(dvo_b200_ctx* ctx, const dvo_b200_config* cfg, int n, dvo_b200_pyramid* const* refs,
                dvo_b200_pyramid* const* curs) {
  if (!cfg || n <= 0 || !refs || !curs) return set_error(ctx, DVO_B200_ERR_INVALID_ARGUMENT, "match: null argument");
  if (cfg->first_level < cfg->last_level || cfg->last_level < 0 || cfg->first_level >= kMaxLevels)
    return set_error(ctx, DVO_B200_ERR_INVALID_ARGUMENT, "match: config not sane (FirstLevel >= LastLevel >= 0 required)");
  if (cfg->max_iterations_per_level < 0) return set_error(ctx, DVO_B200_ERR_INVALID_ARGUMENT, "match: max iterations < 0");
  for (int i = 0; i < n; ++i) {
    if (!refs[i] || !curs[i]) return set_error(ctx, DVO_B200_ERR_INVALID_ARGUMENT, "match: null pyramid");
    // a pyramid built on another ctx's stream: order this stream after its build
    for (const dvo_b200_pyramid* p : {refs[i], curs[i]})
      if (p->slab && p->slab->pool != ctx->pool && p->slab->ready) cudaStreamWaitEvent(ctx->stream, p->slab->ready, 0);
    if (refs[i]->levels <= cfg->first_level || curs[i]->levels <= cfg->first_level)
      return set_error(ctx, DVO_B200_ERR_INVALID_ARGUMENT, "match: pyramid has fewer levels than FirstLevel+1");
    if (refs[i]->L[0].w != refs[0]->L[0].w || refs[i]->L[0].h != refs[0]->L[0].h ||
        curs[i]->L[0].w != refs[0]->L[0].w || curs[i]->L[0].h != refs[0]->L[0].h)
      return set_error(ctx, DVO_B200_ERR_SHAPE_MISMATCH, "match: all pyramids of a batch must share width/height");
  }
  return 0;
}

void fill_pair_levels(PairLevel* h, int n, dvo_b200_pyramid* const* refs, dvo_b200_pyramid* const* curs, int level) {
  for (int i = 0; i < n; ++i) {
    const dvo_b200_pyramid* r = refs[i];
    const dvo_b200_pyramid* c = curs[i];
    const LevelInfo& rl = r->L[level];
    const LevelInfo& cl = c->L[level];
    PairLevel& q = h[i];
    q.r0 = r->planes + rl.plane_off; q.r1 = q.r0 + rl.n;
    q.rmask = r->sel_mask + rl.mask_off;
    q.rsel = r->sel_info + 2 * level;
    q.rtmpl = r->tmpl + rl.tmpl_off;
    q.c0 = c->planes + cl.plane_off; q.c1 = q.c0 + cl.n; q.c2 = q.c1 + cl.n;
    q.cfx = cl.fx; q.cfy = cl.fy; q.cox = cl.ox; q.coy = cl.oy;
    // PointSelection::getMaximumNumberOfPoints (point_selection.cpp:68-71)
    q.max_valid_pixels = (long long)(size_t)((double)r->L[0].n * pow(0.25, (double)level));
  }
}

int upload_pair_levels(dvo_b200_ctx* ctx, int n, dvo_b200_pyramid* const* refs, dvo_b200_pyramid* const* curs, int level) {
  size_t bytes = sizeof(PairLevel) * n;
  int rc = ensure_stage(ctx, 0, bytes);
  if (rc) return rc;
  DVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // previous use of the pinned stage has drained
  fill_pair_levels((PairLevel*)ctx->h_stage, n, refs, curs, level);
  DVO_CUDA(ctx, cudaMemcpyAsync(ctx->ws.d_pair_level, ctx->h_stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->h2d_bytes += bytes;
  return 0;
}

}  // namespace

int tracker_match_batch(dvo_b200_ctx* ctx, const dvo_b200_config* cfg, int n, dvo_b200_pyramid* const* refs,
                        dvo_b200_pyramid* const* curs, const double* T_init, dvo_b200_result* h_results,
                        void* d_results_user, dvo_b200_iteration_stats* iter_stats, int max_iter_stats) {
  int rc = check_batch(ctx, cfg, n, refs, curs);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  Workspace& ws = ctx->ws;
  const int last = cfg->last_level, first = cfg->first_level;
  const int max_log = iter_stats ? max_iter_stats : 0;
  // persistent grid geometry
  if (ctx->num_sms == 0) {
    cudaDeviceProp prop;
    DVO_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    ctx->num_sms = prop.multiProcessorCount;
    int per_sm = 0;
    DVO_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_level_persistent, kSegmentsPerTile * 32, 0));
    if (per_sm < 1) return set_error(ctx, DVO_B200_ERR_CUDA, "persistent kernel does not fit on an SM");
    ctx->ctas_per_sm = per_sm;
  }
  ScratchNeed need;
  for (int level = first; level >= last; --level) {
    const int nl = refs[0]->L[level].n;
    LevelPlan pl = plan_level(nl, ctx->num_sms, ctx->ctas_per_sm, n);
    need.record_floats = std::max(need.record_floats, (size_t)pl.nsquads * kRecordFloatsPerPixel * nl);
    need.export_floats = std::max(need.export_floats, (size_t)pl.nsquads * pl.g * kCtaExportFloats);
    need.segbase_ints = std::max(need.segbase_ints, (size_t)pl.nsquads * pl.g);
    need.partial_floats = std::max(need.partial_floats, (size_t)pl.nsquads * pl.g * kNormalValues);
    need.squads = std::max(need.squads, (size_t)pl.nsquads + 1);
  }
  rc = ensure_workspace(ctx, n, need, max_log);
  if (rc) return rc;

  // selection masks for non-default thresholds (PointSelection caches per pyramid, point_selection.cpp:100-113)
  for (int i = 0; i < n; ++i)
    if ((rc = pyramid_reselect(ctx, refs[i], cfg->intensity_derivative_threshold, cfg->depth_derivative_threshold))) return rc;

  // Pair descriptors of every level and the initial estimates go into the pinned stage once; k_level_begin
  // reads them from there (host-mapped, unified addressing), so the level loop below never touches the H2D copy
  // engine or the host between launches.
  const int nlev = first - last + 1;
  const size_t desc_bytes = sizeof(PairLevel) * (size_t)n * nlev;
  const bool have_init = cfg->use_initial_estimate && T_init;
  const size_t init_bytes = have_init ? sizeof(double) * 16 * (size_t)n : 0;
  if ((rc = ensure_stage(ctx, 0, desc_bytes + init_bytes))) return rc;
  DVO_CUDA(ctx, cudaStreamSynchronize(st));   // previous use of the pinned stage has drained
  PairLevel* h_desc = (PairLevel*)ctx->h_stage;
  for (int level = first, li = 0; level >= last; --level, ++li) fill_pair_levels(h_desc + (size_t)li * n, n, refs, curs, level);
  double* d_Tinit = nullptr;
  if (have_init) {
    d_Tinit = (double*)((char*)ctx->h_stage + desc_bytes);
    std::memcpy(d_Tinit, T_init, init_bytes);
  }
  ctx->h2d_bytes += desc_bytes + init_bytes;

  for (int i = 0; i < 8; ++i) ws.h_active[i] = 0;
  for (int level = first, li = 0; level >= last; --level, ++li) {
    const LevelInfo& L = refs[0]->L[level];
    LevelLaunch lp;
    lp.w = L.w; lp.h = L.h; lp.n = L.n; lp.nseg = (L.n + kSegmentPixels - 1) / kSegmentPixels; lp.ntiles = (lp.nseg + kSegmentsPerTile - 1) / kSegmentsPerTile;
    lp.wmagic = (unsigned)((1ull << 32) / (unsigned)L.w) + 1u;
    lp.level_index = li; lp.level_id = level; lp.max_iterations = cfg->max_iterations_per_level;
    lp.first_level = li == 0; lp.use_initial_estimate = cfg->use_initial_estimate;
    lp.precision = cfg->precision; lp.mu = cfg->mu;
    const LevelPlan plan = plan_level(L.n, ctx->num_sms, ctx->ctas_per_sm, n);
    // squad states, queue head and error flag (last SquadState slot) start at zero
    DVO_CUDA(ctx, cudaMemsetAsync(ws.d_squads, 0, sizeof(SquadState) * (plan.nsquads + 1), st));
    {
      ProfScope prof(ctx, 2);
      k_level_begin<<<(n + 63) / 64, 64, 0, st>>>(ws.d_state, h_desc + (size_t)li * n, ws.d_pair_level, d_Tinit, n, lp);
      ctx->launches++;
    }
    PersistentArgs pa;
    pa.states = ws.d_state; pa.pls = ws.d_pair_level; pa.records = ws.d_records; pa.seg_export = ws.d_scale_export;
    pa.seg_base = ws.d_tile_base; pa.partial = ws.d_normal_partial;
    pa.squads = reinterpret_cast<SquadState*>(ws.d_squads);
    int* tail = reinterpret_cast<int*>(pa.squads + plan.nsquads);
    pa.next_pair = tail; pa.error_flag = tail + 1;
    pa.ilog = ws.d_iter_log; pa.max_log = max_log;
    pa.dbg = ctx->d_dbg ? ctx->d_dbg + 8 * li : nullptr;
    pa.npairs = n; pa.g = plan.g; pa.squads_per_slot = plan.squads_per_slot; pa.num_sms = ctx->num_sms; pa.rpw = plan.rpw;
    pa.nseg = plan.nseg; pa.lp = lp;
    {
      ProfScope prof(ctx, 0);
      void* args[] = {&pa};
      DVO_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_level_persistent, dim3(ctx->num_sms * ctx->ctas_per_sm),
                                                dim3(kSegmentsPerTile * 32), args, 0, st));
      ctx->launches++;
    }
    DVO_CUDA(ctx, cudaMemcpyAsync(&ws.h_active[level_flag_slot(li)], tail + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  // results
  dvo_b200_result* d_res = (dvo_b200_result*)d_results_user;
  if (!d_res) {
    size_t bytes = sizeof(dvo_b200_result) * n;
    if ((rc = ensure_stage(ctx, bytes, 0))) return rc;
    d_res = (dvo_b200_result*)ctx->d_stage;
  }
  {
    ProfScope prof(ctx, 2);
    k_finalize<<<(n + 63) / 64, 64, 0, st>>>(ws.d_state, d_res, n);
    ctx->launches++;
  }
  DVO_CUDA(ctx, cudaGetLastError());
  if (h_results) {
    size_t bytes = sizeof(dvo_b200_result) * n;
    if (bytes > ctx->h_results_bytes) {
      if (ctx->h_results) cudaFreeHost(ctx->h_results);
      ctx->h_results = nullptr; ctx->h_results_bytes = 0;
      DVO_CUDA(ctx, cudaMallocHost(&ctx->h_results, bytes));
      ctx->h_results_bytes = bytes;
    }
    DVO_CUDA(ctx, cudaMemcpyAsync(ctx->h_results, d_res, bytes, cudaMemcpyDeviceToHost, st));
    if (iter_stats) {
      DVO_CUDA(ctx, cudaMemcpyAsync(iter_stats, ws.d_iter_log, sizeof(dvo_b200_iteration_stats) * (size_t)n * max_log,
                                    cudaMemcpyDeviceToHost, st));
      ctx->d2h_bytes += sizeof(dvo_b200_iteration_stats) * (size_t)n * max_log;
    }
    DVO_CUDA(ctx, cudaStreamSynchronize(st));
    std::memcpy(h_results, ctx->h_results, bytes);
    ctx->d2h_bytes += bytes;
    for (int li = 0; li <= first - last && li < 8; ++li)
      if (ws.h_active[li] != 0) return set_error(ctx, DVO_B200_ERR_CUDA, "persistent level kernel: squad barrier timed out");
  }
  return 0;
}

int tracker_linearize(dvo_b200_ctx* ctx, const dvo_b200_config* cfg, dvo_b200_pyramid* ref, dvo_b200_pyramid* cur,
                      int level, const double* T, int use_weights, const float* prev_precision, int64_t* count,
                      float* precision_out, float* ll_out, double* A_out, double* b_out, float* planes7) {
  dvo_b200_config c = *cfg;
  c.first_level = level; c.last_level = level;
  dvo_b200_pyramid* refs[1] = {ref};
  dvo_b200_pyramid* curs[1] = {cur};
  int rc = check_batch(ctx, &c, 1, refs, curs);
  if (rc) return rc;
  if (!T) return set_error(ctx, DVO_B200_ERR_INVALID_ARGUMENT, "linearize: T is null");
  cudaStream_t st = ctx->stream;
  Workspace& ws = ctx->ws;
  const LevelInfo& L = ref->L[level];
  {
    const size_t nseg = (L.n + kSegmentPixels - 1) / kSegmentPixels;
    ScratchNeed need;
    need.record_floats = (size_t)kRecordFloatsPerPixel * L.n;
    need.export_floats = nseg * kCtaExportFloats;
    need.segbase_ints = nseg;
    need.partial_floats = nseg * kNormalValues;
    need.squads = 2;
    if ((rc = ensure_workspace(ctx, 1, need, 0))) return rc;
  }
  if ((rc = pyramid_reselect(ctx, ref, cfg->intensity_derivative_threshold, cfg->depth_derivative_threshold))) return rc;
  LevelLaunch lp;
  lp.w = L.w; lp.h = L.h; lp.n = L.n; lp.nseg = (L.n + kSegmentPixels - 1) / kSegmentPixels; lp.ntiles = (lp.nseg + kSegmentsPerTile - 1) / kSegmentsPerTile;
    lp.wmagic = (unsigned)((1ull << 32) / (unsigned)L.w) + 1u;
  lp.level_index = 0; lp.level_id = level; lp.max_iterations = 1 << 30; lp.first_level = 1;
  lp.use_initial_estimate = 0; lp.precision = 0.0; lp.mu = 0.0;
  if ((rc = upload_pair_levels(ctx, 1, refs, curs, level))) return rc;
  if ((rc = ensure_stage(ctx, 1024, 1024))) return rc;
  DVO_CUDA(ctx, cudaStreamSynchronize(st));
  std::memcpy(ctx->h_stage, T, sizeof(double) * 16);
  float pp[4] = {0, 0, 0, 0};
  if (use_weights && prev_precision) std::memcpy(pp, prev_precision, sizeof(pp));
  std::memcpy((char*)ctx->h_stage + 128, pp, sizeof(pp));
  DVO_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, ctx->h_stage, 256, cudaMemcpyHostToDevice, st));
  ws.h_active[0] = 1;
  DVO_CUDA(ctx, cudaMemcpyAsync(ws.d_active, ws.h_active, sizeof(int), cudaMemcpyHostToDevice, st));
  k_set_state<<<1, 1, 0, st>>>(ws.d_state, ws.d_pair_level, (const double*)ctx->d_stage,
                               (const float*)((char*)ctx->d_stage + 128), use_weights, lp);
  dim3 grid(lp.ntiles, 1);
  k_residual<<<grid, kSegmentsPerTile * 32, 0, st>>>(ws.d_state, ws.d_pair_level, ws.d_records, ws.d_scale_export, lp);
  k_pair_mid<<<1, 32, 0, st>>>(ws.d_state, ws.d_scale_export, ws.d_tile_base, lp.ntiles, ws.d_active, lp, nullptr, 0);
  ctx->launches += 3;
  if (!planes7) {
    k_normal<<<grid, kSegmentsPerTile * 32, 0, st>>>(ws.d_state, ws.d_pair_level, ws.d_records, ws.d_scale_export, ws.d_tile_base,
                                                     ws.d_normal_partial, lp);
    k_pair_end<<<1, kSegmentsPerTile * 32, 0, st>>>(ws.d_state, ws.d_pair_level, ws.d_normal_partial, lp.ntiles, ws.d_active, lp, nullptr, 0);
    ctx->launches += 2;
  }
  DVO_CUDA(ctx, cudaGetLastError());
  PairState* hs = nullptr;
  if ((rc = ensure_stage(ctx, 0, sizeof(PairState) + 64))) return rc;
  hs = (PairState*)ctx->h_stage;
  DVO_CUDA(ctx, cudaMemcpyAsync(hs, ws.d_state, sizeof(PairState), cudaMemcpyDeviceToHost, st));
  DVO_CUDA(ctx, cudaStreamSynchronize(st));
  if (count) *count = hs->n;
  if (precision_out) std::memcpy(precision_out, hs->precision, sizeof(float) * 4);
  if (ll_out) *ll_out = hs->ll;
  if (A_out) std::memcpy(A_out, hs->A, sizeof(double) * 36);
  if (b_out) std::memcpy(b_out, hs->b, sizeof(double) * 6);
  if (planes7) {
    // records: 6 planes + weight; return {ei, ez, gx, gy, hx, hy, z_ref}; invalid -> NaN in every plane
    size_t N = L.n;
    std::vector<float> rec(7 * N), p0(2 * N);
    DVO_CUDA(ctx, cudaMemcpy(rec.data(), ws.d_records, sizeof(float) * 7 * N, cudaMemcpyDeviceToHost));
    DVO_CUDA(ctx, cudaMemcpy(p0.data(), ref->planes + L.plane_off, sizeof(float) * 2 * N, cudaMemcpyDeviceToHost));
    ctx->d2h_bytes += sizeof(float) * 9 * N;
    const float nanv = std::numeric_limits<float>::quiet_NaN();
    // scratch layout: float2 planes E = (e.i, e.z), G = (e.idx, e.idy), H = (e.zdx, e.zdy), then W
    for (size_t i = 0; i < N; ++i) {
      bool valid = rec[2 * i] == rec[2 * i];
      for (int pl = 0; pl < 3; ++pl) {
        planes7[(2 * pl) * N + i] = valid ? rec[pl * 2 * N + 2 * i] : nanv;
        planes7[(2 * pl + 1) * N + i] = valid ? rec[pl * 2 * N + 2 * i + 1] : nanv;
      }
      planes7[6 * N + i] = valid ? p0[2 * i + 1] : nanv;
    }
  }
  return 0;
}

}  // namespace dvo_b200
