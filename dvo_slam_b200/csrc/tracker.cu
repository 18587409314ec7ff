// tracker.cu -- dvo::DenseTracker::match() (dvo_core/src/dense_tracking.cpp:131-376) for a batch of
// independent frame pairs, all state on the device.
//
// Per Gauss-Newton iteration the reference makes five passes over the points
// (computeResidualsSse, computeWeightsSse, computeScaleSse, computeCompleteDataLogLikelihood and the
// normal-equation loop, dense_tracking.cpp:271-343).  Precision P_k is a global reduction that the
// log-likelihood and J^T W J depend on, so there are exactly two data-parallel stages:
//   stage A (k_residual): warp/interpolate/residual/occlusion test, Student-t weight from P_{k-1},
//                         pairwise scale sums, residual record kept for stage B
//   stage B (k_normal):   log-likelihood terms and the 21+6 normal-equation coefficients with W = w*P_k
// with one tiny per-pair kernel after each (k_pair_mid: P_k; k_pair_end: accept test, 6x6 LDL^T solve,
// SE(3) update, termination logic).
#include "common.cuh"

#include <cstdio>
#include <cstring>
#include <limits>
#include <cmath>

namespace dvo_b200 {

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct LevelLaunch {
  int w, h, n, ntiles;
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

__global__ void k_level_begin(PairState* states, const PairLevel* pls, const double* T_init, int npairs,
                              LevelLaunch lp) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  PairState& st = states[p];
  const PairLevel& pl = pls[p];
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
// stage A
// ------------------------------------------------------------------------------------------------
struct TileConsts {
  float kt[12];
  float P[4];
  float c_i, c_gx, c_gy, fx, fy, ubx, uby;
  int first_iteration;
  int S, last_sel;
};

struct PixelOut {
  float ei, ez, gx, gy, hx, hy;
};

// The residual record of one reference pixel: computeResidualsSse (dense_tracking_impl.cpp:133-393).
// Every fp32 operation is an explicit round-to-nearest intrinsic so that the value is defined bit
// for bit (the oracle's MIRROR mode restates exactly this sequence on the CPU):
//   point (x,y,z) = (tx*z, ty*z, z); (X,Y,Z') = fma chains of KT rows; (u,v) = (X,Y) * rcp_rn(Z')
//   bounds 0<=u<=w-2, 0<=v<=h-2; truncation; bilinear blend of the six channels
//   residual record with the weights of dense_tracking.cpp:215-220; occlusion test (line 275)
__device__ __forceinline__ bool pixel_record(int idx, int x, int y, int w, const PairLevel& pl, const TileConsts& tc,
                                             PixelOut& o) {
  float2 a = __ldg(pl.r0 + idx);   // (I_r, Z_r)
  float2 g = __ldg(pl.r1 + idx);   // (Ix_r, Iy_r)
  float z = a.y;
  float tx = __ldg(pl.rtmpl + x), ty = __ldg(pl.rtmpl + w + y);
  float px = __fmul_rn(tx, z), py = __fmul_rn(ty, z);
  float X = __fmaf_rn(tc.kt[0], px, __fmaf_rn(tc.kt[1], py, __fmaf_rn(tc.kt[2], z, tc.kt[3])));
  float Y = __fmaf_rn(tc.kt[4], px, __fmaf_rn(tc.kt[5], py, __fmaf_rn(tc.kt[6], z, tc.kt[7])));
  float Zt = __fmaf_rn(tc.kt[8], px, __fmaf_rn(tc.kt[9], py, __fmaf_rn(tc.kt[10], z, tc.kt[11])));
  float rz = __frcp_rn(Zt);
  float u = __fmul_rn(X, rz), v = __fmul_rn(Y, rz);
  if (!(u >= 0.f && u <= tc.ubx && v >= 0.f && v <= tc.uby)) return false;
  int u0 = __float2int_rz(u), v0 = __float2int_rz(v);
  float fu = __fsub_rn(u, (float)u0), fv = __fsub_rn(v, (float)v0);
  float gu = __fsub_rn(1.0f, fu), gv = __fsub_rn(1.0f, fv);
  int b = v0 * w + u0;
  float2 p00 = __ldg(pl.c0 + b), p10 = __ldg(pl.c0 + b + 1), p01 = __ldg(pl.c0 + b + w), p11 = __ldg(pl.c0 + b + w + 1);
#define DVO_BLEND(c00, c10, c01, c11) \
  __fmaf_rn(fv, __fmaf_rn(fu, c11, __fmul_rn(gu, c01)), __fmul_rn(gv, __fmaf_rn(fu, c10, __fmul_rn(gu, c00))))
  float Zc = DVO_BLEND(p00.y, p10.y, p01.y, p11.y);
  if (Zc != Zc) return false;     // masked depth: any NaN lane of the reference's 8-vector
  float Ic = DVO_BLEND(p00.x, p10.x, p01.x, p11.x);
  o.ez = __fsub_rn(Zc, Zt);
  float s = __fsub_rn(z, 0.4f);
  float sig = __fmaf_rn(__fmul_rn(0.0019f, s), s, 0.0012f);   // depthStdDevZ (dense_tracking_impl.cpp:122-128)
  if (!(o.ez > __fmul_rn(-20.0f, sig))) return false;         // occlusion test
  o.ei = __fmaf_rn(tc.c_i, Ic, __fmul_rn(-tc.c_i, a.x));
  float2 q00 = __ldg(pl.c1 + b), q10 = __ldg(pl.c1 + b + 1), q01 = __ldg(pl.c1 + b + w), q11 = __ldg(pl.c1 + b + w + 1);
  float Ixc = DVO_BLEND(q00.x, q10.x, q01.x, q11.x);
  float Iyc = DVO_BLEND(q00.y, q10.y, q01.y, q11.y);
  o.gx = __fmaf_rn(tc.c_gx, Ixc, __fmul_rn(tc.c_gx, g.x));
  o.gy = __fmaf_rn(tc.c_gy, Iyc, __fmul_rn(tc.c_gy, g.y));
  float2 s00 = __ldg(pl.c2 + b), s10 = __ldg(pl.c2 + b + 1), s01 = __ldg(pl.c2 + b + w), s11 = __ldg(pl.c2 + b + w + 1);
  float Zxc = DVO_BLEND(s00.x, s10.x, s01.x, s11.x);
  float Zyc = DVO_BLEND(s00.y, s10.y, s01.y, s11.y);
#undef DVO_BLEND
  o.hx = __fmul_rn(tc.fx, Zxc);
  o.hy = __fmul_rn(tc.fy, Zyc);
  return true;
}

__device__ __forceinline__ void load_tile_consts(const PairState& st, const PairLevel& pl, int w, int h, TileConsts& tc) {
#pragma unroll
  for (int i = 0; i < 12; ++i) tc.kt[i] = st.kt[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) tc.P[i] = st.precision[i];
  tc.c_i = 1.0f / 255.0f;
  tc.c_gx = __fdiv_rn(__fmul_rn(0.5f, pl.cfx), 255.0f);
  tc.c_gy = __fdiv_rn(__fmul_rn(0.5f, pl.cfy), 255.0f);
  tc.fx = pl.cfx; tc.fy = pl.cfy;
  tc.ubx = (float)(w - 2); tc.uby = (float)(h - 2);
  tc.first_iteration = st.iteration == 0;
  tc.S = pl.rsel[0]; tc.last_sel = pl.rsel[1];
}

// Pairwise scale sum (computeScaleSse, dense_tracking_impl.cpp:590-638).  The reference walks the
// compacted residual list two at a time and, because lines 614-615 re-use the low half of the
// register, adds (w_{2j} + w_{2j+1}) * r_{2j} r_{2j}^T for every pair and w_n r_n r_n^T for an odd
// tail.  Reproducing that needs, for every valid point, the parity of its rank in row-major order
// and the weight of the next valid point.  A contiguous run of pixels is summarised by a ScaleSeg:
// sums under both hypotheses for the parity of its first point (S0: first point is a pair leader,
// S1: it is a follower), its first valid weight and its last valid point (a leader whose partner
// lies in the next run).  Runs combine associatively (combine_seg), so warps, tiles and finally the
// whole image are reduced in a fixed order.
template <typename T>
struct SegT {
  long long n;
  T S0[3], S1[3];
  T wf, wl, ol[3];
};

template <typename T, typename A, typename B>
__host__ __device__ __forceinline__ SegT<T> combine_seg(const A& a, const B& b) {
  SegT<T> r;
  r.n = (long long)a.n + (long long)b.n;
  int hb0 = (int)(a.n & 1), hb1 = (int)((a.n + 1) & 1);
  bool link0 = a.n > 0 && b.n > 0 && (((a.n - 1) & 1) == 0);       // h = 0: last point of a is a leader
  bool link1 = a.n > 0 && b.n > 0 && (((a.n - 1 + 1) & 1) == 0);   // h = 1
  for (int k = 0; k < 3; ++k) {
    T bs0 = hb0 ? (T)b.S1[k] : (T)b.S0[k];
    T bs1 = hb1 ? (T)b.S1[k] : (T)b.S0[k];
    r.S0[k] = (T)a.S0[k] + bs0 + (link0 ? ((T)a.wl + (T)b.wf) * (T)a.ol[k] : (T)0);
    r.S1[k] = (T)a.S1[k] + bs1 + (link1 ? ((T)a.wl + (T)b.wf) * (T)a.ol[k] : (T)0);
  }
  r.wf = a.n > 0 ? (T)a.wf : (T)b.wf;
  if (b.n > 0) { r.wl = (T)b.wl; for (int k = 0; k < 3; ++k) r.ol[k] = (T)b.ol[k]; }
  else         { r.wl = (T)a.wl; for (int k = 0; k < 3; ++k) r.ol[k] = (T)a.ol[k]; }
  return r;
}

__global__ void __launch_bounds__(kTileThreads)
k_residual(const PairState* __restrict__ states, const PairLevel* __restrict__ pls, float* __restrict__ records,
           float* __restrict__ scale_export, int w, int h, int n, int ntiles) {
  const int pair = blockIdx.y, tile = blockIdx.x;
  const PairState& st = states[pair];
  if (!st.level_active) return;
  const PairLevel pl = pls[pair];
  TileConsts tc;
  load_tile_consts(st, pl, w, h, tc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* rec = records + (size_t)pair * 7 * n;
  const bool drop_last = (tc.S & 1) != 0;  // odd number of selected points: last one skipped (dense_tracking_impl.cpp:169)

  float S0[3] = {0.f, 0.f, 0.f}, S1[3] = {0.f, 0.f, 0.f};
  bool pend = false;
  float pw = 0.f, po0 = 0.f, po1 = 0.f, po2 = 0.f;
  int ppar = 0, cnt = 0;
  float wfirst = 0.f;

  const int seg_base = tile * kTilePixels + warp * 128;
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
    const int base = seg_base + r * 32;
    if (base >= n) break;
    const int idx = base + lane;
    bool valid = false;
    PixelOut o;
    float wgt = 1.0f;
    unsigned selw = __ldg(pl.rmask + (base >> 5));
    if ((selw >> lane) & 1u) {
      if (!(drop_last && idx == tc.last_sel)) {
        int y = idx / w, x = idx - y * w;
        valid = pixel_record(idx, x, y, w, pl, tc, o);
      }
    }
    if (valid && !tc.first_iteration) {
      // computeWeightsSse (dense_tracking_impl.cpp:657-707): w = 7 / (5 + r^T P r), nu = 5
      float d = (o.ei * tc.P[0] + o.ez * tc.P[2]) * o.ei + (o.ei * tc.P[1] + o.ez * tc.P[3]) * o.ez;
      wgt = __fdividef(7.0f, 5.0f + d);
    }
    if (idx < n) {
      const float nanf_ = __int_as_float(0x7fc00000);
      rec[idx] = valid ? o.ei : nanf_;
      if (valid) {
        rec[(size_t)n + idx] = o.ez; rec[2 * (size_t)n + idx] = o.gx; rec[3 * (size_t)n + idx] = o.gy;
        rec[4 * (size_t)n + idx] = o.hx; rec[5 * (size_t)n + idx] = o.hy; rec[6 * (size_t)n + idx] = wgt;
      }
    }
    unsigned m = __ballot_sync(kFull, valid);
    if (m) {
      int first = __ffs(m) - 1, last = 31 - __clz(m);
      float w_first = __shfl_sync(kFull, wgt, first);
      if (cnt == 0) wfirst = w_first;
      if (pend && lane == 0) {
        float s = pw + w_first;
        if (ppar) { S1[0] += s * po0; S1[1] += s * po1; S1[2] += s * po2; }
        else      { S0[0] += s * po0; S0[1] += s * po1; S0[2] += s * po2; }
      }
      unsigned above = lane == 31 ? 0u : (m >> (lane + 1));
      int nxt = above ? lane + __ffs(above) : lane;
      float w_next = __shfl_sync(kFull, wgt, nxt);
      float o0 = valid ? o.ei * o.ei : 0.f, o1 = valid ? o.ei * o.ez : 0.f, o2 = valid ? o.ez * o.ez : 0.f;
      if (valid && above) {
        int rank = cnt + __popc(m & ((1u << lane) - 1u));
        float s = wgt + w_next;
        if (rank & 1) { S1[0] += s * o0; S1[1] += s * o1; S1[2] += s * o2; }
        else          { S0[0] += s * o0; S0[1] += s * o1; S0[2] += s * o2; }
      }
      pw = __shfl_sync(kFull, wgt, last);
      po0 = __shfl_sync(kFull, o0, last); po1 = __shfl_sync(kFull, o1, last); po2 = __shfl_sync(kFull, o2, last);
      cnt += __popc(m);
      ppar = (cnt - 1) & 1;
      pend = true;
    }
  }
  // warp reduction of the six sums, then the 8 warp segments of the tile are combined in order
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      S0[k] += __shfl_xor_sync(kFull, S0[k], off);
      S1[k] += __shfl_xor_sync(kFull, S1[k], off);
    }
  }
  __shared__ SegT<float> segs[kTileThreads / 32];
  if (lane == 0) {
    SegT<float>& s = segs[warp];
    s.n = cnt;
    for (int k = 0; k < 3; ++k) { s.S0[k] = S0[k]; s.S1[k] = S1[k]; }
    s.wf = wfirst; s.wl = pw; s.ol[0] = po0; s.ol[1] = po1; s.ol[2] = po2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    SegT<float> acc = segs[0];
    for (int k = 1; k < kTileThreads / 32; ++k) acc = combine_seg<float>(acc, segs[k]);
    float* e = scale_export + ((size_t)pair * ntiles + tile) * kScaleExportFloats;
    e[0] = __int_as_float((int)acc.n);
    e[1] = acc.S0[0]; e[2] = acc.S0[1]; e[3] = acc.S0[2];
    e[4] = acc.S1[0]; e[5] = acc.S1[1]; e[6] = acc.S1[2];
    e[7] = acc.wf; e[8] = acc.wl; e[9] = acc.ol[0]; e[10] = acc.ol[1]; e[11] = acc.ol[2];
  }
}

__device__ __forceinline__ SegT<double> load_seg(const float* e) {
  SegT<double> s;
  s.n = __float_as_int(e[0]);
  s.S0[0] = e[1]; s.S0[1] = e[2]; s.S0[2] = e[3];
  s.S1[0] = e[4]; s.S1[1] = e[5]; s.S1[2] = e[6];
  s.wf = e[7]; s.wl = e[8]; s.ol[0] = e[9]; s.ol[1] = e[10]; s.ol[2] = e[11];
  return s;
}

// one warp per pair: combine tile summaries in order -> covariance -> P_k (dense_tracking.cpp:276-295)
__global__ void k_pair_mid(PairState* states, const float* __restrict__ scale_export, int* __restrict__ tile_base,
                           int ntiles, int* active, LevelLaunch lp, dvo_b200_iteration_stats* ilog, int max_log) {
  const int pair = blockIdx.x, lane = threadIdx.x;
  PairState& st = states[pair];
  if (!st.level_active) return;
  const float* e = scale_export + (size_t)pair * ntiles * kScaleExportFloats;
  int chunk = (ntiles + 31) / 32;
  int t0 = lane * chunk, t1 = min(t0 + chunk, ntiles);
  SegT<double> acc;
  acc.n = 0; acc.wf = acc.wl = 0;
  for (int k = 0; k < 3; ++k) acc.S0[k] = acc.S1[k] = acc.ol[k] = 0;
  for (int t = t0; t < t1; ++t) acc = combine_seg<double>(acc, load_seg(e + (size_t)t * kScaleExportFloats));
  __shared__ SegT<double> lanes[32];
  __shared__ long long lane_base[32];
  lanes[lane] = acc;
  __syncwarp();
  if (lane == 0) {
    SegT<double> all = lanes[0];
    long long run = 0;
    lane_base[0] = 0;
    run = lanes[0].n;
    for (int k = 1; k < 32; ++k) { lane_base[k] = run; run += lanes[k].n; all = combine_seg<double>(all, lanes[k]); }
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
      atomicSub(active, 1);
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
    tile_base[(size_t)pair * ntiles + t] = (int)run;
    run += __float_as_int(e[(size_t)t * kScaleExportFloats]);
  }
}

// ------------------------------------------------------------------------------------------------
// stage B
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
k_normal(const PairState* __restrict__ states, const PairLevel* __restrict__ pls, const float* __restrict__ records,
         const float* __restrict__ scale_export, const int* __restrict__ tile_base, float* __restrict__ partial, int w,
         int h, int n, int ntiles) {
  const int pair = blockIdx.y, tile = blockIdx.x;
  const PairState& st = states[pair];
  if (!st.level_active || !st.phase_ok) return;
  const PairLevel pl = pls[pair];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* rec = records + (size_t)pair * 7 * n;
  const float P0 = st.precision[0], P1 = st.precision[1], P2 = st.precision[2], P3 = st.precision[3];
  const long long n_keep = st.n_keep;
  const int tbase = tile_base[(size_t)pair * ntiles + tile];
  const int tcount = __float_as_int(scale_export[((size_t)pair * ntiles + tile) * kScaleExportFloats]);
  const bool need_rank = (long long)tbase + tcount > n_keep;   // only the last tile(s) of the image

  __shared__ int warp_cnt[kTileThreads / 32];
  const int seg_base = tile * kTilePixels + warp * 128;
  int warp_prefix = 0;
  if (need_rank) {   // block-uniform
    int c = 0;
    for (int r = 0; r < 4; ++r) {
      int idx = seg_base + r * 32 + lane;
      float v = idx < n ? rec[idx] : __int_as_float(0x7fc00000);
      c += __popc(__ballot_sync(kFull, v == v));
    }
    if (lane == 0) warp_cnt[warp] = c;
    __syncthreads();
    for (int k = 0; k < warp; ++k) warp_prefix += warp_cnt[k];
  }

  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  float prod = 1.0f;
  float llsum = 0.f;
  int seen = 0;
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
    const int base = seg_base + r * 32;
    if (base >= n) break;
    const int idx = base + lane;
    float ei = idx < n ? rec[idx] : __int_as_float(0x7fc00000);
    bool valid = ei == ei;
    bool keep = valid;
    if (need_rank) {
      unsigned m = __ballot_sync(kFull, valid);
      long long rank = (long long)tbase + warp_prefix + seen + __popc(m & ((1u << lane) - 1u));
      keep = valid && rank < n_keep;
      seen += __popc(m);
    }
    if (!valid) continue;
    float ez = rec[(size_t)n + idx], gx = rec[2 * (size_t)n + idx], gy = rec[3 * (size_t)n + idx];
    float hx = rec[4 * (size_t)n + idx], hy = rec[5 * (size_t)n + idx], wgt = rec[6 * (size_t)n + idx];
    // log-likelihood term (dense_tracking_impl.cpp:406-425): log(1 + 0.2 r^T P r)
    float d = (ei * P0 + ez * P2) * ei + (ei * P1 + ez * P3) * ez;
    if (keep) prod *= fmaf(0.2f, d, 1.0f);
    // Jacobians at the untransformed reference point (dense_tracking.cpp:448-476, 338-339)
    int y = idx / w, x = idx - y * w;
    float z = __ldg(pl.r0 + idx).y;
    float tx = __ldg(pl.rtmpl + x), ty = __ldg(pl.rtmpl + w + y);
    float px = tx * z, py = ty * z;
    float zi = 1.0f / z, zs = zi * zi;
    float a2 = -px * zs, a3 = a2 * py, a4 = 1.0f - a2 * px, a5 = -py * zi;
    float b2 = -py * zs, b3 = -1.0f + b2 * py, b4 = -a3, b5 = px * zi;
    float J0[6] = {gx * zi, gy * zi, gx * a2 + gy * b2, gx * a3 + gy * b3, gx * a4 + gy * b4, gx * a5 + gy * b5};
    float J1[6] = {hx * zi, hy * zi, hx * a2 + hy * b2 - 1.0f, hx * a3 + hy * b3 - py, hx * a4 + hy * b4 + px,
                   hx * a5 + hy * b5};
    // W = w * P_k ; A += J^T W J ; b -= J^T W r (least_squares.cpp:58-64)
    float W00 = wgt * P0, W01 = wgt * P1, W10 = wgt * P2, W11 = wgt * P3;
    float ua[6], ub[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) { ua[c] = J0[c] * W00 + J1[c] * W10; ub[c] = J0[c] * W01 + J1[c] * W11; }
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = i; j < 6; ++j) { acc[k] += ua[i] * J0[j] + ub[i] * J1[j]; ++k; }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[21 + i] -= ua[i] * ei + ub[i] * ez;
  }
  llsum = logf(prod);
  // block reduction: warp shuffles then shared memory, fixed order
  __shared__ float red[kTileThreads / 32][kNormalPartialFloats];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) llsum += __shfl_xor_sync(kFull, llsum, off);
#pragma unroll
  for (int i = 0; i < 27; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[i] += __shfl_xor_sync(kFull, acc[i], off);
  }
  if (lane == 0) {
    red[warp][0] = llsum;
#pragma unroll
    for (int i = 0; i < 27; ++i) red[warp][1 + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < kNormalPartialFloats) {
    float s = 0.f;
    for (int k = 0; k < kTileThreads / 32; ++k) s += red[k][threadIdx.x];
    partial[((size_t)pair * ntiles + tile) * kNormalPartialFloats + threadIdx.x] = s;
  }
}

// one warp per pair: reduce tile partials, log-likelihood, accept test, solve, termination
// (dense_tracking.cpp:297-363)
__global__ void k_pair_end(PairState* states, const PairLevel* pls, const float* __restrict__ partial, int ntiles,
                           int* active, LevelLaunch lp, dvo_b200_iteration_stats* ilog, int max_log) {
  const int pair = blockIdx.x, lane = threadIdx.x;
  PairState& st = states[pair];
  if (!st.level_active || !st.phase_ok) return;
  double v = 0.0;
  if (lane < kNormalPartialFloats) {
    const float* p = partial + (size_t)pair * ntiles * kNormalPartialFloats + lane;
    for (int t = 0; t < ntiles; ++t) v += (double)p[(size_t)t * kNormalPartialFloats];
  }
  double vals[kNormalPartialFloats];
#pragma unroll
  for (int i = 0; i < kNormalPartialFloats; ++i) vals[i] = __shfl_sync(kFull, v, i);
  if (lane != 0) return;

  const PairLevel& pl = pls[pair];
  LevelSummary& ls = st.levels[lp.level_index];
  // computeCompleteDataLogLikelihood: 0.5 n log det P - 3.5 sum log(1 + 0.2 d), returned as float
  float det = __fsub_rn(__fmul_rn(st.precision[0], st.precision[3]), __fmul_rn(st.precision[1], st.precision[2]));
  float logdet = (float)log((double)det);
  float ll = (float)(0.5 * (double)st.n * (double)logdet - 0.5 * (5.0 + 2.0) * vals[0]);
  st.ll = ll;
  st.nll_cur = -(double)ll;
  double li[6];
  se3_log(st.initial, li);
  double sq = 0;
  for (int i = 0; i < 6; ++i) sq += li[i] * li[i];
  st.prior_cur = lp.mu * sq;                       // dense_tracking.cpp:302
  st.last_error = st.error;                        // dense_tracking.cpp:306-307
  st.error = -(double)ll;
  bool accept = st.error < st.last_error;          // dense_tracking.cpp:312

  // unpack A (upper triangle) and b
  {
    int k = 1;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) { st.A[i * 6 + j] = vals[k]; st.A[j * 6 + i] = vals[k]; ++k; }
    for (int i = 0; i < 6; ++i) st.b[i] = vals[22 + i];
  }
  bool level_done = false;
  if (!accept) {
    st.initial = st.initial_old; st.estimate = st.estimate_old;   // dense_tracking.cpp:314-321
    st.termination = DVO_B200_TERM_LOG_LIKELIHOOD_DECREASED;
    log_iteration(ilog, max_log, pair, st, lp.level_id, false);
    level_done = true;
  } else {
    double A[36], b[6];
    for (int i = 0; i < 36; ++i) A[i] = st.A[i];
    for (int i = 0; i < 6; ++i) { A[i * 6 + i] += lp.mu; b[i] = st.b[i] + lp.mu * li[i]; }   // lines 345-346
    ldlt_solve6(A, b, st.x);                                                                  // line 347
    for (int i = 0; i < 36; ++i) st.A_done[i] = A[i];
    st.nll_done = st.nll_cur; st.prior_done = st.prior_cur; st.have_done = 1;
    ls.last_inc_n = st.n; ls.last_inc_nll = st.nll_cur;
    log_iteration(ilog, max_log, pair, st, lp.level_id, true);
    st.iteration += 1;                                                                        // line 353
  }
  double m = 0; bool nanx = false;
  for (int i = 0; i < 6; ++i) { m = fmax(m, fabs(st.x[i])); nanx |= st.x[i] != st.x[i]; }
  bool big = !nanx && m > lp.precision;
  bool exceeded = st.iteration >= lp.max_iterations;
  if (!(accept && big && !exceeded)) level_done = true;                                       // line 357
  if (level_done) {
    if (!nanx && m <= lp.precision) st.termination = DVO_B200_TERM_INCREMENT_TOO_SMALL;       // line 359
    if (exceeded) st.termination = DVO_B200_TERM_ITERATIONS_EXCEEDED;                         // line 362
    ls.termination = st.termination;
    int need = (st.termination == DVO_B200_TERM_LOG_LIKELIHOOD_DECREASED ||
                st.termination == DVO_B200_TERM_TOO_FEW_CONSTRAINTS) ? 2 : 1;
    ls.has_inc = ls.num_iterations >= need;
    st.level_active = 0;
    atomicSub(active, 1);
  } else {
    prepare_iteration(st, pl);
  }
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
template <typename T>
int grow(dvo_b200_ctx* ctx, T*& ptr, size_t& cap, size_t need) {
  if (need <= cap) return 0;
  if (ptr) { cudaStreamSynchronize(ctx->stream); cudaFree(ptr); ptr = nullptr; cap = 0; }
  DVO_CUDA(ctx, cudaMalloc((void**)&ptr, need * sizeof(T)));
  cap = need;
  return 0;
}

int ensure_workspace(dvo_b200_ctx* ctx, int npairs, int n0, int max_log_per_pair) {
  Workspace& ws = ctx->ws;
  int ntiles0 = (n0 + kTilePixels - 1) / kTilePixels;
  if ((size_t)npairs > ws.cap_pairs) {
    if (ws.d_pair_level) { cudaStreamSynchronize(ctx->stream); cudaFree(ws.d_pair_level); cudaFree(ws.d_state); }
    ws.d_pair_level = nullptr; ws.d_state = nullptr; ws.cap_pairs = 0;
    DVO_CUDA(ctx, cudaMalloc((void**)&ws.d_pair_level, sizeof(PairLevel) * npairs));
    DVO_CUDA(ctx, cudaMalloc((void**)&ws.d_state, sizeof(PairState) * npairs));
    ws.cap_pairs = npairs;
  }
  int rc;
  if ((rc = grow(ctx, ws.d_records, ws.cap_records, (size_t)npairs * 7 * n0))) return rc;
  size_t tiles = (size_t)npairs * ntiles0;
  if (tiles > ws.cap_tiles) {
    if (ws.d_scale_export) { cudaStreamSynchronize(ctx->stream); cudaFree(ws.d_scale_export); cudaFree(ws.d_tile_base); cudaFree(ws.d_normal_partial); }
    ws.d_scale_export = nullptr; ws.d_tile_base = nullptr; ws.d_normal_partial = nullptr; ws.cap_tiles = 0;
    DVO_CUDA(ctx, cudaMalloc((void**)&ws.d_scale_export, tiles * kScaleExportFloats * sizeof(float)));
    DVO_CUDA(ctx, cudaMalloc((void**)&ws.d_tile_base, tiles * sizeof(int)));
    DVO_CUDA(ctx, cudaMalloc((void**)&ws.d_normal_partial, tiles * kNormalPartialFloats * sizeof(float)));
    ws.cap_tiles = tiles;
  }
  if (!ws.d_active) {
    DVO_CUDA(ctx, cudaMalloc((void**)&ws.d_active, sizeof(int) * 4));
    DVO_CUDA(ctx, cudaMallocHost((void**)&ws.h_active, sizeof(int) * 4));
  }
  if (max_log_per_pair > 0) {
    size_t need = (size_t)npairs * max_log_per_pair;
    if ((rc = grow(ctx, ws.d_iter_log, ws.cap_iter_log, need))) return rc;
  }
  return 0;
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
      if (p->ctx != ctx && p->slab && p->slab->ready) cudaStreamWaitEvent(ctx->stream, p->slab->ready, 0);
    if (refs[i]->levels <= cfg->first_level || curs[i]->levels <= cfg->first_level)
      return set_error(ctx, DVO_B200_ERR_INVALID_ARGUMENT, "match: pyramid has fewer levels than FirstLevel+1");
    if (refs[i]->L[0].w != refs[0]->L[0].w || refs[i]->L[0].h != refs[0]->L[0].h ||
        curs[i]->L[0].w != refs[0]->L[0].w || curs[i]->L[0].h != refs[0]->L[0].h)
      return set_error(ctx, DVO_B200_ERR_SHAPE_MISMATCH, "match: all pyramids of a batch must share width/height");
  }
  return 0;
}

int upload_pair_levels(dvo_b200_ctx* ctx, int n, dvo_b200_pyramid* const* refs, dvo_b200_pyramid* const* curs, int level) {
  size_t bytes = sizeof(PairLevel) * n;
  int rc = ensure_stage(ctx, 0, bytes);
  if (rc) return rc;
  DVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // previous use of the pinned stage has drained
  PairLevel* h = (PairLevel*)ctx->h_stage;
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
  DVO_CUDA(ctx, cudaMemcpyAsync(ctx->ws.d_pair_level, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
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
  rc = ensure_workspace(ctx, n, refs[0]->L[last].n, max_log);
  if (rc) return rc;

  // selection masks for non-default thresholds (PointSelection caches per pyramid, point_selection.cpp:100-113)
  for (int i = 0; i < n; ++i)
    if ((rc = pyramid_reselect(ctx, refs[i], cfg->intensity_derivative_threshold, cfg->depth_derivative_threshold))) return rc;

  // T_init upload
  double* d_Tinit = nullptr;
  if (cfg->use_initial_estimate && T_init) {
    size_t bytes = sizeof(double) * 16 * n;
    if ((rc = ensure_stage(ctx, bytes + 256, 0))) return rc;
    d_Tinit = (double*)ctx->d_stage;
    DVO_CUDA(ctx, cudaMemcpyAsync(d_Tinit, T_init, bytes, cudaMemcpyHostToDevice, st));
    DVO_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->h2d_bytes += bytes;
  }

  for (int level = first, li = 0; level >= last; --level, ++li) {
    const LevelInfo& L = refs[0]->L[level];
    LevelLaunch lp;
    lp.w = L.w; lp.h = L.h; lp.n = L.n; lp.ntiles = (L.n + kTilePixels - 1) / kTilePixels;
    lp.level_index = li; lp.level_id = level; lp.max_iterations = cfg->max_iterations_per_level;
    lp.first_level = li == 0; lp.use_initial_estimate = cfg->use_initial_estimate;
    lp.precision = cfg->precision; lp.mu = cfg->mu;
    if ((rc = upload_pair_levels(ctx, n, refs, curs, level))) return rc;
    ws.h_active[0] = n;
    DVO_CUDA(ctx, cudaMemcpyAsync(ws.d_active, ws.h_active, sizeof(int), cudaMemcpyHostToDevice, st));
    {
      ProfScope prof(ctx, 2);
      k_level_begin<<<(n + 63) / 64, 64, 0, st>>>(ws.d_state, ws.d_pair_level, d_Tinit, n, lp);
      ctx->launches++;
    }
    dim3 grid(lp.ntiles, n);
    for (int it = 0; it < lp.max_iterations || it == 0; ++it) {
      {
        ProfScope prof(ctx, 0);
        k_residual<<<grid, kTileThreads, 0, st>>>(ws.d_state, ws.d_pair_level, ws.d_records, ws.d_scale_export, lp.w, lp.h,
                                                  lp.n, lp.ntiles);
      }
      {
        ProfScope prof(ctx, 2);
        k_pair_mid<<<n, 32, 0, st>>>(ws.d_state, ws.d_scale_export, ws.d_tile_base, lp.ntiles, ws.d_active, lp,
                                     ws.d_iter_log, max_log);
      }
      {
        ProfScope prof(ctx, 1);
        k_normal<<<grid, kTileThreads, 0, st>>>(ws.d_state, ws.d_pair_level, ws.d_records, ws.d_scale_export, ws.d_tile_base,
                                                ws.d_normal_partial, lp.w, lp.h, lp.n, lp.ntiles);
      }
      {
        ProfScope prof(ctx, 2);
        k_pair_end<<<n, 32, 0, st>>>(ws.d_state, ws.d_pair_level, ws.d_normal_partial, lp.ntiles, ws.d_active, lp,
                                     ws.d_iter_log, max_log);
      }
      ctx->launches += 4;
      DVO_CUDA(ctx, cudaMemcpyAsync(ws.h_active, ws.d_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      DVO_CUDA(ctx, cudaStreamSynchronize(st));
      if (ws.h_active[0] <= 0) break;
    }
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
  if ((rc = ensure_workspace(ctx, 1, L.n, 0))) return rc;
  if ((rc = pyramid_reselect(ctx, ref, cfg->intensity_derivative_threshold, cfg->depth_derivative_threshold))) return rc;
  LevelLaunch lp;
  lp.w = L.w; lp.h = L.h; lp.n = L.n; lp.ntiles = (L.n + kTilePixels - 1) / kTilePixels;
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
  k_residual<<<grid, kTileThreads, 0, st>>>(ws.d_state, ws.d_pair_level, ws.d_records, ws.d_scale_export, lp.w, lp.h, lp.n, lp.ntiles);
  k_pair_mid<<<1, 32, 0, st>>>(ws.d_state, ws.d_scale_export, ws.d_tile_base, lp.ntiles, ws.d_active, lp, nullptr, 0);
  ctx->launches += 3;
  if (!planes7) {
    k_normal<<<grid, kTileThreads, 0, st>>>(ws.d_state, ws.d_pair_level, ws.d_records, ws.d_scale_export, ws.d_tile_base,
                                            ws.d_normal_partial, lp.w, lp.h, lp.n, lp.ntiles);
    k_pair_end<<<1, 32, 0, st>>>(ws.d_state, ws.d_pair_level, ws.d_normal_partial, lp.ntiles, ws.d_active, lp, nullptr, 0);
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
    for (size_t i = 0; i < N; ++i) {
      bool valid = rec[i] == rec[i];
      for (int k = 0; k < 6; ++k) planes7[k * N + i] = valid ? rec[k * N + i] : nanv;
      planes7[6 * N + i] = valid ? p0[2 * i + 1] : nanv;
    }
  }
  return 0;
}

}  // namespace dvo_b200
