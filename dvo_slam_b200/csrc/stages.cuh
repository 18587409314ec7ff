// stages.cuh -- the two data-parallel stages of one Gauss-Newton iteration of
// dvo::DenseTracker::match() as warp-level device functions (used by every kernel in tracker.cu).
//
//   stage A (stage_a_segment): computeResidualsSse + computeWeightsSse + computeScaleSse
//                              (dense_tracking_impl.cpp:133-393, 657-707, 590-638)
//   stage B (stage_b_segment): computeCompleteDataLogLikelihood + Jacobians + normal equations
//                              (dense_tracking_impl.cpp:406-425, dense_tracking.cpp:333-342, 448-476,
//                               least_squares.cpp:58-64)
//
// A warp owns a contiguous run of pixels (a "segment", row-major order) and walks it 32 pixels at a
// time.  All arithmetic that decides validity is explicit round-to-nearest fp32 in a fixed order
// (packed f32x2 where two channels share an operation), mirrored bit for bit by the oracle's MIRROR
// mode; sums use whatever contraction the compiler picks.
#pragma once
#include "common.cuh"
#include "f32x2.cuh"

namespace dvo_b200 {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kSegmentPixels = 256;     // pixels per warp segment on the test-hook path (8 rounds of 32); the persistent
                                        // kernel sizes its segments per level (LevelPlan in tracker.cu)
constexpr int kSegmentsPerTile = 4;     // warps per CTA: a CTA covers 4 consecutive segments

// record planes of one pair at one level (scratch): E = (e.i, e.z), G = (e.idx, e.idy), H = (e.zdx, e.zdy),
// Z = depth of the reference point (28 B per pixel)
struct RecordPlanes {
  float2* E; float2* G; float2* H; float* Z;
};
__host__ __device__ __forceinline__ RecordPlanes record_planes(float* base, size_t n) {
  RecordPlanes r;
  r.E = reinterpret_cast<float2*>(base);
  r.G = r.E + n;
  r.H = r.G + n;
  r.Z = reinterpret_cast<float*>(r.H + n);
  return r;
}
constexpr int kRecordFloatsPerPixel = 7;

// ---- per pair-iteration constants ---------------------------------------------------------------
struct StageConsts {
  f2 k0, k1, k2, k3;          // (kt[0],kt[4]) (kt[1],kt[5]) (kt[2],kt[6]) (kt[3],kt[7]): rows X and Y of K*T
  float k8, k9, k10, k11;     // row Z
  f2 Pa, Pb;                  // precision used for the weights: (P00,P01), (P10,P11)
  f2 cg, fxy;                 // (0.5 fx/255, 0.5 fy/255), (fx, fy)   (dense_tracking.cpp:215-220)
  float c_i, ubx, uby;
  int first_iteration;
  int drop_idx;               // linear index of the odd selected point that is skipped, or -1
};

__device__ __forceinline__ void load_stage_consts(const PairState& st, const PairLevel& pl, int w, int h, StageConsts& c) {
  // PairState is rewritten between stages by another SM (persistent kernel): read it through L2 (ld.cg)
  float kt[12], P[4];
#pragma unroll
  for (int i = 0; i < 12; ++i) kt[i] = __ldcg(&st.kt[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) P[i] = __ldcg(&st.precision[i]);
  c.k0 = pk(kt[0], kt[4]); c.k1 = pk(kt[1], kt[5]); c.k2 = pk(kt[2], kt[6]); c.k3 = pk(kt[3], kt[7]);
  c.k8 = kt[8]; c.k9 = kt[9]; c.k10 = kt[10]; c.k11 = kt[11];
  c.Pa = pk(P[0], P[1]); c.Pb = pk(P[2], P[3]);
  c.cg = pk(__fdiv_rn(__fmul_rn(0.5f, pl.cfx), 255.0f), __fdiv_rn(__fmul_rn(0.5f, pl.cfy), 255.0f));
  c.fxy = pk(pl.cfx, pl.cfy);
  c.c_i = 1.0f / 255.0f;
  c.ubx = (float)(w - 2); c.uby = (float)(h - 2);
  c.first_iteration = __ldcg(&st.iteration) == 0;
  int S = pl.rsel[0];
  c.drop_idx = (S & 1) ? pl.rsel[1] : -1;   // odd S: last selected point skipped (dense_tracking_impl.cpp:169)
}

// Reference-side inputs of one pixel, loaded one round ahead of their use.
struct RefPixel {
  f2 a;        // (I_r, Z_r)
  f2 g;        // (Ix_r, Iy_r)
  float tx, ty;
};

__device__ __forceinline__ RefPixel load_ref_pixel(const PairLevel& pl, int idx, int w, unsigned wmagic, int n) {
  RefPixel r;
  const int i = min(idx, n - 1);                   // lanes past the end of the image read a valid address
  { const float2 v = __ldcs(pl.r0 + i); r.a = pk(v.x, v.y); }   // reference planes are streamed once per iteration
  { const float2 v = __ldcs(pl.r1 + i); r.g = pk(v.x, v.y); }
  const int y = (int)__umulhi((unsigned)i, wmagic);   // i / w (exact for i*w < 2^32)
  const int x = i - y * w;
  r.tx = __ldg(pl.rtmpl + x);
  r.ty = __ldg(pl.rtmpl + w + y);
  return r;
}

// The residual record of one reference pixel (computeResidualsSse, dense_tracking_impl.cpp:133-393) is
// computed in three steps so that the twelve bilinear taps of round r+1 are in flight while round r
// is blended:
//   project_pixel : point (x,y,z) = (tx*z, ty*z, z); (X,Y,Z') = fma chains over the rows of K*T;
//                   (u,v) = (X,Y)*rcp_rn(Z'); bounds 0<=u<=w-2, 0<=v<=h-2; truncation -> tap index, weights
//   load_taps     : the four neighbours in the three float2 planes of the current image
//   finish_pixel  : bilinear blend, residual weights of dense_tracking.cpp:215-220, NaN test (line 261),
//                   occlusion test (line 275).  E = (e.i, e.z), G = (e.idx, e.idy), H = (e.zdx, e.zdy).
// Branch-free: a rejected point reads tap 0 and is flagged invalid.
struct PixelProjection {
  f2 f, gq;        // (fu, fv), (gu, gv)
  f2 g;            // (Ix_r, Iy_r)
  float Zt, z, Ir;
  int b;           // index of the upper-left tap
  bool inb;
};
struct PixelTaps {
  f2 p00, p10, p01, p11, q00, q10, q01, q11, s00, s10, s01, s11;
};

__device__ __forceinline__ PixelProjection project_pixel(const RefPixel& r, bool selected, int w, const StageConsts& c) {
  PixelProjection p;
  const float z = hi(r.a);
  const f2 pxy = mul2(pk(r.tx, r.ty), bc(z));
  const float px = lo(pxy), py = hi(pxy);
  const f2 XY = fma2(c.k0, bc(px), fma2(c.k1, bc(py), fma2(c.k2, bc(z), c.k3)));
  p.Zt = __fmaf_rn(c.k8, px, __fmaf_rn(c.k9, py, __fmaf_rn(c.k10, z, c.k11)));
  f2 uv = mul2(XY, bc(rcp_rn(p.Zt)));
  const float u = lo(uv), v = hi(uv);
  p.inb = selected && u >= 0.f && u <= c.ubx && v >= 0.f && v <= c.uby;   // NaN compares false
  uv = p.inb ? uv : 0ull;
  // truncation without conversions: for 0 <= t < 2^23, RZ(t + 2^23) carries floor(t) in its mantissa
  const f2 t = add2_rz(uv, bc(8388608.0f));
  p.f = sub2(uv, sub2(t, bc(8388608.0f)));
  p.gq = sub2(bc(1.0f), p.f);
  const int u0 = __float_as_int(lo(t)) - 0x4b000000, v0 = __float_as_int(hi(t)) - 0x4b000000;
  p.b = v0 * w + u0;
  p.g = r.g; p.z = z; p.Ir = lo(r.a);
  return p;
}

__device__ __forceinline__ PixelTaps load_taps(const PairLevel& pl, int b, int w) {
  PixelTaps t;
  t.p00 = ldg_f2(pl.c0 + b); t.p10 = ldg_f2(pl.c0 + b + 1); t.p01 = ldg_f2(pl.c0 + b + w); t.p11 = ldg_f2(pl.c0 + b + w + 1);
  t.q00 = ldg_f2(pl.c1 + b); t.q10 = ldg_f2(pl.c1 + b + 1); t.q01 = ldg_f2(pl.c1 + b + w); t.q11 = ldg_f2(pl.c1 + b + w + 1);
  t.s00 = ldg_f2(pl.c2 + b); t.s10 = ldg_f2(pl.c2 + b + 1); t.s01 = ldg_f2(pl.c2 + b + w); t.s11 = ldg_f2(pl.c2 + b + w + 1);
  return t;
}

__device__ __forceinline__ bool finish_pixel(const PixelProjection& p, const PixelTaps& t, const StageConsts& c, f2& E, f2& G, f2& H) {
  const float fu = lo(p.f), fv = hi(p.f), gu = lo(p.gq), gv = hi(p.gq);
#define DVO_BLEND2(c00, c10, c01, c11) \
  fma2(bc(fv), fma2(bc(fu), c11, mul2(bc(gu), c01)), mul2(bc(gv), fma2(bc(fu), c10, mul2(bc(gu), c00))))
  const f2 IZ = DVO_BLEND2(t.p00, t.p10, t.p01, t.p11);
  const f2 Gc = DVO_BLEND2(t.q00, t.q10, t.q01, t.q11);
  const f2 Hc = DVO_BLEND2(t.s00, t.s10, t.s01, t.s11);
#undef DVO_BLEND2
  const float Zc = hi(IZ);
  const float ez = __fsub_rn(Zc, p.Zt);
  const float s = __fsub_rn(p.z, 0.4f);
  const float sig = __fmaf_rn(__fmul_rn(0.0019f, s), s, 0.0012f);    // depthStdDevZ (lines 122-128)
  const float ei = __fmaf_rn(c.c_i, lo(IZ), __fmul_rn(-c.c_i, p.Ir));
  E = pk(ei, ez);
  G = fma2(c.cg, Gc, mul2(c.cg, p.g));
  H = mul2(c.fxy, Hc);
  return p.inb && Zc == Zc && ez > __fmul_rn(-20.0f, sig);
}

// ---- pairwise scale sum ---------------------------------------------------------------------------
// computeScaleSse (dense_tracking_impl.cpp:590-638) walks the compacted residual list two at a time
// and, because lines 614-615 re-use the low half of the register, adds (w_{2j} + w_{2j+1}) r_{2j} r_{2j}^T
// for every pair plus w_n r_n r_n^T for an odd tail.  That needs, per valid point, the parity of its
// rank in row-major order and the weight of the next valid point.  A contiguous run of pixels is
// summarised by a segment record: the sums under both hypotheses for the parity of its first point
// (S0: the first valid point is a pair leader, S1: it is a follower), its first valid weight and its
// last valid point (a leader whose partner lies in the next run).  Runs combine associatively.
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
  bool link0 = a.n > 0 && b.n > 0 && (((a.n - 1) & 1) == 0);   // hypothesis 0: last point of a is a leader
  bool link1 = a.n > 0 && b.n > 0 && ((a.n & 1) == 0);         // hypothesis 1
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

constexpr int kSegExportFloats = 12;   // n (as int bits), S0[3], S1[3], wf, wl, ol[3]
constexpr int kCtaExportFloats = 16;   // the CTA's four warp summaries combined (12) + the four warp counts (int bits)

__device__ __forceinline__ SegT<double> load_seg_export(const float* e) {
  SegT<double> s;
  // written by other SMs in the same kernel (persistent path): read through L2
  float v[kSegExportFloats];
#pragma unroll
  for (int i = 0; i < kSegExportFloats; ++i) v[i] = __ldcg(e + i);
  s.n = __float_as_int(v[0]);
  s.S0[0] = v[1]; s.S0[1] = v[2]; s.S0[2] = v[3];
  s.S1[0] = v[4]; s.S1[1] = v[5]; s.S1[2] = v[6];
  s.wf = v[7]; s.wl = v[8]; s.ol[0] = v[9]; s.ol[1] = v[10]; s.ol[2] = v[11];
  return s;
}

// Thread 0 of a CTA folds the four warp summaries (shared memory, in pixel order) into one CTA export.
__device__ __forceinline__ void cta_export_segments(const float (*sm_exp)[kSegExportFloats], float* out) {
  SegT<float> acc;
  {
    const float* e = sm_exp[0];
    acc.n = __float_as_int(e[0]);
    for (int k = 0; k < 3; ++k) { acc.S0[k] = e[1 + k]; acc.S1[k] = e[4 + k]; acc.ol[k] = e[9 + k]; }
    acc.wf = e[7]; acc.wl = e[8];
  }
  for (int q = 1; q < kSegmentsPerTile; ++q) {
    const float* e = sm_exp[q];
    SegT<float> b;
    b.n = __float_as_int(e[0]);
    for (int k = 0; k < 3; ++k) { b.S0[k] = e[1 + k]; b.S1[k] = e[4 + k]; b.ol[k] = e[9 + k]; }
    b.wf = e[7]; b.wl = e[8];
    acc = combine_seg<float>(acc, b);
  }
  out[0] = __int_as_float((int)acc.n);
  for (int k = 0; k < 3; ++k) { out[1 + k] = acc.S0[k]; out[4 + k] = acc.S1[k]; out[9 + k] = acc.ol[k]; }
  out[7] = acc.wf; out[8] = acc.wl;
  for (int q = 0; q < kSegmentsPerTile; ++q) out[12 + q] = sm_exp[q][0];
}

// Student-t weight of computeWeightsSse (dense_tracking_impl.cpp:657-707): w = 7 / (5 + r^T P r), nu = 5;
// w = 1 on the first iteration of a level (dense_tracking.cpp:286-289).
__device__ __forceinline__ float student_weight(bool first_iteration, f2 Pa, f2 Pb, float ei, float ez) {
  if (first_iteration) return 1.0f;
  const f2 q = fma2(bc(ez), Pb, mul2(bc(ei), Pa));        // (ei P00 + ez P10, ei P01 + ez P11)
  const float d = fmaf(lo(q), ei, hi(q) * ez);
  return 7.0f * rcp_fast(5.0f + d);
}
__device__ __forceinline__ float student_weight(const StageConsts& c, float ei, float ez) {
  return student_weight(c.first_iteration != 0, c.Pa, c.Pb, ei, ez);
}

__device__ __forceinline__ void store_record(const RecordPlanes& rec, int idx, int n, bool valid, f2 E, f2 G, f2 H, float z) {
  if (idx < n) {
    if (valid) {
      __stcs(rec.E + idx, make_float2(lo(E), hi(E)));    // st.global.cs: streamed, evict-first in L2
      __stcs(rec.G + idx, make_float2(lo(G), hi(G)));
      __stcs(rec.H + idx, make_float2(lo(H), hi(H)));
      __stcs(rec.Z + idx, z);
    } else {
      const float nanf_ = __int_as_float(0x7fc00000);
      __stcs(rec.E + idx, make_float2(nanf_, nanf_));
    }
  }
}

// Running state of the pairwise scale sum of one warp segment (see the comment above SegT).
struct ScaleState {
  // fp64 accumulators: the covariance is inverted and feeds accept/reject decisions, so the sums are kept
  // independent of how the pixels are partitioned over warps (to ~1e-12) at the cost of 12 DADD per 64 pixels
  double sall0, sall1, sall2, salt0, salt1, salt2;
  float pw, po0, po1, po2;   // pending leader: the last valid point seen, waiting for the next valid weight
  float wfirst;
  int psign, cnt;
  bool pend;
};

__device__ __forceinline__ void scale_state_init(ScaleState& s) {
  s.sall0 = s.sall1 = s.sall2 = s.salt0 = s.salt1 = s.salt2 = 0.0;
  s.pw = s.po0 = s.po1 = s.po2 = 0.f; s.wfirst = 0.f; s.psign = 0; s.cnt = 0; s.pend = false;
}

// Adds the 64 points {pixel base+lane: (v0, w0, ei0, ez0)} then {pixel base+32+lane: (v1, ...)} to the state.
__device__ __forceinline__ void scale_round64(ScaleState& st, int lane, bool v0, float w0, float ei0, float ez0,
                                              bool v1, float w1, float ei1, float ez1) {
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned m0 = __ballot_sync(kFullMask, v0), m1 = __ballot_sync(kFullMask, v1);
  if ((m0 | m1) == 0u) return;
  const float w1_first = __shfl_sync(kFullMask, w1, m1 ? __ffs(m1) - 1 : 0);        // first valid weight of the upper half
  const float w_first = m0 ? __shfl_sync(kFullMask, w0, __ffs(m0) - 1) : w1_first;   // first valid weight of the round
  if (st.cnt == 0) st.wfirst = w_first;
  {   // the pending leader of an earlier round pairs with the first valid point of this round
    const float s = st.pend ? st.pw + w_first : 0.f;
    const float sa = __int_as_float(__float_as_int(s) ^ st.psign);
    st.sall0 += (double)(s * st.po0); st.sall1 += (double)(s * st.po1); st.sall2 += (double)(s * st.po2);
    st.salt0 += (double)(sa * st.po0); st.salt1 += (double)(sa * st.po1); st.salt2 += (double)(sa * st.po2);
  }
  const unsigned above0 = (m0 >> lane) >> 1, above1 = (m1 >> lane) >> 1;
  float wn0 = __shfl_sync(kFullMask, w0, above0 ? lane + __ffs(above0) : lane);
  const float wn1 = __shfl_sync(kFullMask, w1, above1 ? lane + __ffs(above1) : lane);
  wn0 = above0 ? wn0 : w1_first;
  const bool next0 = above0 != 0u || m1 != 0u, next1 = above1 != 0u;
  const int c0n = __popc(m0);
  const int sg0 = ((st.cnt + __popc(m0 & lt_mask)) & 1) << 31;          // sign bit set for odd rank
  const int sg1 = ((st.cnt + c0n + __popc(m1 & lt_mask)) & 1) << 31;
  // rejected points carry garbage (possibly NaN) residuals: zero them so that 0 * outer stays 0
  const float xi0 = v0 ? ei0 : 0.f, xz0 = v0 ? ez0 : 0.f, xi1 = v1 ? ei1 : 0.f, xz1 = v1 ? ez1 : 0.f;
  const float a0 = xi0 * xi0, a1 = xi0 * xz0, a2 = xz0 * xz0;
  const float b0 = xi1 * xi1, b1 = xi1 * xz1, b2 = xz1 * xz1;
  {
    const float s = (v0 && next0) ? w0 + wn0 : 0.f;
    const float sa = __int_as_float(__float_as_int(s) ^ sg0);
    st.sall0 += (double)(s * a0); st.sall1 += (double)(s * a1); st.sall2 += (double)(s * a2);
    st.salt0 += (double)(sa * a0); st.salt1 += (double)(sa * a1); st.salt2 += (double)(sa * a2);
  }
  {
    const float s = (v1 && next1) ? w1 + wn1 : 0.f;
    const float sa = __int_as_float(__float_as_int(s) ^ sg1);
    st.sall0 += (double)(s * b0); st.sall1 += (double)(s * b1); st.sall2 += (double)(s * b2);
    st.salt0 += (double)(sa * b0); st.salt1 += (double)(sa * b1); st.salt2 += (double)(sa * b2);
  }
  // new pending leader: the last valid point of the round
  const bool np1 = v1 && !next1, np0 = v0 && !next0;
  st.pend = np0 || np1;
  st.pw = np1 ? w1 : (np0 ? w0 : st.pw);
  st.po0 = np1 ? b0 : (np0 ? a0 : st.po0); st.po1 = np1 ? b1 : (np0 ? a1 : st.po1); st.po2 = np1 ? b2 : (np0 ? a2 : st.po2);
  st.psign = np1 ? sg1 : (np0 ? sg0 : st.psign);
  st.cnt += c0n + __popc(m1);
}

// warp-reduce the sums and write the segment summary (kSegExportFloats floats)
__device__ __forceinline__ void scale_state_export(ScaleState& st, int lane, float* seg_out) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    st.sall0 += __shfl_xor_sync(kFullMask, st.sall0, off); st.sall1 += __shfl_xor_sync(kFullMask, st.sall1, off);
    st.sall2 += __shfl_xor_sync(kFullMask, st.sall2, off); st.salt0 += __shfl_xor_sync(kFullMask, st.salt0, off);
    st.salt1 += __shfl_xor_sync(kFullMask, st.salt1, off); st.salt2 += __shfl_xor_sync(kFullMask, st.salt2, off);
  }
  // leaders at even local rank belong to hypothesis 0, odd to hypothesis 1: S0 = (all + alt)/2, S1 = (all - alt)/2
  if (lane == 0) {
    seg_out[0] = __int_as_float(st.cnt);
    seg_out[1] = (float)(0.5 * (st.sall0 + st.salt0)); seg_out[2] = (float)(0.5 * (st.sall1 + st.salt1)); seg_out[3] = (float)(0.5 * (st.sall2 + st.salt2));
    seg_out[4] = (float)(0.5 * (st.sall0 - st.salt0)); seg_out[5] = (float)(0.5 * (st.sall1 - st.salt1)); seg_out[6] = (float)(0.5 * (st.sall2 - st.salt2));
    seg_out[7] = st.wfirst;
    if (st.cnt == 0) { seg_out[8] = 0.f; seg_out[9] = 0.f; seg_out[10] = 0.f; seg_out[11] = 0.f; }
  }
  if (st.pend) {   // exactly one lane when cnt > 0: the last valid point of the segment
    seg_out[8] = st.pw; seg_out[9] = st.po0; seg_out[10] = st.po1; seg_out[11] = st.po2;
  }
}

// Stage A over the pixels [begin, end) of one pair (begin a multiple of 32): writes the residual
// records and the segment summary (kSegExportFloats floats at `seg_out`).  The warp walks 32 pixels per
// round as a two-deep software pipeline: while round r is blended, the twelve taps of round r+1 and the
// reference-side loads of round r+2 are in flight.  The pairwise scale sum runs once per two rounds.
__device__ __forceinline__ void stage_a_segment(const PairLevel& pl, const StageConsts& c, int w, unsigned wmagic, int n,
                                                int begin, int end, const RecordPlanes& rec, float* seg_out) {
  const int lane = threadIdx.x & 31;
  ScaleState ss;
  scale_state_init(ss);
  if (begin < end) {
    // prologue: project round 0 and issue its taps, load the reference data of round 1
    RefPixel ref = load_ref_pixel(pl, begin + lane, w, wmagic, n);
    unsigned sel = __ldg(pl.rmask + (begin >> 5));
    PixelProjection proj = project_pixel(ref, ((sel >> lane) & 1u) && (begin + lane) != c.drop_idx, w, c);
    PixelTaps taps = load_taps(pl, proj.b, w);
    ref = load_ref_pixel(pl, begin + 32 + lane, w, wmagic, n);
    sel = begin + 32 < end ? __ldg(pl.rmask + (begin >> 5) + 1) : 0u;
    bool sv = false; float sw = 0.f, sei = 0.f, sez = 0.f;   // stashed even round
    bool odd = false;
#pragma unroll 1
    for (int base = begin; base < end; base += 32) {
      const PixelProjection pcur = proj;
      const PixelTaps tcur = taps;
      const int nb = base + 32;
      if (nb < end) {   // warp-uniform
        proj = project_pixel(ref, ((sel >> lane) & 1u) && (nb + lane) != c.drop_idx, w, c);
        taps = load_taps(pl, proj.b, w);
        ref = load_ref_pixel(pl, nb + 32 + lane, w, wmagic, n);
        sel = nb + 32 < end ? __ldg(pl.rmask + (nb >> 5) + 1) : 0u;
      }
      f2 E, G, H;
      const bool v = finish_pixel(pcur, tcur, c, E, G, H);
      const float ei = lo(E), ez = hi(E);
      const float wgt = student_weight(c, ei, ez);
      // 4th plane = reference depth: stage B recomputes the weight (6 flops) instead of reading it plus the
      // reference plane again.  end <= n: never touch another warp's pixels
      store_record(rec, base + lane, end, v, E, G, H, pcur.z);
      if (!odd) { sv = v; sw = wgt; sei = ei; sez = ez; }
      else scale_round64(ss, lane, sv, sw, sei, sez, v, wgt, ei, ez);
      odd = !odd;
    }
    if (odd) scale_round64(ss, lane, sv, sw, sei, sez, false, 0.f, 0.f, 0.f);
  }
  scale_state_export(ss, lane, seg_out);
}

// ---- stage B -----------------------------------------------------------------------------------------
struct StageBConsts {
  float P00, P01, P10, P11;   // P_k
  float l, wd0, wd1;          // P_k = [1 l; 0 1]^T-style factors, see stage_b_segment
  f2 Pa, Pb;                  // P_{k-1} as stage A used it for the weights
  bool first_iteration;
};

constexpr int kNormalValues = 28;   // log-likelihood sum, 21 upper-triangular A (row-major), 6 b

// Accumulators of stage B for one thread.  A is kept as pairs of adjacent columns of one row
// (A[r][2c], A[r][2c+1]); rows 1, 3 and 5 carry one redundant lower-triangle entry so that every
// update is a packed FMA of a broadcast row factor with a column pair.
struct StageBAcc {
  f2 r0[3], r1[3], r2[2], r3[2], r4, r5;   // 12 pairs
  f2 b[3];
  float prod;      // running product of (1 + 0.2 r^T P r) over this thread's kept points
  float llsum;     // sum of logs flushed so far
};

__device__ __forceinline__ void stage_b_init(StageBAcc& a) {
#pragma unroll
  for (int i = 0; i < 3; ++i) { a.r0[i] = 0; a.r1[i] = 0; a.b[i] = 0; }
  a.r2[0] = a.r2[1] = a.r3[0] = a.r3[1] = a.r4 = a.r5 = 0;
  a.prod = 1.0f; a.llsum = 0.f;
}

// A += u v^T (upper triangle, column pairs) and b += u * s for one 6-vector given as three pairs V,
// with u = V * wd.
__device__ __forceinline__ void stage_b_rank1(StageBAcc& acc, const f2 V[3], float wd, float s) {
  const f2 U0 = mul2(V[0], bc(wd)), U1 = mul2(V[1], bc(wd)), U2 = mul2(V[2], bc(wd));
  const float u0 = lo(U0), u1 = hi(U0), u2 = lo(U1), u3 = hi(U1), u4 = lo(U2), u5 = hi(U2);
  acc.r0[0] = fma2(bc(u0), V[0], acc.r0[0]); acc.r0[1] = fma2(bc(u0), V[1], acc.r0[1]); acc.r0[2] = fma2(bc(u0), V[2], acc.r0[2]);
  acc.r1[0] = fma2(bc(u1), V[0], acc.r1[0]); acc.r1[1] = fma2(bc(u1), V[1], acc.r1[1]); acc.r1[2] = fma2(bc(u1), V[2], acc.r1[2]);
  acc.r2[0] = fma2(bc(u2), V[1], acc.r2[0]); acc.r2[1] = fma2(bc(u2), V[2], acc.r2[1]);
  acc.r3[0] = fma2(bc(u3), V[1], acc.r3[0]); acc.r3[1] = fma2(bc(u3), V[2], acc.r3[1]);
  acc.r4 = fma2(bc(u4), V[2], acc.r4);
  acc.r5 = fma2(bc(u5), V[2], acc.r5);
  acc.b[0] = fma2(U0, bc(s), acc.b[0]); acc.b[1] = fma2(U1, bc(s), acc.b[1]); acc.b[2] = fma2(U2, bc(s), acc.b[2]);
}

// Stage B over the pixels [begin, end): log-likelihood terms and normal equations with W = w * P_k.
// rank_base: number of valid points before `begin` in row-major order; points with rank >= n_keep are
// the dropped tail of computeCompleteDataLogLikelihood (dense_tracking_impl.cpp:413-422).
//
// With l = P01/P00, d0 = P00, d1 = P11 - P01^2/P00:
//   J^T P J = d0 j0' j0'^T + d1 J1 J1^T,  j0' = J0 + l J1,     J^T P r = d0 j0' (r0 + l r1) + d1 J1 r1
// so each point contributes two rank-1 updates.  J rows at the untransformed reference point
// (dense_tracking.cpp:448-476): J0 = gx a + gy b, J1 = hx a + hy b - c with
//   a = [1/z, 0, -x/z^2, a2 y, 1 - a2 x, -y/z], b = [0, 1/z, -y/z^2, b2 y - 1, -a3, x/z], c = [0, 0, 1, y, -x, 0].
struct StageBInput {   // everything stage B reads for one pixel; loaded two rounds ahead of its use
  float2 e, g, h;
  float z, tx, ty;
};

__device__ __forceinline__ StageBInput load_stage_b_input(const PairLevel& pl, const RecordPlanes& rec, int idx, int w,
                                                          unsigned wmagic, int limit) {
  StageBInput in;
  const int i = min(idx, limit - 1);
  in.e = __ldcs(rec.E + i);      // ld.global.cs: read once, do not keep in L2
  in.g = __ldcs(rec.G + i);
  in.h = __ldcs(rec.H + i);
  in.z = __ldcs(rec.Z + i);      // the reference depth as stage A left it
  const int y = (int)__umulhi((unsigned)i, wmagic);
  const int x = i - y * w;
  in.tx = __ldg(pl.rtmpl + x);
  in.ty = __ldg(pl.rtmpl + w + y);
  if (idx >= limit) in.e.x = __int_as_float(0x7fc00000);
  return in;
}

__device__ __forceinline__ void stage_b_segment(const PairLevel& pl, const StageBConsts& c, int w, unsigned wmagic, int n,
                                                int begin, int end, const RecordPlanes& rec, long long rank_base,
                                                long long n_keep, bool need_rank, StageBAcc& acc) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  int seen = 0;
  if (begin >= end) return;
  // software pipeline: the loads of rounds r+1 and r+2 are in flight while round r is consumed
  StageBInput in0 = load_stage_b_input(pl, rec, begin + lane, w, wmagic, end);
  StageBInput in1 = load_stage_b_input(pl, rec, begin + 32 + lane, w, wmagic, end);
#pragma unroll 1
  for (int base = begin; base < end; base += 32) {
    const StageBInput in = in0;
    in0 = in1;
    in1 = load_stage_b_input(pl, rec, base + 64 + lane, w, wmagic, end);
    const bool valid = in.e.x == in.e.x;
    bool keep = valid;
    if (need_rank) {   // warp-uniform: only the segments that contain the tail of the point list
      const unsigned m = __ballot_sync(kFullMask, valid);
      keep = valid && (rank_base + seen + __popc(m & lt_mask)) < n_keep;
      seen += __popc(m);
    }
    if (!valid) continue;
    const float ei = in.e.x, ez = in.e.y;
    // log-likelihood term: log(1 + 0.2 r^T P r), accumulated as a product
    const float d = (ei * c.P00 + ez * c.P10) * ei + (ei * c.P01 + ez * c.P11) * ez;
    if (keep) {
      acc.prod *= fmaf(0.2f, d, 1.0f);
      if (acc.prod > 1e18f) { acc.llsum += __logf(acc.prod); acc.prod = 1.0f; }   // keep the product in range
    }
    const float z = in.z;
    const float px = in.tx * z, py = in.ty * z;
    const float zi = rcp_fast(z), zs = zi * zi;
    const float a2 = -px * zs, b2 = -py * zs;
    const float a3 = a2 * py;
    const f2 A23 = pk(a2, a3), B23 = pk(b2, fmaf(b2, py, -1.0f));
    const f2 A45 = pk(fmaf(-a2, px, 1.0f), -py * zi), B45 = pk(-a3, px * zi);
    const f2 NC23 = pk(-1.0f, -py), NC45 = pk(px, 0.0f);     // -c[2..3], -c[4..5]
    const f2 G = pk(in.g.x, in.g.y), H = pk(in.h.x, in.h.y);
    const f2 Gp = fma2(bc(c.l), H, G);                        // (gx + l hx, gy + l hy)
    const float gx = lo(Gp), gy = hi(Gp);
    f2 V0[3], V1[3];
    V0[0] = mul2(Gp, bc(zi));
    V0[1] = fma2(bc(gx), A23, fma2(bc(gy), B23, mul2(bc(c.l), NC23)));
    V0[2] = fma2(bc(gx), A45, fma2(bc(gy), B45, mul2(bc(c.l), NC45)));
    V1[0] = mul2(H, bc(zi));
    V1[1] = fma2(bc(in.h.x), A23, fma2(bc(in.h.y), B23, NC23));
    V1[2] = fma2(bc(in.h.x), A45, fma2(bc(in.h.y), B45, NC45));
    // b -= J^T W r
    const float wgt = student_weight(c.first_iteration, c.Pa, c.Pb, ei, ez);   // same operands, same operations as stage A
    stage_b_rank1(acc, V0, wgt * c.wd0, -fmaf(c.l, ez, ei));
    stage_b_rank1(acc, V1, wgt * c.wd1, -ez);
  }
}

// flush the product of the pending log-likelihood terms and unpack: out[0] = ll sum,
// out[1..21] = A upper triangle (row-major), out[22..27] = b
__device__ __forceinline__ void stage_b_values(const StageBAcc& acc, float out[kNormalValues]) {
  out[0] = acc.llsum + __logf(acc.prod);
  out[1] = lo(acc.r0[0]); out[2] = hi(acc.r0[0]); out[3] = lo(acc.r0[1]); out[4] = hi(acc.r0[1]); out[5] = lo(acc.r0[2]); out[6] = hi(acc.r0[2]);
  out[7] = hi(acc.r1[0]); out[8] = lo(acc.r1[1]); out[9] = hi(acc.r1[1]); out[10] = lo(acc.r1[2]); out[11] = hi(acc.r1[2]);
  out[12] = lo(acc.r2[0]); out[13] = hi(acc.r2[0]); out[14] = lo(acc.r2[1]); out[15] = hi(acc.r2[1]);
  out[16] = hi(acc.r3[0]); out[17] = lo(acc.r3[1]); out[18] = hi(acc.r3[1]);
  out[19] = lo(acc.r4); out[20] = hi(acc.r4);
  out[21] = hi(acc.r5);
  out[22] = lo(acc.b[0]); out[23] = hi(acc.b[0]); out[24] = lo(acc.b[1]); out[25] = hi(acc.b[1]); out[26] = lo(acc.b[2]); out[27] = hi(acc.b[2]);
}

__device__ __forceinline__ void load_stage_b_consts(const PairState& st, StageBConsts& c) {
  c.P00 = __ldcg(&st.precision[0]); c.P01 = __ldcg(&st.precision[1]); c.P10 = __ldcg(&st.precision[2]); c.P11 = __ldcg(&st.precision[3]);
  c.l = c.P01 / c.P00;
  c.wd0 = c.P00;
  c.wd1 = c.P11 - c.P01 * c.l;
  c.Pa = pk(__ldcg(&st.precision_prev[0]), __ldcg(&st.precision_prev[1]));
  c.Pb = pk(__ldcg(&st.precision_prev[2]), __ldcg(&st.precision_prev[3]));
  c.first_iteration = __ldcg(&st.iteration) == 0;
}

}  // namespace dvo_b200
