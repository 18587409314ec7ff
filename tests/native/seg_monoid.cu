// Host-side check of the segment monoid the level kernel uses for the pairwise scale sum (stages.cuh: SegT, combine_seg).
// computeScaleSse (dense_tracking_impl.cpp:590-638) walks the compacted list of valid points two at a time and adds
// (w_{2j} + w_{2j+1}) r_{2j} r_{2j}^T per pair, plus w_n r_n r_n^T for an odd tail.  The kernel summarises runs of points and
// combines the summaries in pixel order; here single-point summaries are combined (a) left to right, (b) through random
// binary trees, (c) strip-wise then across strips like the kernel, and compared with the sequential walk.
// Built and run by tests/test_host.py (nvcc, host code only: no GPU needed).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "stages.cuh"

using dvo_b200::SegT;
using dvo_b200::combine_seg;
typedef SegT<double> Seg;

static unsigned long long rng_state = 88172645463325252ull;
static double uniform() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (double)(rng_state >> 11) / 9007199254740992.0; }

static Seg empty() { Seg s; s.n = 0; s.wf = s.wl = 0; for (int k = 0; k < 3; ++k) s.S0[k] = s.S1[k] = s.ol[k] = 0; return s; }
static Seg single(double w, double ri, double rz) {
  Seg s = empty();
  s.n = 1; s.wf = w; s.wl = w; s.ol[0] = ri * ri; s.ol[1] = ri * rz; s.ol[2] = rz * rz;
  return s;
}
static Seg tree(const std::vector<Seg>& v, size_t lo, size_t hi) {   // random split points: any parenthesisation
  if (hi - lo == 0) return empty();
  if (hi - lo == 1) return v[lo];
  const size_t mid = lo + 1 + (size_t)(uniform() * (double)(hi - lo - 1));
  return combine_seg<double>(tree(v, lo, mid), tree(v, mid, hi));
}
static void total(const Seg& s, double out[3]) {                      // pair_mid_warp: tail term for an odd count
  const bool tail = s.n > 0 && ((s.n - 1) & 1) == 0;
  for (int k = 0; k < 3; ++k) out[k] = s.S0[k] + (tail ? s.wl * s.ol[k] : 0.0);
}

int main() {
  int bad = 0;
  for (int trial = 0; trial < 200; ++trial) {
    const int npix = 1 + (int)(uniform() * 700);
    const double pvalid = trial % 5 == 0 ? 0.05 : (trial % 5 == 1 ? 1.0 : uniform());
    std::vector<Seg> pts;
    std::vector<double> w, ri, rz;
    for (int i = 0; i < npix; ++i) {
      if (uniform() <= pvalid) {
        w.push_back(0.1 + uniform()); ri.push_back(uniform() - 0.5); rz.push_back(0.2 * (uniform() - 0.5));
        pts.push_back(single(w.back(), ri.back(), rz.back()));
      } else {
        pts.push_back(empty());                                      // a rejected pixel: the empty run
      }
    }
    // the reference's walk over the compacted list
    double ref[3] = {0, 0, 0};
    const size_t n = w.size();
    for (size_t j = 0; j + 1 < n; j += 2) { const double s = w[j] + w[j + 1]; ref[0] += s * ri[j] * ri[j]; ref[1] += s * ri[j] * rz[j]; ref[2] += s * rz[j] * rz[j]; }
    if (n & 1) { const size_t j = n - 1; ref[0] += w[j] * ri[j] * ri[j]; ref[1] += w[j] * ri[j] * rz[j]; ref[2] += w[j] * rz[j] * rz[j]; }
    // (a) left to right
    Seg a = empty();
    for (size_t i = 0; i < pts.size(); ++i) a = combine_seg<double>(a, pts[i]);
    // (b) a random parenthesisation
    const Seg b = tree(pts, 0, pts.size());
    // (c) rows of 32 pixels, strips of 7 rows in order, strips in chunks per "lane" then an in-order tree (the kernel's order)
    std::vector<Seg> rows, strips;
    for (size_t i = 0; i < pts.size(); i += 32) { Seg r = empty(); for (size_t k = i; k < pts.size() && k < i + 32; ++k) r = combine_seg<double>(r, pts[k]); rows.push_back(r); }
    for (size_t i = 0; i < rows.size(); i += 7) { Seg s = empty(); for (size_t k = i; k < rows.size() && k < i + 7; ++k) s = combine_seg<double>(s, rows[k]); strips.push_back(s); }
    std::vector<Seg> lanes(32, empty());
    const size_t chunk = (strips.size() + 31) / 32;
    for (size_t l = 0; l < 32; ++l) for (size_t k = l * chunk; k < strips.size() && k < (l + 1) * chunk; ++k) lanes[l] = combine_seg<double>(lanes[l], strips[k]);
    for (int off = 1; off < 32; off <<= 1) for (int l = 0; l < 32; l += 2 * off) lanes[l] = combine_seg<double>(lanes[l], lanes[l + off]);
    const Seg c = lanes[0];
    const Seg* cand[3] = {&a, &b, &c};
    for (int m = 0; m < 3; ++m) {
      double got[3];
      total(*cand[m], got);
      bool ok = cand[m]->n == (long long)n;
      for (int k = 0; k < 3; ++k) ok = ok && std::fabs(got[k] - ref[k]) <= 1e-12 * (1.0 + std::fabs(ref[k]));
      if (!ok) { ++bad; std::printf("trial %d order %d: n %lld vs %zu, %.17g %.17g %.17g vs %.17g %.17g %.17g\n", trial, m, cand[m]->n, n, got[0], got[1], got[2], ref[0], ref[1], ref[2]); }
    }
  }
  std::printf(bad ? "FAILED %d\n" : "ok\n", bad);
  return bad ? 1 : 0;
}
