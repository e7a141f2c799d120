// estimate.cpp -- builds and validates the level model of estimate.h on the host.
#include "estimate.h"

#include <algorithm>
#include <cstring>
#include <limits>

#include "iwb_internal.h"

namespace {

double hi_to_double(int hi) {
  const uint64_t b = (uint64_t)(uint32_t)hi << 32;
  double d;
  memcpy(&d, &b, 8);
  return d;
}
int hi_of(double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  return (int)(uint32_t)(b >> 32);
}
// linear -> target curve WITHOUT the snap of values just below 1 (reference: src/imagew-main.c:74-96); only the shape
// matters here, the table decides what the model is measured against
double forward(double v, int cstype) {
  if (cstype == IW_CSTYPE_REC709) return v < 0.020 ? 4.5 * v : 1.099 * pow(v, 0.45) - 0.099;
  return v <= 0.0031308 ? 12.92 * v : 1.055 * pow(v, 1.0 / 2.4) - 0.055;
}

}  // namespace

bool iwb_build_est_table(int cstype, const std::vector<double> &thr, IwbEstTable *out) {
  if (cstype != IW_CSTYPE_SRGB && cstype != IW_CSTYPE_REC709) return false;   // (pure power laws have no linear toe: the
  const int top = (int)thr.size() - 1;                                        //  code-space and linear-space midpoints part)
  if (top < 255 || top > 65535) return false;
  std::vector<double> mid(top);
  for (int q = 0; q < top; q++) mid[q] = (thr[q] + thr[q + 1]) * 0.5;
  if (!std::is_sorted(mid.begin(), mid.end()) || !(mid[top - 1] < 0.999999999) || !(thr[top] <= 1.0)) return false;

  out->rec.assign(IWB_EST_NREC, IwbEstRec{});
  const double nan = std::numeric_limits<double>::quiet_NaN();
  std::vector<double> err(IWB_EST_NREC, 0.0);
  const int nseg = IWB_EST_NREC - 1;
  for (int s = 0; s < nseg; s++) {
    const int hi = IWB_EST_HI_MIN + (s << IWB_EST_SHIFT);
    const long double a = hi_to_double(hi), b = hi_to_double(hi + (1 << IWB_EST_SHIFT));
    // quadratic through the three Chebyshev nodes of the segment, in u = v - a
    long double u[3], f[3];
    for (int j = 0; j < 3; j++) {
      u[j] = (b - a) * (0.5L + 0.5L * cosl((2 * j + 1) * 3.14159265358979323846L / 6.0L));
      f[j] = (long double)forward((double)(a + u[j]), cstype) * top;
    }
    const long double d01 = (f[1] - f[0]) / (u[1] - u[0]), d12 = (f[2] - f[1]) / (u[2] - u[1]);
    const long double d012 = (d12 - d01) / (u[2] - u[0]);
    IwbEstRec r;
    r.c2 = (float)d012;
    r.c1 = (float)(d01 - d012 * (u[0] + u[1]));
    long double c0 = 0;     // with c1, c2 as stored: the constant that fits the nodes best
    for (int j = 0; j < 3; j++) c0 += f[j] - ((long double)r.c2 * u[j] + (long double)r.c1) * u[j];
    r.c0 = (double)(c0 / 3 + 0.5L);
    out->rec[s] = r;
  }
  out->rec[nseg].c0 = (double)top + 0.5;    // v >= 1 (clamped by the reference): the top code
  out->rec[nseg].c1 = out->rec[nseg].c2 = 0.0f;

  // Validation.  Segment s decides samples of [start, end) (segment 0: everything below its end); the midpoints that
  // bound those samples' code intervals are the ones inside plus the nearest one on either side.
  out->invalid = 0;
  out->worst = 0;
  for (int s = 0; s < nseg; s++) {
    const int hi = IWB_EST_HI_MIN + (s << IWB_EST_SHIFT);
    const double a = s == 0 ? -1.0 : hi_to_double(hi), b = hi_to_double(hi + (1 << IWB_EST_SHIFT));
    int qa = (int)(std::lower_bound(mid.begin(), mid.end(), a) - mid.begin());
    int qb = (int)(std::lower_bound(mid.begin(), mid.end(), b) - mid.begin());
    qa = std::max(qa - 1, 0);
    qb = std::min(qb, top - 1);
    const IwbEstRec &r = out->rec[s];
    double e = 0;
    for (int q = qa; q <= qb; q++) {
      const double g = iwb_est_eval_rec(r, mid[q], hi);
      e = std::max(e, std::fabs(g - (double)(q + 1)));
    }
    // increasing over everything it is asked about: P' = c1 + 2 c2 u is linear in u
    const double v0 = hi_to_double(hi);
    const double ulo = std::min(s == 0 ? 0.0 : a, mid[qa]) - v0, uhi = std::max(b, mid[qb]) - v0;
    const bool rising = (double)r.c1 + 2.0 * (double)r.c2 * ulo > 0.0 && (double)r.c1 + 2.0 * (double)r.c2 * uhi > 0.0;
    if (!(e < 4e-3) || !rising) {
      out->rec[s].c0 = nan;
      out->invalid++;
    } else {
      out->worst = std::max(out->worst, e);
    }
    err[s] = e;
  }
  // slack: the single-precision inner product of the model is not exactly monotone between midpoints -- its rounding is
  // below 2^-22 of the segment's span in codes (< 300 codes at 16 bits) -- plus the double-precision roundings
  out->margin = out->worst + 3e-4;
  // the model must be worth having: most segments valid
  return out->invalid * 8 < nseg;
}

// test hook (include/imagew_b200.h)
extern "C" int iwb200_debug_level_model(int cstype, const double *thr, int ntab, const double *v, int n, int *code, double *stats) {
  IwbEstTable t;
  if (!iwb_build_est_table(cstype, std::vector<double>(thr, thr + ntab), &t)) return 0;
  if (stats) { stats[0] = t.margin; stats[1] = t.worst; stats[2] = (double)t.invalid; }
  const int top = ntab - 1;
  for (int i = 0; i < n; i++) {
    // the kernel's steps (gather4_dev.cuh: emit_est), on the host
    double x = v[i];
    int h = hi_of(x);
    uint64_t b;
    memcpy(&b, &x, 8);
    b = ((uint64_t)(uint32_t)(h & ~(h >> 31)) << 32) | (b & 0xffffffffu);
    memcpy(&x, &b, 8);
    const int hc = iwb_est_clamped_hi(hi_of(x));
    const double g = iwb_est_eval_rec(t.rec[iwb_est_index(hc)], x, hc);
    const double tt = g + IWB_EST_MAGIC;
    const double d = g - (tt - IWB_EST_MAGIC);
    uint64_t tb;
    memcpy(&tb, &tt, 8);
    const int cd = (int)(uint32_t)tb + (hi_of(d) >> 31);
    const bool slow = !(std::fabs(d) > t.margin) || (unsigned)cd > (unsigned)top;
    code[i] = slow ? -1 : cd;
  }
  return 1;
}
