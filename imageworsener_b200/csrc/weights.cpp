// weights.cpp -- per-dimension resampling recipes (filter kernels -> per-output tap lists), built on the host.
//
// Restates the behaviour of the reference's weight-list builder (reference: src/imagew-resize.c:62-185 filter
// kernels, :240-344 iw_create_weightlist_std, :373-410 nearest/null, :412-556 iwpvt_resize_rows_init) with a
// different data layout: a CSR-style table {first[], src[], w[]} in the reference's list order, so that a
// device thread can replay "out = 0; out += in[src]*w" tap by tap in exactly the reference's order.
#include <cmath>

#include "iwb_internal.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

struct filter_spec {
  int family;
  double radius;
  double b, c;   // cubic
  double mix;    // pixel mixing
};

double sinc(double x) { return x <= 0.000000005 ? 1.0 : sin(M_PI * x) / (M_PI * x); }

// Filter value at distance x >= 0 (all filters are symmetric).
double filter_value(const filter_spec &f, double x) {
  switch (f.family) {
    case IW_RESIZETYPE_LANCZOS:
      return x < f.radius ? sinc(x) * sinc(x / f.radius) : 0.0;
    case IW_RESIZETYPE_HANNING:
      return x < f.radius ? sinc(x) * (0.5 * cos(M_PI * x / f.radius) + 0.5) : 0.0;
    case IW_RESIZETYPE_BLACKMAN:
      return x < f.radius
                 ? sinc(x) * (0.5 * cos(M_PI * x / f.radius) + 0.08 * cos(2.0 * M_PI * x / f.radius) + 0.42)
                 : 0.0;
    case IW_RESIZETYPE_SINC:
      return x < f.radius ? sinc(x) : 0.0;
    case IW_RESIZETYPE_GAUSSIAN: {
      if (x >= 2.0) return 0.0;
      double v = exp(-2.0 * x * x) * 0.79788456080286535587989;
      return x <= 1.999 ? v : 1000.0 * (2.0 - x) * v;  // taper to 0 at the cut-off
    }
    case IW_RESIZETYPE_TRIANGLE:
      return x < 1.0 ? 1.0 - x : 0.0;
    case IW_RESIZETYPE_QUADRATIC:
      if (x < 0.5) return 0.75 - x * x;
      if (x < 1.5) return 0.50 * (x - 1.5) * (x - 1.5);
      return 0.0;
    case IW_RESIZETYPE_CUBIC: {
      double b = f.b, c = f.c;
      if (x < 1.0)
        return ((12.0 - 9.0 * b - 6.0 * c) * x * x * x + (-18.0 + 12.0 * b + 6.0 * c) * x * x + (6.0 - 2.0 * b)) / 6.0;
      if (x < 2.0)
        return ((-b - 6.0 * c) * x * x * x + (6.0 * b + 30.0 * c) * x * x + (-12.0 * b - 48.0 * c) * x +
                (8.0 * b + 24.0 * c)) / 6.0;
      return 0.0;
    }
    case IW_RESIZETYPE_HERMITE:
      return x < 1.0 ? 2.0 * x * x * x - 3.0 * x * x + 1.0 : 0.0;
    case IW_RESIZETYPE_BOX:
      return x <= 0.5 ? 1.0 : 0.0;
    case IW_RESIZETYPE_BOXAVG:
      if (x < 0.4999999) return 1.0;
      return x <= 0.5000001 ? 0.5 : 0.0;
    case IW_RESIZETYPE_MIX:
      if (x < (0.5 - f.mix / 2.0)) return 1.0;
      if (x < (0.5 + f.mix / 2.0)) return 0.5 - (x - 0.5) / f.mix;
      return 0.0;
  }
  return 0.0;
}

inline double clampd(double v, double lo, double hi) { return v < lo ? lo : v > hi ? hi : v; }

}  // namespace

int iwb_axis_table::min_src(int o0, int o1) const {
  int m = n_in;
  if (mode == 1) {
    for (int o = o0; o < o1; o++)
      if (map[o] >= 0 && map[o] < m) m = map[o];
  } else {
    for (int t = first[o0]; t < first[o1]; t++)
      if (src[t] >= 0 && src[t] < m) m = src[t];
  }
  return m;
}
int iwb_axis_table::max_src(int o0, int o1) const {
  int m = -1;
  if (mode == 1) {
    for (int o = o0; o < o1; o++)
      if (map[o] > m) m = map[o];
  } else {
    for (int t = first[o0]; t < first[o1]; t++)
      if (src[t] > m) m = src[t];
  }
  return m;
}
int iwb_axis_table::max_taps() const {
  int m = mode == 1 ? 1 : 0;
  if (mode == 0)
    for (int o = 0; o < n_out; o++)
      if (first[o + 1] - first[o] > m) m = first[o + 1] - first[o];
  return m;
}

int iwb_build_axis_table(iw_context *ctx, int dimension, int channeltype, int n_in, int n_out,
                         iwb_axis_table *out) {
  const iwb_axis_request &rq = ctx->axis[dimension];
  iwb_axis_table &T = *out;
  T = iwb_axis_table();
  T.n_in = n_in;
  T.n_out = n_out;

  filter_spec f{rq.family, 0.0, 0.0, 0.0, 0.0};
  bool standard = true, boxhack = false, sincbased = false;
  switch (rq.family) {
    case IW_RESIZETYPE_NULL: standard = false; break;
    case IW_RESIZETYPE_NEAREST: standard = false; boxhack = true; break;
    case IW_RESIZETYPE_MIX:
      f.mix = ((double)n_out) / n_in;
      if (f.mix > 1.0) f.mix = 1.0 / f.mix;
      f.radius = 0.5 + f.mix;
      break;
    case IW_RESIZETYPE_BOX: f.radius = 0.501; boxhack = true; break;
    case IW_RESIZETYPE_BOXAVG: f.radius = 0.501; break;
    case IW_RESIZETYPE_TRIANGLE: f.radius = 1.0; break;
    case IW_RESIZETYPE_QUADRATIC: f.radius = 1.5; break;
    case IW_RESIZETYPE_GAUSSIAN: f.radius = 2.0; break;
    case IW_RESIZETYPE_HERMITE: f.radius = 1.0; break;
    case IW_RESIZETYPE_CUBIC:
      f.radius = 2.0;
      f.b = clampd(rq.param1, -100.0, 100.0);
      f.c = clampd(rq.param2, -100.0, 100.0);
      break;
    case IW_RESIZETYPE_LANCZOS:
    case IW_RESIZETYPE_HANNING:
    case IW_RESIZETYPE_BLACKMAN:
    case IW_RESIZETYPE_SINC: sincbased = true; break;
    default:
      iw_set_error(ctx, "Internal: Unknown resize algorithm");
      return 0;
  }
  if (sincbased) f.radius = clampd(floor(rq.param1 + 0.5), 2.0, 100.0);  // "lobes"

  // Value standing in for a transparent virtual pixel: 0, or this channel's background when the background
  // is composited before resizing.  (The reference indexes its per-intermediate-channel table by channel
  // TYPE here, src/imagew-resize.c:531; the two coincide for R,G,B, the only types that reach this branch
  // with a non-zero colour.)
  T.edge_value = 0.0;
  if (rq.edge_policy == IW_EDGE_POLICY_TRANSPARENT && ctx->apply_bkgd && ctx->bkgd_strategy == IWB_BKGD_EARLY &&
      channeltype >= 0 && channeltype < 4)
    T.edge_value = ctx->mid_ci[channeltype].early_bkgd;

  double blur = clampd(rq.blur, 0.0001, 10000.0);
  double offset = rq.translate;
  if (boxhack) offset -= 0.00000000001;  // keep samples off box boundaries
  if (rq.use_offset && channeltype >= 0 && channeltype <= 2) offset += rq.channel_offset[channeltype];

  if (!standard) {
    T.mode = 1;
    T.map.resize(n_out);
    if (rq.family == IW_RESIZETYPE_NULL) {
      for (int o = 0; o < n_out; o++) T.map[o] = o < n_in ? o : -2;
    } else {  // nearest: ignores blur and out_true_size by design
      for (int o = 0; o < n_out; o++) {
        double centre = (0.5 + (double)o - offset) / (double)n_out;
        int p = (int)floor(centre * (double)n_in);
        T.map[o] = p < 0 ? 0 : p > n_in - 1 ? n_in - 1 : p;
      }
    }
    return 1;
  }

  double reduction = rq.out_true_size < (double)n_in ? ((double)n_in) / rq.out_true_size : 1.0;
  reduction *= blur;
  const double reach = f.radius * reduction;

  T.first.assign(n_out + 1, 0);
  size_t guess = (size_t)(2.0 * reach * n_out) + 16;
  T.src.reserve(guess);
  T.w.reserve(guess);
  for (int o = 0; o < n_out; o++) {
    T.first[o] = (int)T.w.size();
    double centre = (0.5 + (double)o - offset) / rq.out_true_size;
    double pos = centre * (double)n_in - 0.5;
    int lo = (int)ceil(pos - reach - 0.0001);
    int hi = (int)floor(pos + reach + 0.0001);
    double sum = 0.0;
    size_t start = T.w.size();
    for (int p = lo; p <= hi; p++) {
      bool outside = p < 0 || p > n_in - 1;
      if (outside && rq.edge_policy == IW_EDGE_POLICY_STANDARD) continue;
      double d = (((double)p) - pos) / reduction;
      double v = filter_value(f, d < 0 ? -d : d);
      if (v == 0.0) continue;  // exact zeros never enter the list
      sum += v;
      int s = p;
      if (outside) s = rq.edge_policy == IW_EDGE_POLICY_TRANSPARENT ? -1 : (p < 0 ? 0 : n_in - 1);
      T.src.push_back(s);
      T.w.push_back(v);
    }
    if (T.w.size() > start) {
      if (sum != 0.0)
        for (size_t t = start; t < T.w.size(); t++) T.w[t] /= sum;
      else
        for (size_t t = start; t < T.w.size(); t++) T.w[t] = 0.0;
    }
  }
  T.first[n_out] = (int)T.w.size();
  return 1;
}
