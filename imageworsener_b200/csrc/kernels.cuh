// kernels.cuh -- device-side data structures and the per-sample arithmetic of the pixel pipeline.
//
// Parity rules that shape this file:
//   * every tap is  acc = acc + in*w  in list order, FP64, multiply and add rounded SEPARATELY (the reference
//     x86-64 build has no FMA): all parity-critical arithmetic goes through xmul/xadd/xsub/xdiv, which map to
//     __dmul_rn/__dadd_rn/... on the device so no flag or optimiser can contract them;
//   * the vertical pass runs first and its result is rounded to float; the resized alpha plane is rounded to
//     float before it un-premultiplies colour (reference: src/imagew-main.c:905-912, :1057-1061, :1093);
//   * no transcendental is evaluated on the device: colour conversions are host-built tables indexed by
//     integer codes, quantisation searches host-built tables of to_linear(code).
//
// Each kernel is written as a "body" functor run(params, idx) for one logical thread so that the same source
// can be executed by the test-only host emulation build (-DIWB_EMU, see tests/_emu) to debug host logic
// without a GPU.  The product library is compiled by nvcc for sm_100a only and contains no host path.
#pragma once
#include <cstddef>
#include <cstdint>

#include "prng.cuh"

#if defined(__CUDACC__) && !defined(IWB_EMU)
#define IWB_DEVICE_BUILD 1
#else
#define IWB_DEVICE_BUILD 0
#include <cmath>
#endif

// ---- rounding-exact primitives ------------------------------------------------------------------------
IWB_HD double xmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
IWB_HD double xadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
IWB_HD double xsub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
IWB_HD double xdiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}

// ---- plan structures (filled on the host, passed by value) ---------------------------------------------
struct DevAxis {          // one dimension's resampling recipe for one channel
  int mode;               // 0: tap list, 1: index map (nearest / null)
  int n_in, n_out;
  double edge_value;      // value of a virtual sample (src == -1)
  const int *first;       // [n_out+1]
  const int *src;
  const double *w;
  const int *map;         // [n_out]; -2 = zero
};

struct DevInput {
  const uint8_t *pix;     // points at physical row `row_origin`
  size_t bpr;
  int row_origin;
  int width, height;      // LOGICAL size (after the orientation transform)
  int orient;             // 0..7
  int x0, y0;             // crop origin, logical
  int depth;              // 1,2,4,8,16 (uint), 32 (float) or 64 (float input linearised to doubles on the host)
  int nch_phys;
  int is_float;
};

struct DevChanIn {        // how an intermediate channel's samples are produced from the input
  int in_ch;              // first physical channel; >= nch_phys means the virtual (opaque) alpha
  int to_gray;            // combine in_ch, in_ch+1, in_ch+2
  int gsf;                // IW_GSF_WEIGHTED / IW_GSF_ORDERBYVALUE
  double gw[3];
  const double *lut[3];   // code -> linear, per source channel (unused for float / virtual)
  int premult;            // multiply by raw alpha
  int early_bkgd;         // composite over bkgd before resizing
  double bkgd;
  int alpha_ch;           // physical alpha channel or >= nch_phys (virtual)
  const double *alpha_lut;  // code -> code/max
};

struct DevChanOut {       // what happens to a resampled intermediate channel
  int out_ch;             // -1: not stored
  int is_alpha;
  int premult;
  int late_bkgd;
  int bkgd_translucent;
  double bkgd1, bkgd2;
  int qmode;              // 0: linear arithmetic, 1: threshold tables
  int use_nc;             // nearest-colour midpoint table (no dither, 8-bit)
  const double *nc_tbl;   // [255]
  int ntab;               // entries in thr/codeval/linval
  const double *thr;      // linear-space lower bound of level q
  const double *codeval;  // stored code of level q; NULL: the code is q itself (every table that is not posterised)
  const double *linval;   // linear value of that code (the same array as thr when the code is q itself)
  double maxcode;
  int color_count;
  int guess_cs;           // colour space of the table levels (IW_CSTYPE_*), -1: no closed form (plain binary search)
  float guess_gamma;      // 1/gamma for IW_CSTYPE_GAMMA
  int dfamily, dsubtype;
  uint64_t prng_t1;       // PRNG state after the first step for this channel's seed
  int errdiff;            // this channel goes through the sequential error-diffusion kernel
};

struct DevOutput {
  uint8_t *pix;           // points at output row `row_origin`
  size_t bpr;
  int row_origin;
  int width, height;      // full output size
  int nch, depth;         // depth 8/16 (uint) or 32 (float)
  int is_float;
};

struct DevCommon {
  int checker;            // checkerboard background
  int check_size, check_ox, check_oy;
  double bkgd1alpha, bkgd2alpha;
  int intclamp;
};

// ---- raw sample access (reference: src/imagew-main.c:129-276) -------------------------------------------
// logical (x,y) within the crop -> physical (column,row)
IWB_HD void to_physical(const DevInput &in, int x, int y, int *px, int *py) {
  int lx = in.x0 + x, ly = in.y0 + y;
  switch (in.orient) {
    case 0: *px = lx; *py = ly; break;
    case 1: *px = in.width - 1 - lx; *py = ly; break;
    case 2: *px = lx; *py = in.height - 1 - ly; break;
    case 3: *px = in.width - 1 - lx; *py = in.height - 1 - ly; break;
    case 4: *px = ly; *py = lx; break;
    case 5: *px = ly; *py = in.width - 1 - lx; break;
    case 6: *px = in.height - 1 - ly; *py = lx; break;
    case 7: *px = in.height - 1 - ly; *py = in.width - 1 - lx; break;
    default: *px = 0; *py = 0; break;
  }
}

IWB_HD unsigned raw_code(const DevInput &in, int px, int py, int ch) {
  const uint8_t *row = in.pix + (size_t)(py - in.row_origin) * in.bpr;
  switch (in.depth) {
    case 8: return row[(size_t)in.nch_phys * px + ch];
    case 16: {
      size_t z = ((size_t)in.nch_phys * px + ch) * 2;
      return ((unsigned)row[z] << 8) | row[z + 1];   // big-endian
    }
    case 4: { unsigned b = row[px / 2]; return (px & 1) ? (b & 0x0f) : (b >> 4); }
    case 2: return (row[px / 4] >> ((3 - px % 4) * 2)) & 0x03;
    case 1: return (row[px / 8] >> (7 - px % 8)) & 1;
  }
  return 0;
}

IWB_HD double raw_float(const DevInput &in, int px, int py, int ch) {
  if (in.depth == 64)   // linearised on the host (pipeline.cu: float_prepass): aligned doubles
    return *reinterpret_cast<const double *>(in.pix + (size_t)(py - in.row_origin) * in.bpr + ((size_t)in.nch_phys * px + ch) * 8);
  const uint8_t *p = in.pix + (size_t)(py - in.row_origin) * in.bpr + ((size_t)in.nch_phys * px + ch) * 4;
  union { uint8_t b[4]; float f; } u;
  u.b[0] = p[0]; u.b[1] = p[1]; u.b[2] = p[2]; u.b[3] = p[3];
  return (double)u.f;
}

// one source channel -> linear
IWB_HD double one_linear(const DevInput &in, const DevChanIn &ci, int k, int px, int py) {
  int ch = ci.in_ch + k;
  if (ch >= in.nch_phys) return 1.0;                  // virtual alpha: opaque
  if (in.is_float) return raw_float(in, px, py, ch);  // only linear float input is accepted by the plan
  return ci.lut[k][raw_code(in, px, py, ch)];
}

// reference: iw_color_to_grayscale, src/imagew-main.c:278-305
IWB_HD double gray_of(const DevChanIn &ci, double r, double g, double b) {
  if (ci.gsf == 3) {  // order by value
    double v0, v1, v2;
    if (g <= r) { v0 = r; v1 = g; } else { v0 = g; v1 = r; }
    if (b <= v1) { v2 = b; }
    else { v2 = v1; if (b <= v0) { v1 = b; } else { v1 = v0; v0 = b; } }
    r = v0; g = v1; b = v2;
  } else if (ci.gsf != 2) {
    return 0.0;
  }
  return xadd(xadd(xmul(ci.gw[0], r), xmul(ci.gw[1], g)), xmul(ci.gw[2], b));
}

// Linear sample of intermediate channel `ci` at logical (x,y), with premultiply / early background
// (reference: src/imagew-main.c:348-383 and :876-894)
IWB_HD double fetch_linear(const DevInput &in, const DevChanIn &ci, int x, int y) {
  int px, py;
  to_physical(in, x, y, &px, &py);
  double v;
  if (ci.to_gray) {
    double r = one_linear(in, ci, 0, px, py), g = one_linear(in, ci, 1, px, py), b = one_linear(in, ci, 2, px, py);
    v = gray_of(ci, r, g, b);
  } else {
    v = one_linear(in, ci, 0, px, py);
  }
  if (ci.premult || ci.early_bkgd) {
    double a;
    if (ci.alpha_ch >= in.nch_phys) a = 1.0;
    else if (in.is_float) a = raw_float(in, px, py, ci.alpha_ch);
    else a = ci.alpha_lut[raw_code(in, px, py, ci.alpha_ch)];
    if (ci.premult) v = xmul(v, a);
    else v = xadd(xmul(a, v), xmul(xsub(1.0, a), ci.bkgd));
  }
  return v;
}

IWB_HD double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

// ---- vertical pass (reference: iw_process_cols_to_intermediate + iw_resize_row_std) ---------------------
struct VPassParams {
  DevInput in;
  DevCommon cm;
  int nch;                 // channels in this launch
  DevChanIn ch[4];
  DevAxis ax[4];
  float *inter[4];         // [rows][in_w] planes for output rows row0..row1
  int in_w;
  int row0, row1;          // output rows (vertical axis outputs) covered
};

struct VPassBody {
  static IWB_HD void run(const VPassParams &p, size_t idx) {
    const int rows = p.row1 - p.row0;
    const size_t per = (size_t)rows * p.in_w;
    const int c = (int)(idx / per);
    const size_t rem = idx - (size_t)c * per;
    const int jr = (int)(rem / p.in_w);
    const int i = (int)(rem - (size_t)jr * p.in_w);
    const int j = p.row0 + jr;
    const DevAxis &ax = p.ax[c];
    const DevChanIn &ci = p.ch[c];
    double acc = 0.0;
    if (ax.mode == 1) {
      int s = ax.map[j];
      acc = s >= 0 ? fetch_linear(p.in, ci, i, s) : 0.0;
    } else {
      const int t1 = ax.first[j + 1];
      for (int t = ax.first[j]; t < t1; t++) {
        int s = ax.src[t];
        double v = s >= 0 ? fetch_linear(p.in, ci, i, s) : ax.edge_value;
        acc = xadd(acc, xmul(v, ax.w[t]));
      }
    }
    if (p.cm.intclamp) acc = acc < 0.0 ? 0.0 : (acc > 1.0 ? 1.0 : acc);
    p.inter[c][(size_t)jr * p.in_w + i] = (float)acc;
  }
};

// ---- quantisation (reference: src/imagew-main.c:570-765) ------------------------------------------------
struct Candidates {
  int exact;
  double code_lo, code_hi;   // stored codes (as the reference keeps them: doubles)
  double lin_lo, lin_hi;     // their linear values
};

IWB_HD Candidates candidates_linear(const DevChanOut &o, double s) {
  Candidates c;
  double cvt = s > 0.999999999 ? 1.0 : s;
  if (o.color_count == 0) {
    double e = xmul(cvt, o.maxcode);
    if (e > o.maxcode) e = o.maxcode;
    if (e < 0.0) e = 0.0;
    c.code_lo = floor(e);
    c.code_hi = ceil(e);
  } else {
    double pm = (double)(o.color_count - 1);
    double e = xmul(cvt, pm);
    if (e > pm) e = pm;
    if (e < 0.0) e = 0.0;
    double step = xdiv(o.maxcode, pm);
    c.code_lo = floor(xadd(0.5000000001, xmul(floor(e), step)));
    c.code_hi = floor(xadd(0.5000000001, xmul(ceil(e), step)));
  }
  unsigned lo = (unsigned)c.code_lo, hi = (unsigned)c.code_hi;
  c.exact = lo == hi;
  c.lin_lo = xdiv((double)lo, o.maxcode);
  c.lin_hi = xdiv((double)hi, o.maxcode);
  return c;
}

// Non-linear target space: level q is the largest one whose linear lower bound thr[q] is <= s.
// (identity tables carry no code array: two gathers fewer per sample, and linval aliases thr, so the bounds the
// search has just read are the candidates' linear values)
IWB_HD double code_of_level(const DevChanOut &o, int q) { return o.codeval ? o.codeval[q] : (double)q; }

IWB_HD Candidates candidates_table(const DevChanOut &o, double s) {
  Candidates c;
  const int top = o.ntab - 1;
  if (s > 0.999999999) {
    c.exact = 1;
    c.code_lo = c.code_hi = code_of_level(o, top);
    c.lin_lo = c.lin_hi = o.linval[top];
    return c;
  }
  int lo = 0, hi = top;   // invariant: thr[lo] <= s (thr[0] == 0 <= s after clamping)
  if (o.guess_cs >= 0 && top >= 64) {
    // Full-depth tables (up to 65 536 levels): start from a single-precision estimate of the level -- the target
    // curve evaluated in FP32 -- and walk to the exact answer.  The estimate only chooses where the walk starts;
    // the result is decided by the same exact comparisons against the host-built table as the binary search.
    const float sf = (float)s;
    float v;
    if (o.guess_cs == 0) v = sf <= 0.0031308f ? 12.92f * sf : 1.055f * powf(sf, 1.0f / 2.4f) - 0.055f;
    else if (o.guess_cs == 3) v = sf < 0.020f ? 4.5f * sf : 1.099f * powf(sf, 0.45f) - 0.099f;
    else if (o.guess_cs == 2) v = powf(sf, o.guess_gamma);
    else v = sf;
    int g = (int)(v * (float)top);
    g = g < 0 ? 0 : (g > top ? top : g);
    int steps = 0;
    while (g < top && o.thr[g + 1] <= s && steps < 8) { g++; steps++; }
    while (g > 0 && o.thr[g] > s && steps < 16) { g--; steps++; }
    if (steps < 8 && o.thr[g] <= s && (g == top || o.thr[g + 1] > s)) {
      lo = g;
      hi = g;                                   // found: skip the search below
    }
  }
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (o.thr[mid] <= s) lo = mid; else hi = mid;
  }
  if (lo != top && o.thr[top] <= s) lo = top;
  if (lo >= top) {
    c.exact = 1;
    c.code_lo = c.code_hi = code_of_level(o, top);
    c.lin_lo = c.lin_hi = o.linval[top];
    return c;
  }
  c.exact = 0;
  c.code_lo = code_of_level(o, lo);
  c.code_hi = code_of_level(o, lo + 1);
  c.lin_lo = o.linval[lo];
  c.lin_hi = o.linval[lo + 1];
  return c;
}

IWB_HD double nearest_by_midpoints(const double *tbl, double s) {  // src/imagew-main.c:638-666
  unsigned x = 127, d = 64;
  for (;;) {
    if (x > 254 || tbl[x] > s) x -= d; else x += d;
    if (d == 1) return (x > 254 || tbl[x] > s) ? (double)x : (double)(x + 1);
    d >>= 1;
  }
}

// ordered-dither thresholds are (k+0.5)/64, exact in float (src/imagew-main.c:460-480)
IWB_HD double ordered_threshold(int subtype, int x, int y) {
  const unsigned char k0[64] = {0, 48, 12, 60, 3, 51, 15, 63, 32, 16, 44, 28, 35, 19, 47, 31,
                                8, 56, 4, 52, 11, 59, 7, 55, 40, 24, 36, 20, 43, 27, 39, 23,
                                2, 50, 14, 62, 1, 49, 13, 61, 34, 18, 46, 30, 33, 17, 45, 29,
                                10, 58, 6, 54, 9, 57, 5, 53, 42, 26, 38, 22, 41, 25, 37, 21};
  const unsigned char k1[64] = {3, 9, 17, 27, 25, 15, 7, 1, 11, 29, 37, 45, 43, 35, 23, 5,
                                19, 39, 51, 57, 55, 49, 33, 13, 31, 47, 59, 63, 61, 53, 41, 21,
                                30, 46, 58, 62, 60, 52, 40, 20, 18, 38, 50, 56, 54, 48, 32, 12,
                                10, 28, 36, 44, 42, 34, 22, 4, 2, 8, 16, 26, 24, 14, 6, 0};
  int i = (x % 8) + 8 * (y % 8);
  int k = subtype == 1 ? k1[i] : k0[i];
  return ((double)k + 0.5) / 64.0;
}

IWB_HD void store_code(const DevOutput &out, int x, int y, int ch, double code) {
  uint8_t *row = out.pix + (size_t)(y - out.row_origin) * out.bpr;
  if (out.depth == 8) {
    row[(size_t)out.nch * x + ch] = (uint8_t)(0.5 + code);
  } else {
    unsigned short v = (unsigned short)(0.5 + code);
    size_t z = ((size_t)out.nch * x + ch) * 2;
    row[z] = (uint8_t)(v >> 8);
    row[z + 1] = (uint8_t)(v & 0xff);
  }
}
IWB_HD void store_float(const DevOutput &out, int x, int y, int ch, double s) {
  uint8_t *p = out.pix + (size_t)(y - out.row_origin) * out.bpr + ((size_t)out.nch * x + ch) * 4;
  union { uint8_t b[4]; float f; } u;
  u.f = (float)s;
  p[0] = u.b[0]; p[1] = u.b[1]; p[2] = u.b[2]; p[3] = u.b[3];
}

// Compositing that happens after the horizontal pass (reference: src/imagew-main.c:1080-1115).  `a` is the resized
// alpha of the pixel, already rounded to FP32 as the reference's final_alpha32 plane holds it (:1057-1061); it is only
// read for channels that were premultiplied.
IWB_HD double composite_a(const DevCommon &cm, const DevChanOut &o, double s, double a, int x, int y) {
  int alt = 0;
  if (cm.checker)
    alt = (((cm.check_ox + x) / cm.check_size) % 2) != (((cm.check_oy + y) / cm.check_size) % 2);
  double ba = 0.0;
  if (o.bkgd_translucent) ba = alt ? cm.bkgd2alpha : cm.bkgd1alpha;
  if (o.premult) {
    if (a != 0.0) s = xdiv(s, a);
    if (o.late_bkgd) {
      double bk = alt ? o.bkgd2 : o.bkgd1;
      if (o.bkgd_translucent) s = xadd(xmul(s, a), xmul(xmul(bk, ba), xsub(1.0, a)));
      else s = xadd(xmul(s, a), xmul(bk, xsub(1.0, a)));
    }
  } else if (o.is_alpha && o.bkgd_translucent) {
    s = xadd(s, xmul(ba, xsub(1.0, s)));
  }
  return s;
}
IWB_HD double composite(const DevCommon &cm, const DevChanOut &o, double s, const float *final_alpha,
                        int out_w, int x, int y) {
  const double a = o.premult ? (double)final_alpha[(size_t)y * out_w + x] : 0.0;
  return composite_a(cm, o, s, a, x, y);
}

// Quantise + none/ordered/random dither (error diffusion has its own kernel).  `draw` is this sample's PRNG
// output (only read for random dither).  Returns the code to store.
IWB_HD double quantise(const DevChanOut &o, double s, int x, int y, uint32_t draw) {
  s = clamp01(s);
  if (o.use_nc) return nearest_by_midpoints(o.nc_tbl, s);
  Candidates c = o.qmode == 0 ? candidates_linear(o, s) : candidates_table(o, s);
  if (c.exact) return c.code_lo;
  double d_floor = xsub(s, c.lin_lo), d_ceil = xsub(c.lin_hi, s);
  if (o.dfamily == 0) return d_ceil <= d_floor ? c.code_hi : c.code_lo;
  if (o.dfamily == 1) {
    double frac = xdiv(d_floor, xadd(d_floor, d_ceil));
    return frac >= ordered_threshold(o.dsubtype, x, y) ? c.code_hi : c.code_lo;
  }
  if (o.dfamily == 3) {
    double frac = xdiv(d_floor, xadd(d_floor, d_ceil));
    double thr = xdiv((double)draw, (double)0xffffffffu);
    return frac >= thr ? c.code_hi : c.code_lo;
  }
  return 0.0;  // unsupported dither family: the reference stores 0 as well
}

// ---- horizontal pass + epilogue (reference: iw_process_rows_intermediate_to_final) ----------------------
struct HPassParams {
  DevCommon cm;
  DevOutput out;
  int nch;                  // channels in this launch
  DevChanOut ch[4];
  DevAxis ax[4];
  const float *inter[4];    // [rows][in_w]
  int in_w;
  int row0, row1;           // output rows of this band
  int inter_row0;           // output row that inter row 0 corresponds to
  float *final_alpha;       // [out_h_band][out_w]; written by the alpha channel, read by colour channels
  int alpha_row0;           // output row that final_alpha row 0 corresponds to
  double *lin_plane[4];     // errdiff channels: post-composite linear samples [rows][out_w]
  const uint32_t *draws[4]; // random-dither channels: one draw per sample [rows][out_w]
};

struct HPassBody {
  static IWB_HD void run(const HPassParams &p, size_t idx) {
    const int rows = p.row1 - p.row0;
    const int W = p.out.width;
    const size_t per = (size_t)rows * W;
    const int c = (int)(idx / per);
    const size_t rem = idx - (size_t)c * per;
    const int jr = (int)(rem / W);
    const int x = (int)(rem - (size_t)jr * W);
    const int y = p.row0 + jr;
    const DevAxis &ax = p.ax[c];
    const DevChanOut &o = p.ch[c];
    const float *row = p.inter[c] + (size_t)(y - p.inter_row0) * p.in_w;
    double acc = 0.0;
    if (ax.mode == 1) {
      int s = ax.map[x];
      acc = s >= 0 ? (double)row[s] : 0.0;
    } else {
      const int t1 = ax.first[x + 1];
      for (int t = ax.first[x]; t < t1; t++) {
        int s = ax.src[t];
        double v = s >= 0 ? (double)row[s] : ax.edge_value;
        acc = xadd(acc, xmul(v, ax.w[t]));
      }
    }
    if (p.cm.intclamp) acc = acc < 0.0 ? 0.0 : (acc > 1.0 ? 1.0 : acc);
    float *fa = p.final_alpha ? p.final_alpha - (size_t)p.alpha_row0 * W : nullptr;
    if (o.is_alpha && fa) fa[(size_t)y * W + x] = (float)acc;
    if (o.out_ch < 0) return;
    double s = composite(p.cm, o, acc, fa, W, x, y);
    if (p.out.is_float) { store_float(p.out, x, y, o.out_ch, s); return; }
    if (o.errdiff) { p.lin_plane[c][(size_t)jr * W + x] = s; return; }
    uint32_t draw = (o.dfamily == 3 && !o.use_nc) ? p.draws[c][(size_t)jr * W + x] : 0u;
    store_code(p.out, x, y, o.out_ch, quantise(o, s, x, y, draw));
  }
};

// ---- PRNG fill: one thread produces `seg` consecutive draws of one channel -------------------------------
struct PrngParams {
  int nch;
  uint64_t t1[4];
  uint32_t *draws[4];       // [rows][W]
  int W, row0, rows;
  int seg;
  int segs_per_row;
  const iwb_affine_t *jump; // [64]
};
struct PrngBody {
  static IWB_HD void run(const PrngParams &p, size_t idx) {
    const size_t per = (size_t)p.rows * p.segs_per_row;
    const int c = (int)(idx / per);
    const size_t rem = idx - (size_t)c * per;
    const int jr = (int)(rem / p.segs_per_row);
    const int sg = (int)(rem - (size_t)jr * p.segs_per_row);
    const int x0 = sg * p.seg;
    uint64_t n = (uint64_t)(p.row0 + jr) * (uint64_t)p.W + (uint64_t)x0;  // draw index of the first sample
    uint64_t t = p.t1[c];
    if (n) {
      if (t >= IWB_PRNG_P) t -= IWB_PRNG_P;
      t = iwb_prng_jump(t, n, p.jump);
    }
    uint32_t *dst = p.draws[c] + (size_t)jr * p.W;
    int x1 = x0 + p.seg < p.W ? x0 + p.seg : p.W;
    for (int x = x0; x < x1; x++) {
      dst[x] = iwb_prng_draw_of(t);
      t = iwb_prng_step(t);
    }
  }
};

// ---- error diffusion: one logical thread per channel, rows strictly in order, serpentine -----------------
struct ErrDiffParams {
  DevOutput out;
  int nch;
  DevChanOut ch[4];
  const double *lin_plane[4];  // [rows][W]
  double *err[4];              // 3*W doubles per channel, persistent across bands
  double m[4][12];             // diffusion weights per channel: x 0 1 / 2 3 4 5 6 / 7 8 9 10 11
  int row0, row1;
};

struct ErrDiffBody {
  static IWB_HD void run(const ErrDiffParams &p, size_t idx) {
    const int c = (int)idx;
    const DevChanOut &o = p.ch[c];
    const int W = p.out.width;
    double *e0 = p.err[c], *e1 = e0 + W, *e2 = e1 + W;
    const double *m = p.m[c];
    for (int y = p.row0; y < p.row1; y++) {
      const double *src = p.lin_plane[c] + (size_t)(y - p.row0) * W;
      const int fwd = (y % 2) ? -1 : 1;
      for (int z = 0; z < W; z++) {
        const int x = (y % 2) ? W - 1 - z : z;
        double s = clamp01(src[x]);
        s = xadd(s, e0[x]);
        if (s > 1.0) s = 1.0; else if (s < 0.0) s = 0.0;
        Candidates k = o.qmode == 0 ? candidates_linear(o, s) : candidates_table(o, s);
        double code;
        if (k.exact) {
          code = k.code_lo;
        } else {
          double d_floor = xsub(s, k.lin_lo), d_ceil = xsub(k.lin_hi, s);
          double err;
          if (d_ceil <= d_floor) { err = -d_ceil; code = k.code_hi; }
          else { err = d_floor; code = k.code_lo; }
          // distribute (src/imagew-main.c:545-566)
          if ((x - fwd) >= 0 && (x - fwd) < W) {
            if ((x - 2 * fwd) >= 0 && (x - 2 * fwd) < W) {
              e1[x - 2 * fwd] = xadd(e1[x - 2 * fwd], xmul(err, m[2]));
              e2[x - 2 * fwd] = xadd(e2[x - 2 * fwd], xmul(err, m[7]));
            }
            e1[x - fwd] = xadd(e1[x - fwd], xmul(err, m[3]));
            e2[x - fwd] = xadd(e2[x - fwd], xmul(err, m[8]));
          }
          e1[x] = xadd(e1[x], xmul(err, m[4]));
          e2[x] = xadd(e2[x], xmul(err, m[9]));
          if ((x + fwd) >= 0 && (x + fwd) < W) {
            e0[x + fwd] = xadd(e0[x + fwd], xmul(err, m[0]));
            e1[x + fwd] = xadd(e1[x + fwd], xmul(err, m[5]));
            e2[x + fwd] = xadd(e2[x + fwd], xmul(err, m[10]));
            if ((x + 2 * fwd) >= 0 && (x + 2 * fwd) < W) {
              e0[x + 2 * fwd] = xadd(e0[x + 2 * fwd], xmul(err, m[1]));
              e1[x + 2 * fwd] = xadd(e1[x + 2 * fwd], xmul(err, m[6]));
              e2[x + 2 * fwd] = xadd(e2[x + 2 * fwd], xmul(err, m[11]));
            }
          }
        }
        store_code(p.out, x, y, o.out_ch, code);
      }
      for (int x = 0; x < W; x++) { e0[x] = e1[x]; e1[x] = e2[x]; e2[x] = 0.0; }
    }
  }
};

// ---- whole-output touch-ups -------------------------------------------------------------------------------
struct PostParams {
  DevOutput out;
  int row0, row1;
  int negate;               // negate non-alpha channels (src/imagew-main.c:1324-1366)
  int maxcode[4];
  int alpha_ch;             // -1 if none
  int clear_transparent;    // zero colour of alpha==0 pixels (src/imagew-opt.c:1112-1163)
};
struct PostBody {
  static IWB_HD void run(const PostParams &p, size_t idx) {
    const int W = p.out.width;
    const int y = p.row0 + (int)(idx / W), x = (int)(idx % W);
    uint8_t *px = p.out.pix + (size_t)(y - p.out.row_origin) * p.out.bpr;
    const int nch = p.out.nch;
    if (p.negate) {
      for (int c = 0; c < nch; c++) {
        if (c == p.alpha_ch) continue;
        if (p.out.is_float) {
          union { uint8_t b[4]; float f; } u;
          uint8_t *q = px + ((size_t)nch * x + c) * 4;
          for (int k = 0; k < 4; k++) u.b[k] = q[k];
          u.f = ((float)1.0) - u.f;
          for (int k = 0; k < 4; k++) q[k] = u.b[k];
        } else if (p.out.depth == 8) {
          uint8_t *q = px + (size_t)nch * x + c;
          *q = (uint8_t)(p.maxcode[c] - *q);
        } else {
          uint8_t *q = px + ((size_t)nch * x + c) * 2;
          unsigned n = q[0] * 256u + q[1];
          n = (unsigned)p.maxcode[c] - n;
          q[0] = (uint8_t)((n & 0xff00) >> 8);
          q[1] = (uint8_t)(n & 0xff);
        }
      }
    }
    if (p.clear_transparent && p.alpha_ch >= 0) {
      if (p.out.depth == 8) {
        uint8_t *q = px + (size_t)nch * x;
        if (q[nch - 1] == 0) for (int c = 0; c < nch - 1; c++) q[c] = 0;
      } else if (p.out.depth == 16) {
        uint8_t *q = px + (size_t)nch * x * 2;
        if (q[(nch - 1) * 2] == 0 && q[(nch - 1) * 2 + 1] == 0)
          for (int c = 0; c < (nch - 1) * 2; c++) q[c] = 0;
      }
    }
  }
};
