/* iw_oracle.c -- TEST INFRASTRUCTURE ONLY (see iw_oracle.h).
 *
 * A scalar, one-sample-at-a-time restatement of the reference pixel pipeline in the reference's own order of
 * operations: per channel (alpha first), vertical pass with colour->linear fused in, FP32 intermediate,
 * horizontal pass, compositing, linear->target with pow(), quantise by linear-space distance, dither.
 * Unlike the CUDA product it evaluates pow() per sample exactly where the reference does, so it checks the
 * product's table-search quantiser independently.  Every function cites the reference lines it follows.
 * Compile with -ffp-contract=off (no FMA), like the reference's x86-64 build.
 */
#include "iw_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

enum { T_GRAY = 0x001, T_GRAYA = 0x101, T_RGB = 0x010, T_RGBA = 0x110 };
enum { CH_R = 0, CH_G = 1, CH_B = 2, CH_A = 3, CH_GRAY = 4 };
enum { CS_SRGB = 0, CS_LINEAR = 1, CS_GAMMA = 2, CS_REC709 = 3 };
enum { EDGE_REPLICATE = 1, EDGE_STANDARD = 2, EDGE_TRANSPARENT = 3 };
enum { F_AUTO = 0x01, F_NULL = 0x02, F_NEAREST = 0x10, F_MIX = 0x11, F_BOX = 0x20, F_TRIANGLE = 0x21,
       F_QUADRATIC = 0x22, F_GAUSSIAN = 0x23, F_HERMITE = 0x24, F_CUBIC = 0x25, F_BOXAVG = 0x26, F_SINC = 0x30,
       F_LANCZOS = 0x31, F_HANNING = 0x32, F_BLACKMAN = 0x33 };
enum { P_TRANSPARENCY = 0x1, P_16BPS = 0x40, P_ALWAYSLINEAR = 0x2000, P_HDRI = 0x4000, P_REDUCED = 0x8000,
       P_PNG_BKGD = 0x10000, P_RGB8_BKGD = 0x20000, P_RGB16_BKGD = 0x40000 };
enum { D_NONE = 0, D_ORDERED = 1, D_ERRDIFF = 2, D_RANDOM = 3 };
enum { BK_EARLY = 1, BK_LATE = 2 };

static int nchan(int t) { return t == T_RGBA ? 4 : t == T_RGB ? 3 : t == T_GRAYA ? 2 : 1; }
static int has_alpha(int t) { return (t & 0x100) != 0; }
static int alpha_index(int t) { return t == T_RGBA ? 3 : t == T_GRAYA ? 1 : 0; }
static int chtype(int t, int c) { /* reference: src/imagew-main.c:1463-1486 */
  if (t == T_GRAY) return c == 0 ? CH_GRAY : 0;
  if (t == T_GRAYA) return c == 0 ? CH_GRAY : c == 1 ? CH_A : 0;
  if (t == T_RGB) return (c >= 0 && c <= 2) ? c : 0;
  if (t == T_RGBA) return (c >= 0 && c <= 3) ? c : 0;
  return 0;
}

/* ---- colour spaces (reference: src/imagew-main.c:28-96, :427-446) ---- */
double iwo_to_linear(double v, int cs, double gamma) {
  if (cs == CS_LINEAR) return v;
  if (cs == CS_GAMMA) return pow(v, gamma);
  if (cs == CS_REC709) return v < 4.5 * 0.020 ? v / 4.5 : pow((v + 0.099) / 1.099, 1.0 / 0.45);
  return v <= 0.04045 ? v / 12.92 : pow((v + 0.055) / (1.055), 2.4);
}
double iwo_from_linear(double v, int cs, double gamma) {
  if (v > 0.999999999) return 1.0;
  if (cs == CS_LINEAR) return v;
  if (cs == CS_GAMMA) return pow(v, 1.0 / gamma);
  if (cs == CS_REC709) return v < 0.020 ? 4.5 * v : 1.099 * pow(v, 0.45) - 0.099;
  return v <= 0.0031308 ? 12.92 * v : 1.055 * pow(v, 1.0 / 2.4) - 0.055;
}

/* ---- PRNG (reference: src/imagew-util.c:199-233) ---- */
typedef struct { uint32_t mul, carry; } prng_t;
static void prng_seed(prng_t *g, int s) { g->mul = 0x03333333u + (uint32_t)s; g->carry = 0x05555555u + (uint32_t)s; }
static uint32_t prng_next(prng_t *g) {
  uint64_t x = (uint64_t)0xfff0bf23u * g->mul + g->carry;
  g->carry = (uint32_t)(x >> 32);
  g->mul = 0xffffffffu - (uint32_t)(x & 0xffffffffu);
  return g->mul;
}
uint32_t iwo_prng_nth(int seed, uint64_t n) {
  prng_t g;
  uint32_t v = 0;
  prng_seed(&g, seed);
  for (uint64_t i = 0; i <= n; i++) v = prng_next(&g);
  return v;
}

/* ---- resampling recipes (reference: src/imagew-resize.c) ---- */
typedef struct { int src, dst; double w; } tap_t;
typedef struct {
  int kind; /* 0 weight list, 1 nearest, 2 null */
  int n_in, n_out;
  double offset, edge_value;
  tap_t *taps; int ntaps, cap;
} rr_t;

typedef struct { int family; double radius, b, c, mix; } filt_t;

static double sinc(double x) { return x <= 0.000000005 ? 1.0 : sin(M_PI * x) / (M_PI * x); } /* resize.c:62-66 */

static double filt(const filt_t *f, double x) { /* resize.c:68-185 */
  switch (f->family) {
    case F_LANCZOS: return x < f->radius ? sinc(x) * sinc(x / f->radius) : 0.0;
    case F_HANNING: return x < f->radius ? sinc(x) * (0.5 * cos(M_PI * x / f->radius) + 0.5) : 0.0;
    case F_BLACKMAN:
      return x < f->radius ? sinc(x) * (0.5 * cos(M_PI * x / f->radius) + 0.08 * cos(2.0 * M_PI * x / f->radius) + 0.42) : 0.0;
    case F_SINC: return x < f->radius ? sinc(x) : 0.0;
    case F_GAUSSIAN: {
      double v;
      if (x >= 2.0) return 0.0;
      v = exp(-2.0 * x * x) * 0.79788456080286535587989;
      if (x <= 1.999) return v;
      return 1000.0 * (2.0 - x) * v;
    }
    case F_TRIANGLE: return x < 1.0 ? 1.0 - x : 0.0;
    case F_QUADRATIC:
      if (x < 0.5) return 0.75 - x * x;
      if (x < 1.5) return 0.50 * (x - 1.5) * (x - 1.5);
      return 0.0;
    case F_CUBIC: {
      double b = f->b, c = f->c;
      if (x < 1.0) return ((12.0 - 9.0 * b - 6.0 * c) * x * x * x + (-18.0 + 12.0 * b + 6.0 * c) * x * x + (6.0 - 2.0 * b)) / 6.0;
      if (x < 2.0) return ((-b - 6.0 * c) * x * x * x + (6.0 * b + 30.0 * c) * x * x + (-12.0 * b - 48.0 * c) * x + (8.0 * b + 24.0 * c)) / 6.0;
      return 0.0;
    }
    case F_HERMITE: return x < 1.0 ? 2.0 * x * x * x - 3.0 * x * x + 1.0 : 0.0;
    case F_BOX: return x <= 0.5 ? 1.0 : 0.0;
    case F_BOXAVG: return x < 0.4999999 ? 1.0 : (x <= 0.5000001 ? 0.5 : 0.0);
    case F_MIX:
      if (x < (0.5 - f->mix / 2.0)) return 1.0;
      if (x < (0.5 + f->mix / 2.0)) return 0.5 - (x - 0.5) / f->mix;
      return 0.0;
  }
  return 0.0;
}

static void rr_push(rr_t *r, int src, int dst, double w) { /* resize.c:216-227 */
  if (w == 0.0) return;
  if (r->ntaps >= r->cap) {
    r->cap = r->cap ? r->cap * 2 : 256;
    r->taps = (tap_t *)realloc(r->taps, sizeof(tap_t) * (size_t)r->cap);
  }
  r->taps[r->ntaps].src = src; r->taps[r->ntaps].dst = dst; r->taps[r->ntaps].w = w;
  r->ntaps++;
}

/* reference: iwpvt_resize_rows_init (resize.c:412-556) + iw_create_weightlist_std (resize.c:240-344) */
static rr_t *rr_create(const iwo_params *p, int d, int channeltype, int use_offset, int n_in, int n_out, double out_true,
                       double edge_value) {
  rr_t *r = (rr_t *)calloc(1, sizeof(rr_t));
  filt_t f;
  int std = 1, boxhack = 0, sincb = 0, o, i;
  double blur, reduction, reach;
  memset(&f, 0, sizeof f);
  f.family = p->family[d];
  r->n_in = n_in; r->n_out = n_out; r->edge_value = edge_value;
  switch (f.family) {
    case F_NULL: std = 0; r->kind = 2; break;
    case F_NEAREST: std = 0; r->kind = 1; boxhack = 1; break;
    case F_MIX:
      f.mix = ((double)n_out) / n_in;
      if (f.mix > 1.0) f.mix = 1.0 / f.mix;
      f.radius = 0.5 + f.mix;
      break;
    case F_BOX: f.radius = 0.501; boxhack = 1; break;
    case F_BOXAVG: f.radius = 0.501; break;
    case F_TRIANGLE: f.radius = 1.0; break;
    case F_QUADRATIC: f.radius = 1.5; break;
    case F_GAUSSIAN: f.radius = 2.0; break;
    case F_HERMITE: f.radius = 1.0; break;
    case F_CUBIC:
      f.radius = 2.0; f.b = p->param1[d]; f.c = p->param2[d];
      if (f.b < -100.0) f.b = -100.0;
      if (f.b > 100.0) f.b = 100.0;
      if (f.c < -100.0) f.c = -100.0;
      if (f.c > 100.0) f.c = 100.0;
      break;
    case F_LANCZOS: case F_HANNING: case F_BLACKMAN: case F_SINC: sincb = 1; break;
    default: free(r); return NULL;
  }
  if (sincb) {
    f.radius = floor(p->param1[d] + 0.5);
    if (f.radius < 2.0) f.radius = 2.0;
    if (f.radius > 100.0) f.radius = 100.0;
  }
  blur = p->blur[d];
  if (blur < 0.0001) blur = 0.0001;
  if (blur > 10000.0) blur = 10000.0;
  r->offset = p->translate[d];
  if (boxhack) r->offset -= 0.00000000001;
  if (use_offset && channeltype >= 0 && channeltype <= 2) r->offset += p->chan_offset[d][channeltype];
  if (!std) return r;

  reduction = out_true < (double)n_in ? ((double)n_in) / out_true : 1.0;
  reduction *= blur;
  reach = f.radius * reduction;
  for (o = 0; o < n_out; o++) {
    double centre = (0.5 + (double)o - r->offset) / out_true;
    double pos = centre * (double)n_in - 0.5;
    int first = (int)ceil(pos - reach - 0.0001), last = (int)floor(pos + reach + 0.0001);
    int start = r->ntaps, count = 0;
    double sum = 0.0;
    for (i = first; i <= last; i++) {
      double dd, v;
      int s;
      if (p->edge[d] == EDGE_STANDARD && (i < 0 || i >= n_in)) continue;
      dd = (((double)i) - pos) / reduction;
      v = filt(&f, dd < 0 ? -dd : dd);
      if (v == 0.0) continue;
      sum += v; count++;
      if (i < 0 || i > n_in - 1) s = p->edge[d] == EDGE_TRANSPARENT ? -1 : (i < 0 ? 0 : n_in - 1);
      else s = i;
      rr_push(r, s, o, v);
    }
    if (count > 0) {
      for (i = start; i < r->ntaps; i++) {
        if (sum != 0.0) r->taps[i].w /= sum; else r->taps[i].w = 0.0;
      }
    }
  }
  return r;
}
static void rr_free(rr_t *r) { if (r) { free(r->taps); free(r); } }

static void rr_run(const rr_t *r, const double *in, double *out) { /* resize.c:346-410 */
  int i;
  if (r->kind == 2) {
    for (i = 0; i < r->n_out; i++) out[i] = i < r->n_in ? in[i] : 0.0;
  } else if (r->kind == 1) {
    for (i = 0; i < r->n_out; i++) {
      double c = (0.5 + (double)i - r->offset) / (double)r->n_out;
      int s = (int)floor(c * (double)r->n_in);
      if (s < 0) s = 0; else if (s > r->n_in - 1) s = r->n_in - 1;
      out[i] = in[s];
    }
  } else {
    for (i = 0; i < r->n_out; i++) out[i] = 0.0;
    for (i = 0; i < r->ntaps; i++) {
      const tap_t *t = &r->taps[i];
      if (t->src >= 0) out[t->dst] += in[t->src] * t->w;
      else out[t->dst] += r->edge_value * t->w;
    }
  }
}

int iwo_weightlist(const iwo_params *p, int dimension, int channeltype, int n_in, int n_out, double out_true,
                   int use_offset, double edge_value, int *src, int *dst, double *w, int cap) {
  rr_t *r = rr_create(p, dimension, channeltype, use_offset, n_in, n_out, out_true, edge_value);
  int n, i;
  if (!r) return -1;
  n = r->ntaps;
  for (i = 0; i < n && i < cap; i++) { src[i] = r->taps[i].src; dst[i] = r->taps[i].dst; w[i] = r->taps[i].w; }
  rr_free(r);
  return n;
}

/* ---- state: the decisions of iw_prepare_processing (reference: src/imagew-main.c:1899-2248) ---- */
typedef struct { int type, in_ch, to_gray, out_ch, premult; double early_bkgd; } midch_t;
typedef struct { int type, dfam, dsub, cc, maxcode; double bk1, bk2; int use_nc; } outch_t;
typedef struct {
  const iwo_params *p;
  int in_logical, in_nphys, in_nlog, in_alpha;
  int in_max[4], in_slow[4], in_maxfull, reduced_in;
  int x0, y0, w, h; /* crop */
  int mid_type, mid_n, mid_alpha; midch_t mid[4];
  int out_type, out_n, out_depth, out_float; outch_t out[4]; size_t out_bpr;
  int cs_out, cs_in; double g_out, g_in;
  int use_offset[2], percharr[2];
  int apply_bkgd, strategy, checker, bk_source, check_size; double bk1a, bk2a;
  int gsf; double gw[3];
  int family[2];
  int reduced_out, no_out_tables, uses_errdiff;
  double *in_lut, *rev_lut, *nc_tbl;
  double label_lin[4];
  uint8_t *pix;
  float *inter, *inter_alpha, *final_alpha;
  int inter_w, inter_h;
  double *err[3];
  prng_t prng;
  rr_t *cache[2];
  iwo_params pp; /* working copy (auto filter is written back like the reference does) */
} st_t;

static double gray_of(const st_t *s, double r, double g, double b) { /* main.c:278-305 */
  double v0, v1, v2;
  if (s->gsf == 2) return s->gw[0] * r + s->gw[1] * g + s->gw[2] * b;
  if (s->gsf == 3) {
    if (g <= r) { v0 = r; v1 = g; } else { v0 = g; v1 = r; }
    if (b <= v1) { v2 = b; } else { v2 = v1; if (b <= v0) { v1 = b; } else { v1 = v0; v0 = b; } }
    return s->gw[0] * v0 + s->gw[1] * v1 + s->gw[2] * v2;
  }
  return 0.0;
}

static void phys(const st_t *s, int x, int y, int *rx, int *ry) { /* main.c:189-234 */
  const iwo_params *p = s->p;
  int lx = s->x0 + x, ly = s->y0 + y;
  switch (p->orient) {
    case 0: *rx = lx; *ry = ly; break;
    case 1: *rx = p->width - 1 - lx; *ry = ly; break;
    case 2: *rx = lx; *ry = p->height - 1 - ly; break;
    case 3: *rx = p->width - 1 - lx; *ry = p->height - 1 - ly; break;
    case 4: *rx = ly; *ry = lx; break;
    case 5: *rx = ly; *ry = p->width - 1 - lx; break;
    case 6: *rx = p->height - 1 - ly; *ry = lx; break;
    case 7: *rx = p->height - 1 - ly; *ry = p->width - 1 - lx; break;
    default: *rx = 0; *ry = 0;
  }
}

static unsigned raw_int(const st_t *s, int x, int y, int ch) { /* main.c:137-253 */
  const iwo_params *p = s->p;
  int rx, ry;
  const uint8_t *row;
  phys(s, x, y, &rx, &ry);
  row = p->pixels + (size_t)ry * p->bpr;
  switch (p->bit_depth) {
    case 8: return row[s->in_nphys * rx + ch];
    case 16: return ((unsigned)row[(s->in_nphys * rx + ch) * 2] << 8) | row[(s->in_nphys * rx + ch) * 2 + 1];
    case 4: return (rx & 1) ? (row[rx / 2] & 0x0f) : (row[rx / 2] >> 4);
    case 2: return (row[rx / 4] >> ((3 - rx % 4) * 2)) & 3;
    case 1: return (row[rx / 8] & (1 << (7 - rx % 8))) ? 1 : 0;
  }
  return 0;
}

static double raw_sample(const st_t *s, int x, int y, int ch) { /* main.c:257-276 */
  const iwo_params *p = s->p;
  if (ch >= s->in_nphys) return 1.0;
  if (p->sampletype == 1) {
    int rx, ry;
    float f;
    phys(s, x, y, &rx, &ry);
    if (p->bit_depth != 32) return 0.0;
    memcpy(&f, p->pixels + (size_t)ry * p->bpr + (size_t)(s->in_nphys * rx + ch) * 4, 4);
    return (double)f;
  }
  return ((double)raw_int(s, x, y, ch)) / (double)s->in_max[ch];
}

static double int_to_linear(const st_t *s, unsigned v, int cs, double g) { /* main.c:309-327 */
  if (cs == CS_LINEAR) return ((double)v) / (double)s->in_maxfull;
  if (s->in_lut) return s->in_lut[v];
  return iwo_to_linear(((double)v) / (double)s->in_maxfull, cs, g);
}

static double sample_linear(const st_t *s, int x, int y, int midc, int cs, double g) { /* main.c:348-383 */
  int ch = s->mid[midc].in_ch;
  if (s->in_slow[ch]) {
    if (s->mid[midc].to_gray) {
      double r = iwo_to_linear(raw_sample(s, x, y, ch), cs, g), gg = iwo_to_linear(raw_sample(s, x, y, ch + 1), cs, g),
             b = iwo_to_linear(raw_sample(s, x, y, ch + 2), cs, g);
      return gray_of(s, r, gg, b);
    }
    return iwo_to_linear(raw_sample(s, x, y, ch), cs, g);
  }
  if (s->mid[midc].to_gray) {
    double r = int_to_linear(s, raw_int(s, x, y, ch), cs, g), gg = int_to_linear(s, raw_int(s, x, y, ch + 1), cs, g),
           b = int_to_linear(s, raw_int(s, x, y, ch + 2), cs, g);
    return gray_of(s, r, gg, b);
  }
  return int_to_linear(s, raw_int(s, x, y, ch), cs, g);
}

static int bkgd_translucent(const st_t *s) { /* main.c:813-828 */
  const iwo_params *p = s->p;
  if (!s->apply_bkgd) return 0;
  if (!(p->profile & P_TRANSPARENCY)) return 0;
  if (s->strategy == BK_EARLY) return 0;
  if (s->bk_source == 1) { if (p->in_label[3] < 1.0) return 1; }
  else if (s->bk_source == 2) {
    if (s->checker) { if (p->bkgd2[3] < 1.0) return 1; }
    if (p->bkgd[3] < 1.0) return 1;
  }
  return 0;
}

/* ---- vertical pass (reference: iw_process_cols_to_intermediate, main.c:831-927) ---- */
static int vpass(st_t *s, int c, int cs, double g) {
  const iwo_params *p = &s->pp;
  midch_t *m = &s->mid[c];
  int is_alpha = m->type == CH_A, i, j;
  double *in = (double *)malloc(sizeof(double) * (size_t)s->h), *out = (double *)malloc(sizeof(double) * (size_t)s->inter_h);
  rr_t *r = s->cache[1];
  if (!r) {
    double ev = 0.0;
    if (p->edge[1] == EDGE_TRANSPARENT && s->apply_bkgd && s->strategy == BK_EARLY && m->type < 4) ev = s->mid[m->type].early_bkgd;
    r = rr_create(p, 1, m->type, s->use_offset[1], s->h, s->inter_h, p->out_true_valid ? p->out_true_h : (double)p->canvas_h, ev);
    if (!r) { free(in); free(out); return 0; }
    s->cache[1] = r;
  }
  for (i = 0; i < s->w; i++) {
    for (j = 0; j < s->h; j++) {
      in[j] = sample_linear(s, i, j, c, cs, g);
      if (m->premult) {
        double a = raw_sample(s, i, j, s->in_alpha);
        in[j] *= a;
      } else if (s->apply_bkgd && s->strategy == BK_EARLY) {
        double a = raw_sample(s, i, j, s->in_alpha);
        in[j] = (a) * (in[j]) + (1.0 - a) * (m->early_bkgd);
      }
    }
    rr_run(r, in, out);
    if (p->intclamp)
      for (j = 0; j < s->inter_h; j++) { if (out[j] < 0.0) out[j] = 0.0; else if (out[j] > 1.0) out[j] = 1.0; }
    for (j = 0; j < s->inter_h; j++) {
      if (is_alpha) s->inter_alpha[(size_t)j * s->inter_w + i] = (float)out[j];
      else s->inter[(size_t)j * s->inter_w + i] = (float)out[j];
    }
  }
  if (s->percharr[1]) { rr_free(r); s->cache[1] = NULL; }
  free(in); free(out);
  return 1;
}

/* ---- quantisation (reference: main.c:570-666) ---- */
static double lin_of_code(const st_t *s, unsigned v, int cs, double g, double maxcode) { /* main.c:330-344 */
  if (cs == CS_LINEAR) return ((double)v) / maxcode;
  if (s->rev_lut) return s->rev_lut[v];
  return iwo_to_linear(((double)v) / maxcode, cs, g);
}
static int candidates(const st_t *s, double lin, int cs, double g, double *lf, double *lc, double *cf, double *cc_,
                      double maxcode, int color_count) {
  double cvt = iwo_from_linear(lin, cs, g), e;
  unsigned fi, ci;
  if (color_count == 0) {
    e = cvt * maxcode;
    if (e > maxcode) e = maxcode;
    if (e < 0.0) e = 0.0;
    *cf = floor(e); *cc_ = ceil(e);
  } else {
    double pm = (double)(color_count - 1);
    e = cvt * pm;
    if (e > pm) e = pm;
    if (e < 0.0) e = 0.0;
    *cf = floor(0.5000000001 + floor(e) * (maxcode / pm));
    *cc_ = floor(0.5000000001 + ceil(e) * (maxcode / pm));
  }
  fi = (unsigned)(*cf); ci = (unsigned)(*cc_);
  if (fi == ci) return 1;
  *lf = lin_of_code(s, fi, cs, g, maxcode);
  *lc = lin_of_code(s, ci, cs, g, maxcode);
  return 0;
}
static double nc_search(const st_t *s, double lin) { /* main.c:638-666 */
  unsigned x = 127, d = 64;
  for (;;) {
    if (x > 254 || s->nc_tbl[x] > lin) x -= d; else x += d;
    if (d == 1) return (x > 254 || s->nc_tbl[x] > lin) ? (double)x : (double)(x + 1);
    d /= 2;
  }
}
static const float kOrdered[2][64] = { /* main.c:460-480 */
    {0.5 / 64, 48.5 / 64, 12.5 / 64, 60.5 / 64, 3.5 / 64, 51.5 / 64, 15.5 / 64, 63.5 / 64, 32.5 / 64, 16.5 / 64, 44.5 / 64,
     28.5 / 64, 35.5 / 64, 19.5 / 64, 47.5 / 64, 31.5 / 64, 8.5 / 64, 56.5 / 64, 4.5 / 64, 52.5 / 64, 11.5 / 64, 59.5 / 64,
     7.5 / 64, 55.5 / 64, 40.5 / 64, 24.5 / 64, 36.5 / 64, 20.5 / 64, 43.5 / 64, 27.5 / 64, 39.5 / 64, 23.5 / 64, 2.5 / 64,
     50.5 / 64, 14.5 / 64, 62.5 / 64, 1.5 / 64, 49.5 / 64, 13.5 / 64, 61.5 / 64, 34.5 / 64, 18.5 / 64, 46.5 / 64, 30.5 / 64,
     33.5 / 64, 17.5 / 64, 45.5 / 64, 29.5 / 64, 10.5 / 64, 58.5 / 64, 6.5 / 64, 54.5 / 64, 9.5 / 64, 57.5 / 64, 5.5 / 64,
     53.5 / 64, 42.5 / 64, 26.5 / 64, 38.5 / 64, 22.5 / 64, 41.5 / 64, 25.5 / 64, 37.5 / 64, 21.5 / 64},
    {3.5 / 64, 9.5 / 64, 17.5 / 64, 27.5 / 64, 25.5 / 64, 15.5 / 64, 7.5 / 64, 1.5 / 64, 11.5 / 64, 29.5 / 64, 37.5 / 64,
     45.5 / 64, 43.5 / 64, 35.5 / 64, 23.5 / 64, 5.5 / 64, 19.5 / 64, 39.5 / 64, 51.5 / 64, 57.5 / 64, 55.5 / 64, 49.5 / 64,
     33.5 / 64, 13.5 / 64, 31.5 / 64, 47.5 / 64, 59.5 / 64, 63.5 / 64, 61.5 / 64, 53.5 / 64, 41.5 / 64, 21.5 / 64, 30.5 / 64,
     46.5 / 64, 58.5 / 64, 62.5 / 64, 60.5 / 64, 52.5 / 64, 40.5 / 64, 20.5 / 64, 18.5 / 64, 38.5 / 64, 50.5 / 64, 56.5 / 64,
     54.5 / 64, 48.5 / 64, 32.5 / 64, 12.5 / 64, 10.5 / 64, 28.5 / 64, 36.5 / 64, 44.5 / 64, 42.5 / 64, 34.5 / 64, 22.5 / 64,
     4.5 / 64, 2.5 / 64, 8.5 / 64, 16.5 / 64, 26.5 / 64, 24.5 / 64, 14.5 / 64, 6.5 / 64, 0.5 / 64}};
static const double kDiffuse[8][12] = { /* main.c:510-535 */
    {7.0 / 16, 0.0, 0.0, 3.0 / 16, 5.0 / 16, 1.0 / 16, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {7.0 / 48, 5.0 / 48, 3.0 / 48, 5.0 / 48, 7.0 / 48, 5.0 / 48, 3.0 / 48, 1.0 / 48, 3.0 / 48, 5.0 / 48, 3.0 / 48, 1.0 / 48},
    {8.0 / 42, 4.0 / 42, 2.0 / 42, 4.0 / 42, 8.0 / 42, 4.0 / 42, 2.0 / 42, 1.0 / 42, 2.0 / 42, 4.0 / 42, 2.0 / 42, 1.0 / 42},
    {8.0 / 32, 4.0 / 32, 2.0 / 32, 4.0 / 32, 8.0 / 32, 4.0 / 32, 2.0 / 32, 0.0, 0.0, 0.0, 0.0, 0.0},
    {5.0 / 32, 3.0 / 32, 2.0 / 32, 4.0 / 32, 5.0 / 32, 4.0 / 32, 2.0 / 32, 0.0, 2.0 / 32, 3.0 / 32, 2.0 / 32, 0.0},
    {4.0 / 16, 3.0 / 16, 1.0 / 16, 2.0 / 16, 3.0 / 16, 2.0 / 16, 1.0 / 16, 0.0, 0.0, 0.0, 0.0, 0.0},
    {2.0 / 4, 0.0, 0.0, 1.0 / 4, 1.0 / 4, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {1.0 / 8, 1.0 / 8, 0.0, 1.0 / 8, 1.0 / 8, 1.0 / 8, 0.0, 0.0, 0.0, 1.0 / 8, 0.0, 0.0}};

static void diffuse(st_t *s, int sub, double err, int x, int y) { /* main.c:497-567 */
  const double *m = kDiffuse[(sub >= 0 && sub <= 7) ? sub : 0];
  int W = s->p->canvas_w, fwd = (y % 2) ? -1 : 1;
  if ((x - fwd) >= 0 && (x - fwd) < W) {
    if ((x - 2 * fwd) >= 0 && (x - 2 * fwd) < W) { s->err[1][x - 2 * fwd] += err * m[2]; s->err[2][x - 2 * fwd] += err * m[7]; }
    s->err[1][x - fwd] += err * m[3]; s->err[2][x - fwd] += err * m[8];
  }
  s->err[1][x] += err * m[4]; s->err[2][x] += err * m[9];
  if ((x + fwd) >= 0 && (x + fwd) < W) {
    s->err[0][x + fwd] += err * m[0]; s->err[1][x + fwd] += err * m[5]; s->err[2][x + fwd] += err * m[10];
    if ((x + 2 * fwd) >= 0 && (x + 2 * fwd) < W) {
      s->err[0][x + 2 * fwd] += err * m[1]; s->err[1][x + 2 * fwd] += err * m[6]; s->err[2][x + 2 * fwd] += err * m[11];
    }
  }
}

static void put_code(st_t *s, double code, int x, int y, int ch) { /* main.c:386-416 */
  if (s->out_depth == 8) s->pix[(size_t)y * s->out_bpr + (size_t)s->out_n * x + ch] = (uint8_t)(0.5 + code);
  else if (s->out_depth == 16) {
    unsigned short v = (unsigned short)(0.5 + code);
    size_t z = (size_t)y * s->out_bpr + ((size_t)s->out_n * x + ch) * 2;
    s->pix[z] = (uint8_t)(v >> 8); s->pix[z + 1] = (uint8_t)(v & 0xff);
  }
}

static void put_sample(st_t *s, double lin, int x, int y, int ch, int cs, double g) { /* main.c:669-765 */
  outch_t *o = &s->out[ch];
  double lf = 0, lc = 0, cf, cc, df, dc, full;
  if (lin < 0.0) lin = 0.0;
  if (lin > 1.0) lin = 1.0;
  if (o->use_nc) { put_code(s, nc_search(s, lin), x, y, ch); return; }
  if (o->dfam == D_ERRDIFF) {
    lin += s->err[0][x];
    if (lin > 1.0) lin = 1.0; else if (lin < 0.0) lin = 0.0;
  }
  if (candidates(s, lin, cs, g, &lf, &lc, &cf, &cc, (double)o->maxcode, o->cc)) {
    if (o->dfam == D_RANDOM) (void)prng_next(&s->prng);
    put_code(s, cf, x, y, ch);
    return;
  }
  df = lin - lf; dc = lc - lin;
  if (o->dfam == D_NONE) full = dc <= df ? cc : cf;
  else if (o->dfam == D_ERRDIFF) {
    if (dc <= df) { diffuse(s, o->dsub, -dc, x, y); full = cc; }
    else { diffuse(s, o->dsub, df, x, y); full = cf; }
  } else if (o->dfam == D_ORDERED) {
    double thr = kOrdered[o->dsub][(x % 8) + 8 * (y % 8)];
    full = (df / (df + dc)) >= thr ? cc : cf;
  } else if (o->dfam == D_RANDOM) {
    double thr = ((double)prng_next(&s->prng)) / (double)0xffffffff;
    full = (df / (df + dc)) >= thr ? cc : cf;
  } else full = 0.0;
  put_code(s, full, x, y, ch);
}

/* ---- horizontal pass + epilogue (reference: iw_process_rows_intermediate_to_final, main.c:929-1155) ---- */
static int hpass(st_t *s, int c, int cs, double g) {
  const iwo_params *p = &s->pp;
  midch_t *m = &s->mid[c];
  int oc = m->out_ch, is_alpha = m->type == CH_A, W = p->canvas_w, i, j, z, k;
  int translucent = bkgd_translucent(s), errdiff = 0;
  outch_t tmp, *o;
  double *in = (double *)malloc(sizeof(double) * (size_t)s->inter_w), *out = (double *)malloc(sizeof(double) * (size_t)W);
  rr_t *r;
  if (oc >= 0) o = &s->out[oc];
  else { memset(&tmp, 0, sizeof tmp); tmp.type = 11; o = &tmp; }
  o->use_nc = (s->nc_tbl && !is_alpha && o->dfam == D_NONE && o->cc == 0);
  if (o->dfam == D_RANDOM) {
    if (o->dsub == 1 && o->type != CH_A) prng_seed(&s->prng, p->random_seed);
    else prng_seed(&s->prng, p->random_seed + o->type);
  }
  if (oc >= 0 && o->dfam == D_ERRDIFF) {
    errdiff = 1;
    for (k = 0; k < 3; k++) for (i = 0; i < W; i++) s->err[k][i] = 0.0;
  }
  r = s->cache[0];
  if (!r) {
    double ev = 0.0;
    if (p->edge[0] == EDGE_TRANSPARENT && s->apply_bkgd && s->strategy == BK_EARLY && m->type < 4) ev = s->mid[m->type].early_bkgd;
    r = rr_create(p, 0, m->type, s->use_offset[0], s->inter_w, W, p->out_true_valid ? p->out_true_w : (double)p->canvas_w, ev);
    if (!r) { free(in); free(out); return 0; }
    s->cache[0] = r;
  }
  for (j = 0; j < s->inter_h; j++) {
    const float *row = (is_alpha ? s->inter_alpha : s->inter) + (size_t)j * s->inter_w;
    for (i = 0; i < s->inter_w; i++) in[i] = row[i];
    rr_run(r, in, out);
    if (p->intclamp)
      for (i = 0; i < W; i++) { if (out[i] < 0.0) out[i] = 0.0; else if (out[i] > 1.0) out[i] = 1.0; }
    if (is_alpha && s->final_alpha)
      for (i = 0; i < W; i++) s->final_alpha[(size_t)j * W + i] = (float)out[i];
    if (oc == -1) continue;
    for (z = 0; z < W; z++) {
      double v, a, ba = 0.0;
      int alt = 0;
      i = (errdiff && (j % 2)) ? W - 1 - z : z;
      v = out[i];
      if (s->checker) alt = (((p->check_ox + i) / s->check_size) % 2) != (((p->check_oy + j) / s->check_size) % 2);
      if (translucent) ba = alt ? s->bk2a : s->bk1a;
      if (m->premult) {
        a = s->final_alpha[(size_t)j * W + i];
        if (a != 0.0) v /= a;
        if (s->apply_bkgd && s->strategy == BK_LATE) {
          double bk = alt ? o->bk2 : o->bk1;
          if (translucent) v = v * a + bk * ba * (1.0 - a);
          else v = v * a + bk * (1.0 - a);
        }
      } else if (is_alpha && translucent) {
        v = v + ba * (1.0 - v);
      }
      if (s->out_float) {
        float f = (float)v;
        memcpy(s->pix + (size_t)j * s->out_bpr + ((size_t)s->out_n * i + oc) * 4, &f, 4);
      } else put_sample(s, v, i, j, oc, cs, g);
    }
    if (errdiff)
      for (i = 0; i < W; i++) { s->err[0][i] = s->err[1][i]; s->err[1][i] = s->err[2][i]; s->err[2][i] = 0.0; }
  }
  if (s->percharr[0]) { rr_free(r); s->cache[0] = NULL; }
  free(in); free(out);
  return 1;
}

/* ---- decisions (reference: main.c:1546-1611, :1640-1718, :1734-1823, :1842-1880, :1899-2248) ---- */
static int fail(char *err, int n, const char *m) { snprintf(err, (size_t)n, "%s", m); return 0; }

static int prepare(st_t *s, char *err, int errlen) {
  iwo_params *p = &s->pp;
  int i, d, in, mid, out;
  if (p->profile == 0) return fail(err, errlen, "Output profile not set");
  if (p->width > p->max_width || p->height > p->max_height || p->canvas_w > p->max_width || p->canvas_h > p->max_height)
    return fail(err, errlen, "Image dimensions too large");
  if (p->width < 1 || p->height < 1 || p->canvas_w < 1 || p->canvas_h < 1) return fail(err, errlen, "Invalid image dimensions");
  s->gsf = p->gsf; memcpy(s->gw, p->gray_w, sizeof s->gw);
  if (p->to_grayscale && (s->gsf == 0 || s->gsf == 1)) { /* main.c:1883-1895 */
    double r = s->gsf == 0 ? 0.212655 : 0.299, g = s->gsf == 0 ? 0.715158 : 0.587, b = s->gsf == 0 ? 0.072187 : 0.114, tot = r + g + b;
    s->gsf = 2; s->gw[0] = r / tot; s->gw[1] = g / tot; s->gw[2] = b / tot;
  }
  s->in_logical = p->imgtype; /* main.c:1842-1880 */
  if (p->edge[0] == EDGE_TRANSPARENT || p->edge[1] == EDGE_TRANSPARENT) {
    if (p->imgtype == T_GRAY) s->in_logical = T_GRAYA; else if (p->imgtype == T_RGB) s->in_logical = T_RGBA;
  }
  s->in_nphys = nchan(p->imgtype); s->in_nlog = nchan(s->in_logical); s->in_alpha = alpha_index(s->in_logical);
  for (i = 0; i < 4; i++) { s->in_max[i] = p->in_maxcode[i]; s->in_slow[i] = (i < s->in_nlog && i >= s->in_nphys); }
  for (i = 0; i < s->in_nlog; i++) { s->mid[i].type = chtype(s->in_logical, i); s->mid[i].in_ch = i; }
  s->x0 = p->crop_x; s->y0 = p->crop_y; s->w = p->crop_w; s->h = p->crop_h; /* main.c:1948-1958, quirk kept */
  if (s->x0 < 0) s->x0 = 0;
  if (s->y0 < 0) s->y0 = 0;
  if (s->x0 > p->width - 1) s->x0 = p->width - 1;
  if (s->y0 > p->height - 1) s->x0 = p->height - 1;
  if (s->w < 0) s->w = p->width - s->x0;
  if (s->h < 0) s->h = p->height - s->y0;
  if (s->w < 1) s->w = 1;
  if (s->h < 1) s->h = 1;
  if (s->w > p->width - s->x0) s->w = p->width - s->x0;
  if (s->h > p->height - s->y0) s->h = p->height - s->y0;
  if (p->cs_out_valid) { /* main.c:1960-1980 */
    s->cs_out = p->cs_out_type; s->g_out = p->cs_out_gamma;
    if ((p->profile & P_ALWAYSLINEAR) && s->cs_out != CS_LINEAR) { s->cs_out = CS_LINEAR; s->g_out = 0.0; }
  } else { s->cs_out = (p->profile & P_ALWAYSLINEAR) ? CS_LINEAR : CS_SRGB; s->g_out = 0.0; }
  s->cs_in = p->cs_in_type; s->g_in = p->cs_in_gamma;
  if (p->sampletype != 1) { /* main.c:1982-2007 */
    s->in_maxfull = (1 << p->bit_depth) - 1;
    for (i = 0; i < 4; i++) {
      if (s->in_max[i] <= 0) s->in_max[i] = s->in_maxfull;
      if (s->in_max[i] != s->in_maxfull) s->reduced_in = 1;
    }
  }
  if (s->reduced_in || p->sampletype == 1) for (i = 0; i < s->in_nphys; i++) s->in_slow[i] = 1;
  for (d = 0; d < 2; d++) for (i = 0; i < 3; i++) if (fabs(p->chan_offset[d][i]) > 0.00001) s->use_offset[d] = 1;
  if (p->to_grayscale && (s->use_offset[0] || s->use_offset[1])) s->use_offset[0] = s->use_offset[1] = 0;
  /* background strategy, main.c:1734-1823 */
  if (has_alpha(s->in_logical)) {
    if (p->in_label_set && (p->use_label_from_file || !p->bkgd_valid)) s->bk_source = 1;
    else if (p->bkgd_valid) { s->bk_source = 2; s->checker = p->checker_req; }
    s->check_size = p->check_size;
    if (s->checker && s->check_size < 1) s->check_size = 1;
    if (p->bkgd_valid) s->apply_bkgd = 1;
    if (!(p->profile & P_TRANSPARENCY)) s->apply_bkgd = 1;
    if (s->use_offset[0] || s->use_offset[1]) {
      s->apply_bkgd = 1;
      if (s->checker && p->checker_req) s->checker = 0;
      s->strategy = BK_EARLY;
    } else if (s->apply_bkgd) s->strategy = BK_LATE;
  }
  for (d = 0; d < 2; d++)
    if (s->use_offset[d] || (s->apply_bkgd && s->strategy == BK_EARLY && p->edge[d] == EDGE_TRANSPARENT)) s->percharr[d] = 1;
  /* strategy tables, main.c:1640-1718 and :2049-2126 (1=G 2=GA 3=RGB 4=RGBA) */
  in = s->in_logical == T_RGBA ? 4 : s->in_logical == T_RGB ? 3 : s->in_logical == T_GRAYA ? 2 : 1;
  mid = p->to_grayscale ? ((in == 4 || in == 2) ? 2 : 1) : ((in == 4 || in == 2) ? 4 : 3);
  out = mid;
  if (s->apply_bkgd && s->strategy == BK_EARLY) { if (mid == 4) mid = out = 3; else if (mid == 2) mid = out = 1; }
  if (s->apply_bkgd && !bkgd_translucent(s)) { if (mid == 2 && out == 2) out = 1; else if (mid == 4 && out == 4) out = 3; }
  if ((in == 2 || in == 4) && mid == 1) return fail(err, errlen, "Internal error, unknown strategy");
  s->mid_type = mid == 4 ? T_RGBA : mid == 3 ? T_RGB : mid == 2 ? T_GRAYA : T_GRAY;
  if (in == 2 && mid == 4) { s->mid[0].in_ch = s->mid[1].in_ch = s->mid[2].in_ch = 0; s->mid[3].in_ch = 1; }
  if ((in == 1 || in == 2) && mid == 3) s->mid[0].in_ch = s->mid[1].in_ch = s->mid[2].in_ch = 0;
  if (in == 4 && mid == 2) { s->mid[0].to_gray = 1; s->mid[0].in_ch = 0; s->mid[1].in_ch = 3; }
  if (in == 3 && mid == 1) { s->mid[0].to_gray = 1; s->mid[0].in_ch = 0; }
  if (in == 1 && mid == 1) s->mid[0].in_ch = 0;
  s->mid_n = nchan(s->mid_type); s->mid_alpha = alpha_index(s->mid_type);
  for (i = 0; i < s->mid_n; i++) s->mid[i].out_ch = i;
  s->out_type = out == 4 ? T_RGBA : out == 3 ? T_RGB : out == 2 ? T_GRAYA : T_GRAY;
  if (mid == 4 && out == 3) s->mid[3].out_ch = -1;
  if (mid == 2 && out == 1) s->mid[1].out_ch = -1;
  s->out_n = nchan(s->out_type);
  for (i = 0; i < s->mid_n; i++) s->mid[i].type = chtype(s->mid_type, i);
  for (i = 0; i < s->out_n; i++) s->out[i].type = chtype(s->out_type, i);
  if (has_alpha(s->mid_type)) for (i = 0; i < s->mid_n; i++) if (s->mid[i].type != CH_A) s->mid[i].premult = 1;
  /* output depth, main.c:1514-1541 */
  s->out_float = (p->profile & P_HDRI) != 0;
  if (s->out_float) s->out_depth = 32;
  else s->out_depth = (p->out_depth_req > 8 && (p->profile & P_16BPS)) ? 16 : 8;
  if (!s->out_float) { /* main.c:2154-2183 */
    int full = (1 << s->out_depth) - 1;
    for (i = 0; i < s->out_n; i++) s->out[i].maxcode = full;
    if (p->profile & P_REDUCED)
      for (i = 0; i < s->out_n; i++) {
        int mc = p->out_maxcode_req[s->out[i].type];
        if (mc > 0) s->out[i].maxcode = mc > full ? full : mc;
      }
    for (i = 0; i < s->out_n; i++) if (s->out[i].maxcode != full) { s->reduced_out = 1; s->no_out_tables = 1; }
  }
  for (i = 0; i < s->out_n; i++) { /* main.c:2185-2203 */
    outch_t *o = &s->out[i];
    o->cc = p->color_count_req[o->type];
    if (o->cc) { if (o->cc < 2) o->cc = 2; else if (o->cc > o->maxcode) o->cc = o->maxcode; }
    if (o->cc == 1 + o->maxcode) o->cc = 0;
    o->dfam = p->dither_family[o->type]; o->dsub = p->dither_subtype[o->type];
    if (o->dfam == D_ERRDIFF) s->uses_errdiff = 1;
  }
  if (!s->reduced_in && p->sampletype == 0 && s->cs_in != CS_LINEAR && p->bit_depth <= 8 &&
      ((size_t)p->width) * p->height > 512) { /* main.c:1172-1195, :2205-2207 */
    int n = 1 << p->bit_depth;
    s->in_lut = (double *)malloc(sizeof(double) * (size_t)n);
    for (i = 0; i < n; i++) s->in_lut[i] = iwo_to_linear(((double)i) / (n - 1), s->cs_in, s->g_in);
  }
  if (p->in_label_set) {
    for (i = 0; i < 3; i++) s->label_lin[i] = iwo_to_linear(p->in_label[i], s->cs_in, s->g_in);
    s->label_lin[3] = p->in_label[3];
  }
  if (s->apply_bkgd) { /* main.c:1546-1611 */
    double b1[4] = {1.0, 0.0, 1.0, 1.0}, b2[4] = {0.0, 0.0, 0.0, 1.0};
    if (s->bk_source == 1) { memcpy(b1, s->label_lin, sizeof b1); s->checker = 0; }
    else if (s->bk_source == 2) { memcpy(b1, p->bkgd, sizeof b1); if (p->checker_req) memcpy(b2, p->bkgd2, sizeof b2); }
    if (s->strategy == BK_EARLY) s->bk1a = 1.0; else { s->bk1a = b1[3]; s->bk2a = b2[3]; }
    if (s->strategy == BK_LATE && (s->out_type == T_RGB || s->out_type == T_RGBA)) {
      for (i = 0; i < 3; i++) { s->out[i].bk1 = b1[i]; if (s->checker) s->out[i].bk2 = b2[i]; }
    } else if (s->strategy == BK_LATE) {
      s->out[0].bk1 = gray_of(s, b1[0], b1[1], b1[2]);
      if (s->checker) s->out[0].bk2 = gray_of(s, b2[0], b2[1], b2[2]);
    } else if (s->strategy == BK_EARLY && s->out_type == T_RGB) {
      for (i = 0; i < 3; i++) s->mid[i].early_bkgd = b1[i];
    } else if (s->strategy == BK_EARLY && s->out_type == T_GRAY) {
      s->mid[0].early_bkgd = gray_of(s, b1[0], b1[1], b1[2]);
    }
  }
  for (d = 0; d < 2; d++) { /* main.c:1825-1840, :2229-2234 */
    int n_in = d == 0 ? s->w : s->h, n_out = d == 0 ? p->canvas_w : p->canvas_h;
    if (p->family[d] == F_AUTO) {
      if (n_out == n_in && !s->use_offset[d] && !p->out_true_valid && p->translate[d] == 0.0) { p->family[d] = F_NULL; p->blur[d] = 1.0; }
      else { p->family[d] = F_CUBIC; p->blur[d] = 1.0; p->param1[d] = 0.0; p->param2[d] = 0.5; }
    }
  }
  return 1;
}

static unsigned label_code(st_t *s, double lin, double maxcode) { /* main.c:769-799 */
  double lf = 0, lc = 0, cf, cc;
  if (lin < 0.0) lin = 0.0;
  if (lin > 1.0) lin = 1.0;
  if (candidates(s, lin, s->cs_out, s->g_out, &lf, &lc, &cf, &cc, maxcode, 0)) return (unsigned)(0.5 + cf);
  return (unsigned)(0.5 + ((lc - lin) <= (lin - lf) ? cc : cf));
}

void iwo_default_params(iwo_params *p) { /* api.c:209-266 */
  memset(p, 0, sizeof *p);
  p->family[0] = p->family[1] = F_AUTO;
  p->edge[0] = p->edge[1] = EDGE_STANDARD;
  p->blur[0] = p->blur[1] = 1.0;
  p->crop_w = p->crop_h = -1;
  p->cs_in_type = CS_SRGB;
  p->max_width = p->max_height = 40000;
}

void iwo_free_result(iwo_result *r) {
  if (!r) return;
  free(r->pixels); free(r->last_intermediate);
  r->pixels = NULL; r->last_intermediate = NULL;
}

int iwo_process(const iwo_params *params, iwo_result *res, char *err, int errlen) { /* main.c:1368-1461, :2250-2272 */
  st_t S, *s = &S;
  int c, k, ok = 0, W, H;
  iwo_params *p;
  memset(s, 0, sizeof S);
  memset(res, 0, sizeof *res);
  s->pp = *params; s->p = &s->pp; p = &s->pp;
  if (err && errlen) err[0] = 0;
  if (!prepare(s, err, errlen)) goto done;
  W = p->canvas_w; H = p->canvas_h;
  s->inter_w = s->w; s->inter_h = H;
  s->out_bpr = (size_t)(((W * s->out_depth * s->out_n) + 7) / 8);
  s->pix = (uint8_t *)calloc(s->out_bpr, (size_t)H);
  s->inter = (float *)malloc(sizeof(float) * (size_t)s->inter_w * s->inter_h);
  if (s->uses_errdiff) for (k = 0; k < 3; k++) s->err[k] = (double *)malloc(sizeof(double) * (size_t)W);
  if (!s->no_out_tables) { /* main.c:1404-1408, :1172-1236 */
    int ncol = 1 << s->out_depth;
    if (s->cs_out != CS_LINEAR && s->out_depth <= 8 && ((size_t)W) * H > 512) {
      s->rev_lut = (double *)malloc(sizeof(double) * (size_t)ncol);
      for (k = 0; k < ncol; k++) s->rev_lut[k] = iwo_to_linear(((double)k) / (ncol - 1), s->cs_out, s->g_out);
    }
    if (!p->no_gamma && s->cs_out != CS_LINEAR && !s->out_float && s->out_depth <= 8 && ((size_t)W) * H > 512) {
      double prev = 0.0;
      s->nc_tbl = (double *)malloc(sizeof(double) * (size_t)(ncol - 1));
      for (k = 0; k < ncol - 1; k++) {
        double curr = iwo_to_linear(((double)(k + 1)) / (ncol - 1), s->cs_out, s->g_out);
        s->nc_tbl[k] = (prev + curr) / 2.0;
        prev = curr;
      }
    }
  }
  if (has_alpha(s->mid_type)) {
    s->inter_alpha = (float *)malloc(sizeof(float) * (size_t)s->inter_w * s->inter_h);
    s->final_alpha = (float *)malloc(sizeof(float) * (size_t)W * H);
    if (!vpass(s, s->mid_alpha, CS_LINEAR, 0.0) || !hpass(s, s->mid_alpha, CS_LINEAR, 0.0)) { fail(err, errlen, "resize failed"); goto done; }
  }
  for (c = 0; c < s->mid_n; c++) {
    int r1, r2;
    if (s->mid[c].type == CH_A) continue;
    if (p->no_gamma) { r1 = vpass(s, c, CS_LINEAR, 0.0); r2 = r1 && hpass(s, c, CS_LINEAR, 0.0); }
    else { r1 = vpass(s, c, s->cs_in, s->g_in); r2 = r1 && hpass(s, c, s->cs_out, s->g_out); }
    if (!r2) { fail(err, errlen, "resize failed"); goto done; }
  }
  /* background label, main.c:1238-1322 */
  if ((p->profile & (P_PNG_BKGD | P_RGB8_BKGD | P_RGB16_BKGD)) && !p->suppress_out_label && (p->out_label_valid || p->in_label_set)) {
    double clr[4], mx = 0.0;
    memcpy(clr, p->out_label_valid ? p->out_label : s->label_lin, sizeof clr);
    if (p->to_grayscale) { double g = gray_of(s, clr[0], clr[1], clr[2]); clr[0] = clr[1] = clr[2] = g; }
    if (p->profile & P_RGB8_BKGD) mx = 255.0; else if (p->profile & P_RGB16_BKGD) mx = 65535.0;
    else if (s->out_depth == 8) mx = 255.0; else if (s->out_depth == 16) mx = 65535.0;
    if (mx > 0.0) {
      unsigned a;
      /* the reference evaluates the label through its general quantiser with whatever tables exist */
      for (k = 0; k < 3; k++) res->bkgdlabel[k] = ((double)label_code(s, clr[k], mx)) / mx;
      a = clr[3] <= 0.0 ? 0u : clr[3] >= 1.0 ? (unsigned)mx : (unsigned)(0.5 + clr[3] * (unsigned)mx);
      res->bkgdlabel[3] = ((double)a) / mx;
      res->has_bkgdlabel = 1;
    }
  }
  if (p->negate) { /* main.c:1324-1366 */
    int x, y;
    for (c = 0; c < s->out_n; c++) {
      if (s->out[c].type == CH_A) continue;
      for (y = 0; y < H; y++) for (x = 0; x < W; x++) {
        if (s->out_float) {
          float f; size_t z = (size_t)y * s->out_bpr + ((size_t)s->out_n * x + c) * 4;
          memcpy(&f, s->pix + z, 4); f = ((float)1.0) - f; memcpy(s->pix + z, &f, 4);
        } else if (s->out_depth == 8) {
          size_t z = (size_t)y * s->out_bpr + (size_t)s->out_n * x + c;
          s->pix[z] = (uint8_t)(s->out[c].maxcode - s->pix[z]);
        } else {
          size_t z = (size_t)y * s->out_bpr + ((size_t)s->out_n * x + c) * 2;
          unsigned n = s->pix[z] * 256u + s->pix[z + 1];
          n = (unsigned)s->out[c].maxcode - n;
          s->pix[z] = (uint8_t)((n & 0xff00) >> 8); s->pix[z + 1] = (uint8_t)(n & 0xff);
        }
      }
    }
  }
  /* the optimiser's always-on prologue: transparent pixels -> black (opt.c:1112-1163, :1192-1200) */
  if (!s->out_float && !s->reduced_out && has_alpha(s->out_type)) {
    int x, y, bps = s->out_depth / 8, n = s->out_n;
    for (y = 0; y < H; y++) for (x = 0; x < W; x++) {
      uint8_t *q = s->pix + (size_t)y * s->out_bpr + (size_t)x * n * bps;
      int clear = 1;
      for (k = 0; k < bps; k++) if (q[(n - 1) * bps + k]) clear = 0;
      if (clear) memset(q, 0, (size_t)(n - 1) * bps);
    }
  }
  res->imgtype = s->out_type; res->bit_depth = s->out_depth; res->sampletype = s->out_float;
  res->width = W; res->height = H; res->bpr = s->out_bpr; res->pixels = s->pix; s->pix = NULL;
  res->cs_out_type = s->cs_out; res->cs_out_gamma = s->g_out;
  res->reduced_maxcolors = s->reduced_out;
  if (s->reduced_out) {
    if (s->out_type == T_GRAY || s->out_type == T_GRAYA) {
      res->maxcolorcode[CH_GRAY] = (unsigned)s->out[0].maxcode;
      if (has_alpha(s->out_type)) res->maxcolorcode[CH_A] = (unsigned)s->out[1].maxcode;
    } else {
      for (k = 0; k < 3; k++) res->maxcolorcode[k] = (unsigned)s->out[k].maxcode;
      if (has_alpha(s->out_type)) res->maxcolorcode[CH_A] = (unsigned)s->out[3].maxcode;
    }
  }
  res->last_intermediate = s->inter; s->inter = NULL; res->inter_w = s->inter_w; res->inter_h = s->inter_h;
  ok = 1;
done:
  free(s->pix); free(s->inter); free(s->inter_alpha); free(s->final_alpha);
  for (k = 0; k < 3; k++) free(s->err[k]);
  free(s->in_lut); free(s->rev_lut); free(s->nc_tbl);
  rr_free(s->cache[0]); rr_free(s->cache[1]);
  return ok;
}

/* ---- synthetic input + fingerprints for bench.py's reference arm and the tests: xorshift64
 * noise, byte = state >> 24, raster order; FNV-1a-64 of a buffer.  Kept here so that the reference arm never has to
 * load the product library. ---- */
void iwo_synth_xorshift(uint8_t *buf, size_t n, uint64_t seed) {
  uint64_t s = seed;
  for (size_t i = 0; i < n; i++) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    buf[i] = (uint8_t)(s >> 24);
  }
}
uint64_t iwo_fnv1a64(const uint8_t *buf, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; i++) {
    h ^= buf[i];
    h *= 0x100000001b3ull;
  }
  return h;
}
