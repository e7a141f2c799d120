// api.cpp -- the libimagew C API surface (context, setters/getters, allocator and error conventions).
// Behaviour follows the reference entry points cited in include/imagew_b200.h; the implementation is new.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <strings.h>

#include "iwb_internal.h"

// IW_CHANNELTYPE_ALL fans out to every channel type, _NONALPHA to R,G,B,Gray.
template <class F>
static void fan_out(int channeltype, F set) {
  if (channeltype >= 0 && channeltype < IWB_NUM_CHANNELTYPES) set(channeltype);
  if (channeltype == IW_CHANNELTYPE_ALL) set(IW_CHANNELTYPE_ALPHA);
  if (channeltype == IW_CHANNELTYPE_ALL || channeltype == IW_CHANNELTYPE_NONALPHA) {
    set(IW_CHANNELTYPE_RED); set(IW_CHANNELTYPE_GREEN); set(IW_CHANNELTYPE_BLUE); set(IW_CHANNELTYPE_GRAY);
  }
}

extern "C" {

// ------------------------------------------------------------------ allocator (src/imagew-util.c:25-141)
static void *default_malloc(void *, unsigned int flags, size_t n) {
  return (flags & IW_MALLOCFLAG_ZEROMEM) ? calloc(n, 1) : malloc(n);
}
static void default_free(void *, void *mem) { free(mem); }

void *iw_malloc_ex(struct iw_context *ctx, unsigned int flags, size_t n) {
  void *mem = nullptr;
  if (n <= ctx->max_malloc) mem = ctx->mallocfn(ctx->userdata, flags, n);
  if (!mem && !(flags & IW_MALLOCFLAG_NOERRORS)) iw_set_error(ctx, "Out of memory");
  return mem;
}
void *iw_malloc(struct iw_context *ctx, size_t n) { return iw_malloc_ex(ctx, 0, n); }
void *iw_mallocz(struct iw_context *ctx, size_t n) { return iw_malloc_ex(ctx, IW_MALLOCFLAG_ZEROMEM, n); }
void *iw_malloc_large(struct iw_context *ctx, size_t n1, size_t n2) {
  if (n2 != 0 && n1 > ctx->max_malloc / n2) {
    iw_set_error(ctx, "Image too large to process");
    return nullptr;
  }
  // image-sized blocks come from the page-locked pool (pipeline.cu) when the default allocator is in use
  if (ctx->mallocfn == default_malloc && n1 * n2 >= ((size_t)1 << 20) && n1 * n2 <= ctx->max_malloc)
    if (void *p = iwb_pinned_alloc(n1 * n2)) return p;
  return iw_malloc_ex(ctx, 0, n1 * n2);
}
void *iw_realloc_ex(struct iw_context *ctx, unsigned int flags, void *m, size_t oldn, size_t n) {
  if (!m) return iw_malloc_ex(ctx, flags, n);
  void *mem = nullptr;
  if (n <= ctx->max_malloc) {
    // always a fresh block; the old one is released even on failure
    mem = ctx->mallocfn(ctx->userdata, flags, n);
    if (mem) memcpy(mem, m, oldn < n ? oldn : n);
    ctx->freefn(ctx->userdata, m);
  }
  if (!mem && !(flags & IW_MALLOCFLAG_NOERRORS)) iw_set_error(ctx, "Out of memory");
  return mem;
}
void *iw_realloc(struct iw_context *ctx, void *m, size_t oldn, size_t n) {
  return iw_realloc_ex(ctx, 0, m, oldn, n);
}
void iw_free(struct iw_context *ctx, void *mem) {
  if (!mem) return;
  if (ctx->mallocfn == default_malloc && iwb_pinned_free(mem)) return;
  ctx->freefn(ctx->userdata, mem);
}
char *iw_strdup(struct iw_context *ctx, const char *s) {
  if (!s) return nullptr;
  size_t n = strlen(s) + 1;
  char *d = (char *)iw_malloc(ctx, n);
  if (d) memcpy(d, s, n);
  return d;
}
void iw_strlcpy(char *dst, const char *src, size_t dstlen) {
  size_t n = strlen(src);
  if (n > dstlen - 1) n = dstlen - 1;
  memcpy(dst, src, n);
  dst[n] = 0;
}
void iw_snprintf(char *buf, size_t buflen, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, buflen, fmt, ap);
  va_end(ap);
  buf[buflen - 1] = 0;
}
int iw_stricmp(const char *a, const char *b) { return strcasecmp(a, b); }
void iw_zeromem(void *mem, size_t n) { memset(mem, 0, n); }

// ------------------------------------------------------------------ messages (src/imagew-api.c:16-162)
static void translate(struct iw_context *ctx, unsigned int flags, char *dst, size_t dstlen, const char *src) {
  dst[0] = 0;
  if (!(ctx && ctx->translate_fn && ctx->translate_fn(ctx, flags, dst, dstlen, src))) iw_strlcpy(dst, src, dstlen);
}
static void format_msg(struct iw_context *ctx, unsigned int flags, char *dst, size_t dstlen, const char *fmt,
                       va_list ap) {
  if (!ctx || !ctx->translate_fn) {
    vsnprintf(dst, dstlen, fmt, ap);
    dst[dstlen - 1] = 0;
    return;
  }
  char f2[IWB_MSG_MAX], out[IWB_MSG_MAX];
  translate(ctx, IW_TRANSLATEFLAG_FORMAT | flags, f2, sizeof f2, fmt);
  vsnprintf(out, sizeof out, f2, ap);
  out[sizeof out - 1] = 0;
  translate(ctx, IW_TRANSLATEFLAG_POSTFORMAT | flags, dst, dstlen, out);
}
static void record_error(struct iw_context *ctx, const char *s) {
  if (ctx->error_flag) return;  // first error sticks
  ctx->error_flag = 1;
  if (!ctx->error_msg) ctx->error_msg = (char *)iw_malloc_ex(ctx, IW_MALLOCFLAG_NOERRORS, IWB_MSG_MAX);
  if (ctx->error_msg) iw_strlcpy(ctx->error_msg, s, IWB_MSG_MAX);
}
void iw_set_error(struct iw_context *ctx, const char *s) {
  if (ctx->error_flag) return;
  char buf[IWB_MSG_MAX];
  translate(ctx, IW_TRANSLATEFLAG_ERRORMSG, buf, sizeof buf, s);
  record_error(ctx, buf);
}
void iw_set_errorf(struct iw_context *ctx, const char *fmt, ...) {
  if (ctx->error_flag) return;
  char buf[IWB_MSG_MAX];
  va_list ap;
  va_start(ap, fmt);
  format_msg(ctx, IW_TRANSLATEFLAG_ERRORMSG, buf, sizeof buf, fmt, ap);
  va_end(ap);
  record_error(ctx, buf);
}
void iw_warning(struct iw_context *ctx, const char *s) {
  if (!ctx->warning_fn) return;
  char buf[IWB_MSG_MAX];
  translate(ctx, IW_TRANSLATEFLAG_WARNINGMSG, buf, sizeof buf, s);
  ctx->warning_fn(ctx, buf);
}
void iw_warningf(struct iw_context *ctx, const char *fmt, ...) {
  if (!ctx->warning_fn) return;
  char buf[IWB_MSG_MAX];
  va_list ap;
  va_start(ap, fmt);
  format_msg(ctx, IW_TRANSLATEFLAG_WARNINGMSG, buf, sizeof buf, fmt, ap);
  va_end(ap);
  ctx->warning_fn(ctx, buf);
}
const char *iw_get_errormsg(struct iw_context *ctx, char *buf, int buflen) {
  if (ctx->error_msg)
    iw_strlcpy(buf, ctx->error_msg, buflen);
  else
    translate(ctx, IW_TRANSLATEFLAG_ERRORMSG, buf, buflen, "Error message not available");
  return buf;
}
int iw_get_errorflag(struct iw_context *ctx) { return ctx->error_flag; }
void iw_set_translate_fn(struct iw_context *ctx, iw_translatefn_type fn) { ctx->translate_fn = fn; }
void iw_set_warning_fn(struct iw_context *ctx, iw_warningfn_type fn) { ctx->warning_fn = fn; }
void iw_set_userdata(struct iw_context *ctx, void *u) { ctx->userdata = u; }
void *iw_get_userdata(struct iw_context *ctx) { return ctx->userdata; }

// ------------------------------------------------------------------ context (src/imagew-api.c:209-289)
struct iw_context *iw_create_context(struct iw_init_params *params) {
  iw_mallocfn_type mf = (params && params->mallocfn) ? params->mallocfn : default_malloc;
  iw_freefn_type ff = (params && params->mallocfn) ? params->freefn : default_free;
  void *mem = mf(params ? params->userdata : nullptr, IW_MALLOCFLAG_ZEROMEM, sizeof(iw_context));
  if (!mem) return nullptr;
  iw_context *ctx = new (mem) iw_context();
  ctx->mallocfn = mf;
  ctx->freefn = ff;
  if (params) {
    ctx->userdata = params->userdata;
    ctx->caller_api_version = params->api_version;
  }
  iw_make_srgb_csdescr_2(&ctx->img1cs);
  iw_make_srgb_csdescr_2(&ctx->img2cs);
  return ctx;
}

void iw_destroy_context(struct iw_context *ctx) {
  if (!ctx) return;
  if (ctx->img1.pixels) iw_free(ctx, ctx->img1.pixels);
  if (ctx->img2.pixels) iw_free(ctx, ctx->img2.pixels);
  if (ctx->opt_pixels) iw_free(ctx, ctx->opt_pixels);
  if (ctx->palette) iw_free(ctx, ctx->palette);
  if (ctx->error_msg) iw_free(ctx, ctx->error_msg);
  iw_freefn_type ff = ctx->freefn;
  void *ud = ctx->userdata;
  ctx->~iw_context();
  ff(ud, ctx);
}

// ------------------------------------------------------------------ small helpers (src/imagew-api.c:165-207)
int iw_imgtype_num_channels(int t) {
  return t == IW_IMGTYPE_RGBA ? 4 : t == IW_IMGTYPE_RGB ? 3 : t == IW_IMGTYPE_GRAYA ? 2 : 1;
}
size_t iw_calc_bytesperrow(int num_pixels, int bits_per_pixel) {
  return (size_t)(((num_pixels * bits_per_pixel) + 7) / 8);
}
int iw_check_image_dimensions(struct iw_context *ctx, int w, int h) {
  if (w > ctx->max_width || h > ctx->max_height) {
    iw_set_errorf(ctx, "Image dimensions too large (%d\xc3\x97%d)", w, h);
    return 0;
  }
  if (w < 1 || h < 1) {
    iw_set_errorf(ctx, "Invalid image dimensions (%d\xc3\x97%d)", w, h);
    return 0;
  }
  return 1;
}
int iw_is_valid_density(double dx, double dy, int code) {
  if (dx < 0.0001 || dy < 0.0001 || dx > 10000000.0 || dy > 10000000.0) return 0;
  if (dx / 10.0 > dy || dy / 10.0 > dx) return 0;
  return code == IW_DENSITY_UNITS_UNKNOWN || code == IW_DENSITY_UNITS_PER_METER;
}
int iw_get_sample_size(void) { return (int)sizeof(float); }
int iw_get_version_int(void) { return IW_VERSION_INT; }
char *iw_get_version_string(struct iw_context *, char *s, int s_len) {
  int v = IW_VERSION_INT;
  iw_snprintf(s, s_len, "%d.%d.%d", (v >> 16) & 0xff, (v >> 8) & 0xff, v & 0xff);
  return s;
}
unsigned int iw_color_get_int_sample(struct iw_color *clr, int channel, unsigned int maxcolorcode) {
  int n = (int)(0.5 + (clr->c[channel] * (double)maxcolorcode));
  if (n < 0) n = 0;
  if (n > (int)maxcolorcode) n = (int)maxcolorcode;
  return (unsigned int)n;
}

// Output-format capability table (src/imagew-util.c:556-624)
unsigned int iw_get_profile_by_fmt(int fmt) {
  switch (fmt) {
    case IW_FORMAT_PNG:
      return IW_PROFILE_TRANSPARENCY | IW_PROFILE_GRAYSCALE | IW_PROFILE_PALETTETRNS | IW_PROFILE_GRAY1 |
             IW_PROFILE_GRAY2 | IW_PROFILE_GRAY4 | IW_PROFILE_16BPS | IW_PROFILE_BINARYTRNS | IW_PROFILE_PAL1 |
             IW_PROFILE_PAL2 | IW_PROFILE_PAL4 | IW_PROFILE_PAL8 | IW_PROFILE_REDUCEDBITDEPTHS |
             IW_PROFILE_PNG_BKGD;
    case IW_FORMAT_BMP:
      return IW_PROFILE_PAL1 | IW_PROFILE_PAL4 | IW_PROFILE_PAL8 | IW_PROFILE_REDUCEDBITDEPTHS;
    case IW_FORMAT_JPEG:
      return IW_PROFILE_GRAYSCALE;
    case IW_FORMAT_TIFF:
      return IW_PROFILE_TRANSPARENCY | IW_PROFILE_GRAYSCALE | IW_PROFILE_GRAY1 | IW_PROFILE_GRAY4 |
             IW_PROFILE_16BPS | IW_PROFILE_PAL4 | IW_PROFILE_PAL8;
    case IW_FORMAT_MIFF:
      return IW_PROFILE_TRANSPARENCY | IW_PROFILE_GRAYSCALE | IW_PROFILE_ALWAYSLINEAR | IW_PROFILE_HDRI |
             IW_PROFILE_RGB16_BKGD;
    case IW_FORMAT_WEBP:
      return IW_PROFILE_TRANSPARENCY;
    case IW_FORMAT_PAM:
      return IW_PROFILE_16BPS | IW_PROFILE_REDUCEDBITDEPTHS | IW_PROFILE_GRAYSCALE | IW_PROFILE_GRAY1 |
             IW_PROFILE_TRANSPARENCY;
    case IW_FORMAT_PNM:
      return IW_PROFILE_16BPS | IW_PROFILE_REDUCEDBITDEPTHS | IW_PROFILE_GRAYSCALE | IW_PROFILE_GRAY1;
    case IW_FORMAT_PPM:
      return IW_PROFILE_16BPS | IW_PROFILE_REDUCEDBITDEPTHS;
    case IW_FORMAT_PGM:
      return IW_PROFILE_16BPS | IW_PROFILE_REDUCEDBITDEPTHS | IW_PROFILE_GRAYSCALE;
    case IW_FORMAT_PBM:
      return IW_PROFILE_GRAY1;
  }
  return 0;
}

// ------------------------------------------------------------------ setters (src/imagew-api.c:351-758)
void iw_set_output_canvas_size(struct iw_context *ctx, int w, int h) { ctx->canvas_w = w; ctx->canvas_h = h; }
void iw_set_output_image_size(struct iw_context *ctx, double w, double h) {
  ctx->out_true_w = w < 0.01 ? 0.01 : w;
  ctx->out_true_h = h < 0.01 ? 0.01 : h;
  ctx->out_true_valid = 1;
}
void iw_set_input_crop(struct iw_context *ctx, int x, int y, int w, int h) {
  ctx->crop_x = x; ctx->crop_y = y; ctx->crop_w = w; ctx->crop_h = h;
}
void iw_set_output_profile(struct iw_context *ctx, unsigned int n) { ctx->output_profile = n; }
void iw_set_output_depth(struct iw_context *ctx, int bps) { ctx->req_output_depth = bps; }
void iw_set_output_max_color_code(struct iw_context *ctx, int channeltype, int n) {
  if (channeltype >= 0 && channeltype < IWB_NUM_CHANNELTYPES) ctx->req_output_maxcode[channeltype] = n;
}
void iw_set_input_max_color_code(struct iw_context *ctx, int input_channel, int c) {
  if (input_channel >= 0 && input_channel < 4) ctx->in_maxcode_req[input_channel] = c;
}

void iw_set_dither_type(struct iw_context *ctx, int channeltype, int f, int s) {
  fan_out(channeltype, [&](int k) { ctx->dither_family[k] = f; ctx->dither_subtype[k] = s; });
}
void iw_set_color_count(struct iw_context *ctx, int channeltype, int c) {
  fan_out(channeltype, [&](int k) { ctx->req_color_count[k] = c; });
}
void iw_set_channel_offset(struct iw_context *ctx, int channeltype, int dimension, double offs) {
  if (channeltype < 0 || channeltype > 2) return;
  if (dimension < 0 || dimension > 1) dimension = 0;
  ctx->axis[dimension].channel_offset[channeltype] = offs;
}
void iw_set_resize_alg(struct iw_context *ctx, int dimension, int family, double blur, double p1, double p2) {
  if (dimension < 0 || dimension > 1) dimension = 0;
  iwb_axis_request &a = ctx->axis[dimension];
  a.family = family; a.blur = blur; a.param1 = p1; a.param2 = p2;
}
void iw_set_grayscale_weights(struct iw_context *ctx, double r, double g, double b) {
  double tot = r + g + b;
  if (tot == 0.0) tot = 1.0;
  ctx->gray_w[0] = r / tot; ctx->gray_w[1] = g / tot; ctx->gray_w[2] = b / tot;
}

static iw_color opaque_rgb(double r, double g, double b) {
  iw_color c; c.c[0] = r; c.c[1] = g; c.c[2] = b; c.c[3] = 1.0; return c;
}
void iw_set_input_bkgd_label_2(struct iw_context *ctx, const struct iw_color *clr) {
  ctx->in_label_set = 1; ctx->in_label_inputcs = *clr;
}
void iw_set_input_bkgd_label(struct iw_context *ctx, double r, double g, double b) {
  iw_color c = opaque_rgb(r, g, b); iw_set_input_bkgd_label_2(ctx, &c);
}
void iw_set_output_bkgd_label_2(struct iw_context *ctx, const struct iw_color *clr) {
  ctx->out_label_valid = 1; ctx->req_out_label = *clr;
}
void iw_set_output_bkgd_label(struct iw_context *ctx, double r, double g, double b) {
  iw_color c = opaque_rgb(r, g, b); iw_set_output_bkgd_label_2(ctx, &c);
}
void iw_set_apply_bkgd_2(struct iw_context *ctx, const struct iw_color *clr) {
  ctx->bkgd_valid = 1; ctx->req_bkgd = *clr;
}
void iw_set_apply_bkgd(struct iw_context *ctx, double r, double g, double b) {
  iw_color c = opaque_rgb(r, g, b); iw_set_apply_bkgd_2(ctx, &c);
}
void iw_set_bkgd_checkerboard_2(struct iw_context *ctx, int checkersize, const struct iw_color *clr) {
  ctx->req_checkerboard = 1; ctx->check_size = checkersize; ctx->req_bkgd2 = *clr;
}
void iw_set_bkgd_checkerboard(struct iw_context *ctx, int checkersize, double r2, double g2, double b2) {
  iw_color c = opaque_rgb(r2, g2, b2); iw_set_bkgd_checkerboard_2(ctx, checkersize, &c);
}
void iw_set_bkgd_checkerboard_origin(struct iw_context *ctx, int x, int y) {
  ctx->check_origin[IW_DIMENSION_H] = x; ctx->check_origin[IW_DIMENSION_V] = y;
}

// colour-space descriptors; a "gamma" of ~1.0 is really linear (src/imagew-api.c:511-567)
static void gamma_one_is_linear(struct iw_csdescr *cs) {
  if (cs->cstype == IW_CSTYPE_GAMMA && cs->gamma >= 0.999995 && cs->gamma <= 1.000005) cs->cstype = IW_CSTYPE_LINEAR;
}
void iw_make_linear_csdescr(struct iw_csdescr *cs) { cs->cstype = IW_CSTYPE_LINEAR; cs->gamma = 0.0; cs->srgb_intent = 0; }
void iw_make_srgb_csdescr(struct iw_csdescr *cs, int intent) { cs->cstype = IW_CSTYPE_SRGB; cs->gamma = 0.0; cs->srgb_intent = intent; }
void iw_make_srgb_csdescr_2(struct iw_csdescr *cs) { cs->cstype = IW_CSTYPE_SRGB; cs->gamma = 0.0; }
void iw_make_rec709_csdescr(struct iw_csdescr *cs) { cs->cstype = IW_CSTYPE_REC709; cs->gamma = 0.0; }
void iw_make_gamma_csdescr(struct iw_csdescr *cs, double gamma) {
  cs->cstype = IW_CSTYPE_GAMMA;
  cs->gamma = gamma < 0.1 ? 0.1 : gamma > 10.0 ? 10.0 : gamma;
  cs->srgb_intent = 0;
  gamma_one_is_linear(cs);
}
void iw_set_output_colorspace(struct iw_context *ctx, const struct iw_csdescr *cs) {
  ctx->req_out_cs = *cs; gamma_one_is_linear(&ctx->req_out_cs); ctx->out_cs_valid = 1;
}
void iw_set_input_colorspace(struct iw_context *ctx, const struct iw_csdescr *cs) {
  ctx->img1cs = *cs; gamma_one_is_linear(&ctx->img1cs);
}
double iw_convert_sample_to_linear(double v, const struct iw_csdescr *cs) { return iwb_to_linear(v, cs); }
double iw_convert_sample_from_linear(double v, const struct iw_csdescr *cs) { return iwb_from_linear(v, cs); }

int iw_get_input_density(struct iw_context *ctx, double *px, double *py, int *pcode) {
  *px = 1.0; *py = 1.0; *pcode = IW_DENSITY_UNKNOWN;
  if (ctx->img1.density_code == IW_DENSITY_UNKNOWN) return 0;
  if (!iw_is_valid_density(ctx->img1.density_x, ctx->img1.density_y, ctx->img1.density_code)) return 0;
  *px = ctx->img1.density_x; *py = ctx->img1.density_y; *pcode = ctx->img1.density_code;
  return 1;
}
void iw_set_output_density(struct iw_context *ctx, double x, double y, int code) {
  ctx->img2.density_code = code; ctx->img2.density_x = x; ctx->img2.density_y = y;
}
void iw_set_max_malloc(struct iw_context *ctx, size_t n) { ctx->max_malloc = n; }
void iw_set_random_seed(struct iw_context *ctx, int randomize, int seed) { ctx->randomize = randomize; ctx->random_seed = seed; }
void iw_set_allow_opt(struct iw_context *ctx, int opt, int n) {
  if (opt >= 1 && opt <= 5) ctx->opt_flags[opt] = n ? 1 : 0;
}
void iw_set_input_image(struct iw_context *ctx, const struct iw_image *img) { ctx->img1 = *img; }

// Rotate / mirror: only the transform code and the logical size change (src/imagew-api.c:660-684)
void iw_reorient_image(struct iw_context *ctx, unsigned int x) {
  static const unsigned int after_transpose[8] = {4, 6, 5, 7, 0, 2, 1, 3};
  x &= 7;
  if (x & 4) {
    ctx->img1.orient_transform = after_transpose[ctx->img1.orient_transform & 7];
    int t = ctx->img1.width; ctx->img1.width = ctx->img1.height; ctx->img1.height = t;
    double d = ctx->img1.density_x; ctx->img1.density_x = ctx->img1.density_y; ctx->img1.density_y = d;
  }
  ctx->img1.orient_transform ^= (x & 3);
}

void iw_get_output_image(struct iw_context *ctx, struct iw_image *img) {
  memset(img, 0, sizeof *img);
  img->width = ctx->res_w; img->height = ctx->res_h;
  img->imgtype = ctx->res_type; img->sampletype = ctx->img2.sampletype;
  img->bit_depth = ctx->res_depth;
  img->pixels = (iw_byte *)ctx->res_pixels; img->bpr = ctx->res_bpr;
  img->density_code = ctx->img2.density_code;
  img->density_x = ctx->img2.density_x; img->density_y = ctx->img2.density_y;
  img->rendering_intent = ctx->img2.rendering_intent;
  img->has_bkgdlabel = ctx->res_has_label;
  for (int k = 0; k < 4; k++)
    img->bkgdlabel.c[k] = ((double)ctx->res_label[k]) / (ctx->res_depth == 8 ? 255.0 : 65535.0);
  img->has_colorkey_trns = ctx->res_has_colorkey;
  for (int k = 0; k < 3; k++) img->colorkey[k] = ctx->res_colorkey[k];
  if (ctx->reduced_out_maxcode) {
    img->reduced_maxcolors = 1;
    bool alpha = IW_IMGTYPE_HAS_ALPHA(img->imgtype);
    if (IW_IMGTYPE_IS_GRAY(img->imgtype)) {
      img->maxcolorcode[IW_CHANNELTYPE_GRAY] = ctx->out_ci[0].maxcode;
      if (alpha) img->maxcolorcode[IW_CHANNELTYPE_ALPHA] = ctx->out_ci[1].maxcode;
    } else {
      for (int k = 0; k < 3; k++) img->maxcolorcode[k] = ctx->out_ci[k].maxcode;
      if (alpha) img->maxcolorcode[IW_CHANNELTYPE_ALPHA] = ctx->out_ci[3].maxcode;
    }
  }
}
void iw_get_output_colorspace(struct iw_context *ctx, struct iw_csdescr *cs) { *cs = ctx->img2cs; }
const struct iw_palette *iw_get_output_palette(struct iw_context *ctx) { return ctx->palette; }

// ------------------------------------------------------------------ key/value (src/imagew-api.c:760-994)
void iw_set_value(struct iw_context *ctx, int code, int n) {
  switch (code) {
    case IW_VAL_API_VERSION: ctx->caller_api_version = n; break;
    case IW_VAL_CVT_TO_GRAYSCALE: ctx->to_grayscale = n; break;
    case IW_VAL_DISABLE_GAMMA: ctx->no_gamma = n; break;
    case IW_VAL_NO_CSLABEL: ctx->suppress_cslabel = n; break;
    case IW_VAL_INT_CLAMP: ctx->intclamp = n; break;
    case IW_VAL_EDGE_POLICY_X: ctx->axis[IW_DIMENSION_H].edge_policy = n; break;
    case IW_VAL_EDGE_POLICY_Y: ctx->axis[IW_DIMENSION_V].edge_policy = n; break;
    case IW_VAL_PREF_UNITS: ctx->pref_units = n; break;
    case IW_VAL_GRAYSCALE_FORMULA: ctx->grayscale_formula = n; break;
    case IW_VAL_INPUT_NATIVE_GRAYSCALE: ctx->img1.native_grayscale = n; break;
    case IW_VAL_INCLUDE_SCREEN: ctx->include_screen = n; break;
    case IW_VAL_USE_BKGD_LABEL: ctx->use_label_from_file = n; break;
    case IW_VAL_MAX_WIDTH: ctx->max_width = n; break;
    case IW_VAL_MAX_HEIGHT: ctx->max_height = n; break;
    case IW_VAL_NO_BKGD_LABEL: ctx->suppress_out_label = n; break;
    case IW_VAL_INTENT: ctx->req_intent = n; break;
    case IW_VAL_OUTPUT_SAMPLE_TYPE: ctx->req_output_sample_type = n; break;
    case IW_VAL_NEGATE_TARGET: ctx->negate_target = n; break;
    default:
      if (code >= 0 && code < 64) ctx->misc_values[code] = n;  // codec-side settings: kept verbatim
  }
}
int iw_get_value(struct iw_context *ctx, int code) {
  switch (code) {
    case IW_VAL_API_VERSION: return ctx->caller_api_version;
    case IW_VAL_CVT_TO_GRAYSCALE: return ctx->to_grayscale;
    case IW_VAL_DISABLE_GAMMA: return ctx->no_gamma;
    case IW_VAL_NO_CSLABEL: return ctx->suppress_cslabel;
    case IW_VAL_INT_CLAMP: return ctx->intclamp;
    case IW_VAL_EDGE_POLICY_X: return ctx->axis[IW_DIMENSION_H].edge_policy;
    case IW_VAL_EDGE_POLICY_Y: return ctx->axis[IW_DIMENSION_V].edge_policy;
    case IW_VAL_PREF_UNITS: return ctx->pref_units;
    case IW_VAL_GRAYSCALE_FORMULA: return ctx->grayscale_formula;
    case IW_VAL_INPUT_NATIVE_GRAYSCALE: return ctx->img1.native_grayscale;
    case IW_VAL_INPUT_WIDTH: return ctx->img1.width < 1 ? 1 : ctx->img1.width;
    case IW_VAL_INPUT_HEIGHT: return ctx->img1.height < 1 ? 1 : ctx->img1.height;
    case IW_VAL_INPUT_IMAGE_TYPE: return ctx->img1.imgtype;
    case IW_VAL_INPUT_DEPTH: return ctx->img1.bit_depth;
    case IW_VAL_INCLUDE_SCREEN: return ctx->include_screen;
    case IW_VAL_OUTPUT_PALETTE_GRAYSCALE: return ctx->palette_is_gray;
    case IW_VAL_USE_BKGD_LABEL: return ctx->use_label_from_file;
    case IW_VAL_MAX_WIDTH: return ctx->max_width;
    case IW_VAL_MAX_HEIGHT: return ctx->max_height;
    case IW_VAL_PRECISION: return 32;
    case IW_VAL_NO_BKGD_LABEL: return ctx->suppress_out_label;
    case IW_VAL_INTENT: return ctx->req_intent;
    case IW_VAL_OUTPUT_SAMPLE_TYPE: return ctx->req_output_sample_type;
    case IW_VAL_NEGATE_TARGET: return ctx->negate_target;
  }
  return (code >= 0 && code < 64) ? ctx->misc_values[code] : 0;
}
void iw_set_value_dbl(struct iw_context *ctx, int code, double n) {
  if (code == IW_VAL_TRANSLATE_X) ctx->axis[IW_DIMENSION_H].translate = n;
  else if (code == IW_VAL_TRANSLATE_Y) ctx->axis[IW_DIMENSION_V].translate = n;
}
double iw_get_value_dbl(struct iw_context *ctx, int code) {
  if (code == IW_VAL_TRANSLATE_X) return ctx->axis[IW_DIMENSION_H].translate;
  if (code == IW_VAL_TRANSLATE_Y) return ctx->axis[IW_DIMENSION_V].translate;
  return 0.0;
}

// String options: empty value means "1"; same name replaces; at most 32 (src/imagew-api.c:996-1042)
void iw_set_option(struct iw_context *ctx, const char *name, const char *val) {
  if (!val || !val[0]) val = "1";
  for (auto &o : ctx->options)
    if (o.name == name) { o.val = val; return; }
  if (ctx->options.size() >= IWB_MAX_OPTIONS) return;
  ctx->options.push_back({name, val});
}
const char *iw_get_option(struct iw_context *ctx, const char *name) { return iwb_option(ctx, name); }

}  // extern "C"

int iwb_imgtype_alpha_index(int t) { return t == IW_IMGTYPE_RGBA ? 3 : t == IW_IMGTYPE_GRAYA ? 1 : 0; }

const char *iwb_option(const iw_context *ctx, const char *name) {
  for (auto &o : ctx->options)
    if (o.name == name) return o.val.c_str();
  return nullptr;
}
