/* imagew_b200.h -- C ABI of libimagew_b200.so
 *
 * A from-scratch, B200-native implementation of ImageWorsener's pixel pipeline behind the libimagew C API.
 * Every entry point below replaces the reference function of the same name and keeps its argument meaning
 * and error convention (1 = success, 0 = failure, first error sticks on the context); the citation after each
 * group is the reference declaration/definition it stands in for.  Struct layouts and numeric constants are
 * ABI and therefore identical to the reference's (src/imagew.h:40-410).
 *
 * What is new relative to the reference (section "B200 extensions" at the end):
 *   - the string option "backend" (default and only legal value: "cuda"); there is no CPU path in this
 *     library -- iw_process_image() fails with an error if no CUDA device is usable;
 *   - "cuda:device" (ordinal), "cuda:band" ("index/count": process one output-row band only);
 *   - tuning / verification options, none of which may change a result byte (tests/test_stream2.py):
 *     "cuda:gen" (1: first-generation streaming kernels), "cuda:uniform" (0: general path only),
 *     "cuda:fused" (0: generic per-sample kernels only), "cuda:strip2" / "cuda:seg2" (rows / pixels per CTA part of
 *     the generation-2 kernels), "cuda:strip" / "cuda:seg" / "cuda:nc" (generation 1), "cuda:chunk_mb" (size of the
 *     FP32 intermediate), "cuda:pipeline" (sub-bands of the host-buffer pipeline), "cuda:overlap" (two-stream V/H
 *     overlap, off), "cuda:keep_pitch" (1: the device copy of a host image keeps the caller's row stride and whole rows;
 *     default: a 16-byte aligned device pitch, and with a crop in x only the cropped columns are copied, so that
 *     the TMA kernels apply whatever the stride and crop origin are);
 *   - iwb200_job_*: the plan/launch layer under iw_process_image() for callers that keep images resident
 *     in HBM (benchmarks, multi-GPU band sharding).
 */
#ifndef IMAGEW_B200_H
#define IMAGEW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(IWB_BUILDING) && defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this header is exported */
#endif

#define IW_VERSION_INT 0x010305 /* API level implemented (reference: src/imagew.h:40) */

/* ---- iw_set_value / iw_get_value codes (reference: src/imagew.h:51-149) ---- */
enum {
  IW_VAL_API_VERSION = 1, IW_VAL_PRECISION = 10, IW_VAL_CVT_TO_GRAYSCALE = 11, IW_VAL_COMPRESSION = 12,
  IW_VAL_DISABLE_GAMMA = 13, IW_VAL_NO_CSLABEL = 14, IW_VAL_INT_CLAMP = 15, IW_VAL_GRAYSCALE_FORMULA = 16,
  IW_VAL_EDGE_POLICY_X = 17, IW_VAL_EDGE_POLICY_Y = 18, IW_VAL_INPUT_NATIVE_GRAYSCALE = 20,
  IW_VAL_INPUT_WIDTH = 21, IW_VAL_INPUT_HEIGHT = 22, IW_VAL_INPUT_IMAGE_TYPE = 23, IW_VAL_INPUT_DEPTH = 24,
  IW_VAL_PREF_UNITS = 26, IW_VAL_OUTPUT_PALETTE_GRAYSCALE = 35, IW_VAL_OUTPUT_INTERLACED = 36,
  IW_VAL_USE_BKGD_LABEL = 37, IW_VAL_PAGE_TO_READ = 39, IW_VAL_INCLUDE_SCREEN = 40,
  IW_VAL_TRANSLATE_X = 42, IW_VAL_TRANSLATE_Y = 43, IW_VAL_BMP_NO_FILEHEADER = 44, IW_VAL_MAX_WIDTH = 45,
  IW_VAL_MAX_HEIGHT = 46, IW_VAL_NO_BKGD_LABEL = 48, IW_VAL_INTENT = 49, IW_VAL_OUTPUT_SAMPLE_TYPE = 50,
  IW_VAL_OUTPUT_FORMAT = 52, IW_VAL_NEGATE_TARGET = 53
};

/* ---- file formats, only as keys for iw_get_profile_by_fmt (reference: src/imagew.h:152-168) ---- */
enum {
  IW_FORMAT_UNKNOWN = 0, IW_FORMAT_PNG = 1, IW_FORMAT_JPEG = 2, IW_FORMAT_BMP = 3, IW_FORMAT_TIFF = 4,
  IW_FORMAT_MIFF = 5, IW_FORMAT_WEBP = 6, IW_FORMAT_GIF = 7, IW_FORMAT_PNM = 8, IW_FORMAT_PBM = 9,
  IW_FORMAT_PGM = 10, IW_FORMAT_PPM = 11, IW_FORMAT_PAM = 12
};

/* ---- output-format capability bits (reference: src/imagew.h:174-205) ---- */
#define IW_PROFILE_TRANSPARENCY 0x0001
#define IW_PROFILE_GRAYSCALE 0x0002
#define IW_PROFILE_PALETTETRNS 0x0004
#define IW_PROFILE_GRAY1 0x0008
#define IW_PROFILE_GRAY2 0x0010
#define IW_PROFILE_GRAY4 0x0020
#define IW_PROFILE_16BPS 0x0040
#define IW_PROFILE_ALWAYSSRGB 0x0080
#define IW_PROFILE_BINARYTRNS 0x0100
#define IW_PROFILE_PAL1 0x0200
#define IW_PROFILE_PAL2 0x0400
#define IW_PROFILE_PAL4 0x0800
#define IW_PROFILE_PAL8 0x1000
#define IW_PROFILE_ALWAYSLINEAR 0x2000
#define IW_PROFILE_HDRI 0x4000
#define IW_PROFILE_REDUCEDBITDEPTHS 0x8000
#define IW_PROFILE_PNG_BKGD 0x10000
#define IW_PROFILE_RGB8_BKGD 0x20000
#define IW_PROFILE_RGB16_BKGD 0x40000

/* ---- resampling families (reference: src/imagew.h:207-221) ---- */
enum {
  IW_RESIZETYPE_AUTO = 0x01, IW_RESIZETYPE_NULL = 0x02, IW_RESIZETYPE_NEAREST = 0x10, IW_RESIZETYPE_MIX = 0x11,
  IW_RESIZETYPE_BOX = 0x20, IW_RESIZETYPE_TRIANGLE = 0x21, IW_RESIZETYPE_QUADRATIC = 0x22,
  IW_RESIZETYPE_GAUSSIAN = 0x23, IW_RESIZETYPE_HERMITE = 0x24, IW_RESIZETYPE_CUBIC = 0x25,
  IW_RESIZETYPE_BOXAVG = 0x26, IW_RESIZETYPE_SINC = 0x30, IW_RESIZETYPE_LANCZOS = 0x31,
  IW_RESIZETYPE_HANNING = 0x32, IW_RESIZETYPE_BLACKMAN = 0x33
};

/* ---- image / sample / channel types (reference: src/imagew.h:223-254) ---- */
#define IW_IMGTYPE_GRAY 0x0001
#define IW_IMGTYPE_GRAYA 0x0101
#define IW_IMGTYPE_RGB 0x0010
#define IW_IMGTYPE_RGBA 0x0110
#define IW_IMGTYPE_PALETTE 0x1000
#define IW_IMGTYPE_IS_GRAY(x) (((x)&0x001) ? 1 : 0)
#define IW_IMGTYPE_HAS_ALPHA(x) (((x)&0x100) ? 1 : 0)
enum { IW_SAMPLETYPE_UINT = 0, IW_SAMPLETYPE_FLOATINGPOINT = 1 };
enum {
  IW_CHANNELTYPE_RED = 0, IW_CHANNELTYPE_GREEN = 1, IW_CHANNELTYPE_BLUE = 2, IW_CHANNELTYPE_ALPHA = 3,
  IW_CHANNELTYPE_GRAY = 4, IW_CHANNELTYPE_ALL = 10, IW_CHANNELTYPE_NONALPHA = 11
};

/* ---- colour spaces, intents, dimensions (reference: src/imagew.h:257-277) ---- */
enum { IW_CSTYPE_SRGB = 0, IW_CSTYPE_LINEAR = 1, IW_CSTYPE_GAMMA = 2, IW_CSTYPE_REC709 = 3 };
enum { IW_INTENT_UNKNOWN = 0, IW_INTENT_DEFAULT = 1, IW_INTENT_NONE = 2, IW_INTENT_PERCEPTUAL = 10,
       IW_INTENT_RELATIVE = 11, IW_INTENT_SATURATION = 12, IW_INTENT_ABSOLUTE = 13 };
enum { IW_DIMENSION_H = 0, IW_DIMENSION_V = 1 };

/* ---- dithering (reference: src/imagew.h:280-294) ---- */
enum { IW_DITHERFAMILY_NONE = 0, IW_DITHERFAMILY_ORDERED = 1, IW_DITHERFAMILY_ERRDIFF = 2,
       IW_DITHERFAMILY_RANDOM = 3 };
enum { IW_DITHERSUBTYPE_DEFAULT = 0, IW_DITHERSUBTYPE_HALFTONE = 1, IW_DITHERSUBTYPE_SAMEPATTERN = 1 };
enum { IW_DITHERSUBTYPE_FS = 0, IW_DITHERSUBTYPE_JJN = 1, IW_DITHERSUBTYPE_STUCKI = 2,
       IW_DITHERSUBTYPE_BURKES = 3, IW_DITHERSUBTYPE_SIERRA3 = 4, IW_DITHERSUBTYPE_SIERRA2 = 5,
       IW_DITHERSUBTYPE_SIERRA42A = 6, IW_DITHERSUBTYPE_ATKINSON = 7 };

/* ---- misc (reference: src/imagew.h:297-343) ---- */
enum { IW_DENSITY_UNKNOWN = 0, IW_DENSITY_UNITS_UNKNOWN = 1, IW_DENSITY_UNITS_PER_METER = 2 };
enum { IW_GSF_STANDARD = 0, IW_GSF_COMPATIBLE = 1, IW_GSF_WEIGHTED = 2, IW_GSF_ORDERBYVALUE = 3 };
enum { IW_EDGE_POLICY_REPLICATE = 1, IW_EDGE_POLICY_STANDARD = 2, IW_EDGE_POLICY_TRANSPARENT = 3 };
enum { IW_REORIENT_NOCHANGE = 0, IW_REORIENT_FLIP_H = 1, IW_REORIENT_FLIP_V = 2, IW_REORIENT_ROTATE_180 = 3,
       IW_REORIENT_TRANSPOSE = 4, IW_REORIENT_ROTATE_90 = 5, IW_REORIENT_ROTATE_270 = 6,
       IW_REORIENT_TRANSVERSE = 7 };
enum { IW_OPT_GRAYSCALE = 1, IW_OPT_PALETTE = 2, IW_OPT_16_TO_8 = 3, IW_OPT_STRIP_ALPHA = 4,
       IW_OPT_BINARY_TRNS = 5 };
#define IW_TRANSLATEFLAG_FORMAT 0x0001
#define IW_TRANSLATEFLAG_POSTFORMAT 0x0002
#define IW_TRANSLATEFLAG_ERRORMSG 0x0010
#define IW_TRANSLATEFLAG_WARNINGMSG 0x0020
#define IW_MALLOCFLAG_ZEROMEM 0x01
#define IW_MALLOCFLAG_NOERRORS 0x02

typedef uint8_t iw_byte;

/* ---- ABI structs (reference: src/imagew.h:364-419, :482-491) ---- */
struct iw_csdescr {
  int cstype;      /* IW_CSTYPE_* */
  int srgb_intent; /* deprecated, carried for layout */
  double gamma;    /* only for IW_CSTYPE_GAMMA */
};
struct iw_color {
  double c[4]; /* R,G,B,A in 0..1; colour space depends on the call */
};
struct iw_image {
  int imgtype, bit_depth, sampletype;
  int width, height; /* logical size (after orient_transform) */
  iw_byte *pixels;   /* interleaved; 16-bit samples big-endian; float32 native */
  size_t bpr;
  unsigned int orient_transform; /* input only: bit0 mirror-x, bit1 mirror-y, bit2 transpose */
  int native_grayscale;
  int density_code;
  double density_x, density_y;
  int has_colorkey_trns;
  unsigned int colorkey[3];
  int reduced_maxcolors;
  unsigned int maxcolorcode[5];
  int has_bkgdlabel;
  struct iw_color bkgdlabel;
  int rendering_intent;
};
struct iw_rgba8color { iw_byte r, g, b, a; };
struct iw_palette { int num_entries; struct iw_rgba8color entry[256]; };

struct iw_context; /* opaque */

typedef void *(*iw_mallocfn_type)(void *userdata, unsigned int flags, size_t n);
typedef void (*iw_freefn_type)(void *userdata, void *mem);
typedef void (*iw_warningfn_type)(struct iw_context *ctx, const char *msg);
typedef int (*iw_translatefn_type)(struct iw_context *ctx, unsigned int flags, char *dst, size_t dstlen,
                                   const char *src);
struct iw_init_params {
  int api_version;
  void *userdata;
  iw_mallocfn_type mallocfn; /* optional; both or neither */
  iw_freefn_type freefn;
};

/* ---- context life cycle and the hot-path entry (reference: src/imagew.h:497-501; src/imagew-api.c:221-289;
 *      src/imagew-main.c:2250-2272) ---- */
struct iw_context *iw_create_context(struct iw_init_params *params);
void iw_destroy_context(struct iw_context *ctx);
int iw_process_image(struct iw_context *ctx); /* single use per context */

/* ---- errors, warnings, hooks (reference: src/imagew.h:519-536, :658-663; src/imagew-api.c:16-162) ---- */
void iw_set_translate_fn(struct iw_context *ctx, iw_translatefn_type fn);
int iw_get_errorflag(struct iw_context *ctx);
const char *iw_get_errormsg(struct iw_context *ctx, char *buf, int buflen);
void iw_set_userdata(struct iw_context *ctx, void *userdata);
void *iw_get_userdata(struct iw_context *ctx);
void iw_set_warning_fn(struct iw_context *ctx, iw_warningfn_type fn);
void iw_set_error(struct iw_context *ctx, const char *s);
void iw_set_errorf(struct iw_context *ctx, const char *fmt, ...);
void iw_warning(struct iw_context *ctx, const char *s);
void iw_warningf(struct iw_context *ctx, const char *fmt, ...);

/* ---- geometry, profile, depth (reference: src/imagew.h:507, :539-563; src/imagew-api.c:351-389, :660-684) ---- */
void iw_reorient_image(struct iw_context *ctx, unsigned int x);
void iw_set_max_malloc(struct iw_context *ctx, size_t n);
void iw_set_output_canvas_size(struct iw_context *ctx, int w, int h);
void iw_set_output_image_size(struct iw_context *ctx, double w, double h);
void iw_set_input_crop(struct iw_context *ctx, int x, int y, int w, int h);
void iw_set_output_profile(struct iw_context *ctx, unsigned int n);
void iw_set_output_depth(struct iw_context *ctx, int bps);
void iw_set_output_max_color_code(struct iw_context *ctx, int channeltype, int n);
void iw_set_input_max_color_code(struct iw_context *ctx, int input_channel, int c);

/* ---- dither, posterise, resampling, offsets, grayscale (reference: src/imagew.h:565-583;
 *      src/imagew-api.c:391-440, :646-658, :735-748) ---- */
void iw_set_dither_type(struct iw_context *ctx, int channeltype, int family, int subtype);
void iw_set_color_count(struct iw_context *ctx, int channeltype, int c);
void iw_set_resize_alg(struct iw_context *ctx, int dimension, int family, double blur, double param1,
                       double param2);
void iw_set_channel_offset(struct iw_context *ctx, int channeltype, int dimension, double offs);
void iw_set_grayscale_weights(struct iw_context *ctx, double r, double g, double b);

/* ---- background and labels (reference: src/imagew.h:586-603; src/imagew-api.c:449-479, :569-608) ---- */
void iw_set_input_bkgd_label(struct iw_context *ctx, double r, double g, double b);
void iw_set_input_bkgd_label_2(struct iw_context *ctx, const struct iw_color *clr);
void iw_set_output_bkgd_label(struct iw_context *ctx, double r, double g, double b);
void iw_set_output_bkgd_label_2(struct iw_context *ctx, const struct iw_color *clr);
void iw_set_apply_bkgd(struct iw_context *ctx, double r, double g, double b);
void iw_set_apply_bkgd_2(struct iw_context *ctx, const struct iw_color *clr);
void iw_set_bkgd_checkerboard(struct iw_context *ctx, int checkersize, double r2, double g2, double b2);
void iw_set_bkgd_checkerboard_2(struct iw_context *ctx, int checkersize, const struct iw_color *clr);
void iw_set_bkgd_checkerboard_origin(struct iw_context *ctx, int x, int y);

/* ---- colour spaces (reference: src/imagew.h:607-621, :669-670; src/imagew-api.c:511-567;
 *      src/imagew-main.c:53-72, :427-452) ---- */
void iw_set_output_colorspace(struct iw_context *ctx, const struct iw_csdescr *csdescr);
void iw_set_input_colorspace(struct iw_context *ctx, const struct iw_csdescr *csdescr);
void iw_make_linear_csdescr(struct iw_csdescr *cs);
void iw_make_srgb_csdescr(struct iw_csdescr *cs, int srgb_intent);
void iw_make_srgb_csdescr_2(struct iw_csdescr *cs);
void iw_make_rec709_csdescr(struct iw_csdescr *cs);
void iw_make_gamma_csdescr(struct iw_csdescr *cs, double gamma);
double iw_convert_sample_to_linear(double v, const struct iw_csdescr *csdescr);
double iw_convert_sample_from_linear(double v, const struct iw_csdescr *csdescr);

/* ---- density, seed, optimiser switches (reference: src/imagew.h:623-632; src/imagew-api.c:481-508, :615-619,
 *      :721-733) ---- */
int iw_get_input_density(struct iw_context *ctx, double *px, double *py, int *pcode);
void iw_set_output_density(struct iw_context *ctx, double x, double y, int code);
void iw_set_random_seed(struct iw_context *ctx, int randomize, int rand_seed);
void iw_set_allow_opt(struct iw_context *ctx, int opt, int n);

/* ---- images in and out (reference: src/imagew.h:634-647; src/imagew-api.c:291-349, :641-644).
 *      iw_set_input_image: struct is copied, the context takes ownership of img->pixels (must come from
 *      iw_malloc_large).  iw_get_output_image: pixels stay owned by the context. ---- */
void iw_set_input_image(struct iw_context *ctx, const struct iw_image *img);
void iw_get_output_image(struct iw_context *ctx, struct iw_image *img);
void iw_get_output_colorspace(struct iw_context *ctx, struct iw_csdescr *csdescr);
const struct iw_palette *iw_get_output_palette(struct iw_context *ctx);

/* ---- key/value settings and the string-option dictionary (reference: src/imagew.h:649-656;
 *      src/imagew-api.c:760-1042) ---- */
void iw_set_value(struct iw_context *ctx, int code, int n);
int iw_get_value(struct iw_context *ctx, int code);
void iw_set_value_dbl(struct iw_context *ctx, int code, double n);
double iw_get_value_dbl(struct iw_context *ctx, int code);
void iw_set_option(struct iw_context *ctx, const char *name, const char *val);
const char *iw_get_option(struct iw_context *ctx, const char *name);

/* ---- small utilities the callers of the path use (reference: src/imagew.h:667-689, :756-802;
 *      src/imagew-api.c:165-207, :686-703, :750-758; src/imagew-util.c:25-197, :556-624) ---- */
int iw_get_sample_size(void);
unsigned int iw_color_get_int_sample(struct iw_color *clr, int channel, unsigned int maxcolorcode);
int iw_get_version_int(void);
char *iw_get_version_string(struct iw_context *ctx, char *s, int s_len);
unsigned int iw_get_profile_by_fmt(int fmt);
int iw_imgtype_num_channels(int t);
size_t iw_calc_bytesperrow(int num_pixels, int bits_per_pixel);
int iw_check_image_dimensions(struct iw_context *ctx, int w, int h);
int iw_is_valid_density(double density_x, double density_y, int density_code);
void *iw_malloc_ex(struct iw_context *ctx, unsigned int flags, size_t n);
void *iw_malloc(struct iw_context *ctx, size_t n);
void *iw_mallocz(struct iw_context *ctx, size_t n);
void *iw_malloc_large(struct iw_context *ctx, size_t n1, size_t n2);
void *iw_realloc_ex(struct iw_context *ctx, unsigned int flags, void *m, size_t oldn, size_t n);
void *iw_realloc(struct iw_context *ctx, void *m, size_t oldn, size_t n);
char *iw_strdup(struct iw_context *ctx, const char *s);
void iw_free(struct iw_context *ctx, void *mem);
void iw_strlcpy(char *dst, const char *src, size_t dstlen);
void iw_snprintf(char *buf, size_t buflen, const char *fmt, ...);
int iw_stricmp(const char *s1, const char *s2);
void iw_zeromem(void *mem, size_t n);

/* ---- host utilities used by the reference's command-line tool and file-format modules (src/imagew.h:667-810;
 * csrc/util.cpp).  They let the UNMODIFIED imagew-cmd.c and imagew-{pnm,bmp,miff,tiff,gif,allfmts}.c link against this
 * library. ---- */
#include <stdarg.h>
#define IW_ENDIAN_BIG 0
#define IW_ENDIAN_LITTLE 1
#define IW_TRANSLATEFLAG_FORMAT 0x0001
#define IW_TRANSLATEFLAG_POSTFORMAT 0x0002
#define IW_TRANSLATEFLAG_ERRORMSG 0x0010
#define IW_TRANSLATEFLAG_WARNINGMSG 0x0020
struct iw_zlib_module;
void iw_vsnprintf(char *buf, size_t buflen, const char *fmt, va_list ap);
void iw_translate(struct iw_context *ctx, unsigned int flags, char *dst, size_t dstlen, const char *src);
void iw_translatev(struct iw_context *ctx, unsigned int flags, char *dst, size_t dstlen, const char *fmt, va_list ap);
void iw_translatef(struct iw_context *ctx, unsigned int flags, char *dst, size_t dstlen, const char *fmt, ...);
char *iw_get_copyright_string(struct iw_context *ctx, char *dst, int dstlen);
void iw_utf8_to_ascii(const char *src, char *dst, int dstlen);
int iw_get_host_endianness(void);
void iw_set_ui16le(iw_byte *b, unsigned int n);
void iw_set_ui32le(iw_byte *b, unsigned int n);
void iw_set_ui16be(iw_byte *b, unsigned int n);
void iw_set_ui32be(iw_byte *b, unsigned int n);
unsigned int iw_get_ui16le(const iw_byte *b);
unsigned int iw_get_ui32le(const iw_byte *b);
int iw_get_i32le(const iw_byte *b);
unsigned int iw_get_ui16be(const iw_byte *b);
unsigned int iw_get_ui32be(const iw_byte *b);
unsigned int iw_get_ui16_e(const iw_byte *b, int endian);
unsigned int iw_get_ui32_e(const iw_byte *b, int endian);
int iw_max_color_to_bitdepth(unsigned int mc);
int iw_detect_fmt_from_filename(const char *fn);
const char *iw_get_fmt_name(int fmt);
int iw_detect_fmt_of_file(const iw_byte *buf, size_t n);
int iw_is_input_fmt_supported(int fmt);
int iw_is_output_fmt_supported(int fmt);
int iw_parse_number_list(const char *s, int max_numbers, double *results);
double iw_parse_number(const char *s);
int iw_round_to_int(double x);
int iw_parse_int(const char *s);
void iw_set_zlib_module(struct iw_context *ctx, struct iw_zlib_module *z);
struct iw_zlib_module *iw_get_zlib_module(struct iw_context *ctx);

/* =========================================== B200 extensions ============================================ */

/* Library identification: "imagew_b200 <version> sm_100a" (or "... host-emulation" for the test-only build). */
const char *iwb200_version(void);

/* Number of this library's kernels launched so far in this process (all contexts). */
unsigned long long iwb200_kernel_launch_count(void);

/* A job is a compiled plan: the decisions iw_prepare_processing() would make for this context (reference:
 * src/imagew-main.c:1899-2248) plus the device-resident weight lists and colour tables.  It lets a caller keep
 * images in HBM and process arbitrary output-row bands (multi-GPU sharding, benchmarks).  The context's input
 * image header must be set (pixels may be NULL when only device buffers are used). */
typedef struct iwb200_job iwb200_job;

struct iwb200_job_info {
  int out_width, out_height, out_imgtype, out_bit_depth, out_sampletype, out_channels;
  size_t out_bpr;           /* tight: ceil(w*bits*channels/8), as the reference allocates (main.c:1385) */
  int in_width_phys, in_height_phys; /* physical (stored) rows/columns of the input buffer */
  size_t in_bpr;
  int bandable;             /* 0 when the job must run as one band (error-diffusion dither) */
  int uses_fused_kernel;    /* 0: generic per-sample kernels; 1 / 2: streaming kernels of the first / second generation;
                             * 4: gather-form streaming kernels (images with an alpha channel) */
  double taps_v, taps_h;    /* total filter taps (mul+add pairs) per full image, all channels */
};

iwb200_job *iwb200_job_create(struct iw_context *ctx, int device);
void iwb200_job_destroy(iwb200_job *job);
int iwb200_job_get_info(const iwb200_job *job, struct iwb200_job_info *info);

/* Physical input rows [first,last] that output rows [out_row0,out_row1) depend on (band + filter halo). */
int iwb200_job_band_input_rows(const iwb200_job *job, int out_row0, int out_row1, int *first, int *last);

/* Run output rows [out_row0,out_row1) on device buffers, asynchronously on `stream` (a cudaStream_t, NULL =
 * default stream).  d_in points at physical input row `in_row_origin` (so a rank may hold only its slab);
 * d_out points at output row out_row0.  Returns 1/0. */
int iwb200_job_run_device(iwb200_job *job, const void *d_in, size_t in_bpr, int in_row_origin, void *d_out,
                          size_t out_bpr, int out_row0, int out_row1, void *stream);

/* Host-buffer variant: H2D of the needed input rows, run, D2H, synchronous.  This is what
 * iw_process_image() calls.  Returns 1/0. */
int iwb200_job_run_host(iwb200_job *job, const void *h_in, size_t in_bpr, void *h_out, size_t out_bpr,
                        int out_row0, int out_row1);

const char *iwb200_job_last_error(const iwb200_job *job);

/* Verification hook for the "intermediate samples" parity bar: after a run, copy the FP32 intermediate plane
 * (the vertical-pass result the reference keeps in `intermediate32`, src/imagew-main.c:905-912, :1383-1446) of
 * processed channel `ch` (alpha first, as the reference orders channels; -1 = the one processed last) for the LAST chunk of output rows the
 * job ran, as a row-major [rows][width] array; dst may be NULL to query the extent.  Returns 1 on success. */
int iwb200_job_read_intermediate(iwb200_job *job, int ch, float *dst, size_t capacity, int *row0, int *rows,
                                 int *width);

/* Per-kernel device timing.  When enabled, every launch of this job is bracketed by CUDA events on the launching
 * stream; iwb200_job_kernel_stats() synchronises and returns, for kernel class `which` (0 vertical pass,
 * 1 horizontal pass + epilogue, 2 PRNG fill, 3 error diffusion, 4 post, 5 vertical streaming kernel, 6 fused
 * vertical + horizontal kernel), the number of launches and their summed duration in ms since the last reset. */
void iwb200_job_set_profiling(iwb200_job *job, int on);
/* Kernel generation the job's last launch took: 0 generic per-sample kernels, 1 / 2 streaming kernels of the first /
 * second generation (two kernels per chunk), 3 the fused persistent vertical + horizontal kernel, 4 the gather-form
 * streaming kernels (alpha); -1 for a null job.
 * Context option "cuda:gen" = 1 | 2 | 3 forces a generation where the job allows it (3: also for small launches). */
int iwb200_job_last_kernel_generation(const iwb200_job *job);
/* ---- batches of same-shaped images --------------------------------------------------------------------------------
 * One plan for many images of the same shape and settings (thumbnailing a directory, BASELINE configuration 4).  `ctx`
 * carries the settings and the input image HEADER (iw_set_input_image with pixels == NULL is enough); it must outlive
 * the batch.  Every listed device gets `depth` slots; iwb200_batch_submit() queues one image (host -> device copy,
 * kernels, device -> host copy of the result) on the next slot and returns at once unless that slot is still busy, so
 * copies and kernels of consecutive images overlap.  `h_out` receives the image as processed (info.out_bpr bytes per
 * row, iwb200_job_get_info); it is complete after iwb200_batch_wait() or after `depth * ndev` later submissions.
 * Page-locked host buffers (iw_malloc_large) keep the copies asynchronous.  A batch is used from one thread at a time;
 * several batches (e.g. one per device) may run from several threads. */
typedef struct iwb200_batch iwb200_batch;
iwb200_batch *iwb200_batch_create(struct iw_context *ctx, const int *devices, int ndev, int depth);
int iwb200_batch_submit(iwb200_batch *b, const void *h_in, size_t in_bpr, void *h_out, size_t out_bpr);
int iwb200_batch_wait(iwb200_batch *b);
const char *iwb200_batch_last_error(const iwb200_batch *b);
int iwb200_batch_get_info(const iwb200_batch *b, struct iwb200_job_info *info);   /* output size, type, depth, row stride */
void iwb200_batch_destroy(iwb200_batch *b);

/* Diagnostics: every wait inside the fused kernel is bounded (seconds); a warp that gives up records what it waited for
 * and the launch ends with an error instead of hanging the device.  Copies the record as text into buf (empty string
 * if there is none), clears it, returns its length. */
int iwb200_debug_fused_trap(char *buf, int buflen);
int iwb200_job_kernel_stats(iwb200_job *job, int which, int *launches, double *total_ms);
void iwb200_job_reset_stats(iwb200_job *job);

/* Test hook: the host-side split of one fast-path launch into "uniform" and "general" parts (csrc/stream2_host.cuh);
 * pure host code.  out[9] = {step, n_parts, n_u, u_lo, u_hi, u_len, g_len, u_ref_o, u_ref_last}. */
int iwb200_debug_plan_parts(const int *first, const int *last, int n_out, int lo, int hi, int base, int align, int S, int want,
                            int len_override, int g_min, int step_wanted, int *out);

/* Test hook: turn a tap list (tap_first[n_out+1] offsets into tap_src / tap_w, as the weight-list builder emits them)
 * into the streaming schedule of the fast path; pure host code.  Returns 0 if the list does not qualify, else 1 with
 * *slots = accumulator slots S, first/last[n_out] = the (possibly widened) windows and w[n_in * S] = the weight of
 * input r for slot s at w[r * S + s] (w may be NULL). */
int iwb200_debug_build_schedule(const int *tap_first, const int *tap_src, const double *tap_w, int n_in, int n_out,
                                int *slots, int *first, int *last, double *w);

/* Test hook: the level model of the 16-bit epilogue (csrc/estimate.h); pure host code, the same arithmetic as the kernel.
 * Builds the model for a colour space (IW_CSTYPE_*) and a table thr[ntab] (linear value of every code), stores the
 * measured margin / worst error / number of unvalidated segments in stats[3] and, for each of the n samples v[i], the
 * code the model settles on or -1 when the sample is left to the table.  Returns 0 if no model exists. */
int iwb200_debug_level_model(int cstype, const double *thr, int ntab, const double *v, int n, int *code, double *stats);

/* Deterministic synthetic input and output fingerprint used by the benchmarks and the full-size parity tests
 * (SURVEY.md section 8c): xorshift64 noise (s^=s<<13; s^=s>>7; s^=s<<17; byte = s>>24), FNV-1a 64-bit. */
void iwb200_synth_xorshift(uint8_t *buf, size_t n, uint64_t seed);
uint64_t iwb200_fnv1a64(const uint8_t *buf, size_t n);

/* Forget the input pixels without freeing them: for callers that lend a buffer they own (see INTEGRATION.md
 * section 2) instead of transferring ownership as iw_set_input_image() normally does. */
void iwb200_detach_input_image(struct iw_context *ctx);

/* After iw_process_image() with option "cuda:band" set: which output rows the output image holds. */
int iwb200_get_band(struct iw_context *ctx, int *out_row0, int *out_row1);

/* Millisecond timings of the last iw_process_image() on this context: plan build, H2D, kernels, D2H. */
int iwb200_get_timings(struct iw_context *ctx, double *plan_ms, double *h2d_ms, double *kernel_ms,
                       double *d2h_ms);

#if defined(IWB_BUILDING) && defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* IMAGEW_B200_H */
