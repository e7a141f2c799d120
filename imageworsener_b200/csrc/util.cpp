// util.cpp -- the small host utilities of the libimagew API that the reference's command-line tool and its file-format
// modules call (reference: src/imagew.h:667-810, behaviour as documented there and in src/imagew-util.c): number and
// list parsing, little/big-endian field access, format detection by name and by signature, UTF-8 -> ASCII for consoles
// that cannot print the library's messages, string translation hooks.  No pixel ever passes through this file; it
// exists so that the unmodified tool and codecs link against this library (tests/test_cli.py).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "iwb_internal.h"

extern "C" {

// ---- formatted output with a fixed-size destination -----------------------------------------------------------------
void iw_vsnprintf(char *buf, size_t buflen, const char *fmt, va_list ap) {
  if (!buf || buflen == 0) return;
  vsnprintf(buf, buflen, fmt, ap);
  buf[buflen - 1] = '\0';
}

// ---- translation hooks (src/imagew.h:514-519, :746-752): the caller's function may replace a string; if it declines
// (returns 0) or there is none, the string is used as it is
void iw_translate(struct iw_context *ctx, unsigned int flags, char *dst, size_t dstlen, const char *src) {
  dst[0] = '\0';
  const int done = (ctx && ctx->translate_fn) ? ctx->translate_fn(ctx, flags, dst, dstlen, src) : 0;
  if (!done) iw_strlcpy(dst, src, dstlen);
}
void iw_translatev(struct iw_context *ctx, unsigned int flags, char *dst, size_t dstlen, const char *fmt, va_list ap) {
  if (!ctx || !ctx->translate_fn) { iw_vsnprintf(dst, dstlen, fmt, ap); return; }
  char translated_fmt[IWB_MSG_MAX], formatted[IWB_MSG_MAX];
  iw_translate(ctx, IW_TRANSLATEFLAG_FORMAT | flags, translated_fmt, sizeof translated_fmt, fmt);
  iw_vsnprintf(formatted, sizeof formatted, translated_fmt, ap);
  iw_translate(ctx, IW_TRANSLATEFLAG_POSTFORMAT | flags, dst, dstlen, formatted);
}
void iw_translatef(struct iw_context *ctx, unsigned int flags, char *dst, size_t dstlen, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  iw_translatev(ctx, flags, dst, dstlen, fmt, ap);
  va_end(ap);
}
char *iw_get_copyright_string(struct iw_context *ctx, char *dst, int dstlen) {
  iw_translatef(ctx, 0, dst, (size_t)dstlen, "Copyright \xc2\xa9 %s %s", "2011\xe2\x80\x93" "2022", "Jason Summers");
  return dst;
}

// ---- UTF-8 -> ASCII (src/imagew.h:789): code points above 127 become a short ASCII stand-in, '?' if there is none
void iw_utf8_to_ascii(const char *src, char *dst, int dstlen) {
  static const struct { unsigned cp; const char *ascii; } subst[] = {
      {0x00a9, "(c)"}, {0x00d7, "x"}, {0x2013, "-"}, {0x2018, "'"}, {0x2019, "'"}, {0x201c, "\""}, {0x201d, "\""}, {0x2192, "->"}, {0xfeff, ""}};
  int n = 0;
  auto put = [&](char c) { if (n <= dstlen - 2) dst[n++] = c; };
  unsigned cp = 0;
  int more = 0;                      // continuation bytes still expected
  for (const unsigned char *p = (const unsigned char *)src; *p; p++) {
    const unsigned char c = *p;
    if (c < 0x80) { put((char)c); more = 0; }
    else if (c < 0xc0) {             // continuation byte
      if (more > 0) {
        cp = (cp << 6) | (c & 0x3f);
        if (--more == 0) {
          const char *a = "?";
          for (const auto &s : subst) if (s.cp == cp) { a = s.ascii; break; }
          for (; *a; a++) put(*a);
        }
      }
    }
    else if (c < 0xe0) { cp = c & 0x1f; more = 1; }
    else if (c < 0xf0) { cp = c & 0x0f; more = 2; }
    else if (c < 0xf8) { cp = c & 0x07; more = 3; }
  }
  dst[n] = '\0';
}

// ---- byte order ----------------------------------------------------------------------------------------------------
int iw_get_host_endianness(void) {
  const unsigned int one = 1;
  unsigned char first;
  memcpy(&first, &one, 1);
  return first ? IW_ENDIAN_LITTLE : IW_ENDIAN_BIG;
}
void iw_set_ui16le(iw_byte *b, unsigned int n) { b[0] = (iw_byte)(n & 0xff); b[1] = (iw_byte)((n >> 8) & 0xff); }
void iw_set_ui32le(iw_byte *b, unsigned int n) { for (int k = 0; k < 4; k++) b[k] = (iw_byte)((n >> (8 * k)) & 0xff); }
void iw_set_ui16be(iw_byte *b, unsigned int n) { b[0] = (iw_byte)((n >> 8) & 0xff); b[1] = (iw_byte)(n & 0xff); }
void iw_set_ui32be(iw_byte *b, unsigned int n) { for (int k = 0; k < 4; k++) b[k] = (iw_byte)((n >> (8 * (3 - k))) & 0xff); }
unsigned int iw_get_ui16le(const iw_byte *b) { return (unsigned int)b[0] | ((unsigned int)b[1] << 8); }
unsigned int iw_get_ui32le(const iw_byte *b) { return iw_get_ui16le(b) | (iw_get_ui16le(b + 2) << 16); }
int iw_get_i32le(const iw_byte *b) { return (int)iw_get_ui32le(b); }
unsigned int iw_get_ui16be(const iw_byte *b) { return ((unsigned int)b[0] << 8) | (unsigned int)b[1]; }
unsigned int iw_get_ui32be(const iw_byte *b) { return (iw_get_ui16be(b) << 16) | iw_get_ui16be(b + 2); }
unsigned int iw_get_ui16_e(const iw_byte *b, int endian) { return endian == IW_ENDIAN_LITTLE ? iw_get_ui16le(b) : iw_get_ui16be(b); }
unsigned int iw_get_ui32_e(const iw_byte *b, int endian) { return endian == IW_ENDIAN_LITTLE ? iw_get_ui32le(b) : iw_get_ui32be(b); }

// smallest depth (1..16) whose largest code is >= mc
int iw_max_color_to_bitdepth(unsigned int mc) {
  for (int bits = 1; bits < 16; bits++) if (mc < (1u << bits)) return bits;
  return 16;
}

// ---- file formats by name and by signature ---------------------------------------------------------------------------
int iw_detect_fmt_from_filename(const char *fn) {
  static const struct { const char *ext; int fmt; } known[] = {
      {"png", IW_FORMAT_PNG}, {"jpg", IW_FORMAT_JPEG}, {"jpeg", IW_FORMAT_JPEG}, {"bmp", IW_FORMAT_BMP}, {"tif", IW_FORMAT_TIFF},
      {"tiff", IW_FORMAT_TIFF}, {"miff", IW_FORMAT_MIFF}, {"webp", IW_FORMAT_WEBP}, {"gif", IW_FORMAT_GIF}, {"pnm", IW_FORMAT_PNM},
      {"pbm", IW_FORMAT_PBM}, {"pgm", IW_FORMAT_PGM}, {"ppm", IW_FORMAT_PPM}, {"pam", IW_FORMAT_PAM}};
  const char *dot = strrchr(fn, '.');
  if (!dot) return IW_FORMAT_UNKNOWN;
  for (const auto &k : known) if (!iw_stricmp(dot + 1, k.ext)) return k.fmt;
  return IW_FORMAT_UNKNOWN;
}
const char *iw_get_fmt_name(int fmt) {
  switch (fmt) {
    case IW_FORMAT_PNG: return "PNG";
    case IW_FORMAT_JPEG: return "JPEG";
    case IW_FORMAT_BMP: return "BMP";
    case IW_FORMAT_TIFF: return "TIFF";
    case IW_FORMAT_MIFF: return "MIFF";
    case IW_FORMAT_WEBP: return "WebP";
    case IW_FORMAT_GIF: return "GIF";
    case IW_FORMAT_PNM: return "PNM";
    case IW_FORMAT_PBM: return "PBM";
    case IW_FORMAT_PGM: return "PGM";
    case IW_FORMAT_PPM: return "PPM";
    case IW_FORMAT_PAM: return "PAM";
  }
  return nullptr;
}
int iw_detect_fmt_of_file(const iw_byte *buf, size_t n) {
  if (n < 2) return IW_FORMAT_UNKNOWN;
  const iw_byte a = buf[0], b = buf[1];
  if (a == 0x89 && b == 'P') return IW_FORMAT_PNG;
  if (n >= 3 && a == 'G' && b == 'I' && buf[2] == 'F') return IW_FORMAT_GIF;
  if (a == 0xff && b == 0xd8) return IW_FORMAT_JPEG;
  if (a == 'B' && (b == 'M' || b == 'A')) return IW_FORMAT_BMP;
  if ((a == 'I' || a == 'M') && b == a) return IW_FORMAT_TIFF;
  if (a == 'i' && b == 'd') return IW_FORMAT_MIFF;
  if (n >= 12 && !memcmp(buf, "RIFF", 4) && !memcmp(buf + 8, "WEBP", 4)) return IW_FORMAT_WEBP;
  if (a == 'P' && b >= '1' && b <= '6') return IW_FORMAT_PNM;
  if (a == 'P' && b == '7' && buf[2] == 0x0a) return IW_FORMAT_PAM;
  return IW_FORMAT_UNKNOWN;
}

// Which file formats the auxiliary modules that ship with this build can read / write: the dependency-free ones (PNG,
// JPEG and WebP need third-party libraries this library is not built with).
int iw_is_input_fmt_supported(int fmt) {
  return fmt == IW_FORMAT_MIFF || fmt == IW_FORMAT_GIF || fmt == IW_FORMAT_BMP || fmt == IW_FORMAT_PNM || fmt == IW_FORMAT_PAM;
}
int iw_is_output_fmt_supported(int fmt) {
  return fmt == IW_FORMAT_BMP || fmt == IW_FORMAT_TIFF || fmt == IW_FORMAT_MIFF || fmt == IW_FORMAT_PNM || fmt == IW_FORMAT_PPM ||
         fmt == IW_FORMAT_PGM || fmt == IW_FORMAT_PBM || fmt == IW_FORMAT_PAM;
}

// ---- numbers on the command line: decimal, or a ratio "a/b"; lists separated by commas (src/imagew.h:795-802) --------
static bool parse_one(const char *s, double *val, int *used) {
  *val = 0.0;
  int len = 0, slash = -1;
  while (s[len] && s[len] != ',') { if (s[len] == '/') slash = len; len++; }
  *used = 0;
  if (len < 1) return false;
  *used = len;
  if (slash >= 0) {
    const double den = atof(s + slash + 1);
    *val = den == 0.0 ? 0.0 : atof(s) / den;
  } else {
    *val = atof(s);
  }
  return true;
}
int iw_parse_number_list(const char *s, int max_numbers, double *results) {
  int count = 0, pos = 0;
  while (count < max_numbers) {
    int used;
    if (!parse_one(s + pos, &results[count], &used)) break;
    count++;
    pos += used;
    if (s[pos] != ',') break;
    pos++;
  }
  return count;
}
double iw_parse_number(const char *s) {
  double v;
  int used;
  parse_one(s, &v, &used);
  return v;
}
int iw_round_to_int(double x) { return x < 0.0 ? -(int)(0.5 - x) : (int)(0.5 + x); }
int iw_parse_int(const char *s) { return iw_round_to_int(iw_parse_number(s)); }

// ---- the optional zlib module is only carried for the codecs (this library never calls it) ----------------------------
void iw_set_zlib_module(struct iw_context *ctx, struct iw_zlib_module *z) { ctx->zlib_module = z; }
struct iw_zlib_module *iw_get_zlib_module(struct iw_context *ctx) { return ctx->zlib_module; }

}  // extern "C"
