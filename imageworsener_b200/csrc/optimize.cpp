// optimize.cpp -- decision logic of the post-pass optimiser, restated from the reference step by step
// (iwpvt_optimize_image, src/imagew-opt.c:1166-1314; palette / low-gray choice :1000-1110; key-colour search :365-392).
// The data-parallel parts (scan, rewrite) are device kernels (optimize.h, pipeline.cu); this file only turns a scan
// into the list of reductions to apply and composes them into one rewrite recipe.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "iwb_internal.h"
#include "optimize.h"

void iwb_opt_merge(IwbOptScan *into, const IwbOptScan &from) {
  into->flags |= from.flags;
  for (int t = 0; t < 4; t++)
    for (int k = 0; k < 8; k++) into->used[t][k] |= from.used[t][k];
  if (from.ncolors > 256 || into->ncolors > 256) { into->ncolors = 257; return; }
  for (unsigned i = 0; i < from.ncolors; i++) {
    const unsigned key = from.colors[i];
    bool seen = false;
    for (unsigned j = 0; j < into->ncolors && !seen; j++) seen = into->colors[j] == key;
    if (seen) continue;
    if (into->ncolors >= 256) { into->ncolors = 257; return; }
    into->colors[into->ncolors++] = key;
  }
}

namespace {

// reference: iwopt_find_unused, src/imagew-opt.c:365-392 (nearest free value to 192)
bool find_unused(const unsigned (&used)[8], int count, unsigned *key) {
  bool found = false;
  int best = 0;
  for (int i = 0; i < count; i++) {
    if (used[i >> 5] & (1u << (i & 31))) continue;
    const int d = abs(i - 192);
    if (!found || d < best) { *key = (unsigned)i; found = true; best = d; }
  }
  if (!found) *key = 0;
  return found;
}

struct rgba8 { unsigned char r, g, b, a; };
rgba8 unpack(unsigned k) { return {(unsigned char)(k & 255), (unsigned char)((k >> 8) & 255), (unsigned char)((k >> 16) & 255), (unsigned char)(k >> 24)}; }
unsigned pack(const rgba8 &c) { return c.r | (c.g << 8) | (c.b << 16) | ((unsigned)c.a << 24); }

// reference: iwopt_palsortfunc, src/imagew-opt.c:866-897
bool pal_less(const rgba8 &e1, const rgba8 &e2) {
  if (e1.a == 255 && e2.a < 255) return false;
  if (e1.a < 255 && e2.a == 255) return true;
  const int s1 = e1.r + e1.g + e1.b, s2 = e2.r + e2.g + e2.b;
  if (s1 != s2) return s1 < s2;
  if (e1.g != e2.g) return e1.g < e2.g;
  if (e1.r != e2.r) return e1.r < e2.r;
  if (e1.b != e2.b) return e1.b < e2.b;
  return e1.a < e2.a;
}

}  // namespace

int iwb_opt_decide(iw_context *ctx, const IwbOptScan &scan, IwbOptOutcome *out) {
  memset(out, 0, sizeof *out);
  OptRecipe &rc = out->recipe;
  int T = ctx->img2.imgtype, D = ctx->img2.bit_depth;
  const int src_nch = iw_imgtype_num_channels(T);
  rc.src_nch = src_nch; rc.src_bps = D / 8; rc.src_has_alpha = IW_IMGTYPE_HAS_ALPHA(T) ? 1 : 0;
  rc.pal_trns_index = -1;
  out->imgtype = T; out->bit_depth = D; out->bpr = ctx->img2.bpr;
  if (ctx->img2.sampletype != IW_SAMPLETYPE_UINT || ctx->reduced_out_maxcode) return 1;   // :1196-1204

  const unsigned prof = ctx->output_profile;
  const bool opt_gray = ctx->opt_flags[IW_OPT_GRAYSCALE], opt_pal = ctx->opt_flags[IW_OPT_PALETTE], opt_16_8 = ctx->opt_flags[IW_OPT_16_TO_8],
             opt_strip = ctx->opt_flags[IW_OPT_STRIP_ALPHA], opt_btrns = ctx->opt_flags[IW_OPT_BINARY_TRNS];
  bool has_color = scan.flags & IWB_OPTF_COLOR, has_trns = scan.flags & IWB_OPTF_TRNS, has_partial = scan.flags & IWB_OPTF_PARTIAL,
       has_16 = scan.flags & IWB_OPTF_16BIT;
  // the background-colour label constrains the reductions (:1208-1236)
  unsigned label[4] = {ctx->res_label[0], ctx->res_label[1], ctx->res_label[2], ctx->res_label[3]};
  const bool has_label = ctx->res_has_label != 0;
  if (has_label) {
    if (label[0] != label[1] || label[0] != label[2]) has_color = true;
    if (D == 16 && (label[0] % 257 != 0 || label[1] % 257 != 0 || label[2] % 257 != 0)) has_16 = true;
  }
  // (8-bit gray images are not scanned by the reference, :337-362; their flags are all clear, as are ours)

  int map[4] = {0, 1, 2, 3};        // source channel behind each current channel
  int nch = src_nch;
  bool high_byte = false;           // the 8-bit image is the high bytes of the 16-bit one

  // 16 -> 8 bits (:1242-1256, iw_opt_16_to_8 :243-277)
  if (D == 16 && !has_16 && opt_16_8) {
    D = 8; high_byte = true;
    if (has_label) for (int k = 0; k < 4; k++) label[k] >>= 8;
  }
  // drop an alpha channel that is opaque everywhere (:1258-1272)
  if (T == IW_IMGTYPE_RGBA && !has_trns && opt_strip) { T = IW_IMGTYPE_RGB; nch = 3; }
  if (T == IW_IMGTYPE_GRAYA && !has_trns && opt_strip) { T = IW_IMGTYPE_GRAY; nch = 1; }
  // colour -> gray (:1274-1296)
  if (T == IW_IMGTYPE_RGB && !has_color && (prof & IW_PROFILE_GRAYSCALE) && opt_gray) { T = IW_IMGTYPE_GRAY; nch = 1; map[0] = 0; }
  if (T == IW_IMGTYPE_RGBA && !has_color && (prof & IW_PROFILE_GRAYSCALE) && opt_gray) { T = IW_IMGTYPE_GRAYA; nch = 2; map[0] = 0; map[1] = 3; }

  bool has_colorkey = false;
  unsigned colorkey[3] = {0, 0, 0};
  bool palette_image = false;
  rgba8 pal[256];
  int pal_n = 0;
  bool pal_is_gray = false;

  // palette / low-depth gray (iwopt_try_pal_lowgray_optimization, :1000-1110)
  if ((prof & (IW_PROFILE_PAL1 | IW_PROFILE_PAL2 | IW_PROFILE_PAL4 | IW_PROFILE_PAL8 | IW_PROFILE_GRAY1 | IW_PROFILE_GRAY2 | IW_PROFILE_GRAY4)) && D == 8 &&
      scan.ncolors >= 1 && scan.ncolors <= 256) {
    pal_n = (int)scan.ncolors;
    for (int i = 0; i < pal_n; i++) pal[i] = unpack(scan.colors[i]);
    bool usable = true;
    if (has_label) {                // the label needs an entry it can use (:776-794)
      bool found = false;
      for (int i = 0; i < pal_n && !found; i++)
        found = pal[i].a == 0 || (pal[i].r == label[0] && pal[i].g == label[1] && pal[i].b == label[2]);
      if (!found) {
        if (pal_n < 256) pal[pal_n++] = {(unsigned char)label[0], (unsigned char)label[1], (unsigned char)label[2], 255};
        else usable = false;
      }
    }
    if (usable) {
      auto pal_ok = [&](int bpp) {  // iwopt_palette_opt_ok, :899-930
        if (!opt_pal) return false;
        if (has_trns && !(prof & IW_PROFILE_PALETTETRNS)) return false;
        switch (bpp) {
          case 1: return (prof & IW_PROFILE_PAL1) && pal_n <= 2;
          case 2: return (prof & IW_PROFILE_PAL2) && pal_n <= 4;
          case 4: return (prof & IW_PROFILE_PAL4) && pal_n <= 16;
          case 8: return (prof & IW_PROFILE_PAL8) && pal_n <= 256;
        }
        return false;
      };
      bool binary_trns = false;
      unsigned trns_shade = 0;
      auto valid_gray = [&](int bpp) {   // iwopt_palette_is_valid_gray, :932-995
        binary_trns = false; trns_shade = 0;
        if (has_color || has_partial) return false;
        if (has_trns && !opt_btrns) return false;
        if (has_trns && !(prof & IW_PROFILE_BINARYTRNS)) return false;
        int factor, max_entries;
        switch (bpp) {
          case 1: factor = 255; max_entries = 2; break;
          case 2: factor = 85; max_entries = 4; break;
          case 4: factor = 17; max_entries = 16; break;
          case 8: factor = 1; max_entries = 256; break;
          default: return false;
        }
        if (pal_n > max_entries) return false;
        if (has_label && bpp < 8 && label[0] % factor != 0) return false;
        unsigned used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < pal_n; i++)
          if (pal[i].a > 0) {
            if (pal[i].r % factor) return false;
            const int v = pal[i].r / factor;
            used[v >> 5] |= 1u << (v & 31);
          }
        if (has_trns) {
          if (!find_unused(used, max_entries, &trns_shade)) return false;
          binary_trns = true;
        }
        return true;
      };
      auto gray_palette = [&](int bpp) {   // iwopt_make_gray_palette, :997-1017
        const int factor = bpp == 1 ? 255 : (bpp == 2 ? 85 : (bpp == 4 ? 17 : 1));
        pal_n = bpp == 1 ? 2 : (bpp == 2 ? 4 : (bpp == 4 ? 16 : 256));
        for (int i = 0; i < pal_n; i++) pal[i] = {(unsigned char)(factor * i), (unsigned char)(factor * i), (unsigned char)(factor * i), 255};
        pal_is_gray = true;
        if (binary_trns) { has_colorkey = true; colorkey[0] = colorkey[1] = colorkey[2] = trns_shade; }
      };
      int choice = 0;               // order of preference: gray1, pal1, gray2, pal2, gray4, pal4, gray8 (not here), pal8
      if (opt_gray && (prof & IW_PROFILE_GRAY1) && valid_gray(1)) { gray_palette(1); choice = 1; }
      else if (pal_ok(1)) choice = 1;
      else if (opt_gray && (prof & IW_PROFILE_GRAY2) && valid_gray(2)) { gray_palette(2); choice = 1; }
      else if (pal_ok(2)) choice = 1;
      else if (opt_gray && (prof & IW_PROFILE_GRAY4) && valid_gray(4)) { gray_palette(4); choice = 1; }
      else if (pal_ok(4)) choice = 1;
      else if (opt_gray && (prof & IW_PROFILE_GRAYSCALE) && valid_gray(8)) choice = 0;   // best left as 8-bit gray
      else if (pal_ok(8)) choice = 1;
      if (choice) {
        if (!pal_is_gray) std::sort(pal, pal + pal_n, pal_less);
        palette_image = true;
      }
    }
  }

  if (palette_image) {
    T = IW_IMGTYPE_PALETTE; D = 8;
    rc.palette = 1;
    rc.pal_n = pal_n;
    for (int i = 0; i < pal_n; i++) rc.pal_keys[i] = pack(pal[i]);
    if (has_colorkey) rc.pal_trns_index = (int)colorkey[0];
    iw_palette *P = (iw_palette *)iw_malloc(ctx, sizeof(iw_palette));
    if (!P) return 0;
    memset(P, 0, sizeof *P);
    P->num_entries = pal_n;
    for (int i = 0; i < pal_n; i++) { P->entry[i].r = pal[i].r; P->entry[i].g = pal[i].g; P->entry[i].b = pal[i].b; P->entry[i].a = pal[i].a; }
    if (ctx->palette) iw_free(ctx, ctx->palette);
    ctx->palette = P;
    ctx->palette_is_gray = pal_is_gray ? 1 : 0;
    out->has_palette = 1;
  } else {
    // alpha -> colour-key transparency (:1300-1313 and :410-700)
    const bool can = (prof & IW_PROFILE_BINARYTRNS) && opt_btrns && !has_partial;
    unsigned key = 0;
    if (can && T == IW_IMGTYPE_RGBA && D == 8 && find_unused(scan.used[0], 256, &key)) {
      T = IW_IMGTYPE_RGB; nch = 3; has_colorkey = true;
      colorkey[0] = key; colorkey[1] = 192; colorkey[2] = 192;
      rc.key_on = 1; rc.key_len = 3; rc.key_bytes[0] = (unsigned char)key; rc.key_bytes[1] = 192; rc.key_bytes[2] = 192;
    } else if (can && T == IW_IMGTYPE_RGBA && D == 16 && find_unused(scan.used[1], 256, &key)) {
      T = IW_IMGTYPE_RGB; nch = 3; has_colorkey = true;
      colorkey[0] = 192 * 256 + key; colorkey[1] = colorkey[2] = 192 * 256 + 192;
      rc.key_on = 1; rc.key_len = 6;
      const unsigned char kb[6] = {192, (unsigned char)key, 192, 192, 192, 192};
      memcpy(rc.key_bytes, kb, 6);
    } else if (can && T == IW_IMGTYPE_GRAYA && D == 8 && find_unused(scan.used[2], 256, &key)) {
      T = IW_IMGTYPE_GRAY; nch = 1; has_colorkey = true;
      colorkey[0] = colorkey[1] = colorkey[2] = key;
      rc.key_on = 1; rc.key_len = 1; rc.key_bytes[0] = (unsigned char)key;
    } else if (can && T == IW_IMGTYPE_GRAYA && D == 16 && find_unused(scan.used[3], 256, &key)) {
      T = IW_IMGTYPE_GRAY; nch = 1; has_colorkey = true;
      colorkey[0] = colorkey[1] = colorkey[2] = 192 * 256 + key;
      rc.key_on = 1; rc.key_len = 2; rc.key_bytes[0] = 192; rc.key_bytes[1] = (unsigned char)key;
    }
    rc.newnc = nch;
    for (int k = 0; k < nch; k++) rc.cmap[k] = map[k];
    rc.dst_bps = D / 8;
  }
  (void)high_byte;                  // (an 8-bit destination of a 16-bit source takes the first, i.e. high, byte: OptRewriteBody)

  out->imgtype = T; out->bit_depth = D;
  out->bpr = palette_image ? (size_t)ctx->img2.width : iw_calc_bytesperrow(ctx->img2.width, D * nch);
  out->has_colorkey = has_colorkey;
  for (int k = 0; k < 3; k++) out->colorkey[k] = colorkey[k];
  out->changed = (T != ctx->img2.imgtype) || (D != ctx->img2.bit_depth);
  // the label follows the image's depth
  if (has_label) for (int k = 0; k < 4; k++) ctx->res_label[k] = label[k];
  return 1;
}
