// optimize.h -- the post-pass output optimiser (reference: iwpvt_optimize_image and helpers, src/imagew-opt.c:18-1314).
//
// The reference walks the finished image on the host several times: one scan for "is it really gray / opaque / 8-bit",
// an O(pixels x palette) search for the distinct colours, and one full copy per reduction it applies (16 -> 8 bits, drop
// alpha, RGB -> gray, palette indices, colour-key transparency).  Here the image is still in HBM when the pipeline ends,
// so the two data-parallel parts run there:
//
//   scan     one pass over the output: the four flags, the set of distinct colours (a small hash set, abandoned as
//            soon as it outgrows a palette), and the four "which key colours are free" bitmaps of the binary-
//            transparency search (src/imagew-opt.c:18-241, :719-797, :410-700);
//   rewrite  one pass that applies the COMPOSITION of all reductions the decision logic chose (src/imagew-opt.c:243-363,
//            :799-864 and the patching loops of :410-700), producing the final bytes directly.
//
// The decision logic between the two (which reductions, the palette's order, the key colour) is the reference's, step
// by step, on the host (optimize.cpp).  With several devices every band is scanned where it lives and the host merges
// the scans (flags and bitmaps OR, colour sets union) before deciding, so all bands are rewritten the same way.
#pragma once
#include <cstdint>

#include "kernels.cuh"

#define IWB_OPT_HASH 1024          // slots of the colour hash set (power of two, > 4 x 257)
#define IWB_OPT_MAXCOL 320         // colours recorded before the set is abandoned (a palette holds 256)

enum { IWB_OPTF_COLOR = 1, IWB_OPTF_TRNS = 2, IWB_OPTF_PARTIAL = 4, IWB_OPTF_16BIT = 8 };

struct IwbOptScan {                // lives in device memory during the scan, copied to the host afterwards
  unsigned flags;
  unsigned ncolors;                // distinct colours seen; > 256 means "too many" (the list is then incomplete)
  unsigned colors[IWB_OPT_MAXCOL]; // r | g << 8 | b << 16 | a << 24, 8-bit values (high bytes of 16-bit samples)
  unsigned used[4][8];             // 256-bit maps: rgb8, rgb16, gray8, gray16 key colours in use (src/imagew-opt.c:410-700)
  unsigned hash[IWB_OPT_HASH];     // colour + 1 (0 = empty); colour 0xffffffff is kept in `has_ffffffff`
  unsigned has_ffffffff;
};

struct OptScanParams {
  DevOutput out;                   // pix points at row `row_origin`
  int row0, row1;
  int nch, bps;                    // channels, bytes per sample (1 or 2)
  int has_alpha;
  int want_colors;
  IwbOptScan *res;
};

IWB_HD void iwb_opt_or(unsigned *p, unsigned v) {
#if defined(__CUDA_ARCH__)
  if (v & ~*(volatile unsigned *)p) atomicOr(p, v);
#else
  *p |= v;
#endif
}

// insert a colour into the hash set; appends it to the list if it is new
IWB_HD void iwb_opt_add_color(IwbOptScan *r, unsigned key) {
  if (key == 0xffffffffu) {
#if defined(__CUDA_ARCH__)
    if (atomicExch(&r->has_ffffffff, 1u) != 0u) return;
#else
    if (r->has_ffffffff) return;
    r->has_ffffffff = 1;
#endif
  } else {
    unsigned h = (key * 2654435761u) >> 22;                     // 10 bits
    int probes = 0;
    for (;; probes++) {
      if (probes == IWB_OPT_HASH) break;                        // table full (far more than a palette's worth): just count
#if defined(__CUDA_ARCH__)
      const unsigned prev = atomicCAS(&r->hash[h], 0u, key + 1u);
#else
      const unsigned prev = r->hash[h];
      if (prev == 0u) r->hash[h] = key + 1u;
#endif
      if (prev == key + 1u) return;                             // already there
      if (prev == 0u) break;                                    // inserted
      h = (h + 1) & (IWB_OPT_HASH - 1);
    }
  }
#if defined(__CUDA_ARCH__)
  const unsigned idx = atomicAdd(&r->ncolors, 1u);
#else
  const unsigned idx = r->ncolors++;
#endif
  if (idx < IWB_OPT_MAXCOL) r->colors[idx] = key;
}

// one thread per pixel
struct OptScanBody {
  static IWB_HD void run(const OptScanParams &p, size_t idx) {
    const int W = p.out.width;
    const int y = p.row0 + (int)(idx / W), x = (int)(idx % W);
    const uint8_t *px = p.out.pix + (size_t)(y - p.out.row_origin) * p.out.bpr + (size_t)x * p.nch * p.bps;
    unsigned hi[4] = {0, 0, 0, 255}, lo[4] = {0, 0, 0, 255};
    for (int c = 0; c < p.nch; c++) {
      hi[c] = px[c * p.bps];
      lo[c] = p.bps == 2 ? px[c * 2 + 1] : hi[c];
    }
    const int nc = p.has_alpha ? p.nch - 1 : p.nch;             // colour channels: 1 or 3
    unsigned r = hi[0], g = nc == 3 ? hi[1] : hi[0], b = nc == 3 ? hi[2] : hi[0];
    unsigned rl = lo[0], gl = nc == 3 ? lo[1] : lo[0], bl = nc == 3 ? lo[2] : lo[0];
    unsigned a = 255, al = 255;
    if (p.has_alpha) { a = hi[p.nch - 1]; al = lo[p.nch - 1]; }
    unsigned f = 0;
    if (nc == 3 && (r != g || r != b || (p.bps == 2 && (rl != gl || rl != bl)))) f |= IWB_OPTF_COLOR;
    if (p.has_alpha) {
      const bool opaque = a == 255 && al == 255, clear = a == 0 && al == 0;
      if (!opaque) { f |= IWB_OPTF_TRNS; if (!clear) f |= IWB_OPTF_PARTIAL; }
    }
    if (p.bps == 2) for (int c = 0; c < p.nch; c++) if (hi[c] != lo[c]) f |= IWB_OPTF_16BIT;
    if (f) iwb_opt_or(&p.res->flags, f);
    // key colours in use among the pixels that are not fully transparent (src/imagew-opt.c:438-460, :503-528, :575-592, :641-662)
    const bool clear = p.has_alpha && a == 0 && al == 0;
    if (p.has_alpha && !clear) {
      if (nc == 3) {
        if (g == 192 && b == 192) iwb_opt_or(&p.res->used[0][r >> 5], 1u << (r & 31));
        if (p.bps == 2 && r == 192 && g == 192 && gl == 192 && b == 192 && bl == 192) iwb_opt_or(&p.res->used[1][rl >> 5], 1u << (rl & 31));
      }
      iwb_opt_or(&p.res->used[2][r >> 5], 1u << (r & 31));
      if (p.bps == 2 && r == 192) iwb_opt_or(&p.res->used[3][rl >> 5], 1u << (rl & 31));
    }
    if (p.want_colors && *(volatile unsigned *)&p.res->ncolors <= 257u) {
      unsigned key = clear ? 0u : (r | (g << 8) | (b << 16) | (a << 24));
      iwb_opt_add_color(p.res, key);
    }
  }
};

// ---- rewrite: the composition of the reductions, one thread per pixel ------------------------------------------------
struct OptRecipe {
  int src_nch, src_bps, src_has_alpha;
  int palette;                     // 1: destination is one palette index per pixel
  int newnc, cmap[4];              // otherwise: destination channels and the source channel behind each
  int dst_bps;                     // 1 or 2 (2 only if the source is 16-bit and keeps its precision)
  int key_on;                      // fully transparent pixels are written as the key colour (colour-key transparency)
  int key_len;
  unsigned char key_bytes[8];
  int pal_n;
  int pal_trns_index;              // >= 0: palette index of fully transparent pixels (gray palette with a key shade)
  unsigned pal_keys[256];          // r | g << 8 | b << 16 | a << 24 of every palette entry, in palette order
};
struct OptRewriteParams {
  DevOutput out;                   // source image
  int row0, row1;
  uint8_t *dst;                    // destination rows [row0, row1), tightly pitched by dst_bpr
  size_t dst_bpr;
  const OptRecipe *rc;             // device memory
};
struct OptRewriteBody {
  static IWB_HD void run(const OptRewriteParams &p, size_t idx) {
    const OptRecipe &rc = *p.rc;
    const int W = p.out.width;
    const int y = p.row0 + (int)(idx / W), x = (int)(idx % W);
    const uint8_t *px = p.out.pix + (size_t)(y - p.out.row_origin) * p.out.bpr + (size_t)x * rc.src_nch * rc.src_bps;
    bool clear = false;
    if (rc.src_has_alpha) {
      const uint8_t *ap = px + (rc.src_nch - 1) * rc.src_bps;
      clear = ap[0] == 0 && (rc.src_bps == 1 || ap[1] == 0);
    }
    if (rc.palette) {
      uint8_t *d = p.dst + (size_t)(y - p.row0) * p.dst_bpr + x;
      if (clear && rc.pal_trns_index >= 0) { *d = (uint8_t)rc.pal_trns_index; return; }
      const int nc = rc.src_has_alpha ? rc.src_nch - 1 : rc.src_nch;
      const unsigned r = px[0], g = nc == 3 ? px[rc.src_bps] : r, b = nc == 3 ? px[2 * rc.src_bps] : r;
      const unsigned a = rc.src_has_alpha ? px[(rc.src_nch - 1) * rc.src_bps] : 255u;
      const unsigned key = clear ? 0u : (r | (g << 8) | (b << 16) | (a << 24));
      int e = 0;                                                // (not found: index 0, as the reference)
      for (int i = 0; i < rc.pal_n; i++) if (rc.pal_keys[i] == key) { e = i; break; }
      *d = (uint8_t)e;
      return;
    }
    uint8_t *d = p.dst + (size_t)(y - p.row0) * p.dst_bpr + (size_t)x * rc.newnc * rc.dst_bps;
    if (clear && rc.key_on) {
      for (int k = 0; k < rc.key_len; k++) d[k] = rc.key_bytes[k];
      return;
    }
    for (int k = 0; k < rc.newnc; k++) {
      const uint8_t *s = px + rc.cmap[k] * rc.src_bps;
      if (rc.dst_bps == 2) { d[2 * k] = s[0]; d[2 * k + 1] = s[1]; }
      else d[k] = s[0];                                         // (8-bit source, or the high byte of a 16-bit one)
    }
  }
};

// host side (optimize.cpp)
struct iw_context;
struct IwbOptOutcome {
  int changed;                     // 0: the image stays as processed
  int imgtype, bit_depth;
  size_t bpr;
  int has_colorkey;
  unsigned colorkey[3];
  int has_palette;
  OptRecipe recipe;
};
// the reference's decision logic (src/imagew-opt.c:1166-1314) on the merged scan; fills `out` and, when a palette is
// used, ctx's palette
int iwb_opt_decide(iw_context *ctx, const IwbOptScan &scan, IwbOptOutcome *out);
void iwb_opt_merge(IwbOptScan *into, const IwbOptScan &from);
