// errdiff.cuh -- error-diffusion dither (reference: iw_errdiff_dither + put_sample_convert_from_linear,
// src/imagew-main.c:497-567, :690-765; row rotation :1124-1136).
//
// The algorithm is one dependent chain per channel: rows in order, serpentine, every sample's rounding error feeds its
// neighbours.  Only channels (and images) are parallel, so the job is to make the chain short:
//
//   * one warp per channel; lane 0 walks the chain, the other lanes move data: the row of linear samples is prefetched
//     (cp.async) into shared memory one row ahead, the finished row of codes leaves as coalesced stores, the error row
//     that becomes "two rows below" is cleared -- the three error rows live in shared memory and rotate by index;
//   * the error that reaches the NEXT sample of the current row never touches memory: the two cells ahead on the current
//     row are registers (a1, a2), so the chain per sample is
//         s = v + a0;  clamp;  locate s among 4 thresholds already in registers;  d_floor, d_ceil;  err;  a1 += err * m0
//     i.e. about ten dependent FP64 operations;
//   * the level search is taken off the chain by speculation: before err(x) is known, the next sample's value is known up
//     to err(x) * m0, which is less than half a level wide, so its level estimate and the four thresholds around it are
//     fetched while the chain of sample x is still in flight (one step of look-ahead inside the one thread); a miss
//     (estimate off by more than one level) falls back to a full search of the table;
//   * updates with a zero coefficient are skipped: the reference performs `cell += err * 0.0`, which cannot change a cell
//     (a cell is never -0.0: cells start at +0.0 and a sum of two doubles is -0.0 only if both are).
//
// The order of the additions into every cell is the reference's: contributions of earlier rows first (they are in the
// cell when the row starts), then err(x - 2 fwd) * m1, then err(x - fwd) * m0.
#pragma once
#if !defined(IWB_EMU)
#include <cuda_runtime.h>

#include "kernels.cuh"

struct ErrDiffFastParams {
  DevOutput out;
  int nch;
  DevChanOut ch[4];
  const DevChanOut *ch_g;      // the same in global memory (out-of-line slow path)
  const double *lin_plane[4];  // [rows][W]
  double *err[4];              // 3 * W doubles per channel, persistent across chunks of rows
  double m[4][12];
  int row0, row1;
};

// rows of doubles in shared memory: 3 error rows (2 cells of padding on either side, so that updates at the image border
// need no bounds checks: out-of-range targets are cells nobody reads), 2 sample rows, 1 row of codes (16-bit)
// + the level thresholds of a table of up to 256 levels (8-bit output)
static inline size_t iwb_errdiff_smem(int W) { return (size_t)3 * (W + 4) * 8 + (size_t)2 * W * 8 + (size_t)(260 * 8) + (size_t)W * 2 + 64; }

static __device__ __noinline__ Candidates errdiff_full_search(const DevChanOut *og, double s) {
  return og->qmode == 0 ? candidates_linear(*og, s) : candidates_table(*og, s);
}

// FS: the Floyd-Steinberg matrix (x 7/16 | 3/16 5/16 1/16): four coefficients, the row two below is never touched, and
// the three cells of the row below that a sample feeds are registers too (b_prev, b_cur, b_next: a cell is written to
// shared memory once, when its third and last contribution is in), so the chain's loop has no read-after-write through
// memory at all.
template <bool FS>
__global__ void __launch_bounds__(32) iwb_errdiff_warp(const ErrDiffFastParams p) {
  extern __shared__ __align__(16) unsigned char ed_smem[];
  const int c = blockIdx.x, lane = threadIdx.x;
  const int W = p.out.width;
  const DevChanOut &o = p.ch[c];
  double *E = (double *)ed_smem;                                // [3][W + 4]
  double *S = E + 3 * (W + 4);                                  // [2][W]
  double *T = S + 2 * W;                                        // [260] thresholds of a table of <= 256 levels
  unsigned short *C = (unsigned short *)(T + 260);              // [W]
  const int EP = W + 4;
  double *gerr = p.err[c];
  for (int k = 0; k < 3; k++) {
    for (int i = lane; i < W; i += 32) E[k * EP + 2 + i] = gerr[(size_t)k * W + i];
    if (lane < 2) { E[k * EP + lane] = 0.0; E[k * EP + W + 2 + lane] = 0.0; }
  }
  const double *m = p.m[c];
  const double m0 = m[0], m1 = m[1];
  bool nz[12];
#pragma unroll
  for (int k = 0; k < 12; k++) nz[k] = m[k] != 0.0;
  // plain full-depth table: level q is code q and its linear value is thr[q]
  const bool plain = o.qmode == 1 && !o.codeval && o.guess_cs >= 0 && o.ntab - 1 >= 8;
  const int top = o.ntab - 1;
  const float topf = (float)top;
  const bool small_tab = plain && top <= 255;
  if (small_tab) {
    for (int i = lane; i < 260; i += 32) T[i] = i <= top ? o.thr[i] : 2.0;
  }
  const double *thr = small_tab ? T : o.thr;
  const int bps = p.out.depth == 16 ? 2 : 1;

  auto prefetch_row = [&](int y, int buf) {                     // linear samples of row y -> S[buf] (all lanes)
    const double *src = p.lin_plane[c] + (size_t)(y - p.row0) * W;
    double *dst = S + (size_t)buf * W;
    for (int i = lane; i < W; i += 32) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src + i));
    }
    asm volatile("cp.async.commit_group;" ::);
  };
  auto estimate = [&](double s) {                               // FP32 estimate of the level of s, minus one, clamped
    const float sf = (float)s;
    float v;
    if (o.guess_cs == 0) v = sf <= 0.0031308f ? 12.92f * sf : 1.055f * __powf(sf, 1.0f / 2.4f) - 0.055f;
    else if (o.guess_cs == 3) v = sf < 0.020f ? 4.5f * sf : 1.099f * __powf(sf, 0.45f) - 0.099f;
    else if (o.guess_cs == 2) v = __powf(sf, o.guess_gamma);
    else v = sf;
    int b = (int)(v * topf) - 1;
    return b < 0 ? 0 : (b > top - 3 ? top - 3 : b);
  };

  if (p.row0 < p.row1) prefetch_row(p.row0, 0);
  int rb = 0;                                                   // error row of the current image row: E[rb], then rb+1, rb+2 (mod 3)
  for (int y = p.row0; y < p.row1; y++) {
    const int buf = (y - p.row0) & 1;
    if (y + 1 < p.row1) prefetch_row(y + 1, buf ^ 1);
    if (y + 1 < p.row1) asm volatile("cp.async.wait_group 1;" ::); else asm volatile("cp.async.wait_group 0;" ::);
    __syncwarp();
    if (FS && plain && lane == 0) {
      const double *src = S + (size_t)buf * W;
      double *e0 = E + ((rb + 0) % 3) * EP + 2, *e1 = E + ((rb + 1) % 3) * EP + 2;
      const int fwd = (y & 1) ? -1 : 1;
      int x = (y & 1) ? W - 1 : 0;
      const double m3 = m[3], m4 = m[4], m5 = m[5];
      double a0 = e0[x], a1 = e0[x + fwd];
      double b_prev = e1[x - fwd], b_cur = e1[x], b_next = e1[x + fwd];
      double v = src[x];
      v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
      double sp = v + a0;
      sp = sp > 1.0 ? 1.0 : (sp < 0.0 ? 0.0 : sp);
      int qb = estimate(sp);
      double t0 = thr[qb], t1 = thr[qb + 1], t2 = thr[qb + 2], t3 = thr[qb + 3];
      for (int z = 0; z < W; z++) {
        const int xn = x + fwd;
        // off the chain: next sample, its speculative thresholds, the cells that enter the register windows
        double vn = src[z + 1 < W ? xn : x];
        vn = vn < 0.0 ? 0.0 : (vn > 1.0 ? 1.0 : vn);
        double spn = vn + a1;
        spn = spn > 1.0 ? 1.0 : (spn < 0.0 ? 0.0 : spn);
        const int qbn = estimate(spn);
        const double u0 = thr[qbn], u1 = thr[qbn + 1], u2 = thr[qbn + 2], u3 = thr[qbn + 3];
        const double a2 = e0[x + 2 * fwd], bn2 = e1[x + 2 * fwd];
        // the chain
        double s = __dadd_rn(v, a0);
        s = s > 1.0 ? 1.0 : (s < 0.0 ? 0.0 : s);
        const bool b1 = t1 <= s, b2 = t2 <= s;
        int q = qb + (b1 ? 1 : 0) + (b2 ? 1 : 0);
        const double lo = b2 ? t2 : (b1 ? t1 : t0), hi = b2 ? t3 : (b1 ? t2 : t1);
        const double d_floor = __dsub_rn(s, lo), d_ceil = __dsub_rn(hi, s);
        const bool up = d_ceil <= d_floor;
        double err = up ? -d_ceil : d_floor;
        double code = (double)(q + (up ? 1 : 0));
        if (!(s <= 0.999999999 && t0 <= s && s < t3)) {         // rare: top of the range, or the estimate missed the level
          const Candidates k = errdiff_full_search(p.ch_g + c, s);
          if (k.exact) { code = k.code_lo; err = 0.0; }         // (no error is spread: adding err * m = +-0 changes nothing)
          else {
            const double df = __dsub_rn(s, k.lin_lo), dc = __dsub_rn(k.lin_hi, s);
            if (dc <= df) { err = -dc; code = k.code_hi; } else { err = df; code = k.code_lo; }
          }
        }
        C[x] = (unsigned short)(0.5 + code);
        a1 = __dadd_rn(a1, __dmul_rn(err, m0));                 // the next sample's error: the only memory-free hand-over
        b_prev = __dadd_rn(b_prev, __dmul_rn(err, m3));
        b_cur = __dadd_rn(b_cur, __dmul_rn(err, m4));
        b_next = __dadd_rn(b_next, __dmul_rn(err, m5));
        e1[x - fwd] = b_prev;                                   // third and last contribution is in
        a0 = a1; a1 = a2;
        b_prev = b_cur; b_cur = b_next; b_next = bn2;
        v = vn; qb = qbn; t0 = u0; t1 = u1; t2 = u2; t3 = u3;
        x = xn;
      }
      e1[x - fwd] = b_prev;                                     // (the last cell of the row: x is one past the end here)
    } else if (lane == 0) {
      const double *src = S + (size_t)buf * W;
      double *e0 = E + ((rb + 0) % 3) * EP + 2, *e1 = E + ((rb + 1) % 3) * EP + 2, *e2 = E + ((rb + 2) % 3) * EP + 2;
      const int fwd = (y & 1) ? -1 : 1;
      int x = (y & 1) ? W - 1 : 0;
      // a0: error that reaches sample x; a1, a2: the cells one and two ahead on this row
      double a0 = e0[x], a1 = e0[x + fwd], a2 = e0[x + 2 * fwd];   // (padding cells read as 0 and are never used)
      double v = src[x];
      v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
      int qb = 0;
      double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
      if (plain) {
        double sp = v + a0;
        sp = sp > 1.0 ? 1.0 : (sp < 0.0 ? 0.0 : sp);
        qb = estimate(sp);
        t0 = thr[qb]; t1 = thr[qb + 1]; t2 = thr[qb + 2]; t3 = thr[qb + 3];
      }
      for (int z = 0; z < W; z++) {
        // ---- off the chain: the next sample's value and the thresholds around its speculative level
        const int xn = x + fwd;
        double vn = 0.0;
        int qbn = 0;
        double u0 = 0, u1 = 0, u2 = 0, u3 = 0;
        if (z + 1 < W) {
          vn = src[xn];
          vn = vn < 0.0 ? 0.0 : (vn > 1.0 ? 1.0 : vn);
          if (plain) {
            double sp = vn + a1;                                // (without err(x) * m0: less than half a level away)
            sp = sp > 1.0 ? 1.0 : (sp < 0.0 ? 0.0 : sp);
            qbn = estimate(sp);
            u0 = thr[qbn]; u1 = thr[qbn + 1]; u2 = thr[qbn + 2]; u3 = thr[qbn + 3];
          }
        }
        const double a3 = e0[x + 3 * fwd > W + 1 ? W + 1 : (x + 3 * fwd < -2 ? -2 : x + 3 * fwd)];   // cell three ahead (earlier rows only)
        // ---- the chain (reference: src/imagew-main.c:703-749)
        double s = __dadd_rn(v, a0);
        if (s > 1.0) s = 1.0; else if (s < 0.0) s = 0.0;
        double code, err = 0.0;
        bool spread = false;
        if (plain && s <= 0.999999999 && t0 <= s && s < t3) {
          const bool b1 = t1 <= s, b2 = t2 <= s;
          const int q = qb + (b1 ? 1 : 0) + (b2 ? 1 : 0);
          const double lo = b2 ? t2 : (b1 ? t1 : t0), hi = b2 ? t3 : (b1 ? t2 : t1);
          const double d_floor = __dsub_rn(s, lo), d_ceil = __dsub_rn(hi, s);
          if (d_ceil <= d_floor) { err = -d_ceil; code = (double)(q + 1); }
          else { err = d_floor; code = (double)q; }
          spread = true;
        } else {
          const Candidates k = errdiff_full_search(p.ch_g + c, s);
          if (k.exact) code = k.code_lo;
          else {
            const double d_floor = __dsub_rn(s, k.lin_lo), d_ceil = __dsub_rn(k.lin_hi, s);
            if (d_ceil <= d_floor) { err = -d_ceil; code = k.code_hi; }
            else { err = d_floor; code = k.code_lo; }
            spread = true;
          }
        }
        C[x] = (unsigned short)(0.5 + code);
        if (spread) {
          // the two cells ahead on this row stay in registers; everything else is a read-modify-write in shared memory,
          // off the chain (src/imagew-main.c:545-566; zero coefficients skipped)
          if (nz[0]) a1 = __dadd_rn(a1, __dmul_rn(err, m0));
          if (nz[1]) a2 = __dadd_rn(a2, __dmul_rn(err, m1));
          if (nz[2]) e1[x - 2 * fwd] = __dadd_rn(e1[x - 2 * fwd], __dmul_rn(err, m[2]));
          if (nz[7]) e2[x - 2 * fwd] = __dadd_rn(e2[x - 2 * fwd], __dmul_rn(err, m[7]));
          if (nz[3]) e1[x - fwd] = __dadd_rn(e1[x - fwd], __dmul_rn(err, m[3]));
          if (nz[8]) e2[x - fwd] = __dadd_rn(e2[x - fwd], __dmul_rn(err, m[8]));
          if (nz[4]) e1[x] = __dadd_rn(e1[x], __dmul_rn(err, m[4]));
          if (nz[9]) e2[x] = __dadd_rn(e2[x], __dmul_rn(err, m[9]));
          if (nz[5]) e1[x + fwd] = __dadd_rn(e1[x + fwd], __dmul_rn(err, m[5]));
          if (nz[10]) e2[x + fwd] = __dadd_rn(e2[x + fwd], __dmul_rn(err, m[10]));
          if (nz[6]) e1[x + 2 * fwd] = __dadd_rn(e1[x + 2 * fwd], __dmul_rn(err, m[6]));
          if (nz[11]) e2[x + 2 * fwd] = __dadd_rn(e2[x + 2 * fwd], __dmul_rn(err, m[11]));
        }
        // next sample
        a0 = a1; a1 = a2; a2 = a3;
        v = vn; qb = qbn; t0 = u0; t1 = u1; t2 = u2; t3 = u3;
        x = xn;
      }
    }
    __syncwarp();
    // the row of codes leaves; the error row of this image row becomes the (empty) row two below
    {
      uint8_t *row = p.out.pix + (size_t)(y - p.out.row_origin) * p.out.bpr;
      const int nch = p.out.nch, oc = o.out_ch;
      if (bps == 1) for (int i = lane; i < W; i += 32) row[(size_t)nch * i + oc] = (uint8_t)C[i];
      else for (int i = lane; i < W; i += 32) { const unsigned cv = C[i]; row[((size_t)nch * i + oc) * 2] = (uint8_t)(cv >> 8); row[((size_t)nch * i + oc) * 2 + 1] = (uint8_t)(cv & 255u); }
      double *old0 = E + (rb % 3) * EP;
      for (int i = lane; i < EP; i += 32) old0[i] = 0.0;
      rb = (rb + 1) % 3;
    }
    __syncwarp();
  }
  // error rows back to global memory, in the order (this row, next, next + 1) the following chunk expects
  for (int k = 0; k < 3; k++)
    for (int i = lane; i < W; i += 32) gerr[(size_t)k * W + i] = E[((rb + k) % 3) * EP + 2 + i];
}
#endif  // !IWB_EMU
