// gather5.cu -- "tile" kernels for images with alpha whose filter windows are longer than the register windows of
// gather4 (more than 8 taps: any reduction beyond 2x with a cubic filter, Lanczos3 at any reduction -- thumbnails of RGBA
// images).  Same data flow and the same schedules as gather4 (premultiply in the vertical pass, reference:
// src/imagew-main.c:880-885; un-premultiply / background / quantise in the horizontal epilogue, :1057-1061, :1091-1122;
// every output = sum of window[t] * w[o][t] in ascending source order, one DMUL + one DADD per tap, absent taps carry the
// weight 0.0 and are skipped, which leaves the sum bit-identical: src/imagew-resize.c:346-361), but the window is read
// from a shared-memory tile instead of a register ring, so its length is a run-time value (up to 64 here):
//
//   vertical:   warp = 128 adjacent sample columns (lane = 4 of them = whole pixels), CTA row = a strip of output rows
//               whose input rows fit a per-warp tile; one TMA transaction brings the tile; per output row and tap: codes
//               from the tile -> linear (shared-memory table for 8-bit input) -> colour x alpha -> multiply-add;
//   horizontal: warp = 32 ADJACENT OUTPUT PIXELS of one row (lane = pixel, so the stores of a row are coalesced without
//               staging), walking down a block of rows; the lane's W weights sit in shared memory (transposed, loaded
//               once per warp), the row segment of the FP32 intermediate arrives by cp.async, double-buffered; epilogue
//               per pixel with the exact generic compositing + quantiser of kernels.cuh (any depth, colour count, no /
//               ordered dither; random dither stays on the generic kernels: its draws are sequential along a row).
//
// Every sample costs a few times more instructions than in gather4's streaming kernels; against the generic
// one-thread-per-sample kernels these jobs ran on before, it is the difference between reading operands from shared
// memory and from L2.
#include <algorithm>
#include <cmath>

#include "gather4_dev.cuh"
#include "gather5.cuh"
#include "stream2_host.cuh"

#if !defined(IWB_EMU)

#define IWT_BOX_ROWS 32          // vertical: rows per TMA box
#define IWT_WARPS 4

template <bool D16, int NCHP>
__global__ void __launch_bounds__(32 * IWT_WARPS) iwt_v(const __grid_constant__ CUtensorMap tmap, const GVParams p, const int W, const int tile_rows) {
  constexpr int TW = D16 ? 256 : 128;
  constexpr int LUTB = D16 ? 0 : 2 * 256 * 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o0 = p.row_lo + blockIdx.y * p.strip, o1 = min(o0 + p.strip, p.row_hi);
  double *lut_c_s = (double *)smem_raw, *lut_a_s = lut_c_s + 256;      // (NCHP = 0: an image without alpha, one table)
  unsigned char *tile = smem_raw + LUTB + (size_t)warp * tile_rows * TW;
  const unsigned tile0 = s2_addr(tile), bar = s2_addr(smem_raw + LUTB + (size_t)IWT_WARPS * tile_rows * TW) + 8 * warp;
  const int wcol0 = (blockIdx.x * IWT_WARPS + warp) * 128;
  const bool active = wcol0 < p.ncols;
  const int col0 = wcol0 + lane * 4;
  const bool full = col0 + 4 <= p.ncols;
  const int r0 = p.sc.first[o0];
  const int nr = p.sc.first[o1 - 1] + W - r0;                   // input rows the strip's windows span (<= tile_rows)
  if (active) {
    if (lane == 0) {
      s2_mbar_init(bar, 1);
      s2_fence_async();
      const int nbox = (nr + IWT_BOX_ROWS - 1) / IWT_BOX_ROWS;
      s2_mbar_expect(bar, nbox * IWT_BOX_ROWS * TW);
      for (int k = 0; k < nbox; k++)                            // rows / columns outside the tensor arrive as zeros
        s2_tma_2d(tile0 + k * IWT_BOX_ROWS * TW, &tmap, wcol0 * (TW / 128), r0 + k * IWT_BOX_ROWS - p.tm_row0, bar);
    }
    __syncwarp();
  }
  if (!D16)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { lut_c_s[i] = p.lut_c[i]; if (NCHP > 0) lut_a_s[i] = p.lut_a[i]; }
  __syncthreads();
  if (!active) return;
  s2_mbar_wait(bar, 0);

  float *outp = p.inter + (ptrdiff_t)(o0 - p.row_base) * p.pitch + col0;
  const unsigned lane_off = tile0 + lane * (TW / 32);
  for (int o = o0; o < o1; o++) {
    const int b = __ldg(p.sc.first + o) - r0;
    const double *wp = p.sc.w + (size_t)o * W;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    // the weight and the codes of tap t + 1 are fetched while tap t is computed (the row after the strip's last one is
    // read from the tile's padding and never used)
    double wn = __ldg(wp);
    unsigned an, cn = 0;
    if (D16) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(an), "=r"(cn) : "r"(lane_off + b * TW));
    else an = s2_lds32(lane_off + b * TW);
    for (int t = 0; t < W; t++) {
      const double w = wn;                                      // (warp-uniform)
      const unsigned a = an, c2 = cn;
      wn = __ldg(wp + min(t + 1, W - 1));
      if (D16) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(an), "=r"(cn) : "r"(lane_off + (b + t + 1) * TW));
      else an = s2_lds32(lane_off + (b + t + 1) * TW);
      if (w == 0.0) continue;                                   // an absent tap: adding v * 0.0 cannot change the sum
      unsigned code[4];
      if (D16) {
        code[0] = __byte_perm(a, 0, 0x4401); code[1] = __byte_perm(a, 0, 0x4423);   // big-endian 16-bit samples
        code[2] = __byte_perm(c2, 0, 0x4401); code[3] = __byte_perm(c2, 0, 0x4423);
      } else {
        code[0] = a & 255u; code[1] = (a >> 8) & 255u; code[2] = (a >> 16) & 255u; code[3] = a >> 24;
      }
      double v[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const bool is_alpha = NCHP > 0 && (c % (NCHP > 0 ? NCHP : 1)) == NCHP - 1;
        if (D16) v[c] = __ldg((is_alpha ? p.lut_a : p.lut_c) + code[c]);
        else v[c] = (is_alpha ? lut_a_s : lut_c_s)[code[c]];
      }
      // colour x opacity (reference: in_pix[j] *= tmp_alpha, src/imagew-main.c:880-885)
      if (NCHP == 4) {
        v[0] = __dmul_rn(v[0], v[3]); v[1] = __dmul_rn(v[1], v[3]); v[2] = __dmul_rn(v[2], v[3]);
      } else if (NCHP == 2) {
        v[0] = __dmul_rn(v[0], v[1]); v[2] = __dmul_rn(v[2], v[3]);
      }
#pragma unroll
      for (int c = 0; c < 4; c++) acc[c] = __dadd_rn(acc[c], __dmul_rn(v[c], w));
    }
    float f[4];
#pragma unroll
    for (int c = 0; c < 4; c++) f[c] = (float)acc[c];           // reference: src/imagew-main.c:905-912
    if (p.intclamp) {
#pragma unroll
      for (int c = 0; c < 4; c++) f[c] = g4_clamp01f(acc[c]);
    }
    if (full) {
      *reinterpret_cast<float4 *>(outp) = make_float4(f[0], f[1], f[2], f[3]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; c++) if (col0 + c < p.ncols) outp[c] = f[c];
    }
    outp += p.pitch;
  }
}

// One output sample of the horizontal kernel: the reference's steps after the resampling (compositing, then quantise +
// none / ordered dither), on the exact generic code of kernels.cuh; the channel description is read from global memory.
static __device__ __noinline__ unsigned iwt_finish_sample(const DevCommon *cm, const DevChanOut *og, double s, double a, int x, int y) {
  s = composite_a(*cm, *og, s, a, x, y);
  const DevChanOut &o = *og;
  if (o.use_nc && o.guess_cs >= 0) {
    // 8 bits without dither: the stored code is the number of nearest-colour midpoints <= s (src/imagew-main.c:638-666).
    // A single-precision estimate of it is checked against its two neighbouring midpoints -- the same comparisons the
    // reference's search ends with -- and anything else takes the search.
    const double c = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
    const float sf = (float)c;
    float e;
    if (o.guess_cs == 0) e = sf <= 0.0031308f ? 12.92f * sf : 1.055f * __powf(sf, 1.0f / 2.4f) - 0.055f;
    else if (o.guess_cs == 3) e = sf < 0.020f ? 4.5f * sf : 1.099f * __powf(sf, 0.45f) - 0.099f;
    else if (o.guess_cs == 2) e = __powf(sf, o.guess_gamma);
    else e = sf;
    int q = (int)(e * 255.0f + 0.5f);
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    const bool lo_ok = q == 0 || !(__ldg(o.nc_tbl + q - 1) > c);
    const bool hi_ok = q == 255 || __ldg(o.nc_tbl + q) > c;
    if (lo_ok && hi_ok) return (unsigned)q;
    return (unsigned)(0.5 + quantise(o, s, x, y, 0));
  }
  if (g4_table_is_plain(o) && o.dfamily != 3) {
    const G4Q r = g4_q_issue(o, s);
    return (unsigned)(0.5 + g4_q_decide(o, og, r, x, y, 0));
  }
  return (unsigned)(0.5 + quantise(o, s, x, y, 0));
}

template <int NCH, int NCO>
__global__ void __launch_bounds__(32 * IWT_WARPS) iwt_h(const GHParams p, const DevCommon *cm_g, const int W, const int rows_per_warp, const int max_nb) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg0 = (blockIdx.x * IWT_WARPS + warp) * 32;        // this warp's 32 adjacent output pixels
  if (seg0 >= p.out_w) return;
  const int seg1 = min(seg0 + 32, p.out_w);
  const int xo = seg0 + lane;
  const bool valid = xo < p.out_w;
  const int rowbuf_floats = (max_nb * NCH + 3) & ~3;
  const size_t warp_bytes = (size_t)W * 32 * 8 + (size_t)2 * rowbuf_floats * 4;
  unsigned char *wbase = smem_raw + warp * warp_bytes;
  double *wts = (double *)wbase;                                // [W][32]
  float *rowbuf = (float *)(wbase + (size_t)W * 32 * 8);        // [2][rowbuf_floats]
  const int b0 = __ldg(p.sc.first + seg0);
  const int nb = __ldg(p.sc.first + seg1 - 1) + W - b0;          // input pixels the segment's windows span (<= max_nb)
  const int mine = (valid ? __ldg(p.sc.first + xo) : b0) - b0;  // this lane's window start within the segment
  static_assert(NCH >= 1 && NCH <= 4, "channels");
  for (int t = 0; t < W; t++) wts[t * 32 + lane] = valid ? __ldg(p.sc.w + (size_t)xo * W + t) : 0.0;
  const int y0 = p.row_lo + blockIdx.y * rows_per_warp, y1 = min(y0 + rows_per_warp, p.row_hi);
  // the segment's pixels of intermediate row y -> rowbuf[slot], 8 bytes per copy (rows start on 16 bytes, segments of a
  // gray + alpha image on 8); pixels beyond the row's end only meet zero weights and are left as they are
  const int nfl = min(nb, p.ncols / NCH - b0) * NCH;            // floats to copy (an even number)
  auto fetch = [&](int y, int slot) {
    const float *src = p.inter + (size_t)(y - p.row_lo) * p.pitch + (size_t)b0 * NCH;
    const unsigned dst = s2_addr(rowbuf + (size_t)slot * rowbuf_floats);
    if (NCH % 2 == 0) {
      for (int i = lane * 2; i < nfl; i += 64)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + i * 4), "l"(src + i) : "memory");
    } else {                                                    // (no alpha: 1 or 3 floats per pixel, segments start on 4 bytes)
      for (int i = lane; i < nfl; i += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + i * 4), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (y0 < y1) fetch(y0, 0);
  __syncwarp();
  const int bps = p.out.depth == 16 ? 2 : 1, bpp = NCO * bps;
  for (int y = y0; y < y1; y++) {
    const int slot = (y - y0) & 1;
    if (y + 1 < y1) {
      fetch(y + 1, slot ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const unsigned px0 = s2_addr(rowbuf + (size_t)slot * rowbuf_floats) + (unsigned)mine * NCH * 4;
    double acc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) acc[c] = 0.0;
    for (int t = 0; t < W; t++) {
      const double w = wts[t * 32 + lane];
      float f[NCH];
      if (NCH == 4) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[NCH - 1]) : "r"(px0 + t * 16));
      else if (NCH == 2) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f[0]), "=f"(f[NCH - 1]) : "r"(px0 + t * 8));
      else {
#pragma unroll
        for (int c = 0; c < NCH; c++) f[c] = s2_ldsf(px0 + (t * NCH + c) * 4);
      }
      // (a lane whose weight is 0.0 here adds +-0 to a sum that is never -0: bit-identical to skipping the tap; what
      // the buffer holds beyond the row's end is therefore irrelevant as long as it is finite or the product is dropped)
      if (w != 0.0) {
#pragma unroll
        for (int c = 0; c < NCH; c++) acc[c] = __dadd_rn(acc[c], __dmul_rn((double)f[c], w));
      }
    }
    if (p.cm.intclamp) {
#pragma unroll
      for (int c = 0; c < NCH; c++) acc[c] = acc[c] < 0.0 ? 0.0 : (acc[c] > 1.0 ? 1.0 : acc[c]);
    }
    if (valid) {
      const double a = NCH % 2 == 0 ? (double)(float)acc[NCH - 1] : 0.0;   // the resized alpha plane is FP32 (:1057-1061); odd NCH: no alpha
      uint8_t *dst = p.out.pix + (size_t)(y - p.out.row_origin) * p.out.bpr + (size_t)xo * bpp;
#pragma unroll
      for (int k = 0; k < NCO; k++) {
        const unsigned code = iwt_finish_sample(cm_g, p.ch_g + k, acc[k], a, xo, y);
        if (bps == 2) { dst[2 * k] = (unsigned char)(code >> 8); dst[2 * k + 1] = (unsigned char)(code & 255u); }
        else dst[k] = (unsigned char)code;
      }
    }
    __syncwarp();                                               // the row buffer of this slot is refilled two rows on
  }
}

namespace {

bool t5_input_map(const GVParams &p, bool d16, CUtensorMap *tm) {
  const int bpc = d16 ? 2 : 1;
  iwb2_encode_fn enc = iwb2_tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)p.ncols * bpc, (cuuint64_t)p.tm_rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)p.bpr};
  const cuuint32_t box[2] = {(cuuint32_t)(128 * bpc), (cuuint32_t)IWT_BOX_ROWS};
  const cuuint32_t estr[2] = {1, 1};
  const void *base = p.src + (size_t)p.tm_row0 * p.bpr;
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool D16, int NCHP>
bool t5_run_v(const GVJob &job, const int *first_host, cudaStream_t st) {
  GVParams p = job.p;
  const int W = job.W;
  const int threads = 32 * IWT_WARPS;
  const int colblocks = (p.ncols + 128 * IWT_WARPS - 1) / (128 * IWT_WARPS);
  const int rows = p.row_hi - p.row_lo;
  const size_t tw = D16 ? 256 : 128;
  // strips: as long as the per-warp tile allows (96 rows + one of padding: 12 / 24 KB per warp, so that several CTAs
  // share an SM -- the taps of an output are one dependent chain per channel and want many warps), and short enough for
  // about two waves
  const int max_tile_rows = 96;
  if (W + 1 > max_tile_rows) return false;
  auto span = [&](int o0, int o1) { return first_host[o1 - 1] + W - first_host[o0]; };
  int strip = std::max(1, std::min(rows, std::max(1, rows / std::max(1, 2 * iwb2_sm_count() * 2 / colblocks))));
  for (;;) {                                                    // shrink until every strip's rows fit the tile
    bool ok = true;
    for (int o0 = p.row_lo; o0 < p.row_hi && ok; o0 += strip) ok = span(o0, std::min(o0 + strip, p.row_hi)) <= max_tile_rows;
    if (ok) break;
    if (strip == 1) return false;
    strip = std::max(1, strip * 3 / 4);
  }
  int need = 0;
  for (int o0 = p.row_lo; o0 < p.row_hi; o0 += strip) need = std::max(need, span(o0, std::min(o0 + strip, p.row_hi)));
  const int tile_rows = (need + 1 + IWT_BOX_ROWS - 1) / IWT_BOX_ROWS * IWT_BOX_ROWS;   // (+1: the row the reader runs ahead)
  p.strip = strip;
  const size_t smem = (D16 ? 0 : 2 * 256 * 8) + (size_t)IWT_WARPS * tile_rows * tw + 8 * IWT_WARPS + 16;
  CUtensorMap tm;
  if (!t5_input_map(p, D16, &tm)) return false;
  auto kernel = iwt_v<D16, NCHP>;
  iwb2_ctas_per_sm(kernel, threads, smem);
  kernel<<<dim3((unsigned)colblocks, (unsigned)((rows + strip - 1) / strip)), threads, smem, st>>>(tm, p, W, tile_rows);
  return true;
}

template <int NCH, int NCO>
bool t5_run_h(const GHJob &job, const DevCommon *cm_g, int max_nb, cudaStream_t st) {
  const GHParams &p = job.p;
  const int W = job.W;
  const int threads = 32 * IWT_WARPS;
  const int segblocks = (p.out_w + 32 * IWT_WARPS - 1) / (32 * IWT_WARPS);
  const int rows = p.row_hi - p.row_lo;
  const int rowbuf_floats = (max_nb * NCH + 3) & ~3;
  const size_t smem = (size_t)IWT_WARPS * ((size_t)W * 32 * 8 + (size_t)2 * rowbuf_floats * 4) + 16;
  if (smem > (size_t)200 * 1024) return false;
  auto kernel = iwt_h<NCH, NCO>;
  const int slots = iwb2_ctas_per_sm(kernel, threads, smem) * iwb2_sm_count();
  // rows per warp: short enough for about three waves of CTAs; at least 8 rows to amortise the weights, unless the
  // image is so small that even 2-row blocks do not fill the device
  int rpw = std::max(8, (int)std::ceil((double)rows * segblocks / (3.0 * slots)));
  if ((rows + rpw - 1) / rpw * segblocks < slots) rpw = std::max(2, (int)std::ceil((double)rows * segblocks / (double)slots));
  rpw = std::min(rpw, std::max(1, rows));
  kernel<<<dim3((unsigned)segblocks, (unsigned)((rows + rpw - 1) / rpw)), threads, smem, st>>>(p, cm_g, W, rpw, max_nb);
  return true;
}

}  // namespace

bool iwt_supported(int WV, int WH, int nch) {
  return WV >= 2 && WH >= 2 && WV <= IWT_MAXW && WH <= IWT_MAXW && nch >= 1 && nch <= 4 && iwb2_tensor_map_encoder() != nullptr;
}
bool iwt_run_v(const GVJob &job, const int *first_host, cudaStream_t st) {
  if (job.nchp == 0) return job.p.depth16 ? t5_run_v<true, 0>(job, first_host, st) : t5_run_v<false, 0>(job, first_host, st);
  if (job.p.depth16) return job.nchp == 4 ? t5_run_v<true, 4>(job, first_host, st) : t5_run_v<true, 2>(job, first_host, st);
  return job.nchp == 4 ? t5_run_v<false, 4>(job, first_host, st) : t5_run_v<false, 2>(job, first_host, st);
}
bool iwt_run_h(const GHJob &job, const DevCommon *cm_g, int max_nb, cudaStream_t st) {
  if (job.nch == 4) {
    if (job.nco == 4) return t5_run_h<4, 4>(job, cm_g, max_nb, st);
    if (job.nco == 3) return t5_run_h<4, 3>(job, cm_g, max_nb, st);
    return false;
  }
  if (job.nch == 3) return job.nco == 3 ? t5_run_h<3, 3>(job, cm_g, max_nb, st) : false;
  if (job.nch == 1) return job.nco == 1 ? t5_run_h<1, 1>(job, cm_g, max_nb, st) : false;
  if (job.nco == 2) return t5_run_h<2, 2>(job, cm_g, max_nb, st);
  if (job.nco == 1) return t5_run_h<2, 1>(job, cm_g, max_nb, st);
  return false;
}

#endif  // !IWB_EMU
