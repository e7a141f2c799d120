// stream2.cuh -- second-generation streaming kernels (8-bit in, 8-bit out, STANDARD edges, <= 8 open windows).
//
// Same arithmetic as fused.cuh (register slot accumulators, taps in list order, DMUL + DADD rounded separately;
// reference: src/imagew-resize.c:346-361, src/imagew-main.c:831-927, :929-1155) but every warp is an independent
// pipeline fed by the TMA engine:
//
//   * each warp owns a ring of NST stages in shared memory; one elected lane posts the byte count on the
//     stage's mbarrier and a few lanes issue cp.async.bulk (TMA, SASS UBLKCP) copies: G input rows of 128 bytes
//     + the G schedule rows of weights (vertical), or KP pixel columns x 32 rows of the FP32 intermediate as ONE
//     contiguous block + KP schedule rows (horizontal).  No thread ever issues a per-sample load, there is no
//     __syncthreads in the steady state, and shared memory does not grow with the strip / segment length;
//   * the FP32 intermediate is stored "blocked transposed": inter[row group of 32][column][32 rows], so the
//     horizontal kernel (lanes = 32 rows, schedule warp-uniform) gets a column range of its row group with one
//     contiguous bulk copy and reads it with conflict-free LDS.32 at immediate offsets, while the vertical
//     kernel still writes whole 32-byte sectors (8 finished rows of a column);
//   * the 8-bit epilogue is branch-free in the common case: clamp by saturation, bucket -> level by one u16 and
//     one float4 read, fraction by multiplication with the stored reciprocal width, decision against the
//     ordered / random / 0.5 threshold unless it falls inside the FP32 error margin (then the exact FP64
//     quantiser of kernels.cuh decides, so codes stay bit-identical).
#pragma once

#if !defined(IWB_EMU)

#ifndef IWB2_V_WARPS
#define IWB2_V_WARPS 6
#endif
#ifndef IWB2_V_MINB
#define IWB2_V_MINB 2            // resident CTAs per SM the register allocation aims at
#endif
#ifndef IWB2_V_G
#define IWB2_V_G 8               // input rows per stage
#endif
#ifndef IWB2_V_NST
#define IWB2_V_NST 4
#endif
#ifndef IWB2_H_WARPS
#define IWB2_H_WARPS 4
#endif
#ifndef IWB2_H_KP
#define IWB2_H_KP 4              // pixel columns per stage
#endif
#ifndef IWB2_H_NST
#define IWB2_H_NST 4
#endif
#define IWB2_OST_PITCH 132       // bytes per output staging row (33 words: lane-private banks)
#define IWB2_NHINT 4096

struct Sched2 {                  // device pointers
  const double *w;               // [n_in + pad][SS], SS = S rounded up to even; idle slots and padding rows = +0.0
  const int *first, *last;       // [n_out] window of each output
  int n_in, n_out;
};

// Decision tables of one output channel, one 16-byte aligned blob (copied to shared memory by one bulk copy):
//   float4   e4[258]              (lo_q, 1/(hi_q - lo_q), lo_{q+1}, 1/(hi_{q+1} - lo_{q+1})) as floats
//   uint16_t hint[IWB2_NHINT + 8] lowest level intersecting bucket i (the level is hint or hint + 1)
//   double   thr[258]             exact lower bounds (the FP64 decision for samples inside the FP32 margin)
#define IWB2_TAB_HINT (258 * 16)
#define IWB2_TAB_THR (IWB2_TAB_HINT + (IWB2_NHINT + 8) * 2)
#define IWB2_TAB_BYTES ((IWB2_TAB_THR + 258 * 8 + 127) & ~127)
struct Quant2 {
  const unsigned char *blob;     // [IWB2_TAB_BYTES]
  float margin;
  int top;
};

struct V2Params {
  const uint8_t *src;            // byte (row r, column c) of the cropped input is src[r * bpr + c]; 16-byte aligned
  int tm_row0, tm_rows;          // input rows [tm_row0, tm_row0 + tm_rows) are in memory for this launch (TMA tensor extent)
  size_t bpr;                    // multiple of 16
  int ncols;                     // byte columns (= in_w * channels)
  const double *lut;             // [256] code -> linear
  Sched2 sc;
  float *inter;                  // blocked transposed intermediate, rows relative to row_lo
  int ncols_pad;                 // columns per row group in `inter`
  int row_lo, row_hi;            // output rows of this launch
  int strip;                     // output rows per CTA row (multiple of 32)
};

struct H2Params {
  const float *inter;
  int ncols_pad;
  int row_lo, row_hi;
  Sched2 sc;                     // in pixels
  int seg;                       // output pixels per CTA (multiple of 4)
  int out_w;
  DevOutput out;
  DevChanOut ch[4];
  const DevChanOut *ch_g;        // same, in global memory (exact path)
  Quant2 q[4];
  int same_tables;               // every channel uses q[0]'s tables
  const iwb_affine_t *jump;
};

static inline size_t v2_smem_bytes(int S, int nout) {
  const int SS = (S + 1) & ~1;
  const size_t stage = (size_t)IWB2_V_G * 128 + (size_t)IWB2_V_G * SS * 8;
  const size_t warp_bytes = (IWB2_V_NST * stage + 8 * 4 * 32 * 4 + IWB2_V_NST * 8 + 127) & ~(size_t)127;
  return 256 * 16 * 8 + IWB2_V_WARPS * warp_bytes + (size_t)nout * 4 + 16;
}

static inline size_t h2_smem_bytes(int S, int nch, int ntab, int nout) {
  const int SS = (S + 1) & ~1;
  const size_t stage = (size_t)IWB2_H_KP * nch * 128 + (size_t)IWB2_H_KP * SS * 8;
  const size_t warp_bytes = (IWB2_H_NST * stage + 32 * IWB2_OST_PITCH + IWB2_H_NST * 8 + 127) & ~(size_t)127;
  return IWB2_H_WARPS * warp_bytes + (size_t)ntab * IWB2_TAB_BYTES + 64 * 4 + 16 + (size_t)nout * 4 + 16;
}

// launchers (stream2.cu); false = no kernel variant for this shape
bool iwb2_tensor_maps_available();
int iwb2_sm_count();
int iwb2_v_ctas_per_sm(int S, size_t smem);
int iwb2_h_ctas_per_sm(int S, int nch, int epi, size_t smem);
bool iwb2_launch_v(int S, const V2Params &p, dim3 grid, size_t smem, cudaStream_t st);
bool iwb2_launch_h(int S, int nch, int epi, const H2Params &p, dim3 grid, size_t smem, cudaStream_t st);

#endif  // !IWB_EMU
