// gather4.cuh -- streaming kernels in "gather form" for images with an alpha channel: premultiply in the vertical pass,
// un-premultiply / background compositing / any output depth and dither in the horizontal epilogue.
//
// The generation-2 kernels (stream2.cuh) are in scatter form: every input sample is multiplied into the S outputs whose
// windows contain it, S x channels FP64 accumulators per thread.  That is the right shape for reductions, and it keeps
// the channels of a pixel apart.  Jobs with alpha want the opposite: all channels of a pixel in one thread (colour is
// multiplied by the pixel's alpha before resampling and divided by the resampled alpha after it; reference:
// src/imagew-main.c:880-885, :1057-1061, :1091-1110), and they are typically enlargements or mild reductions (few taps
// per output, many outputs per input) where S would be 12-16.  Here a thread keeps the last W input samples of each of
// its channels in registers (W = 2, 4, 6 or 8 >= the longest tap list) and evaluates every output as
//     acc = 0;  acc += window[t] * w[o][t]   for t = 0 .. W-1, ascending source order, one DMUL + one DADD per tap
// (reference: src/imagew-resize.c:346-361; absent taps carry the weight 0.0, which leaves the sum bit-identical, see
// pipeline.cu:build_sched).  The window is a register ring: the code is unrolled over the W positions of the window
// start modulo W, so every tap reads a compile-time register.
//
//   vertical:   warp = 128 adjacent sample columns (lane = 4 of them = one RGBA pixel or two gray+alpha pixels), strips
//               of output rows; input rows arrive through a per-warp TMA ring (2-D tiles of G rows x 128 columns);
//               codes -> linear through the colour table and the alpha table, colour x alpha, FP32 rows out, coalesced;
//   horizontal: warp = 32 rows x one segment of output pixels (lane = row, so the weights are warp-uniform); the FP32
//               intermediate arrives as 2-D TMA tiles of 32 rows x KP pixels; epilogue per pixel: resampled alpha
//               rounded to FP32, colour / alpha, background (solid or checkerboard, opaque or translucent), quantise +
//               dither with the exact quantiser of kernels.cuh (any depth, colour count, none / ordered / random
//               dither), bytes staged per warp and written as coalesced words.
#pragma once
#include <cuda_runtime.h>

#include "estimate.h"
#include "kernels.cuh"
#include "prng.cuh"

#if !defined(IWB_EMU)

#define IWG_MAXW 8               // longest window the kernels are built for
#define IWG_V_G 8                // vertical: input rows per pipeline stage
#define IWG_V_NST 4
#define IWG_H_KP 4               // horizontal: input pixels per pipeline stage
#define IWG_H_NST 4
#define IWG_WARPS 4
#define IWG_OST_PITCH 132        // bytes per output staging row (33 words: lane-private banks); <= 128 bytes staged

struct GSched {                  // device pointers; one dimension
  const int *first;              // [n_out] first source of each output's window
  const double *w;               // [n_out][W] weights of sources first[o] .. first[o] + W - 1 (0.0 where the list has none)
  int n_in, n_out;
};

struct GVParams {
  const uint8_t *src;            // byte (row r, column c) of the cropped input is src[r * bpr + c]; 16-byte aligned
  size_t bpr;                    // multiple of 16
  int tm_row0, tm_rows;          // input rows in memory for this launch (TMA tensor extent)
  int ncols;                     // sample columns (= in_w * channels)
  const double *lut_c, *lut_a;   // code -> linear for colour / alpha samples: 256 entries (8-bit) or 65536 (16-bit)
  int depth16;
  GSched sc;
  float *inter;                  // row-major FP32 intermediate, column c at float c
  int pitch;                     // floats per row (multiple of 4)
  int row_base, row_lo, row_hi;  // output row stored as row 0 of `inter`; output rows of this launch
  int strip;                     // output rows per CTA row
  int intclamp;
};

struct GHParams {
  const float *inter;
  int pitch, ncols;
  int row_lo, row_hi;            // output rows of this launch; row_lo is row 0 of `inter`
  GSched sc;                     // in pixels
  int out_w, seg;                // segment = output pixels per CTA column (multiple of 4)
  DevOutput out;
  DevCommon cm;
  DevChanOut ch[4];              // by interleaved (= physical input) channel
  const DevChanOut *ch_g;        // the same four in global memory (for out-of-line code)
  int nco;                       // output channels: output channel k is interleaved channel k (alpha last, kept or dropped)
  const iwb_affine_t *jump;
  int any_random;
  const IwbEstRec *est;          // level model (estimate.h) of the output channels' common table, or NULL
  double est_margin;
};

struct GVJob { GVParams p; int W, nchp; };
struct GHJob { GHParams p; int W, nch, nco; };

bool iwg_supported(int WV, int WH, int nch);
bool iwg_run_v(const GVJob &job, cudaStream_t st);
bool iwg_run_h(const GHJob &job, cudaStream_t st);

#endif  // !IWB_EMU
