// fused.cuh -- the fast path: two streaming kernels whose FP32 intermediate stays in L2.
//
// Why this shape.  Parity forbids FMA and fixes the tap order, so every tap costs one DMUL and
// one DADD and the FP64 pipe (64 lanes/SM) is the binding resource; shared memory moves only 128 B/clk/SM, i.e.
// one 8-byte operand per lane takes as long as the two FP64 instructions of a tap.  A kernel that reloads a
// sample or a weight per tap is therefore load-bound at 25-50 % of the FP64 roof.  These kernels load and convert
// every input sample ONCE and keep it in a register while it feeds all the output samples whose filter windows
// contain it:
//
//   * a thread walks its column (vertical pass) or its row (horizontal pass) in input order;
//   * output sample o accumulates in register slot o mod S, S = the largest number of windows open at once
//     (6 for a 2x Lanczos3 reduction, 16 for a 4x cubic enlargement), so slot indices are compile-time;
//   * the per-input "schedule" -- the S weights to apply at input r, idle slots = +0.0, and which outputs
//     complete at r -- is identical for all lanes of a warp, lives in shared memory and is read by broadcast;
//   * taps still reach each accumulator in ascending source order, one DMUL + one DADD each, so results are
//     bit-identical to the generic kernels.  (An idle slot adds v*0.0 = +-0 to an accumulator that is never -0,
//     which leaves it unchanged for the finite samples this path accepts.)
//
// Vertical kernel: lanes = adjacent byte columns (coalesced rows in, coalesced float rows out), colour->linear by
// a 16-way replicated shared-memory table (bank-conflict free).  Horizontal kernel: lanes = 32 different rows so
// that the schedule is warp-uniform; 32x(32*channels) blocks of the intermediate are transposed through shared
// memory; the finished sample goes straight into the quantise/dither epilogue and leaves through a byte staging
// tile as coalesced row segments.
#pragma once
#include <algorithm>
#include <vector>

struct iwb200_job;

struct AxisSched {            // device pointers
  int S, n_in, n_out;
  const double *w;            // [n_in][S]
  const int *emit_lo;         // [n_in] first output completing at input r
  const int *emit_n;          // [n_in] how many complete at r
  const int *first, *last;    // [n_out] window of each output
};

struct FastQuant {            // float fast path of the 8-bit epilogue for one output channel
  int mode;                   // 0: always the exact generic quantiser; 1: level + fraction test; 2: midpoint table
  const float2 *pairs;        // [levels] (lower bound, upper bound) of each level in linear space, as floats
  const uint16_t *hint;       // [nhint+1] level whose lower bound is <= i/nhint
  int nhint, top;             // top: highest level
  float margin;               // decision margin inside which the exact path is taken
};

struct VStreamParams {
  const uint8_t *src;         // input pixels; row `row_origin` of the stored image
  size_t bpr;
  int row_origin;
  int y0;                     // crop origin (rows)
  size_t x0_bytes;            // crop origin in bytes
  int ncols;                  // byte columns (= in_w * channels)
  const double *lut;          // [256] code -> linear
  AxisSched sc;
  float *interT;              // transposed intermediate: [ncols][pitch_j], row index relative to row_lo
  size_t pitch_j;             // floats per column (multiple of 32)
  int row_lo, row_hi;         // output rows (vertical outputs) of this launch
  int strip;                  // output rows per CTA (multiple of 8)
  int intclamp;
};

struct HStreamParams {
  const float *interT;
  size_t pitch_j;
  int row_lo, row_hi;
  AxisSched sc;
  int seg;                    // outputs per CTA segment
  DevOutput out;
  DevChanOut ch[4];           // exact epilogue description (generic structure)
  const DevChanOut *ch_g;     // the same four structs in global memory (addressable by the out-of-line exact path)
  FastQuant fq[4];
  int intclamp;
  int out_w;
  const iwb_affine_t *jump;
};

struct FastPlan {
  int enabled = 0;
  int nch = 0;
  int nco = 0;                   // output channels: nch, or 3 for gray input written as RGB (generation 2 only)
  int SV = 0, SH = 0, NC = 0;
  AxisSched sv{}, sh{};
  std::vector<int> vfirst, vlast, hfirst, hlast;   // host copies for band planning / uniform-range detection
  const double *lut = nullptr;
  const double *lut16 = nullptr;                   // generation 2: the same table, each entry repeated 16 times
  FastQuant fq[4]{};
  float *interT = nullptr;
  size_t pitch_j = 0;
  int strip = 32, seg = 64;
  int epi = 0;
  const DevChanOut *ch_g = nullptr;
  int chunk_rows = 0;
  size_t v_smem = 0, h_smem = 0;
  int v_nr_max = 0, h_nb_max = 0;
  // second-generation kernels (stream2.cuh); used when the call's buffers also meet the TMA alignment rules
  int v2 = 0;
  struct Sched2 *s2v = nullptr, *s2h = nullptr;   // (pointers: stream2.cuh is included after this header)
  struct Quant2 *q2 = nullptr;                    // [4]
  int same_tables = 0;
  int strip2 = 0, seg2 = 0;                        // 0: sized per launch to one wave of CTAs
  int uniform = 1;
  int depth16 = 0;                                 // 16-bit big-endian input (generation 2 only)
  int overlap = 1;                                 // sub-chunks pipelined over two streams (V of i+1 beside H of i)
  void *aux_stream = nullptr;
  std::vector<void *> ev;                          // [0..63] V done, [64..127] H done, [128] entry
  int ncols_pad = 0;
};

#if !defined(IWB_EMU)

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------------------
// vertical streaming kernel
//
// Control structure ("phases").  Outputs complete in order o, o+1, ... and output o lives in slot o mod S, so
// the slot that completes next cycles 0,1,..,S-1.  The code for one cycle is unrolled S times; phase k runs a
// tight loop over the input rows up to last[o] (every slot gets its tap: S*NC DMUL+DADD pairs per row, operands
// in registers), then stores and clears slot k -- a compile-time register.  No per-row emission tests, no
// accumulator shuffling.  Input bytes arrive through a per-thread cp.async ring RING_D rows deep.
// ---------------------------------------------------------------------------------------------------------
#define IWB_V_THREADS 128
#ifndef IWB_RING_D
#define IWB_RING_D 8
#endif
#ifndef IWB_V_TB
#define IWB_V_TB 8                // finished rows buffered per thread before a transposed store (8 = one 32-byte sector)
#endif
#ifndef IWB_V_PIPE
#define IWB_V_PIPE 0
#endif
#ifndef IWB_H_PIPE
#define IWB_H_PIPE 0
#endif
#ifndef IWB_LUT_REP
#define IWB_LUT_REP 16           // shared-memory copies of the colour table: one per lane of a half-warp (conflict free)
#endif
#ifndef IWB_V_MINBLOCKS
#define IWB_V_MINBLOCKS 4
#endif

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem_src));
}

// NC byte columns per thread (8 when rows are 8-byte aligned: the per-row overhead -- weights, ring, loop -- is
// then amortised over 8*S accumulator updates).  RING: input through the cp.async ring (needs NC in {4,8} and
// aligned rows); otherwise plain loads.  TB: finished rows buffered per thread before a transposed store.
template <int S, int NC, bool RING>
__global__ void __launch_bounds__(IWB_V_THREADS, (NC >= 8 ? 3 : IWB_V_MINBLOCKS)) iwb_vstream(const VStreamParams p) {
  constexpr int NW = NC >= 4 ? NC / 4 : 1;                // 32-bit ring words per thread and row
  constexpr int TB = IWB_V_TB;                            // rows per transposed store group
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int o0 = p.row_lo + blockIdx.y * p.strip;
  const int o1 = min(o0 + p.strip, p.row_hi);
  const int r0 = p.sc.first[o0], r1 = p.sc.last[o1 - 1];
  const int nr = r1 - r0 + 1;
  int ob = o0;                                            // first output still open at r0
  while (ob > 0 && p.sc.last[ob - 1] >= r0) ob--;
  const int nout = o1 - ob;
  double *lut_s = (double *)smem_raw;                     // [256][IWB_LUT_REP]
  double *w_s = lut_s + 256 * IWB_LUT_REP;                // [nr][S]
  int *last_s = (int *)(w_s + (size_t)nr * S);            // [nout] last input row of each output, relative to r0
  unsigned *ring = (unsigned *)(last_s + ((nout + 3) & ~3));          // [RING_D][threads][NW]
  float *tbuf = (float *)(ring + IWB_RING_D * IWB_V_THREADS * NW);    // [TB][NC][threads]
  for (int i = threadIdx.x; i < 256 * IWB_LUT_REP; i += blockDim.x) lut_s[i] = p.lut[i / IWB_LUT_REP];
  for (int i = threadIdx.x; i < nr * S; i += blockDim.x) w_s[i] = p.sc.w[(size_t)r0 * S + i];
  for (int i = threadIdx.x; i < nout; i += blockDim.x) last_s[i] = p.sc.last[ob + i] - r0;
  __syncthreads();

  const int col0 = (blockIdx.x * blockDim.x + threadIdx.x) * NC;
  if (col0 >= p.ncols) return;
  const bool full = col0 + NC <= p.ncols;
  const double *lut_t = lut_s + (threadIdx.x % IWB_LUT_REP);   // this lane's bank pair
  const uint8_t *colp = p.src + (size_t)(p.y0 + r0 - p.row_origin) * p.bpr + p.x0_bytes + col0;
  unsigned *myring = ring + threadIdx.x * NW;
  constexpr int RSTRIDE = IWB_V_THREADS * NW;             // words between ring slots

  double acc[S][NC];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < NC; c++) acc[s][c] = 0.0;

  const uint8_t *gnext = colp;                            // next row to request
  int requested = 0;
  auto request = [&]() {
    if (requested < nr) {
      if (NW == 2) cp_async8(myring + (requested % IWB_RING_D) * RSTRIDE, gnext);
      else cp_async4(myring + (requested % IWB_RING_D) * RSTRIDE, gnext);
      gnext += p.bpr;
    }
    requested++;
    cp_async_commit();
  };
  if (RING) {
#pragma unroll
    for (int d = 0; d < IWB_RING_D - 1; d++) request();
  }

  int i = 0;                                              // next input row (relative to r0)
  int fetched = 0;                                        // rows whose values have been looked up
  const double *wr = w_s;
  // look up the linear values of row `fetched` (software pipelined one row ahead of the arithmetic, so the
  // ring read -> unpack -> table read chain never sits in front of the FP64 block that needs it)
  auto fetch = [&](double (&v)[NC]) {
    unsigned code[NC];
    if (RING) {
      request();
      cp_async_wait<IWB_RING_D - 1>();
      const unsigned *slot = myring + (fetched % IWB_RING_D) * RSTRIDE;
#pragma unroll
      for (int wd = 0; wd < NW; wd++) {
        const unsigned x = slot[wd];
#pragma unroll
        for (int c = 0; c < 4; c++)
          if (wd * 4 + c < NC) code[wd * 4 + c] = (x >> (8 * c)) & 0xff;
      }
    } else {
      const uint8_t *q = colp + (size_t)min(fetched, nr - 1) * p.bpr;
#pragma unroll
      for (int c = 0; c < NC; c++) code[c] = (col0 + c < p.ncols) ? __ldg(q + c) : 0u;
    }
#pragma unroll
    for (int c = 0; c < NC; c++) v[c] = lut_t[code[c] * IWB_LUT_REP];
    fetched++;
  };
  double vcur[NC];
  if (IWB_V_PIPE) fetch(vcur);
  auto row_step = [&]() {
    double vnext[NC];
    if (IWB_V_PIPE) fetch(vnext);                         // row i+1 (a harmless re-read past the last row)
    else fetch(vcur);
#pragma unroll
    for (int s = 0; s < S; s++) {
      const double w = wr[s];
#pragma unroll
      for (int c = 0; c < NC; c++) acc[s][c] = __dadd_rn(acc[s][c], __dmul_rn(vcur[c], w));
    }
    if (IWB_V_PIPE) {
#pragma unroll
      for (int c = 0; c < NC; c++) vcur[c] = vnext[c];
    }
    wr += S;
    i++;
  };

  for (int og = ob - (ob % S); og < o1; og += S) {
#pragma unroll
    for (int k = 0; k < S; k++) {
      const int o = og + k;
      if (o >= ob && o < o1) {                             // warp-uniform
        const int iend = last_s[o - ob];
        while (i <= iend) row_step();
        if (o >= o0) {
          const int orel = o - p.row_lo;
          if (p.intclamp) {
#pragma unroll
            for (int c = 0; c < NC; c++) acc[k][c] = acc[k][c] < 0.0 ? 0.0 : (acc[k][c] > 1.0 ? 1.0 : acc[k][c]);
          }
#pragma unroll
          for (int c = 0; c < NC; c++) tbuf[((orel % TB) * NC + c) * IWB_V_THREADS + threadIdx.x] = (float)acc[k][c];
          if ((orel % TB) == TB - 1 || o == o1 - 1) {      // a full group: TB consecutive rows of each column
            const int base = orel - (orel % TB), cnt = (orel % TB) + 1;
#pragma unroll
            for (int c = 0; c < NC; c++) {
              if (full || col0 + c < p.ncols) {
                float *dst = p.interT + (size_t)(col0 + c) * p.pitch_j + base;
                float g[TB];
#pragma unroll
                for (int r = 0; r < TB; r++) g[r] = tbuf[(r * NC + c) * IWB_V_THREADS + threadIdx.x];
                if (cnt == TB) {
#pragma unroll
                  for (int r = 0; r < TB; r += 4) *reinterpret_cast<float4 *>(dst + r) = make_float4(g[r], g[r + 1], g[r + 2], g[r + 3]);
                } else {
#pragma unroll
                  for (int r = 0; r < TB; r++)
                    if (r < cnt) dst[r] = g[r];
                }
              }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < NC; c++) acc[k][c] = 0.0;
      }
    }
  }
  if (RING) cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------
// horizontal streaming kernel + quantise/dither epilogue
//
// Same phase structure; lanes = different rows so the schedule is warp-uniform, R rows per thread (a warp covers
// 32*R rows) so that one weight fetch and one loop iteration feed R*NCH*S accumulator updates.  Because the
// vertical kernel stores the intermediate transposed, lane j reads interT[column][row j]: 32 lanes = one
// 128-byte line, fetched by a per-thread cp.async ring IWB_RING_D columns deep (no staging, no barriers).
// A finished sample is parked in shared memory; after each cycle of S phases ONE copy of the epilogue code
// quantises the parked samples in increasing output order (random dither consumes its draws in that order) and
// puts the bytes into a staging tile that is flushed as coalesced row segments.
//
// Epilogue fast path: level search and dither decision in FP32 against float copies of the
// host tables; whenever the FP32 fraction is within `margin` of the decision threshold (a few 1e-4 of the
// samples) the exact FP64 quantiser of kernels.cuh decides instead, so stored codes are bit-identical.
// EPI = 1: every channel is in threshold-table mode with the same dither family; EPI = 2: every channel uses the
// nearest-colour midpoint table; EPI = 0: anything else, decided per sample at run time.
// ---------------------------------------------------------------------------------------------------------
#ifndef IWB_H_WARPS
#define IWB_H_WARPS 4
#endif
#define IWB_H_THREADS (32 * IWB_H_WARPS)
#define IWB_OSTAGE_PITCH 132   // bytes per staging row (33 words: lane-private banks)

__device__ __noinline__ int exact_quantise(const DevChanOut *o, double s, int x, int y, uint32_t draw) {
  return (int)(0.5 + quantise(*o, s, x, y, draw));
}

// FP32 level search: largest level whose float lower bound is <= sf, starting from the bucket hint
__device__ __forceinline__ int level_of(const float2 *pairs_s, const uint16_t *hint_s, float sf, int top, float2 *pr) {
  int q = hint_s[(int)(sf * 1024.0f)];
  float2 v = pairs_s[q];
  while (v.y <= sf && q < top) { q++; v = pairs_s[q]; }
  *pr = v;
  return q;
}

// EPI: 0 = per-sample run-time choice (mixed channels); 1 = threshold tables, no dither; 2 = nearest-colour
// midpoints; 3 = threshold tables, ordered dither; 4 = threshold tables, random dither.
template <int EPI>
__device__ __forceinline__ int quantise8_fast(const DevChanOut &o, const DevChanOut *og, const FastQuant &fq, const float2 *pairs_s,
                                              const uint16_t *hint_s, const float *ordrow, double s, int x, int y,
                                              uint32_t draw) {
  const int mode = EPI == 0 ? fq.mode : (EPI == 2 ? 2 : 1);
  if (mode == 0) return exact_quantise(og, s, x, y, draw);
  const int fam = EPI == 0 ? o.dfamily : (EPI == 3 ? 1 : (EPI == 4 ? 3 : 0));
  const float sf = (float)s;
  // s <= 0 clamps to 0 and converts to exactly code 0; s above the snap point converts to exactly the top code
  // (reference: src/imagew-main.c:684-685, :429-433, :618-621).  Without random dither the whole float
  // interval that rounds to 1.0f also ends at the top code (its fraction is > 0.9999 and every threshold is
  // <= 63.5/64), so the cheap float tests suffice; random dither keeps the exact FP64 tests.
  if (fam == 3) {
    if (!(s > 0.0)) return 0;
    if (s > 0.999999999) return fq.top;
  } else {
    if (!(sf > 0.0f)) return 0;
    if (sf >= 1.0f) return fq.top;
  }
  float2 pr;
  const int q = level_of(pairs_s, hint_s, sf, fq.top, &pr);
  if (mode == 2) {             // nearest colour: level q holds [midpoint below, midpoint above)
    if (fminf(sf - pr.x, pr.y - sf) > fq.margin) return q;
    return exact_quantise(og, s, x, y, draw);
  }
  if (q >= fq.top) return exact_quantise(og, s, x, y, draw);
  const float fr = __fdividef(sf - pr.x, pr.y - pr.x);
  float t;
  if (fam == 0) t = 0.5f;
  else if (fam == 1) t = ordrow[x & 7];
  else t = (float)draw * (1.0f / 4294967295.0f);
  const float d = fr - t;
  if (fabsf(d) > fq.margin) return q + (d >= 0.0f ? 1 : 0);
  return exact_quantise(og, s, x, y, draw);
}

template <int S, int NCH, int R, int EPI>
__global__ void __launch_bounds__(IWB_H_THREADS) iwb_hstream(const HStreamParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int xo0 = blockIdx.x * p.seg;
  const int xo1 = min(xo0 + p.seg, p.out_w);
  const int b0 = p.sc.first[xo0], b1 = p.sc.last[xo1 - 1];
  const int nb = b1 - b0 + 1;
  int xb = xo0;                                          // first output still open at b0
  while (xb > 0 && p.sc.last[xb - 1] >= b0) xb--;
  const int nout = xo1 - xb;
  constexpr int NV = NCH * R;                            // samples per thread and column
  double *w_s = (double *)smem_raw;                                        // [nb][S]
  constexpr int NPARK = S > 8 ? S * NV * 32 : 0;                           // parked samples only when not inlined
  double *scr_all = w_s + (size_t)nb * S;                                  // [warps][S][NV][32]
  double *scr = scr_all + (size_t)warp * NPARK;
  float2 *pairs_s = (float2 *)(scr_all + (size_t)IWB_H_WARPS * NPARK);     // [NCH][256]
  float *ordf_s = (float *)(pairs_s + NCH * 256);                          // [2][64]
  float *ring = ordf_s + 128;                                              // [RING_D][NV][threads]
  int *last_s = (int *)(ring + IWB_RING_D * NV * IWB_H_THREADS);           // [nout]
  uint16_t *hint_s = (uint16_t *)(last_s + nout);                          // [NCH][1026]
  unsigned char *ost_all = (unsigned char *)(hint_s + NCH * 1026 + ((NCH * 1026) & 1));
  unsigned char *ost = ost_all + (size_t)warp * 32 * R * IWB_OSTAGE_PITCH; // [32*R rows][pitch]

  for (int i = threadIdx.x; i < nb * S; i += blockDim.x) w_s[i] = p.sc.w[(size_t)b0 * S + i];
  for (int i = threadIdx.x; i < nout; i += blockDim.x) last_s[i] = p.sc.last[xb + i] - b0;
  for (int i = threadIdx.x; i < NCH * 256; i += blockDim.x) {
    const FastQuant &fq = p.fq[i >> 8];
    pairs_s[i] = (fq.mode && (i & 255) <= fq.top) ? fq.pairs[i & 255] : make_float2(3.0f, 4.0f);
  }
  for (int i = threadIdx.x; i < NCH * 1026; i += blockDim.x) {
    const FastQuant &fq = p.fq[i / 1026];
    hint_s[i] = (fq.mode && (i % 1026) <= fq.nhint) ? fq.hint[i % 1026] : (uint16_t)0;
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    const unsigned char k0[64] = {0, 48, 12, 60, 3, 51, 15, 63, 32, 16, 44, 28, 35, 19, 47, 31,
                                  8, 56, 4, 52, 11, 59, 7, 55, 40, 24, 36, 20, 43, 27, 39, 23,
                                  2, 50, 14, 62, 1, 49, 13, 61, 34, 18, 46, 30, 33, 17, 45, 29,
                                  10, 58, 6, 54, 9, 57, 5, 53, 42, 26, 38, 22, 41, 25, 37, 21};
    const unsigned char k1[64] = {3, 9, 17, 27, 25, 15, 7, 1, 11, 29, 37, 45, 43, 35, 23, 5,
                                  19, 39, 51, 57, 55, 49, 33, 13, 31, 47, 59, 63, 61, 53, 41, 21,
                                  30, 46, 58, 62, 60, 52, 40, 20, 18, 38, 50, 56, 54, 48, 32, 12,
                                  10, 28, 36, 44, 42, 34, 22, 4, 2, 8, 16, 26, 24, 14, 6, 0};
    ordf_s[i] = ((float)(i < 64 ? k0[i] : k1[i - 64]) + 0.5f) / 64.0f;   // exact in float
  }
  __syncthreads();

  const int jb = p.row_lo + (blockIdx.y * IWB_H_WARPS + warp) * 32 * R;  // first row of this warp
  if (jb >= p.row_hi) return;
  int yr[R];
  bool okr[R];
#pragma unroll
  for (int r = 0; r < R; r++) { yr[r] = jb + r * 32 + lane; okr[r] = yr[r] < p.row_hi; }

  double acc[S][NV];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int v = 0; v < NV; v++) acc[s][v] = 0.0;

  // random dither: this thread's draws are consecutive along each of its rows, one per output pixel and channel
  uint64_t rng[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) {
    const int c = v % NCH, r = v / NCH;
    rng[v] = 0;
    if (p.ch[c].dfamily == 3 && okr[r]) {
      uint64_t t = p.ch[c].prng_t1;
      const uint64_t n = (uint64_t)yr[r] * (uint64_t)p.out_w + (uint64_t)xo0;
      if (n) { if (t >= IWB_PRNG_P) t -= IWB_PRNG_P; t = iwb_prng_jump(t, n, p.jump); }
      rng[v] = t;
    }
  }
  const float *ordrow[R];
#pragma unroll
  for (int r = 0; r < R; r++) ordrow[r] = ordf_s + (p.ch[0].dsubtype == 1 ? 64 : 0) + (yr[r] & 7) * 8;

  // per-thread ring of intermediate samples: column b0+i, channel c, row r of this thread
  float *myring = ring + threadIdx.x;
  const float *gnext[R];
#pragma unroll
  for (int r = 0; r < R; r++) gnext[r] = p.interT + (size_t)b0 * NCH * p.pitch_j + (size_t)(yr[r] - p.row_lo);
#pragma unroll
  for (int r = 0; r < R; r++)
    if (!okr[r]) {
#pragma unroll
      for (int d = 0; d < IWB_RING_D; d++)
#pragma unroll
        for (int c = 0; c < NCH; c++) myring[(d * NV + r * NCH + c) * IWB_H_THREADS] = 0.0f;
    }
  int requested = 0;
  auto request = [&]() {
    if (requested < nb) {
      float *slot = myring + (requested % IWB_RING_D) * NV * IWB_H_THREADS;
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (okr[r]) {
#pragma unroll
          for (int c = 0; c < NCH; c++) cp_async4(slot + (r * NCH + c) * IWB_H_THREADS, gnext[r] + (size_t)c * p.pitch_j);
        }
        gnext[r] += (size_t)NCH * p.pitch_j;
      }
    }
    requested++;
    cp_async_commit();
  };
#pragma unroll
  for (int d = 0; d < IWB_RING_D - 1; d++) request();

  int staged_from = xo0;        // first output pixel currently held in the staging tile
  const int px_per_flush = 128 / NCH;
  auto flush = [&](int upto) {  // write staged pixels [staged_from, upto) of this warp's rows, coalesced per row
    const int nbytes = (upto - staged_from) * NCH;
    __syncwarp();
    for (int k = 0; k < 32 * R; k++) {
      const int yy = jb + k;
      if (yy >= p.row_hi) break;
      uint8_t *dst = p.out.pix + (size_t)(yy - p.out.row_origin) * p.out.bpr + (size_t)staged_from * NCH;
      for (int b = lane; b < nbytes; b += 32) dst[b] = ost[k * IWB_OSTAGE_PITCH + b];
    }
    __syncwarp();
    staged_from = upto;
  };

  int i = 0;                    // next input column (relative to b0)
  int fetched = 0;
  const double *wr = w_s;
  auto fetch = [&](double (&v)[NV]) {   // one column ahead of the arithmetic (see the vertical kernel)
    request();
    cp_async_wait<IWB_RING_D - 1>();
    const float *slot = myring + (fetched % IWB_RING_D) * NV * IWB_H_THREADS;
#pragma unroll
    for (int q = 0; q < NV; q++) v[q] = (double)slot[q * IWB_H_THREADS];
    fetched++;
  };
  double vcur[NV];
  if (IWB_H_PIPE) fetch(vcur);
  auto col_step = [&]() {
    double vnext[NV];
    if (IWB_H_PIPE) fetch(vnext);
    else fetch(vcur);
#pragma unroll
    for (int s = 0; s < S; s++) {
      const double w = wr[s];
#pragma unroll
      for (int q = 0; q < NV; q++) acc[s][q] = __dadd_rn(acc[s][q], __dmul_rn(vcur[q], w));
    }
    if (IWB_H_PIPE) {
#pragma unroll
      for (int q = 0; q < NV; q++) vcur[q] = vnext[q];
    }
    wr += S;
    i++;
  };

  // quantise one finished pixel (this thread's R rows x NCH channels) and stage its bytes
  unsigned char *ost_t = ost + lane * IWB_OSTAGE_PITCH;
  auto emit = [&](const double *a, int astride, int xo) {
    if (xo - staged_from >= px_per_flush) flush(xo);
    unsigned char *dstb = ost_t + (xo - staged_from) * NCH;
#pragma unroll
    for (int q = 0; q < NV; q++) {
      const int c = q % NCH, r = q / NCH;
      const DevChanOut &o = p.ch[c];
      uint32_t draw = 0;
      if (EPI == 4 || (EPI == 0 && o.dfamily == 3)) { draw = iwb_prng_draw_of(rng[q]); rng[q] = iwb_prng_step(rng[q]); }
      const float *orow = (EPI == 3) ? ordrow[r] : ordf_s + (o.dsubtype == 1 ? 64 : 0) + (yr[r] & 7) * 8;
      const int code = quantise8_fast<EPI>(o, p.ch_g + c, p.fq[c], pairs_s + c * 256, hint_s + c * 1026, orow,
                                           a[q * astride], xo, yr[r], draw);
      dstb[r * 32 * IWB_OSTAGE_PITCH + c] = (uint8_t)code;
    }
  };
  constexpr bool PARK = S > 8;   // many slots: one out-of-line copy of the epilogue instead of S inlined ones

  for (int og = xb - (xb % S); og < xo1; og += S) {
#pragma unroll
    for (int k = 0; k < S; k++) {
      const int o = og + k;
      if (o >= xb && o < xo1) {                            // warp-uniform
        const int iend = last_s[o - xb];
        while (i <= iend) col_step();
        if (PARK) {
#pragma unroll
          for (int q = 0; q < NV; q++) scr[(k * NV + q) * 32 + lane] = acc[k][q];
        } else if (o >= xo0) {
          emit(&acc[k][0], 1, o);
        }
#pragma unroll
        for (int q = 0; q < NV; q++) acc[k][q] = 0.0;
      }
    }
    if (PARK) {   // epilogue for the outputs of this cycle, in increasing order (own values only: no barrier)
#pragma unroll 1
      for (int k = 0; k < S; k++) {
        const int xo = og + k;
        if (xo < xo0 || xo >= xo1) continue;
        emit(scr + (size_t)k * NV * 32 + lane, 32, xo);
      }
    }
  }
  cp_async_wait<0>();
  flush(xo1);
}

#endif  // !IWB_EMU
