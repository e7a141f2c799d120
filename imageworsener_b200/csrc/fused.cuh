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
  FastQuant fq[4];
  int intclamp;
  int out_w;
  const iwb_affine_t *jump;
};

struct FastPlan {
  int enabled = 0;
  int nch = 0;
  int SV = 0, SH = 0, NC = 0;
  AxisSched sv{}, sh{};
  std::vector<int> vfirst, vlast;   // host copies for band planning
  const double *lut = nullptr;
  FastQuant fq[4]{};
  float *interT = nullptr;
  size_t pitch_j = 0;
  int strip = 32, seg = 64;
  int chunk_rows = 0;
  size_t v_smem = 0, h_smem = 0;
  int v_nr_max = 0, h_nb_max = 0;
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
#define IWB_RING_D 8

template <int S, int NC, bool RING>
__global__ void __launch_bounds__(IWB_V_THREADS) iwb_vstream(const VStreamParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int o0 = p.row_lo + blockIdx.y * p.strip;
  const int o1 = min(o0 + p.strip, p.row_hi);
  const int r0 = p.sc.first[o0], r1 = p.sc.last[o1 - 1];
  const int nr = r1 - r0 + 1;
  int ob = o0;                                            // first output still open at r0
  while (ob > 0 && p.sc.last[ob - 1] >= r0) ob--;
  const int nout = o1 - ob;
  double *lut_s = (double *)smem_raw;                     // [256][16]
  double *w_s = lut_s + 256 * 16;                         // [nr][S]
  int *last_s = (int *)(w_s + (size_t)nr * S);            // [nout] last input row of each output, relative to r0
  unsigned *ring = (unsigned *)(last_s + ((nout + 3) & ~3));   // [RING_D][threads]
  float *tbuf = (float *)(ring + IWB_RING_D * IWB_V_THREADS);  // [8][NC][threads] finished rows awaiting a sector
  for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) lut_s[i] = p.lut[i >> 4];
  for (int i = threadIdx.x; i < nr * S; i += blockDim.x) w_s[i] = p.sc.w[(size_t)r0 * S + i];
  for (int i = threadIdx.x; i < nout; i += blockDim.x) last_s[i] = p.sc.last[ob + i] - r0;
  __syncthreads();

  const int col0 = (blockIdx.x * blockDim.x + threadIdx.x) * NC;
  if (col0 >= p.ncols) return;
  const bool full = col0 + NC <= p.ncols;
  const int rep = threadIdx.x & 15;
  const uint8_t *colp = p.src + (size_t)(p.y0 + r0 - p.row_origin) * p.bpr + p.x0_bytes + col0;
  unsigned *myring = ring + threadIdx.x;

  double acc[S][NC];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < NC; c++) acc[s][c] = 0.0;

  if (RING) {
#pragma unroll
    for (int d = 0; d < IWB_RING_D - 1; d++) {
      if (d < nr) cp_async4(myring + d * IWB_V_THREADS, colp + (size_t)d * p.bpr);
      cp_async_commit();
    }
  }

  int i = 0;                                              // next input row (relative to r0)
  auto row_step = [&]() {
    unsigned code[NC];
    if (RING) {
      const int ahead = i + IWB_RING_D - 1;
      if (ahead < nr) cp_async4(myring + (ahead % IWB_RING_D) * IWB_V_THREADS, colp + (size_t)ahead * p.bpr);
      cp_async_commit();
      cp_async_wait<IWB_RING_D - 1>();
      const unsigned v = myring[(i % IWB_RING_D) * IWB_V_THREADS];
#pragma unroll
      for (int c = 0; c < NC; c++) code[c] = (v >> (8 * c)) & 0xff;
    } else {
      const uint8_t *q = colp + (size_t)i * p.bpr;
#pragma unroll
      for (int c = 0; c < NC; c++) code[c] = (col0 + c < p.ncols) ? __ldg(q + c) : 0u;
    }
    double v[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) v[c] = lut_s[code[c] * 16 + rep];
    const double *wr = w_s + (size_t)i * S;
#pragma unroll
    for (int s = 0; s < S; s++) {
      const double w = wr[s];
#pragma unroll
      for (int c = 0; c < NC; c++) acc[s][c] = __dadd_rn(acc[s][c], __dmul_rn(v[c], w));
    }
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
          float f[NC];
#pragma unroll
          for (int c = 0; c < NC; c++) {
            double a = acc[k][c];
            if (p.intclamp) a = a < 0.0 ? 0.0 : (a > 1.0 ? 1.0 : a);
            f[c] = (float)a;
          }
#pragma unroll
          for (int c = 0; c < NC; c++) tbuf[((orel & 7) * NC + c) * IWB_V_THREADS + threadIdx.x] = f[c];
          if ((orel & 7) == 7 || o == o1 - 1) {            // a full 8-row group (one 32-byte sector per column)
            const int base = orel & ~7, cnt = (orel & 7) + 1;
#pragma unroll
            for (int c = 0; c < NC; c++) {
              if (col0 + c < p.ncols) {
                float *dst = p.interT + (size_t)(col0 + c) * p.pitch_j + base;
                float g[8];
#pragma unroll
                for (int r = 0; r < 8; r++) g[r] = tbuf[(r * NC + c) * IWB_V_THREADS + threadIdx.x];
                if (cnt == 8) {
                  reinterpret_cast<float4 *>(dst)[0] = make_float4(g[0], g[1], g[2], g[3]);
                  reinterpret_cast<float4 *>(dst)[1] = make_float4(g[4], g[5], g[6], g[7]);
                } else {
#pragma unroll
                  for (int r = 0; r < 8; r++)
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
// Same phase structure; lanes = 32 different rows so the schedule is warp-uniform.  Because the vertical kernel
// stores the intermediate transposed, lane j reads interT[column][row j]: 32 lanes = one 128-byte line, fetched
// by a per-thread cp.async ring IWB_RING_D columns deep (no staging, no barriers).  A finished sample is parked
// in shared memory; after each cycle of S phases ONE copy of the epilogue code quantises the parked samples in
// increasing output order (random dither consumes its draws in that order) and puts the bytes into a staging
// tile that is flushed as coalesced row segments.
//
// Epilogue fast path: level search and dither decision in FP32 against float copies of the
// host tables; whenever the FP32 fraction is within `margin` of the decision threshold (a few 1e-4 of the
// samples) the exact FP64 quantiser of kernels.cuh decides instead, so stored codes are bit-identical.
// ---------------------------------------------------------------------------------------------------------
#define IWB_H_WARPS 4
#define IWB_H_THREADS (32 * IWB_H_WARPS)
#define IWB_OSTAGE_PITCH 132   // bytes per staging row (33 words: lane-private banks)

__device__ __noinline__ double exact_quantise(const DevChanOut &o, double s, int x, int y, uint32_t draw) {
  return quantise(o, s, x, y, draw);
}

__device__ __forceinline__ int quantise8_fast(const DevChanOut &o, const FastQuant &fq, const float2 *pairs_s,
                                              const uint16_t *hint_s, const float *ordf_s, double s, int x, int y,
                                              uint32_t draw) {
  if (fq.mode == 0) return (int)(0.5 + exact_quantise(o, s, x, y, draw));
  // s <= 0 clamps to 0, which converts to exactly code 0; s above the snap point converts to exactly the top
  // code (reference: src/imagew-main.c:684-685, :429-433, :618-621) -- both "exact", no dither decision.
  if (!(s > 0.0)) return 0;
  if (s > 0.999999999) return fq.top;
  const float sf = (float)s;
  int q = hint_s[(int)(sf * (float)fq.nhint)];
  float2 pr = pairs_s[q];
  while (pr.y <= sf && q < fq.top) { q++; pr = pairs_s[q]; }
  if (fq.mode == 2) {          // nearest colour: level q holds [midpoint below, midpoint above)
    const float d = fminf(sf - pr.x, pr.y - sf);
    if (d > fq.margin) return q;
    return (int)(0.5 + exact_quantise(o, s, x, y, draw));
  }
  if (q >= fq.top) return (int)(0.5 + exact_quantise(o, s, x, y, draw));
  const float fr = __fdividef(sf - pr.x, pr.y - pr.x);
  float t;
  if (o.dfamily == 0) t = 0.5f;
  else if (o.dfamily == 1) t = ordf_s[(y & 7) * 8 + (x & 7)];
  else t = (float)draw * (1.0f / 4294967295.0f);
  if (fabsf(fr - t) > fq.margin) return q + (fr >= t ? 1 : 0);
  return (int)(0.5 + exact_quantise(o, s, x, y, draw));
}

template <int S, int NCH>
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
  double *w_s = (double *)smem_raw;                                        // [nb][S]
  double *scr_all = w_s + (size_t)nb * S;                                  // [warps][S][NCH][32] parked samples
  double *scr = scr_all + (size_t)warp * S * NCH * 32;
  float2 *pairs_s = (float2 *)(scr_all + (size_t)IWB_H_WARPS * S * NCH * 32);   // [NCH][256]
  float *ordf_s = (float *)(pairs_s + NCH * 256);                          // [2][64]
  float *ring = ordf_s + 128;                                              // [RING_D][NCH][threads]
  int *last_s = (int *)(ring + IWB_RING_D * NCH * IWB_H_THREADS);          // [nout]
  uint16_t *hint_s = (uint16_t *)(last_s + nout);                          // [NCH][1026]
  unsigned char *ost_all = (unsigned char *)(hint_s + NCH * 1026 + ((NCH * 1026) & 1));
  unsigned char *ost = ost_all + (size_t)warp * 32 * IWB_OSTAGE_PITCH;

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

  const int jb = p.row_lo + (blockIdx.y * IWB_H_WARPS + warp) * 32;   // first row of this warp
  if (jb >= p.row_hi) return;
  const int y = jb + lane;
  const bool row_ok = y < p.row_hi;

  double acc[S][NCH];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < NCH; c++) acc[s][c] = 0.0;

  // random dither: this thread's draws are consecutive along its row, one per output pixel and channel
  uint64_t rng[NCH];
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    rng[c] = 0;
    if (p.ch[c].dfamily == 3 && row_ok) {
      uint64_t t = p.ch[c].prng_t1;
      const uint64_t n = (uint64_t)y * (uint64_t)p.out_w + (uint64_t)xo0;
      if (n) { if (t >= IWB_PRNG_P) t -= IWB_PRNG_P; t = iwb_prng_jump(t, n, p.jump); }
      rng[c] = t;
    }
  }

  // per-thread ring of intermediate samples: column b0+i, channel c, this thread's row
  float *myring = ring + threadIdx.x;
  const float *colp = p.interT + (size_t)b0 * NCH * p.pitch_j + (size_t)(y - p.row_lo);
  if (!row_ok) {
#pragma unroll
    for (int d = 0; d < IWB_RING_D * NCH; d++) myring[d * IWB_H_THREADS] = 0.0f;
  }
  auto issue = [&](int col) {
    if (row_ok && col < nb) {
#pragma unroll
      for (int c = 0; c < NCH; c++)
        cp_async4(myring + ((col % IWB_RING_D) * NCH + c) * IWB_H_THREADS, colp + ((size_t)col * NCH + c) * p.pitch_j);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int d = 0; d < IWB_RING_D - 1; d++) issue(d);

  int staged_from = xo0;        // first output pixel currently held in the staging tile
  const int px_per_flush = 128 / NCH;
  auto flush = [&](int upto) {  // write staged pixels [staged_from, upto) of all 32 rows, coalesced per row
    const int nbytes = (upto - staged_from) * NCH;
    __syncwarp();
    for (int k = 0; k < 32; k++) {
      const int yy = jb + k;
      if (yy >= p.row_hi) break;
      uint8_t *dst = p.out.pix + (size_t)(yy - p.out.row_origin) * p.out.bpr + (size_t)staged_from * NCH;
      for (int b = lane; b < nbytes; b += 32) dst[b] = ost[k * IWB_OSTAGE_PITCH + b];
    }
    __syncwarp();
    staged_from = upto;
  };

  int i = 0;                    // next input column (relative to b0)
  auto col_step = [&]() {
    issue(i + IWB_RING_D - 1);
    cp_async_wait<IWB_RING_D - 1>();
    double v[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) v[c] = (double)myring[((i % IWB_RING_D) * NCH + c) * IWB_H_THREADS];
    const double *wr = w_s + (size_t)i * S;
#pragma unroll
    for (int s = 0; s < S; s++) {
      const double w = wr[s];
#pragma unroll
      for (int c = 0; c < NCH; c++) acc[s][c] = __dadd_rn(acc[s][c], __dmul_rn(v[c], w));
    }
    i++;
  };

  for (int og = xb - (xb % S); og < xo1; og += S) {
#pragma unroll
    for (int k = 0; k < S; k++) {
      const int o = og + k;
      if (o >= xb && o < xo1) {                            // warp-uniform
        const int iend = last_s[o - xb];
        while (i <= iend) col_step();
#pragma unroll
        for (int c = 0; c < NCH; c++) { scr[(k * NCH + c) * 32 + lane] = acc[k][c]; acc[k][c] = 0.0; }
      }
    }
    // epilogue for the outputs of this cycle, in increasing order (each thread reads back only its own values)
#pragma unroll 1
    for (int k = 0; k < S; k++) {
      const int xo = og + k;
      if (xo < xo0 || xo >= xo1) continue;
      if (xo - staged_from >= px_per_flush) flush(xo);
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        double a = scr[(k * NCH + c) * 32 + lane];
        if (p.intclamp) a = a < 0.0 ? 0.0 : (a > 1.0 ? 1.0 : a);
        const DevChanOut &o = p.ch[c];
        uint32_t draw = 0;
        if (o.dfamily == 3) { draw = iwb_prng_draw_of(rng[c]); rng[c] = iwb_prng_step(rng[c]); }
        const int code = quantise8_fast(o, p.fq[c], pairs_s + c * 256, hint_s + c * 1026, ordf_s + (o.dsubtype == 1 ? 64 : 0),
                                        a, xo, y, draw);
        ost[lane * IWB_OSTAGE_PITCH + (xo - staged_from) * NCH + o.out_ch] = (uint8_t)code;
      }
    }
  }
  cp_async_wait<0>();
  flush(xo1);
}

#endif  // !IWB_EMU
