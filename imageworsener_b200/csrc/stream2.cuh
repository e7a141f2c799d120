// stream2.cuh -- second-generation streaming kernels (8-bit in, 8-bit out, STANDARD edges, <= 8 open windows).
//
// Same arithmetic as fused.cuh (register slot accumulators, taps in list order, DMUL + DADD rounded separately;
// reference: src/imagew-resize.c:346-361, src/imagew-main.c:831-927, :929-1155) but every warp is an independent
// pipeline fed by the TMA engine:
//
//   * each warp owns a ring of stages in shared memory; one elected lane posts the byte count on the stage's
//     mbarrier and issues the copies (SASS UTMALDG / UBLKCP): a 2-D tensor tile of G input rows x 128 bytes + the G
//     schedule rows of weights (vertical), or a tile of KP pixel columns x 32 rows of the FP32 intermediate + KP
//     schedule rows (horizontal).  No thread ever issues a per-sample global load, there is
//     no __syncthreads in the steady state, and shared memory does not grow with the strip / segment length;
//   * the FP32 intermediate is plain row-major: the vertical kernel stores each finished row of a warp as one
//     coalesced 512-byte segment straight from registers; the horizontal kernel (lanes = 32 rows, so that the
//     schedule is warp-uniform) gets [32 rows][columns] tiles by 2-D TMA and reads its row 16 bytes at a time --
//     with a tile pitch of an odd number of 16-byte chunks that access is bank-conflict free, i.e. the
//     transposition between the two passes costs nothing;
//   * where the schedule is regular (integer reduction factors away from the border) a whole cycle of S outputs
//     is straight-line code (stream2_dev.cuh); elsewhere per-output trip counts come from the schedule tables;
//   * the 8-bit epilogue is branch-free in the common case: clamp by saturation, bucket -> level by one u16 and
//     one float4 read, fraction by multiplication with the stored reciprocal width, decision against the
//     ordered / random / 0.5 threshold unless it falls inside the FP32 error margin (then an exact FP64
//     decision on the same tables is taken, so codes stay bit-identical).
#pragma once

#if !defined(IWB_EMU)

#ifndef IWB2_V_WARPS
#define IWB2_V_WARPS 4
#endif
#ifndef IWB2_V_MINB
#define IWB2_V_MINB 4            // resident CTAs per SM the register allocation aims at (16 warps per SM)
#endif
#ifndef IWB2_V_G
#define IWB2_V_G 8               // general path: input rows per stage
#endif
#ifndef IWB2_V_NST
#define IWB2_V_NST 4
#endif
#ifndef IWB2_H_WARPS
#define IWB2_H_WARPS 4
#endif
#ifndef IWB2_H_MINB
#define IWB2_H_MINB 0            // resident CTAs per SM the horizontal kernel's register allocation aims at (0: no cap)
#endif
#ifndef IWB2_PARTS_FULL
#define IWB2_PARTS_FULL 1        // the uniform parts of a launch fill the wave by themselves (stream2_host.cuh)
#endif
#ifndef IWB2_H_SWAP
#define IWB2_H_SWAP 1            // horizontal launch order: row blocks fastest, so that the short border parts come last
#endif
#ifndef IWB2_H_KP
#define IWB2_H_KP 4              // general path: pixel columns per stage
#endif
#ifndef IWB2_H_NST
#define IWB2_H_NST 3
#endif
#ifndef IWB2_U_NST
#define IWB2_U_NST 2             // uniform path: stages (one cycle of S outputs each)
#endif
// Deferred adds (stream2_dev.cuh) make the FP64 instructions of one input independent of each other at the price
// of one more register pair per accumulator.  Measured on cfg2: the vertical kernel (4 columns x S slots per thread)
// is faster WITHOUT them at 16 warps per SM (0.117 ms) than with them at 12 (0.127 ms); the horizontal kernel, whose
// occupancy is bounded by shared memory, keeps them.
#ifndef IWB2_DEFER_V
#define IWB2_DEFER_V 0
#endif
#ifndef IWB2_DEFER_H
#define IWB2_DEFER_H 1
#endif
// Uniform path: the first product of an output is ASSIGNED to its accumulator instead of being added to a zeroed one
// (one DADD and the zeroing less per output and column; with deferred adds two DADDs less).  `+0.0 + x` and `x` differ
// only for x = -0.0, and an accumulator that starts as -0.0 where the reference has +0.0 meets it again at the first
// tap whose product is not -0.0: a tap with a positive weight (every normalised window has one) on a sample that is
// >= +0 -- colour tables and premultiplied samples are never negative and the vertical pass never produces -0.0f for
// that reason -- so the vertical sums are bit-identical; a horizontal sum could at most end as -0.0 instead of +0.0
// (all samples of the window exact zeros), which clamps, compares and quantises like +0.0.
#ifndef IWB2_FIRST_ASSIGN
#define IWB2_FIRST_ASSIGN 1
#endif
// the slot whose output takes its first tap at row r of a uniform cycle (the slot emitted after row r-1), or -1
__host__ __device__ constexpr int iwb2_first_slot(int r, int S, int stepq) {
  return IWB2_FIRST_ASSIGN && r % stepq == 0 ? (r / stepq + S - 1) % S : -1;
}
// Horizontal uniform path: the epilogues of two adjacent pixels are evaluated together (twice the independent FP32
// decision chains per basic block; the kernel is latency-bound there at 3 warps per scheduler)
#ifndef IWB2_H_PAIR
#define IWB2_H_PAIR 1
#endif
#ifndef IWB2_OST_PITCH
#define IWB2_OST_PITCH 100       // bytes per output staging row (25 words, odd: lane-private banks); <= 96 bytes staged
#endif
#define IWB2_NHINT 4096
#define IWB2_WPAD 64             // schedule rows of zero weights before row 0 and after the last row
#define IWB2_MAX_GU 16           // longest uniform cycle (inputs) the kernels are built for

// pipeline geometry of the general path as a function of the kernel's uniform step (0 = general-only kernel)
// (16-bit input: a shallow ring.  That kernel is bound by its lookups in the 65 536-entry code -> linear table, i.e. by
// the L1 hit rate of those lookups -- profiles/r2_gather_probe.txt -- and every KB of shared memory not used is L1)
#ifndef IWB2_V16_G
#define IWB2_V16_G 4
#endif
#ifndef IWB2_V16_NST
#define IWB2_V16_NST 3
#endif
__host__ __device__ constexpr int iwb2_v_g(int step, bool d16 = false) { return d16 ? IWB2_V16_G : (step > 0 ? IWB2_V_G : 2 * IWB2_V_G); }
__host__ __device__ constexpr int iwb2_v_nst(bool d16 = false) { return d16 ? IWB2_V16_NST : IWB2_V_NST; }
__host__ __device__ constexpr int iwb2_h_kp(int step) { return step > 0 ? IWB2_H_KP : 2 * IWB2_H_KP; }
__host__ __device__ constexpr int iwb2_h_nst(int step) { return step > 0 ? IWB2_H_NST : IWB2_H_NST + 1; }
// resident CTAs per SM the vertical kernel's register allocation aims at: 4 columns x S slots of FP64 accumulators
// per thread leave room for IWB2_V_MINB CTAs up to 8 slots; enlargements (12 / 16 open windows) get the whole file
__host__ __device__ constexpr int iwb2_v_minb(int S) { return S <= 8 ? IWB2_V_MINB : 2; }
// 16-bit input, up to 8 slots: ONE CTA of 16 warps per SM whose shared memory, beside the pipeline rings, holds the first
// `lut_ns` entries (~14 000) of the 65 536-entry code -> linear table.  Lookups in global memory are bound by the L1 miss
// path at ~0.9 per clock per SM and a 512 KB table of doubles does not stay in a sectored L1
// (profiles/r2y_cfg5_v_summary.txt: 6.7 % hits); shared memory serves ~6 per clock (profiles/r2_gather_probe.txt): every
// code found there is a request less on the slow path.  The budget stops at 188 KB: the L1 that remains also tracks the
// misses in flight, and with the last 32 KB handed to the table the kernel is 40 % SLOWER (measured on cfg5: vertical
// pass 4.85 ms without a shared table, 4.19 ms with 16 warps / 188 KB, 4.33 ms / 150 KB, 5.98 ms / 220 KB, 4.58 ms with
// 12 warps, 4.49 ms with 8).  Enlargements with 12 / 16 slots need the whole register file per thread and keep the
// 4-warp layout without a shared table.
__host__ __device__ constexpr bool iwb2_v16_shared_table(int S, bool d16) { return d16 && S <= 8; }
#ifndef IWB2_V16_WARPS
#define IWB2_V16_WARPS 16
#endif
#ifndef IWB2_V16_BUDGET_KB
#define IWB2_V16_BUDGET_KB 188
#endif
__host__ __device__ constexpr int iwb2_v_warps(int S, bool d16) { return iwb2_v16_shared_table(S, d16) ? IWB2_V16_WARPS : IWB2_V_WARPS; }
__host__ __device__ constexpr int iwb2_v_minb2(int S, bool d16) { return iwb2_v16_shared_table(S, d16) ? 1 : iwb2_v_minb(S); }

struct Sched2 {                  // device pointers
  const double *w;               // row r at w + r*SS, r in [-IWB2_WPAD, n_in + IWB2_WPAD); SS = S rounded up to even
  const int *first, *last;       // [n_out] window of each output
  int n_in, n_out;
};

// How one launch splits its outputs over CTAs (rows for the vertical kernel, pixels for the horizontal one):
// n_u uniform parts of u_len outputs covering [u_lo, u_hi), then general parts of g_len outputs covering what is
// left before u_lo and after u_hi.  In the uniform range output o completes at input u_ref_last + step*(o - u_ref_o).
struct Part2 {
  int n_u, u_lo, u_hi, u_len;
  int g_len;
  int u_ref_o, u_ref_last;
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
  int ncols;                     // sample columns (= in_w * channels)
  const double *lut16;           // 8-bit input: [256][16] code -> linear, every entry repeated 16 times (the shared-memory layout)
  const double *lut64k;          // 16-bit input: [65536] code -> linear (global memory; entries below lut_ns also in shared memory)
  int lut_ns;
  int depth16;                   // samples are 16-bit big-endian
  Sched2 sc;
  float *inter;                  // row-major FP32 intermediate
  int pitch;                     // floats per row of `inter` (multiple of 4)
  int sh;                        // column c is stored at float index c + sh (0..3, see iwb2_h_shift)
  int row_base;                  // output row stored as row 0 of `inter`
  int row_lo, row_hi;            // output rows of this launch
  Part2 part;                    // strips
};

struct H2Params {
  const float *inter;            // row-major FP32 intermediate (read through a tensor map built at launch)
  int pitch, ncols, sh;          // floats per row of `inter`; valid floats per row (= in_w * channels); column shift
  int row_lo, row_hi;            // output rows of this launch; row_lo is row 0 of `inter`
  Sched2 sc;                     // in pixels
  Part2 part;                    // segments of output pixels (lengths are multiples of 4)
  int out_w;
  DevOutput out;
  DevChanOut ch[4];
  Quant2 q[4];
  int same_tables;               // every channel uses q[0]'s tables
  const iwb_affine_t *jump;
};

// Host-side description of one pass over a chunk of output rows; the module sizes strips / segments (one wave of
// resident CTAs, uniform parts where the schedule allows) and launches.
struct V2Job {
  V2Params p;                    // everything but `part`
  int S;
  const int *first_host, *last_host;   // [n_out] host copies of the schedule's windows
  int strip_override;            // > 0: rows per uniform strip (tuning)
  int allow_uniform;
};
struct H2Job {
  H2Params p;                    // everything but `part`
  int S, nch, epi;
  int nco;                       // output channels: nch, or 3 for one gray channel written as RGB
  const int *first_host, *last_host;
  int seg_override;
  int allow_uniform;
};

bool iwb2_tensor_maps_available();
bool iwb2_run_v(const V2Job &job, cudaStream_t st);   // false: no kernel variant / launch problem
bool iwb2_run_h(const H2Job &job, cudaStream_t st);
int iwb2_h_shift(const H2Job &job);                    // the column shift (0..3) this horizontal launch wants from the vertical one
bool iwb2_supported(int SV, int SH, int nch, int nco, int epi);
bool iwb2_v_supported(int SV);   // the vertical kernel alone (feeding another horizontal kernel: gather5.cu)

#endif  // !IWB_EMU
