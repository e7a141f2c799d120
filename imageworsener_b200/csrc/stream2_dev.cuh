// stream2_dev.cuh -- device code of the second-generation streaming kernels (design: stream2.cuh).
//
// Each kernel has two code paths chosen per CTA:
//   * "uniform": the strip / segment lies where the schedule is regular -- output o completes exactly STEP inputs
//     after output o-1 (any integer reduction factor, away from the image border).  One cycle of S outputs then
//     consumes exactly S*STEP inputs = one pipeline stage, and the whole cycle is straight-line code: every
//     shared-memory access has an immediate offset, slots are compile-time registers, there is no loop or ring
//     bookkeeping per input, and the only branches are the (warp-uniform) emission guards;
//   * "general": any monotone contiguous-window schedule (per-output trip counts from the schedule tables).
// Both paths use deferred adds (the product of input i is added while input i+1 is multiplied), so no FP64
// instruction of an input depends on another one of the same input; and in-place rotation of the operands (the
// value of input i+1 replaces the value of input i right after its last use).  The order of the additions into
// each accumulator is the reference's (ascending source index, one DMUL + one DADD per tap, no FMA).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "iwb_internal.h"
#include "kernels.cuh"
#include "stream2.cuh"

// ---- mbarrier + bulk copy primitives -----------------------------------------------------------------
__device__ __forceinline__ unsigned s2_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s2_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void s2_mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s2_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra W_%=;\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void s2_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void s2_tma_2d(unsigned dst, const CUtensorMap *tm, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(tm), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool s2_elect() {                    // one lane of the (converged) warp
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void s2_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ unsigned s2_lds32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float s2_ldsf(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 s2_ldsw(unsigned a) {        // one pair of weights
  double x, y;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
  return make_double2(x, y);
}
__device__ __forceinline__ int s2_floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// Which strip / segment does CTA index `id` process?  Uniform ones first, then the general head and tail.
__device__ __forceinline__ void s2_decode_part(const Part2 &pt, int id, int lo, int hi, int *a, int *b, bool *uni) {
  if (id < pt.n_u) {
    *uni = true;
    *a = pt.u_lo + id * pt.u_len;
    *b = min(*a + pt.u_len, pt.u_hi);
    return;
  }
  *uni = false;
  int g = id - pt.n_u;
  const int nh = (pt.u_lo - lo + pt.g_len - 1) / pt.g_len;
  if (g < nh) {
    *a = lo + g * pt.g_len;
    *b = min(*a + pt.g_len, pt.u_lo);
  } else {
    g -= nh;
    *a = pt.u_hi + g * pt.g_len;
    *b = min(*a + pt.g_len, hi);
  }
}

// 16-bit input: lookups in the code -> linear table keep their lines in L1 as long as possible, the stores of the
// intermediate do not allocate there
__device__ __forceinline__ double s2_ldg_keep(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void s2_stg128_stream(float *p, float a, float b, float c, float d) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// ---------------------------------------------------------------------------------------------------------
// vertical kernel.  Warp = 128 adjacent byte columns (lane = 4 of them, one 32-bit word per row), strip of
// output rows; walks the input rows of the strip once.  Loads run ahead of the arithmetic: while row i is
// multiplied, the table values of row i+1 and the ring word of row i+2 are already in flight.
// ---------------------------------------------------------------------------------------------------------
// D16: the input samples are 16-bit big-endian (a warp's 128 columns are 256 bytes per row) and the code ->
// linear table has 65 536 entries: it stays in global memory (L2-resident, read-only path) instead of being
// replicated in shared memory.
template <int S, int STEP, int SH, bool D16>
__global__ void __launch_bounds__(32 * iwb2_v_warps(S, D16), iwb2_v_minb2(S, D16))
    iwb_v2(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_u, const V2Params p) {
  constexpr int SS = (S + 1) & ~1;
  // general path: rows per stage, stages.  A general-only kernel (non-integer factors: the whole image takes this
  // path) gets a deeper pipeline than the general parts of a uniform launch (a few border rows).
  constexpr int G = iwb2_v_g(STEP, D16), NST = iwb2_v_nst(D16);
  constexpr int TW = D16 ? 256 : 128;                           // bytes of a warp's 128 columns in one input row
  constexpr int STAGE = (G * TW + G * SS * 8 + 127) & ~127;     // bytes: input rows, then weights
  constexpr int GU = STEP > 0 ? S * STEP : 4, NSTU = IWB2_U_NST;   // uniform path: one cycle per stage
  constexpr int STEPQ = STEP > 0 ? STEP : 1;
  constexpr int STAGE_U = (GU * TW + GU * SS * 8 + 127) & ~127;
  constexpr int RING = NST * STAGE > NSTU * STAGE_U ? NST * STAGE : NSTU * STAGE_U;
  constexpr int WARP_BYTES = (RING + 8 * 8 + 127) & ~127;       // TMA tiles land on 128-byte boundaries
  static_assert(G >= 4 && NST >= 3 && GU >= 4 && NSTU >= 2, "pipeline depth");
  static_assert((STAGE % 128) == 0 && (STAGE_U % 128) == 0, "stage alignment");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int o0, o1;
  bool uni;
  s2_decode_part(p.part, blockIdx.y, p.row_lo, p.row_hi, &o0, &o1, &uni);
  if (STEP == 0) uni = false;

  constexpr int NWARPS = iwb2_v_warps(S, D16);
  constexpr bool TAB16 = iwb2_v16_shared_table(S, D16);
  // shared-memory colour table: 8-bit input [256][16]; 16-bit input (12-warp layout) the first lut_ns entries (a multiple of 16)
  const int LUTB = D16 ? (TAB16 ? p.lut_ns * 8 : 0) : 256 * 16 * 8;
  double *lut_s = (double *)smem_raw;
  unsigned char *wbase = smem_raw + LUTB + (size_t)warp * WARP_BYTES;
  const unsigned bar0 = s2_addr(wbase + RING);
  int *last_s = (int *)(smem_raw + LUTB + (size_t)NWARPS * WARP_BYTES);   // general path: [nout]
  const int wcol0 = (blockIdx.x * NWARPS + warp) * 128;
  const bool active = wcol0 < p.ncols;
  const unsigned ring0 = s2_addr(wbase);
  const int col0 = wcol0 + lane * 4;
  const bool full = col0 + 4 <= p.ncols;

  // general-path geometry (from the schedule tables)
  int r0 = 0, nr = 0, ob = o0, nout = 0, nstg = 0;
  // uniform-path geometry (closed form): cycle c = outputs [og0 + c*S, +S) = rows [R0 + c*GU, +GU)
  int og0 = 0, ncyc = 0, R0 = 0;
  if (uni) {
    og0 = s2_floor_div(o0 - (S - 1), S) * S;
    ncyc = (o1 - og0 + S - 1) / S;
    R0 = p.part.u_ref_last + STEP * (og0 - p.part.u_ref_o) - STEP + 1;
  } else {
    r0 = p.sc.first[o0];
    nr = p.sc.last[o1 - 1] - r0 + 1;
    while (ob > 0 && p.sc.last[ob - 1] >= r0) ob--;             // first output still open at r0
    nout = o1 - ob;
    nstg = (nr + G - 1) / G;
  }
  // stage n of the general path = input rows [r0 + n*G, +G) x this warp's 128 byte columns as one TMA tile (rows /
  // columns outside the tensor are zero-filled and never consumed) + the G schedule rows of weights
  auto fill = [&](int n, int slot) {
    if (s2_elect()) {
      s2_fence_async();   // this warp's earlier (generic-proxy) reads of the slot are ordered before the bulk copy into it
      const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE;
      s2_mbar_expect(bar, G * TW + G * SS * 8);
      s2_tma_2d(dst, &tmap, wcol0 * (TW / 128), r0 + n * G - p.tm_row0, bar);
      s2_bulk_g2s(dst + G * TW, p.sc.w + (ptrdiff_t)(r0 + n * G) * SS, G * SS * 8, bar);
    }
  };
  auto fill_u = [&](int n, int slot) {                          // uniform path: stage n = cycle n (+1 of overrun)
    if (s2_elect()) {
      s2_fence_async();   // this warp's earlier (generic-proxy) reads of the slot are ordered before the bulk copy into it
      const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE_U;
      s2_mbar_expect(bar, GU * TW + GU * SS * 8);
      s2_tma_2d(dst, &tmap_u, wcol0 * (TW / 128), R0 + n * GU - p.tm_row0, bar);
      s2_bulk_g2s(dst + GU * TW, p.sc.w + (ptrdiff_t)(R0 + n * GU) * SS, GU * SS * 8, bar);
    }
  };
  if (active) {                                                 // start the pipeline before the table loads
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < 8; s++) s2_mbar_init(bar0 + 8 * s, 1);
      s2_fence_async();
    }
    __syncwarp();
    if (uni) {
#pragma unroll
      for (int s = 0; s < NSTU; s++)
        if (s <= ncyc) fill_u(s, s);
    } else {
#pragma unroll
      for (int s = 0; s < NST; s++)
        if (s < nstg) fill(s, s);
    }
  }
  // the colour table arrives already replicated 16 ways ([256][16] doubles = 32 KB) by one bulk copy
  uint64_t *lbar = (uint64_t *)(last_s + ((nout + 1) & ~1));
  if (!D16 && threadIdx.x == 0) {
    const unsigned lb = s2_addr(lbar);
    s2_mbar_init(lb, 1);
    s2_fence_async();
    s2_mbar_expect(lb, 256 * 16 * 8);
    s2_bulk_g2s(s2_addr(lut_s), p.lut16, 256 * 16 * 8, lb);
  }
  if (!uni)
    for (int i = threadIdx.x; i < nout; i += blockDim.x) last_s[i] = p.sc.last[ob + i] - r0;
  if (TAB16)
    for (int i = threadIdx.x; i < p.lut_ns / 2; i += blockDim.x)
      reinterpret_cast<double2 *>(lut_s)[i] = __ldg(reinterpret_cast<const double2 *>(p.lut64k) + i);
  __syncthreads();
  if (!active) return;
  if (!D16) s2_mbar_wait(s2_addr(lbar), 0);

  double acc[S][4], prod[S][4];                                 // slot accumulators and deferred products
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < 4; c++) { acc[s][c] = 0.0; prod[s][c] = 0.0; }

  const unsigned lut_lane = (lane & 15) * 8;                    // this lane's bank pair: bits 3..6 of the table offset
  const char *lut_b = (const char *)lut_s;
  auto lut_at = [&](unsigned off) { return *(const double *)(lut_b + off); };   // (the table base folds into the LDS immediate)
  // a thread's 4 samples of one row: 4 bytes (8-bit) or 8 bytes (16-bit big-endian)
  struct word { unsigned a, b; };
  auto ring_word = [](unsigned addr) {
    word x;
    if (D16) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x.a), "=r"(x.b) : "r"(addr));
    else { x.a = s2_lds32(addr); x.b = 0; }
    return x;
  };
  auto lookup1 = [&](const word &x, int c) {
    if (D16) {                                                  // bytes (hi, lo) of sample c -> code -> global table
      const unsigned wv = c < 2 ? x.a : x.b;
      const unsigned code = __byte_perm(wv, 0, (c & 1) ? 0x4423 : 0x4401);
      if (TAB16) {                                              // (both loads predicated; no branch)
        double v = 0.0;
        if (code < (unsigned)p.lut_ns) v = lut_s[code];
        else v = s2_ldg_keep(p.lut64k + code);
        return v;
      }
      return s2_ldg_keep(p.lut64k + code);
    }
    // 8-bit: code * 128 bytes (16 copies of 8 bytes) + lane copy
    const unsigned sh = c == 0 ? (x.a << 7) : (x.a >> (8 * c - 7));
    return lut_at((sh & 0x7f80u) | lut_lane);
  };
  // one input row: drain last row's products into the accumulators, form this row's products (deferred), and
  // replace each column's value by the next row's as soon as its products are formed
  // `first` / `second` (uniform path only, compile-time constants after unrolling; -1 elsewhere): the slot whose output
  // takes its FIRST tap at this row / took it at the row before.  A first product is assigned instead of added to a
  // zeroed accumulator (see IWB2_FIRST_ASSIGN in stream2.cuh for why the bytes cannot change).
  auto taps = [&](double (&vc)[4], const word &xn, unsigned wrow, int first, int second) {
    double2 w[SS / 2];
#pragma unroll
    for (int s2 = 0; s2 < SS / 2; s2++) w[s2] = s2_ldsw(wrow + 16 * s2);
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
      for (int s = 0; s < S; s++) {
        const double wv = (s & 1) ? w[s >> 1].y : w[s >> 1].x;
        if (IWB2_DEFER_V) {
          if (s == second) acc[s][c] = prod[s][c];
          else if (s != first) acc[s][c] = __dadd_rn(acc[s][c], prod[s][c]);
          prod[s][c] = __dmul_rn(vc[c], wv);
        } else {
          if (s == first) acc[s][c] = __dmul_rn(vc[c], wv);
          else acc[s][c] = __dadd_rn(acc[s][c], __dmul_rn(vc[c], wv));
        }
      }
      vc[c] = lookup1(xn, c);
    }
  };
  // slot k is complete: add its pending products, round to FP32 (reference: src/imagew-main.c:905-912) and store the
  // warp's 128 columns of this intermediate row as one coalesced 512-byte segment; restart the slot.
  // Column c is stored at float index c + SH: the shift (0..3, chosen per launch, a template parameter here) puts
  // the horizontal kernel's tiles on 16-byte boundaries, which its TMA loads need; here it only splits the store.
  float *outp = p.inter + (ptrdiff_t)(o0 - p.row_base) * p.pitch + SH + col0;   // this thread's 4 floats of row o0
  const ptrdiff_t out_step = p.pitch;
  auto finish = [&](int k, int o, bool restart) {
#pragma unroll
    for (int c = 0; c < 4; c++) if (IWB2_DEFER_V) { acc[k][c] = __dadd_rn(acc[k][c], prod[k][c]); if (restart) prod[k][c] = 0.0; }
    if (o >= o0 && o < o1) {                                    // emitted rows are consecutive: outp walks down the plane
      const float f0 = (float)acc[k][0], f1 = (float)acc[k][1], f2 = (float)acc[k][2], f3 = (float)acc[k][3];
      if (full && D16 && SH == 0) {
        s2_stg128_stream(outp, f0, f1, f2, f3);                 // (keeps the intermediate out of L1: that is the table's)
      } else if (full) {
        if (SH == 0) {
          *reinterpret_cast<float4 *>(outp) = make_float4(f0, f1, f2, f3);
        } else if (SH == 2) {
          *reinterpret_cast<float2 *>(outp) = make_float2(f0, f1);
          *reinterpret_cast<float2 *>(outp + 2) = make_float2(f2, f3);
        } else {                                                // SH odd: outp + 1 is 8-byte aligned
          outp[0] = f0;
          *reinterpret_cast<float2 *>(outp + 1) = make_float2(f1, f2);
          outp[3] = f3;
        }
      } else {
        if (col0 + 0 < p.ncols) outp[0] = f0;
        if (col0 + 1 < p.ncols) outp[1] = f1;
        if (col0 + 2 < p.ncols) outp[2] = f2;
        if (col0 + 3 < p.ncols) outp[3] = f3;
      }
      outp += out_step;
    }
    if (restart) {
#pragma unroll
      for (int c = 0; c < 4; c++) acc[k][c] = 0.0;
    }
  };

  if (STEP > 0 && uni) {
    // ---- uniform path --------------------------------------------------------------------------------
    int slot = 0;
    unsigned par = 0;
    s2_mbar_wait(bar0, 0);
    double vc[4];
    {
      const word x0 = ring_word(ring0 + lane * (TW / 32));
#pragma unroll
      for (int c = 0; c < 4; c++) vc[c] = lookup1(x0, c);       // row 0
    }
    word xn = ring_word(ring0 + lane * (TW / 32) + TW);         // row 1
    for (int cyc = 0; cyc < ncyc; cyc++) {
      const int nslot = slot + 1 == NSTU ? 0 : slot + 1;
      const unsigned npar = nslot == 0 ? par ^ 1 : par;
      const unsigned sb = ring0 + slot * STAGE_U + lane * (TW / 32), sbn = ring0 + nslot * STAGE_U + lane * (TW / 32);
      const unsigned wb = ring0 + slot * STAGE_U + GU * TW;
      const int og = og0 + cyc * S;
#pragma unroll
      for (int r = 0; r < GU; r++) {
        if (r == GU - 2) s2_mbar_wait(bar0 + 8 * nslot, npar);  // the last two rows prefetch from the next stage
        const word xcur = xn;
        xn = ring_word(r + 2 < GU ? sb + (r + 2) * TW : sbn + (r + 2 - GU) * TW);   // row r+2
        taps(vc, xcur, wb + r * SS * 8, iwb2_first_slot(r, S, STEPQ), iwb2_first_slot((r + GU - 1) % GU, S, STEPQ));   // row r (values of row r+1 from xcur)
        if ((r + 1) % STEPQ == 0) finish((r + 1) / STEPQ - 1, og + (r + 1) / STEPQ - 1, !IWB2_FIRST_ASSIGN);
      }
      __syncwarp();
      if (cyc + NSTU <= ncyc) fill_u(cyc + NSTU, slot);
      slot = nslot;
      par = npar;
    }
    return;
  }

  // ---- general path ------------------------------------------------------------------------------------
  // read position (ring words): stage / slot / parity; weights position: slot.  The read position runs two rows
  // ahead of the weights of the row being consumed; a stage's slot is refilled when the read position leaves the
  // FOLLOWING stage (the lag is shorter than a stage), so NST-2 stages are in flight.
  int ld_stage = 0, ld_slot = 0, w_slot = 0;
  unsigned parity = 0;
  s2_mbar_wait(bar0, 0);
  unsigned pin = ring0 + lane * (TW / 32);
  unsigned pin_end = pin + G * TW;
  unsigned pw = ring0 + G * TW;
  unsigned pw_end = pw + G * SS * 8;
  auto advance_load = [&]() {                                   // the read position leaves stage ld_stage
    __syncwarp();
    if (ld_stage >= 1 && ld_stage - 1 + NST < nstg) fill(ld_stage - 1 + NST, ld_slot == 0 ? NST - 1 : ld_slot - 1);
    ld_stage++;
    if (++ld_slot == NST) { ld_slot = 0; parity ^= 1; }
    if (ld_stage < nstg) s2_mbar_wait(bar0 + 8 * ld_slot, parity);
    pin = ring0 + ld_slot * STAGE + lane * (TW / 32);
    pin_end = pin + G * TW;
  };
  auto next_word = [&]() {
    const word x = ring_word(pin);
    pin += TW;
    if (pin == pin_end) advance_load();
    return x;
  };
  double vc[4];
  {
    const word x0 = next_word();
#pragma unroll
    for (int c = 0; c < 4; c++) vc[c] = lookup1(x0, c);         // row 0
  }
  word xn = next_word();                                        // row 1
  int i = 0;                                                    // row being consumed (relative to r0)
  auto row_step = [&]() {
    taps(vc, xn, pw, -1, -1);
    xn = next_word();                                           // row i+2
    pw += SS * 8;
    if (pw == pw_end) {
      if (++w_slot == NST) w_slot = 0;
      pw = ring0 + w_slot * STAGE + G * TW;
      pw_end = pw + G * SS * 8;
    }
    i++;
  };
  for (int og = ob - (ob % S); og < o1; og += S) {
#pragma unroll
    for (int k = 0; k < S; k++) {
      const int o = og + k;
      if (o >= ob && o < o1) {                                  // warp-uniform
        const int iend = last_s[o - ob];
        while (i <= iend) row_step();
        finish(k, o, true);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// horizontal kernel + 8-bit epilogue.  Warp = 32 rows (one row group) x one segment of output pixels.
// EPI: 1 = no dither, 3 = ordered, 4 = random (all channels threshold-table mode, same dither); 2 = nearest-colour
// midpoint table (the reference's fast path for undithered 8-bit output).
// ---------------------------------------------------------------------------------------------------------
// ordered-dither matrices (reference: src/imagew-main.c:460-480), entry k means threshold (k + 0.5) / 64
__constant__ unsigned char iwb2_ord[128] = {
    0, 48, 12, 60, 3, 51, 15, 63, 32, 16, 44, 28, 35, 19, 47, 31, 8, 56, 4, 52, 11, 59, 7, 55, 40, 24, 36, 20, 43, 27, 39, 23,
    2, 50, 14, 62, 1, 49, 13, 61, 34, 18, 46, 30, 33, 17, 45, 29, 10, 58, 6, 54, 9, 57, 5, 53, 42, 26, 38, 22, 41, 25, 37, 21,
    3, 9, 17, 27, 25, 15, 7, 1, 11, 29, 37, 45, 43, 35, 23, 5, 19, 39, 51, 57, 55, 49, 33, 13, 31, 47, 59, 63, 61, 53, 41, 21,
    30, 46, 58, 62, 60, 52, 40, 20, 18, 38, 50, 56, 54, 48, 32, 12, 10, 28, 36, 44, 42, 34, 22, 4, 2, 8, 16, 26, 24, 14, 6, 0};

// The exact FP64 decision for a sample the FP32 test could not settle (reference: put_sample_convert_from_linear,
// src/imagew-main.c:669-765, in the table form of kernels.cuh:candidates_table/quantise for levels whose stored
// code is the level index).  `q` is the FP32 guess of the level; the true one is found by stepping from it.
template <int FAM>
__device__ __noinline__ int exact8_v2(const double *thr, int top, double s, int q, float t, uint32_t draw) {
  s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
  if (s > 0.999999999) return top;
  q = max(0, min(q, top));
  while (q < top && thr[q + 1] <= s) q++;
  while (q > 0 && thr[q] > s) q--;
  if (q >= top) return top;
  if (FAM == 4) return q;                                       // nearest colour: the cell index is the code
  const double d_floor = __dsub_rn(s, thr[q]), d_ceil = __dsub_rn(thr[q + 1], s);
  if (FAM == 0) return d_ceil <= d_floor ? q + 1 : q;
  const double frac = __ddiv_rn(d_floor, __dadd_rn(d_floor, d_ceil));
  const double th = FAM == 1 ? (double)t : __ddiv_rn((double)draw, (double)0xffffffffu);
  return frac >= th ? q + 1 : q;
}

// FP32 decision, branch-free: returns the code and whether it is settled (outside the error margin).  The
// channels of a pixel are evaluated side by side (no branch between them, so their table-lookup chains overlap)
// and the rare unsettled samples are redone exactly afterwards, behind one warp vote.
template <int EPI>
__device__ __forceinline__ int quantise8_try(unsigned tab, float margin, int top, float t, double s, bool *sure, int *qguess) {
  // (tab = 32-bit shared-memory address of the channel's table blob: explicit ld.shared keeps the address
  // arithmetic in 32 bits and avoids generic-pointer conversions in the hot path)
  const float sf = __saturatef((float)s);                       // NaN/negative -> 0, > 1 -> 1
  const int b = (int)(sf * (float)IWB2_NHINT);
  unsigned q0u;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(q0u) : "r"(tab + IWB2_TAB_HINT + 2 * b));
  const int q0 = (int)q0u;
  float4 e;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(e.x), "=f"(e.y), "=f"(e.z), "=f"(e.w) : "r"(tab + 16 * q0));
  const bool up = sf >= e.z;
  const float lo = up ? e.z : e.x, inv = up ? e.w : e.y;
  const int q = q0 + (up ? 1 : 0);
  const float fr = (sf - lo) * inv;
  const float d = fr - t;
  int code = q + (d >= 0.0f ? 1 : 0);
  // random thresholds reach 0 and 1: a sample within the margin of a level boundary also goes the exact way
  bool ok = EPI == 4 ? (fabsf(d) > margin && fr > margin && fr < 1.0f - margin) : fabsf(d) > margin;
  if (EPI == 2) {                                               // nearest colour: the cell is the code, if safely inside it
    code = q;
    ok = fr > margin && fr < 1.0f - margin;
  }
  if (EPI == 4) {
    // random dither compares against draw/0xffffffff which can be exactly 0 or 1: keep the reference's exact
    // end cases (src/imagew-main.c:618-621, :684-685)
    if (!(s > 0.0)) { code = 0; ok = true; }
    if (s > 0.999999999) { code = top; ok = true; }
  }
  *sure = ok;
  *qguess = q;
  return code;
}

// Write the warp's staged output bytes: `nrows` rows of `nbytes` bytes, coalesced per row (4-byte words when the
// destination allows).  Out of line: it runs once per ~40 pixels and would otherwise be inlined into every phase.
static __device__ __noinline__ void h2_flush(const unsigned char *ost, uint8_t *dst, size_t bpr, int nbytes, int nrows, bool walign, int lane) {
  __syncwarp();
  if (walign && (nbytes & 3) == 0) {
    if (lane * 4 < nbytes) {
      for (int k = 0; k < nrows; k++)
        *reinterpret_cast<unsigned *>(dst + (size_t)k * bpr + lane * 4) = *reinterpret_cast<const unsigned *>(ost + k * IWB2_OST_PITCH + lane * 4);
    }
  } else {
    for (int k = 0; k < nrows; k++)
      for (int b = lane; b < nbytes; b += 32) dst[(size_t)k * bpr + b] = ost[k * IWB2_OST_PITCH + b];
  }
  __syncwarp();
}

// Shared-memory pitch (floats) of a tile of `k` floats per row: a multiple of 4 with an odd number of 16-byte
// chunks, so that 32 lanes reading 16 bytes each from 32 different rows touch every bank exactly once per
// quarter-warp (the row-major intermediate is transposed for free by this access pattern).
__host__ __device__ constexpr int iwb2_tile_pitch(int k) { return (((k + 3) / 4) | 1) * 4; }

// NCO = output channels: NCH, or 3 for a gray image written as RGB (the reference widens gray input to RGB unless asked
// for grayscale, src/imagew-main.c:1676-1684: three intermediate channels reading the same input channel, so one
// resampled value serves the three; they differ only in their random-dither draws)
template <int S, int STEP, int NCH, int EPI, int NCO = NCH>
__global__ void
#if IWB2_H_MINB > 0
    __launch_bounds__(32 * IWB2_H_WARPS, IWB2_H_MINB)
#else
    __launch_bounds__(32 * IWB2_H_WARPS)
#endif
    iwb_h2(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_u, const H2Params p) {
  constexpr int SS = (S + 1) & ~1;
  constexpr bool DEFER = IWB2_DEFER_H && S * NCH <= 24;         // (deferred adds double the accumulator registers)
  constexpr int KP = iwb2_h_kp(STEP), NST = iwb2_h_nst(STEP);   // (deeper for a general-only kernel, see iwb_v2)
  constexpr int TP = iwb2_tile_pitch(KP * NCH + 3);             // general path: floats per tile row (+3: start rounded down)
  static_assert((KP * NCH) % 4 == 0, "general stages advance by whole 16-byte chunks");
  constexpr int STAGE = (32 * TP * 4 + KP * SS * 8 + 127) & ~127;
  constexpr int GU = STEP > 0 ? S * STEP : 4, NSTU = IWB2_U_NST;
  constexpr int STEPQ = STEP > 0 ? STEP : 1;
  constexpr int TPU = iwb2_tile_pitch(GU * NCH);                // uniform path
  constexpr int STAGE_U = (32 * TPU * 4 + GU * SS * 8 + 127) & ~127;
  static_assert(STEP == 0 || (GU * NCH) % 4 == 0, "a uniform stage must end on a 16-byte chunk");
  constexpr int RING = NST * STAGE > NSTU * STAGE_U ? NST * STAGE : NSTU * STAGE_U;
  constexpr int WARP_BYTES = (RING + 32 * IWB2_OST_PITCH + 8 * 8 + 127) & ~127;
  static_assert(NCO == NCH || (NCH == 1 && NCO == 3), "output channels");
  constexpr int FU = ((IWB2_OST_PITCH - 4) / NCO) & ~3;         // pixels per flush (multiple of 4; FU * NCO bytes fit a staging row)
  static_assert(KP >= 2 && NST >= 3 && GU >= 2 && NSTU >= 2, "pipeline depth");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int xo0, xo1;
  bool uni;
  const int part_id = IWB2_H_SWAP ? blockIdx.y : blockIdx.x, rowblock = IWB2_H_SWAP ? blockIdx.x : blockIdx.y;
  s2_decode_part(p.part, part_id, 0, p.out_w, &xo0, &xo1, &uni);
  if (STEP == 0) uni = false;
  const int ntab = p.same_tables ? 1 : NCO;

  unsigned char *wbase = smem_raw + (size_t)warp * WARP_BYTES;
  unsigned char *ost = wbase + RING;                            // [32 rows][pitch]
  const unsigned bar0 = s2_addr(wbase + RING + 32 * IWB2_OST_PITCH);
  unsigned char *tab_s = smem_raw + (size_t)IWB2_H_WARPS * WARP_BYTES;          // [ntab][IWB2_TAB_BYTES]
  float *ord_s = (float *)(tab_s + (size_t)ntab * IWB2_TAB_BYTES);              // [64]
  uint64_t *tbar = (uint64_t *)(ord_s + 64);
  int *last_s = (int *)(tbar + 2);                                              // general path: [nout]

  const int grp = rowblock * IWB2_H_WARPS + warp;               // row group of this warp
  const int jb = p.row_lo + grp * 32;
  const bool active = jb < p.row_hi;
  const int y = jb + lane;
  const unsigned ring0 = s2_addr(wbase);

  int b0 = 0, nb = 0, xb = xo0, nout = 0, nstg = 0;             // general-path geometry
  int og0 = 0, ncyc = 0, B0 = 0;                                // uniform-path geometry (see the vertical kernel)
  if (uni) {
    og0 = s2_floor_div(xo0 - (S - 1), S) * S;
    ncyc = (xo1 - og0 + S - 1) / S;
    B0 = p.part.u_ref_last + STEP * (og0 - p.part.u_ref_o) - STEP + 1;
  } else {
    b0 = p.sc.first[xo0];
    nb = p.sc.last[xo1 - 1] - b0 + 1;
    while (xb > 0 && p.sc.last[xb - 1] >= b0) xb--;             // first output still open at b0
    nout = xo1 - xb;
    nstg = (nb + KP - 1) / KP;
  }
  // stage n = KP (GU) pixel columns x this warp's 32 rows of the row-major intermediate as one 2-D TMA tile
  // (columns / rows outside the tensor are zero-filled: they meet zero weights or rows that are not stored) + the
  // schedule rows of weights
  // (pixel x, channel c is float x*NCH + c + p.sh of a row.  TMA tiles must start on 16 bytes: the general path
  // rounds the start down and skips the 0..3 leading floats; uniform stages start aligned by the choice of p.sh.)
  const int gskip = (b0 * NCH + p.sh) & 3;                      // (KP*NCH is a multiple of 4: the same for every stage)
  auto fill = [&](int n, int slot) {
    if (s2_elect()) {
      s2_fence_async();   // this warp's earlier (generic-proxy) reads of the slot are ordered before the bulk copy into it
      const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE;
      s2_mbar_expect(bar, 32 * TP * 4 + KP * SS * 8);
      s2_tma_2d(dst, &tmap, (b0 + n * KP) * NCH + p.sh - gskip, grp * 32, bar);
      s2_bulk_g2s(dst + 32 * TP * 4, p.sc.w + (ptrdiff_t)(b0 + n * KP) * SS, KP * SS * 8, bar);
    }
  };
  auto fill_u = [&](int n, int slot) {
    if (s2_elect()) {
      s2_fence_async();   // this warp's earlier (generic-proxy) reads of the slot are ordered before the bulk copy into it
      const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE_U;
      s2_mbar_expect(bar, 32 * TPU * 4 + GU * SS * 8);
      s2_tma_2d(dst, &tmap_u, (B0 + n * GU) * NCH + p.sh, grp * 32, bar);
      s2_bulk_g2s(dst + 32 * TPU * 4, p.sc.w + (ptrdiff_t)(B0 + n * GU) * SS, GU * SS * 8, bar);
    }
  };
  // prologue: the (asynchronous) copy of the decision tables, this warp's pipeline, then the small loops
  if (threadIdx.x == 0) {
    const unsigned tb = s2_addr(tbar);
    s2_mbar_init(tb, 1);
    s2_fence_async();
    s2_mbar_expect(tb, (unsigned)ntab * IWB2_TAB_BYTES);
    for (int k = 0; k < ntab; k++) s2_bulk_g2s(s2_addr(tab_s + (size_t)k * IWB2_TAB_BYTES), p.q[k].blob, IWB2_TAB_BYTES, tb);
  }
  if (active) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < 8; s++) s2_mbar_init(bar0 + 8 * s, 1);
      s2_fence_async();
    }
    __syncwarp();
    if (uni) {
#pragma unroll
      for (int s = 0; s < NSTU; s++)
        if (s <= ncyc) fill_u(s, s);
    } else {
#pragma unroll
      for (int s = 0; s < NST; s++)
        if (s < nstg) fill(s, s);
    }
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x)
    ord_s[i] = ((float)iwb2_ord[(p.ch[0].dsubtype == 1 ? 64 : 0) + i] + 0.5f) / 64.0f;   // exact in float
  if (!uni)
    for (int i = threadIdx.x; i < nout; i += blockDim.x) last_s[i] = p.sc.last[xb + i] - b0;
  __syncthreads();
  if (!active) return;
  s2_mbar_wait(s2_addr(tbar), 0);

  double acc[S][NCH], prod[S][NCH];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < NCH; c++) { acc[s][c] = 0.0; prod[s][c] = 0.0; }

  // random dither: this thread's draws are consecutive along its row, one per output pixel and channel
  uint64_t rng[NCO];
#pragma unroll
  for (int c = 0; c < NCO; c++) {
    rng[c] = 0;
    if (EPI == 4 && y < p.row_hi) {
      uint64_t t = p.ch[c].prng_t1;
      const uint64_t n = (uint64_t)y * (uint64_t)p.out_w + (uint64_t)xo0;
      if (n) { if (t >= IWB_PRNG_P) t -= IWB_PRNG_P; t = iwb_prng_jump(t, n, p.jump); }
      rng[c] = t;
    }
  }
  const unsigned ord_a = s2_addr(ord_s) + (y & 7) * 32;         // this row's eight ordered-dither thresholds
  unsigned tab_c[NCO];                                          // per output channel: table blob address, margin, top level
  float qmargin[NCO];
  int qtop[NCO];
#pragma unroll
  for (int c = 0; c < NCO; c++) {
    const int tb = p.same_tables ? 0 : c;
    tab_c[c] = s2_addr(tab_s) + tb * IWB2_TAB_BYTES;
    qmargin[c] = p.q[tb].margin;
    qtop[c] = p.q[tb].top;
  }

  // one input pixel column; nf[] = the next column's samples (already on their way from shared memory)
  // (`first` / `second`: as in the vertical kernel)
  auto taps = [&](double (&vc)[NCH], const float (&nf)[NCH], unsigned wrow, int first, int second) {
    double2 w[SS / 2];
#pragma unroll
    for (int s2 = 0; s2 < SS / 2; s2++) w[s2] = s2_ldsw(wrow + 16 * s2);
#pragma unroll
    for (int c = 0; c < NCH; c++) {
#pragma unroll
      for (int s = 0; s < S; s++) {
        const double wv = (s & 1) ? w[s >> 1].y : w[s >> 1].x;
        if (DEFER) {
          if (s == second) acc[s][c] = prod[s][c];
          else if (s != first) acc[s][c] = __dadd_rn(acc[s][c], prod[s][c]);
          prod[s][c] = __dmul_rn(vc[c], wv);
        } else {
          if (s == first) acc[s][c] = __dmul_rn(vc[c], wv);
          else acc[s][c] = __dadd_rn(acc[s][c], __dmul_rn(vc[c], wv));
        }
      }
      vc[c] = (double)nf[c];                                    // next column, in place
    }
  };

  // output staging: the warp's 32 rows x up to FU pixels, flushed as 4-byte words (coalesced per row)
  int staged_from = xo0;
  const unsigned ost_a = s2_addr(ost) + lane * IWB2_OST_PITCH;
  const bool walign = (((uintptr_t)p.out.pix | p.out.bpr) & 3) == 0;
  auto flush = [&](int upto) {
    h2_flush(ost, p.out.pix + (size_t)(jb - p.out.row_origin) * p.out.bpr + (size_t)staged_from * NCO, p.out.bpr,
             (upto - staged_from) * NCO, min(32, p.row_hi - jb), walign, lane);
    staged_from = upto;
  };
  // quantise + dither NP (1 or 2) adjacent output pixels xo, xo+1 whose sums are sv[0], sv[1], and stage their bytes.
  // The decision chains of all their channels sit in one basic block (no branch between them), so their table lookups
  // overlap; one warp vote covers the rare samples that need the exact decision.  valid1: pixel xo+1 is inside the
  // segment (the last pixel of an odd-width image has no partner).
  auto emit = [&](const double (&sv)[2][NCH], int xo, int np, bool valid1) {
    if (xo - staged_from >= FU) flush(xo);
    const unsigned dstb = ost_a + (xo - staged_from) * NCO;
    float t[2][NCO];
    uint32_t draw[2][NCO];
    int code[2][NCO], qg[2][NCO];
    bool sure[2][NCO], all_sure = true;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (q >= np) continue;
#pragma unroll
      for (int c = 0; c < NCO; c++) {
        const int ci = NCO == NCH ? c : 0;                       // the intermediate channel behind output channel c
        if (NCO != NCH && EPI != 4 && c > 0) {                   // same sample, same threshold: same code
          t[q][c] = t[q][0]; draw[q][c] = 0; code[q][c] = code[q][0]; sure[q][c] = sure[q][0]; qg[q][c] = qg[q][0];
          continue;
        }
        t[q][c] = EPI == 3 ? s2_ldsf(ord_a + 4 * ((xo + q) & 7)) : 0.5f;
        draw[q][c] = 0;
        if (EPI == 4 && (q == 0 || valid1)) {                    // (warp-uniform: draws are consumed by stored pixels only)
          draw[q][c] = iwb_prng_draw_of(rng[c]);
          rng[c] = iwb_prng_step(rng[c]);
          t[q][c] = (float)draw[q][c] * (1.0f / 4294967295.0f);
        }
        code[q][c] = quantise8_try<EPI>(tab_c[c], qmargin[c], qtop[c], t[q][c], sv[q][ci], &sure[q][c], &qg[q][c]);
        all_sure = all_sure && (sure[q][c] || (q == 1 && !valid1));
      }
    }
    if (__any_sync(0xffffffffu, !all_sure)) {                   // a few 1e-4 of the samples: exact FP64 decision
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (q >= np) continue;
#pragma unroll
        for (int c = 0; c < NCO; c++) {
          const int ci = NCO == NCH ? c : 0;
          if (NCO != NCH && EPI != 4 && c > 0) { code[q][c] = code[q][0]; continue; }
          if (!sure[q][c] && (q == 0 || valid1)) {
            const int tb = p.same_tables ? 0 : c;
            code[q][c] = exact8_v2<EPI == 1 ? 0 : (EPI == 2 ? 4 : (EPI == 3 ? 1 : 3))>((const double *)(tab_s + (size_t)tb * IWB2_TAB_BYTES + IWB2_TAB_THR),
                                                                   p.q[tb].top, sv[q][ci], qg[q][c], t[q][c], draw[q][c]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (q >= np || (q == 1 && !valid1)) continue;
#pragma unroll
      for (int c = 0; c < NCO; c++) asm volatile("st.shared.u8 [%0], %1;" ::"r"(dstb + q * NCO + c), "r"(code[q][c]) : "memory");
    }
  };
  // slot k is complete: add its pending products (deferred adds); `restart`: zero the slot for its next output
  auto drain = [&](int k, bool restart) {
#pragma unroll
    for (int c = 0; c < NCH; c++) if (DEFER) { acc[k][c] = __dadd_rn(acc[k][c], prod[k][c]); if (restart) prod[k][c] = 0.0; }
  };
  auto finish = [&](int k, int xo, bool restart) {
    drain(k, restart);
    if (xo >= xo0 && xo < xo1) {
      double sv[2][NCH];
#pragma unroll
      for (int c = 0; c < NCH; c++) { sv[0][c] = acc[k][c]; sv[1][c] = 0.0; }
      emit(sv, xo, 1, false);
    }
    if (restart) {
#pragma unroll
      for (int c = 0; c < NCH; c++) acc[k][c] = 0.0;
    }
  };

  if (STEP > 0 && uni) {
    // ---- uniform path --------------------------------------------------------------------------------
    int slot = 0;
    unsigned par = 0;
    s2_mbar_wait(bar0, 0);
    // This lane's row of the tile is read 16 bytes at a time (conflict-free, see iwb2_tile_pitch); `ch` is the
    // chunk holding the next sample to hand out.  Every index below is a compile-time constant after unrolling.
    const unsigned rowoff = lane * (TPU * 4);
    float4 ch;
    auto lds128 = [](unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; };
    auto elem = [](const float4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); };
    ch = lds128(ring0 + rowoff);
    double vc[NCH];
    double hold[2][NCH];
    constexpr bool PAIR = IWB2_H_PAIR && S % 2 == 0;
#pragma unroll
    for (int c = 0; c < NCH; c++) { vc[c] = (double)elem(ch, c); hold[0][c] = hold[1][c] = 0.0; }   // column 0 (NCH <= 4: inside the first chunk)
    for (int cyc = 0; cyc < ncyc; cyc++) {
      const int nslot = slot + 1 == NSTU ? 0 : slot + 1;
      const unsigned npar = nslot == 0 ? par ^ 1 : par;
      const unsigned sb = ring0 + slot * STAGE_U + rowoff, sbn = ring0 + nslot * STAGE_U + rowoff;
      const unsigned wb = ring0 + slot * STAGE_U + 32 * TPU * 4;
      const int og = og0 + cyc * S;
#pragma unroll
      for (int r = 0; r < GU; r++) {
        if (r == GU - 1) s2_mbar_wait(bar0 + 8 * nslot, npar);  // the last column prefetches from the next stage
        float nf[NCH];
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          const int f = (r + 1 < GU ? (r + 1) * NCH : 0) + c;   // flat index of the sample within its stage
          if (f % 4 == 0) ch = lds128((r + 1 < GU ? sb : sbn) + f * 4);
          nf[c] = elem(ch, f % 4);
        }
        taps(vc, nf, wb + r * SS * 8, iwb2_first_slot(r, S, STEPQ), iwb2_first_slot((r + GU - 1) % GU, S, STEPQ));
        if ((r + 1) % STEPQ == 0) {
          const int k = (r + 1) / STEPQ - 1, xo = og + k;
          if (!PAIR) {
            finish(k, xo, !IWB2_FIRST_ASSIGN);
          } else {
            // pixels are decided two at a time (og and S are even: pixel og + k with k even is the first of its pair,
            // and a segment starts on a multiple of 4): the even one waits in `hold` for its neighbour
            drain(k, !IWB2_FIRST_ASSIGN);
            if (k % 2 == 0) {
#pragma unroll
              for (int c = 0; c < NCH; c++) hold[0][c] = acc[k][c];
            } else {
#pragma unroll
              for (int c = 0; c < NCH; c++) hold[1][c] = acc[k][c];
              if (xo - 1 >= xo0 && xo - 1 < xo1) emit(hold, xo - 1, 2, xo < xo1);
            }
            if (!IWB2_FIRST_ASSIGN) {
#pragma unroll
              for (int c = 0; c < NCH; c++) acc[k][c] = 0.0;
            }
          }
        }
      }
      __syncwarp();
      if (cyc + NSTU <= ncyc) fill_u(cyc + NSTU, slot);
      slot = nslot;
      par = npar;
    }
    flush(xo1);
    return;
  }

  // ---- general path ------------------------------------------------------------------------------------
  // read position (samples, one pixel ahead of the arithmetic) and weights position, as in the vertical kernel
  int ld_stage = 0, ld_slot = 0, w_slot = 0;
  unsigned parity = 0;
  s2_mbar_wait(bar0, 0);
  // (this path reads its row of the tile 4 bytes at a time, which costs a 4-way bank conflict per load; it serves
  // the image borders and non-integer scale factors)
  const unsigned rowoff = lane * (TP * 4) + gskip * 4;
  unsigned pin = ring0 + rowoff;
  unsigned pin_end = pin + KP * NCH * 4;
  unsigned pw = ring0 + 32 * TP * 4;
  unsigned pw_end = pw + KP * SS * 8;
  auto advance_load = [&]() {
    __syncwarp();
    if (ld_stage >= 1 && ld_stage - 1 + NST < nstg) fill(ld_stage - 1 + NST, ld_slot == 0 ? NST - 1 : ld_slot - 1);
    ld_stage++;
    if (++ld_slot == NST) { ld_slot = 0; parity ^= 1; }
    if (ld_stage < nstg) s2_mbar_wait(bar0 + 8 * ld_slot, parity);
    pin = ring0 + ld_slot * STAGE + rowoff;
    pin_end = pin + KP * NCH * 4;
  };
  double vc[NCH];
#pragma unroll
  for (int c = 0; c < NCH; c++) vc[c] = (double)s2_ldsf(pin + 4 * c);         // pixel column 0
  pin += NCH * 4;
  if (pin == pin_end) advance_load();
  int i = 0;                                                    // pixel column being consumed (relative to b0)
  auto col_step = [&]() {
    float nf[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) nf[c] = s2_ldsf(pin + 4 * c);
    taps(vc, nf, pw, -1, -1);
    pin += NCH * 4;
    if (pin == pin_end) advance_load();
    pw += SS * 8;
    if (pw == pw_end) {
      if (++w_slot == NST) w_slot = 0;
      pw = ring0 + w_slot * STAGE + 32 * TP * 4;
      pw_end = pw + KP * SS * 8;
    }
    i++;
  };
  for (int og = xb - (xb % S); og < xo1; og += S) {
#pragma unroll
    for (int k = 0; k < S; k++) {
      const int o = og + k;
      if (o >= xb && o < xo1) {                                 // warp-uniform
        const int iend = last_s[o - xb];
        while (i <= iend) col_step();
        finish(k, o, true);
      }
    }
  }
  flush(xo1);
}
