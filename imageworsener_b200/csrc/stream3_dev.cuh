// stream3_dev.cuh -- the fused vertical + horizontal kernel (design: stream3.cuh).
//
// One persistent CTA per SM.  Its warps are specialised: V warps walk strips of the image in input-row order and write
// the FP32 intermediate (reference: src/imagew-main.c:831-927, rounded to float exactly as :905-912), H warps turn
// finished groups of 32 intermediate rows into output pixels (reference: src/imagew-main.c:929-1155).  Both kinds are
// resident together, so the horizontal epilogue (FP32 / integer / shared-memory work) issues in the slots the FP64 pipe
// leaves free under the vertical pass.  The intermediate goes through global memory but is consumed while it is still in
// L2; a group's lines are never read again after its H items, and the launch that follows overwrites them in L2.
//
// Work comes from two queues (atomic tickets): V items = (128-column warp column, row part), H items = (group of 32
// output rows, segment of output pixels) in the order their groups are expected to complete.  A V warp announces the
// rows it finished in a group by a release-add on the group's counter (one counter per group and block of warp
// columns); an H warp acquires the counters of the blocks its segment reads before it starts its pipeline.  V items
// never wait for anything, every H item only waits for V items, so the scheme cannot deadlock whatever the order in
// which the hardware runs the warps.
//
// The arithmetic of an item is the one of stream2_dev.cuh, statement for statement: same schedules, same slot
// accumulators, taps in ascending source order, one DMUL + one DADD per tap, same epilogue.
#pragma once
#include "stream2_dev.cuh"
#include "stream3.cuh"

__device__ __forceinline__ void s3_mbar_inval(unsigned bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void s3_fence_proxy_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// (polling with a relaxed load and one acquire fence once the value is there: an acquire load costs a full L1
// invalidation per poll)
__device__ __forceinline__ int s3_ld_acquire(const int *p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void s3_acquire_fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void s3_red_release(int *p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Every wait of this kernel is bounded: a warp that has waited ~4 s (wall clock) for a copy or for another warp's
// announcement records what it was waiting for in the launch's diagnostic slot (page-locked host memory) and traps, so a
// protocol bug ends the launch with an error instead of hanging the device.
__device__ __forceinline__ unsigned long long s3_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
static __device__ __noinline__ void s3_die(int *dbg, int what, int a, int b, int c, int d) {
  if (dbg) {
    const int k = atomicAdd(dbg, 1);
    if (k < 15) {
      volatile int *r = dbg + 8 + 8 * k;
      r[0] = what; r[1] = (int)blockIdx.x; r[2] = (int)(threadIdx.x >> 5); r[3] = a; r[4] = b; r[5] = c; r[6] = d;
    }
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ void s3_mbar_wait(unsigned bar, unsigned parity, int *dbg, int what, int a, int b) {
  unsigned long long t0 = 0;
  for (unsigned spins = 0;; spins++) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((spins & 1023) == 1023) {
      const unsigned long long t = s3_now();
      if (!t0) t0 = t;
      else if (t - t0 > 4000000000ull) s3_die(dbg, what, a, b, (int)parity, (int)spins);
    }
  }
}
// The pipeline barriers of a warp live for the whole kernel.  Every copy a warp starts is waited for exactly once (both
// paths of both passes), so a slot's barrier alternates fill, wait, fill, wait ... across items, and one phase bit per slot
// (`ph`, a register that travels with the warp from item to item) is all the bookkeeping needed.
__device__ __forceinline__ void s3_init_bars(unsigned bar0, int lane) {
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < 8; s++) s2_mbar_init(bar0 + 8 * s, 1);
    s2_fence_async();
  }
  __syncwarp();
}
__device__ __forceinline__ void s3_slot_wait(unsigned bar0, int slot, unsigned &ph, int *dbg, int what, int a, int b) {
  s3_mbar_wait(bar0 + 8 * slot, (ph >> slot) & 1u, dbg, what, a, b);
  ph ^= 1u << slot;
}

// ---------------------------------------------------------------------------------------------------------
// One vertical item: output rows [o0, o1) of the warp column that starts at sample column wcol0.
// ---------------------------------------------------------------------------------------------------------
template <int S, int STEP, bool D16>
__device__ __forceinline__ void s3_v_item(const V2Params &p, const CUtensorMap *tmap, const CUtensorMap *tmap_u, const VH3Sync &sy,
                                          unsigned char *wbase, int ring_bytes, const double *lut_s, int wc, int o0, int o1, bool uni,
                                          unsigned &ph) {
  constexpr int SS = (S + 1) & ~1;
  constexpr int G = iwb2_v_g(STEP), NST = IWB2_V_NST;
  constexpr int TW = D16 ? 256 : 128;
  constexpr int STAGE = G * TW + G * SS * 8;
  constexpr int GU = STEP > 0 ? S * STEP : 4, NSTU = IWB2_U_NST;
  constexpr int STEPQ = STEP > 0 ? STEP : 1;
  constexpr int STAGE_U = (GU * TW + GU * SS * 8 + 127) & ~127;
  const int lane = threadIdx.x & 31;
  const int wcol0 = wc * 128;
  // a warp's region: [8 pipeline barriers, 128 bytes][ring][window table of the general path]
  const unsigned bar0 = s2_addr(wbase);
  const unsigned ring0 = bar0 + 128;
  int *last_s = (int *)(wbase + 128 + ring_bytes);              // general path: [nout]
  const int col0 = wcol0 + lane * 4;
  const bool full = col0 + 4 <= p.ncols;
  if (STEP == 0) uni = false;

  int r0 = 0, nr = 0, ob = o0, nout = 0, nstg = 0;
  int og0 = 0, ncyc = 0, R0 = 0;
  if (uni) {
    og0 = s2_floor_div(o0 - (S - 1), S) * S;
    ncyc = (o1 - og0 + S - 1) / S;
    R0 = p.part.u_ref_last + STEP * (og0 - p.part.u_ref_o) - STEP + 1;
  } else {
    r0 = p.sc.first[o0];
    nr = p.sc.last[o1 - 1] - r0 + 1;
    while (ob > 0 && p.sc.last[ob - 1] >= r0) ob--;
    nout = o1 - ob;
    nstg = (nr + G - 1) / G;
  }
  auto fill = [&](int n, int slot) {
    if (s2_elect()) {
      s2_fence_async();   // this warp's earlier (generic-proxy) reads of the slot are ordered before the bulk copy into it
      const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE;
      s2_mbar_expect(bar, G * TW + G * SS * 8);
      s2_tma_2d(dst, tmap, wcol0 * (TW / 128), r0 + n * G - p.tm_row0, bar);
      s2_bulk_g2s(dst + G * TW, p.sc.w + (ptrdiff_t)(r0 + n * G) * SS, G * SS * 8, bar);
    }
  };
  auto fill_u = [&](int n, int slot) {
    if (s2_elect()) {
      s2_fence_async();   // this warp's earlier (generic-proxy) reads of the slot are ordered before the bulk copy into it
      const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE_U;
      s2_mbar_expect(bar, GU * TW + GU * SS * 8);
      s2_tma_2d(dst, tmap_u, wcol0 * (TW / 128), R0 + n * GU - p.tm_row0, bar);
      s2_bulk_g2s(dst + GU * TW, p.sc.w + (ptrdiff_t)(R0 + n * GU) * SS, GU * SS * 8, bar);
    }
  };
  __syncwarp();                                                 // (every lane is done with the previous item's ring and window table)
  if (uni) {
#pragma unroll
    for (int s = 0; s < NSTU; s++)
      if (s <= ncyc) fill_u(s, s);
  } else {
#pragma unroll
    for (int s = 0; s < NST; s++)
      if (s < nstg) fill(s, s);
    for (int i = lane; i < nout; i += 32) last_s[i] = p.sc.last[ob + i] - r0;
    __syncwarp();
  }

  double acc[S][4];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < 4; c++) acc[s][c] = 0.0;

  const unsigned lut_lane = (lane & 15) * 8;
  const char *lut_b = (const char *)lut_s;
  auto lut_at = [&](unsigned off) { return *(const double *)(lut_b + off); };
  struct word { unsigned a, b; };
  auto ring_word = [](unsigned addr) {
    word x;
    if (D16) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x.a), "=r"(x.b) : "r"(addr));
    else { x.a = s2_lds32(addr); x.b = 0; }
    return x;
  };
  auto lookup1 = [&](const word &x, int c) {
    if (D16) {
      const unsigned wv = c < 2 ? x.a : x.b;
      const unsigned code = __byte_perm(wv, 0, (c & 1) ? 0x4423 : 0x4401);
      return __ldg(p.lut64k + code);
    }
    const unsigned sh = c == 0 ? (x.a << 7) : (x.a >> (8 * c - 7));
    return lut_at((sh & 0x7f80u) | lut_lane);
  };
  auto taps = [&](double (&vc)[4], const word &xn, unsigned wrow) {
    double2 w[SS / 2];
#pragma unroll
    for (int s2 = 0; s2 < SS / 2; s2++) w[s2] = s2_ldsw(wrow + 16 * s2);
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
      for (int s = 0; s < S; s++) acc[s][c] = __dadd_rn(acc[s][c], __dmul_rn(vc[c], (s & 1) ? w[s >> 1].y : w[s >> 1].x));
      vc[c] = lookup1(xn, c);
    }
  };
  // emitted rows are consecutive: outp walks down the plane.  Column c is stored at float index c + sh (see
  // stream2_dev.cuh); the shift is a launch parameter here, the store shape follows it (warp-uniform switch).
  const int sh = p.sh;
  float *outp = p.inter + (ptrdiff_t)(o0 - p.row_base) * p.pitch + sh + col0;
  const ptrdiff_t out_step = p.pitch;
  int *cnt = sy.cnt + (wc / sy.cbw);                            // this warp column's counter of group 0
  int pending = 0;                                              // rows finished in the current group, not yet announced
  // one finished intermediate row of this thread's 4 columns (already rounded to FP32, reference: src/imagew-main.c:905-912)
  auto emit = [&](int o, float f0, float f1, float f2, float f3) {
    if (o >= o0 && o < o1) {
      if (full) {
        if (sh == 0) {
          *reinterpret_cast<float4 *>(outp) = make_float4(f0, f1, f2, f3);
        } else if (sh == 2) {
          *reinterpret_cast<float2 *>(outp) = make_float2(f0, f1);
          *reinterpret_cast<float2 *>(outp + 2) = make_float2(f2, f3);
        } else {
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
      pending++;
      const int rel = o - p.row_lo;
      if ((rel & 31) == 31 || o == o1 - 1) {                    // the group (or the item) ends here: announce the rows
        __syncwarp();                                           // every lane's stores are ordered before lane 0's release
        if (lane == 0) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          s3_red_release(cnt + (ptrdiff_t)(rel >> 5) * sy.ncb, pending);
        }
        pending = 0;
      }
    }
  };
  auto finish = [&](int k, int o) {                             // (k is a compile-time constant at every call)
    const float f0 = (float)acc[k][0], f1 = (float)acc[k][1], f2 = (float)acc[k][2], f3 = (float)acc[k][3];
#pragma unroll
    for (int c = 0; c < 4; c++) acc[k][c] = 0.0;
    emit(o, f0, f1, f2, f3);
  };

  if (STEP > 0 && uni) {
    // One cycle = S outputs = S*STEP input rows = one stage.  The loop over the cycle's outputs is NOT unrolled: the
    // tap code is the same for every output (all S slots take a tap from every row), only the slot that completes
    // differs, and that is a small switch.  The body is ~1/S of the straight-line version of stream2_dev.cuh, so the
    // vertical and the horizontal loop fit the instruction cache together (ncu on the unrolled version of this kernel:
    // 2.6 warps stalled on instruction fetch per issue, against 0.07 / 0.24 in the separate kernels).
    int slot = 0;
    s3_slot_wait(bar0, 0, ph, sy.dbg, 1, wc, o0);
    double vc[4];
    {
      const word x0 = ring_word(ring0 + lane * (TW / 32));
#pragma unroll
      for (int c = 0; c < 4; c++) vc[c] = lookup1(x0, c);
    }
    word xn = ring_word(ring0 + lane * (TW / 32) + TW);
    for (int cyc = 0; cyc < ncyc; cyc++) {
      const int nslot = slot + 1 == NSTU ? 0 : slot + 1;
      const unsigned sb = ring0 + slot * STAGE_U + lane * (TW / 32), sbn = ring0 + nslot * STAGE_U + lane * (TW / 32);
      const unsigned wb = ring0 + slot * STAGE_U + GU * TW;
      const int og = og0 + cyc * S;
#pragma unroll 1
      for (int k = 0; k < S; k++) {
#pragma unroll
        for (int j = 0; j < STEPQ; j++) {
          const int r = k * STEPQ + j;
          if (r == GU - 2) s3_slot_wait(bar0, nslot, ph, sy.dbg, 2, wc, cyc);   // the last two rows prefetch from the next stage
          const word xcur = xn;
          xn = ring_word(r + 2 < GU ? sb + (r + 2) * TW : sbn + (r + 2 - GU) * TW);
          taps(vc, xcur, wb + r * SS * 8);
        }
        float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
        switch (k) {
#define IWB3_VCASE(K)                                                                                   \
  case K:                                                                                               \
    if (K < S) {                                                                                        \
      f0 = (float)acc[K < S ? K : 0][0]; f1 = (float)acc[K < S ? K : 0][1];                             \
      f2 = (float)acc[K < S ? K : 0][2]; f3 = (float)acc[K < S ? K : 0][3];                             \
      acc[K < S ? K : 0][0] = 0.0; acc[K < S ? K : 0][1] = 0.0; acc[K < S ? K : 0][2] = 0.0; acc[K < S ? K : 0][3] = 0.0; \
    }                                                                                                   \
    break;
          IWB3_VCASE(0) IWB3_VCASE(1) IWB3_VCASE(2) IWB3_VCASE(3) IWB3_VCASE(4) IWB3_VCASE(5) IWB3_VCASE(6) IWB3_VCASE(7)
#undef IWB3_VCASE
        }
        emit(og + k, f0, f1, f2, f3);
      }
      __syncwarp();
      if (cyc + NSTU <= ncyc) fill_u(cyc + NSTU, slot);
      slot = nslot;
    }
    return;
  }

  int ld_stage = 0, ld_slot = 0, w_slot = 0;
  s3_slot_wait(bar0, 0, ph, sy.dbg, 3, wc, o0);
  unsigned pin = ring0 + lane * (TW / 32);
  unsigned pin_end = pin + G * TW;
  unsigned pw = ring0 + G * TW;
  unsigned pw_end = pw + G * SS * 8;
  auto advance_load = [&]() {
    __syncwarp();
    if (ld_stage >= 1 && ld_stage - 1 + NST < nstg) fill(ld_stage - 1 + NST, ld_slot == 0 ? NST - 1 : ld_slot - 1);
    ld_stage++;
    if (++ld_slot == NST) ld_slot = 0;
    if (ld_stage < nstg) s3_slot_wait(bar0, ld_slot, ph, sy.dbg, 4, wc, ld_stage);
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
    for (int c = 0; c < 4; c++) vc[c] = lookup1(x0, c);
  }
  word xn = next_word();
  int i = 0;
  auto row_step = [&]() {
    taps(vc, xn, pw);
    xn = next_word();
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
      if (o >= ob && o < o1) {
        const int iend = last_s[o - ob];
        while (i <= iend) row_step();
        finish(k, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// One horizontal item: output pixels [xo0, xo1) of the 32 rows of group `grp`.
// ---------------------------------------------------------------------------------------------------------
template <int S, int STEP, int NCH, int EPI, int NCO>
__device__ __forceinline__ void s3_h_item(const H2Params &p, const CUtensorMap *tmap, const CUtensorMap *tmap_u, int *dbg, unsigned char *wbase,
                                          int ring_bytes, unsigned char *tab_s, const float *ord_s, int grp, int xo0, int xo1, bool uni,
                                          unsigned &ph) {
  constexpr int SS = (S + 1) & ~1;
  constexpr int KP = iwb2_h_kp(STEP), NST = iwb2_h_nst(STEP);
  constexpr int TP = iwb2_tile_pitch(KP * NCH + 3);
  constexpr int STAGE = (32 * TP * 4 + KP * SS * 8 + 127) & ~127;
  constexpr int GU = STEP > 0 ? S * STEP : 4, NSTU = IWB2_U_NST;
  constexpr int STEPQ = STEP > 0 ? STEP : 1;
  constexpr int TPU = iwb2_tile_pitch(GU * NCH);
  constexpr int STAGE_U = (32 * TPU * 4 + GU * SS * 8 + 127) & ~127;
  constexpr int FU = ((IWB2_OST_PITCH - 4) / NCO) & ~3;
  const int lane = threadIdx.x & 31;
  if (STEP == 0) uni = false;

  // a warp's region: [8 pipeline barriers, 128 bytes][ring][output staging][window table of the general path]
  unsigned char *ost = wbase + 128 + ring_bytes;
  const unsigned bar0 = s2_addr(wbase);
  const unsigned ring0 = bar0 + 128;
  int *last_s = (int *)(wbase + 128 + ring_bytes + 32 * IWB2_OST_PITCH);

  const int jb = p.row_lo + grp * 32;
  const int y = jb + lane;

  int b0 = 0, nb = 0, xb = xo0, nout = 0, nstg = 0;
  int og0 = 0, ncyc = 0, B0 = 0;
  if (uni) {
    og0 = s2_floor_div(xo0 - (S - 1), S) * S;
    ncyc = (xo1 - og0 + S - 1) / S;
    B0 = p.part.u_ref_last + STEP * (og0 - p.part.u_ref_o) - STEP + 1;
  } else {
    b0 = p.sc.first[xo0];
    nb = p.sc.last[xo1 - 1] - b0 + 1;
    while (xb > 0 && p.sc.last[xb - 1] >= b0) xb--;
    nout = xo1 - xb;
    nstg = (nb + KP - 1) / KP;
  }
  const int gskip = (b0 * NCH + p.sh) & 3;
  auto fill = [&](int n, int slot) {
    if (s2_elect()) {
      s2_fence_async();   // this warp's earlier (generic-proxy) reads of the slot are ordered before the bulk copy into it
      const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE;
      s2_mbar_expect(bar, 32 * TP * 4 + KP * SS * 8);
      s2_tma_2d(dst, tmap, (b0 + n * KP) * NCH + p.sh - gskip, grp * 32, bar);
      s2_bulk_g2s(dst + 32 * TP * 4, p.sc.w + (ptrdiff_t)(b0 + n * KP) * SS, KP * SS * 8, bar);
    }
  };
  auto fill_u = [&](int n, int slot) {
    if (s2_elect()) {
      s2_fence_async();   // this warp's earlier (generic-proxy) reads of the slot are ordered before the bulk copy into it
      const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE_U;
      s2_mbar_expect(bar, 32 * TPU * 4 + GU * SS * 8);
      s2_tma_2d(dst, tmap_u, (B0 + n * GU) * NCH + p.sh, grp * 32, bar);
      s2_bulk_g2s(dst + 32 * TPU * 4, p.sc.w + (ptrdiff_t)(B0 + n * GU) * SS, GU * SS * 8, bar);
    }
  };
  __syncwarp();                                                 // (every lane is done with the previous item's ring and window table)
  s3_fence_proxy_all();                                         // the rows were written through the generic proxy (other SMs)
  if (uni) {
#pragma unroll
    for (int s = 0; s < NSTU; s++)
      if (s <= ncyc) fill_u(s, s);
  } else {
#pragma unroll
    for (int s = 0; s < NST; s++)
      if (s < nstg) fill(s, s);
    for (int i = lane; i < nout; i += 32) last_s[i] = p.sc.last[xb + i] - b0;
    __syncwarp();
  }

  double acc[S][NCH];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < NCH; c++) acc[s][c] = 0.0;

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
  const unsigned ord_a = s2_addr(ord_s) + (y & 7) * 32;
  unsigned tab_c[NCO];
  float qmargin[NCO];
  int qtop[NCO];
#pragma unroll
  for (int c = 0; c < NCO; c++) {
    const int tb = p.same_tables ? 0 : c;
    tab_c[c] = s2_addr(tab_s) + tb * IWB2_TAB_BYTES;
    qmargin[c] = p.q[tb].margin;
    qtop[c] = p.q[tb].top;
  }
  auto taps = [&](double (&vc)[NCH], const float (&nf)[NCH], unsigned wrow) {
    double2 w[SS / 2];
#pragma unroll
    for (int s2 = 0; s2 < SS / 2; s2++) w[s2] = s2_ldsw(wrow + 16 * s2);
#pragma unroll
    for (int c = 0; c < NCH; c++) {
#pragma unroll
      for (int s = 0; s < S; s++) acc[s][c] = __dadd_rn(acc[s][c], __dmul_rn(vc[c], (s & 1) ? w[s >> 1].y : w[s >> 1].x));
      vc[c] = (double)nf[c];
    }
  };
  int staged_from = xo0;
  const unsigned ost_a = s2_addr(ost) + lane * IWB2_OST_PITCH;
  const bool walign = (((uintptr_t)p.out.pix | p.out.bpr) & 3) == 0;
  auto flush = [&](int upto) {
    h2_flush(ost, p.out.pix + (size_t)(jb - p.out.row_origin) * p.out.bpr + (size_t)staged_from * NCO, p.out.bpr,
             (upto - staged_from) * NCO, min(32, p.row_hi - jb), walign, lane);
    staged_from = upto;
  };
  // one finished output pixel: the resampled linear samples of its channels -> quantise + dither -> staged bytes
  auto emit = [&](int xo, const double (&sv)[NCH]) {
    if (xo >= xo0 && xo < xo1) {
      if (xo - staged_from >= FU) flush(xo);
      const unsigned dstb = ost_a + (xo - staged_from) * NCO;
      float t[NCO];
      uint32_t draw[NCO];
      int code[NCO], qg[NCO];
      bool sure[NCO], all_sure = true;
#pragma unroll
      for (int c = 0; c < NCO; c++) {
        const int ci = NCO == NCH ? c : 0;
        if (NCO != NCH && EPI != 4 && c > 0) {
          t[c] = t[0]; draw[c] = 0; code[c] = code[0]; sure[c] = sure[0]; qg[c] = qg[0];
          continue;
        }
        t[c] = EPI == 3 ? s2_ldsf(ord_a + 4 * (xo & 7)) : 0.5f;
        draw[c] = 0;
        if (EPI == 4) {
          draw[c] = iwb_prng_draw_of(rng[c]);
          rng[c] = iwb_prng_step(rng[c]);
          t[c] = (float)draw[c] * (1.0f / 4294967295.0f);
        }
        code[c] = quantise8_try<EPI>(tab_c[c], qmargin[c], qtop[c], t[c], sv[ci], &sure[c], &qg[c]);
        all_sure = all_sure && sure[c];
      }
      if (__any_sync(0xffffffffu, !all_sure)) {
#pragma unroll
        for (int c = 0; c < NCO; c++) {
          const int ci = NCO == NCH ? c : 0;
          if (NCO != NCH && EPI != 4 && c > 0) { code[c] = code[0]; continue; }
          if (!sure[c]) {
            const int tb = p.same_tables ? 0 : c;
            code[c] = exact8_v2<EPI == 1 ? 0 : (EPI == 2 ? 4 : (EPI == 3 ? 1 : 3))>((const double *)(tab_s + (size_t)tb * IWB2_TAB_BYTES + IWB2_TAB_THR),
                                                                 p.q[tb].top, sv[ci], qg[c], t[c], draw[c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NCO; c++) asm volatile("st.shared.u8 [%0], %1;" ::"r"(dstb + c), "r"(code[c]) : "memory");
    }
  };
  auto finish = [&](int k, int xo) {                            // (k is a compile-time constant at every call)
    double sv[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) { sv[c] = acc[k][c]; acc[k][c] = 0.0; }
    emit(xo, sv);
  };

  if (STEP > 0 && uni) {
    // The loop over a cycle's outputs is rolled, U outputs per iteration (U = the smallest count whose samples fill whole
    // 16-byte chunks of the tile row, so that the position of a sample within its chunk stays a compile-time constant):
    // the tap code does not depend on which slot completes, the epilogue exists U times instead of S times (see the
    // vertical item for why).
    constexpr int U = (STEPQ * NCH) % 4 == 0 ? 1 : ((2 * STEPQ * NCH) % 4 == 0 ? 2 : 4);
    static_assert(STEP == 0 || S % U == 0, "a cycle must be a whole number of loop iterations");
    int slot = 0;
    s3_slot_wait(bar0, 0, ph, dbg, 5, grp, xo0);
    const unsigned rowoff = lane * (TPU * 4);
    float4 ch;
    auto lds128 = [](unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; };
    auto elem = [](const float4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); };
    ch = lds128(ring0 + rowoff);
    double vc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) vc[c] = (double)elem(ch, c);
    for (int cyc = 0; cyc < ncyc; cyc++) {
      const int nslot = slot + 1 == NSTU ? 0 : slot + 1;
      const unsigned sb = ring0 + slot * STAGE_U + rowoff, sbn = ring0 + nslot * STAGE_U + rowoff;
      const unsigned wb = ring0 + slot * STAGE_U + 32 * TPU * 4;
      const int og = og0 + cyc * S;
#pragma unroll 1
      for (int kk = 0; kk < S / U; kk++) {
        double sv[U][NCH];
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
          for (int j = 0; j < STEPQ; j++) {
            const int r = (kk * U + u) * STEPQ + j;             // pixel column of the stage being consumed
            if (r == GU - 1) s3_slot_wait(bar0, nslot, ph, dbg, 6, grp, cyc);   // the last column prefetches from the next stage
            float nf[NCH];
#pragma unroll
            for (int c = 0; c < NCH; c++) {
              // flat index of the next column's sample within its stage: ((r + 1) mod GU) * NCH + c; its position in
              // the 16-byte chunk only depends on u, j, c (kk * U * STEP * NCH is a multiple of 4)
              constexpr int dummy = 0; (void)dummy;
              const int fpos = ((u * STEPQ + j + 1) * NCH + c) % 4;
              const int f = (r + 1 < GU ? (r + 1) * NCH : 0) + c;
              if (fpos == 0) ch = lds128((r + 1 < GU ? sb : sbn) + f * 4);
              nf[c] = elem(ch, fpos);
            }
            taps(vc, nf, wb + r * SS * 8);
          }
          // slot kk*U + u is complete
          switch (kk) {
#define IWB3_HCASE(K)                                                                                  \
  case K:                                                                                              \
    if (K * U + u < S) {                                                                               \
      _Pragma("unroll") for (int c = 0; c < NCH; c++) {                                                \
        sv[u][c] = acc[K * U + u < S ? K * U + u : 0][c];                                              \
        acc[K * U + u < S ? K * U + u : 0][c] = 0.0;                                                   \
      }                                                                                                \
    }                                                                                                  \
    break;
            IWB3_HCASE(0) IWB3_HCASE(1) IWB3_HCASE(2) IWB3_HCASE(3) IWB3_HCASE(4) IWB3_HCASE(5) IWB3_HCASE(6) IWB3_HCASE(7)
#undef IWB3_HCASE
          }
          emit(og + kk * U + u, sv[u]);
        }
      }
      __syncwarp();
      if (cyc + NSTU <= ncyc) fill_u(cyc + NSTU, slot);
      slot = nslot;
    }
    flush(xo1);
    return;
  }

  int ld_stage = 0, ld_slot = 0, w_slot = 0;
  s3_slot_wait(bar0, 0, ph, dbg, 7, grp, xo0);
  const unsigned rowoff = lane * (TP * 4) + gskip * 4;
  unsigned pin = ring0 + rowoff;
  unsigned pin_end = pin + KP * NCH * 4;
  unsigned pw = ring0 + 32 * TP * 4;
  unsigned pw_end = pw + KP * SS * 8;
  auto advance_load = [&]() {
    __syncwarp();
    if (ld_stage >= 1 && ld_stage - 1 + NST < nstg) fill(ld_stage - 1 + NST, ld_slot == 0 ? NST - 1 : ld_slot - 1);
    ld_stage++;
    if (++ld_slot == NST) ld_slot = 0;
    if (ld_stage < nstg) s3_slot_wait(bar0, ld_slot, ph, dbg, 8, grp, ld_stage);
    pin = ring0 + ld_slot * STAGE + rowoff;
    pin_end = pin + KP * NCH * 4;
  };
  double vc[NCH];
#pragma unroll
  for (int c = 0; c < NCH; c++) vc[c] = (double)s2_ldsf(pin + 4 * c);
  pin += NCH * 4;
  if (pin == pin_end) advance_load();
  int i = 0;
  auto col_step = [&]() {
    float nf[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) nf[c] = s2_ldsf(pin + 4 * c);
    taps(vc, nf, pw);
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
      if (o >= xb && o < xo1) {
        const int iend = last_s[o - xb];
        while (i <= iend) col_step();
        finish(k, o);
      }
    }
  }
  flush(xo1);
}

// which part of a launch's outputs is part index `id`: general parts first (the image borders: short items that
// should not end up last in the queue), then the uniform ones
__device__ __forceinline__ void s3_decode_part(const Part2 &pt, int n_g, int id, int lo, int hi, int *a, int *b, bool *uni) {
  s2_decode_part(pt, id < n_g ? pt.n_u + id : id - n_g, lo, hi, a, b, uni);
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
template <int S, int STEP, bool D16, int NCH, int EPI, int NCO>
__global__ void __launch_bounds__(32 * IWB3_WARPS, 1)
    iwb_vh3(const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap vmap_u, const __grid_constant__ CUtensorMap hmap,
            const __grid_constant__ CUtensorMap hmap_u, const VH3Params q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ntab = q.h.same_tables ? 1 : NCO;
  constexpr int LUTB = D16 ? 0 : 256 * 16 * 8;
  double *lut_s = (double *)smem_raw;
  unsigned char *tab_s = smem_raw + LUTB;
  float *ord_s = (float *)(tab_s + (size_t)ntab * IWB2_TAB_BYTES);
  uint64_t *cbar = (uint64_t *)(ord_s + 64);
  unsigned char *warps0 = (unsigned char *)(cbar + 16);         // 128 bytes of CTA barriers keep the warp regions on 128-byte boundaries
  // warps [0, nv) are vertical, the rest horizontal, each with a shared-memory region of its kind
  const bool is_v = warp < q.nv;
  unsigned char *wbase = is_v ? warps0 + (size_t)warp * q.v_warp_bytes
                              : warps0 + (size_t)q.nv * q.v_warp_bytes + (size_t)(warp - q.nv) * q.h_warp_bytes;

  // CTA prologue: colour table and decision tables arrive by bulk copies, the ordered-dither row table is computed
  if (threadIdx.x == 0) {
    const unsigned cb = s2_addr(cbar);
    s2_mbar_init(cb, 1);
    s2_fence_async();
    unsigned bytes = (unsigned)ntab * IWB2_TAB_BYTES + (D16 ? 0u : 256u * 16 * 8);
    s2_mbar_expect(cb, bytes);
    if (!D16) s2_bulk_g2s(s2_addr(lut_s), q.v.lut16, 256 * 16 * 8, cb);
    for (int k = 0; k < ntab; k++) s2_bulk_g2s(s2_addr(tab_s + (size_t)k * IWB2_TAB_BYTES), q.h.q[k].blob, IWB2_TAB_BYTES, cb);
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x)
    ord_s[i] = ((float)iwb2_ord[(q.h.ch[0].dsubtype == 1 ? 64 : 0) + i] + 0.5f) / 64.0f;
  __syncthreads();
  s3_mbar_wait(s2_addr(cbar), 0, q.sy.dbg, 10, 0, 0);

  unsigned ph = 0;                                              // phase bit of each pipeline slot of this warp
  s3_init_bars(s2_addr(wbase), lane);

  auto run_v = [&](int it) {
    const int part = it / q.nwc, wc = it - part * q.nwc;
    int o0, o1;
    bool uni;
    s3_decode_part(q.v.part, q.n_gv, part, q.v.row_lo, q.v.row_hi, &o0, &o1, &uni);
    s3_v_item<S, STEP, D16>(q.v, &vmap, &vmap_u, q.sy, wbase, q.v_ring_bytes, lut_s, wc, o0, o1, uni, ph);
  };
  struct hitem { int grp, xo0, xo1; bool uni; };
  auto h_decode = [&](int it) {
    hitem h;
    const int gi = it / q.nparts_h, xp = it - gi * q.nparts_h;
    h.grp = q.h_order[gi];
    s3_decode_part(q.h.part, q.n_gh, xp, 0, q.h.out_w, &h.xo0, &h.xo1, &h.uni);
    return h;
  };
  // Have the vertical items announced every row of the group in the columns this segment reads?  One counter per block
  // of warp columns, target = launch number x rows of the group x warp columns of the block (see VH3Sync).  Warp-uniform
  // result; `have` / `want` / `blk` describe a block that is still short (diagnostics).
  auto h_ready = [&](const hitem &h, int *blk, int *have, int *want) {
    const int ca = (q.h.sc.first[h.xo0] * NCH) / 128, cb = (q.h.sc.last[h.xo1 - 1] * NCH + NCH - 1) / 128;
    const int ba = ca / q.sy.cbw, bb = min(cb / q.sy.cbw, q.sy.ncb - 1);
    const int rows = min(32, q.h.row_hi - q.h.row_lo - h.grp * 32);
    const int *cnt = q.sy.cnt + (ptrdiff_t)h.grp * q.sy.ncb;
    bool ok = true;
    for (int b = ba + lane; b - lane <= bb; b += 32) {
      if (b <= bb) {
        const int w = q.sy.gen * rows * min(q.sy.cbw, q.nwc - b * q.sy.cbw);
        const int v = s3_ld_acquire(cnt + b);
        if (v < w) { ok = false; *blk = b; *have = v; *want = w; }
      }
    }
    return __all_sync(0xffffffffu, ok) != 0;
  };
  auto run_h = [&](const hitem &h) {
    s3_acquire_fence();
    s3_h_item<S, STEP, NCH, EPI, NCO>(q.h, &hmap, &hmap_u, q.sy.dbg, wbase, q.h_ring_bytes, tab_s, ord_s, h.grp, h.xo0, h.xo1, h.uni, ph);
  };
  auto wait_h = [&](const hitem &h) {
    int blk = 0, have = 0, want = 0;
    unsigned long long t0 = 0;
    for (unsigned spins = 0; !h_ready(h, &blk, &have, &want); spins++) {
      __nanosleep(500);
      if ((spins & 255) == 255) {
        const unsigned long long t = s3_now();
        if (!t0) t0 = t;
        else if (t - t0 > 4000000000ull) s3_die(q.sy.dbg, 9, h.grp, blk, have, want);
      }
    }
  };
  auto pop = [&](int *head, int base) {                         // one ticket (warp-uniform)
    int it = 0;
    if (lane == 0) it = atomicAdd(head, 1) - base;
    return __shfl_sync(0xffffffffu, it, 0);
  };

  if (is_v) {
    for (;;) {
      const int it = pop(q.sy.v_head, q.sy.v_base);
      if (it >= q.n_vitems) break;
      run_v(it);
    }
    return;
  }
  for (;;) {
    const int it = pop(q.sy.h_head, q.sy.h_base);
    if (it >= q.n_hitems) break;
    const hitem h = h_decode(it);
    wait_h(h);
    run_h(h);
  }
}
