// prng.cuh -- the dither PRNG with exact skip-ahead, usable on host and device.
//
// The reference generator (reference: src/imagew-util.c:199-233) keeps 32-bit (multiply, carry):
//     x = A*multiply + carry;  carry = x >> 32;  multiply = 0xffffffff - (x & 0xffffffff);  return multiply
// with A = 0xfff0bf23.  Writing t = carry*2^32 + (0xffffffff - multiply) -- which, for any state produced by a
// step, is just that step's x -- one step is the affine map  t' = (-A*t - (A+1)) mod p,  p = A*2^32 + 1
// (because A*2^32 == -1 mod p).  Affine maps compose, so draw number n of a channel is obtained by applying
// step^(n) (binary powering over a precomputed table of step^(2^k)) to t_1, the state after the first step:
//     draw_n = 0xffffffff - (t_{n+1} & 0xffffffff),   t_{n+1} = step^n(t_1),   t_1 = A*m0 + c0.
// Random dither consumes exactly one draw per output sample in raster order (reference:
// src/imagew-main.c:719-723, :755), so sample (x, y) of a W-wide image uses draw y*W + x.
#pragma once
#include <cstdint>

#ifndef IWB_HD
#if defined(__CUDACC__) && !defined(IWB_EMU)
#define IWB_HD __host__ __device__ __forceinline__
#else
#define IWB_HD inline
#endif
#endif

#define IWB_PRNG_A 0xfff0bf23ull
#define IWB_PRNG_P 0xfff0bf2300000001ull

IWB_HD void iwb_mul64(uint64_t a, uint64_t b, uint64_t *hi, uint64_t *lo) {
#if defined(__CUDA_ARCH__)
  *lo = a * b;
  *hi = __umul64hi(a, b);
#else
  unsigned __int128 r = (unsigned __int128)a * b;
  *lo = (uint64_t)r;
  *hi = (uint64_t)(r >> 64);
#endif
}

// (hi*2^64 + lo) mod p for hi < p.  Long division by the 32-bit A in three 64/32 steps gives
// q = floor(x / (A*2^32)) and r = x mod (A*2^32); since A*2^32 == -1 (mod p), x == r - q.
IWB_HD uint64_t iwb_mod_p(uint64_t hi, uint64_t lo) {
  const uint64_t A = IWB_PRNG_A;
  // x >> 32 as three 32-bit limbs d2:d1:d0 plus carry-in of hi's upper half
  uint64_t d3 = hi >> 32, d2 = hi & 0xffffffffull, d1 = lo >> 32, d0 = lo & 0xffffffffull;
  // x = ((d3*2^32 + d2)*2^32 + d1)*2^32 + d0.  Divide N = d3:d2:d1 (96 bits) by A; remainder r1 < A.
  uint64_t r = d3 % A;               // d3 < 2^32, contributes q3 = d3 / A (0 or 1) at weight 2^64
  uint64_t q3 = d3 / A;
  uint64_t n2 = (r << 32) | d2;      // r < A < 2^32
  uint64_t q2 = n2 / A; r = n2 % A;
  uint64_t n1 = (r << 32) | d1;
  uint64_t q1 = n1 / A; r = n1 % A;
  // quotient Q = q3*2^64 + q2*2^32 + q1 (q2,q1 < 2^32), remainder R = r*2^32 + d0 < A*2^32 < p
  uint64_t R = (r << 32) | d0;
  // result = (R - Q) mod p.  Reduce Q mod p first: 2^64 mod p = 2^64 - p.
  const uint64_t P = IWB_PRNG_P;
  uint64_t Q = (q2 << 32) + q1;      // q2 < 2^32 so no overflow; Q < 2^64
  if (Q >= P) Q -= P;
  if (q3) {                          // add q3 * (2^64 mod p)
    uint64_t c = (0ull - P);         // 2^64 - p
    Q += c;                          // Q < p and c < 2^52: no overflow
    if (Q >= P) Q -= P;
  }
  return R >= Q ? R - Q : R + (P - Q);
}

IWB_HD uint64_t iwb_mulmod_p(uint64_t a, uint64_t b) {
  uint64_t hi, lo;
  iwb_mul64(a, b, &hi, &lo);
  return iwb_mod_p(hi, lo);
}
IWB_HD uint64_t iwb_addmod_p(uint64_t a, uint64_t b) {  // a,b < p
  uint64_t s = a + b;
  if (s < a || s >= IWB_PRNG_P) s -= IWB_PRNG_P;
  return s;
}

struct iwb_affine_t { uint64_t mul, add; };

// t after n more steps, given the table tbl[k] = step^(2^k).
IWB_HD uint64_t iwb_prng_jump(uint64_t t, uint64_t n, const iwb_affine_t *tbl) {
  for (int k = 0; n; k++, n >>= 1)
    if (n & 1) t = iwb_addmod_p(iwb_mulmod_p(tbl[k].mul, t), tbl[k].add);
  return t;
}
// One direct step in the reference's own form; t is a valid (post-step) state.
IWB_HD uint64_t iwb_prng_step(uint64_t t) {
  uint32_t carry = (uint32_t)(t >> 32);
  uint32_t mult = 0xffffffffu - (uint32_t)t;
  return IWB_PRNG_A * (uint64_t)mult + carry;
}
IWB_HD uint32_t iwb_prng_draw_of(uint64_t t) { return 0xffffffffu - (uint32_t)t; }
