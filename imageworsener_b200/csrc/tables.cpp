// tables.cpp -- host-side scalar colour maths and PRNG jump tables.
//
// Every transcendental that can influence an output bit is evaluated HERE, on the host, with the same libm the
// reference uses, as a function of an integer code; the device only ever reads the resulting tables.
#include <cmath>

#include "iwb_internal.h"
#include "prng.cuh"

// input/any colour space -> linear (reference: src/imagew-main.c:28-66)
double iwb_to_linear(double v, const iw_csdescr *cs) {
  switch (cs->cstype) {
    case IW_CSTYPE_LINEAR:
      return v;
    case IW_CSTYPE_GAMMA:
      return pow(v, cs->gamma);
    case IW_CSTYPE_REC709:
      return (v < 4.5 * 0.020) ? v / 4.5 : pow((v + 0.099) / 1.099, 1.0 / 0.45);
    default:  // sRGB, and anything unknown
      return (v <= 0.04045) ? v / 12.92 : pow((v + 0.055) / (1.055), 2.4);
  }
}

// linear -> target colour space; values just below 1 snap to 1 (reference: src/imagew-main.c:74-96, :427-446)
double iwb_from_linear(double v, const iw_csdescr *cs) {
  if (v > 0.999999999) return 1.0;
  switch (cs->cstype) {
    case IW_CSTYPE_LINEAR:
      return v;
    case IW_CSTYPE_GAMMA:
      return pow(v, 1.0 / cs->gamma);
    case IW_CSTYPE_REC709:
      return (v < 0.020) ? 4.5 * v : 1.099 * pow(v, 0.45) - 0.099;
    default:
      return (v <= 0.0031308) ? 12.92 * v : 1.055 * pow(v, 1.0 / 2.4) - 0.055;
  }
}

// colour -> gray with the context's formula (reference: src/imagew-main.c:278-305)
double iwb_gray_of(const iw_context *ctx, double r, double g, double b) {
  const double *w = ctx->gray_w;
  if (ctx->grayscale_formula == IW_GSF_WEIGHTED) return w[0] * r + w[1] * g + w[2] * b;
  if (ctx->grayscale_formula == IW_GSF_ORDERBYVALUE) {
    double v0, v1, v2;  // descending
    if (g <= r) { v0 = r; v1 = g; } else { v0 = g; v1 = r; }
    if (b <= v1) {
      v2 = b;
    } else {
      v2 = v1;
      if (b <= v0) { v1 = b; } else { v1 = v0; v0 = b; }
    }
    return w[0] * v0 + w[1] * v1 + w[2] * v2;
  }
  return 0.0;
}

void iwb_prng_jump_table(iwb_affine tbl[64]) {
  // one step: t -> (-A)*t - (A+1)  (mod p)
  uint64_t m = IWB_PRNG_P - IWB_PRNG_A, a = IWB_PRNG_P - (IWB_PRNG_A + 1);
  for (int k = 0; k < 64; k++) {
    tbl[k].mul = m;
    tbl[k].add = a;
    // (m,a) o (m,a) = (m*m, m*a + a)
    uint64_t m2 = iwb_mulmod_p(m, m);
    uint64_t a2 = iwb_addmod_p(iwb_mulmod_p(m, a), a);
    m = m2;
    a = a2;
  }
}

// Seeding (reference: src/imagew-util.c:214-218) followed by the first step, reduced mod p.
uint64_t iwb_prng_first_state(int seed) {
  uint32_t m0 = ((uint32_t)0x03333333) + (uint32_t)seed;
  uint32_t c0 = ((uint32_t)0x05555555) + (uint32_t)seed;
  uint64_t x = IWB_PRNG_A * (uint64_t)m0 + c0;
  // x's (carry, low) split is what the next direct step uses, so keep x itself when it is a valid state
  // (carry <= A, always true here because x < (A+1)*2^32); reduce only for the affine path.
  return x;
}
