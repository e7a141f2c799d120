// estimate.h -- a piecewise-quadratic model of "linear sample -> output level" that lets the 16-bit epilogue decide
// almost every sample without touching the 65 536-entry threshold table.
//
// The exact rule (reference: src/imagew-main.c:669-765, no dither): with thr[q] the linear value of code q, a sample v
// with thr[q] <= v < thr[q+1] is stored as q+1 iff thr[q+1] - v <= v - thr[q], i.e. the stored code is the number of
// linear-space midpoints m[q] = (thr[q] + thr[q+1]) / 2 that are <= v.  The model P(v) approximates that step function
// continuously: P(m[q]) = q + 1 up to an error E that the HOST MEASURES on every midpoint of the job's actual table,
// with the device's own arithmetic (iwb_est_eval below runs on both sides, operation for operation).  P is increasing on
// every segment, so for m[q-1] <= v < m[q]:  q - E <= P(v) < q + 1 + E, and floor(P(v)) is the reference's code
// whenever P(v) is farther than `margin` = E + slack from an integer.  Samples closer than that (about 2 in 1000),
// samples in segments the host could not validate (the knee of the sRGB / Rec.709 curves), NaNs and anything outside
// [0, 1) take the exact decision on the table.  Nothing here is a tolerance on the output: every path ends in the
// reference's comparison or in a decision that provably equals it.
//
// Segments: IWB_EST_PER_OCTAVE per binary octave of v from 2^-IWB_EST_OCTAVES to 1 (the index is a shift of the high
// word of v); samples below the first octave use the first segment's polynomial (the curves are linear there; checked
// like everything else), samples >= 1 use a constant record (top + 1/2).  16 bytes per record: one LDS.128.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

#define IWB_EST_OCTAVES 12
#define IWB_EST_SHIFT 14                                    // 20 - log2(segments per octave)
#define IWB_EST_PER_OCTAVE (1 << (20 - IWB_EST_SHIFT))
#define IWB_EST_NREC (IWB_EST_OCTAVES * IWB_EST_PER_OCTAVE + 1)
#define IWB_EST_HI_MIN ((1023 - IWB_EST_OCTAVES) << 20)     // high word of 2^-IWB_EST_OCTAVES
#define IWB_EST_HI_MAX (1023 << 20)                         // high word of 1.0
#define IWB_EST_MAGIC 6755399441055744.0                    // 2^52 + 2^51: adding it rounds to an integer in the low word

struct IwbEstRec { double c0; float c1, c2; };              // P = c0 + (c1 + c2 u) u,  u = v - (start of the segment)

#if defined(__CUDACC__)
#define IWB_EST_HD __host__ __device__ __forceinline__
#else
#define IWB_EST_HD inline
#endif

// The model's value at v (+ 1/2: floor of it is the estimated code).  NaN when the segment is not validated.
IWB_EST_HD double iwb_est_eval_rec(const IwbEstRec &r, double v, int hc) {
  uint64_t b0 = (uint64_t)(uint32_t)(hc & ~((1 << IWB_EST_SHIFT) - 1)) << 32;
  double v0;
#if defined(__CUDA_ARCH__)
  v0 = __longlong_as_double((long long)b0);
  const double u = __dsub_rn(v, v0);
  const float w = __fmaf_rn(r.c2, (float)u, r.c1);
  return __fma_rn((double)w, u, r.c0);
#else
  memcpy(&v0, &b0, 8);
  const double u = v - v0;
  const float w = fmaf(r.c2, (float)u, r.c1);
  return fma((double)w, u, r.c0);
#endif
}
IWB_EST_HD int iwb_est_clamped_hi(int hi) { return hi < IWB_EST_HI_MIN ? IWB_EST_HI_MIN : (hi > IWB_EST_HI_MAX ? IWB_EST_HI_MAX : hi); }
IWB_EST_HD int iwb_est_index(int hc) { return (int)((unsigned)(hc - IWB_EST_HI_MIN) >> IWB_EST_SHIFT); }

struct IwbEstTable {
  std::vector<IwbEstRec> rec;    // IWB_EST_NREC records
  double margin = 0;             // distance from an integer below which the model does not decide
  double worst = 0;              // largest |P(m[q]) - (q + 1)| over the validated segments
  int invalid = 0;               // segments left to the table
};
// thr: the job's table (linear value of every code, thr.size() = top + 1).  False: no model for this colour space / table.
bool iwb_build_est_table(int cstype, const std::vector<double> &thr, IwbEstTable *out);
