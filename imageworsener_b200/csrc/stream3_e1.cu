// stream3_e1.cu -- fused vertical + horizontal kernel, epilogue variant 1 (see stream3.inc).
#define IWB3_EPI 1
#include "stream3.inc"

bool iwb3_run_e2(const V2Job &, const H2Job &, VH3State *, int, int, cudaStream_t);
bool iwb3_run_e3(const V2Job &, const H2Job &, VH3State *, int, int, cudaStream_t);
bool iwb3_run_e4(const V2Job &, const H2Job &, VH3State *, int, int, cudaStream_t);

bool iwb3_supported(int SV, int SH, int nch, int nco, int epi, int depth16) {
  if (SV != SH || depth16 || nch != nco || (nch != 1 && nch != 3) || epi < 1 || epi > 4) return false;
#define X(a, b) if (SV == a) return true;
  IWB3_PAIRS(X)
#undef X
  return false;
}

bool iwb3_run(const V2Job &vj, const H2Job &hj, VH3State *st, int nv_override, int seg_override, cudaStream_t stream, bool *launched) {
  bool ok = false;
  switch (hj.epi) {
    case 1: ok = iwb3_run_e1(vj, hj, st, nv_override, seg_override, stream); break;
    case 2: ok = iwb3_run_e2(vj, hj, st, nv_override, seg_override, stream); break;
    case 3: ok = iwb3_run_e3(vj, hj, st, nv_override, seg_override, stream); break;
    case 4: ok = iwb3_run_e4(vj, hj, st, nv_override, seg_override, stream); break;
  }
  if (launched) *launched = ok;
  return ok;
}
