// stream2_v.cu -- vertical streaming kernel: instantiations, planning, launch (design: stream2.cuh).
#define IWB2_D16 0
#include "stream2_v.inc"

bool iwb2_tensor_maps_available() { return iwb2_tensor_map_encoder() != nullptr; }

bool iwb2_supported(int SV, int SH, int nch, int nco, int epi) {
  return iwb2_s_built(SV) && iwb2_s_built(SH) && (nch == 1 || nch == 3) && (nco == nch || (nch == 1 && nco == 3)) &&
         (epi == 1 || epi == 2 || epi == 3 || epi == 4);
}

bool iwb2_v_supported(int SV) { return iwb2_s_built(SV) && iwb2_tensor_map_encoder() != nullptr; }

bool iwb2_run_v(const V2Job &job, cudaStream_t st) { return job.p.depth16 ? iwb2_run_v16(job, st) : iwb2_run_v8(job, st); }

static bool run_h_any(const H2Job &job, cudaStream_t st, int *shift_only) {
  switch (job.epi) {
    case 1: return iwb2_run_h_e1(job, st, shift_only);
    case 2: return iwb2_run_h_e2(job, st, shift_only);
    case 3: return iwb2_run_h_e3(job, st, shift_only);
    case 4: return iwb2_run_h_e4(job, st, shift_only);
  }
  return false;
}
bool iwb2_run_h(const H2Job &job, cudaStream_t st) { return run_h_any(job, st, nullptr); }
int iwb2_h_shift(const H2Job &job) {
  int sh = 0;
  return run_h_any(job, 0, &sh) ? sh : 0;
}

// Test hook (tests/test_planner.py): the host-side split of a launch into uniform / general parts, without touching
// the device.  out[] = {step, n_parts, n_u, u_lo, u_hi, u_len, g_len, u_ref_o, u_ref_last}.
extern "C" __attribute__((visibility("default"))) int iwb200_debug_plan_parts(const int *first, const int *last, int n_out, int lo, int hi, int base,
                                                                             int align, int S, int want, int len_override, int g_min,
                                                                             int step_wanted, int *out) {
  Part2 pt{};
  int nparts = 0;
  const int step = iwb2_plan_parts(first, last, n_out, lo, hi, base, align, S, want, len_override, g_min, step_wanted, &pt, &nparts);
  out[0] = step; out[1] = nparts; out[2] = pt.n_u; out[3] = pt.u_lo; out[4] = pt.u_hi; out[5] = pt.u_len; out[6] = pt.g_len;
  out[7] = pt.u_ref_o; out[8] = pt.u_ref_last;
  return 1;
}
