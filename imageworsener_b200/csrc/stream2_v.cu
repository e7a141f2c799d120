// stream2_v.cu -- vertical streaming kernel: instantiations, planning, launch (design: stream2.cuh).
#include "stream2_host.cuh"

bool iwb2_tensor_maps_available() { return iwb2_tensor_map_encoder() != nullptr; }

bool iwb2_supported(int SV, int SH, int nch, int epi) {
  return iwb2_s_built(SV) && iwb2_s_built(SH) && (nch == 1 || nch == 3) && (epi == 1 || epi == 2 || epi == 3 || epi == 4);
}

// The input as a 2-D byte tensor: [tm_rows][ncols] with the caller's row stride, tiles of `box_rows` rows x 128 columns.
static bool make_input_map(const V2Params &p, int box_rows, CUtensorMap *tm) {
  iwb2_encode_fn enc = iwb2_tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)p.ncols, (cuuint64_t)p.tm_rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)p.bpr};
  const cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const void *base = p.src + (size_t)p.tm_row0 * p.bpr;
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int S, int STEP, int SH>
static bool run_v_sh(const V2Job &job, int step_found, cudaStream_t st) {
  V2Params p = job.p;
  const int threads = 32 * IWB2_V_WARPS;
  const int colblocks = (p.ncols + 128 * IWB2_V_WARPS - 1) / (128 * IWB2_V_WARPS);
  const int slots = iwb2_sm_count() * iwb2_ctas_per_sm(iwb_v2<S, STEP, SH>, threads, iwb2_v_smem(S, STEP, 64 + S));
  int nparts = 0;
  iwb2_plan_parts(job.first_host, job.last_host, p.sc.n_out, p.row_lo, p.row_hi, p.row_base, 1, S, slots / colblocks, job.strip_override, 16,
                  step_found == STEP ? STEP : 0, &p.part, &nparts);
  const int nout = std::max(p.part.g_len, 8) + S;
  CUtensorMap tm, tmu;
  if (!make_input_map(p, iwb2_v_g(STEP), &tm)) return false;
  if (!make_input_map(p, STEP > 0 ? S * STEP : iwb2_v_g(STEP), &tmu)) return false;
  iwb_v2<S, STEP, SH><<<dim3((unsigned)colblocks, (unsigned)nparts), threads, iwb2_v_smem(S, STEP, nout), st>>>(tm, tmu, p);
  return true;
}
template <int S, int STEP>
static bool run_v(const V2Job &job, int step_found, cudaStream_t st) {
  switch (job.p.sh & 3) {                                       // the column shift is a compile-time constant of the kernel
    case 0: return run_v_sh<S, STEP, 0>(job, step_found, st);
    case 1: return run_v_sh<S, STEP, 1>(job, step_found, st);
    case 2: return run_v_sh<S, STEP, 2>(job, step_found, st);
  }
  return run_v_sh<S, STEP, 3>(job, step_found, st);
}

bool iwb2_run_v(const V2Job &job, cudaStream_t st) {
  // the uniform step of this launch's middle decides the instantiation; the planner re-derives the exact range
  int step = 0;
  const V2Params &p = job.p;
  const int mid = p.row_lo + (p.row_hi - p.row_lo) / 2;
  if (job.allow_uniform && mid + 1 < p.sc.n_out) step = job.last_host[mid + 1] - job.last_host[mid];
#define X(a, b) if (job.S == a && step == b) return run_v<a, b>(job, step, st);
  IWB2_UNIFORM_PAIRS(X)
#undef X
#define X(a) if (job.S == a) return run_v<a, 0>(job, 0, st);
  IWB2_GENERAL_S(X)
#undef X
  return false;
}

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
