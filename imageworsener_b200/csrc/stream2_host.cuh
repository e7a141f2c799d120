// stream2_host.cuh -- host-side planning shared by the stream2 translation units: tensor maps, occupancy, and the
// split of a launch into uniform and general parts (see Part2 in stream2.cuh).
#pragma once
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "stream2_dev.cuh"

// (slots, uniform step) pairs the kernels are instantiated for; a cycle is S*STEP inputs: a multiple of 4 (the
// horizontal kernel reads 16-byte chunks, and a stage must end on one for 1 and 3 channels alike), 4..IWB2_MAX_GU
#ifdef IWB_QUICK
#define IWB2_UNIFORM_PAIRS(X) X(6, 2)
#define IWB2_GENERAL_S(X) X(6) X(12)
#else
#define IWB2_UNIFORM_PAIRS(X) \
  X(2, 2) X(4, 2) X(6, 2) X(8, 2) X(4, 3) X(2, 4) X(3, 4) X(4, 4) X(4, 1) X(8, 1)
#define IWB2_GENERAL_S(X) X(2) X(3) X(4) X(5) X(6) X(8) X(12) X(16)
#endif

static inline bool iwb2_pair_built(int S, int step) {
#define X(a, b) if (S == a && step == b) return true;
  IWB2_UNIFORM_PAIRS(X)
#undef X
  return false;
}
static inline bool iwb2_s_built(int S) {
#define X(a) if (S == a) return true;
  IWB2_GENERAL_S(X)
#undef X
  return false;
}

typedef CUresult (*iwb2_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline iwb2_encode_fn iwb2_tensor_map_encoder() {        // driver entry point without linking libcuda
  // (a function-local static with an initialiser: contexts run concurrently from several host threads, and the first
  // use must not be seen half done by a second thread)
  static const iwb2_encode_fn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      return (iwb2_encode_fn)f;
    return (iwb2_encode_fn) nullptr;
  }();
  return fn;
}

static inline int iwb2_sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!sms[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = v > 0 ? v : 148;
  }
  return sms[dev];
}

// Resident CTAs per SM of a kernel at a given dynamic shared-memory size (also raises the kernel's limit).  Cached
// per (device, kernel, size): the query costs microseconds of host time, which matters for small images.
template <class K>
static int iwb2_ctas_per_sm(K kernel, int threads, size_t smem) {
  struct entry { const void *fn; int dev; size_t smem; int n; };
  static std::mutex mu;
  static std::vector<entry> cache;
  const void *fn = (const void *)kernel;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  for (const entry &e : cache)
    if (e.fn == fn && e.dev == dev && e.smem == smem) return e.n;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess) n = 1;
  n = n < 1 ? 1 : n;
  cache.push_back({fn, dev, smem, n});
  return n;
}

// Split outputs [lo, hi) into parts.  `first` / `last` = host copies of the schedule's windows, `base`/`align`: part
// boundaries must be base + k*align, `want` = how many parts make one full wave.  An output o may be emitted by the
// uniform path if (with GU = S*step, the inputs of one cycle)
//   (1) its window ends exactly `step` inputs after its predecessor's (the progression E(o) the kernels assume);
//   (2) its slot is free in time: the kernels restart slot o mod S after input E(o) - GU, so the output that used
//       the slot before (o - S) must end at or before that input, and o's own window must begin after it.
// Inside a long progression (2) follows from (1); it matters for the first S outputs after a break in the
// progression (non-integer factors close to an integer) and at the image border.  (The weights themselves come
// from the schedule table: they are not bitwise periodic in general.)  Returns the uniform step (0: none).
static inline int iwb2_plan_parts(const int *first, const int *last, int n_out, int lo, int hi, int base, int align, int S, int want,
                                  int len_override, int g_min, int step_wanted, Part2 *pt, int *nparts) {
  auto up = [&](int v) { return base + ((v - base + align - 1) / align) * align; };
  auto down = [&](int v) { return base + ((v - base) / align) * align; };
  int step = 0, a = lo, b = lo;
  const int mid = lo + (hi - lo) / 2;
  if (step_wanted > 0 && mid + 1 < n_out && mid >= 1 && last[mid + 1] - last[mid] == step_wanted && iwb2_pair_built(S, step_wanted)) {
    step = step_wanted;
    const int GU = S * step;
    auto slot_free = [&](int o) { return (o - S < 0 || last[o - S] <= last[o] - GU) && first[o] > last[o] - GU; };
    if (!slot_free(mid)) {
      step = 0;
    } else {
      a = mid;
      while (a - 1 >= 0 && last[a] - last[a - 1] == step && slot_free(a - 1)) a--;
      b = mid + 1;
      while (b < n_out && last[b] - last[b - 1] == step && slot_free(b)) b++;
      pt->u_ref_o = mid;                                         // any output of the range anchors the progression
      pt->u_ref_last = last[mid];
    }
  }
  int u_lo = lo, u_hi = lo;
  if (step) {
    u_lo = up(std::max(lo, a));
    u_hi = std::min(hi, b);
    if (u_hi < hi) u_hi = down(u_hi);
    if (u_hi - u_lo < std::max(4 * align, 4 * S)) { step = 0; u_lo = u_hi = lo; }
  }
  want = std::max(1, want);
  // Every part re-reads the inputs of about S outputs in front of it (the windows still open at its start).  When
  // the launch fills the GPU that halo is pure overhead and parts are as long as one wave allows; a small image
  // cannot fill the GPU anyway, and then shorter parts (down to 2*S outputs) shorten the critical path instead.
  (void)g_min;
  g_min = std::max(align, ((2 * S + align - 1) / align) * align);
  if (!step) { pt->u_ref_o = 0; pt->u_ref_last = 0; }
  if (step) {
    // what is left outside the uniform range: parts of at least g_min outputs, at most about a tenth of the wave
    const int rest = (u_lo - lo) + (hi - u_hi);
    int glen = (rest + std::max(1, want / 10) - 1) / std::max(1, want / 10);
    pt->g_len = std::max(std::max(g_min, align), ((glen + align - 1) / align) * align);
    const int n_g = (u_lo - lo + pt->g_len - 1) / pt->g_len + (hi - u_hi + pt->g_len - 1) / pt->g_len;
    // The parts outside the uniform range are a few outputs each (the image border): a CTA slot given to one of them
    // idles for the rest of the launch, so the UNIFORM parts alone fill the wave and the short ones run behind it (they
    // come last in the launch order).  Measured on cfg2: with the border parts counted into the wave 90 of the
    // vertical kernel's 585 CTAs were idle slots (2.83 of 4 warps per scheduler resident on average).
    const int n_u_want = IWB2_PARTS_FULL ? want : std::max(1, want - n_g);
    int len = len_override > 0 ? len_override : std::max(g_min, (u_hi - u_lo + n_u_want - 1) / n_u_want);
    len = std::max(align, ((len + align - 1) / align) * align);
    pt->u_lo = u_lo; pt->u_hi = u_hi; pt->u_len = len;
    pt->n_u = (u_hi - u_lo + len - 1) / len;
    *nparts = pt->n_u + n_g;
  } else {
    int len = len_override > 0 ? len_override : (hi - lo + want - 1) / want;
    len = std::max(std::max(g_min, align), ((len + align - 1) / align) * align);
    pt->n_u = 0; pt->u_lo = pt->u_hi = lo; pt->u_len = align;
    pt->g_len = len;
    *nparts = (hi - lo + len - 1) / len;
  }
  return step;
}

static inline size_t iwb2_v_smem(int S, int step, int nout, bool d16, int lut_ns = 0) {
  const int SS = (S + 1) & ~1;
  const size_t tw = d16 ? 256 : 128;
  const size_t stage = ((size_t)iwb2_v_g(step, d16) * tw + (size_t)iwb2_v_g(step, d16) * SS * 8 + 127) & ~(size_t)127;
  const int gu = step > 0 ? S * step : 4;
  const size_t stage_u = ((size_t)gu * tw + (size_t)gu * SS * 8 + 127) & ~(size_t)127;
  const size_t ring = std::max((size_t)iwb2_v_nst(d16) * stage, (size_t)IWB2_U_NST * stage_u);
  const size_t warp_bytes = (ring + 8 * 8 + 127) & ~(size_t)127;
  return (d16 ? (size_t)lut_ns * 8 : 256 * 16 * 8) + iwb2_v_warps(S, d16) * warp_bytes + (size_t)((nout + 1) & ~1) * 4 + 16 + 16;
}
// 16-bit input, 12-warp layout: how many table entries fit beside the rings (a multiple of 16; the launch may use up to
// 220 KB of dynamic shared memory, see iwb2_ctas_per_sm)
static inline int iwb2_v16_table_entries(int S, int step, int nout) {
  if (!iwb2_v16_shared_table(S, true)) return 0;
  const size_t base = iwb2_v_smem(S, step, nout, true, 0);
  const size_t budget = (size_t)IWB2_V16_BUDGET_KB * 1024;
  if (base + 16 * 8 > budget) return 0;
  return (int)std::min<size_t>(65536, ((budget - base) / 8) & ~(size_t)15);
}
static inline size_t iwb2_h_smem(int S, int step, int nch, int ntab, int nout) {
  const int SS = (S + 1) & ~1;
  const size_t stage = ((size_t)32 * iwb2_tile_pitch(iwb2_h_kp(step) * nch + 3) * 4 + (size_t)iwb2_h_kp(step) * SS * 8 + 127) & ~(size_t)127;
  const int gu = step > 0 ? S * step : 4;
  const size_t stage_u = ((size_t)32 * iwb2_tile_pitch(gu * nch) * 4 + (size_t)gu * SS * 8 + 127) & ~(size_t)127;
  const size_t ring = std::max((size_t)iwb2_h_nst(step) * stage, (size_t)IWB2_U_NST * stage_u);
  const size_t warp_bytes = (ring + 32 * IWB2_OST_PITCH + 8 * 8 + 127) & ~(size_t)127;
  return IWB2_H_WARPS * warp_bytes + (size_t)ntab * IWB2_TAB_BYTES + 64 * 4 + 16 + (size_t)nout * 4 + 16;
}

// The FP32 intermediate of a launch as a 2-D tensor [rows][ncols] (row pitch in floats), tiles of 32 rows x box_cols.
static inline bool iwb2_make_inter_map(const H2Params &p, int box_cols, CUtensorMap *tm) {
  iwb2_encode_fn enc = iwb2_tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)(p.ncols + p.sh), (cuuint64_t)(p.row_hi - p.row_lo)};
  const cuuint64_t gstride[1] = {(cuuint64_t)p.pitch * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(p.inter), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Uniform horizontal stages must start on 16 bytes: with window ends E(o) = u_ref_last + step*(o - u_ref_o) the first
// input pixel of the cycle that starts at output og is E(og) - step + 1, and all cycles of a launch are S*step pixels
// apart (a multiple of 4), so one shift per launch aligns them all.
static inline int iwb2_shift_for(const Part2 &pt, int S, int step, int nch) {
  if (!step || pt.n_u == 0) return 0;
  int og0 = pt.u_lo - (S - 1);
  og0 = (og0 >= 0 ? og0 / S : -((-og0 + S - 1) / S)) * S;
  const long B0 = (long)pt.u_ref_last + (long)step * (og0 - pt.u_ref_o) - step + 1;
  const long m = ((B0 * nch) % 4 + 4) % 4;
  return (int)((4 - m) % 4);
}

bool iwb2_run_v8(const V2Job &job, cudaStream_t st);
bool iwb2_run_v16(const V2Job &job, cudaStream_t st);
bool iwb2_run_h_e1(const H2Job &job, cudaStream_t st, int *shift_only);
bool iwb2_run_h_e2(const H2Job &job, cudaStream_t st, int *shift_only);
bool iwb2_run_h_e3(const H2Job &job, cudaStream_t st, int *shift_only);
bool iwb2_run_h_e4(const H2Job &job, cudaStream_t st, int *shift_only);
