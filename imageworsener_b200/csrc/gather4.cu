// gather4.cu -- gather-form streaming kernels: instantiations, tensor maps, launch (design: gather4.cuh).
#include <cmath>

#include "gather4_dev.cuh"
#include "stream2_host.cuh"

namespace {

bool g4_input_map(const GVParams &p, bool d16, CUtensorMap *tm) {
  const int bpc = d16 ? 2 : 1;
  iwb2_encode_fn enc = iwb2_tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)p.ncols * bpc, (cuuint64_t)p.tm_rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)p.bpr};
  const cuuint32_t box[2] = {(cuuint32_t)(128 * bpc), (cuuint32_t)IWG_V_G};
  const cuuint32_t estr[2] = {1, 1};
  const void *base = p.src + (size_t)p.tm_row0 * p.bpr;
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool g4_inter_map(const GHParams &p, int box_cols, CUtensorMap *tm) {
  iwb2_encode_fn enc = iwb2_tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)p.ncols, (cuuint64_t)(p.row_hi - p.row_lo)};
  const cuuint64_t gstride[1] = {(cuuint64_t)p.pitch * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(p.inter), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// How many parts to cut one dimension into, given the CTAs the other dimension contributes per part and the CTAs the
// device holds at once: about `target_waves` waves, and among the candidates the count whose last wave is fullest.
static int g4_pick_parts(int slots, int per_part, int total_len, int min_len, double target_waves) {
  const int nmax = std::max(1, total_len / std::max(1, min_len));            // parts of at least min_len
  const int want = std::min(nmax, std::max(1, (int)(target_waves * slots) / std::max(1, per_part)));
  int best = want;
  double best_eff = -1.0;
  for (int n = std::max(1, want / 2); n <= std::min(nmax, want + want / 2 + 1); n++) {
    const double waves = (double)n * per_part / slots;
    const double eff = waves / std::ceil(waves) - 0.01 * std::abs(n - want) / want;
    if (eff > best_eff) { best_eff = eff; best = n; }
  }
  return best;
}

template <int W, bool D16, int NCHP>
bool run_v(const GVJob &job, cudaStream_t st) {
  GVParams p = job.p;
  const int threads = 32 * IWG_WARPS;
  const int colblocks = (p.ncols + 128 * IWG_WARPS - 1) / (128 * IWG_WARPS);
  const int rows = p.row_hi - p.row_lo;
  const size_t tw = D16 ? 256 : 128;
  const size_t smem_fixed = (D16 ? 0 : 2 * 256 * 8) + (size_t)IWG_WARPS * (IWG_V_NST * IWG_V_G * tw + 128) + 16;
  auto kernel = iwg_v<W, D16, NCHP>;
  // strips of output rows: about two waves of CTAs with a full last wave, at least 64 rows each (a strip re-reads W input rows)
  const int slots = iwb2_ctas_per_sm(kernel, threads, smem_fixed + 4 * 256) * iwb2_sm_count();
  const int nstrips = g4_pick_parts(slots, colblocks, rows, 64, 2.0);
  int strip = std::max(64, (rows + nstrips - 1) / nstrips);
  strip = std::min(strip, 2048);
  p.strip = strip;
  const size_t smem = smem_fixed + (size_t)strip * 4;
  CUtensorMap tm;
  if (!g4_input_map(p, D16, &tm)) return false;
  iwb2_ctas_per_sm(kernel, threads, smem);
  kernel<<<dim3((unsigned)colblocks, (unsigned)((rows + strip - 1) / strip)), threads, smem, st>>>(tm, p);
  return true;
}

template <int W, int NCH, int NCO, int FAST>
bool run_h(const GHJob &job, cudaStream_t st) {
  GHParams p = job.p;
  constexpr int NW = iwg_h_warps(FAST), NST = iwg_h_nst(FAST);
  const int threads = 32 * NW;
  const int groups = (p.row_hi - p.row_lo + 31) / 32;
  const int rowblocks = (groups + NW - 1) / NW;
  constexpr int TP = iwb2_tile_pitch(IWG_H_KP * NCH + 3);
  const size_t stage = ((size_t)32 * TP * 4 + 127) & ~(size_t)127;
  const size_t smem_fixed = (size_t)NW * (NST * stage + 32 * IWG_OST_PITCH + 128) + (FAST == 2 ? IWB_EST_NREC * sizeof(IwbEstRec) : 0) + 16;
  auto kernel = iwg_h<W, NCH, NCO, FAST>;
  // column segments: about 2.5 waves of CTAs with a full last wave
  const int slots = iwb2_ctas_per_sm(kernel, threads, smem_fixed + 4 * 512) * iwb2_sm_count();
  const int best = g4_pick_parts(slots, rowblocks, p.out_w, 64, 2.5);
  int seg = std::max(64, (p.out_w + best - 1) / best);
  seg = std::min((seg + 3) & ~3, 4096);
  p.seg = seg;
  const size_t smem = smem_fixed + (size_t)seg * 4;
  CUtensorMap tm;
  if (!g4_inter_map(p, TP, &tm)) return false;
  iwb2_ctas_per_sm(kernel, threads, smem);
  kernel<<<dim3((unsigned)((p.out_w + seg - 1) / seg), (unsigned)rowblocks), threads, smem, st>>>(tm, p);
  return true;
}

template <int W>
bool run_v_w(const GVJob &job, cudaStream_t st) {
  if (job.p.depth16) return job.nchp == 4 ? run_v<W, true, 4>(job, st) : run_v<W, true, 2>(job, st);
  return job.nchp == 4 ? run_v<W, false, 4>(job, st) : run_v<W, false, 2>(job, st);
}
// the straight-line epilogue's assumptions (gather4_dev.cuh: FAST)
bool fast_epilogue(const GHJob &job) {
  const GHParams &p = job.p;
  if (job.nco != job.nch - 1 || p.cm.checker || p.cm.intclamp || p.out.is_float) return false;
  for (int k = 0; k < job.nco; k++) {
    const DevChanOut &o = p.ch[k];
    if (!(o.qmode == 1 && !o.use_nc && !o.codeval && o.guess_cs >= 0 && o.ntab - 1 >= 8)) return false;
    if (o.dfamily != 0 && o.dfamily != 1 && o.dfamily != 3) return false;
    if (!o.premult || o.bkgd_translucent || o.is_alpha) return false;
    if (o.ntab != p.ch[0].ntab || o.guess_cs != p.ch[0].guess_cs || o.guess_gamma != p.ch[0].guess_gamma || o.dfamily != p.ch[0].dfamily ||
        o.dsubtype != p.ch[0].dsubtype || o.late_bkgd != p.ch[0].late_bkgd) return false;
  }
  return true;
}
template <int W>
bool run_h_w(const GHJob &job, cudaStream_t st) {
  const bool fast = fast_epilogue(job);
  const bool model = fast && job.p.est && job.p.ch[0].dfamily == 0 && job.p.ch[0].late_bkgd && job.p.out.depth == 16;
  if (job.nch == 4) {
    if (job.nco == 4) return run_h<W, 4, 4, 0>(job, st);
    if (job.nco == 3) return model ? run_h<W, 4, 3, 2>(job, st) : (fast ? run_h<W, 4, 3, 1>(job, st) : run_h<W, 4, 3, 0>(job, st));
    return false;
  }
  if (job.nco == 2) return run_h<W, 2, 2, 0>(job, st);
  if (job.nco == 1) return model ? run_h<W, 2, 1, 2>(job, st) : (fast ? run_h<W, 2, 1, 1>(job, st) : run_h<W, 2, 1, 0>(job, st));
  return false;
}

}  // namespace

bool iwg_supported(int WV, int WH, int nch) {
  auto okw = [](int w) { return w == 2 || w == 4 || w == 6 || w == 8; };
  return okw(WV) && okw(WH) && (nch == 2 || nch == 4) && iwb2_tensor_map_encoder() != nullptr;
}
bool iwg_run_v(const GVJob &job, cudaStream_t st) {
  switch (job.W) {
    case 2: return run_v_w<2>(job, st);
    case 4: return run_v_w<4>(job, st);
    case 6: return run_v_w<6>(job, st);
    case 8: return run_v_w<8>(job, st);
  }
  return false;
}
bool iwg_run_h(const GHJob &job, cudaStream_t st) {
  switch (job.W) {
    case 2: return run_h_w<2>(job, st);
    case 4: return run_h_w<4>(job, st);
    case 6: return run_h_w<6>(job, st);
    case 8: return run_h_w<8>(job, st);
  }
  return false;
}
