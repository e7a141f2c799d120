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

// ---- diagnostics of a launch that trapped (stream3_dev.cuh: s3_die) ----
static int *g_dbg_host[64] = {nullptr};
int *iwb3_debug_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!g_dbg_host[dev]) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!g_dbg_host[dev]) {
      int *p = nullptr;
      if (cudaHostAlloc((void **)&p, 160 * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      memset(p, 0, 160 * sizeof(int));
      g_dbg_host[dev] = p;
    }
  }
  return g_dbg_host[dev];                                       // (unified addressing: the host pointer is valid on the device)
}
std::string iwb3_debug_text() {
  std::string out;
  for (int dev = 0; dev < 64; dev++) {
    int *p = g_dbg_host[dev];
    if (!p || !p[0]) continue;
    static const char *what[] = {"?", "V uniform first stage", "V uniform stage", "V general first stage", "V general stage", "H uniform first stage",
                                 "H uniform stage", "H general first stage", "H general stage", "H group counter", "CTA tables"};
    char line[256];
    snprintf(line, sizeof line, "fused kernel: %d warp(s) gave up waiting on device %d;", p[0], dev);
    out += line;
    for (int k = 0; k < std::min(p[0], 15); k++) {
      const int *r = p + 8 + 8 * k;
      snprintf(line, sizeof line, " [%s: cta %d warp %d a=%d b=%d c=%d d=%d]", what[r[0] >= 0 && r[0] <= 10 ? r[0] : 0], r[1], r[2], r[3], r[4], r[5], r[6]);
      out += line;
    }
    memset(p, 0, 160 * sizeof(int));
  }
  return out;
}
