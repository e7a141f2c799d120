// pipeline.cu -- jobs: turn a prepared context into device tables and launch the kernels.
//
// Data layout in HBM: the input image is kept exactly as the caller stores it (interleaved,
// caller's row stride, big-endian 16-bit); the FP32 intermediate of the vertical pass is one row-major plane per
// channel covering only the current chunk of output rows, sized to stay resident in the 126 MB L2 between the
// vertical pass that writes it and the horizontal pass that reads it; the output is written in its final
// interleaved byte layout.  Chunks are processed in order on one stream.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "iwb_internal.h"
#include "kernels.cuh"
#include "optimize.h"
#if !defined(IWB_EMU)
#include <cuda_runtime.h>
#include "fused.cuh"
#include "stream2.cuh"
#include "stream3.cuh"
#include "gather4.cuh"
#include "gather5.cuh"
#include "errdiff.cuh"
#endif

// ------------------------------------------------------------------------------------------------------
// Thin device layer.  The product build is CUDA only.  -DIWB_EMU (tests/_emu, never shipped, never loaded by the
// package) swaps it for plain host memory + a loop so host logic can be debugged on a machine without a GPU.
// ------------------------------------------------------------------------------------------------------
static std::atomic<unsigned long long> g_launches{0};   // (contexts may run concurrently from several threads)

#if !defined(IWB_EMU)
#include <cuda_runtime.h>
typedef cudaStream_t iwb_stream;
static const char *dev_check(cudaError_t e) { return e == cudaSuccess ? nullptr : cudaGetErrorString(e); }
// Device memory comes from CUDA's stream-ordered pool (cudaMallocAsync on the default stream) with the release
// threshold raised, so that consecutive iw_process_image() calls reuse the same HBM blocks instead of paying
// cudaMalloc/cudaFree (milliseconds, and an implicit device sync) every time.
static void dev_pool_init(int device) {
  static bool done[64] = {false};
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[device] = true;
}
static void *dev_alloc(size_t n, std::string &err) {
  void *p = nullptr;
  if (const char *m = dev_check(cudaMallocAsync(&p, n ? n : 1, 0))) { err = std::string("cudaMallocAsync: ") + m; return nullptr; }
  return p;
}
static void dev_free(void *p) { if (p) cudaFreeAsync(p, 0); }
static bool dev_h2d(void *d, const void *h, size_t n, iwb_stream s, std::string &err) {
  if (const char *m = dev_check(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s))) { err = std::string("H2D copy: ") + m; return false; }
  return true;
}
static bool dev_d2h(void *h, const void *d, size_t n, iwb_stream s, std::string &err) {
  if (const char *m = dev_check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s))) { err = std::string("D2H copy: ") + m; return false; }
  return true;
}
static bool dev_zero(void *d, size_t n, iwb_stream s) { return cudaMemsetAsync(d, 0, n, s) == cudaSuccess; }
static bool dev_sync(iwb_stream s, std::string &err) {
  if (const char *m = dev_check(cudaStreamSynchronize(s))) {
    err = std::string("kernel execution: ") + m;
    const std::string d = iwb3_debug_text();                    // (a fused-kernel warp that gave up waiting says why)
    if (!d.empty()) { err += "; "; err += d; fprintf(stderr, "%s\n", d.c_str()); }
    return false;
  }
  return true;
}
template <class Body, class P>
__global__ void iwb_thread_per_item(const P p, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) Body::run(p, i);
}
template <class Body, class P>
__global__ void iwb_block_per_item(const P p) {
  if (threadIdx.x == 0) Body::run(p, (size_t)blockIdx.x);
}
template <class Body, class P>
static void launch(const P &p, size_t n, iwb_stream s) {
  if (!n) return;
  const int block = 256;
  iwb_thread_per_item<Body, P><<<(unsigned)((n + block - 1) / block), block, 0, s>>>(p, n);
  g_launches++;
}
template <class Body, class P>
static void launch_serial(const P &p, int n, iwb_stream s) {  // n independent sequential chains
  if (!n) return;
  iwb_block_per_item<Body, P><<<n, 32, 0, s>>>(p);
  g_launches++;
}
// The library switches to the job's device for the duration of a call and puts the caller's device back on return
// (a multi-GPU application keeps its own current device).
struct dev_guard {
  int prev = -1;
  explicit dev_guard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    if (device < 0 || prev == device) { prev = -1; return; }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); prev = -1; }   // (reported by dev_select)
  }
  ~dev_guard() { if (prev >= 0) cudaSetDevice(prev); }
};
// Run `f` exactly once per (call site, device), and let nobody past this point before it has finished: kernel
// attributes are per device, and contexts run concurrently from several threads (a launch that overtakes the attribute
// call of another thread fails with "too much shared memory").
static std::mutex g_once_mu;
template <class F>
static void once_per_device(std::atomic<unsigned long long> &mask, F f) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask.load(std::memory_order_acquire) & bit) return;
  std::lock_guard<std::mutex> lk(g_once_mu);
  if (mask.load(std::memory_order_relaxed) & bit) return;
  f();
  mask.fetch_or(bit, std::memory_order_release);
}
static bool dev_select(int device, std::string &err) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
    err = "no CUDA device available (this library has no CPU path)";
    return false;
  }
  if (device < 0 || device >= count) { err = "invalid CUDA device ordinal"; return false; }
  if (const char *m = dev_check(cudaSetDevice(device))) { err = std::string("cudaSetDevice: ") + m; return false; }
  dev_pool_init(device);
  return true;
}
#else
typedef void *iwb_stream;
static void *dev_alloc(size_t n, std::string &) { return malloc(n ? n : 1); }
static void dev_free(void *p) { free(p); }
static bool dev_h2d(void *d, const void *h, size_t n, iwb_stream, std::string &) { memcpy(d, h, n); return true; }
static bool dev_d2h(void *h, const void *d, size_t n, iwb_stream, std::string &) { memcpy(h, d, n); return true; }
static bool dev_zero(void *d, size_t n, iwb_stream) { memset(d, 0, n); return true; }
static bool dev_sync(iwb_stream, std::string &) { return true; }
template <class Body, class P>
static void launch(const P &p, size_t n, iwb_stream) { for (size_t i = 0; i < n; i++) Body::run(p, i); g_launches++; }
template <class Body, class P>
static void launch_serial(const P &p, int n, iwb_stream) { for (int i = 0; i < n; i++) Body::run(p, (size_t)i); g_launches++; }
static bool dev_select(int, std::string &) { return true; }
struct dev_guard { explicit dev_guard(int) {} };
#endif

bool iwb_bkgd_is_translucent(const iw_context *ctx);  // plan.cpp

// ------------------------------------------------------------------------------------------------------
// Pinned host pool behind iw_malloc_large()/iw_free() (api.cpp).  Image buffers handed to and returned by the
// library are page-locked and recycled across contexts, so H2D/D2H run as DMA at PCIe speed and no
// cudaHostAlloc is paid per image.  Only used with the default allocator; disable with IWB200_PINNED=0.
// ------------------------------------------------------------------------------------------------------
#include <map>
#include <mutex>
static std::mutex g_pin_mu;
static std::multimap<size_t, void *> g_pin_free;   // size -> block
static std::map<void *, size_t> g_pin_all;         // every live pinned block -> size
static size_t g_pin_cached = 0;

void *iwb_pinned_alloc(size_t n) {
#if defined(IWB_EMU)
  (void)n;
  return nullptr;
#else
  static int enabled = -1;
  if (enabled < 0) {
    const char *e = getenv("IWB200_PINNED");
    int count = 0;
    enabled = !(e && e[0] == '0') && cudaGetDeviceCount(&count) == cudaSuccess && count > 0;
  }
  if (!enabled) return nullptr;
  size_t want = (n + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto it = g_pin_free.lower_bound(want);
    if (it != g_pin_free.end() && it->first <= want + want / 4) {
      void *p = it->second;
      g_pin_cached -= it->first;
      g_pin_free.erase(it);
      return p;
    }
  }
  void *p = nullptr;
  if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lk(g_pin_mu);
  g_pin_all[p] = want;
  return p;
#endif
}

bool iwb_pinned_free(void *p) {
#if defined(IWB_EMU)
  (void)p;
  return false;
#else
  std::lock_guard<std::mutex> lk(g_pin_mu);
  auto it = g_pin_all.find(p);
  if (it == g_pin_all.end()) return false;
  if (g_pin_cached + it->second > ((size_t)16 << 30)) {  // keep at most 16 GiB page-locked and idle
    cudaFreeHost(p);
    g_pin_all.erase(it);
    return true;
  }
  g_pin_free.emplace(it->second, p);
  g_pin_cached += it->second;
  return true;
#endif
}

// ------------------------------------------------------------------------------------------------------
struct iwb200_job {
  iw_context *ctx = nullptr;
  int device = 0;
  std::string err;
  std::vector<void *> allocs;

  int nproc = 0;             // intermediate channels in processing order (alpha first)
  int mid_index[4] = {0, 0, 0, 0};
  DevChanIn cin[4];
  DevChanOut cout[4];
  DevAxis vax[4], hax[4];
  std::vector<iwb_axis_table> vtab, htab;   // host copies (band planning)
  int vtab_of[4] = {0, 0, 0, 0};
  DevCommon cm{};
  DevInput in{};
  DevOutput out{};
  int in_w = 0, in_h = 0;    // cropped logical input size = axis inputs
  int phys_w = 0, phys_h = 0;
  int has_alpha = 0, any_random = 0, any_errdiff = 0;
  int negate = 0, clear_transparent = 0, out_alpha_ch = -1;
  iwb_affine_t *d_jump = nullptr;
  double errdiff_m[4][12];
  iwb200_job_info info{};
  int building_channel = 0;
  std::map<const double *, std::vector<double>> lut_host;   // 8-bit input tables by device address
  std::vector<double> thr_host[4];
  std::vector<double> nc_host;              // nearest-colour midpoints (when that table exists)          // per processed channel: level thresholds (table-mode channels)

  // scratch, sized for `chunk_rows` output rows
  int chunk_rows = 0;
  float *inter[4] = {nullptr, nullptr, nullptr, nullptr};
  float *final_alpha = nullptr;
  double *lin_plane[4] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t *draws[4] = {nullptr, nullptr, nullptr, nullptr};
  double *errrows[4] = {nullptr, nullptr, nullptr, nullptr};
  int errdiff_next_row = 0;
  // where the FP32 intermediate of the last chunk lives (iwb200_job_read_intermediate)
  int last_path = -1, last_c0 = 0, last_c1 = 0, last_sh = 0;
  int last_gen = 0;                 // kernel generation of the last launch (0 generic, 1 / 2 streaming, 3 fused)
  int x0_override = -1;             // set by run_host while its device copy holds only the cropped columns
  DevChanOut *ed_ch_g = nullptr;    // error-diffusion kernel: channel descriptions in global memory
  // post-pass optimiser (optimize.h): scan result, rewrite recipe and rewritten band in device memory
  IwbOptScan *d_scan = nullptr;
  OptRecipe *d_recipe = nullptr;
  void *d_opt = nullptr; size_t d_opt_bytes = 0;
  iwb_timings *tm_sink = nullptr;   // where run_host reports its phase times (default: the context's record)
  int float_prepass = 0;            // float input in a non-linear colourspace: run_host linearises it on the host first
  int prepass_done = 0;             // (set by run_host around its run_device calls)
  size_t last_pitch = 0;
  const float *last_inter = nullptr;
  // optional per-kernel event timing
  int profiling = 0;
  struct timed { int which; void *a, *b; };
  std::vector<timed> pending;
  int stat_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double stat_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::string key;            // plan-cache key (empty: not cacheable)
  int reused = 0, used_external_stream = 0;
  void *streams[3] = {nullptr, nullptr, nullptr};   // run_host pipeline: copy-in, compute, copy-out
  std::vector<void *> events;
  // staging for run_host
  void *d_in = nullptr; size_t d_in_bytes = 0;
  void *d_out = nullptr; size_t d_out_bytes = 0;
#if !defined(IWB_EMU)
  FastPlan fused;
  // gather-form streaming kernels for images with alpha (gather4.cuh)
  struct gather_plan {
    int enabled = 0;
    int WV = 0, WH = 0, nch = 0, depth16 = 0;
    GSched sv{}, sh{};
    std::vector<int> vfirst, vlast;               // host copies (band planning)
    const double *lut_c = nullptr, *lut_a = nullptr;
    DevChanOut ch[4];
    const DevChanOut *ch_g = nullptr;             // ch[] in device memory
    int nco = 0, out_src[4] = {0, 0, 0, 0};
    int chan_of_proc[4] = {0, 0, 0, 0};           // interleaved channel of processed channel p
    float *inter = nullptr;
    int chunk_rows = 0;
    size_t pitch = 0;
    const IwbEstRec *est = nullptr;               // level model of the colour channels' table (estimate.h), 16-bit output
    double est_margin = 0.0;
    int tile = 0;                                 // windows longer than gather4's register windows: the tile kernels (gather5.cu)
    int tile_max_nb = 0;
    const DevCommon *cm_g = nullptr;
    // images without alpha: the vertical pass runs on the scatter-form kernel of stream2.cuh (one table lookup per
    // sample instead of one per tap) and feeds the tile horizontal kernel
    int v2 = 0, v2_S = 0;
    std::vector<int> v2first, v2last;
    const double *lut16 = nullptr;
  } gather;
  std::vector<std::unique_ptr<VH3State>> vh3;    // fused-kernel state per row range (stream3.cuh)
  int gen3 = 0;                                 // the fused kernel is opt-in (cuda:gen=3): 0 never, 2 whenever supported
  int gen3_nv = 0, gen3_seg = 0;
  struct s2_tables { Sched2 v, h; Quant2 q[4]; };
  std::unique_ptr<s2_tables> s2_store;
#endif

  // All plan tables (weight lists, colour tables, schedules) are packed into a few arena chunks and copied
  // to the device with one H2D per chunk (flush_tables) instead of one allocation + copy per table.
  struct chunk { std::vector<uint8_t> host; uint8_t *dev; size_t used, cap; };
  std::vector<chunk> chunks;
  template <class T>
  T *upload(const T *h, size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    if (chunks.empty() || chunks.back().used + bytes > chunks.back().cap) {
      chunk c;
      c.cap = bytes > ((size_t)4 << 20) ? bytes : ((size_t)4 << 20);
      c.dev = (uint8_t *)dev_alloc(c.cap, err);
      if (!c.dev) return nullptr;
      allocs.push_back(c.dev);
      c.host.resize(c.cap);
      c.used = 0;
      chunks.push_back(std::move(c));
    }
    chunk &c = chunks.back();
    if (count) memcpy(c.host.data() + c.used, h, count * sizeof(T));
    T *d = (T *)(c.dev + c.used);
    c.used += bytes;
    return d;
  }
  bool flush_tables() {
    for (chunk &c : chunks) {
      if (c.used && !dev_h2d(c.dev, c.host.data(), c.used, 0, err)) return false;
      // pageable source: the copy is staged before cudaMemcpyAsync returns, so the host copy can go
      std::vector<uint8_t>().swap(c.host);
      c.cap = c.used;  // closed
    }
    s0_dirty = true;
    return true;
  }
  void *scratch(size_t bytes) {
    void *d = dev_alloc(bytes, err);
    if (d) allocs.push_back(d);
    s0_dirty = true;
    return d;
  }
  // Tables are uploaded and scratch is allocated on the default stream (stream-ordered pool).  A caller's own
  // non-blocking stream does not synchronise with it implicitly: before the first launch that follows such work the
  // stream is ordered after it with an event (once per job and stream, not per call).
  bool s0_dirty = true;
  void *s0_event = nullptr, *s0_stream = nullptr;
  void order_after_stream0(iwb_stream st) {
#if !defined(IWB_EMU)
    if (!st) return;                                            // the default stream itself
    if (!s0_dirty && s0_stream == (void *)st) return;
    if (!s0_event) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); s0_event = e; }
    cudaEventRecord((cudaEvent_t)s0_event, 0);
    cudaStreamWaitEvent((cudaStream_t)st, (cudaEvent_t)s0_event, 0);
    s0_dirty = false;
    s0_stream = (void *)st;
#else
    (void)st;
#endif
  }
};

namespace {

bool fail(iwb200_job *job, const std::string &m) {
  if (job->err.empty()) job->err = m;
  return false;
}

// Bracket one launch with events when profiling is on.
struct prof_scope {
  iwb200_job *job; int which; iwb_stream s; void *a = nullptr;
  prof_scope(iwb200_job *j, int w, iwb_stream st) : job(j), which(w), s(st) {
#if !defined(IWB_EMU)
    if (job->profiling) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); a = e; }
#endif
  }
  ~prof_scope() {
#if !defined(IWB_EMU)
    if (job->profiling && a) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); job->pending.push_back({which, a, (void *)e}); }
#endif
  }
};

bool upload_axis(iwb200_job *job, const iwb_axis_table &t, DevAxis *d) {
  d->mode = t.mode;
  d->n_in = t.n_in;
  d->n_out = t.n_out;
  d->edge_value = t.edge_value;
  d->first = d->src = d->map = nullptr;
  d->w = nullptr;
  if (t.mode == 1) {
    d->map = job->upload(t.map.data(), t.map.size());
    return d->map != nullptr;
  }
  d->first = job->upload(t.first.data(), t.first.size());
  d->src = job->upload(t.src.data(), t.src.size());
  d->w = job->upload(t.w.data(), t.w.size());
  return d->first && d->src && d->w;
}

// code -> linear table for one physical input channel (see kernels.cuh one_linear)
struct lut_key { int maxcode, cstype; double gamma; const double *dev; };

const double *input_lut(iwb200_job *job, std::vector<lut_key> &cache, int depth, int maxcode, const iw_csdescr &cs) {
  for (auto &k : cache)
    if (k.maxcode == maxcode && k.cstype == cs.cstype && (cs.cstype != IW_CSTYPE_GAMMA || k.gamma == cs.gamma)) return k.dev;
  const int n = 1 << depth;
  std::vector<double> t(n);
  const double mx = (double)maxcode;
  for (int i = 0; i < n; i++) t[i] = iwb_to_linear(((double)i) / mx, &cs);
  const double *d = job->upload(t.data(), t.size());
  cache.push_back({maxcode, cs.cstype, cs.gamma, d});
  if (n == 256) job->lut_host[d] = t;                            // host copy: the streaming kernels want it replicated
  return d;
}

const double kErrDiff[8][12] = {  // x 0 1 / 2 3 4 5 6 / 7 8 9 10 11  (reference: src/imagew-main.c:506-535)
    {7.0 / 16, 0.0, 0.0, 3.0 / 16, 5.0 / 16, 1.0 / 16, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {7.0 / 48, 5.0 / 48, 3.0 / 48, 5.0 / 48, 7.0 / 48, 5.0 / 48, 3.0 / 48, 1.0 / 48, 3.0 / 48, 5.0 / 48, 3.0 / 48, 1.0 / 48},
    {8.0 / 42, 4.0 / 42, 2.0 / 42, 4.0 / 42, 8.0 / 42, 4.0 / 42, 2.0 / 42, 1.0 / 42, 2.0 / 42, 4.0 / 42, 2.0 / 42, 1.0 / 42},
    {8.0 / 32, 4.0 / 32, 2.0 / 32, 4.0 / 32, 8.0 / 32, 4.0 / 32, 2.0 / 32, 0.0, 0.0, 0.0, 0.0, 0.0},
    {5.0 / 32, 3.0 / 32, 2.0 / 32, 4.0 / 32, 5.0 / 32, 4.0 / 32, 2.0 / 32, 0.0, 2.0 / 32, 3.0 / 32, 2.0 / 32, 0.0},
    {4.0 / 16, 3.0 / 16, 1.0 / 16, 2.0 / 16, 3.0 / 16, 2.0 / 16, 1.0 / 16, 0.0, 0.0, 0.0, 0.0, 0.0},
    {2.0 / 4, 0.0, 0.0, 1.0 / 4, 1.0 / 4, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {1.0 / 8, 1.0 / 8, 0.0, 1.0 / 8, 1.0 / 8, 1.0 / 8, 0.0, 0.0, 0.0, 1.0 / 8, 0.0, 0.0}};

bool build_output_tables(iwb200_job *job, const iw_context *ctx, const iwb_out_channel &oc, const iw_csdescr &cs,
                         DevChanOut *o) {
  // rev[k] = linear value of stored code k (reference: cvt_int_sample_to_linear_output, main.c:330-344)
  const int maxcode = oc.maxcode;
  const double mx = (double)maxcode;
  std::vector<double> thr, code, lin;
  if (oc.color_count == 0) {
    thr.resize(maxcode + 1);
    code.resize(maxcode + 1);
    for (int k = 0; k <= maxcode; k++) {
      thr[k] = iwb_to_linear(((double)k) / mx, &cs);
      code[k] = (double)k;
    }
    lin = thr;
  } else {
    // posterised levels q = 0..cc-1 (reference: get_nearest_valid_colors, main.c:596-616)
    const int cc = oc.color_count;
    const double pm = (double)(cc - 1);
    thr.resize(cc); code.resize(cc); lin.resize(cc);
    for (int q = 0; q < cc; q++) {
      thr[q] = iwb_to_linear(((double)q) / pm, &cs);
      code[q] = floor(0.5000000001 + ((double)q) * (mx / pm));
      lin[q] = iwb_to_linear(((double)(unsigned int)code[q]) / mx, &cs);
    }
  }
  o->ntab = (int)thr.size();
  // levels of a full-depth table follow the target curve: the device may start its search from an FP32 estimate
  o->guess_cs = oc.color_count == 0 ? (int)cs.cstype : -1;
  o->guess_gamma = (cs.cstype == IW_CSTYPE_GAMMA && cs.gamma > 0.0) ? (float)(1.0 / cs.gamma) : 1.0f;
  job->thr_host[job->building_channel] = thr;
  o->thr = job->upload(thr.data(), thr.size());
  if (oc.color_count == 0) {                                     // code q is level q and its linear value is the level's bound
    o->codeval = nullptr;
    o->linval = o->thr;
  } else {
    o->codeval = job->upload(code.data(), code.size());
    o->linval = job->upload(lin.data(), lin.size());
    if (!o->codeval) return false;
  }
  (void)ctx;
  return o->thr && o->linval;
}

}  // namespace


// ------------------------------------------------------------------------------------------------------
// Fast path, host side: streaming schedules, eligibility, launches (kernels in fused.cuh)
// ------------------------------------------------------------------------------------------------------
#if !defined(IWB_EMU)
namespace {

struct HostSched {
  int S = 0, max_emit = 0;
  std::vector<double> w;
  std::vector<int> emit_lo, emit_n, first, last;
};

// Turn a tap list into the per-input streaming schedule.  Only lists whose windows are increasing runs of real
// source samples, non-empty and monotone qualify (the STANDARD edge policy away from degenerate cases);
// anything else stays on the generic kernels.
// A window may have holes: the reference drops taps whose weight is exactly zero (src/imagew-resize.c:302; a cubic
// filter at |x| = 1 for odd integer reduction factors).  The schedule gives such an input the weight 0.0, as it
// does for every input outside a slot's window, which leaves the sum bit-identical: an accumulator starts at +0.0
// and a sum of two doubles is -0.0 only if both are, so it never holds -0.0, and x + (+-0.0) = x for any other x.
// The windows of a tap list: [first[o], last[o]] for every output, widened (with zero weights) until starts and ends
// are monotone.  false: the list does not qualify (virtual samples, repeated or decreasing sources, real breaks of
// monotonicity).
bool build_windows(const iwb_axis_table &t, HostSched *hs) {
  if (t.mode != 0 || t.n_out < 1) return false;
  const int n_in = t.n_in, n_out = t.n_out;
  hs->first.resize(n_out);
  hs->last.resize(n_out);
  for (int o = 0; o < n_out; o++) {
    const int a = t.first[o], b = t.first[o + 1];
    if (b <= a) return false;
    if (t.src[a] < 0) return false;
    for (int k = a + 1; k < b; k++)
      if (t.src[k] <= t.src[k - 1]) return false;
    hs->first[o] = t.src[a];
    hs->last[o] = t.src[b - 1];
    if (hs->last[o] >= n_in) return false;
  }
  // a zero tap dropped at an end of a list can make a window start after (end before) its neighbour's: widen such
  // windows, again with zero weights, until starts and ends are monotone; real breaks of monotonicity stay generic
  for (int o = n_out - 2; o >= 0; o--) {
    if (hs->first[o] > hs->first[o + 1] + 1) return false;
    hs->first[o] = std::min(hs->first[o], hs->first[o + 1]);
  }
  for (int o = 1; o < n_out; o++) {
    if (hs->last[o] < hs->last[o - 1] - 1) return false;
    hs->last[o] = std::max(hs->last[o], hs->last[o - 1]);
  }
  return true;
}

bool build_sched(const iwb_axis_table &t, int max_slots, HostSched *hs) {
  if (!build_windows(t, hs)) return false;
  const int n_in = t.n_in, n_out = t.n_out;
  // most windows open at once
  int S = 1;
  for (int o = 0, lo = 0; o < n_out; o++) {
    while (hs->last[lo] < hs->first[o]) lo++;
    S = std::max(S, o - lo + 1);
  }
  if (S > max_slots) return false;
  static const int sizes[] = {2, 3, 4, 5, 6, 8, 12, 16};
  int SS = 0;
  for (int v : sizes)
    if (v >= S) { SS = v; break; }
  if (!SS) return false;
  hs->S = SS;
  hs->w.assign((size_t)n_in * SS, 0.0);
  hs->emit_lo.assign(n_in, 0);
  hs->emit_n.assign(n_in, 0);
  for (int o = 0; o < n_out; o++) {
    const int a = t.first[o], b = t.first[o + 1];
    for (int k = a; k < b; k++) hs->w[(size_t)t.src[k] * SS + (o % SS)] = t.w[k];
    const int r = hs->last[o];
    if (hs->emit_n[r] == 0) hs->emit_lo[r] = o;
    hs->emit_n[r]++;
    hs->max_emit = std::max(hs->max_emit, hs->emit_n[r]);
  }
  return true;
}

bool upload_sched(iwb200_job *job, const HostSched &hs, int n_in, int n_out, AxisSched *d) {
  d->S = hs.S; d->n_in = n_in; d->n_out = n_out;
  d->w = job->upload(hs.w.data(), hs.w.size());
  d->emit_lo = job->upload(hs.emit_lo.data(), hs.emit_lo.size());
  d->emit_n = job->upload(hs.emit_n.data(), hs.emit_n.size());
  d->first = job->upload(hs.first.data(), hs.first.size());
  d->last = job->upload(hs.last.data(), hs.last.size());
  return d->w && d->emit_lo && d->emit_n && d->first && d->last;
}

template <int S, int NC>
void launch_v(const VStreamParams &p, dim3 grid, size_t smem, iwb_stream st) {
  // the cp.async ring moves NC bytes per thread and row: needs NC in {4,8} and NC-byte aligned rows
  const bool ring = NC >= 4 && (p.bpr % NC == 0) && (((uintptr_t)p.src + p.x0_bytes) % NC == 0) && (p.ncols % NC == 0);
  static std::atomic<unsigned long long> attr{0};          // the attribute is per device
  once_per_device(attr, [] {
    cudaFuncSetAttribute(iwb_vstream<S, NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(iwb_vstream<S, NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  if (ring) iwb_vstream<S, NC, true><<<grid, IWB_V_THREADS, smem, st>>>(p);
  else iwb_vstream<S, NC, false><<<grid, IWB_V_THREADS, smem, st>>>(p);
}
template <int S, int NCH, int R, int EPI>
void launch_h(const HStreamParams &p, dim3 grid, size_t smem, iwb_stream st) {
  static std::atomic<unsigned long long> attr{0};
  once_per_device(attr, [] { cudaFuncSetAttribute(iwb_hstream<S, NCH, R, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
  iwb_hstream<S, NCH, R, EPI><<<grid, IWB_H_THREADS, smem, st>>>(p);
}

#ifdef IWB_QUICK   // tuning builds: only the variants the headline configuration uses
bool dispatch_v(int S, int NC, const VStreamParams &p, dim3 g, size_t sm, iwb_stream st) {
  if (S == 6 && NC == 4) { launch_v<6, 4>(p, g, sm, st); return true; }
  return false;
}
bool dispatch_h(int S, int NCH, int EPI, const HStreamParams &p, dim3 g, size_t sm, iwb_stream st) {
  if (S == 6 && NCH == 3 && EPI == 3) { launch_h<6, 3, 1, 3>(p, g, sm, st); return true; }
  return false;
}
int h_rows_per_thread(int, int) { return 1; }
#else
bool dispatch_v(int S, int NC, const VStreamParams &p, dim3 g, size_t sm, iwb_stream st) {
  if (NC == 8) {
    switch (S) {
      case 2: launch_v<2, 8>(p, g, sm, st); return true;
      case 3: launch_v<3, 8>(p, g, sm, st); return true;
      case 4: launch_v<4, 8>(p, g, sm, st); return true;
      case 5: launch_v<5, 8>(p, g, sm, st); return true;
      case 6: launch_v<6, 8>(p, g, sm, st); return true;
    }
  } else if (NC == 4) {
    switch (S) {
      case 2: launch_v<2, 4>(p, g, sm, st); return true;
      case 3: launch_v<3, 4>(p, g, sm, st); return true;
      case 4: launch_v<4, 4>(p, g, sm, st); return true;
      case 5: launch_v<5, 4>(p, g, sm, st); return true;
      case 6: launch_v<6, 4>(p, g, sm, st); return true;
      case 8: launch_v<8, 4>(p, g, sm, st); return true;
    }
  } else if (NC == 2) {
    if (S == 12) { launch_v<12, 2>(p, g, sm, st); return true; }
    if (S == 16) { launch_v<16, 2>(p, g, sm, st); return true; }
  }
  return false;
}
template <int NCH, int EPI>
bool dispatch_h_s(int S, const HStreamParams &p, dim3 g, size_t sm, iwb_stream st) {
  switch (S) {   // one row per thread: two rows per thread measured slower (fewer resident warps)
    case 2: launch_h<2, NCH, 1, EPI>(p, g, sm, st); return true;
    case 3: launch_h<3, NCH, 1, EPI>(p, g, sm, st); return true;
    case 4: launch_h<4, NCH, 1, EPI>(p, g, sm, st); return true;
    case 5: launch_h<5, NCH, 1, EPI>(p, g, sm, st); return true;
    case 6: launch_h<6, NCH, 1, EPI>(p, g, sm, st); return true;
    case 8: launch_h<8, NCH, 1, EPI>(p, g, sm, st); return true;
    case 12: launch_h<12, NCH, 1, EPI>(p, g, sm, st); return true;
    case 16: launch_h<16, NCH, 1, EPI>(p, g, sm, st); return true;
  }
  return false;
}
int h_rows_per_thread(int, int) { return 1; }  // must mirror dispatch_h_s
template <int NCH>
bool dispatch_h_e(int S, int EPI, const HStreamParams &p, dim3 g, size_t sm, iwb_stream st) {
  switch (EPI) {
    case 1: return dispatch_h_s<NCH, 1>(S, p, g, sm, st);
    case 2: return dispatch_h_s<NCH, 2>(S, p, g, sm, st);
    case 3: return dispatch_h_s<NCH, 3>(S, p, g, sm, st);
    case 4: return dispatch_h_s<NCH, 4>(S, p, g, sm, st);
  }
  return dispatch_h_s<NCH, 0>(S, p, g, sm, st);
}
bool dispatch_h(int S, int NCH, int EPI, const HStreamParams &p, dim3 g, size_t sm, iwb_stream st) {
  if (NCH == 1) return dispatch_h_e<1>(S, EPI, p, g, sm, st);
  if (NCH == 3) return dispatch_h_e<3>(S, EPI, p, g, sm, st);
  return false;
}
#endif

// Schedule in the layout of the second-generation kernels: rows [-IWB2_WPAD, n_in + IWB2_WPAD) of SS weights (SS = S
// rounded up to even, so every row starts on 16 bytes for the bulk copies); the padding rows are zero weights for
// stages that start before the first input or end after the last one.
bool upload_sched2(iwb200_job *job, const HostSched &hs, int n_in, int n_out, Sched2 *d) {
  const int S = hs.S, SS = (S + 1) & ~1;
  std::vector<double> w((size_t)(n_in + 2 * IWB2_WPAD) * SS, 0.0);
  for (int r = 0; r < n_in; r++)
    for (int k = 0; k < S; k++) w[(size_t)(r + IWB2_WPAD) * SS + k] = hs.w[(size_t)r * S + k];
  const double *base = job->upload(w.data(), w.size());
  d->w = base ? base + (size_t)IWB2_WPAD * SS : nullptr;
  d->first = job->upload(hs.first.data(), hs.first.size());
  d->last = job->upload(hs.last.data(), hs.last.size());
  d->n_in = n_in; d->n_out = n_out;
  return d->w && d->first && d->last;
}

// FP32 decision tables of one channel (threshold-table mode) for quantise8_v2; false = shape not supported.
// (nearest = true: the levels are the cells between the nearest-colour midpoints, reference: src/imagew-main.c:638-666;
// the top cell then ends inside [0, 1] and there is no dither fraction, only "is the sample safely inside its cell")
bool build_quant2(iwb200_job *job, const std::vector<double> &thr, bool nearest, Quant2 *q) {
  const int n = (int)thr.size();
  if (n < 2 || n > 256) return false;
  const int top = n - 1;
  std::vector<float> lof(n + 2), inv(n + 2);
  double worst = 1.0;
  for (int k = 0; k < n; k++) {
    const double lo = thr[k], hi = k < top ? thr[k + 1] : 2.0;
    lof[k] = (float)lo;
    if (k < top && hi > lo) worst = std::max(worst, hi / (hi - lo));
  }
  if (lof[0] != 0.0f || (!nearest && lof[top] != 1.0f)) return false;   // levels must span exactly [0, 1]
  lof[n] = 3.0f; lof[n + 1] = 3.0f;
  for (int k = 0; k < n; k++) {
    const float hif = k < top ? lof[k + 1] : 2.0f;
    if (!(hif > lof[k])) return false;
    inv[k] = (float)(1.0 / ((double)hif - (double)lof[k]));
  }
  inv[n] = inv[n + 1] = 0.0f;
  std::vector<unsigned char> blob(IWB2_TAB_BYTES, 0);
  float4 *e4 = (float4 *)blob.data();
  for (int k = 0; k < 258; k++) e4[k] = k <= n ? make_float4(lof[k], inv[k], lof[k + 1], inv[k + 1]) : make_float4(3.0f, 0.0f, 3.0f, 0.0f);
  double *thr_d = (double *)(blob.data() + IWB2_TAB_THR);
  for (int k = 0; k < 258; k++) thr_d[k] = k < n ? thr[k] : 2.0;
  // bucket i = [i/NHINT, (i+1)/NHINT): its samples must lie in level hint[i] or hint[i] + 1
  uint16_t *hint = (uint16_t *)(blob.data() + IWB2_TAB_HINT);
  int lev = 0;
  for (int i = 0; i <= IWB2_NHINT; i++) {
    const float a = (float)i / (float)IWB2_NHINT, b = (float)(i + 1) / (float)IWB2_NHINT;
    while (lev + 1 <= top && lof[lev + 1] <= a) lev++;
    if (lev + 2 <= top && lof[lev + 2] < b) return false;       // three levels in one bucket
    hint[i] = (uint16_t)lev;
  }
  for (int i = IWB2_NHINT + 1; i < IWB2_NHINT + 8; i++) hint[i] = hint[IWB2_NHINT];
  q->top = top;
  q->margin = (float)(2e-6 + 8.0 * 2.4e-7 * worst);
  if (q->margin > 0.004f) return false;
  q->blob = job->upload(blob.data(), blob.size());
  return q->blob != nullptr;
}

// nr/nb: input samples of the longest strip/segment; nout: outputs tracked (strip/segment + still-open predecessors)
size_t v_smem_bytes(int nr, int S, int nout, int nc) {
  const int nw = nc >= 4 ? nc / 4 : 1, tb = IWB_V_TB;
  return (size_t)256 * IWB_LUT_REP * 8 + (size_t)nr * S * 8 + (size_t)((nout + 3) & ~3) * 4 +
         (size_t)IWB_RING_D * IWB_V_THREADS * nw * 4 + (size_t)tb * nc * IWB_V_THREADS * 4 + 16;
}
size_t h_smem_bytes(int nb, int S, int nch, int nout, int R) {
  const int nv = nch * R;
  size_t b = (size_t)nb * S * 8;
  if (S > 8) b += (size_t)IWB_H_WARPS * S * nv * 32 * 8;       // parked samples (inlined epilogue otherwise)
  b += (size_t)nch * 256 * 8;                       // float2 level bounds
  b += 128 * 4;                                     // ordered-dither thresholds
  b += (size_t)IWB_RING_D * nv * IWB_H_THREADS * 4;
  b += (size_t)nout * 4;
  b += (size_t)(nch * 1026 + ((nch * 1026) & 1)) * 2;
  b += (size_t)IWB_H_WARPS * 32 * R * IWB_OSTAGE_PITCH;
  return b + 16;
}

}  // namespace

// Decide whether this job can use the streaming kernels and, if so, build and upload their schedules.
static void fused_plan_build(iwb200_job *job, FastPlan *f) {
  iw_context *ctx = job->ctx;
  f->enabled = 0;
  const int nch = ctx->in_nch_phys;
  // 8-bit input: both generations of streaming kernels; 16-bit input: generation 2 only (its table stays in L2)
  const bool d16 = ctx->img1.bit_depth == 16;
  if (ctx->img1.sampletype != IW_SAMPLETYPE_UINT || (ctx->img1.bit_depth != 8 && !d16)) return;
  f->depth16 = d16 ? 1 : 0;
  if ((ctx->img1.orient_transform & 7) != 0) return;
  if (nch != 1 && nch != 3) return;
  if (ctx->in_type_logical != ctx->img1.imgtype) return;
  // gray input written as RGB (the reference's default for gray images, src/imagew-main.c:1676-1684): three processed
  // channels that all read input channel 0 -- one intermediate plane, three output bytes (generation 2 only)
  const bool rep = nch == 1 && job->nproc == 3 && ctx->out_nch == 3;
  const int nco = rep ? 3 : nch;
  if (job->nproc != nco || ctx->out_nch != nco || job->has_alpha) return;
  if (ctx->img2.sampletype != IW_SAMPLETYPE_UINT || ctx->img2.bit_depth != 8) return;
  if (job->any_errdiff || ctx->intclamp) return;
  if (ctx->axis[0].per_channel_tables || ctx->axis[1].per_channel_tables) return;
  for (int p = 0; p < nco; p++) {
    const iwb_mid_channel &mc = ctx->mid_ci[job->mid_index[p]];
    if (job->mid_index[p] != p || mc.in_channel != (rep ? 0 : p) || mc.out_channel != p || mc.to_gray || mc.premultiplied) return;
    if (job->cin[p].early_bkgd || job->cin[p].lut[0] != job->cin[0].lut[0] || !job->cin[p].lut[0]) return;
  }
  HostSched hv, hh;
  if (!build_sched(job->vtab[0], 16, &hv)) return;
  if (!build_sched(job->htab[0], 16, &hh)) return;
  if (hh.S * nch > 48) return;
  {
    // 8 byte columns per thread when the rows allow 8-byte requests and the accumulators still fit
    const bool al8 = (ctx->img1.bpr % 8 == 0) && (((size_t)ctx->in_x0 * nch) % 8 == 0) && ((job->in_w * nch) % 8 == 0);
    f->NC = hv.S <= 8 ? 4 : 2;   // (8 columns per thread measured slower: fewer resident warps)
    (void)al8;
    if (const char *o = iwb_option(ctx, "cuda:nc")) { int v = atoi(o); if ((v == 4 && hv.S <= 8) || (v == 8 && hv.S <= 6)) f->NC = v; }
  }
  f->nch = nch;
  f->nco = nco;
  f->SV = hv.S;
  f->SH = hh.S;
  const int ncols = job->in_w * nch;
  // chunk of output rows per kernel pair (bounds the intermediate; it is written and re-read through L2)
  size_t chunk_mb = 512;
  if (const char *o = iwb_option(ctx, "cuda:chunk_mb")) chunk_mb = (size_t)std::max(1, atoi(o));
  size_t rows = (chunk_mb << 20) / ((size_t)ncols * 4);
  rows = std::max<size_t>(rows, 128);
  rows = (rows / 128) * 128;
  rows = std::min<size_t>(rows, (size_t)((job->out.height + 127) / 128) * 128);
  f->chunk_rows = (int)rows;
  f->pitch_j = rows;
  if (const char *o = iwb_option(ctx, "cuda:strip")) f->strip = std::max(8, atoi(o) & ~7);
  if (const char *o = iwb_option(ctx, "cuda:seg")) f->seg = std::max(4, atoi(o) & ~3);
  // shared-memory needs: longest input range of any strip / segment
  // strips restart at every chunk boundary, so bound by any window of `strip` consecutive outputs
  auto longest_any = [](const std::vector<int> &first, const std::vector<int> &last, int step) {
    int m = 0, n = (int)first.size();
    for (int o0 = 0; o0 < n; o0++) {
      int o1 = std::min(n, o0 + step);
      m = std::max(m, last[o1 - 1] - first[o0] + 1);
    }
    return m;
  };
  while (true) {
    f->v_nr_max = longest_any(hv.first, hv.last, f->strip);
    f->v_smem = v_smem_bytes(f->v_nr_max, f->SV, f->strip + f->SV, f->NC);
    if (f->v_smem <= 96 * 1024 || f->strip <= 4) break;
    f->strip /= 2;
  }
  while (true) {
    f->h_nb_max = longest_any(hh.first, hh.last, f->seg);
    f->h_smem = h_smem_bytes(f->h_nb_max, f->SH, nch, f->seg + f->SH, h_rows_per_thread(f->SH, nch));
    if (f->h_smem <= 160 * 1024 || f->seg <= 8) break;
    f->seg = (f->seg / 2) & ~3;
  }
  if (f->v_smem > 200 * 1024 || f->h_smem > 200 * 1024) return;
  if (!upload_sched(job, hv, job->in_h, job->out.height, &f->sv)) return;
  if (!upload_sched(job, hh, job->in_w, job->out.width, &f->sh)) return;
  f->lut = job->cin[0].lut[0];
  for (int p = 0; p < nco; p++) {
    const DevChanOut &o = job->cout[p];
    FastQuant &q = f->fq[p];
    q = FastQuant();
    // level bounds [lo, hi) in linear space: thresholds (dithered / nearest-by-distance) or colour midpoints
    std::vector<double> lo, hi;
    if (o.use_nc) {
      q.mode = 2;
      lo.resize(256); hi.resize(256);
      for (int c = 0; c < 256; c++) { lo[c] = c ? job->nc_host[c - 1] : 0.0; hi[c] = c < 255 ? job->nc_host[c] : 2.0; }
    } else if (o.qmode == 1 && o.color_count == 0 && o.ntab <= 256 && (o.dfamily == 0 || o.dfamily == 1 || o.dfamily == 3)) {
      q.mode = 1;
      const std::vector<double> &thr = job->thr_host[p];
      lo.assign(thr.begin(), thr.end());
      hi.resize(thr.size());
      for (size_t k = 0; k < thr.size(); k++) hi[k] = k + 1 < thr.size() ? thr[k + 1] : 2.0;
    } else {
      continue;   // linear-space arithmetic, posterisation: exact generic quantiser
    }
    const int n = (int)lo.size();
    q.top = n - 1;
    std::vector<float2> pairs(n);
    double worst = 1.0;
    for (int k = 0; k < n; k++) {
      pairs[k] = make_float2((float)lo[k], (float)hi[k]);
      if (k < n - 1 && hi[k] > lo[k]) worst = std::max(worst, hi[k] / (hi[k] - lo[k]));
    }
    // FP32 rounding of s, lo, hi (2^-24 each) and of the quotient, amplified by hi/(hi-lo); x8 safety
    q.margin = q.mode == 1 ? (float)(2e-6 + 8.0 * 2.4e-7 * worst) : (float)(8.0 * 1.2e-7);
    if (q.mode == 1 && q.margin > 0.004f) { q = FastQuant(); continue; }
    const int nh = 1024;
    std::vector<uint16_t> hint(nh + 2);
    int lev = 0;
    for (int i = 0; i <= nh; i++) {
      const double sv = (double)i / (double)nh;
      while (lev + 1 < n && lo[lev + 1] <= sv) lev++;
      // the float search only walks upward, so start one level low when the float bound rounds above i/nh
      int start = lev;
      while (start > 0 && (double)pairs[start].x > sv) start--;
      hint[i] = (uint16_t)start;
    }
    hint[nh + 1] = hint[nh];
    q.nhint = nh;
    q.pairs = job->upload(pairs.data(), pairs.size());
    q.hint = job->upload(hint.data(), hint.size());
    if (!q.pairs || !q.hint) return;
  }
  // epilogue specialisation: all channels in the same fast mode and dither, else the run-time generic one
  {
    bool same = true;
    for (int p = 1; p < nco; p++)
      if (f->fq[p].mode != f->fq[0].mode || job->cout[p].dfamily != job->cout[0].dfamily ||
          job->cout[p].dsubtype != job->cout[0].dsubtype) same = false;
    f->epi = 0;
    if (same && f->fq[0].mode == 2) f->epi = 2;
    if (same && f->fq[0].mode == 1) f->epi = job->cout[0].dfamily == 0 ? 1 : (job->cout[0].dfamily == 1 ? 3 : 4);
  }
  {
    DevChanOut four[4];
    for (int p = 0; p < 4; p++) four[p] = job->cout[p < nch ? p : 0];
    f->ch_g = job->upload(four, 4);
    if (!f->ch_g) return;
  }
  f->vfirst = hv.first;
  f->vlast = hv.last;
  f->hfirst = hh.first;
  f->hlast = hh.last;
  f->enabled = 1;

  // second generation: threshold-table epilogues with one dither family, at most 8 open windows per axis
  f->v2 = 0;
  f->ncols_pad = (ncols + 3) & ~3;
  {
    const char *o = iwb_option(ctx, "cuda:gen");
    const bool want = !(o && atoi(o) == 1);
    if (want && iwb2_supported(hv.S, hh.S, nch, nco, f->epi) && iwb2_tensor_maps_available()) {
      static_assert(sizeof(Sched2) > 0 && sizeof(Quant2) > 0, "stream2.cuh");
      job->s2_store.reset(new iwb200_job::s2_tables());
      f->s2v = &job->s2_store->v; f->s2h = &job->s2_store->h; f->q2 = job->s2_store->q;
      bool ok = upload_sched2(job, hv, job->in_h, job->out.height, f->s2v) && upload_sched2(job, hh, job->in_w, job->out.width, f->s2h);
      if (!d16) {
        std::vector<double> rep(256 * 16);                      // the vertical kernel's shared-memory table, ready to bulk-copy
        const std::vector<double> &src = job->lut_host[f->lut];
        ok = ok && src.size() == 256;
        for (int i = 0; ok && i < 256 * 16; i++) rep[i] = src[i >> 4];
        f->lut16 = job->upload(rep.data(), rep.size());
        ok = ok && f->lut16 != nullptr;
      }
      f->same_tables = 1;
      for (int p = 0; p < nco && ok; p++) {
        if (f->epi == 2) {                                       // cells between the nearest-colour midpoints (one table for all)
          std::vector<double> cells(1, 0.0);
          cells.insert(cells.end(), job->nc_host.begin(), job->nc_host.end());
          ok = build_quant2(job, cells, true, &f->q2[p]);
        } else {
          ok = build_quant2(job, job->thr_host[p], false, &f->q2[p]);
          if (p && job->thr_host[p] != job->thr_host[0]) f->same_tables = 0;
        }
      }
      if (const char *o2 = iwb_option(ctx, "cuda:strip2")) f->strip2 = std::max(8, atoi(o2) & ~7);
      if (const char *o2 = iwb_option(ctx, "cuda:seg2")) f->seg2 = std::max(4, atoi(o2) & ~3);
      if (const char *o2 = iwb_option(ctx, "cuda:uniform")) f->uniform = atoi(o2) != 0;
      if (const char *o2 = iwb_option(ctx, "cuda:overlap")) f->overlap = std::max(1, std::min(32, atoi(o2)));
      if (const char *o2 = iwb_option(ctx, "cuda:gen")) job->gen3 = atoi(o2) == 3 ? 2 : 0;
      if (const char *o2 = iwb_option(ctx, "cuda:nv")) job->gen3_nv = atoi(o2);
      if (const char *o2 = iwb_option(ctx, "cuda:seg3")) job->gen3_seg = std::max(16, atoi(o2) & ~3);
      if (rep && !f->same_tables) ok = false;                    // (one decision serves the three bytes unless the dither is random)
      f->v2 = ok ? 1 : 0;
    }
  }
  if ((d16 || rep) && !f->v2) f->enabled = 0;                    // no first-generation kernels for 16-bit input / gray as RGB
}

static bool fused_run(iwb200_job *job, FastPlan *f, const DevInput &in, const DevOutput &out, int r0, int r1,
                      iwb_stream stream) {
  if (!f->interT) {
    // (+4 floats per row: gen 2 stores column c at float c + shift; zeroed once so the 0..3 floats in front of column
    // 0 are finite -- they meet zero weights)
    const size_t bytes = (size_t)(std::max(job->in_w * f->nch, f->ncols_pad) + 4) * f->pitch_j * sizeof(float);
    f->interT = (float *)job->scratch(bytes);
    if (!f->interT) return false;
    job->order_after_stream0(stream);
    cudaMemsetAsync(f->interT, 0, bytes, stream);
  }
  job->order_after_stream0(stream);
  const int nch = f->nch;
  // second-generation kernels: TMA needs 16-byte aligned rows
  const uint8_t *src0 = in.pix + (ptrdiff_t)(in.y0 - in.row_origin) * (ptrdiff_t)in.bpr + (size_t)in.x0 * nch * (f->depth16 ? 2 : 1);
  if (f->v2 && in.bpr % 16 == 0 && ((uintptr_t)src0 & 15) == 0) {
    // The two kernels leave different units idle (V: FP64-heavy, H: latency-bound epilogue), so sub-chunks of
    // output rows are software-pipelined over two streams: H of sub-chunk i (auxiliary stream) runs beside V of
    // sub-chunk i+1 (caller's stream); the intermediate buffer is used in two halves (ping-pong).
    int nsub = f->overlap;
    const int rows_all = r1 - r0;
    if (nsub < 2 || rows_all < 64 * nsub || job->profiling) nsub = 1;
    const int half_rows = f->chunk_rows / 2;
    const size_t pitch = (size_t)f->ncols_pad + 4;
    // sub-chunk boundaries: equal parts (multiples of 8 rows), each at most one buffer half / one chunk.  Decided
    // before the streams are chosen: an overlapped run whose parts outgrow the event pool falls back to one stream
    // (and whole-chunk parts) as a whole.
    std::vector<int> cuts;
    auto make_cuts = [&](int n) {
      cuts.clear();
      const int cap = n > 1 ? half_rows : f->chunk_rows;
      const int parts = std::max(n, (rows_all + cap - 1) / cap);
      int len = ((rows_all + parts - 1) / parts + 7) & ~7;
      if (len > cap) len = std::max(8, cap & ~7);
      for (int c = r0; c < r1; c += len) cuts.push_back(c);
      cuts.push_back(r1);
    };
    make_cuts(nsub);
    if (nsub > 1 && (int)cuts.size() - 1 > 64) { nsub = 1; make_cuts(1); }
    cudaStream_t sv = (cudaStream_t)stream, sh = sv;
    if (nsub > 1) {
      if (!f->aux_stream) { cudaStream_t t; cudaStreamCreateWithFlags(&t, cudaStreamNonBlocking); f->aux_stream = t; }
      while ((int)f->ev.size() < 2 * 64 + 1) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); f->ev.push_back(e); }
      sh = (cudaStream_t)f->aux_stream;
      cudaEventRecord((cudaEvent_t)f->ev[128], sv);             // everything queued before this call
      cudaStreamWaitEvent(sh, (cudaEvent_t)f->ev[128], 0);
    }
    const int nparts = (int)cuts.size() - 1;
    for (int i = 0; i < nparts; i++) {
      const int c0 = cuts[i], c1 = cuts[i + 1];
      float *inter = f->interT + (nsub > 1 ? (size_t)(i & 1) * half_rows * pitch : 0);
      H2Job hj{};
      H2Params &hp = hj.p;
      hp.inter = inter; hp.pitch = (int)pitch; hp.ncols = job->in_w * nch; hp.row_lo = c0; hp.row_hi = c1; hp.sc = *f->s2h;
      hp.out_w = out.width; hp.out = out; hp.same_tables = f->same_tables; hp.jump = job->d_jump;
      for (int p = 0; p < 4; p++) { hp.ch[p] = job->cout[p < f->nco ? p : 0]; hp.q[p] = f->q2[p < f->nco ? p : 0]; }
      hj.S = f->SH; hj.nch = nch; hj.nco = f->nco; hj.epi = f->epi; hj.first_host = f->hfirst.data(); hj.last_host = f->hlast.data(); hj.seg_override = f->seg2; hj.allow_uniform = f->uniform;
      hp.sh = 0;
      hp.sh = iwb2_h_shift(hj);            // where the horizontal pass wants column 0 (16-byte aligned TMA tiles)
      V2Job vj{};
      V2Params &vp = vj.p;
      vp.src = src0; vp.bpr = in.bpr; vp.ncols = job->in_w * nch; vp.lut16 = f->lut16; vp.lut64k = f->lut; vp.depth16 = f->depth16; vp.sc = *f->s2v;
      vp.tm_row0 = f->vfirst[c0]; vp.tm_rows = f->vlast[c1 - 1] - f->vfirst[c0] + 1;
      vp.inter = inter; vp.pitch = (int)pitch; vp.sh = hp.sh; vp.row_base = c0; vp.row_lo = c0; vp.row_hi = c1;
      vj.S = f->SV; vj.first_host = f->vfirst.data(); vj.last_host = f->vlast.data(); vj.strip_override = f->strip2; vj.allow_uniform = f->uniform;
      job->last_path = 2; job->last_c0 = c0; job->last_c1 = c1; job->last_sh = hp.sh; job->last_pitch = pitch; job->last_inter = inter;
      if (nsub > 1 && i >= 2) cudaStreamWaitEvent(sv, (cudaEvent_t)f->ev[64 + i - 2], 0);   // this half was read by H(i-2)
      // generation 3: both passes in one persistent kernel (stream3.cuh).  Opt-in: measured slower than the two
      // kernels on cfg2 (0.33 ms against 0.20 ms)
      if (nsub == 1 && job->gen3 && iwb3_supported(f->SV, f->SH, nch, f->nco, f->epi, f->depth16)) {
        VH3State *stt = nullptr;
        for (auto &u : job->vh3) if (u->c0 == c0 && u->c1 == c1) stt = u.get();
        if (!stt) {
          if (job->vh3.size() >= 64) {                           // (callers that keep changing the row range: start over)
            cudaStreamSynchronize(sv);
            for (auto &u : job->vh3) if (u->dev) cudaFreeAsync(u->dev, sv);
            job->vh3.clear();
          }
          job->vh3.emplace_back(new VH3State()); stt = job->vh3.back().get();
        }
        bool launched = false;
        {
          prof_scope ps(job, 6, stream);
          iwb3_run(vj, hj, stt, job->gen3_nv, job->gen3_seg, sv, &launched);
        }
        if (launched) {
          g_launches++;
          job->last_gen = 3;
          if (job->negate) {
            PostParams pp;
            pp.out = out; pp.row0 = c0; pp.row1 = c1; pp.negate = 1; pp.alpha_ch = -1; pp.clear_transparent = 0;
            for (int c = 0; c < 4; c++) pp.maxcode[c] = c < job->ctx->out_nch ? job->ctx->out_ci[c].maxcode : 0;
            prof_scope ps(job, 4, (iwb_stream)sv);
            launch<PostBody>(pp, (size_t)(c1 - c0) * out.width, (iwb_stream)sv);
          }
          continue;
        }
        if (job->profiling && !job->pending.empty()) {           // nothing ran inside that bracket
          cudaEventDestroy((cudaEvent_t)job->pending.back().a); cudaEventDestroy((cudaEvent_t)job->pending.back().b);
          job->pending.pop_back();
        }
      }
      job->last_gen = 2;
      {
        prof_scope ps(job, 5, stream);
        if (!iwb2_run_v(vj, sv)) return fail(job, "fast path: no vertical kernel variant (gen 2)");
        g_launches++;
      }
      if (nsub > 1) {
        cudaEventRecord((cudaEvent_t)f->ev[i], sv);
        cudaStreamWaitEvent(sh, (cudaEvent_t)f->ev[i], 0);
      }
      {
        prof_scope ps(job, 1, (iwb_stream)sh);
        if (!iwb2_run_h(hj, sh)) return fail(job, "fast path: no horizontal kernel variant (gen 2)");
        g_launches++;
      }
      if (job->negate) {
        PostParams pp;
        pp.out = out; pp.row0 = c0; pp.row1 = c1; pp.negate = 1; pp.alpha_ch = -1; pp.clear_transparent = 0;
        for (int c = 0; c < 4; c++) pp.maxcode[c] = c < job->ctx->out_nch ? job->ctx->out_ci[c].maxcode : 0;
        prof_scope ps(job, 4, (iwb_stream)sh);
        launch<PostBody>(pp, (size_t)(c1 - c0) * out.width, (iwb_stream)sh);
      }
      if (nsub > 1) cudaEventRecord((cudaEvent_t)f->ev[64 + i], sh);
    }
    if (nsub > 1) {                                             // the caller's stream continues after the last H passes
      cudaStreamWaitEvent(sv, (cudaEvent_t)f->ev[64 + nparts - 1], 0);
      if (nparts >= 2) cudaStreamWaitEvent(sv, (cudaEvent_t)f->ev[64 + nparts - 2], 0);
    }
    if (cudaGetLastError() != cudaSuccess) return fail(job, "fast path: kernel launch failed (gen 2)");
    return true;
  }
  for (int c0 = r0; c0 < r1; c0 += f->chunk_rows) {
    const int c1 = std::min(r1, c0 + f->chunk_rows);
    job->last_path = 1; job->last_c0 = c0; job->last_c1 = c1; job->last_pitch = f->pitch_j; job->last_inter = f->interT;
    job->last_gen = 1;
    VStreamParams vp;
    vp.src = in.pix; vp.bpr = in.bpr; vp.row_origin = in.row_origin; vp.y0 = in.y0;
    vp.x0_bytes = (size_t)in.x0 * nch; vp.ncols = job->in_w * nch; vp.lut = f->lut; vp.sc = f->sv;
    vp.interT = f->interT; vp.pitch_j = f->pitch_j; vp.row_lo = c0; vp.row_hi = c1; vp.strip = f->strip;
    vp.intclamp = job->cm.intclamp;
    dim3 gv((unsigned)((vp.ncols + IWB_V_THREADS * f->NC - 1) / (IWB_V_THREADS * f->NC)), (unsigned)((c1 - c0 + f->strip - 1) / f->strip));
    {
      prof_scope ps(job, 5, stream);
      if (!dispatch_v(f->SV, f->NC, vp, gv, f->v_smem, stream)) return fail(job, "fast path: no vertical kernel variant");
      g_launches++;
    }
    HStreamParams hp;
    hp.interT = f->interT; hp.pitch_j = f->pitch_j; hp.row_lo = c0; hp.row_hi = c1; hp.sc = f->sh; hp.seg = f->seg;
    hp.ch_g = f->ch_g;
    hp.out = out; hp.intclamp = job->cm.intclamp; hp.out_w = out.width; hp.jump = job->d_jump;
    for (int p = 0; p < 4; p++) { hp.ch[p] = job->cout[p < nch ? p : 0]; hp.fq[p] = f->fq[p < nch ? p : 0]; }
    const int rows_per_cta = 32 * IWB_H_WARPS * h_rows_per_thread(f->SH, nch);
    dim3 gh((unsigned)((out.width + f->seg - 1) / f->seg), (unsigned)((c1 - c0 + rows_per_cta - 1) / rows_per_cta));
    {
      prof_scope ps(job, 1, stream);
      if (!dispatch_h(f->SH, nch, f->epi, hp, gh, f->h_smem, stream)) return fail(job, "fast path: no horizontal kernel variant");
      g_launches++;
    }
    if (job->negate) {
      PostParams pp;
      pp.out = out; pp.row0 = c0; pp.row1 = c1; pp.negate = 1; pp.alpha_ch = -1; pp.clear_transparent = 0;
      for (int c = 0; c < 4; c++) pp.maxcode[c] = c < job->ctx->out_nch ? job->ctx->out_ci[c].maxcode : 0;
      prof_scope ps(job, 4, stream);
      launch<PostBody>(pp, (size_t)(c1 - c0) * out.width, stream);
    }
  }
  if (cudaGetLastError() != cudaSuccess) return fail(job, "fast path: kernel launch failed");
  return true;
}
// ---- gather-form streaming kernels (gather4.cuh): jobs with a real alpha channel -------------------------------------
// Per-output schedule: the window [first[o], first[o] + W) and its W weights (0.0 where the reference's list has no tap).
static bool build_gather_sched(iwb200_job *job, const HostSched &hs, const iwb_axis_table &t, int W, GSched *d) {
  const int n_out = t.n_out;
  std::vector<double> w((size_t)n_out * W, 0.0);
  for (int o = 0; o < n_out; o++)
    for (int k = t.first[o]; k < t.first[o + 1]; k++) {
      const int rel = t.src[k] - hs.first[o];
      if (rel < 0 || rel >= W) return false;
      w[(size_t)o * W + rel] = t.w[k];
    }
  d->first = job->upload(hs.first.data(), hs.first.size());
  d->w = job->upload(w.data(), w.size());
  d->n_in = t.n_in; d->n_out = n_out;
  return d->first && d->w;
}

static void gather_plan_build(iwb200_job *job) {
  iw_context *ctx = job->ctx;
  iwb200_job::gather_plan &g = job->gather;
  g.enabled = 0;
  const int nch = ctx->in_nch_phys;
  if (ctx->img1.sampletype != IW_SAMPLETYPE_UINT || (ctx->img1.bit_depth != 8 && ctx->img1.bit_depth != 16)) return;
  if ((ctx->img1.orient_transform & 7) != 0) return;
  // images WITHOUT alpha that the scatter-form kernels do not take (16-bit output, more than 16 open windows) run on the
  // tile kernels with no premultiplication: gray (1 channel) and RGB (3)
  const bool noalpha = !job->has_alpha;
  if (noalpha ? (nch != 1 && nch != 3) : (nch != 2 && nch != 4)) return;
  if (ctx->in_type_logical != ctx->img1.imgtype || job->nproc != nch || ctx->mid_nch != nch) return;
  if (!noalpha && ctx->in_alpha_index != nch - 1) return;
  if (ctx->img2.sampletype != IW_SAMPLETYPE_UINT || (ctx->img2.bit_depth != 8 && ctx->img2.bit_depth != 16)) return;
  if (job->any_errdiff) return;
  if (ctx->axis[0].per_channel_tables || ctx->axis[1].per_channel_tables) return;
  // processed channel p <-> interleaved (physical) channel: a bijection; colour premultiplied, alpha plain
  bool seen[4] = {false, false, false, false};
  const double *lut_c = nullptr, *lut_a = nullptr;
  for (int p = 0; p < job->nproc; p++) {
    const iwb_mid_channel &mc = ctx->mid_ci[job->mid_index[p]];
    const int c = mc.in_channel;
    if (c < 0 || c >= nch || seen[c] || mc.to_gray) return;
    seen[c] = true;
    const bool is_alpha = mc.channeltype == IW_CHANNELTYPE_ALPHA;
    if (noalpha ? is_alpha : (is_alpha != (c == nch - 1))) return;
    const DevChanIn &ci = job->cin[p];
    if (ci.early_bkgd || !ci.lut[0]) return;
    if (noalpha) {
      if (ci.premult || job->cout[p].premult) return;
      if (lut_c && lut_c != ci.lut[0]) return;
      lut_c = ci.lut[0];
    } else if (is_alpha) {
      if (ci.premult) return;
      if (lut_a && lut_a != ci.lut[0]) return;
      lut_a = ci.lut[0];
    } else {
      if (!ci.premult || !ci.alpha_lut || ci.alpha_ch != nch - 1) return;
      if ((lut_c && lut_c != ci.lut[0]) || (lut_a && lut_a != ci.alpha_lut)) return;
      lut_c = ci.lut[0];
      lut_a = ci.alpha_lut;
    }
    g.chan_of_proc[p] = c;
    g.ch[c] = job->cout[p];
  }
  if (!lut_c || (!noalpha && !lut_a)) return;
  HostSched hv, hh;
  if (!build_windows(job->vtab[0], &hv) || !build_windows(job->htab[0], &hh)) return;
  auto width_of = [](const HostSched &h) {
    int m = 1;
    for (size_t o = 0; o < h.first.size(); o++) m = std::max(m, h.last[o] - h.first[o] + 1);
    return (m + 1) & ~1;
  };
  g.WV = width_of(hv); g.WH = width_of(hh);
  g.tile = 0;
  if (noalpha || !iwg_supported(g.WV, g.WH, nch)) {
    // longer windows (reductions beyond 2x, Lanczos3): the tile kernels, which have no random dither
    if (!iwt_supported(g.WV, g.WH, nch) || job->any_random) return;
    g.tile = 1;
    g.tile_max_nb = 0;
    const int ow = (int)hh.first.size();
    for (int s0 = 0; s0 < ow; s0 += 32) g.tile_max_nb = std::max(g.tile_max_nb, hh.first[std::min(s0 + 32, ow) - 1] + g.WH - hh.first[s0]);
    g.cm_g = job->upload(&job->cm, 1);
    if (!g.cm_g) return;
  }
  g.v2 = 0;
  if (noalpha && !ctx->intclamp) {
    HostSched hv2;
    const char *o = iwb_option(ctx, "cuda:gen");
    if (!(o && atoi(o) == 1) && build_sched(job->vtab[0], 16, &hv2) && iwb2_v_supported(hv2.S)) {
      job->s2_store.reset(new iwb200_job::s2_tables());
      bool ok = upload_sched2(job, hv2, job->in_h, job->out.height, &job->s2_store->v);
      if (ok && ctx->img1.bit_depth == 8) {
        std::vector<double> rep(256 * 16);                      // the vertical kernel's shared-memory table, ready to bulk-copy
        const std::vector<double> &src = job->lut_host[lut_c];
        ok = src.size() == 256;
        for (int i = 0; ok && i < 256 * 16; i++) rep[i] = src[i >> 4];
        g.lut16 = ok ? job->upload(rep.data(), rep.size()) : nullptr;
        ok = ok && g.lut16 != nullptr;
      }
      if (ok) { g.v2 = 1; g.v2_S = hv2.S; g.v2first = hv2.first; g.v2last = hv2.last; }
    }
  }
  g.nco = ctx->out_nch;
  for (int k = 0; k < g.nco; k++) {
    int src = -1;
    for (int c = 0; c < nch; c++) if (g.ch[c].out_ch == k) src = c;
    if (src != k) return;                                        // (output channel k is interleaved channel k)
    g.out_src[k] = src;
  }
  if (g.nco != nch && (noalpha || g.nco != nch - 1)) return;
  if (!build_gather_sched(job, hv, job->vtab[0], g.WV, &g.sv) || !build_gather_sched(job, hh, job->htab[0], g.WH, &g.sh)) return;
  {
    DevChanOut four[4];
    for (int c = 0; c < 4; c++) four[c] = g.ch[c < nch ? c : 0];
    g.ch_g = job->upload(four, 4);
    if (!g.ch_g) return;
  }
  g.est = nullptr;
  {
    // 16-bit output without dither: a validated model of the table's midpoints (estimate.h) settles almost every sample
    const DevChanOut &o = g.ch[0];
    int p0 = -1;
    for (int p = 0; p < job->nproc; p++) if (g.chan_of_proc[p] == 0) p0 = p;
    if (p0 >= 0 && ctx->img2.bit_depth == 16 && o.qmode == 1 && !o.codeval && !o.use_nc && o.dfamily == 0 && o.guess_cs >= 0 &&
        (int)job->thr_host[p0].size() == o.ntab) {
      IwbEstTable t;
      if (iwb_build_est_table(o.guess_cs, job->thr_host[p0], &t)) {
        g.est = job->upload(t.rec.data(), t.rec.size());
        g.est_margin = t.margin;
      }
    }
  }
  g.vfirst = hv.first; g.vlast = hv.last;
  g.lut_c = lut_c; g.lut_a = lut_a;
  g.nch = nch;
  g.depth16 = ctx->img1.bit_depth == 16;
  const int ncols = job->in_w * nch;
  g.pitch = (size_t)((ncols + 3) & ~3);
  size_t rows = ((size_t)512 << 20) / (g.pitch * 4);
  rows = std::max<size_t>(32, (rows / 32) * 32);
  rows = std::min<size_t>(rows, (size_t)((job->out.height + 31) / 32) * 32);
  g.chunk_rows = (int)rows;
  g.enabled = 1;
}

static bool gather_run(iwb200_job *job, const DevInput &in, const DevOutput &out, int r0, int r1, iwb_stream stream) {
  iwb200_job::gather_plan &g = job->gather;
  if (!g.inter) {
    g.inter = (float *)job->scratch(g.pitch * g.chunk_rows * sizeof(float));
    if (!g.inter) return false;
  }
  job->order_after_stream0(stream);
  const int nch = g.nch;
  const uint8_t *src0 = in.pix + (ptrdiff_t)(in.y0 - in.row_origin) * (ptrdiff_t)in.bpr + (size_t)in.x0 * nch * (g.depth16 ? 2 : 1);
  for (int c0 = r0; c0 < r1; c0 += g.chunk_rows) {
    const int c1 = std::min(r1, c0 + g.chunk_rows);
    job->last_path = 3; job->last_c0 = c0; job->last_c1 = c1; job->last_sh = 0; job->last_pitch = g.pitch; job->last_inter = g.inter;
    job->last_gen = 4;
    GVJob vj{};
    GVParams &vp = vj.p;
    vp.src = src0; vp.bpr = in.bpr; vp.ncols = job->in_w * nch; vp.lut_c = g.lut_c; vp.lut_a = g.lut_a; vp.depth16 = g.depth16;
    // the tensor covers exactly the rows the caller must provide for these output rows (iwb200_job_band_input_rows);
    // window rows outside it only meet zero weights and arrive as zeros
    vp.tm_row0 = job->vtab[0].min_src(c0, c1);
    vp.tm_rows = job->vtab[0].max_src(c0, c1) - vp.tm_row0 + 1;
    vp.sc = g.sv; vp.inter = g.inter; vp.pitch = (int)g.pitch; vp.row_base = c0; vp.row_lo = c0; vp.row_hi = c1;
    vp.intclamp = job->cm.intclamp;
    vj.W = g.WV; vj.nchp = job->has_alpha ? nch : 0;            // (0: no alpha, nothing to premultiply)
    if (g.tile && g.v2) {
      V2Job v2j{};
      V2Params &v2p = v2j.p;
      v2p.src = src0; v2p.bpr = in.bpr; v2p.ncols = job->in_w * nch; v2p.lut16 = g.lut16; v2p.lut64k = g.lut_c; v2p.depth16 = g.depth16;
      v2p.sc = job->s2_store->v;
      v2p.tm_row0 = g.v2first[c0]; v2p.tm_rows = g.v2last[c1 - 1] - g.v2first[c0] + 1;
      v2p.inter = g.inter; v2p.pitch = (int)g.pitch; v2p.sh = 0; v2p.row_base = c0; v2p.row_lo = c0; v2p.row_hi = c1;
      v2j.S = g.v2_S; v2j.first_host = g.v2first.data(); v2j.last_host = g.v2last.data(); v2j.strip_override = 0; v2j.allow_uniform = 1;
      prof_scope ps(job, 5, stream);
      job->last_gen = 5;
      if (!iwb2_run_v(v2j, (cudaStream_t)stream)) return fail(job, "tile path: no vertical kernel variant");
      g_launches++;
    } else {
      prof_scope ps(job, 5, stream);
      if (g.tile) job->last_gen = 5;
      if (!(g.tile ? iwt_run_v(vj, g.vfirst.data(), (cudaStream_t)stream) : iwg_run_v(vj, (cudaStream_t)stream)))
        return fail(job, "gather path: no vertical kernel variant");
      g_launches++;
    }
    GHJob hj{};
    GHParams &hp = hj.p;
    hp.inter = g.inter; hp.pitch = (int)g.pitch; hp.ncols = job->in_w * nch; hp.row_lo = c0; hp.row_hi = c1; hp.sc = g.sh;
    hp.out_w = out.width; hp.out = out; hp.cm = job->cm;
    for (int c = 0; c < 4; c++) hp.ch[c] = g.ch[c < nch ? c : 0];
    hp.ch_g = g.ch_g;
    hp.nco = g.nco;
    hp.jump = job->d_jump; hp.any_random = job->any_random;
    hp.est = g.est; hp.est_margin = g.est_margin;
    hj.W = g.WH; hj.nch = nch; hj.nco = g.nco;
    {
      prof_scope ps(job, 1, stream);
      if (!(g.tile ? iwt_run_h(hj, g.cm_g, g.tile_max_nb, (cudaStream_t)stream) : iwg_run_h(hj, (cudaStream_t)stream)))
        return fail(job, "gather path: no horizontal kernel variant");
      g_launches++;
    }
    if (job->negate || job->clear_transparent) {
      PostParams pp;
      pp.out = out; pp.row0 = c0; pp.row1 = c1; pp.negate = job->negate; pp.alpha_ch = job->out_alpha_ch;
      pp.clear_transparent = job->clear_transparent;
      for (int c = 0; c < 4; c++) pp.maxcode[c] = c < job->ctx->out_nch ? job->ctx->out_ci[c].maxcode : 0;
      prof_scope ps(job, 4, stream);
      launch<PostBody>(pp, (size_t)(c1 - c0) * out.width, stream);
    }
  }
  if (cudaGetLastError() != cudaSuccess) return fail(job, "gather path: kernel launch failed");
  return true;
}
#endif  // !IWB_EMU

static void job_free(iwb200_job *job) {
  if (!job) return;
  dev_guard on_device(job->device);
#if !defined(IWB_EMU)
  // memory goes back to the stream-ordered pool on the default stream: nothing of this job may still be running on a
  // caller's stream
  if (job->used_external_stream) { cudaDeviceSynchronize(); job->used_external_stream = 0; }
  if (job->s0_event) cudaEventDestroy((cudaEvent_t)job->s0_event);
  for (int k = 0; k < 3; k++) if (job->streams[k]) cudaStreamDestroy((cudaStream_t)job->streams[k]);
  for (void *e : job->events) cudaEventDestroy((cudaEvent_t)e);
  for (void *e : job->fused.ev) cudaEventDestroy((cudaEvent_t)e);
  if (job->fused.aux_stream) cudaStreamDestroy((cudaStream_t)job->fused.aux_stream);
#endif
#if !defined(IWB_EMU)
  for (auto &u : job->vh3) if (u->dev) cudaFreeAsync(u->dev, 0);
#endif
  for (void *p : job->allocs) dev_free(p);
  dev_free(job->d_in);
  dev_free(job->d_out);
  dev_free(job->d_opt);
  delete job;
}

// ---- plan cache.  A built job (device tables, schedules, scratch, staging buffers, streams) depends only on the
// settings serialised by plan_key(); finished jobs are parked here and handed to the next context with the
// same key, so a batch of same-shaped images builds its plan once.  Disable with IWB200_PLAN_CACHE=0.
static std::mutex g_pool_mu;
static std::vector<iwb200_job *> g_pool;   // idle jobs, most recently used last
static const size_t kPoolMax = 8;

template <class T>
static void key_put(std::string &k, const T &v) { k.append(reinterpret_cast<const char *>(&v), sizeof(T)); }

static std::string plan_key(const iw_context *c, int device) {
  std::string k;
  key_put(k, device);
  key_put(k, c->img1.imgtype); key_put(k, c->img1.bit_depth); key_put(k, c->img1.sampletype);
  key_put(k, c->img1.width); key_put(k, c->img1.height); key_put(k, c->img1.bpr); key_put(k, c->img1.orient_transform);
  key_put(k, c->img1cs); key_put(k, c->img2cs);
  key_put(k, c->in_x0); key_put(k, c->in_y0); key_put(k, c->in_w); key_put(k, c->in_h);
  key_put(k, c->in_type_logical); key_put(k, c->in_nch_phys); key_put(k, c->in_nch_logical); key_put(k, c->in_alpha_index);
  for (int i = 0; i < 4; i++) { key_put(k, c->in_ci[i].channeltype); key_put(k, c->in_ci[i].slow_path); key_put(k, c->in_ci[i].maxcode); }
  key_put(k, c->mid_type); key_put(k, c->mid_nch); key_put(k, c->mid_alpha_index);
  for (int i = 0; i < 4; i++) {
    const iwb_mid_channel &m = c->mid_ci[i];
    key_put(k, m.channeltype); key_put(k, m.to_gray); key_put(k, m.in_channel); key_put(k, m.out_channel);
    key_put(k, m.early_bkgd); key_put(k, m.premultiplied);
  }
  key_put(k, c->img2.imgtype); key_put(k, c->img2.bit_depth); key_put(k, c->img2.sampletype);
  key_put(k, c->img2.width); key_put(k, c->img2.height); key_put(k, c->out_nch);
  for (int i = 0; i < 4; i++) {
    const iwb_out_channel &o = c->out_ci[i];
    key_put(k, o.channeltype); key_put(k, o.dither_family); key_put(k, o.dither_subtype); key_put(k, o.color_count);
    key_put(k, o.maxcode); key_put(k, o.bkgd1); key_put(k, o.bkgd2);
  }
  for (int d = 0; d < 2; d++) {
    const iwb_axis_request &a = c->axis[d];
    key_put(k, a.family); key_put(k, a.edge_policy); key_put(k, a.use_offset); key_put(k, a.per_channel_tables);
    key_put(k, a.param1); key_put(k, a.param2); key_put(k, a.blur); key_put(k, a.out_true_size); key_put(k, a.translate);
    for (int j = 0; j < 3; j++) key_put(k, a.channel_offset[j]);
  }
  key_put(k, c->grayscale_formula); for (int j = 0; j < 3; j++) key_put(k, c->gray_w[j]);
  key_put(k, c->no_gamma); key_put(k, c->intclamp); key_put(k, c->random_seed); key_put(k, c->negate_target);
  key_put(k, c->apply_bkgd); key_put(k, c->bkgd_strategy); key_put(k, c->bkgd_checkerboard); key_put(k, c->bkgd_source);
  key_put(k, c->bkgd1alpha); key_put(k, c->bkgd2alpha); key_put(k, c->check_size);
  key_put(k, c->check_origin[0]); key_put(k, c->check_origin[1]);
  key_put(k, c->disable_out_tables); key_put(k, c->reduced_out_maxcode); key_put(k, c->output_profile);
  key_put(k, c->in_label_inputcs.c[3]); key_put(k, c->req_bkgd.c[3]); key_put(k, c->req_bkgd2.c[3]);
  for (const struct iwb_option &o : c->options)
    if (o.name.compare(0, 5, "cuda:") == 0 && o.name != "cuda:band") { k += o.name; k += '='; k += o.val; k += ';'; }
  return k;
}

static bool pool_enabled() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("IWB200_PLAN_CACHE"); on = !(e && e[0] == '0'); }
  return on != 0;
}


extern "C" {

int iwb200_debug_build_schedule(const int *tap_first, const int *tap_src, const double *tap_w, int n_in, int n_out,
                                int *slots, int *first, int *last, double *w) {
#if defined(IWB_EMU)
  (void)tap_first; (void)tap_src; (void)tap_w; (void)n_in; (void)n_out; (void)slots; (void)first; (void)last; (void)w;
  return 0;
#else
  if (!tap_first || !tap_src || !tap_w || n_out < 1 || n_in < 1) return 0;
  iwb_axis_table t;
  t.mode = 0; t.n_in = n_in; t.n_out = n_out;
  t.first.assign(tap_first, tap_first + n_out + 1);
  t.src.assign(tap_src, tap_src + tap_first[n_out]);
  t.w.assign(tap_w, tap_w + tap_first[n_out]);
  HostSched hs;
  if (!build_sched(t, 16, &hs)) return 0;
  if (slots) *slots = hs.S;
  for (int o = 0; o < n_out; o++) { if (first) first[o] = hs.first[o]; if (last) last[o] = hs.last[o]; }
  if (w) std::copy(hs.w.begin(), hs.w.end(), w);
  return 1;
#endif
}

const char *iwb200_version(void) {
#if defined(IWB_EMU)
  return "imagew_b200 0.1 host-emulation (test build, not the product)";
#else
  return "imagew_b200 0.1 sm_100a";
#endif
}
unsigned long long iwb200_kernel_launch_count(void) { return g_launches.load(); }
const char *iwb200_job_last_error(const iwb200_job *job) { return job ? job->err.c_str() : "null job"; }

void iwb200_job_destroy(iwb200_job *job) {
  if (!job) return;
  if (pool_enabled() && job->err.empty() && !job->key.empty()) {
#if !defined(IWB_EMU)
    if (job->used_external_stream) { dev_guard on_device(job->device); cudaDeviceSynchronize(); job->used_external_stream = 0; }
#endif
    job->ctx = nullptr;
    iwb200_job *evict = nullptr;
    {
      std::lock_guard<std::mutex> lk(g_pool_mu);
      g_pool.push_back(job);
      if (g_pool.size() > kPoolMax) { evict = g_pool.front(); g_pool.erase(g_pool.begin()); }
    }
    job_free(evict);
    return;
  }
  job_free(job);
}

// The context must have been through iwb_prepare() (iw_process_image does that; standalone callers use
// iwb200_job_create which prepares a fresh context itself).
static iwb200_job *job_build(iw_context *ctx, int device) {
  iwb200_job *job = new iwb200_job();
  job->ctx = ctx;
  job->device = device;
  if (!dev_select(device, job->err)) return job;

  const int in_w = ctx->in_w, in_h = ctx->in_h;
  const int out_w = ctx->img2.width, out_h = ctx->img2.height;
  job->in_w = in_w;
  job->in_h = in_h;
  const bool transposed = (ctx->img1.orient_transform & 4) != 0;
  job->phys_w = transposed ? ctx->img1.height : ctx->img1.width;
  job->phys_h = transposed ? ctx->img1.width : ctx->img1.height;

  const bool is_float_in = ctx->img1.sampletype == IW_SAMPLETYPE_FLOATINGPOINT;
  if (is_float_in && ctx->img1.bit_depth != 32) { fail(job, "CUDA backend: floating-point input must be 32-bit"); return job; }
  if (!is_float_in && ctx->img1.bit_depth != 1 && ctx->img1.bit_depth != 2 && ctx->img1.bit_depth != 4 &&
      ctx->img1.bit_depth != 8 && ctx->img1.bit_depth != 16) {
    fail(job, "CUDA backend: unsupported input bit depth");
    return job;
  }
  // Floating-point samples in a non-linear colourspace: the reference evaluates the transfer curve per sample with
  // the host libm (src/imagew-main.c:53-66, :348-372).  A continuous pow() cannot be tabulated and the device's pow is
  // not bit-identical, so iwb200_job_run_host converts such input to linear FP64 samples on the host cores, with the
  // same libm, before the upload (float_prepass); the kernels then read doubles (DevInput::depth 64).
  job->float_prepass = is_float_in && !ctx->no_gamma && ctx->img1cs.cstype != IW_CSTYPE_LINEAR;

  DevInput &in = job->in;
  in.pix = nullptr; in.bpr = ctx->img1.bpr; in.row_origin = 0;
  in.width = ctx->img1.width; in.height = ctx->img1.height;
  in.orient = (int)(ctx->img1.orient_transform & 7);
  in.x0 = ctx->in_x0; in.y0 = ctx->in_y0;
  in.depth = job->float_prepass ? 64 : ctx->img1.bit_depth; in.nch_phys = ctx->in_nch_phys; in.is_float = is_float_in;

  DevOutput &out = job->out;
  out.pix = nullptr; out.row_origin = 0; out.width = out_w; out.height = out_h;
  out.nch = ctx->out_nch; out.depth = ctx->img2.bit_depth;
  out.is_float = ctx->img2.sampletype == IW_SAMPLETYPE_FLOATINGPOINT;
  out.bpr = iw_calc_bytesperrow(out_w, ctx->img2.bit_depth * ctx->out_nch);

  DevCommon &cm = job->cm;
  cm.checker = ctx->bkgd_checkerboard; cm.check_size = ctx->check_size;
  cm.check_ox = ctx->check_origin[IW_DIMENSION_H]; cm.check_oy = ctx->check_origin[IW_DIMENSION_V];
  cm.bkgd1alpha = ctx->bkgd1alpha; cm.bkgd2alpha = ctx->bkgd2alpha; cm.intclamp = ctx->intclamp;

  // processing order: alpha first (reference: src/imagew-main.c:1410-1435)
  job->has_alpha = IW_IMGTYPE_HAS_ALPHA(ctx->mid_type);
  job->nproc = 0;
  if (job->has_alpha) job->mid_index[job->nproc++] = ctx->mid_alpha_index;
  for (int m = 0; m < ctx->mid_nch; m++)
    if (ctx->mid_ci[m].channeltype != IW_CHANNELTYPE_ALPHA) job->mid_index[job->nproc++] = m;

  iw_csdescr lin_cs;
  iw_make_linear_csdescr(&lin_cs);
  const bool translucent = iwb_bkgd_is_translucent(ctx);
  const bool nc_table_exists = !ctx->disable_out_tables && !ctx->no_gamma && ctx->img2cs.cstype != IW_CSTYPE_LINEAR &&
                               !out.is_float && ctx->img2.bit_depth <= 8 && ((size_t)out_w) * out_h > 512;
  const double *nc_dev = nullptr;
  if (nc_table_exists) {  // midpoints between consecutive codes (reference: main.c:1197-1236)
    std::vector<double> t(255);
    double prev = 0.0;
    for (int i = 0; i < 255; i++) {
      double curr = iwb_to_linear(((double)(i + 1)) / 255, &ctx->img2cs);
      t[i] = (prev + curr) / 2.0;
      prev = curr;
    }
    nc_dev = job->upload(t.data(), t.size());
    if (!nc_dev) return job;
    job->nc_host = t;
  }

  std::vector<lut_key> luts;
  double taps_v = 0, taps_h = 0;
  for (int p = 0; p < job->nproc; p++) {
    const int m = job->mid_index[p];
    const iwb_mid_channel &mc = ctx->mid_ci[m];
    const bool is_alpha = mc.channeltype == IW_CHANNELTYPE_ALPHA;
    const iw_csdescr &cs_in = (is_alpha || ctx->no_gamma) ? lin_cs : ctx->img1cs;

    // ---- input side
    DevChanIn &ci = job->cin[p];
    memset(&ci, 0, sizeof ci);
    ci.in_ch = mc.in_channel; ci.to_gray = mc.to_gray; ci.gsf = ctx->grayscale_formula;
    for (int k = 0; k < 3; k++) ci.gw[k] = ctx->gray_w[k];
    const int nsrc = mc.to_gray ? 3 : 1;
    for (int k = 0; k < nsrc; k++) {
      int ch = mc.in_channel + k;
      if (ch < ctx->in_nch_phys && !is_float_in) {
        ci.lut[k] = input_lut(job, luts, ctx->img1.bit_depth, ctx->in_ci[ch].maxcode, cs_in);
        if (!ci.lut[k]) return job;
      }
    }
    ci.premult = mc.premultiplied;
    ci.early_bkgd = !mc.premultiplied && ctx->apply_bkgd && ctx->bkgd_strategy == IWB_BKGD_EARLY;
    ci.bkgd = mc.early_bkgd;
    ci.alpha_ch = ctx->in_alpha_index;
    if ((ci.premult || ci.early_bkgd) && ci.alpha_ch < ctx->in_nch_phys && !is_float_in) {
      ci.alpha_lut = input_lut(job, luts, ctx->img1.bit_depth, ctx->in_ci[ci.alpha_ch].maxcode, lin_cs);
      if (!ci.alpha_lut) return job;
    }

    // ---- resampling tables (shared between channels unless the plan says otherwise)
    for (int d = 1; d >= 0; d--) {
      std::vector<iwb_axis_table> &tabs = d == IW_DIMENSION_V ? job->vtab : job->htab;
      DevAxis *dev = d == IW_DIMENSION_V ? job->vax : job->hax;
      if (p > 0 && !ctx->axis[d].per_channel_tables) {
        dev[p] = dev[0];
        if (d == IW_DIMENSION_V) job->vtab_of[p] = 0;
      } else {
        tabs.emplace_back();
        int n_in = d == IW_DIMENSION_V ? in_h : in_w;
        int n_out = d == IW_DIMENSION_V ? out_h : out_w;
        if (!iwb_build_axis_table(ctx, d, mc.channeltype, n_in, n_out, &tabs.back())) { fail(job, "weight table construction failed"); return job; }
        if (!upload_axis(job, tabs.back(), &dev[p])) return job;
        if (d == IW_DIMENSION_V) job->vtab_of[p] = (int)tabs.size() - 1;
      }
    }
    {
      const iwb_axis_table &vt = job->vtab[job->vtab_of[p]];
      const iwb_axis_table &ht = ctx->axis[IW_DIMENSION_H].per_channel_tables ? job->htab.back() : job->htab[0];
      taps_v += (vt.mode == 0 ? (double)vt.w.size() : (double)out_h) * in_w;
      taps_h += (ht.mode == 0 ? (double)ht.w.size() : (double)out_w) * out_h;
    }

    // ---- output side
    DevChanOut &o = job->cout[p];
    memset(&o, 0, sizeof o);
    job->building_channel = p;
    o.out_ch = mc.out_channel;
    o.is_alpha = is_alpha;
    o.premult = mc.premultiplied;
    o.late_bkgd = ctx->apply_bkgd && ctx->bkgd_strategy == IWB_BKGD_LATE;
    o.bkgd_translucent = translucent;
    o.maxcode = 255.0;
    if (o.out_ch >= 0) {
      const iwb_out_channel &oc = ctx->out_ci[o.out_ch];
      const iw_csdescr &cs_out = (is_alpha || ctx->no_gamma) ? lin_cs : ctx->img2cs;
      o.bkgd1 = oc.bkgd1; o.bkgd2 = oc.bkgd2;
      o.maxcode = (double)oc.maxcode;
      o.color_count = oc.color_count;
      o.dfamily = oc.dither_family; o.dsubtype = oc.dither_subtype;
      o.qmode = cs_out.cstype == IW_CSTYPE_LINEAR ? 0 : 1;
      if (!out.is_float) {
        o.use_nc = nc_table_exists && !is_alpha && o.dfamily == IW_DITHERFAMILY_NONE && o.color_count == 0;
        o.nc_tbl = nc_dev;
        if (o.qmode == 1 && !build_output_tables(job, ctx, oc, cs_out, &o)) return job;
        o.errdiff = o.dfamily == IW_DITHERFAMILY_ERRDIFF;
        if (o.errdiff) {
          job->any_errdiff = 1;
          memcpy(job->errdiff_m[p], kErrDiff[(o.dsubtype >= 0 && o.dsubtype <= 7) ? o.dsubtype : 0], sizeof kErrDiff[0]);
        }
        if (o.dfamily == IW_DITHERFAMILY_RANDOM) {
          job->any_random = 1;
          int seed = (o.dsubtype == IW_DITHERSUBTYPE_SAMEPATTERN && oc.channeltype != IW_CHANNELTYPE_ALPHA)
                         ? ctx->random_seed : ctx->random_seed + oc.channeltype;
          o.prng_t1 = iwb_prng_first_state(seed);
        }
      }
    }
  }

  if (job->any_random) {
    iwb_affine tbl[64];
    iwb_prng_jump_table(tbl);
    job->d_jump = (iwb_affine_t *)job->upload(tbl, 64);
    if (!job->d_jump) return job;
  }

  job->negate = ctx->negate_target;
  job->out_alpha_ch = IW_IMGTYPE_HAS_ALPHA(ctx->img2.imgtype) ? ctx->out_nch - 1 : -1;
  // fully transparent pixels become black, unless reduced max codes are in use (reference: opt.c:1196-1200)
  job->clear_transparent = !out.is_float && !ctx->reduced_out_maxcode && job->out_alpha_ch >= 0;

  // chunk height: keep all FP32 intermediate planes of a chunk within ~48 MB so they live in L2
  {
    size_t per_row = (size_t)in_w * 4 * job->nproc;
    size_t rows = per_row ? (size_t)(48u << 20) / per_row : 64;
    rows = std::max<size_t>(rows, 16);
    rows = std::min<size_t>(rows, (size_t)out_h);
    job->chunk_rows = (int)rows;
  }

  iwb200_job_info &info = job->info;
  info.out_width = out_w; info.out_height = out_h; info.out_imgtype = ctx->img2.imgtype;
  info.out_bit_depth = ctx->img2.bit_depth; info.out_sampletype = ctx->img2.sampletype;
  info.out_channels = ctx->out_nch; info.out_bpr = out.bpr;
  info.in_width_phys = job->phys_w; info.in_height_phys = job->phys_h; info.in_bpr = ctx->img1.bpr;
  info.bandable = !job->any_errdiff;
  info.taps_v = taps_v; info.taps_h = taps_h;
#if !defined(IWB_EMU)
  const char *nofuse = iwb_option(ctx, "cuda:fused");
  if (!(nofuse && !strcmp(nofuse, "0"))) {
    fused_plan_build(job, &job->fused);
    const char *og = iwb_option(ctx, "cuda:gather");
    if (!job->fused.enabled && !(og && !strcmp(og, "0"))) gather_plan_build(job);
  }
  info.uses_fused_kernel = job->fused.enabled ? (job->fused.v2 ? 2 : 1) : (job->gather.enabled ? (job->gather.tile ? 5 : 4) : 0);
#endif
  job->flush_tables();
  return job;
}

static bool ensure_scratch(iwb200_job *job) {
  if (job->inter[0]) return true;
  const int W = job->out.width;
  const size_t rows = job->chunk_rows;
  for (int p = 0; p < job->nproc; p++) {
    job->inter[p] = (float *)job->scratch(rows * job->in_w * sizeof(float));
    if (!job->inter[p]) return false;
    const DevChanOut &o = job->cout[p];
    if (o.errdiff) {
      job->lin_plane[p] = (double *)job->scratch(rows * W * sizeof(double));
      job->errrows[p] = (double *)job->scratch((size_t)3 * W * sizeof(double));
      if (!job->lin_plane[p] || !job->errrows[p]) return false;
    }
    if (o.dfamily == IW_DITHERFAMILY_RANDOM && o.out_ch >= 0 && !job->out.is_float) {
      job->draws[p] = (uint32_t *)job->scratch(rows * W * sizeof(uint32_t));
      if (!job->draws[p]) return false;
    }
  }
  if (job->has_alpha) {
    job->final_alpha = (float *)job->scratch(rows * W * sizeof(float));
    if (!job->final_alpha) return false;
  }
  return true;
}

iwb200_job *iwb200_job_create(struct iw_context *ctx, int device) {
  if (!ctx) return nullptr;
  dev_guard on_device(device);
  if (ctx->use_count == 0) {  // standalone use: decide the plan now; the context is consumed like in iw_process_image
    ctx->use_count++;
    if (!iwb_prepare(ctx)) return nullptr;
  }
  std::string key = plan_key(ctx, device);
  // (context option cuda:plan_cache=0: build this context's job from scratch and do not park it afterwards)
  const char *pc = iwb_option(ctx, "cuda:plan_cache");
  const bool use_pool = pool_enabled() && !(pc && pc[0] == '0');
  if (!use_pool) key.clear();
  if (use_pool) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = g_pool.size(); i-- > 0;)
      if (g_pool[i]->key == key) {
        iwb200_job *job = g_pool[i];
        g_pool.erase(g_pool.begin() + i);
        job->ctx = ctx;
        job->errdiff_next_row = 0;
        job->reused = 1;
        return job;
      }
  }
  iwb200_job *job = job_build(ctx, device);
  if (!job->err.empty()) {
    iw_set_errorf(ctx, "%s", job->err.c_str());
    job_free(job);
    return nullptr;
  }
  job->key = std::move(key);
  return job;
}

int iwb200_job_get_info(const iwb200_job *job, struct iwb200_job_info *info) {
  if (!job || !info) return 0;
  *info = job->info;
  return 1;
}

int iwb200_job_band_input_rows(const iwb200_job *job, int r0, int r1, int *first, int *last) {
  if (!job || r0 < 0 || r1 > job->out.height || r0 >= r1) return 0;
  if (job->in.orient & 4) {  // transposed storage: an output band touches every stored row
    *first = 0; *last = job->phys_h - 1;
    return 1;
  }
  int lo = job->in_h, hi = -1;
  for (const iwb_axis_table &t : job->vtab) {
    lo = std::min(lo, t.min_src(r0, r1));
    hi = std::max(hi, t.max_src(r0, r1));
  }
  if (hi < lo) { lo = 0; hi = 0; }
  int a = job->in.y0 + lo, b = job->in.y0 + hi;  // logical rows
  if (job->in.orient & 2) { int t = a; a = job->in.height - 1 - b; b = job->in.height - 1 - t; }
  *first = a; *last = b;
  return 1;
}

int iwb200_job_run_device(iwb200_job *job, const void *d_in, size_t in_bpr, int in_row_origin, void *d_out,
                          size_t out_bpr, int r0, int r1, void *stream_) {
  if (!job) return 0;
  iwb_stream stream = (iwb_stream)stream_;
  if (r0 < 0 || r1 > job->out.height || r0 >= r1) return fail(job, "invalid output row band");
  if (stream_ && stream_ != job->streams[1]) job->used_external_stream = 1;
  if (job->any_errdiff && r0 != job->errdiff_next_row && r0 != 0)
    return fail(job, "error-diffusion dither must be processed in row order from row 0");
  if (job->float_prepass && !job->prepass_done)
    return fail(job, "device-resident floating-point input must be in a linear colorspace (host buffers are linearised by the library)");
  dev_guard on_device(job->device);
  DevInput in = job->in;
  in.pix = (const uint8_t *)d_in; in.bpr = in_bpr; in.row_origin = in_row_origin;
  if (job->x0_override >= 0) in.x0 = job->x0_override;
  DevOutput out = job->out;
  out.pix = (uint8_t *)d_out; out.bpr = out_bpr; out.row_origin = r0;

#if !defined(IWB_EMU)
  bool fast = job->fused.enabled != 0;
  if (fast && (job->fused.depth16 || job->fused.nco != job->fused.nch)) {   // these have only the TMA kernels: rows must be 16-byte aligned
    const uint8_t *s0 = in.pix + (ptrdiff_t)(in.y0 - in.row_origin) * (ptrdiff_t)in.bpr + (size_t)in.x0 * job->fused.nch * (job->fused.depth16 ? 2 : 1);
    fast = job->fused.v2 && in.bpr % 16 == 0 && ((uintptr_t)s0 & 15) == 0;
  }
  if (fast) {
    if (!fused_run(job, &job->fused, in, out, r0, r1, stream)) return 0;
    return 1;
  }
  if (job->gather.enabled) {                                      // (TMA: rows must be 16-byte aligned)
    const uint8_t *s0 = in.pix + (ptrdiff_t)(in.y0 - in.row_origin) * (ptrdiff_t)in.bpr + (size_t)in.x0 * job->gather.nch * (job->gather.depth16 ? 2 : 1);
    if (in.bpr % 16 == 0 && ((uintptr_t)s0 & 15) == 0) return gather_run(job, in, out, r0, r1, stream) ? 1 : 0;
  }
#endif
  if (!ensure_scratch(job)) return 0;
  job->order_after_stream0(stream);
  const int W = out.width;
  if (job->any_errdiff && r0 == 0)
    for (int p = 0; p < job->nproc; p++)
      if (job->errrows[p]) dev_zero(job->errrows[p], (size_t)3 * W * sizeof(double), stream);

  for (int c0 = r0; c0 < r1; c0 += job->chunk_rows) {
    const int c1 = std::min(r1, c0 + job->chunk_rows);
    const size_t rows = c1 - c0;
    job->last_path = 0; job->last_c0 = c0; job->last_c1 = c1;
    job->last_gen = 0;
    // vertical pass, all channels
    VPassParams vp;
    vp.in = in; vp.cm = job->cm; vp.nch = job->nproc; vp.in_w = job->in_w; vp.row0 = c0; vp.row1 = c1;
    for (int p = 0; p < job->nproc; p++) { vp.ch[p] = job->cin[p]; vp.ax[p] = job->vax[p]; vp.inter[p] = job->inter[p]; }
    { prof_scope ps(job, 0, stream); launch<VPassBody>(vp, (size_t)job->nproc * rows * job->in_w, stream); }

    if (job->any_random) {
      PrngParams pp;
      pp.nch = 0; pp.W = W; pp.row0 = c0; pp.rows = (int)rows; pp.seg = 64; pp.segs_per_row = (W + 63) / 64; pp.jump = job->d_jump;
      for (int p = 0; p < job->nproc; p++)
        if (job->draws[p]) { pp.t1[pp.nch] = job->cout[p].prng_t1; pp.draws[pp.nch] = job->draws[p]; pp.nch++; }
      { prof_scope ps(job, 2, stream); launch<PrngBody>(pp, (size_t)pp.nch * rows * pp.segs_per_row, stream); }
    }

    // horizontal pass + epilogue: alpha alone first, then every colour channel
    for (int pass = 0; pass < 2; pass++) {
      HPassParams hp;
      hp.cm = job->cm; hp.out = out; hp.in_w = job->in_w; hp.row0 = c0; hp.row1 = c1; hp.inter_row0 = c0;
      hp.final_alpha = job->final_alpha; hp.alpha_row0 = c0; hp.nch = 0;
      int lo = pass == 0 ? 0 : (job->has_alpha ? 1 : 0);
      int hi = pass == 0 ? (job->has_alpha ? 1 : 0) : job->nproc;
      for (int p = lo; p < hi; p++) {
        int k = hp.nch++;
        hp.ch[k] = job->cout[p]; hp.ax[k] = job->hax[p]; hp.inter[k] = job->inter[p];
        hp.lin_plane[k] = job->lin_plane[p]; hp.draws[k] = job->draws[p];
      }
      { prof_scope ps(job, 1, stream); launch<HPassBody>(hp, (size_t)hp.nch * rows * W, stream); }
    }

    if (job->any_errdiff) {
      ErrDiffParams ep;
      ep.out = out; ep.nch = 0; ep.row0 = c0; ep.row1 = c1;
      for (int p = 0; p < job->nproc; p++)
        if (job->cout[p].errdiff) {
          int k = ep.nch++;
          ep.ch[k] = job->cout[p]; ep.lin_plane[k] = job->lin_plane[p]; ep.err[k] = job->errrows[p];
          memcpy(ep.m[k], job->errdiff_m[p], sizeof ep.m[k]);
        }
#if !defined(IWB_EMU)
      // one warp per channel with the error rows in shared memory (errdiff.cuh) when a row fits; the per-thread
      // reference kernel otherwise
      const size_t ed_smem = iwb_errdiff_smem(W);
      if (ed_smem <= 220 * 1024 && !iwb_option(job->ctx, "cuda:errdiff_generic")) {
        if (!job->ed_ch_g) {
          job->ed_ch_g = (DevChanOut *)job->scratch(4 * sizeof(DevChanOut));
          if (!job->ed_ch_g) return 0;
          job->order_after_stream0(stream);
        }
        ErrDiffFastParams fp;
        fp.out = ep.out; fp.nch = ep.nch; fp.row0 = ep.row0; fp.row1 = ep.row1; fp.ch_g = job->ed_ch_g;
        for (int k = 0; k < ep.nch; k++) {
          fp.ch[k] = ep.ch[k]; fp.lin_plane[k] = ep.lin_plane[k]; fp.err[k] = ep.err[k];
          memcpy(fp.m[k], ep.m[k], sizeof fp.m[k]);
        }
        cudaMemcpyAsync(job->ed_ch_g, fp.ch, 4 * sizeof(DevChanOut), cudaMemcpyHostToDevice, (cudaStream_t)stream);
        static std::atomic<unsigned long long> attr{0};
        once_per_device(attr, [] {
          cudaFuncSetAttribute(iwb_errdiff_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
          cudaFuncSetAttribute(iwb_errdiff_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        });
        // Floyd-Steinberg on every channel: the specialised chain (errdiff.cuh)
        bool fs = true;
        for (int k = 0; k < ep.nch; k++)
          for (int i = 0; i < 12; i++) fs = fs && (ep.m[k][i] != 0.0) == (i == 0 || i == 3 || i == 4 || i == 5);
        prof_scope ps(job, 3, stream);
        if (fs) iwb_errdiff_warp<true><<<ep.nch, 32, ed_smem, (cudaStream_t)stream>>>(fp);
        else iwb_errdiff_warp<false><<<ep.nch, 32, ed_smem, (cudaStream_t)stream>>>(fp);
        g_launches++;
      } else
#endif
      { prof_scope ps(job, 3, stream); launch_serial<ErrDiffBody>(ep, ep.nch, stream); }
      job->errdiff_next_row = c1;
    }

    if (job->negate || job->clear_transparent) {
      PostParams pp;
      pp.out = out; pp.row0 = c0; pp.row1 = c1; pp.negate = job->negate; pp.alpha_ch = job->out_alpha_ch;
      pp.clear_transparent = job->clear_transparent;
      for (int c = 0; c < 4; c++) pp.maxcode[c] = c < job->ctx->out_nch ? job->ctx->out_ci[c].maxcode : 0;
      { prof_scope ps(job, 4, stream); launch<PostBody>(pp, rows * W, stream); }
    }
  }
#if !defined(IWB_EMU)
  if (cudaGetLastError() != cudaSuccess) return fail(job, "kernel launch failed");
#endif
  return 1;
}

// Verification hook: the FP32 intermediate (vertical-pass result, reference: src/imagew-main.c:905-912) of processed
// channel `ch` for the last chunk of output rows this job ran, as a row-major [rows][in_w] plane.
int iwb200_job_read_intermediate(iwb200_job *job, int ch, float *dst, size_t capacity, int *row0, int *rows, int *width) {
  if (job && ch < 0) ch = job->nproc - 1;                       // the channel processed last
  if (!job || job->last_path < 0 || ch < 0 || ch >= job->nproc) return 0;
  const int nrows = job->last_c1 - job->last_c0, w = job->in_w;
  if (row0) *row0 = job->last_c0;
  if (rows) *rows = nrows;
  if (width) *width = w;
  if (!dst) return 1;
  if (capacity < (size_t)nrows * w) return 0;
  dev_guard on_device(job->device);
#if !defined(IWB_EMU)
  cudaDeviceSynchronize();
#endif
  if (job->last_path == 0) {
    return dev_d2h(dst, job->inter[ch], (size_t)nrows * w * sizeof(float), 0, job->err) && dev_sync(0, job->err);
  }
#if !defined(IWB_EMU)
  if (job->last_path == 3) {                                     // gather path: interleaved by physical channel
    const int gn = job->gather.nch, gc = job->gather.chan_of_proc[ch];
    std::vector<float> tmp((size_t)nrows * job->last_pitch);
    if (!dev_d2h(tmp.data(), job->last_inter, tmp.size() * sizeof(float), 0, job->err) || !dev_sync(0, job->err)) return 0;
    for (int j = 0; j < nrows; j++)
      for (int x = 0; x < w; x++) dst[(size_t)j * w + x] = tmp[(size_t)j * job->last_pitch + (size_t)x * gn + gc];
    return 1;
  }
  const int nch = job->fused.nch;
  if (job->fused.nco != nch) ch = 0;                             // gray written as RGB: the three channels share one plane
  if (job->last_path == 2) {
    std::vector<float> tmp((size_t)nrows * job->last_pitch);
    if (!dev_d2h(tmp.data(), job->last_inter, tmp.size() * sizeof(float), 0, job->err) || !dev_sync(0, job->err)) return 0;
    for (int j = 0; j < nrows; j++)
      for (int x = 0; x < w; x++) dst[(size_t)j * w + x] = tmp[(size_t)j * job->last_pitch + job->last_sh + (size_t)x * nch + ch];
    return 1;
  }
  {
    const size_t pj = job->last_pitch;                          // generation 1: interT[byte column][row]
    std::vector<float> tmp((size_t)w * nch * pj);
    if (!dev_d2h(tmp.data(), job->last_inter, tmp.size() * sizeof(float), 0, job->err) || !dev_sync(0, job->err)) return 0;
    for (int j = 0; j < nrows; j++)
      for (int x = 0; x < w; x++) dst[(size_t)j * w + x] = tmp[((size_t)x * nch + ch) * pj + j];
    return 1;
  }
#else
  return 0;
#endif
}

void iwb200_job_set_profiling(iwb200_job *job, int on) { if (job) job->profiling = on; }
int iwb200_job_last_kernel_generation(const iwb200_job *job) { return job ? job->last_gen : -1; }
int iwb200_debug_fused_trap(char *buf, int buflen) {
#if !defined(IWB_EMU)
  const std::string d = iwb3_debug_text();
  if (buf && buflen > 0) { strncpy(buf, d.c_str(), (size_t)buflen - 1); buf[buflen - 1] = 0; }
  return (int)d.size();
#else
  if (buf && buflen > 0) buf[0] = 0;
  return 0;
#endif
}
void iwb200_job_reset_stats(iwb200_job *job) {
  if (!job) return;
  int n; double ms;
  iwb200_job_kernel_stats(job, 0, &n, &ms);  // drains pending events
  for (int k = 0; k < 8; k++) { job->stat_n[k] = 0; job->stat_ms[k] = 0; }
}
int iwb200_job_kernel_stats(iwb200_job *job, int which, int *launches, double *total_ms) {
  if (!job || which < 0 || which >= 8) return 0;
  dev_guard on_device(job->device);
#if !defined(IWB_EMU)
  for (auto &t : job->pending) {
    cudaEvent_t a = (cudaEvent_t)t.a, b = (cudaEvent_t)t.b;
    cudaEventSynchronize(b);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) == cudaSuccess) { job->stat_n[t.which]++; job->stat_ms[t.which] += ms; }
    cudaEventDestroy(a); cudaEventDestroy(b);
  }
#endif
  job->pending.clear();
  if (launches) *launches = job->stat_n[which];
  if (total_ms) *total_ms = job->stat_ms[which];
  return 1;
}

void iwb200_synth_xorshift(uint8_t *buf, size_t n, uint64_t seed) {
  uint64_t s = seed;
  for (size_t i = 0; i < n; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    buf[i] = (uint8_t)(s >> 24);
  }
}
uint64_t iwb200_fnv1a64(const uint8_t *buf, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; i++) { h ^= buf[i]; h *= 0x100000001b3ull; }
  return h;
}

int iwb200_job_run_host(iwb200_job *job, const void *h_in, size_t in_bpr, void *h_out, size_t out_bpr, int r0,
                        int r1) {
  if (!job) return 0;
  dev_guard on_device(job->device);                             // (every allocation below belongs to the job's device)
  int first, last;
  if (!iwb200_job_band_input_rows(job, r0, r1, &first, &last)) return fail(job, "invalid output row band");
  first = std::max(first, 0);
  last = std::min(last, job->phys_h - 1);
  // float samples in a non-linear colourspace -> linear doubles, on the host cores (see job_build)
  struct pre_buf { void *p = nullptr; bool pinned = false; ~pre_buf() { if (p && !(pinned && iwb_pinned_free(p))) free(p); } } pre;
  if (job->float_prepass) {
    const size_t nsamp = (size_t)job->phys_w * job->in.nch_phys, pre_bpr = nsamp * sizeof(double);
    const size_t bytes = (size_t)(last - first + 1) * pre_bpr;
    pre.p = iwb_pinned_alloc(bytes);
    pre.pinned = pre.p != nullptr;
    if (!pre.p) pre.p = malloc(bytes ? bytes : 1);
    if (!pre.p) return fail(job, "out of memory (linearised input)");
    const iw_context *ctx = job->ctx;
    const int nch = job->in.nch_phys;
    const int alpha_ch = IW_IMGTYPE_HAS_ALPHA(ctx->img1.imgtype) ? nch - 1 : -1;
    const iw_csdescr cs = ctx->img1cs;
    const uint8_t *src = (const uint8_t *)h_in;
    double *dstp = (double *)pre.p;
    const int nrows = last - first + 1;
    auto work = [&](int ra, int rb) {
      for (int r = ra; r < rb; r++) {
        const uint8_t *row = src + (size_t)(first + r) * in_bpr;
        double *d = dstp + (size_t)r * nsamp;
        for (size_t i = 0; i < nsamp; i++) {
          float f;
          memcpy(&f, row + 4 * i, 4);
          const double v = (double)f;
          d[i] = ((int)(i % (size_t)nch) == alpha_ch) ? v : iwb_to_linear(v, &cs);
        }
      }
    };
    int nthr = (int)std::min<size_t>(16, std::max<size_t>(1, (size_t)nrows * nsamp / 65536));
    nthr = std::min(nthr, std::max(1, (int)std::thread::hardware_concurrency()));
    if (nthr <= 1) work(0, nrows);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nthr; t++) th.emplace_back(work, (int)((long long)nrows * t / nthr), (int)((long long)nrows * (t + 1) / nthr));
      for (auto &t : th) t.join();
    }
    h_in = (const uint8_t *)pre.p - (size_t)first * pre_bpr;     // rows are addressed by their physical index below
    in_bpr = pre_bpr;
  }
  job->prepass_done = job->float_prepass;
  // The device copy of the input gets a row pitch that is a multiple of 16 bytes (the TMA kernels need aligned rows);
  // the pitch conversion costs nothing: it is done by the copy engine (2-D H2D copy).
  size_t d_bpr = in_bpr;
  size_t col_off = 0, row_bytes = in_bpr;      // the part of each host row that travels
#if !defined(IWB_EMU)
  if (((job->fused.enabled && job->fused.v2) || job->gather.enabled) && !iwb_option(job->ctx, "cuda:keep_pitch")) {
    // a crop in x: only the cropped columns are copied, and they start an aligned device row (otherwise an odd crop
    // origin would rule the TMA kernels out, and the columns outside the crop would cross PCIe for nothing)
    const size_t bpp = (size_t)job->in.nch_phys * (job->in.depth / 8);
    if (job->in.orient == 0 && job->in.x0 > 0 && !job->in.is_float && (job->in.depth == 8 || job->in.depth == 16)) {
      col_off = (size_t)job->in.x0 * bpp;
      row_bytes = (size_t)job->in_w * bpp;
      d_bpr = row_bytes;
    }
    if (d_bpr % 16 != 0) d_bpr = (d_bpr + 15) & ~(size_t)15;
  }
#endif
  const size_t in_bytes = (size_t)(last - first + 1) * d_bpr;
  const size_t out_bytes = (size_t)(r1 - r0) * out_bpr;
  if (job->d_in_bytes < in_bytes) { dev_free(job->d_in); job->d_in = dev_alloc(in_bytes, job->err); job->d_in_bytes = in_bytes; }
  if (job->d_out_bytes < out_bytes) { dev_free(job->d_out); job->d_out = dev_alloc(out_bytes, job->err); job->d_out_bytes = out_bytes; }
  if (!job->d_in || !job->d_out) { job->d_in_bytes = job->d_out_bytes = 0; return 0; }
  iwb_timings &tm = job->tm_sink ? *job->tm_sink : job->ctx->timings;
#if defined(IWB_EMU)
  (void)tm; (void)col_off; (void)row_bytes;
  if (!dev_h2d(job->d_in, (const uint8_t *)h_in + (size_t)first * in_bpr, in_bytes, 0, job->err)) return 0;
  if (!iwb200_job_run_device(job, job->d_in, in_bpr, first, job->d_out, out_bpr, r0, r1, 0)) return 0;
  if (!dev_d2h(h_out, job->d_out, out_bytes, 0, job->err)) return 0;
  return 1;
#else
  // Software pipeline over output-row sub-bands: while sub-band k computes, the input rows of k+1 arrive and the
  // output rows of k-1 leave (three streams, page-locked host buffers).  The kernels take a fraction of the PCIe
  // time, so the call costs about one H2D of the image.  Sequential (error-diffusion) jobs use one sub-band.
  if (!job->streams[0])
    for (int k = 0; k < 3; k++) { cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking); job->streams[k] = st; }
  cudaStream_t s_in = (cudaStream_t)job->streams[0], s_k = (cudaStream_t)job->streams[1], s_out = (cudaStream_t)job->streams[2];
  int nsub = 1;
  const bool transposed = (job->in.orient & 4) != 0, flipped = (job->in.orient & 2) != 0;
  // (sub-bands of a small image do not pay: three sub-bands of a 1080p -> 256x144 job measured 0.26 ms against 0.24 ms
  // in one piece -- the short kernels are inefficient and their launches cost as much as the copies they would hide)
  if (job->info.bandable && !transposed && !flipped && in_bytes >= ((size_t)8 << 20)) nsub = (int)std::min<size_t>(6, (size_t)(r1 - r0) / 256);
  if (const char *o = iwb_option(job->ctx, "cuda:pipeline")) nsub = std::max(1, atoi(o));
  nsub = std::max(1, std::min(nsub, r1 - r0));
  while ((int)job->events.size() < 2 * nsub + 2) { cudaEvent_t e; cudaEventCreate(&e); job->events.push_back(e); }
  auto ev = [&](int i) { return (cudaEvent_t)job->events[i]; };
  // the default stream may still be uploading this job's tables
  cudaEventRecord(ev(2 * nsub), 0);
  cudaStreamWaitEvent(s_k, ev(2 * nsub), 0);
  cudaStreamWaitEvent(s_in, ev(2 * nsub), 0);                   // (the staging buffers were allocated there too)
  job->s0_dirty = false; job->s0_stream = (void *)s_k;
  cudaEventRecord(ev(2 * nsub + 1), s_in);       // timing origin
  int copied_to = first - 1;                     // last physical input row already requested
  bool ok = true;
  // sub-band boundaries: later sub-bands are shorter, because the call ends with the compute and the copy-out of
  // the LAST one after the last input byte has arrived (everything else hides behind the input copy)
  auto cut = [&](int kk) {
    if (kk <= 0) return r0;
    if (kk >= nsub) return r1;
    const double x = (double)kk / nsub;
    int r = r0 + (int)((double)(r1 - r0) * (1.0 - pow(1.0 - x, 1.6)));
    return std::min(r1, std::max(r0, (r / 8) * 8));
  };
  auto bounds = [&](int k, int *b0, int *b1) {
    *b0 = cut(k); *b1 = cut(k + 1);
    if (*b1 <= *b0 || (nsub > 1 && (r1 - r0) < 64 * nsub)) {     // (degenerate sizes: equal split)
      *b0 = r0 + (int)(((long long)(r1 - r0) * k) / nsub);
      *b1 = r0 + (int)(((long long)(r1 - r0) * (k + 1)) / nsub);
    }
  };
  // all input copies are queued first, so that the copy engine never waits for this thread to get through the kernel
  // launches of an earlier sub-band (that matters for small images, where a launch costs as much as a copy)
  for (int k = 0; k < nsub && ok; k++) {
    int b0, b1, f, l;
    bounds(k, &b0, &b1);
    iwb200_job_band_input_rows(job, b0, b1, &f, &l);
    l = std::min(l, last);
    if (nsub == 1 || k == nsub - 1) l = last;
    if (l > copied_to) {
      const size_t off = (size_t)(copied_to + 1 - first) * d_bpr, n = (size_t)(l - copied_to) * in_bpr;
      if (d_bpr == in_bpr && col_off == 0) {
        ok = dev_h2d((uint8_t *)job->d_in + off, (const uint8_t *)h_in + (size_t)(copied_to + 1) * in_bpr, n, s_in, job->err);
      } else {
        ok = cudaMemcpy2DAsync((uint8_t *)job->d_in + off, d_bpr, (const uint8_t *)h_in + (size_t)(copied_to + 1) * in_bpr + col_off, in_bpr,
                               row_bytes, (size_t)(l - copied_to), cudaMemcpyHostToDevice, s_in) == cudaSuccess;
        if (!ok) job->err = "H2D copy (2-D) failed";
      }
      copied_to = l;
    }
    cudaEventRecord(ev(2 * k), s_in);
  }
  for (int k = 0; k < nsub && ok; k++) {
    int b0, b1;
    bounds(k, &b0, &b1);
    cudaStreamWaitEvent(s_k, ev(2 * k), 0);
    job->x0_override = col_off ? 0 : -1;
    ok = ok && iwb200_job_run_device(job, job->d_in, d_bpr, first, (uint8_t *)job->d_out + (size_t)(b0 - r0) * out_bpr, out_bpr,
                                     b0, b1, s_k);
    job->x0_override = -1;
    cudaEventRecord(ev(2 * k + 1), s_k);
    cudaStreamWaitEvent(s_out, ev(2 * k + 1), 0);
    ok = ok && dev_d2h((uint8_t *)h_out + (size_t)(b0 - r0) * out_bpr, (uint8_t *)job->d_out + (size_t)(b0 - r0) * out_bpr,
                       (size_t)(b1 - b0) * out_bpr, s_out, job->err);
  }
  if (!dev_sync(s_out, job->err)) return 0;
  if (!dev_sync(s_k, job->err) || !dev_sync(s_in, job->err)) return 0;
  if (!ok) return 0;
  float ms = 0;
  // with overlap the three phases are concurrent: report copy-in span, compute span after the last copy-in,
  // and copy-out span after the last kernel
  cudaEventElapsedTime(&ms, ev(2 * nsub + 1), ev(2 * (nsub - 1))); tm.h2d_ms = ms;
  cudaEventElapsedTime(&ms, ev(2 * (nsub - 1)), ev(2 * (nsub - 1) + 1)); tm.kernel_ms = ms;
  cudaEvent_t done; cudaEventCreate(&done); cudaEventRecord(done, s_out); cudaEventSynchronize(done);
  cudaEventElapsedTime(&ms, ev(2 * (nsub - 1) + 1), done); tm.d2h_ms = ms;
  cudaEventDestroy(done);
  return 1;
#endif
}


// ---- post-pass optimiser: device scan -> host decisions -> device rewrite (optimize.h, optimize.cpp) -----------------
// `jobs[i]` has just produced output rows bands[i] = [r0, r1) of the image in its device output buffer (row pitch
// ctx->img2.bpr) and copied them to ctx->img2.pixels.  On return the context's result view (iw_get_output_image) is
// the optimised image; the image as processed stays in ctx->img2.
int iwb_optimize_output(struct iw_context *ctx, iwb200_job **jobs, const int *bands, int n) {
  if (ctx->img2.sampletype != IW_SAMPLETYPE_UINT || ctx->reduced_out_maxcode) return 1;   // (src/imagew-opt.c:1196-1204)
  bool any = false;
  for (int k = 1; k <= 5; k++) any = any || ctx->opt_flags[k];
  if (!any) return 1;
  const int W = ctx->img2.width, H = ctx->img2.height;
  const int nch = iw_imgtype_num_channels(ctx->img2.imgtype), bps = ctx->img2.bit_depth / 8;
  const unsigned palbits = IW_PROFILE_PAL1 | IW_PROFILE_PAL2 | IW_PROFILE_PAL4 | IW_PROFILE_PAL8 | IW_PROFILE_GRAY1 | IW_PROFILE_GRAY2 | IW_PROFILE_GRAY4;
  const int want_colors = (ctx->output_profile & palbits) ? 1 : 0;
  IwbOptScan merged;
  memset(&merged, 0, sizeof merged);
  std::vector<IwbOptScan> part(n);
  for (int i = 0; i < n; i++) {
    iwb200_job *job = jobs[i];
    dev_guard on_device(job->device);
    iwb_stream st = (iwb_stream)job->streams[1];
    if (!job->d_scan) {
      job->d_scan = (IwbOptScan *)job->scratch(sizeof(IwbOptScan));
      job->d_recipe = (OptRecipe *)job->scratch(sizeof(OptRecipe));
      if (!job->d_scan || !job->d_recipe) return 0;
      job->order_after_stream0(st);
    }
    dev_zero(job->d_scan, sizeof(IwbOptScan), st);
    OptScanParams sp;
    sp.out = job->out; sp.out.pix = (uint8_t *)job->d_out; sp.out.bpr = ctx->img2.bpr; sp.out.row_origin = bands[2 * i];
    sp.row0 = bands[2 * i]; sp.row1 = bands[2 * i + 1];
    sp.nch = nch; sp.bps = bps; sp.has_alpha = IW_IMGTYPE_HAS_ALPHA(ctx->img2.imgtype) ? 1 : 0; sp.want_colors = want_colors;
    sp.res = job->d_scan;
    launch<OptScanBody>(sp, (size_t)(sp.row1 - sp.row0) * W, st);
    if (!dev_d2h(&part[i], job->d_scan, sizeof(IwbOptScan), st, job->err)) return 0;
  }
  for (int i = 0; i < n; i++) {
    dev_guard on_device(jobs[i]->device);
    if (!dev_sync((iwb_stream)jobs[i]->streams[1], jobs[i]->err)) return 0;
    if (i == 0) { merged = part[0]; if (merged.ncolors > 256) merged.ncolors = 257; }
    else iwb_opt_merge(&merged, part[i]);
  }
  IwbOptOutcome oc;
  if (!iwb_opt_decide(ctx, merged, &oc)) return 0;
  ctx->res_has_colorkey = oc.has_colorkey;
  for (int k = 0; k < 3; k++) ctx->res_colorkey[k] = oc.colorkey[k];
  if (!oc.changed) return 1;
  iw_byte *dst = (iw_byte *)iw_malloc_large(ctx, oc.bpr, (size_t)H);
  if (!dst) return 0;
  for (int i = 0; i < n; i++) {
    iwb200_job *job = jobs[i];
    dev_guard on_device(job->device);
    iwb_stream st = (iwb_stream)job->streams[1];
    const int r0 = bands[2 * i], r1 = bands[2 * i + 1];
    const size_t bytes = (size_t)(r1 - r0) * oc.bpr;
    if (job->d_opt_bytes < bytes) {
      dev_free(job->d_opt);
      job->d_opt = dev_alloc(bytes, job->err);
      job->d_opt_bytes = job->d_opt ? bytes : 0;
      if (!job->d_opt) { iw_free(ctx, dst); return 0; }
      job->s0_dirty = true;
      job->order_after_stream0(st);
    }
    if (!dev_h2d(job->d_recipe, &oc.recipe, sizeof(OptRecipe), st, job->err)) { iw_free(ctx, dst); return 0; }
    OptRewriteParams rp;
    rp.out = job->out; rp.out.pix = (uint8_t *)job->d_out; rp.out.bpr = ctx->img2.bpr; rp.out.row_origin = r0;
    rp.row0 = r0; rp.row1 = r1; rp.dst = (uint8_t *)job->d_opt; rp.dst_bpr = oc.bpr; rp.rc = job->d_recipe;
    launch<OptRewriteBody>(rp, (size_t)(r1 - r0) * W, st);
    if (!dev_d2h(dst + (size_t)r0 * oc.bpr, job->d_opt, bytes, st, job->err)) { iw_free(ctx, dst); return 0; }
  }
  for (int i = 0; i < n; i++) {
    dev_guard on_device(jobs[i]->device);
    if (!dev_sync((iwb_stream)jobs[i]->streams[1], jobs[i]->err)) { iw_free(ctx, dst); return 0; }
  }
  if (ctx->opt_pixels) iw_free(ctx, ctx->opt_pixels);
  ctx->opt_pixels = dst;
  ctx->res_type = oc.imgtype; ctx->res_depth = oc.bit_depth; ctx->res_bpr = oc.bpr; ctx->res_pixels = dst;
  return 1;
}

// One image on several GPUs (context option cuda:devices, process.cpp).  The job is split into output-row bands, one per
// listed device; each band has its own job (plan tables on its device), its own host thread and stream set, copies only
// its input rows + filter halo host -> device and writes its output rows device -> host straight into the caller's
// buffer, so the PCIe links of the devices work in parallel and nothing is exchanged between GPUs.  Dither phase,
// checkerboard phase and PRNG draw index are absolute coordinates, so the bands concatenate bit-exactly.
// The same device may be listed more than once (several bands in flight on one GPU).
int iwb200_run_host_multi(struct iw_context *ctx, const int *devices, int ndev, const void *h_in, size_t in_bpr, void *h_out, size_t out_bpr) {
  if (!ctx || ndev < 1) return 0;
  const int out_h = ctx->img2.height;
  ndev = std::max(1, std::min(ndev, std::max(1, out_h / 32)));  // (at least 32 rows per band)
  std::vector<iwb200_job *> jobs(ndev, nullptr);
  bool ok = true;
  for (int i = 0; i < ndev && ok; i++) {                        // (sequential: job creation reports errors on the context)
    jobs[i] = iwb200_job_create(ctx, devices[i]);
    if (!jobs[i]) ok = false;
    else if (!jobs[i]->info.bandable) {
      iw_set_error(ctx, "CUDA backend: cuda:devices needs a job that can be split into row bands; error-diffusion dithering cannot");
      ok = false;
    }
  }
  std::vector<iwb_timings> tms(ndev);
  std::vector<int> res(ndev, 0);
  if (ok) {
    std::vector<std::thread> th;
    for (int i = 0; i < ndev; i++) {
      // band boundaries on multiples of 32 rows (whole groups of the horizontal kernels)
      int r0 = (int)(((long long)out_h * i / ndev) / 32 * 32), r1 = i + 1 == ndev ? out_h : (int)(((long long)out_h * (i + 1) / ndev) / 32 * 32);
      jobs[i]->tm_sink = &tms[i];
      th.emplace_back([&, i, r0, r1]() {
        res[i] = r1 > r0 ? iwb200_job_run_host(jobs[i], h_in, in_bpr, (uint8_t *)h_out + (size_t)r0 * out_bpr, out_bpr, r0, r1) : 1;
      });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < ndev; i++) {
      jobs[i]->tm_sink = nullptr;
      if (!res[i]) { ok = false; iw_set_errorf(ctx, "CUDA backend (device %d): %s", devices[i], iwb200_job_last_error(jobs[i])); }
      ctx->timings.h2d_ms = std::max(ctx->timings.h2d_ms, tms[i].h2d_ms);
      ctx->timings.kernel_ms = std::max(ctx->timings.kernel_ms, tms[i].kernel_ms);
      ctx->timings.d2h_ms = std::max(ctx->timings.d2h_ms, tms[i].d2h_ms);
    }
  }
  if (ok) {
    std::vector<int> bands;
    for (int i = 0; i < ndev; i++) {
      bands.push_back((int)(((long long)out_h * i / ndev) / 32 * 32));
      bands.push_back(i + 1 == ndev ? out_h : (int)(((long long)out_h * (i + 1) / ndev) / 32 * 32));
    }
    iwb_set_result_view(ctx, 0, out_h);
    if (!iwb_optimize_output(ctx, jobs.data(), bands.data(), ndev)) {
      ok = false;
      if (!ctx->error_flag) iw_set_error(ctx, "CUDA backend: output optimisation failed");
    }
  }
  for (iwb200_job *j : jobs) if (j) iwb200_job_destroy(j);
  return ok ? 1 : 0;
}


struct iwb200_batch;
void iwb200_batch_destroy(iwb200_batch *b);

// ---- batches of same-shaped images (BASELINE configuration 4: 1024 thumbnails) ------------------------------------------
// One plan for the whole batch (the context describes the job: input header, size, filter, dither ...; its pixels are
// not used).  Every listed device gets `depth` slots, each with its own job (tables on that device), staging buffers and
// stream.  iwb200_batch_submit() queues, on the next slot in round-robin order, the host -> device copy of one image, its
// kernels and the device -> host copy of its result -- all asynchronous, so image k+1 crosses PCIe while image k
// computes and image k-1 leaves; the call only blocks when the slot it wants is still busy with an earlier image.
// Host buffers should be page-locked (iw_malloc_large memory is); pageable memory works but serialises the copies.
struct iwb200_batch {
  struct slot { iwb200_job *job = nullptr; void *d_in = nullptr, *d_out = nullptr; void *stream = nullptr, *done = nullptr; bool busy = false; };
  iw_context *ctx = nullptr;
  std::vector<slot> slots;
  size_t next = 0;
  size_t in_bytes = 0, out_bytes = 0, d_bpr = 0, in_row_bytes = 0;
  int in_rows = 0, out_rows = 0;
  std::string err;
};

static bool batch_finish_slot(iwb200_batch *b, iwb200_batch::slot &s) {
  if (!s.busy) return true;
  s.busy = false;
#if !defined(IWB_EMU)
  dev_guard on_device(s.job->device);
  if (const char *m = dev_check(cudaEventSynchronize((cudaEvent_t)s.done))) { b->err = std::string("batch: ") + m; return false; }
#endif
  return true;
}

iwb200_batch *iwb200_batch_create(struct iw_context *ctx, const int *devices, int ndev, int depth) {
  if (!ctx || ndev < 1 || !devices) return nullptr;
  depth = std::max(1, std::min(depth, 16));
  iwb200_batch *b = new iwb200_batch();
  b->ctx = ctx;
  for (int d = 0; d < ndev; d++)
    for (int k = 0; k < depth; k++) {
      iwb200_batch::slot s;
      s.job = iwb200_job_create(ctx, devices[d]);
      if (!s.job) { b->err = "batch: could not build the plan"; b->slots.push_back(s); iwb200_batch_destroy(b); return nullptr; }
      if (s.job->any_errdiff || s.job->float_prepass) {
        iw_set_error(ctx, "CUDA backend: batches do not take error-diffusion dithering or non-linear floating-point input");
        b->slots.push_back(s); iwb200_batch_destroy(b); return nullptr;
      }
      dev_guard on_device(s.job->device);
      const iwb200_job_info &info = s.job->info;
      b->in_rows = info.in_height_phys; b->out_rows = info.out_height;
      b->in_row_bytes = info.in_bpr;
      b->d_bpr = (info.in_bpr + 15) & ~(size_t)15;              // (rows start on 16 bytes on the device: the TMA kernels)
      b->in_bytes = b->d_bpr * (size_t)info.in_height_phys;
      b->out_bytes = info.out_bpr * (size_t)info.out_height;
      s.d_in = dev_alloc(b->in_bytes, b->err);
      s.d_out = dev_alloc(b->out_bytes, b->err);
#if !defined(IWB_EMU)
      cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking); s.stream = st;
      cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); s.done = ev;
      cudaEventRecord(ev, 0);
      cudaStreamWaitEvent(st, ev, 0);                           // (the buffers were allocated on the default stream)
#endif
      b->slots.push_back(s);
      if (!s.d_in || !s.d_out) { iwb200_batch_destroy(b); return nullptr; }
    }
  return b;
}

int iwb200_batch_submit(iwb200_batch *b, const void *h_in, size_t in_bpr, void *h_out, size_t out_bpr) {
  if (!b || !h_in || !h_out) return 0;
  iwb200_batch::slot &s = b->slots[b->next];
  b->next = (b->next + 1) % b->slots.size();
  if (!batch_finish_slot(b, s)) return 0;
  iwb200_job *job = s.job;
  dev_guard on_device(job->device);
  const iwb200_job_info &info = job->info;
  if (out_bpr != info.out_bpr || in_bpr < b->in_row_bytes) { b->err = "batch: row strides do not match the plan"; return 0; }
#if defined(IWB_EMU)
  for (int r = 0; r < b->in_rows; r++) memcpy((uint8_t *)s.d_in + (size_t)r * b->d_bpr, (const uint8_t *)h_in + (size_t)r * in_bpr, b->in_row_bytes);
  if (!iwb200_job_run_device(job, s.d_in, b->d_bpr, 0, s.d_out, out_bpr, 0, b->out_rows, nullptr)) { b->err = job->err; return 0; }
  memcpy(h_out, s.d_out, b->out_bytes);
#else
  cudaStream_t st = (cudaStream_t)s.stream;
  cudaError_t e;
  if (in_bpr == b->d_bpr) e = cudaMemcpyAsync(s.d_in, h_in, b->in_bytes, cudaMemcpyHostToDevice, st);
  else e = cudaMemcpy2DAsync(s.d_in, b->d_bpr, h_in, in_bpr, b->in_row_bytes, (size_t)b->in_rows, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { b->err = std::string("batch: H2D copy: ") + cudaGetErrorString(e); return 0; }
  if (!iwb200_job_run_device(job, s.d_in, b->d_bpr, 0, s.d_out, out_bpr, 0, b->out_rows, st)) { b->err = job->err; return 0; }
  e = cudaMemcpyAsync(h_out, s.d_out, b->out_bytes, cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) { b->err = std::string("batch: D2H copy: ") + cudaGetErrorString(e); return 0; }
  cudaEventRecord((cudaEvent_t)s.done, st);
#endif
  s.busy = true;
  return 1;
}

int iwb200_batch_wait(iwb200_batch *b) {
  if (!b) return 0;
  bool ok = true;
  for (auto &s : b->slots) ok = batch_finish_slot(b, s) && ok;
  return ok ? 1 : 0;
}

const char *iwb200_batch_last_error(const iwb200_batch *b) { return b ? b->err.c_str() : "null batch"; }
int iwb200_batch_get_info(const iwb200_batch *b, struct iwb200_job_info *info) {
  return (b && !b->slots.empty() && b->slots[0].job) ? iwb200_job_get_info(b->slots[0].job, info) : 0;
}

void iwb200_batch_destroy(iwb200_batch *b) {
  if (!b) return;
  for (auto &s : b->slots) {
    if (!s.job) continue;
    dev_guard on_device(s.job->device);
#if !defined(IWB_EMU)
    if (s.stream) { cudaStreamSynchronize((cudaStream_t)s.stream); cudaStreamDestroy((cudaStream_t)s.stream); }
    if (s.done) cudaEventDestroy((cudaEvent_t)s.done);
    s.job->used_external_stream = 0;                            // (synchronised just above)
#endif
    dev_free(s.d_in);
    dev_free(s.d_out);
    iwb200_job_destroy(s.job);
  }
  delete b;
}

}  // extern "C"
