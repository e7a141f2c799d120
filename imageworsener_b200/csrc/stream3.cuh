// stream3.cuh -- the fused vertical + horizontal streaming kernel: parameters and host entry point.
//
// Generation 2 (stream2.cuh) runs the two passes as two kernels; the vertical one leaves the FP64 pipe idle ~30 % of
// the time, the horizontal one ~60 % (its epilogue is FP32 / integer / shared-memory work), and the FP32 intermediate
// between them makes a round trip through HBM.  Generation 3 is ONE persistent kernel with warp-specialised CTAs (one
// per SM): V warps and H warps are resident together, so the FP64 pipe sees the union of both instruction mixes, and H
// consumes a group of 32 intermediate rows as soon as the V warps that produce it have announced it -- while it is
// still in L2.  See stream3_dev.cuh for the synchronisation; the arithmetic is generation 2's, statement for statement
// (reference: src/imagew-resize.c:346-361, src/imagew-main.c:831-927, :929-1155).
#pragma once
#include <string>

#include "stream2.cuh"

#if !defined(IWB_EMU)

#ifndef IWB3_WARPS
#define IWB3_WARPS 16            // warps per CTA (one CTA per SM): 128 registers per thread
#endif
#ifndef IWB3_NV
#define IWB3_NV 10               // of which vertical (the rest horizontal); run-time tunable (cuda:nv)
#endif

// Cross-CTA state of the launches of one (job, row range): two ticket counters and one counter per (group of 32
// output rows, block of `cbw` warp columns).  Nothing is reset between launches: launch number `gen` (1, 2, ...) of the
// same geometry takes tickets [ (gen-1)*per_launch, ... ) and waits for gen * (rows x warp columns) announcements.
struct VH3Sync {
  int *v_head, *h_head;
  int *cnt;                      // [groups][ncb]
  int ncb, cbw;
  int gen;
  int v_base, h_base;            // first ticket of this launch
  int *dbg;                      // diagnostic slot (page-locked host memory, device-visible): [0] = records, [8 + 8k...] = record k
};

struct VH3Params {
  V2Params v;                    // part = row parts
  H2Params h;                    // part = pixel segments
  VH3Sync sy;
  const int *h_order;            // [groups] groups in the order they are expected to complete
  int nv;                        // vertical warps per CTA
  int v_warp_bytes, h_warp_bytes, v_ring_bytes, h_ring_bytes;   // shared memory per warp / of which pipeline ring
  int n_vitems, n_hitems;
  int nwc;                       // warp columns (128 sample columns each)
  int n_gv, n_gh;                // general parts of the two partitions (queued first)
  int nparts_h;
};

// Host-side state kept by the job for one row range [c0, c1) (device memory behind the counters, launch count)
struct VH3State {
  int c0 = -1, c1 = -1;
  int *dev = nullptr;            // [2 + groups*ncb + groups] ints: heads, counters, h_order
  int gen = 0;
  int groups = 0, ncb = 0, cbw = 0;
  long long v_per_launch = 0, h_per_launch = 0;   // tickets taken per launch (items + one failing pop per warp)
  int nv = 0, grid = 0;
  bool order_uploaded = false;
};
// The diagnostic slot of the current device (128 ints of mapped page-locked memory, allocated on first use), and its
// content as text after a failed launch (empty if nothing was recorded).
int *iwb3_debug_slot();
std::string iwb3_debug_text();

bool iwb3_supported(int SV, int SH, int nch, int nco, int epi, int depth16);
// false: not applicable to this launch (the caller falls back to the two-kernel path) or launch problem (err set)
bool iwb3_run(const V2Job &vj, const H2Job &hj, VH3State *st, int nv_override, int seg_override, cudaStream_t stream, bool *launched);

#endif  // !IWB_EMU
