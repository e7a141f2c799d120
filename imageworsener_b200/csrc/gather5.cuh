// gather5.cuh -- tile kernels for images with alpha and filter windows of more than 8 taps (design: gather5.cu).
#pragma once
#include "gather4.cuh"

#if !defined(IWB_EMU)
#define IWT_MAXW 64              // longest window (taps, rounded up to even)
bool iwt_supported(int WV, int WH, int nch);
// first_host: host copy of the vertical schedule's window starts (strip planning)
bool iwt_run_v(const GVJob &job, const int *first_host, cudaStream_t st);
// cm_g: the job's DevCommon in device memory; max_nb: most input pixels any 32-pixel output segment's windows span
bool iwt_run_h(const GHJob &job, const DevCommon *cm_g, int max_nb, cudaStream_t st);
#endif
