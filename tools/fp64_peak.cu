// fp64_peak.cu -- measures the independent DMUL+DADD issue rate of the GPU (the roof that binds the tap loops,
// since bit-exactness forbids FMA).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, bool FMA>
__global__ void k(double *out, double w, int iters) {
  double a[ILP], v[ILP];
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 1e-9 + i; v[i] = 1.0 + i * 1e-6; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (FMA) a[i] = __fma_rn(v[i], w, a[i]);
      else a[i] = __dadd_rn(a[i], __dmul_rn(v[i], w));
    }
  }
  double s = 0;
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, bool FMA>
void run(const char *name, int blocks, int threads) {
  double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
  const int iters = 4096;
  k<ILP, FMA><<<blocks, threads>>>(d, 1.0000001, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP, FMA><<<blocks, threads>>>(d, 1.0000001, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double instr = (double)blocks * threads * iters * ILP * (FMA ? 1 : 2);
  printf("%-28s blocks=%d threads=%d  %.3f ms  %.2f T lane-instr/s  (%.1f lanes/clk/SM at 1.965 GHz, 148 SMs)\n", name, blocks,
         threads, ms, instr / ms / 1e9, instr / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(d);
}
int main() {
  run<8, false>("dmul+dadd ilp8", 148 * 8, 256);
  run<16, false>("dmul+dadd ilp16", 148 * 8, 256);
  run<16, false>("dmul+dadd ilp16 occ4", 148 * 4, 128);
  run<24, false>("dmul+dadd ilp24 occ4x128", 148 * 4, 128);
  run<16, true>("dfma ilp16", 148 * 8, 256);
  return 0;
}
