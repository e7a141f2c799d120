// issue_probe.cu -- what does an FP64 instruction cost the warp scheduler?  Per DMUL+DADD pair the loop issues NX further
// independent full-rate instructions (FP32 adds).  If they fitted into the second cycle an FP64 instruction occupies its
// half-rate pipe, a pair would cost max(4, 2 + NX) cycles per scheduler; if an FP64 instruction holds the issue port for
// both cycles it costs 4 + NX.  Run with enough warps per scheduler that latency does not matter.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o issue_probe issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NX>
__global__ void __launch_bounds__(128) k(double *out, double w0, int iters, float g0) {
  constexpr int ILP = 16;
  double a[ILP];
  float f[8];
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 8; i++) f[i] = g0 * i;
  double w = w0;
  float g = g0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      a[i] = __dadd_rn(a[i], __dmul_rn(a[(i + ILP / 2) % ILP], w));
#pragma unroll
      for (int x = 0; x < NX; x++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[(i * NX + x) & 7]) : "f"(g));
    }
    w = __longlong_as_double(__double_as_longlong(w) + 1);
    g += 1.0f;
  }
  double s = 0;
  for (int i = 0; i < ILP; i++) s += a[i];
  for (int i = 0; i < 8; i++) s += (double)f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NX>
void run(int warps_per_sched) {
  const int blocks = 148 * warps_per_sched, threads = 128, iters = 2048;
  double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
  k<NX><<<blocks, threads>>>(d, 1.0000001, 64, 1.0f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k<NX><<<blocks, threads>>>(d, 1.0000001, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double cyc = best * 1e-3 * clk_khz * 1e3 / ((double)iters * 16 * warps_per_sched);
  printf("%d warps/scheduler  pair + %d FADD  %.3f ms  %.2f cycles per pair per scheduler   [4+NX = %d, max(4, 2+NX) = %d]\n",
         warps_per_sched, NX, best, cyc, 4 + NX, NX > 2 ? 2 + NX : 4);
  cudaFree(d);
}

int main() {
  for (int w : {4, 8, 12}) {
    run<0>(w); run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w);
  }
  return 0;
}
