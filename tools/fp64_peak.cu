// fp64_peak.cu -- measures the FP64 issue rate of the GPU for the instruction mix of the tap loops: bit-exactness
// forbids FMA, so every tap is one DMUL and one DADD.  Also measures how many non-FP64 instructions fit in the
// shadow of the FP64 pipe (the budget for loads, conversions and the epilogue).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fp64_peak fp64_peak.cu
// (An earlier version multiplied by a loop-invariant factor; the compiler hoisted the DMUL and the "pair" rate it
// printed was really the DADD rate counted twice.  Here the multiplicand changes every iteration.)
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: DMUL+DADD pair   1: DADD only   2: DMUL only   3: DFMA
// XI: extra integer instructions (LOP3/IADD chain, independent of the FP64 work) per FP64 pair
template <int ILP, int MODE, int XI>
__global__ void k(double *out, double w0, int iters, unsigned seed) {
  double a[ILP], v[ILP];
  unsigned z[4] = {seed + threadIdx.x, seed * 3u, seed * 5u, seed * 7u};
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 1e-9 + i; v[i] = 1.0 + i * 1e-6; }
  double w = w0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MODE == 0) a[i] = __dadd_rn(a[i], __dmul_rn(a[(i + ILP / 2) % ILP], w));   // every product has fresh operands
      if (MODE == 1) a[i] = __dadd_rn(a[i], v[i]);
      if (MODE == 2) a[i] = __dmul_rn(a[i], w);
      if (MODE == 3) a[i] = __fma_rn(v[i], w, a[i]);
#pragma unroll
      for (int x = 0; x < XI; x++) z[x & 3] = (z[x & 3] ^ (z[(x + 1) & 3] + 0x9e3779b9u)) + (z[x & 3] >> 3);
    }
    w = __longlong_as_double(__double_as_longlong(w) + 1);   // changes every iteration: the DMULs cannot be hoisted or shared
  }
  double s = 0;
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(z[0] ^ z[1] ^ z[2] ^ z[3]);
}

template <int ILP, int MODE, int XI>
void run(const char *name, int blocks, int threads) {
  double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
  const int iters = 4096;
  k<ILP, MODE, XI><<<blocks, threads>>>(d, 1.0000001, 64, 1);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k<ILP, MODE, XI><<<blocks, threads>>>(d, 1.0000001, iters, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double fp = (double)blocks * threads * iters * ILP * (MODE == 0 ? 2 : 1);
  printf("%-34s blocks=%4d thr=%3d  %.3f ms  %6.2f T fp64 lane-instr/s  (%.1f lanes/clk/SM at 1.965 GHz)\n", name, blocks, threads,
         best, fp / best / 1e9, fp / (best * 1e-3) / 148 / 1.965e9);
  cudaFree(d);
}

int main() {
  run<16, 1, 0>("dadd only ilp16", 148 * 8, 256);
  run<16, 2, 0>("dmul only ilp16", 148 * 8, 256);
  run<16, 3, 0>("dfma ilp16", 148 * 8, 256);
  run<8, 0, 0>("dmul+dadd ilp8", 148 * 8, 256);
  run<16, 0, 0>("dmul+dadd ilp16", 148 * 8, 256);
  run<24, 0, 0>("dmul+dadd ilp24 4x128", 148 * 4, 128);
  run<24, 0, 0>("dmul+dadd ilp24 1 warp/SMSP", 148, 128);
  run<24, 0, 0>("dmul+dadd ilp24 2 warps/SMSP", 148, 256);
  run<16, 0, 1>("pair + 2 int instr", 148 * 8, 256);
  run<16, 0, 2>("pair + 4 int instr", 148 * 8, 256);
  run<16, 0, 3>("pair + 6 int instr", 148 * 8, 256);
  run<16, 0, 4>("pair + 8 int instr", 148 * 8, 256);
  return 0;
}
