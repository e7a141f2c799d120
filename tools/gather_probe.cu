// gather_probe.cu -- how fast can an SM look random codes up in a 65 536-entry table of doubles (512 KB: the
// code -> linear table of 16-bit input, reference: src/imagew-main.c:348-372 evaluates pow() per sample there)?
// Measures the alternatives for the 16-bit vertical kernel: global memory through L1/L2 (today), the part of the table
// that fits in shared memory, and the whole table spread over the shared memory of a 2- or 4-CTA cluster (DSMEM).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned rnd(unsigned &s) { s = s * 1664525u + 1013904223u; return s >> 16; }

// MODE 0: global (read-only path)   1: shared, indices folded into the local part   2: cluster DSMEM, table split over CL CTAs
//      3: hybrid: codes < local part from shared memory, the rest from global
template <int MODE, int ILP, typename T>
__global__ void k(const T *__restrict__ tab, T *out, int iters, int local_entries, int cl) {
  extern __shared__ __align__(16) unsigned char smraw[];
  T *sm = (T *)smraw;
  unsigned rank = 0;
  if (MODE == 2) rank = cg::this_cluster().block_rank();
  if (MODE != 0) {
    for (int i = threadIdx.x; i < local_entries; i += blockDim.x) sm[i] = tab[(MODE == 2 ? rank * local_entries : 0) + i];
    __syncthreads();
    if (MODE == 2) cg::this_cluster().sync();
  }
  unsigned s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1;
  T acc[ILP];
  for (int i = 0; i < ILP; i++) acc[i] = 0;
  const unsigned smbase = (unsigned)__cvta_generic_to_shared(sm);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      const unsigned c = rnd(s);          // 16 bits
      T v;
      if (MODE == 0) v = __ldg(tab + c);
      if (MODE == 1) v = sm[c % (unsigned)local_entries];
      if (MODE == 2) {
        const unsigned cc = c % (unsigned)(cl * local_entries);   // (a cluster that holds part of the table: fold)
        const unsigned r = cc / (unsigned)local_entries, o = cc % (unsigned)local_entries;
        unsigned ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smbase + o * (unsigned)sizeof(T)), "r"(r));
        if (sizeof(T) == 8) { unsigned long long u; asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(u) : "r"(ra)); v = (T)__longlong_as_double((long long)u); }
        else { unsigned u; asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(u) : "r"(ra)); v = (T)__uint_as_float(u); }
      }
      if (MODE == 3) v = c < (unsigned)local_entries ? sm[c] : __ldg(tab + c);
      acc[i] += v;
    }
  }
  T t = 0;
  for (int i = 0; i < ILP; i++) t += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
  if (MODE == 2) cg::this_cluster().sync();
}

template <int MODE, int ILP, typename T>
void run(const char *name, const T *tab, T *out, int blocks, int threads, int local_entries, int cl, int carve) {
  auto kern = k<MODE, ILP, T>;
  const size_t smem = MODE == 0 ? 0 : (size_t)local_entries * sizeof(T);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (carve >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = cl > 1 ? 1 : 0;
  const int iters = 2048;
  cudaLaunchKernelEx(&cfg, kern, tab, out, 16, local_entries, cl);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, kern, tab, out, iters, local_entries, cl);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)blocks * threads * iters * ILP;
  printf("%-58s blocks=%4d thr=%4d  %8.3f ms  %7.1f G lookups/s  (%5.2f per clk per SM at 1.965 GHz)  %s\n", name, blocks, threads, ms,
         n / ms * 1e-6, n / (ms * 1e-3) / 148 / 1.965e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  double *tab; cudaMalloc(&tab, 65536 * 8);
  double *h = (double *)malloc(65536 * 8);
  for (int i = 0; i < 65536; i++) h[i] = i * 1e-5;
  cudaMemcpy(tab, h, 65536 * 8, cudaMemcpyHostToDevice);
  float *tabf = (float *)tab;
  double *out; cudaMalloc(&out, 8u << 20);
  float *outf = (float *)out;
  run<0, 4, double>("global 8 B, 16 warps/SM (today's occupancy), ILP 4", tab, out, 148 * 4, 128, 0, 1, -1);
  run<0, 4, double>("global 8 B, 32 warps/SM, ILP 4", tab, out, 148 * 8, 128, 0, 1, -1);
  run<0, 4, double>("global 8 B, 64 warps/SM, ILP 4, max L1", tab, out, 148 * 8, 256, 0, 1, 0);
  run<0, 8, double>("global 8 B, 64 warps/SM, ILP 8, max L1", tab, out, 148 * 8, 256, 0, 1, 0);
  run<0, 4, float>("global 4 B (256 KB table), 64 warps/SM, ILP 4, max L1", tabf, outf, 148 * 8, 256, 0, 1, 0);
  run<1, 4, double>("shared 8 B (24 576 entries local), 16 warps, ILP 4", tab, out, 148, 512, 24576, 1, -1);
  run<1, 8, double>("shared 8 B (24 576 entries local), 16 warps, ILP 8", tab, out, 148, 512, 24576, 1, -1);
  run<1, 4, double>("shared 8 B (24 576 entries local), 32 warps, ILP 4", tab, out, 148, 1024, 24576, 1, -1);
  run<1, 4, float>("shared 4 B (49 152 entries local), 16 warps, ILP 4", tabf, outf, 148, 512, 49152, 1, -1);
  run<3, 4, double>("hybrid 8 B: 24 576 entries shared + rest global, 16 warps", tab, out, 148, 512, 24576, 1, -1);
  run<3, 4, double>("hybrid 8 B: 24 576 entries shared + rest global, 32 warps", tab, out, 148, 1024, 24576, 1, -1);
  run<3, 4, double>("hybrid 8 B: 20 480 entries shared + rest global, 32 warps", tab, out, 148, 1024, 20480, 1, -1);
  run<2, 4, double>("cluster of 4 x 16 384 entries 8 B (DSMEM), 16 warps, ILP 4", tab, out, 148, 512, 16384, 4, -1);
  run<2, 8, double>("cluster of 4 x 16 384 entries 8 B (DSMEM), 16 warps, ILP 8", tab, out, 148, 512, 16384, 4, -1);
  run<2, 4, double>("cluster of 4 x 16 384 entries 8 B (DSMEM), 32 warps, ILP 4", tab, out, 148, 1024, 16384, 4, -1);
  run<2, 4, double>("cluster of 2 x 24 576 entries 8 B (DSMEM, indices folded)", tab, out, 148, 512, 24576, 2, -1);
  run<2, 4, float>("cluster of 2 x 32 768 entries 4 B (DSMEM), 16 warps, ILP 4", tabf, outf, 148, 512, 32768, 2, -1);
  run<2, 8, float>("cluster of 2 x 32 768 entries 4 B (DSMEM), 32 warps, ILP 8", tabf, outf, 148, 1024, 32768, 2, -1);
  run<2, 4, float>("cluster of 4 x 16 384 entries 4 B (DSMEM), 16 warps, ILP 4", tabf, outf, 148, 512, 16384, 4, -1);
  return 0;
}
