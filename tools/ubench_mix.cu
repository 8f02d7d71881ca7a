// DMMA + DFMA co-issue probe (development tool): do the two FP64 paths share one datapath?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NM, int NF>
__global__ void __launch_bounds__(512) k_mix(double *out, int iters, double a0, double b0) {
  double c[NM > 0 ? NM : 1][2]; double f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < NM; ++i) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (NM > NF ? NM : NF); ++i) {
      if (i < NM) dmma(c[i][0], c[i][1], a, b);
      if (i < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NM, int NF> void run(double *out, const char *name) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000, warps = 16;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    k_mix<NM, NF><<<148, warps * 32>>>(out, iters, 1.0, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fm = 2.0 * 256 * NM * (double)iters * warps * 148, ff = 2.0 * 32 * NF * (double)iters * warps * 148;
    if (rep) printf("%-22s %8.3f ms  dmma %6.2f + dfma %6.2f = %6.2f TFLOP/s\n", name, ms, fm / ms * 1e-9, ff / ms * 1e-9, (fm + ff) / ms * 1e-9);
  }
}
int main() {
  double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 1024));
  run<16, 0>(out, "dmma16");
  run<0, 16>(out, "dfma16");
  run<16, 16>(out, "dmma16+dfma16");
  run<8, 16>(out, "dmma8+dfma16");
  run<4, 16>(out, "dmma4+dfma16");
  run<2, 16>(out, "dmma2+dfma16");
  run<16, 4>(out, "dmma16+dfma4");
  CK(cudaGetLastError());
  return 0;
}
