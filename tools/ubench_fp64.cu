// Micro-benchmarks that size the FP64 design on B200 (sm_100a): DMMA.8x8x4 issue
// rate, DFMA rate, and streaming-read bandwidth with 128-bit loads.  Development
// tool only; results are recorded in profiles/r01_ubench.md.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dmma(double *out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
  double a = a0 + threadIdx.x, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double *out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_stream(const double2 *__restrict__ in, size_t n2, double *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i + 7 * stride < n2; i += 8 * stride) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(in + i + u * stride));
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u].x + v[u].y;
  }
  for (; i < n2; i += stride) s += in[i].x + in[i].y;
  if (s == 123.456) out[0] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 32 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_dmma<16><<<148, warps * 32>>>(out, iters, 1.0, 1e-9);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      double fl = 2.0 * 256 * 16 * (double)iters * warps * 148;
      if (rep) printf("DMMA  warps/SM %2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
  }
  for (int warps : {8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_dfma<16><<<148, warps * 32>>>(out, iters, 1.0, 1e-9);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      double fl = 2.0 * 32 * 16 * (double)iters * warps * 148;
      if (rep) printf("DFMA  warps/SM %2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
  }
  size_t bytes = (size_t)8 << 30;
  double2 *buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
  for (int mult : {2, 4, 8, 16}) {
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      k_stream<<<148 * mult, 256>>>(buf, bytes / 16, out);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 2) printf("stream read 8 GiB grid 148x%-2d : %7.3f ms  %7.1f GB/s\n", mult, ms, bytes / ms * 1e-6);
    }
  }
  // L2-resident re-read (64 MiB)
  size_t small = (size_t)64 << 20;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    k_stream<<<148 * 8, 256>>>(buf, small / 16, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep == 3) printf("L2 re-read 64 MiB : %7.4f ms  %7.1f GB/s\n", ms, small / ms * 1e-6);
  }
  CK(cudaGetLastError());
  return 0;
}
