// Micro-benchmark: MUFU.EX2 / FFMA / FADD issue throughput per SM on the device (lanes per clock).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f));
      if (MODE == 2) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(0.5f));
      if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f));
                       asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f));
                       asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f)); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int MODE> void run(const char* name, int warps, int per_iter) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  int iters = 4096;
  k<MODE><<<148, warps * 32>>>(d, iters); cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(d, iters); cudaDeviceSynchronize();
  float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
  double ops = (double)iters * 8 * per_iter * warps * 32;
  printf("%-22s warps/SM %2d: %.1f lanes/clk/SM\n", name, warps, ops / cyc);
  cudaFree(d);
}
int main() {
  for (int w : {4, 8, 16, 32}) { run<0>("MUFU.EX2", w, 1); }
  for (int w : {4, 8, 16, 32}) { run<1>("FFMA", w, 1); }
  for (int w : {4, 8, 16, 32}) { run<2>("FADD", w, 1); }
  for (int w : {8, 16}) { run<3>("EX2+4FFMA (ops=5)", w, 5); }
  return 0;
}
