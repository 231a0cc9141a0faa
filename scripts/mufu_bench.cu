// Microbenchmark: MUFU throughput on B200 (tanh.f32, tanh.f16x2, ex2.f32), warps/SM sweep.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i * 0.1f; h[i] = 0x3c003800u + threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 3) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo,hi}, %0; tanh.approx.f16 lo, lo; mov.b32 %0, {lo,hi};}" : "+r"(h[i]));
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char* name, int threads, int blocks_per_sm) {
  float* d; cudaMalloc(&d, 4);
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * blocks_per_sm, threads>>>(d, 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * blocks_per_sm, threads>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * blocks_per_sm * threads * (double)iters * 8 * (MODE == 1 ? 2 : 1);
  double clk = 1.965e9;
  printf("%-12s thr=%4d bps=%d  %.3f ms  %.1f results/clk/SM  (instr: %.2f warp-inst/clk/SM)\n", name, threads, blocks_per_sm, ms,
         ops / (ms * 1e-3) / clk / 148, 148.0 * blocks_per_sm * threads / 32 * (double)iters * 8 / (ms * 1e-3) / clk / 148);
  cudaFree(d);
}
int main() {
  for (int bps : {1, 2, 4}) {
    run<0>("tanh.f32", 256, bps);
    run<1>("tanh.f16x2", 256, bps);
    run<2>("ex2.f32", 256, bps);
    run<3>("tanh.f16", 256, bps);
  }
  return 0;
}
