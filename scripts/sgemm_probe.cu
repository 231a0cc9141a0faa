// Stand-alone timing probe of the fp32 GEMMs (csrc/gemm_f32.cuh) at the per-step shapes of the fp32 mode.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I include -I video-captioning_b200/csrc scripts/sgemm_probe.cu -o scripts/probe/sgemm_probe
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "gemm_f32.cuh"

namespace vc {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}
}  // namespace vc
using namespace vc;

__global__ void fill_f32(float* p, size_t n, float scale, unsigned seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)(i * 2654435761u) ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = ((float)(x & 0xffff) / 32768.f - 1.f) * scale;
  }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
  const int shapes[4][4] = {{32, 2048, 512, 2}, {32, 512, 1536, 1}, {32, 10000, 512, 1}, {32, 2048, 1536, 1}};
  float *A, *W, *C, *C2;
  CK(cudaMalloc(&A, (size_t)64 * 4096 * 4));
  CK(cudaMalloc(&W, (size_t)2 * 10000 * 2048 * 4));
  CK(cudaMalloc(&C, (size_t)2 * 64 * 10000 * 4));
  CK(cudaMalloc(&C2, (size_t)2 * 64 * 10000 * 4));
  fill_f32<<<256, 256>>>(A, (size_t)64 * 4096, 1.f, 1u);
  fill_f32<<<1024, 256>>>(W, (size_t)2 * 10000 * 2048, 0.05f, 2u);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int si = 0; si < 4; ++si) {
    const int M = shapes[si][0], N = shapes[si][1], K = shapes[si][2], nz = shapes[si][3];
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A[0] = A; g.A[1] = A + 2048; g.W[0] = W; g.W[1] = W + (size_t)N * K;
    g.lda = 4096; g.ldw = K; g.M = M; g.N = N; g.K = K; g.nz = nz; g.a_split = 1 << 30;
    for (int variant = 0; variant < 2; ++variant) {
      EpiStore<float, false, true> e;
      memset(&e, 0, sizeof(e));
      float* out = variant ? C2 : C;
      e.C[0] = out; e.C[1] = out + (size_t)64 * 10000; e.ldc = N;
      auto run = [&]() -> int {
        if (variant == 0) return launch_sgemm(g, e, 0);
        dim3 grid((N + 127) / 128, 1, nz);
        sgemm_nt_kernel<32, 4, EpiStore<float, false, true>><<<grid, 128>>>(g, e);
        return 0;
      };
      for (int i = 0; i < 3; ++i) run();
      CK(cudaDeviceSynchronize());
      const int reps = 50;
      cudaEventRecord(e0);
      for (int i = 0; i < reps; ++i) run();
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("M=%d N=%d K=%d nz=%d %s: %.2f us per launch (%.1f GFLOP/s)\n", M, N, K, nz, variant ? "sgemm_nt<32,4>" : "skinny", ms / reps * 1e3,
             2.0 * M * N * K * nz / (ms / reps * 1e-3) / 1e9);
    }
    // bit-identity of the two kernels
    const size_t n = (size_t)M * N;
    float* h1 = (float*)malloc(n * 4);
    float* h2 = (float*)malloc(n * 4);
    CK(cudaMemcpy(h1, C, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2, C2, n * 4, cudaMemcpyDeviceToHost));
    printf("  identical: %s\n", memcmp(h1, h2, n * 4) == 0 ? "yes" : "NO");
    free(h1);
    free(h2);
  }
  return 0;
}
