// Stand-alone timing probe of the streaming dot-product attention kernel (csrc/attention_dot.cuh).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a [-DVC_DOT_PROBE=1|2] -I include -I video-captioning_b200/csrc \
//        scripts/attn_dot_probe.cu -o scripts/probe/attn_dot_probe
// VC_DOT_PROBE=1: consumers skip the arithmetic (producer / HBM side alone); =2: the producer skips the loads (compute alone).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "attention_dot.cuh"

namespace vc {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}
}  // namespace vc

__global__ void fill_bf16(vc::bf16* p, size_t n, float scale, unsigned seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)(i * 2654435761u) ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = __float2bfloat16_rn(((float)(x & 0xffff) / 32768.f - 1.f) * scale);
  }
}
__global__ void fill_f32(float* p, size_t n, float scale, unsigned seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)(i * 2654435761u) ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = ((float)(x & 0xffff) / 32768.f - 1.f) * scale;
  }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1024, T = argc > 2 ? atoi(argv[2]) : 80, H = argc > 3 ? atoi(argv[3]) : 1024;
  const int K = argc > 4 ? atoi(argv[4]) : 5, heads = argc > 5 ? atoi(argv[5]) : 1, fp32q = argc > 6 ? atoi(argv[6]) : 0;
  vc::bf16 *keys, *vals, *qa, *ctx;
  float* q;
  const size_t n = (size_t)B * T * H, R = (size_t)B * K;
  CK(cudaMalloc(&keys, n * 2));
  CK(cudaMalloc(&vals, n * 2));
  CK(cudaMalloc(&qa, R * H * 2));
  CK(cudaMalloc(&q, R * H * 4));
  CK(cudaMalloc(&ctx, R * H * 2));
  fill_bf16<<<1024, 256>>>(keys, n, 1.f, 1u);
  fill_bf16<<<1024, 256>>>(vals, n, 1.f, 2u);
  fill_bf16<<<256, 256>>>(qa, R * H, 0.2f, 3u);
  fill_f32<<<256, 256>>>(q, R * H, 0.2f, 4u);
  CK(cudaDeviceSynchronize());
  vc::AttnDotArgs a;
  memset(&a, 0, sizeof(a));
  a.skeys = keys; a.values = heads > 1 ? vals : keys;
  a.q_act = qa; a.q_ld = H;
  a.ctx = ctx; a.ctx_ld = H; a.B = B; a.K = K; a.T = T; a.H = H; a.heads = heads; a.scale = 1.f;
  for (int i = 0; i < 3; ++i)
    if (vc::launch_attn_dot_ws(a, 0) != 0) return 2;
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i)
    if (vc::launch_attn_dot_ws(a, 0) != 0) return 2;
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)n * 2 * (heads > 1 ? 2 : 1);
  printf("B=%d T=%d H=%d K=%d heads=%d fp32q=%d: %.1f us per launch, %.0f GB/s of tile reads\n", B, T, H, K, heads, fp32q, ms / reps * 1e3,
         bytes / (ms / reps * 1e-3) / 1e9);
  return 0;
}
