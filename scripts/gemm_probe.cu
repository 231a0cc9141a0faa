// Stand-alone probe of the persistent CTA-pair LSTM GEMM (csrc/gemm_tc.cuh) at the decoder's shapes: time per launch and clock64
// stamps of CTA 0 per tile (producer starts the tile's loads, MMA issuer gets the accumulator / the first k-block / commits the
// last MMA, epilogue gets the accumulator / finishes the cell math / has its stores read).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -DVC_GEMM_PROBE -I include -I video-captioning_b200/csrc \
//        scripts/gemm_probe.cu -o scripts/probe/gemm_probe -lcuda
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "gemm_tc.cuh"

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

__global__ void fill_bf16(bf16* p, size_t n, float scale, unsigned seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)(i * 2654435761u) ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = __float2bfloat16_rn(((float)(x & 0xffff) / 32768.f - 1.f) * scale);
  }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// vocabulary projection with the selection statistics: gemm_probe vocab [M] [V] [K]
int vocab_main(int M, int V, int K) {
  bf16 *A, *W;
  float *logits, *bias, *cmax;
  float2* part;
  int* rowthr;
  long long* dbg;
  const int tn = (V + 255) / 256;
  CK(cudaMalloc(&A, (size_t)M * K * 2));
  CK(cudaMalloc(&W, (size_t)V * K * 2));
  CK(cudaMalloc(&logits, (size_t)M * V * 4));
  CK(cudaMalloc(&bias, (size_t)V * 4));
  CK(cudaMalloc(&cmax, (size_t)M * 8 * tn * 4));
  CK(cudaMalloc(&part, (size_t)M * 2 * tn * 8));
  CK(cudaMalloc(&rowthr, (size_t)M * 4));
  CK(cudaMalloc(&dbg, 16 * 8 * 8));
  CK(cudaMemset(dbg, 0, 16 * 8 * 8));
  CK(cudaMemset(bias, 0, (size_t)V * 4));
  fill_bf16<<<1024, 256>>>(A, (size_t)M * K, 1.f, 1u);
  fill_bf16<<<1024, 256>>>(W, (size_t)V * K, 0.05f, 2u);
  CK(cudaDeviceSynchronize());
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A[0] = g.A[1] = A; g.W[0] = g.W[1] = W; g.lda = K; g.ldw = K; g.M = M; g.N = V; g.K = K; g.nz = 1; g.a_split = 1 << 30;
  EpiStore<float, false, false> e;
  memset(&e, 0, sizeof(e));
  e.C[0] = e.C[1] = logits; e.ldc = V; e.bias[0] = e.bias[1] = bias;
  tc::VocabStats vs;
  memset(&vs, 0, sizeof(vs));
  vs.cmax = cmax; vs.part = part; vs.nc = 8 * tn; vs.np = 2 * tn; vs.rowthr = rowthr; vs.topk = 5;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int modes[6] = {0, 1, 2, 4, 3, 7};    // VocabStats::dbg bits: 1 no logits stores, 2 no exp sums, 4 no statistics stores
  for (int mi = 0; mi < 6; ++mi) {
    const int dbgmode = modes[mi];
    vs.dbg = dbgmode;
    const int reps = 20;
    float total = 0.f;
    for (int i = 0; i < reps + 3; ++i) {
      CK(cudaMemsetAsync(rowthr, 0x80, (size_t)M * 4, 0));
      if (i >= 3) cudaEventRecord(e0);
      if (tc::launch_gemm_tc(g, K, e, 0, &vs) != 0) return 2;
      if (i >= 3) {
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        total += ms;
      }
    }
    printf("vocab GEMM M=%d V=%d K=%d epilogue parts skipped (dbg bits) %d: %.1f us per launch (%.0f TFLOP/s)\n", M, V, K, dbgmode,
           total / reps * 1e3, 2.0 * M * V * K / (total / reps * 1e-3) / 1e12);
  }
  vs.dbg = 0;
  CK(cudaMemsetAsync(rowthr, 0x80, (size_t)M * 4, 0));
  tc::probe_dbg() = dbg;
  if (tc::launch_gemm_tc(g, K, e, 0, &vs) != 0) return 2;
  CK(cudaDeviceSynchronize());
  static long long h[16 * 8];
  CK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  long long base = 0;
  for (int i = 0; i < 128; ++i) if (h[i] && (base == 0 || h[i] < base)) base = h[i];
  printf("tile: loads_start acc_free first_kb mma_done | epi_acc epi_done   (ns since the first stamp of CTA 0)\n");
  for (int it = 0; it < 12; ++it) {
    printf("%2d:", it);
    for (int ev = 0; ev < 6; ++ev) {
      if (ev == 4) printf(" |");
      printf(" %7.0f", h[it * 8 + ev] ? (double)(h[it * 8 + ev] - base) * 1e6 / khz : -1.0);
    }
    printf("\n");
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && strcmp(argv[1], "vocab") == 0)
    return vocab_main(argc > 2 ? atoi(argv[2]) : 5120, argc > 3 ? atoi(argv[3]) : 10000, argc > 4 ? atoi(argv[4]) : 512);
  const int M = argc > 1 ? atoi(argv[1]) : 5120, H = argc > 2 ? atoi(argv[2]) : 512, K = argc > 3 ? atoi(argv[3]) : 1536;
  const int N = 4 * H;
  bf16 *A, *W, *h0, *h1;
  float *c, *cn, *bias;
  long long* dbg;
  CK(cudaMalloc(&A, (size_t)M * K * 2));
  CK(cudaMalloc(&W, (size_t)N * K * 2));
  CK(cudaMalloc(&h0, (size_t)M * H * 2));
  CK(cudaMalloc(&h1, (size_t)M * 2 * H * 2));
  CK(cudaMalloc(&c, (size_t)M * H * 4));
  CK(cudaMalloc(&cn, (size_t)M * H * 4));
  CK(cudaMalloc(&bias, (size_t)N * 4));
  CK(cudaMalloc(&dbg, 16 * 8 * 8));
  CK(cudaMemset(dbg, 0, 16 * 8 * 8));
  CK(cudaMemset(c, 0, (size_t)M * H * 4));
  CK(cudaMemset(bias, 0, (size_t)N * 4));
  fill_bf16<<<1024, 256>>>(A, (size_t)M * K, 1.f, 1u);
  fill_bf16<<<1024, 256>>>(W, (size_t)N * K, 0.03f, 2u);
  CK(cudaDeviceSynchronize());
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A[0] = g.A[1] = A; g.W[0] = g.W[1] = W; g.lda = K; g.ldw = K; g.M = M; g.N = N; g.K = K; g.nz = 1; g.a_split = 1 << 30;
  EpiLstm<bf16, bf16, false> e;
  memset(&e, 0, sizeof(e));
  e.bias[0] = e.bias[1] = bias;
  e.c_prev[0] = e.c_prev[1] = c; e.c_new[0] = e.c_new[1] = cn; e.c_ld = H;
  e.h_out0[0] = e.h_out0[1] = h0; e.h0_ld = H;
  e.h_out1[0] = e.h_out1[1] = h1; e.h1_ld = 2 * H; e.h1_origin = h1; e.h1_origin_cols = 2 * H;
  e.c_origin_in = c; e.c_origin_out = cn; e.c_tma_cols = H;
  for (int i = 0; i < 3; ++i)
    if (tc::launch_gemm_tc(g, K, e, 0) != 0) return 2;
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i)
    if (tc::launch_gemm_tc(g, K, e, 0) != 0) return 2;
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("LSTM GEMM M=%d N=%d K=%d: %.1f us per launch (%.0f TFLOP/s), back to back\n", M, N, K, ms / reps * 1e3, 2.0 * M * N * K / (ms / reps * 1e-3) / 1e12);
  tc::probe_dbg() = dbg;
  if (tc::launch_gemm_tc(g, K, e, 0) != 0) return 2;
  CK(cudaDeviceSynchronize());
  static long long h[16 * 8];
  CK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  long long base = 0;
  for (int i = 0; i < 128 && base == 0; ++i) base = h[i];
  for (int i = 0; i < 128; ++i) if (h[i] && h[i] < base) base = h[i];
  printf("tile: loads_start acc_free first_kb mma_done | epi_acc epi_math epi_stored   (ns since the first stamp of CTA 0)\n");
  for (int it = 0; it < 4; ++it) {
    printf("%d:", it);
    for (int ev = 0; ev < 7; ++ev) {
      if (ev == 4) printf(" |");
      printf(" %7.0f", h[it * 8 + ev] ? (double)(h[it * 8 + ev] - base) * 1e6 / khz : -1.0);
    }
    printf("\n");
  }
  return 0;
}
