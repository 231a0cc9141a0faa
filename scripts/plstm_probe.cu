// Stand-alone probe of the persistent bi-LSTM layer kernel (csrc/lstm_persistent.cuh): times the layer on synthetic operands
// and prints clock64 stamps of the producer / MMA / epilogue roles of CTA (0,0,0) for steps 40..47 (VC_PLSTM_PROBE build).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -DVC_PLSTM_PROBE -I include -I video-captioning_b200/csrc \
//        scripts/plstm_probe.cu -o scripts/probe/plstm_probe -lcuda
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#ifndef PLSTM_HEADER
#define PLSTM_HEADER "lstm_persistent.cuh"
#endif
#include PLSTM_HEADER

namespace vc {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}
}  // namespace vc

__global__ void fill_kernel(vc::bf16* p, size_t n, float scale, unsigned seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)(i * 2654435761u) ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = __float2bfloat16_rn(((float)(x & 0xffff) / 32768.f - 1.f) * scale);
  }
}
__global__ void checksum_kernel(const vc::bf16* p, size_t n, unsigned long long* out) {
  unsigned long long s = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    s += (unsigned long long)(*reinterpret_cast<const unsigned short*>(p + i)) * (unsigned long long)(i % 1000003 + 1);
  atomicAdd(out, s);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1024, T = argc > 2 ? atoi(argv[2]) : 80, H = argc > 3 ? atoi(argv[3]) : 512;
  vc::bf16 *out, *xp, *W0, *W1;
  unsigned int* flags;
  long long* dbg;
  unsigned long long* cs;
  CK(cudaMalloc(&out, (size_t)B * T * 2 * H * 2));
  CK(cudaMalloc(&xp, (size_t)B * T * 8 * H * 2));
  CK(cudaMalloc(&W0, (size_t)4 * H * H * 2));
  CK(cudaMalloc(&W1, (size_t)4 * H * H * 2));
  CK(cudaMalloc(&flags, 4096));
  CK(cudaMalloc(&dbg, 8 * 32 * sizeof(long long)));
  CK(cudaMalloc(&cs, 8));
  fill_kernel<<<1024, 256>>>(xp, (size_t)B * T * 8 * H, 1.0f, 1u);
  fill_kernel<<<256, 256>>>(W0, (size_t)4 * H * H, 0.04f, 2u);
  fill_kernel<<<256, 256>>>(W1, (size_t)4 * H * H, 0.04f, 3u);
  CK(cudaMemset(dbg, 0, 8 * 32 * sizeof(long long)));
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 10;
  float ms = 0.f;
  const int pair_last = getenv("PROBE_PAIR") ? atoi(getenv("PROBE_PAIR")) : 1;
  const int cluster = getenv("PROBE_CLUSTER") ? atoi(getenv("PROBE_CLUSTER")) : 1;
  for (int mode = 0; mode < 3; ++mode) {
    // single CTA, single CTA in clusters of two (multicast h boxes), CTA pairs: the checksums must agree
    const int pair = mode == 2 ? 1 : 0, cl = mode == 1 ? 2 : 1;
    CK(cudaMemset(out, 0, (size_t)B * T * 2 * H * 2));
    for (int i = 0; i < 3; ++i)
      if (vc::tc::launch_lstm_layer_persistent(out, xp, W0, W1, B, T, H, flags, 0, nullptr, pair, cl) != 0) return 2;
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i)
      if (vc::tc::launch_lstm_layer_persistent(out, xp, W0, W1, B, T, H, flags, 0, nullptr, pair, cl) != 0) return 2;
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaMemset(cs, 0, 8));
    checksum_kernel<<<512, 256>>>(out, (size_t)B * T * 2 * H, cs);
    unsigned long long hcs = 0;
    CK(cudaMemcpy(&hcs, cs, 8, cudaMemcpyDeviceToHost));
    printf("%s B=%d T=%d H=%d: %.3f ms per layer (%.2f us per step), checksum %llx\n", pair ? "pair    " : (cl == 2 ? "cluster2" : "single  "), B, T, H, ms / reps,
           ms / reps * 1000.f / T, hcs);
  }
  if (vc::tc::launch_lstm_layer_persistent(out, xp, W0, W1, B, T, H, flags, 0, dbg, pair_last == 1 && cluster == 1 ? 1 : 0, cluster) != 0) return 2;
  CK(cudaDeviceSynchronize());
  long long h[8 * 32];
  CK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const char* names[9] = {"flag_seen", "loads_issued", "first_full", "mma_commit", "tmem_full", "cell_done", "store_read", "store_done", "red_issued"};
  const long long base = h[0] ? h[0] : h[8];
  printf("stamps of CTA (0,0,0) in ns since step 40's first event (nominal clock %d kHz); columns half0 half1\n", khz);
  for (int t = 0; t < 8; ++t) {
    printf("step %d:", 40 + t);
    for (int ev = 0; ev < 9; ++ev) {
      const double a = (double)(h[t * 32 + 2 * ev] - base) * 1e6 / khz, b = (double)(h[t * 32 + 2 * ev + 1] - base) * 1e6 / khz;
      printf(" %s %.0f/%.0f", names[ev], h[t * 32 + 2 * ev] ? a : -1.0, h[t * 32 + 2 * ev + 1] ? b : -1.0);
    }
    printf("\n");
  }
  return 0;
}
