// Stand-alone probe of the persistent additive attention kernel (csrc/attention.cuh, attn_additive_ws_kernel): times it on
// synthetic operands and prints clock64 stamps of CTA 0 per (video, 16-frame tile) unit: when the scoring group that owns the unit
// starts it, finishes its tanh loop, gets its partial-score slot, and when the context group waits for / gets the partials, gets the
// encoder tile and finishes the tile.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -DVC_ATTN_PROBE -I include -I video-captioning_b200/csrc \
//        scripts/attn_ws_probe.cu -o scripts/probe/attn_ws_probe
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "attention.cuh"

namespace vc {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}
}  // namespace vc

template <class T>
__global__ void fill16(T* p, size_t n, float scale, unsigned seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)(i * 2654435761u) ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = (T)(((float)(x & 0xffff) / 32768.f - 1.f) * scale);
  }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1024, T = 80, D = 512, H = 512, K = 5;
  __half *keys, *q, *v;
  vc::bf16 *vals, *ctx;
  long long* dbg;
  const size_t n = (size_t)B * T * D, R = (size_t)B * K;
  CK(cudaMalloc(&keys, n * 2));
  CK(cudaMalloc(&vals, (size_t)B * T * H * 2));
  CK(cudaMalloc(&q, R * D * 2));
  CK(cudaMalloc(&v, D * 2));
  CK(cudaMalloc(&ctx, R * H * 2));
  CK(cudaMalloc(&dbg, 64 * 16 * 8));
  CK(cudaMemset(dbg, 0, 64 * 16 * 8));
  fill16<<<1024, 256>>>(keys, n, 1.f, 1u);
  fill16<<<1024, 256>>>(vals, (size_t)B * T * H, 1.f, 2u);
  fill16<<<256, 256>>>(q, R * D, 1.f, 3u);
  fill16<<<1, 256>>>(v, (size_t)D, 0.05f, 4u);
  CK(cudaDeviceSynchronize());
  vc::AttnAddArgs a;
  memset(&a, 0, sizeof(a));
  a.keys = keys; a.q = q; a.v = v; a.values = vals; a.ctx = ctx; a.ctx_ld = H; a.B = B; a.T = T; a.D = D; a.H = H;
  for (int i = 0; i < 3; ++i)
    if (vc::launch_attn_additive_ws(a, K, 0) != 0) return 2;
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i)
    if (vc::launch_attn_additive_ws(a, K, 0) != 0) return 2;
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("B=%d T=%d D=%d H=%d K=%d: %.1f us per launch (%.2e tanh/s)\n", B, T, D, H, K, ms / reps * 1e3, (double)R * T * D / (ms / reps * 1e-3));
  {
    // the same launch carrying the decoder's row gather (attention.cuh: RowGather): (h, c) of two layers by parent row + the
    // embedding by token, into strided destinations like run_decode's; checked against the sources afterwards
    const int V = 10000, E = 512;
    const size_t ZW = E + 3 * H;
    vc::bf16 *hn[2], *emb, *Z, *XL;
    float *cn[2], *cd[2];
    int *parent, *tok;
    for (int l = 0; l < 2; ++l) {
      CK(cudaMalloc(&hn[l], R * H * 2)); CK(cudaMalloc(&cn[l], R * H * 4)); CK(cudaMalloc(&cd[l], R * H * 4));
      fill16<<<256, 256>>>(hn[l], R * H, 1.f, 10u + l);
      fill16<<<256, 256>>>(cn[l], R * H, 1.f, 20u + l);
    }
    CK(cudaMalloc(&emb, (size_t)V * E * 2)); CK(cudaMalloc(&Z, R * ZW * 2)); CK(cudaMalloc(&XL, R * 2 * H * 2));
    CK(cudaMemset(Z, 0, R * ZW * 2)); CK(cudaMemset(XL, 0, R * 2 * H * 2));
    fill16<<<256, 256>>>(emb, (size_t)V * E, 1.f, 30u);
    int* hp = (int*)malloc(R * 4 * 2);
    for (size_t r = 0; r < R; ++r) { hp[r] = (int)((r / K) * K + (r * 7 + 3) % K); hp[R + r] = (int)((r * 2654435761u) % V); }
    CK(cudaMalloc(&parent, R * 4)); CK(cudaMalloc(&tok, R * 4));
    CK(cudaMemcpy(parent, hp, R * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(tok, hp + R, R * 4, cudaMemcpyHostToDevice));
    vc::RowGather rg;
    memset(&rg, 0, sizeof(rg));
    bool ok = true;
    for (int l = 0; l < 2; ++l) {
      ok = ok && vc::row_gather_add(rg, hn[l], H * 2, l == 0 ? (void*)(Z + E + H) : (void*)(XL + H), l == 0 ? ZW * 2 : 2 * H * 2, H * 2, false);
      ok = ok && vc::row_gather_add(rg, cn[l], H * 4, cd[l], H * 4, H * 4, false);
    }
    ok = ok && vc::row_gather_add(rg, emb, E * 2, Z, ZW * 2, E * 2, true);
    if (!ok) { printf("gather: rows do not fit\n"); return 3; }
    rg.n_rows = (int)R; rg.V = V; rg.parent = parent; rg.tok = tok;
    if (getenv("GATHER_DBG")) rg.dbg = atoi(getenv("GATHER_DBG"));
    for (int i = 0; i < 3; ++i)
      if (vc::launch_attn_additive_ws(a, K, 0, rg) != 0) return 2;
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i)
      if (vc::launch_attn_additive_ws(a, K, 0, rg) != 0) return 2;
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    printf("  with the row gather (%d pieces of 512 bytes per row): %.1f us per launch\n", rg.n_pieces, ms / reps * 1e3);
    // check
    unsigned short* hz = (unsigned short*)malloc(R * ZW * 2);
    unsigned short* hh = (unsigned short*)malloc(R * H * 2);
    unsigned short* he = (unsigned short*)malloc((size_t)V * E * 2);
    float* hc = (float*)malloc(R * H * 4);
    float* hcd = (float*)malloc(R * H * 4);
    CK(cudaMemcpy(hz, Z, R * ZW * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hh, hn[0], R * H * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(he, emb, (size_t)V * E * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc, cn[1], R * H * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hcd, cd[1], R * H * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t r = 0; r < R; ++r)
      for (int c = 0; c < H; ++c) {
        bad += hz[r * ZW + E + H + c] != hh[(size_t)hp[r] * H + c];
        bad += hz[r * ZW + c] != he[(size_t)hp[R + r] * E + c];
        bad += hcd[r * H + c] != hc[(size_t)hp[r] * H + c];
      }
    printf("  gather check: %zu mismatches\n", bad);
    if (bad && !rg.dbg) return 4;
  }
  a.dbg = dbg;
  if (vc::launch_attn_additive_ws(a, K, 0) != 0) return 2;
  CK(cudaDeviceSynchronize());
  static long long h[64 * 16];
  CK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  long long base = 0;
  for (int u = 0; u < 64 && base == 0; ++u) base = h[u * 16];
  printf("unit: score_start loop_done slot_got | ctx_wait ctx_part ctx_enc ctx_done   (ns since the first stamp, nominal clock)\n");
  for (int u = 0; u < 40; ++u) {
    printf("%2d:", u);
    for (int ev = 0; ev < 8; ++ev) {
      if (ev == 3) { printf(" |"); continue; }
      printf(" %7.0f", h[u * 16 + ev] ? (double)(h[u * 16 + ev] - base) * 1e6 / khz : -1.0);
    }
    printf("\n");
  }
  return 0;
}
