// Host-side feature packing for the ingest path (predictor.py:101-107: np float32 features -> device).
// The per-video input is 1.3 MB of fp32 and the PCIe link is the end-to-end bottleneck; bf16 mode rounds the
// features to bf16 on the device anyway, so part of the batch is rounded on the host cores instead (same
// round-to-nearest-even as __float2bfloat16_rn) and crosses the link at half the size.  Plain C++ (no CUDA):
// compiled by the host compiler, AVX-512 path selected at run time.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

namespace {

inline uint16_t pack_one(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40u);   // NaN stays NaN
  return (uint16_t)((x + 0x7fffu + ((x >> 16) & 1u)) >> 16);
}

void pack_scalar(const float* s, uint16_t* d, size_t n) {
  for (size_t i = 0; i < n; ++i) d[i] = pack_one(s[i]);
}

__attribute__((target("avx512f,avx512bw"))) inline __m256i pack16(__m512i x) {
  const __m512i one = _mm512_set1_epi32(1), bias = _mm512_set1_epi32(0x7fff);
  const __m512i lsb = _mm512_and_si512(_mm512_srli_epi32(x, 16), one);
  const __m512i r = _mm512_srli_epi32(_mm512_add_epi32(_mm512_add_epi32(x, bias), lsb), 16);
  const __mmask16 nan = _mm512_cmpgt_epu32_mask(_mm512_and_si512(x, _mm512_set1_epi32(0x7fffffff)), _mm512_set1_epi32(0x7f800000));
  const __m512i q = _mm512_or_si512(_mm512_srli_epi32(x, 16), _mm512_set1_epi32(0x40));
  return _mm512_cvtepi32_epi16(_mm512_mask_blend_epi32(nan, r, q));
}

__attribute__((target("avx512f,avx512bw"))) void pack_avx512(const float* s, uint16_t* d, size_t n) {
  size_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 31u) != 0) { d[i] = pack_one(s[i]); ++i; }
  for (; i + 32 <= n; i += 32) {
    const __m512i a = _mm512_loadu_si512(reinterpret_cast<const void*>(s + i));
    const __m512i b = _mm512_loadu_si512(reinterpret_cast<const void*>(s + i + 16));
    // non-temporal stores: the packed block is read next by the DMA engine, not by this core
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), pack16(a));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 16), pack16(b));
  }
  _mm_sfence();
  for (; i < n; ++i) d[i] = pack_one(s[i]);
}

}  // namespace

// dst[i] = bf16(src[i]) (round to nearest even), i < n, on `threads` host threads.  Returns 0.
extern "C" int vc_host_pack_bf16(const float* src, uint16_t* dst, size_t n, int32_t threads) {
  if (src == nullptr || dst == nullptr) return 1;
  const bool wide = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
  auto run = [=](size_t lo, size_t hi) {
    if (hi <= lo) return;
    if (wide) pack_avx512(src + lo, dst + lo, hi - lo);
    else pack_scalar(src + lo, dst + lo, hi - lo);
  };
  if (threads < 1) threads = 1;
  if (threads == 1 || n < (size_t)1 << 16) {
    run(0, n);
    return 0;
  }
  const size_t per = ((n + threads - 1) / threads + 63) & ~(size_t)63;
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t) {
    const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
    if (lo >= n) break;
    pool.emplace_back(run, lo, hi);
  }
  for (auto& th : pool) th.join();
  return 0;
}
