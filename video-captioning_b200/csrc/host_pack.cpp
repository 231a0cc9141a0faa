// Host-side feature packing for the ingest path (predictor.py:101-107: np float32 features -> device).
// The per-video input is 1.3 MB of fp32 and the PCIe link is the end-to-end bottleneck; bf16 mode rounds the
// features to bf16 on the device anyway, so part of the batch is rounded on the host cores instead (same
// round-to-nearest-even as __float2bfloat16_rn) and crosses the link at half the size.  Plain C++ (no CUDA):
// compiled by the host compiler, AVX-512 path selected at run time.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
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

// Persistent worker pool: a piece of the ingest is ~1.5 ms of work, creating 16 threads per call would cost a fifth of it.
// One job at a time (callers are serialised by `submit_mu`); workers sleep on a condition variable between jobs.
class Pool {
 public:
  static Pool& get() {
    static Pool* p = new Pool();   // leaked on purpose: worker threads must not be joined during static destruction
    return *p;
  }
  // run fn(t) for t in [0, parts) on the workers (+ the caller for part 0); returns when all parts are done
  void run(int parts, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> submit(submit_mu_);
    if (parts <= 1) { fn(0); return; }
    grow(parts - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      parts_ = parts;
      next_ = 1;
      pending_ = parts - 1;
      ++gen_;
    }
    cv_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void grow(int n) {
    while ((int)workers_.size() < n) {
      workers_.emplace_back([this] { loop(); });
      workers_.back().detach();
    }
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(int)>* fn = nullptr;
      int part = -1;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen && fn_ != nullptr && next_ < parts_; });
        part = next_++;
        if (next_ >= parts_) seen = gen_;     // the last part of this job: wait for the next generation afterwards
        fn = fn_;
      }
      (*fn)(part);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::mutex submit_mu_, mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* fn_ = nullptr;
  int parts_ = 0, next_ = 0, pending_ = 0;
  unsigned long gen_ = 0;
};

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
  const int parts = (int)((n + per - 1) / per);
  Pool::get().run(parts, [&](int t) {
    const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
    run(lo, hi);
  });
  return 0;
}

// ---------------------------------------------------------------- batched resize / pad gather into a staging buffer
// predictor.py:101-107 + :292-315 + data/dataset.py:124-150 for a whole batch: row r of the [n_rows, F] staging buffer
// (n_rows = videos x target frames) is a copy of the source frame src_rows[r] points at (the frame the reference's
// linspace subsampling selects), or zeros where src_rows[r] == 0 (zero padding of a short video).  The copy converts on
// the fly: fp32 -> fp32 (plain), fp32 -> bf16 (round to nearest even: the packed ingest, half the bytes over the link),
// fp16 -> fp16 (features stored as halves are staged without widening).  One pass over the selected source frames only --
// frames the subsampling drops are never touched.  dtype codes: 0 = fp32, 1 = bf16, 2 = fp16.
namespace {
inline float half_to_float(uint16_t h) {
  const uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  uint32_t x;
  if (e == 0) {
    if (m == 0) x = s;
    else {       // subnormal
      int sh = 0;
      uint32_t mm = m;
      while ((mm & 0x400u) == 0) { mm <<= 1; ++sh; }
      x = s | ((uint32_t)(113 - sh) << 23) | ((mm & 0x3ffu) << 13);
    }
  } else if (e == 31) x = s | 0x7f800000u | (m << 13);
  else x = s | ((e + 112) << 23) | (m << 13);
  float f;
  memcpy(&f, &x, 4);
  return f;
}
}  // namespace

extern "C" int vc_host_stage_rows(const uint64_t* src_rows, int64_t n_rows, int64_t F, int32_t src_dtype, void* dst,
                                  int32_t dst_dtype, int32_t threads) {
  if (src_rows == nullptr || dst == nullptr || n_rows < 0 || F <= 0) return 1;
  const bool f32_f32 = src_dtype == 0 && dst_dtype == 0, f32_b16 = src_dtype == 0 && dst_dtype == 1;
  const bool f16_f16 = src_dtype == 2 && dst_dtype == 2, f16_b16 = src_dtype == 2 && dst_dtype == 1, f16_f32 = src_dtype == 2 && dst_dtype == 0;
  if (!(f32_f32 || f32_b16 || f16_f16 || f16_b16 || f16_f32)) return 1;
  const size_t dst_es = dst_dtype == 0 ? 4 : 2;
  const bool wide = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
  auto run = [=](int64_t lo, int64_t hi) {
    for (int64_t r = lo; r < hi; ++r) {
      uint8_t* d = reinterpret_cast<uint8_t*>(dst) + (size_t)r * F * dst_es;
      const void* s = reinterpret_cast<const void*>(static_cast<uintptr_t>(src_rows[r]));
      if (s == nullptr) { memset(d, 0, (size_t)F * dst_es); continue; }
      if (f32_f32) memcpy(d, s, (size_t)F * 4);
      else if (f16_f16) memcpy(d, s, (size_t)F * 2);
      else if (f32_b16) {
        if (wide) pack_avx512(reinterpret_cast<const float*>(s), reinterpret_cast<uint16_t*>(d), (size_t)F);
        else pack_scalar(reinterpret_cast<const float*>(s), reinterpret_cast<uint16_t*>(d), (size_t)F);
      } else {
        const uint16_t* hs = reinterpret_cast<const uint16_t*>(s);
        if (f16_f32) { float* o = reinterpret_cast<float*>(d); for (int64_t i = 0; i < F; ++i) o[i] = half_to_float(hs[i]); }
        else { uint16_t* o = reinterpret_cast<uint16_t*>(d); for (int64_t i = 0; i < F; ++i) o[i] = pack_one(half_to_float(hs[i])); }
      }
    }
  };
  if (threads < 1) threads = 1;
  if (threads == 1 || n_rows * F < ((int64_t)1 << 16)) {
    run(0, n_rows);
    return 0;
  }
  const int64_t per = (n_rows + threads - 1) / threads;
  const int parts = (int)((n_rows + per - 1) / per);
  Pool::get().run(parts, [&](int t) {
    const int64_t lo = (int64_t)t * per, hi = lo + per < n_rows ? lo + per : n_rows;
    run(lo, hi);
  });
  return 0;
}
