// Common device/host helpers for the B200 caption-generation kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vc_b200.h"

namespace vc {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- error plumbing (no exceptions cross the C ABI)
// Status codes are the public VC_* enum of include/vc_b200.h.
void set_error(const char* fmt, ...);   // defined in capi.cu (thread-local message)

#define VC_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      vc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VC_ERR_CUDA;                                                          \
    }                                                                                  \
  } while (0)

#define VC_CHECK(cond, ...)                                                            \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      vc::set_error(__VA_ARGS__);                                                      \
      return VC_ERR_INVALID;                                                       \
    }                                                                                  \
  } while (0)

#define VC_TRY(expr)                 \
  do {                               \
    int _s = (expr);                 \
    if (_s != VC_OK) return _s;  \
  } while (0)

// ---------------------------------------------------------------- type helpers
__device__ __forceinline__ float to_float(float x) { return x; }
__device__ __forceinline__ float to_float(bf16 x) { return __bfloat162float(x); }
template <class T> __device__ __forceinline__ T from_float(float x);
template <> __device__ __forceinline__ float from_float<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_float<bf16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_float<__half>(float x) { return __float2half_rn(x); }
__device__ __forceinline__ float to_float(__half x) { return __half2float(x); }

// Load 4 consecutive elements as floats (pointer 16B- (float) / 8B- (bf16) aligned).
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
__device__ __forceinline__ void store4(__half* p, const float (&v)[4]) {
  __half2 a = __floats2half2_rn(v[0], v[1]);
  __half2 b = __floats2half2_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
// 8 consecutive elements (32B float / 16B bf16 aligned)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}

__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// Precise and fast transcendental flavours.  PRECISE is used by the fp32 parity mode (bit-level
// token parity against the CPU reference needs full-precision tanhf/expf).
template <bool PRECISE> __device__ __forceinline__ float tanh_(float x) {
  if (PRECISE) return tanhf(x);
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool PRECISE> __device__ __forceinline__ float sigmoid_(float x) {
  if (PRECISE) return 1.0f / (1.0f + expf(-x));
  // sigmoid(x) = 0.5*tanh(0.5x)+0.5 : one MUFU op
  return fmaf(0.5f, tanh_<false>(0.5f * x), 0.5f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// monotonic float <-> int key (total order of the finite floats and infinities): integer max on keys == max on floats
__device__ __forceinline__ int f2key(float x) {
  const int b = __float_as_int(x);
  return b >= 0 ? b : (b ^ 0x7fffffff);
}
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); }

// Warp arg-max in the order (value descending, index ascending) with two REDUX instructions instead of a five-stage
// shuffle butterfly.  Lanes without a candidate pass valid = false; if no lane has one, wi = 0x7fffffff.
// (v + 0.0f turns -0.0 into +0.0 so that the key order agrees with the float comparison.)
__device__ __forceinline__ void warp_argmax(bool valid, float v, int idx, float& wv, int& wi) {
  const int key = valid ? f2key(v + 0.0f) : (int)0x80000000;
  const int kmax = __reduce_max_sync(0xffffffffu, key);
  wi = __reduce_min_sync(0xffffffffu, (valid && key == kmax) ? idx : 0x7fffffff);
  wv = (wi == 0x7fffffff) ? -INFINITY : key2f(kmax);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The decode loop is ~9 dependent launches per step, each 10-90 us: launch latency and kernel prologues (barrier init,
// TMEM allocation, tensor-map fetch, loads of step-invariant data) are a visible share of the step.  Kernels launched
// through launch_pdl may start while their predecessor in the stream is still running; they call pdl_wait() before
// touching anything the predecessor writes (or reads: every global store comes after the wait), and
// pdl_launch_dependents() right after it, so that at most one successor is resident ahead of time.
// A kernel launched without the attribute returns from pdl_wait() immediately.  VC_DISABLE_PDL=1: plain launches (A/B).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  const char* e = getenv("VC_DISABLE_PDL");          // read per launch: tests toggle it inside one process
  return !(e != nullptr && e[0] == '1');
}

// cluster_x > 1: thread-block clusters of that many CTAs along x (grid.x must be a multiple of it)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cluster_x > 1) {
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = (unsigned)cluster_x;
    at[1].val.clusterDim.y = 1;
    at[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  return launch_pdl_cluster(kern, grid, block, smem, stream, 1, static_cast<Args&&>(args)...);
}

}  // namespace vc
