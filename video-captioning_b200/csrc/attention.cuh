// Fused attention step: scoring + mask + softmax over T + context, one launch per decode step.
//
// Replaces per step (reference file:line, src/models/attention.py):
//   Bahdanau   :56-57 add,tanh,v-GEMV  :61 masked_fill  :64 softmax  :68-71 bmm
//   Luong      :118-146 score          :174-185 mask/softmax/bmm
//   Multi-head :250 QK^T/sqrt(d)       :253-258 mask/softmax  :262-267 w.V, concat  :273 head mean
// The loop-invariant projections (:52 keys, :140 linear_context, :241-242 K,V) are hoisted to one GEMM
// per video batch (attn_precompute); the query projections (:53, :128, :138, :240) are a GEMM over all
// rows just before this kernel.
//
// One CTA per VIDEO: the K beam rows of a video share its keys/values tile, which is therefore read
// from HBM once per video-step (not once per row).
//   scores : warps take pairs of frames, lanes own 16-byte chunks of the feature dimension (coalesced
//            512B per warp-load); the K queries are read from shared memory once per chunk and reused
//            for both frames; per-(frame,beam) partials are reduced with warp shuffles.
//            bf16 mode evaluates tanh two at a time (tanh.approx.f16x2: one MUFU op per pair) because
//            the additive form is MUFU-bound (R*T*A tanh per step), then accumulates in fp32.
//   softmax: one warp per (beam, head) row.
//   context: every thread owns 8 feature columns and a slice of the frames; slices are combined
//            through shared memory in a fixed order (deterministic).
#pragma once
#include <cuda_fp16.h>

#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace vc {

enum AttnMode : int { ATTN_ADDITIVE = 0, ATTN_DOT = 1, ATTN_MHA = 2 };

template <class T>
struct AttnArgs {
  // scoring operand per video: [B,T,D] (additive: projected keys, D=A; dot: enc_out, D=H; mha: K, D=H).
  // Element type KT of the kernel: T, except fp16 for the additive form in bf16 mode (packed half2 math).
  const void* skeys;
  // value operand per video: [B,T,H] (enc_out; mha: V)
  const T* values;
  const float* q;        // [R, D] fp32 query (projected); nullptr when q_act is used
  const T* q_act;        // [R, *] raw hidden state used as the query (Luong dot), row stride q_ld
  int64_t q_ld;
  const float* v;        // additive: [A] score vector;  nullptr otherwise
  float v_bias;          // additive (Bahdanau) bias of attention_linear
  const float* mask;     // [B,T] (0 -> masked) or nullptr
  T* ctx;                // context destination, row stride ctx_ld (written for every row r = b*K + k)
  int64_t ctx_ld;
  float* attn_out;       // optional attention weights destination [R, attn_ld] (+ offset applied by caller)
  int64_t attn_ld;
  int B, K, T_, D, H, heads;
  float scale;           // mha: 1/sqrt(d)
};

// sum_j v[j] * tanh(e[j] + q[j]) over 8 elements
template <bool PRECISE>
__device__ __forceinline__ float additive8(const float (&e)[8], const float* __restrict__ q, const float (&v)[8]) {
  float s = 0.f;
  if (PRECISE) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(v[j], tanhf(e[j] + q[j]), s);
  } else {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      __half2 x = __floats2half2_rn(e[j] + q[j], e[j + 1] + q[j + 1]);
      uint32_t xi = *reinterpret_cast<uint32_t*>(&x), yi;
      asm("tanh.approx.f16x2 %0, %1;" : "=r"(yi) : "r"(xi));
      float2 y = __half22float2(*reinterpret_cast<__half2*>(&yi));
      s = fmaf(v[j], y.x, s);
      s = fmaf(v[j + 1], y.y, s);
    }
  }
  return s;
}

// Packed fp16 form: sum over 8 elements of v * tanh(key + q), operands as 4 half2 words each.
// 4 HADD2 + 4 MUFU.TANH(f16x2) + 4 HFMA2, the 4-term half2 partial is then widened to fp32.
__device__ __forceinline__ float additive8_h2(const uint4& key, const uint4& q, const uint4& v) {
  const uint32_t kw[4] = {key.x, key.y, key.z, key.w};
  const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
  const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
  __half2 acc = __float2half2_rn(0.f);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    __half2 x = __hadd2(*reinterpret_cast<const __half2*>(&kw[p]), *reinterpret_cast<const __half2*>(&qw[p]));
    uint32_t xi = *reinterpret_cast<uint32_t*>(&x), yi;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(yi) : "r"(xi));
    acc = __hfma2(*reinterpret_cast<const __half2*>(&vw[p]), *reinterpret_cast<__half2*>(&yi), acc);
  }
  const float2 f = __half22float2(acc);
  return f.x + f.y;
}

constexpr int kAttnThreads = 256;
constexpr int kAttnFR = 2;   // frames per warp iteration

// KMAX: compile-time bound on beams handled per CTA (K <= KMAX).
// NCH > 0 selects the register-resident scoring loop of the packed additive form: K == KMAX exactly and
// D <= 256*NCH; every lane keeps its NCH 16-byte chunks of all K queries and of v in registers for the
// whole kernel, so the frame loop is LDG(key) + HADD2/MUFU/HFMA2 only.
template <class T, class KT, int MODE, int KMAX, bool PRECISE, int NCH>
__global__ void __launch_bounds__(kAttnThreads) attn_step_kernel(const AttnArgs<T> a) {
  constexpr bool PACKED = (MODE == ATTN_ADDITIVE) && std::is_same<KT, __half>::value;   // q, v staged as fp16
  static_assert(NCH == 0 || PACKED, "register-resident scoring is implemented for the packed additive form");
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x;
  const int K = a.K, Tn = a.T_, D = a.D, H = a.H;
  const int NH = (MODE == ATTN_MHA) ? a.heads : 1;
  const int dh = D / NH;                                   // scoring columns per head
  float* q_s = smem;                                       // [K][D] fp32 (PACKED: fp16 in the same space)
  float* v_s = q_s + (size_t)K * D;                        // [D] (additive only)
  float* sc = v_s + ((MODE == ATTN_ADDITIVE) ? D : 0);     // [K][NH][Tn]
  float* red = sc + (size_t)K * NH * Tn;                   // [G][K][H] context partials
  __half* q_h = reinterpret_cast<__half*>(q_s);
  __half* v_h = reinterpret_cast<__half*>(v_s);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarp = kAttnThreads / 32;

  for (int i = tid; i < K * D; i += kAttnThreads) {
    const int k = i / D, d = i - k * D;
    const int64_t r = (int64_t)b * K + k;
    const float qv = a.q ? a.q[r * D + d] : to_float(a.q_act[r * a.q_ld + d]);
    if (PACKED) q_h[i] = __float2half_rn(qv);
    else q_s[i] = qv;
  }
  if (MODE == ATTN_ADDITIVE)
    for (int i = tid; i < D; i += kAttnThreads) {
      if (PACKED) v_h[i] = __float2half_rn(a.v[i]);
      else v_s[i] = a.v[i];
    }
  if (MODE == ATTN_MHA)
    for (int i = tid; i < K * NH * Tn; i += kAttnThreads) sc[i] = 0.f;
  __syncthreads();

  // ---- scores
  const KT* sk = reinterpret_cast<const KT*>(a.skeys) + (int64_t)b * Tn * D;
  const int group = (MODE == ATTN_MHA) ? min(32, dh / 8) : 32;   // lanes reducing together (one head)
  if constexpr (NCH > 0) {
    uint4 qreg[KMAX][NCH], vreg[NCH];
    bool live[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d0 = lane * 8 + 256 * c;
      live[c] = d0 < D;
      vreg[c] = live[c] ? *reinterpret_cast<const uint4*>(v_h + d0) : make_uint4(0, 0, 0, 0);   // v = 0: no contribution
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        qreg[k][c] = live[c] ? *reinterpret_cast<const uint4*>(q_h + k * D + d0) : make_uint4(0, 0, 0, 0);
    }
    uint4 key[NCH], nxt[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      key[c] = (live[c] && warp < Tn) ? *reinterpret_cast<const uint4*>(sk + (int64_t)warp * D + lane * 8 + 256 * c)
                                      : make_uint4(0, 0, 0, 0);
    for (int t = warp; t < Tn; t += nwarp) {
      const int tn = t + nwarp;
#pragma unroll
      for (int c = 0; c < NCH; ++c)   // prefetch the next frame of this warp
        nxt[c] = (live[c] && tn < Tn) ? *reinterpret_cast<const uint4*>(sk + (int64_t)tn * D + lane * 8 + 256 * c)
                                      : make_uint4(0, 0, 0, 0);
      float part[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        float p = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) p += additive8_h2(key[c], qreg[k][c], vreg[c]);
        part[k] = p;
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const float sres = warp_sum(part[k]);
        if (lane == 0) sc[(size_t)k * Tn + t] = sres + a.v_bias;
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) key[c] = nxt[c];
    }
  } else
  for (int t0 = warp * kAttnFR; t0 < Tn; t0 += nwarp * kAttnFR) {
    float part[kAttnFR][KMAX];
#pragma unroll
    for (int f = 0; f < kAttnFR; ++f)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) part[f][k] = 0.f;
    if constexpr (PACKED) {
      for (int d0 = lane * 8; d0 < D; d0 += 256) {
        uint4 key[kAttnFR];
#pragma unroll
        for (int f = 0; f < kAttnFR; ++f)
          key[f] = (t0 + f < Tn) ? *reinterpret_cast<const uint4*>(sk + (int64_t)(t0 + f) * D + d0) : make_uint4(0, 0, 0, 0);
        const uint4 v8 = *reinterpret_cast<const uint4*>(v_h + d0);
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) {
            const uint4 q8 = *reinterpret_cast<const uint4*>(q_h + k * D + d0);
#pragma unroll
            for (int f = 0; f < kAttnFR; ++f) part[f][k] += additive8_h2(key[f], q8, v8);
          }
      }
    } else
    for (int d0 = lane * 8; d0 < D || (MODE == ATTN_MHA && d0 - lane * 8 < D); d0 += 256) {
      const bool live = d0 < D;
      float e[kAttnFR][8];
#pragma unroll
      for (int f = 0; f < kAttnFR; ++f) {
        if (live && t0 + f < Tn) load8(sk + (int64_t)(t0 + f) * D + d0, e[f]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) e[f][j] = 0.f;
        }
      }
      if (live) {
        float v8[8];
        if (MODE == ATTN_ADDITIVE) load8(v_s + d0, v8);
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) {
            float q8[8];
            load8(q_s + k * D + d0, q8);
#pragma unroll
            for (int f = 0; f < kAttnFR; ++f) {
              if (MODE == ATTN_ADDITIVE) {
                part[f][k] += additive8<PRECISE>(e[f], q8, v8);
              } else {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s = fmaf(e[f][j], q8[j], s);
                part[f][k] += s;
              }
            }
          }
      }
      if (MODE == ATTN_MHA) {
        // one pass covers 256 columns = 256/dh heads (or part of one head): reduce inside the lane
        // group of this head and accumulate into sc (this warp is the only writer of frames t0..)
        const int hd = live ? d0 / dh : 0;
#pragma unroll
        for (int f = 0; f < kAttnFR; ++f)
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k < K) {
              float s = part[f][k];
              for (int o = group >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
              if (live && (lane % group) == 0 && t0 + f < Tn) sc[((size_t)k * NH + hd) * Tn + t0 + f] += s;
              part[f][k] = 0.f;
            }
      }
    }
    if (MODE != ATTN_MHA) {
#pragma unroll
      for (int f = 0; f < kAttnFR; ++f)
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) {
            const float s = warp_sum(part[f][k]);
            if (lane == 0 && t0 + f < Tn) sc[(size_t)k * Tn + t0 + f] = (MODE == ATTN_ADDITIVE) ? s + a.v_bias : s;
          }
    }
  }
  __syncthreads();

  // ---- masked softmax over T per (beam, head): one warp per row of sc
  for (int row = warp; row < K * NH; row += nwarp) {
    float* s = sc + (size_t)row * Tn;
    float m = -INFINITY;
    for (int t = lane; t < Tn; t += 32) {
      float x = s[t];
      if (MODE == ATTN_MHA) x *= a.scale;
      if (a.mask != nullptr && a.mask[(int64_t)b * Tn + t] == 0.f) x = -1e9f;
      s[t] = x;
      m = fmaxf(m, x);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < Tn; t += 32) {
      const float e = PRECISE ? expf(s[t] - m) : __expf(s[t] - m);
      s[t] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int t = lane; t < Tn; t += 32) s[t] = PRECISE ? s[t] / sum : s[t] * inv;
  }
  __syncthreads();

  // ---- optional attention-weight output (mean over heads for MHA, attention.py:273)
  if (a.attn_out != nullptr) {
    for (int i = tid; i < K * Tn; i += kAttnThreads) {
      const int k = i / Tn, t = i - k * Tn;
      float w = 0.f;
      for (int hd = 0; hd < NH; ++hd) w += sc[((size_t)k * NH + hd) * Tn + t];
      if (NH > 1) w /= (float)NH;
      a.attn_out[((int64_t)b * K + k) * a.attn_ld + t] = w;
    }
  }

  // ---- context: thread = 8 columns x a slice of the frames; G slices combined through smem
  const T* vv = a.values + (int64_t)b * Tn * H;
  const int cols8 = H / 8;                       // threads needed to cover H
  const int G = max(1, kAttnThreads / cols8);    // frame slices
  const int dhv = H / NH;
  for (int c0 = 0; c0 < cols8; c0 += kAttnThreads) {   // more than one pass only when H > 2048
    const int ci = c0 + (tid % min(cols8, kAttnThreads));
    const int g = tid / min(cols8, kAttnThreads);
    const int h0 = ci * 8;
    float acc[KMAX][8];
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
    if (g < G && ci < cols8) {
      const int hd = (MODE == ATTN_MHA) ? (h0 / dhv) : 0;
      for (int t = g; t < Tn; t += G) {
        float e[8];
        load8(vv + (int64_t)t * H + h0, e);
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) {
            const float w = sc[((size_t)k * NH + hd) * Tn + t];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(w, e[j], acc[k][j]);
          }
      }
      if (G > 1) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) {
            float* dst = red + ((size_t)g * K + k) * H + h0;
            *reinterpret_cast<float4*>(dst) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
          }
      }
    }
    if (G > 1) {
      __syncthreads();
      // fixed-order combine: thread i handles 4 consecutive columns of one beam
      for (int i = tid; i < K * (H / 4); i += kAttnThreads) {
        const int k = i / (H / 4), h4 = (i - k * (H / 4)) * 4;
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        for (int gg = 0; gg < G; ++gg) {
          const float4 p = *reinterpret_cast<const float4*>(red + ((size_t)gg * K + k) * H + h4);
          o[0] += p.x; o[1] += p.y; o[2] += p.z; o[3] += p.w;
        }
        store4(a.ctx + ((int64_t)b * K + k) * a.ctx_ld + h4, o);
      }
      __syncthreads();
    } else if (ci < cols8) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) {
          float lo[4] = {acc[k][0], acc[k][1], acc[k][2], acc[k][3]};
          float hi[4] = {acc[k][4], acc[k][5], acc[k][6], acc[k][7]};
          store4(a.ctx + ((int64_t)b * K + k) * a.ctx_ld + h0, lo);
          store4(a.ctx + ((int64_t)b * K + k) * a.ctx_ld + h0 + 4, hi);
        }
    }
  }
}

// ---------------------------------------------------------------- additive attention, bf16 mode (v3)
// The Bahdanau / Luong-concat step of the benchmark configuration.  R*T*A tanh per step make it MUFU-bound
// (16 tanh/clk/SM; tanh.approx.f16x2 retires 2 results per lane but at half the instruction rate), so
// the kernel is built to keep the XU pipe fed: small CTAs (one video, 4 warps, <= 64 registers, ~6 KB smem)
// so that 8 CTAs per SM -- every video of a 1024-video step at once -- interleave their MUFU-heavy scoring
// with the FFMA-heavy context phase of their neighbours.
//   queries  fp16 [R, D], written by the query-projection GEMM's epilogue; every lane keeps its 16-byte
//            chunk of the K queries (and of v) in registers for the whole kernel -- no smem staging
//   scoring  warp = (D half, frame slot); per frame pair: 2 LDG.128 of fp16 keys, K x 2 x 4 (HADD2, MUFU.TANH,
//            HFMA2), fp32 butterfly (first stage folds the two frames), partial sums of the two D halves
//            combined in the softmax phase
//   softmax  one warp per (beam) row
//   context  thread = 4 columns of enc_out over all frames (no cross-thread reduction), fp32 FFMA,
//            weights read as float4 from smem
struct AttnAddArgs {
  const __half* keys;    // [B, T, D] fp16 projected keys (attention.py:52 / :140)
  const __half* q;       // [R, D] fp16 projected queries incl. bias (attention.py:53 / :138)
  const int* q_rows;     // nullptr, or [R]: row r's query is q[q_rows[r]] -- the projection ran on the rows BEFORE the beam
                         // reorder (overlapped with the selection), so the reorder is applied here as an indirection
  const __half* v;       // [D] fp16 score vector (attention.py:28 / :99)
  float v_bias;
  const bf16* values;    // [B, T, H] enc_out
  const float* mask;     // [B, T] or nullptr
  bf16* ctx;             // [R, ctx_ld]
  int64_t ctx_ld;
  float* attn_out;       // optional [R, attn_ld]
  int64_t attn_ld;
  int B, T, D, H;
  // Scoring gate: at most `sem_limit` CTAs per SM are in the MUFU-bound scoring phase at a time (counter per SM
  // in global memory, zero before and after every launch).  All CTAs of a step are resident at once; without the
  // gate they run scoring together (XU saturated, stretched 7x) and then the context phase together (XU idle).
  // With it a CTA scores at nearly full XU rate and its FFMA/memory-bound context phase overlaps the scoring of
  // the next CTAs.  nullptr: no gate.
  int* sm_sem;
  int sem_limit;
#ifdef VC_ATTN_PROBE
  long long* dbg;        // probe builds (scripts/attn_ws_probe.cu): clock64 stamps of CTA 0, [unit][event]
#endif
};
#ifdef VC_ATTN_PROBE
#define ATTN_PROBE(cond, unit, ev) do { if (a.dbg != nullptr && blockIdx.x == 0 && (cond) && (unit) < 64) a.dbg[(unit) * 16 + (ev)] = clock64(); } while (0)
#else
#define ATTN_PROBE(cond, unit, ev) do { } while (0)
#endif

constexpr int kAddThreads = 128;

// 4 half2 words of (key + q) -> tanh -> * v, accumulated in one half2
__device__ __forceinline__ __half2 additive4_h2(const uint4& key, const uint4& q, const uint4& v) {
  const uint32_t kw[4] = {key.x, key.y, key.z, key.w};
  const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
  const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
  __half2 acc = __float2half2_rn(0.f);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    __half2 x = __hadd2(*reinterpret_cast<const __half2*>(&kw[p]), *reinterpret_cast<const __half2*>(&qw[p]));
    uint32_t xi = *reinterpret_cast<uint32_t*>(&x), yi;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(yi) : "r"(xi));
    acc = __hfma2(*reinterpret_cast<const __half2*>(&vw[p]), *reinterpret_cast<__half2*>(&yi), acc);
  }
  return acc;
}
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// the same load as a volatile asm statement: it keeps its place among the other volatile asm statements (the MMAs)
__device__ __forceinline__ uint4 ldg128_pinned(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// coherent 16-byte load: for data written by the previous kernel of the stream and read after pdl_wait()
__device__ __forceinline__ uint4 ld128(const void* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

// K: exact beam count; DH: warps sharing one frame (2: D in (256, 512], 1: D <= 256)
template <int K, int DH>
__global__ void __launch_bounds__(kAddThreads, 7) attn_additive_kernel(const AttnAddArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, D = a.D, H = a.H;
  const int Tp = (T + 3) & ~3;                   // weights padded to a multiple of 4 frames (zeros)
  float* part = smem;                            // [DH][K][Tp] partial scores
  float* wgt = smem + DH * K * Tp;               // [K][Tp] softmax weights
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NFS = 4 / DH;                    // frame slots
  const int dh = (DH == 2) ? (warp & 1) : 0;
  const int fs = (DH == 2) ? (warp >> 1) : warp;

  // ---- scores
  int* sem = nullptr;
  {
    const int d0 = (dh * 32 + lane) * 8;
    const bool live = d0 < D;
    uint4 qreg[K];
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int64_t qr = a.q_rows ? (int64_t)a.q_rows[b * K + k] : (int64_t)b * K + k;
      qreg[k] = live ? ld128(a.q + qr * D + d0) : zero4;
    }
    const uint4 vreg = live ? ldg128(a.v + d0) : zero4;           // v = 0: dead lanes contribute nothing
    const __half* kp = a.keys + (int64_t)b * T * D + d0;
    float* prow = part + dh * K * Tp;
    // software pipeline: the next frame pair's keys are in flight while this pair's tanh work runs
    uint4 key0 = (live && fs < T) ? ldg128(kp + (int64_t)fs * D) : zero4;
    uint4 key1 = (live && fs + NFS < T) ? ldg128(kp + (int64_t)(fs + NFS) * D) : zero4;
    if (a.sm_sem != nullptr) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      sem = a.sm_sem + smid;
      if (tid == 0) {
        while (atomicAdd(sem, 1) >= a.sem_limit) {
          atomicSub(sem, 1);
          __nanosleep(400);
        }
      }
      __syncthreads();
    }
    for (int t = fs; t < T; t += 2 * NFS) {
      const int t1 = t + NFS;
      const int tn0 = t + 2 * NFS, tn1 = t1 + 2 * NFS;
      const uint4 nxt0 = (live && tn0 < T) ? ldg128(kp + (int64_t)tn0 * D) : zero4;
      const uint4 nxt1 = (live && tn1 < T) ? ldg128(kp + (int64_t)tn1 * D) : zero4;
      float s0[K], s1[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float2 f0 = __half22float2(additive4_h2(key0, qreg[k], vreg));
        const float2 f1 = __half22float2(additive4_h2(key1, qreg[k], vreg));
        s0[k] = f0.x + f0.y;
        s1[k] = f1.x + f1.y;
      }
      // butterfly: the xor-16 stage folds the two frames (lanes 0-15 end up with frame t, lanes 16-31 with t1)
      const bool hi = (lane & 16) != 0;
      float s[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float send = hi ? s0[k] : s1[k];
        const float keep = hi ? s1[k] : s0[k];
        s[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
      if ((lane & 15) == 0) {
        const int tt = hi ? t1 : t;
        if (tt < T) {
#pragma unroll
          for (int k = 0; k < K; ++k) prow[k * Tp + tt] = s[k];
        }
      }
      key0 = nxt0;
      key1 = nxt1;
    }
  }
  __syncthreads();
  if (sem != nullptr && tid == 0) atomicSub(sem, 1);

  // ---- masked softmax over T, one warp per beam row (attention.py:61-64)
  for (int k = warp; k < K; k += kAddThreads / 32) {
    float m = -INFINITY;
    for (int t = lane; t < T; t += 32) {
      float x = part[k * Tp + t] + a.v_bias;
      if (DH == 2) x += part[(K + k) * Tp + t];
      if (a.mask != nullptr && a.mask[(int64_t)b * T + t] == 0.f) x = -1e9f;
      wgt[k * Tp + t] = x;
      m = fmaxf(m, x);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float e = __expf(wgt[k * Tp + t] - m);
      wgt[k * Tp + t] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int t = lane; t < Tp; t += 32) {
      const float w = (t < T) ? wgt[k * Tp + t] * inv : 0.f;
      wgt[k * Tp + t] = w;
      if (a.attn_out != nullptr && t < T) a.attn_out[((int64_t)b * K + k) * a.attn_ld + t] = w;
    }
  }
  __syncthreads();

  // ---- context = sum_t w[k][t] * enc[t][:]  (attention.py:68-71): 4 columns per thread, all frames
  const bf16* vv = a.values + (int64_t)b * T * H;
  for (int c0 = tid * 4; c0 < H; c0 += kAddThreads * 4) {
    float acc[K][4];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
    constexpr int FB = 8;                          // frames per batch: 8 independent 8-byte loads in flight per thread
    for (int t = 0; t < Tp; t += FB) {
      uint2 e[FB];
#pragma unroll
      for (int f = 0; f < FB; ++f)
        e[f] = (t + f < T) ? __ldg(reinterpret_cast<const uint2*>(vv + (int64_t)(t + f) * H + c0)) : make_uint2(0u, 0u);
#pragma unroll
      for (int g4 = 0; g4 < FB / 4; ++g4) {
        if (t + 4 * g4 < Tp) {
          float4 w4[K];
#pragma unroll
          for (int k = 0; k < K; ++k) w4[k] = *reinterpret_cast<const float4*>(wgt + k * Tp + t + 4 * g4);
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const uint2 ee = e[4 * g4 + f];
            const float x0 = __uint_as_float(ee.x << 16), x1 = __uint_as_float(ee.x & 0xffff0000u);
            const float x2 = __uint_as_float(ee.y << 16), x3 = __uint_as_float(ee.y & 0xffff0000u);
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const float w = (f == 0) ? w4[k].x : (f == 1) ? w4[k].y : (f == 2) ? w4[k].z : w4[k].w;
              acc[k][0] = fmaf(w, x0, acc[k][0]);
              acc[k][1] = fmaf(w, x1, acc[k][1]);
              acc[k][2] = fmaf(w, x2, acc[k][2]);
              acc[k][3] = fmaf(w, x3, acc[k][3]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) store4(a.ctx + ((int64_t)b * K + k) * a.ctx_ld + c0, acc[k]);
  }
}

// ---------------------------------------------------------------- additive attention, bf16 mode (v4: tensor-core reductions)
// v3 above is bound by instruction issue (4.9 instructions per tanh: packed add, 2 MUFU, pack, packed fma,
// fp16->fp32 conversions and a 5-stage shuffle butterfly per frame and beam), not yet by the XU pipe.  Here both
// reductions of the step run on the tensor cores through mma.sync (register fragments, no shared-memory operands):
//   scores   s[t,k] = sum_j v[j] * tanh(key[t,j] + q[k,j]).  A 16x16 A fragment holds tanh values for 16 frames x
//            16 features, B is v replicated over the 8 columns, the fp32 accumulator tile carries the running dot
//            product over all feature blocks: lane (g, tg) evaluates frames g / g+8 on the 8 features of its tg --
//            the dot product does not care how features are permuted as long as key, q and v agree, so every lane's
//            8 features are one contiguous 16-byte chunk.  2.2 instructions per tanh; no shuffles, no conversions.
//   context  ctx^T[col, k] = sum_t enc[t, col] * w[k, t]: A = enc^T (ldmatrix.trans from a staged 16-frame tile),
//            B = softmax weights split into bf16 hi + lo parts (two MMAs: fp32-grade weights), 8 m-tiles per warp.
// Warps split the feature dimension for the scores (partials summed in the softmax) and the columns for the context.
constexpr int kMmaThreads = 128;
constexpr int kEncPad = 8;   // bf16 elements of row padding in the staged enc tile (ldmatrix rows hit distinct banks)

__device__ __forceinline__ uint32_t tanh_h2(uint32_t key, uint32_t q) {
  __half2 x = __hadd2(*reinterpret_cast<const __half2*>(&key), *reinterpret_cast<const __half2*>(&q));
  uint32_t xi = *reinterpret_cast<uint32_t*>(&x), yi;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(yi) : "r"(xi));
  return yi;
}
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// split two fp32 weights into packed bf16 hi and lo parts (hi + lo reproduces ~16 mantissa bits)
__device__ __forceinline__ void split_bf16x2(float w0, float w1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
  const __nv_bfloat162 h = __halves2bfloat162(h0, h1);
  const __nv_bfloat162 l = __floats2bfloat162_rn(w0 - __bfloat162float(h0), w1 - __bfloat162float(h1));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// K: exact beam count (<= 8).  Requires D % 32 == 0, H % 64 == 0.
template <int K, int NBUF, int MINB>
__global__ void __launch_bounds__(kMmaThreads, MINB) attn_additive_mma_kernel(const AttnAddArgs a) {
  extern __shared__ __align__(16) uint8_t smem_u8[];
  const int T = a.T, D = a.D, H = a.H;
  const int NT = (T + 15) >> 4;                  // 16-frame tiles
  const int Tp = NT * 16;
  const int pitch = H + kEncPad;                 // staged enc row pitch (elements)
  __half* q_s = reinterpret_cast<__half*>(smem_u8);                        // [K][D]
  __half* v_s = q_s + K * D;                                               // [D]
  float* part = reinterpret_cast<float*>(v_s + D);                         // [4][K][Tp]
  float* wgt = part + 4 * K * Tp;                                          // [K][Tp]
  bf16* enc_s = reinterpret_cast<bf16*>(wgt + K * Tp);                     // [2][16][pitch]; later [K][H] output staging
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const bf16* vv = a.values + (int64_t)b * T * H;

  // stage one 16-frame tile of enc_out (frames past T are clamped; their weights are zero)
  auto stage_enc = [&](int ks, int buf) {
    const int chunks_per_row = H / 8;            // 16-byte chunks
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(enc_s + (size_t)buf * 16 * pitch);
    for (int rr = 0; rr < 16; ++rr) {            // (no integer division: it would run on the XU pipe the tanh needs)
      int t = ks * 16 + rr;
      t = t < T ? t : T - 1;
      for (int cc = tid; cc < chunks_per_row; cc += kMmaThreads)
        cp_async16(base + (uint32_t)(rr * pitch + cc * 8) * 2u, vv + (int64_t)t * H + cc * 8);
    }
    cp_async_commit();
  };
  stage_enc(0, 0);                               // lands during the scoring phase

  // score vector + queries -> shared memory (fp16).  enc_out, keys and v do not change during the decode loop:
  // only the queries (and the ctx output) depend on the previous kernel in the stream (PDL, common.cuh).
  for (int i = tid; i < D / 8; i += kMmaThreads) reinterpret_cast<uint4*>(v_s)[i] = ldg128(a.v + (int64_t)i * 8);
  pdl_wait();
  pdl_launch_dependents();
  for (int i = tid; i < K * D / 8; i += kMmaThreads) {
    const int k = (i * 8) / D, off = i * 8 - k * D;
    const int64_t qr = a.q_rows ? (int64_t)a.q_rows[b * K + k] : (int64_t)b * K + k;
    reinterpret_cast<uint4*>(q_s)[i] = ld128(a.q + qr * D + off);
  }
  __syncthreads();

  // ---- scores: warp w owns feature blocks fb = w, w+4, ... (32 features each)
  {
    const int nfb = D / 32;
    const __half* kbase = a.keys + (int64_t)b * T * D;
    for (int ft = 0; ft < NT; ++ft) {
      int t0 = ft * 16 + g, t1 = t0 + 8;
      const int tc0 = t0 < T ? t0 : T - 1, tc1 = t1 < T ? t1 : T - 1;
      const __half* k0p = kbase + (int64_t)tc0 * D + tg * 8;
      const __half* k1p = kbase + (int64_t)tc1 * D + tg * 8;
      float acc[K][4];
#pragma unroll
      for (int k = 0; k < K; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }
      uint4 key0 = make_uint4(0u, 0u, 0u, 0u), key1 = key0;
      if (warp < nfb) { key0 = ldg128(k0p + warp * 32); key1 = ldg128(k1p + warp * 32); }
      for (int fb = warp; fb < nfb; fb += 4) {
        const int feat = fb * 32 + tg * 8;
        uint4 nx0 = make_uint4(0u, 0u, 0u, 0u), nx1 = nx0;
        if (fb + 4 < nfb) { nx0 = ldg128(k0p + (fb + 4) * 32); nx1 = ldg128(k1p + (fb + 4) * 32); }
        const uint4 v4 = *reinterpret_cast<const uint4*>(v_s + feat);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const uint4 q4 = *reinterpret_cast<const uint4*>(q_s + k * D + feat);
          mma_f16(acc[k], tanh_h2(key0.x, q4.x), tanh_h2(key1.x, q4.x), tanh_h2(key0.y, q4.y), tanh_h2(key1.y, q4.y), v4.x, v4.y);
          mma_f16(acc[k], tanh_h2(key0.z, q4.z), tanh_h2(key1.z, q4.z), tanh_h2(key0.w, q4.w), tanh_h2(key1.w, q4.w), v4.z, v4.w);
        }
        key0 = nx0;
        key1 = nx1;
      }
      if (tg == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          part[(warp * K + k) * Tp + t0] = acc[k][0];     // every column of the accumulator tile holds the dot product
          part[(warp * K + k) * Tp + t1] = acc[k][2];
        }
      }
    }
  }
  __syncthreads();

  // ---- masked softmax over T, one warp per beam row (attention.py:61-64)
  for (int k = warp; k < K; k += kMmaThreads / 32) {
    float m = -INFINITY;
    for (int t = lane; t < T; t += 32) {
      float x = part[k * Tp + t] + part[(K + k) * Tp + t] + part[(2 * K + k) * Tp + t] + part[(3 * K + k) * Tp + t] + a.v_bias;
      if (a.mask != nullptr && a.mask[(int64_t)b * T + t] == 0.f) x = -1e9f;
      wgt[k * Tp + t] = x;
      m = fmaxf(m, x);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float e = __expf(wgt[k * Tp + t] - m);
      wgt[k * Tp + t] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int t = lane; t < Tp; t += 32) {
      const float w = (t < T) ? wgt[k * Tp + t] * inv : 0.f;
      wgt[k * Tp + t] = w;
      if (a.attn_out != nullptr && t < T) a.attn_out[((int64_t)b * K + k) * a.attn_ld + t] = w;
    }
  }
  // (the barrier that publishes wgt is the first one of the loop below)

  // ---- context: warp w owns columns [w*H/4, (w+1)*H/4) as m-tiles of 16
  {
    const int cw = H / 4;                        // columns per warp
    const int nmt = cw / 16;                     // m-tiles per warp (8 at H = 512)
    constexpr int MAXMT = 8;
    for (int mt0 = 0; mt0 < nmt; mt0 += MAXMT) { // one pass for H <= 512
      float c[MAXMT][4];
#pragma unroll
      for (int i = 0; i < MAXMT; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
      for (int ks = 0; ks < NT; ++ks) {
        const int buf = (mt0 == 0 && NBUF == 2) ? (ks & 1) : 0;
        if (mt0 == 0 && NBUF == 1) {
          if (ks > 0) { __syncthreads(); stage_enc(ks, 0); }      // tile ks-1 consumed by everyone
          cp_async_wait<0>();
        } else if (mt0 == 0) {
          if (ks + 1 < NT) { stage_enc(ks + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        } else {
          __syncthreads();
          stage_enc(ks, 0);
          cp_async_wait<0>();
        }
        __syncthreads();                         // tile ks visible to all warps (and wgt on the first trip)
        // B fragments: weights of beam g for frames ks*16 + {2tg, 2tg+1} and + 8, hi/lo split
        uint32_t bh0 = 0u, bl0 = 0u, bh1 = 0u, bl1 = 0u;
        if (g < K) {
          const float2 w0 = *reinterpret_cast<const float2*>(wgt + g * Tp + ks * 16 + 2 * tg);
          const float2 w1 = *reinterpret_cast<const float2*>(wgt + g * Tp + ks * 16 + 2 * tg + 8);
          split_bf16x2(w0.x, w0.y, bh0, bl0);
          split_bf16x2(w1.x, w1.y, bh1, bl1);
        }
        const uint32_t tile = (uint32_t)__cvta_generic_to_shared(enc_s + (size_t)buf * 16 * pitch);
        // ldmatrix row address of this lane: matrix j = lane/8 -> frames (j/2)*8 + lane%8, columns + (j%2)*8
        const int lrow = ((lane >> 4) << 3) + (lane & 7), lcol = ((lane >> 3) & 1) << 3;
#pragma unroll
        for (int i = 0; i < MAXMT; ++i) {
          if (mt0 + i < nmt) {
            const int col0 = warp * cw + (mt0 + i) * 16;
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4_trans(tile + (uint32_t)(lrow * pitch + col0 + lcol) * 2u, a0, a1, a2, a3);
            mma_bf16(c[i], a0, a1, a2, a3, bh0, bh1);
            mma_bf16(c[i], a0, a1, a2, a3, bl0, bl1);
          }
        }
        if (mt0 == 0 && NBUF == 2 && ks + 1 < NT) __syncthreads();   // everyone done with buf before it is refilled two trips later
      }
      // c[i] = {ctx[col0+g][2tg], ctx[col0+g][2tg+1], ctx[col0+g+8][2tg], ctx[col0+g+8][2tg+1]} (col, beam)
      __syncthreads();                           // all tiles consumed: reuse enc_s as [K][H] bf16 output staging
      bf16* out_s = enc_s;
#pragma unroll
      for (int i = 0; i < MAXMT; ++i) {
        if (mt0 + i < nmt) {
          const int col0 = warp * cw + (mt0 + i) * 16;
          const int k0 = 2 * tg, k1 = 2 * tg + 1;
          if (k0 < K) { out_s[k0 * H + col0 + g] = __float2bfloat16_rn(c[i][0]); out_s[k0 * H + col0 + g + 8] = __float2bfloat16_rn(c[i][2]); }
          if (k1 < K) { out_s[k1 * H + col0 + g] = __float2bfloat16_rn(c[i][1]); out_s[k1 * H + col0 + g + 8] = __float2bfloat16_rn(c[i][3]); }
        }
      }
      __syncthreads();
      if (mt0 + MAXMT >= nmt) {
#pragma unroll
        for (int k = 0; k < K; ++k)
          for (int c8 = tid * 8; c8 < H; c8 += kMmaThreads * 8)
            *reinterpret_cast<uint4*>(a.ctx + ((int64_t)b * K + k) * a.ctx_ld + c8) = *reinterpret_cast<const uint4*>(out_s + k * H + c8);
      }
    }
  }
}

// ---------------------------------------------------------------- additive attention, bf16 mode (v5: one persistent, warp-specialised CTA per SM)
// v4 above runs the MUFU-bound scoring of all resident CTAs first and their memory/tensor-bound context phases
// afterwards (every CTA of the single wave starts together), so the XU pipe idles ~40% of the kernel.  v5 decouples the
// two and streams over (video, 16-frame tile) units with an online softmax.  One CTA per SM owns the videos
// b = blockIdx.x + i*gridDim.x; its units are numbered u = i*NT + tile:
//   NG scoring groups of 4 warps   group j takes the units u = j (mod NG): tanh tile x v on mma.sync exactly as in v4
//                                  (warp = 4 of the D/32 feature blocks), partial dot products -> a ring of shared-memory
//                                  slots (mbarrier full/empty).  Interleaving by unit, not by video, balances the groups
//                                  to within one tile (7 videos x 5 tiles over 4 groups: 9/9/9/8); every warp keeps a
//                                  private double-buffered copy of the query columns it needs, so scoring warps never
//                                  meet at a block barrier.
//   NCG context groups of 4 warps  group c takes the videos i = c (mod NCG): sums the four partials of a unit, running
//                                  max / sum (flash-attention style rescaling of the fp32 accumulators), weights split
//                                  into bf16 hi+lo, enc^T x weights on mma.sync from a ring of enc tiles.
//   1 producer warp                fills the enc ring with cp.async.bulk (mbarrier tx), as far ahead as the ring allows.
// When the attention weights are requested the raw scores of a video stay in shared memory and are normalised with
// the final (max, sum) at the end of the video.
constexpr int kWsPartSlots = 8;
#ifndef VC_WS_ENC_SLOTS
#define VC_WS_ENC_SLOTS 6
#endif
constexpr int kWsEncSlots = VC_WS_ENC_SLOTS;

__device__ __forceinline__ void amb_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void amb_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void amb_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool amb_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void amb_wait(uint32_t bar, uint32_t parity) {   // bounded: a protocol bug traps instead of hanging
  for (uint32_t n = 0; !amb_try(bar, parity); ++n)
    if (n > (1u << 24)) __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// ring bookkeeping: advance (slot, parity) by n positions in a ring of N slots
__device__ __forceinline__ void ring_adv(int& slot, uint32_t& par, int n, int N) {
  slot += n;
  while (slot >= N) { slot -= N; par ^= 1u; }
}

// Row gather riding along with the persistent attention kernel (capi.cu:run_decode): the beam reorder of the decoder state
// and the next tokens' embeddings (video_captioning_model.py:247-249,269-272, decoder.py:130) are pure row copies that only
// depend on the selection before this kernel and only feed the LSTM GEMM after it.  The attention step is XU-bound and
// leaves two thirds of the HBM bandwidth unused, so two extra warps move the rows while the scoring warps work.  A row is a
// list of 512-byte pieces (one 16-byte chunk per lane): piece j of row r comes from row parent[r] (or row tok[r] of a table)
// of src[j] and goes to row r of dst[j].  All addresses and strides are multiples of 16 bytes.
// (A first version staged the rows through a shared-memory ring with cp.async.bulk in both directions: correct, but the
// extra 21-43 KB of shared memory moved the L1 carve-out and its bulk loads queued in front of the encoder tile producer's,
// +12 us per launch for a 20 us gather; scripts/attn_ws_probe.cu.)
constexpr int kGatherMaxPieces = 16;  // 8 KB per row: (h, c) of two layers at H = 512 + a 512-wide embedding are 14
constexpr int kGatherWarps = 2;
struct RowGather {
  int n_rows;            // 0: nothing to gather
  int n_pieces;
  int V;                 // rows of the token-indexed tables (tokens are clamped)
  int dbg;               // probe only: 1 = no loads, 2 = no stores
  const int* parent;     // [n_rows] source row, or nullptr: identity
  const int* tok;        // [n_rows] token of row r
  const uint8_t* src[kGatherMaxPieces];
  uint8_t* dst[kGatherMaxPieces];
  int64_t src_stride[kGatherMaxPieces], dst_stride[kGatherMaxPieces];   // bytes
  uint8_t by_tok[kGatherMaxPieces];                                     // 1: source row = tok[r], else parent[r]
};
// appends the pieces of one segment (`bytes` per row); false when it is not 512-byte granular / aligned or does not fit
inline bool row_gather_add(RowGather& rg, const void* src, int64_t src_stride, void* dst, int64_t dst_stride, size_t bytes, bool by_tok) {
  if (bytes == 0 || bytes % 512 != 0 || src_stride % 16 != 0 || dst_stride % 16 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(dst) & 15) != 0 || rg.n_pieces + (int)(bytes / 512) > kGatherMaxPieces)
    return false;
  for (size_t off = 0; off < bytes; off += 512) {
    const int j = rg.n_pieces++;
    rg.src[j] = static_cast<const uint8_t*>(src) + off; rg.src_stride[j] = src_stride;
    rg.dst[j] = static_cast<uint8_t*>(dst) + off; rg.dst_stride[j] = dst_stride;
    rg.by_tok[j] = by_tok ? 1 : 0;
  }
  return true;
}
__device__ __forceinline__ uint4 ldcg128(const void* p) {     // L2-coherent: the rows were written by earlier kernels of the loop
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void stcg128(void* p, const uint4& v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// T_weights: number of frames when the attention weights are requested (raw scores are kept per video), else 0
inline size_t attn_ws_smem_bytes(int K, int D, int H, int NG, int NCG, int T_weights) {
  const int Tp = (T_weights + 15) & ~15;
  return (size_t)D * 2 + (size_t)NG * 4 * 2 * K * 128 * 2 + (size_t)kWsPartSlots * 4 * K * 16 * 4 + (size_t)NCG * K * H * 2 +
         (size_t)kWsEncSlots * 16 * (H + kEncPad) * 2 + (size_t)(2 * kWsPartSlots + 2 * kWsEncSlots) * 8 +
         (T_weights > 0 ? (size_t)NCG * (K * Tp * 4 + K * 8) : 0);
}

// K: exact beam count (<= 8).  Requires D % 32 == 0, D <= 512, H % 64 == 0, H <= 512.
template <int K, int NG, int NCG>
__global__ void __launch_bounds__(32 * (4 * NG + 4 * NCG + 1 + kGatherWarps), 1) attn_additive_ws_kernel(const AttnAddArgs a, const RowGather rg) {
  constexpr int kScoreWarps = 4 * NG, kCtxWarps = 4 * NCG;
  constexpr int kThreads = 32 * (kScoreWarps + kCtxWarps + 1 + kGatherWarps);      // + enc tile producer warp + row gather warps
  extern __shared__ __align__(16) uint8_t smem_u8[];
  const int T = a.T, D = a.D, H = a.H, B = a.B;
  const int NT = (T + 15) >> 4;                  // 16-frame tiles
  const int pitch = H + kEncPad;                 // staged enc row pitch (elements)
  __half* v_s = reinterpret_cast<__half*>(smem_u8);                              // [D]
  __half* q_s = v_s + D;                                                         // [scoring warp][2][K][4 blocks][32]
  float* part = reinterpret_cast<float*>(q_s + (size_t)kScoreWarps * 2 * K * 128); // [slots][4 warps][K][16]
  bf16* out_s = reinterpret_cast<bf16*>(part + kWsPartSlots * 4 * K * 16);       // [NCG][K][H]
  bf16* enc_s = out_s + (size_t)NCG * K * H;                                     // [slots][16][pitch]
  uint64_t* bars = reinterpret_cast<uint64_t*>(enc_s + (size_t)kWsEncSlots * 16 * pitch);
  const uint32_t part_full = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t part_empty = part_full + 8u * kWsPartSlots;
  const uint32_t enc_full = part_empty + 8u * kWsPartSlots;
  const uint32_t enc_empty = enc_full + 8u * kWsEncSlots;
  float2* ml_all = reinterpret_cast<float2*>(bars + 2 * kWsPartSlots + 2 * kWsEncSlots);   // [NCG][K] (max, 1/sum)      } only when the
  float* sc_all = reinterpret_cast<float*>(ml_all + NCG * K);                              // [NCG][K][NT*16] raw scores } weights are requested
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int nvid = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // videos of this CTA
  const int total = nvid * NT;                                                    // units of this CTA

  if (tid == 0) {
    for (int i = 0; i < kWsPartSlots; ++i) { amb_init(part_full + 8u * i, 4); amb_init(part_empty + 8u * i, 4); }
    for (int i = 0; i < kWsEncSlots; ++i) { amb_init(enc_full + 8u * i, 1); amb_init(enc_empty + 8u * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < D / 8; i += kThreads) reinterpret_cast<uint4*>(v_s)[i] = ldg128(a.v + (int64_t)i * 8);
  __syncthreads();

  // enc tile `ft` of video index `vi` -> ring slot
  auto issue_enc = [&](int vi, int ft, int slot) {
    const uint32_t fb = enc_full + 8u * slot;
    amb_expect_tx(fb, 16u * (uint32_t)H * 2u);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(enc_s + (size_t)slot * 16 * pitch);
    const bf16* src = a.values + (int64_t)((int)blockIdx.x + vi * (int)gridDim.x) * T * H;
    for (int rr = 0; rr < 16; ++rr) {
      int t = ft * 16 + rr;
      t = t < T ? t : T - 1;                     // frames past T: any finite row (their weights are zero)
      bulk_g2s(dst + (uint32_t)(rr * pitch) * 2u, src + (int64_t)t * H, (uint32_t)H * 2u, fb);
    }
  };
  // producer state (lane 0 of the last warp): the first ring pass does not depend on the previous kernel (PDL, common.cuh)
  int p_u = 0, p_vi = 0, p_ft = 0;
  if (warp == kScoreWarps + kCtxWarps && lane == 0) {
    for (; p_u < total && p_u < kWsEncSlots; ++p_u) {
      issue_enc(p_vi, p_ft, p_u);
      if (++p_ft == NT) { p_ft = 0; ++p_vi; }
    }
  }
  pdl_wait();
  pdl_launch_dependents();

  if (warp < kScoreWarps) {
    // ================= scoring warps
    const int gi = warp >> 2, sw = warp & 3;
    const int nfb = D / 32;
    __half* q_w = q_s + (size_t)warp * 2 * K * 128;
    // this warp's query columns (feature blocks sw, sw+4, ...) of video b -> private buffer
    auto load_q = [&](int b, int buf) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(q_w + (size_t)buf * K * 128);
      for (int i = lane; i < K * 16; i += 32) {  // i = (k*4 + j)*4 + c: 16-byte chunk c of block j of beam k
        const int c = i & 3, j = (i >> 2) & 3, k = i >> 4;
        const int fb = sw + 4 * j;
        const int64_t qr = a.q_rows ? (int64_t)a.q_rows[b * K + k] : (int64_t)b * K + k;
        if (fb < nfb) cp_async16(dst + (uint32_t)i * 16u, a.q + qr * D + fb * 32 + c * 8);
      }
      cp_async_commit();
    };
    int u = gi, vi = gi / NT, ft = gi - (gi / NT) * NT;
    int pslot = gi;
    uint32_t ppar = 0;
    ring_adv(pslot, ppar, 0, kWsPartSlots);
    int qb = 1, q_vi = -1;
    uint4 key0 = make_uint4(0u, 0u, 0u, 0u), key1 = key0;
    if (u < total) {
      const int b = (int)blockIdx.x + vi * (int)gridDim.x;
      load_q(b, 0);
      if (sw < nfb) {
        int t0 = ft * 16 + g, t1 = t0 + 8;
        t0 = t0 < T ? t0 : T - 1;
        t1 = t1 < T ? t1 : T - 1;
        const __half* kb = a.keys + (int64_t)b * T * D + tg * 8 + sw * 32;
        key0 = ldg128(kb + (int64_t)t0 * D);
        key1 = ldg128(kb + (int64_t)t1 * D);
      }
    }
    while (u < total) {
      const int b = (int)blockIdx.x + vi * (int)gridDim.x;
      ATTN_PROBE(sw == 0 && lane == 0, u, 0);
      // next unit of this group
      int vin = vi, ftn = ft + NG;
      while (ftn >= NT) { ftn -= NT; ++vin; }
      const bool has_next = u + NG < total;
      if (vi != q_vi) {                          // queries of this video were prefetched into the other buffer
        cp_async_wait<0>();
        __syncwarp();
        qb ^= 1;
        q_vi = vi;
      }
      if (has_next && vin != vi) load_q((int)blockIdx.x + vin * (int)gridDim.x, qb ^ 1);
      if (sw == 0 && lane == 0 && u + 2 * NG < total) {
        // the key tile of this group's unit after the next one -> L2 (one contiguous block of 16 frames): the
        // register prefetch below then only has to cover the L2 latency, not HBM's
        int vp = vin, fp = ftn + NG;
        while (fp >= NT) { fp -= NT; ++vp; }
        const int rows = T - fp * 16 < 16 ? T - fp * 16 : 16;
        bulk_prefetch_l2(a.keys + ((int64_t)((int)blockIdx.x + vp * (int)gridDim.x) * T + fp * 16) * D, (uint32_t)(rows * D * 2));
      }
      const __half* q = q_w + (size_t)qb * K * 128 + tg * 8;
      int t0 = ft * 16 + g, t1 = t0 + 8;
      t0 = t0 < T ? t0 : T - 1;
      t1 = t1 < T ? t1 : T - 1;
      const __half* kbase = a.keys + (int64_t)b * T * D + tg * 8;
      const __half* k0p = kbase + (int64_t)t0 * D;
      const __half* k1p = kbase + (int64_t)t1 * D;
      int n0 = ftn * 16 + g, n1 = n0 + 8;        // same lanes, first block of the next unit
      n0 = n0 < T ? n0 : T - 1;
      n1 = n1 < T ? n1 : T - 1;
      const __half* nbase = a.keys + (int64_t)((int)blockIdx.x + vin * (int)gridDim.x) * T * D + tg * 8 + sw * 32;
      float acc[K][4];
#pragma unroll
      for (int k = 0; k < K; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }
      // (D <= 512: at most four feature blocks per warp; unrolled with an early exit -- 81.5 -> 77.5 us per launch at D = 512.
      // Compile-time D / H variants of this kernel measured no further gain, unlike the dot-product kernel's consumers.)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int fb = sw + 4 * j;
        if (fb >= nfb) break;
        uint4 nx0 = make_uint4(0u, 0u, 0u, 0u), nx1 = nx0;
        const uint4 v4 = *reinterpret_cast<const uint4*>(v_s + fb * 32 + tg * 8);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const uint4 q4 = *reinterpret_cast<const uint4*>(q + k * 128 + j * 32);
          mma_f16(acc[k], tanh_h2(key0.x, q4.x), tanh_h2(key1.x, q4.x), tanh_h2(key0.y, q4.y), tanh_h2(key1.y, q4.y), v4.x, v4.y);
          mma_f16(acc[k], tanh_h2(key0.z, q4.z), tanh_h2(key1.z, q4.z), tanh_h2(key0.w, q4.w), tanh_h2(key1.w, q4.w), v4.z, v4.w);
          if (k == 0) {
            // The next block's keys are requested only AFTER this block's key registers have been read for the first time
            // (asm volatile loads between the first beam's MMAs and the second's).  Issued in front of the block, ptxas gave
            // them the scoreboard of the loads this block is about to consume, so the first HADD2 of every other block waited
            // for the loads just issued as well (16% of the scoring warps' stall samples, ncu source page of the r3 build).
            if (fb + 4 < nfb) {
              nx0 = ldg128_pinned(k0p + (fb + 4) * 32);
              nx1 = ldg128_pinned(k1p + (fb + 4) * 32);
            } else if (has_next) {
              nx0 = ldg128_pinned(nbase + (int64_t)n0 * D);
              nx1 = ldg128_pinned(nbase + (int64_t)n1 * D);
            }
          }
        }
        key0 = nx0;
        key1 = nx1;
      }
      ATTN_PROBE(sw == 0 && lane == 0, u, 1);
      amb_wait(part_empty + 8u * pslot, ppar ^ 1u);
      ATTN_PROBE(sw == 0 && lane == 0, u, 2);
      if (tg == 0) {
        float* pp = part + (size_t)(pslot * 4 + sw) * K * 16;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          pp[k * 16 + g] = acc[k][0];            // every column of the accumulator tile holds the dot product
          pp[k * 16 + g + 8] = acc[k][2];
        }
      }
      __syncwarp();
      if (lane == 0) amb_arrive(part_full + 8u * pslot);
      ring_adv(pslot, ppar, NG, kWsPartSlots);
      u += NG;
      vi = vin;
      ft = ftn;
    }
  } else if (warp < kScoreWarps + kCtxWarps) {
    // ================= context warps
    const int cg = (warp - kScoreWarps) >> 2, cw = (warp - kScoreWarps) & 3, ctid = tid - 32 * kScoreWarps - 128 * cg;
    const int cwid = H / 4;                      // columns per warp
    const int nmt = cwid / 16;                   // m-tiles per warp (8 at H = 512)
    constexpr int MAXMT = 8;
    constexpr float kL2e = 1.4426950408889634f;
    bf16* outg = out_s + (size_t)cg * K * H;
    float2* ml_s = ml_all + cg * K;
    float* sc_s = sc_all + (size_t)cg * K * NT * 16;
    int pslot = 0, eslot = 0;
    uint32_t ppar = 0, epar = 0;
    ring_adv(pslot, ppar, cg * NT, kWsPartSlots);
    ring_adv(eslot, epar, cg * NT, kWsEncSlots);
    const int lrow = ((lane >> 4) << 3) + (lane & 7), lcol = ((lane >> 3) & 1) << 3;   // ldmatrix row address of this lane
    for (int vi = cg; vi < nvid; vi += NCG) {
      const int b = (int)blockIdx.x + vi * (int)gridDim.x;
      float c[MAXMT][4];
#pragma unroll
      for (int i = 0; i < MAXMT; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
      float m_run = -1e30f, l_run = 0.f;
      for (int ft = 0; ft < NT; ++ft) {
        // scores of beam g for frames ft*16 + {2tg, 2tg+1, 2tg+8, 2tg+9}
        ATTN_PROBE(cw == 0 && lane == 0, vi * NT + ft, 4);
        amb_wait(part_full + 8u * pslot, ppar);
        ATTN_PROBE(cw == 0 && lane == 0, vi * NT + ft, 5);
        float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
        if (g < K) {
          const float* pp = part + (size_t)pslot * 4 * K * 16 + g * 16 + 2 * tg;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float2 lo = *reinterpret_cast<const float2*>(pp + w * K * 16);
            const float2 hi = *reinterpret_cast<const float2*>(pp + w * K * 16 + 8);
            x0 += lo.x; x1 += lo.y; x2 += hi.x; x3 += hi.y;
          }
        }
        __syncwarp();
        if (lane == 0) amb_arrive(part_empty + 8u * pslot);
        ring_adv(pslot, ppar, 1, kWsPartSlots);
        // bias, mask (attention.py:61), frames past T
        const int tb = ft * 16 + 2 * tg;
        float xs[4] = {x0 + a.v_bias, x1 + a.v_bias, x2 + a.v_bias, x3 + a.v_bias};
        const int ts[4] = {tb, tb + 1, tb + 8, tb + 9};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (ts[i] >= T) xs[i] = -INFINITY;
          else if (a.mask != nullptr && a.mask[(int64_t)b * T + ts[i]] == 0.f) xs[i] = -1e9f;
        }
        if (a.attn_out != nullptr && cw == 0 && g < K) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sc_s[g * NT * 16 + ts[i]] = xs[i];
        }
        // online softmax over the tile (the 4 lanes of a group hold the 16 frames of beam g)
        float tmax = fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3]));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
        const float m_new = fmaxf(m_run, tmax);
        float scale = exp2f((m_run - m_new) * kL2e);
        float pv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pv[i] = exp2f((xs[i] - m_new) * kL2e);
        float psum = (pv[0] + pv[1]) + (pv[2] + pv[3]);
        psum += __shfl_xor_sync(0xffffffffu, psum, 1);
        psum += __shfl_xor_sync(0xffffffffu, psum, 2);
        l_run = l_run * scale + psum;
        m_run = m_new;
        if (g >= K) { scale = 1.f; pv[0] = pv[1] = pv[2] = pv[3] = 0.f; }
        // accumulator columns of this lane are beams 2tg, 2tg+1: their rescale factors live in lanes 8tg, 8tg+4
        const float sc0 = __shfl_sync(0xffffffffu, scale, 8 * tg);
        const float sc1 = __shfl_sync(0xffffffffu, scale, 8 * tg + 4);
        if (ft > 0) {
#pragma unroll
          for (int i = 0; i < MAXMT; ++i) { c[i][0] *= sc0; c[i][1] *= sc1; c[i][2] *= sc0; c[i][3] *= sc1; }
        }
        uint32_t bh0, bl0, bh1, bl1;
        split_bf16x2(pv[0], pv[1], bh0, bl0);
        split_bf16x2(pv[2], pv[3], bh1, bl1);
        // context: ctx^T[col, beam] += enc[t, col] * p[beam, t]
        amb_wait(enc_full + 8u * eslot, epar);
        ATTN_PROBE(cw == 0 && lane == 0, vi * NT + ft, 6);
        const uint32_t tile = (uint32_t)__cvta_generic_to_shared(enc_s + (size_t)eslot * 16 * pitch);
#pragma unroll
        for (int i = 0; i < MAXMT; ++i) {
          if (i < nmt) {
            const int col0 = cw * cwid + i * 16;
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4_trans(tile + (uint32_t)(lrow * pitch + col0 + lcol) * 2u, a0, a1, a2, a3);
            mma_bf16(c[i], a0, a1, a2, a3, bh0, bh1);
            mma_bf16(c[i], a0, a1, a2, a3, bl0, bl1);
          }
        }
        __syncwarp();
        ATTN_PROBE(cw == 0 && lane == 0, vi * NT + ft, 7);
        if (lane == 0) amb_arrive(enc_empty + 8u * eslot);
        ring_adv(eslot, epar, 1, kWsEncSlots);
      }
      ring_adv(pslot, ppar, (NCG - 1) * NT, kWsPartSlots);      // the units of the other groups' videos
      ring_adv(eslot, epar, (NCG - 1) * NT, kWsEncSlots);
      // normalise and store: c[i] = {ctx[col0+g][2tg], ctx[col0+g][2tg+1], ctx[col0+g+8][2tg], ctx[col0+g+8][2tg+1]} (col, beam)
      const float inv = (g < K) ? 1.0f / l_run : 0.f;
      const float i0 = __shfl_sync(0xffffffffu, inv, 8 * tg);
      const float i1 = __shfl_sync(0xffffffffu, inv, 8 * tg + 4);
      named_bar_sync(1 + cg, 128);               // the previous video's rows have left the staging buffer
      if (a.attn_out != nullptr && cw == 0 && g < K && tg == 0) ml_s[g] = make_float2(m_run, inv);
      const int k0 = 2 * tg, k1 = 2 * tg + 1;
#pragma unroll
      for (int i = 0; i < MAXMT; ++i) {
        if (i < nmt) {
          const int col0 = cw * cwid + i * 16;
          if (k0 < K) { outg[k0 * H + col0 + g] = __float2bfloat16_rn(c[i][0] * i0); outg[k0 * H + col0 + g + 8] = __float2bfloat16_rn(c[i][2] * i0); }
          if (k1 < K) { outg[k1 * H + col0 + g] = __float2bfloat16_rn(c[i][1] * i1); outg[k1 * H + col0 + g + 8] = __float2bfloat16_rn(c[i][3] * i1); }
        }
      }
      named_bar_sync(1 + cg, 128);
#pragma unroll
      for (int k = 0; k < K; ++k)
        for (int c8 = ctid * 8; c8 < H; c8 += 128 * 8)
          *reinterpret_cast<uint4*>(a.ctx + ((int64_t)b * K + k) * a.ctx_ld + c8) = *reinterpret_cast<const uint4*>(outg + k * H + c8);
      if (a.attn_out != nullptr) {               // attention weights (attention.py:64) from the raw scores and the final (max, sum)
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float2 ml = ml_s[k];
          for (int t = ctid; t < T; t += 128)
            a.attn_out[((int64_t)b * K + k) * a.attn_ld + t] = exp2f((sc_s[k * NT * 16 + t] - ml.x) * kL2e) * ml.y;
        }
        named_bar_sync(1 + cg, 128);             // before the next video's scores overwrite sc_s
      }
    }
  } else if (warp == kScoreWarps + kCtxWarps) {
    if (lane == 0) {
      // ================= enc tile producer: as far ahead as the ring allows
      int slot = 0;
      uint32_t par = 0;
      ring_adv(slot, par, p_u, kWsEncSlots);
      for (; p_u < total; ++p_u) {
        amb_wait(enc_empty + 8u * slot, par ^ 1u);
        issue_enc(p_vi, p_ft, slot);
        if (++p_ft == NT) { p_ft = 0; ++p_vi; }
        ring_adv(slot, par, 1, kWsEncSlots);
      }
    }
  } else if (rg.n_rows > 0) {
    // ================= row gather (RowGather above): warp gw takes rows blockIdx.x + (gw + 2 i) gridDim.x; a row moves in two
    // batches of up to eight 512-byte pieces (all loads of a batch in flight, then its stores)
    const int gw = warp - (kScoreWarps + kCtxWarps + 1);
    const int n_my = (rg.n_rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    for (int i = gw; i < n_my; i += kGatherWarps) {
      const int r = (int)blockIdx.x + i * (int)gridDim.x;
      const int64_t p = rg.parent != nullptr ? (int64_t)__ldcg(rg.parent + r) : (int64_t)r;     // written by the previous kernel
      int tok = __ldcg(rg.tok + r);
      tok = min(max(tok, 0), rg.V - 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pc = half * 8 + j;
          v[j] = make_uint4(0u, 0u, 0u, 0u);
          if (pc < rg.n_pieces && !(rg.dbg & 1))
            v[j] = ldcg128(rg.src[pc] + (rg.by_tok[pc] ? (int64_t)tok : p) * rg.src_stride[pc] + lane * 16);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pc = half * 8 + j;
          if (pc < rg.n_pieces && !(rg.dbg & 2)) stcg128(rg.dst[pc] + (int64_t)r * rg.dst_stride[pc] + lane * 16, v[j]);
        }
      }
    }
  }
}

inline int attn_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
inline int attn_ws_groups() {          // VC_ATTN_GROUPS=NG*10+NCG: scoring / context groups per CTA (A/B testing); default 31
  // (three scoring groups since the CTA also carries the two row-gather warps: 1.63 vs 1.67 ms per 20 launches at C2; before
  // that 41 and 31 measured the same)
  const char* e = getenv("VC_ATTN_GROUPS");
  const int v = e != nullptr ? atoi(e) : 31;
  return (v == 21 || v == 31 || v == 41 || v == 42 || v == 51 || v == 52) ? v : 31;
}
// The persistent kernel needs at least two videos per SM to beat v4 (one CTA per video); VC_ATTN_WS_MIN_B overrides the
// threshold (tests run it on small batches)
inline bool attn_additive_ws_ok(int B, int K, int D, int H, int T, bool weights) {
  const int cfg = attn_ws_groups();
  const char* e = getenv("VC_ATTN_WS_MIN_B");
  const int min_b = e != nullptr ? atoi(e) : 2 * attn_num_sms();
  return B >= min_b && K >= 1 && K <= 8 && D % 32 == 0 && D <= 512 && H % 64 == 0 && H <= 512 && T >= 1 &&
         attn_ws_smem_bytes(K, D, H, cfg / 10, cfg % 10, weights ? T : 0) <= 200 * 1024;
}
template <int NG, int NCG>
int launch_attn_additive_ws_cfg(const AttnAddArgs& a, int K, cudaStream_t stream, const RowGather& rg) {
  const size_t smem = attn_ws_smem_bytes(K, a.D, a.H, NG, NCG, a.attn_out != nullptr ? a.T : 0);
  const int grid = a.B < attn_num_sms() ? a.B : attn_num_sms();
  constexpr int kThreads = 32 * (4 * NG + 4 * NCG + 1 + kGatherWarps);
#define VC_WS_LAUNCH(KK)                                                                               \
  do {                                                                                                 \
    auto kern = attn_additive_ws_kernel<KK, NG, NCG>;                                                  \
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    VC_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), smem, stream, a, rg));                        \
  } while (0)
  switch (K) {
    case 1: VC_WS_LAUNCH(1); break;
    case 2: VC_WS_LAUNCH(2); break;
    case 3: VC_WS_LAUNCH(3); break;
    case 4: VC_WS_LAUNCH(4); break;
    case 5: VC_WS_LAUNCH(5); break;
    case 6: VC_WS_LAUNCH(6); break;
    case 7: VC_WS_LAUNCH(7); break;
    default: VC_WS_LAUNCH(8); break;
  }
#undef VC_WS_LAUNCH
  return VC_OK;
}
inline int launch_attn_additive_ws(const AttnAddArgs& a, int K, cudaStream_t stream, const RowGather& rg = RowGather()) {
  VC_CHECK(attn_additive_ws_ok(a.B, K, a.D, a.H, a.T, a.attn_out != nullptr), "additive attention (ws): B=%d K=%d D=%d H=%d T=%d not supported",
           a.B, K, a.D, a.H, a.T);
  VC_CHECK(a.ctx_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a.ctx) & 15) == 0, "additive attention (ws): ctx must be 16-byte aligned");
  switch (attn_ws_groups()) {
    case 21: return launch_attn_additive_ws_cfg<2, 1>(a, K, stream, rg);
    case 41: return launch_attn_additive_ws_cfg<4, 1>(a, K, stream, rg);
    case 42: return launch_attn_additive_ws_cfg<4, 2>(a, K, stream, rg);
    case 51: return launch_attn_additive_ws_cfg<5, 1>(a, K, stream, rg);
    case 52: return launch_attn_additive_ws_cfg<5, 2>(a, K, stream, rg);
    default: return launch_attn_additive_ws_cfg<3, 1>(a, K, stream, rg);
  }
}

inline int attn_mma_nbuf() {           // VC_ATTN_NBUF=2: double-buffered enc staging, 4 CTAs/SM; 1 (default): single buffer, 7 CTAs/SM
  static int n = 0;
  if (n == 0) { const char* e = getenv("VC_ATTN_NBUF"); n = (e != nullptr && e[0] == '2') ? 2 : 1; }
  return n;
}
inline bool attn_additive_mma_ok(int K, int D, int H, int T) {
  const int Tp = (T + 15) & ~15;
  const size_t smem = (size_t)(K + 1) * D * 2 + (size_t)5 * K * Tp * 4 + (size_t)attn_mma_nbuf() * 16 * (H + kEncPad) * 2;
  return K >= 1 && K <= 8 && D % 32 == 0 && H % 64 == 0 && H <= 512 && smem <= 56 * 1024;
}

inline int launch_attn_additive_mma(const AttnAddArgs& a, int K, cudaStream_t stream) {
  VC_CHECK(attn_additive_mma_ok(K, a.D, a.H, a.T), "additive attention (mma): K=%d D=%d H=%d T=%d not supported", K, a.D, a.H, a.T);
  VC_CHECK(a.ctx_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a.ctx) & 15) == 0, "additive attention (mma): ctx must be 16-byte aligned");
  const int Tp = (a.T + 15) & ~15;
  const int nbuf = attn_mma_nbuf();
  const size_t smem = (size_t)(K + 1) * a.D * 2 + (size_t)5 * K * Tp * 4 + (size_t)nbuf * 16 * (a.H + kEncPad) * 2;
#define VC_MMA_LAUNCH(KK)                                                                                  \
  do {                                                                                                     \
    if (nbuf == 2) {                                                                                       \
      auto kern = attn_additive_mma_kernel<KK, 2, 4>;                                                      \
      VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
      VC_CUDA(launch_pdl(kern, dim3(a.B), dim3(kMmaThreads), smem, stream, a));                            \
    } else {                                                                                               \
      auto kern = attn_additive_mma_kernel<KK, 1, 7>;                                                      \
      VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
      VC_CUDA(launch_pdl(kern, dim3(a.B), dim3(kMmaThreads), smem, stream, a));                            \
    }                                                                                                      \
  } while (0)
  switch (K) {
    case 1: VC_MMA_LAUNCH(1); break;
    case 2: VC_MMA_LAUNCH(2); break;
    case 3: VC_MMA_LAUNCH(3); break;
    case 4: VC_MMA_LAUNCH(4); break;
    case 5: VC_MMA_LAUNCH(5); break;
    case 6: VC_MMA_LAUNCH(6); break;
    case 7: VC_MMA_LAUNCH(7); break;
    default: VC_MMA_LAUNCH(8); break;
  }
#undef VC_MMA_LAUNCH
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

inline bool attn_additive_fast_ok(int K, int D, int H) {
  return (K == 1 || K == 3 || K == 5) && D % 8 == 0 && D <= 512 && H % 4 == 0;
}

inline int launch_attn_additive(const AttnAddArgs& a, int K, cudaStream_t stream) {
  VC_CHECK(attn_additive_fast_ok(K, a.D, a.H), "additive attention fast path: K=%d D=%d H=%d not supported", K, a.D, a.H);
  const int DH = a.D > 256 ? 2 : 1;
  const int Tp = (a.T + 3) & ~3;
  const size_t smem = sizeof(float) * (size_t)(DH + 1) * K * Tp;
  VC_CHECK(smem <= 48 * 1024, "additive attention: K=%d T=%d needs %zu B shared memory", K, a.T, smem);
#define VC_ADD_LAUNCH(KK, DD) attn_additive_kernel<KK, DD><<<a.B, kAddThreads, smem, stream>>>(a)
  if (DH == 2) {
    if (K == 1) VC_ADD_LAUNCH(1, 2);
    else if (K == 3) VC_ADD_LAUNCH(3, 2);
    else VC_ADD_LAUNCH(5, 2);
  } else {
    if (K == 1) VC_ADD_LAUNCH(1, 1);
    else if (K == 3) VC_ADD_LAUNCH(3, 1);
    else VC_ADD_LAUNCH(5, 1);
  }
#undef VC_ADD_LAUNCH
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// KT: element type of a.skeys (see AttnArgs)
template <class T, class KT, int MODE, bool PRECISE>
int launch_attn_step(const AttnArgs<T>& a, cudaStream_t stream) {
  VC_CHECK(a.K >= 1 && a.K <= 16, "attention: beam size %d not in [1,16]", a.K);
  VC_CHECK(a.H % 8 == 0 && a.D % 8 == 0, "attention: dims must be multiples of 8 (D=%d H=%d)", a.D, a.H);
  if (MODE == ATTN_MHA) {
    const int dh = a.heads > 0 ? a.D / a.heads : 0;
    VC_CHECK(a.heads >= 1 && a.D % a.heads == 0 && dh % 8 == 0 && (dh & (dh - 1)) == 0 && a.H == a.D,
             "multi-head attention: head dim %d must be a power of two >= 8 (heads=%d dim=%d)", dh, a.heads, a.D);
  }
  const int NH = (MODE == ATTN_MHA) ? a.heads : 1;
  const int cols8 = a.H / 8;
  const int G = kAttnThreads / cols8 > 1 ? kAttnThreads / cols8 : 1;
  const size_t smem = sizeof(float) * ((size_t)a.K * a.D + (MODE == ATTN_ADDITIVE ? a.D : 0) + (size_t)a.K * NH * a.T_ +
                                       (G > 1 ? (size_t)G * a.K * a.H : 0));
  VC_CHECK(smem <= 200 * 1024, "attention: K=%d D=%d T=%d needs %zu B shared memory", a.K, a.D, a.T_, smem);
#define VC_ATTN_LAUNCH(KM, NC)                                                                           \
  do {                                                                                                   \
    auto kern = attn_step_kernel<T, KT, MODE, KM, PRECISE, NC>;                                          \
    if (smem > 48 * 1024) VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<a.B, kAttnThreads, smem, stream>>>(a);                                                        \
  } while (0)
  constexpr bool kPacked = (MODE == ATTN_ADDITIVE) && std::is_same<KT, __half>::value;
  bool done = false;
  if constexpr (kPacked) {
    // register-resident queries: exact K in {1,3,5}, D <= 512
    if (a.D <= 256) {
      if (a.K == 1) { VC_ATTN_LAUNCH(1, 1); done = true; }
      else if (a.K == 3) { VC_ATTN_LAUNCH(3, 1); done = true; }
      else if (a.K == 5) { VC_ATTN_LAUNCH(5, 1); done = true; }
    } else if (a.D <= 512) {
      if (a.K == 1) { VC_ATTN_LAUNCH(1, 2); done = true; }
      else if (a.K == 3) { VC_ATTN_LAUNCH(3, 2); done = true; }
      else if (a.K == 5) { VC_ATTN_LAUNCH(5, 2); done = true; }
    }
  }
  if (!done) {
    if (a.K == 1) VC_ATTN_LAUNCH(1, 0);
    else if (a.K <= 3) VC_ATTN_LAUNCH(3, 0);
    else if (a.K <= 5) VC_ATTN_LAUNCH(5, 0);
    else if (a.K <= 8) VC_ATTN_LAUNCH(8, 0);
    else VC_ATTN_LAUNCH(16, 0);
  }
#undef VC_ATTN_LAUNCH
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
