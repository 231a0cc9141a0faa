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

#include "common.cuh"

namespace vc {

enum AttnMode : int { ATTN_ADDITIVE = 0, ATTN_DOT = 1, ATTN_MHA = 2 };

template <class T>
struct AttnArgs {
  // scoring operand per video: [B,T,D] (additive: projected keys, D=A; dot: enc_out, D=H; mha: K, D=H)
  const T* skeys;
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

constexpr int kAttnThreads = 256;
constexpr int kAttnFR = 2;   // frames per warp iteration

// KMAX: compile-time bound on beams handled per CTA (K <= KMAX).
template <class T, int MODE, int KMAX, bool PRECISE>
__global__ void __launch_bounds__(kAttnThreads) attn_step_kernel(const AttnArgs<T> a) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x;
  const int K = a.K, Tn = a.T_, D = a.D, H = a.H;
  const int NH = (MODE == ATTN_MHA) ? a.heads : 1;
  const int dh = D / NH;                                   // scoring columns per head
  float* q_s = smem;                                       // [K][D]
  float* v_s = q_s + (size_t)K * D;                        // [D] (additive only)
  float* sc = v_s + ((MODE == ATTN_ADDITIVE) ? D : 0);     // [K][NH][Tn]
  float* red = sc + (size_t)K * NH * Tn;                   // [G][K][H] context partials
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarp = kAttnThreads / 32;

  for (int i = tid; i < K * D; i += kAttnThreads) {
    const int k = i / D, d = i - k * D;
    const int64_t r = (int64_t)b * K + k;
    q_s[i] = a.q ? a.q[r * D + d] : to_float(a.q_act[r * a.q_ld + d]);
  }
  if (MODE == ATTN_ADDITIVE)
    for (int i = tid; i < D; i += kAttnThreads) v_s[i] = a.v[i];
  if (MODE == ATTN_MHA)
    for (int i = tid; i < K * NH * Tn; i += kAttnThreads) sc[i] = 0.f;
  __syncthreads();

  // ---- scores
  const T* sk = a.skeys + (int64_t)b * Tn * D;
  const int group = (MODE == ATTN_MHA) ? min(32, dh / 8) : 32;   // lanes reducing together (one head)
  for (int t0 = warp * kAttnFR; t0 < Tn; t0 += nwarp * kAttnFR) {
    float part[kAttnFR][KMAX];
#pragma unroll
    for (int f = 0; f < kAttnFR; ++f)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) part[f][k] = 0.f;
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

template <class T, int MODE, bool PRECISE>
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
#define VC_ATTN_LAUNCH(KM)                                                                               \
  do {                                                                                                   \
    auto kern = attn_step_kernel<T, MODE, KM, PRECISE>;                                                  \
    if (smem > 48 * 1024) VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<a.B, kAttnThreads, smem, stream>>>(a);                                                        \
  } while (0)
  if (a.K == 1) VC_ATTN_LAUNCH(1);
  else if (a.K <= 3) VC_ATTN_LAUNCH(3);
  else if (a.K <= 5) VC_ATTN_LAUNCH(5);
  else if (a.K <= 8) VC_ATTN_LAUNCH(8);
  else VC_ATTN_LAUNCH(16);
#undef VC_ATTN_LAUNCH
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
