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
