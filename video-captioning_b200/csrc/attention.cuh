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
// from HBM once per video-step (not once per row).  Warps stride over frames t with lanes covering the
// feature dimension in 16/32-byte vectors (coalesced 512B-1KB per warp per frame); per-beam partial
// scores are reduced with warp shuffles.  HBM-bound (plus MUFU-bound for the additive form).
#pragma once
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

// KMAX: compile-time bound on beams handled per CTA (K <= KMAX).
template <class T, int MODE, int KMAX, bool PRECISE>
__global__ void __launch_bounds__(256) attn_step_kernel(const AttnArgs<T> a) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x;
  const int K = a.K, Tn = a.T_, D = a.D, H = a.H;
  const int NH = (MODE == ATTN_MHA) ? a.heads : 1;
  float* q_s = smem;                       // [K][D]
  float* v_s = q_s + (size_t)K * D;        // [D] (additive only)
  float* sc = v_s + ((MODE == ATTN_ADDITIVE) ? D : 0);   // [K][NH][Tn]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp = blockDim.x >> 5;

  for (int i = tid; i < K * D; i += blockDim.x) {
    const int k = i / D, d = i - k * D;
    const int64_t r = (int64_t)b * K + k;
    q_s[i] = a.q ? a.q[r * D + d] : to_float(a.q_act[r * a.q_ld + d]);
  }
  if (MODE == ATTN_ADDITIVE)
    for (int i = tid; i < D; i += blockDim.x) v_s[i] = a.v[i];
  __syncthreads();

  // ---- scores: warp per frame, lanes over D in vectors of 8
  const T* sk = a.skeys + (int64_t)b * Tn * D;
  const int group = (MODE == ATTN_MHA) ? (32 / NH) : 32;   // lanes that reduce together
  for (int t = warp; t < Tn; t += nwarp) {
    float part[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) part[k] = 0.f;
    if (MODE == ATTN_MHA) {
      // each lane owns the contiguous D/32 columns [lane*c, (lane+1)*c): all inside one head
      const int c = D >> 5;
      for (int d0 = lane * c; d0 < (lane + 1) * c; d0 += 4) {
        float e[4];
        load4(sk + (int64_t)t * D + d0, e);
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) {
            const float* qk = q_s + k * D + d0;
            part[k] += e[0] * qk[0] + e[1] * qk[1] + e[2] * qk[2] + e[3] * qk[3];
          }
      }
    } else {
      for (int d0 = lane * 8; d0 < D; d0 += 256) {
        float e[8];
        load8(sk + (int64_t)t * D + d0, e);
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) {
            const float* qk = q_s + k * D + d0;
            float s = 0.f;
            if (MODE == ATTN_ADDITIVE) {
#pragma unroll
              for (int j = 0; j < 8; ++j) s = fmaf(v_s[d0 + j], tanh_<PRECISE>(e[j] + qk[j]), s);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) s = fmaf(e[j], qk[j], s);
            }
            part[k] += s;
          }
      }
    }
    const bool masked = a.mask != nullptr && a.mask[(int64_t)b * Tn + t] == 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        float s = part[k];
        for (int o = group >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((lane % group) == 0) {
          const int hd = lane / group;
          float val = (MODE == ATTN_ADDITIVE) ? s + a.v_bias : ((MODE == ATTN_MHA) ? s * a.scale : s);
          if (masked) val = -1e9f;
          sc[((size_t)k * NH + hd) * Tn + t] = val;
        }
      }
  }
  __syncthreads();

  // ---- softmax over T per (beam, head): one warp per row of sc
  for (int row = warp; row < K * NH; row += nwarp) {
    float* s = sc + (size_t)row * Tn;
    float m = -INFINITY;
    for (int t = lane; t < Tn; t += 32) m = fmaxf(m, s[t]);
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
    for (int i = tid; i < K * Tn; i += blockDim.x) {
      const int k = i / Tn, t = i - k * Tn;
      float w = 0.f;
      for (int hd = 0; hd < NH; ++hd) w += sc[((size_t)k * NH + hd) * Tn + t];
      if (NH > 1) w /= (float)NH;
      a.attn_out[((int64_t)b * K + k) * a.attn_ld + t] = w;
    }
  }

  // ---- context: threads over H in vectors of 4, loop over frames (values read once per video)
  const T* vv = a.values + (int64_t)b * Tn * H;
  const int dh = (MODE == ATTN_MHA) ? (H / NH) : H;
  for (int h0 = tid * 4; h0 < H; h0 += blockDim.x * 4) {
    float acc[KMAX][4];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
    const int hd = (MODE == ATTN_MHA) ? (h0 / dh) : 0;
    for (int t = 0; t < Tn; ++t) {
      float e[4];
      load4(vv + (int64_t)t * H + h0, e);
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) {
          const float w = sc[((size_t)k * NH + hd) * Tn + t];
          acc[k][0] = fmaf(w, e[0], acc[k][0]);
          acc[k][1] = fmaf(w, e[1], acc[k][1]);
          acc[k][2] = fmaf(w, e[2], acc[k][2]);
          acc[k][3] = fmaf(w, e[3], acc[k][3]);
        }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) store4(a.ctx + ((int64_t)b * K + k) * a.ctx_ld + h0, acc[k]);
  }
}

template <class T, int MODE, bool PRECISE>
int launch_attn_step(const AttnArgs<T>& a, cudaStream_t stream) {
  VC_CHECK(a.K >= 1 && a.K <= 16, "attention: beam size %d not in [1,16]", a.K);
  VC_CHECK(a.H % 4 == 0, "attention: hidden dim %d must be a multiple of 4", a.H);
  if (MODE == ATTN_MHA) {
    VC_CHECK(a.heads >= 1 && a.heads <= 32 && (32 % a.heads) == 0 && a.D % 128 == 0,
             "multi-head attention: heads=%d must divide 32 and dim=%d must be a multiple of 128", a.heads, a.D);
  } else {
    VC_CHECK(a.D % 8 == 0, "attention: scoring dim %d must be a multiple of 8", a.D);
  }
  const int NH = (MODE == ATTN_MHA) ? a.heads : 1;
  const size_t smem = sizeof(float) * ((size_t)a.K * a.D + (MODE == ATTN_ADDITIVE ? a.D : 0) + (size_t)a.K * NH * a.T_);
  VC_CHECK(smem <= 200 * 1024, "attention: K=%d D=%d T=%d needs %zu B shared memory", a.K, a.D, a.T_, smem);
#define VC_ATTN_LAUNCH(KM)                                                                               \
  do {                                                                                                   \
    auto kern = attn_step_kernel<T, MODE, KM, PRECISE>;                                                  \
    if (smem > 48 * 1024) VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<a.B, 256, smem, stream>>>(a);                                                                 \
  } while (0)
  if (a.K == 1) VC_ATTN_LAUNCH(1);
  else if (a.K <= 4) VC_ATTN_LAUNCH(4);
  else if (a.K <= 8) VC_ATTN_LAUNCH(8);
  else VC_ATTN_LAUNCH(16);
#undef VC_ATTN_LAUNCH
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
