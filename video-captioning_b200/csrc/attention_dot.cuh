// Dot-product attention step on a streaming ring (bf16 mode): Luong dot / general (attention.py:118-131, :174-185)
// and multi-head attention (attention.py:245-267), any H that is a multiple of 128 up to 1024.
//
// The generic attn_step_kernel (attention.cuh) reads a video's tile twice with CUDA-core FMAs and per-thread dependent
// loads: 0.15-0.19 of the HBM roofline at the H = 1024 and multi-head configurations.  The step is memory-bound by
// construction -- per video and step it reads T*H (Luong) or 2*T*H (K and V) bf16 values once and does 4*K*T*H flops on them
// (10 flop/byte at K = 5: above what FFMA sustains at HBM speed, far below what the tensor cores do) -- so this kernel is
// organised around the stream, like the additive v5 kernel:
//   * persistent CTAs (two per SM when the ring fits twice), videos b = blockIdx.x + i * gridDim.x;
//   * one producer warp keeps a ring of (video, 16-frame tile) slots full with cp.async.bulk row copies (mbarrier tx),
//     as far ahead as the ring allows -- the first ring pass is issued before the dependency wait (PDL);
//   * eight consumer warps take the tiles in order.  Scores: S[16 frames x 8 beams] = tile[16 x H] . q^T on mma.sync
//     (A = ldmatrix of the tile, B = the video's bf16 queries kept in registers as fragments);
//     online softmax over the tiles (running max / sum, flash-attention style rescaling); context:
//     ctx^T[H x 8 beams] += tile^T . p on mma.sync (A = ldmatrix.trans of the same tile -- or of the V tile --, B = the
//     probabilities as bf16 hi + lo).  Every tile byte is read from shared memory twice and from HBM once.
//   * Single-head forms: warp w owns features / columns [w*H/8, (w+1)*H/8); the eight partial score tiles are summed
//     through a double-buffered 8 KB exchange (one named barrier per tile).  Multi-head: a warp owns whole heads
//     (8 or 16 heads), its score tile is complete and only changes layout inside the warp (8 shuffles per head).
// Frames past T are clamped to the last row by the producer and get -inf scores; masked frames get -1e9 (attention.py:61).
// The attention-weight output (explain_prediction) is not produced here: run_attention keeps the generic kernel for it.
#pragma once
#include "attention.cuh"

namespace vc {

struct AttnDotArgs {
  const bf16* skeys;     // [B, T, H] scoring operand (Luong: enc_out; multi-head: K)
  const bf16* values;    // [B, T, H] value operand (Luong: enc_out = skeys; multi-head: V)
  const bf16* q_act;     // [R, *] bf16 queries, row stride q_ld: the hidden state itself (Luong dot) or the projected query as
  int64_t q_ld;          // the projection GEMM's epilogue rounds it (general / multi-head; like every activation of the bf16 mode)
  const float* mask;     // [B, T] (0 -> masked) or nullptr
  bf16* ctx;             // [R, ctx_ld]
  int64_t ctx_ld;
  int B, K, T, H, heads;
  float scale;           // multi-head: 1/sqrt(d), applied to the scores
  int nslots;            // ring slots (host: launch_attn_dot_ws)
};

// Consumer warps per CTA (+ 1 producer warp).  The consumers are latency-bound (dependent ldmatrix -> mma chains, one softmax
// round and one named barrier per tile): measured with scripts/attn_dot_probe.cu, compute alone, B = 1024, T = 80:
// H = 1024: 4 warps x 16 m-tiles 70 us, 8 warps x 8 m-tiles 102 us (96-register cap at two CTAs per SM: spills);
// H = 512: 8 warps x 4 m-tiles 40 us.  The producer / HBM side alone delivers 4.6-4.9 TB/s (36 us at H = 1024).
inline int attn_dot_warps(int H) { return H > 512 ? 4 : 8; }
constexpr int kDotMaxSlots = 6;

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
// coherent loads of data written by the previous kernel of the stream (read after pdl_wait)
__device__ __forceinline__ uint32_t ld_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}

// CW: consumer warps.  MAXMT: m-tiles (= k-steps) of 16 per warp, H == 16 * CW * MAXMT.  HPW: heads per warp (0: single head,
// partial scores summed across the warps, values = skeys; > 0: multi-head with a separate value tile, head dim = 16 * MAXMT / HPW).
template <int CW, int MAXMT, int HPW>
__global__ void __launch_bounds__(32 * (CW + 1), 2) attn_dot_ws_kernel(const AttnDotArgs a) {
  constexpr int kDotCW = CW;
  constexpr bool SEPV = HPW > 0;
  constexpr int nmt = MAXMT;                       // m-tiles = k-steps of a warp
  extern __shared__ __align__(16) uint8_t smem_u8[];
  constexpr int HP = HPW > 0 ? HPW : 1;            // softmax states per warp
  constexpr int NKH = MAXMT / HP;                  // m-tiles / k-steps per head (HPW > 0)
  constexpr float kL2e = 1.4426950408889634f;
  const int T = a.T, H = a.H, B = a.B, K = a.K;
  const int NT = (T + 15) >> 4;                    // 16-frame tiles
  const int pitch = H + kEncPad;                   // staged row pitch (elements): ldmatrix rows hit distinct banks
  const int nslots = a.nslots;
  float* part = reinterpret_cast<float*>(smem_u8);                       // [2][consumer warps][8 beams][16 frames]
  bf16* out_s = reinterpret_cast<bf16*>(part + 2 * kDotCW * 8 * 16);     // [K][H] output staging
  bf16* ring = out_s + (size_t)K * H;
  const size_t tile_elems = (size_t)16 * pitch;
  const size_t slot_elems = tile_elems * (SEPV ? 2 : 1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)nslots * slot_elems);
  const uint32_t full = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t empty = full + 8u * (uint32_t)nslots;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int nvid = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // videos of this CTA
  const int total = nvid * NT;                                                    // units of this CTA

  if (tid == 0) {
    for (int i = 0; i < nslots; ++i) { amb_init(full + 8u * i, 1); amb_init(empty + 8u * i, kDotCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kDotCW) {
    // ================= producer: tile `ft` of video index `vi` -> ring slot (skeys / values never change inside the decode
    // loop, so nothing here waits for the previous kernel)
    if (lane == 0) {
      auto issue = [&](int vi, int ft, int slot) {
        const uint32_t fb = full + 8u * slot;
#if defined(VC_DOT_PROBE) && VC_DOT_PROBE == 2
        amb_arrive(fb);                            // probe: no loads (consumers compute on whatever the ring holds)
        return;
#endif
        amb_expect_tx(fb, 16u * (uint32_t)H * 2u * (SEPV ? 2u : 1u));
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (size_t)slot * slot_elems);
        const int64_t row0 = (int64_t)((int)blockIdx.x + vi * (int)gridDim.x) * T;
        for (int rr = 0; rr < 16; ++rr) {
          int t = ft * 16 + rr;
          t = t < T ? t : T - 1;                   // frames past T: any finite row (their probabilities are zero)
          bulk_g2s(dst + (uint32_t)(rr * pitch) * 2u, a.skeys + (row0 + t) * H, (uint32_t)H * 2u, fb);
          if (SEPV) bulk_g2s(dst + (uint32_t)(tile_elems + rr * pitch) * 2u, a.values + (row0 + t) * H, (uint32_t)H * 2u, fb);
        }
      };
      int p_u = 0, p_vi = 0, p_ft = 0, slot = 0;
      uint32_t par = 0;
      for (; p_u < total; ++p_u) {
        if (p_u >= nslots) amb_wait(empty + 8u * slot, par ^ 1u);
        issue(p_vi, p_ft, slot);
        if (++p_ft == NT) { p_ft = 0; ++p_vi; }
        ring_adv(slot, par, 1, nslots);
      }
    }
    pdl_launch_dependents();
    return;
  }

  // ================= consumers
  pdl_wait();                                      // queries (and the ctx destination) belong to the previous kernels
  pdl_launch_dependents();
  const int w = warp;
  constexpr int cw = 16 * MAXMT;                   // features = context columns of this warp
  const int lrowA = (((lane >> 3) & 1) << 3) + (lane & 7), lcolA = (lane >> 4) << 3;   // ldmatrix (scores): frames x features
  const int lrowT = ((lane >> 4) << 3) + (lane & 7), lcolT = ((lane >> 3) & 1) << 3;   // ldmatrix.trans (context)
  int slot = 0, u = 0;
  uint32_t par = 0;
  for (int vi = 0; vi < nvid; ++vi) {
    const int b = (int)blockIdx.x + vi * (int)gridDim.x;
    // ---- query fragments of beam g: features k0 + {2tg, 2tg+1} and + 8 of every k-step
    uint32_t qh[MAXMT][2];
#pragma unroll
    for (int i = 0; i < MAXMT; ++i) {
      qh[i][0] = qh[i][1] = 0u;
      if (g < K) {
        const int k0 = w * cw + i * 16 + 2 * tg;
        const int64_t r = (int64_t)b * K + g;
        qh[i][0] = ld_u32(a.q_act + r * a.q_ld + k0);
        qh[i][1] = ld_u32(a.q_act + r * a.q_ld + k0 + 8);
      }
    }
    float c[MAXMT][4];
#pragma unroll
    for (int i = 0; i < MAXMT; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
    float m_run[HP], l_run[HP];
#pragma unroll
    for (int hh = 0; hh < HP; ++hh) { m_run[hh] = -1e30f; l_run[hh] = 0.f; }

    for (int ft = 0; ft < NT; ++ft, ++u) {
      amb_wait(full + 8u * slot, par);
#if defined(VC_DOT_PROBE) && VC_DOT_PROBE == 1
      __syncwarp();                                // probe: no compute (the producer / HBM side alone)
      if (lane == 0) amb_arrive(empty + 8u * slot);
      ring_adv(slot, par, 1, nslots);
      continue;
#endif
      const uint32_t ktile = (uint32_t)__cvta_generic_to_shared(ring + (size_t)slot * slot_elems);
      const uint32_t vtile = ktile + (SEPV ? (uint32_t)tile_elems * 2u : 0u);
      // ---- scores of beam g for frames ft*16 + {2tg, 2tg+1, 2tg+8, 2tg+9}, per head of this warp
      float xs[HP][4];
      if constexpr (HPW == 0) {
        // (even and odd k-steps: two independent accumulation chains)
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < MAXMT; ++i) {
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(ktile + (uint32_t)(lrowA * pitch + w * cw + i * 16 + lcolA) * 2u, a0, a1, a2, a3);
          mma_bf16((i & 1) ? s1 : s0, a0, a1, a2, a3, qh[i][0], qh[i][1]);
        }
        // accumulator layout: {S[g][2tg], S[g][2tg+1], S[g+8][2tg], S[g+8][2tg+1]} (frame, beam) -> exchange [beam][frame]
        float* pp = part + (size_t)((u & 1) * kDotCW + w) * 128;
        pp[(2 * tg) * 16 + g] = s0[0] + s1[0];
        pp[(2 * tg + 1) * 16 + g] = s0[1] + s1[1];
        pp[(2 * tg) * 16 + g + 8] = s0[2] + s1[2];
        pp[(2 * tg + 1) * 16 + g + 8] = s0[3] + s1[3];
        named_bar_sync(1, 32 * kDotCW);            // (double-buffered: the next tile's partials go to the other half)
        const float* rp = part + (size_t)(u & 1) * kDotCW * 128 + g * 16 + 2 * tg;
        float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
#pragma unroll
        for (int ww = 0; ww < kDotCW; ++ww) {
          const float2 lo = *reinterpret_cast<const float2*>(rp + ww * 128);
          const float2 hi = *reinterpret_cast<const float2*>(rp + ww * 128 + 8);
          x0 += lo.x; x1 += lo.y; x2 += hi.x; x3 += hi.y;
        }
        xs[0][0] = x0; xs[0][1] = x1; xs[0][2] = x2; xs[0][3] = x3;
      } else {
        const int srcA = (2 * tg) * 4 + (g >> 1), srcB = (2 * tg + 1) * 4 + (g >> 1);
        const bool odd = (g & 1) != 0;
#pragma unroll
        for (int hh = 0; hh < HP; ++hh) {
          float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < NKH; ++j) {
            const int i = hh * NKH + j;
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4(ktile + (uint32_t)(lrowA * pitch + w * cw + i * 16 + lcolA) * 2u, a0, a1, a2, a3);
            mma_bf16((j & 1) ? s1 : s0, a0, a1, a2, a3, qh[i][0], qh[i][1]);
          }
          // the head's score tile is complete in this warp: (frame, beam) accumulator layout -> (beam g, 4 frames) per lane
          const float v0 = s0[0] + s1[0], v1 = s0[1] + s1[1], v2 = s0[2] + s1[2], v3 = s0[3] + s1[3];
          const float a0_ = __shfl_sync(0xffffffffu, v0, srcA), a1_ = __shfl_sync(0xffffffffu, v1, srcA);
          const float a2_ = __shfl_sync(0xffffffffu, v2, srcA), a3_ = __shfl_sync(0xffffffffu, v3, srcA);
          const float b0_ = __shfl_sync(0xffffffffu, v0, srcB), b1_ = __shfl_sync(0xffffffffu, v1, srcB);
          const float b2_ = __shfl_sync(0xffffffffu, v2, srcB), b3_ = __shfl_sync(0xffffffffu, v3, srcB);
          xs[hh][0] = odd ? a1_ : a0_;             // frame 2tg
          xs[hh][1] = odd ? b1_ : b0_;             // frame 2tg + 1
          xs[hh][2] = odd ? a3_ : a2_;             // frame 2tg + 8
          xs[hh][3] = odd ? b3_ : b2_;             // frame 2tg + 9
        }
      }
      // ---- mask (attention.py:61 / :176 / :254), frames past T
      const int tb = ft * 16 + 2 * tg;
      const int ts[4] = {tb, tb + 1, tb + 8, tb + 9};
      bool dead[4], masked[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        dead[i] = ts[i] >= T;
        masked[i] = !dead[i] && a.mask != nullptr && a.mask[(int64_t)b * T + ts[i]] == 0.f;
      }
#pragma unroll
      for (int hh = 0; hh < HP; ++hh) {
#pragma unroll
        for (int i = 0; i < 4; ++i) xs[hh][i] = dead[i] ? -INFINITY : (masked[i] ? -1e9f : xs[hh][i] * a.scale);
        // ---- online softmax over the tile (the 4 lanes of a group hold the 16 frames of beam g)
        float tmax = fmaxf(fmaxf(xs[hh][0], xs[hh][1]), fmaxf(xs[hh][2], xs[hh][3]));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
        const float m_new = fmaxf(m_run[hh], tmax);
        float rs = exp2f((m_run[hh] - m_new) * kL2e);
        float pv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pv[i] = exp2f((xs[hh][i] - m_new) * kL2e);
        float psum = (pv[0] + pv[1]) + (pv[2] + pv[3]);
        psum += __shfl_xor_sync(0xffffffffu, psum, 1);
        psum += __shfl_xor_sync(0xffffffffu, psum, 2);
        l_run[hh] = l_run[hh] * rs + psum;
        m_run[hh] = m_new;
        if (g >= K) { rs = 1.f; pv[0] = pv[1] = pv[2] = pv[3] = 0.f; }
        // accumulator columns of this lane are beams 2tg, 2tg+1: their rescale factors live in lanes 8tg, 8tg+4
        const float sc0 = __shfl_sync(0xffffffffu, rs, 8 * tg);
        const float sc1 = __shfl_sync(0xffffffffu, rs, 8 * tg + 4);
        uint32_t bh0, bl0, bh1, bl1;
        split_bf16x2(pv[0], pv[1], bh0, bl0);
        split_bf16x2(pv[2], pv[3], bh1, bl1);
        // ---- context: ctx^T[col, beam] += tile[t, col] * p[beam, t] over the head's columns
#pragma unroll
        for (int j = 0; j < (HPW > 0 ? NKH : MAXMT); ++j) {
          const int i = (HPW > 0 ? hh * NKH : 0) + j;
          {
            if (ft > 0) { c[i][0] *= sc0; c[i][1] *= sc1; c[i][2] *= sc0; c[i][3] *= sc1; }
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4_trans(vtile + (uint32_t)(lrowT * pitch + w * cw + i * 16 + lcolT) * 2u, a0, a1, a2, a3);
            mma_bf16(c[i], a0, a1, a2, a3, bh0, bh1);
            mma_bf16(c[i], a0, a1, a2, a3, bl0, bl1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) amb_arrive(empty + 8u * slot);
      ring_adv(slot, par, 1, nslots);
    }
    // ---- normalise and store: c[i] = {ctx[col0+g][2tg], ctx[col0+g][2tg+1], ctx[col0+g+8][2tg], ctx[col0+g+8][2tg+1]} (col, beam)
    float i0[HP], i1[HP];
#pragma unroll
    for (int hh = 0; hh < HP; ++hh) {
      const float inv = (g < K) ? 1.0f / l_run[hh] : 0.f;
      i0[hh] = __shfl_sync(0xffffffffu, inv, 8 * tg);
      i1[hh] = __shfl_sync(0xffffffffu, inv, 8 * tg + 4);
    }
    named_bar_sync(2, 32 * kDotCW);                // the previous video's rows have left the staging buffer
    const int k0 = 2 * tg, k1 = 2 * tg + 1;
#pragma unroll
    for (int i = 0; i < MAXMT; ++i) {
      {
        const int hh = HPW > 0 ? i / NKH : 0;
        const int col0 = w * cw + i * 16;
        if (k0 < K) { out_s[k0 * H + col0 + g] = __float2bfloat16_rn(c[i][0] * i0[hh]); out_s[k0 * H + col0 + g + 8] = __float2bfloat16_rn(c[i][2] * i0[hh]); }
        if (k1 < K) { out_s[k1 * H + col0 + g] = __float2bfloat16_rn(c[i][1] * i1[hh]); out_s[k1 * H + col0 + g + 8] = __float2bfloat16_rn(c[i][3] * i1[hh]); }
      }
    }
    named_bar_sync(2, 32 * kDotCW);
    for (int k = 0; k < K; ++k)
      for (int c8 = tid * 8; c8 < H; c8 += 32 * kDotCW * 8)
        *reinterpret_cast<uint4*>(a.ctx + ((int64_t)b * K + k) * a.ctx_ld + c8) = *reinterpret_cast<const uint4*>(out_s + k * H + c8);
  }
}

inline bool attn_dot_enabled() {       // VC_DISABLE_ATTN_DOT=1: generic kernel for the dot-product forms (A/B testing)
  const char* e = getenv("VC_DISABLE_ATTN_DOT");
  return !(e != nullptr && e[0] == '1');
}
// heads = 1 for the Luong forms.  sepv: separate value tile (multi-head).
inline bool attn_dot_ws_ok(int K, int H, int T, int heads, bool sepv, bool weights) {
  if (!attn_dot_enabled() || weights || K < 1 || K > 8 || T < 1 || heads < 1) return false;
  if (!(H == 128 || H == 256 || H == 512 || H == 1024)) return false;
  const int cw = attn_dot_warps(H);
  if (sepv != (heads > 1)) return false;           // (single-head attention over separate K / V tiles: generic kernel)
  if (heads > 1) {
    const int nmt = H / (16 * cw), hpw = heads / cw;
    if (heads % cw != 0 || !(hpw == 1 || hpw == 2) || nmt % hpw != 0) return false;
  }
  const size_t slot = (size_t)(sepv ? 2 : 1) * 16 * (H + kEncPad) * 2;
  const size_t fixed = 1024 * (size_t)cw + (size_t)K * H * 2 + 2 * kDotMaxSlots * 8;
  return fixed + 2 * slot <= 225 * 1024;
}

inline int launch_attn_dot_ws(AttnDotArgs a, cudaStream_t stream) {
  const bool sepv = a.values != a.skeys;
  VC_CHECK(attn_dot_ws_ok(a.K, a.H, a.T, a.heads, sepv, false), "dot attention (ws): K=%d H=%d T=%d heads=%d not supported", a.K, a.H, a.T, a.heads);
  VC_CHECK(a.ctx_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a.ctx) & 15) == 0 && a.q_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(a.q_act) & 3) == 0,
           "dot attention (ws): ctx must be 16-byte aligned, queries 4-byte aligned");
  const int cw = attn_dot_warps(a.H);
  const size_t slot = (size_t)(sepv ? 2 : 1) * 16 * (a.H + kEncPad) * 2;
  const size_t fixed = 1024 * (size_t)cw + (size_t)a.K * a.H * 2 + 2 * kDotMaxSlots * 8;
  // two CTAs per SM when at least 3 slots fit into half an SM's shared memory (the other CTA's tiles keep HBM busy while
  // this one finishes a video), else one CTA with as many slots as fit
  int per_sm = 2;
  size_t ns = fixed < 113 * 1024 ? (113 * 1024 - fixed) / slot : 0;
  if (ns < 3) {
    per_sm = 1;
    ns = (225 * 1024 - fixed) / slot;
  }
  a.nslots = (int)(ns < (size_t)kDotMaxSlots ? ns : (size_t)kDotMaxSlots);
  const size_t smem = fixed + (size_t)a.nslots * slot;
  const int cap = per_sm * attn_num_sms();
  const int grid = a.B < cap ? a.B : cap;
  const int hpw = a.heads > 1 ? a.heads / cw : 0;
#define VC_DOT_LAUNCH(CW, MT, HW)                                                                      \
  do {                                                                                                 \
    auto kern = attn_dot_ws_kernel<CW, MT, HW>;                                                        \
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    VC_CUDA(launch_pdl(kern, dim3(grid), dim3(32 * (CW + 1)), smem, stream, a));                       \
  } while (0)
#define VC_DOT_SINGLE(CW, MT) VC_DOT_LAUNCH(CW, MT, 0)
#define VC_DOT_MHA(CW, MT)                                             \
  do {                                                                 \
    if (hpw == 1) VC_DOT_LAUNCH(CW, MT, 1);                            \
    else VC_DOT_LAUNCH(CW, MT, 2);                                     \
  } while (0)
  if (a.H == 1024) {
    if (hpw == 0) VC_DOT_SINGLE(4, 16);
    else VC_DOT_MHA(4, 16);
  } else if (a.H == 512) {
    if (hpw == 0) VC_DOT_SINGLE(8, 4);
    else VC_DOT_MHA(8, 4);
  } else if (a.H == 256) {
    if (hpw == 0) VC_DOT_SINGLE(8, 2);
    else VC_DOT_MHA(8, 2);
  } else {
    if (hpw == 0) VC_DOT_SINGLE(8, 1);
    else VC_DOT_LAUNCH(8, 1, 1);
  }
#undef VC_DOT_MHA
#undef VC_DOT_SINGLE
#undef VC_DOT_LAUNCH
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
