// fp32 GEMM on the CUDA cores (FFMA), used by the fp32 parity mode.
//
// Why not tensor cores here: the fp32 mode has to reproduce the CPU reference's greedy tokens
// bit-for-bit with top1-top2 logit margins down to ~1e-5, which rules out
// TF32/bf16 operand rounding.  Accumulation is plain fp32 in ascending-k order.
//
// Tiling: BM x 128 x 16 block tile, register-blocked TM x 8 per thread, operands transposed into
// shared memory (k-major) so the inner loop is LDS.128 + FFMA, global loads are 128-bit along K and
// register-prefetched one tile ahead.
#pragma once
#include "gemm_common.cuh"

namespace vc {

template <int BM, int TM, class Epi>
__global__ void __launch_bounds__((BM / TM) * 16)
sgemm_nt_kernel(const GemmArgs g, const Epi epi) {
  constexpr int BN = 128, BK = 16, TN = 8;
  constexpr int NTX = BN / TN;         // 16
  constexpr int NTY = BM / TM;
  constexpr int NT = NTX * NTY;
  constexpr int PAD = 4;
  constexpr int A_F4 = BM * BK / 4;    // float4 per A tile
  constexpr int B_F4 = BN * BK / 4;
  constexpr int A_PER = (A_F4 + NT - 1) / NT;
  constexpr int B_PER = B_F4 / NT;
  static_assert(B_F4 % NT == 0, "tile/thread mismatch");

  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int z = blockIdx.z;
  const float* __restrict__ A = reinterpret_cast<const float*>(g.A[z]);
  const float* __restrict__ W = reinterpret_cast<const float*>(g.W[z]);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid % NTX, ty = tid / NTX;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_PER], rb[B_PER];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int f = tid + i * NT;
      const int r = f >> 2, kq = (f & 3) * 4;
      const int k = k0 + kq, row = m0 + r;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < A_F4 && row < g.M && k < g.K) {
        const int col = g.a_col0 + k + (k >= g.a_split ? g.a_skip : 0);
        ra[i] = *reinterpret_cast<const float4*>(A + (int64_t)row * g.lda + col);
      }
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int f = tid + i * NT;
      const int r = f >> 2, kq = (f & 3) * 4;
      const int k = k0 + kq, row = n0 + r;
      rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < g.N && k < g.K) rb[i] = *reinterpret_cast<const float4*>(W + (int64_t)row * g.ldw + k);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int f = tid + i * NT;
      if (f < A_F4) {
        const int r = f >> 2, kq = (f & 3) * 4;
        As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y;
        As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int f = tid + i * NT;
      const int r = f >> 2, kq = (f & 3) * 4;
      Bs[buf][kq + 0][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y;
      Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w;
    }
  };

  const int nk = (g.K + BK - 1) / BK;
  if (nk > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      if (TM == 8) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][BM / 2 + ty * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        a[TM - 4] = a1.x; a[TM - 3] = a1.y; a[TM - 2] = a1.z; a[TM - 1] = a1.w;
      } else {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      }
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows {ty*4..+3} (+ BM/2 block when TM==8), cols {tx*4..+3} and {64+tx*4..+3}
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + (TM == 8 ? ((i < 4) ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4)) : ty * 4 + i);
    if (row >= g.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = n0 + h * (BN / 2) + tx * 4;
      if (col >= g.N) continue;
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      epi(z, row, col, v);
    }
  }
}

// ---------------------------------------------------------------- skinny GEMM (M <= 64: the fp32 mode's per-step GEMMs)
// Config 1 (greedy, 32 videos) runs every per-step GEMM -- 160 recurrent encoder steps per layer, the decoder's LSTM, context,
// query and vocabulary projections -- with M = 32.  The 32 x 128 tiles of the kernel above leave N/128 CTAs (16 per LSTM
// direction) walking K in 16-wide blocks with one global round trip each: 35 us per launch for 134 MFLOP.  Here a CTA owns a
// 32 x 32 output tile (N/32 CTAs per problem: 128 for both directions of a 512-unit LSTM), streams its [32 x K] slices of A and W
// through a 4-stage cp.async ring of 64-wide k-blocks (48 KB in flight against the L2 latency), and every thread accumulates its
// 2 x 4 outputs in ascending-k order with one fmaf per element and k -- the same summation order, hence bit-identical results,
// as sgemm_nt_kernel (the fp32 mode's token-exact parity with the CPU reference rests on it).
// Shared-memory tiles are [32 rows][64 k] without padding; within each half row the 16-byte piece c of row r sits at piece c ^ f(r) (f = (r >> 1) & 7
// for A, (r >> 2) & 7 for W) so that the LDS.128 of a warp -- 4 distinct A rows, 8 distinct W rows -- hit distinct bank groups, and
// the operands of k-step k+1 are loaded into a second register set before the FMAs of k-step k (one warp per scheduler: nothing
// else hides the LDS latency; the first version stalled ~45 clocks on each of its 48 loads per k-block, 1.3 us per k-block).
// What remains is the shared-memory -> register path: an LDS.128 delivers 512 B per warp whatever the broadcast, 128 B per clock
// and SM.  A 4 x 4 register tile per thread (2 B per FMA instead of 3, 64 threads) measured slower with one CTA per SM (14.5 vs 11.4 us at
// K = 512: half the warps to hide latency) and only 10% faster where several CTAs share an SM (N = 10000): kSkTM = 2 stays.
constexpr int kSkBM = 32, kSkBN = 32, kSkBK = 64, kSkStages = 4, kSkTM = 2, kSkThreads = (kSkBM / kSkTM) * (kSkBN / 4);

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

template <class Epi>
__global__ void __launch_bounds__(kSkThreads) sgemm_skinny_kernel(const GemmArgs g, const Epi epi) {
  extern __shared__ __align__(16) float sk_smem[];
  constexpr int kPieces = kSkBK / 4;                                // 16-byte pieces per row of a k-block
  constexpr int kChunks = kSkBM * kPieces / kSkThreads;             // pieces per thread and operand
  constexpr uint32_t kStageBytes = kSkBM * kSkBK * 4;
  const int z = blockIdx.z;
  const float* __restrict__ A = reinterpret_cast<const float*>(z == 0 ? g.A[0] : g.A[1]);
  const float* __restrict__ W = reinterpret_cast<const float*>(z == 0 ? g.W[0] : g.W[1]);
  const int m0 = blockIdx.y * kSkBM, n0 = blockIdx.x * kSkBN;
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;                            // 4 columns 4tx.., kSkTM rows kSkTM*ty..
  const int nk = (g.K + kSkBK - 1) / kSkBK;
  const uint32_t as_base = (uint32_t)__cvta_generic_to_shared(sk_smem);                 // [stages][32 rows][64 k]
  const uint32_t ws_base = as_base + kSkStages * kStageBytes;                            // [stages][32 cols][64 k]

  // piece c = tid + 128 i: row c / kPieces, piece c % kPieces of the k-block; source pointers at k-block 0, advanced per block
  const float* pa[kChunks];
  const float* pw[kChunks];
  uint32_t da[kChunks], dw[kChunks];
  bool ra[kChunks], rw[kChunks];
  int kp[kChunks];
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    const int c = tid + kSkThreads * i;
    const int r = c / kPieces, pc = c % kPieces;
    kp[i] = pc * 4;
    ra[i] = (m0 + r) < g.M;
    rw[i] = (n0 + r) < g.N;
    pa[i] = A + (int64_t)(ra[i] ? m0 + r : 0) * g.lda + g.a_col0 + pc * 4;
    pw[i] = W + (int64_t)(rw[i] ? n0 + r : 0) * g.ldw + pc * 4;
    da[i] = as_base + (uint32_t)(r * kSkBK + (((pc & ~7) | ((pc & 7) ^ ((r / kSkTM) & 7))) << 2)) * 4u;
    dw[i] = ws_base + (uint32_t)(r * kSkBK + (((pc & ~7) | ((pc & 7) ^ ((r >> 2) & 7))) << 2)) * 4u;
  }
  auto issue = [&](int kb, int stage) {
    const int k0 = kb * kSkBK;
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
      const int k = k0 + kp[i];
      const bool kin = k < g.K;
      const int adj = k0 + (k >= g.a_split ? g.a_skip : 0);
      cp_async16_zfill(da[i] + (uint32_t)stage * kStageBytes, kin ? pa[i] + adj : pa[i], kin && ra[i]);
      cp_async16_zfill(dw[i] + (uint32_t)stage * kStageBytes, kin ? pw[i] + k0 : pw[i], kin && rw[i]);
    }
  };
  for (int s = 0; s < kSkStages - 1; ++s) {
    if (s < nk) issue(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float acc[kSkTM][4];
#pragma unroll
  for (int i = 0; i < kSkTM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int fa = ty & 7, fw = tx;                                    // piece swizzles of this thread's A rows (kSkTM*ty..) / W rows (4tx..)
  const float* As = sk_smem;
  const float* Ws = sk_smem + (size_t)kSkStages * kSkBM * kSkBK;

  for (int kb = 0; kb < nk; ++kb) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kSkStages - 2) : "memory");     // k-block kb has landed (this thread's part)
    __syncthreads();                                                             // ... everyone's; and the slot refilled below is free
    const int nxt = kb + kSkStages - 1;
    if (nxt < nk) issue(nxt, nxt % kSkStages);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* as = As + ((size_t)(kb % kSkStages) * kSkBM + kSkTM * ty) * kSkBK;
    const float* ws = Ws + ((size_t)(kb % kSkStages) * kSkBN + 4 * tx) * kSkBK;
    float4 a[2][kSkTM], w[2][4];                                                 // [register set][row / column]
    auto lds = [&](int set, int pc) {
      const int pa_ = ((pc & ~7) | ((pc & 7) ^ fa)) << 2, pw_ = ((pc & ~7) | ((pc & 7) ^ fw)) << 2;
#pragma unroll
      for (int i = 0; i < kSkTM; ++i) a[set][i] = *reinterpret_cast<const float4*>(as + i * kSkBK + pa_);
#pragma unroll
      for (int j = 0; j < 4; ++j) w[set][j] = *reinterpret_cast<const float4*>(ws + j * kSkBK + pw_);
    };
    lds(0, 0);
#pragma unroll
    for (int pc = 0; pc < kPieces; ++pc) {
      const int cur = pc & 1;
      if (pc + 1 < kPieces) lds(cur ^ 1, pc + 1);
      // ascending k, one fmaf per (element, k); k outermost so that the eight accumulators' chains interleave
#pragma unroll
      for (int i = 0; i < kSkTM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[cur][i].x, w[cur][j].x, acc[i][j]);
#pragma unroll
      for (int i = 0; i < kSkTM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[cur][i].y, w[cur][j].y, acc[i][j]);
#pragma unroll
      for (int i = 0; i < kSkTM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[cur][i].z, w[cur][j].z, acc[i][j]);
#pragma unroll
      for (int i = 0; i < kSkTM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[cur][i].w, w[cur][j].w, acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < kSkTM; ++i) {
    const int row = m0 + kSkTM * ty + i, col = n0 + 4 * tx;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    if (row < g.M && col < g.N) epi(z, row, col, v);
  }
}

// Requirements: K % 4 == 0, N % 4 == 0, lda/ldw % 4 == 0, a_col0/a_split/a_skip % 4 == 0, 16B-aligned bases.
template <class Epi>
int launch_sgemm(const GemmArgs& g, const Epi& epi, cudaStream_t stream) {
  VC_CHECK(g.K % 4 == 0 && g.N % 4 == 0 && g.lda % 4 == 0 && g.ldw % 4 == 0 && g.a_col0 % 4 == 0 &&
               g.a_split % 4 == 0 && g.a_skip % 4 == 0,
           "sgemm: K,N,lda,ldw and A column offsets must be multiples of 4 (M=%d N=%d K=%d)", g.M, g.N, g.K);
  if (g.M == 0 || g.N == 0) return VC_OK;
  const int nz = g.nz;
  const int gx = (g.N + 127) / 128;
  static const bool skinny_off = getenv("VC_DISABLE_SKINNY_SGEMM") != nullptr && getenv("VC_DISABLE_SKINNY_SGEMM")[0] == '1';
  if (g.M <= 64 && !skinny_off) {
    const size_t smem = sizeof(float) * (size_t)kSkStages * (kSkBM + kSkBN) * kSkBK;
    auto kern = sgemm_skinny_kernel<Epi>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((g.N + kSkBN - 1) / kSkBN, (g.M + kSkBM - 1) / kSkBM, nz);
    kern<<<grid, kSkThreads, smem, stream>>>(g, epi);
  } else if (g.M > 64) {
    dim3 grid(gx, (g.M + 127) / 128, nz);
    sgemm_nt_kernel<128, 8, Epi><<<grid, 256, 0, stream>>>(g, epi);
  } else if (g.M > 32) {
    dim3 grid(gx, 1, nz);
    sgemm_nt_kernel<64, 4, Epi><<<grid, 256, 0, stream>>>(g, epi);
  } else {
    dim3 grid(gx, 1, nz);
    sgemm_nt_kernel<32, 4, Epi><<<grid, 128, 0, stream>>>(g, epi);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
