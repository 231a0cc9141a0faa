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

// Requirements: K % 4 == 0, N % 4 == 0, lda/ldw % 4 == 0, a_col0/a_split/a_skip % 4 == 0, 16B-aligned bases.
template <class Epi>
int launch_sgemm(const GemmArgs& g, const Epi& epi, cudaStream_t stream) {
  VC_CHECK(g.K % 4 == 0 && g.N % 4 == 0 && g.lda % 4 == 0 && g.ldw % 4 == 0 && g.a_col0 % 4 == 0 &&
               g.a_split % 4 == 0 && g.a_skip % 4 == 0,
           "sgemm: K,N,lda,ldw and A column offsets must be multiples of 4 (M=%d N=%d K=%d)", g.M, g.N, g.K);
  if (g.M == 0 || g.N == 0) return VC_OK;
  const int nz = g.nz;
  const int gx = (g.N + 127) / 128;
  if (g.M > 64) {
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
