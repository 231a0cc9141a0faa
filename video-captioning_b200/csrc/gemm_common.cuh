// GEMM argument block and fused epilogues shared by the fp32 (FFMA) and bf16 (tcgen05) GEMM kernels.
//
// Every linear layer on the path is  C[M,N] = A[M,K] . W[N,K]^T (+ bias)  with both operands K-major,
// exactly nn.Linear's layout, so reference weights are used as stored (no transposes).
// Epilogues are called with groups of 4 consecutive output columns of one row.
#pragma once
#include "common.cuh"

namespace vc {

// grid.z selects one of (up to) two independent problems sharing shapes: the two LSTM directions.
struct GemmArgs {
  const void* A[2];   // [M, *] row-major, element type = kernel's operand type
  int64_t lda;        // elements
  int a_col0;         // first K column inside a row of A
  int a_split;        // logical k >= a_split reads column (a_col0 + k + a_skip): lets a GEMM read a
  int a_skip;         //   concatenation [seg0 | seg1] whose segments are not adjacent in memory
  const void* W[2];   // [N, K] row-major
  int64_t ldw;
  int M, N, K;
  int nz;             // 1 or 2
  // TMA hint (bf16 path): A[z] points inside a row of a larger 2D buffer whose row 0 starts at a_origin
  // and spans a_origin_cols columns (e.g. timestep t of a [B, T*2H] layer output).  nullptr: A[z] is the origin.
  const void* a_origin;
  int64_t a_origin_cols;
  // Tile-level hand-over from a producer LSTM GEMM still running (EpiLstm::sync_signal): rows [m0, m0+128) of A may be
  // read once sync_wait[m0 >> sync_row_shift] >= sync_target.  nullptr: plain stream order.  Honoured by the 128x128
  // tensor-core kernel only (the context projection).
  const unsigned int* sync_wait;
  unsigned int sync_target;
  int sync_row_shift;
  // ... and the producer side of the same protocol for a plain-store GEMM on the 128x128 tensor-core kernel (the context
  // projection hands its rows over to the vocabulary projection): sync_signal[m0 >> 7] is incremented once per finished tile,
  // after its stores have completed.  nullptr: no signalling.
  unsigned int* sync_signal;
  // bf16 store GEMMs: 1 = take the persistent 128x256-tile kernel even with fewer tiles than SMs, 2 = its CTA-pair form
  // (A/B testing of the context projection, VC_CTX_PERSISTENT); both ignore sync_wait / sync_signal.  3 = 128 x 192 tiles on the
  // one-tile-per-CTA kernel when that makes a single wave at one CTA per SM (gemm_tc.cuh: wide_tiles_ok; keeps the hand-over)
  int force_persistent;
};

// ---------------------------------------------------------------- plain store (+bias, +tanh)
template <class OutT, bool TANH, bool PRECISE>
struct EpiStore {
  OutT* C[2];
  int64_t ldc;
  float* C2[2];        // optional fp32 copy of the result (nullptr to skip)
  int64_t ldc2;
  const float* bias[2];  // nullable
  __device__ __forceinline__ void operator()(int z, int row, int col, float (&v)[4]) const {
    if (bias[z] != nullptr) {
      float4 b = *reinterpret_cast<const float4*>(bias[z] + col);
      v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
    }
    if (TANH) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = tanh_<PRECISE>(v[i]);
    }
    store4(C[z] + (int64_t)row * ldc + col, v);
    if (C2[z] != nullptr) store4(C2[z] + (int64_t)row * ldc2 + col, v);
  }
};

// ---------------------------------------------------------------- fused LSTM cell
// Weight rows are gate-interleaved on the host: output column n = 4*unit + gate, gate order i,f,g,o
// (nn.LSTM order; encoder.py:35 / decoder.py:44), so one 4-column group is one hidden unit.
//   gates = acc + bias + addend ;  c' = f*c + i*g ;  h' = o*tanh(c')
// `addend` is the all-timestep input projection (encoder) -- nullptr for the decoder where the input
// projection is part of the same GEMM (A = [x | h_prev], W = [W_ih | W_hh]).
template <class ActT, class XT, bool PRECISE>
struct EpiLstm {
  const float* bias[2];     // [4H] interleaved, nullable
  const XT* addend[2];      // row stride add_ld, nullable
  int64_t add_ld;
  const float* c_prev[2];   // [M,H]
  float* c_new[2];          // [M,H] (may alias c_prev: each element is read then written by one thread)
  int64_t c_ld;
  ActT* h_out0[2];          // primary h destination
  int64_t h0_ld;
  ActT* h_out1[2];          // optional second h destination
  int64_t h1_ld;
  // masked (packed-sequence) variant, encoder.py:74-82: rows with t >= lengths[row] keep their state
  // and emit zeros.  lengths == nullptr -> every row is valid.
  const int* lengths;
  int t_of_z[2];
  const ActT* h_prev[2];    // needed only when lengths != nullptr
  int64_t hp_ld;
  // TMA hints (bf16 staged epilogue): origin (row 0, column 0) and addressable columns of the 2D buffers the
  // pointers above point into.  c_tma_cols == 0 disables the staged path.
  const void* add_origin; int64_t add_origin_cols;
  const void* c_origin_in; const void* c_origin_out; int64_t c_tma_cols;
  const void* h0_origin; int64_t h0_origin_cols;
  const void* h1_origin; int64_t h1_origin_cols;
  // Tile-level hand-over between two stacked LSTM GEMMs (persistent tensor-core kernel only, see gemm_tc.cuh):
  // sync_signal[mt] is incremented once per finished (m-tile row mt, n-tile, CTA) after the h/c stores are globally
  // visible; a kernel given sync_wait does NOT wait for the previous grid as a whole but, per m-tile row, until
  // sync_wait[mt] >= sync_target.  Both nullptr: plain stream order.
  unsigned int* sync_signal;
  const unsigned int* sync_wait;
  unsigned int sync_per_row;     // out: arrivals per m-tile row and launch (filled by the launcher)
  unsigned int sync_target;      // in: value sync_wait[mt] must reach

  __device__ __forceinline__ void operator()(int z, int row, int col, float (&v)[4]) const {
    const int u = col >> 2;
    if (bias[z] != nullptr) {
      float4 b = *reinterpret_cast<const float4*>(bias[z] + col);
      v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
    }
    if (addend[z] != nullptr) {
      float a[4];
      load4(addend[z] + (int64_t)row * add_ld + col, a);
      v[0] += a[0]; v[1] += a[1]; v[2] += a[2]; v[3] += a[3];
    }
    const float cp = c_prev[z][(int64_t)row * c_ld + u];
    const float ig = sigmoid_<PRECISE>(v[0]);
    const float fg = sigmoid_<PRECISE>(v[1]);
    const float gg = tanh_<PRECISE>(v[2]);
    const float og = sigmoid_<PRECISE>(v[3]);
    float cn = fmaf(fg, cp, ig * gg);
    float hn = og * tanh_<PRECISE>(cn);
    ActT hs = from_float<ActT>(hn);
    if (lengths != nullptr && t_of_z[z] >= lengths[row]) {
      cn = cp;
      hs = from_float<ActT>(0.f);
      // state carried forward unchanged; the layer output at this (padded) frame is zero
      if (h_out1[z] != nullptr) h_out1[z][(int64_t)row * h1_ld + u] = h_prev[z] ? h_prev[z][(int64_t)row * hp_ld + u] : hs;
      c_new[z][(int64_t)row * c_ld + u] = cn;
      h_out0[z][(int64_t)row * h0_ld + u] = hs;
      return;
    }
    c_new[z][(int64_t)row * c_ld + u] = cn;
    h_out0[z][(int64_t)row * h0_ld + u] = hs;
    if (h_out1[z] != nullptr) h_out1[z][(int64_t)row * h1_ld + u] = hs;
  }
};

}  // namespace vc
