// Decode-loop glue kernels: state init, greedy arg-max, beam candidate extraction + selection,
// beam reorder fused with the next step's embedding gather, and final caption assembly.
//
// Replaces (reference file:line): decoder.py:81-106 init_hidden_state, :130 embedding, :265-275 greedy
// pick; video_captioning_model.py:191-196 beam init, :209 log_softmax, :211-220 score+topk+div/mod,
// :226-272 the Python batch x beam bookkeeping loop, :274-300 final selection and START padding.
// Everything runs on the device with no host synchronisation inside the step loop.
#pragma once
#include <math.h>

#include "common.cuh"

namespace vc {

// Per-layer recurrent-state plumbing of the decoder (see DESIGN.md "Decoder row buffers").
template <class ActT>
struct DecState {
  int L;                 // decoder layers (<= 4)
  int H, E;
  ActT* x_rec[4];        // destination of h_prev for layer l inside its GEMM A buffer (row stride x_ld[l])
  int64_t x_ld[4];
  ActT* h_new[4];        // [R,H] h produced by this step's layer-l GEMM epilogue
  float* c[4];           // [R,H] cell state consumed by the next step
  float* c_new[4];       // [R,H] cell state produced by this step
  ActT* emb_dst;         // embedding destination inside layer 0's A buffer (row stride emb_ld)
  int64_t emb_ld;
  const ActT* emb_table; // [V,E]
};

// ---------------------------------------------------------------- init (step -1)
template <class ActT>
__global__ void decode_init_kernel(DecState<ActT> st, const float* __restrict__ enc_final /*[B,H]*/, int R, int K,
                                   int start_id, const int* __restrict__ init_tok, int64_t init_stride,
                                   int* __restrict__ cur_tok, float* __restrict__ scores,
                                   unsigned char* __restrict__ alive, int* __restrict__ done,
                                   float* __restrict__ best_score, int* __restrict__ best_len, int diverse) {
  const int r = blockIdx.x;
  const int b = r / K;
  for (int l = 0; l < st.L; ++l) {
    for (int u = threadIdx.x; u < st.H; u += blockDim.x) {
      st.x_rec[l][(int64_t)r * st.x_ld[l] + u] = from_float<ActT>(enc_final[(int64_t)b * st.H + u]);   // decoder.py:103
      st.c[l][(int64_t)r * st.H + u] = 0.f;                                                            // :104
    }
  }
  const int tok0 = init_tok ? init_tok[(int64_t)r * init_stride] : start_id;   // teacher forcing feeds its own first token
  for (int e = threadIdx.x; e < st.E; e += blockDim.x)
    st.emb_dst[(int64_t)r * st.emb_ld + e] = st.emb_table[(int64_t)tok0 * st.E + e];
  if (threadIdx.x == 0) {
    cur_tok[r] = tok0;
    // reference: all K beam scores start at 0 (video_captioning_model.py:194).  `diverse` is the opt-in
    // standard beam search (only beam 0 live at step 0), SURVEY.md section 8f rank 3.
    if (scores) scores[r] = (diverse && (r % K) != 0) ? -INFINITY : 0.f;
    if (alive) alive[r] = 1;
    if (r % K == 0) {
      if (done) done[b] = 0;
      if (best_score) best_score[b] = -INFINITY;
      if (best_len) best_len[b] = 0;
    }
  }
}

// ---------------------------------------------------------------- greedy arg-max (decoder.py:265-269)
// One CTA per row; ties resolve to the lowest index (torch.argmax CPU behaviour).
__global__ void __launch_bounds__(256) greedy_argmax_kernel(const float* __restrict__ logits, int64_t ld, int V,
                                                            float inv_temp_is_one, float temperature,
                                                            int* __restrict__ cur_tok, int* __restrict__ tokens_out,
                                                            int S, int step) {
  const int r = blockIdx.x;
  const float* row = logits + (int64_t)r * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x * 4; i < V; i += blockDim.x * 4) {
    float4 x = *reinterpret_cast<const float4*>(row + i);
    float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y = (inv_temp_is_one != 0.f) ? v[j] : v[j] / temperature;
      if (y > best) { best = y; bi = i + j; }
    }
  }
  // warp then block reduce on (value desc, index asc)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    cur_tok[r] = bi;
    tokens_out[(int64_t)r * S + step] = bi;
  }
}

// ---------------------------------------------------------------- beam: per-row log-softmax stats + top-K
// (video_captioning_model.py:209 log_softmax, first half of :215 topk).  The top-K over the K*V
// candidates of a video is contained in the union of the per-row top-K, so each row only exports its K
// best log-probs.  One CTA per row, one pass over the logits (HBM-bound: R*V*4 bytes per step):
// per-thread online (max, sum-exp) + sorted K-list, then shuffle-only merges (lanes -> warp -> CTA).
// Order everywhere: value desc, vocabulary index asc (ties resolve to the lower index).
template <int KMAX>
struct TopList {
  float v[KMAX];
  int i[KMAX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { v[k] = -INFINITY; i[k] = 0x7fffffff; }
  }
  // insert (y, idx); caller guarantees y > v[KMAX-1] or (y == v[KMAX-1] && idx < i[KMAX-1])
  __device__ __forceinline__ void insert(float y, int idx) {
    v[KMAX - 1] = y; i[KMAX - 1] = idx;
#pragma unroll
    for (int k = KMAX - 1; k > 0; --k) {
      const bool up = v[k] > v[k - 1] || (v[k] == v[k - 1] && i[k] < i[k - 1]);
      if (up) {
        const float a = v[k]; v[k] = v[k - 1]; v[k - 1] = a;
        const int c = i[k]; i[k] = i[k - 1]; i[k - 1] = c;
      }
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int k = 0; k < KMAX - 1; ++k) { v[k] = v[k + 1]; i[k] = i[k + 1]; }
    v[KMAX - 1] = -INFINITY; i[KMAX - 1] = 0x7fffffff;
  }
};

// K rounds of warp arg-max over the lanes' list heads; lane `k` ends up holding the k-th best of the warp.
template <int KMAX>
__device__ __forceinline__ void warp_merge_topk(TopList<KMAX>& l, int K, int lane, float& out_v, int& out_i) {
  out_v = -INFINITY; out_i = 0x7fffffff;
  for (int k = 0; k < K; ++k) {
    float bv = l.v[0];
    int bi = l.i[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (l.i[0] == bi && l.v[0] == bv) l.pop();     // indices are unique within a row: exactly one lane pops
    if (lane == k) { out_v = bv; out_i = bi; }
  }
}

template <int KMAX, bool PRECISE>
__global__ void __launch_bounds__(256) beam_row_topk_kernel(const float* __restrict__ logits, int64_t ld, int V, int K,
                                                            float* __restrict__ cand_val /*[R,K] log-prob*/,
                                                            int* __restrict__ cand_idx /*[R,K]*/) {
  const int r = blockIdx.x;
  const float* row = logits + (int64_t)r * ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarp = 8;
  TopList<KMAX> tl;
  tl.init();
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x * 4; i < V; i += 256 * 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(row + i));
    const float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float y = v[j];
      if (y > m) {
        s *= PRECISE ? expf(m - y) : __expf(m - y);
        m = y;
      }
      s += PRECISE ? expf(y - m) : __expf(y - m);
      if (y > tl.v[KMAX - 1]) tl.insert(y, i + j);   // strictly greater: the earlier index keeps ties
    }
  }
  // log-sum-exp of the row
  __shared__ float sm[nwarp], ss[nwarp], wv[nwarp][KMAX];
  __shared__ int wi[nwarp][KMAX];
  const float wm = warp_max(m);
  const float wsum = warp_sum(m == -INFINITY ? 0.f : s * (PRECISE ? expf(m - wm) : __expf(m - wm)));
  float ov;
  int oi;
  warp_merge_topk<KMAX>(tl, K, lane, ov, oi);
  if (lane == 0) { sm[warp] = wm; ss[warp] = wsum; }
  if (lane < K) { wv[warp][lane] = ov; wi[warp][lane] = oi; }
  __syncthreads();
  if (warp == 0) {
    float M = sm[0];
#pragma unroll
    for (int w = 1; w < nwarp; ++w) M = fmaxf(M, sm[w]);
    float S = 0.f;
#pragma unroll
    for (int w = 0; w < nwarp; ++w) S += ss[w] * (PRECISE ? expf(sm[w] - M) : __expf(sm[w] - M));
    const float lse = M + logf(S);
    // lanes 0..7 adopt warp w's (already sorted) K-list and the 8 lists are merged the same way
    TopList<KMAX> t2;
    t2.init();
    if (lane < nwarp) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) { t2.v[k] = wv[lane][k]; t2.i[k] = wi[lane][k]; }
    }
    warp_merge_topk<KMAX>(t2, K, lane, ov, oi);
    if (lane < K) {
      cand_val[(int64_t)r * K + lane] = ov - lse;    // log_softmax value of the k-th best token
      cand_idx[(int64_t)r * K + lane] = oi;
    }
  }
}

// ---------------------------------------------------------------- beam: per-video selection
// One thread per video (K*K <= 256 candidates).  video_captioning_model.py:211-272 restated for the
// per-video (B=1) semantics of SURVEY.md section 3.3 with a fixed K-row layout:
//   * candidates come only from live beams:  score[r] + logp            (:211)
//   * top-K by (score desc, flat index = beam*V + token asc)            (:215-220)
//   * token == END: completed, score / (len-1)^length_penalty, first maximum kept   (:237-242, :277-281)
//   * otherwise the candidate becomes the next live beam; live beams are compacted to the front in
//     selection order, as the reference's re-stacking does                         (:244-272)
//   * a video with no live beam left is done                            (:251)
struct BeamState {
  float* scores;            // [R]
  unsigned char* alive;     // [R]
  int* done;                // [B]
  float* best_score;        // [B] best completed (normalised) score
  int* best_len;            // [B] generated length of the best completed hypothesis (0 = none)
  int* best_seq;            // [B,S]
  int* hist[2];             // [R,S] token history, ping-pong
};

__global__ void beam_select_kernel(BeamState bs, const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                   int B, int K, int V, int S, int step, int end_id, float length_penalty,
                                   int* __restrict__ parent /*[R]*/, int* __restrict__ cur_tok /*[R]*/) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int* hin = bs.hist[step & 1];
  int* hout = bs.hist[(step + 1) & 1];
  const int r0 = b * K;
  unsigned used[8];   // bitset over K*K <= 256 candidates
  for (int i = 0; i < 8; ++i) used[i] = 0u;
  int n_alive = 0;
  float new_score[16];
  int new_parent[16], new_tok[16];
  const int n_live_in = [&] { int c = 0; for (int k = 0; k < K; ++k) c += bs.alive[r0 + k] ? 1 : 0; return c; }();
  if (n_live_in > 0) {
    for (int sel = 0; sel < K; ++sel) {
      float bv = -INFINITY;
      long long bflat = 0x7fffffffffffffffLL;
      int bc = -1;
      for (int k = 0; k < K; ++k) {
        if (!bs.alive[r0 + k]) continue;
        const float base = bs.scores[r0 + k];
        for (int j = 0; j < K; ++j) {
          const int c = k * K + j;
          if (used[c >> 5] & (1u << (c & 31))) continue;
          const float v = base + cand_val[(int64_t)(r0 + k) * K + j];
          const long long flat = (long long)k * V + cand_idx[(int64_t)(r0 + k) * K + j];
          if (v > bv || (v == bv && flat < bflat) || bc < 0) { bv = v; bflat = flat; bc = c; }
        }
      }
      if (bc < 0) break;
      used[bc >> 5] |= 1u << (bc & 31);
      const int pk = bc / K;
      const int tok = cand_idx[(int64_t)(r0 + pk) * K + (bc - pk * K)];
      if (tok == end_id) {
        const float denom = (float)pow((double)(step + 1), (double)length_penalty);   // (len(new_seq)-1)**lp
        const float fin = bv / denom;
        if (bs.best_len[b] == 0 || fin > bs.best_score[b]) {
          bs.best_score[b] = fin;
          bs.best_len[b] = step + 1;
          for (int i = 0; i < step; ++i) bs.best_seq[(int64_t)b * S + i] = hin[(int64_t)(r0 + pk) * S + i];
          bs.best_seq[(int64_t)b * S + step] = tok;
        }
      } else {
        new_score[n_alive] = bv;
        new_parent[n_alive] = pk;
        new_tok[n_alive] = tok;
        ++n_alive;
      }
    }
  }
  for (int k = 0; k < K; ++k) {
    const int r = r0 + k;
    if (k < n_alive) {
      bs.scores[r] = new_score[k];
      bs.alive[r] = 1;
      parent[r] = r0 + new_parent[k];
      cur_tok[r] = new_tok[k];
      for (int i = 0; i < step; ++i) hout[(int64_t)r * S + i] = hin[(int64_t)(r0 + new_parent[k]) * S + i];
      hout[(int64_t)r * S + step] = new_tok[k];
    } else {
      // dead slot: keeps computing on benign inputs, never contributes candidates
      bs.alive[r] = 0;
      bs.scores[r] = -INFINITY;
      parent[r] = r;
      cur_tok[r] = end_id;
      for (int i = 0; i < step; ++i) hout[(int64_t)r * S + i] = hin[(int64_t)r * S + i];
      hout[(int64_t)r * S + step] = end_id;
    }
  }
  if (n_alive == 0) bs.done[b] = 1;
}

// ---------------------------------------------------------------- reorder (+ next-step embedding)
// video_captioning_model.py:247-249,269-272 (clone/cat of the kept (h,c) columns) and decoder.py:130 for
// the next step.  One CTA per row: gathers the parent row's new state into this row's GEMM operands.
template <class ActT>
__global__ void __launch_bounds__(128) reorder_embed_kernel(DecState<ActT> st, const int* __restrict__ parent,
                                                            const int* __restrict__ cur_tok, int V) {
  const int r = blockIdx.x;
  const int p = parent ? parent[r] : r;
  for (int l = 0; l < st.L; ++l) {
    const ActT* hs = st.h_new[l] + (int64_t)p * st.H;
    const float* cs = st.c_new[l] + (int64_t)p * st.H;
    ActT* hd = st.x_rec[l] + (int64_t)r * st.x_ld[l];
    float* cd = st.c[l] + (int64_t)r * st.H;
    for (int u = threadIdx.x * 4; u < st.H; u += blockDim.x * 4) {
      float hv[4];
      load4(hs + u, hv);
      store4(hd + u, hv);
      *reinterpret_cast<float4*>(cd + u) = *reinterpret_cast<const float4*>(cs + u);
    }
  }
  int tok = cur_tok[r];
  tok = min(max(tok, 0), V - 1);
  const ActT* e = st.emb_table + (int64_t)tok * st.E;
  ActT* ed = st.emb_dst + (int64_t)r * st.emb_ld;
  for (int i = threadIdx.x * 4; i < st.E; i += blockDim.x * 4) {
    float ev[4];
    load4(e + i, ev);
    store4(ed + i, ev);
  }
}

// ---------------------------------------------------------------- final assembly
// video_captioning_model.py:274-300: best completed hypothesis if any, else live beam 0; index 0 is START;
// rows are right-padded with START to S+1.  lengths[b] counts the leading START.
__global__ void beam_finalize_kernel(BeamState bs, int B, int K, int S, int steps_run, int start_id,
                                     int* __restrict__ out_tokens /*[B,S+1]*/, int* __restrict__ out_len,
                                     float* __restrict__ out_score) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int* hist = bs.hist[steps_run & 1];
  int n;
  const int* src;
  if (bs.best_len[b] > 0) { n = bs.best_len[b]; src = bs.best_seq + (int64_t)b * S; }
  else { n = steps_run; src = hist + (int64_t)(b * K) * S; }
  out_tokens[(int64_t)b * (S + 1)] = start_id;
  for (int i = 0; i < S; ++i) out_tokens[(int64_t)b * (S + 1) + 1 + i] = (i < n) ? src[i] : start_id;
  out_len[b] = n + 1;
  if (out_score) out_score[b] = (bs.best_len[b] > 0) ? bs.best_score[b] : bs.scores[b * K];
}

// fp32 -> bf16 conversion for GEMM operands (features, when not consumed as tf32)
__global__ void convert_f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float v[4];
    load4(in + i * 4, v);
    store4(out + i * 4, v);
  }
}

}  // namespace vc
