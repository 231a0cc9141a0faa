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
  // attention contexts computed ahead of the selection on the rows BEFORE the reorder (capi.cu: early attention):
  // row r takes ctx_src[parent[r]].  nullptr: the attention step writes the context segment itself.
  const ActT* ctx_src;   // [R,H]
  ActT* ctx_dst;         // context segment inside layer 0's A buffer (row stride ctx_ld)
  int64_t ctx_ld;
};

// ---------------------------------------------------------------- init (step -1)
template <class ActT>
__global__ void decode_init_kernel(DecState<ActT> st, const float* __restrict__ enc_final /*[B,H]*/, int R, int K,
                                   int start_id, int V, const int* __restrict__ init_tok, int64_t init_stride,
                                   int* __restrict__ cur_tok, float* __restrict__ scores,
                                   unsigned char* __restrict__ alive, int* __restrict__ done,
                                   float* __restrict__ best_score, int* __restrict__ best_len, int* __restrict__ best_slot,
                                   int diverse) {
  const int r = blockIdx.x;
  const int b = r / K;
  for (int l = 0; l < st.L; ++l) {
    for (int u = threadIdx.x; u < st.H; u += blockDim.x) {
      st.x_rec[l][(int64_t)r * st.x_ld[l] + u] = from_float<ActT>(enc_final[(int64_t)b * st.H + u]);   // decoder.py:103
      st.c[l][(int64_t)r * st.H + u] = 0.f;                                                            // :104
    }
  }
  int tok0 = init_tok ? init_tok[(int64_t)r * init_stride] : start_id;   // teacher forcing feeds its own first token
  tok0 = min(max(tok0, 0), V - 1);     // ids are range-checked on the host (nn.Embedding would raise); never read out of bounds
  for (int e = threadIdx.x; e < st.E; e += blockDim.x)
    st.emb_dst[(int64_t)r * st.emb_ld + e] = st.emb_table[(int64_t)tok0 * st.E + e];
  if (threadIdx.x == 0) {
    cur_tok[r] = tok0;
    // reference: all K beam scores start at 0 (video_captioning_model.py:194).  `diverse` is the opt-in
    // standard beam search (only beam 0 live at step 0), SURVEY.md section 8f rank 3.
    if (scores) scores[r] = (diverse && (r % K) != 0) ? -INFINITY : 0.f;
    if (alive) alive[r] = 1;
    if (r % K == 0 && done) done[b] = 0;
    // finished-hypothesis pool of the video: K entries, row r = b*K + k owns entry k (empty: len 0)
    if (best_score) best_score[r] = -INFINITY;
    if (best_len) best_len[r] = 0;
    if (best_slot) best_slot[r] = r % K;
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
// best log-probs.  One CTA per row, one pass over the logits (HBM-bound: R*V*4 bytes per step).
//
// SIMT-friendly selection: every warp keeps ONE sorted K-list distributed over its lanes (lane j holds the
// j-th best) plus the warp-uniform threshold (the K-th best).  Elements are compared against the threshold
// only; the rare survivors (about K*ln(n/K) per warp) are found with a ballot and inserted one at a time
// with a popc/shfl_up shift.  (Per-thread lists cost ~100 instructions per element: with 32 lanes some lane
// inserts at almost every element.)  Order everywhere: value desc, vocabulary index asc.
struct WarpTopK {
  float lv; int li;        // lane j < K: j-th best so far
  float thr; int thr_i;    // warp-uniform K-th best
  int K, lane;
  __device__ __forceinline__ void init(int K_, int lane_) {
    K = K_; lane = lane_; lv = -INFINITY; li = 0x7fffffff; thr = -INFINITY; thr_i = 0x7fffffff;
  }
  __device__ __forceinline__ bool passes(float y, int idx) const { return y > thr || (y == thr && idx < thr_i); }
  // warp-uniform candidate
  __device__ __forceinline__ void insert(float y, int idx) {
    const bool better = (lane < K) && (lv > y || (lv == y && li < idx));
    const int p = __popc(__ballot_sync(0xffffffffu, better));   // sorted list: the better entries are lanes 0..p-1
    if (p >= K) return;
    const float uv = __shfl_up_sync(0xffffffffu, lv, 1);
    const int ui = __shfl_up_sync(0xffffffffu, li, 1);
    if (lane == p) { lv = y; li = idx; }
    else if (lane > p && lane < K) { lv = uv; li = ui; }
    // the threshold only ever rises (it may have been seeded above the still-unfilled list tail)
    const float nt = __shfl_sync(0xffffffffu, lv, K - 1);
    const int ni = __shfl_sync(0xffffffffu, li, K - 1);
    if (nt >= thr) { thr = nt; thr_i = ni; }
  }
  // Seed the threshold with a lower bound of the K-th best: the K-th largest of the lanes' values `x`
  // (ties removed together, which only lowers the bound).  Equal elements still pass (thr_i = INT_MAX).
  __device__ __forceinline__ void seed(float x) {
    float bound = -INFINITY;
    for (int k = 0; k < K; ++k) {
      bound = warp_max(x);
      if (x == bound) x = -INFINITY;
    }
    if (bound > thr) { thr = bound; thr_i = 0x7fffffff; }
  }
  // each lane offers one (y, idx); survivors are inserted in lane order
  __device__ __forceinline__ void offer(float y, int idx) {
    unsigned mask = __ballot_sync(0xffffffffu, passes(y, idx));
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const float cy = __shfl_sync(0xffffffffu, y, src);
      const int ci = __shfl_sync(0xffffffffu, idx, src);
      if (passes(cy, ci)) insert(cy, ci);
    }
  }
};

// Streaming selection without warp collectives in the hot loop:
//   phase 1  the first 4096 columns are streamed for the log-sum-exp only, each thread keeping the maximum
//            of its 16 values; the K-th largest of the 256 thread maxima is a LOWER BOUND of the K-th best
//            logit of the row (about the 5th best of 4096 -> ~12 of the 10k columns exceed it);
//   phase 2  the remaining columns are streamed with log-sum-exp + one compare per float4; the rare survivors
//            are appended to a shared-memory buffer with a single hand-rolled atomic per thread;
//   phase 3  the first 4096 columns are re-scanned (L1/L2 hits) with the compare only;
//   phase 4  warp 0 picks the exact top-K of the survivors (value desc, index asc).
// If the buffer overflows (e.g. thousands of equal logits) warp 0 falls back to the exact streaming
// selection above.  The row is read from HBM exactly once.
__device__ __forceinline__ void topk_append(float (&y)[4], int i, float thr, int* cnt, float* bv, int* bi, int cap) {
  int c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) c += (y[j] >= thr) ? 1 : 0;
  int p;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(p) : "r"((uint32_t)__cvta_generic_to_shared(cnt)), "r"(c) : "memory");
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (y[j] >= thr) {
      if (p < cap) { bv[p] = y[j]; bi[p] = i + j; }
      ++p;
    }
}

template <bool PRECISE>
__global__ void __launch_bounds__(256) beam_row_topk_kernel(const float* __restrict__ logits, int64_t ld, int V, int K,
                                                            float* __restrict__ cand_val /*[R,K] log-prob*/,
                                                            int* __restrict__ cand_idx /*[R,K]*/) {
  constexpr int CAP = 1024, nwarp = 8, kSampleIters = 4;
  __shared__ float bv[CAP];
  __shared__ int bi[CAP];
  __shared__ int cnt;
  __shared__ float s_top[nwarp][16], sm[nwarp], ss[nwarp];
  const int r = blockIdx.x;
  const float* row = logits + (int64_t)r * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sample_end = min(V, kSampleIters * 1024);

  // ---- phase 1: sample (log-sum-exp + thread maximum)
  float m = -INFINITY, s = 0.f;
  for (int i = tid * 4; i < sample_end; i += 1024) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(row + i));
    const float m4 = fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w));
    if (m4 > m) {
      s *= PRECISE ? expf(m - m4) : __expf(m - m4);
      m = m4;
    }
    s += PRECISE ? (expf(t.x - m) + expf(t.y - m)) + (expf(t.z - m) + expf(t.w - m))
                 : (__expf(t.x - m) + __expf(t.y - m)) + (__expf(t.z - m) + __expf(t.w - m));
  }
  {
    // K largest (distinct) thread maxima of this warp -> smem; uniform control flow
    float x = m;
    for (int k = 0; k < K; ++k) {
      const float b = warp_max(x);
      if (x == b) x = -INFINITY;
      if (lane == 0) s_top[warp][k] = b;
    }
    if (tid == 0) cnt = 0;
  }
  __syncthreads();
  float thr;
  {
    // K-th largest of the nwarp*K collected values (every warp computes it redundantly: no second barrier)
    float a = -INFINITY, b2 = -INFINITY;          // two candidates per lane cover up to 64 >= nwarp*K... K <= 8 here
    const int total = nwarp * K;
    if (lane < total) a = s_top[lane / K][lane % K];
    if (lane + 32 < total) b2 = s_top[(lane + 32) / K][(lane + 32) % K];
    if (lane + 64 < total) b2 = fmaxf(b2, s_top[(lane + 64) / K][(lane + 64) % K]);   // K > 8: looser (still valid) bound
    if (lane + 96 < total) b2 = fmaxf(b2, s_top[(lane + 96) / K][(lane + 96) % K]);
    thr = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const float x = fmaxf(a, b2);
      thr = warp_max(x);
      if (a == thr) a = -INFINITY;
      else if (b2 == thr) b2 = -INFINITY;
    }
  }

  // ---- phase 2: the rest of the row
  for (int i = sample_end + tid * 4; i < V; i += 1024) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(row + i));
    float y[4] = {t.x, t.y, t.z, t.w};
    const float m4 = fmaxf(fmaxf(y[0], y[1]), fmaxf(y[2], y[3]));
    if (m4 > m) {
      s *= PRECISE ? expf(m - m4) : __expf(m - m4);
      m = m4;
    }
    s += PRECISE ? (expf(y[0] - m) + expf(y[1] - m)) + (expf(y[2] - m) + expf(y[3] - m))
                 : (__expf(y[0] - m) + __expf(y[1] - m)) + (__expf(y[2] - m) + __expf(y[3] - m));
    if (m4 >= thr) topk_append(y, i, thr, &cnt, bv, bi, CAP);
  }
  // ---- phase 3: survivors of the sample (cache hits)
  for (int i = tid * 4; i < sample_end; i += 1024) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(row + i));
    float y[4] = {t.x, t.y, t.z, t.w};
    if (fmaxf(fmaxf(y[0], y[1]), fmaxf(y[2], y[3])) >= thr) topk_append(y, i, thr, &cnt, bv, bi, CAP);
  }
  const float wm = warp_max(m);
  const float wsum = warp_sum(m == -INFINITY ? 0.f : s * (PRECISE ? expf(m - wm) : __expf(m - wm)));
  if (lane == 0) { sm[warp] = wm; ss[warp] = wsum; }
  __syncthreads();
  if (warp != 0) return;

  // ---- phase 4
  float M = sm[0];
#pragma unroll
  for (int w = 1; w < nwarp; ++w) M = fmaxf(M, sm[w]);
  float S = 0.f;
#pragma unroll
  for (int w = 0; w < nwarp; ++w) S += (sm[w] == -INFINITY) ? 0.f : ss[w] * (PRECISE ? expf(sm[w] - M) : __expf(sm[w] - M));
  const float lse = M + logf(S);
  const int n = cnt;
  float ov = -INFINITY;
  int oi = 0x7fffffff;
  if (n <= CAP) {
    // exact top-K of the n survivors: K rounds of (lane-local best, warp arg-max, remove)
    for (int k = 0; k < K; ++k) {
      float lbv = -INFINITY;
      int lbi = 0x7fffffff, lpos = -1;
      for (int p = lane; p < n; p += 32) {
        const float v = bv[p];
        const int ix = bi[p];
        if (v > lbv || (v == lbv && ix < lbi)) { lbv = v; lbi = ix; lpos = p; }
      }
      float wv = lbv;
      int wi = lbi;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float tv = __shfl_xor_sync(0xffffffffu, wv, o);
        const int ti = __shfl_xor_sync(0xffffffffu, wi, o);
        if (tv > wv || (tv == wv && ti < wi)) { wv = tv; wi = ti; }
      }
      if (lpos >= 0 && lbi == wi && lbv == wv) { bv[lpos] = -INFINITY; bi[lpos] = 0x7fffffff; }   // unique index: one lane
      __syncwarp();
      if (lane == k) { ov = wv; oi = wi; }
    }
  } else {
    WarpTopK tk;
    tk.init(K, lane);
    for (int base = 0; base < V; base += 128) {
      const int i = base + lane * 4;
      float y[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (i < V) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(row + i));
        y[0] = t.x; y[1] = t.y; y[2] = t.z; y[3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) tk.offer(y[j], i < V ? i + j : 0x7fffffff);
    }
    ov = tk.lv;
    oi = tk.li;
  }
  if (lane < K) {
    cand_val[(int64_t)r * K + lane] = ov - lse;    // log_softmax value of the k-th best token
    cand_idx[(int64_t)r * K + lane] = oi;
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
  // Finished-hypothesis pool, K entries per video, kept sorted by normalised score (descending; among equal scores the
  // hypothesis completed first stays ahead -- the one `max` keeps at video_captioning_model.py:277-281).  Entry 0 is what
  // the reference returns; the whole pool is the n-best list of the opt-in real beam search (vc_beam_nbest).
  float* best_score;        // [B,K] normalised score
  int* best_len;            // [B,K] generated length (0 = empty entry)
  int* best_slot;           // [B,K] which of the video's K sequence slots holds the entry's tokens
  int* best_seq;            // [B,K,S] sequence slots
  int* hist[2];             // [R,S] token history, ping-pong
};

constexpr int kSelWarps = 4;   // videos per CTA (one warp each)

// Per-warp scratch of the selection (shared memory).
struct SelScratch {
  float v[16], ns[16];
  int pk[16], nt[16];     // selection order: parent beam, token
  int np[16], tk[16];     // compacted live beams: parent beam, token
  int jslot[16], jpk[16]; // hypotheses completed this step that entered the pool: sequence slot, parent beam
  int misc[4];
};

// One warp selects for one video.  cand_val / cand_idx: the video's [K][K] per-row candidates (row k's j-th best
// log-prob and token), in global or shared memory.
// pre_score / pre_alive / pre_hist (all or none): the video's K beam scores, alive flags and token-history rows (step <= 32
// entries each), staged in shared memory by the caller before its dependency wait -- they are older than the caller's
// predecessor kernel, so their global round trips need not sit behind the selection's own.
template <int NQ>   // candidate slots per lane: K*K <= 32*NQ
__device__ __forceinline__ void beam_select_video(const BeamState& bs, const float* cand_val, const int* cand_idx,
                                                  SelScratch& sm, int b, int K, int V, int S, int step, int end_id,
                                                  float length_penalty, int* __restrict__ parent, int* __restrict__ cur_tok,
                                                  const float* pre_score = nullptr, const unsigned char* pre_alive = nullptr,
                                                  const int (*pre_hist)[32] = nullptr) {
  const int lane = threadIdx.x & 31;
  const int* hin = bs.hist[step & 1];
  int* hout = bs.hist[(step + 1) & 1];
  const int r0 = b * K, KK = K * K;

  // candidates c = k*K + j (beam k, its j-th best token): lane owns c = lane + 32*q
  float cv[NQ];
  int cf[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = lane + 32 * q;
    cv[q] = -INFINITY;
    cf[q] = 0x7fffffff;
    if (c < KK) {
      const int k = c / K;
      if (pre_alive != nullptr ? pre_alive[k] : bs.alive[r0 + k]) {
        cv[q] = (pre_score != nullptr ? pre_score[k] : bs.scores[r0 + k]) + cand_val[c];  // :211
        cf[q] = k * V + cand_idx[c];                                                      // flat index of :215
      }
    }
  }
  // top-K by (score desc, flat asc): K rounds of warp arg-max  (:215-220)
  int n_sel = 0;
  for (int sel = 0; sel < K; ++sel) {
    float bv = -INFINITY;
    int bf = 0x7fffffff, bq = -1;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (cf[q] != 0x7fffffff && (bq < 0 || cv[q] > bv || (cv[q] == bv && cf[q] < bf))) { bv = cv[q]; bf = cf[q]; bq = q; }
    float wv;
    int wf;
    warp_argmax(bq >= 0, bv, bf, wv, wf);
    if (wf == 0x7fffffff) break;            // no live candidate left (warp-uniform)
    if (bq >= 0 && bf == wf) {              // flat indices are unique: exactly one owner
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (q == bq) cf[q] = 0x7fffffff;
      sm.v[sel] = wv;
      sm.pk[sel] = wf / V;                  // :219
      sm.nt[sel] = wf % V;                  // :220
    }
    ++n_sel;
  }
  __syncwarp();
  // END bookkeeping + compaction of the live beams (:226-272), sequential in selection order
  if (lane == 0) {
    int n_alive = 0, n_jobs = 0;
    float* ps = bs.best_score + r0;
    int* pl = bs.best_len + r0;
    int* pslot = bs.best_slot + r0;
    for (int sel = 0; sel < n_sel; ++sel) {
      const int pk = sm.pk[sel], tok = sm.nt[sel];
      const float v = sm.v[sel];
      if (tok == end_id) {
        const float fin = v / (float)pow((double)(step + 1), (double)length_penalty);   // (len(new_seq)-1)**lp, :238-239
        // sorted insert; an equal score goes behind the entries already there (first maximum kept, :277-281)
        int pos = 0;
        while (pos < K && pl[pos] != 0 && !(fin > ps[pos])) ++pos;
        if (pos < K) {
          const int slot = pslot[K - 1];           // slot of the entry that drops out (an empty one while the pool fills)
          for (int i = K - 1; i > pos; --i) { ps[i] = ps[i - 1]; pl[i] = pl[i - 1]; pslot[i] = pslot[i - 1]; }
          ps[pos] = fin; pl[pos] = step + 1; pslot[pos] = slot;
          sm.jslot[n_jobs] = slot;
          sm.jpk[n_jobs] = pk;
          ++n_jobs;
        }
      } else {
        sm.np[n_alive] = pk;
        sm.ns[n_alive] = v;
        sm.tk[n_alive] = tok;
        ++n_alive;
      }
    }
    sm.misc[0] = n_alive;
    sm.misc[1] = n_jobs;
    if (n_alive == 0) bs.done[b] = 1;       // :251
  }
  __syncwarp();
  const int n_alive = sm.misc[0], n_jobs = sm.misc[1];
  // tokens of the hypotheses that entered the pool (in completion order: a slot freed and re-used within the step is
  // simply overwritten by the later job)
  for (int j = 0; j < n_jobs; ++j) {
    int* dst = bs.best_seq + (int64_t)(r0 + sm.jslot[j]) * S;
    const int* src = hin + (int64_t)(r0 + sm.jpk[j]) * S;
    for (int i = lane; i < step; i += 32) dst[i] = src[i];
    if (lane == 0) dst[step] = end_id;
    __syncwarp();
  }
  // token histories of the kept beams: all loads first, then the stores (hin / hout may alias for the compiler, and a
  // load -> store -> load chain per beam costs one memory round trip each)
  for (int i0 = 0; i0 < step; i0 += 32) {
    const int i = i0 + lane;
    int hv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      hv[k] = 0;
      if (k < K && i < step) {
        const int src = (k < n_alive) ? r0 + sm.np[k] : r0 + k;
        hv[k] = pre_hist != nullptr ? pre_hist[src - r0][i] : hin[(int64_t)src * S + i];
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K && i < step) hout[(int64_t)(r0 + k) * S + i] = hv[k];
  }
  for (int k = 0; k < K; ++k) {
    const int r = r0 + k;
    const bool live = k < n_alive;
    const int src = live ? r0 + sm.np[k] : r;         // dead slot: keeps computing on benign inputs
    const int tok = live ? sm.tk[k] : end_id;
    if (lane == 0) {
      hout[(int64_t)r * S + step] = tok;
      bs.scores[r] = live ? sm.ns[k] : -INFINITY;
      bs.alive[r] = live ? 1 : 0;
      parent[r] = src;
      cur_tok[r] = tok;
    }
  }
}

__global__ void __launch_bounds__(kSelWarps * 32) beam_select_kernel(
    BeamState bs, const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int B, int K, int V, int S, int step,
    int end_id, float length_penalty, int* __restrict__ parent /*[R]*/, int* __restrict__ cur_tok /*[R]*/) {
  __shared__ SelScratch scratch[kSelWarps];
  const int w = threadIdx.x >> 5;
  const int b = blockIdx.x * kSelWarps + w;
  if (b >= B) return;                       // warp-uniform
  if (K * K <= 32)
    beam_select_video<1>(bs, cand_val + (int64_t)b * K * K, cand_idx + (int64_t)b * K * K, scratch[w], b, K, V, S, step, end_id,
                         length_penalty, parent, cur_tok);
  else
    beam_select_video<8>(bs, cand_val + (int64_t)b * K * K, cand_idx + (int64_t)b * K * K, scratch[w], b, K, V, S, step, end_id,
                         length_penalty, parent, cur_tok);
}

// One warp moves row r's next-step inputs into place: the parent row's new (h, c) of every layer into this row's
// GEMM operands (video_captioning_model.py:247-249,269-272) and the embedding of the next token (decoder.py:130).
template <class ActT>
__device__ __forceinline__ void reorder_row_warp(const DecState<ActT>& st, int64_t r, int64_t p, int tok, int V, int lane) {
  for (int l = 0; l < st.L; ++l) {
    const ActT* hs = st.h_new[l] + p * st.H;
    const float* cs = st.c_new[l] + p * st.H;
    ActT* hd = st.x_rec[l] + r * st.x_ld[l];
    float* cd = st.c[l] + r * st.H;
    for (int u = lane * 4; u < st.H; u += 128) {
      float hv[4];
      load4(hs + u, hv);
      store4(hd + u, hv);
      *reinterpret_cast<float4*>(cd + u) = *reinterpret_cast<const float4*>(cs + u);
    }
  }
  tok = min(max(tok, 0), V - 1);
  const ActT* e = st.emb_table + (int64_t)tok * st.E;
  ActT* ed = st.emb_dst + r * st.emb_ld;
  for (int i = lane * 4; i < st.E; i += 128) {
    float ev[4];
    load4(e + i, ev);
    store4(ed + i, ev);
  }
}

// ---------------------------------------------------------------- fused selection from the vocab GEMM's statistics
// bf16 mode.  The vocabulary GEMM's epilogue (gemm_tc.cuh, STATS) leaves, per row, the maximum of every
// 32-column chunk of the logits and a (max, sum exp) pair per 128 columns.  Every element of a row's top-K
// (order: value desc, vocabulary index asc) lies in one of the K chunks with the largest maxima (ties: lower
// chunk first): if it did not, K chunk maxima -- K distinct elements -- would precede it.  So one warp per row
//   1. merges the log-sum-exp partials                               (video_captioning_model.py:209)
//   2. picks the K best chunks from the nc maxima,
//   3. reads those K x 128 bytes of logits and takes the exact top-K (:215, first half),
// and then warp 0 of the CTA (one CTA per video) runs the per-video selection (:211-272) on the K x K
// candidates through shared memory.  Greedy mode (K = 1, decoder.py:269): the top-1 is the next token.
// HBM traffic per row: (nc + 2 np) * 4 + K * 128 bytes instead of 4 * V.
// KMAX: compile-time bound of the beam size (loops are unrolled to it); MAXCL: chunk maxima per lane held in
// registers (nc <= 32 * MAXCL).
// All videos of a 1024-video step are resident at once when 7 CTAs fit an SM (K <= 5: 160 threads, <= 56 registers); with 6
// the step ran as two waves of the kernel's own latency (ncu: launch__waves_per_multiprocessor 1.15, 25 us).
template <int KMAX, int MAXCL>
__global__ void __launch_bounds__(KMAX * 32, KMAX <= 5 ? 7 : 1) select_fused_kernel(BeamState bs, const float* logits, int64_t ld,
                                                                 const float* cmax, const float2* part,
                                                                 int nc, int np, int B, int K, int V, int S, int step, int end_id,
                                                                 float length_penalty, int* __restrict__ parent,
                                                                 int* __restrict__ cur_tok, int greedy, int* __restrict__ tokens_out,
                                                                 int* __restrict__ rowthr, const DecState<bf16> st, int do_reorder) {
  __shared__ float s_cv[KMAX * KMAX];
  __shared__ int s_ci[KMAX * KMAX];
  __shared__ SelScratch scratch;
  __shared__ float s_score[KMAX];
  __shared__ unsigned char s_alive[KMAX];
  __shared__ int s_hist[KMAX][32];
  __shared__ float s_cand_v[KMAX][64];
  __shared__ int s_cand_i[KMAX][64];
  const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;     // warp k <-> beam row k of video b
  const int b = blockIdx.x;
  const int64_t r = (int64_t)b * K + k;
  constexpr float kL2e = 1.4426950408889634f;
  // beam state written by the PREVIOUS step's selection (older than the vocabulary GEMM in front of this kernel): loaded
  // before the dependency wait, staged in shared memory for warp 0's per-video selection
  const bool pre = !greedy && step <= 32;
  int hpre = 0;
  float spre = 0.f;
  unsigned char apre = 0;
  if (pre) {
    if (lane < step) hpre = bs.hist[step & 1][r * S + lane];
    if (lane == 0) { spre = bs.scores[r]; apre = bs.alive[r]; }
  }
  pdl_wait();                 // logits / cmax / part come from the vocabulary GEMM just before (read with ld.cg: PDL, common.cuh)
  pdl_launch_dependents();
  // The GEMM's shared pruning threshold of this row (largest "K-th best chunk maximum" any of its epilogue threads saw: a lower
  // bound of the true K-th best chunk maximum), read before it is reset for the next step.  Chunks below it cannot be among the K
  // best, so instead of K arg-max rounds over all nc maxima (10 or 32 registers per lane live across the rounds: 113 registers and
  // 2.3 waves at V = 30k) the maxima are streamed once, the 30-40 candidates at or above the threshold are compacted into shared
  // memory (two per lane) and the K rounds run on those.  More than 64 candidates (threshold off, many ties): the rounds re-scan
  // the maxima from L2, each picking the next chunk after the previous winner in (value desc, index asc) order.
  // Up to 320 chunks (V <= 10k) the register form is the faster one (20 vs 27 us per step at V = 10k; 58 -> 49 us at V = 30k).
  constexpr bool kRegs = MAXCL <= 10;
  const float thr_row = (rowthr != nullptr) ? key2f(__ldcg(rowthr + r)) : -INFINITY;
  __syncwarp();
  if (lane == 0 && rowthr != nullptr) rowthr[r] = (int)0x80808080;

  // chunk maxima first (the longest dependent chain starts here)
  float* cand_v = s_cand_v[k];
  int* cand_i = s_cand_i[k];
  int n_cand = 0;
  float cvr[kRegs ? MAXCL : 1];
  if constexpr (kRegs) {
#pragma unroll
    for (int i = 0; i < MAXCL; ++i) {
      const int c = lane + 32 * i;
      cvr[i] = (c < nc) ? __ldcg(cmax + r * nc + c) : -INFINITY;
    }
  } else {
    constexpr int G = 8;                                  // maxima per lane in flight at a time
    static_assert(MAXCL % G == 0, "chunk groups");
#pragma unroll 1
    for (int i0 = 0; i0 < MAXCL; i0 += G) {
      if (32 * i0 >= nc) break;
      float cv[G];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int c = lane + 32 * (i0 + i);
        cv[i] = (c < nc) ? __ldcg(cmax + r * nc + c) : -INFINITY;
      }
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const bool is = cv[i] >= thr_row && cv[i] > -INFINITY;
        const unsigned bm = __ballot_sync(0xffffffffu, is);
        if (is) {
          const int pos = n_cand + __popc(bm & ((1u << lane) - 1u));
          if (pos < 64) { cand_v[pos] = cv[i]; cand_i[pos] = lane + 32 * (i0 + i); }
        }
        n_cand += __popc(bm);
      }
    }
    __syncwarp();
  }
  if (pre) {
    s_hist[k][lane] = hpre;
    if (lane == 0) { s_score[k] = spre; s_alive[k] = apre; }
  }
  // 1. log-sum-exp of the row
  float lse;
  {
    float2 p[3];
    float m = -1e30f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int j = lane + 32 * i;
      p[i] = (j < np) ? __ldcg(part + r * np + j) : make_float2(-1e30f, 0.f);
      m = fmaxf(m, p[i].x);
    }
    for (int j = lane + 96; j < np; j += 32) m = fmaxf(m, __ldcg(part + r * np + j).x);
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) sum += p[i].y * exp2f((p[i].x - m) * kL2e);
    for (int j = lane + 96; j < np; j += 32) {
      const float2 pp = __ldcg(part + r * np + j);
      sum += pp.y * exp2f((pp.x - m) * kL2e);
    }
    sum = warp_sum(sum);
    lse = m + logf(sum);
  }
  // 2. K best chunks; 3a. this lane's element of each
  float val[KMAX];
  int cidx[KMAX];
  const bool fast = n_cand <= 64;             // (warp-uniform)
  float v0 = -INFINITY, v1 = -INFINITY;
  int i0 = 0x7fffffff, i1 = 0x7fffffff;
  if (!kRegs && fast) {
    if (lane < n_cand) { v0 = cand_v[lane]; i0 = cand_i[lane]; }
    if (lane + 32 < n_cand) { v1 = cand_v[lane + 32]; i1 = cand_i[lane + 32]; }
  }
  float pv = INFINITY;                        // previous winner (slow path)
  int pi = -1;
#pragma unroll
  for (int sel = 0; sel < KMAX; ++sel) {
    val[sel] = -INFINITY;
    cidx[sel] = 0x7fffffff;
    if (sel < K) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      if constexpr (kRegs) {
#pragma unroll
        for (int i = 0; i < MAXCL; ++i)
          if (cvr[i] > bv) { bv = cvr[i]; bi = lane + 32 * i; }     // ascending i: the first maximum has the lowest chunk index
      } else if (fast) {
        // candidates were compacted in ascending chunk order per maxima pass, not globally: compare indices explicitly
        if (i0 != 0x7fffffff) { bv = v0; bi = i0; }
        if (i1 != 0x7fffffff && (v1 > bv || (v1 == bv && i1 < bi))) { bv = v1; bi = i1; }
      } else {
        for (int c = lane; c < nc; c += 32) {
          const float v = __ldcg(cmax + r * nc + c);
          const bool after = v < pv || (v == pv && c > pi);          // not selected yet
          if (after && v > -INFINITY && (v > bv || (v == bv && c < bi))) { bv = v; bi = c; }
        }
      }
      float wv;
      int wi;
      warp_argmax(bi != 0x7fffffff, bv, bi, wv, wi);
      if constexpr (kRegs) {
#pragma unroll
        for (int i = 0; i < MAXCL; ++i)
          if (lane + 32 * i == wi) cvr[i] = -INFINITY;
      } else if (fast) {
        if (i0 == wi) i0 = 0x7fffffff;
        if (i1 == wi) i1 = 0x7fffffff;
      } else {
        pv = wv;
        pi = wi;
      }
      if (wi != 0x7fffffff) {
        const int col = wi * 32 + lane;
        cidx[sel] = col;
        if (col < V) val[sel] = __ldcg(logits + r * ld + col);
      }
    }
  }
  // 3b. exact top-K of the K x 32 candidates: K rounds of (lane-local best, warp arg-max, remove)
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j < K) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
#pragma unroll
      for (int sel = 0; sel < KMAX; ++sel)
        if (val[sel] > bv || (val[sel] == bv && val[sel] != -INFINITY && cidx[sel] < bi)) { bv = val[sel]; bi = cidx[sel]; }
      float wv;
      int wi;
      warp_argmax(bi != 0x7fffffff, bv, bi, wv, wi);
#pragma unroll
      for (int sel = 0; sel < KMAX; ++sel)
        if (cidx[sel] == wi) val[sel] = -INFINITY;
      if (lane == 0) {
        s_cv[k * K + j] = wv - lse;          // log_softmax value of the j-th best token of row k
        s_ci[k * K + j] = wi;
      }
    }
  }
  if (greedy) {
    __syncwarp();
    const int tok = s_ci[k * K];
    if (lane == 0) {
      cur_tok[r] = tok;
      tokens_out[r * S + step] = tok;
    }
    if (do_reorder) reorder_row_warp(st, r, r, tok, V, lane);
    return;
  }
  __syncthreads();
  if (k == 0)
    beam_select_video<(KMAX * KMAX + 31) / 32>(bs, s_cv, s_ci, scratch, b, K, V, S, step, end_id, length_penalty, parent, cur_tok,
                                               pre ? s_score : nullptr, pre ? s_alive : nullptr, pre ? s_hist : nullptr);
  if (!do_reorder) return;
  // reorder fused into the selection: warp k moves row k of this video (parent / token were written by warp 0 above)
  __syncthreads();
  reorder_row_warp(st, r, (int64_t)parent[r], cur_tok[r], V, lane);
}

// host-side dispatch on (beam size, chunk count)
inline int launch_select_fused(BeamState bs, const float* logits, int64_t ld, const float* cmax, const float2* part, int nc, int np,
                               int B, int K, int V, int S, int step, int end_id, float lp, int* parent, int* cur_tok, int greedy,
                               int* tokens_out, int* rowthr, const DecState<bf16>& st, int do_reorder, cudaStream_t s) {
#define VC_SEL(KM, CL) VC_CUDA(launch_pdl(select_fused_kernel<KM, CL>, dim3(B), dim3(K * 32), 0, s, bs, logits, ld, cmax, part, nc, np, B, K, V, S, step, end_id, lp, parent, cur_tok, greedy, tokens_out, rowthr, st, do_reorder))
#define VC_SEL_K(CL)                     \
  do {                                   \
    if (K == 1) VC_SEL(1, CL);           \
    else if (K <= 3) VC_SEL(3, CL);      \
    else if (K <= 5) VC_SEL(5, CL);      \
    else if (K <= 8) VC_SEL(8, CL);      \
    else VC_SEL(16, CL);                 \
  } while (0)
  VC_CHECK(K >= 1 && K <= 16 && nc <= 1024, "fused selection: K=%d nc=%d out of range", K, nc);
  if (nc <= 320) VC_SEL_K(10);
  else VC_SEL_K(32);
#undef VC_SEL_K
#undef VC_SEL
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ---------------------------------------------------------------- reorder (+ next-step embedding)
// video_captioning_model.py:247-249,269-272 (clone/cat of the kept (h,c) columns) and decoder.py:130 for
// the next step.  One CTA per row: gathers the parent row's new state into this row's GEMM operands.
template <class ActT>
__global__ void __launch_bounds__(128) reorder_embed_kernel(DecState<ActT> st, const int* parent, const int* cur_tok, int V) {
  const int r = blockIdx.x;
  pdl_wait();                 // parent / cur_tok come from the selection kernel just before (PDL, common.cuh)
  pdl_launch_dependents();
  const int p = parent ? __ldcg(parent + r) : r;
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
  if (st.ctx_src != nullptr) {
    const ActT* xs = st.ctx_src + (int64_t)p * st.H;
    ActT* xd = st.ctx_dst + (int64_t)r * st.ctx_ld;
    for (int u = threadIdx.x * 4; u < st.H; u += blockDim.x * 4) {
      float xv[4];
      load4(xs + u, xv);
      store4(xd + u, xv);
    }
  }
  int tok = __ldcg(cur_tok + r);
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
  const int r0 = b * K;
  if (bs.best_len[r0] > 0) { n = bs.best_len[r0]; src = bs.best_seq + (int64_t)(r0 + bs.best_slot[r0]) * S; }
  else { n = steps_run; src = hist + (int64_t)r0 * S; }
  out_tokens[(int64_t)b * (S + 1)] = start_id;
  for (int i = 0; i < S; ++i) out_tokens[(int64_t)b * (S + 1) + 1 + i] = (i < n) ? src[i] : start_id;
  out_len[b] = n + 1;
  if (out_score) out_score[b] = (bs.best_len[r0] > 0) ? bs.best_score[r0] : bs.scores[r0];
}

// n-best list of a finished beam decode (the "multiple hypotheses" inference/predictor.py:353 asks for): the video's
// completed hypotheses in pool order (normalised score descending), then the beams still live after `steps_run` steps in
// beam order (raw score descending) with score / steps_run^length_penalty; entries beyond what exists have length 0,
// score -inf and START tokens.  One thread per (video, entry).
__global__ void beam_nbest_kernel(BeamState bs, int B, int K, int S, int steps_run, int N, int start_id, float length_penalty,
                                  int* __restrict__ out_tokens /*[B,N,S+1]*/, int* __restrict__ out_len /*[B,N]*/,
                                  float* __restrict__ out_score /*[B,N]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, j = i - b * N, r0 = b * K;
  int n_done = 0;
  while (n_done < K && bs.best_len[r0 + n_done] > 0) ++n_done;
  int n_live = 0;
  while (n_live < K && bs.alive[r0 + n_live]) ++n_live;      // live beams are compacted to the front
  int n = 0;
  float sc = -INFINITY;
  const int* src = nullptr;
  if (j < n_done && j < K) {
    n = bs.best_len[r0 + j];
    sc = bs.best_score[r0 + j];
    src = bs.best_seq + (int64_t)(r0 + bs.best_slot[r0 + j]) * S;
  } else if (j - n_done < n_live) {
    const int k = j - n_done;
    n = steps_run;
    sc = bs.scores[r0 + k] / (float)pow((double)steps_run, (double)length_penalty);
    src = bs.hist[steps_run & 1] + (int64_t)(r0 + k) * S;
  }
  int* dst = out_tokens + (int64_t)i * (S + 1);
  dst[0] = start_id;
  for (int t = 0; t < S; ++t) dst[1 + t] = (t < n) ? src[t] : start_id;
  out_len[i] = (n > 0) ? n + 1 : 0;
  out_score[i] = sc;
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
