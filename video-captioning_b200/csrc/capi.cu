// C ABI + host-side engine of the B200 caption-generation path (see include/vc_b200.h).
//
// One translation unit: the engine is templated on the activation/operand type
//   float  -> VC_PREC_FP32: FFMA GEMMs, precise tanhf/expf (token-exact parity mode)
//   bf16   -> VC_PREC_BF16: TMA + tcgen05 GEMMs, bf16 activations, fp32 accumulation/cell state/logits
// and drives: encoder (encoder.py:52-98) -> hoisted attention projections (attention.py:52,241-242) ->
// S decode steps (decoder.py:108-171) with greedy (decoder.py:223-289) or beam
// (video_captioning_model.py:148-302) selection, all enqueued on one stream with no host sync.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "attention.cuh"
#include "attention_dot.cuh"
#include "common.cuh"
#include "decode.cuh"
#include "gemm_common.cuh"
#include "gemm_f32.cuh"
#include "gemm_tc.cuh"
#include "lstm_persistent.cuh"

namespace vc {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------- launch accounting + per-class event timing
// Every kernel launch of the path goes through a LaunchScope: it counts launches (bench.py's
// `gpu_launches`) and, when profiling is enabled (vc_profile_begin), brackets the launch with CUDA events
// recorded on the launching stream so bench.py can attribute device time to kernel classes.
// Process-global (it serves bench.py, one measuring thread); handles may still be driven from several threads: the counter
// is atomic and the event bookkeeping is serialised by a mutex.
struct Profiler {
  std::atomic<bool> on{false};
  std::atomic<long long> launches{0};
  std::mutex mu;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
};
static Profiler g_prof;

struct LaunchScope {
  cudaStream_t s;
  cudaEvent_t b = nullptr;
  LaunchScope(int cls, cudaStream_t stream, int n = 1) : s(stream) {
    g_prof.launches.fetch_add(n, std::memory_order_relaxed);
    if (g_prof.on.load(std::memory_order_relaxed)) {
      std::lock_guard<std::mutex> lock(g_prof.mu);
      cudaEvent_t a = g_prof.get();
      b = g_prof.get();
      cudaEventRecord(a, s);
      g_prof.recs.push_back({cls, a, b});
    }
  }
  ~LaunchScope() {
    if (b) cudaEventRecord(b, s);
  }
};
#define VC_SCOPE(cls) vc::LaunchScope _vc_scope_(cls, s)

// ---------------------------------------------------------------- weight preparation kernels
// dst[rmap(r), dst_col0 + c] = src[r, src_col0 + c];  rmap interleaves LSTM gates: row g*H+u -> 4u+g.
template <class OutT>
__global__ void prep_copy_kernel(OutT* __restrict__ dst, int64_t dst_ld, int dst_col0, const float* __restrict__ src,
                                 int64_t src_ld, int src_col0, int rows, int cols, int interleave_H) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    const int rr = interleave_H ? (4 * (r % interleave_H) + r / interleave_H) : r;
    dst[(int64_t)rr * dst_ld + dst_col0 + c] = from_float<OutT>(src[(int64_t)r * src_ld + src_col0 + c]);
  }
}
__global__ void prep_bias_kernel(float* __restrict__ dst, const float* __restrict__ a, const float* __restrict__ b, int n,
                                 int interleave_H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rr = interleave_H ? (4 * (i % interleave_H) + i / interleave_H) : i;
  dst[rr] = a[i] + (b ? b[i] : 0.f);
}
template <class In, class Out>
__global__ void cast_kernel(const In* __restrict__ in, Out* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = from_float<Out>(to_float(in[i]));
}
__global__ void set_tokens_kernel(int* __restrict__ cur_tok, const int* __restrict__ src, int64_t stride, int col, int R) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) cur_tok[r] = src[(int64_t)r * stride + col];
}

}  // namespace vc

using namespace vc;

// ---------------------------------------------------------------- model handle
struct vc_model {
  vc_model_desc_t d;
  std::map<std::string, std::pair<float*, int64_t>> raw;   // device fp32 copies of the reference state_dict
  std::vector<void*> owned;
  bool finalized = false;
  int num_sms = 148;
  int Vp = 0;                             // vocabulary size padded to a multiple of 4 (float4 / TMA row alignment): pad rows of the embedding and W_v are 0, pad biases -1e30, so a pad column never wins a max and adds 0 to a sum of exponentials
  bool disable_persistent_lstm = false;   // VC_DISABLE_PERSISTENT_LSTM=1: per-timestep launches (A/B testing)
  bool disable_tf32_proj = false;         // VC_DISABLE_TF32_PROJ=1: convert the features to bf16 first (A/B testing)
  int dbg_vocab = 0;                      // VC_DEBUG_VOCAB: timing experiments (gemm_tc.cuh VocabStats::dbg), results invalid
  bool disable_fused_reorder = true;      // VC_FUSED_REORDER=1: the fused selection kernel also does the reorder/embedding gather (measured slower than the separate PDL launch: 54 vs 33 + 20 us per step)
  bool feat_cvt = false;                  // VC_FEAT_CVT=1: feature projection with fp32 -> bf16 converting producer warps instead of tf32 operands
  bool early_attn = false;                // VC_EARLY_ATTN=1: the next step's attention also runs on the second stream, before the reorder
  int ctx_persistent = 0;                 // VC_CTX_PERSISTENT=1|2: context projection on the persistent 128x256-tile kernel / its CTA-pair form (A/B testing); 3: 128x192 tiles, one CTA per SM
  bool disable_lstm_merge = false;        // VC_DISABLE_LSTM_MERGE=1: the decoder's stacked LSTM layers as one launch each (A/B testing)
  bool disable_attn_gather = false;       // VC_DISABLE_ATTN_GATHER=1: reorder / embedding gather as a launch of its own instead of a warp of the attention kernel (A/B testing)
  bool disable_early_q = false;           // VC_DISABLE_EARLY_Q=1: query projection in place, after the reorder (A/B testing)
  cudaStream_t aux_stream = nullptr;      // second stream of the decode loop (early query projection)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool disable_ctx_handover = false;      // VC_DISABLE_CTX_HANDOVER=1: the context projection waits for the last LSTM GEMM as a whole (A/B testing)
  bool disable_layer_sync = false;        // VC_DISABLE_LAYER_SYNC=1: stacked decoder LSTM GEMMs in plain stream order (A/B testing)
  bool disable_vocab_handover = false;    // VC_DISABLE_VOCAB_HANDOVER=1: the vocabulary projection waits for the context projection as a whole (A/B testing)
  bool disable_shared_thr = false;        // VC_DISABLE_SHARED_THR=1: per-CTA pruning thresholds only in the vocab GEMM (A/B testing)
  bool disable_fused_select = false;      // VC_DISABLE_FUSED_SELECT=1: stream the whole logits row in the selection (A/B testing)
  // derived, operand-typed (float or bf16 according to d.precision)
  void* Wp = nullptr; float* bp = nullptr;
  void* enc_Wih[4] = {}; float* enc_bias[4] = {};
  void* enc_Whh[4][2] = {};
  void* Wo = nullptr; float* bo = nullptr;
  void* emb = nullptr;
  void* Wkey = nullptr; float* bkey = nullptr;
  void* Wval = nullptr; float* bval = nullptr;
  void* Wq = nullptr; float* bq = nullptr;
  float* vvec = nullptr; float vbias = 0.f;
  __half* vvec_h = nullptr;               // fp16 copy of vvec (bf16 mode, additive attention fast path)
  int attn_gate = 0;                      // VC_ATTN_GATE: CTAs per SM allowed in the v3 scoring phase at a time (0 = no gate; measured: no gain)
  int attn_variant = 5;                   // VC_ATTN_VARIANT: 5 = persistent warp-specialised (default), 4 = one CTA per video (v4), 3 = shuffle-reduction kernel
  bool disable_attn_v3 = false;           // VC_DISABLE_ATTN_V3=1: generic attention kernel for the additive form (A/B testing)
  void* Wao = nullptr; float* bao = nullptr;
  void* dec_W[4] = {}; float* dec_bias[4] = {};
  void* Wc = nullptr; float* bc = nullptr;
  void* Wv = nullptr; float* bv = nullptr;

  ~vc_model() {
    if (aux_stream) cudaStreamDestroy(aux_stream);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (void* p : owned) cudaFree(p);
    for (auto& kv : raw) cudaFree(kv.second.first);
  }
};

namespace {

template <class T> int dev_alloc(vc_model* m, T** out, size_t n) {
  void* p = nullptr;
  VC_CUDA(cudaMalloc(&p, n * sizeof(T) + 256));
  m->owned.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return VC_OK;
}

int get_raw(vc_model* m, const std::string& key, int64_t expect, const float** out) {
  auto it = m->raw.find(key);
  if (it == m->raw.end()) {
    set_error("state_dict key '%s' was not provided", key.c_str());
    return VC_ERR_STATE;
  }
  if (it->second.second != expect) {
    set_error("state_dict key '%s' has %lld elements, expected %lld", key.c_str(), (long long)it->second.second, (long long)expect);
    return VC_ERR_INVALID;
  }
  *out = it->second.first;
  return VC_OK;
}

template <class W>
int prep_block(cudaStream_t s, W* dst, int64_t dst_ld, int dst_col0, const float* src, int64_t src_ld, int src_col0,
               int rows, int cols, int ilv) {
  const int64_t n = (int64_t)rows * cols;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  prep_copy_kernel<W><<<blocks, 256, 0, s>>>(dst, dst_ld, dst_col0, src, src_ld, src_col0, rows, cols, ilv);
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
int prep_bias(cudaStream_t s, float* dst, const float* a, const float* b, int n, int ilv) {
  prep_bias_kernel<<<(n + 255) / 256, 256, 0, s>>>(dst, a, b, n, ilv);
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// Builds the fused / re-laid-out weights in operand type W.
template <class W>
int finalize_model(vc_model* m, cudaStream_t s) {
  const vc_model_desc_t& d = m->d;
  const int F = d.feature_dim, H = d.hidden_dim, E = d.embed_dim, A = d.attn_dim, V = d.vocab_size;
  const float* src;
  const float* src2;
  auto plain = [&](const std::string& key, int rows, int cols, void** dst) -> int {
    const float* p;
    VC_TRY(get_raw(m, key, (int64_t)rows * cols, &p));
    W* w;
    VC_TRY(dev_alloc(m, &w, (size_t)rows * cols));
    VC_TRY(prep_block<W>(s, w, cols, 0, p, cols, 0, rows, cols, 0));
    *dst = w;
    return VC_OK;
  };
  auto biasv = [&](const std::string& key, int n, float** dst) -> int {
    const float* p;
    VC_TRY(get_raw(m, key, n, &p));
    VC_TRY(dev_alloc(m, dst, (size_t)n));
    VC_TRY(prep_bias(s, *dst, p, nullptr, n, 0));
    return VC_OK;
  };
  // encoder (encoder.py:29-47)
  VC_TRY(plain("encoder.feature_projection.weight", H, F, &m->Wp));
  VC_TRY(biasv("encoder.feature_projection.bias", H, &m->bp));
  for (int l = 0; l < d.enc_layers; ++l) {
    const int in = (l == 0) ? H : 2 * H;
    W* wih;
    VC_TRY(dev_alloc(m, &wih, (size_t)8 * H * in));
    VC_TRY(dev_alloc(m, &m->enc_bias[l], (size_t)8 * H));
    for (int dir = 0; dir < 2; ++dir) {
      const std::string sfx = "_l" + std::to_string(l) + (dir ? "_reverse" : "");
      VC_TRY(get_raw(m, "encoder.lstm.weight_ih" + sfx, (int64_t)4 * H * in, &src));
      VC_TRY(prep_block<W>(s, wih + (size_t)dir * 4 * H * in, in, 0, src, in, 0, 4 * H, in, H));
      VC_TRY(get_raw(m, "encoder.lstm.bias_ih" + sfx, 4 * H, &src));
      VC_TRY(get_raw(m, "encoder.lstm.bias_hh" + sfx, 4 * H, &src2));
      VC_TRY(prep_bias(s, m->enc_bias[l] + (size_t)dir * 4 * H, src, src2, 4 * H, H));
      W* whh;
      VC_TRY(dev_alloc(m, &whh, (size_t)4 * H * H));
      VC_TRY(get_raw(m, "encoder.lstm.weight_hh" + sfx, (int64_t)4 * H * H, &src));
      VC_TRY(prep_block<W>(s, whh, H, 0, src, H, 0, 4 * H, H, H));
      m->enc_Whh[l][dir] = whh;
    }
    m->enc_Wih[l] = wih;
  }
  VC_TRY(plain("encoder.output_projection.weight", H, 2 * H, &m->Wo));
  VC_TRY(biasv("encoder.output_projection.bias", H, &m->bo));
  // decoder (decoder.py:33-59)
  const int Vp = m->Vp;
  auto padded_rows = [&](const std::string& key, int cols, void** dst) -> int {     // [V, cols] -> [Vp, cols], pad rows zero
    const float* p;
    VC_TRY(get_raw(m, key, (int64_t)V * cols, &p));
    W* w;
    VC_TRY(dev_alloc(m, &w, (size_t)Vp * cols));
    if (Vp > V) VC_CUDA(cudaMemsetAsync(w + (size_t)V * cols, 0, sizeof(W) * (size_t)(Vp - V) * cols, s));
    VC_TRY(prep_block<W>(s, w, cols, 0, p, cols, 0, V, cols, 0));
    *dst = w;
    return VC_OK;
  };
  VC_TRY(padded_rows("decoder.embedding.weight", E, &m->emb));
  switch (d.attention) {
    case VC_ATTN_BAHDANAU: {
      VC_TRY(plain("decoder.attention.encoder_projection.weight", A, H, &m->Wkey));
      VC_TRY(biasv("decoder.attention.encoder_projection.bias", A, &m->bkey));
      VC_TRY(plain("decoder.attention.decoder_projection.weight", A, H, &m->Wq));
      VC_TRY(biasv("decoder.attention.decoder_projection.bias", A, &m->bq));
      VC_TRY(biasv("decoder.attention.attention_linear.weight", A, &m->vvec));
      VC_TRY(get_raw(m, "decoder.attention.attention_linear.bias", 1, &src));
      VC_CUDA(cudaMemcpyAsync(&m->vbias, src, sizeof(float), cudaMemcpyDeviceToHost, s));
      VC_CUDA(cudaStreamSynchronize(s));
      break;
    }
    case VC_ATTN_LUONG_DOT: break;
    case VC_ATTN_LUONG_GENERAL: VC_TRY(plain("decoder.attention.linear_in.weight", H, H, &m->Wq)); break;
    case VC_ATTN_LUONG_CONCAT: {
      VC_TRY(plain("decoder.attention.linear_context.weight", A, H, &m->Wkey));
      VC_TRY(biasv("decoder.attention.linear_context.bias", A, &m->bkey));
      VC_TRY(plain("decoder.attention.linear_query.weight", A, H, &m->Wq));
      VC_TRY(biasv("decoder.attention.linear_query.bias", A, &m->bq));
      VC_TRY(biasv("decoder.attention.linear_v.weight", A, &m->vvec));
      m->vbias = 0.f;
      break;
    }
    case VC_ATTN_MULTIHEAD: {
      VC_TRY(plain("decoder.attention.key_linear.weight", H, H, &m->Wkey));
      VC_TRY(biasv("decoder.attention.key_linear.bias", H, &m->bkey));
      VC_TRY(plain("decoder.attention.value_linear.weight", H, H, &m->Wval));
      VC_TRY(biasv("decoder.attention.value_linear.bias", H, &m->bval));
      VC_TRY(plain("decoder.attention.query_linear.weight", H, H, &m->Wq));
      VC_TRY(biasv("decoder.attention.query_linear.bias", H, &m->bq));
      VC_TRY(plain("decoder.attention.output_linear.weight", H, H, &m->Wao));
      VC_TRY(biasv("decoder.attention.output_linear.bias", H, &m->bao));
      break;
    }
    default: set_error("unknown attention type %d", d.attention); return VC_ERR_INVALID;
  }
  if (m->vvec != nullptr && !std::is_same<W, float>::value) {
    VC_TRY(dev_alloc(m, &m->vvec_h, (size_t)A));
    cast_kernel<float, __half><<<(A + 255) / 256, 256, 0, s>>>(m->vvec, m->vvec_h, (int64_t)A);
    VC_CUDA(cudaGetLastError());
  }
  for (int l = 0; l < d.dec_layers; ++l) {
    const int in = (l == 0) ? (E + H) : H;
    const int Kc = in + H;
    const std::string sfx = "_l" + std::to_string(l);
    W* w;
    VC_TRY(dev_alloc(m, &w, (size_t)4 * H * Kc));
    VC_TRY(get_raw(m, "decoder.lstm.weight_ih" + sfx, (int64_t)4 * H * in, &src));
    VC_TRY(prep_block<W>(s, w, Kc, 0, src, in, 0, 4 * H, in, H));       // [W_ih | W_hh], gate-interleaved rows
    VC_TRY(get_raw(m, "decoder.lstm.weight_hh" + sfx, (int64_t)4 * H * H, &src));
    VC_TRY(prep_block<W>(s, w, Kc, in, src, H, 0, 4 * H, H, H));
    VC_TRY(dev_alloc(m, &m->dec_bias[l], (size_t)4 * H));
    VC_TRY(get_raw(m, "decoder.lstm.bias_ih" + sfx, 4 * H, &src));
    VC_TRY(get_raw(m, "decoder.lstm.bias_hh" + sfx, 4 * H, &src2));
    VC_TRY(prep_bias(s, m->dec_bias[l], src, src2, 4 * H, H));
    m->dec_W[l] = w;
  }
  {
    // context_projection columns are [h_top | ctx | emb] (decoder.py:157-161); our row buffer is
    // [emb | ctx | h0_prev | h_top] read as [emb | ctx | h_top], so permute the columns once here.
    const int Kc = 2 * H + E;
    W* w;
    VC_TRY(dev_alloc(m, &w, (size_t)H * Kc));
    VC_TRY(get_raw(m, "decoder.context_projection.weight", (int64_t)H * Kc, &src));
    VC_TRY(prep_block<W>(s, w, Kc, 0, src, Kc, 2 * H, H, E, 0));       // emb
    VC_TRY(prep_block<W>(s, w, Kc, E, src, Kc, H, H, H, 0));           // ctx
    VC_TRY(prep_block<W>(s, w, Kc, E + H, src, Kc, 0, H, H, 0));       // h_top
    m->Wc = w;
    VC_TRY(biasv("decoder.context_projection.bias", H, &m->bc));
  }
  VC_TRY(padded_rows("decoder.output_projection.weight", H, &m->Wv));
  {
    const float* p;
    VC_TRY(get_raw(m, "decoder.output_projection.bias", V, &p));
    VC_TRY(dev_alloc(m, &m->bv, (size_t)Vp));
    if (Vp > V) {
      const float pad[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
      VC_CUDA(cudaMemcpyAsync(m->bv + V, pad, sizeof(float) * (size_t)(Vp - V), cudaMemcpyHostToDevice, s));
    }
    VC_TRY(prep_bias(s, m->bv, p, nullptr, V, 0));
  }
  VC_CUDA(cudaStreamSynchronize(s));
  m->finalized = true;
  return VC_OK;
}

// ---------------------------------------------------------------- workspace carving
struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
  template <class T> T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

template <class ActT>
struct WS {
  bf16* feats_bf16;
  ActT *proj, *xp, *out[2], *zero_h, *hs[2], *enc_act, *keys, *vals;
  unsigned int* flags;
  float *cst, *final_f32;
  ActT *Z, *XL[4], *Hn[4], *ctx_pre, *O;
  float *C[4], *Cn[4], *Q, *logits, *cand_val, *scores, *best_score;
  float* vs_cmax;        // vocab GEMM statistics (gemm_tc.cuh VocabStats): [R, nc] chunk maxima
  float2* vs_part;       // [R, np] log-sum-exp partials
  unsigned int* dec_sync; // [dec_layers][ceil(R/128)] tile-level hand-over counters between stacked decoder LSTM GEMMs
  int* vs_rowthr;        // [R] shared pruning threshold of a row (ordered-int key of a float), reset by the selection kernel
  int *cand_idx, *parent, *cur_tok, *done, *best_len, *best_slot, *best_seq, *hist[2];
  unsigned char* alive;
  size_t total;
};

template <class ActT>
WS<ActT> carve(const vc_model_desc_t& d, void* base, int B, int T, int K, int S) {
  WS<ActT> w;
  memset(&w, 0, sizeof(w));
  Carver c(base);
  const size_t F = d.feature_dim, H = d.hidden_dim, E = d.embed_dim, A = d.attn_dim, V = align_up((size_t)d.vocab_size, 4);   // vc_model::Vp
  const size_t BT = (size_t)B * T, R = (size_t)B * K;
  if (!std::is_same<ActT, float>::value) w.feats_bf16 = c.take<bf16>(BT * F);
  w.proj = c.take<ActT>(BT * H);
  w.xp = c.take<ActT>(BT * 8 * H);
  w.out[0] = c.take<ActT>(BT * 2 * H);
  w.out[1] = c.take<ActT>(BT * 2 * H);
  w.zero_h = c.take<ActT>((size_t)B * 2 * H);
  w.hs[0] = c.take<ActT>((size_t)B * 2 * H);
  w.hs[1] = c.take<ActT>((size_t)B * 2 * H);
  w.cst = c.take<float>((size_t)2 * B * H);
  w.flags = c.take<unsigned int>(1024);
  w.enc_act = c.take<ActT>(BT * H);
  w.final_f32 = c.take<float>((size_t)B * H);
  if (d.attention == VC_ATTN_BAHDANAU || d.attention == VC_ATTN_LUONG_CONCAT) w.keys = c.take<ActT>(BT * A);
  if (d.attention == VC_ATTN_MULTIHEAD) { w.keys = c.take<ActT>(BT * H); w.vals = c.take<ActT>(BT * H); }
  w.Z = c.take<ActT>(R * (E + 3 * H));
  for (int l = 1; l < d.dec_layers; ++l) w.XL[l] = c.take<ActT>(R * 2 * H);
  for (int l = 0; l < d.dec_layers; ++l) {
    w.Hn[l] = c.take<ActT>(R * H);
    w.C[l] = c.take<float>(R * H);
    w.Cn[l] = c.take<float>(R * H);
  }
  w.Q = c.take<float>(R * (A > H ? A : H));
  w.ctx_pre = c.take<ActT>(R * H);
  w.O = c.take<ActT>(R * H);
  w.logits = c.take<float>(R * V);
  {
    const size_t tn = (V + 255) / 256;
    w.vs_cmax = c.take<float>(R * 8 * tn);
    w.vs_part = c.take<float2>(R * 2 * tn);
    w.vs_rowthr = c.take<int>(R);
    w.dec_sync = c.take<unsigned int>((size_t)(d.dec_layers + 1) * ((R + 127) / 128));   // + context -> vocabulary projection
  }
  w.cand_val = c.take<float>(R * K);
  w.cand_idx = c.take<int>(R * K);
  w.parent = c.take<int>(R);
  w.cur_tok = c.take<int>(R);
  w.scores = c.take<float>(R);
  w.alive = c.take<unsigned char>(R);
  w.done = c.take<int>(B);
  w.best_score = c.take<float>(R);          // finished-hypothesis pool: K entries per video (decode.cuh BeamState)
  w.best_len = c.take<int>(R);
  w.best_slot = c.take<int>(R);
  w.best_seq = c.take<int>(R * S);
  w.hist[0] = c.take<int>(R * S);
  w.hist[1] = c.take<int>(R * S);
  w.total = align_up(c.off, 256);
  return w;
}

// ---------------------------------------------------------------- GEMM dispatch
// vocabulary projection with the selection statistics (bf16 mode only)
inline int gemm_vocab_stats(const GemmArgs& g, int64_t a_cols, const EpiStore<float, false, false>& e, cudaStream_t s,
                            const tc::VocabStats& vs) {
  return tc::launch_gemm_tc(g, a_cols, e, s, &vs);
}

template <class ActT, class Epi>
int gemm(const GemmArgs& g, int64_t a_cols, const Epi& e, cudaStream_t s) {
  if constexpr (std::is_same<ActT, float>::value) {
    (void)a_cols;
    return launch_sgemm(g, e, s);
  } else {
    return tc::launch_gemm_tc(g, a_cols, e, s);
  }
}

GemmArgs gargs(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A[0] = g.A[1] = A;
  g.W[0] = g.W[1] = W;
  g.lda = lda; g.ldw = ldw; g.M = M; g.N = N; g.K = K; g.nz = 1;
  g.a_col0 = 0; g.a_split = 1 << 30; g.a_skip = 0;
  g.a_origin = nullptr; g.a_origin_cols = 0;
  return g;
}
template <class OutT, bool TANH, bool P>
EpiStore<OutT, TANH, P> estore(OutT* C, int64_t ldc, const float* bias, float* C2 = nullptr, int64_t ldc2 = 0) {
  EpiStore<OutT, TANH, P> e;
  e.C[0] = e.C[1] = C; e.ldc = ldc;
  e.C2[0] = e.C2[1] = C2; e.ldc2 = ldc2;
  e.bias[0] = e.bias[1] = bias;
  return e;
}

// ---------------------------------------------------------------- encoder  (encoder.py:52-98)
template <class ActT>
int run_encoder(vc_model* m, WS<ActT>& w, const void* feats_any, int feats_dtype, int B, int T, const int* lengths,
                float* enc_out_user, float* final_user, cudaStream_t s) {
  constexpr bool P = std::is_same<ActT, float>::value;
  const vc_model_desc_t& d = m->d;
  const int F = d.feature_dim, H = d.hidden_dim;
  const int BT = B * T;
  const float* feats = reinterpret_cast<const float*>(feats_any);
  // feature projection (:70)
  bool proj_done = false;
  if (feats_dtype == VC_DTYPE_BF16) {
    // features already rounded to bf16 (host-packed ingest, vc_host_pack_bf16 / vc_convert_bf16): plain bf16 GEMM
    VC_CHECK(!P, "bf16 features need the bf16 precision mode");
    if constexpr (!P) {
      VC_SCOPE(VC_CLS_ENC_FEATURE_PROJ);
      VC_TRY((gemm<ActT>(gargs(feats_any, F, m->Wp, F, BT, H, F), F, estore<ActT, false, P>(w.proj, H, m->bp), s)));
    }
    proj_done = true;
  }
  if constexpr (!P) {
    if (!proj_done && m->feat_cvt && H >= 256 && F % 64 == 0 && (reinterpret_cast<uintptr_t>(feats) & 15) == 0) {
      // bf16 mode, fp32 features: the GEMM's producer warps round A to bf16 on the way into shared memory (bf16 MMA rate)
      VC_SCOPE(VC_CLS_ENC_FEATURE_PROJ);
      VC_TRY(tc::launch_gemm_tc_cvt(feats, F, m->Wp, F, BT, H, F, estore<bf16, false, false>(w.proj, H, m->bp), s));
      proj_done = true;
    }
    if (proj_done) {
    } else
    // bf16 mode: the tensor cores read the fp32 features (and the fp32 weight copy) as tf32 -- no conversion pass
    if (auto it = m->raw.find("encoder.feature_projection.weight"); !m->disable_tf32_proj && it != m->raw.end() && H >= 256 && F % 32 == 0 && (reinterpret_cast<uintptr_t>(feats) & 15) == 0) {
      VC_SCOPE(VC_CLS_ENC_FEATURE_PROJ);
      VC_TRY(tc::launch_gemm_tc_tf32(gargs(feats, F, it->second.first, F, BT, H, F), F, estore<bf16, false, false>(w.proj, H, m->bp), s));
      proj_done = true;
    }
  }
  if (!proj_done) {
    const void* Ain = feats;
    if (!P) {
      VC_SCOPE(VC_CLS_CONVERT);
      const int64_t n4 = (int64_t)BT * F / 4;
      convert_f32_to_bf16_kernel<<<(int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16), 256, 0, s>>>(feats, w.feats_bf16, n4);
      VC_CUDA(cudaGetLastError());
      Ain = w.feats_bf16;
    }
    VC_SCOPE(VC_CLS_ENC_FEATURE_PROJ);
    VC_TRY((gemm<ActT>(gargs(Ain, F, m->Wp, F, BT, H, F), F, estore<ActT, false, P>(w.proj, H, m->bp), s)));
  }
  VC_CUDA(cudaMemsetAsync(w.zero_h, 0, sizeof(ActT) * (size_t)B * 2 * H, s));

  const ActT* layer_in = w.proj;
  int in_dim = H;
  ActT* out = nullptr;
  int last_hs = 0;
  for (int l = 0; l < d.enc_layers; ++l) {
    // all-timestep input projections of both directions in one GEMM, N = 8H (:84, nn.LSTM W_ih x + b_ih + b_hh)
    {
      VC_SCOPE(VC_CLS_ENC_INPUT_PROJ);
      VC_TRY((gemm<ActT>(gargs(layer_in, in_dim, m->enc_Wih[l], in_dim, BT, 8 * H, in_dim), in_dim,
                         estore<ActT, false, P>(w.xp, 8 * H, m->enc_bias[l]), s)));
    }
    out = w.out[l & 1];
    if constexpr (!P) {
      // bf16 mode: persistent weights-stationary kernel, one cooperative launch per (layer, batch chunk)
      const int max_b = lengths ? 0 : tc::plstm_max_batch(H, m->num_sms);
      if (max_b > 0 && !m->disable_persistent_lstm) {
        for (int b0 = 0; b0 < B; b0 += max_b) {
          const int bc = (B - b0 < max_b) ? (B - b0) : max_b;
          VC_SCOPE(VC_CLS_ENC_RECURRENT);
          VC_TRY(tc::launch_lstm_layer_persistent(out + (size_t)b0 * T * 2 * H, w.xp + (size_t)b0 * T * 8 * H,
                                                  m->enc_Whh[l][0], m->enc_Whh[l][1], bc, T, H, w.flags, s));
        }
        layer_in = out;
        in_dim = 2 * H;
        continue;
      }
    }
    VC_CUDA(cudaMemsetAsync(w.cst, 0, sizeof(float) * (size_t)2 * B * H, s));
    if (lengths) VC_CUDA(cudaMemsetAsync(w.hs[0], 0, sizeof(ActT) * (size_t)B * 2 * H, s));
    for (int st = 0; st < T; ++st) {
      const int tz[2] = {st, T - 1 - st};
      GemmArgs g = gargs(nullptr, 0, nullptr, H, B, 4 * H, H);
      g.nz = 2;
      EpiLstm<ActT, ActT, P> e;
      memset(&e, 0, sizeof(e));
      int64_t a_cols;
      if (lengths) {
        // packed-sequence variant (:74-82): explicit state buffers, rows past their length hold state
        const int cur = st & 1, nxt = cur ^ 1;
        for (int z = 0; z < 2; ++z) {
          g.A[z] = w.hs[cur] + (size_t)z * H;
          e.h_out1[z] = w.hs[nxt] + (size_t)z * H;
          e.h_prev[z] = w.hs[cur] + (size_t)z * H;
        }
        g.lda = 2 * H; a_cols = H;
        e.h1_ld = 2 * H; e.hp_ld = 2 * H;
        e.lengths = lengths; e.t_of_z[0] = tz[0]; e.t_of_z[1] = tz[1];
        last_hs = nxt;
      } else if (st == 0) {
        g.A[0] = w.zero_h; g.A[1] = w.zero_h + H;
        g.lda = 2 * H; a_cols = H;
        g.a_origin = w.zero_h; g.a_origin_cols = 2 * H;
      } else {
        // h_{t-1} is read straight from the layer output buffer [B, T, 2H]
        g.A[0] = out + (size_t)(tz[0] - 1) * 2 * H;
        g.A[1] = out + (size_t)(tz[1] + 1) * 2 * H + H;
        g.lda = (int64_t)T * 2 * H; a_cols = H;
        g.a_origin = out; g.a_origin_cols = (int64_t)T * 2 * H;
      }
      for (int z = 0; z < 2; ++z) {
        g.W[z] = m->enc_Whh[l][z];
        e.addend[z] = w.xp + (size_t)tz[z] * 8 * H + (size_t)z * 4 * H;
        e.c_prev[z] = e.c_new[z] = w.cst + (size_t)z * H;        // cell state [B, 2H], direction-major columns
        e.h_out0[z] = out + (size_t)tz[z] * 2 * H + (size_t)z * H;
      }
      e.add_ld = (int64_t)T * 8 * H;
      e.c_ld = 2 * H;
      e.h0_ld = (int64_t)T * 2 * H;
      e.add_origin = w.xp; e.add_origin_cols = (int64_t)T * 8 * H;
      e.c_origin_in = e.c_origin_out = w.cst; e.c_tma_cols = 2 * H;
      e.h0_origin = out; e.h0_origin_cols = (int64_t)T * 2 * H;
      VC_SCOPE(VC_CLS_ENC_RECURRENT);
      VC_TRY((gemm<ActT>(g, a_cols, e, s)));
    }
    layer_in = out;
    in_dim = 2 * H;
  }
  // output projection over all frames (:87); fp32 copy to the caller if requested
  {
    VC_SCOPE(VC_CLS_ENC_OUTPUT_PROJ);
    VC_TRY((gemm<ActT>(gargs(out, 2 * H, m->Wo, 2 * H, BT, H, 2 * H), 2 * H, estore<ActT, false, P>(w.enc_act, H, m->bo), s)));
  }
  if (enc_out_user) {
    VC_SCOPE(VC_CLS_MISC);
    cast_kernel<ActT, float><<<1184, 256, 0, s>>>(w.enc_act, enc_out_user, (int64_t)BT * H);
    VC_CUDA(cudaGetLastError());
  }
  // final state [h_fwd(T-1) ; h_bwd(0)] through the same W_o (:92-96)
  VC_SCOPE(VC_CLS_ENC_OUTPUT_PROJ);
  if (lengths) {
    VC_TRY((gemm<ActT>(gargs(w.hs[last_hs], 2 * H, m->Wo, 2 * H, B, H, 2 * H), 2 * H,
                       estore<float, false, P>(w.final_f32, H, m->bo), s)));
  } else {
    GemmArgs g = gargs(out, (int64_t)T * 2 * H, m->Wo, 2 * H, B, H, 2 * H);
    g.a_col0 = (T - 1) * 2 * H;
    g.a_split = H;
    g.a_skip = -(T - 1) * 2 * H;
    VC_TRY((gemm<ActT>(g, (int64_t)T * 2 * H, estore<float, false, P>(w.final_f32, H, m->bo), s)));
  }
  if (final_user) VC_CUDA(cudaMemcpyAsync(final_user, w.final_f32, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, s));
  return VC_OK;
}

// ---------------------------------------------------------------- hoisted attention projections
template <class ActT>
int run_precompute(vc_model* m, WS<ActT>& w, int B, int T, cudaStream_t s) {
  constexpr bool P = std::is_same<ActT, float>::value;
  const vc_model_desc_t& d = m->d;
  const int H = d.hidden_dim, A = d.attn_dim, BT = B * T;
  VC_SCOPE(VC_CLS_ATTN_PRECOMPUTE);
  if (d.attention == VC_ATTN_BAHDANAU || d.attention == VC_ATTN_LUONG_CONCAT) {
    // bf16 mode stores the additive keys as fp16 (same footprint, 3 more mantissa bits) for the packed
    // half2 add/tanh/fma of the attention step
    using KeyT = typename std::conditional<P, float, __half>::type;
    VC_TRY((gemm<ActT>(gargs(w.enc_act, H, m->Wkey, H, BT, A, H), H,
                       estore<KeyT, false, P>(reinterpret_cast<KeyT*>(w.keys), A, m->bkey), s)));
  } else if (d.attention == VC_ATTN_MULTIHEAD) {
    VC_TRY((gemm<ActT>(gargs(w.enc_act, H, m->Wkey, H, BT, H, H), H, estore<ActT, false, P>(w.keys, H, m->bkey), s)));
    VC_TRY((gemm<ActT>(gargs(w.enc_act, H, m->Wval, H, BT, H, H), H, estore<ActT, false, P>(w.vals, H, m->bval), s)));
  }
  return VC_OK;
}

// ---------------------------------------------------------------- one attention step for R = B*K rows
// hq: [R, *] previous top-layer hidden state (row stride hq_ld, `hq_cols` addressable columns from hq).
template <class ActT>
int run_attention(vc_model* m, WS<ActT>& w, const ActT* hq, int64_t hq_ld, int64_t hq_cols, const float* mask, int B, int T,
                  int K, ActT* ctx, int64_t ctx_ld, float* attn_out, int64_t attn_ld, cudaStream_t s, bool q_ready = false,
                  const int* q_rows = nullptr, const RowGather* gather = nullptr) {
  constexpr bool P = std::is_same<ActT, float>::value;
  const vc_model_desc_t& d = m->d;
  const int H = d.hidden_dim, A = d.attn_dim, R = B * K;
  AttnArgs<ActT> a;
  memset(&a, 0, sizeof(a));
  a.values = w.enc_act; a.mask = mask; a.ctx = ctx; a.ctx_ld = ctx_ld; a.attn_out = attn_out; a.attn_ld = attn_ld;
  a.B = B; a.K = K; a.T_ = T; a.H = H; a.heads = 1; a.scale = 1.f;
  switch (d.attention) {
    case VC_ATTN_BAHDANAU:
    case VC_ATTN_LUONG_CONCAT: {
      if constexpr (!P) {
        const bool use_mma = m->attn_variant >= 4 && attn_additive_mma_ok(K, A, H, T) && ctx_ld % 8 == 0;
        const bool use_ws = m->attn_variant >= 5 && attn_additive_ws_ok(B, K, A, H, T, attn_out != nullptr) && ctx_ld % 8 == 0;
        if (!m->disable_attn_v3 && (use_ws || use_mma || attn_additive_fast_ok(K, A, H))) {
          // queries leave the projection GEMM as fp16 and are consumed from registers by the v3 kernel
          __half* q16 = reinterpret_cast<__half*>(w.Q);
          if (!q_ready) {     // (q_ready: run_decode projected the queries early, on the rows before the reorder: q_rows)
            VC_SCOPE(VC_CLS_ATTN_QUERY_PROJ);
            VC_TRY((gemm<ActT>(gargs(hq, hq_ld, m->Wq, H, R, A, H), hq_cols, estore<__half, false, P>(q16, A, m->bq), s)));
          }
          AttnAddArgs aa;
          aa.q_rows = q_rows;
          aa.keys = reinterpret_cast<const __half*>(w.keys); aa.q = q16; aa.v = m->vvec_h; aa.v_bias = m->vbias;
          aa.values = w.enc_act; aa.mask = mask; aa.ctx = ctx; aa.ctx_ld = ctx_ld; aa.attn_out = attn_out; aa.attn_ld = attn_ld;
          aa.B = B; aa.T = T; aa.D = A; aa.H = H;
          aa.sm_sem = m->attn_gate > 0 ? reinterpret_cast<int*>(w.flags) + 512 : nullptr;   // zeroed by run_decode
          aa.sem_limit = m->attn_gate;
          VC_SCOPE(VC_CLS_ATTN_STEP);
          if (use_ws) {
            RowGather none;
            none.n_rows = 0;
            return launch_attn_additive_ws(aa, K, s, gather != nullptr ? *gather : none);
          }
          VC_CHECK(gather == nullptr, "row gather needs the persistent additive attention kernel");
          if (use_mma) return launch_attn_additive_mma(aa, K, s);
          return launch_attn_additive(aa, K, s);
        }
      }
      {
        VC_SCOPE(VC_CLS_ATTN_QUERY_PROJ);
        VC_TRY((gemm<ActT>(gargs(hq, hq_ld, m->Wq, H, R, A, H), hq_cols, estore<float, false, P>(w.Q, A, m->bq), s)));   // attention.py:53 / :138
      }
      a.skeys = w.keys; a.q = w.Q; a.v = m->vvec; a.v_bias = m->vbias; a.D = A;
      VC_SCOPE(VC_CLS_ATTN_STEP);
      using KeyT = typename std::conditional<P, float, __half>::type;
      return launch_attn_step<ActT, KeyT, ATTN_ADDITIVE, P>(a, s);
    }
    case VC_ATTN_LUONG_DOT: {
      if constexpr (!P) {
        if (attn_dot_ws_ok(K, H, T, 1, false, attn_out != nullptr) && ctx_ld % 8 == 0 && hq_ld % 2 == 0) {
          AttnDotArgs da;
          memset(&da, 0, sizeof(da));
          da.skeys = da.values = w.enc_act; da.q_act = hq; da.q_ld = hq_ld; da.mask = mask; da.ctx = ctx; da.ctx_ld = ctx_ld;
          da.B = B; da.K = K; da.T = T; da.H = H; da.heads = 1; da.scale = 1.f;
          VC_SCOPE(VC_CLS_ATTN_STEP);
          return launch_attn_dot_ws(da, s);
        }
      }
      a.skeys = w.enc_act; a.q_act = hq; a.q_ld = hq_ld; a.D = H;
      VC_SCOPE(VC_CLS_ATTN_STEP);
      return launch_attn_step<ActT, ActT, ATTN_DOT, P>(a, s);
    }
    case VC_ATTN_LUONG_GENERAL: {
      if constexpr (!P) {
        if (attn_dot_ws_ok(K, H, T, 1, false, attn_out != nullptr) && ctx_ld % 8 == 0) {
          // streaming kernel: the projected query leaves the GEMM as bf16, like every other activation of the bf16 mode
          bf16* q16 = reinterpret_cast<bf16*>(w.Q);
          {
            VC_SCOPE(VC_CLS_ATTN_QUERY_PROJ);
            VC_TRY((gemm<ActT>(gargs(hq, hq_ld, m->Wq, H, R, H, H), hq_cols, estore<bf16, false, P>(q16, H, nullptr), s)));  // :128
          }
          AttnDotArgs da;
          memset(&da, 0, sizeof(da));
          da.skeys = da.values = w.enc_act; da.q_act = q16; da.q_ld = H; da.mask = mask; da.ctx = ctx; da.ctx_ld = ctx_ld;
          da.B = B; da.K = K; da.T = T; da.H = H; da.heads = 1; da.scale = 1.f;
          VC_SCOPE(VC_CLS_ATTN_STEP);
          return launch_attn_dot_ws(da, s);
        }
      }
      {
        VC_SCOPE(VC_CLS_ATTN_QUERY_PROJ);
        VC_TRY((gemm<ActT>(gargs(hq, hq_ld, m->Wq, H, R, H, H), hq_cols, estore<float, false, P>(w.Q, H, nullptr), s)));  // :128
      }
      a.skeys = w.enc_act; a.q = w.Q; a.D = H;
      VC_SCOPE(VC_CLS_ATTN_STEP);
      return launch_attn_step<ActT, ActT, ATTN_DOT, P>(a, s);
    }
    case VC_ATTN_MULTIHEAD: {
      a.skeys = w.keys; a.values = w.vals; a.q = w.Q; a.D = H; a.heads = d.num_heads;
      a.scale = 1.0f / sqrtf((float)(H / d.num_heads));
      a.ctx = w.ctx_pre; a.ctx_ld = H;
      bool mha_done = false;
      if constexpr (!P) {
        if (attn_dot_ws_ok(K, H, T, d.num_heads, true, attn_out != nullptr)) {
          bf16* q16 = reinterpret_cast<bf16*>(w.Q);      // (bf16 query: see the Luong general case)
          {
            VC_SCOPE(VC_CLS_ATTN_QUERY_PROJ);
            VC_TRY((gemm<ActT>(gargs(hq, hq_ld, m->Wq, H, R, H, H), hq_cols, estore<bf16, false, P>(q16, H, m->bq), s)));    // :240
          }
          AttnDotArgs da;
          memset(&da, 0, sizeof(da));
          da.skeys = w.keys; da.values = w.vals; da.q_act = q16; da.q_ld = H; da.mask = mask; da.ctx = w.ctx_pre; da.ctx_ld = H;
          da.B = B; da.K = K; da.T = T; da.H = H; da.heads = d.num_heads; da.scale = a.scale;
          VC_SCOPE(VC_CLS_ATTN_STEP);
          VC_TRY(launch_attn_dot_ws(da, s));
          mha_done = true;
        }
      }
      if (!mha_done) {
        {
          VC_SCOPE(VC_CLS_ATTN_QUERY_PROJ);
          VC_TRY((gemm<ActT>(gargs(hq, hq_ld, m->Wq, H, R, H, H), hq_cols, estore<float, false, P>(w.Q, H, m->bq), s)));    // :240
        }
        VC_SCOPE(VC_CLS_ATTN_STEP);
        VC_TRY((launch_attn_step<ActT, ActT, ATTN_MHA, P>(a, s)));
      }
      GemmArgs g = gargs(w.ctx_pre, H, m->Wao, H, R, H, H);
      EpiStore<ActT, false, P> e = estore<ActT, false, P>(ctx, ctx_ld, m->bao);                                          // :270
      VC_SCOPE(VC_CLS_ATTN_OUTPUT_PROJ);
      return gemm<ActT>(g, H, e, s);
    }
  }
  set_error("unknown attention type");
  return VC_ERR_INVALID;
}

// true when run_attention takes the additive fast path with fp16 queries in w.Q (the path whose query projection
// run_decode may issue early)
template <class ActT>
bool additive_fast_path(const vc_model* m, int B, int T, int K, int64_t ctx_ld, bool weights) {
  if (std::is_same<ActT, float>::value) return false;
  const vc_model_desc_t& d = m->d;
  if (d.attention != VC_ATTN_BAHDANAU && d.attention != VC_ATTN_LUONG_CONCAT) return false;
  const int H = d.hidden_dim, A = d.attn_dim;
  const bool use_mma = m->attn_variant >= 4 && attn_additive_mma_ok(K, A, H, T) && ctx_ld % 8 == 0;
  const bool use_ws = m->attn_variant >= 5 && attn_additive_ws_ok(B, K, A, H, T, weights) && ctx_ld % 8 == 0;
  return !m->disable_attn_v3 && (use_ws || use_mma || attn_additive_fast_ok(K, A, H));
}

// true when run_attention launches the persistent warp-specialised additive kernel (the one that can carry the row gather)
template <class ActT>
bool additive_ws_path(const vc_model* m, int B, int T, int K, int64_t ctx_ld, bool weights) {
  if (!additive_fast_path<ActT>(m, B, T, K, ctx_ld, weights)) return false;
  return m->attn_variant >= 5 && attn_additive_ws_ok(B, K, m->d.attn_dim, m->d.hidden_dim, T, weights) && ctx_ld % 8 == 0;
}

// ---------------------------------------------------------------- decode loop
enum DecodeMode { DM_GREEDY = 0, DM_BEAM = 1, DM_TEACHER = 2 };

template <class ActT>
int run_decode(vc_model* m, WS<ActT>& w, int B, int T, int K, int S, const float* mask, const vc_decode_params_t& p,
               DecodeMode mode, int* tokens_out, int* lengths_out, float* scores_out, float* attn_out,
               const int* teacher_tokens, float* teacher_logits, cudaStream_t s) {
  constexpr bool P = std::is_same<ActT, float>::value;
  const vc_model_desc_t& d = m->d;
  const int H = d.hidden_dim, E = d.embed_dim, V = m->Vp, L = d.dec_layers;     // V: padded vocabulary (pad columns are inert)
  const int Vu = d.vocab_size;                 // columns of the caller's teacher-forced logits
  const int R = B * K;
  const int64_t ZW = E + 3 * H;

  DecState<ActT> st;
  memset(&st, 0, sizeof(st));
  st.L = L; st.H = H; st.E = E;
  st.x_rec[0] = w.Z + (E + H); st.x_ld[0] = ZW;
  for (int l = 1; l < L; ++l) { st.x_rec[l] = w.XL[l] + H; st.x_ld[l] = 2 * H; }
  for (int l = 0; l < L; ++l) { st.h_new[l] = w.Hn[l]; st.c[l] = w.C[l]; st.c_new[l] = w.Cn[l]; }
  st.emb_dst = w.Z; st.emb_ld = ZW;
  st.emb_table = reinterpret_cast<const ActT*>(m->emb);

  BeamState bs;
  bs.scores = w.scores; bs.alive = w.alive; bs.done = w.done; bs.best_score = w.best_score; bs.best_len = w.best_len;
  bs.best_slot = w.best_slot; bs.best_seq = w.best_seq; bs.hist[0] = w.hist[0]; bs.hist[1] = w.hist[1];

  VC_CUDA(cudaMemsetAsync(w.flags + 512, 0, sizeof(unsigned int) * 512, s));   // attention scoring-gate counters (per SM)
  if (w.vs_rowthr != nullptr) VC_CUDA(cudaMemsetAsync(w.vs_rowthr, 0x80, sizeof(int) * (size_t)R, s));   // key of a very negative float
  // Stacked decoder LSTM GEMMs hand their h rows over tile by tile instead of kernel by kernel: layer l+1 starts on the SMs
  // layer l's last (partial) tile round leaves idle.  Counters grow monotonically over the steps (no reset inside the loop).
  unsigned int sync_arr = 0;
  const int sync_rows = (R + 127) / 128;
  if constexpr (!P) {
    if (L > 1 && !m->disable_layer_sync && w.dec_sync != nullptr) sync_arr = tc::lstm_sync_arrivals(R, 4 * H);
    if (sync_arr) VC_CUDA(cudaMemsetAsync(w.dec_sync, 0, sizeof(unsigned int) * (size_t)(L + 1) * sync_rows, s));
  }
  {
    VC_SCOPE(VC_CLS_MISC);
    decode_init_kernel<ActT><<<R, 128, 0, s>>>(st, w.final_f32, R, K, p.start_token_id, V,
                                             mode == DM_TEACHER ? teacher_tokens : nullptr, (int64_t)S, w.cur_tok, w.scores,
                                               w.alive, w.done, w.best_score, w.best_len, w.best_slot, p.diverse_beams);
  }
  VC_CUDA(cudaGetLastError());

  const ActT* hq = st.x_rec[L - 1];
  const int64_t hq_ld = st.x_ld[L - 1];
  const int64_t hq_cols = H;   // columns addressable from hq within its row

  // Early query projection (additive attention, bf16): the queries of step t+1 only need this step's top-layer h, so their
  // projection is issued right after the LSTM layers on a second stream and runs beside the context / vocabulary
  // projections and the (latency-bound) selection; the beam reorder is applied by the attention kernel as a row
  // indirection (parent).  Fork / join through events, so the pattern also records into a CUDA graph.
  bool early_q = false;
  if constexpr (!P) {
    early_q = !m->disable_early_q && S > 1 && additive_fast_path<ActT>(m, B, T, K, ZW, attn_out != nullptr);
    if (early_q && m->aux_stream == nullptr) {
      int prio_lo = 0, prio_hi = 0;
      VC_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      // lowest priority: work on this stream fills what the main stream's kernels leave free
      VC_CUDA(cudaStreamCreateWithPriority(&m->aux_stream, cudaStreamNonBlocking, prio_lo));
      VC_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
      VC_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    }
  }
  const int* q_rows = nullptr;     // rows of w.Q for the next attention step (the previous selection's parents)
  bool q_pending = false;
  // Early attention (VC_EARLY_ATTN=1): the whole attention step of t+1 -- it only depends on (video, query) -- also runs on
  // the second stream, on the rows before the reorder, into w.ctx_pre; the reorder kernel gathers the contexts by parent.
  const bool early_attn = early_q && m->early_attn && m->disable_fused_reorder;
  if (early_attn) { st.ctx_src = w.ctx_pre; st.ctx_dst = w.Z + E; st.ctx_ld = ZW; }
  bool attn_pending = false;
  // Reorder / embedding gather inside the attention kernel (attention.cuh: RowGather): with the queries projected early the
  // attention step of t+1 reads nothing the gather writes, and the XU-bound kernel leaves most of the HBM bandwidth free, so
  // the rows move while the scores are computed instead of in a launch of their own between the selection and the attention.
  RowGather rg;
  memset(&rg, 0, sizeof(rg));
  bool gather_in_attn = false;
  if constexpr (!P) {
    if (early_q && !early_attn && !m->disable_attn_gather && additive_ws_path<ActT>(m, B, T, K, ZW, attn_out != nullptr)) {
      bool ok = true;
      for (int l = 0; ok && l < L; ++l) {
        ok = row_gather_add(rg, st.h_new[l], (int64_t)H * sizeof(ActT), st.x_rec[l], st.x_ld[l] * (int64_t)sizeof(ActT), H * sizeof(ActT), false) &&
             row_gather_add(rg, st.c_new[l], (int64_t)H * sizeof(float), st.c[l], (int64_t)H * sizeof(float), H * sizeof(float), false);
      }
      ok = ok && row_gather_add(rg, st.emb_table, (int64_t)E * sizeof(ActT), st.emb_dst, st.emb_ld * (int64_t)sizeof(ActT), E * sizeof(ActT), true);
      rg.n_rows = ok ? R : 0; rg.V = V; rg.tok = w.cur_tok;
      gather_in_attn = ok;
    }
  }
  bool gather_pending = false;      // the previous step's selection still has to be applied to the decoder state

  for (int step = 0; step < S; ++step) {
    // attention on the previous step's top-layer h (decoder.py:135-138) -> ctx segment of Z
    float* aw = attn_out ? attn_out + (size_t)step * T : nullptr;
    if (!(early_attn && step > 0)) {
      if (q_pending) VC_CUDA(cudaStreamWaitEvent(s, m->ev_join, 0));
      if (gather_pending) VC_CHECK(q_pending, "row gather without early queries");   // (both are set at the end of every step)
      rg.parent = q_rows;
      VC_TRY((run_attention<ActT>(m, w, hq, hq_ld, hq_cols, mask, B, T, K, w.Z + E, ZW, aw, (int64_t)S * T, s, q_pending, q_rows,
                                  gather_pending ? &rg : nullptr)));
      q_pending = false;
      gather_pending = false;
    }
    // L-layer LSTM, one step (:152): gates = [x | h_prev] . [W_ih | W_hh]^T + (b_ih + b_hh), fused cell
    auto lstm_layer_args = [&](int l, GemmArgs& g, EpiLstm<ActT, ActT, P>& e, int64_t& lda) {
      const int in = (l == 0) ? (E + H) : H;
      const ActT* A = (l == 0) ? w.Z : w.XL[l];
      lda = (l == 0) ? ZW : 2 * H;
      g = gargs(A, lda, m->dec_W[l], in + H, R, 4 * H, in + H);
      memset(&e, 0, sizeof(e));
      e.bias[0] = e.bias[1] = m->dec_bias[l];
      e.c_prev[0] = e.c_prev[1] = w.C[l];
      e.c_new[0] = e.c_new[1] = w.Cn[l];
      e.c_ld = H;
      e.h_out0[0] = e.h_out0[1] = w.Hn[l];
      e.h0_ld = H;
      if (l < L - 1) {
        e.h_out1[0] = e.h_out1[1] = w.XL[l + 1]; e.h1_ld = 2 * H;
        e.h1_origin = w.XL[l + 1]; e.h1_origin_cols = 2 * H;
      } else {
        e.h_out1[0] = e.h_out1[1] = w.Z + (E + 2 * H); e.h1_ld = ZW;
        e.h1_origin = w.Z; e.h1_origin_cols = ZW;
      }
      e.c_origin_in = w.C[l]; e.c_origin_out = w.Cn[l]; e.c_tma_cols = H;
      if (sync_arr) {
        // (the last layer hands over to the context projection)
        e.sync_signal = (l < L - 1 || !m->disable_ctx_handover) ? w.dec_sync + (size_t)l * sync_rows : nullptr;
        e.sync_wait = (l > 0) ? w.dec_sync + (size_t)(l - 1) * sync_rows : nullptr;
        e.sync_target = sync_arr * (unsigned int)(step + 1);
      }
    };
    bool lstm_done = false;
    if constexpr (!P) {
      // two stacked layers: one launch over both layers' tiles (gemm_tc.cuh: TcArgs::dual) instead of two kernels that hand
      // tile rows over -- the same counters, but one balanced tile list, one prologue and one drain
      if (L == 2 && sync_arr != 0 && !m->disable_lstm_merge) {
        GemmArgs g0, g1;
        EpiLstm<ActT, ActT, P> e0, e1;
        int64_t lda0 = 0, lda1 = 0;
        lstm_layer_args(0, g0, e0, lda0);
        lstm_layer_args(1, g1, e1, lda1);
        VC_SCOPE(VC_CLS_DEC_LSTM);
        VC_TRY(tc::launch_gemm_tc_lstm_dual(g0, lda0, e0, g1, lda1, e1, s, &lstm_done));
      }
    }
    for (int l = 0; l < L && !lstm_done; ++l) {
      GemmArgs g;
      EpiLstm<ActT, ActT, P> e;
      int64_t lda = 0;
      lstm_layer_args(l, g, e, lda);
      VC_SCOPE(VC_CLS_DEC_LSTM);
      VC_TRY((gemm<ActT>(g, lda, e, s)));
    }
    if constexpr (!P) {
      if (early_q && step + 1 < S) {
        // queries of step+1 from this step's new top-layer h (rows before the reorder), on the second stream
        VC_CUDA(cudaEventRecord(m->ev_fork, s));
        VC_CUDA(cudaStreamWaitEvent(m->aux_stream, m->ev_fork, 0));
        {
          cudaStream_t s_main = s;
          cudaStream_t s = m->aux_stream;     // (LaunchScope below brackets the launch on the stream it runs on)
          (void)s_main;
          VC_SCOPE(VC_CLS_ATTN_QUERY_PROJ);
          VC_TRY((gemm<ActT>(gargs(w.Z + (E + 2 * H), ZW, m->Wq, H, R, d.attn_dim, H), (int64_t)H,
                             estore<__half, false, P>(reinterpret_cast<__half*>(w.Q), d.attn_dim, m->bq), s)));
        }
        if (early_attn) {
          float* aw_next = attn_out ? attn_out + (size_t)(step + 1) * T : nullptr;
          VC_TRY((run_attention<ActT>(m, w, hq, hq_ld, hq_cols, mask, B, T, K, w.ctx_pre, H, aw_next, (int64_t)S * T, m->aux_stream,
                                      /*q_ready=*/true, /*q_rows=*/nullptr)));
          attn_pending = true;
        } else {
          q_pending = true;
        }
        VC_CUDA(cudaEventRecord(m->ev_join, m->aux_stream));
      }
    }
    // tanh(context_projection([h_top ; ctx ; emb])) (:157-165), operands read in place from Z
    const int vtn = (V + 255) / 256;
    // bf16 mode: the GEMM epilogue also emits per-row chunk maxima + log-sum-exp partials, and the selection
    // reads those instead of the whole logits row (decode.cuh: select_fused_kernel)
    const bool fused_sel = !P && mode != DM_TEACHER && !m->disable_fused_select && V >= 256 && 8 * vtn <= 1024 &&
                           (mode == DM_BEAM || p.temperature == 1.0f);
    // context projection -> vocabulary projection hand-over: the 128x128-tile context kernel signals per 128-row tile, the
    // vocabulary GEMM (statistics form, single CTAs) starts on the tile rows that are complete while the rest is still running
    const bool ctx_pers = !P && (m->ctx_persistent == 1 || m->ctx_persistent == 2) && H >= 256;
    const bool vocab_handover = !P && sync_arr != 0 && fused_sel && !m->disable_vocab_handover && tc::ctx_handover_ok(R, H) && H % 128 == 0 && !ctx_pers;
    int ctx_tiles_n = H / 128;      // arrivals per 128-row tile on the context projection's hand-over counters
    {
      GemmArgs g = gargs(w.Z, ZW, m->Wc, 2 * H + E, R, H, 2 * H + E);
      g.a_split = E + H;
      g.a_skip = H;
      g.force_persistent = ctx_pers ? m->ctx_persistent : (m->ctx_persistent == 3 ? 3 : 0);
      if constexpr (!P) ctx_tiles_n = tc::store_tiles_n(g);
      if (sync_arr && !m->disable_ctx_handover && tc::ctx_handover_ok(R, H) && !ctx_pers) {
        // rows of h_top are taken over from the last LSTM layer's GEMM tile row by tile row (128x128-tile kernel only)
        g.sync_wait = w.dec_sync + (size_t)(L - 1) * sync_rows;
        g.sync_target = sync_arr * (unsigned int)(step + 1);
        g.sync_row_shift = tc::lstm_sync_row_shift(R, 4 * H);
      }
      // ... and its own rows go to the vocabulary projection the same way (flags per 128-row tile, one arrival per n-tile)
      if (vocab_handover) g.sync_signal = w.dec_sync + (size_t)L * sync_rows;
      VC_SCOPE(VC_CLS_DEC_CONTEXT_PROJ);
      VC_TRY((gemm<ActT>(g, ZW, estore<ActT, true, P>(w.O, H, m->bc), s)));
    }
    // vocabulary projection (:169)
    // teacher forcing writes the caller's [B,S,V] logits directly when the rows are 16-byte aligned (V % 4 == 0), else
    // through the padded workspace rows + a strided copy
    const bool tf_direct = mode == DM_TEACHER && Vu == V;
    float* lg = tf_direct ? teacher_logits + (size_t)step * V : w.logits;
    const int64_t ldl = tf_direct ? (int64_t)S * V : V;
    {
      VC_SCOPE(VC_CLS_DEC_VOCAB);
      if constexpr (!P) {
        if (fused_sel) {
          tc::VocabStats vs;
          memset(&vs, 0, sizeof(vs));
          vs.cmax = w.vs_cmax; vs.part = w.vs_part; vs.nc = 8 * vtn; vs.np = 2 * vtn;
          vs.rowthr = m->disable_shared_thr ? nullptr : w.vs_rowthr;
          vs.dbg = m->dbg_vocab;
          vs.topk = K <= 8 ? K : 0;     // chunks that cannot be among the row's K best are never written
          GemmArgs gv = gargs(w.O, H, m->Wv, H, R, V, H);
          if (vocab_handover) {
            gv.sync_wait = w.dec_sync + (size_t)L * sync_rows;
            gv.sync_target = (unsigned int)ctx_tiles_n * (unsigned int)(step + 1);
          }
          VC_TRY(gemm_vocab_stats(gv, H, estore<float, false, false>(lg, ldl, m->bv), s, vs));
        } else {
          VC_TRY((gemm<ActT>(gargs(w.O, H, m->Wv, H, R, V, H), H, estore<float, false, P>(lg, ldl, m->bv), s)));
        }
      } else {
        VC_TRY((gemm<ActT>(gargs(w.O, H, m->Wv, H, R, V, H), H, estore<float, false, P>(lg, ldl, m->bv), s)));
      }
    }
    if (mode == DM_TEACHER && !tf_direct) {
      VC_SCOPE(VC_CLS_MISC);
      VC_CUDA(cudaMemcpy2DAsync(teacher_logits + (size_t)step * Vu, sizeof(float) * (size_t)S * Vu, w.logits, sizeof(float) * (size_t)V,
                                sizeof(float) * (size_t)Vu, (size_t)R, cudaMemcpyDeviceToDevice, s));
    }
    // selection
    const int* parent = nullptr;
    bool reorder_done = false;
    if (fused_sel) {
      if constexpr (!P) {
        // the selection kernel also gathers the parent rows' (h, c) and the next tokens' embeddings (reorder fused in)
        VC_SCOPE(VC_CLS_SELECT);
        VC_TRY(launch_select_fused(bs, lg, ldl, w.vs_cmax, w.vs_part, 8 * vtn, 2 * vtn, B, K, V, S, step, p.end_token_id,
                                   p.length_penalty, w.parent, w.cur_tok, mode == DM_GREEDY ? 1 : 0, tokens_out, w.vs_rowthr, st,
                                   (step + 1 < S && !m->disable_fused_reorder) ? 1 : 0, s));
        reorder_done = !m->disable_fused_reorder;
      }
      if (mode == DM_BEAM) parent = w.parent;
    } else if (mode == DM_GREEDY) {
      VC_SCOPE(VC_CLS_SELECT);
      greedy_argmax_kernel<<<R, 256, 0, s>>>(lg, ldl, V, p.temperature == 1.0f ? 1.f : 0.f, p.temperature, w.cur_tok,
                                             tokens_out, S, step);
    } else if (mode == DM_BEAM) {
      vc::LaunchScope _sel(VC_CLS_SELECT, s, 2);
      beam_row_topk_kernel<P><<<R, 256, 0, s>>>(lg, ldl, V, K, w.cand_val, w.cand_idx);
      beam_select_kernel<<<(B + kSelWarps - 1) / kSelWarps, kSelWarps * 32, 0, s>>>(bs, w.cand_val, w.cand_idx, B, K, V, S, step, p.end_token_id,
                                                      p.length_penalty, w.parent, w.cur_tok);
      parent = w.parent;
    } else if (step + 1 < S) {
      VC_SCOPE(VC_CLS_MISC);
      set_tokens_kernel<<<(R + 127) / 128, 128, 0, s>>>(w.cur_tok, teacher_tokens, S, step + 1, R);
    }
    VC_CUDA(cudaGetLastError());
    q_rows = parent;     // the next step's early-projected queries are indexed by this step's parents (nullptr: identity)
    if (attn_pending) {    // the reorder below gathers the early attention's contexts
      VC_CUDA(cudaStreamWaitEvent(s, m->ev_join, 0));
      attn_pending = false;
    }
    if (step + 1 < S && !reorder_done && gather_in_attn) {
      gather_pending = true;       // applied by the next step's attention kernel
    } else if (step + 1 < S && !reorder_done) {
      VC_SCOPE(VC_CLS_REORDER_EMBED);
      VC_CUDA(launch_pdl(reorder_embed_kernel<ActT>, dim3(R), dim3(128), 0, s, st, parent, (const int*)w.cur_tok, V));
    }
  }
  if (mode == DM_BEAM) {
    VC_SCOPE(VC_CLS_MISC);
    beam_finalize_kernel<<<(B + 63) / 64, 64, 0, s>>>(bs, B, K, S, S, p.start_token_id, tokens_out, lengths_out, scores_out);
    VC_CUDA(cudaGetLastError());
  }
  return VC_OK;
}

int check_common(const vc_model* m, int B, int T, int K, int S, size_t ws_bytes, void* ws) {
  VC_CHECK(m != nullptr && m->finalized, "model handle is not finalized");
  VC_CHECK(B >= 1 && T >= 1 && K >= 1 && K <= 16 && S >= 1, "bad batch/frames/beam/length (B=%d T=%d K=%d S=%d)", B, T, K, S);
  const size_t need = vc_workspace_bytes(m, B, T, K, S);
  if (ws == nullptr || ws_bytes < need) {
    set_error("workspace too small: have %zu bytes, need %zu", ws_bytes, need);
    return VC_ERR_WORKSPACE;
  }
  return VC_OK;
}

}  // namespace

namespace {
template <class ActT>
int attention_step_impl(vc_model_t* m, const float* enc_out, const float* hidden, const float* mask, int B, int T, int K,
                        float* context, float* weights, void* ws, cudaStream_t s) {
  const int H = m->d.hidden_dim, R = B * K;
  WS<ActT> w = carve<ActT>(m->d, ws, B, T, K, 1);
  const int64_t n_enc = (int64_t)B * T * H;
  cast_kernel<float, ActT><<<1024, 256, 0, s>>>(enc_out, w.enc_act, n_enc);
  cast_kernel<float, ActT><<<256, 256, 0, s>>>(hidden, w.Hn[0], (int64_t)R * H);
  VC_CUDA(cudaGetLastError());
  VC_TRY((run_precompute<ActT>(m, w, B, T, s)));
  VC_CUDA(cudaMemsetAsync(w.flags + 512, 0, sizeof(unsigned int) * 512, s));
  VC_TRY((run_attention<ActT>(m, w, w.Hn[0], H, H, mask, B, T, K, w.O, H, weights, T, s)));
  cast_kernel<ActT, float><<<256, 256, 0, s>>>(w.O, context, (int64_t)R * H);
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
}  // namespace

// ================================================================ C ABI
extern "C" {

const char* vc_last_error(void) { return g_err; }
int vc_version(void) { return 100; }

long long vc_launch_count(void) { return g_prof.launches.load(std::memory_order_relaxed); }

int vc_profile_begin(void) {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  g_prof.on = true;
  g_prof.used = 0;
  g_prof.recs.clear();
  return VC_OK;
}

int vc_profile_end(float* ms_per_class, int32_t* launches_per_class) {
  VC_CHECK(ms_per_class != nullptr && launches_per_class != nullptr, "null argument");
  g_prof.on = false;
  VC_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_prof.mu);
  for (int c = 0; c < VC_CLS_COUNT; ++c) { ms_per_class[c] = 0.f; launches_per_class[c] = 0; }
  for (const auto& r : g_prof.recs) {
    float ms = 0.f;
    VC_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_per_class[r.cls] += ms;
    launches_per_class[r.cls] += 1;
  }
  g_prof.recs.clear();
  g_prof.used = 0;
  return VC_OK;
}

int vc_model_create(const vc_model_desc_t* desc, vc_model_t** out) {
  VC_CHECK(desc != nullptr && out != nullptr, "null argument");
  const vc_model_desc_t& d = *desc;
  VC_CHECK(d.hidden_dim > 0 && d.hidden_dim % 8 == 0 && d.feature_dim > 0 && d.feature_dim % 8 == 0 && d.embed_dim > 0 &&
               d.embed_dim % 8 == 0 && d.attn_dim > 0 && d.attn_dim % 8 == 0 && d.vocab_size >= 4,
           "F, H, E, A must be positive multiples of 8 and the vocabulary >= 4 (any size: it is padded internally): "
           "F=%d H=%d E=%d A=%d V=%d", d.feature_dim, d.hidden_dim, d.embed_dim, d.attn_dim, d.vocab_size);
  VC_CHECK(d.enc_layers >= 1 && d.enc_layers <= 4 && d.dec_layers >= 1 && d.dec_layers <= 4, "1..4 LSTM layers supported");
  VC_CHECK(d.precision == VC_PREC_FP32 || d.precision == VC_PREC_BF16, "unknown precision %d", d.precision);
  if (d.precision == VC_PREC_BF16)
    VC_CHECK(d.hidden_dim % 64 == 0 && d.feature_dim % 64 == 0 && d.embed_dim % 64 == 0,
             "bf16 tensor-core mode needs F, H, E multiples of 64 (F=%d H=%d E=%d)", d.feature_dim, d.hidden_dim, d.embed_dim);
  if (d.attention == VC_ATTN_MULTIHEAD)
    VC_CHECK(d.num_heads >= 1 && d.num_heads <= 32 && 32 % d.num_heads == 0 && d.hidden_dim % d.num_heads == 0 &&
                 d.hidden_dim % 128 == 0,
             "multi-head attention: heads=%d must divide 32 and H=%d must be a multiple of 128", d.num_heads, d.hidden_dim);
  vc_model* m = new (std::nothrow) vc_model();
  VC_CHECK(m != nullptr, "out of host memory");
  m->d = d;
  m->Vp = (d.vocab_size + 3) / 4 * 4;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
    m->num_sms = sms;
  cudaGetLastError();   // creating a handle must also work where no device is visible (CPU-side tests)
  const char* env = getenv("VC_DISABLE_PERSISTENT_LSTM");
  m->disable_persistent_lstm = env != nullptr && env[0] == '1';
  env = getenv("VC_ATTN_VARIANT");
  if (env != nullptr) m->attn_variant = atoi(env);
  env = getenv("VC_ATTN_GATE");
  if (env != nullptr) m->attn_gate = atoi(env);
  env = getenv("VC_DISABLE_ATTN_V3");
  m->disable_attn_v3 = env != nullptr && env[0] == '1';
  env = getenv("VC_DISABLE_TF32_PROJ");
  m->disable_tf32_proj = env != nullptr && env[0] == '1';
  env = getenv("VC_DEBUG_VOCAB");
  m->dbg_vocab = env != nullptr ? atoi(env) : 0;
  env = getenv("VC_FUSED_REORDER");
  m->disable_fused_reorder = !(env != nullptr && env[0] == '1');
  env = getenv("VC_FEAT_CVT");
  m->feat_cvt = env != nullptr && env[0] == '1';
  env = getenv("VC_EARLY_ATTN");
  m->early_attn = env != nullptr && env[0] == '1';
  env = getenv("VC_DISABLE_EARLY_Q");
  m->disable_early_q = env != nullptr && env[0] == '1';
  env = getenv("VC_CTX_PERSISTENT");
  m->ctx_persistent = env != nullptr ? atoi(env) : 0;
  env = getenv("VC_DISABLE_LSTM_MERGE");
  m->disable_lstm_merge = env != nullptr && env[0] == '1';
  env = getenv("VC_DISABLE_ATTN_GATHER");
  m->disable_attn_gather = env != nullptr && env[0] == '1';
  env = getenv("VC_DISABLE_CTX_HANDOVER");
  m->disable_ctx_handover = env != nullptr && env[0] == '1';
  env = getenv("VC_DISABLE_LAYER_SYNC");
  m->disable_layer_sync = env != nullptr && env[0] == '1';
  env = getenv("VC_DISABLE_VOCAB_HANDOVER");
  m->disable_vocab_handover = env != nullptr && env[0] == '1';
  env = getenv("VC_DISABLE_SHARED_THR");
  m->disable_shared_thr = env != nullptr && env[0] == '1';
  env = getenv("VC_DISABLE_FUSED_SELECT");
  m->disable_fused_select = env != nullptr && env[0] == '1';
  *out = m;
  return VC_OK;
}

int vc_model_set_weight(vc_model_t* m, const char* key, const float* data, int64_t numel, vc_stream_t stream) {
  VC_CHECK(m != nullptr && key != nullptr && data != nullptr && numel > 0, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto it = m->raw.find(key);
  float* dst = nullptr;
  if (it != m->raw.end() && it->second.second == numel) {
    dst = it->second.first;
  } else {
    if (it != m->raw.end()) { cudaFree(it->second.first); m->raw.erase(it); }
    VC_CUDA(cudaMalloc(reinterpret_cast<void**>(&dst), sizeof(float) * (size_t)numel));
    m->raw[key] = std::make_pair(dst, numel);
  }
  VC_CUDA(cudaMemcpyAsync(dst, data, sizeof(float) * (size_t)numel, cudaMemcpyDefault, s));
  m->finalized = false;
  return VC_OK;
}

int vc_model_finalize(vc_model_t* m, vc_stream_t stream) {
  VC_CHECK(m != nullptr, "null model");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  for (void* p : m->owned) cudaFree(p);
  m->owned.clear();
  if (m->d.precision == VC_PREC_FP32) return finalize_model<float>(m, s);
  return finalize_model<bf16>(m, s);
}

void vc_model_destroy(vc_model_t* m) { delete m; }

size_t vc_workspace_bytes(const vc_model_t* m, int32_t B, int32_t T, int32_t K, int32_t S) {
  if (m == nullptr) return 0;
  if (m->d.precision == VC_PREC_FP32) return carve<float>(m->d, nullptr, B, T, K, S).total;
  return carve<bf16>(m->d, nullptr, B, T, K, S).total;
}

#define VC_DISPATCH(m, expr_f32, expr_bf16) ((m)->d.precision == VC_PREC_FP32 ? (expr_f32) : (expr_bf16))

static int encoder_forward_typed(vc_model_t* m, const void* feats, int32_t feats_dtype, int32_t B, int32_t T, const int32_t* lengths,
                                 float* enc_out, float* enc_final, void* ws, size_t ws_bytes, vc_stream_t stream) {
  VC_CHECK(m != nullptr && feats != nullptr, "null argument");
  VC_CHECK(feats_dtype == VC_DTYPE_F32 || feats_dtype == VC_DTYPE_BF16, "unknown feature dtype %d", feats_dtype);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // the encoder part of the workspace does not depend on K/S; accept any workspace sized for K=1,S=1
  VC_TRY(check_common(m, B, T, 1, 1, ws_bytes, ws));
  if (m->d.precision == VC_PREC_FP32) {
    WS<float> w = carve<float>(m->d, ws, B, T, 1, 1);
    return run_encoder<float>(m, w, feats, feats_dtype, B, T, lengths, enc_out, enc_final, s);
  }
  WS<bf16> w = carve<bf16>(m->d, ws, B, T, 1, 1);
  return run_encoder<bf16>(m, w, feats, feats_dtype, B, T, lengths, enc_out, enc_final, s);
}

int vc_encoder_forward(vc_model_t* m, const float* feats, int32_t B, int32_t T, const int32_t* lengths, float* enc_out,
                       float* enc_final, void* ws, size_t ws_bytes, vc_stream_t stream) {
  return encoder_forward_typed(m, feats, VC_DTYPE_F32, B, T, lengths, enc_out, enc_final, ws, ws_bytes, stream);
}

// fp32 -> bf16 (round to nearest even) on the device: the raw part of a host-packed ingest batch
int vc_convert_bf16(const float* src, void* dst, int64_t n, vc_stream_t stream) {
  VC_CHECK(src != nullptr && dst != nullptr && n >= 0 && n % 4 == 0, "vc_convert_bf16: n must be a multiple of 4");
  VC_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "vc_convert_bf16: unaligned buffers");
  if (n == 0) return VC_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t n4 = n / 4;
  VC_SCOPE(VC_CLS_CONVERT);
  convert_f32_to_bf16_kernel<<<(int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16), 256, 0, s>>>(src, reinterpret_cast<bf16*>(dst), n4);
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

int vc_attn_precompute(vc_model_t* m, int32_t B, int32_t T, void* ws, size_t ws_bytes, vc_stream_t stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  VC_TRY(check_common(m, B, T, 1, 1, ws_bytes, ws));
  if (m->d.precision == VC_PREC_FP32) {
    WS<float> w = carve<float>(m->d, ws, B, T, 1, 1);
    return run_precompute<float>(m, w, B, T, s);
  }
  WS<bf16> w = carve<bf16>(m->d, ws, B, T, 1, 1);
  return run_precompute<bf16>(m, w, B, T, s);
}

int vc_decode_greedy(vc_model_t* m, int32_t B, int32_t T, const float* mask, const vc_decode_params_t* p, int32_t* tokens,
                     float* attn, void* ws, size_t ws_bytes, vc_stream_t stream) {
  VC_CHECK(p != nullptr && tokens != nullptr, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int S = p->max_length;
  VC_TRY(check_common(m, B, T, 1, S, ws_bytes, ws));
  VC_CHECK(p->temperature > 0.f, "temperature must be positive");
  VC_CHECK(p->start_token_id >= 0 && p->start_token_id < m->d.vocab_size, "start_token_id %d outside the vocabulary [0, %d)",
           p->start_token_id, m->d.vocab_size);
  if (m->d.precision == VC_PREC_FP32) {
    WS<float> w = carve<float>(m->d, ws, B, T, 1, S);
    return run_decode<float>(m, w, B, T, 1, S, mask, *p, DM_GREEDY, tokens, nullptr, nullptr, attn, nullptr, nullptr, s);
  }
  WS<bf16> w = carve<bf16>(m->d, ws, B, T, 1, S);
  return run_decode<bf16>(m, w, B, T, 1, S, mask, *p, DM_GREEDY, tokens, nullptr, nullptr, attn, nullptr, nullptr, s);
}

int vc_decode_beam(vc_model_t* m, int32_t B, int32_t T, const float* mask, const vc_decode_params_t* p, int32_t* tokens,
                   int32_t* lengths, float* scores, void* ws, size_t ws_bytes, vc_stream_t stream) {
  VC_CHECK(p != nullptr && tokens != nullptr && lengths != nullptr, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int S = p->max_length, K = p->beam_size;
  VC_TRY(check_common(m, B, T, K, S, ws_bytes, ws));
  VC_CHECK(p->start_token_id >= 0 && p->start_token_id < m->d.vocab_size, "start_token_id %d outside the vocabulary [0, %d)",
           p->start_token_id, m->d.vocab_size);
  if (m->d.precision == VC_PREC_FP32) {
    WS<float> w = carve<float>(m->d, ws, B, T, K, S);
    return run_decode<float>(m, w, B, T, K, S, mask, *p, DM_BEAM, tokens, lengths, scores, nullptr, nullptr, nullptr, s);
  }
  WS<bf16> w = carve<bf16>(m->d, ws, B, T, K, S);
  return run_decode<bf16>(m, w, B, T, K, S, mask, *p, DM_BEAM, tokens, lengths, scores, nullptr, nullptr, nullptr, s);
}

// n-best hypotheses of the beam decode that just ran in `ws` (same B, T, beam size and max_length)
int vc_beam_nbest(vc_model_t* m, int32_t B, int32_t T, const vc_decode_params_t* p, int32_t N, int32_t* tokens, int32_t* lengths,
                  float* scores, void* ws, size_t ws_bytes, vc_stream_t stream) {
  VC_CHECK(p != nullptr && tokens != nullptr && lengths != nullptr && scores != nullptr, "null argument");
  VC_CHECK(p->method == VC_METHOD_BEAM && N >= 1 && N <= 2 * p->beam_size, "vc_beam_nbest: beam decode and 1 <= N <= 2*beam_size expected");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int S = p->max_length, K = p->beam_size;
  VC_TRY(check_common(m, B, T, K, S, ws_bytes, ws));
  BeamState bs;
  if (m->d.precision == VC_PREC_FP32) {
    WS<float> w = carve<float>(m->d, ws, B, T, K, S);
    bs.scores = w.scores; bs.alive = w.alive; bs.done = w.done; bs.best_score = w.best_score; bs.best_len = w.best_len;
    bs.best_slot = w.best_slot; bs.best_seq = w.best_seq; bs.hist[0] = w.hist[0]; bs.hist[1] = w.hist[1];
  } else {
    WS<bf16> w = carve<bf16>(m->d, ws, B, T, K, S);
    bs.scores = w.scores; bs.alive = w.alive; bs.done = w.done; bs.best_score = w.best_score; bs.best_len = w.best_len;
    bs.best_slot = w.best_slot; bs.best_seq = w.best_seq; bs.hist[0] = w.hist[0]; bs.hist[1] = w.hist[1];
  }
  VC_SCOPE(VC_CLS_MISC);
  beam_nbest_kernel<<<(B * N + 127) / 128, 128, 0, s>>>(bs, B, K, S, S, N, p->start_token_id, p->length_penalty, tokens, lengths, scores);
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// NOTE on workspace layout: the encoder-side buffers precede the decode-side ones and their sizes do not
// depend on (K,S), so carving with (K,S) = (1,1) and with the real (K,S) yields identical encoder offsets.
int vc_generate(vc_model_t* m, const float* feats, int32_t B, int32_t T, const int32_t* frame_lengths, const float* mask,
                const vc_decode_params_t* p, int32_t* tokens, int32_t* lengths, float* scores, float* attn, void* ws,
                size_t ws_bytes, vc_stream_t stream) {
  return vc_generate_ex(m, feats, VC_DTYPE_F32, B, T, frame_lengths, mask, p, tokens, lengths, scores, attn, ws, ws_bytes, stream);
}

int vc_generate_ex(vc_model_t* m, const void* feats, int32_t feats_dtype, int32_t B, int32_t T, const int32_t* frame_lengths,
                   const float* mask, const vc_decode_params_t* p, int32_t* tokens, int32_t* lengths, float* scores, float* attn,
                   void* ws, size_t ws_bytes, vc_stream_t stream) {
  VC_CHECK(m != nullptr && p != nullptr && feats != nullptr && tokens != nullptr, "null argument");
  VC_CHECK(p->method == VC_METHOD_GREEDY || p->method == VC_METHOD_BEAM, "Unsupported generation method: %d", p->method);
  const int K = p->method == VC_METHOD_BEAM ? p->beam_size : 1;
  VC_TRY(check_common(m, B, T, K, p->max_length, ws_bytes, ws));
  VC_TRY(encoder_forward_typed(m, feats, feats_dtype, B, T, frame_lengths, nullptr, nullptr, ws, ws_bytes, stream));
  VC_TRY(vc_attn_precompute(m, B, T, ws, ws_bytes, stream));
  if (p->method == VC_METHOD_GREEDY) return vc_decode_greedy(m, B, T, mask, p, tokens, attn, ws, ws_bytes, stream);
  VC_CHECK(lengths != nullptr, "beam decoding needs a lengths output");
  return vc_decode_beam(m, B, T, mask, p, tokens, lengths, scores, ws, ws_bytes, stream);
}

int vc_forward_teacher(vc_model_t* m, const float* feats, int32_t B, int32_t T, const int32_t* frame_lengths,
                       const float* mask, const int32_t* input_tokens, int32_t L, float* logits, float* attn,
                       float* enc_out, void* ws, size_t ws_bytes, vc_stream_t stream) {
  VC_CHECK(m != nullptr && feats != nullptr && input_tokens != nullptr && logits != nullptr, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  VC_TRY(check_common(m, B, T, 1, L, ws_bytes, ws));
  VC_TRY(vc_encoder_forward(m, feats, B, T, frame_lengths, enc_out, nullptr, ws, ws_bytes, stream));
  VC_TRY(vc_attn_precompute(m, B, T, ws, ws_bytes, stream));
  vc_decode_params_t p;
  memset(&p, 0, sizeof(p));
  p.max_length = L; p.beam_size = 1; p.temperature = 1.f; p.length_penalty = 1.f;
  if (m->d.precision == VC_PREC_FP32) {
    WS<float> w = carve<float>(m->d, ws, B, T, 1, L);
    return run_decode<float>(m, w, B, T, 1, L, mask, p, DM_TEACHER, nullptr, nullptr, nullptr, attn, input_tokens, logits, s);
  }
  WS<bf16> w = carve<bf16>(m->d, ws, B, T, 1, L);
  return run_decode<bf16>(m, w, B, T, 1, L, mask, p, DM_TEACHER, nullptr, nullptr, nullptr, attn, input_tokens, logits, s);
}

// ---------------------------------------------------------------- step-level entry points (parity tests)
int vc_linear(int32_t precision, const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
              int32_t apply_tanh, void* ws, size_t ws_bytes, vc_stream_t stream) {
  VC_CHECK(A && W && C && M > 0 && N > 0 && K > 0, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (precision == VC_PREC_FP32) {
    GemmArgs g = gargs(A, K, W, K, M, N, K);
    if (apply_tanh) return launch_sgemm(g, estore<float, true, true>(C, N, bias), s);
    return launch_sgemm(g, estore<float, false, true>(C, N, bias), s);
  }
  VC_CHECK(precision == VC_PREC_BF16, "unknown precision %d", precision);
  const size_t need = align_up(sizeof(bf16) * (size_t)M * K, 256) + sizeof(bf16) * (size_t)N * K;
  if (ws == nullptr || ws_bytes < need) {
    set_error("vc_linear(bf16): workspace too small: have %zu need %zu", ws_bytes, need);
    return VC_ERR_WORKSPACE;
  }
  bf16* a16 = reinterpret_cast<bf16*>(ws);
  bf16* w16 = reinterpret_cast<bf16*>(reinterpret_cast<uint8_t*>(ws) + align_up(sizeof(bf16) * (size_t)M * K, 256));
  cast_kernel<float, bf16><<<1024, 256, 0, s>>>(A, a16, (int64_t)M * K);
  cast_kernel<float, bf16><<<1024, 256, 0, s>>>(W, w16, (int64_t)N * K);
  VC_CUDA(cudaGetLastError());
  GemmArgs g = gargs(a16, K, w16, K, M, N, K);
  if (apply_tanh) return tc::launch_gemm_tc(g, K, estore<float, true, false>(C, N, bias), s);
  return tc::launch_gemm_tc(g, K, estore<float, false, false>(C, N, bias), s);
}


int vc_attention_step(vc_model_t* m, const float* enc_out, const float* hidden, const float* mask, int32_t B, int32_t T,
                      int32_t K, float* context, float* weights, void* ws, size_t ws_bytes, vc_stream_t stream) {
  VC_CHECK(enc_out && hidden && context, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  VC_TRY(check_common(m, B, T, K, 1, ws_bytes, ws));
  if (m->d.precision == VC_PREC_FP32) return attention_step_impl<float>(m, enc_out, hidden, mask, B, T, K, context, weights, ws, s);
  return attention_step_impl<bf16>(m, enc_out, hidden, mask, B, T, K, context, weights, ws, s);
}

int vc_beam_select(const float* logits, const float* scores, int32_t B, int32_t K, int32_t V, int32_t* parent, int32_t* token,
                   float* new_scores, void* ws, size_t ws_bytes, vc_stream_t stream) {
  VC_CHECK(logits && scores && parent && token && new_scores && B >= 1 && K >= 1 && K <= 16 && V % 4 == 0, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t R = (size_t)B * K;
  Carver c(ws);
  float* cand_val = c.take<float>(R * K);
  int* cand_idx = c.take<int>(R * K);
  BeamState bs;
  bs.scores = c.take<float>(R);
  bs.alive = c.take<unsigned char>(R);
  bs.done = c.take<int>(B);
  bs.best_score = c.take<float>(R);
  bs.best_len = c.take<int>(R);
  bs.best_slot = c.take<int>(R);
  bs.best_seq = c.take<int>(R);
  bs.hist[0] = c.take<int>(R);
  bs.hist[1] = c.take<int>(R);
  if (ws == nullptr || ws_bytes < c.off) {
    set_error("vc_beam_select: workspace too small: have %zu need %zu", ws_bytes, c.off);
    return VC_ERR_WORKSPACE;
  }
  VC_CUDA(cudaMemcpyAsync(bs.scores, scores, sizeof(float) * R, cudaMemcpyDeviceToDevice, s));
  VC_CUDA(cudaMemsetAsync(bs.alive, 1, R, s));
  VC_CUDA(cudaMemsetAsync(bs.best_len, 0, sizeof(int) * R, s));
  VC_CUDA(cudaMemsetAsync(bs.best_slot, 0, sizeof(int) * R, s));
  beam_row_topk_kernel<true><<<(int)R, 256, 0, s>>>(logits, V, V, K, cand_val, cand_idx);
  beam_select_kernel<<<(B + kSelWarps - 1) / kSelWarps, kSelWarps * 32, 0, s>>>(bs, cand_val, cand_idx, B, K, V, 1, 0, /*end_id=*/-1, 1.0f, parent, token);
  VC_CUDA(cudaGetLastError());
  VC_CUDA(cudaMemcpyAsync(new_scores, bs.scores, sizeof(float) * R, cudaMemcpyDeviceToDevice, s));
  return VC_OK;
}

}  // extern "C"
