/* vc_b200.h -- C ABI of the B200-native caption-generation path.
 *
 * The reference (angadbawa/Video-Captioning) is pure Python: it has no FFI/operator interface, its
 * boundary is the Python class API VideoCaptioningModel.generate / VideoCaptionPredictor
 * (SURVEY.md section 8b).  This header is the C ABI a drop-in for that path binds to; the Python shim
 * in video-captioning_b200/ (same class names, parameter names and state_dict layout as the reference)
 * calls it through ctypes.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions: plain C, raw device pointers + cudaStream_t, caller-allocated outputs and workspace,
 * int status codes (0 = ok; no exceptions or aborts cross the ABI), thread-compatible: a handle and its
 * workspace are driven by one thread at a time, different handles may be used from different threads
 * (vc_last_error is thread-local; the launch counter is atomic and the vc_profile_* bookkeeping, which is
 * process-global, is mutex-guarded; the VC_* environment switches are read with getenv and must not be
 * changed with setenv concurrently).  All float tensors are fp32 row-major; token ids are int32.  Sizes:
 * F, H, E, A multiples of 8 (64 in VC_PREC_BF16 mode: tensor-core tiles), any vocabulary size >= 4 (padded
 * internally).  Paths below are relative to the reference's src/ directory.
 */
#ifndef VC_B200_H_
#define VC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vc_model vc_model_t;
typedef void* vc_stream_t; /* cudaStream_t */

enum { VC_OK = 0, VC_ERR_INVALID = 1, VC_ERR_CUDA = 2, VC_ERR_WORKSPACE = 3, VC_ERR_STATE = 4 };

/* attention variants: attention.py:9 (Bahdanau), :76 (Luong dot/general/concat), :190 (multi-head) */
enum { VC_ATTN_BAHDANAU = 0, VC_ATTN_LUONG_DOT = 1, VC_ATTN_LUONG_GENERAL = 2, VC_ATTN_LUONG_CONCAT = 3, VC_ATTN_MULTIHEAD = 4 };
/* arithmetic: FP32 = CUDA-core fp32 everywhere (token-exact parity mode); BF16 = tcgen05 tensor cores,
 * bf16 operands / fp32 accumulation, fp32 cell state and logits */
enum { VC_PREC_FP32 = 0, VC_PREC_BF16 = 1 };
enum { VC_METHOD_GREEDY = 0, VC_METHOD_BEAM = 1 };
/* element type of a feature buffer handed to vc_generate_ex */
enum { VC_DTYPE_F32 = 0, VC_DTYPE_BF16 = 1, VC_DTYPE_F16 = 2 /* host staging only */ };

/* Mirrors the config.model.* attributes the path reads (config/config.py:13-31). */
typedef struct {
  int32_t feature_dim;    /* cnn_feature_dim */
  int32_t hidden_dim;     /* encoder_hidden_dim == decoder_hidden_dim (decoder.py:97-99 makes != irreproducible) */
  int32_t embed_dim;      /* embedding_dim */
  int32_t attn_dim;       /* attention_dim */
  int32_t vocab_size;
  int32_t enc_layers;     /* encoder_num_layers (bidirectional) */
  int32_t dec_layers;     /* decoder_num_layers */
  int32_t attention;      /* VC_ATTN_* */
  int32_t num_heads;      /* multi-head only (attention.py:193) */
  int32_t precision;      /* VC_PREC_* */
} vc_model_desc_t;

/* generate() keyword arguments: video_captioning_model.py:79-88, :135, :156-157 */
typedef struct {
  int32_t method;          /* VC_METHOD_* */
  int32_t beam_size;       /* beam only, 1..16 */
  int32_t max_length;      /* S */
  int32_t start_token_id;
  int32_t end_token_id;
  float length_penalty;    /* beam only */
  float temperature;       /* greedy only (decoder.py:265) */
  int32_t diverse_beams;   /* 0 = reference semantics (all beam scores start at 0, :194); 1 = opt-in standard beam search */
} vc_decode_params_t;

const char* vc_last_error(void);           /* thread-local message for the last non-zero status */
int vc_version(void);

/* ---- measurement hooks (bench.py).  Kernel classes of the path; every launch is counted, and between
 * vc_profile_begin/vc_profile_end each class's launches are bracketed with CUDA events on the launching
 * stream.  vc_profile_end synchronises the device and returns summed milliseconds + scope counts. */
enum {
  VC_CLS_CONVERT = 0,           /* fp32 -> bf16 feature cast (bf16 mode only) */
  VC_CLS_ENC_FEATURE_PROJ = 1,  /* encoder.py:70 */
  VC_CLS_ENC_INPUT_PROJ = 2,    /* all-timestep W_ih x, both directions */
  VC_CLS_ENC_RECURRENT = 3,     /* per-timestep h W_hh^T + fused LSTM cell */
  VC_CLS_ENC_OUTPUT_PROJ = 4,   /* encoder.py:87,96 */
  VC_CLS_ATTN_PRECOMPUTE = 5,   /* hoisted keys / K,V projections */
  VC_CLS_ATTN_QUERY_PROJ = 6,
  VC_CLS_ATTN_STEP = 7,         /* fused score+mask+softmax+context */
  VC_CLS_ATTN_OUTPUT_PROJ = 8,  /* multi-head output_linear */
  VC_CLS_DEC_LSTM = 9,          /* decoder LSTM layers: GEMM + fused cell */
  VC_CLS_DEC_CONTEXT_PROJ = 10, /* decoder.py:164-165 */
  VC_CLS_DEC_VOCAB = 11,        /* decoder.py:169 */
  VC_CLS_SELECT = 12,           /* greedy arg-max / beam log-softmax+top-k+select */
  VC_CLS_REORDER_EMBED = 13,    /* beam state gather + next-token embedding */
  VC_CLS_MISC = 14,             /* init / finalize / token feeds */
  VC_CLS_COUNT = 15
};
long long vc_launch_count(void);           /* kernels launched by this library in this process so far */
int vc_profile_begin(void);
int vc_profile_end(float* ms_per_class /*[VC_CLS_COUNT]*/, int32_t* scopes_per_class /*[VC_CLS_COUNT]*/);

/* ---- model handle: replaces VideoCaptioningModel.__init__ + load_state_dict (video_captioning_model.py:13-33,
 * inference/predictor.py:70-74).  Weights are passed under their reference state_dict keys
 *  as fp32 [numel] arrays (host or device pointers); the handle keeps its own
 * re-laid-out device copies (gate-interleaved fused LSTM matrices, bf16 casts). */
int vc_model_create(const vc_model_desc_t* desc, vc_model_t** out);
int vc_model_set_weight(vc_model_t* m, const char* state_dict_key, const float* data, int64_t numel, vc_stream_t stream);
int vc_model_finalize(vc_model_t* m, vc_stream_t stream);
void vc_model_destroy(vc_model_t* m);

/* Bytes of device workspace needed for a batch of B videos x T frames decoded with beam K for S steps. */
size_t vc_workspace_bytes(const vc_model_t* m, int32_t B, int32_t T, int32_t K, int32_t S);

/* ---- VideoEncoder.forward (models/encoder.py:52-98).  feats [B,T,F]; lengths [B] int32 or NULL
 * (NULL = no video_mask; with lengths it follows the pack_padded_sequence branch :74-82 and T_out = T).
 * enc_out [B,T,H] / enc_final [B,H] may be NULL when only the workspace copies are needed. */
int vc_encoder_forward(vc_model_t* m, const float* feats, int32_t B, int32_t T, const int32_t* lengths,
                       float* enc_out, float* enc_final, void* workspace, size_t workspace_bytes, vc_stream_t stream);

/* ---- loop-invariant attention projections, hoisted (attention.py:52, :140, :241-242).  Consumes the
 * encoder outputs left in the workspace by vc_encoder_forward. */
int vc_attn_precompute(vc_model_t* m, int32_t B, int32_t T, void* workspace, size_t workspace_bytes, vc_stream_t stream);

/* ---- CaptionDecoder.generate (decoder.py:223-289): all S steps on the device, no host sync.
 * tokens [B,S] int32 (caller truncates at the first step where every row emitted END, decoder.py:275);
 * attn_weights [B,S,T] or NULL; mask [B,T] fp32 (0 = masked) or NULL. */
int vc_decode_greedy(vc_model_t* m, int32_t B, int32_t T, const float* mask, const vc_decode_params_t* p,
                     int32_t* tokens, float* attn_weights, void* workspace, size_t workspace_bytes, vc_stream_t stream);

/* ---- VideoCaptioningModel._beam_search_generate (video_captioning_model.py:148-302) with the per-video
 * (batch-size-1) semantics the Predictor uses (predictor.py:102).  tokens [B,S+1] int32: index 0 = START,
 * right-padded with START; lengths [B] (including START); scores [B] or NULL. */
int vc_decode_beam(vc_model_t* m, int32_t B, int32_t T, const float* mask, const vc_decode_params_t* p,
                   int32_t* tokens, int32_t* lengths, float* scores, void* workspace, size_t workspace_bytes,
                   vc_stream_t stream);

/* ---- n-best list of the beam decode that last ran in `workspace` (same B, T and *p): what
 * generate_multiple_captions wants from the beam ("modify beam search to return multiple hypotheses",
 * inference/predictor.py:353; semantics of video_captioning_model.py:237-242, :274-286).  Per video: the completed
 * hypotheses by length-normalised score (descending; entry 0 is the hypothesis vc_decode_beam returned), then the beams
 * still live after max_length steps with score / max_length^length_penalty.  tokens [B,N,S+1] (index 0 = START,
 * START-padded), lengths [B,N] (incl. START; 0 = no such hypothesis), scores [B,N] (-inf there).  1 <= N <= 2*beam_size.
 * Meaningful with diverse_beams = 1 (under reference semantics all K hypotheses are copies of each other). */
int vc_beam_nbest(vc_model_t* m, int32_t B, int32_t T, const vc_decode_params_t* p, int32_t N, int32_t* tokens,
                  int32_t* lengths, float* scores, void* workspace, size_t workspace_bytes, vc_stream_t stream);

/* ---- VideoCaptioningModel.generate (video_captioning_model.py:79-125): encoder + precompute + decode.
 * tokens is [B,S] (greedy) or [B,S+1] (beam); lengths/scores/attn_weights may be NULL. */
int vc_generate(vc_model_t* m, const float* feats, int32_t B, int32_t T, const int32_t* frame_lengths,
                const float* mask, const vc_decode_params_t* p, int32_t* tokens, int32_t* lengths, float* scores,
                float* attn_weights, void* workspace, size_t workspace_bytes, vc_stream_t stream);

/* ---- feature ingest (predictor.py:101-107 `torch.FloatTensor(features).to(device)`: 1.3 MB of fp32 per video over
 * PCIe is the end-to-end bottleneck).  bf16 mode rounds the features to bf16 before the first GEMM anyway, so a
 * caller may do that rounding on the HOST for part of a batch and ship half the bytes:
 *   vc_host_pack_bf16  host: dst[i] = bf16(src[i]), round to nearest even, on `threads` host threads (no CUDA call)
 *   vc_convert_bf16    device: the same rounding for the part of the batch that crossed the link as fp32
 *   vc_generate_ex     vc_generate with the feature element type stated (VC_DTYPE_BF16 only in VC_PREC_BF16 mode)
 */
int vc_host_pack_bf16(const float* src, uint16_t* dst, size_t n, int32_t threads);
/*   vc_host_stage_rows  host: the Predictor's per-video `torch.FloatTensor(features)` + `_resize_features`
 *                       (inference/predictor.py:101-107, :292-315; data/dataset.py:124-150) for a whole batch in one pass:
 *                       row r of dst [n_rows, F] = the source frame src_rows[r] points at (host address as uint64; the frame
 *                       the linspace subsampling selects) or zeros where src_rows[r] == 0 (zero padding), converted
 *                       src_dtype -> dst_dtype on the way (VC_DTYPE_*: f32->f32, f32->bf16 RNE, f16->f16 / bf16 / f32). */
int vc_host_stage_rows(const uint64_t* src_rows, int64_t n_rows, int64_t F, int32_t src_dtype, void* dst, int32_t dst_dtype,
                       int32_t threads);
int vc_convert_bf16(const float* src_dev, void* dst_dev, int64_t n, vc_stream_t stream);
int vc_generate_ex(vc_model_t* m, const void* feats, int32_t feats_dtype, int32_t B, int32_t T, const int32_t* frame_lengths,
                   const float* mask, const vc_decode_params_t* p, int32_t* tokens, int32_t* lengths, float* scores,
                   float* attn_weights, void* workspace, size_t workspace_bytes, vc_stream_t stream);

/* ---- teacher-forced forward: VideoCaptioningModel.forward / CaptionDecoder.forward
 * (video_captioning_model.py:35-77, decoder.py:173-221).  input_tokens [B,L] int32;
 * logits [B,L,V]; attn_weights [B,L,T] or NULL; enc_out [B,T,H] or NULL ('encoder_outputs' of the
 * reference's return dict).  Workspace as for K=1, S=L. */
int vc_forward_teacher(vc_model_t* m, const float* feats, int32_t B, int32_t T, const int32_t* frame_lengths,
                       const float* mask, const int32_t* input_tokens, int32_t L, float* logits, float* attn_weights,
                       float* enc_out, void* workspace, size_t workspace_bytes, vc_stream_t stream);

/* ---- step-level entry points used by the parity tests ---------------------------------------- */
/* C[M,N] = A[M,K] . W[N,K]^T + bias  through the GEMM kernel of the given precision (fp32 FFMA or bf16
 * tcgen05; bf16 operands are rounded from the fp32 inputs into `workspace`, >= 2*(M*K+N*K) bytes).
 * Stands for every nn.Linear on the path (encoder.py:70,87,96; attention.py:52-53; decoder.py:164,169). */
int vc_linear(int32_t precision, const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N,
              int32_t K, int32_t apply_tanh, void* workspace, size_t workspace_bytes, vc_stream_t stream);

/* One attention step (attention.py forward of the model's variant) for R = B*K rows:
 * enc_out [B,T,H], hidden [R,H] (top-layer h of the previous step), mask [B,T] or NULL ->
 * context [R,H], weights [R,T] (weights may be NULL: the context-only kernels of the decode loop are taken). */
int vc_attention_step(vc_model_t* m, const float* enc_out, const float* hidden, const float* mask, int32_t B, int32_t T,
                      int32_t K, float* context, float* weights, void* workspace, size_t workspace_bytes,
                      vc_stream_t stream);

/* One beam selection step (video_captioning_model.py:209-220) on given logits [B*K,V] and running
 * scores [B*K]: returns the K selected (parent beam, token, score) triples per video. */
int vc_beam_select(const float* logits, const float* scores, int32_t B, int32_t K, int32_t V, int32_t* parent,
                   int32_t* token, float* new_scores, void* workspace, size_t workspace_bytes, vc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VC_B200_H_ */
