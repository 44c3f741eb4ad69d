/* gava_sm100.h -- C ABI of libgava_sm100.so: hand-written sm_100a (B200) kernels for the
 * GaVA-CLIP / Vita-CLIP video-encoder forward and contrastive logit head.
 *
 * The reference (lisqzqng/GaVA-CLIP) is pure PyTorch and has no FFI of its own; the boundary it
 * exposes is the nn.Module API of training/VitaCLIP_model.py.  Each entry point below replaces the
 * ATen library call(s) at the cited reference lines (paths relative to the reference root).
 *
 * Conventions:
 *   - plain pointers + sizes only; every pointer is DEVICE memory owned by the caller; nothing is
 *     allocated, freed or retained by the library; outputs are caller-preallocated.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), CUDA-graph capturable.
 *   - return value: 0 = OK, >0 = cudaError_t of the launch, <0 = invalid argument / unsupported shape.
 *     gava_last_error_string() returns a thread-local description of the last failure.
 *   - bf16 = __nv_bfloat16 (2 bytes), f32 = float.  All matrices are row-major.
 */
#ifndef GAVA_SM100_H_
#define GAVA_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAVA_ABI_VERSION 4

/* epilogue modes of gava_gemm_bf16_fwd */
enum gava_epilogue_mode {
  GAVA_EPI_BIAS_BF16 = 0,      /* out(bf16) = acc + bias                                  (q/k/v proj)     */
  GAVA_EPI_BIAS_GELU_BF16 = 1, /* out(bf16) = QuickGELU(acc + bias)                       (mlp.fc1 + act)  */
  GAVA_EPI_BIAS_RESID_F32 = 2, /* out(f32)  = resid + acc + bias  (out may alias resid)   (out_proj, fc2)  */
  GAVA_EPI_BIAS_F32 = 3,       /* out(f32)  = acc + bias                                  (cls_proj, proj) */
  GAVA_EPI_PATCH_F32 = 4,      /* out(f32)[(r/group)*group_stride + r%group + row_offset]
                                    = acc + bias + pos[1 + r%group] + time_embed[(r/group)%T] (patch embed) */
  GAVA_EPI_DGELU_BF16 = 5,     /* out(bf16) = acc * QuickGELU'(aux[r, c])   (dgrad through mlp.act; aux = fc1 pre-activations) */
  /* (mode 6, a LayerNorm tail behind the fc2 GEMM, existed in ABI version 2; it measured 5 % slower end to end than the
     separate LayerNorm launch and was removed in version 3) */
  /* LayerNorm folded around a GEMM pair (ABI version 4): the PRODUCER of a residual-stream tensor also emits its bf16
     copy and per-row partial statistics, the CONSUMER contracts that raw copy with W' = W diag(ln.weight) and applies
     the normalisation after the contraction:  LN(x) W^T + b  ==  rstd (x W'^T - mean colsum(W')) + (b + W ln.bias).
     Removes the separate norm1 pass over the stream (VitaCLIP_vision_encoder_utils.py:190, `self.norm1(x)`). */
  GAVA_EPI_LNFOLD_BF16 = 7,           /* out(bf16) = rstd[r] * (acc - mean[r] * colsum[c]) + bias[c]   (consumer; colsum
                                         NULL = 0 for row-centred W')  */
  GAVA_EPI_BIAS_RESID_STATS_F32 = 8   /* mode 2, plus copy_out(bf16) = out and stats_out[r, slot] = (sum, sum sq) (producer) */
};

typedef struct gava_gemm_epilogue {
  int32_t mode;            /* enum gava_epilogue_mode */
  const float* bias;       /* [N] or NULL */
  const float* resid;      /* f32 [M, ldr], mode RESID only */
  int64_t ldr;
  void* out;               /* bf16 or f32, row pitch ldo (elements) */
  int64_t ldo;
  int32_t group;           /* PATCH: patches per frame (196) */
  int32_t group_stride;    /* PATCH: token rows per frame in `out` (197) */
  int32_t row_offset;      /* PATCH: first patch slot inside a frame (1, after cls) */
  int32_t T;               /* PATCH: frames per clip */
  const float* pos;        /* PATCH: f32 [group+1, N]  (visual.pos_embed) */
  const float* time_embed; /* PATCH: f32 [T, N]        (visual.time_embed, already resized to T) */
  const void* aux;         /* DGELU: bf16 [M, ld_aux] pre-activations; RESID (optional): bf16 [M, ld_aux] second addend */
  int64_t ld_aux;
  void* aux_out;           /* GELU (optional): also store the pre-activations acc + bias as bf16 [M, ld_aux] */
  int32_t max_ctas;        /* 0 = all SMs; otherwise cap the persistent grid (BIAS_BF16 / LNFOLD modes) so a concurrent stream keeps SMs */
  /* LayerNorm fold, producer (RESID_STATS): */
  void* copy_out;          /* bf16 [M, ld_copy]: the output again, rounded to bf16 (the consumer GEMM's A operand) */
  int64_t ld_copy;
  float* stats_out;        /* f32 [M, n_slots, 2]: (sum, sum of squares) over each column slice of a row;
                              n_slots must equal gava_gemm_stats_slots(M, N) */
  /* LayerNorm fold, consumer (LNFOLD): */
  const float* stats_in;   /* f32 [rows, n_slots, 2] written by the producer, row r of A at stats_in + r * ld_stats */
  int64_t ld_stats;        /* floats between consecutive A rows (2 * n_slots, times the row step of a strided A view) */
  const float* colsum;     /* f32 [N]: sum_k W'[n, k] of the bf16 operand actually passed as W; NULL when the rows of W' are
                              centred (W' = W diag(g) - rowmean): the mean term then vanishes inside the contraction */
  int32_t n_slots;
  float eps;               /* LayerNorm epsilon (consumer) */
} gava_gemm_epilogue;

int gava_version(void);
const char* gava_last_error_string(void);
/* one-time per-process setup (kernel attributes); optional, every entry point self-initialises */
int gava_init(void);

/* Bytes of caller-owned workspace one clip shape needs (the library itself never allocates): B clips of T frames, N0 live
 * tokens per frame (cls + patches), width D, MLP width Dm, G global prompts, patch-row pitch Kpad, projected width Epad.
 * train_layers == 0: inference buffers only; > 0: plus the activation tape and backward scratch of a fine-tune step over
 * that many blocks.  Returns -1 on invalid arguments.  (Sizing aid: mirrors TowerEngine.workspace / TowerTrainer.train_ws.) */
int64_t gava_workspace_bytes(int32_t B, int32_t T, int32_t N0, int32_t D, int32_t Dm, int32_t G, int32_t Kpad, int32_t Epad,
                             int32_t train_layers);

/* C[M,N] = A[M,K] * W[N,K]^T with a fused epilogue.  A: bf16 [M,K] pitch lda; W: bf16 [N,K] pitch ldw
 * (nn.Linear weight layout).  tcgen05.mma / TMEM accumulators / TMA operand loads.
 * Needs K % 64 == 0, N % 128 == 0, 16-byte aligned pointers and pitches.
 * Replaces nn.Linear at VitaCLIP_vision_encoder_utils.py:66,79,110-115,164-165, Conv2d at :215-219 and
 * `cls_x @ self.proj` at VitaCLIP_vision_encoder.py:127. */
int gava_gemm_bf16_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                       const gava_gemm_epilogue* epi, void* stream);

/* Column slices per row a GAVA_EPI_BIAS_RESID_STATS_F32 launch of shape (M, N) writes to stats_out (0: unsupported). */
int32_t gava_gemm_stats_slots(int64_t M, int64_t N);

/* LayerNorm over the last dim (fp32 statistics, eps inside the sqrt), row r read from
 * x + r * in_row_stride (elements, f32), written to out + r * out_row_stride as bf16 (out_bf16 != 0) or f32.
 * weight == NULL: no normalisation, plain copy/cast (used to gather the cls rows).  In-place (out == x, f32) is safe.
 * D in {128,256,384,512,768,1024}.  Replaces LayerNorm.forward, VitaCLIP_vision_encoder_utils.py:22-28. */
int gava_layernorm_fwd(const float* x, int64_t in_row_stride, const float* weight, const float* bias, void* out,
                       int64_t out_row_stride, int32_t out_bf16, int64_t rows, int32_t D, float eps, void* stream);

/* Two consecutive LayerNorms in one pass: out (f32, may alias x) = LN(x; weight, bias), out2 (bf16) = LN(out; weight2,
 * bias2).  ln_pre followed by block 0's norm1, VitaCLIP_vision_encoder.py:113 + ..._utils.py:190. */
int gava_layernorm2_fwd(const float* x, int64_t in_row_stride, const float* weight, const float* bias, float* out,
                        int64_t out_row_stride, const float* weight2, const float* bias2, void* out2, int64_t out2_row_stride,
                        int64_t rows, int32_t D, float eps, void* stream);

/* gava_layernorm_fwd of (x + delta): delta is a bf16 [rows, D] tensor (row pitch delta_row_stride) holding a residual branch
 * that was not yet added to the f32 stream -- `x + attn(...)` of VitaCLIP_vision_encoder_utils.py:191 with the out_proj result
 * kept in bf16; the MLP's fc2 epilogue later adds the same delta (gava_gemm_bf16_fwd, RESID mode with `aux`). */
int gava_layernorm_resid_fwd(const float* x, int64_t in_row_stride, const void* delta, int64_t delta_row_stride,
                             const float* weight, const float* bias, void* out, int64_t out_row_stride, int32_t out_bf16,
                             int64_t rows, int32_t D, float eps, void* stream);

/* Patch gather (im2col) for Conv2d(3, D, k=s=P): video f32 [B,3,T,H,W] -> A bf16 [B*T*(H/P)*(W/P), Kpad] with
 * column index c*P*P + ph*P + pw (the flattened conv weight order); columns >= 3*P*P are zero.
 * Replaces the frame flatten + conv input read, VitaCLIP_vision_encoder.py:105, ..._utils.py:219. */
int gava_patch_gather_fwd(const float* video, void* A, int32_t B, int32_t T, int32_t H, int32_t W, int32_t P,
                          int32_t Kpad, void* stream);

/* Input stage fused into the patch gather: decoder output uint8 [B,T,H,W,3] (HWC, as PyAV's to_ndarray) ->
 * (u/255 - mean[c]) / std[c] in f32 -> bf16 patch rows `hi` (and, if lo != NULL, the bf16 remainder for the fp32 path).
 * mean/std: HOST pointers to 3 floats each.  Replaces `frames.float()/255; (frames-mean)/std; permute`
 * video_dataset/dataset.py:117-122 plus gava_patch_gather_fwd. */
int gava_patch_gather_u8_fwd(const uint8_t* frames, void* hi, void* lo, const float* mean, const float* std, int32_t B,
                             int32_t T, int32_t H, int32_t W, int32_t P, int32_t Kpad, void* stream);

/* cls rows of the embedded stream: x[f*tokens_per_frame + 0, :] = cls_token + pos_embed[0] + time_embed[f % T].
 * Replaces the cls concat + broadcasts, VitaCLIP_vision_encoder.py:108-111. */
int gava_cls_rows_fwd(float* x, int32_t tokens_per_frame, const float* cls_token, const float* pos_embed,
                      const float* time_embed, int32_t BT, int32_t T, int32_t D, void* stream);

/* Summary-token attention core: for every video b and head h, softmax(q k^T / sqrt(dh)) v over the T cls-derived
 * tokens.  qkv: bf16 [B*T, 3*D] = [q | k | v] of summary_attn_layer; out: bf16 [B*T, D].  dh must be 64.
 * Replaces the einsum/softmax/einsum of Attention.forward for L=T, VitaCLIP_vision_encoder_utils.py:71-77,170. */
int gava_temporal_attention_fwd(const void* qkv, void* out, int32_t B, int32_t T, int32_t H, void* stream);

/* K/V-only prompt rows of one layer, LayerNorm(norm1)-ed and cast to bf16:
 *   row g (g < G)            : global_prompts[g]
 *   row G + b*T + t          : local_prompts[t] + cls_proj_out[b*T + t]
 *   row G + B*T + f          : summary[f]                     (f = b*T + t)
 * out: bf16 [G + 2*B*T, D].  Replaces the global/local/summary token construction + norm1 on those rows,
 * VitaCLIP_vision_encoder.py:117-119 and ..._utils.py:171-190. */
int gava_prompt_rows_fwd(const float* global_prompts, const float* local_prompts, const float* cls_proj_out,
                         const float* summary, const float* ln_weight, const float* ln_bias, void* out, int32_t G,
                         int32_t B, int32_t T, int32_t D, float eps, void* stream);

/* Main multi-head self-attention core with a K/V-only prompt prefix (flash-style, scores never leave the SM).
 * For frame f (video b = f / T) and head h the queries are the first n_q live rows of the frame and the keys/values
 * are: the N0 live rows of the frame (from qkv) + G global rows + the T local rows of video b + the summary row of
 * frame f (from kv_aux, row layout of gava_prompt_rows_fwd).
 *   qkv    : bf16 [B*T*N0, 3*D]  = [q | k | v] (q NOT pre-scaled; 1/sqrt(dh) is applied inside)
 *   kv_aux : bf16 [G + 2*B*T, 2*D] = [k | v]
 *   out    : bf16 [B*T*N0, D]; only rows f*N0 + i, i < n_q are written
 * dh = D / H must be 64.  key_groups selects the prompt groups that take part: bit 0 = the T local rows
 * (use_local_prompts), bit 1 = the summary row (use_summary_token); the G global rows always do; 3 = the full Vita-CLIP
 * configuration.  Rows of a disabled group keep their place in kv_aux (contents ignored, must be finite).
 * Replaces Attention.forward core, VitaCLIP_vision_encoder_utils.py:71-77, for the token set built at :160-188 (dead
 * query rows :194-197 are never computed). */
int gava_attention_fwd(const void* qkv, const void* kv_aux, void* out, int32_t B, int32_t T, int32_t N0, int32_t G,
                       int32_t H, int32_t n_q, int32_t key_groups, void* stream);

/* Logit head.  frame_feat: f32 [B*T, E] (ln_post(cls) @ proj per frame); summary_tok: f32 [B*T, D] or NULL.
 *   video_features[b] = mean_t frame_feat[b*T + t]                               (f32 [B, E], un-normalised)
 *   summary_out[b]    = mean_t summary_tok[b*T + t]                              (f32 [B, D], optional)
 *   logits[b, c]      = mean_{k in class c} scale * <v_hat_b, t_hat_k> + bias    (f32 [B, n_cls])
 * text_hat: f32 [R, E], rows already L2-normalised, grouped by class; class_offsets: int32 [n_cls + 1].
 * Replaces VitaCLIP_vision_encoder.py:128-130 and VitaCLIP_model.py:255,287-289 / :303-306, :308-309. */
int gava_head_fwd(const float* frame_feat, const float* summary_tok, const float* text_hat,
                  const int32_t* class_offsets, float* video_features, float* summary_out, float* logits, int32_t B,
                  int32_t T, int32_t E, int32_t D, int32_t n_cls, float scale, float logit_bias, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * fp32-accuracy path (logits within 1e-3 of the fp32 reference).  Contractions still run on the tensor cores: an fp32
 * operand x is split as hi + lo (two bf16 tensors, lo = bf16(x - hi)) and a product is evaluated as
 * Ahi*Whi + Ahi*Wlo + Alo*Whi by three gava_gemm_bf16_fwd launches accumulating into an fp32 output
 * (GAVA_EPI_BIAS_F32 first, then GAVA_EPI_BIAS_RESID_F32 with resid == out).  The entry points below are the glue. */

/* hi[r, :] = bf16(x[r, :]), lo[r, :] = bf16(x[r, :] - hi[r, :]); x rows at pitch in_row_stride (f32 elements), hi/lo rows
 * at pitch out_row_stride (bf16 elements).  D % 4 == 0. */
int gava_split_bf16_fwd(const float* x, int64_t in_row_stride, void* hi, void* lo, int64_t out_row_stride, int64_t rows,
                        int32_t D, void* stream);

/* out[r, :] = [hi | lo | hi] (bf16, 3D wide, contiguous rows): with a weight packed as [w_hi | w_hi | w_lo] along K, a single
 * gava_gemm_bf16_fwd with K' = 3D evaluates the same 3-term product (one launch instead of three; used on the text side,
 * where rows are few and launches are the cost). */
int gava_split3_bf16_fwd(const float* x, int64_t in_row_stride, void* out, int64_t rows, int32_t D, void* stream);

/* gava_patch_gather_fwd with a hi/lo bf16 pair as output (same row/column layout). */
int gava_patch_gather_split_fwd(const float* video, void* hi, void* lo, int32_t B, int32_t T, int32_t H, int32_t W,
                                int32_t P, int32_t Kpad, void* stream);

/* x[f*(Np+1) + 1 + n, :] = tmp[f*Np + n, :] + pos_embed[1 + n, :] + time_embed[f % T, :]   (all f32)
 * (the patch-embed epilogue of the bf16 path as a separate step; VitaCLIP_vision_encoder.py:108-111). */
int gava_embed_assemble_fwd(const float* tmp, float* x, const float* pos_embed, const float* time_embed, int32_t BT,
                            int32_t T, int32_t Np, int32_t D, void* stream);

/* x <- x * sigmoid(1.702 x) in place, f32, n % 4 == 0   (VitaCLIP_vision_encoder_utils.py:18-20). */
int gava_quickgelu_f32_fwd(float* x, int64_t n, void* stream);

/* gava_attention_fwd in f32 on the CUDA cores: qkv f32 [B*T*N0, 3D], kv_aux f32 [G+2BT, 2D], out f32 [B*T*N0, D].
 * kv_aux == NULL: plain self-attention over the N0 rows of each of the B*T groups (used for the summary-token
 * attention with B*T = videos, N0 = frames).  key_groups as in gava_attention_fwd. */
int gava_attention_f32_fwd(const float* qkv, const float* kv_aux, float* out, int32_t B, int32_t T, int32_t N0,
                           int32_t G, int32_t H, int32_t n_q, int32_t key_groups, void* stream);

/* Causal multi-head self-attention core of the CLIP text transformer in f32: qkv f32 [n_seq*L, 3D] (packed in_proj
 * output), out f32 [n_seq*L, D]; query t attends tokens 0..t of its sequence; dh = 64.
 * Replaces nn.MultiheadAttention's core with the causal mask, VitaCLIP_text_encoder.py:83-88,146-152. */
int gava_causal_attention_f32_fwd(const float* qkv, float* out, int32_t n_seq, int32_t L, int32_t H, void* stream);

/* gava_prompt_rows_fwd with f32 output. */
int gava_prompt_rows_f32_fwd(const float* global_prompts, const float* local_prompts, const float* cls_proj_out,
                             const float* summary, const float* ln_weight, const float* ln_bias, float* out, int32_t G,
                             int32_t B, int32_t T, int32_t D, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fine-tune step: backward through the frozen backbone to the trainable prompt / summary parameters
 * (VitaCLIP_model.py:230-235), replacing autograd of the ops above, plus the optimizer update (train.py:351,485).
 * Dense input gradients dX = dY * W reuse gava_gemm_bf16_fwd with W^T packed as the [N,K] operand; the small weight
 * gradients dW = dY^T X reuse it on operands transposed by gava_transpose_bf16. */

/* LayerNorm input gradient.  Row r: dy at dy + r*dy_stride (bf16 if dy_bf16 else f32), x at x + r*x_stride (the forward
 * INPUT, f32).  dx_out[r] = (add1 ? add1[r] : 0) + (add2 ? add2[r] : 0) + dLN(dy[r]); dx_out may alias add1.
 * dx_bf16 (optional): bf16 copy of dx_out (operand of the next dgrad GEMM).  dw/db (optional, both or neither): f32 [D],
 * atomically accumulated weight / bias gradients (caller zeroes them). */
int gava_layernorm_bwd(const void* dy, int64_t dy_stride, int32_t dy_bf16, const float* x, int64_t x_stride,
                       const float* weight, const float* add1, int64_t add1_stride, const float* add2,
                       int64_t add2_stride, float* dx_out, int64_t dx_stride, void* dx_bf16, int64_t dxb_stride,
                       float* dw, float* db, int64_t rows, int32_t D, float eps, void* stream);

/* gava_attention_fwd for the training forward: additionally writes lse f32 [B*T, H, N0] (rows < n_q) = the log2-domain
 * log-sum-exp of every query row's scaled scores (log2(sum_k 2^(s_k log2 e / 8))), saved for gava_attention_bwd. */
int gava_attention_train_fwd(const void* qkv, const void* kv_aux, void* out, float* lse, int32_t B, int32_t T, int32_t N0,
                             int32_t G, int32_t H, int32_t n_q, int32_t key_groups, void* stream);

/* Backward of gava_attention_fwd for the first n_q query rows of every frame (n_q = N0, or 1 in the last layer where only
 * the cls row reaches the output: dq is then written for that row only, dk/dv still cover every key).
 * out/dout: bf16 [B*T*N0, D] forward output and its gradient (rows >= n_q of a frame are not read).
 * lse: the rows saved by gava_attention_train_fwd, or NULL (recomputed inside, one more pass over the keys).
 * dqkv: bf16 [B*T*N0, 3D] = [dq | dk | dv] of the live rows; dkv_aux: f32 [G + 2*B*T, 2D] = [dk | dv] of the K/V-only rows
 * (global rows collect all frames, local rows the T frames of a video).  scratch: f32 [B*T, H, G+T+1, 128] workspace: every
 * (frame, head) writes its prompt-row contributions to its own slot and a second kernel sums the slots in a fixed order
 * (reproducible); scratch == NULL: f32 atomics straight into dkv_aux (order-dependent rounding).
 * key_groups as in gava_attention_fwd (the value the forward ran with): keys of a disabled prompt group have zero
 * probability, so their dk / dv rows come out as exact zeros. */
int gava_attention_bwd(const void* qkv, const void* kv_aux, const void* out, const void* dout, const float* lse,
                       void* dqkv, float* dkv_aux, float* scratch, int32_t B, int32_t T, int32_t N0, int32_t G, int32_t H,
                       int32_t n_q, int32_t key_groups, void* stream);

/* Backward of gava_temporal_attention_fwd: qkv bf16 [B*T, 3D], dout bf16 [B*T, D] -> dqkv bf16 [B*T, 3D].  T <= 128. */
int gava_temporal_attention_bwd(const void* qkv, const void* dout, void* dqkv, int32_t B, int32_t T, int32_t H,
                                void* stream);

/* Backward of gava_causal_attention_f32_fwd (text transformer, fine-tune step): qkv f32 [n_seq*L, 3D] (the forward's
 * input), dout bf16 [n_seq*L, D] -> dqkv bf16 [n_seq*L, 3D].  L <= 128, dh = 64. */
int gava_causal_attention_bwd(const float* qkv, const void* dout, void* dqkv, int32_t n_seq, int32_t L, int32_t H,
                              void* stream);

/* Backward of gava_prompt_rows_fwd.  d_rows_n: bf16 [G + 2*B*T, D] gradient of the normalised rows.  Outputs (f32):
 * d_global [G, D]; d_local_cp [B*T, D] (gradient of local_prompts[t] + cls_proj_out[f], per frame);
 * d_summary [B*T, D] (+ d_summary_ext[f] if given: the gradient arriving at the encoder's summary output). */
int gava_prompt_rows_bwd(const void* d_rows_n, const float* global_prompts, const float* local_prompts,
                         const float* cls_proj_out, const float* summary, const float* ln_weight, float* d_global,
                         float* d_local_cp, float* d_summary, const float* d_summary_ext, int32_t G, int32_t B, int32_t T,
                         int32_t D, float eps, void* stream);

/* out[j, :] = sum_{o<outer} sum_{n<inner} in[((o*J + j)*inner + n)*ld + :]; in f32 or bf16 (in_bf16), out f32 [J, D]
 * (time_embed gradient: outer=B, J=T, inner=N0; local_prompts: outer=B, J=T, inner=1; bias: outer=rows, J=1, inner=1). */
int gava_group_sum(const void* in, int64_t ld, int32_t in_bf16, float* out, int32_t outer, int32_t J, int32_t inner,
                   int32_t D, void* stream);

/* out bf16 [C, Rpad] = in[R, C]^T (in: f32 or bf16, row pitch ld), columns R..Rpad-1 zero (K padding of a wgrad GEMM). */
int gava_transpose_bf16(const void* in, int64_t ld, int32_t in_bf16, void* out, int32_t R, int32_t C, int32_t Rpad,
                        void* stream);

/* h = QuickGELU(a), bf16 -> bf16, n % 8 == 0 (training forward keeps the fc1 pre-activations for the DGELU epilogue). */
int gava_quickgelu_bf16_fwd(const void* a, void* h, int64_t n, void* stream);

/* Fused AdamW on a flat f32 buffer (torch.optim.AdamW semantics: decoupled decay, bias correction with `step` >= 1);
 * gradients are multiplied by grad_scale first (1/world_size after a SUM all-reduce).  n % 4 == 0. */
int gava_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int32_t step, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Evaluation input stage in one kernel: decoder frames uint8 [T0, H0, W0, 3] -> u/255 ->
 * (x - mean[c]) / std[c] -> bilinear resize of every frame to (new_h, new_w) (align_corners = False, ATen weights) ->
 * spatial crop [h_st, h_st+Sh) x [w_st, w_st+Sw) -> frames t_st + i*t_step (i < T; indices past the end repeat the last
 * frame) -> out f32 [3, T, Sh, Sw].  mean/std: HOST pointers to 3 floats.
 * Replaces video_dataset/dataset.py:117-139 (eval branch) and the crops of :160-196. */
int gava_preprocess_u8_fwd(const uint8_t* frames, float* out, int32_t T0, int32_t H0, int32_t W0, int32_t new_h,
                           int32_t new_w, int32_t h_st, int32_t w_st, int32_t t_st, int32_t t_step, int32_t T, int32_t Sh,
                           int32_t Sw, const float* mean, const float* std, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAVA_SM100_H_ */
