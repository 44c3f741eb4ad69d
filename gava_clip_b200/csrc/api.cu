// extern "C" surface of libgava_sm100.so (declared in include/gava_sm100.h).  Thin shim: argument
// hand-off to the kernels' host launchers, integer status + thread-local error string, no exceptions.
#include "common.cuh"
#include "kernels.h"

#include <stdarg.h>
#include <stddef.h>

namespace gava {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return GAVA_OK;
}

}  // namespace gava

using namespace gava;

extern "C" {

int gava_version(void) { return GAVA_ABI_VERSION; }

const char* gava_last_error_string(void) { return g_err; }

int gava_init(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) { set_error("gava_init: %s", cudaGetErrorString(e)); return (int)e; }
  if (major != 10) { set_error("gava_init: device compute capability %d.x is not sm_100", major); return GAVA_ERR_DRIVER; }
  return GAVA_OK;
}

int64_t gava_workspace_bytes(int32_t B, int32_t T, int32_t N0, int32_t D, int32_t Dm, int32_t G, int32_t Kpad, int32_t Epad,
                             int32_t train_layers) {
  // mirror of TowerEngine.workspace / TowerTrainer.train_ws (vision_encoder.py, train_engine.py): what a caller has to
  // allocate for one clip shape.  Sizes in bytes; bf16 = 2, f32 = 4.
  if (B <= 0 || T <= 0 || N0 <= 1 || D <= 0 || Dm <= 0 || G < 0 || Kpad <= 0 || Epad <= 0 || train_layers < 0) return -1;
  const int64_t BT = (int64_t)B * T, M = BT * N0, R = G + 2 * BT;
  int64_t infer = 0;
  infer += BT * (N0 - 1) * Kpad * 2;            // patches
  infer += M * D * 4 + M * D * 2 * 2;           // x, xn, delta
  infer += M * D * 2 + M * (D / 64) * 2 * 4;    // xb, xstats (LayerNorm fold; at most D/64 column slices per row)
  infer += M * 3 * D * 2 + M * Dm * 2;          // qkv, h
  infer += BT * D * (2 + 2 + 4 + 4 + 2 + 2 + 4) + BT * 3 * D * 2;   // cls_bf, cls_bf2, cls_proj, x_cls, sum_n, smix, summary, sqkv
  infer += R * D * 2 + R * 2 * D * 2;           // aux_n, kvaux
  infer += BT * Epad * 4;                       // frame_feat
  if (train_layers == 0) return infer;
  const int64_t L = train_layers;
  int64_t tape = M * D * 4;                                                     // x0_pre
  tape += L * (M * D * 4) + (L - 1) * (M * D * 4);                              // x_in, x_mid
  tape += L * (M * 3 * D * 2 + R * 2 * D * 2 + M * D * 2 + BT * (D / 64) * N0 * 4);   // qkv, kvaux, o, lse
  tape += (L - 1) * (M * Dm * 2);                                               // a1
  tape += L * (BT * D * (4 + 4 + 2 + 2) + BT * 3 * D * 2);                      // cls_proj, summary, sum_n, smix, sqkv
  int64_t scratch = M * D * 4 + M * D * 2 * 2 + M * Dm * 2 + M * 3 * D * 2;     // dx, dxb, gD, g4, dqkv
  scratch += R * 2 * D * (4 + 2) + R * D * 2 + BT * (D / 64) * (G + T + 1) * 128 * 4;  // dkvaux (+bf16), daux_n, slots
  return infer + tape + scratch;
}

static_assert(sizeof(gava_gemm_epilogue) == 168 && offsetof(gava_gemm_epilogue, max_ctas) == 104 &&
                  offsetof(gava_gemm_epilogue, copy_out) == 112 && offsetof(gava_gemm_epilogue, stats_in) == 136 &&
                  offsetof(gava_gemm_epilogue, colsum) == 152 && offsetof(gava_gemm_epilogue, eps) == 164,
              "gava_gemm_epilogue layout is part of the ABI (ctypes mirror: _lib.GemmEpilogue)");

int gava_gemm_bf16_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                       const gava_gemm_epilogue* epi, void* stream) {
  return gemm_bf16(A, lda, W, ldw, M, N, K, epi, (cudaStream_t)stream);
}

int32_t gava_gemm_stats_slots(int64_t M, int64_t N) { return gemm_stats_slots(M, N); }

int gava_layernorm_fwd(const float* x, int64_t in_row_stride, const float* weight, const float* bias, void* out,
                       int64_t out_row_stride, int32_t out_bf16, int64_t rows, int32_t D, float eps, void* stream) {
  return layernorm(x, in_row_stride, weight, bias, out, out_row_stride, out_bf16, rows, D, eps, (cudaStream_t)stream);
}

int gava_layernorm2_fwd(const float* x, int64_t in_row_stride, const float* weight, const float* bias, float* out,
                        int64_t out_row_stride, const float* weight2, const float* bias2, void* out2, int64_t out2_row_stride,
                        int64_t rows, int32_t D, float eps, void* stream) {
  return layernorm2(x, in_row_stride, weight, bias, out, out_row_stride, weight2, bias2, out2, out2_row_stride, rows, D, eps,
                    (cudaStream_t)stream);
}

int gava_layernorm_resid_fwd(const float* x, int64_t in_row_stride, const void* delta, int64_t delta_row_stride,
                             const float* weight, const float* bias, void* out, int64_t out_row_stride, int32_t out_bf16,
                             int64_t rows, int32_t D, float eps, void* stream) {
  return layernorm_resid(x, in_row_stride, delta, delta_row_stride, weight, bias, out, out_row_stride, out_bf16, rows, D, eps,
                         (cudaStream_t)stream);
}

int gava_patch_gather_fwd(const float* video, void* A, int32_t B, int32_t T, int32_t H, int32_t W, int32_t P,
                          int32_t Kpad, void* stream) {
  return patch_gather(video, A, B, T, H, W, P, Kpad, (cudaStream_t)stream);
}

int gava_patch_gather_u8_fwd(const uint8_t* frames, void* hi, void* lo, const float* mean, const float* std, int32_t B,
                             int32_t T, int32_t H, int32_t W, int32_t P, int32_t Kpad, void* stream) {
  return patch_gather_u8(frames, hi, lo, mean, std, B, T, H, W, P, Kpad, (cudaStream_t)stream);
}

int gava_cls_rows_fwd(float* x, int32_t tokens_per_frame, const float* cls_token, const float* pos_embed,
                      const float* time_embed, int32_t BT, int32_t T, int32_t D, void* stream) {
  return cls_rows(x, tokens_per_frame, cls_token, pos_embed, time_embed, BT, T, D, (cudaStream_t)stream);
}

int gava_temporal_attention_fwd(const void* qkv, void* out, int32_t B, int32_t T, int32_t H, void* stream) {
  return temporal_attention(qkv, out, B, T, H, (cudaStream_t)stream);
}

int gava_prompt_rows_fwd(const float* global_prompts, const float* local_prompts, const float* cls_proj_out,
                         const float* summary, const float* ln_weight, const float* ln_bias, void* out, int32_t G,
                         int32_t B, int32_t T, int32_t D, float eps, void* stream) {
  return prompt_rows(global_prompts, local_prompts, cls_proj_out, summary, ln_weight, ln_bias, out, G, B, T, D, eps,
                     (cudaStream_t)stream);
}

int gava_attention_fwd(const void* qkv, const void* kv_aux, void* out, int32_t B, int32_t T, int32_t N0, int32_t G,
                       int32_t H, int32_t n_q, int32_t key_groups, void* stream) {
  return attention(qkv, kv_aux, out, nullptr, B, T, N0, G, H, n_q, key_groups, (cudaStream_t)stream);
}

int gava_attention_train_fwd(const void* qkv, const void* kv_aux, void* out, float* lse, int32_t B, int32_t T, int32_t N0,
                             int32_t G, int32_t H, int32_t n_q, int32_t key_groups, void* stream) {
  return attention(qkv, kv_aux, out, lse, B, T, N0, G, H, n_q, key_groups, (cudaStream_t)stream);
}

int gava_head_fwd(const float* frame_feat, const float* summary_tok, const float* text_hat,
                  const int32_t* class_offsets, float* video_features, float* summary_out, float* logits, int32_t B,
                  int32_t T, int32_t E, int32_t D, int32_t n_cls, float scale, float logit_bias, void* stream) {
  return head(frame_feat, summary_tok, text_hat, class_offsets, video_features, summary_out, logits, B, T, E, D, n_cls,
              scale, logit_bias, (cudaStream_t)stream);
}

int gava_split_bf16_fwd(const float* x, int64_t in_row_stride, void* hi, void* lo, int64_t out_row_stride, int64_t rows,
                        int32_t D, void* stream) {
  return split_bf16(x, in_row_stride, hi, lo, out_row_stride, rows, D, (cudaStream_t)stream);
}

int gava_split3_bf16_fwd(const float* x, int64_t in_row_stride, void* out, int64_t rows, int32_t D, void* stream) {
  return split3_bf16(x, in_row_stride, out, rows, D, (cudaStream_t)stream);
}

int gava_patch_gather_split_fwd(const float* video, void* hi, void* lo, int32_t B, int32_t T, int32_t H, int32_t W,
                                int32_t P, int32_t Kpad, void* stream) {
  return patch_gather_split(video, hi, lo, B, T, H, W, P, Kpad, (cudaStream_t)stream);
}

int gava_embed_assemble_fwd(const float* tmp, float* x, const float* pos_embed, const float* time_embed, int32_t BT,
                            int32_t T, int32_t Np, int32_t D, void* stream) {
  return embed_assemble(tmp, x, pos_embed, time_embed, BT, T, Np, D, (cudaStream_t)stream);
}

int gava_quickgelu_f32_fwd(float* x, int64_t n, void* stream) { return quickgelu_f32(x, n, (cudaStream_t)stream); }

int gava_attention_f32_fwd(const float* qkv, const float* kv_aux, float* out, int32_t B, int32_t T, int32_t N0,
                           int32_t G, int32_t H, int32_t n_q, int32_t key_groups, void* stream) {
  return attention_f32(qkv, kv_aux, out, B, T, N0, G, H, n_q, key_groups, (cudaStream_t)stream);
}

int gava_causal_attention_f32_fwd(const float* qkv, float* out, int32_t n_seq, int32_t L, int32_t H, void* stream) {
  return causal_attention_f32(qkv, out, n_seq, L, H, (cudaStream_t)stream);
}

int gava_prompt_rows_f32_fwd(const float* global_prompts, const float* local_prompts, const float* cls_proj_out,
                             const float* summary, const float* ln_weight, const float* ln_bias, float* out, int32_t G,
                             int32_t B, int32_t T, int32_t D, float eps, void* stream) {
  return prompt_rows_f32(global_prompts, local_prompts, cls_proj_out, summary, ln_weight, ln_bias, out, G, B, T, D, eps,
                         (cudaStream_t)stream);
}

int gava_layernorm_bwd(const void* dy, int64_t dy_stride, int32_t dy_bf16, const float* x, int64_t x_stride,
                       const float* weight, const float* add1, int64_t add1_stride, const float* add2,
                       int64_t add2_stride, float* dx_out, int64_t dx_stride, void* dx_bf16, int64_t dxb_stride,
                       float* dw, float* db, int64_t rows, int32_t D, float eps, void* stream) {
  return layernorm_bwd(dy, dy_stride, dy_bf16, x, x_stride, weight, add1, add1_stride, add2, add2_stride, dx_out,
                       dx_stride, dx_bf16, dxb_stride, dw, db, rows, D, eps, (cudaStream_t)stream);
}

int gava_attention_bwd(const void* qkv, const void* kv_aux, const void* out, const void* dout, const float* lse,
                       void* dqkv, float* dkv_aux, float* scratch, int32_t B, int32_t T, int32_t N0, int32_t G, int32_t H,
                       int32_t n_q, int32_t key_groups, void* stream) {
  return attention_bwd(qkv, kv_aux, out, dout, lse, dqkv, dkv_aux, scratch, B, T, N0, G, H, n_q, key_groups, (cudaStream_t)stream);
}

int gava_temporal_attention_bwd(const void* qkv, const void* dout, void* dqkv, int32_t B, int32_t T, int32_t H,
                                void* stream) {
  return temporal_attention_bwd(qkv, dout, dqkv, B, T, H, (cudaStream_t)stream);
}

int gava_causal_attention_bwd(const float* qkv, const void* dout, void* dqkv, int32_t n_seq, int32_t L, int32_t H,
                              void* stream) {
  return causal_attention_bwd(qkv, dout, dqkv, n_seq, L, H, (cudaStream_t)stream);
}

int gava_prompt_rows_bwd(const void* d_rows_n, const float* global_prompts, const float* local_prompts,
                         const float* cls_proj_out, const float* summary, const float* ln_weight, float* d_global,
                         float* d_local_cp, float* d_summary, const float* d_summary_ext, int32_t G, int32_t B, int32_t T,
                         int32_t D, float eps, void* stream) {
  return prompt_rows_bwd(d_rows_n, global_prompts, local_prompts, cls_proj_out, summary, ln_weight, d_global, d_local_cp,
                         d_summary, d_summary_ext, G, B, T, D, eps, (cudaStream_t)stream);
}

int gava_group_sum(const void* in, int64_t ld, int32_t in_bf16, float* out, int32_t outer, int32_t J, int32_t inner,
                   int32_t D, void* stream) {
  return group_sum(in, ld, in_bf16, out, outer, J, inner, D, (cudaStream_t)stream);
}

int gava_transpose_bf16(const void* in, int64_t ld, int32_t in_bf16, void* out, int32_t R, int32_t C, int32_t Rpad,
                        void* stream) {
  return transpose_bf16(in, ld, in_bf16, out, R, C, Rpad, (cudaStream_t)stream);
}

int gava_quickgelu_bf16_fwd(const void* a, void* h, int64_t n, void* stream) {
  return quickgelu_bf16(a, h, n, (cudaStream_t)stream);
}

int gava_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int32_t step, float grad_scale, void* stream) {
  return adamw_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                    (cudaStream_t)stream);
}

int gava_preprocess_u8_fwd(const uint8_t* frames, float* out, int32_t T0, int32_t H0, int32_t W0, int32_t new_h, int32_t new_w,
                           int32_t h_st, int32_t w_st, int32_t t_st, int32_t t_step, int32_t T, int32_t Sh, int32_t Sw,
                           const float* mean, const float* std, void* stream) {
  return preprocess_u8(frames, out, T0, H0, W0, new_h, new_w, h_st, w_st, t_st, t_step, T, Sh, Sw, mean, std,
                       (cudaStream_t)stream);
}

}  // extern "C"
