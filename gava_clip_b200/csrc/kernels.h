// Internal C++ declarations shared by the kernel translation units and the C-ABI shim (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include "../../include/gava_sm100.h"

namespace gava {

enum : int {
  EPI_BIAS_BF16 = GAVA_EPI_BIAS_BF16,
  EPI_BIAS_GELU_BF16 = GAVA_EPI_BIAS_GELU_BF16,
  EPI_BIAS_RESID_F32 = GAVA_EPI_BIAS_RESID_F32,
  EPI_BIAS_F32 = GAVA_EPI_BIAS_F32,
  EPI_PATCH_F32 = GAVA_EPI_PATCH_F32,
  EPI_DGELU_BF16 = GAVA_EPI_DGELU_BF16,
  EPI_LNFOLD_BF16 = GAVA_EPI_LNFOLD_BF16,
  EPI_BIAS_RESID_STATS_F32 = GAVA_EPI_BIAS_RESID_STATS_F32,
};

int num_sms();         // of the CURRENT device
int device_slot();     // cudaGetDevice(), clamped to [0, 64): index for per-device caches (kernel attributes are per device)

// 2-D bf16 row-major [rows, cols] (row pitch ld elements) TMA map, box = [box_rows, 64 cols], 128-byte swizzle
int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows);

int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, long long planes, long long rows, long long cols, long long ld,
                      long long ld_plane, int box_rows);

int gemm_bf16(const void* A, long long lda, const void* W, long long ldw, long long M, long long N, long long K,
              const gava_gemm_epilogue* epi, cudaStream_t stream);
int gemm_stats_slots(long long M, long long N);

int layernorm(const float* x, long long in_row_stride, const float* w, const float* b, void* out,
              long long out_row_stride, int out_bf16, long long rows, int D, float eps, cudaStream_t stream);

int layernorm2(const float* x, long long in_row_stride, const float* w, const float* b, float* out, long long out_row_stride,
               const float* w2, const float* b2, void* out2, long long out2_row_stride, long long rows, int D, float eps,
               cudaStream_t stream);

// LayerNorm(x + delta), delta bf16 (may be null)
int layernorm_resid(const float* x, long long in_row_stride, const void* delta, long long delta_row_stride, const float* w,
                    const float* b, void* out, long long out_row_stride, int out_bf16, long long rows, int D, float eps,
                    cudaStream_t stream);

int patch_gather(const float* video, void* A, int B, int T, int H, int W, int P, int Kpad, cudaStream_t stream);

// mean/std are HOST arrays of 3 floats (passed by value to the kernel); lo may be null
int patch_gather_u8(const uint8_t* frames, void* hi, void* lo, const float* mean, const float* std, int B, int T, int H,
                    int W, int P, int Kpad, cudaStream_t stream);

int cls_rows(float* x, int tokens_per_frame, const float* cls_token, const float* pos_embed, const float* time_embed,
             int BT, int T, int D, cudaStream_t stream);

int temporal_attention(const void* qkv, void* out, int B, int T, int H, cudaStream_t stream);

int prompt_rows(const float* global_prompts, const float* local_prompts, const float* cls_proj_out,
                const float* summary, const float* ln_w, const float* ln_b, void* out, int G, int B, int T, int D,
                float eps, cudaStream_t stream);

// tcgen05/TMEM/TMA attention (attention_tc.cu); returns GAVA_ERR_SHAPE when the shape is outside its specialisation
int attention_tc_keys(int T, int N0, int G);   // smem key rows needed; 0 = outside the specialisation
// key_groups: bit 0 = the T local-prompt rows take part, bit 1 = the summary-token row takes part (the global prompts always
// do); 3 = the full Vita-CLIP configuration.  Rows of a disabled group stay in kv_aux's layout but get zero probability.
int attention_tc(const void* qkv, const void* kv_aux, void* out, float* lse, int B, int T, int N0, int G, int H, int n_q,
                 int key_groups, cudaStream_t stream);

int attention(const void* qkv, const void* kv_aux, void* out, float* lse, int B, int T, int N0, int G, int H, int n_q,
              int key_groups, cudaStream_t stream);

int head(const float* frame_feat, const float* summary_tok, const float* text_hat, const int* class_offsets,
         float* video_features, float* summary_out, float* logits, int B, int T, int E, int D, int n_cls, float scale,
         float logit_bias, cudaStream_t stream);

// ---- fp32-accuracy path (precise.cu)
int split_bf16(const float* x, long long in_row_stride, void* hi, void* lo, long long out_row_stride, long long rows,
               int D, cudaStream_t stream);
int split3_bf16(const float* x, long long in_row_stride, void* out, long long rows, int D, cudaStream_t stream);
int patch_gather_split(const float* video, void* hi, void* lo, int B, int T, int H, int W, int P, int Kpad,
                       cudaStream_t stream);
int embed_assemble(const float* tmp, float* x, const float* pos, const float* time_embed, int BT, int T, int Np, int D,
                   cudaStream_t stream);
int quickgelu_f32(float* x, long long n, cudaStream_t stream);
int attention_f32(const float* qkv, const float* kv_aux, float* out, int B, int T, int N0, int G, int H, int n_q,
                  int key_groups, cudaStream_t stream);
int causal_attention_f32(const float* qkv, float* out, int n_seq, int L, int H, cudaStream_t stream);
// tensor-core forms for L <= 128 (seq_attention_mma.cu): fp32-accurate 3-term split-bf16 forward, bf16 backward
bool causal_attention_mma_supported(int L, int H);
int causal_attention_mma(const float* qkv, float* out, int n_seq, int L, int H, cudaStream_t stream);
int causal_attention_bwd_mma(const float* qkv, const void* dout, void* dqkv, int n_seq, int L, int H, cudaStream_t stream);
// the same kernels without a mask on bf16 q|k|v: the summary-token attention for clips of more than 16 frames
int temporal_attention_mma(const void* qkv, void* out, int B, int T, int H, cudaStream_t stream);
int temporal_attention_bwd_mma(const void* qkv, const void* dout, void* dqkv, int B, int T, int H, cudaStream_t stream);
int prompt_rows_f32(const float* global_prompts, const float* local_prompts, const float* cls_proj_out,
                    const float* summary, const float* ln_w, const float* ln_b, float* out, int G, int B, int T, int D,
                    float eps, cudaStream_t stream);

// ---- fine-tune step: backward + optimizer (backward.cu, attention_bwd.cu)
int layernorm_bwd(const void* dy, long long dy_stride, int dy_bf16, const float* x, long long x_stride, const float* w,
                  const float* add1, long long add1_stride, const float* add2, long long add2_stride, float* dx_out,
                  long long dx_stride, void* dx_bf16, long long dxb_stride, float* dw, float* db, long long rows, int D,
                  float eps, cudaStream_t stream);
int attention_bwd(const void* qkv, const void* kv_aux, const void* out, const void* dout, const float* lse, void* dqkv,
                  float* dkv_aux, float* scratch, int B, int T, int N0, int G, int H, int n_q, int key_groups, cudaStream_t stream);
int temporal_attention_bwd(const void* qkv, const void* dout, void* dqkv, int B, int T, int H, cudaStream_t stream);
int causal_attention_bwd(const float* qkv, const void* dout, void* dqkv, int n_seq, int L, int H, cudaStream_t stream);
int prompt_rows_bwd(const void* d_rows_n, const float* global_prompts, const float* local_prompts,
                    const float* cls_proj_out, const float* summary, const float* ln_w, float* d_global,
                    float* d_local_cp, float* d_summary, const float* d_summary_ext, int G, int B, int T, int D, float eps,
                    cudaStream_t stream);
int group_sum(const void* in, long long ld, int in_bf16, float* out, int outer, int J, int inner, int D,
              cudaStream_t stream);
int transpose_bf16(const void* in, long long ld, int in_bf16, void* out, int R, int C, int Rpad, cudaStream_t stream);
int quickgelu_bf16(const void* a, void* h, long long n, cudaStream_t stream);
int adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int step, float grad_scale, cudaStream_t stream);

// ---- input stage (input_stage.cu); mean/std: HOST arrays of 3 floats
int preprocess_u8(const uint8_t* frames, float* out, int T0, int H0, int W0, int new_h, int new_w, int h_st, int w_st,
                  int t_st, int t_step, int T, int Sh, int Sw, const float* mean, const float* std, cudaStream_t stream);

}  // namespace gava
