// fp32-accuracy path ("within 1e-3 of the fp32 reference", BASELINE.json north_star).
//
// The tensor cores still do the contractions: every fp32 operand is split into two bf16 terms (x = hi + lo,
// lo = bf16(x - hi)) and a product is evaluated as  Ahi*Whi + Ahi*Wlo + Alo*Whi  by three launches of the tcgen05 GEMM
// accumulating into an fp32 output (the dropped Alo*Wlo term is ~2^-18 relative).  This file holds the glue kernels of
// that path: operand split, fp32 QuickGELU, fp32 prompt rows, fp32 attention cores (CUDA cores; this path is a parity
// instrument, not the throughput path) and the patch-embed assemble step.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace gava {

// ---------------------------------------------------------------- split: fp32 rows -> hi / lo bf16 rows
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, long long in_stride,
                                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                         __nv_bfloat16* __restrict__ hi2, long long out_stride,
                                                         long long rows, int d4n) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * d4n) return;
  const long long r = idx / d4n;
  const int d4 = (int)(idx - r * d4n);
  const float4 v = *reinterpret_cast<const float4*>(x + r * in_stride + 4 * d4);
  const __nv_bfloat16 h0 = __float2bfloat16(v.x), h1 = __float2bfloat16(v.y), h2 = __float2bfloat16(v.z), h3 = __float2bfloat16(v.w);
  uint2 oh, ol;
  oh.x = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
  oh.y = pack_bf16x2(__bfloat162float(h2), __bfloat162float(h3));
  ol.x = pack_bf16x2(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
  ol.y = pack_bf16x2(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
  *reinterpret_cast<uint2*>(hi + r * out_stride + 4 * d4) = oh;
  *reinterpret_cast<uint2*>(lo + r * out_stride + 4 * d4) = ol;
  if (hi2 != nullptr) *reinterpret_cast<uint2*>(hi2 + r * out_stride + 4 * d4) = oh;
}

// out[r, :] = [hi | lo | hi] (3D wide): with weights packed as [w_hi | w_hi | w_lo] along K, ONE GEMM with K' = 3D evaluates
// the 3-term product a_hi w_hi + a_lo w_hi + a_hi w_lo (used where launches, not bytes, are the cost: the text side)
int split3_bf16(const float* x, long long in_row_stride, void* out, long long rows, int D, cudaStream_t stream) {
  if (!x || !out || rows < 0 || D <= 0 || (D & 3) || (in_row_stride & 3)) { set_error("split3_bf16: bad arguments"); return GAVA_ERR_ARG; }
  if (rows == 0) return GAVA_OK;
  const long long total = rows * (D / 4);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  split_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, in_row_stride, o, o + D, o + 2 * D, 3LL * D, rows, D / 4);
  return check_launch("split3_bf16");
}

int split_bf16(const float* x, long long in_row_stride, void* hi, void* lo, long long out_row_stride, long long rows,
               int D, cudaStream_t stream) {
  if (!x || !hi || !lo || rows < 0 || D <= 0 || (D & 3) || (in_row_stride & 3) || (out_row_stride & 3)) {
    set_error("split_bf16: bad arguments");
    return GAVA_ERR_ARG;
  }
  if (rows == 0) return GAVA_OK;
  const long long total = rows * (D / 4);
  split_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      x, in_row_stride, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), nullptr, out_row_stride, rows, D / 4);
  return check_launch("split_bf16");
}

// ---------------------------------------------------------------- patch gather to hi / lo (any patch size)
__global__ void __launch_bounds__(256) patch_gather_split_kernel(const float* __restrict__ video,
                                                                 __nv_bfloat16* __restrict__ hi,
                                                                 __nv_bfloat16* __restrict__ lo, int B, int T, int H, int W,
                                                                 int P, int Kpad) {
  const long long total = (long long)B * 3 * T * H * W;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  long long r = idx / W;
  const int y = (int)(r % H); r /= H;
  const int t = (int)(r % T); r /= T;
  const int c = (int)(r % 3);
  const int b = (int)(r / 3);
  const int gw = W / P, gh = H / P;
  if (y >= gh * P || x >= gw * P) return;
  const long long arow = ((long long)(b * T + t) * gh + y / P) * gw + x / P;
  const int col = c * P * P + (y % P) * P + (x % P);
  const float v = __ldg(video + idx);
  const __nv_bfloat16 h = __float2bfloat16(v);
  hi[arow * Kpad + col] = h;
  lo[arow * Kpad + col] = __float2bfloat16(v - __bfloat162float(h));
}

int patch_gather_split(const float* video, void* hi, void* lo, int B, int T, int H, int W, int P, int Kpad,
                       cudaStream_t stream) {
  if (!video || !hi || !lo || B <= 0 || T <= 0 || P <= 0 || H < P || W < P || Kpad < 3 * P * P) {
    set_error("patch_gather_split: bad arguments");
    return GAVA_ERR_ARG;
  }
  const long long rows = (long long)B * T * (H / P) * (W / P);
  if (Kpad > 3 * P * P) {
    cudaError_t e = cudaMemsetAsync(hi, 0, (size_t)rows * Kpad * 2, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(lo, 0, (size_t)rows * Kpad * 2, stream);
    if (e != cudaSuccess) { set_error("patch_gather_split memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  const long long total = (long long)B * 3 * T * H * W;
  patch_gather_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      video, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), B, T, H, W, P, Kpad);
  return check_launch("patch_gather_split");
}

// ---------------------------------------------------------------- x[f, 1+n] = tmp[f*Np + n] + pos[1+n] + time[f % T]
__global__ void __launch_bounds__(256) embed_assemble_kernel(const float* __restrict__ tmp, float* __restrict__ x,
                                                             const float* __restrict__ pos, const float* __restrict__ te,
                                                             int BT, int T, int Np, int d4n) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)BT * Np * d4n) return;
  const int d4 = (int)(idx % d4n);
  const long long r = idx / d4n;
  const int f = (int)(r / Np), n = (int)(r - (long long)f * Np);
  const float4 a = reinterpret_cast<const float4*>(tmp)[idx];
  const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + (long long)(n + 1) * d4n + d4);
  const float4 t = __ldg(reinterpret_cast<const float4*>(te) + (long long)(f % T) * d4n + d4);
  reinterpret_cast<float4*>(x)[((long long)f * (Np + 1) + 1 + n) * d4n + d4] =
      make_float4(a.x + p.x + t.x, a.y + p.y + t.y, a.z + p.z + t.z, a.w + p.w + t.w);
}

int embed_assemble(const float* tmp, float* x, const float* pos, const float* time_embed, int BT, int T, int Np, int D,
                   cudaStream_t stream) {
  if (!tmp || !x || !pos || !time_embed || BT <= 0 || T <= 0 || Np <= 0 || (D & 3)) {
    set_error("embed_assemble: bad arguments");
    return GAVA_ERR_ARG;
  }
  const long long total = (long long)BT * Np * (D / 4);
  embed_assemble_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(tmp, x, pos, time_embed, BT, T, Np, D / 4);
  return check_launch("embed_assemble");
}

// ---------------------------------------------------------------- fp32 QuickGELU in place
__global__ void __launch_bounds__(256) quickgelu_f32_kernel(float* __restrict__ x, long long n4) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  float4 v = reinterpret_cast<float4*>(x)[idx];
  v.x = v.x / (1.0f + expf(-1.702f * v.x));
  v.y = v.y / (1.0f + expf(-1.702f * v.y));
  v.z = v.z / (1.0f + expf(-1.702f * v.z));
  v.w = v.w / (1.0f + expf(-1.702f * v.w));
  reinterpret_cast<float4*>(x)[idx] = v;
}

int quickgelu_f32(float* x, long long n, cudaStream_t stream) {
  if (!x || n < 0 || (n & 3)) { set_error("quickgelu_f32: bad arguments"); return GAVA_ERR_ARG; }
  if (n == 0) return GAVA_OK;
  quickgelu_f32_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, stream>>>(x, n / 4);
  return check_launch("quickgelu_f32");
}

// ---------------------------------------------------------------- fp32 attention cores (CUDA cores)
// One block per (frame, head) [main] or (video, head) [summary, n_aux == 0]: keys/values staged in shared memory
// (row pitch 65 floats: conflict-free), one warp per query row: lane-parallel dot products over the keys,
// warp-shuffle softmax, lane-parallel weighted sum over the head dimension.
struct AttnF32Args {
  const float* qkv;      // [rows, 3D]
  const float* kv_aux;   // [G + 2 BT, 2D] or null
  float* out;            // [rows, D]
  int T, N0, G, H, n_q, BT, S;
  int causal;            // 1: query q attends keys 0..q only (text transformer); needs kv_aux == NULL
  int groups;            // bit 0: local-prompt keys take part, bit 1: the summary-token key takes part
};

__global__ void __launch_bounds__(256) attention_f32_kernel(const AttnF32Args p) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, f = blockIdx.y;
  const int D = p.H * 64;
  const int S = p.S;
  float* sK = sm;                  // [S][65]
  float* sV = sK + S * 65;         // [S][65]
  float* sQ = sV + S * 65;         // [8][64]
  float* sP = sQ + 8 * 64;         // [8][S]
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int b = f / p.T;
  for (int i = threadIdx.x; i < S * 64; i += 256) {
    const int row = i >> 6, d = i & 63;
    const float* ksrc;
    if (row < p.N0) {
      ksrc = p.qkv + ((long long)f * p.N0 + row) * 3 * D + D + h * 64 + d;
      sK[row * 65 + d] = ksrc[0];
      sV[row * 65 + d] = ksrc[D];
    } else {
      const int a = row - p.N0;
      const int arow = a < p.G ? a : (a < p.G + p.T ? p.G + b * p.T + (a - p.G) : p.G + p.BT + f);
      ksrc = p.kv_aux + (long long)arow * 2 * D + h * 64 + d;
      sK[row * 65 + d] = ksrc[0];
      sV[row * 65 + d] = ksrc[D];
    }
  }
  __syncthreads();
  for (int q = warp; q < p.n_q; q += 8) {
    const float* qsrc = p.qkv + ((long long)f * p.N0 + q) * 3 * D + h * 64;
    sQ[warp * 64 + lane] = qsrc[lane] * 0.125f;        // q / sqrt(dh), dh = 64 (exact: power of two)
    sQ[warp * 64 + 32 + lane] = qsrc[32 + lane] * 0.125f;
    __syncwarp();
    const int Sq = p.causal ? q + 1 : S;      // keys visible to this query
    float mx = -INFINITY;
    for (int j = lane; j < Sq; j += 32) {
      float acc = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; ++d) acc = fmaf(sQ[warp * 64 + d], sK[j * 65 + d], acc);
      if (p.kv_aux != nullptr && j >= p.N0 + p.G && !((j < p.N0 + p.G + p.T) ? (p.groups & 1) : (p.groups & 2))) acc = -INFINITY;
      sP[warp * S + j] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Sq; j += 32) {
      const float e = expf(sP[warp * S + j] - mx);
      sP[warp * S + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Sq; ++j) {
      const float pj = sP[warp * S + j];
      o0 = fmaf(pj, sV[j * 65 + lane], o0);
      o1 = fmaf(pj, sV[j * 65 + 32 + lane], o1);
    }
    float* o = p.out + ((long long)f * p.N0 + q) * D + h * 64;
    o[lane] = o0 / sum;
    o[32 + lane] = o1 / sum;
    __syncwarp();
  }
}

static int attention_f32_impl(const float* qkv, const float* kv_aux, float* out, int B, int T, int N0, int G, int H,
                              int n_q, int causal, int groups, cudaStream_t stream);

int attention_f32(const float* qkv, const float* kv_aux, float* out, int B, int T, int N0, int G, int H, int n_q,
                  int key_groups, cudaStream_t stream) {
  return attention_f32_impl(qkv, kv_aux, out, B, T, N0, G, H, n_q, 0, key_groups, stream);
}

// causal self-attention over n_seq sequences of L tokens (CLIP text transformer, VitaCLIP_text_encoder.py:67-88)
int causal_attention_f32(const float* qkv, float* out, int n_seq, int L, int H, cudaStream_t stream) {
  static const bool use_mma = [] { const char* e = getenv("GAVA_TEXT_ATTN_MMA"); return !(e && e[0] == '0'); }();
  if (use_mma && qkv && out && n_seq > 0 && causal_attention_mma_supported(L, H))
    return causal_attention_mma(qkv, out, n_seq, L, H, stream);
  return attention_f32_impl(qkv, nullptr, out, n_seq, 1, L, 0, H, L, 1, 3, stream);
}

static int attention_f32_impl(const float* qkv, const float* kv_aux, float* out, int B, int T, int N0, int G, int H,
                              int n_q, int causal, int groups, cudaStream_t stream) {
  if (!qkv || !out || B <= 0 || T <= 0 || N0 <= 0 || G < 0 || H <= 0 || n_q <= 0 || n_q > N0 || (causal && kv_aux)) {
    set_error("attention_f32: bad arguments");
    return GAVA_ERR_ARG;
  }
  AttnF32Args a;
  a.causal = causal;
  a.groups = groups;
  a.qkv = qkv; a.kv_aux = kv_aux; a.out = out;
  a.T = T; a.N0 = N0; a.G = G; a.H = H; a.n_q = n_q; a.BT = B * T;
  a.S = kv_aux ? N0 + G + T + 1 : N0;       // kv_aux == NULL: plain self-attention over the N0 rows (summary attention)
  const size_t smem = (size_t)(2 * a.S * 65 + 8 * 64 + 8 * a.S) * sizeof(float);
  if (smem > 220 * 1024) { set_error("attention_f32: %d keys exceed shared memory", a.S); return GAVA_ERR_SHAPE; }
  static size_t attr_dev[64] = {0};
  size_t& attr = attr_dev[device_slot()];
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(attention_f32): %s", cudaGetErrorString(e)); return (int)e; }
    attr = smem;
  }
  attention_f32_kernel<<<dim3(H, B * T), 256, smem, stream>>>(a);
  return check_launch("attention_f32");
}

// ---------------------------------------------------------------- fp32 prompt rows (+norm1), fp32 out
__global__ void __launch_bounds__(256) prompt_rows_f32_kernel(const float* __restrict__ gp, const float* __restrict__ lp,
                                                              const float* __restrict__ cp, const float* __restrict__ sm,
                                                              const float* __restrict__ w, const float* __restrict__ b,
                                                              float* __restrict__ out, int G, int BT, int T, int D, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= G + 2 * BT) return;
  const int lane = lane_id();
  float* o = out + (long long)row * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    float v;
    if (row < G) v = gp[(long long)row * D + d];
    else if (row < G + BT) v = lp[(long long)((row - G) % T) * D + d] + cp[(long long)(row - G) * D + d];
    else v = sm[(long long)(row - G - BT) * D + d];
    o[d] = v;
    s += v;
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int d = lane; d < D; d += 32) { const float t = o[d] - mean; q += t * t; }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  for (int d = lane; d < D; d += 32) o[d] = (o[d] - mean) * rstd * w[d] + b[d];
}

int prompt_rows_f32(const float* global_prompts, const float* local_prompts, const float* cls_proj_out,
                    const float* summary, const float* ln_w, const float* ln_b, float* out, int G, int B, int T, int D,
                    float eps, cudaStream_t stream) {
  if (!global_prompts || !local_prompts || !cls_proj_out || !summary || !ln_w || !ln_b || !out || G < 0 || B <= 0 || T <= 0 || D <= 0) {
    set_error("prompt_rows_f32: bad arguments");
    return GAVA_ERR_ARG;
  }
  const int rows = G + 2 * B * T;
  prompt_rows_f32_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(global_prompts, local_prompts, cls_proj_out, summary, ln_w,
                                                             ln_b, out, G, B * T, T, D, eps);
  return check_launch("prompt_rows_f32");
}

}  // namespace gava
