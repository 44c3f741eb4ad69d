// Bandwidth-bound row-wise kernels of the video tower: LayerNorm / cast, patch gather, cls rows,
// prompt (K/V-only) rows, summary-token attention, logit head.  One warp per row, 128-bit coalesced
// accesses, warp-shuffle reductions, fp32 statistics.
#include "common.cuh"
#include "kernels.h"
#include "rowwise.cuh"

namespace gava {

template <int V>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long in_stride,
                                                        const __nv_bfloat16* __restrict__ delta, long long delta_stride,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        void* out, long long out_stride, int out_bf16, long long rows,
                                                        float eps, const float* __restrict__ w2 = nullptr,
                                                        const float* __restrict__ b2 = nullptr,
                                                        __nv_bfloat16* out2 = nullptr, long long out2_stride = 0) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = lane_id();
  const float4* src = reinterpret_cast<const float4*>(x + row * in_stride);
  float4 v[V];
#pragma unroll
  for (int j = 0; j < V; ++j) v[j] = src[lane + 32 * j];
  if (delta != nullptr) {   // x + delta: a residual branch kept as a bf16 tensor (summed here instead of in a GEMM epilogue)
    const uint2* d = reinterpret_cast<const uint2*>(delta + row * delta_stride);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint2 u = d[lane + 32 * j];
      const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
      v[j].x += __bfloat162float(lo.x); v[j].y += __bfloat162float(lo.y);
      v[j].z += __bfloat162float(hi.x); v[j].w += __bfloat162float(hi.y);
    }
  }
  if (w != nullptr) ln_normalize<V>(v, w, b, eps);
  char* o = reinterpret_cast<char*>(out) + row * out_stride * (out_bf16 ? 2 : 4);
  store_row<V>(v, o, out_bf16);
  if (w2 != nullptr) {   // a second LayerNorm of the values just produced (ln_pre followed by block 0's norm1): no re-read
    ln_normalize<V>(v, w2, b2, eps);
    store_row<V>(v, out2 + row * out2_stride, 1);
  }
}

// out = LN(x; w, b) (f32) and out2 = LN(out; w2, b2) (bf16) in one pass over the rows
int layernorm2(const float* x, long long in_row_stride, const float* w, const float* b, float* out, long long out_row_stride,
               const float* w2, const float* b2, void* out2, long long out2_row_stride, long long rows, int D, float eps,
               cudaStream_t stream) {
  if (!x || !out || !w || !b || !w2 || !b2 || !out2 || rows < 0) { set_error("layernorm2: bad arguments"); return GAVA_ERR_ARG; }
  if (rows == 0) return GAVA_OK;
  if ((in_row_stride & 3) || (out_row_stride & 3) || (out2_row_stride & 3)) { set_error("layernorm2: row strides must be multiples of 4"); return GAVA_ERR_ARG; }
  const unsigned grid = (unsigned)((rows + 7) / 8);
  __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(out2);
#define GAVA_LN2_CASE(V) \
  case 128 * V: layernorm_kernel<V><<<grid, 256, 0, stream>>>(x, in_row_stride, nullptr, 0, w, b, out, out_row_stride, 0, rows, eps, w2, b2, o2, out2_row_stride); break;
  switch (D) {
    GAVA_LN2_CASE(1) GAVA_LN2_CASE(2) GAVA_LN2_CASE(3) GAVA_LN2_CASE(4) GAVA_LN2_CASE(6) GAVA_LN2_CASE(8)
    default: set_error("layernorm2: unsupported width D=%d", D); return GAVA_ERR_SHAPE;
  }
#undef GAVA_LN2_CASE
  return check_launch("layernorm2");
}

int layernorm_resid(const float* x, long long in_row_stride, const void* delta, long long delta_row_stride, const float* w,
                    const float* b, void* out, long long out_row_stride, int out_bf16, long long rows, int D, float eps,
                    cudaStream_t stream) {
  if (!x || !out || rows < 0 || (w && !b)) { set_error("layernorm: bad arguments"); return GAVA_ERR_ARG; }
  if (rows == 0) return GAVA_OK;
  if ((in_row_stride & 3) || (out_row_stride & 3) || (delta_row_stride & 3)) { set_error("layernorm: row strides must be multiples of 4"); return GAVA_ERR_ARG; }
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const __nv_bfloat16* dl = reinterpret_cast<const __nv_bfloat16*>(delta);
#define GAVA_LN_CASE(V) \
  case 128 * V: layernorm_kernel<V><<<grid, 256, 0, stream>>>(x, in_row_stride, dl, delta_row_stride, w, b, out, out_row_stride, out_bf16, rows, eps); break;
  switch (D) {
    GAVA_LN_CASE(1) GAVA_LN_CASE(2) GAVA_LN_CASE(3) GAVA_LN_CASE(4) GAVA_LN_CASE(6) GAVA_LN_CASE(8)
    default: set_error("layernorm: unsupported width D=%d (128/256/384/512/768/1024)", D); return GAVA_ERR_SHAPE;
  }
#undef GAVA_LN_CASE
  return check_launch("layernorm");
}

int layernorm(const float* x, long long in_row_stride, const float* w, const float* b, void* out,
              long long out_row_stride, int out_bf16, long long rows, int D, float eps, cudaStream_t stream) {
  return layernorm_resid(x, in_row_stride, nullptr, 0, w, b, out, out_row_stride, out_bf16, rows, D, eps, stream);
}

// ------------------------------------------------------------------------------------------------
// Patch gather: video f32 [B,3,T,H,W] -> A bf16 [B*T*gh*gw, Kpad], col = c*P*P + ph*P + pw.
// P % 8 == 0: one thread moves 8 contiguous pixels (2 x 16-byte loads -> one 16-byte store).
__global__ void __launch_bounds__(256) patch_gather_vec8_kernel(const float* __restrict__ video,
                                                                __nv_bfloat16* __restrict__ A, int B, int T, int H,
                                                                int W, int P, int Kpad) {
  const int w8n = W >> 3;
  const long long total = (long long)B * 3 * T * H * w8n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w8 = (int)(idx % w8n);
  long long r = idx / w8n;
  const int y = (int)(r % H); r /= H;
  const int t = (int)(r % T); r /= T;
  const int c = (int)(r % 3);
  const int b = (int)(r / 3);
  const float4* src = reinterpret_cast<const float4*>(video + idx * 8);
  const float4 a0 = __ldg(src), a1 = __ldg(src + 1);
  const int x = w8 * 8;
  const int gw = W / P, gh = H / P;
  const long long arow = ((long long)(b * T + t) * gh + y / P) * gw + x / P;
  const int col = c * P * P + (y % P) * P + (x % P);
  uint4 o;
  o.x = pack_bf16x2(a0.x, a0.y); o.y = pack_bf16x2(a0.z, a0.w);
  o.z = pack_bf16x2(a1.x, a1.y); o.w = pack_bf16x2(a1.z, a1.w);
  *reinterpret_cast<uint4*>(A + arow * Kpad + col) = o;
}
// any P (ViT-L/14): one thread per pixel
__global__ void __launch_bounds__(256) patch_gather_scalar_kernel(const float* __restrict__ video,
                                                                  __nv_bfloat16* __restrict__ A, int B, int T, int H,
                                                                  int W, int P, int Kpad) {
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
  A[arow * Kpad + col] = __float2bfloat16(__ldg(video + idx));
}

int patch_gather(const float* video, void* A, int B, int T, int H, int W, int P, int Kpad, cudaStream_t stream) {
  if (!video || !A || B <= 0 || T <= 0 || P <= 0 || H < P || W < P || Kpad < 3 * P * P) {
    set_error("patch_gather: bad arguments");
    return GAVA_ERR_ARG;
  }
  const long long rows = (long long)B * T * (H / P) * (W / P);
  if (Kpad > 3 * P * P) {  // zero the K padding (ViT-L/14: 588 -> 640)
    cudaError_t e = cudaMemsetAsync(A, 0, (size_t)rows * Kpad * 2, stream);
    if (e != cudaSuccess) { set_error("patch_gather memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (P % 8 == 0 && W % 8 == 0 && H % P == 0 && W % P == 0 && Kpad % 8 == 0) {
    const long long total = (long long)B * 3 * T * H * (W / 8);
    patch_gather_vec8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        video, reinterpret_cast<__nv_bfloat16*>(A), B, T, H, W, P, Kpad);
  } else {
    const long long total = (long long)B * 3 * T * H * W;
    patch_gather_scalar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        video, reinterpret_cast<__nv_bfloat16*>(A), B, T, H, W, P, Kpad);
  }
  return check_launch("patch_gather");
}

// ------------------------------------------------------------------------------------------------
// Input stage fused into the patch gather: decoder output uint8 [B,T,H,W,3] (HWC) ->
// (u/255 - mean[c]) / std[c] (video_dataset/dataset.py:117-121, fp32, true divisions) -> bf16 patch rows.
// One thread moves 8 pixels x 3 channels (24 contiguous input bytes -> three 16-byte stores).
__global__ void __launch_bounds__(256) patch_gather_u8_vec8_kernel(const uint8_t* __restrict__ frames,
                                                                   __nv_bfloat16* __restrict__ A, float m0, float m1,
                                                                   float m2, float s0, float s1, float s2, int B, int T,
                                                                   int H, int W, int P, int Kpad) {
  const int w8n = W >> 3;
  const long long total = (long long)B * T * H * w8n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w8 = (int)(idx % w8n);
  long long r = idx / w8n;
  const int y = (int)(r % H); r /= H;            // r = b*T + t
  const uint2* src = reinterpret_cast<const uint2*>(frames + idx * 24);
  const uint2 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
  const uint32_t wds[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
  float v[3][8];
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const float u = (float)((wds[i >> 2] >> (8 * (i & 3))) & 0xffu);
    v[i % 3][i / 3] = __fdiv_rn(__fdiv_rn(u, 255.0f) - mean[i % 3], sd[i % 3]);
  }
  const int x = w8 * 8;
  const int gw = W / P, gh = H / P;
  const long long arow = (r * gh + y / P) * gw + x / P;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint4 o;
    o.x = pack_bf16x2(v[c][0], v[c][1]); o.y = pack_bf16x2(v[c][2], v[c][3]);
    o.z = pack_bf16x2(v[c][4], v[c][5]); o.w = pack_bf16x2(v[c][6], v[c][7]);
    *reinterpret_cast<uint4*>(A + arow * Kpad + c * P * P + (y % P) * P + (x % P)) = o;
  }
}
// any patch size; optional lo output (fp32-accuracy path)
__global__ void __launch_bounds__(256) patch_gather_u8_scalar_kernel(const uint8_t* __restrict__ frames,
                                                                     __nv_bfloat16* __restrict__ hi,
                                                                     __nv_bfloat16* __restrict__ lo, float m0, float m1,
                                                                     float m2, float s0, float s1, float s2, int B, int T,
                                                                     int H, int W, int P, int Kpad) {
  const long long total = (long long)B * T * H * W * 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % 3);
  long long r = idx / 3;
  const int x = (int)(r % W); r /= W;
  const int y = (int)(r % H); r /= H;
  const int gw = W / P, gh = H / P;
  if (y >= gh * P || x >= gw * P) return;
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
  const float v = __fdiv_rn(__fdiv_rn((float)frames[idx], 255.0f) - mean, sd);
  const long long off = ((r * gh + y / P) * gw + x / P) * Kpad + c * P * P + (y % P) * P + (x % P);
  const __nv_bfloat16 h = __float2bfloat16(v);
  hi[off] = h;
  if (lo != nullptr) lo[off] = __float2bfloat16(v - __bfloat162float(h));
}

int patch_gather_u8(const uint8_t* frames, void* hi, void* lo, const float* mean, const float* std, int B, int T, int H,
                    int W, int P, int Kpad, cudaStream_t stream) {
  if (!frames || !hi || !mean || !std || B <= 0 || T <= 0 || P <= 0 || H < P || W < P || Kpad < 3 * P * P) {
    set_error("patch_gather_u8: bad arguments");
    return GAVA_ERR_ARG;
  }
  for (int c = 0; c < 3; ++c)
    if (!(std[c] > 0.f)) { set_error("patch_gather_u8: std must be positive"); return GAVA_ERR_ARG; }
  const long long rows = (long long)B * T * (H / P) * (W / P);
  if (Kpad > 3 * P * P) {
    cudaError_t e = cudaMemsetAsync(hi, 0, (size_t)rows * Kpad * 2, stream);
    if (e == cudaSuccess && lo) e = cudaMemsetAsync(lo, 0, (size_t)rows * Kpad * 2, stream);
    if (e != cudaSuccess) { set_error("patch_gather_u8 memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (!lo && P % 8 == 0 && W % 8 == 0 && H % P == 0 && W % P == 0 && Kpad % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(frames) & 7u) == 0) {
    const long long total = (long long)B * T * H * (W / 8);
    patch_gather_u8_vec8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        frames, reinterpret_cast<__nv_bfloat16*>(hi), mean[0], mean[1], mean[2], std[0], std[1], std[2], B, T, H, W, P, Kpad);
  } else {
    const long long total = (long long)B * T * H * W * 3;
    patch_gather_u8_scalar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        frames, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), mean[0], mean[1], mean[2],
        std[0], std[1], std[2], B, T, H, W, P, Kpad);
  }
  return check_launch("patch_gather_u8");
}

// ------------------------------------------------------------------------------------------------
// cls rows: x[f*tpf] = cls_token + pos_embed[0] + time_embed[f % T]   (VitaCLIP_vision_encoder.py:108-111)
__global__ void __launch_bounds__(256) cls_rows_kernel(float* __restrict__ x, int tpf, const float* __restrict__ cls,
                                                       const float* __restrict__ pos, const float* __restrict__ te,
                                                       int BT, int T, int D) {
  const int d4n = D >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)BT * d4n) return;
  const int f = (int)(idx / d4n), d4 = (int)(idx % d4n);
  const float4 a = __ldg(reinterpret_cast<const float4*>(cls) + d4);
  const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + d4);
  const float4 t = __ldg(reinterpret_cast<const float4*>(te + (long long)(f % T) * D) + d4);
  reinterpret_cast<float4*>(x + (long long)f * tpf * D)[d4] =
      make_float4(a.x + p.x + t.x, a.y + p.y + t.y, a.z + p.z + t.z, a.w + p.w + t.w);
}

int cls_rows(float* x, int tokens_per_frame, const float* cls_token, const float* pos_embed, const float* time_embed,
             int BT, int T, int D, cudaStream_t stream) {
  if (!x || !cls_token || !pos_embed || !time_embed || BT <= 0 || T <= 0 || (D & 3)) {
    set_error("cls_rows: bad arguments");
    return GAVA_ERR_ARG;
  }
  const long long total = (long long)BT * (D / 4);
  cls_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, tokens_per_frame, cls_token, pos_embed,
                                                                        time_embed, BT, T, D);
  return check_launch("cls_rows");
}

// ------------------------------------------------------------------------------------------------
// Prompt rows (K/V-only tokens) + norm1:  [G global | B*T local (+cls_proj) | B*T summary] -> bf16
template <int V>
__global__ void __launch_bounds__(256) prompt_rows_kernel(const float* __restrict__ gp, const float* __restrict__ lp,
                                                          const float* __restrict__ cp, const float* __restrict__ sm,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          __nv_bfloat16* __restrict__ out, int G, int BT, int T,
                                                          float eps) {
  constexpr int D = 128 * V;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= G + 2 * BT) return;
  const int lane = lane_id();
  float4 v[V];
  if (row < G) {
    const float4* s = reinterpret_cast<const float4*>(gp + (long long)row * D);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = __ldg(s + lane + 32 * j);
  } else if (row < G + BT) {
    const int f = row - G;
    const float4* s0 = reinterpret_cast<const float4*>(lp + (long long)(f % T) * D);
    const float4* s1 = reinterpret_cast<const float4*>(cp + (long long)f * D);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 a = __ldg(s0 + lane + 32 * j), c = s1[lane + 32 * j];
      v[j] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
    }
  } else {
    const float4* s = reinterpret_cast<const float4*>(sm + (long long)(row - G - BT) * D);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = s[lane + 32 * j];
  }
  ln_normalize<V>(v, w, b, eps);
  store_row<V>(v, out + (long long)row * D, 1);
}

int prompt_rows(const float* global_prompts, const float* local_prompts, const float* cls_proj_out,
                const float* summary, const float* ln_w, const float* ln_b, void* out, int G, int B, int T, int D,
                float eps, cudaStream_t stream) {
  if (!global_prompts || !local_prompts || !cls_proj_out || !summary || !ln_w || !ln_b || !out || G < 0 || B <= 0 || T <= 0) {
    set_error("prompt_rows: bad arguments");
    return GAVA_ERR_ARG;
  }
  const int BT = B * T, rows = G + 2 * BT;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#define GAVA_PR_CASE(V) \
  case 128 * V: prompt_rows_kernel<V><<<grid, 256, 0, stream>>>(global_prompts, local_prompts, cls_proj_out, summary, ln_w, ln_b, o, G, BT, T, eps); break;
  switch (D) {
    GAVA_PR_CASE(1) GAVA_PR_CASE(2) GAVA_PR_CASE(3) GAVA_PR_CASE(4) GAVA_PR_CASE(6) GAVA_PR_CASE(8)
    default: set_error("prompt_rows: unsupported width D=%d", D); return GAVA_ERR_SHAPE;
  }
#undef GAVA_PR_CASE
  return check_launch("prompt_rows");
}

// ------------------------------------------------------------------------------------------------
// Summary-token attention over the T frames of one video (VitaCLIP_vision_encoder_utils.py:71-77 with L = T).
// grid (H, B), 4 warps; K/V of the (video, head) staged in smem (row pitch 66 bf16: conflict-free).
constexpr int TA_MAX_T = 128;
constexpr int TA_PITCH = 66;
__global__ void __launch_bounds__(128) temporal_attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                 __nv_bfloat16* __restrict__ out, int T, int H) {
  __shared__ __nv_bfloat16 sk[TA_MAX_T * TA_PITCH];
  __shared__ __nv_bfloat16 sv[TA_MAX_T * TA_PITCH];
  __shared__ float sq[4][64];
  __shared__ float sp[4][TA_MAX_T];
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * 64;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long ld = 3LL * D;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * 64;
  for (int i = threadIdx.x; i < T * 32; i += 128) {  // bf16x2 granules
    const int t = i >> 5, d2 = i & 31;
    const __nv_bfloat162 kk = *reinterpret_cast<const __nv_bfloat162*>(base + t * ld + D + 2 * d2);
    const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(base + t * ld + 2 * D + 2 * d2);
    *reinterpret_cast<__nv_bfloat162*>(&sk[t * TA_PITCH + 2 * d2]) = kk;
    *reinterpret_cast<__nv_bfloat162*>(&sv[t * TA_PITCH + 2 * d2]) = vv;
  }
  __syncthreads();
  const float scale_log2 = 0.125f * 1.4426950408889634f;  // dh^-0.5 * log2(e), dh = 64
  for (int tq = warp; tq < T; tq += 4) {
    const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(base + tq * ld + 2 * lane);
    sq[warp][2 * lane] = __bfloat162float(q2.x);
    sq[warp][2 * lane + 1] = __bfloat162float(q2.y);
    __syncwarp();
    float s[TA_MAX_T / 32];
    float m = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < TA_MAX_T / 32; ++jj) {
      const int j = lane + 32 * jj;
      float acc = -INFINITY;
      if (j < T) {
        acc = 0.f;
#pragma unroll 8
        for (int d2 = 0; d2 < 32; ++d2) {
          const __nv_bfloat162 kk = *reinterpret_cast<const __nv_bfloat162*>(&sk[j * TA_PITCH + 2 * d2]);
          acc += sq[warp][2 * d2] * __bfloat162float(kk.x) + sq[warp][2 * d2 + 1] * __bfloat162float(kk.y);
        }
        acc *= scale_log2;
      }
      s[jj] = acc;
      m = fmaxf(m, acc);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < TA_MAX_T / 32; ++jj) {
      const int j = lane + 32 * jj;
      const float p = (j < T) ? exp2f(s[jj] - m) : 0.f;
      if (j < T) sp[warp][j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float p = sp[warp][j];
      const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(&sv[j * TA_PITCH + 2 * lane]);
      o0 += p * __bfloat162float(vv.x);
      o1 += p * __bfloat162float(vv.y);
    }
    const float inv = 1.0f / sum;
    *reinterpret_cast<uint32_t*>(out + ((long long)b * T + tq) * D + h * 64 + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

int temporal_attention(const void* qkv, void* out, int B, int T, int H, cudaStream_t stream) {
  if (!qkv || !out || B <= 0 || T <= 0 || H <= 0) { set_error("temporal_attention: bad arguments"); return GAVA_ERR_ARG; }
  if (T > TA_MAX_T) { set_error("temporal_attention: T=%d > %d unsupported", T, TA_MAX_T); return GAVA_ERR_SHAPE; }
  // more than one 16-row tile of frames (the 70-frame recipe): tensor-core form (seq_attention_mma.cu); a handful of
  // frames stays on this scalar kernel (one partially filled MMA tile would not be faster)
  static const bool use_mma = [] { const char* e = getenv("GAVA_TEXT_ATTN_MMA"); return !(e && e[0] == '0'); }();
  if (use_mma && T > 16) return temporal_attention_mma(qkv, out, B, T, H, stream);
  temporal_attention_kernel<<<dim3(H, B), 128, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                            reinterpret_cast<__nv_bfloat16*>(out), T, H);
  return check_launch("temporal_attention");
}

// ------------------------------------------------------------------------------------------------
// Logit head: temporal mean, L2 norm, per-class mean cosine * scale (+bias).
// (VitaCLIP_vision_encoder.py:128-130, VitaCLIP_model.py:255,287-289,303-309).  grid = B, 256 threads.
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ ff, const float* __restrict__ stok,
                                                   const float* __restrict__ text, const int* __restrict__ offs,
                                                   float* __restrict__ vf, float* __restrict__ sout,
                                                   float* __restrict__ logits, int T, int E, int D, int n_cls,
                                                   float scale, float bias) {
  extern __shared__ float sv[];  // [E] video feature
  __shared__ float red[8];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const float inv_t = 1.0f / (float)T;
  float ss = 0.f;
  for (int e = threadIdx.x; e < E; e += 256) {
    float a = 0.f;
    for (int t = 0; t < T; ++t) a += ff[((long long)b * T + t) * E + e];
    a *= inv_t;
    sv[e] = a;
    vf[(long long)b * E + e] = a;
    ss += a * a;
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv_norm = rsqrtf(tot);
  if (stok != nullptr && sout != nullptr) {
    for (int d = threadIdx.x; d < D; d += 256) {
      float a = 0.f;
      for (int t = 0; t < T; ++t) a += stok[((long long)b * T + t) * D + d];
      sout[(long long)b * D + d] = a * inv_t;
    }
  }
  for (int c = warp; c < n_cls; c += 8) {
    const int k0 = offs[c], k1 = offs[c + 1];
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) {
      const float* tr = text + (long long)k * E;
      float d = 0.f;
      for (int e = lane; e < E; e += 32) d += sv[e] * __ldg(tr + e);
      acc += warp_sum(d);
    }
    if (lane == 0) logits[(long long)b * n_cls + c] = scale * inv_norm * acc / (float)(k1 - k0) + bias;
  }
}

int head(const float* frame_feat, const float* summary_tok, const float* text_hat, const int* class_offsets,
         float* video_features, float* summary_out, float* logits, int B, int T, int E, int D, int n_cls, float scale,
         float logit_bias, cudaStream_t stream) {
  if (!frame_feat || !text_hat || !class_offsets || !video_features || !logits || B <= 0 || T <= 0 || E <= 0 || n_cls <= 0) {
    set_error("head: bad arguments");
    return GAVA_ERR_ARG;
  }
  if (E > 8192) { set_error("head: embed dim %d too large", E); return GAVA_ERR_SHAPE; }
  head_kernel<<<B, 256, E * sizeof(float), stream>>>(frame_feat, summary_tok, text_hat, class_offsets, video_features,
                                                     summary_out, logits, T, E, D, n_cls, scale, logit_bias);
  return check_launch("head");
}

}  // namespace gava
