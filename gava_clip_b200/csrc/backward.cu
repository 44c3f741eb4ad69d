// Backward kernels of the fine-tune step: gradients flow through the FROZEN backbone (dgrad only)
// to the trainable prompt / summary parameters.  Dense dgrads and the small wgrads reuse the tcgen05 GEMM
// (gemm_tcgen05.cu) on pre-transposed operands; this file holds everything else:
//   layernorm_bwd            LayerNorm input gradient (+ optional weight/bias gradients, + bf16 copy for the next GEMM)
//   attention_bwd            flash-style backward of the prompt-prefixed MHSA core (scores recomputed on the SM)
//   temporal_attention_bwd   backward of the T x T summary-token attention
//   prompt_rows_bwd          backward of the K/V-only prompt-row construction + norm1
//   group_sum                strided row reductions (time_embed / local_prompts / bias gradients)
//   transpose_bf16           [R, C] -> bf16 [C, Rpad] (operands of the wgrad GEMMs)
//   quickgelu_bf16           h = QuickGELU(a1) when the training forward keeps the fc1 pre-activations
//   adamw_step               fused AdamW over a flat parameter buffer (torch.optim.AdamW semantics, train.py:351)
// Restates autograd of training/VitaCLIP_vision_encoder_utils.py:61-81,155-203 for the parameters left trainable by
// VitaCLIP_model.py:230-235 (names containing summary / local / global / time_embed).
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace gava {

// ================================================================================================
// LayerNorm backward.  y = (x - mean) * rstd * w + b;  g = dy * w;
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
// One warp per row, lane owns float4 #(lane + 32 j) (D = 128 V).  dx_out = [add1] + [add2] + dx.
template <int V, bool DY_BF16>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* __restrict__ dy, long long dy_stride,
                                                            const float* __restrict__ x, long long x_stride,
                                                            const float* __restrict__ w,
                                                            const float* add1, long long add1_stride,
                                                            const float* add2, long long add2_stride,
                                                            float* dx_out, long long dx_stride,
                                                            __nv_bfloat16* dx_bf16, long long dxb_stride,
                                                            float* __restrict__ dw, float* __restrict__ db,
                                                            long long rows, float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = lane_id();
  constexpr float inv_d = 1.0f / (128.0f * V);
  float4 xv[V], g[V];
  const float4* xs = reinterpret_cast<const float4*>(x + row * x_stride);
#pragma unroll
  for (int j = 0; j < V; ++j) xv[j] = xs[lane + 32 * j];
  if constexpr (DY_BF16) {
    const uint2* s = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + row * dy_stride);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint2 u = s[lane + 32 * j];
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), c = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
      g[j] = make_float4(__bfloat162float(a.x), __bfloat162float(a.y), __bfloat162float(c.x), __bfloat162float(c.y));
    }
  } else {
    const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * dy_stride);
#pragma unroll
    for (int j = 0; j < V; ++j) g[j] = s[lane + 32 * j];
  }
  // the residual-gradient addend is fetched together with x and dy (one round of memory latency instead of a second
  // one after the three warp reductions); it may alias dx_out (each lane reads exactly the elements it writes)
  float4 av[V];
  const float4* a1 = add1 ? reinterpret_cast<const float4*>(add1 + row * add1_stride) : nullptr;
  if (a1) {
#pragma unroll
    for (int j = 0; j < V; ++j) av[j] = a1[lane + 32 * j];
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) sum += (xv[j].x + xv[j].y) + (xv[j].z + xv[j].w);
  const float mean = warp_sum(sum) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    xv[j].x -= mean; xv[j].y -= mean; xv[j].z -= mean; xv[j].w -= mean;
    q += (xv[j].x * xv[j].x + xv[j].y * xv[j].y) + (xv[j].z * xv[j].z + xv[j].w * xv[j].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    xv[j].x *= rstd; xv[j].y *= rstd; xv[j].z *= rstd; xv[j].w *= rstd;   // xhat
    if (dw != nullptr) {   // weight / bias gradients (summary_ln): a few hundred rows, plain atomics
      float* pw = dw + 4 * (lane + 32 * j);
      float* pb = db + 4 * (lane + 32 * j);
      atomicAdd(pw + 0, g[j].x * xv[j].x); atomicAdd(pw + 1, g[j].y * xv[j].y);
      atomicAdd(pw + 2, g[j].z * xv[j].z); atomicAdd(pw + 3, g[j].w * xv[j].w);
      atomicAdd(pb + 0, g[j].x); atomicAdd(pb + 1, g[j].y); atomicAdd(pb + 2, g[j].z); atomicAdd(pb + 3, g[j].w);
    }
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * j);
    g[j].x *= ww.x; g[j].y *= ww.y; g[j].z *= ww.z; g[j].w *= ww.w;
    s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
    s2 += (g[j].x * xv[j].x + g[j].y * xv[j].y) + (g[j].z * xv[j].z + g[j].w * xv[j].w);
  }
  const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
  float4* o = reinterpret_cast<float4*>(dx_out + row * dx_stride);
  const float4* a2 = add2 ? reinterpret_cast<const float4*>(add2 + row * add2_stride) : nullptr;
  uint2* ob = dx_bf16 ? reinterpret_cast<uint2*>(dx_bf16 + row * dxb_stride) : nullptr;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float4 d;
    d.x = rstd * (g[j].x - c1 - xv[j].x * c2);
    d.y = rstd * (g[j].y - c1 - xv[j].y * c2);
    d.z = rstd * (g[j].z - c1 - xv[j].z * c2);
    d.w = rstd * (g[j].w - c1 - xv[j].w * c2);
    if (a1) { d.x += av[j].x; d.y += av[j].y; d.z += av[j].z; d.w += av[j].w; }
    if (a2) { const float4 t = a2[lane + 32 * j]; d.x += t.x; d.y += t.y; d.z += t.z; d.w += t.w; }
    o[lane + 32 * j] = d;
    if (ob) ob[lane + 32 * j] = make_uint2(pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
  }
}

int layernorm_bwd(const void* dy, long long dy_stride, int dy_bf16, const float* x, long long x_stride, const float* w,
                  const float* add1, long long add1_stride, const float* add2, long long add2_stride, float* dx_out,
                  long long dx_stride, void* dx_bf16, long long dxb_stride, float* dw, float* db, long long rows, int D,
                  float eps, cudaStream_t stream) {
  if (!dy || !x || !w || !dx_out || rows < 0 || ((dw != nullptr) != (db != nullptr))) {
    set_error("layernorm_bwd: bad arguments");
    return GAVA_ERR_ARG;
  }
  if (rows == 0) return GAVA_OK;
  if ((dy_stride & 3) || (x_stride & 3) || (add1_stride & 3) || (add2_stride & 3) || (dx_stride & 3) || (dxb_stride & 3)) {
    set_error("layernorm_bwd: row strides must be multiples of 4");
    return GAVA_ERR_ARG;
  }
  const unsigned grid = (unsigned)((rows + 7) / 8);
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
#define GAVA_LNB_CASE(V)                                                                                              \
  case 128 * V:                                                                                                       \
    if (dy_bf16)                                                                                                      \
      layernorm_bwd_kernel<V, true><<<grid, 256, 0, stream>>>(dy, dy_stride, x, x_stride, w, add1, add1_stride, add2, \
                                                              add2_stride, dx_out, dx_stride, dxb, dxb_stride, dw, db, rows, eps); \
    else                                                                                                              \
      layernorm_bwd_kernel<V, false><<<grid, 256, 0, stream>>>(dy, dy_stride, x, x_stride, w, add1, add1_stride, add2, \
                                                               add2_stride, dx_out, dx_stride, dxb, dxb_stride, dw, db, rows, eps); \
    break;
  switch (D) {
    GAVA_LNB_CASE(1) GAVA_LNB_CASE(2) GAVA_LNB_CASE(3) GAVA_LNB_CASE(4) GAVA_LNB_CASE(6) GAVA_LNB_CASE(8)
    default: set_error("layernorm_bwd: unsupported width D=%d", D); return GAVA_ERR_SHAPE;
  }
#undef GAVA_LNB_CASE
  return check_launch("layernorm_bwd");
}

// ================================================================================================
// Prompt-row backward: rows [G global | BT local (+cls_proj) | BT summary] were norm1-ed and projected to K/V; given the
// gradient of the normalised rows, undo norm1 and scatter:
//   d_global[g]   = d(row g)                                  (trainable, this layer's slice of visual.global_prompts)
//   d_local_cp[f] = d(row G + f)        (-> local_prompts[t] summed over videos by group_sum, and -> cls_proj output)
//   d_summary[f]  = d(row G + BT + f) [+ d_summary_ext[f]]    (ext: gradient of the encoder's `summary` output)
template <int V>
__global__ void __launch_bounds__(256) prompt_rows_bwd_kernel(const __nv_bfloat16* __restrict__ d_rows_n,
                                                              const float* __restrict__ gp, const float* __restrict__ lp,
                                                              const float* __restrict__ cp, const float* __restrict__ sm,
                                                              const float* __restrict__ w, float* __restrict__ d_global,
                                                              float* __restrict__ d_local_cp, float* __restrict__ d_summary,
                                                              const float* __restrict__ d_summary_ext, int G, int BT, int T,
                                                              float eps) {
  constexpr int D = 128 * V;
  constexpr float inv_d = 1.0f / D;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= G + 2 * BT) return;
  const int lane = lane_id();
  float4 v[V], g[V];
  float* dst;
  const float* ext = nullptr;
  if (row < G) {
    const float4* s = reinterpret_cast<const float4*>(gp + (long long)row * D);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = __ldg(s + lane + 32 * j);
    dst = d_global + (long long)row * D;
  } else if (row < G + BT) {
    const int f = row - G;
    const float4* s0 = reinterpret_cast<const float4*>(lp + (long long)(f % T) * D);
    const float4* s1 = reinterpret_cast<const float4*>(cp + (long long)f * D);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 a = __ldg(s0 + lane + 32 * j), c = s1[lane + 32 * j];
      v[j] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
    }
    dst = d_local_cp + (long long)f * D;
  } else {
    const int f = row - G - BT;
    const float4* s = reinterpret_cast<const float4*>(sm + (long long)f * D);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = s[lane + 32 * j];
    dst = d_summary + (long long)f * D;
    if (d_summary_ext) ext = d_summary_ext + (long long)f * D;
  }
  const uint2* gs = reinterpret_cast<const uint2*>(d_rows_n + (long long)row * D);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint2 u = gs[lane + 32 * j];
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), c = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * j);
    g[j] = make_float4(__bfloat162float(a.x) * ww.x, __bfloat162float(a.y) * ww.y, __bfloat162float(c.x) * ww.z,
                       __bfloat162float(c.y) * ww.w);
    sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(sum) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    v[j].x *= rstd; v[j].y *= rstd; v[j].z *= rstd; v[j].w *= rstd;
    s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
    s2 += (g[j].x * v[j].x + g[j].y * v[j].y) + (g[j].z * v[j].z + g[j].w * v[j].w);
  }
  const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
  float4* o = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float4 d;
    d.x = rstd * (g[j].x - c1 - v[j].x * c2);
    d.y = rstd * (g[j].y - c1 - v[j].y * c2);
    d.z = rstd * (g[j].z - c1 - v[j].z * c2);
    d.w = rstd * (g[j].w - c1 - v[j].w * c2);
    if (ext) { const float4 t = reinterpret_cast<const float4*>(ext)[lane + 32 * j]; d.x += t.x; d.y += t.y; d.z += t.z; d.w += t.w; }
    o[lane + 32 * j] = d;
  }
}

int prompt_rows_bwd(const void* d_rows_n, const float* global_prompts, const float* local_prompts,
                    const float* cls_proj_out, const float* summary, const float* ln_w, float* d_global,
                    float* d_local_cp, float* d_summary, const float* d_summary_ext, int G, int B, int T, int D, float eps,
                    cudaStream_t stream) {
  if (!d_rows_n || !global_prompts || !local_prompts || !cls_proj_out || !summary || !ln_w || !d_global || !d_local_cp ||
      !d_summary || G < 0 || B <= 0 || T <= 0) {
    set_error("prompt_rows_bwd: bad arguments");
    return GAVA_ERR_ARG;
  }
  const int BT = B * T, rows = G + 2 * BT;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const __nv_bfloat16* dn = reinterpret_cast<const __nv_bfloat16*>(d_rows_n);
#define GAVA_PRB_CASE(V) \
  case 128 * V: prompt_rows_bwd_kernel<V><<<grid, 256, 0, stream>>>(dn, global_prompts, local_prompts, cls_proj_out, summary, ln_w, d_global, d_local_cp, d_summary, d_summary_ext, G, BT, T, eps); break;
  switch (D) {
    GAVA_PRB_CASE(1) GAVA_PRB_CASE(2) GAVA_PRB_CASE(3) GAVA_PRB_CASE(4) GAVA_PRB_CASE(6) GAVA_PRB_CASE(8)
    default: set_error("prompt_rows_bwd: unsupported width D=%d", D); return GAVA_ERR_SHAPE;
  }
#undef GAVA_PRB_CASE
  return check_launch("prompt_rows_bwd");
}

// ================================================================================================
// group_sum: out[j, :] = sum_{o < outer} sum_{n < inner} in[((o * J + j) * inner + n) * ld + :]   (f32 or bf16 in, f32 out)
// grid (J, splits); a block reduces a slice of the (o, n) pairs in registers and adds it atomically (out zeroed first).
template <bool IN_BF16>
__global__ void __launch_bounds__(256) group_sum_kernel(const void* __restrict__ in, long long ld, float* __restrict__ out,
                                                        int outer, int J, int inner, int D) {
  const int j = blockIdx.x;
  const long long pairs = (long long)outer * inner;
  const long long per = (pairs + gridDim.y - 1) / gridDim.y;
  const long long p0 = per * blockIdx.y, p1 = (p0 + per < pairs) ? p0 + per : pairs;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (long long p = p0; p < p1; ++p) {
      const long long o = p / inner, n = p - o * inner;
      const long long r = (o * J + j) * inner + n;
      if constexpr (IN_BF16) acc += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in)[r * ld + c]);
      else acc += reinterpret_cast<const float*>(in)[r * ld + c];
    }
    if (p1 > p0) atomicAdd(out + (long long)j * D + c, acc);
  }
}

int group_sum(const void* in, long long ld, int in_bf16, float* out, int outer, int J, int inner, int D,
              cudaStream_t stream) {
  if (!in || !out || outer <= 0 || J <= 0 || inner <= 0 || D <= 0) { set_error("group_sum: bad arguments"); return GAVA_ERR_ARG; }
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)J * D, stream);
  if (e != cudaSuccess) { set_error("group_sum memset: %s", cudaGetErrorString(e)); return (int)e; }
  const long long pairs = (long long)outer * inner;
  long long splits = (4LL * num_sms() + J - 1) / J;
  if (splits > (pairs + 7) / 8) splits = (pairs + 7) / 8;
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)J, (unsigned)splits);
  if (in_bf16) group_sum_kernel<true><<<grid, 256, 0, stream>>>(in, ld, out, outer, J, inner, D);
  else group_sum_kernel<false><<<grid, 256, 0, stream>>>(in, ld, out, outer, J, inner, D);
  return check_launch("group_sum");
}

// ================================================================================================
// transpose_bf16: in [R, C] (f32 or bf16, row pitch ld) -> out bf16 [C, Rpad], columns r >= R zero-filled.
template <bool IN_BF16>
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const void* __restrict__ in, long long ld,
                                                             __nv_bfloat16* __restrict__ out, int R, int C, int Rpad) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 row groups
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      if constexpr (IN_BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in)[(long long)r * ld + c]);
      else v = reinterpret_cast<const float*>(in)[(long long)r * ld + c];
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < C && r < Rpad) out[(long long)c * Rpad + r] = __float2bfloat16(tile[tx][ty + 8 * i]);
  }
}

int transpose_bf16(const void* in, long long ld, int in_bf16, void* out, int R, int C, int Rpad, cudaStream_t stream) {
  if (!in || !out || R <= 0 || C <= 0 || Rpad < R) { set_error("transpose_bf16: bad arguments"); return GAVA_ERR_ARG; }
  dim3 grid((unsigned)((Rpad + 31) / 32), (unsigned)((C + 31) / 32));
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (in_bf16) transpose_bf16_kernel<true><<<grid, 256, 0, stream>>>(in, ld, o, R, C, Rpad);
  else transpose_bf16_kernel<false><<<grid, 256, 0, stream>>>(in, ld, o, R, C, Rpad);
  return check_launch("transpose_bf16");
}

// ================================================================================================
// h = QuickGELU(a1), bf16 -> bf16, 8 elements per thread
__global__ void __launch_bounds__(256) quickgelu_bf16_kernel(const uint4* __restrict__ a, uint4* __restrict__ h, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 u = a[i];
  uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
    w[k] = pack_bf16x2(quick_gelu(__bfloat162float(v.x)), quick_gelu(__bfloat162float(v.y)));
  }
  h[i] = u;
}

int quickgelu_bf16(const void* a, void* h, long long n, cudaStream_t stream) {
  if (!a || !h || n < 0 || (n & 7)) { set_error("quickgelu_bf16: bad arguments (n %% 8 == 0)"); return GAVA_ERR_ARG; }
  if (n == 0) return GAVA_OK;
  const long long n8 = n >> 3;
  quickgelu_bf16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(a),
                                                                         reinterpret_cast<uint4*>(h), n8);
  return check_launch("quickgelu_bf16");
}

// ================================================================================================
// Fused AdamW over a flat f32 buffer (decoupled weight decay, bias-corrected; torch.optim.AdamW, train.py:351):
//   g *= grad_scale; p *= 1 - lr wd; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adamw_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                    float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                    float lr, float b1, float b2, float eps, float wd, float inv_bc1,
                                                    float inv_sqrt_bc2, float grad_scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
  const float decay = 1.0f - lr * wd, step = lr * inv_bc1;
#define GAVA_ADAM1(c)                                               \
  {                                                                 \
    const float gr = gg.c * grad_scale;                             \
    mm.c = b1 * mm.c + (1.0f - b1) * gr;                            \
    vv.c = b2 * vv.c + (1.0f - b2) * gr * gr;                       \
    pp.c = pp.c * decay - step * mm.c / (sqrtf(vv.c) * inv_sqrt_bc2 + eps); \
  }
  GAVA_ADAM1(x) GAVA_ADAM1(y) GAVA_ADAM1(z) GAVA_ADAM1(w)
#undef GAVA_ADAM1
  p[i] = pp; m[i] = mm; v[i] = vv;
}

int adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int step, float grad_scale, cudaStream_t stream) {
  if (!p || !g || !m || !v || n < 0 || (n & 3) || step < 1) { set_error("adamw_step: bad arguments (n %% 4 == 0, step >= 1)"); return GAVA_ERR_ARG; }
  if (n == 0) return GAVA_OK;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  const long long n4 = n >> 2;
  adamw_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), n4, lr, beta1, beta2, eps, weight_decay, (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)),
      grad_scale);
  return check_launch("adamw_step");
}

// ================================================================================================
// Small-sequence attention backward: one block per (sequence, head), everything in shared memory in f32.
//   P = softmax(q k^T / 8 [+ causal mask]); o = P v.
//   dV = P^T dO; dP = dO V^T; dS = P (dP - rowsum(P dP)); dQ = dS K / 8; dK = dS^T Q / 8
// Used for the T x T summary-token attention of the video tower (bf16 qkv, no mask) and for the 77 x 77 causal attention
// of the CLIP text transformer (f32 qkv, VitaCLIP_text_encoder.py:83-88,146-152).
constexpr int SAB_THREADS = 256;
template <bool IN_F32, bool CAUSAL>
__global__ void __launch_bounds__(SAB_THREADS) seq_attention_bwd_kernel(const void* __restrict__ qkv_,
                                                                const __nv_bfloat16* __restrict__ dout,
                                                                __nv_bfloat16* __restrict__ dqkv, int T, int H) {
  extern __shared__ float tsm[];
  const int PITCH = 65, PP = T + 1;
  float* sq = tsm;
  float* sk = sq + T * PITCH;
  float* sv = sk + T * PITCH;
  float* sdo = sv + T * PITCH;
  float* sp = sdo + T * PITCH;      // [T][T+1]  P, then dS
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ld = 3LL * D;
  const long long base = (long long)b * T * ld + h * 64;
  const __nv_bfloat16* dbase = dout + (long long)b * T * D + h * 64;
  for (int i = tid; i < T * 64; i += SAB_THREADS) {
    const int t = i >> 6, d = i & 63;
    if constexpr (IN_F32) {
      const float* src = reinterpret_cast<const float*>(qkv_) + base;
      sq[t * PITCH + d] = src[t * ld + d];
      sk[t * PITCH + d] = src[t * ld + D + d];
      sv[t * PITCH + d] = src[t * ld + 2 * D + d];
    } else {
      const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(qkv_) + base;
      sq[t * PITCH + d] = __bfloat162float(src[t * ld + d]);
      sk[t * PITCH + d] = __bfloat162float(src[t * ld + D + d]);
      sv[t * PITCH + d] = __bfloat162float(src[t * ld + 2 * D + d]);
    }
    sdo[t * PITCH + d] = __bfloat162float(dbase[(long long)t * D + d]);
  }
  __syncthreads();
  for (int idx = tid; idx < T * T; idx += SAB_THREADS) {
    const int i = idx / T, j = idx - i * T;
    float acc = 0.f;
    if (!CAUSAL || j <= i) {
#pragma unroll 16
      for (int d = 0; d < 64; ++d) acc += sq[i * PITCH + d] * sk[j * PITCH + d];
    }
    sp[i * PP + j] = acc * 0.125f;
  }
  __syncthreads();
  for (int i = warp; i < T; i += SAB_THREADS / 32) {   // row softmax over the visible keys
    const int n = CAUSAL ? i + 1 : T;
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, sp[i * PP + j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < n; j += 32) { const float e = __expf(sp[i * PP + j] - m); sp[i * PP + j] = e; s += e; }
    s = 1.0f / warp_sum(s);
    for (int j = lane; j < T; j += 32) sp[i * PP + j] = (j < n) ? sp[i * PP + j] * s : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* obase = dqkv + (long long)b * T * ld + h * 64;
  for (int idx = tid; idx < T * 64; idx += SAB_THREADS) {   // dV[j][d] = sum_i P[i][j] dO[i][d]
    const int j = idx >> 6, d = idx & 63;
    float acc = 0.f;
    for (int i = CAUSAL ? j : 0; i < T; ++i) acc += sp[i * PP + j] * sdo[i * PITCH + d];
    obase[j * ld + 2 * D + d] = __float2bfloat16(acc);
  }
  __syncthreads();
  for (int i = warp; i < T; i += SAB_THREADS / 32) {   // dS row i (overwrites P row i)
    const int n = CAUSAL ? i + 1 : T;
    float delta = 0.f;
    float dp_loc[4];   // T <= 128
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + 32 * jj;
      float acc = 0.f;
      if (j < n) {
#pragma unroll 16
        for (int d = 0; d < 64; ++d) acc += sdo[i * PITCH + d] * sv[j * PITCH + d];
        delta += sp[i * PP + j] * acc;
      }
      dp_loc[jj] = acc;
    }
    delta = warp_sum(delta);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + 32 * jj;
      if (j < n) sp[i * PP + j] = sp[i * PP + j] * (dp_loc[jj] - delta);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < T * 64; idx += SAB_THREADS) {
    const int t = idx >> 6, d = idx & 63;
    float aq = 0.f, ak = 0.f;
    for (int j = 0; j < T; ++j) {      // masked entries of dS are exactly 0
      aq += sp[t * PP + j] * sk[j * PITCH + d];   // dQ[t] = sum_j dS[t][j] K[j]
      ak += sp[j * PP + t] * sq[j * PITCH + d];   // dK[t] = sum_i dS[i][t] Q[i]
    }
    obase[t * ld + d] = __float2bfloat16(aq * 0.125f);
    obase[t * ld + D + d] = __float2bfloat16(ak * 0.125f);
  }
}

template <bool IN_F32, bool CAUSAL>
static int launch_seq_attention_bwd(const void* qkv, const void* dout, void* dqkv, int n_seq, int T, int H,
                                    cudaStream_t stream, const char* what) {
  if (!qkv || !dout || !dqkv || n_seq <= 0 || T <= 0 || H <= 0) { set_error("%s: bad arguments", what); return GAVA_ERR_ARG; }
  if (T > 128) { set_error("%s: sequence length %d > 128 unsupported", what, T); return GAVA_ERR_SHAPE; }
  const size_t smem = sizeof(float) * ((size_t)4 * T * 65 + (size_t)T * (T + 1));
  static size_t attr_bytes_dev[64] = {0};
  size_t& attr_bytes = attr_bytes_dev[device_slot()];
  if (attr_bytes == 0) attr_bytes = 48 * 1024;
  auto kern = seq_attention_bwd_kernel<IN_F32, CAUSAL>;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(%s): %s", what, cudaGetErrorString(e)); return (int)e; }
    attr_bytes = smem;
  }
  kern<<<dim3(H, n_seq), SAB_THREADS, smem, stream>>>(qkv, reinterpret_cast<const __nv_bfloat16*>(dout),
                                              reinterpret_cast<__nv_bfloat16*>(dqkv), T, H);
  return check_launch(what);
}

int temporal_attention_bwd(const void* qkv, const void* dout, void* dqkv, int B, int T, int H, cudaStream_t stream) {
  static const bool use_mma = [] { const char* e = getenv("GAVA_TEXT_ATTN_MMA"); return !(e && e[0] == '0'); }();
  if (use_mma && qkv && dout && dqkv && B > 0 && H > 0 && T > 16 && causal_attention_mma_supported(T, H))
    return temporal_attention_bwd_mma(qkv, dout, dqkv, B, T, H, stream);
  return launch_seq_attention_bwd<false, false>(qkv, dout, dqkv, B, T, H, stream, "temporal_attention_bwd");
}

int causal_attention_bwd(const float* qkv, const void* dout, void* dqkv, int n_seq, int L, int H, cudaStream_t stream) {
  static const bool use_mma = [] { const char* e = getenv("GAVA_TEXT_ATTN_MMA"); return !(e && e[0] == '0'); }();
  if (use_mma && qkv && dout && dqkv && n_seq > 0 && causal_attention_mma_supported(L, H))
    return causal_attention_bwd_mma(qkv, dout, dqkv, n_seq, L, H, stream);
  return launch_seq_attention_bwd<true, true>(qkv, dout, dqkv, n_seq, L, H, stream, "causal_attention_bwd");
}

}  // namespace gava
