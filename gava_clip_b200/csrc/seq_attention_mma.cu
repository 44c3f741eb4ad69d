// Short-sequence multi-head self-attention (<= 128 tokens, head_dim 64) on the tensor cores, forward and backward:
//   * the causal attention of the CLIP text transformer (77 tokens) -- nn.MultiheadAttention's core with the causal mask,
//     training/VitaCLIP_text_encoder.py:83-88,146-152 (f32 activations);
//   * the summary-token attention over the T frames of a clip, training/VitaCLIP_vision_encoder_utils.py:168-170, for
//     clips longer than one 16-row tile (the shipped 70-frame recipe; bf16 activations, no mask).
//
// One CTA per (sequence, head); the sequence's q, k, v rows (f32 in HBM: the text side keeps fp32 activations) are staged
// once in swizzled shared memory as bf16 and every contraction is mma.sync m16n8k16 (bf16 operands, f32 accumulate).
// The problem is tiny (15 x 8 items of 77 x 77 x 64) and latency-bound: the point of the tensor-core form is the short
// dependent chain per item (the scalar shared-memory kernels they replace took 40 us forward / 103 us backward per layer).
//
//   forward  : fp32-accurate.  q = qh + ql, k = kh + kl, v = vh + vl (bf16 pairs); S = qh kh^T + qh kl^T + ql kh^T,
//              softmax in f32 over the whole row held in registers, P = ph + pl, O = ph vh + ph vl + pl vh.  The dropped
//              lo x lo terms are 2^-18 relative: the text features stay within 1e-4 of the f32 module forward (test).
//   backward : plain bf16 operands (gradient noise 2^-9, like every other backward contraction of the fine-tune step).
//              pre-pass per 16-query tile: lse and delta = sum_j P_ij dP_ij (the forward's O is not kept); then work
//              units dealt to the warps: 16-query tiles (dQ) and 16-key tiles (dK, dV) as in attention_bwd.cu, restricted
//              to the causal half.
// Padding rows (T..t_pad) are zero-filled: their dO = 0 and K = 0 make every contribution vanish; only the causal mask is
// explicit.
#include "common.cuh"
#include "kernels.h"
#include "mma_sync.cuh"

namespace gava {

constexpr int SAM_WARPS = 8;
constexpr int SAM_THREADS = 32 * SAM_WARPS;
constexpr int SAM_MAX_T = 128;
constexpr int SAM_MAX_TILES = SAM_MAX_T / 16;

__device__ __forceinline__ uint4 pack8_bf16(const float4& a, const float4& b) {
  return make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// hi / lo halves of 8 floats: hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split8_bf16(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  hi = pack8_bf16(a, b);
  const float4 ra = make_float4(a.x - bf16_round(a.x), a.y - bf16_round(a.y), a.z - bf16_round(a.z), a.w - bf16_round(a.w));
  const float4 rb = make_float4(b.x - bf16_round(b.x), b.y - bf16_round(b.y), b.z - bf16_round(b.z), b.w - bf16_round(b.w));
  lo = pack8_bf16(ra, rb);
}

// key `kc` takes part in the softmax of query row `r`: the causal mask, or (self-attention over the whole sequence) only
// the padding keys beyond T are excluded
template <bool CAUSAL>
__device__ __forceinline__ bool key_on(int kc, int r, int T) { return CAUSAL ? kc <= r : kc < T; }

// ============================================================================================== forward
// PRECISE: f32 in / out, 3-term split-bf16 products (text transformer); otherwise bf16 in / out, plain bf16 products (the
// tower's summary-token attention over the T frames of a clip, utils.py:168-170, for clips longer than one 16-row tile).
template <bool PRECISE, bool CAUSAL>
__global__ void __launch_bounds__(SAM_THREADS, 1)
seq_attention_mma_kernel(const void* __restrict__ qkv_, void* __restrict__ out_, int T, int H, int t_pad) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * 64;
  const uint32_t plane = (uint32_t)t_pad * 128u;
  constexpr uint32_t PS = PRECISE ? 2u : 1u;       // planes per operand: hi (+ lo)
  const uint32_t sQh = smem_u32(smem), sQl = sQh + plane, sKh = sQh + PS * plane, sKl = sKh + plane,
                 sVh = sQh + 2u * PS * plane, sVl = sVh + plane;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tg = lane & 3, mi = lane >> 3;
  const long long ld = 3LL * D;

  for (int i = tid; i < t_pad * 8 * 3; i += SAM_THREADS) {       // (row, chunk, which of q/k/v)
    const int which = i / (t_pad * 8), j = i - which * (t_pad * 8);
    const int row = j >> 3, ch = j & 7;
    uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
    if (row < T) {
      const long long e = ((long long)b * T + row) * ld + which * D + h * 64 + ch * 8;
      if constexpr (PRECISE) {
        const float* src = reinterpret_cast<const float*>(qkv_) + e;
        split8_bf16(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + 4), hi, lo);
      } else {
        hi = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(qkv_) + e);
      }
    }
    const uint32_t off = kv_off(row, ch) + (uint32_t)which * PS * plane;
    *reinterpret_cast<uint4*>(smem + off) = hi;
    if constexpr (PRECISE) *reinterpret_cast<uint4*>(smem + off + plane) = lo;
  }
  __syncthreads();

  const float scale_log2 = 0.125f * 1.4426950408889634f;
  const int n_t = t_pad >> 4;
  for (int mt = n_t - 1 - warp; mt >= 0; mt -= SAM_WARPS) {      // longest rows first
    uint32_t qh[4][4], ql[PRECISE ? 4 : 1][4];
    {
      const int row = mt * 16 + (lane & 15);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        ldmatrix_x4(sQh + kv_off(row, 2 * ks + (lane >> 4)), qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3]);
        if constexpr (PRECISE) ldmatrix_x4(sQl + kv_off(row, 2 * ks + (lane >> 4)), ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3]);
      }
    }
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const int kt_end = CAUSAL ? mt : n_t - 1;        // last key tile this query tile attends to
    float s[SAM_MAX_TILES][2][4];
    float m0 = -1e30f, m1 = -1e30f;
#pragma unroll
    for (int kt = 0; kt < SAM_MAX_TILES; ++kt) {
      if (kt <= kt_end) {
#pragma unroll
        for (int n = 0; n < 2; ++n) { s[kt][n][0] = s[kt][n][1] = s[kt][n][2] = s[kt][n][3] = 0.f; }
        const int key = kt * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0, b1, b2, b3;
          if constexpr (PRECISE) {
            ldmatrix_x4(sKl + kv_off(key, 2 * ks + (mi & 1)), b0, b1, b2, b3);      // small terms first
            mma_bf16_16816(s[kt][0], qh[ks], b0, b1);
            mma_bf16_16816(s[kt][1], qh[ks], b2, b3);
          }
          ldmatrix_x4(sKh + kv_off(key, 2 * ks + (mi & 1)), b0, b1, b2, b3);
          if constexpr (PRECISE) {
            mma_bf16_16816(s[kt][0], ql[ks], b0, b1);
            mma_bf16_16816(s[kt][1], ql[ks], b2, b3);
          }
          mma_bf16_16816(s[kt][0], qh[ks], b0, b1);
          mma_bf16_16816(s[kt][1], qh[ks], b2, b3);
        }
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int kc = kt * 16 + 8 * n + 2 * tg;
          s[kt][n][0] = key_on<CAUSAL>(kc, r0, T) ? s[kt][n][0] * scale_log2 : -1e30f;
          s[kt][n][1] = key_on<CAUSAL>(kc + 1, r0, T) ? s[kt][n][1] * scale_log2 : -1e30f;
          s[kt][n][2] = key_on<CAUSAL>(kc, r1, T) ? s[kt][n][2] * scale_log2 : -1e30f;
          s[kt][n][3] = key_on<CAUSAL>(kc + 1, r1, T) ? s[kt][n][3] * scale_log2 : -1e30f;
          m0 = fmaxf(m0, fmaxf(s[kt][n][0], s[kt][n][1]));
          m1 = fmaxf(m1, fmaxf(s[kt][n][2], s[kt][n][3]));
        }
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < SAM_MAX_TILES; ++kt) {
      if (kt <= kt_end) {
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const float p0 = exp2f(s[kt][n][0] - m0), p1 = exp2f(s[kt][n][1] - m0);
          const float p2 = exp2f(s[kt][n][2] - m1), p3 = exp2f(s[kt][n][3] - m1);
          l0 += p0 + p1; l1 += p2 + p3;
          ph[2 * n] = pack_bf16x2(p0, p1);
          ph[2 * n + 1] = pack_bf16x2(p2, p3);
          if constexpr (PRECISE) {
            pl[2 * n] = pack_bf16x2(p0 - bf16_round(p0), p1 - bf16_round(p1));
            pl[2 * n + 1] = pack_bf16x2(p2 - bf16_round(p2), p3 - bf16_round(p3));
          }
        }
        const int krow = kt * 16 + (mi & 1) * 8 + (lane & 7);
#pragma unroll
        for (int nd2 = 0; nd2 < 4; ++nd2) {
          uint32_t b0, b1, b2, b3;
          if constexpr (PRECISE) {
            ldmatrix_x4_trans(sVl + kv_off(krow, 2 * nd2 + (mi >> 1)), b0, b1, b2, b3);
            mma_bf16_16816(o[2 * nd2], ph, b0, b1);
            mma_bf16_16816(o[2 * nd2 + 1], ph, b2, b3);
          }
          ldmatrix_x4_trans(sVh + kv_off(krow, 2 * nd2 + (mi >> 1)), b0, b1, b2, b3);
          if constexpr (PRECISE) {
            mma_bf16_16816(o[2 * nd2], pl, b0, b1);
            mma_bf16_16816(o[2 * nd2 + 1], pl, b2, b3);
          }
          mma_bf16_16816(o[2 * nd2], ph, b0, b1);
          mma_bf16_16816(o[2 * nd2 + 1], ph, b2, b3);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const long long e0 = ((long long)b * T + r0) * D + h * 64 + 2 * tg, e1 = e0 + 8LL * D;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if constexpr (PRECISE) {
        float* out = reinterpret_cast<float*>(out_);
        if (r0 < T) *reinterpret_cast<float2*>(out + e0 + 8 * n) = make_float2(o[n][0] * i0, o[n][1] * i0);
        if (r1 < T) *reinterpret_cast<float2*>(out + e1 + 8 * n) = make_float2(o[n][2] * i1, o[n][3] * i1);
      } else {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_);
        if (r0 < T) *reinterpret_cast<uint32_t*>(out + e0 + 8 * n) = pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
        if (r1 < T) *reinterpret_cast<uint32_t*>(out + e1 + 8 * n) = pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
      }
    }
  }
}

// ============================================================================================== backward
template <bool IN_F32, bool CAUSAL>
__global__ void __launch_bounds__(SAM_THREADS, 1)
seq_attention_bwd_mma_kernel(const void* __restrict__ qkv_, const __nv_bfloat16* __restrict__ dout,
                             __nv_bfloat16* __restrict__ dqkv, int T, int H, int t_pad) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * 64;
  const uint32_t plane = (uint32_t)t_pad * 128u;
  const uint32_t sQ = smem_u32(smem), sK = sQ + plane, sV = sK + plane, sdO = sV + plane;
  float* s_lse = reinterpret_cast<float*>(smem + (size_t)4 * plane);
  float* s_delta = s_lse + t_pad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tg = lane & 3, mi = lane >> 3;
  const long long ld = 3LL * D;

  for (int i = tid; i < t_pad * 8 * 4; i += SAM_THREADS) {       // (row, chunk, which of q/k/v/dO)
    const int which = i / (t_pad * 8), j = i - which * (t_pad * 8);
    const int row = j >> 3, ch = j & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < T) {
      if (which < 3) {
        const long long e = ((long long)b * T + row) * ld + which * D + h * 64 + ch * 8;
        if constexpr (IN_F32) {
          const float* src = reinterpret_cast<const float*>(qkv_) + e;
          v = pack8_bf16(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + 4));
        } else {
          v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(qkv_) + e);
        }
      } else {
        v = *reinterpret_cast<const uint4*>(dout + ((long long)b * T + row) * D + h * 64 + ch * 8);
      }
    }
    *reinterpret_cast<uint4*>(smem + kv_off(row, ch) + (uint32_t)which * plane) = v;
  }
  __syncthreads();

  const float scale = 0.125f;
  const float scale_log2 = 0.125f * 1.4426950408889634f;
  const int n_t = t_pad >> 4;

  // ---- pre-pass: lse (log2 domain) and delta of every query row
  for (int mt = n_t - 1 - warp; mt >= 0; mt -= SAM_WARPS) {
    uint32_t qf[4][4], dof[4][4];
    {
      const int row = mt * 16 + (lane & 15);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        ldmatrix_x4(sQ + kv_off(row, 2 * ks + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        ldmatrix_x4(sdO + kv_off(row, 2 * ks + (lane >> 4)), dof[ks][0], dof[ks][1], dof[ks][2], dof[ks][3]);
      }
    }
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const int kt_end = CAUSAL ? mt : n_t - 1;
    float s[SAM_MAX_TILES][2][4];
    float m0 = -1e30f, m1 = -1e30f;
#pragma unroll
    for (int kt = 0; kt < SAM_MAX_TILES; ++kt) {
      if (kt <= kt_end) {
#pragma unroll
        for (int n = 0; n < 2; ++n) { s[kt][n][0] = s[kt][n][1] = s[kt][n][2] = s[kt][n][3] = 0.f; }
        const int key = kt * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(sK + kv_off(key, 2 * ks + (mi & 1)), b0, b1, b2, b3);
          mma_bf16_16816(s[kt][0], qf[ks], b0, b1);
          mma_bf16_16816(s[kt][1], qf[ks], b2, b3);
        }
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int kc = kt * 16 + 8 * n + 2 * tg;
          s[kt][n][0] = key_on<CAUSAL>(kc, r0, T) ? s[kt][n][0] * scale_log2 : -1e30f;
          s[kt][n][1] = key_on<CAUSAL>(kc + 1, r0, T) ? s[kt][n][1] * scale_log2 : -1e30f;
          s[kt][n][2] = key_on<CAUSAL>(kc, r1, T) ? s[kt][n][2] * scale_log2 : -1e30f;
          s[kt][n][3] = key_on<CAUSAL>(kc + 1, r1, T) ? s[kt][n][3] * scale_log2 : -1e30f;
          m0 = fmaxf(m0, fmaxf(s[kt][n][0], s[kt][n][1]));
          m1 = fmaxf(m1, fmaxf(s[kt][n][2], s[kt][n][3]));
        }
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int kt = 0; kt < SAM_MAX_TILES; ++kt) {
      if (kt <= kt_end) {
        float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const int key = kt * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(sV + kv_off(key, 2 * ks + (mi & 1)), b0, b1, b2, b3);
          mma_bf16_16816(dp[0], dof[ks], b0, b1);
          mma_bf16_16816(dp[1], dof[ks], b2, b3);
        }
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const float p0 = exp2f(s[kt][n][0] - m0), p1 = exp2f(s[kt][n][1] - m0);
          const float p2 = exp2f(s[kt][n][2] - m1), p3 = exp2f(s[kt][n][3] - m1);
          l0 += p0 + p1; l1 += p2 + p3;
          d0 += p0 * dp[n][0] + p1 * dp[n][1];
          d1 += p2 * dp[n][2] + p3 * dp[n][3];
        }
      }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, o); l1 += __shfl_xor_sync(0xffffffffu, l1, o);
      d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    }
    if (tg == 0) {
      s_lse[r0] = m0 + log2f(l0); s_lse[r1] = m1 + log2f(l1);
      s_delta[r0] = d0 / l0; s_delta[r1] = d1 / l1;
    }
  }
  __syncthreads();

  // ---- work units, costliest first: even u = key tile u/2 (dK, dV over the query tiles that see it), odd u = query tile
  // n_t-1-u/2 (dQ over the key tiles it sees)
  for (int u = warp; u < 2 * n_t; u += SAM_WARPS) {
    if (u & 1) {
      const int mt = n_t - 1 - (u >> 1);
      uint32_t qf[4][4], dof[4][4];
      {
        const int row = mt * 16 + (lane & 15);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          ldmatrix_x4(sQ + kv_off(row, 2 * ks + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
          ldmatrix_x4(sdO + kv_off(row, 2 * ks + (lane >> 4)), dof[ks][0], dof[ks][1], dof[ks][2], dof[ks][3]);
        }
      }
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      const float lse0 = s_lse[r0], lse1 = s_lse[r1];
      const float dl0 = s_delta[r0], dl1 = s_delta[r1];
      float dq[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
      const int kt_end = CAUSAL ? mt : n_t - 1;
      for (int kt = 0; kt <= kt_end; ++kt) {
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const int key = kt * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(sK + kv_off(key, 2 * ks + (mi & 1)), b0, b1, b2, b3);
          mma_bf16_16816(s[0], qf[ks], b0, b1);
          mma_bf16_16816(s[1], qf[ks], b2, b3);
          ldmatrix_x4(sV + kv_off(key, 2 * ks + (mi & 1)), b0, b1, b2, b3);
          mma_bf16_16816(dp[0], dof[ks], b0, b1);
          mma_bf16_16816(dp[1], dof[ks], b2, b3);
        }
        uint32_t a[4];
        {
          float ds[2][4];
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            const int kc = kt * 16 + 8 * n + 2 * tg;
            ds[n][0] = key_on<CAUSAL>(kc, r0, T) ? exp2f(s[n][0] * scale_log2 - lse0) * (dp[n][0] - dl0) : 0.f;
            ds[n][1] = key_on<CAUSAL>(kc + 1, r0, T) ? exp2f(s[n][1] * scale_log2 - lse0) * (dp[n][1] - dl0) : 0.f;
            ds[n][2] = key_on<CAUSAL>(kc, r1, T) ? exp2f(s[n][2] * scale_log2 - lse1) * (dp[n][2] - dl1) : 0.f;
            ds[n][3] = key_on<CAUSAL>(kc + 1, r1, T) ? exp2f(s[n][3] * scale_log2 - lse1) * (dp[n][3] - dl1) : 0.f;
          }
          a[0] = pack_bf16x2(ds[0][0], ds[0][1]);
          a[1] = pack_bf16x2(ds[0][2], ds[0][3]);
          a[2] = pack_bf16x2(ds[1][0], ds[1][1]);
          a[3] = pack_bf16x2(ds[1][2], ds[1][3]);
        }
        const int krow = kt * 16 + (mi & 1) * 8 + (lane & 7);
#pragma unroll
        for (int nd2 = 0; nd2 < 4; ++nd2) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(sK + kv_off(krow, 2 * nd2 + (mi >> 1)), b0, b1, b2, b3);
          mma_bf16_16816(dq[2 * nd2], a, b0, b1);
          mma_bf16_16816(dq[2 * nd2 + 1], a, b2, b3);
        }
      }
      __nv_bfloat16* o0 = dqkv + ((long long)b * T + r0) * ld + h * 64 + 2 * tg;
      __nv_bfloat16* o1 = o0 + 8 * ld;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if (r0 < T) *reinterpret_cast<uint32_t*>(o0 + 8 * n) = pack_bf16x2(dq[n][0] * scale, dq[n][1] * scale);
        if (r1 < T) *reinterpret_cast<uint32_t*>(o1 + 8 * n) = pack_bf16x2(dq[n][2] * scale, dq[n][3] * scale);
      }
    } else {
      const int kt = u >> 1;
      uint32_t kf[4][4], vf[4][4];
      {
        const int krow_a = kt * 16 + (lane & 15);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          ldmatrix_x4(sK + kv_off(krow_a, 2 * ks + (lane >> 4)), kf[ks][0], kf[ks][1], kf[ks][2], kf[ks][3]);
          ldmatrix_x4(sV + kv_off(krow_a, 2 * ks + (lane >> 4)), vf[ks][0], vf[ks][1], vf[ks][2], vf[ks][3]);
        }
      }
      const int kr0 = kt * 16 + g, kr1 = kr0 + 8;
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) { dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f; }
      for (int qt = CAUSAL ? kt : 0; qt < n_t; ++qt) {
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const int qrow = qt * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(sQ + kv_off(qrow, 2 * ks + (mi & 1)), b0, b1, b2, b3);
          mma_bf16_16816(s[0], kf[ks], b0, b1);
          mma_bf16_16816(s[1], kf[ks], b2, b3);
          ldmatrix_x4(sdO + kv_off(qrow, 2 * ks + (mi & 1)), b0, b1, b2, b3);
          mma_bf16_16816(dp[0], vf[ks], b0, b1);
          mma_bf16_16816(dp[1], vf[ks], b2, b3);
        }
        uint32_t pa[4], da[4];
        {
          float pt[2][4], ds[2][4];
#pragma unroll
          for (int n = 0; n < 2; ++n) {   // rows are keys kr0 / kr1, columns are queries qc, qc + 1
            const int qc = qt * 16 + 8 * n + 2 * tg;
            const float2 ls = *reinterpret_cast<const float2*>(&s_lse[qc]);
            const float2 dl = *reinterpret_cast<const float2*>(&s_delta[qc]);
            pt[n][0] = key_on<CAUSAL>(kr0, qc, T) ? exp2f(s[n][0] * scale_log2 - ls.x) : 0.f;
            pt[n][1] = key_on<CAUSAL>(kr0, qc + 1, T) ? exp2f(s[n][1] * scale_log2 - ls.y) : 0.f;
            pt[n][2] = key_on<CAUSAL>(kr1, qc, T) ? exp2f(s[n][2] * scale_log2 - ls.x) : 0.f;
            pt[n][3] = key_on<CAUSAL>(kr1, qc + 1, T) ? exp2f(s[n][3] * scale_log2 - ls.y) : 0.f;
            ds[n][0] = pt[n][0] * (dp[n][0] - dl.x); ds[n][1] = pt[n][1] * (dp[n][1] - dl.y);
            ds[n][2] = pt[n][2] * (dp[n][2] - dl.x); ds[n][3] = pt[n][3] * (dp[n][3] - dl.y);
          }
          pa[0] = pack_bf16x2(pt[0][0], pt[0][1]); pa[1] = pack_bf16x2(pt[0][2], pt[0][3]);
          pa[2] = pack_bf16x2(pt[1][0], pt[1][1]); pa[3] = pack_bf16x2(pt[1][2], pt[1][3]);
          da[0] = pack_bf16x2(ds[0][0], ds[0][1]); da[1] = pack_bf16x2(ds[0][2], ds[0][3]);
          da[2] = pack_bf16x2(ds[1][0], ds[1][1]); da[3] = pack_bf16x2(ds[1][2], ds[1][3]);
        }
        const int qr = qt * 16 + (mi & 1) * 8 + (lane & 7);
#pragma unroll
        for (int nd2 = 0; nd2 < 4; ++nd2) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(sdO + kv_off(qr, 2 * nd2 + (mi >> 1)), b0, b1, b2, b3);
          mma_bf16_16816(dv[2 * nd2], pa, b0, b1);
          mma_bf16_16816(dv[2 * nd2 + 1], pa, b2, b3);
          ldmatrix_x4_trans(sQ + kv_off(qr, 2 * nd2 + (mi >> 1)), b0, b1, b2, b3);
          mma_bf16_16816(dk[2 * nd2], da, b0, b1);
          mma_bf16_16816(dk[2 * nd2 + 1], da, b2, b3);
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int kr = kr0 + 8 * half;
        if (kr < T) {
          __nv_bfloat16* ok = dqkv + ((long long)b * T + kr) * ld + D + h * 64 + 2 * tg;
          __nv_bfloat16* ov = ok + D;
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            *reinterpret_cast<uint32_t*>(ok + 8 * n) = pack_bf16x2(dk[n][2 * half] * scale, dk[n][2 * half + 1] * scale);
            *reinterpret_cast<uint32_t*>(ov + 8 * n) = pack_bf16x2(dv[n][2 * half], dv[n][2 * half + 1]);
          }
        }
      }
    }
  }
}

// ============================================================================================== host side
template <typename K>
static int opt_in_smem(K kern, size_t bytes, size_t* granted_dev, const char* what) {
  size_t& granted = granted_dev[device_slot()];
  if (granted == 0) granted = 48 * 1024;
  if (bytes > granted) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(%s): %s", what, cudaGetErrorString(e)); return (int)e; }
    granted = bytes;
  }
  return GAVA_OK;
}

bool causal_attention_mma_supported(int L, int H) { return L > 0 && L <= SAM_MAX_T && H > 0; }

template <bool PRECISE, bool CAUSAL>
static int launch_seq_fwd(const void* qkv, void* out, int n_seq, int L, int H, cudaStream_t stream, const char* what) {
  if (!qkv || !out || n_seq <= 0 || !causal_attention_mma_supported(L, H)) {
    set_error("%s: bad arguments (n_seq=%d L=%d H=%d; L <= %d)", what, n_seq, L, H, SAM_MAX_T);
    return GAVA_ERR_ARG;
  }
  const int t_pad = (L + 15) / 16 * 16;
  const size_t smem = (size_t)(PRECISE ? 6 : 3) * t_pad * 128;
  static size_t granted[64] = {0};
  int rc = opt_in_smem(seq_attention_mma_kernel<PRECISE, CAUSAL>, smem, granted, what);
  if (rc) return rc;
  seq_attention_mma_kernel<PRECISE, CAUSAL><<<dim3(H, n_seq), SAM_THREADS, smem, stream>>>(qkv, out, L, H, t_pad);
  return check_launch(what);
}

template <bool IN_F32, bool CAUSAL>
static int launch_seq_bwd(const void* qkv, const void* dout, void* dqkv, int n_seq, int L, int H, cudaStream_t stream,
                          const char* what) {
  if (!qkv || !dout || !dqkv || n_seq <= 0 || !causal_attention_mma_supported(L, H)) {
    set_error("%s: bad arguments (n_seq=%d L=%d H=%d; L <= %d)", what, n_seq, L, H, SAM_MAX_T);
    return GAVA_ERR_ARG;
  }
  const int t_pad = (L + 15) / 16 * 16;
  const size_t smem = (size_t)4 * t_pad * 128 + (size_t)2 * t_pad * sizeof(float);
  static size_t granted[64] = {0};
  int rc = opt_in_smem(seq_attention_bwd_mma_kernel<IN_F32, CAUSAL>, smem, granted, what);
  if (rc) return rc;
  seq_attention_bwd_mma_kernel<IN_F32, CAUSAL><<<dim3(H, n_seq), SAM_THREADS, smem, stream>>>(
      qkv, reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<__nv_bfloat16*>(dqkv), L, H, t_pad);
  return check_launch(what);
}

// text transformer: f32 q|k|v, causal
int causal_attention_mma(const float* qkv, float* out, int n_seq, int L, int H, cudaStream_t stream) {
  return launch_seq_fwd<true, true>(qkv, out, n_seq, L, H, stream, "causal_attention_mma");
}
int causal_attention_bwd_mma(const float* qkv, const void* dout, void* dqkv, int n_seq, int L, int H, cudaStream_t stream) {
  return launch_seq_bwd<true, true>(qkv, dout, dqkv, n_seq, L, H, stream, "causal_attention_bwd_mma");
}
// summary-token attention over the T frames of each clip: bf16 q|k|v, every frame sees every frame
int temporal_attention_mma(const void* qkv, void* out, int B, int T, int H, cudaStream_t stream) {
  return launch_seq_fwd<false, false>(qkv, out, B, T, H, stream, "temporal_attention_mma");
}
int temporal_attention_bwd_mma(const void* qkv, const void* dout, void* dqkv, int B, int T, int H, cudaStream_t stream) {
  return launch_seq_bwd<false, false>(qkv, dout, dqkv, B, T, H, stream, "temporal_attention_bwd_mma");
}

}  // namespace gava
