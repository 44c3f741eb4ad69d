// Backward of the prompt-prefixed multi-head self-attention core (attention.cu / attention_tc.cu), flash-style: the score
// matrix is never written to HBM; it is recomputed on the SM from the saved q/k/v.
//
// One CTA per (frame, head).  Q, dO (the first n_q live query rows of the frame: all N0, or 1 in the last block) and K, V
// (N0 live rows + G global + T local + 1 summary row, gathered exactly as in the forward) are staged once in swizzled
// shared memory (cp.async).  Then
//   prologue : delta[q] = <dO[q], O[q]> (O read once from global); lse[q] from the forward's saved rows, or -- without
//              them -- an extra pass over the keys (online max/sum per 16-query tile) followed by a CTA barrier
//   work units, dealt round-robin to the 4 warps (independent of each other once lse / delta sit in shared memory):
//     16-query tile : per 16-key tile  S = QK^T, P = exp2(S - lse), dP = dO V^T, dS = P (dP - delta), dQ += dS K  -> dq (bf16)
//     16-key tile   : per 16-query tile  S^T = KQ^T, P^T, dV += P^T dO, dP^T = V dO^T, dS^T, dK += dS^T Q
//                     live key rows -> dk, dv (bf16, plain stores); K/V-only rows (global / local / summary prompts,
//                     shared by several frames) -> this CTA's slot of `scratch`, summed in a fixed order by
//                     attention_bwd_reduce_kernel (reproducible), or f32 atomics into d_kv_aux when scratch == NULL.
// All seven contractions (S and dP are formed twice, once per unit type) run on the tensor cores (mma.sync m16n8k16 bf16, f32 accumulate).
// Padding rows are zero-filled, which makes their contributions vanish without masks: K_pad = 0 => dS K_pad = 0,
// dO_pad = 0 => P^T dO_pad = 0 and dP = delta = 0 => dS = 0.  Only the lse pass masks padded keys.
// Measured (config 2, 6144 items of 197 x 214): 0.96 ms per launch, HMMA pipe 42 % busy; 8 warps with <= 128 registers
// (GAVA_AB_WARPS=8) is slower (1.08 ms), see DESIGN.md.
//
// Restates autograd of Attention.forward core, training/VitaCLIP_vision_encoder_utils.py:71-77.
#include "common.cuh"
#include "kernels.h"
#include "mma_sync.cuh"

namespace gava {

struct AttnBwdArgs {
  const __nv_bfloat16* qkv;     // [BT*N0, 3D]
  const __nv_bfloat16* kv_aux;  // [G+2BT, 2D]
  const __nv_bfloat16* out;     // [BT*N0, D]   forward output O
  const __nv_bfloat16* dout;    // [BT*N0, D]
  __nv_bfloat16* dqkv;          // [BT*N0, 3D]
  float* dkv_aux;               // [G+2BT, 2D]  atomically accumulated (zeroed by the launcher)
  const float* lse;             // [BT, H, N0] log2-domain log-sum-exp rows saved by the forward, or null (recomputed)
  float* scratch;               // optional [BT, H, G+T+1, 128]: per-(frame, head) prompt-row gradients, reduced by a 2nd kernel
  int T, N0, G, H, S, BT, q_pad, s_pad;
  int n_q;   // query rows per frame that carry a gradient (N0, or 1 in the last layer: only the cls row reaches the output)
  int groups;  // bit 0: local-prompt keys took part in the forward, bit 1: the summary-token key did (3 = full Vita-CLIP)
};

// key row `key` belongs to a prompt group the forward masked out (zero probability): its P, dS, dK, dV are exactly zero
__device__ __forceinline__ bool key_disabled(const AttnBwdArgs& p, int key) {
  const int a = key - (p.N0 + p.G);
  return a >= 0 && key < p.S && ((a < p.T) ? !(p.groups & 1) : !(p.groups & 2));
}

// Launch shapes: <4 warps, 2 CTAs per SM> (255 registers per thread, K/V A-fragments of a key-tile unit live in
// registers) while two CTAs' shared memory fits an SM; items with more keys (the 70-frame recipe: 276 keys, ViT-L/14 at
// 16 frames: 282) only fit once per SM and run <8 warps, 1 CTA> with the same register budget.  <8 warps, 2 CTAs> (<= 128
// registers, key-tile units re-read their K/V fragments every trip) measured slower (1.08 vs 0.96 ms at config 2) and is
// kept for experiments only (GAVA_AB_WARPS=8).
#ifndef GAVA_AB_WARPS
#define GAVA_AB_WARPS 4
#endif

// MASKED: partial prompt-flag configurations (use_summary_token / use_local_prompts off): a separate instantiation, so the
// full configuration's register allocation (209 of 255) is untouched.
template <bool HAVE_LSE, int WARPS, int MIN_CTAS, bool MASKED>
__global__ void __launch_bounds__(32 * WARPS, MIN_CTAS) attention_bwd_kernel(const AttnBwdArgs p) {
  constexpr int AB_WARPS = WARPS;
  constexpr bool AB_RELOAD_KV = WARPS * MIN_CTAS > 8;   // <= 128 registers per thread: no room for resident K/V fragments
  constexpr int AB_UNROLL = AB_RELOAD_KV ? 1 : 2;       // key / query tiles per loop trip: independent MMA chains to overlap
  extern __shared__ __align__(128) uint8_t smem[];
  const int h = blockIdx.x, f = blockIdx.y;
  const int D = p.H * 64;
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sdO = sQ + (uint32_t)p.q_pad * 128u;
  const uint32_t sK = sdO + (uint32_t)p.q_pad * 128u;
  const uint32_t sV = sK + (uint32_t)p.s_pad * 128u;
  float* s_lse = reinterpret_cast<float*>(smem + (size_t)(2 * p.q_pad + 2 * p.s_pad) * 128);
  float* s_delta = s_lse + p.q_pad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tg = lane & 3;
  const long long ldq = 3LL * D, lda = 2LL * D;
  const int b = f / p.T;
  constexpr int NT = 32 * AB_WARPS;

  // ---- stage Q, dO
  for (int i = tid; i < p.q_pad * 8; i += NT) {
    const int row = i >> 3, ch = i & 7;
    const uint32_t off = kv_off(row, ch);
    if (row < p.n_q) {
      cp_async16(sQ + off, p.qkv + ((long long)f * p.N0 + row) * ldq + h * 64 + ch * 8);
      cp_async16(sdO + off, p.dout + ((long long)f * p.N0 + row) * D + h * 64 + ch * 8);
    } else {
      *reinterpret_cast<uint4*>(smem + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(smem + (size_t)p.q_pad * 128 + off) = make_uint4(0, 0, 0, 0);
    }
  }
  // ---- stage K, V (same gather as the forward)
  for (int i = tid; i < p.s_pad * 8; i += NT) {
    const int row = i >> 3, ch = i & 7;
    const __nv_bfloat16* ksrc = nullptr;
    if (row < p.N0) {
      ksrc = p.qkv + ((long long)f * p.N0 + row) * ldq + D + h * 64;
    } else if (row < p.S) {
      const int a = row - p.N0;
      int arow;
      if (a < p.G) arow = a;
      else if (a < p.G + p.T) arow = p.G + b * p.T + (a - p.G);
      else arow = p.G + p.BT + f;
      ksrc = p.kv_aux + (long long)arow * lda + h * 64;
    }
    const uint32_t off = kv_off(row, ch);
    if (ksrc != nullptr) {
      cp_async16(sK + off, ksrc + ch * 8);
      cp_async16(sV + off, ksrc + D + ch * 8);
    } else {
      *reinterpret_cast<uint4*>(smem + (size_t)2 * p.q_pad * 128 + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(smem + (size_t)(2 * p.q_pad + p.s_pad) * 128 + off) = make_uint4(0, 0, 0, 0);
    }
  }
  // ---- phase 0: delta[q] = sum_d dO[q,d] O[q,d]  (8 lanes per row, 16 bytes each, straight from global)
  for (int i = tid; i < p.q_pad * 8; i += NT) {
    const int row = i >> 3, ch = i & 7;
    float acc = 0.f;
    if (row < p.n_q) {
      const long long off = ((long long)f * p.N0 + row) * D + h * 64 + ch * 8;
      const uint4 o = *reinterpret_cast<const uint4*>(p.out + off);
      const uint4 d = *reinterpret_cast<const uint4*>(p.dout + off);
      const uint32_t* ow = reinterpret_cast<const uint32_t*>(&o);
      const uint32_t* dw = reinterpret_cast<const uint32_t*>(&d);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&ow[k]);
        const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(&dw[k]);
        acc += __bfloat162float(a.x) * __bfloat162float(c.x) + __bfloat162float(a.y) * __bfloat162float(c.y);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (ch == 0) s_delta[row] = acc;
  }
  if constexpr (HAVE_LSE) {
    const float* src = p.lse + ((long long)f * p.H + h) * p.N0;
    for (int i = tid; i < p.q_pad; i += NT) s_lse[i] = (i < p.n_q) ? src[i] : 0.f;   // padded rows: dO = 0 kills them
  }
  cp_async_wait_all();
  __syncthreads();

  const float scale = 0.125f;                                 // dh^-0.5, dh = 64
  const float scale_log2 = 0.125f * 1.4426950408889634f;
  const int n_qt = p.q_pad >> 4, n_kt = p.s_pad >> 4;
  const int mi = lane >> 3;

  // =========================================================== lse pass (only when the forward did not save it)
  if constexpr (!HAVE_LSE) {
  for (int mt = warp; mt < n_qt; mt += AB_WARPS) {
    uint32_t qf[4][4];
    {
      const int row = mt * 16 + (lane & 15);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldmatrix_x4(sQ + kv_off(row, 2 * ks + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    // ---- lse (log2 domain) of rows g and g+8
    float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;
    for (int kt = 0; kt < n_kt; ++kt) {
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      const int key = kt * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(sK + kv_off(key, 2 * ks + (mi & 1)), b0, b1, b2, b3);
        mma_bf16_16816(s[0], qf[ks], b0, b1);
        mma_bf16_16816(s[1], qf[ks], b2, b3);
      }
      float mx0 = m0, mx1 = m1;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int kc = kt * 16 + 8 * n + 2 * tg;
        const bool on0 = kc < p.S && !(MASKED && key_disabled(p, kc)), on1 = kc + 1 < p.S && !(MASKED && key_disabled(p, kc + 1));
        s[n][0] = on0 ? s[n][0] * scale_log2 : -1e30f;
        s[n][1] = on1 ? s[n][1] * scale_log2 : -1e30f;
        s[n][2] = on0 ? s[n][2] * scale_log2 : -1e30f;
        s[n][3] = on1 ? s[n][3] * scale_log2 : -1e30f;
        mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
      }
      l0 = l0 * exp2f(m0 - mx0) + exp2f(s[0][0] - mx0) + exp2f(s[0][1] - mx0) + exp2f(s[1][0] - mx0) + exp2f(s[1][1] - mx0);
      l1 = l1 * exp2f(m1 - mx1) + exp2f(s[0][2] - mx1) + exp2f(s[0][3] - mx1) + exp2f(s[1][2] - mx1) + exp2f(s[1][3] - mx1);
      m0 = mx0; m1 = mx1;
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {   // combine the 4 lanes of a quad
      const float om0 = __shfl_xor_sync(0xffffffffu, m0, o), ol0 = __shfl_xor_sync(0xffffffffu, l0, o);
      const float om1 = __shfl_xor_sync(0xffffffffu, m1, o), ol1 = __shfl_xor_sync(0xffffffffu, l1, o);
      const float n0 = fmaxf(m0, om0), n1 = fmaxf(m1, om1);
      l0 = l0 * exp2f(m0 - n0) + ol0 * exp2f(om0 - n0);
      l1 = l1 * exp2f(m1 - n1) + ol1 * exp2f(om1 - n1);
      m0 = n0; m1 = n1;
    }
    const float lse0 = m0 + log2f(l0), lse1 = m1 + log2f(l1);
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    if (tg == 0) { s_lse[r0] = m0 + log2f(l0); s_lse[r1] = m1 + log2f(l1); }
  }
  __syncthreads();   // s_lse of every query tile is complete
  }

  // =========================================================== work units: 16-query tiles (dQ) and 16-key tiles (dK, dV),
  // independent of each other once lse / delta are in shared memory: dealt round-robin to the warps
  for (int u = warp; u < n_qt + n_kt; u += AB_WARPS) {
  if (u < n_qt) {
    const int mt = u;
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
#pragma unroll AB_UNROLL
    for (int kt = 0; kt < n_kt; ++kt) {
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
          ds[n][0] = exp2f(s[n][0] * scale_log2 - lse0) * (dp[n][0] - dl0);
          ds[n][1] = exp2f(s[n][1] * scale_log2 - lse0) * (dp[n][1] - dl0);
          ds[n][2] = exp2f(s[n][2] * scale_log2 - lse1) * (dp[n][2] - dl1);
          ds[n][3] = exp2f(s[n][3] * scale_log2 - lse1) * (dp[n][3] - dl1);
          if constexpr (MASKED) {
            const int kc = kt * 16 + 8 * n + 2 * tg;
            if (key_disabled(p, kc)) { ds[n][0] = 0.f; ds[n][2] = 0.f; }
            if (key_disabled(p, kc + 1)) { ds[n][1] = 0.f; ds[n][3] = 0.f; }
          }
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
    __nv_bfloat16* o0 = p.dqkv + ((long long)f * p.N0 + r0) * ldq + h * 64 + 2 * tg;
    __nv_bfloat16* o1 = o0 + 8 * ldq;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (r0 < p.n_q) *reinterpret_cast<uint32_t*>(o0 + 8 * n) = pack_bf16x2(dq[n][0] * scale, dq[n][1] * scale);
      if (r1 < p.n_q) *reinterpret_cast<uint32_t*>(o1 + 8 * n) = pack_bf16x2(dq[n][2] * scale, dq[n][3] * scale);
    }
  } else {
    const int kt = u - n_qt;
    uint32_t kf[AB_RELOAD_KV ? 1 : 4][4], vf[AB_RELOAD_KV ? 1 : 4][4];
    const int krow_a = kt * 16 + (lane & 15);
    if constexpr (!AB_RELOAD_KV) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        ldmatrix_x4(sK + kv_off(krow_a, 2 * ks + (lane >> 4)), kf[ks][0], kf[ks][1], kf[ks][2], kf[ks][3]);
        ldmatrix_x4(sV + kv_off(krow_a, 2 * ks + (lane >> 4)), vf[ks][0], vf[ks][1], vf[ks][2], vf[ks][3]);
      }
    }
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f; }
#pragma unroll AB_UNROLL
    for (int qt = 0; qt < n_qt; ++qt) {
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      const int qrow = qt * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b0, b1, b2, b3;
        constexpr int fi = AB_RELOAD_KV ? 0 : 1;
        if constexpr (AB_RELOAD_KV) {
          ldmatrix_x4(sK + kv_off(krow_a, 2 * ks + (lane >> 4)), kf[0][0], kf[0][1], kf[0][2], kf[0][3]);
          ldmatrix_x4(sV + kv_off(krow_a, 2 * ks + (lane >> 4)), vf[0][0], vf[0][1], vf[0][2], vf[0][3]);
        }
        ldmatrix_x4(sQ + kv_off(qrow, 2 * ks + (mi & 1)), b0, b1, b2, b3);
        mma_bf16_16816(s[0], kf[fi * ks], b0, b1);
        mma_bf16_16816(s[1], kf[fi * ks], b2, b3);
        ldmatrix_x4(sdO + kv_off(qrow, 2 * ks + (mi & 1)), b0, b1, b2, b3);
        mma_bf16_16816(dp[0], vf[fi * ks], b0, b1);
        mma_bf16_16816(dp[1], vf[fi * ks], b2, b3);
      }
      uint32_t pa[4], da[4];
      {
        float pt[2][4], ds[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {   // columns are queries: qt*16 + 8n + 2tg (+1)
          const int qc = qt * 16 + 8 * n + 2 * tg;
          const float2 ls = *reinterpret_cast<const float2*>(&s_lse[qc]);
          const float2 dl = *reinterpret_cast<const float2*>(&s_delta[qc]);
          pt[n][0] = exp2f(s[n][0] * scale_log2 - ls.x); pt[n][1] = exp2f(s[n][1] * scale_log2 - ls.y);
          pt[n][2] = exp2f(s[n][2] * scale_log2 - ls.x); pt[n][3] = exp2f(s[n][3] * scale_log2 - ls.y);
          if constexpr (MASKED) {     // rows are keys kt*16 + g (+8)
            if (key_disabled(p, kt * 16 + g)) { pt[n][0] = 0.f; pt[n][1] = 0.f; }
            if (key_disabled(p, kt * 16 + g + 8)) { pt[n][2] = 0.f; pt[n][3] = 0.f; }
          }
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
    // ---- write-out: live rows -> dqkv (bf16), K/V-only rows -> atomics into dkv_aux (f32)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int kr = kt * 16 + g + 8 * half;
      if (kr < p.N0) {
        __nv_bfloat16* ok = p.dqkv + ((long long)f * p.N0 + kr) * ldq + D + h * 64 + 2 * tg;
        __nv_bfloat16* ov = ok + D;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          *reinterpret_cast<uint32_t*>(ok + 8 * n) = pack_bf16x2(dk[n][2 * half] * scale, dk[n][2 * half + 1] * scale);
          *reinterpret_cast<uint32_t*>(ov + 8 * n) = pack_bf16x2(dv[n][2 * half], dv[n][2 * half + 1]);
        }
      } else if (kr < p.S && p.scratch != nullptr) {
        // deterministic path: this CTA's contribution goes to its own slot, attention_bwd_reduce_kernel sums the slots
        float* ok = p.scratch + (((long long)f * p.H + h) * (p.S - p.N0) + (kr - p.N0)) * 128 + 2 * tg;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          *reinterpret_cast<float2*>(ok + 8 * n) = make_float2(dk[n][2 * half] * scale, dk[n][2 * half + 1] * scale);
          *reinterpret_cast<float2*>(ok + 64 + 8 * n) = make_float2(dv[n][2 * half], dv[n][2 * half + 1]);
        }
      } else if (kr < p.S) {
        const int a = kr - p.N0;
        int arow;
        if (a < p.G) arow = a;
        else if (a < p.G + p.T) arow = p.G + b * p.T + (a - p.G);
        else arow = p.G + p.BT + f;
        float* ok = p.dkv_aux + (long long)arow * lda + h * 64 + 2 * tg;
        float* ov = ok + D;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          atomicAdd(ok + 8 * n, dk[n][2 * half] * scale);
          atomicAdd(ok + 8 * n + 1, dk[n][2 * half + 1] * scale);
          atomicAdd(ov + 8 * n, dv[n][2 * half]);
          atomicAdd(ov + 8 * n + 1, dv[n][2 * half + 1]);
        }
      }
    }
  }   // key-tile unit
  }   // unit loop
}

// dkv_aux[arow, :] = sum of the per-(frame, head) slots that reference prompt row `arow`, in a fixed order:
//   global row g: all BT frames; local row (b, t'): the T frames of video b; summary row f: frame f alone.
// 128 threads along x = 64 dk + 64 dv values of one head.  Global rows: grid (G, H), 8 thread groups along y take every
// 8th frame and are combined through shared memory in index order; local / summary rows: grid (2 BT, H), blockDim.y = 1.
__global__ void __launch_bounds__(1024) attention_bwd_reduce_kernel(const float* __restrict__ scratch, float* __restrict__ dkv_aux,
                                                                    int row0, int G, int T, int BT, int H) {
  __shared__ float part[8][128];
  const int arow = row0 + blockIdx.x, h = blockIdx.y, c = threadIdx.x, y = threadIdx.y, ny = blockDim.y;
  const int n_aux = G + T + 1, D = H * 64;
  float acc0 = 0.f, acc1 = 0.f;
  auto slot = [&](int f, int a) { return scratch + (((long long)f * H + h) * n_aux + a) * 128 + c; };
  if (arow < G) {
    int f = y;
    for (; f + ny < BT; f += 2 * ny) { acc0 += *slot(f, arow); acc1 += *slot(f + ny, arow); }
    if (f < BT) acc0 += *slot(f, arow);
  } else if (arow < G + BT) {
    const int b = (arow - G) / T, tp = (arow - G) - b * T;
    for (int t = y; t < T; t += ny) acc0 += *slot(b * T + t, G + tp);
  } else if (y == 0) {
    acc0 = *slot(arow - G - BT, G + T);
  }
  float v = acc0 + acc1;
  if (ny > 1) {
    part[y][c] = v;
    __syncthreads();
    if (y != 0) return;
    v = 0.f;
    for (int i = 0; i < ny; ++i) v += part[i][c];
  }
  dkv_aux[(long long)arow * 2 * D + (c < 64 ? h * 64 + c : D + h * 64 + (c - 64))] = v;
}

template <int WARPS, int MIN_CTAS, bool MASKED>
static int launch_attention_bwd_m(const AttnBwdArgs& a, bool have_lse, size_t smem_bytes, cudaStream_t stream) {
  static size_t attr_bytes_dev[64] = {0};
  size_t& attr_bytes = attr_bytes_dev[device_slot()];
  if (smem_bytes > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel<true, WARPS, MIN_CTAS, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_kernel<false, WARPS, MIN_CTAS, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(attention_bwd): %s", cudaGetErrorString(e)); return (int)e; }
    attr_bytes = smem_bytes;
  }
  const dim3 grid(a.H, a.BT);
  if (have_lse) attention_bwd_kernel<true, WARPS, MIN_CTAS, MASKED><<<grid, 32 * WARPS, smem_bytes, stream>>>(a);
  else attention_bwd_kernel<false, WARPS, MIN_CTAS, MASKED><<<grid, 32 * WARPS, smem_bytes, stream>>>(a);
  return GAVA_OK;
}

template <int WARPS, int MIN_CTAS>
static int launch_attention_bwd(const AttnBwdArgs& a, bool have_lse, size_t smem_bytes, cudaStream_t stream) {
  return (a.groups & 3) == 3 ? launch_attention_bwd_m<WARPS, MIN_CTAS, false>(a, have_lse, smem_bytes, stream)
                             : launch_attention_bwd_m<WARPS, MIN_CTAS, true>(a, have_lse, smem_bytes, stream);
}

int attention_bwd(const void* qkv, const void* kv_aux, const void* out, const void* dout, const float* lse, void* dqkv,
                  float* dkv_aux, float* scratch, int B, int T, int N0, int G, int H, int n_q, int key_groups,
                  cudaStream_t stream) {
  if (!qkv || !kv_aux || !out || !dout || !dqkv || !dkv_aux || B <= 0 || T <= 0 || N0 <= 0 || G < 0 || H <= 0 || n_q <= 0 ||
      n_q > N0) {
    set_error("attention_bwd: bad arguments");
    return GAVA_ERR_ARG;
  }
  AttnBwdArgs a;
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  a.kv_aux = reinterpret_cast<const __nv_bfloat16*>(kv_aux);
  a.out = reinterpret_cast<const __nv_bfloat16*>(out);
  a.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  a.dkv_aux = dkv_aux;
  a.lse = lse;
  a.scratch = scratch;
  a.T = T; a.N0 = N0; a.G = G; a.H = H; a.BT = B * T;
  a.S = N0 + G + T + 1;
  a.n_q = n_q;
  a.groups = key_groups;
  a.q_pad = (n_q + 15) / 16 * 16;
  a.s_pad = (a.S + 15) / 16 * 16;
  const size_t smem_bytes = (size_t)(2 * a.q_pad + 2 * a.s_pad) * 128 + (size_t)a.q_pad * 8;
  if (smem_bytes > 226 * 1024) { set_error("attention_bwd: %d keys x %d queries exceed the shared-memory budget", a.S, N0); return GAVA_ERR_SHAPE; }
  const long long D = 64LL * H;
  if (scratch == nullptr) {   // atomics path: accumulate straight into dkv_aux
    cudaError_t e = cudaMemsetAsync(dkv_aux, 0, sizeof(float) * (size_t)(G + 2 * B * T) * 2 * D, stream);
    if (e != cudaSuccess) { set_error("attention_bwd memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  // two CTAs per SM while their shared memory (+1 KB each of system reservation) fits 228 KB; else one CTA of 8 warps
  const bool two_per_sm = 2 * (smem_bytes + 1024) <= 228 * 1024;
  int lrc;
#if GAVA_AB_WARPS > 4
  lrc = two_per_sm ? launch_attention_bwd<8, 2>(a, lse != nullptr, smem_bytes, stream)
                   : launch_attention_bwd<8, 1>(a, lse != nullptr, smem_bytes, stream);
#else
  lrc = two_per_sm ? launch_attention_bwd<4, 2>(a, lse != nullptr, smem_bytes, stream)
                   : launch_attention_bwd<8, 1>(a, lse != nullptr, smem_bytes, stream);
#endif
  if (lrc) return lrc;
  int rc = check_launch("attention_bwd");
  if (rc || scratch == nullptr) return rc;
  if (G > 0) attention_bwd_reduce_kernel<<<dim3(G, H), dim3(128, 8), 0, stream>>>(scratch, dkv_aux, 0, G, T, B * T, H);
  attention_bwd_reduce_kernel<<<dim3(2 * B * T, H), dim3(128, 1), 0, stream>>>(scratch, dkv_aux, G, G, T, B * T, H);
  return check_launch("attention_bwd_reduce");
}

}  // namespace gava
