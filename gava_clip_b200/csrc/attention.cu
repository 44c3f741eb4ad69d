// Fused multi-head self-attention core with a K/V-only prompt prefix (flash-style: scores stay on the SM).
//
// One CTA per (frame, head).  Keys/values of the frame -- its N0 live rows plus the G global-prompt rows
// (shared by every frame), the T local-prompt rows of the frame's video and the frame's summary row --
// are gathered once into swizzled shared memory with cp.async; each warp then owns 16-query tiles and runs
// QK^T -> online softmax (exp2, fp32) -> PV on the tensor cores with the scores held in registers.
// Query rows of prompt tokens are never computed (their outputs are discarded by the reference,
// VitaCLIP_vision_encoder_utils.py:194-197, VitaCLIP_vision_encoder.py:121).
//
// Restates Attention.forward core, training/VitaCLIP_vision_encoder_utils.py:71-77:
//   aff = einsum('nqhc,nkhc->nqkh', q / sqrt(dh), k); softmax over keys; mix = einsum('nqlh,nlhc->nqhc', aff, v)
#include "common.cuh"
#include "kernels.h"
#include "mma_sync.cuh"

#include <stdlib.h>

namespace gava {

struct AttnArgs {
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* kv_aux;
  __nv_bfloat16* out;
  float* lse;   // optional [BT, H, N0] log2-domain log-sum-exp rows
  int T, N0, G, H, n_q, S, n_chunks, BT;
  int groups;   // bit 0: local-prompt keys take part, bit 1: the summary-token key takes part
};

template <int CT>  // 16-key tiles per softmax chunk
__global__ void __launch_bounds__(128, 3) attention_kernel(const AttnArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int h = blockIdx.x, f = blockIdx.y;
  const int D = p.H * 64;
  const int S_pad = p.n_chunks * CT * 16;
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + (uint32_t)S_pad * 128u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tg = lane & 3;
  const long long ldq = 3LL * D, lda = 2LL * D;
  const int b = f / p.T;

  // ---- gather K/V rows into shared memory
  for (int i = tid; i < S_pad * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    const __nv_bfloat16* ksrc = nullptr;
    const __nv_bfloat16* vsrc = nullptr;
    if (row < p.N0) {
      ksrc = p.qkv + ((long long)f * p.N0 + row) * ldq + D + h * 64;
      vsrc = ksrc + D;
    } else if (row < p.S) {
      const int a = row - p.N0;
      int arow;
      if (a < p.G) arow = a;                                  // global prompts
      else if (a < p.G + p.T) arow = p.G + b * p.T + (a - p.G);  // local prompts of this video
      else arow = p.G + p.BT + f;                             // summary token of this frame
      ksrc = p.kv_aux + (long long)arow * lda + h * 64;
      vsrc = ksrc + D;
    }
    const uint32_t off = kv_off(row, ch);
    if (ksrc != nullptr) {
      cp_async16(sK + off, ksrc + ch * 8);
      cp_async16(sV + off, vsrc + ch * 8);
    } else {
      *reinterpret_cast<uint4*>(smem + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(smem + (size_t)S_pad * 128 + off) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const float scale_log2 = 0.125f * 1.4426950408889634f;  // dh^-0.5 * log2(e), dh = 64
  const int n_tiles = (p.n_q + 15) >> 4;
  for (int mt = warp; mt < n_tiles; mt += 4) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    // ---- Q fragments straight from global memory (16 rows x 64)
    uint32_t qf[4][4];
    {
      const __nv_bfloat16* q0 = p.qkv + ((long long)f * p.N0 + r0) * ldq + h * 64 + 2 * tg;
      const __nv_bfloat16* q1 = q0 + 8 * ldq;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        qf[s][0] = (r0 < p.n_q) ? *reinterpret_cast<const uint32_t*>(q0 + 16 * s) : 0u;
        qf[s][1] = (r1 < p.n_q) ? *reinterpret_cast<const uint32_t*>(q1 + 16 * s) : 0u;
        qf[s][2] = (r0 < p.n_q) ? *reinterpret_cast<const uint32_t*>(q0 + 16 * s + 8) : 0u;
        qf[s][3] = (r1 < p.n_q) ? *reinterpret_cast<const uint32_t*>(q1 + 16 * s + 8) : 0u;
      }
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }

    for (int c = 0; c < p.n_chunks; ++c) {
      const int kbase = c * CT * 16;
      float s[2 * CT][4];
#pragma unroll
      for (int n = 0; n < 2 * CT; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
      // ---- S = Q K^T
#pragma unroll
      for (int nt2 = 0; nt2 < CT; ++nt2) {
        const int mi = lane >> 3;
        const int key = kbase + 16 * nt2 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(sK + kv_off(key, 2 * ks + (mi & 1)), b0, b1, b2, b3);
          mma_bf16_16816(s[2 * nt2], qf[ks], b0, b1);
          mma_bf16_16816(s[2 * nt2 + 1], qf[ks], b2, b3);
        }
      }
      // ---- scale, mask, chunk max
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < 2 * CT; ++n) {
        const int key = kbase + 8 * n + 2 * tg;
        s[n][0] = (key < p.S) ? s[n][0] * scale_log2 : -INFINITY;
        s[n][1] = (key + 1 < p.S) ? s[n][1] * scale_log2 : -INFINITY;
        s[n][2] = (key < p.S) ? s[n][2] * scale_log2 : -INFINITY;
        s[n][3] = (key + 1 < p.S) ? s[n][3] * scale_log2 : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      // a fully masked chunk (only possible for padding chunks) keeps the running state unchanged
      const float sub0 = (mn0 == -INFINITY) ? 0.f : mn0, sub1 = (mn1 == -INFINITY) ? 0.f : mn1;
      const float al0 = exp2f(m0 - sub0), al1 = exp2f(m1 - sub1);
      m0 = mn0; m1 = mn1;
      l0 *= al0; l1 *= al1;
#pragma unroll
      for (int n = 0; n < 8; ++n) { o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1; }
#pragma unroll
      for (int n = 0; n < 2 * CT; ++n) {
        s[n][0] = exp2f(s[n][0] - sub0); s[n][1] = exp2f(s[n][1] - sub0);
        s[n][2] = exp2f(s[n][2] - sub1); s[n][3] = exp2f(s[n][3] - sub1);
        l0 += s[n][0] + s[n][1];
        l1 += s[n][2] + s[n][3];
      }
      // ---- O += P V
#pragma unroll
      for (int kt = 0; kt < CT; ++kt) {
        uint32_t a[4];
        a[0] = pack_bf16x2(s[2 * kt][0], s[2 * kt][1]);
        a[1] = pack_bf16x2(s[2 * kt][2], s[2 * kt][3]);
        a[2] = pack_bf16x2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
        a[3] = pack_bf16x2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
        const int mi = lane >> 3;
        const int key = kbase + 16 * kt + (mi & 1) * 8 + (lane & 7);
#pragma unroll
        for (int nd2 = 0; nd2 < 4; ++nd2) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(sV + kv_off(key, 2 * nd2 + (mi >> 1)), b0, b1, b2, b3);
          mma_bf16_16816(o[2 * nd2], a, b0, b1);
          mma_bf16_16816(o[2 * nd2 + 1], a, b2, b3);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    if (p.lse != nullptr && tg == 0) {
      float* ls = p.lse + ((long long)f * p.H + h) * p.N0;
      if (r0 < p.n_q) ls[r0] = m0 + log2f(l0);
      if (r1 < p.n_q) ls[r1] = m1 + log2f(l1);
    }
    __nv_bfloat16* o0 = p.out + ((long long)f * p.N0 + r0) * D + h * 64 + 2 * tg;
    __nv_bfloat16* o1 = o0 + 8LL * D;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (r0 < p.n_q) *reinterpret_cast<uint32_t*>(o0 + 8 * n) = pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
      if (r1 < p.n_q) *reinterpret_cast<uint32_t*>(o1 + 8 * n) = pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Single-query attention (n_q == 1): the last block only needs the cls row of every frame (vision_encoder.py:126), i.e. one
// query against the frame's ~214 keys per (frame, head) -- 110 KB of K/V traffic for 55 kFLOP.  The tensor-core kernels
// above still stage whole 128-row query tiles for it (178 us at config 2); this one is a plain bandwidth kernel: one warp per
// (frame, head), lanes over keys for the scores (each lane reads whole 128-byte K rows), lanes over channels for P V.
__global__ void __launch_bounds__(256) attention_q1_kernel(const AttnArgs p) {
  __shared__ float s_sc[8][320];   // scores / probabilities of the warp's item (S <= 320)
  const int wid = threadIdx.x >> 5;
  const int item = blockIdx.x * 8 + wid;
  if (item >= p.BT * p.H) return;
  const int f = item / p.H, h = item - f * p.H;
  const int lane = threadIdx.x & 31;
  const int sub = lane >> 3, ch = lane & 7;   // 4 key rows per step, 8 lanes (16 bytes each) per 128-byte row
  const int D = p.H * 64;
  const long long ldq = 3LL * D, lda = 2LL * D;
  const int b = f / p.T;
  float* sc = s_sc[wid];
  // this lane's 8 channels of the query row (cls) of the frame / head
  float q[8];
  {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.qkv + (long long)f * p.N0 * ldq + h * 64) + ch);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
      q[2 * j] = __bfloat162float(v.x); q[2 * j + 1] = __bfloat162float(v.y);
    }
  }
  auto key_ptr = [&](int row) -> const __nv_bfloat16* {
    if (row < p.N0) return p.qkv + ((long long)f * p.N0 + row) * ldq + D + h * 64;
    const int a = row - p.N0;
    int arow;
    if (a < p.G) arow = a;
    else if (a < p.G + p.T) arow = p.G + b * p.T + (a - p.G);
    else arow = p.G + p.BT + f;
    return p.kv_aux + (long long)arow * lda + h * 64;
  };
  const float scale_log2 = 0.125f * 1.4426950408889634f;
  const int steps = (p.S + 3) >> 2;
  // ---- scores: every warp instruction reads 4 whole K rows (512 contiguous-by-row bytes)
  float mx = -INFINITY;
#pragma unroll 4
  for (int st = 0; st < steps; ++st) {
    const int row = 4 * st + sub;
    float d = 0.f;
    if (row < p.S) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(key_ptr(row)) + ch);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        d += __bfloat162float(v.x) * q[2 * j] + __bfloat162float(v.y) * q[2 * j + 1];
      }
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    bool on = row < p.S;
    if (on && row >= p.N0 + p.G) on = (row < p.N0 + p.G + p.T) ? (p.groups & 1) : (p.groups & 2);   // disabled prompt group
    const float v = on ? d * scale_log2 : -INFINITY;
    if (ch == 0 && row < p.S) sc[row] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  __syncwarp();
  float sum = 0.f;
  for (int r = lane; r < p.S; r += 32) {
    const float e = exp2f(sc[r] - mx);
    sc[r] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  // ---- O = P V with the same 4-rows-per-instruction access; the 4 row sub-groups are combined at the end
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll 4
  for (int st = 0; st < steps; ++st) {
    const int row = 4 * st + sub;
    if (row < p.S) {
      const float pk = sc[row];
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(key_ptr(row) + D) + ch);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        o[2 * j] += pk * __bfloat162float(v.x); o[2 * j + 1] += pk * __bfloat162float(v.y);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    o[j] += __shfl_xor_sync(0xffffffffu, o[j], 8);
    o[j] += __shfl_xor_sync(0xffffffffu, o[j], 16);
  }
  if (sub == 0) {
    const float inv = 1.0f / sum;
    uint4 w;
    w.x = pack_bf16x2(o[0] * inv, o[1] * inv); w.y = pack_bf16x2(o[2] * inv, o[3] * inv);
    w.z = pack_bf16x2(o[4] * inv, o[5] * inv); w.w = pack_bf16x2(o[6] * inv, o[7] * inv);
    *(reinterpret_cast<uint4*>(p.out + (long long)f * p.N0 * D + h * 64) + ch) = w;
  }
  if (p.lse != nullptr && lane == 0) p.lse[((long long)f * p.H + h) * p.N0] = mx + log2f(sum);
}

template <int CT>
static int launch_attention(const AttnArgs& a, int frames, size_t smem_bytes, cudaStream_t stream) {
  static size_t attr_bytes_dev[64] = {0};
  size_t& attr_bytes = attr_bytes_dev[device_slot()];
  auto kern = attention_kernel<CT>;
  if (smem_bytes > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e)); return (int)e; }
    attr_bytes = smem_bytes;
  }
  kern<<<dim3(a.H, frames), 128, smem_bytes, stream>>>(a);
  return check_launch("attention");
}

int attention(const void* qkv, const void* kv_aux, void* out, float* lse, int B, int T, int N0, int G, int H, int n_q,
              int key_groups, cudaStream_t stream) {
  if (!qkv || !kv_aux || !out || B <= 0 || T <= 0 || N0 <= 0 || G < 0 || H <= 0 || n_q <= 0 || n_q > N0) {
    set_error("attention: bad arguments");
    return GAVA_ERR_ARG;
  }
  // tcgen05 path for every shape inside its specialisation (G = 8, <= 320 smem key rows: all BASELINE configs incl.
  // ViT-L/14 T=16 and the 70-frame recipe); the mma.sync kernel below covers what is left (other G, longer clips).
  // GAVA_ATTN_TC=0 forces the latter (A/B tests).
  static const bool use_q1 = [] { const char* e = getenv("GAVA_ATTN_Q1"); return !(e && e[0] == '0'); }();
  if (use_q1 && n_q == 1 && N0 + G + T + 1 <= 320) {
    AttnArgs a;
    a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
    a.kv_aux = reinterpret_cast<const __nv_bfloat16*>(kv_aux);
    a.out = reinterpret_cast<__nv_bfloat16*>(out);
    a.lse = lse;
    a.T = T; a.N0 = N0; a.G = G; a.H = H; a.n_q = 1; a.BT = B * T; a.S = N0 + G + T + 1; a.n_chunks = 0; a.groups = key_groups;
    attention_q1_kernel<<<(unsigned)((B * T * H + 7) / 8), 256, 0, stream>>>(a);
    return check_launch("attention_q1");
  }
  static const bool use_tc = [] { const char* e = getenv("GAVA_ATTN_TC"); return !(e && e[0] == '0'); }();
  if (use_tc && attention_tc_keys(T, N0, G) != 0 && n_q <= 384) return attention_tc(qkv, kv_aux, out, lse, B, T, N0, G, H, n_q, key_groups, stream);
  if ((key_groups & 3) != 3) {
    set_error("attention: the mma.sync kernel (shapes outside the tcgen05 specialisation) needs all prompt groups (key_groups = 3)");
    return GAVA_ERR_SHAPE;
  }
  AttnArgs a;
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  a.kv_aux = reinterpret_cast<const __nv_bfloat16*>(kv_aux);
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.lse = lse;
  a.groups = 3;
  a.T = T; a.N0 = N0; a.G = G; a.H = H; a.n_q = n_q; a.BT = B * T;
  a.S = N0 + G + T + 1;
  const int tiles16 = (a.S + 15) / 16;
  const int ct = tiles16 <= 14 ? 7 : 6;
  a.n_chunks = (tiles16 + ct - 1) / ct;
  const size_t smem_bytes = (size_t)a.n_chunks * ct * 16 * 128 * 2;
  if (smem_bytes > 200 * 1024) { set_error("attention: %d keys per frame exceed the shared-memory budget", a.S); return GAVA_ERR_SHAPE; }
  return ct == 7 ? launch_attention<7>(a, B * T, smem_bytes, stream) : launch_attention<6>(a, B * T, smem_bytes, stream);
}

}  // namespace gava
