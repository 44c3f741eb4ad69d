// tcgen05 / TMEM / TMA version of the prompt-prefix attention core (same contract as attention.cu).
//
// Work item = (frame f, head h); a persistent CTA streams items through a 2-stage TMA ring:
//   smem stage : Q tiles 0/1 [128 x 64], K [NK x 64], V [NK x 64]  (bf16, 128-byte swizzle, K/V rows gathered by
//                several TMA boxes: live rows of the frame | 8 global prompts | T local prompts | summary token)
//   TMEM       : S_w = Q_w K^T (128 x NK fp32) for the two query tiles w = 0,1 at columns 0 / 256;
//                P_w (bf16, unnormalised probabilities) overwrites S_w in place, each key-column half inside its own
//                columns (keys [0,NK/2) -> cols [0,NK/4), keys [NK/2,NK) -> cols [NK/2, 3NK/4));
//                O_w = P_w V (128 x 64 fp32) lives at columns +192..+255 of the same region
//   roles      : warp 0 TMA producer, warps 1/2 MMA issuers of query tile 0/1 (warp 2 also allocates TMEM), warps 3-18 softmax/output:
//                8 warps per query tile, TWO threads per query row (each owns one half of the key columns; row max and
//                row sum are exchanged through shared memory), so four independent softmax streams hide the
//                TMEM-load / MUFU latency chains
// QK^T is an SS MMA (both operands K-major in smem); P V is a TS MMA (A = P from TMEM, B = V MN-major in smem).
//
// Restates Attention.forward core, training/VitaCLIP_vision_encoder_utils.py:71-77 (see attention.cu).
#include "common.cuh"
#include "kernels.h"

#ifndef GAVA_ATTN_DEBUG
#define GAVA_ATTN_DEBUG 0
#endif
#if GAVA_ATTN_DEBUG
#define ATC_T0() const long long _t0 = clock64()
#define ATC_ACC(x) x += clock64() - _t0
#else
#define ATC_T0()
#define ATC_ACC(x)
#endif

namespace gava {

constexpr int ATC_THREADS = 608;   // 19 warps
constexpr int ATC_QTILE_BYTES = 128 * 128;  // 128 rows x 64 bf16

struct AttnTcArgs {
  __nv_bfloat16* out;
  float* lse;      // optional [BT, H, N0]: log2-domain log-sum-exp of every query row (saved for the backward)
  int T, N0, G, H, n_q, BT, D;
  int kb_live;     // N0 rounded up to 8: smem row where the prompt rows start
  int num_items;   // BT * H
  uint32_t stage_tx_bytes;
};

// TS MMA: D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major (rows = K index, 64 contiguous N elements per 128-byte row) smem descriptor, 128-byte swizzle:
// SBO = 1024 B between 8-row groups along K; LBO = stride between 64-wide N atoms (single atom here).
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(uint32_t M, uint32_t N) {  // B operand MN-major
  return make_idesc_bf16(M, N) | (1u << 16);
}

template <bool MASK>
__device__ __forceinline__ float chunk16_max(const uint32_t (&v)[16], int j0, int v0, int v1, int v2, float mx) {
  float m0 = mx, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
    if (MASK) {
      const int k = j0 + j;
      if (!(k < v0 || (k >= v1 && k <= v2))) a0 = -INFINITY;
      if (!(k + 1 < v0 || (k + 1 >= v1 && k + 1 <= v2))) a1 = -INFINITY;
    }
    m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1);
  }
  return fmaxf(m0, m1);
}
template <bool MASK>
__device__ __forceinline__ float chunk16_exp(const uint32_t (&v)[16], uint32_t (&pk)[8], int j0, int v0, int v1, int v2,
                                             float c, float mc) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    float p0 = exp2f(fmaf(__uint_as_float(v[j]), c, -mc));
    float p1 = exp2f(fmaf(__uint_as_float(v[j + 1]), c, -mc));
    if (MASK) {
      const int k = j0 + j;
      if (!(k < v0 || (k >= v1 && k <= v2))) p0 = 0.f;
      if (!(k + 1 < v0 || (k + 1 >= v1 && k + 1 <= v2))) p1 = 0.f;
    }
    s0 += p0; s1 += p1;
    pk[j >> 1] = pack_bf16x2(p0, p1);
  }
  return s0 + s1;
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

template <int NK>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    const __grid_constant__ CUtensorMap tm_aux, const AttnTcArgs p) {
  constexpr int kKVBytes = NK * 128;
  constexpr int kStageBytes = 2 * ATC_QTILE_BYTES + 2 * kKVBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytes);
  uint64_t* kv_full = bars;         // [2] TMA -> MMA
  uint64_t* kv_empty = bars + 2;    // [2] MMA -> TMA
  uint64_t* s_full = bars + 4;      // [2] MMA -> softmax group w
  uint64_t* p_ready = bars + 6;     // [2] softmax group w -> MMA
  uint64_t* o_full = bars + 8;      // [2] MMA -> softmax group w
  uint64_t* s_free = bars + 10;     // [2] softmax group w -> MMA (S/P/O region of tile w reusable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* xch = reinterpret_cast<float*>(bars + 16);   // [tile 2][max|sum 2][half 2][row 128] row statistics exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = (p.n_q + 127) >> 7;  // query tiles per item (1 or 2)

  // K/V rows that no TMA box covers (and nothing else) must still be finite: zero both stages once
  for (int i = threadIdx.x; i < 2 * kStageBytes / 16; i += ATC_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_aux);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], n_units);   // one tcgen05.commit per query tile
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 256);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int f = item / p.H, h = item - f * p.H;
        const int b = f / p.T;
        uint8_t* sQ = smem + st * kStageBytes;
        uint8_t* sK = sQ + 2 * ATC_QTILE_BYTES;
        uint8_t* sV = sK + kKVBytes;
        mbar_wait(&kv_empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&kv_full[st], p.stage_tx_bytes);
        const int xq = h * 64, xk = p.D + h * 64, xv = 2 * p.D + h * 64;
        const int row = f * p.N0;
        tma_load_2d(sQ, &tm_q, &kv_full[st], xq, row);
        tma_load_2d(sQ + ATC_QTILE_BYTES, &tm_q, &kv_full[st], xq, row + 128);
        tma_load_2d(sK, &tm_kv, &kv_full[st], xk, row);
        tma_load_2d(sV, &tm_kv, &kv_full[st], xv, row);
        int r = p.kb_live;                       // prompt rows, 8 per box
        tma_load_2d(sK + r * 128, &tm_aux, &kv_full[st], h * 64, 0);
        tma_load_2d(sV + r * 128, &tm_aux, &kv_full[st], p.D + h * 64, 0);
        r += 8;
        for (int i = 0; i < p.T; i += 8, r += 8) {
          tma_load_2d(sK + r * 128, &tm_aux, &kv_full[st], h * 64, p.G + b * p.T + i);
          tma_load_2d(sV + r * 128, &tm_aux, &kv_full[st], p.D + h * 64, p.G + b * p.T + i);
        }
        tma_load_2d(sK + r * 128, &tm_aux, &kv_full[st], h * 64, p.G + p.BT + f);
        tma_load_2d(sV + r * 128, &tm_aux, &kv_full[st], p.D + h * 64, p.G + p.BT + f);
        if (++st == 2) { st = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: warp 1 drives query tile 0, warp 2 query tile 1 =====================
    // One in-order issuer for both tiles would couple them (S of tile 0 for the next item could not be issued while the
    // issuer waits for tile 1's probabilities); with one issuer per tile the two softmax groups run fully decoupled.
    const int w = warp - 1;
    if (lane == 0 && w < n_units) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, NK);
      constexpr uint32_t idesc_o = make_idesc_bf16_bmn(128, 64);
      int st = 0;
      uint32_t ph = 0, it_ph = 0;
#if GAVA_ATTN_DEBUG
      long long d_kv = 0, d_sfree = 0, d_pready = 0, d_start = clock64();
#endif
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        uint8_t* sQ = smem + st * kStageBytes;
        uint8_t* sK = sQ + 2 * ATC_QTILE_BYTES;
        uint8_t* sV = sK + kKVBytes;
        { ATC_T0(); mbar_wait(&kv_full[st], ph); ATC_ACC(d_kv); }
        { ATC_T0(); mbar_wait(&s_free[w], it_ph ^ 1u); ATC_ACC(d_sfree); }
        tc_fence_after();
        const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK));
        const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ + w * ATC_QTILE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + (uint32_t)(w * 256), dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
        umma_commit(&s_full[w]);
        const uint64_t dv = make_sw128_mnmajor_desc(smem_u32(sV), (uint32_t)kKVBytes);
        { ATC_T0(); mbar_wait(&p_ready[w], it_ph); ATC_ACC(d_pready); }
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < NK / 16; ++k) {  // 16 keys per MMA; V advances 16 rows (2048 B); P of key half 0 sits at
          // columns 8k, P of key half 1 at NK/2 + 8(k - NK/32) (each half writes P inside its own S columns)
          const int pcol = k < NK / 32 ? 8 * k : NK / 2 + 8 * (k - NK / 32);
          umma_bf16_ts(tmem_base + (uint32_t)(w * 256 + 192), tmem_base + (uint32_t)(w * 256 + pcol),
                       dv + (uint64_t)(128 * k), idesc_o, (uint32_t)(k != 0));
        }
        umma_commit(&o_full[w]);
        umma_commit(&kv_empty[st]);   // one arrival per tile: the stage is free once both tiles' MMAs have retired
        it_ph ^= 1u;
        if (++st == 2) { st = 0; ph ^= 1u; }
      }
#if GAVA_ATTN_DEBUG
      if (blockIdx.x == 3) printf("[attn mma %d] total %lld: wait kv_full %lld, s_free %lld, p_ready %lld\n", w, clock64() - d_start, d_kv, d_sfree, d_pready);
#endif
    }
    __syncwarp();
  } else if (warp >= 3) {
    // ===================== softmax + output: 2 threads per query row =====================
    const int idx = warp - 3;               // 0..15
    const int w = idx >> 3;                 // query tile
    const int hh = (idx >> 2) & 1;          // key-column half owned by this thread
    const int quarter = warp & 3;           // TMEM lane quarter (hardware rule: warp % 4)
    const int row_in_tile = quarter * 32 + lane;
    constexpr int HALF = NK / 2;            // keys per half (112 or 128): HALF/16 chunks of 16
    const uint32_t t_s = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(w * 256);
    const uint32_t t_mine = t_s + (uint32_t)(hh * HALF);
    const float c = 0.125f * 1.4426950408889634f;  // dh^-0.5 * log2(e)
    const int v0 = p.N0, v1 = p.kb_live, v2 = p.kb_live + 8 + p.T;   // valid keys: [0,v0) u [v1,v2]
    float* my_max = xch + ((w * 2 + 0) * 2 + hh) * 128 + row_in_tile;
    float* ot_max = xch + ((w * 2 + 0) * 2 + (hh ^ 1)) * 128 + row_in_tile;
    float* my_sum = xch + ((w * 2 + 1) * 2 + hh) * 128 + row_in_tile;
    float* ot_sum = xch + ((w * 2 + 1) * 2 + (hh ^ 1)) * 128 + row_in_tile;
    uint32_t it_ph = 0;
#if GAVA_ATTN_DEBUG
    long long d_sfull = 0, d_bar = 0, d_ofull = 0, d_start = clock64();
#endif
    if (w < n_units) {
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int f = item / p.H, h = item - f * p.H;
        { ATC_T0(); mbar_wait(&s_full[w], it_ph); ATC_ACC(d_sfull); }
        tc_fence_after();
        if (w * 128 + quarter * 32 >= p.n_q) {
          // All 32 query rows of this warp lie beyond n_q (e.g. rows 224..255 of a 197-row frame): nothing to compute.
          // MMA rows are independent, so the garbage left in these P/O rows never reaches a stored row; the warp only
          // keeps the barrier protocol going (named barrier with its partner warps, p_ready, s_free).
          asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");
          mbar_arrive(&p_ready[w]);
          mbar_wait(&o_full[w], it_ph);
          mbar_arrive(&s_free[w]);
          it_ph ^= 1u;
          continue;
        }
        // ---- pass 1: maximum over this thread's half of the keys, then exchange with the partner thread
        float mx = -INFINITY;
        {   // the reduction per chunk is ~16 instructions, far shorter than a TMEM load: keep THREE loads in flight
          uint32_t va[16], vb[16], vc[16];
          constexpr int NCH = HALF / 16;
          tmem_ld_32x16(t_mine, va);
          if (NCH > 1) tmem_ld_32x16(t_mine + 16u, vb);
          if (NCH > 2) tmem_ld_32x16(t_mine + 32u, vc);
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            // tcgen05.wait::ld waits for ALL outstanding loads of the thread, so loads are issued in groups of three
            if (ch % 3 == 0) tmem_ld_wait();
            uint32_t (&cur)[16] = (ch % 3 == 0) ? va : ((ch % 3 == 1) ? vb : vc);
            const int j0 = hh * HALF + ch * 16;
            if (j0 + 16 <= v0) mx = chunk16_max<false>(cur, j0, v0, v1, v2, mx);
            else mx = chunk16_max<true>(cur, j0, v0, v1, v2, mx);
            if (ch % 3 == 2 && ch + 1 < NCH) {   // group consumed: issue the next three
              tmem_ld_32x16(t_mine + (uint32_t)((ch + 1) * 16), va);
              if (ch + 2 < NCH) tmem_ld_32x16(t_mine + (uint32_t)((ch + 2) * 16), vb);
              if (ch + 3 < NCH) tmem_ld_32x16(t_mine + (uint32_t)((ch + 3) * 16), vc);
            }
          }
        }
        *my_max = mx;
        { ATC_T0(); asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory"); ATC_ACC(d_bar); }
        mx = fmaxf(mx, *ot_max);
        const float mc = mx * c;
        // ---- pass 2: p = exp2(s*c - m*c); P (bf16) overwrites S inside this thread's own columns
        float sum = 0.f;
        {
          uint32_t va[16], vb[16];
          tmem_ld_32x16(t_mine, va);
#pragma unroll
          for (int ch = 0; ch < HALF / 16; ++ch) {
            tmem_ld_wait();
            uint32_t (&cur)[16] = (ch & 1) ? vb : va;
            uint32_t (&nxt)[16] = (ch & 1) ? va : vb;
            if (ch + 1 < HALF / 16) tmem_ld_32x16(t_mine + (uint32_t)((ch + 1) * 16), nxt);
            uint32_t pk[8];
            const int j0 = hh * HALF + ch * 16;
            if (j0 + 16 <= v0) sum += chunk16_exp<false>(cur, pk, j0, v0, v1, v2, c, mc);
            else sum += chunk16_exp<true>(cur, pk, j0, v0, v1, v2, c, mc);
            tmem_st_32x8(t_mine + (uint32_t)(ch * 8), pk);   // columns [8 ch, 8 ch + 8) of the own half: already consumed
          }
        }
        *my_sum = sum;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[w]);
        // ---- O = P V done: normalise, cast, store this thread's 32 of the row's 64 output values
        { ATC_T0(); mbar_wait(&o_full[w], it_ph); ATC_ACC(d_ofull); }
        tc_fence_after();
        const float tot = sum + *ot_sum;            // partner's sum was written before its p_ready arrive
        const float inv = 1.0f / tot;
        const int qrow = w * 128 + row_in_tile;
        if (p.lse != nullptr && hh == 0 && qrow < p.n_q) p.lse[((long long)f * p.H + h) * p.N0 + qrow] = mc + log2f(tot);
        __nv_bfloat16* orow = p.out + ((long long)f * p.N0 + qrow) * p.D + h * 64 + hh * 32;
#pragma unroll
        for (int d0 = 0; d0 < 32; d0 += 16) {
          uint32_t v[16];
          tmem_ld_32x16(t_s + 192u + (uint32_t)(hh * 32 + d0), v);
          tmem_ld_wait();
          if (qrow < p.n_q) {
            uint4 a, b2;
            a.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
            a.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
            a.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
            a.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
            b2.x = pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
            b2.y = pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
            b2.z = pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
            b2.w = pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
            *reinterpret_cast<uint4*>(orow + d0) = a;
            *reinterpret_cast<uint4*>(orow + d0 + 8) = b2;
          }
        }
        tc_fence_before();
        mbar_arrive(&s_free[w]);
        it_ph ^= 1u;
      }
#if GAVA_ATTN_DEBUG
      if (blockIdx.x == 3 && lane == 0 && (idx == 0 || idx == 12))
        printf("[attn smx warp %d] total %lld: wait s_full %lld, bar %lld, o_full %lld\n", idx, clock64() - d_start, d_sfull, d_bar, d_ofull);
#endif
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NK>
static int launch_attention_tc(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& taux,
                               const AttnTcArgs& a, cudaStream_t stream) {
  constexpr int smem_bytes = 2 * (2 * ATC_QTILE_BYTES + 2 * NK * 128) + 1024 + 128 + 2 * 2 * 2 * 128 * 4;
  static bool attr_set = false;
  auto kern = attention_tc_kernel<NK>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(attention_tc): %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  const int grid = a.num_items < num_sms() ? a.num_items : num_sms();
  kern<<<grid, ATC_THREADS, smem_bytes, stream>>>(tq, tkv, taux, a);
  return check_launch("attention_tc");
}

int attention_tc(const void* qkv, const void* kv_aux, void* out, float* lse, int B, int T, int N0, int G, int H, int n_q,
                 cudaStream_t stream) {
  const int kb_live = (N0 + 7) / 8 * 8;
  const int keys = kb_live + 8 + T + 8;          // live | global | local | summary box
  if (G != 8 || (T % 8) != 0 || kb_live > 256 || keys > 256 || N0 > 256) {
    set_error("attention_tc: shape outside the tcgen05 specialisation (G=%d T=%d N0=%d)", G, T, N0);
    return GAVA_ERR_SHAPE;
  }
  const int nk = keys <= 224 ? 224 : 256;   // softmax works in 32-key chunks
  const int D = H * 64, BT = B * T;
  CUtensorMap tq, tkv, taux;
  int rc = make_tmap_bf16(&tq, qkv, (long long)BT * N0, 3LL * D, 3LL * D, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tkv, qkv, (long long)BT * N0, 3LL * D, 3LL * D, kb_live);
  if (rc) return rc;
  rc = make_tmap_bf16(&taux, kv_aux, (long long)G + 2LL * BT, 2LL * D, 2LL * D, 8);
  if (rc) return rc;
  AttnTcArgs a;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.lse = lse;
  a.T = T; a.N0 = N0; a.G = G; a.H = H; a.n_q = n_q; a.BT = BT; a.D = D;
  a.kb_live = kb_live;
  a.num_items = BT * H;
  a.stage_tx_bytes = (uint32_t)(2 * ATC_QTILE_BYTES + 2 * 128 * keys);
  if (nk == 224) return launch_attention_tc<224>(tq, tkv, taux, a, stream);
  return launch_attention_tc<256>(tq, tkv, taux, a, stream);
}

}  // namespace gava
