// tcgen05 / TMEM / TMA version of the prompt-prefix attention core (same contract as attention.cu).
//
// Work item = (frame f, head h); its query rows are cut into 128-row UNITS (1..3 per item).  A persistent CTA streams
// items through a 2-stage K/V ring and units through a ring of Q tiles:
//   smem       : K [NB*KB x 64], V [NB*KB x 64] per K/V stage (bf16, 128-byte swizzle; rows gathered by TMA boxes:
//                live rows of the frame | 8 global prompts | T local prompts | summary token), Q tiles [128 x 64]
//   TMEM       : two SLOTS of 256 columns; unit u runs in slot u & 1:
//                S = Q K_blk^T (128 x KB fp32) at columns [0, KB) of the slot, one key block (KB keys) at a time;
//                P (bf16, unnormalised probabilities) overwrites S in place, each key-column half inside its own columns
//                (keys [0,KB/2) -> cols [0,KB/4), keys [KB/2,KB) -> cols [KB/2, 3KB/4));
//                O += P V_blk (128 x 64 fp32) at columns [192, 256) of the slot
//   roles      : warp 0 TMA producer, warps 1/2 MMA issuers of slot 0/1 (warp 2 also allocates TMEM), warps 3-18
//                softmax/output: 8 warps per slot, TWO threads per query row (each owns one half of the key columns)
// QK^T is an SS MMA (both operands K-major in smem); P V is a TS MMA (A = P from TMEM, B = V MN-major in smem).
//
// ONE pass over S.  The kernel is bound by MUFU.EX2 (16 exp2 per clock per SM, tools/cuda/tmem_bw.cu) and by the chain of
// round trips MMA -> TMEM -> registers -> TMEM -> MMA, so S is read exactly once: softmax is shift invariant, and instead
// of the exact row maximum (a second read of S) every row uses m = max over its first 32 scores (16 of each key half,
// exchanged through shared memory).  m <= true maximum, hence the largest probability is >= 1 (no underflow of the row
// sum) and all of exp2(s - m), their bf16 roundings, the fp32 row sum and the fp32 P V accumulation keep full relative
// precision as long as nothing overflows.  A row whose sum exceeds 2^100 (a later score beat the first 32 by > 69 nats:
// not seen with real weights, but possible) is recomputed exactly by its two threads on the CUDA cores from global memory
// (attn_row_exact).  With the same stale m for every key block, long rows (NB = 2 blocks: 257..320 keys, ViT-L/14 and
// the 70-frame recipe) need no rescaling of O either.
//
// Key masking costs the softmax warps nothing: invalid smem key rows (padding between the live rows and the prompt rows,
// the tails of the 8-row prompt boxes) get a score of -1e30 from a FIFTH k-step of the S MMA, whose A operand is a constant
// tile with a column of ones and whose B operand holds 0 / -1e30 per key row (the pattern is the same for every item).
// Both are tiny no-swizzle K-major tiles built once per CTA.  exp2 then returns exact zeros for those keys.
//
// The two softmax groups take TURNS on the exp pass (mbarrier token, one key block each): left alone they run in lock
// step -- both in the MUFU-bound exp pass, then both in the epilogue, then both waiting for their MMAs -- and the MUFU
// pipe idles two thirds of the time; alternating, one group's exp pass covers the other's MMA round trips and epilogue.
// Epilogue: O (fp32, TMEM) -> registers -> normalised bf16 rows staged in the unit's own Q tile (free once the unit's
// MMAs retired; same 128-byte swizzle) -> ONE TMA store per unit through a 3-D map [frame][row < n_q][D] whose bounds
// clip the rows beyond n_q (32 scattered 16-byte stores per warp instruction made the epilogue LSU-bound before).
//
// Restates Attention.forward core, training/VitaCLIP_vision_encoder_utils.py:71-77 (see attention.cu).
#include "common.cuh"
#include "kernels.h"
#include "attention_tc.cuh"

#ifndef GAVA_ATTN_DEBUG
#define GAVA_ATTN_DEBUG 0
#endif
#ifndef GAVA_ATTN_SAFE_ORDER
#define GAVA_ATTN_SAFE_ORDER 0   // 1: wait for P V of block b to retire before S of block b+1 is issued over P's columns
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
constexpr float ATC_SUM_LIMIT = 1.2676506e30f;   // 2^100: above this a row is recomputed exactly

struct AttnTcArgs {
  const __nv_bfloat16* qkv;     // [BT*N0, 3D]
  const __nv_bfloat16* kvaux;   // [G + 2BT, 2D]
  float* lse;      // optional [BT, H, N0]: log2-domain log-sum-exp of every query row (saved for the backward)
  int T, N0, G, H, n_q, BT, D;
  int kb_live;     // N0 rounded up to 8: smem row where the prompt rows start
  int live_box;    // rows per TMA box of the live K/V rows (divides kb_live, <= 256)
  int t_pad;       // T rounded up to 8 (local prompts are fetched in 8-row boxes)
  int num_items;   // BT * H
  int groups;      // bit 0: local-prompt keys valid, bit 1: summary-token key valid
  int n_units;     // query tiles per item
  uint32_t kv_tx_bytes;
};

__device__ __forceinline__ float chunk16_max(const uint32_t (&v)[16]) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    m0 = fmaxf(m0, __uint_as_float(v[j])); m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
  }
  return fmaxf(m0, m1);
}
// p = exp2(s c - mc) for 16 scores -> 8 packed bf16 pairs; returns their fp32 sum (masked keys arrive as s = -1e30: p = 0)
__device__ __forceinline__ float chunk16_exp(const uint32_t (&v)[16], uint32_t (&pk)[8], float c, float mc) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float p0 = exp2f(fmaf(__uint_as_float(v[j]), c, -mc));
    const float p1 = exp2f(fmaf(__uint_as_float(v[j + 1]), c, -mc));
    s0 += p0; s1 += p1;
    pk[j >> 1] = pack_bf16x2(p0, p1);
  }
  return s0 + s1;
}

// Exact two-pass attention of ONE query row on the CUDA cores, straight from global memory: the cold path for rows whose
// single-pass sum overflowed.  Thread `hh` of the row's pair produces output dims [32 hh, 32 hh + 32) into the staging
// row.  Scalar arguments only: a reference to the kernel's parameter struct would force a local-memory copy of it (and an
// LDL for every p.x in the hot loops).
struct ExactRowDims { int T, N0, G, H, BT, D, groups; };
__device__ __noinline__ void attn_row_exact(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ kvaux,
                                            uint8_t* stage_row, float* __restrict__ lse, ExactRowDims dm, int f, int h,
                                            int qrow, int hh) {
  const float c = 0.125f * 1.4426950408889634f;
  const int D = dm.D, b = f / dm.T;
  const int S = dm.N0 + dm.G + dm.T + 1;
  auto key_ptr = [&](int i) -> const __nv_bfloat16* {   // K of key i (V sits D elements further)
    if (i < dm.N0) return qkv + ((long long)f * dm.N0 + i) * 3 * D + D + h * 64;
    i -= dm.N0;
    long long r;
    if (i < dm.G) r = i;
    else if (i < dm.G + dm.T) r = dm.G + (long long)b * dm.T + (i - dm.G);
    else r = dm.G + (long long)dm.BT + f;
    return kvaux + r * 2 * D + h * 64;
  };
  const __nv_bfloat16* qp = qkv + ((long long)f * dm.N0 + qrow) * 3 * D + h * 64;
  auto score = [&](int i) {
    if (i >= dm.N0 + dm.G && ((i < dm.N0 + dm.G + dm.T) ? !(dm.groups & 1) : !(dm.groups & 2))) return -INFINITY;   // disabled group
    const __nv_bfloat16* kp = key_ptr(i);
    float s = 0.f;
#pragma unroll 1
    for (int d = 0; d < 64; ++d) s = fmaf(__bfloat162float(qp[d]), __bfloat162float(kp[d]), s);
    return s;
  };
  float mx = -INFINITY;
#pragma unroll 1
  for (int i = 0; i < S; ++i) mx = fmaxf(mx, score(i));
  const float mc = mx * c;
  float tot = 0.f;
#pragma unroll 1
  for (int i = 0; i < S; ++i) tot += exp2f(fmaf(score(i), c, -mc));
  const float inv = 1.0f / tot;
  const int sw = qrow & 7;            // 128-byte swizzle of the staging tile: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
#pragma unroll 1
  for (int d = 0; d < 32; ++d) {      // one output value at a time: slow and small (no register arrays)
    float acc = 0.f;
#pragma unroll 1
    for (int i = 0; i < S; ++i) {
      const float pb = __bfloat162float(__float2bfloat16_rn(exp2f(fmaf(score(i), c, -mc))));
      acc = fmaf(pb, __bfloat162float(key_ptr(i)[D + hh * 32 + d]), acc);
    }
    const int chunk = (hh * 4 + (d >> 3)) ^ sw;
    *reinterpret_cast<__nv_bfloat16*>(stage_row + chunk * 16 + (d & 7) * 2) = __float2bfloat16_rn(acc * inv);
  }
  if (lse != nullptr && hh == 0) lse[((long long)f * dm.H + h) * dm.N0 + qrow] = mc + log2f(tot);
}

// KB keys per block (S columns), NB key blocks per item, QR Q tiles in the ring
template <int KB, int NB, int QR>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    const __grid_constant__ CUtensorMap tm_aux, const __grid_constant__ CUtensorMap tm_out,
                    const AttnTcArgs p) {
  constexpr int NKT = KB * NB;                 // smem key rows per item
  constexpr int kKVBytes = NKT * 128;          // K (or V) of one item
  constexpr int kStageBytes = 2 * kKVBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem + 2 * kStageBytes;
  uint8_t* a_ext = smem_q + QR * ATC_QTILE_BYTES;      // [128 x 16] ones column (A operand of the masking k-step), 4 KB
  uint8_t* b_ext = a_ext + 4096;                       // [NKT x 16] 0 / -1e30 per key row (B operand), NKT * 32 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ext + NKT * 32);
  uint64_t* kv_full = bars;         // [2] TMA -> MMA
  uint64_t* kv_empty = bars + 2;    // [2] MMA -> TMA (n_units commits)
  uint64_t* s_full = bars + 4;      // [2] MMA -> softmax group of slot s (once per key block)
  uint64_t* p_ready = bars + 6;     // [2] softmax group -> MMA (once per key block)
  uint64_t* o_full = bars + 8;      // [2] MMA -> softmax group (once per unit)
  uint64_t* s_free = bars + 10;     // [2] softmax group -> MMA (O is in registers: the slot is reusable)
  uint64_t* pv_done = bars + 12;    // [2] only with GAVA_ATTN_SAFE_ORDER
  uint64_t* turn = bars + 14;       // [2] exp-pass token: turn[s] completes when the OTHER group finished an exp pass
  uint64_t* q_full = bars + 16;     // [QR] TMA -> MMA
  uint64_t* q_empty = bars + 16 + QR;   // [QR] softmax group (after its TMA store read the staged O tile) -> TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16 + 2 * QR);
  float* xch = reinterpret_cast<float*>(bars + 18 + 2 * QR);   // [slot 2][half 2][row 128] row statistics exchange (max, later sum)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_units;

  // K/V rows that no TMA box covers must still be finite (their probabilities are exact zeros): zero both stages once
  for (int i = threadIdx.x; i < 2 * kStageBytes / 16; i += ATC_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  {   // the masking operands
    KeyMask km0;
    km0.v0 = p.N0; km0.v1 = p.kb_live; km0.v2 = p.kb_live + 8 + ((p.groups & 1) ? p.T : 0);
    km0.v3 = (p.groups & 2) ? p.kb_live + 8 + p.t_pad : -1;
    const __nv_bfloat16 one = __float2bfloat16_rn(1.0f), zero = __float2bfloat16_rn(0.0f), neg = __float2bfloat16_rn(-1e30f);
    for (int i = threadIdx.x; i < 128 * 16; i += ATC_THREADS)
      *reinterpret_cast<__nv_bfloat16*>(a_ext + nosw_offset(i >> 4, i & 15)) = (i & 15) == 0 ? one : zero;
    for (int i = threadIdx.x; i < NKT * 16; i += ATC_THREADS)
      *reinterpret_cast<__nv_bfloat16*>(b_ext + nosw_offset(i >> 4, i & 15)) = ((i & 15) == 0 && !km0.ok(i >> 4)) ? neg : zero;
  }
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_aux);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], n_units);   // one tcgen05.commit per unit of the item
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 256);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 256);
      mbar_init(&pv_done[i], 1);
      mbar_init(&turn[i], 256);
    }
    for (int i = 0; i < QR; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
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
      int j = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++j) {
        const int f = item / p.H, h = item - f * p.H;
        const int b = f / p.T;
        const int st = j & 1;
        uint8_t* sK = smem + st * kStageBytes;
        uint8_t* sV = sK + kKVBytes;
        mbar_wait(&kv_empty[st], ((uint32_t)(j >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&kv_full[st], p.kv_tx_bytes);
        const int xk = p.D + h * 64, xv = 2 * p.D + h * 64;
        const int row = f * p.N0;
        for (int r = 0; r < p.kb_live; r += p.live_box) {
          tma_load_2d(sK + r * 128, &tm_kv, &kv_full[st], xk, row + r);
          tma_load_2d(sV + r * 128, &tm_kv, &kv_full[st], xv, row + r);
        }
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
        for (int tile = 0; tile < n_units; ++tile) {
          const int u = j * n_units + tile;
          const int qs = u % QR;
          mbar_wait(&q_empty[qs], ((uint32_t)(u / QR) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&q_full[qs], ATC_QTILE_BYTES);
          tma_load_2d(smem_q + qs * ATC_QTILE_BYTES, &tm_q, &q_full[qs], h * 64, row + tile * 128);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: warp 1 drives slot 0, warp 2 slot 1 =====================
    // One in-order issuer for both slots would couple them (S of the next unit could not be issued while the issuer
    // waits for the other slot's probabilities); with one issuer per slot the two softmax groups run decoupled.
    const int s = warp - 1;
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KB);
      constexpr uint32_t idesc_o = make_idesc_bf16_bmn(128, 64);
      const uint32_t t_slot = tmem_base + (uint32_t)(s * 256);
      uint32_t n_unit = 0, n_blk = 0;
#if GAVA_ATTN_DEBUG
      long long d_kv = 0, d_sfree = 0, d_pready = 0, d_start = clock64();
#endif
      int j = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++j) {
        const int st = j & 1;
        uint8_t* sK = smem + st * kStageBytes;
        uint8_t* sV = sK + kKVBytes;
        for (int tile = 0; tile < n_units; ++tile) {
          const int u = j * n_units + tile;
          if ((u & 1) != s) continue;
          const int qs = u % QR;
          { ATC_T0(); mbar_wait(&kv_full[st], (uint32_t)(j >> 1) & 1u); mbar_wait(&q_full[qs], (uint32_t)(u / QR) & 1u); ATC_ACC(d_kv); }
          { ATC_T0(); mbar_wait(&s_free[s], (n_unit & 1u) ^ 1u); ATC_ACC(d_sfree); }
          tc_fence_after();
          const uint64_t dq = make_sw128_kmajor_desc(smem_u32(smem_q + qs * ATC_QTILE_BYTES));
          const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK));
          const uint64_t dv = make_sw128_mnmajor_desc(smem_u32(sV), (uint32_t)kKVBytes);
          const uint64_t dae = make_nosw_kmajor_desc(smem_u32(a_ext));
          const uint64_t dbe = make_nosw_kmajor_desc(smem_u32(b_ext));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(t_slot, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
          umma_bf16_ss(t_slot, dae, dbe, idesc_s, 1u);       // + (-1e30) on the invalid key columns
          umma_commit(&s_full[s]);
#pragma unroll
          for (int blk = 0; blk < NB; ++blk) {
            { ATC_T0(); mbar_wait(&p_ready[s], n_blk & 1u); ATC_ACC(d_pready); }
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < KB / 16; ++k) {  // 16 keys per MMA; V advances 16 rows (2048 B); P of key half 0 sits
              // at columns 8k, P of key half 1 at KB/2 + 8(k - KB/32) (each half writes P inside its own S columns)
              const int pcol = k < KB / 32 ? 8 * k : KB / 2 + 8 * (k - KB / 32);
              umma_bf16_ts(t_slot + 192u, t_slot + (uint32_t)pcol, dv + (uint64_t)(128 * (blk * (KB / 16) + k)), idesc_o,
                           (uint32_t)(blk != 0 || k != 0));
            }
            if (blk + 1 < NB) {
#if GAVA_ATTN_SAFE_ORDER
              umma_commit(&pv_done[s]);
              mbar_wait(&pv_done[s], (n_unit * (uint32_t)(NB - 1) + (uint32_t)blk) & 1u);   // one completion per non-final block
              tc_fence_after();
#endif
              // S of the next key block overwrites P: tcgen05.mma instructions of one thread execute in issue order
              const uint64_t dkb = dk + (uint64_t)((blk + 1) * KB * 128 / 16);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss(t_slot, dq + (uint64_t)(2 * k), dkb + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
              umma_bf16_ss(t_slot, dae, dbe + (uint64_t)((blk + 1) * KB * 32 / 16), idesc_s, 1u);
              umma_commit(&s_full[s]);
            }
            ++n_blk;
          }
          umma_commit(&o_full[s]);
          umma_commit(&kv_empty[st]);   // one arrival per unit: the stage is free once all units' MMAs have retired
          ++n_unit;                     // (the Q tile is released by the softmax group: it stages the unit's output there)
        }
      }
#if GAVA_ATTN_DEBUG
      if (blockIdx.x == 3) printf("[attn mma %d] total %lld: wait kv/q %lld, s_free %lld, p_ready %lld\n", s, clock64() - d_start, d_kv, d_sfree, d_pready);
#endif
    }
    __syncwarp();
  } else if (warp >= 3) {
    // ===================== softmax + output: 2 threads per query row =====================
    const int idx = warp - 3;               // 0..15
    const int s = idx >> 3;                 // TMEM slot
    const int hh = (idx >> 2) & 1;          // key-column half owned by this thread
    const int quarter = warp & 3;           // TMEM lane quarter (hardware rule: warp % 4)
    const int row_in_tile = quarter * 32 + lane;
    constexpr int HALF = KB / 2;            // keys per half and block: HALF/16 chunks of 16
    constexpr int NCH = HALF / 16;
    const uint32_t t_slot = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 256);
    const uint32_t t_mine = t_slot + (uint32_t)(hh * HALF);
    const float c = 0.125f * 1.4426950408889634f;  // dh^-0.5 * log2(e)
    KeyMask km;
    km.v0 = p.N0; km.v1 = p.kb_live; km.v2 = p.kb_live + 8 + ((p.groups & 1) ? p.T : 0);
    km.v3 = (p.groups & 2) ? p.kb_live + 8 + p.t_pad : km.v2 - 1;      // here: the LAST valid smem key row (bounds the exp pass)
    // one slot per (group, half, row): holds the first-chunk maximum until the named barrier, the row sum afterwards
    float* my_max = xch + (s * 2 + hh) * 128 + row_in_tile;
    float* ot_max = xch + (s * 2 + (hh ^ 1)) * 128 + row_in_tile;
    float* my_sum = my_max;
    float* ot_sum = ot_max;
    const bool storer = quarter == 0 && hh == 0 && lane == 0;   // issues the unit's TMA store (quarter 0 is never dead)
    int pending_qs = -1;                    // Q tile whose TMA store (by this thread) may still be reading shared memory
    uint32_t n_unit = 0, n_blk = 0;
    // key blocks the OTHER group (slot 1) will run in this CTA: group 0 owns the odd unit out and must not wait for a
    // token that is never handed over
    const int my_items = (p.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t blocks_b = (uint32_t)((my_items * n_units) / 2) * NB;
#if GAVA_ATTN_DEBUG
    long long d_sfull = 0, d_bar = 0, d_ofull = 0, d_exp = 0, d_epi = 0, d_turn = 0, d_start = clock64();
#endif
    int j = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++j) {
      const int f = item / p.H, h = item - f * p.H;
      for (int tile = 0; tile < n_units; ++tile) {
        if (((j * n_units + tile) & 1) != s) continue;
        if (tile * 128 + quarter * 32 >= p.n_q) {
          // All 32 query rows of this warp lie beyond n_q (e.g. rows 224..255 of a 197-row frame): nothing to compute.
          // MMA rows are independent, so the garbage left in these P/O rows never reaches a stored row; the warp only
          // keeps the barrier protocol going (named barriers with its partner warps, token, p_ready, s_free).
          for (int blk = 0; blk < NB; ++blk) {
            mbar_wait(&s_full[s], n_blk & 1u);
            if (blk == 0) asm volatile("bar.sync %0, 256;" ::"r"(1 + s) : "memory");
            if (s == 0) { if (n_blk > 0 && n_blk - 1u < blocks_b) mbar_wait(&turn[0], (n_blk - 1u) & 1u); }
            else mbar_wait(&turn[1], n_blk & 1u);
            ++n_blk;
            mbar_arrive(&turn[s ^ 1]);
            mbar_arrive(&p_ready[s]);
          }
          mbar_wait(&o_full[s], n_unit & 1u);
          mbar_arrive(&s_free[s]);
          ++n_unit;
          asm volatile("bar.sync %0, 256;" ::"r"(3 + s) : "memory");
          continue;
        }
        float mc = 0.f, sum = 0.f;
#pragma unroll
        for (int blk = 0; blk < NB; ++blk) {
          { ATC_T0(); mbar_wait(&s_full[s], n_blk & 1u); ATC_ACC(d_sfull); }
          ++n_blk;                                   // this is key block n_blk - 1 of the slot
          tc_fence_after();
          if (blk == 0 && storer && pending_qs >= 0) {
            // The previous unit's TMA store has finished reading its staging tile by now (a whole S MMA round trip
            // later): hand that Q tile back to the producer.  Not later than here: with two key blocks the other
            // group's next unit may need this very ring slot before it can hand over the token this unit's second
            // block waits for.
            tma_store_wait_read();
            mbar_arrive(&q_empty[pending_qs]);
            pending_qs = -1;
          }
          uint32_t va[16], vb[16];
          tmem_ld_32x16(t_mine, va);
          const int kbase = blk * KB + hh * HALF;    // smem key row of this thread's first column
          if (blk == 0) {
            // the row's shift: maximum of the first 16 scores of each key half (key 0, the cls token, is always valid)
            tmem_ld_wait();
            float mx = chunk16_max(va);
            *my_max = mx;
            { ATC_T0(); asm volatile("bar.sync %0, 256;" ::"r"(1 + s) : "memory"); ATC_ACC(d_bar); }
            mx = fmaxf(mx, *ot_max);
            mc = mx * c;
          }
          // ---- this group's turn on the MUFU pipe
          { ATC_T0();
            if (s == 0) { if (n_blk > 1 && n_blk - 2u < blocks_b) mbar_wait(&turn[0], n_blk & 1u); }   // block n_blk - 2 of group 1
            else mbar_wait(&turn[1], (n_blk - 1u) & 1u);                                             // block n_blk - 1 of group 0
            ATC_ACC(d_turn); }
          // ---- p = exp2(s*c - m*c); P (bf16) overwrites S inside this thread's own columns: chunk ch (16 keys) is read
          // from columns [16 ch, 16 ch + 16) of the own half and written back as 8 columns at [8 ch, 8 ch + 8), which were
          // consumed already.  A rolled loop on purpose: fully unrolled (7-8 chunks x {unmasked, masked, dead} variants)
          // the pass was ~30 KB of code and stalled on instruction fetch (ncu: no_instruction was its top stall reason).
#if GAVA_ATTN_DEBUG
          const long long t_exp0 = clock64();
#endif
          int n_live = (km.v3 + 16 - kbase) >> 4;       // chunks that hold at least one valid key (the rest: P = 0, no exps)
          n_live = n_live < 0 ? 0 : (n_live > NCH ? NCH : n_live);
          int ch = 0;
#pragma unroll 1
          for (; ch + 2 <= n_live; ch += 2) {           // branch-free body, two chunks per trip: the TMEM load of one
            uint32_t pk[8];                             // overlaps the exps of the other
            tmem_ld_wait();
            tmem_ld_32x16(t_mine + (uint32_t)((ch + 1) * 16), vb);
            sum += chunk16_exp(va, pk, c, mc);
            tmem_st_32x8(t_mine + (uint32_t)(ch * 8), pk);
            tmem_ld_wait();
            if (ch + 2 < n_live) tmem_ld_32x16(t_mine + (uint32_t)((ch + 2) * 16), va);
            sum += chunk16_exp(vb, pk, c, mc);
            tmem_st_32x8(t_mine + (uint32_t)((ch + 1) * 8), pk);
          }
          tmem_ld_wait();
          if (ch < n_live) {                            // the odd chunk out
            uint32_t pk[8];
            sum += chunk16_exp(va, pk, c, mc);
            tmem_st_32x8(t_mine + (uint32_t)(ch * 8), pk);
            ++ch;
          }
#pragma unroll 1
          for (; ch < NCH; ++ch) {
            const uint32_t zz[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st_32x8(t_mine + (uint32_t)(ch * 8), zz);
          }
          mbar_arrive(&turn[s ^ 1]);          // exp pass done: the other group's turn
          if (blk == NB - 1) *my_sum = sum;
          tmem_st_wait();
#if GAVA_ATTN_DEBUG
          d_exp += clock64() - t_exp0;
#endif
          tc_fence_before();
          mbar_arrive(&p_ready[s]);
        }
        // ---- O = P V done: normalise, cast, store this thread's 32 of the row's 64 output values
        { ATC_T0(); mbar_wait(&o_full[s], n_unit & 1u); ATC_ACC(d_ofull); }
        ++n_unit;
        tc_fence_after();
#if GAVA_ATTN_DEBUG
        const long long t_epi0 = clock64();
#endif
        const float tot = sum + *ot_sum;            // partner's sum was written before its p_ready arrive
        const float inv = 1.0f / tot;
        const int qrow = tile * 128 + row_in_tile;
        const bool exact = !(tot < ATC_SUM_LIMIT) && qrow < p.n_q;    // overflow (or NaN): recompute the row below
        if (p.lse != nullptr && hh == 0 && qrow < p.n_q && !exact) p.lse[((long long)f * p.H + h) * p.N0 + qrow] = mc + log2f(tot);
        uint32_t o0[16], o1[16];
        tmem_ld_32x16(t_slot + 192u + (uint32_t)(hh * 32), o0);
        tmem_ld_32x16(t_slot + 192u + (uint32_t)(hh * 32 + 16), o1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[s]);                    // O is in registers: the slot's next S MMA may start
        // stage the normalised bf16 row in the unit's Q tile (its MMAs have retired), 128-byte swizzled like a TMA tile
        const int qs = (j * n_units + tile) % QR;
        uint8_t* stage = smem_q + qs * ATC_QTILE_BYTES;
        uint8_t* stage_row = stage + row_in_tile * 128;
        if (!exact) {
          const int sw = row_in_tile & 7;
#define GAVA_PK(v, i) pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv)
          st_shared_v4(stage_row + (((hh * 4 + 0) ^ sw) << 4), GAVA_PK(o0, 0), GAVA_PK(o0, 2), GAVA_PK(o0, 4), GAVA_PK(o0, 6));
          st_shared_v4(stage_row + (((hh * 4 + 1) ^ sw) << 4), GAVA_PK(o0, 8), GAVA_PK(o0, 10), GAVA_PK(o0, 12), GAVA_PK(o0, 14));
          st_shared_v4(stage_row + (((hh * 4 + 2) ^ sw) << 4), GAVA_PK(o1, 0), GAVA_PK(o1, 2), GAVA_PK(o1, 4), GAVA_PK(o1, 6));
          st_shared_v4(stage_row + (((hh * 4 + 3) ^ sw) << 4), GAVA_PK(o1, 8), GAVA_PK(o1, 10), GAVA_PK(o1, 12), GAVA_PK(o1, 14));
#undef GAVA_PK
        } else {
          ExactRowDims dm;
          dm.T = p.T; dm.N0 = p.N0; dm.G = p.G; dm.H = p.H; dm.BT = p.BT; dm.D = p.D; dm.groups = p.groups;
          attn_row_exact(p.qkv, p.kvaux, stage_row, p.lse, dm, f, h, qrow, hh);
        }
        fence_proxy_async();                        // generic-proxy smem writes -> visible to the TMA store
        asm volatile("bar.sync %0, 256;" ::"r"(3 + s) : "memory");
        if (storer) {
          tma_store_3d(&tm_out, stage, h * 64, tile * 128, f);     // rows >= n_q are clipped by the map's bounds
          tma_store_commit();
          pending_qs = qs;                          // released one unit later (or at the end), when the read is surely over
        }
#if GAVA_ATTN_DEBUG
        d_epi += clock64() - t_epi0;
#endif
      }
    }
    if (storer) {
      if (pending_qs >= 0) {
        tma_store_wait_read();
        mbar_arrive(&q_empty[pending_qs]);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all of this thread's output tiles are written
    }
#if GAVA_ATTN_DEBUG
    if (blockIdx.x == 3 && lane == 0 && (idx == 1 || idx == 5 || idx == 9))
      printf("[attn smx warp %d] total %lld: wait s_full %lld, bar %lld, o_full %lld; exp pass %lld, epilogue %lld, turn %lld\n", idx, clock64() - d_start, d_sfull, d_bar, d_ofull, d_exp, d_epi, d_turn);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int KB, int NB, int QR>
static int launch_attention_tc(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& taux,
                               const CUtensorMap& tout, const AttnTcArgs& a, cudaStream_t stream) {
  constexpr int smem_bytes = 2 * (2 * KB * NB * 128) + QR * ATC_QTILE_BYTES + 1024 + 4096 + KB * NB * 32 + (18 + 2 * QR) * 8 + 2 * 2 * 128 * 4;
  static_assert(smem_bytes <= 227 * 1024, "attention_tc: shared memory budget");
  static bool attr_set_dev[64] = {false};
  bool& attr_set = attr_set_dev[device_slot()];
  auto kern = attention_tc_kernel<KB, NB, QR>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(attention_tc): %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  const int grid = a.num_items < num_sms() ? a.num_items : num_sms();
  kern<<<grid, ATC_THREADS, smem_bytes, stream>>>(tq, tkv, taux, tout, a);
  return check_launch("attention_tc");
}

// smem key rows the tcgen05 kernel needs for a shape (0: outside its specialisation)
int attention_tc_keys(int T, int N0, int G) {
  if (G != 8) return 0;
  const int kb_live = (N0 + 7) / 8 * 8, t_pad = (T + 7) / 8 * 8;
  const int keys = kb_live + 8 + t_pad + 8;          // live | global | local | summary box
  return keys <= 320 ? keys : 0;
}

int attention_tc(const void* qkv, const void* kv_aux, void* out, float* lse, int B, int T, int N0, int G, int H, int n_q,
                 int key_groups, cudaStream_t stream) {
  const int keys = attention_tc_keys(T, N0, G);
  if (keys == 0 || n_q > 3 * 128) {
    set_error("attention_tc: shape outside the tcgen05 specialisation (G=%d T=%d N0=%d n_q=%d)", G, T, N0, n_q);
    return GAVA_ERR_SHAPE;
  }
  const int kb_live = (N0 + 7) / 8 * 8, t_pad = (T + 7) / 8 * 8;
  int g = kb_live / 8;                           // live rows per TMA box: the largest divisor of kb_live/8 with <= 256 rows
  while (g > 32 || (kb_live / 8) % g) --g;
  const int D = H * 64, BT = B * T;
  CUtensorMap tq, tkv, taux, tout;
  int rc = make_tmap_bf16_3d(&tout, out, BT, n_q, D, D, (long long)N0 * D, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tq, qkv, (long long)BT * N0, 3LL * D, 3LL * D, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tkv, qkv, (long long)BT * N0, 3LL * D, 3LL * D, 8 * g);
  if (rc) return rc;
  rc = make_tmap_bf16(&taux, kv_aux, (long long)G + 2LL * BT, 2LL * D, 2LL * D, 8);
  if (rc) return rc;
  AttnTcArgs a;
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  a.kvaux = reinterpret_cast<const __nv_bfloat16*>(kv_aux);
  a.lse = lse;
  a.T = T; a.N0 = N0; a.G = G; a.H = H; a.n_q = n_q; a.BT = BT; a.D = D;
  a.kb_live = kb_live; a.live_box = 8 * g; a.t_pad = t_pad;
  a.num_items = BT * H;
  a.groups = key_groups;
  a.n_units = (n_q + 127) / 128;
  a.kv_tx_bytes = (uint32_t)(2 * 128 * keys);
  if (keys <= 224) return launch_attention_tc<224, 1, 4>(tq, tkv, taux, tout, a, stream);
  if (keys <= 256) return launch_attention_tc<256, 1, 4>(tq, tkv, taux, tout, a, stream);
  return launch_attention_tc<160, 2, 3>(tq, tkv, taux, tout, a, stream);
}

}  // namespace gava
