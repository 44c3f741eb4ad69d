// tcgen05 / TMEM / TMA backward of the prompt-prefix attention core: the dQ half.
//
//   delta[q] = <dO[q], O[q]>                                  attention_bwd_delta_kernel (one pass over O and dO)
//   dQ       = scale * ( P o (dO V^T - delta) ) K             attention_bwd_dq_kernel    (this file)
//   dK, dV                                                    key-tile units of attention_bwd_kernel (attention_bwd.cu, mma.sync)
// with P = exp2(S c - lse) recomputed on the SM from the forward's saved log-sum-exp rows (flash-style: nothing of size
// queries x keys touches HBM).
//
// Same machine as the forward kernel (attention_tc.cu): work item = (frame, head), query rows cut into 128-row units that
// alternate between two TMEM slots, a persistent CTA with a TMA producer warp, one MMA-issuer warp per slot and two
// softmax groups of 8 warps (two threads per query row, each owns one half of a key block).  Per slot (256 columns):
//   S_blk  = Q K_blk^T   (128 x 96 fp32)  columns [0, 96)     + the masking k-step (-1e30 on invalid key rows: P = 0 there)
//   dP_blk = dO V_blk^T  (128 x 96 fp32)  columns [96, 192)
//   dS_blk = P o (dP - delta), bf16, written in place over S_blk (keys [0,48) -> cols [0,24), keys [48,96) -> cols [48,72))
//   dQ    += dS_blk K_blk (TS MMA: A = dS from TMEM, B = K MN-major in smem)  columns [192, 256)
// Key blocks of 96 (three per item up to 288 smem key rows: every ViT-B/16 shape incl. the 70-frame recipe); the scale
// dh^-0.5 is applied when dQ leaves TMEM.  dQ rows are staged (bf16, 128-byte swizzle) in the unit's Q tile and written
// with one TMA store per unit through a 3-D map over dqkv whose bounds clip the rows >= n_q.
//
// Restates autograd of Attention.forward core, training/VitaCLIP_vision_encoder_utils.py:71-77.
#include "common.cuh"
#include "kernels.h"
#include "attention_tc.cuh"

namespace gava {

constexpr int ABT_THREADS = 608;              // 19 warps, roles as in attention_tc.cu
constexpr int ABT_TILE_BYTES = 128 * 128;     // 128 rows x 64 bf16
constexpr int ABT_KB = 96, ABT_NB = 3, ABT_NKT = ABT_KB * ABT_NB;

struct AttnBwdTcArgs {
  const float* lse;       // [BT, H, N0] log2-domain log-sum-exp rows of the forward
  const float* delta;     // [BT, H, N0] <dO, O> rows
  int T, N0, G, H, n_q, BT, D;
  int kb_live, live_box, t_pad;
  int num_items, n_units;
  uint32_t kv_tx_bytes;
};

// delta[(f, h), row] = sum_d dO[row, h*64 + d] * O[row, h*64 + d]: 8 lanes per (row, head), 16 bytes each
__global__ void __launch_bounds__(256) attention_bwd_delta_kernel(const __nv_bfloat16* __restrict__ out,
                                                                   const __nv_bfloat16* __restrict__ dout,
                                                                   float* __restrict__ delta, int BT, int N0, int H, int n_q) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int sub = (int)(gid & 7);
  const long long cell = gid >> 3;                    // (frame-row, head)
  const long long rows = (long long)BT * N0;
  const bool live = cell < rows * H;
  float acc = 0.f;
  long long row = 0;
  int h = 0;
  if (live) {
    row = cell / H;
    h = (int)(cell - row * H);
    const long long off = row * (long long)H * 64 + h * 64 + sub * 8;
    const uint4 o = *reinterpret_cast<const uint4*>(out + off);
    const uint4 d = *reinterpret_cast<const uint4*>(dout + off);
    const uint32_t* ow = reinterpret_cast<const uint32_t*>(&o);
    const uint32_t* dw = reinterpret_cast<const uint32_t*>(&d);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&ow[k]);
      const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(&dw[k]);
      acc = fmaf(__bfloat162float(a.x), __bfloat162float(c.x), acc);
      acc = fmaf(__bfloat162float(a.y), __bfloat162float(c.y), acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (live && sub == 0) {
    const long long f = row / N0;
    const int r = (int)(row - f * N0);
    if (r < n_q) delta[(f * H + h) * N0 + r] = acc;
  }
}

// ds = exp2(s c - lse) * (dp - delta) for 16 scores -> 8 packed bf16 pairs
__device__ __forceinline__ void chunk16_ds(const uint32_t (&sv)[16], const uint32_t (&dv)[16], uint32_t (&pk)[8], float c,
                                           float lse, float delta) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float p0 = exp2f(fmaf(__uint_as_float(sv[j]), c, -lse));
    const float p1 = exp2f(fmaf(__uint_as_float(sv[j + 1]), c, -lse));
    pk[j >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[j]) - delta), p1 * (__uint_as_float(dv[j + 1]) - delta));
  }
}

__global__ void __launch_bounds__(ABT_THREADS, 1)
attention_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_aux,
                        const __grid_constant__ CUtensorMap tm_dq, const AttnBwdTcArgs p) {
  constexpr int KB = ABT_KB, NB = ABT_NB, NKT = ABT_NKT;
  constexpr int kKVBytes = NKT * 128;
  constexpr int kStageBytes = 2 * kKVBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_qd = smem + 2 * kStageBytes;           // [slot 2][Q tile | dO tile]
  uint8_t* a_ext = smem_qd + 4 * ABT_TILE_BYTES;       // [128 x 16] ones column
  uint8_t* b_ext = a_ext + 4096;                       // [NKT x 16] 0 / -1e30 per key row
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ext + NKT * 32);
  uint64_t* kv_full = bars;         // [2] TMA -> MMA
  uint64_t* kv_empty = bars + 2;    // [2] MMA -> TMA (n_units commits)
  uint64_t* qd_full = bars + 4;     // [2] TMA -> MMA: Q and dO tile of the slot's next unit
  uint64_t* qd_empty = bars + 6;    // [2] softmax group (its dQ store has read the staging tile) -> TMA
  uint64_t* s_full = bars + 8;      // [2] MMA -> softmax group: S_blk and dP_blk ready
  uint64_t* p_ready = bars + 10;    // [2] softmax group -> MMA: dS_blk written
  uint64_t* o_full = bars + 12;     // [2] MMA -> softmax group: dQ complete
  uint64_t* s_free = bars + 14;     // [2] softmax group -> MMA: dQ read out
  uint64_t* turn = bars + 16;       // [2] exp-pass token between the two groups
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_units;

  for (int i = threadIdx.x; i < 2 * kStageBytes / 16; i += ABT_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  {
    KeyMask km0;
    km0.v0 = p.N0; km0.v1 = p.kb_live; km0.v2 = p.kb_live + 8 + p.T; km0.v3 = p.kb_live + 8 + p.t_pad;
    const __nv_bfloat16 one = __float2bfloat16_rn(1.0f), zero = __float2bfloat16_rn(0.0f), neg = __float2bfloat16_rn(-1e30f);
    for (int i = threadIdx.x; i < 128 * 16; i += ABT_THREADS)
      *reinterpret_cast<__nv_bfloat16*>(a_ext + nosw_offset(i >> 4, i & 15)) = (i & 15) == 0 ? one : zero;
    for (int i = threadIdx.x; i < NKT * 16; i += ABT_THREADS)
      *reinterpret_cast<__nv_bfloat16*>(b_ext + nosw_offset(i >> 4, i & 15)) = ((i & 15) == 0 && !km0.ok(i >> 4)) ? neg : zero;
  }
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_kv); tma_prefetch_desc(&tm_aux);
    tma_prefetch_desc(&tm_dq);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], n_units);
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 256);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 256);
      mbar_init(&turn[i], 256);
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
      uint32_t n_slot[2] = {0u, 0u};        // units loaded so far per slot
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
        int r = p.kb_live;
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
          const int s = (j * n_units + tile) & 1;
          mbar_wait(&qd_empty[s], (n_slot[s] & 1u) ^ 1u);
          mbar_arrive_expect_tx(&qd_full[s], 2 * ABT_TILE_BYTES);
          tma_load_2d(smem_qd + s * 2 * ABT_TILE_BYTES, &tm_q, &qd_full[s], h * 64, row + tile * 128);
          tma_load_2d(smem_qd + s * 2 * ABT_TILE_BYTES + ABT_TILE_BYTES, &tm_do, &qd_full[s], h * 64, row + tile * 128);
          ++n_slot[s];
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuer of slot s =====================
    const int s = warp - 1;
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KB);
      constexpr uint32_t idesc_o = make_idesc_bf16_bmn(128, 64);
      const uint32_t t_slot = tmem_base + (uint32_t)(s * 256);
      const uint64_t dae = make_nosw_kmajor_desc(smem_u32(a_ext));
      const uint64_t dbe = make_nosw_kmajor_desc(smem_u32(b_ext));
      const uint64_t dq = make_sw128_kmajor_desc(smem_u32(smem_qd + s * 2 * ABT_TILE_BYTES));
      const uint64_t ddo = make_sw128_kmajor_desc(smem_u32(smem_qd + s * 2 * ABT_TILE_BYTES + ABT_TILE_BYTES));
      uint32_t n_unit = 0, n_blk = 0;
      int j = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++j) {
        const int st = j & 1;
        uint8_t* sK = smem + st * kStageBytes;
        const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK));
        const uint64_t dv = make_sw128_kmajor_desc(smem_u32(sK + kKVBytes));
        const uint64_t dkm = make_sw128_mnmajor_desc(smem_u32(sK), (uint32_t)kKVBytes);      // K as the MN-major B of dS K
        for (int tile = 0; tile < n_units; ++tile) {
          if (((j * n_units + tile) & 1) != s) continue;
          mbar_wait(&kv_full[st], (uint32_t)(j >> 1) & 1u);
          mbar_wait(&qd_full[s], n_unit & 1u);
          mbar_wait(&s_free[s], (n_unit & 1u) ^ 1u);
          tc_fence_after();
#pragma unroll 1
          for (int blk = 0; blk < NB; ++blk) {
            const uint64_t off = (uint64_t)(blk * KB * 128 / 16);
            // S_blk (+ masking k-step) and dP_blk; issued right behind the previous block's dS K MMAs (in order)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(t_slot, dq + (uint64_t)(2 * k), dk + off + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
            umma_bf16_ss(t_slot, dae, dbe + (uint64_t)(blk * KB * 32 / 16), idesc_s, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(t_slot + (uint32_t)KB, ddo + (uint64_t)(2 * k), dv + off + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
            umma_commit(&s_full[s]);
            mbar_wait(&p_ready[s], n_blk & 1u);
            ++n_blk;
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < KB / 16; ++k) {     // dQ += dS_blk K_blk, 16 keys per MMA
              const int pcol = k < KB / 32 ? 8 * k : KB / 2 + 8 * (k - KB / 32);
              umma_bf16_ts(t_slot + 192u, t_slot + (uint32_t)pcol, dkm + (uint64_t)(128 * (blk * (KB / 16) + k)), idesc_o,
                           (uint32_t)(blk != 0 || k != 0));
            }
          }
          umma_commit(&o_full[s]);
          umma_commit(&kv_empty[st]);
          ++n_unit;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 3) {
    // ===================== dS + output: 2 threads per query row =====================
    const int idx = warp - 3;
    const int s = idx >> 3;
    const int hh = (idx >> 2) & 1;
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    constexpr int HALF = KB / 2;            // 48 keys = 3 chunks of 16 per thread and block
    constexpr int NCH = HALF / 16;
    const uint32_t t_slot = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 256);
    const uint32_t t_mine = t_slot + (uint32_t)(hh * HALF);
    const float c = 0.125f * 1.4426950408889634f;
    const int v3 = p.kb_live + 8 + p.t_pad;
    const bool storer = quarter == 0 && hh == 0 && lane == 0;
    bool pending = false;                   // this thread's last dQ store may still be reading the slot's staging tile
    uint32_t n_unit = 0, n_blk = 0;
    const int my_items = (p.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t blocks_b = (uint32_t)((my_items * n_units) / 2) * NB;
    uint8_t* stage = smem_qd + s * 2 * ABT_TILE_BYTES;       // the slot's Q tile doubles as the dQ staging tile
    int j = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++j) {
      const int f = item / p.H, h = item - f * p.H;
      for (int tile = 0; tile < n_units; ++tile) {
        if (((j * n_units + tile) & 1) != s) continue;
        const bool dead = tile * 128 + quarter * 32 >= p.n_q;
        const int qrow = tile * 128 + row_in_tile;
        float lse = 0.f, dl = 0.f;
        if (!dead && qrow < p.n_q) {
          const long long o = ((long long)f * p.H + h) * p.N0 + qrow;
          lse = p.lse[o];
          dl = p.delta[o];
        }
#pragma unroll 1
        for (int blk = 0; blk < NB; ++blk) {
          const uint32_t g = n_blk;
          mbar_wait(&s_full[s], g & 1u);
          ++n_blk;
          tc_fence_after();
          if (s == 0) { if (g > 0 && g - 1u < blocks_b) mbar_wait(&turn[0], (g - 1u) & 1u); }
          else mbar_wait(&turn[1], g & 1u);
          if (!dead) {
            const int kbase = blk * KB + hh * HALF;
            int n_live = (v3 + 16 - kbase) >> 4;
            n_live = n_live < 0 ? 0 : (n_live > NCH ? NCH : n_live);
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
              uint32_t pk[8];
              if (ch < n_live) {
                uint32_t sv[16], dv[16];
                tmem_ld_32x16(t_mine + (uint32_t)(ch * 16), sv);
                tmem_ld_32x16(t_mine + (uint32_t)(KB + ch * 16), dv);
                tmem_ld_wait();
                chunk16_ds(sv, dv, pk, c, lse, dl);
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) pk[q] = 0u;
              }
              tmem_st_32x8(t_mine + (uint32_t)(ch * 8), pk);
            }
          }
          mbar_arrive(&turn[s ^ 1]);
          if (!dead) {
            tmem_st_wait();
            tc_fence_before();
          }
          mbar_arrive(&p_ready[s]);
        }
        mbar_wait(&o_full[s], n_unit & 1u);
        ++n_unit;
        tc_fence_after();
        if (dead) {
          mbar_arrive(&s_free[s]);
          asm volatile("bar.sync %0, 256;" ::"r"(3 + s) : "memory");
          continue;
        }
        uint32_t o0[16], o1[16];
        tmem_ld_32x16(t_slot + 192u + (uint32_t)(hh * 32), o0);
        tmem_ld_32x16(t_slot + 192u + (uint32_t)(hh * 32 + 16), o1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[s]);
        uint8_t* stage_row = stage + row_in_tile * 128;
        const int sw = row_in_tile & 7;
        const float sc = 0.125f;
#define GAVA_PK(v, i) pack_bf16x2(__uint_as_float(v[i]) * sc, __uint_as_float(v[i + 1]) * sc)
        st_shared_v4(stage_row + (((hh * 4 + 0) ^ sw) << 4), GAVA_PK(o0, 0), GAVA_PK(o0, 2), GAVA_PK(o0, 4), GAVA_PK(o0, 6));
        st_shared_v4(stage_row + (((hh * 4 + 1) ^ sw) << 4), GAVA_PK(o0, 8), GAVA_PK(o0, 10), GAVA_PK(o0, 12), GAVA_PK(o0, 14));
        st_shared_v4(stage_row + (((hh * 4 + 2) ^ sw) << 4), GAVA_PK(o1, 0), GAVA_PK(o1, 2), GAVA_PK(o1, 4), GAVA_PK(o1, 6));
        st_shared_v4(stage_row + (((hh * 4 + 3) ^ sw) << 4), GAVA_PK(o1, 8), GAVA_PK(o1, 10), GAVA_PK(o1, 12), GAVA_PK(o1, 14));
#undef GAVA_PK
        fence_proxy_async();
        asm volatile("bar.sync %0, 256;" ::"r"(3 + s) : "memory");
        if (storer) {
          tma_store_3d(&tm_dq, stage, h * 64, tile * 128, f);      // rows >= n_q are clipped by the map's bounds
          tma_store_commit();
          tma_store_wait_read();                                   // the slot's Q / dO tiles may be refilled
          mbar_arrive(&qd_empty[s]);
          pending = true;
        }
      }
    }
    if (storer && pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// smem key rows x layout the dQ kernel supports
bool attention_bwd_tc_supported(int T, int N0, int G, int n_q) {
  if (G != 8 || n_q > 3 * 128) return false;
  const int kb_live = (N0 + 7) / 8 * 8, t_pad = (T + 7) / 8 * 8;
  return kb_live + 8 + t_pad + 8 <= ABT_NKT;
}

int attention_bwd_delta(const void* out, const void* dout, float* delta, int B, int T, int N0, int H, int n_q, cudaStream_t stream) {
  const long long cells = (long long)B * T * N0 * H * 8;
  attention_bwd_delta_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), delta, B * T, N0, H, n_q);
  return check_launch("attention_bwd_delta");
}

int attention_bwd_dq_tc(const void* qkv, const void* kv_aux, const void* dout, const float* lse, const float* delta, void* dqkv,
                        int B, int T, int N0, int G, int H, int n_q, cudaStream_t stream) {
  if (!attention_bwd_tc_supported(T, N0, G, n_q)) {
    set_error("attention_bwd_dq_tc: shape outside the tcgen05 specialisation (G=%d T=%d N0=%d n_q=%d)", G, T, N0, n_q);
    return GAVA_ERR_SHAPE;
  }
  const int kb_live = (N0 + 7) / 8 * 8, t_pad = (T + 7) / 8 * 8;
  const int keys = kb_live + 8 + t_pad + 8;
  int g = kb_live / 8;
  while (g > 32 || (kb_live / 8) % g) --g;
  const int D = H * 64, BT = B * T;
  CUtensorMap tq, tdo, tkv, taux, tdq;
  int rc = make_tmap_bf16(&tq, qkv, (long long)BT * N0, 3LL * D, 3LL * D, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tdo, dout, (long long)BT * N0, D, D, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tkv, qkv, (long long)BT * N0, 3LL * D, 3LL * D, 8 * g);
  if (rc) return rc;
  rc = make_tmap_bf16(&taux, kv_aux, (long long)G + 2LL * BT, 2LL * D, 2LL * D, 8);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tdq, dqkv, BT, n_q, 3LL * D, 3LL * D, (long long)N0 * 3 * D, 128);
  if (rc) return rc;
  AttnBwdTcArgs a;
  a.lse = lse; a.delta = delta;
  a.T = T; a.N0 = N0; a.G = G; a.H = H; a.n_q = n_q; a.BT = BT; a.D = D;
  a.kb_live = kb_live; a.live_box = 8 * g; a.t_pad = t_pad;
  a.num_items = BT * H;
  a.n_units = (n_q + 127) / 128;
  a.kv_tx_bytes = (uint32_t)(2 * 128 * keys);
  constexpr int smem_bytes = 2 * (2 * ABT_NKT * 128) + 4 * ABT_TILE_BYTES + 1024 + 4096 + ABT_NKT * 32 + 20 * 8;
  static_assert(smem_bytes <= 227 * 1024, "attention_bwd_dq: shared memory budget");
  static bool attr_set_dev[64] = {false};
  bool& attr_set = attr_set_dev[device_slot()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(attention_bwd_dq): %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  const int grid = a.num_items < num_sms() ? a.num_items : num_sms();
  attention_bwd_dq_kernel<<<grid, ABT_THREADS, smem_bytes, stream>>>(tq, tdo, tkv, taux, tdq, a);
  return check_launch("attention_bwd_dq_tc");
}

}  // namespace gava
