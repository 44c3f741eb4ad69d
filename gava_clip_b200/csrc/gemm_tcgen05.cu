// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue)
//
//   A, W : bf16, K-major (row-major [rows, K]); moved by TMA (128-byte swizzle) into a 4..6 stage smem ring
//   D    : fp32 accumulators in TMEM, 2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   roles: warp 0 = TMA producer (1 lane), warp 1 = MMA issuer (1 lane), warp 2 = TMEM allocator,
//          warps 4..11 = epilogue (TMEM lane == output row; transposed through smem for coalesced global access)
//
// Replaces the nn.Linear / Conv2d(k=s=patch) library calls of the reference hot path:
//   q/k/v/out_proj  training/VitaCLIP_vision_encoder_utils.py:66,79
//   mlp fc1/fc2     training/VitaCLIP_vision_encoder_utils.py:110-115
//   cls_proj        training/VitaCLIP_vision_encoder_utils.py:164-165
//   patch_embed     training/VitaCLIP_vision_encoder_utils.py:215-219
//   final proj      training/VitaCLIP_vision_encoder.py:127
#include <memory>
#include "common.cuh"
#include "kernels.h"
#include "rowwise.cuh"

#include <mutex>
#include <stdlib.h>

#ifndef GAVA_GEMM_DEBUG
#define GAVA_GEMM_DEBUG 0
#endif

namespace gava {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle atom
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;                       // 2 warps per TMEM lane quarter, each takes half of the N columns
constexpr int GEMM_THREADS = 128 + 32 * EPI_WARPS;  // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4..: epilogue
constexpr int EPI_STAGE_BYTES = 4096;               // per-warp transpose buffer

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kBytesA = BM * BK * 2;
  static constexpr int kBytesB = BN * BK * 2;
  static constexpr int kBytesStage = kBytesA + kBytesB;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages
  static constexpr int kSmemBytes = kStages * kBytesStage + EPI_WARPS * EPI_STAGE_BYTES + 2 * BN * 4 /*bias*/ + 256 /*barriers*/;
};

struct GemmArgs {
  int M, N, K;
  const float* bias;    // [N] or null
  const float* resid;   // fp32 [M, ldr] (EPI_BIAS_RESID_F32)
  long long ldr;
  void* out;            // bf16 or fp32 [M(out rows), ldo]
  long long ldo;
  int group, group_stride, row_offset;  // EPI_PATCH_F32: out_row = (r/group)*group_stride + r%group + row_offset
  const float* pos;     // [group+1, N]
  const float* time_embed;  // [T, N]
  int T;
  const __nv_bfloat16* aux;  // EPI_DGELU_BF16: fc1 pre-activations bf16 [M, ld_aux]
  __nv_bfloat16* aux_out;    // EPI_BIAS_GELU_BF16 (optional): also store the pre-activations acc + bias here, pitch ld_aux
  long long ld_aux;
  // LayerNorm folding.  Producer (EPI_BIAS_RESID_STATS_F32): also bf16(out) -> copy_out and per-row partial
  // (sum, sum of squares) of each epilogue warp's column slice -> stats_out [M, n_slots, 2].
  __nv_bfloat16* copy_out;
  long long ld_copy;
  float* stats_out;
  // Consumer (EPI_LNFOLD_BF16): A is the raw bf16 copy, W carries the norm's weight; out = rstd_r (acc - mean_r colsum[n]) + bias[n]
  const float* stats_in;     // [rows, n_slots, 2], row pitch ld_stats floats
  long long ld_stats;
  const float* colsum;       // [N]
  int n_slots;
  float eps;
  int tma_out;               // bf16 output of modes 0 / 1 / 7 leaves through TMA stores of the swizzled staging tile
};

// ---- epilogue: TMEM -> registers (one row per thread) -> swizzled smem staging -> transposed, coalesced global access.
// After tcgen05.ld a thread owns 32 consecutive columns of ONE row, so direct global stores would touch 32 different
// rows per warp instruction (measured: l1tex 77 % busy, tensor pipe 64 %).  Each warp therefore transposes its
// 32x32 block through a private 4 KB staging buffer (XOR-swizzled 16-byte chunks: conflict-free both ways) so that
// every warp-level global load/store covers whole 64/128-byte row segments.
// (dy0, dy1) * QuickGELU'(a0, a1) on packed bf16 pairs; g'(a) = s (1 + 1.702 a (1 - s)), s = sigmoid(1.702 a)
__device__ __forceinline__ uint32_t dgelu_mul2(uint32_t dy2, uint32_t a2) {
  const __nv_bfloat162 d = *reinterpret_cast<const __nv_bfloat162*>(&dy2);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&a2);
  return pack_bf16x2(__bfloat162float(d.x) * quick_gelu_grad(__bfloat162float(a.x)),
                     __bfloat162float(d.y) * quick_gelu_grad(__bfloat162float(a.y)));
}

// Row statistics carried through one warp's share of a tile: fold = (rstd, -mean * rstd) of this thread's row and the
// staged colsum slice (consumer); rs / rq = running sum / sum of squares of the 8 rows this lane touches in the transposed
// store loop (producer).
struct FoldRow { float rstd, nmr; const float* colsum_s; };

// Epilogue output stores.  GAVA_EPI_STCS=1 (experiment build): streaming (evict-first) stores, on the theory that outputs
// far larger than L2 need not displace the operand tiles the mainloop re-reads.  Measured (tools/gemm_epi_bench.py,
// profiles/r02_gemm_epilogue_cost.txt): no change for the bf16 epilogues, the in-place fp32 stream update of fc2 615 vs
// 398 us, the whole step 3314 vs 3549 clips/s -- plain stores stay.
#ifndef GAVA_EPI_STCS
#define GAVA_EPI_STCS 0
#endif
template <typename T>
__device__ __forceinline__ void epi_store(T* ptr, const T& v) {
#if GAVA_EPI_STCS
  __stcs(ptr, v);
#else
  *ptr = v;
#endif
}

template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& p, uint8_t* stage, int row0, int col0,
                                               const uint32_t (&v)[32], const float4 (&resid)[8], const float* bias_s,
                                               int lane, const FoldRow& fold, float (&rs)[8], float (&rq)[8],
                                               const uint4 (&auxv)[4], const CUtensorMap* tmap_out, const CUtensorMap* tmap_aux, int parity) {
  if constexpr (MODE == EPI_BIAS_BF16 || MODE == EPI_BIAS_GELU_BF16 || MODE == EPI_DGELU_BF16 || MODE == EPI_LNFOLD_BF16) {
    float acc[32];
    {  // bias of this chunk from shared memory (staged per tile before the accumulator wait; broadcast LDS.128)
      const float4* b4 = reinterpret_cast<const float4*>(bias_s);
      if constexpr (MODE == EPI_LNFOLD_BF16) {
        // LayerNorm applied after the contraction: rstd (acc - mean colsum) + bias'.  With row-centred weights
        // (colsum == NULL: sum_k W'[n,k] = 0) the mean drops out inside the contraction and one FMA per element is left.
        if (fold.colsum_s != nullptr) {
          const float4* c4 = reinterpret_cast<const float4*>(fold.colsum_s);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j], c = c4[j];
            acc[4 * j + 0] = fmaf(fold.rstd, __uint_as_float(v[4 * j + 0]), fmaf(fold.nmr, c.x, b.x));
            acc[4 * j + 1] = fmaf(fold.rstd, __uint_as_float(v[4 * j + 1]), fmaf(fold.nmr, c.y, b.y));
            acc[4 * j + 2] = fmaf(fold.rstd, __uint_as_float(v[4 * j + 2]), fmaf(fold.nmr, c.z, b.z));
            acc[4 * j + 3] = fmaf(fold.rstd, __uint_as_float(v[4 * j + 3]), fmaf(fold.nmr, c.w, b.w));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            acc[4 * j + 0] = fmaf(fold.rstd, __uint_as_float(v[4 * j + 0]), b.x);
            acc[4 * j + 1] = fmaf(fold.rstd, __uint_as_float(v[4 * j + 1]), b.y);
            acc[4 * j + 2] = fmaf(fold.rstd, __uint_as_float(v[4 * j + 2]), b.z);
            acc[4 * j + 3] = fmaf(fold.rstd, __uint_as_float(v[4 * j + 3]), b.w);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = b4[j];
          acc[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x; acc[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
          acc[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z; acc[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
        }
      }
    }
    // stage as bf16 (row pitch 64 B = 4 chunks of 16 B, chunk' = chunk ^ ((row >> 1) & 3)), then store transposed
    auto stage_store = [&](const float (&a)[32], __nv_bfloat16* base, long long ld, bool dgelu, bool tma,
                           const CUtensorMap* tmap, int parity) {
      // TMA drain: the 32 x 32 bf16 block is staged exactly as below (64-byte rows, 16-byte chunk index XOR (row >> 1) & 3
      // == CU_TENSOR_MAP_SWIZZLE_64B) in one of two 2 KB slots and leaves with ONE bulk tensor store issued by lane 0 (rows
      // beyond M are clipped by the tensor map): no shared-memory read-back, no per-thread global stores.
      uint8_t* slot = stage + (tma ? parity * 2048 : 0);
      if (tma) {
        if (lane == 0) tma_store_wait_read1();     // the store that used this slot two chunks ago has read it
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 w;
        w.x = pack_bf16x2(a[8 * j + 0], a[8 * j + 1]);
        w.y = pack_bf16x2(a[8 * j + 2], a[8 * j + 3]);
        w.z = pack_bf16x2(a[8 * j + 4], a[8 * j + 5]);
        w.w = pack_bf16x2(a[8 * j + 6], a[8 * j + 7]);
        *reinterpret_cast<uint4*>(slot + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = w;
      }
      if (tma) {
        fence_proxy_async();                       // generic-proxy smem writes -> visible to the bulk store
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmap, slot, col0, row0);
          tma_store_commit();
        }
        return;
      }
      __syncwarp();
      const int c = lane & 3;
      __nv_bfloat16* obase = base + col0 + 8 * c;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + (lane >> 2);
        uint4 w = *reinterpret_cast<const uint4*>(stage + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
        if (dgelu) {   // dgrad through QuickGELU: multiply by g'(a1); a1 was prefetched as coalesced row segments
          const uint4 a1 = auxv[i];
          w.x = dgelu_mul2(w.x, a1.x); w.y = dgelu_mul2(w.y, a1.y); w.z = dgelu_mul2(w.z, a1.z); w.w = dgelu_mul2(w.w, a1.w);
        }
        if (row0 + r < p.M) epi_store(reinterpret_cast<uint4*>(obase + (long long)(row0 + r) * ld), w);
      }
      __syncwarp();
    };
    if constexpr (MODE == EPI_BIAS_GELU_BF16) {
      // training forward keeps the pre-activations: with the TMA drain the two outputs of a chunk use one slot each
      if (p.aux_out != nullptr) stage_store(acc, p.aux_out, p.ld_aux, false, p.tma_out != 0, tmap_aux, 0);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = quick_gelu(acc[j]);
    }
    stage_store(acc, reinterpret_cast<__nv_bfloat16*>(p.out), p.ldo, MODE == EPI_DGELU_BF16,
                MODE != EPI_DGELU_BF16 && p.tma_out != 0, tmap_out,
                (MODE == EPI_BIAS_GELU_BF16 && p.aux_out != nullptr) ? 1 : parity);
  } else {
    // stage raw fp32 accumulators: row pitch 128 B (8 chunks of 16 B), chunk' = chunk ^ (row & 7)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
          make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int c = lane & 7;
    const int col = col0 + 4 * c;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + (lane >> 3);
      const int grow = row0 + r;
      float4 a = *reinterpret_cast<const float4*>(stage + r * 128 + ((c ^ (r & 7)) << 4));
      a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
      if (grow < p.M) {
        if constexpr (MODE == EPI_BIAS_RESID_F32 || MODE == EPI_BIAS_RESID_STATS_F32) {
          a.x += resid[i].x; a.y += resid[i].y; a.z += resid[i].z; a.w += resid[i].w;
          epi_store(reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)grow * p.ldo + col), a);
          if constexpr (MODE == EPI_BIAS_RESID_STATS_F32) {
            // the next LayerNorm's operand (raw bf16 copy) and its row statistics, from the values still in registers
            epi_store(reinterpret_cast<uint2*>(p.copy_out + (long long)grow * p.ld_copy + col),
                      make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w)));
            rs[i] += (a.x + a.y) + (a.z + a.w);
            rq[i] = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, rq[i]))));
          }
        } else if constexpr (MODE == EPI_BIAS_F32) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)grow * p.ldo + col) = a;
        } else {  // EPI_PATCH_F32: + pos_embed[1+n] + time_embed[t], scatter to the frame's token slot
          const int grp = grow / p.group, n = grow - grp * p.group;
          const long long orow = (long long)grp * p.group_stride + n + p.row_offset;
          const float4 pe = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(n + 1) * p.N + col));
          const float4 te = __ldg(reinterpret_cast<const float4*>(p.time_embed + (long long)(grp % p.T) * p.N + col));
          a.x += pe.x + te.x; a.y += pe.y + te.y; a.z += pe.z + te.z; a.w += pe.w + te.w;
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + col) = a;
        }
      }
    }
    __syncwarp();
  }
}

// One warp's share of one accumulator tile: NCH chunks of 32 columns starting at column `col_base`, rows row0..row0+31.
// bf16-output modes: the tile's bias slice is staged in shared memory BEFORE the accumulator wait (its global-load latency
// disappears under the mainloop) and the TMEM load of chunk c+1 is in flight while chunk c is converted and stored.
template <int MODE, int NCH>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& p, uint8_t* stage, float* bias_s, uint64_t* tmem_full_bar,
                                              uint32_t full_phase, uint32_t t_addr, int row0, int col_base, int lane,
                                              long long& dbg_wait, const CUtensorMap* tmap_out, const CUtensorMap* tmap_aux) {
  FoldRow fold;
  fold.rstd = 1.f; fold.nmr = 0.f; fold.colsum_s = nullptr;
  float rs[8], rq[8];
  if constexpr (MODE == EPI_BIAS_BF16 || MODE == EPI_BIAS_GELU_BF16 || MODE == EPI_DGELU_BF16 || MODE == EPI_LNFOLD_BF16) {
    static_assert(NCH * 32 <= 128, "one float4 per lane covers the warp's columns");
    if (lane < NCH * 8) {
      const float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col_base) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(bias_s)[lane] = b;   // the 4 lane-quarter warps of a column half write identical values
    }
    if constexpr (MODE == EPI_LNFOLD_BF16) {
      // colsum slice into the unused upper half of this warp's staging buffer (bf16 staging needs 2 KB of the 4 KB);
      // this thread's row statistics from the producer's partials -- all before the accumulator wait
      if (p.colsum != nullptr) {
        float* cs = reinterpret_cast<float*>(stage + 2048);
        if (lane < NCH * 8) reinterpret_cast<float4*>(cs)[lane] = __ldg(reinterpret_cast<const float4*>(p.colsum + col_base) + lane);
        fold.colsum_s = cs;
      }
      const int grow = row0 + lane;
      if (grow < p.M) {
        const float2* st = reinterpret_cast<const float2*>(p.stats_in + (long long)grow * p.ld_stats);
        float sm = 0.f, sq = 0.f;
        for (int j = 0; j < p.n_slots; ++j) { const float2 t = __ldg(st + j); sm += t.x; sq += t.y; }
        const float inv_k = 1.0f / (float)p.K;
        const float mean = sm * inv_k;
        const float var = fmaxf(fmaf(-mean, mean, sq * inv_k), 0.f);
        fold.rstd = rsqrtf(var + p.eps);
        fold.nmr = -mean * fold.rstd;
      }
    }
    __syncwarp();
    // DGELU: the fc1 pre-activations this lane multiplies with after the smem transpose (rows 8 i + lane / 4, 8 columns
    // at 8 (lane % 4)) for all the warp's chunks, fetched before the accumulator wait: the latency hides under the
    // mainloop.  With the loads issued inside the store loop the K = 768 dgrad (fc2^T) ran at 546 TFLOP/s, epilogue-bound
    // on four serial global-load latencies per chunk; one chunk of look-ahead gave 958.
    uint4 aux_all[MODE == EPI_DGELU_BF16 ? NCH : 1][4];
    auto aux_prefetch = [&](int cc, uint4 (&dst)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int grow = row0 + 8 * i + (lane >> 2);
        dst[i] = (grow < p.M) ? __ldg(reinterpret_cast<const uint4*>(p.aux + (long long)grow * p.ld_aux + col_base + cc * 32 + 8 * (lane & 3)))
                              : make_uint4(0, 0, 0, 0);
      }
    };
    if constexpr (MODE == EPI_DGELU_BF16) {   // every chunk of the tile, all before the accumulator wait
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) aux_prefetch(cc, aux_all[cc]);
    }
#if GAVA_GEMM_DEBUG == 4
    const long long t0 = clock64();
#endif
    mbar_wait(tmem_full_bar, full_phase);
#if GAVA_GEMM_DEBUG == 4
    dbg_wait += clock64() - t0;
#endif
    tc_fence_after();
    uint32_t va[32], vb[32];
    float4 resid[8];
    tmem_ld_32x32(t_addr, va);
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) {
      tmem_ld_wait();
      uint32_t (&cur)[32] = (cc & 1) ? vb : va;
      uint32_t (&nxt)[32] = (cc & 1) ? va : vb;
      if (cc + 1 < NCH) tmem_ld_32x32(t_addr + (uint32_t)((cc + 1) * 32), nxt);
#if GAVA_GEMM_DEBUG != 1
      FoldRow fc = fold;
      if constexpr (MODE == EPI_LNFOLD_BF16) fc.colsum_s = fold.colsum_s ? fold.colsum_s + cc * 32 : nullptr;
      epilogue_chunk<MODE>(p, stage, row0, col_base + cc * 32, cur, resid, bias_s + cc * 32, lane, fc, rs, rq,
                           aux_all[MODE == EPI_DGELU_BF16 ? cc : 0], tmap_out, tmap_aux, cc & 1);
#endif
    }
  } else {
    if constexpr (MODE == EPI_BIAS_RESID_STATS_F32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
    }
#if GAVA_GEMM_DEBUG == 4
    const long long t0 = clock64();
#endif
    mbar_wait(tmem_full_bar, full_phase);
#if GAVA_GEMM_DEBUG == 4
    dbg_wait += clock64() - t0;
#endif
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < NCH; ++cc) {
      const int col0 = col_base + cc * 32;
      float4 resid[8];
      if constexpr (MODE == EPI_BIAS_RESID_F32 || MODE == EPI_BIAS_RESID_STATS_F32) {  // coalesced residual loads, in flight while TMEM is read
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int grow = row0 + 4 * i + (lane >> 3);
#if GAVA_GEMM_DEBUG == 5   // experiment: no residual loads
          resid[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#else
          resid[i] = (grow < p.M) ? *reinterpret_cast<const float4*>(p.resid + (long long)grow * p.ldr + col0 + 4 * (lane & 7))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.aux != nullptr && grow < p.M) {   // second, bf16 addend (the attention branch kept as a bf16 delta)
            const uint2 u = *reinterpret_cast<const uint2*>(p.aux + (long long)grow * p.ld_aux + col0 + 4 * (lane & 7));
            const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
            resid[i].x += __bfloat162float(lo.x); resid[i].y += __bfloat162float(lo.y);
            resid[i].z += __bfloat162float(hi.x); resid[i].w += __bfloat162float(hi.y);
          }
#endif
        }
      }
      uint32_t v[32];
      tmem_ld_32x32(t_addr + (uint32_t)(cc * 32), v);
      tmem_ld_wait();
#if GAVA_GEMM_DEBUG == 6   // experiment: residual loads but no stores
      if (v[0] == 0x7fc12345u && resid[0].x == 1.2345f) reinterpret_cast<uint32_t*>(p.out)[0] = v[3];
#elif GAVA_GEMM_DEBUG != 1
      const uint4 no_aux[4] = {};
      epilogue_chunk<MODE>(p, stage, row0, col0, v, resid, nullptr, lane, fold, rs, rq, no_aux, tmap_out, tmap_aux, 0);
#endif
    }
    if constexpr (MODE == EPI_BIAS_RESID_STATS_F32) {
      // the 8 lanes that share a row in the transposed store loop combine their partials; slot = this warp's column slice
      const int slot = col_base / (NCH * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
          rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);
          rq[i] += __shfl_xor_sync(0xffffffffu, rq[i], o);
        }
        const int grow = row0 + 4 * i + (lane >> 3);
        if ((lane & 7) == 0 && grow < p.M)
          *reinterpret_cast<float2*>(p.stats_out + ((long long)grow * p.n_slots + slot) * 2) = make_float2(rs[i], rq[i]);
      }
    }
  }
}

template <int BN, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux, const GemmArgs p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128-byte-swizzle tiles need 1024-byte alignment
  uint8_t* smem = smem_raw;
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kBytesA;
  uint8_t* smem_epi = smem + kStages * Cfg::kBytesStage;
  float* smem_bias = reinterpret_cast<float*>(smem_epi + EPI_WARPS * EPI_STAGE_BYTES);   // [2 accumulator stages][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bias + 2 * BN);
  uint64_t* full_bar = bars;                      // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;           // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;       // [2]        MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]        epilogue -> MMA
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_n = p.N / BN;
  const int num_m = (p.M + BM - 1) / BM;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 32 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kBytesStage);
          tma_load_2d(smem_a + stage * Cfg::kBytesA, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(smem_b + stage * Cfg::kBytesB, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(smem_u32(smem_a + stage * Cfg::kBytesA));
          const uint64_t db = make_sw128_kmajor_desc(smem_u32(smem_b + stage * Cfg::kBytesB));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in 16-byte units
            umma_bf16_ss(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;
    const int quarter = ew & 3;   // == warp % 4: the TMEM lane quarter this warp may access
    const int half = ew >> 2;     // which half of the tile's columns
    constexpr int kChunksPerWarp = (BN / 32) / (EPI_WARPS / 4);
    uint8_t* stage = smem_epi + ew * EPI_STAGE_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long g_dbg_wait = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      const int row0 = m_blk * BM + quarter * 32;
      const int col_base = n_blk * BN + half * kChunksPerWarp * 32;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * kChunksPerWarp * 32);
      epilogue_tile<MODE, kChunksPerWarp>(p, stage, smem_bias + acc * BN + half * kChunksPerWarp * 32, &tmem_full[acc],
                                          acc_phase, t_addr, row0, col_base, lane, g_dbg_wait, &tmap_out, &tmap_aux);
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (p.tma_out && lane == 0) tma_store_wait_all();   // bulk stores of the staging slots have completed
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------- CTA-pair (cta_group::2) kernel
// One 256x256 output tile per 2-CTA cluster: CTA r loads A rows [m0 + 128 r, +128) and W rows [n0 + 128 r, +128)
// (half of the B operand each), the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) and each CTA drains the
// 128 accumulator rows that live in its own TMEM.  Halving the per-SM B traffic turns the 4-stage ring into a 6-stage
// ring in the same shared memory (more TMA latency cover) and halves the operand read bandwidth per SM.
struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr int kStages = 6;
  static constexpr int kBytesA = BM * BK * 2;        // 128 rows of A
  static constexpr int kBytesB = (BN / 2) * BK * 2;  // 128 rows of W (this CTA's half of N)
  static constexpr int kBytesStage = kBytesA + kBytesB;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kBytesStage + EPI_WARPS * EPI_STAGE_BYTES + 2 * BN * 4 /*bias*/ + 256 /*barriers*/;
};

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux, const GemmArgs p) {
  using Cfg = Gemm2Cfg;
  constexpr int kStages = Cfg::kStages;
  constexpr int BN = Cfg::BN;
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128-byte-swizzle tiles need 1024-byte alignment
  uint8_t* smem = smem_raw;
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kBytesA;
  uint8_t* smem_epi = smem + kStages * Cfg::kBytesStage;
  float* smem_bias = reinterpret_cast<float*>(smem_epi + EPI_WARPS * EPI_STAGE_BYTES);   // [2 accumulator stages][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bias + 2 * BN);
  uint64_t* full_bar = bars;                      // [kStages]  TMA (both CTAs) -> MMA; only the leader's copy is used
  uint64_t* empty_bar = bars + kStages;           // [kStages]  MMA -> TMA; multicast to both CTAs
  uint64_t* tmem_full = bars + 2 * kStages;       // [2]        MMA -> epilogue; multicast to both CTAs
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]        epilogue (both CTAs) -> MMA; only the leader's copy is used
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_n = p.N / BN;
  const int num_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_tiles = num_m * num_n;
  const int num_kb = p.K / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  // Work enumeration of this cluster: iteration `it` -> tile cluster_id + it * num_clusters (round-robin over the tiles)
  auto tile_at = [&](int it, int& m_blk, int& n_blk) -> bool {
    const int t = cluster_id + it * num_clusters;
    if (t >= num_tiles) return false;
    m_blk = t / num_n;
    n_blk = t - m_blk * num_n;
    return true;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * EPI_WARPS);  // one arrive per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_base_slot, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();  // peer barriers initialised, both TMEM allocations done
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int m_blk, n_blk;
      for (int it = 0; tile_at(it, m_blk, n_blk); ++it) {
        const int row_a = m_blk * 2 * BM + (int)rank * BM;
        const int row_b = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
#if GAVA_GEMM_DEBUG == 2   // experiment: no operand loads (mainloop on stale smem)
          if (leader) mbar_arrive(&full_bar[stage]);
#else
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kBytesStage);  // bytes of BOTH CTAs
          tma_load_2d_2sm(smem_a + stage * Cfg::kBytesA, &tmap_a, &full_bar[stage], kb * BK, row_a);
          tma_load_2d_2sm(smem_b + stage * Cfg::kBytesB, &tmap_b, &full_bar[stage], kb * BK, row_b);
#endif
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
#if GAVA_GEMM_DEBUG == 4
      long long w_empty = 0, w_full = 0, t_start = clock64();
#endif
      int m_blk_, n_blk_;
      for (int it = 0; tile_at(it, m_blk_, n_blk_); ++it) {
#if GAVA_GEMM_DEBUG == 4
        long long t0 = clock64();
#endif
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
#if GAVA_GEMM_DEBUG == 4
        w_empty += clock64() - t0;
#endif
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
#if GAVA_GEMM_DEBUG == 4
          long long t1 = clock64();
#endif
          mbar_wait(&full_bar[stage], phase);
#if GAVA_GEMM_DEBUG == 4
          w_full += clock64() - t1;
#endif
          tc_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(smem_u32(smem_a + stage * Cfg::kBytesA));
          const uint64_t db = make_sw128_kmajor_desc(smem_u32(smem_b + stage * Cfg::kBytesB));
#if GAVA_GEMM_DEBUG != 3   // experiment 3: no MMAs (TMA + epilogue only)
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_ss_2sm(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
#endif
          umma_commit_2sm(&empty_bar[stage], 0b11);  // slot free in both CTAs
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&tmem_full[acc], 0b11);  // accumulator halves ready in both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
#if GAVA_GEMM_DEBUG == 4
      if (blockIdx.x == 0 || blockIdx.x == 64)
        printf("[mma  cta %d] total %lld cyc, wait tmem_empty %lld, wait full %lld\n", blockIdx.x, clock64() - t_start, w_empty, w_full);
#endif
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own TMEM half) =====================
    const int ew = warp - 4;
    const int quarter = ew & 3;
    const int half = ew >> 2;
    constexpr int kChunksPerWarp = (BN / 32) / (EPI_WARPS / 4);
    uint8_t* stage = smem_epi + ew * EPI_STAGE_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
#if GAVA_GEMM_DEBUG == 4
    long long e_start = clock64();
#endif
    long long g_dbg_wait = 0;
    int m_blk, n_blk;
    for (int it = 0; tile_at(it, m_blk, n_blk); ++it) {
      const int row0 = m_blk * 2 * BM + (int)rank * BM + quarter * 32;
      const int col_base = n_blk * BN + half * kChunksPerWarp * 32;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * kChunksPerWarp * 32);
      epilogue_tile<MODE, kChunksPerWarp>(p, stage, smem_bias + acc * BN + half * kChunksPerWarp * 32, &tmem_full[acc],
                                          acc_phase, t_addr, row0, col_base, lane, g_dbg_wait, &tmap_out, &tmap_aux);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);  // tell the leader this warp drained its part
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (p.tma_out && lane == 0) tma_store_wait_all();   // bulk stores of the staging slots have completed
#if GAVA_GEMM_DEBUG == 4
    if ((blockIdx.x == 0 || blockIdx.x == 65) && lane == 0 && (ew == 0 || ew == 7))
      printf("[epi  cta %d warp %d] total %lld cyc, waiting tmem_full %lld\n", blockIdx.x, ew, clock64() - e_start, g_dbg_wait);
#endif
  }

  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still touch this CTA's smem/TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

static int encode_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows);
static int encode_tmap_bf16_3d(CUtensorMap* m, const void* ptr, long long planes, long long rows, long long cols, long long ld,
                               long long ld_plane, int box_rows);

// Encoded descriptors are a pure function of (address, shape, pitch, box): the eager fine-tune step issues ~700 GEMM /
// attention launches over a few hundred distinct operands per step, so each host thread keeps a small direct-mapped cache
// instead of calling the driver's encoder two or three times per launch.
struct TmapKey {
  const void* ptr;
  long long rows, cols, ld, plane;   // plane: planes << 40 | ld_plane for the 3-D form, -1 for 2-D
  int box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && plane == o.plane && box_rows == o.box_rows;
  }
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool valid; };
constexpr int kTmapSlots = 1024;
static TmapSlot* tmap_slot(const TmapKey& k) {
  static thread_local std::unique_ptr<TmapSlot[]> slots;     // heap, not TLS: 1024 x 176 bytes per calling thread
  if (!slots) slots.reset(new TmapSlot[kTmapSlots]());
  uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
  h ^= (uint64_t)k.rows * 0xC2B2AE3D27D4EB4Full + (uint64_t)k.cols * 0x165667B19E3779F9ull + (uint64_t)k.ld * 31 +
       (uint64_t)k.plane * 131 + (uint64_t)k.box_rows;
  return &slots[(h >> 40) & (kTmapSlots - 1)];
}

// 2-D bf16 row-major [rows, cols] with row pitch ld (elements); box = [box_rows, 64 cols], 128B swizzle
int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  const TmapKey key{ptr, rows, cols, ld, -1, box_rows};
  TmapSlot* slot = tmap_slot(key);
  if (slot->valid && slot->key == key) { *m = slot->map; return GAVA_OK; }
  const int rc = encode_tmap_bf16(m, ptr, rows, cols, ld, box_rows);
  if (rc == GAVA_OK) { slot->key = key; slot->map = *m; slot->valid = true; }
  return rc;
}

static int encode_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return GAVA_ERR_DRIVER; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || ((ld * 2) & 15)) {
    set_error("TMA operand must be 16-byte aligned (ptr %p, ld %lld)", ptr, ld);
    return GAVA_ERR_ARG;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return GAVA_ERR_DRIVER; }
  return GAVA_OK;
}

// 2-D bf16 output map for the epilogue's TMA drain: box = [32 rows, 32 cols] (64-byte rows), 64-byte swizzle
int make_tmap_bf16_out(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld) {
  const TmapKey key{ptr, rows, cols, ld, -2, 32};
  TmapSlot* slot = tmap_slot(key);
  if (slot->valid && slot->key == key) { *m = slot->map; return GAVA_OK; }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return GAVA_ERR_DRIVER; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || ((ld * 2) & 15)) {
    set_error("TMA output must be 16-byte aligned (ptr %p, ld %lld)", ptr, ld);
    return GAVA_ERR_ARG;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out) failed (%d)", (int)r); return GAVA_ERR_DRIVER; }
  slot->key = key; slot->map = *m; slot->valid = true;
  return GAVA_OK;
}

// 3-D bf16 [planes, rows, cols] (row pitch ld, plane pitch ld_plane elements); box = [1, box_rows, 64 cols], 128B swizzle.
// Used for TMA STORES of per-frame tiles: rows beyond `rows` are clipped instead of spilling into the next plane.
int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, long long planes, long long rows, long long cols, long long ld,
                      long long ld_plane, int box_rows) {
  const TmapKey key{ptr, rows, cols, ld, (planes << 40) | ld_plane, box_rows};
  TmapSlot* slot = tmap_slot(key);
  if (slot->valid && slot->key == key) { *m = slot->map; return GAVA_OK; }
  const int rc = encode_tmap_bf16_3d(m, ptr, planes, rows, cols, ld, ld_plane, box_rows);
  if (rc == GAVA_OK) { slot->key = key; slot->map = *m; slot->valid = true; }
  return rc;
}

static int encode_tmap_bf16_3d(CUtensorMap* m, const void* ptr, long long planes, long long rows, long long cols, long long ld,
                               long long ld_plane, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return GAVA_ERR_DRIVER; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || ((ld * 2) & 15) || ((ld_plane * 2) & 15)) {
    set_error("TMA operand must be 16-byte aligned (ptr %p, ld %lld, plane %lld)", ptr, ld, ld_plane);
    return GAVA_ERR_ARG;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld_plane * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(3d) failed (%d)", (int)r); return GAVA_ERR_DRIVER; }
  return GAVA_OK;
}

int device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}

int num_sms() {
  static int n[64] = {0};
  const int d = device_slot();
  if (n[d] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d);
    n[d] = v > 0 ? v : 148;
  }
  return n[d];
}

template <int BN, int MODE>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& taux, const GemmArgs& args, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set_dev[64] = {false};
  bool& attr_set = attr_set_dev[device_slot()];
  auto kern = gemm_tcgen05_kernel<BN, MODE>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  const int tiles = ((args.M + BM - 1) / BM) * (args.N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(ta, tb, tout, taux, args);
  return check_launch("gemm_tcgen05");
}

template <int MODE>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& taux, const GemmArgs& args, cudaStream_t stream, int max_ctas = 0) {
  using Cfg = Gemm2Cfg;
  static bool attr_set_dev[64] = {false};
  bool& attr_set = attr_set_dev[device_slot()];
  auto kern = gemm2_tcgen05_kernel<MODE>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(gemm2): %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  const int tiles = ((args.M + 2 * BM - 1) / (2 * BM)) * (args.N / Cfg::BN);
  int max_clusters = num_sms() / 2;
  if (max_ctas > 1 && max_ctas / 2 < max_clusters) max_clusters = max_ctas / 2;   // leave SMs to a concurrent stream
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  kern<<<2 * clusters, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(ta, tb, tout, taux, args);
  return check_launch("gemm2_tcgen05");
}

static int dispatch_mode2(int mode, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx, const GemmArgs& a, cudaStream_t s, int max_ctas) {
  switch (mode) {
    case EPI_BIAS_BF16: return launch_gemm2<EPI_BIAS_BF16>(ta, tb, to, tx, a, s, max_ctas);
    case EPI_BIAS_GELU_BF16: return launch_gemm2<EPI_BIAS_GELU_BF16>(ta, tb, to, tx, a, s);
    case EPI_BIAS_RESID_F32: return launch_gemm2<EPI_BIAS_RESID_F32>(ta, tb, to, tx, a, s);
    case EPI_BIAS_F32: return launch_gemm2<EPI_BIAS_F32>(ta, tb, to, tx, a, s);
    case EPI_PATCH_F32: return launch_gemm2<EPI_PATCH_F32>(ta, tb, to, tx, a, s);
    case EPI_DGELU_BF16: return launch_gemm2<EPI_DGELU_BF16>(ta, tb, to, tx, a, s);
    case EPI_LNFOLD_BF16: return launch_gemm2<EPI_LNFOLD_BF16>(ta, tb, to, tx, a, s, max_ctas);
    case EPI_BIAS_RESID_STATS_F32: return launch_gemm2<EPI_BIAS_RESID_STATS_F32>(ta, tb, to, tx, a, s);
  }
  set_error("gemm: unknown epilogue mode %d", mode);
  return GAVA_ERR_ARG;
}

template <int BN>
static int dispatch_mode(int mode, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx, const GemmArgs& a, cudaStream_t s) {
  switch (mode) {
    case EPI_BIAS_BF16: return launch_gemm<BN, EPI_BIAS_BF16>(ta, tb, to, tx, a, s);
    case EPI_BIAS_GELU_BF16: return launch_gemm<BN, EPI_BIAS_GELU_BF16>(ta, tb, to, tx, a, s);
    case EPI_BIAS_RESID_F32: return launch_gemm<BN, EPI_BIAS_RESID_F32>(ta, tb, to, tx, a, s);
    case EPI_BIAS_F32: return launch_gemm<BN, EPI_BIAS_F32>(ta, tb, to, tx, a, s);
    case EPI_PATCH_F32: return launch_gemm<BN, EPI_PATCH_F32>(ta, tb, to, tx, a, s);
    case EPI_DGELU_BF16: return launch_gemm<BN, EPI_DGELU_BF16>(ta, tb, to, tx, a, s);
    case EPI_LNFOLD_BF16: return launch_gemm<BN, EPI_LNFOLD_BF16>(ta, tb, to, tx, a, s);
    case EPI_BIAS_RESID_STATS_F32: return launch_gemm<BN, EPI_BIAS_RESID_STATS_F32>(ta, tb, to, tx, a, s);
  }
  set_error("gemm: unknown epilogue mode %d", mode);
  return GAVA_ERR_ARG;
}

static long long gemm_small_m() {
  static const long long v = [] { const char* e = getenv("GAVA_SMALL_M"); return e ? atoll(e) : 2048LL; }();
  return v;
}

// Column slices per row of the statistics a EPI_BIAS_RESID_STATS_F32 launch of this shape writes: one per epilogue warp
// column share (128 columns with the 256-wide tiles, 64 with the 128-wide ones the small-M shapes use).
int gemm_stats_slots(long long M, long long N) {
  if (M <= 0 || N <= 0 || (N % 128) != 0) return 0;
  const bool wide = (N % 256 == 0) && M > gemm_small_m();
  return (int)(N / (wide ? 128 : 64));
}

int gemm_bf16(const void* A, long long lda, const void* W, long long ldw, long long M, long long N, long long K,
              const gava_gemm_epilogue* epi, cudaStream_t stream) {
  if (!A || !W || !epi || !epi->out) { set_error("gemm: null pointer"); return GAVA_ERR_ARG; }
  if (M <= 0 || N <= 0 || K <= 0 || (K % BK) != 0 || (N % 128) != 0) {
    set_error("gemm: unsupported shape M=%lld N=%lld K=%lld (need K%%64==0, N%%128==0)", M, N, K);
    return GAVA_ERR_SHAPE;
  }
  if ((epi->mode == EPI_BIAS_RESID_F32 || epi->mode == EPI_BIAS_RESID_STATS_F32) && !epi->resid) {
    set_error("gemm: residual epilogue without resid");
    return GAVA_ERR_ARG;
  }
  if (epi->mode == EPI_BIAS_RESID_STATS_F32 &&
      (!epi->copy_out || !epi->stats_out || (epi->ld_copy & 3) || epi->n_slots != gemm_stats_slots(M, N))) {
    set_error("gemm: stats epilogue needs copy_out (pitch %% 4 == 0), stats_out and n_slots == gava_gemm_stats_slots(M, N)");
    return GAVA_ERR_ARG;
  }
  if (epi->mode == EPI_LNFOLD_BF16 && (!epi->stats_in || epi->n_slots <= 0 || epi->ld_stats < 2LL * epi->n_slots ||
                                       (epi->ld_stats & 1) || !(epi->eps > 0.f))) {
    set_error("gemm: LayerNorm-fold epilogue needs stats_in (pitch >= 2 n_slots, even), n_slots > 0, eps > 0");
    return GAVA_ERR_ARG;
  }
  if (epi->mode == EPI_PATCH_F32 && (!epi->pos || !epi->time_embed || epi->group <= 0 || epi->T <= 0)) {
    set_error("gemm: patch epilogue needs pos/time_embed/group/T");
    return GAVA_ERR_ARG;
  }
  if (epi->mode == EPI_DGELU_BF16 && (!epi->aux || (epi->ld_aux & 7))) { set_error("gemm: dgelu epilogue needs aux (pitch %% 8 == 0)"); return GAVA_ERR_ARG; }
  const int bn = (N % 256 == 0) ? 256 : 128;
  static const bool use_pair = [] { const char* e = getenv("GAVA_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  // CTA-pair kernel for every N % 256 == 0 shape; the handful-of-tiles GEMMs of the temporal-message chain (M = B*T rows)
  // are latency-bound and spread over more SMs with the single-CTA 128x128 tiles (GAVA_SMALL_M: row threshold)
  const bool small = M <= gemm_small_m();
  const bool pair = use_pair && bn == 256 && !small;
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16(&ta, A, M, K, lda, BM);
  if (rc) return rc;
  const int bn_eff = small ? 128 : bn;
  rc = make_tmap_bf16(&tb, W, N, K, ldw, pair ? 128 : bn_eff);
  if (rc) return rc;
  GemmArgs a;
  a.M = (int)M; a.N = (int)N; a.K = (int)K;
  a.bias = epi->bias; a.resid = epi->resid; a.ldr = epi->ldr;
  a.out = epi->out; a.ldo = epi->ldo;
  a.group = epi->group; a.group_stride = epi->group_stride; a.row_offset = epi->row_offset;
  a.pos = epi->pos; a.time_embed = epi->time_embed; a.T = epi->T;
  a.aux = (epi->mode == EPI_DGELU_BF16 || epi->mode == EPI_BIAS_RESID_F32 || epi->mode == EPI_BIAS_RESID_STATS_F32)
              ? reinterpret_cast<const __nv_bfloat16*>(epi->aux) : nullptr;
  a.copy_out = reinterpret_cast<__nv_bfloat16*>(epi->copy_out); a.ld_copy = epi->ld_copy;
  a.stats_out = epi->stats_out; a.stats_in = epi->stats_in; a.ld_stats = epi->ld_stats;
  a.colsum = epi->colsum; a.n_slots = epi->n_slots; a.eps = epi->eps;
  a.ld_aux = epi->ld_aux;
  a.aux_out = epi->mode == EPI_BIAS_GELU_BF16 ? reinterpret_cast<__nv_bfloat16*>(epi->aux_out) : nullptr;
  // TMA drain of the bf16 output (modes 0 / 1 / 7 without a second output; GAVA_EPI_TMA=0 restores per-thread stores):
  // the staged 32x32 blocks leave through bulk tensor stores -- no shared-memory read-back, no per-thread global stores.
  // Measured: fc1 370 -> 354 us in isolation, the inference step +0.7 ... +2.8 % over four same-box A/B pairs.
  static const bool epi_tma = [] { const char* e = getenv("GAVA_EPI_TMA"); return !(e && e[0] == '0'); }();
  CUtensorMap tout = ta, taux = ta;
  a.tma_out = 0;
  if (epi_tma && (epi->mode == EPI_BIAS_BF16 || epi->mode == EPI_BIAS_GELU_BF16 || epi->mode == EPI_LNFOLD_BF16) &&
      !a.colsum && !(reinterpret_cast<uintptr_t>(epi->out) & 15u) && !((epi->ldo * 2) & 15) &&
      (!a.aux_out || (!(reinterpret_cast<uintptr_t>(a.aux_out) & 15u) && !((a.ld_aux * 2) & 15)))) {
    rc = make_tmap_bf16_out(&tout, epi->out, M, N, epi->ldo);
    if (rc) return rc;
    if (a.aux_out) {                               // fc1 of the training forward: the pre-activations leave the same way
      rc = make_tmap_bf16_out(&taux, a.aux_out, M, N, a.ld_aux);
      if (rc) return rc;
    }
    a.tma_out = 1;
  }
  if (pair) return dispatch_mode2(epi->mode, ta, tb, tout, taux, a, stream, epi->max_ctas);
  return bn_eff == 256 ? dispatch_mode<256>(epi->mode, ta, tb, tout, taux, a, stream) : dispatch_mode<128>(epi->mode, ta, tb, tout, taux, a, stream);
}

}  // namespace gava
