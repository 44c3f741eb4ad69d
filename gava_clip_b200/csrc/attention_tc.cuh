// tcgen05 / TMEM helpers shared by the attention forward (attention_tc.cu) and backward (attention_bwd_tc.cu) kernels.
#pragma once
#include "common.cuh"

namespace gava {

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
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void st_shared_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

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

// valid keys of the smem row space: live rows [0,v0) | global + local prompts [v1,v2) | summary token at v3
struct KeyMask {
  int v0, v1, v2, v3;
  __device__ __forceinline__ bool ok(int k) const { return k < v0 || (k >= v1 && k < v2) || k == v3; }
};

// K-major tile WITHOUT swizzle, 16 bf16 wide: rows in groups of 8; a group is [k-chunk 0: 8 rows x 16 B][k-chunk 1: 8 rows x
// 16 B] = 256 B.  Canonical layout ((8,n),(8,2)):((16 B, SBO),(2 B, LBO)) with LBO = 128 B, SBO = 256 B.
__device__ __forceinline__ uint64_t make_nosw_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128u >> 4) << 16;                    // leading byte offset: between the two 16-byte k-chunks
  d |= (uint64_t)(256u >> 4) << 32;                    // stride byte offset: between 8-row groups
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  return d;                                            // layout type 0 = no swizzle
}
__device__ __forceinline__ uint32_t nosw_offset(int row, int col) {   // byte offset of element (row, col < 16) in such a tile
  return (uint32_t)((row >> 3) * 256 + (col >> 3) * 128 + (row & 7) * 16 + (col & 7) * 2);
}

}  // namespace gava
