// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM as a function of warps and load shape, and the
// MUFU.EX2 rate beside it.  Decides how the attention softmax is organised (attention_tc.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cuda/tmem_bw tools/cuda/tmem_bw.cu && tools/cuda/tmem_bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld<8>(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}

// mode 0: loads only; mode 1: loads + one exp2 per element; mode 2: exp2 only (registers)
template <int X, int MODE>
__global__ void __launch_bounds__(1024, 1) k(long long* cycles, float* sink, int iters, int cols_span) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  uint32_t va[X], vb[X];
  for (int j = 0; j < X; ++j) { va[j] = threadIdx.x + j; vb[j] = j; }
  __syncthreads();
  const long long t0 = clock64();
  uint32_t col = (uint32_t)((warp >> 2) * X) % (uint32_t)cols_span;
  for (int i = 0; i < iters; ++i) {
    if (MODE != 2) {
      ld<X>(base + col, va);
      ld<X>(base + ((col + X) % (uint32_t)cols_span), vb);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      col = (col + 2 * X) % (uint32_t)cols_span;
    }
    if (MODE == 0) {
      acc += __uint_as_float(va[0] ^ vb[X - 1]);
    } else {
#pragma unroll
      for (int j = 0; j < X; ++j) {
        acc += exp2f(__uint_as_float(va[j]) * 0.18f - 3.f);
        acc += exp2f(__uint_as_float(vb[j]) * 0.18f - 3.f);
        if (MODE == 2) { va[j] += 1; vb[j] += 3; }
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 1234.5f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int X, int MODE>
void run(int warps, const char* name) {
  long long* d; float* s;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  k<X, MODE><<<148, warps * 32, 0>>>(d, s, iters, 256);
  k<X, MODE><<<148, warps * 32, 0>>>(d, s, iters, 256);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)warps * iters * 2 * X * 32 * 4;   // per SM
  printf("%-28s X=%2d warps=%2d: %8lld cycles, %7.1f B/clk/SM, %6.2f elems/clk/SM (%s)\n", name, X, warps, h[0], bytes / h[0],
         bytes / 4 / h[0], cudaGetErrorString(e));
  cudaFree(d); cudaFree(s);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<16, 0>(w, "ldtm only");
  for (int w : {4, 8, 16}) run<32, 0>(w, "ldtm only");
  for (int w : {4, 8, 16}) run<8, 0>(w, "ldtm only");
  for (int w : {4, 8, 16}) run<16, 1>(w, "ldtm + exp2 per element");
  for (int w : {4, 8, 16}) run<16, 2>(w, "exp2 only");
  return 0;
}
