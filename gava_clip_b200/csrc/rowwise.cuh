// Warp-per-row LayerNorm helpers shared by the row-wise kernels and the GEMM's LayerNorm tail.
#pragma once
#include "common.cuh"

namespace gava {

// ------------------------------------------------------------------------------------------------
// LayerNorm (VitaCLIP_vision_encoder_utils.py:22-28): stats in fp32, eps inside the sqrt, affine.
// A lane owns float4 #(lane + 32 j), j < V, of the row (D = 128 V).
template <int V>
__device__ __forceinline__ void ln_normalize(float4 (&x)[V], const float* __restrict__ w, const float* __restrict__ b,
                                             float eps) {
  constexpr float inv_d = 1.0f / (128.0f * V);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
  const float mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    x[j].x -= mean; x[j].y -= mean; x[j].z -= mean; x[j].w -= mean;
    q += (x[j].x * x[j].x + x[j].y * x[j].y) + (x[j].z * x[j].z + x[j].w * x[j].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
  const int lane = lane_id();
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * j);
    const float4 o = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * j);
    x[j].x = x[j].x * rstd * g.x + o.x;
    x[j].y = x[j].y * rstd * g.y + o.y;
    x[j].z = x[j].z * rstd * g.z + o.z;
    x[j].w = x[j].w * rstd * g.w + o.w;
  }
}

template <int V>
__device__ __forceinline__ void store_row(const float4 (&x)[V], void* out_row, int out_bf16) {
  const int lane = lane_id();
  if (out_bf16) {
    uint2* o = reinterpret_cast<uint2*>(out_row);
#pragma unroll
    for (int j = 0; j < V; ++j) o[lane + 32 * j] = make_uint2(pack_bf16x2(x[j].x, x[j].y), pack_bf16x2(x[j].z, x[j].w));
  } else {
    float4* o = reinterpret_cast<float4*>(out_row);
#pragma unroll
    for (int j = 0; j < V; ++j) o[lane + 32 * j] = x[j];
  }
}

}  // namespace gava
