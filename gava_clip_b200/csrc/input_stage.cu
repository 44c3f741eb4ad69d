// Input stage of the evaluation path fused into one bandwidth-bound kernel:
//   decoder output uint8 [T0, H0, W0, 3]  ->  /255, (x - mean) / std  ->  bilinear resize (short side -> spatial size,
//   align_corners = False, no antialias)  ->  spatial crop  ->  temporal crop (start, stride, last frame repeated)
//   ->  f32 clip [3, T, Sh, Sw]  (the (C, T, H, W) layout VitaCLIP.forward takes)
// Restates video_dataset/dataset.py:117-139 (normalise, permute, F.interpolate), :160-196 (temporal / spatial crops).
// The corner pixels are normalised first and interpolated second, in the reference's order and with ATen's
// upsample_bilinear2d weights (src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out).
#include "common.cuh"
#include "kernels.h"

namespace gava {

struct PreArgs {
  const uint8_t* frames;
  float* out;
  int T0, H0, W0, h_st, w_st, t_st, t_step, T, Sh, Sw;
  float scale_h, scale_w;
  float mean[3], std[3];
};

__global__ void __launch_bounds__(256) preprocess_u8_kernel(const PreArgs p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.T * p.Sh * p.Sw;
  if (idx >= total) return;
  const int x = (int)(idx % p.Sw);
  const int y = (int)((idx / p.Sw) % p.Sh);
  const int t = (int)(idx / ((long long)p.Sw * p.Sh));
  int ts = p.t_st + t * p.t_step;
  if (ts > p.T0 - 1) ts = p.T0 - 1;                     // short videos: the last frame is repeated (dataset.py:163-164)
  // ATen area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
  float sy = p.scale_h * ((float)(y + p.h_st) + 0.5f) - 0.5f;
  float sx = p.scale_w * ((float)(x + p.w_st) + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < p.H0 - 1 ? 1 : 0), x1 = x0 + (x0 < p.W0 - 1 ? 1 : 0);
  const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const uint8_t* f = p.frames + (long long)ts * p.H0 * p.W0 * 3;
  const uint8_t* p00 = f + ((long long)y0 * p.W0 + x0) * 3;
  const uint8_t* p01 = f + ((long long)y0 * p.W0 + x1) * 3;
  const uint8_t* p10 = f + ((long long)y1 * p.W0 + x0) * 3;
  const uint8_t* p11 = f + ((long long)y1 * p.W0 + x1) * 3;
  const long long plane = (long long)p.T * p.Sh * p.Sw;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float m = p.mean[c], s = p.std[c];
    // IEEE divisions (the build uses --use_fast_math): same values as `frames.float() / 255.; (frames - mean) / std`
    const float v00 = __fdiv_rn(__fdiv_rn((float)p00[c], 255.f) - m, s), v01 = __fdiv_rn(__fdiv_rn((float)p01[c], 255.f) - m, s);
    const float v10 = __fdiv_rn(__fdiv_rn((float)p10[c], 255.f) - m, s), v11 = __fdiv_rn(__fdiv_rn((float)p11[c], 255.f) - m, s);
    p.out[c * plane + idx] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
  }
}

int preprocess_u8(const uint8_t* frames, float* out, int T0, int H0, int W0, int new_h, int new_w, int h_st, int w_st,
                  int t_st, int t_step, int T, int Sh, int Sw, const float* mean, const float* std, cudaStream_t stream) {
  if (!frames || !out || !mean || !std || T0 <= 0 || H0 <= 0 || W0 <= 0 || new_h <= 0 || new_w <= 0 || T <= 0 || Sh <= 0 ||
      Sw <= 0 || h_st < 0 || w_st < 0 || h_st + Sh > new_h || w_st + Sw > new_w || t_st < 0 || t_step <= 0) {
    set_error("preprocess_u8: bad arguments");
    return GAVA_ERR_ARG;
  }
  PreArgs a;
  a.frames = frames; a.out = out;
  a.T0 = T0; a.H0 = H0; a.W0 = W0; a.h_st = h_st; a.w_st = w_st; a.t_st = t_st; a.t_step = t_step; a.T = T; a.Sh = Sh; a.Sw = Sw;
  a.scale_h = (float)H0 / (float)new_h;
  a.scale_w = (float)W0 / (float)new_w;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean[c]; a.std[c] = std[c]; }
  const long long total = (long long)T * Sh * Sw;
  preprocess_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a);
  return check_launch("preprocess_u8");
}

}  // namespace gava
