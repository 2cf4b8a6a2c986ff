// GPU image pipeline of the crop frontier: Pillow-exact separable BICUBIC resampling of uint8 RGB crops
// (antialiased: filter support scales with the down-scale factor), fused with /255 + CLIP mean/std normalisation
// and the bf16 cast, reading the crop straight out of the search image resident in HBM.
//
// Replaces the reference's per-crop host work (/root/reference/visual_search.py:186-194): PIL crop + deepcopy,
// expand2square (VisualSearch/utils/utils.py:28-39), CLIPImageProcessor / OwlViTProcessor bicubic resizes (Pillow
// ImagingResample, 8bpc path), rescale, normalise, .cuda(), .bfloat16().
//
// Bit-exactness: Pillow's 8-bit path is integer arithmetic — coefficients are computed in double, normalised and
// rounded to 22-bit fixed point (PRECISION_BITS = 32-8-2), each pass accumulates pixel*coef in int32 starting from
// 1<<21 and clips (acc >> 22) to [0,255], horizontal pass first, uint8 intermediate.  The coefficient tables are
// built on the host with the same double arithmetic (vstar_b200/image.py); the kernels below reproduce the integer
// passes, so the uint8 result equals PIL's and the normalised fp32 value equals numpy's (IEEE fp32 mul/sub/div).
#include "common.cuh"
#include "vstar_b200.h"

namespace {

constexpr int PRECISION_BITS = 22;

__device__ __forceinline__ int clip8(int v) {
  v >>= PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass over a (virtually padded) crop: tmp[y, xo, c] for y in [0,in_h), xo in [0,out_w)
// virtual input pixel (x,y) = src crop pixel if x < cw && y < ch else bg (expand2square pads bottom/right)
__global__ void resample_h_kernel(const uint8_t* __restrict__ src, long long row_stride, int x0, int y0, int cw, int ch, int in_h,
                                  int bg0, int bg1, int bg2, const int* __restrict__ coefs, const int* __restrict__ bounds, int ksize,
                                  int out_w, uint8_t* __restrict__ tmp) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (xo >= out_w || y >= in_h) return;
  const int xmin = bounds[2 * xo], xcnt = bounds[2 * xo + 1];
  const int* k = coefs + (long long)xo * ksize;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  if (y < ch) {
    const uint8_t* row = src + (long long)(y0 + y) * row_stride + (long long)x0 * 3;
    for (int i = 0; i < xcnt; ++i) {
      const int x = xmin + i;
      const int kk = k[i];
      if (x < cw) {
        s0 += row[3 * x] * kk; s1 += row[3 * x + 1] * kk; s2 += row[3 * x + 2] * kk;
      } else {
        s0 += bg0 * kk; s1 += bg1 * kk; s2 += bg2 * kk;
      }
    }
  } else {
    for (int i = 0; i < xcnt; ++i) { const int kk = k[i]; s0 += bg0 * kk; s1 += bg1 * kk; s2 += bg2 * kk; }
  }
  uint8_t* o = tmp + ((long long)y * out_w + xo) * 3;
  o[0] = (uint8_t)clip8(s0); o[1] = (uint8_t)clip8(s1); o[2] = (uint8_t)clip8(s2);
}

// vertical pass + (optional) normalisation: out_u8 [out_h,out_w,3] and/or out_bf16 CHW [3,out_h,out_w]
__global__ void resample_v_kernel(const uint8_t* __restrict__ tmp, int out_w, const int* __restrict__ coefs,
                                  const int* __restrict__ bounds, int ksize, int out_h, uint8_t* __restrict__ out_u8,
                                  bf16* __restrict__ out_bf16, float* __restrict__ out_f32, float m0, float m1, float m2, float d0,
                                  float d1, float d2) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x;
  const int yo = blockIdx.y;
  if (xo >= out_w || yo >= out_h) return;
  const int ymin = bounds[2 * yo], ycnt = bounds[2 * yo + 1];
  const int* k = coefs + (long long)yo * ksize;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  const uint8_t* p = tmp + ((long long)ymin * out_w + xo) * 3;
  for (int j = 0; j < ycnt; ++j) {
    const int kk = k[j];
    s0 += p[0] * kk; s1 += p[1] * kk; s2 += p[2] * kk;
    p += (long long)out_w * 3;
  }
  const int c0 = clip8(s0), c1 = clip8(s1), c2 = clip8(s2);
  if (out_u8) {
    uint8_t* o = out_u8 + ((long long)yo * out_w + xo) * 3;
    o[0] = (uint8_t)c0; o[1] = (uint8_t)c1; o[2] = (uint8_t)c2;
  }
  if (out_bf16 || out_f32) {
    // numpy: x.astype(float32) * float32(1/255); (x - mean) / std   (all fp32, IEEE)
    const float r = 0.00392156862745098f;
    const float f0 = __fdiv_rn(__fsub_rn(__fmul_rn((float)c0, r), m0), d0);
    const float f1 = __fdiv_rn(__fsub_rn(__fmul_rn((float)c1, r), m1), d1);
    const float f2 = __fdiv_rn(__fsub_rn(__fmul_rn((float)c2, r), m2), d2);
    const long long plane = (long long)out_h * out_w, idx = (long long)yo * out_w + xo;
    if (out_bf16) { out_bf16[idx] = f2bf(f0); out_bf16[plane + idx] = f2bf(f1); out_bf16[2 * plane + idx] = f2bf(f2); }
    if (out_f32) { out_f32[idx] = f0; out_f32[plane + idx] = f1; out_f32[2 * plane + idx] = f2; }
  }
}

}  // namespace

extern "C" int vsb_resample_h_u8(const void* src, long long row_stride, int x0, int y0, int cw, int ch, int in_h, int bg0, int bg1,
                                 int bg2, const void* coefs, const void* bounds, int ksize, int out_w, void* tmp, void* stream) {
  VSB_CHECK_ARG(src && coefs && bounds && tmp && cw > 0 && ch > 0 && in_h > 0 && out_w > 0 && ksize > 0, "vsb_resample_h_u8: bad args");
  dim3 grid((out_w + 127) / 128, in_h);
  resample_h_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const uint8_t*)src, row_stride, x0, y0, cw, ch, in_h, bg0, bg1,
                                                                              bg2, (const int*)coefs, (const int*)bounds, ksize, out_w,
                                                                              (uint8_t*)tmp);
  VSB_LAUNCH_CHECK();
  return VSB_OK;
}

extern "C" int vsb_resample_v_u8(const void* tmp, int out_w, const void* coefs, const void* bounds, int ksize, int out_h, void* out_u8,
                                 void* out_bf16_chw, void* out_f32_chw, const float* mean3, const float* std3, void* stream) {
  VSB_CHECK_ARG(tmp && coefs && bounds && out_w > 0 && out_h > 0 && ksize > 0 && mean3 && std3, "vsb_resample_v_u8: bad args");
  VSB_CHECK_ARG(out_u8 || out_bf16_chw || out_f32_chw, "vsb_resample_v_u8: no output");
  dim3 grid((out_w + 127) / 128, out_h);
  resample_v_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const uint8_t*)tmp, out_w, (const int*)coefs, (const int*)bounds,
                                                                              ksize, out_h, (uint8_t*)out_u8, (bf16*)out_bf16_chw,
                                                                              (float*)out_f32_chw, mean3[0], mean3[1], mean3[2], std3[0],
                                                                              std3[1], std3[2]);
  VSB_LAUNCH_CHECK();
  return VSB_OK;
}
