// scene-image preprocessing in front of the hot path, fused into one kernel per scene:
//   resize (cv2.INTER_AREA, utils/image_utils.py:85-92) -> pad to a multiple of 32 with zeros (95-107) ->
//   segmentation-backbone normalisation (x / 255 - mean) / std and HWC -> CHW float32 (66-82; trainer.py:578-582),
// and for segmentation masks resize (INTER_NEAREST) -> pad -> one-hot.
// Byte work, bit-exact against OpenCV: the area tables are the ones cv::computeResizeAreaTab builds (host side,
// utils/image_utils.py::area_table), the float32 accumulation keeps OpenCV's order -- per source row the x pass
// buf = ((0 + s0 a0) + s1 a1) + ..., then sum += buf * beta over the rows -- with separate multiply and add (no FMA
// contraction), the result is rounded half-to-even like saturate_cast<uchar>.  Integer scale factors take OpenCV's
// integer path (block sum, * float(1 / area), or (sum + 2) >> 2 for 2 x 2).  HBM-bound: reads 3 B per source pixel once
// (uint8 HWC, coalesced through the per-row x walk), writes 12 B per padded output pixel.
#include <stdint.h>

#include "common.cuh"

namespace ynet {

struct AreaTab {
  const int32_t* start;   // CSR over destination indices: entries [start[d], start[d + 1])
  const int32_t* src;     // source index of every entry
  const float* w;         // float32 weight (alpha / beta)
};

// Augmented views (data_utils.py:115-233: rot90 x k, then fliplr) are read straight from the ORIGINAL image: pixel (y, x)
// of the view cv2.flip(cv2.rotate(img) x k, 1) [orient = k + 4 * flip] is pixel src_pixel(...) of the stored (H0, W0)
// image, so the eight views of a scene cost one upload and no rotated copies; the resize arithmetic runs in the view's
// frame, which keeps it bit-exact against resizing the rotated image.
struct Orient {
  int k, flip, H0, W0;      // H0 x W0: the stored image; the view is W0 x H0 for odd k
};
__device__ __forceinline__ size_t src_pixel(const Orient& o, int W, int y, int x) {
  if (o.flip) x = W - 1 - x;
  int oy, ox;
  switch (o.k) {
    case 1: oy = x; ox = o.W0 - 1 - y; break;               // cv2.ROTATE_90_COUNTERCLOCKWISE: out(i, j) = in(j, W0 - 1 - i)
    case 2: oy = o.H0 - 1 - y; ox = o.W0 - 1 - x; break;
    case 3: oy = o.H0 - 1 - x; ox = y; break;
    default: oy = y; ox = x; break;
  }
  return (size_t)oy * o.W0 + ox;
}

// mode 0: table path; 1: integer block sum * fast_scale; 2: 2 x 2 block, (sum + 2) >> 2.  H x W: the VIEW's size.
__global__ void __launch_bounds__(256)
scene_preprocess_kernel(const uint8_t* __restrict__ img, int H, int W, int dh, int dw, int Hp, int Wp, AreaTab xt, AreaTab yt,
                        int mode, int isc, float fast_scale, double m0, double m1, double m2, double s0, double s1, double s2,
                        float* __restrict__ out_chw, uint8_t* __restrict__ out_u8, Orient o) {
  const long long total = (long long)Hp * Wp;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(t / Wp), x = (int)(t - (long long)y * Wp);
    int v[3] = {0, 0, 0};                         // padded pixels are zeros BEFORE the normalisation (image_utils.py:104-106)
    if (y < dh && x < dw) {
      if (mode == 0) {
        float acc[3] = {0.f, 0.f, 0.f};
        for (int r = yt.start[y]; r < yt.start[y + 1]; ++r) {
          const int sy = yt.src[r];
          float bx[3] = {0.f, 0.f, 0.f};
          for (int c = xt.start[x]; c < xt.start[x + 1]; ++c) {
            const uint8_t* px = img + src_pixel(o, W, sy, xt.src[c]) * 3;
            const float a = xt.w[c];
#pragma unroll
            for (int k = 0; k < 3; ++k) bx[k] = __fadd_rn(bx[k], __fmul_rn((float)px[k], a));
          }
          const float b = yt.w[r];
#pragma unroll
          for (int k = 0; k < 3; ++k) acc[k] = __fadd_rn(acc[k], __fmul_rn(bx[k], b));
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = min(max(__float2int_rn(acc[k]), 0), 255);
      } else {
        int sum[3] = {0, 0, 0};
        for (int ky = 0; ky < isc; ++ky) {
          for (int kx = 0; kx < isc; ++kx) {
            const uint8_t* px = img + src_pixel(o, W, y * isc + ky, x * isc + kx) * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) sum[k] += px[k];
          }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
          v[k] = (mode == 2) ? ((sum[k] + 2) >> 2) : min(max(__float2int_rn(__fmul_rn((float)sum[k], fast_scale)), 0), 255);
      }
      if (out_u8 != nullptr) {
        uint8_t* q = out_u8 + ((size_t)y * dw + x) * 3;
        q[0] = (uint8_t)v[0];
        q[1] = (uint8_t)v[1];
        q[2] = (uint8_t)v[2];
      }
    }
    if (out_chw != nullptr) {
      // smp preprocess_input: float64 (x / 255 - mean) / std, then astype(float32)
      const size_t plane = (size_t)Hp * Wp;
      out_chw[t] = (float)(((double)v[0] / 255.0 - m0) / s0);
      out_chw[plane + t] = (float)(((double)v[1] / 255.0 - m1) / s1);
      out_chw[2 * plane + t] = (float)(((double)v[2] / 255.0 - m2) / s2);
    }
  }
}

// segmentation masks: INTER_NEAREST (src = min(floor(dst / f), size - 1)) -> zero pad -> one-hot over `classes`
__global__ void __launch_bounds__(256)
scene_onehot_kernel(const uint8_t* __restrict__ mask, int H, int W, int dh, int dw, int Hp, int Wp, double inv_f, int classes,
                    float* __restrict__ out) {
  const long long total = (long long)Hp * Wp;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(t / Wp), x = (int)(t - (long long)y * Wp);
    int v = 0;
    if (y < dh && x < dw) {
      const int sy = min((int)floor((double)y * inv_f), H - 1), sx = min((int)floor((double)x * inv_f), W - 1);
      v = mask[(size_t)sy * W + sx];
    }
    for (int c = 0; c < classes; ++c) out[(size_t)c * total + t] = (v == c) ? 1.f : 0.f;
  }
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int ynet_scene_preprocess_oriented_u8(const uint8_t* img_hwc, int32_t H0, int32_t W0, int32_t orient, int32_t dh, int32_t dw,
                                      int32_t Hp, int32_t Wp, const int32_t* xt_start, const int32_t* xt_src,
                                      const float* xt_w, const int32_t* yt_start, const int32_t* yt_src, const float* yt_w,
                                      int32_t int_scale, const double* mean3_host, const double* std3_host, float* out_chw,
                                      uint8_t* out_u8_hwc, void* stream) {
  YNET_CHECK_ARG(orient >= 0 && orient < 8, "orient = k + 4 * flip, k in [0, 3]");
  YNET_CHECK_ARG(img_hwc && (out_chw || out_u8_hwc), "null pointer");
  const int32_t H = (orient & 1) ? W0 : H0, W = (orient & 1) ? H0 : W0;       // the view's size
  YNET_CHECK_ARG(H > 0 && W > 0 && dh > 0 && dw > 0 && Hp >= dh && Wp >= dw, "bad shape");
  YNET_CHECK_ARG(int_scale >= 0 && (int_scale > 0 || (xt_start && xt_src && xt_w && yt_start && yt_src && yt_w)),
                 "area tables missing");
  YNET_CHECK_ARG(int_scale == 0 || ((long long)dh * int_scale <= H && (long long)dw * int_scale <= W),
                 "integer-scale path needs dsize * scale <= ssize");
  YNET_CHECK_ARG(out_chw == nullptr || (mean3_host && std3_host), "normalisation constants missing");
  AreaTab xt{xt_start, xt_src, xt_w}, yt{yt_start, yt_src, yt_w};
  const int mode = int_scale == 0 ? 0 : (int_scale == 2 ? 2 : 1);
  const float fs = int_scale ? (float)(1.0 / ((double)int_scale * int_scale)) : 0.f;
  const double zero3[3] = {0, 0, 0}, one3[3] = {1, 1, 1};
  const double* m = mean3_host ? mean3_host : zero3;
  const double* s = std3_host ? std3_host : one3;
  const long long total = (long long)Hp * Wp;
  const unsigned grid = (unsigned)tmax<long long>(1, tmin<long long>(ceil_div<long long>(total, 256), 16LL * sm_count()));
  const Orient o{orient & 3, orient >> 2, H0, W0};
  scene_preprocess_kernel<<<grid, 256, 0, as_stream(stream)>>>(img_hwc, H, W, dh, dw, Hp, Wp, xt, yt, mode, int_scale, fs, m[0],
                                                               m[1], m[2], s[0], s[1], s[2], out_chw, out_u8_hwc, o);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_scene_preprocess_u8(const uint8_t* img_hwc, int32_t H, int32_t W, int32_t dh, int32_t dw, int32_t Hp, int32_t Wp,
                             const int32_t* xt_start, const int32_t* xt_src, const float* xt_w, const int32_t* yt_start,
                             const int32_t* yt_src, const float* yt_w, int32_t int_scale, const double* mean3_host,
                             const double* std3_host, float* out_chw, uint8_t* out_u8_hwc, void* stream) {
  return ynet_scene_preprocess_oriented_u8(img_hwc, H, W, 0, dh, dw, Hp, Wp, xt_start, xt_src, xt_w, yt_start, yt_src, yt_w,
                                           int_scale, mean3_host, std3_host, out_chw, out_u8_hwc, stream);
}

int ynet_scene_onehot_u8(const uint8_t* mask, int32_t H, int32_t W, int32_t dh, int32_t dw, int32_t Hp, int32_t Wp,
                         double inv_factor, int32_t classes, float* out, void* stream) {
  YNET_CHECK_ARG(mask && out, "null pointer");
  YNET_CHECK_ARG(H > 0 && W > 0 && dh > 0 && dw > 0 && Hp >= dh && Wp >= dw && classes > 0 && inv_factor > 0, "bad shape");
  const long long total = (long long)Hp * Wp;
  const unsigned grid = (unsigned)tmax<long long>(1, tmin<long long>(ceil_div<long long>(total, 256), 16LL * sm_count()));
  scene_onehot_kernel<<<grid, 256, 0, as_stream(stream)>>>(mask, H, W, dh, dw, Hp, Wp, inv_factor, classes, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

}  // extern "C"
