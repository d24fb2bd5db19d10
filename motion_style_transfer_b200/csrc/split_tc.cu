// Split-bf16 ("bf16x3") activations: the reference-grade (<= 1e-3 against the fp32 oracle) engine ON the tensor cores.
//
// A float32 activation x is carried as the pair hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits together) in ONE C8
// tensor [N][2 * cp / 8][H][W][8]: hi planes first, then lo planes (cp = channels padded to 16).  A conv is
//   y = W_hi x_hi + W_hi x_lo + W_lo x_hi          (the dropped W_lo x_lo term is 2^-18 relative),
// which is the multi-source tcgen05 conv kernel of conv_tc.cu as it stands, fed with K = [x_hi | x_lo] against
// [W_hi | W_hi] and a second source that re-reads the hi planes against W_lo (ynet_tc_conv3x3_split); its epilogue adds
// the bias, applies the ReLU in float32 and splits the fp32 accumulator again.  kind::tf32 MMAs (10-bit mantissa
// operands) cannot hold 1e-3 through the 30 stacked layers of models/ynet.py:302-470; three bf16 MMAs can,
// at the same bytes per operand element.  This file: the bandwidth-bound companions (pack / unpack / 2x2 max-pool
// (ynet.py:241-245) / bilinear x2 (ynet.py:463)) on that layout, each reading hi + lo, computing in float32 and splitting.
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"

namespace ynet {

namespace {

__device__ __forceinline__ void split_load8(const uint4* hi, long long lo_off, float* f) {
  const uint4 h = __ldg(hi), l = __ldg(hi + lo_off);
  const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
  const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = __bfloat1622float2(hp[k]), b = __bfloat1622float2(lp[k]);
    f[2 * k] = a.x + b.x;          // exact: hi + lo has at most 17 significant bits
    f[2 * k + 1] = a.y + b.y;
  }
}

__device__ __forceinline__ void split_store8(uint4* hi, long long lo_off, const float* f) {
  __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    h[k] = __float2bfloat16_rn(f[k]);
    l[k] = __float2bfloat16_rn(f[k] - __bfloat162float(h[k]));
  }
  *hi = *reinterpret_cast<const uint4*>(h);
  hi[lo_off] = *reinterpret_cast<const uint4*>(l);
}

inline unsigned split_grid(long long total) {
  const long long b = (total + 255) / 256;
  return (unsigned)(b < 1 ? 1 : (b > 148LL * 32 ? 148LL * 32 : b));
}

}  // namespace

// NCHW f32 -> split planes.  One thread = one pixel x one 8-channel chunk (two 16 B stores).
__global__ void __launch_bounds__(256)
split_pack_kernel(const float* __restrict__ x, const float* __restrict__ mask, int C, int H, int W, long long batch_stride,
                  uint4* __restrict__ out, int chunks, int chunks_total, int chunk_off, long long total) {
  const long long S = (long long)H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pix = t % S;
    const long long rest = t / S;
    const int chunk = (int)(rest % chunks);
    const long long n = rest / chunks;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = chunk * 8 + k;
      f[k] = (c < C) ? __ldg(x + n * batch_stride + (long long)c * S + pix) : 0.f;
      // ReLU backward folded into the split (train_epoch.py:109-110 through autograd): dy where the activation was positive
      if (mask != nullptr && c < C && !(__ldg(mask + (n * C + c) * S + pix) > 0.f)) f[k] = 0.f;
    }
    split_store8(out + ((n * 2 * chunks_total + chunk_off + chunk) * S + pix), (long long)chunks_total * S, f);
  }
}

// split planes -> NCHW f32.  One thread = one pixel x one 8-channel chunk: two 16 B loads, eight coalesced 4 B stores
// (one per channel plane).
__global__ void __launch_bounds__(256)
split_unpack_kernel(const uint4* __restrict__ x, int C, int chunks, int H, int W, float* __restrict__ out, long long total) {
  const long long S = (long long)H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int used = (C + 7) >> 3;               // chunks that hold real channels (total = N * used * S)
    const long long pix = t % S;
    const long long rest = t / S;
    const int chunk = (int)(rest % used);
    const long long n = rest / used;
    float f[8];
    split_load8(x + ((n * 2 * chunks + chunk) * S + pix), (long long)chunks * S, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = chunk * 8 + k;
      if (c < C) out[(n * C + c) * S + pix] = f[k];
    }
  }
}

// 2x2 max-pool: the maximum of the four float32 values hi + lo, split again (exact: it is one of the inputs)
__global__ void __launch_bounds__(256)
split_maxpool_kernel(const uint4* __restrict__ x, long long N, int chunks, int H, int W, uint4* __restrict__ out) {
  const int h = H >> 1, w = W >> 1;
  const long long total = N * chunks * h * w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / (h * w);
    const int r = (int)(t - pl * h * w);
    const int y = r / w, xx = r - y * w;
    const long long n = pl / chunks;
    const int chunk = (int)(pl - n * chunks);
    const uint4* q = x + ((n * 2 * chunks + chunk) * H + 2 * y) * W + 2 * xx;
    const long long lo_in = (long long)chunks * H * W;
    float a[8], b[8], c[8], d[8], o[8];
    split_load8(q, lo_in, a);
    split_load8(q + 1, lo_in, b);
    split_load8(q + W, lo_in, c);
    split_load8(q + W + 1, lo_in, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
    split_store8(out + ((n * 2 * chunks + chunk) * h + y) * w + xx, (long long)chunks * h * w, o);
  }
}

// bilinear x2 (align_corners=False), float32 arithmetic in the order of F.interpolate (x pass, then y pass): same
// thread-to-output mapping as c8_upsample_kernel (conv_tc.cu): 4 loads feed a 2x2 output block
__global__ void __launch_bounds__(256)
split_upsample_kernel(const uint4* __restrict__ x, long long N, int chunks, int H, int W, uint4* __restrict__ out) {
  const int OW = 2 * W;
  const int CH = H + 1, CW = W + 1;
  const long long total = N * chunks * CH * CW;
  const long long lo_in = (long long)chunks * H * W, lo_out = 4 * lo_in;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / ((long long)CH * CW);
    const int r = (int)(t - pl * CH * CW);
    const int ci = r / CW, cj = r - ci * CW;
    const int i0 = max(ci - 1, 0), i1 = min(ci, H - 1);
    const int j0 = max(cj - 1, 0), j1 = min(cj, W - 1);
    const long long n = pl / chunks;
    const int chunk = (int)(pl - n * chunks);
    const uint4* q = x + (n * 2 * chunks + chunk) * (long long)H * W;
    float a[8], b[8], c[8], d[8];
    split_load8(q + (size_t)i0 * W + j0, lo_in, a);
    split_load8(q + (size_t)i0 * W + j1, lo_in, b);
    split_load8(q + (size_t)i1 * W + j0, lo_in, c);
    split_load8(q + (size_t)i1 * W + j1, lo_in, d);
    float top_l[8], top_r[8], bot_l[8], bot_r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      top_l[k] = 0.75f * a[k] + 0.25f * b[k];
      top_r[k] = 0.25f * a[k] + 0.75f * b[k];
      bot_l[k] = 0.75f * c[k] + 0.25f * d[k];
      bot_r[k] = 0.25f * c[k] + 0.75f * d[k];
    }
    uint4* o = out + (n * 2 * chunks + chunk) * 4LL * H * W;
    const bool has_l = cj >= 1, has_r = cj <= W - 1;
    float v[8];
    if (ci >= 1) {
      const size_t row = (size_t)(2 * ci - 1) * OW;
      if (has_l) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.75f * top_l[k] + 0.25f * bot_l[k];
        split_store8(o + row + 2 * cj - 1, lo_out, v);
      }
      if (has_r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.75f * top_r[k] + 0.25f * bot_r[k];
        split_store8(o + row + 2 * cj, lo_out, v);
      }
    }
    if (ci <= H - 1) {
      const size_t row = (size_t)(2 * ci) * OW;
      if (has_l) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.25f * top_l[k] + 0.75f * bot_l[k];
        split_store8(o + row + 2 * cj - 1, lo_out, v);
      }
      if (has_r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.25f * top_r[k] + 0.75f * bot_r[k];
        split_store8(o + row + 2 * cj, lo_out, v);
      }
    }
  }
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int ynet_split_pack_f32(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, int64_t batch_stride, void* out,
                        int32_t C_pad, int32_t out_total_pad, int32_t out_channel_off, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && C_pad >= C && C_pad % 16 == 0, "bad shape (C_pad % 16 == 0)");
  if (out_total_pad == 0) out_total_pad = C_pad;
  YNET_CHECK_ARG(out_total_pad % 16 == 0 && out_channel_off % 16 == 0 && out_channel_off >= 0 &&
                     out_channel_off + C_pad <= out_total_pad,
                 "output slice outside the activation (multiples of 16)");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x && out, "null pointer");
  YNET_CHECK_ALIGN(out, 16);
  const long long total = (long long)N * (C_pad / 8) * H * W;
  split_pack_kernel<<<split_grid(total), 256, 0, as_stream(stream)>>>(x, nullptr, C, H, W, batch_stride,
                                                                      reinterpret_cast<uint4*>(out), C_pad / 8,
                                                                      out_total_pad / 8, out_channel_off / 8, total);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_split_pack_masked_f32(const float* x, const float* relu_out, int32_t N, int32_t C, int32_t H, int32_t W, void* out,
                               int32_t C_pad, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && C_pad >= C && C_pad % 16 == 0, "bad shape (C_pad % 16 == 0)");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x && relu_out && out, "null pointer");
  YNET_CHECK_ALIGN(out, 16);
  const long long total = (long long)N * (C_pad / 8) * H * W;
  split_pack_kernel<<<split_grid(total), 256, 0, as_stream(stream)>>>(x, relu_out, C, H, W, (long long)C * H * W,
                                                                      reinterpret_cast<uint4*>(out), C_pad / 8, C_pad / 8, 0,
                                                                      total);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_split_unpack_f32(const void* x, int32_t N, int32_t C, int32_t C_pad, int32_t H, int32_t W, float* out, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && C_pad >= C && C_pad % 16 == 0, "bad shape");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x && out, "null pointer");
  const long long total = (long long)N * ((C + 7) / 8) * H * W;
  split_unpack_kernel<<<split_grid(total), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(x), C, C_pad / 8, H, W,
                                                                        out, total);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_split_maxpool2x2(const void* x, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C_pad % 16 == 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "bad shape");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x && out, "null pointer");
  const long long total = (long long)N * (C_pad / 8) * (H / 2) * (W / 2);
  split_maxpool_kernel<<<split_grid(total), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(x), N, C_pad / 8, H, W,
                                                                         reinterpret_cast<uint4*>(out));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_split_upsample2x(const void* x, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C_pad % 16 == 0 && H >= 1 && W >= 1, "bad shape");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x && out, "null pointer");
  const long long total = (long long)N * (C_pad / 8) * (H + 1) * (W + 1);
  split_upsample_kernel<<<split_grid(total), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(x), N, C_pad / 8, H,
                                                                          W, reinterpret_cast<uint4*>(out));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

}  // extern "C"
