// a10/a12/a13: sigmoid map, per-channel spatial soft-argmax / softmax / expectation.
// HBM-bound: one streaming read of each map (online softmax), 128-bit loads, warp-shuffle reductions.
#include <float.h>

#include "common.cuh"

namespace ynet {

struct SoftPartial {  // running (max, sum e, sum e*x, sum e*y)
  float m, s, sx, sy;
};

__device__ __forceinline__ SoftPartial combine(const SoftPartial& a, const SoftPartial& b) {
  SoftPartial r;
  r.m = fmaxf(a.m, b.m);
  const float fa = (a.m == -FLT_MAX) ? 0.f : __expf(a.m - r.m);
  const float fb = (b.m == -FLT_MAX) ? 0.f : __expf(b.m - r.m);
  r.s = a.s * fa + b.s * fb;
  r.sx = a.sx * fa + b.sx * fb;
  r.sy = a.sy * fa + b.sy * fb;
  return r;
}

__device__ __forceinline__ SoftPartial warp_combine(SoftPartial p) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    SoftPartial q;
    q.m = __shfl_xor_sync(0xffffffffu, p.m, o);
    q.s = __shfl_xor_sync(0xffffffffu, p.s, o);
    q.sx = __shfl_xor_sync(0xffffffffu, p.sx, o);
    q.sy = __shfl_xor_sync(0xffffffffu, p.sy, o);
    p = combine(p, q);
  }
  return p;
}

__device__ __forceinline__ void online_update(SoftPartial& p, float v, float x, float y) {
  // caller guarantees p.m >= v
  const float e = __expf(v - p.m);
  p.s += e;
  p.sx = fmaf(e, x, p.sx);
  p.sy = fmaf(e, y, p.sy);
}

__device__ __forceinline__ void raise_max(SoftPartial& p, float nm) {
  if (nm > p.m) {
    const float f = (p.m == -FLT_MAX) ? 0.f : __expf(p.m - nm);
    p.s *= f;
    p.sx *= f;
    p.sy *= f;
    p.m = nm;
  }
}

constexpr int kSoftThreads = 256;

// grid = (splits, rows).  Each CTA reduces a contiguous slice of one H*W map to one SoftPartial.
template <bool VEC4>
__global__ void __launch_bounds__(kSoftThreads)
softargmax_partial_kernel(const float* __restrict__ x, long long row_stride, int H, int W, int splits,
                          SoftPartial* __restrict__ part) {
  const int row = blockIdx.y, split = blockIdx.x;
  const int S = H * W;
  const float* src = x + (size_t)row * row_stride;
  SoftPartial p{-FLT_MAX, 0.f, 0.f, 0.f};
  if (VEC4) {
    const int q_total = S >> 2;
    const int per = ceil_div(q_total, splits);
    const int q0 = split * per, q1 = min(q_total, q0 + per);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int q = q0 + threadIdx.x; q < q1; q += 2 * kSoftThreads) {
      const int qb = q + kSoftThreads;
      const bool hb = qb < q1;
      const float4 a = ld_stream(s4 + q);
      const float4 b = hb ? ld_stream(s4 + qb) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
      float mx = fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w));
      mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
      raise_max(p, mx);
      {
        const int base = q << 2;
        const int yy = base / W;
        const float fy = (float)yy, fx = (float)(base - yy * W);
        online_update(p, a.x, fx, fy);
        online_update(p, a.y, fx + 1.f, fy);
        online_update(p, a.z, fx + 2.f, fy);
        online_update(p, a.w, fx + 3.f, fy);
      }
      if (hb) {
        const int base = qb << 2;
        const int yy = base / W;
        const float fy = (float)yy, fx = (float)(base - yy * W);
        online_update(p, b.x, fx, fy);
        online_update(p, b.y, fx + 1.f, fy);
        online_update(p, b.z, fx + 2.f, fy);
        online_update(p, b.w, fx + 3.f, fy);
      }
    }
  } else {
    const int per = ceil_div(S, splits);
    const int i0 = split * per, i1 = min(S, i0 + per);
    for (int i = i0 + threadIdx.x; i < i1; i += kSoftThreads) {
      const float v = src[i];
      raise_max(p, v);
      const int yy = i / W;
      online_update(p, v, (float)(i - yy * W), (float)yy);
    }
  }
  p = warp_combine(p);
  __shared__ SoftPartial sh[kSoftThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = p;
  __syncthreads();
  if (warp == 0) {
    SoftPartial q = (lane < kSoftThreads / 32) ? sh[lane] : SoftPartial{-FLT_MAX, 0.f, 0.f, 0.f};
    q = warp_combine(q);
    if (lane == 0) part[(size_t)row * splits + split] = q;
  }
}

// one warp per row: combine `splits` partials, apply 1/(sum + 1e-6)  (softargmax.py:68)
__global__ void softargmax_finalize_kernel(const SoftPartial* __restrict__ part, int rows, int splits,
                                           float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  SoftPartial p{-FLT_MAX, 0.f, 0.f, 0.f};
  for (int i = lane; i < splits; i += 32) p = combine(p, part[(size_t)row * splits + i]);
  p = warp_combine(p);
  if (lane == 0) {
    const float inv = 1.0f / (p.s + 1e-6f);
    out[2 * row + 0] = p.sx * inv;
    out[2 * row + 1] = p.sy * inv;
  }
}

// a13: softmax over the flattened map, one CTA per row (second read is served by L2).
__global__ void __launch_bounds__(1024) spatial_softmax_kernel(const float* __restrict__ x, long long S,
                                                               float* __restrict__ out) {
  const float* src = x + (size_t)blockIdx.x * S;
  float* dst = out + (size_t)blockIdx.x * S;
  float m = -FLT_MAX, s = 0.f;
  for (long long i = threadIdx.x; i < S; i += blockDim.x) {
    const float v = src[i];
    if (v > m) {
      s *= (m == -FLT_MAX) ? 0.f : __expf(m - v);
      m = v;
    }
    s += __expf(v - m);
  }
  __shared__ float shm[32], shs[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float M = fmaxf(m, m2);
    s = s * ((m == -FLT_MAX) ? 0.f : __expf(m - M)) + s2 * ((m2 == -FLT_MAX) ? 0.f : __expf(m2 - M));
    m = M;
  }
  if (lane == 0) {
    shm[warp] = m;
    shs[warp] = s;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float M = -FLT_MAX;
  for (int w = 0; w < nw; ++w) M = fmaxf(M, shm[w]);
  float tot = 0.f;
  for (int w = 0; w < nw; ++w) tot += shs[w] * ((shm[w] == -FLT_MAX) ? 0.f : __expf(shm[w] - M));
  const float inv = 1.0f / tot;
  for (long long i = threadIdx.x; i < S; i += blockDim.x) dst[i] = __expf(src[i] - M) * inv;
}

// softargmax_on_softmax_map (ynet.py:588-600): (sum p*x, sum p*y), no epsilon.
__global__ void __launch_bounds__(512) expectation2d_kernel(const float* __restrict__ p, int H, int W,
                                                            float* __restrict__ out) {
  const int S = H * W;
  const float* src = p + (size_t)blockIdx.x * S;
  float sx = 0.f, sy = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float v = src[i];
    const int y = i / W;
    sx = fmaf(v, (float)(i - y * W), sx);
    sy = fmaf(v, (float)y, sy);
  }
  sx = warp_sum(sx);
  sy = warp_sum(sy);
  __shared__ float shx[16], shy[16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    shx[warp] = sx;
    shy[warp] = sy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ax = 0.f, ay = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      ax += shx[w];
      ay += shy[w];
    }
    out[2 * blockIdx.x + 0] = ax;
    out[2 * blockIdx.x + 1] = ay;
  }
}

struct ChanSel {
  int ch[32];
};

// out[b,k,:] = sigmoid(logits[b, ch[k], :] / T)
__global__ void __launch_bounds__(256)
sigmoid_select_kernel(const float* __restrict__ logits, int C, long long S, ChanSel sel, int n_ch, float T,
                      float* __restrict__ out) {
  const int b = blockIdx.z, k = blockIdx.y;
  const float* src = logits + ((size_t)b * C + sel.ch[k]) * S;
  float* dst = out + ((size_t)b * n_ch + k) * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) {
    const float z = src[i] / T;
    dst[i] = 1.0f / (1.0f + __expf(-z));
  }
}

static int softargmax_splits(int rows, int S) {
  // enough CTAs for ~4 waves, but never slices shorter than 8 K elements
  const int target = 4 * sm_count();
  int splits = ceil_div(target, max(rows, 1));
  const int max_splits = max(1, S / 8192);
  splits = max(1, min(splits, max_splits));
  return min(splits, 64);
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int64_t ynet_softargmax2d_workspace_bytes(int32_t rows, int32_t H, int32_t W) {
  return (int64_t)rows * 64 * sizeof(SoftPartial);
}

int ynet_softargmax2d(const float* x, int32_t rows, int64_t row_stride, int32_t H, int32_t W, float* out,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  YNET_CHECK_ARG(x && out, "null pointer");
  YNET_CHECK_ARG(rows >= 0 && H > 0 && W > 0 && row_stride >= (int64_t)H * W, "bad shape");
  if (rows == 0) return YNET_OK;
  const int S = H * W;
  const int splits = softargmax_splits(rows, S);
  if (workspace == nullptr || workspace_bytes < (int64_t)rows * splits * (int64_t)sizeof(SoftPartial)) {
    set_error("ynet_softargmax2d: workspace too small");
    return YNET_E_WORKSPACE;
  }
  SoftPartial* part = reinterpret_cast<SoftPartial*>(workspace);
  const bool vec = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (row_stride % 4 == 0);
  for (int r0 = 0; r0 < rows; r0 += 65535) {
    const int rr = min(65535, rows - r0);
    dim3 grid(splits, rr);
    if (vec)
      softargmax_partial_kernel<true><<<grid, kSoftThreads, 0, as_stream(stream)>>>(
          x + (size_t)r0 * row_stride, row_stride, H, W, splits, part + (size_t)r0 * splits);
    else
      softargmax_partial_kernel<false><<<grid, kSoftThreads, 0, as_stream(stream)>>>(
          x + (size_t)r0 * row_stride, row_stride, H, W, splits, part + (size_t)r0 * splits);
    YNET_LAUNCH_CHECK();
  }
  softargmax_finalize_kernel<<<ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(part, rows, splits, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_spatial_softmax(const float* x, int32_t rows, int64_t S, float* out, void* stream) {
  YNET_CHECK_ARG(x && out && rows >= 0 && S > 0, "bad argument");
  if (rows == 0) return YNET_OK;
  spatial_softmax_kernel<<<rows, 1024, 0, as_stream(stream)>>>(x, S, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_expectation2d(const float* p, int32_t rows, int32_t H, int32_t W, float* out, void* stream) {
  YNET_CHECK_ARG(p && out && rows >= 0 && H > 0 && W > 0, "bad argument");
  if (rows == 0) return YNET_OK;
  expectation2d_kernel<<<rows, 512, 0, as_stream(stream)>>>(p, H, W, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_sigmoid_select(const float* logits, int32_t B, int32_t C, int64_t S, const int32_t* ch_host, int32_t n_ch,
                        float temperature, float* out, void* stream) {
  YNET_CHECK_ARG(logits && out && ch_host, "null pointer");
  YNET_CHECK_ARG(B >= 0 && C > 0 && S > 0 && n_ch > 0 && n_ch <= 32, "bad shape (n_ch <= 32)");
  YNET_CHECK_ARG(B <= 65535, "B must be <= 65535 per call");
  if (B == 0) return YNET_OK;
  ChanSel sel;
  for (int i = 0; i < n_ch; ++i) {
    YNET_CHECK_ARG(ch_host[i] >= 0 && ch_host[i] < C, "channel index out of range");
    sel.ch[i] = ch_host[i];
  }
  dim3 grid((unsigned)tmin<long long>(ceil_div<long long>(S, 256 * 4), 1024), n_ch, B);
  sigmoid_select_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, C, S, sel, n_ch, temperature, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

}  // extern "C"
