// a1/a3/a9: template creation, trajectory-heatmap rasterisation, waypoint pyramid.
// HBM-bound kernels: coalesced 128-bit stores, template reads served from L2 (4.4-7.7 MB << 126 MB).
#include "common.cuh"

namespace ynet {

// ---- a3: gather rasteriser ---------------------------------------------------------------------
// out[n,i,j] = tmpl[mid_y - y_n + i][mid_x - x_n + j].  One float4 of output per thread-iteration;
// the 4 source floats are contiguous but not 16 B aligned (x_n is arbitrary) -> scalar L2 loads.
template <int VEC>
__global__ void __launch_bounds__(256)
rasterize_gather_kernel(const float* __restrict__ tmpl, int th, int tw, const float* __restrict__ coords,
                        float* __restrict__ out, int H, int W, int* __restrict__ oob) {
  const int n = blockIdx.y;
  const int x = __float2int_rn(coords[2 * n + 0]);  // round half to even == np.round
  const int y = __float2int_rn(coords[2 * n + 1]);
  int yl = th / 2 - y;
  int xl = tw / 2 - x;
  const bool bad = (yl < 0) | (xl < 0) | (yl + H > th) | (xl + W > tw);
  if (bad && oob != nullptr && threadIdx.x == 0 && blockIdx.x == 0) atomicExch(oob, 1);
  const int wv = W / VEC;
  const int total = H * wv;
  float* o = out + (size_t)n * H * W;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int i = t / wv;
    const int j = (t - i * wv) * VEC;
    int sy = yl + i;
    sy = min(max(sy, 0), th - 1);
    const float* src = tmpl + (size_t)sy * tw;
    float v[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      int sx = xl + j + k;
      sx = min(max(sx, 0), tw - 1);
      v[k] = __ldg(src + sx);
    }
    if (VEC == 4) {
      st_stream(reinterpret_cast<float4*>(o + (size_t)i * W + j), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[(size_t)i * W + j + k] = v[k];
    }
  }
}

// ---- a3 (analytic) / a1: distance map in fp64, bit-identical to the numpy template ------------------
__global__ void __launch_bounds__(256)
rasterize_analytic_kernel(int mid, const float* __restrict__ coords, float* __restrict__ out, int H, int W) {
  const int n = blockIdx.y;
  const int x = __float2int_rn(coords[2 * n + 0]);
  const int y = __float2int_rn(coords[2 * n + 1]);
  const double mx = sqrt((double)(2LL * mid * mid));
  const int total = H * W;
  float* o = out + (size_t)n * total;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int i = t / W, j = t - i * W;
    const int di = i - y, dj = j - x;
    const double s = (double)((long long)di * di + (long long)dj * dj);
    o[t] = (float)(__ddiv_rn(__dsqrt_rn(s), mx) * 2.0);
  }
}

__global__ void __launch_bounds__(256) dist_template_kernel(int size, float* __restrict__ out) {
  const int mid = size / 2;
  const double mx = sqrt((double)(2LL * mid * mid));
  const long long total = (long long)size * size;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / size), j = (int)(t - (long long)i * size);
    const int di = i - mid, dj = j - mid;
    const double s = (double)((long long)di * di + (long long)dj * dj);
    out[t] = (float)(__ddiv_rn(__dsqrt_rn(s), mx) * 2.0);
  }
}

// ---- a9: AvgPool2d(2^i) pyramid, one read of the full-resolution map ----------------------------
struct PyramidOuts {
  float* p[5];
};

__global__ void __launch_bounds__(256)
avgpool_pyramid_kernel(const float* __restrict__ in, int H, int W, int n_levels, PyramidOuts outs) {
  __shared__ float s0[32][33];
  __shared__ float s1[16][17];
  __shared__ float s2[8][9];
  __shared__ float s3[4][5];
  __shared__ float s4[2][3];
  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * 32, tx0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = in + (size_t)n * H * W;
#pragma unroll
  for (int r = 0; r < 4; ++r) s0[ty + 8 * r][tx] = src[(size_t)(ty0 + ty + 8 * r) * W + tx0 + tx];
  __syncthreads();
  const int t = threadIdx.x;
  {  // level 1: 16 x 16
    const int y = t >> 4, x = t & 15;
    const float v = 0.25f * ((s0[2 * y][2 * x] + s0[2 * y][2 * x + 1]) + (s0[2 * y + 1][2 * x] + s0[2 * y + 1][2 * x + 1]));
    s1[y][x] = v;
    const int h = H >> 1, w = W >> 1;
    outs.p[0][(size_t)n * h * w + (size_t)((ty0 >> 1) + y) * w + (tx0 >> 1) + x] = v;
  }
  if (n_levels <= 2) return;
  __syncthreads();
  if (t < 64) {
    const int y = t >> 3, x = t & 7;
    const float v = 0.25f * ((s1[2 * y][2 * x] + s1[2 * y][2 * x + 1]) + (s1[2 * y + 1][2 * x] + s1[2 * y + 1][2 * x + 1]));
    s2[y][x] = v;
    const int h = H >> 2, w = W >> 2;
    outs.p[1][(size_t)n * h * w + (size_t)((ty0 >> 2) + y) * w + (tx0 >> 2) + x] = v;
  }
  if (n_levels <= 3) return;
  __syncthreads();
  if (t < 16) {
    const int y = t >> 2, x = t & 3;
    const float v = 0.25f * ((s2[2 * y][2 * x] + s2[2 * y][2 * x + 1]) + (s2[2 * y + 1][2 * x] + s2[2 * y + 1][2 * x + 1]));
    s3[y][x] = v;
    const int h = H >> 3, w = W >> 3;
    outs.p[2][(size_t)n * h * w + (size_t)((ty0 >> 3) + y) * w + (tx0 >> 3) + x] = v;
  }
  if (n_levels <= 4) return;
  __syncthreads();
  if (t < 4) {
    const int y = t >> 1, x = t & 1;
    const float v = 0.25f * ((s3[2 * y][2 * x] + s3[2 * y][2 * x + 1]) + (s3[2 * y + 1][2 * x] + s3[2 * y + 1][2 * x + 1]));
    s4[y][x] = v;
    const int h = H >> 4, w = W >> 4;
    outs.p[3][(size_t)n * h * w + (size_t)((ty0 >> 4) + y) * w + (tx0 >> 4) + x] = v;
  }
  if (n_levels <= 5) return;
  __syncthreads();
  if (t == 0) {
    const float v = 0.25f * ((s4[0][0] + s4[0][1]) + (s4[1][0] + s4[1][1]));
    const int h = H >> 5, w = W >> 5;
    outs.p[4][(size_t)n * h * w + (size_t)(ty0 >> 5) * w + (tx0 >> 5)] = v;
  }
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int ynet_rasterize_patches(const float* tmpl, int32_t th, int32_t tw, const float* coords, int32_t n, float* out,
                           int32_t H, int32_t W, int32_t* oob_flag, void* stream) {
  YNET_CHECK_ARG(n >= 0 && H > 0 && W > 0 && th >= H && tw >= W, "bad shape (template smaller than window?)");
  if (n == 0) return YNET_OK;
  YNET_CHECK_ARG(tmpl && coords && out, "null pointer");
  const bool vec = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const int work = vec ? H * (W / 4) : H * W;
  int gx = ceil_div(work, 256 * 4);
  // keep >= ~2 waves of CTAs when n is small
  const int min_ctas = 2 * sm_count();
  if ((long long)gx * n < min_ctas) gx = (int)min((long long)ceil_div(work, 256), (long long)ceil_div(min_ctas, n));
  gx = max(gx, 1);
  for (int n0 = 0; n0 < n; n0 += 65535) {
    const int nn = min(65535, n - n0);
    dim3 grid(gx, nn);
    const float* c = coords + 2 * (size_t)n0;
    float* o = out + (size_t)n0 * H * W;
    if (vec)
      rasterize_gather_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(tmpl, th, tw, c, o, H, W, oob_flag);
    else
      rasterize_gather_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(tmpl, th, tw, c, o, H, W, oob_flag);
    YNET_LAUNCH_CHECK();
  }
  return YNET_OK;
}

int ynet_rasterize_dist_analytic(int32_t tmpl_size, const float* coords, int32_t n, float* out, int32_t H, int32_t W,
                                 void* stream) {
  YNET_CHECK_ARG(coords && out, "null pointer");
  YNET_CHECK_ARG(n >= 0 && H > 0 && W > 0 && tmpl_size > 0, "bad shape");
  if (n == 0) return YNET_OK;
  for (int n0 = 0; n0 < n; n0 += 65535) {
    const int nn = min(65535, n - n0);
    dim3 grid(max(1, min(ceil_div(H * W, 256 * 2), 4096)), nn);
    rasterize_analytic_kernel<<<grid, 256, 0, as_stream(stream)>>>(tmpl_size / 2, coords + 2 * (size_t)n0,
                                                                    out + (size_t)n0 * H * W, H, W);
    YNET_LAUNCH_CHECK();
  }
  return YNET_OK;
}

int ynet_create_dist_template(int32_t size, float* out, void* stream) {
  YNET_CHECK_ARG(out && size > 0, "bad argument");
  dist_template_kernel<<<sm_count() * 8, 256, 0, as_stream(stream)>>>(size, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_avgpool_pyramid(const float* in, int32_t n, int32_t H, int32_t W, int32_t n_levels, float* const* outs_host,
                         void* stream) {
  YNET_CHECK_ARG(in && outs_host, "null pointer");
  YNET_CHECK_ARG(n_levels >= 2 && n_levels <= 6, "n_levels must be in [2, 6]");
  if (H % 32 != 0 || W % 32 != 0) {
    set_error("ynet_avgpool_pyramid: H and W must be multiples of 32 (trainer.py:60,581 pads to 2^len(enc))");
    return YNET_E_UNSUPPORTED;
  }
  if (n == 0) return YNET_OK;
  for (int n0 = 0; n0 < n; n0 += 65535) {
    const int nn = min(65535, n - n0);
    PyramidOuts o;
    for (int i = 0; i < 5; ++i)
      o.p[i] = (i < n_levels - 1) ? outs_host[i] + (size_t)n0 * (H >> (i + 1)) * (W >> (i + 1)) : nullptr;
    dim3 grid(W / 32, H / 32, nn);
    avgpool_pyramid_kernel<<<grid, 256, 0, as_stream(stream)>>>(in + (size_t)n0 * H * W, H, W, n_levels, o);
    YNET_LAUNCH_CHECK();
  }
  return YNET_OK;
}

}  // extern "C"
