// a1/a3/a9: template creation, trajectory-heatmap rasterisation, waypoint pyramid.
// HBM-bound kernels: coalesced 128-bit stores, template reads served from L2 (4.4-7.7 MB << 126 MB).
#include <cuda_bf16.h>

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

// ---- a3 + a9 fused for the tensor-core engine: waypoint maps -> bf16 C8 pyramid ------------------------------
// One CTA = one 32x32 full-resolution block of one image: gathers the n_ch (<= 8) template windows into shared
// memory, then writes every pyramid level as C8 planes (16 B = 8 bf16 channels per pixel; channels >= n_ch zero).
// Replaces rasterize_gather + avgpool_pyramid + n_levels pack_c8 launches (and their float32 round trips).
struct PyramidC8Outs {
  uint4* p[6];
};

__device__ __forceinline__ uint4 pack8_bf16(const float* f) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c);
  o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}

template <int NCH>
__global__ void __launch_bounds__(256)
wp_pyramid_c8_kernel(const float* __restrict__ tmpl, int th, int tw, const float* __restrict__ coords, int H, int W,
                     int n_levels, int chunks, int write_pad, PyramidC8Outs outs, int* __restrict__ oob) {
  __shared__ float s0[NCH][32][33];
  __shared__ float s1[NCH][16][17];
  __shared__ float s2[NCH][8][9];
  __shared__ float s3[NCH][4][5];
  __shared__ float s4[NCH][2][3];
  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * 32, tx0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int x = __float2int_rn(coords[2 * (n * NCH + c) + 0]);  // round half to even == np.round
    const int y = __float2int_rn(coords[2 * (n * NCH + c) + 1]);
    const int yl = th / 2 - y, xl = tw / 2 - x;
    const bool bad = (yl < 0) | (xl < 0) | (yl + H > th) | (xl + W > tw);
    if (bad && oob != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) atomicExch(oob, 1);
    const int sx = min(max(xl + tx0 + tx, 0), tw - 1);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int sy = min(max(yl + ty0 + ty + 8 * r, 0), th - 1);
      s0[c][ty + 8 * r][tx] = __ldg(tmpl + (size_t)sy * tw + sx);
    }
  }
  __syncthreads();
  if (outs.p[0] != nullptr) {  // level 0: 32 x 32, 4 pixels per thread (null = level not wanted: the row-marching conv
                               // gathers its waypoint planes itself, rowconv_tc.cu)
    uint4* o = outs.p[0] + (size_t)n * chunks * H * W;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] = (c < NCH) ? s0[c < NCH ? c : 0][ty + 8 * r][tx] : 0.f;
      const size_t pix = (size_t)(ty0 + ty + 8 * r) * W + tx0 + tx;
      o[pix] = pack8_bf16(f);
      if (write_pad)
        for (int k = 1; k < chunks; ++k) o[(size_t)k * H * W + pix] = zero;
    }
  }
  const int t = threadIdx.x;
  auto reduce = [&](auto& src, auto& dst, int side, int lvl) {   // side = output block edge at this level
    if (t < side * side) {
      const int y = t / side, x = t - y * side;
      float f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < NCH) {
          const int cc = c < NCH ? c : 0;
          const float v = 0.25f * ((src[cc][2 * y][2 * x] + src[cc][2 * y][2 * x + 1]) +
                                   (src[cc][2 * y + 1][2 * x] + src[cc][2 * y + 1][2 * x + 1]));
          dst[cc][y][x] = v;
          f[c] = v;
        } else {
          f[c] = 0.f;
        }
      }
      const int h = H >> lvl, w = W >> lvl;
      const size_t pix = (size_t)((ty0 >> lvl) + y) * w + (tx0 >> lvl) + x;
      if (outs.p[lvl] != nullptr) {
        uint4* o = outs.p[lvl] + (size_t)n * chunks * h * w;
        o[pix] = pack8_bf16(f);
        if (write_pad)
          for (int k = 1; k < chunks; ++k) o[(size_t)k * h * w + pix] = zero;
      }
    }
  };
  if (n_levels <= 1) return;
  reduce(s0, s1, 16, 1);
  if (n_levels <= 2) return;
  __syncthreads();
  reduce(s1, s2, 8, 2);
  if (n_levels <= 3) return;
  __syncthreads();
  reduce(s2, s3, 4, 3);
  if (n_levels <= 4) return;
  __syncthreads();
  reduce(s3, s4, 2, 4);
  if (n_levels <= 5) return;
  __syncthreads();
  if (t == 0) {
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      f[c] = (c < NCH) ? 0.25f * ((s4[c < NCH ? c : 0][0][0] + s4[c < NCH ? c : 0][0][1]) +
                                  (s4[c < NCH ? c : 0][1][0] + s4[c < NCH ? c : 0][1][1]))
                       : 0.f;
    const int h = H >> 5, w = W >> 5;
    const size_t pix = (size_t)(ty0 >> 5) * w + (tx0 >> 5);
    uint4* o = outs.p[5] + (size_t)n * chunks * h * w;
    o[pix] = pack8_bf16(f);
    if (write_pad)
      for (int k = 1; k < chunks; ++k) o[(size_t)k * h * w + pix] = zero;
  }
}


// ---- the distance template as bf16 C8 planes (16 B per template pixel: channel 0 = value, 1-7 zero), so that the row-
// marching conv's TMA can load a waypoint channel's window straight into its K-major operand chunk (rowconv_tc.cu):
// level 0 as it is, level 1 as the four parity planes of the 2x2 average (same float32 arithmetic as the pyramid kernels)
__global__ void __launch_bounds__(256)
wp_template_c8_kernel(const float* __restrict__ tmpl, int th, int tw, uint4* __restrict__ l0, uint4* __restrict__ l1) {
  const long long S = (long long)th * tw, S1 = (long long)(th / 2) * (tw / 2);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < S + 4 * S1; t += (long long)gridDim.x * blockDim.x) {
    float v;
    uint4* dst;
    if (t < S) {
      v = tmpl[t];
      dst = l0 + t;
    } else {
      const long long r = t - S;
      const int par = (int)(r / S1);
      const long long q = r - par * S1;
      const int i = (int)(q / (tw / 2)), j = (int)(q - (long long)i * (tw / 2));
      const int y0 = min(2 * i + (par >> 1), th - 1), y1 = min(y0 + 1, th - 1);
      const int x0 = min(2 * j + (par & 1), tw - 1), x1 = min(x0 + 1, tw - 1);
      v = 0.25f * ((tmpl[(size_t)y0 * tw + x0] + tmpl[(size_t)y0 * tw + x1]) +
                   (tmpl[(size_t)y1 * tw + x0] + tmpl[(size_t)y1 * tw + x1]));
      dst = l1 + r;
    }
    const __nv_bfloat162 b = __floats2bfloat162_rn(v, 0.f);
    *dst = make_uint4(*reinterpret_cast<const uint32_t*>(&b), 0u, 0u, 0u);
  }
}

// ---- the same pyramid with 2x2-NEIGHBOURHOOD planes at the finest levels ---------------------------------------
// The n_wp <= 2 waypoint channels leave six of the eight channels of the plane empty, and the conv spends nine MMAs per
// tile on that K block.  Levels l < quad_levels store, at pixel (y, x), channel (dy*2 + dx) * NCH + c = map_c[y+dy][x+dx]
// (dy, dx in {0, 1}; zero outside the image = the conv's zero padding): the 3x3 window is then covered by the FOUR taps
// anchored at (-1,-1) (-1,0) (0,-1) (0,0) (ynet_tc_src.tap_mask = YNET_TC_TAPS_QUAD) at the same 16 B per pixel.
// Pooled levels use the same arithmetic as wp_pyramid_c8_kernel; levels >= quad_levels keep the plain layout.
template <int NCH>
__global__ void __launch_bounds__(256)
wp_pyramid_quad_c8_kernel(const float* __restrict__ tmpl, int th, int tw, const float* __restrict__ coords, int H, int W,
                          int n_levels, int quad_levels, PyramidC8Outs outs, int* __restrict__ oob) {
  static_assert(NCH * 4 <= 8, "2x2 neighbourhood of NCH channels must fit one 8-channel plane");
  __shared__ float s0[NCH][34][35];
  __shared__ float s1[NCH][17][18];
  __shared__ float s2[NCH][8][9];
  __shared__ float s3[NCH][4][5];
  __shared__ float s4[NCH][2][3];
  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * 32, tx0 = blockIdx.x * 32;
  const int t = threadIdx.x;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int x = __float2int_rn(coords[2 * (n * NCH + c) + 0]);  // round half to even == np.round
    const int y = __float2int_rn(coords[2 * (n * NCH + c) + 1]);
    const int yl = th / 2 - y, xl = tw / 2 - x;
    const bool bad = (yl < 0) | (xl < 0) | (yl + H > th) | (xl + W > tw);
    if (bad && oob != nullptr && t == 0 && blockIdx.x == 0 && blockIdx.y == 0) atomicExch(oob, 1);
    // 34 x 34 region (block + two pixels of apron): a warp per row, lanes 0-1 also fetch the two apron columns
    // all ten loads of a thread are issued before the first shared-memory store (one L2 round trip per block)
    const int lane = t & 31, wrp = t >> 5;
    float v[5], va[5];
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int ry = wrp + 8 * it, gy = ty0 + ry;
      const int sy = min(max(yl + gy, 0), th - 1);
      const float* row = tmpl + (size_t)sy * tw;
      const bool rok = ry < 34 && gy < H;
      const int gx = tx0 + 32 + lane;
      v[it] = rok ? __ldg(row + min(max(xl + tx0 + lane, 0), tw - 1)) : 0.f;
      va[it] = (rok && lane < 2 && gx < W) ? __ldg(row + min(max(xl + gx, 0), tw - 1)) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int ry = wrp + 8 * it;
      if (ry < 34) {
        s0[c][ry][lane] = v[it];
        if (lane < 2) s0[c][ry][32 + lane] = va[it];
      }
    }
  }
  __syncthreads();
  {  // level 0: 32 x 32, 4 pixels per thread
    const int tx = t & 31, ty = t >> 5;
    uint4* o = outs.p[0] + (size_t)n * H * W;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int y = ty + 8 * r;
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.f;
      if (quad_levels > 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int c = 0; c < NCH; ++c) f[q * NCH + c] = s0[c][y + (q >> 1)][tx + (q & 1)];
      } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c) f[c] = s0[c][y][tx];
      }
      o[(size_t)(ty0 + y) * W + tx0 + tx] = pack8_bf16(f);
    }
  }
  if (n_levels <= 1) return;
  // level 1 incl. one pooled pixel of apron: 17 x 17 per channel; threads 0..255 take the 16 x 16 core, the next
  // 33 * NCH (< 256) the apron row and column
  {
    const int py = t >> 4, px = t & 15;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      s1[c][py][px] = 0.25f * ((s0[c][2 * py][2 * px] + s0[c][2 * py][2 * px + 1]) +
                               (s0[c][2 * py + 1][2 * px] + s0[c][2 * py + 1][2 * px + 1]));
    if (t < 33 * NCH) {
      const int c = t / 33, r = t - c * 33;
      const int ay = r < 17 ? 16 : r - 17, ax = r < 17 ? r : 16;      // row 16 (17 entries), then column 16 (16 entries)
      s1[c][ay][ax] = 0.25f * ((s0[c][2 * ay][2 * ax] + s0[c][2 * ay][2 * ax + 1]) +
                               (s0[c][2 * ay + 1][2 * ax] + s0[c][2 * ay + 1][2 * ax + 1]));
    }
  }
  __syncthreads();
  {
    const int y = t >> 4, x = t & 15;
    const int h = H >> 1, w = W >> 1;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 0.f;
    if (quad_levels > 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < NCH; ++c) f[q * NCH + c] = s1[c][y + (q >> 1)][x + (q & 1)];
    } else {
#pragma unroll
      for (int c = 0; c < NCH; ++c) f[c] = s1[c][y][x];
    }
    outs.p[1][(size_t)n * h * w + (size_t)((ty0 >> 1) + y) * w + (tx0 >> 1) + x] = pack8_bf16(f);
  }
  auto reduce = [&](auto& src, auto& dst, int side, int lvl) {   // side = output block edge at this level
    if (t < side * side) {
      const int y = t / side, x = t - y * side;
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float v = 0.25f * ((src[c][2 * y][2 * x] + src[c][2 * y][2 * x + 1]) +
                                 (src[c][2 * y + 1][2 * x] + src[c][2 * y + 1][2 * x + 1]));
        dst[c][y][x] = v;
        f[c] = v;
      }
      const int h = H >> lvl, w = W >> lvl;
      outs.p[lvl][(size_t)n * h * w + (size_t)((ty0 >> lvl) + y) * w + (tx0 >> lvl) + x] = pack8_bf16(f);
    }
  };
  if (n_levels <= 2) return;
  reduce(s1, s2, 8, 2);
  if (n_levels <= 3) return;
  __syncthreads();
  reduce(s2, s3, 4, 3);
  if (n_levels <= 4) return;
  __syncthreads();
  reduce(s3, s4, 2, 4);
  if (n_levels <= 5) return;
  __syncthreads();
  if (t == 0) {
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) f[c] = 0.25f * ((s4[c][0][0] + s4[c][0][1]) + (s4[c][1][0] + s4[c][1][1]));
    const int h = H >> 5, w = W >> 5;
    outs.p[5][(size_t)n * h * w + (size_t)(ty0 >> 5) * w + (tx0 >> 5)] = pack8_bf16(f);
  }
}


// ---- a3 + a9 in im2col form for the tensor-core engine -------------------------------------------------------
// The n_wp (1-2) waypoint channels of a trajectory-decoder input occupy a whole 16-channel K block of the conv: nine
// MMAs per tile for two real channels.  Written as im2col instead -- channel c * 9 + kh * 3 + kw of pixel (y, x) =
// map_c[y + kh - 1][x + kw - 1], 0 outside the image (the conv's zero padding) -- the same contribution is ONE 1x1
// (centre-tap) K block per 16 im2col channels.  Level L in {0, 1}: the map is the 2^L average pool of the rasterised
// distance map, pooled with the same arithmetic as wp_pyramid_c8_kernel.
template <int NCH, int L>
__global__ void __launch_bounds__(256)
wp_im2col_c8_kernel(const float* __restrict__ tmpl, int th, int tw, const float* __restrict__ coords, int H, int W,
                    int chunks, uint4* __restrict__ out) {
  constexpr int S = 1 << L;             // pooling factor
  constexpr int OB = 32 >> L;           // output block edge
  constexpr int R = 32 + 2 * S;         // level-0 region edge (block + one pooled pixel of halo)
  constexpr int RP = OB + 2;            // pooled region edge
  __shared__ float s0[NCH][R][R + 1];
  __shared__ float sp[NCH][RP][RP + 1];
  const int n = blockIdx.z;
  const int by0 = blockIdx.y * 32, bx0 = blockIdx.x * 32;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int x = __float2int_rn(coords[2 * (n * NCH + c) + 0]);  // round half to even == np.round
    const int y = __float2int_rn(coords[2 * (n * NCH + c) + 1]);
    const int yl = th / 2 - y, xl = tw / 2 - x;
    for (int idx = threadIdx.x; idx < R * R; idx += 256) {
      const int ry = idx / R, rx = idx - ry * R;
      const int gy = by0 - S + ry, gx = bx0 - S + rx;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const int sy = min(max(yl + gy, 0), th - 1), sx = min(max(xl + gx, 0), tw - 1);
        v = __ldg(tmpl + (size_t)sy * tw + sx);
      }
      s0[c][ry][rx] = v;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < NCH * RP * RP; idx += 256) {
    const int c = idx / (RP * RP), r = idx - c * RP * RP;
    const int py = r / RP, px = r - py * RP;
    float v;
    if (L == 0)
      v = s0[c][py][px];
    else
      v = 0.25f * ((s0[c][2 * py][2 * px] + s0[c][2 * py][2 * px + 1]) + (s0[c][2 * py + 1][2 * px] + s0[c][2 * py + 1][2 * px + 1]));
    sp[c][py][px] = v;
  }
  __syncthreads();
  const int h = H >> L, w = W >> L;
  uint4* o = out + (size_t)n * chunks * h * w;
  for (int idx = threadIdx.x; idx < chunks * OB * OB; idx += 256) {
    const int k = idx / (OB * OB), r = idx - k * OB * OB;
    const int oy = r / OB, ox = r - oy * OB;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = 8 * k + j;
      const int c = ch / 9, tap = ch - 9 * c;
      f[j] = (c < NCH) ? sp[c < NCH ? c : 0][oy + tap / 3][ox + tap % 3] : 0.f;
    }
    o[(size_t)k * h * w + (size_t)((by0 >> L) + oy) * w + (bx0 >> L) + ox] = pack8_bf16(f);
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

int ynet_tc_rasterize_pyramid_c8(const float* tmpl, int32_t th, int32_t tw, const float* coords, int32_t n_img,
                                 int32_t n_ch, int32_t H, int32_t W, int32_t n_levels, void* const* outs_host,
                                 int32_t C_pad, int32_t write_pad, int32_t quad_levels, int32_t* oob_flag, void* stream) {
  YNET_CHECK_ARG(n_img >= 0 && n_ch >= 1 && n_ch <= 8 && H > 0 && W > 0 && th >= H && tw >= W, "bad shape (n_ch <= 8)");
  YNET_CHECK_ARG(quad_levels >= 0 && quad_levels <= 2 && (quad_levels == 0 || (n_ch <= 2 && C_pad == 8)),
                 "quad_levels in [0, 2]; 2x2-neighbourhood planes need n_ch <= 2 and C_pad == 8");
  YNET_CHECK_ARG(n_levels >= 1 && n_levels <= 6 && C_pad >= 8 && C_pad % 8 == 0, "n_levels in [1, 6], C_pad % 8 == 0");
  if (H % 32 != 0 || W % 32 != 0) {
    set_error("ynet_tc_rasterize_pyramid_c8: H and W must be multiples of 32 (trainer.py:60,581)");
    return YNET_E_UNSUPPORTED;
  }
  if (n_img == 0) return YNET_OK;
  YNET_CHECK_ARG(tmpl && coords && outs_host, "null pointer");
  const int chunks = C_pad / 8;
  for (int n0 = 0; n0 < n_img; n0 += 65535) {
    const int nn = min(65535, n_img - n0);
    PyramidC8Outs o;
    for (int i = 0; i < 6; ++i) {
      o.p[i] = nullptr;
      if (i < n_levels) {
        // a null level is skipped (plain layout only): ynet_tc_rowconv3x3_wp gathers the finest levels itself
        YNET_CHECK_ARG(outs_host[i] != nullptr || (quad_levels == 0 && i < 5), "null output level");
        YNET_CHECK_ALIGN(outs_host[i], 16);
        if (outs_host[i] != nullptr)
          o.p[i] = reinterpret_cast<uint4*>(outs_host[i]) + (size_t)n0 * chunks * (H >> i) * (W >> i);
      }
    }
    dim3 grid(W / 32, H / 32, nn);
    const float* c = coords + 2 * (size_t)n0 * n_ch;
    cudaStream_t st = as_stream(stream);
    if (quad_levels > 0) {
      if (n_ch == 1)
        wp_pyramid_quad_c8_kernel<1><<<grid, 256, 0, st>>>(tmpl, th, tw, c, H, W, n_levels, quad_levels, o, oob_flag);
      else
        wp_pyramid_quad_c8_kernel<2><<<grid, 256, 0, st>>>(tmpl, th, tw, c, H, W, n_levels, quad_levels, o, oob_flag);
      YNET_LAUNCH_CHECK();
      continue;
    }
#define YNET_WP_CASE(K)                                                                                              \
  case K:                                                                                                            \
    wp_pyramid_c8_kernel<K><<<grid, 256, 0, st>>>(tmpl, th, tw, c, H, W, n_levels, chunks, write_pad, o, oob_flag); \
    break;
    switch (n_ch) {
      YNET_WP_CASE(1) YNET_WP_CASE(2) YNET_WP_CASE(3) YNET_WP_CASE(4) YNET_WP_CASE(5) YNET_WP_CASE(6) YNET_WP_CASE(7)
      YNET_WP_CASE(8)
    }
#undef YNET_WP_CASE
    YNET_LAUNCH_CHECK();
  }
  return YNET_OK;
}

int ynet_tc_wp_template_c8(const float* tmpl, int32_t th, int32_t tw, void* out_l0, void* out_l1, void* stream) {
  YNET_CHECK_ARG(tmpl && out_l0 && out_l1, "null pointer");
  YNET_CHECK_ARG(th >= 2 && tw >= 2 && th % 2 == 0 && tw % 2 == 0, "template size must be even");
  YNET_CHECK_ALIGN(out_l0, 16);
  YNET_CHECK_ALIGN(out_l1, 16);
  const long long total = 2LL * th * tw;
  const long long blocks = (total + 255) / 256;
  wp_template_c8_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, as_stream(stream)>>>(
      tmpl, th, tw, reinterpret_cast<uint4*>(out_l0), reinterpret_cast<uint4*>(out_l1));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_rasterize_im2col_c8(const float* tmpl, int32_t th, int32_t tw, const float* coords, int32_t n_img, int32_t n_ch,
                                int32_t H, int32_t W, int32_t level, void* out_c8, int32_t C_pad, void* stream) {
  YNET_CHECK_ARG(n_img >= 0 && n_ch >= 1 && n_ch <= 3 && H > 0 && W > 0 && th >= H && tw >= W, "bad shape (n_ch <= 3)");
  YNET_CHECK_ARG((level == 0 || level == 1) && C_pad >= 9 * n_ch && C_pad % 16 == 0, "level in {0, 1}, C_pad >= 9 n_ch, % 16");
  if (H % 32 != 0 || W % 32 != 0) {
    set_error("ynet_tc_rasterize_im2col_c8: H and W must be multiples of 32 (trainer.py:60,581)");
    return YNET_E_UNSUPPORTED;
  }
  if (n_img == 0) return YNET_OK;
  YNET_CHECK_ARG(tmpl && coords && out_c8, "null pointer");
  YNET_CHECK_ALIGN(out_c8, 16);
  const int chunks = C_pad / 8;
  cudaStream_t st = as_stream(stream);
  for (int n0 = 0; n0 < n_img; n0 += 65535) {
    const int nn = min(65535, n_img - n0);
    dim3 grid(W / 32, H / 32, nn);
    const float* c = coords + 2 * (size_t)n0 * n_ch;
    uint4* o = reinterpret_cast<uint4*>(out_c8) + (size_t)n0 * chunks * (H >> level) * (W >> level);
#define YNET_I2C_CASE(K, LV)                                                                   \
  if (n_ch == K && level == LV) wp_im2col_c8_kernel<K, LV><<<grid, 256, 0, st>>>(tmpl, th, tw, c, H, W, chunks, o);
    YNET_I2C_CASE(1, 0) YNET_I2C_CASE(2, 0) YNET_I2C_CASE(3, 0) YNET_I2C_CASE(1, 1) YNET_I2C_CASE(2, 1) YNET_I2C_CASE(3, 1)
#undef YNET_I2C_CASE
    YNET_LAUNCH_CHECK();
  }
  return YNET_OK;
}

}  // extern "C"
