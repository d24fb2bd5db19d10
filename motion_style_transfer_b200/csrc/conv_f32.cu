// a4-a8: reference-grade float32 network engine (CUDA cores, fp32 FMA, fp32 accumulate).
//
// This is the <=1e-3 parity path and the on-device checker for the tensor-core engine (conv_tc.cu).
// One kernel = one 3x3 conv of the reference (ynet.py:192-211, 419-447) with its producer ops fused
// into the tile loader: channel concat (torch.cat), 2x2 max-pool (nn.MaxPool2d) and bilinear x2
// (F.interpolate, align_corners=False) never touch HBM as separate tensors.
#include <float.h>

#include "common.cuh"

namespace ynet {

struct SrcDev {
  const float* ptr;
  int channels;
  int mode;
  long long batch_stride;
  long long batch_mod;
};

struct Conv3Params {
  SrcDev src[YNET_MAX_SOURCES];
  int n_src, C_in, N, H, W, C_out, relu;
  const float* weight;  // packed [C_in][9][C_out]
  const float* bias;
  float* out;
};

constexpr int TW = 32, TH = 8;      // output tile
constexpr int CI_T = 8;             // input channels per staged chunk
constexpr int CO_BLK = 32;          // output channels per CTA
constexpr int IN_P = 35;            // smem row pitch (conflict-free: 35 mod 32 = 3)
constexpr int IN_ROWS = TH + 2, IN_COLS = TW + 2;

// value of logical input channel `c` of the concatenated input at (y, x); zero outside the image
__device__ __forceinline__ float fetch_input(const Conv3Params& p, int n, int c, int y, int x) {
  if (y < 0 || y >= p.H || x < 0 || x >= p.W) return 0.f;
  int s = 0;
  while (s < p.n_src - 1 && c >= p.src[s].channels) {  // torch.cat: walk the source list
    c -= p.src[s].channels;
    ++s;
  }
  const SrcDev& sd = p.src[s];
  if (sd.batch_mod > 0) n = (int)(n % sd.batch_mod);
  if (sd.mode == YNET_SRC_DIRECT) {
    return __ldg(sd.ptr + (size_t)n * sd.batch_stride + ((size_t)c * p.H + y) * p.W + x);
  } else if (sd.mode == YNET_SRC_POOL2) {
    const int W2 = 2 * p.W;
    const float* q = sd.ptr + (size_t)n * sd.batch_stride + ((size_t)c * 2 * p.H + 2 * y) * W2 + 2 * x;
    const float2 a = __ldg(reinterpret_cast<const float2*>(q));
    const float2 b = __ldg(reinterpret_cast<const float2*>(q + W2));
    return fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y));
  } else {  // YNET_SRC_UP2: bilinear x2, align_corners=False (ATen upsample_bilinear2d)
    const int h = p.H >> 1, w = p.W >> 1;
    const float fy = fmaxf(0.f, ((float)y + 0.5f) * 0.5f - 0.5f);
    const float fx = fmaxf(0.f, ((float)x + 0.5f) * 0.5f - 0.5f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* q = sd.ptr + (size_t)n * sd.batch_stride + (size_t)c * h * w;
    const float v00 = __ldg(q + (size_t)y0 * w + x0), v01 = __ldg(q + (size_t)y0 * w + x1);
    const float v10 = __ldg(q + (size_t)y1 * w + x0), v11 = __ldg(q + (size_t)y1 * w + x1);
    return (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

__global__ void __launch_bounds__(256, 2) conv3x3_f32_kernel(const Conv3Params p) {
  __shared__ float s_in[CI_T][IN_ROWS][IN_P];
  __shared__ __align__(16) float s_w[CI_T][9][CO_BLK];

  const int co_blocks = ceil_div(p.C_out, CO_BLK);
  const int n = blockIdx.z / co_blocks;
  const int co0 = (blockIdx.z - n * co_blocks) * CO_BLK;
  const int ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;

  const int tid = threadIdx.x;
  const int cog = tid >> 6;       // 4 groups of 8 output channels (warp-uniform)
  const int pt = tid & 63;
  const int py = pt >> 3;         // 0..7
  const int px0 = (pt & 7) * 4;   // 0,4,..,28

  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  for (int ci0 = 0; ci0 < p.C_in; ci0 += CI_T) {
    const int cn = min(CI_T, p.C_in - ci0);
    // stage the input tile (with halo) of `cn` channels
    for (int e = tid; e < cn * IN_ROWS * IN_COLS; e += 256) {
      const int c = e / (IN_ROWS * IN_COLS);
      const int r = e - c * (IN_ROWS * IN_COLS);
      const int yy = r / IN_COLS, xx = r - yy * IN_COLS;
      s_in[c][yy][xx] = fetch_input(p, n, ci0 + c, ty0 + yy - 1, tx0 + xx - 1);
    }
    // stage the weights: packed [C_in][9][C_out] -> [ci][tap][co]
    for (int e = tid; e < cn * 9 * CO_BLK; e += 256) {
      const int co = e & (CO_BLK - 1);
      const int ct = e >> 5;  // ci * 9 + tap
      const int gco = co0 + co;
      (&s_w[0][0][0])[e] = (gco < p.C_out) ? __ldg(p.weight + ((size_t)ci0 * 9 + ct) * p.C_out + gco) : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < cn; ++c) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        float v[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) v[k] = s_in[c][py + kh][px0 + k];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 w0 = *reinterpret_cast<const float4*>(&s_w[c][kh * 3 + kw][cog * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&s_w[c][kh * 3 + kw][cog * 8 + 4]);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(v[a + kw], wv[b], acc[a][b]);
        }
      }
    }
    __syncthreads();
  }

  const int y = ty0 + py;
  if (y >= p.H) return;
  const int x = tx0 + px0;
  const bool vec = ((p.W & 3) == 0) && (x + 3 < p.W);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int co = co0 + cog * 8 + b;
    if (co >= p.C_out) break;
    const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
    float o[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      o[a] = acc[a][b] + bv;
      if (p.relu) o[a] = fmaxf(o[a], 0.f);
    }
    float* dst = p.out + (((size_t)n * p.C_out + co) * p.H + y) * p.W + x;
    if (vec) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a)
        if (x + a < p.W) dst[a] = o[a];
    }
  }
}

// ---- 1x1 predictor -----------------------------------------------------------------------------------
constexpr int P_CO = 16;

__global__ void __launch_bounds__(256)
conv1x1_f32_kernel(const float* __restrict__ x, int C_in, long long S, const float* __restrict__ weight,
                   const float* __restrict__ bias, int C_out, float* __restrict__ out) {
  extern __shared__ float s_w1[];  // [C_in][P_CO]
  const int n = blockIdx.z;
  const int co0 = blockIdx.y * P_CO;
  for (int e = threadIdx.x; e < C_in * P_CO; e += blockDim.x) {
    const int ci = e / P_CO, co = e - ci * P_CO;
    s_w1[e] = (co0 + co < C_out) ? weight[(size_t)(co0 + co) * C_in + ci] : 0.f;
  }
  __syncthreads();
  const float* xn = x + (size_t)n * C_in * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) {
    float acc[P_CO];
#pragma unroll
    for (int b = 0; b < P_CO; ++b) acc[b] = 0.f;
    for (int ci = 0; ci < C_in; ++ci) {
      const float v = __ldg(xn + (size_t)ci * S + i);
#pragma unroll
      for (int b = 0; b < P_CO; ++b) acc[b] = fmaf(v, s_w1[ci * P_CO + b], acc[b]);
    }
#pragma unroll
    for (int b = 0; b < P_CO; ++b)
      if (co0 + b < C_out) out[((size_t)n * C_out + co0 + b) * S + i] = acc[b] + (bias ? bias[co0 + b] : 0.f);
  }
}

// ---- predictor + SoftArgmax2D fused: logits never leave the SM ----------------------------------------
constexpr int PS_PIX = 256;   // pixels per tile (= threads)
constexpr int PS_MAXC = 32;   // max C_in and C_out

struct SoftP {
  float m, s, sx, sy;
};

__device__ __forceinline__ SoftP softp_combine(const SoftP& a, const SoftP& b) {
  SoftP r;
  r.m = fmaxf(a.m, b.m);
  const float fa = (a.m == -FLT_MAX) ? 0.f : __expf(a.m - r.m);
  const float fb = (b.m == -FLT_MAX) ? 0.f : __expf(b.m - r.m);
  r.s = a.s * fa + b.s * fb;
  r.sx = a.sx * fa + b.sx * fb;
  r.sy = a.sy * fa + b.sy * fb;
  return r;
}

// grid = (splits, N); each CTA walks its pixel slice in tiles of 256 pixels
__global__ void __launch_bounds__(PS_PIX)
predictor_softargmax_kernel(const float* __restrict__ x, int C_in, int H, int W, const float* __restrict__ weight,
                            const float* __restrict__ bias, int C_out, int splits, SoftP* __restrict__ part) {
  __shared__ __align__(16) float s_w[PS_MAXC][PS_MAXC];   // [ci][co]
  __shared__ float s_b[PS_MAXC];
  __shared__ float s_logit[PS_MAXC][PS_PIX];   // [co][pixel]
  const int n = blockIdx.y, split = blockIdx.x;
  const int S = H * W;
  for (int e = threadIdx.x; e < PS_MAXC * PS_MAXC; e += PS_PIX) {
    const int co = e / PS_MAXC, ci = e - co * PS_MAXC;
    s_w[ci][co] = (co < C_out && ci < C_in) ? weight[co * C_in + ci] : 0.f;
  }
  if (threadIdx.x < PS_MAXC) s_b[threadIdx.x] = (bias && threadIdx.x < C_out) ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int per = ceil_div(ceil_div(S, splits), PS_PIX) * PS_PIX;
  const int i0 = split * per, i1 = min(S, i0 + per);
  const float* xn = x + (size_t)n * C_in * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = PS_PIX / 32;
  constexpr int CPW = PS_MAXC / NW;  // channels per warp (4)
  SoftP st[CPW];
#pragma unroll
  for (int k = 0; k < CPW; ++k) st[k] = SoftP{-FLT_MAX, 0.f, 0.f, 0.f};

  for (int t0 = i0; t0 < i1; t0 += PS_PIX) {
    const int i = t0 + threadIdx.x;
    {
      float acc[PS_MAXC];
#pragma unroll
      for (int b = 0; b < PS_MAXC; ++b) acc[b] = 0.f;
      if (i < i1) {
        for (int ci = 0; ci < C_in; ++ci) {
          const float v = __ldg(xn + (size_t)ci * S + i);
#pragma unroll
          for (int b4 = 0; b4 < PS_MAXC / 4; ++b4) {
            const float4 w = *reinterpret_cast<const float4*>(&s_w[ci][4 * b4]);
            acc[4 * b4 + 0] = fmaf(v, w.x, acc[4 * b4 + 0]);
            acc[4 * b4 + 1] = fmaf(v, w.y, acc[4 * b4 + 1]);
            acc[4 * b4 + 2] = fmaf(v, w.z, acc[4 * b4 + 2]);
            acc[4 * b4 + 3] = fmaf(v, w.w, acc[4 * b4 + 3]);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < PS_MAXC; ++b) s_logit[b][threadIdx.x] = (i < i1) ? acc[b] + s_b[b] : -FLT_MAX;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CPW; ++k) {
      const int co = warp + k * NW;
      if (co < C_out) {
        float v[PS_PIX / 32];
        float mx = -FLT_MAX;
#pragma unroll
        for (int q = 0; q < PS_PIX / 32; ++q) {
          v[q] = s_logit[co][q * 32 + lane];
          mx = fmaxf(mx, v[q]);
        }
        if (mx > st[k].m) {
          const float f = (st[k].m == -FLT_MAX) ? 0.f : __expf(st[k].m - mx);
          st[k].s *= f;
          st[k].sx *= f;
          st[k].sy *= f;
          st[k].m = mx;
        }
#pragma unroll
        for (int q = 0; q < PS_PIX / 32; ++q) {
          const int pi = t0 + q * 32 + lane;
          if (pi < i1) {
            const int yy = pi / W;
            const float e = __expf(v[q] - st[k].m);
            st[k].s += e;
            st[k].sx = fmaf(e, (float)(pi - yy * W), st[k].sx);
            st[k].sy = fmaf(e, (float)yy, st[k].sy);
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < CPW; ++k) {
    const int co = warp + k * NW;
    SoftP p = st[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      SoftP q;
      q.m = __shfl_xor_sync(0xffffffffu, p.m, o);
      q.s = __shfl_xor_sync(0xffffffffu, p.s, o);
      q.sx = __shfl_xor_sync(0xffffffffu, p.sx, o);
      q.sy = __shfl_xor_sync(0xffffffffu, p.sy, o);
      p = softp_combine(p, q);
    }
    if (lane == 0 && co < C_out) part[((size_t)n * C_out + co) * splits + split] = p;
  }
}

__global__ void predictor_softargmax_finalize_kernel(const SoftP* __restrict__ part, int rows, int splits,
                                                     float* __restrict__ out) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  SoftP p{-FLT_MAX, 0.f, 0.f, 0.f};
  for (int k = 0; k < splits; ++k) p = softp_combine(p, part[(size_t)row * splits + k]);
  const float inv = 1.0f / (p.s + 1e-6f);
  out[2 * row + 0] = p.sx * inv;
  out[2 * row + 1] = p.sy * inv;
}

// ---- standalone pool / upsample (API completeness, training path) ----------------------------------------
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const float* __restrict__ x, long long planes, int H, int W,
                                                         float* __restrict__ out) {
  const int h = H >> 1, w = W >> 1;
  const long long total = planes * h * w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / (h * w);
    const int r = (int)(t - pl * h * w);
    const int y = r / w, xx = r - y * w;
    const float* q = x + (pl * H + 2 * y) * W + 2 * xx;
    out[t] = fmaxf(fmaxf(q[0], q[1]), fmaxf(q[W], q[W + 1]));
  }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const float* __restrict__ x, long long planes, int H, int W,
                                                         float* __restrict__ out) {
  const int OH = 2 * H, OW = 2 * W;
  const long long total = planes * OH * OW;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / ((long long)OH * OW);
    const int r = (int)(t - pl * OH * OW);
    const int y = r / OW, xx = r - y * OW;
    const float fy = fmaxf(0.f, ((float)y + 0.5f) * 0.5f - 0.5f);
    const float fx = fmaxf(0.f, ((float)xx + 0.5f) * 0.5f - 0.5f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* q = x + pl * H * W;
    out[t] = (1.f - ly) * ((1.f - lx) * q[(size_t)y0 * W + x0] + lx * q[(size_t)y0 * W + x1]) +
             ly * ((1.f - lx) * q[(size_t)y1 * W + x0] + lx * q[(size_t)y1 * W + x1]);
  }
}

// ---- LoRA fold: W + (B @ A).view(W.shape) * (1/r); output OIHW (layout 0) or packed [C_in][k*k][C_out] (1) ---
__global__ void __launch_bounds__(256)
lora_fold_kernel(const float* __restrict__ weight, const float* __restrict__ A, const float* __restrict__ Bm, int C_out,
                 int C_in, int ks, int rank, int layout, float* __restrict__ out) {
  const int total = C_out * C_in * ks * ks;
  const int rk = rank * ks;       // inner dimension of B @ A
  const int cols = C_in * ks;     // columns of (B @ A)
  const float scale = rank > 0 ? 1.0f / (float)rank : 0.f;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
    float w = weight[f];
    if (A != nullptr && Bm != nullptr && rank > 0) {
      const int row = f / cols, col = f - row * cols;
      float d = 0.f;
      for (int j = 0; j < rk; ++j) d = fmaf(Bm[(size_t)row * rk + j], A[(size_t)j * cols + col], d);
      w = w + d * scale;
    }
    if (layout == 0) {
      out[f] = w;
    } else {
      const int kk = ks * ks;
      const int o = f / (C_in * kk);
      const int rem = f - o * C_in * kk;
      const int i = rem / kk, tap = rem - i * kk;
      out[((size_t)i * kk + tap) * C_out + o] = w;
    }
  }
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int ynet_conv3x3_f32(const ynet_conv_src* srcs, int32_t n_src, int32_t N, int32_t H, int32_t W,
                     const float* weight_packed, const float* bias, int32_t C_out, int32_t relu, float* out,
                     void* stream) {
  YNET_CHECK_ARG(srcs && weight_packed && out, "null pointer");
  YNET_CHECK_ARG(n_src >= 1 && n_src <= YNET_MAX_SOURCES, "n_src must be in [1, 4]");
  YNET_CHECK_ARG(N >= 0 && H > 0 && W > 0 && C_out > 0, "bad shape");
  if (N == 0) return YNET_OK;
  Conv3Params p;
  memset(&p, 0, sizeof(p));
  int cin = 0;
  for (int i = 0; i < n_src; ++i) {
    YNET_CHECK_ARG(srcs[i].ptr != nullptr && srcs[i].channels > 0, "bad source");
    YNET_CHECK_ARG(srcs[i].mode >= 0 && srcs[i].mode <= 2, "bad source mode");
    if (srcs[i].mode == YNET_SRC_UP2) YNET_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "UP2 source needs even H, W");
    if (srcs[i].mode == YNET_SRC_POOL2) YNET_CHECK_ALIGN(srcs[i].ptr, 8);
    p.src[i].ptr = reinterpret_cast<const float*>(srcs[i].ptr);
    p.src[i].channels = srcs[i].channels;
    p.src[i].mode = srcs[i].mode;
    p.src[i].batch_stride = srcs[i].batch_stride;
    p.src[i].batch_mod = srcs[i].batch_mod;
    cin += srcs[i].channels;
  }
  p.n_src = n_src;
  p.C_in = cin;
  p.N = N;
  p.H = H;
  p.W = W;
  p.C_out = C_out;
  p.relu = relu;
  p.weight = weight_packed;
  p.bias = bias;
  p.out = out;
  const int co_blocks = ceil_div(C_out, CO_BLK);
  const long long gz = (long long)N * co_blocks;
  YNET_CHECK_ARG(gz <= 65535, "N * ceil(C_out/32) must be <= 65535 per call");
  dim3 grid(ceil_div(W, TW), ceil_div(H, TH), (unsigned)gz);
  conv3x3_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_conv1x1_f32(const float* x, int32_t N, int32_t C_in, int64_t S, const float* weight, const float* bias,
                     int32_t C_out, float* out, void* stream) {
  YNET_CHECK_ARG(x && weight && out, "null pointer");
  YNET_CHECK_ARG(N >= 0 && N <= 65535 && C_in > 0 && C_in <= 1024 && S > 0 && C_out > 0, "bad shape");
  if (N == 0) return YNET_OK;
  dim3 grid((unsigned)tmin<long long>(ceil_div<long long>(S, 256), 2048), ceil_div(C_out, P_CO), N);
  conv1x1_f32_kernel<<<grid, 256, C_in * P_CO * sizeof(float), as_stream(stream)>>>(x, C_in, S, weight, bias, C_out,
                                                                                    out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

static int ps_splits(int N, int S) {
  int splits = ceil_div(3 * sm_count(), max(N, 1));
  splits = max(1, min(splits, max(1, S / (4 * PS_PIX))));
  return min(splits, 128);
}

int64_t ynet_predictor_softargmax_workspace_bytes(int32_t N, int32_t C_out, int32_t H, int32_t W) {
  return (int64_t)N * C_out * 128 * (int64_t)sizeof(SoftP);
}

int ynet_predictor_softargmax_f32(const float* x, int32_t N, int32_t C_in, int32_t H, int32_t W, const float* weight,
                                  const float* bias, int32_t C_out, float* out, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
  YNET_CHECK_ARG(x && weight && out, "null pointer");
  YNET_CHECK_ARG(N >= 0 && N <= 65535 && H > 0 && W > 0, "bad shape");
  YNET_CHECK_ARG(C_in > 0 && C_in <= PS_MAXC && C_out > 0 && C_out <= PS_MAXC, "C_in, C_out must be <= 32");
  if (N == 0) return YNET_OK;
  const int S = H * W;
  const int splits = ps_splits(N, S);
  if (workspace == nullptr || workspace_bytes < (int64_t)N * C_out * splits * (int64_t)sizeof(SoftP)) {
    set_error("ynet_predictor_softargmax_f32: workspace too small");
    return YNET_E_WORKSPACE;
  }
  SoftP* part = reinterpret_cast<SoftP*>(workspace);
  predictor_softargmax_kernel<<<dim3(splits, N), PS_PIX, 0, as_stream(stream)>>>(x, C_in, H, W, weight, bias, C_out,
                                                                                 splits, part);
  YNET_LAUNCH_CHECK();
  predictor_softargmax_finalize_kernel<<<ceil_div(N * C_out, 128), 128, 0, as_stream(stream)>>>(part, N * C_out, splits,
                                                                                               out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_maxpool2x2_f32(const float* x, int64_t planes, int32_t H, int32_t W, float* out, void* stream) {
  YNET_CHECK_ARG(x && out && planes >= 0 && H >= 2 && W >= 2, "bad argument");
  if (planes == 0) return YNET_OK;
  const long long total = planes * (H / 2) * (W / 2);
  maxpool2x2_kernel<<<(unsigned)tmin<long long>(ceil_div<long long>(total, 256), 16LL * sm_count()), 256, 0,
                      as_stream(stream)>>>(x, planes, H, W, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_upsample_bilinear2x_f32(const float* x, int64_t planes, int32_t H, int32_t W, float* out, void* stream) {
  YNET_CHECK_ARG(x && out && planes >= 0 && H >= 1 && W >= 1, "bad argument");
  if (planes == 0) return YNET_OK;
  const long long total = planes * 4LL * H * W;
  upsample2x_kernel<<<(unsigned)tmin<long long>(ceil_div<long long>(total, 256), 16LL * sm_count()), 256, 0,
                      as_stream(stream)>>>(x, planes, H, W, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_lora_fold(const float* weight, const float* lora_A, const float* lora_B, int32_t C_out, int32_t C_in,
                   int32_t ksize, int32_t rank, int32_t out_layout, float* out, void* stream) {
  YNET_CHECK_ARG(weight && out, "null pointer");
  YNET_CHECK_ARG(C_out > 0 && C_in > 0 && ksize > 0 && rank >= 0, "bad shape");
  YNET_CHECK_ARG(out_layout == 0 || out_layout == 1, "out_layout: 0 = OIHW, 1 = packed [C_in][k*k][C_out]");
  const int total = C_out * C_in * ksize * ksize;
  lora_fold_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(weight, lora_A, lora_B, C_out, C_in, ksize,
                                                                        rank, out_layout, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

}  // extern "C"
