// a18: pieces of the MoSA fine-tuning step (utils/train_epoch.py:86-115, models/trainer.py:197-206):
// fused BCE-with-logits forward+backward, conv dgrad / wgrad (fp32), pool / bilinear backward,
// LoRA gradient projection and Adam.  All fp32; reductions are fixed-order (deterministic).
#include <float.h>

#include "common.cuh"

namespace ynet {

// ---- BCEWithLogitsLoss(mean) forward + backward -------------------------------------------------------
constexpr int kBceBlocks = 1024;

__global__ void __launch_bounds__(256)
bce_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n, float gscale,
           double* __restrict__ partial, float* __restrict__ grad) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = x[i], tv = t[i];
    // max(x,0) - x*t + log(1 + exp(-|x|))   (ATen binary_cross_entropy_with_logits)
    const float l = fmaxf(xv, 0.f) - xv * tv + log1pf(expf(-fabsf(xv)));
    acc += (double)l;
    if (grad != nullptr) {
      const float s = 1.0f / (1.0f + expf(-xv));
      grad[i] = (s - tv) * gscale;
    }
  }
  acc = warp_sum(acc);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sh[w];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) bce_final_kernel(const double* __restrict__ partial, int nb, long long n,
                                                        float* __restrict__ loss) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
  acc = warp_sum(acc);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sh[w];
    *loss = (float)(s / (double)n);
  }
}

// ---- dgrad weight transform: OIHW -> packed [C_out][9][C_in] with flipped taps ---------------------------
__global__ void __launch_bounds__(256)
dgrad_weight_kernel(const float* __restrict__ w, int C_out, int C_in, float* __restrict__ packed) {
  const int total = C_out * C_in * 9;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
    const int co = f / (C_in * 9);
    const int rem = f - co * C_in * 9;
    const int ci = rem / 9, tap = rem - ci * 9;
    const int kh = tap / 3, kw = tap - kh * 3;
    const int tap_f = (2 - kh) * 3 + (2 - kw);
    packed[((size_t)co * 9 + tap_f) * C_in + ci] = w[f];
  }
}

__global__ void __launch_bounds__(256)
relu_mask_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (y[i] > 0.f) ? dy[i] : 0.f;
}

// ---- wgrad: dW[co][ci][tap] = sum_{n,y,x} dy[n,co,y,x] * x[n,ci,y+kh-1,x+kw-1] ------------------------------
constexpr int WG_TW = 16, WG_TH = 8;
constexpr int WG_C = 32;     // channel chunk (both C_in and C_out)
constexpr int WG_P = 33;     // smem pitch over channels

// grid.x = persistent CTAs, grid.y = co chunk, grid.z = ci chunk.  Each CTA walks spatial tiles and keeps
// 4 co x 1 ci x 9 taps in registers per thread; one partial per CTA, reduced in fixed order afterwards.
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ relu_out, int N, int H,
             int W, int C_in, int C_out, float* __restrict__ partial /* (grid.x, C_out, C_in, 9) */,
             float* __restrict__ partial_db /* (grid.x, C_out) or null */) {
  __shared__ float s_x[(WG_TH + 2) * (WG_TW + 2) * WG_P];
  __shared__ float s_dy[WG_TH * WG_TW * WG_P];
  const int co0 = blockIdx.y * WG_C, ci0 = blockIdx.z * WG_C;
  const int ci = threadIdx.x & 31;   // lane
  const int cog = threadIdx.x >> 5;  // warp: 4 output channels each
  float acc[4][9];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[a][k] = 0.f;
  float db[4] = {0.f, 0.f, 0.f, 0.f};

  const int tiles_x = ceil_div(W, WG_TW), tiles_y = ceil_div(H, WG_TH);
  const long long n_tiles = (long long)N * tiles_x * tiles_y;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n = (int)(tile / (tiles_x * tiles_y));
    const int r = (int)(tile - (long long)n * tiles_x * tiles_y);
    const int ty0 = (r / tiles_x) * WG_TH, tx0 = (r % tiles_x) * WG_TW;
    __syncthreads();
    for (int e = threadIdx.x; e < WG_C * (WG_TH + 2) * (WG_TW + 2); e += 256) {
      const int c = e / ((WG_TH + 2) * (WG_TW + 2));
      const int q = e - c * ((WG_TH + 2) * (WG_TW + 2));
      const int yy = q / (WG_TW + 2), xx = q - yy * (WG_TW + 2);
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1, gc = ci0 + c;
      float v = 0.f;
      if (gc < C_in && gy >= 0 && gy < H && gx >= 0 && gx < W) v = x[(((size_t)n * C_in + gc) * H + gy) * W + gx];
      s_x[q * WG_P + c] = v;
    }
    for (int e = threadIdx.x; e < WG_C * WG_TH * WG_TW; e += 256) {
      const int c = e / (WG_TH * WG_TW);
      const int q = e - c * (WG_TH * WG_TW);
      const int yy = q / WG_TW, xx = q - yy * WG_TW;
      const int gy = ty0 + yy, gx = tx0 + xx, gc = co0 + c;
      float v = 0.f;
      if (gc < C_out && gy < H && gx < W) {
        const size_t o = (((size_t)n * C_out + gc) * H + gy) * W + gx;
        v = dy[o];
        if (relu_out != nullptr && !(relu_out[o] > 0.f)) v = 0.f;
      }
      s_dy[q * WG_P + c] = v;
    }
    __syncthreads();
    for (int yy = 0; yy < WG_TH; ++yy) {
      for (int xx = 0; xx < WG_TW; ++xx) {
        float g[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) g[a] = s_dy[(yy * WG_TW + xx) * WG_P + cog * 4 + a];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float xv = s_x[((yy + kh) * (WG_TW + 2) + xx + kw) * WG_P + ci];
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[a][kh * 3 + kw] = fmaf(g[a], xv, acc[a][kh * 3 + kw]);
          }
        if (ci == 0) {
#pragma unroll
          for (int a = 0; a < 4; ++a) db[a] += g[a];
        }
      }
    }
  }
  const int gci = ci0 + ci;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gco = co0 + cog * 4 + a;
    if (gco < C_out && gci < C_in) {
      float* o = partial + (((size_t)blockIdx.x * C_out + gco) * C_in + gci) * 9;
#pragma unroll
      for (int k = 0; k < 9; ++k) o[k] = acc[a][k];
    }
    if (partial_db != nullptr && ci == 0 && blockIdx.z == 0 && gco < C_out)
      partial_db[(size_t)blockIdx.x * C_out + gco] = db[a];
  }
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int n_part, long long len, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < n_part; ++p) s += partial[(size_t)p * len + i];
    out[i] = s;
  }
}

// ---- MaxPool2d(2,2) backward: first maximum in (h, w) scan order takes the gradient ----------------------
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long planes, int H, int W,
                   float* __restrict__ dx) {
  const int h = H >> 1, w = W >> 1;
  const long long total = planes * h * w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / (h * w);
    const int r = (int)(t - pl * h * w);
    const int y = r / w, xx = r - y * w;
    const size_t base = ((size_t)pl * H + 2 * y) * W + 2 * xx;
    const float v[4] = {x[base], x[base + 1], x[base + W], x[base + W + 1]};
    int am = 0;
    float m = v[0];
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (v[k] > m) {
        m = v[k];
        am = k;
      }
    const float g = dy[t];
    dx[base] = (am == 0) ? g : 0.f;
    dx[base + 1] = (am == 1) ? g : 0.f;
    dx[base + W] = (am == 2) ? g : 0.f;
    dx[base + W + 1] = (am == 3) ? g : 0.f;
  }
}

// ---- bilinear x2 (align_corners=False) backward: gather form --------------------------------------------
// 1-D: dx[i] = w(2i) dy[2i] + w(2i+1) dy[2i+1] + 0.25 dy[2i+2] (i+1 < n) + 0.25 dy[2i-1] (i > 0),
// with w(2i) = 1 if i == 0 else 0.75 and w(2i+1) = 1 if i == n-1 else 0.75.
__device__ __forceinline__ int up_taps(int i, int n, int* idx, float* wt) {
  int k = 0;
  idx[k] = 2 * i;
  wt[k++] = (i == 0) ? 1.0f : 0.75f;
  idx[k] = 2 * i + 1;
  wt[k++] = (i == n - 1) ? 1.0f : 0.75f;
  if (i + 1 < n) {
    idx[k] = 2 * i + 2;
    wt[k++] = 0.25f;
  }
  if (i > 0) {
    idx[k] = 2 * i - 1;
    wt[k++] = 0.25f;
  }
  return k;
}

__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const float* __restrict__ dy, long long planes, int H, int W, float* __restrict__ dx) {
  const int OW = 2 * W;
  const long long total = planes * H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / ((long long)H * W);
    const int r = (int)(t - pl * H * W);
    const int i = r / W, j = r - i * W;
    int iy[4], ix[4];
    float wy[4], wx[4];
    const int ny = up_taps(i, H, iy, wy), nx = up_taps(j, W, ix, wx);
    const float* g = dy + (size_t)pl * 4 * H * W;
    float s = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) s = fmaf(wy[a] * wx[b], g[(size_t)iy[a] * OW + ix[b]], s);
    dx[t] = s;
  }
}

// ---- LoRA gradient projection ------------------------------------------------------------------------------
// dM = dW.view(C_out*k, C_in*k); dA[j,c] = s * sum_r B[r,j] dM[r,c]; dB[r,j] = s * sum_c dM[r,c] A[j,c]
__global__ void __launch_bounds__(256)
lora_grad_kernel(const float* __restrict__ dW, const float* __restrict__ A, const float* __restrict__ Bm, int rows,
                 int cols, int rk, float scale, float* __restrict__ dA, float* __restrict__ dB) {
  const int nA = rk * cols, nB = rows * rk;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nA + nB; t += gridDim.x * blockDim.x) {
    if (t < nA) {
      const int j = t / cols, c = t - j * cols;
      float s = 0.f;
      for (int r = 0; r < rows; ++r) s = fmaf(Bm[(size_t)r * rk + j], dW[(size_t)r * cols + c], s);
      dA[t] = s * scale;
    } else {
      const int u = t - nA;
      const int r = u / rk, j = u - r * rk;
      float s = 0.f;
      for (int c = 0; c < cols; ++c) s = fmaf(dW[(size_t)r * cols + c], A[(size_t)j * cols + c], s);
      dB[u] = s * scale;
    }
  }
}

// ---- Adam (torch.optim.Adam defaults, no weight decay, no amsgrad) -----------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float gscale, float bc1, float bc2_sqrt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

static inline unsigned grid_for(long long n) {
  return (unsigned)tmax<long long>(1, tmin<long long>(ceil_div<long long>(n, 256), 16LL * sm_count()));
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int ynet_conv3x3_f32(const ynet_conv_src* srcs, int32_t n_src, int32_t N, int32_t H, int32_t W,
                     const float* weight_packed, const float* bias, int32_t C_out, int32_t relu, float* out,
                     void* stream);

int64_t ynet_bce_workspace_bytes(int64_t n) { return (int64_t)kBceBlocks * sizeof(double); }

int ynet_bce_logits_fwd_bwd(const float* logits, const float* target, int64_t n, float grad_scale, float* loss_out,
                            float* grad, void* workspace, int64_t workspace_bytes, void* stream) {
  YNET_CHECK_ARG(logits && target && loss_out && n > 0, "bad argument");
  if (workspace == nullptr || workspace_bytes < ynet_bce_workspace_bytes(n)) {
    set_error("ynet_bce_logits_fwd_bwd: workspace too small");
    return YNET_E_WORKSPACE;
  }
  double* partial = reinterpret_cast<double*>(workspace);
  const int nb = (int)tmin<long long>(kBceBlocks, ceil_div<long long>(n, 256));
  bce_kernel<<<nb, 256, 0, as_stream(stream)>>>(logits, target, n, grad_scale / (float)n, partial, grad);
  YNET_LAUNCH_CHECK();
  bce_final_kernel<<<1, 256, 0, as_stream(stream)>>>(partial, nb, n, loss_out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

// The scratch of the fp32 data gradient comes from the device's stream-ordered pool.  By default that pool hands unused
// memory back to the OS at every synchronisation point (release threshold 0), so the ~200 MB masked-gradient buffer was
// unmapped at each epoch-end read of the loss and mapped again by the next step: single steps of 100-840 ms in a run of
// 48 ms steps (profiles/finetune_fp32_excursions_r02.log).  Told once per device to keep what it has.
static void keep_pool_memory() {
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[dev] = true;
}

int ynet_conv3x3_dgrad_f32(const float* dy, const float* relu_out, int32_t N, int32_t H, int32_t W, const float* weight,
                           int32_t C_out, int32_t C_in, float* dx, void* stream) {
  YNET_CHECK_ARG(dy && weight && dx, "null pointer");
  YNET_CHECK_ARG(N > 0 && H > 0 && W > 0 && C_out > 0 && C_in > 0, "bad shape");
  cudaStream_t st = as_stream(stream);
  keep_pool_memory();
  float* packed = nullptr;
  float* masked = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&packed), (size_t)C_out * C_in * 9 * sizeof(float), st);
  if (e != cudaSuccess) return cuda_fail(e, "ynet_conv3x3_dgrad_f32(cudaMallocAsync)");
  dgrad_weight_kernel<<<ceil_div(C_out * C_in * 9, 256), 256, 0, st>>>(weight, C_out, C_in, packed);
  const float* src = dy;
  if (relu_out != nullptr) {
    const long long n = (long long)N * C_out * H * W;
    e = cudaMallocAsync(reinterpret_cast<void**>(&masked), (size_t)n * sizeof(float), st);
    if (e != cudaSuccess) {
      cudaFreeAsync(packed, st);
      return cuda_fail(e, "ynet_conv3x3_dgrad_f32(cudaMallocAsync)");
    }
    relu_mask_kernel<<<grid_for(n), 256, 0, st>>>(dy, relu_out, n, masked);
    src = masked;
  }
  ynet_conv_src s;
  s.ptr = src;
  s.channels = C_out;
  s.mode = YNET_SRC_DIRECT;
  s.batch_stride = (int64_t)C_out * H * W;
  s.batch_mod = 0;
  int rc = YNET_OK;
  // N * ceil(C_in/32) must fit gridDim.z: split the batch if needed
  const int per = tmax(1, 65535 / ceil_div(C_in, 32));
  for (int n0 = 0; n0 < N && rc == YNET_OK; n0 += per) {
    const int nn = tmin(per, N - n0);
    ynet_conv_src sc = s;
    sc.ptr = src + (size_t)n0 * C_out * H * W;
    rc = ynet_conv3x3_f32(&sc, 1, nn, H, W, packed, nullptr, C_in, 0, dx + (size_t)n0 * C_in * H * W, stream);
  }
  cudaFreeAsync(packed, st);
  if (masked) cudaFreeAsync(masked, st);
  return rc;
}

static int wgrad_ctas() { return 2 * sm_count(); }

int64_t ynet_conv3x3_wgrad_workspace_bytes(int32_t N, int32_t H, int32_t W, int32_t C_out, int32_t C_in) {
  return (int64_t)wgrad_ctas() * ((int64_t)C_out * C_in * 9 + C_out) * (int64_t)sizeof(float);
}

int ynet_conv3x3_wgrad_f32(const float* x, const float* dy, const float* relu_out, int32_t N, int32_t H, int32_t W,
                           int32_t C_in, int32_t C_out, float* dW, float* db, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  YNET_CHECK_ARG(x && dy && dW, "null pointer");
  YNET_CHECK_ARG(N > 0 && H > 0 && W > 0 && C_out > 0 && C_in > 0, "bad shape");
  if (workspace == nullptr || workspace_bytes < ynet_conv3x3_wgrad_workspace_bytes(N, H, W, C_out, C_in)) {
    set_error("ynet_conv3x3_wgrad_f32: workspace too small");
    return YNET_E_WORKSPACE;
  }
  const int P = wgrad_ctas();
  float* partial = reinterpret_cast<float*>(workspace);
  float* partial_db = partial + (size_t)P * C_out * C_in * 9;
  dim3 grid(P, ceil_div(C_out, WG_C), ceil_div(C_in, WG_C));
  wgrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, dy, relu_out, N, H, W, C_in, C_out, partial,
                                                    db ? partial_db : nullptr);
  YNET_LAUNCH_CHECK();
  const long long len = (long long)C_out * C_in * 9;
  reduce_partials_kernel<<<grid_for(len), 256, 0, as_stream(stream)>>>(partial, P, len, dW);
  YNET_LAUNCH_CHECK();
  if (db) {
    reduce_partials_kernel<<<1, 256, 0, as_stream(stream)>>>(partial_db, P, C_out, db);
    YNET_LAUNCH_CHECK();
  }
  return YNET_OK;
}

int ynet_maxpool2x2_bwd_f32(const float* x, const float* dy, int64_t planes, int32_t H, int32_t W, float* dx,
                            void* stream) {
  YNET_CHECK_ARG(x && dy && dx && planes > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "bad argument");
  maxpool_bwd_kernel<<<grid_for(planes * (H / 2) * (W / 2)), 256, 0, as_stream(stream)>>>(x, dy, planes, H, W, dx);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_upsample_bilinear2x_bwd_f32(const float* dy, int64_t planes, int32_t H, int32_t W, float* dx, void* stream) {
  YNET_CHECK_ARG(dy && dx && planes > 0 && H >= 1 && W >= 1, "bad argument");
  upsample_bwd_kernel<<<grid_for(planes * H * W), 256, 0, as_stream(stream)>>>(dy, planes, H, W, dx);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_lora_grad(const float* dW, const float* lora_A, const float* lora_B, int32_t C_out, int32_t C_in, int32_t ksize,
                   int32_t rank, float* dA, float* dB, void* stream) {
  YNET_CHECK_ARG(dW && lora_A && lora_B && dA && dB, "null pointer");
  YNET_CHECK_ARG(C_out > 0 && C_in > 0 && ksize > 0 && rank > 0, "bad shape");
  const int rows = C_out * ksize, cols = C_in * ksize, rk = rank * ksize;
  lora_grad_kernel<<<grid_for((long long)rk * cols + (long long)rows * rk), 256, 0, as_stream(stream)>>>(
      dW, lora_A, lora_B, rows, cols, rk, 1.0f / (float)rank, dA, dB);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t step, float lr,
                   float beta1, float beta2, float eps, float grad_scale, void* stream) {
  YNET_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "bad argument");
  if (n == 0) return YNET_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                          grad_scale, (float)bc1, (float)sqrt(bc2));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

}  // extern "C"
