// a16/a17: CWS waypoint conditioning (utils/evaluate.py:9-34, 172-224) and ADE/FDE (276-291).
//
// The reference builds, for every (goal, agent), an H x W oriented Gaussian with ~15 tiny torch ops,
// multiplies it with the sigmoid map, renormalises and takes the expectation: 20*B launch storms.
// Here one CTA slice reads the agent's sigmoid map ONCE for all goals (one warp per goal) and
// accumulates {sum w, sum w*x, sum w*y}; the Gaussian's own normaliser cancels in the expectation.
#include "common.cuh"

namespace ynet {

struct CwsPrior {
  float mx, my;          // Gaussian mean (x, y)
  float t00, t01, t11;   // inverse covariance (symmetric)
};

// evaluate.py:9-30 restated; Sigma^-1 = R diag(1/a^2, 1/b^2) R^T analytically (R is orthogonal).
__device__ __forceinline__ CwsPrior make_prior(float wx, float wy, float lx, float ly, float length_ratio,
                                              float sigma_factor, float ratio, int rot) {
  const float dx = lx - wx, dy = ly - wy;  // dist = last_observed - waypoint   (evaluate.py:182)
  CwsPrior p;
  p.mx = wx + dx * length_ratio;           // gauss_mean (evaluate.py:190)
  p.my = wy + dy * length_ratio;
  const float rad = atan2f(dx, dy);        // torch.atan2(dist[0], dist[1])
  const float c = cosf(rad), s = sinf(rad);
  float r00 = c, r01 = s, r10 = -s, r11 = c;
  if (rot) {  // [[0,-1],[1,0]] @ R
    const float a00 = -r10, a01 = -r11, a10 = r00, a11 = r01;
    r00 = a00; r01 = a01; r10 = a10; r11 = a11;
  }
  const float dn = sqrtf(dx * dx + dy * dy) + 5.0f;
  const float a = dn / sigma_factor / ratio, b = dn / sigma_factor;
  const float ia = 1.0f / (a * a), ib = 1.0f / (b * b);
  p.t00 = r00 * r00 * ia + r01 * r01 * ib;
  p.t01 = r00 * r10 * ia + r01 * r11 * ib;
  p.t11 = r10 * r10 * ia + r11 * r11 * ib;
  return p;
}

// torch.linspace(0, n, n)[i] (fp32, symmetric evaluation) -- evaluate.py:13-14: spacing n/(n-1), not 1
__device__ __forceinline__ float linspace0n(int i, int n) {
  const float step = (float)n / (float)(n - 1);
  return (i < n / 2) ? step * (float)i : (float)n - step * (float)(n - 1 - i);
}

__device__ __forceinline__ float prior_value(const CwsPrior& p, int i, int j, int H, int W) {
  const float u = linspace0n(j, W) - p.mx;  // meshgrid[..., 0] = yy = ay (x offsets)
  const float v = linspace0n(i, H) - p.my;  // meshgrid[..., 1] = xx = ax (y offsets)
  const float q = p.t00 * u * u + 2.0f * p.t01 * u * v + p.t11 * v * v;
  return __expf(-0.5f * q);
}

// grid = (splits, B), block = 32 * G threads (warp g handles goal g).
__global__ void cws_partial_kernel(const float* __restrict__ sig, int H, int W, const float* __restrict__ wp_in, int B,
                                   int G, const float* __restrict__ last_obs, float length_ratio,
                                   const float* __restrict__ sigma_factor, float ratio, int rot, int splits,
                                   float* __restrict__ partial /* (B, splits, G, 3) */) {
  const int b = blockIdx.y, split = blockIdx.x;
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = H * W;
  const float* src = sig + (size_t)b * S;
  const CwsPrior p = make_prior(wp_in[((size_t)g * B + b) * 2 + 0], wp_in[((size_t)g * B + b) * 2 + 1],
                                last_obs[2 * b + 0], last_obs[2 * b + 1], length_ratio, sigma_factor[g], ratio, rot);
  // whole rows per split: the row terms of the quadratic form are hoisted and no pixel index is divided
  const int rows_per = ceil_div(H, splits);
  const int r0 = split * rows_per, r1 = min(H, r0 + rows_per);
  float s = 0.f, sx = 0.f, sy = 0.f;
  for (int i = r0; i < r1; ++i) {
    const float v = linspace0n(i, H) - p.my;
    const float qv = p.t11 * v * v, tv = 2.0f * p.t01 * v;
    const float* row = src + (size_t)i * W;
    float rs = 0.f;
    for (int j = lane; j < W; j += 32) {
      const float u = linspace0n(j, W) - p.mx;
      // same operation order as prior_value: t00*u*u + 2*t01*u*v + t11*v*v
      const float q = p.t00 * u * u + tv * u + qv;
      const float w = row[j] * __expf(-0.5f * q);
      rs += w;
      sx = fmaf(w, (float)j, sx);
    }
    s += rs;
    sy = fmaf(rs, (float)i, sy);
  }
  s = warp_sum(s);
  sx = warp_sum(sx);
  sy = warp_sum(sy);
  if (lane == 0) {
    float* o = partial + (((size_t)b * splits + split) * G + g) * 3;
    o[0] = s;
    o[1] = sx;
    o[2] = sy;
  }
}

__global__ void cws_finalize_kernel(const float* __restrict__ partial, int B, int G, int splits,
                                    float* __restrict__ out /* (G, B, 2) */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * G) return;
  const int b = t / G, g = t - b * G;
  float s = 0.f, sx = 0.f, sy = 0.f;
  for (int k = 0; k < splits; ++k) {
    const float* o = partial + (((size_t)b * splits + k) * G + g) * 3;
    s += o[0];
    sx += o[1];
    sy += o[2];
  }
  out[((size_t)g * B + b) * 2 + 0] = sx / s;
  out[((size_t)g * B + b) * 2 + 1] = sy / s;
}

// normalised waypoint map of ONE goal (n_traj > 1 re-sampling path): out = sig*k / sum(sig*k)
__global__ void __launch_bounds__(512)
cws_map_kernel(const float* __restrict__ sig, int H, int W, const float* __restrict__ wp_in_g,
               const float* __restrict__ last_obs, float length_ratio, float sigma_factor, float ratio, int rot,
               float* __restrict__ out) {
  const int b = blockIdx.x;
  const int S = H * W;
  const float* src = sig + (size_t)b * S;
  float* dst = out + (size_t)b * S;
  const CwsPrior p = make_prior(wp_in_g[2 * b + 0], wp_in_g[2 * b + 1], last_obs[2 * b + 0], last_obs[2 * b + 1],
                                length_ratio, sigma_factor, ratio, rot);
  // the reference first normalises the kernel (k / k.sum()), then the product (evaluate.py:34, 203-205)
  float ks = 0.f;
  for (int t = threadIdx.x; t < S; t += blockDim.x) ks += prior_value(p, t / W, t % W, H, W);
  __shared__ float sh[16];
  __shared__ float s_tot;
  ks = warp_sum(ks);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ks;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 16; ++w) t += sh[w];
    s_tot = t;
  }
  __syncthreads();
  const float inv_k = 1.0f / s_tot;
  float ws = 0.f;
  for (int t = threadIdx.x; t < S; t += blockDim.x) ws += src[t] * (prior_value(p, t / W, t % W, H, W) * inv_k);
  ws = warp_sum(ws);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ws;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 16; ++w) t += sh[w];
    s_tot = t;
  }
  __syncthreads();
  const float tot = s_tot;
  for (int t = threadIdx.x; t < S; t += blockDim.x)
    dst[t] = (src[t] * (prior_value(p, t / W, t % W, H, W) * inv_k)) / tot;
}

// a17: one warp per agent, lane = sampled future k (each lane keeps the reference's sequential sum over t; the min over
// k is order independent)
__global__ void ade_fde_kernel(const float* __restrict__ gt, const float* __restrict__ trajs,
                               const float* __restrict__ wps, int K, int B, int T, int n_wp, float resize,
                               float* __restrict__ ade, float* __restrict__ fde) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  float best_a = 3.4e38f, best_f = 3.4e38f;
  const float gx = gt[((size_t)b * T + T - 1) * 2 + 0], gy = gt[((size_t)b * T + T - 1) * 2 + 1];
  for (int k = lane; k < K; k += 32) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const float ex = (gt[((size_t)b * T + t) * 2 + 0] - trajs[(((size_t)k * B + b) * T + t) * 2 + 0]) / resize;
      const float ey = (gt[((size_t)b * T + t) * 2 + 1] - trajs[(((size_t)k * B + b) * T + t) * 2 + 1]) / resize;
      acc += sqrtf(ex * ex + ey * ey);
    }
    best_a = fminf(best_a, acc / (float)T);
    const float fx = (gx - wps[(((size_t)k * B + b) * n_wp + n_wp - 1) * 2 + 0]) / resize;
    const float fy = (gy - wps[(((size_t)k * B + b) * n_wp + n_wp - 1) * 2 + 1]) / resize;
    best_f = fminf(best_f, sqrtf(fx * fx + fy * fy));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    best_a = fminf(best_a, __shfl_xor_sync(0xffffffffu, best_a, o));
    best_f = fminf(best_f, __shfl_xor_sync(0xffffffffu, best_f, o));
  }
  if (lane == 0) {
    ade[b] = best_a;
    fde[b] = best_f;
  }
}

constexpr int kCwsSplits = 8;

}  // namespace ynet

using namespace ynet;

extern "C" {

int64_t ynet_cws_waypoint_workspace_bytes(int32_t B, int32_t G) {
  return (int64_t)B * kCwsSplits * G * 3 * (int64_t)sizeof(float);
}

int ynet_cws_waypoint(const float* sig, int32_t B, int32_t H, int32_t W, const float* wp_in, int32_t G,
                      const float* last_obs, float length_ratio, const float* sigma_factor, float ratio, int32_t rot,
                      float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  YNET_CHECK_ARG(sig && wp_in && last_obs && sigma_factor && out, "null pointer");
  YNET_CHECK_ARG(B > 0 && B <= 65535 && H > 1 && W > 1 && G > 0 && G <= 32, "bad shape (G <= 32 per call)");
  if (workspace == nullptr || workspace_bytes < ynet_cws_waypoint_workspace_bytes(B, G)) {
    set_error("ynet_cws_waypoint: workspace too small");
    return YNET_E_WORKSPACE;
  }
  float* partial = reinterpret_cast<float*>(workspace);
  cws_partial_kernel<<<dim3(kCwsSplits, B), 32 * G, 0, as_stream(stream)>>>(
      sig, H, W, wp_in, B, G, last_obs, length_ratio, sigma_factor, ratio, rot, kCwsSplits, partial);
  YNET_LAUNCH_CHECK();
  cws_finalize_kernel<<<ceil_div(B * G, 256), 256, 0, as_stream(stream)>>>(partial, B, G, kCwsSplits, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_cws_waypoint_map(const float* sig, int32_t B, int32_t H, int32_t W, const float* wp_in_g,
                          const float* last_obs, float length_ratio, float sigma_factor, float ratio, int32_t rot,
                          float* out, void* stream) {
  YNET_CHECK_ARG(sig && wp_in_g && last_obs && out, "null pointer");
  YNET_CHECK_ARG(B > 0 && H > 1 && W > 1, "bad shape");
  cws_map_kernel<<<B, 512, 0, as_stream(stream)>>>(sig, H, W, wp_in_g, last_obs, length_ratio, sigma_factor, ratio,
                                                   rot, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_ade_fde(const float* gt, const float* trajs, const float* wps, int32_t K, int32_t B, int32_t T, int32_t n_wp,
                 float resize_factor, float* ade, float* fde, void* stream) {
  YNET_CHECK_ARG(gt && trajs && wps && ade && fde, "null pointer");
  YNET_CHECK_ARG(K > 0 && B >= 0 && T > 0 && n_wp > 0 && resize_factor > 0.f, "bad shape");
  if (B == 0) return YNET_OK;
  ade_fde_kernel<<<ceil_div(B, 4), 128, 0, as_stream(stream)>>>(gt, trajs, wps, K, B, T, n_wp, resize_factor, ade,
                                                                fde);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

}  // extern "C"
