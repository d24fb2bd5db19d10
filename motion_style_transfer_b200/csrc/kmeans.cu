// a14: batched Lloyd k-means for TTST (utils/kmeans.py:22-108), one CTA per agent.
//
// The reference clusters each agent's 10 000 sampled goals in a Python loop that bounces every
// iteration through the CPU (kmeans.py:146-148).  Here the points of one agent (80 KB) are staged
// once into shared memory and all iterations run out of SMEM; agents run concurrently on all SMs.
// Arithmetic follows the oracle bit for bit for integer-valued pixel coordinates: distances
// fl(fl(dx*dx)+fl(dy*dy)) without FMA contraction, first-minimum argmin, exact sums (< 2^24),
// one IEEE division per centre, sequential shift sum, stop when shift^2 < tol.
#include <float.h>

#include "common.cuh"

namespace ynet {

constexpr int kKmThreads = 512;
constexpr int kKmMaxK = 64;

__global__ void __launch_bounds__(kKmThreads)
kmeans_kernel(const float* __restrict__ X, int N, int K, const int* __restrict__ init_idx,
              const int* __restrict__ reseed_idx, int R, float tol, int iter_limit, float* __restrict__ centres,
              int* __restrict__ assign, int* __restrict__ iters, int* __restrict__ status, int points_in_smem) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float cx[kKmMaxK], cy[kKmMaxK];
  __shared__ float sumx[kKmMaxK], sumy[kKmMaxK];
  __shared__ int cnt[kKmMaxK];
  __shared__ int s_done, s_reseed_used, s_status;

  const int b = blockIdx.x;
  const float2* Xg = reinterpret_cast<const float2*>(X) + (size_t)b * N;
  float2* Xs = reinterpret_cast<float2*>(smem);
  if (points_in_smem)
    for (int i = threadIdx.x; i < N; i += kKmThreads) Xs[i] = Xg[i];
  const float2* P = points_in_smem ? Xs : Xg;
  if (threadIdx.x < K) {
    const float2 c = Xg[init_idx[(size_t)b * K + threadIdx.x]];
    cx[threadIdx.x] = c.x;
    cy[threadIdx.x] = c.y;
  }
  if (threadIdx.x == 0) {
    s_done = 0;
    s_reseed_used = 0;
    s_status = 0;
  }
  __syncthreads();

  int it = 0;
  while (true) {
    if (threadIdx.x < K) {
      sumx[threadIdx.x] = 0.f;
      sumy[threadIdx.x] = 0.f;
      cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += kKmThreads) {
      const float2 pt = P[i];
      float best = FLT_MAX;
      int bk = 0;
      for (int k = 0; k < K; ++k) {
        const float dx = __fsub_rn(pt.x, cx[k]);
        const float dy = __fsub_rn(pt.y, cy[k]);
        const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (d < best) {  // strict: first minimum wins (torch.argmin)
          best = d;
          bk = k;
        }
      }
      if (assign != nullptr) assign[(size_t)b * N + i] = bk;
      atomicAdd(&sumx[bk], pt.x);
      atomicAdd(&sumy[bk], pt.y);
      atomicAdd(&cnt[bk], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float shift = 0.f;
      for (int k = 0; k < K; ++k) {
        float nx, ny;
        if (cnt[k] == 0) {  // kmeans.py:82-83: empty cluster -> X[randint]
          int ridx = 0;
          if (reseed_idx != nullptr && s_reseed_used < R) {
            ridx = reseed_idx[(size_t)b * R + s_reseed_used];
            ridx = min(max(ridx, 0), N - 1);
          } else {
            s_status |= 1;
          }
          s_reseed_used++;
          const float2 pt = Xg[ridx];
          nx = pt.x;  // mean of a single point
          ny = pt.y;
        } else {
          const float c = (float)cnt[k];
          nx = __fdiv_rn(sumx[k], c);
          ny = __fdiv_rn(sumy[k], c);
        }
        const float ddx = __fsub_rn(nx, cx[k]), ddy = __fsub_rn(ny, cy[k]);
        const float d2 = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
        shift = __fadd_rn(shift, __fsqrt_rn(d2));
        cx[k] = nx;
        cy[k] = ny;
      }
      const bool stop = (__fmul_rn(shift, shift) < tol) || (iter_limit != 0 && it + 1 >= iter_limit);
      s_done = stop ? 1 : 0;
    }
    __syncthreads();
    ++it;
    if (s_done) break;
  }
  if (threadIdx.x < K) {
    centres[((size_t)b * K + threadIdx.x) * 2 + 0] = cx[threadIdx.x];
    centres[((size_t)b * K + threadIdx.x) * 2 + 1] = cy[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    if (iters != nullptr) iters[b] = it;
    if (status != nullptr) status[b] = s_status;
  }
}

}  // namespace ynet

using namespace ynet;

extern "C" int ynet_kmeans_batched(const float* X, int32_t B, int32_t N, int32_t K, const int32_t* init_idx,
                                   const int32_t* reseed_idx, int32_t R, float tol, int32_t iter_limit,
                                   float* centres, int32_t* assign, int32_t* iters, int32_t* status, void* stream) {
  YNET_CHECK_ARG(X && init_idx && centres, "null pointer");
  YNET_CHECK_ARG(B >= 0 && N > 0 && K > 0 && K <= kKmMaxK && K <= N, "bad shape (K <= 64, K <= N)");
  YNET_CHECK_ALIGN(X, 8);
  if (B == 0) return YNET_OK;
  const size_t need = (size_t)N * sizeof(float2);
  const int in_smem = need <= 200 * 1024 ? 1 : 0;
  const size_t dyn = in_smem ? need : 0;
  static size_t configured = 0;
  if (dyn > 48 * 1024 && dyn > configured) {
    cudaError_t e = cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "ynet_kmeans_batched(cudaFuncSetAttribute)");
    configured = 200 * 1024;
  }
  kmeans_kernel<<<B, kKmThreads, dyn, as_stream(stream)>>>(X, N, K, init_idx, reseed_idx, R, tol, iter_limit, centres,
                                                           assign, iters, status, in_smem);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}
