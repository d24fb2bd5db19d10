// a14: batched Lloyd k-means for TTST (utils/kmeans.py:22-108), one CTA per agent.
//
// The reference clusters each agent's 10 000 sampled goals in a Python loop that bounces every
// iteration through the CPU (kmeans.py:146-148).  Here the points of one agent (80 KB) are staged
// once into shared memory and all iterations run out of SMEM; agents run concurrently on all SMs.
// Arithmetic follows the oracle bit for bit for integer-valued pixel coordinates: distances
// fl(fl(dx*dx)+fl(dy*dy)) without FMA contraction, first-minimum argmin, exact sums (< 2^24),
// one IEEE division per centre, sequential shift sum, stop when shift^2 < tol.
//
// Cluster sums are accumulated in REGISTERS (K compile-time bounded, fully unrolled select-adds) and
// reduced with warp shuffles; shared-memory atomics on 19 hot addresses serialised the first version
// (75 us per iteration).
#include <float.h>

#include "common.cuh"

namespace ynet {

constexpr int kKmThreads = 512;
constexpr int kKmMaxK = 64;

struct KmShared {
  float cx[kKmMaxK], cy[kKmMaxK];
  float sumx[kKmMaxK], sumy[kKmMaxK];
  int cnt[kKmMaxK];
  int done, reseed_used, status;
};

// thread 0: new centres, empty-cluster reseed (kmeans.py:82-83), shift and the stop test
__device__ __forceinline__ void km_update_centres(KmShared& sh, const float2* Xg, int N, int K, const int* reseed_idx,
                                                  int R, int b, float tol, int iter_limit, int it) {
  float shift = 0.f;
  for (int k = 0; k < K; ++k) {
    float nx, ny;
    if (sh.cnt[k] == 0) {
      int ridx = 0;
      if (reseed_idx != nullptr && sh.reseed_used < R) {
        ridx = reseed_idx[(size_t)b * R + sh.reseed_used];
        ridx = min(max(ridx, 0), N - 1);
      } else {
        sh.status |= 1;
      }
      sh.reseed_used++;
      const float2 pt = Xg[ridx];
      nx = pt.x;  // mean of a single point
      ny = pt.y;
    } else {
      const float c = (float)sh.cnt[k];
      nx = __fdiv_rn(sh.sumx[k], c);
      ny = __fdiv_rn(sh.sumy[k], c);
    }
    const float ddx = __fsub_rn(nx, sh.cx[k]), ddy = __fsub_rn(ny, sh.cy[k]);
    const float d2 = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
    shift = __fadd_rn(shift, __fsqrt_rn(d2));
    sh.cx[k] = nx;
    sh.cy[k] = ny;
  }
  const bool stop = (__fmul_rn(shift, shift) < tol) || (iter_limit != 0 && it + 1 >= iter_limit);
  sh.done = stop ? 1 : 0;
}

template <int KMAX>
__global__ void __launch_bounds__(kKmThreads)
kmeans_kernel(const float* __restrict__ X, int N, int K, const int* __restrict__ init_idx,
              const int* __restrict__ reseed_idx, int R, float tol, int iter_limit, float* __restrict__ centres,
              int* __restrict__ assign, int* __restrict__ iters, int* __restrict__ status, int points_in_smem) {
  extern __shared__ __align__(16) float smem[];
  __shared__ KmShared sh;

  const int b = blockIdx.x;
  const float2* Xg = reinterpret_cast<const float2*>(X) + (size_t)b * N;
  float2* Xs = reinterpret_cast<float2*>(smem);
  if (points_in_smem)
    for (int i = threadIdx.x; i < N; i += kKmThreads) Xs[i] = Xg[i];
  const float2* P = points_in_smem ? Xs : Xg;
  if (threadIdx.x < K) {
    const float2 c = Xg[init_idx[(size_t)b * K + threadIdx.x]];
    sh.cx[threadIdx.x] = c.x;
    sh.cy[threadIdx.x] = c.y;
  }
  if (threadIdx.x == 0) {
    sh.done = 0;
    sh.reseed_used = 0;
    sh.status = 0;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  int it = 0;
  while (true) {
    if (threadIdx.x < K) {
      sh.sumx[threadIdx.x] = 0.f;
      sh.sumy[threadIdx.x] = 0.f;
      sh.cnt[threadIdx.x] = 0;
    }
    float ax[KMAX], ay[KMAX];
    int an[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      ax[k] = 0.f;
      ay[k] = 0.f;
      an[k] = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += kKmThreads) {
      const float2 pt = P[i];
      float best = FLT_MAX;
      int bk = 0;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          const float dx = __fsub_rn(pt.x, sh.cx[k]);
          const float dy = __fsub_rn(pt.y, sh.cy[k]);
          const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          if (d < best) {  // strict: first minimum wins (torch.argmin)
            best = d;
            bk = k;
          }
        }
      }
      if (assign != nullptr) assign[(size_t)b * N + i] = bk;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const bool hit = (bk == k);
        ax[k] += hit ? pt.x : 0.f;   // integer-valued coordinates: exact in any order (< 2^24)
        ay[k] += hit ? pt.y : 0.f;
        an[k] += hit ? 1 : 0;
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        float vx = ax[k], vy = ay[k];
        int vn = an[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          vx += __shfl_xor_sync(0xffffffffu, vx, o);
          vy += __shfl_xor_sync(0xffffffffu, vy, o);
          vn += __shfl_xor_sync(0xffffffffu, vn, o);
        }
        if (lane == 0 && vn > 0) {
          atomicAdd(&sh.sumx[k], vx);
          atomicAdd(&sh.sumy[k], vy);
          atomicAdd(&sh.cnt[k], vn);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) km_update_centres(sh, Xg, N, K, reseed_idx, R, b, tol, iter_limit, it);
    __syncthreads();
    ++it;
    if (sh.done) break;
  }
  if (threadIdx.x < K) {
    centres[((size_t)b * K + threadIdx.x) * 2 + 0] = sh.cx[threadIdx.x];
    centres[((size_t)b * K + threadIdx.x) * 2 + 1] = sh.cy[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    if (iters != nullptr) iters[b] = it;
    if (status != nullptr) status[b] = sh.status;
  }
}

// K > 32: shared-memory atomics per point (no register-resident accumulators)
__global__ void __launch_bounds__(kKmThreads)
kmeans_atomic_kernel(const float* __restrict__ X, int N, int K, const int* __restrict__ init_idx,
                     const int* __restrict__ reseed_idx, int R, float tol, int iter_limit, float* __restrict__ centres,
                     int* __restrict__ assign, int* __restrict__ iters, int* __restrict__ status, int points_in_smem) {
  extern __shared__ __align__(16) float smem[];
  __shared__ KmShared sh;
  const int b = blockIdx.x;
  const float2* Xg = reinterpret_cast<const float2*>(X) + (size_t)b * N;
  float2* Xs = reinterpret_cast<float2*>(smem);
  if (points_in_smem)
    for (int i = threadIdx.x; i < N; i += kKmThreads) Xs[i] = Xg[i];
  const float2* P = points_in_smem ? Xs : Xg;
  if (threadIdx.x < K) {
    const float2 c = Xg[init_idx[(size_t)b * K + threadIdx.x]];
    sh.cx[threadIdx.x] = c.x;
    sh.cy[threadIdx.x] = c.y;
  }
  if (threadIdx.x == 0) {
    sh.done = 0;
    sh.reseed_used = 0;
    sh.status = 0;
  }
  __syncthreads();
  int it = 0;
  while (true) {
    if (threadIdx.x < K) {
      sh.sumx[threadIdx.x] = 0.f;
      sh.sumy[threadIdx.x] = 0.f;
      sh.cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += kKmThreads) {
      const float2 pt = P[i];
      float best = FLT_MAX;
      int bk = 0;
      for (int k = 0; k < K; ++k) {
        const float dx = __fsub_rn(pt.x, sh.cx[k]);
        const float dy = __fsub_rn(pt.y, sh.cy[k]);
        const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (d < best) {
          best = d;
          bk = k;
        }
      }
      if (assign != nullptr) assign[(size_t)b * N + i] = bk;
      atomicAdd(&sh.sumx[bk], pt.x);
      atomicAdd(&sh.sumy[bk], pt.y);
      atomicAdd(&sh.cnt[bk], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) km_update_centres(sh, Xg, N, K, reseed_idx, R, b, tol, iter_limit, it);
    __syncthreads();
    ++it;
    if (sh.done) break;
  }
  if (threadIdx.x < K) {
    centres[((size_t)b * K + threadIdx.x) * 2 + 0] = sh.cx[threadIdx.x];
    centres[((size_t)b * K + threadIdx.x) * 2 + 1] = sh.cy[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    if (iters != nullptr) iters[b] = it;
    if (status != nullptr) status[b] = sh.status;
  }
}

}  // namespace ynet

using namespace ynet;

extern "C" int ynet_kmeans_batched(const float* X, int32_t B, int32_t N, int32_t K, const int32_t* init_idx,
                                   const int32_t* reseed_idx, int32_t R, float tol, int32_t iter_limit,
                                   float* centres, int32_t* assign, int32_t* iters, int32_t* status, void* stream) {
  YNET_CHECK_ARG(B >= 0 && N > 0 && K > 0 && K <= kKmMaxK && K <= N, "bad shape (K <= 64, K <= N)");
  if (B == 0) return YNET_OK;
  YNET_CHECK_ARG(X && init_idx && centres, "null pointer");
  YNET_CHECK_ALIGN(X, 8);
  const size_t need = (size_t)N * sizeof(float2);
  const int in_smem = need <= 200 * 1024 ? 1 : 0;
  const size_t dyn = in_smem ? need : 0;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kmeans_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kmeans_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kmeans_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kmeans_atomic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "ynet_kmeans_batched(cudaFuncSetAttribute)");
    configured = true;
  }
  cudaStream_t st = as_stream(stream);
#define KM_ARGS X, N, K, init_idx, reseed_idx, R, tol, iter_limit, centres, assign, iters, status, in_smem
  if (K <= 8)
    kmeans_kernel<8><<<B, kKmThreads, dyn, st>>>(KM_ARGS);
  else if (K <= 20)
    kmeans_kernel<20><<<B, kKmThreads, dyn, st>>>(KM_ARGS);
  else if (K <= 32)
    kmeans_kernel<32><<<B, kKmThreads, dyn, st>>>(KM_ARGS);
  else
    kmeans_atomic_kernel<<<B, kKmThreads, dyn, st>>>(KM_ARGS);
#undef KM_ARGS
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}
