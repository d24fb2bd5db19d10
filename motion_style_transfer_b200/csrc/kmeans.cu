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
#include <cooperative_groups.h>
#include <float.h>

#include <stdlib.h>
#include <string.h>

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

// Packed per-cluster accumulator of integer-valued points: sum x in bits 0-23, sum y in 24-47, count in 48-63.
// Valid when every coordinate is a non-negative integer with N * max < 2^24 and N < 2^16 (TTST: 10 000 pixel
// coordinates < 1386) -- then the sums are exact in any order, like the reference's fp32 sums.
__device__ __forceinline__ unsigned long long km_pack(float x, float y) {
  return (unsigned long long)(unsigned)x | ((unsigned long long)(unsigned)y << 24) | (1ull << 48);
}

// One CLUSTER of `cs` CTAs per agent (cs = 1, 2 or 4, chosen so that B * cs fills the SMs): each CTA keeps its slice of
// the points in shared memory for all iterations.  Per iteration and point: K distance evaluations against centres
// held in REGISTERS, then ONE 64-bit read-modify-write of the thread's private bin acc[k][tid] (conflict-free; the
// first version spent as many instructions on K select-adds per point as on the distances).  Bins are reduced per
// CTA, exchanged through distributed shared memory (double-buffered by iteration parity: one cluster barrier per
// iteration) and every CTA of the cluster recomputes the identical centre update.
// Points per thread in flight: the argmin over the centres is a dependent compare-select chain per point; four independent
// chains hide its latency at 16 warps per SM (14.4 -> 11.5 us per Lloyd iteration at 10 000 points, K = 19; the packed
// FADD2 / FMUL2 form of the distance was tried on top and changed nothing: the loop is not bound by FP32 issue).
constexpr int kKmUnroll = 4;

template <int KMAX>
__global__ void __launch_bounds__(kKmThreads)
kmeans_kernel(const float* __restrict__ X, int N, int K, const int* __restrict__ init_idx,
              const int* __restrict__ reseed_idx, int R, float tol, int iter_limit, float* __restrict__ centres,
              int* __restrict__ assign, int* __restrict__ iters, int* __restrict__ status, int cs, int per) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char km_smem[];
  __shared__ KmShared sh;
  __shared__ unsigned long long red[2][kKmMaxK];    // this CTA's packed sums (exact mode), by iteration parity
  __shared__ float fred[2][kKmMaxK][2];             // generic mode: float sums ...
  __shared__ int cred[2][kKmMaxK];                  // ... and counts
  __shared__ float term[kKmMaxK];
  __shared__ int s_exact;

  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / cs;
  const float2* Xg = reinterpret_cast<const float2*>(X) + (size_t)b * N;
  float2* Xs = reinterpret_cast<float2*>(km_smem);
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(km_smem + (((size_t)per * 8 + 15) & ~(size_t)15));
  const int i0 = rank * per;
  const int nloc = max(0, min(N, i0 + per) - i0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // stage the slice; decide exact-integer mode (uniform over the cluster)
  int ok = 1;
  float cmax = 0.f;
  for (int i = tid; i < nloc; i += kKmThreads) {
    const float2 pt = Xg[i0 + i];
    Xs[i] = pt;
    ok &= (pt.x >= 0.f && pt.y >= 0.f && pt.x == floorf(pt.x) && pt.y == floorf(pt.y) && pt.x < 16777216.f &&
           pt.y < 16777216.f) ? 1 : 0;
    cmax = fmaxf(cmax, fmaxf(pt.x, pt.y));
  }
  ok &= (cmax * (float)N < 16777216.f && N < 65536) ? 1 : 0;
  ok = __syncthreads_and(ok);
  if (tid == 0) s_exact = ok;
  if (tid < K) {
    const float2 c = Xg[init_idx[(size_t)b * K + tid]];
    sh.cx[tid] = c.x;
    sh.cy[tid] = c.y;
  }
  if (tid == 0) {
    sh.done = 0;
    sh.reseed_used = 0;
    sh.status = 0;
  }
  cluster.sync();
  bool exact = true;
  for (int r = 0; r < cs; ++r) exact = exact && (*cluster.map_shared_rank(&s_exact, r) != 0);

  int it = 0;
  while (true) {
    const int par = it & 1;
    float cx[KMAX], cy[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      cx[k] = (k < K) ? sh.cx[k] : 0.f;
      cy[k] = (k < K) ? sh.cy[k] : 0.f;
    }
    if (exact) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) acc[k * kKmThreads + tid] = 0ull;
    } else if (tid < K) {
      sh.sumx[tid] = 0.f;
      sh.sumy[tid] = 0.f;
      sh.cnt[tid] = 0;
    }
    __syncthreads();
    // kKmUnroll points per thread at a time: their argmin chains are independent
    for (int i = tid; i < nloc; i += kKmUnroll * kKmThreads) {
      float2 pt[kKmUnroll];
      float best[kKmUnroll];
      int bk[kKmUnroll];
#pragma unroll
      for (int u = 0; u < kKmUnroll; ++u) {
        const int iu = i + u * kKmThreads;
        pt[u] = Xs[iu < nloc ? iu : i];
        best[u] = FLT_MAX;
        bk[u] = 0;
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
#pragma unroll
          for (int u = 0; u < kKmUnroll; ++u) {
            const float dx = __fsub_rn(pt[u].x, cx[k]);
            const float dy = __fsub_rn(pt[u].y, cy[k]);
            const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            if (d < best[u]) {  // strict: first minimum wins (torch.argmin)
              best[u] = d;
              bk[u] = k;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kKmUnroll; ++u) {
        const int iu = i + u * kKmThreads;
        if (iu < nloc) {
          if (assign != nullptr) assign[(size_t)b * N + i0 + iu] = bk[u];
          if (exact) {
            acc[bk[u] * kKmThreads + tid] += km_pack(pt[u].x, pt[u].y);
          } else {
            atomicAdd(&sh.sumx[bk[u]], pt[u].x);
            atomicAdd(&sh.sumy[bk[u]], pt[u].y);
            atomicAdd(&sh.cnt[bk[u]], 1);
          }
        }
      }
    }
    __syncthreads();
    if (exact) {
      for (int k = warp; k < K; k += kKmThreads / 32) {
        unsigned long long v = 0ull;
#pragma unroll
        for (int j = 0; j < kKmThreads / 32; ++j) v += acc[k * kKmThreads + j * 32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[par][k] = v;
      }
    } else if (tid < K) {
      fred[par][tid][0] = sh.sumx[tid];
      fred[par][tid][1] = sh.sumy[tid];
      cred[par][tid] = sh.cnt[tid];
    }
    cluster.sync();      // every CTA's partial sums of this iteration are visible cluster-wide
    if (warp == 0) {
      // lane k: new centre k (kmeans.py:72-84); empty clusters take the next reseed indices in k order
      float sumx = 0.f, sumy = 0.f;
      int cnt = 0;
      if (lane < K) {
        if (exact) {
          unsigned long long tot = 0ull;
          for (int r = 0; r < cs; ++r) tot += *cluster.map_shared_rank(&red[par][lane], r);
          sumx = (float)(unsigned)(tot & 0xFFFFFFull);
          sumy = (float)(unsigned)((tot >> 24) & 0xFFFFFFull);
          cnt = (int)(tot >> 48);
        } else {
          for (int r = 0; r < cs; ++r) {
            sumx += *cluster.map_shared_rank(&fred[par][lane][0], r);
            sumy += *cluster.map_shared_rank(&fred[par][lane][1], r);
            cnt += *cluster.map_shared_rank(&cred[par][lane], r);
          }
        }
      }
      const unsigned empty = __ballot_sync(0xffffffffu, lane < K && cnt == 0);
      const int used0 = sh.reseed_used;
      float t = 0.f;
      if (lane < K) {
        float nx, ny;
        if (cnt == 0) {
          const int slot = used0 + __popc(empty & ((1u << lane) - 1u));
          int ridx = 0;
          if (reseed_idx != nullptr && slot < R) {
            ridx = reseed_idx[(size_t)b * R + slot];
            ridx = min(max(ridx, 0), N - 1);
          } else {
            atomicOr(&sh.status, 1);
          }
          const float2 pt = Xg[ridx];
          nx = pt.x;  // mean of a single point
          ny = pt.y;
        } else {
          const float c = (float)cnt;
          nx = __fdiv_rn(sumx, c);
          ny = __fdiv_rn(sumy, c);
        }
        const float ddx = __fsub_rn(nx, sh.cx[lane]), ddy = __fsub_rn(ny, sh.cy[lane]);
        t = __fsqrt_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));
        sh.cx[lane] = nx;
        sh.cy[lane] = ny;
        term[lane] = t;
      }
      __syncwarp();
      if (lane == 0) {
        sh.reseed_used = used0 + __popc(empty);
        float shift = 0.f;
        for (int k = 0; k < K; ++k) shift = __fadd_rn(shift, term[k]);     // sequential, like the reference's sum
        const bool stop = (__fmul_rn(shift, shift) < tol) || (iter_limit != 0 && it + 1 >= iter_limit);
        sh.done = stop ? 1 : 0;
      }
    }
    __syncthreads();
    ++it;
    if (sh.done) break;
  }
  if (rank == 0) {
    if (tid < K) {
      centres[((size_t)b * K + tid) * 2 + 0] = sh.cx[tid];
      centres[((size_t)b * K + tid) * 2 + 1] = sh.cy[tid];
    }
    if (tid == 0) {
      if (iters != nullptr) iters[b] = it;
      if (status != nullptr) status[b] = sh.status;
    }
  }
  cluster.sync();        // no CTA leaves while a peer may still read its shared memory
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

template <int KMAX>
static cudaError_t km_launch(int B, int cs, size_t dyn, cudaStream_t st, const float* X, int N, int K, const int* init_idx,
                             const int* reseed_idx, int R, float tol, int iter_limit, float* centres, int* assign, int* iters,
                             int* status, int per) {
  cudaError_t e = cudaFuncSetAttribute(kmeans_kernel<KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(B * cs));
  cfg.blockDim = dim3(kKmThreads);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = (unsigned)cs;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kmeans_kernel<KMAX>, X, N, K, init_idx, reseed_idx, R, tol, iter_limit, centres, assign,
                            iters, status, cs, per);
}

extern "C" int ynet_kmeans_batched(const float* X, int32_t B, int32_t N, int32_t K, const int32_t* init_idx,
                                   const int32_t* reseed_idx, int32_t R, float tol, int32_t iter_limit,
                                   float* centres, int32_t* assign, int32_t* iters, int32_t* status, void* stream) {
  YNET_CHECK_ARG(B >= 0 && N > 0 && K > 0 && K <= kKmMaxK && K <= N, "bad shape (K <= 64, K <= N)");
  if (B == 0) return YNET_OK;
  YNET_CHECK_ARG(X && init_idx && centres, "null pointer");
  YNET_CHECK_ALIGN(X, 8);
  cudaStream_t st = as_stream(stream);
  if (K <= 32) {
    // cluster size: spread one agent over 2 or 4 SMs while B * cs still fits the machine in one wave
    const int kmax = K <= 8 ? 8 : (K <= 20 ? 20 : 32);
    int cs = 1;
    while (cs < 4 && (long long)B * cs * 2 <= sm_count() && N / (cs * 2) >= 4 * kKmThreads) cs *= 2;
    if (const char* e = getenv("YNET_KMEANS_CLUSTER")) cs = tmax(1, tmin(4, atoi(e)));
    int per = ceil_div(N, cs);
    size_t dyn = (((size_t)per * 8 + 15) & ~(size_t)15) + (size_t)kmax * kKmThreads * 8;
    while (dyn > 220 * 1024 && cs < 8) {      // very large N: more CTAs per agent so that a slice fits shared memory
      cs *= 2;
      per = ceil_div(N, cs);
      dyn = (((size_t)per * 8 + 15) & ~(size_t)15) + (size_t)kmax * kKmThreads * 8;
    }
    if (dyn <= 220 * 1024) {
      cudaError_t e;
      if (kmax == 8)
        e = km_launch<8>(B, cs, dyn, st, X, N, K, init_idx, reseed_idx, R, tol, iter_limit, centres, assign, iters, status, per);
      else if (kmax == 20)
        e = km_launch<20>(B, cs, dyn, st, X, N, K, init_idx, reseed_idx, R, tol, iter_limit, centres, assign, iters, status, per);
      else
        e = km_launch<32>(B, cs, dyn, st, X, N, K, init_idx, reseed_idx, R, tol, iter_limit, centres, assign, iters, status, per);
      if (e != cudaSuccess) return cuda_fail(e, "ynet_kmeans_batched");
      return YNET_OK;
    }
  }
  const size_t need = (size_t)N * sizeof(float2);
  const int in_smem = need <= 200 * 1024 ? 1 : 0;
  const size_t dyn = in_smem ? need : 0;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kmeans_atomic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "ynet_kmeans_batched(cudaFuncSetAttribute)");
    configured = true;
  }
  kmeans_atomic_kernel<<<B, kKmThreads, dyn, st>>>(X, N, K, init_idx, reseed_idx, R, tol, iter_limit, centres, assign, iters,
                                                   status, in_smem);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}
