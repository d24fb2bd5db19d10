// a11: TTST / goal sampling = threshold + normalise + torch.multinomial (CPU semantics) + unravel.
//
// Bit-exactness contract: given the same probability map and the same supplied randoms
// the indices equal ATen's CPU multinomial.  With replacement that kernel builds the CDF as a
// SEQUENTIAL float32 running sum, which no parallel scan reproduces, so the running sum is done by
// one lane per row (4-cycle dependent FADD chain) while the other 31 lanes stage the row through
// shared memory with coalesced loads/stores; rows run concurrently on all SMs.
#include <float.h>

#include "common.cuh"

namespace ynet {

// ---- prepare: per-row max and masked global sum (image_utils.py:114-119) -------------------------
__global__ void __launch_bounds__(1024) rowmax_kernel(const float* __restrict__ p, long long S,
                                                      float* __restrict__ rowmax) {
  const float* src = p + (size_t)blockIdx.x * S;
  float m = -FLT_MAX;
  for (long long i = threadIdx.x; i < S; i += blockDim.x) m = fmaxf(m, src[i]);
  m = warp_max(m);
  __shared__ float sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : -FLT_MAX;
    v = warp_max(v);
    if (threadIdx.x == 0) rowmax[blockIdx.x] = v;
  }
}

constexpr int kSumChunks = 8;

// grid = (kSumChunks, rows): fp64 partial sums of the kept entries, fixed reduction order.
__global__ void __launch_bounds__(256)
masked_sum_partial_kernel(const float* __restrict__ p, long long S, float rel, const float* __restrict__ rowmax,
                          double* __restrict__ partial) {
  const int row = blockIdx.y;
  const float* src = p + (size_t)row * S;
  const float thr = rowmax[row] * rel;
  const long long per = ceil_div<long long>(S, kSumChunks);
  const long long i0 = blockIdx.x * per, i1 = min(S, i0 + per);
  double acc = 0.0;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const float v = src[i];
    acc += (v < thr) ? 0.0 : (double)v;
  }
  acc = warp_sum(acc);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[(size_t)row * kSumChunks + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) masked_sum_final_kernel(const double* __restrict__ partial, long long n,
                                                               float* __restrict__ gsum) {
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  acc = warp_sum(acc);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    *gsum = (float)t;
  }
}

__device__ __forceinline__ float normalised(float v, bool use_thr, float thr, float gsum) {
  if (!use_thr) return v;
  const float kept = (v < thr) ? __fmul_rn(v, 0.0f) : v;  // prob * (~mask).int()
  return __fdiv_rn(kept, gsum);                           // prob / prob.sum()
}

// ---- sequential float32 CDF: one CTA per row, warp-specialised -------------------------------------
// torch.multinomial's CPU CDF is a sequential fp32 running sum, so each row is ONE dependent FADD chain
// (173 056 adds at 4 cycles each is the floor).  Everything else is taken off that chain: lane 0 of warp 0
// only walks a staged chunk in shared memory; warps 1-3 meanwhile write the previous chunk's prefix sums
// to HBM and load + threshold + normalise the next chunk into the buffer that just became free.
constexpr int kCdfChunk = 2048;       // elements per staged chunk (two buffers)
constexpr int kCdfThreads = 128;
constexpr int kCdfHelpers = kCdfThreads - 32;

__global__ void __launch_bounds__(kCdfThreads)
cdf_sequential_kernel(const float* __restrict__ p, int rows, long long S, bool use_thr, float rel,
                      const float* __restrict__ rowmax, const float* __restrict__ gsum_ptr, float* __restrict__ cdf) {
  __shared__ __align__(16) float buf[2][kCdfChunk];
  const int row = blockIdx.x;
  const float* src = p + (size_t)row * S;
  float* dst = cdf + (size_t)row * S;
  const float thr = use_thr ? __fmul_rn(rowmax[row], rel) : 0.f;
  const float gsum = use_thr ? *gsum_ptr : 1.f;
  const long long n_chunks = ceil_div<long long>(S, kCdfChunk);
  const int helper = (int)threadIdx.x - 32;          // >= 0: staging / write-out thread

  auto stage = [&](long long c) {                    // chunk c -> buf[c & 1] (zero tail: x + 0 is exact)
    float* b = buf[c & 1];
    const long long base = c * kCdfChunk;
    constexpr int PER = (kCdfChunk + kCdfHelpers - 1) / kCdfHelpers;
    float r[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {                  // all loads in flight before the first division
      const int i = helper + k * kCdfHelpers;
      r[k] = (i < kCdfChunk && base + i < S) ? __ldg(src + base + i) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = helper + k * kCdfHelpers;
      if (i < kCdfChunk) b[i] = (base + i < S) ? normalised(r[k], use_thr, thr, gsum) : 0.f;
    }
  };
  if (helper >= 0) stage(0);
  __syncthreads();
  float carry = 0.f;
  for (long long c = 0; c < n_chunks; ++c) {
    if (threadIdx.x == 0) {
      float4* b4 = reinterpret_cast<float4*>(buf[c & 1]);
      float acc = carry;
#pragma unroll 16
      for (int i = 0; i < kCdfChunk / 4; ++i) {
        float4 v = b4[i];
        v.x = acc = __fadd_rn(acc, v.x);
        v.y = acc = __fadd_rn(acc, v.y);
        v.z = acc = __fadd_rn(acc, v.z);
        v.w = acc = __fadd_rn(acc, v.w);
        b4[i] = v;
      }
      carry = acc;
    } else if (helper >= 0) {
      if (c > 0) {                                   // prefix sums of chunk c - 1 -> HBM
        const float* b = buf[(c - 1) & 1];
        const long long base = (c - 1) * kCdfChunk;
        for (int i = helper; i < kCdfChunk; i += kCdfHelpers)
          if (base + i < S) dst[base + i] = b[i];
      }
      if (c + 1 < n_chunks) stage(c + 1);            // same buffer as chunk c - 1, now free
    }
    __syncthreads();
  }
  if (helper >= 0) {
    const float* b = buf[(n_chunks - 1) & 1];
    const long long base = (n_chunks - 1) * kCdfChunk;
    for (int i = helper; i < kCdfChunk; i += kCdfHelpers)
      if (base + i < S) dst[base + i] = b[i];
  }
}

// ---- inverse-CDF lookup: lower bound of u in c / c[S-1] (last bucket forced to 1) ------------------
__global__ void __launch_bounds__(256)
cdf_search_kernel(const float* __restrict__ cdf, long long S, const double* __restrict__ uniforms, int n,
                  long long* __restrict__ idx, float* __restrict__ xy, int W) {
  const int row = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* c = cdf + (size_t)row * S;
  const float total = c[S - 1];
  const double u = uniforms[(size_t)row * n + j];
  long long lo = 0, hi = S;
  while (hi - lo > 0) {
    const long long mid = lo + (hi - lo) / 2;
    const float cp = (mid == S - 1) ? 1.0f : __fdiv_rn(c[mid], total);
    if ((double)cp < u)
      lo = mid + 1;
    else
      hi = mid;
  }
  const size_t o = (size_t)row * n + j;
  idx[o] = lo;
  if (xy != nullptr) {
    xy[2 * o + 0] = (float)(lo % W);
    xy[2 * o + 1] = floorf(__fdiv_rn((float)lo, (float)W));
  }
}

// ---- without replacement / n == 1: top-n of p / q -------------------------------------------------
constexpr int kTopkThreads = 512;

__global__ void __launch_bounds__(kTopkThreads)
topk_ratio_kernel(const float* __restrict__ p, const float* __restrict__ q, long long S, bool use_thr, float rel,
                  const float* __restrict__ rowmax, const float* __restrict__ gsum_ptr, int n,
                  long long* __restrict__ idx, float* __restrict__ xy, int W) {
  const int row = blockIdx.x;
  const float* pr = p + (size_t)row * S;
  const float* qr = q + (size_t)row * S;
  const float thr = use_thr ? __fmul_rn(rowmax[row], rel) : 0.f;
  const float gsum = use_thr ? *gsum_ptr : 1.f;
  __shared__ unsigned long long sh[kTopkThreads / 32];
  __shared__ unsigned long long s_prev;
  if (threadIdx.x == 0) s_prev = ~0ull;
  __syncthreads();
  for (int t = 0; t < n; ++t) {
    const unsigned long long prev = s_prev;
    unsigned long long best = 0ull;
    for (long long i = threadIdx.x; i < S; i += kTopkThreads) {
      const float r = __fdiv_rn(normalised(pr[i], use_thr, thr, gsum), qr[i]);
      // non-negative floats order like their bit patterns; ties -> lowest index
      const unsigned long long key =
          ((unsigned long long)__float_as_uint(r) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
      if (key < prev && key > best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = max(best, other);
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long b = 0ull;
      for (int w = 0; w < kTopkThreads / 32; ++w) b = max(b, sh[w]);
      s_prev = b;
      const long long id = (long long)(0xffffffffu - (unsigned)(b & 0xffffffffull));
      const size_t o = (size_t)row * n + t;
      idx[o] = id;
      if (xy != nullptr) {
        xy[2 * o + 0] = (float)(id % W);
        xy[2 * o + 1] = floorf(__fdiv_rn((float)id, (float)W));
      }
    }
    __syncthreads();
  }
}

// ---- Philox4x32-10 counter-based generators --------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// graph-safe streams: the effective seed mixes in a DEVICE-resident epoch counter, so that replaying a captured
// CUDA graph (same kernel arguments) still draws fresh numbers every step
__device__ __forceinline__ uint64_t mix_epoch(uint64_t seed, const uint64_t* epoch) {
  return epoch ? seed + (*epoch) * 0x9E3779B97F4A7C15ull : seed;
}
__global__ void counter_add_kernel(uint64_t* ctr, uint64_t inc) { *ctr += inc; }

__global__ void __launch_bounds__(256) rng_uniform_f64_kernel(uint64_t seed, const uint64_t* epoch, uint64_t offset, long long n,
                                                              double* __restrict__ out) {
  seed = mix_epoch(seed, epoch);
  const long long pairs = (n + 1) / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs;
       i += (long long)gridDim.x * blockDim.x) {
    const uint64_t c = offset + (uint64_t)i;
    const uint4 r = philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), 0x5954u, 0u),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const uint64_t a = ((uint64_t)r.x << 32) | r.y, b = ((uint64_t)r.z << 32) | r.w;
    out[2 * i] = (double)(a >> 11) * (1.0 / 9007199254740992.0);
    if (2 * i + 1 < n) out[2 * i + 1] = (double)(b >> 11) * (1.0 / 9007199254740992.0);
  }
}

__global__ void __launch_bounds__(256) rng_exponential_f32_kernel(uint64_t seed, const uint64_t* epoch, uint64_t offset, long long n,
                                                                  float* __restrict__ out) {
  seed = mix_epoch(seed, epoch);
  const long long quads = (n + 3) / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < quads;
       i += (long long)gridDim.x * blockDim.x) {
    const uint64_t c = offset + (uint64_t)i;
    const uint4 r = philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), 0x4558u, 0u),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const unsigned v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (4 * i + k < n) {
        const float u = ((float)(v[k] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
        out[4 * i + k] = -__logf(u);
      }
    }
  }
}

// K distinct indices in [0, N) per row (the device analogue of np.random.choice(N, K, replace=False))
__global__ void __launch_bounds__(128) rng_choice_kernel(uint64_t seed, const uint64_t* epoch, uint64_t offset, int rows, int N, int K,
                                                         int* __restrict__ out) {
  seed = mix_epoch(seed, epoch);
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  int* o = out + (size_t)row * K;
  uint64_t ctr = 0;
  for (int k = 0; k < K; ++k) {
    while (true) {
      const uint4 r = philox4x32_10(make_uint4((unsigned)ctr, (unsigned)(ctr >> 32), 0x4b4du, (unsigned)row),
                                    make_uint2((unsigned)(seed + offset), (unsigned)((seed + offset) >> 32)));
      ++ctr;
      const int cand = (int)(((uint64_t)r.x * (uint64_t)N) >> 32);
      bool dup = false;
      for (int j = 0; j < k; ++j) dup |= (o[j] == cand);
      if (!dup) {
        o[k] = cand;
        break;
      }
    }
  }
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int64_t ynet_sampling_prepare_workspace_bytes(int32_t rows, int64_t S) {
  return (int64_t)rows * kSumChunks * (int64_t)sizeof(double);
}

int ynet_sampling_prepare(const float* prob, int32_t rows, int64_t S, float rel_threshold, float* rowmax, float* gsum,
                          void* workspace, int64_t workspace_bytes, void* stream) {
  YNET_CHECK_ARG(prob && rowmax && gsum, "null pointer");
  YNET_CHECK_ARG(rows > 0 && rows <= 65535 && S > 0, "bad shape (rows in [1, 65535])");
  if (workspace == nullptr || workspace_bytes < ynet_sampling_prepare_workspace_bytes(rows, S)) {
    set_error("ynet_sampling_prepare: workspace too small");
    return YNET_E_WORKSPACE;
  }
  double* partial = reinterpret_cast<double*>(workspace);
  rowmax_kernel<<<rows, 1024, 0, as_stream(stream)>>>(prob, S, rowmax);
  YNET_LAUNCH_CHECK();
  masked_sum_partial_kernel<<<dim3(kSumChunks, rows), 256, 0, as_stream(stream)>>>(prob, S, rel_threshold, rowmax,
                                                                                   partial);
  YNET_LAUNCH_CHECK();
  masked_sum_final_kernel<<<1, 256, 0, as_stream(stream)>>>(partial, (long long)rows * kSumChunks, gsum);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_multinomial_replacement(const float* prob, int32_t rows, int64_t S, float rel_threshold, const float* rowmax,
                                 const float* gsum, const double* uniforms, int32_t n, float* cdf_ws, int64_t* idx,
                                 float* xy, int32_t W, void* stream) {
  YNET_CHECK_ARG(prob && uniforms && cdf_ws && idx, "null pointer");
  YNET_CHECK_ARG(rows > 0 && rows <= 65535 && S > 0 && n > 0 && W > 0, "bad shape");
  const bool use_thr = rel_threshold >= 0.f;
  YNET_CHECK_ARG(!use_thr || (rowmax && gsum), "threshold requested without rowmax/gsum (ynet_sampling_prepare)");
  cdf_sequential_kernel<<<rows, kCdfThreads, 0, as_stream(stream)>>>(
      prob, rows, S, use_thr, rel_threshold, rowmax, gsum, cdf_ws);
  YNET_LAUNCH_CHECK();
  cdf_search_kernel<<<dim3(ceil_div(n, 256), rows), 256, 0, as_stream(stream)>>>(
      cdf_ws, S, uniforms, n, reinterpret_cast<long long*>(idx), xy, W);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_multinomial_topk(const float* prob, const float* expo, int32_t rows, int64_t S, float rel_threshold,
                          const float* rowmax, const float* gsum, int32_t n, int64_t* idx, float* xy, int32_t W,
                          void* stream) {
  YNET_CHECK_ARG(prob && expo && idx, "null pointer");
  YNET_CHECK_ARG(rows > 0 && S > 0 && S < 0xffffffffLL && n > 0 && n <= S && W > 0, "bad shape");
  const bool use_thr = rel_threshold >= 0.f;
  YNET_CHECK_ARG(!use_thr || (rowmax && gsum), "threshold requested without rowmax/gsum (ynet_sampling_prepare)");
  topk_ratio_kernel<<<rows, kTopkThreads, 0, as_stream(stream)>>>(prob, expo, S, use_thr, rel_threshold, rowmax, gsum,
                                                                  n, reinterpret_cast<long long*>(idx), xy, W);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_counter_add(uint64_t* counter, uint64_t inc, void* stream) {
  YNET_CHECK_ARG(counter, "null pointer");
  counter_add_kernel<<<1, 1, 0, as_stream(stream)>>>(counter, inc);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_rng_uniform_f64(uint64_t seed, const uint64_t* epoch, uint64_t offset, int64_t n, double* out, void* stream) {
  YNET_CHECK_ARG(out && n >= 0, "bad argument");
  if (n == 0) return YNET_OK;
  const int blocks = (int)tmin<long long>(ceil_div<long long>((n + 1) / 2, 256), 8LL * sm_count());
  rng_uniform_f64_kernel<<<blocks, 256, 0, as_stream(stream)>>>(seed, epoch, offset, n, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_rng_choice(uint64_t seed, const uint64_t* epoch, uint64_t offset, int32_t rows, int32_t N, int32_t K, int32_t* out,
                    void* stream) {
  YNET_CHECK_ARG(out && rows >= 0 && N > 0 && K > 0 && K <= N, "bad argument");
  if (rows == 0) return YNET_OK;
  rng_choice_kernel<<<ceil_div(rows, 128), 128, 0, as_stream(stream)>>>(seed, epoch, offset, rows, N, K, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_rng_exponential_f32(uint64_t seed, const uint64_t* epoch, uint64_t offset, int64_t n, float* out, void* stream) {
  YNET_CHECK_ARG(out && n >= 0, "bad argument");
  if (n == 0) return YNET_OK;
  const int blocks = (int)tmin<long long>(ceil_div<long long>((n + 3) / 4, 256), 8LL * sm_count());
  rng_exponential_f32_kernel<<<blocks, 256, 0, as_stream(stream)>>>(seed, epoch, offset, n, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

}  // extern "C"
