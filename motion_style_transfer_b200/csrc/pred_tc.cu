// a8 + a12: trajectory-decoder predictor (1x1 conv, ynet.py:450-451,469) fused with SoftArgmax2D
// (softargmax.py:55-81, ynet.py:582-583) on the tensor cores -- the logits never reach HBM.
//
// Operand roles are SWAPPED with respect to conv_tc.cu: the predictor weights are the M operand and the pixels the
// N operand, so the accumulator comes out TRANSPOSED: TMEM lane = output channel, TMEM column = pixel.  A thread
// that owns one lane therefore holds a run of pixels of ONE channel in registers after tcgen05.ld, which is exactly
// the shape a per-channel spatial soft-max wants: no shared-memory transposition, no cross-thread reduction, no
// CTA barrier in the epilogue.  The <= 32 weight rows are replicated into the four 32-lane quadrants of the M = 128
// tile so that all four TMEM lane quadrants (i.e. all four SM sub-partitions) carry epilogue work; each of the 16
// epilogue warps reduces one 16-pixel row of every 16 x 16-pixel tile (accumulator column = row * 16 + x).
//
// Per 128 pixels: 2 MMAs (K = 32) of 64 cycles, 8 KB of HBM (356 cycles at the measured copy bandwidth), 4096
// exponentials (256 cycles of MUFU): the kernel is HBM-bound by design.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace ynet {

constexpr int PR_THREADS = 64 + 32 * PR_EPI_WARPS;
constexpr int PR_STAGES = 14;
constexpr int PR_STAGE_BYTES = 2 * PR_TH * PR_TW * 16;   // one 16-channel K block of a tile: [chunk][row][px][8 ch]

struct PredParams {
  int N, H, W, c_out, n_pad, kblocks;
  int tiles_x, tiles_y;
  long long total_tiles, tiles_per_cta;
  int slots;                      // partial slots per (image, channel)
  const unsigned char* wpacked;   // [kb][1 tap][2][n_pad][8] bf16 (ynet_tc_pack_weights, ksize 1)
  const float* bias;              // n_pad floats
  float4* partial;                // (m, s, sx, sy) [(n * c_out + c) * slots + slot]
};

__global__ void __launch_bounds__(PR_THREADS, 1)
tc_pred_softargmax_kernel(const __grid_constant__ CUtensorMap map, const PredParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  unsigned char* s_w = smem;                                                   // [kb][2][128][8] bf16
  unsigned char* s_stage = smem + PR_MAX_KB * PR_WBLK_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stage + (size_t)PR_STAGES * PR_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + PR_STAGES;
  uint64_t* tfull_bar = empty_bar + PR_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  pred_stage_weights(s_w, p.wpacked, p.kblocks, p.n_pad, threadIdx.x, PR_THREADS);
  if (threadIdx.x == 0) {
    for (int s = 0; s < PR_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), PR_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const long long t0 = (long long)blockIdx.x * p.tiles_per_cta;
  const long long t1 = tmin<long long>(p.total_tiles, t0 + p.tiles_per_cta);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int n = (int)(t0 / tiles_per_img);
      const int r0t = (int)(t0 - (long long)n * tiles_per_img);
      int ty = r0t / p.tiles_x, tx = r0t - ty * p.tiles_x;
      for (long long tile = t0; tile < t1; ++tile) {
        const int y0 = ty * PR_TH, x0 = tx * PR_TW;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, nullptr);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, PR_STAGE_BYTES);
          tma_load_4d(smem_u32(s_stage + (size_t)stage * PR_STAGE_BYTES), &map, fb, 8 * x0, y0, 2 * kb, n);
          if (++stage == PR_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++tx == p.tiles_x) {
          tx = 0;
          if (++ty == p.tiles_y) {
            ty = 0;
            ++n;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {   // whole warp, uniform control flow; the MMA / commit instructions elect one lane
      // D = F32, A = B = BF16, both K-major, N = 256 pixels (the whole 16 x 16 tile), M = 128 (4 x 32 replicated channels)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long tile = t0; tile < t1; ++tile, ++it) {
        const int acc = (int)(it & 1);
        mbar_wait(smem_u32(&tempty_bar[acc]), (uint32_t)((it >> 1) & 1) ^ 1, nullptr);
        tc_fence_after();
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase, nullptr);
          tc_fence_after();
          const uint32_t st = smem_u32(s_stage + (size_t)stage * PR_STAGE_BYTES);
          const uint64_t adesc = make_desc(smem_u32(s_w + (size_t)kb * PR_WBLK_BYTES), 128 * 16, 128);
          // a tile row is 16 px = two contiguous 128-byte core matrices: all 32 eight-pixel groups of the tile are
          // 128 B apart, so ONE N = 256 instruction covers the tile; accumulator column = row * 16 + x
          const uint64_t bdesc = make_desc(st, PR_TH * PR_TW * 16, 128);
          tc_mma_bf16_elect(tmem_base + (uint32_t)(acc * 256), adesc, bdesc, idesc, kb > 0 ? 1u : 0u);
          tc_commit_elect(smem_u32(&empty_bar[stage]));
          if (++stage == PR_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit_elect(smem_u32(&tfull_bar[acc]));
      }
    }
  } else {
    // ===================== epilogue: 16 warps; warp w may touch TMEM lanes 32 (w % 4) .. +31 =====================
    const int e = warp - 2;
    const int q = warp & 3;
    const int row = q * 4 + (e >> 2);              // tile row reduced by this warp (4 warps per lane quadrant)
    const bool active = lane < p.c_out;
    const float bias = active ? p.bias[lane] : 0.f;
    SoftState st{PR_NEG, 0.f, 0.f, 0.f};
    int cur_n = -1;
    auto flush = [&](int n_img) {
      if (active) {
        const long long first_cta = ((long long)n_img * tiles_per_img) / p.tiles_per_cta;
        const int slot = (int)(blockIdx.x - first_cta) * PR_EPI_WARPS + e;
        p.partial[((size_t)n_img * p.c_out + lane) * p.slots + slot] = make_float4(st.m, st.s, st.sx, st.sy);
      }
    };
    // tile coordinates advance incrementally (no per-tile 64-bit divisions in the hot loop)
    int n = (int)(t0 / tiles_per_img);
    int r0t = (int)(t0 - (long long)n * tiles_per_img);
    int ty = r0t / p.tiles_x, tx = r0t - ty * p.tiles_x;
    uint32_t it = 0;
    for (long long tile = t0; tile < t1; ++tile, ++it) {
      const int y0 = ty * PR_TH, x0 = tx * PR_TW;
      const int acc = (int)(it & 1u);
      if (n != cur_n) {
        if (cur_n >= 0) flush(cur_n);
        st = SoftState{PR_NEG, 0.f, 0.f, 0.f};
        cur_n = n;
      }
      mbar_wait(smem_u32(&tfull_bar[acc]), (it >> 1) & 1u, nullptr);
      tc_fence_after();
      const int y = y0 + row;
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + row * 16), v);
      if (y < p.H) {
        if (x0 + PR_TW <= p.W)
          softargmax_row16(st, v, bias, x0, y);
        else
          softargmax_row16_masked(st, v, bias, x0, y, p.W);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++tx == p.tiles_x) {
        tx = 0;
        if (++ty == p.tiles_y) {
          ty = 0;
          ++n;
        }
      }
    }
    if (cur_n >= 0) flush(cur_n);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

__global__ void __launch_bounds__(256) pred_partial_init_kernel(float4* __restrict__ part, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    part[i] = make_float4(PR_NEG, 0.f, 0.f, 0.f);
}

// one warp per (image, channel): combine the partials, apply 1/(sum + 1e-6)  (softargmax.py:68)
__global__ void __launch_bounds__(256)
pred_partial_finalize_kernel(const float4* __restrict__ part, int rows, int slots, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float m = PR_NEG, sm = 0.f, sx = 0.f, sy = 0.f;
  for (int i = lane; i < slots; i += 32) {
    const float4 v = part[(size_t)row * slots + i];
    const float M = fmaxf(m, v.x);
    const float fa = (m == PR_NEG) ? 0.f : __expf(m - M);
    const float fb = (v.x == PR_NEG) ? 0.f : __expf(v.x - M);
    sm = sm * fa + v.y * fb;
    sx = sx * fa + v.z * fb;
    sy = sy * fa + v.w * fb;
    m = M;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, sm, o);
    const float x2 = __shfl_xor_sync(0xffffffffu, sx, o), y2 = __shfl_xor_sync(0xffffffffu, sy, o);
    const float M = fmaxf(m, m2);
    const float fa = (m == PR_NEG) ? 0.f : __expf(m - M);
    const float fb = (m2 == PR_NEG) ? 0.f : __expf(m2 - M);
    sm = sm * fa + s2 * fb;
    sx = sx * fa + x2 * fb;
    sy = sy * fa + y2 * fb;
    m = M;
  }
  if (lane == 0) {
    const float inv = 1.0f / (sm + 1e-6f);
    out[2 * row + 0] = sx * inv;
    out[2 * row + 1] = sy * inv;
  }
}

PredPlan pred_plan(int N, int H, int W) {
  PredPlan pl;
  pl.tiles_x = ceil_div(W, PR_TW);
  pl.tiles_y = ceil_div(H, PR_TH);
  const long long per_img = (long long)pl.tiles_x * pl.tiles_y;
  pl.total_tiles = (long long)N * per_img;
  const long long g0 = tmax<long long>(1, tmin<long long>(pl.total_tiles, sm_count()));
  pl.tiles_per_cta = tmax<long long>(1, ceil_div<long long>(pl.total_tiles, g0));
  pl.grid = (int)tmax<long long>(1, ceil_div<long long>(pl.total_tiles, pl.tiles_per_cta));
  // CTAs own contiguous tile ranges: an image is touched by at most ceil(per_img / tiles_per_cta) + 1 of them
  pl.slots = (int)(ceil_div<long long>(per_img, pl.tiles_per_cta) + 1) * PR_EPI_WARPS;
  return pl;
}

cudaError_t pred_partial_init(float4* part, long long n_part, cudaStream_t st) {
  const unsigned ig = (unsigned)tmax<long long>(1, tmin<long long>(ceil_div<long long>(n_part, 256), 8LL * sm_count()));
  pred_partial_init_kernel<<<ig, 256, 0, st>>>(part, n_part);
  return cudaGetLastError();
}
cudaError_t pred_partial_finalize(const float4* part, int rows, int slots, float* out, cudaStream_t st) {
  pred_partial_finalize_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(part, rows, slots, out);
  return cudaGetLastError();
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int64_t ynet_tc_conv1x1_softargmax_workspace_bytes(int32_t N, int32_t C_out, int32_t H, int32_t W) {
  if (N <= 0 || C_out <= 0 || H <= 0 || W <= 0) return 0;
  const PredPlan pl = pred_plan(N, H, W);
  return (int64_t)N * C_out * pl.slots * (int64_t)sizeof(float4);
}

int ynet_tc_conv1x1_softargmax(const ynet_tc_src* srcs, int32_t n_src, int32_t N, int32_t H, int32_t W,
                               const void* packed_weight, const float* bias, int32_t C_out, float* out, void* workspace,
                               int64_t workspace_bytes, int32_t tune, void* stream) {
  (void)tune;
  YNET_CHECK_ARG(out != nullptr || N == 0, "null output");
  YNET_CHECK_ARG(srcs && packed_weight && bias, "null pointer");
  YNET_CHECK_ARG(n_src == 1, "the fused predictor takes exactly one source");
  YNET_CHECK_ARG(C_out > 0 && C_out <= 32, "C_out must be <= 32 for the fused soft-argmax epilogue");
  YNET_CHECK_ARG(N >= 0 && H > 0 && W > 0, "bad shape");
  if (N == 0) return YNET_OK;
  const int cp = srcs[0].channels_pad;
  YNET_CHECK_ARG(srcs[0].ptr && cp > 0 && cp % 16 == 0 && cp / 16 <= PR_MAX_KB, "channels_pad must be 16..128, multiple of 16");
  YNET_CHECK_ARG(srcs[0].batch_stride != 0 && srcs[0].batch_mod == 0, "broadcast / repeated sources are not supported here");
  YNET_CHECK_ALIGN(srcs[0].ptr, 16);
  YNET_CHECK_ALIGN(packed_weight, 16);
  if (workspace == nullptr || workspace_bytes < ynet_tc_conv1x1_softargmax_workspace_bytes(N, C_out, H, W)) {
    set_error("ynet_tc_conv1x1_softargmax: workspace too small");
    return YNET_E_WORKSPACE;
  }
  YNET_CHECK_ALIGN(workspace, 16);
  EncodeTiledFn encode = tc_get_encode();
  if (encode == nullptr) {
    set_error("ynet_tc_conv1x1_softargmax: cuTensorMapEncodeTiled is not available from the driver");
    return YNET_E_UNSUPPORTED;
  }
  const PredPlan pl = pred_plan(N, H, W);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)(cp / 8), (cuuint64_t)N};
  const cuuint64_t bs = (cuuint64_t)srcs[0].batch_stride * 2;
  YNET_CHECK_ARG(bs % 16 == 0, "batch stride must be a multiple of 8 elements");
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, bs};
  const cuuint32_t box[4] = {(cuuint32_t)PR_TW * 8, (cuuint32_t)PR_TH, 2, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(srcs[0].ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("ynet_tc_conv1x1_softargmax: cuTensorMapEncodeTiled failed (%d) (W=%d H=%d C=%d)", (int)r, W, H, cp);
    return YNET_E_CUDA;
  }
  PredParams p;
  memset(&p, 0, sizeof(p));
  p.N = N;
  p.H = H;
  p.W = W;
  p.c_out = C_out;
  p.n_pad = ceil_div(C_out, 16) * 16;
  p.kblocks = cp / 16;
  p.tiles_x = pl.tiles_x;
  p.tiles_y = pl.tiles_y;
  p.total_tiles = pl.total_tiles;
  p.tiles_per_cta = pl.tiles_per_cta;
  p.slots = pl.slots;
  p.wpacked = reinterpret_cast<const unsigned char*>(packed_weight);
  p.bias = bias;
  p.partial = reinterpret_cast<float4*>(workspace);

  const size_t smem_bytes = (size_t)PR_MAX_KB * PR_WBLK_BYTES + (size_t)PR_STAGES * PR_STAGE_BYTES +
                            (2 * PR_STAGES + 4) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_pred_softargmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "ynet_tc_conv1x1_softargmax(cudaFuncSetAttribute)");
    configured = true;
  }
  cudaStream_t st = as_stream(stream);
  const long long n_part = (long long)N * C_out * pl.slots;
  cudaError_t le = pred_partial_init(p.partial, n_part, st);
  if (le != cudaSuccess) return cuda_fail(le, "ynet_tc_conv1x1_softargmax");
  tc_pred_softargmax_kernel<<<pl.grid, PR_THREADS, smem_bytes, st>>>(map, p);
  YNET_LAUNCH_CHECK();
  le = pred_partial_finalize(p.partial, N * C_out, pl.slots, out, st);
  if (le != cudaSuccess) return cuda_fail(le, "ynet_tc_conv1x1_softargmax");
  return YNET_OK;
}

}  // extern "C"
