// a4-a8: tensor-core network engine for sm_100a -- tcgen05.mma (kind::f16, bf16 operands, fp32
// accumulators in TMEM) fed by TMA, warp-specialised, persistent.
//
// Layout.  Activations live in HBM as bf16 "C8" planes  [N][C/8][H][W][8]  (C padded to 16).  A TMA box
// {8 ch, 10 px, 18 rows, 2 chunks} of that tensor lands in shared memory as [chunk][row][px][8 ch]: every
// run of 8 consecutive pixels x 8 channels is one 128-byte UMMA core matrix (no-swizzle, K-major), rows of
// the image tile are 160 B apart (SBO) and the two 8-channel chunks of a K=16 step are 2880 B apart (LBO).
// The 3x3 taps are therefore NINE DESCRIPTOR OFFSETS into ONE halo tile: start += (kh*10 + kw)*16 bytes.
// The tile is loaded once (1.4x halo, zero-filled out of bounds by TMA = the conv's padding) instead of the
// 9x shared-memory refill of an im2col pipeline -- with C_out = 32 that refill would be the bottleneck.
//
// One CTA = one 16x8-pixel output tile (M = 128) x all output channels (N = C_out padded to 16, <= 256)
// per iteration, persistent over tiles.  warp 0: TMA producer; warp 1: MMA issuer (one elected lane)
// + TMEM allocator; warps 2-5: epilogue (tcgen05.ld -> bias -> ReLU -> bf16 -> C8 store).  Accumulators
// are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ynet {

constexpr int TC_TH = 16, TC_TW = 8;                 // output tile (pixels)
constexpr int TC_BH = TC_TH + 2, TC_BW = TC_TW + 2;  // halo box
constexpr int TC_KB = 16;                            // channels per pipeline stage (one UMMA K step)
constexpr int TC_A_BYTES = TC_BH * TC_BW * TC_KB * 2;            // 5760
constexpr int TC_A_LBO = TC_BH * TC_BW * 16;                     // 2880: chunk (8 ch) stride
constexpr int TC_A_SBO = TC_BW * 16;                             // 160: image-row stride = 8-pixel-group stride
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_STAGES = 8;
constexpr unsigned TC_SPIN_LIMIT = 4u * 1000u * 1000u;             // bounded waits: trap instead of hanging

struct TcSrcDev {
  int kblocks;      // channels_pad / 16
  int bcast;        // source has batch 1
  int batch_mod;    // > 0: image n reads n % batch_mod
  int pad_;
};

struct TcParams {
  TcSrcDev src[YNET_MAX_SOURCES];
  int n_src, N, H, W, n_pad /* C_out padded */, relu;
  int tiles_x, tiles_y;
  long long total_tiles;
  int kb_total;           // sum of kblocks
  int resident;           // weights resident in smem
  int stages, stage_bytes, wres_bytes, tmem_cols;
  const unsigned char* wpacked;   // [kb][tap][2][n_pad][8] bf16
  const float* bias;              // n_pad floats (pad = 0)
  __nv_bfloat16* out;             // C8 planes, n_pad channels
  int* err;
};

// ---- PTX wrappers --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
  uint32_t done = 0;
  unsigned spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > TC_SPIN_LIMIT) {  // a protocol bug must not hang the GPU
      if (err) atomicExch(err, 1);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- the conv kernel -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv3x3_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                  const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3,
                  const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // TMA destinations need 128 B alignment; align the carve-up base to 1 KB explicitly
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  unsigned char* s_w = smem;                                       // resident weights (may be empty)
  unsigned char* s_stage = smem + p.wres_bytes;                    // stages
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + (size_t)p.stages * p.stage_bytes);
  uint64_t* full_bar = s_bar;
  uint64_t* empty_bar = s_bar + TC_MAX_STAGES;
  uint64_t* tfull_bar = s_bar + 2 * TC_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(w_bar + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 4);  // one arrive per epilogue warp
    }
    mbar_init(smem_u32(w_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int wblk_bytes = 9 * 2 * p.n_pad * 16;  // weights of one 16-channel K block: [tap][2][n_pad][8] bf16

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (p.resident) {
        const uint32_t total = (uint32_t)p.kb_total * wblk_bytes;
        mbar_expect_tx(smem_u32(w_bar), total);
        for (uint32_t off = 0; off < total; off += 32768) {
          const uint32_t n = min(32768u, total - off);
          bulk_load(smem_u32(s_w + off), p.wpacked + off, n, smem_u32(w_bar));
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n = (int)(tile / tiles_per_img);
        const int r = (int)(tile - (long long)n * tiles_per_img);
        const int y0 = (r / p.tiles_x) * TC_TH, x0 = (r % p.tiles_x) * TC_TW;
        int kb = 0;
        for (int s = 0; s < p.n_src; ++s) {
          const CUtensorMap* map = (s == 0) ? &map0 : (s == 1) ? &map1 : (s == 2) ? &map2 : &map3;
          const int ns = p.src[s].bcast ? 0 : (p.src[s].batch_mod > 0 ? n % p.src[s].batch_mod : n);
          for (int b = 0; b < p.src[s].kblocks; ++b, ++kb) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.err);
            unsigned char* st = s_stage + (size_t)stage * p.stage_bytes;
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_expect_tx(fb, TC_A_BYTES + (p.resident ? 0 : wblk_bytes));
            tma_load_5d(smem_u32(st), map, fb, 0, x0 - 1, y0 - 1, 2 * b, ns);
            if (!p.resident) bulk_load(smem_u32(st + TC_A_BYTES), p.wpacked + (size_t)kb * wblk_bytes, wblk_bytes, fb);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = BF16, K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((128u >> 4) << 24);
      if (p.resident) mbar_wait(smem_u32(w_bar), 0, p.err);
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = (int)(it & 1);
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1, p.err);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_pad);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase, p.err);
          tc_fence_after();
          unsigned char* st = s_stage + (size_t)stage * p.stage_bytes;
          const uint32_t a_base = smem_u32(st);
          const uint32_t b_base = p.resident ? smem_u32(s_w + (size_t)kb * wblk_bytes) : smem_u32(st + TC_A_BYTES);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int kh = tap / 3, kw = tap - kh * 3;
            const uint64_t adesc = make_desc(a_base + (uint32_t)(kh * TC_BW + kw) * 16, TC_A_LBO, TC_A_SBO);
            const uint64_t bdesc = make_desc(b_base + (uint32_t)tap * 2 * p.n_pad * 16, (uint32_t)p.n_pad * 16, 128);
            tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb > 0 || tap > 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(&empty_bar[stage]));  // frees the smem slot when these MMAs retire
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(smem_u32(&tfull_bar[acc]));      // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;         // accumulator row = pixel of the tile
    const int py = m >> 3, px = m & 7;
    const int n_chunks = p.n_pad >> 3;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int n = (int)(tile / tiles_per_img);
      const int r = (int)(tile - (long long)n * tiles_per_img);
      const int y = (r / p.tiles_x) * TC_TH + py, x = (r % p.tiles_x) * TC_TW + px;
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, p.err);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.n_pad);
      const bool inb = (y < p.H) && (x < p.W);
      for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_row + (uint32_t)c0, v);
        float f[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          f[k] = __uint_as_float(v[k]) + __ldg(p.bias + c0 + k);
          if (p.relu) f[k] = fmaxf(f[k], 0.f);
        }
        if (inb) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int chunk = (c0 >> 3) + h;
            uint4 o;
            o.x = pack_bf16(f[8 * h + 0], f[8 * h + 1]);
            o.y = pack_bf16(f[8 * h + 2], f[8 * h + 3]);
            o.z = pack_bf16(f[8 * h + 4], f[8 * h + 5]);
            o.w = pack_bf16(f[8 * h + 6], f[8 * h + 7]);
            __nv_bfloat16* dst = p.out + ((((size_t)n * n_chunks + chunk) * p.H + y) * p.W + x) * 8;
            *reinterpret_cast<uint4*>(dst) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ---- layout conversion and the bandwidth-bound companions of the C8 engine ------------------------------------
// NCHW f32 -> C8 bf16 (channels >= C are zero).  One thread = one pixel x one 8-channel chunk (16 B store).
__global__ void __launch_bounds__(256)
pack_c8_kernel(const float* __restrict__ x, int C, int H, int W, long long batch_stride, __nv_bfloat16* __restrict__ out,
               int chunks, long long total) {
  const long long S = (long long)H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pix = t % S;
    const long long rest = t / S;
    const int chunk = (int)(rest % chunks);
    const long long n = rest / chunks;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = chunk * 8 + k;
      f[k] = (c < C) ? __ldg(x + n * batch_stride + (long long)c * S + pix) : 0.f;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]);
    o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]);
    o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + t * 8) = o;
  }
}

__global__ void __launch_bounds__(256)
unpack_c8_kernel(const __nv_bfloat16* __restrict__ x, int C, int chunks, int H, int W, float* __restrict__ out,
                 long long total) {
  const long long S = (long long)H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pix = t % S;
    const long long rest = t / S;
    const int c = (int)(rest % C);
    const long long n = rest / C;
    out[t] = __bfloat162float(x[(((n * chunks + (c >> 3)) * S) + pix) * 8 + (c & 7)]);
  }
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]);
  o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]);
  o.w = pack_bf16(f[6], f[7]);
  return o;
}

// 2x2 max-pool on C8 planes: planes = N * chunks; (H, W) input size
__global__ void __launch_bounds__(256)
c8_maxpool_kernel(const uint4* __restrict__ x, long long planes, int H, int W, uint4* __restrict__ out) {
  const int h = H >> 1, w = W >> 1;
  const long long total = planes * h * w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / (h * w);
    const int r = (int)(t - pl * h * w);
    const int y = r / w, xx = r - y * w;
    const uint4* q = x + (pl * H + 2 * y) * W + 2 * xx;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8(__ldg(q), a);
    unpack8(__ldg(q + 1), b);
    unpack8(__ldg(q + W), c);
    unpack8(__ldg(q + W + 1), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
    out[t] = pack8(o);
  }
}

// bilinear x2 (align_corners=False) on C8 planes: (H, W) input size
__global__ void __launch_bounds__(256)
c8_upsample_kernel(const uint4* __restrict__ x, long long planes, int H, int W, uint4* __restrict__ out) {
  const int OH = 2 * H, OW = 2 * W;
  const long long total = planes * OH * OW;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / ((long long)OH * OW);
    const int r = (int)(t - pl * OH * OW);
    const int y = r / OW, xx = r - y * OW;
    const float fy = fmaxf(0.f, ((float)y + 0.5f) * 0.5f - 0.5f);
    const float fx = fmaxf(0.f, ((float)xx + 0.5f) * 0.5f - 0.5f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const uint4* q = x + pl * H * W;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8(__ldg(q + (size_t)y0 * W + x0), a);
    unpack8(__ldg(q + (size_t)y0 * W + x1), b);
    unpack8(__ldg(q + (size_t)y1 * W + x0), c);
    unpack8(__ldg(q + (size_t)y1 * W + x1), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (1.f - ly) * ((1.f - lx) * a[k] + lx * b[k]) + ly * ((1.f - lx) * c[k] + lx * d[k]);
    out[t] = pack8(o);
  }
}

// weights OIHW f32 -> [kb][tap][2][n_pad][8] bf16 over the concatenated, per-source padded input channels
struct PackSrc {
  int real[YNET_MAX_SOURCES], pad[YNET_MAX_SOURCES];
  int n_src;
};
__global__ void __launch_bounds__(256)
tc_pack_weights_kernel(const float* __restrict__ w, int C_out, int C_in, int n_pad, PackSrc ps, int kb_total,
                       __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)kb_total * 9 * 2 * n_pad * 8;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int k8 = (int)(t & 7);
    long long r = t >> 3;
    const int n = (int)(r % n_pad);
    r /= n_pad;
    const int c = (int)(r & 1);
    r >>= 1;
    const int tap = (int)(r % 9);
    const int kb = (int)(r / 9);
    int kpad = kb * 16 + c * 8 + k8;     // index in the padded concatenation
    int ci = -1, off_real = 0;
    for (int s = 0; s < ps.n_src; ++s) {
      if (kpad < ps.pad[s]) {
        if (kpad < ps.real[s]) ci = off_real + kpad;
        break;
      }
      kpad -= ps.pad[s];
      off_real += ps.real[s];
    }
    float v = 0.f;
    if (ci >= 0 && n < C_out) v = w[((size_t)n * C_in + ci) * 9 + tap];
    out[t] = __float2bfloat16_rn(v);
  }
}

// 1x1 predictor on C8 input -> NCHW f32 logits
constexpr int TCP_MAXC = 32;
__global__ void __launch_bounds__(256)
c8_predictor_kernel(const uint4* __restrict__ x, int chunks, int C_in, long long S, const float* __restrict__ weight,
                    const float* __restrict__ bias, int C_out, float* __restrict__ out) {
  __shared__ float s_w[TCP_MAXC][TCP_MAXC];  // [ci][co]
  __shared__ float s_b[TCP_MAXC];
  for (int e = threadIdx.x; e < TCP_MAXC * TCP_MAXC; e += blockDim.x) {
    const int co = e / TCP_MAXC, ci = e - co * TCP_MAXC;
    s_w[ci][co] = (co < C_out && ci < C_in) ? weight[co * C_in + ci] : 0.f;
  }
  if (threadIdx.x < TCP_MAXC) s_b[threadIdx.x] = (bias && threadIdx.x < C_out) ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const uint4* xn = x + (size_t)n * chunks * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) {
    float acc[TCP_MAXC];
#pragma unroll
    for (int b = 0; b < TCP_MAXC; ++b) acc[b] = s_b[b];
    for (int ch = 0; ch * 8 < C_in; ++ch) {
      float f[8];
      unpack8(__ldg(xn + (size_t)ch * S + i), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ci = ch * 8 + k;
#pragma unroll
        for (int b = 0; b < TCP_MAXC; ++b) acc[b] = fmaf(f[k], s_w[ci][b], acc[b]);
      }
    }
    for (int b = 0; b < C_out; ++b) out[((size_t)n * C_out + b) * S + i] = acc[b];
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static inline unsigned grid_1d(long long n) {
  return (unsigned)tmax<long long>(1, tmin<long long>(ceil_div<long long>(n, 256), 16LL * sm_count()));
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int ynet_tc_supported(void) {
  int dev = 0, ma = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev);
  return (ma == 10 && get_encode() != nullptr) ? 1 : 0;
}

int ynet_tc_pack_f32_to_c8(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, int64_t batch_stride,
                           void* out_c8, int32_t C_pad, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && C_pad >= C && C_pad % 16 == 0, "bad shape (C_pad % 16 == 0)");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x && out_c8, "null pointer");
  YNET_CHECK_ALIGN(out_c8, 16);
  const long long total = (long long)N * (C_pad / 8) * H * W;
  pack_c8_kernel<<<grid_1d(total), 256, 0, as_stream(stream)>>>(x, C, H, W, batch_stride,
                                                                reinterpret_cast<__nv_bfloat16*>(out_c8), C_pad / 8, total);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_unpack_c8_to_f32(const void* x_c8, int32_t N, int32_t C, int32_t C_pad, int32_t H, int32_t W, float* out,
                             void* stream) {
  YNET_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && C_pad >= C && C_pad % 8 == 0, "bad shape");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x_c8 && out, "null pointer");
  const long long total = (long long)N * C * H * W;
  unpack_c8_kernel<<<grid_1d(total), 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x_c8), C,
                                                                  C_pad / 8, H, W, out, total);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_maxpool2x2(const void* x_c8, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out_c8, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C_pad % 8 == 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "bad shape");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x_c8 && out_c8, "null pointer");
  const long long planes = (long long)N * (C_pad / 8);
  c8_maxpool_kernel<<<grid_1d(planes * (H / 2) * (W / 2)), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x_c8), planes, H, W, reinterpret_cast<uint4*>(out_c8));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_upsample2x(const void* x_c8, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out_c8, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C_pad % 8 == 0 && H >= 1 && W >= 1, "bad shape");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x_c8 && out_c8, "null pointer");
  const long long planes = (long long)N * (C_pad / 8);
  c8_upsample_kernel<<<grid_1d(planes * 4LL * H * W), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x_c8), planes, H, W, reinterpret_cast<uint4*>(out_c8));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_predictor_f32(const void* x_c8, int32_t N, int32_t C_pad, int32_t C_in, int32_t H, int32_t W,
                          const float* weight, const float* bias, int32_t C_out, float* out, void* stream) {
  YNET_CHECK_ARG(N >= 0 && N <= 65535 && C_pad % 8 == 0 && C_in > 0 && C_in <= C_pad && C_in <= TCP_MAXC && C_out > 0 &&
                     C_out <= TCP_MAXC && H > 0 && W > 0,
                 "bad shape (C_in, C_out <= 32)");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x_c8 && weight && out, "null pointer");
  const long long S = (long long)H * W;
  dim3 grid((unsigned)tmin<long long>(ceil_div<long long>(S, 256), 1024), N);
  c8_predictor_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(x_c8), C_pad / 8, C_in, S,
                                                           weight, bias, C_out, out);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int64_t ynet_tc_packed_weight_bytes(int32_t C_out, int32_t n_src, const int32_t* src_channels_pad_host) {
  if (C_out <= 0 || n_src <= 0 || n_src > YNET_MAX_SOURCES || !src_channels_pad_host) return 0;
  const int n_pad = ceil_div(C_out, 16) * 16;
  long long kb = 0;
  for (int i = 0; i < n_src; ++i) kb += src_channels_pad_host[i] / 16;
  return kb * 9 * 2 * n_pad * 16;
}

int ynet_tc_pack_weights(const float* weight, int32_t C_out, int32_t n_src, const int32_t* src_channels_host,
                         const int32_t* src_channels_pad_host, void* packed, void* stream) {
  YNET_CHECK_ARG(weight && packed && src_channels_host && src_channels_pad_host, "null pointer");
  YNET_CHECK_ARG(C_out > 0 && n_src >= 1 && n_src <= YNET_MAX_SOURCES, "bad shape");
  PackSrc ps;
  memset(&ps, 0, sizeof(ps));
  ps.n_src = n_src;
  int cin = 0, kb = 0;
  for (int i = 0; i < n_src; ++i) {
    YNET_CHECK_ARG(src_channels_pad_host[i] % 16 == 0 && src_channels_pad_host[i] >= src_channels_host[i] &&
                       src_channels_host[i] > 0,
                   "source channels must be padded to a multiple of 16");
    ps.real[i] = src_channels_host[i];
    ps.pad[i] = src_channels_pad_host[i];
    cin += src_channels_host[i];
    kb += src_channels_pad_host[i] / 16;
  }
  const int n_pad = ceil_div(C_out, 16) * 16;
  const long long total = (long long)kb * 9 * 2 * n_pad * 8;
  tc_pack_weights_kernel<<<grid_1d(total), 256, 0, as_stream(stream)>>>(weight, C_out, cin, n_pad, ps, kb,
                                                                        reinterpret_cast<__nv_bfloat16*>(packed));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_conv3x3(const ynet_tc_src* srcs, int32_t n_src, int32_t N, int32_t H, int32_t W, const void* packed_weight,
                    const float* bias, int32_t C_out, int32_t relu, void* out_c8, int32_t C_out_pad, void* stream) {
  YNET_CHECK_ARG(srcs && packed_weight && bias && out_c8, "null pointer");
  YNET_CHECK_ARG(n_src >= 1 && n_src <= YNET_MAX_SOURCES && N >= 0 && H > 0 && W > 0, "bad shape");
  YNET_CHECK_ARG(C_out > 0 && C_out_pad % 16 == 0 && C_out_pad >= C_out && C_out_pad <= 256, "C_out_pad must be a multiple of 16, <= 256");
  YNET_CHECK_ALIGN(out_c8, 16);
  YNET_CHECK_ALIGN(packed_weight, 16);
  if (N == 0) return YNET_OK;
  EncodeTiledFn encode = get_encode();
  if (encode == nullptr) {
    set_error("ynet_tc_conv3x3: cuTensorMapEncodeTiled is not available from the driver");
    return YNET_E_UNSUPPORTED;
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[YNET_MAX_SOURCES];
  memset(maps, 0, sizeof(maps));
  int kb_total = 0;
  for (int i = 0; i < n_src; ++i) {
    const int cp = srcs[i].channels_pad;
    YNET_CHECK_ARG(srcs[i].ptr && cp > 0 && cp % 16 == 0, "source channels_pad must be a positive multiple of 16");
    YNET_CHECK_ALIGN(srcs[i].ptr, 16);
    const bool bcast = srcs[i].batch_stride == 0;
    const int nsrc = bcast ? 1 : (srcs[i].batch_mod > 0 ? srcs[i].batch_mod : N);
    const cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(cp / 8), (cuuint64_t)nsrc};
    const cuuint64_t bs = bcast ? (cuuint64_t)(cp / 8) * H * W * 16 : (cuuint64_t)srcs[i].batch_stride * 2;
    YNET_CHECK_ARG(bs % 16 == 0, "batch stride must be a multiple of 8 elements");
    const cuuint64_t strides[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16, bs};
    const cuuint32_t box[5] = {8, TC_BW, TC_BH, 2, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(srcs[i].ptr), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("ynet_tc_conv3x3: cuTensorMapEncodeTiled failed (%d) for source %d (W=%d H=%d C=%d)", (int)r, i, W, H, cp);
      return YNET_E_CUDA;
    }
    p.src[i].kblocks = cp / 16;
    p.src[i].bcast = bcast ? 1 : 0;
    p.src[i].batch_mod = srcs[i].batch_mod;
    kb_total += cp / 16;
  }
  for (int i = n_src; i < YNET_MAX_SOURCES; ++i) maps[i] = maps[0];
  p.n_src = n_src;
  p.N = N;
  p.H = H;
  p.W = W;
  p.n_pad = C_out_pad;
  p.relu = relu;
  p.tiles_x = ceil_div(W, TC_TW);
  p.tiles_y = ceil_div(H, TC_TH);
  p.total_tiles = (long long)N * p.tiles_x * p.tiles_y;
  p.kb_total = kb_total;
  p.wpacked = reinterpret_cast<const unsigned char*>(packed_weight);
  p.bias = bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_c8);
  p.err = nullptr;

  const int wblk = 9 * 2 * C_out_pad * 16;
  const long long wall = (long long)kb_total * wblk;
  const int budget = 200 * 1024;
  const int tail = (2 * TC_MAX_STAGES + 5) * 8 + 16;
  const char* force_stream = getenv("YNET_TC_FORCE_STREAMED");
  p.resident = (wall + 4 * TC_A_BYTES + tail <= budget) && !(force_stream && force_stream[0] == '1');
  p.wres_bytes = p.resident ? (int)ceil_div<long long>(wall, 1024) * 1024 : 0;
  p.stage_bytes = ceil_div(TC_A_BYTES + (p.resident ? 0 : wblk), 128) * 128;
  p.stages = tmin(TC_MAX_STAGES, (budget - p.wres_bytes - tail) / p.stage_bytes);
  if (p.stages < 2) {
    set_error("ynet_tc_conv3x3: layer does not fit shared memory (C_out_pad=%d)", C_out_pad);
    return YNET_E_UNSUPPORTED;
  }
  int cols = 32;
  while (cols < 2 * C_out_pad) cols *= 2;
  p.tmem_cols = cols;
  const size_t smem_bytes = (size_t)p.wres_bytes + (size_t)p.stages * p.stage_bytes + tail + 1024;

  static size_t configured = 0;
  if (smem_bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "ynet_tc_conv3x3(cudaFuncSetAttribute)");
    configured = 227 * 1024;
  }
  const long long grid = tmin<long long>(p.total_tiles, sm_count());
  tc_conv3x3_kernel<<<(unsigned)grid, TC_THREADS, smem_bytes, as_stream(stream)>>>(maps[0], maps[1], maps[2], maps[3], p);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

}  // extern "C"
