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
#include "tc_common.cuh"

namespace ynet {

constexpr int TC_TH = 16;                            // output tile height; width = 8 * J pixels (J accumulators)
constexpr int TC_KB = 16;                            // channels per pipeline stage (one UMMA K step)
constexpr int TC_MAX_J = 3;                          // 4-D TMA box: (8J+2)*8 elements <= 256
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_STAGES = 32;

struct TcSrcDev {
  int kblocks;      // channels_pad / 16
  int bcast;        // source has batch 1
  int batch_mod;    // > 0: image n reads n % batch_mod; < 0: image n reads n / -batch_mod
  int center;       // 1: only the centre tap is applied to this source (hoisted partial sums, one-tap weights)
  int off;          // 1: the tensor carries a one-pixel replicate-padded ring (dims H + 2, W + 2): shift the TMA box
  int pad_[3];
};
constexpr int TC_MAX_KB = 64;

struct TcParams {
  TcSrcDev src[YNET_MAX_SOURCES];
  int n_src, N, H, W, n_pad /* C_out padded */, relu;
  int tiles_x, tiles_y;
  long long total_tiles;
  int kb_total;           // sum of kblocks
  int with_lo;            // EPI_HILO: also write the low halves
  int hl_planes, hl_off;  // EPI_HILO: 8-channel planes per image of the output tensor (hi then lo halves) and the first hi
                          // plane this conv writes (concat-on-write: the output may be a slice of a wider activation)
  int pad_out;            // EPI_C8: write (H + 2, W + 2) planes with a replicated one-pixel ring (input of an upconv)
  uint64_t center_mask;   // bit kb: K block kb belongs to a centre-tap-only source
  uint64_t quad_mask;     // bit kb: K block kb belongs to a 2x2-neighbourhood source: taps (0,0) (0,1) (1,0) (1,1) only
  int w_total;            // bytes of the packed weights
  int wofs[TC_MAX_KB];    // byte offset of K block kb in the packed weights (9-tap and 1-tap blocks are mixed)
  int resident;           // weights resident in smem
  int stages, stage_bytes, wres_bytes, tmem_cols;
  int j;                  // accumulators (8-pixel column blocks) per tile: consecutive MMAs hit different ones
  int bw;                 // halo box width in pixels = 8 j + 2
  int a_bytes, a_lbo, a_sbo;   // A stage bytes, chunk stride, image-row stride (= 8-pixel-group stride)
  const unsigned char* wpacked;   // [kb][tap][2][n_pad][8] bf16
  const float* bias;              // n_pad floats (pad = 0)
  __nv_bfloat16* out;             // EPI 0: C8 planes, n_pad channels
  float* out_f32;                 // EPI 1: NCHW float32, c_out channels
  float4* partial;                // EPI 2: soft-argmax partials (m, s, sx, sy) [(n * c_out + c) * gridDim.x + cta]
  int c_out;                      // real output channels
  int* err;
  int dbg;                        // profiling aid (YNET_TC_DBG): 1 = epilogue skips the global stores, 2 = skips everything
};

constexpr int EPI_C8 = 0, EPI_NCHW_F32 = 1, EPI_UP2 = 3, EPI_HILO = 4, EPI_PRED = 5;

// ---- the conv kernel -------------------------------------------------------------------------------------
template <int J, int TAPS, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 2)
tc_conv_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
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
  float* s_bias = reinterpret_cast<float*>(w_bar + 2);
  for (int i = threadIdx.x; i < p.n_pad; i += TC_THREADS) s_bias[i] = (p.bias != nullptr) ? p.bias[i] : 0.f;

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
  constexpr int HALO = (TAPS == 9) ? 1 : 0;
  constexpr int BW = 8 * J + 2 * HALO, BH = TC_TH + 2 * HALO;   // TMA box (pixels)
  const int wblk_bytes = TAPS * 2 * p.n_pad * 16;  // weights of one 16-channel K block: [tap][2][n_pad][8] bf16

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // single-thread region ptxas keeps on the uniform datapath (tc_common.cuh)
      if (p.resident) {
        const uint32_t total = (uint32_t)p.w_total;
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
        const int y0 = (r / p.tiles_x) * TC_TH, x0 = (r % p.tiles_x) * 8 * J;
        int kb = 0;
        for (int s = 0; s < p.n_src; ++s) {
          const CUtensorMap* map = (s == 0) ? &map0 : (s == 1) ? &map1 : (s == 2) ? &map2 : &map3;
          const int bm = p.src[s].batch_mod;
          const int ns = p.src[s].bcast ? 0 : (bm > 0 ? n % bm : (bm < 0 ? n / (-bm) : n));
          for (int b = 0; b < p.src[s].kblocks; ++b, ++kb) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.err);
            unsigned char* st = s_stage + (size_t)stage * p.stage_bytes;
            const uint32_t fb = smem_u32(&full_bar[stage]);
            const uint32_t wb = ((p.center_mask >> kb) & 1ull) ? (uint32_t)(wblk_bytes / TAPS)
                                : ((p.quad_mask >> kb) & 1ull) ? (uint32_t)(4 * (wblk_bytes / TAPS))
                                                             : (uint32_t)wblk_bytes;
            mbar_expect_tx(fb, p.a_bytes + (p.resident ? 0u : wb));
            tma_load_4d(smem_u32(st), map, fb, 8 * (x0 - HALO + p.src[s].off), y0 - HALO + p.src[s].off, 2 * b, ns);
            if (!p.resident) bulk_load(smem_u32(st + p.a_bytes), p.wpacked + p.wofs[kb], wb, fb);
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
    if (elect_one()) {   // single-thread region ptxas keeps on the uniform datapath (tc_common.cuh)
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
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * J * p.n_pad);
        // The issuing lane is a scalar instruction stream: keep it to ~3 instructions per MMA.  Descriptors
        // differ only in the start-address field (16-byte units, low word): tap (kh, kw) and column block jj
        // are compile-time offsets into the halo tile, the weight tap advances by a fixed stride.
        constexpr uint32_t A_HI = (uint32_t)((BW * 16) >> 4) | (1u << 14);                 // SBO | version
        constexpr uint32_t A_LBO_FIELD = (uint32_t)((BH * BW * 16) >> 4) << 16;             // LBO
        const uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t b_lbo_field = (uint32_t)((p.n_pad * 16) >> 4) << 16;
        const uint32_t b_tap_step = (uint32_t)(2 * p.n_pad);                                // 16-byte units per tap
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase, p.err);
          tc_fence_after();
          unsigned char* st = s_stage + (size_t)stage * p.stage_bytes;
          const uint32_t a_lo0 = ((smem_u32(st) >> 4) & 0x3FFF) | A_LBO_FIELD;
          const uint32_t b_base = p.resident ? smem_u32(s_w + p.wofs[kb]) : smem_u32(st + p.a_bytes);
          uint32_t b_lo = ((b_base >> 4) & 0x3FFF) | b_lbo_field;
          const uint32_t acc_first = (kb > 0) ? 1u : 0u;
          if (TAPS == 9 && ((p.center_mask >> kb) & 1ull)) {
            // hoisted partial sums: identity weights on the centre tap only
            const uint64_t bdesc = ((uint64_t)b_hi << 32) | b_lo;
#pragma unroll
            for (int jj = 0; jj < J; ++jj) {
              const uint64_t adesc = ((uint64_t)A_HI << 32) | (a_lo0 + (uint32_t)(BW + 1 + 8 * jj));
              tc_mma_bf16(d_tmem + (uint32_t)(jj * p.n_pad), adesc, bdesc, idesc, acc_first);
            }
          } else if (TAPS == 9 && ((p.quad_mask >> kb) & 1ull)) {
            // 2x2-neighbourhood planes (waypoint maps): the four taps anchored at (-1,-1) (-1,0) (0,-1) (0,0) cover
            // the 3x3 window; the packed weights hold exactly these four taps
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
              const uint64_t bdesc = ((uint64_t)b_hi << 32) | b_lo;
#pragma unroll
              for (int jj = 0; jj < J; ++jj) {
                const uint64_t adesc = ((uint64_t)A_HI << 32) | (a_lo0 + (uint32_t)((t4 >> 1) * BW + (t4 & 1) + 8 * jj));
                tc_mma_bf16(d_tmem + (uint32_t)(jj * p.n_pad), adesc, bdesc, idesc, t4 == 0 ? acc_first : 1u);
              }
              b_lo += b_tap_step;
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint64_t bdesc = ((uint64_t)b_hi << 32) | b_lo;
#pragma unroll
              for (int jj = 0; jj < J; ++jj) {
                // consecutive MMAs target different accumulators (column blocks of the tile)
                const uint64_t adesc =
                    ((uint64_t)A_HI << 32) | (a_lo0 + (uint32_t)((TAPS == 9 ? (tap / 3) * BW + (tap % 3) : 0) + 8 * jj));
                tc_mma_bf16(d_tmem + (uint32_t)(jj * p.n_pad), adesc, bdesc, idesc, tap == 0 ? acc_first : 1u);
              }
              b_lo += b_tap_step;
            }
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
    // ===================== epilogue: 4 warps; warp w owns TMEM lanes 32*(w%4) .. +31 =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;         // accumulator row = pixel of the tile
    const int py = m >> 3, px = m & 7;
    const int n_chunks = p.n_pad >> 3;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int n = (int)(tile / tiles_per_img);
      const int r = (int)(tile - (long long)n * tiles_per_img);
      const int y0 = (r / p.tiles_x) * TC_TH, x0 = (r % p.tiles_x) * 8 * J;
      const int y = y0 + py, xb = x0 + px;
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, p.err);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * J * p.n_pad);
#pragma unroll
      for (int jj = 0; jj < J; ++jj) {
        const int x = xb + 8 * jj;
        const bool inb = (y < p.H) && (x < p.W) && p.dbg == 0;
        if (p.dbg == 2) continue;
        if (EPI == EPI_UP2) {
          // phase-decomposed bilinear-x2 + conv: the four phase groups of a low-resolution pixel (y, x) land on the
          // high-resolution pixels (2y + a, 2x + b).  The two b phases of one row are ADJACENT pixels of the C8 plane:
          // both are fetched and written as one 32-byte store (full sectors instead of two half-filled ones).
          const int cp = p.n_pad >> 2;                 // padded C_out of one phase (multiple of 16)
          for (int a = 0; a < 2; ++a) {
            for (int o0 = 0; o0 < cp; o0 += 16) {
              uint32_t v0[16], v1[16];
              tmem_ld16_nowait(t_row + (uint32_t)(jj * p.n_pad + (2 * a) * cp + o0), v0);
              tmem_ld16_nowait(t_row + (uint32_t)(jj * p.n_pad + (2 * a + 1) * cp + o0), v1);
              tmem_ld_wait(v0);
              tmem_ld_wait(v1);
              if (inb) {
                const float4* b0 = reinterpret_cast<const float4*>(s_bias + (2 * a) * cp + o0);
                const float4* b1 = reinterpret_cast<const float4*>(s_bias + (2 * a + 1) * cp + o0);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const float4 ba = b0[2 * h], bb = b0[2 * h + 1], bc = b1[2 * h], bd = b1[2 * h + 1];
                  uint32_t o[8];
                  o[0] = pack_bf16(__uint_as_float(v0[8 * h + 0]) + ba.x, __uint_as_float(v0[8 * h + 1]) + ba.y);
                  o[1] = pack_bf16(__uint_as_float(v0[8 * h + 2]) + ba.z, __uint_as_float(v0[8 * h + 3]) + ba.w);
                  o[2] = pack_bf16(__uint_as_float(v0[8 * h + 4]) + bb.x, __uint_as_float(v0[8 * h + 5]) + bb.y);
                  o[3] = pack_bf16(__uint_as_float(v0[8 * h + 6]) + bb.z, __uint_as_float(v0[8 * h + 7]) + bb.w);
                  o[4] = pack_bf16(__uint_as_float(v1[8 * h + 0]) + bc.x, __uint_as_float(v1[8 * h + 1]) + bc.y);
                  o[5] = pack_bf16(__uint_as_float(v1[8 * h + 2]) + bc.z, __uint_as_float(v1[8 * h + 3]) + bc.w);
                  o[6] = pack_bf16(__uint_as_float(v1[8 * h + 4]) + bd.x, __uint_as_float(v1[8 * h + 5]) + bd.y);
                  o[7] = pack_bf16(__uint_as_float(v1[8 * h + 6]) + bd.z, __uint_as_float(v1[8 * h + 7]) + bd.w);
                  const int chunk = (o0 >> 3) + h;
                  __nv_bfloat16* dst =
                      p.out + ((((size_t)n * (cp >> 3) + chunk) * (2 * p.H) + (2 * y + a)) * (size_t)(2 * p.W) + 2 * x) * 8;
                  st_global_v8(dst, o);
                }
              }
            }
          }
          continue;
        }
        uint32_t v[16], vn[16];
        tmem_ld16(t_row + (uint32_t)(jj * p.n_pad), v);
        for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
          // the next 16 columns are in flight while these are converted and stored
          const bool more = c0 + 16 < p.n_pad;
          if (more) tmem_ld16_nowait(t_row + (uint32_t)(jj * p.n_pad + c0 + 16), vn);
          float f[16];
          {
            // s_bias is 16-byte aligned (barrier block of 70 x 8 bytes after 128-byte-aligned stages): four LDS.128
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const float4 b = b4[k4];
              f[4 * k4 + 0] = __uint_as_float(v[4 * k4 + 0]) + b.x;
              f[4 * k4 + 1] = __uint_as_float(v[4 * k4 + 1]) + b.y;
              f[4 * k4 + 2] = __uint_as_float(v[4 * k4 + 2]) + b.z;
              f[4 * k4 + 3] = __uint_as_float(v[4 * k4 + 3]) + b.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int k = 0; k < 16; ++k) f[k] = fmaxf(f[k], 0.f);
          }
          if (EPI == EPI_C8) {
            if (inb) {
              const int po = p.pad_out;
              const int Hp = p.H + 2 * po, Wp = p.W + 2 * po;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int chunk = (c0 >> 3) + h;
                uint4 o;
                o.x = pack_bf16(f[8 * h + 0], f[8 * h + 1]);
                o.y = pack_bf16(f[8 * h + 2], f[8 * h + 3]);
                o.z = pack_bf16(f[8 * h + 4], f[8 * h + 5]);
                o.w = pack_bf16(f[8 * h + 6], f[8 * h + 7]);
                uint4* dst = reinterpret_cast<uint4*>(p.out) + (((size_t)n * n_chunks + chunk) * Hp + (y + po)) * Wp + (x + po);
                *dst = o;
                if (po) {
                  // replicate the border pixels into the ring (the bilinear index clamping of the consumer)
                  const int dyv = (y == 0) ? -1 : ((y == p.H - 1) ? 1 : 0);
                  const int dxv = (x == 0) ? -1 : ((x == p.W - 1) ? 1 : 0);
                  if (dyv != 0) dst[dyv * Wp] = o;
                  if (dxv != 0) dst[dxv] = o;
                  if (dyv != 0 && dxv != 0) dst[dyv * Wp + dxv] = o;
                  if (p.H == 1 && y == 0) {      // a one-row image is both borders
                    dst[Wp] = o;
                    if (dxv != 0) dst[Wp + dxv] = o;
                  }
                  if (p.W == 1 && x == 0) {
                    dst[1] = o;
                    if (dyv != 0) dst[dyv * Wp + 1] = o;
                    if (p.H == 1) dst[Wp + 1] = o;
                  }
                }
              }
            }
          } else if (EPI == EPI_HILO) {
            // raw partial sums as a bf16 pair: hi = bf16(f), lo = bf16(f - hi); channels [hi: n_pad | lo: n_pad]
            if (inb) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int chunk = (c0 >> 3) + h;
                float hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  hi[k] = __bfloat162float(__float2bfloat16_rn(f[8 * h + k]));
                  lo[k] = f[8 * h + k] - hi[k];
                }
                uint4 o, ol;
                o.x = pack_bf16(hi[0], hi[1]);
                o.y = pack_bf16(hi[2], hi[3]);
                o.z = pack_bf16(hi[4], hi[5]);
                o.w = pack_bf16(hi[6], hi[7]);
                ol.x = pack_bf16(lo[0], lo[1]);
                ol.y = pack_bf16(lo[2], lo[3]);
                ol.z = pack_bf16(lo[4], lo[5]);
                ol.w = pack_bf16(lo[6], lo[7]);
                __nv_bfloat16* dst = p.out + ((((size_t)n * p.hl_planes + p.hl_off + chunk) * p.H + y) * p.W + x) * 8;
                *reinterpret_cast<uint4*>(dst) = o;
                if (p.with_lo) *reinterpret_cast<uint4*>(dst + (size_t)(p.hl_planes >> 1) * p.H * p.W * 8) = ol;
              }
            }
          } else if (EPI == EPI_NCHW_F32) {
            if (inb) {
#pragma unroll
              for (int k = 0; k < 16; ++k)
                if (c0 + k < p.c_out) p.out_f32[(((size_t)n * p.c_out + c0 + k) * p.H + y) * p.W + x] = f[k];
            }
          }
          if (more) {
            tmem_ld_wait(vn);
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = vn[k];
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

// ---- 3x3 conv + ReLU -> 1x1 predictor -> SoftArgmax2D in ONE kernel (decoder.4.2 + predictor + softargmax) ----------
// The conv's bf16 output tile never leaves the SM: the conv epilogue writes it to shared memory in the K-major C8
// layout, the MMA warp multiplies it (as the N = 256-pixel operand) with the replicated predictor weights (M = 128:
// see pred_tc.cu) and 16 soft-argmax warps reduce the transposed logits (TMEM lane = channel) straight from
// tcgen05.ld.  Saves the 64 B/pixel write + 64 B/pixel read of the activation and one launch; the soft-argmax warps
// (MUFU / issue bound) run concurrently with the conv's MMAs (shared-memory-operand bound).
// TMEM: conv accumulators 2 x (2 x 32) columns at [0, 128), predictor accumulator 256 columns at [256, 512).
struct PredFuse {
  const unsigned char* pw;   // predictor weights [kb][1][2][pn_pad][8] bf16
  const float* pbias;        // pn_pad floats
  float4* partial;
  int c_pred, pn_pad, slots;
  long long tiles_per_cta;
};
constexpr int FP_CONV_EPI_WARPS = 8;   // two per TMEM lane quadrant: one per column block (4 warps alone bound the tile)
constexpr int FP_J = 2;
// warp 0: TMA; warps 1..2: MMA issuers (one per column block: a single issuing thread needs ~45 cycles of descriptor
// set-up per MMA and would bound the tile); 8 conv-epilogue warps; 16 soft-argmax warps
constexpr int FP_THREADS = 32 * (1 + FP_J + FP_CONV_EPI_WARPS + PR_EPI_WARPS);   // 864

__global__ void __launch_bounds__(FP_THREADS, 1)
tc_conv_pred_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                    const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3, const TcParams p,
                    const PredFuse f) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int J = FP_J, BW = 8 * J + 2, BH = TC_TH + 2;
  const int y_chunks = p.n_pad >> 3;                 // 8-channel planes of the conv output
  const int y_bytes = y_chunks * 256 * 16;           // one Y buffer: [chunk][16 rows][16 px][8 ch]
  const int kbp = p.n_pad >> 4;                      // predictor K blocks

  unsigned char* s_w = smem;
  unsigned char* s_stage = smem + p.wres_bytes;
  unsigned char* s_y = s_stage + (size_t)p.stages * p.stage_bytes;
  unsigned char* s_pw = s_y + 2 * y_bytes;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_pw + (size_t)kbp * PR_WBLK_BYTES);
  uint64_t* full_bar = s_bar;
  uint64_t* empty_bar = s_bar + TC_MAX_STAGES;
  uint64_t* tfull_bar = s_bar + 2 * TC_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint64_t* y_bar = w_bar + 1;          // [2] conv output tile written (4 epilogue warps)
  uint64_t* yfree_bar = y_bar + 2;      // [2] predictor MMAs that read the buffer have retired
  uint64_t* pfull_bar = yfree_bar + 2;  // predictor accumulator complete
  uint64_t* pempty_bar = pfull_bar + 1; // predictor accumulator drained (16 soft-argmax warps)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(pempty_bar + 1);
  float* s_bias = reinterpret_cast<float*>(pempty_bar + 2);
  for (int i = threadIdx.x; i < p.n_pad; i += FP_THREADS) s_bias[i] = p.bias[i];
  pred_stage_weights(s_pw, f.pw, kbp, f.pn_pad, threadIdx.x, FP_THREADS);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), J);     // one tcgen05.commit per MMA-issuing warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), J);
      mbar_init(smem_u32(&tempty_bar[a]), FP_CONV_EPI_WARPS);
      mbar_init(smem_u32(&y_bar[a]), FP_CONV_EPI_WARPS);
      mbar_init(smem_u32(&yfree_bar[a]), 1);
    }
    mbar_init(smem_u32(w_bar), 1);
    mbar_init(smem_u32(pfull_bar), 1);
    mbar_init(smem_u32(pempty_bar), PR_EPI_WARPS);
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
  const uint32_t pacc = tmem_base + 256u;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const long long t0 = (long long)blockIdx.x * f.tiles_per_cta;
  const long long t1 = tmin<long long>(p.total_tiles, t0 + f.tiles_per_cta);
  // tile coordinates advance incrementally in every role
  int n = (int)(t0 / tiles_per_img);
  const int r0t = (int)(t0 - (long long)n * tiles_per_img);
  int ty = r0t / p.tiles_x, tx = r0t - ty * p.tiles_x;
  auto next_tile = [&]() {
    if (++tx == p.tiles_x) {
      tx = 0;
      if (++ty == p.tiles_y) {
        ty = 0;
        ++n;
      }
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // single-thread region ptxas keeps on the uniform datapath (tc_common.cuh)
      const uint32_t total = (uint32_t)p.w_total;
      mbar_expect_tx(smem_u32(w_bar), total);
      for (uint32_t off = 0; off < total; off += 32768) {
        const uint32_t nb = min(32768u, total - off);
        bulk_load(smem_u32(s_w + off), p.wpacked + off, nb, smem_u32(w_bar));
      }
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = t0; tile < t1; ++tile) {
        const int y0 = ty * TC_TH, x0 = tx * 8 * J;
        int kb = 0;
        for (int s = 0; s < p.n_src; ++s) {
          const CUtensorMap* map = (s == 0) ? &map0 : (s == 1) ? &map1 : (s == 2) ? &map2 : &map3;
          const int bm = p.src[s].batch_mod;
          const int ns = p.src[s].bcast ? 0 : (bm > 0 ? n % bm : (bm < 0 ? n / (-bm) : n));
          for (int b = 0; b < p.src[s].kblocks; ++b, ++kb) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, nullptr);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_expect_tx(fb, p.a_bytes);
            tma_load_4d(smem_u32(s_stage + (size_t)stage * p.stage_bytes), map, fb, 8 * (x0 - 1), y0 - 1, 2 * b, ns);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        next_tile();
      }
    }
  } else if (warp <= J) {
    // ===================== MMA issuers: warp 1 + jj issues the conv MMAs of column block jj; warp 1 also issues the
    // predictor MMAs of tile i - 1 after the conv MMAs of tile i =====================
    const int jj = warp - 1;
    if (elect_one()) {   // single-thread region ptxas keeps on the uniform datapath (tc_common.cuh)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_p = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
      auto issue_pred = [&](uint32_t j) {
        const uint32_t b = j & 1u;
        mbar_wait(smem_u32(&y_bar[b]), (j >> 1) & 1u, nullptr);          // conv output tile j is in shared memory
        mbar_wait(smem_u32(pempty_bar), (j & 1u) ^ 1u, nullptr);         // soft-argmax of tile j - 1 has drained pacc
        tc_fence_after();
        for (int kb = 0; kb < kbp; ++kb) {
          const uint64_t adesc = make_desc(smem_u32(s_pw + (size_t)kb * PR_WBLK_BYTES), 128 * 16, 128);
          const uint64_t bdesc = make_desc(smem_u32(s_y + (size_t)b * y_bytes + (size_t)kb * 2 * 4096), 4096, 128);
          tc_mma_bf16(pacc, adesc, bdesc, idesc_p, kb > 0 ? 1u : 0u);
        }
        tc_commit(smem_u32(pfull_bar));
        tc_commit(smem_u32(&yfree_bar[b]));
      };
      mbar_wait(smem_u32(w_bar), 0, nullptr);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      constexpr uint32_t A_HI = (uint32_t)((BW * 16) >> 4) | (1u << 14);
      constexpr uint32_t A_LBO_FIELD = (uint32_t)((BH * BW * 16) >> 4) << 16;
      const uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);
      const uint32_t b_lbo_field = (uint32_t)((p.n_pad * 16) >> 4) << 16;
      const uint32_t b_tap_step = (uint32_t)(2 * p.n_pad);
      for (long long tile = t0; tile < t1; ++tile, ++it) {
        const uint32_t acc = it & 1u;
        mbar_wait(smem_u32(&tempty_bar[acc]), ((it >> 1) & 1u) ^ 1u, nullptr);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)(J * p.n_pad);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase, nullptr);
          tc_fence_after();
          const uint32_t a_lo0 = ((smem_u32(s_stage + (size_t)stage * p.stage_bytes) >> 4) & 0x3FFF) | A_LBO_FIELD;
          uint32_t b_lo = ((smem_u32(s_w + p.wofs[kb]) >> 4) & 0x3FFF) | b_lbo_field;
          const uint32_t acc_first = (kb > 0) ? 1u : 0u;
          if ((p.center_mask >> kb) & 1ull) {
            const uint64_t bdesc = ((uint64_t)b_hi << 32) | b_lo;
            {
              const uint64_t adesc = ((uint64_t)A_HI << 32) | (a_lo0 + (uint32_t)(BW + 1 + 8 * jj));
              tc_mma_bf16(d_tmem + (uint32_t)(jj * p.n_pad), adesc, bdesc, idesc, acc_first);
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint64_t bdesc = ((uint64_t)b_hi << 32) | b_lo;
              {
                const uint64_t adesc = ((uint64_t)A_HI << 32) | (a_lo0 + (uint32_t)((tap / 3) * BW + (tap % 3) + 8 * jj));
                tc_mma_bf16(d_tmem + (uint32_t)(jj * p.n_pad), adesc, bdesc, idesc, tap == 0 ? acc_first : 1u);
              }
              b_lo += b_tap_step;
            }
          }
          tc_commit(smem_u32(&empty_bar[stage]));
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(smem_u32(&tfull_bar[acc]));
        if (jj == 0 && it >= 1) issue_pred(it - 1);
      }
      if (jj == 0 && it >= 1) issue_pred(it - 1);
    }
  } else if (warp < 1 + J + FP_CONV_EPI_WARPS) {
    // ===================== conv epilogue: accumulator -> bias -> ReLU -> bf16 -> shared memory (K-major C8) =============
    const int q = warp & 3;
    const int jj = (warp - (1 + J)) >> 2;          // column block of this warp
    const int m = q * 32 + lane;
    const int py = m >> 3, px = m & 7;
    uint32_t it = 0;
    for (long long tile = t0; tile < t1; ++tile, ++it) {
      const uint32_t acc = it & 1u;
      mbar_wait(smem_u32(&tfull_bar[acc]), (it >> 1) & 1u, nullptr);
      mbar_wait(smem_u32(&yfree_bar[acc]), ((it >> 1) & 1u) ^ 1u, nullptr);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)(J * p.n_pad);
      unsigned char* yb = s_y + (size_t)acc * y_bytes + py * 256 + px * 16;
      {
        for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_row + (uint32_t)(jj * p.n_pad + c0), v);
          float fv[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            fv[k] = __uint_as_float(v[k]) + s_bias[c0 + k];
            if (p.relu) fv[k] = fmaxf(fv[k], 0.f);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 o;
            o.x = pack_bf16(fv[8 * h + 0], fv[8 * h + 1]);
            o.y = pack_bf16(fv[8 * h + 2], fv[8 * h + 3]);
            o.z = pack_bf16(fv[8 * h + 4], fv[8 * h + 5]);
            o.w = pack_bf16(fv[8 * h + 6], fv[8 * h + 7]);
            *reinterpret_cast<uint4*>(yb + ((c0 >> 3) + h) * 4096 + jj * 128) = o;
          }
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&tempty_bar[acc]));
        mbar_arrive(smem_u32(&y_bar[acc]));
      }
    }
  } else {
    // ===================== soft-argmax: 16 warps, one tile row each; TMEM lane = channel =====================
    const int e = warp - (1 + J + FP_CONV_EPI_WARPS);
    const int q = warp & 3;
    const int row = q * 4 + (e >> 2);
    const bool active = lane < f.c_pred;
    const float bias = active ? f.pbias[lane] : 0.f;
    SoftState st{PR_NEG, 0.f, 0.f, 0.f};
    int cur_n = -1;
    auto flush = [&](int n_img) {
      if (active) {
        const long long first_cta = ((long long)n_img * tiles_per_img) / f.tiles_per_cta;
        const int slot = (int)(blockIdx.x - first_cta) * PR_EPI_WARPS + e;
        f.partial[((size_t)n_img * f.c_pred + lane) * f.slots + slot] = make_float4(st.m, st.s, st.sx, st.sy);
      }
    };
    uint32_t it = 0;
    for (long long tile = t0; tile < t1; ++tile, ++it) {
      const int y = ty * TC_TH + row, x0 = tx * 8 * J;
      if (n != cur_n) {
        if (cur_n >= 0) flush(cur_n);
        st = SoftState{PR_NEG, 0.f, 0.f, 0.f};
        cur_n = n;
      }
      mbar_wait(smem_u32(pfull_bar), it & 1u, nullptr);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(pacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(row * 16), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(pempty_bar));      // the values are in registers: release the accumulator
      if (y < p.H) {
        if (x0 + 8 * J <= p.W)
          softargmax_row16(st, v, bias, x0, y);
        else
          softargmax_row16_masked(st, v, bias, x0, y, p.W);
      }
      next_tile();
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

// bilinear x2 (align_corners=False) on C8 planes: (H, W) input size.
// out[2i] = 0.25 in[i-1] + 0.75 in[i], out[2i+1] = 0.75 in[i] + 0.25 in[i+1] (indices clamped).  One thread owns the
// 2x2 output block between input cells (ci-1, ci) x (cj-1, cj): 4 loads feed 4 outputs (the naive form needs 16).
__global__ void __launch_bounds__(256)
c8_upsample_kernel(const uint4* __restrict__ x, long long planes, int H, int W, uint4* __restrict__ out) {
  const int OW = 2 * W;
  const int CH = H + 1, CW = W + 1;
  const long long total = planes * CH * CW;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / ((long long)CH * CW);
    const int r = (int)(t - pl * CH * CW);
    const int ci = r / CW, cj = r - ci * CW;
    const int i0 = max(ci - 1, 0), i1 = min(ci, H - 1);
    const int j0 = max(cj - 1, 0), j1 = min(cj, W - 1);
    const uint4* q = x + pl * H * W;
    float a[8], b[8], c[8], d[8];
    unpack8(__ldg(q + (size_t)i0 * W + j0), a);
    unpack8(__ldg(q + (size_t)i0 * W + j1), b);
    unpack8(__ldg(q + (size_t)i1 * W + j0), c);
    unpack8(__ldg(q + (size_t)i1 * W + j1), d);
    float top_l[8], top_r[8], bot_l[8], bot_r[8];   // horizontally interpolated rows: left = col 2cj-1, right = col 2cj
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      top_l[k] = 0.75f * a[k] + 0.25f * b[k];
      top_r[k] = 0.25f * a[k] + 0.75f * b[k];
      bot_l[k] = 0.75f * c[k] + 0.25f * d[k];
      bot_r[k] = 0.25f * c[k] + 0.75f * d[k];
    }
    uint4* o = out + pl * 4 * H * W;
    const bool has_l = cj >= 1, has_r = cj <= W - 1;
    float v[8];
    if (ci >= 1) {                // output row 2ci-1 = 0.75 r0 + 0.25 r1
      const size_t row = (size_t)(2 * ci - 1) * OW;
      if (has_l) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.75f * top_l[k] + 0.25f * bot_l[k];
        o[row + 2 * cj - 1] = pack8(v);
      }
      if (has_r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.75f * top_r[k] + 0.25f * bot_r[k];
        o[row + 2 * cj] = pack8(v);
      }
    }
    if (ci <= H - 1) {            // output row 2ci = 0.25 r0 + 0.75 r1
      const size_t row = (size_t)(2 * ci) * OW;
      if (has_l) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.25f * top_l[k] + 0.75f * bot_l[k];
        o[row + 2 * cj - 1] = pack8(v);
      }
      if (has_r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.25f * top_r[k] + 0.75f * bot_r[k];
        o[row + 2 * cj] = pack8(v);
      }
    }
  }
}

// one-pixel replicate padding of C8 planes: (planes, H, W) -> (planes, H + 2, W + 2)
__global__ void __launch_bounds__(256)
c8_pad_replicate_kernel(const uint4* __restrict__ x, long long planes, int H, int W, uint4* __restrict__ out) {
  const int Hp = H + 2, Wp = W + 2;
  const long long total = planes * Hp * Wp;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / ((long long)Hp * Wp);
    const int r = (int)(t - pl * Hp * Wp);
    const int y = min(max(r / Wp - 1, 0), H - 1), xx = min(max(r % Wp - 1, 0), W - 1);
    out[t] = __ldg(x + (pl * H + y) * W + xx);
  }
}

// weights OIHW f32 -> [kb][tap][2][n_pad][8] bf16 over the concatenated, per-source padded input channels
struct PackSrc {
  int real[YNET_MAX_SOURCES], pad[YNET_MAX_SOURCES];
  int n_src;
};
__global__ void __launch_bounds__(256)
tc_pack_weights_kernel(const float* __restrict__ w, int C_out, int C_in, int n_pad, PackSrc ps, int kb_total, int taps,
                       __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)kb_total * taps * 2 * n_pad * 8;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int k8 = (int)(t & 7);
    long long r = t >> 3;
    const int n = (int)(r % n_pad);
    r /= n_pad;
    const int c = (int)(r & 1);
    r >>= 1;
    const int tap = (int)(r % taps);
    const int kb = (int)(r / taps);
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
    if (ci >= 0 && n < C_out) v = w[((size_t)n * C_in + ci) * taps + tap];
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

// ---- bilinear x2 + 3x3 conv as ONE low-resolution conv with 4 x C_out phase channels ------------------------
// F.interpolate(scale 2, bilinear, align_corners=False) followed by conv3x3(pad 1) (ynet.py:463-464) is linear in
// the low-resolution input x: the output pixel (2i + a, 2j + b) is a 3x3 stencil over x[i-1..i+1][j-1..j+1] whose
// weights are  W_eff^{ab}[o,c,p,q] = sum_{dy,dx} w[o,c,dy,dx] alpha_a[dy][p] alpha_b[dx][q]  with
//   alpha_0 = {{.75,.25,0},{.25,.75,0},{0,.75,.25}}   (rows: conv tap -1,0,+1; cols: low-res offset -1,0,+1)
//   alpha_1 = {{.25,.75,0},{0,.75,.25},{0,.25,.75}}.
// The upsampled tensor (4x the pixels) is never materialised and the MMA N dimension becomes 4 x C_out.
// Exact in the interior; the one-pixel low-resolution border ring (index clamping of the interpolation and the
// zero padding of the conv do not commute with the stencil) is recomputed by upconv_border_kernel.
__constant__ float kAlpha[2][3][3] = {{{.75f, .25f, 0.f}, {.25f, .75f, 0.f}, {0.f, .75f, .25f}},
                                      {{.25f, .75f, 0.f}, {0.f, .75f, .25f}, {0.f, .25f, .75f}}};

// w (C_out, C_in, 3, 3) -> w_eff (4 * cp, C_in, 3, 3), rows (a, b, o) with o padded to cp; bias -> bias_eff (4 * cp)
__global__ void __launch_bounds__(256)
upconv_phase_weights_kernel(const float* __restrict__ w, const float* __restrict__ bias, int C_out, int C_in, int cp,
                            float* __restrict__ w_eff, float* __restrict__ bias_eff) {
  const int total = 4 * cp * C_in * 9;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int q = t % 3, pp = (t / 3) % 3;
    const int c = (t / 9) % C_in;
    const int row = t / (9 * C_in);
    const int o = row % cp, phase = row / cp;
    const int a = phase >> 1, b = phase & 1;
    float v = 0.f;
    if (o < C_out) {
      const float* wk = w + ((size_t)o * C_in + c) * 9;
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) v += wk[dy * 3 + dx] * kAlpha[a][dy][pp] * kAlpha[b][dx][q];
    }
    w_eff[t] = v;
  }
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 4 * cp; t += gridDim.x * blockDim.x)
    bias_eff[t] = (t % cp < C_out && bias != nullptr) ? bias[t % cp] : 0.f;
}

// Exact recomputation of the high-resolution outputs that belong to low-resolution border pixels (CUDA cores; the
// ring is ~2 % of the pixels at 416^2).  One CTA column (blockIdx.y) = one group of OG output channels whose float32
// weights [padded input channel][tap][OG] stay resident in shared memory; threads walk (image, ring pixel, phase).
constexpr int UPB_THREADS = 256;
struct UpBorderSrc {
  const uint4* ptr[YNET_MAX_SOURCES];
  long long batch_stride[YNET_MAX_SOURCES];   // in uint4 (pixels x chunks); 0 = broadcast
  int real_chunks[YNET_MAX_SOURCES];          // ceil(real channels / 8)
  int batch_mod[YNET_MAX_SOURCES];
  int n_src;
};
struct UpBorderPack {
  int real[YNET_MAX_SOURCES];
  int n_src;
};

// weight (C_out, C_in, 3, 3) -> [group][padded channel][tap][OG] float32, channels padded per source to 8
__global__ void __launch_bounds__(256)
upconv_border_weights_kernel(const float* __restrict__ w, int C_out, int C_in, UpBorderPack ps, int cpad_total, int og,
                             int groups, float* __restrict__ out) {
  const int total = groups * cpad_total * 9 * og;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int o_in = t % og;
    const int tap = (t / og) % 9;
    const int cpad = (t / (og * 9)) % cpad_total;
    const int g = t / (og * 9 * cpad_total);
    int c = cpad, ci = -1, base = 0;
    for (int s = 0; s < ps.n_src; ++s) {
      const int padded = (ps.real[s] + 7) / 8 * 8;
      if (c < padded) {
        if (c < ps.real[s]) ci = base + c;
        break;
      }
      c -= padded;
      base += ps.real[s];
    }
    const int o = g * og + o_in;
    out[t] = (ci >= 0 && o < C_out) ? w[((size_t)o * C_in + ci) * 9 + tap] : 0.f;
  }
}

template <int OG>   // output channels per thread: 16 or 32
__global__ void __launch_bounds__(UPB_THREADS, 2)
upconv_border_kernel(UpBorderSrc src, int N, int h, int w, const float* __restrict__ bw /* [g][cpad][9][OG] */,
                     const float* __restrict__ bias, int C_out, int cpad_total, __nv_bfloat16* __restrict__ out, int cp) {
  extern __shared__ __align__(16) float s_wt[];   // [cpad_total][9][OG]
  const int grp = blockIdx.y;
  {
    const float4* g4 = reinterpret_cast<const float4*>(bw + (size_t)grp * cpad_total * 9 * OG);
    float4* s4 = reinterpret_cast<float4*>(s_wt);
    for (int e = threadIdx.x; e < cpad_total * 9 * OG / 4; e += UPB_THREADS) s4[e] = __ldg(g4 + e);
  }
  __syncthreads();
  const int ring = (h >= 2 && w >= 2) ? 2 * w + 2 * (h - 2) : h * w;
  const long long total = (long long)N * ring * 4;
  const int H2 = 2 * h, W2 = 2 * w;
  for (long long t = (long long)blockIdx.x * UPB_THREADS + threadIdx.x; t < total; t += (long long)gridDim.x * UPB_THREADS) {
    const int phase = (int)(t & 3);
    const long long r0 = t >> 2;
    const int n = (int)(r0 / ring);
    int r = (int)(r0 - (long long)n * ring);
    int i, j;
    if (h < 2 || w < 2) {
      i = r / w;
      j = r - i * w;
    } else if (r < w) {
      i = 0;
      j = r;
    } else if (r < 2 * w) {
      i = h - 1;
      j = r - w;
    } else {
      r -= 2 * w;
      i = 1 + (r >> 1);
      j = (r & 1) ? w - 1 : 0;
    }
    const int v = 2 * i + (phase >> 1), u = 2 * j + (phase & 1);   // high-resolution output pixel
    // The 3x3 hi-res window around (v, u) draws on the low-res rows i-1..i+1 / columns j-1..j+1 (clamped into the
    // image = the interpolation's index clamping).  ay[d][r]: weight of low-res row slot r (0..2) in hi-res row
    // v + d - 1; all-zero when that hi-res row lies in the conv's zero padding.  Same for columns.
    float ay[3][3], ax[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int vv = v + d - 1, uu = u + d - 1;
      const bool vin = vv >= 0 && vv < H2, uin = uu >= 0 && uu < W2;
      // hi-res index 2k: .25 x[k-1] + .75 x[k]; 2k+1: .75 x[k] + .25 x[k+1]  (slots relative to i-1 / j-1)
      const int ky = (vv >> 1) - (i - 1), kx = (uu >> 1) - (j - 1);
      const int ya = (vv & 1) ? ky : ky - 1, yb = (vv & 1) ? ky + 1 : ky;
      const int xa = (uu & 1) ? kx : kx - 1, xb = (uu & 1) ? kx + 1 : kx;
      const float fa = (vv & 1) ? .75f : .25f, fb = 1.f - fa;
      const float ga = (uu & 1) ? .75f : .25f, gb = 1.f - ga;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        // slot r holds low-res row clamp(i - 1 + r): an (unclamped) tap row i - 1 + r lands on its own slot
        ay[d][r] = vin ? ((ya == r ? fa : 0.f) + (yb == r ? fb : 0.f)) : 0.f;
        ax[d][r] = uin ? ((xa == r ? ga : 0.f) + (xb == r ? gb : 0.f)) : 0.f;
      }
    }
    int rows[3], cols[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      rows[r] = min(max(i - 1 + r, 0), h - 1);
      cols[r] = min(max(j - 1 + r, 0), w - 1);
    }
    float acc[OG];
#pragma unroll
    for (int o = 0; o < OG; ++o) {
      const int oo = grp * OG + o;
      acc[o] = (bias != nullptr && oo < C_out) ? bias[oo] : 0.f;
    }
    int cpad = 0;
    for (int s = 0; s < src.n_src; ++s) {
      const int bm = src.batch_mod[s];
      const int ns = (src.batch_stride[s] == 0) ? 0 : (bm > 0 ? n % bm : (bm < 0 ? n / (-bm) : n));
      const uint4* xs = src.ptr[s] + (size_t)ns * src.batch_stride[s];
      for (int ch = 0; ch < src.real_chunks[s]; ++ch, cpad += 8) {
        const uint4* xc = xs + (size_t)ch * h * w;
        const float* wq = s_wt + (size_t)cpad * 9 * OG;
        uint4 X[3][3];   // packed 3x3 low-res neighbourhood (row slot, column slot)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) X[r][c] = __ldg(xc + (size_t)rows[r] * w + cols[c]);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          // vertical interpolation: R[c] = hi-res row (v + dy - 1) at low-res column slot c
          float R[3][8];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float x0[8], x1[8], x2[8];
            unpack8(X[0][c], x0);
            unpack8(X[1][c], x1);
            unpack8(X[2][c], x2);
#pragma unroll
            for (int k = 0; k < 8; ++k) R[c][k] = ay[dy][0] * x0[k] + ay[dy][1] * x1[k] + ay[dy][2] * x2[k];
          }
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            float uvals[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) uvals[k] = ax[dx][0] * R[0][k] + ax[dx][1] * R[1][k] + ax[dx][2] * R[2][k];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4* wv = reinterpret_cast<const float4*>(wq + (size_t)(k * 9 + dy * 3 + dx) * OG);
#pragma unroll
              for (int o4 = 0; o4 < OG / 4; ++o4) {
                const float4 ww = wv[o4];
                acc[4 * o4 + 0] = fmaf(uvals[k], ww.x, acc[4 * o4 + 0]);
                acc[4 * o4 + 1] = fmaf(uvals[k], ww.y, acc[4 * o4 + 1]);
                acc[4 * o4 + 2] = fmaf(uvals[k], ww.z, acc[4 * o4 + 2]);
                acc[4 * o4 + 3] = fmaf(uvals[k], ww.w, acc[4 * o4 + 3]);
              }
            }
          }
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out) + (((size_t)n * (cp >> 3) + grp * (OG / 8)) * H2 + v) * (size_t)W2 + u;
#pragma unroll
    for (int o = 0; o < OG; ++o)
      if (grp * OG + o >= C_out) acc[o] = 0.f;
#pragma unroll
    for (int q = 0; q < OG / 8; ++q) dst[(size_t)q * H2 * W2] = pack8(acc + 8 * q);
  }
}

// ---- exact border of the phase-decomposed upconv on a REPLICATE-PADDED low-resolution input ----------------------
// With the one-pixel replicated ring the low-resolution stencil reproduces the bilinear index clamping everywhere; what
// remains is the conv's ZERO padding: the stencil behaves as if the upsampled image continued (clamped) outside
// [0, 2h) x [0, 2w).  Only the outermost high-resolution ring is affected, by exactly the taps that fall outside:
//   out[v, u] -= sum_{(dy, dx) outside} w[o, c, dy, dx] * U~[v + dy - 1, u + dx - 1][c],
// U~[v', u'] = the bilinear value computed from the padded tensor without any clamping (valid for -1 <= v' <= 2h).
// Three taps per edge pixel, five per corner, instead of the nine-tap recomputation of a two-pixel ring.
// One thread = TWO consecutive ring pixels x OG output channels: every weight vector read from shared memory feeds both
// pixels, and the multiply-adds are packed fp32x2 (FFMA2) -- the kernel is instruction bound on the CUDA cores.
template <int OG>
__global__ void __launch_bounds__(UPB_THREADS, 2)
upconv_ringfix_kernel(UpBorderSrc src, int N, int h, int w, const float* __restrict__ bw /* [g][cpad][9][OG] */,
                      int cpad_total, __nv_bfloat16* __restrict__ out, int cp) {
  extern __shared__ __align__(16) float s_wt[];   // [cpad_total][9][OG]
  const int grp = blockIdx.y;
  {
    const float4* g4 = reinterpret_cast<const float4*>(bw + (size_t)grp * cpad_total * 9 * OG);
    float4* s4 = reinterpret_cast<float4*>(s_wt);
    for (int e = threadIdx.x; e < cpad_total * 9 * OG / 4; e += UPB_THREADS) s4[e] = __ldg(g4 + e);
  }
  __syncthreads();
  const int H2 = 2 * h, W2 = 2 * w;
  const int ring = 2 * W2 + 2 * (H2 - 2);        // even
  const int wp = w + 2;                          // padded low-resolution pitch
  const long long total = (long long)N * (ring / 2);
  for (long long t = (long long)blockIdx.x * UPB_THREADS + threadIdx.x; t < total; t += (long long)gridDim.x * UPB_THREADS) {
    const int n = (int)(t / (ring / 2));
    const int r2 = 2 * (int)(t - (long long)n * (ring / 2));
    int pv[2], pu[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = r2 + i;
      if (r < W2) {
        pv[i] = 0;
        pu[i] = r;
      } else if (r < 2 * W2) {
        pv[i] = H2 - 1;
        pu[i] = r - W2;
      } else {
        r -= 2 * W2;
        pv[i] = 1 + (r >> 1);
        pu[i] = (r & 1) ? W2 - 1 : 0;
      }
    }
    float2 acc[2][OG / 2];
#pragma unroll
    for (int o = 0; o < OG / 2; ++o) acc[0][o] = acc[1][o] = make_float2(0.f, 0.f);
    for (int tap = 0; tap < 9; ++tap) {
      bool need[2];
      int r0[2], c0[2];
      float fy0[2], fx0[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int vv = pv[i] + tap / 3 - 1, uu = pu[i] + tap % 3 - 1;
        need[i] = !(vv >= 0 && vv < H2 && uu >= 0 && uu < W2);       // inside taps are already exact
        // U~[vv, uu] from the padded tensor: vv = 2k + a (floor division, k may be -1 or h)
        const int kv = (vv + 2) / 2 - 1, av = (vv + 2) & 1;
        const int ku = (uu + 2) / 2 - 1, au = (uu + 2) & 1;
        r0[i] = kv + av;                          // padded rows r0, r0 + 1
        c0[i] = ku + au;
        fy0[i] = av ? 0.75f : 0.25f;
        fx0[i] = au ? 0.75f : 0.25f;
      }
      if (!need[0] && !need[1]) continue;
      int cpad = 0;
      for (int s = 0; s < src.n_src; ++s) {
        const int bm = src.batch_mod[s];
        const int ns = (src.batch_stride[s] == 0) ? 0 : (bm > 0 ? n % bm : (bm < 0 ? n / (-bm) : n));
        const uint4* xs = src.ptr[s] + (size_t)ns * src.batch_stride[s];
        for (int ch = 0; ch < src.real_chunks[s]; ++ch, cpad += 8) {
          float uv[2][8];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (need[i]) {
              const uint4* xc = xs + (size_t)ch * (h + 2) * wp + (size_t)r0[i] * wp + c0[i];
              float a[8], b[8], c[8], d[8];
              unpack8(__ldg(xc), a);
              unpack8(__ldg(xc + 1), b);
              unpack8(__ldg(xc + wp), c);
              unpack8(__ldg(xc + wp + 1), d);
              const float fy1 = 1.f - fy0[i], fx1 = 1.f - fx0[i];
#pragma unroll
              for (int k = 0; k < 8; ++k)
                uv[i][k] = fy0[i] * (fx0[i] * a[k] + fx1 * b[k]) + fy1 * (fx0[i] * c[k] + fx1 * d[k]);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) uv[i][k] = 0.f;
            }
          }
          const float* wq = s_wt + ((size_t)cpad * 9 + tap) * OG;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4* wv = reinterpret_cast<const float4*>(wq + (size_t)k * 9 * OG);
            const float2 u0 = make_float2(uv[0][k], uv[0][k]), u1 = make_float2(uv[1][k], uv[1][k]);
#pragma unroll
            for (int o4 = 0; o4 < OG / 4; ++o4) {
              const float4 ww = wv[o4];
              const float2 wa = make_float2(ww.x, ww.y), wb = make_float2(ww.z, ww.w);
              acc[0][2 * o4 + 0] = __ffma2_rn(u0, wa, acc[0][2 * o4 + 0]);
              acc[0][2 * o4 + 1] = __ffma2_rn(u0, wb, acc[0][2 * o4 + 1]);
              acc[1][2 * o4 + 0] = __ffma2_rn(u1, wa, acc[1][2 * o4 + 0]);
              acc[1][2 * o4 + 1] = __ffma2_rn(u1, wb, acc[1][2 * o4 + 1]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4* dst =
          reinterpret_cast<uint4*>(out) + (((size_t)n * (cp >> 3) + grp * (OG / 8)) * H2 + pv[i]) * (size_t)W2 + pu[i];
#pragma unroll
      for (int q = 0; q < OG / 8; ++q) {
        float cur[8];
        unpack8(dst[(size_t)q * H2 * W2], cur);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          cur[2 * k] -= acc[i][4 * q + k].x;
          cur[2 * k + 1] -= acc[i][4 * q + k].y;
        }
        dst[(size_t)q * H2 * W2] = pack8(cur);
      }
    }
  }
}

// ---- the same ring correction on the tensor cores (warp-level mma.sync) ------------------------------------------
// The CUDA-core kernel above is instruction bound: 3 taps x C_in x C_out multiply-adds per ring pixel is 2.4 GMAC per
// launch at 208^2 -> 416^2, and every U~ value is interpolated three times (once per neighbouring pixel).  Here one CTA
// owns one image and walks its four edges.  Per edge: (1) the outside line U~ (one hi-res row / column just beyond the
// image, with a one-pixel apron) is interpolated ONCE into shared memory as bf16 [position][channel]; (2) the edge's
// three outside taps are staged as mma B fragments; (3) each warp takes 16 consecutive ring pixels: A = the line at
// positions q + t (t = tap along the edge) read with ldmatrix, K = 3 taps x C_in, N = C_out; (4) the fp32 result is
// subtracted from the tensor-core output in place.  Corner pixels: the row edges take their three row taps (including
// the diagonal one), the column edges skip line positions -1 and L (apron zeroed), and __syncthreads() between the
// edges orders the two read-modify-writes.
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int RFM_THREADS = 256;
struct RfmGeom {
  int cpad;        // padded input channels of all sources (multiple of 8)
  int cpad16;      // ... rounded up to the K step
  int pitch;       // bytes per line position (cpad16 * 2 + 16: conflict-free ldmatrix rows)
  int line_rows;   // positions held per line: 16 * tiles + 2
  int line_bytes;  // line_rows * pitch rounded up to 16
};

// mma B fragments of the three outside taps of every edge, appended to the float32 border weights:
// [edge][kstep][n tile][lane][4] bf16, element (k = t * cpad16 + c, o) = w[o][c][tap(edge, t)]
__global__ void __launch_bounds__(256)
upconv_ringfix_frag_kernel(const float* __restrict__ w, int C_out, int C_in, UpBorderPack ps, int cpad_total, int cpad16,
                           int cp, __nv_bfloat16* __restrict__ frag) {
  const int per_edge = 3 * cpad16 * cp;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 4 * per_edge; idx += gridDim.x * blockDim.x) {
    const int edge = idx / per_edge, rem = idx - edge * per_edge;
    const int o = rem % cp, k = rem / cp;
    const int t = k / cpad16, cpd = k - t * cpad16;
    const int tap = (edge == 0) ? t : (edge == 1) ? 6 + t : (edge == 2) ? 3 * t : 3 * t + 2;
    int c = cpd, ci = -1, base = 0;
    if (cpd < cpad_total) {
      for (int sidx = 0; sidx < ps.n_src; ++sidx) {
        const int padded = (ps.real[sidx] + 7) / 8 * 8;
        if (c < padded) {
          if (c < ps.real[sidx]) ci = base + c;
          break;
        }
        c -= padded;
        base += ps.real[sidx];
      }
    }
    const float v = (ci >= 0 && o < C_out) ? w[((size_t)o * C_in + ci) * 9 + tap] : 0.f;
    const int ks = k >> 4, kk = k & 15, nt_total = cp >> 3;
    const int fl = (o & 7) * 4 + ((kk & 7) >> 1);
    frag[((size_t)((edge * (3 * cpad16 >> 4) + ks) * nt_total + (o >> 3)) * 32 + fl) * 4 + (kk >> 3) * 2 + (kk & 1)] =
        __float2bfloat16_rn(v);
  }
}

template <int NT>   // output-channel tiles of 8 (cp / 8)
__global__ void __launch_bounds__(RFM_THREADS)
upconv_ringfix_mma_kernel(UpBorderSrc src, int N, int h, int w, const uint2* __restrict__ frag, RfmGeom gm,
                          __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char rf_smem[];   // two lines: one edge pair at a time
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H2 = 2 * h, W2 = 2 * w, wp = w + 2;
  const int ksteps_per_tap = gm.cpad16 >> 4, ksteps = 3 * ksteps_per_tap;
  const int chunks16 = gm.cpad16 >> 3;

  for (int pair = 0; pair < 2; ++pair) {      // rows (top, bottom), then columns (left, right)
    const bool row_edge = pair == 0;
    const int L = row_edge ? W2 : H2;
    const int tiles = (L + 15) >> 4;
    // ring pixel (rows g / g + 8 of tile `tile2`) this thread corrects; nullptr beyond the edge
    const int g = lane >> 2, tq = lane & 3;
    auto ring_ptr = [&](int tile2, int half) -> __nv_bfloat162* {
      const int side = tile2 >= tiles ? 1 : 0;
      const int q = ((tile2 - side * tiles) << 4) + g + half * 8;
      if (q >= L) return nullptr;
      const int v = row_edge ? (side ? H2 - 1 : 0) : q;
      const int u = row_edge ? q : (side ? W2 - 1 : 0);
      return reinterpret_cast<__nv_bfloat162*>(out + (((size_t)n * NT * H2 + v) * (size_t)W2 + u) * 8 + 2 * tq);
    };
    constexpr int RB = NT >= 8 ? 2 : 4;          // tiles whose ring pixels are in flight per warp
    __nv_bfloat162 cur[RB][2][NT];
    auto prefetch = [&](int first) {
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int tile2 = first + i * (RFM_THREADS / 32);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const __nv_bfloat162* src = (tile2 < 2 * tiles) ? ring_ptr(tile2, half) : nullptr;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
            if (src != nullptr) cur[i][half][nt] = src[(size_t)nt * H2 * W2 * 4];
        }
      }
    };
    prefetch(warp);
    // (1) the two outside lines: position p <-> hi-res coordinate p - 1 along the edge
    const int items = gm.line_rows * chunks16;
#pragma unroll 4
    for (int item = threadIdx.x; item < 2 * items; item += RFM_THREADS) {
      const int side = item >= items ? 1 : 0;
      const int it = item - side * items;
      const int pos = it / chunks16, chunk = it - pos * chunks16;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      const bool live = pos < L + 2 && chunk * 8 < gm.cpad && (row_edge || (pos != 0 && pos != L + 1));
      if (live) {
        const int q = pos - 1;
        const int vv = row_edge ? (side ? H2 : -1) : q;
        const int uu = row_edge ? q : (side ? W2 : -1);
        // U~[vv, uu] from the padded tensor (same arithmetic as upconv_ringfix_kernel)
        const int kv = (vv + 2) / 2 - 1, av = (vv + 2) & 1;
        const int ku = (uu + 2) / 2 - 1, au = (uu + 2) & 1;
        const int r0 = kv + av, c0 = ku + au;
        const float fy0 = av ? 0.75f : 0.25f, fx0 = au ? 0.75f : 0.25f;
        const float fy1 = 1.f - fy0, fx1 = 1.f - fx0;
        int ch = chunk, s = 0;
        while (ch >= src.real_chunks[s]) ch -= src.real_chunks[s++];
        const int bm = src.batch_mod[s];
        const int ns = (src.batch_stride[s] == 0) ? 0 : (bm > 0 ? n % bm : (bm < 0 ? n / (-bm) : n));
        const uint4* xc = src.ptr[s] + (size_t)ns * src.batch_stride[s] + (size_t)ch * (h + 2) * wp + (size_t)r0 * wp + c0;
        float a[8], b[8], c[8], d[8], u[8];
        unpack8(__ldg(xc), a);
        unpack8(__ldg(xc + 1), b);
        unpack8(__ldg(xc + wp), c);
        unpack8(__ldg(xc + wp + 1), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = fy0 * (fx0 * a[k] + fx1 * b[k]) + fy1 * (fx0 * c[k] + fx1 * d[k]);
        val = pack8(u);
      }
      *reinterpret_cast<uint4*>(rf_smem + (size_t)side * gm.line_bytes + (size_t)pos * gm.pitch + chunk * 16) = val;
    }
    __syncthreads();
    // (3) 16 ring pixels per warp step.  The ring pixels a thread corrects (rows g, g + 8 of its tiles; channels 2 tq,
    // 2 tq + 1 of every 8-channel tile) do not depend on the line: the first RB tiles of every warp were fetched before
    // the interpolation stage above, so their HBM latency is hidden behind it.
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;
    for (int first = warp; first < 2 * tiles; first += RB * (RFM_THREADS / 32)) {
      if (first != warp) prefetch(first);
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int tile2 = first + i * (RFM_THREADS / 32);
        if (tile2 >= 2 * tiles) break;
        const int side = tile2 >= tiles ? 1 : 0;
        const int q0 = (tile2 - side * tiles) << 4;
        const int edge = 2 * pair + side;       // 0 top, 1 bottom, 2 left, 3 right
        const unsigned char* line = rf_smem + (size_t)side * gm.line_bytes;
        const uint2* bf = frag + (size_t)edge * ksteps * NT * 32 + lane;
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ++ks) {
          const int t = ks / ksteps_per_tap, c0 = (ks - t * ksteps_per_tap) << 4;
          uint32_t a[4];
          ldmatrix_x4(smem_u32(line + (size_t)(q0 + lrow + t) * gm.pitch + (c0 + lcol) * 2), a);
          uint2 b[NT];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) b[nt] = __ldg(bf + (size_t)(ks * NT + nt) * 32);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], a, b[nt].x, b[nt].y);
        }
        // (4) out -= correction
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          __nv_bfloat162* dst = ring_ptr(tile2, half);
          if (dst != nullptr) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              float2 f = __bfloat1622float2(cur[i][half][nt]);
              f.x -= acc[nt][half * 2 + 0];
              f.y -= acc[nt][half * 2 + 1];
              dst[(size_t)nt * H2 * W2 * 4] = __floats2bfloat162_rn(f.x, f.y);
            }
          }
        }
      }
    }
    __syncthreads();   // the column pair reuses the lines and rewrites the corner pixels
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
EncodeTiledFn tc_get_encode() {
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
  return (ma == 10 && tc_get_encode() != nullptr) ? 1 : 0;
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

int ynet_tc_pad_replicate(const void* x_c8, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out_c8, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C_pad % 8 == 0 && H >= 1 && W >= 1, "bad shape");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x_c8 && out_c8, "null pointer");
  const long long planes = (long long)N * (C_pad / 8);
  c8_pad_replicate_kernel<<<grid_1d(planes * (long long)(H + 2) * (W + 2)), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x_c8), planes, H, W, reinterpret_cast<uint4*>(out_c8));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_upsample2x(const void* x_c8, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out_c8, void* stream) {
  YNET_CHECK_ARG(N >= 0 && C_pad % 8 == 0 && H >= 1 && W >= 1, "bad shape");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ARG(x_c8 && out_c8, "null pointer");
  const long long planes = (long long)N * (C_pad / 8);
  c8_upsample_kernel<<<grid_1d(planes * (long long)(H + 1) * (W + 1)), 256, 0, as_stream(stream)>>>(
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

int64_t ynet_tc_packed_weight_bytes(int32_t C_out, int32_t n_src, const int32_t* src_channels_pad_host, int32_t ksize) {
  if (C_out <= 0 || n_src <= 0 || n_src > YNET_MAX_SOURCES || !src_channels_pad_host || (ksize != 1 && ksize != 2 && ksize != 3))
    return 0;
  const int n_pad = ceil_div(C_out, 16) * 16;
  long long kb = 0;
  for (int i = 0; i < n_src; ++i) kb += src_channels_pad_host[i] / 16;
  return kb * ksize * ksize * 2 * n_pad * 16;
}

int ynet_tc_pack_weights(const float* weight, int32_t C_out, int32_t n_src, const int32_t* src_channels_host,
                         const int32_t* src_channels_pad_host, int32_t ksize, void* packed, void* stream) {
  YNET_CHECK_ARG(weight && packed && src_channels_host && src_channels_pad_host, "null pointer");
  YNET_CHECK_ARG(C_out > 0 && n_src >= 1 && n_src <= YNET_MAX_SOURCES && (ksize >= 1 && ksize <= 3), "bad shape");
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
  const int taps = ksize * ksize;
  const long long total = (long long)kb * taps * 2 * n_pad * 8;
  tc_pack_weights_kernel<<<grid_1d(total), 256, 0, as_stream(stream)>>>(weight, C_out, cin, n_pad, ps, kb, taps,
                                                                        reinterpret_cast<__nv_bfloat16*>(packed));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

}  // extern "C"

namespace ynet {

struct TcOut {
  void* c8;
  float* f32;
  float4* partial;
  PredFuse* fuse;      // EPI_PRED: predictor weights / partial buffer (tiles_per_cta and slots are filled in by tc_launch)
};

template <int TAPS, int EPI>
static cudaError_t tc_configure() {
  cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<1, TAPS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_conv_kernel<2, TAPS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_conv_kernel<3, TAPS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  return e;
}

template <int TAPS, int EPI>
static void tc_dispatch(int j, unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap* maps, const TcParams& p) {
  if (j == 1)
    tc_conv_kernel<1, TAPS, EPI><<<grid, TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  else if (j == 2)
    tc_conv_kernel<2, TAPS, EPI><<<grid, TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  else
    tc_conv_kernel<3, TAPS, EPI><<<grid, TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
}

// Shared launcher of the tcgen05 conv kernel.  taps = 9 (3x3, padding 1) or 1 (1x1); epi = EPI_*.
static int tc_launch(const char* who, const ynet_tc_src* srcs, int n_src, int N, int H, int W, const void* packed_weight,
                     const float* bias, int C_out, int relu, int C_out_pad, int tune, int taps, int epi, TcOut out,
                     int* grid_out, void* stream, int hl_total_pad = 0, int hl_channel_off = 0) {
  if (!(srcs && packed_weight && (bias || epi == EPI_HILO))) {
    set_error("%s: null pointer", who);
    return YNET_E_INVALID;
  }
  if (!(n_src >= 1 && n_src <= YNET_MAX_SOURCES && N >= 0 && H > 0 && W > 0 && C_out > 0 && C_out_pad % 16 == 0 &&
        C_out_pad >= C_out && C_out_pad <= 256)) {
    set_error("%s: bad shape (C_out_pad must be a multiple of 16, <= 256)", who);
    return YNET_E_INVALID;
  }
  if (reinterpret_cast<uintptr_t>(packed_weight) % 16 != 0) {
    set_error("%s: packed_weight not 16-byte aligned", who);
    return YNET_E_ALIGN;
  }
  if (N == 0) return YNET_OK;
  EncodeTiledFn encode = tc_get_encode();
  if (encode == nullptr) {
    set_error("%s: cuTensorMapEncodeTiled is not available from the driver", who);
    return YNET_E_UNSUPPORTED;
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int halo = (taps == 9) ? 1 : 0;
  // accumulators per tile: independent MMA chains, bounded by double-buffered TMEM (512 columns)
  const int j_cap = tmax(1, tmin(TC_MAX_J, 512 / (2 * C_out_pad)));
  int j = tmin(j_cap, ceil_div(W, 8));
  if (const char* e = getenv("YNET_TC_J")) j = tmin(j_cap, atoi(e));
  if (tune & 0xF) j = tmin(j_cap, tune & 0xF);
  j = tmax(1, j);
  if (epi == EPI_PRED) j = FP_J;
  p.j = j;
  p.bw = 8 * j + 2 * halo;
  const int bh = TC_TH + 2 * halo;
  p.a_bytes = bh * p.bw * TC_KB * 2;
  p.a_lbo = bh * p.bw * 16;
  p.a_sbo = p.bw * 16;
  CUtensorMap maps[YNET_MAX_SOURCES];
  memset(maps, 0, sizeof(maps));
  int kb_total = 0, w_total = 0;
  for (int i = 0; i < n_src; ++i) {
    const int cp = srcs[i].channels_pad;
    if (!(srcs[i].ptr && cp > 0 && cp % 16 == 0) || reinterpret_cast<uintptr_t>(srcs[i].ptr) % 16 != 0) {
      set_error("%s: source %d must be 16-byte aligned with channels_pad a positive multiple of 16", who, i);
      return YNET_E_INVALID;
    }
    const bool bcast = srcs[i].batch_stride == 0;
    const int bmod = srcs[i].batch_mod;
    const int nsrc = bcast ? 1 : (bmod > 0 ? bmod : (bmod < 0 ? ceil_div(N, -bmod) : N));
    // A pixel row of one 8-channel chunk is W*8 contiguous bf16: that is the innermost TMA dimension, so one
    // box row is ONE (8J+2)*16-byte request instead of 8J+2 sixteen-byte ones.
    // chunks_stored < cp / 8: the tensor holds fewer 8-channel planes than the K padding (e.g. 2 waypoint channels =
    // ONE plane inside a 16-channel K block); TMA zero-fills the missing planes without touching HBM
    const int stored = (srcs[i].chunks_stored > 0) ? srcs[i].chunks_stored : cp / 8;
    if (stored > cp / 8) {
      set_error("%s: source %d stores more planes than channels_pad / 8", who, i);
      return YNET_E_INVALID;
    }
    const int po = srcs[i].padded ? 1 : 0;       // tensor dims (H + 2, W + 2): replicate-padded ring
    const cuuint64_t Hs = (cuuint64_t)H + 2 * po, Ws = (cuuint64_t)W + 2 * po;
    const cuuint64_t dims[4] = {Ws * 8, Hs, (cuuint64_t)stored, (cuuint64_t)nsrc};
    const cuuint64_t bs = bcast ? (cuuint64_t)stored * Hs * Ws * 16 : (cuuint64_t)srcs[i].batch_stride * 2;
    if (bs % 16 != 0) {
      set_error("%s: batch stride must be a multiple of 8 elements", who);
      return YNET_E_ALIGN;
    }
    const cuuint64_t strides[3] = {Ws * 16, Hs * Ws * 16, bs};
    const cuuint32_t box[4] = {(cuuint32_t)p.bw * 8, (cuuint32_t)bh, 2, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(srcs[i].ptr), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("%s: cuTensorMapEncodeTiled failed (%d) for source %d (W=%d H=%d C=%d)", who, (int)r, i, W, H, cp);
      return YNET_E_CUDA;
    }
    p.src[i].kblocks = cp / 16;
    p.src[i].bcast = bcast ? 1 : 0;
    p.src[i].batch_mod = srcs[i].batch_mod;
    p.src[i].center = (taps == 9 && srcs[i].center_only) ? 1 : 0;
    const bool quad = taps == 9 && !srcs[i].center_only && srcs[i].tap_mask == YNET_TC_TAPS_QUAD;
    if (srcs[i].tap_mask != 0 && !quad) {
      set_error("%s: source %d: tap_mask must be 0 or YNET_TC_TAPS_QUAD (3x3 convs, not centre-only)", who, i);
      return YNET_E_INVALID;
    }
    if (quad && (epi == EPI_PRED || po)) {
      set_error("%s: source %d: 2x2-neighbourhood sources are not supported by this variant", who, i);
      return YNET_E_UNSUPPORTED;
    }
    p.src[i].off = po;
    if (kb_total + cp / 16 > TC_MAX_KB) {
      set_error("%s: more than %d input K blocks", who, TC_MAX_KB);
      return YNET_E_UNSUPPORTED;
    }
    for (int b = 0; b < cp / 16; ++b) {
      p.wofs[kb_total + b] = w_total;
      w_total += (p.src[i].center ? 1 : (quad ? 4 : taps)) * 2 * C_out_pad * 16;
      if (p.src[i].center) p.center_mask |= 1ull << (kb_total + b);
      if (quad) p.quad_mask |= 1ull << (kb_total + b);
    }
    kb_total += cp / 16;
  }
  p.w_total = w_total;
  for (int i = n_src; i < YNET_MAX_SOURCES; ++i) maps[i] = maps[0];
  p.n_src = n_src;
  p.N = N;
  p.H = H;
  p.W = W;
  p.n_pad = C_out_pad;
  p.c_out = C_out;
  p.relu = relu & 1;
  p.pad_out = (epi == EPI_C8 && (relu & 2)) ? 1 : 0;
  p.with_lo = (epi == EPI_HILO && (relu & 2)) ? 1 : 0;
  p.hl_planes = hl_total_pad > 0 ? 2 * hl_total_pad / 8 : (p.with_lo ? 2 : 1) * (C_out_pad / 8);
  p.hl_off = hl_channel_off / 8;
  p.tiles_x = ceil_div(W, 8 * p.j);
  p.tiles_y = ceil_div(H, TC_TH);
  p.total_tiles = (long long)N * p.tiles_x * p.tiles_y;
  p.kb_total = kb_total;
  p.wpacked = reinterpret_cast<const unsigned char*>(packed_weight);
  p.bias = bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(out.c8);
  p.out_f32 = out.f32;
  p.partial = out.partial;
  p.err = nullptr;
  if (const char* e = getenv("YNET_TC_DBG")) p.dbg = atoi(e);

  const int wblk = taps * 2 * C_out_pad * 16;
  const long long wall = w_total;
  const int budget = 200 * 1024;
  const int tail = (2 * TC_MAX_STAGES + 5) * 8 + 16 + 256 * 4;
  const char* force_stream = getenv("YNET_TC_FORCE_STREAMED");
  p.resident = (wall + 4 * p.a_bytes + tail <= budget) && !(force_stream && force_stream[0] == '1');
  p.wres_bytes = p.resident ? (int)ceil_div<long long>(wall, 1024) * 1024 : 0;
  p.stage_bytes = ceil_div(p.a_bytes + (p.resident ? 0 : wblk), 128) * 128;
  int max_stages = 6;   // measured: depth beyond ~4 does not matter; small rings let two CTAs share an SM
  if (const char* e = getenv("YNET_TC_STAGES")) max_stages = tmax(2, tmin(TC_MAX_STAGES, atoi(e)));
  if ((tune >> 8) & 0xFF) max_stages = tmax(2, tmin(TC_MAX_STAGES, (tune >> 8) & 0xFF));
  p.stages = tmin(max_stages, (budget - p.wres_bytes - tail) / p.stage_bytes);
  if (p.stages < 2) {
    set_error("%s: layer does not fit shared memory (C_out_pad=%d)", who, C_out_pad);
    return YNET_E_UNSUPPORTED;
  }
  int cols = 32;
  while (cols < 2 * p.j * C_out_pad) cols *= 2;
  p.tmem_cols = cols;
  const size_t smem_bytes = (size_t)p.wres_bytes + (size_t)p.stages * p.stage_bytes + tail + 1024;

  static bool configured = false;
  if (!configured) {
    cudaError_t e = tc_configure<9, EPI_C8>();
    if (e == cudaSuccess) e = tc_configure<9, EPI_UP2>();
    if (e == cudaSuccess) e = tc_configure<9, EPI_HILO>();
    if (e == cudaSuccess) e = tc_configure<1, EPI_NCHW_F32>();
    if (e != cudaSuccess) return cuda_fail(e, "tc_launch(cudaFuncSetAttribute)");
    configured = true;
  }
  if (epi == EPI_PRED) {
    if (!p.resident || C_out_pad > 64 || out.fuse == nullptr) {
      set_error("%s: the fused conv + predictor kernel needs resident weights and C_out_pad <= 64", who);
      return YNET_E_UNSUPPORTED;
    }
    const PredPlan pl = pred_plan(N, H, W);
    PredFuse f = *out.fuse;
    f.tiles_per_cta = pl.tiles_per_cta;
    f.slots = pl.slots;
    const int y_bytes = (C_out_pad / 8) * 4096;
    const int kbp = C_out_pad / 16;
    const int tail_f = (2 * TC_MAX_STAGES + 12) * 8 + 16 + 256 * 4;
    p.stages = tmin(10, (200 * 1024 - p.wres_bytes - 2 * y_bytes - kbp * PR_WBLK_BYTES - tail_f) / p.stage_bytes);
    if (p.stages < 2) {
      set_error("%s: layer does not fit shared memory", who);
      return YNET_E_UNSUPPORTED;
    }
    const size_t smem_f = (size_t)p.wres_bytes + (size_t)p.stages * p.stage_bytes + 2 * (size_t)y_bytes +
                          (size_t)kbp * PR_WBLK_BYTES + tail_f + 1024;
    static bool configured_f = false;
    if (!configured_f) {
      cudaError_t e = cudaFuncSetAttribute(tc_conv_pred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "tc_launch(cudaFuncSetAttribute)");
      configured_f = true;
    }
    tc_conv_pred_kernel<<<pl.grid, FP_THREADS, smem_f, as_stream(stream)>>>(maps[0], maps[1], maps[2], maps[3], p, f);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return cuda_fail(le, who);
    return YNET_OK;
  }
  // co-resident CTAs per SM: every CTA brings its own MMA-issuing thread (a single thread needs ~45 cycles of
  // descriptor set-up per small-N MMA), bounded by shared memory, the 512 TMEM columns and 2048 threads
  const int fit = (int)tmin<size_t>(4, tmin<size_t>((size_t)(225 * 1024) / (smem_bytes + 1024), (size_t)(512 / p.tmem_cols)));
  int ctas_per_sm = tmax(1, tmin(fit, 2));
  if (const char* e = getenv("YNET_TC_CTAS_PER_SM")) ctas_per_sm = tmax(1, tmin(fit, atoi(e)));
  if ((tune >> 4) & 0xF) ctas_per_sm = tmax(1, tmin(fit, (tune >> 4) & 0xF));
  long long grid = tmin<long long>(p.total_tiles, (long long)sm_count() * ctas_per_sm);
  if (grid_out != nullptr) {
    if (*grid_out > 0) grid = tmin<long long>(grid, *grid_out);   // caller sized its partial buffer for this many CTAs
    *grid_out = (int)grid;
  }
  cudaStream_t st = as_stream(stream);
  if (taps == 9 && epi == EPI_C8)
    tc_dispatch<9, EPI_C8>(p.j, (unsigned)grid, smem_bytes, st, maps, p);
  else if (taps == 9 && epi == EPI_UP2)
    tc_dispatch<9, EPI_UP2>(p.j, (unsigned)grid, smem_bytes, st, maps, p);
  else if (taps == 9 && epi == EPI_HILO)
    tc_dispatch<9, EPI_HILO>(p.j, (unsigned)grid, smem_bytes, st, maps, p);
  else if (taps == 1 && epi == EPI_NCHW_F32)
    tc_dispatch<1, EPI_NCHW_F32>(p.j, (unsigned)grid, smem_bytes, st, maps, p);
  else {
    set_error("%s: unsupported (taps, epilogue) combination", who);
    return YNET_E_UNSUPPORTED;
  }
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return cuda_fail(le, who);
  return YNET_OK;
}

}  // namespace ynet

extern "C" {

int ynet_tc_conv3x3(const ynet_tc_src* srcs, int32_t n_src, int32_t N, int32_t H, int32_t W, const void* packed_weight,
                    const float* bias, int32_t C_out, int32_t relu, void* out_c8, int32_t C_out_pad, int32_t tune,
                    void* stream) {
  YNET_CHECK_ARG(out_c8 != nullptr || N == 0, "null output");
  YNET_CHECK_ALIGN(out_c8, 16);
  TcOut o{out_c8, nullptr, nullptr, nullptr};
  return tc_launch("ynet_tc_conv3x3", srcs, n_src, N, H, W, packed_weight, bias, C_out, relu, C_out_pad, tune, 9, EPI_C8, o,
                   nullptr, stream);
}

int ynet_tc_conv3x3_pred_softargmax(const ynet_tc_src* srcs, int32_t n_src, int32_t N, int32_t H, int32_t W,
                                    const void* packed_weight, const float* bias, int32_t C_out, int32_t relu,
                                    const void* packed_pred_weight, const float* pred_bias, int32_t C_pred, float* out,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
  YNET_CHECK_ARG(out != nullptr || N == 0, "null output");
  YNET_CHECK_ARG(packed_pred_weight && pred_bias, "null pointer");
  YNET_CHECK_ARG(C_pred > 0 && C_pred <= 32 && C_out > 0 && C_out <= 64, "C_pred <= 32, C_out <= 64");
  if (N == 0) return YNET_OK;
  YNET_CHECK_ALIGN(packed_pred_weight, 16);
  if (workspace == nullptr || workspace_bytes < ynet_tc_conv1x1_softargmax_workspace_bytes(N, C_pred, H, W)) {
    set_error("ynet_tc_conv3x3_pred_softargmax: workspace too small");
    return YNET_E_WORKSPACE;
  }
  YNET_CHECK_ALIGN(workspace, 16);
  const PredPlan pl = pred_plan(N, H, W);
  PredFuse f;
  memset(&f, 0, sizeof(f));
  f.pw = reinterpret_cast<const unsigned char*>(packed_pred_weight);
  f.pbias = pred_bias;
  f.partial = reinterpret_cast<float4*>(workspace);
  f.c_pred = C_pred;
  f.pn_pad = ceil_div(C_pred, 16) * 16;
  cudaStream_t st = as_stream(stream);
  cudaError_t le = pred_partial_init(f.partial, (long long)N * C_pred * pl.slots, st);
  if (le != cudaSuccess) return cuda_fail(le, "ynet_tc_conv3x3_pred_softargmax");
  TcOut o{nullptr, nullptr, nullptr, &f};
  int rc = tc_launch("ynet_tc_conv3x3_pred_softargmax", srcs, n_src, N, H, W, packed_weight, bias, C_out, relu,
                     ceil_div(C_out, 16) * 16, 0, 9, EPI_PRED, o, nullptr, stream);
  if (rc != YNET_OK) return rc;
  le = pred_partial_finalize(f.partial, N * C_pred, pl.slots, out, st);
  if (le != cudaSuccess) return cuda_fail(le, "ynet_tc_conv3x3_pred_softargmax");
  return YNET_OK;
}

int ynet_tc_conv3x3_hilo(const ynet_tc_src* srcs, int32_t n_src, int32_t N, int32_t H, int32_t W, const void* packed_weight,
                         int32_t C_out, void* out_c8, int32_t C_out_pad, int32_t with_lo, int32_t tune, void* stream) {
  YNET_CHECK_ARG(out_c8 != nullptr || N == 0, "null output");
  YNET_CHECK_ALIGN(out_c8, 16);
  TcOut o{out_c8, nullptr, nullptr, nullptr};
  return tc_launch("ynet_tc_conv3x3_hilo", srcs, n_src, N, H, W, packed_weight, nullptr, C_out, with_lo ? 2 : 0, C_out_pad,
                   tune, 9, EPI_HILO, o, nullptr, stream);
}

// A full layer of the split-bf16 engine (split_tc.cu): conv over the sources (for an activation x: [x_hi | x_lo] against
// [W_hi | W_hi], then the hi planes again against W_lo), + bias, ReLU, fp32 accumulator split into hi | lo planes.
int ynet_tc_conv3x3_split(const ynet_tc_src* srcs, int32_t n_src, int32_t N, int32_t H, int32_t W, const void* packed_weight,
                          const float* bias, int32_t C_out, int32_t relu, void* out_split, int32_t C_out_pad,
                          int32_t out_total_pad, int32_t out_channel_off, int32_t tune, void* stream) {
  YNET_CHECK_ARG(out_split != nullptr || N == 0, "null output");
  YNET_CHECK_ARG(bias != nullptr, "null bias");
  if (out_total_pad == 0) out_total_pad = C_out_pad;
  YNET_CHECK_ARG(out_total_pad % 16 == 0 && out_channel_off % 16 == 0 && out_channel_off >= 0 &&
                     out_channel_off + C_out_pad <= out_total_pad,
                 "output slice outside the activation (multiples of 16)");
  YNET_CHECK_ALIGN(out_split, 16);
  TcOut o{out_split, nullptr, nullptr, nullptr};
  return tc_launch("ynet_tc_conv3x3_split", srcs, n_src, N, H, W, packed_weight, bias, C_out, (relu ? 1 : 0) | 2, C_out_pad,
                   tune, 9, EPI_HILO, o, nullptr, stream, out_total_pad, out_channel_off);
}

int ynet_tc_upconv_phase_weights(const float* weight, const float* bias, int32_t C_out, int32_t C_in, float* w_eff,
                                 float* bias_eff, void* stream) {
  YNET_CHECK_ARG(weight && w_eff && bias_eff, "null pointer");
  YNET_CHECK_ARG(C_out > 0 && C_in > 0, "bad shape");
  const int cp = ceil_div(C_out, 16) * 16;
  upconv_phase_weights_kernel<<<grid_1d((long long)4 * cp * C_in * 9), 256, 0, as_stream(stream)>>>(
      weight, bias, C_out, C_in, cp, w_eff, bias_eff);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

static inline int upb_og(int cp) { return cp >= 32 ? 32 : 16; }

int64_t ynet_tc_upconv_border_weight_bytes(int32_t C_out, int32_t n_src, const int32_t* src_channels_host) {
  if (C_out <= 0 || n_src <= 0 || n_src > YNET_MAX_SOURCES || !src_channels_host) return 0;
  const int cp = ceil_div(C_out, 16) * 16;
  long long cpad = 0;
  for (int i = 0; i < n_src; ++i) cpad += ceil_div(src_channels_host[i], 8) * 8;
  // float32 [g][cpad][9][og] image of the CUDA-core border kernels, then the bf16 mma fragments of the ring fix
  const long long cpad16 = (cpad + 15) / 16 * 16;
  return cpad * 9 * cp * (long long)sizeof(float) + 4 * 3 * cpad16 * cp * 2;
}

int ynet_tc_upconv_border_weights(const float* weight, int32_t C_out, int32_t n_src, const int32_t* src_channels_host,
                                  float* out, void* stream) {
  YNET_CHECK_ARG(weight && out && src_channels_host, "null pointer");
  YNET_CHECK_ARG(C_out > 0 && n_src >= 1 && n_src <= YNET_MAX_SOURCES, "bad shape");
  UpBorderPack ps;
  memset(&ps, 0, sizeof(ps));
  ps.n_src = n_src;
  int cin = 0, cpad = 0;
  for (int i = 0; i < n_src; ++i) {
    YNET_CHECK_ARG(src_channels_host[i] > 0, "bad source channels");
    ps.real[i] = src_channels_host[i];
    cin += src_channels_host[i];
    cpad += ceil_div(src_channels_host[i], 8) * 8;
  }
  const int cp = ceil_div(C_out, 16) * 16;
  const int og = upb_og(cp);
  upconv_border_weights_kernel<<<grid_1d((long long)cpad * 9 * cp), 256, 0, as_stream(stream)>>>(
      weight, C_out, cin, ps, cpad, og, cp / og, out);
  YNET_LAUNCH_CHECK();
  const int cpad16 = ceil_div(cpad, 16) * 16;
  upconv_ringfix_frag_kernel<<<grid_1d((long long)12 * cpad16 * cp), 256, 0, as_stream(stream)>>>(
      weight, C_out, cin, ps, cpad, cpad16, cp, reinterpret_cast<__nv_bfloat16*>(out + (size_t)cpad * 9 * cp));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_upconv3x3(const ynet_tc_src* srcs, const int32_t* src_channels_host, int32_t n_src, int32_t N, int32_t h,
                      int32_t w, const void* packed_phase_weight, const float* bias_eff, const float* border_weight,
                      const float* bias, int32_t C_out, int32_t relu, void* out_c8, int32_t tune, void* stream) {
  YNET_CHECK_ARG(out_c8 != nullptr || N == 0, "null output");
  YNET_CHECK_ALIGN(out_c8, 16);
  YNET_CHECK_ARG(srcs && src_channels_host && border_weight, "null pointer");
  YNET_CHECK_ALIGN(border_weight, 16);
  YNET_CHECK_ARG(C_out > 0 && C_out <= 64, "C_out must be <= 64 (4 * C_out_pad <= 256 accumulator columns)");
  if (relu != 0) {
    set_error("ynet_tc_upconv3x3: ReLU is not supported (ynet.py:464 applies none after upsample_conv)");
    return YNET_E_UNSUPPORTED;
  }
  const int cp = ceil_div(C_out, 16) * 16;
  TcOut o{out_c8, nullptr, nullptr, nullptr};
  int rc = tc_launch("ynet_tc_upconv3x3", srcs, n_src, N, h, w, packed_phase_weight, bias_eff, 4 * cp, 0, 4 * cp, tune, 9,
                     EPI_UP2, o, nullptr, stream);
  if (rc != YNET_OK || N == 0) return rc;
  if (const char* e = getenv("YNET_UPCONV_SKIP_BORDER"))   // profiling aid: time the tensor-core part alone
    if (e[0] == '1') return YNET_OK;
  UpBorderSrc bs;
  memset(&bs, 0, sizeof(bs));
  bs.n_src = n_src;
  int cpad = 0;
  for (int i = 0; i < n_src; ++i) {
    YNET_CHECK_ARG(src_channels_host[i] > 0 && src_channels_host[i] <= srcs[i].channels_pad, "bad source channels");
    bs.ptr[i] = reinterpret_cast<const uint4*>(srcs[i].ptr);
    bs.batch_stride[i] = srcs[i].batch_stride / 8;
    bs.real_chunks[i] = ceil_div(src_channels_host[i], 8);
    bs.batch_mod[i] = srcs[i].batch_mod;
    cpad += bs.real_chunks[i] * 8;
  }
  const int og = upb_og(cp);
  const size_t smem = (size_t)cpad * 9 * og * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("ynet_tc_upconv3x3: border weights (%d channels x 9 x %d) exceed shared memory", cpad, og);
    return YNET_E_UNSUPPORTED;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(upconv_border_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(upconv_border_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(upconv_ringfix_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(upconv_ringfix_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "ynet_tc_upconv3x3(cudaFuncSetAttribute)");
    configured = true;
  }
  int n_padded = 0;
  for (int i = 0; i < n_src; ++i) n_padded += srcs[i].padded ? 1 : 0;
  YNET_CHECK_ARG(n_padded == 0 || n_padded == n_src, "either all or none of the sources carry the replicate-padded ring");
  if (n_padded == n_src) {
    // replicate-padded inputs: the tensor-core result is exact except for the zero padding of the outermost ring
    {
      const char* env = getenv("YNET_RINGFIX_MMA");     // "0": the CUDA-core kernel (kept as the cross-check)
      const bool use_mma = !(env != nullptr && env[0] == '0');
      RfmGeom gm;
      gm.cpad = cpad;
      gm.cpad16 = ceil_div(cpad, 16) * 16;
      gm.pitch = gm.cpad16 * 2 + 16;
      gm.line_rows = 16 * ceil_div(2 * tmax(h, w), 16) + 2;
      gm.line_bytes = ceil_div(gm.line_rows * gm.pitch, 16) * 16;
      const size_t rf_smem = (size_t)2 * gm.line_bytes;
      const int nt = cp / 8;
      if (use_mma && rf_smem <= 200 * 1024 && (nt == 2 || nt == 4 || nt == 8)) {
        static bool rf_configured = false;
        if (!rf_configured) {
          cudaError_t e = cudaFuncSetAttribute(upconv_ringfix_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          if (e == cudaSuccess)
            e = cudaFuncSetAttribute(upconv_ringfix_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          if (e == cudaSuccess)
            e = cudaFuncSetAttribute(upconv_ringfix_mma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          if (e != cudaSuccess) return cuda_fail(e, "ynet_tc_upconv3x3(cudaFuncSetAttribute, ring fix)");
          rf_configured = true;
        }
        __nv_bfloat16* outm = reinterpret_cast<__nv_bfloat16*>(out_c8);
        const uint2* frag = reinterpret_cast<const uint2*>(border_weight + (size_t)cpad * 9 * cp);
        if (nt == 2)
          upconv_ringfix_mma_kernel<2><<<N, RFM_THREADS, rf_smem, as_stream(stream)>>>(bs, N, h, w, frag, gm, outm);
        else if (nt == 4)
          upconv_ringfix_mma_kernel<4><<<N, RFM_THREADS, rf_smem, as_stream(stream)>>>(bs, N, h, w, frag, gm, outm);
        else
          upconv_ringfix_mma_kernel<8><<<N, RFM_THREADS, rf_smem, as_stream(stream)>>>(bs, N, h, w, frag, gm, outm);
        YNET_LAUNCH_CHECK();
        return YNET_OK;
      }
    }
    const int ring_hi = 2 * (2 * w) + 2 * (2 * h - 2);
    const long long total_hi = (long long)N * (ring_hi / 2);        // two ring pixels per thread
    const int groups_hi = cp / og;
    const int per_sm_hi = smem > 100 * 1024 ? 1 : (smem > 64 * 1024 ? 2 : 3);
    const unsigned gxh = (unsigned)tmax<long long>(
        1, tmin<long long>(ceil_div<long long>(total_hi, UPB_THREADS), (long long)ceil_div(per_sm_hi * sm_count(), groups_hi)));
    dim3 gridh(gxh, groups_hi);
    __nv_bfloat16* outh = reinterpret_cast<__nv_bfloat16*>(out_c8);
    if (og == 16)
      upconv_ringfix_kernel<16><<<gridh, UPB_THREADS, smem, as_stream(stream)>>>(bs, N, h, w, border_weight, cpad, outh, cp);
    else
      upconv_ringfix_kernel<32><<<gridh, UPB_THREADS, smem, as_stream(stream)>>>(bs, N, h, w, border_weight, cpad, outh, cp);
    YNET_LAUNCH_CHECK();
    return YNET_OK;
  }
  const int ring = (h >= 2 && w >= 2) ? 2 * w + 2 * (h - 2) : h * w;
  const long long total = (long long)N * ring * 4;
  const int groups = cp / og;
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 64 * 1024 ? 2 : 3);
  const unsigned gx = (unsigned)tmax<long long>(
      1, tmin<long long>(ceil_div<long long>(total, UPB_THREADS), (long long)ceil_div(per_sm * sm_count(), groups)));
  dim3 grid(gx, groups);
  __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(out_c8);
  if (og == 16)
    upconv_border_kernel<16><<<grid, UPB_THREADS, smem, as_stream(stream)>>>(bs, N, h, w, border_weight, bias, C_out, cpad,
                                                                             outp, cp);
  else
    upconv_border_kernel<32><<<grid, UPB_THREADS, smem, as_stream(stream)>>>(bs, N, h, w, border_weight, bias, C_out, cpad,
                                                                             outp, cp);
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_conv1x1_f32(const ynet_tc_src* srcs, int32_t n_src, int32_t N, int32_t H, int32_t W, const void* packed_weight,
                        const float* bias, int32_t C_out, float* out, int32_t tune, void* stream) {
  YNET_CHECK_ARG(out != nullptr || N == 0, "null output");
  TcOut o{nullptr, out, nullptr, nullptr};
  return tc_launch("ynet_tc_conv1x1_f32", srcs, n_src, N, H, W, packed_weight, bias, C_out, 0, ceil_div(C_out, 16) * 16, tune,
                   1, EPI_NCHW_F32, o, nullptr, stream);
}

}  // extern "C"
