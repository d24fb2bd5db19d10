// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"

namespace ynet {

constexpr unsigned TC_SPIN_LIMIT = 4u * 1000u * 1000u;             // bounded waits: trap instead of hanging
constexpr long long TC_WAIT_CYCLES = 8LL * 1000 * 1000 * 1000;    // ~4 s of SM clock

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
// mbarrier.try_wait WITHOUT a suspend-time hint: the instruction (SASS SYNCS...TRYWAIT) suspends the warp in hardware until
// the phase completes or a system time limit expires, so a waiting warp issues nothing.  Round 1 passed a 2 us hint: ptxas
// then emits TRYWAIT + NANOSLEEP.SYNCS + PHASECHK + a spin counter, and NANOSLEEP.SYNCS returns on ANY barrier activity of
// the CTA -- with ~35 barriers toggling per row the waiting warps of rowconv_tc.cu re-issued that 7-instruction loop 20-40
// times per wait and took more than half of the SM's issue slots away from the single-thread TMA / MMA roles
// (profiles/ncu_r02_rowconv_fused_v3.md).  The spin bound stays: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
  uint32_t done = 0;
  unsigned spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    // a protocol bug must not hang the GPU: give up after TC_SPIN_LIMIT failed waits or ~4 s, whichever comes first
    if (spins == 0) t0 = clock64();
    if (++spins > TC_SPIN_LIMIT || clock64() - t0 > TC_WAIT_CYCLES) {
      if (err) atomicExch(err, 1);
      __trap();
    }
  }
}
// Non-blocking probe of a phase (1 = complete).  The MMA issuers probe the NEXT row's barriers before they issue the current
// row's MMAs, so the ~100-cycle latency of the probe overlaps the issue work instead of stalling the scalar thread.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// Whole-warp wait with uniform control flow afterwards: one lane spins, the warp reconverges on __syncwarp().  Used by
// the MMA-issuing warp so that ptxas keeps the descriptor arithmetic on the uniform datapath.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity, nullptr);
  __syncwarp();
}
// One elected lane of a converged warp.  Guard single-thread roles (TMA producer, MMA issuer) with THIS, not with
// `lane == 0`: ptxas recognises the elect.sync-guarded region as single-threaded and keeps descriptors, barrier addresses
// and the tcgen05 / TMA instructions on the uniform datapath (2-3 instructions per MMA); behind `lane == 0` every
// tcgen05.mma is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (8-12 instructions), and a single issuing thread
// then takes ~100 cycles per MMA (measured: the issue thread, not the tensor pipe, bounded rowconv_tc.cu at first).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// Variants for a warp that runs the issue loop with all 32 lanes (uniform control flow) and elects one lane per
// instruction.  Measured on B200: good for a loop with a few large MMAs per tile (pred_tc.cu: 0.90 -> 0.63 ms), bad
// for the conv kernels' 36+ small MMAs per tile (2.6 -> 11.6 ms): those issue from `if (lane == 0)`.
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

// two loads in flight before one wait (the epilogue pairs the two horizontal phases of an upconv pixel)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// The wait names the destination registers as in/out operands: no consumer of v can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
// 32-byte global store (sm_100: STG.256); the address must be 32-byte aligned
__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t (&o)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
               "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_get_encode();   // cuTensorMapEncodeTiled through the runtime's driver entry point (conv_tc.cu)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- predictor + soft-argmax building blocks shared by pred_tc.cu and the fused conv + predictor kernel ----------
constexpr int PR_TH = 16;            // tile rows
constexpr int PR_J = 2;              // 8-pixel column blocks per tile: tile = 16 x 16 pixels = ONE N = 256 accumulator
constexpr int PR_TW = 8 * PR_J;
constexpr int PR_EPI_WARPS = 16;           // one tile row (16 pixels) per warp
constexpr int PR_WBLK_BYTES = 2 * 128 * 16;              // weights of one K block: [chunk][128 rows][8 ch]
constexpr int PR_MAX_KB = 8;
constexpr float PR_NEG = -3.402823466e+38f;

struct PredPlan {
  int tiles_x, tiles_y, grid, slots;
  long long total_tiles, tiles_per_cta;
};
PredPlan pred_plan(int N, int H, int W);
cudaError_t pred_partial_init(float4* part, long long n_part, cudaStream_t st);
cudaError_t pred_partial_finalize(const float4* part, int rows, int slots, float* out, cudaStream_t st);

// Online soft-argmax state of ONE channel (softargmax.py:67-81): running max, sum e, sum e x, sum e y.
struct SoftState {
  float m, s, sx, sy;
};

// One tile row of 16 logits (raw accumulators a[i], logit = a[i] + bias) at pixels (x0 + i, y): the lean interior path
// (~5 instructions per pixel: FFMA, MUFU.EX2 and three accumulations on four independent chains).
__device__ __forceinline__ void softargmax_row16(SoftState& st, const uint32_t (&v)[16], float bias, int x0, int y) {
  constexpr float LOG2E = 1.4426950408889634f;
  float mx = __uint_as_float(v[0]);
#pragma unroll
  for (int i = 1; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
  mx += bias;
  if (mx > st.m) {     // rare after the first tiles of an image
    const float sc = (st.m == PR_NEG) ? 0.f : __expf(st.m - mx);
    st.s *= sc;
    st.sx *= sc;
    st.sy *= sc;
    st.m = mx;
  }
  const float k = (bias - st.m) * LOG2E;       // e = 2^(a log2e + (bias - m) log2e) = exp(a + bias - m)
  float ex[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ex[i] = ex2_approx(fmaf(__uint_as_float(v[i]), LOG2E, k));
  float s0 = ex[0] + ex[1], s1 = ex[4] + ex[5], s2 = ex[8] + ex[9], s3 = ex[12] + ex[13];
  float x0s = ex[1], x1s = ex[4] * 4.f, x2s = ex[8] * 8.f, x3s = ex[12] * 12.f;
  s0 += ex[2];  s1 += ex[6];  s2 += ex[10];  s3 += ex[14];
  x0s = fmaf(ex[2], 2.f, x0s);   x1s = fmaf(ex[5], 5.f, x1s);   x2s = fmaf(ex[9], 9.f, x2s);   x3s = fmaf(ex[13], 13.f, x3s);
  s0 += ex[3];  s1 += ex[7];  s2 += ex[11];  s3 += ex[15];
  x0s = fmaf(ex[3], 3.f, x0s);   x1s = fmaf(ex[6], 6.f, x1s);   x2s = fmaf(ex[10], 10.f, x2s); x3s = fmaf(ex[14], 14.f, x3s);
  x1s = fmaf(ex[7], 7.f, x1s);   x2s = fmaf(ex[11], 11.f, x2s); x3s = fmaf(ex[15], 15.f, x3s);
  const float se = (s0 + s1) + (s2 + s3);
  const float sxl = (x0s + x1s) + (x2s + x3s);
  st.s += se;
  st.sx += fmaf((float)x0, se, sxl);
  st.sy = fmaf((float)y, se, st.sy);
}

// The same for a row that crosses the right border (pixels x0 + i >= W are masked): the masked accumulators are replaced
// by a huge negative value (their exponential is exactly 0) and the lean path does the rest.
__device__ __forceinline__ void softargmax_row16_masked(SoftState& st, const uint32_t (&v)[16], float bias, int x0, int y,
                                                        int W) {
  if (x0 >= W) return;
  uint32_t vm[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) vm[i] = (x0 + i < W) ? v[i] : __float_as_uint(-1.0e30f);
  softargmax_row16(st, vm, bias, x0, y);
}

// Predictor weights [kb][2][n_pad][8] bf16 -> shared memory [kb][2][128][8] with the (<= 32) rows replicated into the
// four 32-lane quadrants of the M = 128 operand (row r carries channel r % 32).
__device__ __forceinline__ void pred_stage_weights(unsigned char* s_w, const unsigned char* wpacked, int kblocks, int n_pad,
                                                   int tid, int nthreads) {
  const uint4* g = reinterpret_cast<const uint4*>(wpacked);
  uint4* s = reinterpret_cast<uint4*>(s_w);
  const int total = kblocks * 2 * 128;
  for (int e = tid; e < total; e += nthreads) {
    const int r = e & 127, kc = e >> 7;          // kc = kb * 2 + chunk
    const int ch = r & 31;
    s[e] = (ch < n_pad) ? __ldg(g + (size_t)kc * n_pad + ch) : make_uint4(0, 0, 0, 0);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
}

}  // namespace ynet
