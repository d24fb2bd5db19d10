// a7/a8 tail: ROW-MARCHING, kh-STACKED 3x3 convolution for the C_out <= 32 layers of the decoders' full-resolution
// levels (ynet.py:466-468: decoder.4.2, decoder.3.2), optionally fused with the 1x1 predictor and SoftArgmax2D
// (ynet.py:469 + 582-583, softargmax.py:55-81) so that neither the conv output nor the logits ever reach HBM.
//
// Why another conv kernel.  conv_tc.cu computes a 16 x 8-pixel tile with nine M = 128 x N = C_out MMAs per 16-channel K
// block.  With C_out = 32 every MMA is 16 cycles of math behind 5 KB of shared-memory operand fetch (4 KB of pixels,
// 1 KB of weights): the operand fetch caps the tensor pipe at 40 % (profiles/ncu_full_r01d_tail416.md).  Useful flops
// per fetched pixel byte = N, so N has to grow.  Here the three kernel ROWS are stacked along N:
//
//   D[x, (kh, co)] += sum_ci  in[Y, x + kw - 1, ci] * W[co, ci, kh, kw]          (one MMA per (K block, kw), N = 96)
//
// for ONE input row Y of a 128-pixel strip.  Column group kh of that product belongs to OUTPUT row Y + 1 - kh, and the
// accumulators of consecutive output rows are laid out as a ring of 32-column TMEM slots in DESCENDING row order, so
// the three groups of one N = 96 MMA land on the slots of rows Y + 1, Y, Y - 1 directly: the tensor core does the
// row-shifted accumulation, the epilogue of an output row is a plain 32-column read once input row Y + 1 has been
// multiplied.  Per K block and row: 3 MMAs of 48 math cycles behind 7 KB of operands (56 cycles) instead of 9 MMAs of
// 16 cycles behind 5 KB (40 cycles each): 168 instead of 360 cycles, all columns useful.
// Slots are zeroed by the epilogue after it has drained them (tcgen05.st), so every MMA accumulates; where the slot
// triple wraps around the ring (2 rows in 8) or touches the image's first / last row the MMA is split / narrowed.
//
// Work item = (image, 126-pixel column strip); a CTA marches down all H rows of its strip.  An input row arrives by ONE
// TMA: the C8 planes are described to the TMA unit as 8-byte elements (2 per pixel), so a box line may be 256 elements
// = 128 pixels long and start at ANY pixel; box {128 px, 1 row, all chunks} at x = xs0 - 1 lands as [chunk][128 px][8 ch]
// -- the K-major operand layout with the kw taps as 16-byte start offsets, zero-filled outside the image (= the conv's
// padding).  (A first version viewed the row as 16-pixel segments, 40 box lines of 256 B per row: the TMA unit's
// per-line rate, ~25 cycles, then bounded the kernel at 1 000 cycles per row.)  Lanes 126 / 127 of the M = 128 MMA read
// past the window and are discarded.
//
// Fused tail (FUSE = true): the conv epilogue writes the bf16 row to shared memory as the N = 128-pixel operand of
// the predictor MMA (weights = M operand, replicated into the four lane quadrants as in pred_tc.cu), whose transposed
// accumulator (TMEM lane = channel, column = pixel) is reduced by eight soft-argmax warps straight from tcgen05.ld.
// Both epilogues run as two warp sets that alternate rows, so the per-row latency chain (barrier wake-up, tcgen05.ld,
// tcgen05.st) of one row overlaps the next row's.
//
// Round 2, second half (further down in this file):
//  * WP kernels (ynet_tc_rowconv3x3_wp): the waypoint channels of the trajectory decoder's input are windows of the
//    distance template, so their K block is loaded by TMA from bf16 planes of the TEMPLATE (L2-resident) instead of
//    per-image planes in HBM; edge strips are aligned to 8 pixels so that the conv's zero padding is a TMA load of zeros.
//  * tc_rowconv2_kernel (ynet_tc_rowconv2_wp*): decoder.i.0 and decoder.i.2 in one launch -- conv A's epilogue writes its
//    row into a shared-memory ring that conv B marches over with the same issuer code (templated on the TMEM ring size);
//    the hoisted partial sums arrive through their own TMA row ring.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace ynet {

constexpr int RC_M = 128;                       // MMA M = pixels of a window row
constexpr int RC_VW = 126;                      // valid output pixels per strip row (window = [xs0 - 1, xs0 + 127))
constexpr int RC_CHUNK_BYTES = RC_M * 16;       // one 8-channel chunk of a window row
constexpr int RC_CO = 32;                       // output channels = columns per TMEM slot
constexpr int RC_NS = 8;                        // TMEM slot ring (256 columns)
constexpr int RC_STAGES_PLAIN = 8;              // input-row ring: two co-resident CTAs per SM keep 16 rows (128 KB) in flight
constexpr int RC_STAGES_FUSED = 16;             // one CTA per SM: with 8 rows in flight the fused kernel was bound by the TMA
                                                // round trip (~2 us under load / 7 free stages = 580 cycles per row, even with
                                                // all math stripped: YNET_RC_DBG=7); 16 rows = 128 KB in flight per SM
constexpr int RC_NY = 4;                        // conv-output row ring (fused tail)
constexpr int RC_MAX_KB = 4;                    // <= 64 input channels
constexpr int RC_WBLK = 2 * 96 * 16;            // weights of one (K block, kw): [2 chunks][96 = (kh, co)][8 ch] bf16
constexpr int RC_YROW = 4 * RC_M * 16;          // one conv-output row: [4 chunks][128 px][8 ch] bf16
constexpr int RC_EPI_WARPS = 8;                 // two sets of four (one warp per TMEM lane quadrant), alternating rows
constexpr int RC_SOFT_WARPS = 16;               // two sets x 4 lane quadrants x 2 sixteen-pixel halves
// warp 0: TMA; warp 1: conv MMA issuer.  (rc_conv_issuer can split the rows between two issuing warps, ISS = 2: measured on
// B200 it buys nothing once the barrier probes are hoisted -- 1.567 vs 1.570 ms per 320 images -- and the fp32 accumulation
// order of a slot then depends on how the two threads' MMAs interleave, i.e. results change in the last bit from run to
// run.  One issuer keeps the kernel deterministic.)
constexpr int RC_THREADS_PLAIN = 32 * (2 + RC_EPI_WARPS);        // TMA, MMA, epilogue warps
constexpr int RC_THREADS_FUSED = 32 * (2 + RC_EPI_WARPS + 1 + RC_SOFT_WARPS);   // TMA, MMA, epilogue, predictor MMA, soft-argmax
constexpr uint32_t RC_PACC = 256;               // first TMEM column of the two predictor accumulators (2 x 128).  (Four 64-pixel
                                                // accumulators were tried to shorten the soft-argmax -> MMA hand-off: no gain.)

struct RcParams {
  int N, H, W, kb, chunks, strips;
  long long items;
  int relu, out_chunks, pad_out;
  int row_bytes;               // bytes of one row buffer = chunks * 2 KB
  int row_tx;                  // bytes the TMA transactions of one row deliver (= stored chunks * 2 KB)
  int n_src;                   // conv sources (<= 3): their chunks are laid side by side in the row buffer (torch.cat)
  int src_stored[3];           // 8-channel planes source s really holds (TMA box depth)
  int src_coff[3];             // first chunk slot of source s in the row buffer
  int src_bcast[3], src_mod[3];
  int zero_fill;               // some chunk slots are never written by TMA (K padding): zero the ring once
  // waypoint source (ynet_tc_rowconv3x3_wp): the last K block is get_patch (image_utils.py:40-63) [+ one 2x2 AvgPool,
  // evaluate.py:255-257] of the distance template, loaded by TMA straight from the template's bf16 C8 planes (L2-resident)
  // instead of being rasterised to HBM per image and read back; channel c occupies chunk wp_coff + c (K index 8 c)
  const float* wp_coords;      // (N * wp_nch, 2) = (x, y) at full resolution
  int wp_nch, wp_level, wp_th, wp_tw, wp_coff;
  const uint4* part;           // hoisted partial sums (goal-loop hoisting, ynet_tc_conv3x3_hilo): bf16 C8, 4 chunks hi
  long long part_bs;           //   [+ 4 chunks lo], added by the epilogue before the activation; batch stride in uint4
  int part_mod, part_chunks;
  const unsigned char* w96;    // [kb][kw][2][96][8] bf16
  const float* bias;           // 32 floats (zero beyond C_out)
  __nv_bfloat16* out;          // plain: C8 planes (N, out_chunks, H + 2 pad, W + 2 pad, 8)
  const unsigned char* pw;     // fused: predictor weights [kbp][2][pn_pad][8] bf16 (ynet_tc_pack_weights, ksize 1)
  const float* pbias;
  float4* partial;             // (m, s, sx, sy) [(n * c_pred + c) * slots + strip * RC_SOFT_WARPS + warp]
  int c_pred, pn_pad, slots;
  int kbp;                     // predictor K blocks = ceil(C_out / 16) (1 or 2)
  int dbg;                     // profiling aid (YNET_RC_DBG): 1 = soft-argmax warps skip the math, 2 = conv epilogue skips the
                               // conversion + shared-memory stores, 4 = one conv MMA per row instead of 3 * kb (results are garbage)
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 columns <- the bias (the drained accumulator slot is handed back holding the bias: every MMA
// accumulates, and the epilogue needs no add)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
      "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// bf16x2 = (lo, hi) with optional ReLU folded into the conversion
template <bool RELU>
__device__ __forceinline__ uint32_t pack_bf16_act(uint32_t lo, uint32_t hi) {
  uint32_t d;
  if (RELU)
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return d;
}

template <int NS = RC_NS>
__device__ __forceinline__ int rc_slot(int row) { return NS - 1 - (row & (NS - 1)); }

// ---- single-thread issue paths with compile-time ring positions -----------------------------------------------------------
// The issuing thread is a scalar instruction stream (~8 cycles per dependent instruction): at ~150 instructions per row it,
// not the tensor pipe, bounded the kernel (profiles/ncu_r02_rowconv_fused_v3.md).  Rows repeat with period 8 (slot ring =
// stage ring = 8), so an interior row's barrier addresses, TMEM columns and descriptor offsets are immediates.
constexpr uint32_t RC_A_HI = (uint32_t)(128 >> 4) | (1u << 14);                  // SBO = 128 B | descriptor version
constexpr uint32_t RC_B_HI = (uint32_t)(128 >> 4) | (1u << 14);
constexpr uint32_t RC_IDESC0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);

// kernel rows KH0 .. KH0 + LEN - 1 of one input row -> LEN consecutive slots starting at SLOT
template <int KH0, int LEN, int SLOT, int KB>
__device__ __forceinline__ void rc_run(uint32_t tmem_base, uint32_t a_lo, uint32_t w_lo0, int kbn) {
  constexpr uint32_t idesc = RC_IDESC0 | ((uint32_t)(LEN * (RC_CO >> 3)) << 17);
  const uint32_t d_tmem = tmem_base + (uint32_t)(SLOT * RC_CO);
  const uint32_t b_lo0 = w_lo0 + (uint32_t)(KH0 * RC_CO);
  if (KB > 0) {
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
        tc_mma_bf16(d_tmem, ((uint64_t)RC_A_HI << 32) | (a_lo + (uint32_t)(kb * 2 * (RC_CHUNK_BYTES >> 4) + kw)),
                    ((uint64_t)RC_B_HI << 32) | (b_lo0 + (uint32_t)((kb * 3 + kw) * (RC_WBLK >> 4))), idesc, 1u);
  } else {
    uint32_t a = a_lo, b = b_lo0;
    for (int kb = 0; kb < kbn; ++kb) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
        tc_mma_bf16(d_tmem, ((uint64_t)RC_A_HI << 32) | (a + (uint32_t)kw),
                    ((uint64_t)RC_B_HI << 32) | (b + (uint32_t)(kw * (RC_WBLK >> 4))), idesc, 1u);
      a += 2 * (RC_CHUNK_BYTES >> 4);
      b += 3 * (RC_WBLK >> 4);
    }
  }
}

// Strip geometry of the WP kernels (ynet_tc_rowconv3x3_wp).  The waypoint K block comes from the TEMPLATE's planes, which
// hold values where the conv's zero padding must be (x = -1, x = W).  TMA destinations are 128-byte (8-pixel) aligned,
// so the first strip's window starts at x = -8 and the last strip's at x = W - 120: pixels 0-7 / 120-127 of those
// windows lie outside the image and are written by a TMA load of zeros, the other 120 by a 120-pixel box of the template
// (the tensor sources are zero-filled by their own out-of-bounds handling).  Interior strips: 128-pixel box, window
// [126 s - 8, 126 s + 120).  The last strip overlaps its neighbours' windows; it alone writes the columns [W - 119, W).
__device__ __forceinline__ int rc_wp_window(int s, int strips, int W) {
  return s == 0 ? -8 : (s == strips - 1 ? W - 120 : RC_VW * s - 8);
}

// an INTERIOR input row (1 <= Y <= H - 2) whose running index g has g % 8 == POS.  bars = shared address of in_full[0];
// ph = (g >> 3) & 1.  Barrier block layout: in_full[8] in_empty[8] acc_full[8] acc_empty[8] ...
// Two issuing threads share the rows (issuer k: g % 2 == k), so an output row's accumulator is complete when BOTH the
// issuer of its own input row (middle kernel row) and the issuer of the next input row (last contribution) have committed:
// acc_full counts two arrivals.
// pre: bit 0 = acc_empty of this row already seen complete, bit 1 = in_full (probed while this issuer's previous row was
// issued); returns the same two bits for row g + 2 (probed before this row's MMAs are issued) when PROBE, else 0.
template <int POS, int KB, bool PROBE, int ISS, int NSTG, int NS>
__device__ __forceinline__ uint32_t rc_fast_row(uint32_t tmem_base, uint32_t a_base, uint32_t a_step, uint32_t w_lo0, int kbn,
                                                uint32_t bars, int g, uint32_t pre) {
  constexpr int S_UP = NS - 1 - ((POS + 1) & (NS - 1));      // slot of output row g + 1 (first written here)
  constexpr int S_MID = NS - 1 - POS;
  constexpr int S_DN = NS - 1 - ((POS + NS - 1) & (NS - 1));   // slot of output row g - 1 (completed here)
  const uint32_t abars = bars + 16u * NSTG;                       // acc_full[0]; acc_empty[0] = abars + 8 * NS
  const uint32_t ph = (uint32_t)((g / NS) & 1);                    // slot-ring turn of row g
  const int stage = g & (NSTG - 1);
  const uint32_t ph_in = (uint32_t)((g / NSTG) & 1);
  if (!(pre & 1u))
    mbar_wait(abars + 8u * (NS + S_UP), (POS == NS - 1) ? ph : (ph ^ 1u), nullptr);   // acc_empty[S_UP], use (g + 1) / 8
  if (!(pre & 2u)) mbar_wait(bars + 8u * (uint32_t)stage, ph_in, nullptr);                   // in_full[stage]
  tc_fence_after();
  uint32_t nxt = 0;
  if (PROBE) {
    constexpr int NP = (POS + ISS) & (NS - 1);
    constexpr int S_UPN = NS - 1 - ((NP + 1) & (NS - 1));
    const uint32_t phn = (POS >= NS - ISS) ? (ph ^ 1u) : ph;
    const int gn = g + ISS;
    nxt = mbar_test(abars + 8u * (NS + S_UPN), (NP == NS - 1) ? phn : (phn ^ 1u)) |
          (mbar_test(bars + 8u * (uint32_t)(gn & (NSTG - 1)), (uint32_t)((gn / NSTG) & 1)) << 1);
  }
  const uint32_t a_lo = a_base + (uint32_t)stage * a_step;
  if (POS == NS - 1) {          // the ring wraps between rows g + 1 and g
    rc_run<0, 1, S_UP, KB>(tmem_base, a_lo, w_lo0, kbn);
    rc_run<1, 2, S_MID, KB>(tmem_base, a_lo, w_lo0, kbn);
  } else if (POS == 0) {           // ... between rows g and g - 1
    rc_run<0, 2, S_UP, KB>(tmem_base, a_lo, w_lo0, kbn);
    rc_run<2, 1, S_DN, KB>(tmem_base, a_lo, w_lo0, kbn);
  } else {
    rc_run<0, 3, S_UP, KB>(tmem_base, a_lo, w_lo0, kbn);
  }
  tc_commit(bars + 8u * (uint32_t)(NSTG + stage));   // in_empty[stage]
  if (ISS == 2) tc_commit(abars + 8u * S_MID);       // acc_full[S_MID]: this issuer's share of output row g
  tc_commit(abars + 8u * S_DN);                      // acc_full[S_DN]: the last contribution to output row g - 1
  return nxt;
}

// ISS = 2: issuer `who` (0 / 1) takes the input rows with g % 2 == who; ISS = 1: one issuer (who = 0) takes all rows
// NS = slots of the TMEM ring at tmem_base (8, or 4 in the two-conv kernel); a_step = bytes of one input row >> 4;
// barrier block at `bars`: in_full[NSTG] in_empty[NSTG] acc_full[NS] acc_empty[NS]
template <int KB, int ISS, int NSTG, int NS>
__device__ __forceinline__ void rc_conv_issuer(int H, long long items, int kbn, uint32_t a_step, uint32_t tmem_base,
                                               uint32_t a_base, uint32_t w_lo0, uint32_t bars, int who) {
  const uint32_t in_full = bars, in_empty = bars + 8u * NSTG, acc_full = bars + 16u * NSTG, acc_empty = acc_full + 8u * NS;
  int g0 = 0;                                // running index of the strip's first row: slot ring and stage ring position
  for (long long item = blockIdx.x; item < items; item += gridDim.x, g0 += H) {
    for (int Y = (ISS == 2) ? ((g0 ^ who) & 1) : 0; Y < H; Y += ISS) {
      const int g = g0 + Y;
      const int pos = g & (NS - 1);
      const uint32_t ph = (uint32_t)((g / NS) & 1);
      const int stage = g & (NSTG - 1);
      const uint32_t ph_in = (uint32_t)((g / NSTG) & 1);
      const uint32_t a_lo = a_base + (uint32_t)stage * a_step;
      if (Y >= 1 && Y + 1 < H) {
        if (pos == who && Y + NS < H) {     // an aligned group of eight interior rows (this issuer's share of it):
          uint32_t pre = 0;                    // straight-line code, barriers probed one row ahead
          if constexpr (NS == 4) {
            static_assert(NS != 4 || ISS == 1, "the four-slot ring is driven by one issuer");
            pre = rc_fast_row<0, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, pre);
            pre = rc_fast_row<1, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 1, pre);
            pre = rc_fast_row<2, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 2, pre);
            rc_fast_row<3, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 3, pre);
          } else if (ISS == 1) {
            pre = rc_fast_row<0, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, pre);
            pre = rc_fast_row<1, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 1, pre);
            pre = rc_fast_row<2, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 2, pre);
            pre = rc_fast_row<3, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 3, pre);
            pre = rc_fast_row<4, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 4, pre);
            pre = rc_fast_row<5, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 5, pre);
            pre = rc_fast_row<6, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 6, pre);
            rc_fast_row<7, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 7, pre);
          } else if (who == 0) {
            pre = rc_fast_row<0, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, pre);
            pre = rc_fast_row<2, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 2, pre);
            pre = rc_fast_row<4, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 4, pre);
            rc_fast_row<6, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 6, pre);
          } else {
            pre = rc_fast_row<1, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, pre);
            pre = rc_fast_row<3, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 2, pre);
            pre = rc_fast_row<5, KB, true, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 4, pre);
            rc_fast_row<7, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g + 6, pre);
          }
          Y += NS - ISS;
          continue;
        }
        if constexpr (NS == 4) {
          switch (pos) {
            case 0: rc_fast_row<0, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            case 1: rc_fast_row<1, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            case 2: rc_fast_row<2, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            default: rc_fast_row<3, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
          }
        } else {
          switch (pos) {
            case 0: rc_fast_row<0, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            case 1: rc_fast_row<1, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            case 2: rc_fast_row<2, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            case 3: rc_fast_row<3, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            case 4: rc_fast_row<4, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            case 5: rc_fast_row<5, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            case 6: rc_fast_row<6, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
            default: rc_fast_row<7, KB, false, ISS, NSTG, NS>(tmem_base, a_base, a_step, w_lo0, kbn, bars, g, 0u); break;
          }
        }
        continue;
      }
      // ---- first / last row of a strip: generic path (2 rows in H) ----
      // the slots this row writes FIRST must have been drained: row g + 1 (and row g at the top of a strip)
      if (Y == 0) mbar_wait(acc_empty + 8u * rc_slot<NS>(g), ph ^ 1u, nullptr);
      if (Y + 1 < H) mbar_wait(acc_empty + 8u * rc_slot<NS>(g + 1), (uint32_t)((((g + 1) / NS) & 1) ^ 1), nullptr);
      mbar_wait(in_full + 8u * (uint32_t)stage, ph_in, nullptr);
      tc_fence_after();
      // kernel rows kh_lo..kh_hi contribute (output row g + 1 - kh must exist).  The slots of rows g + 1, g, g - 1 are
      // consecutive except where the ring wraps: after kh = 0 when g % 8 == 7, after kh = 1 when g % 8 == 0.
      const int kh_lo = (Y + 1 < H) ? 0 : 1, kh_hi = (Y >= 1) ? 2 : 1;
      const int brk = (pos == NS - 1) ? 1 : (pos == 0 ? 2 : 3);          // first kernel row of a second run
      auto issue = [&](int kh0, int len) {
        const uint32_t idesc = RC_IDESC0 | ((uint32_t)(len * (RC_CO >> 3)) << 17);
        const uint32_t d_tmem = tmem_base + (uint32_t)(rc_slot<NS>(g + 1 - kh0) * RC_CO);
        uint32_t b_lo = w_lo0 + (uint32_t)(kh0 * RC_CO);
        uint32_t a = a_lo;
        for (int kb = 0; kb < kbn; ++kb) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            tc_mma_bf16(d_tmem, ((uint64_t)RC_A_HI << 32) | (a + (uint32_t)kw),
                        ((uint64_t)RC_B_HI << 32) | (b_lo + (uint32_t)(kw * (RC_WBLK >> 4))), idesc, 1u);
          a += 2 * (RC_CHUNK_BYTES >> 4);
          b_lo += 3 * (RC_WBLK >> 4);
        }
      };
      const int a_end = min(brk, kh_hi + 1);
      if (a_end > kh_lo) issue(kh_lo, a_end - kh_lo);
      const int b0 = max(brk, kh_lo);
      if (b0 <= kh_hi) issue(b0, kh_hi - b0 + 1);
      tc_commit(in_empty + 8u * (uint32_t)stage);
      if (ISS == 2) tc_commit(acc_full + 8u * rc_slot<NS>(g));         // this issuer's share of output row g
      if (Y >= 1) tc_commit(acc_full + 8u * rc_slot<NS>(g - 1));       // last contribution to output row g - 1
      if (Y == H - 1) tc_commit(acc_full + 8u * rc_slot<NS>(g));       // no input row below: the last arrival comes from here
    }
  }
}

// WP: the LAST K block is the waypoint source (ynet_tc_rowconv3x3_wp, n_ch <= 2): mapw / mapw120 describe the template's
// bf16 C8 planes with a 128- / 120-pixel box, mapz eight pixels of zeros.  A template parameter so that the plain kernels
// carry none of its code.
template <bool FUSE, bool WP = false>
__global__ void __launch_bounds__(FUSE ? RC_THREADS_FUSED : RC_THREADS_PLAIN, FUSE ? 1 : 2)
tc_rowconv_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                  const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap mapw,
                  const __grid_constant__ CUtensorMap mapw120, const __grid_constant__ CUtensorMap mapz, const RcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NTHREADS = FUSE ? RC_THREADS_FUSED : RC_THREADS_PLAIN;
  constexpr int W_EPI = 2;                          // first epilogue warp
  constexpr int NSTG = FUSE ? RC_STAGES_FUSED : RC_STAGES_PLAIN;
  constexpr int W_L4 = W_EPI + RC_EPI_WARPS;        // predictor MMA warp
  constexpr int W_SOFT = W_L4 + 1;                  // first soft-argmax warp

  unsigned char* s_w = smem;                                                    // kb * 3 * RC_WBLK
  unsigned char* s_in = s_w + (size_t)p.kb * 3 * RC_WBLK;                       // RC_STAGES * row_bytes (+ 1 KB slack)
  unsigned char* s_y = s_in + (size_t)NSTG * p.row_bytes + 1024;           // fused: RC_NY * RC_YROW
  unsigned char* s_pw = s_y + (FUSE ? RC_NY * RC_YROW : 0);                     // fused: 2 * PR_WBLK_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_pw + (FUSE ? 2 * PR_WBLK_BYTES : 0));
  uint64_t* in_full = bars;
  uint64_t* in_empty = in_full + NSTG;
  uint64_t* acc_full = in_empty + NSTG;
  uint64_t* acc_empty = acc_full + RC_NS;
  uint64_t* y_full = acc_empty + RC_NS;
  uint64_t* y_empty = y_full + RC_NY;
  uint64_t* p_full = y_empty + RC_NY;
  uint64_t* p_empty = p_full + 2;
  uint64_t* w_bar = p_empty + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(w_bar + 1);

  if (FUSE) pred_stage_weights(s_pw, p.pw, p.kbp, p.pn_pad, threadIdx.x, NTHREADS);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(smem_u32(&in_full[s]), 1);
      mbar_init(smem_u32(&in_empty[s]), 1);
    }
    for (int s = 0; s < RC_NS; ++s) {
      mbar_init(smem_u32(&acc_full[s]), 1);
      mbar_init(smem_u32(&acc_empty[s]), 4);
    }
    for (int s = 0; s < RC_NY; ++s) {
      mbar_init(smem_u32(&y_full[s]), 4);
      mbar_init(smem_u32(&y_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&p_full[s]), 1);
      mbar_init(smem_u32(&p_empty[s]), RC_SOFT_WARPS / 2);
    }
    mbar_init(smem_u32(w_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = FUSE ? 512u : 256u;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // the slot ring starts out holding the bias: the first four epilogue warps initialise their lane quadrants
  uint32_t bz[32];
  if (warp >= W_EPI && warp < W_EPI + RC_EPI_WARPS) {
#pragma unroll
    for (int i = 0; i < 32; ++i) bz[i] = __float_as_uint(p.bias[i]);
    if (warp < W_EPI + 4) {
      const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
      for (int s = 0; s < RC_NS; ++s) tmem_st32(t_row + (uint32_t)(s * RC_CO), bz);
    }
  }
  if (p.zero_fill) {     // K-padding chunk slots that no TMA transaction ever writes must read as zero
    uint4* z = reinterpret_cast<uint4*>(s_in);
    for (int i = threadIdx.x; i < NSTG * (p.row_bytes >> 4); i += NTHREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ===================== TMA producer: one input row per transaction =====================
    if (elect_one()) {
      const uint32_t wtotal = (uint32_t)(p.kb * 3 * RC_WBLK);
      mbar_expect_tx(smem_u32(w_bar), wtotal);
      for (uint32_t off = 0; off < wtotal; off += 18432) {
        const uint32_t nb = min(18432u, wtotal - off);
        bulk_load(smem_u32(s_w + off), p.w96 + off, nb, smem_u32(w_bar));
      }
      int stage = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = (int)(item / p.strips), s = (int)(item - (long long)n * p.strips);
        int ns[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int bm = p.src_mod[i];
          ns[i] = p.src_bcast[i] ? 0 : (bm > 0 ? n % bm : (bm < 0 ? n / (-bm) : n));
        }
        const int xw = WP ? rc_wp_window(s, p.strips, p.W) : s * RC_VW - 1;      // first pixel of the window
        const int c0 = 2 * xw;                              // 8-byte elements: two per pixel
        const int edge = WP ? (s == 0 ? 1 : (s == p.strips - 1 ? 2 : 0)) : 0;
        int wrow[2] = {0, 0}, wcol[2] = {0, 0}, wpar[2] = {0, 0};
        if (WP) {
          // window of channel c in the template planes: level 0 T[yl + Y][xl + x]; level 1 the 2x2 average whose top-left
          // corner is T[yl + 2Y][xl + 2x] = plane (yl & 1, xl & 1) at [(yl >> 1) + Y][(xl >> 1) + x]
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const size_t k = (size_t)n * p.wp_nch + min(c, p.wp_nch - 1);
            const int yl = p.wp_th / 2 - __float2int_rn(__ldg(p.wp_coords + 2 * k + 1));     // half to even == np.round
            const int xl = p.wp_tw / 2 - __float2int_rn(__ldg(p.wp_coords + 2 * k));
            wpar[c] = p.wp_level ? ((yl & 1) * 2 + (xl & 1)) : 0;
            wrow[c] = p.wp_level ? (yl >> 1) : yl;
            wcol[c] = 2 * (p.wp_level ? (xl >> 1) : xl) + c0;
          }
        }
        for (int Y = 0; Y < p.H; ++Y) {
          mbar_wait(smem_u32(&in_empty[stage]), phase ^ 1, nullptr);
          const uint32_t fb = smem_u32(&in_full[stage]);
          const uint32_t dst = smem_u32(s_in + (size_t)stage * p.row_bytes);
          mbar_expect_tx(fb, (uint32_t)p.row_tx);
          tma_load_4d(dst + (uint32_t)(p.src_coff[0] * RC_CHUNK_BYTES), &map0, fb, c0, Y, 0, ns[0]);
          if (p.n_src > 1) tma_load_4d(dst + (uint32_t)(p.src_coff[1] * RC_CHUNK_BYTES), &map1, fb, c0, Y, 0, ns[1]);
          if (p.n_src > 2) tma_load_4d(dst + (uint32_t)(p.src_coff[2] * RC_CHUNK_BYTES), &map2, fb, c0, Y, 0, ns[2]);
          if (WP) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c >= p.wp_nch) break;
              const uint32_t d = dst + (uint32_t)((p.wp_coff + c) * RC_CHUNK_BYTES);
              if (edge == 0) {
                tma_load_4d(d, &mapw, fb, wcol[c], wrow[c] + Y, wpar[c], 0);
              } else if (edge == 1) {        // window pixels 0-7 = x in [-8, 0): zeros; 8-127 = x in [0, 120)
                tma_load_4d(d, &mapz, fb, 0, 0, 0, 0);
                tma_load_4d(d + 128u, &mapw120, fb, wcol[c] + 16, wrow[c] + Y, wpar[c], 0);
              } else {                       // window pixels 0-119 = x in [W - 120, W); 120-127 = x >= W: zeros
                tma_load_4d(d, &mapw120, fb, wcol[c], wrow[c] + Y, wpar[c], 0);
                tma_load_4d(d + 1920u, &mapz, fb, 0, 0, 0, 0);
              }
            }
          }
          if (++stage == NSTG) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp < W_EPI) {
    // ===================== conv MMA issuers (warps 1, 2; the plain kernel issues from warp 1 alone) =====================
    if (elect_one()) {
      mbar_wait(smem_u32(w_bar), 0, nullptr);
      // descriptors: K-major, no swizzle; LBO = chunk stride, SBO = stride of 8-row groups (128 B: rows are contiguous)
      constexpr uint32_t A_LBO = (uint32_t)(RC_CHUNK_BYTES >> 4) << 16;
      constexpr uint32_t B_LBO = (uint32_t)((96 * 16) >> 4) << 16;
      const uint32_t w_lo0 = ((smem_u32(s_w) >> 4) & 0x3FFF) | B_LBO;
      const uint32_t a_base = ((smem_u32(s_in) >> 4) & 0x3FFF) | A_LBO;
      const uint32_t a_step = (uint32_t)(p.row_bytes >> 4);
      if (p.kb == 2 && !(p.dbg & 4))
        rc_conv_issuer<2, 1, NSTG, RC_NS>(p.H, p.items, 2, a_step, tmem_base, a_base, w_lo0, smem_u32(in_full), warp - 1);
      else
        rc_conv_issuer<0, 1, NSTG, RC_NS>(p.H, p.items, (p.dbg & 4) ? 1 : p.kb, a_step, tmem_base, a_base, w_lo0,
                                          smem_u32(in_full), warp - 1);
    }
  } else if (warp < W_EPI + RC_EPI_WARPS) {
    // ===================== conv epilogue: 2 sets x 4 warps; warp w owns TMEM lanes 32 (w % 4) .. +31 = pixels; set k
    // takes the rows with g % 2 == k =====================
    const int q = warp & 3, set = (warp - W_EPI) >> 2;
    const int m = q * 32 + lane;                    // pixel of the window row (window starts at xs0 - 1: lane = x - xs0)
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    int g0 = 0;                                     // running index of the strip's first row
    const int H = p.H;
    for (long long item = blockIdx.x; item < p.items; item += gridDim.x, g0 += H) {
      const int n = (int)(item / p.strips), s = (int)(item - (long long)n * p.strips);
      const int x = (WP ? rc_wp_window(s, p.strips, p.W) + 1 : s * RC_VW) + m;
      // WP: the columns [W - 119, W) belong to the last strip alone (an interior strip sees template values, not zeros,
      // beyond x = W - 1)
      const bool live = m < RC_VW && x < p.W && (!WP || (x >= 0 && (s == p.strips - 1 || x < p.W - 119)));
      const int npart = p.part_mod > 0 ? n % p.part_mod : (p.part_mod < 0 ? n / (-p.part_mod) : n);
      const uint4* part_px = p.part + (size_t)npart * p.part_bs + x;
      const bool has_part = p.part != nullptr && live;
      const size_t part_cs = (size_t)p.H * p.W;
      uint4 ph[4];                 // this pixel's partial sums of the row being processed (loaded one row ahead)
      if (!FUSE && has_part) {
        const uint4* pp = part_px + (size_t)((g0 ^ set) & 1) * p.W;
#pragma unroll
        for (int c = 0; c < 4; ++c) ph[c] = __ldg(pp + (size_t)c * part_cs);
      }
      for (int Y = (g0 ^ set) & 1; Y < H; Y += 2) {          // this set's rows: g % 2 == set
        const int g = g0 + Y;
        const int sl = rc_slot(g);
        mbar_wait(smem_u32(&acc_full[sl]), (uint32_t)((g / RC_NS) & 1), nullptr);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)(sl * RC_CO), v);
        tmem_st32(t_row + (uint32_t)(sl * RC_CO), bz);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty[sl]));
        if (!FUSE && has_part) {     // + the agent's hoisted encoder-feature share of this conv (fp32 add)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t hw[4] = {ph[c].x, ph[c].y, ph[c].z, ph[c].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              v[8 * c + 2 * k] = __float_as_uint(__uint_as_float(v[8 * c + 2 * k]) + __uint_as_float(hw[k] << 16));
              v[8 * c + 2 * k + 1] =
                  __float_as_uint(__uint_as_float(v[8 * c + 2 * k + 1]) + __uint_as_float(hw[k] & 0xFFFF0000u));
            }
          }
          if (p.part_chunks > 4) {   // the low halves of the partial sums (YNET_HOIST_LO=1)
            const uint4* pl = part_px + (size_t)Y * p.W + 4 * part_cs;
#pragma unroll
            for (int c = 0; c < 4; ++c) ph[c] = __ldg(pl + (size_t)c * part_cs);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t hw[4] = {ph[c].x, ph[c].y, ph[c].z, ph[c].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                v[8 * c + 2 * k] = __float_as_uint(__uint_as_float(v[8 * c + 2 * k]) + __uint_as_float(hw[k] << 16));
                v[8 * c + 2 * k + 1] =
                    __float_as_uint(__uint_as_float(v[8 * c + 2 * k + 1]) + __uint_as_float(hw[k] & 0xFFFF0000u));
              }
            }
          }
          if (Y + 2 < H) {           // the next row of this set: in flight during the stores and the next accumulator wait
            const uint4* pn = part_px + (size_t)(Y + 2) * p.W;
#pragma unroll
            for (int c = 0; c < 4; ++c) ph[c] = __ldg(pn + (size_t)c * part_cs);
          }
        }
        uint4 o[4];
        if (FUSE && (p.dbg & 2)) {
          const int b = g & (RC_NY - 1);
          if (lane == 0) mbar_wait(smem_u32(&y_empty[b]), (uint32_t)(((g / RC_NY) & 1) ^ 1), nullptr);
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&y_full[b]));
          continue;
        }
        if (p.relu) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            o[c].x = pack_bf16_act<true>(v[8 * c + 0], v[8 * c + 1]);
            o[c].y = pack_bf16_act<true>(v[8 * c + 2], v[8 * c + 3]);
            o[c].z = pack_bf16_act<true>(v[8 * c + 4], v[8 * c + 5]);
            o[c].w = pack_bf16_act<true>(v[8 * c + 6], v[8 * c + 7]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            o[c].x = pack_bf16_act<false>(v[8 * c + 0], v[8 * c + 1]);
            o[c].y = pack_bf16_act<false>(v[8 * c + 2], v[8 * c + 3]);
            o[c].z = pack_bf16_act<false>(v[8 * c + 4], v[8 * c + 5]);
            o[c].w = pack_bf16_act<false>(v[8 * c + 6], v[8 * c + 7]);
          }
        }
        if (FUSE) {
          const int b = g & (RC_NY - 1);
          if (lane == 0) mbar_wait(smem_u32(&y_empty[b]), (uint32_t)(((g / RC_NY) & 1) ^ 1), nullptr);
          __syncwarp();
          unsigned char* yb = s_y + (size_t)b * RC_YROW + m * 16;
#pragma unroll
          for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(yb + c * (RC_M * 16)) = o[c];
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&y_full[b]));
        } else if (live) {
          const int po = p.pad_out;
          const int Hp = p.H + 2 * po, Wp = p.W + 2 * po;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c >= p.out_chunks) break;
            uint4* dst = reinterpret_cast<uint4*>(p.out) + (((size_t)n * p.out_chunks + c) * Hp + (Y + po)) * Wp + (x + po);
            *dst = o[c];
            if (po) {      // replicate the border pixels into the ring (input of the phase-decomposed upconv)
              const int dyv = (Y == 0) ? -1 : ((Y == p.H - 1) ? 1 : 0);
              const int dxv = (x == 0) ? -1 : ((x == p.W - 1) ? 1 : 0);
              if (dyv != 0) dst[dyv * Wp] = o[c];
              if (dxv != 0) dst[dxv] = o[c];
              if (dyv != 0 && dxv != 0) dst[dyv * Wp + dxv] = o[c];
            }
          }
        }
      }
    }
  } else if (FUSE && warp == W_L4) {
    // ===================== predictor MMA issuer: D[channel, pixel] = Wp (M = 128, replicated) x row (N = 128) ==========
    if (elect_one()) {
      constexpr uint32_t IDESC_P = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      // descriptors as (hi, lo) halves; the row buffer b and the K block only move the start-address field
      constexpr uint32_t P_HI = (uint32_t)(128 >> 4) | (1u << 14);
      constexpr uint32_t P_LBO = (uint32_t)((RC_M * 16) >> 4) << 16;
      const uint32_t pa_lo = ((smem_u32(s_pw) >> 4) & 0x3FFF) | P_LBO;
      const uint32_t pb_lo = ((smem_u32(s_y) >> 4) & 0x3FFF) | P_LBO;
      const uint32_t yf = smem_u32(y_full), ye = smem_u32(y_empty), pf = smem_u32(p_full), pe = smem_u32(p_empty);
      const int kbpn = p.kbp;
      // One row: wait (unless the probe taken a row earlier already saw both barriers complete), probe the next row's
      // barriers, issue, commit: the probes overlap the ~100-cycle barrier latency with the issue work.
      uint32_t pre = 0;
      long long total_rows = 0;
      for (long long item = blockIdx.x; item < p.items; item += gridDim.x) total_rows += p.H;
      for (long long gl = 0; gl < total_rows; ++gl) {
        const int g = (int)gl;
        const uint32_t b = (uint32_t)(g & (RC_NY - 1)), a = (uint32_t)(g & 1);
        if (!(pre & 1u)) mbar_wait(pe + 8u * a, (uint32_t)(((g >> 1) & 1) ^ 1), nullptr);
        if (!(pre & 2u)) mbar_wait(yf + 8u * b, (uint32_t)((g / RC_NY) & 1), nullptr);
        tc_fence_after();
        {
          const int gn = g + 1;
          pre = mbar_test(pe + 8u * (uint32_t)(gn & 1), (uint32_t)(((gn >> 1) & 1) ^ 1)) |
                (mbar_test(yf + 8u * (uint32_t)(gn & (RC_NY - 1)), (uint32_t)((gn / RC_NY) & 1)) << 1);
        }
        const uint32_t d = tmem_base + RC_PACC + a * 128u;
        const uint32_t bl = pb_lo + b * (uint32_t)(RC_YROW >> 4);
        tc_mma_bf16(d, ((uint64_t)P_HI << 32) | pa_lo, ((uint64_t)P_HI << 32) | bl, IDESC_P, 0u);
        if (kbpn > 1)
          tc_mma_bf16(d, ((uint64_t)P_HI << 32) | (pa_lo + (uint32_t)(PR_WBLK_BYTES >> 4)),
                      ((uint64_t)P_HI << 32) | (bl + (uint32_t)((2 * RC_M * 16) >> 4)), IDESC_P, 1u);
        tc_commit(pf + 8u * a);
        tc_commit(ye + 8u * b);
      }
    }
  } else if (FUSE) {
    // ===================== soft-argmax: 2 sets x 4 lane quadrants x 2 halves = 16 warps; TMEM lane = channel, column =
    // pixel of the window row; set k takes the rows with g % 2 == k (= predictor accumulator k), each warp 16 pixels =====
    const int e = warp - W_SOFT;
    const int q = warp & 3, set = (e >> 2) & 1, half = e >> 3;
    const int col0 = 32 * q + 16 * half;
    const bool active = lane < p.c_pred;
    const float bias = active ? p.pbias[lane] : 0.f;
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + RC_PACC + (uint32_t)(set * 128 + col0);
    const uint32_t pf = smem_u32(&p_full[set]), pe = smem_u32(&p_empty[set]);
    const int H = p.H, W = p.W;
    int g0 = 0;
    for (long long item = blockIdx.x; item < p.items; item += gridDim.x, g0 += H) {
      const int n = (int)(item / p.strips), s = (int)(item - (long long)n * p.strips);
      const int x0 = s * RC_VW + col0;
      const int lim = min(W, s * RC_VW + RC_VW);          // first pixel beyond this strip's valid columns
      SoftState st{PR_NEG, 0.f, 0.f, 0.f};
      if (x0 < lim) {
        const bool full = x0 + 16 <= lim;
        for (int Y = (g0 ^ set) & 1; Y < H; Y += 2) {
          const int g = g0 + Y;
          mbar_wait(pf, (uint32_t)((g >> 1) & 1), nullptr);
          tc_fence_after();
          uint32_t v[16];
          tmem_ld16(t_addr, v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(pe);      // the values are in registers: release the accumulator
          if (p.dbg & 1) continue;
          if (full)
            softargmax_row16(st, v, bias, x0, Y);
          else
            softargmax_row16_masked(st, v, bias, x0, Y, lim);
        }
      } else {
        // nothing to reduce in these columns (right of the image / of the strip), but the accumulator hand-shake goes on
        for (int Y = (g0 ^ set) & 1; Y < H; Y += 2) {
          mbar_wait(pf, (uint32_t)(((g0 + Y) >> 1) & 1), nullptr);
          __syncwarp();
          if (lane == 0) mbar_arrive(pe);
        }
      }
      if (active) p.partial[((size_t)n * p.c_pred + lane) * p.slots + s * RC_SOFT_WARPS + e] = make_float4(st.m, st.s, st.sx, st.sy);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ================================================================================================================
// TWO stacked convs of a decoder block in one kernel (ynet.py:466-468: decoder.i.0 + ReLU + decoder.i.2 + ReLU), with the
// predictor + SoftArgmax2D behind them at the last level (TAIL): the 32-channel activation between the two convs never
// leaves the SM.  Conv A is the WP kernel above (sources [up, waypoint planes from the template], hoisted partial sums
// added by its epilogue); its epilogue writes the bf16 row into a four-row shared-memory ring in the K-major operand
// layout -- zeroed outside the image, which is conv B's padding -- and conv B marches over that ring exactly as conv A
// marches over the TMA ring (same issuer code, other barriers).  Conv B's output lane m is pixel xw + 2 + m: 124 valid
// pixels per strip.  TMEM: two rings of FOUR 32-column slots (A: columns 0-127, B: 128-255) + the two predictor
// accumulators (256-511).  Warps: 0 TMA | 1 conv A issuer | 2 conv B issuer | 3 predictor issuer | 4-11 conv A epilogue |
// 12-19 conv B epilogue (two sets of four each, alternating rows: one set's per-row latency chain of ~1 200 cycles --
// barrier wake-ups, tcgen05.ld / st, proxy fence -- would bound the kernel) | 20-27 soft-argmax (two sets of four, 32
// pixels per warp).
// The hoisted partial sums (64 B per pixel) reach conv A's epilogue through their own TMA row ring: with ONE CTA per SM
// and one epilogue warp set, per-lane global loads keep only 8 KB in flight per SM, and the kernel ran at the L2
// round trip per row (3.7 ms per 320 images against 2.1 ms without the partial sums).  The producer thread drives the
// input ring and the partial ring with two independent cursors (non-blocking barrier probes).
constexpr int R2_NS = 4;                         // slots per accumulator ring
constexpr int R2_NSTG = 8;                       // TMA ring of conv A's input rows
constexpr int R2_NA = 4;                         // ring of conv A's output rows = conv B's input stages
constexpr int R2_VW = 124;                       // valid output pixels per strip row
constexpr int R2_NY = 2;                         // conv B -> predictor row ring (tail)
constexpr int R2_PROW = 4 * RC_CHUNK_BYTES;      // one row of partial sums: [4 chunks (hi)][128 px][8 ch] bf16
constexpr int R2_THREADS_TAIL = 32 * 28;
constexpr int R2_THREADS_PLAIN = 32 * 20;
constexpr int R2_SOFT_WARPS = 16;
constexpr uint32_t R2_BCOL = 128;                // first TMEM column of conv B's ring

struct Rc2Params {
  RcParams a;                  // conv A: sources, waypoint source, partial sums, weights (w96), bias, geometry
  const unsigned char* wb96;   // conv B weights [2][kw][2][96][8]
  const float* bias_b;         // 32 floats
  int relu_b;
};

// window start (x of input pixel 0) of strip s: see rc_wp_window; conv B's valid columns shrink the stride to 124
__device__ __forceinline__ int rc2_window(int s, int strips, int W) {
  return s == 0 ? -8 : (s == strips - 1 ? W - 120 : R2_VW * s - 8);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* b) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "f"(b[0]), "f"(b[1]), "f"(b[2]), "f"(b[3]), "f"(b[4]), "f"(b[5]), "f"(b[6]), "f"(b[7]), "f"(b[8]), "f"(b[9]),
      "f"(b[10]), "f"(b[11]), "f"(b[12]), "f"(b[13]), "f"(b[14]), "f"(b[15])
      : "memory");
}

// soft-argmax of a 16-pixel row segment restricted to the columns [lo, hi)
__device__ __forceinline__ void softargmax_row16_range(SoftState& st, const uint32_t (&v)[16], float bias, int x0, int y, int lo,
                                                       int hi) {
  uint32_t vm[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) vm[i] = (x0 + i >= lo && x0 + i < hi) ? v[i] : __float_as_uint(-1.0e30f);
  softargmax_row16(st, vm, bias, x0, y);
}

// shared-memory accesses by 32-bit shared address (one register instead of a 64-bit generic pointer: the tail variant runs
// 896 threads at 72 registers, and spilled loop state in the epilogues costs an L2 round trip per row)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 32 lanes x 32 columns <- 0: a drained accumulator slot is handed back zeroed (one operand register; the bias is added
// by the epilogue, four channels at a time from shared memory, instead of living in 32 registers or in the slot)
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr),
      "r"(0u)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 columns <- 16 floats at shared address b (plain variant: the slot is handed back holding the bias, like
// tc_rowconv_kernel's, which keeps the two-conv block bit-identical to the two separate launches)
__device__ __forceinline__ void tmem_st16_s(uint32_t taddr, uint32_t b) {
  const uint4 b0 = lds128(b), b1 = lds128(b + 16), b2 = lds128(b + 32), b3 = lds128(b + 48);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(b0.x), "r"(b0.y), "r"(b0.z), "r"(b0.w), "r"(b1.x), "r"(b1.y), "r"(b1.z), "r"(b1.w), "r"(b2.x), "r"(b2.y), "r"(b2.z),
      "r"(b2.w), "r"(b3.x), "r"(b3.y), "r"(b3.z), "r"(b3.w)
      : "memory");
}
template <bool ZERO>
__device__ __forceinline__ void slot_reset(uint32_t taddr, uint32_t sbias) {
  if (ZERO) {
    tmem_zero32(taddr);
  } else {
    tmem_st16_s(taddr, sbias);
    tmem_st16_s(taddr + 16, sbias + 64);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
}
// v[8c .. 8c + 7] += the eight bias floats at shared address b
__device__ __forceinline__ void add_bias8(uint32_t (&v)[32], int c, uint32_t b) {
  const uint4 b0 = lds128(b), b1 = lds128(b + 16);
  v[8 * c + 0] = __float_as_uint(__uint_as_float(v[8 * c + 0]) + __uint_as_float(b0.x));
  v[8 * c + 1] = __float_as_uint(__uint_as_float(v[8 * c + 1]) + __uint_as_float(b0.y));
  v[8 * c + 2] = __float_as_uint(__uint_as_float(v[8 * c + 2]) + __uint_as_float(b0.z));
  v[8 * c + 3] = __float_as_uint(__uint_as_float(v[8 * c + 3]) + __uint_as_float(b0.w));
  v[8 * c + 4] = __float_as_uint(__uint_as_float(v[8 * c + 4]) + __uint_as_float(b1.x));
  v[8 * c + 5] = __float_as_uint(__uint_as_float(v[8 * c + 5]) + __uint_as_float(b1.y));
  v[8 * c + 6] = __float_as_uint(__uint_as_float(v[8 * c + 6]) + __uint_as_float(b1.z));
  v[8 * c + 7] = __float_as_uint(__uint_as_float(v[8 * c + 7]) + __uint_as_float(b1.w));
}

template <bool TAIL>
__global__ void __launch_bounds__(TAIL ? R2_THREADS_TAIL : R2_THREADS_PLAIN, 1)
tc_rowconv2_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                   const __grid_constant__ CUtensorMap mapw, const __grid_constant__ CUtensorMap mapw120,
                   const __grid_constant__ CUtensorMap mapz, const __grid_constant__ CUtensorMap mapp, const Rc2Params pp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const RcParams& p = pp.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NTHREADS = TAIL ? R2_THREADS_TAIL : R2_THREADS_PLAIN;
  constexpr int SETS = TAIL ? 1 : 2;               // epilogue warp sets per conv (tail: the thread budget goes to 16 soft-argmax warps)
  constexpr int W_EPA = 4, W_EPB = W_EPA + 4 * SETS, W_SOFT = W_EPB + 4 * SETS;
  constexpr int NPR = TAIL ? 8 : 4;                // rows of the partial-sum ring
  constexpr int NS = TAIL ? R2_NS : 8;             // slots per accumulator ring: the tail shares the 512 columns with the predictor
  constexpr uint32_t BCOL = TAIL ? R2_BCOL : 256u; // first column of conv B's ring

  unsigned char* s_wa = smem;                                                   // kb * 3 * RC_WBLK
  unsigned char* s_wb = s_wa + (size_t)p.kb * 3 * RC_WBLK;                      // 2 * 3 * RC_WBLK
  unsigned char* s_in = s_wb + 2 * 3 * RC_WBLK;                                 // R2_NSTG * row_bytes (+ 1 KB slack)
  unsigned char* s_a = s_in + (size_t)R2_NSTG * p.row_bytes + 1024;             // R2_NA * RC_YROW (+ 1 KB slack)
  unsigned char* s_pr = s_a + R2_NA * RC_YROW + 1024;                           // NPR * R2_PROW
  unsigned char* s_y = s_pr + NPR * R2_PROW;                                    // tail: R2_NY * RC_YROW
  unsigned char* s_pw = s_y + (TAIL ? R2_NY * RC_YROW : 0);                     // tail: 2 * PR_WBLK_BYTES
  float* s_bias = reinterpret_cast<float*>(s_pw + (TAIL ? 2 * PR_WBLK_BYTES : 0));   // [A: 32][B: 32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 64);
  uint64_t* in_full = bars;                        // block A: in_full[8] in_empty[8] acca_full[4] acca_empty[4]
  uint64_t* in_empty = in_full + R2_NSTG;
  uint64_t* acca_full = in_empty + R2_NSTG;
  uint64_t* acca_empty = acca_full + NS;
  uint64_t* a_full = acca_empty + NS;           // block B: a_full[4] a_empty[4] accb_full[4] accb_empty[4]
  uint64_t* a_empty = a_full + R2_NA;
  uint64_t* accb_full = a_empty + R2_NA;
  uint64_t* accb_empty = accb_full + NS;
  uint64_t* y_full = accb_empty + NS;
  uint64_t* y_empty = y_full + R2_NY;
  uint64_t* p_full = y_empty + R2_NY;
  uint64_t* p_empty = p_full + 2;
  uint64_t* pr_full = p_empty + 2;
  uint64_t* pr_empty = pr_full + NPR;
  uint64_t* w_bar = pr_empty + NPR;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(w_bar + 1);

  if (TAIL) pred_stage_weights(s_pw, p.pw, p.kbp, p.pn_pad, threadIdx.x, NTHREADS);
  if (threadIdx.x < 64) s_bias[threadIdx.x] = threadIdx.x < 32 ? p.bias[threadIdx.x] : pp.bias_b[threadIdx.x - 32];
  if (threadIdx.x == 0) {
    for (int s = 0; s < R2_NSTG; ++s) {
      mbar_init(smem_u32(&in_full[s]), 1);
      mbar_init(smem_u32(&in_empty[s]), 1);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&acca_full[s]), 1);
      mbar_init(smem_u32(&acca_empty[s]), 4);
      mbar_init(smem_u32(&accb_full[s]), 1);
      mbar_init(smem_u32(&accb_empty[s]), 4);
    }
    for (int s = 0; s < R2_NA; ++s) {
      mbar_init(smem_u32(&a_full[s]), 4);
      mbar_init(smem_u32(&a_empty[s]), 1);
    }
    for (int s = 0; s < R2_NY; ++s) {
      mbar_init(smem_u32(&y_full[s]), 4);
      mbar_init(smem_u32(&y_empty[s]), 1);
    }
    for (int s = 0; s < NPR; ++s) {
      mbar_init(smem_u32(&pr_full[s]), 1);
      mbar_init(smem_u32(&pr_empty[s]), 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&p_full[s]), 1);
      mbar_init(smem_u32(&p_empty[s]), R2_SOFT_WARPS / 2);
    }
    mbar_init(smem_u32(w_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = 512u;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {   // K-padding chunk slots of the TMA ring that no transaction writes, and the slack behind the rings, must read as zero
    uint4* z = reinterpret_cast<uint4*>(s_in);
    const int nz = (int)((s_pr - s_in) >> 4);
    for (int i = threadIdx.x; i < nz; i += NTHREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if ((warp >= W_EPA && warp < W_EPA + 4) || (warp >= W_EPB && warp < W_EPB + 4)) {   // both rings start out zeroed (tail) / holding their bias (plain)
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >= W_EPB ? BCOL : 0u);
#pragma unroll
    for (int s = 0; s < NS; ++s) slot_reset<TAIL>(t_row + (uint32_t)(s * RC_CO), smem_u32(s_bias) + (warp >= W_EPB ? 128u : 0u));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ===================== TMA producer (conv A's input rows; see tc_rowconv_kernel<false, true>) =====================
    if (elect_one()) {
      const uint32_t wa = (uint32_t)(p.kb * 3 * RC_WBLK), wb = (uint32_t)(2 * 3 * RC_WBLK);
      mbar_expect_tx(smem_u32(w_bar), wa + wb);
      for (uint32_t off = 0; off < wa; off += 18432) bulk_load(smem_u32(s_wa + off), p.w96 + off, min(18432u, wa - off), smem_u32(w_bar));
      bulk_load(smem_u32(s_wb), pp.wb96, wb, smem_u32(w_bar));
      // two cursors over the same (item, row) sequence: conv A's input rows and the rows of partial sums
      const bool has_part = p.part != nullptr && !(p.dbg & 8);
      long long it_i = blockIdx.x, it_p = has_part ? (long long)blockIdx.x : p.items;
      int Yi = 0, Yp = 0, st_i = 0, st_p = 0;
      uint32_t ph_i = 0, ph_p = 0;
      int ns[2] = {0, 0}, c0 = 0, edge = 0, wrow[2] = {0, 0}, wcol[2] = {0, 0}, wpar[2] = {0, 0};
      int np_img = 0, c0p = 0;
      bool new_i = true, new_p = true;
      while (it_i < p.items || it_p < p.items) {
        if (it_i < p.items) {
          if (new_i) {
            const int n = (int)(it_i / p.strips), s = (int)(it_i - (long long)n * p.strips);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int bm = p.src_mod[i];
              ns[i] = p.src_bcast[i] ? 0 : (bm > 0 ? n % bm : (bm < 0 ? n / (-bm) : n));
            }
            c0 = 2 * rc2_window(s, p.strips, p.W);          // 8-byte elements: two per pixel
            edge = s == 0 ? 1 : (s == p.strips - 1 ? 2 : 0);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const size_t k = (size_t)n * p.wp_nch + min(c, p.wp_nch - 1);
              const int yl = p.wp_th / 2 - __float2int_rn(__ldg(p.wp_coords + 2 * k + 1));     // half to even == np.round
              const int xl = p.wp_tw / 2 - __float2int_rn(__ldg(p.wp_coords + 2 * k));
              wpar[c] = p.wp_level ? ((yl & 1) * 2 + (xl & 1)) : 0;
              wrow[c] = p.wp_level ? (yl >> 1) : yl;
              wcol[c] = 2 * (p.wp_level ? (xl >> 1) : xl) + c0;
            }
            new_i = false;
          }
          if (mbar_test(smem_u32(&in_empty[st_i]), ph_i ^ 1)) {
            const uint32_t fb = smem_u32(&in_full[st_i]);
            const uint32_t dst = smem_u32(s_in + (size_t)st_i * p.row_bytes);
            mbar_expect_tx(fb, (uint32_t)(p.row_tx - ((p.dbg & 16) ? p.wp_nch * RC_CHUNK_BYTES : 0)));
            tma_load_4d(dst + (uint32_t)(p.src_coff[0] * RC_CHUNK_BYTES), &map0, fb, c0, Yi, 0, ns[0]);
            if (p.n_src > 1) tma_load_4d(dst + (uint32_t)(p.src_coff[1] * RC_CHUNK_BYTES), &map1, fb, c0, Yi, 0, ns[1]);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c >= p.wp_nch || (p.dbg & 16)) break;
              const uint32_t d = dst + (uint32_t)((p.wp_coff + c) * RC_CHUNK_BYTES);
              if (edge == 0) {
                tma_load_4d(d, &mapw, fb, wcol[c], wrow[c] + Yi, wpar[c], 0);
              } else if (edge == 1) {
                tma_load_4d(d, &mapz, fb, 0, 0, 0, 0);
                tma_load_4d(d + 128u, &mapw120, fb, wcol[c] + 16, wrow[c] + Yi, wpar[c], 0);
              } else {
                tma_load_4d(d, &mapw120, fb, wcol[c], wrow[c] + Yi, wpar[c], 0);
                tma_load_4d(d + 1920u, &mapz, fb, 0, 0, 0, 0);
              }
            }
            if (++st_i == R2_NSTG) {
              st_i = 0;
              ph_i ^= 1;
            }
            if (++Yi == p.H) {
              Yi = 0;
              it_i += gridDim.x;
              new_i = true;
            }
          }
        }
        if (it_p < p.items) {
          if (new_p) {
            const int n = (int)(it_p / p.strips), s = (int)(it_p - (long long)n * p.strips);
            np_img = p.part_mod > 0 ? n % p.part_mod : (p.part_mod < 0 ? n / (-p.part_mod) : n);
            c0p = 2 * rc2_window(s, p.strips, p.W);
            new_p = false;
          }
          if (mbar_test(smem_u32(&pr_empty[st_p]), ph_p ^ 1)) {
            const uint32_t fb = smem_u32(&pr_full[st_p]);
            mbar_expect_tx(fb, (uint32_t)R2_PROW);
            tma_load_4d(smem_u32(s_pr + (size_t)st_p * R2_PROW), &mapp, fb, c0p, Yp, 0, np_img);
            if (++st_p == NPR) {
              st_p = 0;
              ph_p ^= 1;
            }
            if (++Yp == p.H) {
              Yp = 0;
              it_p += gridDim.x;
              new_p = true;
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== conv A / conv B MMA issuers =====================
    if (elect_one()) {
      mbar_wait(smem_u32(w_bar), 0, nullptr);
      constexpr uint32_t A_LBO = (uint32_t)(RC_CHUNK_BYTES >> 4) << 16;
      constexpr uint32_t B_LBO = (uint32_t)((96 * 16) >> 4) << 16;
      if (warp == 1) {
        const uint32_t w_lo0 = ((smem_u32(s_wa) >> 4) & 0x3FFF) | B_LBO;
        const uint32_t a_base = ((smem_u32(s_in) >> 4) & 0x3FFF) | A_LBO;
        const uint32_t a_step = (uint32_t)(p.row_bytes >> 4);
        if (p.kb == 2 && !(p.dbg & 4))
          rc_conv_issuer<2, 1, R2_NSTG, NS>(p.H, p.items, 2, a_step, tmem_base, a_base, w_lo0, smem_u32(in_full), 0);
        else
          rc_conv_issuer<0, 1, R2_NSTG, NS>(p.H, p.items, (p.dbg & 4) ? 1 : p.kb, a_step, tmem_base, a_base, w_lo0,
                                               smem_u32(in_full), 0);
      } else {
        const uint32_t w_lo0 = ((smem_u32(s_wb) >> 4) & 0x3FFF) | B_LBO;
        const uint32_t a_base = ((smem_u32(s_a) >> 4) & 0x3FFF) | A_LBO;
        rc_conv_issuer<2, 1, R2_NA, NS>(p.H, p.items, 2, (uint32_t)(RC_YROW >> 4), tmem_base + BCOL, a_base, w_lo0,
                                           smem_u32(a_full), 0);
      }
    }
  } else if (warp == 3) {
    // ===================== predictor MMA issuer (tail): D[channel, pixel] = Wp (M = 128, replicated) x row (N = 128) =====
    if (TAIL && elect_one()) {
      constexpr uint32_t IDESC_P = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      constexpr uint32_t P_HI = (uint32_t)(128 >> 4) | (1u << 14);
      constexpr uint32_t P_LBO = (uint32_t)((RC_M * 16) >> 4) << 16;
      const uint32_t pa_lo = ((smem_u32(s_pw) >> 4) & 0x3FFF) | P_LBO;
      const uint32_t pb_lo = ((smem_u32(s_y) >> 4) & 0x3FFF) | P_LBO;
      const uint32_t yf = smem_u32(y_full), ye = smem_u32(y_empty), pf = smem_u32(p_full), pe = smem_u32(p_empty);
      const int kbpn = p.kbp;
      uint32_t pre = 0;
      long long total_rows = 0;
      for (long long item = blockIdx.x; item < p.items; item += gridDim.x) total_rows += p.H;
      for (long long gl = 0; gl < total_rows; ++gl) {
        const int g = (int)gl;
        const uint32_t b = (uint32_t)(g & (R2_NY - 1)), a = (uint32_t)(g & 1);
        if (!(pre & 1u)) mbar_wait(pe + 8u * a, (uint32_t)(((g >> 1) & 1) ^ 1), nullptr);
        if (!(pre & 2u)) mbar_wait(yf + 8u * b, (uint32_t)((g / R2_NY) & 1), nullptr);
        tc_fence_after();
        {
          const int gn = g + 1;
          pre = mbar_test(pe + 8u * (uint32_t)(gn & 1), (uint32_t)(((gn >> 1) & 1) ^ 1)) |
                (mbar_test(yf + 8u * (uint32_t)(gn & (R2_NY - 1)), (uint32_t)((gn / R2_NY) & 1)) << 1);
        }
        const uint32_t d = tmem_base + RC_PACC + a * 128u;
        const uint32_t bl = pb_lo + b * (uint32_t)(RC_YROW >> 4);
        tc_mma_bf16(d, ((uint64_t)P_HI << 32) | pa_lo, ((uint64_t)P_HI << 32) | bl, IDESC_P, 0u);
        if (kbpn > 1)
          tc_mma_bf16(d, ((uint64_t)P_HI << 32) | (pa_lo + (uint32_t)(PR_WBLK_BYTES >> 4)),
                      ((uint64_t)P_HI << 32) | (bl + (uint32_t)((2 * RC_M * 16) >> 4)), IDESC_P, 1u);
        tc_commit(pf + 8u * a);
        tc_commit(ye + 8u * b);
      }
    }
  } else if (warp < W_EPB) {
    // ===================== conv A epilogue: + partial sums, ReLU, zero outside the image -> row ring of conv B ==========
    const int q = warp & 3, set = (warp - W_EPA) >> 2;
    const int m = q * 32 + lane;                    // lane m = pixel xw + 1 + m
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t bar0 = smem_u32(bars);
    constexpr uint32_t B_ACCA_FULL = 8u * (2 * R2_NSTG), B_ACCA_EMPTY = B_ACCA_FULL + 8u * NS, B_A_FULL = B_ACCA_EMPTY + 8u * NS,
                       B_A_EMPTY = B_A_FULL + 8u * R2_NA;
    const uint32_t b_pr_full = smem_u32(pr_full), b_pr_empty = smem_u32(pr_empty);
    const uint32_t sa_lane = smem_u32(s_a) + (uint32_t)(m * 16), sbias = smem_u32(s_bias);
    const uint32_t spr_lane = smem_u32(s_pr) + (uint32_t)(min(m + 1, RC_M - 1) * 16);   // this lane's pixel in a partial-sum row
    const bool has_part = p.part != nullptr && !(p.dbg & 8);
    const int H = p.H, W = p.W, strips = p.strips, items = (int)p.items;
    int g0 = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, g0 += H) {
      const int x = rc2_window(item % strips, strips, W) + 1 + m;
      const bool inside = x >= 0 && x < W;          // (lanes 126 / 127 hold garbage that conv B's valid lanes never read)
      for (int Y = (SETS == 2) ? ((g0 ^ set) & 1) : 0; Y < H; Y += SETS) {          // this set's rows: g % SETS == set
        const int g = g0 + Y;
        const uint32_t sl = (uint32_t)rc_slot<NS>(g);
        mbar_wait(bar0 + B_ACCA_FULL + 8u * sl, (uint32_t)((g / NS) & 1), nullptr);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(t_row + sl * RC_CO, v);
        slot_reset<TAIL>(t_row + sl * RC_CO, sbias);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + B_ACCA_EMPTY + 8u * sl);
        if (has_part) {              // + the agent's hoisted encoder-feature share of conv A (fp32 add), from its TMA ring
          const uint32_t pb = (uint32_t)(g & (NPR - 1));
          mbar_wait(b_pr_full + 8u * pb, (uint32_t)((g / NPR) & 1), nullptr);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 h4 = lds128(spr_lane + pb * R2_PROW + (uint32_t)(c * RC_CHUNK_BYTES));
            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              v[8 * c + 2 * k] = __float_as_uint(__uint_as_float(v[8 * c + 2 * k]) + __uint_as_float(hw[k] << 16));
              v[8 * c + 2 * k + 1] =
                  __float_as_uint(__uint_as_float(v[8 * c + 2 * k + 1]) + __uint_as_float(hw[k] & 0xFFFF0000u));
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(b_pr_empty + 8u * pb);
        }
        const uint32_t b = (uint32_t)(g & (R2_NA - 1));
        if (lane == 0) mbar_wait(bar0 + B_A_EMPTY + 8u * b, (uint32_t)(((g / R2_NA) & 1) ^ 1), nullptr);
        __syncwarp();
        if (!(p.dbg & 2)) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (TAIL) add_bias8(v, c, sbias + (uint32_t)(c * 32));
            uint4 o;
            o.x = inside ? pack_bf16_act<true>(v[8 * c + 0], v[8 * c + 1]) : 0u;
            o.y = inside ? pack_bf16_act<true>(v[8 * c + 2], v[8 * c + 3]) : 0u;
            o.z = inside ? pack_bf16_act<true>(v[8 * c + 4], v[8 * c + 5]) : 0u;
            o.w = inside ? pack_bf16_act<true>(v[8 * c + 6], v[8 * c + 7]) : 0u;
            sts128(sa_lane + b * RC_YROW + (uint32_t)(c * (RC_M * 16)), o);
          }
        }
        if (!(p.dbg & 16)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + B_A_FULL + 8u * b);
      }
    }
  } else if (warp < W_SOFT) {
    // ===================== conv B epilogue: ReLU -> predictor operand ring (tail) or the C8 planes (plain) ==============
    const int q = warp & 3, set = (warp - W_EPB) >> 2;
    const int m = q * 32 + lane;                    // lane m = pixel xw + 2 + m
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + BCOL;
    const uint32_t b_accb_full = smem_u32(accb_full), b_accb_empty = smem_u32(accb_empty);
    const uint32_t b_y_full = smem_u32(y_full), b_y_empty = smem_u32(y_empty);
    const uint32_t sy_lane = smem_u32(s_y) + (uint32_t)(m * 16), sbias = smem_u32(s_bias) + 128u;
    const int H = p.H, W = p.W, strips = p.strips, items = (int)p.items;
    const bool relu_b = pp.relu_b != 0;
    int g0 = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, g0 += H) {
      const int n = item / strips, s = item - n * strips;
      const int x = rc2_window(s, strips, W) + 2 + m;
      // the columns [W - 118, W) belong to the last strip alone (see tc_rowconv_kernel<false, true>)
      const bool live = m < R2_VW && x >= 0 && x < W && (s == strips - 1 || x < W - 118);
      for (int Y = (SETS == 2) ? ((g0 ^ set) & 1) : 0; Y < H; Y += SETS) {          // this set's rows: g % SETS == set
        const int g = g0 + Y;
        const uint32_t sl = (uint32_t)rc_slot<NS>(g);
        mbar_wait(b_accb_full + 8u * sl, (uint32_t)((g / NS) & 1), nullptr);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(t_row + sl * RC_CO, v);
        slot_reset<TAIL>(t_row + sl * RC_CO, sbias);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_accb_empty + 8u * sl);
        uint4 o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (TAIL) add_bias8(v, c, sbias + (uint32_t)(c * 32));
          if (relu_b) {
            o[c].x = pack_bf16_act<true>(v[8 * c + 0], v[8 * c + 1]);
            o[c].y = pack_bf16_act<true>(v[8 * c + 2], v[8 * c + 3]);
            o[c].z = pack_bf16_act<true>(v[8 * c + 4], v[8 * c + 5]);
            o[c].w = pack_bf16_act<true>(v[8 * c + 6], v[8 * c + 7]);
          } else {
            o[c].x = pack_bf16_act<false>(v[8 * c + 0], v[8 * c + 1]);
            o[c].y = pack_bf16_act<false>(v[8 * c + 2], v[8 * c + 3]);
            o[c].z = pack_bf16_act<false>(v[8 * c + 4], v[8 * c + 5]);
            o[c].w = pack_bf16_act<false>(v[8 * c + 6], v[8 * c + 7]);
          }
        }
        if (TAIL) {
          const uint32_t b = (uint32_t)(g & (R2_NY - 1));
          if (lane == 0) mbar_wait(b_y_empty + 8u * b, (uint32_t)(((g / R2_NY) & 1) ^ 1), nullptr);
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; ++c) sts128(sy_lane + b * RC_YROW + (uint32_t)(c * (RC_M * 16)), o[c]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(b_y_full + 8u * b);
        } else if (live) {
          const int po = p.pad_out;
          const int Hp = H + 2 * po, Wp = W + 2 * po;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c >= p.out_chunks) break;
            uint4* dst = reinterpret_cast<uint4*>(p.out) + (((size_t)n * p.out_chunks + c) * Hp + (Y + po)) * Wp + (x + po);
            *dst = o[c];
            if (po) {      // replicate the border pixels into the ring (input of the phase-decomposed upconv)
              const int dyv = (Y == 0) ? -1 : ((Y == H - 1) ? 1 : 0);
              const int dxv = (x == 0) ? -1 : ((x == W - 1) ? 1 : 0);
              if (dyv != 0) dst[dyv * Wp] = o[c];
              if (dxv != 0) dst[dxv] = o[c];
              if (dyv != 0 && dxv != 0) dst[dyv * Wp + dxv] = o[c];
            }
          }
        }
      }
    }
  } else if (TAIL) {
    // ===================== soft-argmax: 2 sets x 4 lane quadrants x 2 halves = 16 warps, 16 pixels each; TMEM lane =
    // channel, column m of the accumulator = pixel xw + 2 + m; set k takes the rows with g % 2 == k ==========
    const int e = warp - W_SOFT;
    const int q = warp & 3, set = (e >> 2) & 1, half = e >> 3;
    const int col0 = 32 * q + 16 * half;
    const bool active = lane < p.c_pred;
    const float bias = active ? p.pbias[lane] : 0.f;
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + RC_PACC + (uint32_t)(set * 128 + col0);
    const uint32_t pf = smem_u32(&p_full[set]), pe = smem_u32(&p_empty[set]);
    const int H = p.H, W = p.W, strips = p.strips, items = (int)p.items;
    int g0 = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, g0 += H) {
      const int n = item / strips, s = item - n * strips;
      const int xs = rc2_window(s, strips, W) + 2;             // pixel of accumulator column 0
      const int x0 = xs + col0;
      const int lo = max(0, xs), hi = (s == strips - 1) ? min(W, xs + R2_VW) : min(W - 118, xs + R2_VW);
      SoftState st{PR_NEG, 0.f, 0.f, 0.f};
      const bool any = x0 < hi && x0 + 16 > lo, full = x0 >= lo && x0 + 16 <= hi;
      for (int Y = (g0 ^ set) & 1; Y < H; Y += 2) {
        const int g = g0 + Y;
        mbar_wait(pf, (uint32_t)((g >> 1) & 1), nullptr);
        if (!any) {                  // nothing to reduce in these columns, but the accumulator hand-shake goes on
          __syncwarp();
          if (lane == 0) mbar_arrive(pe);
          continue;
        }
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16(t_addr, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pe);      // the values are in registers: release the accumulator
        if (p.dbg & 1) continue;
        if (full)
          softargmax_row16(st, v, bias, x0, Y);
        else
          softargmax_row16_range(st, v, bias, x0, Y, lo, hi);
      }
      if (active) p.partial[((size_t)n * p.c_pred + lane) * p.slots + s * R2_SOFT_WARPS + e] = make_float4(st.m, st.s, st.sx, st.sy);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// weights OIHW f32 (C_out <= 32, C_in <= 64, 3, 3) -> [kb][kw][2 chunks][96 = kh * 32 + co][8 ch] bf16
__global__ void __launch_bounds__(256)
rc_pack_weights_kernel(const float* __restrict__ w, int C_out, int C_in, int kb, __nv_bfloat16* __restrict__ out) {
  const int total = kb * 3 * 2 * 96 * 8;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int k8 = t & 7;
    int r = t >> 3;
    const int nrow = r % 96;
    r /= 96;
    const int c = r & 1;
    r >>= 1;
    const int kw = r % 3, b = r / 3;
    const int kh = nrow / RC_CO, co = nrow - kh * RC_CO;
    const int ci = b * 16 + c * 8 + k8;
    const float v = (co < C_out && ci < C_in) ? w[(((size_t)co * C_in + ci) * 3 + kh) * 3 + kw] : 0.f;
    out[t] = __float2bfloat16_rn(v);
  }
}

struct Rc2Host {              // the second conv of tc_rowconv2_kernel
  const void* wb96;
  const float* bias_b;
  int relu_b;
  bool tail;
};

static int rc_launch(const char* who, bool fuse, const ynet_tc_src* srcs, int n_src, const ynet_tc_src* partial, int N,
                     int H, int W, RcParams& p, void* stream, const ynet_tc_wp_src* wp = nullptr,
                     const Rc2Host* two = nullptr) {
  if (!(srcs && n_src >= 1 && n_src <= 3)) {
    set_error("%s: 1..3 conv sources", who);
    return YNET_E_INVALID;
  }
  if (!(N >= 0 && H >= 2 && W >= 1)) {
    set_error("%s: needs H >= 2 (got %d x %d)", who, H, W);
    return YNET_E_UNSUPPORTED;
  }
  if (N == 0) return YNET_OK;
  EncodeTiledFn encode = tc_get_encode();
  if (encode == nullptr) {
    set_error("%s: cuTensorMapEncodeTiled is not available from the driver", who);
    return YNET_E_UNSUPPORTED;
  }
  CUtensorMap maps[3];
  memset(maps, 0, sizeof(maps));
  int chunks = 0, stored_total = 0;
  for (int i = 0; i < n_src; ++i) {
    const ynet_tc_src& sc = srcs[i];
    const int cp = sc.channels_pad;
    if (!sc.ptr || reinterpret_cast<uintptr_t>(sc.ptr) % 16 != 0 || !(cp > 0 && cp % 16 == 0) || sc.padded ||
        sc.center_only || sc.tap_mask) {
      set_error("%s: source %d must be a plain 16-byte aligned C8 tensor with channels_pad a multiple of 16", who, i);
      return YNET_E_INVALID;
    }
    const bool bcast = sc.batch_stride == 0;
    const int bmod = sc.batch_mod;
    const int nsrc = bcast ? 1 : (bmod > 0 ? bmod : (bmod < 0 ? ceil_div(N, -bmod) : N));
    const int stored = (sc.chunks_stored > 0) ? sc.chunks_stored : cp / 8;
    if (stored > cp / 8) {
      set_error("%s: source %d stores more planes than channels_pad / 8", who, i);
      return YNET_E_INVALID;
    }
    // The planes as 8-byte elements, two per pixel: a box line may then span 128 pixels (256 elements) from any pixel.
    const cuuint64_t dims[4] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)stored, (cuuint64_t)nsrc};
    const cuuint64_t bs = bcast ? (cuuint64_t)stored * H * W * 16 : (cuuint64_t)sc.batch_stride * 2;
    if (bs % 16 != 0) {
      set_error("%s: batch stride must be a multiple of 8 elements", who);
      return YNET_E_ALIGN;
    }
    const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, bs};
    const cuuint32_t box[4] = {(cuuint32_t)RC_M * 2, 1, (cuuint32_t)stored, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(sc.ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("%s: cuTensorMapEncodeTiled failed (%d) for source %d (W=%d H=%d C=%d)", who, (int)r, i, W, H, cp);
      return YNET_E_CUDA;
    }
    p.src_stored[i] = stored;
    p.src_coff[i] = chunks;
    p.src_bcast[i] = bcast ? 1 : 0;
    p.src_mod[i] = bmod;
    chunks += cp / 8;
    stored_total += stored;
  }
  for (int i = n_src; i < 3; ++i) maps[i] = maps[0];
  CUtensorMap mapw = maps[0], mapw120 = maps[0], mapz = maps[0];
  if (wp != nullptr) {
    if ((fuse && two == nullptr) || !(wp->tmpl_c8 && wp->coords) || reinterpret_cast<uintptr_t>(wp->tmpl_c8) % 16 != 0 || wp->n_ch < 1 ||
        wp->n_ch > 2 || wp->level < 0 || wp->level > 1 || n_src > 2 || wp->th < (H << wp->level) ||
        wp->tw < (W << wp->level) || (wp->level == 1 && ((wp->th | wp->tw) & 1)) || W < 120) {
      set_error("%s: waypoint source: 1 or 2 channels, level 0 or 1, (even-sized) template at least as large as the "
                "full-resolution image, W >= 120", who);
      return YNET_E_INVALID;
    }
    static void* zeros = nullptr;              // 128 bytes of zeros: the conv's padding pixels of the edge strips
    if (zeros == nullptr) {
      cudaError_t ze = cudaMalloc(&zeros, 256);
      if (ze == cudaSuccess) ze = cudaMemset(zeros, 0, 256);
      if (ze != cudaSuccess) return cuda_fail(ze, who);
    }
    // level 0: planes (1, th, tw); level 1: the four parity planes (4, th / 2, tw / 2) of ynet_tc_wp_template_c8
    const cuuint64_t tH = (cuuint64_t)(wp->th >> wp->level), tW = (cuuint64_t)(wp->tw >> wp->level);
    const cuuint64_t dims[4] = {tW * 2, tH, wp->level ? 4u : 1u, 1};
    const cuuint64_t strides[3] = {tW * 16, tH * tW * 16, (wp->level ? 4u : 1u) * tH * tW * 16};
    const cuuint32_t box[4] = {(cuuint32_t)RC_M * 2, 1, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&mapw, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(wp->tmpl_c8), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const cuuint32_t box120[4] = {120 * 2, 1, 1, 1};
    if (r == CUDA_SUCCESS)
      r = encode(&mapw120, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(wp->tmpl_c8), dims, strides, box120, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const cuuint64_t zdims[4] = {16, 1, 1, 1};
    const cuuint64_t zstr[3] = {128, 128, 128};
    const cuuint32_t zbox[4] = {16, 1, 1, 1};
    if (r == CUDA_SUCCESS)
      r = encode(&mapz, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, zeros, zdims, zstr, zbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("%s: cuTensorMapEncodeTiled failed (%d) for the template planes (%d x %d)", who, (int)r, wp->th, wp->tw);
      return YNET_E_CUDA;
    }
    p.wp_coords = wp->coords;
    p.wp_nch = wp->n_ch;
    p.wp_level = wp->level;
    p.wp_th = wp->th;
    p.wp_tw = wp->tw;
    p.wp_coff = chunks;
    chunks += 2;                 // one 16-channel K block: channel c = K index 8 c (chunk wp_coff + c), the rest zero
    stored_total += wp->n_ch;
  }
  if (chunks / 2 > RC_MAX_KB) {
    set_error("%s: more than %d input channels", who, RC_MAX_KB * 16);
    return YNET_E_UNSUPPORTED;
  }
  p.n_src = n_src;
  p.zero_fill = (stored_total != chunks) ? 1 : 0;
  p.row_tx = stored_total * RC_CHUNK_BYTES;
  if (partial != nullptr) {
    const int pc = partial->channels_pad / 8;
    if (!partial->ptr || reinterpret_cast<uintptr_t>(partial->ptr) % 16 != 0 || !(pc == 4 || pc == 8) ||
        partial->batch_stride <= 0 || (fuse && two == nullptr)) {
      set_error("%s: the partial-sum source must be a 32- or 64-channel (hi | lo) C8 tensor", who);
      return YNET_E_INVALID;
    }
    p.part = reinterpret_cast<const uint4*>(partial->ptr);
    p.part_bs = partial->batch_stride / 8;
    p.part_mod = partial->batch_mod;
    p.part_chunks = pc;
  }
  if (const char* e = getenv("YNET_RC_DBG")) p.dbg = atoi(e);
  p.N = N;
  p.H = H;
  p.W = W;
  p.kb = chunks / 2;
  p.chunks = chunks;
  p.strips = (wp != nullptr) ? 2 + ceil_div(tmax(0, W - 238), RC_VW) : ceil_div(W, RC_VW);     // WP: see rc_wp_window
  if (two != nullptr) p.strips = 2 + ceil_div(tmax(0, W - 236), R2_VW);                        // see rc2_window
  p.items = (long long)N * p.strips;
  p.row_bytes = p.chunks * RC_CHUNK_BYTES;
  if (two != nullptr) {
    if (wp == nullptr || n_src > 2) {
      set_error("%s: the two-conv kernel takes one or two tensor sources and the waypoint source", who);
      return YNET_E_INVALID;
    }
    Rc2Params pp;
    pp.a = p;
    pp.wb96 = reinterpret_cast<const unsigned char*>(two->wb96);
    pp.bias_b = two->bias_b;
    pp.relu_b = two->relu_b;
    CUtensorMap mapp = maps[0];
    if (partial != nullptr) {
      if (p.part_chunks != 4) {
        set_error("%s: the two-conv kernel takes hi-only partial sums (32 channels)", who);
        return YNET_E_UNSUPPORTED;
      }
      const int bmod = partial->batch_mod;
      const int nimg = bmod > 0 ? bmod : (bmod < 0 ? ceil_div(N, -bmod) : N);
      const cuuint64_t dims[4] = {(cuuint64_t)W * 2, (cuuint64_t)H, 4, (cuuint64_t)nimg};
      const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)partial->batch_stride * 2};
      const cuuint32_t box[4] = {(cuuint32_t)RC_M * 2, 1, 4, 1};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&mapp, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(partial->ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed (%d) for the partial sums", who, (int)r);
        return YNET_E_CUDA;
      }
    }
    const size_t smem2 = (size_t)p.kb * 3 * RC_WBLK + 2 * 3 * RC_WBLK + (size_t)R2_NSTG * p.row_bytes + 1024 +
                         (size_t)R2_NA * RC_YROW + 1024 + (size_t)(two->tail ? 8 : 4) * R2_PROW +
                         (two->tail ? (size_t)R2_NY * RC_YROW + 2 * PR_WBLK_BYTES : 0) + 64 * 4 + 80 * 8 + 16 + 1024;
    static bool configured2 = false;
    if (!configured2) {
      cudaError_t e = cudaFuncSetAttribute(tc_rowconv2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(tc_rowconv2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, who);
      configured2 = true;
    }
    if (smem2 > 226 * 1024) {
      set_error("%s: layer does not fit shared memory", who);
      return YNET_E_UNSUPPORTED;
    }
    const long long grid2 = tmin<long long>(p.items, (long long)sm_count());
    if (two->tail)
      tc_rowconv2_kernel<true><<<(unsigned)grid2, R2_THREADS_TAIL, smem2, as_stream(stream)>>>(maps[0], maps[1], mapw, mapw120,
                                                                                              mapz, mapp, pp);
    else
      tc_rowconv2_kernel<false><<<(unsigned)grid2, R2_THREADS_PLAIN, smem2, as_stream(stream)>>>(maps[0], maps[1], mapw,
                                                                                                mapw120, mapz, mapp, pp);
    cudaError_t le2 = cudaGetLastError();
    if (le2 != cudaSuccess) return cuda_fail(le2, who);
    return YNET_OK;
  }
  const size_t smem = (size_t)p.kb * 3 * RC_WBLK + (size_t)(fuse ? RC_STAGES_FUSED : RC_STAGES_PLAIN) * p.row_bytes + 1024 +
                      (fuse ? (size_t)RC_NY * RC_YROW + 2 * PR_WBLK_BYTES : 0) + 80 * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_rowconv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tc_rowconv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tc_rowconv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, who);
    configured = true;
  }
  if (smem > 200 * 1024) {
    set_error("%s: layer does not fit shared memory", who);
    return YNET_E_UNSUPPORTED;
  }
  // plain: 256 TMEM columns and < 110 KB per CTA -> two co-resident CTAs per SM; fused: one (512 columns)
  const int per_sm = fuse ? 1 : ((smem + 1024 <= 110 * 1024) ? 2 : 1);
  const long long grid = tmin<long long>(p.items, (long long)sm_count() * per_sm);
  cudaStream_t st = as_stream(stream);
  if (fuse)
    tc_rowconv_kernel<true><<<(unsigned)grid, RC_THREADS_FUSED, smem, st>>>(maps[0], maps[1], maps[2], mapw, mapw120, mapz, p);
  else
    if (wp != nullptr)
      tc_rowconv_kernel<false, true><<<(unsigned)grid, RC_THREADS_PLAIN, smem, st>>>(maps[0], maps[1], maps[2], mapw, mapw120, mapz, p);
    else
      tc_rowconv_kernel<false><<<(unsigned)grid, RC_THREADS_PLAIN, smem, st>>>(maps[0], maps[1], maps[2], mapw, mapw120, mapz, p);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return cuda_fail(le, who);
  return YNET_OK;
}

}  // namespace ynet

using namespace ynet;

extern "C" {

int64_t ynet_tc_rowconv_packed_weight_bytes(int32_t C_in_pad) {
  if (C_in_pad <= 0 || C_in_pad % 16 != 0 || C_in_pad / 16 > RC_MAX_KB) return 0;
  return (int64_t)(C_in_pad / 16) * 3 * RC_WBLK;
}

int ynet_tc_rowconv_pack_weights(const float* weight, int32_t C_out, int32_t C_in, int32_t C_in_pad, void* packed,
                                 void* stream) {
  YNET_CHECK_ARG(weight && packed, "null pointer");
  YNET_CHECK_ARG(C_out > 0 && C_out <= RC_CO && C_in > 0 && C_in <= C_in_pad && C_in_pad % 16 == 0 &&
                     C_in_pad / 16 <= RC_MAX_KB,
                 "C_out <= 32, C_in <= C_in_pad <= 64");
  YNET_CHECK_ALIGN(packed, 16);
  const int kb = C_in_pad / 16;
  rc_pack_weights_kernel<<<ceil_div(kb * 3 * 2 * 96 * 8, 256), 256, 0, as_stream(stream)>>>(
      weight, C_out, C_in, kb, reinterpret_cast<__nv_bfloat16*>(packed));
  YNET_LAUNCH_CHECK();
  return YNET_OK;
}

int ynet_tc_rowconv3x3(const ynet_tc_src* srcs, int32_t n_src, const ynet_tc_src* partial, int32_t N, int32_t H, int32_t W,
                       const void* packed_weight, const float* bias32, int32_t C_out, int32_t relu, void* out_c8,
                       int32_t C_out_pad, void* stream) {
  YNET_CHECK_ARG(packed_weight && bias32 && (out_c8 || N == 0), "null pointer");
  YNET_CHECK_ARG(C_out > 0 && C_out <= RC_CO && C_out_pad % 16 == 0 && C_out_pad >= C_out && C_out_pad <= RC_CO, "C_out <= 32");
  YNET_CHECK_ALIGN(packed_weight, 16);
  YNET_CHECK_ALIGN(out_c8, 16);
  RcParams p;
  memset(&p, 0, sizeof(p));
  p.w96 = reinterpret_cast<const unsigned char*>(packed_weight);
  p.bias = bias32;
  p.relu = relu & 1;
  p.pad_out = (relu & 2) ? 1 : 0;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_c8);
  p.out_chunks = C_out_pad / 8;
  return rc_launch("ynet_tc_rowconv3x3", false, srcs, n_src, partial, N, H, W, p, stream);
}

int ynet_tc_rowconv3x3_wp(const ynet_tc_src* srcs, int32_t n_src, const ynet_tc_src* partial, const ynet_tc_wp_src* wp,
                          int32_t N, int32_t H, int32_t W, const void* packed_weight, const float* bias32, int32_t C_out,
                          int32_t relu, void* out_c8, int32_t C_out_pad, void* stream) {
  YNET_CHECK_ARG(packed_weight && bias32 && wp && (out_c8 || N == 0), "null pointer");
  YNET_CHECK_ARG(C_out > 0 && C_out <= RC_CO && C_out_pad % 16 == 0 && C_out_pad >= C_out && C_out_pad <= RC_CO, "C_out <= 32");
  YNET_CHECK_ARG(n_src <= 2, "at most two tensor sources next to the waypoint source");
  YNET_CHECK_ALIGN(packed_weight, 16);
  YNET_CHECK_ALIGN(out_c8, 16);
  RcParams p;
  memset(&p, 0, sizeof(p));
  p.w96 = reinterpret_cast<const unsigned char*>(packed_weight);
  p.bias = bias32;
  p.relu = relu & 1;
  p.pad_out = (relu & 2) ? 1 : 0;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_c8);
  p.out_chunks = C_out_pad / 8;
  return rc_launch("ynet_tc_rowconv3x3_wp", false, srcs, n_src, partial, N, H, W, p, stream, wp);
}

int ynet_tc_rowconv2_wp(const ynet_tc_src* srcs, int32_t n_src, const ynet_tc_src* partial, const ynet_tc_wp_src* wp, int32_t N,
                        int32_t H, int32_t W, const void* packed_weight_a, const float* bias32_a, const void* packed_weight_b,
                        const float* bias32_b, int32_t C_out, int32_t relu, void* out_c8, int32_t C_out_pad, void* stream) {
  YNET_CHECK_ARG(packed_weight_a && bias32_a && packed_weight_b && bias32_b && wp && (out_c8 || N == 0), "null pointer");
  YNET_CHECK_ARG(C_out > 0 && C_out <= RC_CO && C_out_pad % 16 == 0 && C_out_pad >= C_out && C_out_pad <= RC_CO, "C_out <= 32");
  YNET_CHECK_ALIGN(packed_weight_a, 16);
  YNET_CHECK_ALIGN(packed_weight_b, 16);
  YNET_CHECK_ALIGN(out_c8, 16);
  RcParams p;
  memset(&p, 0, sizeof(p));
  p.w96 = reinterpret_cast<const unsigned char*>(packed_weight_a);
  p.bias = bias32_a;
  p.relu = 1;
  p.pad_out = (relu & 2) ? 1 : 0;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_c8);
  p.out_chunks = C_out_pad / 8;
  const Rc2Host two{packed_weight_b, bias32_b, relu & 1, false};
  return rc_launch("ynet_tc_rowconv2_wp", false, srcs, n_src, partial, N, H, W, p, stream, wp, &two);
}

int64_t ynet_tc_rowconv2_softargmax_workspace_bytes(int32_t N, int32_t C_pred, int32_t W) {
  if (N <= 0 || C_pred <= 0 || W < 120) return 0;
  return (int64_t)N * C_pred * (2 + ceil_div(tmax(0, W - 236), R2_VW)) * R2_SOFT_WARPS * (int64_t)sizeof(float4);
}

int ynet_tc_rowconv2_wp_pred_softargmax(const ynet_tc_src* srcs, int32_t n_src, const ynet_tc_src* partial,
                                        const ynet_tc_wp_src* wp, int32_t N, int32_t H, int32_t W, const void* packed_weight_a,
                                        const float* bias32_a, const void* packed_weight_b, const float* bias32_b,
                                        int32_t relu_b, const void* packed_pred_weight, const float* pred_bias, int32_t C_pred,
                                        float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  YNET_CHECK_ARG(packed_weight_a && bias32_a && packed_weight_b && bias32_b && packed_pred_weight && pred_bias && wp &&
                     (out || N == 0),
                 "null pointer");
  YNET_CHECK_ARG(C_pred > 0 && C_pred <= 32, "C_pred <= 32");
  YNET_CHECK_ALIGN(packed_weight_a, 16);
  YNET_CHECK_ALIGN(packed_weight_b, 16);
  YNET_CHECK_ALIGN(packed_pred_weight, 16);
  if (N == 0) return YNET_OK;
  if (workspace == nullptr || workspace_bytes < ynet_tc_rowconv2_softargmax_workspace_bytes(N, C_pred, W) || W < 120) {
    set_error("ynet_tc_rowconv2_wp_pred_softargmax: workspace too small (or W < 120)");
    return YNET_E_WORKSPACE;
  }
  YNET_CHECK_ALIGN(workspace, 16);
  RcParams p;
  memset(&p, 0, sizeof(p));
  p.w96 = reinterpret_cast<const unsigned char*>(packed_weight_a);
  p.bias = bias32_a;
  p.relu = 1;
  p.pw = reinterpret_cast<const unsigned char*>(packed_pred_weight);
  p.pbias = pred_bias;
  p.partial = reinterpret_cast<float4*>(workspace);
  p.c_pred = C_pred;
  p.pn_pad = ceil_div(C_pred, 16) * 16;
  p.kbp = 2;
  p.slots = (2 + ceil_div(tmax(0, W - 236), R2_VW)) * R2_SOFT_WARPS;
  const Rc2Host two{packed_weight_b, bias32_b, relu_b & 1, true};
  int rc = rc_launch("ynet_tc_rowconv2_wp_pred_softargmax", true, srcs, n_src, partial, N, H, W, p, stream, wp, &two);
  if (rc != YNET_OK) return rc;
  cudaError_t le = pred_partial_finalize(p.partial, N * C_pred, p.slots, out, as_stream(stream));
  if (le != cudaSuccess) return cuda_fail(le, "ynet_tc_rowconv2_wp_pred_softargmax");
  return YNET_OK;
}

int64_t ynet_tc_rowconv_softargmax_workspace_bytes(int32_t N, int32_t C_pred, int32_t W) {
  if (N <= 0 || C_pred <= 0 || W <= 0) return 0;
  return (int64_t)N * C_pred * ceil_div(W, RC_VW) * RC_SOFT_WARPS * (int64_t)sizeof(float4);
}

int ynet_tc_rowconv3x3_pred_softargmax(const ynet_tc_src* src, int32_t N, int32_t H, int32_t W, const void* packed_weight,
                                       const float* bias32, int32_t C_out, int32_t relu, const void* packed_pred_weight,
                                       const float* pred_bias, int32_t C_pred, float* out, void* workspace,
                                       int64_t workspace_bytes, void* stream) {
  YNET_CHECK_ARG(packed_weight && bias32 && packed_pred_weight && pred_bias && (out || N == 0), "null pointer");
  YNET_CHECK_ARG(C_pred > 0 && C_pred <= 32 && C_out > 0 && C_out <= RC_CO, "C_pred <= 32, C_out <= 32");
  YNET_CHECK_ALIGN(packed_weight, 16);
  YNET_CHECK_ALIGN(packed_pred_weight, 16);
  if (N == 0) return YNET_OK;
  if (workspace == nullptr || workspace_bytes < ynet_tc_rowconv_softargmax_workspace_bytes(N, C_pred, W)) {
    set_error("ynet_tc_rowconv3x3_pred_softargmax: workspace too small");
    return YNET_E_WORKSPACE;
  }
  YNET_CHECK_ALIGN(workspace, 16);
  RcParams p;
  memset(&p, 0, sizeof(p));
  p.w96 = reinterpret_cast<const unsigned char*>(packed_weight);
  p.bias = bias32;
  p.relu = relu & 1;
  p.pw = reinterpret_cast<const unsigned char*>(packed_pred_weight);
  p.pbias = pred_bias;
  p.partial = reinterpret_cast<float4*>(workspace);
  p.c_pred = C_pred;
  p.pn_pad = ceil_div(C_pred, 16) * 16;
  p.kbp = ceil_div(C_out, 16);
  p.slots = ceil_div(W, RC_VW) * RC_SOFT_WARPS;
  int rc = rc_launch("ynet_tc_rowconv3x3_pred_softargmax", true, src, 1, nullptr, N, H, W, p, stream);
  if (rc != YNET_OK) return rc;
  cudaError_t le = pred_partial_finalize(p.partial, N * C_pred, p.slots, out, as_stream(stream));
  if (le != cudaSuccess) return cuda_fail(le, "ynet_tc_rowconv3x3_pred_softargmax");
  return YNET_OK;
}

}  // extern "C"
