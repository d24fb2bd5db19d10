/*
 * ynet_b200.h -- C ABI of libynet_b200.so: the B200 (sm_100a) implementation of the Y-Net
 * scene-heatmap forecasting hot path of vita-epfl/motion-style-transfer.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; its seam for this path is the
 * method set of models/ynet.py::YNet (ynet.py:551-600) plus the free functions get_patch, sampling,
 * kmeans, SoftArgmax2D and torch_multivariate_gaussian_heatmap.  Each entry point below cites the
 * reference lines it replaces (paths relative to the reference tree).  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all buffers, including workspaces (sizes via the *_workspace_bytes queries);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous w.r.t. the host;
 *   - return value: 0 = success, negative = YNET_E_* (message via ynet_last_error_string());
 *   - activations are float32 NCHW unless a function says otherwise; weights are float32 OIHW
 *     exactly as stored in the reference's state_dict.
 */
#ifndef YNET_B200_H_
#define YNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YNET_OK 0
#define YNET_E_INVALID (-1)   /* bad shape / argument                                        */
#define YNET_E_ALIGN (-2)     /* pointer or extent not aligned as required                   */
#define YNET_E_ARCH (-3)      /* device is not sm_100                                         */
#define YNET_E_CUDA (-4)      /* CUDA runtime / driver error (see ynet_last_error_string)    */
#define YNET_E_WORKSPACE (-5) /* workspace too small                                          */
#define YNET_E_UNSUPPORTED (-6)

#define YNET_MAX_SOURCES 4

/* How a conv source is read (fuses nn.MaxPool2d / F.interpolate / torch.cat into the loader). */
#define YNET_SRC_DIRECT 0 /* source has the conv's spatial size                               */
#define YNET_SRC_POOL2 1  /* source is 2H x 2W; loader takes the 2x2 max   (ynet.py:202,215)  */
#define YNET_SRC_UP2 2    /* source is H/2 x W/2; loader does bilinear x2, align_corners=False
                             (ynet.py:463)                                                     */

typedef struct ynet_conv_src {
  const void* ptr;      /* float32 NCHW                                                       */
  int32_t channels;     /* channels contributed by this source (concat order = array order)   */
  int32_t mode;         /* YNET_SRC_*                                                          */
  int64_t batch_stride; /* elements between consecutive images; 0 = broadcast (Tensor.expand,
                           evaluate.py:117)                                                    */
  int64_t batch_mod;    /* > 0: image n reads source image n % batch_mod (the encoder features
                           are shared by the n_goal trajectory-decoder passes, evaluate.py:259) */
} ynet_conv_src;

int ynet_version(void);
const char* ynet_last_error_string(void);
/* Fills sm count / compute capability of the current device; YNET_E_ARCH if it is not sm_100. */
int ynet_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * a3  rasterisation: get_patch + torch.stack  (utils/image_utils.py:40-63, evaluate.py:112-114,
 *     250-253, train_epoch.py:63-78).  out[n,i,j] = tmpl[mid_y - y_n + i, mid_x - x_n + j] with
 *     (x_n, y_n) = round-half-even(coords[n]).  Bit-exact copy of template values.
 *     coords: (n, 2) float32 (x, y).  out: (n, H, W) float32.  oob_flag (optional int32): set to 1
 *     if any window leaves the template (the reference would silently mis-slice).
 * ------------------------------------------------------------------------------------------- */
int ynet_rasterize_patches(const float* tmpl, int32_t tmpl_h, int32_t tmpl_w, const float* coords,
                           int32_t n, float* out, int32_t H, int32_t W, int32_t* oob_flag, void* stream);

/* Analytic variant for the distance template of create_dist_mat(size) (image_utils.py:30-37):
 * float32(sqrt(double(di^2+dj^2)) / sqrt(double(2 mid^2)) * 2), bit-identical to the gather. */
int ynet_rasterize_dist_analytic(int32_t tmpl_size, const float* coords, int32_t n, float* out, int32_t H,
                                 int32_t W, void* stream);

/* a1  create_dist_mat(size) as float32 (trainer.py:209,326). out: (size, size). */
int ynet_create_dist_template(int32_t size, float* out, void* stream);

/* a9  waypoint pyramid: nn.AvgPool2d(2^i), i = 1..n_levels-1, of the full-resolution map
 *     (evaluate.py:255-257, train_epoch.py:97-100).  in: (n, H, W); outs_host[i-1]: (n, H>>i, W>>i). */
int ynet_avgpool_pyramid(const float* in, int32_t n, int32_t H, int32_t W, int32_t n_levels,
                         float* const* outs_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a12 SoftArgmax2D.forward (utils/softargmax.py:55-81): rows = B*C maps of H x W ->
 *     out (rows, 2) = (x, y); eps = 1e-6 added to the exp-sum after max subtraction.
 *     row r starts at x + r * row_stride (row_stride = H*W for a dense (B*C, H, W) block; C*H*W
 *     to read one channel of every image, evaluate.py:142-143).
 *     workspace: ynet_softargmax2d_workspace_bytes(rows, H, W).
 * a13 YNet.softmax (ynet.py:578-579) and softargmax_on_softmax_map (ynet.py:588-600).
 * ------------------------------------------------------------------------------------------- */
int64_t ynet_softargmax2d_workspace_bytes(int32_t rows, int32_t H, int32_t W);
int ynet_softargmax2d(const float* x, int32_t rows, int64_t row_stride, int32_t H, int32_t W, float* out,
                      void* workspace, int64_t workspace_bytes, void* stream);
int ynet_spatial_softmax(const float* x, int32_t rows, int64_t S, float* out, void* stream);
int ynet_expectation2d(const float* p, int32_t rows, int32_t H, int32_t W, float* out, void* stream);

/* a10 YNet.sigmoid with temperature on selected channels (evaluate.py:128-131):
 *     out[b, k] = sigmoid(logits[b, ch[k]] / T).  logits (B, C, S); ch_host: n_ch host ints. */
int ynet_sigmoid_select(const float* logits, int32_t B, int32_t C, int64_t S, const int32_t* ch_host,
                        int32_t n_ch, float temperature, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a11 sampling (utils/image_utils.py:110-135) = threshold/normalise + torch.multinomial + unravel.
 *
 * ynet_sampling_prepare: per-row max and the masked GLOBAL sum (image_utils.py:114-119); the sum
 *     is float64-accumulated in a fixed order, rounded to float32.  rowmax (rows), gsum (1 float).
 * ynet_multinomial_replacement: ATen CPU multinomial, replacement=True: sequential float32 CDF,
 *     c /= c[S-1], c[S-1] = 1, idx = first j with double(c[j]) >= u.  Bit-exact.
 *     prob (rows, S); rel_threshold < 0 disables the mask/normalise step (then rowmax/gsum unused);
 *     uniforms (rows, n) float64; cdf_ws (rows, S) float32 scratch; idx (rows, n) int64;
 *     xy (rows, n, 2) float32 = (idx % W, floor(idx / W)) or NULL.
 * ynet_multinomial_topk: replacement=False or n == 1: top-n of p / q by descending value
 *     (q ~ Exp(1) supplied, float32 (rows, S)); ties -> lowest index.  Bit-exact.
 * ------------------------------------------------------------------------------------------- */
int64_t ynet_sampling_prepare_workspace_bytes(int32_t rows, int64_t S);
int ynet_sampling_prepare(const float* prob, int32_t rows, int64_t S, float rel_threshold, float* rowmax,
                          float* gsum, void* workspace, int64_t workspace_bytes, void* stream);
int ynet_multinomial_replacement(const float* prob, int32_t rows, int64_t S, float rel_threshold,
                                 const float* rowmax, const float* gsum, const double* uniforms, int32_t n,
                                 float* cdf_ws, int64_t* idx, float* xy, int32_t W, void* stream);
int ynet_multinomial_topk(const float* prob, const float* expo, int32_t rows, int64_t S, float rel_threshold,
                          const float* rowmax, const float* gsum, int32_t n, int64_t* idx, float* xy,
                          int32_t W, void* stream);
/* Counter-based device generators for the production path (the parity path supplies randoms). */
/* `epoch` (device uint64, may be NULL) is mixed into the seed on the device so that a captured CUDA graph draws
 * fresh numbers on every replay; ynet_counter_add advances it. */
int ynet_counter_add(uint64_t* counter, uint64_t inc, void* stream);
int ynet_rng_uniform_f64(uint64_t seed, const uint64_t* epoch, uint64_t offset, int64_t n, double* out, void* stream);
int ynet_rng_exponential_f32(uint64_t seed, const uint64_t* epoch, uint64_t offset, int64_t n, float* out,
                             void* stream);
/* K distinct indices in [0, N) per row: device analogue of np.random.choice(N, K, replace=False)
 * (utils/kmeans.py:17).  out (rows, K) int32. */
int ynet_rng_choice(uint64_t seed, const uint64_t* epoch, uint64_t offset, int32_t rows, int32_t N, int32_t K,
                    int32_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a14 kmeans (utils/kmeans.py:22-108, euclidean), batched over agents (the reference loops in
 *     Python, evaluate.py:147-155).  X (B, N, 2) float32; init_idx (B, K) int32 (np.random.choice,
 *     kmeans.py:17); reseed_idx (B, R) int32 consumed in (iteration, cluster) order for empty
 *     clusters (torch.randint, kmeans.py:83) or NULL; stop when shift^2 < tol or iter_limit.
 *     centres (B, K, 2); assign (B, N) int32 or NULL; iters (B) int32 or NULL;
 *     status (B) int32 or NULL: bit0 = reseed stream exhausted.
 * ------------------------------------------------------------------------------------------- */
int ynet_kmeans_batched(const float* X, int32_t B, int32_t N, int32_t K, const int32_t* init_idx,
                        const int32_t* reseed_idx, int32_t R, float tol, int32_t iter_limit, float* centres,
                        int32_t* assign, int32_t* iters, int32_t* status, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a16 CWS (utils/evaluate.py:9-34, 172-224), one waypoint level for all goals and agents:
 *     prior_g,b = oriented Gaussian of torch_multivariate_gaussian_heatmap; returns the
 *     softargmax_on_softmax_map of sigmoid_map[b] * prior (first trajectory per goal).
 *     sig (B, H, W) = sigmoid map of this waypoint; wp_in (G, B, 2) current waypoint (x, y);
 *     last_obs (B, 2); length_ratio = 1/(waypoint_num+2); sigma_factor (G) float per goal
 *     (sigma_factor - traj_idx); out (G, B, 2).
 * ------------------------------------------------------------------------------------------- */
int64_t ynet_cws_waypoint_workspace_bytes(int32_t B, int32_t G);
int ynet_cws_waypoint(const float* sig, int32_t B, int32_t H, int32_t W, const float* wp_in, int32_t G,
                      const float* last_obs, float length_ratio, const float* sigma_factor, float ratio,
                      int32_t rot, float* out, void* workspace, int64_t workspace_bytes, void* stream);
/* Materialises the normalised waypoint map (for the n_traj > 1 re-sampling, evaluate.py:213-216):
 *     out (B, H, W) for ONE goal g. */
int ynet_cws_waypoint_map(const float* sig, int32_t B, int32_t H, int32_t W, const float* wp_in_g,
                          const float* last_obs, float length_ratio, float sigma_factor, float ratio,
                          int32_t rot, float* out, void* stream);

/* a17 ADE / FDE (evaluate.py:276-277, 290-291): gt (B, T, 2); trajs (K, B, T, 2); wps (K, B, n_wp, 2).
 *     ade (B), fde (B): min over K of mean_t ||.||/resize and ||gt_goal - wp_last||/resize. */
int ynet_ade_fde(const float* gt, const float* trajs, const float* wps, int32_t K, int32_t B, int32_t T,
                 int32_t n_wp, float resize_factor, float* ade, float* fde, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4-a8 network, reference-grade float32 engine (CUDA cores, fp32 accumulate).
 *
 * ynet_conv3x3_f32: F.conv2d(cat(sources), W, b, padding=1) [+ ReLU]  (ynet.py:192-211, 419-447)
 *     with the producer ops fused into the loader (YNET_SRC_*).  weight_packed is the layout-1
 *     output of ynet_lora_fold: [sum C_s][3*3][C_out] (OIHW transposed so that CTAs stage it coalesced).
 * ynet_conv1x1_f32: the predictor (ynet.py:450-451,469).
 * ynet_predictor_softargmax_f32: predictor + SoftArgmax2D in one pass (ynet.py:469 + 582-583):
 *     x (N, C_in, H, W) -> out (N, C_out, 2); logits are never written to HBM.
 * ynet_lora_fold: loralib 0.1.1 Conv2d.forward weight (ynet.py:143): W + (B @ A).view(W.shape)/r
 *     (lora_A / lora_B may be NULL = plain conv).  out_layout 0 = OIHW, 1 = packed [C_in][k*k][C_out].
 * ------------------------------------------------------------------------------------------- */
int ynet_conv3x3_f32(const ynet_conv_src* srcs_host, int32_t n_src, int32_t N, int32_t H, int32_t W,
                     const float* weight_packed, const float* bias, int32_t C_out, int32_t relu, float* out,
                     void* stream);
int ynet_conv1x1_f32(const float* x, int32_t N, int32_t C_in, int64_t S, const float* weight,
                     const float* bias, int32_t C_out, float* out, void* stream);
int64_t ynet_predictor_softargmax_workspace_bytes(int32_t N, int32_t C_out, int32_t H, int32_t W);
int ynet_predictor_softargmax_f32(const float* x, int32_t N, int32_t C_in, int32_t H, int32_t W,
                                  const float* weight, const float* bias, int32_t C_out, float* out,
                                  void* workspace, int64_t workspace_bytes, void* stream);
int ynet_maxpool2x2_f32(const float* x, int64_t planes, int32_t H, int32_t W, float* out, void* stream);
int ynet_upsample_bilinear2x_f32(const float* x, int64_t planes, int32_t H, int32_t W, float* out,
                                 void* stream);
int ynet_lora_fold(const float* weight, const float* lora_A, const float* lora_B, int32_t C_out, int32_t C_in,
                   int32_t ksize, int32_t rank, int32_t out_layout, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4-a8 network, tensor-core engine (tcgen05.mma kind::f16, bf16 operands, fp32 TMEM accumulate).
 *
 * Activations live in HBM as bf16 "C8" planes: [N][C/8][H][W][8] (C padded to a multiple of 16),
 * so that a TMA box of pixels is directly a no-swizzle K-major UMMA operand.
 * ynet_tc_pack_weights: OIHW fp32 -> the per-CTA shared-memory image (bf16 core matrices).
 * ynet_tc_conv3x3: one 3x3 conv (+bias, +ReLU) over up to YNET_MAX_SOURCES C8 sources.
 * See DESIGN.md "Tensor-core conv engine".
 * ------------------------------------------------------------------------------------------- */
typedef struct ynet_tc_src {
  const void* ptr;      /* bf16 C8 planes                                                     */
  int32_t channels_pad; /* multiple of 16                                                     */
  int32_t batch_mod;    /* > 0: image n reads source image n % batch_mod (goal-major stacking);
                           < 0: image n reads source image n / (-batch_mod) (agent-major stacking:
                           the n_goal decoder passes of one agent are consecutive images, so the
                           agent's encoder features are re-read from L2, evaluate.py:259)        */
  int64_t batch_stride; /* elements (bf16); 0 = broadcast                                     */
  int32_t center_only;  /* 1: the source carries hoisted partial sums -- only the centre tap is applied
                           and its K blocks hold ONE tap in the packed weights (pack it with ksize 1) */
  int32_t chunks_stored; /* 0 = channels_pad / 8; else the number of 8-channel planes the tensor really holds
                           (< channels_pad / 8): the missing planes read as zero (TMA out-of-bounds fill) */
  int32_t padded;       /* 1: planes are (H + 2) x (W + 2) with a one-pixel REPLICATED ring around the H x W image
                           (ynet_tc_pad_replicate, or a conv's padded output): input of ynet_tc_upconv3x3 only */
  int32_t tap_mask;     /* 0: all taps (or the centre tap when center_only).  YNET_TC_TAPS_QUAD: the planes hold the
                           2x2 neighbourhood of every pixel (channel (dy*2+dx)*n + c = map_c[y+dy][x+dx], zero outside
                           the image; ynet_tc_rasterize_pyramid_c8 with quad_levels) so that the four taps anchored at
                           (-1,-1) (-1,0) (0,-1) (0,0) cover the 3x3 window: 4 MMAs per K block instead of 9.  The K
                           blocks hold FOUR taps in the packed weights (pack a (C_out, C, 2, 2) weight, ksize 2)      */
} ynet_tc_src;
#define YNET_TC_TAPS_QUAD 0x1B

int ynet_tc_supported(void);

/* a3 + a9 fused for the bf16 engine: get_patch of n_img x n_ch coordinates (image_utils.py:40-63) and the
 * AvgPool2d(2^i) pyramid of the result (evaluate.py:255-257), written straight as bf16 C8 planes.
 * coords: (n_img * n_ch, 2) float32 (x, y), n_ch <= 8.  outs_host[l], l < n_levels: (n_img, C_pad/8, H>>l, W>>l, 8)
 * bf16; channel c < n_ch of chunk 0 holds the map, the other channels of chunk 0 are written as zero.
 * write_pad = 0: chunks >= 1 are NOT written (the caller keeps them zero across calls); 1: zero-filled.
 * quad_levels in [0, 2] (n_ch <= 2, C_pad == 8): levels l < quad_levels are written as 2x2-neighbourhood planes --
 * channel (dy*2 + dx) * n_ch + c of pixel (y, x) = map_c[y + dy][x + dx], zero outside the image -- for conv sources
 * with tap_mask = YNET_TC_TAPS_QUAD (4 MMAs per K block instead of 9 at the same 16 B per pixel). */
int ynet_tc_rasterize_pyramid_c8(const float* tmpl, int32_t tmpl_h, int32_t tmpl_w, const float* coords,
                                 int32_t n_img, int32_t n_ch, int32_t H, int32_t W, int32_t n_levels,
                                 void* const* outs_host, int32_t C_pad, int32_t write_pad, int32_t quad_levels,
                                 int32_t* oob_flag, void* stream);
/* The same maps in im2col form for level 0 (full resolution) or 1 (2x2 average pool): channel c*9 + kh*3 + kw of
 * pixel (y, x) = map_c[y+kh-1][x+kw-1], 0 outside the image.  A conv reads it as a `center_only` source with the
 * 1x1 weights W[:, wp channels].reshape(C_out, 9 n_ch): one MMA per 16 im2col channels instead of nine per
 * (mostly empty) 16-channel K block.  out_c8: (n_img, C_pad/8, H>>level, W>>level, 8), C_pad >= 9 n_ch; n_ch <= 3. */
int ynet_tc_rasterize_im2col_c8(const float* tmpl, int32_t tmpl_h, int32_t tmpl_w, const float* coords, int32_t n_img,
                                int32_t n_ch, int32_t H, int32_t W, int32_t level, void* out_c8, int32_t C_pad,
                                void* stream);
int ynet_tc_pack_f32_to_c8(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, int64_t batch_stride,
                           void* out_c8, int32_t C_pad, void* stream);
int ynet_tc_unpack_c8_to_f32(const void* x_c8, int32_t N, int32_t C, int32_t C_pad, int32_t H, int32_t W,
                             float* out, void* stream);
/* nn.MaxPool2d(2,2) / bilinear x2 (align_corners=False) on C8 planes; (H, W) = input size. */
int ynet_tc_maxpool2x2(const void* x_c8, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out_c8, void* stream);
int ynet_tc_upsample2x(const void* x_c8, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out_c8, void* stream);
/* 1x1 predictor (ynet.py:450-451,469) reading C8 bf16, writing float32 NCHW logits (fp32 FMA). */
int ynet_tc_predictor_f32(const void* x_c8, int32_t N, int32_t C_pad, int32_t C_in, int32_t H, int32_t W,
                          const float* weight, const float* bias, int32_t C_out, float* out, void* stream);
/* F.interpolate(scale 2, bilinear, align_corners=False) + upsample_conv (ynet.py:463-464) as ONE low-resolution
 * 3x3 conv with 4 x C_out phase channels and a depth-to-space store: the upsampled tensor is never materialised.
 *   ynet_tc_upconv_phase_weights: weight (C_out, C_in, 3, 3), bias -> w_eff (4*cp, C_in, 3, 3), bias_eff (4*cp),
 *     cp = C_out padded to 16; pack w_eff with ynet_tc_pack_weights(C_out = 4*cp, ksize = 3).
 *   ynet_tc_upconv3x3: srcs are the LOW-resolution (h, w) C8 sources (src_channels_host = their real channel
 *     counts); out_c8: (N, cp/8, 2h, 2w, 8).  `border_weight` (ynet_tc_upconv_border_weights of the original
 *     float32 weight; ynet_tc_upconv_border_weight_bytes bytes) and the original `bias` are used to recompute the
 *     one-pixel border ring exactly (index clamping and zero padding do not commute with the stencil).
 *     When the sources are `padded` (replicated one-pixel ring) the stencil is exact up to the conv's zero padding
 *     and only the outermost high-resolution ring is corrected (3-5 taps per pixel instead of 9 on a 2-pixel ring):
 *     per image, the outside line of each edge is interpolated once into shared memory and the outside taps are
 *     applied with warp-level bf16 mma (fragments appended to `border_weight`); YNET_RINGFIX_MMA=0 selects the
 *     float32 CUDA-core variant.
 *     C_out <= 64; relu must be 0.
 *   ynet_tc_pad_replicate: (N, C_pad/8, H, W, 8) -> (N, C_pad/8, H+2, W+2, 8); ynet_tc_conv3x3 writes the same layout
 *     directly when bit 1 of `relu` is set (relu: bit 0 = ReLU, bit 1 = padded output). */
int ynet_tc_pad_replicate(const void* x_c8, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out_c8, void* stream);
int64_t ynet_tc_upconv_border_weight_bytes(int32_t C_out, int32_t n_src, const int32_t* src_channels_host);
int ynet_tc_upconv_border_weights(const float* weight, int32_t C_out, int32_t n_src, const int32_t* src_channels_host,
                                  float* out, void* stream);
int ynet_tc_upconv_phase_weights(const float* weight, const float* bias, int32_t C_out, int32_t C_in, float* w_eff,
                                 float* bias_eff, void* stream);
int ynet_tc_upconv3x3(const ynet_tc_src* srcs_host, const int32_t* src_channels_host, int32_t n_src, int32_t N,
                      int32_t h, int32_t w, const void* packed_phase_weight, const float* bias_eff,
                      const float* border_weight, const float* bias, int32_t C_out, int32_t relu, void* out_c8,
                      int32_t tune, void* stream);
/* ksize = 3 (3x3 conv), 1 (the 1x1 predictor, centre-only sources) or 2 (2x2-neighbourhood sources, YNET_TC_TAPS_QUAD). */
int64_t ynet_tc_packed_weight_bytes(int32_t C_out, int32_t n_src, const int32_t* src_channels_pad_host,
                                    int32_t ksize);
int ynet_tc_pack_weights(const float* weight, int32_t C_out, int32_t n_src, const int32_t* src_channels_host,
                         const int32_t* src_channels_pad_host, int32_t ksize, void* packed, void* stream);
/* tune: 0 = built-in heuristics; else (accumulators per tile J in bits 0-3) | (max CTAs per SM in bits 4-7) |
 * (pipeline stages in bits 8-15) -- the host side autotunes these per layer shape. */
int ynet_tc_conv3x3(const ynet_tc_src* srcs_host, int32_t n_src, int32_t N, int32_t H, int32_t W,
                    const void* packed_weight, const float* bias, int32_t C_out, int32_t relu, void* out_c8,
                    int32_t C_out_pad, int32_t tune, void* stream);

/* Goal-loop hoisting: the encoder-feature channels of every trajectory-decoder input
 * (evaluate.py:259, ynet.py:466) are identical for the n_goal passes of an agent, so their share of
 * decoder.i.0 / center.0 is computed ONCE per agent by ynet_tc_conv3x3_hilo -- the raw fp32 partial sums
 * (no bias, no ReLU) leave as bf16 -- with_lo = 1: a (hi, lo) pair, out_c8: (N, 2*C_out_pad/8, H, W, 8), channels
 * [hi | lo], ~2^-17 relative; with_lo = 0: hi only, (N, C_out_pad/8, H, W, 8) -- and re-enter the per-goal conv
 * as a `center_only` source with identity weights.
 * Packed weights of a conv with mixed sources = the per-source ynet_tc_pack_weights images (ksize 3 for
 * 3x3 sources, ksize 1 for centre-only ones) concatenated in source order. */
int ynet_tc_conv3x3_hilo(const ynet_tc_src* srcs_host, int32_t n_src, int32_t N, int32_t H, int32_t W,
                         const void* packed_weight, int32_t C_out, void* out_c8, int32_t C_out_pad, int32_t with_lo,
                         int32_t tune, void* stream);

/* Split-bf16 ("bf16x3") engine: the <= 1e-3 parity mode of models/ynet.py:302-470 on the tensor cores (split_tc.cu).
 * A float32 activation x travels as hi = bf16(x), lo = bf16(x - hi) in ONE tensor (N, 2*C_pad/8, H, W, 8): hi planes,
 * then lo planes (C_pad = channels padded to 16).  y = W_hi x_hi + W_hi x_lo + W_lo x_hi (+ bias, ReLU in float32):
 *   ynet_tc_conv3x3_split: per activation TWO ynet_tc_src entries -- {ptr, channels_pad = 2*C_pad} against the packed
 *       [W_hi | W_hi] and {ptr, channels_pad = C_pad, same batch_stride} (the hi planes again) against W_lo -- so the
 *       packed weight is ynet_tc_pack_weights over sources (2*C_pad, C_pad, ...); out_split: (N, 2*C_out_pad/8, H, W, 8).
 *       Concat-on-write (torch.cat of ynet.py:466 / evaluate.py:259 without a copy): out_total_pad > 0 makes out_split a
 *       wider activation (N, 2*out_total_pad/8, H, W, 8) of which this call fills channels [out_channel_off,
 *       out_channel_off + C_out_pad) of both halves; ynet_split_pack_f32 takes the same pair.  0, 0 = a tensor of its own.
 *       The 1x1 predictor (ynet.py:450-451) is ynet_tc_conv1x1_f32 over the same source pairs (float32 NCHW logits).
 *   ynet_split_pack_f32 / ynet_split_unpack_f32: NCHW float32 <-> split planes (batch_stride in floats, 0 = broadcast);
 *   ynet_split_maxpool2x2 (nn.MaxPool2d(2, 2), ynet.py:241-245), ynet_split_upsample2x (F.interpolate(scale_factor=2,
 *       mode='bilinear', align_corners=False), ynet.py:463): read hi + lo, compute in float32, split again. */
int ynet_tc_conv3x3_split(const ynet_tc_src* srcs_host, int32_t n_src, int32_t N, int32_t H, int32_t W,
                          const void* packed_weight, const float* bias, int32_t C_out, int32_t relu, void* out_split,
                          int32_t C_out_pad, int32_t out_total_pad, int32_t out_channel_off, int32_t tune, void* stream);
int ynet_split_pack_f32(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, int64_t batch_stride, void* out_split,
                        int32_t C_pad, int32_t out_total_pad, int32_t out_channel_off, void* stream);
/* dy * (relu_out > 0) -> split planes: the ReLU backward folded into the split of a gradient (contiguous NCHW float32) */
int ynet_split_pack_masked_f32(const float* x, const float* relu_out, int32_t N, int32_t C, int32_t H, int32_t W,
                               void* out_split, int32_t C_pad, void* stream);
int ynet_split_unpack_f32(const void* x_split, int32_t N, int32_t C, int32_t C_pad, int32_t H, int32_t W, float* out,
                          void* stream);
int ynet_split_maxpool2x2(const void* x_split, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out_split, void* stream);
int ynet_split_upsample2x(const void* x_split, int32_t N, int32_t C_pad, int32_t H, int32_t W, void* out_split, void* stream);

/* The 1x1 predictor (ynet.py:450-451,469) on the tensor cores:
 *   ynet_tc_conv1x1_f32        -> float32 NCHW logits (goal decoder: sigmoid / sampling need the map);
 *   ynet_tc_conv1x1_softargmax -> predictor + SoftArgmax2D (ynet.py:582-583, softargmax.py:55-81) in one
 *       kernel (pred_tc.cu): the MMA runs with the weights as the M operand and the pixels as the N
 *       operand, so the accumulator is transposed (TMEM lane = channel, column = pixel) and every epilogue
 *       thread reduces pixels of ONE channel straight from tcgen05.ld -- the logits never reach shared
 *       memory or HBM.  One source (C8, <= 128 channels); out (N, C_out, 2) = (x, y); C_out <= 32.
 *       workspace: ynet_tc_conv1x1_softargmax_workspace_bytes(N, C_out, H, W). */
int ynet_tc_conv1x1_f32(const ynet_tc_src* srcs_host, int32_t n_src, int32_t N, int32_t H, int32_t W,
                        const void* packed_weight, const float* bias, int32_t C_out, float* out, int32_t tune,
                        void* stream);
/* decoder.4.2 + predictor + SoftArgmax2D in ONE kernel (ynet.py:448-451,468-469 + 582-583): the 3x3 conv's bf16
 * output tile stays in shared memory, is multiplied there with the (replicated) predictor weights and reduced by
 * the soft-argmax warps; neither the activation (64 B/pixel out + in) nor the logits reach HBM.  C_out <= 64,
 * C_pred <= 32; packed_pred_weight = ynet_tc_pack_weights(predictor weight, ksize 1, one source of C_out channels);
 * workspace: ynet_tc_conv1x1_softargmax_workspace_bytes(N, C_pred, H, W).  out (N, C_pred, 2) = (x, y). */
int ynet_tc_conv3x3_pred_softargmax(const ynet_tc_src* srcs_host, int32_t n_src, int32_t N, int32_t H, int32_t W,
                                    const void* packed_weight, const float* bias, int32_t C_out, int32_t relu,
                                    const void* packed_pred_weight, const float* pred_bias, int32_t C_pred,
                                    float* out, void* workspace, int64_t workspace_bytes, void* stream);
int64_t ynet_tc_conv1x1_softargmax_workspace_bytes(int32_t N, int32_t C_out, int32_t H, int32_t W);
int ynet_tc_conv1x1_softargmax(const ynet_tc_src* srcs_host, int32_t n_src, int32_t N, int32_t H, int32_t W,
                               const void* packed_weight, const float* bias, int32_t C_out, float* out,
                               void* workspace, int64_t workspace_bytes, int32_t tune, void* stream);

/* Row-marching, kh-stacked 3x3 conv for the C_out <= 32 layers of the decoders' full-resolution levels
 * (ynet.py:466-468: decoder.4.2, decoder.3.2) -- rowconv_tc.cu.  The three kernel rows are stacked along the MMA's N
 * dimension (N = 96) and the accumulators of consecutive output rows form a ring of TMEM slots, so one input row of a
 * 128-pixel strip costs 3 MMAs per 16-channel K block instead of 9 (the N = 32 MMAs of ynet_tc_conv3x3 are bound by
 * the shared-memory operand fetch).  One plain C8 source with channels_pad <= 64; H >= 2.
 *   ynet_tc_rowconv_pack_weights: OIHW float32 (C_out <= 32, C_in, 3, 3), LoRA already folded -> bf16
 *       [K block][kw][2][96 = kh * 32 + co][8]; size = ynet_tc_rowconv_packed_weight_bytes(C_in_pad).
 *   ynet_tc_rowconv3x3: conv + bias + ReLU (relu & 1) -> C8 (N, C_out_pad / 8, H, W, 8); relu & 2: the
 *       replicate-padded (H + 2, W + 2) layout that ynet_tc_upconv3x3 consumes.  bias32: 32 floats, zero beyond C_out.
 *       Up to 3 plain C8 sources = torch.cat along channels (ynet.py:466, evaluate.py:259), <= 64 padded channels in
 *       total, weights packed over the per-source padded concatenation; partial_host (optional): the agent's hoisted
 *       encoder-feature share of this conv (ynet_tc_conv3x3_hilo output, 32 channels hi [+ 32 lo]), added in fp32 by the
 *       epilogue before the activation instead of re-entering through identity-weight MMAs.
 *   ynet_tc_rowconv3x3_pred_softargmax: conv + bias + ReLU -> 1x1 predictor -> SoftArgmax2D (ynet.py:468-469 + 582-583,
 *       softargmax.py:55-81) in ONE kernel: the conv output row stays in shared memory as the N operand of the predictor
 *       MMA, whose transposed accumulator (TMEM lane = channel) is reduced by the soft-argmax warps; neither the
 *       activation nor the logits reach HBM.  packed_pred_weight = ynet_tc_pack_weights(predictor weight, ksize 1, one
 *       source of C_out channels); C_pred <= 32; out (N, C_pred, 2) = (x, y);
 *       workspace: ynet_tc_rowconv_softargmax_workspace_bytes(N, C_pred, W). */
int64_t ynet_tc_rowconv_packed_weight_bytes(int32_t C_in_pad);
int ynet_tc_rowconv_pack_weights(const float* weight, int32_t C_out, int32_t C_in, int32_t C_in_pad, void* packed,
                                 void* stream);
int ynet_tc_rowconv3x3(const ynet_tc_src* srcs_host, int32_t n_src, const ynet_tc_src* partial_host, int32_t N,
                       int32_t H, int32_t W, const void* packed_weight, const float* bias32, int32_t C_out, int32_t relu,
                       void* out_c8, int32_t C_out_pad, void* stream);
int64_t ynet_tc_rowconv_softargmax_workspace_bytes(int32_t N, int32_t C_pred, int32_t W);
/* The same with the waypoint channels of the trajectory decoder's input (evaluate.py:248-259: get_patch of the sampled
 * waypoints, image_utils.py:40-63, and its AvgPool2d(2) level) taken STRAIGHT FROM THE DISTANCE TEMPLATE instead of being
 * rasterised to HBM per image and read back: the kernel's TMA loads the window of every waypoint channel from bf16 C8
 * planes of the template (a few MB, L2-resident, built once per template by ynet_tc_wp_template_c8) and the issuing
 * thread zeroes the conv's padding pixels.  The waypoint source is the LAST 16-channel K block of the conv: channel c
 * (n_ch <= 2) sits at K index 8 c -- pack the weights as two 8-channel sources -- and is bit-identical to the planes
 * ynet_tc_rasterize_pyramid_c8 writes at that level (windows that leave the template read zero instead of clamping;
 * the rasteriser's oob flag reports those).  n_src <= 2 tensor sources precede it.
 *   ynet_tc_wp_template_c8: (th, tw) float32 template -> out_l0 (th, tw, 8) bf16, channel 0 = bf16(T), channels 1-7 zero;
 *       out_l1 (4, th/2, tw/2, 8): plane 2 py + px at [i][j] = bf16(0.25 ((T[2i+py][2j+px] + T[2i+py][2j+px+1]) +
 *       (T[2i+py+1][2j+px] + T[2i+py+1][2j+px+1]))), the AvgPool2d(2) of a window whose corner has parity (py, px). */
typedef struct ynet_tc_wp_src {
  const void* tmpl_c8;   /* out_l0 (level 0) or out_l1 (level 1) of ynet_tc_wp_template_c8                      */
  const float* coords;   /* (N * n_ch, 2) float32 (x, y) at full resolution, rounded half-to-even like np.round */
  int32_t th, tw;        /* size of the float32 template (even for level 1)                                     */
  int32_t n_ch;          /* 1 or 2                                                                              */
  int32_t level;         /* 0: the conv runs at the full resolution; 1: at half of it (2x2 average)             */
} ynet_tc_wp_src;
int ynet_tc_wp_template_c8(const float* tmpl, int32_t th, int32_t tw, void* out_l0, void* out_l1, void* stream);
int ynet_tc_rowconv3x3_wp(const ynet_tc_src* srcs_host, int32_t n_src, const ynet_tc_src* partial_host,
                          const ynet_tc_wp_src* wp_host, int32_t N, int32_t H, int32_t W, const void* packed_weight,
                          const float* bias32, int32_t C_out, int32_t relu, void* out_c8, int32_t C_out_pad, void* stream);
/* A whole decoder block of the trajectory decoder in ONE kernel (ynet.py:466-468: decoder.i.0 + ReLU + decoder.i.2
 * [+ ReLU]), and at the last level the predictor + SoftArgmax2D behind it (ynet.py:469 + 582-583): conv A =
 * ynet_tc_rowconv3x3_wp (tensor sources + waypoint source + hoisted partial sums, 32 output channels, ReLU), whose rows
 * stay in shared memory as the operand of conv B (32 -> C_out <= 32, packed with ynet_tc_rowconv_pack_weights over 32
 * channels).  Neither the 32-channel activation between the convs nor (tail) conv B's output or the logits reach HBM.
 * W >= 120.  relu bits as in ynet_tc_rowconv3x3 (they apply to conv B).  Workspace of the tail:
 * ynet_tc_rowconv2_softargmax_workspace_bytes(N, C_pred, W); out (N, C_pred, 2) = (x, y). */
int ynet_tc_rowconv2_wp(const ynet_tc_src* srcs_host, int32_t n_src, const ynet_tc_src* partial_host,
                        const ynet_tc_wp_src* wp_host, int32_t N, int32_t H, int32_t W, const void* packed_weight_a,
                        const float* bias32_a, const void* packed_weight_b, const float* bias32_b, int32_t C_out,
                        int32_t relu, void* out_c8, int32_t C_out_pad, void* stream);
int64_t ynet_tc_rowconv2_softargmax_workspace_bytes(int32_t N, int32_t C_pred, int32_t W);
int ynet_tc_rowconv2_wp_pred_softargmax(const ynet_tc_src* srcs_host, int32_t n_src, const ynet_tc_src* partial_host,
                                        const ynet_tc_wp_src* wp_host, int32_t N, int32_t H, int32_t W,
                                        const void* packed_weight_a, const float* bias32_a, const void* packed_weight_b,
                                        const float* bias32_b, int32_t relu_b, const void* packed_pred_weight,
                                        const float* pred_bias, int32_t C_pred, float* out, void* workspace,
                                        int64_t workspace_bytes, void* stream);
int ynet_tc_rowconv3x3_pred_softargmax(const ynet_tc_src* src_host, int32_t N, int32_t H, int32_t W,
                                       const void* packed_weight, const float* bias32, int32_t C_out, int32_t relu,
                                       const void* packed_pred_weight, const float* pred_bias, int32_t C_pred,
                                       float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * scene-image preprocessing in front of the path (trainer.py:578-582), one launch per scene.
 * ynet_scene_preprocess_u8: uint8 HWC (3 channels, as cv2.imread returns it) -> cv2.resize(fx = fy = f, INTER_AREA)
 *     (utils/image_utils.py:85-92) -> zero pad at the bottom / right to (Hp, Wp) (95-107) -> smp normalisation
 *     (x / 255 - mean) / std in float64 and HWC -> CHW float32 (66-82).  Bit-exact against OpenCV.  (dh, dw) =
 *     cvRound(size * f); area tables = cv::computeResizeAreaTab for scale 1 / f as CSR arrays (device) -- or int_scale > 0
 *     for an integer scale whose blocks fit (OpenCV's integer path).  out_chw (3, Hp, Wp) and / or out_u8_hwc (dh, dw, 3)
 *     (the resized image alone: the reference-named resize()).
 * ynet_scene_preprocess_oriented_u8 (the image half of utils/data_utils.py::augment_data, 163-233): the
 *     same chain applied to the augmented VIEW cv2.flip(cv2.rotate(img, ROTATE_90_COUNTERCLOCKWISE) x k, 1) with
 *     orient = k + 4 * flip, read straight from the stored (H0, W0) image: the eight views of a scene cost one upload
 *     and no rotated copies.  (dh, dw), the tables and (Hp, Wp) refer to the view (W0 x H0 for odd k).
 * ynet_scene_onehot_u8: segmentation masks: cv2.INTER_NEAREST -> zero pad -> one-hot float32 (classes, Hp, Wp).
 * ------------------------------------------------------------------------------------------- */
int ynet_scene_preprocess_u8(const uint8_t* img_hwc, int32_t H, int32_t W, int32_t dh, int32_t dw, int32_t Hp, int32_t Wp,
                             const int32_t* xt_start, const int32_t* xt_src, const float* xt_w, const int32_t* yt_start,
                             const int32_t* yt_src, const float* yt_w, int32_t int_scale, const double* mean3_host,
                             const double* std3_host, float* out_chw, uint8_t* out_u8_hwc, void* stream);
int ynet_scene_preprocess_oriented_u8(const uint8_t* img_hwc, int32_t H0, int32_t W0, int32_t orient, int32_t dh, int32_t dw,
                                      int32_t Hp, int32_t Wp, const int32_t* xt_start, const int32_t* xt_src,
                                      const float* xt_w, const int32_t* yt_start, const int32_t* yt_src, const float* yt_w,
                                      int32_t int_scale, const double* mean3_host, const double* std3_host, float* out_chw,
                                      uint8_t* out_u8_hwc, void* stream);
int ynet_scene_onehot_u8(const uint8_t* mask, int32_t H, int32_t W, int32_t dh, int32_t dw, int32_t Hp, int32_t Wp,
                         double inv_factor, int32_t classes, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a18 fine-tuning step pieces (utils/train_epoch.py:86-115, models/trainer.py:197-206).
 * ------------------------------------------------------------------------------------------- */
/* BCEWithLogitsLoss(mean) forward + d(loss*scale)/dlogits in one pass.  loss_out: 1 float (mean,
 * unscaled).  grad may be NULL.  workspace: ynet_bce_workspace_bytes(n). */
int64_t ynet_bce_workspace_bytes(int64_t n);
int ynet_bce_logits_fwd_bwd(const float* logits, const float* target, int64_t n, float grad_scale,
                            float* loss_out, float* grad, void* workspace, int64_t workspace_bytes,
                            void* stream);
/* dgrad of conv3x3 (padding 1): dx (N, C_in, H, W) = conv3x3(dy, flip/transpose(W)); dy is first
 * masked by (y > 0) when relu_out != NULL (ReLU backward fused in the loader). */
int ynet_conv3x3_dgrad_f32(const float* dy, const float* relu_out, int32_t N, int32_t H, int32_t W,
                           const float* weight, int32_t C_out, int32_t C_in, float* dx, void* stream);
/* wgrad: dW (C_out, C_in, 3, 3) = sum_n,pixels x (*) dy ; db (C_out). */
int64_t ynet_conv3x3_wgrad_workspace_bytes(int32_t N, int32_t H, int32_t W, int32_t C_out, int32_t C_in);
int ynet_conv3x3_wgrad_f32(const float* x, const float* dy, const float* relu_out, int32_t N, int32_t H,
                           int32_t W, int32_t C_in, int32_t C_out, float* dW, float* db, void* workspace,
                           int64_t workspace_bytes, void* stream);
int ynet_maxpool2x2_bwd_f32(const float* x, const float* dy, int64_t planes, int32_t H, int32_t W, float* dx,
                            void* stream);
int ynet_upsample_bilinear2x_bwd_f32(const float* dy, int64_t planes, int32_t H, int32_t W, float* dx,
                                     void* stream);
/* dA = s * B^T dM, dB = s * dM A^T with dM = dW.view(C_out*k, C_in*k)  (SURVEY 3.2). */
int ynet_lora_grad(const float* dW, const float* lora_A, const float* lora_B, int32_t C_out, int32_t C_in,
                   int32_t ksize, int32_t rank, float* dA, float* dB, void* stream);
/* torch.optim.Adam defaults on one flat buffer; grad_scale folds the 1/world_size average. */
int ynet_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   int32_t step, float lr, float beta1, float beta2, float eps, float grad_scale,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YNET_B200_H_ */
