"""torch-tensor front end of the C ABI (include/ynet_b200.h).

torch is plumbing only: it owns device memory and the stream; every computation below is a call
into libynet_b200.so with raw pointers.  Nothing here runs on the CPU -- CPU tensors are rejected.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ConvSrc, SRC_DIRECT, SRC_POOL2, SRC_UP2, check  # noqa: F401


def _L():
    return _lib.load()


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
_cur_device = getattr(torch._C, '_cuda_getDevice', None)


def _stream():
    """cudaStream_t of torch's current stream.  The raw accessor: building a torch.cuda.Stream object per launch
    (``torch.cuda.current_stream()``) cost 10-17 us each, ~4 ms of the ~390 launches of a fine-tuning step."""
    if _raw_stream is not None and _cur_device is not None:
        return ctypes.c_void_p(_raw_stream(_cur_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _req(t, dtype=torch.float32, name='tensor'):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name}: expected a torch.Tensor, got {type(t)}')
    if not t.is_cuda:
        raise RuntimeError(f'{name}: motion_style_transfer_b200 runs on CUDA only (no CPU fallback); got {t.device}')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected dtype {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        t = t.contiguous()
    return t


_ws_cache = {}
_ws_retired = []      # outgrown buffers stay alive: a captured CUDA graph may have their addresses baked in


def _workspace(nbytes, device, tag):
    """Caller-owned scratch, grown on demand and reused per (device, tag).

    A buffer that a larger request outgrows is retired, never freed: ``GraphedForecaster`` captures these pointers into
    a CUDA graph, and a later eager call with a bigger batch must not hand the memory the graph still replays into back
    to the allocator.  (Scratch sizes are a few MB; the retired list grows only when a size record is broken.)"""
    key = (device, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ws_retired.append(buf)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


launch_count = 0      # kernels-launching C calls issued (bench.py reports it)
_profile = None       # list of per-launch records while profile_begin() is active


def profile_begin():
    global _profile
    _profile = []


def profile_end():
    """Per-launch device times (CUDA events on the launching stream) with algorithmic flops / bytes."""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    out = []
    for r in rec or []:
        out.append(dict(kernel=r['kernel'], tag=r['tag'], ms=r['e0'].elapsed_time(r['e1']), flops=r['flops'],
                        bytes=r['bytes']))
    return out


class _timed:
    def __init__(self, kernel, flops=0.0, nbytes=0.0, tag=''):
        self.kernel, self.flops, self.nbytes, self.tag = kernel, float(flops), float(nbytes), tag

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _profile is not None and exc[0] is None:
            self.e1.record()
            _profile.append(dict(kernel=self.kernel, tag=self.tag, e0=self.e0, e1=self.e1, flops=self.flops,
                                 bytes=self.nbytes))
        return False


def _count(n=1):
    global launch_count
    launch_count += n


# ------------------------------------------------------------------------------------------------ a1/a3/a9
def create_dist_template(size, device):
    out = torch.empty(size, size, dtype=torch.float32, device=device)
    check(_L().ynet_create_dist_template(size, _ptr(out), _stream()), 'create_dist_template')
    _count()
    return out


def rasterize_patches(template, coords, H, W, check_bounds=False):
    """get_patch + stack: template (th, tw) f32, coords (n, 2) f32 (x, y) -> (n, H, W)."""
    template = _req(template, name='template')
    coords = _req(coords.reshape(-1, 2), name='coords')
    n = coords.shape[0]
    out = torch.empty(n, H, W, dtype=torch.float32, device=template.device)
    flag = torch.zeros(1, dtype=torch.int32, device=template.device) if check_bounds else None
    with _timed('rasterize_gather_kernel', 0, 4.0 * n * H * W):
        check(_L().ynet_rasterize_patches(_ptr(template), template.shape[0], template.shape[1], _ptr(coords), n,
                                          _ptr(out), H, W, _ptr(flag), _stream()), 'rasterize_patches')
    _count()
    if check_bounds and int(flag.item()):
        raise ValueError('get_patch: window leaves the template (coordinate outside the image)')
    return out


def rasterize_dist_analytic(template_size, coords, H, W):
    coords = _req(coords.reshape(-1, 2), name='coords')
    n = coords.shape[0]
    out = torch.empty(n, H, W, dtype=torch.float32, device=coords.device)
    check(_L().ynet_rasterize_dist_analytic(template_size, _ptr(coords), n, _ptr(out), H, W, _stream()),
          'rasterize_dist_analytic')
    _count()
    return out


def avgpool_pyramid(maps, n_levels):
    """maps (B, C, H, W) -> [maps, AvgPool2d(2)(maps), ..., AvgPool2d(2^(n_levels-1))(maps)]."""
    maps = _req(maps, name='maps')
    B, C, H, W = maps.shape
    outs = [torch.empty(B, C, H >> i, W >> i, dtype=torch.float32, device=maps.device) for i in range(1, n_levels)]
    arr = (ctypes.c_void_p * max(1, len(outs)))(*[o.data_ptr() for o in outs])
    if n_levels >= 2:
        with _timed('avgpool_pyramid_kernel', 0, 4.0 * B * C * H * W * 4 / 3):
            check(_L().ynet_avgpool_pyramid(_ptr(maps), B * C, H, W, n_levels, arr, _stream()), 'avgpool_pyramid')
        _count()
    return [maps] + outs


# ------------------------------------------------------------------------------------------------ 8f rank 2: scene images
def scene_preprocess_u8(img_u8, dh, dw, Hp, Wp, xt, yt, int_scale, mean3=None, std3=None, want_chw=True, want_u8=False,
                        orient=0):
    """uint8 HWC CUDA image -> (float32 (3, Hp, Wp) normalised | None, uint8 (dh, dw, 3) resized | None).
    xt / yt: (start, src, weight) CUDA tensors of cv::computeResizeAreaTab (None with int_scale > 0).
    orient = k + 4 * flip: the chain is applied to the augmented view cv2.flip(rot90 x k (img), 1) without materialising
    it (data_utils.py:115-233); dh, dw, Hp, Wp and the tables then refer to the view."""
    img_u8 = _req(img_u8, torch.uint8, 'image')
    H, W, C = img_u8.shape
    if C != 3:
        raise ValueError(f'scene_preprocess_u8: expected an (H, W, 3) image, got {tuple(img_u8.shape)}')
    out = torch.empty(3, Hp, Wp, dtype=torch.float32, device=img_u8.device) if want_chw else None
    out8 = torch.empty(dh, dw, 3, dtype=torch.uint8, device=img_u8.device) if want_u8 else None
    m = (ctypes.c_double * 3)(*mean3) if mean3 is not None else None
    sd = (ctypes.c_double * 3)(*std3) if std3 is not None else None
    tabs = [None] * 6 if int_scale else [_ptr(t) for t in (xt + yt)]
    with _timed('scene_preprocess_kernel', 0, 3.0 * H * W + 12.0 * Hp * Wp):
        check(_L().ynet_scene_preprocess_oriented_u8(_ptr(img_u8), H, W, int(orient), dh, dw, Hp, Wp, *tabs, int(int_scale), m,
                                                     sd, _ptr(out), _ptr(out8), _stream()), 'scene_preprocess_oriented_u8')
    _count()
    return out, out8


def scene_onehot_u8(mask_u8, dh, dw, Hp, Wp, inv_factor, classes):
    """uint8 (H, W) CUDA segmentation mask -> float32 (classes, Hp, Wp): INTER_NEAREST resize, zero pad, one-hot."""
    mask_u8 = _req(mask_u8, torch.uint8, 'mask')
    H, W = mask_u8.shape
    out = torch.empty(classes, Hp, Wp, dtype=torch.float32, device=mask_u8.device)
    check(_L().ynet_scene_onehot_u8(_ptr(mask_u8), H, W, dh, dw, Hp, Wp, float(inv_factor), classes, _ptr(out), _stream()),
          'scene_onehot_u8')
    _count()
    return out


# ------------------------------------------------------------------------------------------------ a10/a12/a13
def softargmax2d(x, channel=None):
    """(B, C, H, W) -> (B, C, 2); with `channel` only that channel of every image -> (B, 1, 2)."""
    x = _req(x, name='input')
    B, C, H, W = x.shape
    if channel is None:
        rows, stride, base, oc = B * C, H * W, x.data_ptr(), C
    else:
        rows, stride, base, oc = B, C * H * W, x.data_ptr() + (int(channel) % C) * H * W * 4, 1
    out = torch.empty(B, oc, 2, dtype=torch.float32, device=x.device)
    nb = _L().ynet_softargmax2d_workspace_bytes(rows, H, W)
    ws = _workspace(nb, x.device, 'softargmax')
    with _timed('softargmax_partial_kernel', 0, 4.0 * rows * H * W):
        check(_L().ynet_softargmax2d(ctypes.c_void_p(base), rows, stride, H, W, _ptr(out), _ptr(ws), ws.numel(),
                                     _stream()), 'softargmax2d')
    _count(2)
    return out


def spatial_softmax(x):
    x = _req(x, name='input')
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    check(_L().ynet_spatial_softmax(_ptr(x), B * C, H * W, _ptr(out), _stream()), 'spatial_softmax')
    _count()
    return out


def expectation2d(p):
    p = _req(p, name='input')
    B, C, H, W = p.shape
    out = torch.empty(B, C, 2, dtype=torch.float32, device=p.device)
    check(_L().ynet_expectation2d(_ptr(p), B * C, H, W, _ptr(out), _stream()), 'expectation2d')
    _count()
    return out


def sigmoid_select(logits, channels, temperature):
    logits = _req(logits, name='logits')
    B, C, H, W = logits.shape
    ch = (ctypes.c_int32 * len(channels))(*[int(c) % C for c in channels])
    out = torch.empty(B, len(channels), H, W, dtype=torch.float32, device=logits.device)
    with _timed('sigmoid_select_kernel', 0, 8.0 * B * len(channels) * H * W):
        check(_L().ynet_sigmoid_select(_ptr(logits), B, C, H * W, ch, len(channels), float(temperature), _ptr(out),
                                       _stream()), 'sigmoid_select')
    _count()
    return out


# ------------------------------------------------------------------------------------------------ a11
def _prepare(p2d, rel_threshold):
    rows, S = p2d.shape
    rowmax = torch.empty(rows, dtype=torch.float32, device=p2d.device)
    gsum = torch.empty(1, dtype=torch.float32, device=p2d.device)
    nb = _L().ynet_sampling_prepare_workspace_bytes(rows, S)
    ws = _workspace(nb, p2d.device, 'sampling_prepare')
    check(_L().ynet_sampling_prepare(_ptr(p2d), rows, S, float(rel_threshold), _ptr(rowmax), _ptr(gsum), _ptr(ws),
                                     ws.numel(), _stream()), 'sampling_prepare')
    _count(3)
    return rowmax, gsum


def multinomial_replacement(prob_map, uniforms, rel_threshold=None):
    """prob_map (B, C, H, W) f32, uniforms (B*C, n) f64 -> idx (B*C, n) int64, xy (B, C, n, 2) f32."""
    prob_map = _req(prob_map, name='probability_map')
    B, C, H, W = prob_map.shape
    rows, S = B * C, H * W
    p2d = prob_map.view(rows, S)
    uniforms = _req(uniforms, torch.float64, 'uniforms').view(rows, -1)
    n = uniforms.shape[1]
    rowmax = gsum = None
    if rel_threshold is not None:
        rowmax, gsum = _prepare(p2d, rel_threshold)
    cdf = _workspace(rows * S * 4, prob_map.device, 'cdf')
    idx = torch.empty(rows, n, dtype=torch.int64, device=prob_map.device)
    xy = torch.empty(B, C, n, 2, dtype=torch.float32, device=prob_map.device)
    with _timed('cdf_sequential_kernel+cdf_search_kernel', 0, 4.0 * rows * S + 8.0 * rows * n):
        check(_L().ynet_multinomial_replacement(_ptr(p2d), rows, S,
                                                -1.0 if rel_threshold is None else float(rel_threshold),
                                                _ptr(rowmax), _ptr(gsum), _ptr(uniforms), n, _ptr(cdf), _ptr(idx),
                                                _ptr(xy), W, _stream()), 'multinomial_replacement')
    _count(2)
    return idx, xy


def multinomial_topk(prob_map, expo, n, rel_threshold=None):
    """replacement=False (or n == 1): top-n of p / q.  expo (B*C, H*W) f32 ~ Exp(1)."""
    prob_map = _req(prob_map, name='probability_map')
    B, C, H, W = prob_map.shape
    rows, S = B * C, H * W
    p2d = prob_map.view(rows, S)
    expo = _req(expo, name='exponentials').view(rows, S)
    rowmax = gsum = None
    if rel_threshold is not None:
        rowmax, gsum = _prepare(p2d, rel_threshold)
    idx = torch.empty(rows, n, dtype=torch.int64, device=prob_map.device)
    xy = torch.empty(B, C, n, 2, dtype=torch.float32, device=prob_map.device)
    check(_L().ynet_multinomial_topk(_ptr(p2d), _ptr(expo), rows, S, -1.0 if rel_threshold is None else float(rel_threshold),
                                     _ptr(rowmax), _ptr(gsum), n, _ptr(idx), _ptr(xy), W, _stream()),
          'multinomial_topk')
    _count()
    return idx, xy


def rng_uniform_f64(seed, offset, n, device, epoch=None):
    out = torch.empty(n, dtype=torch.float64, device=device)
    check(_L().ynet_rng_uniform_f64(seed, _ptr(epoch), offset, n, _ptr(out), _stream()), 'rng_uniform_f64')
    _count()
    return out


def rng_exponential_f32(seed, offset, n, device, epoch=None):
    out = torch.empty(n, dtype=torch.float32, device=device)
    check(_L().ynet_rng_exponential_f32(seed, _ptr(epoch), offset, n, _ptr(out), _stream()), 'rng_exponential_f32')
    _count()
    return out


def counter_add(counter, inc=1):
    """counter: 1-element int64 CUDA tensor (device-resident epoch of a graph-safe DeviceRng)."""
    check(_L().ynet_counter_add(_ptr(counter), int(inc), _stream()), 'counter_add')
    _count()


def rng_choice(seed, offset, rows, N, K, device, epoch=None):
    out = torch.empty(rows, K, dtype=torch.int32, device=device)
    check(_L().ynet_rng_choice(seed, _ptr(epoch), offset, rows, N, K, _ptr(out), _stream()), 'rng_choice')
    _count()
    return out


# ------------------------------------------------------------------------------------------------ a14
def kmeans_batched(X, init_idx, reseed_idx=None, tol=1e-4, iter_limit=0, want_assign=False):
    """X (B, N, 2) f32, init_idx (B, K) int32 -> centres (B, K, 2), assign (B, N)|None, iters (B), status (B)."""
    X = _req(X, name='X')
    B, N, D = X.shape
    if D != 2:
        raise NotImplementedError('kmeans_batched: only 2-D points (pixel coordinates) are supported')
    init_idx = _req(init_idx, torch.int32, 'init_idx')
    K = init_idx.shape[1]
    R = 0
    if reseed_idx is not None:
        reseed_idx = _req(reseed_idx, torch.int32, 'reseed_idx')
        R = reseed_idx.shape[1]
    centres = torch.empty(B, K, 2, dtype=torch.float32, device=X.device)
    assign = torch.empty(B, N, dtype=torch.int32, device=X.device) if want_assign else None
    iters = torch.empty(B, dtype=torch.int32, device=X.device)
    status = torch.empty(B, dtype=torch.int32, device=X.device)
    with _timed('kmeans_kernel', 0, 8.0 * B * N):
        check(_L().ynet_kmeans_batched(_ptr(X), B, N, K, _ptr(init_idx), _ptr(reseed_idx), R, float(tol),
                                       int(iter_limit), _ptr(centres), _ptr(assign), _ptr(iters), _ptr(status),
                                       _stream()), 'kmeans_batched')
    _count()
    return centres, assign, iters, status


# ------------------------------------------------------------------------------------------------ a16/a17
def cws_waypoint(sig, wp_in, last_obs, length_ratio, sigma_factor, ratio, rot):
    """sig (B, H, W); wp_in (G, B, 2); last_obs (B, 2); sigma_factor (G,) -> (G, B, 2)."""
    sig = _req(sig, name='sig')
    B, H, W = sig.shape
    wp_in = _req(wp_in, name='wp_in')
    G = wp_in.shape[0]
    last_obs = _req(last_obs, name='last_obs')
    sigma_factor = _req(sigma_factor, name='sigma_factor')
    out = torch.empty(G, B, 2, dtype=torch.float32, device=sig.device)
    ws = _workspace(_L().ynet_cws_waypoint_workspace_bytes(B, min(G, 32)), sig.device, 'cws')
    for g0 in range(0, G, 32):
        g1 = min(G, g0 + 32)
        check(_L().ynet_cws_waypoint(_ptr(sig), B, H, W, _ptr(wp_in[g0:g1]), g1 - g0, _ptr(last_obs),
                                     float(length_ratio), _ptr(sigma_factor[g0:g1]), float(ratio), int(bool(rot)),
                                     _ptr(out[g0:g1]), _ptr(ws), ws.numel(), _stream()), 'cws_waypoint')
        _count(2)
    return out


def cws_waypoint_map(sig, wp_in_g, last_obs, length_ratio, sigma_factor, ratio, rot):
    sig = _req(sig, name='sig')
    B, H, W = sig.shape
    out = torch.empty(B, H, W, dtype=torch.float32, device=sig.device)
    check(_L().ynet_cws_waypoint_map(_ptr(sig), B, H, W, _ptr(_req(wp_in_g, name='wp_in_g')),
                                     _ptr(_req(last_obs, name='last_obs')), float(length_ratio), float(sigma_factor),
                                     float(ratio), int(bool(rot)), _ptr(out), _stream()), 'cws_waypoint_map')
    _count()
    return out


def ade_fde(gt_future, trajs, wps, resize_factor):
    gt_future = _req(gt_future, name='gt_future')
    trajs = _req(trajs, name='trajs')
    wps = _req(wps, name='waypoints')
    K, B, T, _ = trajs.shape
    n_wp = wps.shape[2]
    ade = torch.empty(B, dtype=torch.float32, device=trajs.device)
    fde = torch.empty(B, dtype=torch.float32, device=trajs.device)
    check(_L().ynet_ade_fde(_ptr(gt_future), _ptr(trajs), _ptr(wps), K, B, T, n_wp, float(resize_factor), _ptr(ade),
                            _ptr(fde), _stream()), 'ade_fde')
    _count()
    return ade, fde


# ------------------------------------------------------------------------------------------------ a4-a8 fp32 engine
def lora_fold(weight, lora_A=None, lora_B=None, packed=True):
    """W + (B @ A).view(W.shape) / r  ->  OIHW (packed=False) or [C_in][k*k][C_out] (packed=True)."""
    weight = _req(weight, name='weight')
    C_out, C_in, k, _ = weight.shape
    rank = 0
    if lora_A is not None:
        lora_A = _req(lora_A, name='lora_A')
        lora_B = _req(lora_B, name='lora_B')
        rank = lora_A.shape[0] // k
    out = torch.empty(C_in * k * k * C_out, dtype=torch.float32, device=weight.device)
    check(_L().ynet_lora_fold(_ptr(weight), _ptr(lora_A), _ptr(lora_B), C_out, C_in, k, rank, 1 if packed else 0,
                              _ptr(out), _stream()), 'lora_fold')
    _count()
    return out.view(C_in, k * k, C_out) if packed else out.view(C_out, C_in, k, k)


def conv3x3_f32(sources, weight_packed, bias, relu, N, H, W):
    """sources: list of (tensor NCHW f32, mode); weight_packed [C_in][9][C_out]; -> (N, C_out, H, W).

    A source whose batch is 1 is broadcast; a source whose batch B divides N is read as n % B
    (features shared by stacked goal passes)."""
    C_out = weight_packed.shape[2]
    arr = (ConvSrc * len(sources))()
    keep = []
    cin = 0
    for i, (t, mode) in enumerate(sources):
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError('conv3x3_f32: sources must be float32 CUDA tensors')
        if t.dim() != 4:
            raise ValueError('conv3x3_f32: sources must be 4-D')
        bs = t.stride(0)
        if t.shape[0] == 1 and N > 1:
            bs = 0                                   # broadcast over the batch (Tensor.expand)
        if not t[0].is_contiguous():
            t = t.contiguous()
            bs = 0 if (t.shape[0] == 1 and N > 1) else t.stride(0)
        keep.append(t)
        arr[i].ptr = t.data_ptr()
        arr[i].channels = t.shape[1]
        arr[i].mode = mode
        arr[i].batch_stride = bs
        arr[i].batch_mod = t.shape[0] if (1 < t.shape[0] < N and N % t.shape[0] == 0 and bs != 0) else 0
        if 1 < t.shape[0] < N and arr[i].batch_mod == 0:
            raise ValueError(f'conv3x3_f32: source batch {t.shape[0]} does not divide N={N}')
        cin += t.shape[1]
    if cin != weight_packed.shape[0]:
        raise ValueError(f'conv3x3_f32: sources give {cin} channels, weight expects {weight_packed.shape[0]}')
    out = torch.empty(N, C_out, H, W, dtype=torch.float32, device=weight_packed.device)
    co_blocks = (C_out + 31) // 32
    per = max(1, 65535 // co_blocks)
    for n0 in range(0, N, per):
        nn = min(per, N - n0)
        if n0:
            for i, t in enumerate(keep):
                if arr[i].batch_mod:
                    raise ValueError('conv3x3_f32: modulo-batched sources need N * ceil(C_out/32) <= 65535')
                arr[i].ptr = t.data_ptr() + n0 * arr[i].batch_stride * 4
        with _timed('conv3x3_f32_kernel', 2.0 * 9 * cin * C_out * H * W * nn, 4.0 * (cin + C_out) * H * W * nn,
                    tag=f'{cin}->{C_out}@{H}x{W} N={nn}'):
            check(_L().ynet_conv3x3_f32(arr, len(sources), nn, H, W, _ptr(weight_packed), _ptr(bias), C_out,
                                        1 if relu else 0, _ptr(out[n0:]), _stream()), 'conv3x3_f32')
        _count()
    return out


def conv1x1_f32(x, weight, bias):
    x = _req(x, name='x')
    N, C_in, H, W = x.shape
    weight = _req(weight, name='weight')
    C_out = weight.shape[0]
    out = torch.empty(N, C_out, H, W, dtype=torch.float32, device=x.device)
    with _timed('conv1x1_f32_kernel', 2.0 * C_in * C_out * H * W * N, 4.0 * (C_in + C_out) * H * W * N):
        check(_L().ynet_conv1x1_f32(_ptr(x), N, C_in, H * W, _ptr(weight), _ptr(bias), C_out, _ptr(out), _stream()),
              'conv1x1_f32')
    _count()
    return out


def predictor_softargmax_f32(x, weight, bias):
    """1x1 predictor + SoftArgmax2D fused: x (N, C_in, H, W) -> (N, C_out, 2)."""
    x = _req(x, name='x')
    N, C_in, H, W = x.shape
    weight = _req(weight, name='weight')
    C_out = weight.shape[0]
    out = torch.empty(N, C_out, 2, dtype=torch.float32, device=x.device)
    nb = _L().ynet_predictor_softargmax_workspace_bytes(N, C_out, H, W)
    ws = _workspace(nb, x.device, 'pred_softargmax')
    with _timed('predictor_softargmax_kernel', 0, 4.0 * C_in * H * W * N):
        check(_L().ynet_predictor_softargmax_f32(_ptr(x), N, C_in, H, W, _ptr(weight), _ptr(bias), C_out, _ptr(out),
                                                 _ptr(ws), ws.numel(), _stream()), 'predictor_softargmax_f32')
    _count(2)
    return out


def maxpool2x2(x):
    x = _req(x, name='x')
    N, C, H, W = x.shape
    out = torch.empty(N, C, H // 2, W // 2, dtype=torch.float32, device=x.device)
    check(_L().ynet_maxpool2x2_f32(_ptr(x), N * C, H, W, _ptr(out), _stream()), 'maxpool2x2')
    _count()
    return out


def upsample_bilinear2x(x):
    x = _req(x, name='x')
    N, C, H, W = x.shape
    out = torch.empty(N, C, 2 * H, 2 * W, dtype=torch.float32, device=x.device)
    check(_L().ynet_upsample_bilinear2x_f32(_ptr(x), N * C, H, W, _ptr(out), _stream()), 'upsample_bilinear2x')
    _count()
    return out


# ------------------------------------------------------------------------------------------------ a18 training pieces
def bce_logits_fwd_bwd(logits, target, grad_scale, want_grad=True):
    logits = _req(logits, name='logits')
    target = _req(target, name='target')
    n = logits.numel()
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits) if want_grad else None
    ws = _workspace(_L().ynet_bce_workspace_bytes(n), logits.device, 'bce')
    check(_L().ynet_bce_logits_fwd_bwd(_ptr(logits), _ptr(target), n, float(grad_scale), _ptr(loss), _ptr(grad),
                                       _ptr(ws), ws.numel(), _stream()), 'bce_logits_fwd_bwd')
    _count(2)
    return loss, grad


def conv3x3_dgrad_f32(dy, relu_out, weight_oihw):
    dy = _req(dy, name='dy')
    N, C_out, H, W = dy.shape
    weight_oihw = _req(weight_oihw, name='weight')
    C_in = weight_oihw.shape[1]
    if relu_out is not None:
        relu_out = _req(relu_out, name='relu_out')
    dx = torch.empty(N, C_in, H, W, dtype=torch.float32, device=dy.device)
    check(_L().ynet_conv3x3_dgrad_f32(_ptr(dy), _ptr(relu_out), N, H, W, _ptr(weight_oihw), C_out, C_in, _ptr(dx),
                                      _stream()), 'conv3x3_dgrad_f32')
    _count(3)
    return dx


def conv3x3_wgrad_f32(x, dy, relu_out, want_bias=True):
    x = _req(x, name='x')
    dy = _req(dy, name='dy')
    N, C_in, H, W = x.shape
    C_out = dy.shape[1]
    if relu_out is not None:
        relu_out = _req(relu_out, name='relu_out')
    dW = torch.empty(C_out, C_in, 3, 3, dtype=torch.float32, device=x.device)
    db = torch.empty(C_out, dtype=torch.float32, device=x.device) if want_bias else None
    nb = _L().ynet_conv3x3_wgrad_workspace_bytes(N, H, W, C_out, C_in)
    ws = _workspace(nb, x.device, 'wgrad')
    check(_L().ynet_conv3x3_wgrad_f32(_ptr(x), _ptr(dy), _ptr(relu_out), N, H, W, C_in, C_out, _ptr(dW), _ptr(db),
                                      _ptr(ws), ws.numel(), _stream()), 'conv3x3_wgrad_f32')
    _count(3)
    return dW, db


def maxpool2x2_bwd(x, dy):
    x = _req(x, name='x')
    dy = _req(dy, name='dy')
    N, C, H, W = x.shape
    dx = torch.empty_like(x)
    check(_L().ynet_maxpool2x2_bwd_f32(_ptr(x), _ptr(dy), N * C, H, W, _ptr(dx), _stream()), 'maxpool2x2_bwd')
    _count()
    return dx


def upsample_bilinear2x_bwd(dy):
    dy = _req(dy, name='dy')
    N, C, OH, OW = dy.shape
    dx = torch.empty(N, C, OH // 2, OW // 2, dtype=torch.float32, device=dy.device)
    check(_L().ynet_upsample_bilinear2x_bwd_f32(_ptr(dy), N * C, OH // 2, OW // 2, _ptr(dx), _stream()),
          'upsample_bilinear2x_bwd')
    _count()
    return dx


def lora_grad(dW, lora_A, lora_B):
    dW = _req(dW, name='dW')
    C_out, C_in, k, _ = dW.shape
    lora_A = _req(lora_A, name='lora_A')
    lora_B = _req(lora_B, name='lora_B')
    rank = lora_A.shape[0] // k
    dA = torch.empty_like(lora_A)
    dB = torch.empty_like(lora_B)
    check(_L().ynet_lora_grad(_ptr(dW), _ptr(lora_A), _ptr(lora_B), C_out, C_in, k, rank, _ptr(dA), _ptr(dB),
                              _stream()), 'lora_grad')
    _count()
    return dA, dB


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    for t, nme in ((param, 'param'), (grad, 'grad'), (exp_avg, 'exp_avg'), (exp_avg_sq, 'exp_avg_sq')):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError(f'adam_step: {nme} must be a contiguous float32 CUDA tensor')
    check(_L().ynet_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), int(step),
                              float(lr), float(beta1), float(beta2), float(eps), float(grad_scale), _stream()),
          'adam_step')
    _count()


# ------------------------------------------------------------------------------------------------ a4-a8 tensor-core engine
class C8:
    """bf16 activation in C8 planes: data (N, C_pad/8, H, W, 8) bfloat16, logical channel count C.

    ``rep`` > 1 makes a lazy ``repeat_interleave(rep, dim=0)``: as a conv source, image n reads data[n // rep]
    (agent-major stacking of the n_goal decoder passes; nothing is copied).
    """

    __slots__ = ('data', 'C', 'rep', 'center', 'pad', 'taps')

    def __init__(self, data, C, rep=1, center=False, pad=0, taps=0):
        # center: hoisted partial sums (hi | lo); a conv applies identity weights on its centre tap only
        # pad = 1: planes are (H + 2, W + 2) with a one-pixel replicated ring (input of tc_upconv3x3 only)
        # taps = TAPS_QUAD: 2x2-neighbourhood planes (tc_rasterize_pyramid quad levels); C counts the stored channels
        self.data, self.C, self.rep, self.center, self.pad, self.taps = data, C, rep, center, pad, taps

    @property
    def N(self):
        return self.data.shape[0] * self.rep

    def repeat_interleave(self, rep):
        return C8(self.data, self.C, self.rep * rep, self.center, self.pad, self.taps)

    def batch_slice(self, b0, b1):
        if self.rep != 1:
            raise ValueError('batch_slice of a repeated C8')
        return C8(self.data[b0:b1], self.C, 1, self.center, self.pad, self.taps)

    def unpadded(self):
        """The H x W interior of a replicate-padded C8 as a plain (contiguous) C8."""
        if not self.pad:
            return self
        return C8(self.data[:, :, 1:-1, 1:-1].contiguous(), self.C, self.rep, self.center, 0)

    @property
    def C_pad(self):
        return self.data.shape[1] * 8

    @property
    def K_pad(self):
        """Channels a conv sees (multiple of 16); planes beyond the stored ones read as zero (TMA fill)."""
        return (self.C_pad + 15) // 16 * 16

    @property
    def H(self):
        return self.data.shape[2] - 2 * self.pad

    @property
    def W(self):
        return self.data.shape[3] - 2 * self.pad


def _pad16(c):
    return (c + 15) // 16 * 16


def tc_rasterize_im2col(template, coords, n_img, n_ch, H, W, level):
    """Waypoint maps of pyramid level 0 or 1 in im2col form (see ynet_tc_rasterize_im2col_c8): a ``center`` C8 with
    9 * n_ch channels whose conv weights are W[:, wp channels].reshape(C_out, 9 * n_ch, 1, 1)."""
    template = _req(template, name='template')
    coords = _req(coords, name='coords').reshape(-1, 2)
    if coords.shape[0] != n_img * n_ch:
        raise ValueError(f'tc_rasterize_im2col: expected {n_img * n_ch} coordinates, got {coords.shape[0]}')
    cp = _pad16(9 * n_ch)
    out = torch.empty(n_img, cp // 8, H >> level, W >> level, 8, dtype=torch.bfloat16, device=template.device)
    with _timed('wp_im2col_c8_kernel', 0, (2.0 * cp * ((H >> level) * (W >> level)) + 4.0 * n_ch * H * W) * n_img):
        check(_L().ynet_tc_rasterize_im2col_c8(_ptr(template), template.shape[0], template.shape[1], _ptr(coords), n_img,
                                               n_ch, H, W, level, _ptr(out), cp, _stream()), 'tc_rasterize_im2col_c8')
    _count()
    return C8(out, 9 * n_ch, 1, True)


def tc_supported():
    return bool(_L().ynet_tc_supported())


def tc_pack(x):
    """NCHW float32 (batch may be broadcast: stride 0 or N == 1) -> C8 bf16."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
        raise RuntimeError('tc_pack: expected a 4-D float32 CUDA tensor')
    if x.shape[0] > 1 and x.stride(0) == 0:
        x = x[:1]
    if not x[0].is_contiguous():
        x = x.contiguous()
    N, C, H, W = x.shape
    cp = _pad16(C)
    out = torch.empty(N, cp // 8, H, W, 8, dtype=torch.bfloat16, device=x.device)
    with _timed('pack_c8_kernel', 0, (4.0 * C + 2.0 * cp) * N * H * W):
        check(_L().ynet_tc_pack_f32_to_c8(_ptr(x), N, C, H, W, x.stride(0) if N > 1 else C * H * W, _ptr(out), cp,
                                          _stream()), 'tc_pack_f32_to_c8')
    _count()
    return C8(out, C)


TAPS_QUAD = 0x1B      # ynet_b200.h YNET_TC_TAPS_QUAD: taps (0,0) (0,1) (1,0) (1,1) of the 3x3 window


def tc_quad_weights(weight_oihw, c0, n_ch):
    """(C_out, C_in, 3, 3) weight -> the (C_out, 4 * n_ch, 2, 2) weight of a 2x2-neighbourhood source holding input
    channels [c0, c0 + n_ch): stored channel (dy*2 + dx) * n_ch + c under the tap anchored at (ty - 1, tx - 1) carries
    w[:, c0 + c, ty + dy, tx + dx]; every 3x3 tap is assigned to exactly one (anchor, offset) pair -- the anchors at -1
    use only offset 0, so a zero-filled out-of-image anchor loses nothing that lies inside the image."""
    C_out = weight_oihw.shape[0]
    q = torch.zeros(C_out, 4 * n_ch, 2, 2, dtype=torch.float32, device=weight_oihw.device)
    for ty in range(2):
        for tx in range(2):
            for dy in range(2):
                for dx in range(2):
                    if (ty == 0 and dy == 1) or (tx == 0 and dx == 1):
                        continue
                    q[:, (dy * 2 + dx) * n_ch:(dy * 2 + dx + 1) * n_ch, ty, tx] = weight_oihw[:, c0:c0 + n_ch, ty + dy, tx + dx]
    return q


_wp_template_cache = {}


def tc_wp_template(template):
    """bf16 C8 planes of the distance template for the row-marching conv's waypoint source (ynet_tc_wp_template_c8):
    (level-0 planes (th, tw, 8), level-1 parity planes (4, th/2, tw/2, 8)); built once per template tensor."""
    key = (template.data_ptr(), template._version, tuple(template.shape), template.device)
    hit = _wp_template_cache.get(key)
    if hit is None:
        th, tw = template.shape
        l0 = torch.empty(th, tw, 8, dtype=torch.bfloat16, device=template.device)
        l1 = torch.empty(4, th // 2, tw // 2, 8, dtype=torch.bfloat16, device=template.device)
        check(_L().ynet_tc_wp_template_c8(_ptr(template), th, tw, _ptr(l0), _ptr(l1), _stream()), 'tc_wp_template_c8')
        _count()
        if len(_wp_template_cache) > 8:
            _wp_template_cache.clear()
        hit = _wp_template_cache[key] = (l0, l1, template)       # (keeps the keyed tensor alive: no pointer reuse)
    return hit[0], hit[1]


class WpPlanes:
    """Level ``level`` (0 or 1) of a waypoint pyramid that is NOT materialised: the row-marching conv loads it straight
    from the distance template's bf16 planes (ynet_tc_rowconv3x3_wp).  Stands where the C8 planes of that level would:
    (N, n_ch <= 2 channels, H, W) at the level's resolution, one 16-channel K block with channel c at K index 8 c."""

    __slots__ = ('template', 'coords', 'N', 'C', 'level', 'H', 'W', 'planes')
    rep, pad, center, taps, K_pad, C_pad = 1, 0, False, 0, 16, 8

    def __init__(self, template, coords, N, C, level, H, W):
        self.template, self.coords, self.N, self.C, self.level, self.H, self.W = template, coords, N, C, level, H, W
        self.planes = tc_wp_template(template)[level]

    def weight_parts(self, c0):
        """tc_rowconv_pack_weights_cat parts of this source: its channels are input channels [c0, c0 + C) of the conv."""
        return [(c0, c0 + 1, 16)] if self.C == 1 else [(c0, c0 + 1, 8), (c0 + 1, c0 + 2, 8)]

    def materialize(self):
        """The C8 planes of this level (what tc_rasterize_pyramid would have written)."""
        return tc_rasterize_pyramid(self.template, self.coords, self.N, self.C, self.H << self.level, self.W << self.level,
                                    self.level + 1, only_level=self.level)[self.level]


def tc_rasterize_pyramid(template, coords, n_img, n_ch, H, W, n_levels, slot=0, quad_levels=0, lazy_levels=0,
                         only_level=None):
    """get_patch + AvgPool pyramid of ``n_img x n_ch`` waypoint coordinates written straight as bf16 C8 planes
    (image_utils.py:40-63 + evaluate.py:255-257).  Returns n_levels C8 (n_img, ONE 8-channel plane, H>>l, W>>l):
    16 B per pixel; a conv pads the K block to 16 channels through the TMA zero fill (``C8.K_pad``).
    ``slot`` is accepted for compatibility and ignored (every call returns fresh tensors).
    ``quad_levels`` (n_ch <= 2): the finest levels are written as 2x2-neighbourhood planes (C8.taps = TAPS_QUAD,
    4 * n_ch stored channels): a conv spends four MMAs per K block on them instead of nine.
    """
    template = _req(template, name='template')
    coords = _req(coords, name='coords').reshape(-1, 2)
    if coords.shape[0] != n_img * n_ch:
        raise ValueError(f'tc_rasterize_pyramid: expected {n_img * n_ch} coordinates, got {coords.shape[0]}')
    # n_ch <= 8 channels fit ONE 8-channel plane; the conv's TMA zero-fills the other plane of the 16-channel K block
    # lazy_levels (<= 2, n_ch <= 2): the finest levels come back as WpPlanes (gathered inside the row-marching conv)
    # and are not written; only_level: write that level alone
    lazy_levels = lazy_levels if (n_ch <= 2 and template.shape[0] % 2 == 0 and template.shape[1] % 2 == 0) else 0
    skip = [(l < lazy_levels) or (only_level is not None and l != only_level) for l in range(n_levels)]
    bufs = [None if skip[l] else torch.empty(n_img, 1, H >> l, W >> l, 8, dtype=torch.bfloat16, device=template.device)
            for l in range(n_levels)]
    write_pad = 0
    outs = (ctypes.c_void_p * n_levels)(*[None if b is None else b.data_ptr() for b in bufs])
    S = sum((H >> l) * (W >> l) for l in range(n_levels) if not skip[l])
    with _timed('wp_pyramid_c8_kernel', 0, 16.0 * S * n_img):
        check(_L().ynet_tc_rasterize_pyramid_c8(_ptr(template), template.shape[0], template.shape[1], _ptr(coords), n_img,
                                                n_ch, H, W, n_levels, outs, 8, write_pad, quad_levels, None, _stream()),
              'tc_rasterize_pyramid_c8')
    _count()
    return [(WpPlanes(template, coords, n_img, n_ch, l, H >> l, W >> l) if l < lazy_levels else None) if b is None
            else (C8(b, 4 * n_ch, taps=TAPS_QUAD) if l < quad_levels else C8(b, n_ch)) for l, b in enumerate(bufs)]


def tc_pad_replicate(a):
    """One-pixel replicate padding of a C8 (the bilinear index clamping of tc_upconv3x3 made explicit)."""
    if a.pad:
        return a
    if a.rep != 1:
        raise ValueError('tc_pad_replicate of a repeated C8')
    out = torch.empty(a.data.shape[0], a.C_pad // 8, a.H + 2, a.W + 2, 8, dtype=torch.bfloat16, device=a.data.device)
    with _timed('c8_pad_replicate_kernel', 0, 4.0 * a.C_pad * a.data.shape[0] * a.H * a.W):
        check(_L().ynet_tc_pad_replicate(_ptr(a.data), a.data.shape[0], a.C_pad, a.H, a.W, _ptr(out), _stream()),
              'tc_pad_replicate')
    _count()
    return C8(out, a.C, 1, a.center, 1)


def tc_unpack(a):
    a = a.unpadded()
    out = torch.empty(a.N, a.C, a.H, a.W, dtype=torch.float32, device=a.data.device)
    check(_L().ynet_tc_unpack_c8_to_f32(_ptr(a.data), a.N, a.C, a.C_pad, a.H, a.W, _ptr(out), _stream()),
          'tc_unpack_c8_to_f32')
    _count()
    return out


def tc_maxpool(a):
    a = a.unpadded()
    out = torch.empty(a.N, a.C_pad // 8, a.H // 2, a.W // 2, 8, dtype=torch.bfloat16, device=a.data.device)
    with _timed('c8_maxpool_kernel', 0, 2.5 * a.C_pad * a.N * a.H * a.W):
        check(_L().ynet_tc_maxpool2x2(_ptr(a.data), a.N, a.C_pad, a.H, a.W, _ptr(out), _stream()), 'tc_maxpool2x2')
    _count()
    return C8(out, a.C)


def tc_upsample(a):
    a = a.unpadded()
    out = torch.empty(a.N, a.C_pad // 8, a.H * 2, a.W * 2, 8, dtype=torch.bfloat16, device=a.data.device)
    with _timed('c8_upsample_kernel', 0, 10.0 * a.C_pad * a.N * a.H * a.W):
        check(_L().ynet_tc_upsample2x(_ptr(a.data), a.N, a.C_pad, a.H, a.W, _ptr(out), _stream()), 'tc_upsample2x')
    _count()
    return C8(out, a.C)


def tc_pack_weights(weight_oihw, src_channels):
    """Effective OIHW float32 weight (3x3 or 1x1) -> bf16 [kb][tap][2][C_out_pad][8] over per-source padded channels."""
    weight_oihw = _req(weight_oihw, name='weight')
    C_out, ksize = weight_oihw.shape[0], weight_oihw.shape[2]
    n = len(src_channels)
    real = (ctypes.c_int32 * n)(*src_channels)
    pad = (ctypes.c_int32 * n)(*[_pad16(c) for c in src_channels])
    nbytes = _L().ynet_tc_packed_weight_bytes(C_out, n, pad, ksize)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=weight_oihw.device)
    check(_L().ynet_tc_pack_weights(_ptr(weight_oihw), C_out, n, real, pad, ksize, _ptr(packed), _stream()),
          'tc_pack_weights')
    _count()
    return packed


def _tc_batch_mod(s, N):
    """ynet_tc_src.batch_mod of source ``s`` inside a conv over N images (see include/ynet_b200.h)."""
    if s.rep > 1:
        if s.N != N:
            raise ValueError(f'repeated source covers {s.N} images, the conv has N={N}')
        return -s.rep
    if 1 < s.N < N:
        if N % s.N != 0:
            raise ValueError(f'source batch {s.N} does not divide N={N}')
        return s.N
    return 0


def _tc_src_array(sources, N):
    arr = (_lib.TcSrc * len(sources))()
    for i, s in enumerate(sources):
        arr[i].ptr = s.data.data_ptr()
        arr[i].channels_pad = s.C_pad
        arr[i].batch_stride = 0 if (s.N == 1 and N > 1) else s.data.stride(0)
        arr[i].batch_mod = _tc_batch_mod(s, N)
    return arr


def tc_conv1x1_f32(a, packed_weight, bias_pad, C_out):
    """1x1 predictor on the tensor cores -> float32 NCHW logits."""
    out = torch.empty(a.N, C_out, a.H, a.W, dtype=torch.float32, device=a.data.device)
    arr = _tc_src_array([a], a.N)
    with _timed('tc_conv_kernel<1x1,f32>', 2.0 * a.C * C_out * a.H * a.W * a.N,
                (2.0 * a.C_pad + 4.0 * C_out) * a.H * a.W * a.N):
        check(_L().ynet_tc_conv1x1_f32(arr, 1, a.N, a.H, a.W, _ptr(packed_weight), _ptr(bias_pad), C_out, _ptr(out), 0,
                                       _stream()), 'tc_conv1x1_f32')
    _count()
    return out


def tc_conv1x1_softargmax(a, packed_weight, bias_pad, C_out):
    """1x1 predictor + SoftArgmax2D fused on the tensor-core kernel: C8 activation -> (N, C_out, 2)."""
    out = torch.empty(a.N, C_out, 2, dtype=torch.float32, device=a.data.device)
    nb = _L().ynet_tc_conv1x1_softargmax_workspace_bytes(a.N, C_out, a.H, a.W)
    ws = _workspace(nb, a.data.device, 'tc_softargmax')
    arr = _tc_src_array([a], a.N)
    with _timed('tc_pred_softargmax_kernel', 2.0 * a.C * C_out * a.H * a.W * a.N, 2.0 * a.C_pad * a.H * a.W * a.N):
        check(_L().ynet_tc_conv1x1_softargmax(arr, 1, a.N, a.H, a.W, _ptr(packed_weight), _ptr(bias_pad), C_out,
                                              _ptr(out), _ptr(ws), ws.numel(), 0, _stream()), 'tc_conv1x1_softargmax')
    _count(3)
    return out


def tc_conv3x3(sources, packed_weight, bias_pad, C_out, relu, pad_out=False):
    """sources: list of C8 (batch N, 1 = broadcast, or a divisor of N = modulo).  Returns C8 (N, C_out);
    pad_out: write the replicate-padded (H + 2, W + 2) layout that tc_upconv3x3 consumes."""
    if any(s.pad for s in sources):
        raise ValueError('tc_conv3x3: replicate-padded C8 tensors are inputs of tc_upconv3x3 only')
    N = max(s.N for s in sources)
    H, W = sources[0].H, sources[0].W
    arr = (_lib.TcSrc * len(sources))()
    for i, s in enumerate(sources):
        if s.H != H or s.W != W:
            raise ValueError('tc_conv3x3: sources must share the spatial size')
        arr[i].ptr = s.data.data_ptr()
        arr[i].channels_pad = s.K_pad
        arr[i].chunks_stored = s.C_pad // 8
        arr[i].batch_stride = 0 if (s.N == 1 and N > 1) else s.data.stride(0)
        arr[i].batch_mod = _tc_batch_mod(s, N)
        arr[i].center_only = 1 if s.center else 0
        arr[i].tap_mask = s.taps
    cp = _pad16(C_out)
    po = 1 if pad_out else 0
    out = torch.empty(N, cp // 8, H + 2 * po, W + 2 * po, 8, dtype=torch.bfloat16, device=sources[0].data.device)
    cin_pad = sum(s.C_pad for s in sources)
    args = (arr, len(sources), N, H, W, _ptr(packed_weight), _ptr(bias_pad), C_out, (1 if relu else 0) | (2 * po),
            _ptr(out), cp)
    key = (tuple(-s.K_pad if s.center else (s.K_pad, s.taps) if s.taps else s.K_pad for s in sources), cp, H, W,
           min(N, 64), po)
    tune = _tc_tune.get(key)
    if tune is None:
        tune = _tc_autotune(key, args) if (tc_autotune_enabled and not torch.cuda.is_current_stream_capturing()) else 0
    hoisted = ('+P' if any(s.center for s in sources) else '') + ('+Q' if any(s.taps for s in sources) else '')
    real_c = sum((s.C // 4 if s.taps else s.C) for s in sources if not s.center)     # quad planes store 4 copies
    # algorithmic bytes: every DISTINCT source image read once (a repeated / broadcast source is shared by its images)
    in_bytes = sum(2.0 * s.C_pad * min(s.data.shape[0], N) for s in sources) * H * W
    with _timed('tc_conv3x3_kernel', 2.0 * 9 * real_c * C_out * H * W * N,
                in_bytes + 2.0 * cp * H * W * N, tag=f'{cin_pad}{hoisted}->{cp}@{H}x{W} N={N}'):
        check(_L().ynet_tc_conv3x3(*args, tune, _stream()), 'tc_conv3x3')
    _count()
    return C8(out, C_out, 1, False, po)


def tc_conv3x3_pred_softargmax(sources, packed_weight, bias_pad, C_out, relu, packed_pred, pred_bias_pad, C_pred):
    """conv3x3(+bias, +ReLU) -> 1x1 predictor -> SoftArgmax2D in one kernel: C8 sources -> (N, C_pred, 2)."""
    N = max(s.N for s in sources)
    H, W = sources[0].H, sources[0].W
    arr = (_lib.TcSrc * len(sources))()
    for i, s in enumerate(sources):
        if s.H != H or s.W != W:
            raise ValueError('tc_conv3x3_pred_softargmax: sources must share the spatial size')
        arr[i].ptr = s.data.data_ptr()
        arr[i].channels_pad = s.K_pad
        arr[i].chunks_stored = s.C_pad // 8
        arr[i].batch_stride = 0 if (s.N == 1 and N > 1) else s.data.stride(0)
        arr[i].batch_mod = _tc_batch_mod(s, N)
        arr[i].center_only = 1 if s.center else 0
    out = torch.empty(N, C_pred, 2, dtype=torch.float32, device=sources[0].data.device)
    nb = _L().ynet_tc_conv1x1_softargmax_workspace_bytes(N, C_pred, H, W)
    ws = _workspace(nb, out.device, 'tc_softargmax')
    cin_pad = sum(s.K_pad for s in sources)
    flops = 2.0 * (9 * sum(s.C for s in sources if not s.center) * C_out + C_out * C_pred) * H * W * N
    with _timed('tc_conv_pred_kernel', flops, 2.0 * cin_pad * H * W * N, tag=f'{cin_pad}->{_pad16(C_out)}->{C_pred}@{H}x{W} N={N}'):
        check(_L().ynet_tc_conv3x3_pred_softargmax(arr, len(sources), N, H, W, _ptr(packed_weight), _ptr(bias_pad), C_out,
                                                   1 if relu else 0, _ptr(packed_pred), _ptr(pred_bias_pad), C_pred,
                                                   _ptr(out), _ptr(ws), ws.numel(), _stream()),
              'tc_conv3x3_pred_softargmax')
    _count(3)
    return out


def tc_conv3x3_hilo(sources, packed_weight, C_out, with_lo=True):
    """Raw 3x3 partial sums (no bias / ReLU) of C8 sources as a bf16 (hi | lo) pair: C8 with 2 * pad16(C_out)
    channels, flagged ``center`` so that a later tc_conv3x3 adds it through identity weights (goal-loop hoisting)."""
    N = max(s.N for s in sources)
    H, W = sources[0].H, sources[0].W
    arr = (_lib.TcSrc * len(sources))()
    for i, s in enumerate(sources):
        if s.H != H or s.W != W:
            raise ValueError('tc_conv3x3_hilo: sources must share the spatial size')
        arr[i].ptr = s.data.data_ptr()
        arr[i].channels_pad = s.K_pad
        arr[i].chunks_stored = s.C_pad // 8
        arr[i].batch_stride = 0 if (s.N == 1 and N > 1) else s.data.stride(0)
        arr[i].batch_mod = _tc_batch_mod(s, N)
    cp = _pad16(C_out)
    k = 2 if with_lo else 1
    out = torch.empty(N, k * cp // 8, H, W, 8, dtype=torch.bfloat16, device=sources[0].data.device)
    cin_pad = sum(s.C_pad for s in sources)
    with _timed('tc_conv3x3_hilo_kernel', 2.0 * 9 * sum(s.C for s in sources) * C_out * H * W * N,
                2.0 * (cin_pad + k * cp) * H * W * N, tag=f'{cin_pad}->{cp}x{k}@{H}x{W} N={N}'):
        check(_L().ynet_tc_conv3x3_hilo(arr, len(sources), N, H, W, _ptr(packed_weight), C_out, _ptr(out), cp, k - 1, 0,
                                        _stream()), 'tc_conv3x3_hilo')
    _count()
    return C8(out, k * cp, 1, True)


# ---- split-bf16 ("bf16x3") activations: the <= 1e-3 engine on the tensor cores (csrc/split_tc.cu) ---------------------
class Split:
    """float32 activation as hi = bf16(x), lo = bf16(x - hi): data (N, 2 * cp / 8, H, W, 8) bfloat16, hi planes then lo
    planes, cp = pad16(C).  As a conv source a batch of 1 is broadcast and a divisor of N is read modulo (goal-major
    stacking of the decoder passes, like the fp32 engine)."""

    __slots__ = ('data', 'C', 'layout')
    rep = 1

    def __init__(self, data, C, layout=None):
        # layout: [(C_i, cp_i)] of the activations concatenated inside (split_cat); their channels sit at sum(cp_<i)
        self.data, self.C = data, C
        self.layout = [(C, data.shape[1] * 4)] if layout is None else layout

    N = property(lambda self: self.data.shape[0])
    H = property(lambda self: self.data.shape[2])
    W = property(lambda self: self.data.shape[3])
    cp = property(lambda self: self.data.shape[1] * 4)


def split_empty(N, layout, H, W, device):
    """Uninitialised Split holding the channel concatenation ``layout`` = [(C_i, cp_i)]: the target of concat-on-write
    (``into=`` of split_pack / tc_conv3x3_split), every slice of which must then be written."""
    tot = sum(cp for _, cp in layout)
    return Split(torch.empty(N, tot // 4, H, W, 8, dtype=torch.bfloat16, device=device), sum(c for c, _ in layout),
                 list(layout))


def split_pack(x, into=None):
    """NCHW float32 (batch may be broadcast) -> Split; into = (Split buffer, channel offset): fill that slice instead."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
        raise RuntimeError('split_pack: expected a 4-D float32 CUDA tensor')
    if x.shape[0] > 1 and x.stride(0) == 0:
        x = x[:1]
    if not x[0].is_contiguous():
        x = x.contiguous()
    N, C, H, W = x.shape
    cp = _pad16(C)
    if into is None:
        out, tot, off = torch.empty(N, cp // 4, H, W, 8, dtype=torch.bfloat16, device=x.device), 0, 0
    else:
        out, tot, off = into[0].data, into[0].cp, into[1]
        if out.shape[0] != N or out.shape[2] != H or out.shape[3] != W:
            raise ValueError('split_pack: the target activation has another shape')
    with _timed('split_pack_kernel', 0, (4.0 * C + 4.0 * cp) * N * H * W):
        check(_L().ynet_split_pack_f32(_ptr(x), N, C, H, W, x.stride(0) if N > 1 else C * H * W, _ptr(out), cp, tot, off,
                                       _stream()), 'split_pack_f32')
    _count()
    return into[0] if into is not None else Split(out, C)


def split_pack_masked(x, relu_out):
    """dy * (relu_out > 0) -> Split (ReLU backward folded into the split of a gradient); contiguous NCHW float32."""
    x, relu_out = _req(x, name='x'), _req(relu_out, name='relu_out')
    if x.shape != relu_out.shape or x.dim() != 4:
        raise ValueError('split_pack_masked: gradient and activation must share one 4-D shape')
    N, C, H, W = x.shape
    cp = _pad16(C)
    out = torch.empty(N, cp // 4, H, W, 8, dtype=torch.bfloat16, device=x.device)
    with _timed('split_pack_kernel', 0, (8.0 * C + 4.0 * cp) * N * H * W):
        check(_L().ynet_split_pack_masked_f32(_ptr(x), _ptr(relu_out), N, C, H, W, _ptr(out), cp, _stream()),
              'split_pack_masked_f32')
    _count()
    return Split(out, C)


def split_unpack(a):
    out = torch.empty(a.N, a.C, a.H, a.W, dtype=torch.float32, device=a.data.device)
    with _timed('split_unpack_kernel', 0, 8.0 * a.C * a.N * a.H * a.W):
        check(_L().ynet_split_unpack_f32(_ptr(a.data), a.N, a.C, a.cp, a.H, a.W, _ptr(out), _stream()), 'split_unpack_f32')
    _count()
    return out


def split_maxpool(a):
    out = torch.empty(a.N, a.cp // 4, a.H // 2, a.W // 2, 8, dtype=torch.bfloat16, device=a.data.device)
    with _timed('split_maxpool_kernel', 0, 5.0 * a.cp * a.N * a.H * a.W):
        check(_L().ynet_split_maxpool2x2(_ptr(a.data), a.N, a.cp, a.H, a.W, _ptr(out), _stream()), 'split_maxpool2x2')
    _count()
    return Split(out, a.C)


def split_upsample(a):
    out = torch.empty(a.N, a.cp // 4, a.H * 2, a.W * 2, 8, dtype=torch.bfloat16, device=a.data.device)
    with _timed('split_upsample_kernel', 0, 20.0 * a.cp * a.N * a.H * a.W):
        check(_L().ynet_split_upsample2x(_ptr(a.data), a.N, a.cp, a.H, a.W, _ptr(out), _stream()), 'split_upsample2x')
    _count()
    return Split(out, a.C)


def split_cat(parts, N=None):
    """torch.cat along channels of Split activations (hi planes of all parts, then their lo planes).  Parts with a
    smaller batch are tiled (``n % batch``: goal-major stacking)."""
    N = max(p.N for p in parts) if N is None else N
    his, los = [], []
    for p in parts:
        d = p.data if p.N == N else p.data.repeat(N // p.N, 1, 1, 1, 1)
        k = d.shape[1] // 2
        his.append(d[:, :k])
        los.append(d[:, k:])
    return Split(torch.cat(his + los, dim=1), sum(p.C for p in parts), [l for p in parts for l in p.layout])


_split_sel_cache = {}


def _split_weight_select(src_layouts, c_in, device):
    """For the channel-concatenation [W_hi (c_in) | W_lo (c_in) | one zero channel]: the gather index that lays the input
    channels out as ynet_tc_conv3x3_split reads them -- per source [hi | hi] over its 2 * sum(cp) stored channels, then lo
    over its sum(cp) hi channels, every part padded to its cp -- and the channel count of each of those blocks.  Cached
    per layout: fine-tuning re-packs the adapted layers' weights after every optimiser step."""
    key = (tuple(tuple(tuple(e) for e in layout) for layout in src_layouts), c_in, str(device))
    hit = _split_sel_cache.get(key)
    if hit is None:
        zero = 2 * c_in
        sel, chans, c0 = [], [], 0
        for layout in src_layouts:
            hi, lo = [], []
            for C, cp in layout:
                hi += list(range(c0, c0 + C)) + [zero] * (cp - C)
                lo += list(range(c_in + c0, c_in + c0 + C)) + [zero] * (cp - C)
                c0 += C
            sel += hi + hi + lo
            chans += [2 * len(hi), len(hi)]
        if c0 != c_in:
            raise ValueError(f'split_pack_weights: sources hold {c0} channels, the weight expects {c_in}')
        if len(_split_sel_cache) > 256:
            _split_sel_cache.clear()
        hit = _split_sel_cache[key] = (torch.tensor(sel, dtype=torch.long, device=device), chans)
    return hit


def split_pack_weights(weight_oihw, src_layouts):
    """OIHW float32 weight (3x3 or 1x1), input channels = the concatenation of the sources' real channels.
    src_layouts: per Split source its ``layout``.  Packed for the source pairs of ynet_tc_conv3x3_split: per source [W_hi | W_hi] over its
    2 * sum(cp) stored channels, then W_lo over its sum(cp) hi channels."""
    w = _req(weight_oihw, name='weight')
    C_out, c_in, kh, kw = w.shape
    sel, chans = _split_weight_select(src_layouts, c_in, w.device)
    w_hi = w.to(torch.bfloat16).to(torch.float32)
    w_lo = w - w_hi                                   # rounded to bf16 by the packer: W = hi + lo to 2^-17
    wcat = torch.cat([w_hi, w_lo, w.new_zeros(C_out, 1, kh, kw)], dim=1)
    return tc_pack_weights(wcat.index_select(1, sel), chans)


def _split_src_array(sources, N):
    arr = (_lib.TcSrc * (2 * len(sources)))()
    for i, s in enumerate(sources):
        for j, ch in enumerate((2 * s.cp, s.cp)):
            e = arr[2 * i + j]
            e.ptr = s.data.data_ptr()
            e.channels_pad = ch
            e.chunks_stored = ch // 8
            e.batch_stride = 0 if (s.N == 1 and N > 1) else s.data.stride(0)
            e.batch_mod = _tc_batch_mod(s, N)
    return arr


def tc_conv3x3_split(sources, packed_weight, bias_pad, C_out, relu, into=None):
    """conv3x3(cat(sources)) + bias (+ReLU) on Split activations -> Split: three bf16 tcgen05 MMAs per product term.
    into = (Split buffer, channel offset): write the result as that channel slice of a wider activation."""
    if not 1 <= len(sources) <= 2:
        raise ValueError('tc_conv3x3_split: one or two Split sources (concatenate more with split_cat)')
    N = max(s.N for s in sources)
    H, W = sources[0].H, sources[0].W
    if any(s.H != H or s.W != W for s in sources):
        raise ValueError('tc_conv3x3_split: sources must share the spatial size')
    arr = _split_src_array(sources, N)
    cpo = _pad16(C_out)
    if into is None:
        out, tot, off = torch.empty(N, cpo // 4, H, W, 8, dtype=torch.bfloat16, device=sources[0].data.device), 0, 0
    else:
        out, tot, off = into[0].data, into[0].cp, into[1]
        if out.shape[0] != N or out.shape[2] != H or out.shape[3] != W:
            raise ValueError('tc_conv3x3_split: the target activation has another shape')
    cin = sum(s.cp for s in sources)
    with _timed('tc_conv3x3_split_kernel', 3 * 2.0 * 9 * cin * C_out * H * W * N,
                sum(4.0 * s.cp * min(s.N, N) for s in sources) * H * W + 4.0 * cpo * H * W * N,
                tag=f'{cin}->{cpo}@{H}x{W} N={N}'):
        check(_L().ynet_tc_conv3x3_split(arr, 2 * len(sources), N, H, W, _ptr(packed_weight), _ptr(bias_pad), C_out,
                                         1 if relu else 0, _ptr(out), cpo, tot, off, 0, _stream()), 'tc_conv3x3_split')
    _count()
    return into[0] if into is not None else Split(out, C_out)


def tc_conv1x1_split_f32(a, packed_weight, bias_pad, C_out):
    """1x1 predictor on a Split activation -> float32 NCHW logits."""
    out = torch.empty(a.N, C_out, a.H, a.W, dtype=torch.float32, device=a.data.device)
    arr = _split_src_array([a], a.N)
    with _timed('tc_conv_kernel<1x1,f32,split>', 3 * 2.0 * a.cp * C_out * a.H * a.W * a.N,
                (4.0 * a.cp + 4.0 * C_out) * a.H * a.W * a.N):
        check(_L().ynet_tc_conv1x1_f32(arr, 2, a.N, a.H, a.W, _ptr(packed_weight), _ptr(bias_pad), C_out, _ptr(out), 0,
                                       _stream()), 'tc_conv1x1_f32')
    _count()
    return out


def tc_pack_hoisted_weights(weight_oihw, parts):
    """Packed weights of a conv whose sources mix 3x3 inputs and hoisted partial sums.

    parts: list in source order of ('conv', (c0, c1)) -- input channels [c0, c1) of ``weight_oihw`` applied as 3x3 --,
    ('i2c', (c0, c1)) -- the same channels read from an im2col ``center`` source (tc_rasterize_im2col) --,
    ('quad', (c0, n)) -- channels [c0, c0 + n) read from 2x2-neighbourhood planes (tc_rasterize_pyramid quad levels) --
    or ('partial', channels) -- a ``center`` source with ``channels`` = pad16(C_out) (hi) or twice that (hi | lo):
    identity on the centre tap."""
    C_out = weight_oihw.shape[0]
    cp = _pad16(C_out)
    bufs = []
    for kind, arg in parts:
        if kind == 'conv':
            c0, c1 = arg
            bufs.append(tc_pack_weights(weight_oihw[:, c0:c1].contiguous(), [c1 - c0]))
        elif kind == 'quad':      # 2x2-neighbourhood source: (first input channel, channels)
            c0, n_ch = arg
            bufs.append(tc_pack_weights(tc_quad_weights(weight_oihw, c0, n_ch), [4 * n_ch]))
        elif kind == 'i2c':       # im2col source: channel c * 9 + kh * 3 + kw <-> weight[:, c0 + c, kh, kw]
            c0, c1 = arg
            w1 = weight_oihw[:, c0:c1].reshape(C_out, (c1 - c0) * 9, 1, 1).contiguous()
            bufs.append(tc_pack_weights(w1, [(c1 - c0) * 9]))
        else:
            eye = torch.zeros(C_out, arg, 1, 1, dtype=torch.float32, device=weight_oihw.device)
            idx = torch.arange(C_out, device=weight_oihw.device)
            eye[idx, idx, 0, 0] = 1.0
            if arg == 2 * cp:
                eye[idx, cp + idx, 0, 0] = 1.0
            elif arg != cp:
                raise ValueError(f'partial source has {arg} channels, expected {cp} or {2 * cp}')
            bufs.append(tc_pack_weights(eye, [arg]))
    return torch.cat(bufs)


def tc_upconv_phase_weights(weight_oihw, bias):
    """(C_out, C_in, 3, 3) weight of an upsample_conv -> (w_eff (4*cp, C_in, 3, 3), bias_eff (4*cp,)) of the
    equivalent low-resolution phase conv (see ynet_tc_upconv_phase_weights)."""
    weight_oihw = _req(weight_oihw, name='weight')
    C_out, C_in = weight_oihw.shape[:2]
    cp = _pad16(C_out)
    w_eff = torch.empty(4 * cp, C_in, 3, 3, dtype=torch.float32, device=weight_oihw.device)
    b_eff = torch.empty(4 * cp, dtype=torch.float32, device=weight_oihw.device)
    check(_L().ynet_tc_upconv_phase_weights(_ptr(weight_oihw), _ptr(bias), C_out, C_in, _ptr(w_eff), _ptr(b_eff),
                                            _stream()), 'tc_upconv_phase_weights')
    _count()
    return w_eff, b_eff


def tc_upconv_border_weights(weight_oihw, src_channels):
    """float32 OIHW weight -> the shared-memory image of the border-ring kernel (see ynet_tc_upconv_border_weights)."""
    weight_oihw = _req(weight_oihw, name='weight')
    n = len(src_channels)
    real = (ctypes.c_int32 * n)(*src_channels)
    nbytes = _L().ynet_tc_upconv_border_weight_bytes(weight_oihw.shape[0], n, real)
    out = torch.empty(nbytes // 4, dtype=torch.float32, device=weight_oihw.device)
    check(_L().ynet_tc_upconv_border_weights(_ptr(weight_oihw), weight_oihw.shape[0], n, real, _ptr(out), _stream()),
          'tc_upconv_border_weights')
    _count()
    return out


def tc_upconv3x3(sources, packed_phase_weight, bias_eff, border_weight, bias, C_out, exact_ring=True):
    """bilinear x2 + conv3x3 (ynet.py:463-464) of low-resolution C8 sources -> C8 (N, C_out, 2h, 2w), without
    materialising the upsampled tensor."""
    if exact_ring and all(s.rep == 1 for s in sources):
        sources = [tc_pad_replicate(s) for s in sources]     # no-op for producers that already wrote the padded layout
    else:
        sources = [s.unpadded() for s in sources]
    N = max(s.N for s in sources)
    h, w = sources[0].H, sources[0].W
    arr = (_lib.TcSrc * len(sources))()
    for i, s in enumerate(sources):
        if s.H != h or s.W != w:
            raise ValueError('tc_upconv3x3: sources must share the spatial size')
        arr[i].ptr = s.data.data_ptr()
        arr[i].channels_pad = s.C_pad
        arr[i].batch_stride = 0 if (s.N == 1 and N > 1) else s.data.stride(0)
        arr[i].batch_mod = _tc_batch_mod(s, N)
        arr[i].padded = s.pad
    real = (ctypes.c_int32 * len(sources))(*[s.C for s in sources])
    cp = _pad16(C_out)
    out = torch.empty(N, cp // 8, 2 * h, 2 * w, 8, dtype=torch.bfloat16, device=sources[0].data.device)
    cin_pad = sum(s.C_pad for s in sources)
    args = (arr, real, len(sources), N, h, w, _ptr(packed_phase_weight), _ptr(bias_eff), _ptr(border_weight), _ptr(bias),
            C_out, 0, _ptr(out))
    key = ('up', tuple(s.C_pad for s in sources), cp, h, w, min(N, 64), sources[0].pad)
    tune = _tc_tune.get(key)
    if tune is None:
        tune = (_tc_autotune(key, args, fn='ynet_tc_upconv3x3', n_pad=4 * cp)
                if (tc_autotune_enabled and not torch.cuda.is_current_stream_capturing()) else 0)
    with _timed('tc_upconv3x3_kernel', 2.0 * 9 * sum(s.C for s in sources) * C_out * 4 * h * w * N,
                (2.0 * cin_pad + 8.0 * cp) * h * w * N, tag=f'{cin_pad}->{cp}@up{2 * h}x{2 * w} N={N}'):
        check(_L().ynet_tc_upconv3x3(*args, tune, _stream()), 'tc_upconv3x3')
    _count(2)
    return C8(out, C_out)


# ---- row-marching kh-stacked conv (rowconv_tc.cu): the C_out <= 32 layers of the full-resolution decoder levels ----------
def tc_rowconv_supported(a, C_out):
    """Can ``tc_rowconv3x3*`` take the plain C8 activation(s) ``a`` (one C8 or a list = channel concat)?"""
    srcs = a if isinstance(a, (list, tuple)) else [a]
    if not (1 <= len(srcs) <= 3) or C_out > 32:
        return False
    for i, s in enumerate(srcs):
        if isinstance(s, WpPlanes) and i == len(srcs) - 1 and 1 <= i <= 2 and s.W >= 120:
            continue                 # loaded from the template planes: last source, behind one or two tensor sources
        if not isinstance(s, C8) or s.pad or s.center or s.taps:
            return False
    return (sum(s.K_pad for s in srcs) <= 64 and srcs[0].H >= 2
            and all(s.H == srcs[0].H and s.W == srcs[0].W for s in srcs))


def tc_rowconv_pack_weights(weight_oihw, C_in_pad):
    """Effective OIHW float32 3x3 weight (C_out <= 32) -> bf16 [K block][kw][2][96 = kh * 32 + co][8]."""
    weight_oihw = _req(weight_oihw, name='weight')
    C_out, C_in = weight_oihw.shape[:2]
    packed = torch.empty(_L().ynet_tc_rowconv_packed_weight_bytes(C_in_pad), dtype=torch.uint8, device=weight_oihw.device)
    check(_L().ynet_tc_rowconv_pack_weights(_ptr(weight_oihw), C_out, C_in, C_in_pad, _ptr(packed), _stream()),
          'tc_rowconv_pack_weights')
    _count()
    return packed


def tc_rowconv_pack_weights_cat(weight_oihw, parts):
    """Packed weights of a multi-source row conv: ``parts`` = [(c0, c1, K_pad)] per source, in source order -- source i
    carries input channels [c0, c1) of ``weight_oihw`` and occupies K_pad (multiple of 16) channels of the K axis."""
    w = weight_oihw
    cols = []
    for c0, c1, kp in parts:
        blk = torch.zeros(w.shape[0], kp, 3, 3, dtype=torch.float32, device=w.device)
        blk[:, :c1 - c0] = w[:, c0:c1]
        cols.append(blk)
    wcat = torch.cat(cols, dim=1).contiguous()
    return tc_rowconv_pack_weights(wcat, wcat.shape[1])


def _rowconv_srcs(srcs, N):
    arr = (_lib.TcSrc * len(srcs))()
    for i, a in enumerate(srcs):
        arr[i].ptr = a.data.data_ptr()
        arr[i].channels_pad = a.K_pad
        arr[i].chunks_stored = a.C_pad // 8
        arr[i].batch_stride = 0 if (a.N == 1 and N > 1) else a.data.stride(0)
        arr[i].batch_mod = _tc_batch_mod(a, N)
    return arr


def tc_rowconv3x3(a, packed_weight, bias32, C_out, relu, pad_out=False, partial=None):
    """conv3x3 + bias (+ ReLU) with C_out <= 32 of one plain C8 source or a list of <= 3 (= channel concat) -> C8
    (see ynet_tc_rowconv3x3).  ``partial``: hoisted partial sums (tc_conv3x3_hilo output) added before the activation."""
    srcs = list(a) if isinstance(a, (list, tuple)) else [a]
    wp = srcs.pop() if isinstance(srcs[-1], WpPlanes) else None      # waypoint planes gathered in the kernel
    N = max(s.N for s in srcs + ([partial] if partial is not None else []))
    H, W = srcs[0].H, srcs[0].W
    cp = _pad16(C_out)
    po = 1 if pad_out else 0
    out = torch.empty(N, cp // 8, H + 2 * po, W + 2 * po, 8, dtype=torch.bfloat16, device=srcs[0].data.device)
    parr = None
    in_bytes = sum(2.0 * s.C_pad * min(s.data.shape[0], N) for s in srcs) * H * W
    if partial is not None:
        parr = _rowconv_srcs([partial], N)
        parr[0].channels_pad = partial.C_pad
        in_bytes += 2.0 * partial.C_pad * min(partial.data.shape[0], N) * H * W
    k_pad = sum(s.K_pad for s in srcs) + (16 if wp is not None else 0)
    tag = f'{k_pad}{"+P" if partial is not None else ""}{"+W" if wp is not None else ""}->{cp}@{H}x{W} N={N}'
    with _timed('tc_rowconv_kernel', 2.0 * 9 * (sum(s.C for s in srcs) + (wp.C if wp is not None else 0)) * C_out * H * W * N,
                in_bytes + 2.0 * cp * H * W * N, tag=tag):
        if wp is None:
            check(_L().ynet_tc_rowconv3x3(_rowconv_srcs(srcs, N), len(srcs), parr, N, H, W, _ptr(packed_weight),
                                          _ptr(bias32), C_out, (1 if relu else 0) | (2 * po), _ptr(out), cp, _stream()),
                  'tc_rowconv3x3')
        else:
            if wp.N != N or wp.H != H or wp.W != W:
                raise ValueError('tc_rowconv3x3: the waypoint planes cover another batch / resolution')
            w = _lib.TcWpSrc(wp.planes.data_ptr(), wp.coords.data_ptr(), wp.template.shape[0], wp.template.shape[1],
                             wp.C, wp.level)
            check(_L().ynet_tc_rowconv3x3_wp(_rowconv_srcs(srcs, N), len(srcs), parr, ctypes.byref(w), N, H, W,
                                             _ptr(packed_weight), _ptr(bias32), C_out, (1 if relu else 0) | (2 * po),
                                             _ptr(out), cp, _stream()), 'tc_rowconv3x3_wp')
    _count()
    return C8(out, C_out, 1, False, po)


def _rowconv2_args(srcs, partial):
    srcs = list(srcs)
    wp = srcs.pop()
    if not isinstance(wp, WpPlanes) or not 1 <= len(srcs) <= 2:
        raise ValueError('tc_rowconv2: sources = one or two C8 tensors followed by WpPlanes')
    N = max(s.N for s in srcs + ([partial] if partial is not None else []))
    H, W = srcs[0].H, srcs[0].W
    if wp.N != N or wp.H != H or wp.W != W:
        raise ValueError('tc_rowconv2: the waypoint planes cover another batch / resolution')
    parr = None
    in_bytes = sum(2.0 * s.C_pad * min(s.data.shape[0], N) for s in srcs) * H * W
    if partial is not None:
        parr = _rowconv_srcs([partial], N)
        parr[0].channels_pad = partial.C_pad
        in_bytes += 2.0 * partial.C_pad * min(partial.data.shape[0], N) * H * W
    w = _lib.TcWpSrc(wp.planes.data_ptr(), wp.coords.data_ptr(), wp.template.shape[0], wp.template.shape[1], wp.C, wp.level)
    k_pad = sum(s.K_pad for s in srcs) + 16
    c_in = sum(s.C for s in srcs) + wp.C
    return srcs, wp, w, parr, N, H, W, in_bytes, k_pad, c_in


def tc_rowconv2_supported(srcs, C_mid, C_out):
    """Can the two-conv row kernel take conv A's sources (C8 tensors + WpPlanes last) with 32 channels in between?"""
    return (isinstance(srcs[-1], WpPlanes) and tc_rowconv_supported(srcs, C_mid) and C_mid == 32 and C_out <= 32
            and srcs[0].W >= 120)


def tc_rowconv2_wp(srcs, packed_a, bias32_a, packed_b, bias32_b, C_out, relu, pad_out=False, partial=None):
    """conv3x3(cat(srcs)) + partial + ReLU -> conv3x3 (+ReLU) in one kernel (ynet_tc_rowconv2_wp) -> C8 (N, C_out)."""
    srcs, wp, w, parr, N, H, W, in_bytes, k_pad, c_in = _rowconv2_args(srcs, partial)
    cp = _pad16(C_out)
    po = 1 if pad_out else 0
    out = torch.empty(N, cp // 8, H + 2 * po, W + 2 * po, 8, dtype=torch.bfloat16, device=srcs[0].data.device)
    tag = f'{k_pad}{"+P" if partial is not None else ""}+W->32->{cp}@{H}x{W} N={N}'
    with _timed('tc_rowconv2_kernel', 2.0 * 9 * (c_in * 32 + 32 * C_out) * H * W * N, in_bytes + 2.0 * cp * H * W * N, tag=tag):
        check(_L().ynet_tc_rowconv2_wp(_rowconv_srcs(srcs, N), len(srcs), parr, ctypes.byref(w), N, H, W, _ptr(packed_a),
                                       _ptr(bias32_a), _ptr(packed_b), _ptr(bias32_b), C_out, (1 if relu else 0) | (2 * po),
                                       _ptr(out), cp, _stream()), 'tc_rowconv2_wp')
    _count()
    return C8(out, C_out, 1, False, po)


def tc_rowconv2_wp_pred_softargmax(srcs, packed_a, bias32_a, packed_b, bias32_b, relu_b, packed_pred, pred_bias_pad, C_pred,
                                   partial=None):
    """conv + partial + ReLU -> conv (+ReLU) -> 1x1 predictor -> SoftArgmax2D in one kernel -> (N, C_pred, 2)."""
    srcs, wp, w, parr, N, H, W, in_bytes, k_pad, c_in = _rowconv2_args(srcs, partial)
    out = torch.empty(N, C_pred, 2, dtype=torch.float32, device=srcs[0].data.device)
    nb = _L().ynet_tc_rowconv2_softargmax_workspace_bytes(N, C_pred, W)
    ws = _workspace(nb, out.device, 'tc_rowconv2_softargmax')
    tag = f'{k_pad}{"+P" if partial is not None else ""}+W->32->32->{C_pred}@{H}x{W} N={N}'
    with _timed('tc_rowconv2_kernel<pred,softargmax>', 2.0 * (9 * (c_in * 32 + 32 * 32) + 32 * C_pred) * H * W * N, in_bytes,
                tag=tag):
        check(_L().ynet_tc_rowconv2_wp_pred_softargmax(_rowconv_srcs(srcs, N), len(srcs), parr, ctypes.byref(w), N, H, W,
                                                       _ptr(packed_a), _ptr(bias32_a), _ptr(packed_b), _ptr(bias32_b),
                                                       1 if relu_b else 0, _ptr(packed_pred), _ptr(pred_bias_pad), C_pred,
                                                       _ptr(out), _ptr(ws), ws.numel(), _stream()),
              'tc_rowconv2_wp_pred_softargmax')
    _count(2)
    return out


def tc_rowconv3x3_pred_softargmax(a, packed_weight, bias32, C_out, relu, packed_pred, pred_bias_pad, C_pred):
    """conv3x3 (+bias, +ReLU) -> 1x1 predictor -> SoftArgmax2D in one kernel: C8 (N, <= 64 ch) -> (N, C_pred, 2)."""
    out = torch.empty(a.N, C_pred, 2, dtype=torch.float32, device=a.data.device)
    nb = _L().ynet_tc_rowconv_softargmax_workspace_bytes(a.N, C_pred, a.W)
    ws = _workspace(nb, out.device, 'tc_rowconv_softargmax')
    flops = 2.0 * (9 * a.C * C_out + C_out * C_pred) * a.H * a.W * a.N
    with _timed('tc_rowconv_kernel<pred,softargmax>', flops, 2.0 * a.C_pad * min(a.data.shape[0], a.N) * a.H * a.W,
                tag=f'{a.K_pad}->{_pad16(C_out)}->{C_pred}@{a.H}x{a.W} N={a.N}'):
        check(_L().ynet_tc_rowconv3x3_pred_softargmax(_rowconv_srcs([a], a.N), a.N, a.H, a.W, _ptr(packed_weight), _ptr(bias32),
                                                      C_out, 1 if relu else 0, _ptr(packed_pred), _ptr(pred_bias_pad), C_pred,
                                                      _ptr(out), _ptr(ws), ws.numel(), _stream()),
              'tc_rowconv3x3_pred_softargmax')
    _count(2)
    return out


tc_autotune_enabled = os.environ.get('YNET_TC_AUTOTUNE', '1') != '0'
_tc_tune = {}


def _tc_autotune(key, args, fn='ynet_tc_conv3x3', n_pad=None):
    """Pick (accumulators per tile, CTAs per SM, stages) for one layer shape by timing the candidates once."""
    cp = key[1] if n_pad is None else n_pad
    launch = getattr(_L(), fn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cands = [j | (ctas << 4) | (stages << 8) for j in (1, 2, 3) if j <= max(1, 512 // (2 * cp))
             for ctas in (2, 1) for stages in (6, 3)]

    def run(tune, reps):
        e0.record()
        for _ in range(reps):
            check(launch(*args, tune, _stream()), fn + '(autotune)')
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    # bring the SM clocks up first: timing the first candidates on an idle GPU would bias the choice
    spent = 0.0
    while spent < 40.0:
        spent += run(0, 4) * 4
    times = {t: float('inf') for t in cands}
    for _ in range(2):
        for t in cands:
            times[t] = min(times[t], run(t, 3))
    best = min(times, key=times.get)
    _tc_tune[key] = best
    return best


def tc_predictor_f32(a, weight, bias):
    """1x1 predictor on a C8 activation -> float32 NCHW logits."""
    weight = _req(weight, name='weight')
    C_out, C_in = weight.shape
    out = torch.empty(a.N, C_out, a.H, a.W, dtype=torch.float32, device=a.data.device)
    with _timed('c8_predictor_kernel', 2.0 * C_in * C_out * a.H * a.W * a.N, (2.0 * a.C_pad + 4.0 * C_out) * a.H * a.W * a.N):
        check(_L().ynet_tc_predictor_f32(_ptr(a.data), a.N, a.C_pad, C_in, a.H, a.W, _ptr(weight), _ptr(bias), C_out,
                                         _ptr(out), _stream()), 'tc_predictor_f32')
    _count()
    return out
