"""Training-mode executor: the same layer walk as ``engine.YNetEngine`` but every layer is a
``torch.autograd.Function`` whose forward AND backward are libynet_b200.so kernels, so that
``loss.backward()`` in the reference's train_epoch.py:109-115 works unchanged.

torch.autograd is plumbing here (graph bookkeeping, gradient accumulation of the few LoRA
tensors); conv forward / dgrad / wgrad, pool / bilinear backward, BCE and the LoRA gradient
projection (dA = s B^T dM, dB = s dM A^T) are CUDA kernels.
"""
import os
import weakref

import torch

from . import ops
from .engine import ChannelCat, _parts, _parallel_as_3x3
from .ops import SRC_DIRECT, SRC_POOL2, SRC_UP2


def _materialize(t, mode, N):
    """Concrete (N, C, H, W) view of a conv source at the conv's resolution (only needed for wgrad)."""
    if mode == SRC_POOL2:
        t = ops.maxpool2x2(t.contiguous())
    elif mode == SRC_UP2:
        t = ops.upsample_bilinear2x(t.contiguous())
    if t.shape[0] == 1 and N > 1:
        t = t.expand(N, -1, -1, -1)
    return t


# Tensor-core training convs (model.set_backend('bf16x3'), or YNET_TRAIN_TC=1): the forward conv and the data gradient of
# every layer run as split-bf16 tcgen05 convs (three bf16 MMAs per product, fp32 accumulation: ~1e-5 of the fp32 result,
# inside the fixture tolerances of tests/test_gpu_train.py) instead of the CUDA-core fp32 kernels; the weight gradient of
# the few trainable (LoRA / adapter) layers stays on the fp32 wgrad kernel.  The autograd interface stays float32 NCHW:
# operands are split on entry and results joined on exit (bandwidth-bound passes).
TRAIN_TC = os.environ.get('YNET_TRAIN_TC', '0') == '1'
_tc_wcache = {}
_zero_bias = {}


def _select_executor(model):
    """The conv kernels of the training graph follow the model's backend (called at every entry point below)."""
    global TRAIN_TC
    TRAIN_TC = getattr(model, '_backend', 'fp32') == 'bf16x3' or os.environ.get('YNET_TRAIN_TC', '0') == '1'


def _tc_packed(kind, weight, ver, make):
    """Packed split weights of a PARAMETER, cached per tensor object (weak reference: a new model whose parameter lands on
    a recycled address must not hit) and version; temporaries (adapter sums) are packed every time."""
    if weight is None:
        return make()
    key = (kind, id(weight))
    hit = _tc_wcache.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == ver:
        return hit[2]
    if len(_tc_wcache) > 512:
        for k in [k for k, v in _tc_wcache.items() if v[0]() is None]:
            del _tc_wcache[k]
    val = make()
    _tc_wcache[key] = (weakref.ref(weight), ver, val)
    return val


_range_idx_cache = {}


def _range_index(ranges, device):
    """arange over the concatenated channel ranges, cached (built once per layer instead of once per optimiser step)."""
    key = (tuple(ranges), str(device))
    hit = _range_idx_cache.get(key)
    if hit is None:
        if len(_range_idx_cache) > 256:
            _range_idx_cache.clear()
        hit = _range_idx_cache[key] = torch.cat([torch.arange(c0, c1, device=device) for c0, c1 in ranges])
    return hit


def _tc_conv_forward(fold, C_out, weight, wver, bias, relu, parts):
    """conv3x3(cat(parts)) + bias (+ReLU): float32 NCHW parts (batch 1 = broadcast) -> float32 NCHW.  ``fold()`` yields the
    effective OIHW weight (LoRA folded); it only runs when the packed weights are not cached for this version."""
    from .engine import YNetEngineSplit
    sp = [ops.split_pack(t) for t in parts]
    sources, ranges = YNetEngineSplit._group(sp)
    layouts = tuple(tuple(s.layout) for s in sources)

    def make():
        w_eff = fold()
        idx = _range_index(ranges, w_eff.device)
        bias_pad = torch.zeros(ops._pad16(C_out), dtype=torch.float32, device=w_eff.device)
        if bias is not None:
            bias_pad[:C_out] = bias.detach()
        return ops.split_pack_weights(w_eff.index_select(1, idx).contiguous(), [s.layout for s in sources]), bias_pad
    bver = None if bias is None else (id(bias), bias._version)
    packed, bias_pad = _tc_packed('f', weight, (wver, bver, layouts, ranges), make)
    return ops.split_unpack(ops.tc_conv3x3_split(sources, packed, bias_pad, C_out, relu))


def _tc_conv_dgrad(fold, C_in, weight, wver, dy, relu_out):
    """d/dx of conv3x3 (+ReLU): dy (masked by the activation) convolved with the flipped, transposed weight."""
    dys = ops.split_pack_masked(dy, relu_out) if relu_out is not None else ops.split_pack(dy)

    def make():
        w_t = fold().flip(2, 3).transpose(0, 1).contiguous()                # (C_in, C_out, 3, 3)
        return ops.split_pack_weights(w_t, [dys.layout])
    packed = _tc_packed('d', weight, (wver, tuple(dys.layout)), make)
    zero = _zero_bias.get((dy.device, C_in))
    if zero is None:
        zero = _zero_bias[(dy.device, C_in)] = torch.zeros(ops._pad16(C_in), dtype=torch.float32, device=dy.device)
    return ops.split_unpack(ops.tc_conv3x3_split([dys], packed, zero, C_in, False))


class Conv3x3Fn(torch.autograd.Function):
    """conv3x3(cat(sources)) (+ReLU) with LoRA-folded weight; sources may be pooled / upsampled on load."""

    @staticmethod
    def forward(ctx, weight, bias, lora_A, lora_B, relu, modes, H, W, *sources):
        N = max(s.shape[0] for s in sources)
        ctx.tc = (TRAIN_TC and ops.tc_supported() and weight.shape[0] <= 256 and 3 * sum(
            ops._pad16(s.shape[1]) for s in sources) // 16 <= 64)
        if ctx.tc:
            ctx.wver = (weight._version, None if lora_A is None else (id(lora_A), lora_A._version),
                        None if lora_B is None else (id(lora_B), lora_B._version))
            parts = [_materialize(s.contiguous(), m, 1) for s, m in zip(sources, modes)]
            y = _tc_conv_forward(lambda: ops.lora_fold(weight, lora_A, lora_B, packed=False), weight.shape[0],
                                 weight if weight.is_leaf else None, ctx.wver, bias, relu, parts)
        else:
            packed = ops.lora_fold(weight, lora_A, lora_B, packed=True)
            y = ops.conv3x3_f32(list(zip(sources, modes)), packed, bias, relu, N, H, W)
        ctx.relu, ctx.modes, ctx.N = relu, modes, N
        ctx.has_lora = lora_A is not None
        ctx.save_for_backward(weight, bias, lora_A, lora_B, y, *sources)
        return y

    @staticmethod
    def backward(ctx, dy):
        weight, bias, lora_A, lora_B, y, *sources = ctx.saved_tensors
        dy = dy.contiguous()
        relu_out = y if ctx.relu else None
        need = ctx.needs_input_grad
        d_weight = d_bias = d_A = d_B = None
        d_sources = [None] * len(sources)
        if any(need[8:]):
            def fold():
                return ops.lora_fold(weight, lora_A, lora_B, packed=False)
            dx = (_tc_conv_dgrad(fold, weight.shape[1], weight if weight.is_leaf else None, ctx.wver, dy, relu_out) if ctx.tc
                  else ops.conv3x3_dgrad_f32(dy, relu_out, fold()))
            c0 = 0
            for i, (s, mode) in enumerate(zip(sources, ctx.modes)):
                c1 = c0 + s.shape[1]
                if need[8 + i]:
                    g = dx[:, c0:c1]
                    if mode == SRC_POOL2:
                        g = ops.maxpool2x2_bwd(s.contiguous(), g.contiguous())
                    elif mode == SRC_UP2:
                        g = ops.upsample_bilinear2x_bwd(g.contiguous())
                    if s.shape[0] == 1 and ctx.N > 1:
                        g = g.sum(dim=0, keepdim=True)
                    d_sources[i] = g
                c0 = c1
        want_w = need[0] or (ctx.has_lora and (need[2] or need[3]))
        want_b = bias is not None and need[1]
        if want_w or want_b:
            parts = [_materialize(s, m, ctx.N) for s, m in zip(sources, ctx.modes)]
            x = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
            dW, db = ops.conv3x3_wgrad_f32(x.contiguous(), dy, relu_out, want_bias=want_b)
            if want_b:
                d_bias = db
            if need[0]:
                d_weight = dW
            if ctx.has_lora and (need[2] or need[3]):
                d_A, d_B = ops.lora_grad(dW, lora_A, lora_B)
        return (d_weight, d_bias, d_A, d_B, None, None, None, None) + tuple(d_sources)


class Conv1x1Fn(torch.autograd.Function):
    """The 1x1 predictor (ynet.py:450-451).  Frozen on the MoSA path: backward = dgrad only."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w2 = weight.reshape(weight.shape[0], -1)
        ctx.save_for_backward(x, weight)
        return ops.conv1x1_f32(x, w2, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        w2 = weight.reshape(weight.shape[0], -1)
        dx = ops.conv1x1_f32(dy, w2.t().contiguous(), None) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1]:      # full training only ('all'/'train'): off the MoSA hot path
            dw = torch.einsum('nohw,nihw->oi', dy, x).reshape(weight.shape)
        if ctx.needs_input_grad[2]:
            db = dy.sum(dim=(0, 2, 3))
        return dx, dw, db


class MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.maxpool2x2(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.maxpool2x2_bwd(x, dy.contiguous())


class BCEWithLogitsFn(torch.autograd.Function):
    """mean BCE-with-logits; d/dlogits computed in the same pass as the loss."""

    @staticmethod
    def forward(ctx, logits, target):
        loss, grad = ops.bce_logits_fwd_bwd(logits, target, 1.0, want_grad=True)
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return grad * dloss, None


class BCEWithLogitsLoss(torch.nn.Module):
    """Stand-in for nn.BCEWithLogitsLoss() (models/trainer.py:206) running the fused kernel."""

    def forward(self, logits, target):
        return BCEWithLogitsFn.apply(logits, target)


def _conv(module, sources, relu, H, W):
    tensors = [t for t, _ in sources]
    modes = tuple(m for _, m in sources)
    weight = module.weight
    if hasattr(module, 'serial_layer'):
        raise NotImplementedError('serial adapters (AdapterLayer) normalise with batch statistics in training mode: '
                                  'fine-tuning them is outside the B200 hot path')
    if hasattr(module, 'parallel_layer'):
        # AdapterLayer, parallel (ynet.py:121-130): conv(x, W) + sum_k conv_k(x, W_k) = conv(x, W + sum_k pad(W_k)); the
        # weight-space sum is differentiable, so dW_k falls out of the conv's wgrad kernel through autograd
        weight = weight + _parallel_as_3x3(module.parallel_layer, differentiable=True)
    return Conv3x3Fn.apply(weight, module.bias, getattr(module, 'lora_A', None), getattr(module, 'lora_B', None),
                           relu, modes, H, W, *tensors)


def _run_stages(stages, x_parts):
    feats = []
    cur = x_parts
    for stage in stages:
        mods = list(stage)
        convs = [m for m in mods if isinstance(m, torch.nn.Conv2d)]
        has_pool = any(isinstance(m, torch.nn.MaxPool2d) for m in mods)
        H, W = cur[0].shape[2], cur[0].shape[3]
        if has_pool:
            H, W = H // 2, W // 2
        if not convs:
            x = cur[0] if len(cur) == 1 else ChannelCat(cur).materialize()
            y = MaxPoolFn.apply(x.contiguous())
            feats.append(y)
            cur = [y]
            continue
        for ci, conv in enumerate(convs):
            srcs = [(t, (SRC_POOL2 if has_pool else SRC_DIRECT) if ci == 0 else SRC_DIRECT) for t in cur]
            cur = [_conv(conv, srcs, True, H, W)]
        feats.append(cur[0])
    return feats


def pred_features(model, scene_map, motion_map):
    _select_executor(model)
    enc = model.encoder
    scene, motion = _parts(scene_map), _parts(motion_map)
    if model.network == 'fusion':
        sf = _run_stages(enc.scene_stages, scene)
        mf = _run_stages(enc.motion_stages, motion)
        feats = [ChannelCat((a, b)) for a, b in zip(sf, mf)]
        return feats + _run_stages(enc.fusion_stages, list(feats[-1]))
    if getattr(enc, 'adapters', None) is not None:
        raise NotImplementedError('fine-tuning the block-level serial / parallel adapters of YNetEncoderB is outside the '
                                  'B200 hot path (inference is supported)')
    return _run_stages(enc.stages, scene + motion)


def decoder_logits(model, decoder, key, features):
    _select_executor(model)
    feats = [_parts(f) for f in features][::-1]
    c = feats[0]
    H, W = c[0].shape[2], c[0].shape[3]
    x = _conv(decoder.center[0], [(t, SRC_DIRECT) for t in c], True, H, W)
    x = _conv(decoder.center[2], [(x, SRC_DIRECT)], True, H, W)
    for i, skip in enumerate(feats[1:]):
        H, W = skip[0].shape[2], skip[0].shape[3]
        up = _conv(decoder.upsample_conv[i], [(x, SRC_UP2)], False, H, W)
        x = _conv(decoder.decoder[i][0], [(up, SRC_DIRECT)] + [(t, SRC_DIRECT) for t in skip], True, H, W)
        x = _conv(decoder.decoder[i][2], [(x, SRC_DIRECT)], True, H, W)
    p = decoder.predictor
    return Conv1x1Fn.apply(x, p.weight, p.bias)
