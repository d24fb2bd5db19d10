"""Micro-benchmark of the row-marching conv kernel (rowconv_tc.cu) next to the tile kernel on the same layer (CUDA events).

    N=640 python tools/bench_rowconv.py            # env: N images, MODE=plain|fused|both|tile, HW=416
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion_style_transfer_b200 import ops  # noqa: E402

N = int(os.environ.get('N', 320))
HW = int(os.environ.get('HW', 416))
MODE = os.environ.get('MODE', 'both')
REPS = int(os.environ.get('REPS', 5))
torch.manual_seed(0)
_w = torch.randn(4096, 4096, device='cuda')
for _ in range(20):
    _w @ _w          # clock warm-up
torch.cuda.synchronize()
x = ops.tc_pack(torch.relu(torch.randn(N, 32, HW, HW, device='cuda')))
w = torch.randn(32, 32, 3, 3, device='cuda') * 0.1
bias = torch.randn(32, device='cuda') * 0.1
wp = torch.randn(30, 32, 1, 1, device='cuda') * 0.5
pb = torch.zeros(32, device='cuda')
packed_row = ops.tc_rowconv_pack_weights(w, 32)
packed_tile = ops.tc_pack_weights(w, [32])
ppacked = ops.tc_pack_weights(wp, [32])


def timeit(name, fn, flops, nbytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    print(f'{name:>28} N={N} @{HW}: {ms:8.3f} ms  {flops / ms / 1e9:7.1f} TF/s  {nbytes / ms / 1e6:7.0f} GB/s')


S = HW * HW * N
f_conv = 2.0 * 9 * 32 * 32 * S
if MODE in ('both', 'tile'):
    timeit('tile conv', lambda: ops.tc_conv3x3([x], packed_tile, bias, 32, True), f_conv, 128.0 * S)
    y = ops.tc_conv3x3([x], packed_tile, bias, 32, True)
    timeit('tile pred+softargmax', lambda: ops.tc_conv1x1_softargmax(y, ppacked, pb, 30), 2.0 * 32 * 30 * S, 64.0 * S)
if MODE in ('both', 'plain'):
    timeit('row conv', lambda: ops.tc_rowconv3x3(x, packed_row, bias, 32, True), f_conv, 128.0 * S)
if MODE in ('both', 'fused'):
    timeit('row conv+pred+softargmax', lambda: ops.tc_rowconv3x3_pred_softargmax(x, packed_row, bias, 32, True, ppacked, pb, 30),
           f_conv + 2.0 * 32 * 30 * S, 64.0 * S)

if MODE in ('both', 'l2'):
    # decoder.4.0 of the trajectory decoder: cat(up 16, feature 32 (hoisted), waypoints 2) -> 32, agent-major (G = 20)
    G = 20
    nb = N // G
    up = ops.tc_pack(torch.randn(nb * G, 16, HW, HW, device='cuda'))
    wpl = ops.tc_pack(torch.rand(nb * G, 2, HW, HW, device='cuda'))
    wpl = ops.C8(wpl.data[:, :1].contiguous(), 2)          # one stored plane, as tc_rasterize_pyramid writes it
    feat = ops.tc_pack(torch.relu(torch.randn(nb, 32, HW, HW, device='cuda')))
    w2 = torch.randn(32, 50, 3, 3, device='cuda') * 0.1
    part = ops.tc_conv3x3_hilo([feat], ops.tc_pack_weights(w2[:, 16:48].contiguous(), [32]), 32, False).repeat_interleave(G)
    packed2 = ops.tc_rowconv_pack_weights_cat(w2, [(0, 16, 16), (48, 50, 16)])
    S2 = HW * HW * nb * G
    timeit('row conv [up|P|wp] -> 32', lambda: ops.tc_rowconv3x3([up, wpl], packed2, bias, 32, True, partial=part),
           2.0 * 9 * 18 * 32 * S2, (32.0 + 16.0 + 64.0) * S2)

if MODE in ('wp',):
    # the same with the waypoint planes gathered inside the kernel (ynet_tc_rowconv3x3_wp)
    G = 20
    nb = N // G
    up = ops.tc_pack(torch.randn(nb * G, 16, HW, HW, device='cuda'))
    tmpl = ops.create_dist_template(int(2.5 * HW), 'cuda')
    coords = (torch.rand(nb * G * 2, 2, device='cuda') * (HW - 1)).contiguous()
    pyr = ops.tc_rasterize_pyramid(tmpl, coords, nb * G, 2, HW, HW, 1)
    lazy = ops.tc_rasterize_pyramid(tmpl, coords, nb * G, 2, HW, HW, 1, lazy_levels=1)
    feat = ops.tc_pack(torch.relu(torch.randn(nb, 32, HW, HW, device='cuda')))
    w2 = torch.randn(32, 50, 3, 3, device='cuda') * 0.1
    part = ops.tc_conv3x3_hilo([feat], ops.tc_pack_weights(w2[:, 16:48].contiguous(), [32]), 32, False).repeat_interleave(G)
    packed2 = ops.tc_rowconv_pack_weights_cat(w2, [(0, 16, 16), (48, 50, 16)])
    S2 = HW * HW * nb * G
    timeit('row conv [up|P|wp planes]', lambda: ops.tc_rowconv3x3([up, pyr[0]], packed2, bias, 32, True, partial=part),
           2.0 * 9 * 18 * 32 * S2, (32.0 + 16.0 + 64.0) * S2)
    packed3 = ops.tc_rowconv_pack_weights_cat(w2, [(0, 16, 16)] + lazy[0].weight_parts(48))
    timeit('row conv [up|P|wp template]', lambda: ops.tc_rowconv3x3([up, lazy[0]], packed3, bias, 32, True, partial=part),
           2.0 * 9 * 18 * 32 * S2, (32.0 + 64.0) * S2)

if MODE in ('rc2',):
    # decoder.4.0 + decoder.4.2 (+ predictor + soft-argmax) in one kernel against the separate launches
    G = 20
    nb = N // G
    up = ops.tc_pack(torch.randn(nb * G, 16, HW, HW, device='cuda'))
    tmpl = ops.create_dist_template(int(2.5 * HW), 'cuda')
    coords = (torch.rand(nb * G * 2, 2, device='cuda') * (HW - 1)).contiguous()
    lazy = ops.tc_rasterize_pyramid(tmpl, coords, nb * G, 2, HW, HW, 1, lazy_levels=1)[0]
    feat = ops.tc_pack(torch.relu(torch.randn(nb, 32, HW, HW, device='cuda')))
    w2 = torch.randn(32, 50, 3, 3, device='cuda') * 0.1
    part = ops.tc_conv3x3_hilo([feat], ops.tc_pack_weights(w2[:, 16:48].contiguous(), [32]), 32, False).repeat_interleave(G)
    pa = ops.tc_rowconv_pack_weights_cat(w2, [(0, 16, 16)] + lazy.weight_parts(48))
    S2 = HW * HW * nb * G
    fl = 2.0 * 9 * (18 * 32 + 32 * 32) * S2
    timeit('L1 row conv [up|P|wp template]', lambda: ops.tc_rowconv3x3([up, lazy], pa, bias, 32, True, partial=part),
           2.0 * 9 * 18 * 32 * S2, (32.0 + 64.0) * S2)
    timeit('two-conv block -> HBM', lambda: ops.tc_rowconv2_wp([up, lazy], pa, bias, packed_row, bias, 32, True, partial=part),
           fl, (32.0 + 64.0) * S2)
    timeit('two-conv block + pred + softargmax', lambda: ops.tc_rowconv2_wp_pred_softargmax(
        [up, lazy], pa, bias, packed_row, bias, True, ppacked, pb, 30, partial=part), fl + 2.0 * 32 * 30 * S2, 32.0 * S2)
