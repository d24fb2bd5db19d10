"""Micro-benchmark of the tensor-core conv kernel on the dominant Y-Net layer shapes (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion_style_transfer_b200 import ops  # noqa: E402

SHAPES = [  # (source channels, C_out, H, W)
    ((16, 32, 2), 32, 416, 416), ((64,), 32, 416, 416), ((16, 32, 16), 32, 416, 416), ((32,), 32, 416, 416), ((32,), 16, 416, 416), ((32, 32, 2), 32, 208, 208),
    ((64,), 32, 208, 208), ((32, 64, 2), 64, 104, 104), ((64,), 64, 104, 104), ((130,), 130, 13, 13),
]
N = int(os.environ.get('N', 64))
if os.environ.get('SHAPES'):          # e.g. SHAPES=2,3 for an ncu capture of single layers
    SHAPES = [SHAPES[int(i)] for i in os.environ['SHAPES'].split(',')]
torch.manual_seed(0)
print(f'env: CTAS_PER_SM={os.environ.get("YNET_TC_CTAS_PER_SM")} STAGES={os.environ.get("YNET_TC_STAGES")} '
      f'J={os.environ.get("YNET_TC_J")} N={N}')
_w = torch.randn(4096, 4096, device='cuda')
for _ in range(20):
    _w @ _w          # clock warm-up
torch.cuda.synchronize()
for cins, cout, H, W in SHAPES:
    srcs = [ops.tc_pack(torch.randn(N, c, H, W, device='cuda')) for c in cins]
    w = torch.randn(cout, sum(cins), 3, 3, device='cuda') * 0.1
    packed = ops.tc_pack_weights(w, list(cins))
    bias = torch.zeros((cout + 15) // 16 * 16, device='cuda')
    for _ in range(10):
        ops.tc_conv3x3(srcs, packed, bias, cout, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        out = ops.tc_conv3x3(srcs, packed, bias, cout, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * 9 * sum(cins) * cout * H * W * N
    cin_pad = sum((c + 15) // 16 * 16 for c in cins)
    by = 2.0 * (cin_pad + (cout + 15) // 16 * 16) * H * W * N
    print(f'{str(cins):>14} -> {cout:3d} @{H:3d}  {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TF/s  {by / ms / 1e6:7.1f} GB/s  tune={[hex(v) for v in ops._tc_tune.values()]}')
    ops._tc_tune.clear()

UP_SHAPES = [((32,), 16, 208, 208), ((64,), 32, 104, 104), ((64,), 32, 52, 52), ((64,), 32, 26, 26), ((128,), 64, 13, 13)]
if os.environ.get('UP', '1') != '0':
    for cins, cout, h, w in UP_SHAPES:
        srcs = [ops.tc_pack(torch.randn(N, c, h, w, device='cuda')) for c in cins]
        wgt = (torch.randn(cout, sum(cins), 3, 3, device='cuda') * 0.1).contiguous()
        b = torch.zeros(cout, device='cuda')
        w_eff, b_eff = ops.tc_upconv_phase_weights(wgt, b)
        packed = ops.tc_pack_weights(w_eff, list(cins))
        bw = ops.tc_upconv_border_weights(wgt, list(cins))
        for _ in range(10):
            ops.tc_upconv3x3(srcs, packed, b_eff, bw, b, cout)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.tc_upconv3x3(srcs, packed, b_eff, bw, b, cout)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * 9 * sum(cins) * cout * 4 * h * w * N
        print(f'up {str(cins):>10} -> {cout:3d} @{h:3d}->{2 * h:3d}  {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TF/s  tune={ops._tc_tune}')
        ops._tc_tune.clear()

if os.environ.get('PRED', '1') != '0':
    for cin, cout, H, W in [(32, 30, 416, 416), (32, 12, 416, 416)]:
        a = ops.tc_pack(torch.relu(torch.randn(N, cin, H, W, device='cuda')))
        wgt = torch.randn(cout, cin, 1, 1, device='cuda') * 0.5
        packed = ops.tc_pack_weights(wgt, [cin])
        bias = torch.zeros((cout + 15) // 16 * 16, device='cuda')
        for _ in range(5):
            ops.tc_conv1x1_softargmax(a, packed, bias, cout)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.tc_conv1x1_softargmax(a, packed, bias, cout)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        by = 2.0 * a.C_pad * H * W * N
        print(f'pred+softargmax {cin} -> {cout} @{H}  {ms:8.3f} ms  {by / ms / 1e6:7.1f} GB/s')

if os.environ.get('FUSED', '1') != '0':
    cin = cmid = 32
    cout, H, W = 30, 416, 416
    a = ops.tc_pack(torch.relu(torch.randn(N, cin, H, W, device='cuda')))
    w = torch.randn(cmid, cin, 3, 3, device='cuda') * 0.1
    packed = ops.tc_pack_weights(w, [cin])
    bias = torch.zeros(32, device='cuda')
    ppacked = ops.tc_pack_weights(torch.randn(cout, cmid, 1, 1, device='cuda') * 0.5, [cmid])
    pbias = torch.zeros(32, device='cuda')
    for _ in range(5):
        ops.tc_conv3x3_pred_softargmax([a], packed, bias, cmid, True, ppacked, pbias, cout)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.tc_conv3x3_pred_softargmax([a], packed, bias, cmid, True, ppacked, pbias, cout)
    e1.record()
    torch.cuda.synchronize()
    print(f'fused conv {cin}->{cmid} + pred {cout} + softargmax @{H}  {e0.elapsed_time(e1) / 10:8.3f} ms')
