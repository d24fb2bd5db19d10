"""compute-sanitizer target: one small call of every kernel that changed this round (memcheck / racecheck).

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion_style_transfer_b200 import ops  # noqa: E402

ops.tc_autotune_enabled = False
torch.manual_seed(0)
dev = 'cuda'
H, W = 64, 96
# quad rasteriser + conv over a quad source
tmpl = ops.create_dist_template(3 * W, dev)
coords = torch.stack([torch.rand(6) * (W - 1), torch.rand(6) * (H - 1)], 1).to(dev)
pyr = ops.tc_rasterize_pyramid(tmpl, coords, 3, 2, H, W, 4, quad_levels=2)
x = ops.tc_pack(torch.randn(3, 16, H, W, device=dev))
wgt = torch.randn(32, 18, 3, 3, device=dev) * 0.1
pk = ops.tc_pack_hoisted_weights(wgt, [('conv', (0, 16)), ('quad', (16, 2))])
y = ops.tc_conv3x3([x, pyr[0]], pk, torch.zeros(32, device=dev), 32, True, pad_out=True)
# upconv with the mma ring fix, NT = 2, 4, 8, odd sizes
for cin, cout, h, w in [(32, 16, 13, 21), (64, 32, 7, 40), (128, 64, 5, 3), (24, 8, 1, 1)]:
    src = ops.tc_pad_replicate(ops.tc_pack(torch.randn(2, cin, h, w, device=dev)))
    wu = torch.randn(cout, cin, 3, 3, device=dev) * 0.1
    bu = torch.randn(cout, device=dev)
    w_eff, b_eff = ops.tc_upconv_phase_weights(wu, bu)
    out = ops.tc_upconv3x3([src], ops.tc_pack_weights(w_eff, [cin]), b_eff, ops.tc_upconv_border_weights(wu, [cin]), bu, cout)
# sequential CDF sampler
p = torch.rand(3, 1, H, W, device=dev)
u = torch.rand(3, 1000, dtype=torch.float64, device=dev)
idx, xy = ops.multinomial_replacement(p, u, rel_threshold=0.01)
# CWS + ADE/FDE
sig = torch.rand(3, H, W, device=dev)
wp = torch.rand(20, 3, 2, device=dev) * 50
cw = ops.cws_waypoint(sig, wp, torch.rand(3, 2, device=dev) * 50, 0.5, torch.full((20,), 6.0, device=dev), 2.0, True)
ade, fde = ops.ade_fde(torch.rand(3, 30, 2, device=dev), torch.rand(20, 3, 30, 2, device=dev), torch.rand(20, 3, 2, 2, device=dev), 0.33)
# round 2: row-marching conv (plain, padded output, multi-source + hoisted partial), fused conv + predictor + soft-argmax tail,
# scene-image preprocessing
xr = ops.tc_pack(torch.relu(torch.randn(3, 32, 34, 150, device=dev)))
wr = torch.randn(32, 32, 3, 3, device=dev) * 0.1
b32 = torch.randn(32, device=dev) * 0.1
pkr = ops.tc_rowconv_pack_weights(wr, 32)
yr = ops.tc_rowconv3x3(xr, pkr, b32, 32, True, pad_out=True)
up = ops.tc_pack(torch.randn(6, 16, 34, 150, device=dev))
wpl = ops.tc_pack(torch.rand(6, 2, 34, 150, device=dev))
feat = ops.tc_pack(torch.relu(torch.randn(2, 32, 34, 150, device=dev)))
w2 = torch.randn(32, 50, 3, 3, device=dev) * 0.1
part = ops.tc_conv3x3_hilo([feat], ops.tc_pack_weights(w2[:, 16:48].contiguous(), [32]), 32, True).repeat_interleave(3)
ym = ops.tc_rowconv3x3([up, wpl], ops.tc_rowconv_pack_weights_cat(w2, [(0, 16, 16), (48, 50, 16)]), b32, 32, True, partial=part)
wp1 = torch.randn(30, 32, 1, 1, device=dev) * 0.5
sa = ops.tc_rowconv3x3_pred_softargmax(xr, pkr, b32, 32, True, ops.tc_pack_weights(wp1, [32]), torch.zeros(32, device=dev), 30)
from motion_style_transfer_b200.utils import image_utils as U  # noqa: E402
import numpy as np  # noqa: E402
img = np.random.RandomState(0).randint(0, 256, (97, 131, 3)).astype(np.uint8)
pre = U.preprocess_scene_image(img, 0.33, 32)
pre4 = U.preprocess_scene_image(img, 0.25, 32)
msk = U.preprocess_scene_image(img[:, :, 0] % 6, 0.33, 32, seg_mask=True)
# round 2, second half: waypoint planes from the template, two-conv block (plain + tail), split-bf16 engine ops and the
# tensor-core training convs, oriented preprocessing, evaluate() through the cached graph
H2, W2 = 64, 256
coords2 = torch.stack([torch.rand(6 * 2) * (W2 - 1), torch.rand(6 * 2) * (H2 - 1)], 1).to(dev).contiguous()
tmpl2 = ops.create_dist_template(3 * W2, dev)
lazy = ops.tc_rasterize_pyramid(tmpl2, coords2, 6, 2, H2, W2, 2, lazy_levels=2)
up2 = ops.tc_pack(torch.randn(6, 16, H2, W2, device=dev))
feat2 = ops.tc_pack(torch.relu(torch.randn(2, 32, H2, W2, device=dev)))
wa = torch.randn(32, 50, 3, 3, device=dev) * 0.1
part2 = ops.tc_conv3x3_hilo([feat2], ops.tc_pack_weights(wa[:, 16:48].contiguous(), [32]), 32, False).repeat_interleave(3)
pa = ops.tc_rowconv_pack_weights_cat(wa, [(0, 16, 16)] + lazy[0].weight_parts(48))
y1 = ops.tc_rowconv3x3([up2, lazy[0]], pa, b32, 32, True, partial=part2)
y2 = ops.tc_rowconv2_wp([up2, lazy[0]], pa, b32, pkr, b32, 32, True, pad_out=True, partial=part2)
y3 = ops.tc_rowconv2_wp_pred_softargmax([up2, lazy[0]], pa, b32, pkr, b32, True, ops.tc_pack_weights(wp1, [32]),
                                        torch.zeros(32, device=dev), 30, partial=part2)
up1 = ops.tc_pack(torch.randn(6, 16, H2 // 2, W2 // 2, device=dev))
y4 = ops.tc_rowconv3x3([up1, lazy[1]], pa, b32, 32, True)
xs = ops.split_pack(torch.randn(3, 21, 20, 36, device=dev))
ws = torch.randn(30, 21, 3, 3, device=dev) * 0.1
ys = ops.tc_conv3x3_split([xs], ops.split_pack_weights(ws, [xs.layout]), torch.zeros(32, device=dev), 30, True)
ys2 = ops.split_unpack(ops.split_upsample(ops.split_maxpool(ys)))
ymask = ops.split_pack_masked(torch.randn(3, 30, 20, 36, device=dev), ops.split_unpack(ys))
pre_o = U.preprocess_scene_image(img, 0.33, 32, orient=5)
# k-means with four points per thread in flight: N not a multiple of 4 x 512, K at and below the template bounds, 2-CTA clusters
for n_pts, k in ((10000, 19), (777, 5), (4100, 32)):
    pts = torch.randint(0, 400, (3, n_pts, 2), device=dev).float()
    init_idx = torch.stack([torch.randperm(n_pts, device=dev)[:k] for _ in range(3)]).int()
    cen, _, n_it, _ = ops.kmeans_batched(pts, init_idx, torch.randint(0, n_pts, (3, 64), dtype=torch.int32, device=dev), 0.001, 40,
                                         want_assign=True)
torch.cuda.synchronize()
print('sanitize_small r02b: ok', float(ops.tc_unpack(y1).abs().sum()), float(ops.tc_unpack(y2).abs().sum()), float(y3.sum()),
      float(ops.tc_unpack(y4).abs().sum()), float(ys2.abs().sum()), float(ops.split_unpack(ymask).abs().sum()), float(pre_o.sum()),
      float(cen.sum()), int(n_it.max()))
print('sanitize_small r02: ok', float(ops.tc_unpack(yr).abs().sum()), float(ops.tc_unpack(ym).abs().sum()), float(sa.sum()),
      float(pre.sum()), float(pre4.sum()), float(msk.sum()))
print('sanitize_small: ok', float(ops.tc_unpack(out).abs().sum()), int(idx.sum()), float(cw.sum()), float(ade.sum()))
