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
torch.cuda.synchronize()
print('sanitize_small: ok', float(ops.tc_unpack(out).abs().sum()), int(idx.sum()), float(cw.sum()), float(ade.sum()))
