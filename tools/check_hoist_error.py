"""Trajectory soft-argmax error of the bf16 engine variants against the fp32 engine (same inputs, same weights).

    python tools/check_hoist_error.py        # prints max / mean |dx| in pixels for hoist off / hi / hi+lo
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from motion_style_transfer_b200 import ops  # noqa: E402
from motion_style_transfer_b200 import synthetic as O  # noqa: E402  (synthetic input generators)

cfg = bench.WORKLOADS['ind_long_ttst_cws']
H = W = 416
B, G = int(os.environ.get('B', 6)), 20
dev = torch.device('cuda')
model = bench.build_model_state(cfg).to(dev).eval()
scene = O.synthetic_scene(H, W, seed=0)[None].to(dev)
traj = O.synthetic_tracks(B, cfg['obs'] + cfg['pred'], H, W, seed=5).to(dev)
tmpl = ops.create_dist_template(int(4200 * cfg['resize']), dev)
g = torch.Generator().manual_seed(3)
wps = (torch.rand(G, B, 2, 2, generator=g) * 300 + 50).to(dev)
obs = ops.rasterize_patches(tmpl, traj[:, :cfg['obs']].reshape(-1, 2), H, W).view(B, cfg['obs'], H, W)


def run(backend, hoist=None, lo=None):
    model.set_backend(backend)
    if hoist is not None:
        model.engine.hoist = hoist
        model.engine.hoist_lo = bool(lo)
    with torch.no_grad():
        feats = model.pred_features(scene, obs)
        return model.engine.decode_trajectories(feats, wps, tmpl, H, W, 256).float().cpu()


ref = run('fp32')
for name, kw in [('bf16 direct', dict(hoist=False)), ('bf16 hoisted hi', dict(hoist=True, lo=False)),
                 ('bf16 hoisted hi+lo', dict(hoist=True, lo=True))]:
    d = (run('bf16', **kw) - ref).abs()
    print(f'{name:22s} max |d| {d.max().item():.4f} px   mean |d| {d.mean().item():.5f} px')
