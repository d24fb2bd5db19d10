"""Kernel-time breakdown of one fine-tuning step (torch.profiler, CUDA activities): python tools/profile_finetune.py [backend]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from motion_style_transfer_b200 import parallel  # noqa: E402
from motion_style_transfer_b200.autograd_engine import BCEWithLogitsLoss  # noqa: E402
from motion_style_transfer_b200.models.trainer import FusedAdam, apply_freeze_policy  # noqa: E402
from motion_style_transfer_b200.utils.image_utils import create_dist_mat, create_gaussian_heatmap_template  # noqa: E402
from motion_style_transfer_b200.utils.train_epoch import train_epoch  # noqa: E402

backend = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
cfg = bench.WORKLOADS['sdd_short']
dev = torch.device('cuda', 0)
model = bench.build_model_state(cfg).to(dev)
model.set_backend(backend)
apply_freeze_policy(model, 'mosa_1', [0, 1, 2, 3, 4], 'original')
opt = FusedAdam(model.parameters(), lr=0.003)
loader, images = bench._scene_loader(cfg, 10, seed=20)
size = int(4200 * cfg['resize'])
tmpl = torch.Tensor(create_dist_mat(size=size)).to(dev)
gt = torch.Tensor(create_gaussian_heatmap_template(size=size, kernlen=31, nsig=4, normalize=False)).to(dev)


def epoch(e):
    return train_epoch(model, loader, images, opt, BCEWithLogitsLoss(), 1000, dev, 'sdd', None, gt, tmpl, cfg['wps'], e, cfg['obs'],
                       cfg['pred'], 10, 10000, cfg['resize'], None)


for e in range(3):
    epoch(e)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    epoch(3)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))

if os.environ.get('HOST_PROFILE', '0') == '1':
    # where the host time of a step goes (the bf16x3 step is issue-bound: ~24 ms of Python against ~18 ms of kernels)
    import cProfile
    import pstats
    import time
    t0 = time.perf_counter()
    for e in range(4, 7):
        epoch(e)
    t_issue = (time.perf_counter() - t0) / 3
    torch.cuda.synchronize()
    print(f'host issue time per epoch (1 step of 10 agents): {1000 * t_issue:.2f} ms')
    pr = cProfile.Profile()
    pr.enable()
    for e in range(7, 10):
        epoch(e)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats('tottime').print_stats(35)
    st.sort_stats('cumulative').print_stats(45)
