"""Throughput of the reference-named drop-in call itself: utils/evaluate.py::evaluate() (23-argument signature,
DataLoader of scenes, host-side reference RNG semantics, eager launches) on synthetic inD-long TTST+CWS scenes.

    python tools/bench_evaluate.py [--agents 128] [--scenes 3]

bench.py's `e2e` goes through the CUDA-graph replay of the same batch body (GraphedForecaster); this script shows what a
user of the unchanged reference scripts gets from evaluate() without opting into the graph.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import pandas as pd
import torch
from torch.utils.data import DataLoader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from motion_style_transfer_b200 import ops  # noqa: E402
from motion_style_transfer_b200.utils.dataloader import SceneDataset, scene_collate  # noqa: E402
from motion_style_transfer_b200.utils.evaluate import evaluate  # noqa: E402
from motion_style_transfer_b200.utils.image_utils import create_dist_mat  # noqa: E402
from motion_style_transfer_b200 import synthetic as O  # noqa: E402  (synthetic input generators)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=128)
    ap.add_argument('--scenes', type=int, default=3)
    args = ap.parse_args()
    cfg = bench.WORKLOADS['ind_long_ttst_cws']
    dev = torch.device('cuda')
    model = bench.build_model_state(cfg).to(dev).eval().set_backend('bf16')
    total = cfg['obs'] + cfg['pred']
    rows = []
    images = {}
    for s in range(args.scenes + 1):                       # scene 0 is the warm-up
        tr = O.synthetic_tracks(args.agents, total, bench.H, bench.W, seed=10 + s).numpy() / cfg['resize']
        for b in range(args.agents):
            for t in range(total):
                rows.append(dict(frame=t, trackId=b, x=float(tr[b, t, 0]), y=float(tr[b, t, 1]), sceneId=f's{s}',
                                 metaId=s * args.agents + b))
        images[f's{s}'] = O.synthetic_scene(bench.H, bench.W, seed=s)
    df = pd.DataFrame(rows)
    tmpl = torch.Tensor(create_dist_mat(size=int(4200 * cfg['resize'])))

    def run(frame):
        loader = DataLoader(SceneDataset(frame, resize=cfg['resize'], total_len=total), batch_size=1,
                            collate_fn=scene_collate)
        torch.manual_seed(1)
        np.random.seed(2)
        return evaluate(model, loader, images, dev, 'ind-dataset-v1.0', None, tmpl, cfg['wps'], 'test', cfg['n_goal'],
                        cfg['n_traj'], cfg['obs'], args.agents, cfg['resize'], cfg['T'], cfg['ttst'], cfg['cws'],
                        cfg['thr'], cfg['cwsp'])

    run(df[df.sceneId == 's0'])                            # autotune + allocator warm-up
    torch.cuda.synchronize()
    l0 = ops.launch_count
    t0 = time.perf_counter()
    ade, fde, out, _ = run(df[df.sceneId != 's0'])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = args.scenes * args.agents * cfg['n_goal'] * cfg['n_traj']
    print(json.dumps({'api': 'utils.evaluate.evaluate (eager, host RNG)', 'agent_trajectories_per_s': n / dt,
                      'ms_per_scene': 1000 * dt / args.scenes, 'agents_per_scene': args.agents, 'scenes': args.scenes,
                      'launches': ops.launch_count - l0, 'ade': float(ade), 'fde': float(fde), 'rows': len(out)}))


if __name__ == '__main__':
    main()
