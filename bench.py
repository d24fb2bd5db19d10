#!/usr/bin/env python
"""bench.py -- agent-trajectories/sec of the Y-Net+MoSA forecasting hot path (TTST + CWS) on B200.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A "step" = one pass of the evaluate() batch body (rasterise -> encoder -> goal decoder -> sigmoid ->
TTST 10k multinomial + k-means(19) -> CWS -> 20 trajectory-decoder passes -> soft-argmax -> ADE/FDE)
over one batch of `--agents` synthetic agents per GPU on a 416x416 synthetic semantic map.  Agents are
independent, so ranks shard them with no data-path collective (weak scaling).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2] shape: inD long-term eval with TTST + CWS (config/inD_longterm_eval.yaml)
    'ind_long_ttst_cws': dict(obs=5, pred=30, wps=[14, 29], resize=0.33, T=1.8, thr=0.002, ttst=True, cws=True,
                              cwsp=dict(sigma_factor=6, ratio=2, rot=True), n_goal=20, n_traj=1),
    # BASELINE.json configs[0] shape: SDD short-term eval (config/sdd_shortterm_eval.yaml)
    'sdd_short': dict(obs=8, pred=12, wps=[11], resize=0.25, T=1.0, thr=0.01, ttst=False, cws=False, cwsp=None,
                      n_goal=20, n_traj=1),
    # BASELINE.json configs[3] shape: Y-Net-Mod (network=fusion, n_fusion=2) on config/inD_shortterm_eval.yaml
    # (scripts/inD/scene1_car_to_truck/ynetmod/*.sh:4-11), MoSA on the scene / motion / fusion branches
    'ind_short_ynetmod': dict(obs=8, pred=12, wps=[11], resize=0.33, T=1.0, thr=0.002, ttst=False, cws=False, cwsp=None,
                              n_goal=20, n_traj=1, network='fusion', n_fusion=2, position=['scene', 'motion', 'fusion']),
}
ENC, DEC = [32, 32, 64, 64, 64], [64, 64, 64, 32, 32]
H = W = 416
GF_PER_AGENT = {'ind_long_ttst_cws': 371.6, 'sdd_short': 364.7,     # reference-executed GFLOP
                'ind_short_ynetmod': 364.7 - 4.68 + 2.44}             # Y-Net-Mod encoder: 2.44 instead of 4.68 GF


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], bf16_tflops=d['bf16_tflops'], src='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, src='fallback')


def load_traffic(kernel, agents):
    """Average DRAM bytes per launch of `kernel` from the committed ncu pass of this command
    (profiles/ncu_traffic_r02.json, tools/ncu_traffic.py); None when the capture was made at another batch size."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic_r02.json')
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    k = d.get('kernels', {}).get(kernel)
    if d.get('agents_per_step') != agents or not k or not k.get('launches'):
        return None
    return k['dram_bytes'] / k['launches']


def build_model_state(cfg, seed=0):
    """Random-init Y-Net + MoSA r=1 on encoder stages 0-4 (reference default init, LoRA B ~ N(0, 0.02),
    predictors x50 so that the heat maps are peaky: SURVEY 8d)."""
    from motion_style_transfer_b200.models.ynet import YNet
    torch.manual_seed(seed)
    m = YNet(obs_len=cfg['obs'], pred_len=cfg['pred'], segmentation_model_fp=None, encoder_channels=ENC,
             decoder_channels=DEC, n_waypoints=len(cfg['wps']), train_net='mosa_1',
             position=cfg.get('position', [0, 1, 2, 3, 4]), network=cfg.get('network', 'original'),
             n_fusion=cfg.get('n_fusion'))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if 'lora_B' in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        m.goal_decoder.predictor.weight.mul_(50.0)
        m.traj_decoder.predictor.weight.mul_(50.0)
    return m


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc = [], None
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)        # (its exit must not overlap the end-to-end loop that follows)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except (ValueError, IndexError):
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm), sm_min_mhz=float(min(sm)) if sm else None)


def _reference_evaluate(model_state, cfg, B, device, seeds, ns):
    """One call of the UNMODIFIED reference's utils/evaluate.py::evaluate (evaluate.py:37-315) on one synthetic scene of B
    agents through its own SceneDataset / DataLoader; returns (seconds, ade, fde).  `ns` = oracle.ref_harness.load()."""
    import pandas as pd
    import warnings
    from torch.utils.data import DataLoader
    from motion_style_transfer_b200 import synthetic as S
    m = ns.ynet.YNet(obs_len=cfg['obs'], pred_len=cfg['pred'], segmentation_model_fp=None, encoder_channels=ENC,
                     decoder_channels=DEC, n_waypoints=len(cfg['wps']), train_net='mosa_1',
                     position=cfg.get('position', [0, 1, 2, 3, 4]), network=cfg.get('network', 'original'),
                     n_fusion=cfg.get('n_fusion'))
    m.load_state_dict(model_state, strict=True)
    m = m.to(device).eval()
    total = cfg['obs'] + cfg['pred']
    tracks = S.synthetic_tracks(B, total, H, W, seed=seeds[0]) / cfg['resize']
    rows = [dict(frame=t, trackId=b, x=float(tracks[b, t, 0]), y=float(tracks[b, t, 1]), sceneId='s0', metaId=b)
            for b in range(B) for t in range(total)]
    ds = ns.dataloader.SceneDataset(pd.DataFrame(rows), resize=cfg['resize'], total_len=total)
    dl = DataLoader(ds, batch_size=1, collate_fn=ns.dataloader.scene_collate)
    tmpl = torch.Tensor(ns.image_utils.create_dist_mat(int(4200 * cfg['resize']))).to(device)      # trainer.py:326
    images = {'s0': S.synthetic_scene(H, W, seed=0)}
    torch.manual_seed(seeds[1])
    np.random.seed(seeds[2])
    if torch.device(device).type == 'cuda':
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ade, fde, _, _ = ns.evaluate.evaluate(m, dl, images, device, 'sdd', None, tmpl, cfg['wps'], 'test', cfg['n_goal'],
                                              cfg['n_traj'], cfg['obs'], B, cfg['resize'], cfg['T'], cfg['ttst'], cfg['cws'],
                                              cfg['thr'], cfg['cwsp'], network=cfg.get('network', 'original'))
    if torch.device(device).type == 'cuda':
        torch.cuda.synchronize()
    return time.perf_counter() - t0, float(ade), float(fde)


def _torch_network_only(model_state, cfg, B, device, ns):
    """The reference's own network calls of one evaluate() batch -- pred_features, pred_goal, n_goal x (pred_traj +
    softargmax) (evaluate.py:123-126, 248-266) -- on `device`, inputs resident, under fp32 (TF32 convs) and bf16 autocast:
    the cuDNN number the hand-written engine has to beat on the same silicon."""
    m = ns.ynet.YNet(obs_len=cfg['obs'], pred_len=cfg['pred'], segmentation_model_fp=None, encoder_channels=ENC,
                     decoder_channels=DEC, n_waypoints=len(cfg['wps']), train_net='mosa_1',
                     position=cfg.get('position', [0, 1, 2, 3, 4]), network=cfg.get('network', 'original'),
                     n_fusion=cfg.get('n_fusion'))
    m.load_state_dict(model_state, strict=True)
    m = m.to(device).eval()
    g = torch.Generator(device='cpu').manual_seed(0)
    sem = torch.softmax(torch.randn(1, 6, H, W, generator=g), 1).to(device).expand(B, -1, -1, -1)
    obs = torch.rand(B, cfg['obs'], H, W, generator=g).to(device)
    wp = torch.rand(B, len(cfg['wps']), H, W, generator=g).to(device)
    pools = [torch.nn.AvgPool2d(2 ** i, 2 ** i) for i in range(1, 6)]
    out = {}

    def body():
        feats = m.pred_features(sem, obs)
        m.pred_goal(feats)
        for _ in range(cfg['n_goal'] * cfg['n_traj']):
            pyr = [wp] + [p(wp) for p in pools]
            m.softargmax(m.pred_traj([torch.cat([f, w], dim=1) for f, w in zip(feats, pyr)]))

    for name, ctx in (('tf32', torch.autocast('cuda', enabled=False)), ('bf16_autocast', torch.autocast('cuda', dtype=torch.bfloat16))):
        with torch.no_grad(), ctx:
            body()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                body()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        out[name] = dict(ms_per_batch=ms, agent_trajectories_per_s=B * cfg['n_goal'] * cfg['n_traj'] / (ms / 1000))
    out['what'] = f'{B} agents, inputs resident, CUDA events; no sampling / k-means / CWS / rasterisation'
    return out


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host threads.  When the unmodified
    reference is importable (oracle/_ref staged by oracle/build_ref.py, or /root/reference) this is its
    utils/evaluate.py::evaluate itself (kind "reference"); otherwise the oracle port (kind "port")."""
    if rank != 0:
        return
    from oracle import ref_harness
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = build_model_state(cfg)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    B = args.ref_agents
    kind = 'reference' if ref_harness.available() else 'port'
    times = []
    if kind == 'reference':
        ns = ref_harness.load()
        for it in range(args.warmup + args.steps):
            dt, _, _ = _reference_evaluate(sd, cfg, B, 'cpu', (100 + it, 1000 + it, 2000 + it), ns)
            if it >= args.warmup:
                times.append(dt)
    else:
        from oracle import ynet_oracle as O
        scene = O.synthetic_scene(H, W, seed=0)[None]
        tmpl = O.create_dist_mat(int(4200 * cfg['resize'])).astype(np.float32)
        for it in range(args.warmup + args.steps):
            traj = O.synthetic_tracks(B, cfg['obs'] + cfg['pred'], H, W, seed=100 + it)
            torch.manual_seed(1000 + it)
            np.random.seed(2000 + it)
            t0 = time.perf_counter()
            O.evaluate_batch(sd, scene, traj, tmpl, cfg['wps'], cfg['n_goal'], cfg['n_traj'], cfg['obs'], cfg['resize'],
                             cfg['T'], cfg['ttst'], cfg['cws'], cfg['thr'], cfg['cwsp'])
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
    ms = 1000 * sum(times) / len(times)
    val = B * cfg['n_goal'] * cfg['n_traj'] / (ms / 1000)
    what = ('unmodified reference utils/evaluate.py::evaluate on device cpu' if kind == 'reference'
            else 'oracle port (reference not staged)')
    sample = (f'{what}: {B} agents x {cfg["n_goal"] * cfg["n_traj"]} trajectories per step (one scene, one batch), '
              f'{args.steps} steps, 416x416, torch CPU fp32 with {threads} threads; the GPU arm runs the same workload with '
              f'{args.agents} agents per step -- a CPU step of that size would take ~{ms / 1000 * args.agents / B / 60:.0f} min')
    print(json.dumps({
        'impl': 'reference', 'metric': 'agent-trajectories/sec', 'value': val, 'unit': 'agent-trajectories/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args, cfg), 'agents_per_step': B,
                   'agents_per_step_note': 'bounded sample of the GPU arm\'s workload (same scene size, model, TTST/CWS '
                                           'settings); throughput per agent-trajectory does not depend on the batch size '
                                           'on the CPU (per-agent Python loops)'},
        'cpu_baseline': {'value': val, 'unit': 'agent-trajectories/s', 'cores': threads, 'kind': kind,
                         'sample': sample},
        'e2e': {'value': val, 'unit': 'agent-trajectories/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def _scene_loader(cfg, n_agents, seed=10):
    """One synthetic scene of n_agents agents as the drop-in SceneDataset / DataLoader pair evaluate() and train_epoch()
    take (dataloader.py:8-50)."""
    import pandas as pd
    from torch.utils.data import DataLoader
    from motion_style_transfer_b200 import synthetic as S
    from motion_style_transfer_b200.utils.dataloader import SceneDataset, scene_collate
    total = cfg['obs'] + cfg['pred']
    tr = S.synthetic_tracks(n_agents, total, H, W, seed=seed).numpy() / cfg['resize']
    df = pd.DataFrame({'frame': np.tile(np.arange(total), n_agents), 'trackId': np.repeat(np.arange(n_agents), total),
                       'x': tr[:, :, 0].reshape(-1), 'y': tr[:, :, 1].reshape(-1), 'sceneId': 's0',
                       'metaId': np.repeat(np.arange(n_agents), total)})
    ds = SceneDataset(df, resize=cfg['resize'], total_len=total)
    return DataLoader(ds, batch_size=1, collate_fn=scene_collate), {'s0': S.synthetic_scene(H, W, seed=0)}


def _dist_max(ms, world, dev):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


def run_evaluate_mode(args, cfg, rank, world, dev):
    """--mode evaluate: ONE scene of --agents agents through the reference-named drop-in utils/evaluate.py::evaluate()
    (23-argument signature, DataLoader, host RNG semantics).  Under torchrun the agents of the scene are sharded over the
    ranks (parallel.shard_bounds) and the per-agent rows are gathered at the end of the scene (parallel.gather_rows, NCCL)
    INSIDE the timed region: strong scaling of one fixed job."""
    import torch.distributed as dist
    from motion_style_transfer_b200.utils.evaluate import evaluate
    from motion_style_transfer_b200.utils.image_utils import create_dist_mat
    model = build_model_state(cfg).to(dev).eval().set_backend(args.backend)
    loader, images = _scene_loader(cfg, args.agents)
    tmpl = torch.Tensor(create_dist_mat(size=int(4200 * cfg['resize'])))
    bs = args.chunk_agents if args.chunk_agents > 0 else 128

    def call():
        return evaluate(model, loader, images, dev, 'sdd', None, tmpl, cfg['wps'], 'test', cfg['n_goal'], cfg['n_traj'],
                        cfg['obs'], bs, cfg['resize'], cfg['T'], cfg['ttst'], cfg['cws'], cfg['thr'], cfg['cwsp'])

    for _ in range(max(2, args.warmup)):      # (the drop-in captures its CUDA graph on the second call that meets a shape)
        torch.manual_seed(1)
        np.random.seed(2)
        call()
    times, res = [], None
    for it in range(args.steps):
        torch.manual_seed(100 + it)
        np.random.seed(200 + it)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = call()
        e1.record()
        torch.cuda.synchronize()
        times.append(_dist_max(e0.elapsed_time(e1), world, dev))
    ms = sum(times) / len(times)
    if rank == 0:
        print(json.dumps({
            'metric': 'agent-trajectories/sec (sharded evaluate() drop-in, one scene, rows gathered)', 'mode': 'evaluate',
            'value': args.agents * cfg['n_goal'] * cfg['n_traj'] / (ms / 1000), 'unit': 'agent-trajectories/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': model.engine.backend_dtype(), 'data': 'synthetic',
            'config': {'workload': workload_name(args, cfg), 'agents_per_scene': args.agents, 'batch_size': bs,
                       'agents_per_rank': -(-args.agents // world), 'collective': 'all_gather of (ade, fde) rows per scene'},
            'ade': float(res[0]), 'fde': float(res[1]), 'step_ms': times}))
    if world > 1:
        dist.destroy_process_group()


def run_finetune_mode(args, cfg, rank, world, dev):
    """--mode finetune: optimiser steps of the drop-in utils/train_epoch.py::train_epoch (BASELINE configs 2 / 4):
    rasterise -> encoder -> goal + trajectory decoders -> 2 x BCEWithLogits x 1000 -> backward -> FusedAdam, batches of
    --chunk-agents (default 10, scripts/*/tune_mosa*.sh) agents; under torchrun the agents of every batch are sharded over
    the ranks and the flattened adapter gradient is summed with ONE NCCL all-reduce per step (parallel.allreduce_flat)."""
    import torch.distributed as dist
    from motion_style_transfer_b200 import parallel
    from motion_style_transfer_b200.autograd_engine import BCEWithLogitsLoss
    from motion_style_transfer_b200.models.trainer import FusedAdam, apply_freeze_policy
    from motion_style_transfer_b200.utils.image_utils import create_dist_mat, create_gaussian_heatmap_template
    from motion_style_transfer_b200.utils.train_epoch import train_epoch
    model = build_model_state(cfg).to(dev)
    if args.backend == 'bf16x3':          # forward + data-gradient convs on the tensor cores (split-bf16); default: fp32 kernels
        model.set_backend('bf16x3')
    net = cfg.get('network', 'original')
    pos = cfg.get('position', [0, 1, 2, 3, 4])
    apply_freeze_policy(model, 'mosa_1', pos, net)
    trainable = [q for q in model.parameters() if q.requires_grad]
    parallel.broadcast_params(trainable)
    opt = FusedAdam(model.parameters(), lr=0.003)
    crit = BCEWithLogitsLoss()
    bs = args.chunk_agents if args.chunk_agents > 0 else 10
    n_batches = max(1, args.agents // bs)
    loader, images = _scene_loader(cfg, n_batches * bs, seed=20)
    size = int(4200 * cfg['resize'])
    tmpl = torch.Tensor(create_dist_mat(size=size)).to(dev)
    gt_tmpl = torch.Tensor(create_gaussian_heatmap_template(size=size, kernlen=31, nsig=4, normalize=False)).to(dev)

    def epoch(e):
        return train_epoch(model, loader, images, opt, crit, 1000, dev, 'sdd', None, gt_tmpl, tmpl, cfg['wps'], e, cfg['obs'],
                           cfg['pred'], bs, 10000, cfg['resize'], net if net != 'original' else None)

    for e in range(max(1, args.warmup)):
        epoch(e)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    t_wall0 = time.time()
    times, out = [], None
    for it in range(args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = epoch(100 + it)
        e1.record()
        torch.cuda.synchronize()
        times.append(_dist_max(e0.elapsed_time(e1), world, dev) / n_batches)
    ms = sum(times) / len(times)
    clocks = sampler.stop(t_wall0, time.time()) if sampler else None
    mem = torch.cuda.memory_stats(dev)
    if rank == 0:
        print(json.dumps({
            'metric': 'fine-tune optimiser steps/sec (train_epoch drop-in, MoSA r=1 adapters)', 'mode': 'finetune',
            'clocks': clocks, 'mem': {k: mem.get(k, 0) for k in ('num_alloc_retries', 'num_device_alloc', 'num_device_free')},
            'value': 1000.0 / ms, 'unit': 'steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'bf16x3 (forward + dgrad), f32 (wgrad, Adam)' if args.backend == 'bf16x3' else 'f32',
            'data': 'synthetic',
            'config': {'workload': workload_name(args, cfg).replace(' eval,', ' fine-tune,'), 'agents_per_step': bs,
                       'steps_per_epoch': n_batches, 'trainable_floats': int(sum(q.numel() for q in trainable)),
                       'collective': 'one all-reduce of the flattened adapter gradient per step'},
            'train_ade': out[0], 'train_fde': out[1], 'train_loss': out[2], 'step_ms': times}))
    if world > 1:
        dist.destroy_process_group()


def workload_name(args, cfg):
    net = 'Y-Net-Mod+MoSA(mosa_1, scene/motion/fusion branches)' if cfg.get('network') == 'fusion' else \
        'Y-Net+MoSA(mosa_1, encoder stages 0-4)'
    return (f'{net} {args.workload} eval, obs {cfg["obs"]}/pred {cfg["pred"]}, '
            f'{cfg["n_goal"]} goals, TTST={cfg["ttst"]} (10k samples + k-means 19), CWS={cfg["cws"]}, '
            f'synthetic {H}x{W} semantic map, random-init weights')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='ind_long_ttst_cws', choices=sorted(WORKLOADS))
    ap.add_argument('--agents', type=int, default=128, help='agents per GPU per step')
    ap.add_argument('--chunk-agents', type=int, default=0,
                    help='agents per captured graph replay (0 = all of --agents in one replay); the throughput sweep runs '
                         '--agents 1024 ... 65536 in chunks of 128')
    ap.add_argument('--mode', default='forecast', choices=['forecast', 'evaluate', 'finetune'],
                    help='forecast = the evaluate() batch body (driver default); evaluate = one scene of --agents agents '
                         'through the sharded drop-in evaluate() (strong scaling); finetune = train_epoch optimiser steps')
    ap.add_argument('--ref-agents', type=int, default=2, help='agents per step of the CPU reference arm')
    ap.add_argument('--cpu-agents', type=int, default=4, help='agents of the bounded cpu_baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--torch-cuda-agents', type=int, default=32,
                    help='agents of the secondary bar: the unmodified reference on device cuda (0 = skip)')
    ap.add_argument('--no-roofline', action='store_true', help='skip the per-launch eager timing pass')
    ap.add_argument('--profile-layers', default=None, help='write a per-layer timing table to this path')
    ap.add_argument('--no-graph', dest='graph', action='store_false',
                    help='issue every kernel from Python instead of replaying the captured CUDA graph')
    ap.add_argument('--backend', default='bf16', choices=['bf16', 'fp32', 'bf16x3'],
                    help='bf16 = tcgen05 tensor-core engine (default), fp32 = CUDA-core reference-grade engine, bf16x3 = '
                         'split-bf16 tensor-core engine at the fp32 engine\'s parity (three MMAs per product)')
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        return run_reference(args, cfg, rank, world)

    import torch.distributed as dist
    from motion_style_transfer_b200 import ops, _lib
    from motion_style_transfer_b200.utils.evaluate import forecast_batch
    from motion_style_transfer_b200.utils.image_utils import DeviceRng
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()
    if args.mode == 'evaluate':
        return run_evaluate_mode(args, cfg, rank, world, dev)
    if args.mode == 'finetune':
        return run_finetune_mode(args, cfg, rank, world, dev)

    model = build_model_state(cfg).to(dev).eval().set_backend(args.backend)
    from motion_style_transfer_b200 import synthetic as O      # synthetic-input generators (the product never imports oracle)
    scene_host = O.synthetic_scene(H, W, seed=0)[None].contiguous().pin_memory()
    tmpl = ops.create_dist_template(int(4200 * cfg['resize']), dev)
    B = args.agents
    n_iter = args.warmup + args.steps
    total_len = cfg['obs'] + cfg['pred']
    traj_host = [O.synthetic_tracks(B, total_len, H, W, seed=1 + rank * 1000 + it).pin_memory() for it in range(n_iter)]
    traj_dev = [t.to(dev) for t in traj_host]
    scene_dev = scene_host.to(dev)
    rng = DeviceRng(seed=1234 + rank)
    launches_per_step = None
    C = args.chunk_agents if args.chunk_agents > 0 else B
    assert B % C == 0, '--agents must be a multiple of --chunk-agents'
    n_chunks = B // C
    ade_all = torch.empty(B, dtype=torch.float32, device=dev)
    fde_all = torch.empty(B, dtype=torch.float32, device=dev)

    def chunked(fn):
        """One step = all B agents, C at a time through `fn` (whose outputs are overwritten by the next replay)."""
        if n_chunks == 1:
            return fn
        def run(scene, traj):
            for c in range(n_chunks):
                r = fn(scene, traj[c * C:(c + 1) * C])
                ade_all[c * C:(c + 1) * C].copy_(r['ade'])
                fde_all[c * C:(c + 1) * C].copy_(r['fde'])
            return dict(ade=ade_all, fde=fde_all)
        return run

    if args.graph:
        # the ~450 launches of one batch are captured once and replayed (utils/evaluate.py::GraphedForecaster)
        from motion_style_transfer_b200.utils.evaluate import GraphedForecaster
        gf = GraphedForecaster(model, tmpl, tuple(scene_dev.shape), (C,) + tuple(traj_dev[0].shape[1:]), cfg['wps'], cfg['n_goal'],
                               cfg['n_traj'], cfg['obs'], cfg['resize'], cfg['T'], cfg['ttst'], cfg['cws'], cfg['thr'],
                               cfg['cwsp'], seed=1234 + rank)
        l0 = ops.launch_count
        gf.capture(scene_dev, traj_dev[0][:C])
        launches_per_step = (ops.launch_count - l0) // 3 * n_chunks      # 2 eager warm-ups + the captured pass, per chunk

        step = chunked(lambda scene, traj: gf(scene, traj))
    else:
        step = chunked(lambda scene, traj: forecast_batch(
            model, scene, traj, tmpl, cfg['wps'], cfg['n_goal'], cfg['n_traj'], cfg['obs'], cfg['resize'], cfg['T'],
            cfg['ttst'], cfg['cws'], cfg['thr'], cfg['cwsp'], rng=rng))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident throughput (`value`): inputs already in HBM ------------------------------------
    for it in range(args.warmup):
        step(scene_dev, traj_dev[it])
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ops.launch_count
    t_wall0 = time.time()
    from motion_style_transfer_b200.utils import kmeans as km_mod
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks, km_iters, host_ms = [], [], []
    mem0 = torch.cuda.memory_stats(dev)
    torch.cuda.cudart().cudaProfilerStart()       # `ncu --profile-from-start off` sees the timed region only
    e0.record()
    for it in range(args.warmup, n_iter):
        th = time.perf_counter()
        res = step(scene_dev, traj_dev[it])
        host_ms.append(1000 * (time.perf_counter() - th))
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append(ev)
        km_iters.append(km_mod.last_iters)
    e1.record()
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    t_wall1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    step_list = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    km_diag = [int(k.max().item()) for k in km_iters if k is not None]
    mem1 = torch.cuda.memory_stats(dev)
    mem_diag = {k: mem1.get(k, 0) - mem0.get(k, 0) for k in ('num_alloc_retries', 'num_device_alloc', 'num_device_free')}
    mem_diag['reserved_gb'] = mem1.get('reserved_bytes.all.peak', 0) / 2 ** 30
    mem_diag['allocated_peak_gb'] = mem1.get('allocated_bytes.all.peak', 0) / 2 ** 30
    launches = (ops.launch_count - launches0) if launches_per_step is None else launches_per_step * args.steps
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_step = ms_total / args.steps
    traj_per_step = world * B * cfg['n_goal'] * cfg['n_traj']
    value = traj_per_step / (ms_step / 1000)

    # ---- end to end through the public call with HOST buffers (H2D of inputs + D2H of metrics inside) -----
    # (every step ends in a blocking D2H read, so a host-side pause lands in the measurement: no garbage collection inside
    # the timed loop, and the clock sampler's nvidia-smi child has exited before it starts)
    import gc
    gc.collect()
    gc.disable()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_marks, e2e_km = [], []
    e2.record()
    for it in range(args.warmup, n_iter):
        sc = scene_host.to(dev, non_blocking=True)
        tr = traj_host[it].to(dev, non_blocking=True)
        r = step(sc, tr)
        ade_h, fde_h = r['ade'].cpu(), r['fde'].cpu()
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        e2e_marks.append(ev)
        e2e_km.append(int(km_mod.last_iters.max().item()) if km_mod.last_iters is not None else None)
    e3.record()
    barrier()
    gc.enable()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3)) / args.steps
    e2e_value = traj_per_step / (ms_e2e / 1000)
    h2d = scene_host.numel() * 4 + traj_host[0].numel() * 4
    d2h = 2 * B * 4

    # ---- roofline of the dominant kernel: per-launch CUDA-event timing of one extra step -----------------
    roofline, layer_table, kernel_table = None, None, []
    if rank == 0 and not args.no_roofline:
        peaks = load_peaks()
        _eager = lambda: forecast_batch(model, scene_dev, traj_dev[-1][:C], tmpl, cfg['wps'], cfg['n_goal'], cfg['n_traj'],
                                        cfg['obs'], cfg['resize'], cfg['T'], cfg['ttst'], cfg['cws'], cfg['thr'],
                                        cfg['cwsp'], rng=rng)
        _eager()                  # re-warm the eager allocator pool (graph capture emptied the cache)
        # eager passes (not the graph): per-launch CUDA events.  An event pair also spans any host stall between the two
        # records (allocator growth, lazy module load) while the stream is idle, so the pass runs three times and every
        # launch keeps its minimum.
        prof = None
        for _ in range(3):
            ops.profile_begin()
            _eager()
            cur = ops.profile_end()
            if prof is None or len(cur) != len(prof):
                prof = cur
            else:
                for a, b in zip(prof, cur):
                    a['ms'] = min(a['ms'], b['ms'])
        layer_table = prof
        if prof:
            by_kernel = {}
            for p in prof:
                k = by_kernel.setdefault(p['kernel'], dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
                k['ms'] += p['ms']
                k['flops'] += p['flops']
                k['bytes'] += p['bytes']
                k['n'] += 1
            top_name, top = max(by_kernel.items(), key=lambda kv: kv[1]['ms'])
            step_ms = sum(p['ms'] for p in prof)
            # which roof bounds the dominant kernel: its arithmetic intensity (executed flops / algorithmic bytes) against the
            # ridge of the MEASURED peaks.  Below the ridge the kernel cannot exceed AI x HBM bandwidth: it is judged against
            # the HBM copy bandwidth; above it against the bf16 tensor peak.  Both fractions are reported.
            ridge = peaks['bf16_tflops'] * 1e12 / (peaks['hbm_gbs'] * 1e9)
            ai = top['flops'] / top['bytes'] if top['bytes'] > 0 else float('inf')
            tf = top['flops'] / (top['ms'] / 1000) / 1e12
            gbs = top['bytes'] / (top['ms'] / 1000) / 1e9
            common = dict(kernel=top_name, traffic=load_traffic(top_name, C), algorithmic_bytes_per_launch=top['bytes'] / top['n'],
                          flops_per_launch=top['flops'] / top['n'], arithmetic_intensity=ai, ridge_flop_per_byte=ridge,
                          tensor_tflops=tf, tensor_frac=tf / peaks['bf16_tflops'], hbm_gbs=gbs,
                          hbm_frac=gbs / peaks['hbm_gbs'], launches=top['n'], avg_launch_ms=top['ms'] / top['n'],
                          share_of_step=top['ms'] / step_ms)
            if top['flops'] > 0 and ai >= ridge:
                roofline = dict(bound='tensor', achieved=tf, peak=peaks['bf16_tflops'], unit='TFLOP/s',
                                frac=tf / peaks['bf16_tflops'], peak_source=peaks['src'] + ' (burst bf16 cuBLAS)', **common)
            else:
                roofline = dict(bound='hbm', achieved=gbs, peak=peaks['hbm_gbs'], unit='GB/s', frac=gbs / peaks['hbm_gbs'],
                                peak_source=peaks['src'] + ' (copy bandwidth)', **common)
        if prof:
            # every kernel class of the step against its own roofline: convs vs the tensor peak, the
            # rasterise / soft-argmax / sampling / CWS kernels vs the measured HBM copy bandwidth (algorithmic bytes)
            for name, k in sorted(by_kernel.items(), key=lambda kv: -kv[1]['ms'])[:10]:
                row = dict(kernel=name, launches=k['n'], ms=k['ms'], share_of_step=k['ms'] / step_ms)
                if k['flops'] > 0:
                    row.update(tflops=k['flops'] / k['ms'] / 1e9, tensor_frac=k['flops'] / k['ms'] / 1e9 / peaks['bf16_tflops'])
                if k['bytes'] > 0:
                    row.update(hbm_gbs=k['bytes'] / k['ms'] / 1e6, hbm_frac=k['bytes'] / k['ms'] / 1e6 / peaks['hbm_gbs'])
                kernel_table.append(row)
        if args.profile_layers:
            with open(args.profile_layers, 'w') as f:
                json.dump(dict(step_ms=sum(p['ms'] for p in prof), launches=prof), f, indent=1)

    # ---- bounded CPU baseline on this host, rank 0 at N=1 only: the unmodified reference's evaluate() when staged
    # (oracle/_ref, kind "reference"), else the oracle port.  The only place the b200 arm touches oracle/ (as the checker
    # being TIMED next to the product, never as part of it) --------------------------------------------------------------
    cpu_baseline, torch_cuda = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref_harness
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        Bc = args.cpu_agents
        if ref_harness.available():
            ns = ref_harness.load()
            dt, _, _ = _reference_evaluate(sd, cfg, Bc, 'cpu', (77, 1, 2), ns)
            kind, what = 'reference', 'unmodified reference utils/evaluate.py::evaluate, device cpu'
        else:
            from oracle import ynet_oracle as ORC
            traj = ORC.synthetic_tracks(Bc, total_len, H, W, seed=77)
            tm = ORC.create_dist_mat(int(4200 * cfg['resize'])).astype(np.float32)
            torch.manual_seed(1)
            np.random.seed(2)
            t0 = time.perf_counter()
            ORC.evaluate_batch(sd, scene_host, traj, tm, cfg['wps'], cfg['n_goal'], cfg['n_traj'], cfg['obs'],
                               cfg['resize'], cfg['T'], cfg['ttst'], cfg['cws'], cfg['thr'], cfg['cwsp'])
            dt = time.perf_counter() - t0
            kind, what = 'port', 'oracle port'
        cpu_baseline = dict(value=Bc * cfg['n_goal'] * cfg['n_traj'] / dt, unit='agent-trajectories/s', cores=threads,
                            kind=kind, sample=f'{what}: {Bc} agents x {cfg["n_goal"]} trajectories, one batch, '
                                              f'{dt:.1f} s, torch CPU fp32 with {threads} threads')
        # ---- the secondary bar: the SAME unmodified reference on device='cuda', i.e. stock
        # PyTorch / cuDNN on this very B200 (TF32 convs = torch's default; host-side per-agent k-means as the reference
        # has it), and its network alone (encoder + goal decoder + 20 trajectory-decoder passes) under bf16 autocast
        if ref_harness.available() and args.torch_cuda_agents > 0:
            try:
                Bt = args.torch_cuda_agents
                _reference_evaluate(sd, cfg, min(Bt, 4), dev, (78, 3, 4), ns)              # warm-up (cuDNN autotune)
                dt, _, _ = _reference_evaluate(sd, cfg, Bt, dev, (79, 5, 6), ns)
                torch_cuda = dict(value=Bt * cfg['n_goal'] * cfg['n_traj'] / dt, unit='agent-trajectories/s',
                                  what='unmodified reference evaluate() on device cuda (PyTorch eager + cuDNN, TF32 convs, '
                                       f'host k-means as in kmeans.py:146-148), {Bt} agents in one batch, {dt:.2f} s')
                torch_cuda['network_only'] = _torch_network_only(sd, cfg, Bt, dev, ns)
            except Exception as e:      # the secondary bar must never break the bench line
                torch_cuda = dict(error=repr(e)[:300])

    if rank == 0:
        out = {
            'metric': 'agent-trajectories/sec', 'value': value, 'unit': 'agent-trajectories/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': model.engine.backend_dtype(), 'data': 'synthetic',
            'config': {'workload': workload_name(args, cfg), 'agents_per_gpu_per_step': B, 'agents_per_graph_replay': C,
                       'cuda_graph': bool(args.graph),
                       'global_agents_per_step': world * B, 'parallelism': f'dp{world} (agents sharded, no collective)',
                       'l2': 'inputs + activations per step >> 126 MB L2 (fresh synthetic tracks every step)',
                       'gflop_per_agent_reference_executed': GF_PER_AGENT[args.workload]},
            'e2e': {'value': e2e_value, 'unit': 'agent-trajectories/s', 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
            'torch_cuda_baseline': torch_cuda,
            'kernels': kernel_table,
            'diag': {'step_ms': step_list, 'host_issue_ms': host_ms, 'mem': mem_diag, 'kmeans_max_iters': km_diag,
                     'e2e_step_ms': [a.elapsed_time(b) for a, b in zip([e2] + e2e_marks[:-1], e2e_marks)],
                     'e2e_kmeans_max_iters': e2e_km},
            'effective_tflops_reference_normalised': value / (cfg['n_goal'] * cfg['n_traj']) *
                                                      GF_PER_AGENT[args.workload] / 1e3,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
