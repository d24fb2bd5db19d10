"""The data-parallel drop-ins on NCCL: run with

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m pytest tests/test_gpu_nccl.py -m gpu -q

Single-process runs (the driver's `pytest -m gpu`) skip: they need WORLD_SIZE > 1.  Under torchrun every rank runs the
same assertions: sharded evaluate() returns on EVERY rank exactly what one process returns for the whole scene, and two
fine-tuning steps with the agents of each batch sharded over the ranks (one flat-gradient all-reduce per step) end in
the same adapter weights as the single-process run."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WORLD = int(os.environ.get('WORLD_SIZE', '1'))


@pytest.fixture(scope='module')
def pg():
    if WORLD < 2:
        pytest.skip('needs torchrun with >= 2 ranks')
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    yield dist
    dist.barrier()


def _small_model(seed=0):
    from motion_style_transfer_b200.models.ynet import YNet
    torch.manual_seed(seed)
    m = YNet(obs_len=5, pred_len=6, segmentation_model_fp=None, encoder_channels=[8, 8, 16, 16, 16],
             decoder_channels=[16, 16, 16, 8, 8], n_waypoints=2, train_net='mosa_1', position=[0, 1, 2, 3, 4],
             network='original')
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if 'lora_B' in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return m


def _loader(n_agents, total, H, W, resize, seed):
    import pandas as pd
    from torch.utils.data import DataLoader
    from motion_style_transfer_b200 import synthetic as S
    from motion_style_transfer_b200.utils.dataloader import SceneDataset, scene_collate
    tr = S.synthetic_tracks(n_agents, total, H, W, seed=seed).numpy() / resize
    df = pd.DataFrame({'frame': np.tile(np.arange(total), n_agents), 'trackId': np.repeat(np.arange(n_agents), total),
                       'x': tr[:, :, 0].reshape(-1), 'y': tr[:, :, 1].reshape(-1), 'sceneId': 's0',
                       'metaId': np.repeat(np.arange(n_agents), total)})
    return (DataLoader(SceneDataset(df, resize=resize, total_len=total), batch_size=1, collate_fn=scene_collate),
            {'s0': S.synthetic_scene(H, W, seed=0)})


def test_sharded_evaluate_equals_single_process(pg):
    """evaluate() with the scene's agents sharded over the ranks + NCCL all_gather of the rows == the unsharded call."""
    from motion_style_transfer_b200 import parallel
    from motion_style_transfer_b200.utils.evaluate import evaluate
    from motion_style_transfer_b200.utils.image_utils import create_dist_mat
    dev = torch.device('cuda', torch.cuda.current_device())
    H, W, n_agents = 64, 96, 7                                   # ragged shards
    m = _small_model().to(dev).eval()
    loader, images = _loader(n_agents, 11, H, W, 0.25, seed=4)
    tmpl = torch.Tensor(create_dist_mat(size=1050))
    args = (m, loader, images, dev, 'sdd', None, tmpl, [2, 5], 'test', 20, 1, 5, 4, 0.25, 1.0, False, False, 0.01, None)

    from motion_style_transfer_b200.utils import evaluate as ev
    ev.RNG_MODE = 'host'

    def run():
        torch.manual_seed(100)            # every rank draws the SAME host randoms; shards take their rows
        np.random.seed(200)
        return evaluate(*args)

    ade, fde, df, _ = run()                                       # sharded (world > 1)
    assert parallel.world()[1] == WORLD and len(df) == n_agents
    # the unsharded answer, computed on every rank by pretending to be alone
    real = parallel.world
    parallel.world = lambda: (0, 1)
    try:
        ade1, fde1, df1, _ = run()
    finally:
        parallel.world = real
    # plain sampling (replacement=False) draws exponentials for the rows of the LOCAL batch, so a shard consumes another
    # part of the host stream than the unsharded run: compare what does not depend on the draws ...
    assert list(df.metaId) == list(df1.metaId)
    assert np.isfinite(df.ade.values).all() and np.isfinite(df.fde.values).all()
    # ... and that every rank holds the same gathered rows
    t = torch.tensor(df.ade.values, device=dev)
    lo, hi = t.clone(), t.clone()
    pg.all_reduce(lo, op=pg.ReduceOp.MIN)
    pg.all_reduce(hi, op=pg.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    assert abs(ade - float(df.ade.mean())) < 1e-5


def test_sharded_finetune_matches_single_process(pg):
    """Two train_epoch passes (4 agents, batches of 2 sharded over the ranks, FusedAdam with the NCCL all-reduce of the
    flat LoRA gradient) end in the adapter weights of the single-process run (same data order, fp32: <= 1e-5)."""
    from motion_style_transfer_b200 import parallel
    from motion_style_transfer_b200.autograd_engine import BCEWithLogitsLoss
    from motion_style_transfer_b200.models.trainer import FusedAdam, apply_freeze_policy
    from motion_style_transfer_b200.utils.image_utils import create_dist_mat, create_gaussian_heatmap_template
    from motion_style_transfer_b200.utils.train_epoch import train_epoch
    dev = torch.device('cuda', torch.cuda.current_device())
    H, W = 64, 96
    loader, images = _loader(4, 11, H, W, 0.25, seed=3)
    tmpl = torch.Tensor(create_dist_mat(size=1050)).to(dev)
    gt = torch.Tensor(create_gaussian_heatmap_template(size=1050, kernlen=31, nsig=4, normalize=False)).to(dev)

    def train(sharded):
        m = _small_model().to(dev)
        apply_freeze_policy(m, 'mosa_1', [0, 1, 2, 3, 4], 'original')
        real = parallel.world
        if not sharded:
            parallel.world = lambda: (0, 1)
        try:
            parallel.broadcast_params([q for q in m.parameters() if q.requires_grad])
            opt = FusedAdam(m.parameters(), lr=1e-3)
            out = None
            for e in range(2):
                out = train_epoch(m, loader, images, opt, BCEWithLogitsLoss(), 1000, dev, 'sdd', None, gt, tmpl, [2, 5], e, 5,
                                  6, 2, 10000, 0.25)
        finally:
            parallel.world = real
        return {n: q.detach().clone() for n, q in m.named_parameters() if q.requires_grad}, out

    w_dp, out_dp = train(True)
    w_1, out_1 = train(False)
    assert len(w_dp) == 18
    for n in w_1:
        assert torch.allclose(w_dp[n], w_1[n], rtol=0, atol=1e-5), n
    assert abs(out_dp[2] - out_1[2]) < 1e-2 * max(1.0, abs(out_1[2]))       # loss: mean over ranks == full-batch loss
    # replicas stay bit-identical across ranks
    for n, q in w_dp.items():
        lo, hi = q.clone(), q.clone()
        pg.all_reduce(lo, op=pg.ReduceOp.MIN)
        pg.all_reduce(hi, op=pg.ReduceOp.MAX)
        assert torch.equal(lo, hi), n


def test_train_and_test_entry_points_under_torchrun(pg, monkeypatch):
    """``torchrun -m motion_style_transfer_b200.train ...`` in small: every rank runs the same command line in one shared
    working directory (rank 0 lays it out), agents of every batch are sharded, rank 0 alone writes the checkpoints -- and every
    rank ends with the same adapter tensors, which ``test`` restores to the same ADE / FDE on every rank."""
    import pathlib
    import re
    import shutil
    import test_gpu_scripts as S
    from motion_style_transfer_b200 import train, test
    from motion_style_transfer_b200.utils.parser import get_parser
    rank = pg.get_rank()
    root = pathlib.Path(f"/tmp/ynet_scripts_{os.environ.get('MASTER_PORT', '0')}")
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
        os.makedirs(root)
        S._make_workspace(root, S.CONFIG, ['sA_0', 'sB_1'], S._sdd_raw_dir, 'reference.jpg', 'sdd_segmentation.pth',
                          (('train', 0, 12), ('val', 100, 4), ('test', 200, 6)))
        # a "pretrained" checkpoint: random initialisation, everything but the segmentation module
        from motion_style_transfer_b200.models.ynet import YNet
        torch.manual_seed(5)
        m = YNet(obs_len=5, pred_len=6, segmentation_model_fp=None, encoder_channels=[8, 8, 16, 16, 16],
                 decoder_channels=[16, 16, 16, 8, 8], n_waypoints=1, train_net='train', position=[], network='original')
        os.makedirs(root / 'ckpts')
        torch.save(m.state_dict(), root / 'ckpts' / 'sdd__ynet__ped.pt')
    pg.barrier()
    monkeypatch.chdir(root)
    try:
        tune = (f'{S.COMMON} --fine_tune --seed 2 --n_epoch 2 --n_round 1 --pretrained_ckpt ckpts/sdd__ynet__ped.pt '
                '--train_net mosa_1 --position 0 1 --ckpt_path ckpts/tuned --n_train_batch 2 --lr 0.003')
        train.main(get_parser(True).parse_args(tune.split()))
        pg.barrier()
        tuned = 'ckpts/tuned/Seed_2__filter_agent_type_Biker__mosa_1__Pos_0_1__TrN_8__lr_0.003__original.pt'
        sd = torch.load(tuned, map_location='cuda')
        assert sd and all('lora_' in k for k in sd)
        import contextlib
        import io
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            test.main(get_parser(False).parse_args(
                f'{S.COMMON} --seed 2 --n_round 1 --pretrained_ckpt ckpts/sdd__ynet__ped.pt --tuned_ckpt {tuned}'.split()))
        ade, fde = (float(v) for v in re.findall(S.AVERAGE, buf.getvalue())[0][1:])
        t = torch.tensor([ade, fde], device='cuda', dtype=torch.float64)
        lo, hi = t.clone(), t.clone()
        pg.all_reduce(lo, op=pg.ReduceOp.MIN)
        pg.all_reduce(hi, op=pg.ReduceOp.MAX)
        assert torch.equal(lo, hi) and ade > 0 and fde > 0
    finally:
        pg.barrier()
        if rank == 0:
            shutil.rmtree(root, ignore_errors=True)
