"""GPU parity tests of the network engine and the evaluate() driver, through the C ABI.

Tolerances (north_star): logits / heat maps <= 1e-3 relative in the fp32 engine; ADE/FDE within
0.05 px; sampled indices bit-exact at the op boundary (tests/test_gpu_ops.py).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, golden_state_dict
from helpers import build_product_model, ReplayRng, eval_cfg, rel_err
from oracle import ynet_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-3


@pytest.fixture(scope='module')
def ops(cuda_device):
    from motion_style_transfer_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------------ conv kernels
@pytest.mark.parametrize('cins,cout,H,W,N,relu', [((14,), 32, 32, 64, 2, True), ((5, 6), 12, 33, 37, 3, False),
                                                  ((16, 32, 1), 32, 64, 32, 2, True), ((130,), 130, 13, 13, 2, True),
                                                  ((3,), 70, 8, 40, 1, True)])
def test_conv3x3_direct_vs_torch(ops, cins, cout, H, W, N, relu):
    torch.manual_seed(0)
    xs = [torch.randn(N, c, H, W) for c in cins]
    w = torch.randn(cout, sum(cins), 3, 3) * 0.1
    b = torch.randn(cout)
    ref = F.conv2d(torch.cat(xs, 1), w, b, padding=1)
    ref = F.relu(ref) if relu else ref
    packed = ops.lora_fold(w.cuda(), packed=True)
    got = ops.conv3x3_f32([(x.cuda(), ops.SRC_DIRECT) for x in xs], packed, b.cuda(), relu, N, H, W)
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-5


def test_conv3x3_fused_pool_upsample_broadcast_modulo(ops):
    torch.manual_seed(1)
    N, H, W = 4, 16, 32
    a = torch.randn(N, 6, 2 * H, 2 * W)          # read through a fused 2x2 max-pool
    b = torch.randn(N, 5, H // 2, W // 2)        # read through fused bilinear x2
    c = torch.randn(1, 3, H, W)                  # broadcast over the batch (Tensor.expand)
    d = torch.randn(2, 4, H, W)                  # modulo-batched: image n reads d[n % 2]
    w = torch.randn(9, 18, 3, 3) * 0.1
    bias = torch.randn(9)
    x = torch.cat([F.max_pool2d(a, 2, 2), F.interpolate(b, scale_factor=2, mode='bilinear', align_corners=False),
                   c.expand(N, -1, -1, -1), d.repeat(2, 1, 1, 1)], 1)
    ref = F.relu(F.conv2d(x, w, bias, padding=1))
    got = ops.conv3x3_f32([(a.cuda(), ops.SRC_POOL2), (b.cuda(), ops.SRC_UP2), (c.cuda(), ops.SRC_DIRECT),
                           (d.cuda(), ops.SRC_DIRECT)], ops.lora_fold(w.cuda()), bias.cuda(), True, N, H, W)
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-5
    # expanded (stride-0) tensors are broadcast without a copy
    got2 = ops.conv3x3_f32([(a.cuda(), ops.SRC_POOL2), (b.cuda(), ops.SRC_UP2),
                            (c.cuda().expand(N, -1, -1, -1)[:1], ops.SRC_DIRECT), (d.cuda(), ops.SRC_DIRECT)],
                           ops.lora_fold(w.cuda()), bias.cuda(), True, N, H, W)
    assert torch.equal(got, got2)


def test_pool_upsample_conv1x1_predictor(ops):
    torch.manual_seed(2)
    x = torch.randn(2, 5, 16, 24)
    assert torch.equal(ops.maxpool2x2(x.cuda()).cpu(), F.max_pool2d(x, 2, 2))
    up = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
    assert torch.allclose(ops.upsample_bilinear2x(x.cuda()).cpu(), up, rtol=1e-5, atol=1e-6)
    w = torch.randn(30, 5, 1, 1)
    b = torch.randn(30)
    ref = F.conv2d(x, w, b)
    got = ops.conv1x1_f32(x.cuda(), w.reshape(30, 5).cuda(), b.cuda())
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-5
    sa = ops.predictor_softargmax_f32(x.cuda(), (w.reshape(30, 5) * 3).cuda(), b.cuda()).cpu()
    ref_sa = O.softargmax2d(F.conv2d(x, w * 3, b))
    np.testing.assert_allclose(sa.numpy(), ref_sa.numpy(), rtol=0, atol=5e-3)


def test_lora_fold_matches_loralib_semantics(ops):
    torch.manual_seed(3)
    for cout, cin, r in ((32, 14, 1), (64, 32, 2), (8, 5, 4)):
        w = torch.randn(cout, cin, 3, 3)
        A = torch.randn(3 * r, 3 * cin)
        Bm = torch.randn(3 * cout, 3 * r) * 0.02
        ref = w + (Bm @ A).view(w.shape) * (1.0 / r)
        got = ops.lora_fold(w.cuda(), A.cuda(), Bm.cuda(), packed=False).cpu()
        assert rel_err(got.numpy(), ref.numpy()) < 1e-6
        packed = ops.lora_fold(w.cuda(), A.cuda(), Bm.cuda(), packed=True).cpu()
        assert torch.equal(packed, got.permute(1, 2, 3, 0).reshape(cin, 9, cout))
    # B = 0 is an exact no-op (train.py:46-59 --init_check)
    got0 = ops.lora_fold(w.cuda(), A.cuda(), torch.zeros_like(Bm).cuda(), packed=False).cpu()
    assert torch.equal(got0, w)


# ------------------------------------------------------------------------------------------ whole network
@pytest.mark.parametrize('tag,network,kw', [('ynet', 'original', {}),
                                            ('ynetmod', 'fusion', dict(n_fusion=2, position=('scene', 'motion', 'fusion')))])
def test_network_golden(ops, tag, network, kw):
    from motion_style_transfer_b200.engine import ChannelCat
    g = load_golden(f'network_{tag}')
    sd = golden_state_dict(g)
    m = build_product_model(sd, 5, 6, 2, network=network, **kw)
    scene = torch.from_numpy(g['scene']).cuda()                    # (1, 6, H, W): broadcast over agents
    motion = torch.from_numpy(g['motion']).cuda()
    with torch.no_grad():
        feats = m.pred_features(scene, motion)
        assert len(feats) == 6
        for i, f in enumerate(feats):
            f = f.materialize() if isinstance(f, ChannelCat) else f
            assert f.shape == g[f'feat{i}'].shape
            assert rel_err(f.cpu().numpy(), g[f'feat{i}']) < REL, f'feature {i}'
        goal = m.pred_goal(feats)
        assert rel_err(goal.cpu().numpy(), g['goal']) < REL
        pyr = ops.avgpool_pyramid(torch.from_numpy(g['wp']).cuda(), 6)
        tin = [ChannelCat(tuple(f) + (p,)) if isinstance(f, tuple) else ChannelCat((f, p)) for f, p in zip(feats, pyr)]
        traj = m.pred_traj(tin)
        assert rel_err(traj.cpu().numpy(), g['traj']) < REL
        # drop-in form: the caller concatenates itself (evaluate.py:259)
        tin2 = [torch.cat([f.materialize() if isinstance(f, ChannelCat) else f, p], 1) for f, p in zip(feats, pyr)]
        assert rel_err(m.pred_traj(tin2).cpu().numpy(), g['traj']) < REL
        # fused predictor + soft-argmax
        sa = m.pred_traj_softargmax(tin).cpu().numpy()
        np.testing.assert_allclose(sa, O.softargmax2d(g['traj']).numpy(), rtol=0, atol=0.02)


def test_network_full_size_against_oracle(ops):
    """Full-width Y-Net (32/64 channels, mosa_1 on stages 0-4) at 416x416, 2 agents: logits <= 1e-3."""
    from motion_style_transfer_b200.models.ynet import YNet
    torch.manual_seed(0)
    m = YNet(obs_len=8, pred_len=12, segmentation_model_fp=None, encoder_channels=[32, 32, 64, 64, 64],
             decoder_channels=[64, 64, 64, 32, 32], n_waypoints=1, train_net='mosa_1', position=[0, 1, 2, 3, 4],
             network='original')
    gen = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if 'lora_B' in n:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.02)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    scene = O.synthetic_scene(416, 416, seed=0)[None]
    tracks = O.synthetic_tracks(2, 20, 416, 416, seed=1)
    tmpl = O.create_dist_mat(1050).astype(np.float32)
    obs = torch.from_numpy(O.get_patch_stack(tmpl, tracks[:, :8].reshape(-1, 2).numpy(), 416, 416)).view(2, 8, 416, 416)
    torch.set_num_threads(8)
    with torch.no_grad():
        feats_o = O.pred_features(sd, scene.expand(2, -1, -1, -1), obs)
        goal_o = O.pred_goal(sd, feats_o)
    m = m.cuda().eval()
    with torch.no_grad():
        feats = m.pred_features(scene.cuda(), obs.cuda())
        goal = m.pred_goal(feats)
    for i, (a, b) in enumerate(zip(feats, feats_o)):
        assert rel_err(a.cpu().numpy(), b.numpy()) < REL, f'feature {i}'
    assert rel_err(goal.cpu().numpy(), goal_o.numpy()) < REL


def test_state_dict_roundtrip_and_partial_checkpoint(ops, tmp_path):
    """Checkpoint layout (trainer.py:586-614): full dict and LoRA-only dict, strict=False loads."""
    g = load_golden('network_ynet')
    sd = golden_state_dict(g)
    m = build_product_model(sd, 5, 6, 2)
    assert list(m.state_dict().keys()) == list(sd.keys())
    lora_only = {k: v for k, v in m.state_dict().items() if 'lora' in k}
    assert len(lora_only) == 18
    torch.save(lora_only, tmp_path / 'tuned.pt')
    res = m.load_state_dict(torch.load(tmp_path / 'tuned.pt'), strict=False)
    assert not res.unexpected_keys


# ------------------------------------------------------------------------------------------ evaluate()
@pytest.mark.parametrize('name', ['eval_sdd_short', 'eval_ind_long_ttst_cws'])
def test_forecast_batch_golden(ops, name):
    from motion_style_transfer_b200.utils.evaluate import forecast_batch
    g = load_golden(name)
    c = eval_cfg(g)
    m = build_product_model(golden_state_dict(g), c['obs'], c['pred'], len(c['wps']))
    tmpl = ops.create_dist_template(int(g['template_size']), 'cuda')
    res = forecast_batch(m, torch.from_numpy(g['scene'])[None].cuda(), torch.from_numpy(g['trajectory']).cuda(), tmpl,
                         c['wps'], c['n_goal'], c['n_traj'], c['obs'], c['resize'], c['T'], c['ttst'], c['cws'],
                         c['thr'], c['cwsp'], rng=ReplayRng(g),
                         kmeans_init=g['kmeans_init'] if c['ttst'] else None, want_maps=True)
    assert rel_err(res['goal_map'].cpu().numpy(), g['goal_map']) < REL
    wps = res['waypoint_samples'].permute(1, 2, 0, 3).cpu().numpy()      # reference layout (B, n_wp, G, 2)
    if c['ttst']:
        # goal 0 = soft-argmax, goals 1.. = k-means centres of 10k draws; fp32-level logit differences can
        # move a handful of draws -> centres within a small fraction of a pixel
        np.testing.assert_allclose(wps, g['waypoint_sample'], rtol=0, atol=0.25)
    else:
        # top-k of p/q: logits agree to 1e-3, ties in p/q ordering are measure-zero -> same pixels
        assert (wps == g['waypoint_sample']).mean() > 0.97
    np.testing.assert_allclose(res['ade'].cpu().numpy(), g['ade'], rtol=0, atol=0.05 if not c['ttst'] else 0.5)
    np.testing.assert_allclose(res['fde'].cpu().numpy(), g['fde'], rtol=0, atol=0.05 if not c['ttst'] else 1.0)


def test_evaluate_dropin_signature_against_fixture(ops, monkeypatch):
    """The 23-argument evaluate() with a DataLoader, seeded like the reference run that wrote the fixture
    (RNG_MODE 'host': torch's / numpy's global generators consumed in the reference's order)."""
    import pandas as pd
    from torch.utils.data import DataLoader, Dataset
    from motion_style_transfer_b200.utils import evaluate as ev
    from motion_style_transfer_b200.utils.evaluate import evaluate
    monkeypatch.setattr(ev, 'RNG_MODE', 'host')
    g = load_golden('eval_sdd_short')
    c = eval_cfg(g)
    m = build_product_model(golden_state_dict(g), c['obs'], c['pred'], len(c['wps']))
    traj = torch.from_numpy(g['trajectory'])
    B = traj.shape[0]

    class OneScene(Dataset):
        def __len__(self):
            return 1

        def __getitem__(self, i):
            meta = pd.DataFrame({'metaId': np.repeat(np.arange(B), traj.shape[1])})
            return traj, meta, 's0'

    loader = DataLoader(OneScene(), batch_size=1, collate_fn=lambda b: (b[0][0], [b[0][1]], b[0][2]))
    tmpl = torch.from_numpy(O.create_dist_mat(int(g['template_size'])).astype(np.float32))
    torch.manual_seed(100)
    np.random.seed(200)
    ade, fde, df, td = evaluate(m, loader, {'s0': torch.from_numpy(g['scene'])}, 'cuda', 'sdd', None, tmpl, c['wps'],
                                'test', c['n_goal'], c['n_traj'], c['obs'], c['B'], c['resize'], c['T'], c['ttst'],
                                c['cws'], c['thr'], c['cwsp'], return_preds=True, return_samples=True)
    assert list(df.columns) == ['metaId', 'sceneId', 'ade', 'fde'] and len(df) == B
    np.testing.assert_allclose(df.ade.values, g['ade'], rtol=0, atol=0.05)
    np.testing.assert_allclose(df.fde.values, g['fde'], rtol=0, atol=0.05)
    assert td['goal_map'].shape == g['goal_map'].shape and td['waypoint_sample'].shape == g['waypoint_sample'].shape
    assert abs(ade - g['ade'].mean()) < 0.05 and abs(fde - g['fde'].mean()) < 0.05


# ---- serial / parallel adapter baselines (ynet.py:15-131, 237-283) -----------------------------------
from helpers import ADAPTER_TAGS, build_adapter_model     # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize('backend,tol', [('fp32', REL), ('bf16', 2.5e-2)])
@pytest.mark.parametrize('tag', ADAPTER_TAGS)
def test_adapter_baselines_against_reference_fixture(ops, tag, backend, tol):
    """Layer-level adapters folded into their conv, block-level adapters as one extra conv launch: encoder features and
    goal logits of the live reference (eval mode)."""
    g = load_golden(f'adapter_{tag}')
    m = build_adapter_model(g, 'cuda').set_backend(backend)
    scene = torch.from_numpy(g['scene']).cuda()
    motion = torch.from_numpy(g['motion']).cuda()
    with torch.no_grad():
        feats = m.pred_features(scene, motion)
        for i, f in enumerate(feats):
            f = ops.tc_unpack(f) if isinstance(f, ops.C8) else f
            assert rel_err(f.cpu().numpy(), g[f'feat{i}']) < tol, f'feature {i}'
        goal = m.pred_goal(feats)
        assert rel_err(goal.cpu().numpy(), g['goal']) < tol


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['parallelLayer_3x3', 'parallelLayer_1x1_3x3'])
def test_parallel_layer_adapter_gradients_against_reference(ops, tag):
    """Fine-tuning the parallel layer adapters (scripts/sdd/ped_to_biker/tune_pa.sh): gradients of the trainable tensors
    (trainer.py:133-135) from the device autograd engine vs the reference's autograd."""
    from motion_style_transfer_b200.models.trainer import apply_freeze_policy
    g = load_golden(f'adapter_{tag}')
    m = build_adapter_model(g, 'cuda')
    apply_freeze_policy(m, str(g['train_net']), [int(p) for p in g['position']], 'original')
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert names and all('parallel' in n for n in names)
    m.train()
    feats = m.pred_features(torch.from_numpy(g['scene']).cuda(), torch.from_numpy(g['motion']).cuda())
    goal = m.pred_goal(feats)
    assert rel_err(goal.detach().cpu().numpy(), g['goal']) < REL
    (goal.square().mean() * 100.0).backward()
    for n, p in m.named_parameters():
        if p.requires_grad:
            assert rel_err(p.grad.cpu().numpy(), g['grad/' + n]) < 2e-3, n


@pytest.mark.gpu
def test_serial_adapters_raise_in_training_mode(ops):
    g = load_golden('adapter_serialLayer')
    m = build_adapter_model(g, 'cuda')
    for p in m.encoder.parameters():
        p.requires_grad = True
    m.train()
    with pytest.raises(NotImplementedError):
        m.pred_features(torch.from_numpy(g['scene']).cuda(), torch.from_numpy(g['motion']).cuda())


@pytest.mark.gpu
def test_embed_network_against_reference_fixture(ops):
    """network='embed': the embedding layers (forward and backward through the float32 conv kernels) and the goal logits
    behind them vs the live reference."""
    from motion_style_transfer_b200.models.ynet import YNet
    g = load_golden('embed')
    torch.manual_seed(int(g['seed']))
    m = YNet(obs_len=5, pred_len=6, segmentation_model_fp=None, encoder_channels=[8, 8, 16, 16, 16],
             decoder_channels=[16, 16, 16, 8, 8], n_waypoints=2, train_net='all', position=[], network='embed')
    missing, unexpected = m.load_state_dict(golden_state_dict(g), strict=False)
    assert not unexpected and all(k.startswith(('goal_decoder.', 'traj_decoder.')) for k in missing)
    m = m.cuda()
    scene, motion = torch.from_numpy(g['scene']).cuda(), torch.from_numpy(g['motion']).cuda()
    with torch.no_grad():
        assert rel_err(m.scene_embedding(scene).cpu().numpy(), g['scene_emb']) < REL
        assert rel_err(m.motion_embedding(motion).cpu().numpy(), g['motion_emb']) < REL
    for p in m.parameters():
        p.requires_grad = True
    m.train()
    sc, mo = m.scene_embedding(scene), m.motion_embedding(motion)
    goal = m.pred_goal(m.pred_features(sc, mo))
    assert rel_err(goal.detach().cpu().numpy(), g['goal']) < REL
    (goal.square().mean() * 100.0).backward()
    for n, p in m.named_parameters():
        if 'embedding' in n:
            assert rel_err(p.grad.cpu().numpy(), g['grad/' + n]) < 2e-3, n


def test_evaluate_device_rng_graph_equals_eager_and_is_seeded(ops, monkeypatch):
    """evaluate() in its default mode (device generator; full batches replayed from ONE captured CUDA graph): the graph
    path returns exactly what the eager launches return, a re-seeded repeat of the same call index reproduces it, the next
    call draws other numbers, and a weight update is picked up (a new graph is captured)."""
    import pandas as pd
    from torch.utils.data import DataLoader, Dataset
    from motion_style_transfer_b200.utils import evaluate as ev
    g = load_golden('eval_ind_long_ttst_cws')
    c = eval_cfg(g)
    m = build_product_model(golden_state_dict(g), c['obs'], c['pred'], len(c['wps'])).set_backend('bf16')
    traj = torch.from_numpy(g['trajectory'])
    traj = torch.cat([traj + 0.37 * i for i in range(5)])[:9]            # 9 agents: four full batches of 2 and a tail of 1
    B = traj.shape[0]

    class OneScene(Dataset):
        def __len__(self):
            return 1

        def __getitem__(self, i):
            return traj, pd.DataFrame({'metaId': np.repeat(np.arange(B), traj.shape[1])}), 's0'

    loader = DataLoader(OneScene(), batch_size=1, collate_fn=lambda b: (b[0][0], [b[0][1]], b[0][2]))
    tmpl = torch.from_numpy(O.create_dist_mat(int(g['template_size'])).astype(np.float32))
    images = {'s0': torch.from_numpy(g['scene'])}

    def run(graph, call):
        monkeypatch.setattr(ev, 'USE_GRAPH', graph)
        monkeypatch.setattr(ev, '_eval_calls', call)
        torch.manual_seed(7)
        return ev.evaluate(m, loader, images, 'cuda', 'ind', None, tmpl, c['wps'], 'test', c['n_goal'], c['n_traj'], c['obs'],
                           2, c['resize'], c['T'], c['ttst'], c['cws'], c['thr'], c['cwsp'], return_preds=True)

    assert ev.RNG_MODE == 'device'
    monkeypatch.setattr(ev, 'EVAL_MIN_BATCH', 1)             # (default 128: small caller batches are merged)
    monkeypatch.setattr(ev, 'GRAPH_MIN_BATCHES', 3)          # (default 14: a capture must pay for itself)
    a = run(True, 0)
    assert len(m.__dict__['_forecast_graphs']) == 1
    first = next(iter(m.__dict__['_forecast_graphs'].values()))
    b = run(False, 0)
    assert np.array_equal(a[2].ade.values, b[2].ade.values) and np.array_equal(a[2].fde.values, b[2].fde.values)
    assert np.array_equal(a[3]['prediction'], b[3]['prediction'])
    a2 = run(True, 0)
    assert np.array_equal(a[2].ade.values, a2[2].ade.values)
    nxt = run(True, 1)                                                   # the next round draws other samples
    assert not np.array_equal(a[2].fde.values, nxt[2].fde.values)
    with torch.no_grad():
        m.goal_decoder.predictor.bias.add_(0.5)
    upd = run(True, 0)
    graphs = list(m.__dict__['_forecast_graphs'].values())
    assert len(graphs) == 1 and graphs[0] is not first and graphs[0].pool_bytes >= 0      # re-captured for the new weights
    ref = run(False, 0)
    assert np.array_equal(upd[2].ade.values, ref[2].ade.values)


def test_evaluate_merges_small_batches(ops, monkeypatch):
    """evaluate(batch_size=2) with the device generator forecasts EVAL_MIN_BATCH agents per launch sequence: same per-agent
    rows as one call with the large batch (same streams: the batches coincide), fewer launches than 2-agent batches."""
    import pandas as pd
    from torch.utils.data import DataLoader, Dataset
    from motion_style_transfer_b200.utils import evaluate as ev
    g = load_golden('eval_sdd_short')
    c = eval_cfg(g)
    m = build_product_model(golden_state_dict(g), c['obs'], c['pred'], len(c['wps'])).set_backend('bf16')
    traj = torch.cat([torch.from_numpy(g['trajectory']) + 0.41 * i for i in range(4)])[:7]
    B = traj.shape[0]

    class OneScene(Dataset):
        def __len__(self):
            return 1

        def __getitem__(self, i):
            return traj, pd.DataFrame({'metaId': np.repeat(np.arange(B), traj.shape[1])}), 's0'

    loader = DataLoader(OneScene(), batch_size=1, collate_fn=lambda b: (b[0][0], [b[0][1]], b[0][2]))
    tmpl = torch.from_numpy(O.create_dist_mat(int(g['template_size'])).astype(np.float32))
    images = {'s0': torch.from_numpy(g['scene'])}
    monkeypatch.setattr(ev, 'USE_GRAPH', False)

    def run(bs, min_batch):
        monkeypatch.setattr(ev, 'EVAL_MIN_BATCH', min_batch)
        monkeypatch.setattr(ev, '_eval_calls', 0)
        torch.manual_seed(3)
        l0 = ops.launch_count
        out = ev.evaluate(m, loader, images, 'cuda', 'sdd', None, tmpl, c['wps'], 'test', c['n_goal'], c['n_traj'], c['obs'], bs,
                          c['resize'], c['T'], c['ttst'], c['cws'], c['thr'], c['cwsp'])
        return out, ops.launch_count - l0

    run(B, 1)                                            # (weight packing, template planes, autotuning: one-time launches)
    merged, n_merged = run(2, 128)
    whole, n_whole = run(B, 1)
    small, n_small = run(2, 1)
    assert np.array_equal(merged[2].ade.values, whole[2].ade.values) and n_merged == n_whole
    assert n_small > 2 * n_merged and len(small[2]) == B
