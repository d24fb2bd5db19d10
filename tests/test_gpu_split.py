"""GPU tests of the split-bf16 ("bf16x3") engine -- the <= 1e-3 parity mode on the tensor cores (csrc/split_tc.cu,
ynet_tc_conv3x3_split) -- through the C ABI: layout round trip, pooling / bilinear companions and the three-MMA conv
against float64 torch, then the whole network against the live-reference fixtures and, at full width and 416^2, against
the oracle (models/ynet.py:302-470).

STATED TOLERANCE: one conv <= 2e-5 of max|ref| (2^-17 operand splits, float32 accumulation); network logits <= 1e-3 of
max|ref| like the fp32 engine (tests/test_gpu_network.py); soft-argmax coordinates <= 0.02 px.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, golden_state_dict
from helpers import build_product_model, rel_err
from oracle import ynet_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-3


@pytest.fixture(scope='module')
def ops(cuda_device):
    from motion_style_transfer_b200 import ops as _ops
    if not _ops.tc_supported():
        pytest.fail('tensor-core engine unavailable on this device (needs sm_100 + cuTensorMapEncodeTiled)')
    return _ops


def test_split_layout_roundtrip_pool_upsample(ops):
    torch.manual_seed(0)
    x = torch.randn(3, 21, 12, 20) * torch.logspace(-3, 3, 21).view(1, -1, 1, 1)
    a = ops.split_pack(x.cuda())
    assert a.data.shape == (3, 8, 12, 20, 8) and a.C == 21 and a.cp == 32
    back = ops.split_unpack(a).cpu()
    assert ((back - x).abs() <= 2.0 ** -16 * x.abs()).all()
    # hi planes are the plain bf16 rounding; padded channels are zero
    hi = a.data[:, :4].float().permute(0, 1, 4, 2, 3).reshape(3, 32, 12, 20).cpu()
    assert torch.equal(hi[:, :21], x.to(torch.bfloat16).float()) and not hi[:, 21:].any()
    x16 = back                                            # values exactly representable as hi + lo
    mp = ops.split_unpack(ops.split_maxpool(ops.split_pack(x16.cuda()))).cpu()
    assert torch.equal(mp, F.max_pool2d(x16, 2, 2))
    up = ops.split_unpack(ops.split_upsample(ops.split_pack(x16.cuda()))).cpu()
    ref = F.interpolate(x16.double(), scale_factor=2, mode='bilinear', align_corners=False)
    assert rel_err(up.numpy(), ref.numpy()) < 2e-5
    # fewer real channels than padded planes (C = 8 -> one used chunk of two), several images
    x8 = torch.randn(3, 8, 6, 10)
    assert ((ops.split_unpack(ops.split_pack(x8.cuda())).cpu() - x8).abs() <= 2.0 ** -16 * x8.abs()).all()
    # broadcast batch (stride 0) packs one image
    b = ops.split_pack(x[:1].cuda().expand(4, -1, -1, -1))
    assert b.N == 1


@pytest.mark.parametrize('cins,cout,H,W,N,relu', [
    ([32], 32, 24, 40, 2, True),
    ([14], 30, 17, 23, 3, False),            # padded channels on both sides, odd sizes
    ([64, 18], 64, 16, 16, 2, True),         # two sources = torch.cat
    ([64, 64, 2], 64, 26, 26, 4, True),      # three parts: concatenated into one source
    ([144], 144, 13, 13, 2, True),           # 27 K blocks
    ([32], 32, 416, 416, 1, True),
])
def test_conv3x3_split_vs_float64(ops, cins, cout, H, W, N, relu):
    from motion_style_transfer_b200.engine import YNetEngineSplit
    torch.manual_seed(1)
    xs = [torch.randn(N, c, H, W) for c in cins]
    w = torch.randn(cout, sum(cins), 3, 3) * 0.1
    b = torch.randn(cout)
    ref = F.conv2d(torch.cat(xs, 1).double(), w.double(), b.double(), padding=1)
    ref = F.relu(ref) if relu else ref
    parts = [ops.split_pack(x.cuda()) for x in xs]
    sources, ranges = YNetEngineSplit._group(parts)
    assert len(sources) <= 2
    idx = torch.cat([torch.arange(c0, c1) for c0, c1 in ranges])
    packed = ops.split_pack_weights(w[:, idx].contiguous().cuda(), [s.layout for s in sources])
    bias = torch.zeros((cout + 15) // 16 * 16)
    bias[:cout] = b
    out = ops.tc_conv3x3_split(sources, packed, bias.cuda(), cout, relu)
    torch.cuda.synchronize()
    got = ops.split_unpack(out).cpu()
    assert got.shape == ref.shape
    err = rel_err(got.numpy(), ref.numpy())
    # the plain bf16 engine on the same operands, for scale
    print(f'split conv {cins}->{cout}@{H}x{W}: rel err {err:.2e}')
    assert err < 2e-5


def test_conv3x3_split_modulo_batch_and_1x1(ops):
    """Goal-major stacking (evaluate.py:248-266): image g * B + b reads the per-agent source b; 1x1 predictor -> float32."""
    torch.manual_seed(2)
    B, G, H, W = 3, 4, 20, 28
    feat = torch.randn(B, 16, H, W)
    up = torch.randn(G * B, 16, H, W)
    wp = torch.rand(G * B, 2, H, W)
    w = torch.randn(32, 34, 3, 3) * 0.1
    b = torch.randn(32)
    ref = F.relu(F.conv2d(torch.cat([up, feat.repeat(G, 1, 1, 1), wp], 1).double(), w.double(), b.double(), padding=1))
    from motion_style_transfer_b200.engine import YNetEngineSplit
    parts = [ops.split_pack(t.cuda()) for t in (up, feat, wp)]
    sources, ranges = YNetEngineSplit._group(parts)
    assert [s.N for s in sources] == [G * B, B] and ranges == ((0, 16), (32, 34), (16, 32))
    idx = torch.cat([torch.arange(c0, c1) for c0, c1 in ranges])
    packed = ops.split_pack_weights(w[:, idx].contiguous().cuda(), [s.layout for s in sources])
    y = ops.tc_conv3x3_split(sources, packed, b.cuda(), 32, True)
    assert rel_err(ops.split_unpack(y).cpu().numpy(), ref.numpy()) < 2e-5
    # concat-on-write: a conv and a pack fill the two ends of one activation; same bits as the copied concatenation
    w_up = torch.randn(16, 8, 3, 3) * 0.2
    x0 = ops.split_pack(torch.randn(G * B, 8, H, W).cuda())
    pk_up = ops.split_pack_weights(w_up.cuda(), [x0.layout])
    up_s = ops.tc_conv3x3_split([x0], pk_up, torch.zeros(16).cuda(), 16, False)
    buf = ops.split_empty(G * B, [(16, 16), (2, 16)], H, W, 'cuda')
    ops.tc_conv3x3_split([x0], pk_up, torch.zeros(16).cuda(), 16, False, into=(buf, 0))
    ops.split_pack(wp.cuda(), into=(buf, 16))
    cat = ops.split_cat([up_s, parts[2]])
    assert buf.layout == cat.layout and torch.equal(buf.data, cat.data)
    wp1 = torch.randn(30, 32, 1, 1)
    bp = torch.randn(30)
    bias = torch.zeros(32)
    bias[:30] = bp
    logits = ops.tc_conv1x1_split_f32(y, ops.split_pack_weights(wp1.cuda(), [y.layout]), bias.cuda(), 30)
    ref1 = F.conv2d(ops.split_unpack(y).cpu().double(), wp1.double(), bp.double())
    assert rel_err(logits.cpu().numpy(), ref1.numpy()) < 2e-5


@pytest.mark.parametrize('tag,network,kw', [('ynet', 'original', {}),
                                            ('ynetmod', 'fusion', dict(n_fusion=2, position=('scene', 'motion', 'fusion')))])
def test_network_golden_split(ops, tag, network, kw):
    """Live-reference fixtures network_ynet / network_ynetmod (encoder features, goal and trajectory logits)."""
    from motion_style_transfer_b200.engine import ChannelCat
    g = load_golden(f'network_{tag}')
    m = build_product_model(golden_state_dict(g), 5, 6, 2, network=network, **kw).set_backend('bf16x3')
    scene = torch.from_numpy(g['scene']).cuda()
    motion = torch.from_numpy(g['motion']).cuda()
    with torch.no_grad():
        feats = m.pred_features(scene, motion)
        assert len(feats) == 6
        for i, f in enumerate(feats):
            parts = [ops.split_unpack(t) for t in (f if isinstance(f, ChannelCat) else (f,))]
            N = max(t.shape[0] for t in parts)
            f32 = torch.cat([t.expand(N, -1, -1, -1) for t in parts], 1)
            assert f32.shape == g[f'feat{i}'].shape
            assert rel_err(f32.cpu().numpy(), g[f'feat{i}']) < REL, f'feature {i}'
        goal = m.pred_goal(feats)
        assert rel_err(goal.cpu().numpy(), g['goal']) < REL
        pyr = ops.avgpool_pyramid(torch.from_numpy(g['wp']).cuda(), 6)
        tin = [ChannelCat(tuple(f) + (p,)) if isinstance(f, tuple) else ChannelCat((f, p)) for f, p in zip(feats, pyr)]
        traj = m.pred_traj(tin)
        err = rel_err(traj.cpu().numpy(), g['traj'])
        print(f'[bf16x3] {tag}: goal logits rel {rel_err(goal.cpu().numpy(), g["goal"]):.2e}, trajectory logits rel {err:.2e}')
        assert err < REL
        sa = m.pred_traj_softargmax(tin).cpu().numpy()
        np.testing.assert_allclose(sa, O.softargmax2d(g['traj']).numpy(), rtol=0, atol=0.02)


def test_network_full_size_split_against_oracle(ops):
    """Full-width Y-Net (32/64 channels, mosa_1 on stages 0-4) at 416x416, 2 agents: features and goal logits <= 1e-3
    (the same case as tests/test_gpu_network.py::test_network_full_size_against_oracle runs on the fp32 engine)."""
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
    m = m.cuda().eval().set_backend('bf16x3')
    with torch.no_grad():
        feats = m.pred_features(scene.cuda(), obs.cuda())
        goal = m.pred_goal(feats)
    for i, (a, b) in enumerate(zip(feats, feats_o)):
        assert rel_err(ops.split_unpack(a).cpu().numpy(), b.numpy()) < REL, f'feature {i}'
    err = rel_err(goal.cpu().numpy(), goal_o.numpy())
    print(f'[bf16x3] full width 416^2: goal logits rel {err:.2e}')
    assert err < REL
