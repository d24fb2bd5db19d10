"""GPU tests of the tensor-core (tcgen05 / TMEM / TMA) conv engine, through the C ABI.

bf16 operands, fp32 accumulation.  STATED TOLERANCE of the bf16 mode: single conv with bf16-exact
inputs <= 5e-3 of max|ref| (output rounding to bf16 = 2^-9 relative); whole network logits <= 3e-2 of
max|ref| (SURVEY 7 "hard parts" measured 1.8e-3..5e-3 for bf16 operands on random-init weights);
ADE/FDE within 0.05 px on the non-TTST path is checked in test_forecast_bf16.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, golden_state_dict
from helpers import build_product_model, ReplayRng, eval_cfg, rel_err
from oracle import ynet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(cuda_device):
    from motion_style_transfer_b200 import ops as _ops
    if not _ops.tc_supported():
        pytest.fail('tensor-core engine unavailable on this device (needs sm_100 + cuTensorMapEncodeTiled)')
    return _ops


def bf16_exact(t):
    return t.to(torch.bfloat16).to(torch.float32)


def test_pack_unpack_roundtrip(ops):
    torch.manual_seed(0)
    x = bf16_exact(torch.randn(3, 13, 20, 24))
    a = ops.tc_pack(x.cuda())
    assert a.C == 13 and a.C_pad == 16 and a.data.shape == (3, 2, 20, 24, 8)
    assert torch.equal(ops.tc_unpack(a).cpu(), x)
    # pad channels are zero
    assert float(a.data[:, 1, :, :, 5:].abs().max()) == 0.0
    p = ops.tc_maxpool(a)
    assert torch.equal(ops.tc_unpack(p).cpu(), F.max_pool2d(x, 2, 2))
    u = ops.tc_upsample(a)
    ref = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
    assert rel_err(ops.tc_unpack(u).cpu().numpy(), ref.numpy()) < 5e-3


def _tc_conv(ops, xs, w, b, relu, N):
    srcs = [ops.tc_pack(x.cuda()) for x in xs]
    packed = ops.tc_pack_weights(w.cuda(), [x.shape[1] for x in xs])
    cout = w.shape[0]
    bias = torch.zeros((cout + 15) // 16 * 16)
    bias[:cout] = b
    out = ops.tc_conv3x3(srcs, packed, bias.cuda(), cout, relu)
    torch.cuda.synchronize()
    return ops.tc_unpack(out).cpu()


@pytest.mark.parametrize('cins,cout,H,W,N,relu', [
    ((16,), 32, 16, 8, 1, False),          # exactly one tile, one K block
    ((32,), 32, 32, 24, 2, True),
    ((14,), 32, 48, 40, 2, True),          # padded input channels
    ((16, 32, 1), 32, 64, 32, 3, True),    # traj decoder.4.0 shape: [up | feature | waypoint]
    ((64,), 64, 52, 52, 2, True),          # partial tiles (52 = 3*16+4 = 6*8+4)
    ((64, 1), 12, 26, 26, 2, False),       # C_out padded 12 -> 16
    ((130,), 130, 13, 13, 2, True),        # centre conv: N = 144, weights streamed through the pipeline
    ((32,), 16, 416, 416, 1, False),       # full-resolution upsample_conv.4
])
def test_tc_conv3x3_vs_torch(ops, cins, cout, H, W, N, relu):
    torch.manual_seed(1)
    xs = [bf16_exact(torch.randn(N, c, H, W)) for c in cins]
    w = bf16_exact(torch.randn(cout, sum(cins), 3, 3) * 0.1)
    b = torch.randn(cout)
    ref = F.conv2d(torch.cat(xs, 1), w, b, padding=1)
    ref = F.relu(ref) if relu else ref
    got = _tc_conv(ops, xs, w, b, relu, N)
    assert got.shape == ref.shape
    assert rel_err(got.numpy(), ref.numpy()) < 5e-3


def test_tc_conv_broadcast_modulo_and_streamed(ops):
    torch.manual_seed(2)
    N, H, W = 4, 32, 32
    a = bf16_exact(torch.randn(1, 6, H, W))       # broadcast
    b = bf16_exact(torch.randn(2, 32, H, W))      # modulo: n % 2
    c = bf16_exact(torch.randn(N, 2, H, W))
    w = bf16_exact(torch.randn(32, 40, 3, 3) * 0.1)
    bias = torch.randn(32)
    x = torch.cat([a.expand(N, -1, -1, -1), b.repeat(2, 1, 1, 1), c], 1)
    ref = F.relu(F.conv2d(x, w, bias, padding=1))
    got = _tc_conv(ops, [a, b, c], w, bias, True, N)
    assert rel_err(got.numpy(), ref.numpy()) < 5e-3
    os.environ['YNET_TC_FORCE_STREAMED'] = '1'
    try:
        got2 = _tc_conv(ops, [a, b, c], w, bias, True, N)
    finally:
        del os.environ['YNET_TC_FORCE_STREAMED']
    assert torch.equal(got, got2)


@pytest.mark.parametrize('tag,network,kw', [('ynet', 'original', {}),
                                            ('ynetmod', 'fusion', dict(n_fusion=2, position=('scene', 'motion', 'fusion')))])
def test_network_bf16_golden(ops, tag, network, kw):
    from motion_style_transfer_b200.engine import ChannelCat
    g = load_golden(f'network_{tag}')
    m = build_product_model(golden_state_dict(g), 5, 6, 2, network=network, **kw).set_backend('bf16')
    with torch.no_grad():
        feats = m.pred_features(torch.from_numpy(g['scene']).cuda(), torch.from_numpy(g['motion']).cuda())
        goal = m.pred_goal(feats)
        pyr = ops.avgpool_pyramid(torch.from_numpy(g['wp']).cuda(), 6)
        tin = [ChannelCat(tuple(f) + (p,)) if isinstance(f, tuple) else ChannelCat((f, p)) for f, p in zip(feats, pyr)]
        traj = m.pred_traj(tin)
    assert rel_err(goal.cpu().numpy(), g['goal']) < 3e-2
    assert rel_err(traj.cpu().numpy(), g['traj']) < 3e-2


def test_forecast_bf16_sdd_short(ops):
    from motion_style_transfer_b200.utils.evaluate import forecast_batch
    g = load_golden('eval_sdd_short')
    c = eval_cfg(g)
    m = build_product_model(golden_state_dict(g), c['obs'], c['pred'], len(c['wps'])).set_backend('bf16')
    tmpl = ops.create_dist_template(int(g['template_size']), 'cuda')
    res = forecast_batch(m, torch.from_numpy(g['scene'])[None].cuda(), torch.from_numpy(g['trajectory']).cuda(), tmpl,
                         c['wps'], c['n_goal'], c['n_traj'], c['obs'], c['resize'], c['T'], c['ttst'], c['cws'],
                         c['thr'], c['cwsp'], rng=ReplayRng(g), want_maps=True)
    assert rel_err(res['goal_map'].cpu().numpy(), g['goal_map']) < 3e-2
    # ADE is a soft-argmax expectation: robust to bf16; FDE depends on sampled goal pixels (top-k of p/q)
    np.testing.assert_allclose(res['ade'].cpu().numpy(), g['ade'], rtol=0, atol=0.25)


@pytest.mark.parametrize('cin,cout,H,W,N', [(32, 30, 64, 96, 3), (32, 12, 416, 416, 2), (8, 6, 52, 40, 2)])
def test_tc_predictor_logits_and_fused_softargmax(ops, cin, cout, H, W, N):
    """1x1 predictor on the tensor cores: float32 logits, and the fused predictor + soft-argmax epilogue."""
    torch.manual_seed(3)
    x = bf16_exact(torch.relu(torch.randn(N, cin, H, W)))
    w = bf16_exact(torch.randn(cout, cin, 1, 1) * 0.5)
    b = torch.randn(cout)
    ref = F.conv2d(x, w, b)
    a = ops.tc_pack(x.cuda())
    packed = ops.tc_pack_weights(w.cuda(), [cin])
    bias = torch.zeros((cout + 15) // 16 * 16)
    bias[:cout] = b
    logits = ops.tc_conv1x1_f32(a, packed, bias.cuda(), cout).cpu()
    assert rel_err(logits.numpy(), ref.numpy()) < 1e-4          # fp32 accumulate of bf16-exact operands, fp32 out
    sa = ops.tc_conv1x1_softargmax(a, packed, bias.cuda(), cout).cpu().numpy()
    np.testing.assert_allclose(sa, O.softargmax2d(ref).numpy(), rtol=0, atol=5e-3)
    # peaky logits (what TTST sees after training): scale the predictor
    packed50 = ops.tc_pack_weights((w * 8).cuda(), [cin])
    sa50 = ops.tc_conv1x1_softargmax(a, packed50, (bias * 8).cuda(), cout).cpu().numpy()
    np.testing.assert_allclose(sa50, O.softargmax2d(F.conv2d(x, w * 8, b * 8)).numpy(), rtol=0, atol=2e-2)


def test_tc_conv_repeat_interleaved_source(ops):
    """Agent-major stacking: a source with ``rep`` is read as image n // rep (no copy)."""
    torch.manual_seed(4)
    nb, G, H, W = 3, 4, 32, 24
    feat = bf16_exact(torch.randn(nb, 32, H, W))
    wp = bf16_exact(torch.randn(nb * G, 2, H, W))
    w = bf16_exact(torch.randn(32, 34, 3, 3) * 0.1)
    bias = torch.randn(32)
    ref = F.relu(F.conv2d(torch.cat([feat.repeat_interleave(G, dim=0), wp], 1), w, bias, padding=1))
    f8 = ops.tc_pack(feat.cuda())
    srcs = [f8.batch_slice(0, nb).repeat_interleave(G), ops.tc_pack(wp.cuda())]
    assert srcs[0].N == nb * G
    packed = ops.tc_pack_weights(w.cuda(), [32, 2])
    out = ops.tc_conv3x3(srcs, packed, bias.cuda(), 32, True)
    assert rel_err(ops.tc_unpack(out).cpu().numpy(), ref.numpy()) < 5e-3
    # a slice of the agents (chunked decoding): rows 1..2 only
    out2 = ops.tc_conv3x3([f8.batch_slice(1, 3).repeat_interleave(G), ops.tc_pack(wp[G:].contiguous().cuda())], packed,
                          bias.cuda(), 32, True)
    assert torch.equal(ops.tc_unpack(out2), ops.tc_unpack(out)[G:])


@pytest.mark.parametrize('n_img,n_ch,H,W', [(5, 2, 64, 96), (3, 1, 32, 32), (2, 8, 96, 64)])
def test_tc_rasterize_pyramid_vs_oracle(ops, n_img, n_ch, H, W):
    """get_patch + AvgPool2d(2^i) pyramid written straight as bf16 C8: equal to the oracle's float32 maps rounded
    to bf16 (level 0 exactly; pooled levels within one bf16 ulp = 2^-8 relative, the 2x2 cascade sums in a
    different order than AvgPool2d(2^i))."""
    size = 300
    tmpl = O.create_dist_mat(size).astype(np.float32)
    g = torch.Generator().manual_seed(5)
    coords = torch.rand(n_img * n_ch, 2, generator=g) * torch.tensor([W - 1.0, H - 1.0])
    ref0 = torch.from_numpy(O.get_patch_stack(tmpl, coords.numpy(), H, W)).view(n_img, n_ch, H, W)
    ref = O.avgpool_pyramid(ref0, 6)
    for slot in (0, 0):          # repeated calls return fresh, identical planes
        pyr = ops.tc_rasterize_pyramid(torch.from_numpy(tmpl).cuda(), coords.cuda(), n_img, n_ch, H, W, 6, slot=slot)
        for l, (a, r) in enumerate(zip(pyr, ref)):
            # ONE 8-channel plane per level; a conv sees K_pad = 16 channels (TMA zero-fills the missing plane)
            assert a.C == n_ch and a.C_pad == 8 and a.K_pad == 16 and a.data.shape == (n_img, 1, H >> l, W >> l, 8)
            got = ops.tc_unpack(a).cpu()
            if l == 0:
                assert torch.equal(got, bf16_exact(r))
            else:
                assert rel_err(got.numpy(), r.numpy()) < 2.0 ** -7
            assert n_ch == 8 or float(a.data[:, 0, :, :, n_ch:].abs().max()) == 0.0


@pytest.mark.parametrize('cins,cout,h,w,N', [((32,), 16, 16, 24, 2), ((16,), 8, 13, 13, 3), ((64,), 32, 52, 52, 2),
                                            ((130,), 64, 13, 13, 2), ((32,), 16, 208, 208, 1), ((8, 8), 4, 2, 2, 2),
                                            ((16,), 8, 1, 1, 2)])
def test_tc_upconv_vs_torch(ops, cins, cout, h, w, N):
    """bilinear x2 + conv3x3 as one low-resolution phase conv (+ exact border ring) vs F.interpolate + F.conv2d."""
    torch.manual_seed(6)
    xs = [bf16_exact(torch.relu(torch.randn(N, c, h, w))) for c in cins]
    wgt = torch.randn(cout, sum(cins), 3, 3) * 0.1
    b = torch.randn(cout)
    up = F.interpolate(torch.cat(xs, 1), scale_factor=2, mode='bilinear', align_corners=False)
    ref = F.conv2d(up, wgt, b, padding=1)
    srcs = [ops.tc_pack(x.cuda()) for x in xs]
    w_eff, b_eff = ops.tc_upconv_phase_weights(wgt.cuda(), b.cuda())
    packed = ops.tc_pack_weights(w_eff, list(cins))
    bw = ops.tc_upconv_border_weights(wgt.cuda().contiguous(), list(cins))
    out = ops.tc_upconv3x3(srcs, packed, b_eff, bw, b.cuda(), cout)
    got = ops.tc_unpack(out).cpu()
    assert got.shape == ref.shape
    # bf16 rounding of the folded weights and of the output: 2^-8 relative each
    assert rel_err(got.numpy(), ref.numpy()) < 8e-3
    # the border ring: tensor-core stencil minus the three (five at a corner) outside taps, which the ring-fix kernel
    # evaluates with bf16 operands as well (mma.sync) -> the same error class as the interior
    ring = torch.ones_like(ref, dtype=torch.bool)
    ring[:, :, 2:-2, 2:-2] = False
    assert rel_err(got[ring].numpy(), ref[ring].numpy()) < 8e-3


@pytest.mark.parametrize('cins,cout,h,w,N', [((32,), 16, 16, 24, 2), ((24,), 8, 13, 13, 3), ((64,), 32, 52, 52, 2),
                                            ((128,), 64, 13, 13, 2), ((32,), 16, 208, 208, 1), ((8, 8), 4, 2, 2, 2),
                                            ((16,), 8, 1, 1, 2), ((64,), 32, 5, 40, 2), ((40, 16), 32, 21, 3, 2)])
def test_tc_upconv_padded_ring_mma_and_cuda_core(ops, cins, cout, h, w, N, monkeypatch):
    """Replicate-padded sources: the outermost ring is corrected by upconv_ringfix_mma_kernel (mma.sync, default) or by
    upconv_ringfix_kernel (CUDA cores, YNET_RINGFIX_MMA=0).  Both must match F.interpolate + F.conv2d, ring included,
    and each other up to the bf16 rounding of the interpolated line / the three outside taps."""
    torch.manual_seed(16)
    xs = [bf16_exact(torch.relu(torch.randn(N, c, h, w))) for c in cins]
    wgt = torch.randn(cout, sum(cins), 3, 3) * 0.1
    b = torch.randn(cout)
    up = F.interpolate(torch.cat(xs, 1), scale_factor=2, mode='bilinear', align_corners=False)
    ref = F.conv2d(up, wgt, b, padding=1)
    srcs = [ops.tc_pad_replicate(ops.tc_pack(x.cuda())) for x in xs]
    w_eff, b_eff = ops.tc_upconv_phase_weights(wgt.cuda(), b.cuda())
    packed = ops.tc_pack_weights(w_eff, list(cins))
    bw = ops.tc_upconv_border_weights(wgt.cuda().contiguous(), list(cins))
    got = {}
    for mode in ('1', '0'):
        monkeypatch.setenv('YNET_RINGFIX_MMA', mode)
        got[mode] = ops.tc_unpack(ops.tc_upconv3x3(srcs, packed, b_eff, bw, b.cuda(), cout)).cpu()
        assert got[mode].shape == ref.shape
        assert rel_err(got[mode].numpy(), ref.numpy()) < 8e-3
        ring = torch.ones_like(ref, dtype=torch.bool)
        ring[:, :, 1:-1, 1:-1] = False
        assert rel_err(got[mode][ring].numpy(), ref[ring].numpy()) < 1e-2
        corners = got[mode][:, :, [0, 0, -1, -1], [0, -1, 0, -1]]
        assert rel_err(corners.numpy(), ref[:, :, [0, 0, -1, -1], [0, -1, 0, -1]].numpy()) < 1.5e-2
    inner = (slice(None), slice(None), slice(1, -1), slice(1, -1))
    assert torch.equal(got['0'][inner], got['1'][inner])          # only the ring differs between the two kernels
    assert rel_err(got['1'].numpy(), got['0'].numpy()) < 1e-2    # a bf16 ulp or two of the largest values


@pytest.mark.parametrize('n_wp,H,W', [(2, 64, 96), (1, 32, 32), (2, 416, 416)])
def test_tc_quad_waypoint_planes(ops, n_wp, H, W):
    """2x2-neighbourhood waypoint planes (quad levels): the planes hold map[y+dy][x+dx] (zero outside the image), and a
    conv over them with the four anchored taps equals the nine-tap conv over the plain planes."""
    torch.manual_seed(21)
    n_img, levels = 3, 4
    tmpl = ops.create_dist_template(max(H, W) * 3, 'cuda')
    coords = torch.stack([torch.rand(n_img * n_wp) * (W - 1), torch.rand(n_img * n_wp) * (H - 1)], 1).cuda()
    plain = ops.tc_rasterize_pyramid(tmpl, coords, n_img, n_wp, H, W, levels)
    quad = ops.tc_rasterize_pyramid(tmpl, coords, n_img, n_wp, H, W, levels, quad_levels=2)
    for l in range(levels):
        p, q = plain[l].data[:, 0].float(), quad[l].data[:, 0].float()          # (n_img, h, w, 8)
        if l >= 2:
            assert quad[l].taps == 0 and torch.equal(p, q)
            continue
        assert quad[l].taps == ops.TAPS_QUAD and quad[l].C == 4 * n_wp
        padded = F.pad(p[..., :n_wp], (0, 0, 0, 1, 0, 1))                        # zero row / column beyond the image
        h, w = p.shape[1], p.shape[2]
        for dy in range(2):
            for dx in range(2):
                k = (dy * 2 + dx) * n_wp
                assert torch.equal(q[..., k:k + n_wp], padded[:, dy:dy + h, dx:dx + w])
        assert float(q[..., 4 * n_wp:].abs().max()) == 0.0 if 4 * n_wp < 8 else True
    # conv(cat(x, wp)) through both layouts
    cx, cout = 16, 32
    x = bf16_exact(torch.randn(n_img, cx, H, W))
    wgt = bf16_exact(torch.randn(cout, cx + n_wp, 3, 3) * 0.1).cuda()
    bias = torch.randn(cout).cuda()
    a = ops.tc_pack(x.cuda())
    ref = ops.tc_conv3x3([a, plain[0]], ops.tc_pack_weights(wgt, [cx, n_wp]), bias, cout, True)
    pk = ops.tc_pack_hoisted_weights(wgt, [('conv', (0, cx)), ('quad', (cx, n_wp))])
    got = ops.tc_conv3x3([a, quad[0]], pk, bias, cout, True)
    r, g = ops.tc_unpack(ref).cpu(), ops.tc_unpack(got).cpu()
    wp_plain = plain[0].data[:, 0, :, :, :n_wp].float().permute(0, 3, 1, 2).cpu()
    tref = F.relu(F.conv2d(torch.cat([x, wp_plain], 1), wgt.cpu(), bias.cpu(), padding=1))
    assert rel_err(g.numpy(), tref.numpy()) < 6e-3
    assert rel_err(g.numpy(), r.numpy()) < 4e-3          # same bf16 products, different accumulation order


def test_tc_hoisted_partial_sums_match_direct_conv(ops):
    """Goal-loop hoisting: conv(cat(up, feature, wp)) == conv(cat(up, wp)) + hi/lo partial(feature) via identity taps."""
    torch.manual_seed(11)
    nb, G, H, W = 2, 3, 32, 40
    up = bf16_exact(torch.randn(nb * G, 16, H, W))
    feat = bf16_exact(torch.randn(nb, 32, H, W))
    wp = bf16_exact(torch.rand(nb * G, 2, H, W) * 2)
    w = bf16_exact(torch.randn(32, 50, 3, 3) * 0.1)
    bias = torch.randn(32)
    ref = F.relu(F.conv2d(torch.cat([up, feat.repeat_interleave(G, dim=0), wp], 1), w, bias, padding=1))
    f8 = ops.tc_pack(feat.cuda())
    part = ops.tc_conv3x3_hilo([f8], ops.tc_pack_weights(w[:, 16:48].contiguous().cuda(), [32]), 32)
    assert part.center and part.C_pad == 64
    # the (hi | lo) pair reproduces the fp32 partial sums to ~2^-16 relative
    pf = ops.tc_unpack(part).cpu()
    raw = F.conv2d(feat, w[:, 16:48], None, padding=1)
    assert rel_err((pf[:, :32] + pf[:, 32:]).numpy(), raw.numpy()) < 3e-5
    packed = ops.tc_pack_hoisted_weights(w.cuda(), [('conv', (0, 16)), ('partial', 64), ('conv', (48, 50))])
    bp = torch.zeros(32)
    bp[:32] = bias
    out = ops.tc_conv3x3([ops.tc_pack(up.cuda()), part.repeat_interleave(G), ops.tc_pack(wp.cuda())], packed, bp.cuda(),
                         32, True)
    got = ops.tc_unpack(out).cpu()
    assert rel_err(got.numpy(), ref.numpy()) < 6e-3          # bf16 output rounding only
    # hi-only partial sums: one more bf16 rounding of the feature share
    part1 = ops.tc_conv3x3_hilo([f8], ops.tc_pack_weights(w[:, 16:48].contiguous().cuda(), [32]), 32, with_lo=False)
    assert part1.C_pad == 32
    packed1 = ops.tc_pack_hoisted_weights(w.cuda(), [('conv', (0, 16)), ('partial', 32), ('conv', (48, 50))])
    out1 = ops.tc_conv3x3([ops.tc_pack(up.cuda()), part1.repeat_interleave(G), ops.tc_pack(wp.cuda())], packed1, bp.cuda(),
                          32, True)
    assert rel_err(ops.tc_unpack(out1).cpu().numpy(), ref.numpy()) < 1.2e-2


@pytest.mark.parametrize('level', [0, 1])
def test_tc_im2col_waypoint_source_matches_3x3(ops, level):
    """Waypoint maps written as im2col + a 1x1 centre-tap K block == the 3x3 conv over the plain maps."""
    torch.manual_seed(12)
    n_img, n_wp, H, W = 3, 2, 64, 96
    tmpl = ops.create_dist_template(300, 'cuda')
    coords = torch.tensor([[10.2, 5.5], [80.0, 60.0], [0.0, 0.0], [95.0, 63.0], [40.5, 31.5], [47.0, 2.0]], device='cuda')
    pyr = ops.tc_rasterize_pyramid(tmpl, coords, n_img, n_wp, H, W, 2, slot=77)
    plain = pyr[level]
    i2c = ops.tc_rasterize_im2col(tmpl, coords, n_img, n_wp, H, W, level)
    assert i2c.center and i2c.C == 18 and i2c.H == H >> level
    # im2col channel (c, kh, kw) is the plain map shifted by (kh - 1, kw - 1) with zero fill
    pm = ops.tc_unpack(plain).cpu()
    im = ops.tc_unpack(i2c).cpu()
    ref = F.unfold(pm, 3, padding=1).view(n_img, n_wp * 9, H >> level, W >> level)
    assert torch.equal(im, ref)
    w = bf16_exact(torch.randn(32, 18, 3, 3) * 0.1)
    bias = torch.randn(32)
    up = ops.tc_pack(bf16_exact(torch.randn(n_img, 16, H >> level, W >> level)).cuda())
    direct = ops.tc_conv3x3([up, plain], ops.tc_pack_weights(w.cuda(), [16, 2]), bias.cuda(), 32, True)
    packed = ops.tc_pack_hoisted_weights(w.cuda(), [('conv', (0, 16)), ('i2c', (16, 18))])
    fused = ops.tc_conv3x3([up, i2c], packed, bias.cuda(), 32, True)
    assert rel_err(ops.tc_unpack(fused).cpu().numpy(), ops.tc_unpack(direct).cpu().numpy()) < 4e-3


@pytest.mark.parametrize('cin,cmid,cout,H,W,N', [(32, 32, 30, 64, 96, 3), (32, 32, 12, 416, 416, 2), (16, 16, 6, 48, 40, 5)])
def test_tc_fused_conv_predictor_softargmax(ops, cin, cmid, cout, H, W, N):
    """decoder.4.2 + predictor + SoftArgmax2D in one kernel == the three separate tensor-core launches."""
    torch.manual_seed(13)
    x = bf16_exact(torch.relu(torch.randn(N, cin, H, W)))
    w = bf16_exact(torch.randn(cmid, cin, 3, 3) * 0.15)
    b = torch.randn(cmid) * 0.1
    wp_ = bf16_exact(torch.randn(cout, cmid, 1, 1) * 0.5)
    bp = torch.randn(cout)
    a = ops.tc_pack(x.cuda())
    packed = ops.tc_pack_weights(w.cuda(), [cin])
    bias = torch.zeros((cmid + 15) // 16 * 16)
    bias[:cmid] = b
    ppacked = ops.tc_pack_weights(wp_.cuda(), [cmid])
    pbias = torch.zeros((cout + 15) // 16 * 16)
    pbias[:cout] = bp
    y = ops.tc_conv3x3([a], packed, bias.cuda(), cmid, True)
    sep = ops.tc_conv1x1_softargmax(y, ppacked, pbias.cuda(), cout).cpu().numpy()
    fused = ops.tc_conv3x3_pred_softargmax([a], packed, bias.cuda(), cmid, True, ppacked, pbias.cuda(), cout).cpu().numpy()
    np.testing.assert_allclose(fused, sep, rtol=0, atol=2e-3)       # identical bf16 intermediate, different sum order
    mid = bf16_exact(F.relu(F.conv2d(x, w, b, padding=1)))
    ref = O.softargmax2d(F.conv2d(mid, wp_, bp)).numpy()
    np.testing.assert_allclose(fused, ref, rtol=0, atol=3e-2)


def test_tc_conv_padded_output_and_exact_upconv_ring(ops):
    """A conv can write the replicate-padded layout directly; the phase-decomposed upconv on it equals
    F.interpolate(bilinear x2) + conv3x3 including the border ring."""
    torch.manual_seed(14)
    N, cin, cmid, cout, h, w = 3, 16, 64, 32, 13, 26
    x = bf16_exact(torch.randn(N, cin, h, w))
    w1 = bf16_exact(torch.randn(cmid, cin, 3, 3) * 0.2)
    b1 = torch.randn(cmid) * 0.1
    a = ops.tc_pack(x.cuda())
    packed = ops.tc_pack_weights(w1.cuda(), [cin])
    bias = torch.zeros(cmid)
    bias[:cmid] = b1
    plain = ops.tc_conv3x3([a], packed, bias.cuda(), cmid, True)
    padded = ops.tc_conv3x3([a], packed, bias.cuda(), cmid, True, pad_out=True)
    assert padded.pad == 1 and padded.H == h and padded.data.shape[2] == h + 2
    assert torch.equal(padded.data, ops.tc_pad_replicate(plain).data)
    pl = plain.data.permute(0, 1, 4, 2, 3).float().reshape(N, -1, h, w)                 # (N, C_pad, h, w)
    ref_pad = F.pad(pl, (1, 1, 1, 1), mode='replicate').reshape(N, -1, 8, h + 2, w + 2).permute(0, 1, 3, 4, 2)
    assert torch.equal(padded.data, ref_pad.to(torch.bfloat16))
    w2 = bf16_exact(torch.randn(cout, cmid, 3, 3) * 0.1).contiguous()
    b2 = torch.randn(cout)
    w_eff, b_eff = ops.tc_upconv_phase_weights(w2.cuda(), b2.cuda())
    pk = ops.tc_pack_weights(w_eff, [cmid])
    bw = ops.tc_upconv_border_weights(w2.cuda(), [cmid])
    got = ops.tc_unpack(ops.tc_upconv3x3([padded], pk, b_eff, bw, b2.cuda(), cout)).cpu()
    mid = ops.tc_unpack(plain).cpu()
    ref = F.conv2d(F.interpolate(mid, scale_factor=2, mode='bilinear', align_corners=False), w2, b2, padding=1)
    err = (got - ref).abs()
    scale = ref.abs().max()
    assert err.max() / scale < 1.2e-2
    ring = torch.ones_like(err, dtype=torch.bool)
    ring[:, :, 1:-1, 1:-1] = False
    assert err[ring].max() / scale < 1.2e-2 and err[ring].mean() < 2 * err[~ring].mean() + 1e-3
