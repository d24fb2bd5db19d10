"""GPU tests of the row-marching, kh-stacked conv kernel (csrc/rowconv_tc.cu) through the C ABI: the plain conv against
F.conv2d and against the tile kernel (ynet_tc_conv3x3), and the fused conv + predictor + SoftArgmax2D tail
(ynet.py:468-469 + 582-583, softargmax.py:55-81) against a float32 torch restatement and the unfused launches.

STATED TOLERANCE: bf16 operands, fp32 accumulation: conv output <= 5e-3 of max|ref| (bf16 output rounding); soft-argmax
coordinates <= 2e-2 px against float32 logits computed from the bf16-rounded activation.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err
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


def _bias32(b):
    out = torch.zeros(32)
    out[:b.numel()] = b
    return out.cuda()


@pytest.mark.parametrize('cin,cout,H,W,N,relu', [
    (32, 32, 16, 16, 1, False),       # one strip, 16 of 126 lanes live
    (32, 32, 13, 13, 2, True),        # any width
    (32, 32, 30, 127, 2, True),       # 126 + 1 pixels
    (32, 32, 64, 64, 2, True),
    (16, 32, 48, 32, 3, True),        # one K block
    (64, 32, 40, 160, 2, True),       # four K blocks, two strips (126 + 34 pixels)
    (32, 16, 33, 48, 2, False),       # odd height, C_out 16
    (24, 30, 21, 272, 2, True),       # padded channels on both sides, three strips
    (32, 32, 2, 16, 5, True),         # minimal height: both rows are border rows
    (32, 32, 416, 416, 1, True),      # decoder.4.2
])
def test_rowconv3x3_vs_torch(ops, cin, cout, H, W, N, relu):
    torch.manual_seed(1)
    x = bf16_exact(torch.randn(N, cin, H, W))
    w = bf16_exact(torch.randn(cout, cin, 3, 3) * 0.1)
    b = torch.randn(cout)
    ref = F.conv2d(x, w, b, padding=1)
    ref = F.relu(ref) if relu else ref
    a = ops.tc_pack(x.cuda())
    assert ops.tc_rowconv_supported(a, cout)
    packed = ops.tc_rowconv_pack_weights(w.cuda(), a.K_pad)
    out = ops.tc_rowconv3x3(a, packed, _bias32(b), cout, relu)
    torch.cuda.synchronize()
    got = ops.tc_unpack(out).cpu()
    assert got.shape == ref.shape
    assert rel_err(got.numpy(), ref.numpy()) < 5e-3
    # same operands, same fp32 accumulation up to the summation order: the tile kernel agrees to a bf16 ulp
    packed_t = ops.tc_pack_weights(w.cuda(), [cin])
    bias_t = torch.zeros((cout + 15) // 16 * 16)
    bias_t[:cout] = b
    tile = ops.tc_unpack(ops.tc_conv3x3([a], packed_t, bias_t.cuda(), cout, relu)).cpu()
    assert rel_err(got.numpy(), tile.numpy()) < 2.0 ** -7
    # a second launch reuses a clean accumulator ring (the epilogue hands every slot back zeroed)
    out2 = ops.tc_rowconv3x3(a, packed, _bias32(b), cout, relu)
    assert torch.equal(out2.data, out.data)


def test_rowconv3x3_padded_output_and_repeated_source(ops):
    """pad_out: the replicate-padded layout the phase-decomposed upconv consumes; rep: agent-major stacked source."""
    torch.manual_seed(2)
    nb, G, H, W = 2, 3, 24, 48
    x = bf16_exact(torch.randn(nb, 32, H, W))
    w = bf16_exact(torch.randn(32, 32, 3, 3) * 0.1)
    b = torch.randn(32)
    ref = F.relu(F.conv2d(x, w, b, padding=1))
    a = ops.tc_pack(x.cuda())
    packed = ops.tc_rowconv_pack_weights(w.cuda(), a.K_pad)
    out = ops.tc_rowconv3x3(a, packed, _bias32(b), 32, True, pad_out=True)
    assert out.pad == 1 and out.data.shape == (nb, 4, H + 2, W + 2, 8)
    inner = ops.tc_unpack(out).cpu()
    assert rel_err(inner.numpy(), ref.numpy()) < 5e-3
    full = ops.tc_unpack(ops.C8(out.data, 32)).cpu()              # the (H + 2, W + 2) planes as they are
    assert torch.equal(full, F.pad(inner, (1, 1, 1, 1), mode='replicate'))
    rep = ops.tc_rowconv3x3(a.repeat_interleave(G), packed, _bias32(b), 32, True)
    assert rep.N == nb * G
    assert torch.equal(ops.tc_unpack(rep).cpu(), ops.tc_unpack(out).cpu().repeat_interleave(G, dim=0))


@pytest.mark.parametrize('cin,cpred,H,W,N', [(32, 30, 64, 96, 3), (32, 12, 416, 416, 2), (16, 6, 34, 144, 2),
                                             (32, 30, 208, 208, 5)])
def test_rowconv_pred_softargmax(ops, cin, cpred, H, W, N):
    torch.manual_seed(3)
    x = bf16_exact(torch.relu(torch.randn(N, cin, H, W)))
    w = bf16_exact(torch.randn(32, cin, 3, 3) * 0.1)
    b = torch.randn(32) * 0.1
    wp = bf16_exact(torch.randn(cpred, 32, 1, 1) * 0.5)
    bp = torch.randn(cpred)
    y = bf16_exact(F.relu(F.conv2d(x, w, b, padding=1)))          # the activation as the kernel rounds it
    ref = O.softargmax2d(F.conv2d(y, wp, bp)).numpy()
    a = ops.tc_pack(x.cuda())
    packed = ops.tc_rowconv_pack_weights(w.cuda(), a.K_pad)
    ppacked = ops.tc_pack_weights(wp.cuda(), [32])
    pb = torch.zeros((cpred + 15) // 16 * 16)
    pb[:cpred] = bp
    got = ops.tc_rowconv3x3_pred_softargmax(a, packed, _bias32(b), 32, True, ppacked, pb.cuda(), cpred)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert got.shape == (N, cpred, 2)
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-2)
    # two MMA-issuing warps accumulate into the same TMEM slots: the result must not depend on how their MMAs interleave
    for _ in range(3):
        again = ops.tc_rowconv3x3_pred_softargmax(a, packed, _bias32(b), 32, True, ppacked, pb.cuda(), cpred).cpu().numpy()
        assert np.array_equal(again, got)
    # the unfused launches (tile conv -> predictor + soft-argmax kernel)
    packed_t = ops.tc_pack_weights(w.cuda(), [cin])
    yt = ops.tc_conv3x3([a], packed_t, _bias32(b), 32, True)
    unf = ops.tc_conv1x1_softargmax(yt, ppacked, pb.cuda(), cpred).cpu().numpy()
    np.testing.assert_allclose(got, unf, rtol=0, atol=2e-2)
    # peaky logits (what the trained / x50 model produces)
    ppacked8 = ops.tc_pack_weights((wp * 8).cuda(), [32])
    got8 = ops.tc_rowconv3x3_pred_softargmax(a, packed, _bias32(b), 32, True, ppacked8, (pb * 8).cuda(), cpred).cpu().numpy()
    # (a bf16 rounding boundary of one activation -- fp32 sums in another order than torch's -- moves a logit by
    # 8 x 0.4 %; with two competing peaks that shifts the expectation by ~0.1 px.  The unfused kernels see the same.)
    np.testing.assert_allclose(got8, O.softargmax2d(F.conv2d(y, wp * 8, bp * 8)).numpy(), rtol=0, atol=0.25)
    unf8 = ops.tc_conv1x1_softargmax(yt, ppacked8, (pb * 8).cuda(), cpred).cpu().numpy()
    np.testing.assert_allclose(got8, unf8, rtol=0, atol=0.25)


def test_rowconv_multi_source_with_hoisted_partial(ops):
    """decoder.i.0 of the trajectory decoder as the row kernel runs it: cat(up, feature, waypoints) with the feature
    share hoisted (tc_conv3x3_hilo, once per agent) and added by the epilogue; sources [up, waypoint plane] side by side
    on the K axis; agent-major stacking (the partial is read as image n // G)."""
    torch.manual_seed(4)
    nb, G, H, W = 2, 3, 40, 150
    c_up, c_feat, n_wp = 16, 32, 2
    up = bf16_exact(torch.randn(nb * G, c_up, H, W))
    feat = bf16_exact(torch.relu(torch.randn(nb, c_feat, H, W)))
    wp = bf16_exact(torch.rand(nb * G, n_wp, H, W))
    w = bf16_exact(torch.randn(32, c_up + c_feat + n_wp, 3, 3) * 0.1)
    b = torch.randn(32) * 0.1
    ref = F.relu(F.conv2d(torch.cat([up, feat.repeat_interleave(G, dim=0), wp], 1), w, b, padding=1))
    a_up, a_wp = ops.tc_pack(up.cuda()), ops.tc_pack(wp.cuda())
    f8 = ops.tc_pack(feat.cuda())
    wf = w[:, c_up:c_up + c_feat].contiguous().cuda()
    for with_lo, tol in ((True, 5e-3), (False, 8e-3)):
        part = ops.tc_conv3x3_hilo([f8], ops.tc_pack_weights(wf, [c_feat]), 32, with_lo).repeat_interleave(G)
        packed = ops.tc_rowconv_pack_weights_cat(w.cuda(), [(0, c_up, a_up.K_pad), (c_up + c_feat, c_up + c_feat + n_wp, a_wp.K_pad)])
        out = ops.tc_rowconv3x3([a_up, a_wp], packed, _bias32(b), 32, True, partial=part)
        torch.cuda.synchronize()
        got = ops.tc_unpack(out).cpu()
        assert got.shape == ref.shape
        assert rel_err(got.numpy(), ref.numpy()) < tol
    # three conv sources, no partial: plain concatenation
    x3 = bf16_exact(torch.randn(nb * G, 8, H, W))
    w3 = bf16_exact(torch.randn(24, c_up + 8 + n_wp, 3, 3) * 0.1)
    ref3 = F.conv2d(torch.cat([up, x3, wp], 1), w3, None, padding=1)
    a3 = ops.tc_pack(x3.cuda())
    packed3 = ops.tc_rowconv_pack_weights_cat(w3.cuda(), [(0, 16, 16), (16, 24, a3.K_pad), (24, 26, a_wp.K_pad)])
    out3 = ops.tc_rowconv3x3([a_up, a3, a_wp], packed3, torch.zeros(32).cuda(), 24, False)
    assert rel_err(ops.tc_unpack(out3).cpu().numpy(), ref3.numpy()) < 5e-3


@pytest.mark.parametrize('n_wp,level,H0,W0,N', [(2, 0, 64, 160, 3), (2, 1, 64, 288, 3), (1, 0, 32, 416, 2), (1, 1, 96, 480, 2),
                                                (2, 0, 416, 416, 2), (2, 1, 416, 416, 2), (2, 0, 32, 128, 2), (2, 0, 32, 256, 2), (1, 1, 64, 512, 2)])
def test_rowconv_gathered_waypoint_planes_bit_exact(ops, n_wp, level, H0, W0, N):
    """ynet_tc_rowconv3x3_wp: the waypoint planes (get_patch, image_utils.py:40-63, + AvgPool2d(2), evaluate.py:255-257)
    loaded by the kernel's TMA from the distance template's bf16 planes == the same conv reading the rasterised planes, bit for bit;
    coordinates near the image border put zero-padded / clamped pixels in every strip."""
    torch.manual_seed(5)
    H, W = H0 >> level, W0 >> level
    tmpl = ops.create_dist_template(3 * max(H0, W0), 'cuda')
    coords = torch.stack([torch.rand(N * n_wp) * (W0 - 1), torch.rand(N * n_wp) * (H0 - 1)], 1)
    coords[0] = torch.tensor([0.5, 1.5])                       # half-to-even rounding, window against the template edge
    coords[-1] = torch.tensor([W0 - 1.0, H0 - 1.0])
    coords = coords.cuda().contiguous()
    planes = ops.tc_rasterize_pyramid(tmpl, coords, N, n_wp, H0, W0, level + 1)[level]
    lazy = ops.tc_rasterize_pyramid(tmpl, coords, N, n_wp, H0, W0, level + 1, lazy_levels=level + 1)[level]
    assert isinstance(lazy, ops.WpPlanes) and (lazy.H, lazy.W, lazy.N, lazy.C) == (H, W, N, n_wp)
    # more than two waypoint channels are not gathered in the kernel: the planes are written as before
    c3 = torch.rand(N * 3, 2).cuda() * 20
    assert isinstance(ops.tc_rasterize_pyramid(tmpl, c3, N, 3, H0, W0, 2, lazy_levels=2)[0], ops.C8)
    assert torch.equal(lazy.materialize().data, planes.data)
    up = ops.tc_pack(bf16_exact(torch.randn(N, 16, H, W)).cuda())
    w = bf16_exact(torch.randn(32, 16 + n_wp, 3, 3) * 0.1)
    b = torch.randn(32) * 0.1
    packed = ops.tc_rowconv_pack_weights_cat(w.cuda(), [(0, 16, 16), (16, 16 + n_wp, 16)])
    packed_l = ops.tc_rowconv_pack_weights_cat(w.cuda(), [(0, 16, 16)] + lazy.weight_parts(16))    # channel c at K index 8 c
    ref = ops.tc_rowconv3x3([up, planes], packed, _bias32(b), 32, True)
    got = ops.tc_rowconv3x3([up, lazy], packed_l, _bias32(b), 32, True)
    torch.cuda.synchronize()
    assert torch.equal(got.data, ref.data)
    # against torch on the float32 maps (bf16 operands, fp32 accumulation)
    wmap = ops.tc_unpack(planes).cpu()
    full = F.relu(F.conv2d(torch.cat([ops.tc_unpack(up).cpu(), wmap], 1), w, b, padding=1))
    assert rel_err(ops.tc_unpack(got).cpu().numpy(), full.numpy()) < 5e-3
    # with the hoisted partial sums and a second tensor source in front (decoder.i.0 of the trajectory decoder)
    feat = ops.tc_pack(bf16_exact(torch.relu(torch.randn(N, 32, H, W))).cuda())
    part = ops.tc_conv3x3_hilo([feat], ops.tc_pack_weights(bf16_exact(torch.randn(32, 32, 3, 3) * 0.05).cuda(), [32]), 32, False)
    ref2 = ops.tc_rowconv3x3([up, planes], packed, _bias32(b), 32, True, pad_out=True, partial=part)
    got2 = ops.tc_rowconv3x3([up, lazy], packed_l, _bias32(b), 32, True, pad_out=True, partial=part)
    assert torch.equal(got2.data, ref2.data)


@pytest.mark.parametrize('n_wp,level,H0,W0,N,G,c_up', [(2, 0, 64, 160, 4, 2, 16), (2, 1, 64, 416, 4, 2, 32), (1, 0, 32, 416, 2, 1, 16),
                                                       (2, 0, 416, 416, 2, 2, 16), (2, 0, 64, 256, 6, 3, 16)])
def test_rowconv2_block_equals_unfused_chain(ops, n_wp, level, H0, W0, N, G, c_up):
    """ynet_tc_rowconv2_wp: decoder.i.0 (+ partial sums, ReLU) and decoder.i.2 in one kernel == the two row-kernel launches
    with the activation round-tripping through HBM, bit for bit (same MMAs in the same order per output pixel); plain and
    replicate-padded output; agent-major partial sums (image n reads partial n // G)."""
    torch.manual_seed(6)
    H, W = H0 >> level, W0 >> level
    tmpl = ops.create_dist_template(3 * max(H0, W0), 'cuda')
    coords = torch.stack([torch.rand(N * n_wp) * (W0 - 1), torch.rand(N * n_wp) * (H0 - 1)], 1)
    coords[0] = torch.tensor([0.5, 1.5])
    coords[-1] = torch.tensor([W0 - 1.0, H0 - 1.0])
    coords = coords.cuda().contiguous()
    lazy = ops.tc_rasterize_pyramid(tmpl, coords, N, n_wp, H0, W0, level + 1, lazy_levels=level + 1)[level]
    up = ops.tc_pack(bf16_exact(torch.randn(N, c_up, H, W)).cuda())
    feat = ops.tc_pack(bf16_exact(torch.relu(torch.randn(N // G, 32, H, W))).cuda())
    wa = bf16_exact(torch.randn(32, c_up + 32 + n_wp, 3, 3) * 0.1)
    ba, bb = torch.randn(32) * 0.1, torch.randn(32) * 0.1
    wb = bf16_exact(torch.randn(32, 32, 3, 3) * 0.1)
    part = ops.tc_conv3x3_hilo([feat], ops.tc_pack_weights(wa[:, c_up:c_up + 32].contiguous().cuda(), [32]), 32,
                               False).repeat_interleave(G)
    pa = ops.tc_rowconv_pack_weights_cat(wa.cuda(), [(0, c_up, c_up)] + lazy.weight_parts(c_up + 32))
    pb = ops.tc_rowconv_pack_weights(wb.cuda(), 32)
    assert ops.tc_rowconv2_supported([up, lazy], 32, 32)
    mid = ops.tc_rowconv3x3([up, lazy], pa, _bias32(ba), 32, True, partial=part)
    for pad_out in (False, True):
        ref = ops.tc_rowconv3x3(mid, pb, _bias32(bb), 32, True, pad_out=pad_out)
        got = ops.tc_rowconv2_wp([up, lazy], pa, _bias32(ba), pb, _bias32(bb), 32, True, pad_out=pad_out, partial=part)
        torch.cuda.synchronize()
        assert got.pad == ref.pad and torch.equal(got.data, ref.data)
    # no ReLU behind conv B, fewer output channels, and a second launch on a clean accumulator ring
    wb2 = bf16_exact(torch.randn(20, 32, 3, 3) * 0.1)
    pb2 = ops.tc_rowconv_pack_weights(wb2.cuda(), 32)
    ref2 = ops.tc_rowconv3x3(mid, pb2, _bias32(bb[:20]), 20, False)
    got2 = ops.tc_rowconv2_wp([up, lazy], pa, _bias32(ba), pb2, _bias32(bb[:20]), 20, False, partial=part)
    assert torch.equal(got2.data, ref2.data)
    if c_up > 16:
        return                     # (the tail kernel keeps an 8-row partial-sum ring: <= 32 input channels in conv A)
    # tail: predictor + soft-argmax behind conv B (other strip partition of the sums: float32 order differs)
    wp1 = bf16_exact(torch.randn(30, 32, 1, 1) * 0.5)
    ppk = ops.tc_pack_weights(wp1.cuda(), [32])
    pbias = torch.zeros(32)
    pbias[:30] = torch.randn(30)
    refs = ops.tc_rowconv3x3_pred_softargmax(mid, pb, _bias32(bb), 32, True, ppk, pbias.cuda(), 30)
    gots = ops.tc_rowconv2_wp_pred_softargmax([up, lazy], pa, _bias32(ba), pb, _bias32(bb), True, ppk, pbias.cuda(), 30,
                                              partial=part)
    torch.cuda.synchronize()
    assert gots.shape == (N, 30, 2)
    # (the tail variant adds the conv biases after the accumulation instead of starting from them: a few bf16 roundings of
    # the activations flip by one ulp; stated tolerance of the fused predictor + soft-argmax kernels: 2e-2 px)
    np.testing.assert_allclose(gots.cpu().numpy(), refs.cpu().numpy(), rtol=0, atol=2e-2)
    again = ops.tc_rowconv2_wp_pred_softargmax([up, lazy], pa, _bias32(ba), pb, _bias32(bb), True, ppk, pbias.cuda(), 30,
                                               partial=part)
    assert torch.equal(again, gots)


@pytest.mark.parametrize('n_wp,H0,W0,N,mod', [(1, 32, 128, 4, 2), (2, 32, 128, 6, 3), (2, 64, 352, 4, 0)])
def test_rowconv2_edge_geometry(ops, n_wp, H0, W0, N, mod):
    """Two-conv block at the smallest width (two overlapping edge strips), with one waypoint channel, with a goal-major
    partial-sum source (image n reads partial n % mod), and with an interior strip: always bit-identical to the two
    launches, and within bf16 tolerance of torch on the same operands."""
    torch.manual_seed(7)
    tmpl = ops.create_dist_template(3 * max(H0, W0), 'cuda')
    coords = torch.stack([torch.rand(N * n_wp) * (W0 - 1), torch.rand(N * n_wp) * (H0 - 1)], 1).cuda().contiguous()
    lazy = ops.tc_rasterize_pyramid(tmpl, coords, N, n_wp, H0, W0, 1, lazy_levels=1)[0]
    up = ops.tc_pack(bf16_exact(torch.randn(N, 16, H0, W0)).cuda())
    wa = bf16_exact(torch.randn(32, 16 + 32 + n_wp, 3, 3) * 0.1)
    wb = bf16_exact(torch.randn(32, 32, 3, 3) * 0.1)
    ba, bb = torch.randn(32) * 0.1, torch.randn(32) * 0.1
    n_part = mod if mod else N
    feat = ops.tc_pack(bf16_exact(torch.relu(torch.randn(n_part, 32, H0, W0))).cuda())
    part = ops.tc_conv3x3_hilo([feat], ops.tc_pack_weights(wa[:, 16:48].contiguous().cuda(), [32]), 32, False)
    pa = ops.tc_rowconv_pack_weights_cat(wa.cuda(), [(0, 16, 16)] + lazy.weight_parts(48))
    pb = ops.tc_rowconv_pack_weights(wb.cuda(), 32)
    mid = ops.tc_rowconv3x3([up, lazy], pa, _bias32(ba), 32, True, partial=part)
    ref = ops.tc_rowconv3x3(mid, pb, _bias32(bb), 32, True)
    got = ops.tc_rowconv2_wp([up, lazy], pa, _bias32(ba), pb, _bias32(bb), 32, True, partial=part)
    torch.cuda.synchronize()
    assert torch.equal(got.data, ref.data)
    # against torch on the unpacked operands (bf16 operands, fp32 accumulation, bf16 activation in between)
    wmap = ops.tc_unpack(lazy.materialize()).cpu()
    x = torch.cat([ops.tc_unpack(up).cpu(), ops.tc_unpack(feat).cpu().repeat(N // n_part, 1, 1, 1), wmap], 1)
    y1 = bf16_exact(F.relu(F.conv2d(x, wa, ba, padding=1)))
    y2 = F.relu(F.conv2d(y1, wb, bb, padding=1))
    assert rel_err(ops.tc_unpack(got).cpu().numpy(), y2.numpy()) < 1.5e-2
