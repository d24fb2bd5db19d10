"""Randomised shapes through the row kernels (rowconv_tc.cu) against their slower equivalents: the plain row conv against
the tile kernel (same operands, fp32 accumulation: <= 1 bf16 ulp), the template-fed waypoint source against the rasterised
planes (bit-exact), the two-conv block against the two launches (bit-exact).  Seeds are fixed: the sweep is deterministic.
"""
import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(cuda_device):
    from motion_style_transfer_b200 import ops as _ops
    if not _ops.tc_supported():
        pytest.fail('tensor-core engine unavailable on this device (needs sm_100 + cuTensorMapEncodeTiled)')
    return _ops


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _b32(b):
    out = torch.zeros(32)
    out[:b.numel()] = b
    return out.cuda()


@pytest.mark.parametrize('seed', range(12))
def test_plain_rowconv_random_shapes(ops, seed):
    rng = np.random.RandomState(100 + seed)
    N, H, W = int(rng.randint(1, 5)), int(rng.randint(2, 70)), int(rng.randint(1, 300))
    cin, cout = int(rng.choice([8, 16, 24, 32, 40, 64])), int(rng.randint(1, 33))
    relu, pad_out = bool(rng.randint(2)), bool(rng.randint(2))
    torch.manual_seed(seed)
    x = _bf(torch.randn(N, cin, H, W))
    w = _bf(torch.randn(cout, cin, 3, 3) * 0.1)
    b = torch.randn(cout)
    a = ops.tc_pack(x.cuda())
    got = ops.tc_rowconv3x3(a, ops.tc_rowconv_pack_weights(w.cuda(), a.K_pad), _b32(b), cout, relu, pad_out=pad_out)
    bias_t = torch.zeros((cout + 15) // 16 * 16)
    bias_t[:cout] = b
    ref = ops.tc_conv3x3([a], ops.tc_pack_weights(w.cuda(), [cin]), bias_t.cuda(), cout, relu, pad_out=pad_out)
    torch.cuda.synchronize()
    assert got.data.shape == ref.data.shape, (N, H, W, cin, cout)
    assert rel_err(ops.tc_unpack(got).cpu().numpy(), ops.tc_unpack(ref).cpu().numpy()) < 2.0 ** -7, (N, H, W, cin, cout)
    if pad_out:      # the replicated ring too
        assert rel_err(got.data.float().cpu().numpy(), ref.data.float().cpu().numpy()) < 2.0 ** -7


@pytest.mark.parametrize('seed', range(10))
def test_waypoint_source_and_two_conv_random_shapes(ops, seed):
    rng = np.random.RandomState(200 + seed)
    level = int(rng.randint(2))
    H0, W0 = 32 * int(rng.randint(1, 5)) << level, 32 * int(rng.randint(4, 14)) << level
    G = int(rng.choice([1, 2, 3]))
    N = G * int(rng.randint(1, 3))
    n_wp, c_up = int(rng.randint(1, 3)), int(rng.choice([16, 32]))
    cout = int(rng.randint(1, 33))
    H, W = H0 >> level, W0 >> level
    torch.manual_seed(seed)
    tmpl = ops.create_dist_template(2 * max(H0, W0) + 64, 'cuda')
    coords = torch.stack([torch.rand(N * n_wp) * (W0 - 1), torch.rand(N * n_wp) * (H0 - 1)], 1).cuda().contiguous()
    planes = ops.tc_rasterize_pyramid(tmpl, coords, N, n_wp, H0, W0, level + 1)[level]
    lazy = ops.tc_rasterize_pyramid(tmpl, coords, N, n_wp, H0, W0, level + 1, lazy_levels=level + 1)[level]
    up = ops.tc_pack(_bf(torch.randn(N, c_up, H, W)).cuda())
    feat = ops.tc_pack(_bf(torch.relu(torch.randn(N // G, 32, H, W))).cuda())
    wa = _bf(torch.randn(32, c_up + 32 + n_wp, 3, 3) * 0.1)
    wb = _bf(torch.randn(cout, 32, 3, 3) * 0.1)
    ba, bb = torch.randn(32) * 0.1, torch.randn(cout) * 0.1
    part = ops.tc_conv3x3_hilo([feat], ops.tc_pack_weights(wa[:, c_up:c_up + 32].contiguous().cuda(), [32]), 32,
                               False).repeat_interleave(G)
    p_planes = ops.tc_rowconv_pack_weights_cat(wa.cuda(), [(0, c_up, c_up), (c_up + 32, c_up + 32 + n_wp, 16)])
    p_lazy = ops.tc_rowconv_pack_weights_cat(wa.cuda(), [(0, c_up, c_up)] + lazy.weight_parts(c_up + 32))
    ref_mid = ops.tc_rowconv3x3([up, planes], p_planes, _b32(ba), 32, True, partial=part)
    mid = ops.tc_rowconv3x3([up, lazy], p_lazy, _b32(ba), 32, True, partial=part)
    torch.cuda.synchronize()
    assert torch.equal(mid.data, ref_mid.data), (level, H0, W0, N, n_wp, c_up)
    pb = ops.tc_rowconv_pack_weights(wb.cuda(), 32)
    relu, pad_out = bool(rng.randint(2)), bool(rng.randint(2))
    ref = ops.tc_rowconv3x3(mid, pb, _b32(bb), cout, relu, pad_out=pad_out)
    got = ops.tc_rowconv2_wp([up, lazy], p_lazy, _b32(ba), pb, _b32(bb), cout, relu, pad_out=pad_out, partial=part)
    torch.cuda.synchronize()
    assert torch.equal(got.data, ref.data), (level, H0, W0, N, n_wp, c_up, cout, relu, pad_out)


@pytest.mark.parametrize('seed', range(10))
def test_split_conv_random_shapes(ops, seed):
    """The split-bf16 conv (bf16x3 engine) against float64 torch: random part counts / channels / sizes / batch sharing."""
    import torch.nn.functional as F
    from motion_style_transfer_b200.engine import YNetEngineSplit
    rng = np.random.RandomState(300 + seed)
    n_parts = int(rng.randint(1, 4))
    B = int(rng.randint(1, 4))
    G = int(rng.randint(1, 4))
    H, W = int(rng.randint(1, 40)), int(rng.randint(1, 70))
    cins = [int(rng.randint(1, 70)) for _ in range(n_parts)]
    cout = int(rng.randint(1, 130))
    relu = bool(rng.randint(2))
    torch.manual_seed(seed)
    # part 0 covers all G * B images; later parts may be per-agent (B images, read modulo) or broadcast (1 image)
    batches = [G * B] + [int(rng.choice([G * B, B, 1])) for _ in range(n_parts - 1)]
    xs = [torch.randn(n, c, H, W) for n, c in zip(batches, cins)]
    w = torch.randn(cout, sum(cins), 3, 3) * 0.1
    b = torch.randn(cout)
    full = [x if x.shape[0] == G * B else x.repeat(G * B // x.shape[0], 1, 1, 1) for x in xs]
    ref = F.conv2d(torch.cat(full, 1).double(), w.double(), b.double(), padding=1)
    ref = F.relu(ref) if relu else ref
    parts = [ops.split_pack(x.cuda()) for x in xs]
    sources, ranges = YNetEngineSplit._group(parts)
    idx = torch.cat([torch.arange(c0, c1) for c0, c1 in ranges])
    packed = ops.split_pack_weights(w[:, idx].contiguous().cuda(), [s.layout for s in sources])
    bias = torch.zeros((cout + 15) // 16 * 16)
    bias[:cout] = b
    out = ops.tc_conv3x3_split(sources, packed, bias.cuda(), cout, relu)
    torch.cuda.synchronize()
    got = ops.split_unpack(out).cpu()
    assert got.shape == ref.shape, (batches, cins, cout, H, W)
    assert rel_err(got.numpy(), ref.numpy()) < 3e-5, (batches, cins, cout, H, W)
