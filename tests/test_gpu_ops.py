"""GPU parity tests of the bandwidth-bound ops, through the C ABI (ops.py -> libynet_b200.so).

Compared against (1) the committed golden fixtures written by the live reference and (2) the oracle
restatement on seeded inputs.  Integer / index results must be bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import rel_err
from oracle import ynet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(cuda_device):
    from motion_style_transfer_b200 import ops as _ops
    return _ops


def dev(a, dtype=None):
    t = torch.as_tensor(np.asarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ------------------------------------------------------------------------------------------ a1 / a3 / a9
@pytest.mark.parametrize('size', [1050, 1386, 131])
def test_dist_template_bit_exact(ops, size):
    t = ops.create_dist_template(size, 'cuda').cpu().numpy()
    assert np.array_equal(t, O.create_dist_mat(size).astype(np.float32))


def test_get_patch_golden_bit_exact(ops):
    g = load_golden('get_patch')
    out = ops.rasterize_patches(dev(g['template']), dev(g['traj']), int(g['H']), int(g['W']), check_bounds=True)
    assert np.array_equal(out.cpu().numpy(), g['out'])
    ana = ops.rasterize_dist_analytic(130, dev(g['traj']), int(g['H']), int(g['W']))
    assert np.array_equal(ana.cpu().numpy(), g['out'])


def test_get_patch_full_size_vs_oracle(ops):
    rs = np.random.RandomState(0)
    for size, H, W in ((1050, 416, 416), (1386, 416, 448)):
        tmpl = O.create_dist_mat(size).astype(np.float32)
        traj = np.concatenate([rs.uniform(0, W, (37, 1)), rs.uniform(0, H, (37, 1))], 1).astype(np.float32)
        traj[:4] = [[0.5, 1.5], [2.5, 3.5], [W - 0.51, H - 0.51], [0, 0]]       # half-even ties, corners
        out = ops.rasterize_patches(dev(tmpl), dev(traj), H, W, check_bounds=True).cpu().numpy()
        assert np.array_equal(out, O.get_patch_stack(tmpl, traj, H, W))
        ana = ops.rasterize_dist_analytic(size, dev(traj), H, W).cpu().numpy()
        assert np.array_equal(ana, out)


def test_get_patch_gaussian_template_and_odd_width(ops):
    tmpl = O.create_gaussian_heatmap_template(260, 31, 4, normalize=False).astype(np.float32)
    traj = np.array([[10.2, 7.7], [50.5, 31.5], [0, 0]], dtype=np.float32)
    out = ops.rasterize_patches(dev(tmpl), dev(traj), 64, 99).cpu().numpy()      # W % 4 != 0 -> scalar path
    assert np.array_equal(out, O.get_patch_stack(tmpl, traj, 64, 99))


def test_get_patch_out_of_template_raises(ops):
    tmpl = O.create_dist_mat(100).astype(np.float32)
    with pytest.raises(ValueError):
        ops.rasterize_patches(dev(tmpl), dev(np.array([[90., 5.]], dtype=np.float32)), 64, 64, check_bounds=True)


def test_empty_inputs(ops):
    tmpl = dev(O.create_dist_mat(100).astype(np.float32))
    out = ops.rasterize_patches(tmpl, torch.zeros(0, 2, device='cuda'), 32, 32)
    assert out.shape == (0, 32, 32)


def test_avgpool_pyramid(ops):
    torch.manual_seed(0)
    x = torch.rand(3, 2, 64, 96) * 2
    ref = O.avgpool_pyramid(x, 6)
    got = ops.avgpool_pyramid(x.cuda(), 6)
    assert len(got) == 6
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=5e-6, atol=1e-7)   # summation order only
    got3 = ops.avgpool_pyramid(x.cuda(), 3)
    assert len(got3) == 3 and torch.allclose(got3[2].cpu(), ref[2], atol=1e-6)


# ------------------------------------------------------------------------------------------ a10 / a12 / a13
def test_softargmax_golden(ops):
    g = load_golden('softargmax')
    out = ops.softargmax2d(dev(g['x'])).cpu().numpy()
    np.testing.assert_allclose(out, g['out'], rtol=0, atol=2e-3)
    one = ops.softargmax2d(dev(g['x']), channel=-1).cpu().numpy()
    np.testing.assert_allclose(one[:, 0], g['out'][:, -1], rtol=0, atol=2e-3)
    sm = ops.spatial_softmax(dev(g['x'])).cpu().numpy()
    np.testing.assert_allclose(sm, g['softmax'], rtol=2e-4, atol=1e-9)
    ex = ops.expectation2d(dev(g['softmax'])).cpu().numpy()
    np.testing.assert_allclose(ex, g['on_softmax_map'], rtol=0, atol=2e-3)


@pytest.mark.parametrize('shape', [(2, 12, 416, 416), (5, 3, 64, 100), (1, 1, 33, 7)])
def test_softargmax_vs_oracle_peaky_and_flat(ops, shape):
    torch.manual_seed(1)
    for scale in (0.2, 30.0):                     # random-init-like flat logits and peaky ones
        x = torch.randn(*shape) * scale
        ref = O.softargmax2d(x).numpy()
        got = ops.softargmax2d(x.cuda()).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=0, atol=5e-3)     # pixels; north_star bound is 0.05 px


def test_softargmax_properties_full_size(ops):
    torch.manual_seed(2)
    x = (torch.randn(4, 30, 416, 416) * 5).cuda()
    a = ops.softargmax2d(x)
    b = ops.softargmax2d(x + 7.5)                 # shift invariance of softmax
    assert (a - b).abs().max().item() < 1e-2
    assert a[..., 0].min() >= 0 and a[..., 0].max() <= 415 and a[..., 1].min() >= 0 and a[..., 1].max() <= 415
    delta = torch.full((1, 1, 416, 416), -1e4, device='cuda')
    delta[0, 0, 123, 321] = 50.0                  # one-hot heat map -> its coordinate (x=321, y=123)
    c = ops.softargmax2d(delta)[0, 0].cpu().numpy()
    np.testing.assert_allclose(c, [321, 123], atol=1e-3)


def test_sigmoid_select(ops):
    torch.manual_seed(3)
    x = torch.randn(3, 30, 32, 64) * 4
    got = ops.sigmoid_select(x.cuda(), [14, 29], 1.8).cpu()
    ref = torch.sigmoid(x[:, [14, 29]] / 1.8)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
    got2 = ops.sigmoid_select(x.cuda(), [-1], 1.0).cpu()
    assert torch.allclose(got2, torch.sigmoid(x[:, -1:]), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------ a11
def test_sampling_golden_bit_exact(ops):
    g = load_golden('sampling')
    p = dev(g['prob'])
    idx, xy = ops.multinomial_replacement(p, dev(g['repl_uniforms']), 0.3)
    assert np.array_equal(xy.cpu().numpy(), g['repl_out'])
    idx, xy = ops.multinomial_replacement(p, dev(g['repl_nothr_uniforms']), None)
    assert np.array_equal(xy.cpu().numpy(), g['repl_nothr_out'])
    idx, xy = ops.multinomial_topk(p, dev(g['norepl_expo']), 20, None)
    assert np.array_equal(xy.cpu().numpy(), g['norepl_out'])
    idx, xy = ops.multinomial_topk(p, dev(g['one_expo']), 1, 0.05)
    assert np.array_equal(xy.cpu().numpy(), g['one_out'])


@pytest.mark.parametrize('H,W,B,n,thr', [(416, 416, 3, 10000, 0.002), (416, 416, 2, 10000, 0.01),
                                          (64, 100, 4, 777, None), (33, 7, 2, 50, 0.5)])
def test_multinomial_replacement_vs_oracle_bit_exact(ops, H, W, B, n, thr):
    torch.manual_seed(10)
    logits = torch.randn(B, 1, H, W) * 3
    logits[:, :, H // 3, W // 2] += 8             # a peak so that the threshold bites
    p = torch.sigmoid(logits)
    u = torch.rand(B, n, dtype=torch.float64)
    ref = O.sampling(p.numpy(), n, thr, True, u.numpy())
    idx, xy = ops.multinomial_replacement(p.cuda(), u.cuda(), thr)
    assert np.array_equal(xy.cpu().numpy(), ref)
    assert idx.min() >= 0 and idx.max() < H * W


@pytest.mark.parametrize('H,W,B,n,thr', [(416, 416, 3, 20, None), (64, 96, 5, 20, None), (32, 48, 3, 1, 0.05)])
def test_multinomial_topk_vs_oracle_bit_exact(ops, H, W, B, n, thr):
    torch.manual_seed(11)
    p = torch.sigmoid(torch.randn(B, 1, H, W) * 2)
    q = torch.empty(B, H * W).exponential_(1)
    ref = O.sampling(p.numpy(), n, thr, False, q.numpy())
    idx, xy = ops.multinomial_topk(p.cuda(), q.cuda(), n, thr)
    assert np.array_equal(xy.cpu().numpy(), ref)


def test_sampling_dropin_replays_global_rng(ops):
    """sampling() draws from torch's global generator like torch.multinomial on CPU would."""
    from motion_style_transfer_b200.utils.image_utils import sampling
    torch.manual_seed(5)
    p = torch.sigmoid(torch.randn(2, 1, 32, 48) * 3)
    torch.manual_seed(99)
    ref = O.sampling(p.numpy(), 300, 0.01, True, O.HostRng.uniforms(2, 300))
    torch.manual_seed(99)
    got = sampling(p.cuda(), 300, rel_threshold=0.01, replacement=True).cpu().numpy()
    assert np.array_equal(got, ref)


def test_device_rng(ops):
    u = ops.rng_uniform_f64(7, 0, 100001, 'cuda')
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean().item() - 0.5) < 0.01
    u2 = ops.rng_uniform_f64(7, 0, 100001, 'cuda')
    assert torch.equal(u, u2)
    e = ops.rng_exponential_f32(7, 5, 200003, 'cuda')
    assert e.min() > 0 and abs(e.mean().item() - 1.0) < 0.02


# ------------------------------------------------------------------------------------------ a14
def test_kmeans_golden_bit_exact(ops):
    g = load_golden('kmeans')
    X = dev(g['X'])[None]
    c, a, it, st = ops.kmeans_batched(X, dev(g['init'], torch.int32)[None], None, 0.001, 1000, want_assign=True)
    assert np.array_equal(c[0].cpu().numpy(), g['centres'])
    assert np.array_equal(a[0].cpu().numpy(), g['ids'])
    # empty-cluster reseed path (utils/kmeans.py:82-83) with the recorded torch.randint stream
    X2 = dev(g['X2'])[None]
    c2, a2, it2, st2 = ops.kmeans_batched(X2, dev(g['init2'], torch.int32)[None], dev(g['reseeds2'], torch.int32)[None],
                                          0.001, 1000, want_assign=True)
    assert np.array_equal(c2[0].cpu().numpy(), g['centres2'])
    assert np.array_equal(a2[0].cpu().numpy(), g['ids2'])
    assert int(st2[0]) == 0


def test_kmeans_batched_vs_oracle_bit_exact(ops):
    rs = np.random.RandomState(3)
    B, N, K = 6, 10000, 19
    X = np.stack([np.stack([rs.randint(0, 416, N), rs.randint(0, 416, N)], 1) for _ in range(B)]).astype(np.float32)
    # make it multi-modal like TTST samples
    X[:, : N // 2] = (X[:, : N // 2] * 0.1 + 150).round()
    init = np.stack([rs.choice(N, K, replace=False) for _ in range(B)]).astype(np.int32)
    reseeds = rs.randint(0, N, (B, 16)).astype(np.int32)
    c, a, it, st = ops.kmeans_batched(dev(X), dev(init), dev(reseeds), 0.001, 1000, want_assign=True)
    for b in range(B):
        stream = iter(reseeds[b])
        ids, cen, n_it = O.kmeans(X[b], K, init[b], reseed_fn=lambda: int(next(stream)), tol=0.001, iter_limit=1000)
        assert np.array_equal(c[b].cpu().numpy(), cen), f'agent {b}'
        assert np.array_equal(a[b].cpu().numpy(), ids)
        assert int(it[b]) == n_it


def test_kmeans_iter_limit_and_dropin(ops):
    from motion_style_transfer_b200.utils.kmeans import kmeans
    rs = np.random.RandomState(4)
    X = rs.randint(0, 64, (500, 2)).astype(np.float32)
    np.random.seed(1)
    ids, cen = kmeans(X=dev(X), num_clusters=4, distance='euclidean', device=torch.device('cuda'), tqdm_flag=False,
                      tol=0.001, iter_limit=3)
    np.random.seed(1)
    init = O.HostRng.kmeans_init(500, 4)
    ids_o, cen_o, n_it = O.kmeans(X, 4, init, tol=0.001, iter_limit=3)
    assert n_it <= 3 and np.array_equal(cen.cpu().numpy(), cen_o) and np.array_equal(ids.cpu().numpy(), ids_o)
    with pytest.raises(NotImplementedError):
        kmeans(X=dev(X), num_clusters=4, distance='cosine')


# ------------------------------------------------------------------------------------------ a16 / a17
def test_cws_gaussian_golden(ops):
    from motion_style_transfer_b200.utils.evaluate import torch_multivariate_gaussian_heatmap as gauss
    g = load_golden('cws_gaussian')
    a = gauss([40.3, 20.2], 32, 48, [13.0, -7.5], 6, 2, 'cuda', True).cpu().numpy()
    b = gauss([10.0, 30.0], 32, 48, [-3.0, 4.0], 5, 2, 'cuda', False).cpu().numpy()
    assert rel_err(a, g['g1']) < 1e-3 and rel_err(b, g['g2']) < 1e-3


def test_cws_waypoints_vs_oracle(ops):
    torch.manual_seed(6)
    B, H, W, n_goal = 3, 64, 96, 5
    sig = torch.sigmoid(torch.randn(B, 2, H, W) * 2)
    goals = torch.stack([torch.rand(n_goal, B) * W, torch.rand(n_goal, B) * H], -1)
    last = torch.stack([torch.rand(B) * W, torch.rand(B) * H], -1)
    ref = O.cws_waypoints(sig, goals, last, n_goal, 6, 2, True)               # (G, B, 2, 2)
    sf = torch.full((n_goal,), 6.0)
    got = ops.cws_waypoint(sig[:, 0].contiguous().cuda(), goals.cuda(), last.cuda(), 0.5, sf.cuda(), 2, True)
    np.testing.assert_allclose(got.cpu().numpy(), ref[:, :, 0].numpy(), rtol=0, atol=5e-3)


def test_ade_fde_vs_oracle(ops):
    torch.manual_seed(7)
    K, B, T, n_wp = 20, 7, 12, 2
    gt = torch.rand(B, T, 2) * 100
    trajs = torch.rand(K, B, T, 2) * 100
    wps = torch.rand(K, B, n_wp, 2) * 100
    a, f = O.ade_fde(gt, trajs, wps, 0.25)
    ga, gf = ops.ade_fde(gt.cuda(), trajs.cuda(), wps.cuda(), 0.25)
    np.testing.assert_allclose(ga.cpu().numpy(), a.numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(gf.cpu().numpy(), f.numpy(), rtol=1e-5, atol=1e-4)
