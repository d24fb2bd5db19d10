"""The oracle restatement (oracle/ynet_oracle.py) against fixtures produced by the LIVE
reference (oracle/gen_golden.py) -- and, when /root/reference is present, against the
reference itself.  CPU only."""
import ast
import warnings

import numpy as np
import pytest
import torch

from conftest import load_golden, golden_state_dict
from oracle import ynet_oracle as O
from oracle import ref_harness


@pytest.mark.parametrize('size', [1050, 1386])
def test_dist_template(size):
    g = load_golden(f'dist_template_{size}')
    t = O.create_dist_mat(size).astype(np.float32)
    assert np.array_equal(t[::97, ::89], g['sample'])
    assert np.array_equal(t[[0, size // 2, size - 1]], g['rows'])
    assert np.bitwise_xor.reduce(t.view(np.uint32).ravel()) == g['xor']
    assert t.astype(np.float64).sum() == g['sum64']
    assert t.max() == np.float32(2.0) and t[size // 2, size // 2] == 0


def test_gauss_template():
    g = load_golden('gauss_template_1050')
    t = O.create_gaussian_heatmap_template(1050, 31, 4, normalize=False).astype(np.float32)
    assert np.array_equal(t[525 - 16:525 + 16, 525 - 16:525 + 16], g['centre'])
    assert t.astype(np.float64).sum() == g['sum64']


def test_get_patch_bit_exact_and_half_even():
    g = load_golden('get_patch')
    out = O.get_patch_stack(g['template'], g['traj'], int(g['H']), int(g['W']))
    assert np.array_equal(out, g['out'])
    # analytic fp64 form is the same bits
    assert np.array_equal(O.dist_patch_analytic(g['traj'], int(g['H']), int(g['W']), 130), g['out'])
    x, y = O.round_coords(np.array([[0.5, 1.5], [2.5, 101.5]], dtype=np.float32))
    assert list(x) == [0, 2] and list(y) == [2, 102]


def test_sampling_bit_exact():
    g = load_golden('sampling')
    p = g['prob']
    assert np.array_equal(O.sampling(p, 500, 0.3, True, g['repl_uniforms']), g['repl_out'])
    assert np.array_equal(O.sampling(p, 64, None, True, g['repl_nothr_uniforms']), g['repl_nothr_out'])
    assert np.array_equal(O.sampling(p, 20, None, False, g['norepl_expo']), g['norepl_out'])
    assert np.array_equal(O.sampling(p, 1, 0.05, False, g['one_expo']), g['one_out'])


def test_softargmax_softmax():
    g = load_golden('softargmax')
    np.testing.assert_allclose(O.softargmax2d(g['x']).numpy(), g['out'], rtol=0, atol=1e-4)
    np.testing.assert_allclose(O.spatial_softmax(g['x']).numpy(), g['softmax'], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(O.softargmax_on_softmax_map(g['softmax']).numpy(), g['on_softmax_map'],
                               rtol=0, atol=1e-4)


def test_kmeans_bit_exact():
    g = load_golden('kmeans')
    ids, c, it = O.kmeans(g['X'], 7, g['init'], tol=0.001, iter_limit=1000)
    assert np.array_equal(c, g['centres']) and np.array_equal(ids, g['ids'])


def test_kmeans_empty_cluster_reseed():
    g = load_golden('kmeans')
    stream = iter(g['reseeds2'])
    used = []

    def reseed():
        v = int(next(stream))
        used.append(v)
        return v
    ids, c, it = O.kmeans(g['X2'], 5, g['init2'], reseed_fn=reseed, tol=0.001, iter_limit=1000)
    assert len(used) >= 1, 'fixture must exercise utils/kmeans.py:82-83'
    assert np.array_equal(c, g['centres2']) and np.array_equal(ids, g['ids2'])


def test_cws_gaussian():
    g = load_golden('cws_gaussian')
    a = O.cws_gaussian([40.3, 20.2], 32, 48, [13.0, -7.5], 6, 2, True).numpy()
    b = O.cws_gaussian([10.0, 30.0], 32, 48, [-3.0, 4.0], 5, 2, False).numpy()
    np.testing.assert_allclose(a, g['g1'], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(b, g['g2'], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize('tag,network', [('ynet', 'original'), ('ynetmod', 'fusion')])
def test_network_forward(tag, network):
    g = load_golden(f'network_{tag}')
    sd = golden_state_dict(g)
    torch.set_num_threads(1)
    scene = torch.from_numpy(g['scene']).expand(2, -1, -1, -1)
    feats = O.pred_features(sd, scene, g['motion'], network)
    for i, f in enumerate(feats):
        np.testing.assert_allclose(f.numpy(), g[f'feat{i}'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.pred_goal(sd, feats).numpy(), g['goal'], rtol=1e-5, atol=1e-6)
    pyr = O.avgpool_pyramid(g['wp'], len(feats))
    traj = O.pred_traj(sd, [torch.cat([f, p], 1) for f, p in zip(feats, pyr)])
    np.testing.assert_allclose(traj.numpy(), g['traj'], rtol=1e-5, atol=1e-6)


# ---- serial / parallel adapter baselines (ynet.py:15-131, 237-283) -----------------------------------
from helpers import ADAPTER_TAGS, build_adapter_model     # noqa: E402


@pytest.mark.parametrize('tag', ADAPTER_TAGS)
def test_adapter_baselines_oracle_against_reference_fixture(tag):
    """The oracle's eval-mode restatement of AdapterLayer / AdapterBlock vs the live reference's features and goal logits."""
    g = load_golden(f'adapter_{tag}')
    torch.set_num_threads(1)
    m = build_adapter_model(g)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    scene = torch.from_numpy(g['scene']).expand(2, -1, -1, -1)
    feats = O.pred_features(sd, scene, g['motion'], 'original', adapter_position=g['position'])
    for i, f in enumerate(feats):
        np.testing.assert_allclose(f.numpy(), g[f'feat{i}'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.pred_goal(sd, feats).numpy(), g['goal'], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize('tag', ADAPTER_TAGS)
def test_adapter_folding_is_the_adapter_forward(tag):
    """Host logic of the engines (weight-space folding, engine.fold_adapter_layer / block_adapter_weights): one 3x3 conv
    with the folded weights equals the adapter's own forward as restated by the oracle."""
    import torch.nn.functional as F
    from motion_style_transfer_b200 import engine as E
    g = load_golden(f'adapter_{tag}')
    m = build_adapter_model(g).double()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    torch.manual_seed(0)
    if 'Layer' in str(g['train_net']):
        stage = int(g['position'][-1])
        conv = m.encoder.stages[stage][1 if stage else 0]
        prefix = f'encoder.stages.{stage}.{1 if stage else 0}'
        x = torch.randn(2, conv.weight.shape[1], 12, 10, dtype=torch.float64)
        w, b = E.fold_adapter_layer(conv, conv.weight.detach(), conv.bias.detach())
        np.testing.assert_allclose(F.conv2d(x, w, b, padding=1).numpy(), O._conv(sd, prefix, x, relu=False).numpy(),
                                   rtol=1e-10, atol=1e-12)
    else:
        ai = len(g['position']) - 1
        adapter = m.encoder.adapters[ai]
        cout = m.encoder.stages[int(g['position'][ai])][-2].weight.shape[0]
        w, b, needs_input = E.block_adapter_weights(adapter, cout)
        y = torch.randn(2, cout, 12, 10, dtype=torch.float64)
        if needs_input:
            xin = torch.randn(2, w.shape[1] - cout, 12, 10, dtype=torch.float64)
            ref = y + O._parallel_adapter(sd, f'encoder.adapters.{ai}', xin)
            got = F.conv2d(torch.cat([y, xin], 1), w.double(), None, padding=1)
        else:
            ref = O._serial_adapter(sd, f'encoder.adapters.{ai}', y)
            got = F.conv2d(y, w.double(), b.double(), padding=1)
        np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-10, atol=1e-12)


@pytest.mark.skipif(not ref_harness.available(), reason='needs the reference tree (build container only)')
@pytest.mark.parametrize('train_net,position', [('parallelLayer_3x3', [0, 1, 2, 3, 4]), ('parallelLayer_1x1_3x3', [1, 3]),
                                                ('serialLayer', [0, 2, 4]), ('serial', [0, 1, 2, 3, 4]),
                                                ('parallel_1x1', [0, 2]), ('parallel', [0, 1])])
def test_live_reference_adapter_state_dict_is_identical(train_net, position):
    """Same keys in the same order and the same default initialisation under one seed: checkpoints interchange."""
    from motion_style_transfer_b200.models.ynet import YNet
    ns = ref_harness.load()
    kw = dict(obs_len=5, pred_len=6, segmentation_model_fp=None, encoder_channels=[8, 8, 16, 16, 16],
              decoder_channels=[16, 16, 16, 8, 8], n_waypoints=2, train_net=train_net, position=list(position),
              network='original')
    torch.manual_seed(3)
    a = ns.ynet.YNet(**kw).state_dict()
    torch.manual_seed(3)
    b = YNet(**kw).state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_embed_network_oracle_against_reference_fixture():
    g = load_golden('embed')
    sd = golden_state_dict(g)
    torch.set_num_threads(1)
    np.testing.assert_allclose(O.embedding_forward(sd, 'scene_embedding', g['scene']).numpy(), g['scene_emb'],
                               rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.embedding_forward(sd, 'motion_embedding', g['motion']).numpy(), g['motion_emb'],
                               rtol=1e-5, atol=1e-6)


def test_semantic_adapter_mirrors_the_reference_failure():
    """ynet.py:516 passes position=None into get_conv2d (ynet.py:140): the reference raises TypeError; so does the drop-in."""
    from motion_style_transfer_b200.models.ynet import YNet
    kw = dict(obs_len=5, pred_len=6, segmentation_model_fp=None, encoder_channels=[8, 8, 16, 16, 16],
              decoder_channels=[16, 16, 16, 8, 8], n_waypoints=2, train_net='semantic_3x3', position=[], network='original')
    with pytest.raises(TypeError):
        YNet(**kw)
    if ref_harness.available():
        with pytest.raises(TypeError):
            ref_harness.load().ynet.YNet(**kw)


def test_swap_pavement_terrain_host_logic():
    from motion_style_transfer_b200.utils.image_utils import swap_pavement_terrain
    x = torch.arange(2 * 6 * 3 * 4, dtype=torch.float32).reshape(2, 6, 3, 4)
    assert torch.equal(swap_pavement_terrain(x), O.swap_pavement_terrain(x))
    if ref_harness.available():
        assert torch.equal(swap_pavement_terrain(x), ref_harness.load().image_utils.swap_pavement_terrain(x.clone()))
    with pytest.raises(ValueError):
        swap_pavement_terrain(x[0])


class _ReplayRng:
    """Feeds the randoms recorded in the fixture instead of drawing from global RNGs."""

    def __init__(self, g):
        self.g = g
        self.k = 0

    def uniforms(self, rows, n):
        return self.g['uniforms']

    def exponentials(self, rows, S):
        return self.g['expo']

    def kmeans_init(self, N, K):
        self.k += 1
        return self.g['kmeans_init'][self.k - 1]

    def reseed(self, N):
        raise AssertionError('fixture has no empty cluster')


@pytest.mark.parametrize('name', ['eval_sdd_short', 'eval_ind_long_ttst_cws'])
def test_evaluate_batch_against_reference_fixture(name):
    g = load_golden(name)
    c = ast.literal_eval(str(g['cfg']))
    sd = golden_state_dict(g)
    torch.set_num_threads(1)
    tmpl = O.create_dist_mat(int(g['template_size'])).astype(np.float32)
    ade, fde, aux = O.evaluate_batch(
        sd, g['scene'][None], g['trajectory'], tmpl, c['wps'], c['n_goal'], c['n_traj'], c['obs'],
        c['resize'], c['T'], c['ttst'], c['cws'], c['thr'], c['cwsp'], rng=_ReplayRng(g), return_all=True)
    np.testing.assert_allclose(aux['goal_map'].numpy(), g['goal_map'], rtol=1e-4, atol=1e-5)
    # (n_goal, B, n_wp, 2) -> reference stores (B, n_wp, n_goal, 2)
    # The reference divides by a thread-count-dependent global fp32 sum (image_utils.py:119); the
    # oracle DEFINES that sum as fp64-accumulated ("global-sum quirk"), which can flip
    # ~1 in 10^4 TTST draws and so move a k-means centre by <0.02 px.  Bound: the 0.05 px of north_star.
    np.testing.assert_allclose(aux['waypoint_samples'].permute(1, 2, 0, 3).numpy(), g['waypoint_sample'],
                               rtol=0, atol=0.05)
    np.testing.assert_allclose(ade.numpy(), g['ade'], rtol=0, atol=0.05)
    np.testing.assert_allclose(fde.numpy(), g['fde'], rtol=0, atol=0.05)


@pytest.mark.skipif(not ref_harness.available(), reason='live reference only in the build container')
def test_live_reference_sampling_and_kmeans():
    ns = ref_harness.load()
    torch.set_num_threads(1)
    torch.manual_seed(31)
    p = torch.sigmoid(torch.randn(2, 1, 32, 64) * 4)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        torch.manual_seed(77)
        ref = ns.image_utils.sampling(p, 3000, rel_threshold=0.01, replacement=True).numpy()
    torch.manual_seed(77)
    mine = O.sampling(p.numpy(), 3000, 0.01, True, O.HostRng.uniforms(2, 3000))
    assert np.array_equal(ref, mine)
    X = torch.from_numpy(ref[0, 0])
    np.random.seed(5)
    ids, cen = ns.kmeans.kmeans(X=X.clone(), num_clusters=19, distance='euclidean',
                                device=torch.device('cpu'), tqdm_flag=False, tol=0.001, iter_limit=1000)
    np.random.seed(5)
    init = O.HostRng.kmeans_init(3000, 19)
    ids2, c2, _ = O.kmeans(X.numpy(), 19, init, reseed_fn=lambda: O.HostRng.reseed(3000),
                           tol=0.001, iter_limit=1000)
    assert np.array_equal(cen.numpy(), c2) and np.array_equal(ids.numpy(), ids2)


def test_lora_init_is_noop():
    """train.py:46-59 --init_check: lora_B = 0 must be an exact no-op."""
    from oracle.loralib_restatement import Conv2d
    torch.manual_seed(0)
    conv = Conv2d(5, 7, 3, r=2, stride=1, padding=1)
    x = torch.randn(1, 5, 8, 8)
    base = torch.nn.functional.conv2d(x, conv.weight, conv.bias, padding=1)
    assert torch.equal(conv(x), base)
    assert conv.lora_A.shape == (6, 15) and conv.lora_B.shape == (21, 6)
    assert not conv.weight.requires_grad and conv.lora_A.requires_grad


@pytest.mark.parametrize('tag,network', [('ynet', 'original'), ('ynetmod', 'fusion')])
def test_train_epoch_against_reference_fixture(tag, network):
    """Oracle restatement of train_epoch.py:44-126 (two Adam steps) vs the live reference's fixture."""
    import ast
    g = load_golden(f'train_{tag}')
    c = ast.literal_eval(str(g['cfg']))
    sd = golden_state_dict(g)
    size = int(g['template_size'])
    dist_t = O.create_dist_mat(size).astype(np.float32)
    gauss_t = O.create_gaussian_heatmap_template(size, kernlen=c['kernlen'], nsig=c['nsig'],
                                                 normalize=False).astype(np.float32)
    torch.set_num_threads(1)
    ade, fde, loss, sd_after, grads = O.train_epoch(sd, g['scene'], g['trajectory'], dist_t, gauss_t, c['wps'], c['obs'],
                                                    c['pred'], c['batch_size'], c['lr'], c['loss_scale'], c['resize'],
                                                    network)
    assert abs(ade - float(g['train_ade'])) < 1e-3 and abs(fde - float(g['train_fde'])) < 1e-3
    assert abs(loss - float(g['train_loss'])) < 1e-3 * abs(float(g['train_loss']))
    keys = [k[8:] for k in g.files if k.startswith('trained/')]
    assert keys and sorted(keys) == sorted(grads)
    for k in keys:
        np.testing.assert_allclose(sd_after[k].numpy(), g['trained/' + k], rtol=0, atol=2e-6)
        ref = g['lastgrad/' + k]
        np.testing.assert_allclose(grads[k].numpy(), ref, rtol=0, atol=1e-4 * max(np.abs(ref).max(), 1e-6))


def test_product_synthetic_generators_match_the_oracles():
    """bench.py / tools feed the product from motion_style_transfer_b200.synthetic (the product never imports oracle);
    the oracle's own generators must produce the same tensors so that both arms see identical inputs."""
    from motion_style_transfer_b200 import synthetic as S
    assert torch.equal(S.synthetic_scene(32, 64, seed=3), O.synthetic_scene(32, 64, seed=3))
    for jitter in (0.0, 0.5):
        assert torch.equal(S.synthetic_tracks(5, 35, 64, 96, seed=7, jitter=jitter),
                           O.synthetic_tracks(5, 35, 64, 96, seed=7, jitter=jitter))
