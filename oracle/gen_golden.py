"""Generate tests/golden/*.npz from the LIVE, UNMODIFIED reference (build container only).

    python -m oracle.gen_golden          # needs /root/reference (read-only)

Every fixture stores the inputs, the explicit random numbers that the reference
consumed from its global generators (replayed from the seed)
and the reference's outputs.  ``tests/test_oracle_golden.py`` checks the oracle
restatement against them (CPU) and ``tests/test_gpu_*.py`` check the CUDA path.
The reference ships no golden vectors of its own.
"""
import os
import warnings

import numpy as np
import pandas as pd
import torch
from torch.utils.data import DataLoader

from oracle import ref_harness
from oracle import ynet_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

SMALL_ENC = [8, 8, 16, 16, 16]
SMALL_DEC = [16, 16, 16, 8, 8]


def build_ref_model(ns, obs, pred, n_wp, enc=SMALL_ENC, dec=SMALL_DEC, train_net='mosa_1',
                    position=(0, 1, 2, 3, 4), network='original', n_fusion=None, seed=0, peaky=1.0):
    torch.manual_seed(seed)
    m = ns.ynet.YNet(obs_len=obs, pred_len=pred, segmentation_model_fp=None,
                     encoder_channels=list(enc), decoder_channels=list(dec), n_waypoints=n_wp,
                     train_net=train_net, position=list(position), network=network, n_fusion=n_fusion)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if 'lora_B' in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        for d in (m.goal_decoder, m.traj_decoder):
            d.predictor.weight.mul_(peaky)
    return m


def make_df(tracks, scene='s0'):
    B, T, _ = tracks.shape
    rows = [dict(frame=t, trackId=b, x=float(tracks[b, t, 0]), y=float(tracks[b, t, 1]),
                 sceneId=scene, metaId=b) for b in range(B) for t in range(T)]
    return pd.DataFrame(rows)


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


def gen_templates(ns):
    for size in (1050, 1386):
        d = ns.image_utils.create_dist_mat(size)
        t = torch.Tensor(d).numpy()
        # strided sample + exact checksums of the float32 bit patterns
        save(f'dist_template_{size}', sample=t[::97, ::89], rows=t[[0, size // 2, size - 1]],
             xor=np.bitwise_xor.reduce(t.view(np.uint32).ravel()),
             sum64=np.float64(t.astype(np.float64).sum()))
    g = ns.image_utils.create_gaussian_heatmap_template(1050, kernlen=31, nsig=4, normalize=False)
    gt = torch.Tensor(g).numpy()
    save('gauss_template_1050', centre=gt[525 - 16:525 + 16, 525 - 16:525 + 16],
         sum64=np.float64(gt.astype(np.float64).sum()))


def gen_patch(ns):
    tmpl = torch.Tensor(ns.image_utils.create_dist_mat(130))
    traj = np.array([[10.5, 20.5], [0.5, 1.5], [47.49, 31.0], [11.5, 7.0], [0.0, 0.0], [2.5, 3.5]],
                    dtype=np.float32)
    out = torch.stack(ns.image_utils.get_patch(tmpl, traj, 32, 48)).numpy()
    save('get_patch', template=tmpl.numpy(), traj=traj, out=out, H=32, W=48)


def gen_sampling(ns):
    torch.manual_seed(3)
    p = torch.sigmoid(torch.randn(3, 1, 16, 24) * 3)
    cases = {}
    torch.manual_seed(11)
    cases['repl_out'] = ns.image_utils.sampling(p, 500, rel_threshold=0.3, replacement=True).numpy()
    torch.manual_seed(11)
    cases['repl_uniforms'] = O.HostRng.uniforms(3, 500)
    torch.manual_seed(12)
    cases['norepl_out'] = ns.image_utils.sampling(p, 20).numpy()
    torch.manual_seed(12)
    cases['norepl_expo'] = O.HostRng.exponentials(3, 16 * 24)
    torch.manual_seed(13)
    cases['one_out'] = ns.image_utils.sampling(p, 1, rel_threshold=0.05).numpy()
    torch.manual_seed(13)
    cases['one_expo'] = O.HostRng.exponentials(3, 16 * 24)
    torch.manual_seed(14)
    cases['repl_nothr_out'] = ns.image_utils.sampling(p, 64, replacement=True).numpy()
    torch.manual_seed(14)
    cases['repl_nothr_uniforms'] = O.HostRng.uniforms(3, 64)
    save('sampling', prob=p.numpy(), **cases)


def gen_softargmax(ns):
    torch.manual_seed(4)
    x = torch.randn(2, 3, 32, 48) * 4
    sa = ns.softargmax.SoftArgmax2D(normalized_coordinates=False)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = sa(x).numpy()
        sm = torch.softmax(x.view(2, 3, -1), 2).view_as(x)
        m = ns.ynet.YNet.softargmax_on_softmax_map(None, sm).numpy()
        smx = ns.ynet.YNet.softmax(None, x).numpy()
    save('softargmax', x=x.numpy(), out=out, softmax=smx, on_softmax_map=m)


def gen_kmeans(ns):
    torch.manual_seed(5)
    X = torch.stack([torch.randint(0, 96, (2000,)), torch.randint(0, 64, (2000,))], 1).float()
    np.random.seed(7)
    ids, cen = ns.kmeans.kmeans(X=X.clone(), num_clusters=7, distance='euclidean',
                                device=torch.device('cpu'), tqdm_flag=False, tol=0.001, iter_limit=1000)
    np.random.seed(7)
    init = O.HostRng.kmeans_init(2000, 7)
    # crafted duplicate points -> an empty cluster -> reseed path (kmeans.py:82-83)
    X2 = torch.cat([torch.tensor([[5., 5.]] * 6), X[:200]], 0)
    perm = np.arange(206)

    class _Fixed:
        pass
    orig_choice = np.random.choice
    np.random.choice = lambda n, k, replace=False: np.array([0, 1, 2, 50, 100])
    try:
        torch.manual_seed(21)
        ids2, cen2 = ns.kmeans.kmeans(X=X2.clone(), num_clusters=5, distance='euclidean',
                                      device=torch.device('cpu'), tqdm_flag=False, tol=0.001, iter_limit=1000)
    finally:
        np.random.choice = orig_choice
    torch.manual_seed(21)
    reseeds = np.array([int(torch.randint(206, (1,))) for _ in range(16)])
    save('kmeans', X=X.numpy(), init=init, ids=ids.numpy(), centres=cen.numpy(),
         X2=X2.numpy(), init2=np.array([0, 1, 2, 50, 100]), ids2=ids2.numpy(), centres2=cen2.numpy(),
         reseeds2=reseeds)


def gen_cws(ns):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        g1 = ns.evaluate.torch_multivariate_gaussian_heatmap(
            torch.tensor([40.3, 20.2]), 32, 48, torch.tensor([13.0, -7.5]), 6, 2, 'cpu', True)
        g2 = ns.evaluate.torch_multivariate_gaussian_heatmap(
            torch.tensor([10.0, 30.0]), 32, 48, torch.tensor([-3.0, 4.0]), 5, 2, 'cpu', False)
    save('cws_gaussian', g1=g1.numpy(), g2=g2.numpy())


def gen_network(ns):
    """Per-module forward of the reference model: features, goal logits, traj logits."""
    for tag, kw in (('ynet', dict(network='original')),
                    ('ynetmod', dict(network='fusion', n_fusion=2, position=('scene', 'motion', 'fusion')))):
        m = build_ref_model(ns, 5, 6, 2, seed=2, **kw).eval()
        torch.manual_seed(9)
        scene = torch.softmax(torch.randn(1, 6, 64, 96), 1).expand(2, -1, -1, -1)
        motion = torch.rand(2, 5, 64, 96) * 2
        wp = torch.rand(2, 2, 64, 96) * 2
        with torch.no_grad():
            feats = m.pred_features(scene, motion)
            goal = m.pred_goal(feats)
            pyr = O.avgpool_pyramid(wp, len(feats))
            traj = m.pred_traj([torch.cat([f, g], 1) for f, g in zip(feats, pyr)])
        sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
        save(f'network_{tag}', scene=scene[:1].numpy(), motion=motion.numpy(), wp=wp.numpy(),
             goal=goal.numpy(), traj=traj.numpy(),
             **{f'feat{i}': f.numpy() for i, f in enumerate(feats)},
             **{'sd/' + k: v for k, v in sd.items()})


ADAPTER_VARIANTS = (   # tag, train_net, position  (the comparison baselines, ynet.py:15-131, 237-283)
    ('parallelLayer_3x3', 'parallelLayer_3x3', (0, 1, 2, 3, 4)),      # scripts/sdd/ped_to_biker/tune_pa.sh:22
    ('parallelLayer_1x1_3x3', 'parallelLayer_1x1_3x3', (1, 3)),
    ('serialLayer', 'serialLayer', (0, 2, 4)),
    ('serial_block', 'serial', (0, 1, 2, 3, 4)),
    ('parallel_block_1x1', 'parallel_1x1', (0, 2)),
    ('parallel_block_1x1_3x3', 'parallel_1x1_3x3', (1, 4)),
)


def gen_adapters(ns):
    """Encoder features (+ goal logits) of the reference with the serial / parallel adapter baselines, eval mode, and
    the reference's autograd gradients of the parallel layer adapters (the only adapter baseline that fine-tunes here)."""
    for tag, train_net, position in ADAPTER_VARIANTS:
        m = build_ref_model(ns, 5, 6, 2, train_net=train_net, position=position, seed=4)
        g = torch.Generator().manual_seed(11)
        with torch.no_grad():          # adapters are zero-initialised (ynet.py:43-50): make them non-trivial
            for n, p in m.encoder.named_parameters():
                if 'serial_layer.1' in n or 'parallel_layer' in n:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1)
                elif 'serial_layer.0' in n:
                    p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=g) if n.endswith('weight')
                            else 0.2 * torch.randn(p.shape, generator=g))
            for n, b in m.encoder.named_buffers():
                if n.endswith('running_mean'):
                    b.copy_(0.3 * torch.randn(b.shape, generator=g))
                elif n.endswith('running_var'):
                    b.copy_(0.5 + torch.rand(b.shape, generator=g))
        m.eval()
        torch.manual_seed(12)
        scene = torch.softmax(torch.randn(1, 6, 32, 32), 1).expand(2, -1, -1, -1)
        motion = torch.rand(2, 5, 32, 32) * 2
        extra = {}
        if 'parallelLayer' in train_net:
            for p in m.parameters():
                p.requires_grad = False
            train = {n: p for n, p in m.encoder.named_parameters() if 'parallel' in n}     # trainer.py:133-135
            for p in train.values():
                p.requires_grad = True
            feats = m.pred_features(scene, motion)
            goal = m.pred_goal(feats)
            (goal.square().mean() * 100.0).backward()
            extra = {'grad/encoder.' + n: p.grad.numpy() for n, p in train.items()}
            feats = [f.detach() for f in feats]
            goal = goal.detach()
        else:
            with torch.no_grad():
                feats = m.pred_features(scene, motion)
                goal = m.pred_goal(feats)
        # only the encoder is stored: the decoders are the default initialisation under seed 4, which the drop-in
        # reproduces bit for bit (same modules, same creation order)
        sd = {k: v.detach().numpy() for k, v in m.state_dict().items() if k.startswith('encoder.')}
        save(f'adapter_{tag}', scene=scene[:1].numpy(), motion=motion.numpy(), goal=goal.numpy(), seed=np.array(4),
             train_net=np.array(train_net), position=np.array(position),
             **{f'feat{i}': f.numpy() for i, f in enumerate(feats)}, **extra,
             **{'sd/' + k: v for k, v in sd.items()})


def gen_embed_semantic(ns):
    """network='embed' (ynet.py:154-167, 529-531; evaluate.py:99-100,120-121) of the reference: embedded maps, goal logits
    and the autograd gradients of the embedding layers.  (The semantic adapter, ynet.py:513-519, cannot be constructed:
    get_conv2d iterates over position=None.)"""
    torch.manual_seed(12)
    scene = torch.softmax(torch.randn(1, 6, 32, 32), 1)
    motion = torch.rand(2, 5, 32, 32) * 2
    m = build_ref_model(ns, 5, 6, 2, train_net='all', position=(), network='embed', seed=5)
    sc = m.scene_embedding(scene)
    mo = m.motion_embedding(motion)
    feats = m.pred_features(sc.expand(2, -1, -1, -1), mo)
    goal = m.pred_goal(feats)
    (goal.square().mean() * 100.0).backward()
    grads = {'grad/' + n: p.grad.numpy() for n, p in m.named_parameters() if 'embedding' in n}
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items() if k.startswith(('encoder.', 'scene_', 'motion_'))}
    save('embed', scene=scene.numpy(), motion=motion.numpy(), scene_emb=sc.detach().numpy(), motion_emb=mo.detach().numpy(),
         goal=goal.detach().numpy(), seed=np.array(5), **grads, **{'sd/' + k: v for k, v in sd.items()})


def gen_evaluate(ns):
    cfgs = [
        dict(name='eval_sdd_short', H=64, W=96, obs=8, pred=12, wps=[11], B=3, resize=0.25, n_goal=20,
             n_traj=1, T=1.0, ttst=False, cws=False, thr=0.01, cwsp=None, peaky=1.0),
        dict(name='eval_ind_long_ttst_cws', H=64, W=96, obs=5, pred=30, wps=[14, 29], B=3, resize=0.33,
             n_goal=20, n_traj=1, T=1.8, ttst=True, cws=True, thr=0.002,
             cwsp=dict(sigma_factor=6, ratio=2, rot=True), peaky=30.0),
    ]
    for c in cfgs:
        m = build_ref_model(ns, c['obs'], c['pred'], len(c['wps']), peaky=c['peaky'])
        scene = O.synthetic_scene(c['H'], c['W'], seed=0)
        tracks = O.synthetic_tracks(c['B'], c['obs'] + c['pred'], c['H'], c['W'], seed=1)
        df = make_df(tracks / c['resize'])
        ds = ns.dataloader.SceneDataset(df, resize=c['resize'], total_len=c['obs'] + c['pred'])
        dl = DataLoader(ds, batch_size=1, collate_fn=ns.dataloader.scene_collate)
        size = int(4200 * c['resize'])
        tmpl = torch.Tensor(ns.image_utils.create_dist_mat(size))
        torch.manual_seed(100)
        np.random.seed(200)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ade, fde, dfo, td = ns.evaluate.evaluate(
                m, dl, {'s0': scene}, 'cpu', 'sdd', None, tmpl, c['wps'], 'test', c['n_goal'], c['n_traj'],
                c['obs'], c['B'], c['resize'], c['T'], c['ttst'], c['cws'], c['thr'], c['cwsp'],
                return_preds=True, return_samples=True, network='original')
            traj = next(iter(dl))[0]
        # replay the randoms the reference consumed (after the DataLoader's base_seed draw)
        torch.manual_seed(100)
        np.random.seed(200)
        torch.empty((), dtype=torch.int64).random_()
        rnd = {}
        if c['ttst']:
            rnd['uniforms'] = O.HostRng.uniforms(c['B'], 10000)
            rnd['kmeans_init'] = np.stack([O.HostRng.kmeans_init(10000, c['n_goal'] - 1) for _ in range(c['B'])])
        else:
            rnd['expo'] = O.HostRng.exponentials(c['B'], c['H'] * c['W'])
        sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
        save(c['name'], scene=scene.numpy(), trajectory=traj.numpy(), template_size=size,
             ade=dfo.ade.values.astype(np.float32), fde=dfo.fde.values.astype(np.float32),
             goal_map=td['goal_map'], goal_sigmoid_map=td['goal_sigmoid_map'],
             waypoint_sample=td['waypoint_sample'], prediction=td['prediction'],
             cfg=np.array(repr({k: v for k, v in c.items()})),
             **rnd, **{'sd/' + k: v for k, v in sd.items()})


def gen_preprocess():
    """Scene-image preprocessing (image_utils.py:66-107) recorded from the installed cv2 + the published smp constants."""
    import cv2
    from oracle import preprocess_oracle as P
    rng = np.random.RandomState(7)
    out = {}
    for tag, (H, W, f) in (('a', (97, 131, 0.33)), ('b', (64, 96, 0.25))):
        img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
        r = cv2.resize(img, (0, 0), fx=f, fy=f, interpolation=cv2.INTER_AREA)
        p = cv2.copyMakeBorder(r, 0, (-r.shape[0]) % 32, 0, (-r.shape[1]) % 32, cv2.BORDER_CONSTANT)
        x = ((p / 255.0) - P.IMAGENET_MEAN) / P.IMAGENET_STD
        out.update({f'img_{tag}': img, f'factor_{tag}': f, f'resized_{tag}': r,
                    f'chw_{tag}': x.transpose(2, 0, 1).astype('float32')})
    mask = rng.randint(0, 6, (97, 131)).astype(np.uint8)
    m = cv2.resize(mask, (0, 0), fx=0.33, fy=0.33, interpolation=cv2.INTER_NEAREST)
    m = cv2.copyMakeBorder(m, 0, (-m.shape[0]) % 32, 0, (-m.shape[1]) % 32, cv2.BORDER_CONSTANT)
    out['mask'] = mask
    out['onehot'] = np.stack([(m == v) for v in range(6)], axis=-1).transpose(2, 0, 1).astype('float32')
    save('preprocess', **out)


def gen_augment(ns):
    """utils/data_utils.py::augment_data (163-233) run LIVE on two synthetic scenes written to a temporary directory:
    the augmented DataFrame (8x: three rotations, then the mirror image of all four) and, for every augmented scene, the
    reference's rotated / flipped image pushed through resize -> pad -> normalise (cv2 + the published smp constants)."""
    import importlib
    import sys
    import tempfile
    import cv2
    from oracle import preprocess_oracle as P
    sys.path.insert(0, ref_harness.REF_PATH)
    try:
        du = importlib.import_module('utils.data_utils')
    finally:
        sys.path.remove(ref_harness.REF_PATH)
    rng = np.random.RandomState(11)
    sizes = {'sA': (60, 84), 's_B': (50, 50)}
    f = 0.5
    rows = []
    with tempfile.TemporaryDirectory() as tmp:
        out = {}
        for scene, (H, W) in sizes.items():
            img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
            os.makedirs(os.path.join(tmp, scene))
            cv2.imwrite(os.path.join(tmp, scene, 'reference.png'), img)
            out['img/' + scene] = img
            for meta in range(2):
                mid = len(rows) // 4
                for t in range(4):
                    rows.append(dict(frame=t, trackId=mid, x=float(rng.uniform(0, W)), y=float(rng.uniform(0, H)),
                                     sceneId=scene, metaId=mid))
        df = pd.DataFrame(rows)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            data, images = du.augment_data(df.copy(), image_path=tmp, images={}, image_file='reference.png')
    out['in/x'], out['in/y'] = df.x.values, df.y.values
    out['in/metaId'], out['in/sceneId'] = df.metaId.values.astype(np.int64), np.array(list(df.sceneId.values), dtype='U32')
    out['out/x'], out['out/y'] = data.x.values, data.y.values
    out['out/metaId'], out['out/sceneId'] = data.metaId.values.astype(np.int64), np.array(list(data.sceneId.values), dtype='U32')
    out['out/frame'] = data.frame.values.astype(np.int64)
    for key, im in images.items():
        r = cv2.resize(im, (0, 0), fx=f, fy=f, interpolation=cv2.INTER_AREA)
        pd_ = cv2.copyMakeBorder(r, 0, (-r.shape[0]) % 32, 0, (-r.shape[1]) % 32, cv2.BORDER_CONSTANT)
        x = ((pd_ / 255.0) - P.IMAGENET_MEAN) / P.IMAGENET_STD
        out['chw/' + key] = x.transpose(2, 0, 1).astype('float32')
    save('augment', factor=f, **out)


def gen_train(ns):
    """Two optimiser steps of the reference's train_epoch (train_epoch.py:8-136) with MoSA r=1 on stages 0-4:
    freeze policy of trainer.py:117,137-139, Adam(lr) + BCEWithLogitsLoss as trainer.py:197,206."""
    if ns.train_epoch is None:
        print('train_epoch not importable:', ns.trainer_error)
        return
    for tag, network, kw in (('ynet', 'original', {}),
                             ('ynetmod', 'fusion', dict(n_fusion=2, position=('scene', 'motion', 'fusion')))):
        c = dict(H=64, W=96, obs=5, pred=6, wps=[2, 5], B=4, batch_size=2, resize=0.25, lr=1e-3, loss_scale=1000,
                 kernlen=31, nsig=4)
        m = build_ref_model(ns, c['obs'], c['pred'], len(c['wps']), network=network, **kw)
        for p_ in m.parameters():
            p_.requires_grad = False
        for n, p_ in m.encoder.named_parameters():
            if 'lora' in n:
                p_.requires_grad = True
        sd0 = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
        scene = O.synthetic_scene(c['H'], c['W'], seed=0)
        tracks = O.synthetic_tracks(c['B'], c['obs'] + c['pred'], c['H'], c['W'], seed=3)
        df = make_df(tracks / c['resize'])
        ds = ns.dataloader.SceneDataset(df, resize=c['resize'], total_len=c['obs'] + c['pred'])
        dl = DataLoader(ds, batch_size=1, collate_fn=ns.dataloader.scene_collate)
        size = int(4200 * c['resize'])
        tmpl = torch.Tensor(ns.image_utils.create_dist_mat(size))
        gt_tmpl = torch.Tensor(ns.image_utils.create_gaussian_heatmap_template(size=size, kernlen=c['kernlen'],
                                                                               nsig=c['nsig'], normalize=False))
        opt = torch.optim.Adam(m.parameters(), lr=c['lr'])
        crit = torch.nn.BCEWithLogitsLoss()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ade, fde, loss = ns.train_epoch.train_epoch(
                m, dl, {'s0': scene}, opt, crit, c['loss_scale'], 'cpu', 'sdd', None, gt_tmpl, tmpl, c['wps'], 0,
                c['obs'], c['pred'], c['batch_size'], 100, c['resize'], network)
            traj = next(iter(dl))[0]
        trained = {n: p_.detach().numpy() for n, p_ in m.named_parameters() if p_.requires_grad}
        grads = {n: p_.grad.detach().numpy() for n, p_ in m.named_parameters() if p_.requires_grad}
        save(f'train_{tag}', scene=scene.numpy(), trajectory=traj.numpy(), template_size=size,
             train_ade=np.float32(ade), train_fde=np.float32(fde), train_loss=np.float32(loss),
             cfg=np.array(repr(c)),
             **{'sd/' + k: v for k, v in sd0.items()},
             **{'trained/' + k: v for k, v in trained.items()},
             **{'lastgrad/' + k: v for k, v in grads.items()})


def gen_forward_batch(ns):
    """The saliency forward of trainer.py:445-516 (``YNetTrainer._forward_batch``, unbound, on a stand-in ``self``):
    logits of both decoders, both losses and the gradient of a fixed linear functional of the logits with respect to the
    scene image (``set_input = ['scene', 'traj']``, no noise: the only configuration trainer.py:357-443 runs through)."""
    if ns.trainer is None:
        print('trainer not importable:', ns.trainer_error)
        return
    import types
    c = dict(H=64, W=96, obs=5, pred=6, wps=[2, 5], B=3, resize=0.25, loss_scale=1000, kernlen=31, nsig=4)
    m = build_ref_model(ns, c['obs'], c['pred'], len(c['wps']))
    for p_ in m.parameters():
        p_.requires_grad = False
    sd0 = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
    scene = O.synthetic_scene(c['H'], c['W'], seed=0)
    traj = torch.as_tensor(O.synthetic_tracks(c['B'], c['obs'] + c['pred'], c['H'], c['W'], seed=5), dtype=torch.float32)
    size = int(4200 * c['resize'])
    tmpl = torch.Tensor(ns.image_utils.create_dist_mat(size))
    gt_tmpl = torch.Tensor(ns.image_utils.create_gaussian_heatmap_template(size=size, kernlen=c['kernlen'],
                                                                           nsig=c['nsig'], normalize=False))
    crit = torch.nn.BCEWithLogitsLoss()
    me = types.SimpleNamespace(model=m, device='cpu')
    fb = ns.trainer.YNetTrainer._forward_batch
    img = scene.clone().unsqueeze(0).requires_grad_(True)
    goal, trj = fb(me, img, traj, tmpl, gt_tmpl, crit, c['obs'], c['pred'], c['wps'], c['loss_scale'], 'cpu',
                   ['scene', 'traj'], None, True)
    g = torch.Generator().manual_seed(11)
    r1 = torch.randn(goal.shape, generator=g)
    r2 = torch.randn(trj.shape, generator=g)
    ((goal * r1).sum() + (trj * r2).sum()).backward()
    grad_maps = img.grad.detach().clone()
    img2 = scene.clone().unsqueeze(0).requires_grad_(True)
    gl, tl = fb(me, img2, traj, tmpl, gt_tmpl, crit, c['obs'], c['pred'], c['wps'], c['loss_scale'], 'cpu',
                ['scene', 'traj'], None, False)
    (gl + tl).backward()
    save('forward_batch', scene=scene.numpy(), trajectory=traj.numpy(), template_size=size, cfg=np.array(repr(c)),
         goal_map=goal.detach().numpy(), traj_map=trj.detach().numpy(), functional_seed=11,
         grad_maps=grad_maps.numpy(), goal_loss=np.float32(gl.item()), traj_loss=np.float32(tl.item()),
         grad_loss=img2.grad.detach().numpy(), **{'sd/' + k: v for k, v in sd0.items()})


def _agents_frame(seed, n_agents=37, T=5):
    rng = np.random.RandomState(seed)
    ids = rng.permutation(200)[:n_agents]
    rows = [dict(frame=t, trackId=int(m), x=float(rng.rand()), y=float(rng.rand()), sceneId='s%d' % (m % 3), metaId=int(m))
            for m in ids for t in range(T)]
    return pd.DataFrame(rows)


def _xs(df):
    return None if df is None else [float(v) for v in df.x.values]


def _log_of_run(seed, exp, n_param, early, ade, fde, pre='ckpts/sdd__ynet__ped.pt', tuned=None):
    """What one train.py / test.py run prints, reduced to the lines the scraper reads (models/trainer.py:203,280,344,351;
    train.py:33; util.py:59)."""
    s = str({'save_every_n': 121, 'resize_factor': 0.25, 'seed': seed, 'pretrained_ckpt': pre, 'tuned_ckpt': tuned,
             'lr': 0.003}) + '\n'
    if exp:
        s += f'Experiment {exp} has started\n'
    if n_param is not None:
        s += 'The number of trainable parameters: {:d}\n'.format(n_param)
    if early is not None:
        s += f'Early stop at epoch {early}\n'
    s += f'Round 0: \nTest ADE: {ade + 1} \nTest FDE: {fde + 1}\n'
    s += f'\nAverage performance (by 3): \nTest ADE: {ade} \nTest FDE: {fde}\n'
    return s


def gen_scripts_host(ns):
    """the host side of train.py / test.py run LIVE -- utils/data_utils.py:14-48,754-912,955-964 (splits,
    sample limits, with numpy's global generator seeded), utils/parser.py, utils/util.py (names, parameter dictionaries)
    and utils/extract_log.py (scraped CSVs).  One JSON file: inputs and the reference's outputs / printed lines."""
    import contextlib
    import io
    import json
    import tempfile
    R, RU, RP, RX = ns.data_utils, ns.util, ns.parser, ns.extract_log
    out = {}

    def captured(fn, *a, **kw):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = fn(*a, **kw)
        return res, buf.getvalue()

    frames = {'main': _agents_frame(0), 'Biker.pkl': _agents_frame(1), 'Car.pkl': _agents_frame(2, 20),
              'train.pkl': _agents_frame(3, 30), 'val.pkl': _agents_frame(4, 8), 'test.pkl': _agents_frame(5, 9)}
    out['frames'] = {k: v.to_dict(orient='list') for k, v in frames.items()}
    df = frames['main']
    out['split_by_ratio'] = []
    for kw in [dict(val_split=0.1, test_split=0.2), dict(val_split=5, test_split=10, shuffle=True),
               dict(val_split=0.1, test_split=12, share_val_test=True), dict(val_split=0, test_split=12, share_val_test=True),
               dict(val_split=10, test_split=12, share_val_test=True, shuffle=True), dict(val_split=0.3),
               dict(val_split=0.1, test_split=0.2, given_test_meta_ids=[3, 5, 7, 88])]:
        np.random.seed(4)
        call = dict(kw)
        if 'given_test_meta_ids' in call:
            call['given_test_meta_ids'] = np.array(call['given_test_meta_ids'])
        parts, printed = captured(R.dataset_split_by_ratio, df, **call)
        out['split_by_ratio'].append(dict(kw=kw, seed=4, parts=[_xs(p) for p in parts], printed=printed,
                                          next_random=float(np.random.rand())))
    np.random.seed(1)
    out['limit_samples'] = dict(seed=1, num=3, batch_size=4, xs=_xs(R.limit_samples(df, 3, 4)),
                                xs_ordered=_xs(R.limit_samples(df, 2, 4, False)))
    out['downsample'] = dict(step=2, xs=_xs(R.downsample(df, 2)))
    short = df.drop(index=[0, 1, 7, 30, 31, 32, 33])
    out['filter_short'] = dict(dropped=[0, 1, 7, 30, 31, 32, 33], threshold=4, xs=_xs(R.filter_short_trajectories(short, 4)))
    out['prepare'] = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, f in frames.items():
            if name != 'main':
                f.to_pickle(os.path.join(tmp, name))
        two = ['Biker.pkl', 'Car.pkl']
        for args in [('sequential', 4, 2, two, two, 0.1, [5, 6], False, False, 'train', True),
                     ('sequential', 4, None, two[:1], two[:1], 0.1, [5], True, True, 'train', False),
                     ('sequential', 4, None, None, two, 0.1, [5, 7], False, True, 'eval', False),
                     ('predefined', 4, 2, None, None, None, None, True, False, 'train', False),
                     ('predefined', 4, None, None, None, None, None, False, False, 'eval', True)]:
            np.random.seed(7)
            parts, printed = captured(R.prepare_dataeset, tmp, *args)
            out['prepare'].append(dict(args=list(args), seed=7, parts=[_xs(p) for p in parts], printed=printed))
        _, printed = captured(R.split_train_val_test_randomly, tmp, 'Biker.pkl', 0.1, 0.2, seed=3)
        out['split_randomly'] = dict(val_split=0.1, test_split=0.2, seed=3, printed=printed, parts=[
            _xs(pd.read_pickle(os.path.join(tmp, 'Biker', n + '.pkl'))) for n in ('train', 'val', 'test')])
        sel, printed = captured(R.dataset_split_given_scenes, tmp, two, ['s1'])
        out['given_scenes'] = dict(scenes=['s1'], xs=_xs(sel), printed=printed)
    # flags, names
    train_argv = [
        '--fine_tune --config_filename sdd_shortterm_train.yaml --seed 3 --batch_size 10 --n_epoch 100 --n_early_stop 30 '
        '--n_round 3 --dataset_path filter/shortterm/agent_type/deathCircle_0/Biker --network original --load_data predefined '
        '--pretrained_ckpt ckpts/sdd__ynet__ped.pt --train_net mosa_1 --position 0 1 2 3 4 --ckpt_path ckpts/sdd/ped_to_biker '
        '--n_train_batch 3 --lr 0.003 --steps 20 --smooth_val',
        '--config_filename inD_longterm_train.yaml --dataset_path filter/longterm/agent_type/scene1 --network fusion '
        '--n_fusion 2 --train_files car.pkl truck.pkl --val_files car.pkl truck.pkl --test_splits 10 20 --val_split 0.2 '
        '--augment --ynet_bias --n_train_batch 0.5 --lr 0.00005 --n_epoch 50 --n_early_stop 300 --share_val_test --shuffle',
        '--config_filename x.yaml --dataset_path a/b --network embed --train_net all --load_data predefined']
    test_argv = ['--config_filename sdd_shortterm_eval.yaml --seed 2 --batch_size 10 --n_round 3 --dataset_path p/q --network '
                 'original --load_data predefined --pretrained_ckpt ckpts/a__b__ped.pt --tuned_ckpt '
                 'ckpts/x/Seed_1__p_q__mosa_1__Pos_0_1__TrN_20__lr_0.003__AUG__bias__original.pt']
    out['parser'] = []
    for is_train, argvs in ((True, train_argv), (False, test_argv)):
        for argv in argvs:
            a = RP.get_parser(is_train).parse_args(argv.split())
            rec = dict(is_train=is_train, argv=argv, namespace=dict(vars(a)))
            if is_train:
                rec['experiment'] = {str(n): RU.get_experiment_name(a, n) for n in (17, 3)}
            out['parser'].append(rec)
    names = ['ckpts/x/Seed_1__p_q__mosa_1__Pos_0_1__TrN_20__lr_0.003__AUG__bias__original.pt',
             'Seed_2__a__all__TrN_40__lr_0.00005__original.pt', 'Seed_1__p__encoder__TrN_10__original.pt']
    out['ckpt_names'] = [dict(path=n, name=RU.get_ckpt_name(n), position=RU.get_position(n), position_str=RU.get_position(
        n, return_list=False), train_net=RX.get_train_net(n), n_train=RX.get_n_train(n), lr=RX.get_lr(n),
        bias=RX.get_bool_bias(n), aug=RX.get_bool_aug(n)) for n in names]
    out['update_params'] = []
    for pre in ('ckpts/sdd__ynet__ped.pt', 'ckpts/sdd__ynet__ped_embed.pt'):
        base = {'pretrained_ckpt': pre, 'train_net': 'train', 'position': []}
        out['update_params'].append(dict(tuned=names[0], params=base, updated=RU.update_params(names[0], base)))
    out['ckpts_and_names'] = [dict(args=a, result=[list(r) for r in RU.get_ckpts_and_names(*a)]) for a in
                              [(['a', 'b'], ['A', 'B'], None, [None]), (None, None, 'pre.pt', names[:2])]]
    # get_params / get_image_and_data_path in a scratch working directory
    with tempfile.TemporaryDirectory() as tmp:
        for d in ('config', 'data/sdd/raw/annotations', 'data/sdd/p/q'):
            os.makedirs(os.path.join(tmp, d))
        yaml_text = ('save_every_n: 121\nresize_factor: 0.25\ndata_dir: data/\ndataset_name: sdd\nwaypoints:\n  - 11\n'
                     'CWS_params: None\n')
        with open(os.path.join(tmp, 'config', 'sdd_shortterm_eval.yaml'), 'w') as f:
            f.write(yaml_text)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            a = RP.get_parser(False).parse_args(test_argv[0].split())
            params, printed = captured(RU.get_params, a)
            out['get_params'] = dict(yaml=yaml_text, argv=test_argv[0], params=params, printed=printed,
                                     paths=list(RU.get_image_and_data_path(params)))
        finally:
            os.chdir(cwd)
    # logs -> CSV
    tuned = names[0]
    logs = {
        'sdd_train': _log_of_run(1, 'Seed_1__p_q__mosa_1__Pos_0_1_2_3_4__TrN_30__lr_0.003__smooth__early_30__original', 8190,
                                 41, 12.3456, 20.5)
        + _log_of_run(2, 'Seed_2__p_q__all__TrN_20__lr_0.00005__AUG__bias__original', None, None, 11.0, 19.25)
        + _log_of_run(3, 'Seed_3__p_q__encoder__TrN_10__fusion_2', 120, 7, 10.5, 18.0),
        'sdd_eval': _log_of_run(1, None, None, None, 12.5, 20.5, tuned=tuned)
        + _log_of_run(2, None, None, None, 13.5, 21.5, tuned='ckpts/x/Seed_2__p_q__all__TrN_40__original.pt'),
    }
    imp = ("{'save_every_n': 121, 'seed': 1, 'pretrained_ckpt': 'ckpts/pre.pt', 'tuned_ckpts': ['ckpts/t.pt'], 'x': 1}\n"
           + ''.join(f'Replacing encoder.stages.{i}.0\n\nAverage performance (by 1): \nTest ADE: {10 + i}.5 \n'
                     f'Test FDE: {20 + i}.25\n' for i in range(3)))
    logs['sdd_imp'] = imp + imp.replace("'seed': 1", "'seed': 2")
    out['logs'] = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, text in logs.items():
            with open(os.path.join(tmp, name + '.out'), 'w') as f:
                f.write(text)
            captured(RX.extract_file, os.path.join(tmp, name + '.out'), os.path.join(tmp, 'csv'))
            with open(os.path.join(tmp, 'csv', name + '.csv')) as f:
                out['logs'][name] = dict(text=text, csv=f.read())
    path = os.path.join(OUT, 'scripts_host.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(f'scripts_host: {os.path.getsize(path) / 1024:.1f} KiB')


VARFS = ['avg_vel', 'max_acc', 'abs+max_acc', 'min_vel', 'tot_vel', 'abs+min_acc', 'max_vel', 'avg_acc']


def _frame_arrays(prefix, df):
    out = {}
    for col in df.columns:
        v = df[col].to_numpy()
        out[f'{prefix}/{col}'] = np.array(list(v), dtype='U32') if v.dtype.kind in 'OUT' or str(df[col].dtype).startswith(
            ('str', 'object')) else v
    out[f'{prefix}/__columns__'] = np.array(list(df.columns), dtype='U32')
    return out


def gen_raw_datasets(ns):
    """The converters from the raw recordings (utils/sdd_dataset.py, inD_dataset.py, filter_dataset.py; data_utils.py:279-413)
    run LIVE on the synthetic recordings of oracle/synth_raw.py: the raw frames, the variation-factor table and the
    per-agent-type datasets.  ``load_and_window_*`` of the reference cannot run here (pandas 3, see oracle/data_oracle.py):
    the windowed frame the later stages are fed is the restatement's (oracle/data_oracle.py) and is stored with the fixture."""
    import contextlib
    import importlib
    import io
    import sys
    import tempfile
    from oracle import data_oracle, synth_raw
    sys.path.insert(0, ref_harness.REF_PATH)
    try:
        RS = importlib.import_module('utils.sdd_dataset')
        RI = importlib.import_module('utils.inD_dataset')
        RF = importlib.import_module('utils.filter_dataset')
    finally:
        sys.path.remove(ref_harness.REF_PATH)
    R = ns.data_utils
    out = {}

    def quiet(fn, *a, **kw):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = fn(*a, **kw)
        return res, buf.getvalue()

    with tempfile.TemporaryDirectory() as tmp:
        synth_raw.write_sdd(tmp, seed=0)
        raw = RS.load_raw_sdd(tmp)
        out.update(_frame_arrays('sdd/raw', raw))
        # sdd_dataset.py:44-50 with the two functions the reference cannot run replaced by their per-agent restatement
        w = data_oracle.split_fragmented(raw)
        w = R.downsample(w, step=12)
        w = R.filter_short_trajectories(w, threshold=10)
        w = data_oracle.sliding_window(w, window_size=10, stride=10)
        out.update(_frame_arrays('sdd/window', w))
        for obs in (4, 0):
            table, printed = quiet(R.get_varf_table, w, VARFS, obs)
            out.update(_frame_arrays(f'sdd/varf{obs}', table))
            out[f'sdd/varf{obs}/__printed__'] = np.array(printed)
        for tag, sel in (('all', None), ('sel', ['bookstore_0', 'coupa_3'])):
            d = os.path.join(tmp, 'agent_' + tag)
            _, printed = quiet(R.create_dataset_by_agent_type, w, ['Biker', 'Pedestrian'], d, False, selected_scenes=sel)
            files = sorted(os.path.relpath(os.path.join(dp, f), d) for dp, _, fs in os.walk(d) for f in fs)
            out[f'sdd/agent_{tag}/files'] = np.array(files, dtype='U64')
            out[f'sdd/agent_{tag}/printed'] = np.array(printed)
            for f in files:
                out[f'sdd/agent_{tag}/{f}'] = pd.read_pickle(os.path.join(d, f)).index.to_numpy()
        _, printed = quiet(R.create_dataset_by_agent_type, w, ['Pedestrian'], os.path.join(tmp, 'stat'), True)
        out['sdd/agent_stat/printed'] = np.array(printed)
        table.to_pickle(os.path.join(tmp, 'varf.pkl'))
        w[w.label == 'Pedestrian'].to_pickle(os.path.join(tmp, 'ped.pkl'))
        lo, hi = float(table.avg_vel.quantile(0.3)), float(table.avg_vel.quantile(0.8))
        _, printed = quiet(RF.filter_by_avg_vel, os.path.join(tmp, 'ped.pkl'), os.path.join(tmp, 'varf.pkl'), lo, hi)
        out['sdd/filter/bounds'] = np.array([lo, hi])
        out['sdd/filter/printed'] = np.array(printed)
        out['sdd/filter/index'] = pd.read_pickle(os.path.join(tmp, 'ped_filter.pkl')).index.to_numpy()
    with tempfile.TemporaryDirectory() as tmp:
        synth_raw.write_ind(tmp, seed=1)
        out.update(_frame_arrays('ind/raw', RI.load_raw_inD(tmp, recordings=['00', '07'])))
    save('raw_datasets', **out)


def main(only=None):
    torch.set_num_threads(1)     # reference's global-sum quirk depends on thread count
    ns = ref_harness.load()
    gen_templates(ns)
    gen_patch(ns)
    gen_sampling(ns)
    gen_softargmax(ns)
    gen_kmeans(ns)
    gen_cws(ns)
    gen_network(ns)
    gen_adapters(ns)
    gen_embed_semantic(ns)
    gen_evaluate(ns)
    gen_train(ns)
    gen_forward_batch(ns)
    gen_augment(ns)
    gen_scripts_host(ref_harness.load_scripts_host())
    gen_raw_datasets(ref_harness.load_scripts_host())
    gen_preprocess()


if __name__ == '__main__':
    main()
