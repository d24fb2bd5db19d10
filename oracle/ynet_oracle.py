"""CPU restatement of the reference's Y-Net forecasting hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference lines it follows (paths relative to /root/reference).  Integer / index
work is written with numpy so that it is bit-reproducible; the convolutional
network uses torch CPU fp32 ops exactly as the reference does (the reference is
PyTorch: models/ynet.py).

All random numbers are EXPLICIT inputs (uniforms, exponentials, init indices);
helpers at the bottom draw them from the global torch / numpy generators in the
same order and quantity as the reference would, so that the live reference can be
replayed under a seed.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# a1/a2  templates  (utils/image_utils.py:7-37)
# --------------------------------------------------------------------------------------


def create_dist_mat(size, normalize=True):
    """utils/image_utils.py:30-37 -- Euclidean distance to the template centre.

    ``np.linalg.norm`` of an int64 index grid is ``sqrt(float64(di^2 + dj^2))``;
    the maximum is at index (0, 0): ``sqrt(2 * mid^2)``.  Returned as float64 like
    the reference; callers cast with ``torch.Tensor(...)`` / ``astype(float32)``
    (models/trainer.py:209,326).
    """
    mid = size // 2
    i = np.arange(size, dtype=np.int64) - mid
    sq = (i[:, None] ** 2 + i[None, :] ** 2).astype(np.float64)
    d = np.sqrt(sq)
    if normalize:
        d = d / d.max() * 2
    return d


def gkern(kernlen=31, nsig=4):
    """utils/image_utils.py:7-12."""
    ax = np.linspace(-(kernlen - 1) / 2., (kernlen - 1) / 2., kernlen)
    xx, yy = np.meshgrid(ax, ax)
    k = np.exp(-0.5 * (np.square(xx) + np.square(yy)) / np.square(nsig))
    return k / np.sum(k)


def create_gaussian_heatmap_template(size, kernlen=81, nsig=4, normalize=True):
    """utils/image_utils.py:15-27."""
    t = np.zeros([size, size])
    k = gkern(kernlen, nsig)
    m = k.shape[0]
    lo = size // 2 - int(np.floor(m / 2))
    hi = size // 2 + int(np.ceil(m / 2))
    t[lo:hi, lo:hi] = k
    if normalize:
        t = t / t.max()
    return t


# --------------------------------------------------------------------------------------
# a3  rasterisation  (utils/image_utils.py:40-63 + torch.stack at evaluate.py:113-114)
# --------------------------------------------------------------------------------------


def round_coords(traj):
    """``np.round(..).astype('int')`` -- round half to even (image_utils.py:52-53)."""
    traj = np.asarray(traj, dtype=np.float32)
    return np.rint(traj[:, 0]).astype(np.int64), np.rint(traj[:, 1]).astype(np.int64)


def get_patch_stack(template, traj, H, W):
    """get_patch + torch.stack: out[n] = template[mid-y : mid-y+H, mid-x : mid-x+W]."""
    template = np.asarray(template)
    x, y = round_coords(traj)
    mid_x = template.shape[1] // 2
    mid_y = template.shape[0] // 2
    out = np.empty((len(x), H, W), dtype=template.dtype)
    for n in range(len(x)):
        yl, xl = mid_y - y[n], mid_x - x[n]
        if yl < 0 or xl < 0 or yl + H > template.shape[0] or xl + W > template.shape[1]:
            raise ValueError('window leaves the template (coordinate outside the image)')
        out[n] = template[yl:yl + H, xl:xl + W]
    return out


def dist_patch_analytic(traj, H, W, size):
    """Analytic form of get_patch(create_dist_mat(size)): fp64 then cast."""
    x, y = round_coords(traj)
    mid = size // 2
    i = np.arange(H, dtype=np.int64)[None, :, None] - y[:, None, None]
    j = np.arange(W, dtype=np.int64)[None, None, :] - x[:, None, None]
    d = np.sqrt((i * i + j * j).astype(np.float64)) / np.sqrt(np.float64(2 * mid * mid)) * 2
    return d.astype(np.float32)


def avgpool_pyramid(maps, n_levels):
    """evaluate.py:255-257 / train_epoch.py:97-100: AvgPool2d(2^i) of the FULL-res map."""
    t = torch.as_tensor(maps)
    return [t] + [F.avg_pool2d(t, kernel_size=2 ** i, stride=2 ** i) for i in range(1, n_levels)]


# --------------------------------------------------------------------------------------
# a11  sampling  (utils/image_utils.py:110-135 + ATen multinomial CPU kernel)
# --------------------------------------------------------------------------------------


def threshold_normalise(prob, rel_threshold):
    """image_utils.py:113-119.  prob: (R, S) float32.

    The reference divides by the GLOBAL fp32 sum of the whole (R, S) view; that sum
    depends on torch's thread count ("global-sum normalisation quirk"), so
    the sum is DEFINED here as the float64-accumulated sum rounded to float32.
    """
    prob = np.asarray(prob, dtype=np.float32)
    mx = prob.max(axis=1, keepdims=True)
    thr = (mx * np.float32(rel_threshold)).astype(np.float32)
    keep = ~(prob < thr)
    p = (prob * keep.astype(np.float32)).astype(np.float32)
    s = np.float32(p.sum(dtype=np.float64))
    return (p / s).astype(np.float32)


def multinomial_with_replacement(prob, uniforms):
    """ATen MultinomialKernel.cpp (CPU, replacement=True), restated:

    per row: c[j] = sequential float32 running sum; c /= c[-1]; c[-1] = 1;
    idx = first j with double(c[j]) >= u  (lower bound), u ~ U[0,1) float64.
    prob: (R, S) float32, uniforms: (R, n) float64 -> (R, n) int64.
    """
    prob = np.asarray(prob, dtype=np.float32)
    uniforms = np.asarray(uniforms, dtype=np.float64)
    R, S = prob.shape
    out = np.empty(uniforms.shape, dtype=np.int64)
    for r in range(R):
        c = np.cumsum(prob[r], dtype=np.float32)  # sequential fp32 accumulation
        tot = c[-1]
        c = (c / tot).astype(np.float32)
        c[-1] = np.float32(1.0)
        out[r] = np.searchsorted(c.astype(np.float64), uniforms[r], side='left')
    return out


def multinomial_without_replacement(prob, expo, n):
    """ATen multinomial (replacement=False or n == 1): topk(p / q), q ~ Exp(1) float32.

    Returns indices by descending p/q (ties: lowest index first).
    """
    prob = np.asarray(prob, dtype=np.float32)
    expo = np.asarray(expo, dtype=np.float32)
    with np.errstate(divide='ignore', invalid='ignore'):
        r = (prob / expo).astype(np.float32)
    order = np.argsort(-r, axis=1, kind='stable')
    return order[:, :n].astype(np.int64)


def unravel_samples(idx, B, C, W):
    """image_utils.py:125-133: x = idx % W, y = floor(idx / W), float32, (B, C, n, 2)."""
    idx = idx.reshape(B, C, -1)
    out = np.empty(idx.shape + (2,), dtype=np.float32)
    out[..., 0] = (idx % W).astype(np.float32)
    out[..., 1] = np.floor(idx.astype(np.float32) / np.float32(W))
    return out


def sampling(prob_map, num_samples, rel_threshold=None, replacement=False, randoms=None):
    """utils/image_utils.py:110-135 with explicit randoms.

    randoms: float64 uniforms (R, n) when replacement and n > 1, else float32
    exponentials (R, S).  Returns (B, C, n, 2) float32 (x, y).
    """
    prob_map = np.asarray(prob_map, dtype=np.float32)
    B, C, H, W = prob_map.shape
    p = prob_map.reshape(B * C, H * W)
    if rel_threshold is not None:
        p = threshold_normalise(p, rel_threshold)
    if replacement and num_samples > 1:
        idx = multinomial_with_replacement(p, randoms)
    else:
        idx = multinomial_without_replacement(p, randoms, num_samples)
    return unravel_samples(idx, B, C, W)


# --------------------------------------------------------------------------------------
# a12/a13  soft-argmax, softmax  (utils/softargmax.py:55-81, models/ynet.py:578-600)
# --------------------------------------------------------------------------------------


def softargmax2d(x):
    """softargmax.py:64-81: e = exp(x - max); inv = 1/(sum(e) + 1e-6); (sum pos_x e inv, sum pos_y e inv)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    B, C, H, W = x.shape
    v = x.reshape(B, C, -1)
    e = torch.exp(v - v.max(dim=-1, keepdim=True)[0])
    inv = 1.0 / (e.sum(dim=-1, keepdim=True) + 1e-6)
    ys, xs = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing='ij')
    ex = ((xs.reshape(-1) * e) * inv).sum(dim=-1, keepdim=True)
    ey = ((ys.reshape(-1) * e) * inv).sum(dim=-1, keepdim=True)
    return torch.cat([ex, ey], dim=-1)


def spatial_softmax(x):
    """models/ynet.py:578-579."""
    x = torch.as_tensor(x, dtype=torch.float32)
    return torch.softmax(x.reshape(*x.shape[:2], -1), dim=2).view_as(x)


def softargmax_on_softmax_map(p):
    """models/ynet.py:588-600: plain expectation, no epsilon."""
    p = torch.as_tensor(p, dtype=torch.float32)
    _, _, H, W = p.shape
    ys, xs = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing='ij')
    v = p.flatten(2)
    return torch.cat([(xs.reshape(-1) * v).sum(-1, keepdim=True),
                      (ys.reshape(-1) * v).sum(-1, keepdim=True)], dim=-1)


# --------------------------------------------------------------------------------------
# a14  k-means  (utils/kmeans.py:9-108, 146-159)
# --------------------------------------------------------------------------------------


def kmeans(X, num_clusters, init_idx, reseed_fn=None, tol=1e-4, iter_limit=0):
    """Lloyd iterations exactly as utils/kmeans.py:72-106 (euclidean).

    X: (N, D) float32; init_idx: (K,) indices (np.random.choice at kmeans.py:17);
    reseed_fn(): returns the index used for an empty cluster (torch.randint at :83).
    Distances fl(fl(dx*dx)+fl(dy*dy)) (no FMA), argmin = first minimum, mean = fp32
    sum / count, shift = sum_k sqrt(sum_d delta^2), stop when shift^2 < tol.
    Returns (assignments (N,), centres (K, D), n_iterations).
    """
    X = np.ascontiguousarray(X, dtype=np.float32)
    c = X[np.asarray(init_idx)].copy()
    it = 0
    while True:
        diff = X[:, None, :] - c[None, :, :]
        sq = (diff * diff).astype(np.float32)
        dis = sq[..., 0]
        for d in range(1, X.shape[1]):
            dis = (dis + sq[..., d]).astype(np.float32)
        choice = np.argmin(dis, axis=1)
        pre = c.copy()
        for k in range(num_clusters):
            sel = X[choice == k]
            if sel.shape[0] == 0:
                sel = X[[int(reseed_fn())]]
            # torch mean on CPU: fp32 sum then one division by the count
            c[k] = (sel.sum(axis=0, dtype=np.float32) / np.float32(sel.shape[0])).astype(np.float32)
        d2 = ((c - pre) * (c - pre)).astype(np.float32)
        per = d2[:, 0]
        for d in range(1, X.shape[1]):
            per = (per + d2[:, d]).astype(np.float32)
        per = np.sqrt(per).astype(np.float32)
        shift = np.float32(0)
        for k in range(num_clusters):
            shift = np.float32(shift + per[k])
        it += 1
        if np.float32(shift * shift) < tol:
            break
        if iter_limit != 0 and it >= iter_limit:
            break
    return choice, c, it


# --------------------------------------------------------------------------------------
# a16  CWS  (utils/evaluate.py:9-34, 172-224)
# --------------------------------------------------------------------------------------


def cws_gaussian(mean_xy, H, W, dist, sigma_factor, ratio, rot=False):
    """torch_multivariate_gaussian_heatmap, evaluate.py:9-34 (torch CPU fp32)."""
    mean_xy = torch.as_tensor(mean_xy, dtype=torch.float32)
    dist = torch.as_tensor(dist, dtype=torch.float32)
    ax = torch.linspace(0, H, H) - mean_xy[1]
    ay = torch.linspace(0, W, W) - mean_xy[0]
    xx, yy = torch.meshgrid([ax, ay], indexing='ij')
    mesh = torch.stack([yy, xx], dim=-1)
    rad = torch.atan2(dist[0], dist[1])
    c, s = torch.cos(rad), torch.sin(rad)
    R = torch.tensor([[c, s], [-s, c]], dtype=torch.float32)
    if rot:
        R = torch.matmul(torch.tensor([[0., -1.], [1., 0.]]), R)
    dn = dist.square().sum(-1).sqrt() + 5
    conv = torch.tensor([[dn / sigma_factor / ratio, 0.], [0., dn / sigma_factor]], dtype=torch.float32)
    conv = torch.square(conv)
    T = torch.matmul(torch.matmul(R, conv), R.T)
    k = (torch.matmul(mesh, torch.inverse(T)) * mesh).sum(-1)
    k = torch.exp(-0.5 * k)
    return k / k.sum()


def cws_waypoints(sig_maps, goals, last_obs, n_goal, sigma_factor, ratio, rot, expo_fn=None):
    """evaluate.py:172-224.  sig_maps (B, n_wp, H, W) sigmoid maps; goals (G, B, 2)
    with G = n_goal * n_traj (already repeated); last_obs (B, 2).
    Returns waypoint samples (G, B, n_wp, 2).  For traj_idx > 0 the reference draws
    with sampling(.., 1, rel_threshold=0.05); ``expo_fn(B, S)`` supplies exponentials.
    """
    sig_maps = torch.as_tensor(sig_maps, dtype=torch.float32)
    goals = torch.as_tensor(goals, dtype=torch.float32)
    last_obs = torch.as_tensor(last_obs, dtype=torch.float32)
    B, n_wp, H, W = sig_maps.shape
    out = []
    for g_num in range(goals.shape[0]):
        wp = goals[g_num]
        lst = [wp]
        traj_idx = g_num // n_goal
        for wnum in reversed(range(n_wp - 1)):
            distance = last_obs - wp
            hm = []
            for dist, coord in zip(distance, wp):
                mean = coord + dist * (1 / (wnum + 2))
                hm.append(cws_gaussian(mean, H, W, dist, sigma_factor - traj_idx, ratio, rot))
            hm = torch.stack(hm)
            m = sig_maps[:, wnum] * hm
            m = (m.flatten(1) / m.flatten(1).sum(-1, keepdim=True)).view_as(m)
            if traj_idx == 0:
                wp = softargmax_on_softmax_map(m.unsqueeze(0)).squeeze(0)
            else:
                s = sampling(m.unsqueeze(1).numpy(), 1, rel_threshold=0.05,
                             randoms=expo_fn(B, H * W))
                wp = torch.from_numpy(s).permute(2, 0, 1, 3).squeeze(2).squeeze(0)
            lst.append(wp)
        out.append(torch.stack(lst[::-1]).permute(1, 0, 2))
    return torch.stack(out)


# --------------------------------------------------------------------------------------
# a4-a8  network forward from a state dict  (models/ynet.py)
# --------------------------------------------------------------------------------------


def effective_weight(sd, prefix):
    """loralib 0.1.1 Conv2d.forward (see oracle/loralib_restatement.py); ynet.py:143."""
    w = sd[prefix + '.weight']
    if prefix + '.lora_A' in sd:
        A, Bm = sd[prefix + '.lora_A'], sd[prefix + '.lora_B']
        r = A.shape[0] // w.shape[2]
        w = w + (Bm @ A).view(w.shape) * (1.0 / r)
    return w


def _serial_adapter(sd, prefix, x):
    """Adapter serial branch in eval mode (ynet.py:26-28, 63-65, 119-121): conv1x1(BatchNorm(x)) + x."""
    p = prefix + '.serial_layer'
    y = F.batch_norm(x, sd[p + '.0.running_mean'], sd[p + '.0.running_var'], sd[p + '.0.weight'], sd[p + '.0.bias'],
                     training=False, eps=1e-5)
    return F.conv2d(y, sd[p + '.1.weight'], sd.get(p + '.1.bias')) + x


def _parallel_adapter(sd, prefix, x):
    """Adapter parallel branch (ynet.py:30-41, 55-62, 122-129): sum of the k x k convs of x (same padding, no bias)."""
    p = prefix + '.parallel_layer'
    if p + '.weight' in sd:
        ws = [sd[p + '.weight']]
    else:
        ws, i = [], 0
        while f'{p}.{i}.weight' in sd:
            ws.append(sd[f'{p}.{i}.weight'])
            i += 1
    y = 0
    for w in ws:
        y = y + F.conv2d(x, w, None, padding=w.shape[-1] // 2)
    return y


def _conv(sd, prefix, x, relu=True, pad=1):
    y = F.conv2d(x, effective_weight(sd, prefix), sd[prefix + '.bias'], padding=pad)
    if prefix + '.serial_layer.1.weight' in sd:                       # AdapterLayer.forward, ynet.py:117-131
        y = _serial_adapter(sd, prefix, y)
    elif prefix + '.parallel_layer.weight' in sd or prefix + '.parallel_layer.0.weight' in sd:
        y = y + _parallel_adapter(sd, prefix, x)
    return F.relu(y) if relu else y


def _run_stage_list(sd, prefix, n_stage_total, x, first_has_pool, adapter_position=None):
    """A ModuleList of Sequential stages as built in ynet.py:192-215 / 309-367.  ``adapter_position``: stage ids that
    carry a block-level AdapterBlock ``encoder.adapters.{j}`` (YNetEncoderB.forward, ynet.py:258-283)."""
    feats = []
    i = 0
    pos = [int(p) for p in adapter_position] if adapter_position is not None else []
    while True:
        base = f'{prefix}.{i}'
        has0 = f'{base}.0.weight' in sd
        has1 = f'{base}.1.weight' in sd
        if has0:                      # [conv, relu]
            xin = x
            x = _conv(sd, f'{base}.0', x)
        elif has1:                    # [pool, conv, relu, conv, relu]
            x = F.max_pool2d(x, 2, 2)
            xin = x
            x = _conv(sd, f'{base}.1', x)
            x = _conv(sd, f'{base}.3', x)
        else:
            break
        if i in pos:
            a = f'encoder.adapters.{pos.index(i)}'
            if a + '.serial_layer.1.weight' in sd:        # x = adapter(stage(x)), ynet.py:262-266
                x = _serial_adapter(sd, a, x)
            else:                                         # x = stage(x) + adapter(stage input), ynet.py:267-279
                x = x + _parallel_adapter(sd, a, xin)
        feats.append(x)
        i += 1
    return feats, x, i


def pred_features(sd, scene_map, motion_map, network='original', adapter_position=None):
    """models/ynet.py:570-575, 229-234 (Y-Net), 258-283 (Y-Net with block-level adapters at ``adapter_position``) and
    369-395 (Y-Net-Mod)."""
    scene_map = torch.as_tensor(scene_map, dtype=torch.float32)
    motion_map = torch.as_tensor(motion_map, dtype=torch.float32)
    if network == 'fusion':
        sf, _, _ = _run_stage_list(sd, 'encoder.scene_stages', 0, scene_map, False)
        mf, _, _ = _run_stage_list(sd, 'encoder.motion_stages', 0, motion_map, False)
        feats = [torch.cat([a, b], dim=1) for a, b in zip(sf, mf)]
        ff, x, _ = _run_stage_list(sd, 'encoder.fusion_stages', 0, feats[-1], True)
        feats += ff
        feats.append(F.max_pool2d(x, 2, 2))     # trailing pool-only stage
        return feats
    x = torch.cat([scene_map, motion_map], dim=1)
    has_blocks = any(k.startswith('encoder.adapters.') for k in sd)
    feats, x, _ = _run_stage_list(sd, 'encoder.stages', 0, x, False, adapter_position if has_blocks else None)
    feats.append(F.max_pool2d(x, 2, 2))
    return feats


def decoder_forward(sd, prefix, features):
    """models/ynet.py:453-471."""
    f = features[::-1]
    x = _conv(sd, f'{prefix}.center.0', f[0])
    x = _conv(sd, f'{prefix}.center.2', x)
    for i, skip in enumerate(f[1:]):
        x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
        x = F.conv2d(x, sd[f'{prefix}.upsample_conv.{i}.weight'],
                     sd[f'{prefix}.upsample_conv.{i}.bias'], padding=1)
        x = torch.cat([x, skip], dim=1)
        x = _conv(sd, f'{prefix}.decoder.{i}.0', x)
        x = _conv(sd, f'{prefix}.decoder.{i}.2', x)
    return F.conv2d(x, sd[f'{prefix}.predictor.weight'], sd[f'{prefix}.predictor.bias'])


def pred_goal(sd, features):
    return decoder_forward(sd, 'goal_decoder', features)


def pred_traj(sd, features):
    return decoder_forward(sd, 'traj_decoder', features)


# --------------------------------------------------------------------------------------
# a17  ADE / FDE  (utils/evaluate.py:276-277, 290-291)
# --------------------------------------------------------------------------------------


def ade_fde(gt_future, trajs_samples, waypoint_samples, resize_factor):
    """gt_future (B, T, 2); trajs_samples (K, B, T, 2); waypoint_samples (K, B, n_wp, 2)."""
    gt_future = torch.as_tensor(gt_future, dtype=torch.float32)
    trajs_samples = torch.as_tensor(trajs_samples, dtype=torch.float32)
    waypoint_samples = torch.as_tensor(waypoint_samples, dtype=torch.float32)
    gt_goal = gt_future[:, -1:]
    ade_b = ((((gt_future - trajs_samples) / resize_factor) ** 2).sum(dim=3) ** 0.5).mean(dim=2)
    fde_b = ((((gt_goal - waypoint_samples[:, :, -1:]) / resize_factor) ** 2).sum(dim=3) ** 0.5)
    return ade_b.min(dim=0)[0], fde_b.min(dim=0)[0][:, 0]


# --------------------------------------------------------------------------------------
# a3->a17  one batch of the evaluate() inner loop  (utils/evaluate.py:109-291)
# --------------------------------------------------------------------------------------


class HostRng:
    """Draws randoms from the GLOBAL torch / numpy generators in the reference's order."""

    @staticmethod
    def uniforms(rows, n):            # consumed by multinomial(replacement=True)
        return torch.empty(rows * n, dtype=torch.float64).uniform_().reshape(rows, n).numpy()

    @staticmethod
    def exponentials(rows, S):        # consumed by multinomial(replacement=False | n == 1)
        return torch.empty(rows, S, dtype=torch.float32).exponential_(1).numpy()

    @staticmethod
    def kmeans_init(N, K):            # utils/kmeans.py:17
        return np.random.choice(N, K, replace=False)

    @staticmethod
    def reseed(N):                    # utils/kmeans.py:83
        return int(torch.randint(N, (1,)))


def evaluate_batch(sd, scene_image, trajectory, template, waypoints, n_goal, n_traj, obs_len,
                   resize_factor=0.25, temperature=1.0, use_TTST=False, use_CWS=False,
                   rel_thresh=0.002, CWS_params=None, network='original', rng=HostRng,
                   ttst_samples=10000, return_all=False):
    """One iteration of the batch loop at utils/evaluate.py:109-291 (non-'eth' datasets).

    scene_image (1, n_cls, H, W); trajectory (B, obs+pred, 2) already resized.
    Returns (ade (B,), fde (B,)) [+ dict of intermediates].
    """
    with torch.no_grad():
        scene_image = torch.as_tensor(scene_image, dtype=torch.float32)
        trajectory = torch.as_tensor(trajectory, dtype=torch.float32)
        template = np.asarray(template, dtype=np.float32)
        _, _, H, W = scene_image.shape
        B = trajectory.shape[0]
        observed = trajectory[:, :obs_len].reshape(-1, 2).numpy()
        observed_map = torch.from_numpy(get_patch_stack(template, observed, H, W)).reshape(-1, obs_len, H, W)
        gt_future = trajectory[:, obs_len:]
        semantic = scene_image.expand(B, -1, -1, -1)
        feats = pred_features(sd, semantic, observed_map, network)
        goal_map = pred_goal(sd, feats)
        wp_map = goal_map[:, waypoints]
        sig = torch.sigmoid(wp_map / temperature)

        if use_TTST:
            gs = sampling(sig[:, -1:].numpy(), ttst_samples, rel_threshold=rel_thresh, replacement=True,
                          randoms=rng.uniforms(B, ttst_samples))
            gs = torch.from_numpy(gs).permute(2, 0, 1, 3)            # (S, B, 1, 2)
            soft = softargmax2d(wp_map[:, -1:])                      # (B, 1, 2)
            centres = []
            for person in range(B):
                X = gs[:, person, 0].numpy()
                init = rng.kmeans_init(X.shape[0], n_goal - 1)
                _, c, _ = kmeans(X, n_goal - 1, init, reseed_fn=lambda: rng.reseed(X.shape[0]),
                                 tol=0.001, iter_limit=1000)
                centres.append(torch.from_numpy(c))
            gs = torch.stack(centres).permute(1, 0, 2).unsqueeze(2)
            goal_samples = torch.cat([soft.unsqueeze(0), gs], dim=0)
        else:
            gs = sampling(sig[:, -1:].numpy(), n_goal, randoms=rng.exponentials(B, H * W))
            goal_samples = torch.from_numpy(gs).permute(2, 0, 1, 3)  # (n_goal, B, 1, 2)

        if use_CWS and len(waypoints) > 1:
            goal_samples = goal_samples.repeat(n_traj, 1, 1, 1)
            last_obs = trajectory[:, obs_len - 1]
            waypoint_samples = cws_waypoints(sig, goal_samples.squeeze(2), last_obs, n_goal,
                                             CWS_params['sigma_factor'], CWS_params['ratio'],
                                             CWS_params['rot'], expo_fn=rng.exponentials)
        elif not use_CWS and len(waypoints) > 1:
            ws = sampling(sig[:, :-1].numpy(), n_goal * n_traj,
                          randoms=rng.exponentials(B * (len(waypoints) - 1), H * W))
            ws = torch.from_numpy(ws).permute(2, 0, 1, 3)
            goal_samples = goal_samples.repeat(n_traj, 1, 1, 1)
            waypoint_samples = torch.cat([ws, goal_samples], dim=2)
        else:
            waypoint_samples = goal_samples

        trajs = []
        for wp in waypoint_samples:
            wmap = torch.from_numpy(get_patch_stack(template, wp.reshape(-1, 2).numpy(), H, W))
            wmap = wmap.reshape(-1, len(waypoints), H, W)
            pyr = avgpool_pyramid(wmap, len(feats))
            tin = [torch.cat([f, g], dim=1) for f, g in zip(feats, pyr)]
            tmap = pred_traj(sd, tin)
            trajs.append(softargmax2d(tmap))
        trajs = torch.stack(trajs)
        ade, fde = ade_fde(gt_future, trajs, waypoint_samples, resize_factor)
    if return_all:
        return ade, fde, dict(goal_map=goal_map, sig=sig, goal_samples=goal_samples,
                              waypoint_samples=waypoint_samples, trajs=trajs, feats=feats)
    return ade, fde


# --------------------------------------------------------------------------------------
# a18  training step pieces  (utils/train_epoch.py:86-115)
# --------------------------------------------------------------------------------------


def bce_with_logits_mean(logits, target):
    """nn.BCEWithLogitsLoss() default (mean) -- models/trainer.py:206."""
    return F.binary_cross_entropy_with_logits(torch.as_tensor(logits), torch.as_tensor(target))


def adam_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults, single-tensor form (models/trainer.py:197).

    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
    p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
    """
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def train_epoch(sd, scene_image, trajectory, dist_template, gauss_template, waypoints, obs_len, pred_len,
                batch_size, lr, loss_scale=1000.0, resize_factor=0.25, network='original', trainable=None):
    """One epoch of utils/train_epoch.py:44-126 over ONE scene, from a state dict.

    ``sd`` maps names to float32 tensors; ``trainable`` lists the names that train (default: every key containing
    'lora' under 'encoder.', the MoSA freeze policy of models/trainer.py:117,137-139).  Gradients come from torch
    autograd over the restated forward (the reference does the same, train_epoch.py:109-115); Adam is the
    restated single-tensor update above.  Returns (train_ADE, train_FDE, train_loss, sd_after, last_grads).
    """
    sd = {k: torch.as_tensor(v).clone() for k, v in sd.items()}
    if trainable is None:
        trainable = [k for k in sd if k.startswith('encoder.') and 'lora' in k]
    state = {k: (torch.zeros_like(sd[k]), torch.zeros_like(sd[k])) for k in trainable}
    scene_image = torch.as_tensor(scene_image, dtype=torch.float32)
    if scene_image.dim() == 3:
        scene_image = scene_image[None]
    trajectory = torch.as_tensor(trajectory, dtype=torch.float32)
    _, _, H, W = scene_image.shape
    total_loss, ades, fdes, step, grads = 0.0, [], [], 0, {}
    for i in range(0, trajectory.shape[0], batch_size):
        traj = trajectory[i:i + batch_size]
        B = traj.shape[0]
        leaf = {k: (v.clone().requires_grad_(True) if k in state else v) for k, v in sd.items()}
        observed_map = torch.from_numpy(get_patch_stack(dist_template, traj[:, :obs_len].reshape(-1, 2).numpy(), H, W))
        observed_map = observed_map.reshape(B, obs_len, H, W)
        gt_future = traj[:, obs_len:]
        gt_future_map = torch.from_numpy(get_patch_stack(gauss_template, gt_future.reshape(-1, 2).numpy(), H, W))
        gt_future_map = gt_future_map.reshape(B, pred_len, H, W)
        gt_wp = gt_future[:, waypoints]
        gt_wp_map = torch.from_numpy(get_patch_stack(dist_template, gt_wp.reshape(-1, 2).numpy(), H, W))
        gt_wp_map = gt_wp_map.reshape(B, len(waypoints), H, W)
        feats = pred_features(leaf, scene_image.expand(B, -1, -1, -1), observed_map, network)
        goal_map = pred_goal(leaf, feats)
        goal_loss = bce_with_logits_mean(goal_map, gt_future_map) * loss_scale
        pyr = avgpool_pyramid(gt_wp_map, len(feats))
        traj_map = pred_traj(leaf, [torch.cat([f, p], dim=1) for f, p in zip(feats, pyr)])
        traj_loss = bce_with_logits_mean(traj_map, gt_future_map) * loss_scale
        loss = goal_loss + traj_loss
        names = list(state)
        gs = torch.autograd.grad(loss, [leaf[k] for k in names])
        step += 1
        for k, g in zip(names, gs):
            m, v = state[k]
            sd[k], m, v = adam_step(sd[k], g, m, v, step, lr)
            state[k] = (m, v)
            grads[k] = g
        with torch.no_grad():
            total_loss += float(loss)
            pt = softargmax2d(traj_map.detach())
            pg = softargmax2d(goal_map.detach()[:, -1:])
            ades.append(((((gt_future - pt) / resize_factor) ** 2).sum(dim=2) ** 0.5).mean(dim=1))
            fdes.append(((((gt_future[:, -1:] - pg) / resize_factor) ** 2).sum(dim=2) ** 0.5).mean(dim=1))
    return float(torch.cat(ades).mean()), float(torch.cat(fdes).mean()), total_loss, sd, grads


# --------------------------------------------------------------------------------------
# synthetic workloads  
# --------------------------------------------------------------------------------------


def synthetic_scene(H=416, W=416, n_cls=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.softmax(torch.randn(n_cls, H, W, generator=g), dim=0)


def synthetic_tracks(B, total_len, H=416, W=416, seed=0, jitter=0.0):
    """Constant-velocity tracks in pixel units of the (resized) image."""
    g = torch.Generator().manual_seed(seed)
    start = torch.rand(B, 1, 2, generator=g) * torch.tensor([0.4 * W, 0.4 * H]) + torch.tensor([0.3 * W, 0.3 * H])
    vel = (torch.rand(B, 1, 2, generator=g) * 2 - 1) * torch.tensor([0.01 * W, 0.01 * H]) * 20 / total_len
    t = torch.arange(total_len, dtype=torch.float32).view(1, -1, 1)
    tr = start + vel * t
    if jitter:
        tr = tr + torch.randn(B, total_len, 2, generator=g) * jitter
    return tr.float()


# --------------------------------------------------------------------------------------
# network='embed' (models/ynet.py:154-167; evaluate.py:99-100,120-121) and --swap_semantic
# --------------------------------------------------------------------------------------


def embedding_forward(sd, prefix, x):
    """Embedding.forward (ynet.py:166-167): three (conv3x3 + ReLU) layers ``<prefix>.conv.{0,2,4}``."""
    x = torch.as_tensor(x, dtype=torch.float32)
    for i in (0, 2, 4):
        x = F.relu(F.conv2d(x, sd[f'{prefix}.conv.{i}.weight'], sd[f'{prefix}.conv.{i}.bias'], padding=1))
    return x


def swap_pavement_terrain(semantic_img):
    """image_utils.py:165-171: exchange channels 1 and 2."""
    out = torch.as_tensor(semantic_img).clone()
    out[:, 1], out[:, 2] = torch.as_tensor(semantic_img)[:, 2], torch.as_tensor(semantic_img)[:, 1]
    return out
