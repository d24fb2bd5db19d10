"""Drop-in for the hot-path functions of the reference's utils/image_utils.py.

``create_dist_mat`` / ``create_gaussian_heatmap_template`` / ``gkern`` are one-off host numpy code
(identical formulas, image_utils.py:7-37); ``get_patch`` and ``sampling`` run on the device.
Scene-image preprocessing (image_utils.py:66-107: resize / pad / segmentation-backbone normalisation)
runs as ONE fused CUDA launch per scene (``preprocess_scene_images``); ``resize`` / ``pad`` /
``preprocess_image_for_segmentation`` keep the reference's names and in-place dict semantics on top of the same kernels.
image2world (ETH/UCY homography) stays outside.
"""
import numpy as np
import torch

from .. import ops


def gkern(kernlen=31, nsig=4):
    ax = np.linspace(-(kernlen - 1) / 2., (kernlen - 1) / 2., kernlen)
    xx, yy = np.meshgrid(ax, ax)
    kernel = np.exp(-0.5 * (np.square(xx) + np.square(yy)) / np.square(nsig))
    return kernel / np.sum(kernel)


def create_gaussian_heatmap_template(size, kernlen=81, nsig=4, normalize=True):
    template = np.zeros([size, size])
    kernel = gkern(kernlen=kernlen, nsig=nsig)
    m = kernel.shape[0]
    lo = size // 2 - int(np.floor(m / 2))
    hi = size // 2 + int(np.ceil(m / 2))
    template[lo:hi, lo:hi] = kernel
    if normalize:
        template = template / template.max()
    return template


def create_dist_mat(size, normalize=True):
    middle = size // 2
    idx = np.arange(size, dtype=np.int64) - middle
    dist_mat = np.sqrt((idx[:, None] ** 2 + idx[None, :] ** 2).astype(np.float64))
    if normalize:
        dist_mat = dist_mat / dist_mat.max() * 2
    return dist_mat


def create_dist_template_device(size, device):
    """float32 create_dist_mat(size) built on the device (bit-identical, fp64 sqrt/div in-kernel)."""
    return ops.create_dist_template(size, device)


def get_patch_stack(template, traj, H, W):
    """Batched form used by the drivers: (n, H, W) device tensor, coordinates stay on the device."""
    if not torch.is_tensor(traj):
        traj = torch.as_tensor(np.asarray(traj, dtype=np.float32))
    traj = traj.to(device=template.device, dtype=torch.float32)
    return ops.rasterize_patches(template, traj, H, W)


def get_patch(template, traj, H, W):
    """image_utils.py:40-63: list of (H, W) windows of `template` centred on round(traj).

    Returns views of one stacked device tensor, so ``torch.stack(get_patch(...))`` is what the
    reference computes (bit-exact copies of template values, round-half-even like np.round).
    """
    return list(get_patch_stack(template, traj, H, W).unbind(0))


def swap_pavement_terrain(semantic_img):
    """image_utils.py:165-173: exchange semantic classes 1 and 2 of a (B, C, H, W) map (``--swap_semantic``).  Returns a
    new tensor (the reference swaps in place on the tensor it has just produced)."""
    if semantic_img.dim() != 4:
        raise ValueError(f'semanctic image has shape {semantic_img.shape} but should have 4 dimensions')
    order = list(range(semantic_img.shape[1]))
    order[1], order[2] = order[2], order[1]
    return semantic_img[:, order].contiguous()


class DeviceRng:
    """Counter-based (Philox) device generator for the production path.

    With ``graph_safe=True`` the per-step stream is selected by a DEVICE-resident epoch counter (mixed into
    the seed inside the kernels), so a captured CUDA graph draws fresh numbers on every replay: call
    ``next_step()`` once per step (inside the captured region).  Within a step, offsets restart from 0.
    """

    def __init__(self, seed=0, graph_safe=False):
        self.seed = int(seed)
        self.offset = 0
        self.graph_safe = graph_safe
        self.epoch = None

    def _epoch(self, device):
        if not self.graph_safe:
            return None
        if self.epoch is None:
            self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        return self.epoch

    def next_step(self, device):
        if self.graph_safe:
            ops.counter_add(self._epoch(device), 1)
            self.offset = 0

    def uniforms(self, rows, n, device):
        out = ops.rng_uniform_f64(self.seed, self.offset, rows * n, device, self._epoch(device)).view(rows, n)
        self.offset += (rows * n + 1) // 2
        return out

    def exponentials(self, rows, S, device):
        out = ops.rng_exponential_f32(self.seed, self.offset, rows * S, device, self._epoch(device)).view(rows, S)
        self.offset += (rows * S + 3) // 4
        return out

    def kmeans_init(self, rows, N, K, device):
        out = ops.rng_choice(self.seed, self.offset, rows, N, K, device, self._epoch(device))
        self.offset += 1
        return out

    def reseeds(self, rows, N, R, device):
        out = ops.rng_choice(self.seed ^ 0x5EED, self.offset, rows, N, R, device, self._epoch(device))
        self.offset += 1
        return out


class HostRng:
    """Draws from torch's GLOBAL CPU generator in the quantity and order torch.multinomial would
    (SURVEY App. A.6), so a seeded run reproduces the reference's CPU results index for index."""

    @staticmethod
    def uniforms(rows, n, device):
        return torch.empty(rows * n, dtype=torch.float64).uniform_().view(rows, n).to(device, non_blocking=True)

    @staticmethod
    def exponentials(rows, S, device):
        return torch.empty(rows, S, dtype=torch.float32).exponential_(1).to(device, non_blocking=True)


_default_rng = HostRng


def sampling(probability_map, num_samples, rel_threshold=None, replacement=False, rng=None):
    """image_utils.py:110-135: (B, C, H, W) probabilities -> (B, C, num_samples, 2) float32 (x, y).

    Threshold + global-sum normalise, multinomial (ATen CPU semantics: sequential fp32 CDF for
    replacement=True, exponential-race top-k otherwise) and index unravel, all on the device.
    """
    rng = rng or _default_rng
    B, C, H, W = probability_map.shape
    dev = probability_map.device
    if replacement and num_samples > 1:
        u = rng.uniforms(B * C, num_samples, dev)
        _, xy = ops.multinomial_replacement(probability_map, u, rel_threshold)
    else:
        q = rng.exponentials(B * C, H * W, dev)
        _, xy = ops.multinomial_topk(probability_map, q, num_samples, rel_threshold)
    return xy


# ---- scene-image preprocessing (image_utils.py:66-107, trainer.py:578-582) -----------------------------
# smp.encoders.get_preprocessing_fn('resnet101', 'imagenet') (segmentation_models_pytorch 0.1.0, requirements.txt:7; the
# package is absent, the constants are its published ones): input_space RGB, input_range [0, 1]
SMP_MEAN = (0.485, 0.456, 0.406)
SMP_STD = (0.229, 0.224, 0.225)


def area_table(ssize, dsize, scale):
    """cv::computeResizeAreaTab for one axis as CSR arrays (start (dsize + 1,), src, float32 weight), in OpenCV's
    accumulation order (host logic; the kernel walks it)."""
    start, src, w = [0], [], []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1, s2 = int(np.ceil(f1)), int(np.floor(f2))
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            src.append(s1 - 1)
            w.append((s1 - f1) / cell)
        for sx in range(s1, s2):
            src.append(sx)
            w.append(1.0 / cell)
        if f2 - s2 > 1e-3:
            src.append(s2)
            w.append(min(min(f2 - s2, 1.0), cell) / cell)
        start.append(len(src))
    return (np.asarray(start, np.int32), np.asarray(src, np.int32), np.asarray(w, np.float64).astype(np.float32))


def _resize_plan(H, W, factor):
    """(dh, dw, scale, integer scale or 0) of cv2.resize(img, (0, 0), fx=factor, fy=factor, INTER_AREA)."""
    dh, dw = int(np.rint(H * factor)), int(np.rint(W * factor))
    scale = 1.0 / factor
    isc = int(np.rint(scale))
    fast = abs(scale - isc) < np.finfo(np.float64).eps and dh * isc <= H and dw * isc <= W
    return dh, dw, scale, (isc if fast else 0)


def _ceil_to(v, d):
    return int(np.ceil(v / d) * d)


_tab_cache = {}


def _device_tabs(H, W, dh, dw, scale, device):
    key = (H, W, dh, dw, scale, str(device))
    hit = _tab_cache.get(key)
    if hit is None:
        xt = tuple(torch.from_numpy(a).to(device) for a in area_table(W, dw, scale))
        yt = tuple(torch.from_numpy(a).to(device) for a in area_table(H, dh, scale))
        hit = _tab_cache[key] = (xt, yt)
    return hit


def preprocess_scene_image(im, factor, division_factor=32, seg_mask=False, classes=6, device='cuda', orient=0):
    """resize -> pad -> preprocess_image_for_segmentation of ONE scene image (trainer.py:578-582) in one launch:
    uint8 (H, W, 3) numpy / tensor (or (H, W) mask with seg_mask) -> float32 (C, Hp, Wp) CUDA tensor.
    orient = k + 4 * flip: the same for the augmented view fliplr^flip(rot90^k(im)) (data_utils.py:115-233), read from
    the stored image itself."""
    t = torch.as_tensor(np.ascontiguousarray(im) if isinstance(im, np.ndarray) else im)
    if t.dtype != torch.uint8:
        raise TypeError(f'scene images are uint8 (cv2.imread), got {t.dtype}')
    t = t.to(device).contiguous()
    H, W = (t.shape[1], t.shape[0]) if (orient & 1) else t.shape[:2]        # the view's size
    dh, dw, scale, isc = _resize_plan(H, W, factor)
    Hp, Wp = _ceil_to(dh, division_factor), _ceil_to(dw, division_factor)
    if seg_mask:
        if orient:
            raise NotImplementedError('augmented views of segmentation masks (ETH/UCY) are outside the B200 hot path')
        return ops.scene_onehot_u8(t, dh, dw, Hp, Wp, scale, classes)
    xt, yt = (None, None) if isc else _device_tabs(H, W, dh, dw, scale, t.device)
    return ops.scene_preprocess_u8(t, dh, dw, Hp, Wp, xt, yt, isc, SMP_MEAN, SMP_STD, orient=orient)[0]


AUGMENT_SUFFIX = {0: '', 1: '_rot90', 2: '_rot180', 3: '_rot270'}


def augment_data(data, images):
    """utils/data_utils.py::augment_data (163-233), the trajectory half: every scene three more times rotated
    counter-clockwise by k * 90 degrees about the image centre (``rot``, 115-142), then all of those mirrored
    horizontally (``fliplr``, 145-160): 8x the rows, new ``sceneId`` suffixes and ``metaId`` offsets as the reference
    assigns them.  Same float64 numpy arithmetic as the reference (``np.dot(xy, R)`` with R from cos / sin of -k pi / 2),
    so the coordinates are bit-identical.  ``images``: {sceneId: uint8 image as read from disk}.

    Returns (augmented DataFrame, {augmented sceneId: (original sceneId, orient)}): the images themselves are NOT
    rotated on the host -- ``preprocess_scene_image(..., orient=)`` reads the view from the stored image."""
    import pandas as pd

    def size(scene, k):
        h, w = images[scene].shape[:2]
        return (w, h) if k % 2 else (h, w)                 # (y0, x0) of the view after k rotations

    views = {scene: (scene, 0) for scene in data.sceneId.unique()}
    data_ = data.copy()
    for k in (1, 2, 3):
        meta_max = data['metaId'].max()
        for scene in data_.sceneId.unique():
            xy = data_[data_.sceneId == scene].copy()
            y0, x0 = images[scene].shape[:2]
            xy.loc()[:, 'x'] = xy['x'] - x0 / 2
            xy.loc()[:, 'y'] = xy['y'] - y0 / 2
            c, s = np.cos(-k * np.pi / 2), np.sin(-k * np.pi / 2)
            xy.loc()[:, ['x', 'y']] = np.dot(xy[['x', 'y']], np.array([[c, s], [-s, c]]))
            y1, x1 = size(scene, k)
            xy.loc()[:, 'x'] = xy['x'] + x1 / 2
            xy.loc()[:, 'y'] = xy['y'] + y1 / 2
            xy['sceneId'] = scene + AUGMENT_SUFFIX[k]
            xy['metaId'] = xy['metaId'] + meta_max + 1
            views[scene + AUGMENT_SUFFIX[k]] = (scene, k)
            data = pd.concat([data, xy], axis=0)
    meta_max = data['metaId'].max()
    for scene in data.sceneId.unique():
        base, k = views[scene]
        xy = data[data.sceneId == scene].copy()
        y0, x0 = size(base, k)
        xy.loc()[:, 'x'] = xy['x'] - x0 / 2
        xy.loc()[:, 'y'] = xy['y'] - y0 / 2
        xy.loc()[:, ['x', 'y']] = np.dot(xy[['x', 'y']], np.array([[-1, 0], [0, 1]]))
        xy.loc()[:, 'x'] = xy['x'] + x0 / 2
        xy.loc()[:, 'y'] = xy['y'] + y0 / 2
        xy['sceneId'] = xy['sceneId'] + '_fliplr'
        xy['metaId'] = xy['metaId'] + meta_max + 1
        views[scene + '_fliplr'] = (base, k + 4)
        data = pd.concat([data, xy], axis=0)
    return data, views


def preprocess_scene_images(images, factor, division_factor=32, seg_mask=False, classes=6, device='cuda'):
    """The three preprocessing calls of trainer.py:578-582 over a dict of scene images, fused: {scene: uint8 image} ->
    {scene: float32 (C, Hp, Wp) CUDA tensor} (in place, like the reference's helpers)."""
    for key, im in images.items():
        images[key] = preprocess_scene_image(im, factor, division_factor, seg_mask, classes, device)
    return images


def resize(images, factor, seg_mask=False):
    """image_utils.py:85-92 (in place): cv2.resize(fx = fy = factor, INTER_AREA; INTER_NEAREST for masks) on the device;
    the dict keeps numpy uint8 images like the reference's."""
    for key, im in images.items():
        t = torch.as_tensor(np.ascontiguousarray(im)).cuda()
        H, W = t.shape[:2]
        dh, dw, scale, isc = _resize_plan(H, W, factor)
        if seg_mask:
            ys = torch.clamp(torch.floor(torch.arange(dh, dtype=torch.float64, device=t.device) * scale).long(), max=H - 1)
            xs = torch.clamp(torch.floor(torch.arange(dw, dtype=torch.float64, device=t.device) * scale).long(), max=W - 1)
            images[key] = t[ys][:, xs].cpu().numpy()
            continue
        xt, yt = (None, None) if isc else _device_tabs(H, W, dh, dw, scale, t.device)
        images[key] = ops.scene_preprocess_u8(t, dh, dw, dh, dw, xt, yt, isc, want_chw=False, want_u8=True)[1].cpu().numpy()


def pad(images, division_factor=32):
    """image_utils.py:95-107 (in place): zero border at the bottom / right up to a multiple of division_factor."""
    for key, im in images.items():
        H, W = im.shape[:2]
        widths = ((0, _ceil_to(H, division_factor) - H), (0, _ceil_to(W, division_factor) - W)) + ((0, 0),) * (im.ndim - 2)
        images[key] = np.pad(im, widths, mode='constant')


def preprocess_image_for_segmentation(images, encoder='resnet101', encoder_weights='imagenet', seg_mask=False, classes=6):
    """image_utils.py:66-82 (in place): uint8 HWC -> normalised float32 CHW tensor (one-hot for masks)."""
    if (encoder, encoder_weights) != ('resnet101', 'imagenet'):
        raise NotImplementedError('only the resnet101 / imagenet preprocessing the reference uses (image_utils.py:66)')
    for key, im in images.items():
        t = torch.as_tensor(np.ascontiguousarray(im)).cuda()
        H, W = t.shape[:2]
        if seg_mask:
            images[key] = ops.scene_onehot_u8(t, H, W, H, W, 1.0, classes)
        else:
            images[key] = ops.scene_preprocess_u8(t, H, W, H, W, None, None, 1, SMP_MEAN, SMP_STD)[0]
