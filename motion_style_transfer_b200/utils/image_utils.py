"""Drop-in for the hot-path functions of the reference's utils/image_utils.py.

``create_dist_mat`` / ``create_gaussian_heatmap_template`` / ``gkern`` are one-off host numpy code
(identical formulas, image_utils.py:7-37); ``get_patch`` and ``sampling`` run on the device.
Image preprocessing (resize/pad/smp preprocessing, image2world, swap_pavement_terrain) is outside
the hot path.
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
