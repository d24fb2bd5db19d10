"""Synthetic workloads of the forecasting path: what bench.py and the tools feed the product.

The oracle keeps its own copy of these two generators (oracle/ynet_oracle.py::synthetic_scene / synthetic_tracks) so
that the product never imports ``oracle``; tests/test_oracle_golden.py checks that both produce the same tensors.
"""
import torch


def synthetic_scene(H=416, W=416, n_cls=6, seed=0):
    """Semantic map ``softmax(randn(n_cls, H, W), dim=0)`` float32 (what ``YNet.segmentation`` would hand over)."""
    g = torch.Generator().manual_seed(seed)
    return torch.softmax(torch.randn(n_cls, H, W, generator=g), dim=0)


def synthetic_tracks(B, total_len, H=416, W=416, seed=0, jitter=0.0):
    """Constant-velocity tracks (B, total_len, 2) in pixel units of the resized image: start ~ U(0.3, 0.7) of the
    image, velocity ~ U(-0.01, 0.01) of the image per step x 20 / total_len, optional N(0, jitter) noise."""
    g = torch.Generator().manual_seed(seed)
    start = torch.rand(B, 1, 2, generator=g) * torch.tensor([0.4 * W, 0.4 * H]) + torch.tensor([0.3 * W, 0.3 * H])
    vel = (torch.rand(B, 1, 2, generator=g) * 2 - 1) * torch.tensor([0.01 * W, 0.01 * H]) * 20 / total_len
    t = torch.arange(total_len, dtype=torch.float32).view(1, -1, 1)
    tr = start + vel * t
    if jitter:
        tr = tr + torch.randn(B, total_len, 2, generator=g) * jitter
    return tr.float()
