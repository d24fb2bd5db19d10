"""Drop-in for the reference's utils/softargmax.py (SoftArgmax2D, create_meshgrid)."""
from typing import Optional

import torch
import torch.nn as nn

from .. import ops


def create_meshgrid(x: torch.Tensor, normalized_coordinates: Optional[bool]):
    """softargmax.py:10-23 -- returns (pos_y, pos_x).  Host-side helper kept for API parity."""
    assert len(x.shape) == 4, x.shape
    _, _, height, width = x.shape
    if normalized_coordinates:
        xs = torch.linspace(-1.0, 1.0, width, device=x.device, dtype=x.dtype)
        ys = torch.linspace(-1.0, 1.0, height, device=x.device, dtype=x.dtype)
    else:
        xs = torch.linspace(0, width - 1, width, device=x.device, dtype=x.dtype)
        ys = torch.linspace(0, height - 1, height, device=x.device, dtype=x.dtype)
    return torch.meshgrid(ys, xs, indexing='ij')


class SoftArgmax2D(nn.Module):
    """Spatial soft-argmax (softargmax.py:26-81): (B, N, H, W) -> (B, N, 2) as (x, y).

    exp(x - max) / (sum + 1e-6) expectation of the pixel grid, computed in one streaming pass
    (online softmax) by ``ynet_softargmax2d``.
    """

    def __init__(self, normalized_coordinates: Optional[bool] = True) -> None:
        super().__init__()
        self.normalized_coordinates = normalized_coordinates
        self.eps = 1e-6

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(input):
            raise TypeError('Input input type is not a torch.Tensor. Got {}'.format(type(input)))
        if not len(input.shape) == 4:
            raise ValueError('Invalid input shape, we expect BxCxHxW. Got: {}'.format(input.shape))
        out = ops.softargmax2d(input)
        if self.normalized_coordinates:
            _, _, H, W = input.shape
            scale = torch.tensor([2.0 / (W - 1), 2.0 / (H - 1)], device=out.device, dtype=out.dtype)
            out = out * scale - 1.0
        return out
