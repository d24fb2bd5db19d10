"""CWS edge cases the reference's loop has but round 1 left untested on the GPU (evaluate.py:172-224):
more than two waypoints (the levels are conditioned sequentially, last to first) and n_traj > 1 (trajectories 1.. of every
goal RE-SAMPLE each earlier waypoint from the thresholded product map, sampling(.., 1, rel_threshold=0.05), with
sigma_factor reduced by the trajectory index)."""
import numpy as np
import pytest
import torch

from oracle import ynet_oracle as O

pytestmark = pytest.mark.gpu


class _RecordedExpo:
    """Exp(1) draws recorded once; the oracle and the product consume them in the same (goal-major, level-minor) order."""

    def __init__(self, n_draws, rows, S, seed):
        g = torch.Generator().manual_seed(seed)
        self.q = [torch.empty(rows, S, dtype=torch.float32).exponential_(1, generator=g) for _ in range(n_draws)]
        self.k = 0

    def _next(self):
        self.k += 1
        return self.q[self.k - 1]

    def oracle(self, rows, S):
        return self._next().numpy()

    def exponentials(self, rows, S, device):
        return self._next().to(device)

    def uniforms(self, rows, n, device):
        raise AssertionError('CWS re-sampling draws exponentials only (num_samples == 1)')


@pytest.mark.parametrize('n_wp,n_traj', [(3, 1), (2, 2), (3, 3)])
def test_cws_sequential_levels_and_resampling(cuda_device, n_wp, n_traj):
    from motion_style_transfer_b200.utils.evaluate import _cws
    torch.manual_seed(8)
    B, H, W, n_goal = 3, 64, 96, 4
    sig = torch.sigmoid(torch.randn(B, n_wp, H, W) * 2)
    goals = torch.stack([torch.rand(n_goal, B) * W, torch.rand(n_goal, B) * H], -1)      # (n_goal, B, 2)
    last = torch.stack([torch.rand(B) * W, torch.rand(B) * H], -1)
    params = dict(sigma_factor=6, ratio=2, rot=True)
    n_draws = n_goal * (n_traj - 1) * (n_wp - 1)
    rec_o, rec_p = _RecordedExpo(n_draws, B, H * W, 3), _RecordedExpo(n_draws, B, H * W, 3)
    ref = O.cws_waypoints(sig, goals.repeat(n_traj, 1, 1), last, n_goal, 6, 2, True, expo_fn=rec_o.oracle)   # (G, B, n_wp, 2)
    sig_list = [sig[:, i:i + 1].contiguous().cuda() for i in range(n_wp)]
    got, gs = _cws(None, sig_list, goals.unsqueeze(2).cuda(), last.cuda(), n_goal, n_traj, n_wp, params, rec_p)
    torch.cuda.synchronize()
    got = got.cpu()
    assert got.shape == ref.shape == (n_goal * n_traj, B, n_wp, 2)
    assert rec_o.k == rec_p.k == n_draws                      # same number of draws, same order
    # trajectory 0 of every goal: expectation of sigmoid x prior at every level (no draws), level by level
    np.testing.assert_allclose(got[:n_goal].numpy(), ref[:n_goal].numpy(), rtol=0, atol=1e-2)
    assert torch.equal(got[:, :, -1], goals.repeat(n_traj, 1, 1))      # the last waypoint is the goal itself
    if n_traj > 1:
        # re-sampled waypoints are pixel coordinates: top-1 of p / q.  The maps agree to ~1e-3, so nearly all draws land on
        # the oracle's pixel; a near-tie may flip one (and then moves the levels conditioned on it)
        same = (got[n_goal:, :, :-1] == ref[n_goal:, :, :-1]).all(-1).float().mean().item()
        assert same >= 0.85, same
        assert float((got[n_goal:, :, :-1] % 1).abs().max()) == 0.0     # integer pixels (idx % W, idx // W)
