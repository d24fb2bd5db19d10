"""Drop-in for the reference's utils/kmeans.py (euclidean k-means) + a batched device entry point."""
import numpy as np
import torch

from .. import ops


last_iters = None


def initialize(X, num_clusters):
    """kmeans.py:9-19: initial centres = X[np.random.choice(N, K, replace=False)]."""
    indices = np.random.choice(len(X), num_clusters, replace=False)
    return X[indices]


def _reseed_stream(B, R, N, device):
    """Empty-cluster reseeds (kmeans.py:83 draws torch.randint lazily).  They are drawn from a side
    generator keyed on torch's initial seed so that the common case (no empty cluster) leaves the
    global stream exactly where the reference would."""
    g = torch.Generator().manual_seed(torch.initial_seed() % (2 ** 63 - 1))
    return torch.randint(N, (B, R), generator=g, dtype=torch.int32).to(device)


def kmeans_batched(X, num_clusters, init_idx=None, tol=1e-4, iter_limit=0, reseed_idx=None, want_assign=False):
    """X (B, N, 2) on the device -> (assignments (B, N) | None, centres (B, K, 2)).

    One CTA per agent, points resident in shared memory for all iterations.  init_idx (B, K):
    defaults to np.random.choice per agent, drawn in agent order like the reference's Python loop
    (evaluate.py:147-155).
    """
    B, N, _ = X.shape
    if init_idx is None:
        init_idx = np.stack([np.random.choice(N, num_clusters, replace=False) for _ in range(B)])
    if not torch.is_tensor(init_idx):
        init_idx = torch.as_tensor(np.asarray(init_idx))
    init_idx = init_idx.to(device=X.device, dtype=torch.int32)
    if reseed_idx is None:
        reseed_idx = _reseed_stream(B, 64, N, X.device)
    centres, assign, iters, status = ops.kmeans_batched(X.float(), init_idx, reseed_idx, tol, iter_limit, want_assign)
    global last_iters
    last_iters = iters          # device tensor (B,): Lloyd iterations per agent, for diagnostics
    return assign, centres


def kmeans(X, num_clusters, distance='euclidean', cluster_centers=[], tol=1e-4, tqdm_flag=True, iter_limit=0,
           device=torch.device('cpu')):
    """kmeans.py:22-108 signature.  Returns (cluster ids (N,), cluster centres (K, D))."""
    if distance != 'euclidean':
        raise NotImplementedError
    if type(cluster_centers) != list:
        raise NotImplementedError('resuming from given centres is not on the hot path (no caller in the reference)')
    X = X.float()
    if not X.is_cuda:
        raise RuntimeError('kmeans: X must live on the CUDA device (no CPU fallback)')
    assign, centres = kmeans_batched(X[None], num_clusters, tol=tol, iter_limit=iter_limit, want_assign=True)
    return assign[0].long(), centres[0]
