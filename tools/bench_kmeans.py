"""k-means (TTST, a14) alone: B agents x 10 000 integer pixel points, K = 19, on diffuse maps like the benchmark's
(~150-200 Lloyd iterations for the slowest agent).  Prints ms, the slowest agent's iterations and us per iteration, and a
checksum of the centres (identical across kernel revisions: the arithmetic is bit-exact by contract).

    python tools/bench_kmeans.py [B]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion_style_transfer_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N, K, H, W = 10000, 19, 416, 416
rng = np.random.RandomState(0)
X = np.empty((B, N, 2), np.float32)
for b in range(B):
    n_blob = rng.randint(2, 6)
    cen = rng.uniform(40, 376, (n_blob, 2))
    sig = rng.uniform(15, 60, n_blob)
    which = rng.randint(0, n_blob + 1, N)                 # last "blob" = uniform background
    pts = np.where((which == n_blob)[:, None], rng.uniform(0, W, (N, 2)),
                   cen[np.minimum(which, n_blob - 1)] + rng.normal(size=(N, 2)) * sig[np.minimum(which, n_blob - 1)][:, None])
    X[b] = np.clip(np.floor(pts), 0, W - 1)
init = np.stack([rng.choice(N, K, replace=False) for _ in range(B)]).astype(np.int32)
Xd, initd = torch.from_numpy(X).cuda(), torch.from_numpy(init).cuda()
reseed = torch.randint(0, N, (B, 64), dtype=torch.int32, device='cuda')
for _ in range(2):
    c, _, iters, st = ops.kmeans_batched(Xd, initd, reseed, 0.001, 1000)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    c, _, iters, st = ops.kmeans_batched(Xd, initd, reseed, 0.001, 1000)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
mx = int(iters.max())
print(f'B={B} {ms:.3f} ms, max iters {mx}, mean iters {float(iters.float().mean()):.1f}, {1000 * ms / mx:.2f} us/iter, '
      f'centres checksum {float(c.double().sum()):.6f}')
