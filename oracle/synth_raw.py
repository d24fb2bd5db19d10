"""TEST INFRASTRUCTURE ONLY -- tiny synthetic recordings in the RAW on-disk formats the reference's converters read
(utils/sdd_dataset.py:11-30: ``annotations/<scene>/video<k>/annotations.txt``, space separated, no header;
utils/inD_dataset.py:35-45: ``<rec>_tracks.csv`` + ``<rec>_tracksMeta.csv``).  Deterministic in ``seed``: the fixture
generator (oracle/gen_golden.py::gen_raw_datasets) and tests/test_raw_datasets.py write the same files."""
import os

import numpy as np
import pandas as pd

SDD_LABELS = ['Pedestrian', 'Biker', 'Cart', 'Pedestrian', 'Biker']


def write_sdd(root, seed=0, start=(100, 900), speed_x=(0.3, 2.0), speed_y=(-1.0, 1.0)):
    """3 videos, 5 tracks each, 280-520 frames at 30 fps; track 1 has a frame gap, track 2 a run of lost boxes.  (The
    defaults are the fixture's; narrower ``start`` / ``speed_*`` keep the tracks inside a small image.)"""
    rng = np.random.RandomState(seed)
    for scene, videos in (('bookstore', [0, 1]), ('coupa', [3])):
        for v in videos:
            d = os.path.join(root, 'annotations', scene, f'video{v}')
            os.makedirs(d)
            rows = []
            for tid, label in enumerate(SDD_LABELS):
                f0, n = int(rng.randint(0, 50)), int(rng.randint(280, 520))
                x, y = rng.uniform(start[0], start[1], 2)
                vx, vy = rng.uniform(*speed_x), rng.uniform(*speed_y)
                for k in range(n):
                    frame = f0 + k + (40 if (tid == 1 and k > 250) else 0)
                    x += vx + rng.normal(0, 0.2)
                    y += vy + rng.normal(0, 0.2)
                    lost = int(tid == 2 and 100 <= k < 106)
                    rows.append(f'{tid} {int(x)} {int(y)} {int(x) + 20} {int(y) + 31} {frame} {lost} 0 1 "{label}"')
            with open(os.path.join(d, 'annotations.txt'), 'w') as f:
                f.write('\n'.join(rows) + '\n')


IND_CLASSES = ['car', 'pedestrian', 'truck_bus', 'bicycle', 'pedestrian', 'pedestrian']
_IND_UNUSED = ['heading', 'width', 'length', 'xVelocity', 'yVelocity', 'xAcceleration', 'yAcceleration', 'lonVelocity',
               'latVelocity', 'lonAcceleration', 'latAcceleration']


def write_ind(root, seed=0, recordings=('00', '07')):
    """6 road users per recording, 200-420 frames at 25 fps, metres with y pointing up (mostly negative); one of them
    starts outside the frame (x < 0)."""
    rng = np.random.RandomState(seed)
    os.makedirs(root, exist_ok=True)
    for rec in recordings:
        rows, meta = [], []
        for tid, cls in enumerate(IND_CLASSES):
            n, f0 = int(rng.randint(200, 420)), int(rng.randint(0, 100))
            x, y = rng.uniform(5, 60), rng.uniform(-60, -5)
            if tid == 3:
                x = -1.0
            vx, vy = rng.uniform(0.02, 0.12), rng.uniform(-0.05, 0.05)
            meta.append({'recordingId': int(rec), 'trackId': tid, 'initialFrame': f0, 'finalFrame': f0 + n - 1, 'numFrames': n,
                         'width': 1.0, 'length': 2.0, 'class': cls})
            for k in range(n):
                x += vx + rng.normal(0, 0.01)
                y += vy + rng.normal(0, 0.01)
                row = {'recordingId': int(rec), 'trackId': tid, 'frame': f0 + k, 'trackLifetime': k, 'xCenter': x, 'yCenter': y}
                row.update({c: 0.0 for c in _IND_UNUSED})
                rows.append(row)
        pd.DataFrame(rows).to_csv(os.path.join(root, f'{rec}_tracks.csv'), index=False)
        pd.DataFrame(meta).to_csv(os.path.join(root, f'{rec}_tracksMeta.csv'), index=False)
