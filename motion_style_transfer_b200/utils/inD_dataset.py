"""inD: raw ``<rec>_tracks.csv`` + ``<rec>_tracksMeta.csv`` -> the windowed trajectory frame and the per-agent-type pickles
(utils/inD_dataset.py:1-181).

    python -m motion_style_transfer_b200.utils.inD_dataset --labels pedestrian --selected_scenes scene1

Columns of the result: trackId, frame, x, y (pixels of the scene image), sceneId (scene1..4), metaId, label, recId.
"""
import os

import numpy as np
import pandas as pd

from . import raw_dataset
from .data_utils import downsample, filter_short_trajectories, sliding_window

# recordings of the four intersections (inD_dataset.py:24-27, 82-89)
SCENE_RECORDINGS = {1: range(0, 7), 2: range(7, 18), 3: range(18, 30), 4: range(30, 33)}
_UNUSED = ['trackLifetime', 'heading', 'width', 'length', 'xVelocity', 'yVelocity', 'xAcceleration', 'yAcceleration',
           'lonVelocity', 'latVelocity', 'lonAcceleration', 'latAcceleration']


def _recordings(scenes):
    return ['%02d' % r for s in scenes for r in SCENE_RECORDINGS[s]]


def load_raw_inD(path='inD-dataset-v1.0/data', scenes=[1], recordings=None):
    """inD_dataset.py:11-70: all road users of the chosen recordings with their class as ``label``, y mirrored (image rows
    grow downwards), points outside the frame dropped, agents numbered in order of appearance; ``sceneId`` is still the
    recording here."""
    frames = []
    for rec in (recordings if recordings is not None else _recordings(scenes)):
        track = pd.read_csv(os.path.join(path, f'{rec}_tracks.csv')).drop(columns=_UNUSED)
        meta = pd.read_csv(os.path.join(path, f'{rec}_tracksMeta.csv')).set_index('trackId')
        track['label'] = meta['class'].reindex(track['trackId']).to_numpy()
        track['rec&trackId'] = track.recordingId.astype(str) + '_' + track.trackId.astype(str).str.zfill(6)
        track['sceneId'] = rec
        track['yCenter'] = -track['yCenter']
        frames.append(track[(track['yCenter'] >= 0) & (track['xCenter'] >= 0)])
    data = pd.concat(frames, ignore_index=True)
    data['metaId'] = pd.factorize(data['rec&trackId'], sort=False)[0]
    data = data.rename(columns={'xCenter': 'x', 'yCenter': 'y'})
    return data.reindex(columns=['trackId', 'frame', 'x', 'y', 'sceneId', 'metaId', 'label'])


def load_and_window_inD(step, window_size, stride, scenes=[1, 2, 3, 4], path='inD-dataset-v1.0/data'):
    """inD_dataset.py:73-100: downsample (25 fps / step), drop short tracks, cut into windows, name the scene, scale metres
    to pixels of the (12x reduced) background image.  The two scale constants are applied as the reference's CODE does
    (scene1: 0.0127 * 12, the others: 0.00814 * 12), not as its comment says."""
    df = load_raw_inD(path=path, scenes=scenes, recordings=None)
    df = downsample(df, step=step)
    df = filter_short_trajectories(df, threshold=window_size)
    df = sliding_window(df, window_size=window_size, stride=stride)
    df['recId'] = df['sceneId'].copy()
    df['sceneId'] = df['recId'].map({rec: f'scene{s}' for s in SCENE_RECORDINGS for rec in _recordings([s])})
    scale = np.where(df.sceneId == 'scene1', 0.0127 * 12, 0.00814 * 12)
    df.x /= scale
    df.y /= scale
    return df


def main(argv=None):
    parser = raw_dataset.make_parser(
        'data/inD-dataset-v1.0/data', 'data_5_30_1fps.pkl', 'data/inD-dataset-v1.0/filter/longterm', step=25, window=35,
        obs_len=5, varf=['agent_type'], varf_ranges=[(0.25, 0.7), (1, 3)], labels=['pedestrian'],
        label_choices=['truck_bus', 'car', 'pedestrian', 'bicycle'], scenes=['scene1'])
    args = parser.parse_args(argv)
    return raw_dataset.build(args, lambda: load_and_window_inD(args.step, args.window_size, args.stride, scenes=[1, 2, 3, 4],
                                                               path=args.raw_data_dir))


if __name__ == '__main__':
    main()
