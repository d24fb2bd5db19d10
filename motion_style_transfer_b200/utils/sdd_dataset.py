"""Stanford Drone Dataset: raw ``annotations/<scene>/video<k>/annotations.txt`` -> the windowed trajectory frame and the
per-agent-type pickles (utils/sdd_dataset.py:1-126).

    python -m motion_style_transfer_b200.utils.sdd_dataset --varf agent_type --labels Pedestrian Biker [--selected_scenes ...]

Columns of the result: trackId, frame, label, x, y (bounding-box centres), sceneId (``<scene>_<k>``), metaId.
"""
import os

import pandas as pd

from . import raw_dataset
from .data_utils import downsample, filter_short_trajectories, sliding_window, split_fragmented

SDD_COLUMNS = ['trackId', 'xmin', 'ymin', 'xmax', 'ymax', 'frame', 'lost', 'occluded', 'generated', 'label']


def load_raw_sdd(path):
    """sdd_dataset.py:11-41.  Lost boxes are dropped; agents are numbered in order of appearance over the scenes (sorted by
    name).  ``header=0`` with explicit names is the reference's call: the first annotation line of every file is consumed as
    a header -- kept, a dataset built here has to equal one built there row for row."""
    root = os.path.join(path, 'annotations')
    frames = []
    for scene in sorted(os.listdir(root)):
        for video in sorted(os.listdir(os.path.join(root, scene))):
            a = pd.read_csv(os.path.join(root, scene, video, 'annotations.txt'), header=0, names=SDD_COLUMNS, delimiter=' ')
            a = a[a['lost'] == 0]
            frames.append(pd.DataFrame({
                'trackId': a['trackId'], 'frame': a['frame'], 'label': a['label'], 'x': (a['xmax'] + a['xmin']) / 2,
                'y': (a['ymax'] + a['ymin']) / 2, 'sceneId': f"{scene}_{video.split('video')[1]}"}))
    data = pd.concat(frames, ignore_index=True)
    data['metaId'] = pd.factorize(data['sceneId'] + '_' + data['trackId'].astype(str).str.zfill(4), sort=False)[0]
    return data


def load_and_window_sdd(path, step, window_size, stride):
    """sdd_dataset.py:44-50: split tracks at frame gaps, downsample (30 fps / step), drop tracks shorter than one window,
    cut into windows."""
    df = split_fragmented(load_raw_sdd(path=path))
    df = downsample(df, step=step)
    df = filter_short_trajectories(df, threshold=window_size)
    return sliding_window(df, window_size=window_size, stride=stride)


def main(argv=None):
    parser = raw_dataset.make_parser(
        'data/sdd/raw', 'data_8_12_2_5fps.pkl', 'data/sdd/filter/shortterm', step=12, window=20, obs_len=8, varf=None,
        varf_ranges=[(0.5, 3.5), (4, 8)], labels=['Pedestrian', 'Biker'],
        label_choices=['Biker', 'Bus', 'Car', 'Cart', 'Pedestrian', 'Skater'], scenes=None)
    args = parser.parse_args(argv)
    return raw_dataset.build(args, lambda: load_and_window_sdd(args.raw_data_dir, args.step, args.window_size, args.stride))


if __name__ == '__main__':
    main()
