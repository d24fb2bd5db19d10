"""Keep the agents of a dataset whose variation factor lies in a range (utils/filter_dataset.py:1-34):

    python -m motion_style_transfer_b200.utils.filter_dataset --data_path .../scene1/pedestrian.pkl \\
        --varf_path data/inD-dataset-v1.0/data/varf_8_12_2_5fps.pkl --lower_bound 0.2

writes ``<data_path minus .pkl>_filter.pkl`` -- the ``*_filter`` datasets of the reference's training scripts (e.g.
``scene1/truck_bus_filter``).  The factor table is the one ``sdd_dataset`` / ``inD_dataset`` write beside the windowed pickle
(``data_utils.get_varf_table``: metaId, label, sceneId, scene, avg_vel, max_acc).
"""
import argparse

import numpy as np
import pandas as pd


def filter_by_avg_vel(data_path, varf_path, lower_bound=None, upper_bound=None, factor='avg_vel'):
    """filter_dataset.py:5-18.  The name is the reference's; ``factor`` may be any column of the table.  Bounds are
    inclusive, a missing bound is open."""
    trajectories = pd.read_pickle(data_path)
    table = pd.read_pickle(varf_path)
    table = table[np.isin(table['metaId'].to_numpy(), trajectories['metaId'].unique())]
    value = table[factor].to_numpy()
    inside = np.ones(len(table), dtype=bool)
    for bound, side in ((lower_bound, np.greater_equal), (upper_bound, np.less_equal)):
        if bound is not None:
            inside &= side(value, bound)
    kept = trajectories[np.isin(trajectories['metaId'].to_numpy(), table['metaId'].to_numpy()[inside])]
    print(f'Before filter: #={trajectories.shape[0]}')
    print(f'After filter: #={kept.shape[0]}')
    kept.to_pickle(data_path.replace('.pkl', '_filter.pkl'))


_FLAGS = (('data_path', str, None, 'the dataset pickle to filter'), ('varf_path', str, None, 'variation-factor table'),
          ('factor', str, 'avg_vel', 'column of the table'), ('lower_bound', float, None, None), ('upper_bound', float, None, None))


def main(argv=None):
    cli = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    for name, kind, default, text in _FLAGS:
        cli.add_argument('--' + name, type=kind, default=default, help=text)
    opt = cli.parse_args(argv)
    filter_by_avg_vel(opt.data_path, opt.varf_path, opt.lower_bound, opt.upper_bound, opt.factor)


if __name__ == '__main__':
    main()
