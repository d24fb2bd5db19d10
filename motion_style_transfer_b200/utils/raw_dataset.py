"""Shared second half of ``utils/sdd_dataset.py`` / ``utils/inD_dataset.py`` (their ``__main__`` blocks, sdd_dataset.py:84-126,
inD_dataset.py:132-181): the windowed trajectory pickle, the variation-factor table (avg_vel, max_acc) beside it, and the
per-agent-type datasets the training scripts read.  Host-side pandas; nothing here reaches the GPU."""
import argparse
import os

import pandas as pd

from .data_utils import create_dataset_by_agent_type, create_dataset_given_range, get_varf_table

VARF_HELP = ("Variation factors from: 'avg_vel', 'max_vel', 'avg_acc', 'max_acc', 'abs+max_acc', 'abs+avg_acc', "
             "'agent_type' (the neighbour-distance factors 'min_dist', 'avg_den*' are not built)")


def make_parser(data_dir, filename, filter_dir, step, window, obs_len, varf, varf_ranges, labels, label_choices, scenes):
    """The flags both converters share; only the defaults differ (sdd_dataset.py:56-82, inD_dataset.py:103-129)."""
    p = argparse.ArgumentParser()
    p.add_argument('--additional_data_dir', default=data_dir, type=str, help='where the variation-factor table goes')
    p.add_argument('--raw_data_dir', default=data_dir, type=str, help='the raw recordings (or a subset of them)')
    p.add_argument('--raw_data_filename', default=filename, type=str)
    p.add_argument('--filter_data_dir', default=filter_dir, type=str)
    p.add_argument('--reload', action='store_true', help='read the windowed pickle written by an earlier run')
    p.add_argument('--statistic_only', action='store_true', help='print the agent counts, write no dataset')
    p.add_argument('--step', default=step, type=int)
    p.add_argument('--window_size', default=window, type=int)
    p.add_argument('--stride', default=window, type=int)
    p.add_argument('--obs_len', default=obs_len, type=int)
    p.add_argument('--varf', default=varf, nargs='+', help=VARF_HELP)
    p.add_argument('--varf_ranges', default=varf_ranges, help='range of varation factor to take')
    p.add_argument('--labels', default=labels, nargs='+', type=str, choices=label_choices)
    p.add_argument('--selected_scenes', default=scenes, type=str, nargs='+')
    return p


def build(args, load_and_window):
    """Load (or reload) the windowed frame, write it and its variation-factor table, then the per-agent-type pickles."""
    args.labels.sort()
    print(args)
    pickle_path = os.path.join(args.raw_data_dir, args.raw_data_filename)
    if args.reload:
        df = pd.read_pickle(pickle_path)
        print('Reloaded raw dataset')
    else:
        if args.varf is not None and any('dist' in f or 'den' in f for f in args.varf):
            raise NotImplementedError('neighbour-distance variation factors (data_utils.py:520-540) are not built')
        df = load_and_window()
        print('Loaded raw dataset')
        df.to_pickle(pickle_path)
        print(f'Saved data to {pickle_path}')
        varf_path = os.path.join(args.additional_data_dir, args.raw_data_filename.replace('data', 'varf'))
        get_varf_table(df, ['avg_vel', 'max_acc'], args.obs_len).to_pickle(varf_path)
        print(f'Saved variation factor data to {varf_path}')
    if args.varf is None:
        return df
    if args.varf == ['agent_type']:
        create_dataset_by_agent_type(df, args.labels, os.path.join(args.filter_data_dir, 'agent_type'),
                                     statistic_only=args.statistic_only, selected_scenes=args.selected_scenes)
    else:
        if any('dist' in f or 'den' in f for f in args.varf):
            raise NotImplementedError('neighbour-distance variation factors (data_utils.py:520-540) are not built')
        out_dir = os.path.join(args.filter_data_dir, '__'.join(args.varf), '_'.join(args.labels))
        create_dataset_given_range(df, args.varf, args.varf_ranges, args.labels, out_dir, obs_len=args.obs_len,
                                   statistic_only=args.statistic_only)
    print(f'Created dataset: \nVariation factor = {args.varf} \nAgents = {args.labels}')
    return df
