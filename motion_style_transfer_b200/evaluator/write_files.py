"""One-number-per-line result files (evaluator/write_files.py:1-48)."""
import csv
import os
import pathlib

import numpy as np


def round_val(num, ndig=4):
    """write_files.py:35-39: None -> 0.0, otherwise the rounded number as a string."""
    return 0.0 if num is None else str(round(num, ndig))


def convert_to_str(in_list):
    """write_files.py:42-48: one single-cell row per value."""
    return [[i if isinstance(i, str) else str(i)] for i in in_list]


def write_csv(out_dir, out_name, ade, fde, ade_final=None, fde_final=None):
    """write_files.py:8-21: [final ADE,] best ADE, mean ADE, then the ADE of every epoch -- one per line (the FDE arguments
    are accepted and unused, as in the reference)."""
    pathlib.Path(out_dir).mkdir(parents=True, exist_ok=True)
    rows = [round_val(min(ade)), round_val(np.mean(ade))] + [round_val(v) for v in ade]
    if ade_final is not None:
        rows = [round_val(ade_final)] + rows
    with open(os.path.join(out_dir, out_name), 'w') as f:
        csv.writer(f, dialect='excel').writerows(convert_to_str(rows))


def get_out_dir(out_dir, dataset_path, seed, train_net, val_files, train_files=None):
    """write_files.py:24-32: <out_dir>/<dataset_path>/<train files>/<val files>/<train_net>/<seed>."""
    def joined(files):
        return '_'.join('_' + f.split('.pkl')[0] + '_' for f in files)
    train_name = joined(train_files) if train_files else 'None'
    return os.path.join(out_dir, dataset_path, train_name.strip('_'), joined(val_files).strip('_'), train_net, str(seed))
