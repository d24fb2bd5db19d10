"""The variation-factor table of an existing trajectory pickle (utils/generate_varf.py:1-34):

    python -m motion_style_transfer_b200.utils.generate_varf --raw_data_dir data/sdd/raw --raw_data_filename data_8_12_2_5fps.pkl

writes ``<additional_data_dir>/df_varfs.pkl`` (or ``--varf_path``) with the observed average speed of every agent."""
import argparse
import os

import pandas as pd

from .data_utils import get_varf_table


def main(argv=None):
    cli = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    for name, kind, default in (('additional_data_dir', str, 'data/sdd/raw'), ('raw_data_dir', str, None),
                                ('raw_data_filename', str, None), ('varf_path', str, None), ('obs_len', int, 8)):
        cli.add_argument('--' + name, type=kind, default=default)
    opt = cli.parse_args(argv)
    print(opt)
    df = pd.read_pickle(os.path.join(opt.raw_data_dir, opt.raw_data_filename))
    print('Loaded raw dataset')
    table = get_varf_table(df, ['avg_vel'], opt.obs_len)
    out_path = opt.varf_path if opt.varf_path is not None else os.path.join(opt.additional_data_dir, 'df_varfs.pkl')
    table.to_pickle(out_path)
    print(f'Saved variation factor data to {out_path}')


if __name__ == '__main__':
    main()
