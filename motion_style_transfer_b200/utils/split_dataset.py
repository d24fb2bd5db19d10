"""Random train / val / test split of one trajectory pickle, as a command (utils/split_dataset.py:1-19; called fifteen times
by the reference's scripts/*/preprocessing.sh):

    python -m motion_style_transfer_b200.utils.split_dataset --data_dir <dir> --data_filename pedestrian.pkl \\
        --val_split 0.1 --test_split 0.2 --seed 1

A split > 1 is a number of agents, otherwise a fraction.  Output: ``<dir>/pedestrian/{train,val,test}.pkl`` -- what
``--load_data predefined`` reads (data_utils.split_train_val_test_randomly).
"""
import argparse

from .data_utils import split_train_val_test_randomly

_FLAGS = (('data_dir', str, None), ('data_filename', str, None), ('val_split', float, None), ('test_split', float, None),
          ('seed', int, 1))


def main(argv=None):
    cli = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    for name, kind, default in _FLAGS:
        cli.add_argument('--' + name, type=kind, default=default)
    opt = cli.parse_args(argv)
    split_train_val_test_randomly(opt.data_dir, opt.data_filename, opt.val_split, opt.test_split, opt.seed)


if __name__ == '__main__':
    main()
