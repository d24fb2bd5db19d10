"""``python -m motion_style_transfer_b200.utils.split_dataset --data_dir ... --data_filename pedestrian.pkl --val_split 0.1
--test_split 0.2 --seed 1`` (utils/split_dataset.py:1-19, used by scripts/inD/preprocessing.sh): writes the predefined
train.pkl / val.pkl / test.pkl that ``--load_data predefined`` reads."""
import argparse

from .data_utils import split_train_val_test_randomly


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--data_dir', default=None, type=str,
                        help='Path to the raw data, can be a subset of the entire dataset')
    parser.add_argument('--data_filename', default=None, type=str)
    parser.add_argument('--val_split', default=None, type=float)
    parser.add_argument('--test_split', default=None, type=float)
    parser.add_argument('--seed', default=1, type=int)
    args = parser.parse_args(argv)
    split_train_val_test_randomly(args.data_dir, args.data_filename, args.val_split, args.test_split, args.seed)


if __name__ == '__main__':
    main()
