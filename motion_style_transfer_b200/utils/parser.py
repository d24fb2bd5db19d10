"""Command-line flags of the reference's ``train.py`` / ``test.py`` (utils/parser.py:1-79): same flag names, types and
defaults, so the shell scripts under the reference's ``scripts/`` drive ``python -m motion_style_transfer_b200.train`` /
``.test`` unchanged.  One table instead of four builder functions; ``--backend`` is the only addition."""
import argparse

__all__ = ['get_parser']

_FLAG = dict(action='store_true')


def _multi(kind, default=None):
    return dict(default=default, type=kind, nargs='+')


# (flag, argparse keywords), grouped like parser.py:6-69
_DATA = [
    ('--dataset_path', dict(default=None, type=str)),
    ('--ckpt_path', dict(default='ckpts')),
    ('--shuffle', _FLAG),
    ('--augment', _FLAG),
    ('--load_data', dict(default='sequential', choices=['sequential', 'predefined'])),
    ('--show_details', _FLAG),
    # load_data == 'sequential'
    ('--val_split', dict(default=0.1, type=float)),
    ('--test_splits', dict(help='agents held out for testing, one number per file of --val_files', **_multi(int))),
    ('--val_files', _multi(str)),
    ('--share_val_test', _FLAG),
]
_MODEL = [
    # either a list of whole checkpoints ...
    ('--ckpts', _multi(str)),
    ('--ckpts_name', _multi(str)),
    # ... or a pretrained one plus the tuned parameters saved beside it
    ('--pretrained_ckpt', dict(default=None, type=str)),
    ('--tuned_ckpt', dict(default=None, type=str)),
    ('--tuned_ckpts', _multi(str)),
    ('--network', dict(choices=['original', 'embed', 'fusion'])),
    ('--n_fusion', dict(default=None, type=int)),
    ('--swap_semantic', _FLAG),
    ('--position', _multi(str, default=[])),
    ('--ynet_bias', _FLAG),
    ('--train_net', dict(default='train', type=str, help='which parameters train: all | train | encoder | mosa_<r> | serial* | parallel* | bias* | scene | motion | fusion ...')),
]
_GENERAL = [
    ('--seed', dict(default=1, type=int)),
    ('--batch_size', dict(default=8, type=int)),
    ('--gpu', dict(default=None, type=int, help='CUDA_VISIBLE_DEVICES for this run')),
    ('--n_round', dict(default=1, type=int, help='test rounds to average (TTST / CWS draw random numbers)')),
    ('--config_filename', dict(default=None, type=str)),
    ('--backend', dict(default=None, choices=['fp32', 'bf16x3', 'bf16'],
                       help='B200 engine (INTEGRATION.md section 2); default: YNET_BACKEND or fp32')),
]
_TRAIN = [
    ('--fine_tune', _FLAG),
    ('--n_epoch', dict(default=100, type=int)),
    ('--n_early_stop', dict(default=300, type=int)),
    ('--n_train_batch', dict(default=None, type=float,
                             help='low-shot fine-tuning: train on this many batches of agents only (may be fractional)')),
    ('--lr', dict(default=0.0001, type=float)),
    ('--steps', _multi(int, default=[])),
    ('--lr_decay_ratio', dict(default=0.1)),
    ('--init_check', _FLAG),
    ('--window_size', dict(default=9, type=int)),
    ('--smooth_val', _FLAG),
    # load_data == 'sequential'
    ('--train_files', _multi(str)),
]


def get_parser(is_train):
    """parser.py:72-79."""
    parser = argparse.ArgumentParser()
    for flag, kw in _DATA + _MODEL + _GENERAL + (_TRAIN if is_train else []):
        parser.add_argument(flag, **kw)
    return parser
