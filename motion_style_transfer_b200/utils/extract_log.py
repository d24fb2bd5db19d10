"""Result scraper (utils/extract_log.py:1-170): turns the stdout of a batch of ``train`` / ``test`` runs (one ``.out`` file
per shell script of the reference's ``scripts/``) into one CSV row per run.  A run starts where its parameter dictionary
is printed (``get_params``: the first key is ``save_every_n``); the fields are the lines the trainer prints
(models/trainer.py: 'The number of trainable parameters', 'Early stop at epoch', 'Average performance (by n)') and the
experiment / checkpoint name grammar of ``utils/util.py``.

    python -m motion_style_transfer_b200.utils.extract_log --file_path logs/sdd_train.out --out_dir csv/log

One declarative field table per log kind instead of three hand-rolled loops.  The only deviation: a seed may have more than
one digit (the reference's pattern ``[\\d+]`` matches a single character, and its ``astype(int)`` then fails on None).
"""
import argparse
import pathlib
import re

import pandas as pd

from .util import get_position

_RUN = 'save_every_n'
_SEED = r"'seed': (\d+),"
_PRETRAINED = r"'pretrained_ckpt': '(.*?)',"
_METRIC = r'Average performance \(by [\d]+\): \nTest ADE: ([\d\.]+) \nTest FDE: ([\d\.]+)'


def _basename(path):
    return path.split('/')[-1]


def _field(msg, pattern, group=1, default=None, post=None):
    m = re.search(pattern, msg)
    if m is None:
        return default
    return post(m.group(group)) if post else m.group(group)


# ---- fields of a checkpoint / experiment name (extract_log.py:100-146) -------------------------------------------
def get_train_net(ckpt_path):
    return ckpt_path.split('__')[2] if ckpt_path is not None else None


def get_n_train(ckpt_path):
    return int(ckpt_path.split('TrN_')[-1].split('_')[0]) if ckpt_path is not None else None


def get_lr(ckpt_path):
    if ckpt_path is None:
        return None
    if 'lr' not in ckpt_path:
        return 0.00005                      # runs older than the lr field used this rate
    return ckpt_path.split('lr_')[1].split('_')[0].split('.pt')[0]


def get_bool_bias(ckpt_path):
    return 'bias' in ckpt_path.split('TrN')[-1] if ckpt_path is not None else None


def get_bool_aug(ckpt_path):
    return 'AUG' in ckpt_path if ckpt_path is not None else None


_NAME_FIELDS = [('train_net', get_train_net), ('n_train', get_n_train),
                ('position', lambda name: get_position(name, return_list=False)), ('lr', get_lr),
                ('is_ynet_bias', get_bool_bias), ('is_augment', get_bool_aug)]


def _runs(msgs):
    return re.split(_RUN, msgs)[1:]


def _table(rows, columns, casts, name_column, order):
    df = pd.DataFrame(rows, columns=columns)
    for col, kind in casts.items():
        df[col] = df[col].astype(kind)
    for col, fn in _NAME_FIELDS:
        df[col] = df[name_column].apply(fn)
    return df.reindex(columns=order)


def extract_train_msg(msgs):
    """extract_log.py:8-42: one row per training run."""
    rows = [dict(seed=_field(m, _SEED), pretrained_ckpt=_field(m, _PRETRAINED, post=_basename),
                 experiment=_field(m, r'Experiment (.*?) has started'),
                 n_param=_field(m, r'The number of trainable parameters: ([\d]+)', default=0),
                 n_epoch=_field(m, r'Early stop at epoch ([\d]+)', default=99),
                 ade=_field(m, _METRIC, 1), fde=_field(m, _METRIC, 2)) for m in _runs(msgs)]
    return _table(rows, ['seed', 'pretrained_ckpt', 'experiment', 'n_param', 'n_epoch', 'ade', 'fde'],
                  dict(seed=int, n_param=int, n_epoch=int, ade=float, fde=float), 'experiment',
                  ['seed', 'train_net', 'n_train', 'position', 'n_param', 'n_epoch', 'lr', 'is_ynet_bias', 'is_augment',
                   'ade', 'fde', 'experiment', 'pretrained_ckpt'])


def extract_test_msg(test_msg):
    """extract_log.py:45-73: one row per evaluation run of a tuned checkpoint."""
    rows = [dict(seed=_field(m, _SEED), pretrained_ckpt=_field(m, _PRETRAINED, post=_basename),
                 tuned_ckpt=_field(m, r"'tuned_ckpt': '(.*?)',", post=_basename),
                 ade=_field(m, _METRIC, 1), fde=_field(m, _METRIC, 2)) for m in _runs(test_msg)]
    return _table(rows, ['seed', 'pretrained_ckpt', 'tuned_ckpt', 'ade', 'fde'], dict(seed=int, ade=float, fde=float),
                  'tuned_ckpt', ['seed', 'train_net', 'n_train', 'position', 'lr', 'is_ynet_bias', 'is_augment', 'ade', 'fde',
                                 'tuned_ckpt', 'pretrained_ckpt'])


def extract_imp_msg(imp_msg):
    """extract_log.py:76-97: layer-importance runs -- one row per 'Replacing <layer>' line and its test result (strings,
    as the reference leaves them)."""
    rows = []
    for m in _runs(imp_msg):
        layers = re.findall('Replacing (.*?)\n', m)
        metrics = re.findall(_METRIC, m)
        if len(layers) != len(metrics):
            raise ValueError(f'{len(layers)} replaced layers but {len(metrics)} results in one run')
        common = dict(seed=_field(m, _SEED), tuned_ckpt=_field(m, r"'tuned_ckpts': \['(.*?)'\],"),
                      pretrained_ckpt=_field(m, _PRETRAINED))
        rows += [dict(layer=layer, ade=ade, fde=fde, **common) for layer, (ade, fde) in zip(layers, metrics)]
    return pd.DataFrame(rows, columns=['seed', 'layer', 'ade', 'fde', 'tuned_ckpt', 'pretrained_ckpt'])


def extract_file(file_path, out_dir):
    """extract_log.py:149-166: the kind of log is read off the file name ('eval' / 'train' / 'imp')."""
    with open(file_path, 'r') as f:
        msgs = f.read()
    for key, parse in (('eval', extract_test_msg), ('train', extract_train_msg), ('imp', extract_imp_msg)):
        if key in file_path:
            df = parse(msgs)
            break
    else:
        raise NotImplementedError
    pathlib.Path(out_dir).mkdir(parents=True, exist_ok=True)
    file_name = re.search('/([^/]+).out', file_path).group(1) if '/' in file_path else file_path.replace('.out', '')
    out_name = f'{out_dir}/{file_name}.csv'
    print(f'Saved {out_name}')
    df.to_csv(out_name, index=False)


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--file_path', default=None, type=str)
    parser.add_argument('--out_dir', default='csv/log', type=str)
    args = parser.parse_args()
    extract_file(args.file_path, args.out_dir)
