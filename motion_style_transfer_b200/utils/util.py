"""Run bookkeeping of the reference's scripts (utils/util.py:1-147): the parameter dictionary from YAML + flags, the
data / image directories, experiment and checkpoint names, and restoring a model from a pretrained checkpoint plus the
separately saved tuned parameters.  The checkpoint-name grammar is what ``utils/extract_log.py`` parses back:

    Seed_<s>__<dataset path with _>__<train_net>[__Pos_<p0>_<p1>...][__TrN_<n>__lr_<lr>[__smooth][__early_<k>][__AUG][__bias]]__<network | fusion_<n>>
"""
import os

import numpy as np
import yaml


def get_experiment_name(args, n_data):
    """util.py:7-31."""
    parts = [f'Seed_{args.seed}']
    where = args.dataset_path.replace('/', '_')
    if args.load_data == 'sequential':
        where += '_' + '_'.join(f.replace('.pkl', '') for f in args.train_files)
    parts += [where, str(args.train_net)]
    if args.position != []:
        parts.append('Pos_' + '_'.join(map(str, args.position)))
    if args.n_train_batch is not None:
        parts += [f'TrN_{n_data}', 'lr_' + np.format_float_positional(args.lr, trim='-')]
        if args.smooth_val:
            parts.append('smooth')
        if args.n_early_stop < args.n_epoch:
            parts.append(f'early_{args.n_early_stop}')
        if args.augment:
            parts.append('AUG')
        if args.ynet_bias:
            parts.append('bias')
    parts.append(args.network if args.network in ('original', 'embed') else f'fusion_{args.n_fusion}')
    return '__'.join(parts)


def _dataset_kind(params):
    name = params['dataset_name'].lower()
    for kind in ('sdd', 'ind'):
        if kind in name:
            return kind
    raise ValueError(f'Invalid {name}')


def get_params(args):
    """util.py:34-60: ``config/<config_filename>`` (relative to the working directory, like the reference) overlaid with the
    command-line flags; the segmentation checkpoint is chosen by dataset."""
    if args.network == 'fusion':
        assert args.n_fusion is not None
    with open(os.path.join('config', args.config_filename)) as file:
        params = yaml.load(file, Loader=yaml.FullLoader)
    seg = {'sdd': 'sdd_segmentation.pth', 'ind': 'inD_segmentation.pth'}[_dataset_kind(params)]
    params['segmentation_model_fp'] = os.path.join(params['data_dir'], params['dataset_name'], seg)
    n = getattr(args, 'n_train_batch', None)
    if n is not None and int(n) == n:                 # 2.0 -> 2 (the flag is a float: 0.5 batches are allowed)
        args.n_train_batch = int(n)
    params.update(vars(args))
    print(params)
    return params


def get_image_and_data_path(params):
    """util.py:63-76."""
    root = os.path.join(params['data_dir'], params['dataset_name'])
    sub = {'sdd': ('raw', 'annotations'), 'ind': ('images',)}[_dataset_kind(params)]
    image_path = os.path.join(root, *sub)
    assert os.path.isdir(image_path), f'image dir error: {image_path}'
    data_path = os.path.join(root, params['dataset_path'])
    assert os.path.isdir(data_path), f'data dir error: {data_path}'
    return image_path, data_path


def _name_fields(ckpt_path):
    """The fields of a checkpoint file name (grammar in the module docstring) that the helpers below need."""
    name = ckpt_path.split('/')[-1]
    position = name.split('Pos_')[-1].split('__')[0] if 'Pos' in name else None
    return dict(name=name, train_net=name.split('__')[2], position=position)


def get_position(ckpt_path, return_list=True):
    """util.py:79-91: the ``Pos_0_1_2`` field of a checkpoint name -> '0_1_2' or ['0', '1', '2']; None without one (or
    without a path)."""
    if ckpt_path is None or 'Pos' not in ckpt_path:
        return None
    pos = ckpt_path.split('Pos_')[-1].split('__')[0]
    return pos.split('_') if return_list else pos


def get_ckpt_name(ckpt_path):
    """util.py:94-104: 'mosa_1[0_1_2](20)' / 'all(20)' -- the column name of a tuned checkpoint in result tables: what was
    trained, where, on how many agents."""
    f = _name_fields(ckpt_path)
    n_train = int(f['name'].split('TrN_')[-1].split('_')[0])
    where = '' if f['position'] is None else f"[{f['position']}]"
    return f"{f['train_net']}{where}({n_train})"


def update_params(ckpt_path, params):
    """util.py:107-123: the constructor arguments a separately saved set of tuned parameters needs, read back from its file
    name: ``train_net`` and ``position`` decide where the adapter tensors live in the module tree."""
    f = _name_fields(ckpt_path)
    updated = dict(params, train_net=f['train_net'].split('.')[0])
    base_arch = params['pretrained_ckpt'].split('_')[-1].split('.')[0]
    if base_arch == 'embed':
        updated['add_embedding'] = True
    elif 'fusion' in base_arch:       # (unreachable with the name grammar above: the last '_' field is the number)
        updated['n_fusion'] = int(base_arch.split('_')[-1])
    if f['position'] is not None:
        updated['position'] = f['position'].split('_')
    return updated


def get_ckpts_and_names(ckpts, ckpts_name, pretrained_ckpt, tuned_ckpts):
    """util.py:126-136: (paths, display names, is the file a set of separately saved tuned parameters)."""
    if ckpts is not None:
        return ckpts, ckpts_name, [False] * len(ckpts)
    if pretrained_ckpt is None:
        raise ValueError('No checkpoint provided')
    return ([pretrained_ckpt] + tuned_ckpts, ['OODG'] + [get_ckpt_name(c) for c in tuned_ckpts],
            [False] + [True] * len(tuned_ckpts))


def restore_model(params, is_file_separated, base_ckpt, separated_ckpt=None):
    """util.py:139-147."""
    from ..models.trainer import YNetTrainer
    if not is_file_separated:
        model = YNetTrainer(params=params)
        model.load_params(base_ckpt)
    else:
        model = YNetTrainer(params=update_params(separated_ckpt, params))
        model.load_separated_params(base_ckpt, separated_ckpt)
    return model
