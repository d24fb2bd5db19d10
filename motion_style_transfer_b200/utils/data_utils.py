"""Host side of the data formats either side of the path: the reference's ``utils/data_utils.py``.

On-disk format (``dataloader.py:30-39``, ``data_utils.py:859-872``): pandas pickles ``train.pkl / val.pkl / test.pkl`` (or
one ``<agent>.pkl`` per agent type split sequentially), one row per (agent, time step) with the columns
``frame, trackId, x, y, sceneId, metaId`` and ``obs_len + pred_len`` consecutive rows per ``metaId``.

Everything here is host bookkeeping on a DataFrame that is read once per run; nothing reaches the GPU.  Same names,
arguments, printed lines and -- where a function draws random numbers -- the same calls on numpy's global generator in the
same order as the reference, so a seeded run selects the same agents.  What differs is how the work is done:

  * ``reduce_df_meta_ids`` is a hash join (``np.isin``) instead of an (ids x rows) boolean matrix (`:812-813`: 64k agents x
    1.3M rows would be 84 GB);
  * ``downsample`` / ``filter_short_trajectories`` / ``sliding_window`` / ``split_fragmented`` (`:14-112`) are index
    arithmetic on the grouped frame instead of ``groupby.apply`` of a Python function per agent.  The reference's
    ``sliding_window`` and ``split_fragmented`` raise on the pandas of this image (3.0: ``groupby.apply`` no longer hands
    the grouping column to the function), so those two are pinned against ``oracle/data_oracle.py`` (per-agent loops
    restating `:51-112`) instead of the live reference; the others are pinned against the live reference.

Of the analysis half of the reference file (`:279-751`) the velocity / acceleration variation factors and the per-agent-type
datasets (what ``utils/sdd_dataset.py`` / ``inD_dataset.py`` / ``filter_dataset.py`` need to go from the raw recordings to
the pickles above) and the datasets by factor range are here; neighbour-distance factors and the plots are not.
"""
import os
import pathlib
import random

import numpy as np
import pandas as pd
import torch

from .image_utils import AUGMENT_SUFFIX, augment_data  # noqa: F401  (data_utils.py:115-233: lives with the image code)


# ------------------------------------------------------------------------------------------ preprocessing (14-112)
def mask_step(x, step):
    """data_utils.py:14-20: True at every ``step``-th position, starting from the first."""
    mask = np.zeros(len(x), dtype=bool)
    mask[::step] = True
    return mask


def downsample(df, step):
    """data_utils.py:23-33: keep every ``step``-th row of each agent (metaId), e.g. 30 fps -> 2.5 fps with step 12."""
    nth = df.groupby('metaId').cumcount().to_numpy()
    return df[nth % step == 0]


def filter_short_trajectories(df, threshold):
    """data_utils.py:36-48: drop the agents with fewer than ``threshold`` (non-null) frames."""
    n_frames = df.groupby('metaId')['frame'].transform('count').to_numpy()
    return df[n_frames >= threshold]


def _grouped_positions(df):
    """Row positions ordered by (metaId, original order), the first position of every agent in that order and the
    agents' lengths -- the frame the reference's ``groupby(['metaId'])`` iterates over."""
    codes, uniques = pd.factorize(df['metaId'], sort=True)
    order = np.argsort(codes, kind='stable')
    counts = np.bincount(codes, minlength=len(uniques))
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    return order, starts, counts


def sliding_window(df, window_size, stride):
    """data_utils.py:51-78: cut every (downsampled) trajectory into chunks of ``window_size`` rows, ``stride`` rows apart
    (overlapping when stride < window_size); every chunk becomes an agent of its own, numbered in order of appearance
    (agents by ascending metaId, chunks by time).  Agents shorter than one window vanish."""
    order, starts, counts = _grouped_positions(df)
    n_chunk = np.maximum((counts - window_size) // stride + 1, 0)
    total = int(n_chunk.sum())
    agent = np.repeat(np.arange(len(counts)), n_chunk)
    first_chunk = np.concatenate([[0], np.cumsum(n_chunk)[:-1]]).astype(np.int64)
    i_chunk = np.arange(total) - first_chunk[agent]
    rows = (starts[agent] + i_chunk * stride)[:, None] + np.arange(window_size)[None, :]
    out = df.iloc[order[rows.reshape(-1)]].copy()
    out['metaId'] = np.repeat(np.arange(total, dtype=np.int64), window_size)
    return out.reset_index(drop=True)


def split_fragmented(df):
    """data_utils.py:81-112: a trajectory with a gap (frame[t+1] - frame[t] != 1) is split there; the rows from each gap
    onwards get a new metaId, ids are renumbered in order of appearance.  Row order and index are kept; the result carries
    the helper column ``frame_diff`` like the reference's."""
    out = df.copy()
    diff = out.groupby('metaId')['frame'].diff().fillna(value=1.0).to_numpy()
    out['frame_diff'] = diff
    gap = pd.Series(diff != 1.0, index=out.index)
    piece = gap.groupby(out['metaId'].to_numpy()).cumsum().to_numpy().astype(np.int64)     # 0 before the first gap
    agent = pd.factorize(out['metaId'], sort=False)[0].astype(np.int64)
    out['metaId'] = pd.factorize(agent * (int(piece.max(initial=0)) + 1) + piece, sort=False)[0]
    return out


# ------------------------------------------------------------------------------------------ scene images (236-276)
def resize_and_pad_image(images, size, pad=2019):
    """data_utils.py:236-245: pad to a (pad x pad) square at the bottom / right, then INTER_AREA-resize to (size x size),
    in place."""
    import cv2
    for key, im in images.items():
        h, w = im.shape[:2]
        im = cv2.copyMakeBorder(im, 0, pad - h, 0, pad - w, cv2.BORDER_CONSTANT)
        images[key] = cv2.resize(im, (size, size), interpolation=cv2.INTER_AREA)


def _scene_file(image_path, scene, image_file, use_raw_data):
    if use_raw_data and image_file != 'oracle.png':          # SDD raw layout: <scene name>/video<idx>/reference.jpg
        name, idx = scene.split('_')
        return os.path.join(image_path, name, f'video{idx}', image_file)
    return os.path.join(image_path, scene, image_file)


def create_images_dict(unique_scene, image_path, image_file='reference.jpg', use_raw_data=False):
    """data_utils.py:248-263: {scene: BGR uint8 image} (``oracle.png`` semantic maps are read as one channel)."""
    import cv2
    flags = (0,) if image_file == 'oracle.png' else ()
    return {scene: cv2.imread(_scene_file(image_path, scene, image_file, use_raw_data), *flags) for scene in unique_scene}


def load_images(scenes, image_path, image_file='reference.jpg'):
    """data_utils.py:266-276."""
    return create_images_dict(set(scenes) if isinstance(scenes, list) else scenes, image_path, image_file)


# ------------------------------------------------------------------------------------------ variation factors (279-357)
_REDUCE = {'max': np.max, 'avg': np.mean, 'min': np.min, 'tot': np.sum,
           'abs+max': lambda a, axis: np.max(np.abs(a), axis=axis), 'abs+avg': lambda a, axis: np.mean(np.abs(a), axis=axis),
           'abs+min': lambda a, axis: np.mean(np.abs(a), axis=axis)}        # ('abs+min' is a mean in the reference, :351-352)


def aggregate_per_varf_value(df, varf, obs_len):
    """data_utils.py:293-357: one statistic per agent, ``varf = '<op>_<attr>'`` with attr = vel (speed between consecutive
    rows / frame step) or acc (difference of consecutive speeds / frame step), taken over the first ``obs_len`` rows (all rows
    if falsy) and reduced by op (max / avg / min / tot / abs+...).  Returns the frame [metaId, <varf>, label], agents in
    ascending metaId.  All agents at once on an (agents, rows) array -- the data are windowed, every agent has the same
    number of rows -- instead of a Python function per agent.  The neighbour-distance factors (dist, den*) are not built."""
    op, attr = varf.split('_')
    if attr not in ('vel', 'acc'):
        raise NotImplementedError(f'variation factor {varf!r}: only the velocity / acceleration factors are built')
    if op not in _REDUCE:
        raise ValueError(f'Cannot compute {op} operation')
    order, starts, counts = _grouped_positions(df)
    if len(counts) == 0 or (counts != counts[0]).any():
        raise ValueError('aggregate_per_varf_value: every agent needs the same number of rows (windowed data)')
    T = int(counts[0])
    take = df.iloc[order]
    x, y = take['x'].to_numpy().reshape(-1, T), take['y'].to_numpy().reshape(-1, T)
    frame = take['frame'].to_numpy().reshape(-1, T)
    label = take['label'].to_numpy().reshape(-1, T)
    step = frame[:, 1:] - frame[:, :-1]
    assert (step == step[:, :1]).all() and (label == label[:, :1]).all()     # the reference's sanity checks (:306-313)
    step = step[:, :1].astype(np.float64) if step.dtype.kind != 'f' else step[:, :1]
    n = obs_len if obs_len else T
    vel = np.sqrt((x[:, :n - 1] - x[:, 1:n]) ** 2 + (y[:, :n - 1] - y[:, 1:n]) ** 2) / step
    seq = vel if attr == 'vel' else (vel[:, :n - 2] - vel[:, 1:n - 1]) / step
    stats = _REDUCE[op](np.ascontiguousarray(seq), axis=1)
    meta = take['metaId'].to_numpy().reshape(-1, T)[:, 0]
    return pd.DataFrame({'metaId': meta, varf: stats, 'label': label[:, 0]})


def get_varf_table(df, varf_list, obs_len):
    """data_utils.py:279-290: [metaId, label, sceneId, scene, <varf>...], one row per agent."""
    print('Computing variation fatcor by obs_len' if obs_len else 'Computing variation fatcor by obs_len + pred_len')
    table = df.groupby(['metaId', 'label', 'sceneId']).size().reset_index()[['metaId', 'label', 'sceneId']]
    table['scene'] = table.sceneId.apply(lambda s: s.split('_')[0])
    for varf in varf_list:
        table = table.merge(aggregate_per_varf_value(df, varf, obs_len)[['metaId', varf]], on='metaId')
    return table


# ------------------------------------------------------------------------------------------ datasets per agent type (367-413)
def convert_df_to_dict(df_gb):
    """data_utils.py:367-373: {group: {'metaId': [...], 'sceneId': [...], 'label': [...]}}, one entry per agent."""
    out = {}
    for key in df_gb.groups.keys():
        agents = df_gb.get_group(key)[['metaId', 'sceneId', 'label']].drop_duplicates()
        assert agents.metaId.nunique() == agents.shape[0]
        out[key] = agents.to_dict('list')
    return out


def create_dataset_by_agent_type(df, labels, out_dir, statistic_only, same_group_size=False, selected_scenes=None):
    """data_utils.py:376-413: ``<out_dir>/<label>.pkl`` per agent type, or with ``selected_scenes``
    ``<out_dir>/<scene>/<label>.pkl`` per scene plus ``<out_dir>/<scene>__<scene>.../<label>.pkl`` for their union -- the
    directories ``--dataset_path filter/.../agent_type/<scene>/`` of the training scripts point at."""
    if same_group_size:
        raise NotImplementedError('same_group_size (data_utils.py:468-517) has no caller in the reference')
    pathlib.Path(out_dir).mkdir(parents=True, exist_ok=True)
    df_label = df[df.label.isin(labels)]
    groups = df_label.groupby(by='label', dropna=True)
    rows_per_agent = df_label[df_label.metaId == df_label.metaId.unique()[0]].shape[0]
    n_agents = groups.count()['metaId'] / rows_per_agent
    print('Statistics:\n', n_agents)
    print('# total:', n_agents.sum())
    if statistic_only:
        return
    for agent, group in convert_df_to_dict(groups).items():
        rows = df_label[df_label.metaId.isin(group['metaId'])]
        if selected_scenes is None:
            rows.to_pickle(os.path.join(out_dir, f'{agent}.pkl'))
            continue
        parts = []
        for scene_id in selected_scenes:
            scene_dir = os.path.join(out_dir, scene_id)
            pathlib.Path(scene_dir).mkdir(parents=True, exist_ok=True)
            part = rows[rows.sceneId == scene_id]
            parts.append(part)
            print(f'scene_id = {scene_id}, label = {agent}, #= {part.metaId.unique().shape[0]}')
            part.to_pickle(os.path.join(scene_dir, f'{agent}.pkl'))
        union_dir = os.path.join(out_dir, '__'.join(selected_scenes))
        pathlib.Path(union_dir).mkdir(parents=True, exist_ok=True)
        union = pd.concat(parts, axis=0)
        print(f'scene_id = {selected_scenes}, label = {agent}, #= {union.metaId.unique().shape[0]}')
        union.to_pickle(os.path.join(union_dir, f'{agent}.pkl'))


def add_range_column(df, varf, varf_ranges, obs_len, inclusive='both'):
    """data_utils.py:359-364: column ``<varf>_range`` = '<lo>_<hi>' of the range the agent's factor falls into (a later
    range wins where ranges overlap, agents outside every range get NaN).  The merge renumbers the index."""
    stats = aggregate_per_varf_value(df, varf, obs_len)
    for lo, hi in varf_ranges:
        stats.loc[stats[varf].between(lo, hi, inclusive=inclusive), f'{varf}_range'] = f'{lo}_{hi}'
    return df.merge(stats[['metaId', f'{varf}_range']], on='metaId')


def create_dataset_given_range(df, varf, varf_ranges, labels, out_dir, obs_len, statistic_only, inclusive='both',
                               same_group_size=False):
    """data_utils.py:415-465: one pickle per range of a variation factor, ``<out_dir>/<lo>_<hi>.pkl`` (``varf_ranges`` a
    list of tuples, one factor) or per combination of ranges of several factors, ``<lo>_<hi>__<lo>_<hi>.pkl``
    (``varf_ranges`` a list of lists of tuples).  The 'Statistics' lines print what the reference prints (the number of
    distinct group sizes)."""
    if same_group_size:
        raise NotImplementedError('same_group_size (data_utils.py:468-517) has no caller in the reference')
    pathlib.Path(out_dir).mkdir(parents=True, exist_ok=True)
    df_label = df[df.label.isin(labels)]
    if isinstance(varf_ranges[0], tuple):
        varf = varf[0]
        df_label = add_range_column(df_label, varf, varf_ranges, obs_len, inclusive=inclusive)
        column = f'{varf}_range'
    elif isinstance(varf_ranges[0], list):
        for f, r in zip(varf, varf_ranges):
            df_label = add_range_column(df_label, f, r, obs_len, inclusive=inclusive)
        column = '__'.join(varf) + '_range'
        complete = ~df_label.isna().any(axis=1)
        df_label.loc[complete, column] = df_label.loc[complete, [f + '_range' for f in varf]].agg('__'.join, axis=1)
    else:
        raise ValueError(f'Cannot process {varf}.')
    groups = df_label.groupby(by=column, dropna=True)
    n_sizes = groups.count()['metaId'].unique().shape[0]
    print('Statistics:\n', n_sizes)
    print('# total:', n_sizes)
    if statistic_only:
        return
    for name, group in convert_df_to_dict(groups).items():
        df_label[df_label.metaId.isin(group['metaId'])].to_pickle(os.path.join(out_dir, f'{name}.pkl'))


# ------------------------------------------------------------------------------------------ splits (754-912, 955-964)
def _count(split, n):
    """A split > 1 is a number of agents, otherwise a fraction of them (data_utils.py:779,782)."""
    return int(split) if split > 1 else int(split * n)


def reduce_df_meta_ids(df, meta_ids):
    """data_utils.py:812-813: the rows of the given agents, in the frame's own order."""
    return df[np.isin(df['metaId'].to_numpy(), np.asarray(meta_ids))]


def dataset_split_by_ratio(df, val_split, test_split=None, shuffle=False, share_val_test=False, given_test_meta_ids=None):
    """data_utils.py:770-809.  Agents in ascending metaId order (shuffled by numpy's global generator on request) are cut
    into train | val | test from the front.  ``share_val_test``: validation is every ``n_test // n_val``-th test agent
    (every 3rd if that is <= 1).  Without a test split the reference hands the FIRST ``n - n_val`` agents to validation and
    the last ``n_val`` to training (`:802-804`); kept, the scripts never take that branch."""
    ids = np.unique(df['metaId'])
    if shuffle:
        print('Shuffling data')
        np.random.shuffle(ids)
    n = ids.shape[0]
    n_val = _count(val_split, n)
    if test_split is None:
        val_ids, train_ids = np.split(ids, [n - n_val])
        return reduce_df_meta_ids(df, train_ids), reduce_df_meta_ids(df, val_ids), None
    n_test = _count(test_split, n)
    if share_val_test:
        print('Share validation and test set')
        train_ids, test_ids = np.split(ids, [n - n_test])
        df_val = None
        if n_val != 0:
            every = n_test // n_val
            df_val = reduce_df_meta_ids(df, test_ids[::every if every > 1 else 3])
    else:
        print('Validation and test sets are independent')
        n_train = n - n_val - n_test
        train_ids, val_ids, test_ids = np.split(ids, [n_train, n_train + n_val])
        if given_test_meta_ids is not None:
            test_ids = given_test_meta_ids
            print('Replaced test set by given test meta ids')
        df_val = reduce_df_meta_ids(df, val_ids)
    return reduce_df_meta_ids(df, train_ids), df_val, reduce_df_meta_ids(df, test_ids)


def split_train_val_test_sequentially(data_path, train_files, val_split, test_splits=None, shuffle=False,
                                      share_val_test=False):
    """data_utils.py:754-767: one pickle per agent type, each split on its own, the parts concatenated."""
    print(f"Split {train_files} given val_split={val_split}, test_split={test_splits}")
    parts = ([], [], [])
    for train_file, test_split in zip(train_files, test_splits):
        df = pd.read_pickle(os.path.join(data_path, train_file))
        for acc, part in zip(parts, dataset_split_by_ratio(df, val_split, test_split, shuffle, share_val_test)):
            if part is not None:
                acc.append(part)
    return tuple(pd.concat(acc) if acc else pd.DataFrame([]) for acc in parts)


def dataset_split_given_scenes(data_path, files, scenes):
    """data_utils.py:816-820."""
    print(f"Split {files} given scenes={scenes}")
    df = pd.concat([pd.read_pickle(os.path.join(data_path, file)) for file in files])
    return df[df.sceneId.isin(scenes)]


def split_train_val_test_randomly(data_dir, data_filename, val_split, test_split, seed=1):
    """data_utils.py:823-856: writes train.pkl / val.pkl / test.pkl into ``<data_dir>/<data_filename minus .pkl>/``."""
    out_dir = f"{data_dir}/{data_filename.replace('.pkl', '')}"
    pathlib.Path(out_dir).mkdir(parents=True, exist_ok=True)
    df = pd.read_pickle(f'{data_dir}/{data_filename}')
    ids = np.unique(df['metaId'])
    n = ids.shape[0]
    n_val, n_test = _count(val_split, n), _count(test_split, n)
    n_train = n - n_val - n_test
    set_random_seeds(seed)
    np.random.shuffle(ids)
    parts = np.split(ids, [n_train, n_train + n_val])
    print(f'# data = {n}')
    for name, part in zip(('train', 'val', 'test'), parts):
        print(f'# {name} = {part.shape[0]}')
    for name, part in zip(('train', 'val', 'test'), parts):
        reduce_df_meta_ids(df, part).to_pickle(f'{out_dir}/{name}.pkl')
    print('Split train/val/test set')


def load_predefined_train_val_test(data_path, batch_size, n_train_batch=None, shuffle=False):
    """data_utils.py:859-872: the three pickles; ``n_train_batch`` limits training to batch_size * n_train_batch agents
    (the low-shot fine-tuning sets: ``--n_train_batch 2`` with batch_size 10 = 20 trajectories)."""
    df_train, df_val, df_test = (pd.read_pickle(f'{data_path}/{name}.pkl') for name in ('train', 'val', 'test'))
    if n_train_batch is not None:
        n_sample = int(batch_size * n_train_batch)
        ids = df_train.metaId.unique()
        n_train = ids.shape[0]
        assert n_sample <= n_train, f'Training set size ({n_train}) < Sample size ({n_sample})'
        if shuffle:
            np.random.shuffle(ids)
        df_train = reduce_df_meta_ids(df_train, ids[:n_sample])
    return df_train, df_val, df_test


def limit_samples(df, num, batch_size, random_ids=True):
    """data_utils.py:955-964: ``num`` batches worth of agents, drawn with numpy's global generator."""
    if num is None:
        return df
    ids = np.unique(df['metaId'])
    if random_ids:
        np.random.shuffle(ids)
    return reduce_df_meta_ids(df, ids[:num * batch_size])


def _describe(name, df):
    if df is not None:
        print(f"{name}: {df.shape}; #={df.metaId.unique().shape[0]}")


def prepare_dataeset(data_path, load_data, batch_size, n_train_batch, train_files, val_files, val_split, test_splits,
                     shuffle, share_val_test, mode='train', show_details=False):
    """data_utils.py:875-912 (the name is the reference's): the entry ``train.py:22-25`` / ``test.py:17-19`` call."""
    if load_data == 'predefined':
        print('Loading predefined train/val/test sets')
        df_train, df_val, df_test = load_predefined_train_val_test(data_path, batch_size=batch_size,
                                                                   n_train_batch=n_train_batch, shuffle=shuffle)
    else:
        print('Splitting train/val/test sets sequentially')
        if mode == 'train':
            assert train_files is not None, 'No train file is provided'
            assert val_files is not None, 'No val file is provided'
            assert val_split is not None, 'No val split is provided'
            if train_files != val_files:
                raise NotImplementedError
            df_train, df_val, df_test = split_train_val_test_sequentially(data_path, train_files, val_split, test_splits,
                                                                          shuffle, share_val_test)
            df_train = limit_samples(df_train, n_train_batch, batch_size)
        elif mode == 'eval':
            assert val_files is not None, 'No val file is provided'
            df_train, df_val, df_test = split_train_val_test_sequentially(data_path, val_files, val_split, test_splits,
                                                                          shuffle, share_val_test)
        else:
            raise NotImplementedError
    if show_details:
        for name, df in (('train', df_train), ('val', df_val), ('test', df_test)):
            print(f'{name}_meta_ids: {df.metaId.unique()}')
    if mode == 'train':
        _describe('df_train', df_train)
        _describe('df_val', df_val)
    _describe('df_test', df_test)
    return df_train, df_val, df_test


def get_meta_ids_focus(df=None, given_meta_ids=None, given_csv=None, random_n=None):
    """data_utils.py:914-942: which agents a visualisation / saliency run looks at: given ids, the ``n_limited`` largest
    (absolute) differences between two columns of a result CSV (``name = '<col1>__<col2>__diff|abs_diff'``), ``random_n``
    random ones, or all."""
    if given_meta_ids is not None:
        if isinstance(given_meta_ids, int):
            focus = [given_meta_ids]
        elif isinstance(given_meta_ids, list):
            focus = given_meta_ids
        else:
            raise ValueError(f'Invalid given_meta_ids={given_meta_ids}')
    elif given_csv['path'] is not None:
        col1, col2, op = given_csv['name'].split('__')
        result = pd.read_csv(given_csv['path'])
        delta = result[col1].values - result[col2].values
        if op == 'abs_diff':
            delta = np.abs(delta)
        elif op != 'diff':
            raise ValueError(f'Invalid op={op}')
        result.loc[:, 'diff'] = delta
        focus = result.sort_values(by='diff', ascending=False).head(given_csv['n_limited']).metaId.values
    elif random_n is not None:
        ids = df.metaId.unique()
        np.random.shuffle(ids)
        focus = ids[:random_n]
    else:
        focus = df.metaId.unique()
    print('Focusing on meta_ids=', focus)
    return focus


def set_random_seeds(random_seed=0):
    """data_utils.py:945-952.  Also rewinds the evaluate() call counter the device generator's stream id is derived from
    (utils/evaluate.py), so that -- as with the reference's global CPU generators -- two test rounds that each start with
    ``set_random_seeds(s)`` draw the same numbers (``train.py:49-57``, the initialisation check, relies on it)."""
    import cv2
    torch.manual_seed(random_seed)
    torch.cuda.manual_seed(random_seed)
    np.random.seed(random_seed)
    random.seed(random_seed)
    cv2.setRNGSeed(random_seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    from . import evaluate as _evaluate
    _evaluate.reset_rng_stream()
