"""SURVEY 8f rank 4, the host side of the reference's train.py / test.py: dataset splits and sample limits
(utils/data_utils.py:14-112,754-964), flags (utils/parser.py), run / checkpoint names and parameter dictionaries
(utils/util.py), and the log scraper (utils/extract_log.py) -- against ``tests/golden/scripts_host.json``, recorded from
the LIVE reference by oracle/gen_golden.py::gen_scripts_host (same inputs, numpy's global generator seeded the same way:
the selected agents, the printed lines and the scraped CSV bytes have to be identical)."""
import contextlib
import io
import json
import os
import random

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope='module')
def G():
    with open(os.path.join(GOLDEN, 'scripts_host.json')) as f:
        return json.load(f)


@pytest.fixture(scope='module')
def frames(G):
    return {k: pd.DataFrame(v) for k, v in G['frames'].items()}


def xs(df):
    return None if df is None else [float(v) for v in df.x.values]


def captured(fn, *a, **kw):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = fn(*a, **kw)
    return res, buf.getvalue()


def test_dataset_split_by_ratio_selects_the_reference_agents(G, frames):
    from motion_style_transfer_b200.utils import data_utils as D
    for case in G['split_by_ratio']:
        kw = dict(case['kw'])
        if 'given_test_meta_ids' in kw:
            kw['given_test_meta_ids'] = np.array(kw['given_test_meta_ids'])
        np.random.seed(case['seed'])
        parts, printed = captured(D.dataset_split_by_ratio, frames['main'], **kw)
        assert [xs(p) for p in parts] == case['parts'], case['kw']
        assert printed == case['printed']
        assert float(np.random.rand()) == case['next_random']        # the global generator was advanced identically
    # train | val | test partition the agents when the three are independent
    tr, va, te = D.dataset_split_by_ratio(frames['main'], 0.1, 0.2)
    ids = [set(p.metaId) for p in (tr, va, te)]
    assert not (ids[0] & ids[1]) and not (ids[0] & ids[2]) and not (ids[1] & ids[2])
    assert ids[0] | ids[1] | ids[2] == set(frames['main'].metaId)


def test_limit_downsample_filter(G, frames):
    from motion_style_transfer_b200.utils import data_utils as D
    df = frames['main']
    c = G['limit_samples']
    np.random.seed(c['seed'])
    assert xs(D.limit_samples(df, c['num'], c['batch_size'])) == c['xs']
    assert xs(D.limit_samples(df, 2, c['batch_size'], False)) == c['xs_ordered']
    assert D.limit_samples(df, None, 4) is df
    assert xs(D.downsample(df, G['downsample']['step'])) == G['downsample']['xs']
    assert np.array_equal(D.mask_step(np.arange(7), 3), [True, False, False, True, False, False, True])
    c = G['filter_short']
    assert xs(D.filter_short_trajectories(df.drop(index=c['dropped']), c['threshold'])) == c['xs']
    # reduce_df_meta_ids: rows of the given agents in the frame's own order; nothing for no agents
    some = np.array(sorted(set(df.metaId))[:3][::-1])
    got = D.reduce_df_meta_ids(df, some)
    assert got.equals(df[df.metaId.isin(some)])
    assert len(D.reduce_df_meta_ids(df, np.array([], dtype=np.int64))) == 0


def test_prepare_dataeset_and_file_splits(G, frames, tmp_path):
    from motion_style_transfer_b200.utils import data_utils as D
    for name, f in frames.items():
        if name != 'main':
            f.to_pickle(tmp_path / name)
    for case in G['prepare']:
        np.random.seed(case['seed'])
        parts, printed = captured(D.prepare_dataeset, str(tmp_path), *case['args'])
        assert [xs(p) for p in parts] == case['parts'], case['args']
        assert printed == case['printed'], case['args']
    c = G['split_randomly']
    _, printed = captured(D.split_train_val_test_randomly, str(tmp_path), 'Biker.pkl', c['val_split'], c['test_split'],
                          seed=c['seed'])
    assert printed == c['printed']
    assert [xs(pd.read_pickle(tmp_path / 'Biker' / f'{n}.pkl')) for n in ('train', 'val', 'test')] == c['parts']
    # the command line of scripts/inD/preprocessing.sh writes the same three files
    import shutil
    from motion_style_transfer_b200.utils import split_dataset
    shutil.rmtree(tmp_path / 'Biker')
    captured(split_dataset.main, ['--data_dir', str(tmp_path), '--data_filename', 'Biker.pkl', '--val_split', str(c['val_split']),
                                  '--test_split', str(c['test_split']), '--seed', str(c['seed'])])
    assert [xs(pd.read_pickle(tmp_path / 'Biker' / f'{n}.pkl')) for n in ('train', 'val', 'test')] == c['parts']
    c = G['given_scenes']
    sel, printed = captured(D.dataset_split_given_scenes, str(tmp_path), ['Biker.pkl', 'Car.pkl'], c['scenes'])
    assert xs(sel) == c['xs'] and printed == c['printed']
    # error behaviour of the reference: unknown mode, different train / val files, more samples than agents
    with pytest.raises(NotImplementedError):
        D.prepare_dataeset(str(tmp_path), 'sequential', 4, None, ['Biker.pkl'], ['Car.pkl'], 0.1, [5], False, False, 'train')
    with pytest.raises(NotImplementedError):
        D.prepare_dataeset(str(tmp_path), 'sequential', 4, None, None, ['Car.pkl'], 0.1, [5], False, False, 'plot')
    with pytest.raises(AssertionError, match='Training set size'):
        D.load_predefined_train_val_test(str(tmp_path), batch_size=16, n_train_batch=2)
    with pytest.raises(AssertionError, match='No val file'):
        D.prepare_dataeset(str(tmp_path), 'sequential', 4, None, None, None, 0.1, [5], False, False, 'eval')


def _gappy_frame():
    rows = []
    for meta, frames_ in ((5, list(range(15)) + list(range(20, 30)) + list(range(33, 48))), (2, list(range(7))),
                          (9, list(range(100, 125)))):
        for i, fr in enumerate(frames_):
            rows.append(dict(frame=fr, trackId=meta, x=float(meta * 1000 + i), y=0.5, sceneId='a', metaId=meta))
    return pd.DataFrame(rows)


def test_sliding_window_and_split_fragmented():
    """The reference's own two functions raise on pandas 3 (groupby.apply no longer passes the grouping column): pinned
    against the per-agent restatement in oracle/data_oracle.py and against cases worked out by hand."""
    from motion_style_transfer_b200.utils import data_utils as D
    from oracle import data_oracle as O
    df = _gappy_frame()
    for window, stride in ((8, 4), (20, 20), (5, 1), (41, 3)):
        got = D.sliding_window(df, window, stride)
        pd.testing.assert_frame_equal(got, O.sliding_window(df, window, stride), check_dtype=False)
        assert (got.groupby('metaId').size() == window).all()
    got = D.sliding_window(df, 8, 4)
    # agents in ascending metaId (2: 7 rows, no window; 5: 40 rows, 9 windows; 9: 25 rows, 5 windows), chunks in time order
    assert got.metaId.nunique() == 14 and list(got.metaId.unique()) == list(range(14))
    assert list(got.x[:8]) == [5000.0 + i for i in range(8)] and list(got.x[8:16]) == [5004.0 + i for i in range(8)]
    assert list(got.x[9 * 8:9 * 8 + 2]) == [9000.0, 9001.0]
    assert list(got.index) == list(range(len(got)))
    assert len(D.sliding_window(df, 41, 3)) == 0
    before = df.copy()
    got = D.split_fragmented(df)
    pd.testing.assert_frame_equal(df, before)                       # the caller's frame is not modified
    pd.testing.assert_frame_equal(got, O.split_fragmented(df), check_dtype=False)
    # agent 5 has gaps after 15 and 25 rows -> three agents; ids renumbered in order of appearance; rows keep their place
    assert list(got.metaId.unique()) == [0, 1, 2, 3, 4]
    assert list(got.metaId[:15]) == [0] * 15 and list(got.metaId[15:25]) == [1] * 10 and list(got.metaId[25:40]) == [2] * 15
    assert list(got.metaId[40:47]) == [3] * 7 and list(got.metaId[47:]) == [4] * 25
    assert list(got.frame_diff[[15, 25]]) == [6.0, 4.0] and (got.frame_diff.drop([15, 25]) == 1.0).all()
    assert list(got.x) == list(df.x)
    # shuffled row order within the frame (agents interleaved): same per-agent result
    mixed = df.sample(frac=1.0, random_state=0).sort_values('frame', kind='stable')
    pd.testing.assert_frame_equal(D.split_fragmented(mixed), O.split_fragmented(mixed), check_dtype=False)
    pd.testing.assert_frame_equal(D.sliding_window(mixed, 8, 4), O.sliding_window(mixed, 8, 4), check_dtype=False)


def test_parser_flags_and_names(G):
    from motion_style_transfer_b200.utils import parser as P, util as U, extract_log as X
    for case in G['parser']:
        args = P.get_parser(case['is_train']).parse_args(case['argv'].split())
        got = dict(vars(args))
        assert got.pop('backend') is None                      # the one flag the reference does not have
        assert got == case['namespace'], case['argv']
        for n, name in case.get('experiment', {}).items():
            assert U.get_experiment_name(args, int(n)) == name
    assert P.get_parser(True).parse_args(['--backend', 'bf16x3']).backend == 'bf16x3'
    for c in G['ckpt_names']:
        p = c['path']
        assert U.get_ckpt_name(p) == c['name']
        assert U.get_position(p) == c['position'] and U.get_position(p, return_list=False) == c['position_str']
        assert (X.get_train_net(p), X.get_n_train(p), X.get_lr(p), X.get_bool_bias(p), X.get_bool_aug(p)) == \
            (c['train_net'], c['n_train'], c['lr'], c['bias'], c['aug'])
    assert U.get_position(None) is None
    assert all(f(None) is None for f in (X.get_train_net, X.get_n_train, X.get_lr, X.get_bool_bias, X.get_bool_aug))
    for c in G['update_params']:
        assert U.update_params(c['tuned'], c['params']) == c['updated']
    for c in G['ckpts_and_names']:
        assert [list(r) for r in U.get_ckpts_and_names(*c['args'])] == c['result']
    with pytest.raises(ValueError, match='No checkpoint provided'):
        U.get_ckpts_and_names(None, None, None, [None])


def test_get_params_and_paths(G, tmp_path, monkeypatch):
    from motion_style_transfer_b200.utils import parser as P, util as U
    c = G['get_params']
    for d in ('config', 'data/sdd/raw/annotations', 'data/sdd/p/q'):
        os.makedirs(tmp_path / d)
    (tmp_path / 'config' / 'sdd_shortterm_eval.yaml').write_text(c['yaml'])
    monkeypatch.chdir(tmp_path)
    args = P.get_parser(False).parse_args(c['argv'].split())
    del args.backend
    params, printed = captured(U.get_params, args)
    assert params == c['params'] and printed == c['printed']
    assert list(U.get_image_and_data_path(params)) == c['paths']
    with pytest.raises(AssertionError, match='data dir error'):
        U.get_image_and_data_path({**params, 'dataset_path': 'nowhere'})
    with pytest.raises(ValueError, match='Invalid'):
        U.get_image_and_data_path({**params, 'dataset_name': 'eth'})
    args = P.get_parser(True).parse_args(c['argv'].split() + ['--n_train_batch', '2'])
    assert args.n_train_batch == 2.0 and isinstance(args.n_train_batch, float)
    captured(U.get_params, args)
    assert args.n_train_batch == 2 and isinstance(args.n_train_batch, int)          # util.py:52-56
    args = P.get_parser(True).parse_args(c['argv'].split() + ['--n_train_batch', '0.5'])
    captured(U.get_params, args)
    assert args.n_train_batch == 0.5


def test_extract_log_csv_bytes(G, tmp_path):
    from motion_style_transfer_b200.utils import extract_log as X
    for name, c in G['logs'].items():
        (tmp_path / f'{name}.out').write_text(c['text'])
        _, printed = captured(X.extract_file, str(tmp_path / f'{name}.out'), str(tmp_path / 'csv'))
        assert printed == f"Saved {tmp_path / 'csv'}/{name}.csv\n"
        assert (tmp_path / 'csv' / f'{name}.csv').read_text() == c['csv'], name
    with pytest.raises(NotImplementedError):
        (tmp_path / 'other.out').write_text('x')
        X.extract_file(str(tmp_path / 'other.out'), str(tmp_path / 'csv'))
    # two-digit seeds (the reference's pattern cannot read them)
    text = G['logs']['sdd_train']['text'].replace("'seed': 1,", "'seed': 12,")
    assert list(X.extract_train_msg(text).seed) == [12, 2, 3]


def test_get_meta_ids_focus(tmp_path, frames):
    from motion_style_transfer_b200.utils import data_utils as D
    df = frames['main']
    quiet = lambda *a, **kw: captured(D.get_meta_ids_focus, *a, **kw)[0]     # noqa: E731
    assert quiet(given_meta_ids=7) == [7] and quiet(given_meta_ids=[1, 2]) == [1, 2]
    with pytest.raises(ValueError):
        quiet(given_meta_ids='7')
    pd.DataFrame({'metaId': [1, 2, 3, 4], 'a': [1.0, 5.0, 2.0, 0.0], 'b': [0.0, 1.0, 7.0, 0.5]}).to_csv(tmp_path / 'r.csv')
    csv = dict(path=str(tmp_path / 'r.csv'), n_limited=2)
    assert list(quiet(given_csv=dict(name='a__b__diff', **csv))) == [2, 1]
    assert list(quiet(given_csv=dict(name='a__b__abs_diff', **csv))) == [3, 2]
    with pytest.raises(ValueError):
        quiet(given_csv=dict(name='a__b__ratio', **csv))
    none = dict(path=None)
    np.random.seed(5)
    ids = df.metaId.unique()
    np.random.shuffle(ids)
    np.random.seed(5)
    assert list(quiet(df=df, given_csv=none, random_n=4)) == list(ids[:4])
    assert list(quiet(df=df, given_csv=none)) == list(df.metaId.unique())


def test_set_random_seeds_rewinds_every_generator():
    from motion_style_transfer_b200.utils import data_utils as D, evaluate as E
    E._eval_calls = 5
    D.set_random_seeds(3)
    a = (float(np.random.rand()), random.random(), float(torch.rand(1)), torch.initial_seed(), E._eval_calls)
    E._eval_calls = 9
    D.set_random_seeds(3)
    b = (float(np.random.rand()), random.random(), float(torch.rand(1)), torch.initial_seed(), E._eval_calls)
    assert a == b and a[3] == 3 and a[4] == 0
    assert torch.backends.cudnn.deterministic and not torch.backends.cudnn.benchmark


def test_write_files_and_round_means(tmp_path):
    """evaluator/write_files.py and the per-agent mean over rounds of evaluator/evaluate_multickpts.py:50-57 (outputs of the
    live reference for these inputs)."""
    from motion_style_transfer_b200.evaluator import write_files as W
    from motion_style_transfer_b200.evaluator.evaluate_multickpts import mean_over_rounds
    W.write_csv(str(tmp_path / 'r'), 'a.csv', [1.23456, 2.5, 0.333333], [1, 2, 3])
    assert (tmp_path / 'r' / 'a.csv').read_text() == '0.3333\n1.356\n1.2346\n2.5\n0.3333\n'
    W.write_csv(str(tmp_path / 'r'), 'b.csv', [1.23456, 2.5, 0.333333], [1, 2, 3], 3.14159265)
    assert (tmp_path / 'r' / 'b.csv').read_text() == '3.1416\n0.3333\n1.356\n1.2346\n2.5\n0.3333\n'
    assert W.round_val(None) == 0.0 and W.round_val(2.0) == '2.0' and W.convert_to_str([1, 'a']) == [['1'], ['a']]
    assert W.get_out_dir('o', 'a/b', 1, 'mosa', ['x.pkl', 'y.pkl'], ['t.pkl']) == 'o/a/b/t/x___y/mosa/1'
    assert W.get_out_dir('o', 'a/b', 1, 'mosa', ['x.pkl']) == 'o/a/b/None/x/mosa/1'
    rounds = [pd.DataFrame({'metaId': [3, 4], 'sceneId': ['s', 's'], 'ade': [1.0, 2.0], 'fde': [4.0, 8.0]}),
              pd.DataFrame({'metaId': [3, 4], 'sceneId': ['s', 's'], 'ade': [3.0, 2.0], 'fde': [0.0, 4.0]})]
    t = mean_over_rounds(rounds, 'A')
    assert list(t.columns) == ['metaId', 'sceneId', 'ade_A', 'fde_A']
    assert list(t.ade_A) == [2.0, 2.0] and list(t.fde_A) == [2.0, 6.0] and list(rounds[0].ade) == [1.0, 2.0]
