"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two dataset-preparation functions of the reference that cannot be
run live in this image (pandas 3.0 dropped the behaviour of ``groupby.apply`` they rely on: the grouping column is no
longer part of the frame handed to the function, /root/reference/utils/data_utils.py:51-112 raise AttributeError).

Per-agent Python loops following the reference line by line.  Parity status: UNPINNED against the live reference (it
cannot execute); pinned against hand-derived cases in tests/test_data_utils.py.  Only tests may import this module.
"""
import pandas as pd


def sliding_window(df, window_size, stride):
    """data_utils.py:51-78.  ``groupby(['metaId'])`` visits the agents in ascending metaId; for each, chunk i takes the
    rows [i*stride, i*stride + window_size) and the label '<metaId>_<i>' (`:51-61`); labels are factorised in order of
    appearance (`:75`), the index is reset (`:77`)."""
    pieces, labels = [], []
    for meta_id, x in df.groupby('metaId', sort=True):
        n_chunk = (len(x) - window_size) // stride + 1
        for i in range(n_chunk):
            pieces.append(x.iloc[i * stride:i * stride + window_size])
            labels += ['{}_{}'.format(meta_id, i)] * window_size
    out = pd.concat(pieces) if pieces else df.iloc[:0]
    out = out.copy()
    out['metaId'] = pd.factorize(pd.Series(labels, dtype=object), sort=False)[0]
    return out.reset_index(drop=True)


def split_fragmented(df):
    """data_utils.py:81-112.  frame_diff = per-agent frame difference, 1 for the first row (`:103`); every row with
    frame_diff != 1 starts a fragment: from that row's label to the end of the agent the id becomes '<metaId>_<counter>',
    counter = 0, 1, ... (`:81-90`); ids factorised in order of appearance (`:110`); the frame keeps its order, its index
    and the ``frame_diff`` column."""
    out = df.copy()
    out['frame_diff'] = out.groupby('metaId')['frame'].diff().fillna(value=1.0).to_numpy()
    new_id = out['metaId'].astype(object).copy()
    for meta_id, x in out.groupby('metaId', sort=True):
        counter = 0
        for label in x.index[x['frame_diff'] != 1.0]:
            rest = x.loc[label:].index
            new_id.loc[rest] = '{}_{}'.format(meta_id, counter)
            counter += 1
    out['metaId'] = pd.factorize(new_id, sort=False)[0]
    return out
