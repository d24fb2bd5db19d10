"""Per-agent ADE / FDE of several checkpoints side by side (evaluator/evaluate_multickpts.py:1-108):

    python -m motion_style_transfer_b200.evaluator.evaluate_multickpts --config_filename inD_shortterm_eval.yaml \\
        --dataset_path filter/shortterm/agent_type/scene1/pedestrian_filter_s1_t524 --load_data predefined --network original \\
        --pretrained_ckpt ckpts/sdd__ynet__ped.pt --tuned_ckpts ckpts/.../...mosa_2...pt ckpts/.../...all...pt --n_round 3

writes ``csv/comparison/<seed>__<dataset path>[__<val files>]/<names>__N<agents>_R<rounds>.csv`` with one row per agent and the
columns ``metaId, sceneId, ade_<name>, fde_<name>`` per checkpoint (means over the rounds) -- the table
``get_meta_ids_focus(given_csv=...)`` reads back to pick the agents on which two models differ most.  Every checkpoint is
evaluated under the same seed (same draws of the device generator), so the columns are paired samples.  ``--viz`` (matplotlib
overlays, evaluator/visualization.py) is outside the path and refused.
"""
import pathlib

from .. import parallel
from ..utils.data_utils import get_meta_ids_focus, prepare_dataeset, set_random_seeds
from ..utils.parser import get_parser
from ..utils.util import get_ckpts_and_names, get_image_and_data_path, get_params, restore_model


def mean_over_rounds(list_metrics, ckpt_name):
    """evaluate_multickpts.py:50-57: the per-agent mean of ade / fde over the rounds, columns tagged with the model name."""
    table = list_metrics[0].copy()
    for more in list_metrics[1:]:
        table[['ade', 'fde']] = more[['ade', 'fde']] + table[['ade', 'fde']]
    table[['ade', 'fde']] = table[['ade', 'fde']] / len(list_metrics)
    return table.rename({'ade': f'ade_{ckpt_name}', 'fde': f'fde_{ckpt_name}'}, axis=1)


def main(args):
    if args.viz:
        raise NotImplementedError('--viz draws matplotlib overlays (evaluator/visualization.py): not part of the B200 path')
    parallel.init_from_env()           # under torchrun: one process per GPU, agents sharded (parallel.py)
    set_random_seeds(args.seed)
    params = get_params(args)
    image_path, data_path = get_image_and_data_path(params)
    _, _, df_test = prepare_dataeset(data_path, args.load_data, args.batch_size, None, None, args.val_files, args.val_split,
                                     args.test_splits, args.shuffle, args.share_val_test, 'eval', show_details=False)
    focus = get_meta_ids_focus(
        df_test, given_csv={'path': args.result_path, 'name': args.result_name, 'n_limited': args.result_limited},
        given_meta_ids=args.given_meta_ids, random_n=args.random_n)
    df_test = df_test[df_test.metaId.isin(focus)]
    print('meta_ids_focus: #=', len(focus))
    print(f"df_test_limited: {df_test.shape}; #={df_test.metaId.unique().shape[0]}")

    ckpts, names, separated = get_ckpts_and_names(args.ckpts, args.ckpts_name, args.pretrained_ckpt, args.tuned_ckpts)
    result = None
    for ckpt, name, sep in zip(ckpts, names, separated):
        print(f'====== Testing for {name} ======')
        model = restore_model(params, sep, ckpts[0] if sep else ckpt, ckpt if sep else None)
        if getattr(args, 'backend', None) is not None:
            model.model.set_backend(args.backend)
        set_random_seeds(args.seed)
        _, _, list_metrics, _ = model.test(df_test, image_path, True, False)
        table = mean_over_rounds(list_metrics[:args.n_round], name)
        result = table if result is None else result.merge(table, on=['metaId', 'sceneId'])

    folder = f"{args.seed}__{'_'.join(args.dataset_path.split('/'))}"
    if args.val_files is not None:
        folder += f"__{'_'.join(args.val_files).rstrip('.pkl')}"       # (rstrip of a character set, as in the reference)
    out_dir = f'csv/comparison/{folder}'
    pathlib.Path(out_dir).mkdir(parents=True, exist_ok=True)
    out_name = f"{out_dir}/{'_'.join(names)}__N{len(focus)}_R{args.n_round}.csv"
    result.to_csv(out_name, index=False)
    print(f'Saved {out_name}')
    return result


def get_multickpts_parser():
    """evaluate_multickpts.py:93-103: the flags of test.py plus the choice of agents."""
    parser = get_parser(False)
    parser.add_argument('--given_meta_ids', default=None, type=int, nargs='+')
    parser.add_argument('--result_path', default=None, type=str)
    parser.add_argument('--result_name', default=None, type=str)
    parser.add_argument('--result_limited', default=None, type=int)
    parser.add_argument('--random_n', default=None, type=int)
    parser.add_argument('--viz', action='store_true')
    return parser


if __name__ == '__main__':
    main(get_multickpts_parser().parse_args())
