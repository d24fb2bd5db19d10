"""Pretraining and fine-tuning (MoSA / adapters / encoder / all) from the command line -- the flags, the printed lines and
the files written are those of the reference's ``train.py`` (1-79), so its ``scripts/*/{pretrain,tune_*}.sh`` run after
replacing ``python train.py`` by

    python -m motion_style_transfer_b200.train <same flags> [--backend fp32|bf16x3|bf16]

from a directory that holds ``config/`` and ``data/``.  ``utils/extract_log.py`` scrapes the output.
"""
import os
import time

from . import parallel
from .utils import data_utils, util
from .utils.parser import get_parser


def _new_trainer(params, backend):
    from .models.trainer import YNetTrainer
    trainer = YNetTrainer(params=params)
    if backend is not None:
        trainer.model.set_backend(backend)
    return trainer


def _initialization_check(args, params, adapted, df_test, image_dir):
    """train.py:45-60: freshly inserted adapters (LoRA B = 0, identity-initialised adapter convs) must leave the forecasts
    of the pretrained network unchanged -- same seed, same test agents, equal ADE and FDE to the last bit."""
    plain = _new_trainer({**params, 'position': []}, args.backend)
    plain.load_params(args.pretrained_ckpt)
    scores = []
    for trainer in (plain, adapted):
        data_utils.set_random_seeds(args.seed)
        ade, fde, _, _ = trainer.test(df_test, image_dir)
        scores.append((ade, fde))
    if scores[0] != scores[1]:
        raise RuntimeError('Wrong model initialization')
    print('Passed initialization check')


def main(args):
    started = time.time()
    parallel.init_from_env()           # under torchrun: one process per GPU, agents sharded (parallel.py)
    data_utils.set_random_seeds(args.seed)
    if args.gpu:                        # (as in the reference, device 0 needs no flag)
        os.environ['CUDA_VISIBLE_DEVICES'] = str(args.gpu)
    params = util.get_params(args)
    image_dir, data_dir = util.get_image_and_data_path(params)
    df_train, df_val, df_test = data_utils.prepare_dataeset(
        data_dir, args.load_data, args.batch_size, args.n_train_batch, args.train_files, args.val_files, args.val_split,
        args.test_splits, args.shuffle, args.share_val_test, 'train', args.show_details)
    run_name = util.get_experiment_name(args, df_train.metaId.unique().shape[0])
    print(f"Experiment {run_name} has started")

    trainer = _new_trainer(params, args.backend)
    if args.pretrained_ckpt is None:
        print("Training from scratch")
    else:
        trainer.load_params(args.pretrained_ckpt)
        print(f"Loaded checkpoint {args.pretrained_ckpt}")
    if args.init_check:
        _initialization_check(args, params, trainer, df_test, image_dir)

    print('############ Train model ##############')
    trainer.train(df_train, df_val, image_dir, image_dir, run_name)
    print('############ Test leftout data ##############')
    data_utils.set_random_seeds(args.seed)
    trainer.test(df_test, image_dir)
    print('Time spent:', time.strftime("%Hh%Mm%Ss", time.gmtime(time.time() - started)))


if __name__ == '__main__':
    main(get_parser(True).parse_args())
