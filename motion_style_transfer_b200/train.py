"""``python -m motion_style_transfer_b200.train <flags of the reference's train.py>`` (train.py:1-79): pretraining and
MoSA / adapter / encoder fine-tuning on the B200 engines.  Run from a directory holding ``config/`` and ``data/`` like the
reference; the printed lines are the reference's (``utils/extract_log.py`` scrapes them)."""
import os
import time

from .utils.data_utils import prepare_dataeset, set_random_seeds
from .utils.parser import get_parser
from .utils.util import get_experiment_name, get_image_and_data_path, get_params


def _trainer(params, backend):
    from .models.trainer import YNetTrainer
    trainer = YNetTrainer(params=params)
    if backend is not None:
        trainer.model.set_backend(backend)
    return trainer


def main(args):
    tic = time.time()
    set_random_seeds(args.seed)
    if args.gpu:                                       # (train.py:17: gpu 0 is the default device anyway)
        os.environ['CUDA_VISIBLE_DEVICES'] = str(args.gpu)
    params = get_params(args)
    image_path, data_path = get_image_and_data_path(params)

    df_train, df_val, df_test = prepare_dataeset(
        data_path, args.load_data, args.batch_size, args.n_train_batch, args.train_files, args.val_files, args.val_split,
        args.test_splits, args.shuffle, args.share_val_test, 'train', args.show_details)
    experiment = get_experiment_name(args, df_train.metaId.unique().shape[0])
    print(f"Experiment {experiment} has started")

    model = _trainer(params, args.backend)
    if args.pretrained_ckpt is not None:
        model.load_params(args.pretrained_ckpt)
        print(f"Loaded checkpoint {args.pretrained_ckpt}")
    else:
        print("Training from scratch")

    if args.init_check:
        # train.py:45-60: a freshly adapted model (LoRA B = 0, adapters at identity) must forecast exactly what the
        # pretrained one does under the same seed
        plain = _trainer({**params, 'position': []}, args.backend)
        plain.load_params(args.pretrained_ckpt)
        set_random_seeds(args.seed)
        ade_pre, fde_pre, _, _ = plain.test(df_test, image_path)
        set_random_seeds(args.seed)
        ade_cur, fde_cur, _, _ = model.test(df_test, image_path)
        if ade_pre != ade_cur or fde_pre != fde_cur:
            raise RuntimeError('Wrong model initialization')
        print('Passed initialization check')

    print('############ Train model ##############')
    model.train(df_train, df_val, image_path, image_path, experiment)

    print('############ Test leftout data ##############')
    set_random_seeds(args.seed)
    model.test(df_test, image_path)
    print('Time spent:', time.strftime("%Hh%Mm%Ss", time.gmtime(time.time() - tic)))


if __name__ == '__main__':
    main(get_parser(True).parse_args())
