"""``python -m motion_style_transfer_b200.test <flags of the reference's test.py>`` (test.py:1-55): ADE / FDE of a whole
checkpoint (``--ckpts``) or of a pretrained checkpoint plus separately saved tuned parameters (``--pretrained_ckpt`` +
``--tuned_ckpt``) on the held-out agents."""
import time

from . import parallel
from .utils.data_utils import prepare_dataeset, set_random_seeds
from .utils.parser import get_parser
from .utils.util import get_ckpts_and_names, get_image_and_data_path, get_params, restore_model


def main(args):
    tic = time.time()
    parallel.init_from_env()           # under torchrun: one process per GPU, agents sharded (parallel.py)
    set_random_seeds(args.seed)
    params = get_params(args)
    image_path, data_path = get_image_and_data_path(params)
    _, _, df_test = prepare_dataeset(data_path, args.load_data, args.batch_size, None, None, args.val_files, args.val_split,
                                     args.test_splits, args.shuffle, args.share_val_test, 'eval', args.show_details)

    ckpts, names, separated = get_ckpts_and_names(args.ckpts, args.ckpts_name, args.pretrained_ckpt, [args.tuned_ckpt])
    print(ckpts, names)
    # test.py:30-40: with several checkpoints the LAST one that is not the pretrained baseline ('OODG') is the one tested
    model = None
    for ckpt, name, sep in zip(ckpts, names, separated):
        if len(names) == 1 or name != 'OODG':
            model = restore_model(params, sep, ckpts[0] if sep else ckpt, ckpt if sep else None)
    if args.backend is not None:
        model.model.set_backend(args.backend)

    print('############ Test model ##############')
    set_random_seeds(args.seed)
    model.test(df_test, image_path)
    print('Time spent:', time.strftime("%Hh%Mm%Ss", time.gmtime(time.time() - tic)))


if __name__ == '__main__':
    main(get_parser(False).parse_args())
