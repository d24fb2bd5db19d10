"""The reference's command lines end to end on the B200 path (SURVEY 8f rank 4 + 8b callers): a scratch working directory
laid out like the reference's (``config/*.yaml``, ``data/sdd/raw/annotations/<scene>/video<k>/reference.jpg``, the pickled
segmentation module, ``train.pkl / val.pkl / test.pkl``), then

    train  (from scratch, 1 epoch)                          scripts/*/pretrain.sh
    train  --fine_tune --train_net mosa_1 --init_check       scripts/*/tune_mosa.sh
    test   --pretrained_ckpt ... --tuned_ckpt ...            scripts/*/generalize.sh, evaluate
    extract_log on both logs

and the numbers have to agree with each other: a freshly adapted model forecasts exactly what the pretrained one does
(``--init_check``), the tuned checkpoint restored by ``test`` reproduces the ADE / FDE ``train`` printed for the same seed,
and the scraper reads those numbers back.
"""
import os
import re

import numpy as np
import pandas as pd
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

CONFIG = dict(
    save_every_n=121, resize_factor=0.25, viz_epoch=10, encoder_channels=[8, 8, 16, 16, 16], decoder_channels=[16, 16, 16, 8, 8],
    waypoints=[5], temperature=1.0, n_semantic_classes=6, loss_scale=1000, kernlen=31, nsig=4, use_features_only=False,
    e_unfreeze=10000, use_TTST=True, rel_threshold=0.01, use_CWS=False, CWS_params='None', obs_len=5, pred_len=6, n_goal=20,
    n_traj=1, use_raw_data=True, data_dir='data/', dataset_name='sdd')
DATASET_PATH = 'filter/agent_type/Biker'
H0, W0 = 200, 264                                     # x 0.25 -> 50 x 66 -> padded to 64 x 96


def _agents(rng, first_id, n, scenes):
    rows = []
    T = CONFIG['obs_len'] + CONFIG['pred_len']
    for k in range(n):
        start = rng.uniform([40, 40], [W0 - 120, H0 - 90])
        step = rng.uniform([2, -3], [8, 6])
        wobble = rng.normal(0, 0.6, (T, 2))
        for t in range(T):
            x, y = start + step * t + wobble[t]
            rows.append(dict(frame=12 * t, trackId=first_id + k, x=float(x), y=float(y), sceneId=scenes[k % len(scenes)],
                             metaId=first_id + k))
    return pd.DataFrame(rows)


def _make_workspace(root, config, scenes, image_dir, image_name, seg_name, pickles):
    """config/tiny.yaml, one image per scene, the pickled segmentation module and the trajectory pickles."""
    import cv2
    from helpers import TinySeg
    rng = np.random.RandomState(0)
    os.makedirs(root / 'config')
    with open(root / 'config' / 'tiny.yaml', 'w') as f:
        yaml.safe_dump(config, f, sort_keys=False)
    data = root / 'data' / config['dataset_name']
    for scene in scenes:
        d = data / image_dir(scene)
        os.makedirs(d)
        assert cv2.imwrite(str(d / image_name), rng.randint(0, 256, (H0, W0, 3)).astype(np.uint8))
    torch.manual_seed(0)
    torch.save(TinySeg(6), data / seg_name)
    os.makedirs(data / DATASET_PATH)
    for name, first, n in pickles:
        _agents(rng, first, n, scenes).to_pickle(data / DATASET_PATH / f'{name}.pkl')


def _sdd_raw_dir(scene):
    name, idx = scene.split('_')
    return os.path.join('raw', 'annotations', name, f'video{idx}')


@pytest.fixture()
def workspace(tmp_path, monkeypatch, cuda_device):
    _make_workspace(tmp_path, CONFIG, ['sA_0', 'sB_1'], _sdd_raw_dir, 'reference.jpg', 'sdd_segmentation.pth',
                    (('train', 0, 12), ('val', 100, 4), ('test', 200, 6)))
    monkeypatch.chdir(tmp_path)
    return tmp_path


# inD layout (config/inD_longterm_eval.yaml: images/<scene>/reference.png, two waypoints, CWS parameters as a dictionary)
CONFIG_IND = dict(CONFIG, waypoints=[2, 5], temperature=1.8, use_TTST=True, rel_threshold=0.002, use_CWS=True,
                  CWS_params=dict(sigma_factor=6, ratio=2, rot=True), use_raw_data=False, dataset_name='inD-dataset-v1.0')


@pytest.fixture()
def workspace_ind(tmp_path, monkeypatch, cuda_device):
    _make_workspace(tmp_path, CONFIG_IND, ['scene1'], lambda scene: os.path.join('images', scene), 'reference.png',
                    'inD_segmentation.pth', (('car', 0, 16),))
    monkeypatch.chdir(tmp_path)
    return tmp_path


COMMON = f'--config_filename tiny.yaml --dataset_path {DATASET_PATH} --network original --load_data predefined --batch_size 4'
AVERAGE = r'Average performance \(by (\d+)\): \nTest ADE: ([\d\.]+) \nTest FDE: ([\d\.]+)'


def test_pretrain_finetune_test_and_scrape(workspace, capsys):
    from motion_style_transfer_b200 import train, test
    from motion_style_transfer_b200.utils import extract_log
    from motion_style_transfer_b200.utils.parser import get_parser

    # ---- pretraining from scratch (train_net = 'train': everything but the segmentation module) -------------------
    train.main(get_parser(True).parse_args(f'{COMMON} --seed 1 --n_epoch 1 --n_round 1 --ckpt_path ckpts/pre'.split()))
    out = capsys.readouterr().out
    assert 'Training from scratch' in out and 'Loading predefined train/val/test sets' in out
    assert 'df_train: (132, 6); #=12' in out and 'df_val: (44, 6); #=4' in out and 'df_test: (66, 6); #=6' in out
    experiment = re.search('Experiment (.*?) has started', out).group(1)
    assert experiment == 'Seed_1__filter_agent_type_Biker__train__original'
    pre = f'ckpts/pre/{experiment}.pt'
    saved = torch.load(pre)
    assert not any('segmentation' in k for k in saved) and 'encoder.stages.0.0.weight' in saved
    assert os.path.exists(f'ckpts/pre/{experiment}_weights.pt')                 # best validation epoch (trainer.py:262-266)
    n_all = int(re.search(r'The number of trainable parameters: (\d+)', out).group(1))
    assert n_all == sum(v.numel() for k, v in saved.items())

    # ---- MoSA fine-tuning on 2 batches of 4 agents, with the initialisation check --------------------------------
    tune = (f'{COMMON} --fine_tune --seed 2 --n_epoch 3 --n_early_stop 30 --n_round 2 --pretrained_ckpt {pre} --train_net mosa_1 '
            '--position 0 1 --ckpt_path ckpts/tuned --n_train_batch 2 --lr 0.003 --steps 2 --init_check')
    train.main(get_parser(True).parse_args(tune.split()))
    train_out = capsys.readouterr().out
    assert 'Passed initialization check' in train_out and f'Loaded checkpoint {pre}' in train_out
    assert 'df_train: (88, 6); #=8' in train_out
    tuned_name = 'Seed_2__filter_agent_type_Biker__mosa_1__Pos_0_1__TrN_8__lr_0.003__original'   # (early stop 30 >= 3 epochs)
    assert f'Experiment {tuned_name} has started' in train_out
    tuned = f'ckpts/tuned/{tuned_name}.pt'
    sd = torch.load(tuned)
    assert sd and all('lora_' in k for k in sd)                                 # only the adapters are written
    n_lora = int(re.search(r'The number of trainable parameters: (\d+)', train_out).group(1))
    assert n_lora == sum(v.numel() for v in sd.values())
    assert any(float(v.detach().abs().max()) > 0 for k, v in sd.items() if 'lora_B' in k)   # training moved B away from zero
    averages = re.findall(AVERAGE, train_out)
    assert len(averages) == 3 and averages[0] == averages[1]                    # pretrained == freshly adapted, same seed
    assert all(a[0] == '2' for a in averages)
    assert len(re.findall(r'lr=0\.003\n', train_out)) == 2 and 'lr=0.0003' in train_out      # MultiStepLR(steps=[2], 0.1)

    # ---- test.py: pretrained + separately saved tuned parameters --------------------------------------------------
    test.main(get_parser(False).parse_args(
        f'{COMMON} --seed 2 --n_round 2 --pretrained_ckpt {pre} --tuned_ckpt {tuned}'.split()))
    eval_out = capsys.readouterr().out
    assert f"['{pre}', '{tuned}'] ['OODG', 'mosa_1[0_1](8)']" in eval_out
    evals = re.findall(AVERAGE, eval_out)
    assert len(evals) == 1 and evals[0] == averages[2]          # same weights, same seed -> the numbers train printed
    # ... and a whole checkpoint through --ckpts: the pretrained model's numbers (first average of the init check)
    test.main(get_parser(False).parse_args(f'{COMMON} --seed 2 --n_round 2 --ckpts {pre} --ckpts_name pre'.split()))
    assert re.findall(AVERAGE, capsys.readouterr().out) == [averages[0]]

    # ---- scraper --------------------------------------------------------------------------------------------------
    os.makedirs('logs')
    for name, text in (('tiny_train', train_out), ('tiny_eval', eval_out)):
        with open(f'logs/{name}.out', 'w') as f:
            f.write(text)
        extract_log.extract_file(f'logs/{name}.out', 'csv/log')
    row = pd.read_csv('csv/log/tiny_train.csv', dtype={'position': str}, float_precision='round_trip').iloc[0]
    assert (row.seed, row.train_net, row.n_train, str(row.position), row.n_param, row.n_epoch) == (2, 'mosa_1', 8, '0_1', n_lora, 99)
    assert (row.ade, row.fde) == (float(averages[0][1]), float(averages[0][2]))   # the scraper takes the first average
    assert row.experiment == tuned_name and row.pretrained_ckpt == f'{experiment}.pt' and not row.is_augment
    row = pd.read_csv('csv/log/tiny_eval.csv', dtype={'position': str}, float_precision='round_trip').iloc[0]
    assert (row.seed, row.train_net, row.n_train, str(row.position), float(row.lr)) == (2, 'mosa_1', 8, '0_1', 0.003)
    assert (row.ade, row.fde) == (float(evals[0][1]), float(evals[0][2]))
    assert row.tuned_ckpt == f'{tuned_name}.pt'

    # ---- evaluate_multickpts: per-agent table of both models under one seed, then the agents on which they differ most ----
    from motion_style_transfer_b200.evaluator import evaluate_multickpts as multi
    table = multi.main(multi.get_multickpts_parser().parse_args(
        f'{COMMON} --seed 2 --n_round 2 --pretrained_ckpt {pre} --tuned_ckpts {tuned}'.split()))
    out = capsys.readouterr().out
    csv_path = 'csv/comparison/2__filter_agent_type_Biker/OODG_mosa_1[0_1](8)__N6_R2.csv'
    assert f'Saved {csv_path}' in out and '====== Testing for mosa_1[0_1](8) ======' in out
    saved = pd.read_csv(csv_path, float_precision='round_trip')
    assert list(saved.columns) == ['metaId', 'sceneId', 'ade_OODG', 'fde_OODG', 'ade_mosa_1[0_1](8)', 'fde_mosa_1[0_1](8)']
    assert sorted(saved.metaId) == list(range(200, 206)) and len(table) == 6
    # paired samples: the column means are the numbers test.py printed for the two models under this seed
    np.testing.assert_allclose(saved['ade_OODG'].mean(), float(averages[0][1]), rtol=1e-5)
    np.testing.assert_allclose(saved['fde_mosa_1[0_1](8)'].mean(), float(evals[0][2]), rtol=1e-5)
    multi.main(multi.get_multickpts_parser().parse_args(
        f'{COMMON} --seed 2 --n_round 1 --ckpts {pre} --ckpts_name pre --result_path {csv_path} '
        '--result_name ade_OODG__ade_mosa_1[0_1](8)__abs_diff --result_limited 2'.split()))
    out = capsys.readouterr().out
    worst = saved.assign(d=(saved['ade_OODG'] - saved['ade_mosa_1[0_1](8)']).abs()).sort_values('d', ascending=False).metaId[:2]
    assert 'meta_ids_focus: #= 2' in out and 'df_test_limited: (22, 6); #=2' in out
    assert sorted(pd.read_csv('csv/comparison/2__filter_agent_type_Biker/pre__N2_R1.csv').metaId) == sorted(worst)
    with pytest.raises(NotImplementedError):
        multi.main(multi.get_multickpts_parser().parse_args(f'{COMMON} --ckpts {pre} --ckpts_name pre --viz'.split()))


def test_augmented_pretraining_epoch(workspace, capsys):
    """``--augment`` (scripts/*/pretrain.sh): 8 views per scene through the oriented preprocessing launch."""
    from motion_style_transfer_b200 import train
    from motion_style_transfer_b200.utils.parser import get_parser
    train.main(get_parser(True).parse_args(
        f'{COMMON} --seed 1 --n_epoch 1 --n_round 1 --ckpt_path ckpts/aug --augment --backend bf16x3'.split()))
    out = capsys.readouterr().out
    assert 'Augmented data and images' in out and 'Best epoch at 0' in out
    assert len(re.findall(AVERAGE, out)) == 1


def test_ind_ynetmod_sequential_split_smoothed_validation_cws(workspace_ind, capsys):
    """scripts/inD/scene1_car_to_truck/ynetmod/{pretrain,tune_mosa_A,generalize}.sh in small: Y-Net-Mod (``--network fusion
    --n_fusion 2``), MoSA on the motion branch, one pickle split sequentially with a shared validation / test set,
    validation smoothed over a window, evaluation with TTST + CWS."""
    from motion_style_transfer_b200 import train, test
    from motion_style_transfer_b200.utils import extract_log
    from motion_style_transfer_b200.utils.parser import get_parser
    common = (f'--config_filename tiny.yaml --dataset_path {DATASET_PATH} --network fusion --n_fusion 2 --batch_size 4 '
              '--load_data sequential --val_files car.pkl --val_split 2 --test_splits 6 --share_val_test')
    train.main(get_parser(True).parse_args(
        f'{common} --train_files car.pkl --seed 1 --n_epoch 1 --n_round 1 --ckpt_path ckpts'.split()))
    out = capsys.readouterr().out
    assert "Split ['car.pkl'] given val_split=2.0, test_split=[6]" in out and 'Share validation and test set' in out
    assert 'df_train: (110, 6); #=10' in out and 'df_val: (22, 6); #=2' in out and 'df_test: (66, 6); #=6' in out
    pre_name = 'Seed_1__filter_agent_type_Biker_car__train__fusion_2'
    assert f'Experiment {pre_name} has started' in out
    pre = 'ckpts/inD__ynetmod__car.pt'
    os.rename(f'ckpts/{pre_name}.pt', pre)

    tune = (f'{common} --train_files car.pkl --fine_tune --seed 1 --n_epoch 5 --n_early_stop 3000 --n_round 2 --pretrained_ckpt {pre} '
            '--train_net mosa_1 --position motion --ckpt_path ckpts/tuned --n_train_batch 2 --lr 0.001 --smooth_val --window_size 3 '
            '--init_check')
    train.main(get_parser(True).parse_args(tune.split()))
    train_out = capsys.readouterr().out
    tuned_name = 'Seed_1__filter_agent_type_Biker_car__mosa_1__Pos_motion__TrN_8__lr_0.001__smooth__fusion_2'
    assert f'Experiment {tuned_name} has started' in train_out and 'Passed initialization check' in train_out
    assert 'df_train: (88, 6); #=8' in train_out
    assert re.search(r'Best epoch at [23]\n', train_out)       # the first smoothed value exists at epoch 3 (centre: 2)
    sd = torch.load(f'ckpts/tuned/{tuned_name}.pt')
    assert sd and all('lora_' in k and k.startswith('encoder.motion_stages.') for k in sd)
    averages = re.findall(AVERAGE, train_out)
    assert len(averages) == 3 and averages[0] == averages[1]

    test.main(get_parser(False).parse_args(
        f'{common} --seed 1 --n_round 2 --pretrained_ckpt {pre} --tuned_ckpt ckpts/tuned/{tuned_name}.pt'.split()))
    eval_out = capsys.readouterr().out
    assert "['OODG', 'mosa_1[motion](8)']" in eval_out and 'TTST setting: True' in eval_out
    assert re.findall(AVERAGE, eval_out) == [averages[2]]

    os.makedirs('logs')
    with open('logs/ind_eval.out', 'w') as f:
        f.write(eval_out)
    extract_log.extract_file('logs/ind_eval.out', 'csv')
    row = pd.read_csv('csv/ind_eval.csv', float_precision='round_trip').iloc[0]
    assert (row.train_net, row.n_train, row.position, float(row.lr)) == ('mosa_1', 8, 'motion', 0.001)
    assert (row.ade, row.fde) == (float(averages[2][1]), float(averages[2][2]))


def test_raw_recordings_to_trained_model(tmp_path, monkeypatch, capsys, cuda_device):
    """The whole journey of scripts/sdd/preprocessing.sh + pretrain.sh in small: raw ``annotations.txt`` files -> ``sdd_dataset``
    (split at gaps, downsample, window, per-agent-type pickles, factor table) -> ``split_dataset`` -> ``train`` from scratch on
    the B200 path, with the scene images beside the annotations as in the SDD download."""
    import cv2
    from helpers import TinySeg
    from oracle import synth_raw
    from motion_style_transfer_b200 import train
    from motion_style_transfer_b200.utils import sdd_dataset, split_dataset
    from motion_style_transfer_b200.utils.parser import get_parser
    raw = tmp_path / 'data' / 'sdd' / 'raw'
    synth_raw.write_sdd(str(raw), seed=4, start=(250, 600), speed_x=(0.2, 0.9), speed_y=(-0.4, 0.4))    # x < 910, y < 600
    rng = np.random.RandomState(1)
    for dp, _, files in os.walk(raw / 'annotations'):
        if 'annotations.txt' in files:
            assert cv2.imwrite(os.path.join(dp, 'reference.jpg'), rng.randint(0, 256, (640, 960, 3)).astype(np.uint8))
    torch.manual_seed(0)
    torch.save(TinySeg(6), tmp_path / 'data' / 'sdd' / 'sdd_segmentation.pth')
    os.makedirs(tmp_path / 'config')
    with open(tmp_path / 'config' / 'tiny.yaml', 'w') as f:
        yaml.safe_dump(CONFIG, f, sort_keys=False)
    monkeypatch.chdir(tmp_path)
    sdd_dataset.main('--raw_data_dir data/sdd/raw --additional_data_dir data/sdd/raw --filter_data_dir data/sdd/filter/shortterm '
                     '--window_size 11 --stride 11 --obs_len 5 --varf agent_type --labels Pedestrian Biker'.split())
    split_dataset.main('--data_dir data/sdd/filter/shortterm/agent_type --data_filename Pedestrian.pkl --val_split 0.2 '
                       '--test_split 0.2 --seed 1'.split())
    out = capsys.readouterr().out
    assert '# data = 16' in out and '# train = 10' in out and '# val = 3' in out and '# test = 3' in out
    train.main(get_parser(True).parse_args(
        '--config_filename tiny.yaml --dataset_path filter/shortterm/agent_type/Pedestrian --network original --load_data '
        'predefined --batch_size 4 --seed 1 --n_epoch 2 --n_round 1 --ckpt_path ckpts'.split()))
    out = capsys.readouterr().out
    assert 'df_train: (110, 8); #=10' in out and 'df_test: (33, 8); #=3' in out       # (+ label, frame_diff columns)
    ade, fde = (float(v) for v in re.findall(AVERAGE, out)[0][1:])
    assert np.isfinite(ade) and np.isfinite(fde) and 0 < ade < 2000
    epochs = re.findall(r'Epoch (\d+): \tTrain \(Top-1\) ADE: ([\d\.]+) FDE: ([\d\.]+) \t\tVal \(Top-k\) ADE: ([\d\.]+)', out)
    assert [e[0] for e in epochs] == ['0', '1']
    assert os.path.exists('ckpts/Seed_1__filter_shortterm_agent_type_Pedestrian__train__original.pt')
