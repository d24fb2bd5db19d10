"""Drop-in ``YNetTrainer`` for the reference's models/trainer.py:45-614.

Constructor arguments, attributes, the train/test/load/save methods, the freeze-policy DSL
(``train_net`` / ``position`` / ``ynet_bias``), the checkpoint layout and the stdout formats scraped
by the reference's log tools are preserved; the compute runs through libynet_b200.so.
``prepare_data`` (trainer.py:518-584) reads the scene images with cv2 and preprocesses them in one fused CUDA launch
each; ``train_prepared`` / ``test_prepared`` / ``forward_test_prepared`` accept ready-made (images dict, DataLoader)
pairs.
"""
import os
import pathlib
import re
from collections import OrderedDict, deque
from copy import deepcopy

import torch
from torch.utils.data import DataLoader
from tqdm import tqdm

from .. import ops, parallel
from ..engine import ChannelCat
from ..autograd_engine import BCEWithLogitsLoss
from ..utils.dataloader import SceneDataset, scene_collate
from ..utils.evaluate import evaluate
from ..utils.image_utils import create_dist_mat, create_gaussian_heatmap_template
from ..utils.train_epoch import train_epoch
from .ynet import YNet


def _mark_bias(module):
    for name, p in module.named_parameters():
        if 'bias' in name:
            p.requires_grad = True


def mark_encoder_bias_trainable(model):
    _mark_bias(model.encoder)
    return model


def mark_goal_bias_trainable(model):
    _mark_bias(model.goal_decoder)
    return model


def mark_traj_bias_trainable(model):
    _mark_bias(model.traj_decoder)
    return model


def mark_ynet_bias_trainable(model):
    return mark_traj_bias_trainable(mark_goal_bias_trainable(mark_encoder_bias_trainable(model)))


def apply_freeze_policy(model, train_net, position, network, ynet_bias=False):
    """trainer.py:112-195: which parameters train for a given ``train_net`` / ``position``."""
    for p in model.semantic_segmentation.parameters():
        p.requires_grad = False
    if train_net in ('all', 'train'):
        return model
    for p in model.parameters():
        p.requires_grad = False
    position = [str(i) for i in position]
    enc = model.encoder
    fusion_sets = {
        'scene': ('scene_stages',), 'motion': ('motion_stages',), 'fusion': ('fusion_stages',),
        'scene_fusion': ('scene_stages', 'fusion_stages'), 'motion_fusion': ('motion_stages', 'fusion_stages'),
        'scene_motion': ('scene_stages', 'motion_stages'),
        'scene_motion_fusion': ('scene_stages', 'motion_stages', 'fusion_stages'),
    }
    if train_net == 'encoder' and len(position) == 0:
        for p in enc.parameters():
            p.requires_grad = True
    elif train_net == 'encoder':
        for name, p in enc.named_parameters():
            if name.split('.')[1] in position:
                p.requires_grad = True
    elif 'serial' in train_net or 'parallel' in train_net:
        key = 'serial' if 'serial' in train_net else 'parallel'
        for name, p in enc.named_parameters():
            if key in name:
                p.requires_grad = True
    elif 'mosa' in train_net:
        for name, p in enc.named_parameters():
            if 'lora' in name:
                p.requires_grad = True
    elif 'semantic' in train_net:
        for name, p in model.named_parameters():
            if 'semantic_adapter' in name:
                p.requires_grad = True
    elif network == 'fusion' and train_net in fusion_sets:
        for attr in fusion_sets[train_net]:
            for p in getattr(enc, attr).parameters():
                p.requires_grad = True
    elif train_net == 'biasEncoder':
        mark_encoder_bias_trainable(model)
    elif train_net == 'biasGoal':
        mark_goal_bias_trainable(model)
    elif train_net == 'biasTraj':
        mark_traj_bias_trainable(model)
    elif train_net == 'bias':
        mark_ynet_bias_trainable(model)
    elif 'segmentation' in train_net:
        layer = train_net.split('_')[1]
        for name, p in model.semantic_segmentation.named_parameters():
            if (layer in ['head', 'bias', 'bn'] and layer in name) or \
                    re.search(rf'decoder.blocks.\d.{layer}', name) is not None:
                p.requires_grad = True
    else:
        raise NotImplementedError
    if ynet_bias:
        mark_ynet_bias_trainable(model)
    return model


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam defaults (trainer.py:197) with the update done by ``ynet_adam_step``.

    Under torch.distributed (one process per GPU, agents sharded across ranks) the gradients of all
    trainable tensors -- 8 190 floats for Y-Net mosa_1 -- are flattened into one buffer, summed with a
    single NCCL all-reduce over NVLink and averaged inside the Adam kernel (grad_scale = 1/world).
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        dp = parallel.world()[1] > 1
        for group in self.param_groups:
            # Under data parallelism the flat buffer has ONE layout on every rank (all trainable tensors, zeros where a
            # rank has no gradient) followed by one "has a gradient" flag per tensor: after the SUM a tensor is stepped
            # iff some rank produced a gradient for it -- torch.optim.Adam's skip-if-None rule, decided globally.
            flat, ps = parallel.flatten_grads(group['params'], fixed_layout=dp)
            if not ps:
                continue
            n_grad = sum(p.numel() for p in ps)
            if dp:
                flags = torch.tensor([0.0 if p.grad is None else 1.0 for p in ps], dtype=flat.dtype, device=flat.device)
                flat = torch.cat([flat, flags])
            grad_scale = parallel.allreduce_flat(flat)          # ONE collective per step; mean taken in the kernel
            live = (flat[n_grad:] > 0).tolist() if dp else [True] * len(ps)
            off = 0
            for p, has_grad in zip(ps, live):
                n = p.numel()
                if not has_grad:
                    off += n
                    continue
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['step'] += 1
                ops.adam_step(p.data, flat[off:off + n], st['exp_avg'], st['exp_avg_sq'], st['step'], group['lr'],
                              group['betas'][0], group['betas'][1], group['eps'], grad_scale)
                off += n
                # the kernel wrote through a raw pointer: bump the version so cached folded weights refresh
                if hasattr(torch.autograd.graph, 'increment_version'):
                    torch.autograd.graph.increment_version(p)
                else:  # pragma: no cover
                    p.add_(0)
        return None


class YNetTrainer:
    def __init__(self, params, device=None):
        self.params = params
        self.device = device if device else torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        print(f'Working on {self.device}')
        self.division_factor = 2 ** len(params['encoder_channels'])
        self.template_size = int(4200 * params['resize_factor'])
        self.model = YNet(
            obs_len=params['obs_len'], pred_len=params['pred_len'],
            segmentation_model_fp=params['segmentation_model_fp'],
            use_features_only=params['use_features_only'],
            n_semantic_classes=params['n_semantic_classes'],
            encoder_channels=params['encoder_channels'], decoder_channels=params['decoder_channels'],
            n_waypoints=len(params['waypoints']), train_net=params['train_net'], position=params['position'],
            network=params['network'], n_fusion=params['n_fusion'])
        self.homo_mat = None

    # ------------------------------------------------------------------------------------------ train
    def train(self, df_train, df_val, train_image_path, val_image_path, experiment_name):
        p = self.params
        train_images, train_loader, self.homo_mat = self.prepare_data(
            df_train, train_image_path, p['dataset_name'], 'train', p['obs_len'], p['pred_len'], p['resize_factor'],
            p.get('use_raw_data', False), p.get('augment', False))
        val_images, val_loader, _ = self.prepare_data(
            df_val, val_image_path, p['dataset_name'], 'val', p['obs_len'], p['pred_len'], p['resize_factor'],
            p.get('use_raw_data', False), False)
        return self.train_prepared(train_images, train_loader, val_images, val_loader, experiment_name)

    def train_prepared(self, train_images, train_loader, val_images, val_loader, experiment_name):
        return self._train(train_images, train_loader, val_images, val_loader, experiment_name, **self.params)

    def _train(self, train_images, train_loader, val_images, val_loader, experiment_name, ckpt_path, dataset_name,
               resize_factor, obs_len, pred_len, batch_size, lr, n_epoch, waypoints, n_goal, n_traj, kernlen, nsig,
               e_unfreeze, loss_scale, temperature, use_raw_data=False, save_every_n=10, train_net='all', position=[],
               fine_tune=False, augment=False, ynet_bias=False, use_CWS=False, resl_thresh=0.002, CWS_params=None,
               n_early_stop=5, steps=[20], lr_decay_ratio=0.1, network=None, swap_semantic=False, window_size=9,
               smooth_val=False, **kwargs):
        model = self.model.to(self.device)
        apply_freeze_policy(model, train_net, position, network, ynet_bias)
        if 'segmentation' in str(train_net):
            # trainer.py:176-184 marks segmentation-backbone tensors trainable, and train_epoch.py:34-47 then runs the
            # backbone with gradients once epoch >= e_unfreeze.  The backbone is outside the B200 path (it runs under
            # no_grad here), so such a run would silently train nothing.
            raise NotImplementedError("train_net='segmentation_*' fine-tunes the segmentation backbone, which is outside "
                                      'the B200 hot path')
        # one process per GPU: every rank starts from rank 0's trainable tensors (LoRA A is random per process)
        parallel.broadcast_params([p for p in model.parameters() if p.requires_grad])
        optimizer = FusedAdam(model.parameters(), lr=lr)
        if fine_tune:
            print('LR Schedular because finetuning')
            lr_scheduler = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=steps, gamma=lr_decay_ratio)
        print('The number of trainable parameters: {:d}'.format(
            sum(param.numel() for param in model.parameters() if param.requires_grad)))
        criterion = BCEWithLogitsLoss()
        input_template = torch.Tensor(create_dist_mat(size=self.template_size)).to(self.device)
        gt_template = torch.Tensor(create_gaussian_heatmap_template(
            size=self.template_size, kernlen=kernlen, nsig=nsig, normalize=False)).to(self.device)

        best_val_ADE, best_epoch = 99999999999999, 0
        self.val_ADE, self.val_FDE = [], []
        state_dicts = deque()
        half_window_size = (window_size // 2) + 1
        curr_model_dict = best_state_dict = None
        print('Start training')
        for e in tqdm(range(n_epoch), desc='Epoch'):
            train_ADE, train_FDE, train_loss = train_epoch(
                model, train_loader, train_images, optimizer, criterion, loss_scale, self.device, dataset_name,
                self.homo_mat, gt_template, input_template, waypoints, e, obs_len, pred_len, batch_size, e_unfreeze,
                resize_factor, network, swap_semantic)
            # like the reference, validation skips TTST (trainer.py:229-235)
            val_ADE, val_FDE, _, _ = evaluate(
                model, val_loader, val_images, self.device, dataset_name, self.homo_mat, input_template, waypoints,
                'val', n_goal, n_traj, obs_len, batch_size, resize_factor, temperature, False, use_CWS, resl_thresh,
                CWS_params, network=network, swap_semantic=swap_semantic)
            if fine_tune:
                print(f'Epoch {e}: \tTrain (Top-1) ADE: {train_ADE:.2f} FDE: {train_FDE:.2f} \t\tVal (Top-k) ADE: '
                      f'{val_ADE:.2f} FDE: {val_FDE:.2f}   lr={lr_scheduler.get_last_lr()[0]}')
            else:
                print(f'Epoch {e}: \tTrain (Top-1) ADE: {train_ADE:.2f} FDE: {train_FDE:.2f} \t\tVal (Top-k) ADE: '
                      f'{val_ADE:.2f} FDE: {val_FDE:.2f}')
            self.val_ADE.append(val_ADE)
            self.val_FDE.append(val_FDE)
            if fine_tune:
                lr_scheduler.step()
            if smooth_val:
                print('Length: ', len(state_dicts))
                if len(state_dicts) == half_window_size:
                    curr_model_dict = state_dicts.popleft()
                state_dicts.append(deepcopy(model.state_dict()))
                val_ADE = best_val_ADE + 1 if e < window_size else sum(self.val_ADE[-window_size:]) / window_size
            else:
                curr_model_dict = deepcopy(model.state_dict())
            if val_ADE < best_val_ADE:
                best_val_ADE = val_ADE
                best_epoch = e - half_window_size + 1 if smooth_val else e
                best_state_dict = curr_model_dict
                if not fine_tune:
                    print(f'Best Epoch {e}: \nVal ADE: {val_ADE} \nVal FDE: {val_FDE}')
                    if parallel.world()[0] == 0:        # replicas are identical: one writer
                        pathlib.Path(ckpt_path).mkdir(parents=True, exist_ok=True)
                        torch.save(model.state_dict(), f'{ckpt_path}/{experiment_name}_weights.pt')
                    parallel.barrier()
            if (e + 1) % save_every_n == 0:
                self.save_params(f'{ckpt_path}/{experiment_name}__epoch_{e}.pt', train_net)
            if fine_tune and (best_val_ADE < min(self.val_ADE[-n_early_stop:])):
                print(f'Early stop at epoch {e}')
                break
        print(f'Best epoch at {best_epoch}')
        if best_epoch != 0 and best_state_dict is not None:
            model.load_state_dict(best_state_dict, strict=True)
        self.save_params(f'{ckpt_path}/{experiment_name}.pt', train_net)
        return self.val_ADE, self.val_FDE

    # ------------------------------------------------------------------------------------------ test
    def test(self, df_test, image_path, return_preds=False, return_samples=False):
        p = self.params
        test_images, test_loader, self.homo_mat = self.prepare_data(
            df_test, image_path, p['dataset_name'], 'test', p['obs_len'], p['pred_len'], p['resize_factor'],
            p.get('use_raw_data', False))
        return self.test_prepared(test_images, test_loader, return_preds, return_samples)

    def test_prepared(self, test_images, test_loader, return_preds=False, return_samples=False):
        return self._test(test_images, test_loader, return_preds=return_preds, return_samples=return_samples,
                          **self.params)

    def _test(self, test_images, test_loader, dataset_name, resize_factor, batch_size, n_round, obs_len, pred_len,
              waypoints, n_goal, n_traj, temperature, rel_threshold, use_TTST, use_CWS, CWS_params,
              use_raw_data=False, return_preds=False, return_samples=False, network=None, swap_semantic=False,
              **kwargs):
        model = self.model.to(self.device)
        input_template = torch.Tensor(create_dist_mat(size=self.template_size)).to(self.device)
        self.eval_ADE, self.eval_FDE = [], []
        list_metrics, list_trajs = [], []
        print('TTST setting:', use_TTST)
        print('Start testing')
        for e in tqdm(range(n_round), desc='Round'):
            test_ADE, test_FDE, df_metrics, trajs_dict = evaluate(
                model, test_loader, test_images, self.device, dataset_name, self.homo_mat, input_template, waypoints,
                'test', n_goal, n_traj, obs_len, batch_size, resize_factor, temperature, use_TTST, use_CWS,
                rel_threshold, CWS_params, return_preds=return_preds, return_samples=return_samples, network=network,
                swap_semantic=swap_semantic)
            list_metrics.append(df_metrics)
            list_trajs.append(trajs_dict)
            print(f'Round {e}: \nTest ADE: {test_ADE} \nTest FDE: {test_FDE}')
            self.eval_ADE.append(test_ADE)
            self.eval_FDE.append(test_FDE)
        avg_ade = sum(self.eval_ADE) / len(self.eval_ADE)
        avg_fde = sum(self.eval_FDE) / len(self.eval_FDE)
        print(f'\nAverage performance (by {n_round}): \nTest ADE: {avg_ade} \nTest FDE: {avg_fde}')
        return avg_ade, avg_fde, list_metrics, list_trajs

    # ------------------------------------------------------------------------------------------ saliency
    def forward_test(self, df_test, image_path, set_input, noisy_std_frac):
        """trainer.py:354-355: one differentiable forward of one scene, for input-gradient (saliency) studies."""
        p = self.params
        test_images, test_loader, self.homo_mat = self.prepare_data(
            df_test, image_path, p['dataset_name'], 'test', p['obs_len'], p['pred_len'], p['resize_factor'],
            p.get('use_raw_data', False))
        return self.forward_test_prepared(test_images, test_loader, set_input, noisy_std_frac)

    def forward_test_prepared(self, test_images, test_loader, set_input, noisy_std_frac):
        return self._forward_test(test_images, test_loader, set_input, noisy_std_frac, **self.params)

    def _forward_test(self, test_images, test_loader, set_input, noisy_std_frac, decision, obs_len, pred_len, waypoints,
                      kernlen, nsig, loss_scale, **kwargs):
        """trainer.py:357-443.  ``decision`` = 'loss' -> (goal_loss, traj_loss, scene image[, noisy scene image]);
        'map' -> (goal logits, trajectory logits, scene image[, noisy scene image, cat(semantic, observed maps)]).

        Which tensor carries ``requires_grad`` follows the reference: the noisy copy of the scene image when noise goes
        on the semantic input, else the scene image itself if 'scene' is in ``set_input``; noise on the semantic map /
        the observed maps is added inside ``_forward_batch``.  (In the reference only ``noisy_std_frac=None`` with
        ``decision='map'`` runs to the end -- 'loss' passes ``False`` as ``set_input`` (trainer.py:381-383, TypeError) and
        the noisy branches return names that were never bound (trainer.py:436-440); here every combination returns what
        those lines evidently meant.)"""
        if decision not in ('loss', 'map'):
            raise ValueError(f'No support for decision={decision}')
        if len(test_loader) == 0:
            raise ValueError('No data is provided')
        if len(test_loader) != 1:
            raise ValueError(f'Received more than 1 scene ({len(test_loader)})')
        input_template = torch.Tensor(create_dist_mat(size=self.template_size)).to(self.device)
        gt_template = torch.Tensor(create_gaussian_heatmap_template(
            size=self.template_size, kernlen=kernlen, nsig=nsig, normalize=False)).to(self.device)
        criterion = BCEWithLogitsLoss()
        traj, _, scene_id = next(iter(test_loader))
        scene_raw_img = test_images[scene_id].to(self.device).unsqueeze(0)
        noisy_scene_img = None
        fed = scene_raw_img
        if noisy_std_frac is not None and 'semantic' in set_input:
            std = float(noisy_std_frac * (scene_raw_img.max() - scene_raw_img.min()))
            noisy_scene_img = scene_raw_img + scene_raw_img.new(scene_raw_img.size()).normal_(0, std)
            noisy_scene_img.requires_grad = True
            fed = noisy_scene_img
        elif noisy_std_frac is None:
            scene_raw_img.requires_grad = 'scene' in set_input
        out = self._forward_batch(fed, traj, input_template, gt_template, criterion, obs_len, pred_len, waypoints,
                                  loss_scale, self.device, set_input, noisy_std_frac, decision == 'map')
        head = tuple(out[:2])
        if noisy_std_frac is None:
            return head + (scene_raw_img,)
        if decision == 'loss':
            return head + (scene_raw_img, noisy_scene_img)
        return head + (scene_raw_img, noisy_scene_img, out[2] if len(out) > 2 else None)

    def _forward_batch(self, scene_raw_img, traj, input_template, gt_template, criterion, obs_len, pred_len, waypoints,
                       loss_scale, device, set_input=None, noisy_std_frac=None, return_pred_map=False):
        """trainer.py:445-516: goal decoder on (semantic map, observed maps), trajectory decoder on the features plus
        the AvgPool pyramid of the PREDICTED waypoint logits, BCE of both against the Gaussian ground-truth maps.

        Runs on the differentiable executor (``autograd_engine``: conv forward / dgrad kernels of libynet_b200.so) as
        soon as an input requires a gradient, so ``.backward()`` on the returned losses or maps reaches
        ``scene_raw_img`` through the segmentation backbone.  Heat maps are rasterised on the device (constants)."""
        model = self.model.to(self.device)
        set_input = () if set_input is None else set_input
        _, _, H, W = scene_raw_img.shape
        traj = traj.to(device=self.device, dtype=torch.float32)
        B = traj.shape[0]
        observed_map = ops.rasterize_patches(input_template, traj[:, :obs_len].reshape(-1, 2), H, W).view(B, obs_len, H, W)
        gt_future_map = ops.rasterize_patches(gt_template, traj[:, obs_len:].reshape(-1, 2), H, W).view(B, pred_len, H, W)
        semantic_image = model.adapt_semantic(model.segmentation(scene_raw_img))     # (1, C, H, W): broadcast over agents

        scene_in, motion_in = semantic_image, observed_map
        noisy = False
        if noisy_std_frac is not None and 'semantic' in set_input:
            semantic_image = semantic_image.expand(B, -1, -1, -1)
            std = float(noisy_std_frac * (semantic_image.detach().max() - semantic_image.detach().min()))
            scene_in = semantic_image + semantic_image.new(semantic_image.size()).normal_(0, std)
            scene_in.requires_grad_(True)
            noisy = True
        if 'traj' in set_input:
            if noisy_std_frac is not None:
                std = float(noisy_std_frac * (observed_map.max() - observed_map.min()))
                motion_in = observed_map + observed_map.new(observed_map.size()).normal_(0, std)
                motion_in.requires_grad_(True)
                if scene_in is not semantic_image:
                    # trainer.py:487-488: with both inputs noisy the reference feeds the noisy SEMANTIC map as the
                    # motion input too; kept (it only type-checks when the two have the same channel count)
                    motion_in = scene_in
                noisy = True
            else:
                observed_map.requires_grad_(True)

        features = model.pred_features(scene_in.float().contiguous(), motion_in.contiguous())
        pred_goal_map = model.pred_goal(features)
        goal_loss = criterion(pred_goal_map, gt_future_map) * loss_scale
        pred_waypoint_map = pred_goal_map[:, waypoints]
        # nn.AvgPool2d(2^i) of the predicted logits, differentiable (torch's pooling: this tool is not on the hot path)
        pyr = [pred_waypoint_map] + [torch.nn.functional.avg_pool2d(pred_waypoint_map, 2 ** i, 2 ** i)
                                     for i in range(1, len(features))]
        traj_input = [ChannelCat(tuple(f) + (g.contiguous(),)) if isinstance(f, tuple) else ChannelCat((f, g.contiguous()))
                      for f, g in zip(features, pyr)]
        pred_traj_map = model.pred_traj(traj_input)
        traj_loss = criterion(pred_traj_map, gt_future_map) * loss_scale
        if return_pred_map:
            if noisy:
                sem = semantic_image if semantic_image.shape[0] == B else semantic_image.expand(B, -1, -1, -1)
                return pred_goal_map, pred_traj_map, torch.cat([sem, observed_map], dim=1)
            return pred_goal_map, pred_traj_map
        return goal_loss, traj_loss

    # ------------------------------------------------------------------------------------------ data
    def prepare_data(self, df, image_path, dataset_name, mode, obs_len, pred_len, resize_factor, use_raw_data,
                     augment=False):
        """trainer.py:518-584.  Scene images are read with cv2 (host I/O, data_utils.py:248-263) and go through ONE fused
        CUDA launch each -- resize (INTER_AREA) -> pad to a multiple of ``division_factor`` -> segmentation-backbone
        normalisation, bit-exact against the reference's cv2 / numpy chain (utils/image_utils.py); the
        returned dict holds device tensors.  ``augment`` (data_utils.py:115-233): see below."""
        import cv2
        from ..utils.image_utils import preprocess_scene_images
        dataset_name = dataset_name.lower()
        names = {'sdd': 'reference.jpg', 'ind-dataset-v1.0': 'reference.png', 'eth': 'oracle.png'}
        if dataset_name not in names:
            raise ValueError(f'{dataset_name} dataset is not supported')
        if dataset_name == 'eth':
            raise NotImplementedError('ETH/UCY homography path is outside the B200 hot path')
        images_dict = {}
        for scene in df.sceneId.unique():
            if use_raw_data:
                scene_name, scene_idx = scene.split('_')
                im_path = os.path.join(image_path, scene_name, f'video{scene_idx}', names[dataset_name])
            else:
                im_path = os.path.join(image_path, scene, names[dataset_name])
            im = cv2.imread(im_path)                 # channels: blue, green, red (data_utils.py:262)
            if im is None:
                raise FileNotFoundError(im_path)
            images_dict[scene] = im
        if not augment:
            print('No data and images augmentation')
            preprocess_scene_images(images_dict, resize_factor, self.division_factor, seg_mask=False, device=self.device)
        else:
            # data_utils.py:163-233: x8 (three rotations, then the mirror image of all four).  The trajectories are
            # transformed on the host in the reference's float64 arithmetic; the eight image views of a scene are read
            # by the preprocessing kernel from ONE uploaded image (no rotated copies)
            from ..utils.image_utils import augment_data, preprocess_scene_image
            df, views = augment_data(df, images_dict)
            stored = {scene: torch.as_tensor(im).to(self.device) for scene, im in images_dict.items()}
            images_dict = {view: preprocess_scene_image(stored[base], resize_factor, self.division_factor, device=self.device,
                                                        orient=orient) for view, (base, orient) in views.items()}
            print('Augmented data and images')
        dataset = SceneDataset(df, resize=resize_factor, total_len=obs_len + pred_len)
        dataloader = DataLoader(dataset, batch_size=1, collate_fn=scene_collate, shuffle=(mode == 'train'),
                                generator=parallel.shared_generator() if mode == 'train' else None)
        return images_dict, dataloader, None

    # ------------------------------------------------------------------------------------------ checkpoints
    def load_params(self, path):
        self.model.load_state_dict(torch.load(path, map_location=self.device), strict=False)
        print(f'Loaded ynet model to {"GPU" if torch.device(self.device).type == "cuda" else "CPU"}')

    def save_params(self, path, train_net):
        if train_net == 'all' or train_net == 'train':
            state_dict = {k: v for k, v in self.model.state_dict().items() if 'segmentation' not in k}
        else:
            state_dict = OrderedDict()
            for name, param in self.model.named_parameters():
                if param.requires_grad:
                    state_dict[name] = param
        # one process per GPU: the replicas hold identical tensors, rank 0 alone writes (concurrent torch.save calls on
        # one path corrupt the file); everyone waits so that a following load sees the complete file
        if parallel.world()[0] == 0:
            parent = os.path.dirname(str(path))
            if parent:
                pathlib.Path(parent).mkdir(parents=True, exist_ok=True)
            torch.save(state_dict, path)
        parallel.barrier()

    def load_separated_params(self, pretrained_path, tuned_path):
        self.model.load_state_dict(torch.load(pretrained_path, map_location=self.device), strict=False)
        self.model.load_state_dict(torch.load(tuned_path, map_location=self.device), strict=False)
        print(f'Loaded ynet model to {"GPU" if torch.device(self.device).type == "cuda" else "CPU"}')
