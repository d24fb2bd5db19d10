"""Drop-in for the reference's utils/train_epoch.py (one fine-tuning epoch), device-resident.

Same 20-argument signature and return value as train_epoch.py:8-12.  Heat maps are rasterised on the
device from device coordinates (no numpy round trip), the forward/backward of every layer and the
BCE run in libynet_b200.so kernels through ``autograd_engine``; ``optimizer.step()`` is whatever the
trainer passes (``FusedAdam`` = Adam kernel + optional NCCL all-reduce of the LoRA gradients).
"""
import torch

from .. import ops, parallel
from ..engine import ChannelCat
from .image_utils import swap_pavement_terrain


def train_epoch(model, train_loader, train_images, optimizer, criterion, loss_scale, device, dataset_name, homo_mat,
                gt_template, input_template, waypoints, epoch, obs_len, pred_len, batch_size, e_unfreeze,
                resize_factor, network=None, swap_semantic=False):
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('motion_style_transfer_b200.train_epoch runs on CUDA only (no CPU fallback)')
    train_loss = torch.zeros((), device=device)
    world_size = parallel.world()[1]
    trainable = [p for p in model.parameters() if p.requires_grad]
    if any(p.requires_grad for p in model.semantic_segmentation.parameters()):
        # train_epoch.py:34-47 re-runs the segmentation backbone WITH gradients once epoch >= e_unfreeze; here it runs
        # once per scene under no_grad (the backbone is outside the B200 path), which would train nothing, silently.
        raise NotImplementedError('fine-tuning the segmentation backbone (train_net=segmentation_*, e_unfreeze) is '
                                  'outside the B200 hot path')
    train_ADE, train_FDE = [], []
    model.train()
    input_template = input_template.to(device=device, dtype=torch.float32)
    gt_template = gt_template.to(device=device, dtype=torch.float32)

    # train_epoch.py:34-47 switches the whole model to eval() for the segmentation pass of every scene and back to train();
    # the only module that runs in between is the (frozen) backbone, so only that one is switched -- once, not a walk over
    # ~110 modules twice per scene (2 ms of host time per step) -- and handed back in train mode at the end
    model.semantic_segmentation.eval()
    for batch, (trajectory, meta, scene) in enumerate(train_loader):
        with torch.no_grad():
            scene_image = train_images[scene].to(device).unsqueeze(0)
            scene_image = model.segmentation(scene_image).float().contiguous()
        trajectory = trajectory.to(device=device, dtype=torch.float32)
        for i in range(0, len(trajectory), batch_size):
            semantic_img = model.adapt_semantic(scene_image)
            if swap_semantic:                      # train_epoch.py:57-58
                semantic_img = swap_pavement_terrain(semantic_img)
            _, _, H, W = scene_image.shape
            batch_traj = trajectory[i:i + batch_size]
            # one process per GPU: this rank takes a contiguous share of the batch's agents; its mean loss is
            # weighted so that the rank-average of the gradients is the gradient of the full-batch mean
            lo, hi = parallel.shard_bounds(batch_traj.shape[0])
            traj = batch_traj[lo:hi]
            B = traj.shape[0]
            shard_weight = B * world_size / batch_traj.shape[0]
            if B == 0:      # more ranks than agents: contribute a zero gradient, stay in the collective
                for p in trainable:
                    p.grad = torch.zeros_like(p)
                optimizer.step()
                continue
            observed_map = ops.rasterize_patches(input_template, traj[:, :obs_len].reshape(-1, 2), H, W)
            observed_map = observed_map.view(B, obs_len, H, W)
            gt_future = traj[:, obs_len:].contiguous()
            gt_future_map = ops.rasterize_patches(gt_template, gt_future.reshape(-1, 2), H, W).view(B, pred_len, H, W)
            gt_waypoints = gt_future[:, waypoints]
            gt_waypoint_map = ops.rasterize_patches(input_template, gt_waypoints.reshape(-1, 2), H, W)
            gt_waypoint_map = gt_waypoint_map.view(B, gt_waypoints.shape[1], H, W)

            if network == 'embed':                 # train_epoch.py:81-83
                semantic_img = model.scene_embedding(semantic_img)
                observed_map = model.motion_embedding(observed_map)
            features = model.pred_features(semantic_img, observed_map)            # scene broadcast over B
            pred_goal_map = model.pred_goal(features)
            goal_loss = criterion(pred_goal_map, gt_future_map) * loss_scale

            pyr = ops.avgpool_pyramid(gt_waypoint_map, len(features))
            traj_input = [ChannelCat(tuple(f) + (g,)) if isinstance(f, tuple) else ChannelCat((f, g))
                          for f, g in zip(features, pyr)]
            pred_traj_map = model.pred_traj(traj_input)
            traj_loss = criterion(pred_traj_map, gt_future_map) * loss_scale

            loss = goal_loss + traj_loss
            if world_size > 1:
                loss = loss * shard_weight
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()

            with torch.no_grad():
                train_loss += loss.detach()
                pred_traj = model.softargmax(pred_traj_map.detach())
                pred_goal = ops.softargmax2d(pred_goal_map.detach(), channel=-1)
                train_ADE.append(((((gt_future - pred_traj) / resize_factor) ** 2).sum(dim=2) ** 0.5).mean(dim=1))
                train_FDE.append(((((gt_future[:, -1:] - pred_goal[:, -1:]) / resize_factor) ** 2).sum(dim=2) ** 0.5)
                                 .mean(dim=1))

    model.semantic_segmentation.train()
    train_ADE = torch.cat(train_ADE) if train_ADE else torch.zeros(0, device=device)
    train_FDE = torch.cat(train_FDE) if train_FDE else torch.zeros(0, device=device)
    if world_size == 1:
        return train_ADE.mean().item(), train_FDE.mean().item(), train_loss.item()
    ade, fde, _ = parallel.reduce_metric_sums(train_ADE.sum().item(), train_FDE.sum().item(), train_ADE.numel(), device)
    loss_sum, _, _ = parallel.reduce_metric_sums(train_loss.item(), 0.0, 1, device)   # mean over ranks (count 1 each)
    return ade, fde, loss_sum
