"""Scene-per-item dataset (host side) with the interface of the reference's utils/dataloader.py:8-50."""
import numpy as np
import torch
from torch.utils.data import Dataset


class SceneDataset(Dataset):
    """One element = all trajectories of one scene, already multiplied by the resize factor."""

    def __init__(self, data, resize, total_len):
        self.trajectories, self.meta, self.scene_list = self.split_trajectories_by_scene(data, total_len)
        self.trajectories = [t * resize for t in self.trajectories]

    def __len__(self):
        return len(self.trajectories)

    def __getitem__(self, idx):
        return self.trajectories[idx], self.meta[idx], self.scene_list[idx]

    def split_trajectories_by_scene(self, data, total_len):
        trajectories, meta, scene_list = [], [], []
        for _, meta_df in data.groupby('sceneId', as_index=False):
            trajectories.append(meta_df[['x', 'y']].to_numpy().astype('float32').reshape(-1, total_len, 2))
            meta.append(meta_df)
            scene_list.append(meta_df.iloc()[0:1].sceneId.item())
        return trajectories, meta, scene_list


def scene_collate(batch):
    """batch_size is 1 scene: returns (trajectories (N, T, 2) float32 tensor, [meta df], scene id)."""
    trajectories = [b[0] for b in batch]
    meta = [b[1] for b in batch]
    scene = [b[2] for b in batch]
    return torch.from_numpy(np.ascontiguousarray(trajectories[0], dtype=np.float32)), meta, scene[0]
