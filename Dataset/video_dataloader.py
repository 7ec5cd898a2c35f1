"""`.npy` face-clip loader with the reference's layout (video_dataloader.py:6-68): files `<label>_*.npy` of uint8
(T,H,W,3) -> float32 (T,3,H,W)/255; label from the file-name prefix; zero-padding collate."""
import os

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .synthetic import SyntheticClips, collate_clips as collate_fn, dataset_missing, label_from_name, synthetic_requested  # noqa: F401


class FaceDataset(Dataset):
    """raw_uint8=True returns the frames exactly as stored -- uint8 (T,H,W,3) -- for the models' uint8 ingest path
    (4x fewer host->device bytes, no permute; the 1/255 scaling happens inside the stem kernel)."""

    def __init__(self, folder_path, raw_uint8=False):
        self.raw_uint8 = raw_uint8
        self.files = sorted(os.path.join(folder_path, f) for f in os.listdir(folder_path) if f.endswith(".npy"))

    def __len__(self):
        return len(self.files)

    def __getitem__(self, idx):
        arr = np.load(self.files[idx])                                        # (T,H,W,3) uint8
        label = torch.tensor([label_from_name(self.files[idx])], dtype=torch.float32)       # (1,) like video_dataloader.py:37
        if self.raw_uint8:
            return torch.from_numpy(arr), label
        frames = torch.from_numpy(arr).permute(0, 3, 1, 2).float().div_(255.0)
        return frames, label


def get_face_dataloader(folder_path, batch_size=4, shuffle=True, num_workers=0, raw_uint8=False):
    if folder_path and os.path.isdir(folder_path):
        ds = FaceDataset(folder_path, raw_uint8)
    elif synthetic_requested():
        ds = SyntheticClips(raw_uint8=raw_uint8)
    else:
        raise dataset_missing("get_face_dataloader", folder_path)
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, collate_fn=collate_fn, pin_memory=True)
