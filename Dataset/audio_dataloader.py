"""MFCC loader with the reference's layout (audio_dataloader.py:6-47): `.npy` (120,13) -> (T,3,13) by channel repeat,
label from the file-name prefix, zero-padding collate -> (B,Tmax,3,13), labels (B,1)."""
import os

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .synthetic import SyntheticAudio, collate_audio as collate_fn, dataset_missing, label_from_name, synthetic_requested  # noqa: F401


class AudioDataset(Dataset):
    def __init__(self, folder_path):
        self.files = sorted(os.path.join(folder_path, f) for f in os.listdir(folder_path) if f.endswith(".npy"))

    def __len__(self):
        return len(self.files)

    def __getitem__(self, idx):
        mf = torch.from_numpy(np.load(self.files[idx])).float()              # (T,13)
        return mf.unsqueeze(1).repeat(1, 3, 1), torch.tensor([label_from_name(self.files[idx])], dtype=torch.float32)


def get_audio_dataloader(folder_path, batch_size=8, shuffle=True, waveforms=False):
    """waveforms=True (no reference counterpart): yield raw 16 kHz waveforms (B, samples) for the GPU
    MFCC front-end (multimodal_deepfake_detection_b200.audio_frontend.MFCC) instead of pre-computed MFCC files."""
    if waveforms:
        # the reference defines no on-disk waveform format (its pre-processor goes mp4 -> wav -> MFCC .npy in one pass,
        # wavfake_audio_dataset.py:30-44), so this route only exists on synthetic waveforms
        if not synthetic_requested():
            raise dataset_missing("get_audio_dataloader(waveforms=True)", folder_path)
        from .synthetic import SyntheticWaveforms, collate_waveforms
        return DataLoader(SyntheticWaveforms(), batch_size=batch_size, shuffle=shuffle, collate_fn=collate_waveforms)
    if folder_path and os.path.isdir(folder_path):
        ds = AudioDataset(folder_path)
    elif synthetic_requested():
        ds = SyntheticAudio()
    else:
        raise dataset_missing("get_audio_dataloader", folder_path)
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, collate_fn=collate_fn)
