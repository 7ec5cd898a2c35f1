"""Synthetic stand-ins with the reference loaders' tuple layouts (video_dataloader.py:53-68, audio_dataloader.py:34-47,
SURVEY App. C).  Values follow the real value ranges: frames in [0,1] (uint8/255), MFCC-like audio."""
import os

import torch
from torch.utils.data import DataLoader, Dataset


def synthetic_requested() -> bool:
    """Synthetic data is served only on explicit request (XCP_SYNTHETIC=1): a missing / mistyped dataset path must fail
    like the reference's loaders do (os.listdir raises), not train a confident model on noise and overwrite checkpoints."""
    return os.environ.get("XCP_SYNTHETIC", "0") == "1"


def dataset_missing(what: str, path) -> FileNotFoundError:
    return FileNotFoundError("%s: no usable dataset at %r (expected a folder of `real_*.npy` / `fake_*.npy` files as written by the "
                             "reference's pre-processors); set XCP_SYNTHETIC=1 to run on synthetic data instead" % (what, path))


def label_from_name(path: str) -> int:
    """video_dataloader.py:29-32 / audio_dataloader.py:22: `<label>_...npy`, 'real' -> 0, anything else -> 1."""
    return 0 if os.path.basename(path).split("_")[0].lower() == "real" else 1


class SyntheticClips(Dataset):
    def __init__(self, n=64, frames=16, size=299, seed=0, variable_length=False, raw_uint8=False):
        self.n, self.frames, self.size, self.seed, self.var = n, frames, size, seed, variable_length
        self.raw_uint8 = raw_uint8            # uint8 (T,H,W,3) frames as stored on disk instead of float (T,3,H,W)/255
        g = torch.Generator().manual_seed(seed)
        self.labels = torch.randint(0, 2, (n,), generator=g).tolist()
        self.samples = [("synthetic_%d" % i, self.labels[i], None) for i in range(n)]   # label at index 1 (train_visual.py:525)
        self.all_labels = self.labels

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        t = self.frames if not self.var else max(2, self.frames - (i % 3))
        if self.raw_uint8:
            return (torch.randint(0, 256, (t, self.size, self.size, 3), generator=g, dtype=torch.uint8),
                    torch.tensor(self.labels[i], dtype=torch.float32))
        return torch.rand(t, 3, self.size, self.size, generator=g), torch.tensor(self.labels[i], dtype=torch.float32)


def collate_clips(batch):
    """video_dataloader.py:53-68: zero-pad to the longest clip -> (B,Tmax,3,H,W), labels (B,)."""
    vids, labs = zip(*batch)
    tmax = max(v.shape[0] for v in vids)
    out = torch.zeros(len(vids), tmax, *vids[0].shape[1:], dtype=vids[0].dtype)     # float (T,3,H,W) or raw uint8 (T,H,W,3) clips
    for i, v in enumerate(vids):
        out[i, :v.shape[0]] = v
    return out, torch.stack(labs)


def collate_clips_with_lengths(batch):
    """The 'enhanced' collate of train_visual.py:563 -> (video, labels (B,), seq_lengths): the labels feed CrossEntropyLoss
    directly (train_visual.py:571-572), so they are flattened whatever the per-item shape."""
    vids, labs = zip(*batch)
    video, labels = collate_clips(batch)
    return video, labels.reshape(-1), torch.tensor([v.shape[0] for v in vids], dtype=torch.long)


class SyntheticAudio(Dataset):
    def __init__(self, n=64, steps=120, n_mfcc=13, seed=0):
        self.n, self.steps, self.n_mfcc, self.seed = n, steps, n_mfcc, seed
        g = torch.Generator().manual_seed(seed)
        self.labels = torch.randint(0, 2, (n,), generator=g).tolist()

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 7919 + i)
        mf = torch.randn(self.steps, 1, self.n_mfcc, generator=g) * 20.0
        mf[:, :, 0] = mf[:, :, 0] * 5.0 - 300.0                       # c0 ~ N(-300, 100)
        return mf.repeat(1, 3, 1), torch.tensor([float(self.labels[i])])   # (T,3,13) by channel repeat, audio_dataloader.py:25-26


class SyntheticWaveforms(Dataset):
    """16 kHz mono waveforms (what `ffmpeg -ar 16000 -ac 1` hands librosa, wavfake_audio_dataset.py:30-41) for the GPU
    front-end route: (samples,) fp32 + label; `frames` MFCC frames need (frames - 1) * 160 samples."""

    def __init__(self, n=64, frames=120, sr=16000, seed=0):
        self.n, self.samples, self.sr, self.seed = n, (frames - 1) * int(0.010 * sr) + int(0.010 * sr) // 2, sr, seed
        g = torch.Generator().manual_seed(seed)
        self.labels = torch.randint(0, 2, (n,), generator=g).tolist()

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 104729 + i)
        t = torch.arange(self.samples) / self.sr
        f0 = 110.0 + 200.0 * torch.rand((), generator=g)
        wav = 0.3 * torch.sin(2 * torch.pi * f0 * t) + 0.1 * torch.sin(2 * torch.pi * 3.1 * f0 * t) + 0.02 * torch.randn(self.samples, generator=g)
        return wav.float(), torch.tensor([float(self.labels[i])])


def collate_waveforms(batch):
    wavs, labs = zip(*batch)
    return torch.stack(wavs), torch.stack(labs)


def collate_audio(batch):
    feats, labs = zip(*batch)
    tmax = max(f.shape[0] for f in feats)
    out = torch.zeros(len(feats), tmax, *feats[0].shape[1:])
    for i, f in enumerate(feats):
        out[i, :f.shape[0]] = f
    return out, torch.stack(labs)


def synthetic_loader(ds, batch_size, shuffle, collate_fn):
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, collate_fn=collate_fn, pin_memory=torch.cuda.is_available())
