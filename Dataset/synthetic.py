"""Synthetic stand-ins with the reference loaders' tuple layouts (video_dataloader.py:53-68, audio_dataloader.py:34-47,
SURVEY App. C).  Values follow the real value ranges: frames in [0,1] (uint8/255), MFCC-like audio."""
import torch
from torch.utils.data import DataLoader, Dataset


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
    """The 'enhanced' collate of train_visual.py:563 -> (video, labels, seq_lengths)."""
    vids, labs = zip(*batch)
    video, labels = collate_clips(batch)
    return video, labels, torch.tensor([v.shape[0] for v in vids], dtype=torch.long)


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
