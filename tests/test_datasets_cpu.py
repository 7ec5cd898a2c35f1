"""Host-side data layer (CPU tier): the loaders keep the reference's tuple layouts (video_dataloader.py:6-68,
audio_dataloader.py:6-47) and add the two ingest routes of SURVEY.md §8 rows f-2 / f-4 (raw uint8 frames, raw waveforms)."""
import subprocess
import sys

import numpy as np
import pytest
import torch

from Dataset import audio_dataloader, video_dataloader, video_dataloader_enhanced
from Dataset.synthetic import SyntheticAudio, SyntheticClips, SyntheticWaveforms, collate_audio, collate_clips_with_lengths


def test_face_dataset_float_and_raw_uint8(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (5, 8, 8, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (3, 8, 8, 3), dtype=np.uint8)
    np.save(tmp_path / "real_a.npy", a)
    np.save(tmp_path / "fake_b.npy", b)
    ds = video_dataloader.FaceDataset(str(tmp_path))                       # sorted: fake_b, real_a
    x, y = ds[0]
    assert x.shape == (3, 3, 8, 8) and x.dtype == torch.float32 and float(y) == 1.0 and y.shape == (1,)     # video_dataloader.py:37
    assert torch.equal(x, torch.from_numpy(b).permute(0, 3, 1, 2).float() / 255.0)         # video_dataloader.py:27-35
    x, y = ds[1]
    assert float(y) == 0.0 and x.shape == (5, 3, 8, 8)
    raw = video_dataloader.FaceDataset(str(tmp_path), raw_uint8=True)
    xr, _ = raw[1]
    assert xr.dtype == torch.uint8 and xr.shape == (5, 8, 8, 3) and torch.equal(xr, torch.from_numpy(a))
    video, labels = video_dataloader.collate_fn([raw[0], raw[1]])           # zero-pad to the longest clip, dtype kept
    assert video.shape == (2, 5, 8, 8, 3) and video.dtype == torch.uint8 and labels.tolist() == [[1.0], [0.0]]   # (B,1) like the reference
    assert torch.equal(video[0, 3:], torch.zeros(2, 8, 8, 3, dtype=torch.uint8))
    loader = video_dataloader.get_face_dataloader(str(tmp_path), batch_size=2, shuffle=False)
    v, l = next(iter(loader))
    assert v.shape == (2, 5, 3, 8, 8) and v.dtype == torch.float32 and 0.0 <= float(v.min()) and float(v.max()) <= 1.0


def test_audio_dataset_channel_repeat_and_collate(tmp_path):
    m = np.random.default_rng(1).standard_normal((120, 13)).astype(np.float32)
    np.save(tmp_path / "real_x.npy", m)
    np.save(tmp_path / "fake_y.npy", m[:24])
    ds = audio_dataloader.AudioDataset(str(tmp_path))
    x, y = ds[1]
    assert x.shape == (120, 3, 13) and y.shape == (1,) and float(y) == 0.0                  # audio_dataloader.py:20-28
    assert torch.equal(x[:, 0], x[:, 1]) and torch.equal(x[:, 1], x[:, 2]) and torch.equal(x[:, 0], torch.from_numpy(m))
    feats, labs = collate_audio([ds[0], ds[1]])
    assert feats.shape == (2, 120, 3, 13) and labs.shape == (2, 1) and torch.equal(feats[0, 24:], torch.zeros(96, 3, 13))


def test_labels_follow_the_reference_file_name_rule(tmp_path):
    """video_dataloader.py:29-32: label = 0 iff the part before the first '_' is 'real' (case-insensitive)."""
    for name in ("real_1.npy", "REAL_2.npy", "realistic_3.npy", "fake_4.npy", "real.npy"):
        np.save(tmp_path / name, np.zeros((1, 4, 4, 3), dtype=np.uint8))
    ds = video_dataloader.FaceDataset(str(tmp_path))
    got = {f.split("/")[-1]: float(ds[i][1]) for i, f in enumerate(ds.files)}
    assert got == {"real_1.npy": 0.0, "REAL_2.npy": 0.0, "realistic_3.npy": 1.0, "fake_4.npy": 1.0, "real.npy": 1.0}


def test_missing_dataset_raises_unless_synthetic_is_requested(tmp_path, monkeypatch):
    """a mistyped path must not silently train on noise (the reference's loaders raise from os.listdir)."""
    monkeypatch.delenv("XCP_SYNTHETIC", raising=False)
    missing = str(tmp_path / "nope")
    with pytest.raises(FileNotFoundError):
        video_dataloader.get_face_dataloader(missing)
    with pytest.raises(FileNotFoundError):
        audio_dataloader.get_audio_dataloader(missing)
    with pytest.raises(FileNotFoundError):
        audio_dataloader.get_audio_dataloader(missing, waveforms=True)
    with pytest.raises(FileNotFoundError):
        video_dataloader_enhanced.get_face_dataloader(missing)
    (tmp_path / "raw").mkdir()
    (tmp_path / "raw" / "clip.mp4").write_bytes(b"")
    with pytest.raises(FileNotFoundError, match="not decoded"):
        video_dataloader_enhanced.get_face_dataloader(str(tmp_path / "raw"))
    from Dataset.AuVidDataset import get_joint_dataloader
    with pytest.raises(FileNotFoundError):
        get_joint_dataloader(video_root=missing)
    monkeypatch.setenv("XCP_SYNTHETIC", "1")
    assert len(audio_dataloader.get_audio_dataloader(missing).dataset) > 0
    assert len(get_joint_dataloader(video_root=missing)[0].dataset) > 0


def test_synthetic_sets_have_the_loaders_layouts(monkeypatch):
    monkeypatch.setenv("XCP_SYNTHETIC", "1")
    clips = SyntheticClips(n=4, frames=5, size=16, variable_length=True)
    video, labels, lengths = collate_clips_with_lengths([clips[i] for i in range(4)])
    assert video.shape == (4, 5, 3, 16, 16) and labels.shape == (4,) and lengths.tolist() == [5, 4, 3, 5]
    assert clips.samples[2][1] == clips.labels[2]                                            # label at index 1 (train_visual.py:525)
    audio = SyntheticAudio(n=3, steps=7)
    x, y = audio[0]
    assert x.shape == (7, 3, 13) and torch.equal(x[:, 0], x[:, 2]) and y.shape == (1,)
    wav = SyntheticWaveforms(n=3, frames=120)
    w, y = wav[1]
    assert w.dtype == torch.float32 and 1 + w.numel() // 160 == 120 and float(w.abs().max()) < 1.0
    loader = audio_dataloader.get_audio_dataloader(None, batch_size=2, shuffle=False, waveforms=True)
    wb, yb = next(iter(loader))
    assert wb.shape == (2, w.numel()) and yb.shape == (2, 1)
    assert torch.equal(SyntheticWaveforms(n=3)[1][0], w)                                     # seeded: same item every time


def test_enhanced_loader_is_reproducible_across_processes(monkeypatch):
    monkeypatch.setenv("XCP_SYNTHETIC", "1")
    code = ("from Dataset.video_dataloader_enhanced import get_face_dataloader as g;"
            "d=g(subset='test',frame_size=(8,8),synthetic_clips=2).dataset;print(d.seed, float(d[0][0].sum()))")
    outs = {subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env={"PYTHONHASHSEED": s, "PATH": "/usr/bin:/bin", "XCP_SYNTHETIC": "1"},
                           cwd=str(__import__("pathlib").Path(__file__).resolve().parents[1])).stdout for s in ("1", "2")}
    assert len(outs) == 1 and "" not in outs, outs
    loader = video_dataloader_enhanced.get_face_dataloader(subset="train", frame_size=(8, 8), synthetic_clips=3, batch_size=3)
    video, labels, lengths = next(iter(loader))
    assert video.shape[0] == 3 and video.shape[2:] == (3, 8, 8) and lengths.shape == (3,)
