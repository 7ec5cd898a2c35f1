"""GPU audio front-end (csrc/mfcc.cu, audio_frontend.MFCC) against the numpy oracle and the committed golden vectors
(reference arithmetic: librosa.feature.mfcc as called in wavfake_audio_dataset.py:43).

Tolerance: the kernels compute in fp32 (direct DFT, fp32 mel sums, log10f); values range over +-1000, the bound is
1e-3 absolute (1e-6 of the range; measured 1.2e-4), stated per assert."""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mfcc_oracle as M  # noqa: E402
from multimodal_deepfake_detection_b200 import XceptionLSTMA  # noqa: E402
from multimodal_deepfake_detection_b200._lib import XcpError  # noqa: E402
from multimodal_deepfake_detection_b200.audio_frontend import MFCC  # noqa: E402

DEV = "cuda"
ATOL = 1e-3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "mfcc_golden.npz"))


@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_mfcc_matches_golden_and_oracle(gold, pad_mode):
    front = MFCC(pad_mode=pad_mode).to(DEV)
    worst = 0.0
    for name in ("noise", "tones", "gated", "short"):
        y = gold["wav::" + name]
        got = front(torch.from_numpy(y).to(DEV)).cpu().numpy()
        ref = gold["mfcc::%s::%s" % (name, pad_mode)]
        assert got.shape == ref.shape
        worst = max(worst, float(np.abs(got - ref).max()), float(np.abs(got - M.mfcc(y, pad_mode=pad_mode)).max()))
    print("mfcc %s: worst abs err %.2e" % (pad_mode, worst))
    assert worst < ATOL


def test_mfcc_batch_has_per_waveform_top_db_and_ragged_length():
    """power_to_db clips at (max - 80 dB) of each file's own spectrogram: a loud and a quiet waveform in one batch must not
    share the maximum.  L is not a multiple of the hop."""
    rng = np.random.default_rng(7)
    L = 16000 + 77
    loud = (rng.standard_normal(L) * 0.5).astype(np.float32)
    quiet = (rng.standard_normal(L) * 1e-4).astype(np.float32)
    quiet[4000:9000] = 0.0
    mixed = (0.3 * np.sin(2 * np.pi * 440 * np.arange(L) / 16000)).astype(np.float32)
    wav = np.stack([loud, quiet, mixed])
    front = MFCC().to(DEV)
    got = front(torch.from_numpy(wav).to(DEV)).cpu().numpy()
    assert got.shape == (3, 1 + L // 160, 13)
    for b in range(3):
        assert np.abs(got[b] - M.mfcc(wav[b])).max() < ATOL, b
    again = front(torch.from_numpy(wav).to(DEV)).cpu().numpy()
    assert np.array_equal(got, again)                                  # deterministic (max via atomicMax is order-free)


def test_mfcc_other_geometry():
    rng = np.random.default_rng(8)
    y = (rng.standard_normal(22050) * 0.1).astype(np.float32)
    front = MFCC(sr=22050, n_mfcc=20, n_fft=512, hop_length=256, n_mels=64).to(DEV)
    got = front(torch.from_numpy(y).to(DEV)).cpu().numpy()
    ref = M.mfcc(y, sr=22050, n_mfcc=20, n_fft=512, hop=256, n_mels=64)
    assert got.shape == ref.shape and np.abs(got - ref).max() < ATOL


def test_mfcc_clips_feed_xception_lstma_like_the_npy_files():
    """Waveform -> MFCC.clips() -> XceptionLSTMA.extract_features == the reference's offline route
    (librosa .npy -> AudioDataset channel repeat, audio_dataloader.py:20-28) with the oracle standing in for librosa."""
    rng = np.random.default_rng(9)
    wav = (rng.standard_normal((2, 3200)) * 0.1).astype(np.float32)
    front = MFCC().to(DEV)
    clips = front.clips(torch.from_numpy(wav).to(DEV), frames=6, offset=2)
    assert clips.shape == (2, 6, 3, 13)
    ref_items = np.stack([M.dataset_item(M.mfcc(w)[2:8]) for w in wav])
    assert np.abs(clips.cpu().numpy() - ref_items).max() < ATOL
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = XceptionLSTMA(32).to(DEV).eval()
    with torch.no_grad():
        a = m(m.extract_features(clips, torch.device(DEV)))
        b = m(m.extract_features(torch.from_numpy(ref_items).to(DEV), torch.device(DEV)))
    assert a.shape == (2, 1) and (a - b).abs().max().item() < 2e-3
    with pytest.raises(XcpError):
        front.clips(torch.from_numpy(wav).to(DEV), frames=120)         # too short: the reference skips such files


def test_mfcc_rejects_cpu_and_bad_arguments():
    with pytest.raises(XcpError):
        MFCC()(torch.zeros(1, 1600))                                    # no CPU path
    with pytest.raises(XcpError):
        MFCC(pad_mode="edge")
    with pytest.raises(XcpError):
        MFCC(n_fft=4096).to(DEV)(torch.zeros(1, 16000, device=DEV))     # n_fft beyond the kernel's shared-memory table
