"""CPU tier for the audio front-end: the numpy oracle (oracle/mfcc_oracle.py, librosa's published
algorithm) against golden vectors made by an independent implementation (oracle/gen_mfcc_golden.py: transformers.audio_utils
+ scipy), and the host-built constants of the product (audio_frontend.slaney_mel_filterbank) against both."""
import os

import numpy as np
import pytest

from oracle import mfcc_oracle as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "mfcc_golden.npz"))


@pytest.mark.parametrize("name", ["noise", "tones", "gated", "short"])
@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_oracle_matches_independent_golden(gold, name, pad_mode):
    ref = gold["mfcc::%s::%s" % (name, pad_mode)]
    got = M.mfcc(gold["wav::" + name], pad_mode=pad_mode)
    assert got.shape == ref.shape == (1 + len(gold["wav::" + name]) // 160, 13)
    assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def test_mel_filterbank_and_dct_constants(gold):
    fb = M.mel_filterbank()
    assert fb.shape == (128, 201) and np.abs(fb - gold["melfb"]).max() < 1e-7
    from multimodal_deepfake_detection_b200.audio_frontend import slaney_mel_filterbank
    mine = slaney_mel_filterbank(16000, 400, 128)
    assert mine.dtype == np.float32 and np.abs(mine - gold["melfb"]).max() < 1e-7
    # other geometries the module accepts
    for sr, n_fft, n_mels in ((22050, 512, 64), (8000, 200, 40)):
        assert np.abs(slaney_mel_filterbank(sr, n_fft, n_mels) - M.mel_filterbank(sr, n_fft, n_mels)).max() < 1e-7
    scipy_fftpack = pytest.importorskip("scipy.fftpack")
    x = np.random.default_rng(1).standard_normal((128, 5))
    assert np.abs(M.dct_ortho_matrix(13, 128) @ x - scipy_fftpack.dct(x, axis=0, type=2, norm="ortho")[:13]).max() < 1e-10


def test_live_independent_implementation_when_available(gold):
    """Same comparison against transformers.audio_utils computed now (guards the committed fixture against drift)."""
    au = pytest.importorskip("transformers.audio_utils")
    scipy_fftpack = pytest.importorskip("scipy.fftpack")
    y = gold["wav::tones"].astype(np.float64)
    fb = au.mel_filter_bank(201, 128, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")
    db = au.spectrogram(y, M.hann_periodic(400), 400, 160, fft_length=400, power=2.0, center=True, pad_mode="constant", mel_filters=fb,
                        mel_floor=1e-10, log_mel="dB", reference=1.0, min_value=1e-10, db_range=80.0, dtype=np.float64)
    ref = scipy_fftpack.dct(db, axis=0, type=2, norm="ortho")[:13].T
    assert np.abs(M.mfcc(y) - ref).max() < 1e-3


def test_dataset_item_layout():
    m = M.mfcc(np.random.default_rng(3).standard_normal(16000 * 2).astype(np.float32) * 0.1)
    item = M.dataset_item(m[:120])
    assert item.shape == (120, 3, 13) and np.array_equal(item[:, 0], item[:, 2])
