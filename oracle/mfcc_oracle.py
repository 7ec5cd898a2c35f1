"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.

CPU restatement (numpy, float64 internally) of the audio front-end the reference runs offline
(wavfake_audio_dataset.py:17-19,40-44):

    librosa.feature.mfcc(y=audio, sr=16000, n_mfcc=13, n_fft=int(0.025*sr)=400, hop_length=int(0.010*sr)=160).T

librosa is a third-party dependency that is NOT vendored in /root/reference and NOT installed in this image, and the
reference pins no version (it has no requirements file).  This file therefore restates librosa's published algorithm
(librosa 0.10.x defaults):

  feature.mfcc        -> melspectrogram(power=2.0, n_mels=128, fmin=0, fmax=sr/2, htk=False, norm="slaney")
                         -> power_to_db(ref=1.0, amin=1e-10, top_db=80.0) -> scipy.fftpack.dct(type=2, norm="ortho")[:n_mfcc]
  core.stft           -> win_length=n_fft, window="hann" (periodic), center=True, pad_mode="constant" (zeros; librosa < 0.10
                         used "reflect": both are implemented, `pad_mode`), frames t = 0 .. len(y)//hop
  filters.mel         -> Slaney mel scale (linear below 1 kHz, log above), triangular filters, area ("slaney") normalisation

PARITY PINNING: librosa itself cannot be run here, so the restatement is pinned against an INDEPENDENT implementation of
the same published algorithm that is installed -- `transformers.audio_utils` (`mel_filter_bank(norm="slaney",
mel_scale="slaney")`, `spectrogram(power=2, log_mel="dB", db_range=80)`), written to reproduce librosa's output -- and
against scipy's DCT (tests/test_mfcc_oracle_cpu.py).  Against librosa proper it is "parity unpinned"; DESIGN.md says so.

Only tests/, bench-side CPU legs and __graft_entry__.smoke() may import this file.
"""
from __future__ import annotations

import numpy as np


def hz_to_mel(f):
    """librosa.hz_to_mel(htk=False): Slaney's Auditory Toolbox scale."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int = 16000, n_fft: int = 400, n_mels: int = 128, fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """librosa.filters.mel(norm="slaney", htk=False) -> float32 [n_mels, 1 + n_fft//2]."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window("hann", n, fftbins=True)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def power_spectrogram(y: np.ndarray, n_fft: int = 400, hop: int = 160, pad_mode: str = "constant") -> np.ndarray:
    """|STFT|^2 with librosa.stft's framing (center=True): -> [1 + n_fft//2, T], T = 1 + len(y)//hop."""
    y = np.asarray(y, dtype=np.float64)
    ypad = np.pad(y, n_fft // 2, mode="constant" if pad_mode == "constant" else "reflect")
    T = 1 + (len(ypad) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    frames = ypad[idx] * hann_periodic(n_fft)[None, :]
    spec = np.fft.rfft(frames, n=n_fft, axis=1)
    return (spec.real ** 2 + spec.imag ** 2).T


def dct_ortho_matrix(n_out: int, n_in: int) -> np.ndarray:
    """Rows of scipy.fftpack.dct(type=2, norm="ortho"): out[c] = sum_m x[m] * D[c, m]."""
    m = np.arange(n_in)[None, :]
    c = np.arange(n_out)[:, None]
    D = np.cos(np.pi * (2 * m + 1) * c / (2.0 * n_in)) * np.sqrt(2.0 / n_in)
    D[0] *= np.sqrt(0.5)
    return D


def mfcc(y: np.ndarray, sr: int = 16000, n_mfcc: int = 13, n_fft: int = 400, hop: int = 160, n_mels: int = 128,
         amin: float = 1e-10, top_db: float = 80.0, pad_mode: str = "constant") -> np.ndarray:
    """wavfake_audio_dataset.py:43-44: librosa.feature.mfcc(...).T -> float32 [T, n_mfcc]."""
    S = mel_filterbank(sr, n_fft, n_mels).astype(np.float64) @ power_spectrogram(y, n_fft, hop, pad_mode)     # [n_mels, T]
    db = 10.0 * np.log10(np.maximum(amin, S))                  # ref = 1.0 contributes -10*log10(max(amin, 1)) = 0
    db = np.maximum(db, db.max() - top_db)
    return (dct_ortho_matrix(n_mfcc, n_mels) @ db).T.astype(np.float32)


def dataset_item(mfcc_frames: np.ndarray) -> np.ndarray:
    """audio_dataloader.py:20-28: (T, 13) -> (T, 3, 13) by channel repeat (what XceptionLSTMA.extract_features consumes)."""
    return np.repeat(mfcc_frames[:, None, :], 3, axis=1)
