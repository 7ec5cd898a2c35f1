"""GPU audio front-end: waveforms resident in HBM -> the MFCC tensors XceptionLSTMA consumes.

The reference computes MFCCs offline with librosa (wavfake_audio_dataset.py:17-19,40-44:
``librosa.feature.mfcc(y, sr=16000, n_mfcc=13, n_fft=400, hop_length=160).T``), stores 120-frame ``.npy`` files and the
loader repeats the plane into three channels (audio_dataloader.py:20-28).  ``MFCC`` does the same arithmetic in two
kernels of libxcp_sm100.so (csrc/mfcc.cu) so configs 4/5 can run end to end from 16 kHz waveforms:

    front = MFCC().to("cuda")
    feats = model.extract_features(front.clips(waveforms, frames=120), device)     # (B,120,3,13) -> (B,120,2048)

Only constants are prepared on the host (the Slaney mel filterbank of ``librosa.filters.mel``); there is no CPU path.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import XcpError


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm="slaney")`` -> float32 [n_mels, 1 + n_fft//2]:
    Slaney's mel scale (66.67 Hz per mel below 1 kHz, log-spaced above with 27 mels per factor 6.4), triangular
    filters on the FFT bin centres, each scaled by 2 / bandwidth."""
    fmax = 0.5 * sr if fmax is None else float(fmax)
    lin = 200.0 / 3.0
    knee_hz, knee_mel, log_k = 1000.0, 1000.0 / lin, math.log(6.4) / 27.0

    def to_mel(f):
        return f / lin if f < knee_hz else knee_mel + math.log(f / knee_hz) / log_k

    edges_mel = np.linspace(to_mel(fmin), to_mel(fmax), n_mels + 2)
    edges = np.where(edges_mel < knee_mel, edges_mel * lin, knee_hz * np.exp(log_k * (edges_mel - knee_mel)))
    bins = np.arange(1 + n_fft // 2, dtype=np.float64) * (sr / float(n_fft))
    lo, mid, hi = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    rising = (bins[None, :] - lo) / (mid - lo)
    falling = (hi - bins[None, :]) / (hi - mid)
    tri = np.clip(np.minimum(rising, falling), 0.0, None)
    return (tri * (2.0 / (hi - lo))).astype(np.float32)


class MFCC(nn.Module):
    """Waveform [B, L] (or [L]) fp32 on a B200 -> MFCC [B, T, n_mfcc], T = 1 + L // hop_length."""

    def __init__(self, sr: int = 16000, n_mfcc: int = 13, n_fft: int | None = None, hop_length: int | None = None, n_mels: int = 128,
                 pad_mode: str = "constant", amin: float = 1e-10, top_db: float = 80.0):
        super().__init__()
        self.sr, self.n_mfcc, self.n_mels = sr, n_mfcc, n_mels
        self.n_fft = int(0.025 * sr) if n_fft is None else n_fft                 # 25 ms window (wavfake_audio_dataset.py:18)
        self.hop_length = int(0.010 * sr) if hop_length is None else hop_length   # 10 ms hop     (wavfake_audio_dataset.py:19)
        if pad_mode not in ("constant", "reflect"):
            raise XcpError("MFCC: pad_mode 'constant' (librosa >= 0.10) or 'reflect' (older librosa), got %r" % (pad_mode,))
        self.pad_mode, self.amin, self.top_db = pad_mode, float(amin), float(top_db)
        fb = slaney_mel_filterbank(sr, self.n_fft, n_mels)
        self.register_buffer("melfb_t", torch.from_numpy(np.ascontiguousarray(fb.T)), persistent=False)    # [bins, mels]

    def num_frames(self, samples: int) -> int:
        return 1 + samples // self.hop_length

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        if not wav.is_cuda or not self.melfb_t.is_cuda:
            raise XcpError("MFCC: waveform and module must live on a CUDA device (this package has no CPU path)")
        squeeze = wav.dim() == 1
        if squeeze:
            wav = wav.unsqueeze(0)
        if wav.dim() != 2:
            raise XcpError("MFCC: expected waveforms of shape [B, L], got %s" % (tuple(wav.shape),))
        wav = wav.to(torch.float32).contiguous()
        B, L = wav.shape
        T = self.num_frames(L)
        dev = wav.device
        logmel = torch.empty((B, T, self.n_mels), device=dev, dtype=torch.float32)
        gmax = torch.empty((B,), device=dev, dtype=torch.int32)
        out = torch.empty((B, T, self.n_mfcc), device=dev, dtype=torch.float32)
        _lib.call("xcp_mfcc", ctypes.c_void_p(wav.data_ptr()), B, L, ctypes.c_void_p(self.melfb_t.data_ptr()), self.n_fft,
                  self.hop_length, self.n_mels, self.n_mfcc, int(self.pad_mode == "reflect"), self.amin, self.top_db,
                  ctypes.c_void_p(logmel.data_ptr()), ctypes.c_void_p(gmax.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                  dev.index if dev.index is not None else torch.cuda.current_device(),
                  ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        return out[0] if squeeze else out

    def clips(self, wav: torch.Tensor, frames: int = 120, offset: int = 0) -> torch.Tensor:
        """The tensor AudioDataset yields (audio_dataloader.py:20-28): frames [offset, offset+frames) of every waveform's MFCC
        (the reference's train split is the first 120 frames, wavfake_audio_dataset.py:67-70), repeated to 3 channels:
        [B, frames, 3, n_mfcc] (an expanded view; XceptionLSTMA.extract_features makes it contiguous)."""
        m = self.forward(wav if wav.dim() == 2 else wav.unsqueeze(0))
        if m.shape[1] < offset + frames:
            raise XcpError("MFCC.clips: %d frames available, %d requested (the reference skips such files, "
                           "wavfake_audio_dataset.py:82-83)" % (m.shape[1], offset + frames))
        return m[:, offset:offset + frames].unsqueeze(2).expand(-1, -1, 3, -1)
