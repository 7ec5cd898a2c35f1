"""Generates tests/golden/mfcc_golden.npz: MFCCs of seeded waveforms computed by an implementation that is INDEPENDENT of
oracle/mfcc_oracle.py -- transformers.audio_utils (mel_filter_bank / spectrogram, written to reproduce librosa) + scipy's
DCT -- following librosa.feature.mfcc's published pipeline (wavfake_audio_dataset.py:43).  librosa itself is not installed
in this image, so these vectors pin the oracle against a second implementation, not against librosa proper.

    python oracle/gen_mfcc_golden.py
"""
import os

import numpy as np
import scipy.fftpack
from transformers import audio_utils as au

SR, N_FFT, HOP, N_MELS, N_MFCC = 16000, 400, 160, 128, 13


def waveforms():
    rng = np.random.default_rng(2024)
    t = np.arange(int(0.8 * SR)) / SR
    noise = (rng.standard_normal(t.size) * 0.05).astype(np.float32)
    tones = (0.4 * np.sin(2 * np.pi * 220 * t) + 0.2 * np.sin(2 * np.pi * 3100 * t + 1.0) + 0.02 * rng.standard_normal(t.size)).astype(np.float32)
    gated = noise.copy()
    gated[2000:7000] = 0.0                      # digital silence: exercises amin and the top_db clip
    return {"noise": noise, "tones": tones, "gated": gated, "short": noise[:1234].copy()}


def reference_mfcc(y, pad_mode):
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT)
    fb = au.mel_filter_bank(1 + N_FFT // 2, N_MELS, 0.0, SR / 2.0, SR, norm="slaney", mel_scale="slaney")
    db = au.spectrogram(y.astype(np.float64), window, N_FFT, HOP, fft_length=N_FFT, power=2.0, center=True, pad_mode=pad_mode,
                        mel_filters=fb, mel_floor=1e-10, log_mel="dB", reference=1.0, min_value=1e-10, db_range=80.0,
                        dtype=np.float64)
    return scipy.fftpack.dct(db, axis=0, type=2, norm="ortho")[:N_MFCC].T.astype(np.float32)


def main():
    out = {}
    for name, y in waveforms().items():
        out["wav::" + name] = y
        for pm in ("constant", "reflect"):
            out["mfcc::%s::%s" % (name, pm)] = reference_mfcc(y, pm)
    fb = au.mel_filter_bank(1 + N_FFT // 2, N_MELS, 0.0, SR / 2.0, SR, norm="slaney", mel_scale="slaney")
    out["melfb"] = fb.T.astype(np.float32)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mfcc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("mfcc")})


if __name__ == "__main__":
    main()
