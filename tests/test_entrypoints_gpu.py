"""The reference's entry points (train_visual.main, test_visual.test, train_audio, train_au_face.main, test_au_face.main,
train_au_patch / test_au_patch) run end to end on the sm_100a path on small synthetic data: one bounded epoch each,
checkpoint formats as in the reference, evaluation scripts load what the training scripts wrote."""
import importlib
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def small_run(tmp_path, monkeypatch):
    monkeypatch.chdir(ROOT)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    env = {"XCP_SYNTHETIC": "1", "XCP_EPOCHS": "2", "XCP_FREEZE_EPOCHS": "1", "XCP_SYNTH_CLIPS": "8", "XCP_FRAME_SIZE": "75", "XCP_WORKERS": "0",
           "XCP_CKPT_DIR": str(tmp_path / "ck"), "XCP_OUTPUT_DIR": str(tmp_path / "out"), "XCP_EVAL_EVERY": "1",
           "XCP_AUDIO_HIDDEN": "64", "XCP_FUSION_HIDDEN": "64", "XCP_MAX_FRAMES": "4", "XCP_PATCH_STEPS": "6", "XCP_N_MELS": "16"}
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    return tmp_path


def _fresh(name):
    sys.modules.pop(name, None)
    return importlib.import_module(name)


def test_train_then_test_visual(small_run):
    tv = _fresh("train_visual")
    best = tv.main()
    ck = torch.load(small_run / "ck" / "XceptionLSTMV_ArcFace_Best.pth")
    assert set(ck) == {"model", "arcface"} and len(ck["model"]) == 288 and ck["arcface"]["weight"].shape == (2, 128)
    assert best == best and best < 50.0          # finite loss
    m = _fresh("test_visual").test()
    assert 0.0 <= m["ACC"] <= 1.0 and 0.0 <= m["AUC"] <= 1.0


def test_train_audio(small_run):
    best = _fresh("train_audio").main()
    sd = torch.load(small_run / "ck" / "best_model_audio.pth")
    assert len(sd) == 288 and sd["lstm.weight_ih_l0"].shape == (256, 2048)
    assert best == best and best < 50.0


def test_train_audio_from_waveforms(small_run, monkeypatch):
    """the same script fed with raw 16 kHz waveforms, MFCCs computed on the GPU every step."""
    monkeypatch.setenv("XCP_AUDIO_FROM_WAV", "1")
    monkeypatch.setenv("XCP_EPOCHS", "1")
    best = _fresh("train_audio").main()
    assert (small_run / "ck" / "best_model_audio.pth").exists()
    assert best == best and best < 50.0


def test_train_then_test_au_face(small_run):
    auc = _fresh("train_au_face").main()
    ck = torch.load(small_run / "ck" / "auface_cross_best_auc_arcface_cb.pth")
    assert {"model", "embed", "arcface", "best_auc"} <= set(ck)
    assert 0.0 <= auc <= 1.0
    m = _fresh("test_au_face").main()
    assert 0.0 <= m["AUC"] <= 1.0
    assert (small_run / "out" / "eval_scores_and_labels.npz").exists()


def test_train_then_test_au_patch(small_run):
    best = _fresh("train_au_patch").main()
    assert best == best and best < 50.0
    m = _fresh("test_au_patch").main()
    assert 0.0 <= m["AUC"] <= 1.0
