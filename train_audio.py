"""train_audio.py -- BCE training of XceptionLSTMA(512) on MFCC sequences (the reference's module-level script,
train_audio.py:1-94, wrapped in `main()`), on the sm_100a path.

Protocol kept: batch 8, BCELoss on the sigmoid output, Adam(lr=1e-4), evaluation every 10 epochs with
ReduceLROnPlateau(min, 0.5, patience 5), best `state_dict` to Checkpoints/best_model_audio.pth, early stopping after
10 stale evaluations.  (The reference wraps the model in nn.DataParallel on multi-GPU hosts; here multi-GPU training is
one process per GPU with ddp.GradBucketer -- see bench.py -- and this script drives a single device.)"""
import os

import torch

from Dataset.audio_dataloader import get_audio_dataloader
from Models.XceptionLSTMA import XceptionLSTMA
from multimodal_deepfake_detection_b200 import FusedAdam
from multimodal_deepfake_detection_b200.audio_frontend import MFCC
from multimodal_deepfake_detection_b200.loops import audio_epoch, env_int, require_b200

CKPT_DIR = os.environ.get("XCP_CKPT_DIR", "Checkpoints")


def main():
    device = require_b200()
    from_wav = bool(env_int("XCP_AUDIO_FROM_WAV", 0))      # 1: raw 16 kHz waveforms -> MFCC on the GPU (no offline librosa pass)
    frontend = MFCC().to(device) if from_wav else None
    train_dataloader = get_audio_dataloader("Dataset/processed_audio/train", batch_size=8, shuffle=False, waveforms=from_wav)
    eval_dataloader = get_audio_dataloader("Dataset/processed_audio/eval", batch_size=8, shuffle=False, waveforms=from_wav)
    model = XceptionLSTMA(hidden_dim=env_int("XCP_AUDIO_HIDDEN", 512)).to(device)
    optimizer = FusedAdam(model.parameters(), lr=0.0001)
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.5, patience=5)
    best_eval_loss, early_stop_count, patience = float("inf"), 0, 10
    num_epochs, eval_every = env_int("XCP_EPOCHS", 100), env_int("XCP_EVAL_EVERY", 10)
    for epoch in range(num_epochs):
        model.train()
        loss, _ = audio_epoch(model, train_dataloader, device, optimizer, frontend=frontend)
        print(f"Epoch [{epoch + 1}/{num_epochs}], Train Loss: {loss:.4f}")
        if (epoch + 1) % eval_every == 0:
            model.eval()
            eval_loss, eval_accuracy = audio_epoch(model, eval_dataloader, device, None, frontend=frontend)
            print(f"Evaluation Loss: {eval_loss:.4f}, Accuracy: {eval_accuracy:.4f}")
            prev_lr = optimizer.param_groups[0]["lr"]
            scheduler.step(eval_loss)
            if optimizer.param_groups[0]["lr"] < prev_lr:
                print(f"Learning rate reduced to {optimizer.param_groups[0]['lr']:.6f}")
            if eval_loss < best_eval_loss:
                best_eval_loss, early_stop_count = eval_loss, 0
                print("New best model found. Saving...")
                os.makedirs(CKPT_DIR, exist_ok=True)
                torch.save(model.state_dict(), os.path.join(CKPT_DIR, "best_model_audio.pth"))
            else:
                early_stop_count += 1
            if early_stop_count >= patience:
                print("Early stopping triggered. Training stopped.")
                break
    return best_eval_loss


if __name__ == "__main__":
    main()
