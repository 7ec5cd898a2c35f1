"""train_audio.py -- BCE training of XceptionLSTMA(512) on MFCC sequences (the reference's module-level script,
train_audio.py:1-94, wrapped in `main()`), on the sm_100a path.

Protocol kept: batch 8, BCELoss on the sigmoid output, Adam(lr=1e-4), evaluation every 10 epochs with
ReduceLROnPlateau(min, 0.5, patience 5), best `state_dict` to Checkpoints/best_model_audio.pth, early stopping after
10 stale evaluations.  The reference wraps the model in nn.DataParallel on multi-GPU hosts (train_audio.py:16-18); here
multi-GPU training is `torchrun --nproc-per-node N train_audio.py`: one process per GPU, the training files sharded by a
DistributedSampler (batch 8 per GPU), gradients averaged by ddp.GradBucketer, rank 0 prints and saves."""
import os

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader
from torch.utils.data.distributed import DistributedSampler

from Dataset.audio_dataloader import get_audio_dataloader
from Models.XceptionLSTMA import XceptionLSTMA
from multimodal_deepfake_detection_b200 import FusedAdam
from multimodal_deepfake_detection_b200.audio_frontend import MFCC
from multimodal_deepfake_detection_b200.ddp import GradBucketer
from multimodal_deepfake_detection_b200.loops import audio_epoch, broadcast_module_state, env_int, init_data_parallel, require_b200

CKPT_DIR = os.environ.get("XCP_CKPT_DIR", "Checkpoints")


def main():
    world, rank = init_data_parallel()
    device = require_b200()
    say = print if rank == 0 else (lambda *a, **k: None)
    from_wav = bool(env_int("XCP_AUDIO_FROM_WAV", 0))      # 1: raw 16 kHz waveforms -> MFCC on the GPU (no offline librosa pass)
    frontend = MFCC().to(device) if from_wav else None
    train_dataloader = get_audio_dataloader("Dataset/processed_audio/train", batch_size=8, shuffle=False, waveforms=from_wav)
    eval_dataloader = get_audio_dataloader("Dataset/processed_audio/eval", batch_size=8, shuffle=False, waveforms=from_wav)
    model = XceptionLSTMA(hidden_dim=env_int("XCP_AUDIO_HIDDEN", 512)).to(device)
    optimizer = FusedAdam(model.parameters(), lr=0.0001)
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.5, patience=5)
    sampler, after_backward = None, None
    if world > 1:
        ds = train_dataloader.dataset
        sampler = DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=False, drop_last=True)
        train_dataloader = DataLoader(ds, batch_size=8, sampler=sampler, collate_fn=train_dataloader.collate_fn)
        broadcast_module_state([model])
        after_backward = GradBucketer(model, backbone=model.feature_extractor).finish
    best_eval_loss, early_stop_count, patience = float("inf"), 0, 10
    num_epochs, eval_every = env_int("XCP_EPOCHS", 100), env_int("XCP_EVAL_EVERY", 10)
    for epoch in range(num_epochs):
        model.train()
        if sampler is not None:
            sampler.set_epoch(epoch)
        loss, _ = audio_epoch(model, train_dataloader, device, optimizer, frontend=frontend, after_backward=after_backward)
        say(f"Epoch [{epoch + 1}/{num_epochs}], Train Loss: {loss:.4f}")
        if (epoch + 1) % eval_every == 0:
            model.eval()
            if world > 1:                                 # rank 0's running statistics and numbers drive every rank's control flow
                broadcast_module_state([model], buffers_only=True)
            eval_loss, eval_accuracy = audio_epoch(model, eval_dataloader, device, None, frontend=frontend)
            if world > 1:
                ctl = torch.tensor([eval_loss], device=device, dtype=torch.float64)
                dist.broadcast(ctl, 0)
                eval_loss = float(ctl[0])
            say(f"Evaluation Loss: {eval_loss:.4f}, Accuracy: {eval_accuracy:.4f}")
            prev_lr = optimizer.param_groups[0]["lr"]
            scheduler.step(eval_loss)
            if optimizer.param_groups[0]["lr"] < prev_lr:
                say(f"Learning rate reduced to {optimizer.param_groups[0]['lr']:.6f}")
            if eval_loss < best_eval_loss:
                best_eval_loss, early_stop_count = eval_loss, 0
                say("New best model found. Saving...")
                if rank == 0:
                    os.makedirs(CKPT_DIR, exist_ok=True)
                    torch.save(model.state_dict(), os.path.join(CKPT_DIR, "best_model_audio.pth"))
            else:
                early_stop_count += 1
            if early_stop_count >= patience:
                say("Early stopping triggered. Training stopped.")
                break
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    return best_eval_loss


if __name__ == "__main__":
    main()
