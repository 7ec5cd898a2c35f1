"""train_visual.py -- ArcFace/CE training of XceptionLSTMV on face clips (entry point `main()` as in the reference,
train_visual.py:489-649), running on the sm_100a path.

Same hyper-parameters and protocol as the reference: XceptionLSTMV(128) + ArcFaceHead(128, 2, s=30, m=0.5),
CrossEntropy on the margin logits, Adam(lr=1e-5, weight_decay=1e-4) over model + head, clip_grad_norm_ 1.0,
ReduceLROnPlateau(min, 0.5, patience 3), backbone frozen for the first 3 epochs, batch 4, best checkpoint
``{"model", "arcface"}`` on (eval loss, EER), early stopping after 6 stale epochs.  Without the LAV-DF tree the loaders
serve synthetic clips (Dataset/); XCP_EPOCHS / XCP_SYNTH_CLIPS / XCP_FRAME_SIZE bound a smoke run.

Multi-GPU (no counterpart in the reference): `torchrun --nproc-per-node N train_visual.py` runs one
process per GPU; clips are sharded by a DistributedSampler (batch 4 per GPU), gradients are averaged by the bucketed NCCL
all-reduce of ddp.GradBucketer overlapped with backward, BatchNorm statistics stay per rank during training and rank 0's
are broadcast before each evaluation; rank 0 prints and writes the checkpoint.
"""
import os
from collections import Counter

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch.utils.data import DataLoader
from torch.utils.data.distributed import DistributedSampler

from Dataset.video_dataloader_enhanced import collate_fn, get_face_dataloader
from Models.XceptionLSTMV import XceptionLSTMV
from multimodal_deepfake_detection_b200 import ArcFaceHead, FusedAdam
from multimodal_deepfake_detection_b200.ddp import GradBucketer
from multimodal_deepfake_detection_b200.loops import (broadcast_module_state, env_int, init_data_parallel, require_b200,
                                                     set_backbone_trainable, visual_epoch)

CKPT_DIR = os.environ.get("XCP_CKPT_DIR", "Checkpoints")
CKPT_NAME = "XceptionLSTMV_ArcFace_Best.pth"


def _dataset(subset, folder, lavdf_json):
    size = env_int("XCP_FRAME_SIZE", 224)
    return get_face_dataloader(folder_path=folder, mode="lavdf_raw", subset=subset, lavdf_json=lavdf_json, batch_size=1,
                               augment_minority=False, shuffle=False, raw_video=True, use_face_detection=True,
                               frame_size=(size, size), max_frames=50, synthetic_clips=env_int("XCP_SYNTH_CLIPS", 32)).dataset


def main():
    world, rank = init_data_parallel()
    device = require_b200()
    say = print if rank == 0 else (lambda *a, **k: None)
    train_folder = eval_folder = os.environ.get("XCP_LAVDF_ROOT", "/media/rt0706/Lab/LAV-DF")
    lavdf_json = "Dataset/LAV-DF/metadata.json"

    say("Loading training data...")
    train_dataset = _dataset("train", train_folder, lavdf_json)
    say("Loading eval data...")
    eval_dataset = _dataset("dev", eval_folder, lavdf_json)
    say("Class counts:", Counter(lbl for _, lbl, _ in train_dataset.samples))

    model = XceptionLSTMV(hidden_dim=128).to(device)
    arcface_head = ArcFaceHead(128, 2, s=30.0, m=0.5).to(device)
    params = list(model.parameters()) + list(arcface_head.parameters())
    optimizer = FusedAdam(params, lr=1e-5, weight_decay=1e-4, max_norm=1.0)       # Adam + clip_grad_norm_(…, 1.0), one launch
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.5, patience=3)

    best_eval_loss = best_eer = float("inf")
    patience, early_stop_count = 6, 0
    num_epochs, freeze_epochs = env_int("XCP_EPOCHS", 50), env_int("XCP_FREEZE_EPOCHS", 3)
    workers = env_int("XCP_WORKERS", 2)
    sampler, after_backward = None, None
    if world > 1:                                         # one replica per GPU: shard the clips, average the gradients
        sampler = DistributedSampler(train_dataset, num_replicas=world, rank=rank, shuffle=True, drop_last=True)
        broadcast_module_state([model, arcface_head])
        bucketer = GradBucketer(model, backbone=model.feature_extractor)
        backbone_ids = {id(p) for p in model.feature_extractor.parameters()}
        extra = [p for p in params if id(p) not in backbone_ids]               # LSTM + head + ArcFace: one final bucket
        after_backward = lambda: bucketer.finish(extra)  # noqa: E731
    train_loader = DataLoader(train_dataset, batch_size=4, shuffle=(sampler is None), sampler=sampler, num_workers=workers,
                              collate_fn=collate_fn, pin_memory=True)
    eval_loader = DataLoader(eval_dataset, batch_size=4, shuffle=False, num_workers=workers, collate_fn=collate_fn, pin_memory=True)

    for epoch in range(num_epochs):
        say(f"\nEpoch {epoch + 1}/{num_epochs}")
        if sampler is not None:
            sampler.set_epoch(epoch)
        set_backbone_trainable(model, epoch >= freeze_epochs)
        model.train(); arcface_head.train()
        loss, m, (cr, tr, cf, tf, acc) = visual_epoch(model, arcface_head, train_loader, device, optimizer, after_backward)
        say(f"Train: Loss={loss:.4f}, Acc={acc:.4f}, AUC={m['AUC']:.4f}, pAUC={m['pAUC']:.4f}, AP={m['AP']:.4f}, EER={m['EER']:.4f}")
        say(f"Train Correct Real: {cr}/{tr} | Correct Fake: {cf}/{tf}")

        model.eval(); arcface_head.eval()
        if world > 1:                                     # same running statistics everywhere => same decisions below
            broadcast_module_state([model], buffers_only=True)
        eval_loss, m, (cr, tr, cf, tf, acc) = visual_epoch(model, arcface_head, eval_loader, device, None)
        say(f"Eval: Loss={eval_loss:.4f}, Acc={acc:.4f}, AUC={m['AUC']:.4f}, pAUC={m['pAUC']:.4f}, AP={m['AP']:.4f}, EER={m['EER']:.4f}")
        say(f"Eval Correct Real: {cr}/{tr} | Correct Fake: {cf}/{tf}")
        if world > 1:                                     # belt and braces: rank 0's numbers drive every rank's control flow
            ctl = torch.tensor([eval_loss, m["EER"]], device=device, dtype=torch.float64)
            dist.broadcast(ctl, 0)
            eval_loss, m["EER"] = float(ctl[0]), float(ctl[1])
        scheduler.step(eval_loss)

        if eval_loss < best_eval_loss and m["EER"] < best_eer:
            best_eval_loss, best_eer, early_stop_count = eval_loss, m["EER"], 0
            if rank == 0:
                os.makedirs(CKPT_DIR, exist_ok=True)
                torch.save({"model": model.state_dict(), "arcface": arcface_head.state_dict()}, os.path.join(CKPT_DIR, CKPT_NAME))
            say("New best model saved.")
        else:
            early_stop_count += 1
            say(f"Early stopping counter: {early_stop_count}/{patience}")
            if early_stop_count >= patience:
                say("Early stopping triggered.")
                break
    say("Training finished.")
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    return best_eval_loss


if __name__ == "__main__":
    mp.set_start_method("spawn", force=True)
    main()
