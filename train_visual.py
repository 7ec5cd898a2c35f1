"""train_visual.py -- ArcFace/CE training of XceptionLSTMV on face clips (entry point `main()` as in the reference,
train_visual.py:489-649), running on the sm_100a path.

Same hyper-parameters and protocol as the reference: XceptionLSTMV(128) + ArcFaceHead(128, 2, s=30, m=0.5),
CrossEntropy on the margin logits, Adam(lr=1e-5, weight_decay=1e-4) over model + head, clip_grad_norm_ 1.0,
ReduceLROnPlateau(min, 0.5, patience 3), backbone frozen for the first 3 epochs, batch 4, best checkpoint
``{"model", "arcface"}`` on (eval loss, EER), early stopping after 6 stale epochs.  Without the LAV-DF tree the loaders
serve synthetic clips (Dataset/); XCP_EPOCHS / XCP_SYNTH_CLIPS / XCP_FRAME_SIZE bound a smoke run.
"""
import os
from collections import Counter

import torch
import torch.multiprocessing as mp
from torch.utils.data import DataLoader

from Dataset.video_dataloader_enhanced import collate_fn, get_face_dataloader
from Models.XceptionLSTMV import XceptionLSTMV
from multimodal_deepfake_detection_b200 import ArcFaceHead, FusedAdam
from multimodal_deepfake_detection_b200.loops import env_int, require_b200, set_backbone_trainable, visual_epoch

CKPT_DIR = os.environ.get("XCP_CKPT_DIR", "Checkpoints")
CKPT_NAME = "XceptionLSTMV_ArcFace_Best.pth"


def _dataset(subset, folder, lavdf_json):
    size = env_int("XCP_FRAME_SIZE", 224)
    return get_face_dataloader(folder_path=folder, mode="lavdf_raw", subset=subset, lavdf_json=lavdf_json, batch_size=1,
                               augment_minority=False, shuffle=False, raw_video=True, use_face_detection=True,
                               frame_size=(size, size), max_frames=50, synthetic_clips=env_int("XCP_SYNTH_CLIPS", 32)).dataset


def main():
    device = require_b200()
    train_folder = eval_folder = os.environ.get("XCP_LAVDF_ROOT", "/media/rt0706/Lab/LAV-DF")
    lavdf_json = "Dataset/LAV-DF/metadata.json"

    print("Loading training data...")
    train_dataset = _dataset("train", train_folder, lavdf_json)
    print("Loading eval data...")
    eval_dataset = _dataset("dev", eval_folder, lavdf_json)
    print("Class counts:", Counter(lbl for _, lbl, _ in train_dataset.samples))

    model = XceptionLSTMV(hidden_dim=128).to(device)
    arcface_head = ArcFaceHead(128, 2, s=30.0, m=0.5).to(device)
    params = list(model.parameters()) + list(arcface_head.parameters())
    optimizer = FusedAdam(params, lr=1e-5, weight_decay=1e-4, max_norm=1.0)       # Adam + clip_grad_norm_(…, 1.0), one launch
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.5, patience=3)

    best_eval_loss = best_eer = float("inf")
    patience, early_stop_count = 6, 0
    num_epochs, freeze_epochs = env_int("XCP_EPOCHS", 50), env_int("XCP_FREEZE_EPOCHS", 3)
    workers = env_int("XCP_WORKERS", 2)
    train_loader = DataLoader(train_dataset, batch_size=4, shuffle=True, num_workers=workers, collate_fn=collate_fn, pin_memory=True)
    eval_loader = DataLoader(eval_dataset, batch_size=4, shuffle=False, num_workers=workers, collate_fn=collate_fn, pin_memory=True)

    for epoch in range(num_epochs):
        print(f"\nEpoch {epoch + 1}/{num_epochs}")
        set_backbone_trainable(model, epoch >= freeze_epochs)
        model.train(); arcface_head.train()
        loss, m, (cr, tr, cf, tf, acc) = visual_epoch(model, arcface_head, train_loader, device, optimizer)
        print(f"Train: Loss={loss:.4f}, Acc={acc:.4f}, AUC={m['AUC']:.4f}, pAUC={m['pAUC']:.4f}, AP={m['AP']:.4f}, EER={m['EER']:.4f}")
        print(f"Train Correct Real: {cr}/{tr} | Correct Fake: {cf}/{tf}")

        model.eval(); arcface_head.eval()
        eval_loss, m, (cr, tr, cf, tf, acc) = visual_epoch(model, arcface_head, eval_loader, device, None)
        print(f"Eval: Loss={eval_loss:.4f}, Acc={acc:.4f}, AUC={m['AUC']:.4f}, pAUC={m['pAUC']:.4f}, AP={m['AP']:.4f}, EER={m['EER']:.4f}")
        print(f"Eval Correct Real: {cr}/{tr} | Correct Fake: {cf}/{tf}")
        scheduler.step(eval_loss)

        if eval_loss < best_eval_loss and m["EER"] < best_eer:
            best_eval_loss, best_eer, early_stop_count = eval_loss, m["EER"], 0
            os.makedirs(CKPT_DIR, exist_ok=True)
            torch.save({"model": model.state_dict(), "arcface": arcface_head.state_dict()}, os.path.join(CKPT_DIR, CKPT_NAME))
            print("New best model saved.")
        else:
            early_stop_count += 1
            print(f"Early stopping counter: {early_stop_count}/{patience}")
            if early_stop_count >= patience:
                print("Early stopping triggered.")
                break
    print("Training finished.")
    return best_eval_loss


if __name__ == "__main__":
    mp.set_start_method("spawn", force=True)
    main()
