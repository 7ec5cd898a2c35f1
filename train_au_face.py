"""train_au_face.py -- audio-face fusion training (entry point `main()` as in the reference, train_au_face.py:558-768) on
the sm_100a path: AUFaceCrossDetector tokens -> embed head -> ArcFace(m=0.30) -> CB-Focal + 0.2*align MSE +
0.1*temporal smoothness, AdamW(1e-4, wd 0.01) under OneCycleLR(max 1e-3, pct_start 0.3), 4-step gradient
accumulation, clip 1.0, EMA weights for evaluation, best-AUC checkpoint {"model","embed","arcface","best_auc"}.
The head / loss region is one fused module (modules.FusionHead).  Synthetic paired loaders stand in for the dataset."""
import math
import os
from collections import Counter

import torch
from torch.optim.swa_utils import AveragedModel

from Dataset.AuVidDataset import get_joint_dataloader
from Models.AUFaceModel import AUFaceCrossDetector, FusionHead
from multimodal_deepfake_detection_b200 import FusedAdam
from multimodal_deepfake_detection_b200.loops import (binary_metrics, collect_scores, env_int, fusion_forward, require_b200,
                                                      youden_threshold)

accum_steps, patience, grad_clip, seed = 4, 8, 1.0, 42
lambda_align, lambda_temp = 0.2, 0.1
CKPT_DIR = os.environ.get("XCP_CKPT_DIR", "Checkpoints")
CKPT_NAME = "auface_cross_best_auc_arcface_cb.pth"


def main():
    device = require_b200()
    torch.manual_seed(seed)
    os.makedirs(CKPT_DIR, exist_ok=True)
    epochs = env_int("XCP_EPOCHS", 100)
    hidden = env_int("XCP_FUSION_HIDDEN", 256)
    train_loader, test_loader, eval_loader = get_joint_dataloader(
        video_root="/media/rt0706/Media/VCBSL-Dataset/FAVC_Whole/frames", au_root="Dataset/AU_Files/fakeavceleb_whole_image_patches",
        batch_size=2, shuffle=True, max_frames=env_int("XCP_MAX_FRAMES", 75), max_aus=17, image_size=env_int("XCP_FRAME_SIZE", 128),
        num_workers=0, csv_path="Dataset/meta_data.csv", return_weights=True, n_train=env_int("XCP_SYNTH_CLIPS", 16))

    model = AUFaceCrossDetector(num_aus=17, face_dim=512, au_dim=512, lstm_hidden=hidden).to(device)
    for p in model.parameters():                         # the fusion script trains both streams end to end
        p.requires_grad = True
    counts = Counter(getattr(train_loader.dataset, "all_labels", []))
    samples_per_cls = [max(counts.get(0, 1), 1), max(counts.get(1, 1), 1)]
    print(f"[Info] Class counts (for CB-Focal): real={samples_per_cls[0]}, fake={samples_per_cls[1]}")
    head = FusionHead(hidden, samples_per_cls, s=30.0, m=0.30, beta=0.9999, gamma=2.0, lambda_align=lambda_align,
                      lambda_temp=lambda_temp, p_drop=0.2).to(device)
    ema_model, ema_head = AveragedModel(model), AveragedModel(head)

    params = list(model.parameters()) + list(head.parameters())
    optimizer = FusedAdam(params, lr=1e-4, weight_decay=0.01, decoupled=True, max_norm=grad_clip)       # AdamW + clip, one launch
    steps_per_epoch = math.ceil(len(train_loader) / max(1, accum_steps))
    scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr=1e-3, epochs=epochs, steps_per_epoch=steps_per_epoch, pct_start=0.3)

    best_auc, early_stop_count = 0.0, 0
    for epoch in range(epochs):
        model.train(); head.train()
        print(f"\nEpoch {epoch + 1}")
        optimizer.zero_grad(set_to_none=True)
        running = torch.zeros((), device=device)
        probs_all, labels_all = [], []
        for i, batch in enumerate(train_loader):
            loss, probs, labels = fusion_forward(model, head, batch, device, True)
            loss.backward()                              # gradients accumulate across the accum_steps micro-batches
            if (i + 1) % accum_steps == 0 or (i + 1) == len(train_loader):
                optimizer.step()
                optimizer.zero_grad(set_to_none=True)
                scheduler.step()
                ema_model.update_parameters(model)
                ema_head.update_parameters(head)
            running += loss.detach()
            probs_all.append(probs.float()); labels_all.append(labels.float())
        y, p = torch.cat(labels_all).cpu().numpy(), torch.cat(probs_all).cpu().numpy()
        m = binary_metrics(y, p)
        print(f"Train: Loss={float(running) / max(1, len(train_loader)):.4f}, AUC={m['AUC']:.4f}, pAUC={m['pAUC']:.4f}, EER={m['EER']:.4f}")

        ema_model.eval(); ema_head.eval()
        with torch.no_grad():
            ep, ey = collect_scores(eval_loader, lambda b: fusion_forward(ema_model.module, ema_head.module, b, device, False)[1:])
        m = binary_metrics(ey, ep)
        thr, fpr, tpr = youden_threshold(ey, ep)
        print(f"Eval: AUC={m['AUC']:.4f}, pAUC={m['pAUC']:.4f}, EER={m['EER']:.4f}, AP={m['AP']:.4f}, thr={thr:.3f}, FPR={fpr:.3f}, TPR={tpr:.3f}")
        if m["AUC"] > best_auc:                          # strict, like train_au_face.py:748
            best_auc, early_stop_count = m["AUC"], 0
            # the ArcFace weights that scored this AUC (the evaluation above ran on the EMA head): test_au_face.py must score
            # with the same ones that picked the checkpoint
            torch.save({"model": ema_model.state_dict(), "embed": ema_head.module.embed_head.state_dict(),
                        "arcface": ema_head.module.arcface.state_dict(), "best_auc": best_auc}, os.path.join(CKPT_DIR, CKPT_NAME))
            print(f"New best AUC: {m['AUC']:.4f} - Model saved.")
        else:
            early_stop_count += 1
            if early_stop_count >= patience:
                print(f"Early stopping at AUC {best_auc:.4f}")
                break
    print("Training Complete.")
    return best_auc


if __name__ == "__main__":
    main()
