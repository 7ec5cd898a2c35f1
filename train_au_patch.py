"""train_au_patch.py -- patch-sequence training with label-smoothed BCE-with-logits (the reference's module-level
script train_au_patch.py:160-330 wrapped in `main()`), on the sm_100a path.

The reference's `Models.ResNetLSTM.AUPatchResNetClassifierWithAUAttention` is not part of the repository (SURVEY
App. C) and is out of scope; the in-scope configuration of this script (BASELINE.json configs[3]) runs the
XceptionLSTMA path on log-mel / MFCC patch sequences (B, T, 3, n) with the script's own criterion and protocol:
LabelSmoothingBCEWithLogitsLoss(0.1) on logits, Adam(lr=1e-4, weight_decay=1e-4), ReduceLROnPlateau(min, 0.5,
patience 4), batch 2, early stopping after 5 stale epochs, best checkpoint by eval loss."""
import os

import torch

from Dataset.audio_dataloader import collate_fn
from Dataset.synthetic import SyntheticAudio, synthetic_loader
from Models.XceptionLSTMA import XceptionLSTMA
from multimodal_deepfake_detection_b200 import FusedAdam, LabelSmoothingBCEWithLogitsLoss
from multimodal_deepfake_detection_b200.loops import binary_metrics, env_int, require_b200

CKPT_DIR = os.environ.get("XCP_CKPT_DIR", "Checkpoints")
CKPT_NAME = "au_patch_xception_lstma_best.pth"


def _epoch(model, criterion, loader, device, optimizer=None):
    train = optimizer is not None
    total = torch.zeros((), device=device)
    ps, ys, n = [], [], 0
    with torch.set_grad_enabled(train):
        for patches, labels in loader:
            patches, labels = patches.to(device, non_blocking=True), labels.to(device, non_blocking=True).float().view(-1, 1)
            # head + criterion (train_au_patch.py:203-214) in one launch: same loss / gradients as criterion(forward_logits(.), labels)
            loss, logits = model.forward_loss(model.extract_features(patches, device), labels, smoothing=criterion.smoothing)
            if train:
                optimizer.zero_grad(set_to_none=True)
                loss.backward()
                optimizer.step()
            total += loss.detach()
            ps.append(torch.sigmoid(logits.detach()).view(-1)); ys.append(labels.view(-1))
            n += 1
    m = binary_metrics(torch.cat(ys).cpu().numpy(), torch.cat(ps).cpu().numpy())
    return float(total) / max(n, 1), m


def main():
    device = require_b200()
    from Dataset.synthetic import dataset_missing, synthetic_requested
    if not synthetic_requested():      # the reference's AU-patch loader is absent: synthetic patch sequences only
        raise dataset_missing("AUPatchFeatureLoader", None)
    n, steps, n_mels = env_int("XCP_SYNTH_CLIPS", 16), env_int("XCP_PATCH_STEPS", 120), env_int("XCP_N_MELS", 64)
    train_loader = synthetic_loader(SyntheticAudio(n, steps, n_mels, seed=0), 2, True, collate_fn)
    eval_loader = synthetic_loader(SyntheticAudio(max(n // 2, 2), steps, n_mels, seed=1), 2, False, collate_fn)
    model = XceptionLSTMA(hidden_dim=env_int("XCP_AUDIO_HIDDEN", 128)).to(device)
    for p in model.feature_extractor.parameters():
        p.requires_grad = True
    criterion = LabelSmoothingBCEWithLogitsLoss()
    optimizer = FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.5, patience=4)
    best, stale, patience = float("inf"), 0, 5
    for epoch in range(env_int("XCP_EPOCHS", 50)):
        model.train()
        loss, m = _epoch(model, criterion, train_loader, device, optimizer)
        print(f"Epoch {epoch + 1}: Train Loss={loss:.4f}, AUC={m['AUC']:.4f}, EER={m['EER']:.4f}")
        model.eval()
        eval_loss, m = _epoch(model, criterion, eval_loader, device, None)
        print(f"Eval: Loss={eval_loss:.4f}, AUC={m['AUC']:.4f}, pAUC={m['pAUC']:.4f}, EER={m['EER']:.4f}")
        scheduler.step(eval_loss)
        if eval_loss < best:
            best, stale = eval_loss, 0
            os.makedirs(CKPT_DIR, exist_ok=True)
            torch.save(model.state_dict(), os.path.join(CKPT_DIR, CKPT_NAME))
            print("New best model saved.")
        else:
            stale += 1
            print(f"Early stopping patience: {stale}/{patience}")
            if stale >= patience:
                break
    return best


if __name__ == "__main__":
    main()
