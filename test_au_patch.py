"""test_au_patch.py -- evaluation counterpart of train_au_patch.py (reference: test_au_patch.py, module-level script):
loads the best checkpoint, scores the evaluation split and prints AUC / pAUC / AP / EER on the sm_100a path."""
import os

import torch

from Dataset.audio_dataloader import collate_fn
from Dataset.synthetic import SyntheticAudio, synthetic_loader
from Models.XceptionLSTMA import XceptionLSTMA
from multimodal_deepfake_detection_b200.loops import binary_metrics, env_int, require_b200, strip_module_prefix

CKPT_PATH = os.path.join(os.environ.get("XCP_CKPT_DIR", "Checkpoints"), "au_patch_xception_lstma_best.pth")


def main():
    device = require_b200()
    from Dataset.synthetic import dataset_missing, synthetic_requested
    if not synthetic_requested():      # the reference's AU-patch loader is absent: synthetic patch sequences only
        raise dataset_missing("AUPatchFeatureLoader", None)
    n, steps, n_mels = env_int("XCP_SYNTH_CLIPS", 16), env_int("XCP_PATCH_STEPS", 120), env_int("XCP_N_MELS", 64)
    loader = synthetic_loader(SyntheticAudio(max(n // 2, 2), steps, n_mels, seed=1), 2, False, collate_fn)
    model = XceptionLSTMA(hidden_dim=env_int("XCP_AUDIO_HIDDEN", 128)).to(device)
    model.load_state_dict(strip_module_prefix(torch.load(CKPT_PATH, map_location=device)))
    model.eval()
    ps, ys = [], []
    with torch.no_grad():
        for patches, labels in loader:
            probs = model(model.extract_features(patches.to(device), device))       # sigmoid(fc_out(.)), XceptionLSTMA.py:55-59
            ps.append(probs.view(-1)); ys.append(labels.view(-1).to(device))
    m = binary_metrics(torch.cat(ys).cpu().numpy(), torch.cat(ps).cpu().numpy())
    print("=== AU/log-mel patch test ===")
    for k, v in m.items():
        print(f"{k}: {v:.4f}")
    return m


if __name__ == "__main__":
    main()
