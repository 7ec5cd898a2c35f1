"""test_au_face.py -- evaluation of the fusion checkpoint (entry point `main()` as in the reference,
test_au_face.py:228-345) on the sm_100a path: loads {"model","embed","arcface"}, scores the eval (fallback: test) split,
prints AUC / pAUC / EER / AP and the Youden operating point, saves scores + labels.  (The reference's t-SNE plots are
outside the hot path and not reproduced.)"""
import os

import numpy as np
import torch

from Dataset.AuVidDataset import get_joint_dataloader
from Models.AUFaceModel import AUFaceCrossDetector, FusionHead
from multimodal_deepfake_detection_b200.loops import (binary_metrics, collect_scores, env_int, fusion_forward, require_b200,
                                                      strip_module_prefix, youden_threshold)

WEIGHTS_PATH = os.path.join(os.environ.get("XCP_CKPT_DIR", "Checkpoints"), "auface_cross_best_auc_arcface_cb.pth")
OUTPUT_DIR = os.environ.get("XCP_OUTPUT_DIR", "eval_outputs")
PRIMARY_SPLIT, SEED, MAX_AUS = "eval", 42, 17


def main():
    device = require_b200()
    os.makedirs(OUTPUT_DIR, exist_ok=True)
    torch.manual_seed(SEED); np.random.seed(SEED)
    hidden = env_int("XCP_FUSION_HIDDEN", 256)
    _, test_loader, eval_loader = get_joint_dataloader(video_root=None, au_root=None, batch_size=2, shuffle=False,
                                                       max_frames=env_int("XCP_MAX_FRAMES", 75), max_aus=MAX_AUS,
                                                       image_size=env_int("XCP_FRAME_SIZE", 128), num_workers=0,
                                                       csv_path="Dataset/meta_data.csv", n_train=env_int("XCP_SYNTH_CLIPS", 16))
    loader, split_used = (eval_loader, "eval") if PRIMARY_SPLIT == "eval" else (test_loader, "test")
    if loader is None or len(loader.dataset) == 0:
        loader, split_used = (test_loader, "test") if split_used == "eval" else (eval_loader, "eval")
        if loader is None or len(loader.dataset) == 0:
            raise RuntimeError("No valid data in either eval or test splits.")
    print(f"[Data] Evaluating split: {split_used}  |  N={len(loader.dataset)}")

    model = AUFaceCrossDetector(num_aus=MAX_AUS, face_dim=512, au_dim=512, lstm_hidden=hidden).to(device)
    head = FusionHead(hidden).to(device)
    assert os.path.isfile(WEIGHTS_PATH), f"Missing weights: {WEIGHTS_PATH}"
    ckpt = torch.load(WEIGHTS_PATH, map_location=device)
    model.load_state_dict(strip_module_prefix(ckpt["model"]))
    head.embed_head.load_state_dict(strip_module_prefix(ckpt["embed"]))
    head.arcface.load_state_dict(ckpt["arcface"])
    model.eval(); head.eval()
    with torch.no_grad():
        scores, labels = collect_scores(loader, lambda b: fusion_forward(model, head, b, device, False)[1:])
    m = binary_metrics(labels, scores)
    thr, fpr, tpr = youden_threshold(labels, scores)
    print(f"[{split_used.upper()}] AUC={m['AUC']:.4f} | pAUC@0.1={m['pAUC']:.4f} | EER={m['EER']:.4f} | AP={m['AP']:.4f}")
    print(f"[{split_used.upper()}] thr={thr:.3f}, FPR={fpr:.3f}, TPR={tpr:.3f}")
    np.savez(os.path.join(OUTPUT_DIR, f"{split_used}_scores_and_labels.npz"), scores=scores, labels=labels)
    print("Done.")
    return m


if __name__ == "__main__":
    main()
