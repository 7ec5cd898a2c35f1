"""test_visual.py -- FakeAVCeleb-style evaluation of the XceptionLSTMV + ArcFace checkpoint (entry point `test()` as in
the reference, test_visual.py:572-650): inference logits s*cos (labels=None), softmax P(fake), accuracy / AUC / pAUC /
AP / EER and class-wise counts.  Runs on the sm_100a path; synthetic clips stand in for the dataset when it is absent."""
import os

import torch
import torch.multiprocessing as mp
from torch.utils.data import DataLoader

from Dataset.video_dataloader_enhanced import collate_fn, get_face_dataloader
from Models.XceptionLSTMV import XceptionLSTMV
from multimodal_deepfake_detection_b200 import ArcFaceHead
from multimodal_deepfake_detection_b200.loops import ClassCounter, binary_metrics, env_int, require_b200, visual_batch

CKPT_PATH = os.path.join(os.environ.get("XCP_CKPT_DIR", "Checkpoints"), "XceptionLSTMV_ArcFace_Best.pth")


def test():
    device = require_b200()
    size = env_int("XCP_FRAME_SIZE", 224)
    print("Loading FakeAVCeleb test data...")
    test_dataset = get_face_dataloader(folder_path=os.environ.get("XCP_FAVC_ROOT", "/media/rt0706/Media/VCBSL-Dataset/FAVC_Whole/frames"),
                                       mode="fakeavceleb", subset="test", csv_path="Dataset/meta_data.csv", batch_size=1,
                                       augment_minority=False, shuffle=False, raw_video=False, use_face_detection=False,
                                       frame_size=(size, size), max_frames=75,
                                       synthetic_clips=env_int("XCP_SYNTH_CLIPS", 32)).dataset
    test_loader = DataLoader(test_dataset, batch_size=4, shuffle=False, num_workers=env_int("XCP_WORKERS", 2), collate_fn=collate_fn)

    model = XceptionLSTMV(hidden_dim=128).to(device)
    arcface_head = ArcFaceHead(128, 2, s=30.0, m=0.5).to(device)
    ckpt = torch.load(CKPT_PATH, map_location=device)
    model.load_state_dict(ckpt["model"])
    arcface_head.load_state_dict(ckpt["arcface"])
    print(f"Loaded checkpoint from {CKPT_PATH}")
    model.eval(); arcface_head.eval()

    counter = ClassCounter(device)
    probs_all, labels_all = [], []
    with torch.no_grad():
        for video_batch, labels, seq_lengths in test_loader:
            video_batch, labels = video_batch.to(device, non_blocking=True), labels.to(device, non_blocking=True)
            _, probs = visual_batch(model, arcface_head, video_batch, None, seq_lengths.to(device), False)
            counter.update(probs, labels)
            probs_all.append(probs.float()); labels_all.append(labels.float())
    cr, tr, cf, tf, acc = counter.result()
    metrics = binary_metrics(torch.cat(labels_all).cpu().numpy(), torch.cat(probs_all).cpu().numpy())
    print("\n=== FakeAVCeleb Test Results ===")
    print(f"Accuracy: {acc:.4f}")
    for k, v in metrics.items():
        print(f"{k}: {v:.4f}")
    print(f"Classwise: Real {cr}/{tr}, Fake {cf}/{tf}")
    return dict(metrics, ACC=acc)


if __name__ == "__main__":
    mp.set_start_method("spawn", force=True)
    test()
