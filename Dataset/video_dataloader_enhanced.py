"""`Dataset.video_dataloader_enhanced` as imported by train_visual.py:451 / test_visual.py:466 (absent from the reference;
signature from the call sites).  Serves a folder of pre-processed `.npy` clips; synthetic clips only on
explicit request (XCP_SYNTHETIC=1); raw LAV-DF / FakeAVCeleb trees (mp4 + metadata) are not decoded here and raise."""
import os
import zlib

from torch.utils.data import DataLoader

from .synthetic import (SyntheticClips, collate_clips_with_lengths as collate_fn, dataset_missing, label_from_name,  # noqa: F401
                        synthetic_requested)
from .video_dataloader import FaceDataset


def get_face_dataloader(folder_path=None, mode="lavdf_raw", subset="train", lavdf_json=None, csv_path=None, batch_size=1,
                        augment_minority=False, shuffle=False, raw_video=True, use_face_detection=True, frame_size=(224, 224),
                        max_frames=50, synthetic_clips=32):
    if folder_path and os.path.isdir(folder_path) and any(f.endswith(".npy") for f in os.listdir(folder_path)):
        ds = FaceDataset(folder_path)
        ds.samples = [(f, label_from_name(f), None) for f in ds.files]
    elif not synthetic_requested():
        if folder_path and os.path.isdir(folder_path):
            raise FileNotFoundError("get_face_dataloader: %r holds no `.npy` clips; raw %s trees (mp4 + metadata) are not decoded by this "
                                    "loader -- run the reference's pre-processor first, or set XCP_SYNTHETIC=1" % (folder_path, mode))
        raise dataset_missing("get_face_dataloader", folder_path)
    else:
        ds = SyntheticClips(n=synthetic_clips, frames=min(max_frames, 16), size=frame_size[0], seed=zlib.crc32(subset.encode()) % 1000)   # (str hash() is salted per process)
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, collate_fn=collate_fn)
