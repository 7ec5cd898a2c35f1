"""`Dataset.AuVidDataset.get_joint_dataloader` as imported by train_au_face.py:410 (absent from the reference).
Serves paired synthetic (videos[B,3,T,H,W], audio[B,Ta,3,13], labels[B]) batches."""
import torch
from torch.utils.data import DataLoader, Dataset

from .synthetic import SyntheticAudio, SyntheticClips, dataset_missing, synthetic_requested


class JointSynthetic(Dataset):
    def __init__(self, n, frames, size, seed):
        self.v = SyntheticClips(n, frames, size, seed)
        self.a = SyntheticAudio(n, steps=frames, seed=seed + 1)
        self.all_labels = self.v.labels

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        vid, lab = self.v[i]
        aud, _ = self.a[i]
        return vid.permute(1, 0, 2, 3).contiguous(), aud, lab.long()          # (3,T,H,W): train_au_face.py:643-644 layout


def _collate(batch):
    v, a, y = zip(*batch)
    return torch.stack(v), torch.stack(a), torch.stack(y)


def get_joint_dataloader(video_root=None, au_root=None, batch_size=2, shuffle=True, max_frames=16, max_aus=17, image_size=128,
                         num_workers=0, csv_path=None, return_weights=False, n_train=16, **_):
    if not synthetic_requested():          # the reference's joint dataset class is absent: synthetic pairs only
        raise dataset_missing("get_joint_dataloader", video_root)
    mk = lambda n, seed, sh: DataLoader(JointSynthetic(n, min(max_frames, 16), image_size, seed), batch_size=batch_size,
                                        shuffle=sh, collate_fn=_collate)
    return mk(n_train, 0, shuffle), mk(max(n_train // 2, batch_size), 1, False), mk(max(n_train // 2, batch_size), 2, False)
