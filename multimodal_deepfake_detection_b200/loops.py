"""Host-side step/epoch mechanics shared by the entry-point scripts (train_visual.py, test_visual.py, train_audio.py,
train_au_face.py, test_au_face.py, train_au_patch.py at the repo root), i.e. what surrounds the hot path in the
reference's loops: batch movement, the fused loss call, the fused clip+Adam step, metric accumulation
(train_visual.py:476-487,563-590; train_au_face.py:462-507,633-705; train_audio.py:33-46).

Differences from the reference loops that follow from running on the sm_100a path (documented in INTEGRATION.md):
  * no autocast / GradScaler: the kernels compute in bf16 with fp32 accumulation natively, gradients are fp32 and
    unscaled, so ``clip_grad_norm_(…, 1.0)`` acts on the true gradients (fused into FusedAdam);
  * the per-step ``loss.item()`` is kept (the scripts print running losses) but metrics are accumulated on the device
    and read back once per epoch.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import XcpError


# ------------------------------------------------------------------------------------------------ metrics (numpy)
def _roc(labels: np.ndarray, scores: np.ndarray):
    """ROC points with sklearn.metrics.roc_curve's conventions (descending thresholds, a leading (0,0) point)."""
    order = np.argsort(-scores, kind="mergesort")
    y, s = labels[order].astype(np.float64), scores[order]
    distinct = np.where(np.diff(s))[0]
    idx = np.r_[distinct, y.size - 1]
    tps = np.cumsum(y)[idx]
    fps = 1 + idx - tps
    thr = s[idx]
    if tps.size > 2:                                   # roc_curve(drop_intermediate=True): drop collinear points
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        tps, fps, thr = tps[keep], fps[keep], thr[keep]
    tps, fps = np.r_[0.0, tps], np.r_[0.0, fps]
    thr = np.r_[np.inf, thr]
    P, N = max(tps[-1], 1e-12), max(fps[-1], 1e-12)
    return fps / N, tps / P, thr


def binary_metrics(labels: Sequence, probs: Sequence) -> Dict[str, float]:
    """AUC, pAUC (FPR <= 0.1, normalised), AP, EER and the EER threshold -- the quantities the reference prints every
    epoch (train_visual.py:476-487; train_au_face.py:462-473).  Degenerate single-class input returns the reference's
    sentinel values."""
    y = np.asarray(labels).astype(np.int64).ravel()
    p = np.asarray(probs, dtype=np.float64).ravel()
    if y.size == 0 or np.unique(y).size <= 1:
        return {"AUC": 0.0, "pAUC": 0.0, "AP": 0.0, "EER": 1.0, "THR": 0.5}
    fpr, tpr, thr = _roc(y, p)
    auc = float(np.trapezoid(tpr, fpr)) if hasattr(np, "trapezoid") else float(np.trapz(tpr, fpr))
    m = fpr <= 0.1
    pauc = 0.0
    if m.sum() >= 2:
        pauc = float((np.trapezoid if hasattr(np, "trapezoid") else np.trapz)(tpr[m], fpr[m]) / 0.1)
    fnr = 1.0 - tpr
    k = int(np.nanargmin(np.abs(fpr - fnr)))
    eer = float((fpr[k] + fnr[k]) / 2.0)
    # average precision = sum over distinct thresholds of (recall step) x precision (sklearn.metrics.average_precision_score)
    order = np.argsort(-p, kind="mergesort")
    ys, ss = y[order].astype(np.float64), p[order]
    idx = np.r_[np.where(np.diff(ss))[0], ys.size - 1]
    tp_c = np.cumsum(ys)[idx]
    prec = tp_c / (1.0 + idx)
    ap = float(np.sum(np.diff(np.r_[0.0, tp_c / max(tp_c[-1], 1e-12)]) * prec))
    return {"AUC": auc, "pAUC": pauc, "AP": ap, "EER": eer, "THR": float(thr[k])}


def youden_threshold(labels: Sequence, probs: Sequence) -> Tuple[float, float, float]:
    """Threshold maximising TPR - FPR (train_au_face.py:475-490 ``pick_threshold(mode="youden")``) -> (thr, fpr, tpr)."""
    y = np.asarray(labels).astype(np.int64).ravel()
    p = np.asarray(probs, dtype=np.float64).ravel()
    if y.size == 0 or np.unique(y).size <= 1:
        return 0.5, 0.0, 0.0
    fpr, tpr, thr = _roc(y, p)
    k = int(np.argmax(tpr - fpr))
    t = float(thr[k]) if np.isfinite(thr[k]) else 1.0
    return t, float(fpr[k]), float(tpr[k])


class ClassCounter:
    """Per-class hit counts accumulated on the device (the reference does four .cpu() round trips per step)."""

    def __init__(self, device):
        self.t = torch.zeros(4, device=device, dtype=torch.long)     # correct_real, total_real, correct_fake, total_fake

    def update(self, probs: torch.Tensor, labels: torch.Tensor, thr: float = 0.5):
        pred = (probs > thr).long().view(-1)
        lab = labels.long().view(-1)
        self.t += torch.stack([((pred == 0) & (lab == 0)).sum(), (lab == 0).sum(), ((pred == 1) & (lab == 1)).sum(), (lab == 1).sum()])

    def result(self):
        cr, tr, cf, tf = (int(v) for v in self.t.tolist())
        return cr, tr, cf, tf, (cr + cf) / (tr + tf + 1e-6)


# ------------------------------------------------------------------------------------------------ shared plumbing
def require_b200() -> torch.device:
    if not torch.cuda.is_available():
        raise XcpError("this entry point runs the sm_100a path only: no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def init_data_parallel():
    """One process per GPU under torchrun (WORLD_SIZE / RANK / LOCAL_RANK): select the GPU, join the NCCL group.
    Returns (world, rank).  A plain `python train_visual.py` run is world 1 and touches nothing."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 1, 0
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, int(os.environ.get("RANK", "0"))


def broadcast_module_state(modules, src: int = 0, buffers_only: bool = False):
    """Replica synchronisation: parameters + buffers at start-up, BatchNorm running statistics (which stay per-rank during
    training, like the reference's plain BatchNorm2d under DDP) before every evaluation so all ranks take the same
    scheduler / early-stopping decisions."""
    import torch.distributed as dist
    for m in modules:
        tensors = list(m.buffers()) if buffers_only else list(m.parameters()) + list(m.buffers())
        for t in tensors:
            dist.broadcast(t.data, src)


def env_int(name: str, default: int) -> int:
    """Bounded-run knobs for smoke tests on synthetic data (XCP_EPOCHS, XCP_SYNTH_CLIPS, XCP_FRAME_SIZE, ...)."""
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def set_backbone_trainable(model, trainable: bool):
    """train_visual.py:551-556: requires_grad toggling of the frozen/unfrozen backbone, per epoch."""
    for p in model.feature_extractor.parameters():
        p.requires_grad = trainable


def strip_module_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """nn.DataParallel / AveragedModel checkpoints carry a ``module.`` prefix (train_audio.py:87; train_au_face.py:735)."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items() if k != "n_averaged"}


# ------------------------------------------------------------------------------------------------ visual (ArcFace-CE)
def visual_batch(model, arcface, video, labels, seq_lengths, train: bool):
    """One batch of train_visual.py:563-572 / test_visual.py:614-619 -> (loss or None, P(fake) per clip)."""
    feats = model.extract_features(video, seq_lengths)
    emb = model.last_step(model.lstm(feats)[0], seq_lengths)        # [:, -1, :] unless model.use_seq_lengths (row f-2)
    if labels is None:
        logits = arcface(emb)
        return None, torch.softmax(logits, dim=1)[:, 1]
    lab = labels.long().view(-1)
    logits, loss = arcface.loss(emb, lab)                 # margin logits + CrossEntropy + both gradients in one kernel
    return loss, torch.softmax(logits.detach(), dim=1)[:, 1]


def visual_epoch(model, arcface, loader, device, optimizer=None, after_backward=None):
    """A full pass; optimizer=None evaluates.  Returns (mean loss, metrics dict, ClassCounter result).
    `after_backward` (data parallel: ddp.GradBucketer.finish) runs between loss.backward() and optimizer.step()."""
    train = optimizer is not None
    total = torch.zeros((), device=device)
    counter = ClassCounter(device)
    probs_all, labels_all = [], []
    n = 0
    with torch.set_grad_enabled(train):
        for video, labels, seq_lengths in loader:
            video = video.to(device, non_blocking=True)
            labels = labels.to(device, non_blocking=True)
            seq_lengths = seq_lengths.to(device, non_blocking=True)
            if train:
                optimizer.zero_grad(set_to_none=True)
            loss, probs = visual_batch(model, arcface, video, labels, seq_lengths, train)
            if train:
                loss.backward()
                if after_backward is not None:
                    after_backward()
                optimizer.step()                         # FusedAdam(max_norm=1.0): clip_grad_norm_ + Adam in one launch
            total += loss.detach()
            counter.update(probs, labels)
            probs_all.append(probs.detach().float())
            labels_all.append(labels.detach().float())
            n += 1
    p = torch.cat(probs_all).cpu().numpy() if probs_all else np.zeros(0)
    y = torch.cat(labels_all).cpu().numpy() if labels_all else np.zeros(0)
    return float(total) / max(n, 1), binary_metrics(y, p), counter.result()


# ------------------------------------------------------------------------------------------------ audio (BCE on sigmoid)
def audio_epoch(model, loader, device, optimizer=None, frontend=None, frames: int = 120, after_backward=None):
    """train_audio.py:33-46 / 55-67: BCELoss on the sigmoid output; returns (mean loss, accuracy).
    With `frontend` (audio_frontend.MFCC) a 2-D batch is taken as raw waveforms (B, samples) and turned into the
    (B, frames, 3, 13) MFCC tensor on the device (row f-4) instead of coming from the offline librosa files."""
    train = optimizer is not None
    total = torch.zeros((), device=device)
    hits = torch.zeros((), device=device)
    seen = 0
    n = 0
    with torch.set_grad_enabled(train):
        for audio, labels in loader:
            audio, labels = audio.to(device, non_blocking=True), labels.to(device, non_blocking=True)
            if audio.dim() == 2:
                if frontend is None:
                    raise XcpError("audio_epoch: got raw waveforms %s but no MFCC front-end" % (tuple(audio.shape),))
                audio = frontend.clips(audio, frames=frames)
            feats = model.extract_features(audio, device)
            # head + nn.BCELoss in one launch: same values and gradients as criterion(model(feats), labels)
            loss, out = model.forward_loss(feats, labels.float().view(-1, 1))
            if train:
                optimizer.zero_grad(set_to_none=True)
                loss.backward()
                if after_backward is not None:           # data parallel: ddp.GradBucketer.finish
                    after_backward()
                optimizer.step()
            total += loss.detach()
            hits += ((out.detach() > 0.5).float() == labels).sum()
            seen += labels.shape[0]
            n += 1
    return float(total) / max(n, 1), float(hits) / max(seen, 1)


# ------------------------------------------------------------------------------------------------ fusion (train_au_face)
def unpack_joint(batch):
    """3-tuple (videos, au, labels) or 5-tuple (+ au_mask, au_weight) batches (train_au_face.py:509-518)."""
    if len(batch) == 5:
        return batch
    v, a, y = batch
    return v, a, y, None, None


def fusion_forward(model, head, batch, device, train: bool):
    """train_au_face.py:643-674: two-stream tokens -> fused head/loss.  Returns (loss or None, P(fake), labels)."""
    videos, au, labels, au_mask, au_weight = unpack_joint(batch)
    if videos.dim() == 5 and videos.size(1) != 3 and videos.size(2) == 3:       # (B,T,C,H,W) -> (B,C,T,H,W)
        videos = videos.permute(0, 2, 1, 3, 4).contiguous()
    videos = videos.to(device, non_blocking=True)
    au = au.to(device, non_blocking=True)
    labels = labels.long().to(device, non_blocking=True)
    _, v_tokens, au_tokens = model(videos, au, au_mask=au_mask, au_weight=au_weight)
    if train:
        loss, logits = head(v_tokens, au_tokens, labels)
    else:
        loss, logits = None, head.predict_logits(v_tokens, au_tokens)
    return loss, torch.softmax(logits.detach(), dim=1)[:, 1], labels


def collect_scores(batches: Iterable, fn) -> Tuple[np.ndarray, np.ndarray]:
    ps, ys = [], []
    for b in batches:
        p, y = fn(b)
        ps.append(p.detach().float()); ys.append(y.detach().float())
    if not ps:
        return np.zeros(0), np.zeros(0)
    return torch.cat(ps).cpu().numpy(), torch.cat(ys).cpu().numpy()
