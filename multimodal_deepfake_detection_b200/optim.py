"""Fused optimizer step for the hot path.

The reference's loops run ``clip_grad_norm_`` + ``torch.optim.Adam(lr=1e-5, weight_decay=1e-4).step()``
(train_visual.py:533,574-577) or AdamW (train_au_face.py:616-619,678-693) over ~300 parameter tensors.  Here the
whole update -- global gradient norm, clip factor, L2 / decoupled weight decay, moments, bias correction, parameter
write -- is ONE multi-tensor launch of ``xcp_adam_multi`` driven by a device-side pointer table.  The step counter
lives in device memory, so the step is CUDA-graph capturable and never synchronises with the host.

``state_dict()`` keeps torch.optim.Adam's layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``) so
optimizer checkpoints interchange with the reference's.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib
from ._lib import XcpError
from .executor import bump_param_epoch

_CHUNK = 8192   # must match ADAM_CHUNK in csrc/optim.cu


class FusedAdam(torch.optim.Optimizer):
    """Adam / AdamW with torch.optim.Adam's arithmetic.  ``decoupled=True`` gives AdamW; ``max_norm`` fuses
    ``clip_grad_norm_(params, max_norm)`` into the step (the norm is taken over every gradient handed to the step)."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 decoupled: bool = False, max_norm: Optional[float] = None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled, max_norm=max_norm)
        super().__init__(params, defaults)
        self._tables = {}        # group index -> dict(key, table, chunks, n_chunks, keep)
        self._chunk_cache = {}
        self._keep = []          # tables a CUDA-graph capture has seen: referenced by raw pointer from the graph, never freed

    # ---- state -----------------------------------------------------------------------------------------------
    def _init_group_state(self, gi: int, group):
        ps = [p for p in group["params"]]
        if not ps:
            return
        dev = ps[0].device
        if dev.type != "cuda":
            raise XcpError("FusedAdam: parameters must live on a CUDA device (no CPU path)")
        total = sum((p.numel() + 3) // 4 * 4 for p in ps)
        m = torch.zeros((total,), device=dev, dtype=torch.float32)
        v = torch.zeros((total,), device=dev, dtype=torch.float32)
        steps = torch.zeros((len(ps),), device=dev, dtype=torch.int32)      # one counter per parameter, like torch.optim.Adam
        off = 0
        for pi, p in enumerate(ps):
            if p.dtype != torch.float32:
                raise XcpError("FusedAdam: fp32 master parameters expected, got %s" % p.dtype)
            n = p.numel()
            st = self.state[p]
            old_m, old_v, old_step = st.get("exp_avg"), st.get("exp_avg_sq"), st.get("step")
            st["exp_avg"] = m[off:off + n].view(p.shape)
            st["exp_avg_sq"] = v[off:off + n].view(p.shape)
            if old_m is not None:           # load_state_dict() before the first step
                st["exp_avg"].copy_(old_m); st["exp_avg_sq"].copy_(old_v)
            st["step"] = steps[pi]
            if old_step is not None:
                st["step"].fill_(int(old_step))
            off += (n + 3) // 4 * 4
        group["_xcp_state"] = (m, v, steps)
        self._tables.pop(gi, None)

    def _group_ready(self, gi: int, group) -> bool:
        s = group.get("_xcp_state")
        if s is None:
            return False
        for p in group["params"]:
            st = self.state.get(p)
            if not st or "exp_avg" not in st or st["exp_avg"].device != p.device or not torch.is_tensor(st.get("step")) \
                    or st["step"].dtype != torch.int32 or st["step"].device != p.device:
                return False
        return True

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for group in self.param_groups:     # re-pack the loaded per-parameter state into flat arenas on the next step
            group.pop("_xcp_state", None)
        self._tables.clear()

    # ---- step ------------------------------------------------------------------------------------------------
    def _table(self, gi: int, group, live):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for p in live)
        ent = self._tables.get(gi)
        if ent is not None and ent["key"] == key:
            self._pin_if_capturing(ent)
            return ent
        rows = np.empty((len(live), 6), dtype=np.uint64)
        for ti, p in enumerate(live):
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                raise XcpError("FusedAdam: contiguous fp32 parameters and gradients expected")
            st = self.state[p]
            rows[ti] = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), st["step"].data_ptr())
        sizes = tuple(int(n) for n in rows[:, 4])
        chunks = self._chunk_cache.get(sizes)
        if chunks is None:
            per = [(n + _CHUNK - 1) // _CHUNK for n in sizes]
            chunks = np.empty((sum(per), 2), dtype=np.int32)
            chunks[:, 0] = np.repeat(np.arange(len(per), dtype=np.int32), per)
            chunks[:, 1] = np.concatenate([np.arange(k, dtype=np.int32) for k in per])
            self._chunk_cache = {sizes: chunks}
        dev = live[0].device
        raw = rows.tobytes() + chunks.tobytes()
        host = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
        buf = torch.empty((len(raw),), device=dev, dtype=torch.uint8)
        buf.copy_(host, non_blocking=True)
        ent = {"key": key, "buf": buf, "host": host, "n_tensors": len(live), "n_chunks": len(chunks), "chunk_off": 48 * len(live)}
        self._tables[gi] = ent
        self._pin_if_capturing(ent)
        return ent

    def _pin_if_capturing(self, ent):
        """A graph captured while this table is current keeps its raw device pointer (and, if the upload itself was captured,
        a memcpy node that re-reads the pinned source at every replay): such a table must outlive the graph, so it is never
        freed.  Tables only ever used by eager steps are dropped when the gradient pointers change."""
        if not ent.get("pinned") and torch.cuda.is_current_stream_capturing():
            ent["pinned"] = True
            self._keep.append((ent["host"], ent["buf"]))

    def _hyper(self, gi: int, group, dev):
        """Device copy of (lr, weight_decay) of a group.  The two floats live in a pinned host buffer that is refreshed from
        ``group['lr']`` / ``group['weight_decay']`` and copied to the device on the stepping stream at every step().  Under
        CUDA-graph capture that copy becomes a memcpy node which re-reads the pinned buffer at every replay, so a replay
        follows the LR scheduler as long as ``refresh_hyper()`` runs before it (graph.GraphedTrainStep does that)."""
        ent = group.get("_xcp_hyper")
        if ent is None or ent[1].device != dev:
            host = torch.zeros((2,), dtype=torch.float32).pin_memory()
            ent = group["_xcp_hyper"] = (host, torch.zeros((2,), device=dev, dtype=torch.float32))
        host, devt = ent
        host[0] = float(group["lr"]); host[1] = float(group["weight_decay"])
        devt.copy_(host, non_blocking=True)
        if torch.cuda.is_current_stream_capturing() and not any(k[0] is host for k in self._keep):
            self._keep.append((host, devt))               # the captured memcpy node reads `host` for the graph's lifetime
        return devt

    def refresh_hyper(self):
        """Write the groups' current lr / weight_decay into the pinned buffers a captured step re-reads (call before a
        CUDA-graph replay; eager steps refresh by themselves)."""
        for group in self.param_groups:
            ent = group.get("_xcp_hyper")
            if ent is not None:
                ent[0][0] = float(group["lr"]); ent[0][1] = float(group["weight_decay"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if not self._group_ready(gi, group):
                self._init_group_state(gi, group)
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            ent = self._table(gi, group, live)
            dev = live[0].device
            b1, b2 = group["betas"]
            mx = group.get("max_norm")
            ws = None
            if mx:
                ws = group.get("_xcp_sumsq")
                if ws is None or ws.device != dev:
                    ws = group["_xcp_sumsq"] = torch.zeros((), device=dev, dtype=torch.float32)
            base = ent["buf"].data_ptr()
            hyper = self._hyper(gi, group, dev)
            _lib.call("xcp_adam_multi", ctypes.c_void_p(base), ent["n_tensors"], ctypes.c_void_p(base + ent["chunk_off"]), ent["n_chunks"],
                      float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                      float(group["weight_decay"]), int(bool(group["decoupled"])), ctypes.c_void_p(ws.data_ptr() if ws is not None else 0),
                      float(mx or 0.0), 1.0, ctypes.c_void_p(hyper.data_ptr()),
                      dev.index if dev.index is not None else torch.cuda.current_device(),
                      ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if mx:
                _lib.add_launches(1)
        bump_param_epoch()               # parameters changed behind Tensor._version: derived bf16 packs are stale
        return loss

    def grad_norm(self, gi: int = 0) -> torch.Tensor:
        """Global gradient norm measured by the last clipped step of group gi (device scalar; what clip_grad_norm_ returns)."""
        ws = self.param_groups[gi].get("_xcp_sumsq")
        if ws is None:
            raise XcpError("FusedAdam.grad_norm: the group was not stepped with max_norm set")
        return ws.sqrt()

    def state_dict(self):
        sd = super().state_dict()
        for g in sd["param_groups"]:
            for k in [k for k in g if k.startswith("_xcp_")]:
                del g[k]
        # torch.optim.Adam layout: a float `step` tensor per parameter, detached copies of the moments.  Optimizer.state_dict()
        # hands out the SAME inner per-parameter dicts as self.state, so the export is built from copies of them: writing the
        # clones back would detach the live state from the flat arenas (and from a captured graph that keeps updating them).
        out = {}
        for k, live in sd["state"].items():
            st = dict(live)
            if torch.is_tensor(st.get("step")):
                st["step"] = st["step"].detach().clone().to(torch.float32)
            for name in ("exp_avg", "exp_avg_sq"):
                if name in st:
                    st[name] = st[name].detach().clone()
            out[k] = st
        sd["state"] = out
        return sd
