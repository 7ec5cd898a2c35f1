"""Drop-in nn.Module classes for the reference's hot path, executing on hand-written sm_100a kernels.

Same class names, constructor arguments, attribute names, registration order (=> identical seeded init and
parameters() order) and state_dict keys/shapes/dtypes as the reference:

    SeparableConv2d, Block, Xception, xception      Xception.py:37-47, 50-99, 103-201, 205-213
    XceptionLSTMV                                   XceptionLSTMV.py:9-70
    XceptionLSTMA                                   XceptionLSTMA.py:5-59
    ArcFaceHead, CBFocalLoss, LabelSmoothingBCEWithLogitsLoss
                                                    train_visual.py:455-474, train_au_face.py:423-458, train_au_patch.py:203-211

The leaf nn.Conv2d / nn.BatchNorm2d / nn.Linear / nn.LSTM objects only *hold* parameters and buffers (so
checkpoints interchange with the reference); forward never calls them -- it runs the fused plan in executor.py.
There is no CPU or eager fallback: a CPU tensor or an unsupported configuration raises XcpError.
"""
from __future__ import annotations

import math
import os
import warnings
from typing import List, Optional

import torch
import torch.nn as nn

from . import executor as ex
from . import ops
from ._lib import XcpError

__all__ = ["SeparableConv2d", "Block", "Xception", "xception", "XceptionLSTMV", "XceptionLSTMA", "ArcFaceHead",
           "CBFocalLoss", "LabelSmoothingBCEWithLogitsLoss", "FusedLSTM", "FusedLinear", "FusionHead", "AUFaceCrossDetector",
           "model_urls"]

model_urls = {
    # same URL as the reference (Xception.py:31-34); only consulted through the local torch hub cache
    "xception": "http://data.lip6.fr/cadene/pretrainedmodels/xception-43020ad28.pth"
}


def _require_cuda(x: torch.Tensor, who: str):
    if not x.is_cuda:
        raise XcpError(f"{who}: input is on {x.device}; this implementation only runs on an sm_100a (B200) device "
                       f"and has no CPU fallback")


def _cache_of(mod: nn.Module) -> ex.PackCache:
    c = mod.__dict__.get("_pack_cache")
    if c is None:
        c = ex.PackCache()
        mod.__dict__["_pack_cache"] = c
    return c


def _precision_of(mod: nn.Module) -> str:
    return mod.__dict__.get("_precision") or os.environ.get("XCP_PRECISION", "bf16")


def _set_precision(mod: nn.Module, precision: str):
    if precision not in ("bf16", "fp32"):
        raise XcpError("%s.set_precision: 'bf16' or 'fp32', got %r" % (type(mod).__name__, precision))
    mod.__dict__["_precision"] = precision
    return mod


def _to_nhwc(x: torch.Tensor, fp32: bool) -> torch.Tensor:
    """NCHW fp32 -> NHWC with the physical channel pitch: bf16 through the layout kernel, or (fp32 validation plan) a plain
    fp32 copy (data movement only)."""
    x = x.float().contiguous()
    if not fp32:
        return ops.nchw_to_nhwc(x, pad=True)
    F_, C, H, W = x.shape
    out = torch.zeros((F_, H, W, ops.phys(C)), device=x.device, dtype=torch.float32)
    out[..., :C] = x.permute(0, 2, 3, 1)
    return out


def _to_nchw(x: torch.Tensor, C: int) -> torch.Tensor:
    if x.dtype == torch.float32:
        return x[..., :C].permute(0, 3, 1, 2).contiguous()
    return ops.nhwc_to_nchw(x, C)


# ================================================================================================ SeparableConv2d
class _SepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, w_dw, w_pw):
        cache = _cache_of(mod)
        xin = _to_nhwc(x, _precision_of(mod) == "fp32")
        spec = ex.SepSpec(mod, None, w_dw.shape[0], w_pw.shape[0], False)
        t = ex.sep_forward(cache, spec, xin, None, False, [])
        ctx.mod, ctx.tape = mod, t
        return _to_nchw(t.y, w_pw.shape[0])

    @staticmethod
    def backward(ctx, dout):
        mod, t = ctx.mod, ctx.tape
        cache = _cache_of(mod)
        sink = ex.GradSink([mod.conv1.weight, mod.pointwise.weight], dout.device)
        dy = _to_nhwc(dout, t.y.dtype == torch.float32)
        dd = ex._pw_backward(cache, sink, mod.pointwise.weight, dy, t.d)
        dz, _ = ex._dw_backward(cache, sink, t, dd)
        dx = _to_nchw(dz, mod.conv1.weight.shape[0]) if ctx.needs_input_grad[1] else None
        return None, dx, sink.view(mod.conv1.weight), sink.view(mod.pointwise.weight)


class SeparableConv2d(nn.Module):
    """Depthwise k x k conv followed by a 1x1 conv (Xception.py:37-47)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, bias=False):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, in_channels, kernel_size, stride, padding, dilation, groups=in_channels, bias=bias)
        self.pointwise = nn.Conv2d(in_channels, out_channels, 1, 1, 0, 1, 1, bias=bias)

    def set_precision(self, precision: str):
        return _set_precision(self, precision)

    def _check_supported(self):
        c = self.conv1
        if not (c.kernel_size == (3, 3) and c.stride == (1, 1) and c.padding == (1, 1) and c.dilation == (1, 1)
                and c.bias is None and self.pointwise.bias is None and c.in_channels % 8 == 0
                and self.pointwise.out_channels % 8 == 0):
            raise XcpError("SeparableConv2d: the sm_100a path implements the configuration every Xception layer uses "
                           "(k=3, s=1, p=1, d=1, bias=False, channels % 8 == 0); got %r" % (c,))

    def forward(self, x):
        _require_cuda(x, "SeparableConv2d")
        self._check_supported()
        return _SepFn.apply(self, x, self.conv1.weight, self.pointwise.weight)


# ================================================================================================ Block
class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        cache = _cache_of(mod)
        inp = _to_nhwc(x, _precision_of(mod) == "fp32")
        nbt: list = []
        bt = ex.block_forward(cache, mod._spec(), inp, nbt, save=True)
        ex._bump_nbt(nbt)
        ctx.mod, ctx.tape, ctx.nparams = mod, bt, len(params)
        return _to_nchw(bt.out, mod._out)

    @staticmethod
    def backward(ctx, dout):
        mod, bt = ctx.mod, ctx.tape
        params = list(mod.parameters())
        sink = ex.GradSink(params, dout.device)
        G = _to_nhwc(dout, bt.out.dtype == torch.float32)
        gin = ex.block_backward(_cache_of(mod), sink, bt, G)
        dx = _to_nchw(gin, mod._in) if ctx.needs_input_grad[1] else None
        return (None, dx) + tuple(sink.view(p) for p in params)


class Block(nn.Module):
    """Residual block of separable convs (Xception.py:50-99); same Sequential slot layout as the reference."""

    def __init__(self, in_filters, out_filters, reps, strides=1, start_with_relu=True, grow_first=True):
        super().__init__()
        if out_filters != in_filters or strides != 1:
            self.skip = nn.Conv2d(in_filters, out_filters, 1, stride=strides, bias=False)
            self.skipbn = nn.BatchNorm2d(out_filters)
        else:
            self.skip = None
        self.relu = nn.ReLU(inplace=True)
        rep: List[nn.Module] = []
        filters = in_filters
        if grow_first:
            rep += [self.relu, SeparableConv2d(in_filters, out_filters, 3, stride=1, padding=1, bias=False), nn.BatchNorm2d(out_filters)]
            filters = out_filters
        for _ in range(reps - 1):
            rep += [self.relu, SeparableConv2d(filters, filters, 3, stride=1, padding=1, bias=False), nn.BatchNorm2d(filters)]
        if not grow_first:
            rep += [self.relu, SeparableConv2d(in_filters, out_filters, 3, stride=1, padding=1, bias=False), nn.BatchNorm2d(out_filters)]
        if not start_with_relu:
            rep = rep[1:]
        else:
            rep[0] = nn.ReLU(inplace=False)
        if strides != 1:
            rep.append(nn.MaxPool2d(3, strides, 1))
        self.rep = nn.Sequential(*rep)
        self._in, self._out, self._strides = in_filters, out_filters, strides

    def set_precision(self, precision: str):
        return _set_precision(self, precision)

    def _spec(self) -> ex.BlockSpec:
        units = []
        pending_relu = False
        mods = list(self.rep)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.ReLU):
                pending_relu = True
                i += 1
            elif isinstance(m, SeparableConv2d):
                m._check_supported()
                bn = mods[i + 1]
                units.append(ex.SepSpec(m, bn, m.conv1.in_channels, m.pointwise.out_channels, pending_relu))
                pending_relu = False
                i += 2
            elif isinstance(m, nn.MaxPool2d):
                i += 1
            else:
                raise XcpError("Block.rep holds an unexpected module %r" % (m,))
        return ex.BlockSpec(units, self._strides, self.skip, self.skipbn if self.skip is not None else None, self._in, self._out)

    def forward(self, inp):
        _require_cuda(inp, "Block")
        return _BlockFn.apply(self, inp, *self.parameters())


# ================================================================================================ small linear
class _LinearFn(torch.autograd.Function):
    """y = x W^T + b for classifier-sized layers (B rows processed 32 at a time, warp per output neuron)."""

    @staticmethod
    def forward(ctx, x, W, b):
        x2 = x.reshape(-1, x.shape[-1]).float().contiguous()
        outs = [ops.linear_small_fwd(x2[i:i + 32], W.detach(), b.detach() if b is not None else None, 0)
                for i in range(0, x2.shape[0], 32)]
        ctx.save_for_backward(x2, W)
        ctx.has_bias, ctx.xshape = b is not None, x.shape
        return torch.cat(outs, 0).view(*x.shape[:-1], W.shape[0]) if len(outs) > 1 else outs[0].view(*x.shape[:-1], W.shape[0])

    @staticmethod
    def backward(ctx, dout):
        x2, W = ctx.saved_tensors
        d2 = dout.reshape(-1, dout.shape[-1]).float().contiguous()
        dW = torch.zeros_like(W)
        db = torch.zeros((W.shape[0],), device=W.device, dtype=W.dtype) if ctx.has_bias else None
        dins = [ops.linear_small_bwd(d2[i:i + 32], None, 1.0, x2[i:i + 32], W.detach(), dW, db, want_din=ctx.needs_input_grad[0])
                for i in range(0, x2.shape[0], 32)]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = (torch.cat(dins, 0) if len(dins) > 1 else dins[0]).view(ctx.xshape)
        return dx, dW, db


class FusedLinear(nn.Linear):
    """nn.Linear with the same parameters/keys (Xception.fc, Xception.py:149,199) running on the C-ABI kernels."""

    def forward(self, x):
        _require_cuda(x, "FusedLinear")
        return _LinearFn.apply(x, self.weight, self.bias)


# ================================================================================================ Xception
class _XceptionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, *params):
        # needs_input_grad ignores torch.no_grad(); the caller records the grad mode it was entered with (Xception.features)
        save = any(ctx.needs_input_grad[2:]) and net.__dict__.get("_want_tape", True)
        feat, tape = ex.xception_forward(net, x if x.dtype == torch.uint8 else x.float(), save=save)
        ctx.net, ctx.tape = net, tape
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        net = ctx.net
        params = net._backbone_params()
        sink = ex.GradSink(params, dfeat.device, scratch_floats=2 * sum(ops.phys(p.numel()) + 4 for p in params if p.dim() == 1))
        hook = net.__dict__.get("_grad_ready_hook")
        net.__dict__["_last_sink"] = sink      # the arena of the latest backward (every p.grad is a view of it; tools / bench inspect it)
        if hook is not None:
            sink.on_ready = lambda lo, hi: hook(sink, lo, hi)
        ex.xception_backward(net, ctx.tape, dfeat.float(), sink)
        ctx.tape = None
        if hook is not None:
            hook(sink, -1, -1)     # flush
        grads = tuple(sink.view(p) if need else None for p, need in zip(params, ctx.needs_input_grad[2:]))
        return (None, None) + grads


class Xception(nn.Module):
    """Xception backbone (Xception.py:103-201) on the fused sm_100a plan."""

    def __init__(self, num_classes=1000):
        super().__init__()
        self.num_classes = num_classes
        self.conv1 = nn.Conv2d(3, 32, 3, 2, 0, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(32, 64, 3, bias=False)
        self.bn2 = nn.BatchNorm2d(64)
        self.block1 = Block(64, 128, 2, 2, start_with_relu=False, grow_first=True)
        self.block2 = Block(128, 256, 2, 2, start_with_relu=True, grow_first=True)
        self.block3 = Block(256, 728, 2, 2, start_with_relu=True, grow_first=True)
        self.block4 = Block(728, 728, 3, 1, start_with_relu=True, grow_first=True)
        self.block5 = Block(728, 728, 3, 1, start_with_relu=True, grow_first=True)
        self.block6 = Block(728, 728, 3, 1, start_with_relu=True, grow_first=True)
        self.block7 = Block(728, 728, 3, 1, start_with_relu=True, grow_first=True)
        self.block8 = Block(728, 728, 3, 1, start_with_relu=True, grow_first=True)
        self.block9 = Block(728, 728, 3, 1, start_with_relu=True, grow_first=True)
        self.block10 = Block(728, 728, 3, 1, start_with_relu=True, grow_first=True)
        self.block11 = Block(728, 728, 3, 1, start_with_relu=True, grow_first=True)
        self.block12 = Block(728, 1024, 2, 2, start_with_relu=True, grow_first=False)
        self.conv3 = SeparableConv2d(1024, 1536, 3, 1, 1)
        self.bn3 = nn.BatchNorm2d(1536)
        self.conv4 = SeparableConv2d(1536, 2048, 3, 1, 1)
        self.bn4 = nn.BatchNorm2d(2048)
        self.fc = FusedLinear(2048, num_classes)
        # ------- init weights (Xception.py:155-160) --------
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    # ---- plan plumbing (not part of state_dict)
    @property
    def _pack_cache(self) -> ex.PackCache:
        return _cache_of(self)

    @property
    def _block_specs(self):
        return [getattr(self, "block%d" % i)._spec() for i in range(1, 13)]

    @property
    def _exit_specs(self):
        return [ex.SepSpec(self.conv3, self.bn3, 1024, 1536, False), ex.SepSpec(self.conv4, self.bn4, 1536, 2048, True)]

    def _backbone_params(self) -> List[nn.Parameter]:
        """Every parameter except fc.*, in registration order."""
        fc_ids = {id(p) for p in self.fc.parameters()} if isinstance(self.fc, nn.Module) else set()
        return [p for p in self.parameters() if id(p) not in fc_ids]

    # ---- arithmetic: "bf16" = the tcgen05 production plan; "fp32" = the validation arithmetic: the SAME plan (executor.py,
    # forward and backward) on fp32 activations through the fp32 kernels of csrc/f32.cu + csrc/f32_bwd.cu
    def set_precision(self, precision: str):
        return _set_precision(self, precision)

    @property
    def precision(self) -> str:
        return _precision_of(self)

    def features(self, x):
        """conv1 ... bn4/relu/GAP of Xception.forward (Xception.py:168-198): [F,3,H,W] -> [F,2048]."""
        _require_cuda(x, "Xception")
        if x.dtype == torch.uint8:                     # raw frames, NHWC [F,H,W,3]: scaled by 1/255 inside the stem kernel
            if x.dim() != 4 or x.shape[3] != 3:
                raise XcpError("Xception: uint8 frames must be NHWC [F,H,W,3], got %s" % (tuple(x.shape),))
        elif x.dim() != 4 or x.shape[1] != 3:
            raise XcpError("Xception: expected an input of shape [F,3,H,W], got %s" % (tuple(x.shape),))
        self.__dict__["_want_tape"] = torch.is_grad_enabled()      # no_grad evaluation: nothing saved, inference plan eligible
        return _XceptionFn.apply(self, x, *self._backbone_params())

    def forward(self, x):
        x = self.features(x)
        x = self.fc(x)                                                   # Xception.py:199
        return x


def xception(pretrained=False, **kwargs):
    """Construct Xception (Xception.py:205-213).  ``pretrained=True`` loads the reference's checkpoint from the local
    torch hub cache ($TORCH_HOME/hub/checkpoints/xception-43020ad28.pth); the reference would download it, this
    environment has no network, so a missing cache entry keeps the seeded random init and warns."""
    model = Xception(**kwargs)
    if pretrained:
        # cache-only: never open a socket (the reference's model_zoo.load_url would; a box with DNS would then silently
        # swap the seeded weights for the ImageNet checkpoint and change every seeded test / bench number)
        path = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(model_urls["xception"]))
        if os.path.isfile(path):
            try:
                model.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))
            except Exception as e:  # a malformed cache entry
                warnings.warn("xception(pretrained=True): cannot load %s (%s: %s); keeping the random initialisation"
                              % (path, type(e).__name__, e))
        else:
            warnings.warn("xception(pretrained=True): no cached checkpoint at %s (this implementation never downloads); "
                          "keeping the random initialisation" % path)
    return model


# ================================================================================================ LSTM
class _LSTMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, w_ih, w_hh, b_ih, b_hh):
        cache = _cache_of(mod)
        B, T, I = x.shape
        H = w_hh.shape[1]
        xb = x if x.dtype == torch.bfloat16 else ops.cast_bf16(x.float().contiguous())
        xb = xb.contiguous().view(B * T, I)
        wih_b, _ = cache.linear(w_ih)
        _, whh_t = cache.linear(w_hh)
        xproj, _ = ops.gemm_tn(xb, wih_b, ops.EPI_F32)
        h, gates, cst, hn, cn = ops.lstm_fwd(xproj, b_ih.detach(), b_hh.detach(), whh_t, B, T, H)
        # (detached alias: keeping the Function's own output in ctx would create a reference cycle output -> grad_fn -> ctx
        #  that only the cyclic GC frees, pinning the previous step's autograd graph and its AccumulateGrad streams)
        ctx.mod, ctx.saved = mod, (xb, gates, cst, h.detach(), B, T, H, I)
        ctx.mark_non_differentiable(cn)
        return h, hn, cn

    @staticmethod
    def backward(ctx, dh, dhn, dcn):
        mod = ctx.mod
        cache = _cache_of(mod)
        xb, gates, cst, h, B, T, H, I = ctx.saved
        w_ih, w_hh, b_ih, b_hh = mod.weight_ih_l0, mod.weight_hh_l0, mod.bias_ih_l0, mod.bias_hh_l0
        sink = ex.GradSink([w_ih, w_hh, b_ih, b_hh], h.device)
        whh_b, _ = cache.linear(w_hh)
        dgates, hprev = ops.lstm_bwd(dh.float().contiguous() if dh is not None else None,
                                     dhn.float().contiguous() if dhn is not None else None, None, gates, cst, h, whh_b,
                                     sink.view(b_ih), sink.view(b_hh), B, T, H)
        ops.gemm_wgrad(dgates, xb, sink.view(w_ih))
        ops.gemm_wgrad(dgates, hprev, sink.view(w_hh))
        dx = None
        if ctx.needs_input_grad[1]:
            _, wih_t = cache.linear(w_ih)
            dx, _ = ops.gemm_tn(dgates, wih_t, ops.EPI_F32)
            dx = dx.view(B, T, I)
        return None, dx, sink.view(w_ih), sink.view(w_hh), sink.view(b_ih), sink.view(b_hh)


class FusedLSTM(nn.LSTM):
    """nn.LSTM(input, hidden, 1, batch_first=True) with identical parameters / state_dict keys
    (XceptionLSTMV.py:18-23); forward = one tcgen05 input-projection GEMM + one persistent recurrent kernel."""

    def set_precision(self, precision: str):
        if precision not in ("bf16", "fp32"):
            raise XcpError("FusedLSTM.set_precision: 'bf16' or 'fp32', got %r" % (precision,))
        self.__dict__["_precision"] = precision
        return self

    @property
    def precision(self) -> str:
        return self.__dict__.get("_precision") or os.environ.get("XCP_PRECISION", "bf16")

    def forward(self, input, hx=None):
        _require_cuda(input, "FusedLSTM")
        if hx is not None or self.num_layers != 1 or self.bidirectional or not self.batch_first or self.proj_size != 0:
            raise XcpError("FusedLSTM: only the reference's configuration is implemented "
                           "(1 layer, unidirectional, batch_first, zero initial state)")
        if input.dim() != 3:
            raise XcpError("FusedLSTM: expected [B,T,%d]" % self.input_size)
        if self.precision == "fp32":                 # forward-only validation arithmetic (fp32_plan.py)
            if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in self.parameters())):
                raise XcpError("FusedLSTM: the fp32 validation plan is forward-only -- call it under torch.no_grad()")
            from . import fp32_plan
            out, hn, cn = fp32_plan.lstm(self, input)
            return out, (hn.unsqueeze(0), cn.unsqueeze(0))
        out, hn, cn = _LSTMFn.apply(self, input, self.weight_ih_l0, self.weight_hh_l0, self.bias_ih_l0, self.bias_hh_l0)
        return out, (hn.unsqueeze(0), cn.unsqueeze(0))


# ================================================================================================ MLP head
def _head_forward(ctx, lstm_out, row_index, training, p_drop, want_logits, labels, smoothing, masks, wb):
    """fc_layers (4 x Linear+ReLU+Dropout(0.3)) + fc_out (+ sigmoid) (+ criterion) of XceptionLSTMV.py:25-44,66-70 as ONE
    launch (csrc/head_fused.cu): the last-step select is an address into the LSTM output, dropout masks are drawn in the
    kernel, the loss and dL/dz come out of the same launch when ``labels`` is given."""
    x = lstm_out if lstm_out.dim() == 3 else lstm_out.unsqueeze(1)
    x = x.float().contiguous()
    B = x.shape[0]
    if B > 32:
        raise XcpError("classifier head: at most 32 clips per call (got %d)" % B)
    ctx.set_materialize_grads(False)
    p = float(p_drop) if training else 0.0
    loss_mode = 0 if labels is None else (2 if want_logits else 1)
    y = labels.detach().float().contiguous() if labels is not None else None
    acts, z, prob, loss, dz = ops.head_mlp_fwd(x, row_index, [t.detach() for t in wb], p, masks, loss_mode, y,
                                               float(smoothing or 0.0))
    ctx.x, ctx.row_index, ctx.acts, ctx.dz, ctx.wb = x, row_index, acts, dz, wb
    ctx.prob = None if want_logits else prob.detach()        # detached: no output -> grad_fn -> ctx cycle
    ctx.scale = 1.0 / (1.0 - p) if p > 0 else 1.0
    ctx.in_shape = lstm_out.shape
    return (z if want_logits else prob), loss


def _head_backward(ctx, dout, dloss):
    wb = ctx.wb
    n_fixed = 8
    if dout is None and dloss is None:
        return (None,) * (n_fixed + len(wb))
    if dout is None:                       # the fused criterion: dz was formed by the forward launch, scaled by dloss in the kernel
        dsrc, prob, gscale = ctx.dz, None, dloss.detach().float().contiguous()
    elif dloss is None:                    # an external criterion on the sigmoid output / the logits
        dsrc, prob, gscale = dout.detach().float().contiguous(), ctx.prob, None
    else:
        d2 = dout.detach().float()
        if ctx.prob is not None:
            d2 = d2 * ctx.prob * (1.0 - ctx.prob)
        dsrc, prob, gscale = (ctx.dz * dloss.detach().float() + d2).contiguous(), None, None
    sink = ex.GradSink(list(wb), ctx.x.device)
    dx = ops.head_mlp_bwd(dsrc, prob, gscale, ctx.x, ctx.row_index, ctx.acts, ctx.scale, [t.detach() for t in wb],
                          [sink.view(t) for t in wb], want_dx=ctx.needs_input_grad[0])
    if dx is not None:
        dx = dx.view(ctx.in_shape)
    return (dx,) + (None,) * (n_fixed - 1) + tuple(sink.view(t) for t in wb)


class _HeadFn(torch.autograd.Function):
    """Head without a criterion: -> probabilities (or fc_out logits)."""

    @staticmethod
    def forward(ctx, lstm_out, row_index, training, p_drop, want_logits, labels, smoothing, masks, *wb):
        return _head_forward(ctx, lstm_out, row_index, training, p_drop, want_logits, None, None, masks, wb)[0]

    @staticmethod
    def backward(ctx, dout):
        return _head_backward(ctx, dout, None)


class _HeadLossFn(torch.autograd.Function):
    """Head + criterion in one launch: -> (probabilities or logits, loss)."""

    @staticmethod
    def forward(ctx, lstm_out, row_index, training, p_drop, want_logits, labels, smoothing, masks, *wb):
        return _head_forward(ctx, lstm_out, row_index, training, p_drop, want_logits, labels, smoothing, masks, wb)

    @staticmethod
    def backward(ctx, dout, dloss):
        return _head_backward(ctx, dout, dloss)


class _XceptionLSTMBase(nn.Module):
    def __init__(self, hidden_dim):
        super().__init__()
        self.feature_extractor = xception(pretrained=True)
        self.feature_extractor.fc = nn.Identity()
        for param in self.feature_extractor.parameters():
            param.requires_grad = False                                   # frozen at construction (XceptionLSTMV.py:15-16)
        self.lstm = FusedLSTM(input_size=2048, hidden_size=hidden_dim, num_layers=1, batch_first=True)
        self.fc_layers = nn.Sequential(
            nn.Linear(hidden_dim, 1024), nn.ReLU(), nn.Dropout(0.3),
            nn.Linear(1024, 1024), nn.ReLU(), nn.Dropout(0.3),
            nn.Linear(1024, 1024), nn.ReLU(), nn.Dropout(0.3),
            nn.Linear(1024, 1024), nn.ReLU(), nn.Dropout(0.3),
        )
        self.fc_out = nn.Linear(1024, 1)
        self.sigmoid = nn.Sigmoid()

    def set_precision(self, precision: str):
        """"bf16": tensor-core production plan; "fp32": forward-only validation arithmetic for the backbone and the LSTM
        (the MLP head is fp32 in both)."""
        self.feature_extractor.set_precision(precision)
        self.lstm.set_precision(precision)
        return self

    def _head_params(self):
        ps = []
        for k in (0, 3, 6, 9):
            ps += [self.fc_layers[k].weight, self.fc_layers[k].bias]
        return ps + [self.fc_out.weight, self.fc_out.bias]

    # ---- variable-length clips (second half).  The shipped class ignores lengths: zero-padded frames run
    # through the backbone (and, in train mode, into the BatchNorm statistics) and the head reads the LSTM output at the last,
    # possibly padded, step (XceptionLSTMV.py:55,68; video_dataloader.py:59-64).  That stays the default (drop-in).  With
    # ``use_seq_lengths = True`` (or XCP_USE_SEQ_LENGTHS=1) a ``seq_lengths`` tensor handed to extract_features / last_step /
    # forward is honoured: only the valid frames are packed through the backbone, and the clip embedding is the LSTM output at
    # step length-1 (a causal LSTM: identical to running the unpadded clip).  Pack / unpack / select are index copies.
    use_seq_lengths = False

    def _lengths_on(self, seq_lengths) -> bool:
        on = self.use_seq_lengths or os.environ.get("XCP_USE_SEQ_LENGTHS", "0") == "1"
        return bool(on) and torch.is_tensor(seq_lengths) and not seq_lengths.dtype.is_floating_point

    def _packed_features(self, frames: torch.Tensor, b: int, t: int, seq_lengths: torch.Tensor) -> torch.Tensor:
        """frames [b*t, ...] (padded) -> features [b, t, 2048] with zeros at the padded steps; only valid frames are computed."""
        lens = seq_lengths.detach().to("cpu").long().clamp_(0, t)              # the collate builds them on the host
        if int(lens.min()) < 1:
            raise XcpError("seq_lengths: every clip needs at least one valid frame, got %s" % lens.tolist())
        valid = (torch.arange(t).unsqueeze(0) < lens.unsqueeze(1)).reshape(-1)
        idx = torch.nonzero(valid).squeeze(1).to(frames.device)
        feats_v = self.feature_extractor(frames.index_select(0, idx))
        feats = torch.zeros((b * t, feats_v.shape[-1]), device=frames.device, dtype=feats_v.dtype)
        return feats.index_copy(0, idx, feats_v).view(b, t, -1)

    def last_step(self, lstm_out: torch.Tensor, seq_lengths=None) -> torch.Tensor:
        """The clip embedding: lstm_out[:, -1, :] (XceptionLSTMV.py:68; train_visual.py:569), or -- lengths honoured -- the output
        at each clip's last VALID step."""
        if not self._lengths_on(seq_lengths):
            return lstm_out[:, -1, :]
        b, t, _ = lstm_out.shape
        last = (seq_lengths.detach().long().clamp(1, t) - 1).to(lstm_out.device)
        return lstm_out[torch.arange(b, device=lstm_out.device), last]

    def _row_index(self, lstm_out, seq_lengths):
        if not self._lengths_on(seq_lengths):
            return None
        t = lstm_out.shape[1]
        return (seq_lengths.detach().long().clamp(1, t) - 1).to(lstm_out.device).contiguous()

    def _head(self, lstm_out, row_index, want_logits: bool):
        """The head on every clip of the batch: one launch per 32 clips (the kernels hold 32 rows)."""
        drop = self.fc_layers[2]
        args = (drop.training, float(drop.p), want_logits, None, None, None)
        B = lstm_out.shape[0]
        if B <= 32:
            return _HeadFn.apply(lstm_out, row_index, *args, *self._head_params())
        outs = [_HeadFn.apply(lstm_out[i:i + 32], None if row_index is None else row_index[i:i + 32].contiguous(), *args,
                              *self._head_params()) for i in range(0, B, 32)]
        return torch.cat(outs, 0)

    def forward(self, features, seq_lengths=None):
        """lstm -> last step -> fc_layers -> sigmoid(fc_out) (XceptionLSTMV.py:66-70); the head is one launch.  The optional
        second argument exists because train_visual.py's older variants pass seq_lengths; like the shipped class, lengths
        are not used unless ``use_seq_lengths`` is set."""
        lstm_out, _ = self.lstm(features)
        return self._head(lstm_out, self._row_index(lstm_out, seq_lengths), False)

    def forward_logits(self, features):
        """Same path as forward() without the final sigmoid: fc_out logits (B,1) for logit-space criteria such as
        LabelSmoothingBCEWithLogitsLoss (train_au_patch.py:203-214)."""
        lstm_out, _ = self.lstm(features)
        return self._head(lstm_out, None, True)

    def forward_loss(self, features, labels, seq_lengths=None, smoothing=None):
        """forward() and its criterion as ONE kernel (BASELINE north_star (3)): ``smoothing=None`` -> (nn.BCELoss()(probabilities,
        labels), probabilities) as train_audio.py:20,39 computes them; a float -> (LabelSmoothingBCEWithLogitsLoss(smoothing)
        (logits, labels), logits) as train_au_patch.py:203-214 does.  Same values and gradients as the two-call form (which is
        what runs for more than 32 clips: the mean over the batch spans several head launches)."""
        lstm_out, _ = self.lstm(features)
        if labels.numel() != lstm_out.shape[0]:
            raise XcpError("forward_loss: %d labels for %d clips" % (labels.numel(), lstm_out.shape[0]))
        if lstm_out.shape[0] > 32:
            out = self._head(lstm_out, self._row_index(lstm_out, seq_lengths), smoothing is not None)
            y = labels.float().view(-1, 1)
            crit = LabelSmoothingBCEWithLogitsLoss(smoothing) if smoothing is not None else BCELoss()
            return crit(out, y), out
        drop = self.fc_layers[2]
        out, loss = _HeadLossFn.apply(lstm_out, self._row_index(lstm_out, seq_lengths), drop.training, float(drop.p),
                                      smoothing is not None, labels, smoothing, None, *self._head_params())
        return loss, out


class XceptionLSTMV(_XceptionLSTMBase):
    """Video model (XceptionLSTMV.py:9-70)."""

    def extract_features(self, video_batch, device=None):
        """(B,T,3,H,W) -> (B,T,2048).  The 2nd positional argument may be a device (shipped signature,
        XceptionLSTMV.py:46) or a seq_lengths tensor (what train_visual.py:568 passes); tensors are ignored."""
        if device is not None and not torch.is_tensor(device):
            self.feature_extractor.to(device)
            video_batch = video_batch.to(device)
        if video_batch.dtype == torch.uint8:          # raw clips (B,T,H,W,3) as stored on disk (video_dataloader.py:27-35)
            b, t, h, w, c = video_batch.shape
            frames = video_batch.reshape(b * t, h, w, c)
        else:
            b, t, c, h, w = video_batch.shape
            frames = video_batch.reshape(b * t, c, h, w)
        if self._lengths_on(device):                  # honour seq_lengths: padded frames never reach the backbone
            return self._packed_features(frames, b, t, device)
        feats = self.feature_extractor(frames)
        return feats.view(b, t, -1)


class XceptionLSTMA(_XceptionLSTMBase):
    """Audio model (XceptionLSTMA.py:5-59): (B,T,3,n_mfcc) -> bilinear 64x64 -> Xception -> LSTM -> head."""

    def extract_features(self, audio_batch, device=None):
        if device is not None and not torch.is_tensor(device):
            self.feature_extractor.to(device)
            audio_batch = audio_batch.to(device)
        _require_cuda(audio_batch, "XceptionLSTMA")
        b, t, c, n = audio_batch.shape
        frames = audio_batch.reshape(b * t, c, n, 1).float().contiguous()
        frames = ops.bilinear_up(frames, 64)                              # XceptionLSTMA.py:46
        feats = self.feature_extractor(frames)
        return feats.view(b, t, feats.shape[-1])


# ================================================================================================ heads / losses
class _ArcFaceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, weight, labels, s, m, loss_mode, class_w, gamma):
        x = feats.float().contiguous()
        dw = torch.zeros_like(weight)
        logits, loss, dx = ops.arcface_loss(x, weight.detach(), labels, s, m, loss_mode, class_w, gamma, dw=dw)
        ctx.save_for_backward(dx, dw)
        ctx.mark_non_differentiable(logits)
        return logits, loss

    @staticmethod
    def backward(ctx, _dlogits, dloss):
        dx, dw = ctx.saved_tensors
        return dx * dloss, dw * dloss, None, None, None, None, None, None


class ArcFaceHead(nn.Module):
    """ArcFace logits (train_visual.py:455-474; train_au_face.py:423-442).  ``forward(features, labels)`` returns the
    margin logits like the reference.  ``loss(features, labels, ...)`` additionally fuses CrossEntropy / CB-focal and
    their gradients into the same kernel."""

    def __init__(self, feat_dim, num_classes=2, s=30.0, m=0.5):
        super().__init__()
        self.num_classes, self.s, self.m = num_classes, s, m
        self.weight = nn.Parameter(torch.randn(num_classes, feat_dim))
        nn.init.xavier_uniform_(self.weight)
        if num_classes != 2:
            raise XcpError("ArcFaceHead: the fused kernel implements the reference's binary (real/fake) head")

    def loss(self, features, labels, class_weights=None, gamma=2.0):
        """(logits, loss): CrossEntropyLoss(ArcFace(features, labels), labels) if class_weights is None, else
        CBFocalLoss (train_au_face.py:445-458)."""
        _require_cuda(features, "ArcFaceHead")
        mode = 0 if class_weights is None else 1
        return _ArcFaceFn.apply(features, self.weight, labels.long().contiguous(), self.s, self.m, mode, class_weights, gamma)

    def forward(self, features, labels=None):
        _require_cuda(features, "ArcFaceHead")
        if labels is None:
            logits, _, _ = ops.arcface_loss(features.detach().float().contiguous(), self.weight.detach(), None, self.s, self.m)
            return logits
        # differentiable logits: route through autograd with an external loss -> use torch-free composite
        return _ArcLogitsFn.apply(features, self.weight, labels.long().contiguous(), self.s, self.m)


class _ArcLogitsFn(torch.autograd.Function):
    """Margin logits as a differentiable output (for callers that apply their own criterion, train_visual.py:571-572).
    Backward re-runs the fused kernel once per class column with a one-hot upstream gradient folded in via the
    linearity of the chain rule: dL/dx = sum_c dL/dlogit_c * dlogit_c/dx."""

    @staticmethod
    def forward(ctx, feats, weight, labels, s, m):
        x = feats.float().contiguous()
        logits, _, _ = ops.arcface_loss(x, weight.detach(), labels, s, m, want_dx=False)
        ctx.save_for_backward(x, weight, labels)
        ctx.s, ctx.m = s, m
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, weight, labels = ctx.saved_tensors
        dx, dw = ops.arcface_logits_bwd(x, weight.detach(), labels, ctx.s, ctx.m, dlogits.float().contiguous())
        return dx, dw, None, None, None


class CBFocalLoss(nn.Module):
    """Class-balanced focal loss on ArcFace logits (train_au_face.py:445-458).  Weights are computed exactly as the
    reference does; use ``ArcFaceHead.loss(features, labels, class_weights=cb.class_weights, gamma=cb.gamma)``."""

    def __init__(self, samples_per_cls, beta=0.9999, gamma=2.0):
        super().__init__()
        n = [float(v) for v in samples_per_cls]
        w = [(1.0 - beta) / (1.0 - beta ** v) for v in n]
        tot = sum(w)
        w = [v / tot * len(n) for v in w]
        self.register_buffer("class_weights", torch.tensor(w, dtype=torch.float32))
        self.gamma = gamma


class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, smoothing):
        z = logits.float().contiguous().view(-1)
        probs, loss, dz = ops.bce_fwd_bwd(z, targets.float().contiguous().view(-1), smoothing)
        ctx.save_for_backward(dz)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dz,) = ctx.saved_tensors
        return (dz.view(ctx.shape) * dloss), None, None


class _BCEProbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, targets):
        p = probs.float().contiguous()
        y = targets.float().contiguous()
        if p.numel() != y.numel():
            raise XcpError("BCELoss: input %s and target %s differ in size" % (tuple(probs.shape), tuple(targets.shape)))
        loss, dp = ops.bce_prob_fwd_bwd(p, y, want_grad=ctx.needs_input_grad[0])
        ctx.dp, ctx.shape = dp, probs.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        return (ctx.dp.view(ctx.shape) * dloss), None


class BCELoss(nn.Module):
    """nn.BCELoss() (mean reduction, log clamped at -100) as the reference applies it to the sigmoid output
    (train_audio.py:20,39): loss and dL/dp in one kernel."""

    def forward(self, input, target):
        _require_cuda(input, "BCELoss")
        return _BCEProbFn.apply(input, target)


class LabelSmoothingBCEWithLogitsLoss(nn.Module):
    """train_au_patch.py:203-211 -- BCE-with-logits on targets*(1-s)+0.5*s, loss and gradient in one kernel."""

    def __init__(self, smoothing=0.1):
        super().__init__()
        self.smoothing = smoothing

    def forward(self, logits, targets):
        _require_cuda(logits, "LabelSmoothingBCEWithLogitsLoss")
        return _BCEFn.apply(logits, targets, float(self.smoothing))


# ================================================================================================ fusion head (train_au_face)
class _FusionHeadFn(torch.autograd.Function):
    """The fused region of train_au_face.py:659-674: token mean-pooling, concat, embed_head (Linear-ReLU-Dropout-
    Linear), ArcFace margin logits, CB-focal loss, alignment MSE and temporal smoothness -- ONE launch for the forward and
    the loss, ONE for every gradient (tokens, embed_head, ArcFace weight), csrc/head_fused.cu."""

    @staticmethod
    def forward(ctx, v_tok, a_tok, labels, W0, b0, W3, b3, arc_w, class_w, cfg, mask=None):
        s, m, gamma, la, lt, training, p_drop = cfg
        v = v_tok.float().contiguous(); a = a_tok.float().contiguous()
        p = float(p_drop) if training else 0.0
        out = ops.fusion_head_fwd(v, a, labels, W0.detach(), b0.detach(), W3.detach(), b3.detach(), arc_w.detach(), class_w,
                                  s, m, gamma, la, lt, p, mask)
        # only what the backward launch reads -- NOT the returned loss / logits: an output kept on ctx closes a cycle
        # (output -> grad_fn -> ctx -> output) that keeps the step's autograd graph, and with it the parameters' AccumulateGrad
        # nodes of an eager warm-up step on the default stream, alive into a later CUDA-graph capture (which then fails)
        ctx.saved = {k: out[k] for k in ("pooled", "h", "de", "darc")}
        ctx.v, ctx.a, ctx.params = v, a, (W0, b0, W3, b3, arc_w)
        ctx.cfg = (1.0 / (1.0 - p) if p > 0 else 1.0, la, lt)
        ctx.shapes = (v_tok.shape, a_tok.shape)
        ctx.mark_non_differentiable(out["logits"])
        return out["loss"], out["logits"]

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        W0, b0, W3, b3, arc_w = ctx.params
        scale, la, lt = ctx.cfg
        sink = ex.GradSink([W0, b0, W3, b3, arc_w], ctx.v.device)
        dv, da = ops.fusion_head_bwd(dloss.detach().float().contiguous(), ctx.v, ctx.a, W0.detach(), W3.detach(), ctx.saved, scale, la, lt,
                                     sink.view(W0), sink.view(b0), sink.view(W3), sink.view(b3), sink.view(arc_w),
                                     want_dv=ctx.needs_input_grad[0], want_da=ctx.needs_input_grad[1])
        if dv is not None:
            dv = dv.view(ctx.shapes[0])
        if da is not None:
            da = da.view(ctx.shapes[1])
        return (dv, da, None, sink.view(W0), sink.view(b0), sink.view(W3), sink.view(b3), sink.view(arc_w), None, None, None)


class FusionHead(nn.Module):
    """embed_head + ArcFace + CB-focal + regularisers of train_au_face.py:598-613,659-674 as one module.
    Submodule names mirror the script's objects so its checkpoint dict ({"embed", "arcface"}) maps 1:1."""

    def __init__(self, token_dim, samples_per_cls=(1, 1), s=30.0, m=0.30, beta=0.9999, gamma=2.0, lambda_align=0.2,
                 lambda_temp=0.1, p_drop=0.2):
        super().__init__()
        self.embed_head = nn.Sequential(nn.Linear(2 * token_dim, 256), nn.ReLU(inplace=True), nn.Dropout(p_drop), nn.Linear(256, 128))
        self.arcface = ArcFaceHead(128, 2, s=s, m=m)
        self.cbfocal = CBFocalLoss(list(samples_per_cls), beta=beta, gamma=gamma)
        self.lambda_align, self.lambda_temp = lambda_align, lambda_temp

    def forward(self, v_tokens, au_tokens, labels):
        """-> (loss, logits_arc)"""
        _require_cuda(v_tokens, "FusionHead")
        e = self.embed_head
        cfg = (self.arcface.s, self.arcface.m, self.cbfocal.gamma, self.lambda_align, self.lambda_temp, self.training, float(e[2].p))
        return _FusionHeadFn.apply(v_tokens, au_tokens, labels.long().contiguous(), e[0].weight, e[0].bias, e[3].weight, e[3].bias,
                                   self.arcface.weight, self.cbfocal.class_weights, cfg, None)

    @torch.no_grad()
    def predict_logits(self, v_tokens, au_tokens):
        """Inference logits s*cos (labels=None path of ArcFaceHead, train_au_face.py:715-716): the same single launch without labels."""
        e = self.embed_head
        out = ops.fusion_head_fwd(v_tokens.float().contiguous(), au_tokens.float().contiguous(), None, e[0].weight, e[0].bias,
                                  e[3].weight, e[3].bias, self.arcface.weight, None, self.arcface.s, self.arcface.m, 0.0, 0.0, 0.0)
        return out["logits"]


class AUFaceCrossDetector(nn.Module):
    """Two-stream detector with the call signature train_au_face.py:594,656 expects (the reference's
    Models/AUFaceModel.py is absent).  Built from the in-scope hot path only: a face stream
    (Xception per frame -> FusedLSTM tokens) and an audio/AU stream (bilinear 64x64 -> Xception -> FusedLSTM tokens).
    forward(videos[B,3,T,H,W], au_patches[B,Ta,3,n]) -> (logits[B,1], v_tokens[B,T,D], au_tokens[B,Ta,D])."""

    def __init__(self, num_aus=17, face_dim=512, au_dim=512, lstm_hidden=256):
        super().__init__()
        self.num_aus = num_aus
        self.face_stream = XceptionLSTMV(lstm_hidden)
        self.au_stream = XceptionLSTMA(lstm_hidden)
        self.classifier = FusedLinear(2 * lstm_hidden, 1)

    def forward(self, videos, au_patches, au_mask=None, au_weight=None):
        _require_cuda(videos, "AUFaceCrossDetector")
        if videos.dim() == 5 and videos.size(1) == 3:                 # (B,C,T,H,W) -> (B,T,C,H,W), train_au_face.py:643-644
            videos = videos.permute(0, 2, 1, 3, 4).contiguous()
        vf = self.face_stream.extract_features(videos)
        v_tokens = self.face_stream.lstm(vf)[0]
        af = self.au_stream.extract_features(au_patches)
        au_tokens = self.au_stream.lstm(af)[0]
        logits = self.classifier(torch.cat([v_tokens.mean(1), au_tokens.mean(1)], dim=1))
        return logits, v_tokens, au_tokens
