"""fp32 validation plan of the Xception forward (SURVEY.md §8 row d: "fp32 logits within 1e-4 relative").

The production plan (executor.py) computes in bf16 on the tensor cores.  This module walks the SAME module tree and the
same plan structure (stem, 12 blocks with their skip paths, exit flow, GAP, fc), but every arithmetic step runs in the
plain-fp32 kernels of csrc/f32.cu on fp32 NHWC activations, reading the fp32 master parameters directly.  It exists so
that a maintainer can check a checkpoint / a layout change against the reference at fp32 tolerance on the GPU; it is
forward-only and not tuned (select it with ``Xception.set_precision("fp32")`` or ``XCP_PRECISION=fp32``).

BatchNorm follows the module's mode exactly like the production plan: train mode takes batch statistics (double
accumulation), updates running_mean / running_var (unbiased) / num_batches_tracked; eval mode folds the running stats.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import XcpError
from .ops import _p, _s, F32


def _dev(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def conv3x3(x, w, stride: int, nchw: bool):
    """Dense 3x3, padding 0 (Xception.py:117,121).  x: fp32 NCHW (stem input) or NHWC; returns NHWC."""
    if nchw:
        F_, Ci, H, W = x.shape
    else:
        F_, H, W, Ci = x.shape
    Co = w.shape[0]
    out = torch.empty((F_, (H - 3) // stride + 1, (W - 3) // stride + 1, Co), device=x.device, dtype=F32)
    _lib.call("xcp_f32_conv3x3", _p(x), int(nchw), _p(w), _p(out), F_, H, W, Ci, Co, stride, _dev(x), _s())
    return out


def dw3x3(x, w):
    F_, H, W, C = x.shape
    out = torch.empty_like(x)
    _lib.call("xcp_f32_dw3x3", _p(x), _p(w), _p(out), F_, H, W, C, _dev(x), _s())
    return out


def gemm(a2d, w2d, bias=None):
    M, K = a2d.shape
    N = w2d.shape[0]
    out = torch.empty((M, N), device=a2d.device, dtype=F32)
    _lib.call("xcp_f32_gemm", _p(a2d), _p(w2d), _p(bias), _p(out), M, N, K, _dev(a2d), _s())
    return out


def affine(y, scale, shift, relu: bool):
    out = torch.empty_like(y)
    _lib.call("xcp_f32_affine", _p(y), _p(scale), _p(shift), int(relu), _p(out), y.numel(), y.shape[-1], _dev(y), _s())
    return out


def batchnorm(bn, y, relu: bool):
    """nn.BatchNorm2d forward on an NHWC fp32 tensor (+ optional ReLU), with the module's train/eval semantics."""
    C = y.shape[-1]
    M = y.numel() // C
    training = bn.training or bn.running_mean is None
    g, b = bn.weight.detach(), bn.bias.detach()
    if training:
        nparts = _lib.call("xcp_f32_bn_stats_parts", M)
        parts = torch.empty((nparts, 2, C), device=y.device, dtype=F32)
        _lib.call("xcp_f32_bn_stats", _p(y), _p(parts), M, C, _dev(y), _s())
        track = bn.track_running_stats and bn.running_mean is not None
        mom = bn.momentum if bn.momentum is not None else 0.1
        st = ops.bn_finalize(parts, M, g, b, bn.running_mean if track else None, bn.running_var if track else None, True, mom, bn.eps, C=C)
        if track:
            bn.num_batches_tracked += 1
    else:
        st = ops.bn_finalize(None, M, g, b, bn.running_mean, bn.running_var, False, 0.0, bn.eps, C=C)
    return affine(y, st.scale, st.shift, relu)


def sepconv(sep, x):
    """SeparableConv2d.forward (Xception.py:44-47): depthwise 3x3 s1 p1, then the 1x1."""
    F_, H, W, C = x.shape
    d = dw3x3(x, sep.conv1.weight.detach())
    pw = sep.pointwise.weight.detach()
    return gemm(d.view(-1, C), pw.view(pw.shape[0], C)).view(F_, H, W, pw.shape[0])


def block(spec, inp):
    """Block.forward (Xception.py:89-99)."""
    x = inp
    for u in spec.units:
        if u.relu:
            x = affine(x, None, None, True)
        x = batchnorm(u.bn, sepconv(u.sep, x), False)
    F_, H, W, C = x.shape
    if spec.skip is not None:
        s = spec.stride
        Hi, Wi, Ci = inp.shape[1], inp.shape[2], inp.shape[3]
        if s != 1:
            g = torch.empty((F_, (Hi - 1) // s + 1, (Wi - 1) // s + 1, Ci), device=inp.device, dtype=F32)
            _lib.call("xcp_f32_gather", _p(inp), _p(g), F_, Hi, Wi, Ci, s, _dev(inp), _s())
        else:
            g = inp
        sw = spec.skip.weight.detach()
        sk = gemm(g.view(-1, Ci), sw.view(sw.shape[0], Ci)).view(g.shape[0], g.shape[1], g.shape[2], sw.shape[0])
        sk = batchnorm(spec.skipbn, sk, False)
    else:
        sk = inp
    if spec.stride != 1:
        if spec.stride != 2:
            raise XcpError("fp32 plan: only stride 1 and 2 blocks exist in Xception")
        out = torch.empty_like(sk)
        _lib.call("xcp_f32_pool_add", _p(x), _p(sk), _p(out), F_, H, W, C, _dev(x), _s())
        return out
    out = torch.empty_like(x)
    _lib.call("xcp_f32_add", _p(x), _p(sk), _p(out), x.numel(), _dev(x), _s())
    return out


@torch.no_grad()
def xception_features(net, x: torch.Tensor) -> torch.Tensor:
    """conv1 ... bn4/relu/GAP (Xception.py:168-198) in fp32: [F,3,H,W] fp32 -> [F,2048] fp32."""
    if not x.is_cuda:
        raise XcpError("fp32 plan: expected a CUDA tensor (this package has no CPU path)")
    if x.dtype == torch.uint8:                               # raw NHWC frames -> the reference's [0,1] NCHW tensor
        x = x.permute(0, 3, 1, 2).to(F32).div_(255.0)
    x = x.to(F32).contiguous()
    y = batchnorm(net.bn1, conv3x3(x, net.conv1.weight.detach(), 2, True), True)
    y = batchnorm(net.bn2, conv3x3(y, net.conv2.weight.detach(), 1, False), True)
    for spec in net._block_specs:
        y = block(spec, y)
    e3, e4 = net._exit_specs
    y = batchnorm(e3.bn, sepconv(e3.sep, y), True)
    y = batchnorm(e4.bn, sepconv(e4.sep, y), True)
    F_, H, W, C = y.shape
    feat = torch.empty((F_, C), device=y.device, dtype=F32)
    _lib.call("xcp_f32_gap", _p(y), _p(feat), F_, H * W, C, _dev(y), _s())
    return feat


@torch.no_grad()
def linear(x2d, weight, bias=None):
    return gemm(x2d.contiguous(), weight.detach().contiguous(), None if bias is None else bias.detach())


@torch.no_grad()
def lstm(mod, x):
    """nn.LSTM(2048, H, 1, batch_first=True) forward in fp32 (XceptionLSTMV.py:18-23,67): -> (out [B,T,H], h_n, c_n)."""
    B, T, I = x.shape
    H = mod.hidden_size
    xproj = gemm(x.to(F32).contiguous().view(B * T, I), mod.weight_ih_l0.detach().contiguous())
    out = torch.empty((B, T, H), device=x.device, dtype=F32)
    hn = torch.empty((B, H), device=x.device, dtype=F32)
    cn = torch.empty((B, H), device=x.device, dtype=F32)
    _lib.call("xcp_f32_lstm_fwd", _p(xproj), _p(mod.bias_ih_l0.detach()), _p(mod.bias_hh_l0.detach()),
              _p(mod.weight_hh_l0.detach().contiguous()), _p(out), _p(hn), _p(cn), B, T, H, _dev(x), _s())
    return out, hn, cn
