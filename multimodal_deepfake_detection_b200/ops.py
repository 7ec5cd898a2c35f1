"""Thin tensor-level wrappers over the C ABI (include/xcp.h).  PyTorch is used only for device memory
(the caching allocator owns every buffer) and for the current stream; all arithmetic happens in
libxcp_sm100.so.  Activations are bf16 NHWC tensors of shape [F, H, W, C].

Every activation-level wrapper also accepts fp32 NHWC activations and then calls the fp32 validation kernel with the same
fused signature (csrc/f32.cu, csrc/f32_bwd.cu): executor.py walks ONE plan for both arithmetics, so the chain rule that
trains in bf16 is the one checked against the fp32 oracle at 1e-4 (Xception.set_precision("fp32"))."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32

EPI_BF16, EPI_BF16_STATS, EPI_F32 = 0, 1, 2
SRC_DIRECT, SRC_RELU, SRC_POOL, SRC_GAP_RELU = 0, 1, 2, 3
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def phys(C: int) -> int:
    """Physical channel pitch of an NHWC activation with C logical channels: widths above 64 are rounded up to a
    multiple of 64 (728 -> 768) so that every pixel row is whole 128-byte lines (TMA boxes, 16-byte vectors); the
    pad channels are kept exactly zero (include/xcp.h, "Channel padding")."""
    return C if C <= 64 else (C + 63) // 64 * 64


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.XcpError(f"{name}: expected a CUDA tensor (this package has no CPU path)")
    if t.dtype != dtype:
        raise _lib.XcpError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.XcpError(f"{name}: expected a contiguous tensor")
    return t


def check_device(device: torch.device):
    _lib.call("xcp_check_device", device.index if device.index is not None else torch.cuda.current_device())


# ------------------------------------------------------------------------------------------------ GEMMs
def gemm_tn(a: torch.Tensor, b: torch.Tensor, epi: int = EPI_BF16, bias: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None, n_real: int = 0, k_real: int = 0) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """out[M,N] = a[M,K] @ b[N,K]^T.  Returns (out, stats_partials or None).  n_real / k_real: logical channel counts when N / K
    are zero-padded channel pitches (728 in 768): the MMAs skip the padding, the padded output columns come out as zeros."""
    if a.dtype == F32:
        return _gemm_tn_f32(a, b, epi)
    _chk(a, BF16, "gemm_tn.a"); _chk(b, BF16, "gemm_tn.b")
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=F32 if epi == EPI_F32 else BF16)
    stats = None
    if epi == EPI_BF16_STATS:
        stats = torch.empty((_lib.call("xcp_gemm_stats_parts", M, N, a.device.index), 2, N), device=a.device, dtype=F32)
    _lib.call("xcp_gemm_tn", _p(a), K, _p(b), K, _p(out), N, M, N, K, epi, _p(stats), _p(bias), int(n_real), int(k_real),
              a.device.index, _s())
    return out, stats


def gemm_tn_bias(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor, relu: bool, residual: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None, n_real: int = 0, k_real: int = 0) -> torch.Tensor:
    """Inference plan: out[M,N] = relu?(a[M,K] @ b[N,K]^T + bias[N] + residual[M,N]) in bf16 (b = BN-folded weights)."""
    _chk(a, BF16, "gemm_tn_bias.a"); _chk(b, BF16, "gemm_tn_bias.b"); _chk(bias, F32, "gemm_tn_bias.bias")
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and bias.numel() >= N
    if residual is not None:
        _chk(residual, BF16, "gemm_tn_bias.residual")
        assert residual.shape == (M, N)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=BF16)
    _lib.call("xcp_gemm_tn_bias", _p(a), K, _p(b), K, _p(out), N, M, N, K, _p(bias), int(relu), _p(residual), N,
              int(n_real), int(k_real), a.device.index, _s())
    return out


def pack_weight_scaled(w: torch.Tensor, row_scale: torch.Tensor) -> torch.Tensor:
    """fp32 [N,K] -> bf16 [phys(N), phys(K)] with row n multiplied by row_scale[n] (BatchNorm scale folded into the weights)."""
    _chk(w, F32, "pack_weight_scaled.w"); _chk(row_scale, F32, "pack_weight_scaled.scale")
    R, Cc = w.shape
    Rp, Cp = phys(R), phys(Cc)
    assert row_scale.numel() >= R
    out = torch.empty((Rp, Cp), device=w.device, dtype=BF16)
    _lib.call("xcp_pack_weight_scaled", _p(w), _p(row_scale), _p(out), R, Cc, Rp, Cp, w.device.index, _s())
    return out


def _f32_bn_stats(y2d: torch.Tensor) -> torch.Tensor:
    """Per-channel (sum, sum-sq) partials of an fp32 [M,C] matrix in the layout xcp_bn_finalize reads."""
    M, C = y2d.shape
    parts = torch.empty((_lib.call("xcp_f32_bn_stats_parts", M), 2, C), device=y2d.device, dtype=F32)
    _lib.call("xcp_f32_bn_stats", _p(y2d), _p(parts), M, C, y2d.device.index, _s())
    return parts


# fp32 validation GEMMs: "ffma" = plain fp32 FMA kernels (default); "split3" = the PRODUCTION tcgen05 kernels fed 3-way bf16
# splits of both fp32 operands (6-fold concatenation along the reduction dimension; csrc/f32_bwd.cu split3_kernel)
FP32_GEMM = [__import__("os").environ.get("XCP_FP32_GEMM", "ffma")]


def split3(x2d: torch.Tensor, side: int, along_rows: bool) -> torch.Tensor:
    _chk(x2d, F32, "split3.x")
    R, Cc = x2d.shape
    out = torch.empty((6 * R, Cc) if along_rows else (R, 6 * Cc), device=x2d.device, dtype=BF16)
    _lib.call("xcp_split3_bf16", _p(x2d), _p(out), R, Cc, side, int(along_rows), x2d.device.index, _s())
    return out


def _gemm_tn_f32(a: torch.Tensor, b: torch.Tensor, epi: int):
    _chk(a, F32, "gemm_tn.a"); _chk(b, F32, "gemm_tn.b (fp32 plan: fp32 weights [N,K])")
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K
    if FP32_GEMM[0] == "split3":
        out, _ = gemm_tn(split3(a, 0, False), split3(b, 1, False), EPI_F32)
    else:
        out = torch.empty((M, N), device=a.device, dtype=F32)
        _lib.call("xcp_f32_gemm", _p(a), _p(b), _p(None), _p(out), M, N, K, a.device.index, _s())
    return out, (_f32_bn_stats(out) if epi == EPI_BF16_STATS else None)


def gemm_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, ld_dw: Optional[int] = None):
    """dw[P,Q] += dy[R,P]^T @ x[R,Q]  (dw fp32, accumulated in place)."""
    assert dw.dtype == F32
    R = dy.shape[0]
    assert x.shape[0] == R
    P, Q = (dw.shape[0], dw.shape[1]) if dw.dim() == 2 else (dy.shape[1], x.shape[1])   # logical (the operands may be channel-padded)
    assert P <= dy.shape[1] and Q <= x.shape[1]
    if dy.dtype == F32:
        _chk(dy, F32, "gemm_wgrad.dy"); _chk(x, F32, "gemm_wgrad.x")
        if FP32_GEMM[0] == "split3":
            return gemm_wgrad(split3(dy, 0, True), split3(x, 1, True), dw, ld_dw)
        _lib.call("xcp_f32_gemm_wgrad", _p(dy), dy.shape[1], _p(x), x.shape[1], _p(dw), ld_dw if ld_dw is not None else Q, R, P, Q,
                  dy.device.index, _s())
        return
    _chk(dy, BF16, "gemm_wgrad.dy"); _chk(x, BF16, "gemm_wgrad.x")
    _lib.call("xcp_gemm_wgrad", _p(dy), dy.shape[1], _p(x), x.shape[1], _p(dw), ld_dw if ld_dw is not None else Q, R, P, Q,
              dy.device.index, _s())


def gemm_ref(a, b, mn_major=False):
    if not mn_major:
        M, K = a.shape; N = b.shape[0]
        out = torch.empty((M, N), device=a.device, dtype=F32)
        _lib.call("xcp_gemm_ref", _p(a), K, _p(b), K, _p(out), N, M, N, K, 0, a.device.index, _s())
    else:
        K, M = a.shape; N = b.shape[1]
        out = torch.empty((M, N), device=a.device, dtype=F32)
        _lib.call("xcp_gemm_ref", _p(a), M, _p(b), N, _p(out), N, M, N, K, 1, a.device.index, _s())
    return out


def conv3x3_gemm_fwd(x: torch.Tensor, wk: torch.Tensor, want_stats: bool = True):
    """Dense 3x3 s1 p0 conv of NHWC x [F,Hg,Wg,Cin] with packed weights wk [Cout, 9*Cin] (fp32 plan: the raw fp32
    [Cout,Cin,3,3] weight)."""
    if x.dtype == F32:
        _chk(x, F32, "conv3x3.x"); _chk(wk, F32, "conv3x3.w")
        F_, Hg, Wg, Cin = x.shape
        Cout = wk.shape[0]
        out = torch.empty((F_, Hg - 2, Wg - 2, Cout), device=x.device, dtype=F32)
        _lib.call("xcp_f32_conv3x3", _p(x), 0, _p(wk), _p(out), F_, Hg, Wg, Cin, Cout, 1, x.device.index, _s())
        return out, (_f32_bn_stats(out.view(-1, Cout)) if want_stats else None)
    _chk(x, BF16, "conv3x3.x"); _chk(wk, BF16, "conv3x3.wk")
    F_, Hg, Wg, Cin = x.shape
    Cout = wk.shape[0]
    Ho, Wo = Hg - 2, Wg - 2
    out = torch.empty((F_, Ho, Wo, Cout), device=x.device, dtype=BF16)
    stats = None
    if want_stats:
        stats = torch.empty((_lib.call("xcp_gemm_stats_parts", F_ * Hg * Wg, Cout, x.device.index), 2, Cout), device=x.device, dtype=F32)
    _lib.call("xcp_conv3x3_gemm", _p(x), _p(wk), _p(out), _p(stats), F_, Hg, Wg, Cin, Cout, Ho, Wo, 1, x.device.index, _s())
    return out, stats


def conv3x3_gemm_dgrad(dy_grid: torch.Tensor, wk_t: torch.Tensor):
    """dy_grid [F,Hg,Wg,Cout] (zero outside the valid window) -> dx [F,Hg,Wg,Cin]; wk_t [Cin, 9*Cout].
    fp32 plan: dy_grid is the plain gradient [F,Hg-2,Wg-2,Cout] and wk_t the raw fp32 [Cout,Cin,3,3] weight."""
    if dy_grid.dtype == F32:
        _chk(dy_grid, F32, "conv3x3_dgrad.dy"); _chk(wk_t, F32, "conv3x3_dgrad.w")
        F_, Ho, Wo, Cout = dy_grid.shape
        Cin = wk_t.shape[1]
        out = torch.empty((F_, Ho + 2, Wo + 2, Cin), device=dy_grid.device, dtype=F32)
        _lib.call("xcp_f32_conv3x3_dgrad", _p(dy_grid), _p(wk_t), _p(out), F_, Ho + 2, Wo + 2, Cin, Cout, dy_grid.device.index, _s())
        return out
    F_, Hg, Wg, Cout = dy_grid.shape
    Cin = wk_t.shape[0]
    out = torch.empty((F_, Hg, Wg, Cin), device=dy_grid.device, dtype=BF16)
    _lib.call("xcp_conv3x3_gemm", _p(dy_grid), _p(wk_t), _p(out), _p(None), F_, Hg, Wg, Cout, Cin, Hg, Wg, -1,
              dy_grid.device.index, _s())
    return out


def conv3x3_wgrad(dy_grid: torch.Tensor, x: torch.Tensor, gk: torch.Tensor):
    """gk[Cout, 9*Cin] (fp32) += per-tap dy_grid^T @ shifted(x): one MN-major tcgen05 launch, the 9 taps are N tiles that
    stream the same rows of both operands together."""
    _chk(dy_grid, BF16, "conv3x3_wgrad.dy"); _chk(x, BF16, "conv3x3_wgrad.x")
    F_, Hg, Wg, Cout = dy_grid.shape
    Cin = x.shape[3]
    assert gk.dtype == F32 and gk.shape == (Cout, 9 * Cin)
    _lib.call("xcp_conv3x3_wgrad", _p(dy_grid), _p(x), _p(gk), F_, Hg, Wg, Cin, Cout, x.device.index, _s())


def conv3x3_wgrad_f32(x: torch.Tensor, x_nchw: bool, dy: torch.Tensor, dw: torch.Tensor, stride: int):
    """fp32 plan: dw[Cout,Cin,3,3] += weight gradient of a dense 3x3 p0 conv (conv1: x fp32 NCHW, stride 2; conv2: NHWC, 1)."""
    _chk(x, F32, "conv3x3_wgrad_f32.x"); _chk(dy, F32, "conv3x3_wgrad_f32.dy")
    if x_nchw:
        F_, Ci, H, W = x.shape
    else:
        F_, H, W, Ci = x.shape
    _lib.call("xcp_f32_conv3x3_wgrad", _p(x), int(x_nchw), _p(dy), _p(dw), F_, H, W, Ci, dy.shape[3], stride, x.device.index, _s())


# ------------------------------------------------------------------------------------------------ stem / dw
def _stem_input(x: torch.Tensor, who: str):
    """-> (is_u8, F, H, W).  fp32 NCHW [F,3,H,W] (the reference's tensor) or uint8 NHWC [F,H,W,3] (raw frames, scaled by
    1/255 inside the kernels)."""
    if not x.is_cuda or not x.is_contiguous():
        raise _lib.XcpError(f"{who}: expected a contiguous CUDA tensor (this package has no CPU path)")
    if x.dtype == torch.uint8:
        if x.dim() != 4 or x.shape[3] != 3:
            raise _lib.XcpError(f"{who}: uint8 frames must be NHWC [F,H,W,3], got {tuple(x.shape)}")
        return True, x.shape[0], x.shape[1], x.shape[2]
    if x.dtype != F32 or x.dim() != 4 or x.shape[1] != 3:
        raise _lib.XcpError(f"{who}: expected fp32 NCHW [F,3,H,W] or uint8 NHWC [F,H,W,3], got {x.dtype} {tuple(x.shape)}")
    return False, x.shape[0], x.shape[2], x.shape[3]


def stem_conv1_fwd(x: torch.Tensor, w: torch.Tensor, act_dtype=BF16):
    _chk(w, F32, "stem_conv1.w")
    u8, F_, H, W = _stem_input(x, "stem_conv1.x")
    H1, W1 = (H - 3) // 2 + 1, (W - 3) // 2 + 1
    if act_dtype == F32:                    # fp32 plan (x already the reference's fp32 NCHW tensor)
        if u8:
            raise _lib.XcpError("stem_conv1 (fp32 plan): convert uint8 frames to the fp32 NCHW tensor first")
        y = torch.empty((F_, H1, W1, 32), device=x.device, dtype=F32)
        _lib.call("xcp_f32_conv3x3", _p(x), 1, _p(w), _p(y), F_, H, W, 3, 32, 2, x.device.index, _s())
        return y, _f32_bn_stats(y.view(-1, 32))
    y = torch.empty((F_, H1, W1, 32), device=x.device, dtype=BF16)
    parts = torch.empty((_lib.call("xcp_stem_conv1_parts", F_, H, W, x.device.index), 2, 32), device=x.device, dtype=F32)
    _lib.call("xcp_stem_conv1_fwd", _p(x), int(u8), _p(w), _p(y), _p(parts), F_, H, W, x.device.index, _s())
    return y, parts


def stem_conv1_fwd_affine(x: torch.Tensor, w: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """Inference form of the stem: relu(scale * conv1(x) + shift) in one pass (eval-mode bn1 folded), bf16 NHWC."""
    _chk(w, F32, "stem_conv1.w"); _chk(scale, F32, "stem_conv1.scale"); _chk(shift, F32, "stem_conv1.shift")
    u8, F_, H, W = _stem_input(x, "stem_conv1.x")
    out = torch.empty((F_, (H - 3) // 2 + 1, (W - 3) // 2 + 1, 32), device=x.device, dtype=BF16)
    _lib.call("xcp_stem_conv1_fwd_affine", _p(x), int(u8), _p(w), _p(scale), _p(shift), _p(out), F_, H, W, x.device.index, _s())
    return out


def stem_conv1_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor):
    u8, F_, H, W = _stem_input(x, "stem_conv1_wgrad.x")
    if dy.dtype == F32:
        return conv3x3_wgrad_f32(x, True, dy, dw, 2)
    ws = torch.empty((_lib.call("xcp_stem_conv1_wgrad_ws_bytes", F_, H, W),), device=x.device, dtype=torch.uint8)
    _lib.call("xcp_stem_conv1_wgrad", _p(x), int(u8), _p(dy), _p(dw), _p(ws), F_, H, W, x.device.index, _s())


def dw3x3_fwd(x: torch.Tensor, w9: torch.Tensor, scale=None, shift=None, relu: bool = False, out=None):
    _chk(w9, F32, "dw3x3.w9")
    F_, H, W, C = x.shape
    if out is None:
        out = torch.empty_like(x)
    if x.dtype == F32:
        _chk(x, F32, "dw3x3.x")
        _lib.call("xcp_f32_dw3x3_fused", _p(x), _p(w9), _p(scale), _p(shift), int(relu), _p(out), F_, H, W, C, x.device.index, _s())
        return out
    _chk(x, BF16, "dw3x3.x")
    _lib.call("xcp_dw3x3_fwd", _p(x), _p(w9), _p(scale), _p(shift), int(relu), _p(out), F_, H, W, C, x.device.index, _s())
    return out


def dw3x3_bwd(dD, xin, w9, scale, shift, relu, dw, add_full=None, add_half=None, bnsum=None):
    """dw: fp32 [C_real,1,3,3] gradient buffer (accumulated); bnsum: zero-filled fp32 [2,C] when scale/shift are given."""
    F_, H, W, C = xin.shape
    dz = torch.empty_like(xin)
    if scale is not None and bnsum is None:
        bnsum = torch.zeros((2, C), device=xin.device, dtype=F32)
    if xin.dtype == F32:
        _chk(dD, F32, "dw3x3_bwd.dD"); _chk(xin, F32, "dw3x3_bwd.xin")
        _lib.call("xcp_f32_dw3x3_bwd", _p(dD), _p(xin), _p(w9), _p(scale), _p(shift), int(relu), _p(dz), _p(add_full), _p(add_half),
                  _p(dw), _p(bnsum), F_, H, W, C, dw.shape[0], xin.device.index, _s())
        return dz, bnsum
    _chk(dD, BF16, "dw3x3_bwd.dD"); _chk(xin, BF16, "dw3x3_bwd.xin")
    _lib.call("xcp_dw3x3_bwd", _p(dD), _p(xin), _p(w9), _p(scale), _p(shift), int(relu), _p(dz), _p(add_full), _p(add_half),
              _p(dw), _p(bnsum), F_, H, W, C, dw.shape[0], xin.device.index, _s())
    return dz, bnsum


# ------------------------------------------------------------------------------------------------ BN & friends
class BNState:
    """Per-call BatchNorm quantities: folded affine (scale, shift) and saved (mean, rstd)."""
    __slots__ = ("scale", "shift", "mean", "rstd", "count", "training")

    def __init__(self, C, device):
        buf = torch.empty((4, C), device=device, dtype=F32)
        self.scale, self.shift, self.mean, self.rstd = buf[0], buf[1], buf[2], buf[3]
        self.count = 0.0
        self.training = True


def bn_finalize(parts: torch.Tensor, count: float, gamma, beta, running_mean, running_var, training: bool,
                momentum: float = BN_MOMENTUM, eps: float = BN_EPS, C: Optional[int] = None) -> BNState:
    """C = physical channel pitch of the activation (default: the width of `parts`, else len(gamma))."""
    Cr = gamma.shape[0]
    if C is None:
        C = parts.shape[-1] if parts is not None else Cr
    st = BNState(C, gamma.device)
    st.count = float(count)
    st.training = training
    if training:
        _lib.call("xcp_bn_finalize", _p(parts), parts.shape[0], C, Cr, float(count), _p(gamma), _p(beta), _p(running_mean),
                  _p(running_var), momentum, eps, _p(st.scale), _p(st.shift), _p(st.mean), _p(st.rstd), gamma.device.index, _s())
    else:
        _lib.call("xcp_bn_eval_affine", _p(gamma), _p(beta), _p(running_mean), _p(running_var), eps, _p(st.scale), _p(st.shift),
                  _p(st.mean), _p(st.rstd), C, Cr, gamma.device.index, _s())
    return st


def bn_act(y, scale, shift, relu: bool):
    out = torch.empty_like(y)
    if y.dtype == F32:
        _lib.call("xcp_f32_affine", _p(y), _p(scale), _p(shift), int(relu), _p(out), y.numel(), y.shape[-1], y.device.index, _s())
        return out
    _lib.call("xcp_bn_act", _p(y), _p(scale), _p(shift), int(relu), _p(out), y.numel(), y.shape[-1], y.device.index, _s())
    return out


def gather_s2(x, scale=None, shift=None, relu=False):
    F_, H, W, C = x.shape
    if x.dtype == F32:
        if scale is not None or relu:
            raise _lib.XcpError("gather_s2 (fp32 plan): plain gather only")
        out = torch.empty((F_, (H + 1) // 2, (W + 1) // 2, C), device=x.device, dtype=F32)
        _lib.call("xcp_f32_gather", _p(x), _p(out), F_, H, W, C, 2, x.device.index, _s())
        return out
    out = torch.empty((F_, (H + 1) // 2, (W + 1) // 2, C), device=x.device, dtype=BF16)
    _lib.call("xcp_gather_s2", _p(x), _p(scale), _p(shift), int(relu), _p(out), F_, H, W, C, x.device.index, _s())
    return out


def pool_add_fwd(y, scale, shift, ys, scale_s, shift_s, want_idx=True, want_ymax=False):
    """-> (out, idx) or, with want_ymax, (out, idx, ymax): ymax = the raw y at the arg-max (saved for bn_bwd_sums)."""
    F_, H, W, C = y.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty((F_, Ho, Wo, C), device=y.device, dtype=y.dtype)
    idx = torch.empty((F_, Ho, Wo, C), device=y.device, dtype=torch.uint8) if want_idx else None
    ymax = torch.empty((F_, Ho, Wo, C), device=y.device, dtype=y.dtype) if want_ymax else None
    if y.dtype == F32:
        _lib.call("xcp_f32_pool_add_fused", _p(y), _p(scale), _p(shift), _p(ys), _p(scale_s), _p(shift_s), _p(out), _p(idx), _p(ymax),
                  F_, H, W, C, y.device.index, _s())
    else:
        _lib.call("xcp_pool_add_fwd", _p(y), _p(scale), _p(shift), _p(ys), _p(scale_s), _p(shift_s), _p(out), _p(idx), _p(ymax), F_, H,
                  W, C, y.device.index, _s())
    return (out, idx, ymax) if want_ymax else (out, idx)


def bn_bwd_sums(y: torch.Tensor, G: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[2, C] fp32 = (sum G, sum G*y) over two same-shape [.., C] activations: pass 1 of the BatchNorm backward on its own."""
    C = y.shape[-1]
    n_pix = y.numel() // C
    assert G.shape == y.shape and G.dtype == y.dtype
    sums = out if out is not None else torch.empty((2, C), device=y.device, dtype=F32)
    if y.dtype == F32:
        _lib.call("xcp_f32_bn_bwd_sums", _p(y), _p(G), _p(sums), n_pix, C, y.device.index, _s())
        return sums
    _chk(y, BF16, "bn_bwd_sums.y"); _chk(G, BF16, "bn_bwd_sums.G")
    ws = torch.empty((_lib.call("xcp_bnbwd_num_parts"), 2, C), device=y.device, dtype=F32)
    _lib.call("xcp_bn_bwd_sums", _p(y), _p(G), _p(ws), _p(sums), n_pix, C, y.device.index, _s())
    return sums


def bn_add_fwd(y, scale, shift, skip, scale_s=None, shift_s=None):
    out = torch.empty_like(y)
    if y.dtype == F32:
        _lib.call("xcp_f32_bn_add", _p(y), _p(scale), _p(shift), _p(skip), _p(scale_s), _p(shift_s), _p(out), y.numel(), y.shape[-1],
                  y.device.index, _s())
        return out
    _lib.call("xcp_bn_add_fwd", _p(y), _p(scale), _p(shift), _p(skip), _p(scale_s), _p(shift_s), _p(out), y.numel(),
              y.shape[-1], y.device.index, _s())
    return out


def bn_relu_gap(y, scale, shift):
    F_, H, W, C = y.shape
    feat = torch.empty((F_, C), device=y.device, dtype=F32)
    if y.dtype == F32:
        _lib.call("xcp_f32_bn_relu_gap", _p(y), _p(scale), _p(shift), _p(feat), F_, H * W, C, y.device.index, _s())
        return feat
    _lib.call("xcp_bn_relu_gap", _p(y), _p(scale), _p(shift), _p(feat), F_, H * W, C, y.device.index, _s())
    return feat


def bn_bwd(mode: int, y, st: BNState, gamma, dgamma, dbeta, G=None, idx=None, dfeat=None, presums=None, want_dy=True,
           grid_hw: Optional[Tuple[int, int]] = None):
    """Backward through z = BN(y) (batch stats if st.training) given the gradient source described by `mode`.
    Returns dy (bf16, same shape as y, or on the (grid_h, grid_w) padded layout when grid_hw is given)."""
    F_, H, W, C = y.shape
    dev = y.device
    coef = torch.empty((3, C), device=dev, dtype=F32)
    if y.dtype == F32:
        ws = torch.empty((2, C), device=dev, dtype=F32) if presums is None else None
        dy = None
        gh = gw = 0
        if want_dy:
            if grid_hw is not None:
                gh, gw = grid_hw
                dy = torch.zeros((F_, gh, gw, C), device=dev, dtype=F32)
            else:
                dy = torch.empty_like(y)
        _lib.call("xcp_f32_bn_bwd", mode, _p(y), _p(G), _p(idx), _p(dfeat), _p(st.scale), _p(st.shift), _p(gamma), _p(st.mean),
                  _p(st.rstd), int(st.training), _p(presums), _p(ws), _p(coef), _p(dgamma), _p(dbeta), _p(dy), F_, H, W, C,
                  gamma.shape[0], gw, gh, dev.index, _s())
        return dy
    ws = None
    if presums is None:
        ws = torch.empty((_lib.call("xcp_bnbwd_num_parts"), 2, C), device=dev, dtype=F32)
    dy = None
    gh = gw = 0
    if want_dy:
        if grid_hw is not None:
            gh, gw = grid_hw
            dy = torch.empty((F_, gh, gw, C), device=dev, dtype=BF16)     # the kernel fills [0:H, 0:W]; only the border
            if gh > H:                                                     # strips need the zeros (memset plumbing)
                dy[:, H:].zero_()
            if gw > W:
                dy[:, :H, W:].zero_()
        else:
            dy = torch.empty_like(y)
    _lib.add_launches(-(1 if presums is not None else 0) - (0 if want_dy else 1))
    _lib.call("xcp_bn_bwd", mode, _p(y), _p(G), _p(idx), _p(dfeat), _p(st.scale), _p(st.shift), _p(gamma), _p(st.mean),
              _p(st.rstd), int(st.training), _p(presums), _p(ws), _p(coef), _p(dgamma), _p(dbeta), _p(dy), F_, H, W, C,
              gamma.shape[0], gw, gh, dev.index, _s())
    return dy


# ------------------------------------------------------------------------------------------------ layout / packing
def nchw_to_nhwc(x: torch.Tensor, pad: bool = False) -> torch.Tensor:
    """fp32 NCHW -> bf16 NHWC; pad=True stores the channels with the physical pitch phys(C) (zero pad channels)."""
    _chk(x, F32, "nchw_to_nhwc.x")
    F_, C, H, W = x.shape
    Cp = phys(C) if pad else C
    out = torch.empty((F_, H, W, Cp), device=x.device, dtype=BF16)
    _lib.call("xcp_nchw_to_nhwc", _p(x), _p(out), F_, C, Cp, H * W, x.device.index, _s())
    return out


def nhwc_to_nchw(x: torch.Tensor, C: Optional[int] = None) -> torch.Tensor:
    """bf16 NHWC (channel pitch x.shape[3]) -> fp32 NCHW with the first C channels (default all)."""
    _chk(x, BF16, "nhwc_to_nchw.x")
    F_, H, W, Cp = x.shape
    C = Cp if C is None else C
    out = torch.empty((F_, C, H, W), device=x.device, dtype=F32)
    _lib.call("xcp_nhwc_to_nchw", _p(x), _p(out), F_, C, Cp, H * W, x.device.index, _s())
    return out


def pack_weight(w2d: torch.Tensor, want_t: bool = True, pad: bool = False):
    """fp32 [R,Cc] -> bf16 [Rp,Cp] (+ transpose [Cp,Rp]); pad=True zero-pads both dims to phys()."""
    _chk(w2d, F32, "pack_weight.w")
    R, Cc = w2d.shape
    Rp, Cp = (phys(R), phys(Cc)) if pad else (R, Cc)
    out = torch.empty((Rp, Cp), device=w2d.device, dtype=BF16)
    out_t = torch.empty((Cp, Rp), device=w2d.device, dtype=BF16) if want_t else None
    _lib.call("xcp_pack_weight", _p(w2d), _p(out), _p(out_t), R, Cc, Rp, Cp, w2d.device.index, _s())
    return out, out_t


def pack_dw(w: torch.Tensor, pad: bool = False):
    C = w.shape[0]
    Cp = phys(C) if pad else C
    w9 = torch.empty((9, Cp), device=w.device, dtype=F32)
    _lib.call("xcp_pack_dw", _p(w), _p(w9), C, Cp, w.device.index, _s())
    return w9


def pack_multi(table: torch.Tensor, n_tensors: int, n_tiles: int):
    """One launch over a device-side table of pointwise / depthwise weights (executor.PackCache.prefetch builds it)."""
    _lib.call("xcp_pack_multi", _p(table), n_tensors, n_tiles, table.device.index, _s())


def unpack_dw_grad(g9: torch.Tensor, gw: torch.Tensor, accumulate: bool):
    _lib.call("xcp_unpack_dw_grad", _p(g9), _p(gw), g9.shape[1], int(accumulate), g9.device.index, _s())


def pack_conv3x3(w: torch.Tensor, want_t: bool = True):
    O, I = w.shape[0], w.shape[1]
    wk = torch.empty((O, 9 * I), device=w.device, dtype=BF16)
    wk_t = torch.empty((I, 9 * O), device=w.device, dtype=BF16) if want_t else None
    _lib.call("xcp_pack_conv3x3", _p(w), _p(wk), _p(wk_t), O, I, w.device.index, _s())
    return wk, wk_t


def unpack_conv3x3_grad(gk: torch.Tensor, gw: torch.Tensor):
    O, I = gw.shape[0], gw.shape[1]
    _lib.call("xcp_unpack_conv3x3_grad", _p(gk), _p(gw), O, I, gw.device.index, _s())


def bilinear_up(x: torch.Tensor, S: int = 64) -> torch.Tensor:
    """x: [F, C, n, 1] fp32 -> [F, C, S, S] fp32 (XceptionLSTMA.py:46)."""
    _chk(x, F32, "bilinear_up.x")
    F_, C, n, one = x.shape
    assert one == 1
    out = torch.empty((F_, C, S, S), device=x.device, dtype=F32)
    _lib.call("xcp_bilinear_up", _p(x), _p(out), F_ * C, n, S, x.device.index, _s())
    return out


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _chk(x, F32, "cast_bf16.x")
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _lib.call("xcp_cast_f32_bf16", _p(x), _p(out), x.numel(), x.device.index, _s())
    return out


# ------------------------------------------------------------------------------------------------ LSTM / head
def lstm_fwd(xproj, b_ih, b_hh, w_hh_t, B, T, H):
    dev = xproj.device
    h_out = torch.empty((B, T, H), device=dev, dtype=F32)
    gates = torch.empty((B, T, 4 * H), device=dev, dtype=F32)
    cst = torch.empty((B, T, H), device=dev, dtype=F32)
    hn = torch.empty((B, H), device=dev, dtype=F32)
    cn = torch.empty((B, H), device=dev, dtype=F32)
    _lib.call("xcp_lstm_fwd", _p(xproj), _p(b_ih), _p(b_hh), _p(w_hh_t), _p(h_out), _p(gates), _p(cst), _p(hn), _p(cn), B, T, H,
              dev.index, _s())
    return h_out, gates, cst, hn, cn


def lstm_bwd(dout, dhn, dcn, gates, cst, hst, w_hh_bf16, dbias_ih, dbias_hh, B, T, H):
    dev = gates.device
    dgates = torch.empty((B * T, 4 * H), device=dev, dtype=BF16)
    hprev = torch.empty((B * T, H), device=dev, dtype=BF16)
    _lib.call("xcp_lstm_bwd", _p(dout), _p(dhn), _p(dcn), _p(gates), _p(cst), _p(hst), _p(w_hh_bf16), _p(dgates), _p(hprev),
              _p(dbias_ih), _p(dbias_hh), B, T, H, dev.index, _s())
    return dgates, hprev


def linear_small_fwd(a, W, bias, act: int, mask=None, drop_scale: float = 1.0):
    B, K = a.shape
    N = W.shape[0]
    out = torch.empty((B, N), device=a.device, dtype=F32)
    _lib.call("xcp_linear_small_fwd", _p(a), _p(W), _p(bias), _p(mask), drop_scale, act, _p(out), B, N, K, a.device.index, _s())
    return out


def linear_small_bwd(delta_raw, out_act, drop_scale, a, W, dW, db, want_din=True):
    B, K = a.shape
    N = W.shape[0]
    din = torch.zeros((B, K), device=a.device, dtype=F32) if want_din else None
    _lib.call("xcp_linear_small_bwd", _p(delta_raw), _p(out_act), drop_scale, _p(a), _p(W), _p(dW), _p(db), _p(din), B, N, K,
              a.device.index, _s())
    return din


_HEAD_STATE = {}
_HEAD_BARS = {}
_HEAD_POOL = {}


def head_state(device: torch.device):
    """Per-device state of the fused head kernels -> (barrier counters, dropout generator).  The grid-barrier counters of the
    forward and the backward launch (zero at rest) are per (device, stream): launches on different streams may overlap and must
    not share a counter (a CUDA graph keeps the counters of the stream it was captured on).  The dropout generator {seed,
    launch counter} is device resident, so CUDA-graph replays draw fresh masks; the seed is torch's at first use
    (torch.manual_seed makes the masks reproducible), decorrelated across data-parallel ranks."""
    rng = _HEAD_STATE.get(device.index)
    if rng is None:
        import os
        rank = int(os.environ.get("RANK", device.index or 0))
        seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * rank) & 0x7FFFFFFFFFFFFFFF
        rng = _HEAD_STATE[device.index] = torch.tensor([seed, 0], dtype=torch.int64).to(device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    bars = _HEAD_BARS.get(key)
    if bars is None:
        # slots come out of one zeroed pool per device made at the first (eager) call, so that a stream first seen DURING a CUDA-graph
        # capture needs no allocation (which would live in that graph's private memory pool) and no captured fill kernel
        pool = _HEAD_POOL.get(device.index)
        if pool is None or pool[1] >= pool[0].shape[0]:
            with torch.cuda.device(device):
                pool = _HEAD_POOL[device.index] = [torch.zeros((16, 8), device=device, dtype=torch.int32), 0]
        bars = _HEAD_BARS[key] = pool[0][pool[1]]
        pool[1] += 1
    return bars, rng


def head_reseed(device: torch.device, seed: int):
    bars, rng = head_state(device)
    rng.copy_(torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64))


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else 0 for t in tensors])


def _head_rows(x, row_index):
    """x [B,T,H] fp32 contiguous -> (pointer to row 0 of the selected rows, row stride): step T-1 of every clip, or step
    row_index[b] when an int64 device tensor of per-clip steps is given."""
    B, T, H = x.shape
    base = x.data_ptr() + (0 if row_index is not None else (T - 1) * H * 4)
    return ctypes.c_void_p(base), T * H


def head_mlp_fwd(x, row_index, wb, p_drop: float = 0.0, masks=None, loss_mode: int = 0, y=None, smoothing: float = 0.0):
    """Whole classifier head (+ loss) in one launch.  x: LSTM output [B,T,H] fp32 (the last step, or step row_index[b], is
    read in place).  wb: [W0,b0,...,W4,b4].  -> (acts [4,B,Wd], z [B,1], prob [B,1], loss or None, dz or None)."""
    _chk(x, F32, "head_mlp_fwd.x")
    dev = x.device
    B, T, H = x.shape
    xp, row_stride = _head_rows(x, row_index)
    Wd = wb[0].shape[0]
    for t in wb:
        _chk(t, F32, "head_mlp_fwd.weights")
    bars, rng = head_state(dev)
    acts = torch.empty((4, B, Wd), device=dev, dtype=F32)
    z = torch.empty((B, 1), device=dev, dtype=F32)
    prob = torch.empty((B, 1), device=dev, dtype=F32)
    loss = dz = None
    if loss_mode:
        _chk(y, F32, "head_mlp_fwd.y")
        if y.numel() != B:
            raise _lib.XcpError("head loss: %d targets for %d rows" % (y.numel(), B))
        loss = torch.empty((), device=dev, dtype=F32)
        dz = torch.empty((B, 1), device=dev, dtype=F32)
    if masks is not None:
        _chk(masks, torch.uint8, "head_mlp_fwd.masks")
        if tuple(masks.shape) != (4, B, Wd):
            raise _lib.XcpError("head_mlp_fwd: masks must be [4, %d, %d]" % (B, Wd))
    ptrs = _ptr_array(wb)
    _lib.call("xcp_head_mlp_fwd", xp, int(row_stride), _p(row_index), ctypes.c_void_p(ctypes.addressof(ptrs)), _p(masks),
              _p(rng) if (masks is None and p_drop > 0) else ctypes.c_void_p(0), float(p_drop), _p(acts), _p(z), _p(prob),
              int(loss_mode), _p(y), float(smoothing), _p(loss), _p(dz), _p(bars), B, H, Wd, dev.index, _s())
    return acts, z, prob, loss, dz


def head_mlp_bwd(dsrc, prob, gscale, x, row_index, acts, drop_scale: float, wb, dwb, want_dx: bool = True):
    """Backward of head_mlp_fwd in one launch: accumulates into the gradient slots dwb (entries may be None) and returns the
    gradient wrt the whole LSTM output [B,T,H] (zero-filled by the kernel, the selected rows added) or None."""
    _chk(dsrc, F32, "head_mlp_bwd.dsrc"); _chk(x, F32, "head_mlp_bwd.x")
    dev = x.device
    B, T, H = x.shape
    xp, row_stride = _head_rows(x, row_index)
    dx = torch.empty_like(x) if want_dx else None
    dxp = ctypes.c_void_p(dx.data_ptr() + (xp.value - x.data_ptr())) if want_dx else ctypes.c_void_p(0)
    Wd = wb[0].shape[0]
    bars, _ = head_state(dev)
    dacts = torch.empty((4, B, Wd), device=dev, dtype=F32)
    wp, gp = _ptr_array(wb), _ptr_array(dwb)
    _lib.call("xcp_head_mlp_bwd", _p(dsrc), _p(prob), _p(gscale), xp, int(row_stride), _p(row_index), _p(acts),
              float(drop_scale), ctypes.c_void_p(ctypes.addressof(wp)), ctypes.c_void_p(ctypes.addressof(gp)), _p(dacts), dxp, _p(dx),
              int(dx.numel()) if dx is not None else 0, ctypes.c_void_p(bars.data_ptr() + 16), B, H, Wd, dev.index, _s())
    return dx


def fusion_head_fwd(v, a, labels, W0, b0, W3, b3, arc_w, class_w, s: float, m: float, gamma: float, lambda_align: float,
                    lambda_temp: float, p_drop: float = 0.0, mask=None):
    """train_au_face.py:659-674 in one launch: token pooling, embed_head, ArcFace logits, CE / CB-focal (class_w given) + the
    regularisers.  labels None: inference logits only.  -> dict(pooled, h, e, logits, loss, de, darc) (de / darc: unit-loss
    gradients wrt the embedding / the ArcFace weight, consumed by fusion_head_bwd)."""
    _chk(v, F32, "fusion_head.v"); _chk(a, F32, "fusion_head.a")
    B, Tv, D = v.shape
    Ta = a.shape[1]
    if a.shape[0] != B or a.shape[2] != D:
        raise _lib.XcpError("fusion head: token streams %s and %s differ in batch / width" % (tuple(v.shape), tuple(a.shape)))
    N0, N3 = W0.shape[0], W3.shape[0]
    if W0.shape[1] != 2 * D or W3.shape[1] != N0 or tuple(arc_w.shape) != (2, N3):
        raise _lib.XcpError("fusion head: embed_head / ArcFace shapes do not match the tokens")
    dev = v.device
    bars, rng = head_state(dev)
    out = {"pooled": torch.empty((B, 2 * D), device=dev, dtype=F32), "h": torch.empty((B, N0), device=dev, dtype=F32),
           "e": torch.empty((B, N3), device=dev, dtype=F32), "logits": torch.empty((B, 2), device=dev, dtype=F32),
           "loss": None, "de": None, "darc": None}
    rows = torch.empty((2 * B,), device=dev, dtype=F32)
    if labels is not None:
        _chk(labels, torch.int64, "fusion_head.labels")
        out["loss"] = torch.empty((), device=dev, dtype=F32)
        out["de"] = torch.empty((B, N3), device=dev, dtype=F32)
        out["darc"] = torch.empty((B, 2, N3), device=dev, dtype=F32)
    if mask is not None:
        _chk(mask, torch.uint8, "fusion_head.mask")
    _lib.call("xcp_fusion_head_fwd", _p(v), _p(a), B, Tv, Ta, D, _p(W0), _p(b0), _p(W3), _p(b3), N0, N3, _p(arc_w), _p(labels),
              float(s), float(m), 1 if class_w is not None else 0, _p(class_w), float(gamma), float(lambda_align), float(lambda_temp),
              _p(mask), _p(rng) if (mask is None and p_drop > 0) else ctypes.c_void_p(0), float(p_drop), _p(out["pooled"]),
              _p(out["h"]), _p(out["e"]), _p(out["logits"]), _p(out["loss"]), _p(out["de"]), _p(out["darc"]), _p(rows),
              ctypes.c_void_p(bars.data_ptr() + 8), dev.index, _s())
    return out


def fusion_head_bwd(gscale, v, a, W0, W3, saved, drop_scale: float, lambda_align: float, lambda_temp: float, dW0, db0, dW3, db3,
                    darc, want_dv: bool = True, want_da: bool = True):
    """Backward of fusion_head_fwd in one launch; parameter gradients are accumulated into the given slots."""
    B, Tv, D = v.shape
    Ta = a.shape[1]
    N0, N3 = W0.shape[0], W3.shape[0]
    dev = v.device
    bars, _ = head_state(dev)
    dh = torch.empty((B, N0), device=dev, dtype=F32)
    dpooled = torch.empty((B, 2 * D), device=dev, dtype=F32)
    dv = torch.empty_like(v) if want_dv else None
    da = torch.empty_like(a) if want_da else None
    _lib.call("xcp_fusion_head_bwd", _p(gscale), _p(v), _p(a), B, Tv, Ta, D, _p(W0), _p(W3), N0, N3, _p(saved["pooled"]), _p(saved["h"]),
              _p(saved["de"]), _p(saved["darc"]), float(drop_scale), float(lambda_align), float(lambda_temp), _p(dW0), _p(db0), _p(dW3),
              _p(db3), _p(darc), _p(dh), _p(dpooled), _p(dv), _p(da), ctypes.c_void_p(bars.data_ptr() + 24), dev.index, _s())
    return dv, da


def sigmoid_fwd(z):
    p = torch.empty_like(z)
    _lib.call("xcp_sigmoid_fwd", _p(z), _p(p), z.numel(), z.device.index, _s())
    return p


def sigmoid_bwd(p, dp):
    dz = torch.empty_like(p)
    _lib.call("xcp_sigmoid_bwd", _p(p), _p(dp), _p(dz), p.numel(), p.device.index, _s())
    return dz


def bce_fwd_bwd(z, y, smoothing: float = 0.0, want_grad: bool = True):
    B = z.numel()
    probs = torch.empty((B, 1), device=z.device, dtype=F32)
    loss = torch.empty((), device=z.device, dtype=F32)
    dz = torch.empty((B, 1), device=z.device, dtype=F32) if want_grad else None
    _lib.call("xcp_bce_fwd_bwd", _p(z), _p(y), smoothing, _p(probs), _p(loss), _p(dz), B, z.device.index, _s())
    return probs, loss, dz


def bce_prob_fwd_bwd(p, y, want_grad: bool = True):
    """nn.BCELoss() on probabilities: -> (loss scalar, dL/dp)."""
    _chk(p, F32, "bce_prob.p"); _chk(y, F32, "bce_prob.y")
    n = p.numel()
    loss = torch.empty((), device=p.device, dtype=F32)
    dp = torch.empty_like(p) if want_grad else None
    _lib.call("xcp_bce_prob_fwd_bwd", _p(p), _p(y), _p(loss), _p(dp), n, p.device.index, _s())
    return loss, dp


def arcface_loss(x, w, labels, s, m, loss_mode=0, class_w=None, gamma=2.0, dw=None, want_dx=True, gscale=1.0):
    B, D = x.shape
    dev = x.device
    logits = torch.empty((B, 2), device=dev, dtype=F32)
    loss = torch.zeros((), device=dev, dtype=F32)
    rows = torch.empty((B,), device=dev, dtype=F32)
    dx = torch.empty((B, D), device=dev, dtype=F32) if (want_dx and labels is not None) else None
    _lib.call("xcp_arcface_loss", _p(x), _p(w), _p(labels), s, m, loss_mode, _p(class_w), gamma, _p(None), _p(logits), _p(loss),
              _p(rows), _p(dx), _p(dw), B, D, gscale, dev.index, _s())
    return logits, loss, dx


def arcface_logits_bwd(x, w, labels, s, m, dlogits):
    """Gradient of the margin logits wrt (x, w) for an upstream dL/dlogits computed by an external criterion."""
    B, D = x.shape
    dev = x.device
    logits = torch.empty((B, 2), device=dev, dtype=F32)
    dx = torch.empty((B, D), device=dev, dtype=F32)
    dw = torch.zeros((2, D), device=dev, dtype=F32)
    _lib.call("xcp_arcface_loss", _p(x), _p(w), _p(labels), s, m, 0, _p(None), 0.0, _p(dlogits), _p(logits), _p(None), _p(None),
              _p(dx), _p(dw), B, D, 1.0, dev.index, _s())
    return dx, dw


def fusion_pool_reg(v, a, lambda_align, lambda_temp, want_grad=True, gscale=1.0):
    B, T, D = v.shape
    dev = v.device
    pooled = torch.empty((B, 2 * D), device=dev, dtype=F32)
    loss = torch.empty((), device=dev, dtype=F32)
    dv = torch.empty_like(v) if want_grad else None
    da = torch.empty_like(a) if want_grad else None
    _lib.call("xcp_fusion_pool_reg", _p(v), _p(a), _p(pooled), _p(loss), _p(dv), _p(da), B, T, D, lambda_align, lambda_temp,
              gscale, dev.index, _s())
    return pooled, loss, dv, da


def fusion_pool_bwd(dpooled, dv, da):
    B, T, D = dv.shape
    _lib.call("xcp_fusion_pool_bwd", _p(dpooled), _p(dv), _p(da), B, T, D, dv.device.index, _s())


def grad_sumsq(g: torch.Tensor, out: torch.Tensor, zero_first=True):
    _lib.call("xcp_grad_sumsq", _p(g), g.numel(), _p(out), int(zero_first), g.device.index, _s())


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, decoupled, step, sumsq=None, max_norm=0.0, grad_scale=1.0):
    _lib.call("xcp_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, int(decoupled), step,
              _p(sumsq), max_norm, grad_scale, p.device.index, _s())
