"""Thin torch-tensor wrappers over the C ABI (include/mv3d_b200.h).

Tensors are containers for device memory only; every computation below happens in
libmv3d_b200.so.  Nothing here falls back to torch math or to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import GemmDesc, WgradDesc, check, current_stream, lib, ptr

BF16 = torch.bfloat16


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def pad_channels(c: int) -> int:
    """Channel padding of the PAD layout: 16 for <= 16 channels (RGB, 9-slice BEV: K step 16, SWIZZLE_32B boxes),
    otherwise multiples of 64 (K step 64, SWIZZLE_128B boxes -- wide TMA rows beat the 25 % less padding of 48)."""
    return 16 if c <= 16 else round_up(c, 64)


FMT_BF16X2, FMT_F16E5 = 0, 1   # MV3D_FMT_* of include/mv3d_b200.h
F16E5_MAX_WEIGHT = 15.9        # fp16(4096 w) must stay finite


@dataclass
class PadAct:
    """Activation in the PAD layout, shape (B, H+1, W+1, c_pad): bf16 hi (+ optional lo) planes (fmt FMT_BF16X2), or an
    fp16 plane + an e5m2 byte plane of the same pitch (fmt FMT_F16E5; both held in bf16-typed tensors as raw bits)."""

    hi: torch.Tensor
    lo: Optional[torch.Tensor]
    B: int
    H: int
    W: int
    C: int
    fmt: int = FMT_BF16X2

    @property
    def c_pad(self) -> int:
        return self.hi.shape[-1]

    @property
    def rows(self) -> int:
        return self.B * (self.H + 1) * (self.W + 1)


def _new_pad(B, H, W, c_pad, precise, device):
    hi = torch.empty((B, H + 1, W + 1, c_pad), dtype=BF16, device=device)
    lo = torch.empty_like(hi) if precise else None
    return hi, lo


def pad_nhwc(x: torch.Tensor, precise: bool = True, fmt: int = FMT_BF16X2) -> PadAct:
    """(B,H,W,C) float32 -> PAD."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    B, H, W, Cc = x.shape
    cp = pad_channels(Cc) if fmt == FMT_BF16X2 else round_up(Cc, 64)
    hi, lo = _new_pad(B, H, W, cp, precise or fmt == FMT_F16E5, x.device)
    check(lib().mv3d_pad_nhwc_fmt(ptr(x), B, H, W, Cc, cp, ptr(hi), ptr(lo), fmt, current_stream()), "mv3d_pad_nhwc_fmt")
    return PadAct(hi, lo, B, H, W, Cc, fmt)


def im2col3x3(x: torch.Tensor, precise: bool = True, k_pad: int = 32) -> PadAct:
    """(B,H,W,C<=3) float32 -> PAD rows whose 'channels' are the 9*C im2col taps (see mv3d_im2col3x3_pad)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    B, H, W, Cc = x.shape
    assert 9 * Cc <= k_pad
    hi, lo = _new_pad(B, H, W, k_pad, precise, x.device)
    check(lib().mv3d_im2col3x3_pad(ptr(x), B, H, W, Cc, k_pad, ptr(hi), ptr(lo), current_stream()), "mv3d_im2col3x3_pad")
    return PadAct(hi, lo, B, H, W, 9 * Cc)


def conv3x3_small_cin(x: torch.Tensor, w_hwio: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = True,
                      precise: bool = True, out_fmt: int = FMT_BF16X2) -> PadAct:
    """First conv layer on a dense (B,H,W,C<=4) float32 image: one direct kernel -> PAD activation in `out_fmt`."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    B, H, W, Cc = x.shape
    cout = w_hwio.shape[-1]
    assert tuple(w_hwio.shape[:3]) == (3, 3, Cc) and cout % 8 == 0 and w_hwio.is_contiguous()
    n_pad = pad_channels(cout)
    hi, lo = _new_pad(B, H, W, n_pad, precise or out_fmt == FMT_F16E5, x.device)
    if n_pad != cout:
        hi.zero_()
        if lo is not None:
            lo.zero_()
    check(lib().mv3d_conv3x3_small_cin(ptr(x), B, H, W, Cc, ptr(w_hwio), ptr(bias), cout, int(relu), ptr(hi), ptr(lo),
                                       n_pad, out_fmt, current_stream()), "mv3d_conv3x3_small_cin")
    return PadAct(hi, lo, B, H, W, cout, out_fmt)


def unpad_nhwc(a: PadAct) -> torch.Tensor:
    out = torch.empty((a.B, a.H, a.W, a.C), dtype=torch.float32, device=a.hi.device)
    check(lib().mv3d_unpad_nhwc_fmt(ptr(a.hi), ptr(a.lo), a.B, a.H, a.W, a.C, a.c_pad, ptr(out), a.fmt, current_stream()),
          "mv3d_unpad_nhwc_fmt")
    return out


def maxpool2x2(a: PadAct) -> PadAct:
    """Network.max_pool(2,2,2,2,'VALID') on the PAD layout."""
    Ho, Wo = a.H // 2, a.W // 2
    hi, lo = _new_pad(a.B, Ho, Wo, a.c_pad, a.lo is not None, a.hi.device)
    check(lib().mv3d_maxpool2x2_pad_fmt(ptr(a.hi), ptr(a.lo), a.B, a.H, a.W, a.c_pad, ptr(hi), ptr(lo), a.fmt,
                                        current_stream()), "mv3d_maxpool2x2_pad_fmt")
    return PadAct(hi, lo, a.B, Ho, Wo, a.C, a.fmt)


@dataclass
class PackedWeight:
    """bf16 hi/lo (N, taps*cin_pad) K-major weight + fp32 bias."""

    hi: torch.Tensor
    lo: torch.Tensor
    bias: Optional[torch.Tensor]
    taps: int
    cin: int
    cin_pad: int
    cout: int
    fmt: int = FMT_BF16X2


def pack_weights(w_hwio: torch.Tensor, bias: Optional[torch.Tensor], cin_pad: Optional[int] = None,
                 fmt: int = FMT_BF16X2) -> PackedWeight:
    """HWIO (kh,kw,Cin,Cout) or (Cin,Cout) float32 -> PackedWeight (network.py:119 / :388 layouts)."""
    assert w_hwio.is_cuda and w_hwio.dtype == torch.float32
    w = w_hwio.contiguous()
    if w.dim() == 2:
        w = w.view(1, 1, *w.shape)
    kh, kw, cin, cout = w.shape
    taps = kh * kw
    cp = cin_pad or (pad_channels(cin) if fmt == FMT_BF16X2 else round_up(cin, 64))
    if fmt == FMT_F16E5 and not float(w.abs().max()) < F16E5_MAX_WEIGHT:  # once per weight version (range guard only)
        raise ValueError("f16e5 weights must satisfy |w| < %g (fp16 plane holds 4096 w)" % F16E5_MAX_WEIGHT)
    hi = torch.empty((cout, taps * cp), dtype=BF16, device=w.device)
    lo = torch.empty_like(hi)
    check(lib().mv3d_pack_weights_fmt(ptr(w), taps, cin, cout, cp, ptr(hi), ptr(lo), fmt, current_stream()),
          "mv3d_pack_weights_fmt")
    b = None if bias is None else bias.to(torch.float32).contiguous()
    return PackedWeight(hi, lo, b, taps, cin, cp, cout, fmt)


PAIR_MODE = os.environ.get("MV3D_PAIR", "1") != "0"  # mirrors pair_mode() in csrc/conv_gemm_tcgen05.cu


def set_pair_mode(on: bool) -> bool:
    """CTA-pair (cta_group::2) 3x3 conv kernel on/off; returns the previous setting."""
    global PAIR_MODE
    prev = bool(lib().mv3d_gemm_set_pair_mode(1 if on else 0))
    PAIR_MODE = bool(on)
    return prev


GEMM_EVENTS = None  # bench.py sets this to a list to time every GEMM launch with CUDA events on its stream


def gemm_kernel_name(taps: int, k_per_tap: int, n: int, passes: int, split_k: int = 1, m: int = 0, hp: int = 0) -> str:
    """Which template instantiation mv3d_conv_gemm dispatches to (mirrors dispatch_bn / launch_gemm in
    csrc/conv_gemm_tcgen05.cu) -- used to attribute per-launch timings to kernels in bench.py."""
    if (taps == 1 and hp == 0 and passes in (2, 3) and split_k > 1 and k_per_tap % 64 == 0 and n % 256 == 0
            and 64 <= m <= 512 and PAIR_MODE):
        return "fc_swapped_pair_kernel"
    if passes == 2:
        if PAIR_MODE and (n % 256 == 0 or n in (64, 128)):
            return "conv3x3_pair_kernel<%d,2>" % min(n, 256)
        return "conv3x3_reuse_kernel<%d,2>" % (128 if n > 64 else 64)
    if passes == 1 and n > 128:
        bn = 256
    else:
        bn = 128 if n > 64 else (64 if n > 32 else 32)
    if taps == 9 and k_per_tap % 64 == 0 and split_k <= 1:
        if PAIR_MODE and (n % 256 == 0 or n in (64, 128)):
            return "conv3x3_pair_kernel<%d,%d>" % (min(n, 256), passes)
        return "conv3x3_reuse_kernel<%d,%d>" % (bn, passes)
    return "conv_gemm_kernel<%d,%d,%d>" % (bn, 64 if k_per_tap % 64 == 0 else (32 if k_per_tap % 32 == 0 else 16), passes)


def _run_gemm(_flops=0.0, **kw):
    d = GemmDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    if GEMM_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
    check(lib().mv3d_conv_gemm(C.byref(d), current_stream()), "mv3d_conv_gemm")
    if GEMM_EVENTS is not None:
        e1.record(torch.cuda.current_stream())
        GEMM_EVENTS.append((e0, e1, gemm_kernel_name(d.taps, d.Cin, d.N, d.passes, d.split_k, d.M, d.Hp), float(_flops)))


def _cin_valid(a: PadAct, w: PackedWeight) -> int:
    """mv3d_gemm_desc.cin_valid: a single zero-padded 64-channel chunk with fewer real channels (the 36-channel BEV map)."""
    c = min(a.C, w.cin)
    return c if (w.taps == 9 and a.c_pad == 64 and c < 64) else 0


def conv(a: PadAct, w: PackedWeight, relu: bool = True, precise: bool = True, out_pad: bool = True,
         out_f32_dense: bool = False, mask: Optional["PadAct"] = None, mask_scale: float = 1.0,
         addend: Optional[torch.Tensor] = None, use_bias: bool = True, out_fmt: int = FMT_BF16X2, softmax_cols: int = 0,
         pool: bool = False):
    """3x3 SAME or 1x1 convolution (+bias, +ReLU) on the PAD layout.  Returns (PadAct | None, dense f32 | None).
    mask / addend: the backward-data epilogue (out = (acc + addend) gated by mask > 0), see mv3d_gemm_desc."""
    assert w.cin_pad == a.c_pad, (w.cin_pad, a.c_pad)
    assert (not precise) or a.lo is not None
    assert a.fmt == w.fmt, "activation / weight operand formats differ"
    assert a.fmt == FMT_BF16X2 or (w.taps == 9 and a.c_pad % 64 == 0 and a.lo is not None)
    passes = 2 if a.fmt == FMT_F16E5 else (3 if precise else 1)
    dev = a.hi.device
    n_pad = pad_channels(w.cout)
    out = None
    if pool:   # 2x2/2 VALID max-pool fused into the epilogue: only the pooled PAD activation is written
        assert pool_fusable(a, w) and out_pad and not out_f32_dense and mask is None and addend is None
        hi, lo = _new_pad(a.B, a.H // 2, a.W // 2, n_pad, precise or out_fmt == FMT_F16E5, dev)
        out = PadAct(hi, lo, a.B, a.H // 2, a.W // 2, w.cout, out_fmt)
        _run_gemm(_flops=2.0 * a.B * a.H * a.W * w.taps * min(a.C, w.cin) * w.cout,
                  M=a.rows, N=w.cout, Cin=a.c_pad, taps=w.taps, Hp=a.H + 1, Wp=a.W + 1, passes=passes, out_fmt=out_fmt,
                  d_a_hi=ptr(a.hi), d_a_lo=ptr(a.lo), d_w_hi=ptr(w.hi), d_w_lo=ptr(w.lo),
                  d_bias=ptr(w.bias) if use_bias else None, relu=int(relu), d_out_hi=ptr(hi), d_out_lo=ptr(lo), ld_out=n_pad,
                  d_out_f32=None, ld_f32=0, f32_dense=0, split_k=1, pool=1, cin_valid=_cin_valid(a, w))
        return out, None
    if out_pad:
        assert out_fmt == FMT_BF16X2 or w.cout % 64 == 0
        hi, lo = _new_pad(a.B, a.H, a.W, n_pad, precise or out_fmt == FMT_F16E5, dev)
        if n_pad != w.cout:
            hi.zero_()
            if lo is not None:
                lo.zero_()
        out = PadAct(hi, lo, a.B, a.H, a.W, w.cout, out_fmt)
    dense = torch.empty((a.B, a.H, a.W, w.cout), dtype=torch.float32, device=dev) if out_f32_dense else None
    _run_gemm(_flops=2.0 * a.B * a.H * a.W * w.taps * min(a.C, w.cin) * w.cout,
              M=a.rows, N=w.cout, Cin=a.c_pad, taps=w.taps, Hp=a.H + 1, Wp=a.W + 1, passes=passes, out_fmt=out_fmt,
              d_a_hi=ptr(a.hi), d_a_lo=ptr(a.lo), d_w_hi=ptr(w.hi), d_w_lo=ptr(w.lo),
              d_bias=ptr(w.bias) if use_bias else None,
              relu=int(relu), d_out_hi=ptr(out.hi) if out else None, d_out_lo=ptr(out.lo) if out else None,
              ld_out=n_pad, d_out_f32=ptr(dense), ld_f32=w.cout, f32_dense=1 if out_f32_dense else 0, split_k=1,
              d_mask_hi=ptr(mask.hi) if mask is not None else None, ld_mask=mask.c_pad if mask is not None else 0,
              mask_scale=float(mask_scale), d_addend_f32=ptr(addend),
              ld_addend=addend.shape[-1] if addend is not None else 0, softmax_cols=int(softmax_cols),
              cin_valid=_cin_valid(a, w))
    return out, dense


POOL_FUSION = os.environ.get("MV3D_POOL_FUSION", "1") != "0"   # A/B switch
POOL_MAX_WASTE = 1.15   # tiles are 128 columns wide: fuse only where the ragged last block costs < 15 % extra MMA work


def pool_fusable(a: PadAct, w: PackedWeight) -> bool:
    """Can maxpool2x2(conv3x3(a, w)) run as ONE kernel (mv3d_gemm_desc.pool)?  CTA-pair kernel with a single 64- or
    128-channel N tile, operands in a 2-/3-pass format, and a width that 128-column blocks cover without much waste."""
    if not (POOL_FUSION and PAIR_MODE and w.taps == 9 and w.cout in (64, 128) and a.c_pad % 64 == 0 and a.lo is not None):
        return False
    if a.H < 2 or a.W < 2:
        return False
    wo2 = 2 * (a.W // 2)
    return (-(-wo2 // 128)) * 128 <= POOL_MAX_WASTE * wo2


def linear(a_hi: torch.Tensor, a_lo: Optional[torch.Tensor], w: PackedWeight, relu: bool, precise: bool = True,
           out_bf16: bool = True, out_f32: bool = False, split_k: int = 1, mask_hi: Optional[torch.Tensor] = None,
           mask_scale: float = 1.0, use_bias: bool = True):
    """Network.fc: (M,K) bf16 hi/lo rows x PackedWeight -> (hi, lo, f32).  K must equal w.cin_pad."""
    M, K = a_hi.shape
    assert K == w.cin_pad and w.taps == 1
    dev = a_hi.device
    n_pad = round_up(w.cout, 16)
    hi = lo = f32 = None
    if out_bf16:
        hi = torch.zeros((M, n_pad), dtype=BF16, device=dev)
        lo = torch.zeros_like(hi) if precise else None
    if w.fmt == FMT_F16E5:
        # f16e5 rows (a_hi = fp16 plane, a_lo = byte plane, both in 16-bit containers) x f16e5 weights: the swapped
        # split-K CTA-pair kernel only (a few hundred rows, wide output)
        assert split_k > 1 and mask_hi is None and a_lo is not None and 64 <= M <= 512 and w.cout % 256 == 0 and K % 64 == 0 \
            and PAIR_MODE, "f16e5 fc operands need the swapped split-K kernel"
    if split_k > 1 and mask_hi is None:
        acc = torch.zeros((M, w.cout), dtype=torch.float32, device=dev)
        _run_gemm(_flops=2.0 * M * w.cin * w.cout,
                  M=M, N=w.cout, Cin=K, taps=1, Hp=0, Wp=0, passes=2 if w.fmt == FMT_F16E5 else (3 if precise else 1), d_a_hi=ptr(a_hi),
                  d_a_lo=ptr(a_lo), d_w_hi=ptr(w.hi), d_w_lo=ptr(w.lo), d_bias=None, relu=0, d_out_hi=None,
                  d_out_lo=None, ld_out=0, d_out_f32=ptr(acc), ld_f32=w.cout, f32_dense=0, split_k=split_k)
        if out_f32:
            f32 = torch.empty((M, w.cout), dtype=torch.float32, device=dev)
        check(lib().mv3d_bias_act(ptr(acc), M, w.cout, w.cout, ptr(w.bias), int(relu), ptr(hi), ptr(lo), n_pad,
                                  ptr(f32), w.cout, current_stream()), "mv3d_bias_act")
        return hi, lo, f32
    if out_f32:
        f32 = torch.empty((M, w.cout), dtype=torch.float32, device=dev)
    _run_gemm(_flops=2.0 * M * w.cin * w.cout,
              M=M, N=w.cout, Cin=K, taps=1, Hp=0, Wp=0, passes=3 if precise else 1, d_a_hi=ptr(a_hi),
              d_a_lo=ptr(a_lo), d_w_hi=ptr(w.hi), d_w_lo=ptr(w.lo), d_bias=ptr(w.bias) if use_bias else None,
              relu=int(relu), d_out_hi=ptr(hi), d_out_lo=ptr(lo), ld_out=n_pad, d_out_f32=ptr(f32), ld_f32=w.cout,
              f32_dense=0, split_k=1, d_mask_hi=ptr(mask_hi), ld_mask=mask_hi.stride(0) if mask_hi is not None else 0,
              mask_scale=float(mask_scale))
    return hi, lo, f32


def softmax_pairs(x: torch.Tensor, n_pairs: int) -> torch.Tensor:
    """(rows, 2*n_pairs) float32 -> pairwise softmax, same shape.  2-D inputs may be row-strided views."""
    if x.dim() == 2 and x.stride(1) == 1:
        x2, ld = x, x.stride(0)
    else:
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        ld = x2.shape[1]
    out = torch.empty((x2.shape[0], x2.shape[1]), dtype=torch.float32, device=x.device)
    check(lib().mv3d_softmax_pairs(ptr(x2), x2.shape[0], ld, n_pairs, ptr(out), out.shape[1],
                                   current_stream()), "mv3d_softmax_pairs")
    return out.view(x.shape)


# ----------------------------------------------------------------------------------------------------------------
# training-side wrappers (backward data / backward filter / pooling / bias)
# ----------------------------------------------------------------------------------------------------------------
def pack_weights_dgrad(w_hwio: torch.Tensor, cout_pad: Optional[int] = None) -> PackedWeight:
    """HWIO (kh,kw,Cin,Cout) or (Cin,Cout) float32 -> the backward-data operand: rows = Cin, K = taps*cout_pad with the
    taps flipped, so that `conv(G, packed)` / `linear(G, packed)` yields dLoss/dInput."""
    w = w_hwio.contiguous()
    if w.dim() == 2:
        w = w.view(1, 1, *w.shape)
    kh, kw, cin, cout = w.shape
    taps = kh * kw
    cp = cout_pad or pad_channels(cout)
    hi = torch.empty((cin, taps * cp), dtype=BF16, device=w.device)
    lo = torch.empty_like(hi)
    check(lib().mv3d_pack_weights_dgrad(ptr(w), taps, cin, cout, cp, ptr(hi), ptr(lo), current_stream()),
          "mv3d_pack_weights_dgrad")
    return PackedWeight(hi, lo, None, taps, cout, cp, cin)


def _run_wgrad(**kw):
    d = WgradDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    if GEMM_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
    check(lib().mv3d_conv_wgrad(C.byref(d), current_stream()), "mv3d_conv_wgrad")
    if GEMM_EVENTS is not None:
        e1.record(torch.cuda.current_stream())
        GEMM_EVENTS.append((e0, e1, "wgrad_kernel<%d,%d>" % (128 if d.Cx % 128 == 0 else (64 if d.Cx % 64 == 0 else 16), d.passes),
                            2.0 * d.P * d.taps * d.cin * d.cout))


def conv_wgrad(x: PadAct, g: PadAct, dw: torch.Tensor, precise: bool = True, accumulate: bool = True,
               tap_window: bool = True, split_rows: int = 0) -> None:
    """dw (taps, cin, cout) float32 (+)= sum_p x[p + shift_tap] (x) g[p]   (x: layer input, g: gated output gradient)."""
    taps, cin, cout = (dw.shape[0] * dw.shape[1], dw.shape[2], dw.shape[3]) if dw.dim() == 4 else dw.shape
    assert dw.is_contiguous() and dw.dtype == torch.float32
    assert x.rows == g.rows and (x.H, x.W) == (g.H, g.W)
    _run_wgrad(P=x.rows, Cx=x.c_pad, Cg=g.c_pad, cin=cin, cout=cout, taps=taps, Wp=x.W + 1, passes=3 if precise else 1,
               d_x_hi=ptr(x.hi), d_x_lo=ptr(x.lo), d_g_hi=ptr(g.hi), d_g_lo=ptr(g.lo), d_dw=ptr(dw), ld_dw=cout,
               accumulate=int(accumulate), split_rows=split_rows, tap_window=int(tap_window))


def linear_wgrad(x_hi: torch.Tensor, x_lo: Optional[torch.Tensor], g_hi: torch.Tensor, g_lo: Optional[torch.Tensor],
                 dw: torch.Tensor, precise: bool = True, accumulate: bool = True) -> None:
    """dw (in, out) float32 (+)= x^T g for row-major bf16 hi/lo x (R, in_pad), g (R, out_pad)."""
    R, kin = x_hi.shape
    assert g_hi.shape[0] == R and dw.dim() == 2 and dw.is_contiguous()
    _run_wgrad(P=R, Cx=kin, Cg=g_hi.shape[1], cin=dw.shape[0], cout=dw.shape[1], taps=1, Wp=0,
               passes=3 if precise else 1, d_x_hi=ptr(x_hi), d_x_lo=ptr(x_lo), d_g_hi=ptr(g_hi), d_g_lo=ptr(g_lo),
               d_dw=ptr(dw), ld_dw=dw.shape[1], accumulate=int(accumulate), split_rows=0, tap_window=0)


def maxpool2x2_bwd(x: PadAct, g: PadAct) -> PadAct:
    """Gradient of maxpool2x2(x) routed to the arg-max positions and gated by x > 0 (the producing conv's ReLU)."""
    assert (g.H, g.W) == (x.H // 2, x.W // 2) and g.c_pad == x.c_pad
    hi, lo = _new_pad(x.B, x.H, x.W, x.c_pad, x.lo is not None, x.hi.device)
    check(lib().mv3d_maxpool2x2_bwd_pad(ptr(x.hi), ptr(x.lo), ptr(g.hi), ptr(g.lo), x.B, x.H, x.W, x.c_pad, ptr(hi),
                                        ptr(lo), current_stream()), "mv3d_maxpool2x2_bwd_pad")
    return PadAct(hi, lo, x.B, x.H, x.W, x.C)


def bias_grad(g_hi: torch.Tensor, g_lo: Optional[torch.Tensor], n: int, db: torch.Tensor) -> None:
    """db (n) float32 += column sums of the bf16 hi/lo gradient (any leading shape, last dim = row pitch)."""
    ld = g_hi.shape[-1]
    rows = g_hi.numel() // ld
    check(lib().mv3d_bias_grad(ptr(g_hi), ptr(g_lo), rows, ld, n, ptr(db), current_stream()), "mv3d_bias_grad")


def pad_nhwc_masked(x: torch.Tensor, mask: Optional[PadAct], precise: bool = True) -> PadAct:
    B, H, W, Cc = x.shape
    cp = mask.c_pad if mask is not None else pad_channels(Cc)
    hi, lo = _new_pad(B, H, W, cp, precise, x.device)
    check(lib().mv3d_pad_nhwc_masked(ptr(x), B, H, W, Cc, cp, ptr(mask.hi) if mask is not None else None, ptr(hi),
                                     ptr(lo), current_stream()), "mv3d_pad_nhwc_masked")
    return PadAct(hi, lo, B, H, W, Cc)
