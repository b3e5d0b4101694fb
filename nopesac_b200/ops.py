"""PyTorch-tensor wrappers over the C ABI (include/nopesac_b200.h): tensors in, tensors out, everything
enqueued on the current CUDA stream, no host synchronisation.  `launch_count()` counts the kernels of
this library that were enqueued (bench.py's `gpu_launches`)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
CAM_TYPES = {"soft": 0, "avg-all": 1, "min-cost": 2, "max-score": 3}

_launches = 0


def launch_count() -> int:
    return _launches


def reset_launch_count():
    global _launches
    _launches = 0


def _count(n=1):
    global _launches
    _launches += n


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (nopesac_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t


def _c(t: torch.Tensor, name: str, dtype=torch.float32):
    return _chk(t, name, dtype).contiguous()


def weights_version(module) -> int:
    """Sum of the in-place version counters of a module's parameters and buffers: changes whenever any of them is
    modified in place (`p.data.copy_()`, `p.mul_()`, BN running-stat edits ...), so packed weight copies keyed on it
    can never go stale silently."""
    v = 0
    for t in module.parameters():
        v += t._version
    for t in module.buffers():
        v += t._version
    return v


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
           out: Optional[torch.Tensor] = None, bias_group_rows: int = 0) -> torch.Tensor:
    """out[M,N] = act(x[M,K] @ w[N,K]^T + bias).  `x` / `out` may be column slices of wider row-major
    buffers (stride(0) is passed as the leading dimension)."""
    _chk(x, "x"); _chk(w, "w")
    assert x.dim() == 2 and w.dim() == 2 and x.stride(1) == 1 and w.is_contiguous(), "linear: layout"
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K, f"linear: K mismatch {w.shape} vs {x.shape}"
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    assert out.shape == (M, N) and out.stride(1) == 1
    if bias is not None:
        _chk(bias, "bias")
        assert bias.is_contiguous() and bias.shape[-1] == N
    ldx = x.stride(0) if M > 1 else max(K, x.stride(0))
    ldo = out.stride(0) if M > 1 else max(N, out.stride(0))
    st = _lib.lib().nsac_linear(_p(x), ldx, _p(w), _p(bias), bias_group_rows, _p(out), ldo, M, N, K, act, _stream())
    _lib.check(st, "nsac_linear")
    _count()
    return out


SPLIT_F16, SPLIT_BF16 = 0, 1
_SPLIT_DTYPE = {SPLIT_F16: torch.float16, SPLIT_BF16: torch.bfloat16}


class Split:
    """16-bit hi / lo planes of an fp32 matrix [rows, K]: x * scale ~ hi + lo — the operand format of the tcgen05
    GEMM engine.  fp16 planes (default) carry 22 significant bits and need |x * scale| <= 65504; bf16 planes carry
    16 bits with fp32's range.  Row stride `ld` is a multiple of 64 for freshly made ones (columns K..ld zero).
    `scale` is a power of two (weights are pre-scaled so that small entries stay out of fp16's subnormals)."""
    __slots__ = ("hi", "lo", "K", "fmt", "scale")

    def __init__(self, hi: torch.Tensor, lo: torch.Tensor, K: int, fmt: int = SPLIT_F16, scale: float = 1.0):
        self.hi, self.lo, self.K, self.fmt, self.scale = hi, lo, K, fmt, scale

    @property
    def rows(self):
        return self.hi.shape[0]

    def rows_view(self, a: int, b: int) -> "Split":
        return Split(self.hi[a:b], self.lo[a:b], self.K, self.fmt, self.scale)

    def cols(self, a: int, b: int) -> "Split":
        """Column slice view [rows, a:b] (keeps the parent's row stride)."""
        return Split(self.hi[:, a:b], self.lo[:, a:b], b - a, self.fmt, self.scale)

    @staticmethod
    def empty(rows: int, K: int, device, fmt: int = SPLIT_F16) -> "Split":
        ld = (K + 63) // 64 * 64
        mk = torch.zeros if ld != K else torch.empty
        return Split(mk(rows, ld, device=device, dtype=_SPLIT_DTYPE[fmt]), mk(rows, ld, device=device, dtype=_SPLIT_DTYPE[fmt]),
                     K, fmt)

    def float(self) -> torch.Tensor:
        lo = 0.0 if self.lo is None else self.lo[:, :self.K].float()
        return (self.hi[:, :self.K].float() + lo) / self.scale


def plane_overflow(clear: bool = True) -> bool:
    """True if, since the last clearing call, a finite value beyond fp16 range (|x| > 65504) was written into an fp16 plane
    (inputs or activations far outside a trained network's scale): every result since then is invalid.  Synchronises the
    current stream - a validation aid, never called on the hot path."""
    flag = C.c_int(0)
    st = _lib.lib().nsac_plane_overflow(C.byref(flag), 1 if clear else 0, _stream())
    _lib.check(st, "nsac_plane_overflow")
    return bool(flag.value)


def split(x: torch.Tensor, fmt: int = SPLIT_F16, scale: float = 1.0) -> Split:
    """fp32 [rows, K] -> Split of x * scale (zero-padded to a multiple of 64 columns)."""
    _chk(x, "x")
    assert x.dim() == 2 and x.stride(1) == 1
    rows, K = x.shape
    ld = (K + 63) // 64 * 64
    out = Split(torch.empty(rows, ld, device=x.device, dtype=_SPLIT_DTYPE[fmt]),
                torch.empty(rows, ld, device=x.device, dtype=_SPLIT_DTYPE[fmt]), K, fmt, scale)
    st = _lib.lib().nsac_split16(_p(x), x.stride(0) if rows > 1 else max(K, x.stride(0)), rows, K, scale, fmt,
                                 _p(out.hi), _p(out.lo), ld, _stream())
    _lib.check(st, "nsac_split16")
    _count()
    return out


def split_weight(w: torch.Tensor, fmt: int = SPLIT_F16) -> Split:
    """Weight planes, pre-scaled by a power of two so that max|w| lands in [4096, 8192) for fp16 planes (one-off,
    at prepare() time: reads max|w| back to the host)."""
    scale = 1.0
    if fmt == SPLIT_F16:
        mx = float(w.detach().abs().max())
        if mx > 0:
            import math
            scale = 2.0 ** math.floor(math.log2(8192.0 / mx))
    return split(w.detach().contiguous(), fmt, scale)


def gemm_tc(a: Split, w: Split, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, passes: int = 3,
            want_f32: bool = True, want_split: bool = False, out_f32: Optional[torch.Tensor] = None,
            out_split: Optional[Split] = None, bias_group_rows: int = 0, residual: Optional[Split] = None):
    """tcgen05 split-precision GEMM: act(a @ w^T + bias [+ residual]) -> (fp32 [M,N] or None, Split or None).
    `a.lo is None`: A is exact in one 16-bit plane (the lo.hi pass is skipped).  `residual`: planes [M,N] added before the
    activation (the shortcut of a bottleneck block, fused into the epilogue)."""
    M, N = a.rows, w.rows
    assert a.K == w.K, f"gemm_tc: K mismatch {a.K} vs {w.K}"
    assert a.fmt == w.fmt, "gemm_tc: operand plane formats differ"
    K = (a.K + 63) // 64 * 64
    assert a.hi.shape[1] >= K or a.hi.stride(0) >= K, "gemm_tc: A planes must be zero-padded to a multiple of 64 columns"
    assert w.hi.shape[1] >= K, "gemm_tc: W planes must be zero-padded to a multiple of 64 columns"
    dev = a.hi.device
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty(M, N, device=dev, dtype=torch.float32)
    if want_split and out_split is None:
        out_split = Split.empty(M, N, dev, a.fmt)
    if out_split is not None:
        assert out_split.fmt == a.fmt and out_split.scale == 1.0
    if bias is not None:
        _chk(bias, "bias")
        assert bias.is_contiguous() and bias.shape[-1] == N
    ldo = 0 if out_f32 is None else (out_f32.stride(0) if M > 1 else max(N, out_f32.stride(0)))
    lds = 0 if out_split is None else out_split.hi.stride(0)
    if residual is not None:
        assert bias_group_rows == 0 and residual.fmt == a.fmt and residual.scale == 1.0 and residual.rows == M and residual.K == N
        st = _lib.lib().nsac_gemm_split_residual(_p(a.hi), _p(a.lo), a.hi.stride(0), _p(w.hi), _p(w.lo), w.hi.stride(0), _p(bias),
                                                 M, N, K, act, passes, a.fmt, 1.0 / (a.scale * w.scale), _p(residual.hi),
                                                 _p(residual.lo), residual.hi.stride(0), _p(out_f32), ldo,
                                                 None if out_split is None else _p(out_split.hi),
                                                 None if out_split is None else _p(out_split.lo), lds, _stream())
        _lib.check(st, "nsac_gemm_split_residual")
        _count()
        return out_f32, out_split
    st = _lib.lib().nsac_gemm_split(_p(a.hi), _p(a.lo), a.hi.stride(0), _p(w.hi), _p(w.lo), w.hi.stride(0), _p(bias),
                                    bias_group_rows, M, N, K, act, passes, a.fmt, 1.0 / (a.scale * w.scale),
                                    _p(out_f32), ldo, None if out_split is None else _p(out_split.hi),
                                    None if out_split is None else _p(out_split.lo), lds, _stream())
    _lib.check(st, "nsac_gemm_split")
    _count()
    return out_f32, out_split


# ---------------------------------------------------------------------------------------------------
# pixel pose network pieces (NHWC activations: [N*H*W, C] rows)
# ---------------------------------------------------------------------------------------------------
def _empty_split(rows, cols, device, fmt=SPLIT_F16):
    return Split(torch.empty(rows, cols, device=device, dtype=_SPLIT_DTYPE[fmt]),
                 torch.empty(rows, cols, device=device, dtype=_SPLIT_DTYPE[fmt]), cols, fmt)


def nchw_to_planes(x: torch.Tensor, fmt: int = SPLIT_F16, out: Optional[Split] = None, row_offset: int = 0) -> Split:
    """[N,C,H,W] fp32 -> NHWC planes [N*H*W, C] (optionally into rows [row_offset, row_offset + N*H*W) of `out`)."""
    x = _c(x, "x")
    N, Cc, H, W = x.shape
    if out is None:
        out = _empty_split(N * H * W, Cc, x.device, fmt)
    assert out.hi.is_contiguous() and out.hi.shape[1] == Cc and out.rows >= row_offset + N * H * W
    dst = Split(out.hi[row_offset:], out.lo[row_offset:], Cc, fmt)
    st = _lib.lib().nsac_nchw_to_planes(_p(x), N, Cc, H * W, fmt, _p(dst.hi), _p(dst.lo), _stream())
    _lib.check(st, "nsac_nchw_to_planes")
    _count()
    return out


def conv3x3_tc(x: Split, N: int, H: int, W: int, w: Split, bias=None, act: int = ACT_NONE, passes: int = 3,
               want_f32: bool = True, want_split: bool = False, stride: int = 1):
    """3x3 / pad 1 convolution, stride 1 or 2 (implicit GEMM, 4-D TMA gather; the stride is the tensor map's traversal stride).
    x: NHWC planes [N*H*W, Cin] (contiguous), w: planes [Cout, 9*Cin] in (ky, kx, cin) order.
    -> (fp32 [N*Ho*Wo, Cout] or None, Split or None), Ho = (H-1)//stride + 1."""
    Cin, Cout = x.hi.shape[1], w.rows
    assert x.hi.is_contiguous() and x.lo.is_contiguous() and x.rows == N * H * W and Cin % 64 == 0
    assert w.hi.is_contiguous() and w.hi.shape[1] == 9 * Cin and x.fmt == w.fmt and stride in (1, 2)
    dev = x.hi.device
    rows = N * ((H - 1) // stride + 1) * ((W - 1) // stride + 1)
    out_f32 = torch.empty(rows, Cout, device=dev, dtype=torch.float32) if want_f32 else None
    out_split = Split.empty(rows, Cout, dev, x.fmt) if want_split else None
    if bias is not None:
        _chk(bias, "bias")
    tail = (act, passes, x.fmt, 1.0 / (x.scale * w.scale), _p(out_f32), Cout if want_f32 else 0,
            None if out_split is None else _p(out_split.hi), None if out_split is None else _p(out_split.lo),
            0 if out_split is None else out_split.hi.stride(0), _stream())
    if stride == 1:
        st = _lib.lib().nsac_conv3x3_split(_p(x.hi), _p(x.lo), _p(w.hi), _p(w.lo), _p(bias), N, H, W, Cin, Cout, *tail)
    else:
        st = _lib.lib().nsac_conv3x3_split_strided(_p(x.hi), _p(x.lo), _p(w.hi), _p(w.lo), _p(bias), N, H, W, Cin, Cout, stride, *tail)
    _lib.check(st, "nsac_conv3x3_split")
    _count()
    return out_f32, out_split


def conv1x1_tc_strided(x: Split, N: int, H: int, W: int, w: Split, bias=None, act: int = ACT_NONE, passes: int = 3, stride: int = 2):
    """1x1 convolution with a stride (no padding) on NHWC planes [N*H*W, Cin] -> planes [N*Ho*Wo, Cout]: the strided projection
    shortcut of a bottleneck block, gathered by TMA (no subsampled copy)."""
    Cin, Cout = x.hi.shape[1], w.rows
    assert x.hi.is_contiguous() and x.lo.is_contiguous() and x.rows == N * H * W and Cin % 64 == 0
    assert w.hi.is_contiguous() and w.hi.shape[1] == Cin and x.fmt == w.fmt and stride in (1, 2)
    rows = N * ((H - 1) // stride + 1) * ((W - 1) // stride + 1)
    out = Split.empty(rows, Cout, x.hi.device, x.fmt)
    st = _lib.lib().nsac_conv1x1_split_strided(_p(x.hi), _p(x.lo), _p(w.hi), _p(w.lo), _p(bias), N, H, W, Cin, Cout, stride, act, passes,
                                               x.fmt, 1.0 / (x.scale * w.scale), None, 0, _p(out.hi), _p(out.lo), out.hi.stride(0),
                                               _stream())
    _lib.check(st, "nsac_conv1x1_split_strided")
    _count()
    return out


def groupnorm_nhwc(x: torch.Tensor, N: int, H: int, W: int, gamma, beta, groups: int = 32, eps: float = 1e-5,
                   relu: bool = False, skip: Optional[torch.Tensor] = None, want_f32: bool = True, want_split: bool = False,
                   fmt: int = SPLIT_F16):
    """GroupNorm over NHWC rows [N*H*W, C] (+ReLU) (+ nearest-2x-upsampled `skip` [N*(H/2)*(W/2), C])."""
    x = _c(x, "x")
    Cc = x.shape[1]
    out_f32 = torch.empty_like(x) if want_f32 else None
    out_split = _empty_split(x.shape[0], Cc, x.device, fmt) if want_split else None
    ws_bytes = _lib.lib().nsac_groupnorm_ws_bytes(N, H, W, groups)
    ws = torch.empty(max(ws_bytes, 16), device=x.device, dtype=torch.uint8)
    st = _lib.lib().nsac_groupnorm_nhwc(_p(x), N, H, W, Cc, groups, _p(gamma), _p(beta), eps, int(relu), _p(skip), fmt,
                                        _p(out_f32), None if out_split is None else _p(out_split.hi),
                                        None if out_split is None else _p(out_split.lo), _p(ws), ws_bytes, _stream())
    _lib.check(st, "nsac_groupnorm_nhwc")
    _count()
    return out_f32, out_split


def maxpool2_planes(x: torch.Tensor, N: int, H: int, W: int, fmt: int = SPLIT_F16) -> Split:
    x = _c(x, "x")
    Cc = x.shape[1]
    out = _empty_split(N * (H // 2) * (W // 2), Cc, x.device, fmt)
    st = _lib.lib().nsac_maxpool2_planes(_p(x), N, H, W, Cc, fmt, _p(out.hi), _p(out.lo), _stream())
    _lib.check(st, "nsac_maxpool2_planes")
    _count()
    return out


def corr_softmax(f1: torch.Tensor, f2: torch.Tensor, B: int, H: int, W: int, fmt: int = SPLIT_F16) -> Split:
    """compute_corr_softmax on NHWC features [B*H*W, C] -> planes [B*H*W, Cp] (Cp = H*W rounded up to 64)."""
    f1, f2 = _c(f1, "f1"), _c(f2, "f2")
    Cp = (H * W + 63) // 64 * 64
    out = _empty_split(B * H * W, Cp, f1.device, fmt)
    out.K = Cp
    st = _lib.lib().nsac_corr_softmax(_p(f1), _p(f2), B, H, W, f1.shape[1], Cp, fmt, _p(out.hi), _p(out.lo), _stream())
    _lib.check(st, "nsac_corr_softmax")
    _count()
    return out


def im2col3x3_planes(x: torch.Tensor, N: int, H: int, W: int, stride: int, fmt: int = SPLIT_F16):
    """Explicit 3x3 / pad 1 im2col of an fp32 NHWC map -> (planes [N*Ho*Wo, Kp], Ho, Wo)."""
    x = _c(x, "x")
    Cc = x.shape[1]
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    Kp = (9 * Cc + 63) // 64 * 64
    out = _empty_split(N * Ho * Wo, Kp, x.device, fmt)
    out.K = 9 * Cc
    st = _lib.lib().nsac_im2col3x3_planes(_p(x), N, H, W, Cc, stride, Kp, fmt, _p(out.hi), _p(out.lo), _stream())
    _lib.check(st, "nsac_im2col3x3_planes")
    _count()
    return out, Ho, Wo


# ---------------------------------------------------------------------------------------------------
# ResNet-50 backbone glue (row f2; csrc/backbone.cu)
# ---------------------------------------------------------------------------------------------------
def stem_im2col_planes(img: torch.Tensor, mean, std, fmt: int = SPLIT_F16):
    """[N,3,H,W] fp32 -> ((img - mean) / std) im2col planes of the 7x7 / stride 2 / pad 3 stem conv: (Split [N*Ho*Wo, 192] with
    K = 147 in (ky,kx,c) order, Ho, Wo).  mean / std: 3 Python floats each (cfg.MODEL.PIXEL_MEAN / PIXEL_STD)."""
    img = _c(img, "img")
    N, Cc, H, W = img.shape
    assert Cc == 3
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = _empty_split(N * Ho * Wo, 192, img.device, fmt)
    out.K = 147
    m3, s3 = (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std])
    st = _lib.lib().nsac_stem_im2col_planes(_p(img), N, H, W, m3, s3, fmt, _p(out.hi), _p(out.lo), _stream())
    _lib.check(st, "nsac_stem_im2col_planes")
    _count()
    return out, Ho, Wo


def stem_im2col_u8(img: torch.Tensor, border_classes: bool = False):
    """[N,3,H,W] uint8 -> (Split with ONE fp16 plane [N*Ho*Wo, 192] of raw pixel values (K = 147 in (ky,kx,c) order, `lo` is
    None: exact), Ho, Wo).  Out-of-image taps are 0 — see stem_border_fix, or `border_classes=True`: 24 one-hot border-class
    columns follow (K = 171), see stem_border_classes / nsac_stem_im2col_u8_cls."""
    img = _c(img, "img", torch.uint8)
    N, Cc, H, W = img.shape
    assert Cc == 3
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    hi = torch.empty(N * Ho * Wo, 192, device=img.device, dtype=torch.float16)
    fn = _lib.lib().nsac_stem_im2col_u8_cls if border_classes else _lib.lib().nsac_stem_im2col_u8
    st = fn(_p(img), N, H, W, _p(hi), _stream())
    _lib.check(st, "nsac_stem_im2col_u8")
    _count()
    return Split(hi, None, 171 if border_classes else 147), Ho, Wo


def stem_border_classes(H: int, W: int):
    """The 24 border classes of nsac_stem_im2col_u8_cls for an H x W image: list of (idx, oob [7,7] bool) — which taps of the 7x7 /
    stride 2 / pad 3 window fall outside the image for output pixels of class idx."""
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    rep = lambda n: (0, 1, 2, n - 2, n - 1)                       # a representative output coordinate per class
    out = []
    for rc, yo in enumerate(rep(Ho)):
        for cc, xo in enumerate(rep(Wo)):
            cls = rc * 5 + cc
            if cls == 12:
                continue
            oob = torch.zeros(7, 7, dtype=torch.bool)
            for k in range(7):
                if not 0 <= 2 * yo + k - 3 < H:
                    oob[k, :] = True
                if not 0 <= 2 * xo + k - 3 < W:
                    oob[:, k] = True
            out.append((cls if cls < 12 else cls - 1, oob))
    return out


def stem_border_fix(img: torch.Tensor, w_folded: torch.Tensor, bias: torch.Tensor, mean, std, out: torch.Tensor):
    """Recomputes in exact fp32 the rows of the stem output `out` [N*Ho*Wo, 64] whose 7x7 window leaves the image (the zero
    padding applies to the normalised image); w_folded [64,147] fp32, (ky,kx,c) order, for normalised input."""
    img = _c(img, "img", torch.uint8)
    _chk(w_folded, "w_folded"); _chk(bias, "bias"); _chk(out, "out")
    N, _, H, W = img.shape
    assert w_folded.is_contiguous() and tuple(w_folded.shape) == (64, 147) and out.is_contiguous() and out.shape[1] == 64
    m3, s3 = (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std])
    st = _lib.lib().nsac_stem_border_fix(_p(img), _p(w_folded), _p(bias), N, H, W, m3, s3, _p(out), _stream())
    _lib.check(st, "nsac_stem_border_fix")
    _count()
    return out


def im2col3x3_from_planes(x: Split, N: int, H: int, W: int, stride: int):
    """3x3 / pad 1 im2col of contiguous NHWC planes [N*H*W, C] -> (Split [N*Ho*Wo, 9*C], Ho, Wo): 16-byte copies."""
    Cc = x.hi.shape[1]
    assert x.hi.is_contiguous() and x.lo.is_contiguous() and x.rows == N * H * W and Cc % 8 == 0 and (9 * Cc) % 64 == 0
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = Split(torch.empty(N * Ho * Wo, 9 * Cc, device=x.hi.device, dtype=x.hi.dtype),
                torch.empty(N * Ho * Wo, 9 * Cc, device=x.hi.device, dtype=x.hi.dtype), 9 * Cc, x.fmt, x.scale)
    st = _lib.lib().nsac_im2col3x3_from_planes(_p(x.hi), _p(x.lo), N, H, W, Cc, stride, _p(out.hi), _p(out.lo), _stream())
    _lib.check(st, "nsac_im2col3x3_from_planes")
    _count()
    return out, Ho, Wo


def maxpool3x3s2_nhwc(x: torch.Tensor, N: int, H: int, W: int, fmt: int = SPLIT_F16, want_f32: bool = True, want_split: bool = True):
    """MaxPool2d(3, 2, 1) of an fp32 NHWC map [N*H*W, C] -> (fp32 [N*Ho*Wo, C] or None, Split or None, Ho, Wo)."""
    x = _c(x, "x")
    Cc = x.shape[1]
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(N * Ho * Wo, Cc, device=x.device, dtype=torch.float32) if want_f32 else None
    sp = _empty_split(N * Ho * Wo, Cc, x.device, fmt) if want_split else None
    st = _lib.lib().nsac_maxpool3x3s2_nhwc(_p(x), N, H, W, Cc, fmt, _p(out), None if sp is None else _p(sp.hi),
                                           None if sp is None else _p(sp.lo), _stream())
    _lib.check(st, "nsac_maxpool3x3s2_nhwc")
    _count()
    return out, sp, Ho, Wo


def subsample2_planes(x: Split, N: int, H: int, W: int):
    """Every second pixel of contiguous NHWC planes [N*H*W, C] -> (Split [N*Ho*Wo, C], Ho, Wo)."""
    Cc = x.hi.shape[1]
    assert x.hi.is_contiguous() and x.lo.is_contiguous() and x.rows == N * H * W and Cc % 8 == 0
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = Split(torch.empty(N * Ho * Wo, Cc, device=x.hi.device, dtype=x.hi.dtype),
                torch.empty(N * Ho * Wo, Cc, device=x.hi.device, dtype=x.hi.dtype), x.K, x.fmt, x.scale)
    st = _lib.lib().nsac_subsample2_planes(_p(x.hi), _p(x.lo), N, H, W, Cc, _p(out.hi), _p(out.lo), _stream())
    _lib.check(st, "nsac_subsample2_planes")
    _count()
    return out, Ho, Wo


def add_relu_nhwc(a: torch.Tensor, b: torch.Tensor, fmt: int = SPLIT_F16, want_f32: bool = True, want_split: bool = True):
    """relu(a + b) of two contiguous fp32 [rows, C] maps (C % 64 == 0 for the planes) -> (fp32 or None, Split or None)."""
    a, b = _c(a, "a"), _c(b, "b")
    assert a.shape == b.shape and a.dim() == 2 and (not want_split or a.shape[1] % 64 == 0)
    out = torch.empty_like(a) if want_f32 else None
    sp = Split(torch.empty(a.shape, device=a.device, dtype=_SPLIT_DTYPE[fmt]), torch.empty(a.shape, device=a.device, dtype=_SPLIT_DTYPE[fmt]),
               a.shape[1], fmt) if want_split else None
    st = _lib.lib().nsac_add_relu_nhwc(_p(a), _p(b), a.numel(), fmt, _p(out), None if sp is None else _p(sp.hi),
                                       None if sp is None else _p(sp.lo), _stream())
    _lib.check(st, "nsac_add_relu_nhwc")
    _count()
    return out, sp


# ---------------------------------------------------------------------------------------------------
# PlaneTRHead glue (row f1; csrc/planetr.cu)
# ---------------------------------------------------------------------------------------------------
def row_op(x: torch.Tensor, y: Optional[torch.Tensor] = None, ln=None, pos: Optional[torch.Tensor] = None, T: int = 0,
           sum_out: Optional[torch.Tensor] = None, want_f32: bool = False, want_split: bool = False, want_pos_split: bool = False):
    """Per row of fp32 [rows, C]: s = x (+ y); t = LayerNorm(s) if `ln` = (gamma, beta, eps) else s.
    -> (t fp32 or None, t planes or None, (t + pos[row % T]) planes or None); `sum_out` (may alias x) receives s."""
    _chk(x, "x")
    rows, Cc = x.shape
    assert x.stride(1) == 1 and (y is None or (y.shape == x.shape and y.stride(1) == 1))
    dev = x.device
    t32 = torch.empty(rows, Cc, device=dev, dtype=torch.float32) if want_f32 else None
    tp = Split.empty(rows, Cc, dev) if want_split else None
    pp = Split.empty(rows, Cc, dev) if want_pos_split else None
    if want_pos_split:
        _chk(pos, "pos")
        assert pos.is_contiguous() and pos.shape == (T, Cc)
    g, b, eps = (ln[0], ln[1], float(ln[2])) if ln is not None else (None, None, 0.0)
    st = _lib.lib().nsac_row_op(_p(x), x.stride(0), _p(y), 0 if y is None else y.stride(0), _p(g), _p(b), eps, 0 if ln is None else 1,
                                _p(pos) if want_pos_split else None, T, _p(sum_out), 0 if sum_out is None else sum_out.stride(0),
                                _p(t32), Cc, None if tp is None else _p(tp.hi), None if tp is None else _p(tp.lo),
                                0 if tp is None else tp.hi.stride(0), None if pp is None else _p(pp.hi),
                                None if pp is None else _p(pp.lo), 0 if pp is None else pp.hi.stride(0), rows, Cc, _stream())
    _lib.check(st, "nsac_row_op")
    _count()
    return t32, tp, pp


def attention_tiled(q, k, v, B: int, L: int, S: int, H: int = 8, D: int = 32, want_f32: bool = False, out_split: Optional[Split] = None):
    """softmax(q k^T / sqrt(D)) v per (batch element, head); q [B*L, H*D], k / v [B*S, H*D] (same row stride), S <= 320.
    -> fp32 [B*L, H*D] (want_f32) and / or planes written into `out_split`."""
    _chk(q, "q"); _chk(k, "k"); _chk(v, "v")
    assert k.stride(0) == v.stride(0) and q.stride(1) == 1 and k.stride(1) == 1 and v.stride(1) == 1
    out = torch.empty(B * L, H * D, device=q.device, dtype=torch.float32) if want_f32 else None
    st = _lib.lib().nsac_attention_tiled(_p(q), q.stride(0), _p(k), _p(v), k.stride(0), _p(out), 0 if out is None else out.stride(0),
                                         None if out_split is None else _p(out_split.hi), None if out_split is None else _p(out_split.lo),
                                         0 if out_split is None else out_split.hi.stride(0), B, L, S, H, D, _stream())
    _lib.check(st, "nsac_attention_tiled")
    _count()
    return out if want_f32 else out_split


def upsample2x_relu_add(a: torch.Tensor, b: torch.Tensor, N: int, h: int, w: int, want_f32: bool = False, want_split: bool = True):
    """relu(bilinear_2x(a)) + b on NHWC fp32 rows: a [N*h*w, C], b [N*2h*2w, C] -> (fp32 or None, planes or None)."""
    a, b = _c(a, "a"), _c(b, "b")
    Cc = a.shape[1]
    assert a.shape[0] == N * h * w and b.shape == (N * 4 * h * w, Cc)
    out = torch.empty_like(b) if want_f32 else None
    sp = _empty_split(b.shape[0], Cc, b.device) if want_split else None
    st = _lib.lib().nsac_upsample2x_relu_add(_p(a), _p(b), N, h, w, Cc, _p(out), None if sp is None else _p(sp.hi),
                                             None if sp is None else _p(sp.lo), _stream())
    _lib.check(st, "nsac_upsample2x_relu_add")
    _count()
    return out, sp


def gemm_tc_rowbias(a: Split, w: Split, row_bias: torch.Tensor, out_f32: torch.Tensor, passes: int = 3):
    """out_f32[M, N] = a @ w^T + row_bias[m] (per-ROW scalar): the mask-logit GEMM whose rows are plane queries."""
    M, N = a.rows, w.rows
    K = (a.K + 63) // 64 * 64
    assert a.K == w.K and a.fmt == w.fmt and out_f32.shape == (M, N) and out_f32.stride(1) == 1 and row_bias.numel() == M
    st = _lib.lib().nsac_gemm_split_rowbias(_p(a.hi), _p(a.lo), a.hi.stride(0), _p(w.hi), _p(w.lo), w.hi.stride(0), None, _p(row_bias),
                                            M, N, K, ACT_NONE, passes, a.fmt, 1.0 / (a.scale * w.scale), _p(out_f32),
                                            out_f32.stride(0), None, None, 0, _stream())
    _lib.check(st, "nsac_gemm_split_rowbias")
    _count()
    return out_f32


def layernorm(x, gamma, beta, res=None, out=None, out_split: Optional[Split] = None):
    """out = (res or 0) + LayerNorm(x); optionally also written as fp16 hi/lo planes (`out_split` view)."""
    _chk(x, "x")
    rows, Cdim = x.shape
    if out is None:
        out = torch.empty(rows, Cdim, device=x.device, dtype=torch.float32)
    st = _lib.lib().nsac_layernorm(_p(x), x.stride(0), _p(gamma), _p(beta), _p(res),
                                   0 if res is None else res.stride(0), _p(out), out.stride(0),
                                   None if out_split is None else _p(out_split.hi),
                                   None if out_split is None else _p(out_split.lo),
                                   0 if out_split is None else out_split.hi.stride(0), rows, Cdim, _stream())
    _lib.check(st, "nsac_layernorm")
    _count()
    return out


def attention(q, k, v, B, L, S, H=8, D=32, out=None, out_split: Optional[Split] = None, want_f32: bool = True,
              kv_count: Optional[torch.Tensor] = None):
    """q [B*L, H*D] (row stride ldq), k/v [B*S, H*D] (same row stride) -> [B*L, H*D] fp32 and / or planes.
    `kv_count` (int32 [B]): ragged batch — only the first kv_count[b] source tokens of batch element b are attended to."""
    _chk(q, "q"); _chk(k, "k"); _chk(v, "v")
    assert k.stride(0) == v.stride(0)
    if out is None and want_f32:
        out = torch.empty(B * L, H * D, device=q.device, dtype=torch.float32)
    tail = (None if out_split is None else _p(out_split.hi), None if out_split is None else _p(out_split.lo),
            0 if out_split is None else out_split.hi.stride(0), B, L, S, H, D)
    head = (_p(q), q.stride(0), _p(k), _p(v), k.stride(0), _p(out), 0 if out is None else out.stride(0))
    if kv_count is None:
        st = _lib.lib().nsac_attention(*head, *tail, _stream())
    else:
        kv_count = _c(kv_count, "kv_count", torch.int32)
        assert kv_count.numel() == B
        st = _lib.lib().nsac_attention_ragged(*head, *tail, _p(kv_count), _stream())
    _lib.check(st, "nsac_attention")
    _count()
    return out if out is not None else out_split


def match_sinkhorn_assign(desc1, desc2, planes1, planes2, cam, bin_score, offset_mult=4.0, normal_mult=8.0,
                          iters=200, threshold=0.2, count1: Optional[torch.Tensor] = None,
                          count2: Optional[torch.Tensor] = None):
    """-> (log_scores_padded [B,n1+1,n2+1], assign [B,n1,n2]).  `count1` / `count2` (int32 [B]): ragged batch — pair b has
    count1[b] x count2[b] planes in the first rows of the padded inputs; its result is the top-left block (rest -inf / 0)."""
    desc1, desc2 = _c(desc1, "desc1"), _c(desc2, "desc2")
    planes1, planes2, cam = _c(planes1, "planes1"), _c(planes2, "planes2"), _c(cam, "cam")
    bin_score = _c(bin_score.reshape(1), "bin_score")
    B, n1, Cd = desc1.shape
    n2 = desc2.shape[1]
    lsp = torch.empty(B, n1 + 1, n2 + 1, device=desc1.device, dtype=torch.float32)
    assign = torch.empty(B, n1, n2, device=desc1.device, dtype=torch.float32)
    args = (_p(desc1), _p(desc2), _p(planes1), _p(planes2), _p(cam), _p(bin_score), offset_mult, normal_mult, iters, threshold,
            B, n1, n2, Cd)
    if count1 is None and count2 is None:
        st = _lib.lib().nsac_match_sinkhorn_assign(*args, _p(lsp), _p(assign), _stream())
    else:
        count1 = None if count1 is None else _c(count1, "count1", torch.int32)
        count2 = None if count2 is None else _c(count2, "count2", torch.int32)
        assert (count1 is None or count1.numel() == B) and (count2 is None or count2.numel() == B)
        st = _lib.lib().nsac_match_sinkhorn_assign_ragged(*args, _p(count1), _p(count2), _p(lsp), _p(assign), _stream())
    _lib.check(st, "nsac_match_sinkhorn_assign")
    _count()
    return lsp, assign


def geo_sequence(planes1, planes2, assign, t0, q0, num_queries, hyp_pairs=None):
    planes1, planes2, t0, q0 = _c(planes1, "planes1"), _c(planes2, "planes2"), _c(t0, "t0"), _c(q0, "q0")
    B, n1, _ = planes1.shape
    n2 = planes2.shape[1]
    dev = planes1.device
    NQ = num_queries
    H = 0
    if hyp_pairs is not None:
        hyp_pairs = _c(hyp_pairs, "hyp_pairs", torch.int32)
        H = hyp_pairs.shape[0]
    else:
        assign = _c(assign, "assign")
    geo_local = torch.empty(B, NQ, 6, device=dev)
    geo_global = torch.empty(B, NQ, 6, device=dev)
    sig = torch.empty(B, NQ, device=dev)
    geo8 = torch.empty(B, NQ, 8, device=dev)
    mnum = torch.empty(B, device=dev, dtype=torch.int32)
    pidx = torch.empty(B, NQ, 2, device=dev, dtype=torch.int32)
    st = _lib.lib().nsac_geo_sequence(_p(planes1), _p(planes2), _p(assign) if hyp_pairs is None else None,
                                      _p(hyp_pairs), H, _p(t0), _p(q0), B, n1, n2, NQ, _p(geo_local), _p(geo_global),
                                      _p(sig), _p(geo8), _p(mnum), _p(pidx), _stream())
    _lib.check(st, "nsac_geo_sequence")
    _count()
    return geo_local, geo_global, sig, geo8, mnum, pidx


def pose_heads(feat_rot, feat_tran, w_rots, b_rots, w_trans, b_trans):
    rows = (feat_rot if feat_rot is not None else feat_tran).shape[0]
    dev = (feat_rot if feat_rot is not None else feat_tran).device
    q = torch.empty(rows, 4, device=dev) if feat_rot is not None else None
    t = torch.empty(rows, 3, device=dev) if feat_tran is not None else None
    fr = None if feat_rot is None else _c(feat_rot, "feat_rot")
    ft = None if feat_tran is None else _c(feat_tran, "feat_tran")
    st = _lib.lib().nsac_pose_heads(_p(fr), _p(ft), _p(w_rots), _p(b_rots), _p(w_trans), _p(b_trans), rows, 256,
                                    _p(q), _p(t), _stream())
    _lib.check(st, "nsac_pose_heads")
    _count()
    return q, t


def _score_mlp_struct(ws):
    s = _lib.ScoreMLP()
    for name, t in zip(("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4"), ws):
        _chk(t, name)
        assert t.is_contiguous()
        setattr(s, name, t.data_ptr())
    return s


def score_pack(rot_mlp, tran_mlp, num_queries: int) -> torch.Tensor:
    """Device-side weight pack of the tensor-core scoring path (build once per weight version)."""
    L = _lib.lib()
    dev = rot_mlp[0].device
    pack = torch.empty(L.nsac_score_pack_bytes(num_queries), device=dev, dtype=torch.uint8)
    rs, ts = _score_mlp_struct(rot_mlp), _score_mlp_struct(tran_mlp)
    st = L.nsac_score_pack(C.byref(rs), C.byref(ts), num_queries, _p(pack), _stream())
    _lib.check(st, "nsac_score_pack")
    _count()
    return pack


def score_pack_host_vectors(pack: torch.Tensor, num_queries: int) -> Optional[torch.Tensor]:
    """Host (pinned) mirror of the pack's bias / folded-output vectors (6 x 128 floats): passed to score_aggregate as
    `vecs_host`, they travel as kernel parameters (constant bank) instead of shared-memory broadcast loads.  One-off device ->
    host copy per weight version (synchronises), next to split_weight's own one-off range probe."""
    off = _lib.lib().nsac_score_pack_vecs_offset(num_queries)
    if off == 0:
        return None
    v = pack[off:off + 6 * 128 * 4].view(torch.float32).cpu().contiguous()
    return v.pin_memory() if pack.is_cuda else v


def score_aggregate(geo_local, q_h, t_h, q0, t0, feat_rot, feat_tran, feat_rot0, feat_tran0, matched_num,
                    rot_mlp, tran_mlp, w_rots, b_rots, w_trans, b_trans, out_cam_type="soft",
                    want_scores=True, want_diag=False, precision="fp16", pack=None, exchange=None, vecs_host=None):
    """rot_mlp / tran_mlp: 8-tuples (w1,b1,w2,b2,w3,b3,w4,b4). Returns dict(pose, score_rot, score_tran,
    sel_idx, diag).  `exchange` (nopesac_b200.dist.FusedResultExchange) makes the selection kernel also store
    every result row into all ranks' result buffers over NVLink.  precision "fp16" = tcgen05 path (score MLPs single-pass fp16, fp32 accumulate);
    "fp32" = exact CUDA-core path (always used when the diagnostic outputs are requested)."""
    B, NQ, _ = geo_local.shape
    dev = geo_local.device
    args = [_c(a, n) for a, n in ((geo_local, "geo_local"), (q_h, "q_h"), (t_h, "t_h"), (q0, "q0"), (t0, "t0"),
                                  (feat_rot, "feat_rot"), (feat_tran, "feat_tran"), (feat_rot0, "feat_rot0"),
                                  (feat_tran0, "feat_tran0"))]
    matched_num = _c(matched_num, "matched_num", torch.int32)
    pose = torch.empty(B, 16, device=dev)
    sr = torch.empty(B, NQ + 1, device=dev) if want_scores else None
    stt = torch.empty(B, NQ + 1, device=dev) if want_scores else None
    sel = torch.empty(B, 2, device=dev, dtype=torch.int32)
    diag = torch.zeros(3, B, NQ + 1, NQ, device=dev) if want_diag else None
    L = _lib.lib()
    rs, ts = _score_mlp_struct(rot_mlp), _score_mlp_struct(tran_mlp)
    if precision == "fp16" and not want_diag:
        if pack is None:
            pack = score_pack(rot_mlp, tran_mlp, NQ)
        ws = torch.empty(L.nsac_score_tc_workspace_bytes(B, NQ), device=dev, dtype=torch.uint8)
        if os.environ.get("NSAC_SCORE_NO_CVEC"):                                   # A/B knob (scripts/score_quick.py)
            vecs_host = None
        vh = None if vecs_host is None else C.c_void_p(vecs_host.data_ptr())       # HOST pointer (see score_pack_host_vectors)
        st = L.nsac_score_aggregate_tc_cv(*[_p(a) for a in args], _p(matched_num), _p(pack), vh,
                                       _p(w_rots), _p(b_rots), _p(w_trans), _p(b_trans), B, NQ, CAM_TYPES[out_cam_type],
                                       _p(pose), _p(sr), _p(stt), _p(sel), _p(ws),
                                       None if exchange is None else C.c_void_p(exchange.peer_ptrs_dev),
                                       0 if exchange is None else exchange.world, 0 if exchange is None else exchange.row_offset,
                                       _stream())
        _lib.check(st, "nsac_score_aggregate_tc")
        _count(3)
        return {"pose": pose, "score_rot": sr, "score_tran": stt, "sel_idx": sel, "diag": None}
    if exchange is not None:
        raise RuntimeError("the fused result exchange needs the tensor-core scoring path (precision='fp16', no diagnostics)")
    ws = torch.empty(L.nsac_score_workspace_bytes(B, NQ), device=dev, dtype=torch.uint8)
    st = L.nsac_score_aggregate(*[_p(a) for a in args], _p(matched_num), C.byref(rs), C.byref(ts),
                                _p(w_rots), _p(b_rots), _p(w_trans), _p(b_trans), B, NQ, CAM_TYPES[out_cam_type],
                                _p(pose), _p(sr), _p(stt), _p(sel), _p(diag), _p(ws), _stream())
    _lib.check(st, "nsac_score_aggregate")
    _count(3)
    return {"pose": pose, "score_rot": sr, "score_tran": stt, "sel_idx": sel, "diag": diag}


def prune_assignment(assign, planes1, planes2, pose):
    assign, planes1, planes2 = _c(assign, "assign"), _c(planes1, "planes1"), _c(planes2, "planes2")
    _chk(pose, "pose")
    assert pose.stride(1) == 1
    B, n1, n2 = assign.shape
    out = torch.empty_like(assign)
    st = _lib.lib().nsac_prune_assignment(_p(assign), _p(planes1), _p(planes2), _p(pose), pose.stride(0), B, n1, n2,
                                          _p(out), _stream())
    _lib.check(st, "nsac_prune_assignment")
    _count()
    return out


# ---------------------------------------------------------------------------------------------------
# whole-stage entry: the one-plane RANSAC refinement (K6 .. K10) behind ONE C call (csrc/forward.cu)
# ---------------------------------------------------------------------------------------------------
def tc_layer(w: Split, bias: Optional[torch.Tensor]) -> "_lib.TcLayer":
    """nsac_tc_layer of a split weight (borrowed pointers: the caller keeps `w` / `bias` alive)."""
    t = _lib.TcLayer()
    t.w_hi, t.w_lo = w.hi.data_ptr(), w.lo.data_ptr()
    t.bias = None if bias is None else bias.data_ptr()
    t.N, t.K, t.ldw, t.w_scale = w.rows, (w.K + 63) // 64 * 64, w.hi.stride(0), float(w.scale)
    return t


def refine_forward(weights, planes1, planes2, assign, t0, q0, rot_feat0, trans_feat0, num_queries: int, out_cam_type: str = "soft",
                   hyp_pairs=None, want_scores: bool = True, prune: bool = True, exchange=None):
    """nsac_refine_forward: geo sequences -> hypothesis MLP chain -> per-hypothesis poses -> scoring / selection -> pruning, one
    call, no host round trip.  `weights`: a _lib.RefineWeights (PlaneCameraHead.refine_weights()).  Returns a dict with the
    tensors of the stage (pose [B,16], assign_pruned, geo_local, geo_global, sig, matched_num, pair_idx, q_h, t_h, score_rot,
    score_tran, sel_idx)."""
    planes1, planes2, t0, q0 = _c(planes1, "planes1"), _c(planes2, "planes2"), _c(t0, "t0"), _c(q0, "q0")
    rot_feat0, trans_feat0 = _c(rot_feat0, "rot_feat0"), _c(trans_feat0, "trans_feat0")
    B, n1, _ = planes1.shape
    n2, NQ, dev = planes2.shape[1], num_queries, planes1.device
    H = 0
    if hyp_pairs is not None:
        hyp_pairs = _c(hyp_pairs, "hyp_pairs", torch.int32)
        H = hyp_pairs.shape[0]
    if assign is not None:
        assign = _c(assign, "assign")
    f = lambda *shape: torch.empty(*shape, device=dev)
    i32 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.int32)
    out = {"pose": f(B, 16), "assign_pruned": torch.empty_like(assign) if (prune and assign is not None) else None,
           "geo_local": f(B, NQ, 6), "geo_global": f(B, NQ, 6), "sig": f(B, NQ), "matched_num": i32(B), "pair_idx": i32(B, NQ, 2),
           "q_h": f(B * NQ, 4), "t_h": f(B * NQ, 3), "score_rot": f(B, NQ + 1) if want_scores else None,
           "score_tran": f(B, NQ + 1) if want_scores else None, "sel_idx": i32(B, 2)}
    L = _lib.lib()
    nbytes = L.nsac_refine_workspace_bytes(B, NQ)
    ws, ws_ptr = _aligned_workspace(nbytes, dev)
    n = C.c_int(0)
    st = L.nsac_refine_forward(C.byref(weights), _p(planes1), _p(planes2), _p(assign), _p(hyp_pairs), H, _p(t0), _p(q0), _p(rot_feat0),
                               _p(trans_feat0), B, n1, n2, NQ, CAM_TYPES[out_cam_type],
                               *[_p(out[k]) for k in ("pose", "assign_pruned", "geo_local", "geo_global", "sig", "matched_num", "pair_idx",
                                                      "q_h", "t_h", "score_rot", "score_tran", "sel_idx")],
                               C.c_void_p(ws_ptr), nbytes, None if exchange is None else C.c_void_p(exchange.peer_ptrs_dev),
                               0 if exchange is None else exchange.world, 0 if exchange is None else exchange.row_offset,
                               C.byref(n), _stream())
    _lib.check(st, "nsac_refine_forward")
    _count(n.value)
    return out


def _aligned_workspace(nbytes: int, device):
    ws = torch.empty(nbytes + 256, device=device, dtype=torch.uint8)
    return ws, (ws.data_ptr() + 255) // 256 * 256                    # 256-byte aligned (CUDA allocations already are)


def match_forward(weights, app1, app2, planes1, planes2, cam, threshold: float, count1=None, count2=None):
    """nsac_match_forward: the whole MatchingHead forward (projection -> 18 GNN layers -> descriptors -> Sinkhorn + mutual-NN
    assignment) in one call.  `weights`: a _lib.MatchWeights (MatchingHead.match_weights()).
    -> (log_scores_padded [B,n1+1,n2+1], assign [B,n1,n2])."""
    app1, app2 = _c(app1, "app1"), _c(app2, "app2")
    planes1, planes2, cam = _c(planes1, "planes1"), _c(planes2, "planes2"), _c(cam, "cam")
    B, n1, Cd = app1.shape
    n2, dev = app2.shape[1], app1.device
    assert Cd == 256 and app2.shape[2] == 256
    if count1 is not None:
        count1, count2 = _c(count1, "count1", torch.int32), _c(count2, "count2", torch.int32)
        assert count1.numel() == B and count2.numel() == B
    lsp = torch.empty(B, n1 + 1, n2 + 1, device=dev)
    assign = torch.empty(B, n1, n2, device=dev)
    L = _lib.lib()
    nbytes = L.nsac_match_workspace_bytes(B, n1, n2)
    ws, ws_ptr = _aligned_workspace(nbytes, dev)
    n = C.c_int(0)
    st = L.nsac_match_forward(C.byref(weights), _p(app1), _p(app2), _p(planes1), _p(planes2), _p(cam), _p(count1), _p(count2),
                              float(threshold), B, n1, n2, _p(lsp), _p(assign), C.c_void_p(ws_ptr), nbytes, C.byref(n), _stream())
    _lib.check(st, "nsac_match_forward")
    _count(n.value)
    return lsp, assign


def pixel_forward(weights, res3, res4, res5, B: int, H3: int, W3: int, initial_pose=None):
    """nsac_pixel_forward: K1 (pixel pose network on the stacked views' res3 / res4 / res5 planes) + w >= 0 canonicalisation +
    K2 (AIM), one call.  `res*`: ops.Split NHWC planes of the 2B images, or all None with `initial_pose = (tran [B,3], rot [B,4])`
    (the network is skipped).  -> dict(init_tran, init_rot, pix_tran_feat, pix_rot_feat, t0, q0, rot_feat0, trans_feat0)."""
    L = _lib.lib()
    if res5 is None:
        t_in, q_in = initial_pose
        dev = t_in.device
        out = {"init_tran": _c(t_in, "initial tran").clone(), "init_rot": _c(q_in, "initial rot").clone(),
               "pix_tran_feat": None, "pix_rot_feat": None}
        assert out["init_tran"].shape == (B, 3) and out["init_rot"].shape == (B, 4)
        planes = [None] * 6
        H3 = W3 = 4
    else:
        dev = res5.hi.device
        for r, c, hw in ((res3, 512, H3 * W3), (res4, 1024, H3 * W3 // 4), (res5, 2048, H3 * W3 // 16)):
            assert r.hi.is_contiguous() and r.lo.is_contiguous() and tuple(r.hi.shape) == (2 * B * hw, c) and r.scale == 1.0, \
                f"pixel_forward: expected contiguous planes [{2 * B * hw}, {c}], got {tuple(r.hi.shape)}"
        out = {"init_tran": torch.empty(B, 3, device=dev), "init_rot": torch.empty(B, 4, device=dev),
               "pix_tran_feat": torch.empty(B, 256, device=dev), "pix_rot_feat": torch.empty(B, 256, device=dev)}
        planes = [_p(res3.hi), _p(res3.lo), _p(res4.hi), _p(res4.lo), _p(res5.hi), _p(res5.lo)]
    out.update(t0=torch.empty(B, 3, device=dev), q0=torch.empty(B, 4, device=dev), rot_feat0=torch.empty(B, 256, device=dev),
               trans_feat0=torch.empty(B, 256, device=dev))
    nbytes = L.nsac_pixel_workspace_bytes(B, H3, W3)
    ws, ws_ptr = _aligned_workspace(nbytes, dev)
    n = C.c_int(0)
    st = L.nsac_pixel_forward(C.byref(weights), *planes, B, H3, W3,
                              *[_p(out[k]) for k in ("init_tran", "init_rot", "pix_tran_feat", "pix_rot_feat", "t0", "q0", "rot_feat0",
                                                     "trans_feat0")], C.c_void_p(ws_ptr), nbytes, C.byref(n), _stream())
    _lib.check(st, "nsac_pixel_forward")
    _count(n.value)
    return out


def head_forward(pixel_w, match_w, refine_w, res3, res4, res5, B: int, H3: int, W3: int, planes1, planes2, app1, app2,
                 num_queries: int, match_threshold: float, out_cam_type: str = "soft", count1=None, count2=None, hyp_pairs=None,
                 initial_pose=None, want_scores: bool = True, exchange=None):
    """nsac_head_forward: PlaneCameraHead.inference_Joint + MatchingHead in ONE C call (pixel pose network + AIM -> matcher ->
    one-plane RANSAC refinement).  Arguments as in pixel_forward / match_forward / refine_forward; returns one dict with every
    output tensor of the three stages."""
    planes1, planes2, app1, app2 = _c(planes1, "planes1"), _c(planes2, "planes2"), _c(app1, "app1"), _c(app2, "app2")
    n1, n2, NQ, dev = planes1.shape[1], planes2.shape[1], num_queries, planes1.device
    assert planes1.shape[0] == B and tuple(app1.shape) == (B, n1, 256) and tuple(app2.shape) == (B, n2, 256)
    f = lambda *shape: torch.empty(*shape, device=dev)
    i32 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.int32)
    if res5 is None:
        t_in, q_in = initial_pose
        init_tran, init_rot = _c(t_in, "initial tran").clone(), _c(q_in, "initial rot").clone()
        assert init_tran.shape == (B, 3) and init_rot.shape == (B, 4)
        planes = [None] * 6
        H3 = W3 = 0
    else:
        for r, c, hw in ((res3, 512, H3 * W3), (res4, 1024, H3 * W3 // 4), (res5, 2048, H3 * W3 // 16)):
            assert r.hi.is_contiguous() and r.lo.is_contiguous() and tuple(r.hi.shape) == (2 * B * hw, c) and r.scale == 1.0, \
                f"head_forward: expected contiguous planes [{2 * B * hw}, {c}], got {tuple(r.hi.shape)}"
        init_tran, init_rot = f(B, 3), f(B, 4)
        planes = [_p(res3.hi), _p(res3.lo), _p(res4.hi), _p(res4.lo), _p(res5.hi), _p(res5.lo)]
    Hn = 0
    if hyp_pairs is not None:
        hyp_pairs = _c(hyp_pairs, "hyp_pairs", torch.int32)
        Hn = hyp_pairs.shape[0]
    if count1 is not None:
        count1, count2 = _c(count1, "count1", torch.int32), _c(count2, "count2", torch.int32)
    out = {"init_tran": init_tran, "init_rot": init_rot, "t0": f(B, 3), "q0": f(B, 4), "rot_feat0": f(B, 256), "trans_feat0": f(B, 256),
           "log_scores_padded": f(B, n1 + 1, n2 + 1), "assign": f(B, n1, n2), "pose": f(B, 16), "assign_pruned": f(B, n1, n2),
           "geo_local": f(B, NQ, 6), "geo_global": f(B, NQ, 6), "sig": f(B, NQ), "matched_num": i32(B), "pair_idx": i32(B, NQ, 2),
           "q_h": f(B * NQ, 4), "t_h": f(B * NQ, 3), "score_rot": f(B, NQ + 1) if want_scores else None,
           "score_tran": f(B, NQ + 1) if want_scores else None, "sel_idx": i32(B, 2)}
    L = _lib.lib()
    W = _lib.HeadWeights(C.pointer(pixel_w), C.pointer(match_w), C.pointer(refine_w))
    nbytes = L.nsac_head_workspace_bytes(B, H3, W3, n1, n2, NQ)
    ws, ws_ptr = _aligned_workspace(nbytes, dev)
    n = C.c_int(0)
    st = L.nsac_head_forward(C.byref(W), *planes, B, H3, W3, _p(planes1), _p(planes2), _p(app1), _p(app2), _p(count1), _p(count2), n1, n2,
                             _p(hyp_pairs), Hn, NQ, float(match_threshold), CAM_TYPES[out_cam_type], *[_p(v) for v in out.values()],
                             C.c_void_p(ws_ptr), nbytes, None if exchange is None else C.c_void_p(exchange.peer_ptrs_dev),
                             0 if exchange is None else exchange.world, 0 if exchange is None else exchange.row_offset,
                             C.byref(n), _stream())
    _lib.check(st, "nsac_head_forward")
    _count(n.value)
    return out


def backbone_forward(weights, images: torch.Tensor, keep=("res2", "res3", "res4", "res5")):
    """nsac_backbone_forward: ResNet-50 from uint8 images [N,3,H,W] (not normalised) in one call -> {name: (Split NHWC planes, h, w)}
    for the levels in `keep`.  `weights`: a _lib.BackboneWeights for this image size (ResNet50Backbone.backbone_weights(H, W))."""
    images = _c(images, "images", torch.uint8)
    N, Cc, H, W = images.shape
    assert Cc == 3
    dev = images.device
    up = lambda v: (v - 1) // 2 + 1
    h, w = up(up(H)), up(up(W))
    out, ptrs = {}, []
    for name, ch in (("res2", 256), ("res3", 512), ("res4", 1024), ("res5", 2048)):
        if name in keep:
            sp = Split(torch.empty(N * h * w, ch, device=dev, dtype=torch.float16), torch.empty(N * h * w, ch, device=dev, dtype=torch.float16), ch)
            out[name] = (sp, h, w)
            ptrs += [_p(sp.hi), _p(sp.lo)]
        else:
            ptrs += [None, None]
        h, w = up(h), up(w)
    L = _lib.lib()
    nbytes = L.nsac_backbone_workspace_bytes(N, H, W)
    ws, ws_ptr = _aligned_workspace(nbytes, dev)
    n = C.c_int(0)
    st = L.nsac_backbone_forward(C.byref(weights), _p(images), N, H, W, *ptrs, C.c_void_p(ws_ptr), nbytes, C.byref(n), _stream())
    _lib.check(st, "nsac_backbone_forward")
    _count(n.value)
    return out


def model_forward(backbone_w, pixel_w, match_w, refine_w, images: torch.Tensor, planes1, planes2, app1, app2, num_queries: int,
                  match_threshold: float, out_cam_type: str = "soft", count1=None, count2=None, hyp_pairs=None, want_scores: bool = True,
                  exchange=None):
    """nsac_model_forward: stage set S5 in ONE C call — ResNet-50 on the 2B uint8 images (first views, then second views; H, W
    multiples of 32) -> pixel pose network + AIM -> matcher -> one-plane RANSAC refinement.  Returns the dict of head_forward."""
    images = _c(images, "images", torch.uint8)
    planes1, planes2, app1, app2 = _c(planes1, "planes1"), _c(planes2, "planes2"), _c(app1, "app1"), _c(app2, "app2")
    B, n1, n2, NQ, dev = planes1.shape[0], planes1.shape[1], planes2.shape[1], num_queries, planes1.device
    N2, Cc, H, W = images.shape
    assert N2 == 2 * B and Cc == 3 and H % 32 == 0 and W % 32 == 0, f"model_forward: images {tuple(images.shape)} for {B} pairs"
    assert tuple(app1.shape) == (B, n1, 256) and tuple(app2.shape) == (B, n2, 256)
    f = lambda *shape: torch.empty(*shape, device=dev)
    i32 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.int32)
    Hn = 0
    if hyp_pairs is not None:
        hyp_pairs = _c(hyp_pairs, "hyp_pairs", torch.int32)
        Hn = hyp_pairs.shape[0]
    if count1 is not None:
        count1, count2 = _c(count1, "count1", torch.int32), _c(count2, "count2", torch.int32)
    out = {"init_tran": f(B, 3), "init_rot": f(B, 4), "t0": f(B, 3), "q0": f(B, 4), "rot_feat0": f(B, 256), "trans_feat0": f(B, 256),
           "log_scores_padded": f(B, n1 + 1, n2 + 1), "assign": f(B, n1, n2), "pose": f(B, 16), "assign_pruned": f(B, n1, n2),
           "geo_local": f(B, NQ, 6), "geo_global": f(B, NQ, 6), "sig": f(B, NQ), "matched_num": i32(B), "pair_idx": i32(B, NQ, 2),
           "q_h": f(B * NQ, 4), "t_h": f(B * NQ, 3), "score_rot": f(B, NQ + 1) if want_scores else None,
           "score_tran": f(B, NQ + 1) if want_scores else None, "sel_idx": i32(B, 2)}
    L = _lib.lib()
    HW_ = _lib.HeadWeights(C.pointer(pixel_w), C.pointer(match_w), C.pointer(refine_w))
    nbytes = L.nsac_model_workspace_bytes(B, H, W, n1, n2, NQ)
    ws, ws_ptr = _aligned_workspace(nbytes, dev)
    n = C.c_int(0)
    st = L.nsac_model_forward(C.byref(backbone_w), C.byref(HW_), _p(images), B, H, W, _p(planes1), _p(planes2), _p(app1), _p(app2),
                              _p(count1), _p(count2), n1, n2, _p(hyp_pairs), Hn, NQ, float(match_threshold), CAM_TYPES[out_cam_type],
                              *[_p(v) for v in out.values()], C.c_void_p(ws_ptr), nbytes,
                              None if exchange is None else C.c_void_p(exchange.peer_ptrs_dev), 0 if exchange is None else exchange.world,
                              0 if exchange is None else exchange.row_offset, C.byref(n), _stream())
    _lib.check(st, "nsac_model_forward")
    _count(n.value)
    return out
