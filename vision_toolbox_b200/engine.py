"""Graph planner + executor behind the drop-in modules.

A module tree (ConvNormAct, DarknetBlock, CSPDarknetStage, OSABlock, Darknet, VoVNet ...) *emits* itself into
a :class:`Graph` once per (input shape, mode, precision).  The graph is a static list of ops over NHWC views
(bf16, or fp32 in parity mode) that live in one activation arena (concatenations are channel slices of one buffer, so
``torch.cat`` of reference darknet.py:53 / vovnet.py:55 never runs).  :class:`Runner` replays that list through the C ABI
(``include/vtb.h``) on the current CUDA stream, forward and backward; autograd sees ONE node per call.

Plan-level transformations (all decided when the graph is built, all covered by tests/test_precision.py and the CPU dry
run of tests/test_dry_run_plan.py): sibling units that read the same tensor run as one convolution
(``conv_norm_act_pair``), 3x3 RGB stems run as a 1x1 GEMM over a gathered operand, residual gradients alias the block
output's gradient memory, weight-gradient GEMMs go to a second stream in single-process plans, every convolution's bf16
operand packs are rebuilt by one launch per forward.

No op here has a torch/cuDNN implementation: if libvtb_b200.so is missing, loading it raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Optional

import torch
from torch import nn

from . import _lib
from ._lib import VtbBnTrain, VtbConv, VtbDgradBn, check

BF16 = 2
F32 = 4

# ----------------------------------------------------------------------------------------------------
# precision mode
#   "bf16" (default): NHWC bf16 activations, tcgen05 implicit-GEMM kernels - the reference under torch.autocast(bf16)
#   "fp32"          : NHWC fp32 activations, CUDA-core FMA kernels of csrc/parity_f32.cu - the reference WITHOUT autocast,
#                     the north star's "fp32 mode" (1e-4 relative); a parity instrument, not a fast path
#   "auto"          : what the reference would do with the same call: bf16 inside torch.autocast(dtype=bfloat16), else fp32
# ----------------------------------------------------------------------------------------------------
import contextlib as _contextlib
import os as _os

_PRECISIONS = ("bf16", "fp32", "auto")
_precision = _os.environ.get("VTB_PRECISION", "bf16")
if _precision not in _PRECISIONS:
    raise ValueError(f"VTB_PRECISION must be one of {_PRECISIONS}, got {_precision!r}")


def set_precision(mode: str) -> None:
    global _precision
    if mode not in _PRECISIONS:
        raise KeyError(f"unknown precision {mode!r}; expected one of {_PRECISIONS}")
    _precision = mode


def get_precision() -> str:
    return _precision


@_contextlib.contextmanager
def precision(mode: str):
    """``with precision("fp32"): model(x)`` - run CUDA tensors through the fp32 parity kernels."""
    old = _precision
    set_precision(mode)
    try:
        yield
    finally:
        set_precision(old)


def _resolve_f32() -> bool:
    if _precision == "auto":
        return not (torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16)
    return _precision == "fp32"


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def _check_norm_tensors(mod) -> None:
    """The kernels read and write BatchNorm parameters / statistics as fp32 arrays through raw pointers: anything else
    (model.half(), model.bfloat16(), a non-contiguous view) must fail loudly instead of corrupting memory."""
    norm = mod.norm
    for name in ("weight", "bias", "running_mean", "running_var"):
        t = getattr(norm, name, None)
        if t is None:
            if name in ("weight", "bias"):
                raise NotImplementedError("BatchNorm2d(affine=False) has no sm_100a kernel here")
            continue
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError(f"BatchNorm2d.{name} must be a contiguous float32 tensor for the native path (got {t.dtype}); "
                            "keep normalisation parameters and buffers in fp32 (model.half()/bfloat16() are not supported)")


def _bn_signature(module: nn.Module) -> tuple:
    """Per-BatchNorm (batch statistics?, parameter dtype) flags of a module tree: part of the plan key, so toggling
    sub-module modes (bn.eval(), stem.eval()) or casting the model builds a new plan instead of reusing a stale one."""
    sig = []
    for m in module.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            sig.append((bool(m.training or m.running_mean is None), m.weight.dtype if m.weight is not None else None))
    return tuple(sig)


# ----------------------------------------------------------------------------------------------------
# symbolic tensors
# ----------------------------------------------------------------------------------------------------
class Buffer:
    """A dense NHWC bf16 allocation of `pixels` x `c` elements inside the activation (or gradient) arena."""

    def __init__(self, idx: int, n: int, h: int, w: int, c: int, kind: str, esize: int = BF16):
        self.idx, self.n, self.h, self.w, self.c, self.kind = idx, n, h, w, c, kind
        self.esize = esize  # bytes per element: 2 (bf16 mode) or 4 (fp32 parity mode)
        self.offset = -1  # bytes, assigned by Graph.finalize
        self.exclusive_owner: Optional["TView"] = None

    @property
    def pixels(self) -> int:
        return self.n * self.h * self.w

    @property
    def nbytes(self) -> int:
        return self.pixels * self.c * self.esize


class TView:
    """Channel slice [coff, coff+c) of a Buffer: what every kernel sees as (pointer, pixel pitch)."""

    def __init__(self, buf: Buffer, coff: int, c: int):
        self.buf, self.coff, self.c = buf, coff, c
        self.consumers: list[int] = []
        self.is_output = False
        self.is_input = False
        self.grad_alias: Optional["TView"] = None  # gradient lives in another view's gradient memory

    n = property(lambda s: s.buf.n)
    h = property(lambda s: s.buf.h)
    w = property(lambda s: s.buf.w)
    ld = property(lambda s: s.buf.c)
    pixels = property(lambda s: s.buf.pixels)

    def byte_offset(self) -> int:
        return self.buf.offset + self.coff * self.buf.esize

    def __repr__(self):
        return f"TView(buf{self.buf.idx}[{self.coff}:{self.coff + self.c}] {self.n}x{self.h}x{self.w} ld{self.ld})"


@dataclass
class ConvOp:
    mod: Any  # ConvNormAct
    x: TView
    y: Optional[TView]  # raw conv output (None in fused eval mode)
    out: TView
    residual: Optional[TView]
    relu: bool
    geom: VtbConv
    cin_real: int
    pidx: int = -1  # index of (weight, gamma, beta) in the parameter list
    st: dict = field(default_factory=dict)  # float offsets into the stat arena
    kind: str = "conv"
    # side-by-side pair (two units reading the same tensor, run as ONE convolution): set on the first / second op
    pair: Any = None        # first op -> its partner
    pair_of: Any = None     # second op -> the first
    pair_geom: Any = None   # VtbConv with cout = both units (first op only)
    # gathered-operand stem: (taps, image channels) when this op is the image's first convolution run as a 1x1 GEMM
    col: Any = None
    # BatchNorm2d decides per LAYER whether it normalises with batch statistics (torch: `self.training or running stats
    # are None`) - the frozen-BN fine-tuning pattern is model.train() followed by bn.eval() on some layers
    batch_stats: bool = True
    # plain nn.Conv2d WITH bias and no normalisation (the lateral 1x1 convolutions of the reference necks, necks.py:60-65):
    # the bias rides in the fused affine epilogue (scale = 1, shift = bias); parameters are (weight, bias)
    bias: bool = False
    nparams: int = 3
    # BatchNorm-backward statistics carried by this op's dgrad (Graph._plan_dgrad_bn): ([producer ConvOp, ...], split) when
    # the dgrad into x is the LAST contribution to the gradient of the tensor(s) those producers wrote
    dgrad_bn: Any = None
    bwd_stats_from: Any = None   # producer side: the consumer op whose dgrad epilogue reduces this layer's (dz, dz*xhat) sums


@dataclass
class PoolOp:
    x: TView
    out: TView
    idx: Optional[TView] = None   # saved argmax positions (uint8 per element), only when gradients are needed
    kind: str = "pool"


@dataclass
class CopyOp:
    """out = x: places a tensor the plan does not own (the plan input) into a channel slice of a concat buffer."""
    x: TView
    out: TView
    kind: str = "copy"


@dataclass
class EseOp:
    mod: Any  # ESEBlock
    x: TView
    out: TView
    residual: Optional[TView]
    pidx: int = -1
    st: dict = field(default_factory=dict)
    kind: str = "ese"


# ----------------------------------------------------------------------------------------------------
# graph builder
# ----------------------------------------------------------------------------------------------------
class Graph:
    def __init__(self, training: bool, need_grad: bool, f32: bool = False, pair_ok: bool = False,
                 col_stem: bool = False):
        # the image's first convolution as a 1x1 GEMM over a gathered operand (vtb_im2col_input): tensor-core plans whose
        # input needs no gradient
        self.col_stem = col_stem and not f32
        self.input_col: Optional[TView] = None
        self.input_unused = False   # set by finalize(): the padded NHWC image left the arena (gathered-operand stem)
        self.training = training
        self.need_grad = need_grad
        self.f32 = f32
        # sibling ConvNormAct units (CSP conv1 | conv2) may run as one convolution: training-mode tensor-core plans whose
        # BatchNorm finalisation happens inside the conv kernel (single GPU, or SyncBN over peer memory)
        self.pair_ok = pair_ok and training and not f32
        self.esize = F32 if f32 else BF16
        # BN uses batch statistics only in training mode; the 1-kernel fused epilogue needs frozen statistics
        # (fp32 parity mode keeps conv and normalise separate: it mirrors the reference's op order)
        self.fused_eval = (not training) and (not need_grad) and (not f32)
        self.buffers: list[Buffer] = []
        self.ops: list[Any] = []
        self.params: list[torch.Tensor] = []
        self.outputs: list[TView] = []
        self.input: Optional[TView] = None
        self.input_c = 0
        self.stat_floats = 0
        self.ws_bytes = 0
        self.dy_bytes = 0
        self.act_bytes = 0

    # -- allocation
    def new_buffer(self, n: int, h: int, w: int, c: int, kind: str = "act") -> Buffer:
        b = Buffer(len(self.buffers), n, h, w, c, kind, self.esize)
        self.buffers.append(b)
        return b

    def new_tensor(self, n: int, h: int, w: int, c: int, kind: str = "act") -> TView:
        b = self.new_buffer(n, h, w, c, kind)
        t = TView(b, 0, c)
        b.exclusive_owner = t
        return t

    def slice(self, buf: Buffer, coff: int, c: int) -> TView:
        assert coff % 8 == 0 and c % 8 == 0 and coff + c <= buf.c
        return TView(buf, coff, c)

    def rehome(self, t: TView, buf: Buffer, coff: int) -> bool:
        """Move a not-yet-allocated standalone tensor into a channel slice of `buf` (in-place concat)."""
        old = t.buf
        if old.exclusive_owner is not t or t.is_input or old is buf:
            return False
        if (old.n, old.h, old.w) != (buf.n, buf.h, buf.w):
            return False
        self.buffers.remove(old)
        for i, b in enumerate(self.buffers):
            b.idx = i
        t.buf, t.coff = buf, coff
        return True

    def copy_into(self, x: TView, buf: Buffer, coff: int) -> TView:
        """`x` as channel slice [coff, coff + c) of `buf` when it cannot be re-homed (it is the plan's input): one copy
        launch forward, one gradient add backward (reference vovnet.py:51: ``outputs = [x]`` of a standalone OSABlock)."""
        c = self.input_c if (x.is_input and x is self.input) else x.c
        if c != x.c:
            raise NotImplementedError(f"a block input with {c} channels cannot be placed in a concat buffer (multiple of 16)")
        out = self.slice(buf, coff, c)
        x.consumers.append(len(self.ops))
        self.ops.append(CopyOp(x, out))
        return out

    def _stat(self, op, name: str, floats: int) -> None:
        self.stat_floats = _round_up(self.stat_floats, 64)
        op.st[name] = self.stat_floats
        self.stat_floats += floats

    # -- ops
    def input_image(self, n: int, c: int, h: int, w: int) -> TView:
        cpad = _round_up(c, 16)
        t = self.new_tensor(n, h, w, cpad, "act")
        t.is_input = True
        self.input, self.input_c = t, c
        return t

    def conv_norm_act(self, mod, x: TView, residual: Optional[TView] = None, out: Optional[TView] = None,
                      y: Optional[TView] = None) -> TView:
        conv, norm = mod.conv, mod.norm
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        cin_real, cout = conv.in_channels, conv.out_channels
        if not mod.native_supported():
            raise NotImplementedError(
                "ConvNormAct options outside the Darknet/VoVNet hot path (groups/dilation/norm='none'/act other than "
                "relu|none) have no sm_100a kernel; run that module on its own"
            )
        have = self.input_c if (x.is_input and x is self.input) else x.c
        if have != cin_real:
            # what aten::convolution raises for the reference (components.py:26): the module tree decides, not the data
            raise RuntimeError(f"Given groups=1, weight of size {list(conv.weight.shape)}, expected input to have "
                               f"{cin_real} channels, but got {have} channels instead")
        if cout % 16 or x.c < cin_real or x.c % 16:
            raise NotImplementedError(f"channel counts must be multiples of 16 (got {cin_real}->{cout})")
        geom = VtbConv(x.n, x.h, x.w, x.c, cout, k, s, p)
        ho = (x.h + 2 * p - k) // s + 1
        wo = (x.w + 2 * p - k) // s + 1
        col = None
        if (self.col_stem and x.is_input and x is self.input and cin_real == 3 and k == 3
                and self.input_col is None):
            # 3x3 RGB stems (darknet.py:74; vovnet.py:85): gather the 27 taps of every output pixel once
            # (vtb_im2col_input) and run the convolution as a 1x1 GEMM: one TMA request per tile instead of one per tap
            # (the 6x6 YOLOv5 stem keeps the im2col descriptors: 108 gathered columns cost more than they save, 4.38 ->
            # 4.62 ms per 32 x 640^2 forward)
            kcols = k * k * cin_real
            xc = self.new_tensor(x.n, ho, wo, _round_up(kcols, 16))
            xc.is_input = True
            xc.col_of = (k, s, p, cin_real)
            self.input_col = xc
            col = (k * k, cin_real)
            x, cin_real = xc, kcols
            geom = VtbConv(xc.n, ho, wo, xc.c, cout, 1, 1, 0)
        if out is None:
            out = self.new_tensor(x.n, ho, wo, cout)
        assert (out.n, out.h, out.w, out.c) == (x.n, ho, wo, cout)
        if y is None:
            y = None if self.fused_eval else self.new_tensor(x.n, ho, wo, cout, "raw")
        op = ConvOp(mod, x, y, out, residual, mod.act_is_relu(), geom, cin_real)
        op.col = col
        op.batch_stats = bool(norm.training or norm.running_mean is None)
        if op.batch_stats and norm.momentum is None and norm.running_mean is not None:
            # cumulative moving average (factor 1 / num_batches_tracked): the factor lives on the device; not built
            raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative moving average) has no sm_100a kernel here")
        _check_norm_tensors(mod)
        idx = len(self.ops)
        x.consumers.append(idx)
        if residual is not None:
            assert (residual.n, residual.h, residual.w, residual.c) == (out.n, out.h, out.w, out.c)
            residual.consumers.append(idx)
        op.pidx = len(self.params)
        self.params += [conv.weight, norm.weight, norm.bias]
        L = _lib.lib()
        if self.f32:
            # fp32 parity mode: two-level fp64 sums (vtb_f32_bn_stats / vtb_f32_bn_bwd_reduce), rows of double[c][2]
            rows_f = rows_b = rows_b_alloc = L.vtb_f32_bn_rows(out.pixels, cout)
            if rows_f <= 0:
                check(-1, "vtb_f32_bn_rows")
            per_row = cout * 4
        else:
            rows_f = L.vtb_conv_stats_rows(C.byref(geom))
            if rows_f <= 0:
                check(-1, "vtb_conv_stats_rows")
            rows_b = L.vtb_bn_bwd_rows(out.pixels, cout)            # rows written by vtb_bn_bwd_reduce
            rows_b_alloc = max(rows_b, L.vtb_bn_bwd_fused_rows(out.pixels, cout))   # scratch also serves the fused kernel
            per_row = cout * 2
        for name in ("mean", "invstd", "scale", "shift"):
            self._stat(op, name, cout)
        self._stat(op, "partial_f", rows_f * per_row)
        self._stat(op, "sums", cout * 4)  # double[c][2]
        if self.need_grad:
            self._stat(op, "partial_b", (rows_b_alloc + 1) * per_row)   # + one row: global means under SyncBN
            self._stat(op, "coef", cout * 2)
            self._stat(op, "sums_b", cout * 4)
            self._stat(op, "lsums_b", cout * 4)
            if col is not None:
                self._stat(op, "dw_col", cout * cin_real)   # weight gradient in (tap, ci) column order
            wsq = L.vtb_f32_conv_wgrad_workspace_bytes if self.f32 else L.vtb_conv_wgrad_workspace_bytes
            self.ws_bytes = max(self.ws_bytes, int(wsq(C.byref(geom))))
            self.dy_bytes = max(self.dy_bytes, out.pixels * cout * self.esize)
        op.rows_f, op.rows_b = rows_f, rows_b
        self.ops.append(op)
        return out

    def conv_bias(self, unit, x: TView, out: Optional[TView] = None) -> TView:
        """A bare ``nn.Conv2d(..., bias=True)`` (reference necks.py:60-65 lateral convolutions): `unit.conv` is the module.
        One launch forward (bias in the affine epilogue); backward = dgrad + wgrad straight from the output gradient and a
        per-channel sum for the bias gradient."""
        conv = unit.conv
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        cin_real, cout = conv.in_channels, conv.out_channels
        ok = (conv.bias is not None and conv.groups == 1 and conv.dilation == (1, 1) and conv.kernel_size[0] == conv.kernel_size[1]
              and conv.stride[0] == conv.stride[1] and s in (1, 2) and conv.padding[0] == conv.padding[1]
              and conv.padding_mode == "zeros" and k * k <= 36 and not self.f32)
        if not ok:
            raise NotImplementedError("this nn.Conv2d configuration has no sm_100a kernel here (bf16 plans, groups=1, "
                                      "dilation=1, square kernel, stride 1|2, with bias)")
        have = self.input_c if (x.is_input and x is self.input) else x.c
        if have != cin_real:
            raise RuntimeError(f"Given groups=1, weight of size {list(conv.weight.shape)}, expected input to have "
                               f"{cin_real} channels, but got {have} channels instead")
        if cout % 16 or x.c % 16:
            raise NotImplementedError(f"channel counts must be multiples of 16 (got {cin_real}->{cout})")
        for t in (conv.weight, conv.bias):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError("nn.Conv2d weight and bias must be contiguous float32 tensors for the native path")
        geom = VtbConv(x.n, x.h, x.w, x.c, cout, k, s, p)
        ho, wo = (x.h + 2 * p - k) // s + 1, (x.w + 2 * p - k) // s + 1
        if out is None:
            out = self.new_tensor(x.n, ho, wo, cout)
        op = ConvOp(unit, x, None, out, None, False, geom, cin_real)
        op.bias, op.nparams, op.batch_stats = True, 2, False
        x.consumers.append(len(self.ops))
        op.pidx = len(self.params)
        self.params += [conv.weight, conv.bias]
        L = _lib.lib()
        self._stat(op, "scale", cout)           # ones
        if self.need_grad:
            rows_b = L.vtb_bn_bwd_rows(out.pixels, cout)
            self._stat(op, "partial_b", (rows_b + 1) * cout * 2)
            self._stat(op, "coef", cout * 2)
            op.rows_b = rows_b
            self.ws_bytes = max(self.ws_bytes, int(L.vtb_conv_wgrad_workspace_bytes(C.byref(geom))))
        self.ops.append(op)
        return out

    def conv_norm_act_pair(self, mod_a, mod_b, x: TView, out_a: Optional[TView] = None,
                           out_b: Optional[TView] = None) -> tuple[TView, TView]:
        """Two ConvNormAct units applied to the SAME tensor (CSPDarknetStage conv1 / conv2, reference darknet.py:52-53).

        In training-mode tensor-core plans they become one implicit GEMM over side-by-side packed weights: one fprop
        (statistics + BatchNorm finalisation of both units in its epilogue), one dgrad without a read-modify-write
        fan-in, one wgrad - `x` is read once instead of twice in each of the three.  The normalise / BatchNorm-backward
        passes stay per unit (their outputs live in different buffers).  Elsewhere: two ordinary units."""
        ca, cb, na, nb = mod_a.conv, mod_b.conv, mod_a.norm, mod_b.norm
        bs_a, bs_b = (bool(n_.training or n_.running_mean is None) for n_ in (na, nb))
        compatible = (
            self.pair_ok and bs_a and bs_b and mod_a.native_supported() and mod_b.native_supported()
            and ca.kernel_size == cb.kernel_size and ca.stride == cb.stride and ca.padding == cb.padding
            and ca.in_channels == cb.in_channels and na.eps == nb.eps and na.momentum == nb.momentum
            and na.track_running_stats == nb.track_running_stats and (na.running_mean is None) == (nb.running_mean is None)
        )
        if not compatible:
            return self.conv_norm_act(mod_a, x, out=out_a), self.conv_norm_act(mod_b, x, out=out_b)
        k, s, p = ca.kernel_size[0], ca.stride[0], ca.padding[0]
        ho, wo = (x.h + 2 * p - k) // s + 1, (x.w + 2 * p - k) // s + 1
        c_a, c_b = ca.out_channels, cb.out_channels
        tot = c_a + c_b
        ybuf = self.new_buffer(x.n, ho, wo, tot, "raw")
        first = len(self.ops)
        o_a = self.conv_norm_act(mod_a, x, out=out_a, y=self.slice(ybuf, 0, c_a))
        o_b = self.conv_norm_act(mod_b, x, out=out_b, y=self.slice(ybuf, c_a, c_b))
        op_a, op_b = self.ops[first], self.ops[first + 1]
        op_a.pair, op_b.pair_of = op_b, op_a
        op_a.pair_geom = VtbConv(x.n, x.h, x.w, x.c, tot, k, s, p)
        # the fused launch finalises BatchNorm for all `tot` channels into ONE set of arrays; each unit uses its slice
        L = _lib.lib()
        for name in ("mean", "invstd", "scale", "shift"):
            self._stat(op_a, name, tot)
            op_b.st[name] = op_a.st[name] + c_a
        rows = L.vtb_conv_stats_rows(C.byref(op_a.pair_geom))
        if rows <= 0:
            check(-1, "vtb_conv_stats_rows")
        self._stat(op_a, "pair_partial_f", rows * tot * 2)
        if self.need_grad:
            self.ws_bytes = max(self.ws_bytes, int(L.vtb_conv_wgrad_workspace_bytes(C.byref(op_a.pair_geom))))
            self.dy_bytes = max(self.dy_bytes, o_a.pixels * tot * self.esize)
        return o_a, o_b

    def maxpool(self, x: TView, out: Optional[TView] = None) -> TView:
        ho, wo = (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1
        if out is None:
            out = self.new_tensor(x.n, ho, wo, x.c)
        x.consumers.append(len(self.ops))
        # one byte per output element, carved out of the arena as a (c/2)-channel bf16 tensor
        idx = self.new_tensor(x.n, ho, wo, _round_up(-(-x.c // self.esize), 8), "raw") if self.need_grad else None
        self.ops.append(PoolOp(x, out, idx))
        return out

    def ese(self, mod, x: TView, residual: Optional[TView] = None, out: Optional[TView] = None) -> TView:
        if out is None:
            out = self.new_tensor(x.n, x.h, x.w, x.c)
        for t in (mod.linear.weight, mod.linear.bias):
            if t is None or t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError("ESEBlock.linear weight and bias must be contiguous float32 tensors for the native path "
                                "(model.half()/bfloat16() are not supported)")
        op = EseOp(mod, x, out, residual)
        idx = len(self.ops)
        x.consumers.append(idx)
        if residual is not None:
            residual.consumers.append(idx)
        op.pidx = len(self.params)
        self.params += [mod.linear.weight, mod.linear.bias]
        for name in ("pool", "z", "gate"):
            self._stat(op, name, x.n * x.c)
        if self.need_grad:
            self._stat(op, "scratch", 3 * x.n * x.c)
        self.ops.append(op)
        return out

    def mark_output(self, t: TView) -> None:
        t.is_output = True
        self.outputs.append(t)

    def finalize(self) -> None:
        if self.input_col is not None and self.input is not None and not self.input.consumers and not self.input.is_output:
            self.buffers.remove(self.input.buf)   # the padded NHWC image is not needed: the stem reads the gathered operand
            for i, b in enumerate(self.buffers):
                b.idx = i
            self.input_unused = True
        # materialised activations first (the gradient arena mirrors exactly this prefix), raw conv outputs after
        off = 0
        for b in self.buffers:
            if b.kind == "act":
                b.offset = off
                off += _round_up(b.nbytes, 1024)
        self.grad_bytes = off
        for b in self.buffers:
            if b.kind != "act":
                b.offset = off
                off += _round_up(b.nbytes, 1024)
        self.act_bytes = off
        # Residual gradient aliasing: for out = res + f(res) the gradient of `res` can live in the gradient memory
        # of `out` (every other contribution is accumulated into it) iff nothing consumes `res` after this op
        # and `res` is a standalone tensor.
        for i, op in enumerate(self.ops):
            res = getattr(op, "residual", None)
            if (res is not None and not res.is_output and not res.is_input and max(res.consumers) == i
                    and res.buf.exclusive_owner is res):  # a concat slice also receives gradient through the full view
                res.grad_alias = op.out
        self._plan_dgrad_bn()

    def _plan_dgrad_bn(self) -> None:
        """Decide which dgrad launches also reduce the BatchNorm(+ReLU) backward sums of the layer(s) that produced their
        output tensor (``vtb_conv_dgrad_bn``; autograd chain ConvolutionBackward0 -> ReluBackward0 ->
        NativeBatchNormBackward0 of reference components.py:26-39).

        The backward pass visits the ops in reverse order; op i's dgrad writes (or accumulates into) the gradient of its
        input view X.  That write completes the gradient of a producer's output region R inside X iff i is the
        smallest-index op that reads any view overlapping R (every other contribution - later ops' dgrads, residual
        identities, the seeds of graph outputs - has been applied before op i is visited).  X must be tiled exactly by
        the outputs of one or two ConvNormAct units (two: the CSP concat buffer, darknet.py:53)."""
        # Opt-in (VTB_DGRAD_BN=1).  Measured on B200 (profiles/r02_dgrad_bn_*.txt): the per-element mask / sum work costs the
        # 16 epilogue warps of the GEMM more than the standalone reduce pass saves (1x1 128->128 @22^2: dgrad 16.7 -> 38.6 us
        # per launch for 10 us saved in the BatchNorm kernel), so the default plan keeps vtb_bn_bwd_fused.
        if not (self.training and self.need_grad) or self.f32 or _os.environ.get("VTB_DGRAD_BN", "0") != "1":
            return
        L = _lib.lib()
        readers: dict[int, list[tuple[int, int, int, str]]] = {}   # buffer idx -> (op idx, c0, c1, role)
        producers: dict[int, list[Any]] = {}
        for i, op in enumerate(self.ops):
            for role in ("x", "residual"):
                t = getattr(op, role, None)
                if t is not None:
                    readers.setdefault(id(t.buf), []).append((i, t.coff, t.coff + t.c, role))
            if op.kind == "conv" and op.y is not None and op.batch_stats:
                producers.setdefault(id(op.out.buf), []).append(op)
        for i, op in enumerate(self.ops):
            if op.kind != "conv" or op.pair_of is not None or op.col is not None:
                continue
            x = op.x
            if x.is_input:
                continue
            geom = op.pair_geom if op.pair is not None else op.geom
            c0, c1 = x.coff, x.coff + x.c
            prods = sorted((q for q in producers.get(id(x.buf), ())
                            if q.out.coff >= c0 and q.out.coff + q.out.c <= c1 and q.bwd_stats_from is None),
                           key=lambda q: q.out.coff)
            if not prods or len(prods) > 2 or prods[0].out.coff != c0 or prods[-1].out.coff + prods[-1].out.c != c1:
                continue
            if len(prods) == 2 and prods[0].out.coff + prods[0].out.c != prods[1].out.coff:
                continue
            ok = True
            for q in prods:
                r0, r1 = q.out.coff, q.out.coff + q.out.c
                first = min((j, role) for j, a, b, role in readers.get(id(x.buf), ()) if a < r1 and b > r0)
                ok = ok and first == (i, "x") and (q.out.n, q.out.h, q.out.w) == (x.n, x.h, x.w)
            split = prods[0].out.c if len(prods) == 2 else 0
            pw = L.vtb_conv_dgrad_panel_w(C.byref(geom))
            if not ok or pw <= 0 or split % pw:
                continue
            rows = L.vtb_conv_dgrad_stats_rows(C.byref(geom))
            if rows <= 0:
                check(-1, "vtb_conv_dgrad_stats_rows")
            op.dgrad_bn = (prods, split)
            for q in prods:
                q.bwd_stats_from = op
            self._stat(op, "partial_d", rows * x.c * 2)


# ----------------------------------------------------------------------------------------------------
# executor
# ----------------------------------------------------------------------------------------------------
class Run:
    """Device state of one forward call that backward needs (activation + statistics arenas)."""

    __slots__ = ("act", "stat", "x_shape", "x_dtype", "x_requires_grad", "count_scale", "bwd_stats_done")


class DistConfig:
    """Data-parallel hooks (vision_toolbox_b200.parallel fills this in).

    SyncBN statistics travel either through the one-kernel NVLink peer-memory exchange (``vtb_bn_sync_*``; buffers from
    ``torch.distributed._symmetric_memory``) or, when peer memory cannot be set up (``VTB_SYNCBN=nccl`` forces it),
    through one NCCL all-reduce per exchange.
    """

    def __init__(self, group=None, sync_bn: bool = True, device: Optional[torch.device] = None):
        import os

        import torch.distributed as dist

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.sync_bn = sync_bn and self.world > 1
        self.sync = None          # _lib.VtbSyncBn when the peer-memory path is active
        self.sync_error = None
        self.on_grads_ready = None  # callable(list[Tensor]) : parameter gradients that are final (parallel.Trainer)
        if (self.sync_bn and device is not None and device.type == "cuda"
                and os.environ.get("VTB_SYNCBN", "p2p") == "p2p" and self.world <= 8):
            try:
                self._setup_peer_memory(device)
            except Exception as e:  # noqa: BLE001 - any failure selects the NCCL exchange, loudly
                self.sync_error = f"{type(e).__name__}: {e}"
                if self.rank == 0:
                    print(f"[vision_toolbox_b200] SyncBN peer-memory exchange unavailable ({self.sync_error}); "
                          "using NCCL all-reduce per layer", flush=True)
            # every rank must agree on the path
            flag = torch.tensor([1 if self.sync is not None else 0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag) == 0:
                self.sync = None

    def _setup_peer_memory(self, device: torch.device) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        L = _lib.lib()
        nbytes = int(L.vtb_bn_sync_buffer_bytes())
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)   # every buffer is zeroed before anyone pushes into it
        sync = _lib.VtbSyncBn()
        sync.rank, sync.world = int(hdl.rank), int(hdl.world_size)
        ptrs = list(hdl.buffer_ptrs)
        assert sync.world == self.world and len(ptrs) == self.world
        for r, ptr in enumerate(ptrs):
            sync.peer_buffers[r] = int(ptr)
        self._symm_buf, self._symm_hdl = buf, hdl   # keep the mapping alive
        self.sync = sync

    def all_reduce_(self, t: torch.Tensor) -> None:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


class Runner:
    def __init__(self, graph: Graph, device: torch.device, dist_cfg: Optional["DistConfig"] = None):
        self.g = graph
        self.device = device
        self.L = _lib.lib()
        self.dist: Optional[DistConfig] = dist_cfg
        self.grad_sink = False
        self.on_grads_ready = None   # single-process hook: callable(list[Parameter]) - gradients that are final
        self.last_all_direct = False
        # ticket counters of the fused conv + BatchNorm-finalize kernels (self-cleaning, shared by all layers)
        self.tickets = torch.zeros(512, dtype=torch.int32, device=device)   # [0,256): conv kernels, [256,512): BN bwd
        # element type of the activation / gradient arenas and the entry points whose two precisions share a signature
        L = self.L
        self.f32 = graph.f32
        self._conv_ops = [op for op in graph.ops if op.kind == "conv"]
        # Weight-gradient GEMMs on a second stream (VTB_WGRAD_STREAM=0 switches it off): wgrad(i) only needs dy(i) and
        # x(i), so it fills the ramp-up / tail bubbles of the dependent chain bn_bwd -> dgrad -> bn_bwd ... on the main
        # stream (15.47 -> 15.16 ms per CSPDarknet-53 step; with SyncBN it also fills the exchange waits: 15.15 -> 14.78 ms
        # at 2 GPUs).  With more than one rank it needs the SM reserve (VTB_SM_RESERVE > 0, parallel.nccl_pg_options):
        # without it a SyncBN kernel spinning for its peers on every SM, an NCCL kernel waiting for its peer and a third
        # stream both depend on closed a cross-rank wait cycle (exchange timeout at 2 GPUs).  VTB_WGRAD_STREAM_MULTI=0
        # keeps multi-rank plans on one stream.
        # normalise + ReLU (+ residual) of a unit inside its convolution's launch (VTB_FUSED_NORM=1; single units of bf16
        # plans): the conv kernel's CTAs apply the finished coefficients to their own tiles after a flag wait
        self.fused_norm = _os.environ.get("VTB_FUSED_NORM", "0") == "1" and not graph.f32
        # stride-2 dgrad of the few-channel layers as one GEMM over 2x2 super-pixels (VTB_DGRAD_S2_MERGED, csrc/api_conv.cu)
        self.dgrad_s2 = _os.environ.get("VTB_DGRAD_S2_MERGED", "1") == "1" and not graph.f32
        self._s2_ws: dict = {}
        self._side = None
        multi_rank = dist_cfg is not None and dist_cfg.world > 1
        multi_ok = (_os.environ.get("VTB_WGRAD_STREAM_MULTI", "1") == "1"
                    and _os.environ.get("VTB_SM_RESERVE", "16").strip() not in ("0", ""))
        if (_os.environ.get("VTB_WGRAD_STREAM", "1") == "1" and not graph.f32 and device.type == "cuda"
                and (not multi_rank or multi_ok)):
            self._side = torch.cuda.Stream(device)
        self._dy_slots = 3 if self._side is not None else 1
        self._dy_free: list = []
        self._pack_key, self._pack_jobs, self._pack_launches, self._pack_srcs = None, None, [], None
        self._packs_token = None   # set by parallel.NativeSGD after it re-packed the operands itself
        self.tdtype = torch.float32 if self.f32 else torch.bfloat16
        self.fn_grad_add = L.vtb_f32_grad_add if self.f32 else L.vtb_grad_add
        self.fn_to_nhwc = L.vtb_f32_nchw_to_nhwc if self.f32 else L.vtb_nchw_to_nhwc
        self.fn_pool_fwd = L.vtb_f32_maxpool3s2_fwd if self.f32 else L.vtb_maxpool3s2_fwd
        self.fn_pool_bwd = L.vtb_f32_maxpool3s2_bwd if self.f32 else L.vtb_maxpool3s2_bwd
        self.fn_ese_fwd = L.vtb_f32_ese_fwd if self.f32 else L.vtb_ese_fwd
        self.fn_ese_bwd = L.vtb_f32_ese_bwd if self.f32 else L.vtb_ese_bwd

    # -- helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _refresh_packs(self, st: int) -> None:
        """Rebuild the bf16 operand packs (``[Cout][tap][Cin]`` for fprop / wgrad order, ``[Cin][tap][Cout]`` for dgrad) of
        EVERY convolution of the plan from the fp32 master weights: one launch at the start of every forward.

        Not cached on ``Parameter._version``: fused optimizers (``torch.optim.SGD(fused=True)``, fused AdamW) update
        parameters in place WITHOUT bumping the version counter, so a version-keyed cache silently serves stale weights.
        """
        L = self.L
        convs = self._conv_ops
        if not convs:
            return
        if self._packs_token is not None and self._packs_token == self.pack_token():
            # the native optimizer (parallel.NativeSGD -> vtb_sgd_pack_weights) cut the operands from the weights it has
            # just written, and nothing touched the masters since (torch bumps Parameter._version on every in-place
            # write; the token is only ever set by that optimizer): the packs are current
            return
        self._packs_token = None
        srcs = []
        for op in convs:
            w = op.mod.conv.weight.detach()
            if w.dtype != torch.float32 or not w.is_contiguous():
                w = w.float().contiguous()
            srcs.append(w)
        key = tuple(w.data_ptr() for w in srcs)
        if key != self._pack_key:
            table, launches = self.pack_job_table(srcs)
            self._pack_jobs = table
            self._pack_launches, self._pack_key = launches, key
        self._pack_srcs = srcs   # converted copies (non-fp32 masters) must outlive the launch
        base, rec = self._pack_jobs.data_ptr(), C.sizeof(_lib.VtbPackJob)
        for j0, n, blocks in self._pack_launches:
            check(L.vtb_pack_weights(base + j0 * rec, n, blocks, st), "vtb_pack_weights")

    def pack_token(self) -> tuple:
        """(storage, version) of every convolution master weight of the plan."""
        return tuple((op.mod.conv.weight.data_ptr(), op.mod.conv.weight._version) for op in self._conv_ops)

    def pack_job_table(self, srcs, sgd=None, subset=None):
        """Device-resident VtbPackJob table of every convolution of the plan (+ the launches that walk it).  `sgd`:
        callable(parameter) -> (gradient pointer, momentum pointer, weight decay) for vtb_sgd_pack_weights.  `subset`:
        indices into the plan's convolutions (a table of just those, e.g. one gradient bucket)."""
        L = self.L
        convs = self._conv_ops
        if subset is not None:
            convs, srcs = [convs[i] for i in subset], [srcs[i] for i in subset]
        if True:
            jobs = (_lib.VtbPackJob * max(1, len(convs)))()
            blk, launches = 0, []
            for j, (op, w) in enumerate(zip(convs, srcs)):
                g = op.geom
                n = g.cout * g.k * g.k * g.cin
                first = op if op.pair is not None else op.pair_of
                if op.col is not None:
                    # gathered-operand stem: rows of k*k*c columns in (tap, ci) order, padded to g.cin (pads stay zero)
                    taps, creal = op.col
                    cache = op.mod.__dict__.get("_vtb_wpack_col")
                    if cache is None or cache[0].numel() != g.cout * g.cin or cache[0].device != w.device:
                        cache = (torch.zeros(g.cout * g.cin, dtype=torch.bfloat16, device=w.device), None)
                        op.mod.__dict__["_vtb_wpack_col"] = cache
                    jobs[j] = _lib.VtbPackJob(w.data_ptr(), cache[0].data_ptr(), None, g.cout, creal, creal, taps, g.cout, 0,
                                              g.cin, blk)
                    nb = int(L.vtb_pack_job_blocks(g.cout, creal, taps))
                elif first is not None:
                    # side-by-side pair: both units pack into ONE [tot][tap][cin] / [cin][tap][tot] operand pair
                    tot = first.pair_geom.cout
                    npair = tot * g.k * g.k * g.cin
                    cache = first.mod.__dict__.get("_vtb_wpack_pair")
                    if cache is None or cache[0].numel() != npair or cache[0].device != w.device:
                        cache = (torch.empty(npair, dtype=torch.bfloat16, device=w.device),
                                 torch.empty(npair, dtype=torch.bfloat16, device=w.device))
                        first.mod.__dict__["_vtb_wpack_pair"] = cache
                    co_off = 0 if op is first else first.geom.cout
                    wf_ptr = cache[0].data_ptr() + co_off * g.k * g.k * g.cin * BF16
                    jobs[j] = _lib.VtbPackJob(w.data_ptr(), wf_ptr, cache[1].data_ptr(), g.cout, op.cin_real,
                                              g.cin, g.k * g.k, tot, co_off, g.k * g.k * g.cin, blk)
                    nb = int(L.vtb_pack_job_blocks(g.cout, g.cin, g.k * g.k))
                else:
                    cache = op.mod.__dict__.get("_vtb_wpack")
                    if cache is None or cache[0].numel() != n or cache[0].device != w.device:
                        cache = (torch.empty(n, dtype=torch.bfloat16, device=w.device),
                                 torch.empty(n, dtype=torch.bfloat16, device=w.device))
                        op.mod.__dict__["_vtb_wpack"] = cache
                    jobs[j] = _lib.VtbPackJob(w.data_ptr(), cache[0].data_ptr(), cache[1].data_ptr(), g.cout, op.cin_real,
                                              g.cin, g.k * g.k, g.cout, 0, g.k * g.k * g.cin, blk)
                    nb = int(L.vtb_pack_job_blocks(g.cout, g.cin, g.k * g.k))
                if nb <= 0:
                    check(-1, "vtb_pack_job_blocks")
                if sgd is not None:
                    jobs[j].g, jobs[j].m, jobs[j].weight_decay = sgd(op.mod.conv.weight)
                blk += nb
                if (j + 1) % 256 == 0 or j == len(convs) - 1:   # a launch serves at most 256 jobs
                    launches.append((j // 256 * 256, j % 256 + 1, blk))
                    blk = 0
            table = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8)
            return table.to(self.device), launches

    @staticmethod
    def _packed(op: ConvOp):
        """(wf, wd) bf16 packs of this convolution, current as of the last forward of its module."""
        return op.mod.__dict__["_vtb_wpack_col" if op.col is not None else "_vtb_wpack"]

    def view_tensor(self, arena: torch.Tensor, t: TView) -> torch.Tensor:
        """Zero-copy logical-NCHW (channels_last strided) tensor over a view of the arena."""
        base = arena[t.buf.offset : t.buf.offset + t.buf.nbytes].view(self.tdtype)
        ld = t.ld
        return base.as_strided((t.n, t.c, t.h, t.w), (t.h * t.w * ld, 1, t.w * ld, ld), base.storage_offset() + t.coff)

    # -- forward
    def forward(self, x: torch.Tensor):
        g, L = self.g, self.L
        st = self._stream()
        dev = x.device
        run = Run()
        run.act = torch.empty(g.act_bytes, dtype=torch.uint8, device=dev)
        run.stat = torch.empty(g.stat_floats + 64, dtype=torch.float32, device=dev)
        run.x_shape, run.x_dtype, run.x_requires_grad = tuple(x.shape), x.dtype, x.requires_grad
        abase, sbase = run.act.data_ptr(), run.stat.data_ptr()
        sbase = (sbase + 255) // 256 * 256

        # input layout conversion (NCHW float -> NHWC bf16, channels padded to a multiple of 16)
        xin = x.detach()
        n, c, h, w = xin.shape
        t_in = g.input
        if (not self.f32 and g.input_col is None and xin.dtype == torch.bfloat16 and c == t_in.c and xin.stride(1) == 1
                and xin.stride(3) % 8 == 0 and xin.stride(2) == w * xin.stride(3) and xin.stride(0) == h * w * xin.stride(3)
                and xin.data_ptr() % 16 == 0):
            # already an NHWC bf16 view (a feature map of another native plan): one strided copy into the arena
            check(self.fn_grad_add(abase + t_in.byte_offset(), t_in.c, xin.data_ptr(), xin.stride(3), n * h * w, c, 0, st),
                  "vtb_grad_add(input)")
            xin = None
        elif xin.dtype != torch.float32 or not xin.is_contiguous():
            xin = xin.float().contiguous()
        if g.input_col is not None:
            xc = g.input_col
            kk_, ss_, pp_, _ = xc.col_of
            check(L.vtb_im2col_input(xin.data_ptr(), n, c, h, w, kk_, ss_, pp_, abase + xc.byte_offset(), xc.c, st),
                  "vtb_im2col_input")
        if not g.input_unused and xin is not None:
            check(self.fn_to_nhwc(xin.data_ptr(), n, c, h, w, abase + t_in.byte_offset(), t_in.c, st), "vtb_nchw_to_nhwc")

        if not self.f32:
            self._refresh_packs(st)
        world = self.dist.world if (self.dist is not None and self.dist.sync_bn) else 1
        for op in g.ops:
            if op.kind == "conv":
                if self.f32:
                    self._conv_forward_f32(op, abase, sbase, run, st, world)
                elif op.pair is not None:
                    self._conv_forward_pair(op, abase, sbase, run, st, world)
                elif op.pair_of is None:    # the second unit of a pair ran together with the first
                    self._conv_forward(op, abase, sbase, run, st, world)
            elif op.kind == "copy":
                xx, oo = op.x, op.out
                check(self.fn_grad_add(abase + oo.byte_offset(), oo.ld, abase + xx.byte_offset(), xx.ld, xx.pixels, xx.c, 0,
                                       st), "vtb_grad_add(copy)")
            elif op.kind == "pool":
                xx, oo = op.x, op.out
                check(self.fn_pool_fwd(abase + xx.byte_offset(), xx.ld, xx.n, xx.h, xx.w, xx.c,
                                           abase + oo.byte_offset(), oo.ld,
                                           abase + op.idx.byte_offset() if op.idx is not None else 0, st),
                      "vtb_maxpool3s2_fwd")
            elif op.kind == "ese":
                xx, oo, rr = op.x, op.out, op.residual
                lin = op.mod.linear
                check(self.fn_ese_fwd(abase + xx.byte_offset(), xx.ld, xx.n, xx.h * xx.w, xx.c,
                                    lin.weight.data_ptr(), lin.bias.data_ptr(),
                                    0 if rr is None else abase + rr.byte_offset(), 0 if rr is None else rr.ld,
                                    abase + oo.byte_offset(), oo.ld, sbase + 4 * op.st["pool"], sbase + 4 * op.st["z"],
                                    sbase + 4 * op.st["gate"], st), "vtb_ese_fwd")
        outs = [self.view_tensor(run.act, t) for t in g.outputs]
        return outs, run

    def _conv_forward(self, op: ConvOp, abase: int, sbase: int, run: Run, st: int, world: int) -> None:
        g, L = self.g, self.L
        geom, norm = op.geom, (None if op.bias else op.mod.norm)
        wf, _ = self._packed(op)
        x, out, res = op.x, op.out, op.residual
        f = lambda name: sbase + 4 * op.st[name]
        cout = geom.cout
        res_p = 0 if res is None else abase + res.byte_offset()
        res_ld = 0 if res is None else res.ld
        if op.bias:
            ones = run.stat_view_f32(op.st["scale"], cout, sbase)
            ones.fill_(1.0)
            check(L.vtb_conv_fprop(C.byref(geom), abase + x.byte_offset(), x.ld, wf.data_ptr(), abase + out.byte_offset(),
                                   out.ld, 0, ones.data_ptr(), op.mod.conv.bias.data_ptr(), 0, 0, 0, st),
                  "vtb_conv_fprop(bias)")
            return
        if g.fused_eval:
            check(L.vtb_bn_eval_affine(cout, norm.weight.data_ptr(), norm.bias.data_ptr(), norm.running_mean.data_ptr(),
                                       norm.running_var.data_ptr(), norm.eps, f("scale"), f("shift"), st),
                  "vtb_bn_eval_affine")
            check(L.vtb_conv_fprop(C.byref(geom), abase + x.byte_offset(), x.ld, wf.data_ptr(),
                                   abase + out.byte_offset(), out.ld, 0, f("scale"), f("shift"), int(op.relu), res_p,
                                   res_ld, st), "vtb_conv_fprop(eval)")
            return
        y = op.y
        use_batch_stats = op.batch_stats
        if use_batch_stats:
            mom = norm.momentum if norm.momentum is not None else 0.1
            track = norm.track_running_stats and norm.running_mean is not None
            rm = norm.running_mean.data_ptr() if track else 0
            rv = norm.running_var.data_ptr() if track else 0
            nbt = norm.num_batches_tracked.data_ptr() if track else 0
            count = float(out.pixels)
        peer_sync = self.dist.sync if (world > 1 and self.dist is not None) else None
        fused = use_batch_stats and (world == 1 or peer_sync is not None)
        if fused:
            # conv + statistics (+ SyncBN exchange over NVLink) + BatchNorm finalisation in ONE launch
            bn = VtbBnTrain(count * world, norm.weight.data_ptr(), norm.bias.data_ptr(), norm.eps, mom, rm, rv, nbt,
                            f("mean"), f("invstd"), f("scale"), f("shift"), self.tickets.data_ptr(),
                            C.addressof(peer_sync) if peer_sync is not None else None)
            if self.fused_norm:
                # the normalise + ReLU (+ residual) pass rides in the same launch (VtbBnTrain.act_*): no vtb_bn_act
                bn.act_out, bn.act_ld, bn.act_relu = abase + out.byte_offset(), out.ld, int(op.relu)
                bn.act_residual, bn.act_ldr = (res_p or None), res_ld
            check(L.vtb_conv_fprop_bn(C.byref(geom), abase + x.byte_offset(), x.ld, wf.data_ptr(),
                                      abase + y.byte_offset(), y.ld, f("partial_f"), C.byref(bn), st),
                  "vtb_conv_fprop_bn")
            if self.fused_norm:
                return
        else:
            check(L.vtb_conv_fprop(C.byref(geom), abase + x.byte_offset(), x.ld, wf.data_ptr(),
                                   abase + y.byte_offset(), y.ld, f("partial_f") if use_batch_stats else 0, 0, 0, 0, 0,
                                   0, st), "vtb_conv_fprop")
        if use_batch_stats and not fused:
            # SyncBN through NCCL: local sums -> cross-rank sum -> finalise with the GLOBAL element count
            check(L.vtb_bn_stats_reduce(f("partial_f"), op.rows_f, cout, f("sums"), st), "vtb_bn_stats_reduce")
            sums = run.stat_view_f64(op.st["sums"], cout * 2, sbase)
            self.dist.all_reduce_(sums)
            check(L.vtb_bn_finalize(0, 0, f("sums"), count * world, cout, norm.weight.data_ptr(),
                                    norm.bias.data_ptr(), norm.eps, mom, rm, rv, nbt, f("mean"), f("invstd"),
                                    f("scale"), f("shift"), st), "vtb_bn_finalize")
        elif not use_batch_stats:
            # frozen statistics but autograd requested: affine from running stats, mean/invstd kept for backward
            check(L.vtb_bn_eval_affine(cout, norm.weight.data_ptr(), norm.bias.data_ptr(), norm.running_mean.data_ptr(),
                                       norm.running_var.data_ptr(), norm.eps, f("scale"), f("shift"), st),
                  "vtb_bn_eval_affine")
            mean = run.stat_view_f32(op.st["mean"], cout, sbase)
            invstd = run.stat_view_f32(op.st["invstd"], cout, sbase)
            mean.copy_(norm.running_mean)
            invstd.copy_(torch.rsqrt(norm.running_var + norm.eps))
        check(L.vtb_bn_act(abase + y.byte_offset(), y.ld, out.pixels, cout, f("scale"), f("shift"), int(op.relu),
                           res_p, res_ld, abase + out.byte_offset(), out.ld, st), "vtb_bn_act")

    def _conv_forward_pair(self, op_a: ConvOp, abase: int, sbase: int, run: Run, st: int, world: int) -> None:
        """Both units of a side-by-side pair: ONE conv + statistics + BatchNorm finalisation launch, then one
        normalise+ReLU pass per unit (their destinations differ: a concat slice and a standalone tensor)."""
        L = self.L
        op_b, geom = op_a.pair, op_a.pair_geom
        na, nb = op_a.mod.norm, op_b.mod.norm
        peer_sync = self.dist.sync if (world > 1 and self.dist is not None) else None
        if world > 1 and peer_sync is None:
            raise RuntimeError("paired plan built for the in-kernel BatchNorm finalisation, but SyncBN runs over NCCL")
        f = lambda name: sbase + 4 * op_a.st[name]
        mom = na.momentum if na.momentum is not None else 0.1
        track = na.track_running_stats and na.running_mean is not None
        ptr = lambda t: t.data_ptr() if track else 0
        bn = VtbBnTrain(float(op_a.out.pixels) * world, na.weight.data_ptr(), na.bias.data_ptr(), na.eps, mom,
                        ptr(na.running_mean), ptr(na.running_var), ptr(na.num_batches_tracked),
                        f("mean"), f("invstd"), f("scale"), f("shift"), self.tickets.data_ptr(),
                        C.addressof(peer_sync) if peer_sync is not None else None)
        bn.split = op_a.geom.cout
        bn.gamma2, bn.beta2 = nb.weight.data_ptr(), nb.bias.data_ptr()
        if track:
            bn.running_mean2, bn.running_var2 = nb.running_mean.data_ptr(), nb.running_var.data_ptr()
            bn.num_batches_tracked2 = nb.num_batches_tracked.data_ptr()
        wf, _ = op_a.mod.__dict__["_vtb_wpack_pair"]
        x, y = op_a.x, op_a.y   # y: slice 0 of the shared raw buffer, pitch = both units
        check(L.vtb_conv_fprop_bn(C.byref(geom), abase + x.byte_offset(), x.ld, wf.data_ptr(), abase + y.byte_offset(),
                                  y.ld, f("pair_partial_f"), C.byref(bn), st), "vtb_conv_fprop_bn(pair)")
        for op in (op_a, op_b):
            fo = lambda name, op=op: sbase + 4 * op.st[name]
            yy, out, res = op.y, op.out, op.residual
            check(L.vtb_bn_act(abase + yy.byte_offset(), yy.ld, out.pixels, op.geom.cout, fo("scale"), fo("shift"),
                               int(op.relu), 0 if res is None else abase + res.byte_offset(),
                               0 if res is None else res.ld, abase + out.byte_offset(), out.ld, st), "vtb_bn_act")

    def _conv_backward_pair(self, op_a: ConvOp, abase, sbase, gp, gld, is_init, mark, dybase, wsbase, pgrads, run, st, world):
        """Backward of a side-by-side pair: BatchNorm+ReLU backward per unit into the two column ranges of ONE dy
        buffer, then one dgrad (plain store: no fan-in accumulate) and one wgrad (rows split over the two weights)."""
        L = self.L
        op_b, geom = op_a.pair, op_a.pair_geom
        tot = geom.cout
        peer_sync = self.dist.sync if (world > 1 and self.dist is not None) else None
        for op, off in ((op_b, op_a.geom.cout), (op_a, 0)):
            fo = lambda name, op=op: sbase + 4 * op.st[name]
            out, yy = op.out, op.y
            if id(op) in run.bwd_stats_done:
                check(L.vtb_bn_bwd_apply(gp(out), gld(out), abase + yy.byte_offset(), yy.ld, out.pixels, op.geom.cout,
                                         fo("scale"), fo("shift"), fo("mean"), fo("invstd"), int(op.relu), fo("coef"),
                                         dybase + off * BF16, tot, st), "vtb_bn_bwd_apply(pair)")
                continue
            check(L.vtb_bn_bwd_fused(gp(out), gld(out), abase + yy.byte_offset(), yy.ld, out.pixels, op.geom.cout,
                                     fo("scale"), fo("shift"), fo("mean"), fo("invstd"), int(op.relu),
                                     float(out.pixels) * world, fo("partial_b"), pgrads[op.pidx + 1].data_ptr(),
                                     pgrads[op.pidx + 2].data_ptr(), 0, self.tickets.data_ptr() + 1024,
                                     dybase + off * BF16, tot, C.byref(peer_sync) if peer_sync is not None else None, st),
                  "vtb_bn_bwd_fused(pair)")
        _, wd = op_a.mod.__dict__["_vtb_wpack_pair"]
        x = op_a.x
        self._launch_wgrad(lambda s_: check(L.vtb_conv_wgrad_pair(C.byref(geom), dybase, tot, abase + x.byte_offset(), x.ld,
                                                                  wsbase, pgrads[op_a.pidx].data_ptr(),
                                                                  pgrads[op_b.pidx].data_ptr(), op_a.geom.cout,
                                                                  op_a.cin_real, 0, s_), "vtb_conv_wgrad_pair"), st)
        if not (x.is_input and not run.x_requires_grad):
            self._dgrad(op_a, geom, dybase, tot, wd, abase, gp, gld, is_init, pgrads, run, st)
            mark(x)
        for op in (op_a, op_b):
            self._residual_grad(op, gp, gld, is_init, mark, st)

    def _conv_forward_f32(self, op: ConvOp, abase: int, sbase: int, run: Run, st: int, world: int) -> None:
        """fp32 parity mode of one ConvNormAct (reference components.py:26-39 without autocast): conv -> fp64 statistics
        (-> all-reduce of the sums under SyncBN) -> finalise -> normalise + ReLU (+ residual)."""
        g, L = self.g, self.L
        geom, norm, conv = op.geom, op.mod.norm, op.mod.conv
        x, y, out, res = op.x, op.y, op.out, op.residual
        f = lambda name: sbase + 4 * op.st[name]
        cout = geom.cout
        w = self._master_weight(op)
        check(L.vtb_f32_conv_fprop(C.byref(geom), abase + x.byte_offset(), x.ld, w.data_ptr(), op.cin_real,
                                   abase + y.byte_offset(), y.ld, st), "vtb_f32_conv_fprop")
        if op.batch_stats:
            mom = norm.momentum if norm.momentum is not None else 0.1
            track = norm.track_running_stats and norm.running_mean is not None
            rm = norm.running_mean.data_ptr() if track else 0
            rv = norm.running_var.data_ptr() if track else 0
            nbt = norm.num_batches_tracked.data_ptr() if track else 0
            check(L.vtb_f32_bn_stats(abase + y.byte_offset(), y.ld, out.pixels, cout, f("partial_f"), f("sums"), st),
                  "vtb_f32_bn_stats")
            if world > 1:
                self.dist.all_reduce_(run.stat_view_f64(op.st["sums"], cout * 2, sbase))
            check(L.vtb_bn_finalize(0, 0, f("sums"), float(out.pixels) * world, cout, norm.weight.data_ptr(),
                                    norm.bias.data_ptr(), norm.eps, mom, rm, rv, nbt, f("mean"), f("invstd"),
                                    f("scale"), f("shift"), st), "vtb_bn_finalize")
        else:
            run.stat_view_f32(op.st["mean"], cout, sbase).copy_(norm.running_mean)
            run.stat_view_f32(op.st["invstd"], cout, sbase).copy_(torch.rsqrt(norm.running_var + norm.eps))
        check(L.vtb_f32_bn_act(abase + y.byte_offset(), y.ld, out.pixels, cout, f("mean"), f("invstd"),
                               norm.weight.data_ptr(), norm.bias.data_ptr(), int(op.relu),
                               0 if res is None else abase + res.byte_offset(), 0 if res is None else res.ld,
                               abase + out.byte_offset(), out.ld, st), "vtb_f32_bn_act")

    @staticmethod
    def _master_weight(op: ConvOp) -> torch.Tensor:
        """The OIHW fp32 master weight as the fp32 kernels read it (no packing in parity mode)."""
        w = op.mod.conv.weight.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            w = w.float().contiguous()
            op.mod.__dict__["_vtb_w32"] = w   # keep the converted copy alive until the stream has consumed it
        return w

    def _conv_backward_f32(self, op: ConvOp, abase, sbase, gp, gld, is_init, mark, dybase, wsbase, pgrads, run, st, world):
        g, L = self.g, self.L
        geom, cout, norm = op.geom, op.geom.cout, op.mod.norm
        x, y, out, res = op.x, op.y, op.out, op.residual
        f = lambda name: sbase + 4 * op.st[name]
        dout_p, dout_ld = gp(out), gld(out)
        bn = (f("mean"), f("invstd"), norm.weight.data_ptr(), norm.bias.data_ptr(), int(op.relu))
        count = float(out.pixels)
        dgamma, dbeta = pgrads[op.pidx + 1].data_ptr(), pgrads[op.pidx + 2].data_ptr()
        check(L.vtb_f32_bn_bwd_reduce(dout_p, dout_ld, abase + y.byte_offset(), y.ld, out.pixels, cout, *bn,
                                      f("partial_b"), f("lsums_b"), st), "vtb_f32_bn_bwd_reduce")
        if op.batch_stats and world > 1:
            sums = run.stat_view_f64(op.st["sums_b"], cout * 2, sbase)
            sums.copy_(run.stat_view_f64(op.st["lsums_b"], cout * 2, sbase))
            self.dist.all_reduce_(sums)
            check(L.vtb_bn_bwd_finalize(0, 0, f("sums_b"), f("lsums_b"), count * world, cout, dgamma, dbeta, 0,
                                        f("coef"), 0, st), "vtb_bn_bwd_finalize")
        else:
            check(L.vtb_bn_bwd_finalize(0, 0, f("lsums_b"), 0, count, cout, dgamma, dbeta, 0, f("coef"), 0, st),
                  "vtb_bn_bwd_finalize")
            if not op.batch_stats:
                # frozen statistics: mean/var are constants, so dy = gamma * invstd * dz
                run.stat_view_f32(op.st["coef"], cout * 2, sbase).zero_()
        check(L.vtb_f32_bn_bwd_apply(dout_p, dout_ld, abase + y.byte_offset(), y.ld, out.pixels, cout, *bn, f("coef"),
                                     dybase, cout, st), "vtb_f32_bn_bwd_apply")
        w = self._master_weight(op)
        if not (x.is_input and not run.x_requires_grad):
            check(L.vtb_f32_conv_dgrad(C.byref(geom), dybase, cout, w.data_ptr(), op.cin_real, gp(x), gld(x),
                                       int(is_init(x)), st), "vtb_f32_conv_dgrad")
            mark(x)
        check(L.vtb_f32_conv_wgrad(C.byref(geom), dybase, cout, abase + x.byte_offset(), x.ld, wsbase,
                                   pgrads[op.pidx].data_ptr(), op.cin_real, 0, st), "vtb_f32_conv_wgrad")
        self._residual_grad(op, gp, gld, is_init, mark, st)

    def _grad_stream_ctx(self):
        """Stream whose completion implies "the gradients reported so far are final" (side stream when wgrads run there)."""
        import contextlib

        return torch.cuda.stream(self._side) if self._side is not None else contextlib.nullcontext()

    def _launch_wgrad(self, launch, st: int) -> None:
        """Enqueue a weight-gradient GEMM (`launch(stream_handle)`) for the dy that the main stream has just produced:
        on the main stream, or - two-stream mode - on the side stream, ordered after the dy producer; the dy buffer is
        handed back to the rotation once this wgrad has read it."""
        if self._side is None:
            launch(st)
            return
        ready = torch.cuda.Event()
        ready.record()
        self._side.wait_event(ready)
        launch(self._side.cuda_stream)
        done = torch.cuda.Event()
        done.record(self._side)
        self._dy_free[self._dy_slot] = done

    def _residual_grad(self, op: ConvOp, gp, gld, is_init, mark, st) -> None:
        """Gradient of the post-activation residual add (darknet.py:28): identity into `res` unless it is aliased."""
        out, res = op.out, op.residual
        if res is None:
            return
        rv = res
        while rv.grad_alias is not None:
            rv = rv.grad_alias
        ov = out
        while ov.grad_alias is not None:
            ov = ov.grad_alias
        if rv is not ov:
            check(self.fn_grad_add(gp(res), gld(res), gp(out), gld(out), out.pixels, out.c, int(is_init(res)), st),
                  "vtb_grad_add")
            mark(res)
        # aliased: the gradient of `res` already sits in out's gradient memory (initialised by definition)

    # -- backward
    def backward(self, run: Run, gouts):
        g, L = self.g, self.L
        if not g.need_grad:
            raise RuntimeError("this plan was built without gradient support")
        st = self._stream()
        dev = run.act.device
        abase, sbase = run.act.data_ptr(), (run.stat.data_ptr() + 255) // 256 * 256
        run.bwd_stats_done = set()   # units whose BatchNorm-backward sums a consumer's dgrad epilogue has produced
        gact = torch.empty(g.grad_bytes, dtype=torch.uint8, device=dev)  # mirrors the "act" prefix of the arena
        gbase = gact.data_ptr()
        dy_stride = _round_up(g.dy_bytes, 1024)
        dyraw = torch.empty(self._dy_slots * dy_stride + 1024, dtype=torch.uint8, device=dev)
        dybase0 = (dyraw.data_ptr() + 1023) // 1024 * 1024
        dybase = dybase0
        self._dy_free = [None] * self._dy_slots   # per dy buffer: event after which its last wgrad reader is done
        dy_turn = 0
        ws = torch.empty(g.ws_bytes + 1024, dtype=torch.uint8, device=dev)
        wsbase = (ws.data_ptr() + 1023) // 1024 * 1024
        # parameter gradients: fresh fp32 tensors handed to autograd, or (grad_sink mode, used by parallel.Trainer)
        # written straight into the existing fp32 `.grad` storage — e.g. views of one flat all-reduce buffer —
        # which skips autograd's per-parameter accumulate kernels.  Sink mode OVERWRITES .grad.
        pgrads, direct = [], []
        for p in g.params:
            gr = p.grad if self.grad_sink else None
            if gr is not None and gr.dtype == torch.float32 and gr.is_contiguous() and gr.device == dev:
                pgrads.append(gr); direct.append(True)
            else:
                pgrads.append(torch.empty_like(p, dtype=torch.float32, memory_format=torch.contiguous_format))
                direct.append(False)
        world = self.dist.world if (self.dist is not None and self.dist.sync_bn) else 1
        # every parameter gradient of this backward OVERWRITES the caller's .grad storage (parallel.Trainer then skips
        # zeroing its flat gradient buffer)
        self.last_all_direct = bool(self.grad_sink and all(direct))

        # which gradient memory is already valid: buffer idx -> list of (c0, c1)
        init: dict[int, list[tuple[int, int]]] = {}

        def gview(t: TView) -> TView:
            while t.grad_alias is not None:
                t = t.grad_alias
            return t

        def is_init(t: TView) -> bool:
            t = gview(t)
            return any(a <= t.coff and t.coff + t.c <= b for a, b in init.get(t.buf.idx, ()))

        def mark(t: TView) -> None:
            t = gview(t)
            init.setdefault(t.buf.idx, []).append((t.coff, t.coff + t.c))

        def gp(t: TView) -> int:
            return gbase + gview(t).byte_offset()

        def gld(t: TView) -> int:
            return gview(t).ld

        # seed with the incoming feature-map gradients
        for t, go in zip(g.outputs, gouts):
            if go is None:
                continue
            go = go.detach()
            if go.dtype != self.tdtype:
                go = go.to(self.tdtype)
            go = go.contiguous(memory_format=torch.channels_last)
            check(self.fn_grad_add(gp(t), gld(t), go.data_ptr(), t.c, t.pixels, t.c, int(is_init(t)), st), "vtb_grad_add")
            mark(t)

        def next_dy() -> int:
            """dy buffer for the next BatchNorm backward (round robin; waits for the side-stream wgrad that read it last)."""
            nonlocal dy_turn
            slot = dy_turn % self._dy_slots
            dy_turn += 1
            self._dy_slot = slot
            ev = self._dy_free[slot]
            if ev is not None:
                torch.cuda.current_stream(dev).wait_event(ev)
                self._dy_free[slot] = None
            return dybase0 + slot * dy_stride

        pending: list[tuple[TView, TView]] = []
        # gradient all-reduce overlap: only meaningful when gradients land directly in the caller's flat buffer
        # (single-process plans report too when the caller registered a hook: parallel.Trainer steps the optimizer per bucket)
        hook = self.dist.on_grads_ready if self.dist is not None else self.on_grads_ready
        ready_cb = hook if (self.grad_sink and all(direct)) else None

        def flush_pending(force_for: Optional[TView]) -> None:
            for item in list(pending):
                rr, src = item
                if is_init(rr) or rr is force_for:
                    check(self.fn_grad_add(gp(rr), gld(rr), gp(src), gld(src), src.pixels, src.c, int(is_init(rr)), st),
                          "vtb_grad_add")
                    mark(rr)
                    pending.remove(item)

        for i in range(len(g.ops) - 1, -1, -1):
            op = g.ops[i]
            if op.kind == "conv" and op.pair is not None:
                continue   # handled together with its partner (visited just before, in reverse order)
            if op.kind == "conv" and op.pair_of is not None:
                op_a = op.pair_of
                flush_pending(op.out)
                flush_pending(op_a.out)
                have = [is_init(o.out) for o in (op_a, op)]
                if any(have):
                    for o, h in zip((op_a, op), have):
                        if not h:   # one branch without gradient: it contributes zeros to the shared dy buffer
                            tg = gview(o.out)
                            self.view_tensor(gact, tg).zero_()
                            mark(o.out)
                    self._conv_backward_pair(op_a, abase, sbase, gp, gld, is_init, mark, next_dy(), wsbase, pgrads, run, st, world)
                else:
                    for o in (op_a, op):
                        for j in range(3):
                            pgrads[o.pidx + j].zero_()
                if ready_cb is not None:
                    with self._grad_stream_ctx():
                        ready_cb(g.params[op.pidx : op.pidx + 3])
                        ready_cb(g.params[op_a.pidx : op_a.pidx + 3])
                flush_pending(False)
                continue
            flush_pending(op.out)
            if not is_init(op.out):
                # no gradient reaches this op: its parameters get zero gradient
                if op.kind in ("conv", "ese"):
                    n_p = op.nparams if op.kind == "conv" else 2
                    for j in range(n_p):
                        pgrads[op.pidx + j].zero_()
                    if ready_cb is not None:
                        ready_cb(g.params[op.pidx : op.pidx + n_p])
                continue
            if op.kind == "conv":
                bwd = self._conv_backward_f32 if self.f32 else self._conv_backward
                bwd(op, abase, sbase, gp, gld, is_init, mark, next_dy(), wsbase, pgrads, run, st, world)
                if ready_cb is not None:
                    with self._grad_stream_ctx():   # the all-reduce must also wait for the side-stream wgrad
                        ready_cb(g.params[op.pidx : op.pidx + op.nparams])
            elif op.kind == "copy":
                xx, oo = op.x, op.out
                if not (xx.is_input and not run.x_requires_grad):
                    check(self.fn_grad_add(gp(xx), gld(xx), gp(oo), gld(oo), xx.pixels, xx.c, int(is_init(xx)), st),
                          "vtb_grad_add(copy)")
                    mark(xx)
            elif op.kind == "pool":
                xx, oo = op.x, op.out
                if not (xx.is_input and not run.x_requires_grad):
                    check(self.fn_pool_bwd(abase + xx.byte_offset(), xx.ld, xx.n, xx.h, xx.w, xx.c, gp(oo), gld(oo),
                                               gp(xx), gld(xx), int(is_init(xx)),
                                               abase + op.idx.byte_offset() if op.idx is not None else 0, st),
                          "vtb_maxpool3s2_bwd")
                    mark(xx)
            elif op.kind == "ese":
                xx, oo, rr = op.x, op.out, op.residual
                lin = op.mod.linear
                check(self.fn_ese_bwd(abase + xx.byte_offset(), xx.ld, xx.n, xx.h * xx.w, xx.c, lin.weight.data_ptr(),
                                    sbase + 4 * op.st["pool"], sbase + 4 * op.st["z"], sbase + 4 * op.st["gate"],
                                    gp(oo), gld(oo), gp(xx), gld(xx), int(is_init(xx)),
                                    pgrads[op.pidx].data_ptr(), pgrads[op.pidx + 1].data_ptr(), 0,
                                    sbase + 4 * op.st["scratch"], st), "vtb_ese_bwd")
                mark(xx)
                if ready_cb is not None:
                    ready_cb(g.params[op.pidx : op.pidx + 2])
                if rr is not None and gview(rr) is not gview(oo):
                    # `rr` may be slice 0 of a concat buffer whose gradient is about to be (over)written as a whole
                    # by out_conv's dgrad: defer the identity contribution until that memory is initialised
                    pending.append((rr, oo))
            flush_pending(False)

        if self._side is not None:
            torch.cuda.current_stream(dev).wait_stream(self._side)   # every weight gradient is final past this point
        gx = None
        if run.x_requires_grad:
            t = g.input
            flush_pending(t)
            if is_init(t):
                gin = self.view_tensor(gact, t)[:, : run.x_shape[1]]
                gx = gin.to(run.x_dtype).contiguous()
            else:
                gx = torch.zeros(run.x_shape, dtype=run.x_dtype, device=dev)
        # cast parameter gradients to the parameter dtype if a user keeps non-fp32 masters
        out_grads = [None if d else (pg if pg.dtype == p.dtype else pg.to(p.dtype))
                     for pg, p, d in zip(pgrads, g.params, direct)]
        return gx, out_grads

    def _conv_backward(self, op: ConvOp, abase, sbase, gp, gld, is_init, mark, dybase, wsbase, pgrads, run, st, world):
        g, L = self.g, self.L
        geom, cout = op.geom, op.geom.cout
        x, y, out, res = op.x, op.y, op.out, op.residual
        f = lambda name: sbase + 4 * op.st[name]
        dout_p, dout_ld = gp(out), gld(out)
        peer_sync = self.dist.sync if (world > 1 and self.dist is not None) else None
        if op.bias:
            # no normalisation, no activation: dy IS the output gradient; the bias gradient is its per-channel sum (the
            # reduce kernel of the BatchNorm backward with the mask off - its second sum is not used)
            check(L.vtb_bn_bwd_reduce(dout_p, dout_ld, dout_p, dout_ld, out.pixels, cout, f("scale"), f("scale"), f("scale"),
                                      f("scale"), 0, f("partial_b"), st), "vtb_bn_bwd_reduce(bias)")
            check(L.vtb_bn_bwd_finalize(f("partial_b"), op.rows_b, 0, 0, 1.0, cout, 0, pgrads[op.pidx + 1].data_ptr(), 0,
                                        f("coef"), 0, st), "vtb_bn_bwd_finalize(bias)")
            self._conv_backward_gemms(op, abase, gp, gld, is_init, mark, dout_p, wsbase, pgrads, run, st, lddy=dout_ld)
            return
        if id(op) in run.bwd_stats_done:
            # the dgrad that completed this unit's output gradient already reduced (and exchanged) the sums: coef, dgamma
            # and dbeta are final - BatchNorm+ReLU backward is one apply pass
            check(L.vtb_bn_bwd_apply(dout_p, dout_ld, abase + y.byte_offset(), y.ld, out.pixels, cout, f("scale"),
                                     f("shift"), f("mean"), f("invstd"), int(op.relu), f("coef"), dybase, cout, st),
                  "vtb_bn_bwd_apply")
            self._conv_backward_gemms(op, abase, gp, gld, is_init, mark, dybase, wsbase, pgrads, run, st)
            return
        if op.batch_stats and (world == 1 or peer_sync is not None):
            # reduce -> finalize (-> SyncBN exchange over NVLink) -> apply in one cooperative launch
            check(L.vtb_bn_bwd_fused(dout_p, dout_ld, abase + y.byte_offset(), y.ld, out.pixels, cout, f("scale"),
                                     f("shift"), f("mean"), f("invstd"), int(op.relu), float(out.pixels) * world,
                                     f("partial_b"), pgrads[op.pidx + 1].data_ptr(), pgrads[op.pidx + 2].data_ptr(), 0,
                                     self.tickets.data_ptr() + 1024, dybase, cout,
                                     C.byref(peer_sync) if peer_sync is not None else None, st), "vtb_bn_bwd_fused")
            self._conv_backward_gemms(op, abase, gp, gld, is_init, mark, dybase, wsbase, pgrads, run, st)
            return
        check(L.vtb_bn_bwd_reduce(dout_p, dout_ld, abase + y.byte_offset(), y.ld, out.pixels, cout, f("scale"),
                                  f("shift"), f("mean"), f("invstd"), int(op.relu), f("partial_b"), st),
              "vtb_bn_bwd_reduce")
        dgamma, dbeta = pgrads[op.pidx + 1].data_ptr(), pgrads[op.pidx + 2].data_ptr()
        count = float(out.pixels)
        if op.batch_stats and world > 1:
            check(L.vtb_bn_bwd_finalize(f("partial_b"), op.rows_b, 0, 0, count, cout, 0, 0, 0, 0, f("lsums_b"), st),
                  "vtb_bn_bwd_finalize(local)")
            sums = run.stat_view_f64(op.st["sums_b"], cout * 2, sbase)
            sums.copy_(run.stat_view_f64(op.st["lsums_b"], cout * 2, sbase))
            self.dist.all_reduce_(sums)
            check(L.vtb_bn_bwd_finalize(0, 0, f("sums_b"), f("lsums_b"), count * world, cout, dgamma, dbeta, 0,
                                        f("coef"), 0, st), "vtb_bn_bwd_finalize")
        else:
            check(L.vtb_bn_bwd_finalize(f("partial_b"), op.rows_b, 0, 0, count, cout, dgamma, dbeta, 0, f("coef"), 0,
                                        st), "vtb_bn_bwd_finalize")
            if not op.batch_stats:
                # frozen statistics: mean/var are constants, so dy = scale * dz (no mean-subtraction terms)
                run.stat_view_f32(op.st["coef"], cout * 2, sbase).zero_()
        check(L.vtb_bn_bwd_apply(dout_p, dout_ld, abase + y.byte_offset(), y.ld, out.pixels, cout, f("scale"),
                                 f("shift"), f("mean"), f("invstd"), int(op.relu), f("coef"), dybase, cout, st),
              "vtb_bn_bwd_apply")
        self._conv_backward_gemms(op, abase, gp, gld, is_init, mark, dybase, wsbase, pgrads, run, st)

    def _conv_backward_gemms(self, op: ConvOp, abase, gp, gld, is_init, mark, dybase, wsbase, pgrads, run, st, lddy=None):
        L = self.L
        geom, cout = op.geom, op.geom.cout
        lddy = cout if lddy is None else lddy
        x, out, res = op.x, op.out, op.residual
        dout_p, dout_ld = gp(out), gld(out)
        _, wd = self._packed(op)
        if op.col is not None:
            # gathered-operand stem: the GEMM yields [cout][taps * c] in (tap, ci) column order; permute to OIHW
            taps, creal = op.col
            dw_col = (run.stat.data_ptr() + 255) // 256 * 256 + 4 * op.st["dw_col"]

            def launch(s_):
                check(L.vtb_conv_wgrad(C.byref(geom), dybase, cout, abase + x.byte_offset(), x.ld, wsbase, dw_col,
                                       op.cin_real, 0, s_), "vtb_conv_wgrad(stem)")
                check(L.vtb_dw_from_col(dw_col, cout, creal, taps, pgrads[op.pidx].data_ptr(), 0, s_), "vtb_dw_from_col")

            self._launch_wgrad(launch, st)
        else:
            self._launch_wgrad(lambda s_: check(L.vtb_conv_wgrad(C.byref(geom), dybase, lddy, abase + x.byte_offset(), x.ld,
                                                                 wsbase, pgrads[op.pidx].data_ptr(), op.cin_real, 0, s_),
                                                "vtb_conv_wgrad"), st)
        if not (x.is_input and not run.x_requires_grad):
            self._dgrad(op, geom, dybase, lddy, wd, abase, gp, gld, is_init, pgrads, run, st)
            mark(x)
        self._residual_grad(op, gp, gld, is_init, mark, st)

    def _dgrad_s2_workspace(self, op: ConvOp, geom, lddx: int):
        """Workspace of the merged stride-2 dgrad of `op` (vtb_conv_dgrad_s2), or None when the layer does not qualify:
        3x3 / stride 2 / pad 1, even H and W, cin <= 64, dense input gradient (pitch == channels)."""
        key = id(op)
        if key not in self._s2_ws:
            ws = None
            if geom.k == 3 and geom.stride == 2 and geom.pad == 1 and lddx == geom.cin:
                nbytes = int(self.L.vtb_conv_dgrad_s2_workspace_bytes(C.byref(geom)))
                if nbytes > 0:
                    ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._s2_ws[key] = ws
        return self._s2_ws[key]

    def _dgrad(self, op: ConvOp, geom, dybase: int, lddy: int, wd, abase: int, gp, gld, is_init, pgrads, run, st) -> None:
        """dgrad of `op` (or of a side-by-side pair, `geom` = the pair geometry) into its input; when the plan marked this
        write as the last contribution to the producers' output gradient it also reduces their BatchNorm-backward sums."""
        L = self.L
        x = op.x
        world = self.dist.world if (self.dist is not None and self.dist.sync_bn) else 1
        peer_sync = self.dist.sync if (world > 1 and self.dist is not None) else None
        if op.dgrad_bn is None or not (world == 1 or peer_sync is not None):
            ws2 = self._dgrad_s2_workspace(op, geom, gld(x)) if self.dgrad_s2 else None
            if ws2 is not None:
                # 3x3 stride-2 layer with few channels: one dense GEMM over 2x2 super-pixels (dy read once, not 4 times)
                check(L.vtb_conv_dgrad_s2(C.byref(geom), dybase, lddy, wd.data_ptr(), ws2.data_ptr(), gp(x), gld(x),
                                          int(is_init(x)), st), "vtb_conv_dgrad_s2")
                return
            check(L.vtb_conv_dgrad(C.byref(geom), dybase, lddy, wd.data_ptr(), gp(x), gld(x), int(is_init(x)), st),
                  "vtb_conv_dgrad")
            return
        prods, split = op.dgrad_bn
        sbase = (run.stat.data_ptr() + 255) // 256 * 256
        bn = VtbDgradBn()
        bn.split = split
        for l, q in enumerate(prods):
            fq = lambda name, q=q: sbase + 4 * q.st[name]
            lay = bn.layer[l]
            lay.y, lay.ldy = abase + q.y.byte_offset(), q.y.ld
            lay.scale, lay.shift, lay.mean, lay.invstd = fq("scale"), fq("shift"), fq("mean"), fq("invstd")
            lay.relu = int(q.relu)
            lay.dgamma, lay.dbeta = pgrads[q.pidx + 1].data_ptr(), pgrads[q.pidx + 2].data_ptr()
            lay.coef = fq("coef")
        bn.count = float(x.pixels) * world
        bn.partial = sbase + 4 * op.st["partial_d"]
        bn.tickets = self.tickets.data_ptr()
        bn.sync = C.addressof(peer_sync) if peer_sync is not None else None
        check(L.vtb_conv_dgrad_bn(C.byref(geom), dybase, lddy, wd.data_ptr(), gp(x), gld(x), int(is_init(x)),
                                  C.byref(bn), st), "vtb_conv_dgrad_bn")
        for q in prods:
            run.bwd_stats_done.add(id(q))


def _stat_view_f32(self, off_floats: int, n: int, sbase: int) -> torch.Tensor:
    start = (sbase - self.stat.data_ptr()) // 4 + off_floats
    return self.stat[start : start + n]


def _stat_view_f64(self, off_floats: int, n: int, sbase: int) -> torch.Tensor:
    start = (sbase - self.stat.data_ptr()) // 4 + off_floats
    return self.stat[start : start + 2 * n].view(torch.float64)


Run.stat_view_f32 = _stat_view_f32
Run.stat_view_f64 = _stat_view_f64


# ----------------------------------------------------------------------------------------------------
# autograd glue + plan cache
# ----------------------------------------------------------------------------------------------------
class _NativeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner: Runner, x: torch.Tensor, *params):
        outs, run = runner.forward(x)
        ctx.runner, ctx.run = runner, run
        ctx.set_materialize_grads(False)  # unused feature maps get no zero-filled gradient tensors
        return tuple(outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gouts):
        gx, pgrads = ctx.runner.backward(ctx.run, gouts)
        ctx.run = None
        return (None, gx, *pgrads)


def run_native(module: nn.Module, x: torch.Tensor) -> list[torch.Tensor]:
    """Execute `module` (anything with ``_emit(graph, tview) -> tview | list[tview]``) on a CUDA tensor."""
    if not x.is_cuda:
        raise ValueError("run_native expects a CUDA tensor")
    if x.dim() != 4:
        raise ValueError(f"expected an (N, C, H, W) tensor, got shape {tuple(x.shape)}")
    need_grad = torch.is_grad_enabled() and (
        x.requires_grad or any(p.requires_grad for p in module.parameters())
    )
    f32 = _resolve_f32()
    dcfg = module.__dict__.get("_vtb_dist")
    # batch statistics are decided per BatchNorm layer (reference: nn.BatchNorm2d at components.py:36 looks at ITS OWN
    # .training): the plan is "training" as soon as one layer uses them, and the per-layer flags are part of the key
    bn_sig = _bn_signature(module)
    training = any(flag for flag, _ in bn_sig) if bn_sig else module.training
    # sibling units run as one convolution when BatchNorm is finalised inside the conv kernel (VTB_PAIR=0: A/B switch)
    pair_ok = (_os.environ.get("VTB_PAIR", "1") == "1" and training and not f32
               and (dcfg is None or not dcfg.sync_bn or dcfg.world == 1 or dcfg.sync is not None))
    # the image's first convolution as a 1x1 GEMM over a gathered operand, unless the image itself needs a gradient
    col_stem = (_os.environ.get("VTB_COL_STEM", "1") == "1" and not f32
                and not (torch.is_grad_enabled() and x.requires_grad))
    key = (tuple(x.shape), bn_sig, need_grad, x.device.index, f32, pair_ok,
           dcfg is not None and dcfg.world > 1, col_stem, _os.environ.get("VTB_DGRAD_BN", "0"))
    plans = module.__dict__.setdefault("_vtb_plans", {})
    runner = plans.get(key)
    if runner is None:
        with torch.cuda.device(x.device):
            g = Graph(training, need_grad, f32, pair_ok, col_stem)
            t_in = g.input_image(*x.shape)
            outs = module._emit(g, t_in)
            if isinstance(outs, TView):
                outs = [outs]
            for t in outs:
                g.mark_output(t)
            g.finalize()
            runner = Runner(g, x.device, dcfg)
        plans[key] = runner
        if len(plans) > 8:  # bound the cache (each plan only holds metadata)
            plans.pop(next(iter(plans)))
    runner.dist = module.__dict__.get("_vtb_dist")
    runner.grad_sink = bool(module.__dict__.get("_vtb_grad_sink", False))
    runner.on_grads_ready = module.__dict__.get("_vtb_on_grads_ready")
    with torch.cuda.device(x.device):
        if need_grad:
            outs = _NativeFn.apply(runner, x, *runner.g.params)
        else:
            outs, _ = runner.forward(x)
    return list(outs)
