"""Dry run of the executor on the CPU: the planner + Runner drive a RECORDING stand-in for libvtb_b200.so (host-only
queries go to the real library), so every Python path of forward / backward is executed without a GPU and the launch list
can be checked for what a kernel would have touched:

  * every activation / gradient view handed to a kernel (pointer, pixel pitch, pixels, channels) lies inside one of the
    buffers the executor allocated for the call;
  * a read-modify-write gradient fan-in (dgrad / grad_add / pool / eSE with accumulate = 1) only targets gradient memory
    that an earlier launch of the same backward initialised, and every gradient view that is read was written before;
  * launch counts per entry point match the model (one fprop per unit or sibling pair, one BatchNorm backward per unit...).

No compute call reaches the real library here (there is no GPU); numerical parity is tests/test_gpu_*.py."""
import ctypes as C
from collections import Counter
from unittest import mock

import pytest
import torch

from vision_toolbox_b200 import _lib, backbones, engine
from vision_toolbox_b200.backbones.darknet import CSPDarknetStage

HOST_ONLY = {"vtb_last_error", "vtb_conv_stats_rows", "vtb_bn_bwd_rows", "vtb_bn_bwd_fused_rows", "vtb_conv_out_hw",
             "vtb_conv_wgrad_workspace_bytes", "vtb_f32_conv_wgrad_workspace_bytes", "vtb_f32_bn_rows",
             "vtb_pack_job_blocks", "vtb_launch_count", "vtb_version", "vtb_num_sms", "vtb_bn_sync_buffer_bytes",
             "vtb_conv_dgrad_stats_rows", "vtb_conv_dgrad_panel_w", "vtb_conv_tiling_info", "vtb_sgd_job_blocks",
             "vtb_conv_dgrad_s2_workspace_bytes"}


class RecordingLib:
    def __init__(self):
        self.real = _lib.lib()
        self.calls = []

    def __getattr__(self, name):
        if name in HOST_ONLY:
            return getattr(self.real, name)
        assert name in _lib.SIGNATURES, f"{name} is not declared in include/vtb.h"

        def call(*args):
            assert len(args) == len(_lib.SIGNATURES[name][1]), (name, len(args))
            self.calls.append((name, args))
            return 0

        return call


class Allocations:
    """Every torch.empty / torch.zeros the executor performs during the call (kept alive: no address reuse)."""

    def __init__(self):
        self.spans = []
        self.keep = []

    def add(self, t: torch.Tensor):
        self.keep.append(t)
        self.spans.append((t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()))

    def inside(self, lo: int, hi: int) -> bool:
        return any(a <= lo and hi <= b for a, b in self.spans)


def _geom(g):
    g = g._obj if hasattr(g, "_obj") else g
    ho = (g.h + 2 * g.pad - g.k) // g.stride + 1
    wo = (g.w + 2 * g.pad - g.k) // g.stride + 1
    return g, g.n * g.h * g.w, g.n * ho * wo


def _views(name, a, es):
    """(pointer, pitch, pixels, channels, mode) of every activation / gradient view of a call; mode r / w / rw."""
    f32 = name.startswith("vtb_f32_")
    base = name.replace("vtb_f32_", "vtb_")
    if base == "vtb_im2col_input":
        n, h, w, k, st, pad = a[1], a[3], a[4], a[5], a[6], a[7]
        return [(a[8], a[9], n * ((h + 2 * pad - k) // st + 1) * ((w + 2 * pad - k) // st + 1), a[9], "w")]
    if base == "vtb_nchw_to_nhwc":
        return [(a[5], a[6], a[1] * a[3] * a[4], a[6], "w")]
    if base in ("vtb_conv_fprop", "vtb_conv_fprop_bn"):
        g, pin, pout = _geom(a[0])
        y = (a[5], a[6]) if f32 else (a[4], a[5])
        v = [(a[1], a[2], pin, g.cin, "r"), (y[0], y[1], pout, g.cout, "w")]
        if base == "vtb_conv_fprop" and not f32 and a[10]:      # eval-mode fused epilogue: + residual
            v.append((a[10], a[11], pout, g.cout, "r"))
        if base == "vtb_conv_fprop_bn" and not f32:
            bn = a[7]._obj if hasattr(a[7], "_obj") else a[7]
            if getattr(bn, "act_out", None):                    # fused normalise: the unit's activation (+ residual)
                v.append((bn.act_out, bn.act_ld, pout, g.cout, "w"))
                if bn.act_residual:
                    v.append((bn.act_residual, bn.act_ldr, pout, g.cout, "r"))
        return v
    if base == "vtb_conv_dgrad":
        g, pin, pout = _geom(a[0])
        dx, ld, acc = (a[5], a[6], a[7]) if f32 else (a[4], a[5], a[6])
        return [(a[1], a[2], pout, g.cout, "r"), (dx, ld, pin, g.cin, "rw" if acc else "w")]
    if base == "vtb_conv_dgrad_s2":
        g, pin, pout = _geom(a[0])
        return [(a[1], a[2], pout, g.cout, "r"), (a[5], a[6], pin, g.cin, "rw" if a[7] else "w")]
    if base == "vtb_conv_dgrad_bn":
        g, pin, pout = _geom(a[0])
        bn = a[7]._obj if hasattr(a[7], "_obj") else a[7]
        v = [(a[1], a[2], pout, g.cout, "r"), (a[4], a[5], pin, g.cin, "rw" if a[6] else "w")]
        bounds = [(0, g.cin)] if bn.split == 0 else [(0, bn.split), (bn.split, g.cin)]
        for lay, (c0, c1) in zip(bn.layer, bounds):     # the producers' raw conv outputs, on dx's pixel lattice
            v.append((lay.y, lay.ldy, pin, c1 - c0, "r"))
        return v
    if base in ("vtb_conv_wgrad", "vtb_conv_wgrad_pair"):
        g, pin, pout = _geom(a[0])
        return [(a[1], a[2], pout, g.cout, "r"), (a[3], a[4], pin, g.cin, "r")]
    if base == "vtb_bn_act":
        if f32:
            v = [(a[0], a[1], a[2], a[3], "r"), (a[11], a[12], a[2], a[3], "w")]
            return v + ([(a[9], a[10], a[2], a[3], "r")] if a[9] else [])
        v = [(a[0], a[1], a[2], a[3], "r"), (a[9], a[10], a[2], a[3], "w")]
        return v + ([(a[7], a[8], a[2], a[3], "r")] if a[7] else [])
    if base == "vtb_bn_bwd_fused":
        return [(a[0], a[1], a[4], a[5], "r"), (a[2], a[3], a[4], a[5], "r"), (a[17], a[18], a[4], a[5], "w")]
    if base in ("vtb_bn_bwd_reduce",):
        return [(a[0], a[1], a[4], a[5], "r"), (a[2], a[3], a[4], a[5], "r")]
    if base == "vtb_bn_bwd_apply":
        return [(a[0], a[1], a[4], a[5], "r"), (a[2], a[3], a[4], a[5], "r"), (a[12], a[13], a[4], a[5], "w")]
    if base == "vtb_grad_add":
        return [(a[2], a[3], a[4], a[5], "r"), (a[0], a[1], a[4], a[5], "rw" if a[6] else "w")]
    if base == "vtb_maxpool3s2_fwd":
        n, h, w, c = a[2], a[3], a[4], a[5]
        return [(a[0], a[1], n * h * w, c, "r"), (a[6], a[7], n * ((h - 1) // 2 + 1) * ((w - 1) // 2 + 1), c, "w")]
    if base == "vtb_maxpool3s2_bwd":
        n, h, w, c = a[2], a[3], a[4], a[5]
        po = n * ((h - 1) // 2 + 1) * ((w - 1) // 2 + 1)
        return [(a[6], a[7], po, c, "r"), (a[8], a[9], n * h * w, c, "rw" if a[10] else "w")]
    if base == "vtb_ese_fwd":
        v = [(a[0], a[1], a[2] * a[3], a[4], "r"), (a[9], a[10], a[2] * a[3], a[4], "w")]
        return v + ([(a[7], a[8], a[2] * a[3], a[4], "r")] if a[7] else [])
    if base == "vtb_ese_bwd":
        return [(a[0], a[1], a[2] * a[3], a[4], "r"), (a[9], a[10], a[2] * a[3], a[4], "r"),
                (a[11], a[12], a[2] * a[3], a[4], "rw" if a[13] else "w")]
    return []


class FakeDist:
    """What engine.DistConfig looks like to the Runner when SyncBN goes through one all-reduce per exchange."""

    def __init__(self, world=2):
        self.world, self.rank, self.sync_bn, self.sync, self.on_grads_ready = world, 0, True, None, None
        self.reduced = 0

    def all_reduce_(self, t):
        assert t.dtype == torch.float64 and t.numel() % 2 == 0
        self.reduced += 1


def _dry_run(model, shape, f32, need_input_grad=False, dist=None, training=True, need_grad=True):
    model.train(training)
    g = engine.Graph(training, need_grad, f32, pair_ok=dist is None, col_stem=not need_input_grad)
    outs = model._emit(g, g.input_image(*shape))
    for t in ([outs] if isinstance(outs, engine.TView) else outs):
        g.mark_output(t)
    g.finalize()
    lib = RecordingLib()
    with mock.patch.object(engine._lib, "lib", return_value=lib):   # the Runner binds its entry points when it is built
        runner = engine.Runner(g, torch.device("cpu"))
    assert runner.L is lib
    runner.dist = dist
    runner._stream = lambda: 0
    allocs = Allocations()
    real_empty, real_zeros = torch.empty, torch.zeros

    def rec(fn):
        def wrapped(*a, **k):
            t = fn(*a, **k)
            allocs.add(t)
            return t
        return wrapped

    x = torch.rand(shape, requires_grad=need_input_grad)
    with mock.patch.object(engine.torch, "empty", rec(real_empty)), mock.patch.object(engine.torch, "zeros", rec(real_zeros)):
        outs_t, run = runner.forward(x)
        n_fwd = len(lib.calls)
        if not need_grad:
            for i, (name, a) in enumerate(lib.calls):
                for ptr, ld, pixels, c, mode in _views(name, a, 4 if f32 else 2):
                    assert allocs.inside(ptr, ptr + ((pixels - 1) * ld + c) * (4 if f32 else 2)), (i, name)
            return g, lib.calls, n_fwd
        tdt = torch.float32 if f32 else torch.bfloat16
        gouts = [torch.ones(o.shape, dtype=tdt).contiguous(memory_format=torch.channels_last) for o in outs_t]
        for t in gouts:
            allocs.add(t)
        gx, pgrads = runner.backward(run, gouts)
    assert (gx is not None) == need_input_grad
    assert len(pgrads) == len(g.params) and all(p is not None for p in pgrads)
    es = 4 if f32 else 2
    written = []   # (lo, hi of the first pixel row, pitch) of every view some launch has written

    def covered(ptr, ld, c):
        """[ptr, ptr + c) of the first pixel row is covered by the union of written views of the same pitch (a concat
        buffer is written slice by slice and read as a whole)."""
        need, end = ptr, ptr + c * es
        for lo, hi in sorted((lo, hi) for lo, hi, pitch in written if pitch == ld and hi > ptr and lo < end):
            if lo > need:
                return False
            need = max(need, hi)
            if need >= end:
                return True
        return need >= end

    for t in gouts:   # the incoming feature-map gradients are inputs of the backward, not products of a launch
        written.append((t.data_ptr(), t.data_ptr() + t.shape[1] * es, t.shape[1]))
    for i, (name, a) in enumerate(lib.calls):
        for ptr, ld, pixels, c, mode in _views(name, a, es):
            assert ptr and ld >= c and pixels > 0, (name, ptr, ld, c)
            assert allocs.inside(ptr, ptr + ((pixels - 1) * ld + c) * es), (i, name, "view outside every allocation")
            if i >= n_fwd and "r" in mode and not name.endswith("_fwd"):
                # backward reads: activations were written by the forward, gradients by an earlier backward launch
                assert covered(ptr, ld, c), (i, name, mode, "reads memory no launch has written")
            if "w" in mode:
                written.append((ptr, ptr + c * es, ld))
    return g, lib.calls, n_fwd


CASES = {
    "cspdarknet53": (lambda: backbones.cspdarknet53(), (2, 3, 64, 64)),
    "darknet19": (lambda: backbones.darknet19(), (2, 3, 64, 64)),
    "darknet_yolov5s": (lambda: backbones.darknet_yolov5s(), (2, 3, 64, 64)),
    "vovnet39_ese": (lambda: backbones.vovnet39_ese(), (2, 3, 64, 64)),
    "vovnet27_slim": (lambda: backbones.vovnet27_slim(), (2, 3, 64, 64)),
}


@pytest.mark.parametrize("f32", [False, True], ids=["bf16", "fp32"])
@pytest.mark.parametrize("name", list(CASES))
def test_launch_list_is_memory_safe(name, f32):
    build, shape = CASES[name]
    torch.manual_seed(0)
    _dry_run(build(), shape, f32)


def test_launch_list_with_image_gradient():
    _dry_run(backbones.cspdarknet53(), (2, 3, 64, 64), False, need_input_grad=True)
    _dry_run(backbones.vovnet27_slim(), (2, 3, 64, 64), True, need_input_grad=True)


@pytest.mark.parametrize("f32", [False, True], ids=["bf16", "fp32"])
def test_launch_list_with_all_reduce_syncbn(f32):
    """The fallback exchange (VTB_SYNCBN=nccl, and every fp32-mode multi-rank run): statistics leave the kernels as fp64
    sums, are all-reduced, and BatchNorm is finalised by separate launches - two exchanges per unit."""
    dist = FakeDist(world=2)
    g, calls, n_fwd = _dry_run(backbones.Darknet(16, [(1, 32), (2, 64)], CSPDarknetStage), (2, 3, 32, 32), f32, dist=dist)
    units = sum(op.kind == "conv" for op in g.ops)
    assert dist.reduced == 2 * units
    names = Counter(n for n, _ in calls)
    assert names["vtb_bn_finalize"] == units and names["vtb_bn_bwd_finalize"] == (units if f32 else 2 * units)
    assert names["vtb_conv_fprop_bn"] == 0 and names["vtb_bn_bwd_fused"] == 0


@pytest.mark.parametrize("f32", [False, True], ids=["bf16", "fp32"])
def test_eval_and_frozen_statistics_plans(f32):
    """eval() without gradients: one fused conv + affine + ReLU (+ residual) launch per unit in bf16 mode, conv + normalise
    in fp32 mode; eval() with gradients (frozen BatchNorm): the unfused backward."""
    m = backbones.Darknet(16, [(1, 32), (2, 64)], CSPDarknetStage)
    g, calls, _ = _dry_run(m, (2, 3, 32, 32), f32, training=False, need_grad=False)
    units = sum(op.kind == "conv" for op in g.ops)
    names = Counter(n for n, _ in calls)
    if f32:
        assert names["vtb_f32_conv_fprop"] == units == names["vtb_f32_bn_act"] and g.fused_eval is False
    else:
        assert names["vtb_conv_fprop"] == units == names["vtb_bn_eval_affine"] and names["vtb_bn_act"] == 0
        assert all(op.pair is None for op in g.ops if op.kind == "conv")        # no pairing outside training plans
    g, calls, n_fwd = _dry_run(m, (2, 3, 32, 32), f32, training=False, need_grad=True)
    bwd = Counter(n for n, _ in calls[n_fwd:])
    pre = "vtb_f32_" if f32 else "vtb_"
    assert bwd[pre + "bn_bwd_apply"] == units and bwd["vtb_bn_bwd_finalize"] == units and bwd["vtb_bn_bwd_fused"] == 0
    vov = backbones.VoVNet(32, [(1, 16, 2, 32), (2, 16, 3, 32)], ese=True)
    _dry_run(vov, (2, 3, 32, 32), f32, training=False, need_grad=False)
    _dry_run(vov, (2, 3, 32, 32), f32, training=False, need_grad=True)


def test_launch_counts_cspdarknet53(monkeypatch):
    """67 units: 5 sibling pairs run as one convolution each -> 62 fprop, 61 + 4 x 5 dgrad launches happen inside the
    library (stride-2 phases are one C call), the stem needs no dgrad."""
    g, calls, n_fwd = _dry_run(backbones.cspdarknet53(), (2, 3, 64, 64), False)
    bwd = Counter(n for n, _ in calls[n_fwd:])
    units = sum(op.kind == "conv" for op in g.ops)
    pairs = sum(op.kind == "conv" and op.pair is not None for op in g.ops)
    # default plan: one fused reduce + apply BatchNorm-backward kernel per unit, plain dgrads - the two few-channel
    # stride-2 layers (32 -> 64, 64 -> 128; 128 -> 256 and wider stay on the four-phase path) as one GEMM over 2x2 super-pixels
    assert bwd["vtb_bn_bwd_fused"] == units and bwd["vtb_conv_dgrad_bn"] == 0
    assert bwd["vtb_conv_dgrad"] + bwd["vtb_conv_dgrad_s2"] == units - pairs - 1 and bwd["vtb_conv_dgrad_s2"] == 2
    monkeypatch.setenv("VTB_DGRAD_S2_MERGED", "0")
    _, calls0, n_fwd0 = _dry_run(backbones.cspdarknet53(), (2, 3, 64, 64), False)
    bwd0 = Counter(n for n, _ in calls0[n_fwd0:])
    assert bwd0["vtb_conv_dgrad"] == units - pairs - 1 and bwd0["vtb_conv_dgrad_s2"] == 0
    monkeypatch.delenv("VTB_DGRAD_S2_MERGED")
    # fused normalise (VTB_FUSED_NORM=1): the normalise + ReLU (+ residual) pass of every single unit rides in its
    # convolution's launch; only the two units of each side-by-side pair keep their vtb_bn_act
    monkeypatch.setenv("VTB_FUSED_NORM", "1")
    g, calls, n_fwd = _dry_run(backbones.cspdarknet53(), (2, 3, 64, 64), False)
    fwd = Counter(n for n, _ in calls[:n_fwd])
    assert fwd["vtb_conv_fprop_bn"] == units - pairs and fwd["vtb_bn_act"] == 2 * pairs
    monkeypatch.setenv("VTB_FUSED_NORM", "0")
    # opt-in plan (VTB_DGRAD_BN=1): the BatchNorm-backward sums ride in the dgrad epilogues
    monkeypatch.setenv("VTB_DGRAD_BN", "1")
    g, calls, n_fwd = _dry_run(backbones.cspdarknet53(), (2, 3, 64, 64), False)
    fwd = Counter(n for n, _ in calls[:n_fwd])
    bwd = Counter(n for n, _ in calls[n_fwd:])
    units = sum(op.kind == "conv" for op in g.ops)
    pairs = sum(op.kind == "conv" and op.pair is not None for op in g.ops)
    assert (units, pairs) == (67, 5)
    assert fwd == Counter({"vtb_pack_weights": 1, "vtb_im2col_input": 1, "vtb_conv_fprop_bn": units - pairs,
                           "vtb_bn_act": units})
    # BatchNorm backward: every dgrad is the last contribution to its input's gradient and carries the (dz, dz*xhat) sums
    # of the unit(s) that produced it (the CSP concat buffer: two units per out_conv dgrad) -> one apply pass per unit;
    # only the last unit, whose gradient arrives from outside the plan, keeps the reduce + apply kernel
    assert bwd["vtb_bn_bwd_fused"] == 1 and bwd["vtb_bn_bwd_apply"] == units - 1
    assert bwd["vtb_conv_wgrad"] + bwd["vtb_conv_wgrad_pair"] == units - pairs and bwd["vtb_conv_wgrad_pair"] == pairs
    assert bwd["vtb_conv_dgrad_bn"] == units - pairs - 1 and bwd["vtb_conv_dgrad"] == 0   # every convolution but the stem
    two = sum(1 for n, a in calls[n_fwd:] if n == "vtb_conv_dgrad_bn" and a[7]._obj.split > 0)
    assert two == pairs                                       # the five out_conv dgrads serve conv1 | last block
    # stem weight gradient; one injection per feature map that received a gradient (the trainer only uses the last one)
    assert bwd["vtb_dw_from_col"] == 1 and bwd["vtb_grad_add"] == len(g.outputs) == 5
    # the residual of every DarknetBlock is aliased into its output's gradient memory: no extra fan-out copies
    assert set(bwd) == {"vtb_bn_bwd_fused", "vtb_bn_bwd_apply", "vtb_conv_wgrad", "vtb_conv_wgrad_pair",
                        "vtb_conv_dgrad_bn", "vtb_dw_from_col", "vtb_grad_add"}
