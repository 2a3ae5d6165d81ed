"""Host-side logic of the precision switch (bf16 tensor-core path | fp32 parity mode): no GPU needed."""
import pytest
import torch

import vision_toolbox_b200 as vtb
from vision_toolbox_b200 import engine
from vision_toolbox_b200.backbones import Darknet, VoVNet
from vision_toolbox_b200.backbones.darknet import CSPDarknetStage


def test_precision_api():
    assert vtb.get_precision() == "bf16"          # default: the tensor-core path
    with vtb.precision("fp32"):
        assert vtb.get_precision() == "fp32" and engine._resolve_f32()
        with vtb.precision("bf16"):
            assert not engine._resolve_f32()
        assert vtb.get_precision() == "fp32"
    assert vtb.get_precision() == "bf16"
    with pytest.raises(KeyError):
        vtb.set_precision("fp16")
    with vtb.precision("auto"):                   # what the reference would compute for the same call
        assert engine._resolve_f32()
        if torch.cuda.is_available():             # torch disables CUDA autocast on a machine without CUDA
            with torch.autocast("cuda", dtype=torch.bfloat16):
                assert not engine._resolve_f32()


def _plan(m, shape, training, need_grad, f32):
    m.train(training)
    g = engine.Graph(training, need_grad, f32)
    outs = m._emit(g, g.input_image(*shape))
    for t in ([outs] if isinstance(outs, engine.TView) else outs):
        g.mark_output(t)
    g.finalize()
    return g


@pytest.mark.parametrize("build", [lambda: Darknet(16, [(1, 32), (2, 32)], CSPDarknetStage),
                                   lambda: VoVNet(32, [(1, 16, 2, 32), (2, 16, 3, 32)], ese=True)])
def test_fp32_plan_mirrors_bf16_plan(build):
    m = build()
    g16 = _plan(m, (2, 3, 32, 32), True, True, False)
    g32 = _plan(m, (2, 3, 32, 32), True, True, True)
    assert len(g16.ops) == len(g32.ops) and len(g16.buffers) == len(g32.buffers)
    assert all(b.esize == 2 for b in g16.buffers) and all(b.esize == 4 for b in g32.buffers)
    for a, b in zip(g16.ops, g32.ops):
        assert a.kind == b.kind
        assert (a.out.coff, a.out.c, a.out.ld) == (b.out.coff, b.out.c, b.out.ld)   # same concat slices
    assert g32.grad_bytes >= 2 * g16.grad_bytes - 1024 * len(g16.buffers)
    assert g32.dy_bytes == 2 * g16.dy_bytes
    # views never overlap inside the arena
    spans = sorted((b.offset, b.offset + b.nbytes) for b in g32.buffers)
    assert all(e0 <= s1 for (_, e0), (s1, _) in zip(spans, spans[1:]))
    # eval without gradients: bf16 uses the fused epilogue (no raw conv output), fp32 keeps conv and normalise apart
    e16 = _plan(m, (2, 3, 32, 32), False, False, False)
    e32 = _plan(m, (2, 3, 32, 32), False, False, True)
    assert e16.fused_eval and not e32.fused_eval
    assert all(op.y is None for op in e16.ops if op.kind == "conv")
    assert all(op.y is not None for op in e32.ops if op.kind == "conv")


def test_sibling_pair_plan():
    """CSP conv1 | conv2 read the same tensor: training-mode tensor-core plans run them as ONE convolution whose raw
    output, BatchNorm arrays and dy buffer hold both units side by side; every other plan keeps two ordinary units."""
    m = Darknet(16, [(1, 32), (2, 64)], CSPDarknetStage)
    m.train()
    g = engine.Graph(True, True, False, pair_ok=True)
    outs = m._emit(g, g.input_image(2, 3, 32, 32))
    for t in outs:
        g.mark_output(t)
    g.finalize()
    firsts = [op for op in g.ops if op.kind == "conv" and op.pair is not None]
    assert len(firsts) == 2                                     # one pair per CSP stage
    for a in firsts:
        b = a.pair
        assert b.pair_of is a and g.ops.index(b) == g.ops.index(a) + 1
        assert a.x is b.x or (a.x.buf is b.x.buf and a.x.coff == b.x.coff)
        assert a.y.buf is b.y.buf and a.y.coff == 0 and b.y.coff == a.geom.cout and a.y.ld == a.geom.cout + b.geom.cout
        assert a.pair_geom.cout == a.geom.cout + b.geom.cout and a.pair_geom.cin == a.geom.cin
        for name in ("mean", "invstd", "scale", "shift"):
            assert b.st[name] == a.st[name] + a.geom.cout       # one finalised array, two slices
        assert a.out.buf is not b.out.buf                       # concat slice vs. standalone tensor
        assert g.dy_bytes >= a.out.pixels * a.pair_geom.cout * 2
    # same module, plans that must NOT pair: eval, fp32 parity mode, pairing switched off
    for args in ((False, False, False, True), (True, True, True, True), (True, True, False, False)):
        gg = engine.Graph(*args)
        m.train(args[0])
        m._emit(gg, gg.input_image(2, 3, 32, 32))
        assert all(op.pair is None and op.pair_of is None for op in gg.ops if op.kind == "conv")
    m.train()


def test_gathered_operand_stem_plan():
    """The image's first convolution runs as a 1x1 GEMM over a gathered operand (k*k*3 taps -> columns, padded to 16) in
    tensor-core plans whose input needs no gradient; the padded NHWC image buffer disappears from the arena."""
    from vision_toolbox_b200.backbones import DarknetYOLOv5

    cases = [(Darknet(16, [(1, 32)], CSPDarknetStage), (9, 3), 32, (32, 32)),
             (VoVNet(32, [(1, 16, 2, 32)], ese=False), (9, 3), 32, (16, 16))]
    # the 6x6 stride-2 YOLOv5 stem (108 taps) keeps the im2col descriptors
    gy = engine.Graph(True, True, False, True, col_stem=True)
    DarknetYOLOv5(16, [(1, 32)]).train()._emit(gy, gy.input_image(2, 3, 32, 32))
    assert gy.input_col is None and all(op.col is None for op in gy.ops if op.kind == "conv")
    for m, col, kp, hw in cases:
        m.train()
        g = engine.Graph(True, True, False, True, col_stem=True)
        t_in = g.input_image(2, 3, 32, 32)
        outs = m._emit(g, t_in)
        for t in outs:
            g.mark_output(t)
        g.finalize()
        stem = next(op for op in g.ops if op.kind == "conv")
        assert stem.col == col and stem.x is g.input_col and stem.x.is_input
        assert (stem.geom.k, stem.geom.stride, stem.geom.pad, stem.geom.cin) == (1, 1, 0, kp)
        assert (stem.x.h, stem.x.w, stem.x.c) == (hw[0], hw[1], kp) and stem.cin_real == col[0] * col[1]
        assert "dw_col" in stem.st and t_in.buf not in g.buffers and g.input_unused
        assert sum(op.col is not None for op in g.ops if op.kind == "conv") == 1
        # the same model with an image that needs a gradient, or in fp32 parity mode: the ordinary stem
        for args in ((True, True, False, True, False), (True, True, True, True, True)):
            gg = engine.Graph(*args)
            m._emit(gg, gg.input_image(2, 3, 32, 32))
            assert all(op.col is None for op in gg.ops if op.kind == "conv") and gg.input_col is None


def test_channel_mismatch_raises_like_the_reference():
    """A 4-channel image into the 3-channel stem: aten::convolution raises RuntimeError for the reference
    (components.py:26); the planner must not silently read the first three channels of its padded NHWC buffer."""
    m = Darknet(16, [(1, 32)], CSPDarknetStage).train()
    for col in (False, True):
        g = engine.Graph(True, True, False, True, col_stem=col)
        with pytest.raises(RuntimeError, match="expected input to have 3 channels, but got 4"):
            m._emit(g, g.input_image(2, 4, 32, 32))
    from vision_toolbox_b200.components import ConvNormAct
    unit = ConvNormAct(32, 64, 1)
    g = engine.Graph(True, True)
    with pytest.raises(RuntimeError, match="expected input to have 32 channels, but got 16"):
        unit._emit(g, g.new_tensor(2, 8, 8, 16))


def test_half_precision_models_are_rejected_not_corrupted():
    """BatchNorm / eSE tensors reach the kernels as raw fp32 pointers - a cast model must raise at plan build."""
    from vision_toolbox_b200 import engine
    from vision_toolbox_b200.backbones import VoVNet
    from vision_toolbox_b200.components import ConvNormAct

    for cast in ("half", "bfloat16"):
        m = getattr(ConvNormAct(16, 32), cast)()
        g = engine.Graph(True, False)
        with pytest.raises(TypeError, match="float32"):
            m._emit(g, g.input_image(2, 16, 8, 8))
    v = VoVNet(32, [(1, 16, 2, 32)], ese=True)
    v.stages[0].module_0.ese.half()
    g = engine.Graph(False, False)
    with pytest.raises(TypeError, match="float32"):
        v._emit(g, g.input_image(2, 3, 32, 32))


def test_per_layer_batchnorm_mode_is_part_of_the_plan():
    """model.train() followed by bn.eval() (frozen statistics) must use running statistics for THAT layer and
    must not reuse the plan of the all-training model."""
    from vision_toolbox_b200 import engine
    from vision_toolbox_b200.backbones.darknet import CSPDarknetStage

    m = CSPDarknetStage(1, 16, 32).train()
    sig_train = engine._bn_signature(m)
    m.conv1.norm.eval()
    m.blocks[0].conv2.eval()
    sig_mixed = engine._bn_signature(m)
    assert sig_train != sig_mixed and sum(f for f, _ in sig_train) == len(sig_train)
    assert sum(f for f, _ in sig_mixed) == len(sig_train) - 2
    g = engine.Graph(True, True, pair_ok=True)
    m._emit(g, g.input_image(2, 16, 16, 16))
    flags = {id(op.mod): op.batch_stats for op in g.ops if op.kind == "conv"}
    assert flags[id(m.conv1)] is False and flags[id(m.blocks[0].conv2)] is False and flags[id(m.conv2)] is True
    # conv1 | conv2 disagree on the statistics they use: they must not be fused into one side-by-side convolution
    assert all(op.pair is None for op in g.ops if op.kind == "conv")
    bn_none = CSPDarknetStage(1, 16, 32).eval()
    bn_none.conv.norm.running_mean = None
    bn_none.conv.norm.running_var = None
    assert engine._bn_signature(bn_none)[0][0] is True       # eval without running statistics -> batch statistics
    mom = CSPDarknetStage(1, 16, 32).train()
    mom.conv.norm.momentum = None
    with pytest.raises(NotImplementedError, match="momentum"):
        g = engine.Graph(True, False)
        mom._emit(g, g.input_image(2, 16, 16, 16))
