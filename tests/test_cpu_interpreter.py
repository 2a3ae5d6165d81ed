"""The executor on the CPU against the reference's golden vectors, through tests/fake_vtb.InterpreterLib (a torch-CPU
implementation of the C ABI's documented semantics).  Verifies the HOST side of the product - plans, buffer slicing,
gradient routing, sibling pairing, the gathered-operand stem, the pack job table - with real numbers and no GPU; the CUDA
kernels meet the same goldens in tests/test_gpu_parity.py."""
from unittest import mock

import pytest
import torch

from conftest import GOLDEN_CASES, load_golden, rel_err
from fake_vtb import InterpreterLib
from helpers import BUILDERS
from vision_toolbox_b200 import engine
from vision_toolbox_b200.backbones.base import BaseBackbone


class TwinRankDist:
    """A second rank that holds the SAME batch: all-reducing a statistic doubles it.  Global mean / biased variance / the
    dx formula are then those of the single batch, so the multi-rank code path must reproduce the single-process numbers
    (only the unbiased running-variance factor differs: 2M / (2M - 1))."""

    def __init__(self):
        self.world, self.rank, self.sync_bn, self.sync, self.on_grads_ready = 2, 0, True, None, None
        self.calls = 0

    def all_reduce_(self, t):
        t.mul_(2)
        self.calls += 1


def _run(name, g, training, need_grad, x_grad, pair=True, col=True, f32=False, dist=None):
    m = BUILDERS[name]()
    m.load_state_dict(g["state_dict"])
    m.train(training)
    graph = engine.Graph(training, need_grad, f32, pair_ok=pair, col_stem=col and not x_grad)
    t_in = graph.input_image(*g["x"].shape)
    outs = m._emit(graph, t_in) if not isinstance(m, BaseBackbone) else m._emit(graph, t_in)
    for t in ([outs] if isinstance(outs, engine.TView) else outs):
        graph.mark_output(t)
    graph.finalize()
    lib = InterpreterLib()
    with mock.patch.object(engine._lib, "lib", return_value=lib):
        runner = engine.Runner(graph, torch.device("cpu"))
    runner._stream = lambda: 0
    runner.dist = dist
    x = g["x"].clone().requires_grad_(x_grad)
    outs_t, run = runner.forward(x)
    outs_f = [o.float().clone() for o in outs_t]
    if not need_grad:
        return m, graph, outs_f, None, None
    gouts = [c.to(torch.float32 if f32 else torch.bfloat16).contiguous(memory_format=torch.channels_last)
             for c in g["cotangents"]]
    gx, pgrads = runner.backward(run, gouts)
    grads = {}
    names = {id(p): k for k, p in m.named_parameters()}
    for p, gr in zip(graph.params, pgrads):
        grads[names[id(p)]] = gr.clone()
    return m, graph, outs_f, grads, gx


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_train_forward_and_statistics(name):
    g = load_golden(name)
    with torch.no_grad():
        m, graph, outs, _, _ = _run(name, g, True, False, False)
    for o, ref in zip(outs, g["train_bf16_outs"]):
        assert tuple(o.shape) == tuple(ref.shape) and rel_err(o, ref) < 2e-2, (name, rel_err(o, ref))
    sd = m.state_dict()
    for k, ref in g["buffers_after_bf16_step"].items():
        assert rel_err(sd[k].float(), ref) < 2e-3, k
    for k, ref in g["buffers_after_step"].items():
        if "num_batches" in k:
            assert int(sd[k]) == int(ref), k


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_eval_forward(name):
    g = load_golden(name)
    with torch.no_grad():
        m, graph, outs, _, _ = _run(name, g, False, False, False)
    assert graph.fused_eval
    for o, ref in zip(outs, g["eval_fp32_outs"]):
        assert rel_err(o, ref) < 2e-2
    for k, v in g["state_dict"].items():
        assert torch.equal(m.state_dict()[k], v), k


@pytest.mark.parametrize("dgrad_bn", ["0", "1"], ids=["bn-bwd-kernel", "bn-bwd-sums-in-dgrad"])
@pytest.mark.parametrize("x_grad", [False, True], ids=["stem-gathered", "image-gradient"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_train_backward(name, x_grad, dgrad_bn, monkeypatch):
    monkeypatch.setenv("VTB_DGRAD_BN", dgrad_bn)   # opt-in plan: vtb_conv_dgrad_bn + vtb_bn_bwd_apply
    g = load_golden(name)
    m, graph, outs, grads, gx = _run(name, g, True, True, x_grad)
    fused = sum(op.dgrad_bn is not None for op in graph.ops if op.kind == "conv")
    assert (fused > 0) == (dgrad_bn == "1" and sum(op.kind == "conv" for op in graph.ops) > 1), fused
    ref16, ref32 = dict(g["train_bf16_dparams"]), dict(g["train_fp32_dparams"])
    if x_grad:
        grads["__dx__"], ref16["__dx__"], ref32["__dx__"] = gx, g["train_bf16_dx"], g["train_fp32_dx"]
    else:
        assert gx is None
    assert set(grads) == set(ref32)
    for k in ref32:
        assert grads[k].shape == ref32[k].shape and torch.isfinite(grads[k]).all(), k
        e_ours, e_ref = rel_err(grads[k], ref32[k]), rel_err(ref16[k], ref32[k])
        assert e_ours < 2.0 * e_ref + 1e-2, (name, k, e_ours, e_ref)     # SURVEY.md Appendix B criterion


@pytest.mark.parametrize("name", ["block_darknet_32", "model_cspdarknet", "model_vovnet_ese", "unit_3x3s2_32_32_odd"])
def test_fused_normalise_plan_is_the_same_function(name, monkeypatch):
    """VTB_FUSED_NORM=1: the unit's normalise + ReLU (+ residual) pass rides in vtb_conv_fprop_bn (VtbBnTrain.act_*) instead
    of its own vtb_bn_act call - forward maps, statistics and gradients are those of the default plan, bit for bit."""
    g = load_golden(name)
    res = []
    for fused in ("0", "1"):
        monkeypatch.setenv("VTB_FUSED_NORM", fused)
        m, graph, outs, grads, gx = _run(name, g, True, True, True)
        res.append((outs, grads, gx, {k: v.clone() for k, v in m.state_dict().items() if "running" in k}))
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    for k in res[0][1]:
        assert torch.equal(res[0][1][k], res[1][1][k]), k
    assert torch.equal(res[0][2], res[1][2])
    for k in res[0][3]:
        assert torch.equal(res[0][3][k], res[1][3][k]), k


@pytest.mark.parametrize("nb", [1, 4])
def test_native_sgd_job_tables_against_torch_sgd(nb):
    """parallel.NativeSGD (vtb_sgd_pack_weights / vtb_sgd_step job tables, one set per gradient bucket) through the
    interpreter: after two steps the parameters equal torch.optim.SGD's (momentum 0.9, weight decay on conv weights only),
    whether the optimizer steps everything at once (nb = 1) or bucket by bucket in any order (nb = 4)."""
    import copy

    from vision_toolbox_b200 import parallel

    name = "model_cspdarknet"
    g = load_golden(name)
    m = BUILDERS[name]()
    m.load_state_dict(g["state_dict"])
    m.train()
    ref = copy.deepcopy(m)
    graph = engine.Graph(True, True, False, pair_ok=True, col_stem=True)
    t_in = graph.input_image(*g["x"].shape)
    for t in m._emit(graph, t_in):
        graph.mark_output(t)
    graph.finalize()
    lib = InterpreterLib()
    with mock.patch.object(engine._lib, "lib", return_value=lib):
        runner = engine.Runner(graph, torch.device("cpu"))
        runner._stream = lambda: 0
        m.__dict__["_vtb_plans"] = {"train": runner}
        params = list(m.parameters())
        decay, no_decay = parallel.split_decay_groups([m])
        opt = parallel.NativeSGD(m, params, decay, lr=0.1, momentum=0.9, weight_decay=1e-2)
        if nb > 1:
            opt.set_buckets({id(p): i % nb for i, p in enumerate(params)}, nb)
        rdecay, rno = parallel.split_decay_groups([ref])
        ropt = torch.optim.SGD([{"params": rdecay, "weight_decay": 1e-2}, {"params": rno, "weight_decay": 0.0}],
                               lr=0.1, momentum=0.9)
        gen = torch.Generator().manual_seed(3)
        for step in range(2):
            for p, q in zip(params, ref.parameters()):
                gr = torch.randn(p.shape, generator=gen)
                if p.grad is None:
                    p.grad = gr            # the job tables hold the gradients' addresses: later steps write in place
                else:
                    p.grad.copy_(gr)
                q.grad = gr.clone()
            if nb > 1 and step == 1:
                assert opt.ready()
                for b in (2, 0, 3):             # some buckets early, in any order; step() finishes the rest
                    opt.step_bucket(b)
            opt.step()
            ropt.step()
        for (k, p), q in zip(m.named_parameters(), ref.parameters()):
            assert rel_err(p, q) < 2e-6, k
        # the bf16 operands the fused step left behind are those of the new weights: same forward as a fresh re-pack
        with torch.no_grad():
            a = [o.float().clone() for o in runner.forward(g["x"])[0]]
            runner._packs_token = None
            b = [o.float().clone() for o in runner.forward(g["x"])[0]]
        for u, v in zip(a, b):
            assert torch.equal(u, v)


def test_pairing_and_gathered_stem_do_not_change_the_result():
    """The plan-level transformations are exact rewrites: same numbers with and without them (up to the summation order of
    the torch kernels the interpreter uses)."""
    name = "model_cspdarknet"
    g = load_golden(name)
    _, g1, o1, gr1, _ = _run(name, g, True, True, False, pair=True, col=True)
    _, g0, o0, gr0, _ = _run(name, g, True, True, False, pair=False, col=False)
    assert any(op.pair is not None for op in g1.ops if op.kind == "conv") and g1.input_col is not None
    assert all(op.pair is None for op in g0.ops if op.kind == "conv") and g0.input_col is None
    for a, b in zip(o1, o0):
        assert rel_err(a, b) < 1e-2
    for k in gr0:
        assert rel_err(gr1[k], gr0[k]) < 2e-2, k


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fp32_mode_plans(name):
    """The fp32 parity-mode plan (separate statistics / finalise / normalise launches, fp32 arenas) through the interpreter:
    1e-5 on feature maps and running statistics, 1e-4 on gradients, against the reference's fp32 run."""
    g = load_golden(name)
    m, graph, outs, grads, gx = _run(name, g, True, True, True, f32=True)
    assert not any(op.pair is not None or op.col is not None for op in graph.ops if op.kind == "conv")
    for o, ref in zip(outs, g["train_fp32_outs"]):
        assert rel_err(o, ref) < 1e-5, (name, rel_err(o, ref))
    sd = m.state_dict()
    for k, ref in g["buffers_after_step"].items():
        if "num_batches" in k:
            assert int(sd[k]) == int(ref)
        else:
            assert rel_err(sd[k], ref) < 1e-5, k
    ref = dict(g["train_fp32_dparams"])
    grads["__dx__"], ref["__dx__"] = gx, g["train_fp32_dx"]
    for k in ref:
        assert rel_err(grads[k], ref[k]) < 1e-4, (name, k, rel_err(grads[k], ref[k]))
    with torch.no_grad():
        m2, graph2, outs2, _, _ = _run(name, g, False, False, False, f32=True)
    assert not graph2.fused_eval
    for o, r in zip(outs2, g["eval_fp32_outs"]):
        assert rel_err(o, r) < 1e-5


@pytest.mark.parametrize("f32", [False, True], ids=["bf16", "fp32"])
@pytest.mark.parametrize("name", ["stage_csp_2_16_32", "model_cspdarknet", "model_vovnet_ese"])
def test_all_reduce_syncbn_path_with_a_twin_rank(name, f32):
    """SyncBN through one all-reduce per exchange (the NCCL fallback of bf16 plans, the only multi-rank path of fp32
    plans), fed by a twin rank: same feature maps and gradients as the single process."""
    g = load_golden(name)
    dist = TwinRankDist()
    m, graph, outs, grads, _ = _run(name, g, True, True, False, pair=False, f32=f32, dist=dist)
    units = sum(op.kind == "conv" for op in graph.ops)
    assert dist.calls == 2 * units
    for o, ref in zip(outs, g["train_fp32_outs" if f32 else "train_bf16_outs"]):
        assert rel_err(o, ref) < (1e-5 if f32 else 2e-2)
    ref32, ref16 = g["train_fp32_dparams"], g["train_bf16_dparams"]
    for k in ref32:
        if f32:
            assert rel_err(grads[k], ref32[k]) < 1e-4, k
        else:
            assert rel_err(grads[k], ref32[k]) < 2.0 * rel_err(ref16[k], ref32[k]) + 1e-2, k
    sd = m.state_dict()
    for k, ref in g["buffers_after_step" if f32 else "buffers_after_bf16_step"].items():
        if "running_mean" in k:
            assert rel_err(sd[k], ref) < (1e-5 if f32 else 2e-3), k


def test_bias_conv_unit_plan():
    """The planner's bias unit (a bare nn.Conv2d with bias: the lateral convolutions of the reference necks, necks.py:60-65)
    on the CPU interpreter: forward, input / weight / bias gradients against torch."""
    from vision_toolbox_b200.necks import _BiasConvUnit

    torch.manual_seed(0)
    for k, s in ((1, 1), (3, 2)):
        conv = torch.nn.Conv2d(32, 48, k, s, padding=k // 2)
        unit = _BiasConvUnit(conv)
        x = torch.rand(2, 32, 9, 10)
        cot = torch.randn(2, 48, (9 + 2 * (k // 2) - k) // s + 1, (10 + 2 * (k // 2) - k) // s + 1)
        graph = engine.Graph(True, True)
        out = unit._emit(graph, graph.input_image(*x.shape))
        graph.mark_output(out)
        graph.finalize()
        lib = InterpreterLib()
        with mock.patch.object(engine._lib, "lib", return_value=lib):
            runner = engine.Runner(graph, torch.device("cpu"))
        runner._stream = lambda: 0
        xx = x.clone().requires_grad_(True)
        outs, run = runner.forward(xx)
        gx, pg = runner.backward(run, [cot.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)])
        xr = x.clone().requires_grad_(True)
        ref = conv(xr)
        (ref * cot).sum().backward()
        assert rel_err(outs[0].float(), ref) < 1e-2
        assert rel_err(gx, xr.grad) < 2e-2
        assert rel_err(pg[0], conv.weight.grad) < 2e-2 and rel_err(pg[1], conv.bias.grad) < 2e-2
        assert [n for n in lib.calls if "bn_act" in n or "fprop_bn" in n] == []      # one conv launch, no normalise pass
