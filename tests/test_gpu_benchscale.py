"""Parity in the tile regime the benchmark runs (VERDICT r01 "what's weak" #1): single ConvNormAct units at the real
CSPDarknet-53 shapes of SURVEY.md table T-C3 with the benchmark's batch (N = 256), so that persistent CTAs walk several
tiles, the two TMEM accumulator sets alternate, mbarrier phases wrap, the per-CTA BatchNorm statistics accumulate across
tiles, K-split accumulator chains and split-K weight gradients are all exercised - and compared with the ORACLE
(oracle/vt_oracle.py, bf16 rounding points; reference components.py:26-46) on the same tensors.

The oracle runs on the GPU in fp32 (TF32 off) for speed; one case is also run on the CPU to pin the GPU yardstick to the
CPU oracle.  Tolerances: north_star bf16 mode 2e-2 relative (teacher-forced single units), running statistics against the
bf16-mode oracle (same rounding points)."""
import ctypes as C

import pytest
import torch

from conftest import rel_err
from oracle import vt_oracle as O

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2
N = 256

# name: (cin, cout, k, stride, H)  -- SURVEY.md T-C3 rows (CSPDarknet-53 @176, batch 256)
UNITS = {
    "c3x3_128_128_22": (128, 128, 3, 1, 22),     # the 19.2 % row: tensor-bound, multi-tile CTAs, two TMEM sets
    "c3x3_256_256_11": (256, 256, 3, 1, 11),     # 19.2 % row: N = 256 tile, single TMEM set
    "c3x3s2_512_1024_11": (512, 1024, 3, 2, 11), # odd input, stride 2: 4-phase dgrad, 4 n-blocks
    "c1x1_64_32_88": (64, 32, 1, 1, 88),         # HBM-bound 1x1, 7744 tiles
    "c3x3_64_64_44": (64, 64, 3, 1, 44),         # narrow tile: K-split accumulator chains (ksplit 2)
    "c1x1_512_512_6": (512, 512, 1, 1, 6),       # 6x6 stage: one tile per CTA, split-K wgrad
    "c3x3s2_32_64_176": (32, 64, 3, 2, 176),     # stage-1 stride-2 conv, M = 1.98 M
}
_seen = {}


def _tiling(cin, cout, k, s, h, op):
    from vision_toolbox_b200 import _lib

    L = _lib.lib()
    g = _lib.VtbConv(N, h, h, cin, cout, k, s, (k - s + 1) // 2)
    info = (C.c_int * 8)()
    _lib.check(L.vtb_conv_tiling_info(C.byref(g), op, info), "vtb_conv_tiling_info")
    return list(info)


def _make_unit(cin, cout, k, s, seed=0):
    from vision_toolbox_b200.components import ConvNormAct

    torch.manual_seed(seed)
    m = ConvNormAct(cin, cout, k, s)
    with torch.no_grad():
        m.norm.weight.uniform_(0.5, 1.5)
        m.norm.bias.uniform_(-0.2, 0.2)
        m.norm.running_mean.uniform_(-0.1, 0.1)
        m.norm.running_var.uniform_(0.5, 1.5)
    return m


def _oracle_unit(sd, x, cot, stride, device):
    """bf16-mode oracle of one unit: train forward + backward of <out, cot> (+ the running-statistics update)."""
    sd = {k: v.to(device) for k, v in sd.items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    full["__stride__"] = {"": stride}
    xg = x.to(device).clone().requires_grad_(True)
    new_stats = {}
    out = O.conv_norm_act(full, "", xg, True, "bf16", new_stats=new_stats)
    grads = torch.autograd.grad((out * cot.to(device)).sum(), [xg] + list(params.values()))
    return out.detach(), grads[0], dict(zip(params.keys(), grads[1:])), new_stats


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("name", list(UNITS))
def test_unit_at_benchmark_batch(name):
    cin, cout, k, s, h = UNITS[name]
    fp, dg, wg = _tiling(cin, cout, k, s, h, 0), _tiling(cin, cout, k, s, h, 1), _tiling(cin, cout, k, s, h, 2)
    _seen[name] = (fp, dg, wg)
    print(f"{name}: fprop tiling {fp}  dgrad {dg}  wgrad {wg}")
    m = _make_unit(cin, cout, k, s)
    sd = {kk: v.clone() for kk, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.rand(N, cin, h, h, generator=g)
    ho = (h + 2 * ((k - s + 1) // 2) - k) // s + 1
    cot = torch.randn(N, cout, ho, ho, generator=g)
    m = m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    out = m(xg)
    (out.float() * cot.cuda()).sum().backward()
    torch.cuda.synchronize()
    ref_out, ref_dx, ref_dp, ref_stats = _oracle_unit(sd, x, cot, s, "cuda")
    e = rel_err(out.float(), ref_out)
    print(f"  forward rel err {e:.2e}")
    assert e < BF16_TOL / 2
    msd = m.state_dict()
    for key in ("norm.running_mean", "norm.running_var"):
        es = rel_err(msd[key], ref_stats[key])
        print(f"  {key} rel err {es:.2e}")
        assert es < 5e-4, key
    assert int(msd["norm.num_batches_tracked"]) == int(sd["norm.num_batches_tracked"]) + 1
    errs = {"dx": rel_err(xg.grad, ref_dx)}
    for kk, p in m.named_parameters():
        errs[kk] = rel_err(p.grad, ref_dp[kk])
    print("  gradient rel errs " + ", ".join(f"{a} {b:.2e}" for a, b in errs.items()))
    for kk, v in errs.items():
        assert v < BF16_TOL, (kk, v)


def test_stem_at_benchmark_batch():
    """3 -> 32 stem at 176^2, batch 256 (7.9 M pixels: where the sum / sum-of-squares cancellation of the BatchNorm
    statistics bites) through the gathered-operand path (image without gradient) against the bf16-mode oracle."""
    m = _make_unit(3, 32, 3, 1)
    sd = {kk: v.clone() for kk, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    x = torch.rand(N, 3, 176, 176, generator=g)
    cot = torch.randn(N, 32, 176, 176, generator=g)
    m = m.cuda().train()
    out = m(x.cuda())
    (out.float() * cot.cuda()).sum().backward()
    torch.cuda.synchronize()
    ref_out, _, ref_dp, ref_stats = _oracle_unit(sd, x, cot, 1, "cuda")
    assert rel_err(out.float(), ref_out) < BF16_TOL / 2
    msd = m.state_dict()
    for key in ("norm.running_mean", "norm.running_var"):
        es = rel_err(msd[key], ref_stats[key])
        print(f"  stem {key} rel err {es:.2e}")
        assert es < 5e-4, key
    for kk, p in m.named_parameters():
        e = rel_err(p.grad, ref_dp[kk])
        print(f"  stem {kk} rel err {e:.2e}")
        assert e < BF16_TOL, (kk, e)


def test_gpu_yardstick_matches_cpu_oracle():
    """The GPU (cuDNN fp32, TF32 off) run of the oracle used above agrees with its CPU run on a benchmark-shaped unit."""
    cin, cout, k, s, h = UNITS["c3x3_128_128_22"]
    m = _make_unit(cin, cout, k, s)
    sd = {kk: v.clone() for kk, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    n = 32
    x = torch.rand(n, cin, h, h, generator=g)
    cot = torch.randn(n, cout, h, h, generator=g)
    a = _oracle_unit(sd, x, cot, s, "cpu")
    b = _oracle_unit(sd, x, cot, s, "cuda")
    assert rel_err(b[0], a[0]) < 1e-3
    # gradients differ through ReLU-mask flips of bf16-rounded pre-activations: 5e-3 floor
    assert rel_err(b[1], a[1]) < 1e-2
    for kk in a[2]:
        assert rel_err(b[2][kk], a[2][kk]) < 1e-2, kk


def _oracle_module(fn, sd, strides, x, cot, mode, device):
    sd = {k: v.to(device) for k, v in sd.items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    full["__stride__"] = strides
    xg = x.to(device).clone().requires_grad_(True)
    out = fn(full, "", xg, True, mode, None)
    grads = torch.autograd.grad((out * cot.to(device)).sum(), [xg] + list(params.values()))
    return out.detach(), grads[0], dict(zip(params.keys(), grads[1:]))


@pytest.mark.parametrize("kind", ["block", "csp_stage"])
def test_block_and_stage_at_benchmark_batch(kind):
    """DarknetBlock(128, expansion 1) at 22^2 and CSPDarknetStage(2, 64, 128) at 44 -> 22, batch 256 (reference
    darknet.py:20-28, 39-55): forward against the bf16-mode oracle, gradients per SURVEY.md Appendix B (error against the
    fp32 oracle no worse than twice the bf16 oracle's own error)."""
    from vision_toolbox_b200.backbones.darknet import CSPDarknetStage, DarknetBlock

    torch.manual_seed(3)
    if kind == "block":
        m, fn, strides, cin, h, cout, ho = DarknetBlock(128, 1), O.darknet_block, {}, 128, 22, 128, 22
    else:
        m, fn, strides, cin, h, cout, ho = CSPDarknetStage(2, 64, 128), O.csp_stage, {"conv.": 2}, 64, 44, 128, 22
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.2, 0.2)
    sd = {kk: v.clone() for kk, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    x = torch.rand(N, cin, h, h, generator=g)
    cot = torch.randn(N, cout, ho, ho, generator=g)
    m = m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    out = m(xg)
    (out.float() * cot.cuda()).sum().backward()
    torch.cuda.synchronize()
    o16, dx16, dp16 = _oracle_module(fn, sd, strides, x, cot, "bf16", "cuda")
    o32, dx32, dp32 = _oracle_module(fn, sd, strides, x, cot, "fp32", "cuda")
    e = rel_err(out.float(), o16)
    print(f"{kind}: forward vs bf16 oracle {e:.2e}; bf16 oracle vs fp32 oracle {rel_err(o16, o32):.2e}")
    assert e < BF16_TOL
    ours = {kk: p.grad for kk, p in m.named_parameters()}
    ours["__dx__"], dp16["__dx__"], dp32["__dx__"] = xg.grad, dx16, dx32
    worst = 0.0
    for kk in dp32:
        e_ours, e_ref = rel_err(ours[kk], dp32[kk]), rel_err(dp16[kk], dp32[kk])
        worst = max(worst, e_ours / max(e_ref, 1e-6))
        assert e_ours < 2.0 * e_ref + 1e-2, (kk, e_ours, e_ref)
    print(f"{kind}: worst (our err)/(bf16 oracle err) vs fp32 oracle = {worst:.2f}")


def test_benchmark_regime_is_what_ran():
    """The cases above really were multi-tile / double-buffered / K-split / split-K plans (vtb_conv_tiling_info)."""
    info = {n: (_tiling(*UNITS[n], 0), _tiling(*UNITS[n], 1), _tiling(*UNITS[n], 2)) for n in UNITS}
    fp = {n: v[0] for n, v in info.items()}
    # {block_m, block_n, grid, tiles, max tiles per CTA, TMEM sets, ksplit, stages}
    assert fp["c3x3_128_128_22"][4] > 1 and fp["c3x3_128_128_22"][5] == 2          # several tiles, alternating TMEM sets
    assert fp["c1x1_64_32_88"][4] >= 8                                             # mbarrier phases wrap many times
    assert any(v[0][6] >= 2 or v[1][6] >= 2 for v in info.values())                # K-split accumulator chains
    assert any(v[0][5] == 1 for v in info.values())                                # single-set (wide tile) plans too
    assert fp["c3x3s2_512_1024_11"][2] // max(1, 1024 // fp["c3x3s2_512_1024_11"][1]) >= 1
    assert fp["c3x3s2_512_1024_11"][1] < 1024                                      # several n-blocks
    wg = {n: v[2] for n, v in info.items()}
    assert any(v[3] > 1 for v in wg.values()) and any(v[3] == 1 and v[4] >= 64 for v in wg.values())   # split-K and one long-K split
