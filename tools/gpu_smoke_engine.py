"""Quick engine smoke on the GPU box: one golden case at a time with verbose errors (dev aid)."""
import sys, traceback
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from conftest import GOLDEN_CASES, load_golden, rel_err
from helpers import BUILDERS, module_outputs

for name in GOLDEN_CASES:
    try:
        g = load_golden(name)
        m = BUILDERS[name](); m.load_state_dict(g["state_dict"]); m = m.cuda().train()
        x = g["x"].cuda().requires_grad_(True)
        outs = module_outputs(m, x)
        loss = sum((o.float() * c.cuda()).sum() for o, c in zip(outs, g["cotangents"]))
        loss.backward(); torch.cuda.synchronize()
        fe = [rel_err(o.float(), r) for o, r in zip(outs, g["train_bf16_outs"])]
        ge = {k: rel_err(p.grad, g["train_bf16_dparams"][k]) for k, p in m.named_parameters()}
        gr = {k: rel_err(g["train_bf16_dparams"][k], g["train_fp32_dparams"][k]) for k in ge}
        wk = max(ge, key=ge.get)
        print(f"{name}: fwd {max(fe):.2e} dx {rel_err(x.grad, g['train_bf16_dx']):.2e} worst dparam {wk} {ge[wk]:.2e} (ref bf16-vs-fp32 {gr[wk]:.2e})")
    except Exception:
        print(name, "EXCEPTION"); traceback.print_exc()
        break
