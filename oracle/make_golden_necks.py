"""Golden vectors for FPN / PAN: runs the UNMODIFIED reference (/root/reference/vision_toolbox/necks.py:45-120) in the
build container on fixed inputs and stores state_dict, inputs, outputs and gradients (fp32, train mode) in
tests/golden_extras/necks.pt.  TEST INFRASTRUCTURE ONLY.      python oracle/make_golden_necks.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, "/root/reference")
from vision_toolbox import necks as R  # noqa: E402

cases = {}
for name, cls, kw in (("fpn_sum", "FPN", {}), ("fpn_concat", "FPN", {"fuse_fn": "concat"}), ("pan_sum", "PAN", {}),
                      ("fpn_bottom_up", "FPN", {"top_down": False})):
    torch.manual_seed(0)
    m = getattr(R, cls)([16, 32, 48], 16, **kw).train()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    xs = [torch.rand(2, 16, 16, 12), torch.rand(2, 32, 8, 6), torch.rand(2, 48, 4, 3)]
    xg = [x.clone().requires_grad_(True) for x in xs]
    outs = m(list(xg))
    cots = [torch.randn_like(o) for o in outs]
    sum((o * c).sum() for o, c in zip(outs, cots)).backward()
    cases[name] = dict(cls=cls, kw=kw, state_dict=sd, xs=xs, outs=[o.detach() for o in outs], cots=cots,
                       dxs=[x.grad for x in xg], dparams={k: p.grad.clone() for k, p in m.named_parameters()})
torch.save(cases, ROOT / "tests" / "golden_extras" / "necks.pt")
print("wrote necks.pt", list(cases))
