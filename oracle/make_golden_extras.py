"""Golden vectors for the batch transforms (RandomMixup / RandomCutmix / RandomCutMixMixUp): runs the UNMODIFIED reference
(/root/reference/extras.py:14-109) in the build container with fixed seeds and stores inputs, seeds and outputs in
tests/golden_extras/mix.pt.  TEST INFRASTRUCTURE ONLY (the reference does not exist on the GPU box).

    python oracle/make_golden_extras.py
"""
import importlib.util
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("ref_extras", "/root/reference/extras.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

torch.manual_seed(0)
x = torch.rand(4, 3, 10, 12)
y = torch.randint(0, 10, (4,))
cases = []
for name, args in (("RandomMixup", (10, 0.7, 0.4)), ("RandomCutmix", (10, 0.7, 1.0)), ("RandomCutMixMixUp", (10, 1.0, 0.2))):
    for seed in range(8):
        torch.manual_seed(seed)
        b, t = getattr(ref, name)(*args)(x, y)
        cases.append(dict(cls=name, args=args, seed=seed, batch=b.clone(), target=t.clone()))
out = ROOT / "tests" / "golden_extras"
out.mkdir(exist_ok=True)
torch.save(dict(x=x, y=y, cases=cases), out / "mix.pt")
print("wrote", out / "mix.pt", len(cases), "cases")
