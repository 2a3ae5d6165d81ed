"""A few training-mode forwards of ONE ConvNormAct unit (the command ncu wraps for a single-launch capture).
    python tools/one_unit.py cin cout k stride hw [batch] [iters]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from vision_toolbox_b200.components import ConvNormAct

cin, cout, k, s, hw = (int(v) for v in sys.argv[1:6])
nb = int(sys.argv[6]) if len(sys.argv) > 6 else 256
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
torch.manual_seed(0)
m = ConvNormAct(cin, cout, k, s).cuda().train()
x = torch.rand(nb, cin, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    for i in range(iters):
        y = m(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.float().mean()))
