"""Two training steps of a bench model (first = warm-up incl. weight packing) — the command ncu wraps.
    python tools/one_step.py [model] [batch] [res] [steps]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from vision_toolbox_b200 import backbones, parallel

name = sys.argv[1] if len(sys.argv) > 1 else "cspdarknet53"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
res = int(sys.argv[3]) if len(sys.argv) > 3 else 176
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = getattr(backbones, name)().to(dev).train()
head = torch.nn.Linear(model.out_channels_list[-1], 1000).to(dev)
tr = parallel.Trainer(model, head)
x = torch.rand(nb, 3, res, res, device=dev)
y = torch.randint(0, 1000, (nb,), device=dev)
for i in range(steps):
    loss = tr.step(x, y)
    torch.cuda.synchronize()
    print(f"step {i} loss {float(loss):.4f}", flush=True)
