"""Does the result for an image depend on the batch it is computed in? (tile configuration / K-split order)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vision_toolbox_b200.backbones import Darknet
from vision_toolbox_b200.backbones.darknet import CSPDarknetStage
from vision_toolbox_b200.components import ConvNormAct

torch.manual_seed(0)
m = Darknet(16, [(1, 32), (2, 64), (1, 128)], CSPDarknetStage).cuda().eval()
g = torch.Generator().manual_seed(7)
X = torch.rand(64, 3, 64, 64, generator=g).cuda()
with torch.no_grad():
    full = m.get_feature_maps(X)
    for nb in (32, 16, 8):
        parts = [m.get_feature_maps(X[i:i + nb]) for i in range(0, 64, nb)]
        for lvl in range(len(full)):
            cat = torch.cat([p[lvl] for p in parts])
            d = (cat.float() - full[lvl].float()).abs()
            print(f"chunk {nb:2d} level {lvl}: max abs diff {float(d.max()):.3e}  differing {float((d > 0).float().mean()):.2e}  (max |ref| {float(full[lvl].float().abs().max()):.2f})")
# single units, train mode, raw statistics: same input split in 1 vs 4 chunks is not comparable (batch stats) -> compare
# the per-image conv output through eval mode with identity BN
for (cin, cout, k, s, hw) in [(16, 32, 3, 1, 64), (32, 64, 3, 2, 64), (64, 64, 1, 1, 32), (64, 128, 3, 2, 16), (128, 128, 3, 1, 8)]:
    u = ConvNormAct(cin, cout, k, s).cuda().eval()
    x = torch.rand(64, cin, hw, hw, device="cuda")
    with torch.no_grad():
        a = u(x)
        b = torch.cat([u(x[i:i + 8]) for i in range(0, 64, 8)])
    d = (a.float() - b.float()).abs()
    print(f"unit {cin}->{cout} k{k}s{s} {hw}px: max abs diff {float(d.max()):.3e} differing {float((d > 0).float().mean()):.2e}")
