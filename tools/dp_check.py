"""Data-parallel parity on real GPUs (run under torchrun, >= 2 ranks):
SyncBN + gradient mean over R ranks must equal ONE process running the concatenated global batch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

from vision_toolbox_b200 import parallel
from vision_toolbox_b200.backbones import Darknet
from vision_toolbox_b200.backbones.darknet import CSPDarknetStage


def build():
    torch.manual_seed(0)
    m = Darknet(16, [(1, 32), (2, 64), (1, 128)], CSPDarknetStage)
    head = torch.nn.Linear(128, 10)
    return m, head


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nb = 8
    g = torch.Generator().manual_seed(7)
    X = torch.rand(nb * world, 3, 64, 64, generator=g)
    Y = torch.randint(0, 10, (nb * world,), generator=g)

    # reference: one process, the whole global batch, no process group
    m1, h1 = build()
    t1 = parallel.Trainer(m1.to(dev).train(), h1.to(dev), lr=0.0, momentum=0.0, weight_decay=0.0)
    t1.flat.zero_()
    l1 = t1.forward_loss(X.to(dev), Y.to(dev))
    l1.backward()
    ref = t1.flat.clone()
    ref_stats = {k: v.clone() for k, v in m1.state_dict().items() if "running" in k}

    for mode in (os.environ.get("VTB_SYNCBN", "p2p"),):
        m2, h2 = build()
        t2 = parallel.Trainer(m2.to(dev).train(), h2.to(dev), lr=0.0, momentum=0.0, weight_decay=0.0, sync_bn=True,
                              process_group=dist.group.WORLD, bucket_mb=0.05)
        path = "peer-memory" if t2.dist_cfg.sync is not None else f"nccl ({t2.dist_cfg.sync_error})"
        xs, ys = X[rank * nb:(rank + 1) * nb].to(dev), Y[rank * nb:(rank + 1) * nb].to(dev)
        for it in range(2):   # twice: exercises the parity double-buffering of the exchange
            t2.flat.zero_()
            if it == 1:
                m2.load_state_dict(build()[0].state_dict())
            l2 = t2.forward_loss(xs, ys)
            l2.backward()
            t2._finish_exchange()
        torch.cuda.synchronize()
        err = float((t2.flat - ref).norm() / ref.norm())
        lerr = torch.tensor([float(l2)], device=dev)
        dist.all_reduce(lerr)
        loss_err = abs(float(lerr) / world - float(l1))
        serr = max(float((m2.state_dict()[k] - v).abs().max() / v.abs().max().clamp_min(1e-6)) for k, v in ref_stats.items())
        ok = err < 2e-2 and loss_err < 1e-3 and serr < 1e-3
        print(f"rank {rank}: SyncBN via {path}: grad rel err vs single-process global batch {err:.3e}, "
              f"loss diff {loss_err:.2e}, running-stat max rel err {serr:.2e}, buckets {len(t2.buckets)} -> "
              f"{'OK' if ok else 'FAIL'}", flush=True)
        if not ok:
            dist.destroy_process_group()
            sys.exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
