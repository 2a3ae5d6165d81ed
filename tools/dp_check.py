"""Data-parallel parity on real GPUs (run under torchrun, >= 2 ranks).

  * SyncBN + gradient mean over R ranks vs ONE process running the concatenated global batch:
    a single ConvNormAct + head is held to the bf16 budget (2e-2); the deep model is reported and loosely bounded, because
    the tile configuration (hence the fp32 summation order inside the tensor-core K loop) depends on the per-rank batch,
    bf16 roundings flip for ~1e-5 of the activations per layer and train-mode gradients amplify that.
  * the peer-memory SyncBN exchange vs the NCCL all-reduce exchange in the SAME world: must agree to fp32 round-off.
  * BatchNorm running statistics identical on every rank and equal to the single-process ones.
  * fp32 parity mode (vision_toolbox_b200.precision("fp32"), SyncBN through one all-reduce of the fp64 sums): the same
    comparison at fp32 accuracy - a single unit to 1e-4 (measured ~1e-6), the deep model to 2e-3, running statistics 1e-5.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

import vision_toolbox_b200 as vtb
from vision_toolbox_b200 import parallel
from vision_toolbox_b200.backbones import Darknet
from vision_toolbox_b200.backbones.darknet import CSPDarknetStage


def build(kind):
    torch.manual_seed(0)
    if kind == "unit":
        m = Darknet(32, [(0, 64)])            # stem + one stride-2 ConvNormAct
        return m, torch.nn.Linear(64, 10)
    m = Darknet(16, [(1, 32), (2, 64), (1, 128)], CSPDarknetStage)
    return m, torch.nn.Linear(128, 10)


def run(kind, dev, X, Y, group, mode):
    if mode is not None:
        os.environ["VTB_SYNCBN"] = mode
    m, h = build(kind)
    t = parallel.Trainer(m.to(dev).train(), h.to(dev), lr=0.0, momentum=0.0, weight_decay=0.0, sync_bn=True,
                         process_group=group, bucket_mb=0.05)
    for it in range(2):   # twice: exercises the parity double-buffering of the exchange
        t.flat.zero_()
        if it == 1:
            m.load_state_dict(build(kind)[0].state_dict())
        loss = t.forward_loss(X, Y)
        loss.backward()
        if group is not None:
            t._finish_exchange()
    torch.cuda.synchronize()
    path = "single" if group is None else ("peer-memory" if t.dist_cfg.sync is not None else "nccl")
    stats = {k: v.clone() for k, v in m.state_dict().items() if "running" in k}
    return t.flat.clone(), float(loss.detach()), stats, path


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from vision_toolbox_b200.parallel import nccl_pg_options

    dist.init_process_group("nccl", device_id=dev, pg_options=nccl_pg_options())
    nb = 8
    g = torch.Generator().manual_seed(7)
    X = torch.rand(nb * world, 3, 64, 64, generator=g).to(dev)
    Y = torch.randint(0, 10, (nb * world,), generator=g).to(dev)
    xs, ys = X[rank * nb:(rank + 1) * nb], Y[rank * nb:(rank + 1) * nb]
    ok_all = True
    for kind, tol in (("unit", 2e-2), ("deep", 0.5)):
        ref, l1, s1, _ = run(kind, dev, X, Y, None, None)
        gp, lp, sp, path_p = run(kind, dev, xs, ys, dist.group.WORLD, "p2p")
        # exchange mechanism alone: same kernels on both sides (the NCCL path finalises BatchNorm in separate launches, so
        # CSP sibling units are not paired there; pairing changes the tile configuration, hence bf16 roundings)
        # The flag protocol of the peer-memory exchange is bit-identical to the NCCL all-reduce; the default tagged
        # protocol perturbs the fp64 sums by 2^-36 (its tags live in the low mantissa bits), which a deep bf16 network
        # may amplify through a flipped rounding - its deviation is reported, the bit-identity claim is checked on flags.
        os.environ["VTB_PAIR"] = "0"
        gt, _, _, _ = run(kind, dev, xs, ys, dist.group.WORLD, "p2p")
        os.environ["VTB_SYNC_TAGGED"] = "0"
        gq, _, _, _ = run(kind, dev, xs, ys, dist.group.WORLD, "p2p")
        os.environ.pop("VTB_SYNC_TAGGED")
        gn, ln, sn, path_n = run(kind, dev, xs, ys, dist.group.WORLD, "nccl")
        os.environ["VTB_PAIR"] = "1"
        err = float((gp - ref).norm() / ref.norm())
        cross = float((gq - gn).norm() / gn.norm())
        tagged_dev = float((gt - gn).norm() / gn.norm())
        lsum = torch.tensor([lp], device=dev)
        dist.all_reduce(lsum)
        loss_err = abs(float(lsum) / world - l1)
        serr = max(float((sp[k] - v).abs().max() / v.abs().max().clamp_min(1e-6)) for k, v in s1.items())
        # identical statistics on every rank
        vec = torch.cat([v.flatten() for v in sp.values()])
        lo, hi = vec.clone(), vec.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        ok = (err < tol and cross < 1e-5 and tagged_dev < tol and loss_err < 1e-3 and serr < 1e-3 and same
              and path_p == "peer-memory")
        ok_all &= ok
        print(f"rank {rank} [{kind}] SyncBN {path_p}: grads vs single-process global batch {err:.3e} (tol {tol}), "
              f"flag protocol vs {path_n} exchange {cross:.1e} (tagged protocol {tagged_dev:.1e}), loss diff {loss_err:.1e}, running stats {serr:.1e}, identical on all ranks "
              f"{same} -> {'OK' if ok else 'FAIL'}", flush=True)
    for kind, tol in (("unit", 1e-4), ("deep", 2e-3)):
        with vtb.precision("fp32"):
            ref, l1, s1, _ = run(kind, dev, X, Y, None, None)
            gp, lp, sp, _ = run(kind, dev, xs, ys, dist.group.WORLD, "nccl")
        err = float((gp - ref).norm() / ref.norm())
        lsum = torch.tensor([lp], device=dev)
        dist.all_reduce(lsum)
        loss_err = abs(float(lsum) / world - l1)
        serr = max(float((sp[k] - v).abs().max() / v.abs().max().clamp_min(1e-6)) for k, v in s1.items())
        ok = err < tol and loss_err < 1e-5 and serr < 1e-5
        ok_all &= ok
        print(f"rank {rank} [{kind}] fp32 mode: grads vs single-process global batch {err:.3e} (tol {tol}), loss diff "
              f"{loss_err:.1e}, running stats {serr:.1e} -> {'OK' if ok else 'FAIL'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
