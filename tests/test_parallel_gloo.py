"""Data-parallel host logic on CPU: world_size-2 `gloo` runs of parallel.Trainer (bucketed gradient mean, parameter
broadcast, identical replicas after a step) plus the bucket bookkeeping that drives the all-reduce/backward overlap."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vision_toolbox_b200 import parallel
from vision_toolbox_b200.backbones import Darknet
from vision_toolbox_b200.backbones.darknet import CSPDarknetStage


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(seed: int):
    torch.manual_seed(seed)
    m = Darknet(16, [(1, 32), (1, 32)], CSPDarknetStage)
    head = torch.nn.Linear(32, 10)
    return m, head


def _data(rank: int):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.rand(4, 3, 32, 32, generator=g), torch.randint(0, 10, (4,), generator=g)


def _local_grads(state, head_state, rank):
    """Gradients one replica computes on its own shard (no SyncBN on the CPU path), as a flat vector."""
    m, head = _model(0)
    m.load_state_dict(state); head.load_state_dict(head_state)
    tr = parallel.Trainer(m, head, lr=0.0, momentum=0.0, weight_decay=0.0)
    x, y = _data(rank)
    loss = tr.forward_loss(x, y)
    loss.backward()
    return tr.flat.clone()


def _worker(rank: int, world: int, port: int, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, head = _model(seed=rank)            # DIFFERENT init per rank: the Trainer must broadcast rank 0's
        tr = parallel.Trainer(m, head, lr=0.1, momentum=0.9, weight_decay=1e-4, process_group=dist.group.WORLD,
                              bucket_mb=0.01)   # tiny buckets -> several of them
        assert len(tr.buckets) > 3
        state0 = {k: v.clone() for k, v in m.state_dict().items()}
        head0 = {k: v.clone() for k, v in head.state_dict().items()}
        ref0, _ = _model(seed=0)
        for k, v in ref0.state_dict().items():
            assert torch.equal(state0[k], v), f"rank {rank}: {k} was not broadcast from rank 0"
        x, y = _data(rank)
        # expected: mean over ranks of the per-replica gradients
        expect = sum(_local_grads(state0, head0, r) for r in range(world)) / world
        tr.flat.zero_()
        loss = tr.forward_loss(x, y)
        loss.backward()
        tr._finish_exchange()
        err = float((tr.flat - expect).abs().max() / expect.abs().max())
        assert err < 1e-6, err
        assert all(not l for l in tr._launched) and tr._pending == tr._bucket_size   # bookkeeping reset
        # a full step keeps the replicas identical
        tr.step(x, y)
        vec = torch.cat([p.detach().flatten() for p in tr.params])
        gathered = [torch.empty_like(vec) for _ in range(world)]
        dist.all_gather(gathered, vec)
        assert torch.equal(gathered[0], gathered[1])
        out.put((rank, "ok", err))
    except Exception as e:  # noqa: BLE001
        out.put((rank, f"{type(e).__name__}: {e}", None))
    finally:
        dist.destroy_process_group()


def test_trainer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_bucket_ranges_cover_everything_in_order():
    sizes = [5, 100, 7, 300, 2, 2, 50]
    b = parallel.bucket_ranges(sizes, 100)
    assert b[0][0] == 0 and b[-1][1] == sum(sizes)
    assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
    assert all(e - a >= 100 for a, e in b[:-1])


def test_overlap_bucket_ranges_keep_the_last_finishing_bucket_small():
    """Backward completes the flat buffer from its end: the bucket at the start is the exposed one and must be small."""
    sizes = [5, 100, 7, 300, 2, 2, 50]
    for be, le in [(100, 10), (100, 120), (50, 1), (10_000, 10)]:
        b = parallel.overlap_bucket_ranges(sizes, be, le)
        assert b[0][0] == 0 and b[-1][1] == sum(sizes)
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        bounds = {0}
        acc = 0
        for s_ in sizes:
            acc += s_
            bounds.add(acc)
        assert all(a in bounds and e in bounds for a, e in b)          # cut at parameter boundaries only
        assert b[0][1] - b[0][0] <= max(le, sizes[0])                  # the exposed bucket is small
        assert all(e - a >= min(be, sum(sizes) - b[0][1]) // 2 for a, e in b[1:])   # no confetti behind it
    assert parallel.overlap_bucket_ranges([7], 100, 10) == [(0, 7)]
    assert parallel.overlap_bucket_ranges([], 100, 10) == []


def test_sm_reserve_and_nccl_options(monkeypatch):
    """VTB_SM_RESERVE is read by both sides (library: BatchNorm-backward grid; Python: NCCL's CTA cap).  0 switches the cap
    off - NCCL rejects maxCTAs = 0, so no options object must be produced - and garbage falls back to the default."""
    monkeypatch.setenv("VTB_SM_RESERVE", "0")
    assert parallel.sm_reserve() == 0 and parallel.nccl_pg_options() is None
    monkeypatch.setenv("VTB_SM_RESERVE", "nonsense")
    assert parallel.sm_reserve() == 16
    monkeypatch.setenv("VTB_SM_RESERVE", "24")
    assert parallel.sm_reserve() == 24
    opts = parallel.nccl_pg_options()
    if opts is not None:                       # torch builds without ProcessGroupNCCL.Options give None
        assert opts.config.max_ctas == 24


def test_weight_decay_groups_follow_reference_rule():
    m, head = _model(0)
    decay, no_decay = parallel.split_decay_groups([m, head])
    n_conv = sum(1 for mod in m.modules() if isinstance(mod, torch.nn.Conv2d))
    assert len(decay) == n_conv + 1                      # conv weights + the linear weight (classifier.py:141-169)
    assert all(p.dim() in (2, 4) for p in decay)
    assert len(decay) + len(no_decay) == len(list(m.parameters())) + len(list(head.parameters()))


def test_overlap_bookkeeping_launches_each_bucket_once_when_complete():
    m, head = _model(0)
    tr = parallel.Trainer(m, head, bucket_mb=0.01)
    launched = []
    tr._launch_bucket = lambda b: (launched.append(b), tr._launched.__setitem__(b, True))
    backbone_params = list(m.parameters())
    for i in range(len(backbone_params) - 1, -1, -1):    # reverse layer order, like the native backward
        before = list(launched)
        tr._grads_ready([backbone_params[i]])
        for b in set(launched) - set(before):
            a, e = tr.buckets[b]
            # every parameter of a launched bucket has been reported
            assert all(tr._offset[id(p)] >= a or tr._offset[id(p)] + p.numel() <= a for p in backbone_params[:i])
            assert tr._pending[b] == 0
    assert len(launched) == len(set(launched))
    head_bucket = tr._bucket_of[id(head.weight)]
    assert set(launched) >= set(range(len(tr.buckets))) - {head_bucket, tr._bucket_of[id(head.bias)]}
