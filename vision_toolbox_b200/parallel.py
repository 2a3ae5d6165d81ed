"""Data-parallel training step around a native backbone — the semantics the reference gets from Lightning
(``configs/base.yaml:16-23``: DDP gradient mean + SyncBatchNorm) and its trainer (``classifier.py:59-64``
head, ``:83-95`` step, ``:141-169`` SGD with weight decay only on conv/linear weights).

One process per GPU.  Two exchanges per step, both over NCCL/NVLink:
  * SyncBN: per BatchNorm layer the (sum, sum-of-squares) and (sum dz, sum dz*xhat) vectors are all-reduced
    inside the plan (engine.DistConfig) so every rank normalises with GLOBAL batch statistics;
  * gradients: all parameter gradients live in ONE flat fp32 buffer (parameter ``.grad`` are views), reduced
    in buckets and averaged.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .engine import DistConfig


class _HeadCEFn(torch.autograd.Function):
    """Pooling + linear head + label-smoothed cross entropy of the reference trainer (classifier.py:59-64, 92) as two
    C-ABI calls (``vtb_head_ce_fwd`` / ``vtb_head_ce_bwd``, 9 kernel launches per step instead of ~35 torch ones).

    `f`: the backbone's last feature map - a logical (N, C, H, W) bf16 tensor with NHWC strides (what the native backbone
    returns).  `sink`: write the head's parameter gradients straight into the existing fp32 ``.grad`` tensors (Trainer's
    flat all-reduce buffer, overwriting) instead of handing fresh tensors to autograd; `on_ready` is then told that they
    are final (the gradient exchange of their bucket can start while the backbone's backward runs)."""

    @staticmethod
    def forward(ctx, f, weight, bias, labels, label_smoothing: float, sink: bool, on_ready=None):
        from . import _lib
        from ._lib import check

        L = _lib.lib()
        n, c, h, w = f.shape
        k = weight.shape[0]
        assert f.dtype == torch.bfloat16 and f.stride(1) == 1 and f.stride(2) == w * f.stride(3), "NHWC bf16 view expected"
        assert weight.dtype == torch.float32 and weight.is_contiguous() and bias.dtype == torch.float32
        dev = f.device
        st = torch.cuda.current_stream(dev).cuda_stream
        pooled = torch.empty(n, c, dtype=torch.float32, device=dev)
        logits = torch.empty(n, k, dtype=torch.float32, device=dev)
        dlogits = torch.empty(n, k, dtype=torch.float32, device=dev)
        row_loss = torch.empty(n, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        labels = labels.to(torch.int64).contiguous()
        check(L.vtb_head_ce_fwd(f.data_ptr(), f.stride(3), n, h * w, c, weight.data_ptr(), bias.data_ptr(), k,
                                labels.data_ptr(), float(label_smoothing), pooled.data_ptr(), logits.data_ptr(),
                                dlogits.data_ptr(), row_loss.data_ptr(), loss.data_ptr(), st), "vtb_head_ce_fwd")
        ctx.save_for_backward(pooled, dlogits, weight, bias)
        ctx.geom = (n, c, h, w, k)
        ctx.sink = sink
        ctx.on_ready = on_ready   # callable([weight, bias]) once their gradients sit in .grad (Trainer's bucket overlap)
        ctx.need_df = f.requires_grad
        ctx.logits = logits
        return loss.view(())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gloss):
        from . import _lib
        from ._lib import check

        L = _lib.lib()
        pooled, dlogits, weight, bias = ctx.saved_tensors
        n, c, h, w, k = ctx.geom
        dev = pooled.device
        st = torch.cuda.current_stream(dev).cuda_stream
        g = gloss.detach().to(torch.float32).contiguous()
        direct = (ctx.sink and weight.grad is not None and bias.grad is not None and weight.grad.dtype == torch.float32
                  and weight.grad.is_contiguous() and bias.grad.is_contiguous())
        dW = weight.grad if direct else torch.empty_like(weight)
        db = bias.grad if direct else torch.empty_like(bias)
        df = torch.empty((n, c, h, w), dtype=torch.bfloat16, device=dev,
                         memory_format=torch.channels_last) if ctx.need_df else None
        scratch = torch.empty(n * k + n * c, dtype=torch.float32, device=dev)
        check(L.vtb_head_ce_bwd(pooled.data_ptr(), dlogits.data_ptr(), weight.data_ptr(), n, h * w, c, k, g.data_ptr(),
                                dW.data_ptr(), db.data_ptr(), 0, 0 if df is None else df.data_ptr(), c,
                                scratch.data_ptr(), st), "vtb_head_ce_bwd")
        if direct and ctx.on_ready is not None:
            ctx.on_ready([weight, bias])
        return df, (None if direct else dW), (None if direct else db), None, None, None, None


def sm_reserve() -> int:
    """SMs the SyncBN kernels leave free for whatever else must be able to start (``VTB_SM_RESERVE``, default 16; the
    library reads the same variable: csrc/elementwise.cu ``vtb_bn_bwd_fused``)."""
    import os

    try:
        return max(0, int(os.environ.get("VTB_SM_RESERVE", "16")))
    except ValueError:
        return 16


def nccl_pg_options():
    """``pg_options`` for ``dist.init_process_group("nccl", ...)``: keeps NCCL's kernels within the SMs that the SyncBN
    kernels leave free (``max_ctas`` = ``sm_reserve()``).  A SyncBN kernel that spins for its peers while holding every
    SM, and a collective on another stream that cannot start next to it, can close a cross-rank wait cycle (seen at 8
    GPUs); with the reserve NCCL can always start.  Returns None when the reserve is 0 or this torch build has no such
    option (NCCL's defaults then apply)."""
    import torch.distributed as dist

    r = sm_reserve()
    if r <= 0:
        return None
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = r
        return opts
    except Exception:  # noqa: BLE001
        return None


def split_decay_groups(modules: list[nn.Module]):
    """classifier.py:141-169 — weight decay on conv / linear weights only, none on norm parameters and biases."""
    decay, no_decay = [], []
    for root in modules:
        for m in root.modules():
            for name, p in m.named_parameters(recurse=False):
                if isinstance(m, (nn.Conv2d, nn.Linear)) and name == "weight":
                    decay.append(p)
                else:
                    no_decay.append(p)
    return decay, no_decay


class NativeSGD:
    """SGD with momentum (torch.optim.SGD semantics, dampening 0, no Nesterov; reference classifier.py:141-169) through the
    C ABI: ONE launch updates every convolution weight of the backbone and cuts the bf16 operands of the next forward from
    the new values (``vtb_sgd_pack_weights`` - no separate re-pack pass), ONE launch updates everything else
    (``vtb_sgd_step``: BatchNorm weights / biases, eSE and head parameters).  Gradients are read from the parameters'
    ``.grad`` (Trainer: views of its flat all-reduce buffer); lr / momentum live in device memory, so a schedule can change
    them under a captured CUDA graph (``set_lr``)."""

    def __init__(self, backbone: nn.Module, params: list, decay: list, lr: float, momentum: float, weight_decay: float):
        self.backbone, self.params = backbone, params
        self.decay_ids = {id(p) for p in decay}
        self.weight_decay = weight_decay
        dev = params[0].device
        self.hyper = torch.tensor([lr, momentum], dtype=torch.float32, device=dev)
        self.mom = {id(p): torch.zeros_like(p, memory_format=torch.contiguous_format) for p in params}
        self._runner = None
        self._tables = None
        self._bucket_of, self._nb = {}, 1
        self._done: set = set()

    def set_lr(self, lr: float) -> None:
        self.hyper[0] = lr

    def set_buckets(self, bucket_of: dict, n: int) -> None:
        """Group the parameters (id -> bucket index, `n` buckets): `step_bucket(b)` then updates one group on its own, so a
        trainer can step every gradient bucket as soon as it is final (and reduced) while backward still runs."""
        self._bucket_of, self._nb = dict(bucket_of), int(n)
        self._tables = None

    def _build(self, runner) -> None:
        import ctypes as C

        from . import _lib

        L = _lib.lib()
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.grad is None or p.grad.dtype != torch.float32 \
                    or not p.grad.is_contiguous():
                raise TypeError("NativeSGD needs contiguous float32 parameters and gradients")
        wd = lambda p: self.weight_decay if id(p) in self.decay_ids else 0.0
        packed = {id(op.mod.conv.weight) for op in runner._conv_ops}
        srcs = [op.mod.conv.weight.detach() for op in runner._conv_ops]
        sgd = lambda p: (p.grad.data_ptr(), self.mom[id(p)].data_ptr(), wd(p))
        bucket = lambda p: self._bucket_of.get(id(p), 0) if self._nb > 1 else 0
        groups = []
        for b in range(max(1, self._nb)):
            idx = [i for i, w in enumerate(srcs) if bucket(runner._conv_ops[i].mod.conv.weight) == b]
            table, launches = runner.pack_job_table(srcs, sgd=sgd, subset=idx) if idx else (None, [])
            rest = [p for p in self.params if id(p) not in packed and bucket(p) == b]
            plain, blk, plain_launches = (_lib.VtbSgdJob * max(1, len(rest)))(), 0, []
            for j, p in enumerate(rest):
                plain[j] = _lib.VtbSgdJob(p.data_ptr(), p.grad.data_ptr(), self.mom[id(p)].data_ptr(), p.numel(), wd(p), blk)
                blk += int(L.vtb_sgd_job_blocks(p.numel()))
                if (j + 1) % 256 == 0 or j == len(rest) - 1:
                    plain_launches.append((j // 256 * 256, j % 256 + 1, blk))
                    blk = 0
            ptab = torch.frombuffer(bytearray(bytes(plain)), dtype=torch.uint8).to(self.hyper.device)
            groups.append((table, launches, ptab, plain_launches if rest else []))
        self._runner = runner
        self._tables = (groups, C.sizeof(_lib.VtbPackJob), C.sizeof(_lib.VtbSgdJob))
        self._key = self._current_key()
        self._done = set()

    def _current_key(self):
        return tuple(p.data_ptr() for p in self.params) + tuple(p.grad.data_ptr() for p in self.params)

    def _training_runner(self):
        plans = self.backbone.__dict__.get("_vtb_plans", {})
        return next((r for r in plans.values() if r.g.need_grad and not r.g.f32), None)

    def ready(self) -> bool:
        """True when `step_bucket` can run right now: the job tables exist and still describe the live tensors (they are
        built by the first `step()`)."""
        return (self._tables is not None and self._runner is self._training_runner() and self._runner is not None
                and self._key == self._current_key())

    def step_bucket(self, b: int) -> None:
        """Update the parameters of bucket `b` on the current stream (tables must be `ready()`); `step()` skips them."""
        from . import _lib
        from ._lib import check

        L = _lib.lib()
        groups, rec, prec = self._tables
        table, launches, ptab, plain_launches = groups[b]
        dev = self.hyper.device
        st = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0
        for j0, n, blocks in launches:
            check(L.vtb_sgd_pack_weights(table.data_ptr() + j0 * rec, n, blocks, self.hyper.data_ptr(), st),
                  "vtb_sgd_pack_weights")
        for j0, n, blocks in plain_launches:
            check(L.vtb_sgd_step(ptab.data_ptr() + j0 * prec, n, blocks, self.hyper.data_ptr(), st), "vtb_sgd_step")
        self._done.add(b)

    def step(self) -> None:
        runner = self._training_runner()
        if runner is None:
            raise RuntimeError("NativeSGD.step() needs a native training plan (run forward + backward first)")
        if self._tables is None or self._runner is not runner or self._key != self._current_key():
            if self._done:
                raise RuntimeError("NativeSGD: parameters or gradients were re-allocated in the middle of a step")
            self._build(runner)
        for b in range(len(self._tables[0])):
            if b not in self._done:
                self.step_bucket(b)
        self._done = set()
        plans = self.backbone.__dict__.get("_vtb_plans", {})
        # the operands of `runner` are current; other plans of the module (other shapes / modes) re-pack as usual
        for r in plans.values():
            r._packs_token = r.pack_token() if r is runner else None


def bucket_ranges(sizes: list[int], bucket_elems: int) -> list[tuple[int, int]]:
    """Greedy contiguous buckets over a flat buffer: [(start, end)] in elements, each >= bucket_elems except the last."""
    out, start, acc = [], 0, 0
    for s in sizes:
        acc += s
        if acc - start >= bucket_elems:
            out.append((start, acc))
            start = acc
    if acc > start:
        out.append((start, acc))
    return out


def overlap_bucket_ranges(sizes: list[int], bucket_elems: int, last_elems: int) -> list[tuple[int, int]]:
    """Bucket layout for an exchange overlapped with backward.  Backward finishes parameters from the END of the flat
    buffer towards its start, so the bucket at the START completes last and its all-reduce is the one nothing can hide:
    it is kept small (about `last_elems`, cut at a parameter boundary); the rest is cut greedily from the end into buckets
    of >= `bucket_elems` (the leftover joins the bucket next to the small one).  [(start, end)] in elements, ascending."""
    total = sum(sizes)
    if not sizes:
        return []
    head_end, j = 0, 0
    while j < len(sizes) and (head_end == 0 or head_end + sizes[j] <= last_elems):
        head_end += sizes[j]
        j += 1
    if head_end >= total:
        return [(0, total)]
    cuts, acc, end = [], 0, total
    for s in reversed(sizes[j:]):
        acc += s
        if acc >= bucket_elems:
            cuts.append((end - acc, end))
            end -= acc
            acc = 0
    if acc > 0:
        if cuts and acc < bucket_elems // 2:
            a, e = cuts.pop()
            cuts.append((head_end, e))       # a small leftover joins its neighbour
        else:
            cuts.append((head_end, end))
    return [(0, head_end)] + cuts[::-1]


class Trainer:
    """One data-parallel training step: forward, label-smoothed CE, backward, gradient mean, SGD.

    Gradient exchange: every parameter ``.grad`` is a view of ONE flat fp32 buffer, cut into ~``bucket_mb`` buckets.
    The native backward reports which parameter gradients are final (reverse layer order); as soon as a bucket is
    complete its all-reduce is enqueued on a side stream, so the exchange overlaps the rest of backward (what DDP's
    reducer does for the reference, configs/base.yaml:19).  The native head reports its gradients first (they are the
    first ones backward produces); the bucket that completes last - the stem side - is kept small (`last_bucket_mb`),
    because its all-reduce is the only one that nothing overlaps.
    """

    def __init__(self, backbone: nn.Module, head: nn.Module, lr: float = 0.05, momentum: float = 0.9,
                 weight_decay: float = 2e-5, label_smoothing: float = 0.1, sync_bn: bool = True,
                 process_group=None, bucket_mb: float = 25.0, mixup_cutmix: Optional[nn.Module] = None,
                 last_bucket_mb: float = 1.0):
        self.backbone, self.head = backbone, head
        self.label_smoothing = label_smoothing
        # classifier.py:66-67, 86-87: RandomCutMixMixUp on the batch before the forward (vision_toolbox_b200.extras: sampled
        # and applied on the device, capturable); targets become probability vectors
        self.mixup_cutmix = mixup_cutmix
        self.group = process_group
        self.world = 1
        self.params = [p for p in list(backbone.parameters()) + list(head.parameters()) if p.requires_grad]
        dev = self.params[0].device
        self.dist_cfg: Optional[DistConfig] = None
        self.avg_in_collective = False
        self._graph = None
        self.graph_launches = 0
        import os

        # classifier head + loss through the C ABI (vtb_head_ce_*: pooling, 64-wide fp32 GEMM tiles, label-smoothed CE and
        # their backward in 8 launches; parity: tests/test_gpu_parity.py) instead of ~35 torch / cuBLAS kernels, so that the
        # step contains no library kernel.  VTB_NATIVE_HEAD=0 selects the torch composition.
        self.native_head = (os.environ.get("VTB_NATIVE_HEAD", "1") == "1" and isinstance(head, nn.Linear)
                            and head.bias is not None and head.weight.dtype == torch.float32
                            and head.weight.device.type == "cuda")
        self._seed = None
        self._zeroed = True
        self._zero_needed = True   # until a step has shown that every gradient is overwritten in place (see _step_eager)
        self._used_native_head = False
        if process_group is not None:
            import torch.distributed as dist

            self.world = dist.get_world_size(process_group)
            # DDP broadcasts rank 0's parameters and buffers when it wraps the module
            for t in list(backbone.state_dict().values()) + list(head.state_dict().values()):
                dist.broadcast(t, src=0, group=process_group)
            self.dist_cfg = DistConfig(process_group, sync_bn=sync_bn, device=dev)
            self.dist_cfg.on_grads_ready = self._grads_ready
            backbone.__dict__["_vtb_dist"] = self.dist_cfg
            self.avg_in_collective = dist.get_backend(process_group) == "nccl"
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        off = 0
        self._offset = {}
        for p, n in zip(self.params, sizes):
            p.grad = self.flat[off : off + n].view_as(p)
            self._offset[id(p)] = off
            off += n
        # the native backward writes backbone gradients straight into these views (engine.Runner grad_sink mode)
        backbone.__dict__["_vtb_grad_sink"] = True
        self.buckets = overlap_bucket_ranges(sizes, int(bucket_mb * 1024 * 1024 / 4),
                                             int(min(last_bucket_mb, bucket_mb) * 1024 * 1024 / 4))
        # bucket bookkeeping for the overlap: parameters per bucket, the bucket of every parameter
        starts = [a for a, _ in self.buckets]
        self._bucket_of = {}
        self._bucket_size = [0] * len(self.buckets)
        import bisect

        for p in self.params:
            b = bisect.bisect_right(starts, self._offset[id(p)]) - 1
            self._bucket_of[id(p)] = b
            self._bucket_size[b] += 1
        self._pending = list(self._bucket_size)
        self._launched = [False] * len(self.buckets)
        self._works = []
        self.comm_stream = torch.cuda.Stream(dev) if (dev.type == "cuda" and self.world > 1) else None
        decay, no_decay = split_decay_groups([backbone, head])
        # optimizer: the native fused SGD + operand re-pack on CUDA (VTB_NATIVE_SGD=0: torch's fused SGD + a re-pack launch
        # at the start of every forward), torch SGD for CPU modules
        self.native_sgd = (dev.type == "cuda" and os.environ.get("VTB_NATIVE_SGD", "1") == "1"
                           and all(p.dtype == torch.float32 for p in self.params))
        # optimizer overlap (opt-in, VTB_SGD_OVERLAP=1): a bucket's parameters are updated as soon as its gradients are
        # final (and reduced) instead of after backward.  Safe - every reader of a layer's weights in this step (its dgrad)
        # is enqueued before the layer reports, and the updating stream waits for that point of the main stream - but
        # measured neutral (13.89 vs 13.85 ms on one GPU, 14.483 vs 14.487 ms on two; profiles/r02_bench_*_t14*.json): the
        # update is HBM-bound and so is the end of backward (the 88^2 / 176^2 layers) it would have to hide under.
        self.sgd_overlap = False
        self._sgd_now = False
        self._main_stream = None
        if self.native_sgd:
            self.opt = NativeSGD(backbone, self.params, decay, lr, momentum, weight_decay)
            self.sgd_overlap = (os.environ.get("VTB_SGD_OVERLAP", "0") == "1" and self.native_head
                                and (self.world == 1 or self.avg_in_collective))
            if self.sgd_overlap:
                self.opt.set_buckets(self._bucket_of, len(self.buckets))
                if self.world == 1:
                    backbone.__dict__["_vtb_on_grads_ready"] = self._grads_ready
        else:
            self.opt = torch.optim.SGD(
                [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}],
                lr=lr, momentum=momentum, fused=dev.type == "cuda")

    # -- gradient exchange
    def _after_main(self, stream) -> None:
        """Make `stream` wait for everything the step's main stream has been given so far (the dgrads that read the weights a
        bucket update is about to overwrite)."""
        if self._main_stream is not None and stream != self._main_stream:
            ev = torch.cuda.Event()
            ev.record(self._main_stream)
            stream.wait_event(ev)

    def _launch_bucket(self, b: int) -> None:
        a, e = self.buckets[b]
        self._launched[b] = True
        if self.world == 1:
            # single process: nothing to exchange - the bucket's optimizer step, on the stream its gradients were produced on
            if self._sgd_now:
                self._after_main(torch.cuda.current_stream())
                self.opt.step_bucket(b)
            return
        import torch.distributed as dist

        op = dist.ReduceOp.AVG if self.avg_in_collective else dist.ReduceOp.SUM
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                work = dist.all_reduce(self.flat[a:e], op=op, group=self.group, async_op=True)
                self._works.append(work)
                if self._sgd_now:
                    self._after_main(self.comm_stream)
                    work.wait()                 # orders the exchange stream (not the host) behind the collective
                    self.opt.step_bucket(b)
        else:
            self._works.append(dist.all_reduce(self.flat[a:e], op=op, group=self.group, async_op=True))

    def _grads_ready(self, params) -> None:
        """Called by the native backward when the gradients of `params` are final (written to the flat buffer)."""
        for p in params:
            b = self._bucket_of.get(id(p))
            if b is None:
                continue
            self._pending[b] -= 1
            if self._pending[b] == 0 and not self._launched[b]:
                self._launch_bucket(b)

    def _finish_exchange(self) -> None:
        for b in range(len(self.buckets) - 1, -1, -1):   # whatever the overlap did not cover (the head, CPU modules)
            if not self._launched[b]:
                self._launch_bucket(b)
        if self.world > 1:
            for w in self._works:
                w.wait()
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            if not self.avg_in_collective:
                self.flat.mul_(1.0 / self.world)
        self._works = []
        self._pending = list(self._bucket_size)
        self._launched = [False] * len(self.buckets)

    def forward_loss(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if self.mixup_cutmix is not None:
            x, y = self.mixup_cutmix(x, y)   # y: (N, classes) probabilities from here on
        f = self.backbone(x)  # (N, C, H, W) bf16 on CUDA
        if self.native_head and y.ndim == 1 and f.is_cuda and f.dtype == torch.bfloat16 and f.shape[1] % 8 == 0:
            # pooling + linear + label-smoothed CE in the native library (head gradients land in the flat buffer)
            self._used_native_head = True
            ready = self._grads_ready if (self.world > 1 or self.sgd_overlap) else None
            return _HeadCEFn.apply(f, self.head.weight, self.head.bias, y, self.label_smoothing, True, ready)
        self._used_native_head = False
        if not self._zeroed:
            # the flat buffer was not zeroed (the previous step overwrote every gradient in place), but autograd ACCUMULATES
            # the torch head's gradients
            for p in self.head.parameters():
                if p.grad is not None:
                    p.grad.zero_()
        pooled = f.float().mean(dim=(2, 3))  # AdaptiveAvgPool2d + Flatten (classifier.py:61-62)
        logits = self.head(pooled)
        return F.cross_entropy(logits, y, label_smoothing=self.label_smoothing)

    def _all_grads_overwritten(self) -> bool:
        """True when the step just run wrote EVERY gradient in place (native head + a native backbone plan in sink mode):
        autograd accumulated nothing, so the flat buffer needs no zeroing before the next step (never-written entries keep
        their initial zeros)."""
        if not self._used_native_head:
            return False
        plans = self.backbone.__dict__.get("_vtb_plans", {})
        runners = [r for r in plans.values() if r.g.need_grad]
        return bool(runners) and all(r.last_all_direct for r in runners)

    def _step_eager(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self._zeroed = self._zero_needed
        if self._zero_needed:
            self.flat.zero_()
        if self.sgd_overlap:
            self._main_stream = torch.cuda.current_stream()
            # per-bucket updates need the optimizer's job tables (built by the first step) and in-place gradients
            self._sgd_now = (not self._zero_needed) and self.opt.ready()
        loss = self.forward_loss(x, y)
        if self._seed is None or self._seed.device != loss.device or self._seed.dtype != loss.dtype:
            self._seed = torch.ones((), dtype=loss.dtype, device=loss.device)
        loss.backward(self._seed)   # a cached seed: autograd would launch a fill kernel for its own every step
        self._zero_needed = not self._all_grads_overwritten()
        if self.world > 1 or self.sgd_overlap:
            self._finish_exchange()
        self.opt.step()
        self._sgd_now = False
        return loss.detach()

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if self._graph is not None and tuple(x.shape) == tuple(self._gx.shape) and x.dtype == self._gx.dtype:
            # the whole step (≈600 kernel launches) replays from ONE CUDA graph: inputs are copied into the captured
            # buffers, the loss is read from the captured output
            self._gx.copy_(x, non_blocking=True)
            self._gy.copy_(y, non_blocking=True)
            self._graph.replay()
            return self._gloss
        return self._step_eager(x, y)

    def enable_cuda_graph(self, x: torch.Tensor, y: torch.Tensor, warmup: int = 3) -> None:
        """Capture forward + loss + backward + SGD for inputs shaped like (x, y) into a CUDA graph (single process).

        Every kernel of the step is enqueued through the C ABI on the capturing stream, all buffers come from the
        graph's private pool, parameter gradients land in the flat buffer and the SyncBN/BatchNorm tickets are device
        resident, so a replay is exactly one training step.  `warmup` eager steps run first (they are REAL steps).
        """
        dev = x.device
        if self.world > 1:
            # NCCL all-reduces (side stream, forked from / joined to the capturing stream) and the SyncBN peer-memory
            # kernels (device-resident sequence numbers) are capturable; the NCCL-per-layer SyncBN fallback is not used
            # under capture because it reads tensors back on the host path of torch.distributed
            if self.dist_cfg is not None and self.dist_cfg.sync_bn and self.dist_cfg.sync is None:
                raise NotImplementedError("CUDA-graph capture needs the peer-memory SyncBN exchange")
        self._gx, self._gy = x.clone(), y.clone()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 2)):   # builds the plan, the weight packs and the momentum buffers
                self._step_eager(self._gx, self._gy)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib

        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            self._gloss = self._step_eager(self._gx, self._gy)
        self.graph_launches = _lib.launch_count() - n0   # kernels of this library inside one replay
        self._graph = graph
