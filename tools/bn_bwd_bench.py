"""Per-launch time of the BatchNorm-backward entry points on one tensor shape (CUDA-graph replay of 20 back-to-back
launches, host-free):   python tools/bn_bwd_bench.py"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from vision_toolbox_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda", 0)
for c, pix in ((512, 9216), (1024, 9216), (256, 30976), (128, 123904), (64, 495616), (32, 1982464)):
    dout = torch.randn(pix, c, device=dev).bfloat16()
    y = torch.randn(pix, c, device=dev).bfloat16()
    dy = torch.empty_like(y)
    out = torch.empty_like(y)
    f = lambda n=c: torch.rand(n, device=dev) + 0.5
    scale, shift, mean, invstd = f(), f(), f(), f()
    rows = max(L.vtb_bn_bwd_rows(pix, c), L.vtb_bn_bwd_fused_rows(pix, c))
    partial = torch.zeros((rows + 1) * c * 2, device=dev)
    dgamma, dbeta, coef = torch.zeros(c, device=dev), torch.zeros(c, device=dev), torch.zeros(2 * c, device=dev)
    sync = torch.zeros(512, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(dev)

    def fused(s):
        _lib.check(L.vtb_bn_bwd_fused(dout.data_ptr(), c, y.data_ptr(), c, pix, c, scale.data_ptr(), shift.data_ptr(),
                                      mean.data_ptr(), invstd.data_ptr(), 1, float(pix), partial.data_ptr(), dgamma.data_ptr(),
                                      dbeta.data_ptr(), 0, sync.data_ptr(), dy.data_ptr(), c, None, s))

    def split(s):
        _lib.check(L.vtb_bn_bwd_reduce(dout.data_ptr(), c, y.data_ptr(), c, pix, c, scale.data_ptr(), shift.data_ptr(),
                                       mean.data_ptr(), invstd.data_ptr(), 1, partial.data_ptr(), s))
        _lib.check(L.vtb_bn_bwd_finalize(partial.data_ptr(), L.vtb_bn_bwd_rows(pix, c), None, None, float(pix), c,
                                         dgamma.data_ptr(), dbeta.data_ptr(), 0, coef.data_ptr(), None, s))
        _lib.check(L.vtb_bn_bwd_apply(dout.data_ptr(), c, y.data_ptr(), c, pix, c, scale.data_ptr(), shift.data_ptr(),
                                      mean.data_ptr(), invstd.data_ptr(), 1, coef.data_ptr(), dy.data_ptr(), c, s))

    def act(s):
        _lib.check(L.vtb_bn_act(y.data_ptr(), c, pix, c, scale.data_ptr(), shift.data_ptr(), 1, None, 0, out.data_ptr(), c, s))

    res = {}
    for name, fn in (("bn_bwd_fused", fused), ("reduce+finalize+apply", split), ("bn_act", act)):
        with torch.cuda.stream(st):
            fn(st.cuda_stream)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(20):
                    fn(st.cuda_stream)
            g.replay(); st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(5):
                g.replay()
            e1.record(st)
            st.synchronize()
            res[name] = e0.elapsed_time(e1) * 1e3 / 100
    mb = pix * c * 2 / 1e6
    print(f"c {c:5d} pix {pix:8d} ({mb:6.1f} MB/tensor): " + "  ".join(f"{k} {v:7.1f} us" for k, v in res.items()), flush=True)
