"""A CPU INTERPRETER of the C ABI in include/vtb.h, for tests only.

tests/test_dry_run_plan.py checks WHERE the executor's launches read and write; this stand-in also gives them meaning:
every compute entry point of the bf16 training / eval plans is implemented with plain torch on the CPU, operating on the
very buffers the executor allocated (raw pointers -> torch views).  Running `engine.Runner.forward / backward` against it
therefore yields real numbers without a GPU, which tests/test_cpu_interpreter.py compares with the reference's golden
vectors.  What this verifies is the HOST side - the plan (concat slices, residual aliasing, gradient routing and
accumulate flags, sibling pairing, the gathered-operand stem and its weight-gradient permutation, the pack job table) -
against the semantics documented in vtb.h; the CUDA kernels are verified against the same goldens by tests/test_gpu_*.py.
It is never imported by the package.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from vision_toolbox_b200 import _lib

BF16 = torch.bfloat16
HOST_ONLY = {"vtb_last_error", "vtb_conv_stats_rows", "vtb_bn_bwd_rows", "vtb_bn_bwd_fused_rows", "vtb_conv_out_hw",
             "vtb_conv_wgrad_workspace_bytes", "vtb_f32_conv_wgrad_workspace_bytes", "vtb_f32_bn_rows",
             "vtb_pack_job_blocks", "vtb_launch_count", "vtb_version", "vtb_num_sms", "vtb_bn_sync_buffer_bytes",
             "vtb_conv_dgrad_stats_rows", "vtb_conv_dgrad_panel_w", "vtb_conv_tiling_info", "vtb_sgd_job_blocks",
             "vtb_conv_dgrad_s2_workspace_bytes"}


def _flat(ptr: int, n: int, dt: torch.dtype) -> torch.Tensor:
    es = torch.empty((), dtype=dt).element_size()
    return torch.frombuffer((C.c_uint8 * (n * es)).from_address(ptr), dtype=dt)


def _view(ptr: int, ld: int, pixels: int, c: int, dt: torch.dtype = BF16) -> torch.Tensor:
    """[pixels, c] view with pixel pitch `ld` over raw memory (writes go through)."""
    return _flat(ptr, (pixels - 1) * ld + c, dt).as_strided((pixels, c), (ld, 1))


def _obj(a):
    return a._obj if hasattr(a, "_obj") else a


def _r(t: torch.Tensor) -> torch.Tensor:
    return t.to(BF16).float()


class InterpreterLib:
    def __init__(self):
        self.real = _lib.lib()
        self.calls = []

    def __getattr__(self, name):
        if name in HOST_ONLY:
            return getattr(self.real, name)
        impl = getattr(self, "_" + name, None)
        if impl is None:
            raise AssertionError(f"the CPU interpreter has no implementation of {name}")

        def call(*args):
            self.calls.append(name)
            impl(*args)
            return 0

        return call

    # ---------------------------------------------------------------- layout / packing
    def _vtb_nchw_to_nhwc(self, x, n, c, h, w, out, cpad, st):
        src = _flat(x, n * c * h * w, torch.float32).view(n, c, h, w)
        dst = _view(out, cpad, n * h * w, cpad)
        dst.zero_()
        dst[:, :c] = src.permute(0, 2, 3, 1).reshape(-1, c).to(BF16)

    def _vtb_im2col_input(self, x, n, c, h, w, k, stride, pad, out, kp, st):
        src = _flat(x, n * c * h * w, torch.float32).view(n, c, h, w)
        cols = F.unfold(src, k, padding=pad, stride=stride)            # [n, c*k*k, L], rows ordered (ci, tap)
        L = cols.shape[-1]
        cols = cols.view(n, c, k * k, L).permute(0, 3, 2, 1).reshape(n * L, k * k * c)   # -> (tap, ci) columns
        dst = _view(out, kp, n * L, kp)
        dst.zero_()
        dst[:, : k * k * c] = cols.to(BF16)

    def _vtb_sgd_pack_weights(self, jobs, njobs, blocks, hyper, st):
        lr, mu = _flat(hyper, 2, torch.float32).tolist()
        for j in (_lib.VtbPackJob * njobs).from_address(jobs):
            n = j.cout * j.cin_real * j.kk
            self._sgd(_flat(j.w, n, torch.float32), _flat(j.g, n, torch.float32), _flat(j.m, n, torch.float32),
                      j.weight_decay, lr, mu)
        self._vtb_pack_weights(jobs, njobs, blocks, st)

    def _vtb_sgd_step(self, jobs, njobs, blocks, hyper, st):
        lr, mu = _flat(hyper, 2, torch.float32).tolist()
        for j in (_lib.VtbSgdJob * njobs).from_address(jobs):
            self._sgd(_flat(j.w, j.n, torch.float32), _flat(j.g, j.n, torch.float32), _flat(j.m, j.n, torch.float32),
                      j.weight_decay, lr, mu)

    @staticmethod
    def _sgd(w, g, m, wd, lr, mu):
        m.mul_(mu).add_(g + wd * w)
        w.sub_(lr * m)

    def _vtb_pack_weights(self, jobs, njobs, blocks, st):
        table = (_lib.VtbPackJob * njobs).from_address(jobs)
        for j in table:
            w = _flat(j.w, j.cout * j.cin_real * j.kk, torch.float32).view(j.cout, j.cin_real, j.kk)
            wp = torch.zeros(j.cout, j.kk, j.cin)
            wp[:, :, : j.cin_real] = w.permute(0, 2, 1)
            _view(j.wf, j.wf_ld, j.cout, j.kk * j.cin).copy_(wp.reshape(j.cout, -1).to(BF16))
            if j.wd:
                # wd[(ci * kk + t) * wd_ld + wd_co_off + co]
                _view(j.wd + 2 * j.wd_co_off, j.wd_ld, j.cin * j.kk, j.cout).copy_(
                    wp.permute(2, 1, 0).reshape(j.cin * j.kk, j.cout).to(BF16))

    # ---------------------------------------------------------------- convolution family
    @staticmethod
    def _hw(g):
        return (g.h + 2 * g.pad - g.k) // g.stride + 1, (g.w + 2 * g.pad - g.k) // g.stride + 1

    def _conv(self, g, x, ldx, wf):
        ho, wo = self._hw(g)
        xi = _view(x, ldx, g.n * g.h * g.w, g.cin).float().view(g.n, g.h, g.w, g.cin).permute(0, 3, 1, 2)
        w = _flat(wf, g.cout * g.k * g.k * g.cin, BF16).float().view(g.cout, g.k, g.k, g.cin).permute(0, 3, 1, 2)
        y = F.conv2d(xi, w, None, g.stride, g.pad)
        return y.permute(0, 2, 3, 1).reshape(g.n * ho * wo, g.cout)

    def _vtb_conv_fprop(self, geom, x, ldx, wf, y, ldy, stats, scale, shift, relu, res, ldr, st):
        g = _obj(geom)
        out = self._conv(g, x, ldx, wf)
        if scale:
            out = out * _flat(scale, g.cout, torch.float32) + _flat(shift, g.cout, torch.float32)
        if relu:
            out = out.clamp_min(0)
        if res:
            out = _r(out) + _view(res, ldr, out.shape[0], g.cout).float()
        _view(y, ldy, out.shape[0], g.cout).copy_(out.to(BF16))
        if stats:
            yb = _view(y, ldy, out.shape[0], g.cout).double()
            rows = self.real.vtb_conv_stats_rows(C.byref(g))
            part = _flat(stats, rows * g.cout * 2, torch.float32).view(rows, g.cout, 2)
            part.zero_()
            part[0, :, 0] = yb.sum(0).float()
            part[0, :, 1] = (yb * yb).sum(0).float()

    def _vtb_conv_fprop_bn(self, geom, x, ldx, wf, y, ldy, partial, bn, st):
        g, b = _obj(geom), _obj(bn)
        assert b.sync is None, "the interpreter models a single rank"
        out = self._conv(g, x, ldx, wf)
        yv = _view(y, ldy, out.shape[0], g.cout)
        yv.copy_(out.to(BF16))
        yb = yv.double()
        mean = yb.sum(0) / b.count
        var = ((yb * yb).sum(0) / b.count - mean * mean).clamp_min(0)
        invstd = (1.0 / torch.sqrt(var + b.eps)).float()
        for c0, c1, gam, bet, rm, rv, nbt in self._bn_sets(g.cout, b):
            n = c1 - c0
            gamma, beta = _flat(gam, n, torch.float32), _flat(bet, n, torch.float32)
            sc = gamma * invstd[c0:c1]
            _flat(b.mean + 4 * c0, n, torch.float32).copy_(mean[c0:c1].float())
            _flat(b.invstd + 4 * c0, n, torch.float32).copy_(invstd[c0:c1])
            _flat(b.scale + 4 * c0, n, torch.float32).copy_(sc)
            _flat(b.shift + 4 * c0, n, torch.float32).copy_(beta - mean[c0:c1].float() * sc)
            if rm:
                unbiased = var[c0:c1] * (b.count / (b.count - 1.0)) if b.count > 1 else var[c0:c1]
                r_m, r_v = _flat(rm, n, torch.float32), _flat(rv, n, torch.float32)
                r_m.mul_(1 - b.momentum).add_(b.momentum * mean[c0:c1].float())
                r_v.mul_(1 - b.momentum).add_(b.momentum * unbiased.float())
            if nbt:
                _flat(nbt, 1, torch.int64).add_(1)
        if getattr(b, "act_out", None):   # fused normalise: the unit's vtb_bn_act in the same call
            assert not b.split
            self._vtb_bn_act(y, ldy, out.shape[0], g.cout, b.scale, b.shift, b.act_relu, b.act_residual, b.act_ldr,
                             b.act_out, b.act_ld, st)

    @staticmethod
    def _bn_sets(cout, b):
        if b.split:
            return [(0, b.split, b.gamma, b.beta, b.running_mean, b.running_var, b.num_batches_tracked),
                    (b.split, cout, b.gamma2, b.beta2, b.running_mean2, b.running_var2, b.num_batches_tracked2)]
        return [(0, cout, b.gamma, b.beta, b.running_mean, b.running_var, b.num_batches_tracked)]

    def _vtb_conv_dgrad(self, geom, dy, lddy, wd, dx, lddx, accumulate, st):
        g = _obj(geom)
        ho, wo = self._hw(g)
        gy = _view(dy, lddy, g.n * ho * wo, g.cout).float().view(g.n, ho, wo, g.cout).permute(0, 3, 1, 2)
        w = _flat(wd, g.cin * g.k * g.k * g.cout, BF16).float().view(g.cin, g.k, g.k, g.cout).permute(3, 0, 1, 2)
        gx = torch.nn.grad.conv2d_input((g.n, g.cin, g.h, g.w), w, gy, g.stride, g.pad)
        gx = gx.permute(0, 2, 3, 1).reshape(g.n * g.h * g.w, g.cin)
        dst = _view(dx, lddx, gx.shape[0], g.cin)
        dst.copy_(((_r(gx) + dst.float()) if accumulate else gx).to(BF16))

    def _vtb_conv_dgrad_s2(self, geom, dy, lddy, wd, ws, dx, lddx, accumulate, st):
        """Stride-2 dgrad as one GEMM over 2x2 super-pixels: the same function of (dy, wd) as vtb_conv_dgrad."""
        g = _obj(geom)
        assert g.k == 3 and g.stride == 2 and g.pad == 1 and g.h % 2 == 0 and g.w % 2 == 0 and lddx == g.cin and ws
        self._vtb_conv_dgrad(geom, dy, lddy, wd, dx, lddx, accumulate, st)

    def _vtb_conv_dgrad_bn(self, geom, dy, lddy, wd, dx, lddx, accumulate, bn, st):
        """dgrad, then the BatchNorm(+ReLU) backward sums of the producer layer(s) against the COMPLETED dx (vtb.h)."""
        g, b = _obj(geom), _obj(bn)
        assert b.sync is None, "the interpreter models a single rank"
        self._vtb_conv_dgrad(geom, dy, lddy, wd, dx, lddx, accumulate, st)
        pixels = g.n * g.h * g.w
        bounds = [(0, g.cin)] if b.split == 0 else [(0, b.split), (b.split, g.cin)]
        for lay, (c0, c1) in zip(b.layer, bounds):
            c = c1 - c0
            dz, xhat = self._dz_xhat(dx + 2 * c0, lddx, lay.y, lay.ldy, pixels, c, lay.scale, lay.shift, lay.mean,
                                     lay.invstd, lay.relu)
            s0, s1 = dz.double().sum(0), (dz.double() * xhat.double()).sum(0)
            if lay.dgamma:
                _flat(lay.dgamma, c, torch.float32).copy_(s1.float())
            if lay.dbeta:
                _flat(lay.dbeta, c, torch.float32).copy_(s0.float())
            _flat(lay.coef, 2 * c, torch.float32).view(c, 2).copy_(torch.stack([s0 / b.count, s1 / b.count], 1).float())

    def _dw(self, g, dy, lddy, x, ldx):
        ho, wo = self._hw(g)
        gy = _view(dy, lddy, g.n * ho * wo, g.cout).float().view(g.n, ho, wo, g.cout).permute(0, 3, 1, 2)
        xi = _view(x, ldx, g.n * g.h * g.w, g.cin).float().view(g.n, g.h, g.w, g.cin).permute(0, 3, 1, 2)
        return torch.nn.grad.conv2d_weight(xi, (g.cout, g.cin, g.k, g.k), gy, g.stride, g.pad)   # OIHW

    def _vtb_conv_wgrad(self, geom, dy, lddy, x, ldx, ws, dw, cin_real, accumulate, st):
        g = _obj(geom)
        new = self._dw(g, dy, lddy, x, ldx)[:, :cin_real].reshape(-1)
        dst = _flat(dw, new.numel(), torch.float32)
        dst.copy_(dst + new if accumulate else new)

    def _vtb_conv_wgrad_pair(self, geom, dy, lddy, x, ldx, ws, dw_a, dw_b, split, cin_real, accumulate, st):
        g = _obj(geom)
        new = self._dw(g, dy, lddy, x, ldx)[:, :cin_real]
        for ptr, part in ((dw_a, new[:split]), (dw_b, new[split:])):
            dst = _flat(ptr, part.numel(), torch.float32)
            dst.copy_(dst + part.reshape(-1) if accumulate else part.reshape(-1))

    def _vtb_dw_from_col(self, dw_col, cout, c, kk, dw, accumulate, st):
        src = _flat(dw_col, cout * kk * c, torch.float32).view(cout, kk, c)
        dst = _flat(dw, cout * c * kk, torch.float32).view(cout, c, kk)
        dst.copy_(dst + src.permute(0, 2, 1) if accumulate else src.permute(0, 2, 1))

    # ---------------------------------------------------------------- BatchNorm / ReLU / adds
    def _vtb_bn_eval_affine(self, c, gamma, beta, rm, rv, eps, scale, shift, st):
        g, b = _flat(gamma, c, torch.float32), _flat(beta, c, torch.float32)
        sc = g * torch.rsqrt(_flat(rv, c, torch.float32) + eps)
        _flat(scale, c, torch.float32).copy_(sc)
        _flat(shift, c, torch.float32).copy_(b - _flat(rm, c, torch.float32) * sc)

    def _vtb_bn_act(self, y, ldy, pixels, c, scale, shift, relu, res, ldr, out, ldo, st):
        z = _view(y, ldy, pixels, c).float() * _flat(scale, c, torch.float32) + _flat(shift, c, torch.float32)
        if relu:
            z = z.clamp_min(0)
        if res:
            z = _r(z) + _view(res, ldr, pixels, c).float()
        _view(out, ldo, pixels, c).copy_(z.to(BF16))

    def _dz_xhat(self, dout, lddo, y, ldy, pixels, c, scale, shift, mean, invstd, relu):
        yv = _view(y, ldy, pixels, c).float()
        dz = _view(dout, lddo, pixels, c).float()
        if relu:
            z = yv * _flat(scale, c, torch.float32) + _flat(shift, c, torch.float32)
            dz = torch.where(z > 0, dz, torch.zeros_like(dz))
        xhat = (yv - _flat(mean, c, torch.float32)) * _flat(invstd, c, torch.float32)
        return dz, xhat

    def _vtb_bn_bwd_fused(self, dout, lddo, y, ldy, pixels, c, scale, shift, mean, invstd, relu, count, partial, dgamma,
                          dbeta, accumulate, sync, dy, lddy, peers, st):
        assert peers is None
        dz, xhat = self._dz_xhat(dout, lddo, y, ldy, pixels, c, scale, shift, mean, invstd, relu)
        s0, s1 = dz.double().sum(0), (dz.double() * xhat.double()).sum(0)
        for ptr, val in ((dgamma, s1), (dbeta, s0)):
            if ptr:
                d = _flat(ptr, c, torch.float32)
                d.copy_(d + val.float() if accumulate else val.float())
        sc = _flat(scale, c, torch.float32)
        out = sc * (dz - (s0 / count).float() - xhat * (s1 / count).float())
        _view(dy, lddy, pixels, c).copy_(out.to(BF16))

    def _vtb_bn_bwd_reduce(self, dout, lddo, y, ldy, pixels, c, scale, shift, mean, invstd, relu, partial, st):
        dz, xhat = self._dz_xhat(dout, lddo, y, ldy, pixels, c, scale, shift, mean, invstd, relu)
        rows = self.real.vtb_bn_bwd_rows(pixels, c)
        part = _flat(partial, rows * c * 2, torch.float32).view(rows, c, 2)
        part.zero_()
        part[0, :, 0] = dz.double().sum(0).float()
        part[0, :, 1] = (dz.double() * xhat.double()).sum(0).float()

    def _vtb_bn_bwd_finalize(self, partial, rows, sums, local_sums, count, c, dgamma, dbeta, accumulate, coef, sums_out, st):
        if sums:
            s = _flat(sums, 2 * c, torch.float64).view(c, 2)
        else:
            s = _flat(partial, rows * c * 2, torch.float32).view(rows, c, 2).double().sum(0)
        if sums_out:
            _flat(sums_out, 2 * c, torch.float64).copy_(s.reshape(-1))
            return
        loc = _flat(local_sums, 2 * c, torch.float64).view(c, 2) if local_sums else s
        for ptr, val in ((dgamma, loc[:, 1]), (dbeta, loc[:, 0])):
            if ptr:
                d = _flat(ptr, c, torch.float32)
                d.copy_(d + val.float() if accumulate else val.float())
        _flat(coef, 2 * c, torch.float32).copy_((s / count).float().reshape(-1))

    def _vtb_bn_bwd_apply(self, dout, lddo, y, ldy, pixels, c, scale, shift, mean, invstd, relu, coef, dy, lddy, st):
        dz, xhat = self._dz_xhat(dout, lddo, y, ldy, pixels, c, scale, shift, mean, invstd, relu)
        k = _flat(coef, 2 * c, torch.float32).view(c, 2)
        out = _flat(scale, c, torch.float32) * (dz - k[:, 0] - xhat * k[:, 1])
        _view(dy, lddy, pixels, c).copy_(out.to(BF16))

    def _vtb_resize2_add(self, a, lda, b, ldb, n, h, w, c, hb, wb, up, out, ldo, st):
        bb = _view(b, ldb, n * hb * wb, c).float().view(n, hb, wb, c)
        r = bb.repeat_interleave(2, 1).repeat_interleave(2, 2) if up else bb[:, ::2, ::2][:, :h, :w]
        r = r.reshape(n * h * w, c)
        if a:
            r = r + _view(a, lda, n * h * w, c).float()
        _view(out, ldo, n * h * w, c).copy_(r.to(BF16))

    def _vtb_resize2_add_bwd(self, gout, ldg, n, h, w, c, gb, ldgb, hb, wb, up, accumulate, st):
        g = _view(gout, ldg, n * h * w, c).float().view(n, h, w, c)
        if up:
            new = g.view(n, hb, 2, wb, 2, c).sum(dim=(2, 4))
        else:
            new = torch.zeros(n, hb, wb, c)
            new[:, : 2 * h : 2, : 2 * w : 2] = g
        new = new.reshape(n * hb * wb, c)
        dst = _view(gb, ldgb, n * hb * wb, c)
        dst.copy_(((_r(new) + dst.float()) if accumulate else new).to(BF16))

    def _vtb_mix_images(self, x, out, n, c, h, w, prm, st):
        p = _flat(prm, 6, torch.float32)
        xi = _flat(x, n * c * h * w, torch.float32).view(n, c, h, w)
        prev = xi.roll(1, 0)
        mode, lam = int(p[0]), p[1]
        if mode == 1:
            o = xi * lam + prev * (1.0 - lam)
        elif mode == 2:
            o = xi.clone()
            x1, y1, x2, y2 = (int(v) for v in p[2:6])
            o[:, :, y1:y2, x1:x2] = prev[:, :, y1:y2, x1:x2]
        else:
            o = xi.clone()
        _flat(out, n * c * h * w, torch.float32).copy_(o.reshape(-1))

    def _vtb_grad_add(self, dst, ldd, src, lds, pixels, c, accumulate, st):
        d, s = _view(dst, ldd, pixels, c), _view(src, lds, pixels, c)
        d.copy_((d.float() + s.float()).to(BF16) if accumulate else s)

    # ---------------------------------------------------------------- VoVNet ops
    def _vtb_maxpool3s2_fwd(self, x, ldx, n, h, w, c, out, ldo, idx, st):
        xi = _view(x, ldx, n * h * w, c).float().view(n, h, w, c).permute(0, 3, 1, 2)
        y, ind = F.max_pool2d(xi, 3, 2, 1, return_indices=True)
        ho, wo = y.shape[2:]
        _view(out, ldo, n * ho * wo, c).copy_(y.permute(0, 2, 3, 1).reshape(-1, c).to(BF16))
        if idx:
            ih, iw = ind // w, ind % w
            oh = torch.arange(ho).view(1, 1, ho, 1)
            ow = torch.arange(wo).view(1, 1, 1, wo)
            code = (ih - (oh * 2 - 1)) * 3 + (iw - (ow * 2 - 1))            # window position r*3+s of the maximum
            _flat(idx, n * ho * wo * c, torch.uint8).copy_(code.permute(0, 2, 3, 1).reshape(-1).to(torch.uint8))

    def _vtb_maxpool3s2_bwd(self, x, ldx, n, h, w, c, dout, lddo, dx, lddx, accumulate, idx, st):
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        code = _flat(idx, n * ho * wo * c, torch.uint8).view(n, ho, wo, c).long()
        g = _view(dout, lddo, n * ho * wo, c).float().view(n, ho, wo, c)
        oh = torch.arange(ho).view(1, ho, 1, 1)
        ow = torch.arange(wo).view(1, 1, wo, 1)
        ih, iw = oh * 2 - 1 + code // 3, ow * 2 - 1 + code % 3
        acc = torch.zeros(n, h * w, c)
        acc.scatter_add_(1, (ih * w + iw).view(n, ho * wo, c), g.reshape(n, ho * wo, c))
        dst = _view(dx, lddx, n * h * w, c)
        new = acc.view(n * h * w, c)
        dst.copy_(((_r(new) + dst.float()) if accumulate else new).to(BF16))

    def _vtb_ese_fwd(self, x, ldx, n, hw, c, weight, bias, res, ldr, out, ldo, pool, z, gate, st):
        xv = _view(x, ldx, n * hw, c).float().view(n, hw, c)
        p = xv.sum(1) / hw
        W, b = _flat(weight, c * c, torch.float32).view(c, c), _flat(bias, c, torch.float32)
        zz = _r(_r(p) @ _r(W).t() + _r(b))
        gt = _r((zz / 6 + 0.5).clamp(0, 1))
        for ptr, val in ((pool, p), (z, zz), (gate, gt)):
            _flat(ptr, n * c, torch.float32).copy_(val.reshape(-1))
        o = xv * gt[:, None, :]
        if res:
            o = _r(o) + _view(res, ldr, n * hw, c).float().view(n, hw, c)
        _view(out, ldo, n * hw, c).copy_(o.reshape(n * hw, c).to(BF16))

    def _vtb_ese_bwd(self, x, ldx, n, hw, c, weight, pool, z, gate, dout, lddo, dx, lddx, acc_dx, dweight, dbias, acc_dw,
                     scratch, st):
        xv = _view(x, ldx, n * hw, c).float().view(n, hw, c)
        g = _view(dout, lddo, n * hw, c).float().view(n, hw, c)
        W = _flat(weight, c * c, torch.float32).view(c, c)
        p, zz, gt = (_flat(t, n * c, torch.float32).view(n, c) for t in (pool, z, gate))
        dgate = (g * xv).sum(1)
        t = zz / 6 + 0.5
        dz = torch.where((t > 0) & (t < 1), dgate / 6, torch.zeros_like(dgate))
        dpool = (dz @ _r(W)) / hw
        for ptr, val in ((dweight, dz.t() @ _r(p)), (dbias, dz.sum(0))):
            d = _flat(ptr, val.numel(), torch.float32)
            d.copy_(d + val.reshape(-1) if acc_dw else val.reshape(-1))
        new = (g * gt[:, None, :] + dpool[:, None, :]).reshape(n * hw, c)
        dst = _view(dx, lddx, n * hw, c)
        dst.copy_(((_r(new) + dst.float()) if acc_dx else new).to(BF16))

    # ---------------------------------------------------------------- fp32 parity mode (vtb_f32_*)
    def _vtb_f32_nchw_to_nhwc(self, x, n, c, h, w, out, cpad, st):
        src = _flat(x, n * c * h * w, torch.float32).view(n, c, h, w)
        dst = _view(out, cpad, n * h * w, cpad, torch.float32)
        dst.zero_()
        dst[:, :c] = src.permute(0, 2, 3, 1).reshape(-1, c)

    @staticmethod
    def _w32(g, w, cin_real):
        """OIHW fp32 master, zero-extended to the padded channel count of the operand view."""
        src = _flat(w, g.cout * cin_real * g.k * g.k, torch.float32).view(g.cout, cin_real, g.k, g.k)
        full = torch.zeros(g.cout, g.cin, g.k, g.k)
        full[:, :cin_real] = src
        return full

    def _vtb_f32_conv_fprop(self, geom, x, ldx, w, cin_real, y, ldy, st):
        g = _obj(geom)
        ho, wo = self._hw(g)
        xi = _view(x, ldx, g.n * g.h * g.w, g.cin, torch.float32).view(g.n, g.h, g.w, g.cin).permute(0, 3, 1, 2)
        out = F.conv2d(xi, self._w32(g, w, cin_real), None, g.stride, g.pad)
        _view(y, ldy, g.n * ho * wo, g.cout, torch.float32).copy_(out.permute(0, 2, 3, 1).reshape(-1, g.cout))

    def _vtb_f32_conv_dgrad(self, geom, dy, lddy, w, cin_real, dx, lddx, accumulate, st):
        g = _obj(geom)
        ho, wo = self._hw(g)
        gy = _view(dy, lddy, g.n * ho * wo, g.cout, torch.float32).view(g.n, ho, wo, g.cout).permute(0, 3, 1, 2)
        gx = torch.nn.grad.conv2d_input((g.n, g.cin, g.h, g.w), self._w32(g, w, cin_real), gy, g.stride, g.pad)
        gx = gx.permute(0, 2, 3, 1).reshape(g.n * g.h * g.w, g.cin)
        dst = _view(dx, lddx, gx.shape[0], g.cin, torch.float32)
        dst.copy_(dst + gx if accumulate else gx)

    def _vtb_f32_conv_wgrad(self, geom, dy, lddy, x, ldx, ws, dw, cin_real, accumulate, st):
        g = _obj(geom)
        ho, wo = self._hw(g)
        gy = _view(dy, lddy, g.n * ho * wo, g.cout, torch.float32).view(g.n, ho, wo, g.cout).permute(0, 3, 1, 2)
        xi = _view(x, ldx, g.n * g.h * g.w, g.cin, torch.float32).view(g.n, g.h, g.w, g.cin).permute(0, 3, 1, 2)
        new = torch.nn.grad.conv2d_weight(xi, (g.cout, g.cin, g.k, g.k), gy, g.stride, g.pad)[:, :cin_real].reshape(-1)
        dst = _flat(dw, new.numel(), torch.float32)
        dst.copy_(dst + new if accumulate else new)

    def _vtb_f32_bn_stats(self, y, ldy, pixels, c, partial, sums, st):
        yv = _view(y, ldy, pixels, c, torch.float32).double()
        _flat(sums, 2 * c, torch.float64).copy_(torch.stack([yv.sum(0), (yv * yv).sum(0)], 1).reshape(-1))

    def _vtb_bn_finalize(self, partial, rows, sums, count, c, gamma, beta, eps, momentum, rm, rv, nbt, mean, invstd, scale,
                         shift, st):
        if sums:
            s = _flat(sums, 2 * c, torch.float64).view(c, 2)
        else:
            s = _flat(partial, rows * c * 2, torch.float32).view(rows, c, 2).double().sum(0)
        mu = s[:, 0] / count
        var = (s[:, 1] / count - mu * mu).clamp_min(0)
        istd = (1.0 / torch.sqrt(var + eps)).float()
        sc = _flat(gamma, c, torch.float32) * istd
        _flat(mean, c, torch.float32).copy_(mu.float())
        _flat(invstd, c, torch.float32).copy_(istd)
        _flat(scale, c, torch.float32).copy_(sc)
        _flat(shift, c, torch.float32).copy_(_flat(beta, c, torch.float32) - mu.float() * sc)
        if rm:
            unbiased = var * (count / (count - 1.0)) if count > 1 else var
            _flat(rm, c, torch.float32).mul_(1 - momentum).add_(momentum * mu.float())
            _flat(rv, c, torch.float32).mul_(1 - momentum).add_(momentum * unbiased.float())
        if nbt:
            _flat(nbt, 1, torch.int64).add_(1)

    def _vtb_bn_stats_reduce(self, partial, rows, c, sums, st):
        _flat(sums, 2 * c, torch.float64).copy_(
            _flat(partial, rows * c * 2, torch.float32).view(rows, c, 2).double().sum(0).reshape(-1))

    @staticmethod
    def _bn32(yv, mean, invstd, gamma, beta, c):
        f = lambda p: _flat(p, c, torch.float32)
        return (yv - f(mean)) * f(invstd) * f(gamma) + f(beta)

    def _vtb_f32_bn_act(self, y, ldy, pixels, c, mean, invstd, gamma, beta, relu, res, ldr, out, ldo, st):
        z = self._bn32(_view(y, ldy, pixels, c, torch.float32), mean, invstd, gamma, beta, c)
        if relu:
            z = z.clamp_min(0)
        if res:
            z = z + _view(res, ldr, pixels, c, torch.float32)
        _view(out, ldo, pixels, c, torch.float32).copy_(z)

    def _dz32(self, dout, lddo, y, ldy, pixels, c, mean, invstd, gamma, beta, relu):
        yv = _view(y, ldy, pixels, c, torch.float32)
        dz = _view(dout, lddo, pixels, c, torch.float32).clone()
        if relu:
            dz = torch.where(self._bn32(yv, mean, invstd, gamma, beta, c) > 0, dz, torch.zeros_like(dz))
        return dz, (yv - _flat(mean, c, torch.float32)) * _flat(invstd, c, torch.float32)

    def _vtb_f32_bn_bwd_reduce(self, dout, lddo, y, ldy, pixels, c, mean, invstd, gamma, beta, relu, partial, sums, st):
        dz, xhat = self._dz32(dout, lddo, y, ldy, pixels, c, mean, invstd, gamma, beta, relu)
        _flat(sums, 2 * c, torch.float64).copy_(
            torch.stack([dz.double().sum(0), (dz.double() * xhat.double()).sum(0)], 1).reshape(-1))

    def _vtb_f32_bn_bwd_apply(self, dout, lddo, y, ldy, pixels, c, mean, invstd, gamma, beta, relu, coef, dy, lddy, st):
        dz, xhat = self._dz32(dout, lddo, y, ldy, pixels, c, mean, invstd, gamma, beta, relu)
        k = _flat(coef, 2 * c, torch.float32).view(c, 2)
        out = _flat(gamma, c, torch.float32) * _flat(invstd, c, torch.float32) * (dz - k[:, 0] - xhat * k[:, 1])
        _view(dy, lddy, pixels, c, torch.float32).copy_(out)

    def _vtb_f32_grad_add(self, dst, ldd, src, lds, pixels, c, accumulate, st):
        d, s = _view(dst, ldd, pixels, c, torch.float32), _view(src, lds, pixels, c, torch.float32)
        d.copy_(d + s if accumulate else s)

    def _vtb_f32_maxpool3s2_fwd(self, x, ldx, n, h, w, c, out, ldo, idx, st):
        xi = _view(x, ldx, n * h * w, c, torch.float32).view(n, h, w, c).permute(0, 3, 1, 2)
        y, ind = F.max_pool2d(xi, 3, 2, 1, return_indices=True)
        ho, wo = y.shape[2:]
        _view(out, ldo, n * ho * wo, c, torch.float32).copy_(y.permute(0, 2, 3, 1).reshape(-1, c))
        if idx:
            oh, ow = torch.arange(ho).view(1, 1, ho, 1), torch.arange(wo).view(1, 1, 1, wo)
            code = (ind // w - (oh * 2 - 1)) * 3 + (ind % w - (ow * 2 - 1))
            _flat(idx, n * ho * wo * c, torch.uint8).copy_(code.permute(0, 2, 3, 1).reshape(-1).to(torch.uint8))

    def _vtb_f32_maxpool3s2_bwd(self, x, ldx, n, h, w, c, dout, lddo, dx, lddx, accumulate, idx, st):
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        code = _flat(idx, n * ho * wo * c, torch.uint8).view(n, ho, wo, c).long()
        g = _view(dout, lddo, n * ho * wo, c, torch.float32).view(n, ho, wo, c)
        oh, ow = torch.arange(ho).view(1, ho, 1, 1), torch.arange(wo).view(1, 1, wo, 1)
        pos = (oh * 2 - 1 + code // 3) * w + (ow * 2 - 1 + code % 3)
        acc = torch.zeros(n, h * w, c)
        acc.scatter_add_(1, pos.view(n, ho * wo, c), g.reshape(n, ho * wo, c))
        dst = _view(dx, lddx, n * h * w, c, torch.float32)
        dst.copy_(dst + acc.view(-1, c) if accumulate else acc.view(-1, c))

    def _vtb_f32_ese_fwd(self, x, ldx, n, hw, c, weight, bias, res, ldr, out, ldo, pool, z, gate, st):
        xv = _view(x, ldx, n * hw, c, torch.float32).view(n, hw, c)
        p = xv.sum(1) / hw
        zz = p @ _flat(weight, c * c, torch.float32).view(c, c).t() + _flat(bias, c, torch.float32)
        gt = (zz / 6 + 0.5).clamp(0, 1)
        for ptr, val in ((pool, p), (z, zz), (gate, gt)):
            _flat(ptr, n * c, torch.float32).copy_(val.reshape(-1))
        o = xv * gt[:, None, :]
        if res:
            o = o + _view(res, ldr, n * hw, c, torch.float32).view(n, hw, c)
        _view(out, ldo, n * hw, c, torch.float32).copy_(o.reshape(n * hw, c))

    def _vtb_f32_ese_bwd(self, x, ldx, n, hw, c, weight, pool, z, gate, dout, lddo, dx, lddx, acc_dx, dweight, dbias,
                         acc_dw, scratch, st):
        xv = _view(x, ldx, n * hw, c, torch.float32).view(n, hw, c)
        g = _view(dout, lddo, n * hw, c, torch.float32).view(n, hw, c)
        W = _flat(weight, c * c, torch.float32).view(c, c)
        p, zz, gt = (_flat(t, n * c, torch.float32).view(n, c) for t in (pool, z, gate))
        dgate = (g * xv).sum(1)
        dz = torch.where((zz > -3) & (zz < 3), dgate / 6, torch.zeros_like(dgate))
        dpool = (dz @ W) / hw
        for ptr, val in ((dweight, dz.t() @ p), (dbias, dz.sum(0))):
            d = _flat(ptr, val.numel(), torch.float32)
            d.copy_(d + val.reshape(-1) if acc_dw else val.reshape(-1))
        new = (g * gt[:, None, :] + dpool[:, None, :]).reshape(n * hw, c)
        dst = _view(dx, lddx, n * hw, c, torch.float32)
        dst.copy_(dst + new if acc_dx else new)
