"""ctypes binding of libvtb_b200.so (the C ABI declared in include/vtb.h).

The library is built in-tree by ``make -C vision_toolbox_b200/csrc`` (or ``__graft_entry__.build()``).
There is deliberately NO fallback: a CUDA tensor reaching the backbone without this library raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libvtb_b200.so"
CSRC = _PKG / "csrc"


class VtbConv(C.Structure):
    """struct VtbConv of include/vtb.h (geometry of nn.Conv2d at reference components.py:26-35)."""

    _fields_ = [(k, C.c_int) for k in ("n", "h", "w", "cin", "cout", "k", "stride", "pad")]


class VtbBnTrain(C.Structure):
    """struct VtbBnTrain of include/vtb.h (BatchNorm2d training tensors for the fused conv + finalize call)."""

    _fields_ = [("count", C.c_double), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
                ("momentum", C.c_float), ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
                ("num_batches_tracked", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("tickets", C.c_void_p), ("sync", C.c_void_p),
                ("split", C.c_int), ("gamma2", C.c_void_p), ("beta2", C.c_void_p), ("running_mean2", C.c_void_p),
                ("running_var2", C.c_void_p), ("num_batches_tracked2", C.c_void_p),
                ("act_out", C.c_void_p), ("act_ld", C.c_int), ("act_relu", C.c_int), ("act_residual", C.c_void_p),
                ("act_ldr", C.c_int)]


class VtbBnBwdLayer(C.Structure):
    """struct VtbBnBwdLayer of include/vtb.h (one producer layer of a dgrad that carries BatchNorm-backward statistics)."""

    _fields_ = [("y", C.c_void_p), ("ldy", C.c_int), ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p),
                ("invstd", C.c_void_p), ("relu", C.c_int), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
                ("coef", C.c_void_p)]


class VtbDgradBn(C.Structure):
    """struct VtbDgradBn of include/vtb.h."""

    _fields_ = [("split", C.c_int), ("layer", VtbBnBwdLayer * 2), ("count", C.c_double), ("partial", C.c_void_p),
                ("tickets", C.c_void_p), ("sync", C.c_void_p)]


class VtbPackJob(C.Structure):
    """struct VtbPackJob of include/vtb.h (one convolution's weight re-pack inside the batched launch)."""

    _fields_ = [("w", C.c_void_p), ("wf", C.c_void_p), ("wd", C.c_void_p), ("cout", C.c_int), ("cin_real", C.c_int),
                ("cin", C.c_int), ("kk", C.c_int), ("wd_ld", C.c_int), ("wd_co_off", C.c_int), ("wf_ld", C.c_int),
                ("first_block", C.c_longlong), ("g", C.c_void_p), ("m", C.c_void_p), ("weight_decay", C.c_float)]


class VtbSgdJob(C.Structure):
    """struct VtbSgdJob of include/vtb.h (one plain tensor of the fused SGD-momentum step)."""

    _fields_ = [("w", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("n", C.c_longlong),
                ("weight_decay", C.c_float), ("first_block", C.c_longlong)]


class VtbSyncBn(C.Structure):
    """struct VtbSyncBn of include/vtb.h (peer-mapped SyncBN exchange buffers)."""

    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("peer_buffers", C.c_void_p * 8)]


_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_f = C.c_float
_d = C.c_double
_cp = C.POINTER(VtbConv)

# name -> (restype, argtypes); mirrors include/vtb.h one to one (tests/test_abi.py checks the symbol list)
SIGNATURES = {
    "vtb_last_error": (C.c_char_p, []),
    "vtb_version": (_i, []),
    "vtb_num_sms": (_i, []),
    "vtb_launch_count": (_ll, []),
    "vtb_conv_out_hw": (_i, [_cp, C.POINTER(_i), C.POINTER(_i)]),
    "vtb_conv_stats_rows": (_i, [_cp]),
    "vtb_conv_wgrad_workspace_bytes": (C.c_size_t, [_cp]),
    "vtb_conv_tiling_info": (_i, [_cp, _i, C.POINTER(_i)]),
    "vtb_pack_weight": (_i, [_cp, _p, _i, _p, _p, _p]),
    "vtb_pack_job_blocks": (_ll, [_i, _i, _i]),
    "vtb_pack_weights": (_i, [_p, _i, _ll, _p]),
    "vtb_sgd_pack_weights": (_i, [_p, _i, _ll, _p, _p]),
    "vtb_sgd_job_blocks": (_ll, [_ll]),
    "vtb_sgd_step": (_i, [_p, _i, _ll, _p, _p]),
    "vtb_conv_fprop": (_i, [_cp, _p, _i, _p, _p, _i, _p, _p, _p, _i, _p, _i, _p]),
    "vtb_conv_fprop_bn": (_i, [_cp, _p, _i, _p, _p, _i, _p, C.POINTER(VtbBnTrain), _p]),
    "vtb_conv_dgrad": (_i, [_cp, _p, _i, _p, _p, _i, _i, _p]),
    "vtb_conv_dgrad_s2_workspace_bytes": (C.c_size_t, [_cp]),
    "vtb_conv_dgrad_s2": (_i, [_cp, _p, _i, _p, _p, _p, _i, _i, _p]),
    "vtb_conv_dgrad_stats_rows": (_i, [_cp]),
    "vtb_conv_dgrad_panel_w": (_i, [_cp]),
    "vtb_conv_dgrad_bn": (_i, [_cp, _p, _i, _p, _p, _i, _i, C.POINTER(VtbDgradBn), _p]),
    "vtb_conv_wgrad": (_i, [_cp, _p, _i, _p, _i, _p, _p, _i, _i, _p]),
    "vtb_conv_wgrad_pair": (_i, [_cp, _p, _i, _p, _i, _p, _p, _p, _i, _i, _i, _p]),
    "vtb_bn_stats_reduce": (_i, [_p, _i, _i, _p, _p]),
    "vtb_bn_finalize": (_i, [_p, _i, _p, _d, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vtb_bn_eval_affine": (_i, [_i, _p, _p, _p, _p, _f, _p, _p, _p]),
    "vtb_bn_sync_buffer_bytes": (C.c_size_t, []),
    "vtb_bn_sync_finalize": (_i, [_p, _i, _i, C.POINTER(VtbSyncBn), _d, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vtb_bn_sync_bwd_finalize": (_i, [_p, _i, _i, C.POINTER(VtbSyncBn), _d, _p, _p, _i, _p, _p, _p]),
    "vtb_bn_act": (_i, [_p, _i, _ll, _i, _p, _p, _i, _p, _i, _p, _i, _p]),
    "vtb_bn_bwd_rows": (_i, [_ll, _i]),
    "vtb_bn_bwd_reduce": (_i, [_p, _i, _p, _i, _ll, _i, _p, _p, _p, _p, _i, _p, _p]),
    "vtb_bn_bwd_finalize": (_i, [_p, _i, _p, _p, _d, _i, _p, _p, _i, _p, _p, _p]),
    "vtb_bn_bwd_apply": (_i, [_p, _i, _p, _i, _ll, _i, _p, _p, _p, _p, _i, _p, _p, _i, _p]),
    "vtb_bn_bwd_fused_rows": (_i, [_ll, _i]),
    "vtb_bn_bwd_fused": (_i, [_p, _i, _p, _i, _ll, _i, _p, _p, _p, _p, _i, _d, _p, _p, _p, _i, _p, _p, _i, _p, _p]),
    "vtb_grad_add": (_i, [_p, _i, _p, _i, _ll, _i, _i, _p]),
    "vtb_nchw_to_nhwc": (_i, [_p, _i, _i, _i, _i, _p, _i, _p]),
    "vtb_im2col_input": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "vtb_dw_from_col": (_i, [_p, _i, _i, _i, _p, _i, _p]),
    "vtb_mix_images": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "vtb_resize2_add": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "vtb_resize2_add_bwd": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p]),
    "vtb_maxpool3s2_fwd": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _p]),
    "vtb_maxpool3s2_bwd": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _i, _i, _p, _p]),
    "vtb_ese_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p]),
    "vtb_ese_bwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _i, _i, _p, _p, _i, _p, _p]),
    "vtb_head_ce_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p, _f, _p, _p, _p, _p, _p, _p]),
    "vtb_head_ce_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p]),
    # fp32 parity mode (csrc/parity_f32.cu)
    "vtb_f32_nchw_to_nhwc": (_i, [_p, _i, _i, _i, _i, _p, _i, _p]),
    "vtb_f32_conv_fprop": (_i, [_cp, _p, _i, _p, _i, _p, _i, _p]),
    "vtb_f32_conv_dgrad": (_i, [_cp, _p, _i, _p, _i, _p, _i, _i, _p]),
    "vtb_f32_conv_wgrad_workspace_bytes": (C.c_size_t, [_cp]),
    "vtb_f32_conv_wgrad": (_i, [_cp, _p, _i, _p, _i, _p, _p, _i, _i, _p]),
    "vtb_f32_bn_rows": (_i, [_ll, _i]),
    "vtb_f32_bn_stats": (_i, [_p, _i, _ll, _i, _p, _p, _p]),
    "vtb_f32_bn_act": (_i, [_p, _i, _ll, _i, _p, _p, _p, _p, _i, _p, _i, _p, _i, _p]),
    "vtb_f32_bn_bwd_reduce": (_i, [_p, _i, _p, _i, _ll, _i, _p, _p, _p, _p, _i, _p, _p, _p]),
    "vtb_f32_bn_bwd_apply": (_i, [_p, _i, _p, _i, _ll, _i, _p, _p, _p, _p, _i, _p, _p, _i, _p]),
    "vtb_f32_grad_add": (_i, [_p, _i, _p, _i, _ll, _i, _i, _p]),
    "vtb_f32_maxpool3s2_fwd": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _p]),
    "vtb_f32_maxpool3s2_bwd": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _i, _i, _p, _p]),
    "vtb_f32_ese_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p]),
    "vtb_f32_ese_bwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _i, _i, _p, _p, _i, _p, _p]),
}

_lib = None


class VtbError(RuntimeError):
    pass


def build(verbose: bool = False) -> Path:
    """Compile libvtb_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", str(CSRC), "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise VtbError("building libvtb_b200.so failed")
    return LIB_PATH


def lib():
    """Load (once) and return the ctypes handle; raises VtbError when the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("VTB_AUTOBUILD", "1") == "1" and (CSRC / "Makefile").exists():
            build()
        if not LIB_PATH.exists():
            raise VtbError(
                f"{LIB_PATH} not found: build it with `make -C {CSRC}`; vision_toolbox_b200 has no CPU/torch "
                "fallback for CUDA tensors"
            )
    h = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(h, name)
        fn.restype = res
        fn.argtypes = args
    _lib = h
    return h


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().vtb_last_error().decode(errors="replace")
        raise VtbError(f"{what or 'vtb call'} failed ({rc}): {msg}")


def launch_count() -> int:
    return int(lib().vtb_launch_count())
