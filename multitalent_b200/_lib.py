"""ctypes binding of libmtb200.so (C ABI declared in include/mtb200.h).

The product path has no CPU fallback: if the shared library is missing or fails to load, `lib()` raises -- loudly.
"""
import ctypes as C
import os

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libmtb200.so")

F32, BF16, F16 = 0, 1, 2
MAX_TAPS, MAX_GROUPS = 32, 8

_TORCH2ENUM = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}
_ENUM2TORCH = {v: k for k, v in _TORCH2ENUM.items()}


def dtype_enum(dt):
    return _TORCH2ENUM[dt]


def torch_dtype(e):
    return _ENUM2TORCH[e]


class ConvParams(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("out", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("xform", C.c_void_p),
        ("stats", C.c_void_p),
        ("dtype", C.c_int32), ("wdtype", C.c_int32), ("B", C.c_int32),
        ("Di", C.c_int32), ("Hi", C.c_int32), ("Wi", C.c_int32), ("in_ldc", C.c_int32), ("in_coff", C.c_int32),
        ("Cin", C.c_int32),
        ("Dof", C.c_int32), ("Hof", C.c_int32), ("Wof", C.c_int32), ("out_ldc", C.c_int32), ("out_coff", C.c_int32),
        ("Cout", C.c_int32),
        ("Do", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("is_", C.c_int32 * 3), ("os_", C.c_int32 * 3),
        ("ngroups", C.c_int32),
        ("group_tap_begin", C.c_int32 * (MAX_GROUPS + 1)),
        ("group_ooff", (C.c_int32 * 3) * MAX_GROUPS),
        ("ntaps", C.c_int32),
        ("tap_off", (C.c_int32 * 3) * MAX_TAPS),
        ("tap_widx", C.c_int32 * MAX_TAPS),
        ("accumulate", C.c_int32),
        ("red_y", C.c_void_p), ("red_xform", C.c_void_p), ("red_meanrstd", C.c_void_p), ("red", C.c_void_p),
        ("red_ldc", C.c_int32), ("red_coff", C.c_int32),
        ("impl", C.c_int32),
        ("in_split", C.c_int32), ("out_split", C.c_int32),
    ]


class WgradParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p), ("xform", C.c_void_p),
        ("dtype", C.c_int32), ("B", C.c_int32),
        ("Di", C.c_int32), ("Hi", C.c_int32), ("Wi", C.c_int32), ("in_ldc", C.c_int32), ("in_coff", C.c_int32),
        ("Cin", C.c_int32),
        ("Dof", C.c_int32), ("Hof", C.c_int32), ("Wof", C.c_int32), ("out_ldc", C.c_int32), ("out_coff", C.c_int32),
        ("Cout", C.c_int32),
        ("Do", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("is_", C.c_int32 * 3), ("os_", C.c_int32 * 3),
        ("ngroups", C.c_int32),
        ("group_tap_begin", C.c_int32 * (MAX_GROUPS + 1)),
        ("group_ooff", (C.c_int32 * 3) * MAX_GROUPS),
        ("ntaps", C.c_int32),
        ("tap_off", (C.c_int32 * 3) * MAX_TAPS),
        ("tap_widx", C.c_int32 * MAX_TAPS),
        ("impl", C.c_int32),
        ("in_split", C.c_int32),
    ]


class PackDesc(C.Structure):
    _fields_ = [
        ("w", C.c_void_p), ("packed", C.c_void_p), ("packed_swap", C.c_void_p),
        ("Cout", C.c_int32), ("Cin", C.c_int32), ("ntap", C.c_int32), ("transposed", C.c_int32),
        ("Cout_p", C.c_int32), ("Cin_p", C.c_int32), ("split", C.c_int32), ("split_p", C.c_int32),
        ("blk_begin", C.c_int32), ("reserved", C.c_int32),
    ]


class UnpackDesc(C.Structure):
    _fields_ = [
        ("dw", C.c_void_p), ("grad", C.c_void_p),
        ("Cout", C.c_int32), ("Cin", C.c_int32), ("ntap", C.c_int32), ("transposed", C.c_int32),
        ("Cout_p", C.c_int32), ("Cin_p", C.c_int32), ("split", C.c_int32), ("split_p", C.c_int32),
        ("blk_begin", C.c_int32), ("reserved", C.c_int32),
    ]


MAX_HEAD_BATCH = 16


class HeadBwdParams(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("target", C.c_void_p), ("coef", C.c_void_p), ("gscale", C.c_void_p),
        ("pos_mask", C.c_void_p), ("x", C.c_void_p), ("w_swap", C.c_void_p), ("w_fwd", C.c_void_p), ("dx", C.c_void_p),
        ("dw", C.c_void_p),
        ("nvox", C.c_int64),
        ("dtype", C.c_int32), ("B", C.c_int32), ("z_ldc", C.c_int32), ("C8", C.c_int32), ("n_labels", C.c_int32),
        ("x_ldc", C.c_int32), ("x_coff", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32), ("dx_ldc", C.c_int32),
        ("dx_coff", C.c_int32), ("accumulate", C.c_int32),
        ("win_c0", C.c_int32 * MAX_HEAD_BATCH),
    ]


class HeadFwdParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w_fwd", C.c_void_p), ("target", C.c_void_p), ("valid_mask", C.c_void_p),
        ("pos_mask", C.c_void_p), ("stats", C.c_void_p), ("hard", C.c_void_p),
        ("nvox", C.c_int64),
        ("dtype", C.c_int32), ("B", C.c_int32), ("C8", C.c_int32), ("n_labels", C.c_int32), ("x_ldc", C.c_int32),
        ("x_coff", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
        ("win_c0", C.c_int32 * MAX_HEAD_BATCH),
    ]


class HeadAggParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w_fwd", C.c_void_p), ("bias", C.c_void_p), ("gauss", C.c_void_p), ("acc", C.c_void_p),
        ("nb", C.c_void_p), ("weight", C.c_float),
        ("dtype", C.c_int32), ("x_ldc", C.c_int32), ("x_coff", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
        ("C", C.c_int32), ("pd", C.c_int32), ("ph", C.c_int32), ("pw", C.c_int32), ("flip", C.c_int32),
        ("nonlin", C.c_int32), ("X", C.c_int32), ("Y", C.c_int32), ("Z", C.c_int32), ("x0", C.c_int32),
        ("y0", C.c_int32), ("z0", C.c_int32),
    ]


UNPACK_CHUNK = 4096
_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p

# name -> argtypes; every symbol include/mtb200.h declares (tests/test_abi.py checks the two lists agree)
SIGNATURES = {
    "mtb200_version": [],
    "mtb200_last_error": [],
    "mtb200_has_tcgen05": [],
    "mtb200_last_kernel": [],
    "mtb200_conv_taps": [C.POINTER(ConvParams), _vp],
    "mtb200_wgrad_taps": [C.POINTER(WgradParams), _vp],
    "mtb200_conv_c1_fwd": [_vp, _i64, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "mtb200_conv_c1_wgrad": [_vp, _i64, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "mtb200_colsum": [_vp, _i32, _i64, _i32, _i32, _i32, _vp, _vp],
    "mtb200_in_finalize": [_vp, _vp, _vp, _i32, _i32, _i64, _f32, _f32, _vp, _vp, _vp],
    "mtb200_in_stats": [_vp, _i32, _i32, _i64, _i32, _i32, _i32, _vp, _vp],
    "mtb200_norm_act": [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _f32, _vp],
    "mtb200_in_bwd_reduce": [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _vp],
    "mtb200_in_bwd_apply": [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _vp,
                            _vp, _vp, _vp, _vp],
    "mtb200_lrelu_bwd": [_vp, _vp, _vp, _i32, _i64, _f32, _vp],
    "mtb200_residual_bwd": [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i64,
                            _i32, _f32, _vp],
    "mtb200_dcce_stats": [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i64, _vp, _vp, _vp],
    "mtb200_dcce_bwd": [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i64, _vp, _f32, _vp, _vp, _i32, _vp],
    "mtb200_mt_loss_stats": [_vp, _i32, _i32, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _vp],
    "mtb200_mt_loss_finalize": [_vp, _vp, _vp, _i32, _i32, _i64, _f32, _f32, _vp, _vp, _vp],
    "mtb200_mt_loss_bwd": [_vp, _i32, _i32, _i32, _vp, _i32, _i64, _vp, _i32, _vp, _vp, _vp, _i32, _vp],
    "mtb200_head_bwd_fused": [C.POINTER(HeadBwdParams), _vp],
    "mtb200_head_fwd_stats": [C.POINTER(HeadFwdParams), _vp],
    "mtb200_head_aggregate": [C.POINTER(HeadAggParams), _vp],
    "mtb200_sw_gather_tile": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32,
                              _vp],
    "mtb200_sw_aggregate": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _f32, _i32, _vp, _vp, _i32, _i32, _i32,
                            _i32, _i32, _i32, _vp],
    "mtb200_sw_finalize": [_vp, _vp, _i32, _i64, _vp, _vp, _vp],
    "mtb200_sw_finalize_slab": [_vp, _vp, _i32, _i64, _i64, _vp, _vp, _vp],
    "mtb200_sumsq": [_vp, _i64, _vp, _vp],
    "mtb200_sgd_step": [_vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32, _f32, _i32, _vp, _vp],
    "mtb200_loss_scale_update": [_vp, _vp, _f32, _f32, _i32, _vp],
    "mtb200_pack_weights": [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "mtb200_pack_weights_batched": [_vp, _i32, _i32, _i32, _vp],
    "mtb200_unpack_wgrad_batched": [_vp, _i32, _i32, _vp],
    "mtb200_unpack_wgrad": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp],
    "mtb200_ncdhw_to_ndhwc": [_vp, _i32, _i32, _i64, _vp, _i32, _i32, _i32, _i32, _vp],
    "mtb200_ndhwc_to_ncdhw": [_vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp, _vp],
    "mtb200_crop_pad": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp],
    "mtb200_resize_nearest": [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp],
    "mtb200_resample_probs": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp],
}

_lib = None
launch_count = 0  # number of library kernels launched through this binding (bench.py reports it)


class Mtb200Error(RuntimeError):
    pass


def lib():
    """Load libmtb200.so once.  Raises if it is missing: there is deliberately no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mtb200Error(
                "libmtb200.so not found at %s -- build it with `python -m multitalent_b200.build` "
                "(nvcc, sm_100a). The native CUDA library is required; there is no CPU/eager fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = C.c_char_p if name in ("mtb200_last_error", "mtb200_last_kernel") else C.c_int
        _lib = l
    return _lib


class KernelProfile:
    """Per-launch CUDA-event timing of library calls on the current stream (bench.py's roofline section).
    `with KernelProfile() as kp: ...; kp.summary()` -> {entry point: {launches, ms, flops, bytes}}."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _profile
        _profile = self
        return self

    def __exit__(self, *exc):
        global _profile
        _profile = None

    def per_launch(self):
        """[(tag, info, ms, flops, bytes)] in launch order (tools/layer_profile.py)."""
        torch.cuda.synchronize()
        return [(name, info, e0.elapsed_time(e1), flops, nbytes) for name, e0, e1, flops, nbytes, info, _k in self.records]

    def per_launch_kernels(self):
        """[(tag, info, ms, flops, bytes, CUDA kernel)] in launch order."""
        torch.cuda.synchronize()
        return [(name, info, e0.elapsed_time(e1), flops, nbytes, k) for name, e0, e1, flops, nbytes, info, k in self.records]

    def per_cuda_kernel(self):
        """{CUDA kernel family the library dispatched to (mtb200_last_kernel): {launches, ms, flops}}."""
        torch.cuda.synchronize()
        out = {}
        for _name, e0, e1, flops, nb, _info, kern in self.records:
            kern = kern.split("+")[0]  # "+red" = the same kernel with the fused reduction epilogue
            d = out.setdefault(kern, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nb
        return out

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, flops, nbytes, _info, _k in self.records:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


_profile = None


def call(name, *args, flops=0.0, nbytes=0.0, tag=None, info=None):
    """Invoke an entry point, raise Mtb200Error with the library's message on a negative status.
    `flops` / `nbytes` = ALGORITHMIC work of this launch (only used when a KernelProfile is active)."""
    global launch_count
    l = lib()
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    r = getattr(l, name)(*args)
    if r != 0:
        msg = l.mtb200_last_error()
        raise Mtb200Error("%s failed (%d): %s" % (name, r, msg.decode() if msg else "?"))
    if _profile is not None:
        e1.record()
        kern = l.mtb200_last_kernel()
        _profile.records.append((tag or name, e0, e1, flops, nbytes, info, kern.decode() if kern else name))
    launch_count += 1
    return r


def ptr(t):
    """Raw device pointer of a tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
