"""ctypes binding of ``libsgv3d_b200.so`` (the C ABI declared in ``include/sgv3d_b200.h``).

The library is the only compute path: there is no CPU or eager fallback.  If the shared object
is missing or a symbol cannot be resolved this module raises immediately.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsgv3d_b200.so")

ARITH_SEQ, ARITH_FMA, ARITH_PAIR = 0, 1, 2
DTYPE_F32, DTYPE_BF16 = 0, 1
ABI_VERSION = 2

#: every symbol ``include/sgv3d_b200.h`` declares (checked by tests/test_abi.py)
SYMBOLS = (
    "sgv3d_abi_version", "sgv3d_last_error", "sgv3d_launch_count",
    "sgv3d_voxel_pooling_workspace_bytes", "sgv3d_voxel_pooling_forward",
    "sgv3d_voxel_pooling_backward_workspace_bytes", "sgv3d_voxel_pooling_backward",
    "sgv3d_geometry_quantize", "sgv3d_inverse4x4", "sgv3d_camera_prep",
    "sgv3d_lift_splat_workspace_bytes", "sgv3d_lift_splat_plan", "sgv3d_lift_splat_forward",
    "sgv3d_lift_splat_forward_bsm", "sgv3d_lift_splat_backward", "sgv3d_lift_splat_plan_expand",
    "sgv3d_lift_splat_uses_block_pipeline", "sgv3d_lift_splat_backward_bsm",
    "sgv3d_profile_enable", "sgv3d_profile_report",
)


class LiftSplatDesc(ctypes.Structure):
    """``struct sgv3d_lift_splat_desc`` (include/sgv3d_b200.h)."""
    _fields_ = ([(n, c_int32) for n in ("B", "Nc", "D", "fH", "fW", "C", "X", "Y", "Z", "arith", "ctx_dtype",
                                        "height_is_logits")]
                + [(n, c_int64) for n in ("height_batch_stride", "ctx_batch_stride",
                                          "grad_height_batch_stride", "grad_ctx_batch_stride")]
                + [("reserved", c_int32 * 4)])


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m sgv3d_b200.build` "
            "(nvcc, sm_100a). sgv3d_b200 has no fallback path.")
    L = ctypes.CDLL(LIB_PATH)
    for name in SYMBOLS:
        if not hasattr(L, name):
            raise RuntimeError(f"{LIB_PATH} does not export {name}")
    L.sgv3d_abi_version.restype = c_int
    L.sgv3d_last_error.restype = ctypes.c_char_p
    L.sgv3d_launch_count.restype = c_int64
    L.sgv3d_launch_count.argtypes = [c_int]
    L.sgv3d_voxel_pooling_workspace_bytes.restype = c_size_t
    L.sgv3d_voxel_pooling_workspace_bytes.argtypes = [c_int] * 6
    L.sgv3d_voxel_pooling_forward.restype = c_int
    L.sgv3d_voxel_pooling_forward.argtypes = [c_int] * 6 + [c_void_p] * 5 + [c_size_t, c_void_p]
    L.sgv3d_voxel_pooling_backward_workspace_bytes.restype = c_size_t
    L.sgv3d_voxel_pooling_backward_workspace_bytes.argtypes = [c_int] * 4
    L.sgv3d_voxel_pooling_backward.restype = c_int
    L.sgv3d_voxel_pooling_backward.argtypes = ([c_int] * 5 + [c_void_p] + [c_int64] * 4
                                               + [c_void_p] * 3 + [c_size_t, c_void_p])
    L.sgv3d_geometry_quantize.restype = c_int
    L.sgv3d_geometry_quantize.argtypes = [c_int] * 6 + [c_void_p] * 13
    L.sgv3d_inverse4x4.restype = c_int
    L.sgv3d_inverse4x4.argtypes = [c_int] + [c_void_p] * 7
    L.sgv3d_camera_prep.restype = c_int
    L.sgv3d_camera_prep.argtypes = [c_int, c_int] + [c_void_p] * 8
    P = ctypes.POINTER(LiftSplatDesc)
    L.sgv3d_lift_splat_workspace_bytes.restype = c_size_t
    L.sgv3d_lift_splat_workspace_bytes.argtypes = [P]
    L.sgv3d_lift_splat_uses_block_pipeline.restype = c_int
    L.sgv3d_lift_splat_uses_block_pipeline.argtypes = [P]
    L.sgv3d_lift_splat_plan.restype = c_int
    L.sgv3d_lift_splat_plan.argtypes = [P] + [c_void_p] * 11 + [c_size_t, c_void_p]
    L.sgv3d_lift_splat_forward.restype = c_int
    L.sgv3d_lift_splat_forward.argtypes = [P] + [c_void_p] * 4 + [c_size_t, c_void_p]
    L.sgv3d_lift_splat_forward_bsm.restype = c_int
    L.sgv3d_lift_splat_forward_bsm.argtypes = ([P] + [c_void_p] * 3 + [c_int, c_int64, ctypes.c_float]
                                               + [c_void_p] * 2 + [c_size_t, c_void_p])
    L.sgv3d_lift_splat_backward.restype = c_int
    L.sgv3d_lift_splat_backward.argtypes = [P] + [c_void_p] * 6 + [c_size_t, c_void_p]
    L.sgv3d_lift_splat_backward_bsm.restype = c_int
    L.sgv3d_lift_splat_backward_bsm.argtypes = ([P] + [c_void_p] * 4 + [c_int, c_int64, ctypes.c_float]
                                                + [c_void_p] * 4 + [c_size_t, c_void_p])
    L.sgv3d_lift_splat_plan_expand.restype = c_int
    L.sgv3d_lift_splat_plan_expand.argtypes = [P] + [c_void_p] * 2 + [c_size_t, c_void_p]
    L.sgv3d_profile_enable.restype = c_int
    L.sgv3d_profile_enable.argtypes = [c_int]
    L.sgv3d_profile_report.restype = ctypes.c_long
    L.sgv3d_profile_report.argtypes = [ctypes.c_char_p, c_size_t]
    if L.sgv3d_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libsgv3d_b200 ABI {L.sgv3d_abi_version()} != expected {ABI_VERSION}")
    _lib = L
    return L


def check(status: int) -> None:
    """Non-zero C status -> RuntimeError (the reference calls ``exit(-1)`` instead,
    ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:51-55)."""
    if status != 0:
        msg = lib().sgv3d_last_error()
        raise RuntimeError(f"libsgv3d_b200 error {status}: {msg.decode() if msg else '?'}")


def launch_count(reset: bool = False) -> int:
    return int(lib().sgv3d_launch_count(1 if reset else 0))


def profile_enable(on: bool) -> None:
    lib().sgv3d_profile_enable(1 if on else 0)


def profile_report() -> dict:
    """{kernel name: (launches, total_ms)} since the last report (synchronises the recorded events)."""
    buf = ctypes.create_string_buffer(1 << 16)
    lib().sgv3d_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.rsplit(",", 2)
        out[name] = (int(n), float(ms))
    return out


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def host_f32x3(values):
    """HOST array of three floats for the ``lower3`` / ``size3`` arguments."""
    return (ctypes.c_float * 3)(*[float(v) for v in values])
