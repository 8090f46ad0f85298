"""TEST INFRASTRUCTURE ONLY -- ctypes binding of ``oracle/sgv3d_oracle.c`` (the exact-order C
restatement) and of ``oracle/_ref/libvoxel_pooling_ref.so`` (the reference's own CUDA kernel,
compiled unmodified from ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libsgv3d_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libvoxel_pooling_ref.so")

ARITH_SEQ = 0  # torch-CPU order: separately rounded mul / add
ARITH_FMA = 1  # k-ascending FMA chain
ARITH_PAIR = 2  # fma(a1,b1,a0*b0) + fma(a3,b3,a2*b2): torch CUDA bmm (cuBLAS) on B200

_f = ctypes.POINTER(ctypes.c_float)
_d = ctypes.POINTER(ctypes.c_double)
_i = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> None:
    """(Re)build the C oracle and -- when /root/reference is present -- the reference kernel."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(
            os.path.join(_HERE, "sgv3d_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(_REF)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return lib().oracle_num_threads()


def geometry(mode, u_tab, v_tab, z_tab, ida_inv, m_virtual, m_ego, ref_h, bda):
    """(B, Nc, D, fH, fW, 3) fp32; see oracle_geometry in sgv3d_oracle.c."""
    ida_inv, m_virtual, m_ego = _c32(ida_inv), _c32(m_virtual), _c32(m_ego)
    b, nc = ida_inv.shape[:2]
    u_tab, v_tab, z_tab, ref_h = _c32(u_tab), _c32(v_tab), _c32(z_tab), _c32(ref_h)
    bda = _c32(bda) if bda is not None else None
    d, fh, fw = len(z_tab), len(v_tab), len(u_tab)
    out = np.empty((b, nc, d, fh, fw, 3), np.float32)
    lib().oracle_geometry(ctypes.c_int(mode), b, nc, d, fh, fw, _p(u_tab, _f), _p(v_tab, _f),
                          _p(z_tab, _f), _p(ida_inv, _f), _p(m_virtual, _f), _p(m_ego, _f),
                          _p(bda, _f), _p(ref_h, _f), _p(out, _f))
    return out


def quantize(xyz, lower, size):
    xyz = _c32(xyz)
    lower, size = _c32(lower), _c32(size)
    idx = np.empty(xyz.shape, np.int32)
    lib().oracle_quantize(ctypes.c_long(xyz.size // 3), _p(xyz, _f), _p(lower, _f), _p(size, _f),
                          _p(idx, _i))
    return idx


def voxel_pooling_forward(geom, feat, nx, ny, nz, acc64=False):
    """Returns (out (B,Y,X,C) fp32 or fp64, pos_memo (B,N,3) int32)."""
    geom = np.ascontiguousarray(geom, np.int32)
    feat = _c32(feat)
    b = geom.shape[0]
    geom = geom.reshape(b, -1, 3)
    feat = feat.reshape(b, geom.shape[1], -1)
    n, c = feat.shape[1], feat.shape[2]
    pos = np.full((b, n, 3), -1, np.int32)
    if acc64:
        out = np.zeros((b, ny, nx, c), np.float64)
        lib().oracle_voxel_pooling_forward(b, n, c, nx, ny, nz, _p(geom, _i), _p(feat, _f), None,
                                           _p(pos, _i), _p(out, _d))
    else:
        out = np.zeros((b, ny, nx, c), np.float32)
        lib().oracle_voxel_pooling_forward(b, n, c, nx, ny, nz, _p(geom, _i), _p(feat, _f),
                                           _p(out, _f), _p(pos, _i), None)
    return out, pos


def voxel_pooling_backward(grad_out, pos_memo, c):
    """grad_out (B,C,Y,X) fp32 contiguous -> grad_feat (B,N,C) fp32."""
    grad_out = _c32(grad_out)
    pos_memo = np.ascontiguousarray(pos_memo, np.int32)
    b, n, _ = pos_memo.shape
    _, cc, ny, nx = grad_out.shape
    assert cc == c
    g = np.empty((b, n, c), np.float32)
    lib().oracle_voxel_pooling_backward(b, n, c, nx, ny, _p(grad_out, _f), _p(pos_memo, _i),
                                        _p(g, _f))
    return g


def lift_splat_forward64(idx, height, ctx, nx, ny, nz):
    """idx (B,Nc,D,fH,fW,3) int32, height (B*Nc,D,fH,fW), ctx (B*Nc,C,fH,fW) -> (B,C,Y,X) f64."""
    idx = np.ascontiguousarray(idx, np.int32)
    height, ctx = _c32(height), _c32(ctx)
    b, nc, d, fh, fw, _ = idx.shape
    c = ctx.shape[1]
    out = np.zeros((b, c, ny, nx), np.float64)
    lib().oracle_lift_splat_forward64(b, nc, d, fh, fw, c, nx, ny, nz, _p(idx, _i), _p(height, _f),
                                      _p(ctx, _f), _p(out, _d))
    return out


def lift_splat_backward64(idx, height, ctx, grad_bev, nx, ny, nz):
    idx = np.ascontiguousarray(idx, np.int32)
    height, ctx, grad_bev = _c32(height), _c32(ctx), _c32(grad_bev)
    b, nc, d, fh, fw, _ = idx.shape
    c = ctx.shape[1]
    gh = np.zeros(height.shape, np.float64)
    gc = np.zeros(ctx.shape, np.float64)
    lib().oracle_lift_splat_backward64(b, nc, d, fh, fw, c, nx, ny, nz, _p(idx, _i), _p(height, _f),
                                       _p(ctx, _f), _p(grad_bev, _f), _p(gh, _d), _p(gc, _d))
    return gh, gc


# ---- the reference's own CUDA kernel (needs a GPU; used by -m gpu tests and bench) ----------
_REF_SYMBOL = b"_Z37voxel_pooling_forward_kernel_launcheriiiiiiPKiPKfPfPiP11CUstream_st"
_ref = None


def reference_kernel_available() -> bool:
    return os.path.exists(_REF)


def reference_voxel_pooling_forward(b, n, c, nx, ny, nz, geom_ptr, feat_ptr, out_ptr, pos_ptr,
                                    stream):
    """Launch the UNMODIFIED reference kernel (voxel_pooling_forward_cuda.cu:38-56) on raw device
    pointers.  ``out`` must be zero-filled and ``pos_memo`` -1-filled by the caller
    (voxel_pooling.py:37-40)."""
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(_REF)
    fn = getattr(_ref, _REF_SYMBOL.decode())
    fn.restype = None
    fn(ctypes.c_int(b), ctypes.c_int(n), ctypes.c_int(c), ctypes.c_int(nx), ctypes.c_int(ny),
       ctypes.c_int(nz), ctypes.c_void_p(geom_ptr), ctypes.c_void_p(feat_ptr),
       ctypes.c_void_p(out_ptr), ctypes.c_void_p(pos_ptr), ctypes.c_void_p(stream))
