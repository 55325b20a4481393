"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/_ref/libref_dpm.so: the REAL reference (its own cell.cpp,
Tissue2D.cpp, Tissue3D.cpp and .cl kernels, see oracle/ref_shim.cpp).  Geometry and Disperse run anywhere;
CLEulerUpdate needs an OpenCL device (NVIDIA's ICD exists on the B200 box, not in the build container)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_dpm.so")
_lib = None


def present() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_disperse3d.restype = C.c_float
        _lib.ref_disperse2d.restype = C.c_float
    return _lib


def available() -> bool:
    """True when the reference's OpenCL compute path can run on this machine."""
    return present() and bool(lib().ref_available())


def device_name() -> str:
    buf = C.create_string_buffer(256)
    return buf.value.decode() if lib().ref_device_name(buf, 256) else ""


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def cell3d(start, calA, r0):
    v = np.zeros((162, 3), np.float32); f = np.zeros((320, 3), np.uint32); s = np.zeros(7, np.float32)
    st = np.asarray(start, np.float32)
    lib().ref_cell3d(_p(st), C.c_float(calA), C.c_float(r0), _p(v), _p(f), _p(s))
    return v, f, dict(calA0=s[0], r0=s[1], v0=s[2], sa0=s[3], a0=s[4], Volume=s[5], SurfaceArea=s[6])


def cell2d(x0, y0, calA, nv, r0):
    v = np.zeros((nv, 2), np.float32); s = np.zeros(5, np.float32)
    lib().ref_cell2d(C.c_float(x0), C.c_float(y0), C.c_float(calA), C.c_uint(nv), C.c_float(r0), _p(v), _p(s))
    return v, dict(calA0=s[0], a0=s[1], l0=s[2], r0=s[3], area=s[4])


def reset_drand48():
    C.CDLL(None).seed48((C.c_ushort * 3)(0, 0, 0))


def disperse3d(n, start, calA, r0, phi0):
    v = np.zeros((n * 162, 3), np.float32)
    st = np.asarray(start, np.float32)
    reset_drand48()
    L = lib().ref_disperse3d(n, _p(st), C.c_float(calA), C.c_float(r0), C.c_float(phi0), _p(v))
    return v, np.float32(L)


def disperse2d(n, calA, nv, r0, phi0):
    v = np.zeros((n, nv, 2), np.float32)
    reset_drand48()
    L = lib().ref_disperse2d(n, C.c_float(calA), C.c_uint(nv), C.c_float(r0), C.c_float(phi0), _p(v))
    return v, np.float32(L)


def euler3d(verts4, Kv, Ka, Ks, v0, a0, Kre, PBC, L, nsteps, dt):
    """Reference Tissue3D::CLEulerUpdate. verts4: (nc*162,4). Returns (verts4, forces4, seconds)."""
    V4 = np.asarray(verts4, np.float32).reshape(-1, 4)
    nc = V4.shape[0] // 162
    v3 = np.ascontiguousarray(V4[:, :3]); f3 = np.zeros_like(v3)
    P = [np.ascontiguousarray(np.broadcast_to(np.asarray(x, np.float32), (nc,))) for x in (Kv, Ka, Ks, v0, a0)]
    sec = C.c_double(0)
    rc = lib().ref3d_euler(nc, _p(v3), _p(f3), *[_p(x) for x in P], C.c_float(Kre), int(PBC), C.c_float(L), int(nsteps), C.c_float(dt), C.byref(sec))
    if rc != 0:
        raise RuntimeError("reference CLEulerUpdate failed: " + lib().ref_last_error().decode())
    Vo = np.zeros_like(V4); Fo = np.zeros_like(V4)
    Vo[:, :3] = v3; Fo[:, :3] = f3
    return Vo, Fo, sec.value


def euler2d(verts2, nv, Ka, Kl, Kb, a0, l0, r0, Kre, Kat, PBC, L, nsteps, dt):
    """Reference Tissue2D::CLEulerUpdate. verts2: (nc,S,2). Returns (verts2, forces2, seconds)."""
    V = np.array(verts2, np.float32, copy=True)
    nc, S = V.shape[0], V.shape[1]
    F = np.zeros_like(V)
    n = np.ascontiguousarray(np.broadcast_to(np.asarray(nv, np.int32), (nc,)))
    P = [np.ascontiguousarray(np.broadcast_to(np.asarray(x, np.float32), (nc,))) for x in (Ka, Kl, Kb, a0, l0, r0)]
    sec = C.c_double(0)
    rc = lib().ref2d_euler(nc, S, _p(n), _p(V), _p(F), *[_p(x) for x in P], C.c_float(Kre), C.c_float(Kat), int(PBC), C.c_float(L), int(nsteps),
                           C.c_float(dt), C.byref(sec))
    if rc != 0:
        raise RuntimeError("reference CLEulerUpdate failed: " + lib().ref_last_error().decode())
    return V, F, sec.value


def attract3d(verts4, l0, Kat, PBC, L):
    """The reference's AllVertAttraction kernel (shaders/Cell3D_Kernel.cl:313-364), unmodified text, launched on its own on
    zeroed forces (the reference host never enqueues it).  verts4: (nc*162,4).  Returns forces4."""
    V4 = np.asarray(verts4, np.float32).reshape(-1, 4)
    nc = V4.shape[0] // 162
    v3 = np.ascontiguousarray(V4[:, :3]); f3 = np.zeros_like(v3)
    l0a = np.ascontiguousarray(np.broadcast_to(np.asarray(l0, np.float32), (nc,)))
    rc = lib().ref3d_attract(nc, _p(v3), _p(f3), _p(l0a), C.c_float(L), int(PBC), C.c_float(Kat))
    if rc != 0:
        raise RuntimeError("reference AllVertAttraction failed: " + lib().ref_last_error().decode())
    Fo = np.zeros_like(V4)
    Fo[:, :3] = f3
    return Fo
