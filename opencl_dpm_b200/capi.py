"""ctypes binding of include/dpm_b200.h (the C ABI of libdpm_b200.so).

Fails loudly when the CUDA extension is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdpm_b200.so")

DPM3D_VOLUME, DPM3D_AREA, DPM3D_STICK, DPM3D_REPEL, DPM3D_ALL = 1, 2, 4, 8, 15
DPM2D_AREA, DPM2D_PERIMETER, DPM2D_BENDING, DPM2D_ATTRACT, DPM2D_REPEL, DPM2D_ALL = 1, 2, 4, 8, 16, 31


class DpmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dpm_b200 error {code}: {msg}")
        self.code = code


class Grid(C.Structure):
    _fields_ = [("nb", C.c_int32 * 3), ("periodic", C.c_int32 * 3), ("allpass", C.c_int32 * 3),
                ("origin", C.c_float * 3), ("inv_binw", C.c_float * 3), ("max_ext", C.c_float),
                ("margin", C.c_float), ("nbins", C.c_int32), ("pad", C.c_int32)]

    def as_tuple(self):
        return (tuple(self.nb), tuple(self.periodic), tuple(self.allpass), tuple(self.origin),
                tuple(self.inv_binw), self.max_ext, self.margin, self.nbins)


class Stats(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("launches", C.c_uint64), ("rebuilds", C.c_uint64),
                ("contact_evals", C.c_uint64), ("halo_bytes", C.c_uint64), ("reserved", C.c_uint64 * 3)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        global LIB_PATH
        LIB_PATH = os.environ.get("DPM_B200_LIB", LIB_PATH)  # development: an alternative build of the same library
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m opencl_dpm_b200.build` "
                              "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.dpm_version.restype = C.c_char_p
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        buf = C.create_string_buffer(1024)
        lib().dpm_last_error(buf, 1024)
        raise DpmError(rc, buf.value.decode(errors="replace"))


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f32(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} floats, got {a.size}")
    return a


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().dpm_device_count(C.byref(n)))
    return n.value


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(lib().dpm_nccl_unique_id(buf))
    return bytes(buf)


def icosphere(subdiv: int = 2):
    nv, nf = 10 * 4 ** subdiv + 2, 20 * 4 ** subdiv
    V = np.zeros((nv, 3), np.float32)
    F = np.zeros((nf, 3), np.uint32)
    a, b = C.c_int(0), C.c_int(0)
    _check(lib().dpm_icosphere(subdiv, _fp(V), F.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(a), C.byref(b)))
    assert (a.value, b.value) == (nv, nf)
    return V, F


def cell3d_params(calA: float, r0: float, nf: int):
    out = np.zeros(4, np.float32)
    _check(lib().dpm_cell3d_params(C.c_float(calA), C.c_float(r0), nf, _fp(out)))
    return dict(v0=out[0], sa0=out[1], a0=out[2], l0=out[3])


class Dpm3D:
    """Handle over the 3D path (dpm3d_* entry points)."""

    def __init__(self, ncells: int, nv: int, faces: np.ndarray, device: int = 0):
        self.nc, self.nv = int(ncells), int(nv)
        self.faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        self.nf = self.faces.shape[0]
        self.K = 32
        self._h = C.c_void_p()
        _check(lib().dpm3d_create(C.byref(self._h), device, self.nc, self.nv, self.nf,
                                  self.faces.ctypes.data_as(C.POINTER(C.c_uint32))))

    def close(self):
        if self._h:
            lib().dpm3d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        _check(lib().dpm3d_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_neighbor_params(self, skin_rel: float = 0.1, max_candidates: int = 32):
        _check(lib().dpm3d_set_neighbor_params(self._h, C.c_float(skin_rel), int(max_candidates)))
        self.K = int(max_candidates)

    def set_force_mask(self, mask: int):
        _check(lib().dpm3d_set_force_mask(self._h, C.c_uint(mask)))

    def shard_init(self, rank: int, nranks: int, unique_id: bytes, max_ghost: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().dpm3d_shard_init(self._h, int(rank), int(nranks), buf, int(max_ghost)))

    def set_global_ids(self, gid):
        g = np.ascontiguousarray(gid, dtype=np.int32)
        assert g.size == self.nc
        _check(lib().dpm3d_set_global_ids(self._h, _ip(g)))

    def set_compat(self, stale_volume_from_face: int = -1):
        """>= 0: reproduce the reference's volume race as it resolves on NVIDIA OpenCL (see include/dpm_b200.h)."""
        _check(lib().dpm3d_set_compat(self._h, int(stale_volume_from_face)))

    def _params(self, Kv, Ka, Ks, v0, a0, l0):
        return [_f32(np.broadcast_to(np.asarray(x, np.float32), (self.nc,)), self.nc) for x in (Kv, Ka, Ks, v0, a0, l0)]

    def upload(self, verts4, Kv, Ka, Ks, v0, a0, l0):
        v = _f32(verts4, self.nc * self.nv * 4)
        p = self._params(Kv, Ka, Ks, v0, a0, l0)
        _check(lib().dpm3d_upload(self._h, _fp(v), *[_fp(x) for x in p]))

    def upload_device(self, dev_ptr: int, Kv, Ka, Ks, v0, a0, l0):
        p = self._params(Kv, Ka, Ks, v0, a0, l0)
        _check(lib().dpm3d_upload_device(self._h, C.cast(C.c_void_p(dev_ptr), C.POINTER(C.c_float)), *[_fp(x) for x in p]))

    def step(self, nsteps: int, dt: float, Kre: float, Kat: float = 0.0, pbc: int = 1, L: float = 1.0):
        _check(lib().dpm3d_step(self._h, int(nsteps), C.c_float(dt), C.c_float(Kre), C.c_float(Kat), int(pbc), C.c_float(L)))

    def sync(self):
        _check(lib().dpm3d_sync(self._h))

    def download(self, want_forces: bool = True):
        v = np.empty((self.nc * self.nv, 4), np.float32)
        f = np.empty((self.nc * self.nv, 4), np.float32) if want_forces else None
        _check(lib().dpm3d_download(self._h, _fp(v), _fp(f)))
        return v, f

    def euler_update(self, verts4, Kv, Ka, Ks, v0, a0, l0, nsteps, dt, Kre, Kat=0.0, pbc=1, L=1.0, forces_out=None):
        """The reference seam in one call (host buffers in/out). verts4 is updated in place."""
        assert verts4.dtype == np.float32 and verts4.flags.c_contiguous and verts4.size == self.nc * self.nv * 4
        p = self._params(Kv, Ka, Ks, v0, a0, l0)
        ms = C.c_float(0)
        _check(lib().dpm3d_euler_update(self._h, _fp(verts4), _fp(forces_out), *[_fp(x) for x in p], int(nsteps),
                                        C.c_float(dt), C.c_float(Kre), C.c_float(Kat), int(pbc), C.c_float(L), C.byref(ms)))
        return ms.value

    def rebuild_neighbors(self, pbc: int, L: float):
        _check(lib().dpm3d_rebuild_neighbors(self._h, int(pbc), C.c_float(L)))

    def neighbor_params(self):
        """(skin_rel, max_candidates) in force — euler_update may have grown K after a candidate-list overflow"""
        s, k = C.c_float(0), C.c_int(0)
        _check(lib().dpm3d_get_neighbor_params(self._h, C.byref(s), C.byref(k)))
        self.K = k.value
        return s.value, k.value

    def neighbor_artifacts(self):
        g = Grid()
        self.neighbor_params()  # the library's current K sizes `cand` (never a stale Python-side copy)
        _check(lib().dpm3d_get_neighbor_artifacts(self._h, C.byref(g), None, None, None, None, None))
        bin_id = np.empty(self.nc, np.int32); order = np.empty(self.nc, np.int32)
        bin_start = np.empty(g.nbins + 1, np.int32); cc = np.empty(self.nc, np.int32)
        cand = np.empty((self.nc, self.K), np.int32)
        _check(lib().dpm3d_get_neighbor_artifacts(self._h, C.byref(g), _ip(bin_id), _ip(order), _ip(bin_start), _ip(cc), _ip(cand)))
        return dict(grid=g, bin_id=bin_id, order=order, bin_start=bin_start, cand_count=cc, cand=cand)

    def cell_bounds(self):
        b = np.empty((self.nc, 12), np.float32)
        _check(lib().dpm3d_get_cell_bounds(self._h, _fp(b)))
        return b

    def stats(self) -> Stats:
        s = Stats()
        _check(lib().dpm3d_get_stats(self._h, C.byref(s)))
        return s

    def device_state(self):
        a, b = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
        _check(lib().dpm3d_device_state(self._h, C.byref(a), C.byref(b)))
        return C.cast(a, C.c_void_p).value, C.cast(b, C.c_void_p).value


class Dpm2D:
    """Handle over the 2D path (dpm2d_* entry points)."""

    def __init__(self, ncells: int, max_nv: int, device: int = 0):
        self.nc, self.S = int(ncells), int(max_nv)
        self.K = 32
        self._h = C.c_void_p()
        _check(lib().dpm2d_create(C.byref(self._h), device, self.nc, self.S))

    def close(self):
        if self._h:
            lib().dpm2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        _check(lib().dpm2d_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_neighbor_params(self, skin_rel: float = 0.1, max_candidates: int = 32):
        _check(lib().dpm2d_set_neighbor_params(self._h, C.c_float(skin_rel), int(max_candidates)))
        self.K = int(max_candidates)

    def set_force_mask(self, mask: int):
        _check(lib().dpm2d_set_force_mask(self._h, C.c_uint(mask)))

    def _params(self, *xs):
        return [_f32(np.broadcast_to(np.asarray(x, np.float32), (self.nc,)), self.nc) for x in xs]

    def upload(self, verts2, nv, Ka, Kl, Kb, a0, l0, r0):
        v = _f32(verts2, self.nc * self.S * 2)
        n = np.ascontiguousarray(np.broadcast_to(np.asarray(nv, np.int32), (self.nc,)), dtype=np.int32)
        p = self._params(Ka, Kl, Kb, a0, l0, r0)
        _check(lib().dpm2d_upload(self._h, _fp(v), _ip(n), *[_fp(x) for x in p]))

    def step(self, nsteps, dt, Kre, Kat=0.0, pbc=1, L=1.0):
        _check(lib().dpm2d_step(self._h, int(nsteps), C.c_float(dt), C.c_float(Kre), C.c_float(Kat), int(pbc), C.c_float(L)))

    def sync(self):
        _check(lib().dpm2d_sync(self._h))

    def download(self, want_forces=True):
        v = np.empty((self.nc, self.S, 2), np.float32)
        f = np.empty((self.nc, self.S, 2), np.float32) if want_forces else None
        _check(lib().dpm2d_download(self._h, _fp(v), _fp(f)))
        return v, f

    def euler_update(self, verts2, nv, Ka, Kl, Kb, a0, l0, r0, nsteps, dt, Kre, Kat=0.0, pbc=1, L=1.0, forces_out=None):
        assert verts2.dtype == np.float32 and verts2.flags.c_contiguous and verts2.size == self.nc * self.S * 2
        n = np.ascontiguousarray(np.broadcast_to(np.asarray(nv, np.int32), (self.nc,)), dtype=np.int32)
        p = self._params(Ka, Kl, Kb, a0, l0, r0)
        ms = C.c_float(0)
        _check(lib().dpm2d_euler_update(self._h, _fp(verts2), _fp(forces_out), _ip(n), *[_fp(x) for x in p], int(nsteps),
                                        C.c_float(dt), C.c_float(Kre), C.c_float(Kat), int(pbc), C.c_float(L), C.byref(ms)))
        return ms.value

    def rebuild_neighbors(self, Kat, pbc, L):
        _check(lib().dpm2d_rebuild_neighbors(self._h, C.c_float(Kat), int(pbc), C.c_float(L)))

    def neighbor_params(self):
        s, k = C.c_float(0), C.c_int(0)
        _check(lib().dpm2d_get_neighbor_params(self._h, C.byref(s), C.byref(k)))
        self.K = k.value
        return s.value, k.value

    def neighbor_artifacts(self):
        g = Grid()
        self.neighbor_params()  # the library's current K sizes `cand`
        _check(lib().dpm2d_get_neighbor_artifacts(self._h, C.byref(g), None, None, None, None, None))
        bin_id = np.empty(self.nc, np.int32); order = np.empty(self.nc, np.int32)
        bin_start = np.empty(g.nbins + 1, np.int32); cc = np.empty(self.nc, np.int32)
        cand = np.empty((self.nc, self.K), np.int32)
        _check(lib().dpm2d_get_neighbor_artifacts(self._h, C.byref(g), _ip(bin_id), _ip(order), _ip(bin_start), _ip(cc), _ip(cand)))
        return dict(grid=g, bin_id=bin_id, order=order, bin_start=bin_start, cand_count=cc, cand=cand)

    def stats(self) -> Stats:
        s = Stats()
        _check(lib().dpm2d_get_stats(self._h, C.byref(s)))
        return s
