"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/liboracle_dpm.so (the CPU restatement of the
reference kernels).  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_dpm.so")

_lib = None


def build(force: bool = False) -> None:
    srcs = [os.path.join(HERE, f) for f in ("dpm_oracle.c", "dpm_oracle_impl.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "liboracle_dpm.so"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


class Grid(C.Structure):
    _fields_ = [("nb", C.c_int32 * 3), ("periodic", C.c_int32 * 3), ("allpass", C.c_int32 * 3),
                ("origin", C.c_float * 3), ("inv_binw", C.c_float * 3), ("max_ext", C.c_float),
                ("margin", C.c_float), ("nbins", C.c_int32), ("pad", C.c_int32)]

    def as_tuple(self):
        return (tuple(self.nb), tuple(self.periodic), tuple(self.allpass), tuple(self.origin),
                tuple(self.inv_binw), self.max_ext, self.margin, self.nbins)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _real(dtype):
    return (C.c_float, "_f32") if dtype == np.float32 else (C.c_double, "_f64")


def _arr(x, n, dtype):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype), (n,)), dtype=dtype)


# ---------------------------------------------------------------- 3D
def forces3d(verts4, faces, Kv, Ka, Ks, v0, a0, l0, Kc, PBC, L, which=15, cand_count=None, cand=None,
             dtype=np.float32, want_contacts=False):
    """Forces of one step from positions verts4 [(nc*nv),4]. Returns forces [(nc*nv),4] (and contacts)."""
    ct, sfx = _real(dtype)
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    nf = faces.shape[0]
    nc = len(np.atleast_1d(np.asarray(v0))) if np.ndim(v0) else None
    V = np.ascontiguousarray(verts4, dtype=dtype).reshape(-1, 4)
    if nc is None:
        raise ValueError("per-cell arrays required")
    nv = V.shape[0] // nc
    Fo = np.zeros_like(V)
    P = [_arr(x, nc, dtype) for x in (Kv, Ka, Ks, v0, a0, l0)]
    contacts = np.zeros((nc * nv, 2), np.int32) if want_contacts else None
    K = 0 if cand is None else cand.shape[1]
    fn = getattr(lib(), "oracle3d_forces" + sfx)
    fn.restype = None
    fn(nc, nv, nf, _p(faces, C.c_uint32), _p(V, ct), _p(Fo, ct), *[_p(x, ct) for x in P], ct(Kc), int(PBC), ct(L),
       int(which), _p(None if cand_count is None else np.ascontiguousarray(cand_count, np.int32), C.c_int32),
       _p(None if cand is None else np.ascontiguousarray(cand, np.int32), C.c_int32), int(K), _p(contacts, C.c_int32))
    return (Fo, contacts) if want_contacts else Fo


def forces3d_range(verts4, faces, Kv, Ka, Ks, v0, a0, l0, Kc, PBC, L, c0, c1, which=15):
    """All-pairs forces (fp32) for cells [c0,c1) only — a bounded sample of the reference algorithm."""
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    nc = len(np.asarray(v0))
    V = np.ascontiguousarray(verts4, dtype=np.float32).reshape(-1, 4)
    nv = V.shape[0] // nc
    Fo = np.zeros_like(V)
    P = [_arr(x, nc, np.float32) for x in (Kv, Ka, Ks, v0, a0, l0)]
    fn = lib().oracle3d_forces_range_f32
    fn.restype = None
    fn(nc, nv, faces.shape[0], _p(faces, C.c_uint32), _p(V, C.c_float), _p(Fo, C.c_float), *[_p(x, C.c_float) for x in P],
       C.c_float(Kc), int(PBC), C.c_float(L), int(which), int(c0), int(c1))
    return Fo


def repel_sample3d(verts4, faces, nc, Kc, PBC, L, ci, vi0, vi1):
    """All-pairs RepellingForces (fp32) of vertices [vi0, vi1) of cell ci against all nc cells in verts4 — the bounded
    sample of the reference algorithm that bench.py --impl reference times.  Returns [(vi1-vi0), 4]."""
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    V = np.ascontiguousarray(verts4, dtype=np.float32).reshape(-1, 4)
    nv = V.shape[0] // nc
    out = np.zeros((vi1 - vi0, 4), np.float32)
    fn = lib().oracle3d_repel_sample_f32
    fn.restype = None
    fn(int(nc), int(nv), faces.shape[0], _p(faces, C.c_uint32), _p(V, C.c_float), C.c_float(Kc), int(PBC), C.c_float(L),
       int(ci), int(vi0), int(vi1), _p(out, C.c_float))
    return out


def run3d(verts4, faces, Kv, Ka, Ks, v0, a0, l0, Kc, PBC, L, nsteps, dt, which=15, dtype=np.float32, stale_from=-1):
    """nsteps of the all-pairs reference algorithm. Returns (verts4, last_forces4).
    stale_from >= 0: emulate the reference's volume race as it resolves on NVIDIA OpenCL (faces >= stale_from use
    the previous step's volume, zero on the first step of the call)."""
    ct, sfx = _real(dtype)
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    nf = faces.shape[0]
    nc = len(np.asarray(v0))
    V = np.array(verts4, dtype=dtype, copy=True).reshape(-1, 4)
    nv = V.shape[0] // nc
    Fo = np.zeros_like(V)
    P = [_arr(x, nc, dtype) for x in (Kv, Ka, Ks, v0, a0, l0)]
    if stale_from >= 0:
        fn = getattr(lib(), "oracle3d_run_compat" + sfx)
        fn.restype = None
        fn(nc, nv, nf, _p(faces, C.c_uint32), _p(V, ct), _p(Fo, ct), *[_p(x, ct) for x in P], ct(Kc), int(PBC), ct(L),
           int(nsteps), ct(dt), int(which), int(stale_from))
        return V, Fo
    fn = getattr(lib(), "oracle3d_run" + sfx)
    fn.restype = None
    fn(nc, nv, nf, _p(faces, C.c_uint32), _p(V, ct), _p(Fo, ct), *[_p(x, ct) for x in P], ct(Kc), int(PBC), ct(L),
       int(nsteps), ct(dt), int(which))
    return V, Fo


def attract3d(verts4, l0, Kat, PBC, L, nc, forces=None, dtype=np.float32):
    """AllVertAttraction (shaders/Cell3D_Kernel.cl:313-364), literal scatter form, ACCUMULATED into `forces`
    (zeros if None).  Returns the forces array."""
    ct, sfx = _real(dtype)
    V = np.ascontiguousarray(verts4, dtype=dtype).reshape(-1, 4)
    nv = V.shape[0] // nc
    Fo = np.zeros_like(V) if forces is None else np.ascontiguousarray(forces, dtype=dtype).reshape(-1, 4)
    l0a = _arr(l0, nc, dtype)
    fn = getattr(lib(), "oracle3d_attract" + sfx)
    fn.restype = None
    fn(nc, nv, _p(V, ct), _p(Fo, ct), _p(l0a, ct), ct(L), int(PBC), ct(Kat))
    return Fo


def run3d_attract(verts4, faces, Kv, Ka, Ks, v0, a0, l0, Kc, Kat, PBC, L, nsteps, dt, which=15, dtype=np.float32):
    """nsteps of {the six live kernels + AllVertAttraction; Euler}: the all-pairs reference algorithm with the dead
    attraction kernel enqueued between RepellingForces and EulerPosition.  Returns (verts4, last_forces4)."""
    nc = len(np.asarray(v0))
    V = np.array(verts4, dtype=dtype, copy=True).reshape(-1, 4)
    Fo = np.zeros_like(V)
    for _ in range(int(nsteps)):
        Fo = forces3d(V, faces, Kv, Ka, Ks, v0, a0, l0, Kc, PBC, L, which=which, dtype=dtype)
        Fo = attract3d(V, l0, Kat, PBC, L, nc, forces=Fo, dtype=dtype)
        V[:, :3] += Fo[:, :3] * dtype(dt)
    return V, Fo


def run3d_culled(verts4, faces, Kv, Ka, Ks, v0, a0, l0, Kc, PBC, L, nsteps, dt, dtype=np.float32, rebuild_every=1):
    """nsteps of the CULLED form (CPU cell list + literal kernels; equal to the all-pairs form, see the tests) — fast
    enough for 1e4-step trajectories.  The candidate lists are rebuilt from the current fp32 AABBs every
    `rebuild_every` steps with a generous margin."""
    ct, sfx = _real(dtype)
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    nc = len(np.asarray(v0))
    V = np.array(verts4, dtype=dtype, copy=True).reshape(-1, 4)
    nv = V.shape[0] // nc
    P = [_arr(x, nc, dtype) for x in (Kv, Ka, Ks, v0, a0, l0)]
    Fo = np.zeros_like(V)
    cl = None
    for s in range(int(nsteps)):
        if cl is None or s % rebuild_every == 0:
            V32 = V.astype(np.float32)
            lo, hi = aabb3d(V32, nc)
            Vc = V32.reshape(nc, nv, 4)
            emax = max(float(np.linalg.norm(Vc[:, faces[:, i], :3] - Vc[:, faces[:, (i + 1) % 3], :3], axis=2).max()) for i in range(3))
            cl = cell_list(3, lo, hi, PBC, L, 0.3, 1.5 * 0.34 * emax, 64)
            assert cl["cand_count"].max() <= 64
        Fo = forces3d(V, faces, *P, Kc, PBC, L, cand_count=cl["cand_count"], cand=cl["cand"], dtype=dtype)
        V[:, :3] += Fo[:, :3] * dtype(dt)
    return V, Fo


def aabb3d(verts4, nc):
    V = np.ascontiguousarray(verts4, np.float32).reshape(-1, 4)
    nv = V.shape[0] // nc
    lo = np.zeros((nc, 3), np.float32); hi = np.zeros((nc, 3), np.float32)
    lib().oracle_aabb3d(nc, nv, _p(V, C.c_float), _p(lo, C.c_float), _p(hi, C.c_float))
    return lo, hi


# ---------------------------------------------------------------- 2D
def forces2d(verts2, NV, Ka, Kl, Kb, a0, l0, r0, Kre, Kat, PBC, L, which=31, cand_count=None, cand=None,
             dtype=np.float32, want_inside=False):
    ct, sfx = _real(dtype)
    V = np.ascontiguousarray(verts2, dtype=dtype)
    nc, S = V.shape[0], V.shape[1]
    NV = _arr(NV, nc, np.int32)
    Fo = np.zeros_like(V)
    P = [_arr(x, nc, dtype) for x in (Ka, Kl, Kb, a0, l0, r0)]
    inside = np.zeros((nc, S), np.int32) if want_inside else None
    K = 0 if cand is None else cand.shape[1]
    fn = getattr(lib(), "oracle2d_forces" + sfx)
    fn.restype = None
    fn(nc, S, _p(NV, C.c_int32), _p(V, ct), _p(Fo, ct), *[_p(x, ct) for x in P], ct(Kre), ct(Kat), int(PBC), ct(L),
       int(which), _p(None if cand_count is None else np.ascontiguousarray(cand_count, np.int32), C.c_int32),
       _p(None if cand is None else np.ascontiguousarray(cand, np.int32), C.c_int32), int(K), _p(inside, C.c_int32))
    return (Fo, inside) if want_inside else Fo


def forces2d_range(verts2, NV, Ka, Kl, Kb, a0, l0, r0, Kre, Kat, PBC, L, c0, c1, which=31):
    """All-pairs 2D forces (fp32) of cells [c0, c1) only — a bounded sample of the reference algorithm."""
    V = np.ascontiguousarray(verts2, dtype=np.float32)
    nc, S = V.shape[0], V.shape[1]
    NV = _arr(NV, nc, np.int32)
    Fo = np.zeros_like(V)
    P = [_arr(x, nc, np.float32) for x in (Ka, Kl, Kb, a0, l0, r0)]
    fn = lib().oracle2d_forces_range_f32
    fn.restype = None
    fn(nc, S, _p(NV, C.c_int32), _p(V, C.c_float), _p(Fo, C.c_float), *[_p(x, C.c_float) for x in P], C.c_float(Kre),
       C.c_float(Kat), int(PBC), C.c_float(L), int(which), int(c0), int(c1))
    return Fo


def run2d(verts2, NV, Ka, Kl, Kb, a0, l0, r0, Kre, Kat, PBC, L, nsteps, dt, which=31, dtype=np.float32):
    ct, sfx = _real(dtype)
    V = np.array(verts2, dtype=dtype, copy=True)
    nc, S = V.shape[0], V.shape[1]
    NV = _arr(NV, nc, np.int32)
    Fo = np.zeros_like(V)
    P = [_arr(x, nc, dtype) for x in (Ka, Kl, Kb, a0, l0, r0)]
    fn = getattr(lib(), "oracle2d_run" + sfx)
    fn.restype = None
    fn(nc, S, _p(NV, C.c_int32), _p(V, ct), _p(Fo, ct), *[_p(x, ct) for x in P], ct(Kre), ct(Kat), int(PBC), ct(L),
       int(nsteps), ct(dt), int(which))
    return V, Fo


def run2d_culled(verts2, NV, Ka, Kl, Kb, a0, l0, r0, Kre, Kat, PBC, L, nsteps, dt, dtype=np.float32, rebuild_every=1, K=64):
    """nsteps of the CULLED 2D form (CPU cell list incl. the |d| > L wrap partners + literal kernels); usable at
    config B's 4096 cells.  Candidate lists are rebuilt from the current fp32 AABBs every `rebuild_every` steps."""
    V = np.array(verts2, dtype=dtype, copy=True)
    nc = V.shape[0]
    NVa = _arr(NV, nc, np.int32)
    l0a = _arr(l0, nc, np.float32)
    Fo = np.zeros_like(V)
    cl = None
    for s in range(int(nsteps)):
        if cl is None or s % rebuild_every == 0:
            lo, hi = aabb2d(V.astype(np.float32), NVa)
            cl = cell_list(2, lo, hi, PBC, L, 0.3, float(l0a.max()) if Kat != 0 else 0.0, K, far2d=True)
            assert cl["cand_count"].max() <= K, "2D candidate list overflow in the oracle"
        Fo = forces2d(V, NVa, Ka, Kl, Kb, a0, l0, r0, Kre, Kat, PBC, L, cand_count=cl["cand_count"], cand=cl["cand"], dtype=dtype)
        V += Fo * dtype(dt)
    return V, Fo


def aabb2d(verts2, NV):
    V = np.ascontiguousarray(verts2, np.float32)
    nc, S = V.shape[0], V.shape[1]
    NV = _arr(NV, nc, np.int32)
    lo = np.zeros((nc, 3), np.float32); hi = np.zeros((nc, 3), np.float32)
    lib().oracle_aabb2d(nc, S, _p(NV, C.c_int32), _p(V, C.c_float), _p(lo, C.c_float), _p(hi, C.c_float))
    return lo, hi


# ---------------------------------------------------------------- cell list spec
def cell_list(nd, lo, hi, PBC, L, skin_rel, rng, K, cap=None, far2d=False):
    lo = np.ascontiguousarray(lo, np.float32); hi = np.ascontiguousarray(hi, np.float32)
    nc = lo.shape[0]
    cap = 4 * nc + 1024 if cap is None else cap
    g = Grid()
    bin_id = np.zeros(nc, np.int32); order = np.zeros(nc, np.int32); bin_start = np.zeros(cap + 1, np.int32)
    cc = np.zeros(nc, np.int32); cand = np.zeros((nc, K), np.int32)
    lib().oracle_cell_list(int(nd), nc, _p(lo, C.c_float), _p(hi, C.c_float), int(PBC), C.c_float(L), C.c_float(skin_rel),
                           C.c_float(rng), int(cap), int(K), int(far2d), C.byref(g), _p(bin_id, C.c_int32),
                           _p(order, C.c_int32), _p(bin_start, C.c_int32), _p(cc, C.c_int32), _p(cand, C.c_int32))
    return dict(grid=g, bin_id=bin_id, order=order, bin_start=bin_start[:g.nbins + 1].copy(), cand_count=cc, cand=cand)


# ---------------------------------------------------------------- geometry
def icosphere(subdiv=2):
    nv, nf = 10 * 4 ** subdiv + 2, 20 * 4 ** subdiv
    V = np.zeros((nv, 3), np.float32); F = np.zeros((nf, 3), np.uint32)
    lib().oracle_icosphere.restype = C.c_int
    n = lib().oracle_icosphere(int(subdiv), _p(V, C.c_float), _p(F, C.c_uint32))
    assert n == nv
    return V, F


def cell3d_params(calA, r0, nf):
    out = np.zeros(4, np.float32)
    lib().oracle_cell3d_params(C.c_float(calA), C.c_float(r0), int(nf), _p(out, C.c_float))
    return dict(v0=out[0], sa0=out[1], a0=out[2], l0=out[3])


def cell3d_place(unitV, r0, start):
    unitV = np.ascontiguousarray(unitV, np.float32)
    nv = unitV.shape[0]
    out = np.zeros((nv, 4), np.float32)
    st = np.asarray(start, np.float32)
    lib().oracle_cell3d_place(nv, _p(unitV, C.c_float), C.c_float(r0), _p(st, C.c_float), _p(out, C.c_float))
    return out


def cell2d_init(x0, y0, calA, NV, r0):
    verts = np.zeros((NV, 2), np.float32); out = np.zeros(3, np.float32)
    lib().oracle_cell2d_init(C.c_float(x0), C.c_float(y0), C.c_float(calA), int(NV), C.c_float(r0), _p(verts, C.c_float), _p(out, C.c_float))
    return verts, dict(calA0=out[0], a0=out[1], l0=out[2])
