"""Shared builders for the parity tests: the reference's own demo configurations
(test3D.py, test3D.cpp, test2D.cpp, test2D.py) expressed through the drop-in `clDPM` module,
then flattened to the C ABI's array layout."""
from __future__ import annotations

import numpy as np

import opencl_dpm_b200 as pkg


def cldpm():
    return pkg.load_cldpm()


def reset_drand48():
    """Disperse()/Disperse2D() draw from the UNSEEDED drand48 stream (src/Tissue3D.cpp:51-52); glibc's initial
    state is X0 = 0 (first values 3.9e-14, 9.85e-4, 0.0416, ... — SURVEY §3.3).  Re-arm it so that every
    configuration equals what a fresh reference process would produce, independent of test order."""
    import ctypes

    ctypes.CDLL(None).seed48((ctypes.c_ushort * 3)(0, 0, 0))


def flat3d(T):
    """Tissue3D -> dict of flat arrays (float4-strided vertices, per-cell scalars, faces)."""
    cells = T.Cells
    nc = len(cells)
    V = np.zeros((nc * 162, 4), np.float32)
    for i, c in enumerate(cells):
        V[i * 162:(i + 1) * 162, :3] = np.asarray(c.Verts, np.float32)
    faces = np.asarray(cells[0].GetFaces(), np.uint32)
    return dict(nc=nc, nv=162, faces=faces, verts=V, L=np.float32(T.L), PBC=int(T.PBC))


def params3d(nc, calA, r0, Kv, Ka, Ks, nf=320):
    from opencl_dpm_b200 import capi

    p = capi.cell3d_params(calA, r0, nf)
    one = np.ones(nc, np.float32)
    return dict(Kv=one * np.float32(Kv), Ka=one * np.float32(Ka), Ks=one * np.float32(Ks), v0=one * p["v0"],
                a0=one * p["a0"], l0=one * p["l0"])


def config_test3d_py(ncells=64):
    """reference test3D.py:8-16 (BASELINE config C uses 64 cells instead of 32)."""
    m = cldpm()
    c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
    c.Ka, c.Kv, c.Ks = 2.0, 5.0, 3.0
    T = m.Tissue3D([c] * ncells, 0.35)
    T.Kre = 25.0
    reset_drand48()
    T.Disperse2D()
    d = flat3d(T)
    d.update(params3d(ncells, 1.0, 1.0, 5.0, 2.0, 3.0))
    d.update(Kre=np.float32(25.0), dt=np.float32(0.01))
    return d


def config_test3d_cpp():
    """reference test3D.cpp:8-38: 30 cells, two prototypes, r0 1.8, z0 1.3, K=1, Kre 50, dt 0.005."""
    m = cldpm()
    a = m.Cell3D([7.0, 6.0, 1.3], 1.05, 1.8)
    b = m.Cell3D([4.0, 6.0, 1.3], 1.05, 1.8)
    for c in (a, b):
        c.Kv, c.Ka, c.Ks = 1.0, 1.0, 1.0
    T = m.Tissue3D([a, b] * 15, 0.35)
    T.Kre = 50.0
    reset_drand48()
    T.Disperse2D()
    d = flat3d(T)
    d.update(params3d(30, 1.05, 1.8, 1.0, 1.0, 1.0))
    d.update(Kre=np.float32(50.0), dt=np.float32(0.005))
    return d


def flat2d(T):
    cells = T.Cells
    nc = len(cells)
    nvs = np.array([len(c.Verts) for c in cells], np.int32)
    S = int(nvs.max())
    V = np.zeros((nc, S, 2), np.float32)
    for i, c in enumerate(cells):
        V[i, :nvs[i]] = np.asarray(c.Verts, np.float32)
    return dict(nc=nc, S=S, nv=nvs, verts=V, L=np.float32(T.L), PBC=int(T.PBC))


def params2d(specs, Ka, Kl, Kb):
    """specs: list of (calA, NV, r0) per cell."""
    from oracle import oracle as O

    a0, l0, r0 = [], [], []
    cache = {}
    for s in specs:
        if s not in cache:
            cache[s] = O.cell2d_init(0.0, 0.0, s[0], s[1], s[2])[1]
        a0.append(cache[s]["a0"]); l0.append(cache[s]["l0"]); r0.append(s[2])
    n = len(specs)
    one = np.ones(n, np.float32)
    return dict(Ka=one * np.float32(Ka), Kl=one * np.float32(Kl), Kb=one * np.float32(Kb), a0=np.array(a0, np.float32),
                l0=np.array(l0, np.float32), r0=np.array(r0, np.float32))


def config_test2d(ncells=32):
    """BASELINE config A: 32 x Cell2D(0,0,1.05,32,1.0), Ka=Kl=1, Kb=0.1, phi 0.85, Kre 50 (reference test2D.cpp:9-31)."""
    m = cldpm()
    c = m.Cell2D(0.0, 0.0, 1.05, 32, 1.0)
    c.Ka, c.Kl, c.Kb = 1.0, 1.0, 0.1
    T = m.Tissue2D([c] * ncells, 0.85)
    T.Kre = 50.0
    reset_drand48()
    T.Disperse()
    d = flat2d(T)
    d.update(params2d([(1.05, 32, 1.0)] * ncells, 1.0, 1.0, 0.1))
    d.update(Kre=np.float32(50.0), Kat=np.float32(0.0), dt=np.float32(0.005))
    return d


def config_test2d_py(npairs=40):
    """reference test2D.py:8-21 (mixed 25/22-vertex cells, Kat != 0), shrunk from 200 pairs."""
    m = cldpm()
    c = m.Cell2D(0.0, 0.0, 1.2, 25, 1.0)
    c2 = m.Cell2D(0.0, 0.0, 1.2, 22, 1.3)
    for x in (c, c2):
        x.Ka, x.Kl, x.Kb = 0.1, 1.0, 0.05
    T = m.Tissue2D([c, c2] * npairs, 0.9)
    T.Kre = 1.0
    T.Kat = 0.5
    reset_drand48()
    T.Disperse()
    d = flat2d(T)
    d.update(params2d([(1.2, 25, 1.0), (1.2, 22, 1.3)] * npairs, 0.1, 1.0, 0.05))
    d.update(Kre=np.float32(1.0), Kat=np.float32(0.5), dt=np.float32(0.005))
    return d


def assert_forces_close(F, Fref32, Fref64, what="forces", cond_factor=8.0, max_illcond_frac=2e-3):
    """Per-step force parity (SURVEY §8c iii): every vertex within 1e-5 * max(|F_ref|_inf, 1e-3) of the fp32 oracle.
    The reference's repulsion is ill-conditioned for a vertex lying almost in the plane of a neighbour's face that
    subtends ~pi (2*atan2(num, den) with num, den -> 0): there the fp32 oracle itself is off the fp64 oracle by more
    than the tolerance.  Such vertices must stay within cond_factor x the fp32 oracle's own error (+ tol), and must
    be rare."""
    tol = force_tol(Fref32)
    err = np.abs(F - Fref32).reshape(len(F), -1).max(1)
    own = np.abs(Fref32.astype(np.float64) - Fref64).reshape(len(F), -1).max(1)
    bad = err > tol
    assert (err[bad] <= cond_factor * own[bad] + tol).all(), (
        f"{what}: max error {err.max():.3e} (tol {tol:.3e}); worst ill-conditioned ratio "
        f"{(err[bad] / (own[bad] + 1e-30)).max():.1f}")
    assert bad.mean() <= max_illcond_frac, f"{what}: {bad.sum()} of {len(bad)} vertices beyond tolerance"
    return float(err.max() / tol), int(bad.sum())


def force_tol(Fref, rel=1e-5):
    """max_i |dF|_inf <= rel * max(|F_ref|_inf over the tissue, 1e-3)."""
    return rel * max(float(np.abs(Fref).max()), 1e-3)
