"""Synthetic tissues for BASELINE.json's configurations, as flat arrays in the
C ABI's layout.  The reference's own initialisers (Disperse / Disperse2D) are O(N^2) x 1e5
iterations and unusable beyond ~1e3 cells, so the large configs use a jittered square lattice
that is commensurate with the periodic box."""
from __future__ import annotations

import numpy as np



def monolayer3d(nx: int, ny: int | None = None, subdiv: int = 3, r0: float = 1.0, calA: float = 1.0,
                spacing: float = 1.9, jitter: float = 0.05, seed: int = 12345, Kv: float = 5.0, Ka: float = 2.0,
                Ks: float = 3.0, Kre: float = 25.0, dt: float = 0.01, x_range: tuple[int, int] | None = None,
                y_range: tuple[int, int] | None = None, geom=None):
    """Configs D (64x64) and E (512x512): nx*ny icosphere cells on a square lattice resting on the
    substrate (centre z = r0), spacing 1.9 r0 (~5 % overlap), jitter U(-0.05,0.05) r0, periodic box
    L = nx*spacing.  x_range=(i0,i1) builds only lattice columns i0..i1-1 (a slab, for sharding);
    the jitter stream is indexed by global cell id so slabs agree with the full tissue.  y_range likewise restricts the
    rows (a bounded block of a large tissue for the CPU arm).  geom: the module that provides icosphere() and
    cell3d_params() — the product's C ABI by default; bench.py's reference arm passes the oracle so that the CPU arm never
    maps the product library.  The jitter uses numpy's RandomState(12345) (SURVEY §8d names std::mt19937(12345): the same
    Mersenne twister, numpy's own float conversion)."""
    ny = nx if ny is None else ny
    if geom is None:
        from . import capi as geom
    unit, faces = geom.icosphere(subdiv)
    nv, nf = unit.shape[0], faces.shape[0]
    p = geom.cell3d_params(calA, r0, nf)
    s = np.float32(spacing * r0)
    rng = np.random.RandomState(seed)
    jit = ((rng.random_sample((nx * ny, 2)) * 2.0 - 1.0) * jitter * r0).astype(np.float32)
    i0, i1 = (0, nx) if x_range is None else x_range
    j0, j1 = (0, ny) if y_range is None else y_range
    ii, jj = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), indexing="ij")
    gid = (ii * ny + jj).ravel()
    cx = (ii.ravel().astype(np.float32) + np.float32(0.5)) * s + jit[gid, 0]
    cy = (jj.ravel().astype(np.float32) + np.float32(0.5)) * s + jit[gid, 1]
    cz = np.full_like(cx, np.float32(r0))
    nc = cx.shape[0]
    verts = np.zeros((nc, nv, 4), np.float32)
    verts[:, :, :3] = unit[None, :, :] * np.float32(r0) + np.stack([cx, cy, cz], 1)[:, None, :]
    one = np.ones(nc, np.float32)
    return dict(nc=nc, nv=nv, nf=nf, faces=faces, verts=verts.reshape(nc * nv, 4), gid=gid,
                Kv=one * np.float32(Kv), Ka=one * np.float32(Ka), Ks=one * np.float32(Ks), v0=one * p["v0"],
                a0=one * p["a0"], l0=one * p["l0"], Kre=np.float32(Kre), dt=np.float32(dt), PBC=1,
                L=np.float32(nx) * s, Ly=np.float32(ny) * s)


def tissue2d(nx: int, ny: int | None = None, nv: int = 64, r0: float = 1.0, calA: float = 1.2, spacing: float = 1.9,
             jitter: float = 0.05, seed: int = 12345, Ka: float = 0.1, Kl: float = 1.0, Kb: float = 0.05, Kre: float = 1.0,
             Kat: float = 0.5, dt: float = 0.005):
    """Config B: nx*ny regular nv-gons (reference Cell2D construction, src/cell.cpp:12-33) on a jittered
    square lattice, periodic box L = nx*spacing; stiffnesses of reference test2D.py:11-21."""
    ny = nx if ny is None else ny
    nc = nx * ny
    NV = np.float32(nv)
    calA0 = np.float32(calA * (nv * np.tan(np.pi / nv) / np.pi))
    ang = 2.0 * np.pi * (np.arange(nv) + 1.0) / float(NV)
    ring = np.stack([np.float32(r0) * np.cos(ang), np.float32(r0) * np.sin(ang)], 1).astype(np.float32)
    # a0 = shoelace area of the ring at the origin, accumulated like Cell2D::GetArea (float accumulator)
    area = np.float32(0.0)
    j = nv - 1
    for i in range(nv):
        area = np.float32(np.float64(area) + 0.5 * np.float64(np.float32(ring[j, 0] + ring[i, 0]) * np.float32(ring[j, 1] - ring[i, 1])))
        j = i
    a0 = np.float32(abs(area))
    l0 = np.float32(2.0 * np.sqrt(np.pi * np.float64(calA0) * np.float64(a0)) / float(NV))
    s = np.float32(spacing * r0)
    rng = np.random.RandomState(seed)
    jit = ((rng.random_sample((nc, 2)) * 2.0 - 1.0) * jitter * r0).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    cx = (ii.ravel().astype(np.float32) + np.float32(0.5)) * s + jit[:, 0]
    cy = (jj.ravel().astype(np.float32) + np.float32(0.5)) * s + jit[:, 1]
    verts = (ring[None, :, :] + np.stack([cx, cy], 1)[:, None, :]).astype(np.float32)
    one = np.ones(nc, np.float32)
    return dict(nc=nc, S=nv, nv=np.full(nc, nv, np.int32), verts=verts, Ka=one * np.float32(Ka), Kl=one * np.float32(Kl),
                Kb=one * np.float32(Kb), a0=one * a0, l0=one * l0, r0=one * np.float32(r0), Kre=np.float32(Kre),
                Kat=np.float32(Kat), dt=np.float32(dt), PBC=1, L=np.float32(nx) * s)


def _reset_drand48():
    """Disperse()/Disperse2D() draw from the UNSEEDED drand48 stream (src/Tissue3D.cpp:51-52); re-arm glibc's initial state
    so that the configuration equals what a fresh reference process produces."""
    import ctypes

    ctypes.CDLL(None).seed48((ctypes.c_ushort * 3)(0, 0, 0))


def test3d_config(ncells: int = 64, subdiv: int = 2):
    """BASELINE config C: ncells x Cell3D({0,0,0}, 1.0, 1.0), Ka=2, Kv=5, Ks=3, Tissue3D(cells, 0.35), Kre=25,
    Disperse2D() — reference test3D.py:8-16 — through the drop-in clDPM classes, flattened to the C ABI's layout."""
    from . import capi, load_cldpm

    m = load_cldpm()
    c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0, subdiv)
    c.Ka, c.Kv, c.Ks = 2.0, 5.0, 3.0
    T = m.Tissue3D([c] * ncells, 0.35)
    T.Kre = 25.0
    _reset_drand48()
    T.Disperse2D()
    cells = T.Cells
    nv, nf = cells[0].NV, cells[0].NF
    V = np.zeros((ncells, nv, 4), np.float32)
    for i, x in enumerate(cells):
        V[i, :, :3] = np.asarray(x.Verts, np.float32)
    p = capi.cell3d_params(1.0, 1.0, nf)
    one = np.ones(ncells, np.float32)
    return dict(nc=ncells, nv=nv, nf=nf, faces=np.asarray(cells[0].GetFaces(), np.uint32), verts=V.reshape(ncells * nv, 4),
                gid=np.arange(ncells), Kv=one * np.float32(5.0), Ka=one * np.float32(2.0), Ks=one * np.float32(3.0), v0=one * p["v0"],
                a0=one * p["a0"], l0=one * p["l0"], Kre=np.float32(25.0), dt=np.float32(0.01), PBC=int(T.PBC), L=np.float32(T.L))


def test2d_config(ncells: int = 32):
    """BASELINE config A: ncells x Cell2D(0,0,1.05,32,1.0), Ka=Kl=1, Kb=0.1, Tissue2D(cells, 0.85), Kre=50, Disperse() —
    reference test2D.cpp:9-31 — through the drop-in clDPM classes (the cells' own a0 / l0 as the host pack reads them)."""
    from . import load_cldpm

    m = load_cldpm()
    c = m.Cell2D(0.0, 0.0, 1.05, 32, 1.0)
    c.Ka, c.Kl, c.Kb = 1.0, 1.0, 0.1
    T = m.Tissue2D([c] * ncells, 0.85)
    T.Kre = 50.0
    _reset_drand48()
    T.Disperse()
    cells = T.Cells
    nv = 32
    V = np.zeros((ncells, nv, 2), np.float32)
    for i, x in enumerate(cells):
        V[i] = np.asarray(x.Verts, np.float32)
    ref = tissue2d(1, 1, nv=nv, calA=1.05)  # a0, l0 of the regular 32-gon (Cell2D constructor arithmetic)
    one = np.ones(ncells, np.float32)
    return dict(nc=ncells, S=nv, nv=np.full(ncells, nv, np.int32), verts=V, Ka=one, Kl=one.copy(), Kb=one * np.float32(0.1),
                a0=one * ref["a0"][0], l0=one * ref["l0"][0], r0=one.copy(), Kre=np.float32(50.0), Kat=np.float32(0.0),
                dt=np.float32(0.005), PBC=int(T.PBC), L=np.float32(T.L))
