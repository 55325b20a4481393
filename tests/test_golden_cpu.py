"""CPU tests (-m "not gpu") that PIN the oracle: tests/golden/*.npz hold outputs of the REAL reference
(its unmodified host code + OpenCL kernels, run on a B200 through NVIDIA's OpenCL by tests/golden/make_golden.py).

3D: the reference's VolumeForceUpdate has a work-group race; on NVIDIA's runtime it resolves as
"faces 160..319 use the previous step's volume, zero on the first step".  The oracle reproduces the reference's
outputs with that emulation switched on (stale_from=160); its default is the intended semantics.
"""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from oracle import ref as R

GOLD = os.path.join(ROOT, "tests", "golden")


def _l0(a0):
    return np.float32(np.sqrt(np.float64(np.float32(4.0) * a0)) / np.sqrt(np.float64(np.float32(3.0))))


def load3d(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    nc = g["verts0"].shape[0] // 162
    one = np.ones(nc, np.float32)
    P = [one * g["Kv"], one * g["Ka"], one * g["Ks"], one * g["v0"], one * g["a0"], one * _l0(g["a0"])]
    return g, nc, P


def load2d(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    nc = g["verts0"].shape[0]
    one = np.ones(nc, np.float32)
    P = [one * g["Ka"], one * g["Kl"], one * g["Kb"], one * g["a0"], one * g["l0"], one * g["r0"]]
    return g, nc, P


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLD, "ref*.npz"))) >= 4


@pytest.mark.parametrize("name", ["ref3d_test3dpy_16", "ref3d_test3dcpp_12"])
def test_oracle3d_reproduces_reference_outputs(name):
    g, nc, P = load3d(name)
    for n in g["steps"]:
        V, F = O.run3d(g["verts0"], g["faces"], *P, g["Kre"], 1, g["L"], int(n), g["dt"], stale_from=160)
        Fr, Vr = g[f"forces_{n}"], g[f"verts_{n}"]
        tol = 1e-5 * max(float(np.abs(Fr).max()), 1e-3)
        assert np.abs(F[:, :3] - Fr[:, :3]).max() <= tol, (name, n)
        assert np.abs(V[:, :3] - Vr[:, :3]).max() <= 4e-6, (name, n)


def test_reference_race_is_what_separates_default_oracle_from_reference():
    """Default oracle (current volume for every face) differs from the reference's first step exactly by the
    volume force of faces 160..319 evaluated with volume 0 — nothing else."""
    g, nc, P = load3d("ref3d_test3dcpp_12")
    F = O.forces3d(g["verts0"], g["faces"], *P, g["Kre"], 1, g["L"])
    Fr = g["forces_1"]
    D = (Fr - F)[:, :3].reshape(nc, 162, 3).astype(np.float64)
    V = g["verts0"][:, :3].reshape(nc, 162, 3).astype(np.float64)
    faces = g["faces"]
    Kv, v0 = float(g["Kv"]), float(g["v0"])
    for c in (0, nc - 1):
        com = V[c].mean(0)
        vol = abs(sum(np.dot(np.cross(V[c, a], V[c, b]), V[c, d]) for a, b, d in faces)) / 6
        dstrain = (0.0 / v0 - 1.0) - (vol / v0 - 1.0)
        pred = np.zeros((162, 3))
        for f, (i0, i1, i2) in enumerate(faces):
            if f < 160:
                continue
            a, b, cc = V[c, i1] - com, V[c, i2] - com, V[c, i0] - com
            for idx, gg in ((i0, np.cross(a, b)), (i1, np.cross(b, cc)), (i2, np.cross(cc, a))):
                pred[idx] += -Kv * dstrain * gg / 6.0
        assert np.abs(pred - D[c]).max() < 2e-5
        assert np.abs(D[c]).max() > 0.1


@pytest.mark.parametrize("name", ["ref2d_test2d_32", "ref2d_kat_24"])
def test_oracle2d_reproduces_reference_outputs(name):
    g, nc, P = load2d(name)
    V, F = O.run2d(g["verts0"], g["nv"], *P, g["Kre"], g["Kat"], 1, g["L"], 1, g["dt"])
    Fr, Vr = g["forces_1"], g["verts_1"]
    tol = 1e-5 * max(float(np.abs(Fr).max()), 1e-3)
    assert np.abs(F - Fr).max() <= tol
    assert np.abs(V - Vr).max() <= 5e-7
    n = int(g["steps"][-1])
    V, F = O.run2d(g["verts0"], g["nv"], *P, g["Kre"], g["Kat"], 1, g["L"], n, g["dt"])
    assert np.abs(V - g[f"verts_{n}"]).max() <= 4e-6  # 20 steps: a few ulp of the coordinates
    assert np.abs(F - g[f"forces_{n}"]).max() <= 1e-4


# ---- the reference's own constructors and initialisers, compiled into oracle/_ref (CPU-only parts) ----------
needs_ref = pytest.mark.skipif(not R.present(), reason="oracle/_ref not built (needs /root/reference at build time)")


@needs_ref
def test_geometry_bit_exact_vs_reference_constructors():
    import helpers as H
    from opencl_dpm_b200 import capi

    for start, calA, r0 in (([0.0, 0.0, 0.0], 1.0, 1.0), ([7.0, 6.0, 1.3], 1.05, 1.8)):
        v, f, sc = R.cell3d(start, calA, r0)
        Vu, F = O.icosphere(2)
        assert np.array_equal(F, f)
        assert np.array_equal(O.cell3d_place(Vu, r0, start)[:, :3], v)
        p = O.cell3d_params(calA, r0, 320)
        assert (p["v0"], p["sa0"], p["a0"]) == (sc["v0"], sc["sa0"], sc["a0"])
        pc = capi.cell3d_params(calA, r0, 320)
        assert (pc["v0"], pc["sa0"], pc["a0"]) == (sc["v0"], sc["sa0"], sc["a0"])
        c = H.cldpm().Cell3D(start, calA, r0)
        assert np.array_equal(np.asarray(c.Verts, np.float32), v)
        assert np.float32(c.GetVolume()) == sc["Volume"]
    for calA, nv, r0 in ((1.05, 32, 1.0), (1.2, 64, 1.0), (1.2, 25, 1.0), (1.2, 22, 1.3)):
        v, sc = R.cell2d(0.5, -0.25, calA, nv, r0)
        vo, po = O.cell2d_init(0.5, -0.25, calA, nv, r0)
        assert np.array_equal(v, vo) and (po["calA0"], po["a0"], po["l0"]) == (sc["calA0"], sc["a0"], sc["l0"])
        c = H.cldpm().Cell2D(0.5, -0.25, calA, nv, r0)
        assert np.array_equal(np.asarray(c.Verts, np.float32), v)


@needs_ref
def test_disperse_bit_exact_vs_reference():
    """Tissue3D::Disperse2D and Tissue2D::Disperse of the host mirror produce the reference's initial conditions
    bit for bit (same unseeded drand48 stream, same float/double promotion)."""
    import helpers as H

    m = H.cldpm()
    vr, Lr = R.disperse3d(30, [0.0, 0.0, 0.0], 1.0, 1.0, 0.35)
    c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
    T = m.Tissue3D([c] * 30, 0.35)
    H.reset_drand48()
    T.Disperse2D()
    mine = np.concatenate([np.asarray(x.Verts, np.float32) for x in T.Cells])
    assert np.float32(T.L) == Lr and np.array_equal(mine, vr)
    vr2, Lr2 = R.disperse2d(32, 1.05, 32, 1.0, 0.85)
    c2 = m.Cell2D(0.0, 0.0, 1.05, 32, 1.0)
    T2 = m.Tissue2D([c2] * 32, 0.85)
    H.reset_drand48()
    T2.Disperse()
    mine2 = np.stack([np.asarray(x.Verts, np.float32) for x in T2.Cells])
    assert np.float32(T2.L) == Lr2 and np.array_equal(mine2, vr2)


@needs_ref
def test_binned_disperse_bit_exact_vs_reference_on_larger_tissues():
    """The host classes relax the cell centres over Verlet lists on a bin grid (opencl_dpm_b200/host/disperse.hpp, active
    from 64 cells) instead of the reference's all-pairs loop.  The REAL reference (oracle/_ref) on
    tissues large enough for the binned path: positions must still be equal bit for bit, i.e. same iteration count, same
    summation order."""
    import helpers as H

    m = H.cldpm()
    vr, Lr = R.disperse3d(96, [0.0, 0.0, 0.0], 1.0, 1.0, 0.35)
    c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
    T = m.Tissue3D([c] * 96, 0.35)
    H.reset_drand48()
    T.Disperse2D()
    mine = np.concatenate([np.asarray(x.Verts, np.float32) for x in T.Cells])
    assert np.float32(T.L) == Lr and np.array_equal(mine, vr)
    vr2, Lr2 = R.disperse2d(150, 1.05, 16, 1.0, 0.85)
    c2 = m.Cell2D(0.0, 0.0, 1.05, 16, 1.0)
    T2 = m.Tissue2D([c2] * 150, 0.85)
    H.reset_drand48()
    T2.Disperse()
    mine2 = np.stack([np.asarray(x.Verts, np.float32) for x in T2.Cells])
    assert np.float32(T2.L) == Lr2 and np.array_equal(mine2, vr2)
