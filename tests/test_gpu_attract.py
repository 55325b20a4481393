"""GPU parity tests of the 3D vertex-vertex attraction (AllVertAttraction, shaders/Cell3D_Kernel.cl:313-364; SURVEY §8f
rank 1).  The reference compiles this kernel but never enqueues it, so it is opt-in here (DPM3D_ATTRACT); with the default
mask Kat must have no effect, exactly as in the reference.  Checker: the oracle's literal scatter-form restatement, and the
golden vector produced by the reference's own kernel text on NVIDIA OpenCL (tests/golden/ref3d_attract_12.npz)."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

PKEYS = ("Kv", "Ka", "Ks", "v0", "a0", "l0")
ATTRACT = 16
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref3d_attract_12.npz")


def _oracle():
    from oracle import oracle as O

    return O


def _tissue(nx, ny=None, subdiv=2, l0_scale=None, **kw):
    from opencl_dpm_b200 import synth

    d = synth.monolayer3d(nx, ny, subdiv=subdiv, **kw)
    if l0_scale is not None:  # heterogeneous rest lengths: the cutoff 2*l0 and the force use EACH cell's own l0 (:347-350)
        d["l0"] = (d["l0"] * np.where(np.arange(d["nc"]) % 2 == 0, 1.0, l0_scale)).astype(np.float32)
    return d


def _gpu_forces(d, mask, Kat, pbc, nsteps=1, L=None):
    from opencl_dpm_b200 import Dpm3D

    h = Dpm3D(d["nc"], d["nv"], d["faces"])
    h.set_force_mask(mask)
    h.upload(d["verts"], *[d[k] for k in PKEYS])
    h.step(nsteps, float(d["dt"]), float(d["Kre"]), float(Kat), pbc, float(d["L"] if L is None else L))
    V, F = h.download()
    st = h.stats()
    h.close()
    return V, F, st


@pytest.mark.parametrize("pbc", [0, 1])
@pytest.mark.parametrize("subdiv,l0_scale", [(2, None), (2, 1.3), (3, 1.2)])
def test_attraction_alone_vs_oracle(pbc, subdiv, l0_scale):
    O = _oracle()
    d = _tissue(4, subdiv=subdiv, l0_scale=l0_scale)
    Kat = 0.7
    V1, F, st = _gpu_forces(d, ATTRACT, Kat, pbc)
    Fr = O.attract3d(d["verts"], d["l0"], Kat, pbc, float(d["L"]), d["nc"])
    assert np.abs(Fr).max() > 0.05 and (np.abs(Fr[:, :3]).max(1) > 0).sum() > 50  # the fixture has attracting pairs
    tol = H.force_tol(Fr)
    err = np.abs(F[:, :3] - Fr[:, :3]).max()
    assert err <= tol, f"attraction force error {err:.3e} > {tol:.3e}"
    assert np.array_equal(F[:, 3], np.zeros(len(F), np.float32))


def test_attraction_periodic_images_in_a_tiny_box():
    """2 x 2 cells in a box of two lattice spacings: a neighbour is within reach through BOTH faces of the box, so the
    per-pair minimum image (:338-343) differs from the per-cell COM shift of the contact term; the product keeps every
    vertex for such pairs and evaluates the literal formula."""
    O = _oracle()
    d = _tissue(2, subdiv=2)
    Kat = 0.9
    _, F, _ = _gpu_forces(d, ATTRACT, Kat, 1)
    Fr = O.attract3d(d["verts"], d["l0"], Kat, 1, float(d["L"]), d["nc"])
    Fn = O.attract3d(d["verts"], d["l0"], Kat, 0, float(d["L"]), d["nc"])
    assert np.abs(Fr - Fn).max() > 1e-2  # the periodic images matter here
    assert np.abs(F[:, :3] - Fr[:, :3]).max() <= H.force_tol(Fr)


def test_all_forces_with_attraction_five_steps():
    """The six live kernels + attraction, 5 steps, each step re-seeded from the oracle state."""
    O = _oracle()
    d = _tissue(4, subdiv=2, l0_scale=1.25)
    Kat = 0.5
    from opencl_dpm_b200 import Dpm3D

    h = Dpm3D(d["nc"], d["nv"], d["faces"])
    h.set_force_mask(15 | ATTRACT)
    V = d["verts"].copy()
    for s in range(5):
        h.upload(V, *[d[k] for k in PKEYS])
        h.step(1, float(d["dt"]), float(d["Kre"]), Kat, d["PBC"], float(d["L"]))
        V1, F = h.download()
        args = (V, d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"])
        Fr = O.attract3d(V, d["l0"], Kat, d["PBC"], float(d["L"]), d["nc"], forces=O.forces3d(*args))
        F64 = O.attract3d(V, d["l0"], Kat, d["PBC"], float(d["L"]), d["nc"], forces=O.forces3d(*args, dtype=np.float64), dtype=np.float64)
        H.assert_forces_close(F[:, :3], Fr[:, :3], F64[:, :3], f"step {s}")
        Vn = V.copy()
        Vn[:, :3] += Fr[:, :3] * d["dt"]
        assert np.abs(V1[:, :3] - Vn[:, :3]).max() <= 2e-6
        V = Vn
    h.close()


def test_multi_step_device_resident_with_attraction():
    """25 steps without re-seeding (neighbour lists, skin and rebuilds live on the device) vs the oracle trajectory."""
    O = _oracle()
    d = _tissue(4, subdiv=2)
    Kat = 0.5
    V, F, st = _gpu_forces(d, 15 | ATTRACT, Kat, d["PBC"], nsteps=25)
    Vr, Fr = O.run3d_attract(d["verts"], d["faces"], *[d[k] for k in PKEYS], d["Kre"], Kat, d["PBC"], d["L"], 25, d["dt"])
    assert np.abs(V[:, :3] - Vr[:, :3]).max() <= 2e-5
    assert np.abs(F[:, :3] - Fr[:, :3]).max() <= 20 * H.force_tol(Fr)  # 25 steps of fp32 drift


def test_default_mask_ignores_kat_like_the_reference():
    """The reference host never enqueues AllVertAttraction: Tissue3D.Kat has no effect.  Same here unless opted in."""
    d = _tissue(4, subdiv=2)
    Va, Fa, _ = _gpu_forces(d, 15, 0.0, d["PBC"], nsteps=3)
    Vb, Fb, _ = _gpu_forces(d, 15, 5.0, d["PBC"], nsteps=3)
    assert np.array_equal(Va, Vb) and np.array_equal(Fa, Fb)
    # and with the attraction selected but Kat == 0 the kernel returns at once (:318-319): bit-identical again
    Vc, Fc, _ = _gpu_forces(d, 15 | ATTRACT, 0.0, d["PBC"], nsteps=3)
    assert np.array_equal(Va, Vc) and np.array_equal(Fa, Fc)


def test_contact_term_unchanged_by_the_wider_attraction_cull():
    """With the attraction on, the units kernel admits vertices within the larger attraction pad; the contact term of
    those must still be exactly what the default path computes: F(15|16) - F(16) == F(15) up to one rounding of the sum."""
    d = _tissue(4, subdiv=2)
    _, F15, _ = _gpu_forces(d, 15, 0.0, d["PBC"])
    _, F16, _ = _gpu_forces(d, ATTRACT, 0.7, d["PBC"])
    _, F31, _ = _gpu_forces(d, 15 | ATTRACT, 0.7, d["PBC"])
    assert np.abs(F31 - (F15 + F16)).max() <= 4e-6 * max(1.0, float(np.abs(F31).max()))


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vector of the reference's AllVertAttraction kernel not generated yet")
@pytest.mark.parametrize("pbc", [0, 1])
def test_cuda_attraction_vs_reference_kernel_golden(pbc):
    from opencl_dpm_b200 import Dpm3D, capi

    g = np.load(GOLD)
    nc = len(g["l0"])
    p = capi.cell3d_params(1.0, 1.0, 320)
    one = np.ones(nc, np.float32)
    h = Dpm3D(nc, 162, g["faces"])
    h.set_force_mask(ATTRACT)
    h.upload(g["verts0"], one, one, one, one * p["v0"], one * p["a0"], g["l0"])
    h.step(1, 0.01, 0.0, float(g["Kat"]), pbc, float(g["L"]))
    _, F = h.download()
    h.close()
    Fr = g[f"forces_pbc{pbc}"]
    tol = 1e-5 * max(float(np.abs(Fr).max()), 1e-3)
    err = np.abs(F[:, :3] - Fr[:, :3]).max()
    print(f"AllVertAttraction vs the reference's kernel (pbc={pbc}): err {err:.2e}, tol {tol:.2e}, |F|max {np.abs(Fr).max():.3f}")
    assert err <= tol


def test_attraction_live_reference_kernel():
    """The reference's kernel text run live on the box's OpenCL (when reachable) on the synthetic monolayer."""
    from oracle import ref as R

    if not R.available():
        pytest.skip("no OpenCL device / oracle/_ref on this machine")
    d = _tissue(4, subdiv=2, l0_scale=1.3)
    Kat = 0.7
    Fr = R.attract3d(d["verts"], d["l0"], Kat, 1, float(d["L"]))
    _, F, _ = _gpu_forces(d, ATTRACT, Kat, 1)
    assert np.abs(Fr).max() > 0.05
    assert np.abs(F[:, :3] - Fr[:, :3]).max() <= H.force_tol(Fr)


def test_cldpm_tissue3d_opt_in_through_attraction_method():
    """Drop-in surface: Tissue3D.Kat is inert by default (reference behaviour); attractionMethod = "AllVertAttraction"
    enables the kernel, and the result matches the oracle with the attraction enqueued."""
    O = _oracle()
    m = H.cldpm()

    def run(method, Kat):
        c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
        c.Ka, c.Kv, c.Ks = 2.0, 5.0, 3.0
        T = m.Tissue3D([c] * 16, 0.35)
        T.Kre = 25.0
        T.Kat = Kat
        if method:
            T.attractionMethod = method
        H.reset_drand48()
        T.Disperse2D()
        d0 = H.flat3d(T)
        T.CLEulerUpdate(2, 0.01)
        return d0, H.flat3d(T)["verts"]

    d0, Va = run(None, 0.0)
    _, Vb = run(None, 5.0)
    assert np.array_equal(Va, Vb)
    _, Vc = run("AllVertAttraction", 5.0)
    P = H.params3d(16, 1.0, 1.0, 5.0, 2.0, 3.0)
    Vr, _ = O.run3d_attract(d0["verts"], d0["faces"], *[P[k] for k in PKEYS], 25.0, 5.0, d0["PBC"], d0["L"], 2, np.float32(0.01))
    assert np.abs(Vc[:, :3] - Vr[:, :3]).max() <= 4e-6
    print("attraction moved vertices by up to", np.abs(Vc - Va).max())
