"""GPU tests against the REAL reference: committed golden vectors (tests/golden, produced by the reference's own
OpenCL implementation on a B200) and, when NVIDIA's OpenCL is reachable on the box, the reference run live
through oracle/_ref.  The CUDA path is driven through the C ABI; for 3D the compat switch reproduces the
reference's volume race as it resolves on NVIDIA hardware (include/dpm_b200.h, dpm3d_set_compat)."""
import numpy as np
import pytest

from test_golden_cpu import load2d, load3d

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ref3d_test3dpy_16", "ref3d_test3dcpp_12"])
def test_cuda3d_vs_reference_golden(name):
    from opencl_dpm_b200 import Dpm3D

    g, nc, P = load3d(name)
    h = Dpm3D(nc, 162, g["faces"])
    h.set_compat(160)
    for n in g["steps"]:
        V = g["verts0"].copy()
        F = np.zeros_like(V)
        h.euler_update(V, *P, int(n), float(g["dt"]), float(g["Kre"]), 0.0, 1, float(g["L"]), forces_out=F)
        Fr, Vr = g[f"forces_{n}"], g[f"verts_{n}"]
        tol = 1e-5 * max(float(np.abs(Fr).max()), 1e-3)
        ef, ev = np.abs(F[:, :3] - Fr[:, :3]).max(), np.abs(V[:, :3] - Vr[:, :3]).max()
        print(f"{name} {n} steps: force err {ef:.2e} (tol {tol:.2e}), position err {ev:.2e}")
        assert ef <= tol and ev <= 4e-6
    h.close()


@pytest.mark.parametrize("name", ["ref2d_test2d_32", "ref2d_kat_24"])
def test_cuda2d_vs_reference_golden(name):
    from opencl_dpm_b200 import Dpm2D

    g, nc, P = load2d(name)
    S = g["verts0"].shape[1]
    h = Dpm2D(nc, S)
    for n in g["steps"]:
        V = g["verts0"].copy()
        F = np.zeros_like(V)
        h.euler_update(V, g["nv"], *P, int(n), float(g["dt"]), float(g["Kre"]), float(g["Kat"]), 1, float(g["L"]), forces_out=F)
        Fr, Vr = g[f"forces_{n}"], g[f"verts_{n}"]
        tol = 1e-5 * max(float(np.abs(Fr).max()), 1e-3)
        ef, ev = np.abs(F - Fr).max(), np.abs(V - Vr).max()
        print(f"{name} {n} steps: force err {ef:.2e} (tol {tol:.2e}), position err {ev:.2e}")
        if n == 1:
            assert ef <= tol and ev <= 5e-7
        else:
            assert ev <= 4e-6 and ef <= 1e-4
    h.close()


def test_cuda3d_vs_reference_live():
    """64-cell test3D.py configuration, reference run live on the box's OpenCL, 3 steps."""
    from oracle import ref as R

    if not R.available():
        pytest.skip("no OpenCL device / oracle/_ref on this machine")
    import helpers as H
    from opencl_dpm_b200 import Dpm3D

    d = H.config_test3d_py(64)
    Vr, Fr, sec = R.euler3d(d["verts"], d["Kv"], d["Ka"], d["Ks"], d["v0"], d["a0"], d["Kre"], d["PBC"], d["L"], 3, d["dt"])
    h = Dpm3D(d["nc"], d["nv"], d["faces"])
    h.set_compat(160)
    V = d["verts"].copy()
    F = np.zeros_like(V)
    ms = h.euler_update(V, *[d[k] for k in ("Kv", "Ka", "Ks", "v0", "a0", "l0")], 3, float(d["dt"]), float(d["Kre"]), 0.0, d["PBC"],
                        float(d["L"]), forces_out=F)
    tol = 1e-5 * max(float(np.abs(Fr).max()), 1e-3)
    ef, ev = np.abs(F[:, :3] - Fr[:, :3]).max(), np.abs(V[:, :3] - Vr[:, :3]).max()
    print(f"live reference ({R.device_name()}): {sec * 1e3:.0f} ms for 3 steps; CUDA loop {ms:.2f} ms; force err {ef:.2e} (tol {tol:.2e}), pos err {ev:.2e}")
    assert ef <= 2 * tol and ev <= 4e-6
    h.close()
