"""GPU parity tests for the 2D hot path (pytest -m gpu), through the C ABI.
Tolerance: SURVEY.md §8(c): per step max |dF|_inf <= 1e-5 * max(|F_ref|_inf, 1e-3);
the inside/outside classification and the cell-list artefacts are bit-exact."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
PK = ("Ka", "Kl", "Kb", "a0", "l0", "r0")


def _oracle():
    from oracle import oracle as O

    return O


def _handle(d, **kw):
    from opencl_dpm_b200 import Dpm2D

    h = Dpm2D(d["nc"], d["S"])
    if kw:
        h.set_neighbor_params(**kw)
    return h


def _gpu(h, d, verts, nsteps=1, mask=31, Kat=None):
    h.set_force_mask(mask)
    h.upload(verts, d["nv"], *[d[k] for k in PK])
    h.step(nsteps, float(d["dt"]), float(d["Kre"]), float(d["Kat"] if Kat is None else Kat), d["PBC"], float(d["L"]))
    return h.download()


def _real(d, A):
    """mask of real (non-padding) vertex slots"""
    return (np.arange(d["S"])[None, :] < d["nv"][:, None])


@pytest.mark.parametrize("cfg", ["test2d", "test2d_py"])
@pytest.mark.parametrize("mask", [1, 2, 4, 8, 16, 31])
def test_single_step_force_parity(cfg, mask):
    O = _oracle()
    d = H.config_test2d(32) if cfg == "test2d" else H.config_test2d_py(40)
    h = _handle(d)
    V1, F = _gpu(h, d, d["verts"], 1, mask)
    Fref = O.forces2d(d["verts"], d["nv"], *[d[k] for k in PK], d["Kre"], d["Kat"], d["PBC"], d["L"], which=mask)
    m = _real(d, F)
    tol = H.force_tol(Fref)
    err = np.abs(F - Fref)[m].max()
    assert err <= tol, f"force error {err:.3e} > {tol:.3e} (|F|max {np.abs(Fref).max():.3f})"
    Vref = d["verts"] + Fref * d["dt"]
    assert np.abs(V1 - Vref)[m].max() <= tol * float(d["dt"]) + 4e-7 * np.abs(Vref).max()
    h.close()


@pytest.mark.parametrize("cfg", ["test2d", "test2d_py"])
def test_inside_classification_bit_exact(cfg):
    """The repulsion's even-odd point-in-polygon result per vertex (incl. the |d|>L wrap quirk) must equal
    the all-pairs oracle's exactly: a vertex carries repulsion force iff the oracle says inside."""
    O = _oracle()
    d = H.config_test2d(32) if cfg == "test2d" else H.config_test2d_py(40)
    h = _handle(d)
    _, F = _gpu(h, d, d["verts"], 1, 16)
    _, inside = O.forces2d(d["verts"], d["nv"], *[d[k] for k in PK], d["Kre"], d["Kat"], d["PBC"], d["L"], which=16, want_inside=True)
    m = _real(d, F)
    gpu_in = (np.abs(F).max(2) > 0)
    assert inside[m].sum() > 5, "config has too few overlapping vertices to be a meaningful test"
    assert np.array_equal(gpu_in[m], inside[m].astype(bool))
    h.close()


def test_multi_step_parity_each_step_from_oracle_state():
    O = _oracle()
    d = H.config_test2d_py(40)
    h = _handle(d)
    V = d["verts"].copy()
    m = _real(d, V)
    worst = 0.0
    for s in range(25):
        _, F = _gpu(h, d, V)
        Fref = O.forces2d(V, d["nv"], *[d[k] for k in PK], d["Kre"], d["Kat"], d["PBC"], d["L"])
        worst = max(worst, np.abs(F - Fref)[m].max() / H.force_tol(Fref))
        V = (V + Fref * d["dt"]).astype(np.float32)
    assert worst <= 1.0, f"worst per-step force error is {worst:.2f}x the tolerance"
    h.close()


def test_trajectory_500_steps_vs_oracle():
    """reference test2D.py inner loop (500 steps, dt 0.005) on 80 cells."""
    O = _oracle()
    d = H.config_test2d_py(40)
    h = _handle(d)
    V1, F1 = _gpu(h, d, d["verts"], 500)
    Vr, Fr = O.run2d(d["verts"], d["nv"], *[d[k] for k in PK], d["Kre"], d["Kat"], d["PBC"], d["L"], 500, d["dt"])
    V64, _ = O.run2d(d["verts"], d["nv"], *[d[k] for k in PK], d["Kre"], d["Kat"], d["PBC"], d["L"], 500, d["dt"], dtype=np.float64)
    m = _real(d, V1)
    scale = np.abs(Vr[m]).max()
    g = np.abs(V1 - Vr)[m].max() / scale
    o = np.abs(Vr - V64)[m].max() / scale
    print(f"drift over 500 steps: gpu-vs-f32 {g:.3e}  f32-vs-f64 {o:.3e}")
    assert g <= max(1e-4, 20 * o)
    h.close()


@pytest.mark.parametrize("cfg,kat", [("test2d", 0.0), ("test2d_py", 0.5), ("test2d_py", 0.0)])
def test_neighbor_artifacts_bit_exact(cfg, kat):
    O = _oracle()
    d = H.config_test2d(32) if cfg == "test2d" else H.config_test2d_py(40)
    h = _handle(d)
    h.upload(d["verts"], d["nv"], *[d[k] for k in PK])
    h.rebuild_neighbors(kat, d["PBC"], float(d["L"]))
    art = h.neighbor_artifacts()
    lo, hi = O.aabb2d(d["verts"], d["nv"])
    rng = float(d["l0"].max()) if kat != 0.0 else 0.0
    ref = O.cell_list(2, lo, hi, d["PBC"], d["L"], 0.1, rng, h.K, far2d=True)
    assert art["grid"].as_tuple() == ref["grid"].as_tuple()
    for k in ("bin_id", "order", "bin_start", "cand_count"):
        assert np.array_equal(art[k], ref[k]), k
    for i in range(d["nc"]):
        n = ref["cand_count"][i]
        assert np.array_equal(art["cand"][i, :n], ref["cand"][i, :n]), i
    h.close()


def test_culled_oracle_equals_all_pairs_on_gpu_lists():
    """The GPU's candidate lists fed to the CPU oracle's culled form reproduce the all-pairs forces exactly
    (the lists are a superset of every interacting pair, including the |d|>L wrap partners)."""
    O = _oracle()
    d = H.config_test2d(32)
    h = _handle(d)
    h.upload(d["verts"], d["nv"], *[d[k] for k in PK])
    h.rebuild_neighbors(0.5, d["PBC"], float(d["L"]))
    art = h.neighbor_artifacts()
    Fa = O.forces2d(d["verts"], d["nv"], *[d[k] for k in PK], d["Kre"], 0.5, d["PBC"], d["L"])
    Fc = O.forces2d(d["verts"], d["nv"], *[d[k] for k in PK], d["Kre"], 0.5, d["PBC"], d["L"], cand_count=art["cand_count"], cand=art["cand"])
    assert np.array_equal(Fa, Fc)
    h.close()


def test_euler_update_and_skin_independence():
    d = H.config_test2d_py(40)
    h = _handle(d)
    V = d["verts"].copy()
    F = np.zeros_like(V)
    h.euler_update(V, d["nv"], *[d[k] for k in PK], 300, float(d["dt"]), float(d["Kre"]), float(d["Kat"]), d["PBC"], float(d["L"]), forces_out=F)
    V1, F1 = _gpu(h, d, d["verts"], 300)
    m = _real(d, V)
    assert np.array_equal(V[m], V1[m]) and np.array_equal(F[m], F1[m])
    h2 = _handle(d, skin_rel=0.01, max_candidates=32)
    V2, F2 = _gpu(h2, d, d["verts"], 300)
    print(f"rebuilds: skin 0.1 -> {h.stats().rebuilds}, skin 0.01 -> {h2.stats().rebuilds}")
    assert h2.stats().rebuilds > h.stats().rebuilds
    assert np.array_equal(V1[m], V2[m]) and np.array_equal(F1[m], F2[m])
    h.close(); h2.close()


def test_coincident_vertices_of_two_cells_stay_finite():
    """Attraction pulls vertices of neighbouring cells together and in fp32 they can meet exactly (seen after ~530 steps
    of config B): OpenCL normalize(0) = 0, so the pair adds nothing; the GPU must equal the oracle there, not NaN."""
    O = _oracle()
    d = H.config_test2d(32)
    V = d["verts"].copy()
    # move cell 1 so that one of its vertices lands exactly on a vertex of cell 0
    shift = V[0, 3] - V[1, 11]
    V[1, : d["nv"][1]] += shift
    V[1, 11] = V[0, 3]
    h = _handle(d)
    for mask in (8, 31):
        V1, F = _gpu(h, d, V, 1, mask, Kat=0.5)
        Fref = O.forces2d(V, d["nv"], *[d[k] for k in PK], d["Kre"], 0.5, d["PBC"], d["L"], which=mask)
        m = _real(d, F)
        assert np.isfinite(Fref[m]).all() and np.isfinite(F[m]).all()
        assert np.abs(F - Fref)[m].max() <= H.force_tol(Fref)
    h.close()


def test_step_resident_equals_repeated_cleulerupdate_2d():
    """Tissue2D.StepResident / SyncCells (extension): same trajectory as repeated CLEulerUpdate calls, bit for bit."""
    m = H.cldpm()

    def tissue():
        c = m.Cell2D(0.0, 0.0, 1.2, 25, 1.0)
        c2 = m.Cell2D(0.0, 0.0, 1.2, 22, 1.3)
        for x in (c, c2):
            x.Ka, x.Kl, x.Kb = 0.1, 1.0, 0.05
        T = m.Tissue2D([c, c2] * 12, 0.9)
        T.Kre = 1.0
        T.Kat = 0.5
        H.reset_drand48()
        T.Disperse()
        return T

    A, B = tissue(), tissue()
    for _ in range(3):
        A.CLEulerUpdate(20, 0.005)
        B.StepResident(20, 0.005)
    B.SyncCells()
    assert np.array_equal(H.flat2d(A)["verts"], H.flat2d(B)["verts"])


@pytest.mark.parametrize("nx,pbc", [(2, 1), (3, 1), (4, 0), (6, 1)])
def test_attraction_and_repulsion_in_small_and_open_boxes(nx, pbc):
    """The attraction's near-vertex compaction and the per-vertex culls are only applied when the periodic box is large
    against the cells (`att_cull_ok` / `own_cull_ok`, dpm2d.cu); 2 x 2 and 3 x 3 cells in a box of two / three lattice
    spacings take the no-cull paths (every image is within reach), an open box has no images at all.  All against the
    all-pairs oracle, 5 steps, each re-seeded from the oracle state."""
    from opencl_dpm_b200 import synth

    O = _oracle()
    d = synth.tissue2d(nx, nv=32, Kat=0.7)
    d["PBC"] = pbc
    h = _handle(d, skin_rel=0.1, max_candidates=64)
    V = d["verts"].copy()
    worst = 0.0
    for s in range(5):
        V1, F = _gpu(h, d, V)
        Fr = O.forces2d(V, d["nv"], *[d[k] for k in PK], d["Kre"], d["Kat"], d["PBC"], d["L"])
        tol = H.force_tol(Fr)
        err = float(np.abs(F - Fr).max())
        worst = max(worst, err / tol)
        assert err <= tol, f"step {s}: force error {err:.3e} > {tol:.3e}"
        V = V + Fr * d["dt"]
    Fatt = O.forces2d(d["verts"], d["nv"], *[d[k] for k in PK], d["Kre"], d["Kat"], d["PBC"], d["L"], which=8)
    assert np.abs(Fatt).max() > 1e-3  # the attraction is active in this fixture
    print(f"nx={nx} pbc={pbc}: worst error/tol {worst:.2f}")
    h.close()
