"""GPU parity tests for the 3D hot path (run on the B200 box: pytest -m gpu).

Every compute call goes through the C ABI (libdpm_b200.so via ctypes); the CPU oracle
(oracle/) is only the checker.  Tolerances follow SURVEY.md §8(c):
  integer artefacts bit-exact; forces/positions per step within
  1e-5 * max(|F_ref|_inf over the tissue, 1e-3).
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

PKEYS = ("Kv", "Ka", "Ks", "v0", "a0", "l0")


def _oracle():
    from oracle import oracle as O

    return O


def _handle(d, **kw):
    from opencl_dpm_b200 import Dpm3D

    h = Dpm3D(d["nc"], d["nv"], d["faces"])
    if kw:
        h.set_neighbor_params(**kw)
    return h


def _gpu_step(h, d, verts, nsteps=1, mask=15):
    h.set_force_mask(mask)
    h.upload(verts, *[d[k] for k in PKEYS])
    h.step(nsteps, float(d["dt"]), float(d["Kre"]), 0.0, d["PBC"], float(d["L"]))
    return h.download()


@pytest.mark.parametrize("cfg", ["test3d_py", "test3d_cpp"])
@pytest.mark.parametrize("mask", [1, 2, 4, 8, 15])
def test_single_step_force_parity(cfg, mask):
    """One step from the reference demo's initial state: forces and positions vs the all-pairs oracle."""
    O = _oracle()
    d = H.config_test3d_py(64) if cfg == "test3d_py" else H.config_test3d_cpp()
    h = _handle(d)
    V1, F = _gpu_step(h, d, d["verts"], 1, mask)
    Fref = O.forces3d(d["verts"], d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"], which=mask)
    tol = H.force_tol(Fref)
    err = np.abs(F[:, :3] - Fref[:, :3]).max()
    assert err <= tol, f"force error {err:.3e} > {tol:.3e} (|F|max {np.abs(Fref).max():.3f})"
    Vref = d["verts"].copy()
    Vref[:, :3] += Fref[:, :3] * d["dt"]
    assert np.abs(V1[:, :3] - Vref[:, :3]).max() <= tol * float(d["dt"]) + 4e-7 * np.abs(Vref).max()
    h.close()


def test_multi_step_parity_each_step_from_oracle_state():
    """20 steps; before every step the GPU is re-seeded with the oracle's state, so each step is an
    independent parity check on an evolving configuration (contacts appear and disappear)."""
    O = _oracle()
    d = H.config_test3d_py(64)
    h = _handle(d)
    V = d["verts"].copy()
    worst, nbad = 0.0, 0
    for s in range(20):
        V1, F = _gpu_step(h, d, V)
        args = (V, d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"])
        Fref = O.forces3d(*args)
        F64 = O.forces3d(*args, dtype=np.float64)
        w, b = H.assert_forces_close(F[:, :3], Fref[:, :3], F64[:, :3], f"step {s}")
        worst, nbad = max(worst, w), nbad + b
        V[:, :3] += Fref[:, :3] * d["dt"]
    print(f"worst error/tol {worst:.2f}; ill-conditioned vertex-steps beyond tol: {nbad} of {20 * len(V)}")
    h.close()


def test_trajectory_100_steps_vs_oracle():
    """reference test3D.cpp verbatim: 30 cells, 100 steps, dt 0.005 — final positions vs the fp32 all-pairs
    oracle, with the fp32-vs-fp64 oracle drift reported alongside (chaos vs arithmetic)."""
    O = _oracle()
    d = H.config_test3d_cpp()
    h = _handle(d)
    V1, F1 = _gpu_step(h, d, d["verts"], 100)
    Vr, Fr = O.run3d(d["verts"], d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"], 100, d["dt"])
    V64, _ = O.run3d(d["verts"], d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"], 100, d["dt"], dtype=np.float64)
    scale = np.abs(Vr[:, :3]).max()
    gpu_vs_f32 = np.abs(V1[:, :3] - Vr[:, :3]).max() / scale
    f32_vs_f64 = np.abs(Vr[:, :3] - V64[:, :3]).max() / scale
    print(f"drift over 100 steps: gpu-vs-f32 {gpu_vs_f32:.3e}  f32-vs-f64 {f32_vs_f64:.3e}")
    # stated trajectory tolerance: 1e-4 relative over 100 steps, and never worse than 20x the oracle's own fp32 drift
    assert gpu_vs_f32 <= max(1e-4, 20 * f32_vs_f64)
    assert np.abs(F1[:, :3] - Fr[:, :3]).max() <= 50 * H.force_tol(Fr)
    h.close()


def test_trajectory_10000_steps_vs_oracle():
    """north_star: trajectories must stay within a stated tolerance over 1e4 steps.  reference test3D.cpp tissue (30
    cells), 1e4 steps of dt 0.005: GPU vs the fp32 oracle (culled form, rebuilt every 5 steps with a wide margin) and the
    oracle's own fp32-vs-fp64 drift.  Stated tolerance: 1e-3 of the tissue extent, and within 20x the oracle's own
    precision drift."""
    O = _oracle()
    d = H.config_test3d_cpp()
    h = _handle(d)
    nsteps = 10000
    V1, F1 = _gpu_step(h, d, d["verts"], nsteps)
    args = (d["verts"], d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"], nsteps, d["dt"])
    Vr, Fr = O.run3d_culled(*args, rebuild_every=5)
    V64, _ = O.run3d_culled(*args, dtype=np.float64, rebuild_every=5)
    scale = np.abs(Vr[:, :3]).max()
    g = np.abs(V1[:, :3] - Vr[:, :3]).max() / scale
    o = np.abs(Vr[:, :3] - V64[:, :3]).max() / scale
    print(f"drift over {nsteps} steps: gpu-vs-f32 {g:.3e}  f32-vs-f64 {o:.3e}  (rebuilds on the GPU: {h.stats().rebuilds})")
    assert np.isfinite(V1).all()
    assert g <= max(1e-3, 20 * o)
    h.close()


def test_com_and_volume_are_bit_exact():
    """The two ill-conditioned per-cell sums are evaluated in the reference's serial order with
    individually rounded ops: COM and signed volume must equal the oracle's bit for bit."""
    O = _oracle()
    d = H.config_test3d_cpp()
    h = _handle(d)
    _gpu_step(h, d, d["verts"], 1)
    b = h.cell_bounds()  # per-cell scalars of the CURRENT (new) positions, produced by the step kernel's epilogue
    V1, _ = h.download()
    V1 = V1.reshape(d["nc"], d["nv"], 4)
    faces = d["faces"]
    for ci in range(d["nc"]):
        vol = np.float32(0)
        for f in faces:
            P0, P1, P2 = V1[ci, f[0], :3], V1[ci, f[1], :3], V1[ci, f[2], :3]
            c = np.array([P0[1] * P1[2] - P0[2] * P1[1], P0[2] * P1[0] - P0[0] * P1[2], P0[0] * P1[1] - P0[1] * P1[0]], np.float32)
            t = np.float32(np.float32(np.float32(c[0] * P2[0]) + np.float32(c[1] * P2[1])) + np.float32(c[2] * P2[2]))
            vol = np.float32(vol + np.float32(t / np.float32(6.0)))
        assert np.float32(abs(vol)).tobytes() == b[ci, 11].tobytes(), (ci, vol, b[ci, 11])
    for ci in range(d["nc"]):
        s = np.zeros(3, np.float32)
        for i in range(d["nv"]):
            s = (s + V1[ci, i, :3]).astype(np.float32)
        com = (s * np.float32(np.float32(1.0) / np.float32(d["nv"]))).astype(np.float32)
        assert com.tobytes() == b[ci, 8:11].tobytes(), (ci, com, b[ci, 8:11])
    lo, hi = O.aabb3d(V1.reshape(-1, 4), d["nc"])
    assert lo.tobytes() == b[:, 0:3].astype(np.float32).tobytes()
    assert hi.tobytes() == b[:, 4:7].astype(np.float32).tobytes()
    h.close()


@pytest.mark.parametrize("cfg", ["test3d_py", "test3d_cpp", "nonperiodic"])
def test_neighbor_artifacts_bit_exact(cfg):
    """Cell-list artefacts (grid, bin ids, sorted permutation, bin starts, candidate lists) vs the CPU spec."""
    O = _oracle()
    d = H.config_test3d_py(64) if cfg != "test3d_cpp" else H.config_test3d_cpp()
    pbc = 0 if cfg == "nonperiodic" else 1
    h = _handle(d)
    h.upload(d["verts"], *[d[k] for k in PKEYS])
    h.rebuild_neighbors(pbc, float(d["L"]))
    art = h.neighbor_artifacts()
    b = h.cell_bounds()
    lo, hi = O.aabb3d(d["verts"], d["nc"])
    assert lo.tobytes() == np.ascontiguousarray(b[:, 0:3]).tobytes() and hi.tobytes() == np.ascontiguousarray(b[:, 4:7]).tobytes()
    rng = np.float32(1.25) * b[:, 7].max()  # RANGE_HEADROOM * largest contact pad, as the device computes it
    ref = O.cell_list(3, lo, hi, pbc, d["L"], 0.1, rng, h.K)
    assert art["grid"].as_tuple() == ref["grid"].as_tuple()
    for k in ("bin_id", "order", "bin_start", "cand_count"):
        assert np.array_equal(art[k], ref[k]), k
    for i in range(d["nc"]):
        n = ref["cand_count"][i]
        assert np.array_equal(art["cand"][i, :n], ref["cand"][i, :n]), i
    h.close()


def test_contact_set_matches_all_pairs():
    """{(vertex): number of force-carrying contacts} from the culled GPU path == all-pairs oracle, via the
    repulsion-only force: a vertex has nonzero repulsion iff the oracle has a contact there."""
    O = _oracle()
    d = H.config_test3d_py(64)
    h = _handle(d)
    _, F = _gpu_step(h, d, d["verts"], 1, mask=8)
    Fref, con = O.forces3d(d["verts"], d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"], which=8, want_contacts=True)
    gpu_has = np.abs(F[:, :3]).max(1) > 1e-3 * 0.5 * float(d["Kre"]) * 0.999
    ref_has = con[:, 0] > 0
    assert con[:, 0].sum() > 100, "config has too few contacts to be a meaningful test"
    assert np.array_equal(gpu_has, ref_has), f"{(gpu_has != ref_has).sum()} vertices differ"
    print(f"force-carrying contacts {con[:, 0].sum()}, noise-level contacts {con[:, 1].sum()}")
    h.close()


def test_euler_update_one_call_and_rebuilds():
    """dpm3d_euler_update (the CLEulerUpdate seam): host buffers in/out equals upload+step+download, and
    a long run triggers neighbour rebuilds without changing results vs a run with a huge skin."""
    d = H.config_test3d_cpp()
    h = _handle(d)
    V = d["verts"].copy()
    F = np.zeros_like(V)
    ms = h.euler_update(V, *[d[k] for k in PKEYS], 200, float(d["dt"]), float(d["Kre"]), 0.0, d["PBC"], float(d["L"]), forces_out=F)
    assert ms > 0
    V1, F1 = _gpu_step(h, d, d["verts"], 200)
    assert np.array_equal(V, V1) and np.array_equal(F, F1)
    st = h.stats()
    h2 = _handle(d, skin_rel=0.02, max_candidates=32)
    V2, F2 = _gpu_step(h2, d, d["verts"], 200)
    st2 = h2.stats()
    print(f"rebuilds: skin 0.1 -> {st.rebuilds}, skin 0.02 -> {st2.rebuilds}")
    assert st2.rebuilds > st.rebuilds >= 1
    assert np.array_equal(V1, V2) and np.array_equal(F1, F2), "results must not depend on the skin / rebuild schedule"
    h.close(); h2.close()


def test_errors_and_validation():
    from opencl_dpm_b200 import Dpm3D, DpmError

    d = H.config_test3d_cpp()
    h = _handle(d)
    with pytest.raises(DpmError) as e:
        h.step(1, 0.005, 1.0)
    assert e.value.code == 1  # step before upload
    h.upload(d["verts"], *[d[k] for k in PKEYS])
    for bad in [dict(nsteps=0, dt=0.005), dict(nsteps=1, dt=0.0), dict(nsteps=1, dt=0.2)]:
        with pytest.raises(DpmError) as e:
            h.step(bad["nsteps"], bad["dt"], 1.0)
        assert e.value.code == 1
    bad_faces = d["faces"].copy()
    bad_faces[0, 0] = 9999
    with pytest.raises(DpmError) as e:
        Dpm3D(2, 162, bad_faces)
    assert e.value.code == 2 and "Invalid face indices" in str(e.value)
    open_faces = d["faces"][:-1]
    with pytest.raises(DpmError) as e:
        Dpm3D(2, 162, open_faces)
    assert e.value.code == 4
    h.close()


def test_vertex_exactly_on_a_neighbour_vertex():
    """normalize(0) = 0 in the reference's solid-angle sum (shaders/Cell3D_Kernel.cl:289-291): a vertex that coincides
    with a vertex of the neighbour zeroes the terms of the faces at that corner.  The decomposition of the fast path
    cannot express that, so the unit must take the literal sum and agree with the oracle (and stay finite)."""
    O = _oracle()
    d = H.config_test3d_cpp()
    nv = d["nv"]
    V = d["verts"].copy().reshape(d["nc"], nv, 4)
    # translate cell 1 so that its vertex 40 sits exactly on vertex 12 of cell 0
    V[1, :, :3] += (V[0, 12, :3] - V[1, 40, :3])[None, :]
    V[1, 40, :3] = V[0, 12, :3]
    V = V.reshape(-1, 4)
    h = _handle(d)
    _, F = _gpu_step(h, d, V, 1, 8)
    args = (V, d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"])
    Fref = O.forces3d(*args, which=8)
    Fref64 = O.forces3d(*args, which=8, dtype=np.float64)
    assert np.isfinite(F).all() and np.isfinite(Fref).all()
    H.assert_forces_close(F, Fref, Fref64, what="repulsion with a coincident vertex")
    h.close()


@pytest.mark.parametrize("mask", [2, 15])
def test_degenerate_edges_skip_their_faces(mask):
    """SurfaceAreaForceUpdate drops a whole face when one of its edges is shorter than 1e-12 (shaders/Cell3D_Kernel.cl:151).
    Two edges of cell 0 and one of cell 3 are collapsed to zero length: the step kernel finds the affected vertices through
    the per-vertex flags of the previous epilogue (here: the bounds kernel of the upload) and weights every ring edge by the
    number of its non-degenerate faces; all forces must stay finite and equal the oracle's."""
    O = _oracle()
    d = H.config_test3d_cpp()
    nv = d["nv"]
    V = d["verts"].copy().reshape(d["nc"], nv, 4)
    f = d["faces"]
    V[0, f[10, 1]] = V[0, f[10, 0]]      # an edge of face 10 of cell 0
    V[0, f[200, 2]] = V[0, f[200, 1]]    # ... and of face 200
    V[3, f[77, 0]] = V[3, f[77, 2]]
    V = V.reshape(-1, 4)
    h = _handle(d)
    V1, F = _gpu_step(h, d, V, 1, mask)
    args = (V, d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"])
    Fref = O.forces3d(*args, which=mask)
    Fref64 = O.forces3d(*args, which=mask, dtype=np.float64)
    assert np.isfinite(F).all() and np.isfinite(Fref).all() and np.isfinite(V1).all()
    H.assert_forces_close(F[:, :3], Fref[:, :3], Fref64[:, :3], what="forces with degenerate edges")
    # a second step from the GPU's own state: the flags now come from the step kernel's face pass
    h.step(1, float(d["dt"]), float(d["Kre"]), 0.0, d["PBC"], float(d["L"]))
    V2, F2 = h.download()
    args1 = (V1, d["faces"], *[d[k] for k in PKEYS], d["Kre"], d["PBC"], d["L"])
    Fr1 = O.forces3d(*args1, which=mask)
    Fr164 = O.forces3d(*args1, which=mask, dtype=np.float64)
    assert np.isfinite(F2).all()
    H.assert_forces_close(F2[:, :3], Fr1[:, :3], Fr164[:, :3], what="second step with degenerate edges")
    h.close()


def test_cldpm_tissue3d_with_642_vertex_cells():
    """Extension: the subdivision level through the drop-in classes.  9 cells of the 642-vertex mesh
    (Cell3D(start, calA, r0, 3)) through Tissue3D.CLEulerUpdate vs the all-pairs oracle on the same flat arrays."""
    O = _oracle()
    m = H.cldpm()
    c = m.Cell3D([0.0, 0.0, 1.0], 1.0, 1.0, 3)
    c.Ka, c.Kv, c.Ks = 2.0, 5.0, 3.0
    T = m.Tissue3D([c] * 9, 0.6)
    T.Kre = 25.0
    H.reset_drand48()
    T.Disperse2D()
    cells = T.Cells
    nv = cells[0].NV
    assert nv == 642 and cells[0].NF == 1280
    V0 = np.zeros((9 * nv, 4), np.float32)
    for i, x in enumerate(cells):
        V0[i * nv:(i + 1) * nv, :3] = np.asarray(x.Verts, np.float32)
    faces = np.asarray(cells[0].GetFaces(), np.uint32)
    P = H.params3d(9, 1.0, 1.0, 5.0, 2.0, 3.0, nf=1280)
    T.CLEulerUpdate(3, 0.01)
    out = T.Cells
    V = np.concatenate([np.asarray(x.Verts, np.float32) for x in out])
    F = np.concatenate([np.asarray(x.GetForces(), np.float32).T for x in out])
    Vr, Fr = O.run3d(V0, faces, *[P[k] for k in PKEYS], 25.0, int(T.PBC), np.float32(T.L), 3, np.float32(0.01))
    assert np.abs(Fr).max() > 1.0
    assert np.abs(V - Vr[:, :3]).max() <= 4e-6
    assert np.abs(F - Fr[:, :3]).max() <= 4 * H.force_tol(Fr)


def test_step_resident_equals_repeated_cleulerupdate():
    """Extension: StepResident keeps the tissue on the device between calls (one upload, no per-call
    download); 4 x 5 resident steps + SyncCells must equal 4 CLEulerUpdate(5) calls bit for bit (results do not depend on
    the neighbour-list rebuild schedule), and editing Cells + InvalidateDevice() must be honoured."""
    m = H.cldpm()

    def tissue():
        c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
        c.Ka, c.Kv, c.Ks = 2.0, 5.0, 3.0
        T = m.Tissue3D([c] * 16, 0.35)
        T.Kre = 25.0
        H.reset_drand48()
        T.Disperse2D()
        return T

    A, B = tissue(), tissue()
    for _ in range(4):
        A.CLEulerUpdate(5, 0.01)
        B.StepResident(5, 0.01)
    B.SyncCells()
    Va, Vb = H.flat3d(A)["verts"], H.flat3d(B)["verts"]
    assert np.array_equal(Va, Vb)
    Fa = np.concatenate([np.asarray(x.GetForces(), np.float32).T for x in A.Cells])
    Fb = np.concatenate([np.asarray(x.GetForces(), np.float32).T for x in B.Cells])
    assert np.array_equal(Fa, Fb) and np.abs(Fa).max() > 0
    # edit the host copy, invalidate, continue: both paths see the edit
    for T in (A, B):
        cells = T.Cells
        v = np.asarray(cells[3].Verts, np.float32) + np.float32(0.05)
        cells[3].Verts = v.tolist()
        T.Cells = cells
    B.InvalidateDevice()
    A.CLEulerUpdate(3, 0.01)
    B.StepResident(3, 0.01)
    B.SyncCells()
    assert np.array_equal(H.flat3d(A)["verts"], H.flat3d(B)["verts"])
    with pytest.raises(ValueError):
        B.StepResident(0, 0.01)


def test_zero_copy_views_of_positions_and_forces():
    """Extension: Tissue3D.PositionsView() / ForcesView() are numpy views [NCELLS][NV][4] of the packed host
    arrays of the last CLEulerUpdate / SyncCells — no copy (the next call rewrites the same memory), equal to what the
    reference surface returns as lists (Cell3D.Verts / GetForces, src/CellWrapper.cpp:7-19)."""
    m = H.cldpm()
    c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
    c.Ka, c.Kv, c.Ks = 2.0, 5.0, 3.0
    T = m.Tissue3D([c] * 16, 0.35)
    T.Kre = 25.0
    H.reset_drand48()
    T.Disperse2D()
    assert T.PositionsView() is None  # nothing packed yet
    T.CLEulerUpdate(5, 0.01)
    P, F = T.PositionsView(), T.ForcesView()
    assert P.shape == (16, 162, 4) and F.shape == (16, 162, 4) and P.dtype == np.float32 and not P.flags.owndata
    cells = T.Cells
    V = np.stack([np.asarray(x.Verts, np.float32) for x in cells])
    Fc = np.stack([np.asarray(x.GetForces(), np.float32).T for x in cells])
    assert np.array_equal(P[:, :, :3], V) and np.array_equal(F[:, :, :3], Fc) and np.abs(Fc).max() > 0
    addr = P.__array_interface__["data"][0]
    before = P[:, :, :3].copy()
    T.CLEulerUpdate(5, 0.01)
    P2 = T.PositionsView()
    assert P2.__array_interface__["data"][0] == addr, "the view is the staging array itself, not a copy"
    assert not np.array_equal(P[:, :, :3], before), "the earlier view shows the new state: same memory"
    T.StepResident(3, 0.01)
    T.SyncCells()
    assert np.array_equal(T.PositionsView()[:, :, :3], np.stack([np.asarray(x.Verts, np.float32) for x in T.Cells]))
    del T
    assert np.isfinite(P).all()  # the view keeps the array alive


def test_threaded_pack_unpack_equals_flat_path():
    """A tissue large enough for the host classes' threaded pack/unpack (> 2e5 vertices): Tissue3D.CLEulerUpdate must
    equal the C ABI driven with the same flat arrays, bit for bit (positions, last-step forces, Volume)."""
    from opencl_dpm_b200 import Dpm3D

    m = H.cldpm()
    c = m.Cell3D([0.0, 0.0, 1.0], 1.0, 1.0)
    c.Ka, c.Kv, c.Ks = 2.0, 5.0, 3.0
    n = 1400
    T = m.Tissue3D([c] * n, 0.35)
    T.Kre = 25.0
    H.reset_drand48()
    T.Disperse2D()
    d = H.flat3d(T)
    P = H.params3d(n, 1.0, 1.0, 5.0, 2.0, 3.0)
    T.CLEulerUpdate(4, 0.01)
    out = T.Cells
    Vc = np.concatenate([np.asarray(x.Verts, np.float32) for x in out])
    Fc = np.concatenate([np.asarray(x.GetForces(), np.float32).T for x in out])
    h = Dpm3D(n, 162, d["faces"])
    V = d["verts"].copy()
    F = np.zeros_like(V)
    h.euler_update(V, *[P[k] for k in PKEYS], 4, 0.01, 25.0, 0.0, d["PBC"], float(d["L"]), forces_out=F)
    h.close()
    assert np.array_equal(Vc, V[:, :3]) and np.array_equal(Fc, F[:, :3])
    assert 0.5 < out[5].GetVolume() < 4.2 and out[5].GetVolume() == T.Cells[5].GetVolume()  # unpack refreshed the cell
