"""GPU parity on the EXACT batches bench.py times and on the regimes the headline excludes.

bench.py's batch = synth.monolayer3d(64, subdiv) / synth.tissue2d(64, nv=64) advanced `--equilibrate` = 100 timesteps on the
GPU.  Each test takes that state off the device, lets the GPU do ONE further timestep and compares forces, positions and
the contact set with the CPU oracle in culled form (CPU cell list + the literal kernels of shaders/Cell3D_Kernel.cl:251-310
and shaders/Cell2D_kernel.cl:121-268; the culled form equals the all-pairs form exactly, tests/test_oracle_cpu.py) evaluated
on the same positions.  Tolerance: helpers.assert_forces_close — every vertex within 1e-5 * max(|F_ref|_inf, 1e-3) of the
fp32 oracle, the few ill-conditioned vertices within 8x the oracle's own fp32-vs-fp64 error.
Also: timestep 5 from the raw lattice (contact-dominated: ~10^5 units), and a deliberately NON-star-shaped tissue whose
contacts all take the general (patch-box) evaluation of the contact kernel instead of the star-shaped fast path."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

PK3 = ("Kv", "Ka", "Ks", "v0", "a0", "l0")
PK2 = ("Ka", "Kl", "Kb", "a0", "l0", "r0")


def _oracle():
    from oracle import oracle as O

    return O


def _culled3d(O, d, V, dtype=np.float32, which=15, want_contacts=False):
    nc, nv, f = d["nc"], d["nv"], d["faces"]
    Vc = V.reshape(nc, nv, 4)
    emax = max(float(np.linalg.norm(Vc[:, f[:, i], :3] - Vc[:, f[:, (i + 1) % 3], :3], axis=2).max()) for i in range(3))
    lo, hi = O.aabb3d(V, nc)
    cl = O.cell_list(3, lo, hi, d["PBC"], d["L"], 0.1, 1.25 * 0.34 * emax, 64)
    assert cl["cand_count"].max() <= 64
    return O.forces3d(V, f, *[d[k] for k in PK3], d["Kre"], d["PBC"], d["L"], which=which, cand_count=cl["cand_count"], cand=cl["cand"],
                      dtype=dtype, want_contacts=want_contacts)


def _check_one_more_step_3d(d, V0, presteps, what, min_units=0, min_nonstar=0, min_contacts=0):
    """GPU: upload V0, advance `presteps` timesteps, download -> V; upload V, ONE timestep -> (V1, F).  Oracle on V."""
    from opencl_dpm_b200 import Dpm3D

    O = _oracle()
    h = Dpm3D(d["nc"], d["nv"], d["faces"])
    P = [d[k] for k in PK3]
    args = (float(d["dt"]), float(d["Kre"]), 0.0, d["PBC"], float(d["L"]))
    h.upload(V0, *P)
    if presteps > 0:
        h.step(presteps, *args)
    V, _ = h.download()
    assert np.isfinite(V).all()
    h.upload(V, *P)
    h.step(1, *args)
    V1, F = h.download()
    st = h.stats()
    # reserved[0]: units that fell back to the literal all-faces sum; reserved[1] low word: units whose neighbour is not
    # star-shaped (general patch-box evaluation)
    units, literal, nonstar = int(st.contact_evals), int(st.reserved[0]), int(st.reserved[1] & 0xffffffff)
    Fref, con = _culled3d(O, d, V, want_contacts=True)
    F64 = _culled3d(O, d, V, dtype=np.float64)
    worst, nbad = H.assert_forces_close(F[:, :3], Fref[:, :3], F64[:, :3], what)
    tol = H.force_tol(Fref)
    Vref = V.copy()
    Vref[:, :3] += Fref[:, :3] * d["dt"]
    perr = np.abs(V1[:, :3] - Vref[:, :3]).max(1)
    ferr = np.abs(F[:, :3] - Fref[:, :3]).max(1)
    # positions follow the forces: x += F dt in fp32 (shaders/Cell3D_Kernel.cl:380)
    assert (perr <= (ferr + tol) * float(d["dt"]) + 4e-7 * np.abs(Vref[:, :3]).max()).all(), f"{what}: positions"
    # contact set: a vertex carries repulsion on the GPU iff the oracle has a force-carrying contact there (mask = 8)
    h.set_force_mask(8)
    h.upload(V, *P)
    h.step(1, *args)
    _, F8 = h.download()
    gpu_has = np.abs(F8[:, :3]).max(1) > 1e-3 * 0.5 * float(d["Kre"]) * 0.999
    ref_has = con[:, 0] > 0
    ndiff = int((gpu_has != ref_has).sum())
    # a contact whose |w| sits within rounding of the 1e-3 classification threshold may fall on either side
    assert ndiff <= max(2, int(2e-4 * ref_has.sum())), f"{what}: contact sets differ at {ndiff} vertices of {int(ref_has.sum())} in contact"
    h.close()
    print(f"{what}: |F|max {np.abs(Fref).max():.3f}, worst err/tol {worst:.2f}, ill-conditioned beyond tol {nbad}, units/step {units}, "
          f"non-star units {nonstar}, literal fallbacks {literal}, vertices in contact {int(ref_has.sum())} (set differs at {ndiff})")
    assert units >= min_units, f"{what}: only {units} contact units — not the regime this test is for"
    assert nonstar >= min_nonstar, f"{what}: only {nonstar} units with a non-star-shaped neighbour"
    assert int(ref_has.sum()) >= min_contacts, f"{what}: only {int(ref_has.sum())} vertices carry a contact force"
    return units, nonstar, literal


@pytest.mark.parametrize("subdiv,name", [(3, "D642"), (2, "D162")])
def test_bench_batch_3d_one_more_step(subdiv, name):
    """bench.py's headline batch (config D: 4096 cells, lattice + 100 timesteps) and its 162-vertex sibling."""
    from opencl_dpm_b200 import synth

    d = synth.monolayer3d(64, subdiv=subdiv)
    _check_one_more_step_3d(d, d["verts"], 100, f"{name} batch (lattice + 100 timesteps)")


@pytest.mark.parametrize("presteps", [0, 1, 5])
def test_contact_dominated_timesteps_from_the_raw_lattice(presteps):
    """The first timesteps of config D from the raw lattice: every cell overlaps its neighbours, ~10^5-10^6 contact units
    per timestep, thousands of vertices inside a neighbour at timestep 0 — the regime the headline batch has left behind
    (`trajectory` phase 0-30 of the bench line)."""
    from opencl_dpm_b200 import synth

    d = synth.monolayer3d(64, subdiv=3)
    _check_one_more_step_3d(d, d["verts"], presteps, f"D642 timestep {presteps} from the lattice", min_units=20000,
                            min_contacts=1000 if presteps == 0 else 0)


@pytest.mark.parametrize("subdiv,name", [(2, "C162"), (3, "C642")])
def test_bench_batch_config_c(subdiv, name):
    """BASELINE config C as bench.py runs it (reference test3D.py: 64 cells placed by Disperse2D() with their centres ON the
    substrate plane, + 100 timesteps): the substrate force folds the lower half of every cell inwards, no cell is star-shaped
    any more, every contact unit takes the general patch-box evaluation."""
    from opencl_dpm_b200 import synth

    d = synth.test3d_config(64, subdiv=subdiv)
    _check_one_more_step_3d(d, d["verts"], 100, f"{name} batch (Disperse2D + 100 timesteps)", min_units=1000, min_nonstar=1000)


def _dimpled(d, depth=1.4, cap=0.5):
    """push the cap z > z_c + cap of every cell inwards, past the centre: closed, not self-intersecting, NOT star-shaped
    about the centroid (72 of 320 faces face away from it)"""
    nc, nv = d["nc"], d["nv"]
    V = d["verts"].reshape(nc, nv, 4).copy()
    plane = V[:, :, 2].mean(1, keepdims=True) + np.float32(cap)
    V[:, :, 2] = np.where(V[:, :, 2] > plane, plane - np.float32(depth) * (V[:, :, 2] - plane), V[:, :, 2])
    return V.reshape(-1, 4)


def test_non_star_shaped_tissue_takes_the_general_evaluation():
    """256 dimpled cells (162 vertices) on the overlapping lattice: every neighbour fails the star-shape test, so every
    contact unit (>= 10^3) is evaluated by the general patch-box evaluation — the path a crumpled tissue lives on."""
    from opencl_dpm_b200 import synth

    d = synth.monolayer3d(16, subdiv=2)
    V0 = _dimpled(d)
    units, nonstar, literal = _check_one_more_step_3d(d, V0, 0, "non-star tissue", min_units=1000, min_nonstar=1000, min_contacts=1000)
    assert nonstar == units, "every neighbour is non-star-shaped: every unit takes the general evaluation"
    assert literal <= units // 100, "the literal all-faces sum is only the fallback for coincident vertices"
    # and a few timesteps later (the dimples relax but stay): still in parity
    _check_one_more_step_3d(d, V0, 8, "non-star tissue + 8 timesteps", min_units=1000, min_nonstar=500)


def test_bench_batch_2d_one_more_step():
    """bench.py's config B batch (4096 cells x 64 vertices, lattice + 100 timesteps): one further timestep vs the culled
    CPU oracle, inside-classification (RepulsionForceUpdate :166-196) bit for bit."""
    from opencl_dpm_b200 import Dpm2D, synth

    O = _oracle()
    d = synth.tissue2d(64, nv=64)
    h = Dpm2D(d["nc"], d["S"])
    h.set_neighbor_params(0.1, 64)
    P = [d[k] for k in PK2]
    args = (float(d["dt"]), float(d["Kre"]), float(d["Kat"]), d["PBC"], float(d["L"]))
    h.upload(d["verts"], d["nv"], *P)
    h.step(100, *args)
    V, _ = h.download()
    h.upload(V, d["nv"], *P)
    h.step(1, *args)
    V1, F = h.download()
    lo, hi = O.aabb2d(V, d["nv"])
    cl = O.cell_list(2, lo, hi, d["PBC"], d["L"], 0.1, float(d["l0"].max()), 64, far2d=True)
    assert cl["cand_count"].max() <= 64
    kw = dict(cand_count=cl["cand_count"], cand=cl["cand"])
    Fref, inside = O.forces2d(V, d["nv"], *P, d["Kre"], d["Kat"], d["PBC"], d["L"], want_inside=True, **kw)
    F64 = O.forces2d(V, d["nv"], *P, d["Kre"], d["Kat"], d["PBC"], d["L"], dtype=np.float64, **kw)
    n = d["nc"] * d["S"]
    worst, nbad = H.assert_forces_close(F.reshape(n, 2), Fref.reshape(n, 2), F64.reshape(n, 2), "B2D batch")
    tol = H.force_tol(Fref)
    Vref = V + Fref * d["dt"]
    assert np.abs(V1 - Vref).max() <= 2 * tol * float(d["dt"]) + 4e-7 * np.abs(Vref).max()
    # repulsion-only force is non-zero exactly where the oracle classifies the vertex inside another polygon
    h.set_force_mask(16)
    h.upload(V, d["nv"], *P)
    h.step(1, *args)
    _, F16 = h.download()
    assert np.array_equal(np.abs(F16).max(2) > 0, inside > 0)
    print(f"B2D batch: |F|max {np.abs(Fref).max():.3f}, worst err/tol {worst:.2f}, beyond tol {nbad}, vertices inside a neighbour {int(inside.sum())}")
    h.close()
