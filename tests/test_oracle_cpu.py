"""CPU tests (-m "not gpu"): the oracle against known answers and against itself.

  * geometry constants produced by the reference's own src/cell.cpp (values measured by
    compiling the reference's cell.cpp; also re-checked against oracle/_ref when it is built);
  * analytic known-answer tests that follow from the kernel formulas;
  * the culled form (cell list + exact culls) equals the all-pairs form exactly.
"""
import numpy as np
import pytest

from oracle import oracle as O

PK3 = ("Kv", "Ka", "Ks", "v0", "a0", "l0")


def test_icosphere_matches_reference_constants():
    V, F = O.icosphere(2)
    assert V.shape == (162, 3) and F.shape == (320, 3)
    assert tuple(F[0]) == (0, 42, 44) and tuple(F[319]) == (160, 161, 159)
    np.testing.assert_allclose(V[12], [-0.809017003, 0.44721356, 0.273520619], rtol=0, atol=1e-7)
    r = np.linalg.norm(V.astype(np.float64), axis=1)
    assert 0.9595 < r.min() < 0.9605 and abs(r.max() - 1.0) < 1e-6  # not a unit sphere
    val = np.bincount(F.ravel(), minlength=162)
    assert (val[:12] == 5).all() and (val[12:] == 6).all()
    p = O.cell3d_params(1.0, 1.0, 320)
    np.testing.assert_allclose([p["v0"], p["sa0"], p["a0"], p["l0"]], [4.18879032, 12.5663719, 0.039269913, 0.228822812], rtol=2e-7)
    p = O.cell3d_params(1.05, 1.8, 320)
    np.testing.assert_allclose([p["v0"], p["sa0"], p["a0"]], [24.4290237, 42.0611458, 0.131441087], rtol=2e-7)


def test_icosphere_subdiv3_is_closed_and_oriented():
    V, F = O.icosphere(3)
    assert V.shape == (642, 3) and F.shape == (1280, 3)
    e = {}
    for a, b, c in F:
        for u, v in ((a, b), (b, c), (c, a)):
            e[(u, v)] = e.get((u, v), 0) + 1
    assert all(n == 1 for n in e.values()) and all((v, u) in e for (u, v) in e)  # every edge once per direction
    P = V.astype(np.float64)
    vol6 = np.einsum("ij,ij->i", np.cross(P[F[:, 0]], P[F[:, 1]]), P[F[:, 2]]).sum()
    assert vol6 > 0  # CCW-outward


def test_cell2d_matches_reference_constants():
    v, p = O.cell2d_init(0, 0, 1.05, 32, 1.0)
    np.testing.assert_allclose([p["calA0"], p["a0"], p["l0"]], [1.05338645, 3.12144494, 0.200875357], rtol=2e-7)
    np.testing.assert_allclose(v[0], [0.980785251, 0.195090324], rtol=2e-7)
    _, p = O.cell2d_init(0, 0, 1.2, 64, 1.0)
    np.testing.assert_allclose([p["calA0"], p["a0"], p["l0"]], [1.20096481, 3.13654852, 0.107501894], rtol=2e-7)


def _one_cell(start=(0.0, 0.0, 1.3), calA=1.0, r0=1.0):
    V, F = O.icosphere(2)
    p = O.cell3d_params(calA, r0, 320)
    return O.cell3d_place(V, r0, start), F, p


def test_kat_edge_and_volume_forces_sum_to_zero():
    v4, F, p = _one_cell()
    one = np.ones(1, np.float32)
    for which in (1, 2):
        Fo = O.forces3d(v4, F, one * 5, one * 2, one * 3, one * p["v0"], one * p["a0"], one * p["l0"], 25.0, 0, 10.0, which=which,
                        dtype=np.float64)
        assert np.abs(Fo[:, :3].sum(0)).max() < 1e-12
        assert np.abs(Fo).max() > 1e-3


def test_kat_isolated_cell_has_no_repulsion_and_winding_of_com_is_one():
    v4, F, p = _one_cell()
    one = np.ones(1, np.float32)
    Fo = O.forces3d(v4, F, one, one, one, one * p["v0"], one * p["a0"], one * p["l0"], 25.0, 1, 10.0, which=8)
    assert np.abs(Fo).max() == 0.0
    # two cells: second cell's vertex 0 moved to the first cell's COM -> winding number 1 -> |F| = 0.5*Kc*|dir|
    a, _, _ = _one_cell((0, 0, 0))
    b, _, _ = _one_cell((5, 0, 0))
    b[0, :3] = 0.0
    two = np.ones(2, np.float32)
    Fo = O.forces3d(np.concatenate([a, b]), F, two, two, two, two * p["v0"], two * p["a0"], two * p["l0"], 25.0, 0, 100.0, which=8,
                    dtype=np.float64)
    np.testing.assert_allclose(np.linalg.norm(Fo[162, :3]), 12.5, rtol=1e-7)  # 4*M_PI_F is the float pi


def test_kat_reference_drops_faces_subtending_more_than_pi():
    """Documents SURVEY-level behaviour found while building the oracle: `if (denom < 1e-8) continue`
    (shaders/Cell3D_Kernel.cl:293-295) removes the nearest face for points just outside a mesh, so the
    reference applies a partial contact force slightly OUTSIDE the neighbour's surface."""
    a, F, p = _one_cell((0, 0, 0))
    b, _, _ = _one_cell((5, 0, 0))
    f0 = F[0]
    c = a[f0, :3].astype(np.float64).mean(0)
    n = np.cross(a[f0[1], :3] - a[f0[0], :3], a[f0[2], :3] - a[f0[0], :3]).astype(np.float64)
    n /= np.linalg.norm(n)
    two = np.ones(2, np.float32)
    args = (F, two, two, two, two * p["v0"], two * p["a0"], two * p["l0"], 25.0, 0, 100.0)
    b[0, :3] = c + 0.02 * n  # 0.02 outside the face: face subtends > pi -> dropped -> |wn| ~ 0.4
    Fo = O.forces3d(np.concatenate([a, b]), *args, which=8, dtype=np.float64)
    assert 2.0 < np.linalg.norm(Fo[162, :3]) < 6.3
    b[0, :3] = c + 0.5 * n  # far enough (> longest edge / 3): every face counted, true winding number 0
    Fo = O.forces3d(np.concatenate([a, b]), *args, which=8, dtype=np.float64)
    assert np.abs(Fo[162]).max() == 0.0


def test_kat_2d_fresh_polygon():
    """a0 = GetArea() => zero area strain; perimeter+bending forces radial and equal on every vertex."""
    v, p = O.cell2d_init(3.0, 4.0, 1.05, 32, 1.0)
    V = v[None].copy()
    one = np.ones(1, np.float32)
    args = (np.array([32], np.int32), one, one, one * 0.1, one * p["a0"], one * p["l0"], one)
    Fa = O.forces2d(V, *args, 50.0, 0.0, 1, 20.0, which=1, dtype=np.float64)
    assert np.abs(Fa).max() < 1e-6
    Fp = O.forces2d(V, *args, 50.0, 0.0, 1, 20.0, which=6, dtype=np.float64)[0]
    mag = np.linalg.norm(Fp, axis=1)
    assert mag.std() / mag.mean() < 1e-5
    rad = (V[0] - np.array([3.0, 4.0]))
    cosang = (Fp * rad).sum(1) / (mag * np.linalg.norm(rad, axis=1))
    assert np.abs(np.abs(cosang) - 1).max() < 1e-5
    Fr = O.forces2d(V, *args, 50.0, 0.5, 1, 20.0, which=24)
    assert np.abs(Fr).max() == 0.0  # isolated cell: no contacts


def _tissue3d(n=4, seed=3, subdiv=2):
    rng = np.random.RandomState(seed)
    V, F = O.icosphere(subdiv)
    p = O.cell3d_params(1.0, 1.0, F.shape[0])
    cells = []
    for i in range(n):
        for j in range(n):
            cells.append(O.cell3d_place(V, 1.0, [1.7 * i + 0.2 * rng.rand(), 1.7 * j + 0.2 * rng.rand(), 1.0 + 0.1 * rng.rand()]))
    nc = n * n
    one = np.ones(nc, np.float32)
    return dict(nc=nc, verts=np.concatenate(cells), faces=F, Kv=one * 5, Ka=one * 2, Ks=one * 3, v0=one * p["v0"], a0=one * p["a0"],
                l0=one * p["l0"], L=np.float32(1.7 * n))


@pytest.mark.parametrize("pbc", [0, 1])
def test_culled_equals_all_pairs_3d(pbc):
    d = _tissue3d()
    args = (d["verts"], d["faces"], *[d[k] for k in PK3], 25.0, pbc, d["L"])
    Fa, ca = O.forces3d(*args, want_contacts=True)
    lo, hi = O.aabb3d(d["verts"], d["nc"])
    cl = O.cell_list(3, lo, hi, pbc, d["L"], 0.1, 1.25 * 0.34 * 0.36, 32)
    assert cl["cand_count"].max() <= 32
    Fc, cc = O.forces3d(*args, cand_count=cl["cand_count"], cand=cl["cand"], want_contacts=True)
    assert ca[:, 0].sum() > 50
    assert np.array_equal(ca[:, 0], cc[:, 0])
    assert np.abs(Fa - Fc).max() <= 1e-6 * 12.5  # at most noise-level contacts differ


def _tissue2d(n=6, seed=5, nv=24):
    rng = np.random.RandomState(seed)
    cells, prm = [], None
    for i in range(n):
        for j in range(n):
            v, prm = O.cell2d_init(1.8 * i + 0.3 * rng.rand(), 1.8 * j + 0.3 * rng.rand(), 1.1, nv, 1.0)
            cells.append(v)
    nc = n * n
    one = np.ones(nc, np.float32)
    return dict(nc=nc, verts=np.stack(cells), nv=np.full(nc, nv, np.int32), Ka=one, Kl=one, Kb=one * 0.1, a0=one * prm["a0"],
                l0=one * prm["l0"], r0=one, L=np.float32(1.8 * n))


@pytest.mark.parametrize("pbc", [0, 1])
def test_culled_equals_all_pairs_2d_including_wrap_quirk(pbc):
    d = _tissue2d()
    # with pbc the box is exactly the lattice period, so cells at opposite edges are > L apart in raw
    # coordinates and exercise the floor/round wrap of the reference's repulsion kernel
    args = (d["verts"], d["nv"], d["Ka"], d["Kl"], d["Kb"], d["a0"], d["l0"], d["r0"], 10.0, 0.5, pbc, d["L"])
    Fa, ia = O.forces2d(*args, want_inside=True)
    lo, hi = O.aabb2d(d["verts"], d["nv"])
    cl = O.cell_list(2, lo, hi, pbc, d["L"], 0.1, float(d["l0"].max()), 64, far2d=True)
    Fc, ic = O.forces2d(*args, cand_count=cl["cand_count"], cand=cl["cand"], want_inside=True)
    assert ia.sum() > 10
    assert np.array_equal(ia, ic)
    assert np.array_equal(Fa, Fc)


def test_cell_list_is_sorted_and_complete():
    d = _tissue3d(5)
    lo, hi = O.aabb3d(d["verts"], d["nc"])
    for pbc in (0, 1):
        cl = O.cell_list(3, lo, hi, pbc, d["L"], 0.1, 0.5, 32)
        order, bs, bid = cl["order"], cl["bin_start"], cl["bin_id"]
        assert sorted(order.tolist()) == list(range(d["nc"]))
        keys = [(bid[c], c) for c in order]
        assert keys == sorted(keys)  # stable sort by (bin, id)
        assert bs[0] == 0 and bs[-1] == d["nc"] and (np.diff(bs) >= 0).all()
        for i in range(d["nc"]):
            c = cl["cand"][i, :cl["cand_count"][i]]
            assert (np.diff(c) > 0).all() and i not in c


def test_f32_oracle_tracks_f64_oracle():
    d = _tissue3d(3)
    args = (d["verts"], d["faces"], *[d[k] for k in PK3], 25.0, 1, d["L"])
    V32, _ = O.run3d(*args, 20, 0.01)
    V64, _ = O.run3d(*args, 20, 0.01, dtype=np.float64)
    assert np.abs(V32[:, :3] - V64[:, :3]).max() < 5e-5


def test_kat_coincident_vertices_follow_opencl_normalize_zero():
    """Two vertices of different cells at exactly the same point: OpenCL's normalize() returns a zero vector unchanged
    (measured on the reference's own runtime on the B200 box, profiles/r01_cl_semantics.log), so the attraction adds
    nothing (shaders/Cell2D_kernel.cl:263-264) and the solid-angle terms of the faces at that corner are zero
    (shaders/Cell3D_Kernel.cl:289-298) -- forces stay finite."""
    d = _tissue2d(n=3, nv=24)
    V = d["verts"].copy()
    V[1, 5] = V[0, 17]  # cell 1's vertex 5 lands exactly on cell 0's vertex 17
    args = (d["nv"], d["Ka"], d["Kl"], d["Kb"], d["a0"], d["l0"], d["r0"], 10.0, 0.5, 0, d["L"])
    F = O.forces2d(V, *args, which=8)
    assert np.isfinite(F).all()
    # the coincident pair contributes nothing: same force as with the partner vertex far away, for that pair's share
    V2 = V.copy()
    V2[0, 17] += np.float32(1e-3)
    F2 = O.forces2d(V2, *args, which=8)
    assert np.abs(F[1, 5] - F2[1, 5]).max() < 2e-4  # coefficient ~ Kat/n * d/l0 with d = 1e-3
    t = _tissue3d(n=2)
    W = t["verts"].copy().reshape(t["nc"], -1, 4)
    W[1, 7, :3] = W[0, 30, :3]  # a vertex of cell 1 exactly on a vertex of cell 0
    F3 = O.forces3d(W.reshape(-1, 4), t["faces"], *[t[k] for k in PK3], 25.0, 0, t["L"], which=8)
    assert np.isfinite(F3).all()
