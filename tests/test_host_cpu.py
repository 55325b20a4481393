"""CPU tests (-m "not gpu"): the C ABI library loads and exports every symbol include/dpm_b200.h declares,
and the host-side mirror of the reference interface (clDPM module) behaves like the reference's classes.
No compute entry point is called without a GPU."""
import os
import re

import numpy as np
import pytest

import helpers as H
from conftest import ROOT, has_gpu


def test_library_exports_every_declared_symbol():
    from opencl_dpm_b200 import capi

    hdr = open(os.path.join(ROOT, "include", "dpm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(dpm[23]?d?_\w+)\s*\(", hdr)))
    assert len(names) >= 35, names
    lib = capi.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared but not exported: {missing}"
    assert b"sm_100a" in lib.dpm_version()


def test_no_cpu_fallback_without_device():
    from opencl_dpm_b200 import Dpm3D, DpmError, capi

    if has_gpu():
        pytest.skip("a GPU is present")
    V, F = capi.icosphere(2)
    with pytest.raises(DpmError) as e:
        Dpm3D(2, 162, F)
    assert e.value.code == 3  # DPM_ERR_CUDA: the product path fails loudly, it never computes on the CPU


def test_abi_geometry_helpers_match_oracle_and_reference_constants():
    from opencl_dpm_b200 import capi
    from oracle import oracle as O

    for s in (2, 3):
        V, F = capi.icosphere(s)
        Vo, Fo = O.icosphere(s)
        assert np.array_equal(V, Vo) and np.array_equal(F, Fo)
    p, po = capi.cell3d_params(1.05, 1.8, 320), O.cell3d_params(1.05, 1.8, 320)
    assert all(p[k] == po[k] for k in p)


def test_cldpm_surface_matches_reference_wrappers():
    """names/readwrite-ness of src/CellWrapper.cpp:7-29 and src/TissueWrapper.cpp:7-29"""
    m = H.cldpm()
    c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
    for a in ("Kv", "Ka", "Ks", "Verts"):
        setattr(c, a, getattr(c, a))
    for meth in ("GetVolume", "GetPositions", "GetVesselPositions", "GetFaces", "GetForces"):
        assert callable(getattr(c, meth))
    assert not hasattr(c, "Forces")  # 3D forces are only reachable through GetForces()
    c2 = m.Cell2D(0.0, 0.0, 1.05, 32, 1.0)
    for a in ("Ka", "Kl", "Kb", "Verts", "Forces"):
        setattr(c2, a, getattr(c2, a))
    T = m.Tissue3D([c] * 3, 0.35)
    for a in ("Kre", "Kat", "Cells"):
        setattr(T, a, getattr(T, a))
    for a in ("NCELLS", "L", "PBC"):
        getattr(T, a)
        with pytest.raises(AttributeError):
            setattr(T, a, 1)
    assert callable(T.CLEulerUpdate) and callable(T.Disperse2D)
    T2 = m.Tissue2D([c2] * 3, 0.85)
    for a in ("Kre", "Kat", "Cells"):
        setattr(T2, a, getattr(T2, a))
    assert (T2.Kre, T2.Kat, T2.PBC) == (1.0, 0.0, True)
    assert callable(T2.CLEulerUpdate) and callable(T2.Disperse)


def test_cldpm_geometry_equals_reference_probe_values():
    """values obtained by compiling the reference's own src/cell.cpp"""
    m = H.cldpm()
    c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
    np.testing.assert_allclose(c.GetVolume(), 3.84748006, rtol=3e-7)
    faces = np.asarray(c.GetFaces())
    assert faces.shape == (320, 3) and tuple(faces[0]) == (0, 42, 44) and tuple(faces[319]) == (160, 161, 159)
    np.testing.assert_allclose(np.asarray(c.Verts)[12], [-0.809017003, 0.44721356, 0.273520619], atol=1e-7)
    pos = np.asarray(c.GetPositions())
    assert pos.shape == (3, 162)
    c = m.Cell3D([7.0, 6.0, 1.3], 1.05, 1.8)
    np.testing.assert_allclose(c.GetVolume(), 22.4385052, rtol=3e-6)
    T = m.Tissue3D([m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)] * 64, 0.35)
    np.testing.assert_allclose(T.L, np.cbrt(64 * 4.18879032) / 0.35, rtol=1e-6)
    c2 = m.Cell2D(0.0, 0.0, 1.05, 32, 1.0)
    np.testing.assert_allclose(np.asarray(c2.Verts)[0], [0.980785251, 0.195090324], rtol=2e-7)
    T2 = m.Tissue2D([c2] * 32, 0.85)
    np.testing.assert_allclose(T2.L, np.sqrt(32 * 3.12144494) / 0.85, rtol=1e-6)


def test_cleulerupdate_validation_maps_to_reference_exceptions():
    """src/Tissue3D.cpp:123-135 -> std::invalid_argument (ValueError); :159-171 -> std::runtime_error"""
    m = H.cldpm()
    c = m.Cell3D([0.0, 0.0, 1.0], 1.0, 1.0)
    c.Kv, c.Ka, c.Ks = 5.0, 2.0, 3.0
    T = m.Tissue3D([c] * 2, 0.35)
    T.Kre = 25.0
    with pytest.raises(ValueError, match="nsteps must be positive"):
        T.CLEulerUpdate(0, 0.01)
    with pytest.raises(ValueError, match="dt must be positive and reasonable"):
        T.CLEulerUpdate(1, 0.5)
    with pytest.raises(ValueError):
        T.CLEulerUpdate(1, -0.01)
    bad = m.Cell3D([0.0, 0.0, 1.0], 1.0, 1.0)  # Kv = Ka = 0 as constructed
    T = m.Tissue3D([bad] * 2, 0.35)
    with pytest.raises(RuntimeError, match="Invalid spring constants"):
        T.CLEulerUpdate(1, 0.01)
    nanc = m.Cell3D([0.0, 0.0, 1.0], 1.0, 1.0)
    nanc.Kv, nanc.Ka, nanc.Ks = 5.0, 2.0, 3.0
    v = nanc.Verts
    v[5] = [float("nan"), 0.0, 0.0]
    nanc.Verts = v
    T = m.Tissue3D([nanc] * 2, 0.35)
    with pytest.raises(RuntimeError, match="Non-finite vertex coordinates"):
        T.CLEulerUpdate(1, 0.01)
    if not has_gpu():
        T = m.Tissue3D([c] * 2, 0.35)
        with pytest.raises(RuntimeError):  # no device -> loud failure, never a CPU computation
            T.CLEulerUpdate(1, 0.01)


def test_disperse_separates_cells_and_keeps_shapes():
    m = H.cldpm()
    c = m.Cell2D(0.0, 0.0, 1.05, 32, 1.0)
    T = m.Tissue2D([c] * 16, 0.85)
    T.Disperse()
    cells = T.Cells
    ctr = np.array([np.asarray(x.Verts).mean(0) for x in cells])
    d = ctr[:, None, :] - ctr[None, :, :]
    d -= T.L * np.round(d / T.L)
    dist = np.sqrt((d ** 2).sum(-1)) + 10 * np.eye(16)
    assert dist.min() > 1.5  # soft discs of radius r0 relaxed apart (contact distance 2*r0)
    r = np.linalg.norm(np.asarray(cells[3].Verts) - ctr[3], axis=1)
    np.testing.assert_allclose(r, 1.0, rtol=1e-5)
    c3 = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
    T3 = m.Tissue3D([c3] * 8, 0.35)
    z0 = np.asarray(T3.Cells[0].Verts)[:, 2].copy()
    T3.Disperse2D()
    assert np.array_equal(np.asarray(T3.Cells[0].Verts)[:, 2], z0)  # z untouched (src/Tissue3D.cpp:106-115)


def test_binned_centre_relaxation_equals_all_pairs_bit_for_bit():
    """Disperse()/Disperse2D(): Verlet lists over a bin grid vs the reference's all-pairs loop restated next to it
    (disperse.hpp) — mixed radii, sizes beyond what the real reference finishes quickly; identical floats and the same
    stopping iteration.  (tests/test_golden_cpu.py checks the binned form against the real reference on smaller tissues.)"""
    import time

    import helpers as H

    m = H.cldpm()
    for n, r, phi in ((100, 1.0, 0.85), (500, 1.0, 0.85), (300, 2.0, 0.5)):
        L = float(np.sqrt(n * np.pi * r * r) / phi)
        rad = [r if i % 2 == 0 else 1.3 * r for i in range(n)]
        H.reset_drand48()
        t0 = time.time()
        Xa, Ya, capa = m._relax_centres(rad, L, False)
        ta = time.time() - t0
        H.reset_drand48()
        t0 = time.time()
        Xb, Yb, capb = m._relax_centres(rad, L, True)
        tb = time.time() - t0
        assert capa == capb
        assert np.array_equal(np.asarray(Xa, np.float32), np.asarray(Xb, np.float32))
        assert np.array_equal(np.asarray(Ya, np.float32), np.asarray(Yb, np.float32))
        print(f"n={n}: binned {ta:.2f} s, all-pairs {tb:.2f} s")


def test_disperse_4096_cells_finishes():
    """BASELINE config B/D sizes were out of reach of the reference's own initialiser (O(N^2) x up to 1e5 iterations)."""
    import time

    import helpers as H

    m = H.cldpm()
    c = m.Cell2D(0.0, 0.0, 1.2, 16, 1.0)
    T = m.Tissue2D([c] * 4096, 0.85)
    H.reset_drand48()
    t0 = time.time()
    T.Disperse()
    dt = time.time() - t0
    X = np.array([np.asarray(x.Verts, np.float32).mean(0) for x in T.Cells])
    assert np.isfinite(X).all() and dt < 120
    print(f"Disperse of 4096 cells: {dt:.1f} s")


def test_cell3d_subdivision_levels():
    """Extension: Cell3D(start, calA, r0, subdivisions).  Level 2 is the reference mesh and the
    3-argument constructor; level 3 is the 642-vertex mesh of BASELINE configs D/E; sizes follow V = F/2 + 2."""
    import helpers as H
    from opencl_dpm_b200 import capi

    m = H.cldpm()
    ref = m.Cell3D([1.0, 2.0, 3.0], 1.05, 1.8)
    same = m.Cell3D([1.0, 2.0, 3.0], 1.05, 1.8, 2)
    assert ref.NV == 162 and ref.NF == 320 and ref.subdivisions == 2
    assert np.array_equal(np.asarray(ref.Verts, np.float32), np.asarray(same.Verts, np.float32)) and ref.GetFaces() == same.GetFaces()
    for s, nv, nf in ((0, 12, 20), (1, 42, 80), (3, 642, 1280)):
        c = m.Cell3D([0.0, 0.0, 1.0], 1.0, 1.0, s)
        assert (c.NV, c.NF) == (nv, nf)
        P = c.GetPositions()
        assert len(P) == 3 and len(P[0]) == nv and len(c.GetFaces()) == nf and len(c.GetForces()[0]) == nv
        u, f = capi.icosphere(s)
        assert np.array_equal(np.asarray(c.Verts, np.float32), (u + np.array([0, 0, 1], np.float32)).astype(np.float32))
        assert np.array_equal(np.asarray(c.GetFaces()), f)
        assert 0.6 * 4.18879 < c.GetVolume() <= 4.18879 + 1e-3  # inscribed polyhedron of the unit sphere
    with pytest.raises(ValueError):
        m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0, 4)


def test_tissue3d_rejects_mixed_meshes():
    import helpers as H

    m = H.cldpm()
    a, b = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0, 2), m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0, 3)
    for c in (a, b):
        c.Kv, c.Ka, c.Ks = 1.0, 1.0, 1.0
    T = m.Tissue3D([a, b], 0.35)
    T.Kre = 1.0
    with pytest.raises(RuntimeError, match="share one mesh"):
        T.CLEulerUpdate(1, 0.01)


def test_trajectory_writer_roundtrip(tmp_path):
    """AppendFrame (host/trajectory.cpp) + opencl_dpm_b200.traj.read: the flat binary replacement of the reference's
    PNG-per-frame loop."""
    import helpers as H
    from opencl_dpm_b200 import traj

    m = H.cldpm()
    p3 = str(tmp_path / "t3.dpmt")
    c = m.Cell3D([0.0, 0.0, 1.0], 1.0, 1.0, 3)
    T = m.Tissue3D([c] * 5, 0.35)
    H.reset_drand48()
    T.AppendFrame(p3)
    T.Disperse2D()
    T.AppendFrame(p3)
    r = traj.read(p3)
    assert r["frames"].shape == (2, 5, 642, 3) and r["dim"] == 3 and r["PBC"] == 1 and abs(r["L"] - T.L) < 1e-6
    assert np.array_equal(r["faces"], np.asarray(T.Cells[0].GetFaces()))
    assert np.array_equal(r["frames"][1, 3], np.asarray(T.Cells[3].Verts, np.float32))
    assert np.array_equal(r["frames"][0, 3], np.asarray(c.Verts, np.float32))
    p2 = str(tmp_path / "t2.dpmt")
    a, b = m.Cell2D(0.0, 0.0, 1.2, 25, 1.0), m.Cell2D(1.0, 1.0, 1.2, 22, 1.3)
    T2 = m.Tissue2D([a, b] * 3, 0.9)
    T2.AppendFrame(p2)
    r2 = traj.read(p2, mmap=False)
    assert r2["frames"].shape == (1, 6, 25, 2) and list(r2["NV"]) == [25, 22] * 3
    assert np.array_equal(r2["frames"][0, 1, :22], np.asarray(T2.Cells[1].Verts, np.float32)) and not r2["frames"][0, 1, 22:].any()
    with pytest.raises(RuntimeError):
        T.AppendFrame(p2)  # a 3D tissue into a 2D file


def test_threaded_pack_keeps_the_reference_validation_errors():
    """Tissues above ~2e5 vertices are packed by several threads (host/Tissue3D.cpp:parallel_cells); the threads only run
    the silent happy path, so a bad input must still raise the reference's error for the FIRST offending cell
    (src/Tissue3D.cpp:157-190), from the serial pass."""
    import helpers as H

    m = H.cldpm()
    c = m.Cell3D([0.0, 0.0, 0.0], 1.0, 1.0)
    c.Ka, c.Kv, c.Ks = 2.0, 5.0, 3.0
    n = 1400  # x 162 vertices = 226,800 > the threading threshold
    T = m.Tissue3D([c] * n, 0.35)
    T.Kre = 25.0
    cells = T.Cells
    v = np.asarray(cells[900].Verts, np.float32)
    v[17, 1] = np.nan
    cells[900].Verts = v.tolist()
    cells[1200].Kv = -1.0  # a later problem of another kind: must not be the one reported
    T.Cells = cells
    with pytest.raises(RuntimeError, match="Non-finite vertex coordinates"):
        T.CLEulerUpdate(1, 0.01)
    cells[900].Verts = np.nan_to_num(v).tolist()
    T.Cells = cells
    with pytest.raises(RuntimeError, match="Invalid spring constants"):
        T.CLEulerUpdate(1, 0.01)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's all-pairs algorithm on the host cores, oracle port): one JSON line on
    stdout with the keys the driver reads; non-zero ranks print nothing and exit 0."""
    import json
    import subprocess
    import sys

    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C162", "--steps", "1", "--warmup", "0", "--no-opencl"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "vertex-steps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "vertex-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["name"] == "C162"
    assert d["cpu_baseline_culled"]["value"] > d["value"] and d["cpu_baseline_culled"]["kind"] == "port"
    # the CPU arm never maps the product library: its tissue comes from the oracle's own geometry helpers
    assert d["native_so_loaded"] is not None and all(x.startswith("oracle/") for x in d["native_so_loaded"]), d["native_so_loaded"]
    assert d["gpu_launches"] == 0
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                        timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_bench_reference_arm_is_bounded_on_config_e():
    """Config E (262,144 cells) is the default workload of the N > 1 runs: the CPU arm must answer in about a minute (round 1
    timed out there), from a bounded block of the lattice scaled to the whole tissue, and only on rank 0."""
    import json
    import subprocess
    import sys
    import time

    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8", "--workload", "E162", "--steps", "1",
                        "--warmup", "0", "--no-opencl"], capture_output=True, text=True, timeout=240,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="8"))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["impl"] == "reference" and d["config"]["name"] == "E162" and d["n_gpus"] == 8 and d["value"] > 0
    assert "scaled" in d["cpu_baseline"]["sample"] and time.time() - t0 < 200
