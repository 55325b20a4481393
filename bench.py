#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native OpenCL_DPM hot path.

Metric (BASELINE.json): vertex-steps/s of the fused force+integrate step, plus the fraction
of the HBM roofline at 32 B (3D) / 16 B (2D) algorithmic bytes per vertex-step.

A bench "step" = one CLEulerUpdate-equivalent pass over the tissue: `--inner` timesteps of
{forces, Euler}.  `value` is measured with inputs resident in HBM (CUDA events on the stream
the kernels are launched on, around dpm3d_step only — what the reference's own timer brackets,
src/Tissue3D.cpp:369,454).  `e2e` is the same metric through the reference-facing one-call
seam dpm3d_euler_update with PINNED HOST buffers: H2D of the vertices and per-cell parameters,
the step loop, D2H of vertices and last-step forces, wall-clock.

The same JSON line also carries (single GPU):
  trajectory        one device-resident run of >= 2000 timesteps from the RAW synthetic lattice, per phase
                    (the headline batch is one state of this trajectory; SURVEY §8d asks for nsteps >= 1000)
  other_configs     short lines of BASELINE configs A, B, C (and D at the reference's own 162-vertex mesh)
  e2e_host_classes  Tissue3D.CLEulerUpdate on a 4096-cell Cells vector (pack + C call + unpack): the drop-in call
and at N > 1: sharded_check (slab-sharded run == single-GPU run, bit for bit) and single_gpu_same_problem.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload D642|D162|C162|C642|B2D|A2D|E642|E162] [--inner T]
  python bench.py --impl reference ...   # the reference algorithm on the host cores (oracle arm, CPU only)
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PK3 = ("Kv", "Ka", "Ks", "v0", "a0", "l0")
PK2 = ("Ka", "Kl", "Kb", "a0", "l0", "r0")

DESC = {
    "E642": "3D DPM 262,144-cell monolayer, 642-vertex icospheres (168,296,448 vertices), x-slab decomposed with per-step halo exchange",
    "E162": "3D DPM 262,144-cell monolayer, 162-vertex icospheres (42,467,328 vertices), x-slab decomposed with per-step halo exchange",
    "D642": "3D DPM 4096-cell monolayer, 642-vertex icospheres (2,629,632 vertices), winding-number repulsion, substrate, PBC",
    "D162": "3D DPM 4096-cell monolayer, 162-vertex icospheres (663,552 vertices; the reference's mesh)",
    "C162": "test3D: 3D DPM 64 cells x 162-vertex icospheres, reference test3D.py parameters, Disperse2D() placement",
    "C642": "test3D: 3D DPM 64 cells x 642-vertex icospheres, reference test3D.py parameters, Disperse2D() placement",
    "B2D": "2D DPM 4096 cells x 64 vertices, area+perimeter+bending+attraction+repulsion, PBC",
    "A2D": "test2D: 2D DPM 32 cells x 32 vertices, reference test2D.cpp parameters, Disperse() placement, periodic box",
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def native_so_loaded():
    """in-tree shared objects mapped into this process (evidence of which native code ran)"""
    try:
        with open("/proc/self/maps") as f:
            libs = {line.split()[-1] for line in f if ".so" in line and ROOT in line}
        return sorted(os.path.relpath(x, ROOT) for x in libs)
    except Exception:
        return None


def load_synth():
    """opencl_dpm_b200/synth.py by file path: the CPU reference arm builds its tissues without importing the product
    package (so that libdpm_b200.so is never mapped into that process)."""
    spec = importlib.util.spec_from_file_location("dpm_synth", os.path.join(ROOT, "opencl_dpm_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_workload(name: str, rank: int = 0, world: int = 1, geom=None, synth=None):
    if synth is None:
        from opencl_dpm_b200 import synth
    if name in ("E642", "E162"):
        # BASELINE config E: 512 x 512 = 262,144 cells; each rank builds only its x-slab of lattice columns
        from opencl_dpm_b200 import shard

        i0, i1 = shard.slab_columns(512, rank, world)
        d = synth.monolayer3d(512, subdiv=3 if name == "E642" else 2, x_range=(i0, i1), geom=geom)
        d["nc_global"] = 512 * 512
    elif name in ("D642", "D162"):
        d = synth.monolayer3d(64, subdiv=3 if name == "D642" else 2, geom=geom)
    elif name in ("C162", "C642"):
        d = synth.test3d_config(64, subdiv=2 if name == "C162" else 3)
    elif name == "B2D":
        d = synth.tissue2d(64, nv=64)
    elif name == "A2D":
        d = synth.test2d_config(32)
    else:
        raise SystemExit(f"unknown workload {name}")
    d["name"], d["desc"], d["dim"] = name, DESC[name], (2 if name.endswith("2D") else 3)
    d.setdefault("nc_global", d["nc"])
    return d


class stdout_to_stderr:
    """NCCL prints its version banner on fd 1 when the first communicator is created and the drop-in classes print the
    reference's two timing lines; the contract is ONE JSON line on stdout, so fd 1 is pointed at stderr meanwhile."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# =====================================================================================================================
# CPU arms (oracle port).  Only here — and in tests/ and smoke() — is anything under oracle/ executed.
# =====================================================================================================================
EVALS_PER_CORE_S = 1.7e7  # solid-angle evaluations per second and host core of the fp32 port (measured, round 1)
_SAMPLE_K = {}            # vertices per sample of the all-pairs arm, adapted to the measured rate (per workload)


def cpu_allpairs_sample(d, budget_s: float):
    """The reference's own ALL-PAIRS algorithm (CPU port of the OpenCL kernels, fp32, OpenMP over all host cores) on a
    BOUNDED sample of the workload, sized from `budget_s`:
      3D: RepellingForces (> 99 % of the reference's step) of k vertices of cell 0 against every cell present;
          the cost of a (vertex, cell) pair is the same for every pair, so the sample's rate is the workload's rate.  For
          config E only a block of the tissue is built and the per-vertex time is scaled to all 262,143 partner cells.
      2D: all six kernels for the first k cells against all cells."""
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    if d["dim"] == 3:
        nc, nv, nf = d["nc"], d["nv"], d["nf"]
        scale = (d["nc_global"] - 1) / max(1, nc - 1)  # partner cells of the whole tissue per partner cell present
        # the sample size adapts to the host's real rate (whatever its cores / OpenMP settings are): it starts with one vertex
        # and is rescaled after every sample towards `budget_s` seconds of work
        k = int(max(1, min(nv, _SAMPLE_K.get(d["name"], 1))))
        t0 = time.perf_counter()
        O.repel_sample3d(d["verts"], d["faces"], nc, d["Kre"], d["PBC"], d["L"], 0, 0, k)
        dt = time.perf_counter() - t0
        _SAMPLE_K[d["name"]] = max(1, min(nv, int(k * min(8.0, budget_s / max(dt, 1e-3)))))
        sample = (f"reference all-pairs algorithm (RepellingForces, > 99 % of its step): {k} vertices of cell 0 against "
                  f"{nc - 1} cells, 1 timestep, {dt:.1f} s" + (f"; per-vertex time scaled x{scale:.1f} to the tissue's "
                                                               f"{d['nc_global'] - 1} partner cells" if scale > 1.0001 else ""))
        return {"value": k / (dt * scale), "unit": "vertex-steps/s", "cores": cores, "kind": "port", "sample": sample}
    per_cell = float(d["nv"].mean()) ** 2 * d["nc"] * 2 / (2.5e8 * cores)
    k = int(max(1, min(d["nc"], budget_s / max(per_cell, 1e-9))))
    t0 = time.perf_counter()
    O.forces2d_range(d["verts"], d["nv"], *[d[k2] for k2 in PK2], d["Kre"], d["Kat"], d["PBC"], d["L"], 0, k)
    dt = time.perf_counter() - t0
    return {"value": float(d["nv"][:k].sum()) / dt, "unit": "vertex-steps/s", "cores": cores, "kind": "port",
            "sample": f"reference all-pairs algorithm: all six kernels for the first {k} of {d['nc']} cells against all cells, 1 timestep, {dt:.1f} s"}


def cpu_culled_sample(d):
    """The same kernels behind the CPU restatement of the cell list (culled form = the product's algorithmic complexity),
    whole tissue, one timestep including the list build.  (Config E: the caller passes a 4096-cell block — the per-vertex
    work of the lattice does not depend on the tissue size.)"""
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    if d["dim"] == 3:
        V = d["verts"].reshape(d["nc"], d["nv"], 4)
        f = d["faces"]
        emax = max(float(np.linalg.norm(V[:, f[:, i], :3] - V[:, f[:, (i + 1) % 3], :3], axis=2).max()) for i in range(3))
        lo, hi = O.aabb3d(d["verts"], d["nc"])
        cl = O.cell_list(3, lo, hi, d["PBC"], d["L"], 0.1, 1.25 * 0.34 * emax, 32)
        O.forces3d(d["verts"], d["faces"], *[d[k] for k in PK3], d["Kre"], d["PBC"], d["L"], cand_count=cl["cand_count"], cand=cl["cand"])
        nvert = d["nc"] * d["nv"]
    else:
        lo, hi = O.aabb2d(d["verts"], d["nv"])
        cl = O.cell_list(2, lo, hi, d["PBC"], d["L"], 0.1, float(d["l0"].max()) if d["Kat"] != 0 else 0.0, 64, far2d=True)
        O.forces2d(d["verts"], d["nv"], *[d[k] for k in PK2], d["Kre"], d["Kat"], d["PBC"], d["L"], cand_count=cl["cand_count"], cand=cl["cand"])
        nvert = int(d["nv"].sum())
    dt = time.perf_counter() - t0
    return {"value": nvert / dt, "unit": "vertex-steps/s", "cores": cores, "kind": "port",
            "sample": f"culled form (CPU cell list + literal kernels), {d['nc']} cells, 1 timestep incl. list build ({dt:.1f} s)"}


def run_reference_arm(args):
    """`--impl reference`: the reference's algorithm on the box's host cores.  CPU only: the tissue is built with the
    oracle's own geometry helpers and the product package is never imported.  Every step is one bounded sample; the
    whole run takes about a minute whatever the workload (config E included)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is the ONE process that runs, on all the host's cores.  The
    # variable is read when libgomp is loaded (with the oracle library, below), so it is set first.
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import oracle as O

    synth = load_synth()
    name = args.workload
    t_start = time.perf_counter()
    if name.startswith("E"):  # a block of 16 x 512 cells of the 512 x 512 lattice; the sample is scaled to the whole tissue
        d = synth.monolayer3d(512, subdiv=3 if name == "E642" else 2, x_range=(0, 16), geom=O)
        d["nc_global"] = 512 * 512
        block = "16 x 512-cell block of the lattice built on the host; "
    elif name in ("D642", "D162"):
        d, block = synth.monolayer3d(64, subdiv=3 if name == "D642" else 2, geom=O), ""
    elif name in ("C162", "C642"):  # Disperse2D() lives in the product's host classes: the CPU arm takes the lattice stand-in
        d, block = synth.monolayer3d(8, subdiv=2 if name == "C162" else 3, geom=O), "8 x 8 lattice stand-in for Disperse2D(); "
    elif name == "B2D":
        d, block = synth.tissue2d(64, nv=64), ""
    elif name == "A2D":
        d, block = synth.tissue2d(6, ny=6, nv=32, calA=1.05, Ka=1.0, Kl=1.0, Kb=0.1, Kre=50.0, Kat=0.0), "6 x 6 lattice stand-in for Disperse(); "
    else:
        raise SystemExit(f"unknown workload {name}")
    d["name"], d["desc"], d["dim"] = name, DESC[name], (2 if name.endswith("2D") else 3)
    d.setdefault("nc_global", d["nc"])
    nsamp = max(1, args.warmup + args.steps)
    budget = max(1.0, min(8.0, 40.0 / nsamp))
    vals, cb = [], None
    for _ in range(4):  # untimed: let the sample size settle (each at most `budget` seconds once settled)
        cpu_allpairs_sample(d, budget) if d["dim"] == 3 else None
    for i in range(nsamp):
        cb = cpu_allpairs_sample(d, budget)
        if i >= args.warmup:
            vals.append(cb["value"])
        if time.perf_counter() - t_start > 60.0:  # never run into the driver's limit
            break
    if not vals:
        vals = [cb["value"]]
    v = float(np.mean(vals))
    cb = dict(cb, value=v, sample=block + cb["sample"])
    if name.startswith("E"):
        dc = synth.monolayer3d(64, subdiv=3 if name == "E642" else 2, geom=O)
        dc["dim"] = 3
    else:
        dc = d
    culled = cpu_culled_sample(dc)
    if name.startswith("E"):
        culled["sample"] = "4096-cell lattice block (per-vertex work of the lattice does not depend on the tissue size): " + culled["sample"]
    extra = {}
    if not args.no_opencl:
        try:  # the REAL reference (its own host code + OpenCL kernels, oracle/_ref) when an OpenCL device is reachable
            from oracle import ref as R

            if R.available():
                d64 = synth.monolayer3d(8, subdiv=2, geom=O)
                with stdout_to_stderr():  # the reference prints its own timing lines
                    _, _, sec = R.euler3d(d64["verts"], d64["Kv"], d64["Ka"], d64["Ks"], d64["v0"], d64["a0"], d64["Kre"], 1, d64["L"], 2, d64["dt"])
                extra["reference_opencl"] = {
                    "device": R.device_name(), "workload": "64-cell monolayer x 162 vertices (the reference hard-codes NV=162), 2 timesteps, "
                    "whole CLEulerUpdate call incl. its per-call JIT build", "seconds": sec, "vertex_steps_per_s": 64 * 162 * 2 / sec,
                    "note": "all-pairs contact kernel: cost per vertex-step grows linearly with the cell count (x64 at 4096 cells)"}
        except Exception as e:  # never let the informational leg break the arm
            extra["reference_opencl"] = {"unavailable": str(e)[:200]}
    out = {"impl": "reference", "metric": "vertex-steps/sec (force+integrate)", "value": v, "unit": "vertex-steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": None, "higher_is_better": True, "scaling": "strong" if name.startswith("E") and args.gpus > 1 else "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": d["desc"], "name": d["name"], "algorithm": "reference all-pairs (CPU port of the OpenCL kernels, oracle/)",
                      "note": "the reference has no neighbour search: its cost per vertex-step grows with the cell count, the product's does not; "
                              "cpu_baseline_culled is the like-for-like CPU sibling (same algorithmic complexity as the product)"},
           "cpu_baseline": cb, "cpu_baseline_culled": culled,
           "e2e": {"value": v, "unit": "vertex-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.perf_counter() - t_start, "native_so_loaded": native_so_loaded(), **extra}
    print(json.dumps(out), flush=True)


# =====================================================================================================================
# the product arm
# =====================================================================================================================
class Runner:
    """One workload on one handle: device-resident stepping, re-upload of a batch, the e2e seam."""

    def __init__(self, d, local, stream, rank=0, world=1, sharded=False, max_ghost=None):
        import torch

        from opencl_dpm_b200 import Dpm2D, Dpm3D

        self.d, self.dim, self.torch = d, d["dim"], torch
        self.stream = stream
        self.nvert = d["nc"] * d["nv"] if self.dim == 3 else int(d["nv"].sum())
        self.balg = 32 if self.dim == 3 else 16
        if self.dim == 3:
            h = Dpm3D(d["nc"], d["nv"], d["faces"], device=local)
            if sharded:
                from opencl_dpm_b200 import shard

                h.shard_init(rank, world, shard.broadcast_unique_id(rank), max_ghost=max_ghost)
                h.set_global_ids(d["gid"])
            h.set_stream(stream.cuda_stream)
            self.params = [d[k] for k in PK3]
            self.dev_verts = torch.from_numpy(d["verts"]).cuda()
            self.h2d = d["verts"].nbytes + 6 * 4 * d["nc"]
        else:
            h = Dpm2D(d["nc"], d["S"], device=local)
            h.set_neighbor_params(0.1, 64)
            h.set_stream(stream.cuda_stream)
            self.params = [d[k] for k in PK2]
            self.h2d = d["verts"].nbytes + (6 * 4 + 4) * d["nc"]
        self.d2h = 2 * d["verts"].nbytes
        self.h = h

    def reset(self, src=None):
        d = self.d
        if self.dim == 3:
            self.h.upload_device((self.dev_verts if src is None else src).data_ptr(), *self.params)
        else:
            self.h.upload(d["verts"] if src is None else src, d["nv"], *self.params)

    def run_steps(self, n):
        d = self.d
        if self.dim == 3:
            self.h.step(n, float(d["dt"]), float(d["Kre"]), 0.0, d["PBC"], float(d["L"]))
        else:
            self.h.step(n, float(d["dt"]), float(d["Kre"]), float(d["Kat"]), d["PBC"], float(d["L"]))

    def snapshot(self):
        """the current state in the form reset(src=...) re-uploads (3D: device tensor; the 2D ABI uploads from host memory)"""
        if self.dim == 3:
            return self.torch.from_numpy(self.h.download(want_forces=False)[0]).cuda()
        return self.h.download(want_forces=False)[0].copy()

    def e2e_buffers(self):
        torch = self.torch
        self.host_v = torch.from_numpy(self.d["verts"].copy()).pin_memory()
        self.host_f = torch.zeros_like(self.host_v).pin_memory()

    def e2e_call(self, n):
        d = self.d
        t0 = time.perf_counter()
        if self.dim == 3:
            self.h.euler_update(self.host_v.numpy(), *self.params, n, float(d["dt"]), float(d["Kre"]), 0.0, d["PBC"], float(d["L"]),
                                forces_out=self.host_f.numpy())
        else:
            self.h.euler_update(self.host_v.numpy(), d["nv"], *self.params, n, float(d["dt"]), float(d["Kre"]), float(d["Kat"]), d["PBC"],
                                float(d["L"]), forces_out=self.host_f.numpy())
        return time.perf_counter() - t0

    def stats_dict(self):
        st = self.h.stats()
        return {"steps": int(st.steps), "rebuilds": int(st.rebuilds), "contact_evals": int(st.contact_evals), "literal": int(st.reserved[0]), "nonstar": int(st.reserved[1] & 0xffffffff),
                "halo_bytes": int(st.halo_bytes), "launches": int(st.launches)}

    def timed_windows(self, batch, inner, steps, warmup, flush, barrier=None):
        """`steps` windows of `inner` timesteps, each restarted from `batch` (re-uploaded before the event pair, L2 flushed);
        returns (total ms over the windows, launches inside them, per-window stats of the last window)."""
        torch = self.torch
        for _ in range(warmup):
            self.reset(batch)
            self.run_steps(inner)
        torch.cuda.synchronize()
        l0 = self.h.stats().launches
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if barrier:
            barrier()
        t0 = time.perf_counter()
        for a, b in evs:
            self.reset(batch)   # this window's input batch becomes resident (outside the event pair); its first timestep rebuilds the lists
            flush.fill_(1)      # flush L2 between timed windows (outside the event pair)
            a.record(self.stream)
            self.run_steps(inner)
            b.record(self.stream)
        if barrier:
            barrier()
        else:
            torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        self.h.sync()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        st = self.stats_dict()  # counters are reset by every upload: these are the LAST window's (timed region only)
        launches = st["launches"] - l0 - steps  # minus the bounds kernel of each window's (untimed) batch upload
        return ms, launches, st, wall

    def close(self):
        self.h.close()


def trajectory(R, phases, peak):
    """ONE device-resident run from the raw synthetic lattice, timed per phase with CUDA events on the launching stream
    (no re-upload, no L2 flush: consecutive timesteps as a caller's loop runs them)."""
    torch = R.torch
    R.reset()
    R.h.sync()
    prev = R.stats_dict()
    out, t_total, s0 = [], 0.0, 0
    for s1 in phases:
        n = s1 - s0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(R.stream)
        R.run_steps(n)
        b.record(R.stream)
        R.h.sync()
        ms = a.elapsed_time(b)
        st = R.stats_dict()
        vs = R.nvert * n / (ms * 1e-3)
        out.append({"timesteps": [s0, s1], "ms_per_timestep": ms / n, "vertex_steps_per_s": vs,
                    "roofline_frac": vs * R.balg / 1e9 / peak,
                    "contact_evals_per_timestep": (st["contact_evals"] - prev["contact_evals"]) / n,
                    "literal_fallback_evals_per_timestep": (st["literal"] - prev["literal"]) / n,
                    "nonstar_evals_per_timestep": (st["nonstar"] - prev["nonstar"]) / n,
                    "rebuilds": st["rebuilds"] - prev["rebuilds"]})
        prev, s0, t_total = st, s1, t_total + ms
    vs = R.nvert * phases[-1] / (t_total * 1e-3)
    return {"from": "raw jittered lattice (~5 % overlap everywhere), one upload, device-resident", "timesteps": phases[-1], "ms_total": t_total,
            "ms_per_timestep": t_total / phases[-1], "vertex_steps_per_s": vs, "roofline_frac": vs * R.balg / 1e9 / peak, "phases": out}


def short_line(name, local, stream, flush, peak, equilibrate, inner=25, steps=3, warmup=3):
    """A compact bench line of another BASELINE config under the headline protocol."""
    d = make_workload(name)
    R = Runner(d, local, stream)
    R.reset()
    if equilibrate > 0:
        R.run_steps(equilibrate)
    R.torch.cuda.synchronize()
    batch = R.snapshot()
    ms, launches, st, _ = R.timed_windows(batch, inner, steps, warmup, flush)
    n = inner * steps
    vs = R.nvert * n / (ms * 1e-3)
    R.close()
    return {"name": name, "workload": d["desc"], "vertices": R.nvert, "value": vs, "unit": "vertex-steps/s", "ms_per_timestep": ms / n,
            "roofline_frac": vs * R.balg / 1e9 / peak, "timesteps_per_step": inner, "steps": steps, "equilibrate": equilibrate,
            "contact_evals_per_timestep": st["contact_evals"] / max(1, st["steps"]),
            "literal_fallback_evals_per_timestep": st["literal"] / max(1, st["steps"]),
            "nonstar_evals_per_timestep": st["nonstar"] / max(1, st["steps"])}


def e2e_host_classes(batch_host, d, inner, steps):
    """The drop-in call itself: Tissue3D.CLEulerUpdate(inner, dt) on a Cells vector holding the bench batch — AoS pack,
    the C ABI call, unpack with the reference's checks (src/Tissue3D.cpp:118-522 is what it replaces).  Wall clock."""
    import opencl_dpm_b200 as pkg

    m = pkg.load_cldpm()
    nc, nv = d["nc"], d["nv"]
    sub = {12: 0, 42: 1, 162: 2, 642: 3}[nv]
    c = m.Cell3D([0.0, 0.0, 1.0], 1.0, 1.0, sub)
    c.Kv, c.Ka, c.Ks = float(d["Kv"][0]), float(d["Ka"][0]), float(d["Ks"][0])
    vsum = float(nc) * float(d["v0"][0])
    with stdout_to_stderr():
        T = m.Tissue3D([c] * nc, float(np.cbrt(vsum) / float(d["L"])))  # L = cbrt(sum v0) / phi0 (src/Tissue3D.cpp:25-29); L is read-only from Python
    T.Kre = float(d["Kre"])
    V = batch_host.reshape(nc, nv, 4)[:, :, :3]
    cells = T.Cells
    for i, x in enumerate(cells):
        x.Verts = V[i].tolist()
    times = []
    for _ in range(steps + 1):
        T.Cells = cells  # the caller's Cells hold the batch again (untimed)
        with stdout_to_stderr():
            t0 = time.perf_counter()
            T.CLEulerUpdate(inner, float(d["dt"]))
            times.append(time.perf_counter() - t0)
    times = times[1:]  # the first call also creates the device context
    t = float(np.mean(times))
    return {"value": nc * nv * inner / t, "unit": "vertex-steps/s", "ms_per_call": t * 1e3, "calls": len(times),
            "call": f"clDPM.Tissue3D.CLEulerUpdate({inner}, {float(d['dt'])}) on {nc} Cell3D objects (AoS pack + dpm3d_euler_update + unpack with the reference's checks)",
            "box_L": float(T.L)}


def sharded_check(rank, world, local, stream):
    """A slab-sharded run must reproduce the single-GPU run of the same tissue BIT FOR BIT: 32 x 32 cells x 162 vertices,
    30 timesteps from the overlapping lattice (contact-dominated), all ranks; rank 0 repeats it alone and compares."""
    import torch
    import torch.distributed as dist

    from opencl_dpm_b200 import shard, synth

    nx = 32 if 32 % world == 0 else 4 * world  # equal slabs of >= 4 lattice columns
    i0, i1 = shard.slab_columns(nx, rank, world)
    d = synth.monolayer3d(nx, 32, subdiv=2, x_range=(i0, i1))
    d["dim"] = 3
    R = Runner(d, local, stream, rank, world, sharded=True, max_ghost=256)
    nsteps = 30
    R.reset()
    R.run_steps(nsteps)
    V, F = R.h.download()
    halo = R.stats_dict()["halo_bytes"]
    R.close()
    mine = torch.from_numpy(np.concatenate([V, F], 1)).cuda()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    res = None
    if rank == 0:
        full = synth.monolayer3d(nx, 32, subdiv=2)
        full["dim"] = 3
        R1 = Runner(full, local, stream)
        R1.reset()
        R1.run_steps(nsteps)
        V1, F1 = R1.h.download()
        R1.close()
        got = torch.cat(parts).cpu().numpy()  # slabs are contiguous column ranges: concatenation is global-id order
        same = bool(np.array_equal(got[:, :4], V1) and np.array_equal(got[:, 4:], F1))
        res = {"tissue": f"{nx} x 32 cells x 162 vertices, {nsteps} timesteps from the overlapping lattice", "ranks": world,
               "bit_identical_to_single_gpu": same, "max_abs_force": float(np.abs(F1).max()), "halo_bytes_rank0": int(halo),
               "max_abs_position_diff": float(np.abs(got[:, :4] - V1).max())}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="D642 (default at 1 GPU), D162, C162, C642, B2D, A2D, E642 (default at >1 GPU), E162")
    ap.add_argument("--inner", type=int, default=None, help="timesteps per bench step (one CLEulerUpdate call)")
    ap.add_argument("--equilibrate", type=int, default=None,
                    help="untimed timesteps from the synthetic lattice that produce the batch every bench step processes (default 100). "
                         "The whole trajectory from the raw lattice is reported per phase in `trajectory`")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip trajectory / other_configs / e2e_host_classes / sharded_check (profiling runs)")
    ap.add_argument("--no-opencl", action="store_true", help="reference arm: skip the informational run of the real reference through OpenCL")
    ap.add_argument("--trajectory-steps", type=int, default=2000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        # 1 GPU: BASELINE config D (4096 cells, the north-star roofline target).  N > 1: config E (262,144 cells) is
        # strong-scaled over the ranks; the same line carries the single-GPU rate of the SAME problem (single_gpu_same_problem).
        args.workload = "D642" if world_env == 1 else "E642"
    if args.inner is None:
        # timesteps per CLEulerUpdate-equivalent call: the reference's own 3D demo advances 25 per call (test3D.py:20-22)
        args.inner = 25 if not args.workload.startswith("E") else 10
    if args.equilibrate is None:
        args.equilibrate = 100

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch

    from opencl_dpm_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    redirect = stdout_to_stderr()
    redirect.__enter__()  # until the JSON line: NCCL banners and the drop-in classes' timing prints go to stderr
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    capi.lib()
    d = make_workload(args.workload, rank, world)
    dim = d["dim"]
    sharded = world > 1 and args.workload.startswith("E")
    stream = torch.cuda.Stream()  # a real (non-legacy) stream: the library launches on it and the events are recorded on it
    torch.cuda.set_stream(stream)
    # one lattice column (512 cells) per slab face + headroom; with 2 ranks the single peer is both neighbours
    R = Runner(d, local, stream, rank, world, sharded, max_ghost=(768 if world > 2 else 1280) if sharded else None)
    nvert, balg = R.nvert, R.balg
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) -----------------------------------------------
    # The batch every bench step processes: the synthetic lattice relaxed for `equilibrate` timesteps (untimed, all ranks).
    # Each step re-uploads THAT state (device to device, before the timed region: inputs resident in HBM when it starts)
    # and advances it `inner` timesteps, so a step's work does not depend on how many steps came before it.  The state is
    # one point of the trajectory from the raw lattice; the whole trajectory is timed per phase in `trajectory` below.
    R.reset()
    if args.equilibrate > 0:
        R.run_steps(args.equilibrate)
        torch.cuda.synchronize()
    batch = R.snapshot()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up: nvidia-smi needs ~0.3 s before its first sample; all samples are under load
    ms, launches, st, t_wall = R.timed_windows(batch, args.inner, args.steps, args.warmup, flush, barrier)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tv = torch.tensor([float(nvert)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tv, op=dist.ReduceOp.SUM)
        nvert_all = int(tv.item())  # all ranks' owned vertices (sharded: the whole tissue; replicas: world copies)
    else:
        nvert_all = nvert
    total_vs = nvert_all * args.inner * args.steps
    value = total_vs / (ms * 1e-3)

    # ---- end-to-end through the one-call seam with pinned host buffers (`e2e`) ---------------
    # Every call uploads the same batch from pinned host memory, advances it `inner` timesteps and reads positions and
    # forces back (host buffers in, host buffers out: what a caller of CLEulerUpdate does).
    R.e2e_buffers()
    batch_host = (batch.cpu() if torch.is_tensor(batch) else torch.from_numpy(batch)).view_as(R.host_v)

    def e2e_step():
        R.host_v.copy_(batch_host)  # untimed: the caller's buffer holds the batch again (the call updates it in place)
        return R.e2e_call(args.inner)
    for _ in range(min(2, args.warmup)):
        e2e_step()
    barrier()
    te = [e2e_step() for _ in range(args.steps)]
    barrier()
    e2e_t = float(np.sum(te))
    if rank == 0:
        print("[bench] e2e calls (ms): " + " ".join(f"{x * 1e3:.2f}" for x in te), file=sys.stderr)
    if world > 1:
        t = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    e2e_value = nvert_all * args.inner * args.steps / e2e_t

    # ---- extras (outside every timed region above) ---------------------------------------------------------------
    extras = {}
    if not args.no_extras:
        if world == 1 and dim == 3 and args.trajectory_steps >= 100:
            n = args.trajectory_steps
            phases = [p for p in (30, 300, 1000) if p < n] + [n]
            try:
                extras["trajectory"] = trajectory(R, phases, peak)
            except Exception as e:  # a capacity overflow deep in the crumpled regime must not lose the headline
                extras["trajectory"] = {"error": str(e)[:300]}
        if world == 1 and args.workload == "D642":
            others = []
            for nm, eq in (("A2D", 100), ("B2D", 100), ("C162", 100), ("C642", 100), ("D162", 100)):
                try:
                    others.append(short_line(nm, local, stream, flush, peak, eq))
                except Exception as e:
                    others.append({"name": nm, "error": str(e)[:300]})
            extras["other_configs"] = others
            try:
                extras["e2e_host_classes"] = e2e_host_classes(batch_host.numpy(), d, args.inner, min(3, args.steps))
            except Exception as e:
                extras["e2e_host_classes"] = {"error": str(e)[:300]}
        if sharded:
            try:
                extras["sharded_check"] = sharded_check(rank, world, local, stream)
            except Exception as e:
                extras["sharded_check"] = {"error": str(e)[:300]}
            # the SAME problem on one GPU (rank 0 alone, all of config E), so that the speed-up has a baseline measured in this run
            if rank == 0:
                try:
                    full = make_workload(args.workload, 0, 1)
                    R1 = Runner(full, local, stream)
                    R1.reset()
                    if args.equilibrate > 0:
                        R1.run_steps(args.equilibrate)
                    torch.cuda.synchronize()
                    b1 = R1.snapshot()
                    ms1, _, st1, _ = R1.timed_windows(b1, args.inner, max(2, min(3, args.steps)), 2, flush)
                    n1 = args.inner * max(2, min(3, args.steps))
                    v1 = R1.nvert * n1 / (ms1 * 1e-3)
                    extras["single_gpu_same_problem"] = {"value": v1, "unit": "vertex-steps/s", "ms_per_timestep": ms1 / n1,
                                                         "cells": int(full["nc"]), "protocol": "same batch protocol, rank 0 alone"}
                    extras["speedup_vs_1gpu"] = value / v1
                    R1.close()
                    del b1
                except Exception as e:
                    extras["single_gpu_same_problem"] = {"error": str(e)[:300]}
            barrier()

    redirect.__exit__()
    if rank == 0:
        n_step_kernels = args.inner * args.steps
        t_launch = ms * 1e-3 / n_step_kernels  # step-kernel launches dominate the region (the rebuild kernel is a no-op launch)
        achieved = nvert * balg / t_launch / 1e9  # per GPU: this rank's owned vertices per launch
        traffic, traffic_src = None, "no ncu --set full capture of this workload is committed"
        try:  # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f).get(d["name"], {})
                traffic = tj.get("bytes_per_launch")
                traffic_src = tj.get("source", traffic_src)
        except Exception:
            pass
        out = {
            "metric": "vertex-steps/sec (force+integrate)", "value": value, "unit": "vertex-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": d["desc"], "name": d["name"], "timesteps_per_step": args.inner, "dt": float(d["dt"]),
                       "parallelism": ("single GPU" if world == 1 else
                                       f"x-slab decomposition over {world} GPUs, ghost cells exchanged every timestep over NVLink (ring of slabs)" if sharded
                                       else f"{world} independent replicas (one tissue per GPU)"),
                       "cells_per_gpu": int(d["nc"]), "cells_total": int(d["nc_global"]) if sharded or world == 1 else int(d["nc"]) * world,
                       "l2": "L2 flushed (256 MiB write) between timed steps; within a step consecutive timesteps reuse L2 as in the real loop",
                       "ms_per_timestep": ms / n_step_kernels,
                       "state": f"jittered lattice relaxed for {args.equilibrate} untimed timesteps (stats.contact_evals_per_timestep: (vertex, neighbour) "
                                "evaluations per timestep inside the timed region); every bench step re-uploads that state device-to-device before its "
                                "timed region and advances it timesteps_per_step timesteps; `trajectory` times the whole run from the raw lattice"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "vertex-steps/s", "h2d_bytes_per_step": int(R.h2d), "d2h_bytes_per_step": int(R.d2h),
                    "ms_per_step": e2e_t * 1e3 / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "dpm3d_step_kernel" if dim == 3 else "dpm2d_step_kernel", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "algorithmic_bytes_per_vertex_step": balg, "us_per_launch": t_launch * 1e6},
            "stats": {"rebuilds": st["rebuilds"], "contact_evals_per_timestep": st["contact_evals"] / max(1, st["steps"]),
                      "wall_s_timed_region": t_wall, "halo_bytes_per_timestep_rank0": st["halo_bytes"] / max(1, st["steps"]),
                      "literal_fallback_evals_per_timestep": st["literal"] / max(1, st["steps"]),
                      "nonstar_evals_per_timestep": st["nonstar"] / max(1, st["steps"])},
            **extras,
            "native_so_loaded": native_so_loaded(),
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_culled_sample(d)                                   # same algorithmic complexity as the product
            out["cpu_baseline_reference_algorithm"] = cpu_allpairs_sample(d, 10.0)       # the reference's all-pairs
        print(json.dumps(out), flush=True)
    R.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
