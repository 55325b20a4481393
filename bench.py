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

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload D642|D162|C162|B2D] [--inner T]
  python bench.py --impl reference ...   # the reference algorithm on the host cores (oracle arm)
"""
from __future__ import annotations

import argparse
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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(name: str, rank: int = 0, world: int = 1):
    from opencl_dpm_b200 import shard, synth

    if name in ("E642", "E162"):
        # BASELINE config E: 512 x 512 = 262,144 cells; each rank builds only its x-slab of lattice columns
        sub = 3 if name == "E642" else 2
        i0, i1 = shard.slab_columns(512, rank, world)
        d = synth.monolayer3d(512, subdiv=sub, x_range=(i0, i1))
        d["nc_global"] = 512 * 512
        desc = (f"3D DPM 262,144-cell monolayer, {d['nv']}-vertex icospheres ({512 * 512 * d['nv']:,} vertices), x-slab decomposed over "
                f"{world} GPU(s) with per-step NCCL halo exchange")
    elif name == "D642":
        d = synth.monolayer3d(64, subdiv=3)
        desc = "3D DPM 4096-cell monolayer, 642-vertex icospheres (2,629,632 vertices), winding-number repulsion, substrate, PBC"
    elif name == "D162":
        d = synth.monolayer3d(64, subdiv=2)
        desc = "3D DPM 4096-cell monolayer, 162-vertex icospheres (663,552 vertices; the reference's mesh)"
    elif name == "C162":
        d = synth.monolayer3d(8, subdiv=2)
        desc = "3D DPM 64-cell monolayer, 162-vertex icospheres"
    elif name == "B2D":
        d = synth.tissue2d(64, nv=64)
        desc = "2D DPM 4096 cells x 64 vertices, area+perimeter+bending+attraction+repulsion, PBC"
    else:
        raise SystemExit(f"unknown workload {name}")
    d["name"], d["desc"], d["dim"] = name, desc, (2 if name == "B2D" else 3)
    d.setdefault("nc_global", d["nc"])
    return d


class stdout_to_stderr:
    """NCCL prints its version banner on fd 1 when the first communicator is created; the contract is ONE JSON line on
    stdout, so fd 1 is pointed at stderr while communicators are set up."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

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


def cpu_baseline_sample(d, budget_s: float = 20.0, allpairs: bool = True):
    """CPU baseline on the box's host cores (oracle port, fp32, OpenMP), one timestep of forces.
    allpairs=True : the reference's own ALL-PAIRS algorithm on a bounded sample — forces of the first k cells against
                    all cells (cost per vertex is the same for every cell, so the sample's rate is the workload's rate);
    allpairs=False: the same kernels behind the CPU restatement of the cell list (culled form, whole tissue) — the
                    baseline BASELINE.md plans for configs B/D, where all-pairs takes hours per step."""
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    if d["dim"] == 3 and allpairs:
        per_cell = d["nv"] * (d["nc"] - 1) * d["nf"] / (1.7e7 * cores)
        ns = int(max(1, min(d["nc"], budget_s / max(per_cell, 1e-9))))
        t0 = time.perf_counter()
        O.forces3d_range(d["verts"], d["faces"], *[d[k] for k in PK3], d["Kre"], d["PBC"], d["L"], 0, ns)
        dt = time.perf_counter() - t0
        nvert = ns * d["nv"]
        sample = f"reference all-pairs algorithm: forces of the first {ns} of {d['nc']} cells against all cells, 1 timestep ({dt:.1f} s)"
    elif d["dim"] == 3:
        V = d["verts"].reshape(d["nc"], d["nv"], 4)
        f = d["faces"]
        emax = max(float(np.linalg.norm(V[:, f[:, i], :3] - V[:, f[:, (i + 1) % 3], :3], axis=2).max()) for i in range(3))
        lo, hi = O.aabb3d(d["verts"], d["nc"])
        t0 = time.perf_counter()
        cl = O.cell_list(3, lo, hi, d["PBC"], d["L"], 0.1, 1.25 * 0.34 * emax, 32)
        O.forces3d(d["verts"], d["faces"], *[d[k] for k in PK3], d["Kre"], d["PBC"], d["L"], cand_count=cl["cand_count"], cand=cl["cand"])
        dt = time.perf_counter() - t0
        nvert = d["nc"] * d["nv"]
        sample = f"culled form (CPU cell list + literal kernels), all {d['nc']} cells, 1 timestep incl. list build ({dt:.1f} s)"
    else:
        t0 = time.perf_counter()
        if allpairs:
            O.forces2d(d["verts"], d["nv"], *[d[k] for k in PK2], d["Kre"], d["Kat"], d["PBC"], d["L"])
            kind = "reference all-pairs algorithm"
        else:
            lo, hi = O.aabb2d(d["verts"], d["nv"])
            cl = O.cell_list(2, lo, hi, d["PBC"], d["L"], 0.1, float(d["l0"].max()) if d["Kat"] != 0 else 0.0, 64, far2d=True)
            O.forces2d(d["verts"], d["nv"], *[d[k] for k in PK2], d["Kre"], d["Kat"], d["PBC"], d["L"], cand_count=cl["cand_count"], cand=cl["cand"])
            kind = "culled form (CPU cell list + literal kernels)"
        dt = time.perf_counter() - t0
        nvert = int(d["nv"].sum())
        sample = f"{kind}, all {d['nc']} cells, 1 timestep ({dt:.1f} s)"
    return {"value": nvert / dt, "unit": "vertex-steps/s", "cores": cores, "kind": "port", "sample": sample}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d = make_workload(args.workload)
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline_sample(d, budget_s=max(4.0, 40.0 / max(1, args.steps + args.warmup)))
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    extra = {}
    try:  # the REAL reference (its own host code + OpenCL kernels, oracle/_ref) when an OpenCL device is reachable
        from oracle import ref as R

        if R.available():
            from opencl_dpm_b200 import synth

            d64 = synth.monolayer3d(8, subdiv=2)
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
           "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": {"workload": d["desc"], "name": d["name"], "algorithm": "reference all-pairs (CPU port of the OpenCL kernels)"},
           "cpu_baseline": cb, "e2e": {"value": v, "unit": "vertex-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, **extra}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="D642 (default at 1 GPU), D162, C162, B2D, E642 (default at >1 GPU), E162")
    ap.add_argument("--inner", type=int, default=None, help="timesteps per bench step (one CLEulerUpdate call)")
    ap.add_argument("--equilibrate", type=int, default=None,
                    help="untimed timesteps from the synthetic lattice that produce the batch every bench step processes (default 100). "
                         "The lattice starts with ~5 %% overlap everywhere: its first ~30 timesteps are dominated by the contact kernel "
                         "(0.4-0.7 ms per timestep), from ~40 to ~150 the contacts are active at a steady moderate level, and by ~300 "
                         "the cells have pushed each other apart and no vertex is in contact any more")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        # 1 GPU: BASELINE config D (4096 cells, the north-star roofline target).  N > 1: config E (262,144 cells) is
        # strong-scaled over the ranks; vertex-steps/s of this path is independent of the cell count (same lattice,
        # same per-cell work), so the N=1 D642 value is the single-GPU baseline of the same metric.
        args.workload = "D642" if world_env == 1 else "E642"
    if args.inner is None:
        # timesteps per CLEulerUpdate-equivalent call: the reference's own 3D demo advances 25 per call (test3D.py:20-22)
        args.inner = 25 if not args.workload.startswith("E") else 10

    if args.equilibrate is None:
        args.equilibrate = 100

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch

    from opencl_dpm_b200 import Dpm2D, Dpm3D, capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    redirect = stdout_to_stderr()
    redirect.__enter__()  # until the warm-up is done (NCCL communicators are created lazily)
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
    nvert = d["nc"] * d["nv"] if dim == 3 else int(d["nv"].sum())
    balg = 32 if dim == 3 else 16

    if dim == 3:
        h = Dpm3D(d["nc"], d["nv"], d["faces"], device=local)
        if sharded:
            from opencl_dpm_b200 import shard

            h.shard_init(rank, world, shard.broadcast_unique_id(rank), max_ghost=768 if world > 2 else 1280)
            # one lattice column (512 cells) per slab face + headroom; with 2 ranks the single peer is both neighbours
            h.set_global_ids(d["gid"])
        h.set_stream(stream.cuda_stream)
        params = [d[k] for k in PK3]
        dev_verts = torch.from_numpy(d["verts"]).cuda()

        def reset(src=None):
            h.upload_device((dev_verts if src is None else src).data_ptr(), *params)

        def run_steps(n):
            h.step(n, float(d["dt"]), float(d["Kre"]), 0.0, d["PBC"], float(d["L"]))

        def snapshot():  # the current state as a device tensor (what reset(src=...) re-uploads)
            return torch.from_numpy(h.download(want_forces=False)[0]).cuda()
        host_v = torch.from_numpy(d["verts"].copy()).pin_memory()
        host_f = torch.zeros_like(host_v).pin_memory()
        h2d = d["verts"].nbytes + 6 * 4 * d["nc"]
        d2h = 2 * d["verts"].nbytes

        def e2e_call(n):
            t0 = time.perf_counter()
            h.euler_update(host_v.numpy(), *params, n, float(d["dt"]), float(d["Kre"]), 0.0, d["PBC"], float(d["L"]), forces_out=host_f.numpy())
            return time.perf_counter() - t0
    else:
        h = Dpm2D(d["nc"], d["S"], device=local)
        h.set_neighbor_params(0.1, 64)
        h.set_stream(stream.cuda_stream)
        params = [d[k] for k in PK2]

        def reset(src=None):
            h.upload(d["verts"] if src is None else src, d["nv"], *params)

        def run_steps(n):
            h.step(n, float(d["dt"]), float(d["Kre"]), float(d["Kat"]), d["PBC"], float(d["L"]))

        def snapshot():  # the 2D ABI uploads from host memory
            return h.download(want_forces=False)[0].copy()
        host_v = torch.from_numpy(d["verts"].copy()).pin_memory()
        host_f = torch.zeros_like(host_v).pin_memory()
        h2d = d["verts"].nbytes + (6 * 4 + 4) * d["nc"]
        d2h = 2 * d["verts"].nbytes

        def e2e_call(n):
            t0 = time.perf_counter()
            h.euler_update(host_v.numpy(), d["nv"], *params, n, float(d["dt"]), float(d["Kre"]), float(d["Kat"]), d["PBC"], float(d["L"]),
                           forces_out=host_f.numpy())
            return time.perf_counter() - t0

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) -----------------------------------------------
    # The batch every bench step processes: the synthetic lattice relaxed for `equilibrate` timesteps (untimed, all ranks):
    # past the contact-dominated first steps, with contacts still active (stats.contact_evals_per_timestep says how many
    # (vertex, neighbour) evaluations a timestep of the timed region did).  Each step re-uploads THAT state (device to device, before the timed region: inputs resident in HBM when it
    # starts) and advances it `inner` timesteps, so a step's work does not depend on how many steps came before it — left to
    # run on for thousands of timesteps the D-parameter cells crumple under the substrate force (the model, not the
    # integration: the fp64 CPU oracle does the same), stop being star-shaped, and their contacts take the literal
    # all-faces sum, ten times slower; that regime is reported separately in DESIGN.md, not mixed into the headline.
    reset()
    if args.equilibrate > 0:
        run_steps(args.equilibrate)
        torch.cuda.synchronize()
    batch = snapshot()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up: nvidia-smi needs ~0.3 s before its first sample; all samples are under load
    for _ in range(args.warmup):
        reset(batch)
        run_steps(args.inner)
    torch.cuda.synchronize()
    redirect.__exit__()
    launches0 = h.stats().launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for a, b in evs:
        reset(batch)    # this step's input batch becomes resident (outside the event pair); the first timestep rebuilds the lists
        flush.fill_(1)  # flush L2 between timed steps (outside the event pair)
        a.record(stream)
        run_steps(args.inner)
        b.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    h.sync()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = h.stats().launches - launches0 - args.steps  # minus the bounds kernel of each step's (untimed) batch upload
    st = h.stats()
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if world > 1:
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
    batch_host = (batch.cpu() if torch.is_tensor(batch) else torch.from_numpy(batch)).view_as(host_v)

    def e2e_step():
        host_v.copy_(batch_host)  # untimed: the caller's buffer holds the batch again (the call updates it in place)
        return e2e_call(args.inner)
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

    if rank == 0:
        peak, peak_src = peaks()
        n_step_kernels = args.inner * args.steps
        t_launch = ms * 1e-3 / n_step_kernels  # step-kernel launches dominate the region (the rebuild kernel is a no-op launch)
        achieved = nvert * balg / t_launch / 1e9  # per GPU: this rank's owned vertices per launch
        traffic = None
        try:  # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload, if any
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(d["name"], {}).get("bytes_per_launch")
        except Exception:
            pass
        out = {
            "metric": "vertex-steps/sec (force+integrate)", "value": value, "unit": "vertex-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": d["desc"], "name": d["name"], "timesteps_per_step": args.inner, "dt": float(d["dt"]),
                       "parallelism": ("single GPU" if world == 1 else
                                       f"x-slab decomposition over {world} GPUs, ghost cells exchanged every timestep with ncclSend/ncclRecv "
                                       f"(ring of slabs), one 8-float ncclAllGather per step for the global rebuild decision" if sharded
                                       else f"{world} independent replicas (one tissue per GPU)"),
                       "cells_per_gpu": int(d["nc"]), "cells_total": int(d["nc_global"]) if sharded or world == 1 else int(d["nc"]) * world,
                       "l2": "L2 flushed (256 MiB write) between timed steps; within a step consecutive timesteps reuse L2 as in the real loop",
                       "ms_per_timestep": ms / n_step_kernels,
                       "state": f"jittered lattice relaxed for {args.equilibrate} untimed timesteps (contacts active: see stats.contact_evals_per_timestep); every bench step "
                                "re-uploads that state device-to-device before its timed region and advances it timesteps_per_step timesteps"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "vertex-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_t * 1e3 / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "dpm3d_step_kernel" if dim == 3 else "dpm2d_step_kernel", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_vertex_step": balg, "us_per_launch": t_launch * 1e6},
            "stats": {"rebuilds": int(st.rebuilds), "contact_evals_per_timestep": st.contact_evals / max(1, st.steps),
                      "wall_s_timed_region": t_wall, "halo_bytes_per_timestep_rank0": st.halo_bytes / max(1, st.steps),
                      "literal_fallback_evals_per_timestep": st.reserved[0] / max(1, st.steps)},
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline_sample(d, allpairs=False)                  # same algorithmic complexity
            out["cpu_baseline_reference_algorithm"] = cpu_baseline_sample(d, budget_s=12.0)  # the reference's all-pairs
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
