"""Worker of tests/test_gpu_multi.py — launched with torch.distributed.run, one rank per GPU.
Runs a slab-sharded 3D tissue for a number of steps and writes each rank's owned cells to <out>/rank<r>.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def split_steps(nsteps, calls):
    base = nsteps // calls
    return [base + (1 if i < nsteps % calls else 0) for i in range(calls)]


def main():
    out, nx, ny, subdiv, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    Kat = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0  # != 0: vertex-vertex attraction on (DPM3D_ATTRACT)
    import torch
    import torch.distributed as dist

    from opencl_dpm_b200 import Dpm3D, shard, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = shard.broadcast_unique_id(rank)
    i0, i1 = shard.slab_columns(nx, rank, world)
    d = synth.monolayer3d(nx, ny, subdiv=subdiv, x_range=(i0, i1))
    PK = ("Kv", "Ka", "Ks", "v0", "a0", "l0")
    h = Dpm3D(d["nc"], d["nv"], d["faces"], device=local)
    h.shard_init(rank, world, uid, max_ghost=max(8, 3 * ny))
    h.set_global_ids(d["gid"])
    if Kat != 0.0:
        h.set_force_mask(15 | 16)
    # calls > 1: the timesteps are split over several step() calls; reupload: the downloaded state is uploaded again between them
    calls = int(sys.argv[7]) if len(sys.argv) > 7 else 1
    reupload = bool(int(sys.argv[8])) if len(sys.argv) > 8 else False
    h.upload(d["verts"], *[d[k] for k in PK])
    for i, n in enumerate(split_steps(nsteps, calls)):
        h.step(n, float(d["dt"]), float(d["Kre"]), Kat, d["PBC"], float(d["L"]))
        V, F = h.download()
        if reupload and i + 1 < calls:
            h.upload(V, *[d[k] for k in PK])
    st = h.stats()
    np.savez(os.path.join(out, f"rank{rank}.npz"), gid=d["gid"], verts=V, forces=F, rebuilds=st.rebuilds, halo_bytes=st.halo_bytes,
             contact_evals=st.contact_evals)
    dist.barrier()
    h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
