"""CPU test (-m "not gpu") of the multi-GPU decomposition logic with torch.distributed/gloo, world_size 2:
each rank owns an x-slab, selects its halo with the rule of shard_select_kernel (opencl_dpm_b200/shard.py),
exchanges the boundary cells, and evaluates the reference's all-pairs forces (oracle) on own + ghost cells.
The union over ranks must equal the forces of the undivided tissue."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

PK = ("Kv", "Ka", "Ks", "v0", "a0", "l0")


def _worker(rank, world, port, nx, ny, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from opencl_dpm_b200 import shard, synth
    from oracle import oracle as O

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    i0, i1 = shard.slab_columns(nx, rank, world)
    d = synth.monolayer3d(nx, ny, subdiv=2, x_range=(i0, i1))
    nc, nv = d["nc"], d["nv"]
    lo, hi = O.aabb3d(d["verts"], nc)
    V = d["verts"].reshape(nc, nv, 4)
    emax = 0.0
    for f in d["faces"]:
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            emax = max(emax, float(np.linalg.norm(V[:, a, :3] - V[:, b, :3], axis=1).max()))
    summary = [None] * world
    dist.all_gather_object(summary, dict(rlo=float(lo[:, 0].min()), rhi=float(hi[:, 0].max()), ext=float((hi - lo).max()), pad=0.34 * emax))
    margin = shard.halo_margin(max(s["ext"] for s in summary), max(s["pad"] for s in summary))
    peers = sorted({(rank - 1) % world, (rank + 1) % world} - {rank})
    send = {p: shard.select_halo(lo, hi, (summary[p]["rlo"], summary[p]["rhi"]), margin, 1, float(d["L"])) for p in peers}
    inbox = [None] * world
    dist.all_gather_object(inbox, {p: (d["gid"][send[p]], V[send[p]].copy()) for p in peers})
    ghosts_gid, ghosts_V = [], []
    for p in peers:
        g, v = inbox[p][rank]
        ghosts_gid.append(g)
        ghosts_V.append(v)
    allV = np.concatenate([V] + ghosts_V).reshape(-1, 4)
    ntot = allV.shape[0] // nv
    P = [np.full(ntot, d[k][0], np.float32) for k in PK]
    F = O.forces3d_range(allV, d["faces"], *P, d["Kre"], d["PBC"], d["L"], 0, nc)
    np.savez(os.path.join(out, f"r{rank}.npz"), gid=d["gid"], F=F.reshape(ntot, nv, 4)[:nc], nghost=ntot - nc, nsent=sum(len(s) for s in send.values()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nx,ny", [(6, 3)])
def test_slab_halo_exchange_reproduces_global_forces_gloo(tmp_path, nx, ny):
    import torch.multiprocessing as mp

    from opencl_dpm_b200 import synth
    from oracle import oracle as O

    world = 2
    port = 29600 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, nx, ny, str(tmp_path)), nprocs=world, join=True)
    d = synth.monolayer3d(nx, ny, subdiv=2)
    Fg = O.forces3d(d["verts"], d["faces"], *[d[k] for k in PK], d["Kre"], d["PBC"], d["L"]).reshape(d["nc"], d["nv"], 4)
    assert np.abs(Fg).max() > 1.0
    seen = np.zeros(d["nc"], bool)
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        assert 0 < g["nghost"] < d["nc"]  # a real halo, not the whole tissue... on a 6-column periodic ring of 2 slabs
        assert np.abs(g["F"] - Fg[g["gid"]]).max() <= 1e-5 * np.abs(Fg).max()
        seen[g["gid"]] = True
    assert seen.all()


def test_slab_columns_cover_lattice():
    from opencl_dpm_b200 import shard

    for nx in (8, 13, 512):
        for world in (1, 2, 3, 8):
            cols = [shard.slab_columns(nx, r, world) for r in range(world)]
            assert cols[0][0] == 0 and cols[-1][1] == nx
            assert all(cols[i][1] == cols[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cols]
            assert max(sizes) - min(sizes) <= 1
