"""Reader of the flat binary trajectories written by Tissue2D/Tissue3D.AppendFrame (host/trajectory.cpp): the
replacement for the reference's per-frame matplotlib PNG loop (plot.py) — SURVEY §8f rank 4."""
from __future__ import annotations

import numpy as np

_HDR = np.dtype([("magic", "S8"), ("dim", "<i4"), ("ncells", "<i4"), ("nv", "<i4"), ("nf", "<i4"), ("L", "<f4"), ("pbc", "<i4"),
                 ("reserved", "<i4", (2,))])


def read(path: str, mmap: bool = True) -> dict:
    """Returns dict(dim, ncells, nv, L, PBC, faces (3D) | NV (2D), frames[nframes, ncells, nv, dim]); frames is a
    read-only memory map unless mmap=False."""
    h = np.fromfile(path, dtype=_HDR, count=1)
    if len(h) != 1 or h["magic"][0] != b"DPMTRAJ1":
        raise ValueError(f"{path}: not a DPM trajectory")
    h = h[0]
    dim, nc, nv, nf = int(h["dim"]), int(h["ncells"]), int(h["nv"]), int(h["nf"])
    off = _HDR.itemsize
    out = dict(dim=dim, ncells=nc, nv=nv, L=float(h["L"]), PBC=int(h["pbc"]))
    if dim == 3:
        out["faces"] = np.fromfile(path, dtype="<i4", count=nf * 3, offset=off).reshape(nf, 3)
        off += 4 * nf * 3
    else:
        out["NV"] = np.fromfile(path, dtype="<i4", count=nc, offset=off)
        off += 4 * nc
    per = nc * nv * dim
    import os

    nframes = (os.path.getsize(path) - off) // (4 * per)
    if mmap and nframes > 0:
        fr = np.memmap(path, dtype="<f4", mode="r", offset=off, shape=(nframes, nc, nv, dim))
    else:
        fr = np.fromfile(path, dtype="<f4", count=nframes * per, offset=off).reshape(nframes, nc, nv, dim)
    out["frames"] = fr
    return out
