"""Host-side helpers for the slab-decomposed (multi-GPU) 3D path: x-slab partition of a tissue, the halo
selection rule restated in numpy (used by the CPU/gloo tests), and the NCCL bootstrap over torch.distributed."""
from __future__ import annotations

import numpy as np


def slab_columns(nx: int, rank: int, world: int) -> tuple[int, int]:
    """Lattice columns [i0, i1) owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(nx, world)
    i0 = rank * base + min(rank, rem)
    return i0, i0 + base + (1 if rank < rem else 0)


def partition_by_x(verts4: np.ndarray, nc: int, world: int, L: float) -> np.ndarray:
    """Owner rank of every cell of a flat tissue: slab of its centroid's wrapped x coordinate."""
    V = verts4.reshape(nc, -1, 4)
    cx = V[:, :, 0].mean(1)
    cx = cx - L * np.floor(cx / L)
    return np.minimum((cx / L * world).astype(np.int64), world - 1)


def select_halo(lo: np.ndarray, hi: np.ndarray, region: tuple[float, float], margin: float, pbc: int, L: float) -> np.ndarray:
    """Indices (ascending) of the cells whose x-extent, grown by `margin`, reaches `region` = (qlo, qhi) under the
    periodic image that brings them closest — the rule of shard_select_kernel (csrc/dpm_halo.cu)."""
    d = 0.5 * (lo[:, 0] + hi[:, 0]) - 0.5 * (region[0] + region[1])
    if pbc:
        d = d - L * np.round(d / L)
    ok = np.abs(d) <= 0.5 * (hi[:, 0] - lo[:, 0]) + margin + 0.5 * (region[1] - region[0])
    return np.nonzero(ok)[0]


def halo_margin(max_ext: float, max_pad: float, skin_rel: float = 0.1) -> float:
    return skin_rel * max_ext + 1.25 * max_pad + 1e-4 * max_ext


def broadcast_unique_id(rank: int) -> bytes:
    """rank 0 creates the NCCL unique id, everybody receives it through torch.distributed."""
    import torch.distributed as dist

    from . import capi

    obj = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]
