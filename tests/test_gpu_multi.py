"""Multi-GPU (slab decomposition + NCCL halo exchange) parity: a sharded run must reproduce the single-GPU run of
the same tissue BIT FOR BIT (candidate lists are ordered by global cell id, so the summation order is the same).
Needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    from opencl_dpm_b200 import capi

    try:
        return capi.device_count()
    except Exception:
        return 0


# calls > 1 / reupload: several step() calls on the resident state, or with the downloaded state uploaded again in between (the
# step kernel's fused push into the neighbours' inboxes must not survive an upload, and must carry over between calls)
@pytest.mark.parametrize("world,nx,ny,subdiv,nsteps,Kat,calls,reupload",
                         [(2, 8, 6, 2, 40, 0.0, 1, 0), (2, 6, 4, 3, 12, 0.0, 1, 0), (4, 12, 4, 2, 25, 0.0, 1, 0), (2, 8, 6, 2, 30, 0.5, 1, 0),
                          (2, 8, 6, 2, 40, 0.0, 3, 0), (2, 8, 6, 2, 40, 0.0, 2, 1), (4, 12, 4, 2, 25, 0.0, 2, 1),
                          (2, 8, 6, 2, 40, 0.0, 2, -1)])  # reupload == -1: the NCCL halo path (DPM_HALO_NCCL=1), two calls
def test_sharded_run_equals_single_gpu_bit_for_bit(world, nx, ny, subdiv, nsteps, Kat, calls, reupload):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    from opencl_dpm_b200 import Dpm3D, synth

    with tempfile.TemporaryDirectory() as out:
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), out, str(nx), str(ny), str(subdiv), str(nsteps), str(Kat), str(calls), str(reupload)]
        env = dict(os.environ)
        if reupload < 0:
            env["DPM_HALO_NCCL"] = "1"
            reupload = 0
            cmd[-1] = "0"
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        d = synth.monolayer3d(nx, ny, subdiv=subdiv)
        PK = ("Kv", "Ka", "Ks", "v0", "a0", "l0")
        h = Dpm3D(d["nc"], d["nv"], d["faces"])
        if Kat != 0.0:  # vertex-vertex attraction (gather form: ghosts stay read-only, no reverse exchange)
            h.set_force_mask(15 | 16)
        h.upload(d["verts"], *[d[k] for k in PK])
        from multi_gpu_worker import split_steps

        for i, n in enumerate(split_steps(nsteps, calls)):
            h.step(n, float(d["dt"]), float(d["Kre"]), Kat, d["PBC"], float(d["L"]))
            V1, F1 = h.download()
            if reupload and i + 1 < calls:
                h.upload(V1, *[d[k] for k in PK])
        V1 = V1.reshape(d["nc"], d["nv"], 4)
        F1 = F1.reshape(d["nc"], d["nv"], 4)
        assert np.abs(F1).max() > 1.0  # contacts are active
        halo = 0
        for rk in range(world):
            g = np.load(os.path.join(out, f"rank{rk}.npz"))
            Vs = g["verts"].reshape(-1, d["nv"], 4)
            Fs = g["forces"].reshape(-1, d["nv"], 4)
            assert np.array_equal(Vs, V1[g["gid"]]), f"rank {rk}: positions differ from the single-GPU run"
            assert np.array_equal(Fs, F1[g["gid"]]), f"rank {rk}: forces differ from the single-GPU run"
            halo += int(g["halo_bytes"])
        assert halo > 0
        h.close()
