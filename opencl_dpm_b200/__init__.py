"""opencl_dpm_b200 — B200-native (sm_100a CUDA) replacement for the per-timestep hot path of
sudo-shaka/OpenCL_DPM (Tissue2D/Tissue3D::CLEulerUpdate).

Layers
  include/dpm_b200.h              the C ABI (the drop-in boundary)
  opencl_dpm_b200/csrc            hand-written CUDA kernels + the C ABI implementation (libdpm_b200.so)
  opencl_dpm_b200/host            C++ Cell2D/Cell3D/Tissue2D/Tissue3D with the reference's API + pybind `clDPM`
  opencl_dpm_b200.capi            ctypes binding of the C ABI (flat arrays; used by bench.py and the tests)
  opencl_dpm_b200.synth           synthetic tissues for BASELINE.json's configs

There is no CPU fallback: importing the compute entry points without the built CUDA library raises.
"""
from . import capi  # noqa: F401
from .capi import Dpm2D, Dpm3D, DpmError  # noqa: F401


def load_cldpm():
    """Import the in-tree pybind11 module `clDPM` (the reference's Python surface)."""
    import importlib.util
    import os

    from .build import cldpm_path

    path = cldpm_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -m opencl_dpm_b200.build`")
    capi.lib()  # make sure libdpm_b200.so is resolvable first
    spec = importlib.util.spec_from_file_location("clDPM", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
