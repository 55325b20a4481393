"""TEST INFRASTRUCTURE: CPU oracle of the OpenCL_DPM hot path. Only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package."""
