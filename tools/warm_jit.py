#!/usr/bin/env python
"""Fill the run-time specialisation cache (ldpc_error_floor_b200/jit_cache) here, where NVRTC cross-compiles without a GPU,
for the graphs the GPU tests specialise: Polar(64, 48) and tests/test_gpu_parity.py's random QC graph."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ldpc_error_floor_b200 import _lib
d = np.load(os.path.join(ROOT, "ldpc_error_floor_b200", "data", "base_graphs.npz"))
t0 = time.time()
print("polar", _lib.jit_prebuild(d["graph/polar/proto"], int(d["graph/polar/meta"][0])), f"{time.time() - t0:.0f} s")
import test_gpu_parity
proto, z = test_gpu_parity._random_qc_graph()
t0 = time.time()
print("random QC graph", _lib.jit_prebuild(proto, z), f"{time.time() - t0:.0f} s")
