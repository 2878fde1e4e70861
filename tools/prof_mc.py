#!/usr/bin/env python
"""One fused Monte-Carlo launch (generate + decode + count, early termination) for ncu captures.
usage: python tools/prof_mc.py <graph-key> <snr-dB> [frames] [systematic]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_error_floor_b200 as L
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codes.npz")))
key, snr = sys.argv[1], float(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 21
proto = d[f"graph/{key}/proto"].astype(np.int32); meta = d[f"graph/{key}/meta"]
g = L.BaseGraph(proto, int(meta[0]), (int(meta[1]), int(meta[2])), (int(meta[3]), int(meta[4])))
wk = {"wimax": "wimax_base20"}.get(key)
ws = (L.WeightSet([int(v) for v in d[f"weights/{wk}/sharing"]], {i: d[f"weights/{wk}/block{i}"] for i in range(3)}) if wk
      else L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)}))
dec = L.NMSDecoder(g, ws, iters=20, systematic=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
print(dec.mc_info())
sigma = float(g.sigma([snr])[0])
dec.mc_run(sigma, n, 3, early_term=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); c, _, _ = dec.mc_run(sigma, n, 4, frame_offset=n, early_term=True); e1.record(); torch.cuda.synchronize()
c = c.cpu().numpy()
print(f"{key} {snr} dB: {n / e0.elapsed_time(e1) / 1e3:.2f} Mframes/s avg it {c[4] / c[0]:.2f} FER {c[2] / c[0]:.2e}")
