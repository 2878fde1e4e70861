#!/usr/bin/env python
"""Fused Monte-Carlo throughput vs Eb/N0 (early termination on): shows the floor set by sample generation +
prologue/epilogue when frames converge in 1-2 iterations.  usage: python tools/mc_sweep.py [graph-key]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_error_floor_b200 as L
key = sys.argv[1] if len(sys.argv) > 1 else "wimax"
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codes.npz")))
proto = d[f"graph/{key}/proto"].astype(np.int32); meta = d[f"graph/{key}/meta"]
g = L.BaseGraph(proto, int(meta[0]), (int(meta[1]), int(meta[2])), (int(meta[3]), int(meta[4])))
wk = [k.split("/")[1] for k in d if k.startswith("weights/") and k.endswith("/sharing") and k.split("/")[1].startswith(key)]
ws = (L.WeightSet([int(v) for v in d[f"weights/{wk[0]}/sharing"]], {i: d[f"weights/{wk[0]}/block{i}"] for i in range(3)})
      if wk else L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)}))
dec = L.NMSDecoder(g, ws, iters=20, decoding_type=2, q_bit=5, device=0)
n = 1 << 22
for snr in [float(s) for s in (sys.argv[2].split() if len(sys.argv) > 2 else "3.0 4.0 5.0 6.0 8.0 12.0".split())]:
    sigma = float(g.sigma([snr])[0])
    cnt, _, _ = dec.mc_run(sigma, n, 5, early_term=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cnt.zero_(); e0.record()
    dec.mc_run(sigma, n, 6, frame_offset=n, early_term=True, counters=cnt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = cnt.cpu().numpy()
    print(f"{key} {snr:5.1f} dB: {n/ms/1e3:8.2f} Mframes/s  avg iters {c[4]/c[0]:.2f}  FER {c[2]/c[0]:.2e}  ({dec.kernel_name})", flush=True)
x = dec.generate(float(g.sigma([4.0])[0]), n, 5)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); x = dec.generate(float(g.sigma([4.0])[0]), n, 6); e1.record(); torch.cuda.synchronize()
print(f"stand-alone generator: {n/e0.elapsed_time(e1)/1e3:.2f} Mframes/s ({n*g.NZ/e0.elapsed_time(e1)/1e6:.2f} G samples/s)")
