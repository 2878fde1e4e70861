#!/usr/bin/env python
"""Throughput of the float32 kernel family (decoding_type 1, and the quantised modes the packed kernels do not take:
q_bit 6, per-edge weights).  usage: python tools/f32_sweep.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_error_floor_b200 as L
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codes.npz")))
B = 1 << 17
for key, wkey, dt, qb in (("wimax", "wimax_base20", 1, 5), ("wimax", "wimax_base20", 2, 6), ("5g_r050_z64", "5g_r050_z64_boost50", 1, 5),
                          ("mackay", None, 1, 5)):
    proto = d[f"graph/{key}/proto"].astype(np.int32); meta = d[f"graph/{key}/meta"]
    g = L.BaseGraph(proto, int(meta[0]), (int(meta[1]), int(meta[2])), (int(meta[3]), int(meta[4])))
    if wkey:
        ws = L.WeightSet([int(v) for v in d[f"weights/{wkey}/sharing"]], {i: d[f"weights/{wkey}/block{i}"] for i in range(3)})
    else:
        ws = L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)})
    dec = L.NMSDecoder(g, ws, iters=20, decoding_type=dt, q_bit=qb, device=0)
    llr = dec.generate(float(g.sigma([2.0])[0]), B, 1).reshape(B, -1)
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    for _ in range(2): dec.post_decode(llr, counters=cnt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): dec.post_decode(llr, counters=cnt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{key} decoding_type={dt} q_bit={qb}: {dec.kernel_name} FB={dec.frames_per_cta} cps={dec.ctas_per_sm} thr={dec.threads_per_cta} "
          f"smem={dec.smem_bytes}  {B/ms/1e3:.2f} Mframes/s  {B/ms/1e3*g.E*g.z*20/1e6:.3f} T edge-updates/s", flush=True)
