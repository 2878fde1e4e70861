#!/usr/bin/env python
"""One decode launch of a chosen graph / mode, for `ncu -k regex:nms_ -c 1` captures.
usage: python tools/prof_one.py <graph-key> <decoding_type> <q_bit> [frames] [early_term]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_error_floor_b200 as L
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codes.npz")))
key, dt, qb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 16
WK = {"wimax": "wimax_base20", "wifi": "wifi_boost50", "5g_r050_z64": "5g_r050_z64_boost50"}
proto = d[f"graph/{key}/proto"].astype(np.int32); meta = d[f"graph/{key}/meta"]
g = L.BaseGraph(proto, int(meta[0]), (int(meta[1]), int(meta[2])), (int(meta[3]), int(meta[4])))
wkey = WK.get(key)
if wkey and f"weights/{wkey}/sharing" in d:
    ws = L.WeightSet([int(v) for v in d[f"weights/{wkey}/sharing"]], {i: d[f"weights/{wkey}/block{i}"] for i in range(3)})
else:
    ws = L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)})
dec = L.NMSDecoder(g, ws, iters=20, decoding_type=dt, q_bit=qb, device=0)
llr = dec.generate(float(g.sigma([2.0])[0]), B, 1).reshape(B, -1)
cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
dec.post_decode(llr, counters=cnt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dec.post_decode(llr, counters=cnt); e1.record(); torch.cuda.synchronize()
print(dec.kernel_name, f"{B / e0.elapsed_time(e1) / 1e3:.2f} Mframes/s", cnt.tolist())
