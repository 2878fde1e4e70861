#!/usr/bin/env python
"""Time the WiMAX decode kernel for several launch geometries (spec variants / generic), GPU only.
usage: python tools/geom_sweep.py [graph-key] ["Fp,R Fp,R ..."]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_error_floor_b200 as L

key = sys.argv[1] if len(sys.argv) > 1 else "wimax"
geos = [tuple(int(v) for v in g.split(",")) for g in (sys.argv[2].split() if len(sys.argv) > 2 else ["0,0"])]
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codes.npz")))
proto = d[f"graph/{key}/proto"].astype(np.int32); meta = d[f"graph/{key}/meta"]
z = int(meta[0])
wkey = {"wimax": "wimax_base20"}.get(key)
wkeys = [k.split("/")[1] for k in d if k.startswith("weights/") and k.endswith("/sharing")]
if wkey is None:
    cand = [k for k in wkeys if k.startswith(key)]
    wkey = cand[0] if cand else None
g = L.BaseGraph(proto, z, (int(meta[1]), int(meta[2])), (int(meta[3]), int(meta[4])))
if wkey:
    ws = L.WeightSet([int(v) for v in d[f"weights/{wkey}/sharing"]], {i: d[f"weights/{wkey}/block{i}"] for i in range(3)})
    T = int(os.environ.get("SWEEP_ITERS", "20"))
else:
    ws = L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)}); T = 20
B = int(os.environ.get("SWEEP_FRAMES", 1 << 18))
sigma = float(g.sigma([float(os.environ.get("SWEEP_SNR", "3.5"))])[0])
llr = None
for Fp, R in geos:
    for k, v in (("LDPC_B200_FP", Fp), ("LDPC_B200_R", R)):
        if v: os.environ[k] = str(v)
        else: os.environ.pop(k, None)
    dec = L.NMSDecoder(g, ws, iters=T, decoding_type=2, q_bit=5, device=0)
    if llr is None:
        llr = dec.generate(sigma, B, 1).reshape(B, -1)
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    for _ in range(2): dec.post_decode(llr, counters=cnt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cnt.zero_(); e0.record()
    for _ in range(5): dec.post_decode(llr, counters=cnt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    e0.record()
    for _ in range(5): dec.post_decode(llr, counters=cnt, early_term=True)
    e1.record(); torch.cuda.synchronize()
    ms_et = e0.elapsed_time(e1) / 5
    print(f"{key} Fp={Fp} R={R}: {dec.kernel_name} FB={dec.frames_per_cta} cps={dec.ctas_per_sm} thr={dec.threads_per_cta} "
          f"smem={dec.smem_bytes}  {B/ms/1e3:.2f} Mframes/s ({ms:.3f} ms)  ET: {B/ms_et/1e3:.2f} Mframes/s  "
          f"cnt={cnt.cpu().numpy()[:4].tolist()}", flush=True)
