#!/usr/bin/env python
"""Fused Monte-Carlo throughput vs Eb/N0 with early termination: batch kernels (single launch and two-stage) against the
persistent-slot kernel in every compiled geometry; asserts identical counters.
usage: python tools/mc_sweep.py <graph-key> "<snr list>" [systematic]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_error_floor_b200 as L
from ldpc_error_floor_b200 import montecarlo

key = sys.argv[1] if len(sys.argv) > 1 else "wimax"
snrs = [float(s) for s in (sys.argv[2].split() if len(sys.argv) > 2 else "3.0 4.0 5.0 6.0".split())]
systematic = int(sys.argv[3]) if len(sys.argv) > 3 else (1 if key.startswith("5g") else 0)
geoms = [tuple(int(v) for v in a.split(",")) for a in sys.argv[4].split()] if len(sys.argv) > 4 else [(0, 0)]
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codes.npz")))
proto = d[f"graph/{key}/proto"].astype(np.int32); meta = d[f"graph/{key}/meta"]
g = L.BaseGraph(proto, int(meta[0]), (int(meta[1]), int(meta[2])), (int(meta[3]), int(meta[4])))
wk = [k.split("/")[1] for k in d if k.startswith("weights/") and k.endswith("/sharing") and k.split("/")[1].startswith(key)]
ws = (L.WeightSet([int(v) for v in d[f"weights/{wk[0]}/sharing"]], {i: d[f"weights/{wk[0]}/block{i}"] for i in range(3)}).rows(0, 20)
      if wk else L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)}))
n = 1 << 22


def timed(dec, sigma, **kw):
    dec.mc_run(sigma, n, 5, early_term=True, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cnt, _, _ = dec.mc_run(sigma, n, 6, frame_offset=n, early_term=True, **kw)
    e1.record(); torch.cuda.synchronize()
    return n / e0.elapsed_time(e1) / 1e3, cnt.cpu().numpy()


decs = {}
os.environ["LDPC_B200_NO_PERSIST"] = "1"
base = L.NMSDecoder(g, ws, iters=20, decoding_type=2, q_bit=5, device=0, systematic=systematic)
del os.environ["LDPC_B200_NO_PERSIST"]
for fp, r in geoms:
    if fp:
        os.environ["LDPC_B200_MCP_FP"], os.environ["LDPC_B200_MCP_R"] = str(fp), str(r)
    dd = L.NMSDecoder(g, ws, iters=20, decoding_type=2, q_bit=5, device=0, systematic=systematic)
    info = dd.mc_info()
    if info["persistent"]:
        decs[info["kernel"] + f" {info['ctas_per_sm']}x{info['threads_per_cta']}thr {info['smem_bytes'] // 1024}KB"] = dd
    os.environ.pop("LDPC_B200_MCP_FP", None); os.environ.pop("LDPC_B200_MCP_R", None)
print(f"# {key} systematic={systematic} n={n} batch kernel {base.kernel_name}; persistent: {list(decs)}", flush=True)
for snr in snrs:
    sigma = float(g.sigma([snr])[0])
    os.environ["LDPC_B200_NO_PERSIST"] = "1"
    m0, c0 = timed(base, sigma)
    s1 = montecarlo._pick_stage1(c0, 20)
    m1, c1 = timed(base, sigma, stage1_iters=s1) if s1 else (float("nan"), c0)
    del os.environ["LDPC_B200_NO_PERSIST"]
    assert np.array_equal(c0, c1)
    line = f"{key} {snr:5.2f} dB  avg_it {c0[4] / c0[0]:5.2f} synd_fail {c0[5] / c0[0]:.3f} FER {c0[2] / c0[0]:.2e} | batch {m0:7.2f}  two-stage(s1={s1}) {m1:7.2f}"
    for name, dd in decs.items():
        m2, c2 = timed(dd, sigma)
        assert np.array_equal(c0, c2), (name, c0, c2)
        line += f" | {name.split()[0].replace('nms_mcp_spec_' + key + '_', '')} {m2:7.2f}"
    print(line + "  Mframes/s", flush=True)
