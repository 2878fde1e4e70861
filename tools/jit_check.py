#!/usr/bin/env python
"""Throughput of run-time specialised kernels against the generic table-driven ones and against a shipped graph of similar size
(20 iterations, no early stop, decode of generated frames; Monte-Carlo with early termination for lifted graphs)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ldpc_error_floor_b200 as L
import test_gpu_parity
d = dict(np.load(os.path.join(ROOT, "tests", "golden", "codes.npz")))


def rate(dec, g, B, snr):
    x = dec.generate(float(g.sigma([snr])[0]), B, 1).reshape(B, -1)
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    dec.post_decode(x, counters=cnt); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dec.post_decode(x, counters=cnt); e1.record(); torch.cuda.synchronize()
    fps = B / e0.elapsed_time(e1) * 1e3
    return fps, fps * g.E * g.z * dec.T / 1e12


def mc_rate(dec, g, snr, n=1 << 21):
    s = float(g.sigma([snr])[0])
    dec.mc_run(s, n, 3, early_term=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c, _, _ = dec.mc_run(s, n, 4, frame_offset=n, early_term=True); e1.record(); torch.cuda.synchronize()
    return n / e0.elapsed_time(e1) * 1e3, c.cpu().numpy()


def both(name, g, ws, B, snr, mc_snr=None):
    jit = L.NMSDecoder(g, ws, iters=20)
    os.environ["LDPC_B200_NO_JIT"] = "1"
    gen = L.NMSDecoder(g, ws, iters=20)
    del os.environ["LDPC_B200_NO_JIT"]
    fj, tj = rate(jit, g, B, snr); fg, tg = rate(gen, g, B, snr)
    line = f"{name:28s} {jit.kernel_name:44s} {fj / 1e6:8.2f} Mframes/s {tj:5.2f} T edge-upd/s | generic {gen.kernel_name:24s} {fg / 1e6:8.2f} Mframes/s {tg:5.2f} T  (x{fj / fg:.2f})"
    if mc_snr is not None:
        mj, cj = mc_rate(jit, g, mc_snr); mg, cg = mc_rate(gen, g, mc_snr)
        assert np.array_equal(cj, cg)
        line += f" | MC {mc_snr} dB: {mj / 1e6:.1f} vs {mg / 1e6:.1f} Mframes/s (avg it {cj[4] / cj[0]:.2f})"
    print(line, flush=True)


g = L.BaseGraph(d["graph/polar/proto"].astype(np.int32), 1)
both("Polar(64,48) z=1", g, L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)}), 1 << 21, 4.0)
proto, z = test_gpu_parity._random_qc_graph()
g = L.BaseGraph(proto, z)
both(f"random QC 5x15 z={z}", g, L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)}), 1 << 19, 3.0, mc_snr=5.0)
for key in ("bch", "5g_r073_z32"):
    m = d[f"graph/{key}/meta"]
    g = L.BaseGraph(d[f"graph/{key}/proto"].astype(np.int32), int(m[0]), (int(m[1]), int(m[2])), (int(m[3]), int(m[4])))
    dec = L.NMSDecoder(g, L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)}), iters=20)
    f, t = rate(dec, g, 1 << 20, 4.0)
    print(f"shipped {key:20s} {dec.kernel_name:44s} {f / 1e6:8.2f} Mframes/s {t:5.2f} T edge-upd/s", flush=True)
