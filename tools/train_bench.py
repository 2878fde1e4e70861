#!/usr/bin/env python
"""Time one training batch (ldpc_train_grad: forward + backward of the reference's step, main_Base.py:160-162) at the
reference's batch sizes.  usage: LDPC_B200_TRAIN_THREADS=<n> python tools/train_bench.py"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case
import ldpc_error_floor_b200 as L

for name, B in (("wimax_qms_333_t20", 20), ("wimax_qms_333_t20", 200), ("5g_r073_z72_qms_300_t20_sys", 40), ("5g_r050_z64_qms_222_t50", 20)):
    case = load_case(name)
    g = L.BaseGraph(case["proto"], case["z"], case["punct"], case["short"])
    dec = L.NMSDecoder(g, L.WeightSet(case["sharing"], dict(case["weights"])), iters=case["T"], decoding_type=2, q_bit=case["q_bit"],
                       clip_llr=case["clip"])
    x = dec.generate(float(g.sigma([3.0])[0]), B, seed=1)
    for _ in range(3):
        dec.train_grad(x, loss_type=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        loss, grads, _ = dec.train_grad(x, loss_type=2)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"threads {os.environ.get('LDPC_B200_TRAIN_THREADS', 'default')}: {name} T={case['T']} batch {B}: {dt * 1e3:.2f} ms per step, loss {loss:.6f}")
