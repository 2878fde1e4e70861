"""Oracle for the training step ("next" row N1): the reference's OWN loss (Main_Functions.py:337-378) on the
reference's OWN forward graph, differentiated by torch.autograd.  TEST INFRASTRUCTURE ONLY; development container
only (needs /root/reference).  See oracle/tf_shim_torch/v1_torch.py for why torch's gradients stand in for TF's."""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

from . import ref_runner

_MOD = None


def _load():
    global _MOD
    if _MOD is None:
        mf, _ = ref_runner.load_reference()          # numpy-shimmed copy: init_parameter / init_connecting_matrix
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path.insert(0, os.path.join(here, "tf_shim_torch"))
        import v1_torch
        spec = importlib.util.spec_from_file_location("Main_Functions_torch",
                                                      os.path.join(ref_runner.REFERENCE_ROOT, "Main_Functions.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.tf = v1_torch                             # build_neural_network now runs on torch tensors
        _MOD = (mf, mod, v1_torch)
    return _MOD


def loss_and_grads(code_proto, z, sharing, weights, xa, T, iter_start=0, loss_type=2, etha=0.0, decoding_type=2,
                   q_bit=5, clip_llr=20.0, punct=(0, 0), short=(0, 0), fixed_iter=0, fixed_init=0, target_node=None,
                   iters_max=None):
    """One training batch exactly as main_Base.py:136-137, 160-162 builds it: build_neural_network for
    t = 0..T-1 with sampling_type 0, training block [iter_start, T).  Returns dict(loss, grads {(type, t): array},
    app [T, B, N*z])."""
    import torch
    mf, mod, tf = _load()
    code_proto = np.asarray(code_proto, dtype=int)
    M, N, base, cn_deg, vn_deg, E, rate, sigma = mf.init_parameter(code_proto, np.array([0.0]), z, punct[0], punct[1],
                                                                    short[0], short[1])
    E = int(E)
    mats = mf.init_connecting_matrix(code_proto, base, N, M, E, z, vn_deg, cn_deg, punct[0], punct[1])
    xa = np.asarray(xa, dtype=np.float32)
    B = xa.shape[0]
    target = N if target_node is None else int(target_node)
    net = {"xa": torch.from_numpy(xa), "ya": torch.zeros((B, N * z)), "etha": float(etha), "learn_rate": 0.0,
           "LLRa0": torch.zeros((B, z, E))}
    for i, code in enumerate(sharing):
        if code > 0:
            rows = T if code in (1, 2, 3) else fixed_iter + 1
            w = np.asarray(weights[i], dtype=np.float32)
            for t in range(rows):
                net[f"var_{i}_{t}"] = torch.tensor(np.ascontiguousarray(w[t]).reshape(-1), requires_grad=True)
    tf.train.AdamOptimizer.last = None
    for t in range(T):
        net = mod.build_neural_network(net, list(sharing), decoding_type, 0, loss_type, target, t,
                                       T if iters_max is None else iters_max, fixed_iter, fixed_init, iter_start, T,
                                       N, M, E, z, B, *mats, q_bit, clip_llr)
    loss, var_list = tf.train.AdamOptimizer.last
    grads = torch.autograd.grad(loss, var_list, allow_unused=True)
    names = {id(v): k for k, v in net.items() if isinstance(k, str) and k.startswith("var_")}
    out = {}
    for v, g in zip(var_list, grads):
        _, i, t = names[id(v)].split("_")
        out[(int(i), int(t))] = np.zeros(v.shape, np.float32) if g is None else g.detach().numpy().astype(np.float32)
    app = np.stack([net[f"ya_output{t}"].detach().numpy() for t in range(T)], axis=0)
    return {"loss": float(loss.detach()), "grads": out, "app": app.astype(np.float32)}
