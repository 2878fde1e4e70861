"""Block-wise training of the NMS weights: the training side of main_Base.py / main_Post.py ("next" row N1).

What runs where: the batch forward + loss + backward is ONE CUDA kernel (csrc/nms_train.cu, `ldpc_train_grad`);
validation is the fused Monte-Carlo / decode path; this module is the host loop around them -- the schedule of
main_Base.py:108-202 (blocks of `iter_step` iterations, `epoch_input` epochs, evaluate-before-train at epoch 0,
weight dump every epoch, best-on-validation copy), Adam as TF 1.x defines it, the [Min_weight, Max_weight] clip
constraint of weight_init (Main_Functions.py:434), the eta / learning-rate discounts (:192-196).
Temporal sharing (sharing code 4) ties the rows of the iterations >= fixed_iter; init_weight = -1 draws truncated normals.
Samples: the reference draws numpy MT19937 normals; here the Philox generator of the library, cycling through the
SNR list frame by frame like create_mix_epoch (Print_Functions.py:36).
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import drivers, formats
from .graph import BaseGraph


class AdamTF1:
    """tf.train.AdamOptimizer (beta1 0.9, beta2 0.999, epsilon 1e-8):
    lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m, v moving averages;  var -= lr_t * m / (sqrt(v) + eps)."""

    def __init__(self, shapes: Dict, beta1=0.9, beta2=0.999, eps=1e-8):
        self.b1, self.b2, self.eps, self.t = beta1, beta2, eps, 0
        self.m = {k: np.zeros(s, dtype=np.float64) for k, s in shapes.items()}
        self.v = {k: np.zeros(s, dtype=np.float64) for k, s in shapes.items()}

    def step(self, params: Dict, grads: Dict, lr: float) -> None:
        self.t += 1
        lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, g in grads.items():
            g = np.asarray(g, dtype=np.float64)
            self.m[k] = self.b1 * self.m[k] + (1.0 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1.0 - self.b2) * g * g
            params[k] -= lr_t * self.m[k] / (np.sqrt(self.v[k]) + self.eps)


@dataclass
class BlockResult:
    training_iter_start: int
    training_iter_end: int
    losses: List[float] = field(default_factory=list)          # average training loss per epoch (epoch 0 = 0)
    valid: List[np.ndarray] = field(default_factory=list)      # Results[4, nSNR] per epoch
    opt_value: float = 100000.0
    weights: Optional[formats.WeightSet] = None


def make_batch(dec, SNR_sigma, batch_size: int, seed: int, frame_offset: int):
    """create_mix_epoch's batch (Print_Functions.py:35-66): frame k is drawn at SNR_sigma[k % nSNR]."""
    import torch
    n = len(SNR_sigma)
    out = torch.empty((batch_size, dec.graph.N, dec.graph.z), dtype=torch.float32, device=dec.device)
    for s, sg in enumerate(SNR_sigma):
        k = len(range(s, batch_size, n))
        if k:
            out[s::n] = dec.generate(float(sg), k, seed, frame_offset + s * batch_size)
    return out


def train_block(cfg: drivers.RunConfig, training_iter_start: int, training_iter_end: int, device: Optional[int] = None,
                init: Optional[formats.WeightSet] = None, epochs: Optional[int] = None, log=print,
                valid_frames: Optional[int] = None) -> BlockResult:
    """One pass of the `while training_iter_end <= iters_max` body of main_Base.py:110-202."""
    import torch
    from .decoder import NMSDecoder, check_params
    from .montecarlo import compute_results
    c = cfg
    if c.opt_result_print == 3:
        raise NotImplementedError("opt_result_print = 3 (best epoch by validation loss, Print_Functions.py:167-181): the "
                                  "validation pass here has no loss row -- use 0, 1 or 2")
    SNR_Matrix = check_params(c.sampling_type, c.SNR_Matrix, c.sharing, c.iters_max, c.fixed_iter, c.iter_step)
    proto = formats.read_base_graph(c.path(f"./BaseGraph/{c.filename}.txt"))
    g = BaseGraph(proto, c.z_value, (c.punct_start, c.punct_end), (c.short_start, c.short_end), name=c.filename)
    SNR_sigma = g.sigma(SNR_Matrix)
    T = training_iter_end
    ws_file = drivers.load_block_weights(c, g, training_iter_start, T, init)   # blocks under cfg.sharing's codes, T rows each
    tied = [i for i, code in enumerate(ws_file.sharing) if code in (4, 5)]
    # temporal sharing: the decoder takes the equivalent per-iteration table (code 4 -> per-edge rows; its UCN twin has no
    # branch in build_neural_network and is only carried through to the weight file)
    ws = formats.expand_temporal(ws_file, T, c.fixed_iter) if tied else ws_file
    dec = NMSDecoder(g, ws, iters=T, decoding_type=c.decoding_type, q_bit=c.q_bit, clip_llr=c.clip_LLR, device=device,
                     systematic=c.systematic)
    t_lo = max(training_iter_start - c.fixed_init, c.fixed_iter)                      # Main_Functions.py:342, 368
    data = drivers.process_data(c)
    params = {i: np.array(b, dtype=np.float64) for i, b in ws.blocks.items()}
    adam = AdamTF1({i: p.shape for i, p in params.items()})
    res = BlockResult(training_iter_start, T)
    etha, lr = c.etha_start, c.learn_rate_start
    seed = 1074 + c.seed_in
    opt_flag = False
    nbatch = math.floor(c.training_num / c.batch_size)
    os.makedirs(os.path.dirname(c.perf_filename), exist_ok=True)
    if not os.path.exists(c.perf_filename):
        with open(c.perf_filename, "w") as fh:
            fh.write(drivers.perf_header(c, SNR_Matrix, g.M, g.N, g.E, g.rate_ref))
    n_ep = c.epoch_input if epochs is None else epochs
    for epoch in range(0, n_ep + 1):
        t0 = time.time()
        avg = 0.0
        if epoch > 0:
            for b in range(nbatch):
                if c.sampling_type == 1:
                    rows = data[0][b * c.batch_size:(b + 1) * c.batch_size]
                    xa = torch.from_numpy(formats.uncor_to_llr(rows, g.N, g.z)).to(dec.device)   # read_uncor_llr (:6-10)
                else:
                    # every batch its own slice of the Philox frame space: make_batch spans len(SNR_sigma) * batch_size indices
                    xa = make_batch(dec, SNR_sigma, c.batch_size, seed,
                                    ((epoch - 1) * nbatch + b) * c.batch_size * max(len(SNR_sigma), 1))
                loss, grads, _ = dec.train_grad(xa, iter_lo=t_lo, loss_type=c.loss_type, etha=etha)
                for i in tied:
                    # temporal sharing (codes 4 / 5): ONE variable serves every iteration >= fixed_iter (weight_init
                    # :411-414, build_neural_network :299-304), so its gradient is the sum over those iterations; the
                    # rows start equal and see equal Adam updates, so they stay tied (print_weight writes them expanded)
                    if i in grads and c.fixed_iter < T:
                        grads[i][c.fixed_iter:] = grads[i][c.fixed_iter:].sum(axis=0, keepdims=True)
                adam.step(params, grads, lr)
                for i in params:                                                            # clip constraint (:434)
                    np.clip(params[i], c.Min_weight, c.Max_weight, out=params[i])
                    # frozen iterations; a tied (temporal) block is ONE variable from fixed_iter on, so only the rows
                    # before fixed_iter are frozen there -- otherwise the rows [fixed_iter, t_lo) would drift from the rest
                    lo = min(t_lo, c.fixed_iter) if i in tied else t_lo
                    params[i][:lo] = ws.blocks[i][:lo]
                dec.set_weights(formats.WeightSet(list(ws.sharing), {i: p.astype(np.float32) for i, p in params.items()}))
                avg += loss / nbatch
        t_train = time.time() - t0
        cur = formats.WeightSet(list(ws.sharing), {i: p.astype(np.float32) for i, p in params.items()})
        # print_weight (:74-96): header = the run's sharing codes, temporal blocks written as their T expanded rows
        out_ws = cur if not tied else formats.WeightSet(list(ws_file.sharing), {i: (params[i].astype(np.float32) if i in params
                                                                                     else ws_file.blocks[i]) for i in ws_file.blocks})
        formats.write_weights(c.path(formats.weights_filename(c.out_filename, T)), out_ws)
        txt = (f"* Training_iter_start: {training_iter_start} training_iter_end: {T} epoch: [{epoch}/{n_ep}]\n"
               f"Training loss: {drivers.FTE([avg])}\n")
        with open(c.perf_filename, "a") as fh:
            fh.write(txt)
        if log:
            log(txt, end="")
        t_valid = 0.0
        if c.valid_flag > 0:
            vn = c.valid_num if valid_frames is None else valid_frames
            # fresh validation samples every epoch, as the reference's advancing RandomState gives (main_Base.py:176-178)
            r, t_valid = compute_results(dec, vn, data[2], SNR_sigma, c.batch_size, c.sampling_type, seed=seed + 7919 * (epoch + 1))
            res.valid.append(r)
            res.opt_value, opt_flag = drivers.print_result(r, res.opt_value, c.perf_filename, c.out_filename, T,
                                                           c.opt_result_print, opt_flag, False, root=c.root, quiet=log is None)
            if opt_flag:
                res.weights = cur
        with open(c.perf_filename, "a") as fh:
            fh.write(f"Running time (Train/Valid/Test): {t_train:.2f}/{t_valid:.2f}/{0.0:.2f}\n\n")
        res.losses.append(avg)
        if c.etha_discount != 0 and c.etha_discount_step != 0 and (epoch + 1) % c.etha_discount_step == 0:
            etha *= c.etha_discount
        if c.learn_rate_discount != 0 and c.learn_rate_step != 0 and (epoch + 1) % c.learn_rate_step == 0:
            lr *= c.learn_rate_discount
    if res.weights is None:
        res.weights = cur
    return res


def train(cfg: drivers.RunConfig, **kw) -> List[BlockResult]:
    """The whole schedule: blocks [fixed_iter, fixed_iter + iter_step), ... up to iters_max (main_Base.py:108-110, 200-202);
    every block starts from the previous block's best-on-validation file (Main_Functions.py:390-391)."""
    out = []
    start, end = cfg.fixed_iter, cfg.fixed_iter + cfg.iter_step
    while end <= cfg.iters_max:
        out.append(train_block(cfg, start, end, **kw))
        start += cfg.iter_step
        end += cfg.iter_step
    return out
