"""Callers of the hot path: the evaluation / collection side of main_Base.py and main_Post.py.

The reference drivers are scripts whose configuration is a block of module-level variables
(main_Base.py:22-63, main_Post.py:22-63) followed by: check_params -> process_data -> init_parameter ->
Performance.txt header -> per training block {weight_init, epoch loop {train, print_weight, validation
compute_results, print_result, optional test pass}}.  `RunConfig` mirrors those variables name for name and
`evaluate` runs everything of that flow that is NOT training (the epoch-0 pass of every block: the reference
evaluates before it trains, main_Base.py:143-183): validation Monte-Carlo (`sampling_type` 0), collection of
uncorrected words into ./Uncor.txt (`sampling_type` 2), and the post decoder on Inputs/[Uncor]_* (`sampling_type`
1, main_Post.py), writing `./Weights/C{core_idx}_{filename}_Performance.txt` in the reference's format
(main_Base.py:90-103, Print_Functions.py:185-228).  Training itself (loss, Adam) is out of scope.
"""
from __future__ import annotations

import math
import os
import shutil
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import formats
from .graph import BaseGraph


@dataclass
class RunConfig:
    """The module-level variables of main_Base.py:14,22-63 (same names, same defaults)."""
    core_idx: int = 0
    filename: str = "wman_N0576_R34_z24"
    sharing: List[int] = field(default_factory=lambda: [3, 0, 3])   # CN_Weight, UCN_Weight, VN_Weight
    sampling_type: int = 0      # 0: Default, 1: Read_Uncor, 2: Collect_Uncor
    decoding_type: int = 2      # 0: SP, 1: MS, 2: QMS
    q_bit: int = 5
    systematic: int = 0
    z_value: int = 24
    punct_start: int = 0
    punct_end: int = 0
    short_start: int = 0
    short_end: int = 0
    iters_max: int = 20
    fixed_iter: int = 0
    fixed_init: int = 0
    iter_step: int = 20
    loss_type: int = 2
    opt_result_print: int = 1   # 0: BER_last, 1: FER_last, 2: FER, 3: Loss
    etha_start: float = 0
    etha_discount: float = 0
    etha_discount_step: int = 0
    learn_rate_start: float = 0.001
    learn_rate_discount: float = 0
    learn_rate_step: int = 0
    batch_size: int = 20
    training_num: int = 10000
    epoch_input: int = 200
    valid_flag: int = 1
    valid_num: int = 10000
    test_flag: int = 0
    test_num: int = 400
    init_from_file: int = 0
    init_weight: float = 1
    init_VN_weight: float = 1
    Max_weight: float = 2
    Min_weight: float = 0
    seed_in: int = 2
    SNR_Matrix: np.ndarray = field(default_factory=lambda: np.array([2, 2.5, 3.0, 3.5, 4.0]))
    clip_LLR: float = 20.0      # main_Base.py:69
    root: str = "."             # directory holding BaseGraph/ Weights/ Inputs/ (the reference uses the CWD)

    @classmethod
    def post(cls, **kw) -> "RunConfig":
        """main_Post.py's values where they differ from main_Base.py (:25-26, 35-38, 53-55, 63)."""
        d = dict(sharing=[3, 3, 3], sampling_type=1, iters_max=30, fixed_iter=20, iter_step=10,
                 valid_num=5000, test_flag=1, test_num=5000, SNR_Matrix=np.array([2.0, 2.1, 2.2, 2.3, 2.4, 2.5]))
        d.update(kw)
        return cls(**d)

    @property
    def out_filename(self) -> str:          # main_Base.py:68
        return f"C{self.core_idx}_{self.filename}"

    def path(self, rel: str) -> str:
        return os.path.join(self.root, rel[2:] if rel.startswith("./") else rel)

    @property
    def perf_filename(self) -> str:         # main_Base.py:90
        return self.path(f"./Weights/{self.out_filename}_Performance.txt")


def FTE(arr, precision=2):
    """Print_Functions.FTE (:227-228): format to exponential."""
    return [f"{val:.{precision}e}" for val in arr]


def perf_header(cfg: RunConfig, SNR_Matrix, M_proto, N_proto, Num_edge_proto, code_rate) -> str:
    """The block main_Base.py:91-103 prints at the top of Performance.txt (same text, typos included)."""
    c = cfg
    lines = [
        f"Decoding_type = {c.decoding_type} q_bit = {c.q_bit}",
        f"CN_weight_sharing = {c.sharing[0]} UCW_weight_sharing = {c.sharing[1]} VN_weight_sharing = {c.sharing[2]}",
        f"Init_CN_weight = {c.init_weight} Max_weight = {c.Max_weight} Min_weight = {c.Min_weight} "
        f"Init_VN_weight = {c.init_VN_weight}, init_from_file = {c.init_from_file}",
        f"samping_type = {c.sampling_type} systematic = {c.systematic}",
        f"z_value = {c.z_value} iters_max = {c.iters_max} fixed_iter = {c.fixed_iter} fixed_init = {c.fixed_init} "
        f"iter_step = {c.iter_step}",
        f"puncturing = {c.punct_start} ~ {c.punct_end}, shortening = {c.short_start} ~ {c.short_end}",
        f"etha_start = {c.etha_start} etha_discount = {c.etha_discount} etha_discount_step = {c.etha_discount_step}",
        f"loss_type = {c.loss_type} learn_rate_start = {c.learn_rate_start} learn_rate_discount = "
        f"{c.learn_rate_discount} learn_rate_step = {c.learn_rate_step}",
        f"batch_size = {c.batch_size} epochs = {c.epoch_input} training_num = {c.training_num} valid_flag = "
        f"{c.valid_flag} valid_num = {c.valid_num} test_flag = {c.test_flag} test_num = {c.test_num}",
        f"SNR_Matrix = {SNR_Matrix}",
        f"M_proto = {M_proto} N_proto = {N_proto} Num_edge_proto = {Num_edge_proto} code_rate = {code_rate}",
        "",
    ]
    return "\n".join(lines) + "\n"


def compute_opt_value(opt_value, opt_result_print, ber_last_SNR, fer_last_SNR, fer_SNR, loss_SNR):
    """Print_Functions.compute_opt_value (:167-181)."""
    cand = [np.sum(ber_last_SNR), np.sum(fer_last_SNR), np.sum(fer_SNR), np.sum(loss_SNR)][opt_result_print]
    if opt_value > cand:
        return cand, True
    return opt_value, False


def print_result(Results, opt_value, Perf_filename, out_filename, training_iter_end, opt_result_print,
                 opt_print_flag, test_time, root=".", quiet=False):
    """Print_Functions.print_result (:185-214): appends the Valid_Result / Test_Result block and copies the
    current weight file to ..._Opt_Weight_End{T}.txt when the validation metric improved."""
    label = "Test_Result" if test_time else "Valid_Result"
    if not test_time:
        opt_value, opt_print_flag = compute_opt_value(opt_value, opt_result_print, *Results[:4])
    elif opt_print_flag:
        opt_value, _ = compute_opt_value(100000, opt_result_print, *Results[:4])
    txt = (f"{label}\nBER_last: {FTE(Results[0, :])}\nFER_last: {FTE(Results[1, :])}\nFER: {FTE(Results[2, :])}\n"
           f"loss: {FTE(Results[3, :])}\nopt_value: {FTE([opt_value])}\n\n")
    with open(Perf_filename, "a") as out_file:
        out_file.write(txt)
    if not quiet:
        print(txt, end="")
    if not test_time and opt_print_flag:
        src = os.path.join(root, f"Weights/{out_filename}_Weight_End{training_iter_end}.txt")
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(root, f"Weights/{out_filename}_Opt_Weight_End{training_iter_end}.txt"))
    return opt_value, opt_print_flag


def process_data(cfg: RunConfig):
    """Main_Functions.process_data (:526-576): the three [Uncor] sets, truncated to training_num / valid_num /
    test_num rows, in the file's sign convention; all labels are the all-zero codeword."""
    if cfg.sampling_type != 1:
        return [], [], [], [], [], []
    f_train, f_valid, f_test = (cfg.path(p) for p in formats.uncor_filenames(cfg.filename))
    tr = formats.read_uncor(f_train, cfg.training_num)
    va = formats.read_uncor(f_valid, cfg.valid_num) if cfg.valid_flag == 1 else []
    te = formats.read_uncor(f_test, cfg.test_num) if cfg.test_flag == 1 else []
    zeros = lambda a: np.zeros(np.shape(a), dtype=np.int64) if len(a) else []   # noqa: E731
    return tr, zeros(tr), va, zeros(va), te, zeros(te)


def load_block_weights(cfg: RunConfig, graph: BaseGraph, training_iter_start: int, training_iter_end: int,
                       weights: Optional[formats.WeightSet] = None) -> formats.WeightSet:
    """Main_Functions.weight_init (:387-439) without the TF variables: rows [0, training_iter_start) come from
    ./Weights/{out}_Opt_Weight_End{start}.txt (:390-391, 419-422), rows of the block being trained from
    ..._In_Weight_End{iters_max}.txt when init_from_file == 1 (:388-389, 423-426) else the constants
    init_weight / init_VN_weight (:427-431).  `weights` overrides the files."""
    if weights is not None:
        return weights.rows(0, training_iter_end)
    blocks = {}
    prev = None
    if training_iter_start > 0:
        prev = formats.read_weights(cfg.path(formats.weights_filename(cfg.out_filename, training_iter_start, "Opt_Weight")))
    init = None
    if cfg.init_from_file == 1:
        init = formats.read_weights(cfg.path(formats.weights_filename(cfg.out_filename, cfg.iters_max, "In_Weight")))
    for i, code in enumerate(cfg.sharing):
        if code <= 0:
            continue
        width = formats.weight_width(code, i, graph.M, graph.N, graph.E)
        para_init = cfg.init_VN_weight if i == 2 else cfg.init_weight
        if para_init == -1:
            # tf.truncated_normal_initializer(mean = (min + max) / 2, stddev = 0.1) (:427-428): normal draws, those beyond
            # two standard deviations redrawn; TF's own generator is not reproduced, the seed is the run's noise seed
            rng = np.random.RandomState(1074 + cfg.seed_in + 104729 * i)
            draw = rng.normal(size=(training_iter_end, width))
            while True:
                bad = np.abs(draw) > 2.0
                if not bad.any():
                    break
                draw[bad] = rng.normal(size=int(bad.sum()))
            rows = ((cfg.Min_weight + cfg.Max_weight) / 2 + 0.1 * draw).astype(np.float32)
            if code in (4, 5):   # temporal sharing: one variable serves every iteration >= fixed_iter (:411-414)
                rows[cfg.fixed_iter:] = rows[min(cfg.fixed_iter, training_iter_end - 1)]
        else:
            rows = np.full((training_iter_end, width), para_init, dtype=np.float32)
        if prev is not None:
            rows[:training_iter_start] = prev.blocks[i][:training_iter_start]
        if init is not None:
            rows[training_iter_start:] = init.blocks[i][training_iter_start:training_iter_end]
        blocks[i] = np.clip(rows, cfg.Min_weight, cfg.Max_weight)           # the clip constraint (:434)
    return formats.WeightSet(list(cfg.sharing), blocks)


def evaluate(cfg: RunConfig, weights: Optional[formats.WeightSet] = None, training_iter_end: Optional[int] = None,
             device: Optional[int] = None, group=None, uncor_path: Optional[str] = None, quiet: bool = False):
    """The evaluation pass of one training block of main_Base.py / main_Post.py (epoch 0): returns a dict with
    `valid` / `test` Results f32[4, nSNR] (rows BER_last, FER_last, FER, loss = 0) and the seconds they took."""
    from .decoder import NMSDecoder, check_params
    from .montecarlo import compute_results
    c = cfg
    SNR_Matrix = check_params(c.sampling_type, c.SNR_Matrix, c.sharing, c.iters_max, c.fixed_iter, c.iter_step)
    proto = formats.read_base_graph(c.path(f"./BaseGraph/{c.filename}.txt"))              # main_Base.py:67
    g = BaseGraph(proto, c.z_value, (c.punct_start, c.punct_end), (c.short_start, c.short_end), name=c.filename)
    SNR_sigma = g.sigma(SNR_Matrix)
    T = c.fixed_iter + c.iter_step if training_iter_end is None else int(training_iter_end)
    ws = load_block_weights(c, g, T - c.iter_step if training_iter_end is None else 0, T, weights)
    data = process_data(c)
    os.makedirs(os.path.dirname(c.perf_filename), exist_ok=True)
    rank = 0
    if group is not None or _dist_ready():
        import torch.distributed as dist
        rank = dist.get_rank(group)
    if rank == 0:
        with open(c.perf_filename, "w") as fh:
            fh.write(perf_header(c, SNR_Matrix, g.M, g.N, g.E, g.rate_ref))
    if c.opt_result_print == 3:
        raise NotImplementedError("opt_result_print = 3 selects the best epoch by validation LOSS (Print_Functions.py:151,161, "
                                  "167-181); compute_results here returns no loss row -- use 0, 1 or 2 (BER / FER_last / FER)")
    # systematic = 1: metrics over the first N - M proto columns only (main_Base.py:83-86); temporal sharing needs fixed_iter
    dec = NMSDecoder(g, ws, iters=T, decoding_type=c.decoding_type, q_bit=c.q_bit, clip_llr=c.clip_LLR, device=device,
                     systematic=c.systematic, fixed_iter=c.fixed_iter)
    seed = 1074 + c.seed_in                                                                # noise_seed, :70
    out = {"SNR_Matrix": SNR_Matrix, "SNR_sigma": SNR_sigma, "iters": T, "decoder": dec}
    opt_valid, opt_flag = 100000, False
    if c.valid_flag > 0:
        if uncor_path is None:
            uncor_path = c.path("./Uncor.txt")                                            # Print_Functions.py:122
        res, sec = compute_results(dec, c.valid_num, data[2], SNR_sigma, c.batch_size, c.sampling_type, seed=seed,
                                   uncor_path=uncor_path, group=group)
        out["valid"], out["time_valid"] = res, sec
        if rank == 0:
            opt_valid, opt_flag = print_result(res, opt_valid, c.perf_filename, c.out_filename, T, c.opt_result_print,
                                               opt_flag, False, root=c.root, quiet=quiet)
    if c.sampling_type == 1 and c.test_num > 0 and c.test_flag == 1:
        res, sec = compute_results(dec, c.test_num, data[4], SNR_sigma, c.batch_size, c.sampling_type, seed=seed,
                                   group=group)
        out["test"], out["time_test"] = res, sec
        if rank == 0:
            print_result(res, 100000, c.perf_filename, c.out_filename, T, c.opt_result_print, opt_flag, True,
                         root=c.root, quiet=quiet)
    if rank == 0:
        line = (f"Running time (Train/Valid/Test): {0.0:.2f}/{out.get('time_valid', 0.0):.2f}/"
                f"{out.get('time_test', 0.0):.2f}\n")
        with open(c.perf_filename, "a") as fh:
            fh.write(line + "\n")
    return out


def _dist_ready() -> bool:
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()
    except Exception:
        return False


def split_uncor(path: str, filename: str, n_train: int, n_valid: int, n_test: int, root: str = ".") -> List[str]:
    """The manual step between collection and main_Post.py: ./Uncor.txt -> Inputs/[Uncor]_{filename}.txt,
    _Valid.txt, _Test.txt (consecutive, disjoint row ranges; same text lines)."""
    with open(path) as fh:
        lines = fh.readlines()
    need = n_train + n_valid + n_test
    if len(lines) < need:
        raise ValueError(f"{path}: {len(lines)} rows < {need} requested")
    outs = [os.path.join(root, p[2:]) for p in formats.uncor_filenames(filename)]
    os.makedirs(os.path.dirname(outs[0]), exist_ok=True)
    lo = 0
    for p, n in zip(outs, (n_train, n_valid, n_test)):
        with open(p, "w") as fh:
            fh.writelines(lines[lo:lo + n])
        lo += n
    return outs
