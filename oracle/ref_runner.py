"""Oracle A -- the reference's OWN decode code, imported unmodified.  TEST INFRASTRUCTURE ONLY.

This module exists only in the development container: it needs `/root/reference`
(read-only; absent on the GPU box).  It is used by `tests/golden/make_golden.py`
to mint the committed golden vectors and by the container-only tests that pin
`oracle/nms_oracle.py` / `oracle/nms_oracle.c` (Oracle B) against the reference.
Nothing in the product (`ldpc_error_floor_b200/`), in `bench.py` or in the
`-m gpu` tests imports it.

What runs: `Main_Functions.build_neural_network` (Main_Functions.py:157-385),
`init_parameter` (:8-38), `init_connecting_matrix` (:46-150) and
`Print_Functions.compute_results` (:130-165) exactly as shipped, on top of the
numpy stand-in for `tensorflow.compat.v1` in `oracle/tf_shim`.  TensorFlow
itself cannot be installed here, so "the reference" below always means
"the reference's Python under the shim" -- say so wherever a number is quoted.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("LDPC_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Main_Functions.py"))


def load_reference():
    """Import (Main_Functions, Print_Functions) from the read-only reference tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    mf = importlib.import_module("Main_Functions")
    pf = importlib.import_module("Print_Functions")
    return mf, pf


class ReferenceDecoder:
    """Unrolls `build_neural_network` T times the way main_Base.py:136-137 does,
    but eagerly (numpy arrays instead of TF tensors)."""

    def __init__(self, code_proto, z, sharing, weights, T, decoding_type=2, q_bit=5,
                 clip_llr=20.0, punct=(0, 0), short=(0, 0), snr_db=(0.0,), target_node=None, fixed_iter=0):
        self.mf, self.pf = load_reference()
        self.code_proto = np.asarray(code_proto, dtype=int)
        self.z = int(z)
        self.sharing = list(sharing)
        self.T = int(T)
        self.decoding_type = int(decoding_type)
        self.q_bit = int(q_bit)
        self.clip_llr = float(clip_llr)
        self.punct = tuple(punct)
        self.short = tuple(short)
        (self.M, self.N, self.code_base, self.cn_deg, self.vn_deg, self.E, self.code_rate,
         self.snr_sigma) = self.mf.init_parameter(self.code_proto, np.asarray(snr_db, dtype=float),
                                                  self.z, punct[0], punct[1], short[0], short[1])
        self.E = int(self.E)
        self.target_node = self.N if target_node is None else int(target_node)   # main_Base.py:83-86 (systematic)
        self.fixed_iter = int(fixed_iter)                                        # temporal sharing (code 4)
        self.mats = self.mf.init_connecting_matrix(self.code_proto, self.code_base, self.N, self.M,
                                                   self.E, self.z, self.vn_deg, self.cn_deg,
                                                   punct[0], punct[1])
        # weights: {type_idx: array[T, width]}; names follow weight_init (Main_Functions.py:433)
        self.vars = {}
        for i, code in enumerate(self.sharing):
            if code > 0:
                w = np.asarray(weights[i], dtype=np.float32)
                rows = self.T if code in (1, 2, 3) else self.fixed_iter + 1       # weight_init :411-414
                for t in range(rows):
                    self.vars[f"var_{i}_{t}"] = np.ascontiguousarray(w[t]).reshape(-1)

    def decode(self, xa, ya=None):
        """xa: f32 [B, N, z] channel LLRs (log p1/p0).  Returns dict with
        app [T, B, N*z] (= ya_output{t}), c2v [T, B, z, E] (= LLRa{t+1})."""
        xa = np.asarray(xa, dtype=np.float32)
        B = xa.shape[0]
        if ya is None:
            ya = np.zeros((B, self.N * self.z), dtype=np.float32)
        net = dict(self.vars)
        net["xa"] = xa
        net["ya"] = ya
        net["LLRa0"] = np.zeros((B, self.z, self.E), dtype=np.float32)  # main_Base.py:126
        for t in range(self.T):
            net = self.mf.build_neural_network(
                net, self.sharing, self.decoding_type, 2, 2, self.target_node, t, self.T, self.fixed_iter, 0, 0, self.T,
                self.N, self.M, self.E, self.z, B, *self.mats, self.q_bit, self.clip_llr)
        app = np.stack([net[f"ya_output{t}"] for t in range(self.T)], axis=0)
        c2v = np.stack([net[f"LLRa{t + 1}"] for t in range(self.T)], axis=0)
        return {"app": app.astype(np.float32), "c2v": c2v.astype(np.float32),
                "ya_output_all": np.asarray(net["ya_output_all"], dtype=np.float32)}

    # -- a fake `sess` so Print_Functions.compute_results runs unmodified (:148) --
    class _Sess:
        def __init__(self, outer):
            self.outer = outer

        def run(self, fetches=None, feed_dict=None):
            xa = feed_dict["xa"]
            out = self.outer.decode(xa)["ya_output_all"]
            if isinstance(fetches, (list, tuple)):
                return [out, 0.0]
            return out

    def compute_results(self, sample_num, word_seed, noise_seed, batch_size, sampling_type=2,
                        input_llr=None, input_codeword=None, cwd=None):
        """Runs Print_Functions.compute_results (:130-165) as main_Base.py:177 calls it."""
        net_dict = {"ya_output_all": "ya_output_all", "lossa": "lossa", "xa": "xa", "ya": "ya",
                    "etha": "etha", "learn_rate": "learn_rate"}
        word_random = np.random.RandomState(word_seed)
        noise_random = np.random.RandomState(noise_seed)
        old = os.getcwd()
        if cwd is not None:
            os.chdir(cwd)
        try:
            res, took = self.pf.compute_results(
                sample_num, input_llr if input_llr is not None else [],
                input_codeword if input_codeword is not None else [], self.snr_sigma,
                word_random, noise_random, batch_size, sampling_type, self.N, self.M, self.z, True,
                self.T, self._Sess(self), net_dict, 0, self.decoding_type, self.punct[0],
                self.punct[1], self.short[0], self.short[1], self.q_bit, self.clip_llr)
        finally:
            os.chdir(old)
        return res, took
