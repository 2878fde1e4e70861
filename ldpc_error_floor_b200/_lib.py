"""ctypes binding of libldpc_b200.so (include/ldpc_b200.h).  No fallback: if the shared
library is missing, every entry point raises -- the product never decodes on the CPU."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libldpc_b200.so")
CSRC = os.path.join(_HERE, "csrc")

NUM_COUNTERS = 8
COUNTER_NAMES = ("frames", "frame_err_last", "frame_err_any", "bit_err_last", "iters",
                 "synd_fail", "undetected", "harvested")

FLAG_SYND_OK = 1
FLAG_UNCOR_ANY = 2
FLAG_UNCOR_LAST = 4
FLAG_SYND_OK_EVER = 8

HARVEST_NONE, HARVEST_UNCOR_ANY, HARVEST_UNCOR_LAST, HARVEST_SYND_FAIL = 0, 1, 2, 3


class LdpcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ldpc_b200 error {code}: {msg}")
        self.code = code


class GraphInfo(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int32), ("N", ctypes.c_int32), ("z", ctypes.c_int32), ("E", ctypes.c_int32),
                ("max_dc", ctypes.c_int32), ("max_dv", ctypes.c_int32),
                ("n_ref", ctypes.c_int32), ("k_ref", ctypes.c_int32),
                ("n_true", ctypes.c_int32), ("k_true", ctypes.c_int32),
                ("rate_ref", ctypes.c_double), ("rate_true", ctypes.c_double)]


# every symbol include/ldpc_b200.h declares: (name, restype, argtypes)
_P = ctypes.c_void_p
_I32, _I64, _U32, _U64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
SYMBOLS = {
    "ldpc_last_error": (ctypes.c_char_p, []),
    "ldpc_version": (ctypes.c_int, []),
    "ldpc_graph_create": (ctypes.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, ctypes.POINTER(_P)]),
    "ldpc_graph_destroy": (ctypes.c_int, [_P]),
    "ldpc_graph_info": (ctypes.c_int, [_P, ctypes.POINTER(GraphInfo)]),
    "ldpc_graph_edges": (ctypes.c_int, [_P, _P, _P, _P]),
    "ldpc_graph_sigma": (ctypes.c_int, [_P, _P, _I32, _I32, _P]),
    "ldpc_decoder_create": (ctypes.c_int, [_P, _P, _I32, _P, _P, _P, _I32, _I32, ctypes.c_float, _I32,
                                           ctypes.POINTER(_P)]),
    "ldpc_decoder_create2": (ctypes.c_int, [_P, _P, _I32, _P, _P, _P, _I32, _I32, ctypes.c_float, _I32, _I32,
                                            ctypes.POINTER(_P)]),
    "ldpc_decoder_destroy": (ctypes.c_int, [_P]),
    "ldpc_decoder_uses_packed_kernel": (ctypes.c_int, [_P]),
    "ldpc_decoder_kernel_name": (ctypes.c_char_p, [_P]),
    "ldpc_decoder_geometry": (ctypes.c_int, [_P, ctypes.POINTER(_I32), ctypes.POINTER(_I32),
                                             ctypes.POINTER(_I32), ctypes.POINTER(_I32)]),
    "ldpc_decoder_launch_info": (ctypes.c_int, [_P, _I32, ctypes.POINTER(_I32), ctypes.POINTER(_I32), ctypes.POINTER(_I32),
                                                ctypes.POINTER(_I32), ctypes.c_char_p, _I32]),
    "ldpc_decoder_mc_info": (ctypes.c_int, [_P, ctypes.POINTER(_I32), ctypes.POINTER(_I32), ctypes.POINTER(_I32),
                                            ctypes.POINTER(_I32), ctypes.POINTER(_I32), ctypes.c_char_p, _I32]),
    "ldpc_decode": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P]),
    "ldpc_decode_host": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _P, _I32, _P, _P, _P, _P]),
    "ldpc_llr_generate_cw": (ctypes.c_int, [_P, ctypes.c_double, _I64, _U64, _U64, _P, _I64, _P, _P]),
    "ldpc_decode_cw": (ctypes.c_int, [_P, _P, _P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "ldpc_decoder_q8_step": (ctypes.c_float, [_P]),
    "ldpc_decode_q8": (ctypes.c_int, [_P, _P, ctypes.c_float, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "ldpc_decode_q8_host": (ctypes.c_int, [_P, _P, ctypes.c_float, _I64, _I32, _I32, _P, _P, _P, _P]),
    "ldpc_decode_host_stats": (ctypes.c_int, [_P, _P]),
    "ldpc_pack_q8_host": (ctypes.c_int, [_P, _P, _I64, _P, _P]),
    "ldpc_pack_q8_values": (ctypes.c_int, [_P, _I64, ctypes.c_float, ctypes.c_float, _I32, _P, _P]),
    "ldpc_llr_generate": (ctypes.c_int, [_P, ctypes.c_double, _I64, _U64, _U64, _P, _P]),
    "ldpc_normal_probe": (ctypes.c_int, [_I32, _U64, _U64, _I64, _I32, _P, _P, _P]),
    "ldpc_mc_run": (ctypes.c_int, [_P, ctypes.c_double, _I64, _U64, _U64, _I32, _I32, _I32, _P, _P, _P, _U32, _P]),
    "ldpc_mc_run_staged": (ctypes.c_int, [_P, ctypes.c_double, _I64, _U64, _U64, _I32, _I32, _I32, _P, _P, _P, _U32, _P, _P, _U32, _P]),
    "ldpc_mc_run_host": (ctypes.c_int, [_P, ctypes.c_double, _I64, _U64, _U64, _I32, _I32, _I32, _P, _P, _U32,
                                        ctypes.POINTER(_U32)]),
    "ldpc_post_decode": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P]),
    "ldpc_decoder_set_weights": (ctypes.c_int, [_P, _P, _P, _P]),
    "ldpc_train_grad": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _I32, ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                                       _P, _P, _P, _P]),
    "ldpc_jit_prebuild": (ctypes.c_int, [_P, _I32, _I32, _I32]),
    "ldpc_alu_peak_probe": (ctypes.c_int, [_I32, _I32, ctypes.POINTER(ctypes.c_double)]),
    "ldpc_launch_count": (ctypes.c_uint64, []),
}

_lib = None


def build(force: bool = False, jobs: int = 8) -> str:
    """Compile the CUDA extension in-tree for sm_100a (`make -C csrc`)."""
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", CSRC, f"-j{jobs}"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LdpcError(-3, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                                f"g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ldpc_last_error()
        raise LdpcError(rc, msg.decode() if msg else "unknown")


ALU_PROBE_KINDS = {"ffma": 0, "fmnmx": 1, "lop3": 2, "iadd": 3, "hfma2": 4, "hmnmx2": 5, "fadd": 6}


class HostStats(ctypes.Structure):
    """ldpc_host_stats_t (include/ldpc_b200.h)."""
    _fields_ = [("threads", ctypes.c_int32), ("chunks_total", ctypes.c_int32), ("chunks_q8", ctypes.c_int32), ("chunks_f32", ctypes.c_int32),
                ("chunks_unencodable", ctypes.c_int32), ("float_share", ctypes.c_double), ("s_pack", ctypes.c_double),
                ("s_wait", ctypes.c_double), ("s_wait_feed", ctypes.c_double), ("s_copy_out", ctypes.c_double), ("s_total", ctypes.c_double),
                ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64)]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


def pack_q8_values(x, step: float, qmax: float, lossless: bool):
    """ldpc_pack_q8_values: float32 values -> int8 multiples of `step` (host threads, no GPU).  Returns (int8 array of
    x's shape, number of values without an int8 form)."""
    import numpy as np
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.int8)
    bad = ctypes.c_int64(0)
    check(load().ldpc_pack_q8_values(x.ctypes.data, x.size, float(step), float(qmax), 1 if lossless else 0,
                                     out.ctypes.data, ctypes.byref(bad)))
    return out, int(bad.value)


def jit_prebuild(proto, z: int) -> int:
    """Compile the run-time specialised kernels of a base graph into the on-disk cache (no GPU needed)."""
    import numpy as np
    p = np.ascontiguousarray(proto, dtype=np.int32)
    rc = load().ldpc_jit_prebuild(p.ctypes.data, p.shape[0], p.shape[1], int(z))
    if rc < 0:
        check(rc)
    return rc


def alu_peak_probe(device: int = 0, kinds=("ffma", "fmnmx", "lop3", "hfma2", "hmnmx2")) -> dict:
    """Measured instruction-issue rates (T lane-ops/s) of the SM pipes on `device` (ldpc_alu_peak_probe)."""
    out = {}
    for k in kinds:
        v = ctypes.c_double(0.0)
        check(load().ldpc_alu_peak_probe(int(device), ALU_PROBE_KINDS[k], ctypes.byref(v)))
        out[k] = v.value / 1e12
    return out
