"""Base-graph compiler front-end (host side of ldpc_graph_create in include/ldpc_b200.h).

Replaces Main_Functions.init_parameter (Main_Functions.py:8-38) and the index conventions of
init_connecting_matrix (:46-150).  The compilation itself is native code (csrc/ldpc_capi.cu);
this class owns the handle and exposes the reference's scalars under the reference's names.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib, formats


class BaseGraph:
    """A proto / parity-check matrix with its lifting size and puncture / shorten ranges.

    proto: int [M, N], -1 = no edge, else circulant shift (used mod z); z = 1 for non-QC codes.
    punct / short: (start, end) 1-based inclusive bit ranges, (0, 0) = none (main_Base.py:31-34).
    """

    def __init__(self, proto, z: int, punct: Tuple[int, int] = (0, 0), short: Tuple[int, int] = (0, 0),
                 name: str = ""):
        self.proto = np.ascontiguousarray(np.asarray(proto), dtype=np.int32)
        if self.proto.ndim != 2:
            raise ValueError("proto must be a 2-D matrix")
        self.z = int(z)
        self.punct = (int(punct[0]), int(punct[1]))
        self.short = (int(short[0]), int(short[1]))
        self.name = name
        lib = _lib.load()
        self._h = ctypes.c_void_p()
        M, N = self.proto.shape
        _lib.check(lib.ldpc_graph_create(self.proto.ctypes.data, M, N, self.z, self.punct[0], self.punct[1],
                                         self.short[0], self.short[1], ctypes.byref(self._h)))
        info = _lib.GraphInfo()
        _lib.check(lib.ldpc_graph_info(self._h, ctypes.byref(info)))
        self.info = info
        self.M, self.N, self.E = info.M, info.N, info.E
        self.NZ = self.N * self.z
        row = np.empty(self.E, np.int32)
        col = np.empty(self.E, np.int32)
        shift = np.empty(self.E, np.int32)
        _lib.check(lib.ldpc_graph_edges(self._h, row.ctypes.data, col.ctypes.data, shift.ctypes.data))
        self.edge_row, self.edge_col, self.edge_shift = row, col, shift   # E(C) order

    @classmethod
    def from_file(cls, path: str, z: Optional[int] = None, punct=(0, 0), short=(0, 0)) -> "BaseGraph":
        """BaseGraph/*.txt (format F1).  For the 5G files z / punct / short default to what the
        file name encodes."""
        proto = formats.read_base_graph(path)
        meta = formats.parse_5g_name(path)
        if meta is not None:
            z = meta["z"] if z is None else z
            if tuple(punct) == (0, 0) and tuple(short) == (0, 0):
                punct, short = meta["punct"], meta["short"]
        if z is None:
            raise ValueError("z must be given for this base graph")
        return cls(proto, z, punct, short, name=os.path.splitext(os.path.basename(path))[0])

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().ldpc_graph_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- the reference's scalars (Main_Functions.py:18-29)
    @property
    def code_base(self) -> np.ndarray:
        return (self.proto != -1).astype(self.proto.dtype)

    @property
    def cn_deg(self) -> np.ndarray:
        return self.code_base.sum(axis=1)

    @property
    def vn_deg(self) -> np.ndarray:
        return self.code_base.sum(axis=0)

    @property
    def rate_ref(self) -> float:
        return float(self.info.rate_ref)

    @property
    def rate_true(self) -> float:
        return float(self.info.rate_true)

    @property
    def k_true(self) -> int:
        return int(self.info.k_true)

    @property
    def n_true(self) -> int:
        return int(self.info.n_true)

    def sigma(self, snr_db: Sequence[float], use_ref_rate: bool = True) -> np.ndarray:
        """sqrt(1/(2 R 10^(snr/10))) (Main_Functions.py:35-36); R = the reference's rate by default."""
        snr = np.ascontiguousarray(np.atleast_1d(np.asarray(snr_db, dtype=np.float64)))
        out = np.empty_like(snr)
        _lib.check(_lib.load().ldpc_graph_sigma(self._h, snr.ctypes.data, snr.size, 1 if use_ref_rate else 0,
                                                out.ctypes.data))
        return out

    def syndrome_ok(self, hard_bits: np.ndarray) -> np.ndarray:
        """Host check used by tests: hard_bits bool/int [B, N*z] -> bool [B] all checks satisfied."""
        b = np.asarray(hard_bits).reshape(-1, self.N, self.z).astype(np.uint8)
        s = np.zeros((b.shape[0], self.M, self.z), dtype=np.uint8)
        for e in range(self.E):
            s[:, self.edge_row[e], :] ^= np.roll(b[:, self.edge_col[e], :], -int(self.edge_shift[e]), axis=1)
        return ~s.reshape(b.shape[0], -1).any(axis=1)


def init_parameter(code_Proto, SNR_Matrix, z_value, punct_start, punct_end, short_start, short_end):
    """Drop-in for Main_Functions.init_parameter (:8-38): same arguments, same 8-tuple."""
    g = BaseGraph(code_Proto, z_value, (punct_start, punct_end), (short_start, short_end))
    sigma = g.sigma(np.asarray(SNR_Matrix, dtype=np.float64))
    return g.M, g.N, g.code_base, g.cn_deg, g.vn_deg, g.E, g.rate_ref, sigma
