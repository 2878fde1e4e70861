"""Text data formats of the reference, kept byte-compatible so existing files drop in.

F1  BaseGraph/*.txt          proto / parity-check matrix      (main_Base.py:67)
F2  Weights/*.txt            per-iteration NMS weights         (Print_Functions.py:74-96 writes,
                                                                Main_Functions.py:387-439 reads)
F3  Inputs/[Uncor]_*.txt     harvested uncorrected words       (Print_Functions.py:120-126 writes,
                                                                Main_Functions.py:526-576 and
                                                                Print_Functions.py:6-10 read)
"""
from __future__ import annotations

import io
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# ------------------------------------------------------------------ F1  base graphs


def read_base_graph(path: str) -> np.ndarray:
    """M lines x N tab-separated ints, -1 = no edge, otherwise the circulant shift
    (used mod z).  LF or CRLF, no trailing newline required.  Same call as
    main_Base.py:67 (`np.loadtxt(path, int, delimiter='\\t')`) but tolerant of a single row."""
    with open(path, "r", newline=None) as fh:
        rows = [ln.strip() for ln in fh.read().splitlines() if ln.strip() != ""]
    mat = [[int(tok) for tok in ln.split("\t")] for ln in rows]
    widths = {len(r) for r in mat}
    if len(widths) != 1:
        raise ValueError(f"{path}: ragged base graph rows {sorted(widths)}")
    return np.asarray(mat, dtype=np.int32)


def write_base_graph(path: str, proto: np.ndarray, crlf: bool = False) -> None:
    eol = "\r\n" if crlf else "\n"
    txt = eol.join("\t".join(str(int(v)) for v in row) for row in np.asarray(proto))
    with open(path, "w", newline="") as fh:
        fh.write(txt)  # the shipped files carry no trailing newline


_5G_NAME = re.compile(r"n_dec(\d+)_n(\d+)_k(\d+)_z(\d+)_s(\d+)_(\d+)")


def parse_5g_name(filename: str) -> Optional[dict]:
    """5G file-name convention: n_dec = N*z, n = transmitted, k = info,
    s{a}_{b} = shortening range (1-based inclusive); the first two proto columns are punctured."""
    m = _5G_NAME.search(os.path.basename(filename))
    if not m:
        return None
    n_dec, n, k, z, ss, se = (int(v) for v in m.groups())
    return {"n_dec": n_dec, "n": n, "k": k, "z": z, "punct": (1, 2 * z), "short": (ss, se)}


# --------------------------------------------------------------------- F2  weights

_WIDTH_KIND = ("cn", "ucn", "vn")


def weight_width(code: int, kind_idx: int, M: int, N: int, E: int) -> int:
    """Row length for a sharing code (Main_Functions.py:397-405)."""
    if code in (1, 4):
        return E
    if code in (2, 5):
        return M if kind_idx in (0, 1) else N
    if code == 3:
        return 1
    raise ValueError(f"sharing code {code} carries no weights")


@dataclass
class WeightSet:
    """sharing = [CN, UCN, VN] codes (main_Base.py:24-25); blocks[i] is f32 [T, width_i]
    for every i whose code is non-zero."""
    sharing: List[int]
    blocks: Dict[int, np.ndarray] = field(default_factory=dict)

    @property
    def iterations(self) -> int:
        return 0 if not self.blocks else int(next(iter(self.blocks.values())).shape[0])

    def rows(self, start: int, stop: int) -> "WeightSet":
        """Iterations [start, stop): e.g. rows(0, 20) of a 50-row boosted file = the base decoder."""
        return WeightSet(list(self.sharing), {i: b[start:stop].copy() for i, b in self.blocks.items()})


def expand_temporal(ws: "WeightSet", T: int, fixed_iter: int) -> "WeightSet":
    """Temporal sharing (codes 4 / 5, main_Base.py:24): iterations t >= fixed_iter reuse the variables of
    iteration fixed_iter (Main_Functions.py:170-174, 299-304), so only fixed_iter + 1 rows exist
    (weight_init :411-414).  Returns the equivalent per-iteration set: code 4 -> 1 (per edge) with T rows;
    a UCN block of code 4 is dropped because build_neural_network has no UCN branch for it (:299-304).
    Files written by print_weight already hold the T expanded rows (Print_Functions.py:87-94) and pass through."""
    sharing, blocks = list(ws.sharing), {}
    for i, code in enumerate(ws.sharing):
        if code <= 0:
            continue
        b = np.asarray(ws.blocks[i], dtype=np.float32)
        if code in (4, 5):
            if i == 1:
                sharing[1] = 0
                continue
            if b.shape[0] < T:
                if b.shape[0] < fixed_iter + 1:
                    raise ValueError(f"temporal block {i}: {b.shape[0]} rows < fixed_iter + 1 = {fixed_iter + 1}")
                b = b[np.minimum(np.arange(T), fixed_iter)]
            sharing[i] = 1 if code == 4 else 2
        blocks[i] = b
    return WeightSet(sharing, blocks)


def read_weights(path: str) -> WeightSet:
    """Header "s0 s1 s2", blank line, then one blank-line-terminated block of T lines per
    non-zero sharing code (tab-separated float32 reprs).  The reference addresses the same
    lines by absolute row number (Main_Functions.py:419-426); T is inferred from the block."""
    with open(path, "r", newline=None) as fh:
        lines = fh.read().splitlines()
    if not lines:
        raise ValueError(f"{path}: empty weight file")
    sharing = [int(tok) for tok in lines[0].split()]
    if len(sharing) != 3:
        raise ValueError(f"{path}: header must hold three sharing codes, got {lines[0]!r}")
    blocks_raw: List[List[List[float]]] = []
    cur: List[List[float]] = []
    for ln in lines[1:]:
        if ln.strip() == "":
            if cur:
                blocks_raw.append(cur)
                cur = []
            continue
        cur.append([float(tok) for tok in ln.split("\t")])
    if cur:
        blocks_raw.append(cur)
    active = [i for i, c in enumerate(sharing) if c > 0]
    if len(blocks_raw) != len(active):
        raise ValueError(f"{path}: {len(blocks_raw)} weight blocks for sharing {sharing}")
    blocks = {}
    for i, raw in zip(active, blocks_raw):
        widths = {len(r) for r in raw}
        if len(widths) != 1:
            raise ValueError(f"{path}: ragged rows in the {_WIDTH_KIND[i]} block")
        blocks[i] = np.asarray(raw, dtype=np.float32)
    T = {b.shape[0] for b in blocks.values()}
    if len(T) != 1:
        raise ValueError(f"{path}: blocks disagree on the iteration count {sorted(T)}")
    return WeightSet(sharing, blocks)


def write_weights(path: str, ws: WeightSet) -> None:
    """Same bytes as Print_Functions.print_weight (:74-96): `print("s0 s1 s2\\n")`, then
    np.savetxt(fmt='%s', delimiter='\\t') of each float32 row, a blank line after each block."""
    buf = io.StringIO()
    buf.write("{0} {1} {2}\n\n".format(*ws.sharing))
    for i, code in enumerate(ws.sharing):
        if code > 0:
            for row in np.asarray(ws.blocks[i], dtype=np.float32):
                buf.write("\t".join(str(v) for v in row) + "\n")   # '%s' of np.float32
            buf.write("\n")
    with open(path, "w") as fh:
        fh.write(buf.getvalue())


def weights_filename(out_filename: str, iters: int, kind: str = "Weight") -> str:
    """./Weights/{out}_Weight_End{T}.txt, _Opt_Weight_End{T}, _In_Weight_End{T}
    (main_Base.py:68; Main_Functions.py:389-391; Print_Functions.py:75,199)."""
    return f"./Weights/{out_filename}_{kind}_End{iters}.txt"


# ------------------------------------------------------------ F3  uncorrected words


def read_uncor(path: str, limit: Optional[int] = None) -> np.ndarray:
    """Returns the stored rows WITHOUT the three leading columns, f32 [n, N*z], still in the
    file's sign convention (log p0/p1) -- exactly what process_data returns
    (Main_Functions.py:530-538).  Use `uncor_to_llr` to obtain decoder inputs."""
    data = np.loadtxt(path, dtype=np.float32, delimiter="\t", ndmin=2)
    if data.shape[1] > 3:
        data = data[:, 3:]
    if limit is not None:
        if data.shape[0] < limit:
            raise ValueError(f"{path}: {data.shape[0]} rows < requested {limit}")
        data = data[:limit]
    return np.ascontiguousarray(data)


def uncor_to_llr(rows: np.ndarray, N: int, z: int) -> np.ndarray:
    """Print_Functions.read_uncor_llr (:6-10): negate (file holds log p0/p1) and reshape [B,N,z]."""
    rows = np.asarray(rows, dtype=np.float32)
    return -rows.reshape(rows.shape[0], N, z)


def append_uncor(path: str, llr: np.ndarray) -> int:
    """Print_Functions.write_uncor_file (:120-126): append one line per word, three leading
    0.0 columns, then the NEGATED decoder-input LLRs, fmt '%.1f', tab-separated."""
    llr = np.asarray(llr, dtype=np.float32)
    flat = -llr.reshape(llr.shape[0], -1)
    with open(path, "a") as fh:
        np.savetxt(fh, np.concatenate((np.zeros((flat.shape[0], 3)), flat), axis=1),
                   fmt="%.1f", delimiter="\t")
    return flat.shape[0]


def uncor_filenames(filename: str) -> Tuple[str, str, str]:
    """Inputs/[Uncor]_{filename}{,_Valid,_Test}.txt (Main_Functions.py:529,543,559)."""
    base = f"./Inputs/[Uncor]_{filename}"
    return base + ".txt", base + "_Valid.txt", base + "_Test.txt"


# ------------------------------------------ F3b  binary sidecar of the uncorrected-word sets ("next" row N2)
# The text format costs ~6 bytes and a float parse per value; at 10^6 words it is the bottleneck.  On the
# quantised path every stored value is a small multiple of the quantiser step, so a word is N*z int8.
# Layout (little endian): magic "LDPCQ8\0\1", u32 version, u32 words-per-row (N*z), u64 rows, f32 step,
# f32 Eb/N0 in dB at which the words were harvested (NaN = unknown; the text format cannot record it),
# u64 Philox seed, 24 reserved bytes (64-byte header); then rows * N*z int8 = DECODER-INPUT LLR / step
# (log p1/p0: NOT negated, unlike the text file).

_Q8_MAGIC = b"LDPCQ8\x00\x01"
_Q8_HEADER = 64


def llr_to_q8(llr: np.ndarray, step: float = 0.5) -> np.ndarray:
    """Decoder-input LLRs -> int8 counts; raises if a value is off the grid or out of range."""
    llr = np.asarray(llr, dtype=np.float32)
    q = llr.reshape(llr.shape[0], -1) / np.float32(step)
    r = np.rint(q)
    if not np.array_equal(q, r) or np.abs(r).max(initial=0) > 127:
        raise ValueError("LLRs are not int8 multiples of the step")
    return r.astype(np.int8)


def q8_to_llr(words: np.ndarray, step: float = 0.5) -> np.ndarray:
    return np.asarray(words, dtype=np.int8).astype(np.float32) * np.float32(step)


def write_uncor_q8(path: str, words: np.ndarray, step: float = 0.5, snr_db: float = float("nan"), seed: int = 0,
                   append: bool = False) -> int:
    """Write / append int8 words [n, N*z] (decoder-input sign convention)."""
    import struct
    words = np.ascontiguousarray(np.asarray(words, dtype=np.int8))
    n, width = words.shape
    if append and os.path.exists(path):
        with open(path, "r+b") as fh:
            head = fh.read(_Q8_HEADER)
            if head[:8] != _Q8_MAGIC:
                raise ValueError(f"{path}: not an LDPCQ8 file")
            _, w0, rows = struct.unpack_from("<IIQ", head, 8)
            if w0 != width:
                raise ValueError(f"{path}: row width {w0} != {width}")
            fh.seek(0, os.SEEK_END)
            fh.write(words.tobytes())
            fh.seek(16)
            fh.write(struct.pack("<Q", rows + n))
        return rows + n
    head = _Q8_MAGIC + struct.pack("<IIQffQ", 1, width, n, step, snr_db, seed & (2 ** 64 - 1))
    with open(path, "wb") as fh:
        fh.write(head.ljust(_Q8_HEADER, b"\0"))
        fh.write(words.tobytes())
    return n


def read_uncor_q8(path: str, limit: Optional[int] = None, offset: int = 0):
    """Returns (int8 [n, N*z] memory-mapped, meta dict(step, snr_db, seed, rows))."""
    import struct
    with open(path, "rb") as fh:
        head = fh.read(_Q8_HEADER)
    if head[:8] != _Q8_MAGIC:
        raise ValueError(f"{path}: not an LDPCQ8 file")
    version, width, rows, step, snr_db, seed = struct.unpack_from("<IIQffQ", head, 8)
    if version != 1:
        raise ValueError(f"{path}: version {version}")
    n = rows - offset if limit is None else limit
    if offset + n > rows:
        raise ValueError(f"{path}: {rows} rows < requested {offset + n}")
    data = np.memmap(path, dtype=np.int8, mode="r", offset=_Q8_HEADER + offset * width, shape=(n, width))
    return data, {"step": step, "snr_db": snr_db, "seed": seed, "rows": rows}


def uncor_text_to_q8(text_path: str, q8_path: str, step: float = 0.5, snr_db: float = float("nan"), dedup: bool = False) -> int:
    """Inputs/[Uncor]_*.txt -> sidecar.  The text rows are negated LLRs (Print_Functions.py:124); the sidecar holds
    decoder inputs.  dedup drops repeated words (keeps the first occurrence, order preserved)."""
    rows = read_uncor(text_path)
    words = llr_to_q8(-rows, step)
    if dedup:
        _, first = np.unique(words, axis=0, return_index=True)
        words = words[np.sort(first)]
    return write_uncor_q8(q8_path, words, step, snr_db)


def uncor_q8_to_text(q8_path: str, text_path: str) -> int:
    """Sidecar -> the reference's text format (appends, like write_uncor_file)."""
    words, meta = read_uncor_q8(q8_path)
    return append_uncor(text_path, q8_to_llr(words, meta["step"]))
