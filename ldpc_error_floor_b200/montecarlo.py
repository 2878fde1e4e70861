"""Monte-Carlo FER/BER driver: the host side of Print_Functions.compute_results (:130-165).

The reference loops `for batch: for SNR: create_mix_epoch -> sess.run -> calc_ber_fer -> write_uncor_file`
in Python, 20 frames per device call.  Here one call per chunk runs the fused generator + decoder +
counters + harvest kernel (ldpc_mc_run); the host only shards chunks over ranks and reduces counters.

Sharding: the global frame index space of an SNR point is cut into chunks of
`chunk_frames`; chunk c belongs to rank c % world_size.  The Philox counter is the GLOBAL frame index,
so FER/BER and the harvested words do not depend on the number of GPUs.  Collectives: one all-reduce
(sum) of the 8 uint64 counters per round, one all-gather of the harvested-word counts and padded
payloads at the end of a point -- NCCL over NVLink when the process group is NCCL, gloo in the CPU tests.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib

try:  # torch is plumbing only; the format / host-logic tests run without CUDA
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


@dataclass
class SnrPoint:
    """Counters of one Eb/N0 point (LDPC_CNT_* in include/ldpc_b200.h)."""
    snr_db: float
    sigma: float
    frames: int = 0
    frame_err_last: int = 0
    frame_err_any: int = 0
    bit_err_last: int = 0
    iters: int = 0
    synd_fail: int = 0
    undetected: int = 0
    harvested: int = 0
    bits_per_frame: int = 0
    seconds: float = 0.0
    chunks_done: int = 0           # chunks of the point's frame index space decoded so far (resume granularity)

    @property
    def fer(self) -> float:        # Results[2]: never correct at any iteration (Print_Functions.py:105-111)
        return self.frame_err_any / max(self.frames, 1)

    @property
    def fer_last(self) -> float:   # Results[1] (:115-116)
        return self.frame_err_last / max(self.frames, 1)

    @property
    def ber_last(self) -> float:   # Results[0] (:112-113)
        return self.bit_err_last / max(self.frames * self.bits_per_frame, 1)

    @property
    def avg_iters(self) -> float:
        return self.iters / max(self.frames, 1)

    def fer_ci95(self, which: str = "any"):
        """Wilson 95 % interval of the frame error rate."""
        k = self.frame_err_any if which == "any" else self.frame_err_last
        n = max(self.frames, 1)
        z = 1.959963984540054
        p = k / n
        den = 1 + z * z / n
        c = (p + z * z / (2 * n)) / den
        h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / den
        return max(0.0, c - h), min(1.0, c + h)

    def add(self, counters: Sequence[int]) -> None:
        for name, v in zip(_lib.COUNTER_NAMES, counters):
            setattr(self, name, getattr(self, name) + int(v))


def _world(group=None):
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _all_reduce_sum(vec: np.ndarray, device, group=None) -> np.ndarray:
    """Sum an int64 vector over ranks (tiny: latency-bound).  NCCL wants CUDA tensors, gloo CPU ones."""
    rank, world = _world(group)
    if world == 1:
        return vec
    backend = dist.get_backend(group)
    t = torch.from_numpy(vec.astype(np.int64))
    if backend == "nccl":
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def _all_gather_rows(rows: np.ndarray, device, group=None) -> np.ndarray:
    """Concatenate [n_r, W] float32 blocks of all ranks in rank order (counts first, padded payload second)."""
    rank, world = _world(group)
    if world == 1:
        return rows
    backend = dist.get_backend(group)
    dev = device if backend == "nccl" else "cpu"
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    if mx == 0:
        return rows[:0]
    pad = torch.zeros((mx, rows.shape[1]), dtype=torch.float32, device=dev)
    pad[:rows.shape[0]] = torch.from_numpy(rows).to(dev)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


class MonteCarlo:
    """Sharded Monte-Carlo over one decoder (anything with `.mc_run(...)`, `.graph`, `.device`).

    seed: Philox key shared by all ranks.  chunk_frames: frames per fused kernel launch (per rank)."""

    def __init__(self, decoder, seed: int = 2044, chunk_frames: int = 1 << 20, group=None):
        self.dec = decoder
        self.seed = int(seed)
        self.chunk = int(chunk_frames)
        self.group = group
        self.rank, self.world = _world(group)

    def run_point(self, snr_db: float, n_frames: int, *, sigma: Optional[float] = None, iters: int = 0,
                  early_term: bool = False, harvest: int = _lib.HARVEST_NONE, max_uncor: int = 0,
                  min_frame_errors: Optional[int] = None, frame_base: int = 0, round_chunks: int = 4,
                  stage1_iters: Optional[int] = None, start_chunk: int = 0, init_counters=None, init_rows=None,
                  on_checkpoint=None, checkpoint_rounds: int = 0):
        """Decode frames [frame_base, frame_base + n_frames) of this point's global index space.
        Stops early once `min_frame_errors` frame errors (any-iteration criterion) are seen, checked once
        per round of `round_chunks` chunks per rank.  Returns (SnrPoint, harvested LLR rows [n, N*z]).
        stage1_iters: with early termination, split each launch in two (NMSDecoder.mc_run): None = decide after the
        first chunk from its statistics (see _pick_stage1), 0 = never.  The counters do not depend on it.

        Resume / checkpoint (campaign.py): the point continues at chunk `start_chunk` with the counters (`init_counters`,
        int64[8]) and harvested rows (`init_rows`) of the chunks before it; every `checkpoint_rounds` rounds
        `on_checkpoint(chunks_done, counters, rows)` is called on every rank with the all-reduced state of chunks
        [0, chunks_done).  Chunks are keyed by global frame indices, so a point may be resumed with another number of ranks."""
        g = self.dec.graph
        if sigma is None:
            sigma = float(g.sigma([snr_db])[0])
        pt = SnrPoint(float(snr_db), float(sigma), bits_per_frame=g.NZ)
        t0 = time.time()
        n_chunks = (n_frames + self.chunk - 1) // self.chunk
        counters = ubuf = ucnt = None
        stage1 = stage1_iters if early_term and _takes_stage1(self.dec) else 0
        base = np.zeros(_lib.NUM_COUNTERS, dtype=np.int64) if init_counters is None else np.asarray(init_counters, dtype=np.int64)
        want_rows = harvest != _lib.HARVEST_NONE and max_uncor > 0
        acc_rows = [np.asarray(init_rows, dtype=np.float32).reshape(-1, g.NZ)] if (init_rows is not None and want_rows) else []
        sent = 0                                  # local harvested rows already gathered into acc_rows

        def collect_rows():
            nonlocal sent
            if not want_rows:
                return np.zeros((0, g.NZ), dtype=np.float32)
            n = min(int(self._to_numpy(ucnt)[0]), max_uncor) if ubuf is not None else 0
            new = self._rows(ubuf, n)[sent:] if n > sent else np.zeros((0, g.NZ), dtype=np.float32)
            sent = max(sent, n)
            acc_rows.append(_all_gather_rows(np.ascontiguousarray(new), self.dec.device, self.group))
            return np.concatenate(acc_rows, axis=0)[:max_uncor]          # the cap is global, not per rank

        done_chunks, rounds = int(start_chunk), 0
        while done_chunks < n_chunks:
            hi = min(n_chunks, done_chunks + round_chunks * self.world)
            for c in range(done_chunks, hi):
                if c % self.world != self.rank:
                    continue
                off = c * self.chunk
                n = min(self.chunk, n_frames - off)
                if stage1 is None and early_term and counters is not None:
                    stage1 = _pick_stage1(self._to_numpy(counters), self.dec.T if iters == 0 else iters)
                extra = {"stage1_iters": stage1} if stage1 else {}     # only decoders that know the two-stage form see it
                counters, ubuf, ucnt = self.dec.mc_run(
                    sigma, n, self.seed, frame_offset=frame_base + off, iters=iters, early_term=early_term,
                    harvest=harvest, capacity=max_uncor, counters=counters, uncor_buf=ubuf, uncor_count=ucnt, **extra)
            done_chunks = hi
            rounds += 1
            ckpt = on_checkpoint is not None and checkpoint_rounds > 0 and rounds % checkpoint_rounds == 0
            if (min_frame_errors is not None or ckpt) and done_chunks < n_chunks:
                tot = base + _all_reduce_sum(self._to_numpy(counters), self.dec.device, self.group)
                if ckpt:
                    on_checkpoint(done_chunks, tot, collect_rows())
                if min_frame_errors is not None and tot[_lib.COUNTER_NAMES.index("frame_err_any")] >= min_frame_errors:
                    break
        pt.add(base + _all_reduce_sum(self._to_numpy(counters), self.dec.device, self.group))
        rows = collect_rows()
        pt.seconds = time.time() - t0
        pt.chunks_done = done_chunks
        return pt, rows

    def sweep(self, snr_db_list: Sequence[float], n_frames: int, **kw) -> List[SnrPoint]:
        """One SnrPoint per Eb/N0 (the reference's `for SNR_idx` loop, Print_Functions.py:137).  Each point
        gets its own slice of the frame index space so points are statistically independent."""
        out = []
        for k, s in enumerate(snr_db_list):
            pt, _ = self.run_point(s, n_frames, frame_base=k * (1 << 40), **kw)
            out.append(pt)
        return out

    @staticmethod
    def _to_numpy(t) -> np.ndarray:
        if t is None:
            return np.zeros(_lib.NUM_COUNTERS, dtype=np.int64)
        if torch is not None and isinstance(t, torch.Tensor):
            return t.detach().cpu().numpy().astype(np.int64)
        return np.asarray(t, dtype=np.int64)

    @staticmethod
    def _rows(buf, n: int) -> np.ndarray:
        if torch is not None and isinstance(buf, torch.Tensor):
            return buf[:n].detach().cpu().numpy()
        return np.asarray(buf[:n], dtype=np.float32)


def _takes_stage1(dec) -> bool:
    """Does this decoder's mc_run know the two-stage form?  (Test doubles of NMSDecoder need not.)"""
    import inspect
    try:
        return "stage1_iters" in inspect.signature(dec.mc_run).parameters
    except (TypeError, ValueError):
        return False


def _pick_stage1(counters: np.ndarray, T: int) -> int:
    """Stage-1 iteration count of the two-stage Monte-Carlo launch from the statistics of the chunks decoded so far
    (0 = one launch).  It pays when a visible share f of the frames never reaches a zero syndrome -- each of them keeps the
    CTA it sits in busy for all T iterations -- while the others converge early, after a iterations on average: stage 1
    then runs about 1.3 a iterations and the stragglers are decoded again, densely packed.  Measured on B200
    (profiles/r01_two_stage_mc.txt): 5G n2112 z72 at 4.5-5.5 dB (f = 8-12 %, a = 4-5) +28..40 %, WiMAX at 3 dB (f = 6 %) +10 %;
    with f below 3 % (5G n1024 z64, WiMAX above 3.5 dB) or a above T / 2.6 there is nothing to gain, so it stays off."""
    names = _lib.COUNTER_NAMES
    frames = float(counters[names.index("frames")])
    if frames <= 0:
        return 0
    fail = float(counters[names.index("synd_fail")])
    f = fail / frames
    a = (float(counters[names.index("iters")]) - fail * T) / max(frames - fail, 1.0)
    s1 = int(min(max(3.0, math.ceil(1.3 * a)), T - 1))
    if f < 0.03 or f > 0.5 or s1 > T // 2:
        return 0
    return s1


def create_mix_epoch(decoder, scaling_factor, word_random, seed: int, batch_size: int, code_GM=None,
                     is_zeros_word: bool = True, frame_offset: int = 0):
    """Drop-in for Print_Functions.create_mix_epoch (:29-72): one batch, frame k drawn at scaling_factor[k % len] (:36).
    Returns (X CUDA float32 [B, N, z], Y int64 numpy [B, N*z]).  The noise comes from the device generator (Philox key
    `seed`, global frame indices from `frame_offset`); the codewords -- when `is_zeros_word` is False -- from the caller's
    numpy RandomState exactly as in the reference: infoWord = word_random.randint(0, 2, (1, k*z)), Y = infoWord . code_GM mod 2
    (:40-42), one draw per frame in frame order."""
    g = decoder.graph
    sf = np.atleast_1d(np.asarray(scaling_factor, dtype=np.float64))
    B, n = int(batch_size), sf.size
    Y = np.zeros((B, g.NZ), dtype=np.int64)
    if not is_zeros_word:
        GM = np.asarray(code_GM)
        for k in range(B):
            info = word_random.randint(0, 2, size=(1, GM.shape[0]))
            Y[k] = (np.dot(info, GM) % 2)[0]
    else:
        for k in range(B):
            word_random.randint(0, 2, size=(1, g.NZ))                  # the reference draws and discards (:39)
    X = torch.empty((B, g.N, g.z), dtype=torch.float32, device=decoder.device)
    for s, sg in enumerate(sf):
        idx = np.arange(s, B, n)
        if idx.size:
            X[s::n] = decoder.generate(float(sg), idx.size, seed, frame_offset + s * B,
                                       codeword=None if is_zeros_word else Y[idx])
    return X, Y


def compute_results(decoder, sample_num, input_llr, SNR_sigma, batch_size, sampling_type, seed=2044,
                    uncor_path: Optional[str] = None, iters: int = 0, group=None, input_codeword=None):
    """Drop-in for Print_Functions.compute_results (:130-165): returns (Results f32[4, nSNR], seconds) with
    rows BER_last, FER_last, FER, loss.  The loss row stays 0: the fused Monte-Carlo path has no loss kernel, so
    `opt_result_print = 3` (best epoch by validation loss) is refused by drivers.evaluate / trainer.train_block.

    sampling_type 0/2: `floor(sample_num/batch_size)*batch_size` generated frames per sigma (:135-143);
    2 also appends the never-corrected words to `uncor_path` in the Inputs/[Uncor] format (:155-156).
    sampling_type 1: decodes the stored rows `input_llr` (file sign convention, :145 / :6-10); with `input_codeword`
    (labels [n, N*z], Main_Functions.process_data) the metrics are taken against those words (ldpc_decode_cw), BER with the
    reference's signed sum (:112-113)."""
    from . import formats
    t0 = time.time()
    g = decoder.graph
    SNR_sigma = np.atleast_1d(np.asarray(SNR_sigma, dtype=np.float64))
    res = np.zeros((4, SNR_sigma.size), dtype=np.float32)
    n = int(math.floor(sample_num / batch_size)) * int(batch_size)
    if sampling_type == 1:
        xa = formats.uncor_to_llr(np.asarray(input_llr, dtype=np.float32)[:n], g.N, g.z)
        # stored words of a quantised decoder sit on its grid ('%.1f' of 0.5-steps, Print_Functions.py:124): ship them as
        # int8 -- a quarter of the bytes over PCIe, same results (tests/test_gpu_parity.py::test_q8_words_...)
        if input_codeword is not None and np.any(np.asarray(input_codeword)[:n]):
            Y = np.asarray(input_codeword)[:n].reshape(n, -1)
            r, signed, cnt = decoder.decode_cw(torch.from_numpy(xa).to(decoder.device), Y, iters=iters)
            c = cnt.cpu().numpy()
            nb = max(int(batch_size), 1)
            sg = signed.cpu().numpy().astype(np.int64)[:(n // nb) * nb].reshape(-1, nb)
            res[0, :] = np.abs(sg.sum(axis=1)).sum() / max(n * g.NZ, 1)          # |sum of signed errors| per batch (:112-113, 158)
            res[1, :] = c[1] / max(n, 1)
            res[2, :] = c[2] / max(n, 1)
            return res, time.time() - t0
        step = float(getattr(decoder, "q8_step", 0.0) or 0.0)
        words = None
        if step > 0 and hasattr(decoder, "decode_q8_host"):
            try:
                words = formats.llr_to_q8(xa.reshape(xa.shape[0], -1), step)
            except ValueError:
                words = None                      # off-grid rows (a float decoder's file): float32 transport
        r = decoder.decode_q8_host(words, iters=iters) if words is not None else decoder.decode_host(xa, iters=iters)
        flags = r["flags"]
        res[0, :] = r["biterr"].sum() / max(n * g.NZ, 1)
        res[1, :] = ((flags & _lib.FLAG_UNCOR_LAST) != 0).mean() if n else 0.0
        res[2, :] = ((flags & _lib.FLAG_UNCOR_ANY) != 0).mean() if n else 0.0
        return res, time.time() - t0
    mc = MonteCarlo(decoder, seed=seed, group=group)
    harvest = _lib.HARVEST_UNCOR_ANY if sampling_type == 2 else _lib.HARVEST_NONE
    for k, sg in enumerate(SNR_sigma):
        pt, rows = mc.run_point(0.0, n, sigma=float(sg), iters=iters, harvest=harvest,
                                max_uncor=n if sampling_type == 2 else 0, frame_base=k * (1 << 40))
        res[0, k], res[1, k], res[2, k] = pt.ber_last, pt.fer_last, pt.fer
        if sampling_type == 2 and rows.shape[0] and uncor_path and mc.rank == 0:
            formats.append_uncor(uncor_path, rows)
    return res, time.time() - t0
