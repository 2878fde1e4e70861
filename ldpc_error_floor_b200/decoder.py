"""Host side of the NMS decode op: decoder handle + PyTorch custom op over the C-ABI.

`NMSDecoder.decode` is what a shimmed `sess.run(net_dict["ya_output_all"], feed_dict={xa: ...})`
(Print_Functions.py:148-151) maps onto; `NMSDecoder.ya_output_all` returns exactly that
tensor ([T*B, N*z] float32, iterations concatenated along axis 0, Main_Functions.py:380-383).
PyTorch is only plumbing here (device memory, streams); all arithmetic is in csrc/.
"""

import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .formats import WeightSet, weight_width
from .graph import BaseGraph


def check_params(sampling_type, SNR_Matrix, sharing, iters_max, fixed_iter, iter_step):
    """Main_Functions.check_params (:498-523) with the same rules; raises ValueError where the
    reference prints and sys.exit()s."""
    SNR_Matrix = np.asarray(SNR_Matrix, dtype=np.float64)
    if sampling_type == 1:
        if len(SNR_Matrix) > 1:
            SNR_Matrix = np.array([0.0])
    elif sampling_type == 2:
        if len(SNR_Matrix) > 1:
            raise ValueError("sampling_type == 2 and len(SNR_Matrix) > 1")
    if np.sum(sharing) == 0:
        raise ValueError("np.sum(sharing) == 0")
    if any(v in [4, 5] for v in sharing) and (iters_max - fixed_iter) % iter_step > 0:
        raise ValueError("any(value in [4,5] for value in sharing) and (iters_max - fixed_iter) % iter_step > 0")
    if sharing[2] in [1, 4]:
        raise ValueError("sharing[2] in [1,4]")
    if sharing[1] != 0 and sharing[0] != sharing[1]:
        raise ValueError("sharing[1] != 0 and sharing[0]!=sharing[1])")
    return SNR_Matrix


@dataclass
class DecodeResult:
    hard: Optional[torch.Tensor]      # uint8 [B, N*z] hard decision (bit = APP >= 0) -- unpacked view
    hard_packed: Optional[torch.Tensor]   # int32 [B, ceil(N*z/32)] as the kernel wrote it
    iters: torch.Tensor               # int32 [B] iterations until the first zero syndrome (else T)
    flags: torch.Tensor               # uint8 [B] LDPC_FLAG_* bits
    biterr: torch.Tensor              # int32 [B] ones in the output decision
    app: Optional[torch.Tensor]       # f32 [B, N*z] or [T, B, N*z]

    @property
    def synd_ok(self):
        return (self.flags & _lib.FLAG_SYND_OK) != 0

    @property
    def uncor_any(self):
        return (self.flags & _lib.FLAG_UNCOR_ANY) != 0

    @property
    def uncor_last(self):
        return (self.flags & _lib.FLAG_UNCOR_LAST) != 0


def pack_codewords(y, nbits: int, device) -> "tuple[torch.Tensor, int]":
    """Codeword bits [nbits] (shared by all frames) or [B, nbits] (0/1, any integer / bool type, numpy or torch) ->
    (int32 CUDA tensor of packed words, stride in words): bit k at word k // 32, bit k % 32 (include/ldpc_b200.h)."""
    a = y.detach().cpu().numpy() if isinstance(y, torch.Tensor) else np.asarray(y)
    shared = a.ndim == 1
    a = np.ascontiguousarray(a.reshape(1 if shared else a.shape[0], -1) != 0)
    if a.shape[1] != nbits:
        raise ValueError(f"codeword has {a.shape[1]} bits, expected {nbits}")
    words = (nbits + 31) // 32
    pad = np.zeros((a.shape[0], words * 32), dtype=bool)
    pad[:, :nbits] = a
    packed = np.packbits(pad, axis=1, bitorder="little").view(np.uint32).reshape(a.shape[0], words)
    return torch.from_numpy(packed.view(np.int32).copy()).to(device), (0 if shared else words)


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def unpack_bits(packed: torch.Tensor, nbits: int) -> torch.Tensor:
    """int32 [B, W] -> uint8 [B, nbits]; bit k lives at word k//32, bit k%32."""
    shifts = torch.arange(32, device=packed.device, dtype=torch.int32)
    bits = (packed.unsqueeze(-1) >> shifts) & 1
    return bits.reshape(packed.shape[0], -1)[:, :nbits].to(torch.uint8)


class NMSDecoder:
    """Flooding neural min-sum decoder for one (graph, weight set, arithmetic mode) on one GPU.

    weights: WeightSet (sharing codes + [T, width] blocks, format F2); `iters` defaults to the
    number of weight rows.  decoding_type 0 = sum-product, 1 = min-sum, 2 = quantised min-sum (q_bit as in
    Main_Functions.py:483-492).  A "post decoder" is simply an NMSDecoder built from the boosted
    weight file (base rows followed by post rows)."""

    def __init__(self, graph: BaseGraph, weights: WeightSet, iters: Optional[int] = None, decoding_type: int = 2,
                 q_bit: int = 5, clip_llr: float = 20.0, device: Optional[int] = None, systematic: int = 0,
                 fixed_iter: int = 0):
        """systematic = 1 (main_Base.py:29): the error metrics cover the first N - M proto columns only.
        fixed_iter: for temporal sharing (code 4) given as its fixed_iter + 1 variables."""
        self.graph = graph
        T = weights.iterations if iters is None else int(iters)
        if any(int(c) in (4, 5) for c in weights.sharing):
            from .formats import expand_temporal
            if iters is None:
                raise ValueError("temporal sharing needs the iteration count (iters=...)")
            weights = expand_temporal(weights, T, int(fixed_iter))
        self.target_node = graph.N - graph.M if systematic else 0
        self.sharing = [int(s) for s in weights.sharing]
        if T <= 0:
            raise ValueError("a decoder needs at least one iteration (all-zero sharing carries no rows: pass iters)")
        self.T = T
        self.decoding_type, self.q_bit, self.clip_llr = int(decoding_type), int(q_bit), float(clip_llr)
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        blocks = [None, None, None]
        for i, code in enumerate(self.sharing):
            if code > 0:
                w = np.ascontiguousarray(np.asarray(weights.blocks[i], dtype=np.float32))
                width = weight_width(code, i, graph.M, graph.N, graph.E)
                if w.ndim == 1:
                    w = w.reshape(-1, 1)
                if w.shape[0] < T or w.shape[1] != width:
                    raise ValueError(f"weight block {i}: shape {w.shape}, need [>={T}, {width}]")
                blocks[i] = np.ascontiguousarray(w[:T])
        self._blocks = blocks
        lib = _lib.load()
        sh = (ctypes.c_int32 * 3)(*self.sharing)
        self._h = ctypes.c_void_p()
        _lib.check(lib.ldpc_decoder_create2(
            graph._h, sh, T, *[b.ctypes.data if b is not None else None for b in blocks],
            self.decoding_type, self.q_bit, self.clip_llr, self.device_index, self.target_node, ctypes.byref(self._h)))
        self.packed = bool(lib.ldpc_decoder_uses_packed_kernel(self._h))
        self.kernel_name = lib.ldpc_decoder_kernel_name(self._h).decode()
        fb, cps, thr, smem = (ctypes.c_int32() for _ in range(4))
        _lib.check(lib.ldpc_decoder_geometry(self._h, ctypes.byref(fb), ctypes.byref(cps), ctypes.byref(thr),
                                             ctypes.byref(smem)))
        self.frames_per_cta, self.ctas_per_sm = fb.value, cps.value
        self.threads_per_cta, self.smem_bytes = thr.value, smem.value
        self.hard_words = (graph.NZ + 31) // 32

    def launch_info(self, early_term: bool = False) -> Dict:
        """Kernel and launch geometry a call with / without early termination uses (ldpc_decoder_launch_info)."""
        fb, cps, thr, smem = (ctypes.c_int32() for _ in range(4))
        name = ctypes.create_string_buffer(96)
        _lib.check(_lib.load().ldpc_decoder_launch_info(self._h, 1 if early_term else 0, ctypes.byref(fb), ctypes.byref(cps),
                                                        ctypes.byref(thr), ctypes.byref(smem), name, 96))
        return {"kernel": name.value.decode(), "frames_per_cta": fb.value, "ctas_per_sm": cps.value,
                "threads_per_cta": thr.value, "smem_bytes": smem.value}

    def mc_info(self) -> Dict:
        """Kernel and launch geometry of mc_run(early_term=True) (ldpc_decoder_mc_info); 'persistent' = the
        persistent-slot kernel serves it."""
        pers, fb, cps, thr, smem = (ctypes.c_int32() for _ in range(5))
        name = ctypes.create_string_buffer(96)
        _lib.check(_lib.load().ldpc_decoder_mc_info(self._h, ctypes.byref(pers), ctypes.byref(fb), ctypes.byref(cps),
                                                    ctypes.byref(thr), ctypes.byref(smem), name, 96))
        return {"persistent": bool(pers.value), "kernel": name.value.decode(), "frames_per_cta": fb.value,
                "ctas_per_sm": cps.value, "threads_per_cta": thr.value, "smem_bytes": smem.value}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().ldpc_decoder_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ decode (device buffers)
    def decode(self, llr: torch.Tensor, iters: int = 0, early_term: bool = False, app: Optional[str] = None,
               want_hard: bool = True, unpack: bool = False) -> DecodeResult:
        """llr: CUDA float32 [B, N, z] or [B, N*z] (log p1/p0).  app: None | 'last' | 'all'."""
        if app not in (None, "last", "all"):
            raise ValueError("app must be None, 'last' or 'all'")
        return _decode_via_op(self, llr, iters, early_term, app, want_hard, unpack)

    def _decode_impl(self, llr, iters, early_term, app, want_hard, unpack) -> DecodeResult:
        g = self.graph
        if not llr.is_cuda or llr.dtype != torch.float32:
            raise ValueError("decode: llr must be a CUDA float32 tensor (use decode_host for host buffers)")
        if llr.device.index != self.device_index:
            raise ValueError(f"decode: llr lives on {llr.device}, decoder on {self.device}")
        B = llr.shape[0]
        if llr.numel() != B * g.NZ:
            raise ValueError(f"decode: llr has {llr.numel()} elements, expected {B}x{g.NZ}")
        llr = llr.contiguous()
        T_run = self.T if iters == 0 else iters
        dev = self.device
        hard = torch.empty((B, self.hard_words), dtype=torch.int32, device=dev) if want_hard else None
        it = torch.empty((B,), dtype=torch.int32, device=dev)
        fl = torch.empty((B,), dtype=torch.uint8, device=dev)
        be = torch.empty((B,), dtype=torch.int32, device=dev)
        app_t = None
        if app == "last":
            app_t = torch.empty((B, g.NZ), dtype=torch.float32, device=dev)
        elif app == "all":
            app_t = torch.zeros((T_run, B, g.NZ), dtype=torch.float32, device=dev)
        elif app is not None:
            raise ValueError("app must be None, 'last' or 'all'")
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().ldpc_decode(self._h, _ptr(llr), B, int(iters), 1 if early_term else 0, _ptr(app_t),
                                           1 if app == "all" else 0, _ptr(hard), _ptr(it), _ptr(fl), _ptr(be),
                                           ctypes.c_void_p(stream)))
        return DecodeResult(unpack_bits(hard, g.NZ) if (unpack and hard is not None) else None, hard, it, fl, be, app_t)

    def decode_cw(self, llr: torch.Tensor, codeword, iters: int = 0, early_term: bool = False,
                  counters: Optional[torch.Tensor] = None):
        """decode() with the error metrics taken against `codeword` (bits [N*z] or [B, N*z]) instead of the all-zero word
        (ldpc_decode_cw; calc_ber_fer with Y != 0, Print_Functions.py:100-118).  Returns (DecodeResult, biterr_signed int32 [B],
        counters int64[8]); DecodeResult.biterr is the Hamming distance of the output decision."""
        g = self.graph
        if not llr.is_cuda or llr.dtype != torch.float32:
            raise ValueError("decode_cw: llr must be a CUDA float32 tensor")
        B = llr.shape[0]
        if llr.numel() != B * g.NZ:
            raise ValueError(f"decode_cw: llr has {llr.numel()} elements, expected {B}x{g.NZ}")
        llr = llr.contiguous()
        cw, stride = pack_codewords(codeword, g.NZ, self.device)
        if stride and cw.shape[0] != B:
            raise ValueError("decode_cw: one codeword per frame, or one for all")
        dev = self.device
        hard = torch.empty((B, self.hard_words), dtype=torch.int32, device=dev)
        it = torch.empty((B,), dtype=torch.int32, device=dev)
        fl = torch.empty((B,), dtype=torch.uint8, device=dev)
        be = torch.empty((B,), dtype=torch.int32, device=dev)
        bs = torch.empty((B,), dtype=torch.int32, device=dev)
        if counters is None:
            counters = torch.zeros((_lib.NUM_COUNTERS,), dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().ldpc_decode_cw(self._h, _ptr(llr), _ptr(cw), stride, B, int(iters), 1 if early_term else 0,
                                              _ptr(hard), _ptr(it), _ptr(fl), _ptr(be), _ptr(bs), _ptr(counters),
                                              ctypes.c_void_p(stream)))
        torch.cuda.current_stream(dev).synchronize()                  # cw is a temporary
        return DecodeResult(None, hard, it, fl, be, None), bs, counters

    def ya_output_all(self, xa: torch.Tensor, iters: int = 0) -> torch.Tensor:
        """net_dict["ya_output_all"] of the reference: [T*B, N*z] float32 (Main_Functions.py:380-383)."""
        r = self.decode(xa, iters=iters, early_term=False, app="all", want_hard=False)
        return r.app.reshape(-1, self.graph.NZ)

    # -------------------------------------------------------------------- decode (host buffers)
    def _host_results(self, B: int, out: Optional[Dict]):
        """Result arrays of a host-buffer call: fresh ones, or those of `out` (a dict an earlier call returned) when their
        shapes fit -- a steady-state loop then writes into memory that is already mapped instead of faulting in ~80 bytes
        per frame of new pages per call."""
        shapes = {"hard_packed": ((B, self.hard_words), np.uint32), "iters": ((B,), np.int32), "flags": ((B,), np.uint8),
                  "biterr": ((B,), np.int32)}
        res = []
        for key, (shape, dt) in shapes.items():
            a = out.get(key) if out is not None else None
            if not (isinstance(a, np.ndarray) and a.shape == shape and a.dtype == dt and a.flags.c_contiguous and a.flags.writeable):
                a = np.empty(shape, dtype=dt)
            res.append(a)
        return res

    def decode_host(self, llr, iters: int = 0, early_term: bool = False, app: Optional[str] = None,
                    out: Optional[Dict] = None) -> Dict:
        """End-to-end call: llr is a host array (numpy or CPU torch tensor, pinned or not); results
        come back in host numpy arrays (those of `out`, the dict of an earlier call, when given: they are overwritten).
        Copies and kernels are pipelined inside the library."""
        g = self.graph
        if isinstance(llr, torch.Tensor):
            if llr.is_cuda:
                raise ValueError("decode_host takes host memory")
            arr = llr.contiguous().numpy()
        else:
            arr = np.ascontiguousarray(llr, dtype=np.float32)
        if arr.dtype != np.float32:
            arr = arr.astype(np.float32)
        B = arr.shape[0]
        if arr.size != B * g.NZ:
            raise ValueError(f"decode_host: llr has {arr.size} elements, expected {B}x{g.NZ}")
        T_run = self.T if iters == 0 else iters
        hard, it, fl, be = self._host_results(B, out)
        app_a = None
        if app == "last":
            app_a = np.empty((B, g.NZ), dtype=np.float32)
        elif app == "all":
            app_a = np.zeros((T_run, B, g.NZ), dtype=np.float32)
        _lib.check(_lib.load().ldpc_decode_host(
            self._h, arr.ctypes.data, B, int(iters), 1 if early_term else 0,
            app_a.ctypes.data if app_a is not None else None, 1 if app == "all" else 0, hard.ctypes.data,
            it.ctypes.data, fl.ctypes.data, be.ctypes.data))
        return {"hard_packed": hard, "iters": it, "flags": fl, "biterr": be, "app": app_a}

    # ------------------------------------------------------------------ training step (N1)
    def set_weights(self, weights: WeightSet) -> None:
        """Replace the weights (same sharing codes and shapes): the variable update of a training step."""
        if [int(c) for c in weights.sharing] != self.sharing:
            raise ValueError(f"set_weights: sharing {weights.sharing} != {self.sharing}")
        blocks = [None, None, None]
        for i, code in enumerate(self.sharing):
            if code > 0:
                w = np.ascontiguousarray(np.asarray(weights.blocks[i], dtype=np.float32).reshape(-1, self._blocks[i].shape[1])[:self.T])
                if w.shape != self._blocks[i].shape:
                    raise ValueError(f"set_weights: block {i} has shape {w.shape}, need {self._blocks[i].shape}")
                blocks[i] = w
        _lib.check(_lib.load().ldpc_decoder_set_weights(self._h, *[b.ctypes.data if b is not None else None for b in blocks]))
        self._blocks = blocks

    def train_grad(self, llr: torch.Tensor, iter_lo: int = 0, loss_type: int = 2, etha: float = 0.0, iters: int = 0,
                   want_app: bool = False):
        """One training batch of the reference (main_Base.py:160-162): returns (loss, {i: grad f32 [T, width]}, app).
        llr: CUDA float32 [B, N, z] / [B, N*z]; loss over iterations [iter_lo, iters), gradient for their weights."""
        g = self.graph
        if not llr.is_cuda or llr.dtype != torch.float32:
            raise ValueError("train_grad: llr must be a CUDA float32 tensor")
        B = llr.shape[0]
        if llr.numel() != B * g.NZ:
            raise ValueError(f"train_grad: llr has {llr.numel()} elements, expected {B}x{g.NZ}")
        llr = llr.contiguous()
        T_run = self.T if iters == 0 else int(iters)
        grads = [np.zeros_like(b) if b is not None else None for b in self._blocks]
        app = torch.empty((T_run, B, g.NZ), dtype=torch.float32, device=self.device) if want_app else None
        loss = ctypes.c_double(0.0)
        torch.cuda.current_stream(self.device).synchronize()
        _lib.check(_lib.load().ldpc_train_grad(self._h, _ptr(llr), B, int(iters), int(iter_lo), int(loss_type), float(etha),
                                               ctypes.byref(loss), *[x.ctypes.data if x is not None else None for x in grads],
                                               _ptr(app)))
        return loss.value, {i: x for i, x in enumerate(grads) if x is not None}, app

    # -------------------------------------------------------------- compact int8 words (N2)
    @property
    def q8_step(self) -> float:
        """LLR units of one int8 count: the quantiser step of this decoder (0.0 for float decoders)."""
        return float(_lib.load().ldpc_decoder_q8_step(self._h))

    def decode_q8(self, words: torch.Tensor, iters: int = 0, early_term: bool = False, step: float = 0.0,
                  counters: Optional[torch.Tensor] = None) -> DecodeResult:
        """words: CUDA int8 [B, N*z], llr = words * step (step 0 = the quantiser step).  Same outputs as
        decode() without APP; `counters` (int64[8], accumulated) as in mc_run."""
        g = self.graph
        if not words.is_cuda or words.dtype != torch.int8:
            raise ValueError("decode_q8: words must be a CUDA int8 tensor")
        B = words.shape[0]
        if words.numel() != B * g.NZ:
            raise ValueError(f"decode_q8: {words.numel()} elements, expected {B}x{g.NZ}")
        words = words.contiguous()
        dev = self.device
        hard = torch.empty((B, self.hard_words), dtype=torch.int32, device=dev)
        it = torch.empty((B,), dtype=torch.int32, device=dev)
        fl = torch.empty((B,), dtype=torch.uint8, device=dev)
        be = torch.empty((B,), dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().ldpc_decode_q8(self._h, _ptr(words), float(step), B, int(iters), 1 if early_term else 0,
                                              _ptr(hard), _ptr(it), _ptr(fl), _ptr(be), _ptr(counters),
                                              ctypes.c_void_p(stream)))
        return DecodeResult(None, hard, it, fl, be, None)

    def decode_q8_host(self, words, iters: int = 0, early_term: bool = False, step: float = 0.0,
                       out: Optional[Dict] = None) -> Dict:
        """End-to-end call on host int8 words (numpy or CPU tensor, pinned or not); results in numpy arrays (`out`: see
        decode_host)."""
        g = self.graph
        arr = words.contiguous().numpy() if isinstance(words, torch.Tensor) else np.ascontiguousarray(words)
        if arr.dtype != np.int8:
            raise ValueError("decode_q8_host takes int8 words")
        B = arr.shape[0]
        if arr.size != B * g.NZ:
            raise ValueError(f"decode_q8_host: {arr.size} elements, expected {B}x{g.NZ}")
        hard, it, fl, be = self._host_results(B, out)
        _lib.check(_lib.load().ldpc_decode_q8_host(self._h, arr.ctypes.data, float(step), B, int(iters),
                                                   1 if early_term else 0, hard.ctypes.data, it.ctypes.data,
                                                   fl.ctypes.data, be.ctypes.data))
        return {"hard_packed": hard, "iters": it, "flags": fl, "biterr": be, "app": None}

    def host_stats(self) -> Dict:
        """What the last decode_host / decode_q8_host call did (ldpc_decode_host_stats): chunks sent as int8 / float32,
        host threads, seconds the calling thread packed / waited for the device, bytes copied."""
        st = _lib.HostStats()
        _lib.check(_lib.load().ldpc_decode_host_stats(self._h, ctypes.byref(st)))
        return st.as_dict()

    def pack_q8(self, llr):
        """float32 host words -> (int8 words in units of the quantiser step, number of values with no int8 form):
        the packing pass of decode_host on its own (ldpc_pack_q8_host), for callers that decode a word set repeatedly
        with decode_q8_host."""
        arr = llr.contiguous().numpy() if isinstance(llr, torch.Tensor) else np.ascontiguousarray(llr, dtype=np.float32)
        B = arr.shape[0]
        if arr.dtype != np.float32 or arr.size != B * self.graph.NZ:
            raise ValueError(f"pack_q8: float32 [B, {self.graph.NZ}] expected")
        out = np.empty((B, self.graph.NZ), dtype=np.int8)
        bad = ctypes.c_int64(0)
        _lib.check(_lib.load().ldpc_pack_q8_host(self._h, arr.ctypes.data, B, out.ctypes.data, ctypes.byref(bad)))
        return out, int(bad.value)

    # --------------------------------------------------------------------- generator / MC / post
    def generate(self, sigma: float, n_frames: int, seed: int, frame_offset: int = 0, codeword=None) -> torch.Tensor:
        """BPSK/AWGN LLRs of the all-zero codeword, CUDA float32 [n_frames, N, z]
        (replaces Print_Functions.create_mix_epoch, :29-72; Philox instead of MT19937).
        codeword: bits [N*z] (every frame) or [n_frames, N*z]: the samples of those words instead (is_zeros_word = False, :40-46)."""
        out = torch.empty((n_frames, self.graph.N, self.graph.z), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if codeword is not None:
            cw, stride = pack_codewords(codeword, self.graph.NZ, self.device)
            if stride and cw.shape[0] != n_frames:
                raise ValueError("generate: one codeword per frame, or one for all")
            _lib.check(_lib.load().ldpc_llr_generate_cw(self._h, float(sigma), int(n_frames), int(seed) & (2**64 - 1),
                                                        int(frame_offset), _ptr(cw), stride, _ptr(out), ctypes.c_void_p(stream)))
            torch.cuda.current_stream(self.device).synchronize()      # cw is a temporary
            return out
        _lib.check(_lib.load().ldpc_llr_generate(self._h, float(sigma), int(n_frames), int(seed) & (2**64 - 1),
                                                 int(frame_offset), _ptr(out), ctypes.c_void_p(stream)))
        return out

    def mc_run(self, sigma: float, n_frames: int, seed: int, frame_offset: int = 0, iters: int = 0,
               early_term: bool = False, harvest: int = _lib.HARVEST_NONE, capacity: int = 0,
               counters: Optional[torch.Tensor] = None, uncor_buf: Optional[torch.Tensor] = None,
               uncor_count: Optional[torch.Tensor] = None, stage1_iters: int = 0):
        """Fused generate + decode + count (+ harvest) for one SNR point, asynchronous on the current
        stream.  Returns (counters int64[8] CUDA, uncor_buf f32 [capacity, N*z] CUDA or None,
        uncor_count int32[1] CUDA); pass the tensors back in to keep accumulating.
        stage1_iters > 0 (with early_term): the two-stage form (ldpc_mc_run_staged) -- same counters, the frames that
        have not converged after stage1_iters iterations are decoded in a dense second launch."""
        dev = self.device
        if counters is None:
            counters = torch.zeros((_lib.NUM_COUNTERS,), dtype=torch.int64, device=dev)
        if uncor_count is None:
            uncor_count = torch.zeros((1,), dtype=torch.int32, device=dev)
        if harvest != _lib.HARVEST_NONE and capacity > 0 and uncor_buf is None:
            uncor_buf = torch.empty((capacity, self.graph.NZ), dtype=torch.float32, device=dev)
        cap = 0 if uncor_buf is None else uncor_buf.shape[0]
        stream = torch.cuda.current_stream(dev).cuda_stream
        T = self.T if iters == 0 else int(iters)
        if early_term and 0 < stage1_iters < T:
            if getattr(self, "_defer_list", None) is None or self._defer_list.numel() < n_frames:
                self._defer_list = torch.empty((int(n_frames),), dtype=torch.int64, device=dev)
                self._defer_count = torch.zeros((1,), dtype=torch.int32, device=dev)
            _lib.check(_lib.load().ldpc_mc_run_staged(self._h, float(sigma), int(n_frames), int(seed) & (2**64 - 1),
                                                      int(frame_offset), int(iters), int(stage1_iters), int(harvest),
                                                      _ptr(counters), _ptr(uncor_buf), _ptr(uncor_count), cap,
                                                      _ptr(self._defer_list), _ptr(self._defer_count),
                                                      int(self._defer_list.numel()), ctypes.c_void_p(stream)))
            return counters, uncor_buf, uncor_count
        _lib.check(_lib.load().ldpc_mc_run(self._h, float(sigma), int(n_frames), int(seed) & (2**64 - 1),
                                           int(frame_offset), int(iters), 1 if early_term else 0, int(harvest),
                                           _ptr(counters), _ptr(uncor_buf), _ptr(uncor_count), cap,
                                           ctypes.c_void_p(stream)))
        return counters, uncor_buf, uncor_count

    def mc_run_host(self, sigma: float, n_frames: int, seed: int, frame_offset: int = 0, iters: int = 0,
                    early_term: bool = False, harvest: int = _lib.HARVEST_NONE, capacity: int = 0):
        """Synchronous host-buffer twin: returns (dict of counters, numpy [n_uncor, N*z])."""
        cnt = np.zeros(_lib.NUM_COUNTERS, dtype=np.uint64)
        rows = np.empty((max(capacity, 1), self.graph.NZ), dtype=np.float32)
        n = ctypes.c_uint32(0)
        _lib.check(_lib.load().ldpc_mc_run_host(self._h, float(sigma), int(n_frames), int(seed) & (2**64 - 1),
                                                int(frame_offset), int(iters), 1 if early_term else 0, int(harvest),
                                                cnt.ctypes.data, rows.ctypes.data if capacity > 0 else None,
                                                int(capacity), ctypes.byref(n)))
        return dict(zip(_lib.COUNTER_NAMES, (int(v) for v in cnt))), rows[:n.value]

    def post_decode(self, words: torch.Tensor, iters: int = 0, early_term: bool = False,
                    counters: Optional[torch.Tensor] = None):
        """Run this (boosted) decoder on compacted uncorrected words: CUDA float32 [n, N*z] decoder-input
        LLRs, e.g. the buffer mc_run filled (main_Post.py on Inputs/[Uncor]_*)."""
        dev = self.device
        n = words.shape[0]
        if counters is None:
            counters = torch.zeros((_lib.NUM_COUNTERS,), dtype=torch.int64, device=dev)
        hard = torch.empty((n, self.hard_words), dtype=torch.int32, device=dev)
        it = torch.empty((n,), dtype=torch.int32, device=dev)
        fl = torch.empty((n,), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().ldpc_post_decode(self._h, _ptr(words.contiguous()), n, int(iters),
                                                1 if early_term else 0, _ptr(counters), _ptr(hard), _ptr(it),
                                                _ptr(fl), ctypes.c_void_p(stream)))
        return counters, DecodeResult(None, hard, it, fl, torch.empty(0), None)


def normal_probe(seed: int, n_frames: int, quads_per_frame: int, frame_offset: int = 0, want_normals: bool = False,
                 device: int = 0, counts: Optional[torch.Tensor] = None):
    """The generator's N(0,1) stream on its own (ldpc_normal_probe): returns (counts int64[6] CUDA: samples with
    |n| > 3, 4, 5, 6, 7 sigma and the total, accumulated into `counts` when given; normals f32 CUDA or None)."""
    dev = torch.device("cuda", device)
    if counts is None:
        counts = torch.zeros((6,), dtype=torch.int64, device=dev)
    out = torch.empty((n_frames * quads_per_frame * 4,), dtype=torch.float32, device=dev) if want_normals else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.load().ldpc_normal_probe(int(device), int(seed) & (2**64 - 1), int(frame_offset), int(n_frames),
                                             int(quads_per_frame), _ptr(out), _ptr(counts), ctypes.c_void_p(stream)))
    return counts, out


# ---- PyTorch custom op: torch.ops.ldpc_b200.nms_decode(llr, handle, ...) -> (hard, iters, flags, biterr, app)
# `handle` is a small integer the decoder owns for its whole life (weak registry below), so a traced / exported graph that
# captured the op keeps referring to a live decoder; it is not the C pointer and not id().
import itertools
import weakref

_REGISTRY: "weakref.WeakValueDictionary[int, NMSDecoder]" = weakref.WeakValueDictionary()
_NEXT_HANDLE = itertools.count(1)


def op_handle(dec: NMSDecoder) -> int:
    """The integer that names `dec` in torch.ops.ldpc_b200.nms_decode calls (assigned once, valid while `dec` lives)."""
    h = getattr(dec, "_op_handle", None)
    if h is None:
        h = next(_NEXT_HANDLE)
        dec._op_handle = h
        _REGISTRY[h] = dec
    return h


def _lookup(handle: int) -> NMSDecoder:
    dec = _REGISTRY.get(int(handle))
    if dec is None:
        raise RuntimeError(f"ldpc_b200::nms_decode: decoder handle {handle} is not alive")
    return dec


@torch.library.custom_op("ldpc_b200::nms_decode", mutates_args=(), device_types="cuda")
def _nms_decode_op(llr: torch.Tensor, handle: int, iters: int, early_term: bool, app_mode: int) -> List[torch.Tensor]:
    dec = _lookup(handle)
    app = {0: None, 1: "last", 2: "all"}[app_mode]
    r = dec._decode_impl(llr, iters, early_term, app, True, False)
    app_t = r.app if r.app is not None else torch.empty(0, device=llr.device)
    return [r.hard_packed, r.iters, r.flags, r.biterr, app_t]


@_nms_decode_op.register_fake
def _(llr, handle, iters, early_term, app_mode):
    dec = _lookup(handle)
    B = llr.shape[0]
    T_run = dec.T if iters == 0 else iters
    app = (torch.empty(0, device=llr.device) if app_mode == 0 else
           llr.new_empty((B, dec.graph.NZ)) if app_mode == 1 else llr.new_empty((T_run, B, dec.graph.NZ)))
    return [llr.new_empty((B, dec.hard_words), dtype=torch.int32), llr.new_empty((B,), dtype=torch.int32),
            llr.new_empty((B,), dtype=torch.uint8), llr.new_empty((B,), dtype=torch.int32), app]


def _decode_via_op(dec: NMSDecoder, llr, iters, early_term, app, want_hard, unpack) -> DecodeResult:
    """Routes NMSDecoder.decode through torch.ops.ldpc_b200.nms_decode and re-wraps the outputs."""
    hard, it, fl, be, app_t = torch.ops.ldpc_b200.nms_decode(
        llr, op_handle(dec), int(iters), bool(early_term), {None: 0, "last": 1, "all": 2}[app])
    return DecodeResult(unpack_bits(hard, dec.graph.NZ) if unpack else None, hard if want_hard else None, it, fl,
                        be, app_t if app is not None else None)
