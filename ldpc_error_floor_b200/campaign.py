"""Monte-Carlo FER/BER campaign over Eb/N0 points, sharded over the GPUs of one box.

    python -m ldpc_error_floor_b200.campaign --graph BaseGraph/wman_N0576_R34_z24.txt --z 24 \
        --weights Weights/C0_wman_N0576_R34_z24_Opt_Weight_End20.txt --snr 2 2.5 3 3.5 4 --frames 1e7 --min-errors 100
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m ldpc_error_floor_b200.campaign ... (one rank per GPU)

This is the evaluation loop of Print_Functions.compute_results (:130-165) grown into a campaign: per point it
stops at `--min-errors` frame errors or `--frames` frames, harvests the never-corrected words (the reference's
sampling_type 2, :155-156) and, with `--post-weights`, runs the boosted post decoder on the compacted failures
(main_Post.py).  Frame indices are global, so the result does not depend on the number of ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import List, Optional, Sequence

import numpy as np

from . import _lib, formats
from .graph import BaseGraph


CHECKPOINT_VERSION = 1


def _decoder_signature(dec) -> dict:
    """What a checkpoint must agree on before it may be continued: graph, weights, arithmetic."""
    import hashlib
    h = hashlib.sha256()
    g = dec.graph
    h.update(np.ascontiguousarray(getattr(g, "proto", np.zeros(0)), dtype=np.int32).tobytes())
    for b in getattr(dec, "_blocks", []) or []:
        if b is not None:
            h.update(np.ascontiguousarray(b, dtype=np.float32).tobytes())
    return {"NZ": int(g.NZ), "z": int(getattr(g, "z", 0)), "T": int(getattr(dec, "T", 0)),
            "sharing": [int(v) for v in getattr(dec, "sharing", [])], "decoding_type": int(getattr(dec, "decoding_type", 2)),
            "q_bit": int(getattr(dec, "q_bit", 5)), "target_node": int(getattr(dec, "target_node", 0)),
            "sha256": h.hexdigest()[:16]}


def _save_checkpoint(path: str, state: dict, rows: dict) -> None:
    """state -> <path> (JSON), harvested rows of point k -> <path>.rows<k>.npy; both replaced atomically."""
    for k, r in rows.items():
        tmp = f"{path}.rows{k}.tmp.npy"
        r = np.ascontiguousarray(r, dtype=np.float32)
        h = r.astype(np.float16)
        np.save(tmp, h if np.array_equal(h.astype(np.float32), r) else r)   # on-grid words are exact in half the bytes
        os.replace(tmp, f"{path}.rows{k}.npy")
    tmp = path + ".tmp"
    with open(tmp, "w") as fh:
        json.dump(state, fh, indent=1)
    os.replace(tmp, path)


def load_checkpoint(path: str):
    """(state dict, {point index: rows}) of a campaign checkpoint, or (None, {}) when there is none."""
    if not path or not os.path.exists(path):
        return None, {}
    with open(path) as fh:
        state = json.load(fh)
    if state.get("version") != CHECKPOINT_VERSION:
        raise ValueError(f"{path}: checkpoint version {state.get('version')} != {CHECKPOINT_VERSION}")
    rows = {}
    for k, p in enumerate(state["points"]):
        f = f"{path}.rows{k}.npy"
        if p.get("rows_kept", 0) > 0 and os.path.exists(f):
            rows[k] = np.load(f)[:p["rows_kept"]].astype(np.float32)
    return state, rows


def run_campaign(dec, snr_db: Sequence[float], max_frames: int, min_errors: Optional[int] = None,
                 early_term: bool = True, iters: int = 0, seed: int = 2044, chunk_frames: int = 1 << 21,
                 harvest: bool = False, max_uncor: int = 0, post_dec=None, post_iters: int = 0, group=None,
                 use_ref_rate: bool = True, log=None, checkpoint: Optional[str] = None, resume: bool = False,
                 checkpoint_rounds: int = 8, round_chunks: int = 4) -> List[dict]:
    """Returns one dict per Eb/N0 point (counters, rates, 95 % interval, frames/s; `post` = the post decoder's
    counters on the harvested words when `post_dec` is given) plus the harvested rows under key 'rows'.

    checkpoint: path of a JSON state file rank 0 rewrites every `checkpoint_rounds` rounds of every point (seed, chunk size,
    decoder signature, per point: chunks done, the eight counters, harvested rows in <path>.rows<k>.npy) -- the Monte-Carlo
    counterpart of the reference resuming a training block from its last weight file (Main_Functions.py:390-391, 419-422).
    resume: continue from that file: finished points are taken from it, the interrupted one restarts at its first
    undecoded chunk.  Frame indices are global, so the resumed run may use another number of GPUs and still ends with
    the counters and words of an uninterrupted run."""
    import torch
    from .montecarlo import MonteCarlo
    mc = MonteCarlo(dec, seed=seed, chunk_frames=chunk_frames, group=group)
    g = dec.graph
    sig = {"seed": int(seed), "chunk_frames": int(chunk_frames), "max_frames": int(max_frames), "iters": int(iters),
           "early_term": bool(early_term), "min_errors": None if min_errors is None else int(min_errors),
           "harvest": bool(harvest or post_dec is not None), "max_uncor": int(max_uncor), "use_ref_rate": bool(use_ref_rate),
           "snr_db": [float(v) for v in snr_db], "decoder": _decoder_signature(dec)}
    state, saved_rows = (load_checkpoint(checkpoint) if resume else (None, {}))
    if state is not None and state["config"] != sig:
        diff = [k for k in sig if state["config"].get(k) != sig[k]]
        raise ValueError(f"{checkpoint}: checkpoint was written by a different campaign (differs in {diff})")
    if state is None:
        state = {"version": CHECKPOINT_VERSION, "config": sig, "points": []}
    all_rows = dict(saved_rows)
    out = []
    for k, snr in enumerate(snr_db):
        sigma = float(g.sigma([snr], use_ref_rate=use_ref_rate)[0])
        prev = state["points"][k] if k < len(state["points"]) else None
        if prev is None:
            state["points"].append({"snr_db": float(snr), "sigma": sigma, "frame_base": k * (1 << 40), "chunks_done": 0,
                                    "counters": [0] * _lib.NUM_COUNTERS, "done": False, "seconds": 0.0, "rows_kept": 0})
            prev = state["points"][k]
        st = prev
        t0 = time.time()
        spent0 = float(st["seconds"])             # seconds of the run(s) this point was resumed from

        def on_ckpt(chunks_done, counters, rows, st=st, k=k, t0=t0, spent=spent0):
            st.update(chunks_done=int(chunks_done), counters=[int(v) for v in counters], rows_kept=int(rows.shape[0]),
                      seconds=spent + time.time() - t0)
            if checkpoint and mc.rank == 0:
                all_rows[k] = rows
                _save_checkpoint(checkpoint, state, {k: rows})

        if st["done"]:
            from .montecarlo import SnrPoint
            pt = SnrPoint(float(snr), sigma, bits_per_frame=g.NZ)
            pt.add(st["counters"])
            pt.chunks_done = st["chunks_done"]
            rows = all_rows.get(k, np.zeros((0, g.NZ), dtype=np.float32))
            dt = st["seconds"]
        else:
            pt, rows = mc.run_point(snr, int(max_frames), sigma=sigma, iters=iters, early_term=early_term,
                                    harvest=_lib.HARVEST_UNCOR_ANY if (harvest or post_dec is not None) else _lib.HARVEST_NONE,
                                    max_uncor=max_uncor, min_frame_errors=min_errors, frame_base=st["frame_base"],
                                    start_chunk=st["chunks_done"], init_counters=st["counters"], init_rows=all_rows.get(k),
                                    on_checkpoint=on_ckpt if checkpoint else None, checkpoint_rounds=checkpoint_rounds,
                                    round_chunks=round_chunks)
            dt = spent0 + time.time() - t0
            st.update(chunks_done=int(pt.chunks_done), done=True, seconds=dt, rows_kept=int(rows.shape[0]),
                      counters=[int(getattr(pt, n)) for n in _lib.COUNTER_NAMES])
            if checkpoint and mc.rank == 0:
                all_rows[k] = rows
                _save_checkpoint(checkpoint, state, {k: rows})
        lo, hi = pt.fer_ci95("any")
        rec = {"snr_db": float(snr), "sigma": sigma, "frames": pt.frames, "frame_err_any": pt.frame_err_any,
               "frame_err_last": pt.frame_err_last, "bit_err_last": pt.bit_err_last, "fer": pt.fer,
               "fer_last": pt.fer_last, "ber_last": pt.ber_last, "fer_ci95": [lo, hi], "avg_iters": pt.avg_iters,
               "synd_fail": pt.synd_fail, "undetected": pt.undetected, "harvested": pt.harvested,
               "rows_kept": int(rows.shape[0]), "seconds": dt, "frames_per_s": pt.frames / max(dt, 1e-9)}
        if post_dec is not None and rows.shape[0] > 0 and mc.rank == 0:
            words = torch.from_numpy(np.ascontiguousarray(rows)).to(post_dec.device)
            cnt, pres = post_dec.post_decode(words, iters=post_iters, early_term=False)
            c = dict(zip(_lib.COUNTER_NAMES, (int(v) for v in cnt.cpu().numpy())))
            rec["survivors"] = rows[(pres.flags.cpu().numpy() & _lib.FLAG_UNCOR_ANY) != 0]   # words the post decoder leaves
            rec["post"] = {"words": c["frames"], "still_uncor_any": c["frame_err_any"],
                           "still_uncor_last": c["frame_err_last"], "bit_err_last": c["bit_err_last"],
                           "fer_after_post": pt.fer * c["frame_err_any"] / max(c["frames"], 1)
                           if rows.shape[0] >= pt.harvested else None}
            if rec["post"]["fer_after_post"] is not None:
                # Wilson interval of "base decoder fails AND the post decoder does not repair it" over all frames
                from .montecarlo import SnrPoint
                q = SnrPoint(float(snr), sigma)
                q.add([pt.frames, 0, c["frame_err_any"], 0, 0, 0, 0, 0])
                rec["post"]["fer_after_post_ci95"] = list(q.fer_ci95("any"))
        rec["rows"] = rows
        out.append(rec)
        if log is not None and mc.rank == 0:
            log(rec)
    return out


def _fmt(rec: dict) -> str:
    s = (f"Eb/N0 {rec['snr_db']:5.2f} dB  frames {rec['frames']:>13d}  FER {rec['fer']:.3e} "
         f"[{rec['fer_ci95'][0]:.2e}, {rec['fer_ci95'][1]:.2e}]  FER_last {rec['fer_last']:.3e}  BER_last {rec['ber_last']:.3e}  "
         f"avg it {rec['avg_iters']:5.2f}  undetected {rec['undetected']}  {rec['frames_per_s'] / 1e6:7.2f} Mframes/s")
    if "post" in rec:
        p = rec["post"]
        s += f"  | post: {p['words']} words -> {p['still_uncor_any']} still uncorrected"
        if p.get("fer_after_post") is not None:
            ci = p.get("fer_after_post_ci95", [float("nan")] * 2)
            s += f"  FER after post {p['fer_after_post']:.3e} [{ci[0]:.2e}, {ci[1]:.2e}]"
    return s


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--graph", required=True, help="BaseGraph/*.txt (format F1)")
    ap.add_argument("--z", type=int, default=None)
    ap.add_argument("--punct", type=int, nargs=2, default=(0, 0))
    ap.add_argument("--short", type=int, nargs=2, default=(0, 0))
    ap.add_argument("--weights", default=None, help="Weights/*.txt (format F2); default: plain min-sum, weight --ms-weight")
    ap.add_argument("--ms-weight", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--decoding-type", type=int, default=2)
    ap.add_argument("--q-bit", type=int, default=5)
    ap.add_argument("--snr", type=float, nargs="+", required=True)
    ap.add_argument("--frames", type=float, default=1e7, help="frames per point (upper bound)")
    ap.add_argument("--min-errors", type=int, default=None)
    ap.add_argument("--no-early-term", action="store_true")
    ap.add_argument("--true-rate", action="store_true", help="sigma from k/n without the reference's +1 quirk")
    ap.add_argument("--systematic", action="store_true",
                    help="count errors over the N - M information columns only (main_Base.py:29 systematic = 1)")
    ap.add_argument("--seed", type=int, default=2044)
    ap.add_argument("--chunk", type=int, default=1 << 21)
    ap.add_argument("--harvest", default=None, help="append never-corrected words to this file (Inputs/[Uncor] format)")
    ap.add_argument("--max-uncor", type=int, default=100000)
    ap.add_argument("--post-weights", default=None, help="boosted weight file: run the post decoder on the failures")
    ap.add_argument("--post-iters", type=int, default=0)
    ap.add_argument("--survivors", default=None,
                    help="write the words the post decoder still fails on, all points, as an LDPCQ8 file (formats.write_uncor_q8)")
    ap.add_argument("--json", default=None)
    ap.add_argument("--checkpoint", default=None, help="campaign state file (JSON + .rows<k>.npy), rewritten as the run goes")
    ap.add_argument("--checkpoint-rounds", type=int, default=8, help="rounds (4 chunks per GPU each) between checkpoints")
    ap.add_argument("--resume", action="store_true", help="continue from --checkpoint (any number of GPUs)")
    args = ap.parse_args(argv)

    import torch
    import torch.distributed as dist
    from .decoder import NMSDecoder
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print("campaign needs a CUDA device: this framework has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank = dist.get_rank() if world > 1 else 0

    g = BaseGraph.from_file(args.graph, z=args.z, punct=tuple(args.punct), short=tuple(args.short))
    if args.weights:
        ws = formats.read_weights(args.weights)
    else:
        ws = formats.WeightSet([3, 0, 0], {0: np.full((args.iters, 1), args.ms_weight, dtype=np.float32)})
    dec = NMSDecoder(g, ws, iters=args.iters, decoding_type=args.decoding_type, q_bit=args.q_bit, device=local_rank,
                     systematic=1 if args.systematic else 0)
    post = None
    if args.post_weights:
        pws = formats.read_weights(args.post_weights)
        post = NMSDecoder(g, pws, iters=args.post_iters or None, decoding_type=args.decoding_type, q_bit=args.q_bit,
                          device=local_rank, systematic=1 if args.systematic else 0)
    if rank == 0:
        mk = dec.mc_info()["kernel"] if not args.no_early_term else dec.launch_info(False)["kernel"]
        print(f"# {g.name or args.graph}: M={g.M} N={g.N} z={g.z} E={g.E} k={g.k_true} n={g.n_true}  kernel {mk}  "
              f"{world} GPU(s)", flush=True)
    recs = run_campaign(dec, args.snr, int(args.frames), args.min_errors, early_term=not args.no_early_term,
                        seed=args.seed, chunk_frames=args.chunk, harvest=args.harvest is not None,
                        max_uncor=args.max_uncor if (args.harvest or post is not None) else 0, post_dec=post,
                        post_iters=args.post_iters, use_ref_rate=not args.true_rate,
                        log=lambda r: print(_fmt(r), flush=True), checkpoint=args.checkpoint, resume=args.resume,
                        checkpoint_rounds=args.checkpoint_rounds)
    if rank == 0:
        if args.harvest:
            for r in recs:
                if r["rows"].shape[0]:
                    formats.append_uncor(args.harvest, r["rows"])
        if args.survivors:
            step = dec.q8_step or 0.5
            for r in recs:
                if "survivors" in r and r["survivors"].shape[0]:
                    formats.write_uncor_q8(args.survivors, formats.llr_to_q8(np.clip(r["survivors"], -127 * step, 127 * step), step),
                                           step=step, snr_db=r["snr_db"], seed=args.seed, append=os.path.exists(args.survivors))
        if args.json:
            with open(args.json, "w") as fh:
                json.dump({"world_size": world, "kernel": dec.mc_info(), "graph": g.name or args.graph,
                           "points": [{k: v for k, v in r.items() if k not in ("rows", "survivors")} for r in recs]}, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
