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


def run_campaign(dec, snr_db: Sequence[float], max_frames: int, min_errors: Optional[int] = None,
                 early_term: bool = True, iters: int = 0, seed: int = 2044, chunk_frames: int = 1 << 21,
                 harvest: bool = False, max_uncor: int = 0, post_dec=None, post_iters: int = 0, group=None,
                 use_ref_rate: bool = True, log=None) -> List[dict]:
    """Returns one dict per Eb/N0 point (counters, rates, 95 % interval, frames/s; `post` = the post decoder's
    counters on the harvested words when `post_dec` is given) plus the harvested rows under key 'rows'."""
    import torch
    from .montecarlo import MonteCarlo
    mc = MonteCarlo(dec, seed=seed, chunk_frames=chunk_frames, group=group)
    g = dec.graph
    out = []
    for k, snr in enumerate(snr_db):
        sigma = float(g.sigma([snr], use_ref_rate=use_ref_rate)[0])
        t0 = time.time()
        pt, rows = mc.run_point(snr, int(max_frames), sigma=sigma, iters=iters, early_term=early_term,
                                harvest=_lib.HARVEST_UNCOR_ANY if (harvest or post_dec is not None) else _lib.HARVEST_NONE,
                                max_uncor=max_uncor, min_frame_errors=min_errors, frame_base=k * (1 << 40))
        dt = time.time() - t0
        lo, hi = pt.fer_ci95("any")
        rec = {"snr_db": float(snr), "sigma": sigma, "frames": pt.frames, "frame_err_any": pt.frame_err_any,
               "frame_err_last": pt.frame_err_last, "bit_err_last": pt.bit_err_last, "fer": pt.fer,
               "fer_last": pt.fer_last, "ber_last": pt.ber_last, "fer_ci95": [lo, hi], "avg_iters": pt.avg_iters,
               "synd_fail": pt.synd_fail, "undetected": pt.undetected, "harvested": pt.harvested,
               "rows_kept": int(rows.shape[0]), "seconds": dt, "frames_per_s": pt.frames / max(dt, 1e-9)}
        if post_dec is not None and rows.shape[0] > 0 and mc.rank == 0:
            words = torch.from_numpy(np.ascontiguousarray(rows)).to(post_dec.device)
            cnt, _ = post_dec.post_decode(words, iters=post_iters, early_term=False)
            c = dict(zip(_lib.COUNTER_NAMES, (int(v) for v in cnt.cpu().numpy())))
            rec["post"] = {"words": c["frames"], "still_uncor_any": c["frame_err_any"],
                           "still_uncor_last": c["frame_err_last"], "bit_err_last": c["bit_err_last"],
                           "fer_after_post": pt.fer * c["frame_err_any"] / max(c["frames"], 1)
                           if rows.shape[0] >= pt.harvested else None}
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
    ap.add_argument("--json", default=None)
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
        print(f"# {g.name or args.graph}: M={g.M} N={g.N} z={g.z} E={g.E} k={g.k_true} n={g.n_true}  kernel {dec.kernel_name}  "
              f"{world} GPU(s)", flush=True)
    recs = run_campaign(dec, args.snr, int(args.frames), args.min_errors, early_term=not args.no_early_term,
                        seed=args.seed, chunk_frames=args.chunk, harvest=args.harvest is not None,
                        max_uncor=args.max_uncor if (args.harvest or post is not None) else 0, post_dec=post,
                        post_iters=args.post_iters, use_ref_rate=not args.true_rate,
                        log=lambda r: print(_fmt(r), flush=True))
    if rank == 0:
        if args.harvest:
            for r in recs:
                if r["rows"].shape[0]:
                    formats.append_uncor(args.harvest, r["rows"])
        if args.json:
            with open(args.json, "w") as fh:
                json.dump([{k: v for k, v in r.items() if k != "rows"} for r in recs], fh, indent=1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
