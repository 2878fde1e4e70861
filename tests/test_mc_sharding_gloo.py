"""CPU, world_size 2 over gloo: the Monte-Carlo sharding logic (ldpc_error_floor_b200.montecarlo).
Frames are cut into chunks, chunk c goes to rank c % world, the sample stream is keyed by the GLOBAL
frame index, counters are all-reduced and harvested words all-gathered -- so the result must not
depend on the number of ranks.  The device work is stubbed by the C oracle (tests may use oracle/)."""
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleDecoder:
    """Stands in for NMSDecoder.mc_run: same counters / harvest contract, arithmetic from oracle/nms_oracle.c."""

    def __init__(self, fail_after=None):
        from oracle import c_oracle
        self.fail_after, self.calls = fail_after, 0
        self.c_oracle = c_oracle
        d = dict(np.load(os.path.join(ROOT, "tests", "golden", "decode_mackay_qms_300_t20.npz")))
        self.proto = d["proto"].astype(np.int32)
        self.w = {0: d["w0"]}
        self.N, self.z, self.T = self.proto.shape[1], 1, 20
        self.device = "cpu"
        self.graph = SimpleNamespace(NZ=self.N, N=self.N, z=1,
                                     sigma=lambda snr, use_ref_rate=True: np.sqrt(1.0 / (2.0 * 0.5 * 10 ** (np.asarray(snr, float) / 10))))

    def llr(self, sigma, seed, first, n):
        out = np.empty((n, self.N, 1), np.float32)
        for k in range(n):                                 # keyed by the global frame index, like Philox
            rng = np.random.default_rng([seed, first + k])
            x = 2.0 * (rng.normal(size=self.N) * sigma - 1.0) / sigma ** 2
            out[k, :, 0] = np.clip(np.rint(x * 2) / 2, -7.5, 7.5)
        return out

    def mc_run(self, sigma, n_frames, seed, frame_offset=0, iters=0, early_term=False, harvest=0, capacity=0,
               counters=None, uncor_buf=None, uncor_count=None):
        self.calls += 1
        if self.fail_after is not None and self.calls > self.fail_after:
            raise KeyboardInterrupt("simulated kill")            # every rank dies in the same round, as under torchrun
        xa = self.llr(sigma, seed, frame_offset, n_frames)
        r = self.c_oracle.decode(self.proto, 1, xa, [3, 0, 0], self.w, self.T, 2, 5, 20.0, nthreads=1)
        hard = r["app"] >= 0
        any_one = hard.any(axis=2)
        uncor_any, uncor_last = any_one.all(axis=0), any_one[-1]
        synd_fail = r["synd"][-1]
        c = np.array([n_frames, uncor_last.sum(), uncor_any.sum(), hard[-1].sum(), n_frames * self.T, synd_fail.sum(),
                      (uncor_last & ~synd_fail).sum(), uncor_any.sum() if harvest else 0], dtype=np.int64)
        counters = c if counters is None else counters + c
        if uncor_count is None:
            uncor_count = np.zeros(1, np.int64)
        if harvest and capacity:
            if uncor_buf is None:
                uncor_buf = np.zeros((capacity, self.N), np.float32)
            for row in xa[uncor_any].reshape(-1, self.N):
                if uncor_count[0] < capacity:
                    uncor_buf[uncor_count[0]] = row
                uncor_count[0] += 1
        return counters, uncor_buf, uncor_count


def _worker(rank, world, port, outdir, n_frames, min_err):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from ldpc_error_floor_b200.montecarlo import MonteCarlo
    mc = MonteCarlo(OracleDecoder(), seed=11, chunk_frames=50)
    pt, rows = mc.run_point(2.0, n_frames, early_term=False, harvest=1, max_uncor=n_frames,
                            min_frame_errors=min_err, round_chunks=2)
    np.savez(os.path.join(outdir, f"w{world}_r{rank}.npz"), frames=pt.frames, any=pt.frame_err_any,
             last=pt.frame_err_last, bits=pt.bit_err_last, iters=pt.iters, rows=rows)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _run(world, outdir, n_frames, min_err, port):
    if world == 1:
        _worker(0, 1, port, outdir, n_frames, min_err)
    else:
        mp.spawn(_worker, args=(world, port, outdir, n_frames, min_err), nprocs=world, join=True)
    return [dict(np.load(os.path.join(outdir, f"w{world}_r{r}.npz"))) for r in range(world)]


@pytest.mark.timeout(600)
def test_two_ranks_equal_one_rank():
    with tempfile.TemporaryDirectory() as tmp:
        one = _run(1, tmp, 730, None, 29511)[0]
        two = _run(2, tmp, 730, None, 29512)
    assert int(one["frames"]) == 730 and int(one["any"]) > 0
    for r in two:                                   # every rank holds the reduced result
        for k in ("frames", "any", "last", "bits", "iters"):
            assert int(r[k]) == int(one[k]), k
        assert r["rows"].shape == one["rows"].shape == (int(one["any"]), 96)
    key = lambda a: sorted(map(bytes, a))           # same words, rank-major order
    assert key(two[0]["rows"]) == key(one["rows"]) and np.array_equal(two[0]["rows"], two[1]["rows"])


@pytest.mark.timeout(600)
def test_stop_rule_is_collective():
    """min_frame_errors is checked on the all-reduced counter, so both ranks stop after the same round."""
    with tempfile.TemporaryDirectory() as tmp:
        two = _run(2, tmp, 4000, 5, 29513)
    assert int(two[0]["frames"]) == int(two[1]["frames"]) < 4000
    assert int(two[0]["frames"]) % 50 == 0 and int(two[0]["any"]) >= 5


# ------------------------------------------------------------------------------------------------ checkpoint / resume
def _campaign_worker(rank, world, port, outdir, tag, ckpt, resume, fail_after):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from ldpc_error_floor_b200 import campaign
    killed = False
    try:
        recs = campaign.run_campaign(OracleDecoder(fail_after), [2.0, 3.0], 600, early_term=False, seed=11, chunk_frames=50,
                                     harvest=True, max_uncor=600, checkpoint=ckpt, resume=resume, checkpoint_rounds=1, round_chunks=1)
        np.savez(os.path.join(outdir, f"{tag}_r{rank}.npz"), **{f"{k}_{i}": np.asarray(r[k]) for i, r in enumerate(recs)
                                                                for k in ("frames", "frame_err_any", "frame_err_last",
                                                                          "bit_err_last", "rows")})
    except KeyboardInterrupt:
        killed = True
    if world > 1:
        dist.destroy_process_group()
    if fail_after is not None:
        assert killed


def _campaign(world, outdir, tag, ckpt, resume, fail_after, port):
    if world == 1:
        _campaign_worker(0, 1, port, outdir, tag, ckpt, resume, fail_after)
    else:
        mp.spawn(_campaign_worker, args=(world, port, outdir, tag, ckpt, resume, fail_after), nprocs=world, join=True)
    f = os.path.join(outdir, f"{tag}_r0.npz")
    return dict(np.load(f)) if os.path.exists(f) else None


@pytest.mark.timeout(900)
def test_campaign_kill_and_resume():
    """A 2-rank campaign dies in the middle of its second Eb/N0 point; `resume` on ONE rank picks the state file up and ends
    with exactly the counters and harvested words of an uninterrupted run (checkpoint / resume)."""
    import json
    with tempfile.TemporaryDirectory() as tmp:
        ckpt = os.path.join(tmp, "state.json")
        whole = _campaign(1, tmp, "whole", None, False, None, 29521)
        assert _campaign(2, tmp, "killed", ckpt, False, 8, 29522) is None        # 12 chunks per point: dies inside point 1
        st = json.load(open(ckpt))
        assert st["points"][0]["done"] and not st["points"][1]["done"]
        assert 0 < st["points"][1]["chunks_done"] < 12 and st["points"][1]["counters"][0] == 50 * st["points"][1]["chunks_done"]
        resumed = _campaign(1, tmp, "resumed", ckpt, True, None, 29523)
        st2 = json.load(open(ckpt))
        assert all(p["done"] for p in st2["points"])
        # a different campaign may not continue this file
        from ldpc_error_floor_b200 import campaign
        with pytest.raises(ValueError):
            campaign.run_campaign(OracleDecoder(), [2.0, 3.0], 600, early_term=False, seed=12, chunk_frames=50, harvest=True,
                                  max_uncor=600, checkpoint=ckpt, resume=True)
    key = lambda a: sorted(map(bytes, a))
    for i in range(2):
        for k in ("frames", "frame_err_any", "frame_err_last", "bit_err_last"):
            assert int(resumed[f"{k}_{i}"]) == int(whole[f"{k}_{i}"]), (i, k)
        assert key(resumed[f"rows_{i}"]) == key(whole[f"rows_{i}"])
    assert int(whole["frames_0"]) == 600 and int(whole["frame_err_any_0"]) > 0
