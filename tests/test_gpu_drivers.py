"""GPU: the callers end to end through the reference's file layout -- main_Base.py as collector
(sampling_type 2 -> ./Uncor.txt), the manual split into Inputs/[Uncor]_*, main_Post.py's evaluation of the
30-iteration decoder on those words, Performance.txt -- and the campaign runner."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    import materialize_files
    root = tmp_path_factory.mktemp("ref_layout")
    return str(root), materialize_files.materialize(str(root))


def test_collect_split_post_flow(files):
    from ldpc_error_floor_b200 import drivers, formats
    from oracle import c_oracle
    root, made = files
    # 1. base decoder, collection run (main_Base.py with sampling_type = 2, shipped 20-iteration weights)
    base_w = formats.read_weights(made["w:wimax_base20"])
    cfg = drivers.RunConfig(root=root, sharing=[3, 3, 3], sampling_type=2, SNR_Matrix=np.array([2.5]), valid_num=20000)
    out = drivers.evaluate(cfg, weights=base_w, quiet=True)
    res = out["valid"]
    n_uncor = sum(1 for _ in open(os.path.join(root, "Uncor.txt")))
    assert res.shape == (4, 1) and 0.25 < res[2, 0] < 0.42            # reference anchor at 2.5 dB: FER 0.333
    assert n_uncor == round(float(res[2, 0]) * 20000)
    assert res[1, 0] >= res[2, 0] and res[0, 0] > 0
    perf = open(cfg.perf_filename).read()
    assert perf.startswith("Decoding_type = 2 q_bit = 5\n") and "Valid_Result\nBER_last: ['" in perf
    # 2. the split the authors do by hand, then main_Post.py's evaluation (rows 20..29 = init weight 1.0)
    drivers.split_uncor(os.path.join(root, "Uncor.txt"), cfg.filename, 3000, 1500, 1500, root=root)
    pcfg = drivers.RunConfig.post(root=root, training_num=3000, valid_num=1500, test_num=1500)
    pout = drivers.evaluate(pcfg, quiet=True)
    assert pout["iters"] == 30 and pout["SNR_Matrix"].tolist() == [0.0]
    perf = open(pcfg.perf_filename).read()
    assert "Valid_Result" in perf and "Test_Result" in perf and "Running time (Train/Valid/Test)" in perf
    # 3. same numbers from the CPU oracle on the same stored words
    rows = formats.read_uncor(os.path.join(root, "Inputs", "[Uncor]_wman_N0576_R34_z24_Test.txt"), 1500)
    xa = formats.uncor_to_llr(rows, 24, 24)
    ws = drivers.load_block_weights(pcfg, pout["decoder"].graph, 20, 30)
    ref = c_oracle.decode(pout["decoder"].graph.proto, 24, xa, ws.sharing, ws.blocks, 30, 2, 5, 20.0, want_all=True)
    hard = ref["app"] >= 0                                              # [T, B, NZ]
    uncor_t = hard.any(axis=2)
    fer = uncor_t.min(axis=0).mean()
    fer_last = uncor_t[-1].mean()
    ber_last = hard[-1].sum() / hard[-1].size
    t = pout["test"]
    assert abs(t[2, 0] - fer) < 1e-6 and abs(t[1, 0] - fer_last) < 1e-6 and abs(t[0, 0] - ber_last) < 1e-7
    assert fer < 1.0                                                     # ten more iterations correct some words


def test_evaluate_systematic_counts_information_columns_only(files, codes):
    """systematic = 1 (main_Base.py:29, 83-86): drivers.evaluate must hand it to the decoder -- BER / FER over the first
    N - M proto columns, as the performance header says.  Checked against a decoder built directly with systematic = 1 and
    against the reference's own compute_results run for this configuration (mc_ref_5g_r050_z64.npz)."""
    import ldpc_error_floor_b200 as L
    from conftest import golden_path
    from ldpc_error_floor_b200 import drivers, formats
    from ldpc_error_floor_b200.montecarlo import SnrPoint, compute_results
    root, made = files
    key = "5g_r050_z64"
    stem = str(codes[f"graph/{key}/stem"])
    z, ps, pe, ss, se, _ = (int(v) for v in codes[f"graph/{key}/meta"])
    w = formats.read_weights(made["w:5g_r050_z64_boost50"]).rows(0, 20)
    out = {}
    for systematic in (0, 1):
        cfg = drivers.RunConfig(root=root, filename=stem, sharing=[2, 2, 2], sampling_type=0, systematic=systematic,
                                z_value=z, punct_start=ps, punct_end=pe, short_start=ss, short_end=se,
                                SNR_Matrix=np.array([1.5, 2.0]), valid_num=200000)
        out[systematic] = drivers.evaluate(cfg, weights=w, quiet=True)["valid"]
        assert f"systematic = {systematic}\n" in open(cfg.perf_filename).read()
    g = L.BaseGraph(codes[f"graph/{key}/proto"].astype(np.int32), z, (ps, pe), (ss, se))
    dec = L.NMSDecoder(g, w, iters=20, systematic=1)
    direct, _ = compute_results(dec, 200000, None, g.sigma([1.5, 2.0]), 20, 0, seed=1074 + cfg.seed_in)
    assert np.array_equal(out[1], direct)
    assert (out[1][1] < out[0][1]).all()                 # parity-column errors no longer count
    ref = np.load(golden_path("mc_ref_5g_r050_z64.npz"))
    assert int(ref["systematic"]) == 1
    for k in range(2):
        pt = SnrPoint(float(ref["snr"][k]), float(ref["sigma"][k]))
        pt.add([int(ref["frames"]), int(ref[f"uncor_last_{k}"].sum()), int(ref[f"uncor_any_{k}"].sum()), 0, 0, 0, 0, 0])
        lo, hi = pt.fer_ci95("last")
        assert lo <= out[1][1, k] <= hi, (k, out[1][1, k], lo, hi)


def test_campaign_cli(files, tmp_path):
    import json
    from ldpc_error_floor_b200 import campaign
    root, made = files
    js = str(tmp_path / "c.json")
    harvest = str(tmp_path / "Uncor.txt")
    rc = campaign.main(["--graph", made["5g_r050_z64"], "--weights", made["w:5g_r050_z64_boost50"], "--iters", "20",
                        "--snr", "1.5", "2.5", "--frames", "200000", "--min-errors", "200", "--chunk", "32768",
                        "--harvest", harvest, "--max-uncor", "5000", "--post-weights", made["w:5g_r050_z64_boost50"],
                        "--json", js])
    assert rc == 0
    recs = json.load(open(js))["points"]
    assert [r["snr_db"] for r in recs] == [1.5, 2.5]
    assert recs[0]["fer"] > recs[1]["fer"] > 0 and recs[0]["frames"] < 200000     # stopped on the error target
    for r in recs:
        assert r["fer_ci95"][0] <= r["fer"] <= r["fer_ci95"][1]
        assert r["post"]["words"] == r["rows_kept"] > 0
        assert r["post"]["still_uncor_any"] <= r["post"]["words"]                  # 50 iterations fix some of them
    assert sum(1 for _ in open(harvest)) == sum(r["rows_kept"] for r in recs)
