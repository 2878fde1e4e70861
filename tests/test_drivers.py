"""Host side of the callers (drivers.py): config surface of main_Base.py / main_Post.py, Performance.txt text,
process_data, the Uncor split.  Where /root/reference exists the text is compared with what the reference's
own code prints; the committed golden (tests/golden/perf_text.npz, minted by the same test) stands in elsewhere."""
import io
import os
import re
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import golden_path
from ldpc_error_floor_b200 import drivers, formats
from oracle import ref_runner

GOLD = golden_path("perf_text.npz")
RESULTS = np.array([[5.613e-2, 1.991e-2, 2.528e-3, 9.549e-5, 0], [0.774, 0.334, 0.055, 0.002, 0],
                    [0.774, 0.333, 0.055, 0.002, 0], [0, 0, 0, 0, 0]], dtype=np.float32)


def _reference_text(tmp_path):
    """Header: exec of main_Base.py:91-103 verbatim on the shipped config; result block: Print_Functions.print_result."""
    src = open(os.path.join(ref_runner.REFERENCE_ROOT, "main_Base.py")).read().splitlines()
    cfg_lines = [ln for ln in src[21:63]]                       # :22-63 module-level config
    ns = {"np": np}
    exec("\n".join(cfg_lines), ns)
    ns.update(M_proto=6, N_proto=24, Num_edge_proto=88, code_rate=431 / 574, Perf_filename=str(tmp_path / "perf_ref.txt"))
    exec("\n".join(ln for ln in src[90:103]), ns)               # :91-103 header block
    header = open(ns["Perf_filename"]).read()
    _, pf = ref_runner.load_reference()
    os.makedirs(tmp_path / "Weights", exist_ok=True)
    (tmp_path / "Weights" / "C0_x_Weight_End20.txt").write_text("3 0 3\n\n")   # print_result copies it to _Opt_
    p = str(tmp_path / "perf_ref2.txt")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with redirect_stdout(io.StringIO()):
            opt, flag = pf.print_result(RESULTS, 100000, p, "C0_x", 20, 1, False, False)
            pf.print_result(RESULTS, opt, p, "C0_x", 20, 1, flag, True)
    finally:
        os.chdir(cwd)
    return header, open(p).read()


def test_performance_text_matches_reference(tmp_path):
    cfg = drivers.RunConfig(root=str(tmp_path))
    header = drivers.perf_header(cfg, cfg.SNR_Matrix, 6, 24, 88, 431 / 574)
    p = str(tmp_path / "perf.txt")
    opt, flag = drivers.print_result(RESULTS, 100000, p, "C0_x", 20, 1, False, False, root=str(tmp_path), quiet=True)
    drivers.print_result(RESULTS, opt, p, "C0_x", 20, 1, flag, True, root=str(tmp_path), quiet=True)
    block = open(p).read()
    assert flag and abs(opt - RESULTS[1].sum()) < 1e-6
    if ref_runner.reference_available():
        rh, rb = _reference_text(tmp_path)
        if not os.path.exists(GOLD):
            np.savez(GOLD, header=rh, block=rb)
        assert header == rh
        assert block == rb
    g = np.load(GOLD)
    assert header == str(g["header"]) and block == str(g["block"])


def test_post_config_matches_main_post():
    """RunConfig.post() carries exactly the lines where main_Post.py differs from main_Base.py (:25-26, 35-38, 53-55)."""
    c = drivers.RunConfig.post()
    assert (c.sharing, c.sampling_type, c.iters_max, c.fixed_iter, c.iter_step) == ([3, 3, 3], 1, 30, 20, 10)
    assert (c.valid_num, c.test_flag, c.test_num) == (5000, 1, 5000)
    assert c.out_filename == "C0_wman_N0576_R34_z24"
    if ref_runner.reference_available():
        for fn, cfg in (("main_Base.py", drivers.RunConfig()), ("main_Post.py", c)):
            src = open(os.path.join(ref_runner.REFERENCE_ROOT, fn)).read().splitlines()
            ns = {"np": np}
            exec("\n".join(src[21:63]), ns)
            for name in ("filename", "sharing", "sampling_type", "decoding_type", "q_bit", "systematic", "z_value",
                         "punct_start", "punct_end", "short_start", "short_end", "iters_max", "fixed_iter",
                         "fixed_init", "iter_step", "loss_type", "opt_result_print", "batch_size", "training_num",
                         "epoch_input", "valid_flag", "valid_num", "test_flag", "test_num", "init_from_file",
                         "init_weight", "init_VN_weight", "Max_weight", "Min_weight", "seed_in"):
                assert getattr(cfg, name) == ns[name], (fn, name)
            assert np.array_equal(cfg.SNR_Matrix, ns["SNR_Matrix"])


def test_process_data_and_split(tmp_path):
    rng = np.random.RandomState(0)
    llr = np.clip(np.rint(rng.normal(-3, 3, (30, 576)) * 2) / 2, -7.5, 7.5).astype(np.float32)
    formats.append_uncor(str(tmp_path / "Uncor.txt"), llr)
    outs = drivers.split_uncor(str(tmp_path / "Uncor.txt"), "wman_N0576_R34_z24", 12, 8, 6, root=str(tmp_path))
    assert [os.path.basename(o) for o in outs] == ["[Uncor]_wman_N0576_R34_z24.txt", "[Uncor]_wman_N0576_R34_z24_Valid.txt",
                                                   "[Uncor]_wman_N0576_R34_z24_Test.txt"]
    cfg = drivers.RunConfig.post(root=str(tmp_path), training_num=10, valid_num=8, test_num=5)
    tr, ytr, va, yva, te, yte = drivers.process_data(cfg)
    assert tr.shape == (10, 576) and va.shape == (8, 576) and te.shape == (5, 576)
    assert np.array_equal(tr, -llr[:10]) and np.array_equal(va, -llr[12:20]) and np.array_equal(te, -llr[20:25])
    assert ytr.dtype == np.int64 and not ytr.any() and not yte.any()
    with pytest.raises(ValueError):
        drivers.process_data(drivers.RunConfig.post(root=str(tmp_path), training_num=13))   # "Wrong input" -> sys.exit()
    assert drivers.process_data(drivers.RunConfig(root=str(tmp_path))) == ([], [], [], [], [], [])
    if ref_runner.reference_available():
        mf, _ = ref_runner.load_reference()
        cwd = os.getcwd()
        os.chdir(tmp_path)
        try:
            ref = mf.process_data(1, "wman_N0576_R34_z24", 10, 1, 8, 1, 5)
        finally:
            os.chdir(cwd)
        for a, b in zip(ref, (tr, ytr, va, yva, te, yte)):
            assert np.array_equal(np.asarray(a), np.asarray(b))


def test_load_block_weights_rules(tmp_path, codes):
    """weight_init: rows before the block from ..._Opt_Weight_End{start}.txt, the block's rows = init constants."""
    import types
    os.makedirs(tmp_path / "Weights")
    with open(tmp_path / "Weights" / "C0_wman_N0576_R34_z24_Opt_Weight_End20.txt", "w") as fh:
        fh.write(str(codes["weights/wimax_base20/text"]))
    g = types.SimpleNamespace(M=6, N=24, E=88)
    cfg = drivers.RunConfig.post(root=str(tmp_path))
    ws = drivers.load_block_weights(cfg, g, 20, 30)
    assert ws.sharing == [3, 3, 3] and ws.iterations == 30
    for i in range(3):
        assert np.array_equal(ws.blocks[i][:20], codes[f"weights/wimax_base20/block{i}"].reshape(20, -1))
        assert np.all(ws.blocks[i][20:] == 1.0)
    base = drivers.load_block_weights(drivers.RunConfig(root=str(tmp_path)), g, 0, 20)
    assert base.sharing == [3, 0, 3] and set(base.blocks) == {0, 2} and np.all(base.blocks[0] == 1.0)


def test_random_weight_init_and_temporal_tie(tmp_path):
    """init_weight = -1 (Main_Functions.py:427-428): truncated normal around (Min + Max) / 2, stddev 0.1, nothing beyond
    two standard deviations; with temporal sharing (code 4) the rows of iterations >= fixed_iter are ONE variable."""
    from ldpc_error_floor_b200 import drivers
    from ldpc_error_floor_b200.graph import BaseGraph
    d = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "codes.npz")))
    g = BaseGraph(d["graph/wimax/proto"].astype(np.int32), 24)
    cfg = drivers.RunConfig(root=str(tmp_path), sharing=[1, 0, 2], init_weight=-1, init_VN_weight=-1, iters_max=8, iter_step=8)
    ws = drivers.load_block_weights(cfg, g, 0, 8)
    assert ws.blocks[0].shape == (8, g.E) and ws.blocks[2].shape == (8, g.N)
    for b in (ws.blocks[0], ws.blocks[2]):
        assert np.all(np.abs(b - 1.0) <= 0.2 + 1e-6) and 0.07 < b.std() < 0.1 and abs(b.mean() - 1.0) < 0.02
    assert not np.array_equal(ws.blocks[0][0], ws.blocks[0][1])
    again = drivers.load_block_weights(cfg, g, 0, 8)
    assert np.array_equal(again.blocks[0], ws.blocks[0])            # seeded by the run's seed_in
    cfg4 = drivers.RunConfig(root=str(tmp_path), sharing=[4, 0, 2], init_weight=-1, iters_max=8, iter_step=5, fixed_iter=3)
    w4 = drivers.load_block_weights(cfg4, g, 0, 8)
    assert all(np.array_equal(w4.blocks[0][t], w4.blocks[0][3]) for t in range(3, 8))
    assert not np.array_equal(w4.blocks[0][2], w4.blocks[0][3])
