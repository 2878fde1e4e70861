"""GPU: fused Monte-Carlo path (Philox generator + decode + counters + harvest), post decoder, and the
compute_results drop-in, checked against the oracle and the reference-minted fixtures."""
import os

import numpy as np
import pytest

from conftest import golden_path, load_case

pytestmark = pytest.mark.gpu


def make(case_name, **kw):
    import ldpc_error_floor_b200 as L
    case = load_case(case_name)
    g = L.BaseGraph(case["proto"], case["z"], case["punct"], case["short"])
    ws = L.WeightSet(case["sharing"], dict(case["weights"]))
    dec = L.NMSDecoder(g, ws, iters=case["T"], decoding_type=case["decoding_type"], q_bit=case["q_bit"],
                       clip_llr=case["clip"], **kw)
    return case, g, dec


@pytest.mark.parametrize("name", ["wimax_qms_333_t20", "5g_r050_z64_qms_222_t50", "wimax_float_333_t20"])
def test_generator_distribution_and_shortening(name):
    """llr ~ N(-2/s^2, 4/s^2) before quantisation (Print_Functions.py:45-46); punctured -> 0, shortened -> -clip."""
    case, g, dec = make(name)
    sigma = float(g.sigma([2.0])[0])
    x = dec.generate(sigma, 4000, seed=3).cpu().numpy().reshape(4000, -1)
    mask = np.ones(g.NZ, bool)
    if case["punct"][0] > 0:
        sl = slice(case["punct"][0] - 1, case["punct"][1])
        assert np.all(x[:, sl] == 0)
        mask[sl] = False
    if case["short"][0] > 0:
        sl = slice(case["short"][0] - 1, case["short"][1])
        assert np.all(x[:, sl] == -case["clip"])
        mask[sl] = False
    body = x[:, mask]
    if case["decoding_type"] == 2:
        assert np.all(body * 2 == np.rint(body * 2)) and np.abs(body).max() <= 7.5
        ref = np.clip(np.rint(np.random.RandomState(0).normal(-2 / sigma ** 2, 2 / sigma, body.shape) * 2) / 2, -7.5, 7.5)
        assert abs(body.mean() - ref.mean()) < 0.02 and abs(body.std() - ref.std()) < 0.02
    else:
        n = body.size
        assert abs(body.mean() + 2 / sigma ** 2) < 5 * (2 / sigma) / np.sqrt(n)
        assert abs(body.std() / (2 / sigma) - 1) < 0.01
        z = (body + 2 / sigma ** 2) / (2 / sigma)
        assert abs((z ** 3).mean()) < 0.02 and abs((z ** 4).mean() - 3) < 0.05      # Gaussian moments
    # frames are keyed by the global index: offsets tile the same stream
    a = dec.generate(sigma, 64, seed=3, frame_offset=100).cpu().numpy()
    assert np.array_equal(a.reshape(64, -1), x[100:164])
    assert not np.array_equal(dec.generate(sigma, 64, seed=4, frame_offset=100).cpu().numpy().reshape(64, -1), x[100:164])


@pytest.mark.parametrize("name,et", [("wimax_qms_333_t20", False), ("wimax_qms_333_t20", True),
                                     ("5g_r073_z32_qms_222_t50", True), ("mackay_float_300_t20", False)])
def test_mc_run_equals_generate_then_decode(name, et):
    """The fused kernel draws the same samples as ldpc_llr_generate and counts what ldpc_decode reports."""
    import torch
    import ldpc_error_floor_b200 as L
    case, g, dec = make(name)
    sigma = float(g.sigma([2.5 if "wimax" in name else 3.0])[0])
    n = 3001
    cnt, rows = dec.mc_run_host(sigma, n, seed=9, frame_offset=12345, early_term=et, harvest=L.HARVEST_UNCOR_ANY,
                                capacity=n)
    x = dec.generate(sigma, n, seed=9, frame_offset=12345)
    r = dec.decode(x, early_term=et, unpack=True)
    flags = r.flags.cpu().numpy()
    assert cnt["frames"] == n
    assert cnt["frame_err_any"] == int(((flags & 2) != 0).sum()) > 0
    assert cnt["frame_err_last"] == int(((flags & 4) != 0).sum())
    assert cnt["bit_err_last"] == int(r.biterr.sum().item())
    assert cnt["synd_fail"] == int(((flags & 1) == 0).sum())
    assert cnt["undetected"] == int((((flags & 1) != 0) & ((flags & 4) != 0)).sum())
    iters = r.iters.cpu().numpy()
    assert cnt["iters"] == (int(iters.sum()) if et else n * case["T"])
    # harvested rows are exactly the generated LLRs of the never-corrected frames (any order)
    want = x.cpu().numpy().reshape(n, -1)[(flags & 2) != 0]
    assert cnt["harvested"] == rows.shape[0] == want.shape[0]
    assert sorted(map(bytes, rows)) == sorted(map(bytes, want))
    # two halves of the index space add up to the whole (what the multi-GPU sharding relies on)
    c1, _ = dec.mc_run_host(sigma, 1500, seed=9, frame_offset=12345, early_term=et)
    c2, _ = dec.mc_run_host(sigma, n - 1500, seed=9, frame_offset=12345 + 1500, early_term=et)
    for k in ("frames", "frame_err_any", "frame_err_last", "bit_err_last", "iters"):
        assert c1[k] + c2[k] == cnt[k], k


def test_fer_inside_reference_confidence_interval():
    """WiMAX QMS 20-it FER at 2.5 / 3.0 dB against the reference's own compute_results run (mc_wimax.npz):
    our estimate from 200k frames must lie inside the 95 % interval of the reference's 200-frame estimate."""
    from ldpc_error_floor_b200.montecarlo import MonteCarlo, SnrPoint
    d = np.load(golden_path("mc_wimax.npz"))
    case, g, dec = make("wimax_qms_333_t20")
    mc = MonteCarlo(dec, seed=77, chunk_frames=1 << 16)
    for k, snr in enumerate(d["snr"]):
        assert float(g.sigma([snr])[0]) == pytest.approx(float(d["sigma"][k]), rel=1e-12)
        pt, _ = mc.run_point(float(snr), 200_000)
        ref = SnrPoint(float(snr), float(d["sigma"][k]), bits_per_frame=g.NZ)
        nref = int(d["sample_num"])
        ref.add([nref, round(float(d["results"][1, k]) * nref), round(float(d["results"][2, k]) * nref), 0, 0, 0, 0, 0])
        lo, hi = ref.fer_ci95("any")
        assert lo <= pt.fer <= hi, (snr, pt.fer, lo, hi)
        lo, hi = ref.fer_ci95("last")
        assert lo <= pt.fer_last <= hi
        ber_ref = float(d["results"][0, k])
        assert 0.5 * ber_ref <= pt.ber_last <= 2.0 * ber_ref


def test_post_decoder_on_harvested_words():
    """Config 4 flow on a small scale: base decoder (rows 0-19) harvests its failures, the boosted decoder
    (rows 0-49) re-decodes them; both legs agree with the C oracle on the same words."""
    import torch
    import ldpc_error_floor_b200 as L
    from oracle import c_oracle
    codes = dict(np.load(golden_path("codes.npz")))
    key = "5g_r050_z64"
    z, ps, pe, ss, se, _ = (int(v) for v in codes[f"graph/{key}/meta"])
    proto = codes[f"graph/{key}/proto"].astype(np.int32)
    g = L.BaseGraph(proto, z, (ps, pe), (ss, se))
    boosted = L.WeightSet([2, 2, 2], {i: codes[f"weights/{key}_boost50/block{i}"] for i in range(3)})
    base = L.NMSDecoder(g, boosted.rows(0, 20))
    post = L.NMSDecoder(g, boosted)
    sigma = float(g.sigma([1.5])[0])
    cnt, buf, ucnt = base.mc_run(sigma, 20000, seed=5, harvest=L.HARVEST_UNCOR_ANY, capacity=4096)
    n = min(int(ucnt.item()), 4096)
    assert 20 < n == int(cnt[2].item()) or n == 4096
    words = buf[:n]
    pc, res = post.post_decode(words)
    pc = pc.cpu().numpy()
    w_np = words.cpu().numpy().reshape(n, g.N, g.z)
    ref_base = c_oracle.decode(proto, z, w_np, [2, 2, 2], boosted.rows(0, 20).blocks, 20, 2, 5, 20.0)
    assert (ref_base["app"] >= 0).any(axis=2).all(axis=0).all()        # the base decoder never corrects them
    ref = c_oracle.decode(proto, z, w_np, [2, 2, 2], boosted.blocks, 50, 2, 5, 20.0)
    any_one = (ref["app"] >= 0).any(axis=2)
    assert pc[0] == n and pc[2] == int(any_one.all(axis=0).sum()) and pc[1] == int(any_one[-1].sum())
    assert pc[2] < n                                                   # boosting recovers some of them
    flags = res.flags.cpu().numpy()
    assert np.array_equal((flags & 1) != 0, ~ref["synd"][-1])


def test_compute_results_dropin(tmp_path):
    """sampling_type 2 harvests to an Inputs/[Uncor]-format file; sampling_type 1 decodes such a file."""
    from ldpc_error_floor_b200 import formats
    from ldpc_error_floor_b200.montecarlo import compute_results
    case, g, dec = make("wimax_qms_333_t20")
    sig = g.sigma([2.5])
    path = str(tmp_path / "Uncor.txt")
    res, took = compute_results(dec, 5010, None, sig, 20, 2, seed=3, uncor_path=path)
    assert res.shape == (4, 1) and res.dtype == np.float32 and took > 0
    rows = formats.read_uncor(path)
    assert rows.shape[0] == round(float(res[2, 0]) * 5000) > 100          # floor(5010/20)*20 frames
    res1, _ = compute_results(dec, rows.shape[0], rows, np.array([0.0]), 1, 1)
    assert res1[2, 0] == 1.0                                                # none of them is ever corrected
    assert 0.9 <= res1[1, 0] <= 1.0


def test_alu_peak_probe_is_plausible():
    """The on-box pipe micro-benchmark behind bench.py's roofline denominators: FP32 FMA issue close to 148 SMs x 128
    lanes x clock, the integer / min-max pipe at about half of it, packed half2 instructions at the FP32 rate or half."""
    import torch
    from ldpc_error_floor_b200 import _lib
    p = _lib.alu_peak_probe(0, kinds=("ffma", "fadd", "fmnmx", "lop3", "iadd", "hfma2", "hmnmx2"))
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    nominal = sms * 128 * 1.965e9 / 1e12
    assert 0.6 * nominal < p["ffma"] < 1.05 * nominal, p
    for k in ("fmnmx", "lop3", "iadd", "hmnmx2"):
        assert 0.25 * nominal < p[k] < 1.05 * nominal, (k, p)
    print(p)


@pytest.mark.parametrize("key,snr,systematic", [("wimax", 3.0, 0), ("wimax", 4.0, 0), ("5g_r050_z64", 2.5, 1), ("5g_r073_z72", 4.5, 1)])
def test_two_stage_monte_carlo_equals_single_stage(key, snr, systematic, codes, monkeypatch):
    """ldpc_mc_run_staged: stage 1 defers the frames without a zero syndrome by frame index, stage 2 regenerates and
    decodes them in full -- the eight counters and the set of harvested words are those of the one-launch run.
    (The batch kernels' form; graphs with a persistent-slot kernel no longer need it, hence the switch.)"""
    import torch
    monkeypatch.setenv("LDPC_B200_NO_PERSIST", "1")
    import ldpc_error_floor_b200 as L
    from ldpc_error_floor_b200 import _lib, montecarlo
    proto = codes[f"graph/{key}/proto"].astype(np.int32); meta = codes[f"graph/{key}/meta"]
    g = L.BaseGraph(proto, int(meta[0]), (int(meta[1]), int(meta[2])), (int(meta[3]), int(meta[4])))
    wk = {"wimax": "wimax_base20", "5g_r050_z64": "5g_r050_z64_boost50"}.get(key)
    if wk:
        ws = L.WeightSet([int(v) for v in codes[f"weights/{wk}/sharing"]], {i: codes[f"weights/{wk}/block{i}"] for i in range(3)})
    else:
        ws = L.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)})
    dec = L.NMSDecoder(g, ws, iters=20, systematic=systematic)
    sigma, n = float(g.sigma([snr])[0]), 60000
    c1, b1, u1 = dec.mc_run(sigma, n, 11, frame_offset=12345, early_term=True, harvest=_lib.HARVEST_UNCOR_ANY, capacity=4000)
    for s1 in (3, 8, 19):
        c2, b2, u2 = dec.mc_run(sigma, n, 11, frame_offset=12345, early_term=True, harvest=_lib.HARVEST_UNCOR_ANY, capacity=4000,
                                stage1_iters=s1)
        assert torch.equal(c1, c2), (s1, c1.tolist(), c2.tolist())
        k = int(u1.item())
        assert int(u2.item()) == k
        if 0 < k <= 4000:
            a = b1[:k].cpu().numpy(); b = b2[:k].cpu().numpy()
            assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])
    names = _lib.COUNTER_NAMES
    c = c1.cpu().numpy()
    assert c[names.index("frames")] == n
    s1 = montecarlo._pick_stage1(c, 20)
    assert s1 == 0 or 3 <= s1 <= 10


# ------------------------------------------------------------------------------------------------------------------
# persistent-slot Monte-Carlo kernel (csrc/nms_mcp.cuh)
def _decoder_for(codes, key, systematic, iters=20, weights=None):
    import ldpc_error_floor_b200 as L
    proto = codes[f"graph/{key}/proto"].astype(np.int32); meta = codes[f"graph/{key}/meta"]
    g = L.BaseGraph(proto, int(meta[0]), (int(meta[1]), int(meta[2])), (int(meta[3]), int(meta[4])))
    wk = {"wimax": "wimax_base20", "wifi": "wifi_boost50", "5g_r050_z64": "5g_r050_z64_boost50",
          "5g_r073_z32": "5g_r073_z32_boost50", "5g_r033_z32": "5g_r033_z32_boost50",
          "5g_r050_z32": "5g_r050_z32_boost50"}.get(key) if weights is None else None
    if wk:
        ws = L.WeightSet([int(v) for v in codes[f"weights/{wk}/sharing"]], {i: codes[f"weights/{wk}/block{i}"] for i in range(3)})
    else:
        ws = weights or L.WeightSet([3, 0, 0], {0: np.full((iters, 1), 0.8, np.float32)})
    return g, ws, L.NMSDecoder(g, ws, iters=iters, systematic=systematic)


def _sorted_rows(a):
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("key,snr,systematic,iters,n", [
    ("wimax", 3.0, 0, 20, 60001), ("wimax", 5.0, 0, 20, 200003), ("wimax", 2.0, 0, 20, 7), ("wimax", 2.0, 0, 20, 1),
    ("wimax", 3.0, 0, 5, 20011), ("wifi", 3.5, 0, 20, 50000), ("wifi", 3.5, 0, 50, 9000),
    ("5g_r050_z64", 2.5, 1, 20, 40000), ("5g_r050_z64", 1.5, 0, 50, 6000), ("5g_r073_z72", 4.5, 1, 20, 30011),
    ("5g_r073_z72", 3.0, 0, 20, 5000), ("5g_r073_z32", 3.0, 1, 20, 30000), ("5g_r033_z32", 0.5, 0, 20, 20000),
    ("5g_r050_z32", 1.5, 1, 30, 20000), ("mackay", 3.0, 0, 20, 100003), ("mackay", 6.0, 0, 20, 70), ("bch", 4.0, 0, 20, 50000)])
def test_persistent_kernel_counts_like_the_batch_kernel(key, snr, systematic, iters, n, codes, monkeypatch):
    """ldpc_mc_run with early termination: the persistent-slot kernel (slots refilled frame by frame, frames of one CTA
    at different iterations) against the batch kernel on the same global frame indices -- the eight counters and the
    harvested words, for every harvest criterion.  Covers shipped weights with per-iteration / per-check CN, UCN and VN
    rows, ragged and tiny launches, frames that use up all iterations, and the z = 1 graphs (64 slots per CTA: two-word
    slot masks, two bookkeeping warps)."""
    import torch
    from ldpc_error_floor_b200 import _lib
    g, ws, dec = _decoder_for(codes, key, systematic, iters)
    info = dec.mc_info()
    assert info["persistent"] and info["kernel"].startswith("nms_mcp_spec_"), info
    sigma = float(g.sigma([snr])[0])
    cap = 3000
    for harvest in (_lib.HARVEST_UNCOR_ANY, _lib.HARVEST_UNCOR_LAST, _lib.HARVEST_SYND_FAIL, _lib.HARVEST_NONE):
        monkeypatch.setenv("LDPC_B200_NO_PERSIST", "1")
        assert not dec.mc_info()["persistent"]
        c1, b1, u1 = dec.mc_run(sigma, n, 21, frame_offset=777, early_term=True, harvest=harvest, capacity=cap)
        torch.cuda.synchronize()
        monkeypatch.delenv("LDPC_B200_NO_PERSIST")
        c2, b2, u2 = dec.mc_run(sigma, n, 21, frame_offset=777, early_term=True, harvest=harvest, capacity=cap)
        torch.cuda.synchronize()
        assert torch.equal(c1, c2), (harvest, dict(zip(_lib.COUNTER_NAMES, zip(c1.tolist(), c2.tolist()))))
        k = int(u1.item())
        assert int(u2.item()) == k == int(c1[7].item())
        if harvest != _lib.HARVEST_NONE and 0 < k <= cap:
            assert np.array_equal(_sorted_rows(b1[:k].cpu().numpy()), _sorted_rows(b2[:k].cpu().numpy()))
    assert int(c1[0].item()) == n
    # accumulation over calls and offsets: two halves add up to the whole
    h = n // 2
    ca, _, _ = dec.mc_run(sigma, h, 21, frame_offset=777, early_term=True)
    ca, _, _ = dec.mc_run(sigma, n - h, 21, frame_offset=777 + h, early_term=True, counters=ca)
    assert torch.equal(ca, c2)


def _oracle_counters(app, synd, target_bits, early_term):
    """The eight Monte-Carlo counters from the oracle's per-iteration APPs [T, B, N*z] and syndromes [T, B] (True = some
    check violated), all-zero codeword, errors counted over the first target_bits bits (calc_ber_fer, Print_Functions.py:
    100-118, plus the termination-flag mapping of SURVEY.md 8a D9)."""
    T, B = synd.shape
    hard = app[:, :, :target_bits] >= 0
    ones = hard.sum(axis=2)                                  # [T, B]
    ok = ~synd
    stop = np.where(ok.any(axis=0), ok.argmax(axis=0), T - 1) if early_term else np.full(B, T - 1)   # iteration whose decision is output
    executed = np.where(ok.any(axis=0), ok.argmax(axis=0) + 1, T) if early_term else np.full(B, T)
    idx = np.arange(B)
    ones_out = ones[stop, idx]
    ever = np.array([(ones[:stop[b] + 1, b] == 0).any() for b in range(B)])
    synd_ok = ok[stop, idx]
    return {"frames": B, "frame_err_last": int((ones_out > 0).sum()), "frame_err_any": int((~ever).sum()),
            "bit_err_last": int(ones_out.sum()), "iters": int(executed.sum()), "synd_fail": int((~synd_ok).sum()),
            "undetected": int((synd_ok & (ones_out > 0)).sum())}


@pytest.mark.parametrize("key,snrs,systematic,iters,B", [("5g_r073_z72", (2.4, 2.7, 3.0, 3.5), 1, 20, 1600),
                                                         ("5g_r050_z64", (1.0, 1.5, 2.0), 1, 20, 1200),
                                                         ("wimax", (2.0, 2.5, 3.0, 3.5), 0, 20, 3000)])
def test_monte_carlo_counters_against_c_oracle(key, snrs, systematic, iters, B, codes, monkeypatch):
    """The campaign configurations end to end against oracle/nms_oracle.c: BASELINE config 5 exactly as the campaign runs it
    (5G R0.73 n2112 z72, plain 0.8 min-sum, sharing [3, 0, 0], 20 iterations, systematic = 1), config 4 and config 1.
    The fused Monte-Carlo launch (persistent-slot kernel AND batch kernel, with and without early termination) must report
    the counters computed from the oracle's per-iteration APPs of the very same generated frames; the decode entry point
    must return the oracle's APPs, flags and iteration counts on them (many CTAs, ragged tail)."""
    import torch
    from oracle import c_oracle
    from ldpc_error_floor_b200 import _lib
    g, ws, dec = _decoder_for(codes, key, systematic, iters)
    proto = codes[f"graph/{key}/proto"].astype(np.int32)
    tb = (g.N - g.M if systematic else g.N) * g.z
    for k, snr in enumerate(snrs):
        sigma, off = float(g.sigma([snr])[0]), 1000 * k + 17
        x = dec.generate(sigma, B, seed=31, frame_offset=off)
        ref = c_oracle.decode(proto, g.z, x.cpu().numpy(), ws.sharing, ws.blocks, iters, 2, 5, 20.0, want_all=True)
        r = dec.decode(x, app="last")
        assert np.array_equal(r.app.cpu().numpy(), ref["app_last"])
        assert np.array_equal((r.flags.cpu().numpy() & 1) != 0, ~ref["synd"][iters - 1])
        for et in (True, False):
            want = _oracle_counters(ref["app"], ref["synd"], tb, et)
            for persist in (True, False):
                if persist:
                    monkeypatch.delenv("LDPC_B200_NO_PERSIST", raising=False)
                else:
                    monkeypatch.setenv("LDPC_B200_NO_PERSIST", "1")
                cnt, _ = dec.mc_run_host(sigma, B, seed=31, frame_offset=off, early_term=et)
                got = {n: cnt[n] for n in want}
                assert got == want, (snr, et, persist, got, want)
        monkeypatch.delenv("LDPC_B200_NO_PERSIST", raising=False)
        assert want["frame_err_last"] < B       # the points are not degenerate
    # the float min-sum kernel of the same graph on the last point's frames (float z72 kernel: nms_f32_spec_5g_r073_z72)
    import ldpc_error_floor_b200 as L
    fdec = L.NMSDecoder(g, ws, iters=iters, decoding_type=1, systematic=systematic)
    xf = fdec.generate(sigma, min(B, 1500), seed=32, frame_offset=5)
    reff = c_oracle.decode(proto, g.z, xf.cpu().numpy(), ws.sharing, ws.blocks, iters, 1, 5, 20.0, want_all=False)
    rf = fdec.decode(xf, app="last")
    assert np.array_equal(rf.app.cpu().numpy(), reff["app_last"])
    assert np.array_equal((rf.flags.cpu().numpy() & 1) != 0, ~reff["synd"][iters - 1])
    e = fdec.decode(xf, early_term=True)
    okf = ~reff["synd"]
    assert np.array_equal(e.iters.cpu().numpy(), np.where(okf.any(axis=0), okf.argmax(axis=0) + 1, iters))


# ------------------------------------------------------------------------------------------------------------------
# reference-side Monte-Carlo fixtures (tests/golden/make_mc_fixtures.py): FER / BER inside the reference's 95 % interval
REF_MC = ["wimax", "wifi", "5g_r050_z64", "5g_r073_z72", "wimax_float"]


@pytest.mark.parametrize("name", REF_MC)
def test_fer_ber_inside_reference_intervals(name, codes):
    """north_star: "FER/BER at each Eb/N0 must fall inside the reference's 95 % confidence interval".  The reference side is
    Print_Functions.compute_results run unmodified (5 000 / 2 000 / 1 000 / 300 frames per point, mc_ref_<name>.npz); ours is
    2 M frames per point from the fused Monte-Carlo path, so its own sampling error is negligible.  FER and FER_last: Wilson
    interval of the reference's count.  BER: frames fail in bursts, so the interval comes from the reference's per-frame bit
    error counts (mean +- 1.96 standard errors), not from a per-bit binomial."""
    import ldpc_error_floor_b200 as L
    from ldpc_error_floor_b200.montecarlo import MonteCarlo, SnrPoint
    path = golden_path(f"mc_ref_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not minted")
    d = np.load(path)
    key, T, systematic = str(d["graph"]), int(d["T"]), int(d["systematic"])
    sharing = [int(v) for v in d["sharing"]]
    ws = L.WeightSet(sharing, {i: d[f"w{i}"] for i in range(3) if f"w{i}" in d})
    proto = codes[f"graph/{key}/proto"].astype(np.int32); meta = codes[f"graph/{key}/meta"]
    g = L.BaseGraph(proto, int(meta[0]), (int(meta[1]), int(meta[2])), (int(meta[3]), int(meta[4])))
    dec = L.NMSDecoder(g, ws, iters=T, decoding_type=int(d["decoding_type"]), q_bit=int(d["q_bit"]), systematic=systematic)
    mc = MonteCarlo(dec, seed=1234, chunk_frames=1 << 19)
    nref = int(d["frames"])
    for k, snr in enumerate(d["snr"]):
        assert float(g.sigma([snr])[0]) == pytest.approx(float(d["sigma"][k]), rel=1e-12)
        pt, _ = mc.run_point(float(snr), 1 << 21)                      # no early termination: FER_last needs all iterations
        ref = SnrPoint(float(snr), float(d["sigma"][k]), bits_per_frame=g.NZ)
        ref.add([nref, int(d[f"uncor_last_{k}"].sum()), int(d[f"uncor_any_{k}"].sum()), 0, 0, 0, 0, 0])
        assert ref.fer == pytest.approx(float(d["results"][2, k]), abs=1e-6)        # the fixture is self-consistent
        lo, hi = ref.fer_ci95("any")
        assert lo <= pt.fer <= hi, (name, snr, "FER", pt.fer, lo, hi)
        lo, hi = ref.fer_ci95("last")
        assert lo <= pt.fer_last <= hi, (name, snr, "FER_last", pt.fer_last, lo, hi)
        be = d[f"biterr_{k}"].astype(np.float64) / float(d["ber_divisor"])
        assert be.mean() == pytest.approx(float(d["results"][0, k]), rel=1e-4, abs=1e-9)
        half = 1.96 * be.std(ddof=1) / np.sqrt(nref)
        nfail = int((be > 0).sum())
        if nfail >= 30:
            lo_b, hi_b = be.mean() - half, be.mean() + half
        elif nfail > 0:
            # a handful of failing frames: BER = FER_last x mean burst size; Wilson interval for the first factor, a factor
            # of two either way for the second
            burst = be[be > 0].mean()
            wlo, whi = ref.fer_ci95("last")
            lo_b, hi_b = 0.5 * wlo * burst, 2.0 * whi * burst
        else:
            continue
        assert lo_b <= pt.ber_last <= hi_b, (name, snr, "BER", pt.ber_last, lo_b, hi_b)


# ------------------------------------------------------------------------------------------------------------------
# generator: the distribution the error-floor estimates rest on
def test_generator_tails_and_ks():
    """N(0,1) stream of the fused generator (Philox4x32-10 + Box-Muller on MUFU log2 / rsqrt / sin / cos, u1 refined with a
    second Philox block below 2^-24): (i) Kolmogorov-Smirnov against the normal CDF on 1e7 samples, (ii) two-sided tail
    counts beyond 3, 4, 5, 6 and 7 sigma over 1.07e11 samples against erfc within a 4-sigma Poisson band -- the reference
    draws float64 normals from numpy (Print_Functions.py:45), so this is what "statistically equivalent" has to mean for a
    1e-9 event."""
    import math
    import torch
    from scipy import stats
    from ldpc_error_floor_b200.decoder import normal_probe
    _, x = normal_probe(seed=99, n_frames=10_000, quads_per_frame=250, want_normals=True)
    x = x.cpu().numpy().astype(np.float64)
    assert x.size == 10_000_000
    ks = stats.kstest(x, "norm")
    assert ks.statistic < 1.63 / math.sqrt(x.size), ks            # 1 % critical value of the KS statistic
    assert abs(x.mean()) < 4 / math.sqrt(x.size) and abs(x.var() - 1) < 4 * math.sqrt(2 / x.size)
    # pairs (cos, sin branch) and neighbouring samples are uncorrelated
    assert abs(np.corrcoef(x[0::2], x[1::2])[0, 1]) < 4 / math.sqrt(x.size / 2)
    counts = None
    n_frames, quads = 1 << 20, 1 << 8                             # 2^28 Philox blocks = 1.07e9 samples per launch
    for rep in range(100):
        counts, _ = normal_probe(seed=7, n_frames=n_frames, quads_per_frame=quads, frame_offset=rep * n_frames, counts=counts)
    c = counts.cpu().numpy()
    total = int(c[5])
    assert total == 100 * n_frames * quads * 4
    for k, thr in enumerate((3.0, 4.0, 5.0, 6.0, 7.0)):
        expect = total * math.erfc(thr / math.sqrt(2.0))
        band = 4.0 * math.sqrt(expect) + 1.0
        assert abs(int(c[k]) - expect) <= band + 2e-4 * expect, (thr, int(c[k]), expect)     # 2e-4: MUFU accuracy at 3-4 sigma
    print("tail counts", dict(zip((3, 4, 5, 6, 7), c[:5].tolist())), "of", total)


def test_create_mix_epoch_and_compute_results_with_codewords():
    """create_mix_epoch drop-in with a generator matrix (is_zeros_word = False) and compute_results(sampling_type 1) on labelled
    words: Y is a codeword, the decoder brings most frames back to it, and the metrics are taken against it."""
    import torch
    from ldpc_error_floor_b200 import formats
    from ldpc_error_floor_b200.montecarlo import compute_results, create_mix_epoch
    d = np.load(golden_path("decode_mackay_qms_300_t20_cw.npz"))
    case, g, dec = make("mackay_qms_300_t20_cw")
    H = np.zeros((g.M, g.N), np.int64); H[case["proto"] != -1] = 1         # z = 1: the proto matrix is H
    Yg = d["codeword"].astype(np.int64)
    assert not ((Yg @ H.T) % 2).any()
    GM = Yg[:8]                                                            # eight codewords span a small subcode
    rs = np.random.RandomState(4)
    X, Y = create_mix_epoch(dec, g.sigma([4.0, 5.0]), rs, seed=12, batch_size=600, code_GM=GM, is_zeros_word=False)
    assert X.shape == (600, g.N, g.z) and Y.shape == (600, g.NZ) and not ((Y @ H.T) % 2).any() and Y.any()
    r, signed, cnt = dec.decode_cw(X, Y)
    c = cnt.cpu().numpy()
    assert c[0] == 600 and c[2] <= c[1] < 60                               # 4-5 dB: nearly every frame is brought back to Y
    z, _, _ = dec.decode_cw(X, np.zeros_like(Y))                           # against the wrong word: (almost) all frames "fail"
    assert int(((z.flags.cpu().numpy() & 4) != 0).sum()) > 500
    rows = -X.reshape(600, -1).cpu().numpy()                               # the [Uncor] file sign convention
    res, _ = compute_results(dec, 600, rows, np.array([0.0]), 20, 1, input_codeword=Y)
    assert res[1, 0] == pytest.approx(c[1] / 600) and res[2, 0] == pytest.approx(c[2] / 600)
    Xz, Yz = create_mix_epoch(dec, g.sigma([4.0]), np.random.RandomState(4), seed=12, batch_size=64)
    assert not Yz.any() and torch.equal(Xz, dec.generate(float(g.sigma([4.0])[0]), 64, 12).reshape(64, g.N, g.z))
