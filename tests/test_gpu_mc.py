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
def test_two_stage_monte_carlo_equals_single_stage(key, snr, systematic, codes):
    """ldpc_mc_run_staged: stage 1 defers the frames without a zero syndrome by frame index, stage 2 regenerates and
    decodes them in full -- the eight counters and the set of harvested words are those of the one-launch run."""
    import torch
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
