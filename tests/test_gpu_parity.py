"""GPU parity: the CUDA path (through the C-ABI) against the golden vectors minted from the
reference's own code, and against the C oracle on larger seeded batches.

Pass criteria (BASELINE.json north_star): quantised path bit-exact (APP, hard decisions, syndrome
flags); float path: hard decisions identical, APP within 1e-5 relative."""
import os

import numpy as np
import pytest

from conftest import all_cases, load_case

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5   # float min-sum: |a-b| <= REL_TOL * max(1, |b|)


def build_decoder(case, iters=None):
    import ldpc_error_floor_b200 as L
    g = L.BaseGraph(case["proto"], case["z"], case["punct"], case["short"])
    ws = L.WeightSet(case["sharing"], {i: w for i, w in case["weights"].items()})
    dec = L.NMSDecoder(g, ws, iters=case["T"] if iters is None else iters, decoding_type=case["decoding_type"],
                       q_bit=case["q_bit"], clip_llr=case["clip"])
    return g, dec


def oracle_flags(g, app):
    """app [T,B,NZ] -> per-frame reference observables (Print_Functions.py:100-118 + syndrome)."""
    hard = app >= 0
    T, B, _ = app.shape
    synd_ok = np.stack([g.syndrome_ok(hard[t]) for t in range(T)])          # [T,B]
    any_one = hard.any(axis=2)                                              # [T,B]
    uncor_any = any_one.all(axis=0)
    iters = np.where(synd_ok.any(axis=0), synd_ok.argmax(axis=0) + 1, T)
    return hard, synd_ok, any_one, uncor_any, iters


@pytest.mark.parametrize("name", all_cases())
def test_golden_decode(name):
    import torch
    case = load_case(name)
    g, dec = build_decoder(case)
    xa = torch.from_numpy(case["xa"]).cuda()
    r = dec.decode(xa, app="all", unpack=True)
    app = r.app.cpu().numpy()
    ref = case["app"]
    assert app.shape == ref.shape
    hard_ref, synd_ok, any_one, uncor_any, iters = oracle_flags(g, ref)
    if case["decoding_type"] == 2:
        assert np.array_equal(app, ref), f"max |diff| {np.abs(app - ref).max()}"
    else:
        err = np.abs(app - ref) / np.maximum(1.0, np.abs(ref))
        assert err.max() <= REL_TOL, f"max rel diff {err.max()}"
        assert np.array_equal(app >= 0, hard_ref)
    T = case["T"]
    assert np.array_equal(r.hard.cpu().numpy().astype(bool), hard_ref[T - 1])
    flags = r.flags.cpu().numpy()
    assert np.array_equal((flags & 1) != 0, synd_ok[T - 1])
    assert np.array_equal((flags & 2) != 0, uncor_any)
    assert np.array_equal((flags & 4) != 0, any_one[T - 1])
    assert np.array_equal((flags & 8) != 0, synd_ok.any(axis=0))
    assert np.array_equal(r.iters.cpu().numpy(), iters)
    assert np.array_equal(r.biterr.cpu().numpy(), hard_ref[T - 1].sum(axis=1))
    # ya_output_all layout == concat over iterations (Main_Functions.py:380-383)
    assert np.array_equal(dec.ya_output_all(xa).cpu().numpy(), app.reshape(-1, app.shape[-1]))


@pytest.mark.parametrize("name", all_cases())
def test_golden_fast_path(name):
    """Without APP output the graph-specialised kernels run their unrolled code (with it, the table-driven code): the hard
    decisions, flags, first zero-syndrome iteration and bit-error counts of that path against the reference, with and
    without early termination."""
    import torch
    case = load_case(name)
    g, dec = build_decoder(case)
    xa = torch.from_numpy(case["xa"]).cuda()
    hard_ref, synd_ok, any_one, uncor_any, iters = oracle_flags(g, case["app"])
    T = case["T"]
    r = dec.decode(xa, app=None, unpack=True)
    assert np.array_equal(r.hard.cpu().numpy().astype(bool), hard_ref[T - 1])
    flags = r.flags.cpu().numpy()
    assert np.array_equal((flags & 1) != 0, synd_ok[T - 1])
    assert np.array_equal((flags & 2) != 0, uncor_any)
    assert np.array_equal((flags & 4) != 0, any_one[T - 1])
    assert np.array_equal(r.iters.cpu().numpy(), iters)
    assert np.array_equal(r.biterr.cpu().numpy(), hard_ref[T - 1].sum(axis=1))
    e = dec.decode(xa, early_term=True, app=None, unpack=True)
    stop = iters - 1
    B = hard_ref.shape[1]
    assert np.array_equal(e.iters.cpu().numpy(), iters)
    assert np.array_equal(e.hard.cpu().numpy().astype(bool), np.stack([hard_ref[stop[b], b] for b in range(B)]))
    assert np.array_equal((e.flags.cpu().numpy() & 1) != 0, np.array([synd_ok[stop[b], b] for b in range(B)]))


@pytest.mark.parametrize("name", ["wimax_qms_333_t20", "5g_r073_z32_qms_222_t50", "wimax_float_333_t20",
                                  "mackay_qms_300_t20"])
def test_golden_early_termination(name):
    """With early termination a frame stops at the first zero syndrome: its outputs must equal the
    reference's at that iteration; frames that never converge match the last iteration."""
    import torch
    case = load_case(name)
    g, dec = build_decoder(case)
    xa = torch.from_numpy(case["xa"]).cuda()
    ref = case["app"]
    hard_ref, synd_ok, any_one, uncor_any, iters = oracle_flags(g, ref)
    r = dec.decode(xa, early_term=True, app="last", unpack=True)
    T = case["T"]
    stop = iters - 1   # iteration index whose decision is output
    B = ref.shape[1]
    want_hard = np.stack([hard_ref[stop[b], b] for b in range(B)])
    assert np.array_equal(r.iters.cpu().numpy(), iters)
    assert np.array_equal(r.hard.cpu().numpy().astype(bool), want_hard)
    app = r.app.cpu().numpy()
    want_app = np.stack([ref[stop[b], b] for b in range(B)])
    if case["decoding_type"] == 2:
        assert np.array_equal(app, want_app)
    else:
        assert (np.abs(app - want_app) / np.maximum(1, np.abs(want_app))).max() <= REL_TOL
    flags = r.flags.cpu().numpy()
    assert np.array_equal((flags & 1) != 0, np.array([synd_ok[stop[b], b] for b in range(B)]))
    assert np.array_equal((flags & 4) != 0, np.array([any_one[stop[b], b] for b in range(B)]))


@pytest.mark.parametrize("name,B", [("wimax_qms_333_t20", 3000), ("wifi_qms_333_t50", 700),
                                    ("5g_r050_z64_qms_222_t50", 500), ("wimax_float_333_t20", 1500),
                                    ("bch_qms_333_t10", 900), ("mackay_float_300_t20", 2500)])
def test_against_c_oracle_large(name, B):
    """Seeded batches big enough to exercise many CTAs, ragged tails and the persistent loop,
    checked against oracle/nms_oracle.c (itself pinned to the goldens)."""
    import torch
    from oracle import c_oracle
    case = load_case(name)
    g, dec = build_decoder(case)
    rng = np.random.RandomState(1234)
    sig = float(np.mean(g.sigma([3.0])))
    xa = (2.0 * (rng.normal(size=(B, g.N, g.z)) * sig - 1.0) / sig ** 2).astype(np.float32)
    if case["decoding_type"] == 2:
        xa = np.clip(np.rint(xa * 2) / 2, -7.5, 7.5).astype(np.float32)
    if case["punct"][0] > 0:
        xa.reshape(B, -1)[:, case["punct"][0] - 1:case["punct"][1]] = 0
    if case["short"][0] > 0:
        xa.reshape(B, -1)[:, case["short"][0] - 1:case["short"][1]] = -case["clip"]
    T = case["T"]
    ref = c_oracle.decode(case["proto"], case["z"], xa, case["sharing"], case["weights"], T, case["decoding_type"],
                          case["q_bit"], case["clip"], want_all=False)
    r = dec.decode(torch.from_numpy(xa).cuda(), app="last", unpack=True)
    app = r.app.cpu().numpy()
    # the float kernels sum in the oracle's order (nms_f32.cuh), so the float path is compared bit for bit as well
    assert np.array_equal(app, ref["app_last"]), f"max |diff| {np.abs(app - ref['app_last']).max()}"
    assert np.array_equal((r.flags.cpu().numpy() & 1) != 0, ~ref["synd"][T - 1])
    synd_ok = ~ref["synd"]
    iters = np.where(synd_ok.any(axis=0), synd_ok.argmax(axis=0) + 1, T)
    assert np.array_equal(r.iters.cpu().numpy(), iters)
    # host-buffer entry point gives the same answer
    h = dec.decode_host(xa, app=None)
    assert np.array_equal(h["hard_packed"].view(np.int32), r.hard_packed.cpu().numpy())
    assert np.array_equal(h["flags"], r.flags.cpu().numpy())


@pytest.mark.parametrize("name", ["wimax_qms_333_t20", "5g_r050_z64_qms_222_t50", "mackay_qms_300_t20", "polar_qms_223_t6"])
def test_q8_words_decode_like_float_words(name):
    """ldpc_decode_q8 / ldpc_decode_q8_host on int8 words == ldpc_decode on the same words as float32."""
    import torch
    import ldpc_error_floor_b200 as L
    from ldpc_error_floor_b200 import formats as F
    case = load_case(name)
    g = L.BaseGraph(case["proto"], case["z"], case["punct"], case["short"])
    dec = L.NMSDecoder(g, L.WeightSet(case["sharing"], dict(case["weights"])), iters=case["T"], decoding_type=2,
                       q_bit=case["q_bit"], clip_llr=case["clip"])
    assert dec.q8_step == 0.5
    x = dec.generate(float(g.sigma([2.5])[0]), 1500, seed=11).reshape(1500, -1)
    x = torch.clamp(x, -7.5, 7.5)                         # shortened bits are -clip_LLR = -20: Q() clamps them anyway
    words = torch.from_numpy(F.llr_to_q8(x.cpu().numpy(), 0.5)).cuda()
    for et in (False, True):
        a = dec.decode(x, early_term=et)
        cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
        b = dec.decode_q8(words, early_term=et, counters=cnt)
        h = dec.decode_q8_host(words.cpu().numpy(), early_term=et)
        for r in (b,):
            assert torch.equal(a.hard_packed, r.hard_packed) and torch.equal(a.iters, r.iters)
            assert torch.equal(a.flags, r.flags) and torch.equal(a.biterr, r.biterr)
        assert np.array_equal(h["hard_packed"].view(np.int32), a.hard_packed.cpu().numpy())
        assert np.array_equal(h["flags"], a.flags.cpu().numpy()) and np.array_equal(h["iters"], a.iters.cpu().numpy())
        c = cnt.cpu().numpy()
        assert c[0] == 1500 and c[3] == int(a.biterr.sum().item()) and c[2] == int(((a.flags & 2) != 0).sum().item())


def test_systematic_metrics_match_reference():
    """systematic = 1 (main_Base.py:29, 83-86): ya_output_all and calc_ber_fer only see the first N - M columns.
    Golden: the reference's own calc_ber_fer on its ya_output_all (decode_5g_r050_z64_qms_222_t12_sys.npz)."""
    import torch
    import ldpc_error_floor_b200 as L
    case = load_case("5g_r050_z64_qms_222_t12_sys")
    g = L.BaseGraph(case["proto"], case["z"], case["punct"], case["short"])
    ws = L.WeightSet(case["sharing"], dict(case["weights"]))
    dec = L.NMSDecoder(g, ws, iters=case["T"], decoding_type=2, q_bit=5, clip_llr=case["clip"], systematic=1)
    assert dec.target_node == case["target_node"] == g.N - g.M
    xa = torch.from_numpy(case["xa"]).cuda()
    r = dec.decode(xa, app="all", unpack=True)
    B, tz = xa.shape[0], case["target_node"] * case["z"]
    # ya_output_all = per-iteration APPs truncated to the target columns, iterations stacked on axis 0 (:329-335, 380-383)
    ya = r.app[:, :, :tz].reshape(-1, tz).cpu().numpy()
    assert np.array_equal(ya, case["ya_output_all"])
    assert np.array_equal(r.uncor_any.cpu().numpy(), case["uncor_flag"] > 0)
    assert np.array_equal(r.biterr.cpu().numpy(), case["error_num"])
    full = L.NMSDecoder(g, ws, iters=case["T"], decoding_type=2, q_bit=5, clip_llr=case["clip"]).decode(xa, unpack=True)
    assert torch.equal(full.hard, r.hard) and torch.equal(full.iters, r.iters)       # outputs cover the whole word
    assert int(full.biterr.sum()) >= int(r.biterr.sum()) and int(full.uncor_any.sum()) >= int(r.uncor_any.sum())
    assert np.array_equal(r.hard[:, :tz].sum(dim=1).cpu().numpy(), case["error_num"])
    # same through the fused Monte-Carlo counters and with early termination on device buffers
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    dec.decode_q8(torch.round(torch.clamp(xa.reshape(B, -1), -7.5, 7.5) * 2).to(torch.int8), counters=cnt)
    c = cnt.cpu().numpy()
    assert c[0] == B and c[2] == int((case["uncor_flag"] > 0).sum()) and c[3] == int(case["error_num"].sum())
    ber_last, fer_last, fer = case["metrics"]
    assert c[1] / B == pytest.approx(fer_last) and c[2] / B == pytest.approx(fer)
    assert c[3] / (B * g.NZ) == pytest.approx(ber_last)          # the reference divides by the FULL length (:113)


def test_temporal_sharing_matches_reference():
    """sharing code 4: fixed_iter + 1 per-edge variables, iterations >= fixed_iter reuse the last one
    (Main_Functions.py:299-304).  Golden: the reference run with sharing [4,0,2], fixed_iter = 3, T = 8."""
    import torch
    import ldpc_error_floor_b200 as L
    case = load_case("wimax_qms_402_t8_fixed3")
    assert case["raw_sharing"] == [4, 0, 2] and case["raw_weights"][0].shape[0] == 4 and case["sharing"] == [1, 0, 2]
    g = L.BaseGraph(case["proto"], case["z"])
    raw = L.WeightSet(case["raw_sharing"], dict(case["raw_weights"]))
    with pytest.raises(ValueError):
        L.NMSDecoder(g, raw)                                       # iteration count is not in the variables
    dec = L.NMSDecoder(g, raw, iters=8, fixed_iter=3, decoding_type=2, q_bit=5)
    r = dec.decode(torch.from_numpy(case["xa"]).cuda(), app="all")
    assert np.array_equal(r.app.cpu().numpy(), case["app"])


def _quirk_inputs(g, B, seed, qms):
    """Noisy words with the values the reference treats specially sprinkled in: exact zeros and -0.0 (the zero rule,
    Main_Functions.py:230), magnitudes at and below 1e-4 (the minimum rule, :250), values beyond clip_LLR."""
    rng = np.random.RandomState(seed)
    sig = float(np.mean(g.sigma([3.0])))
    xa = (2.0 * (rng.normal(size=(B, g.N * g.z)) * sig - 1.0) / sig ** 2).astype(np.float32)
    special = np.array([0.0, -0.0, 5e-5, -5e-5, 1e-4, -1e-4, 9.9e-5, 30.0, -30.0, 20.0, -20.0], dtype=np.float32)
    pick = rng.rand(B, g.N * g.z) < 0.08
    xa[pick] = special[rng.randint(0, len(special), size=int(pick.sum()))]
    if qms:
        xa = np.clip(np.rint(xa * 2) / 2, -7.5, 7.5).astype(np.float32)
    return xa


@pytest.mark.parametrize("name,exact", [("wimax_float_333_t20", True), ("mackay_float_300_t20", True),
                                        ("5g_r073_z32_float_222_t50", True), ("wimax_qms_q6_323_t6", True),
                                        ("wimax_qms_112_t6", True), ("wimax_qms_333_t20", True)])
def test_quirk_values_against_c_oracle(name, exact):
    """Zeros, -0.0, |x| <= 1e-4 and |x| > clip_LLR in the channel values exercise the reference's 1e-4 rules; the float
    kernels apply those per check row instead of per edge and sum in the oracle's order, so even the float path is
    compared bit for bit here (all iterations, APP-output path) and by decisions / flags on the unrolled path."""
    import torch
    from oracle import c_oracle
    case = load_case(name)
    g, dec = build_decoder(case)
    T = case["T"]
    xa = _quirk_inputs(g, 600, 77, case["decoding_type"] == 2)
    ref = c_oracle.decode(case["proto"], case["z"], xa.reshape(600, g.N, g.z), case["sharing"], case["weights"], T,
                          case["decoding_type"], case["q_bit"], case["clip"], want_all=True)
    r = dec.decode(torch.from_numpy(xa).cuda(), app="all", unpack=True)
    app = r.app.cpu().numpy()
    if exact:
        assert np.array_equal(app, ref["app"]), f"max |diff| {np.abs(app - ref['app']).max()}"
    else:
        assert (np.abs(app - ref["app"]) / np.maximum(1, np.abs(ref["app"]))).max() <= REL_TOL
    synd_ok = ~ref["synd"]
    iters = np.where(synd_ok.any(axis=0), synd_ok.argmax(axis=0) + 1, T)
    f = dec.decode(torch.from_numpy(xa).cuda(), app=None, unpack=True)      # unrolled path of the specialised kernels
    assert np.array_equal(f.hard.cpu().numpy().astype(bool), ref["app"][T - 1] >= 0)
    assert np.array_equal(f.iters.cpu().numpy(), iters)
    assert np.array_equal((f.flags.cpu().numpy() & 1) != 0, synd_ok[T - 1])


@pytest.mark.parametrize("dt", [2, 1])
def test_full_size_batch_properties(dt):
    """BASELINE-size launch (2^20 WiMAX frames, every CTA slot and the persistent loop busy): copies of the same word
    decode to the same bits wherever they sit in the batch, a second run reproduces the first, the counters equal the
    sums of the per-frame outputs, and a permutation of the frames permutes the results."""
    import torch
    import ldpc_error_floor_b200 as L
    case = load_case("wimax_qms_333_t20")
    g = L.BaseGraph(case["proto"], case["z"], case["punct"], case["short"])
    dec = L.NMSDecoder(g, L.WeightSet(case["sharing"], dict(case["weights"])), iters=20, decoding_type=dt, q_bit=5)
    W, B = 4099, 1 << 20                                    # a prime number of distinct words: copies land in every slot
    words = dec.generate(float(g.sigma([3.0])[0]), W, seed=5).reshape(W, -1)
    idx = torch.arange(B, device="cuda") % W
    llr = words[idx].contiguous()
    a = dec.decode(llr)
    cnt, pd = dec.post_decode(llr)                          # same launch with on-device counters
    assert torch.equal(pd.hard_packed, a.hard_packed) and torch.equal(pd.flags, a.flags) and torch.equal(pd.iters, a.iters)
    base = dec.decode(words)
    for field in ("hard_packed", "iters", "flags", "biterr"):
        assert torch.equal(getattr(a, field), getattr(base, field)[idx]), field
    b = dec.decode(llr)
    assert torch.equal(a.hard_packed, b.hard_packed) and torch.equal(a.flags, b.flags)
    c = cnt.cpu().numpy()
    assert c[0] == B and c[3] == int(a.biterr.sum().item())
    assert c[1] == int(((a.flags & 4) != 0).sum().item()) and c[2] == int(((a.flags & 2) != 0).sum().item())
    perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    p = dec.decode(llr[perm].contiguous(), early_term=True)
    e = dec.decode(llr, early_term=True)
    assert torch.equal(p.hard_packed, e.hard_packed[perm]) and torch.equal(p.iters, e.iters[perm])


@pytest.mark.parametrize("name", all_cases(sum_product=True))
def test_sum_product_golden(name):
    """decoding_type 0 (sum-product, Main_Functions.py:238-245) against the reference's own run (numpy tanh / arctanh).
    Stated criterion: x0 = -2 atanh(prod tanh(-v/2)) has derivative 2 / (1 - x^2), so once messages saturate (|x0| > ~10,
    1 - |x| < 1e-4) one unit in the last place of the product -- which is all that two correct float32 tanh
    implementations can promise each other -- moves the output by up to several percent; TensorFlow's own Eigen tanh and
    the numpy stand-in differ in the same way.  Hence: (1) hard decisions, flags and iteration counts identical on every
    frame and iteration; (2) the iterations before saturation (t <= 2 here) within the 1e-5 bar of the other paths;
    (3) at least 95 % of all APP values within 1e-5 relative; (4) every value within 10 %."""
    import torch
    case = load_case(name)
    g, dec = build_decoder(case)
    assert case["decoding_type"] == 0
    xa = torch.from_numpy(case["xa"]).cuda()
    r = dec.decode(xa, app="all", unpack=True)
    app, ref = r.app.cpu().numpy(), case["app"]
    err = np.abs(app - ref) / np.maximum(1.0, np.abs(ref))
    assert np.array_equal(app >= 0, ref >= 0)
    assert err[:3].max() <= REL_TOL, err[:3].max()
    assert (err <= REL_TOL).mean() >= 0.95 and err.max() <= 0.1, ((err <= REL_TOL).mean(), err.max())
    hard_ref, synd_ok, any_one, uncor_any, iters = oracle_flags(g, ref)
    T = case["T"]
    assert np.array_equal(r.iters.cpu().numpy(), iters)
    flags = r.flags.cpu().numpy()
    assert np.array_equal((flags & 1) != 0, synd_ok[T - 1]) and np.array_equal((flags & 2) != 0, uncor_any)
    f = dec.decode(xa, app=None, early_term=True, unpack=True)      # unrolled VN phase + early termination
    stop = iters - 1
    assert np.array_equal(f.iters.cpu().numpy(), iters)
    assert np.array_equal(f.hard.cpu().numpy().astype(bool), np.stack([hard_ref[stop[b], b] for b in range(ref.shape[1])]))


def test_sum_product_monte_carlo_and_limits():
    """The fused generator gives punctured bits the reference's 0.001 in sum-product mode (Print_Functions.py:53-55),
    sum-product beats min-sum at the same Eb/N0, and the training kernel refuses decoding_type 0 with an error code."""
    import torch
    import ldpc_error_floor_b200 as L
    from ldpc_error_floor_b200 import _lib
    case = load_case("5g_r073_z32_sp_222_t12", )
    g = L.BaseGraph(case["proto"], case["z"], case["punct"], case["short"])
    ws = L.WeightSet([3, 0, 0], {0: np.ones((12, 1), np.float32)})
    sp = L.NMSDecoder(g, ws, iters=12, decoding_type=0)
    ms = L.NMSDecoder(g, ws, iters=12, decoding_type=1)
    x = sp.generate(float(g.sigma([3.0])[0]), 64, seed=3).reshape(64, -1).cpu().numpy()
    ps, pe = case["punct"]
    assert np.all(x[:, ps - 1:pe] == np.float32(0.001))
    assert np.all(ms.generate(float(g.sigma([3.0])[0]), 64, seed=3).reshape(64, -1).cpu().numpy()[:, ps - 1:pe] == 0)
    sigma = float(g.sigma([2.5])[0])
    c_sp, _, _ = sp.mc_run(sigma, 1 << 16, 5)
    c_ms, _, _ = ms.mc_run(sigma, 1 << 16, 5)
    c_sp, c_ms = c_sp.cpu().numpy(), c_ms.cpu().numpy()
    assert c_sp[0] == c_ms[0] == 1 << 16 and 0 < c_sp[2] < c_ms[2]
    with pytest.raises(_lib.LdpcError):
        sp.train_grad(torch.zeros((4, g.NZ), dtype=torch.float32, device="cuda"))


@pytest.mark.parametrize("env", [("LDPC_B200_NO_SPEC",), ("LDPC_B200_NO_SPEC", "LDPC_B200_FORCE_GENERIC")])
@pytest.mark.parametrize("name", ["wimax_qms_333_t20", "wimax_float_333_t20", "wimax_qms_q6_323_t6", "wimax_qms_112_t6",
                                  "5g_r073_z32_float_222_t50", "5g_r050_z64_qms_222_t50", "mackay_float_300_t20"])
def test_generic_kernels_give_the_same_results(name, env, monkeypatch):
    """A graph that is not known at build time is served by the degree-bucketed generic kernels (table-driven code,
    same arithmetic headers).  Forced here for shipped graphs: goldens, and bit-identity with the specialised kernels."""
    import torch
    case = load_case(name)
    xa = torch.from_numpy(case["xa"]).cuda()
    _, spec = build_decoder(case)
    a = spec.decode(xa, app="all", early_term=False)
    for e in env:
        monkeypatch.setenv(e, "1")
    g, gen = build_decoder(case)
    assert "_kernel_" in gen.kernel_name and "_spec_" in spec.kernel_name, (gen.kernel_name, spec.kernel_name)
    if "LDPC_B200_FORCE_GENERIC" in env:
        assert gen.kernel_name.endswith("_0_0")
    b = gen.decode(xa, app="all", early_term=False)
    assert torch.equal(a.app, b.app) and torch.equal(a.hard_packed, b.hard_packed)
    assert torch.equal(a.flags, b.flags) and torch.equal(a.iters, b.iters) and torch.equal(a.biterr, b.biterr)
    ref = case["app"]
    app = b.app.cpu().numpy()
    if case["decoding_type"] == 2:
        assert np.array_equal(app, ref)
    else:
        assert (np.abs(app - ref) / np.maximum(1.0, np.abs(ref))).max() <= REL_TOL and np.array_equal(app >= 0, ref >= 0)
    e1, e2 = spec.decode(xa, early_term=True), gen.decode(xa, early_term=True)
    assert torch.equal(e1.hard_packed, e2.hard_packed) and torch.equal(e1.iters, e2.iters) and torch.equal(e1.flags, e2.flags)


def test_fixed_iteration_geometry_gives_the_same_results(monkeypatch):
    """WiMAX carries a second launch geometry (16 frames per CTA) that serves the calls WITHOUT early termination; calls
    with early termination keep the default (8 frames per CTA).  Same bits from both, and from a decoder without it."""
    import torch
    case = load_case("wimax_qms_333_t20")
    g, dec = build_decoder(case)
    a, b = dec.launch_info(early_term=False), dec.launch_info(early_term=True)
    assert a["kernel"] == "nms_h2_spec_wimax_fp8_r2" and a["frames_per_cta"] == 16
    assert b["kernel"] == "nms_h2_spec_wimax_fp4_r2" and b["frames_per_cta"] == 8
    x = dec.generate(float(g.sigma([3.0])[0]), 5000, seed=9).reshape(5000, -1)
    r1 = dec.decode(x)                                   # fixed iterations: the 16-frame geometry
    cnt1, _ = dec.post_decode(x)
    monkeypatch.setenv("LDPC_B200_NO_ALT", "1")
    _, plain = build_decoder(case)
    assert plain.launch_info(early_term=False)["kernel"] == "nms_h2_spec_wimax_fp4_r2"
    r2 = plain.decode(x)
    cnt2, _ = plain.post_decode(x)
    for f in ("hard_packed", "iters", "flags", "biterr"):
        assert torch.equal(getattr(r1, f), getattr(r2, f)), f
    assert torch.equal(cnt1, cnt2)
    c1, _, _ = dec.mc_run(float(g.sigma([3.0])[0]), 40000, 4, early_term=False)
    c2, _, _ = plain.mc_run(float(g.sigma([3.0])[0]), 40000, 4, early_term=False)
    assert torch.equal(c1, c2)


@pytest.mark.parametrize("name", ["wimax_qms_333_t20", "wimax_float_333_t20", "5g_r050_z64_qms_222_t50"])
def test_partial_iterations_and_ragged_batches(name):
    """iters = t + 1 gives ya_output{t} (main_Base.py:160); an empty batch is a no-op; batches of 1 .. 2 x frames-per-CTA + 1
    frames (ragged last CTA, both launch geometries) reproduce the rows of the full batch."""
    import torch
    case = load_case(name)
    g, dec = build_decoder(case)
    xa = torch.from_numpy(case["xa"]).cuda()
    ref = case["app"]
    for k in (1, 2, case["T"] // 2, case["T"] - 1):
        r = dec.decode(xa, iters=k, app="last")
        a = r.app.cpu().numpy()
        if case["decoding_type"] == 2:
            assert np.array_equal(a, ref[k - 1]), k
        else:
            assert (np.abs(a - ref[k - 1]) / np.maximum(1, np.abs(ref[k - 1]))).max() <= REL_TOL
    empty = dec.decode(xa[:0])
    assert empty.hard_packed.shape[0] == 0 and empty.flags.shape[0] == 0
    big = torch.cat([xa] * 6)[:37]
    full = dec.decode(big)
    full_et = dec.decode(big, early_term=True)
    for n in (1, 2, 3, 7, 8, 9, 15, 16, 17, 33, 37):
        for et, want in ((False, full), (True, full_et)):
            r = dec.decode(big[:n].contiguous(), early_term=et)
            assert torch.equal(r.hard_packed, want.hard_packed[:n]) and torch.equal(r.iters, want.iters[:n])
            assert torch.equal(r.flags, want.flags[:n]) and torch.equal(r.biterr, want.biterr[:n])
    h = dec.decode_host(big[:5].cpu().numpy())
    assert np.array_equal(h["flags"], full.flags[:5].cpu().numpy())


# ------------------------------------------------------------------------------------------------------------------
# "next" row N3: non-zero codewords (Print_Functions.py:40-46 with is_zeros_word = False, metrics against Y :100-118)
@pytest.mark.parametrize("name", ["mackay_qms_300_t20_cw", "wimax_qms_333_t20_cw", "wimax_float_333_t10_cw"])
def test_nonzero_codeword_golden(name):
    """Goldens minted by the reference itself: create_mix_epoch with a generator matrix, build_neural_network,
    calc_ber_fer(ya_output_all, T, Y, B).  ldpc_decode_cw must report the reference's uncor_flag (never equal to Y at any
    iteration), its per-frame signed error sums, FER_last and -- with the reference's cancelling sum -- BER_last."""
    import torch
    from ldpc_error_floor_b200 import _lib
    case = load_case(name)
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", f"decode_{name}.npz"))
    Y, T = d["codeword"], case["T"]
    g, dec = build_decoder(case)
    xa = torch.from_numpy(case["xa"]).cuda()
    B = xa.shape[0]
    r, signed, cnt = dec.decode_cw(xa, Y)
    flags = r.flags.cpu().numpy()
    hard_ref = case["app"] >= 0                                           # [T, B, NZ]
    wrong = (hard_ref != Y[None].astype(bool)).any(axis=2)                # [T, B]
    assert np.array_equal((flags & _lib.FLAG_UNCOR_ANY) != 0, d["uncor_flag"].astype(bool))
    assert np.array_equal((flags & _lib.FLAG_UNCOR_ANY) != 0, wrong.all(axis=0))
    assert np.array_equal((flags & _lib.FLAG_UNCOR_LAST) != 0, wrong[-1])
    assert np.array_equal(signed.cpu().numpy(), d["error_num"].astype(np.int64))
    assert np.array_equal(r.biterr.cpu().numpy(), (hard_ref[-1] != Y.astype(bool)).sum(axis=1))
    ber_last, fer_last, fer = (float(v) for v in d["metrics"])
    assert abs(abs(int(signed.sum().item())) / (B * g.NZ) - ber_last) < 1e-12        # :112-113
    c = cnt.cpu().numpy()
    assert c[0] == B and c[1] == round(fer_last * B) and c[2] == round(fer * B) and c[4] == B * T
    assert c[3] == int(r.biterr.sum().item())
    # the decode itself does not know about Y: same hard decisions / syndromes as the plain entry point
    plain = dec.decode(xa)
    assert torch.equal(plain.hard_packed, r.hard_packed) and torch.equal(plain.iters, r.iters)
    assert torch.equal(plain.flags & 9, r.flags & 9)
    # one shared codeword == that codeword repeated; early termination stops at the first zero syndrome
    r1, s1, _ = dec.decode_cw(xa[:1].repeat(3, 1, 1), Y[0])
    assert torch.equal(r1.flags, r.flags[:1].repeat(3)) and int(s1[0].item()) == int(signed[0].item())
    e, es, ec = dec.decode_cw(xa, Y, early_term=True)
    hard_e = np.stack([hard_ref[int(it) - 1, b] for b, it in enumerate(e.iters.cpu().numpy())])
    assert np.array_equal(e.biterr.cpu().numpy(), (hard_e != Y.astype(bool)).sum(axis=1))
    assert int(ec[4].item()) == int(np.where((e.flags.cpu().numpy() & 1) != 0, e.iters.cpu().numpy(), T).sum())


def test_generator_with_codeword(codes):
    """ldpc_llr_generate_cw: the all-zero codeword gives ldpc_llr_generate's samples bit for bit; a random codeword moves
    the mean to +-2/sigma^2 per bit (Print_Functions.py:45-46) and leaves punctured / shortened positions alone."""
    import torch
    case = load_case("5g_r050_z64_qms_222_t50")
    g, dec = build_decoder(case)
    import ldpc_error_floor_b200 as L
    fdec = L.NMSDecoder(g, L.WeightSet(case["sharing"], dict(case["weights"])), iters=case["T"], decoding_type=1)
    sigma = float(g.sigma([2.0])[0])
    zero = np.zeros(g.NZ, np.uint8)
    a = fdec.generate(sigma, 300, seed=5, frame_offset=9)
    b = fdec.generate(sigma, 300, seed=5, frame_offset=9, codeword=zero)
    assert torch.equal(a, b)
    assert torch.equal(dec.generate(sigma, 300, seed=5), dec.generate(sigma, 300, seed=5, codeword=np.zeros((300, g.NZ), np.uint8)))
    rng = np.random.RandomState(3)
    Y = rng.randint(0, 2, size=(4000, g.NZ)).astype(np.uint8)
    x = fdec.generate(sigma, 4000, seed=6, codeword=Y).cpu().numpy().reshape(4000, -1)
    ps, pe = case["punct"]; ss, se = case["short"]
    assert np.all(x[:, ps - 1:pe] == 0) and np.all(x[:, ss - 1:se] == -case["clip"])
    mask = np.ones(g.NZ, bool); mask[ps - 1:pe] = False; mask[ss - 1:se] = False
    s = (x * (2.0 * Y - 1.0))[:, mask]
    assert abs(s.mean() - 2 / sigma ** 2) < 5 * (2 / sigma) / np.sqrt(s.size)
    assert abs(s.std() / (2 / sigma) - 1) < 0.01
    # shared codeword
    xs = fdec.generate(sigma, 16, seed=6, codeword=Y[0]).cpu().numpy().reshape(16, -1)
    assert np.array_equal(xs[0], x[0])


# ------------------------------------------------------------------------------------------------------------------
# run-time graph specialisation (csrc/nms_jit.cu): a graph the build has never seen gets NVRTC-compiled unrolled kernels
def _random_qc_graph(seed=17, M=5, N=15, z=40):
    rng = np.random.RandomState(seed)
    proto = -np.ones((M, N), dtype=np.int32)
    for i in range(M):
        cols = rng.choice(N - M, size=6, replace=False)
        proto[i, cols] = rng.randint(0, z, size=6)
        proto[i, N - M + i] = 0                              # a staircase over the parity columns
        if i:
            proto[i, N - M + i - 1] = rng.randint(0, z)
    for j in range(N - M):                                   # no unconnected column
        if (proto[:, j] == -1).all():
            proto[rng.randint(M), j] = rng.randint(0, z)
    return proto, z


def test_runtime_specialised_kernels_match_generic_and_oracle(monkeypatch):
    import torch
    import ldpc_error_floor_b200 as L
    from oracle import c_oracle
    proto, z = _random_qc_graph()
    M, N = proto.shape
    rng = np.random.RandomState(3)
    T = 12
    ws = L.WeightSet([2, 2, 2], {0: rng.uniform(0.5, 1.0, (T, M)).astype(np.float32), 1: rng.uniform(0.3, 0.9, (T, M)).astype(np.float32),
                                 2: rng.uniform(0.8, 1.1, (T, N)).astype(np.float32)})
    g = L.BaseGraph(proto, z)
    dec = L.NMSDecoder(g, ws, iters=T)
    assert "jit" in dec.kernel_name and dec.packed, dec.kernel_name
    info = dec.mc_info()
    assert info["persistent"] and "jit" in info["kernel"], info
    monkeypatch.setenv("LDPC_B200_NO_JIT", "1")
    gen = L.NMSDecoder(g, ws, iters=T)                        # the table-driven generic kernels
    monkeypatch.delenv("LDPC_B200_NO_JIT")
    assert "jit" not in gen.kernel_name and not gen.mc_info()["persistent"]
    sigma = float(g.sigma([2.5])[0])
    x = dec.generate(sigma, 2500, seed=8)
    for et in (False, True):
        a, b = dec.decode(x, early_term=et), gen.decode(x, early_term=et)
        assert torch.equal(a.hard_packed, b.hard_packed) and torch.equal(a.flags, b.flags)
        assert torch.equal(a.iters, b.iters) and torch.equal(a.biterr, b.biterr)
    ref = c_oracle.decode(proto, z, x.cpu().numpy(), ws.sharing, ws.blocks, T, 2, 5, 20.0, want_all=False)
    r = dec.decode(x, app="last")
    assert np.array_equal(r.app.cpu().numpy(), ref["app_last"])
    assert 0 < int(((r.flags.cpu().numpy() & 1) == 0).sum()) < 2500          # some frames fail, some converge
    for et in (True, False):
        c1, _ = dec.mc_run_host(sigma, 40001, seed=2, frame_offset=5, early_term=et, harvest=L.HARVEST_UNCOR_ANY, capacity=10)
        c2, _ = gen.mc_run_host(sigma, 40001, seed=2, frame_offset=5, early_term=et, harvest=L.HARVEST_UNCOR_ANY, capacity=10)
        assert c1 == c2, (et, c1, c2)


def test_polar_runs_on_a_runtime_specialised_kernel():
    """Polar(64, 48) has row degree 64 and no compiled-in unrolled kernel: it is specialised at run time (its cubin is
    prebuilt by __graft_entry__.build()); the goldens above already hold it to the reference."""
    case = load_case("polar_qms_223_t6")
    g, dec = build_decoder(case)
    assert "jit" in dec.kernel_name, dec.kernel_name


def test_custom_op_handle_is_stable():
    """torch.ops.ldpc_b200.nms_decode names its decoder by a small integer the decoder keeps for life (not id(), not a pointer
    popped after the call), so a captured call site stays valid while other decoders come and go."""
    import gc
    import torch
    from ldpc_error_floor_b200.decoder import op_handle
    case = load_case("wimax_qms_333_t20")
    g, dec = build_decoder(case)
    xa = torch.from_numpy(case["xa"]).cuda()
    h = op_handle(dec)
    a = torch.ops.ldpc_b200.nms_decode(xa, h, 0, False, 0)
    _, other = build_decoder(load_case("mackay_qms_300_t20"))
    assert op_handle(other) != h
    del other
    gc.collect()
    b = torch.ops.ldpc_b200.nms_decode(xa, h, 0, False, 0)
    assert op_handle(dec) == h and torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
    assert torch.equal(dec.decode(xa).hard_packed, a[0])
    with pytest.raises(RuntimeError):
        torch.ops.ldpc_b200.nms_decode(xa, 10 ** 9, 0, False, 0)


# weights for which float32(k * 0.5 * w) lands exactly on a half step although the exact product does not: the reference
# rounds TWICE (float32 product, then rint to the grid, Main_Functions.py:267-311, 475-494), so a kernel that fuses the
# multiply into the quantiser's add differs by one step -- ptxas 12.9 does that to mul.rn.f32x2 + add.rn.f32x2 on its own
# (nms_device.cuh mul2_rn_unfused).  0.9f, 0.85f and 0.95f are among them.
DOUBLE_ROUNDING_WEIGHTS = [0.9, 0.85, 0.95, 0.8333333730697632, 0.9166666269302368, 1.0714285373687744]


def test_double_rounding_weights_exist():
    """CPU-side sanity of the constants above (float64 holds the exact product of two float32)."""
    k = np.arange(1, 16, dtype=np.float32) * np.float32(0.5)
    for w in DOUBLE_ROUNDING_WEIGHTS:
        p = k.astype(np.float64) * np.float64(np.float32(w))
        fused = np.rint(p * 2) / 2
        twice = np.rint(np.float32(p).astype(np.float64) * 2) / 2
        assert (fused != twice).any(), w


@pytest.mark.parametrize("graph", ["wimax_qms_333_t20", "5g_r073_z72_qms_300_t20_sys", "mackay_qms_300_t20"])
@pytest.mark.parametrize("raw", [False, True])
def test_weights_that_expose_a_fused_multiply_add(graph, raw):
    """Decoders whose CN / UCN / VN weights are the double-rounding constants, against the C oracle: fast path (no APP
    output, unrolled kernels), APP path, persistent-slot Monte-Carlo kernel; on-grid words and raw float32 LLRs."""
    import torch
    import ldpc_error_floor_b200 as L
    from oracle import c_oracle
    case = load_case(graph)
    g = L.BaseGraph(case["proto"], case["z"], case["punct"], case["short"])
    T = 12
    wc = np.array([DOUBLE_ROUNDING_WEIGHTS[t % 6] for t in range(T)], dtype=np.float32).reshape(T, 1)
    wu = np.array([DOUBLE_ROUNDING_WEIGHTS[(t + 2) % 6] for t in range(T)], dtype=np.float32).reshape(T, 1)
    wv = np.array([DOUBLE_ROUNDING_WEIGHTS[(t + 4) % 6] for t in range(T)], dtype=np.float32).reshape(T, 1)
    sharing, blocks = [3, 3, 3], {0: wc, 1: wu, 2: wv}
    dec = L.NMSDecoder(g, L.WeightSet(sharing, blocks), iters=T, decoding_type=2, q_bit=5, clip_llr=20.0)
    assert dec.packed
    B = 2500
    sigma = float(g.sigma([3.0])[0])
    if raw:
        rng = np.random.RandomState(77)
        xa = (2.0 * (rng.normal(size=(B, g.NZ)) * sigma - 1.0) / sigma ** 2).astype(np.float32)
    else:
        xa = dec.generate(sigma, B, seed=5).reshape(B, -1).cpu().numpy()
    ref = c_oracle.decode(case["proto"], case["z"], xa.reshape(B, g.N, g.z), sharing, blocks, T, 2, 5, 20.0, want_all=True)
    hard_ref = ref["app"] >= 0                                   # [T, B, NZ]
    synd_ok = ~ref["synd"]
    iters = np.where(synd_ok.any(axis=0), synd_ok.argmax(axis=0) + 1, T)
    xd = torch.from_numpy(xa).cuda()
    fast = dec.decode(xd, unpack=True)                           # unrolled kernels
    assert np.array_equal(fast.hard.cpu().numpy().astype(bool), hard_ref[T - 1])
    assert np.array_equal(fast.iters.cpu().numpy(), iters)
    slow = dec.decode(xd, app="all")                             # table-driven VN phase
    assert np.array_equal(slow.app.reshape(T, B, -1).cpu().numpy(), ref["app"])
    et = dec.decode(xd, early_term=True, unpack=True)
    stop = np.where(synd_ok.any(axis=0), synd_ok.argmax(axis=0), T - 1)
    assert np.array_equal(et.hard.cpu().numpy().astype(bool), hard_ref[stop, np.arange(B)])
    if not raw:                                                  # the fused Monte-Carlo launch draws these very words
        cnt, _ = dec.mc_run_host(sigma, B, seed=5, early_term=True)
        ones = hard_ref.sum(axis=2)
        assert cnt["iters"] == int(iters.sum())
        assert cnt["frame_err_last"] == int((ones[stop, np.arange(B)] > 0).sum())
        assert cnt["bit_err_last"] == int(ones[stop, np.arange(B)].sum())
