"""Development container only: pins the oracle restatements directly against the reference's OWN code
(Main_Functions.build_neural_network / Print_Functions, imported unmodified under the numpy TF shim).
Skipped where /root/reference does not exist (GPU box): there the committed goldens stand in."""
import numpy as np
import pytest

from oracle import ref_runner

pytestmark = pytest.mark.skipif(not ref_runner.reference_available(), reason="needs /root/reference")


def _setup(codes, gkey, wkey, T, decoding_type=2, q_bit=5):
    proto = codes[f"graph/{gkey}/proto"].astype(int)
    z, ps, pe, ss, se, _ = (int(v) for v in codes[f"graph/{gkey}/meta"])
    sharing = [int(v) for v in codes[f"weights/{wkey}/sharing"]]
    w = {i: codes[f"weights/{wkey}/block{i}"] for i in range(3)}
    rd = ref_runner.ReferenceDecoder(proto, z, sharing, w, T, decoding_type, q_bit, 20.0, (ps, pe), (ss, se), [3.0])
    return proto, z, (ps, pe), (ss, se), sharing, w, rd


@pytest.mark.parametrize("gkey,wkey,T,dt", [("wimax", "wimax_base20", 20, 2), ("wimax", "wimax_base20", 12, 1),
                                            ("5g_r073_z32", "5g_r073_z32_boost50", 25, 2)])
def test_oracles_match_reference_decode(codes, gkey, wkey, T, dt):
    from oracle import c_oracle, nms_oracle as ob
    proto, z, punct, short, sharing, w, rd = _setup(codes, gkey, wkey, T, dt)
    g = ob.OracleGraph(proto, z, punct, short)
    X, _ = ob.create_mix_epoch(rd.snr_sigma, np.random.RandomState(5), np.random.RandomState(6), 5, g.N, z, dt,
                               punct, short, 5, 20.0)
    ref = rd.decode(X)["app"]
    a = ob.decode(g, X, sharing, w, T, dt, 5, 20.0)["app"]
    b = c_oracle.decode(proto, z, X, sharing, w, T, dt, 5, 20.0)["app"]
    if dt == 2:
        assert np.array_equal(a, ref) and np.array_equal(b, ref)
    else:
        for o in (a, b):
            assert (np.abs(o - ref) / np.maximum(1, np.abs(ref))).max() <= 1e-5
            assert np.array_equal(o >= 0, ref >= 0)


def test_sample_generator_restatement_is_identical(codes):
    """oracle.create_mix_epoch draws exactly what Print_Functions.create_mix_epoch draws (same RandomState order)."""
    from oracle import nms_oracle as ob
    _, pf = ref_runner.load_reference()
    for gkey, dt in (("wimax", 2), ("5g_r050_z64", 2), ("wimax", 1)):
        proto = codes[f"graph/{gkey}/proto"]
        z, ps, pe, ss, se, _ = (int(v) for v in codes[f"graph/{gkey}/meta"])
        N = proto.shape[1]
        sig = [0.6, 0.55]
        Xr, Yr = pf.create_mix_epoch(np.array(sig), np.random.RandomState(1), np.random.RandomState(2), 6, N, N, z, [],
                                     True, dt, ps, pe, ss, se, 5, 20.0)
        Xo, Yo = ob.create_mix_epoch(sig, np.random.RandomState(1), np.random.RandomState(2), 6, N, z, dt, (ps, pe),
                                     (ss, se), 5, 20.0)
        assert np.array_equal(np.asarray(Xr, np.float32), Xo) and np.array_equal(Yr, Yo)


def test_calc_ber_fer_restatement(codes):
    from oracle import nms_oracle as ob
    _, pf = ref_runner.load_reference()
    rng = np.random.RandomState(0)
    app = rng.normal(-3, 4, size=(4 * 10, 96)).astype(np.float32)
    app[5] = -1.0
    Y = np.zeros((10, 96), dtype=np.int64)
    r = pf.calc_ber_fer(app, 4, Y, 10)
    o = ob.calc_ber_fer(app, 4, Y, 10)
    for a, b in zip(r, o):
        assert np.array_equal(np.asarray(a), np.asarray(b))


def test_mc_fixture_reproduces(codes):
    """The committed mc_wimax.npz is what compute_results returns today (first 40 frames of the 2.5 dB point)."""
    proto, z, punct, short, sharing, w, _ = _setup(codes, "wimax", "wimax_base20", 20)
    rd = ref_runner.ReferenceDecoder(proto, z, sharing, w, 20, 2, 5, 20.0, punct, short, [2.5])
    from oracle import nms_oracle as ob
    g = ob.OracleGraph(proto, z)
    res, harvested = ob.monte_carlo(g, 40, rd.snr_sigma, np.random.RandomState(2044), np.random.RandomState(1076), 20,
                                    sharing, w, 20, 2, 5, 20.0, collect=True)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        ref, _ = rd.compute_results(40, 2044, 1076, 20, sampling_type=2, cwd=tmp)
    assert np.array_equal(res[:3], ref[:3])


def test_reference_rounds_twice_for_weights_like_0p9(codes):
    """The constants of tests/test_gpu_parity.py (DOUBLE_ROUNDING_WEIGHTS): with these weights a product rounded ONCE into
    the quantiser differs from the float32 product quantised afterwards.  The reference (numpy shim: float32 multiply, then
    round(x * qk) / qk, Main_Functions.py:267-311, 475-494) does the latter, and so do both oracle restatements -- which is
    what the CUDA kernels are held to (nms_device.cuh mul2_rn_unfused)."""
    from oracle import c_oracle, nms_oracle as ob
    proto = codes["graph/wimax/proto"].astype(int)
    z, ps, pe, ss, se, _ = (int(v) for v in codes["graph/wimax/meta"])
    T = 12
    ws = [0.9, 0.85, 0.95, 0.8333333730697632, 0.9166666269302368, 1.0714285373687744]
    w = {0: np.array([ws[t % 6] for t in range(T)], np.float32).reshape(T, 1),
         1: np.array([ws[(t + 2) % 6] for t in range(T)], np.float32).reshape(T, 1),
         2: np.array([ws[(t + 4) % 6] for t in range(T)], np.float32).reshape(T, 1)}
    sharing = [3, 3, 3]
    rd = ref_runner.ReferenceDecoder(proto, z, sharing, w, T, 2, 5, 20.0, (ps, pe), (ss, se), [3.0])
    g = ob.OracleGraph(proto, z, (ps, pe), (ss, se))
    X, _ = ob.create_mix_epoch(rd.snr_sigma, np.random.RandomState(15), np.random.RandomState(16), 40, g.N, z, 2,
                               (ps, pe), (ss, se), 5, 20.0)
    ref = rd.decode(X)["app"]
    a = ob.decode(g, X, sharing, w, T, 2, 5, 20.0)["app"]
    b = c_oracle.decode(proto, z, X, sharing, w, T, 2, 5, 20.0)["app"]
    assert np.array_equal(a, ref) and np.array_equal(b, ref)
    # and the fused form would NOT have matched: count the products where it differs (a float64 product of two float32 is exact)
    k = np.arange(1, 16, dtype=np.float32) * np.float32(0.5)
    p = k[:, None].astype(np.float64) * np.array(ws, np.float32)[None, :].astype(np.float64)
    assert (np.rint(p * 2) != np.rint(np.float32(p).astype(np.float64) * 2)).sum() >= 6
