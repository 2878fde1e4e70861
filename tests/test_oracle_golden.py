"""CPU: pins the oracle restatements (oracle/nms_oracle.py, oracle/nms_oracle.c) to the golden
vectors minted from the reference's own build_neural_network (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import all_cases, load_case

REL_TOL = 1e-5


def _check(case, app):
    ref = case["app"]
    if case["decoding_type"] == 2:
        assert np.array_equal(app, ref), f"max |diff| {np.abs(app - ref).max()}"
    else:
        err = np.abs(app - ref) / np.maximum(1.0, np.abs(ref))
        assert err.max() <= REL_TOL
        assert np.array_equal(app >= 0, ref >= 0)


@pytest.mark.parametrize("name", all_cases())
def test_c_oracle_matches_golden(name):      # all_cases(): min-sum and quantised min-sum (no sum-product branch in C)
    from oracle import c_oracle
    case = load_case(name)
    out = c_oracle.decode(case["proto"], case["z"], case["xa"], case["sharing"], case["weights"], case["T"],
                          case["decoding_type"], case["q_bit"], case["clip"])
    _check(case, out["app"])


@pytest.mark.parametrize("name", all_cases(sum_product=True))
def test_numpy_oracle_sum_product_matches_golden(name):
    """decoding_type 0 (Main_Functions.py:238-245): the numpy restatement uses the same float32 tanh / arctanh as the
    reference under the numpy TF shim and multiplies in E(C) order (numpy's reduce_prod over the dense tile may group the
    factors differently: bit-identical on the 5G case, 4e-6 on WiMAX).  (The C oracle has no sum-product branch: libm's
    tanhf differs from numpy's in the last place, and -2 atanh(x) amplifies one ulp of x near 1 to several percent -- see
    test_gpu_parity.test_sum_product_golden for the criterion that follows from that.)"""
    from oracle import nms_oracle as ob
    case = load_case(name)
    g = ob.OracleGraph(case["proto"], case["z"], case["punct"], case["short"])
    out = ob.decode(g, case["xa"], case["sharing"], case["weights"], case["T"], 0, case["q_bit"], case["clip"])
    _check(case, out["app"])


@pytest.mark.parametrize("name", ["wimax_qms_333_t20", "wimax_float_333_t20", "5g_r073_z32_qms_222_t50",
                                  "wimax_qms_112_t6", "polar_qms_223_t6", "wimax_qms_q3_323_t6", "wimax_qms_q6_323_t6"])
def test_numpy_oracle_matches_golden(name):
    from oracle import nms_oracle as ob
    case = load_case(name)
    g = ob.OracleGraph(case["proto"], case["z"], case["punct"], case["short"])
    out = ob.decode(g, case["xa"][:3], case["sharing"], case["weights"], case["T"], case["decoding_type"],
                    case["q_bit"], case["clip"])
    sub = dict(case)
    sub["app"] = case["app"][:, :3]
    _check(sub, out["app"])


@pytest.mark.parametrize("name", ["wimax_qms_333_t20", "wimax_float_333_t20", "bch_qms_333_t10"])
def test_c_oracle_naive_equals_fast(name):
    from oracle import c_oracle
    case = load_case(name)
    a = c_oracle.decode(case["proto"], case["z"], case["xa"], case["sharing"], case["weights"], case["T"],
                        case["decoding_type"], case["q_bit"], case["clip"], naive=False)
    b = c_oracle.decode(case["proto"], case["z"], case["xa"], case["sharing"], case["weights"], case["T"],
                        case["decoding_type"], case["q_bit"], case["clip"], naive=True)
    assert np.array_equal(a["app"], b["app"]) and np.array_equal(a["synd"], b["synd"])
