import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (development container only)")


def golden_path(name):
    return os.path.join(GOLDEN, name)


def load_case(name):
    """decode_<name>.npz -> dict with python-typed metadata."""
    d = dict(np.load(golden_path(f"decode_{name}.npz")))
    z, ps, pe, ss, se = (int(v) for v in d["meta"])
    case = {
        "name": name, "proto": d["proto"].astype(np.int32), "z": z, "punct": (ps, pe), "short": (ss, se),
        "sharing": [int(v) for v in d["sharing"]], "T": int(d["T"]), "decoding_type": int(d["decoding_type"]),
        "q_bit": int(d["q_bit"]), "clip": float(d["clip"]), "xa": d["xa"], "app": d["app"],
        "weights": {i: d[f"w{i}"] for i in range(3) if f"w{i}" in d},
    }
    case["raw_sharing"], case["raw_weights"] = list(case["sharing"]), dict(case["weights"])
    case["fixed_iter"] = int(d["fixed_iter"]) if "fixed_iter" in d else 0
    if any(c in (4, 5) for c in case["sharing"]):
        # temporal sharing: the reference ran with fixed_iter + 1 variables; everything on our side takes the
        # equivalent per-iteration table (formats.expand_temporal)
        from ldpc_error_floor_b200 import formats
        ws = formats.expand_temporal(formats.WeightSet(case["sharing"], case["weights"]), case["T"], case["fixed_iter"])
        case["sharing"], case["weights"] = ws.sharing, ws.blocks
    if "target_node" in d:
        case.update(target_node=int(d["target_node"]), ya_output_all=d["ya_output_all"], uncor_flag=d["uncor_flag"],
                    error_num=d["error_num"], metrics=d["metrics"])
    return case


def all_cases(sum_product=False):
    """Golden decode cases; the sum-product ones (decoding_type 0, their own parity criterion) only on request."""
    import glob
    names = sorted(os.path.basename(p)[len("decode_"):-len(".npz")] for p in glob.glob(golden_path("decode_*.npz")))
    return [n for n in names if ("_sp_" in n) == sum_product]


@pytest.fixture(scope="session")
def codes():
    return dict(np.load(golden_path("codes.npz")))
