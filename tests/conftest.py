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
    return case


def all_cases():
    import glob
    return sorted(os.path.basename(p)[len("decode_"):-len(".npz")] for p in glob.glob(golden_path("decode_*.npz")))


@pytest.fixture(scope="session")
def codes():
    return dict(np.load(golden_path("codes.npz")))
