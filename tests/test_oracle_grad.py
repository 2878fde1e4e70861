"""Training step ("next" row N1): the hand-restated backward (oracle/nms_grad_oracle.py) against the goldens minted
from torch.autograd of the reference's OWN loss + forward code (tests/golden/grad_*.npz, oracle/ref_grad.py), and --
where /root/reference exists -- against that autograd run directly."""
import glob
import os

import numpy as np
import pytest

from conftest import golden_path
from oracle import nms_grad_oracle as go, nms_oracle as ob, ref_runner

CASES = sorted(os.path.basename(p)[len("grad_"):-len(".npz")] for p in glob.glob(golden_path("grad_*.npz")))


def load_grad_case(name):
    d = dict(np.load(golden_path(f"grad_{name}.npz")))
    z, ps, pe, ss, se = (int(v) for v in d["meta"])
    sharing = [int(v) for v in d["sharing"]]
    tn = int(d["target_node"])
    return {"proto": d["proto"].astype(np.int32), "z": z, "punct": (ps, pe), "short": (ss, se), "sharing": sharing,
            "T": int(d["T"]), "t_lo": int(d["t_lo"]), "loss_type": int(d["loss_type"]), "etha": float(d["etha"]),
            "decoding_type": int(d["decoding_type"]), "q_bit": int(d["q_bit"]), "clip": float(d["clip"]), "xa": d["xa"],
            "loss": float(d["loss"]), "target_node": None if tn < 0 else tn,
            "weights": {i: d[f"w{i}"] for i in range(3) if sharing[i] > 0},
            "grads": {i: d[f"g{i}"] for i in range(3) if sharing[i] > 0}}


def check_grads(got, case, rtol=2e-4):
    """got: {i: [T, width]}; compared per weight type against the golden with a tolerance relative to the largest
    gradient of that type (sums of ~1e4 float32 terms in a different order)."""
    for i, ref in case["grads"].items():
        g = np.asarray(got[i], dtype=np.float64).reshape(ref.shape)
        scale = max(np.abs(ref).max(), 1e-7)
        assert np.abs(g - ref).max() <= rtol * scale, (i, np.abs(g - ref).max(), scale)
        assert not g[:case["t_lo"]].any()                      # frozen iterations get no gradient


def test_there_are_goldens():
    assert len(CASES) >= 8


@pytest.mark.parametrize("name", CASES)
def test_backward_restatement_matches_golden(name):
    c = load_grad_case(name)
    g = ob.OracleGraph(c["proto"], c["z"], c["punct"], c["short"])
    r = go.loss_and_grads(g, c["xa"], c["sharing"], c["weights"], c["T"], c["t_lo"], c["loss_type"], c["etha"],
                          c["decoding_type"], c["q_bit"], c["clip"], c["target_node"])
    assert r["loss"] == pytest.approx(c["loss"], rel=1e-5, abs=1e-7)
    got = {}
    for i in c["grads"]:
        got[i] = np.zeros_like(c["grads"][i], dtype=np.float64)
        for t in range(c["t_lo"], c["T"]):
            got[i][t] = r["grads"][(i, t)]
    check_grads(got, c, rtol=2e-5)


@pytest.mark.skipif(not ref_runner.reference_available(), reason="needs /root/reference")
def test_goldens_are_what_the_reference_loss_gives_today():
    from oracle import ref_grad
    c = load_grad_case("5g_r073_z32_qms_222_fer_t8")
    r = ref_grad.loss_and_grads(c["proto"].astype(int), c["z"], c["sharing"], c["weights"], c["xa"], c["T"],
                                iter_start=c["t_lo"], loss_type=c["loss_type"], etha=c["etha"],
                                decoding_type=c["decoding_type"], punct=c["punct"], short=c["short"])
    assert r["loss"] == pytest.approx(c["loss"], rel=1e-6)
    for (i, t), gr in r["grads"].items():
        assert np.allclose(gr, c["grads"][i][t], rtol=1e-5, atol=1e-9)
