"""CPU: the reference's three text formats (F1 BaseGraph, F2 Weights, F3 Inputs/[Uncor]) round-trip
byte-for-byte against every file the reference ships (carried in tests/golden/codes.npz)."""
import os

import numpy as np
import pytest

from ldpc_error_floor_b200 import formats

GRAPHS = ["wimax", "wifi", "mackay", "bch", "polar", "5g_r033_z32", "5g_r050_z32", "5g_r050_z64", "5g_r073_z32",
          "5g_r073_z72"]
WEIGHTS = ["wimax_base20", "wimax_boost50", "wifi_boost50", "5g_r033_z32_boost50", "5g_r050_z32_boost50",
           "5g_r050_z64_boost50", "5g_r073_z32_boost50"]


@pytest.mark.parametrize("key", GRAPHS)
def test_base_graph_roundtrip(codes, tmp_path, key):
    proto = codes[f"graph/{key}/proto"].astype(np.int32)
    crlf = bool(codes[f"graph/{key}/crlf"])
    path = tmp_path / (str(codes[f"graph/{key}/stem"]) + ".txt")
    formats.write_base_graph(str(path), proto, crlf=crlf)
    raw = path.read_bytes()
    assert not raw.endswith(b"\n")                      # shipped files have no trailing newline
    assert (b"\r\n" in raw) == crlf                     # the 5G files are CRLF, the rest LF
    back = formats.read_base_graph(str(path))
    assert back.dtype == np.int32 and np.array_equal(back, proto)
    # np.loadtxt as main_Base.py:67 calls it reads the same matrix
    assert np.array_equal(np.loadtxt(str(path), int, delimiter="\t"), proto)


def test_5g_name_convention(codes):
    for key in GRAPHS:
        stem = str(codes[f"graph/{key}/stem"])
        meta = formats.parse_5g_name(stem)
        z, ps, pe, ss, se, E = (int(v) for v in codes[f"graph/{key}/meta"])
        if key.startswith("5g"):
            assert meta["z"] == z and meta["punct"] == (ps, pe) == (1, 2 * z) and meta["short"] == (ss, se)
            proto = codes[f"graph/{key}/proto"]
            assert meta["n_dec"] == proto.shape[1] * z
            assert meta["k"] == (proto.shape[1] - proto.shape[0]) * z - (se - ss + 1)
        else:
            assert meta is None


@pytest.mark.parametrize("key", WEIGHTS)
def test_weights_roundtrip(codes, tmp_path, key):
    text = str(codes[f"weights/{key}/text"])
    src = tmp_path / "in.txt"
    src.write_text(text)
    ws = formats.read_weights(str(src))
    assert ws.sharing == [int(v) for v in codes[f"weights/{key}/sharing"]]
    for i in range(3):
        assert np.array_equal(ws.blocks[i], codes[f"weights/{key}/block{i}"])
    out = tmp_path / "out.txt"
    formats.write_weights(str(out), ws)
    written = out.read_text()
    assert written.rstrip("\n") == text.rstrip("\n")
    if text.endswith("\n\n"):                           # files Print_Functions.print_weight wrote itself
        assert written == text                          # (C0_wman...End20 was hand-assembled: one newline short)
    # the reference addresses rows by absolute line number (Main_Functions.py:419-421)
    T = ws.iterations
    row = 0
    for i in range(3):
        for t in range(T):
            row += 1
            data = np.loadtxt(str(src), skiprows=1 + row, max_rows=1, delimiter="\t")
            assert np.allclose(np.atleast_1d(data).astype(np.float32), ws.blocks[i][t])
        row += 1


def test_weights_structure_of_shipped_files(codes):
    """rows 0-19 of every boosted file are the base decoder (CN == UCN block)."""
    for key in WEIGHTS:
        cn, ucn = codes[f"weights/{key}/block0"], codes[f"weights/{key}/block1"]
        assert np.array_equal(cn[:20], ucn[:20])
        assert cn.min() >= 0.0 and cn.max() <= 2.0
    assert np.array_equal(codes["weights/wimax_boost50/block0"][:20], codes["weights/wimax_base20/block0"])
    ws = formats.WeightSet([3, 3, 3], {i: codes[f"weights/wimax_boost50/block{i}"] for i in range(3)})
    base = ws.rows(0, 20)
    assert base.iterations == 20 and np.array_equal(base.blocks[2], codes["weights/wimax_base20/block2"])


def test_weights_omit_zero_sharing_blocks(tmp_path):
    ws = formats.WeightSet([3, 0, 2], {0: np.full((4, 1), 0.75, np.float32), 2: np.ones((4, 5), np.float32)})
    p = tmp_path / "w.txt"
    formats.write_weights(str(p), ws)
    lines = p.read_text().split("\n")
    assert lines[0] == "3 0 2" and lines[1] == "" and lines[2] == "0.75" and lines[6] == ""
    back = formats.read_weights(str(p))
    assert back.sharing == [3, 0, 2] and sorted(back.blocks) == [0, 2] and back.blocks[2].shape == (4, 5)


def test_weights_errors(tmp_path):
    p = tmp_path / "bad.txt"
    p.write_text("3 3\n\n0.5\n")
    with pytest.raises(ValueError):
        formats.read_weights(str(p))
    p.write_text("3 0 3\n\n0.5\n0.6\n\n0.7\n\n")       # blocks disagree on T
    with pytest.raises(ValueError):
        formats.read_weights(str(p))
    p.write_text("")
    with pytest.raises(ValueError):
        formats.read_weights(str(p))


def test_uncor_roundtrip(tmp_path):
    rng = np.random.RandomState(3)
    llr = (rng.randint(-15, 16, size=(7, 24, 24)) / 2.0).astype(np.float32)
    llr[0, 0, 0] = 0.0
    path = tmp_path / "Uncor.txt"
    assert formats.append_uncor(str(path), llr[:4]) == 4
    assert formats.append_uncor(str(path), llr[4:]) == 3   # append mode, like write_uncor_file
    lines = path.read_text().splitlines()
    assert len(lines) == 7 and all(ln.startswith("0.0\t0.0\t0.0\t") for ln in lines)
    assert all(len(ln.split("\t")) == 3 + 576 for ln in lines)
    rows = formats.read_uncor(str(path))
    assert rows.shape == (7, 576) and rows.dtype == np.float32
    assert np.array_equal(rows, -llr.reshape(7, -1))            # the file stores log p0/p1
    assert np.array_equal(formats.uncor_to_llr(rows, 24, 24), llr)
    assert formats.read_uncor(str(path), limit=5).shape == (5, 576)
    with pytest.raises(ValueError):
        formats.read_uncor(str(path), limit=8)                  # Main_Functions.py:534-536
    assert formats.uncor_filenames("wman_N0576_R34_z24")[2] == "./Inputs/[Uncor]_wman_N0576_R34_z24_Test.txt"


def test_uncor_matches_reference_writer():
    """mc_wimax.npz holds the Uncor.txt the reference's write_uncor_file appended (Print_Functions.py:120-126)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "mc_wimax.npz")
    if not os.path.exists(path):
        pytest.skip("mc_wimax.npz not minted yet")
    d = np.load(path)
    text = str(d["uncor_text_0"])
    lines = text.splitlines()
    assert lines, "the 2.5 dB point harvests words"
    vals = np.array([[float(t) for t in ln.split("\t")] for ln in lines], dtype=np.float32)
    llr = formats.uncor_to_llr(vals[:, 3:], 24, 24)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "Uncor.txt")
        formats.append_uncor(p, llr)
        assert open(p).read() == text                           # byte-identical, "-0.0" included


def test_uncor_q8_sidecar_round_trip(tmp_path):
    """N2: int8 binary sidecar of the Inputs/[Uncor] text format -- exact on the quantiser grid (the sign of a
    zero, which the '%.1f' text keeps as '-0.0' / '0.0', is the one thing it drops; the decoder ignores it)."""
    from ldpc_error_floor_b200 import formats as F
    rng = np.random.RandomState(0)
    llr = np.clip(np.rint(rng.normal(-3, 3, (50, 576)) * 2) / 2, -7.5, 7.5).astype(np.float32)
    llr[7] = llr[3]
    txt, q8 = str(tmp_path / "u.txt"), str(tmp_path / "u.q8")
    F.append_uncor(txt, llr)
    assert F.uncor_text_to_q8(txt, q8, snr_db=3.5) == 50
    w, meta = F.read_uncor_q8(q8)
    assert w.dtype == np.int8 and w.shape == (50, 576)
    assert meta == {"step": 0.5, "snr_db": 3.5, "seed": 0, "rows": 50}
    assert np.array_equal(F.q8_to_llr(w, 0.5), llr)
    assert os.path.getsize(q8) == 64 + 50 * 576 and os.path.getsize(txt) > 4 * os.path.getsize(q8)
    assert F.write_uncor_q8(q8, w[:5], append=True) == 55
    w2, m2 = F.read_uncor_q8(q8)
    assert m2["rows"] == 55 and np.array_equal(w2[50:], w[:5])
    part, _ = F.read_uncor_q8(q8, limit=10, offset=20)
    assert np.array_equal(part, w[20:30])
    assert F.uncor_text_to_q8(txt, str(tmp_path / "d.q8"), dedup=True) == 49
    back = str(tmp_path / "back.txt")
    F.uncor_q8_to_text(q8, back)
    assert np.array_equal(F.read_uncor(back)[:50], F.read_uncor(txt))       # numerically identical rows
    with pytest.raises(ValueError):
        F.llr_to_q8(np.array([[0.25]], np.float32), 0.5)                     # off the grid
    with pytest.raises(ValueError):
        F.read_uncor_q8(txt)                                                  # not a sidecar
