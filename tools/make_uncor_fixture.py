#!/usr/bin/env python
"""Regenerates the missing Inputs/[Uncor]_wman_N0576_R34_z24_Test.txt of the reference (.MISSING_LARGE_BLOBS) as a committed
fixture: every word of 1 572 864 generated WiMAX frames at 3.5 dB that the shipped 20-iteration base decoder never corrects
(criterion D9, Print_Functions.py:105-111, 120-126), sorted, in the compact LDPCQ8 form (formats.write_uncor_q8; the text form
is `formats.uncor_q8_to_text`).  bench.py -- both arms -- and the tests read it; needs a GPU:
    python tools/make_uncor_fixture.py            # writes tests/golden/uncor_wimax_3.5dB.q8"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ldpc_error_floor_b200 as L
from ldpc_error_floor_b200 import formats

SNR_DB, SEED, FRAMES = 3.5, 20261018, 3 * (1 << 19)
OUT = os.path.join(ROOT, "tests", "golden", "uncor_wimax_3.5dB.q8")


def harvest():
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "codes.npz")))
    g = L.BaseGraph(d["graph/wimax/proto"].astype(np.int32), int(d["graph/wimax/meta"][0]))
    ws = L.WeightSet([int(v) for v in d["weights/wimax_base20/sharing"]], {i: d[f"weights/wimax_base20/block{i}"] for i in range(3)})
    dec = L.NMSDecoder(g, ws, decoding_type=2, q_bit=5, clip_llr=20.0, device=0)
    cnt, rows = dec.mc_run_host(float(g.sigma([SNR_DB])[0]), FRAMES, seed=SEED, harvest=L.HARVEST_UNCOR_ANY, capacity=20000)
    assert cnt["harvested"] == rows.shape[0]
    rows = rows[np.lexsort(rows.T[::-1])]                    # harvest order depends on CTA scheduling: sort
    return formats.llr_to_q8(rows, 0.5), cnt


if __name__ == "__main__":
    words, cnt = harvest()
    formats.write_uncor_q8(OUT, words, step=0.5, snr_db=SNR_DB, seed=SEED)
    print(OUT, words.shape, cnt)
