#!/usr/bin/env python
"""Recreate the reference's data-file layout (BaseGraph/*.txt, Weights/*.txt, Results/**) from the committed
fixture tests/golden/codes.npz, so the file-based drivers (drivers.evaluate, campaign) can run where
/root/reference does not exist (the GPU box).   usage: python tools/materialize_files.py OUTDIR"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldpc_error_floor_b200 import formats  # noqa: E402

WEIGHT_PATHS = {
    "wimax_base20": "Weights/C0_wman_N0576_R34_z24_Opt_Weight_End20.txt",
    "wimax_boost50": "Results/WiMAX/Weights_Iter50.txt",
    "wifi_boost50": "Results/WIFI/Weights_Iter50.txt",
}


def materialize(outdir):
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "codes.npz")))
    made = {}
    for key in sorted(k.split("/")[1] for k in d if k.endswith("/proto")):
        stem = str(d[f"graph/{key}/stem"])
        p = os.path.join(outdir, "BaseGraph", stem + ".txt")
        os.makedirs(os.path.dirname(p), exist_ok=True)
        formats.write_base_graph(p, d[f"graph/{key}/proto"], crlf=bool(d[f"graph/{key}/crlf"]))
        made[key] = p
    for key in sorted(k.split("/")[1] for k in d if k.startswith("weights/") and k.endswith("/text")):
        rel = WEIGHT_PATHS.get(key)
        if rel is None:   # 5G: Results/5G/<graph stem>_Weight_End50.txt
            rel = "Results/5G/" + str(d[f"graph/{key[:-len('_boost50')]}/stem"]) + "_Weight_End50.txt"
        p = os.path.join(outdir, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w", newline="") as fh:
            fh.write(str(d[f"weights/{key}/text"]))
        made["w:" + key] = p
    os.makedirs(os.path.join(outdir, "Inputs"), exist_ok=True)
    return made


if __name__ == "__main__":
    for k, v in materialize(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/files").items():
        print(k, v)
