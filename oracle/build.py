"""Builds the C oracle into oracle/_build/ (git-ignored, travels with gpurun).
TEST INFRASTRUCTURE ONLY.  `python -m oracle.build` or `oracle.build.build()`.

The reference is pure Python (no C/C++ sources), so there is no `oracle/_ref`
binary to compile from /root/reference -- see DESIGN.md "Oracle"."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SRC = os.path.join(HERE, "nms_oracle.c")


def _compile(out_name: str, extra):
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, out_name)
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(SRC):
        return out
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", out, SRC, "-lm"] + list(extra)
    subprocess.run(cmd, check=True)
    return out


def build() -> str:
    _compile("libnms_oracle_naive.so", ["-DNMS_ORACLE_NAIVE"])
    return _compile("libnms_oracle.so", [])


if __name__ == "__main__":
    print(build())
