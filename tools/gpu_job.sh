mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0, ".")
from ldpc_error_floor_b200 import _lib
for _ in range(2):
    print({k: round(v, 2) for k, v in _lib.alu_peak_probe(0, kinds=("ffma", "fadd", "fmnmx", "lop3", "iadd", "hfma2", "hmnmx2")).items()})
PY
python -m pytest tests/test_gpu_mc.py -x -q 2>&1 | tail -3
