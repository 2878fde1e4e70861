mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
SWEEP_SNR=1.0 python tools/geom_sweep.py wimax "0,0" 2>&1 | tail -1
