mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/f32_sweep.py 2>&1 | tee gpurun_out/f32_sweep.txt
