mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/f32_sweep.py 2>&1 | tee gpurun_out/f32_sweep.txt
python tools/prof_one.py bch 1 5 262144 2>&1 | tail -1
