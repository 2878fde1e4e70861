mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/mc_sweep.py wimax 2>&1 | tee gpurun_out/mc_sweep_wimax.txt
python tools/mc_sweep.py 5g_r073_z72 "3.0 4.0 5.0 6.0" 2>&1 | tee gpurun_out/mc_sweep_z72.txt
