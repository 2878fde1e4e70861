mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/mc_sweep.py wimax 2>&1 | tee gpurun_out/mc_sweep_wimax.txt
python tools/prof_mc.py 5g_r073_z72 5.5 2>&1 | tail -1
python tools/prof_mc.py mackay 5.0 4194304 2>&1 | tail -1
