# scratch job script for gpurun (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/t_all.txt
cat gpurun_out/t_all.txt
timeout 600 python tools/jit_check.py > gpurun_out/jit_check.txt 2>&1; cat gpurun_out/jit_check.txt | cut -c1-330
