# scratch job script for gpurun (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t_all.txt
cat gpurun_out/t_all.txt
timeout 300 python tools/mc_sweep.py 5g_r073_z72 "4.0 5.5 7.0" 1 "2,2 3,2" > gpurun_out/sweep_z72.txt 2>&1; tail -4 gpurun_out/sweep_z72.txt
timeout 300 python tools/mc_sweep.py wimax "3.0 5.0" 0 "4,2" > gpurun_out/sweep_wimax.txt 2>&1; tail -3 gpurun_out/sweep_wimax.txt
python bench.py --skip-cpu > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 1500 gpurun_out/bench_a.json
