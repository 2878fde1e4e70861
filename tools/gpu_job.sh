# scratch job script for gpurun (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t_all.txt
cat gpurun_out/t_all.txt
timeout 300 python tools/mc_sweep.py 5g_r073_z72 "5.5 7.0" 1 "2,2 3,2 4,2 4,3" > gpurun_out/sweep_z72.txt 2>&1; tail -3 gpurun_out/sweep_z72.txt
timeout 300 python tools/mc_sweep.py wimax "3.0 5.0" 0 "4,2" > gpurun_out/sweep_wimax.txt 2>&1; tail -3 gpurun_out/sweep_wimax.txt
python bench.py --skip-cpu > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_c.json"))
print("frames/s %.4e" % j["frames_per_s"], "value", j["value"], "e2e %.3e" % j["e2e"]["frames_per_s"], "q8 %.3e" % j["e2e_q8"]["frames_per_s"], "mc", "%.3e %.3e" % (j["mc"]["frames_per_s"], j["mc"]["frames_per_s_early_stop"]), "float %.3e" % j["float_min_sum"]["frames_per_s"], j["roofline"]["frac"], "cfg5 %.3e" % j["mc_config5"]["frames_per_s"])
PY
