# scratch job script for gpurun (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t_all.txt
cat gpurun_out/t_all.txt
timeout 600 python tools/jit_check.py > gpurun_out/jit_check.txt 2>&1; cat gpurun_out/jit_check.txt | cut -c1-330
python bench.py --skip-cpu > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_d.json"))
print("frames/s %.4e" % j["frames_per_s"], "value", j["value"], "e2e %.3e" % j["e2e"]["frames_per_s"], "q8 %.3e" % j["e2e_q8"]["frames_per_s"], "mc", "%.3e %.3e" % (j["mc"]["frames_per_s"], j["mc"]["frames_per_s_early_stop"]), "float %.3e" % j["float_min_sum"]["frames_per_s"], j["roofline"]["frac"], "cfg5 %.3e" % j["mc_config5"]["frames_per_s"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
