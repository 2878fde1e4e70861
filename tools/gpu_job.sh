mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --skip-cpu > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_final.json"))
print("frames/s %.4e" % j["frames_per_s"], "value", j["value"], "e2e %.3e" % j["e2e"]["frames_per_s"], "q8 %.3e" % j["e2e_q8"]["frames_per_s"], "mc", "%.3e %.3e" % (j["mc"]["frames_per_s"], j["mc"]["frames_per_s_early_stop"]), "float %.3e" % j["float_min_sum"]["frames_per_s"], j["roofline"]["frac"], j["gpu_launches"])
PY
