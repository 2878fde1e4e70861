mkdir -p gpurun_out
python bench.py > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v8_ref.json 2> /dev/null
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_v8.json"))
print("frames/s %.4e" % j["frames_per_s"], "value", j["value"], "e2e %.3e" % j["e2e"]["frames_per_s"], "q8 %.3e" % j["e2e_q8"]["frames_per_s"], "mc", "%.3e %.3e" % (j["mc"]["frames_per_s"], j["mc"]["frames_per_s_early_stop"]), "float %.3e" % j["float_min_sum"]["frames_per_s"], j["roofline"]["frac"], j["clocks"], j["cpu_baseline"]["frames_per_s"])
PY
python bench.py --steps 2 --warmup 3 --frames 262144 --e2e-frames 32768 --skip-cpu > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 2 --warmup 3 --frames 262144 --e2e-frames 32768 --skip-cpu > gpurun_out/ncu_l8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_h2_spec_wimax_fp8 -s 4 -c 1 -o gpurun_out/prof_h2 -f python bench.py --steps 2 --warmup 3 --frames 262144 --e2e-frames 32768 --skip-cpu > gpurun_out/ncu_h2.log 2>&1
tail -2 gpurun_out/ncu_h2.log
