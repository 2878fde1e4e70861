mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --skip-cpu > gpurun_out/bench_fma.json 2> gpurun_out/bench_fma.err
python - <<'PY'
import json
for n in ("fma",):
    try:
        j = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, "frames/s %.3e" % j["frames_per_s"], "e2e %.3e" % j["e2e"]["frames_per_s"], "q8 %.3e" % j["e2e_q8"]["frames_per_s"], "mc", "%.3e %.3e" % (j["mc"]["frames_per_s"], j["mc"]["frames_per_s_early_stop"]), "float %.3e" % j["float_min_sum"]["frames_per_s"], j["geometry"], j["clocks"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-2000:])
PY
ncu --set full --clock-control none --import-source on -k regex:nms_h2 -s 6 -c 1 -o gpurun_out/prof_h2 -f python bench.py --steps 2 --warmup 3 --frames 262144 --e2e-frames 32768 --skip-cpu > gpurun_out/ncu_h2.log 2>&1
tail -2 gpurun_out/ncu_h2.log
