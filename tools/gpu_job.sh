# scratch job script for gpurun (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t_all.txt
cat gpurun_out/t_all.txt
python bench.py --skip-cpu > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_e.json"))
print("frames/s %.4e" % j["frames_per_s"], "value", j["value"], "e2e %.3e" % j["e2e"]["frames_per_s"], "q8 %.3e" % j["e2e_q8"]["frames_per_s"], "mc", "%.3e %.3e" % (j["mc"]["frames_per_s"], j["mc"]["frames_per_s_early_stop"]), "float %.3e" % j["float_min_sum"]["frames_per_s"], j["roofline"]["frac"], "cfg5 %.3e" % j["mc_config5"]["frames_per_s"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_e_n2.json 2> gpurun_out/bench_e_n2.err; tail -c 400 gpurun_out/bench_e_n2.json; tail -2 gpurun_out/bench_e_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_e_ref_n2.json 2> gpurun_out/bench_e_ref_n2.err; cut -c1-300 gpurun_out/bench_e_ref_n2.json
