mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> /dev/null
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value", j["value"], "frames/s %.4e" % j["frames_per_s"], "ms/step", j["ms_per_step"], "e2e %.3e" % j["e2e"]["frames_per_s"], j["check"], j["n_gpus"])
r = json.loads(open("gpurun_out/bench_ref_n2.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["n_gpus"], r["impl"])
PY
tail -3 gpurun_out/bench_n2.err
