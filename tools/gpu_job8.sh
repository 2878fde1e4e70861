mkdir -p gpurun_out
python tools/f32_sweep.py 2>&1 | tee gpurun_out/f32_sweep.txt
for n in 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $n --skip-cpu > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $n --steps 2 --warmup 1 > gpurun_out/bench_ref_n$n.json 2> /dev/null
done
head -c 400 gpurun_out/bench_n8.json; echo; head -c 200 gpurun_out/bench_ref_n8.json; echo; tail -3 gpurun_out/bench_n8.err
