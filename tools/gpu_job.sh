set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
python - <<'PY' > gpurun_out/h2d_probe.txt 2>&1
import torch, time
x = torch.empty(1<<28, dtype=torch.uint8).pin_memory()
d = torch.empty(1<<28, dtype=torch.uint8, device='cuda')
for n in (1<<20, 1<<23, 1<<25, 1<<28):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): d[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('h2d', n, 10*n/dt/1e9, 'GB/s')
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): x[:n].copy_(d[:n], non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('d2h', n, 10*n/dt/1e9, 'GB/s')
import os; print('cpus', os.cpu_count())
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --frames 262144 --e2e-frames 32768 --skip-cpu > gpurun_out/ncu_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_h2 -s 6 -c 1 -o gpurun_out/prof_v3 -f python bench.py --steps 2 --warmup 3 --frames 262144 --e2e-frames 32768 --skip-cpu > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out
