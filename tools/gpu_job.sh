mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; echo rc=$?; tail -3 gpurun_out/bench_cur.err
