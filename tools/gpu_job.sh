# one GPU session: parity tests, plain bench, launch list, one full ncu capture of the top kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; echo rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cur.csv python bench.py --steps 2 --warmup 3 --frames 262144 --e2e-frames 32768 --skip-cpu > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_h2 -s 6 -c 1 -o gpurun_out/prof_cur -f python bench.py --steps 2 --warmup 3 --frames 262144 --e2e-frames 32768 --skip-cpu > gpurun_out/ncu_f.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
