mkdir -p gpurun_out
python tools/materialize_files.py gpurun_out/files > /dev/null
F=gpurun_out/files
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29533 bench.py --gpus 8 --skip-cpu > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
head -c 260 gpurun_out/bench_n8.json; echo
timeout 60 $TR --master-port 29534 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_n8.json 2> /dev/null
head -c 200 gpurun_out/bench_ref_n8.json; echo
# config 5, a deeper point: 5G NR R0.73 n2112 z72 at 5.5 dB, up to 3e9 frames over 8 GPUs
timeout 150 $TR --master-port 29531 -m ldpc_error_floor_b200.campaign --graph $F/BaseGraph/5G_LDPC_R0.73_n_dec2304_n2112_k1536_z72_s1537_1584.txt --z 72 --punct 1 144 --short 1537 1584 --ms-weight 0.8 --systematic --snr 5.5 --frames 3e9 --min-errors 500 --chunk 4194304 --harvest gpurun_out/uncor_z72_8gpu.txt --max-uncor 2000 --json gpurun_out/camp_z72_8gpu_55.json 2>&1 | grep -v "^W\|^\*\|^$" | tee gpurun_out/camp_z72_8gpu_55.txt
wc -l gpurun_out/uncor_z72_8gpu.txt; rm -f gpurun_out/uncor_z72_8gpu.txt
