mkdir -p gpurun_out
nvidia-smi -L | wc -l
python tools/materialize_files.py gpurun_out/files > /dev/null
F=gpurun_out/files
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
# config 5: 5G NR R0.73 n2112 z72, error-floor points sharded over 8 GPUs (plain 0.8 min-sum weights: none are shipped)
timeout 200 $TR --master-port 29531 -m ldpc_error_floor_b200.campaign --graph $F/BaseGraph/5G_LDPC_R0.73_n_dec2304_n2112_k1536_z72_s1537_1584.txt --z 72 --punct 1 144 --short 1537 1584 --ms-weight 0.8 --systematic --snr 4.5 5.0 --frames 3e9 --min-errors 200 --chunk 4194304 --harvest gpurun_out/uncor_z72_8gpu.txt --max-uncor 2000 --json gpurun_out/camp_z72_8gpu.json 2>&1 | grep -v "^W\|^\*\|^$" | tee gpurun_out/camp_z72_8gpu.txt
wc -l gpurun_out/uncor_z72_8gpu.txt; rm -f gpurun_out/uncor_z72_8gpu.txt
# config 4: 5G NR R0.50 n1024 z64, base NMS (rows 0-19) + boosted post decoder (50 rows) on the compacted failures
W=$F/Results/5G/5G_LDPC_R0.50_n_dec1280_n1024_k512_z64_s513_640_Weight_End50.txt
timeout 120 $TR --master-port 29532 -m ldpc_error_floor_b200.campaign --graph $F/BaseGraph/5G_LDPC_R0.50_n_dec1280_n1024_k512_z64_s513_640.txt --z 64 --punct 1 128 --short 513 640 --weights $W --iters 20 --post-weights $W --post-iters 50 --snr 2.5 3.0 --frames 2e9 --min-errors 2000 --max-uncor 20000 --json gpurun_out/camp_z64_8gpu.json 2>&1 | grep -v "^W\|^\*\|^$" | tee gpurun_out/camp_z64_8gpu.txt
timeout 200 $TR --master-port 29533 bench.py --gpus 8 --skip-cpu > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
head -c 300 gpurun_out/bench_n8.json; echo; tail -3 gpurun_out/bench_n8.err
