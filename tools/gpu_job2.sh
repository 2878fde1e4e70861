mkdir -p gpurun_out
nvidia-smi -L
python tools/materialize_files.py gpurun_out/files > /dev/null
F=gpurun_out/files
ARGS="--graph $F/BaseGraph/802_11n_N648_R56_z27.txt --z 27 --weights $F/Results/WIFI/Weights_Iter50.txt --iters 20 --snr 3.0 4.0 --frames 16777216 --harvest gpurun_out/uncor_wifi_N.txt --max-uncor 20000"
python -m ldpc_error_floor_b200.campaign ${ARGS/_N.txt/_1.txt} --json gpurun_out/camp_wifi_1gpu.json 2>&1 | tee gpurun_out/camp_wifi_1gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 -m ldpc_error_floor_b200.campaign ${ARGS/_N.txt/_2.txt} --json gpurun_out/camp_wifi_2gpu.json 2>&1 | grep -v "^W\|^\*" | tee gpurun_out/camp_wifi_2gpu.txt
sort gpurun_out/uncor_wifi_1.txt | md5sum; sort gpurun_out/uncor_wifi_2.txt | md5sum; wc -l gpurun_out/uncor_wifi_1.txt gpurun_out/uncor_wifi_2.txt
rm -f gpurun_out/uncor_wifi_1.txt gpurun_out/uncor_wifi_2.txt
python bench.py --gpus 1 --skip-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 600 gpurun_out/bench_n2.err
