mkdir -p gpurun_out
python tools/materialize_files.py /tmp/files > /dev/null
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29531 -m ldpc_error_floor_b200.campaign --graph /tmp/files/BaseGraph/5G_LDPC_R0.73_n_dec2304_n2112_k1536_z72_s1537_1584.txt --z 72 --punct 1 144 --short 1537 1584 --ms-weight 0.8 --systematic --snr 5.5 6.0 --frames 4e9 --min-errors 500 --chunk 4194304 --harvest /tmp/uncor_z72.txt --max-uncor 2000 --json gpurun_out/camp_z72_8gpu_staged.json 2>&1 | grep -v "^W\|^\*\|^$" | tee gpurun_out/camp_z72_8gpu_staged.txt
wc -l /tmp/uncor_z72.txt
