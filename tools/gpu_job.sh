mkdir -p gpurun_out
python tools/materialize_files.py gpurun_out/files > /dev/null
F=gpurun_out/files
G=$F/BaseGraph/5G_LDPC_R0.50_n_dec1280_n1024_k512_z64_s513_640.txt
W=$F/Results/5G/5G_LDPC_R0.50_n_dec1280_n1024_k512_z64_s513_640_Weight_End50.txt
python -m ldpc_error_floor_b200.campaign --graph $G --z 64 --punct 1 128 --short 513 640 --systematic --weights $W --iters 20 --post-weights $W --post-iters 50 --snr 2.5 3.0 3.5 --frames 3e8 --min-errors 3000 --max-uncor 20000 --json gpurun_out/camp_z64_sys.json 2>&1 | grep -v "^W\|^\*\|^$" | tee gpurun_out/camp_z64_sys.txt
