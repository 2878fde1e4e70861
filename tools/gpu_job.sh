mkdir -p gpurun_out
python -m pytest tests/test_gpu_drivers.py -x -q 2>&1 | tail -15
python tools/materialize_files.py gpurun_out/files > /dev/null
F=gpurun_out/files
python -m ldpc_error_floor_b200.campaign --graph $F/BaseGraph/wman_N0576_R34_z24.txt --z 24 --weights $F/Weights/C0_wman_N0576_R34_z24_Opt_Weight_End20.txt --snr 3.5 4.0 4.5 5.0 --frames 4e8 --min-errors 200 --post-weights $F/Results/WiMAX/Weights_Iter50.txt --json gpurun_out/campaign_wimax.json 2>&1 | tee gpurun_out/campaign_wimax.txt
