#!/bin/bash
# BASELINE configs 2, 3, 4 on one B200 with the round-2 kernels (persistent-slot Monte-Carlo): FER sweeps with the stop rule
# ">= N frame errors or the frame budget", harvest + boosted post decoder where the reference ships 50-row weights.
mkdir -p gpurun_out
python tools/materialize_files.py gpurun_out/files > /dev/null
F=gpurun_out/files
C="python -m ldpc_error_floor_b200.campaign"
{
echo "## config 2: MacKay N96 K48 (z = 1), min-sum weight 1.0 and normalised 0.8, quantised (q_bit 5), 20 iterations"
$C --graph $F/BaseGraph/MACKAY_N96_K48.txt --z 1 --ms-weight 1.0 --iters 20 --snr 1 1.5 2 2.5 3 3.5 4 4.5 5 5.5 6 --frames 2e9 --min-errors 2000
$C --graph $F/BaseGraph/MACKAY_N96_K48.txt --z 1 --ms-weight 0.8 --iters 20 --snr 1 1.5 2 2.5 3 3.5 4 4.5 5 5.5 6 --frames 2e9 --min-errors 2000
echo "## config 3: 802.11n N648 R5/6 z27, shipped weights rows 0-19, harvest + the 50-row boosted post decoder"
$C --graph $F/BaseGraph/802_11n_N648_R56_z27.txt --z 27 --weights $F/Results/WIFI/Weights_Iter50.txt --iters 20 --snr 3 3.5 4 4.5 5 --frames 4e9 --min-errors 3000 --max-uncor 20000 --post-weights $F/Results/WIFI/Weights_Iter50.txt --post-iters 50
echo "## config 4: 5G NR R0.50 n1024 z64, shipped weights rows 0-19 (systematic), harvest + the 50-row boosted post decoder"
$C --graph $F/BaseGraph/5G_LDPC_R0.50_n_dec1280_n1024_k512_z64_s513_640.txt --weights "$F/Results/5G/5G_LDPC_R0.50_n_dec1280_n1024_k512_z64_s513_640_Weight_End50.txt" --iters 20 --systematic --snr 2.5 3 3.5 4 --frames 6e9 --min-errors 3000 --max-uncor 20000 --post-weights "$F/Results/5G/5G_LDPC_R0.50_n_dec1280_n1024_k512_z64_s513_640_Weight_End50.txt" --post-iters 50
} > gpurun_out/r02_campaigns_configs_2_3_4.txt 2>&1
rm -rf gpurun_out/files
grep -v "^$" gpurun_out/r02_campaigns_configs_2_3_4.txt | cut -c1-260
