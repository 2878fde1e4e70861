mkdir -p gpurun_out
python tools/materialize_files.py gpurun_out/files > /dev/null
F=gpurun_out/files
# config 2: MacKay N96 K48 (z = 1, CSR path), plain and normalised min-sum, Eb/N0 1..6 dB, stop at 100 frame errors
for w in 1.0 0.8; do
  echo "## MacKay N96 K48, min-sum weight $w (quantised, q_bit 5)"
  python -m ldpc_error_floor_b200.campaign --graph $F/BaseGraph/MACKAY_N96_K48.txt --z 1 --ms-weight $w --snr 1 1.5 2 2.5 3 3.5 4 4.5 5 5.5 6 --frames 2e9 --min-errors 100 --chunk 8388608 --json gpurun_out/camp_mackay_$w.json 2>&1 | grep "Eb/N0\|^#"
done | tee gpurun_out/camp_mackay.txt
echo "## MacKay N96 K48, min-sum weight 0.8, float messages" | tee -a gpurun_out/camp_mackay.txt
python -m ldpc_error_floor_b200.campaign --graph $F/BaseGraph/MACKAY_N96_K48.txt --z 1 --ms-weight 0.8 --decoding-type 1 --snr 1 2 3 4 5 6 --frames 1e9 --min-errors 100 --chunk 8388608 2>&1 | grep "Eb/N0\|^#" | tee -a gpurun_out/camp_mackay.txt
# config 3: 802.11n N648 R5/6 z27, shipped 50-row weights: rows 0-19 base + all 50 rows as post decoder on the harvested words
W=$F/Results/WIFI/Weights_Iter50.txt
python -m ldpc_error_floor_b200.campaign --graph $F/BaseGraph/802_11n_N648_R56_z27.txt --z 27 --weights $W --iters 20 --post-weights $W --post-iters 50 --snr 3.0 3.5 4.0 4.5 5.0 --frames 2e9 --min-errors 300 --max-uncor 20000 --harvest gpurun_out/uncor_wifi.txt --json gpurun_out/camp_wifi_post.json 2>&1 | grep "Eb/N0\|^#" | tee gpurun_out/camp_wifi_post.txt
wc -l gpurun_out/uncor_wifi.txt | tee -a gpurun_out/camp_wifi_post.txt; rm -f gpurun_out/uncor_wifi.txt
