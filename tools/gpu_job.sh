mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/materialize_files.py /tmp/files > /dev/null
python -m ldpc_error_floor_b200.campaign --graph /tmp/files/BaseGraph/5G_LDPC_R0.73_n_dec2304_n2112_k1536_z72_s1537_1584.txt --z 72 --punct 1 144 --short 1537 1584 --ms-weight 0.8 --systematic --snr 5.5 --frames 6e7 --chunk 4194304 2>&1 | grep "Eb/N0"
