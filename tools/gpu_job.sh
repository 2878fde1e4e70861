mkdir -p gpurun_out
python tools/train_demo.py gpurun_out/train_z72 6 2>&1 | tee gpurun_out/train_z72.txt
