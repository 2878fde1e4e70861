mkdir -p gpurun_out
SWEEP_SNR=4.5 SWEEP_FRAMES=131072 python tools/geom_sweep.py 5g_r073_z72 "3,2 1,2 1,4 2,2 2,3 3,3 3,1" 2>&1 | tee gpurun_out/geom_z72.txt
