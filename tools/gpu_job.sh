mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/f32_sweep.py 2>&1 | tee gpurun_out/f32_sweep.txt
for g in "4 2" "4 3" "4 1" "8 2" "8 1" "2 4"; do set -- $g; LDPC_B200_FP=$1 LDPC_B200_R=$2 python tools/prof_one.py wimax 1 5 131072 2>&1 | tail -1; done | tee gpurun_out/f32_geom.txt
for g in "3 2" "4 2" "4 1"; do set -- $g; LDPC_B200_FP=$1 LDPC_B200_R=$2 python tools/prof_one.py 5g_r073_z72 1 5 32768 2>&1 | tail -1; done | tee -a gpurun_out/f32_geom.txt
for g in "2 2" "4 1" "4 2"; do set -- $g; LDPC_B200_FP=$1 LDPC_B200_R=$2 python tools/prof_one.py 5g_r050_z64 1 5 65536 2>&1 | tail -1; done | tee -a gpurun_out/f32_geom.txt
ncu --set full --clock-control none --import-source on -k regex:nms_f32 -c 1 -o gpurun_out/prof_f32b -f python tools/prof_one.py wimax 1 5 > gpurun_out/ncu_f32b.log 2>&1
tail -2 gpurun_out/ncu_f32b.log
