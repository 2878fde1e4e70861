mkdir -p gpurun_out
python tools/prof_one.py 5g_r073_z72 2 5 65536 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:nms_h2_spec_5g -c 1 -o gpurun_out/prof_z72 -f python tools/prof_one.py 5g_r073_z72 2 5 65536 > gpurun_out/ncu_z72.log 2>&1
tail -1 gpurun_out/ncu_z72.log
