# scratch job script for gpurun (edited per call)
mkdir -p gpurun_out
timeout 1500 python tools/boost_z72.py gpurun_out/boost_z72 4 4.0 8000 "5.0 6.0 7.0" 4e9 > gpurun_out/boost_z72.txt 2>&1
tail -30 gpurun_out/boost_z72.txt
rm -rf gpurun_out/boost_z72/Inputs gpurun_out/boost_z72/Uncor.txt
