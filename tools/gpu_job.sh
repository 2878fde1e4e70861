mkdir -p gpurun_out
timeout 500 python tools/boost_demo.py /tmp/boost_wimax 4 3.5 2>&1 | grep -v "^W\|Warning" | tee gpurun_out/boost_demo.txt
