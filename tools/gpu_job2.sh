mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_i_n2.json 2> gpurun_out/bench_i_n2.err
tail -3 gpurun_out/bench_i_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_i_n2.json').read().strip().splitlines()[-1])
e=d['e2e']; print('N2 dev', d['frames_per_s']/1e6, 'e2e', e['frames_per_s']/1e6, 'f32only', e['float32_only']['frames_per_s']/1e6, 'page', d['e2e_pageable']['frames_per_s']/1e6, 'q8', d['e2e_q8']['frames_per_s']/1e6, 'ceiling', e['h2d_ceiling_gbs'])
print(json.dumps(e['host']))
P
nproc
