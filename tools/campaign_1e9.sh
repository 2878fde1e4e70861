#!/bin/bash
# BASELINE config 5 on one 8-GPU box: 5G NR R0.73 n2112 z72, normalised min-sum 0.8 (quantised, 20 iterations, systematic = 1)
# + the boosted 50-iteration post decoder trained on this box (tools/boost_z72.py), 1e11 frames at 5.5 dB, with a kill after
# ~75 s and a resume from the checkpoint.   usage: gpurun --gpus 8 -- bash tools/campaign_1e9.sh [frames] [snr] [ngpu]
FRAMES=${1:-1e11}; SNR=${2:-5.5}; NG=${3:-8}
mkdir -p gpurun_out /tmp/camp
python tools/materialize_files.py gpurun_out/files > /dev/null
G=gpurun_out/files/BaseGraph/5G_LDPC_R0.73_n_dec2304_n2112_k1536_z72_s1537_1584.txt
W=profiles/r02_5g_r073_z72_boosted_weights_End50.txt
OUT=gpurun_out/r02_campaign_5g_z72_1e-9
ARGS="--graph $G --ms-weight 0.8 --iters 20 --systematic --snr $SNR --frames $FRAMES --max-uncor 60000 --post-weights $W --post-iters 50 --checkpoint /tmp/camp/state.json --checkpoint-rounds 16 --json $OUT.json --survivors $OUT.survivors.q8 --seed 20261018"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
date +%s > $OUT.t0
timeout -s TERM 75 $RUN --master-port 29541 -m ldpc_error_floor_b200.campaign $ARGS > $OUT.phase1.txt 2>&1
echo "phase 1 exit $?" >> $OUT.phase1.txt
cp /tmp/camp/state.json $OUT.state_at_kill.json 2>/dev/null
sleep 2
$RUN --master-port 29542 -m ldpc_error_floor_b200.campaign $ARGS --resume > $OUT.phase2.txt 2>&1
echo "phase 2 exit $?" >> $OUT.phase2.txt
date +%s > $OUT.t1
cp /tmp/camp/state.json $OUT.state_final.json 2>/dev/null
nvidia-smi --query-gpu=index,clocks.sm,clocks_throttle_reasons.active,power.draw --format=csv > $OUT.smi.txt 2>&1
tail -3 $OUT.phase1.txt; tail -5 $OUT.phase2.txt
rm -rf gpurun_out/files
