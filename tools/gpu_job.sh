mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -k short_training 2>&1 | tail -30
python - <<'PY'
import sys, os, numpy as np, time
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import materialize_files
from ldpc_error_floor_b200 import drivers, trainer
root='gpurun_out/train_demo'; materialize_files.materialize(root)
cfg = drivers.RunConfig(root=root, sharing=[3,0,3], decoding_type=2, loss_type=2, etha_start=0.0, iters_max=10, iter_step=10,
                        batch_size=20, training_num=2000, valid_num=100000, SNR_Matrix=np.array([2.0,2.5,3.0,3.5,4.0]))
t=time.time(); res = trainer.train_block(cfg, 0, 10, epochs=3, log=None); print('shipped-style config (QMS, FER loss, batch 20): 3 epochs x 100 batches in', round(time.time()-t,1),'s')
print('losses', res.losses); print('valid FER_last per epoch', [r[1].round(4).tolist() for r in res.valid])
PY
