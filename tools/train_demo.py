#!/usr/bin/env python
"""Train a base decoder on the GPU for a graph that ships without weights (BASELINE config 5: 5G NR R0.73 n2112 z72),
then run a short FER campaign before / after.  usage: python tools/train_demo.py [outdir] [epochs]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import materialize_files
import ldpc_error_floor_b200 as L
from ldpc_error_floor_b200 import campaign, drivers, formats, trainer

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/train_z72"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
made = materialize_files.materialize(out)
stem = "5G_LDPC_R0.73_n_dec2304_n2112_k1536_z72_s1537_1584"
cfg = drivers.RunConfig(root=out, filename=stem, z_value=72, punct_start=1, punct_end=144, short_start=1537, short_end=1584,
                        sharing=[2, 0, 2], systematic=1, decoding_type=2, q_bit=5, loss_type=0, etha_start=1.0, iters_max=20, iter_step=20,
                        batch_size=40, training_num=8000, valid_num=200000, SNR_Matrix=np.array([2.5, 2.75, 3.0, 3.25]),
                        learn_rate_start=0.005, init_weight=0.8, init_VN_weight=1.0, opt_result_print=2)
t0 = time.time()
res = trainer.train_block(cfg, 0, 20, epochs=epochs, log=None)
print(f"trained {epochs} epochs x {cfg.training_num // cfg.batch_size} batches of {cfg.batch_size} frames in {time.time() - t0:.1f} s "
      f"(incl. {epochs + 1} validation passes of {cfg.valid_num} frames x {len(cfg.SNR_Matrix)} SNRs)")
print("training loss per epoch:", [round(x, 5) for x in res.losses])
for k, r in enumerate(res.valid):
    print(f"epoch {k}: valid FER {drivers.FTE(r[2])}  BER_last {drivers.FTE(r[0])}")
g = L.BaseGraph.from_file(made["5g_r073_z72"])
wfile = os.path.join(out, "Weights", f"C0_{stem}_Opt_Weight_End20.txt")
print("best weights:", wfile, os.path.exists(wfile))
snrs = [2.5, 3.0, 3.5, 4.0]
plain = L.NMSDecoder(g, formats.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)}), iters=20, systematic=1)
tuned = L.NMSDecoder(g, formats.read_weights(wfile), iters=20, systematic=1)
for name, dec in (("NMS 0.8 (no trained weights)", plain), ("trained on this box", tuned)):
    recs = campaign.run_campaign(dec, snrs, 1 << 25, min_errors=300, early_term=True, seed=99)
    print(name)
    for r in recs:
        print("   " + campaign._fmt(r))
