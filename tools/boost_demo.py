#!/usr/bin/env python
"""The reference's whole boosting-learning workflow on one B200, WiMAX N576 R3/4:
  1. main_Base.py as collector (sampling_type 2): the shipped 20-iteration base decoder, never-corrected words -> ./Uncor.txt
  2. the split the authors do by hand -> Inputs/[Uncor]_{train, Valid, Test}
  3. main_Post.py: blocks [20,30) [30,40) [40,50) trained on the collected words (sampling_type 1, FER loss, Adam)
  4. the result against the authors' shipped 50-iteration weights (Results/WiMAX/Weights_Iter50.txt) on fresh failures
usage: python tools/boost_demo.py [outdir] [epochs-per-block] [collect-snr]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import materialize_files
import ldpc_error_floor_b200 as L
from ldpc_error_floor_b200 import campaign, drivers, formats, trainer

out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/boost_wimax"   # writes ~100 MB of word files
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
snr = float(sys.argv[3]) if len(sys.argv) > 3 else 3.5
made = materialize_files.materialize(out)
stem = "wman_N0576_R34_z24"
NTR, NVA, NTE = 10000, 2500, 2500

# 1. collection (the evaluation pass of main_Base.py with sampling_type = 2)
base_w = formats.read_weights(made["w:wimax_base20"])
t0 = time.time()
need, frames = NTR + NVA + NTE, 0
upath = os.path.join(out, "Uncor.txt")
if os.path.exists(upath):
    os.remove(upath)
k = 0
while True:
    cfg = drivers.RunConfig(root=out, sharing=[3, 3, 3], sampling_type=2, SNR_Matrix=np.array([snr]), valid_num=1 << 21, seed_in=2 + k)
    res = drivers.evaluate(cfg, weights=base_w, quiet=True)["valid"]
    frames += 1 << 21; k += 1
    n = sum(1 for _ in open(upath))
    if n >= need or k >= 12:
        break
print(f"collected {n} never-corrected words from {frames} frames at {snr} dB in {time.time() - t0:.1f} s (base FER {res[2, 0]:.3e})")
# 2. split
drivers.split_uncor(upath, stem, NTR, NVA, NTE, root=out)
# 3. post training, the reference's configuration (main_Post.py:25-63) with three blocks instead of one and fewer epochs
pcfg = drivers.RunConfig.post(root=out, iters_max=50, fixed_iter=20, iter_step=10, training_num=NTR, valid_num=NVA, test_num=NTE,
                              test_flag=0, epoch_input=epochs, learn_rate_start=0.003)
t0 = time.time()
blocks = trainer.train(pcfg, log=None)
print(f"trained {len(blocks)} blocks x {epochs} epochs x {NTR // pcfg.batch_size} batches of {pcfg.batch_size} words in {time.time() - t0:.1f} s")
for b in blocks:
    print(f"  block [{b.training_iter_start}, {b.training_iter_end}): validation FER_last per epoch",
          [f"{float(r[1, 0]):.3f}" for r in b.valid], " loss", [round(x, 4) for x in b.losses])
ours = os.path.join(out, "Weights", f"C0_{stem}_Opt_Weight_End50.txt")
# 4. fresh failures of the base decoder (another seed), decoded by: 50 iterations of plain continuation (weights 1.0 after
#    iteration 20), our trained post decoder, the authors' shipped post decoder
g = L.BaseGraph.from_file(made["wimax"], z=24)
base = L.NMSDecoder(g, base_w, iters=20)
shipped = formats.read_weights(made["w:wimax_boost50"])
trained = formats.read_weights(ours)
cont = formats.WeightSet([3, 3, 3], {i: np.vstack([base_w.blocks[i][:20], np.ones((30, 1), np.float32)]) for i in range(3)})
for name, ws in (("untrained continuation (weights 1.0)", cont), ("trained on this box", trained), ("shipped Weights_Iter50", shipped)):
    post = L.NMSDecoder(g, ws, iters=50)
    recs = campaign.run_campaign(base, [snr, snr + 0.5], 1 << 27, min_errors=20000, early_term=True, seed=777, harvest=True,
                                 max_uncor=20000, post_dec=post, post_iters=50)
    for r in recs:
        p = r["post"]
        print(f"{name:40s} Eb/N0 {r['snr_db']:.2f} dB: base FER {r['fer']:.3e} ({r['frames']} frames), post decoder leaves "
              f"{p['still_uncor_any']} of {p['words']} words -> FER {r['fer'] * p['still_uncor_any'] / max(p['words'], 1):.3e}")
