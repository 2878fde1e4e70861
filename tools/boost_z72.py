#!/usr/bin/env python
"""BASELINE config 5 end to end on one B200: the reference's boosting-learning workflow for 5G NR R0.73 n2112 z72, a graph the
reference ships no weights for.
  1. base decoder = normalised min-sum 0.8, 20 iterations, written in the reference's per-node file form (header "2 2 2": CN = UCN =
     0.8 per check, VN = 1.0 per column -- the structure of the shipped Results/5G files' rows 0-19)
  2. main_Base.py as collector (sampling_type 2 semantics, systematic = 1): never-corrected words at <collect-snr> -> ./Uncor.txt,
     split into Inputs/[Uncor]_{train, Valid, Test}
  3. main_Post.py: blocks [20,30) [30,40) [40,50) trained on those words (sampling_type 1, FER loss, Adam, per-node weights)
  4. fresh failures of the base decoder at the campaign's Eb/N0 points, re-decoded by the trained 50-iteration decoder and by an
     untrained continuation (weights 1.0): what the boosted decoder does to the error floor
usage: python tools/boost_z72.py [outdir] [epochs-per-block] [collect-snr] [words] [eval-snrs] [eval-frames]"""
import os, shutil, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import materialize_files
import ldpc_error_floor_b200 as L
from ldpc_error_floor_b200 import campaign, drivers, formats, trainer

out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/boost_z72"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
snr_c = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
need = int(sys.argv[4]) if len(sys.argv) > 4 else 8000
eval_snrs = [float(v) for v in (sys.argv[5].split() if len(sys.argv) > 5 else "5.0 6.0".split())]
eval_frames = int(float(sys.argv[6])) if len(sys.argv) > 6 else 1 << 31
made = materialize_files.materialize(out)
stem = "5G_LDPC_R0.73_n_dec2304_n2112_k1536_z72_s1537_1584"
g = L.BaseGraph.from_file(made["5g_r073_z72"])
NTR, NVA = int(need * 0.75), need - int(need * 0.75)

# 1. base decoder in file form
base_file = formats.WeightSet([2, 2, 2], {0: np.full((20, g.M), 0.8, np.float32), 1: np.full((20, g.M), 0.8, np.float32),
                                          2: np.ones((20, g.N), np.float32)})
os.makedirs(os.path.join(out, "Weights"), exist_ok=True)
formats.write_weights(os.path.join(out, "Weights", f"C0_{stem}_Opt_Weight_End20.txt"), base_file)
base = L.NMSDecoder(g, formats.WeightSet([3, 0, 0], {0: np.full((20, 1), 0.8, np.float32)}), iters=20, systematic=1)
chk = L.NMSDecoder(g, base_file, iters=20, systematic=1)
sg = float(g.sigma([3.0])[0])
a, _ = base.mc_run_host(sg, 20000, seed=1, early_term=True)
b, _ = chk.mc_run_host(sg, 20000, seed=1, early_term=True)
assert a == b, "the per-node file form of the base decoder must decode like the scalar 0.8 decoder"
print("base decoder:", base.mc_info())

# 2. collection
t0 = time.time()
recs = campaign.run_campaign(base, [snr_c], 1 << 34, min_errors=need, early_term=True, seed=4242, harvest=True, max_uncor=need)
r = recs[0]
rows = r["rows"][:need]
print(f"collected {rows.shape[0]} never-corrected words from {r['frames']} frames at {snr_c} dB in {time.time() - t0:.1f} s "
      f"(base FER {r['fer']:.3e}, {r['frames_per_s'] / 1e6:.1f} Mframes/s)", flush=True)
upath = os.path.join(out, "Uncor.txt")
if os.path.exists(upath):
    os.remove(upath)
formats.append_uncor(upath, rows)
drivers.split_uncor(upath, stem, NTR, NVA, 0, root=out)

# 3. post training (main_Post.py's configuration with three blocks and per-node weights)
pcfg = drivers.RunConfig.post(root=out, filename=stem, z_value=72, punct_start=1, punct_end=144, short_start=1537, short_end=1584,
                              sharing=[2, 2, 2], systematic=1, iters_max=50, fixed_iter=20, iter_step=10, training_num=NTR,
                              valid_num=NVA, test_num=0, test_flag=0, epoch_input=epochs, learn_rate_start=0.003)
t0 = time.time()
blocks = trainer.train(pcfg, log=None)
print(f"trained {len(blocks)} blocks x {epochs} epochs x {NTR // pcfg.batch_size} batches of {pcfg.batch_size} words in {time.time() - t0:.1f} s")
for bl in blocks:
    print(f"  block [{bl.training_iter_start}, {bl.training_iter_end}): validation FER_last per epoch",
          [f"{float(v[1, 0]):.3f}" for v in bl.valid], " loss", [round(x, 4) for x in bl.losses], flush=True)
ours = os.path.join(out, "Weights", f"C0_{stem}_Opt_Weight_End50.txt")
keep = os.path.join(os.path.dirname(out.rstrip("/")), f"{stem}_Weight_End50_trained_on_b200.txt")
shutil.copy(ours, keep)
print("boosted weights:", keep)

# 4. fresh failures at the campaign points
trained = formats.read_weights(ours)
cont = formats.WeightSet([2, 2, 2], {i: np.vstack([base_file.blocks[i], np.ones((30, base_file.blocks[i].shape[1]), np.float32)]) for i in range(3)})
for name, ws in (("untrained continuation (weights 1.0)", cont), ("trained on this box", trained)):
    post = L.NMSDecoder(g, ws, iters=50, systematic=1)
    recs = campaign.run_campaign(base, eval_snrs, eval_frames, min_errors=3000, early_term=True, seed=777, harvest=True,
                                 max_uncor=20000, post_dec=post, post_iters=50)
    for r in recs:
        print(f"{name:38s} " + campaign._fmt(r), flush=True)
