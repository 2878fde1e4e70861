"""GPU: the training step ("next" row N1) -- loss and weight gradients from the CUDA kernel against the goldens
minted by torch.autograd of the reference's own loss + forward code, and a short training run that lowers the loss."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from test_oracle_grad import CASES, check_grads, load_grad_case

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("name", CASES)
def test_train_grad_matches_reference_autograd(name):
    import torch
    import ldpc_error_floor_b200 as L
    c = load_grad_case(name)
    g = L.BaseGraph(c["proto"], c["z"], c["punct"], c["short"])
    ws = L.WeightSet(c["sharing"], dict(c["weights"]))
    dec = L.NMSDecoder(g, ws, iters=c["T"], decoding_type=c["decoding_type"], q_bit=c["q_bit"], clip_llr=c["clip"],
                       systematic=1 if c["target_node"] else 0)
    xa = torch.from_numpy(c["xa"]).cuda()
    loss, grads, app = dec.train_grad(xa, iter_lo=c["t_lo"], loss_type=c["loss_type"], etha=c["etha"], want_app=True)
    assert loss == pytest.approx(c["loss"], rel=2e-5, abs=1e-7)
    check_grads(grads, c, rtol=5e-4)
    # the training kernel's forward is the decoder the fast kernels run
    ref = dec.decode(xa, app="all").app
    if c["decoding_type"] == 2:
        assert torch.equal(app, ref)
    else:
        assert float((app - ref).abs().max()) <= 1e-4 and torch.equal(app >= 0, ref >= 0)


def test_set_weights_and_a_short_training_run(tmp_path):
    """A few dozen Adam steps on WiMAX starting from plain min-sum (all weights 1.0), float messages, cross-entropy
    on the last iteration: the loss and the validation BER go down (normalised min-sum beats plain min-sum), weights
    stay inside [Min_weight, Max_weight], frozen iterations do not move, the weight file is what print_weight writes
    and the fast decoder sees the new weights."""
    import torch
    import materialize_files
    from ldpc_error_floor_b200 import drivers, formats, trainer
    root = str(tmp_path)
    materialize_files.materialize(root)
    cfg = drivers.RunConfig(root=root, sharing=[3, 3, 3], decoding_type=1, loss_type=0, etha_start=0.0, iters_max=6,
                            iter_step=6, batch_size=128, training_num=128 * 16, valid_num=8192,
                            SNR_Matrix=np.array([2.0, 2.5, 3.0]), learn_rate_start=0.03, init_weight=1.0, init_VN_weight=1.0)
    res = trainer.train_block(cfg, 0, 6, epochs=3, log=None)
    assert len(res.losses) == 4 and res.losses[0] == 0.0
    assert res.losses[3] < res.losses[1]                                   # training lowers the loss
    assert res.valid[-1][0].sum() < res.valid[0][0].sum()                  # and the validation BER_last
    w = formats.read_weights(os.path.join(root, "Weights", "C0_wman_N0576_R34_z24_Weight_End6.txt"))
    assert w.sharing == [3, 3, 3] and w.iterations == 6
    for b in w.blocks.values():
        assert b.min() >= 0.0 and b.max() <= 2.0 and not np.allclose(b, 1.0)
    perf = open(cfg.perf_filename).read()
    assert perf.count("Valid_Result") == 4 and "Training loss:" in perf
    assert os.path.exists(os.path.join(root, "Weights", "C0_wman_N0576_R34_z24_Opt_Weight_End6.txt"))
    # second block: iterations 0..5 frozen at the first block's best weights
    cfg2 = drivers.RunConfig(root=root, sharing=[3, 3, 3], decoding_type=1, loss_type=0, iters_max=8, fixed_iter=6, iter_step=2,
                             batch_size=64, training_num=64 * 4, valid_num=2048, SNR_Matrix=np.array([2.5]),
                             learn_rate_start=0.02, init_weight=0.5)
    res2 = trainer.train_block(cfg2, 6, 8, epochs=1, log=None)
    best = formats.read_weights(os.path.join(root, "Weights", "C0_wman_N0576_R34_z24_Opt_Weight_End6.txt"))
    for i in range(3):
        assert np.array_equal(res2.weights.blocks[i][:6], best.blocks[i])
        assert not np.array_equal(res2.weights.blocks[i][6:], np.full_like(res2.weights.blocks[i][6:], 0.5))


def test_temporal_sharing_training_keeps_the_shared_rows_tied(tmp_path):
    """sharing code 4 (main_Base.py:24, weight_init :411-414): iterations >= fixed_iter share ONE per-edge variable.
    After a few Adam steps the rows 3..5 are still equal to each other, have moved, rows 0..2 (below t_lo) have not, and
    the weight file carries the run's header with the T expanded rows (print_weight :87-94)."""
    import materialize_files
    from ldpc_error_floor_b200 import drivers, formats, trainer
    root = str(tmp_path)
    materialize_files.materialize(root)
    cfg = drivers.RunConfig(root=root, sharing=[4, 0, 2], decoding_type=2, q_bit=5, loss_type=0, etha_start=0.0, iters_max=6,
                            fixed_iter=3, iter_step=3, batch_size=64, training_num=64 * 6, valid_num=2048, valid_flag=1,
                            SNR_Matrix=np.array([2.5, 3.0]), learn_rate_start=0.02, init_weight=0.9, init_VN_weight=1.0)
    E = 88
    start = formats.WeightSet([4, 0, 2], {0: np.full((6, E), 0.9, np.float32), 2: np.ones((6, 24), np.float32)})
    res = trainer.train_block(cfg, 3, 6, init=start, epochs=2, log=None)
    w = res.weights
    assert w.sharing[0] == 1 and w.blocks[0].shape == (6, E)          # the decoder's per-iteration form
    assert np.array_equal(w.blocks[0][3], w.blocks[0][4]) and np.array_equal(w.blocks[0][3], w.blocks[0][5])
    # the dump of the last epoch (res.weights is the best-on-validation copy, possibly the untrained epoch 0)
    f = formats.read_weights(os.path.join(root, "Weights", "C0_wman_N0576_R34_z24_Weight_End6.txt"))
    assert list(f.sharing) == [4, 0, 2] and f.blocks[0].shape == (6, E)
    assert np.array_equal(f.blocks[0][3], f.blocks[0][4]) and np.array_equal(f.blocks[0][3], f.blocks[0][5])
    assert np.abs(f.blocks[0][3] - 0.9).max() > 1e-3                  # the shared variable was trained ...
    assert np.array_equal(f.blocks[0][:3], np.full((3, E), 0.9, np.float32))   # ... the iterations below t_lo were not
    assert np.abs(f.blocks[2][3:] - 1.0).max() > 1e-4 and not np.array_equal(f.blocks[2][3], f.blocks[2][5])   # VN rows: untied
