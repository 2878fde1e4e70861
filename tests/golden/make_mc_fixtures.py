"""Mints the reference-side Monte-Carlo fixtures mc_ref_<case>.npz (development container only).

    python tests/golden/make_mc_fixtures.py <case> [<case> ...]      # cases: see CASES; "all" = every case

Each fixture is a run of the reference's OWN Print_Functions.compute_results (:130-165), unmodified, exactly as
main_Base.py:177 calls it at epoch 0 (seeds 2042+2 / 1074+2, batch 20, sampling_type 0), one call per Eb/N0 point, with
Main_Functions.build_neural_network under the numpy stand-in for TensorFlow (oracle/ref_runner.py).  Besides the
Results[4] column the script records, per frame, what the reference's outputs say about it (bit errors of the last
iteration, "wrong at the last iteration", "never right at any iteration") -- observed from the ya_output_all tensor
the reference hands to calc_ber_fer (:100-118), with calc_ber_fer's own arithmetic -- so that the GPU tests can put
a frame-level confidence interval around the reference's BER as well as its FER.

The GPU box has no /root/reference: tests read only the committed .npz files.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import ref_runner  # noqa: E402
import make_golden as mg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# name: (graph key, weights (shipped key | ("const", sharing, cn)), T, decoding_type, q_bit, systematic, snr list, frames)
CASES = {
    # SURVEY.md Appendix A.3 anchor, regenerated and extended: WiMAX base decoder, shipped weights
    "wimax": ("wimax", "wimax_base20", 20, 2, 5, 0, [2.0, 2.5, 3.0, 3.5], 5000),
    # BASELINE config 3: 802.11n, rows 0-19 of the shipped 50-row file
    "wifi": ("wifi", "wifi_boost50", 20, 2, 5, 0, [3.0, 3.5], 2000),
    # BASELINE config 4: 5G R0.50 n1024 z64, rows 0-19 of the shipped file, systematic = 1
    "5g_r050_z64": ("5g_r050_z64", "5g_r050_z64_boost50", 20, 2, 5, 1, [1.5, 2.0], 1000),
    # BASELINE config 5, the campaign configuration: 5G R0.73 n2112 z72, plain 0.8 min-sum (no weights are shipped
    # for this graph), sharing [3, 0, 0], systematic = 1
    "5g_r073_z72": ("5g_r073_z72", ("const", [3, 0, 0], 0.8), 20, 2, 5, 1, [2.5, 3.0], 300),
    # float min-sum on WiMAX (decoding_type 1)
    "wimax_float": ("wimax", "wimax_base20", 20, 1, 5, 0, [2.5, 3.0], 2000),
}


class RecordingDecoder(ref_runner.ReferenceDecoder):
    """ReferenceDecoder whose fake session also keeps, per frame, what calc_ber_fer derives from the outputs."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.rec_biterr, self.rec_last, self.rec_any = [], [], []

    class _Sess(ref_runner.ReferenceDecoder._Sess):
        def run(self, fetches=None, feed_dict=None):
            out = super().run(fetches, feed_dict)
            ya = out[0] if isinstance(out, list) else out
            o = self.outer
            B = feed_dict["xa"].shape[0]
            hard = (np.asarray(ya) >= 0).reshape(o.T, B, -1)           # Print_Functions.py:106 with Y = 0
            wrong = hard.any(axis=2)                                   # [T, B]
            o.rec_biterr.append(hard[-1].sum(axis=1))
            o.rec_last.append(wrong[-1])
            o.rec_any.append(wrong.min(axis=0))                        # :109 uncor_flag = min over iterations
            return out


def run_point(args):
    name, k = args
    gkey, wsel, T, dt, qb, systematic, snrs, frames = CASES[name]
    stem, proto, z, punct, short = mg.graph_meta(gkey)
    M, N = proto.shape
    E = int((proto != -1).sum())
    if isinstance(wsel, tuple):
        sharing, weights = wsel[1], mg.const_weights(wsel[1], T, M, N, E, cn=wsel[2])
    else:
        sharing, weights = mg.shipped(wsel)
    t0 = time.time()
    rd = RecordingDecoder(proto.astype(int), z, sharing, weights, T, dt, qb, 20.0, punct, short, [snrs[k]],
                          target_node=N - M if systematic else None)
    res, _ = rd.compute_results(frames, 2044, 1076, 20, sampling_type=0)
    be = np.concatenate(rd.rec_biterr).astype(np.int32)
    print(f"{name} {snrs[k]} dB: {be.size} frames  BER_last {res[0, 0]:.4e} FER_last {res[1, 0]:.4f} FER {res[2, 0]:.4f}  "
          f"[{time.time() - t0:.0f} s]", flush=True)
    return res[:, 0], float(rd.snr_sigma[0]), be, np.concatenate(rd.rec_last), np.concatenate(rd.rec_any)


def run_case(name):
    import multiprocessing as mp
    gkey, wsel, T, dt, qb, systematic, snrs, frames = CASES[name]
    stem, proto, z, punct, short = mg.graph_meta(gkey)
    M, N = proto.shape
    E = int((proto != -1).sum())
    if isinstance(wsel, tuple):
        sharing = wsel[1]
        weights = mg.const_weights(sharing, T, M, N, E, cn=wsel[2])
    else:
        sharing, weights = mg.shipped(wsel)
    res_cols, sig, per = [], [], {}
    t0 = time.time()
    with mp.get_context("fork").Pool(len(snrs)) as pool:      # one process per Eb/N0 point (each call is independent)
        parts = pool.map(run_point, [(name, k) for k in range(len(snrs))])
    for k, (col, sg, be, last, anyf) in enumerate(parts):
        res_cols.append(col)
        sig.append(sg)
        per[f"biterr_{k}"], per[f"uncor_last_{k}"], per[f"uncor_any_{k}"] = be, last, anyf
    out = {"snr": np.array(snrs), "sigma": np.array(sig), "results": np.stack(res_cols, axis=1),
           "frames": np.array(frames), "batch": np.array(20), "seeds": np.array([2044, 1076]),
           "T": np.array(T), "decoding_type": np.array(dt), "q_bit": np.array(qb), "systematic": np.array(systematic),
           "sharing": np.array(sharing), "graph": np.array(gkey), "bits_counted": np.array((N - M if systematic else N) * z),
           "ber_divisor": np.array(N * z),       # calc_ber_fer divides by Y_test.shape[1] = N*z even when systematic
           "seconds": np.array(time.time() - t0)}
    for i in range(3):
        if sharing[i] > 0:
            out[f"w{i}"] = np.asarray(weights[i], dtype=np.float32)[:T]
    out.update(per)
    np.savez_compressed(os.path.join(OUT, f"mc_ref_{name}.npz"), **out)
    print(f"mc_ref_{name}.npz written ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    if not ref_runner.reference_available():
        raise SystemExit("needs /root/reference (development container only)")
    names = sys.argv[1:]
    if names == ["all"] or not names:
        names = list(CASES)
    for nm in names:
        run_case(nm)
