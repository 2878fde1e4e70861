"""Mints the committed golden fixtures from the reference's own code (Oracle A).

Run ONLY in the development container (needs /root/reference, read-only):
    python tests/golden/make_golden.py
It imports /root/reference/Main_Functions.py and Print_Functions.py UNMODIFIED on top of the
numpy stand-in for tensorflow.compat.v1 (oracle/tf_shim) -- see oracle/ref_runner.py -- and writes

  codes.npz          every shipped BaseGraph/*.txt proto matrix, every shipped weight table
                     (Weights/*.txt, Results/**.txt) and the init_parameter scalars per graph
  decode_<case>.npz  channel LLRs from Print_Functions.create_mix_epoch + the per-iteration APP
                     tensors build_neural_network produces for them (ya_output{t})
  mc_wimax.npz       a small Print_Functions.compute_results run (Results[4,nSNR]) and the
                     Uncor.txt it appended (format F3)

The GPU box has no /root/reference; tests and bench.py read only these fixtures.
"""
import glob
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ldpc_error_floor_b200 import formats  # noqa: E402
from oracle import ref_runner  # noqa: E402

REF = ref_runner.REFERENCE_ROOT
OUT = os.path.dirname(os.path.abspath(__file__))

GRAPHS = {
    # key: (file stem, z, punct, short)
    "wimax": ("wman_N0576_R34_z24", 24, (0, 0), (0, 0)),
    "wifi": ("802_11n_N648_R56_z27", 27, (0, 0), (0, 0)),
    "mackay": ("MACKAY_N96_K48", 1, (0, 0), (0, 0)),
    "bch": ("BCH_63_51", 1, (0, 0), (0, 0)),
    "polar": ("Polar_64_48", 1, (0, 0), (0, 0)),
    "5g_r033_z32": ("5G_LDPC_R0.33_n_dec896_n768_k256_z32_s257_320", None, None, None),
    "5g_r050_z32": ("5G_LDPC_R0.50_n_dec640_n512_k256_z32_s257_320", None, None, None),
    "5g_r050_z64": ("5G_LDPC_R0.50_n_dec1280_n1024_k512_z64_s513_640", None, None, None),
    "5g_r073_z32": ("5G_LDPC_R0.73_n_dec480_n352_k256_z32_s257_320", None, None, None),
    "5g_r073_z72": ("5G_LDPC_R0.73_n_dec2304_n2112_k1536_z72_s1537_1584", None, None, None),
}
WEIGHTS = {
    "wimax_base20": "Weights/C0_wman_N0576_R34_z24_Opt_Weight_End20.txt",
    "wimax_boost50": "Results/WiMAX/Weights_Iter50.txt",
    "wifi_boost50": "Results/WIFI/Weights_Iter50.txt",
    "5g_r033_z32_boost50": "Results/5G/5G_LDPC_R0.33_n_dec896_n768_k256_z32_s257_320_Weight_End50.txt",
    "5g_r050_z32_boost50": "Results/5G/5G_LDPC_R0.50_n_dec640_n512_k256_z32_s257_320_Weight_End50.txt",
    "5g_r050_z64_boost50": "Results/5G/5G_LDPC_R0.50_n_dec1280_n1024_k512_z64_s513_640_Weight_End50.txt",
    "5g_r073_z32_boost50": "Results/5G/5G_LDPC_R0.73_n_dec480_n352_k256_z32_s257_320_Weight_End50.txt",
}


ONLY = []


def graph_meta(key):
    stem, z, punct, short = GRAPHS[key]
    path = os.path.join(REF, "BaseGraph", stem + ".txt")
    proto = formats.read_base_graph(path)
    if z is None:
        m = formats.parse_5g_name(stem)
        z, punct, short = m["z"], m["punct"], m["short"]
    return stem, proto, z, punct, short


def make_codes():
    mf, _ = ref_runner.load_reference()
    out = {}
    snr = np.array([1.0, 2.0, 2.5, 3.0, 3.5, 4.0, 5.0])
    for key in GRAPHS:
        stem, proto, z, punct, short = graph_meta(key)
        raw = open(os.path.join(REF, "BaseGraph", stem + ".txt"), "rb").read()
        M, N, base, cn, vn, E, rate, sigma = mf.init_parameter(proto.astype(int), snr, z, punct[0], punct[1],
                                                               short[0], short[1])
        out[f"graph/{key}/proto"] = proto.astype(np.int16)
        out[f"graph/{key}/meta"] = np.array([z, punct[0], punct[1], short[0], short[1], int(E)], dtype=np.int64)
        out[f"graph/{key}/stem"] = np.array(stem)
        out[f"graph/{key}/crlf"] = np.array(b"\r\n" in raw)
        out[f"graph/{key}/rate_ref"] = np.array(rate, dtype=np.float64)
        out[f"graph/{key}/sigma_ref"] = np.asarray(sigma, dtype=np.float64)
        out[f"graph/{key}/cn_deg"] = np.asarray(cn, dtype=np.int64)
        out[f"graph/{key}/vn_deg"] = np.asarray(vn, dtype=np.int64)
    out["snr_grid"] = snr
    for key, rel in WEIGHTS.items():
        ws = formats.read_weights(os.path.join(REF, rel))
        out[f"weights/{key}/sharing"] = np.array(ws.sharing, dtype=np.int64)
        for i, b in ws.blocks.items():
            out[f"weights/{key}/block{i}"] = b
        out[f"weights/{key}/text"] = np.array(open(os.path.join(REF, rel), "r").read())
    np.savez_compressed(os.path.join(OUT, "codes.npz"), **out)
    print("codes.npz", len(out), "arrays")


def ref_llrs(pf, sigmas, B, N, z, decoding_type, punct, short, q_bit, clip, seed):
    word = np.random.RandomState(2042 + seed)      # main_Base.py:71-74
    noise = np.random.RandomState(1074 + seed)
    X, _ = pf.create_mix_epoch(np.asarray(sigmas), word, noise, B, N, N, z, [], True, decoding_type,
                               punct[0], punct[1], short[0], short[1], q_bit, clip)
    return np.asarray(X, dtype=np.float32)


def make_decode_case(name, gkey, sharing, weights, T, decoding_type, q_bit, B, snr_db, seed=2, clip=20.0,
                     raw_llr=False, target_node=None, fixed_iter=0):
    if ONLY and not any(o in name for o in ONLY):
        return
    stem, proto, z, punct, short = graph_meta(gkey)
    rd = ref_runner.ReferenceDecoder(proto.astype(int), z, sharing, weights, T, decoding_type, q_bit, clip,
                                     punct, short, snr_db, target_node=target_node, fixed_iter=fixed_iter)
    # raw_llr: feed UNQUANTISED channel LLRs to the quantised decoder (legal in the reference: the
    # graph quantises its input itself, Main_Functions.py:176-177, 321-322)
    X = ref_llrs(rd.pf, rd.snr_sigma, B, rd.N, z, 1 if raw_llr else decoding_type, punct, short, q_bit, clip, seed)
    res = rd.decode(X)
    out = {"proto": proto.astype(np.int16), "meta": np.array([z, punct[0], punct[1], short[0], short[1]]),
           "sharing": np.array(sharing), "T": np.array(T), "decoding_type": np.array(decoding_type),
           "q_bit": np.array(q_bit), "clip": np.array(clip, dtype=np.float32), "xa": X,
           "app": res["app"].astype(np.float32), "sigma": np.asarray(rd.snr_sigma)}
    for i in range(3):
        if sharing[i] > 0:
            out[f"w{i}"] = np.asarray(weights[i], dtype=np.float32)[:T]
    if target_node is not None:
        # systematic = 1 (main_Base.py:83-84): ya_output_all holds the first target_node columns only and
        # calc_ber_fer (Print_Functions.py:100-118) counts over those
        ya_all = res["ya_output_all"]
        Y = np.zeros((B, rd.N * z), dtype=np.int64)
        ber_last, fer_last, fer, uncor, error_num = rd.pf.calc_ber_fer(ya_all, T, Y, B)
        out.update(target_node=np.array(target_node), ya_output_all=ya_all.astype(np.float32),
                   uncor_flag=np.asarray(uncor), error_num=np.asarray(error_num),
                   metrics=np.array([ber_last, fer_last, fer]))
    if fixed_iter:
        out["fixed_iter"] = np.array(fixed_iter)
    np.savez_compressed(os.path.join(OUT, f"decode_{name}.npz"), **out)
    hard_fail = ((res["app"][-1] >= 0).sum(axis=1) > 0).sum()
    print(f"decode_{name}.npz  B={B} T={T}  frames still wrong at the end: {hard_fail}")


def shipped(key):
    ws = formats.read_weights(os.path.join(REF, WEIGHTS[key]))
    return ws.sharing, ws.blocks


def const_weights(sharing, T, M, N, E, cn=0.8, ucn=0.6, vn=1.0, rng=None):
    width = lambda code, kind: {0: 0, 1: E, 2: (N if kind == 2 else M), 3: 1}[code]
    vals = (cn, ucn, vn)
    blocks = {}
    for i in range(3):
        if sharing[i] > 0:
            w = np.full((T, width(sharing[i], i)), vals[i], dtype=np.float32)
            if rng is not None:
                w = (w + rng.uniform(-0.25, 0.25, size=w.shape)).astype(np.float32)
            blocks[i] = w
    return blocks


def make_decode_cases():
    sh, w = shipped("wimax_base20")
    make_decode_case("wimax_qms_333_t20", "wimax", sh, w, 20, 2, 5, 8, [2.5, 3.0, 3.5, 4.0])
    make_decode_case("wimax_float_333_t20", "wimax", sh, w, 20, 1, 5, 8, [2.5, 3.0, 3.5, 4.0])
    make_decode_case("wimax_qms_303_t20", "wimax", [3, 0, 3], {0: w[0], 2: w[2]}, 20, 2, 5, 6, [3.0, 3.5])
    make_decode_case("wimax_qmsraw_333_t20", "wimax", sh, w, 20, 2, 5, 6, [3.0, 3.5], raw_llr=True)
    sh, w = shipped("wimax_boost50")
    make_decode_case("wimax_qms_333_t50", "wimax", sh, w, 50, 2, 5, 6, [2.5, 3.0])
    sh, w = shipped("wifi_boost50")
    make_decode_case("wifi_qms_333_t50", "wifi", sh, w, 50, 2, 5, 6, [3.0, 3.5, 4.0])
    sh, w = shipped("5g_r050_z64_boost50")
    make_decode_case("5g_r050_z64_qms_222_t50", "5g_r050_z64", sh, w, 50, 2, 5, 4, [1.0, 2.0])
    sh, w = shipped("5g_r073_z32_boost50")
    make_decode_case("5g_r073_z32_qms_222_t50", "5g_r073_z32", sh, w, 50, 2, 5, 4, [3.0, 4.0])
    make_decode_case("5g_r073_z32_float_222_t50", "5g_r073_z32", sh, w, 50, 1, 5, 4, [3.0, 4.0])
    sh, w = shipped("5g_r033_z32_boost50")
    make_decode_case("5g_r033_z32_qms_222_t20", "5g_r033_z32", sh, w, 20, 2, 5, 4, [0.0, 1.0])
    # sum-product (decoding_type 0, Main_Functions.py:238-245; punctured LLRs are 0.001, Print_Functions.py:53-55)
    sh, w = shipped("wimax_base20")
    make_decode_case("wimax_sp_333_t10", "wimax", sh, w, 10, 0, 5, 8, [2.0, 2.5, 3.0, 3.5])
    sh, w = shipped("5g_r073_z32_boost50")
    make_decode_case("5g_r073_z32_sp_222_t12", "5g_r073_z32", sh, w, 12, 0, 5, 6, [2.0, 3.0, 4.0])
    rng = np.random.RandomState(7)
    # no weights are shipped for these graphs: synthetic ones
    _, p, z, _, _ = graph_meta("5g_r073_z72")
    M, N = p.shape
    E = int((p != -1).sum())
    make_decode_case("5g_r073_z72_qms_222_t8", "5g_r073_z72", [2, 2, 2],
                     const_weights([2, 2, 2], 8, M, N, E, rng=rng), 8, 2, 5, 2, [3.0, 4.0])
    # BASELINE config 5 as the campaign runs it: plain 0.8 min-sum, 20 iterations, systematic = 1
    make_decode_case("5g_r073_z72_qms_300_t20_sys", "5g_r073_z72", [3, 0, 0], const_weights([3, 0, 0], 20, M, N, E), 20, 2, 5,
                     6, [2.5, 3.0, 3.5], target_node=N - M)
    _, p, z, _, _ = graph_meta("mackay")
    M, N = p.shape
    E = int((p != -1).sum())
    make_decode_case("mackay_float_300_t20", "mackay", [3, 0, 0], const_weights([3, 0, 0], 20, M, N, E), 20, 1, 5,
                     16, [2.0, 3.0, 4.0, 5.0])
    make_decode_case("mackay_qms_300_t20", "mackay", [3, 0, 0], const_weights([3, 0, 0], 20, M, N, E), 20, 2, 5,
                     16, [2.0, 3.0, 4.0, 5.0])
    _, p, z, _, _ = graph_meta("bch")
    M, N = p.shape
    E = int((p != -1).sum())
    make_decode_case("bch_qms_333_t10", "bch", [3, 3, 3], const_weights([3, 3, 3], 10, M, N, E, rng=rng), 10, 2, 5,
                     8, [3.0, 5.0])
    _, p, z, _, _ = graph_meta("polar")
    M, N = p.shape
    E = int((p != -1).sum())
    make_decode_case("polar_qms_223_t6", "polar", [2, 2, 3], const_weights([2, 2, 3], 6, M, N, E, rng=rng), 6, 2, 5,
                     8, [3.0, 5.0])
    make_decode_case("polar_float_300_t6", "polar", [3, 0, 0], const_weights([3, 0, 0], 6, M, N, E), 6, 1, 5,
                     8, [3.0, 5.0])
    # per-edge weights (sharing code 1) and the other quantisers, on WiMAX
    _, p, z, _, _ = graph_meta("wimax")
    M, N = p.shape
    E = int((p != -1).sum())
    make_decode_case("wimax_qms_112_t6", "wimax", [1, 1, 2], const_weights([1, 1, 2], 6, M, N, E, rng=rng), 6, 2, 5,
                     4, [3.0, 3.5])
    make_decode_case("wimax_float_102_t6", "wimax", [1, 0, 2], const_weights([1, 0, 2], 6, M, N, E, rng=rng), 6, 1,
                     5, 4, [3.0, 3.5])
    for qb in (6, -5, 4, 3):
        make_decode_case(f"wimax_qms_q{qb}_323_t6".replace("-", "m"), "wimax", [3, 3, 3],
                         const_weights([3, 3, 3], 6, M, N, E, rng=rng), 6, 2, qb, 4, [3.0, 3.5])
    make_decode_case("wimax_qms_000_t5", "wimax", [0, 0, 0], {}, 5, 2, 5, 4, [3.0, 3.5])
    # "next" row N3: temporal sharing (code 4: per-edge CN weights, rows >= fixed_iter reuse row fixed_iter,
    # Main_Functions.py:299-304) and the systematic metric (target_node = N - M, main_Base.py:83-84)
    stem, proto, z, punct, short = graph_meta("wimax")
    M, N = proto.shape
    E = int((proto != -1).sum())
    rng = np.random.RandomState(11)
    w4 = const_weights([1, 0, 2], 4, M, N, E, rng=rng)          # var_0_0..var_0_3 (fixed_iter = 3), var_2_t for t < 8
    w4[2] = const_weights([1, 0, 2], 8, M, N, E, rng=rng)[2]
    make_decode_case("wimax_qms_402_t8_fixed3", "wimax", [4, 0, 2], w4, 8, 2, 5, 6, [3.0, 3.5], fixed_iter=3)
    sh, w = shipped("5g_r050_z64_boost50")
    stem, proto, z, punct, short = graph_meta("5g_r050_z64")
    M, N = proto.shape
    make_decode_case("5g_r050_z64_qms_222_t12_sys", "5g_r050_z64", sh, w, 12, 2, 5, 12, [0.5, 1.5, 2.5], target_node=N - M)


def lifted_H(proto, z):
    M, N = proto.shape
    H = np.zeros((M * z, N * z), dtype=np.uint8)
    for i in range(M):
        for j in range(N):
            if proto[i, j] != -1:
                s_ = int(proto[i, j]) % z
                for a in range(z):
                    H[i * z + a, j * z + (a + s_) % z] = 1          # check (i, a) -- variable (j, (a + s) mod z)
    return H


def generator_matrix(H, k):
    """k rows of a basis of the null space of H over GF(2): a code_GM for create_mix_epoch (Print_Functions.py:41-42)."""
    Hm = H.copy() % 2
    m, n = Hm.shape
    piv, r = [], 0
    for c in range(n):
        rows = np.nonzero(Hm[r:, c])[0]
        if rows.size == 0:
            continue
        Hm[[r, r + rows[0]]] = Hm[[r + rows[0], r]]
        for rr in np.nonzero(Hm[:, c])[0]:
            if rr != r:
                Hm[rr] ^= Hm[r]
        piv.append(c)
        r += 1
        if r == m:
            break
    free = [c for c in range(n) if c not in piv]
    G = np.zeros((len(free), n), dtype=np.int64)
    for gi, fc in enumerate(free):
        G[gi, fc] = 1
        for ri, pc in enumerate(piv):
            G[gi, pc] = Hm[ri, fc]
    assert not ((G @ H.T) % 2).any()
    return G[:k]


def make_cw_case(name, gkey, sharing, weights, T, decoding_type, q_bit, B, snr_db, seed=5, clip=20.0):
    """Non-zero codewords ("next" row N3): create_mix_epoch with is_zeros_word = False and a generator matrix, the decode, and
    calc_ber_fer against Y -- all by the reference's own code."""
    if ONLY and not any(o in name for o in ONLY):
        return
    stem, proto, z, punct, short = graph_meta(gkey)
    rd = ref_runner.ReferenceDecoder(proto.astype(int), z, sharing, weights, T, decoding_type, q_bit, clip, punct, short, snr_db)
    M, N = proto.shape
    GM = generator_matrix(lifted_H(proto, z), (N - M) * z)
    word = np.random.RandomState(2042 + seed)
    noise = np.random.RandomState(1074 + seed)
    X, Y = rd.pf.create_mix_epoch(np.asarray(rd.snr_sigma), word, noise, B, N, N - M, z, GM, False, decoding_type,
                                  punct[0], punct[1], short[0], short[1], q_bit, clip)
    X = np.asarray(X, dtype=np.float32)
    res = rd.decode(X, ya=Y.astype(np.float32))
    ber_last, fer_last, fer, uncor, error_num = rd.pf.calc_ber_fer(res["ya_output_all"], T, Y, B)
    out = {"proto": proto.astype(np.int16), "meta": np.array([z, punct[0], punct[1], short[0], short[1]]),
           "sharing": np.array(sharing), "T": np.array(T), "decoding_type": np.array(decoding_type),
           "q_bit": np.array(q_bit), "clip": np.array(clip, dtype=np.float32), "xa": X, "app": res["app"].astype(np.float32),
           "sigma": np.asarray(rd.snr_sigma), "codeword": Y.astype(np.uint8), "uncor_flag": np.asarray(uncor),
           "error_num": np.asarray(error_num), "metrics": np.array([ber_last, fer_last, fer])}
    for i in range(3):
        if sharing[i] > 0:
            out[f"w{i}"] = np.asarray(weights[i], dtype=np.float32)[:T]
    np.savez_compressed(os.path.join(OUT, f"decode_{name}.npz"), **out)
    print(f"decode_{name}.npz  B={B} T={T}  ones per codeword {Y.sum(axis=1)[:4]}  FER {fer:.3f} FER_last {fer_last:.3f} BER_last {ber_last:.3e}")


def make_cw_cases():
    _, p, z, _, _ = graph_meta("mackay")
    M, N = p.shape
    E = int((p != -1).sum())
    make_cw_case("mackay_qms_300_t20_cw", "mackay", [3, 0, 0], const_weights([3, 0, 0], 20, M, N, E), 20, 2, 5, 16, [1.0, 2.0, 3.0, 4.0])
    sh, w = shipped("wimax_base20")
    make_cw_case("wimax_qms_333_t20_cw", "wimax", sh, w, 20, 2, 5, 8, [2.0, 2.5, 3.0, 3.5])
    make_cw_case("wimax_float_333_t10_cw", "wimax", sh, w, 10, 1, 5, 6, [2.0, 3.0])


def make_mc():
    """compute_results exactly as main_Base.py:177 calls it at epoch 0 with sampling_type=2."""
    sh, w = shipped("wimax_base20")
    stem, proto, z, punct, short = graph_meta("wimax")
    snr = [2.5, 3.0]
    rd = ref_runner.ReferenceDecoder(proto.astype(int), z, sh, w, 20, 2, 5, 20.0, punct, short, snr)
    with tempfile.TemporaryDirectory() as tmp:
        # sampling_type 2 insists on one SNR point (Main_Functions.py:502-505): run the points one by one
        results, texts = [], []
        for k, s in enumerate(snr):
            rd1 = ref_runner.ReferenceDecoder(proto.astype(int), z, sh, w, 20, 2, 5, 20.0, punct, short, [s])
            res, _ = rd1.compute_results(200, 2044, 1076, 20, sampling_type=2, cwd=tmp)
            results.append(res[:, 0])
            path = os.path.join(tmp, "Uncor.txt")
            texts.append(open(path).read() if os.path.exists(path) else "")
            if os.path.exists(path):
                os.remove(path)
    np.savez_compressed(os.path.join(OUT, "mc_wimax.npz"), snr=np.array(snr), results=np.stack(results, axis=1),
                        sigma=np.asarray(rd.snr_sigma), uncor_text_0=np.array(texts[0]),
                        uncor_text_1=np.array(texts[1]), sample_num=np.array(200), batch=np.array(20),
                        seeds=np.array([2044, 1076]))
    print("mc_wimax.npz", np.stack(results, axis=1))


def make_grad_case(name, gkey, sharing, weights, T, t_lo, loss_type, etha, decoding_type, B, snr_db, seed=7,
                   q_bit=5, clip=20.0, target_node=None):
    """One training batch through the reference's own loss (Main_Functions.py:337-378), differentiated by
    torch.autograd (oracle/ref_grad.py): the golden for the CUDA backward ("next" row N1)."""
    from oracle import ref_grad
    if ONLY and not any(o in name for o in ONLY):
        return
    stem, proto, z, punct, short = graph_meta(gkey)
    rd = ref_runner.ReferenceDecoder(proto.astype(int), z, sharing, weights, T, decoding_type, q_bit, clip, punct,
                                     short, snr_db)
    X = ref_llrs(rd.pf, rd.snr_sigma, B, rd.N, z, decoding_type, punct, short, q_bit, clip, seed)
    r = ref_grad.loss_and_grads(proto.astype(int), z, sharing, weights, X, T, iter_start=t_lo, loss_type=loss_type,
                                etha=etha, decoding_type=decoding_type, q_bit=q_bit, clip_llr=clip, punct=punct,
                                short=short, target_node=target_node)
    out = {"proto": proto.astype(np.int16), "meta": np.array([z, punct[0], punct[1], short[0], short[1]]),
           "sharing": np.array(sharing), "T": np.array(T), "t_lo": np.array(t_lo), "loss_type": np.array(loss_type),
           "etha": np.array(etha, dtype=np.float64), "decoding_type": np.array(decoding_type), "q_bit": np.array(q_bit),
           "clip": np.array(clip, dtype=np.float32), "xa": X, "loss": np.array(r["loss"], dtype=np.float64),
           "target_node": np.array(-1 if target_node is None else target_node)}
    for i in range(3):
        if sharing[i] > 0:
            w = np.asarray(weights[i], dtype=np.float32)[:T]
            out[f"w{i}"] = w
            gr = np.zeros_like(w.reshape(T, -1))
            for t in range(t_lo, T):
                gr[t] = r["grads"][(i, t)]
            out[f"g{i}"] = gr
    np.savez_compressed(os.path.join(OUT, f"grad_{name}.npz"), **out)
    print(f"grad_{name}.npz  loss {r['loss']:.6g}  max|g| {max(np.abs(out[k]).max() for k in out if k[0] == 'g' and k[1:].isdigit()):.4g}")


def make_grad_cases():
    sh, w = shipped("wimax_base20")
    make_grad_case("wimax_qms_333_fer_t6", "wimax", sh, w, 6, 0, 2, 0.0, 2, 6, [2.0, 2.5])
    make_grad_case("wimax_qms_333_bce_eta_t6", "wimax", sh, w, 6, 2, 0, 0.5, 2, 6, [2.0, 2.5])
    make_grad_case("wimax_float_333_sber_t5", "wimax", sh, w, 5, 0, 1, 1.0, 1, 4, [2.0])
    sh5, w5 = shipped("5g_r073_z32_boost50")
    make_grad_case("5g_r073_z32_qms_222_fer_t8", "5g_r073_z32", sh5, w5, 8, 3, 2, 0.7, 2, 4, [1.0, 2.0])
    stem, proto, z, punct, short = graph_meta("5g_r073_z32")
    M, N = proto.shape
    make_grad_case("5g_r073_z32_qms_222_bce_sys_t6", "5g_r073_z32", sh5, w5, 6, 0, 0, 1.0, 2, 4, [1.0, 2.0],
                   target_node=N - M)
    stem, proto, z, punct, short = graph_meta("wimax")
    M, N = proto.shape
    E = int((proto != -1).sum())
    rng = np.random.RandomState(5)
    make_grad_case("wimax_qms_112_fer_t4", "wimax", [1, 1, 2], const_weights([1, 1, 2], 4, M, N, E, rng=rng), 4, 0, 2, 0.0,
                   2, 4, [2.0])
    make_grad_case("wimax_qms_303_bce_t4", "wimax", [3, 0, 3], {0: w[0], 2: w[2]}, 4, 1, 0, 0.0, 2, 4, [2.5])
    stem, proto, z, punct, short = graph_meta("mackay")
    M, N = proto.shape
    E = int((proto != -1).sum())
    make_grad_case("mackay_float_300_bce_t5", "mackay", [3, 0, 0], const_weights([3, 0, 0], 5, M, N, E), 5, 0, 0, 1.0, 1,
                   16, [2.0, 3.0])


if __name__ == "__main__":
    if not ref_runner.reference_available():
        raise SystemExit("needs /root/reference (development container only)")
    # `make_golden.py decode only=sys,fixed` re-mints just the decode cases whose name contains one of the keys
    ONLY[:] = [k for a in sys.argv[1:] if a.startswith("only=") for k in a[5:].split(",")]
    what = [a for a in sys.argv[1:] if not a.startswith("only=")] or ["codes", "decode", "cw", "mc", "grad"]
    if "codes" in what:
        make_codes()
    if "decode" in what:
        make_decode_cases()
    if "cw" in what:
        make_cw_cases()
    if "mc" in what:
        make_mc()
    if "grad" in what:
        make_grad_cases()
