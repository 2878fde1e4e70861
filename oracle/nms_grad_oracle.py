"""Oracle B for the training step ("next" row N1): loss and weight gradients of one batch, restated by hand.
TEST INFRASTRUCTURE ONLY.

Forward = oracle/nms_oracle.decode (build_neural_network, Main_Functions.py:161-335) with every intermediate
kept; loss = Main_Functions.py:337-357; backward = the gradient TensorFlow's autodiff assigns to that graph:
  * Cal_MSA_Q_TF (:475-494) is a straight-through estimator: forward quantise, gradient of clip(x, +-qmax)
    (passes where |x| <= qmax, inclusive);  clip_by_value likewise;
  * reduce_min (:249, :350) splits the gradient evenly among tied minima;  abs -> sign(x);
  * tf.sign, tf.to_float(x > 0) and every comparison carry no gradient, so the sign product (:251-253), the
    UCN indicator (:180-206), the ReLU gate (:308) and the zero rules (:228, :250) are constants;
  * sign_through (:457-460): forward sign(x), gradient of inv_exp(x) = 2 / (1 + exp(-x)) - 1.
Pinned against torch.autograd of the reference's own code (oracle/ref_grad.py) in tests/test_oracle_grad.py and
through the committed goldens tests/golden/grad_*.npz.
"""
from __future__ import annotations

import numpy as np

from .nms_oracle import F32, OracleGraph, quantize, weight_row

QMAX = {5: 7.5, 6: 15.5, -5: 15.0, 4: 7.0, 3: 6.0}


def _reduce_weight_grad(ge, code, kind, g):
    """per-edge (CN/UCN) or per-column (VN) gradient [E] / [N] -> the shape of one weight row."""
    if code == 3:
        return np.array([ge.sum()], dtype=np.float64)
    if code == 2:
        if kind == "vn":
            return ge.astype(np.float64)
        out = np.zeros(g.M, dtype=np.float64)
        np.add.at(out, g.row, ge)
        return out
    return ge.astype(np.float64)      # code 1: per edge


def loss_and_grads(g: OracleGraph, xa, sharing, weights, T, t_lo=0, loss_type=2, etha=0.0, decoding_type=2, q_bit=5,
                   clip_llr=20.0, target_node=None):
    """xa f32 [B, N, z]; weights {i: [T, width]}; trainable iterations / loss iterations = [t_lo, T)
    (t_lo = max(training_iter_start - fixed_init, fixed_iter), Main_Functions.py:342, 368).
    Returns dict(loss, grads {(i, t): f32 row}, app [T, B, N*z])."""
    xa = np.ascontiguousarray(xa, dtype=F32)
    B = xa.shape[0]
    z, E, M, N = g.z, g.E, g.M, g.N
    clip = F32(clip_llr)
    qms = decoding_type == 2
    qmax = F32(QMAX[q_bit]) if qms else clip
    target = N if target_node is None else int(target_node)
    lane = np.arange(z)
    sat = (lambda x: quantize(x, q_bit)) if qms else (lambda x: np.clip(x, -clip, clip))
    xq = quantize(xa, q_bit) if qms else xa
    c2v = np.zeros((B, E, z), dtype=F32)
    rec, app_prev = [], None
    # ------------------------------------------------------------------ forward, keeping the intermediates
    for t in range(T):
        r = {}
        if sharing[2] in (2, 3):
            wv = weight_row(weights[2][t], sharing[2], "vn", g)
            xin_pre = xa * wv[None, :, None]
        else:
            xin_pre = xa
        xin = quantize(xin_pre, q_bit) if qms else xin_pre
        r["xin_mask"] = (np.abs(xin_pre) <= qmax) if qms else np.ones_like(xin_pre, dtype=bool)
        if sharing[1] > 0:
            src = xin if t == 0 else app_prev.reshape(B, N, z)
            sgn = np.where(-src > 0, F32(1), F32(-1))
            par = np.ones((B, M, z), dtype=F32)
            for e in range(E):
                par[:, g.row[e], :] *= sgn[:, g.col[e], (lane + g.shift[e]) % z]
            ucn_check = par < 0
        else:
            ucn_check = np.zeros((B, M, z), dtype=bool)
        v2c = np.empty((B, E, z), dtype=F32)
        for j in range(N):
            es = g.col_edges[j]
            for e in es:
                acc = np.zeros((B, z), dtype=F32)
                for e2 in es:
                    if e2 != e:
                        acc = acc + c2v[:, e2, :]
                v2c[:, e, :] = xin[:, j, :] + acc
        pre = np.empty_like(v2c)                                   # CN lane frame, before saturation
        for e in range(E):
            pre[:, e, :] = v2c[:, e, (lane + g.shift[e]) % z]
        r["v2c_mask"] = np.abs(pre) <= qmax
        v2cc = sat(pre)
        v2cc = v2cc + F32(0.0001) * (F32(1) - (np.abs(v2cc) > 0).astype(F32))
        r["v2cc"] = v2cc
        out = np.empty((B, E, z), dtype=F32)
        sig = np.empty((B, E, z), dtype=F32)
        for i in range(M):
            es = g.row_edges[i]
            vals = v2cc[:, es, :]
            for p, e in enumerate(es):
                others = np.delete(vals, p, axis=1)
                if others.shape[1] == 0:
                    m = np.full((B, z), F32(10000))
                    prod = np.ones((B, z), dtype=F32)
                else:
                    m = np.min(np.abs(others), axis=1)
                    prod = np.prod(np.where(others > 0, F32(-1), F32(1)), axis=1, dtype=F32)
                m = m + F32(-0.0001) * (F32(1) - (np.abs(m) > F32(0.0001)).astype(F32))
                sig[:, e, :] = np.sign(-prod)
                out[:, e, :] = m * sig[:, e, :]
        r["sig"] = sig
        x0 = np.empty_like(out)
        ucn_edge = np.empty((B, E, z), dtype=bool)
        for e in range(E):
            idx = (lane - g.shift[e]) % z
            x0[:, e, :] = out[:, e, idx]
            ucn_edge[:, e, :] = ucn_check[:, g.row[e], :][:, idx]
        mag = np.abs(x0)
        u = ucn_edge.astype(F32)
        if sharing[0] == 0:
            x1 = mag
            weff = np.ones((1, E, 1), dtype=F32)
        else:
            w0 = weight_row(weights[0][t], sharing[0], "cn", g)[None, :, None]
            if sharing[1] == sharing[0]:
                w1 = weight_row(weights[1][t], sharing[1], "cn", g)[None, :, None]
                x1 = (mag * w0) * (F32(1) - u) + (mag * w1) * u
                weff = w0 * (F32(1) - u) + w1 * u
            else:
                x1 = mag * w0
                weff = w0 * np.ones_like(u)
        gate = (x1 > 0).astype(F32)
        x2 = x1 * gate
        r.update(x0=x0, mag=mag, u=u, weff=weff, gate=gate, out_mask=np.abs(x2) <= qmax)
        c2v = (sat(x2) * np.sign(x0)).astype(F32)
        s = np.zeros((B, N, z), dtype=F32)
        for e in range(E):
            s[:, g.col[e], :] = s[:, g.col[e], :] + c2v[:, e, :]
        raw = xq + s
        r["app_mask"] = np.abs(raw) <= clip
        app = np.clip(raw, -clip, clip)
        r["app"] = app
        app_prev = app.reshape(B, N * z)
        rec.append(r)

    # ------------------------------------------------------------------ loss (Main_Functions.py:339-357)
    coefs = {t: float(etha) ** (T - 1 - t) for t in range(t_lo, T)}            # pow(etha, k); 0 ** 0 = 1
    norm = sum(coefs.values())
    tz = target * z
    loss = 0.0
    g_app = {}
    for t in range(t_lo, T):
        x = rec[t]["app"].reshape(B, N * z)[:, :tz].astype(np.float64)
        ga = np.zeros((B, N * z), dtype=np.float64)
        if loss_type == 0:      # sigmoid cross entropy with labels 0 = softplus(x)
            ell = np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))
            loss += coefs[t] / norm * ell.mean()
            ga[:, :tz] = coefs[t] / norm / (B * tz) / (1.0 + np.exp(-x))
        elif loss_type == 1:    # soft BER
            sg = 1.0 / (1.0 + np.exp(-x))
            loss += coefs[t] / norm * sg.mean()
            ga[:, :tz] = coefs[t] / norm / (B * tz) * sg * (1.0 - sg)
        else:                   # FER: 1/2 (1 - sign_through(min(-x)))
            m = (-x).min(axis=1)
            loss += coefs[t] / norm * (0.5 * (1.0 - np.sign(m))).mean()
            ties = (-x) == m[:, None]
            dinv = 2.0 * np.exp(-m) / (1.0 + np.exp(-m)) ** 2
            # d loss / d m = -1/2 inv_exp'(m);  d m / d x_j = -1 / n_ties on the tied maxima of x
            ga[:, :tz] = (coefs[t] / norm / B) * (0.5 * dinv / ties.sum(axis=1))[:, None] * ties
        g_app[t] = ga.reshape(B, N, z)

    # ------------------------------------------------------------------ backward
    grads = {}
    g_next = np.zeros((B, E, z), dtype=np.float64)        # d loss / d c2v_{t+1} from the later iterations
    for t in range(T - 1, t_lo - 1, -1):
        r = rec[t]
        ga = g_app[t] * r["app_mask"]
        g_c2v = g_next.copy()
        for e in range(E):
            g_c2v[:, e, :] += ga[:, g.col[e], :]
        sx0 = np.sign(r["x0"]).astype(np.float64)
        g_x1 = g_c2v * sx0 * r["out_mask"] * r["gate"]
        if sharing[0] != 0:
            if sharing[1] == sharing[0]:
                grads[(0, t)] = _reduce_weight_grad((g_x1 * r["mag"] * (1 - r["u"])).sum(axis=(0, 2)), sharing[0], "cn", g)
                grads[(1, t)] = _reduce_weight_grad((g_x1 * r["mag"] * r["u"]).sum(axis=(0, 2)), sharing[1], "cn", g)
            else:
                grads[(0, t)] = _reduce_weight_grad((g_x1 * r["mag"]).sum(axis=(0, 2)), sharing[0], "cn", g)
                if sharing[1] > 0:
                    grads[(1, t)] = np.zeros(1)
        g_x0 = g_x1 * r["weff"] * sx0
        g_out = np.empty_like(g_x0)
        for e in range(E):
            g_out[:, e, :] = g_x0[:, e, (lane + g.shift[e]) % z]
        g_m = g_out * r["sig"]
        v2cc = r["v2cc"]
        g_abs = np.zeros((B, E, z), dtype=np.float64)
        for i in range(M):
            es = list(g.row_edges[i])
            av = np.abs(v2cc[:, es, :])
            for p, e in enumerate(es):
                if len(es) == 1:
                    continue
                oth = [q for q in range(len(es)) if q != p]
                o = av[:, oth, :]
                m = o.min(axis=1, keepdims=True)
                ind = (o == m)
                share = g_m[:, e, :][:, None, :] * ind / ind.sum(axis=1, keepdims=True)
                for k, q in enumerate(oth):
                    g_abs[:, es[q], :] += share[:, k, :]
        g_pre = g_abs * np.sign(v2cc) * r["v2c_mask"]
        g_v2c = np.empty_like(g_pre)
        for e in range(E):
            g_v2c[:, e, :] = g_pre[:, e, (lane - g.shift[e]) % z]
        g_in = np.zeros((B, E, z), dtype=np.float64)
        g_xin = np.zeros((B, N, z), dtype=np.float64)
        for j in range(N):
            es = g.col_edges[j]
            G = sum(g_v2c[:, e, :] for e in es)
            g_xin[:, j, :] = G
            for e in es:
                g_in[:, e, :] = G - g_v2c[:, e, :]
        if sharing[2] in (2, 3):
            grads[(2, t)] = _reduce_weight_grad((g_xin * r["xin_mask"] * xa).sum(axis=(0, 2)), sharing[2], "vn", g)
        g_next = g_in
    grads = {k: np.asarray(v, dtype=np.float64) for k, v in grads.items() if not (k[0] == 1 and sharing[1] != sharing[0])}
    return {"loss": float(loss), "grads": grads, "app": np.stack([r["app"].reshape(B, N * z) for r in rec])}
