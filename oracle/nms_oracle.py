"""Oracle B -- sparse numpy restatement of the reference decode path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker.  The product
(`ldpc_error_floor_b200/`) never imports anything under `oracle/`.

Parity status: PINNED against the reference's own code (Oracle A =
`oracle/ref_runner.py`, the unmodified `/root/reference/Main_Functions.py` under the
numpy TF shim) by `tests/test_oracle_vs_reference.py` in the dev container and, where
`/root/reference` is absent (GPU box), by the committed golden vectors in
`tests/golden/*.npz` that `tests/golden/make_golden.py` minted from Oracle A.
The reference ships no golden vectors / tests of its own.

Every function cites the reference lines it restates.  Arithmetic is float32
throughout, like the TF graph.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- graph
class OracleGraph:
    """Edge lists of a proto-matrix.  Restates Main_Functions.py:8-38 (init_parameter)
    and the index conventions of init_connecting_matrix (:46-150):
    E(C) = row-major edge order (:69-75), lifted check (i,a) <-> variable (j,(a+s)%z) (:64-66)."""

    def __init__(self, proto, z, punct=(0, 0), short=(0, 0)):
        proto = np.asarray(proto, dtype=np.int64)
        self.proto = proto
        self.M, self.N = proto.shape
        self.z = int(z)
        rows, cols, shifts = [], [], []
        for i in range(self.M):               # E(C): for i: for j  (Main_Functions.py:69-70)
            for j in range(self.N):
                if proto[i, j] != -1:
                    rows.append(i)
                    cols.append(j)
                    shifts.append(int(proto[i, j]) % self.z)   # :72
        self.row = np.array(rows)
        self.col = np.array(cols)
        self.shift = np.array(shifts)
        self.E = len(rows)
        self.cn_deg = np.bincount(self.row, minlength=self.M)
        self.vn_deg = np.bincount(self.col, minlength=self.N)
        self.row_edges = [np.nonzero(self.row == i)[0] for i in range(self.M)]
        self.col_edges = [np.nonzero(self.col == j)[0] for j in range(self.N)]  # ascending E(C)
        self.punct = tuple(int(v) for v in punct)
        self.short = tuple(int(v) for v in short)
        # Main_Functions.py:24-29 -- note the "+1" even when start=end=0
        punct_num = self.punct[1] - self.punct[0] + 1
        short_num = self.short[1] - self.short[0] + 1
        self.n_ref = self.N * self.z - punct_num - short_num
        self.k_ref = (self.N - self.M) * self.z - short_num
        self.rate_ref = 1.0 * self.k_ref / self.n_ref

    def sigma(self, snr_db):
        """Main_Functions.py:35-36."""
        snr_lin = 10.0 ** (np.asarray(snr_db, dtype=np.float64) / 10.0)
        return np.sqrt(1.0 / (2.0 * snr_lin * self.rate_ref))

    def syndrome(self, bits):
        """bits [B, N*z] (index j*z+c) -> [B, M, z] parity of each lifted check."""
        B = bits.shape[0]
        b = np.asarray(bits).reshape(B, self.N, self.z).astype(np.int64)
        s = np.zeros((B, self.M, self.z), dtype=np.int64)
        for e in range(self.E):
            s[:, self.row[e], :] ^= np.roll(b[:, self.col[e], :], -self.shift[e], axis=1)
        return s


# ----------------------------------------------------------------------- quantiser
def quantize(x, q_bit):
    """Forward value of Cal_MSA_Q_TF (Main_Functions.py:475-494) == Cal_MSA_Q
    (Print_Functions.py:12-25).  np.rint == tf.round == half-to-even."""
    x = np.asarray(x, dtype=F32)
    if q_bit == 6:
        return np.clip(np.rint(x), F32(-15.5), F32(15.5))
    if q_bit == 5:
        return np.clip(np.rint(x * F32(2)) / F32(2), F32(-7.5), F32(7.5))
    if q_bit == -5:
        return np.clip(np.rint(x), F32(-15), F32(15))
    if q_bit == 4:
        return np.clip(np.rint(x), F32(-7), F32(7))
    if q_bit == 3:
        return np.clip(np.rint(x / F32(2)) * F32(2), F32(-6), F32(6))
    raise ValueError(f"q_bit {q_bit} has no branch in the reference")


def weight_row(w_t, code, kind, g: OracleGraph):
    """Expand one iteration's weight row to a per-edge (CN/UCN, E(C) order) or per-column (VN)
    vector.  Main_Functions.py:269-304 (CN/UCN) and :168-169 (VN)."""
    w_t = np.asarray(w_t, dtype=F32).reshape(-1)
    if kind == "vn":
        if code == 3:
            return np.full(g.N, w_t[0], dtype=F32)
        if code == 2:
            assert w_t.size == g.N
            return w_t
        raise ValueError("VN sharing must be 0, 2 or 3 (Main_Functions.py:515-517)")
    if code == 3:
        return np.full(g.E, w_t[0], dtype=F32)
    if code == 2:
        assert w_t.size == g.M
        return w_t[g.row]
    if code == 1:
        assert w_t.size == g.E
        return w_t
    raise ValueError(f"sharing code {code} not supported on the decode path")


# -------------------------------------------------------------------------- decode
def decode(g: OracleGraph, xa, sharing, weights, T, decoding_type=2, q_bit=5, clip_llr=20.0,
           return_c2v=False):
    """Restates build_neural_network (Main_Functions.py:161-335) for iterations 0..T-1.

    xa: f32 [B, N, z] (log p1/p0).  weights: {0: cn[T,w], 1: ucn[T,w], 2: vn[T,w]} for the
    non-zero sharing codes.  Returns dict(app [T,B,N*z] f32, synd [T,B] bool (any check
    unsatisfied by APP_t >= 0), ucn [T,B,M,z] bool (mask consumed at iteration t))."""
    assert decoding_type in (0, 1, 2)      # 0 = sum-product (:238-245), 1 = min-sum, 2 = quantised min-sum
    xa = np.ascontiguousarray(xa, dtype=F32)
    B = xa.shape[0]
    z, E, M, N = g.z, g.E, g.M, g.N
    clip = F32(clip_llr)
    qms = decoding_type == 2
    c2v = np.zeros((B, E, z), dtype=F32)          # VN-lane frame, E(C) order (LLRa0, main_Base.py:126)
    xq = quantize(xa, q_bit) if qms else xa       # :321-322
    apps, synds, ucns, c2vs = [], [], [], []
    app_prev = None
    lane = np.arange(z)
    for t in range(T):
        # ---- D2  VN weight + input quantise (:168-177)
        if sharing[2] in (2, 3):
            wv = weight_row(weights[2][t], sharing[2], "vn", g)
            xin = xa * wv[None, :, None]
        else:
            xin = xa
        if qms:
            xin = quantize(xin, q_bit)
        # ---- D3  UCN indicator from the previous hard decision (:180-209)
        if sharing[1] > 0:
            src = xin if t == 0 else app_prev.reshape(B, N, z)
            sgn = np.where(-src > 0, F32(1), F32(-1))            # :187-188  (bit = src >= 0)
            par = np.ones((B, M, z), dtype=F32)
            for e in range(E):
                par[:, g.row[e], :] *= sgn[:, g.col[e], (lane + g.shift[e]) % z]
            ucn_check = par < 0                                   # :200  [B,M,z] in CN lane frame
        else:
            ucn_check = np.zeros((B, M, z), dtype=bool)
        ucns.append(ucn_check)
        # ---- D4  VN update: direct extrinsic sum, then + channel (:213-215)
        v2c = np.empty((B, E, z), dtype=F32)
        for j in range(N):
            es = g.col_edges[j]
            for e in es:
                acc = np.zeros((B, z), dtype=F32)
                for e2 in es:                                     # ascending E(C) index, like the GEMM row
                    if e2 != e:
                        acc = acc + c2v[:, e2, :]
                v2c[:, e, :] = xin[:, j, :] + acc                 # x2 = x0 + x1 (:215)
        # lift into the CN lane frame (:217-221): v2c'[a] = v2c[(a+s)%z]
        v2cc = np.empty_like(v2c)
        for e in range(E):
            v2cc[:, e, :] = v2c[:, e, (lane + g.shift[e]) % z]
        # ---- D5  saturation and the zero rule (:223-230)
        if qms:
            v2cc = quantize(v2cc, q_bit)
        else:
            v2cc = np.clip(v2cc, -clip, clip)
        if decoding_type in (1, 2):                               # :229-230 (not for sum-product)
            v2cc = v2cc + F32(0.0001) * (F32(1) - (np.abs(v2cc) > 0).astype(F32))
        # ---- D6  CN update (:231-254)
        out = np.empty((B, E, z), dtype=F32)
        for i in range(M):
            es = g.row_edges[i]
            vals = v2cc[:, es, :]                                 # [B, dc, z]
            if decoding_type == 0:
                # sum-product (:238-245): tanh(-x/2), a zero factor (a masked entry of the dense tile, but also a true
                # zero message) counts as 1, product over the other edges in E(C) order, clip at +-(1 - 1e-7) as float32
                # (= 1 - 2^-23), -2 atanh
                th = np.tanh(F32(-0.5) * vals).astype(F32)
                th = th + (F32(1) - (np.abs(th) > 0).astype(F32))
                lim = F32(1) - F32(1e-7)
                for p, e in enumerate(es):
                    prod = np.ones((B, z), dtype=F32)
                    for p2 in range(len(es)):
                        if p2 != p:
                            prod = prod * th[:, p2, :]
                    out[:, e, :] = F32(-2) * np.arctanh(np.clip(prod, -lim, lim)).astype(F32)
                continue
            for p, e in enumerate(es):
                others = np.delete(vals, p, axis=1)
                if others.shape[1] == 0:
                    m = np.full((B, z), F32(10000))               # all-masked row -> 10000 (:248)
                    prod = np.ones((B, z), dtype=F32)
                else:
                    m = np.min(np.abs(others), axis=1)            # :248-249
                    prod = np.prod(np.where(others > 0, F32(-1), F32(1)), axis=1, dtype=F32)  # :251-252
                m = m + F32(-0.0001) * (F32(1) - (np.abs(m) > F32(0.0001)).astype(F32))        # :250
                out[:, e, :] = m * np.sign(-prod)                 # :253-254
        # ---- D7  lift back (:259-263): x0[c] = out[(c-s)%z]; weight, ReLU, saturate, sign
        x0 = np.empty_like(out)
        ucn_edge = np.empty((B, E, z), dtype=bool)
        for e in range(E):
            idx = (lane - g.shift[e]) % z
            x0[:, e, :] = out[:, e, idx]
            ucn_edge[:, e, :] = ucn_check[:, g.row[e], :][:, idx]
        mag = np.abs(x0)
        if sharing[0] == 0:
            x1 = mag                                              # :267-268
        else:
            w0 = weight_row(weights[0][t], sharing[0], "cn", g)[None, :, None]
            if sharing[1] == sharing[0]:
                w1 = weight_row(weights[1][t], sharing[1], "cn", g)[None, :, None]
                u = ucn_edge.astype(F32)
                x1 = (mag * w0) * (F32(1) - u) + (mag * w1) * u   # :275 / :285 / :295
            else:
                x1 = mag * w0
        x2 = x1 * (x1 > 0).astype(F32)                            # :308
        if qms:
            x2 = quantize(x2, q_bit)                              # :310-311
        else:
            x2 = np.clip(x2, -clip, clip)                         # :313
        c2v = (x2 * np.sign(x0)).astype(F32)                      # :316
        # ---- D8  APP (:317-327)
        s = np.zeros((B, N, z), dtype=F32)
        for e in range(E):                                        # ascending E(C), like the GEMM column
            s[:, g.col[e], :] = s[:, g.col[e], :] + c2v[:, e, :]
        app = np.clip(xq + s, -clip, clip).reshape(B, N * z)
        apps.append(app)
        app_prev = app
        synds.append(g.syndrome(app >= 0).reshape(B, -1).any(axis=1))
        if return_c2v:
            c2vs.append(c2v.copy())
    res = {"app": np.stack(apps), "synd": np.stack(synds), "ucn": np.stack(ucns)}
    if return_c2v:
        res["c2v"] = np.stack(c2vs)     # [T, B, E, z]
    return res


# ----------------------------------------------------------------- Monte-Carlo side
def create_mix_epoch(sigmas, word_random, noise_random, batch_size, N, z, decoding_type,
                     punct, short, q_bit, clip_llr):
    """Restates Print_Functions.create_mix_epoch (:29-72) for the all-zero codeword, drawing
    from the two numpy RandomState streams in the same order as the reference."""
    X = []
    cur = 0
    while cur < batch_size:
        for sf in sigmas:
            Y_i = 0 * word_random.randint(0, 2, size=(1, N * z))          # :39
            X_p = noise_random.normal(0.0, 1.0, Y_i.shape) * sf + (-1) ** (1 - Y_i)  # :45
            llr = 2 * X_p / (sf ** 2)                                     # :46
            if decoding_type == 2:
                llr = quantize_f64(llr, q_bit)                            # :49-50 (numpy, float64)
            if punct[0] > 0:
                llr[0, punct[0] - 1:punct[1]] = 0.001 if decoding_type == 0 else 0   # :53-57
            if short[0] > 0:
                llr[0, short[0] - 1:short[1]] = -clip_llr                 # :59-60
            X.append(llr.astype(np.float32)[0])                           # np.vstack onto f32 X (:31,:62)
            cur += 1
            if cur == batch_size:
                break
    X = np.stack(X).reshape(batch_size, N, z)
    Y = np.zeros((batch_size, N * z), dtype=np.int64)
    return X, Y


def quantize_f64(x, q_bit):
    """Print_Functions.Cal_MSA_Q (:12-25) in the float64 it is called with."""
    if q_bit == 6:
        return np.clip(np.round(x), -15.5, 15.5)
    if q_bit == 5:
        return np.clip(np.round(x * 2) / 2, -7.5, 7.5)
    if q_bit == -5:
        return np.clip(np.round(x), -15, 15)
    if q_bit == 4:
        return np.clip(np.round(x), -7, 7)
    if q_bit == 3:
        return np.clip(np.round(x / 2) * 2, -6, 6)
    raise ValueError(q_bit)


def calc_ber_fer(app_all, T, Y, batch_size):
    """Restates Print_Functions.calc_ber_fer (:100-118).  app_all: [T*B, L]."""
    L = app_all.shape[1]
    flags = []
    for t in range(T):
        blk = app_all[t * batch_size:(t + 1) * batch_size]
        flags.append(np.abs((blk >= 0) - Y[:, :L]).sum(axis=1) > 0)       # :106
    uncor = np.min(np.stack(flags).astype(np.float64), axis=0)            # :109  genie
    fer = uncor.sum() * 1.0 / batch_size                                  # :111
    last = app_all[(T - 1) * batch_size:T * batch_size]
    error_num = ((last >= 0) - Y[:, :L]).sum(axis=1)                      # :112
    ber_last = np.abs(error_num.sum()) / (Y.shape[0] * Y.shape[1])        # :113
    fer_last = (np.abs((last >= 0) - Y[:, :L]).sum(axis=1) > 0).sum() * 1.0 / Y.shape[0]   # :115-116
    return ber_last, fer_last, fer, uncor, error_num


def monte_carlo(g: OracleGraph, sample_num, sigmas, word_random, noise_random, batch_size,
                sharing, weights, T, decoding_type, q_bit, clip_llr, collect=False):
    """Restates Print_Functions.compute_results (:130-165) for sampling_type 0/2.
    Returns (Results f32[4,nSNR], list of harvested uncorrected LLR rows [N*z])."""
    res = np.zeros((4, len(sigmas)), dtype=np.float32)
    batch_num = math.floor(sample_num / batch_size)
    harvested = []
    for _ in range(batch_num):
        for si, sg in enumerate(sigmas):
            X, Y = create_mix_epoch([sg], word_random, noise_random, batch_size, g.N, g.z,
                                    decoding_type, g.punct, g.short, q_bit, clip_llr)
            out = decode(g, X, sharing, weights, T, decoding_type, q_bit, clip_llr)
            app_all = out["app"].reshape(T * batch_size, g.N * g.z)
            ber_last, fer_last, fer, uncor, _ = calc_ber_fer(app_all, T, Y, batch_size)
            if collect and np.sum(uncor == 1) > 0:
                harvested.append(X[uncor == 1].reshape(-1, g.N * g.z))    # :124 (sign flip is a file detail)
            res[0, si] += ber_last / batch_num
            res[1, si] += fer_last / batch_num
            res[2, si] += fer / batch_num
    return res, harvested
