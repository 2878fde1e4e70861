// nms_f32.cuh -- float32 arithmetic (one frame per 32-bit word) shared by the degree-bucketed generic kernels
// (nms_f32.cu) and the graph-specialised ones (nms_f32_spec.cuh): decoding_type 1 (min-sum, clip +-clip_LLR) and the
// quantised modes the packed kernels do not take (q_bit 6, per-edge weights).
//
// Arithmetic follows the TF graph: V->C = xin + (sum of the OTHER C->V of the variable) formed directly, never as
// total - self (Main_Functions.py:213-215), saturation and the 1e-4 zero rule (:223-230), minimum of the other
// edges with the "<= 1e-4 -> subtract 1e-4" rule (:248-250), sign product (:251-254), |.|*w -> ReLU -> saturate ->
// sign (:267-316), APP = clip(xq + sum) (:317-325).  Choices the reference leaves open or that cost nothing:
//  * summation order (TF's matmul order is unspecified): ascending E(C) order, as in oracle/nms_oracle.c, with the
//    common prefixes of the extrinsic sums formed once (f32_extrinsic) -- bit-identical to the restatement;
//  * the syndrome of the previous hard decision, which selects the unsatisfied-check weight (:180-206) and drives the
//    termination flags, is formed from the ballot words of the hard decisions (one word per column and 32-lane chunk,
//    written by the VN phase): lane p of the warp fetches the 32 hard bits the warp's lanes see through edge p of the row
//    (a funnel shift of two ballot words), one warp-wide XOR reduction gives the row's syndrome word -- about 20
//    instructions per row instead of 8 per edge, and the float messages stay untouched (the packed kernels' trick of
//    parking the bit in the mantissa LSB would cost the float path one unit in the last place per message);
//  * per-edge work moved to the check row where it commutes with the minimum: a V->C value only reaches the output
//    through (min1, min2) and its sign, so on the float path the clip of :225-226 is applied to the two minima instead
//    of to every edge, and the zero rule (:230, "0 counts as +1e-4"; a V->C word is never -0.0 because xin never is)
//    costs one compare per row: only a row whose smallest magnitude is 0 patches its zeros and redoes its minima.
#pragma once
#include "nms_h2.cuh"   // lds32 / ldsf / sts32

namespace nms {

constexpr uint32_t SIGN1 = 0x80000000u;

// per-thread constants (byte units, shared-window addresses)
struct F32Ctx {
    uint32_t sb;      // address of nms_smem[0]
    uint32_t q4;      // q * 4
    uint32_t amask;   // all ones for active lanes, 0 for padding lanes (they never rotate)
    uint32_t Lthr4;   // L*4 for active lanes, 2^30 for padding lanes (they never wrap)
    uint32_t xa4;     // sb + (off_xa + q) * 4   (+ j*LP*4 per column)
    uint32_t xq4;     // sb + (off_xq + q) * 4
    uint32_t et4;     // sb + off_et * 4: per-edge syndrome table {column * C | s_e*Fp << 16}
    uint32_t chunk32; // first lane of this warp's chunk
};

__device__ __forceinline__ F32Ctx f32_ctx(const KParams &P, const Ctx &c) {
    F32Ctx h;
    h.sb = smem_base();
    h.q4 = (uint32_t)c.q * 4u;
    h.amask = c.act ? 0xffffffffu : 0u;
    h.Lthr4 = c.act ? (uint32_t)P.L * 4u : 0x40000000u;
    h.xa4 = h.sb + (uint32_t)P.off_xa * 4u + h.q4;
    h.xq4 = h.sb + (uint32_t)P.off_xq * 4u + h.q4;
    h.et4 = h.sb + (uint32_t)P.off_et * 4u;
    h.chunk32 = (uint32_t)c.chunk * 32u;
    return h;
}

// variable lane -> byte offset of the rotated check lane: (q + rot) mod L, padding lanes stay put
template <bool PAD>
__device__ __forceinline__ uint32_t f32_rot(const F32Ctx &h, uint32_t rot4, uint32_t L4) {
    if (PAD) {
        const uint32_t t1 = h.q4 + (rot4 & h.amask);
        return min(t1, t1 - h.Lthr4);
    } else {
        const uint32_t t1 = h.q4 + rot4;
        return min(t1, t1 - L4);
    }
}

// QM: 0 = float path (clip), 1 = quantised, 2 = decided at run time without a branch (generic kernels:
// sat_magic = 0 makes the rounding step the identity)
template <int QM>
__device__ __forceinline__ float f32_sat(const KParams &P, float x) {
    if constexpr (QM == 0) return fminf(fmaxf(x, -P.clip), P.clip);                              // :225-226, :312-313
    else if constexpr (QM == 1) return fminf(fmaxf(qround(x, P.qmagic), -P.qmax), P.qmax);        // :223-224, :310-311
    else return fminf(fmaxf(qround(x, P.sat_magic), -P.sat_bound), P.sat_bound);
}
template <int QM>
__device__ __forceinline__ float f32_sat_pos(const KParams &P, float x) {   // x >= 0
    if constexpr (QM == 0) return fminf(x, P.clip);
    else if constexpr (QM == 1) return fminf(qround(x, P.qmagic), P.qmax);
    else return fminf(qround(x, P.sat_magic), P.sat_bound);
}
template <int QM>
__device__ __forceinline__ bool f32_is_qms(const KParams &P) { return QM == 2 ? P.qms != 0 : QM == 1; }

// V->C word: xin + ext (quantised path: Q() of it, :223-224; float path: the clip is applied by the check row).
// Never -0.0: xin is never -0.0 (f32_pos_zero) and x + (-x) = +0.
template <int QM>
__device__ __forceinline__ uint32_t f32_v2c(const KParams &P, float xin, float ext) {
    float m = __fadd_rn(xin, ext);
    if constexpr (QM != 0) m = f32_sat<QM>(P, m);
    return __float_as_uint(m);
}
__device__ __forceinline__ float f32_pos_zero(float x) { return __fadd_rn(x, 0.0f); }   // -0.0 -> +0.0, else unchanged

// weighted, saturated output magnitude for a minimum `mb` (bits) of the other edges;
// returns the bits of the value with the sign of the adjusted minimum folded in (:250, :254, :267-313)
// CLAMP: `mb` is a real V->C magnitude (not the 10000 of an all-masked row), to be saturated on the float path
template <int QM, bool CLAMP = true>
__device__ __forceinline__ uint32_t f32_row_mag(const KParams &P, uint32_t mb, float w) {
    float m = __uint_as_float(mb);
    const float Z = 0.0001f;
    if constexpr (QM == 0 && CLAMP) m = fminf(m, P.clip);             // :225-226
    if constexpr (QM == 2 && CLAMP) m = fminf(m, P.sat_bound);        // same; a no-op after Q() on the quantised path
    if constexpr (QM == 1) {
        // on the quantiser grid the smallest non-zero magnitude is a whole step, so a zero (which counts as 1e-4, :230) is
        // still the row's minimum and needs no patching: 0 -> 1e-4 -> 0 (:250)
        m = m > Z ? m : (m == 0.0f ? 0.0f : __fadd_rn(m, -Z));
    } else {
        m = m > Z ? m : __fadd_rn(m, -Z);                             // :250 (a zero V->C arrives here as 1e-4, :230)
    }
    const float x1 = __fmul_rn(fabsf(m), w);                          // :267-298
    const float x2 = f32_sat_pos<QM>(P, x1 > 0.0f ? x1 : 0.0f);       // :308-313
    return __float_as_uint(x2) ^ (__float_as_uint(m) & SIGN1);
}

// syndrome word of one check row for this warp's 32 check lanes: bit l = parity over the row's edges of the previous
// hard decision of the variable lane (chunk32 + l + s_e*Fp) mod L.  hb4: byte address of the ballot buffer hb[buf][.][.]
// (word [j*C + chunk'] holds the hard bits of column j, lanes 32*chunk' ..).  Lane p serves edge p (+32, +64 .. for
// rows longer than a warp); bits of padding lanes are zero in the ballots and ignored by the callers.
template <bool PAD>
__device__ __forceinline__ uint32_t f32_row_syndrome(const KParams &P, const F32Ctx &h, uint32_t hb4, int e0, int dc) {
    const int lane = threadIdx.x & 31;
    uint32_t acc = 0;
    for (int p = lane; p < dc; p += 32) {
        const uint32_t tb = lds32(h.et4 + (uint32_t)(e0 + p) * 4u);
        const uint32_t col4 = hb4 + (tb & 0xffffu) * 4u;
        uint32_t s = h.chunk32 + (tb >> 16);
        s = s >= (uint32_t)P.L ? s - (uint32_t)P.L : s;
        const uint32_t w = s >> 5, r = s & 31u;
        if constexpr (!PAD) {   // L is a multiple of 32: the wrap falls on a word boundary
            const uint32_t w1 = w + 1u == (uint32_t)P.C ? 0u : w + 1u;
            acc ^= __funnelshift_r(lds32(col4 + w * 4u), lds32(col4 + w1 * 4u), r);
        } else {
            const uint32_t w1 = min(w + 1u, (uint32_t)P.C - 1u);
            uint32_t f = __funnelshift_r(lds32(col4 + w * 4u), lds32(col4 + w1 * 4u), r);
            const uint32_t thr = (uint32_t)P.L - s;   // lanes l >= thr wrap to variable lane l - thr
            if (thr < 32u) f = (f & ((1u << thr) - 1u)) | (lds32(col4) << thr);
            acc ^= f;
        }
    }
    return (__reduce_xor_sync(0xffffffffu, acc) >> lane) & 1u;
}

// CN / UCN weight of a row when the weights are not per edge (sharing 0 / 2 / 3)
__device__ __forceinline__ void f32_row_weights(const KParams &P, int t, int i, float &w0, float &w1) {
    w0 = cn_weight(P, t, i, 0);
    w1 = P.sharing1 != 0 ? ucn_weight(P, t, i, 0) : w0;
}

// (min1, min2) of |raw[LO..HI)| as a tournament (see h2_min12): 35 operations for 15 edges instead of 45
template <int DC, int LO, int HI>
__device__ __forceinline__ void f32_min12t(const uint32_t (&raw)[DC], float &m1, float &m2) {
    if constexpr (HI - LO == 1) {
        m1 = fabsf(__uint_as_float(raw[LO]));
        m2 = 10000.0f;   // all-masked row -> 10000 (:248)
    } else if constexpr (HI - LO == 2) {
        const float a = fabsf(__uint_as_float(raw[LO])), b = fabsf(__uint_as_float(raw[LO + 1]));
        m1 = fminf(a, b);
        m2 = fmaxf(a, b);
    } else if constexpr (HI - LO == 3) {
        f32_min12t<DC, LO, LO + 2>(raw, m1, m2);
        const float a = fabsf(__uint_as_float(raw[LO + 2]));
        const float t = fmaxf(m1, a);
        m1 = fminf(m1, a);
        m2 = fminf(m2, t);
    } else {
        constexpr int MID = LO + (((HI - LO) / 2 + 1) & ~1);   // even-sized left half: its leaves are pairs
        float a1, a2, b1, b2;
        f32_min12t<DC, LO, MID>(raw, a1, a2);
        f32_min12t<DC, MID, HI>(raw, b1, b2);
        const float hi = fmaxf(a1, b1);
        m1 = fminf(a1, b1);
        m2 = fminf(fminf(hi, a2), b2);
    }
}
template <int DC>
__device__ __forceinline__ void f32_min12(const uint32_t (&raw)[DC], float &m1, float &m2) { f32_min12t<DC, 0, DC>(raw, m1, m2); }

__device__ __forceinline__ uint32_t opaque_xor(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// one check row held in registers, one weight per row.  a0: byte address of msg[e0][q]; stride4 = LP*4
// par: this lane's syndrome bit of the previous hard decision (f32_row_syndrome)
// ZERO: what to do about a zero V->C, which counts as +1e-4 (:230) --
//   ZERO_PATCH  patch the register array and redo the minima in line (generic kernels, run-time mode);
//   ZERO_BAIL   return false with nothing written (float path of the graph-specialised kernels): the caller then CALLS the
//               patched form, so the rare case costs the common one no register moves -- patching in line made ptxas keep
//               two copies of the array, 56 MOVs per 3 rows.  Rare means: float messages are zero when a channel value
//               is (words on a quantiser grid fed to the float decoder: iteration 0 only).
enum { ZERO_NONE = 0, ZERO_PATCH = 1, ZERO_BAIL = 2 };
template <int DC, int QM, int ZERO>
__device__ __forceinline__ bool cn_row_f32_body(const KParams &P, uint32_t a0, uint32_t stride4, float w0, float w1,
                                                uint32_t par) {
    uint32_t raw[DC];
#pragma unroll
    for (int p = 0; p < DC; ++p) raw[p] = lds32(a0 + p * stride4);
    uint32_t sx = 0;   // bit 31: parity of the negative inputs
#pragma unroll
    for (int p = 0; p < DC; ++p) sx ^= raw[p];
    float m1, m2;
    f32_min12<DC>(raw, m1, m2);
    if constexpr (ZERO == ZERO_BAIL) {
        if (m1 == 0.0f) return false;
    }
    if (ZERO == ZERO_PATCH && m1 == 0.0f) {
#pragma unroll
        for (int p = 0; p < DC; ++p) raw[p] = __uint_as_float(raw[p]) == 0.0f ? __float_as_uint(0.0001f) : raw[p];
        f32_min12<DC>(raw, m1, m2);
    }
    const float w = par ? w1 : w0;   // unsatisfied check -> UCN weight (:275,:285,:295)
    // C->V of edge p is negative iff the number of positive OTHER inputs is even (:251-254; an input is never 0, :230):
    // sign bit = sign(adjusted min) ^ (dc & 1) ^ parity(negative inputs) ^ own sign
    const uint32_t Pbit = (sx ^ ((DC & 1) ? SIGN1 : 0u)) & SIGN1;
    // The row's sign parity is folded into A and B ONCE, through an opaque XOR: left to itself the compiler pulls it back out
    // of the select (select(c, a ^ p, b ^ p) -> select(c, a, b) ^ p), and sel ^ p ^ (raw & SIGN) then has four inputs -- two
    // LOP3 per edge instead of one on the busiest pipe of this kernel.
    const uint32_t A = opaque_xor(f32_row_mag<QM, true>(P, __float_as_uint(m1), w), Pbit),
                   B = opaque_xor(f32_row_mag<QM, (DC >= 2)>(P, __float_as_uint(m2), w), Pbit);   // degree 1: min2 is the 10000 of :248
#pragma unroll
    for (int p = 0; p < DC; ++p) {
        const uint32_t v = fabsf(__uint_as_float(raw[p])) > m1 ? A : B;   // others' minimum: min1 unless this edge is it
        sts32(a0 + p * stride4, v ^ (raw[p] & SIGN1));
    }
    return true;
}
template <int DC, int QM>
static __device__ __noinline__ void cn_row_f32_zero(const KParams &P, uint32_t a0, uint32_t stride4, float w0, float w1,
                                                    uint32_t par) {
    cn_row_f32_body<DC, QM, ZERO_PATCH>(P, a0, stride4, w0, w1, par);
}
template <int DC, int QM>
__device__ __forceinline__ void cn_row_f32(const KParams &P, uint32_t a0, uint32_t stride4, float w0, float w1, uint32_t par) {
    if constexpr (QM == 0) {          // float path, specialised kernels
        if (!cn_row_f32_body<DC, 0, ZERO_BAIL>(P, a0, stride4, w0, w1, par)) cn_row_f32_zero<DC, 0>(P, a0, stride4, w0, w1, par);
    } else if constexpr (QM == 1) {   // quantised twin: a zero stays the row's minimum and needs no patching (f32_row_mag)
        cn_row_f32_body<DC, 1, ZERO_NONE>(P, a0, stride4, w0, w1, par);
    } else {
        cn_row_f32_body<DC, 2, ZERO_PATCH>(P, a0, stride4, w0, w1, par);
    }
}

// magnitude of a V->C word as the check sees it: a zero counts as +1e-4 (:230)
__device__ __forceinline__ float f32_eff(uint32_t r) {
    const float a = fabsf(__uint_as_float(r));
    return a == 0.0f ? 0.0001f : a;
}

// any degree, any weight sharing (per-edge weights included): two passes over shared memory
template <int QM>
static __device__ __noinline__ void cn_row_f32_generic(const KParams &P, uint32_t a0, uint32_t stride4, int dc, int t, int i,
                                                       int e0, uint32_t par) {
    uint32_t sx = 0;
    float m1 = 10000.0f, m2 = 10000.0f;
    for (int p = 0; p < dc; ++p) {
        const uint32_t r = lds32(a0 + p * stride4);
        sx ^= r;
        const float a = f32_eff(r);
        const float tmx = fmaxf(m1, a);
        m1 = fminf(m1, a);
        m2 = fminf(m2, tmx);
    }
    const bool ucn = P.sharing1 != 0 && par;
    const uint32_t Pbit = (sx ^ ((dc & 1) ? SIGN1 : 0u)) & SIGN1;
    if (P.sharing0 == 1) {
        for (int p = 0; p < dc; ++p) {
            const uint32_t r = lds32(a0 + p * stride4);
            const float w = ucn ? ucn_weight(P, t, i, e0 + p) : cn_weight(P, t, i, e0 + p);
            const bool ismin = !(f32_eff(r) > m1);
            const uint32_t v = (ismin && dc < 2) ? f32_row_mag<QM, false>(P, __float_as_uint(m2), w)
                                                 : f32_row_mag<QM, true>(P, __float_as_uint(ismin ? m2 : m1), w);
            sts32(a0 + p * stride4, v ^ Pbit ^ (r & SIGN1));
        }
    } else {
        const float w = ucn ? ucn_weight(P, t, i, e0) : cn_weight(P, t, i, e0);
        const uint32_t A = f32_row_mag<QM, true>(P, __float_as_uint(m1), w) ^ Pbit;
        const uint32_t B = (dc >= 2 ? f32_row_mag<QM, true>(P, __float_as_uint(m2), w) : f32_row_mag<QM, false>(P, __float_as_uint(m2), w)) ^ Pbit;
        for (int p = 0; p < dc; ++p) {
            const uint32_t r = lds32(a0 + p * stride4);
            sts32(a0 + p * stride4, (f32_eff(r) > m1 ? A : B) ^ (r & SIGN1));
        }
    }
}

// sum-product check update (decoding_type 0, Main_Functions.py:238-245): t_p = tanh(-clip(v_p) / 2), a zero factor counts
// as 1 (the reference cannot tell a zero message from a masked entry of its dense tile), product over the OTHER edges in
// E(C) order, clipped at +-(1 - 1e-7) -- in float32 that is 1 - 2^-23 -- and x0 = -2 atanh(product); then the common
// tail |x0| w -> ReLU -> clip -> sign (:267-316).  No zero rule on the inputs (:229 is min-sum only).  Any degree up to
// 64; the factors are staged in the message words themselves.  Not a hot path: O(dc^2) products, libm tanhf / atanhf.
static __device__ __noinline__ void cn_row_f32_sp(const KParams &P, uint32_t a0, uint32_t stride4, int dc, int t, int i, int e0,
                                                  uint32_t par) {
    const float lim = 1.0f - 1e-7f;
    for (int p = 0; p < dc; ++p) {
        const float v = fminf(fmaxf(__uint_as_float(lds32(a0 + p * stride4)), -P.clip), P.clip);   // :225-226
        const float th = tanhf(__fmul_rn(-0.5f, v));
        sts32(a0 + p * stride4, __float_as_uint(th == 0.0f ? 1.0f : th));
    }
    const bool ucn = P.sharing1 != 0 && par;
    float outv[64];
    for (int p = 0; p < dc; ++p) {
        float prod = 1.0f;
        for (int p2 = 0; p2 < dc; ++p2)
            if (p2 != p) prod = __fmul_rn(prod, __uint_as_float(lds32(a0 + p2 * stride4)));
        const float x0 = __fmul_rn(-2.0f, atanhf(fminf(fmaxf(prod, -lim), lim)));
        const float w = P.sharing0 == 0 ? 1.0f : (ucn ? ucn_weight(P, t, i, e0 + p) : cn_weight(P, t, i, e0 + p));
        const float x1 = __fmul_rn(fabsf(x0), w);
        const float x2 = fminf(x1 > 0.0f ? x1 : 0.0f, P.clip);
        outv[p] = x0 > 0.0f ? x2 : (x0 < 0.0f ? -x2 : 0.0f);                                           // :316
    }
    for (int p = 0; p < dc; ++p) sts32(a0 + p * stride4, __float_as_uint(outv[p]));
}

// extrinsic sums of a column held in registers, in the order of the reference restatement (oracle/nms_oracle.c): the
// OTHER C->V values added one by one in ascending E(C) order.  The ascending sum that skips edge u starts with the
// prefix c_0 + .. + c_{u-1}, which all edges share with the APP sum: dv (dv - 1) / 2 + dv additions instead of
// dv (dv - 1), bit for bit the same results (a sum that starts from 0.0f differs at most in the sign of a zero, which
// the addition of xin -- never -0.0 -- removes).  Returns the ascending total c_0 + .. + c_{DV-1} (:317).
template <int DV>
__device__ __forceinline__ float f32_extrinsic(const float (&cv)[DV], float (&ext)[DV]) {
    if constexpr (DV == 1) {
        ext[0] = 0.0f;
        return cv[0];
    } else {
        float pre[DV];   // pre[u] = c_0 + .. + c_{u-1}
        pre[1] = cv[0];
#pragma unroll
        for (int u = 2; u < DV; ++u) pre[u] = __fadd_rn(pre[u - 1], cv[u - 1]);
        ext[DV - 1] = pre[DV - 1];
#pragma unroll
        for (int u = 0; u < DV - 1; ++u) {
            float acc = u == 0 ? cv[1] : __fadd_rn(pre[u], cv[u + 1]);
#pragma unroll
            for (int k = (u == 0 ? 2 : u + 2); k < DV; ++k) acc = __fadd_rn(acc, cv[k]);
            ext[u] = acc;
        }
        return __fadd_rn(pre[DV - 1], cv[DV - 1]);
    }
}

// ---- cold path (optional APP output): out of line, minimal arguments
static __device__ __noinline__ void f32_cold(const KParams &P, int j, int t, float app, long long frame0, int nvalid) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, chunk = warp % P.C, q = chunk * 32 + lane;
    if (q < P.L) {   // ya_output{t} = clip(APP) (:324-327)
        Ctx c;
        c.act = 1; c.a_lane = q / P.Fp; c.frame0 = frame0; c.nvalid = nvalid;
        app_store(P, c, j, t, q - c.a_lane * P.Fp, fminf(fmaxf(app, -P.clip), P.clip));
    }
}

// per-variable part of the VN phase (table-driven code): channel value, APP, hard bit, next iteration's xin.
// INIT: the pass before iteration 0 (C->V = 0): writes xq, the hard bit is taken from xin_0 (:181-182).
// The hard decisions go out as one ballot word per column and chunk: hb[buf][j][chunk], buf = t & 1 (INIT: 1).
template <bool INIT, int QM>
__device__ __forceinline__ float f32_var(const KParams &P, const Ctx &c, const F32Ctx &h, int j, int t, float S, bool want_app,
                                         uint32_t &ones) {
    const uint32_t jl4 = (uint32_t)(j * P.LP) * 4u;
    const float xa = ldsf(h.xa4 + jl4);
    float xqv = xa;
    if (f32_is_qms<QM>(P)) {
        if (INIT) {
            xqv = qf(P, xa);                                         // :321-322
            sts32(h.xq4 + jl4, __float_as_uint(xqv));
        } else {
            xqv = ldsf(h.xq4 + jl4);
        }
    }
    const float app = __fadd_rn(xqv, S);                             // :324; clip_LLR (:325) never changes its sign
    const int tn = INIT ? 0 : min(t + 1, P.T_run - 1);               // the last iteration has no successor: its V->C
    float xin = xa;                                                  // words only carry the hard bits to the syndrome pass
    if (P.sharing2 != 0) xin = __fmul_rn(xa, vn_weight(P, tn, j));   // :168-169
    xin = f32_is_qms<QM>(P) ? qf(P, xin) : f32_pos_zero(xin);        // :176-177
    const bool hb = (INIT ? xin : app) >= 0.0f;                      // Print_Functions.py:106
    if (!INIT && hb && j < P.target_n) ones |= 1u;
    const uint32_t b = __ballot_sync(0xffffffffu, c.act && hb);
    if (c.lane == 0) nms_smem[P.off_hb + ((INIT ? 1 : (t & 1)) * P.N + j) * P.C + c.chunk] = b;
    if (!INIT && want_app) f32_cold(P, j, t, app, c.frame0, c.nvalid);
    return xin;
}

// one column of degree DV held in registers; table-driven: P.vn_edge = {e*LP*4, rot*4} (bytes)
template <int DV, bool INIT, bool PAD, int QM>
__device__ __forceinline__ void vn_col_f32(const KParams &P, const Ctx &c, const F32Ctx &h, int j, int t, bool want_app,
                                           uint32_t &ones) {
    const int c0 = P.col_ptr[j];
    const uint32_t L4 = (uint32_t)P.L * 4u;
    uint32_t addr[DV];
    float cv[DV], ext[DV];
#pragma unroll
    for (int u = 0; u < DV; ++u) {
        const int2 ve = P.vn_edge[c0 + u];
        addr[u] = h.sb + (uint32_t)ve.x + f32_rot<PAD>(h, (uint32_t)ve.y, L4);
        cv[u] = INIT ? 0.0f : ldsf(addr[u]);
    }
    float S = 0.0f;
    if (!INIT) S = f32_extrinsic<DV>(cv, ext);
    const float xin = f32_var<INIT, QM>(P, c, h, j, t, S, want_app, ones);
#pragma unroll
    for (int u = 0; u < DV; ++u) sts32(addr[u], f32_v2c<QM>(P, xin, INIT ? 0.0f : ext[u]));
}

// any column degree up to 64 (host-guarded): same arithmetic, chains staged in local arrays
template <bool INIT, int QM>
static __device__ __noinline__ void vn_col_f32_generic(const KParams &P, const Ctx &c, const F32Ctx &h, int j, int t,
                                                       bool want_app, uint32_t &ones) {
    const int c0 = P.col_ptr[j], dv = min(P.col_ptr[j + 1] - c0, 64);
    const uint32_t L4 = (uint32_t)P.L * 4u;
    float cv[64], pre[64];
    float S = 0.0f;
    if (!INIT) {
        for (int u = 0; u < dv; ++u) {
            const int2 ve = P.vn_edge[c0 + u];
            cv[u] = ldsf(h.sb + (uint32_t)ve.x + f32_rot<true>(h, (uint32_t)ve.y, L4));
        }
        pre[0] = 0.0f;
        if (dv > 1) pre[1] = cv[0];
        for (int u = 2; u < dv; ++u) pre[u] = __fadd_rn(pre[u - 1], cv[u - 1]);
        S = dv > 1 ? __fadd_rn(pre[dv - 1], cv[dv - 1]) : (dv == 1 ? cv[0] : 0.0f);
    }
    const float xin = f32_var<INIT, QM>(P, c, h, j, t, S, want_app, ones);
    for (int u = 0; u < dv; ++u) {
        const int2 ve = P.vn_edge[c0 + u];
        float ext = 0.0f;
        if (!INIT && dv > 1) {
            if (u == dv - 1) {
                ext = pre[u];
            } else {
                ext = u == 0 ? cv[1] : __fadd_rn(pre[u], cv[u + 1]);
                for (int k = (u == 0 ? 2 : u + 2); k < dv; ++k) ext = __fadd_rn(ext, cv[k]);
            }
        }
        sts32(h.sb + (uint32_t)ve.x + f32_rot<true>(h, (uint32_t)ve.y, L4), f32_v2c<QM>(P, xin, ext));
    }
}

// table-driven VN phase over this warp's columns (generic kernels; APP-output / shared-memory-init path of the
// specialised ones).  DVB = 0: any degree.
template <int DVB, bool INIT, int QM, int PADMODE = 2>   // PADMODE 0 / 1: L == LP known at compile time, 2: run time
__device__ __forceinline__ void f32_vn_phase_tab(const KParams &P, const Ctx &c, const F32Ctx &h, int t, uint32_t &ones) {
    const bool cold = !INIT && P.app != nullptr;
    const bool pad = PADMODE == 2 ? P.L != P.LP : PADMODE == 1;
    for (int n = c.slot; n < P.N; n += P.R) {
        const int j = P.vn_order[n];
        if constexpr (DVB == 0) {
            vn_col_f32_generic<INIT, QM>(P, c, h, j, t, cold, ones);
        } else {
            const int dv = P.col_ptr[j + 1] - P.col_ptr[j];
            if (pad) {
                switch (dv) {
#define X(p)                                                                              \
    case (p) + 1:                                                                         \
        if constexpr ((p) < DVB && PADMODE != 0) vn_col_f32<(p) + 1, INIT, true, QM>(P, c, h, j, t, cold, ones); \
        break;
                    NMS_REP_DESC(X)
#undef X
                default: break;
                }
            } else {
                switch (dv) {
#define X(p)                                                                               \
    case (p) + 1:                                                                          \
        if constexpr ((p) < DVB && PADMODE != 1) vn_col_f32<(p) + 1, INIT, false, QM>(P, c, h, j, t, cold, ones); \
        break;
                    NMS_REP_DESC(X)
#undef X
                default: break;
                }
            }
        }
    }
}

// syndrome of the last hard decision (ballot buffer `buf`)
__device__ __forceinline__ uint32_t f32_synd_phase(const KParams &P, const Ctx &c, int buf) {
    const F32Ctx h = f32_ctx(P, c);
    const uint32_t hb4 = h.sb + (uint32_t)(P.off_hb + buf * P.N * P.C) * 4u;
    const bool pad = P.L != P.LP;
    uint32_t bad = 0;
    for (int n = c.slot; n < P.M; n += P.R) {
        const int i = P.cn_order[n];
        const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
        bad |= pad ? f32_row_syndrome<true>(P, h, hb4, e0, dc) : f32_row_syndrome<false>(P, h, hb4, e0, dc);
    }
    return bad;
}

}   // namespace nms
