// nms_h2.cuh -- packed fp16x2 arithmetic (two frames per 32-bit word) shared by the degree-bucketed
// generic kernels (nms_h2.cu) and the graph-specialised kernels (nms_h2_spec.cuh).
//
// Why fp16x2 is exact here: in quantised min-sum (decoding_type 2, q_bit 5/-5/4/3) every message is a
// multiple of the quantiser step and every partial sum stays below 512 steps, all of which fp16
// represents exactly, so HADD2 / HMNMX2 / HSET2 reproduce the reference's float32 results bit for bit --
// at two frames per instruction.  The products with the trained weights and the quantiser rounding (the
// only inexact steps) are done in float32 exactly as the reference does them (Main_Functions.py:267-311,
// 483-492).
//
// Two identities keep the inner loops short:
//  * the hard decision of each variable rides in the (always free) mantissa LSB of its outgoing V->C
//    message, so ONE XOR per edge gives the CN phase both the sign parity and the syndrome of the
//    previous hard decision that selects the unsatisfied-check weights (:180-206);
//  * the V->C saturation Q() of :223-224 is a clamp of an on-grid value; clamping commutes with the
//    minimum and does not touch the sign, so it is applied once per check to (min1, min2) instead of
//    once per edge.
// All shared-memory traffic of the inner loops goes through ld/st.shared with 32-bit byte addresses
// (H2Ctx::sb + offset), so every access is one LDS/STS [reg + immediate].
#pragma once
#include "nms_device.cuh"

namespace nms {

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float ldsf(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }

// per-thread constants of the packed kernels (byte units)
struct H2Ctx {
    uint32_t sb;      // shared-window byte address of nms_smem[0]
    uint32_t q4;      // q * 4
    uint32_t amask;   // all ones for active lanes, 0 for padding lanes (they never rotate)
    uint32_t Lthr4;   // L*4 for active lanes, 2^30 for padding lanes (they never wrap)
    uint32_t xa8;     // sb + off_xa*4 + q*8   (+ j*LP*8 per column)
    uint32_t xq4;     // sb + off_xq*4 + q*4   (+ j*LP*4 per column)
};

// shared-window address of nms_smem[0].  volatile: computed where written, once -- left to itself the compiler rematerialises
// the conversion (S2R SR_CgaCtaId + LEA, ~25 cycles of latency) next to every use that is short of registers
__device__ __forceinline__ uint32_t smem_base() {
    uint32_t sb;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sb) : "l"((const void *)nms_smem));
    return sb;
}

__device__ __forceinline__ H2Ctx h2_ctx(const KParams &P, const Ctx &c) {
    H2Ctx h;
    h.sb = smem_base();
    h.q4 = (uint32_t)c.q * 4u;
    h.amask = c.act ? 0xffffffffu : 0u;
    h.Lthr4 = c.act ? (uint32_t)P.L * 4u : 0x40000000u;
    h.xa8 = h.sb + (uint32_t)P.off_xa * 4u + (uint32_t)c.q * 8u;
    h.xq4 = h.sb + (uint32_t)P.off_xq * 4u + h.q4;
    return h;
}

// byte address of the weight row of iteration t (branch-free: "no weight" is a staged row of ones)
__device__ __forceinline__ uint32_t h2_wrow(const H2Ctx &h, int base_word, int t, int width) {
    return h.sb + (uint32_t)(base_word + t * width) * 4u;
}
__device__ __forceinline__ float h2_w(uint32_t row, int node, int mask) { return ldsf(row + (uint32_t)((node & mask) * 4)); }

// variable lane -> byte offset of the rotated check lane: (q + rot) mod L, padding lanes stay put
template <bool PAD>
__device__ __forceinline__ uint32_t h2_rot(const H2Ctx &h, uint32_t rot4, uint32_t L4) {
    if (PAD) {
        const uint32_t t1 = h.q4 + (rot4 & h.amask);
        return min(t1, t1 - h.Lthr4);          // t1 - Lthr4 wraps high unless the lane must wrap
    } else {
        const uint32_t t1 = h.q4 + rot4;
        return min(t1, t1 - L4);
    }
}

// Q() of two float32 values -> half2 (round to step in fp32, saturate after packing; +-inf saturate too)
__device__ __forceinline__ __half2 q2(const KParams &P, float lo, float hi) {
    const __half2 qm = u2h(P.qmax_h2);
    const float2 q = qround2(make_float2(lo, hi), P.qmagic);
    const __half2 r = __floats2half2_rn(q.x, q.y);
    return __hmax2(__hmin2(r, qm), __hneg2(qm));
}

// weighted, quantised magnitudes for "edge is not the minimum" (A) and "edge is the minimum" (B), with the
// row's sign parity folded in.  par: XOR of all raw V->C words of the row; w0 / w1: CN / UCN weight.
// The two frames of a lane may be at different iterations (persistent-slot Monte-Carlo kernel, nms_mcp.cuh), hence one
// (CN, UCN) weight pair per half.
__device__ __forceinline__ void h2_row_mags(const KParams &P, float w0lo, float w0hi, float w1lo, float w1hi, bool dc_odd,
                                            uint32_t par, __half2 m1, __half2 m2, uint32_t &A, uint32_t &B) {
    const __half2 qm = u2h(P.qmax_h2), zero = __float2half2_rn(0.0f);
    // drop the piggy-backed hard bits; V->C saturation (:223-224) applied to the two minima
    const __half2 m1c = __hmin2(u2h(h2u(m1) & ~LSB2), qm), m2c = __hmin2(u2h(h2u(m2) & ~LSB2), qm);
    const float wlo = (par & 1u) ? w1lo : w0lo;     // unsatisfied check -> UCN weight (:275,:285,:295)
    const float whi = (par & 0x10000u) ? w1hi : w0hi;
    // Q(relu(min * w)) (:308-311)
    const float2 w2 = make_float2(wlo, whi);
    const float2 qa = qround2(mul2_rn_unfused(__half22float2(m1c), w2), P.qmagic);
    const float2 qb = qround2(mul2_rn_unfused(__half22float2(m2c), w2), P.qmagic);
    __half2 magA = __floats2half2_rn(qa.x, qa.y);
    __half2 magB = __floats2half2_rn(qb.x, qb.y);
    magA = __hmax2(__hmin2(magA, qm), zero);
    magB = __hmax2(__hmin2(magB, qm), zero);
    // C->V is negative iff (dc + #negative others) is odd (:251-254; a zero V->C counts as positive, :230)
    const uint32_t s0 = (par & SIGN2) ^ (dc_odd ? SIGN2 : 0u);
    A = h2u(magA) ^ s0;
    B = h2u(magB) ^ s0;
}
__device__ __forceinline__ void h2_row_mags(const KParams &P, float w0, float w1, bool dc_odd, uint32_t par,
                                            __half2 m1, __half2 m2, uint32_t &A, uint32_t &B) {
    h2_row_mags(P, w0, w0, w1, w1, dc_odd, par, m1, m2, A, B);
}

// (min1, min2) of |raw[LO..HI)| as a tournament: pairs are sorted with one min and one max, two sorted pairs merge with
// min, max and one three-input min.  35 operations for 15 edges instead of 45 for the running update, on the busiest
// pipe of the kernel, and a dependency depth of log2(dc) levels instead of dc / 2.  Exact (minimum and maximum are).
template <int DC, int LO, int HI>
__device__ __forceinline__ void h2_min12(const uint32_t (&raw)[DC], __half2 &m1, __half2 &m2) {
    if constexpr (HI - LO == 1) {
        m1 = __habs2(u2h(raw[LO]));
        m2 = __float2half2_rn(10000.0f);   // all-masked row -> 10000 (:248)
    } else if constexpr (HI - LO == 2) {
        const __half2 a = __habs2(u2h(raw[LO])), b = __habs2(u2h(raw[LO + 1]));
        m1 = __hmin2(a, b);
        m2 = __hmax2(a, b);
    } else if constexpr (HI - LO == 3) {
        h2_min12<DC, LO, LO + 2>(raw, m1, m2);
        const __half2 a = __habs2(u2h(raw[LO + 2]));
        const __half2 t = __hmax2(m1, a);
        m1 = __hmin2(m1, a);
        m2 = __hmin2(m2, t);
    } else {
        constexpr int MID = LO + (((HI - LO) / 2 + 1) & ~1);   // even-sized left half: its leaves are pairs
        __half2 a1, a2, b1, b2;
        h2_min12<DC, LO, MID>(raw, a1, a2);
        h2_min12<DC, MID, HI>(raw, b1, b2);
        const __half2 hi = __hmax2(a1, b1);
        m1 = __hmin2(a1, b1);
        m2 = __hmin2(__hmin2(hi, a2), b2);
    }
}

// sum of the packed words cv[LO..HI) as a balanced tree: log2(n) dependent additions instead of n - 1 (the VN phase waits on
// fixed-latency dependencies more than on anything else).  Exact in any order: on-grid values, |partial sums| < 512 steps.
template <int DV, int LO, int HI>
__device__ __forceinline__ __half2 h2_tree_sum(const uint32_t (&cv)[DV]) {
    if constexpr (HI - LO == 1) {
        return u2h(cv[LO]);
    } else {
        constexpr int MID = LO + (HI - LO + 1) / 2;
        return __hadd2(h2_tree_sum<DV, LO, MID>(cv), h2_tree_sum<DV, MID, HI>(cv));
    }
}

// one check row held in registers.  a0: byte address of msg[e0][q]; stride4: bytes between edges (LP*4)
template <int DC>
__device__ __forceinline__ void cn_row_h2(const KParams &P, uint32_t a0, uint32_t stride4, float w0lo, float w0hi,
                                          float w1lo, float w1hi, uint32_t &bad) {
    uint32_t raw[DC];
#pragma unroll
    for (int p = 0; p < DC; ++p) raw[p] = lds32(a0 + p * stride4);
    uint32_t par = 0;
#pragma unroll
    for (int p = 0; p < DC; ++p) par ^= raw[p];
    bad |= par;
    __half2 m1, m2;
    h2_min12<DC, 0, DC>(raw, m1, m2);
    uint32_t A, B;
    h2_row_mags(P, w0lo, w0hi, w1lo, w1hi, (DC & 1) != 0, par, m1, m2, A, B);
    // |v| > min1 -> the others' minimum is min1 (A), else min2 (B).  The select is done as B + g (A - B) with g in {0, 1}
    // on the FMA pipe instead of a second LOP3: the integer / logic pipe runs at half the issue rate and is the busiest
    // pipe of this kernel.  Exact: A, B are on-grid values of one sign, so A - B and the sum are representable.
    const __half2 dAB = __hsub2(u2h(A), u2h(B));
#pragma unroll
    for (int p = 0; p < DC; ++p) {
        const __half2 g = __hgt2(__habs2(u2h(raw[p])), m1);
        sts32(a0 + p * stride4, h2u(__hfma2(g, dAB, u2h(B))) ^ (raw[p] & SIGN2));
    }
}

template <int DC>
__device__ __forceinline__ void cn_row_h2(const KParams &P, uint32_t a0, uint32_t stride4, float w0, float w1,
                                          uint32_t &bad) {
    cn_row_h2<DC>(P, a0, stride4, w0, w0, w1, w1, bad);
}

// any degree: two passes over shared memory instead of a register array
__device__ __forceinline__ void cn_row_h2_generic(const KParams &P, uint32_t a0, uint32_t stride4, int dc, float w0,
                                                      float w1, uint32_t &bad) {
    uint32_t par = 0;
    __half2 m1 = __float2half2_rn(10000.0f), m2 = m1;
    for (int p = 0; p < dc; ++p) {
        const uint32_t r = lds32(a0 + p * stride4);
        par ^= r;
        const __half2 a = __habs2(u2h(r));
        const __half2 tmx = __hmax2(m1, a);
        m1 = __hmin2(m1, a);
        m2 = __hmin2(m2, tmx);
    }
    bad |= par;
    uint32_t A, B;
    h2_row_mags(P, w0, w1, (dc & 1) != 0, par, m1, m2, A, B);
    for (int p = 0; p < dc; ++p) {
        const uint32_t r = lds32(a0 + p * stride4);
        const uint32_t gt = __hgt2_mask(__habs2(u2h(r)), m1);
        sts32(a0 + p * stride4, ((gt & A) | (~gt & B)) ^ (r & SIGN2));
    }
}

// ---- cold path (hard-decision ballots for a possible copy-out, optional float APP output), kept out of
// line with minimal arguments so the iteration loop stays small in the instruction cache
static __device__ __noinline__ void h2_cold(const KParams &P, int j, int t, uint32_t hbw, uint32_t app, int what,
                                            long long frame0, int nvalid) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, chunk = warp % P.C, q = chunk * 32 + lane;
    const int act = q < P.L;
    if (what & 1) {   // ballots -> hb[buf][half][j][chunk]
        const int buf = t < 0 ? 1 : (t & 1);
        const uint32_t lo = __ballot_sync(0xffffffffu, act && (hbw & 1u));
        const uint32_t hi = __ballot_sync(0xffffffffu, act && (hbw >> 16));
        if (lane == 0) {
            nms_smem[P.off_hb + ((buf * 2 + 0) * P.N + j) * P.C + chunk] = lo;
            nms_smem[P.off_hb + ((buf * 2 + 1) * P.N + j) * P.C + chunk] = hi;
        }
    }
    if ((what & 2) && act) {   // ya_output{t} = clip(APP) (:324-327)
        Ctx c;
        c.act = act; c.a_lane = q / P.Fp; c.frame0 = frame0; c.nvalid = nvalid;
        const int fp = q - c.a_lane * P.Fp;
        app_store(P, c, j, t, 2 * fp, fminf(fmaxf(__low2float(u2h(app)), -P.clip), P.clip));
        app_store(P, c, j, t, 2 * fp + 1, fminf(fmaxf(__high2float(u2h(app)), -P.clip), P.clip));
    }
}

// per-variable part of the VN phase, for NC independent columns at once (NC = 2 doubles the ILP)
struct H2Var {
    __half2 xin;      // next iteration's weighted + quantised channel value
    uint32_t hbw;     // hard bits (bit 0 / bit 16) of this slot's two frames
};

// j[k]: column, jlp[k] = j*LP (words); wvrow: byte address of the next iteration's VN weight row; cold: bit 0 =
// ballots wanted, bit 1 = APP output wanted.  INIT: the pass before iteration 0 (C->V = 0): writes xq, and the
// hard bit is taken from xin_0.  Returns has_next.
template <bool INIT, int NC>
__device__ __forceinline__ bool h2_var(const KParams &P, const Ctx &c, const H2Ctx &h, const int (&j)[NC],
                                       const int (&jlp)[NC], int t, const __half2 (&S)[NC], uint32_t wvrow, int cold,
                                       uint32_t &ones, H2Var (&v)[NC]) {
    float2 x[NC];
    __half2 xqh[NC], app[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) x[k] = lds64f(h.xa8 + (uint32_t)jlp[k] * 8u);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        if (INIT || P.no_xq) {   // no_xq: Q(xa) is recomputed from xa instead of kept in its own array
            xqh[k] = q2(P, x[k].x, x[k].y);                       // Q(xa), :321-322
            if (!P.no_xq) sts32(h.xq4 + (uint32_t)jlp[k] * 4u, h2u(xqh[k]));
        } else {
            xqh[k] = u2h(lds32(h.xq4 + (uint32_t)jlp[k] * 4u));
        }
    }
    const int tn = INIT ? 0 : t + 1;
    const bool has_next = tn < P.T_run;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        app[k] = __hadd2(xqh[k], S[k]);   // unclipped APP; clip_LLR never changes its sign
        v[k].xin = xqh[k];
    }
    if (has_next && P.sharing2 != 0) {
        float w[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) w[k] = h2_w(wvrow, j[k], P.h2_mv);
#pragma unroll
        for (int k = 0; k < NC; ++k)
            v[k].xin = q2(P, __fmul_rn(x[k].x, w[k]), __fmul_rn(x[k].y, w[k]));   // Q(xa * w), :168-177
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const __half2 hsrc = INIT ? v[k].xin : app[k];   // iteration 0 takes the syndrome of xin_0 (:181-182)
        v[k].hbw = (~h2u(hsrc) >> 15) & LSB2;            // bit = (value >= 0); a zero here is always +0
        if (!INIT && j[k] < P.target_n) ones |= v[k].hbw;
    }
    if (cold) {
#pragma unroll
        for (int k = 0; k < NC; ++k) h2_cold(P, j[k], t, v[k].hbw, h2u(app[k]), cold, c.frame0, c.nvalid);
    }
    return has_next;
}

// NC variable columns of degree DV held in registers; table-driven (generic kernels): P.vn_edge = {e*LP*4, rot*4}
template <int DV, bool INIT, bool PAD, int NC>
__device__ __forceinline__ void vn_col_h2(const KParams &P, const Ctx &c, const H2Ctx &h, const int (&j)[NC], int t,
                                          uint32_t wvrow, int cold, uint32_t &ones) {
    const uint32_t L4 = (uint32_t)P.L * 4u;
    uint32_t addr[NC][DV], cv[NC][DV];
    int jlp[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c0 = P.col_ptr[j[k]];
        jlp[k] = j[k] * P.LP;
#pragma unroll
        for (int u = 0; u < DV; ++u) {
            const int2 ve = P.vn_edge[c0 + u];
            addr[k][u] = h.sb + (uint32_t)ve.x + h2_rot<PAD>(h, (uint32_t)ve.y, L4);
            cv[k][u] = INIT ? 0u : lds32(addr[k][u]);
        }
    }
    __half2 S[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        S[k] = __float2half2_rn(0.0f);
        if (!INIT) {
#pragma unroll
            for (int u = 0; u < DV; ++u) S[k] = __hadd2(S[k], u2h(cv[k][u]));
        }
    }
    H2Var v[NC];
    const bool has_next = h2_var<INIT, NC>(P, c, h, j, jlp, t, S, wvrow, cold, ones, v);
    if (has_next) {
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const __half2 SX = __hadd2(v[k].xin, S[k]);
#pragma unroll
            for (int u = 0; u < DV; ++u) {
                const __half2 m = INIT ? v[k].xin : __hsub2(SX, u2h(cv[k][u]));   // total - self: exact on the grid
                sts32(addr[k][u], h2u(m) | v[k].hbw);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < NC; ++k)
#pragma unroll
            for (int u = 0; u < DV; ++u) sts32(addr[k][u], v[k].hbw);   // only the final syndrome pass reads these
    }
}

template <bool INIT>
__device__ __forceinline__ void vn_col_h2_generic(const KParams &P, const Ctx &c, const H2Ctx &h, int j, int t,
                                                  uint32_t wvrow, int cold, uint32_t &ones) {
    const int c0 = P.col_ptr[j], dv = P.col_ptr[j + 1] - c0;
    const uint32_t L4 = (uint32_t)P.L * 4u;
    __half2 S[1] = {__float2half2_rn(0.0f)};
    if (!INIT)
        for (int u = 0; u < dv; ++u) {
            const int2 ve = P.vn_edge[c0 + u];
            S[0] = __hadd2(S[0], u2h(lds32(h.sb + (uint32_t)ve.x + h2_rot<true>(h, (uint32_t)ve.y, L4))));
        }
    const int jj[1] = {j}, jlp[1] = {j * P.LP};
    H2Var v[1];
    const bool has_next = h2_var<INIT, 1>(P, c, h, jj, jlp, t, S, wvrow, cold, ones, v);
    const __half2 SX = __hadd2(v[0].xin, S[0]);
    for (int u = 0; u < dv; ++u) {
        const int2 ve = P.vn_edge[c0 + u];
        const uint32_t a = h.sb + (uint32_t)ve.x + h2_rot<true>(h, (uint32_t)ve.y, L4);
        uint32_t o = v[0].hbw;
        if (has_next) o |= h2u(INIT ? v[0].xin : __hsub2(SX, u2h(lds32(a))));
        sts32(a, o);
    }
}

// syndrome parity of the hard bits parked in the message LSBs after the last VN phase
__device__ __forceinline__ uint32_t h2_synd_phase(const KParams &P, const Ctx &c) {
    uint32_t bad = 0;
    for (int n = c.slot; n < P.M; n += P.R) {
        const int i = P.cn_order[n];
        const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
        uint32_t par = 0;
        for (int p = 0; p < dc; ++p) par ^= nms_smem[(e0 + p) * P.LP + c.q];
        bad |= par;
    }
    return bad;
}

// per-phase constants of the VN phase
__device__ __forceinline__ int h2_cold_mask(const KParams &P, bool init, bool need_hb) {
    return (need_hb ? 1 : 0) | ((!init && P.app != nullptr) ? 2 : 0);
}

// table-driven VN phase over this warp's columns (generic kernels; INIT pass of the specialised ones).
// Columns are sorted by degree; slot s owns positions p % R == s, walked with one running position, so the degree
// dispatch happens once per degree class and columns of one class are processed two at a time.
template <int DV, bool INIT, bool PAD>
__device__ __forceinline__ void h2_vn_class(const KParams &P, const Ctx &c, const H2Ctx &h, int &p, int end, int t,
                                            uint32_t wvrow, int cold, uint32_t &ones) {
    const int R = P.R;
    if constexpr (DV <= 3) {
        while (p + R < end) {
            const int jj[2] = {P.vn_order[p], P.vn_order[p + R]};
            vn_col_h2<DV, INIT, PAD, 2>(P, c, h, jj, t, wvrow, cold, ones);
            p += 2 * R;
        }
    }
    while (p < end) {
        const int jj[1] = {P.vn_order[p]};
        vn_col_h2<DV, INIT, PAD, 1>(P, c, h, jj, t, wvrow, cold, ones);
        p += R;
    }
}

template <int DVB, bool INIT>
__device__ __forceinline__ void h2_vn_phase_tab(const KParams &P, const Ctx &c, const H2Ctx &h, int t, bool need_hb,
                                                uint32_t &ones) {
    const bool pad = P.L != P.LP;
    const uint32_t wvrow = h2_wrow(h, P.h2w_v, INIT ? 0 : t + 1, P.h2_wv);
    const int cold = h2_cold_mask(P, INIT, need_hb);
    if constexpr (DVB == 0) {
        for (int n = c.slot; n < P.N; n += P.R) vn_col_h2_generic<INIT>(P, c, h, P.vn_order[n], t, wvrow, cold, ones);
    } else {
        int p = c.slot;
        for (int k = 0; k < P.n_vn_cls; ++k) {
            const ushort4 cl = P.vn_cls[k];
            const int end = cl.z;
            if (p >= end) continue;
            if (pad) {
                switch (cl.x) {
#define X(d)                                                                                          \
    case (d) + 1:                                                                                     \
        if constexpr ((d) < DVB) h2_vn_class<(d) + 1, INIT, true>(P, c, h, p, end, t, wvrow, cold, ones); \
        break;
                    NMS_REP_DESC(X)
#undef X
                default: break;
                }
            } else {
                switch (cl.x) {
#define X(d)                                                                                           \
    case (d) + 1:                                                                                      \
        if constexpr ((d) < DVB) h2_vn_class<(d) + 1, INIT, false>(P, c, h, p, end, t, wvrow, cold, ones); \
        break;
                    NMS_REP_DESC(X)
#undef X
                default: break;
                }
            }
        }
    }
}

}   // namespace nms
