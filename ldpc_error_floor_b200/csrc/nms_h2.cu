// nms_h2.cu -- packed fp16x2 arithmetic back-end (two frames per 32-bit word).
//
// Compiled once per degree bucket: -DNMS_DCB=<max row degree> -DNMS_DVB=<max column degree>
// (0/0 = any degree, runtime loops over shared memory instead of register arrays).
//
// Why fp16x2 is exact here: in quantised min-sum (decoding_type 2) every message is a multiple
// of the quantiser step with magnitude <= 15.5 and every partial sum stays below 512, all of
// which fp16 represents exactly, so HADD2 / HMNMX2 / HSET2 reproduce the reference's float32
// results bit for bit -- at two frames per instruction.  The products with the trained weights
// and the quantiser rounding (the only inexact steps) are done in float32, exactly as the
// reference does them (Main_Functions.py:267-311, 483-492).
// The hard decision of each variable rides in the (always free) mantissa LSB of its outgoing
// V->C message, so one XOR per edge gives the CN phase both the sign parity and the syndrome of
// the previous hard decision that selects the unsatisfied-check weights (:180-206).
#include "nms_device.cuh"

#ifndef NMS_DCB
#define NMS_DCB 16
#endif
#ifndef NMS_DVB
#define NMS_DVB 8
#endif

namespace nms {

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

// two values (already scaled by qk) -> quantised half2
__device__ __forceinline__ __half2 qsym2(const KParams &P, float tlo, float thi) {
    return __hmul2(__floats2half2_rn(qcore(tlo, P.qmaxk), qcore(thi, P.qmaxk)), __float2half2_rn(P.qinv));
}
__device__ __forceinline__ __half2 qpos2(const KParams &P, float tlo, float thi) {   // Q(relu(.)), :308-311
    const float a = fminf(fmaxf(rint_magic(tlo), 0.0f), P.qmaxk);
    const float b = fminf(fmaxf(rint_magic(thi), 0.0f), P.qmaxk);
    return __hmul2(__floats2half2_rn(a, b), __float2half2_rn(P.qinv));
}

// weighted, quantised magnitudes for "edge is not the minimum" (A) and "edge is the minimum" (B),
// with the row's sign parity folded in.  par: XOR of all raw V->C words of the row.
__device__ __forceinline__ void h2_row_mags(const KParams &P, int i, int t, int dc, uint32_t par, __half2 m1,
                                            __half2 m2, uint32_t &A, uint32_t &B) {
    const __half2 m1c = u2h(h2u(m1) & ~LSB2), m2c = u2h(h2u(m2) & ~LSB2);   // drop the piggy-backed hard bits
    float wk0 = P.qk, wk1;
    if (P.sharing0 != 0) wk0 = __fmul_rn(cn_w(P.w_cn, P.sharing0, P.wc, t, i, 0), P.qk);
    wk1 = wk0;
    if (P.sharing1 != 0) wk1 = __fmul_rn(cn_w(P.w_ucn, P.sharing1, P.wu, t, i, 0), P.qk);
    const float wlo = (par & 1u) ? wk1 : wk0;         // unsatisfied check -> UCN weight (:275,:285,:295)
    const float whi = (par & 0x10000u) ? wk1 : wk0;
    const __half2 magA = qpos2(P, __fmul_rn(__low2float(m1c), wlo), __fmul_rn(__high2float(m1c), whi));
    const __half2 magB = qpos2(P, __fmul_rn(__low2float(m2c), wlo), __fmul_rn(__high2float(m2c), whi));
    // C->V is negative iff (dc + #negative others) is odd (:251-254; a zero V->C counts as positive, :230)
    const uint32_t s0 = (par & SIGN2) ^ ((dc & 1) ? SIGN2 : 0u);
    A = h2u(magA) ^ s0;
    B = h2u(magB) ^ s0;
}

template <int DC>
__device__ __forceinline__ void cn_row_h2(const KParams &P, const Ctx &c, int i, int t, uint32_t &bad) {
    const int LP = P.LP;
    uint32_t *base = c.msg + P.row_ptr[i] * LP + c.qe;
    uint32_t raw[DC];
#pragma unroll
    for (int p = 0; p < DC; ++p) raw[p] = base[p * LP];
    uint32_t par = 0;
#pragma unroll
    for (int p = 0; p < DC; ++p) par ^= raw[p];
    bad |= par;
    __half2 m1 = __float2half2_rn(10000.0f), m2 = m1;   // all-masked row -> 10000 (:248)
#pragma unroll
    for (int p = 0; p < DC; ++p) {
        const __half2 a = __habs2(u2h(raw[p]));
        const __half2 tmx = __hmax2(m1, a);
        m1 = __hmin2(m1, a);
        m2 = __hmin2(m2, tmx);
    }
    uint32_t A, B;
    h2_row_mags(P, i, t, DC, par, m1, m2, A, B);
#pragma unroll
    for (int p = 0; p < DC; ++p) {
        const uint32_t gt = __hgt2_mask(__habs2(u2h(raw[p])), m1);   // |v| > min1 -> others' min is min1, else min2
        const uint32_t out = ((gt & A) | (~gt & B)) ^ (raw[p] & SIGN2);
        if (c.active) base[p * LP] = out;
    }
}

// any degree: two passes over shared memory instead of a register array
static __device__ __noinline__ void cn_row_h2_generic(const KParams &P, const Ctx &c, int i, int t, uint32_t &bad) {
    const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0, LP = P.LP;
    uint32_t *base = c.msg + e0 * LP + c.qe;
    uint32_t par = 0;
    __half2 m1 = __float2half2_rn(10000.0f), m2 = m1;
    for (int p = 0; p < dc; ++p) {
        const uint32_t r = base[p * LP];
        par ^= r;
        const __half2 a = __habs2(u2h(r));
        const __half2 tmx = __hmax2(m1, a);
        m1 = __hmin2(m1, a);
        m2 = __hmin2(m2, tmx);
    }
    bad |= par;
    uint32_t A, B;
    h2_row_mags(P, i, t, dc, par, m1, m2, A, B);
    for (int p = 0; p < dc; ++p) {
        const uint32_t r = base[p * LP];
        const uint32_t gt = __hgt2_mask(__habs2(u2h(r)), m1);
        const uint32_t out = ((gt & A) | (~gt & B)) ^ (r & SIGN2);
        if (c.active) base[p * LP] = out;
    }
}

// per-variable part shared by the unrolled and the generic column update
struct H2Var {
    __half2 xin, S;   // next iteration's weighted+quantised channel value; sum of incoming C->V
    uint32_t hbw;     // hard bits (bit 0 / bit 16) of this slot's two frames
    bool has_next;
};

template <bool INIT>
__device__ __forceinline__ H2Var h2_var(const KParams &P, const Ctx &c, int j, int t, __half2 S, uint32_t &ones) {
    H2Var v;
    v.S = S;
    const int slotw = j * P.LP + c.qe;
    const float2 x = reinterpret_cast<const float2 *>(c.xa)[slotw];
    __half2 xqh;
    if (INIT) {
        xqh = qsym2(P, __fmul_rn(x.x, P.qk), __fmul_rn(x.y, P.qk));   // Q(xa), :321-322
        if (c.active) c.xq[slotw] = h2u(xqh);
    } else {
        xqh = u2h(c.xq[slotw]);
    }
    const __half2 app = __hadd2(xqh, S);   // unclipped APP; clip_LLR never changes its sign
    const int tn = INIT ? 0 : t + 1;
    v.has_next = tn < P.T_run;
    v.xin = xqh;
    if (v.has_next && P.sharing2 != 0) {
        const float wk = __fmul_rn(vn_w(P, tn, j), P.qk);   // (xa*w)*qk == xa*(w*qk): qk is a power of two
        v.xin = qsym2(P, __fmul_rn(x.x, wk), __fmul_rn(x.y, wk));     // :168-177
    }
    const __half2 hsrc = INIT ? v.xin : app;   // iteration 0 takes the syndrome of xin_0 (:181-182)
    v.hbw = (~h2u(hsrc) >> 15) & LSB2;         // bit = (value >= 0); a zero here is always +0
    if (!INIT) ones |= v.hbw;
    const uint32_t lo = __ballot_sync(0xffffffffu, c.active && (v.hbw & 1u));
    const uint32_t hi = __ballot_sync(0xffffffffu, c.active && (v.hbw >> 16));
    if (c.lane == 0) {
        const int buf = INIT ? 1 : (t & 1);
        c.hb[((buf * 2 + 0) * P.N + j) * P.C + c.chunk] = lo;
        c.hb[((buf * 2 + 1) * P.N + j) * P.C + c.chunk] = hi;
    }
    if (!INIT && P.app != nullptr && c.active) {   // optional float APP output (ya_output{t}, :324-327)
        app_store(P, c, j, t, c.f0, fminf(fmaxf(__low2float(app), -P.clip), P.clip));
        app_store(P, c, j, t, c.f1, fminf(fmaxf(__high2float(app), -P.clip), P.clip));
    }
    return v;
}

// INIT: the pass before iteration 0 (C->V = 0): writes xq, V->C = Q(xa*wv_0), hard bit of xin_0.
template <int DV, bool INIT>
__device__ __forceinline__ void vn_col_h2(const KParams &P, const Ctx &c, int j, int t, uint32_t &ones) {
    const int c0 = P.col_ptr[j], L = P.L;
    int addr[DV];
    uint32_t cv[DV];
#pragma unroll
    for (int u = 0; u < DV; ++u) {
        const int2 ve = P.vn_edge[c0 + u];
        int qq = c.qe + ve.y;
        qq = (qq >= L) ? qq - L : qq;
        addr[u] = ve.x + qq;
        cv[u] = INIT ? 0u : c.msg[addr[u]];
    }
    __half2 S = __float2half2_rn(0.0f);
    if (!INIT) {
#pragma unroll
        for (int u = 0; u < DV; ++u) S = __hadd2(S, u2h(cv[u]));
    }
    const H2Var v = h2_var<INIT>(P, c, j, t, S, ones);
    if (v.has_next) {
        const __half2 SX = __hadd2(v.xin, S);
        const __half2 hi = __float2half2_rn(P.qmax), lo = __float2half2_rn(-P.qmax);
#pragma unroll
        for (int u = 0; u < DV; ++u) {
            __half2 m = INIT ? v.xin : __hsub2(SX, u2h(cv[u]));     // total - self: exact on the grid (:213-215)
            m = __hmax2(__hmin2(m, hi), lo);                        // Q() of an on-grid value is a clamp (:223-224)
            if (c.active) c.msg[addr[u]] = h2u(m) | v.hbw;
        }
    } else {
#pragma unroll
        for (int u = 0; u < DV; ++u)
            if (c.active) c.msg[addr[u]] = v.hbw;                   // only the final syndrome pass reads these
    }
}

template <bool INIT>
__device__ __noinline__ void vn_col_h2_generic(const KParams &P, const Ctx &c, int j, int t, uint32_t &ones) {
    const int c0 = P.col_ptr[j], dv = P.col_ptr[j + 1] - c0, L = P.L;
    __half2 S = __float2half2_rn(0.0f);
    if (!INIT)
        for (int u = 0; u < dv; ++u) {
            const int2 ve = P.vn_edge[c0 + u];
            int qq = c.qe + ve.y;
            qq = (qq >= L) ? qq - L : qq;
            S = __hadd2(S, u2h(c.msg[ve.x + qq]));
        }
    const H2Var v = h2_var<INIT>(P, c, j, t, S, ones);
    const __half2 SX = __hadd2(v.xin, S);
    const __half2 hi = __float2half2_rn(P.qmax), lo = __float2half2_rn(-P.qmax);
    for (int u = 0; u < dv; ++u) {
        const int2 ve = P.vn_edge[c0 + u];
        int qq = c.qe + ve.y;
        qq = (qq >= L) ? qq - L : qq;
        uint32_t o = v.hbw;
        if (v.has_next) {
            __half2 m = INIT ? v.xin : __hsub2(SX, u2h(c.msg[ve.x + qq]));
            m = __hmax2(__hmin2(m, hi), lo);
            o |= h2u(m);
        }
        if (c.active) c.msg[ve.x + qq] = o;
    }
}

template <int DCB, int DVB>
struct H2Policy {
    static constexpr bool H2 = true;

    static __device__ __forceinline__ void cn_task(const KParams &P, const Ctx &c, int i, int t, uint32_t &bad) {
        if constexpr (DCB == 0) {
            cn_row_h2_generic(P, c, i, t, bad);
        } else {
            const int dc = P.row_ptr[i + 1] - P.row_ptr[i];
            switch (dc) {
#define X(p)                                                       \
    case (p) + 1:                                                  \
        if constexpr ((p) < DCB) cn_row_h2<(p) + 1>(P, c, i, t, bad); \
        break;
                NMS_REP_DESC(X)
#undef X
            default: break;
            }
        }
    }

    template <bool INIT>
    static __device__ __forceinline__ void vn_task(const KParams &P, const Ctx &c, int j, int t, uint32_t &ones) {
        if constexpr (DVB == 0) {
            vn_col_h2_generic<INIT>(P, c, j, t, ones);
        } else {
            const int dv = P.col_ptr[j + 1] - P.col_ptr[j];
            switch (dv) {
#define X(p)                                                                \
    case (p) + 1:                                                           \
        if constexpr ((p) < DVB) vn_col_h2<(p) + 1, INIT>(P, c, j, t, ones); \
        break;
                NMS_REP_DESC(X)
#undef X
            default: break;
            }
        }
    }

    // syndrome parity of the hard bits parked in the message LSBs after the last VN phase
    static __device__ __forceinline__ uint32_t synd_row(const KParams &P, const Ctx &c, int i, int tl) {
        const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
        uint32_t par = 0;
        for (int p = 0; p < dc; ++p) par ^= c.msg[(e0 + p) * P.LP + c.qe];
        return par;
    }
};

#define NMS_CAT2(a, b, c, d) a##b##c##d
#define NMS_KNAME(dc, dv) NMS_CAT2(nms_h2_kernel_, dc, _, dv)

__global__ void __launch_bounds__(512, (NMS_DCB != 0 && NMS_DCB <= 16 && NMS_DVB <= 8) ? 2 : 1)
    NMS_KNAME(NMS_DCB, NMS_DVB)(const __grid_constant__ KParams P) {
    nms_decode_body<H2Policy<NMS_DCB, NMS_DVB>>(P);
}

}   // namespace nms

#define NMS_CAT3(a, b, c, d) a##b##c##d
#define NMS_FNAME(dc, dv) NMS_CAT3(nms_h2_func_, dc, _, dv)
// exported for the launcher: the kernel's address for cudaFuncSetAttribute / occupancy / launch
extern "C" const void *NMS_FNAME(NMS_DCB, NMS_DVB)() { return (const void *)nms::NMS_KNAME(NMS_DCB, NMS_DVB); }
