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
#pragma once
#include "nms_device.cuh"

namespace nms {

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

__device__ __forceinline__ float h2_wcn(const KParams &P, int t, int i) { return smem_f(P.h2w_c + t * P.h2_wc + (i & P.h2_mc)); }
__device__ __forceinline__ float h2_wucn(const KParams &P, int t, int i) { return smem_f(P.h2w_u + t * P.h2_wu + (i & P.h2_mu)); }
__device__ __forceinline__ float h2_wvn(const KParams &P, int t, int j) { return smem_f(P.h2w_v + t * P.h2_wv + (j & P.h2_mv)); }

// Q() of two float32 values -> half2 (round to step in fp32, saturate after packing; +-inf saturate too)
__device__ __forceinline__ __half2 q2(const KParams &P, float lo, float hi) {
    const __half2 qm = __float2half2_rn(P.qmax);
    const __half2 r = __floats2half2_rn(qround(lo, P.qmagic), qround(hi, P.qmagic));
    return __hmax2(__hmin2(r, qm), __hneg2(qm));
}

// weighted, quantised magnitudes for "edge is not the minimum" (A) and "edge is the minimum" (B), with the
// row's sign parity folded in.  par: XOR of all raw V->C words of the row; w0 / w1: CN / UCN weight.
__device__ __forceinline__ void h2_row_mags(const KParams &P, float w0, float w1, bool dc_odd, uint32_t par,
                                            __half2 m1, __half2 m2, uint32_t &A, uint32_t &B) {
    const __half2 qm = __float2half2_rn(P.qmax), zero = __float2half2_rn(0.0f);
    // drop the piggy-backed hard bits; V->C saturation (:223-224) applied to the two minima
    const __half2 m1c = __hmin2(u2h(h2u(m1) & ~LSB2), qm), m2c = __hmin2(u2h(h2u(m2) & ~LSB2), qm);
    const float wlo = (par & 1u) ? w1 : w0;         // unsatisfied check -> UCN weight (:275,:285,:295)
    const float whi = (par & 0x10000u) ? w1 : w0;
    // Q(relu(min * w)) (:308-311)
    __half2 magA = __floats2half2_rn(qround(__fmul_rn(__low2float(m1c), wlo), P.qmagic),
                                     qround(__fmul_rn(__high2float(m1c), whi), P.qmagic));
    __half2 magB = __floats2half2_rn(qround(__fmul_rn(__low2float(m2c), wlo), P.qmagic),
                                     qround(__fmul_rn(__high2float(m2c), whi), P.qmagic));
    magA = __hmax2(__hmin2(magA, qm), zero);
    magB = __hmax2(__hmin2(magB, qm), zero);
    // C->V is negative iff (dc + #negative others) is odd (:251-254; a zero V->C counts as positive, :230)
    const uint32_t s0 = (par & SIGN2) ^ (dc_odd ? SIGN2 : 0u);
    A = h2u(magA) ^ s0;
    B = h2u(magB) ^ s0;
}

// one check row held in registers.  `off`: word index of msg[e0][q]; `stride`: words between edges (LP)
template <int DC>
__device__ __forceinline__ void cn_row_h2(const KParams &P, int off, int stride, float w0, float w1, uint32_t &bad) {
    uint32_t raw[DC];
#pragma unroll
    for (int p = 0; p < DC; ++p) raw[p] = nms_smem[off + p * stride];
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
    h2_row_mags(P, w0, w1, (DC & 1) != 0, par, m1, m2, A, B);
#pragma unroll
    for (int p = 0; p < DC; ++p) {
        const uint32_t gt = __hgt2_mask(__habs2(u2h(raw[p])), m1);   // |v| > min1 -> others' min is min1, else min2
        nms_smem[off + p * stride] = ((gt & A) | (~gt & B)) ^ (raw[p] & SIGN2);
    }
}

// any degree: two passes over shared memory instead of a register array
static __device__ __noinline__ void cn_row_h2_generic(const KParams &P, int off, int stride, int dc, float w0, float w1,
                                                      uint32_t &bad) {
    uint32_t par = 0;
    __half2 m1 = __float2half2_rn(10000.0f), m2 = m1;
    for (int p = 0; p < dc; ++p) {
        const uint32_t r = nms_smem[off + p * stride];
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
        const uint32_t r = nms_smem[off + p * stride];
        const uint32_t gt = __hgt2_mask(__habs2(u2h(r)), m1);
        nms_smem[off + p * stride] = ((gt & A) | (~gt & B)) ^ (r & SIGN2);
    }
}

// per-variable part of the VN phase
struct H2Var {
    __half2 xin;      // next iteration's weighted + quantised channel value
    uint32_t hbw;     // hard bits (bit 0 / bit 16) of this slot's two frames
    bool has_next;
};

// slotw: word index j*LP + q.  wv_next: VN weight of the next iteration (ignored when sharing2 == 0)
template <bool INIT>
__device__ __forceinline__ H2Var h2_var(const KParams &P, const Ctx &c, int j, int t, int slotw, __half2 S,
                                        bool need_hb, uint32_t &ones) {
    H2Var v;
    const float2 x = *reinterpret_cast<const float2 *>(&smem_f(P.off_xa + 2 * slotw));
    __half2 xqh;
    if (INIT) {
        xqh = q2(P, x.x, x.y);                                   // Q(xa), :321-322
        nms_smem[P.off_xq + slotw] = h2u(xqh);
    } else {
        xqh = u2h(nms_smem[P.off_xq + slotw]);
    }
    const __half2 app = __hadd2(xqh, S);   // unclipped APP; clip_LLR never changes its sign
    const int tn = INIT ? 0 : t + 1;
    v.has_next = tn < P.T_run;
    v.xin = xqh;
    if (v.has_next && P.sharing2 != 0) {
        const float w = h2_wvn(P, tn, j);
        v.xin = q2(P, __fmul_rn(x.x, w), __fmul_rn(x.y, w));     // Q(xa * w), :168-177
    }
    const __half2 hsrc = INIT ? v.xin : app;   // iteration 0 takes the syndrome of xin_0 (:181-182)
    v.hbw = (~h2u(hsrc) >> 15) & LSB2;         // bit = (value >= 0); a zero here is always +0
    if (!INIT) ones |= v.hbw;
    if (need_hb) {
        const uint32_t lo = __ballot_sync(0xffffffffu, c.act && (v.hbw & 1u));
        const uint32_t hi = __ballot_sync(0xffffffffu, c.act && (v.hbw >> 16));
        if (c.lane == 0) {
            const int buf = INIT ? 1 : (t & 1);
            nms_smem[P.off_hb + ((buf * 2 + 0) * P.N + j) * P.C + c.chunk] = lo;
            nms_smem[P.off_hb + ((buf * 2 + 1) * P.N + j) * P.C + c.chunk] = hi;
        }
    }
    if (!INIT && P.app != nullptr) {   // optional float APP output (ya_output{t}, :324-327)
        app_store(P, c, j, t, c.f0, fminf(fmaxf(__low2float(app), -P.clip), P.clip));
        app_store(P, c, j, t, c.f1, fminf(fmaxf(__high2float(app), -P.clip), P.clip));
    }
    return v;
}

// INIT: the pass before iteration 0 (C->V = 0): writes xq, V->C = Q(xa*wv_0), hard bit of xin_0.
template <int DV, bool INIT>
__device__ __forceinline__ void vn_col_h2(const KParams &P, const Ctx &c, int j, int t, bool need_hb, uint32_t &ones) {
    const int c0 = P.col_ptr[j], L = P.L;
    int addr[DV];
    uint32_t cv[DV];
#pragma unroll
    for (int u = 0; u < DV; ++u) {
        addr[u] = vn_addr(c, P.vn_edge[c0 + u], L);
        cv[u] = INIT ? 0u : nms_smem[addr[u]];
    }
    __half2 S = __float2half2_rn(0.0f);
    if (!INIT) {
#pragma unroll
        for (int u = 0; u < DV; ++u) S = __hadd2(S, u2h(cv[u]));
    }
    const H2Var v = h2_var<INIT>(P, c, j, t, j * P.LP + c.q, S, need_hb, ones);
    if (v.has_next) {
        const __half2 SX = __hadd2(v.xin, S);
#pragma unroll
        for (int u = 0; u < DV; ++u) {
            const __half2 m = INIT ? v.xin : __hsub2(SX, u2h(cv[u]));   // total - self: exact on the grid (:213-215)
            nms_smem[addr[u]] = h2u(m) | v.hbw;
        }
    } else {
#pragma unroll
        for (int u = 0; u < DV; ++u) nms_smem[addr[u]] = v.hbw;         // only the final syndrome pass reads these
    }
}

template <bool INIT>
__device__ __noinline__ void vn_col_h2_generic(const KParams &P, const Ctx &c, int j, int t, bool need_hb,
                                               uint32_t &ones) {
    const int c0 = P.col_ptr[j], dv = P.col_ptr[j + 1] - c0, L = P.L;
    __half2 S = __float2half2_rn(0.0f);
    if (!INIT)
        for (int u = 0; u < dv; ++u) S = __hadd2(S, u2h(nms_smem[vn_addr(c, P.vn_edge[c0 + u], L)]));
    const H2Var v = h2_var<INIT>(P, c, j, t, j * P.LP + c.q, S, need_hb, ones);
    const __half2 SX = __hadd2(v.xin, S);
    for (int u = 0; u < dv; ++u) {
        const int a = vn_addr(c, P.vn_edge[c0 + u], L);
        uint32_t o = v.hbw;
        if (v.has_next) o |= h2u(INIT ? v.xin : __hsub2(SX, u2h(nms_smem[a])));
        nms_smem[a] = o;
    }
}

}   // namespace nms
