// nms_f32.cuh -- float32 arithmetic (one frame per 32-bit word) shared by the degree-bucketed generic kernels
// (nms_f32.cu) and the graph-specialised ones (nms_f32_spec.cuh).  Operation order follows the TF graph: direct
// extrinsic V->C sums in ascending E(C) order (Main_Functions.py:213-215), the 1e-4 zero / minimum rules (:230,
// :250), |.|*w -> ReLU -> clip/quantise -> sign (:267-316), APP = clip(xq + sum) (:317-325).  Hard decisions are
// kept as ballot words per (column, lane chunk) so the CN phase can form the syndrome of the previous iteration.
#pragma once
#include "nms_device.cuh"

namespace nms {

// previous hard decision of the variable behind E(C) edge e, seen from check lane q
__device__ __forceinline__ uint32_t f32_hbit(const KParams &P, const Ctx &c, int buf, int e) {
    const int jv = P.e_col[e];
    int qv = c.q + P.e_sF[e] * c.act;
    qv = (qv >= c.Lthr) ? qv - P.L : qv;
    return (nms_smem[P.off_hb + (buf * P.N + jv) * P.C + (qv >> 5)] >> (qv & 31)) & 1u;
}

__device__ __forceinline__ float f32_cn_emit(const KParams &P, float raw, float m1, float m2, int npos, float w) {
    float m = fabsf(raw) > m1 ? m1 : m2;                            // min over the other edges (:248-249)
    m = (fabsf(m) > 0.0001f) ? m : __fadd_rn(m, -0.0001f);          // :250
    const int np = npos - (raw > 0.0f ? 1 : 0);
    const float x0 = (np & 1) ? m : -m;                             // :251-254
    const float x1 = __fmul_rn(fabsf(x0), w);                       // :267-298
    float x2 = x1 > 0.0f ? x1 : 0.0f;                               // :308
    x2 = P.qms ? qf(P, x2) : fminf(fmaxf(x2, -P.clip), P.clip);     // :310-313
    return x0 > 0.0f ? x2 : (x0 < 0.0f ? -x2 : 0.0f);               // :316
}

__device__ __forceinline__ float f32_edge_w(const KParams &P, bool ucn, int t, int i, int e) {
    if (P.sharing0 == 0) return 1.0f;
    return ucn ? ucn_weight(P, t, i, e) : cn_weight(P, t, i, e);
}

template <int DC>
__device__ __forceinline__ void cn_row_f32(const KParams &P, const Ctx &c, int i, int t, uint32_t &bad) {
    const int e0 = P.row_ptr[i], LP = P.LP, off = e0 * LP + c.q;
    float raw[DC];
#pragma unroll
    for (int p = 0; p < DC; ++p) raw[p] = smem_f(off + p * LP);
    uint32_t par = 0;
    const int buf = (t + 1) & 1;   // hard bits of APP_{t-1} (the init pass wrote buffer 1)
#pragma unroll
    for (int p = 0; p < DC; ++p) par ^= f32_hbit(P, c, buf, e0 + p);
    bad |= par;
    float m1 = 10000.0f, m2 = 10000.0f;   // all-masked row -> 10000 (:248)
    int npos = 0;
#pragma unroll
    for (int p = 0; p < DC; ++p) {
        const float a = fabsf(raw[p]);
        const float tmx = fmaxf(m1, a);
        m1 = fminf(m1, a);
        m2 = fminf(m2, tmx);
        npos += raw[p] > 0.0f ? 1 : 0;
    }
    const bool ucn = P.sharing1 != 0 && par;
#pragma unroll
    for (int p = 0; p < DC; ++p)
        smem_f(off + p * LP) = f32_cn_emit(P, raw[p], m1, m2, npos, f32_edge_w(P, ucn, t, i, e0 + p));
}

static __device__ __noinline__ void cn_row_f32_generic(const KParams &P, const Ctx &c, int i, int t, uint32_t &bad) {
    const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0, LP = P.LP, off = e0 * LP + c.q;
    uint32_t par = 0;
    const int buf = (t + 1) & 1;
    float m1 = 10000.0f, m2 = 10000.0f;
    int npos = 0;
    for (int p = 0; p < dc; ++p) {
        const float r = smem_f(off + p * LP);
        par ^= f32_hbit(P, c, buf, e0 + p);
        const float a = fabsf(r);
        const float tmx = fmaxf(m1, a);
        m1 = fminf(m1, a);
        m2 = fminf(m2, tmx);
        npos += r > 0.0f ? 1 : 0;
    }
    bad |= par;
    const bool ucn = P.sharing1 != 0 && par;
    for (int p = 0; p < dc; ++p)
        smem_f(off + p * LP) = f32_cn_emit(P, smem_f(off + p * LP), m1, m2, npos, f32_edge_w(P, ucn, t, i, e0 + p));
}

__device__ __forceinline__ float f32_sat(const KParams &P, float v) {
    v = P.qms ? qf(P, v) : fminf(fmaxf(v, -P.clip), P.clip);        // :223-226
    return v == 0.0f ? 0.0001f : v;                                  // :230
}

struct F32Var {
    float xin;
    bool has_next;
};

template <bool INIT>
__device__ __forceinline__ F32Var f32_var(const KParams &P, const Ctx &c, int j, int t, float S, uint32_t &ones) {
    F32Var v;
    const int slotw = j * P.LP + c.q;
    const float xa = smem_f(P.off_xa + slotw);
    float xqv = xa;
    if (P.qms) {
        if (INIT) {
            xqv = qf(P, xa);                                         // :321-322
            smem_f(P.off_xq + slotw) = xqv;
        } else {
            xqv = smem_f(P.off_xq + slotw);
        }
    }
    const float app = fminf(fmaxf(__fadd_rn(xqv, S), -P.clip), P.clip);   // :324-325
    const int tn = INIT ? 0 : t + 1;
    v.has_next = tn < P.T_run;
    v.xin = xa;
    if (v.has_next) {
        if (P.sharing2 != 0) v.xin = __fmul_rn(xa, vn_weight(P, tn, j));  // :168-169
        if (P.qms) v.xin = qf(P, v.xin);                                   // :176-177
    }
    const float hsrc = INIT ? v.xin : app;
    const bool hbit = hsrc >= 0.0f;                                  // Print_Functions.py:106
    if (!INIT && hbit && j < P.target_n) ones |= 1u;
    const uint32_t b = __ballot_sync(0xffffffffu, c.act && hbit);
    if (c.lane == 0) nms_smem[P.off_hb + ((INIT ? 1 : (t & 1)) * P.N + j) * P.C + c.chunk] = b;
    if (!INIT && P.app != nullptr) app_store(P, c, j, t, c.f0, app);
    return v;
}

template <int DV, bool INIT>
__device__ __forceinline__ void vn_col_f32(const KParams &P, const Ctx &c, int j, int t, uint32_t &ones) {
    const int c0 = P.col_ptr[j], L = P.L;
    int addr[DV];
    float cv[DV];
#pragma unroll
    for (int u = 0; u < DV; ++u) {
        addr[u] = vn_addr(c, P.vn_edge[c0 + u], L);
        cv[u] = INIT ? 0.0f : smem_f(addr[u]);
    }
    float S = 0.0f;
#pragma unroll
    for (int u = 0; u < DV; ++u) S = __fadd_rn(S, cv[u]);            // ascending E(C), like the GEMM column (:317)
    const F32Var v = f32_var<INIT>(P, c, j, t, S, ones);
    if (v.has_next) {
#pragma unroll
        for (int u = 0; u < DV; ++u) {
            float acc = 0.0f;                                        // direct extrinsic sum (:214), ascending
#pragma unroll
            for (int u2 = 0; u2 < DV; ++u2)
                if (u2 != u) acc = __fadd_rn(acc, cv[u2]);
            smem_f(addr[u]) = f32_sat(P, __fadd_rn(v.xin, acc));     // :215, :223-230
        }
    }
}

// any column degree up to 64 (host-guarded): same arithmetic, C->V staged in a local array
template <bool INIT>
__device__ __noinline__ void vn_col_f32_generic(const KParams &P, const Ctx &c, int j, int t, uint32_t &ones) {
    const int c0 = P.col_ptr[j], dv = min(P.col_ptr[j + 1] - c0, 64), L = P.L;
    float cv[64];
    for (int u = 0; u < dv; ++u) cv[u] = INIT ? 0.0f : smem_f(vn_addr(c, P.vn_edge[c0 + u], L));
    float S = 0.0f;
    for (int u = 0; u < dv; ++u) S = __fadd_rn(S, cv[u]);
    const F32Var v = f32_var<INIT>(P, c, j, t, S, ones);
    if (v.has_next) {
        for (int u = 0; u < dv; ++u) {
            float acc = 0.0f;
            for (int u2 = 0; u2 < dv; ++u2)
                if (u2 != u) acc = __fadd_rn(acc, cv[u2]);
            smem_f(vn_addr(c, P.vn_edge[c0 + u], L)) = f32_sat(P, __fadd_rn(v.xin, acc));
        }
    }
}

}   // namespace nms
