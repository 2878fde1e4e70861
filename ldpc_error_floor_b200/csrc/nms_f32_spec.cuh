// nms_f32_spec.cuh -- graph-specialised float32 kernels (one frame per 32-bit lane): the float min-sum path
// (decoding_type 1) and the quantised modes the packed kernels do not take (q_bit 6, per-edge weights), for the base
// graphs known at build time.  Same structure as nms_h2_spec.cuh: every row / column offset, degree and circulant
// rotation is an immediate, the VN phase is unrolled per column, the prologue takes the channel LLRs straight from
// global memory.  The arithmetic is the reference-ordered code of nms_f32.cu (direct extrinsic sums in ascending
// E(C) order, the 1e-4 rules, |.|*w -> ReLU -> saturate -> sign), so results equal the generic float kernels bit for
// bit.  Hard decisions travel as ballot words per (column, lane chunk); the previous syndrome a check needs for its
// unsatisfied-check weight is the parity of those bits at the rotated lanes.
#pragma once
#include "nms_f32.cuh"

namespace nms {

template <class G>
struct F32SpecPolicy {
    static constexpr bool H2 = false;
    static constexpr bool FUSED_LOAD = true;
    static constexpr bool PAD = G::L != G::LP;

    // lane q -> (q + ROT) mod L in words; padding lanes stay put
    template <int ROT>
    static __device__ __forceinline__ int rot(const Ctx &c) {
        if constexpr (ROT == 0) return c.q;
        int qq = c.q + (PAD ? ROT * c.act : ROT);
        return qq >= (PAD ? c.Lthr : G::L) ? qq - G::L : qq;
    }

    // ------------------------------------------------------------------------------ CN phase
    static __device__ __forceinline__ float sat(const KParams &P, float x) {   // Q() or clip, no mode branch
        return fminf(fmaxf(__fsub_rn(__fadd_rn(x, P.sat_magic), P.sat_magic), -P.sat_bound), P.sat_bound);
    }
    // weighted, saturated output magnitude for a minimum m of the other edges (the arithmetic of f32_cn_emit, done once
    // per row for min1 and min2 instead of once per edge): returns the bits of the value with sign(m_fixed) folded in
    static __device__ __forceinline__ uint32_t row_mag(const KParams &P, float m, float w) {
        m = m > 0.0001f ? m : __fadd_rn(m, -0.0001f);                 // :250
        const float x1 = __fmul_rn(fabsf(m), w);                      // :267-298
        const float x2 = sat(P, x1 > 0.0f ? x1 : 0.0f);               // :308-313
        return __float_as_uint(x2) ^ (__float_as_uint(m) & 0x80000000u);
    }

    template <int I>
    static __device__ __forceinline__ void cn_row(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        constexpr int E0 = G::row_ptr[I], DC = G::row_ptr[I + 1] - E0;
        const int off = E0 * G::LP + c.q;
        float raw[DC];
#pragma unroll
        for (int p = 0; p < DC; ++p) raw[p] = smem_f(off + p * G::LP);
        // syndrome of the previous hard decision: its bits sit in hb[buf][column][chunk] at the variable's lane
        uint32_t par = 0;
        const uint32_t *hb = nms_smem + P.off_hb + ((t + 1) & 1) * G::N * G::C;
        static_for<0, DC>([&](auto p) {
            constexpr int E = E0 + decltype(p)::v;
            const int qv = rot<G::e_sF[E]>(c);
            par ^= hb[G::e_col[E] * G::C + (qv >> 5)] >> (qv & 31);
        });
        par &= 1u;
        bad |= par;
        float m1 = 10000.0f, m2 = 10000.0f;   // all-masked row -> 10000 (:248)
        uint32_t sx = 0;                       // XOR of the inputs: bit 31 = parity of the negative ones
#pragma unroll
        for (int p = 0; p < DC; ++p) {
            const float a = fabsf(raw[p]);
            const float tmx = fmaxf(m1, a);
            m1 = fminf(m1, a);
            m2 = fminf(m2, tmx);
            sx ^= __float_as_uint(raw[p]);
        }
        // C->V of edge p: magnitude from the minimum of the OTHER edges, negative iff (number of positive others) is even
        // (:251-254; an input is never 0 here, :230).  With P = parity of all positive inputs and s = sign bit of the edge's
        // own input: sign bit of the output = sign(m_fixed) ^ P ^ s.
        const uint32_t Pbit = ((sx >> 31) ^ (uint32_t)(DC & 1)) << 31;
        const bool ucn = P.sharing1 != 0 && par;
        if (P.sharing0 == 1) {     // per-edge weights: nothing to hoist
#pragma unroll
            for (int p = 0; p < DC; ++p) {
                const float w = ucn ? ucn_weight(P, t, I, E0 + p) : cn_weight(P, t, I, E0 + p);
                const uint32_t v = row_mag(P, fabsf(raw[p]) > m1 ? m1 : m2, w);
                smem_f(off + p * G::LP) = __uint_as_float(v ^ Pbit ^ (__float_as_uint(raw[p]) & 0x80000000u));
            }
        } else {
            const float w = f32_edge_w(P, ucn, t, I, E0);   // one weight per row (or none)
            const uint32_t A = row_mag(P, m1, w) ^ Pbit, B = row_mag(P, m2, w) ^ Pbit;
#pragma unroll
            for (int p = 0; p < DC; ++p) {
                const uint32_t v = fabsf(raw[p]) > m1 ? A : B;
                smem_f(off + p * G::LP) = __uint_as_float(v ^ (__float_as_uint(raw[p]) & 0x80000000u));
            }
        }
    }

    static __device__ __forceinline__ void cn_phase(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        static_for<0, G::R>([&](auto s) {
            constexpr int SLOT = decltype(s)::v;
            constexpr int NT = (G::M - SLOT + G::R - 1) / G::R;
            if (c.slot == SLOT) {
                static_for<0, NT>([&](auto n) { cn_row<G::cn_order[SLOT + decltype(n)::v * G::R]>(P, c, t, bad); });
            }
        });
    }

    // ------------------------------------------------------------------------------ VN phase
    // MODE 0: iteration t;  1: pass before iteration 0, xa in shared memory;  2: same, xa = `xg` from global memory
    template <int J, int MODE>
    static __device__ __forceinline__ void vn_col(const KParams &P, const Ctx &c, int t, float xg, uint32_t &ones) {
        constexpr int C0 = G::col_ptr[J], DV = G::col_ptr[J + 1] - C0;
        constexpr bool INIT = MODE != 0;
        int addr[DV];
        float cv[DV];
        static_for<0, DV>([&](auto u) {
            constexpr int U = decltype(u)::v;
            addr[U] = G::vn_e[C0 + U] * G::LP + rot<G::vn_rot[C0 + U]>(c);
            cv[U] = INIT ? 0.0f : smem_f(addr[U]);
        });
        float S = 0.0f;
#pragma unroll
        for (int u = 0; u < DV; ++u) S = __fadd_rn(S, cv[u]);            // ascending E(C), like the GEMM column (:317)
        if constexpr (MODE == 2) {
            if (P.qms) xg = fminf(fmaxf(xg, -XA_BOUND), XA_BOUND);
            smem_f(P.off_xa + J * G::LP + c.q) = xg;
        }
        const F32Var v = f32_var<INIT>(P, c, J, t, S, ones);
        if (v.has_next) {
#pragma unroll
            for (int u = 0; u < DV; ++u) {
                float acc = 0.0f;                                        // direct extrinsic sum (:214), ascending
#pragma unroll
                for (int u2 = 0; u2 < DV; ++u2)
                    if (u2 != u) acc = __fadd_rn(acc, cv[u2]);
                const float m = sat(P, __fadd_rn(v.xin, acc));           // :215, :223-226
                smem_f(addr[u]) = m == 0.0f ? 0.0001f : m;               // :230
            }
        }
    }

    template <int SLOT, int MODE>
    static __device__ __forceinline__ void vn_slot(const KParams &P, const Ctx &c, int t, uint32_t &ones) {
        constexpr int NT = (G::N - SLOT + G::R - 1) / G::R;
        if constexpr (MODE == 2) {
            const bool ok = c.act && c.f0 < c.nvalid;
            const long long o = (c.frame0 + (ok ? c.f0 : 0)) * (long long)P.NZ + c.a_lane;
            float x[NT];
            if (P.llr != nullptr) {
                const float *p0 = P.llr + o;
                static_for<0, NT>([&](auto n) {
                    constexpr int J = G::vn_order[SLOT + decltype(n)::v * G::R];
                    x[decltype(n)::v] = ok ? __ldg(p0 + J * G::z) : 0.0f;
                });
            } else {
                const signed char *p0 = P.llr_q8 + o;
                static_for<0, NT>([&](auto n) {
                    constexpr int J = G::vn_order[SLOT + decltype(n)::v * G::R];
                    x[decltype(n)::v] = ok ? (float)__ldg(p0 + J * G::z) * P.q8_step : 0.0f;
                });
            }
            static_for<0, NT>([&](auto n) {
                vn_col<G::vn_order[SLOT + decltype(n)::v * G::R], MODE>(P, c, t, x[decltype(n)::v], ones);
            });
        } else {
            static_for<0, NT>([&](auto n) { vn_col<G::vn_order[SLOT + decltype(n)::v * G::R], MODE>(P, c, t, 0.0f, ones); });
        }
    }

    template <int MODE>
    static __device__ __forceinline__ void vn_dispatch(const KParams &P, const Ctx &c, int t, uint32_t &ones) {
        static_for<0, G::R>([&](auto s) {
            if (c.slot == decltype(s)::v) vn_slot<decltype(s)::v, MODE>(P, c, t, ones);
        });
    }

    template <bool INIT>
    static __device__ __forceinline__ void vn_phase(const KParams &P, const Ctx &c, int t, bool need_hb, uint32_t &ones) {
        if constexpr (INIT) vn_dispatch<1>(P, c, t, ones);
        else vn_dispatch<0>(P, c, t, ones);
    }

    static __device__ __forceinline__ void load_init(const KParams &P, const Ctx &c) {
        uint32_t dummy = 0;
        vn_dispatch<2>(P, c, -1, dummy);
    }

    static __device__ __forceinline__ uint32_t synd_phase(const KParams &P, const Ctx &c, int tl) {
        uint32_t bad = 0;
        const uint32_t *hb = nms_smem + P.off_hb + ((tl + 1) & 1) * G::N * G::C;
        static_for<0, G::R>([&](auto s) {
            constexpr int SLOT = decltype(s)::v;
            constexpr int NT = (G::M - SLOT + G::R - 1) / G::R;
            if (c.slot == SLOT) {
                static_for<0, NT>([&](auto n) {
                    constexpr int I = G::cn_order[SLOT + decltype(n)::v * G::R];
                    constexpr int E0 = G::row_ptr[I], DC = G::row_ptr[I + 1] - E0;
                    uint32_t par = 0;
                    static_for<0, DC>([&](auto p) {
                        constexpr int E = E0 + decltype(p)::v;
                        const int qv = rot<G::e_sF[E]>(c);
                        par ^= hb[G::e_col[E] * G::C + (qv >> 5)] >> (qv & 31);
                    });
                    bad |= par & 1u;
                });
            }
        });
        return bad;
    }
};

}   // namespace nms
