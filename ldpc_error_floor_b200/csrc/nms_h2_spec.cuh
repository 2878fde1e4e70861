// nms_h2_spec.cuh -- graph-specialised packed kernels.  `G` is a generated struct of constexpr tables
// (csrc/gen_spec.py, one per known base graph and launch geometry): every table lookup, stride and
// rotation becomes an immediate, every row / column is unrolled with its exact degree, and the per-slot
// task lists are resolved at compile time, so the instruction stream is almost only message arithmetic.
// Arithmetic is the shared code of nms_h2.cuh -- results are bit-identical to the generic kernels.
#pragma once
#include "nms_h2.cuh"

namespace nms {

// variable lane q -> check lane (q + ROT) mod L for one edge, ROT a compile-time constant
template <class G, int ROT>
__device__ __forceinline__ int spec_rot(const Ctx &c) {
    if constexpr (ROT == 0) {
        return c.q;
    } else if constexpr (G::L == G::LP) {
        if constexpr ((G::L & (G::L - 1)) == 0) {
            return (c.q + ROT) & (G::L - 1);
        } else {
            const unsigned t1 = (unsigned)(c.q + ROT), t2 = (unsigned)(c.q + (ROT - G::L));   // t2 wraps high if no wrap
            return (int)min(t1, t2);
        }
    } else {
        int qq = c.q + ROT * c.act;      // padding lanes keep their own padding word
        return (qq >= c.Lthr) ? qq - G::L : qq;
    }
}

template <class G>
struct H2SpecPolicy {
    static constexpr bool H2 = true;

    template <int SLOT>
    static __device__ __forceinline__ void cn_slot(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        constexpr int NT = (G::M - SLOT + G::R - 1) / G::R;
        static_for<0, NT>([&](auto n) {
            constexpr int I = G::cn_order[SLOT + decltype(n)::v * G::R];
            constexpr int E0 = G::row_ptr[I], DC = G::row_ptr[I + 1] - E0;
            const float w0 = h2_wcn(P, t, I), w1 = h2_wucn(P, t, I);
            cn_row_h2<DC>(P, E0 * G::LP + c.q, G::LP, w0, w1, bad);
        });
    }

    static __device__ __forceinline__ void cn_phase(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        static_for<0, G::R>([&](auto s) {
            if (c.slot == decltype(s)::v) cn_slot<decltype(s)::v>(P, c, t, bad);
        });
    }

    template <int J, bool INIT>
    static __device__ __forceinline__ void vn_col(const KParams &P, const Ctx &c, int t, bool need_hb, uint32_t &ones) {
        constexpr int C0 = G::col_ptr[J], DV = G::col_ptr[J + 1] - C0;
        int addr[DV];
        uint32_t cv[DV];
        static_for<0, DV>([&](auto u) {
            constexpr int U = decltype(u)::v;
            constexpr int X = G::vn_e[C0 + U] * G::LP, ROT = G::vn_rot[C0 + U];
            addr[U] = X + spec_rot<G, ROT>(c);
            cv[U] = INIT ? 0u : nms_smem[addr[U]];
        });
        __half2 S = __float2half2_rn(0.0f);
        if (!INIT) {
#pragma unroll
            for (int u = 0; u < DV; ++u) S = __hadd2(S, u2h(cv[u]));
        }
        const H2Var v = h2_var<INIT>(P, c, J, t, J * G::LP + c.q, S, need_hb, ones);
        if (v.has_next) {
            const __half2 SX = __hadd2(v.xin, S);
#pragma unroll
            for (int u = 0; u < DV; ++u) {
                const __half2 m = INIT ? v.xin : __hsub2(SX, u2h(cv[u]));
                nms_smem[addr[u]] = h2u(m) | v.hbw;
            }
        } else {
#pragma unroll
            for (int u = 0; u < DV; ++u) nms_smem[addr[u]] = v.hbw;
        }
    }

    template <int SLOT, bool INIT>
    static __device__ __forceinline__ void vn_slot(const KParams &P, const Ctx &c, int t, bool need_hb, uint32_t &ones) {
        constexpr int NT = (G::N - SLOT + G::R - 1) / G::R;
        static_for<0, NT>([&](auto n) {
            constexpr int J = G::vn_order[SLOT + decltype(n)::v * G::R];
            vn_col<J, INIT>(P, c, t, need_hb, ones);
        });
    }

    template <bool INIT>
    static __device__ __forceinline__ void vn_phase(const KParams &P, const Ctx &c, int t, bool need_hb, uint32_t &ones) {
        static_for<0, G::R>([&](auto s) {
            if (c.slot == decltype(s)::v) vn_slot<decltype(s)::v, INIT>(P, c, t, need_hb, ones);
        });
    }

    static __device__ __forceinline__ uint32_t synd_phase(const KParams &P, const Ctx &c, int tl) {
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
};

}   // namespace nms
