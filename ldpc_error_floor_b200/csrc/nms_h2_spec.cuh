// nms_h2_spec.cuh -- graph-specialised packed kernels.  `G` is a generated struct of constexpr tables
// (csrc/gen_spec.py, one per known base graph and launch geometry).  In the iteration loop
//   * the VN phase is unrolled per column with every message row offset and circulant rotation as an
//     immediate (the rotated lane offsets are loop-invariant and end up hoisted into registers);
//   * the CN phase keeps ONE body per distinct row degree (it needs no per-edge constants, only the row's
//     base offset), which keeps the loop small enough for the instruction cache;
//   * the once-per-batch INIT pass and all cold paths use the compact table-driven code of nms_h2.cuh.
// Arithmetic is the shared code of nms_h2.cuh -- results are bit-identical to the generic kernels.
#pragma once
#include "nms_h2.cuh"

namespace nms {

template <class G, int ROT>
__device__ __forceinline__ uint32_t spec_rot(const H2Ctx &h) {
    if constexpr (ROT == 0) {
        return h.q4;
    } else if constexpr (G::L == G::LP) {
        if constexpr ((G::L & (G::L - 1)) == 0) return (h.q4 + ROT * 4u) & (G::L * 4u - 1u);
        else return h2_rot<false>(h, ROT * 4u, G::L * 4u);
    } else {
        return h2_rot<true>(h, ROT * 4u, G::L * 4u);
    }
}

template <class G>
struct H2SpecPolicy {
    static constexpr bool H2 = true;
    static constexpr uint32_t LP4 = G::LP * 4u;

    // rows of one degree share a body; the row's base offset and weights are the only per-row values
    static __device__ __forceinline__ void cn_phase(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        const H2Ctx h = h2_ctx(P, c);
        const uint32_t wcrow = h2_wrow(h, P.h2w_c, t, P.h2_wc), wurow = h2_wrow(h, P.h2w_u, t, P.h2_wu);
#pragma unroll 1
        for (int n = c.slot; n < G::M; n += G::R) {
            const int i = P.cn_order[n];
            const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
            const uint32_t a0 = h.sb + (uint32_t)e0 * LP4 + h.q4;
            const float w0 = h2_w(wcrow, i, P.h2_mc), w1 = h2_w(wurow, i, P.h2_mu);
            static_for<0, G::NDEG>([&](auto k) {
                constexpr int DC = G::cn_degs[decltype(k)::v];
                if (dc == DC) cn_row_h2<DC>(P, a0, LP4, w0, w1, bad);
            });
        }
    }

    // one column, everything constant-folded.  Only used for iterations that are followed by another one and
    // need no cold-path work (ballots / APP output): no branches, no calls.  VNW: VN weights present.
    template <int J, bool VNW>
    static __device__ __forceinline__ void vn_col(const KParams &P, const H2Ctx &h, uint32_t wvrow, uint32_t &ones) {
        constexpr int C0 = G::col_ptr[J], DV = G::col_ptr[J + 1] - C0;
        uint32_t addr[DV], cv[DV];
        static_for<0, DV>([&](auto u) {
            constexpr int U = decltype(u)::v;
            constexpr uint32_t X4 = (uint32_t)G::vn_e[C0 + U] * LP4;
            constexpr int ROT = G::vn_rot[C0 + U];
            addr[U] = h.sb + spec_rot<G, ROT>(h) + X4;
            cv[U] = lds32(addr[U]);
        });
        __half2 S = __float2half2_rn(0.0f);
#pragma unroll
        for (int u = 0; u < DV; ++u) S = __hadd2(S, u2h(cv[u]));
        const __half2 xqh = u2h(lds32(h.xq4 + (uint32_t)(J * G::LP) * 4u));
        const __half2 app = __hadd2(xqh, S);
        __half2 xin = xqh;
        if constexpr (VNW) {
            const float2 x = lds64f(h.xa8 + (uint32_t)(J * G::LP) * 8u);
            const float w = h2_w(wvrow, J, P.h2_mv);
            xin = q2(P, __fmul_rn(x.x, w), __fmul_rn(x.y, w));   // Q(xa * w), :168-177
        }
        const uint32_t hbw = (~h2u(app) >> 15) & LSB2;
        ones |= hbw;
        const __half2 SX = __hadd2(xin, S);
#pragma unroll
        for (int u = 0; u < DV; ++u) sts32(addr[u], h2u(__hsub2(SX, u2h(cv[u]))) | hbw);
    }

    template <int SLOT, bool VNW>
    static __device__ __forceinline__ void vn_slot(const KParams &P, const H2Ctx &h, uint32_t wvrow, uint32_t &ones) {
        constexpr int NT = (G::N - SLOT + G::R - 1) / G::R;
        static_for<0, NT>([&](auto n) {
            constexpr int J = G::vn_order[SLOT + decltype(n)::v * G::R];
            vn_col<J, VNW>(P, h, wvrow, ones);
        });
    }

    template <bool INIT>
    static __device__ __forceinline__ void vn_phase(const KParams &P, const Ctx &c, int t, bool need_hb, uint32_t &ones) {
        const H2Ctx h = h2_ctx(P, c);
        const int cold = h2_cold_mask(P, INIT, need_hb);
        if (INIT || cold != 0 || t + 1 >= P.T_run) {
            // init pass, last iteration, ballots or APP output wanted: the compact table-driven code
            h2_vn_phase_tab<0, INIT>(P, c, h, t, need_hb, ones);
        } else {
            const uint32_t wvrow = h2_wrow(h, P.h2w_v, t + 1, P.h2_wv);
            if (P.sharing2 != 0) {
                static_for<0, G::R>([&](auto s) {
                    if (c.slot == decltype(s)::v) vn_slot<decltype(s)::v, true>(P, h, wvrow, ones);
                });
            } else {
                static_for<0, G::R>([&](auto s) {
                    if (c.slot == decltype(s)::v) vn_slot<decltype(s)::v, false>(P, h, wvrow, ones);
                });
            }
        }
    }

    static __device__ __forceinline__ uint32_t synd_phase(const KParams &P, const Ctx &c, int tl) {
        return h2_synd_phase(P, c);
    }
};

}   // namespace nms
