// nms_h2_spec.cuh -- graph-specialised packed kernels.  `G` is a generated struct of constexpr tables
// (csrc/gen_spec.py, one per known base graph and launch geometry).
//   * VN phase: unrolled per column with every message row offset and circulant rotation as an immediate
//     (the rotated lane offsets are loop-invariant and end up hoisted into registers).  The same unrolled
//     column code serves the iteration loop (with or without hard-decision ballots, i.e. with or without
//     early termination), the pass before iteration 0, and the fused "load channel LLRs from global memory +
//     first V->C messages" prologue -- no table-driven code is left on the path of a decode without APP output.
//   * CN phase: ONE body per distinct row degree (it needs no per-edge constants, only the row's base offset),
//     which keeps the loop small enough for the 32 KB instruction cache.
//   * final syndrome pass: unrolled per row.
// Arithmetic is the shared code of nms_h2.cuh -- results are bit-identical to the generic kernels.
#pragma once
#include "nms_h2.cuh"

namespace nms {

__device__ __forceinline__ void sts64f(uint32_t a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y));
}

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

enum { VN_ITER = 0, VN_INIT_SMEM = 1, VN_INIT_GLOBAL = 2 };

// ------------------------------------------------------------------------------------ CN phase of the specialised kernels
// Rows of one degree share a body (the loop stays small in the instruction cache); a slot's rows come sorted by degree, so
// the phase is one counted loop per degree class -- G::cn_cls_cnt[slot][class] rows -- instead of a compare-and-branch chain
// per row.  The row list {byte offset of the row's first message, degree | row << 16} is host-built (KParams::cn_task) and
// read one entry ahead.  WROW: the weights differ from row to row (sharing code 2); otherwise they are loaded once per phase.
// UCN: unsatisfied-check weights exist.  PERHALF: the two frames of a lane are at different iterations (persistent-slot
// Monte-Carlo kernel), i.e. one weight row per half; otherwise *_hi == *_lo.
template <class G, bool PERHALF, bool WROW, bool UCN>
__device__ __forceinline__ void spec_cn_rows(const KParams &P, const H2Ctx &h, int slot, uint32_t wc_lo, uint32_t wc_hi,
                                             uint32_t wu_lo, uint32_t wu_hi, uint32_t &bad) {
    constexpr uint32_t LP4 = G::LP * 4u;
    constexpr int NT = (G::M + G::R - 1) / G::R;
    float w0l = 1.0f, w0h = 1.0f, w1l = 1.0f, w1h = 1.0f;
    if constexpr (!WROW) {
        w0l = ldsf(wc_lo);
        w0h = PERHALF ? ldsf(wc_hi) : w0l;
        w1l = UCN ? ldsf(wu_lo) : w0l;
        w1h = UCN ? (PERHALF ? ldsf(wu_hi) : w1l) : w0h;
    }
    const uint2 *task = P.cn_task + slot * NT;
    uint2 tk = task[0];
    static_for<0, G::NDEG>([&](auto k) {
        constexpr int K = decltype(k)::v;
        constexpr int DC = G::cn_degs_desc[K];
        int nk = 0;                                         // rows of this class in this warp's slot (compile-time table)
        static_for<0, G::R>([&](auto sl) {
            constexpr int CNT = G::cn_cls_cnt[decltype(sl)::v * G::NDEG + K];
            if (slot == decltype(sl)::v) nk = CNT;
        });
#pragma unroll 1
        for (int n = 0; n < nk; ++n) {
            const uint2 cur = tk;
            tk = *++task;                                   // next row's entry: its latency hides behind this row
            const uint32_t a0 = h.sb + cur.x + h.q4;
            if constexpr (WROW) {
                const uint32_t i4 = (cur.y >> 14) & ~3u;    // row index * 4
                w0l = ldsf(wc_lo + (i4 & (uint32_t)P.h2_mc));
                w0h = PERHALF ? ldsf(wc_hi + (i4 & (uint32_t)P.h2_mc)) : w0l;
                w1l = UCN ? ldsf(wu_lo + (i4 & (uint32_t)P.h2_mu)) : w0l;
                w1h = UCN ? (PERHALF ? ldsf(wu_hi + (i4 & (uint32_t)P.h2_mu)) : w1l) : w0h;
            }
            cn_row_h2<DC>(P, a0, LP4, w0l, w0h, w1l, w1h, bad);
        }
    });
}

template <class G, bool PERHALF>
__device__ __forceinline__ void spec_cn_phase(const KParams &P, const H2Ctx &h, int slot, uint32_t wc_lo, uint32_t wc_hi,
                                              uint32_t wu_lo, uint32_t wu_hi, uint32_t &bad) {
    const bool wrow = (P.h2_mc | P.h2_mu) != 0, ucn = P.h2w_u != P.h2w_c;   // uniform
    if (wrow) {
        if (ucn) spec_cn_rows<G, PERHALF, true, true>(P, h, slot, wc_lo, wc_hi, wu_lo, wu_hi, bad);
        else spec_cn_rows<G, PERHALF, true, false>(P, h, slot, wc_lo, wc_hi, wu_lo, wu_hi, bad);
    } else {
        if (ucn) spec_cn_rows<G, PERHALF, false, true>(P, h, slot, wc_lo, wc_hi, wu_lo, wu_hi, bad);
        else spec_cn_rows<G, PERHALF, false, false>(P, h, slot, wc_lo, wc_hi, wu_lo, wu_hi, bad);
    }
}

template <class G>
struct H2SpecPolicy {
    static constexpr bool H2 = true;
    static constexpr bool FUSED_LOAD = true;
    static constexpr bool TRACKS_GRID = true;   // the init pass reports off-grid channel values (Ctx::og)
    static constexpr uint32_t LP4 = G::LP * 4u;
    static constexpr bool PAD = G::L != G::LP;
    static __device__ __forceinline__ void setup(const KParams &, int) {}

    // ------------------------------------------------------------------------------ CN phase
    static __device__ __forceinline__ void cn_phase(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        const H2Ctx h = h2_ctx(P, c);
        const uint32_t wcrow = h2_wrow(h, P.h2w_c, t, P.h2_wc), wurow = h2_wrow(h, P.h2w_u, t, P.h2_wu);
        spec_cn_phase<G, false>(P, h, c.slot, wcrow, wcrow, wurow, wurow, bad);
    }

    // ------------------------------------------------------------------------------ VN phase
    // One column, everything constant-folded: no branches, no calls.
    //   MODE VN_ITER        iteration t: sum the C->V words, APP sign -> hard bit, xin of iteration t+1, V->C out
    //        VN_INIT_SMEM   pass before iteration 0 (C->V = 0), channel values already in the xa array
    //        VN_INIT_GLOBAL same, channel values `xg` just loaded from global memory (also fills the xa array)
    //   VNW: VN weights present.  HB: publish the hard decisions as ballots (a copy-out may follow).
    // wvs: the VN weight when it does not vary per column (sharing code 3: loaded once per phase), NaN-free sentinel < 0
    // otherwise (sharing code 2: fetched per column from wvrow)
    //   MV:  VN weights vary per column (sharing code 2: fetched per column from wvrow) -- else `wvs` is THE weight of the
    //        iteration (code 3: loaded once per phase)
    //   OG:  this warp's channel values are on the quantiser grid and inside +-qmax (found by the init pass, which reports
    //        off-grid values through `ones`): Q(xa) is then xa itself, one pack instead of round + pack + clamp
    template <int J, int MODE, bool VNW, bool HB, bool MV, bool OG>
    static __device__ __forceinline__ void vn_col(const KParams &P, const H2Ctx &h, uint32_t wvrow, float wvs, uint32_t hbrow,
                                                  float2 xg, uint32_t &ones) {
        constexpr int C0 = G::col_ptr[J], DV = G::col_ptr[J + 1] - C0;
        constexpr bool INIT = MODE != VN_ITER;
        uint32_t addr[DV], cv[DV];
        static_for<0, DV>([&](auto u) {
            constexpr int U = decltype(u)::v;
            constexpr uint32_t X4 = (uint32_t)G::vn_e[C0 + U] * LP4;
            constexpr int ROT = G::vn_rot[C0 + U];
            addr[U] = h.sb + spec_rot<G, ROT>(h) + X4;
            if constexpr (!INIT) cv[U] = lds32(addr[U]);
        });
        __half2 S = __float2half2_rn(0.0f);
        if constexpr (!INIT) S = h2_tree_sum<DV, 0, DV>(cv);
        constexpr uint32_t XA8 = (uint32_t)(J * G::LP) * 8u, XQ4 = (uint32_t)(J * G::LP) * 4u;
        float2 x = xg;
        __half2 xqh;
        if constexpr (MODE == VN_INIT_GLOBAL) {
            x.x = fminf(fmaxf(x.x, -XA_BOUND), XA_BOUND);
            x.y = fminf(fmaxf(x.y, -XA_BOUND), XA_BOUND);
            sts64f(h.xa8 + XA8, x);
        }
        if constexpr (MODE == VN_INIT_SMEM || (MODE == VN_ITER && VNW)) x = lds64f(h.xa8 + XA8);
        // With VN weights xa is loaded anyway, so Q(xa) is recomputed (6 instructions) instead of kept in an array of its
        // own: 9 KB less shared memory per CTA on WiMAX, i.e. a fourth resident CTA (KParams::no_xq, set by the host).
        if constexpr (!INIT && VNW && OG) {
            xqh = __floats2half2_rn(x.x, x.y);           // Q(xa) = xa: exact
        } else if constexpr (INIT || VNW) {
            xqh = q2(P, x.x, x.y);                       // Q(xa), :321-322
            if constexpr (!VNW) sts32(h.xq4 + XQ4, h2u(xqh));
            if constexpr (INIT && VNW) {                 // any value off the grid or outside +-qmax?  (NaN: yes)
                const float2 back = __half22float2(xqh);   // compared as bits: -0.0 (Q gives +0.0) counts as off the grid
                ones |= (__float_as_uint(back.x) ^ __float_as_uint(x.x)) | (__float_as_uint(back.y) ^ __float_as_uint(x.y));
            }
        } else {
            xqh = u2h(lds32(h.xq4 + XQ4));
        }
        __half2 xin = xqh;
        if constexpr (VNW) {
            const float w = MV ? h2_w(wvrow, J, -1) : wvs;
            const float2 xw = mul2_rn_unfused(x, make_float2(w, w));
            xin = q2(P, xw.x, xw.y);                             // Q(xa * w), :168-177
        }
        // hard bit = (value >= 0): of xin_0 before iteration 0 (:181-182), of the APP afterwards (a zero is +0)
        const __half2 hsrc = INIT ? xin : __hadd2(xqh, S);
        const uint32_t hbw = ~(h2u(hsrc) >> 15) & LSB2;
        if constexpr (!INIT) {
            if (J < P.target_n) ones |= hbw;   // uniform: only the first target_node columns count (systematic)
        }
        if constexpr (HB) {
            const bool act = !PAD || h.amask != 0u;
            const uint32_t lo = __ballot_sync(0xffffffffu, act && (hbw & 1u));
            const uint32_t hi = __ballot_sync(0xffffffffu, act && (hbw >> 16));
            sts32(hbrow + (uint32_t)(J * G::C) * 4u, lo);   // every lane stores the same word: no lane-0 branch
            sts32(hbrow + (uint32_t)((G::N + J) * G::C) * 4u, hi);
        }
        if constexpr (INIT) {
#pragma unroll
            for (int u = 0; u < DV; ++u) sts32(addr[u], h2u(xin) | hbw);
        } else {
            const __half2 SX = __hadd2(xin, S);
#pragma unroll
            for (int u = 0; u < DV; ++u) sts32(addr[u], h2u(__hsub2(SX, u2h(cv[u]))) | hbw);   // total - self: exact
        }
    }

    template <int SLOT, int MODE, bool VNW, bool HB, bool MV, bool OG>
    static __device__ __forceinline__ void vn_slot(const KParams &P, const Ctx &c, const H2Ctx &h, uint32_t wvrow, float wvs,
                                                   uint32_t hbrow, uint32_t &ones) {
        constexpr int NT = (G::N - SLOT + G::R - 1) / G::R;
        if constexpr (MODE == VN_INIT_GLOBAL) {
            // this lane's two frames: all of the slot's loads are issued before the first use
            const bool v0 = c.act && c.f0 < c.nvalid, v1 = c.act && c.f1 < c.nvalid;
            const long long o0 = (c.frame0 + (v0 ? c.f0 : 0)) * (long long)P.NZ + c.a_lane;
            const long long o1 = (c.frame0 + (v1 ? c.f1 : 0)) * (long long)P.NZ + c.a_lane;
            float2 x[NT];
            if (P.llr != nullptr) {
                const float *p0 = P.llr + o0, *p1 = P.llr + o1;
                static_for<0, NT>([&](auto n) {
                    constexpr int J = G::vn_order[SLOT + decltype(n)::v * G::R];
                    x[decltype(n)::v].x = v0 ? __ldg(p0 + J * G::z) : 0.0f;
                    x[decltype(n)::v].y = v1 ? __ldg(p1 + J * G::z) : 0.0f;
                });
            } else {   // int8 words in units of q8_step
                const signed char *p0 = P.llr_q8 + o0, *p1 = P.llr_q8 + o1;
                static_for<0, NT>([&](auto n) {
                    constexpr int J = G::vn_order[SLOT + decltype(n)::v * G::R];
                    x[decltype(n)::v].x = v0 ? (float)__ldg(p0 + J * G::z) * P.q8_step : 0.0f;
                    x[decltype(n)::v].y = v1 ? (float)__ldg(p1 + J * G::z) * P.q8_step : 0.0f;
                });
            }
            static_for<0, NT>([&](auto n) {
                constexpr int J = G::vn_order[SLOT + decltype(n)::v * G::R];
                vn_col<J, MODE, VNW, HB, MV, OG>(P, h, wvrow, wvs, hbrow, x[decltype(n)::v], ones);
            });
        } else {
            static_for<0, NT>([&](auto n) {
                constexpr int J = G::vn_order[SLOT + decltype(n)::v * G::R];
                vn_col<J, MODE, VNW, HB, MV, OG>(P, h, wvrow, wvs, hbrow, make_float2(0.0f, 0.0f), ones);
            });
        }
    }

    template <int MODE, bool HB>
    static __device__ __forceinline__ void vn_dispatch(const KParams &P, const Ctx &c, const H2Ctx &h, int trow, int tbuf,
                                                       uint32_t &ones) {
        const uint32_t wvrow = h2_wrow(h, P.h2w_v, trow, P.h2_wv);
        // ballots go to hb[buf][half][j][chunk]
        const uint32_t hbrow = h.sb + (uint32_t)(P.off_hb + tbuf * 2 * G::N * G::C + c.chunk) * 4u;
        if (P.sharing2 != 0) {
            const float wvs = ldsf(wvrow);                       // element 0 of the row: THE weight when sharing code is 3
            const bool og = MODE == VN_ITER && c.og != 0;        // uniform per warp
            static_for<0, G::R>([&](auto s) {
                constexpr int S = decltype(s)::v;
                if (c.slot != S) return;
                if (P.h2_mv != 0) {                              // uniform
                    if (og) vn_slot<S, MODE, true, HB, true, MODE == VN_ITER>(P, c, h, wvrow, wvs, hbrow, ones);
                    else vn_slot<S, MODE, true, HB, true, false>(P, c, h, wvrow, wvs, hbrow, ones);
                } else {
                    if (og) vn_slot<S, MODE, true, HB, false, MODE == VN_ITER>(P, c, h, wvrow, wvs, hbrow, ones);
                    else vn_slot<S, MODE, true, HB, false, false>(P, c, h, wvrow, wvs, hbrow, ones);
                }
            });
        } else {
            static_for<0, G::R>([&](auto s) {
                if (c.slot == decltype(s)::v) vn_slot<decltype(s)::v, MODE, false, HB, false, false>(P, c, h, wvrow, 1.0f, hbrow, ones);
            });
        }
    }

    template <bool INIT>
    static __device__ __forceinline__ void vn_phase(const KParams &P, const Ctx &c, int t, bool need_hb, uint32_t &ones) {
        const H2Ctx h = h2_ctx(P, c);
        if (!INIT && P.app != nullptr) {
            h2_vn_phase_tab<0, INIT>(P, c, h, t, need_hb, ones);   // APP output wanted: the compact table-driven code
            return;
        }
        if constexpr (INIT) {
            vn_dispatch<VN_INIT_SMEM, false>(P, c, h, 0, 0, ones);
        } else {
            // the last iteration has no successor: it reuses its own weight row, and its V->C words only carry the
            // hard bits (mantissa LSB) into the final syndrome pass
            const int trow = min(t + 1, P.T_run - 1);
            if (need_hb) vn_dispatch<VN_ITER, true>(P, c, h, trow, t & 1, ones);
            else vn_dispatch<VN_ITER, false>(P, c, h, trow, t & 1, ones);
        }
    }

    // channel LLRs straight from global memory into the first V->C messages (replaces load + init pass)
    static __device__ __forceinline__ void load_init(const KParams &P, const Ctx &c, uint32_t &offgrid) {
        const H2Ctx h = h2_ctx(P, c);
        vn_dispatch<VN_INIT_GLOBAL, false>(P, c, h, 0, 0, offgrid);
    }

    // ------------------------------------------------------------------ final syndrome pass
    static __device__ __forceinline__ uint32_t synd_phase(const KParams &P, const Ctx &c, int tl) {
        const uint32_t a0 = smem_base() + (uint32_t)c.q * 4u;
        uint32_t bad = 0;
        static_for<0, G::R>([&](auto s) {
            constexpr int SLOT = decltype(s)::v;
            constexpr int NT = (G::M - SLOT + G::R - 1) / G::R;
            if (c.slot == SLOT) {
                static_for<0, NT>([&](auto n) {
                    constexpr int I = G::cn_order[SLOT + decltype(n)::v * G::R];
                    constexpr int E0 = G::row_ptr[I], DC = G::row_ptr[I + 1] - E0;
                    uint32_t par = 0;
#pragma unroll
                    for (int p = 0; p < DC; ++p) par ^= lds32(a0 + (uint32_t)(E0 + p) * LP4);
                    bad |= par;
                });
            }
        });
        return bad;
    }
};

}   // namespace nms
