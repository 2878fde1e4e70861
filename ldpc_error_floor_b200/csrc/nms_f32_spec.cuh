// nms_f32_spec.cuh -- graph-specialised float32 kernels (one frame per 32-bit lane) for the base graphs known at
// build time: the float min-sum path (decoding_type 1; QM = 0) and the quantised modes the packed kernels do not take
// (q_bit 6, per-edge weights; QM = 1).  Same structure as nms_h2_spec.cuh:
//   * VN phase unrolled per column, every message row offset and circulant rotation an immediate, no branches; the
//     same column code serves the iteration loop (with or without hard-decision ballots), and the fused "channel
//     LLRs from global memory -> first V->C messages" prologue;
//   * CN phase: one body per distinct row degree, rows from the host-built task table -- the iteration loop stays
//     inside the 32 KB instruction cache (a fully unrolled CN phase stalled 58 % of the issue slots on fetch);
//   * final syndrome pass unrolled per row.
// The arithmetic is the shared code of nms_f32.cuh, so results equal the generic float kernels bit for bit.
#pragma once
#include "nms_f32.cuh"

namespace nms {

template <class G, int ROT>
__device__ __forceinline__ uint32_t f32_spec_rot(const F32Ctx &h) {
    if constexpr (ROT == 0) {
        return h.q4;
    } else if constexpr (G::L == G::LP) {
        if constexpr ((G::L & (G::L - 1)) == 0) return (h.q4 + ROT * 4u) & (G::L * 4u - 1u);
        else return f32_rot<false>(h, ROT * 4u, G::L * 4u);
    } else {
        return f32_rot<true>(h, ROT * 4u, G::L * 4u);
    }
}

template <class G, int QM>
struct F32SpecPolicy {
    static constexpr bool H2 = false;
    static constexpr bool FUSED_LOAD = true;
    static constexpr bool TRACKS_GRID = false;
    static constexpr uint32_t LP4 = G::LP * 4u;
    static constexpr bool PAD = G::L != G::LP;
    enum { F_ITER = 0, F_INIT_GLOBAL = 2 };   // VN column code: iteration t / pass before iteration 0 fed from global memory
    static constexpr int ETW = PAD ? 2 : 1;   // words per (edge, chunk) entry of the syndrome table

    // Syndrome table, built once per CTA: entry (e, chunk) tells the lane that serves edge e which two ballot words hold
    // the 32 hard bits the chunk's check lanes see through that edge, and how far to funnel-shift them:
    //   word 0 = lo | hi << 11 | r << 22   (word indices into hb[buf][.], r = first bit)
    //   word 1 (L not a multiple of 32) = column's first word | thr << 16: lanes >= thr wrap to variable lane l - thr
    static __device__ __forceinline__ void setup(const KParams &P, int tid) {
        for (int idx = tid; idx < G::E * G::C; idx += blockDim.x) {
            const int e = idx / G::C, ch = idx - e * G::C;
            const int colC = P.e_col[e] * G::C;
            int s = ch * 32 + P.e_sF[e];
            s = s >= G::L ? s - G::L : s;
            const int w = s >> 5, r = s & 31;
            const int w1 = PAD ? min(w + 1, G::C - 1) : (w + 1 == G::C ? 0 : w + 1);
            nms_smem[P.off_et2 + idx * ETW] = (uint32_t)(colC + w) | ((uint32_t)(colC + w1) << 11) | ((uint32_t)r << 22);
            if constexpr (PAD) nms_smem[P.off_et2 + idx * ETW + 1] = (uint32_t)colC | ((uint32_t)min(G::L - s, 32) << 16);
        }
    }

    // syndrome bit of this lane's check in row [e0, e0 + dc): lane p serves edge p, one XOR reduction (see nms_f32.cuh)
    static __device__ __forceinline__ uint32_t row_syndrome(uint32_t et2c, uint32_t hb4, int e0, int dc) {
        const int lane = threadIdx.x & 31;
        uint32_t f = 0;
        if (lane < dc) {   // dc <= 32 for the graphs that get a specialised kernel
            const uint32_t ea = et2c + (uint32_t)((e0 + lane) * (G::C * ETW * 4));
            const uint32_t tb = lds32(ea);
            f = __funnelshift_r(lds32(hb4 + (tb & 0x7ffu) * 4u), lds32(hb4 + ((tb >> 11) & 0x7ffu) * 4u), tb >> 22);
            if constexpr (PAD) {
                const uint32_t t2 = lds32(ea + 4u), thr = t2 >> 16;
                if (thr < 32u) f = (f & ((1u << thr) - 1u)) | (lds32(hb4 + (t2 & 0xffffu) * 4u) << thr);
            }
        }
        return (__reduce_xor_sync(0xffffffffu, f) >> lane) & 1u;
    }

    // ------------------------------------------------------------------------------ CN phase
    // min-sum rows with one weight per row or per iteration: one counted loop per degree class (a slot's rows come sorted by
    // degree, G::cn_cls_cnt[slot][class] of them), the row list read one entry ahead, weights that do not vary per row loaded
    // once per phase (WROW = false) -- the structure of spec_cn_rows in nms_h2_spec.cuh
    template <bool WROW>
    static __device__ __forceinline__ void cn_rows(const KParams &P, const Ctx &c, int t, uint32_t a00, uint32_t hb4, uint32_t et2c,
                                                   uint32_t w0row, int m0, uint32_t w1row, int m1, uint32_t &bad) {
        constexpr int NT = (G::M + G::R - 1) / G::R;
        float w0 = 1.0f, w1 = 1.0f;
        if constexpr (!WROW) { w0 = ldsf(w0row); w1 = ldsf(w1row); }
        const uint2 *task = P.cn_task + c.slot * NT;
        uint2 tk = task[0];
        static_for<0, G::NDEG>([&](auto k) {
            constexpr int K = decltype(k)::v;
            constexpr int DC = G::cn_degs_desc[K];
            int nk = 0;
            static_for<0, G::R>([&](auto sl) {
                constexpr int CNT = G::cn_cls_cnt[decltype(sl)::v * G::NDEG + K];
                if (c.slot == decltype(sl)::v) nk = CNT;
            });
#pragma unroll 1
            for (int n = 0; n < nk; ++n) {
                const uint2 cur = tk;
                tk = *++task;
                const uint32_t par = row_syndrome(et2c, hb4, (int)(cur.x / LP4), DC);
                bad |= par;
                if constexpr (WROW) {
                    const int i = (int)(cur.y >> 16);
                    w0 = ldsf(w0row + (uint32_t)((i & m0) * 4));
                    w1 = ldsf(w1row + (uint32_t)((i & m1) * 4));
                }
                cn_row_f32<DC, QM>(P, a00 + cur.x, LP4, w0, w1, par);
            }
        });
    }

    static __device__ __forceinline__ void cn_phase(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        const F32Ctx h = f32_ctx(P, c);
        const uint32_t a00 = h.sb + h.q4;
        const uint32_t hb4 = h.sb + (uint32_t)(P.off_hb + ((t + 1) & 1) * G::N * G::C) * 4u;   // hard bits of APP_{t-1}
        const uint32_t et2c = h.sb + (uint32_t)(P.off_et2 + c.chunk * ETW) * 4u;
        // weight rows of iteration t, branch-free (the host only picks these kernels when the weights are staged in shared
        // memory): "no CN weight" reads the 1.0f parked behind the syndrome table, "no UCN weight" aliases the CN row
        const uint32_t one4 = h.et4 + (uint32_t)G::E * 4u;
        const uint32_t w0row = P.sharing0 != 0 ? h.sb + (uint32_t)(P.off_w + P.w_off_cn + t * P.wc) * 4u : one4;
        const int m0 = (P.sharing0 != 0 && P.wc > 1) ? -1 : 0;
        const uint32_t w1row = P.sharing1 != 0 ? h.sb + (uint32_t)(P.off_w + P.w_off_ucn + t * P.wu) * 4u : w0row;
        const int m1 = P.sharing1 != 0 ? (P.wu > 1 ? -1 : 0) : m0;
        if (!(QM == 0 && P.sp) && P.sharing0 != 1) {   // uniform
            if ((m0 | m1) != 0) cn_rows<true>(P, c, t, a00, hb4, et2c, w0row, m0, w1row, m1, bad);
            else cn_rows<false>(P, c, t, a00, hb4, et2c, w0row, m0, w1row, m1, bad);
            return;
        }
        constexpr int NT = (G::M + G::R - 1) / G::R;
        const uint2 *task = P.cn_task + c.slot * NT;   // host-built: {row offset in bytes, degree | row index << 16}
#pragma unroll 1
        for (int n = 0; n < NT; ++n) {
            const uint2 tk = task[n];
            const int dc = (int)(tk.y & 0xffffu), i = (int)(tk.y >> 16);
            if (dc == 0) break;
            const uint32_t a0 = a00 + tk.x;
            const uint32_t par = row_syndrome(et2c, hb4, (int)(tk.x / LP4), dc);
            bad |= par;
            if (QM == 0 && P.sp) cn_row_f32_sp(P, a0, LP4, dc, t, i, (int)(tk.x / LP4), par);        // sum-product (decoding_type 0)
            else cn_row_f32_generic<QM>(P, a0, LP4, dc, t, i, (int)(tk.x / LP4), par);              // per-edge weights
        }
    }

    // ------------------------------------------------------------------------------ VN phase
    // One column, everything constant-folded.  MODE F_ITER: iteration t;  F_INIT_GLOBAL: pass before iteration 0 with the
    // channel value `xg` just loaded from global memory (also fills the xa / xq arrays).
    // VNW: VN weights present (wvrow = the next iteration's row).  The hard decisions of the column go out as one
    // ballot word per chunk (hbrow) and are OR-ed into `onesw` (bit l = lane l's decision has a one so far).
    // TALL: every column counts in the error metrics (target_node = N), so the "has a one" word needs no column test.
    // MV: the VN weight varies per column (sharing code 2) -- else element 0 of the row is THE weight (code 3)
    template <int J, int MODE, bool VNW, bool TALL, bool MV>
    static __device__ __forceinline__ void vn_col(const KParams &P, const F32Ctx &h, uint32_t wvrow, float wvs, uint32_t hbrow,
                                                  float xg, uint32_t &onesw) {
        constexpr int C0 = G::col_ptr[J], DV = G::col_ptr[J + 1] - C0;
        constexpr bool INIT = MODE != F_ITER;
        constexpr uint32_t XJ4 = (uint32_t)(J * G::LP) * 4u;
        uint32_t addr[DV];
        float cv[DV], ext[DV];
        static_for<0, DV>([&](auto u) {
            constexpr int U = decltype(u)::v;
            constexpr uint32_t X4 = (uint32_t)G::vn_e[C0 + U] * LP4;
            addr[U] = h.sb + f32_spec_rot<G, G::vn_rot[C0 + U]>(h) + X4;
            if constexpr (!INIT) cv[U] = ldsf(addr[U]);
        });
        float S = 0.0f;
        if constexpr (!INIT) S = f32_extrinsic<DV>(cv, ext);
        float xa = xg, xqv;
        if constexpr (INIT) {
            if constexpr (QM == 1) {
                xa = fminf(fmaxf(xa, -XA_BOUND), XA_BOUND);
                xqv = qf(P, xa);                                     // :321-322
                sts32(h.xq4 + XJ4, __float_as_uint(xqv));
            } else {
                xa = f32_pos_zero(xa);   // -0.0 (it occurs in [Uncor] files) -> +0.0: same value everywhere it is used
                xqv = xa;
            }
            sts32(h.xa4 + XJ4, __float_as_uint(xa));
        } else {
            if constexpr (QM == 1) {
                xqv = ldsf(h.xq4 + XJ4);
                if constexpr (VNW) xa = ldsf(h.xa4 + XJ4);
            } else {
                xa = ldsf(h.xa4 + XJ4);
                xqv = xa;
            }
        }
        float xin = QM == 1 ? xqv : xa;
        if constexpr (VNW) {
            xin = __fmul_rn(xa, MV ? ldsf(wvrow + (uint32_t)(J * 4)) : wvs);   // :168-169
            xin = QM == 1 ? qf(P, xin) : f32_pos_zero(xin);                    // :176-177
        }
        const float hsrc = INIT ? xin : __fadd_rn(xqv, S);          // :181-182 / :324 (clip_LLR never changes the sign)
        const bool hb = hsrc >= 0.0f;
        const uint32_t b = __ballot_sync(0xffffffffu, (!PAD || h.amask != 0u) && hb);   // hard decisions of this column / chunk
        if constexpr (!INIT) {
            if (TALL || J < P.target_n) onesw |= b;   // uniform: only the first target_node columns count (systematic)
        }
        sts32(hbrow + (uint32_t)(J * G::C) * 4u, b);   // every lane stores the same word: no lane-0 branch
#pragma unroll
        for (int u = 0; u < DV; ++u) sts32(addr[u], f32_v2c<QM>(P, xin, INIT ? 0.0f : ext[u]));
    }

    template <int SLOT, int MODE, bool VNW, bool TALL, bool MV>
    static __device__ __forceinline__ void vn_slot(const KParams &P, const Ctx &c, const F32Ctx &h, uint32_t wvrow, float wvs,
                                                   uint32_t hbrow, uint32_t &onesw) {
        constexpr int NT = (G::N - SLOT + G::R - 1) / G::R;
        if constexpr (MODE == F_INIT_GLOBAL) {
            const bool ok = c.act && c.f0 < c.nvalid;
            const long long o = (c.frame0 + (ok ? c.f0 : 0)) * (long long)P.NZ + c.a_lane;
            float x[NT];   // all of the slot's loads are issued before the first use
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
                vn_col<G::vn_order[SLOT + decltype(n)::v * G::R], MODE, VNW, TALL, MV>(P, h, wvrow, wvs, hbrow, x[decltype(n)::v], onesw);
            });
        } else {
            static_for<0, NT>([&](auto n) {
                vn_col<G::vn_order[SLOT + decltype(n)::v * G::R], MODE, VNW, TALL, MV>(P, h, wvrow, wvs, hbrow, 0.0f, onesw);
            });
        }
    }

    template <int MODE>
    static __device__ __forceinline__ void vn_dispatch(const KParams &P, const Ctx &c, const F32Ctx &h, int trow, int tbuf,
                                                       uint32_t &ones) {
        const uint32_t hbrow = h.sb + (uint32_t)(P.off_hb + tbuf * G::N * G::C + c.chunk) * 4u;   // hb[buf][j][chunk]
        uint32_t onesw = 0;
        const uint32_t wvrow = h.sb + (uint32_t)(P.off_w + P.w_off_vn + trow * P.wv) * 4u;
        const bool mv = P.wv > 1;
        const float wvs = P.sharing2 != 0 ? ldsf(wvrow) : 1.0f;   // THE weight of the iteration when it does not vary per column
        const bool tall = MODE != F_ITER || P.target_n >= G::N;
        static_for<0, G::R>([&](auto s) {
            constexpr int S = decltype(s)::v;
            if (c.slot == S) {
                if (P.sharing2 != 0) {
                    if (mv) {
                        if (tall) vn_slot<S, MODE, true, true, true>(P, c, h, wvrow, wvs, hbrow, onesw);
                        else vn_slot<S, MODE, true, false, true>(P, c, h, wvrow, wvs, hbrow, onesw);
                    } else {
                        if (tall) vn_slot<S, MODE, true, true, false>(P, c, h, wvrow, wvs, hbrow, onesw);
                        else vn_slot<S, MODE, true, false, false>(P, c, h, wvrow, wvs, hbrow, onesw);
                    }
                } else {
                    if (tall) vn_slot<S, MODE, false, true, false>(P, c, h, 0u, 1.0f, hbrow, onesw);
                    else vn_slot<S, MODE, false, false, false>(P, c, h, 0u, 1.0f, hbrow, onesw);
                }
            }
        });
        ones |= (onesw >> c.lane) & 1u;
    }

    static __device__ __forceinline__ bool unrolled_ok(const KParams &P) { return P.app == nullptr; }   // APP output: table code

    template <bool INIT>
    static __device__ __forceinline__ void vn_phase(const KParams &P, const Ctx &c, int t, bool need_hb, uint32_t &ones) {
        const F32Ctx h = f32_ctx(P, c);
        if (INIT || !unrolled_ok(P)) {   // channel values already in shared memory (generator), or APP output wanted
            f32_vn_phase_tab<G::DVMAX, INIT, QM, PAD ? 1 : 0>(P, c, h, t, ones);
            return;
        }
        const int trow = min(t + 1, P.T_run - 1);
        vn_dispatch<F_ITER>(P, c, h, trow, t & 1, ones);
    }

    // channel LLRs straight from global memory into the first V->C messages (replaces load + init pass)
    static __device__ __forceinline__ void load_init(const KParams &P, const Ctx &c, uint32_t &) {
        const F32Ctx h = f32_ctx(P, c);
        uint32_t dummy = 0;
        vn_dispatch<F_INIT_GLOBAL>(P, c, h, 0, 1, dummy);   // hard bits of xin_0 -> ballot buffer 1
    }

    // ------------------------------------------------------------------ final syndrome pass
    static __device__ __forceinline__ uint32_t synd_phase(const KParams &P, const Ctx &c, int tl) {
        const F32Ctx h = f32_ctx(P, c);
        const uint32_t hb4 = h.sb + (uint32_t)(P.off_hb + ((tl + 1) & 1) * G::N * G::C) * 4u;
        const uint32_t et2c = h.sb + (uint32_t)(P.off_et2 + c.chunk * ETW) * 4u;
        constexpr int NT = (G::M + G::R - 1) / G::R;
        const uint2 *task = P.cn_task + c.slot * NT;
        uint32_t bad = 0;
#pragma unroll 1
        for (int n = 0; n < NT; ++n) {
            const uint2 tk = task[n];
            const int dc = (int)(tk.y & 0xffffu);
            if (dc == 0) break;
            bad |= row_syndrome(et2c, hb4, (int)(tk.x / LP4), dc);
        }
        return bad;
    }
};

}   // namespace nms
