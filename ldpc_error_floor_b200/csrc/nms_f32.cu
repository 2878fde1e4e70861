// nms_f32.cu -- float32 arithmetic back-end (one frame per 32-bit word): every reference mode.
//
// Compiled once per degree bucket like nms_h2.cu (-DNMS_DCB / -DNMS_DVB, 0/0 = any degree).
// Serves decoding_type 1 (min-sum, clip +-clip_LLR) and every quantised configuration the
// packed kernel does not take (q_bit 6, per-edge weights).  Operation order follows the TF
// graph: direct extrinsic V->C sums in ascending E(C) order (Main_Functions.py:213-215), the
// 1e-4 zero / minimum rules (:230, :250), |.|*w -> ReLU -> clip/quantise -> sign (:267-316),
// APP = clip(xq + sum) (:317-325).  Hard decisions are kept as ballot words per (column, lane
// chunk) so the CN phase can form the syndrome of the previous iteration exactly.
#include "nms_f32.cuh"

#ifndef NMS_DCB
#define NMS_DCB 16
#endif
#ifndef NMS_DVB
#define NMS_DVB 8
#endif

namespace nms {

template <int DCB, int DVB>
struct F32Policy {
    static constexpr bool H2 = false;
    static constexpr bool FUSED_LOAD = false;

    static __device__ __forceinline__ void cn_phase(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        for (int n = c.slot; n < P.M; n += P.R) {
            const int i = P.cn_order[n];
            if constexpr (DCB == 0) {
                cn_row_f32_generic(P, c, i, t, bad);
            } else {
                const int dc = P.row_ptr[i + 1] - P.row_ptr[i];
                switch (dc) {
#define X(p)                                                        \
    case (p) + 1:                                                   \
        if constexpr ((p) < DCB) cn_row_f32<(p) + 1>(P, c, i, t, bad); \
        break;
                    NMS_REP_DESC(X)
#undef X
                default: break;
                }
            }
        }
    }

    template <bool INIT>
    static __device__ __forceinline__ void vn_phase(const KParams &P, const Ctx &c, int t, bool need_hb, uint32_t &ones) {
        for (int n = c.slot; n < P.N; n += P.R) {
            const int j = P.vn_order[n];
            if constexpr (DVB == 0) {
                vn_col_f32_generic<INIT>(P, c, j, t, ones);
            } else {
                const int dv = P.col_ptr[j + 1] - P.col_ptr[j];
                switch (dv) {
#define X(p)                                                                 \
    case (p) + 1:                                                            \
        if constexpr ((p) < DVB) vn_col_f32<(p) + 1, INIT>(P, c, j, t, ones); \
        break;
                    NMS_REP_DESC(X)
#undef X
                default: break;
                }
            }
        }
    }

    static __device__ __forceinline__ uint32_t synd_phase(const KParams &P, const Ctx &c, int tl) {
        uint32_t bad = 0;
        for (int n = c.slot; n < P.M; n += P.R) {
            const int i = P.cn_order[n];
            const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
            uint32_t par = 0;
            for (int p = 0; p < dc; ++p) par ^= f32_hbit(P, c, (tl + 1) & 1, e0 + p);
            bad |= par;
        }
        return bad;
    }
};

#define NMS_CAT2(a, b, c, d) a##b##c##d
#define NMS_KNAME(dc, dv) NMS_CAT2(nms_f32_kernel_, dc, _, dv)

__global__ void __launch_bounds__(512, 1) NMS_KNAME(NMS_DCB, NMS_DVB)(const __grid_constant__ KParams P) {
    nms_decode_body<F32Policy<NMS_DCB, NMS_DVB>>(P);
}

}   // namespace nms

#define NMS_CAT3(a, b, c, d) a##b##c##d
#define NMS_FNAME(dc, dv) NMS_CAT3(nms_f32_func_, dc, _, dv)
extern "C" const void *NMS_FNAME(NMS_DCB, NMS_DVB)() { return (const void *)nms::NMS_KNAME(NMS_DCB, NMS_DVB); }
