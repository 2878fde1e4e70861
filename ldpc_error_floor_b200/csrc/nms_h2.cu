// nms_h2.cu -- degree-bucketed generic packed kernels: any base graph, tables read from the constant bank.
// Compiled once per bucket: -DNMS_DCB=<max row degree> -DNMS_DVB=<max column degree> (0/0 = any degree).
// Graphs known at build time get a graph-specialised kernel instead (nms_h2_spec.cuh + gen_spec.py).
#include "nms_h2.cuh"

#ifndef NMS_DCB
#define NMS_DCB 16
#endif
#ifndef NMS_DVB
#define NMS_DVB 8
#endif

namespace nms {

template <int DCB, int DVB>
struct H2Policy {
    static constexpr bool H2 = true;

    static __device__ __forceinline__ void cn_phase(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        for (int n = c.slot; n < P.M; n += P.R) {
            const int i = P.cn_order[n];
            const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
            const int off = e0 * P.LP + c.q;
            const float w0 = h2_wcn(P, t, i), w1 = h2_wucn(P, t, i);
            if constexpr (DCB == 0) {
                cn_row_h2_generic(P, off, P.LP, dc, w0, w1, bad);
            } else {
                switch (dc) {
#define X(p)                                                              \
    case (p) + 1:                                                         \
        if constexpr ((p) < DCB) cn_row_h2<(p) + 1>(P, off, P.LP, w0, w1, bad); \
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
                vn_col_h2_generic<INIT>(P, c, j, t, need_hb, ones);
            } else {
                const int dv = P.col_ptr[j + 1] - P.col_ptr[j];
                switch (dv) {
#define X(p)                                                                         \
    case (p) + 1:                                                                    \
        if constexpr ((p) < DVB) vn_col_h2<(p) + 1, INIT>(P, c, j, t, need_hb, ones); \
        break;
                    NMS_REP_DESC(X)
#undef X
                default: break;
                }
            }
        }
    }

    // syndrome parity of the hard bits parked in the message LSBs after the last VN phase
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
