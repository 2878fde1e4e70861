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
    static constexpr bool FUSED_LOAD = false;
    static constexpr bool TRACKS_GRID = false;
    static __device__ __forceinline__ void setup(const KParams &, int) {}

    template <int DC>
    static __device__ __forceinline__ void cn_class(const KParams &P, const H2Ctx &h, int &p, int end, uint32_t LP4,
                                                    uint32_t wcrow, uint32_t wurow, uint32_t &bad) {
        do {
            const int i = P.cn_order[p];
            const uint32_t a0 = h.sb + (uint32_t)P.row_ptr[i] * LP4 + h.q4;
            cn_row_h2<DC>(P, a0, LP4, h2_w(wcrow, i, P.h2_mc), h2_w(wurow, i, P.h2_mu), bad);
            p += P.R;
        } while (p < end);
    }

    // rows are sorted by degree; slot s owns positions p % R == s: one dispatch per degree class
    static __device__ __forceinline__ void cn_phase(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        const H2Ctx h = h2_ctx(P, c);
        const uint32_t LP4 = (uint32_t)P.LP * 4u;
        const uint32_t wcrow = h2_wrow(h, P.h2w_c, t, P.h2_wc), wurow = h2_wrow(h, P.h2w_u, t, P.h2_wu);
        if constexpr (DCB == 0) {
            for (int n = c.slot; n < P.M; n += P.R) {
                const int i = P.cn_order[n];
                const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
                cn_row_h2_generic(P, h.sb + (uint32_t)e0 * LP4 + h.q4, LP4, dc, h2_w(wcrow, i, P.h2_mc),
                                  h2_w(wurow, i, P.h2_mu), bad);
            }
        } else {
            int p = c.slot;
            for (int k = 0; k < P.n_cn_cls; ++k) {
                const ushort4 cl = P.cn_cls[k];
                const int end = cl.z;
                if (p >= end) continue;
                switch (cl.x) {
#define X(d)                                                                             \
    case (d) + 1:                                                                        \
        if constexpr ((d) < DCB) cn_class<(d) + 1>(P, h, p, end, LP4, wcrow, wurow, bad); \
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
        const H2Ctx h = h2_ctx(P, c);
        h2_vn_phase_tab<DVB, INIT>(P, c, h, t, need_hb, ones);
    }

    static __device__ __forceinline__ uint32_t synd_phase(const KParams &P, const Ctx &c, int tl) {
        return h2_synd_phase(P, c);
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
