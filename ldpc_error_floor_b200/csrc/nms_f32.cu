// nms_f32.cu -- float32 arithmetic back-end (one frame per 32-bit word), generic kernels: any base graph.
//
// Compiled once per degree bucket like nms_h2.cu (-DNMS_DCB / -DNMS_DVB, 0/0 = any degree).  Serves decoding_type 1
// (min-sum, clip +-clip_LLR) and every quantised configuration the packed kernel does not take (q_bit 6, per-edge
// weights) for graphs without a specialised kernel (nms_f32_spec.cuh); the arithmetic is the shared code of
// nms_f32.cuh, so both kernel families give identical results.
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
    static constexpr bool TRACKS_GRID = false;
    static __device__ __forceinline__ void setup(const KParams &, int) {}

    static __device__ __forceinline__ void cn_phase(const KParams &P, const Ctx &c, int t, uint32_t &bad) {
        const F32Ctx h = f32_ctx(P, c);
        const uint32_t a00 = h.sb + h.q4, stride4 = (uint32_t)P.LP * 4u;
        const uint32_t hb4 = h.sb + (uint32_t)(P.off_hb + ((t + 1) & 1) * P.N * P.C) * 4u;   // hard bits of APP_{t-1} (init pass: buffer 1)
        const bool pad = P.L != P.LP;
        for (int n = c.slot; n < P.M; n += P.R) {
            const int i = P.cn_order[n];
            const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
            const uint32_t a0 = a00 + (uint32_t)e0 * stride4;
            const uint32_t par = pad ? f32_row_syndrome<true>(P, h, hb4, e0, dc) : f32_row_syndrome<false>(P, h, hb4, e0, dc);
            bad |= par;
            if (P.sp) {
                cn_row_f32_sp(P, a0, stride4, dc, t, i, e0, par);
            } else if (DCB == 0 || P.sharing0 == 1) {
                cn_row_f32_generic<2>(P, a0, stride4, dc, t, i, e0, par);
            } else {
                float w0, w1;
                f32_row_weights(P, t, i, w0, w1);
                switch (dc) {
#define X(p)                                                                       \
    case (p) + 1:                                                                  \
        if constexpr ((p) < DCB) cn_row_f32<(p) + 1, 2>(P, a0, stride4, w0, w1, par); \
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
        const F32Ctx h = f32_ctx(P, c);
        f32_vn_phase_tab<DVB, INIT, 2>(P, c, h, t, ones);   // float kernels publish their ballots every iteration
    }

    static __device__ __forceinline__ uint32_t synd_phase(const KParams &P, const Ctx &c, int tl) { return f32_synd_phase(P, c, (tl + 1) & 1); }
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
