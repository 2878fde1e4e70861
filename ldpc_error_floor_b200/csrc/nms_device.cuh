// nms_device.cuh -- device-side skeleton shared by the packed (fp16x2) and float NMS kernels.
//
// One persistent CTA decodes FB frames per batch entirely out of shared memory:
//   load (global LLRs | fused Philox BPSK/AWGN generator)  ->  T x { CN phase ; VN phase }
//   -> final syndrome pass -> outputs (packed hard bits, flags, counters, harvested words).
// HBM sees LLRs in and bits/flags out; every edge message lives in shared memory / registers.
// The arithmetic back-end is a Policy (nms_h2.cu / nms_f32.cu) providing
//   cn_task(P,c,i,t,bad)  vn_task<INIT>(P,c,j,t,ones)  synd_row(P,c,i,tl)
//
// Reference semantics restated: Main_Functions.py:161-335 (steps D2..D8 of SURVEY.md 8a),
// quantiser :475-494, sample generation Print_Functions.py:29-72, metrics :100-118.
#pragma once
#include "nms_common.cuh"

namespace nms {

constexpr float RINT_MAGIC = 12582912.0f;   // 1.5 * 2^23: (t + M) - M == rintf(t) (half-to-even) for |t| < 2^22
constexpr float XA_BOUND = 1.0e5f;          // QMS inputs are clamped here so the magic rint stays exact
constexpr uint32_t SIGN2 = 0x80008000u;
constexpr uint32_t LSB2 = 0x00010001u;

// misc shared words
constexpr int MISC_SYND = 0;      // [2][64] syndrome-bad flag of the previous APP, by iteration parity
constexpr int MISC_ONES = 128;    // [2][64] "hard decision has a one" flag, by iteration parity
constexpr int MISC_BITERR = 256;  // [64]
constexpr int MISC_HIDX = 320;    // [64] harvest row index (or 0xffffffff)
constexpr int MISC_CTRL = 384;    // [16]
constexpr int MISC_WORDS = NMS_MISC_WORDS;
static_assert(MISC_WORDS >= 400, "misc layout");
// ctrl words
constexpr int CTRL_NEWLY = 0;     // [2] frames to copy out now
constexpr int CTRL_FROZEN = 2;    // [2][2] frozen mask by iteration parity
constexpr int CTRL_NEWONES = 6;   // [2] of those, frames whose decision has a one
constexpr int CTRL_HARVEST = 8;   // [2]

__device__ __forceinline__ float rint_magic(float t) { return __fsub_rn(__fadd_rn(t, RINT_MAGIC), RINT_MAGIC); }

// Q(x) with x already multiplied by qk: clamp(rint(t), +-maxk) (caller multiplies by qinv)
__device__ __forceinline__ float qcore(float t, float maxk) {
    return fminf(fmaxf(rint_magic(t), -maxk), maxk);
}
__device__ __forceinline__ float qf(const KParams &P, float x) {   // full float quantiser
    return __fmul_rn(qcore(__fmul_rn(x, P.qk), P.qmaxk), P.qinv);
}

__device__ __forceinline__ float cn_w(const float *w, int code, int width, int t, int i, int e) {
    if (code == 3) return __ldg(w + (size_t)t * width);
    if (code == 2) return __ldg(w + (size_t)t * width + i);
    return __ldg(w + (size_t)t * width + e);
}
__device__ __forceinline__ float vn_w(const KParams &P, int t, int j) {
    if (P.sharing2 == 3) return __ldg(P.w_vn + (size_t)t * P.wv);
    if (P.sharing2 == 2) return __ldg(P.w_vn + (size_t)t * P.wv + j);
    return 1.0f;
}

struct Ctx {
    uint32_t *msg, *xq, *hb, *misc;
    float *xa;
    int lane, chunk, slot, q, qe;
    bool active;
    int f0, f1;   // frame(s) of this lane's slot (packed: f0 = 2fp, f1 = 2fp+1; float: f0 = f1 = fp)
    int a_lane;   // circulant lane of q
    long long frame0;
    int nvalid;
};

// is frame f frozen as far as the VN phase of iteration t can tell?  (only used for the optional APP output)
__device__ __forceinline__ bool app_frozen(const KParams &P, const Ctx &c, int t, int f) {
    const uint32_t *ctrl = c.misc + MISC_CTRL;
    bool frozen = (ctrl[CTRL_FROZEN + ((t + 1) & 1) * 2 + (f >> 5)] >> (f & 31)) & 1u;
    if (P.early_term && t >= 1 && c.misc[MISC_SYND + (t & 1) * 64 + f] == 0u) frozen = true;
    return frozen;
}
__device__ __forceinline__ void app_store(const KParams &P, const Ctx &c, int j, int t, int f, float v) {
    if (f >= c.nvalid || app_frozen(P, c, t, f)) return;
    const long long tt = P.app_all ? t : 0;
    P.app[tt * P.app_stride_t + (c.frame0 + f) * (long long)P.NZ + j * P.z + c.a_lane] = v;
}

// ------------------------------------------------------------------------- sample generation
// One Philox4x32-10 block -> four N(0,1) via Box-Muller -> four channel LLRs of frame F, bits 4*quad..4*quad+3.
// Print_Functions.py:45-60: x = n*sigma - 1 (all-zero word), llr = 2x/sigma^2, quantise, puncture, shorten.
__device__ __forceinline__ void gen_llr4(const KParams &P, unsigned long long F, int quad, float out[4]) {
    uint32_t r[4];
    philox4x32_10((uint32_t)F, (uint32_t)(F >> 32), (uint32_t)quad, 0u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), r);
    float n[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float u1 = fmaf((float)r[2 * h], 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (r+0.5)/2^32
        const float u2 = (float)r[2 * h + 1] * 2.3283064365386963e-10f;
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        n[2 * h] = rad * cs;
        n[2 * h + 1] = rad * sn;
    }
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
        const int k = 4 * quad + k4 + 1;   // 1-based bit index
        float llr = __fmul_rn(__fadd_rn(__fmul_rn(n[k4], P.sigma), -1.0f), P.two_over_s2);
        if (P.qms) llr = qf(P, llr);                                          // :49-50
        if (P.punct_s > 0 && k >= P.punct_s && k <= P.punct_e) llr = 0.0f;    // :53-57
        if (P.short_s > 0 && k >= P.short_s && k <= P.short_e) llr = -P.clip; // :59-60
        out[k4] = llr;
    }
}

template <bool H2>
__device__ __forceinline__ void store_xa(const KParams &P, float *xa, int f, int k, float v) {
    const int j = k / P.z, a = k - j * P.z;
    if (P.qms) v = fminf(fmaxf(v, -XA_BOUND), XA_BOUND);
    if (H2) {
        const int qq = a * P.Fp + (f >> 1);
        xa[(j * P.LP + qq) * 2 + (f & 1)] = v;
    } else {
        xa[j * P.LP + a * P.Fp + f] = v;
    }
}

// gather the packed hard decision of the frames in `mask` from the ballot array `hbuf`
template <bool H2>
__device__ __forceinline__ void copy_out(const KParams &P, const Ctx &c, const uint32_t mask[2], const uint32_t onesm[2],
                                         int hbuf) {
    const int nh = H2 ? 2 : 1;
    for (int item = threadIdx.x; item < P.FB * P.HW; item += blockDim.x) {
        const int f = item / P.HW, w = item - f * P.HW;
        if (!((mask[f >> 5] >> (f & 31)) & 1u)) continue;
        uint32_t outw = 0;
        if ((onesm[f >> 5] >> (f & 31)) & 1u) {
            const int half = H2 ? (f & 1) : 0, fp = H2 ? (f >> 1) : f;
            const uint32_t *hb = c.hb + (size_t)(hbuf * nh + half) * P.N * P.C;
            int k = 32 * w;
            int j = k / P.z, a = k - j * P.z;
            for (int b = 0; b < 32 && k < P.NZ; ++b, ++k) {
                const int qq = a * P.Fp + fp;
                outw |= ((hb[j * P.C + (qq >> 5)] >> (qq & 31)) & 1u) << b;
                if (++a == P.z) { a = 0; ++j; }
            }
            if (outw) atomicAdd(&c.misc[MISC_BITERR + f], (uint32_t)__popc(outw));
        }
        if (P.hard != nullptr) P.hard[(c.frame0 + f) * (long long)P.HW + w] = outw;
    }
}

// =============================================================================== main kernel
template <class Policy>
__device__ __forceinline__ void nms_decode_body(const KParams &P) {
    constexpr bool H2 = Policy::H2;
    extern __shared__ __align__(16) uint32_t smem[];
    Ctx c;
    c.msg = smem + P.off_msg;
    c.xa = reinterpret_cast<float *>(smem + P.off_xa);
    c.xq = smem + P.off_xq;
    c.hb = smem + P.off_hb;
    c.misc = smem + P.off_misc;
    const int tid = threadIdx.x;
    c.lane = tid & 31;
    const int warp = tid >> 5;
    c.chunk = warp % P.C;
    c.slot = warp / P.C;
    c.q = c.chunk * 32 + c.lane;
    c.active = c.q < P.L;
    c.qe = c.active ? c.q : 0;
    c.a_lane = c.qe / P.Fp;
    {
        const int fp = c.qe - c.a_lane * P.Fp;
        c.f0 = H2 ? 2 * fp : fp;
        c.f1 = H2 ? 2 * fp + 1 : fp;
    }
    uint32_t *ctrl = c.misc + MISC_CTRL;
    const long long nbatches = (P.n_frames + P.FB - 1) / P.FB;

    for (long long batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
        c.frame0 = batch * P.FB;
        c.nvalid = (int)min((long long)P.FB, P.n_frames - c.frame0);
        __syncthreads();   // previous batch fully retired before shared memory is reused

        // ---------------- load channel LLRs into xa (zero for padding frames)
        if (P.llr != nullptr) {
            const int tot = P.FB * P.NZ;
            for (int idx = tid; idx < tot; idx += blockDim.x) {
                const int f = idx / P.NZ, k = idx - f * P.NZ;
                const float v = f < c.nvalid ? __ldg(P.llr + (c.frame0 + f) * (long long)P.NZ + k) : 0.0f;
                store_xa<H2>(P, c.xa, f, k, v);
            }
        } else {
            const int nquads = (P.NZ + 3) >> 2, tot = P.FB * nquads;
            for (int idx = tid; idx < tot; idx += blockDim.x) {
                const int f = idx / nquads, quad = idx - f * nquads;
                float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (f < c.nvalid) gen_llr4(P, P.frame_offset + (unsigned long long)(c.frame0 + f), quad, v);
                for (int k4 = 0; k4 < 4; ++k4)
                    if (4 * quad + k4 < P.NZ) store_xa<H2>(P, c.xa, f, 4 * quad + k4, v[k4]);
            }
        }
        for (int idx = tid; idx < MISC_WORDS; idx += blockDim.x) c.misc[idx] = 0;
        // per-frame state lives in the registers of threads 0..63 (thread f owns frame f)
        bool st_frozen = tid >= c.nvalid, st_synd_ever = false, st_ever_correct = false;
        bool st_out_synd_ok = false, st_out_one = false;
        int st_iters = P.T_run, st_executed = P.T_run;
        __syncthreads();
        if (tid < 64) {
            const uint32_t fm = __ballot_sync(0xffffffffu, st_frozen);
            if (c.lane == 0) { ctrl[CTRL_FROZEN + warp] = fm; ctrl[CTRL_FROZEN + 2 + warp] = fm; }
        }

        // ---------------- init pass: xq, first V->C messages, hard bits of xin_0
        {
            uint32_t dummy = 0;
            for (int n = c.slot; n < P.N; n += P.R) Policy::template vn_task<true>(P, c, P.vn_order[n], -1, dummy);
        }
        __syncthreads();

        bool alldone = false;
        int t = 0;
        for (; t < P.T_run; ++t) {
            // ======== CN phase (also yields the syndrome of the previous hard decision)
            uint32_t bad = 0;
            for (int n = c.slot; n < P.M; n += P.R) Policy::cn_task(P, c, P.cn_order[n], t, bad);
            if (t >= 1 && c.active) {
                uint32_t *sy = c.misc + MISC_SYND + (t & 1) * 64;
                if (bad & 1u) sy[c.f0] = 1u;
                if (H2 && (bad & 0x10000u)) sy[c.f1] = 1u;
            }
            __syncthreads();   // A
            // ======== per-frame bookkeeping for APP_{t-1} (threads 0..63), concurrent with the VN phase
            if (tid < 64) {
                bool newly = false;
                if (t >= 1) {
                    const bool fbad = c.misc[MISC_SYND + (t & 1) * 64 + tid] != 0u;
                    const bool one = c.misc[MISC_ONES + ((t - 1) & 1) * 64 + tid] != 0u;
                    c.misc[MISC_ONES + ((t - 1) & 1) * 64 + tid] = 0u;
                    if (!st_frozen) {
                        if (!one) st_ever_correct = true;
                        if (!fbad && !st_synd_ever) { st_synd_ever = true; st_iters = t; }
                        if (P.early_term && !fbad) {
                            st_frozen = true; newly = true;
                            st_out_synd_ok = true; st_out_one = one; st_executed = t;
                        }
                    }
                }
                c.misc[MISC_SYND + ((t + 1) & 1) * 64 + tid] = 0u;
                const uint32_t nm = __ballot_sync(0xffffffffu, newly);
                const uint32_t no = __ballot_sync(0xffffffffu, newly && st_out_one);
                const uint32_t fm = __ballot_sync(0xffffffffu, st_frozen);
                if (c.lane == 0) {
                    ctrl[CTRL_NEWLY + warp] = nm;
                    ctrl[CTRL_NEWONES + warp] = no;
                    ctrl[CTRL_FROZEN + (t & 1) * 2 + warp] = fm;
                }
            }
            // ======== VN phase
            uint32_t ones = 0;
            for (int n = c.slot; n < P.N; n += P.R) Policy::template vn_task<false>(P, c, P.vn_order[n], t, ones);
            if (c.active) {
                uint32_t *on = c.misc + MISC_ONES + (t & 1) * 64;
                if (ones & 1u) on[c.f0] = 1u;
                if (H2 && (ones & 0x10000u)) on[c.f1] = 1u;
            }
            __syncthreads();   // B
            if (P.early_term) {
                const uint32_t nm[2] = {ctrl[CTRL_NEWLY], ctrl[CTRL_NEWLY + 1]};
                if (nm[0] | nm[1]) {
                    const uint32_t no[2] = {ctrl[CTRL_NEWONES], ctrl[CTRL_NEWONES + 1]};
                    copy_out<H2>(P, c, nm, no, (t + 1) & 1);   // hard bits of APP_{t-1}
                }
                alldone = (ctrl[CTRL_FROZEN + (t & 1) * 2] & ctrl[CTRL_FROZEN + (t & 1) * 2 + 1]) == 0xffffffffu;
                if (alldone) break;
            }
        }

        if (!alldone) {
            // ---------------- syndrome of the last hard decision APP_{T-1}
            uint32_t bad = 0;
            const int tl = P.T_run;
            for (int n = c.slot; n < P.M; n += P.R) bad |= Policy::synd_row(P, c, P.cn_order[n], tl);
            if (c.active) {
                uint32_t *sy = c.misc + MISC_SYND + (tl & 1) * 64;
                if (bad & 1u) sy[c.f0] = 1u;
                if (H2 && (bad & 0x10000u)) sy[c.f1] = 1u;
            }
            __syncthreads();
            if (tid < 64) {
                const bool fbad = c.misc[MISC_SYND + (tl & 1) * 64 + tid] != 0u;
                const bool one = c.misc[MISC_ONES + ((tl - 1) & 1) * 64 + tid] != 0u;
                const bool pending = !st_frozen;
                if (pending) {
                    if (!one) st_ever_correct = true;
                    if (!fbad && !st_synd_ever) { st_synd_ever = true; st_iters = tl; }
                    st_out_synd_ok = !fbad; st_out_one = one; st_executed = tl;
                }
                const uint32_t nm = __ballot_sync(0xffffffffu, pending);
                const uint32_t no = __ballot_sync(0xffffffffu, pending && one);
                if (c.lane == 0) { ctrl[CTRL_NEWLY + warp] = nm; ctrl[CTRL_NEWONES + warp] = no; }
            }
            __syncthreads();
            const uint32_t nm[2] = {ctrl[CTRL_NEWLY], ctrl[CTRL_NEWLY + 1]};
            const uint32_t no[2] = {ctrl[CTRL_NEWONES], ctrl[CTRL_NEWONES + 1]};
            copy_out<H2>(P, c, nm, no, (tl - 1) & 1);
        }
        __syncthreads();   // bit-error counts complete

        // ---------------- per-frame results, Monte-Carlo counters, harvest
        if (tid < 64) {
            const bool valid = tid < c.nvalid;
            const uint32_t be = c.misc[MISC_BITERR + tid];
            const bool uncor_any = !st_ever_correct, uncor_last = st_out_one;
            if (valid) {
                const long long F = c.frame0 + tid;
                if (P.iters) P.iters[F] = st_iters;
                if (P.flags)
                    P.flags[F] = (uint8_t)((st_out_synd_ok ? 1u : 0u) | (uncor_any ? 2u : 0u) | (uncor_last ? 4u : 0u) |
                                           (st_synd_ever ? 8u : 0u));
                if (P.biterr) P.biterr[F] = (int)be;
            }
            bool harvest = false;
            if (valid && P.harvest_mode != 0)
                harvest = P.harvest_mode == 1 ? uncor_any : (P.harvest_mode == 2 ? uncor_last : !st_out_synd_ok);
            uint32_t hidx = 0xffffffffu;
            if (harvest && P.uncor_count != nullptr) {
                hidx = atomicAdd(P.uncor_count, 1u);
                if (hidx >= P.uncor_cap || P.uncor_buf == nullptr) hidx = 0xffffffffu;
            }
            c.misc[MISC_HIDX + tid] = hidx;
            const uint32_t hm = __ballot_sync(0xffffffffu, hidx != 0xffffffffu);
            if (c.lane == 0) ctrl[CTRL_HARVEST + warp] = hm;
            if (P.counters != nullptr) {
                const unsigned v0 = __reduce_add_sync(0xffffffffu, valid ? 1u : 0u);
                const unsigned v1 = __reduce_add_sync(0xffffffffu, valid && uncor_last ? 1u : 0u);
                const unsigned v2 = __reduce_add_sync(0xffffffffu, valid && uncor_any ? 1u : 0u);
                const unsigned v3 = __reduce_add_sync(0xffffffffu, valid ? be : 0u);
                const unsigned v4 = __reduce_add_sync(0xffffffffu, valid ? (unsigned)st_executed : 0u);
                const unsigned v5 = __reduce_add_sync(0xffffffffu, valid && !st_out_synd_ok ? 1u : 0u);
                const unsigned v6 = __reduce_add_sync(0xffffffffu, valid && st_out_synd_ok && uncor_last ? 1u : 0u);
                const unsigned v7 = __reduce_add_sync(0xffffffffu, harvest ? 1u : 0u);
                if (c.lane == 0) {
                    if (v0) atomicAdd(P.counters + 0, (unsigned long long)v0);
                    if (v1) atomicAdd(P.counters + 1, (unsigned long long)v1);
                    if (v2) atomicAdd(P.counters + 2, (unsigned long long)v2);
                    if (v3) atomicAdd(P.counters + 3, (unsigned long long)v3);
                    if (v4) atomicAdd(P.counters + 4, (unsigned long long)v4);
                    if (v5) atomicAdd(P.counters + 5, (unsigned long long)v5);
                    if (v6) atomicAdd(P.counters + 6, (unsigned long long)v6);
                    if (v7) atomicAdd(P.counters + 7, (unsigned long long)v7);
                }
            }
        }
        if (P.harvest_mode != 0 && P.uncor_buf != nullptr) {
            __syncthreads();
            const uint32_t hm[2] = {ctrl[CTRL_HARVEST], ctrl[CTRL_HARVEST + 1]};
            if (hm[0] | hm[1]) {
                for (int f = 0; f < P.FB; ++f) {
                    if (!((hm[f >> 5] >> (f & 31)) & 1u)) continue;
                    const uint32_t row = c.misc[MISC_HIDX + f];
                    for (int k = tid; k < P.NZ; k += blockDim.x) {
                        const int j = k / P.z, a = k - j * P.z;
                        const float v = H2 ? c.xa[(j * P.LP + a * P.Fp + (f >> 1)) * 2 + (f & 1)]
                                           : c.xa[j * P.LP + a * P.Fp + f];
                        P.uncor_buf[(size_t)row * P.NZ + k] = v;
                    }
                }
            }
        }
    }
}

// fall-through chains: REP_DESC(X) expands X(31) X(30) ... X(0)
#define NMS_REP_DESC(X)                                                                                        \
    X(31) X(30) X(29) X(28) X(27) X(26) X(25) X(24) X(23) X(22) X(21) X(20) X(19) X(18) X(17) X(16) X(15) X(14) \
    X(13) X(12) X(11) X(10) X(9) X(8) X(7) X(6) X(5) X(4) X(3) X(2) X(1) X(0)

}   // namespace nms
