// nms_device.cuh -- device-side skeleton shared by the packed (fp16x2) and float NMS kernels.
//
// One persistent CTA decodes FB frames per batch entirely out of shared memory:
//   load (global LLRs | fused Philox BPSK/AWGN generator)  ->  T x { CN phase ; VN phase }
//   -> final syndrome pass -> outputs (packed hard bits, flags, counters, harvested words).
// HBM sees LLRs in and bits/flags out; every edge message lives in shared memory / registers.
// The arithmetic back-end is a Policy (nms_h2.cuh / nms_f32.cu / generated graph-specialised
// policies) providing
//   setup(P,tid)   cn_phase(P,c,t,bad)   vn_phase<INIT>(P,c,t,need_hb,ones)   synd_phase(P,c,tl) -> bad
//
// Shared memory is ONE array `nms_smem` addressed by word offsets (message array at word 0), so
// every access compiles to LDS/STS [reg + immediate]; inactive padding lanes (q >= L) work on
// their own padding words instead of being branched around.
//
// Reference semantics restated: Main_Functions.py:161-335 (steps D2..D8 of SURVEY.md 8a),
// quantiser :475-494, sample generation Print_Functions.py:29-72, metrics :100-118.
#pragma once
#include "nms_common.cuh"

extern __shared__ __align__(16) uint32_t nms_smem[];

namespace nms {

constexpr float XA_BOUND = 1.0e5f;          // QMS inputs are clamped here so the magic rounding stays exact
constexpr uint32_t SIGN2 = 0x80008000u;
constexpr uint32_t LSB2 = 0x00010001u;

// misc shared words
constexpr int MISC_SYND = 0;      // [2][2] frame masks "syndrome of the previous APP is bad", by iteration parity
constexpr int MISC_ONES = 8;      // [2][2] frame masks "hard decision has a one", by iteration parity
constexpr int MISC_BITERR = 256;  // [64]
constexpr int MISC_HIDX = 320;    // [64] harvest row index (or 0xffffffff)
constexpr int MISC_CTRL = 384;    // [16]
static_assert(NMS_MISC_WORDS >= 400, "misc layout");
// ctrl words
constexpr int CTRL_NEWLY = 0;     // [2] frames to copy out now
constexpr int CTRL_FROZEN = 2;    // [2][2] frozen mask by iteration parity
constexpr int CTRL_NEWONES = 6;   // [2] of those, frames whose decision has a one
constexpr int CTRL_HARVEST = 8;   // [2]

__device__ __forceinline__ float &smem_f(int w) { return reinterpret_cast<float *>(nms_smem)[w]; }

// round x half-to-even to the quantiser step (== rint(x*qk)/qk of Main_Functions.py:483-492; exact for
// |x| < 2^21/qk): floats in [2^23/qk, 2^24/qk) are spaced exactly one step apart
__device__ __forceinline__ float qround(float x, float magic) { return __fsub_rn(__fadd_rn(x, magic), magic); }
// two values at once with Blackwell's packed fp32x2 adds (FADD2: one issue slot for both) -- same IEEE results
__device__ __forceinline__ float2 qround2(float2 x, float magic) {
    return __fadd2_rn(__fadd2_rn(x, make_float2(magic, magic)), make_float2(-magic, -magic));
}
// x * w for two values, each product rounded to float32 on its own.  NOT __fmul2_rn: ptxas (12.9) contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 -- it keeps the scalar forms apart, and neither -fmad=false nor an asm barrier
// stops it -- and Q(x * w) = rint(float32(x * w) * qk) / qk is a DOUBLE rounding: fused, 2.5 * 0.9f = 2.25000009 rounds to 2.5
// where the reference (float32 product 2.25, tie to even) gives 2.0.  One in 5e6 (level, weight) pairs differs, among them
// weights as plain as 0.9f, 0.85f and 0.95f (tests/test_gpu_parity.py: DOUBLE_ROUNDING_WEIGHTS).  An explicit fused
// multiply-add with a zero addend IS the rounded product (a -0 product becomes +0, which the quantiser's add erases
// anyway), still one instruction for both values, and nothing may fold a further add into it.
__device__ __forceinline__ float2 mul2_rn_unfused(float2 x, float2 w) {
    return __ffma2_rn(x, w, make_float2(0.0f, 0.0f));
}
__device__ __forceinline__ float qf(const KParams &P, float x) {   // full float quantiser
    return fminf(fmaxf(qround(x, P.qmagic), -P.qmax), P.qmax);
}

// weight of iteration t: staged copy in shared memory when it fits, else the global table
__device__ __forceinline__ float wload(const KParams &P, int idx) {
    return P.w_staged ? smem_f(P.off_w + idx) : __ldg(P.w_all + idx);
}
__device__ __forceinline__ float cn_weight(const KParams &P, int t, int i, int e) {
    if (P.sharing0 == 0) return 1.0f;
    return wload(P, P.w_off_cn + t * P.wc + (P.sharing0 == 3 ? 0 : (P.sharing0 == 2 ? i : e)));
}
__device__ __forceinline__ float ucn_weight(const KParams &P, int t, int i, int e) {
    return wload(P, P.w_off_ucn + t * P.wu + (P.sharing1 == 3 ? 0 : (P.sharing1 == 2 ? i : e)));
}
__device__ __forceinline__ float vn_weight(const KParams &P, int t, int j) {
    return wload(P, P.w_off_vn + t * P.wv + (P.sharing2 == 3 ? 0 : j));
}

struct Ctx {
    int lane, chunk, slot, warp;
    int q;        // lane index inside the interleaved block, 0..LP-1 (>= L: padding lane)
    int act;      // 1 for q < L
    int Lthr;     // L for active lanes, INT_MAX for padding lanes (they never wrap / rotate)
    int f0, f1;   // frame(s) of this lane's slot (packed: 2fp, 2fp+1; float: fp, fp); padding lanes: 0
    int fword, fsh;   // f0 / 32, f0 % 32: where this lane's frames sit in the per-CTA frame masks
    int a_lane;   // circulant lane of q
    long long frame0;
    int nvalid;
    int og;       // 1: every channel value of this warp's lanes is on the quantiser grid and inside +-qmax (batch in flight;
                  // only policies with TRACKS_GRID look at it)
};

// is frame f frozen as far as the VN phase of iteration t can tell?  (only used for the optional APP output)
__device__ __forceinline__ bool app_frozen(const KParams &P, int t, int f) {
    const uint32_t *misc = nms_smem + P.off_misc;
    bool frozen = (misc[MISC_CTRL + CTRL_FROZEN + ((t + 1) & 1) * 2 + (f >> 5)] >> (f & 31)) & 1u;
    if (P.early_term && t >= 1 && ((misc[MISC_SYND + (t & 1) * 2 + (f >> 5)] >> (f & 31)) & 1u) == 0u) frozen = true;
    return frozen;
}
__device__ __forceinline__ void app_store(const KParams &P, const Ctx &c, int j, int t, int f, float v) {
    if (!c.act || f >= c.nvalid || app_frozen(P, t, f)) return;
    const long long tt = P.app_all ? t : 0;
    P.app[tt * P.app_stride_t + (c.frame0 + f) * (long long)P.NZ + j * P.z + c.a_lane] = v;
}

// ------------------------------------------------------------------------- sample generation
// Tail refinement of a Box-Muller radius.  u1 = (r + 0.5) / 2^32 from one 32-bit word stops at 2^-33, i.e. at a radius
// of 6.76 sigma, and has only a few hundred distinct values beyond 5.8 sigma.  When the word is below 2^8 (once in
// 1.7e7 pairs) a second Philox block (same frame and quad, counter word c3 = 1 + pair) supplies 32 more bits:
// u1 = (r + (r' + 0.5) / 2^32) / 2^32 >= 2^-65, radius up to 9.5 sigma with a smooth tail.  Out of line: the hot loop
// pays one compare.
static __device__ __noinline__ float gen_tail_u1(unsigned long long F, int quad, int pair, uint32_t r, unsigned long long seed) {
    uint32_t x[4];
    philox4x32_10((uint32_t)F, (uint32_t)(F >> 32), (uint32_t)quad, 1u + (uint32_t)pair, (uint32_t)seed, (uint32_t)(seed >> 32), x);
    const float lowbits = fmaf((float)x[0], 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (r' + 0.5) / 2^32 in (0, 1)
    return ((float)r + lowbits) * 2.3283064365386963e-10f;
}

// four Philox words -> four N(0,1) via Box-Muller
__device__ __forceinline__ void box_muller4(const uint32_t (&r)[4], unsigned long long seed, unsigned long long F, int quad, float n[4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float u1 = fmaf((float)r[2 * h], 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (r+0.5)/2^32
        if (r[2 * h] < 256u) u1 = gen_tail_u1(F, quad, h, r[2 * h], seed);
        const float u2 = (float)r[2 * h + 1] * 2.3283064365386963e-10f;
        // hardware log2 / rsqrt / sin / cos (MUFU): ~1e-6 absolute on the normals, far below the quantiser step and
        // the Monte-Carlo noise; ldpc_llr_generate and the fused loops share this code, so they stay identical
        const float x = -2.0f * __logf(u1);
        const float rad = x > 0.0f ? x * rsqrtf(x) : 0.0f;
        float sn, cs;
        __sincosf(6.283185307179586f * u2, &sn, &cs);
        n[2 * h] = rad * cs;
        n[2 * h + 1] = rad * sn;
    }
}

// One Philox4x32-10 block -> four N(0,1) (frame F, bits 4*quad..4*quad+3)
__device__ __forceinline__ void gen_normal4(unsigned long long seed, unsigned long long F, int quad, float n[4]) {
    uint32_t r[4];
    philox4x32_10((uint32_t)F, (uint32_t)(F >> 32), (uint32_t)quad, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    box_muller4(r, seed, F, quad, n);
}
// the same block with the round keys taken from the kernel parameters
__device__ __forceinline__ void gen_normal4(const KParams &P, unsigned long long F, int quad, float n[4]) {
    uint32_t r[4];
    philox4x32_10_keyed((uint32_t)F, (uint32_t)(F >> 32), (uint32_t)quad, 0u, P.pkeys, r);
    box_muller4(r, P.seed, F, quad, n);
}

// N(0,1) sample -> channel LLR of a zero bit.  Print_Functions.py:45-50: x = n*sigma - 1 (all-zero word), llr = 2x/sigma^2
// (float64 there), quantised on the QMS path.  Here llr = fma(n, 2/sigma, -2/sigma^2): one rounding.
__device__ __forceinline__ float llr_from_normal(const KParams &P, float n) {
    const float llr = fmaf(n, P.two_over_s, -P.two_over_s2);
    return P.qms ? qf(P, llr) : llr;                                          // :49-50
}

// One Philox block -> four channel LLRs (bits 4*quad .. 4*quad+3 of frame F), punctured and shortened (:53-60)
__device__ __forceinline__ void gen_llr4(const KParams &P, unsigned long long F, int quad, float out[4]) {
    float n[4];
    gen_normal4(P, F, quad, n);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
        const int k = 4 * quad + k4 + 1;   // 1-based bit index
        float llr = llr_from_normal(P, n[k4]);
        if (P.punct_s > 0 && k >= P.punct_s && k <= P.punct_e) llr = P.sp ? 0.001f : 0.0f;    // :53-57
        if (P.short_s > 0 && k >= P.short_s && k <= P.short_e) llr = -P.clip; // :59-60
        out[k4] = llr;
    }
}

// global (Philox) index of batch frame k: consecutive from frame_offset, or listed (stage 2 of a two-stage Monte-Carlo run)
__device__ __forceinline__ unsigned long long frame_index(const KParams &P, long long k) {
    return P.frame_list != nullptr ? __ldg(P.frame_list + k) : P.frame_offset + (unsigned long long)k;
}

template <bool H2>
__device__ __forceinline__ void store_xa(const KParams &P, int f, int k, float v) {
    const int j = k / P.z, a = k - j * P.z;
    if (P.qms) v = fminf(fmaxf(v, -XA_BOUND), XA_BOUND);
    if (H2) smem_f(P.off_xa + (j * P.LP + a * P.Fp + (f >> 1)) * 2 + (f & 1)) = v;
    else smem_f(P.off_xa + j * P.LP + a * P.Fp + f) = __fadd_rn(v, 0.0f);   // float kernels: never -0.0 (nms_f32.cuh)
}
template <bool H2>
__device__ __forceinline__ void store_xa_ja(const KParams &P, int f, int j, int a, float v) {   // bit k = j*z + a
    if (P.qms) v = fminf(fmaxf(v, -XA_BOUND), XA_BOUND);
    if (H2) smem_f(P.off_xa + (j * P.LP + a * P.Fp + (f >> 1)) * 2 + (f & 1)) = v;
    else smem_f(P.off_xa + j * P.LP + a * P.Fp + f) = __fadd_rn(v, 0.0f);
}
template <bool H2>
__device__ __forceinline__ float load_xa(const KParams &P, int f, int k) {
    const int j = k / P.z, a = k - j * P.z;
    return H2 ? smem_f(P.off_xa + (j * P.LP + a * P.Fp + (f >> 1)) * 2 + (f & 1)) : smem_f(P.off_xa + j * P.LP + a * P.Fp + f);
}

// gather the packed hard decision of the frames in `mask` from the ballot array `hbuf`
template <bool H2>
__device__ __forceinline__ void copy_out(const KParams &P, const Ctx &c, const uint32_t mask[2], const uint32_t onesm[2],
                                         int hbuf) {
    const int nh = H2 ? 2 : 1;
    uint32_t *misc = nms_smem + P.off_misc;
    for (int item = threadIdx.x; item < P.FB * P.HW; item += blockDim.x) {
        const int f = item / P.HW, w = item - f * P.HW;
        if (!((mask[f >> 5] >> (f & 31)) & 1u)) continue;
        uint32_t outw = 0;
        // frames whose counted columns are clean still need the gather when parity columns are not counted
        if (P.target_n < P.N || ((onesm[f >> 5] >> (f & 31)) & 1u)) {
            const int half = H2 ? (f & 1) : 0, fp = H2 ? (f >> 1) : f;
            const uint32_t *hb = nms_smem + P.off_hb + (hbuf * nh + half) * P.N * P.C;
            int k = 32 * w;
            int j = k / P.z, a = k - j * P.z;
            for (int b = 0; b < 32 && k < P.NZ; ++b, ++k) {
                const int qq = a * P.Fp + fp;
                outw |= ((hb[j * P.C + (qq >> 5)] >> (qq & 31)) & 1u) << b;
                if (++a == P.z) { a = 0; ++j; }
            }
            const int nb = min(32, max(0, P.target_n * P.z - 32 * w));   // bits of this word that count
            const uint32_t cw = nb >= 32 ? outw : (outw & ((1u << nb) - 1u));
            if (cw) atomicAdd(&misc[MISC_BITERR + f], (uint32_t)__popc(cw));
        }
        if (P.hard != nullptr) P.hard[(c.frame0 + f) * (long long)P.HW + w] = outw;
    }
}

// OR this lane's per-frame bits (bit 0: frame f0, bit 16: frame f1 = f0 + 1) into the CTA's frame mask at `base`:
// one warp reduction + at most one shared atomic per warp and mask word
__device__ __forceinline__ void publish(const KParams &P, const Ctx &c, int base, uint32_t bits, bool h2) {
    uint32_t m = h2 ? ((bits & 1u) | ((bits >> 15) & 2u)) : (bits & 1u);
    m = c.act ? m << c.fsh : 0u;
    uint32_t *dst = nms_smem + P.off_misc + base;
    if (P.FB <= 32) {
        const uint32_t r = __reduce_or_sync(0xffffffffu, m);
        if (c.lane == 0 && r) atomicOr(dst, r);
    } else {
        const uint32_t r0 = __reduce_or_sync(0xffffffffu, c.fword == 0 ? m : 0u);
        const uint32_t r1 = __reduce_or_sync(0xffffffffu, c.fword == 1 ? m : 0u);
        if (c.lane == 0 && r0) atomicOr(dst, r0);
        if (c.lane == 0 && r1) atomicOr(dst + 1, r1);
    }
}

// per-frame decoding state of thread f < 64, packed into one register
constexpr uint32_t ST_FROZEN = 1u, ST_SYND_EVER = 2u, ST_EVER_CORRECT = 4u, ST_OUT_SYND_OK = 8u, ST_OUT_ONE = 16u;
constexpr int ST_ITERS_SH = 8, ST_EXEC_SH = 18;   // 10-bit fields: first zero-syndrome iteration, iterations executed
__device__ __forceinline__ uint32_t st_set(uint32_t st, int sh, int v) { return (st & ~(0x3ffu << sh)) | ((uint32_t)v << sh); }
__device__ __forceinline__ int st_get(uint32_t st, int sh) { return (int)((st >> sh) & 0x3ffu); }

// =============================================================================== main kernel
template <class Policy>
__device__ __forceinline__ void nms_decode_body(const KParams &P) {
    constexpr bool H2 = Policy::H2;
    Ctx c;
    const int tid = threadIdx.x;
    c.lane = tid & 31;
    c.warp = tid >> 5;
    c.chunk = c.warp % P.C;
    c.slot = c.warp / P.C;
    c.q = c.chunk * 32 + c.lane;
    c.act = c.q < P.L ? 1 : 0;
    c.Lthr = c.act ? P.L : 0x7fffffff;
    c.a_lane = c.act ? c.q / P.Fp : 0;
    {
        const int fp = c.act ? c.q - c.a_lane * P.Fp : 0;
        c.f0 = H2 ? 2 * fp : fp;
        c.f1 = H2 ? 2 * fp + 1 : fp;
        c.fword = c.f0 >> 5;
        c.fsh = c.f0 & 31;
    }
    uint32_t *misc = nms_smem + P.off_misc;
    uint32_t *ctrl = misc + MISC_CTRL;
    const long long nbatches = (P.n_frames + P.FB - 1) / P.FB;
    if (P.w_staged)
        for (int idx = tid; idx < P.w_words; idx += blockDim.x) smem_f(P.off_w + idx) = __ldg(P.w_all + idx);
    // padding lanes read their own (otherwise unused) words: give those defined contents once
    if (P.LP != P.L) {
        for (int idx = tid; idx < P.off_hb; idx += blockDim.x) nms_smem[idx] = 0u;
    }
    // float kernels: per-edge table of the syndrome pass {column * C | s_e*Fp << 16} (nms_f32.cuh)
    // followed by one word holding 1.0f (the "no weight" row of the specialised float kernels)
    if constexpr (!H2) {
        for (int e = tid; e < P.E; e += blockDim.x) nms_smem[P.off_et + e] = (uint32_t)(P.e_col[e] * P.C) | ((uint32_t)P.e_sF[e] << 16);
        if (tid == 0) nms_smem[P.off_et + P.E] = __float_as_uint(1.0f);
    }
    Policy::setup(P, tid);   // policy-owned tables (visible after the first barrier of the batch loop)

    for (long long batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
        c.frame0 = batch * P.FB;
        c.nvalid = (int)min((long long)P.FB, P.n_frames - c.frame0);
        __syncthreads();   // previous batch fully retired before shared memory is reused

        // ---------------- load channel LLRs into xa (zero for padding frames)
        bool fused = false;   // graph-specialised kernels load global LLRs inside their unrolled init pass
        const bool from_global = P.llr != nullptr || P.llr_q8 != nullptr;
        if constexpr (Policy::FUSED_LOAD) fused = from_global;
        if (fused) {
        } else if (from_global) {
            const int tot = P.FB * P.NZ;
            for (int idx = tid; idx < tot; idx += blockDim.x) {
                const int f = idx / P.NZ, k = idx - f * P.NZ;
                const long long g = (c.frame0 + f) * (long long)P.NZ + k;
                float v = 0.0f;
                if (f < c.nvalid) v = P.llr != nullptr ? __ldg(P.llr + g) : (float)__ldg(P.llr_q8 + g) * P.q8_step;
                store_xa<H2>(P, f, k, v);
            }
        } else {
            const int nquads = (P.NZ + 3) >> 2;
            int f = tid / nquads, quad = tid - f * nquads;   // (frame, quad) advance incrementally: no division per item
            while (f < P.FB) {
                float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (f < c.nvalid) gen_llr4(P, frame_index(P, c.frame0 + f), quad, v);
                int k = 4 * quad, j = k / P.z, a = k - j * P.z;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4, ++k) {
                    if (k < P.NZ) store_xa_ja<H2>(P, f, j, a, v[k4]);
                    if (++a == P.z) { a = 0; ++j; }
                }
                quad += blockDim.x;
                while (quad >= nquads) { quad -= nquads; ++f; }
            }
        }
        for (int idx = tid; idx < NMS_MISC_WORDS; idx += blockDim.x) misc[idx] = 0;
        // per-frame state lives in the registers of threads 0..63 (thread f owns frame f)
        uint32_t st = (tid >= c.nvalid ? ST_FROZEN : 0u) | ((uint32_t)P.T_run << ST_ITERS_SH) | ((uint32_t)P.T_run << ST_EXEC_SH);
        __syncthreads();
        if (tid < 64) {
            const uint32_t fm = __ballot_sync(0xffffffffu, (st & ST_FROZEN) != 0u);
            if (c.lane == 0) { ctrl[CTRL_FROZEN + c.warp] = fm; ctrl[CTRL_FROZEN + 2 + c.warp] = fm; }
        }

        // ---------------- init pass: xq, first V->C messages, hard bits of xin_0
        c.og = 0;
        uint32_t offgrid = 0;   // policies with TRACKS_GRID: OR of "this value has no exact grid form" over the thread's columns
        if constexpr (Policy::FUSED_LOAD) {
            if (fused) Policy::load_init(P, c, offgrid);
            else Policy::template vn_phase<true>(P, c, -1, true, offgrid);
        } else {
            Policy::template vn_phase<true>(P, c, -1, true, offgrid);
        }
        if constexpr (Policy::TRACKS_GRID) c.og = __all_sync(0xffffffffu, offgrid == 0u) ? 1 : 0;
        __syncthreads();

        bool alldone = false;
        int t = 0;
        for (; t < P.T_run; ++t) {
            // ======== CN phase (also yields the syndrome of the previous hard decision)
            uint32_t bad = 0;
            Policy::cn_phase(P, c, t, bad);
            if (t >= 1) publish(P, c, MISC_SYND + (t & 1) * 2, bad, H2);
            // A.  With early termination the barrier also tells whether ANY frame of the CTA still violates a check:
            // if none does, every running frame stops here and the VN phase of this iteration is not needed
            bool all_clean = false;
            if (P.early_term && t >= 1) all_clean = __syncthreads_or(c.act && (bad & (H2 ? LSB2 : 1u)) != 0u) == 0;
            else __syncthreads();
            // ======== per-frame bookkeeping for APP_{t-1} (threads 0..63), concurrent with the VN phase
            if (tid < 64) {
                bool newly = false;
                if (t >= 1) {
                    const bool fbad = (misc[MISC_SYND + (t & 1) * 2 + c.warp] >> c.lane) & 1u;
                    const bool one = (misc[MISC_ONES + ((t - 1) & 1) * 2 + c.warp] >> c.lane) & 1u;
                    if (!(st & ST_FROZEN)) {
                        if (!one) st |= ST_EVER_CORRECT;
                        if (!fbad && !(st & ST_SYND_EVER)) st = st_set(st, ST_ITERS_SH, t) | ST_SYND_EVER;
                        if (P.early_term && !fbad) {
                            newly = true;
                            st = st_set(st, ST_EXEC_SH, t) | ST_FROZEN | ST_OUT_SYND_OK | (one ? ST_OUT_ONE : 0u);
                        }
                    }
                }
                __syncwarp();
                if (c.lane == 0) {   // masks consumed: clear them for their next use
                    misc[MISC_ONES + ((t + 1) & 1) * 2 + c.warp] = 0u;
                    misc[MISC_SYND + ((t + 1) & 1) * 2 + c.warp] = 0u;
                }
                if (P.early_term) {
                    const uint32_t nm = __ballot_sync(0xffffffffu, newly);
                    const uint32_t no = __ballot_sync(0xffffffffu, newly && (st & ST_OUT_ONE));
                    const uint32_t fm = __ballot_sync(0xffffffffu, (st & ST_FROZEN) != 0u);
                    if (c.lane == 0) {
                        ctrl[CTRL_NEWLY + c.warp] = nm;
                        ctrl[CTRL_NEWONES + c.warp] = no;
                        ctrl[CTRL_FROZEN + (t & 1) * 2 + c.warp] = fm;
                    }
                }
            }
            if (all_clean) {
                __syncthreads();   // the bookkeeping's control words
                const uint32_t nm[2] = {ctrl[CTRL_NEWLY], ctrl[CTRL_NEWLY + 1]};
                const uint32_t no[2] = {ctrl[CTRL_NEWONES], ctrl[CTRL_NEWONES + 1]};
                if (nm[0] | nm[1]) copy_out<H2>(P, c, nm, no, (t + 1) & 1);   // hard bits of APP_{t-1}
                alldone = true;
                break;
            }
            // ======== VN phase (hard-decision ballots only when a copy-out can follow)
            uint32_t ones = 0;
            const bool need_hb = !H2 || P.early_term || t == P.T_run - 1;   // float kernels: always (their syndrome reads the ballots)
            Policy::template vn_phase<false>(P, c, t, need_hb, ones);
            publish(P, c, MISC_ONES + (t & 1) * 2, ones, H2);
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
            const int tl = P.T_run;
            const uint32_t bad = Policy::synd_phase(P, c, tl);
            publish(P, c, MISC_SYND + (tl & 1) * 2, bad, H2);
            __syncthreads();
            if (tid < 64) {
                const bool fbad = (misc[MISC_SYND + (tl & 1) * 2 + c.warp] >> c.lane) & 1u;
                const bool one = (misc[MISC_ONES + ((tl - 1) & 1) * 2 + c.warp] >> c.lane) & 1u;
                const bool pending = !(st & ST_FROZEN);
                if (pending) {
                    if (!one) st |= ST_EVER_CORRECT;
                    if (!fbad && !(st & ST_SYND_EVER)) st = st_set(st, ST_ITERS_SH, tl) | ST_SYND_EVER;
                    st = st_set(st, ST_EXEC_SH, tl) | (fbad ? 0u : ST_OUT_SYND_OK) | (one ? ST_OUT_ONE : 0u);
                }
                const uint32_t nm = __ballot_sync(0xffffffffu, pending);
                const uint32_t no = __ballot_sync(0xffffffffu, pending && one);
                if (c.lane == 0) { ctrl[CTRL_NEWLY + c.warp] = nm; ctrl[CTRL_NEWONES + c.warp] = no; }
            }
            __syncthreads();
            const uint32_t nm[2] = {ctrl[CTRL_NEWLY], ctrl[CTRL_NEWLY + 1]};
            const uint32_t no[2] = {ctrl[CTRL_NEWONES], ctrl[CTRL_NEWONES + 1]};
            copy_out<H2>(P, c, nm, no, (tl - 1) & 1);
        }
        __syncthreads();   // bit-error counts complete

        // ---------------- per-frame results, Monte-Carlo counters, harvest
        if (tid < 64) {
            bool valid = tid < c.nvalid;
            if (valid && P.defer_list != nullptr && !(st & ST_OUT_SYND_OK)) {   // stage 1: not converged -> stage 2, not counted here
                const unsigned slot_idx = atomicAdd(P.defer_count, 1u);
                if (slot_idx < P.defer_cap) P.defer_list[slot_idx] = frame_index(P, c.frame0 + tid);   // overflow: the host sees count > capacity
                valid = false;
            }
            const uint32_t be = misc[MISC_BITERR + tid];
            const bool uncor_any = !(st & ST_EVER_CORRECT), uncor_last = (st & ST_OUT_ONE) != 0u;
            const bool st_out_synd_ok = (st & ST_OUT_SYND_OK) != 0u;
            const int st_executed = st_get(st, ST_EXEC_SH);
            if (valid) {
                const long long F = c.frame0 + tid;
                if (P.iters) P.iters[F] = st_get(st, ST_ITERS_SH);
                if (P.flags)
                    P.flags[F] = (uint8_t)((st_out_synd_ok ? 1u : 0u) | (uncor_any ? 2u : 0u) | (uncor_last ? 4u : 0u) |
                                           ((st & ST_SYND_EVER) ? 8u : 0u));
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
            misc[MISC_HIDX + tid] = hidx;
            const uint32_t hm = __ballot_sync(0xffffffffu, hidx != 0xffffffffu);
            if (c.lane == 0) ctrl[CTRL_HARVEST + c.warp] = hm;
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
                    const uint32_t row = misc[MISC_HIDX + f];
                    for (int k = tid; k < P.NZ; k += blockDim.x)
                        P.uncor_buf[(size_t)row * P.NZ + k] = load_xa<H2>(P, f, k);
                }
            }
        }
    }
}

// compile-time loops
template <int V> struct IC { static constexpr int v = V; };
template <int I0, int I1, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I0 < I1) {
        f(IC<I0>{});
        static_for<I0 + 1, I1>(f);
    }
}

// REP_DESC(X) expands X(31) X(30) ... X(0)
#define NMS_REP_DESC(X)                                                                                        \
    X(31) X(30) X(29) X(28) X(27) X(26) X(25) X(24) X(23) X(22) X(21) X(20) X(19) X(18) X(17) X(16) X(15) X(14) \
    X(13) X(12) X(11) X(10) X(9) X(8) X(7) X(6) X(5) X(4) X(3) X(2) X(1) X(0)

}   // namespace nms
