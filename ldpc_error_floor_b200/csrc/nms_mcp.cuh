// nms_mcp.cuh -- persistent-slot Monte-Carlo kernel (packed fp16x2, graph-specialised).
//
// The batch kernels of nms_device.cuh give a CTA FB frames, decode them together and fetch the next FB when ALL of
// them have stopped: with per-frame early termination the slowest frame of a batch sets the pace, and at error-floor
// SNRs (average 2-6 iterations, a few per cent of stragglers that run all T) most lanes idle most of the time.  Here a
// CTA owns FB frame SLOTS instead.  Every loop step is one flooding iteration for whatever frames sit in the slots --
// each at its own iteration index -- and a slot whose frame has stopped (zero syndrome, or T iterations done) is
// refilled on the spot with the next frame of the CTA's share of the global frame index space:
//
//   step s:  [generate the channel values of the frames entering now]      (all threads, only if a slot was refilled)
//            VN phase  -- for a slot that was just refilled the stale C->V words are multiplied by 0 (the sum is an FMA
//                         chain S = cv * keep + S, the extrinsic output SX - cv an FMA cv * (-keep) + SX: no extra
//                         instruction per edge), which makes the phase the "init pass" of that frame
//            CN phase  -- weights of iteration t_f per half (the two frames of a lane need not be at the same iteration)
//            bookkeeping (warp 0, one lane per slot): syndrome / ones of the previous hard decision -> stop or go on,
//                         counters in registers, harvest, next frame index
//
// Samples: exactly those of ldpc_llr_generate / the batch kernels (Philox counter = global frame index), so the eight
// counters and the harvested words of a run do not depend on which kernel, launch geometry or number of GPUs produced
// them (tests/test_gpu_mc.py::test_persistent_kernel_counts_like_the_batch_kernel).
//
// Shared memory: msg[E][LP] (V->C / C->V in place) | xq[N][LP] (quantised channel values, half2) | weights | state.
// No float channel array and no hard-decision ballots: Monte-Carlo samples are on the quantiser grid already
// (xa == Q(xa)), and the all-zero codeword needs bit COUNTS, not bits -- z72 fits 2 CTAs per SM this way.
//
// Reference semantics restated: decode Main_Functions.py:161-335, samples Print_Functions.py:29-72, metrics :100-118,
// loop being replaced :136-161.
#pragma once
#include "nms_h2_spec.cuh"

namespace nms {

// state words, relative to P.off_misc; [2] = double-buffered by step parity; every slot mask is two words (slots 0-31, 32-63)
constexpr int MCP_SYND = 0;    // [2][2] slot mask: the previous hard decision violates a check
constexpr int MCP_GE1 = 4;     // [2][2] slots whose frame is at iteration >= 1 (its syndrome means something)
constexpr int MCP_ATT = 8;     // [2][2] slots whose frame has used up its iterations: it stops whatever the syndrome says
constexpr int MCP_ACT = 12;    // [2][2] slots that hold a frame
constexpr int MCP_HMASK = 16;  // [2] slots whose frame is copied to the harvest buffer in this step
constexpr int MCP_PROD = 20;   // frames of this CTA's share the producer warps have put into the ring so far
constexpr int MCP_CONS = 21;   // ... and how many of them the decoding warps have taken out (their ring rows are free)
constexpr int MCP_TP = 32;     // [2][32] per slot pair: iteration index of the low | high << 16 frame
constexpr int MCP_CNT = 96;    // [2][32] per slot pair: ones in the counted columns of the last hard decision, low | high << 16
constexpr int MCP_HROW = 160;  // [64] harvest row of the slot's frame
constexpr int MCP_ACC = 224;   // [FB][8] uint64: the eight Monte-Carlo counters per slot, flushed once per launch
static_assert(MCP_ACC == NMS_MCP_MISC_WORDS(0), "host and device agree on the state block");

// slot masks: one word up to 32 frames per CTA, two beyond (z = 1 graphs interleave 64 frames per CTA)
template <bool WIDE> struct McpMask { using type = uint32_t; };
template <> struct McpMask<true> { using type = unsigned long long; };
__device__ __forceinline__ int mpopc(uint32_t m) { return __popc(m); }
__device__ __forceinline__ int mpopc(unsigned long long m) { return __popcll(m); }
__device__ __forceinline__ int mffs(uint32_t m) { return __ffs(m); }
__device__ __forceinline__ int mffs(unsigned long long m) { return __ffsll((long long)m); }
__device__ __forceinline__ uint32_t mdrop_top(uint32_t m) { return m & ~(0x80000000u >> __clz(m)); }
__device__ __forceinline__ unsigned long long mdrop_top(unsigned long long m) { return m & ~(0x8000000000000000ull >> __clzll((long long)m)); }
__device__ __forceinline__ void mload(const uint32_t *w, uint32_t &m) { m = w[0]; }
__device__ __forceinline__ void mload(const uint32_t *w, unsigned long long &m) { m = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32); }

__device__ __forceinline__ void sts16(uint32_t a, unsigned short v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v)); }
__device__ __forceinline__ void sts64u(uint32_t a, uint32_t lo, uint32_t hi) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(lo), "r"(hi)); }
__device__ __forceinline__ unsigned short lds16(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t ld_flag(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_flag(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
// barrier among the first `n` threads of the CTA (the decoding warps) / among the producer warps: named barriers 1 and 2
template <int N> __device__ __forceinline__ void bar_named(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory"); }

template <class G>
struct McpKernel {
    static constexpr int FB = 2 * G::Fp;
    static constexpr int NTHR = G::C * G::R * 32;      // decoding threads (threads NTHR .. NTHR + 32 NP - 1 are the producers)
    static constexpr int NP = NMS_MCP_NP;
    static constexpr int RS = ((G::N * G::z + 3) & ~3);  // halves per ring row
    static __device__ __forceinline__ void cbar() {     // barrier of the decoding warps
        if constexpr (NP > 0) bar_named<NTHR>(1);
        else __syncthreads();
    }
    static constexpr uint32_t LP4 = G::LP * 4u;
    static constexpr bool PAD = G::L != G::LP;
    static constexpr int NZ = G::N * G::z;
    static constexpr int NQUADS = (NZ + 3) / 4;
    static constexpr bool WIDE = FB > 32;
    using mask_t = typename McpMask<WIDE>::type;
    static constexpr mask_t FBMASK = FB >= (WIDE ? 64 : 32) ? ~(mask_t)0 : (((mask_t)1 << (FB & (WIDE ? 63 : 31))) - 1);
    static constexpr int NBW = (FB + 31) / 32;          // bookkeeping warps: one lane per slot
    static_assert(FB <= 64 && G::C * G::R >= NBW, "slot masks are at most two words, one bookkeeping warp per word");

    // ---------------------------------------------------------------------------------------- sample generation
    // channel values of the frames entering the slots in `fresh` (rank r in the mask -> global frame first + r), written
    // as halves into the xq array; same samples, same order as gen_llr4 draws them for everyone else
    static __device__ __forceinline__ void generate(const KParams &P, uint32_t sb, mask_t fresh, unsigned long long first) {
        const unsigned short hneg = __half_as_ushort(__float2half_rn(-P.qmax));   // a shortened bit: Q(-clip_LLR) to the decoder
        unsigned long long F = P.frame_offset + first;
        for (mask_t m = fresh; m != 0; m &= m - 1, ++F) {                          // uniform: one refilled slot after the other
            const int f = mffs(m) - 1;
            const uint32_t base = sb + (uint32_t)P.off_xq * 4u + (uint32_t)(f >> 1) * 4u + (uint32_t)(f & 1) * 2u;
            for (int quad = threadIdx.x; quad < NQUADS; quad += NTHR) {
                const int k1 = 4 * quad + 1;                                       // 1-based index of the quad's first bit
                if constexpr (G::z % 4 == 0) {
                    // the four bits of a quad sit in one column at consecutive circulant lanes: one address, immediate offsets
                    constexpr int QPC = G::z / 4;
                    const int j0 = quad / QPC, a0 = (quad - j0 * QPC) * 4;
                    const uint32_t ad = base + (uint32_t)(j0 * G::LP + a0 * G::Fp) * 4u;
                    const bool inp = P.punct_s > 0 && k1 + 3 >= P.punct_s && k1 <= P.punct_e;   // touches the punctured range
                    const bool ins = P.short_s > 0 && k1 + 3 >= P.short_s && k1 <= P.short_e;   // ... the shortened range
                    if (!(inp || ins)) {                                           // the common case: four plain samples
                        float n[4];
                        gen_normal4(P, F, quad, n);
                        const __half2 lo = __floats2half2_rn(llr_from_normal(P, n[0]), llr_from_normal(P, n[1]));
                        const __half2 hi = __floats2half2_rn(llr_from_normal(P, n[2]), llr_from_normal(P, n[3]));
                        sts16(ad, (unsigned short)(h2u(lo) & 0xffffu));
                        sts16(ad + (uint32_t)G::Fp * 4u, (unsigned short)(h2u(lo) >> 16));
                        sts16(ad + (uint32_t)G::Fp * 8u, (unsigned short)(h2u(hi) & 0xffffu));
                        sts16(ad + (uint32_t)G::Fp * 12u, (unsigned short)(h2u(hi) >> 16));
                        continue;
                    }
                    const bool allp = P.punct_s > 0 && k1 >= P.punct_s && k1 + 3 <= P.punct_e;
                    const bool alls = P.short_s > 0 && k1 >= P.short_s && k1 + 3 <= P.short_e;
                    if (allp || alls) {                                            // no sample needed: the value is fixed
                        const unsigned short c = alls ? hneg : (unsigned short)0;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) sts16(ad + (uint32_t)(k4 * G::Fp) * 4u, c);
                        continue;
                    }
                }
                // a quad that straddles a range boundary or a column boundary: the general code
                float v[4];
                gen_llr4(P, F, quad, v);
                int k = k1 - 1, j = k / G::z, a = k - j * G::z;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4, ++k) {
                    const float x = fminf(fmaxf(v[k4], -P.qmax), P.qmax);          // shortened: -clip_LLR -> -qmax
                    if (NZ % 4 == 0 || k < NZ) sts16(base + (uint32_t)(j * G::LP + a * G::Fp) * 4u, __half_as_ushort(__float2half_rn(x)));
                    if (++a == G::z) { a = 0; ++j; }
                }
            }
        }
    }

    // ------------------------------------------------------------------------------------------ producer warps
    // The sample generator as its own warps: frame i of the CTA's share (global frame lo + i) goes into ring row i % FB as
    // NZ halves in bit order, as soon as that row is free; MCP_PROD / MCP_CONS are the two ends of the ring.  Same samples
    // as generate() (and everyone else): gen_normal4 of (global frame, quad).
    static __device__ __forceinline__ void produce(const KParams &P, uint32_t sb, unsigned long long lo, uint32_t total) {
        uint32_t *misc = nms_smem + P.off_misc;
        const int ptid = threadIdx.x - NTHR;
        const unsigned short hneg = __half_as_ushort(__float2half_rn(-P.qmax));
        const uint32_t hneg2 = (uint32_t)hneg | ((uint32_t)hneg << 16);
        for (uint32_t i = 0; i < total; ++i) {
            while ((int)(i - ld_flag(&misc[MCP_CONS])) >= FB) __nanosleep(100);       // ring full: the decoders are busy
            const unsigned long long F = P.frame_offset + lo + i;
            const uint32_t row = sb + (uint32_t)P.off_ring * 4u + (uint32_t)(i % FB) * (uint32_t)(RS * 2);
            for (int quad = ptid; quad < NQUADS; quad += NP * 32) {
                const int k1 = 4 * quad + 1;                                           // 1-based index of the quad's first bit
                const bool inp = P.punct_s > 0 && k1 + 3 >= P.punct_s && k1 <= P.punct_e;
                const bool ins = P.short_s > 0 && k1 + 3 >= P.short_s && k1 <= P.short_e;
                uint32_t w0, w1;
                if (!(inp || ins)) {                                                   // four plain samples
                    float n[4];
                    gen_normal4(P, F, quad, n);
                    w0 = h2u(__floats2half2_rn(llr_from_normal(P, n[0]), llr_from_normal(P, n[1])));
                    w1 = h2u(__floats2half2_rn(llr_from_normal(P, n[2]), llr_from_normal(P, n[3])));
                } else if (P.punct_s > 0 && k1 >= P.punct_s && k1 + 3 <= P.punct_e) {
                    w0 = w1 = 0u;                                                      // punctured: no sample needed
                } else if (P.short_s > 0 && k1 >= P.short_s && k1 + 3 <= P.short_e) {
                    w0 = w1 = hneg2;                                                   // shortened: Q(-clip_LLR)
                } else {                                                               // a quad across a range boundary
                    float v[4];
                    gen_llr4(P, F, quad, v);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) v[k4] = fminf(fmaxf(v[k4], -P.qmax), P.qmax);
                    w0 = h2u(__floats2half2_rn(v[0], v[1]));
                    w1 = h2u(__floats2half2_rn(v[2], v[3]));
                }
                sts64u(row + (uint32_t)quad * 8u, w0, w1);
            }
            __threadfence_block();
            if constexpr (NP > 1) bar_named<NP * 32>(2);
            else __syncwarp();
            if (ptid == 0) st_flag(&misc[MCP_PROD], i + 1u);
        }
    }

    // decoding warps: move the frames entering the slots in `fresh` (rank r -> frame `rel` + r of the CTA's share) from the
    // ring into the xq array; waits for the producers if they are behind
    static __device__ __forceinline__ void take(const KParams &P, uint32_t sb, mask_t fresh, uint32_t rel) {
        uint32_t *misc = nms_smem + P.off_misc;
        const uint32_t need = rel + (uint32_t)mpopc(fresh);
        while (ld_flag(&misc[MCP_PROD]) < need) { }
        __threadfence_block();
        uint32_t i = rel;
        for (mask_t m = fresh; m != 0; m &= m - 1, ++i) {
            const int f = mffs(m) - 1;
            const uint32_t row = sb + (uint32_t)P.off_ring * 4u + (i % FB) * (uint32_t)(RS * 2);
            const uint32_t base = sb + (uint32_t)P.off_xq * 4u + (uint32_t)(f >> 1) * 4u + (uint32_t)(f & 1) * 2u;
            for (int k = threadIdx.x; k < NZ; k += NTHR) {
                const int j = k / G::z, a = k - j * G::z;
                sts16(base + (uint32_t)(j * G::LP + a * G::Fp) * 4u, lds16(row + (uint32_t)k * 2u));
            }
        }
    }

    // ------------------------------------------------------------------------------------------------ VN phase
    // keep / negkeep: {1, 1} / {-1, -1} as half2, 0 / -0 in the half whose slot was just refilled (or is empty).
    // freshsel: 0xffff per refilled half.  wv_lo / wv_hi: byte address of the VN weight row of the iteration each half enters.
    // sh_lo / sh_hi: this lane's bit j*z + a is shortened iff sh_lo <= j*z <= sh_hi (VN weights see the SAMPLE there, -clip_LLR,
    // not Q(-clip_LLR) = -qmax: Print_Functions.py:59-60 vs Main_Functions.py:168-177)
    // MV: VN weights vary per column (sharing code 2); else wsl / wsh are THE weights of the two halves' iterations
    template <int J, bool VNW, bool MV>
    static __device__ __forceinline__ void vn_col(const KParams &P, const H2Ctx &h, uint32_t keep, uint32_t negkeep,
                                                  uint32_t freshsel, uint32_t wv_lo, uint32_t wv_hi, float wsl, float wsh,
                                                  int sh_lo, int sh_hi, uint32_t &cnt) {
        constexpr int C0 = G::col_ptr[J], DV = G::col_ptr[J + 1] - C0;
        uint32_t addr[DV], cv[DV];
        static_for<0, DV>([&](auto u) {
            constexpr int U = decltype(u)::v;
            constexpr uint32_t X4 = (uint32_t)G::vn_e[C0 + U] * LP4;
            constexpr int ROT = G::vn_rot[C0 + U];
            addr[U] = h.sb + spec_rot<G, ROT>(h) + X4;
            cv[U] = lds32(addr[U]);
        });
        // stale words of a refilled half are finite on-grid values, so their sum times 0 is 0: one multiply after the tree
        const __half2 S = __hmul2(h2_tree_sum<DV, 0, DV>(cv), u2h(keep));
        const __half2 xqh = u2h(lds32(h.xq4 + (uint32_t)(J * G::LP) * 4u));
        __half2 xin = xqh;
        uint32_t hs = h2u(__hadd2(xqh, S));                                   // unclipped APP (a refilled half: xq itself)
        if constexpr (VNW) {
            // per-column rows (sharing code 2) are fetched here, a per-iteration scalar (code 3) came in with the phase
            const float wl = MV ? h2_w(wv_lo, J, -1) : wsl, wh = MV ? h2_w(wv_hi, J, -1) : wsh;
            const bool shortened = J * G::z >= sh_lo && J * G::z <= sh_hi;
            const float xl = shortened ? -P.clip : __low2float(xqh), xh = shortened ? -P.clip : __high2float(xqh);
            xin = q2(P, __fmul_rn(xl, wl), __fmul_rn(xh, wh));               // Q(xa * w), :168-177
            hs = (h2u(xin) & freshsel) | (hs & ~freshsel);                    // iteration 0 takes the syndrome of xin_0 (:181-182)
        }
        const uint32_t hbw = ~(hs >> 15) & LSB2;                              // bit = (value >= 0)
        if (J < P.target_n) cnt += hbw;                                       // uniform; 16-bit counters, no carry (<= N per lane)
        const __half2 SX = __hadd2(xin, S);
#pragma unroll
        for (int u = 0; u < DV; ++u) sts32(addr[u], h2u(__hfma2(u2h(cv[u]), u2h(negkeep), SX)) | hbw);   // total - self
    }

    template <int SLOT, bool VNW, bool MV>
    static __device__ __forceinline__ void vn_slot(const KParams &P, const H2Ctx &h, uint32_t keep, uint32_t negkeep,
                                                   uint32_t freshsel, uint32_t wv_lo, uint32_t wv_hi, float wsl, float wsh,
                                                   int sh_lo, int sh_hi, uint32_t &cnt) {
        constexpr int NT = (G::N - SLOT + G::R - 1) / G::R;
        static_for<0, NT>([&](auto n) {
            constexpr int J = G::vn_order[SLOT + decltype(n)::v * G::R];
            vn_col<J, VNW, MV>(P, h, keep, negkeep, freshsel, wv_lo, wv_hi, wsl, wsh, sh_lo, sh_hi, cnt);
        });
    }

    // ------------------------------------------------------------------------------------------------ CN phase
    static __device__ __forceinline__ void cn_phase(const KParams &P, const H2Ctx &h, int slot, int t_lo, int t_hi, uint32_t &bad) {
        spec_cn_phase<G, true>(P, h, slot, h2_wrow(h, P.h2w_c, t_lo, P.h2_wc), h2_wrow(h, P.h2w_c, t_hi, P.h2_wc),
                               h2_wrow(h, P.h2w_u, t_lo, P.h2_wu), h2_wrow(h, P.h2w_u, t_hi, P.h2_wu), bad);
    }

    // ---------------------------------------------------------------------------------------------------- kernel
    static __device__ __forceinline__ void run(const KParams &P) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int chunk = warp % G::C, slot = warp / G::C;
        const int q = chunk * 32 + lane;
        const bool act = q < G::L;
        const int a_lane = act ? q / G::Fp : 0;
        const int fp = act ? q - a_lane * G::Fp : 0;
        H2Ctx h;
        h.sb = smem_base();
        h.q4 = (uint32_t)q * 4u;
        h.amask = act ? 0xffffffffu : 0u;
        h.Lthr4 = act ? (uint32_t)G::L * 4u : 0x40000000u;
        h.xa8 = 0u;
        h.xq4 = h.sb + (uint32_t)P.off_xq * 4u + h.q4;
        uint32_t *misc = nms_smem + P.off_misc;
        const int T = P.T_run;

        // this CTA's share of the launch's frames
        const unsigned long long nfr = (unsigned long long)P.n_frames;
        const unsigned long long lo = nfr * blockIdx.x / gridDim.x, hi = nfr * (blockIdx.x + 1ull) / gridDim.x;

        constexpr int NALL = NTHR + NP * 32;
        for (int idx = tid; idx < P.w_words; idx += NALL) smem_f(P.off_w + idx) = __ldg(P.w_all + idx);
        for (int idx = tid; idx < P.off_w; idx += NALL) nms_smem[idx] = 0u;          // messages, channel values, ring: finite
        for (int idx = tid; idx < NMS_MCP_MISC_WORDS(FB); idx += NALL) misc[idx] = 0u;
        unsigned long long next = lo;                                                  // first frame not yet in a slot
        mask_t fresh = (hi - lo) >= (unsigned long long)FB ? FBMASK : (((mask_t)1 << (int)(hi - lo)) - 1);
        unsigned long long first = next;                                               // frame of the lowest refilled slot
        next += (unsigned long long)mpopc(fresh);
        mask_t empty = ~fresh & FBMASK;
        __syncthreads();                                  // the only barrier all warps share
        if constexpr (NP > 0) {
            if (tid >= NTHR) {                            // producer warps: nothing but samples from here on
                produce(P, h.sb, lo, (uint32_t)(hi - lo));
                return;
            }
        }
        if (tid == 0) {
            misc[MCP_ACT] = (uint32_t)fresh;
            if constexpr (WIDE) misc[MCP_ACT + 1] = (uint32_t)((unsigned long long)fresh >> 32);
        }
        // bookkeeping state of slot `sl` (warps 0 .. NBW-1, one lane per slot): iteration index of its frame, "was right at
        // some iteration", its counters
        const int sl = warp * 32 + lane;
        int t = 0;
        bool ever = false;
        unsigned long long *acc = reinterpret_cast<unsigned long long *>(misc + MCP_ACC) + (sl < FB ? sl : 0) * 8;
        if (fresh == 0) return;

        for (int s = 0;; ++s) {
            const int p = s & 1;
            if (fresh != 0) {
                if constexpr (NP > 0) {
                    take(P, h.sb, fresh, (uint32_t)(first - lo));
                    cbar();
                    if (tid == 0) {                       // those ring rows are free (every copy out of them is behind the barrier)
                        __threadfence_block();
                        st_flag(&misc[MCP_CONS], (uint32_t)(first - lo) + (uint32_t)mpopc(fresh));
                    }
                } else {
                    generate(P, h.sb, fresh, first);
                    cbar();
                }
            }
            // ================================================================ VN phase
            {
                const uint32_t kb = (uint32_t)(~(fresh | empty) >> (2 * fp));          // bit 0 / 1: low / high frame goes on
                const uint32_t keep = ((kb & 1u) ? 0x3c00u : 0u) | ((kb & 2u) ? 0x3c000000u : 0u);
                const uint32_t freshsel = ((kb & 1u) ? 0u : 0xffffu) | ((kb & 2u) ? 0u : 0xffff0000u);
                uint32_t cnt = 0;
                if (P.sharing2 != 0) {
                    const uint32_t tp = misc[MCP_TP + (p ^ 1) * 32 + fp];              // what the last CN phase ran
                    const int tl = (kb & 1u) ? min((int)(tp & 0xffffu) + 1, T - 1) : 0;
                    const int th = (kb & 2u) ? min((int)(tp >> 16) + 1, T - 1) : 0;
                    const uint32_t wv_lo = h2_wrow(h, P.h2w_v, tl, P.h2_wv), wv_hi = h2_wrow(h, P.h2w_v, th, P.h2_wv);
                    // shortened bits are k = j*z + a with short_s <= k + 1 <= short_e
                    const int sh_lo = P.short_s > 0 ? P.short_s - 1 - a_lane : 0x7fffffff, sh_hi = P.short_e - 1 - a_lane;
                    const float wsl = ldsf(wv_lo), wsh = ldsf(wv_hi);
                    static_for<0, G::R>([&](auto sl_) {
                        if (slot != decltype(sl_)::v) return;
                        if (P.h2_mv != 0) vn_slot<decltype(sl_)::v, true, true>(P, h, keep, keep ^ SIGN2, freshsel, wv_lo, wv_hi, wsl, wsh, sh_lo, sh_hi, cnt);
                        else vn_slot<decltype(sl_)::v, true, false>(P, h, keep, keep ^ SIGN2, freshsel, wv_lo, wv_hi, wsl, wsh, sh_lo, sh_hi, cnt);
                    });
                } else {
                    static_for<0, G::R>([&](auto sl_) {
                        if (slot == decltype(sl_)::v) vn_slot<decltype(sl_)::v, false, false>(P, h, keep, keep ^ SIGN2, freshsel, 0u, 0u, 1.0f, 1.0f, 0, 0, cnt);
                    });
                }
                // ones of this hard decision, per slot pair (skipped by warps that saw none: the common case once
                // the channel errors are gone)
                if (!act) cnt = 0u;
                if (__any_sync(0xffffffffu, cnt != 0u)) {
                    if constexpr (G::Fp >= 16) {           // (nearly) every lane of a warp serves another pair: no reduction to do
                        if (cnt) atomicAdd(&misc[MCP_CNT + p * 32 + fp], cnt);
                    } else {
#pragma unroll
                        for (int k = 0; k < G::Fp; ++k) {
                            const uint32_t r = __reduce_add_sync(0xffffffffu, fp == k ? cnt : 0u);
                            if (lane == 0 && r) atomicAdd(&misc[MCP_CNT + p * 32 + k], r);
                        }
                    }
                }
            }
            cbar();
            // ================================================================ CN phase
            {
                const uint32_t tp = misc[MCP_TP + p * 32 + fp];
                uint32_t bad = 0;
                cn_phase(P, h, slot, min((int)(tp & 0xffffu), T - 1), min((int)(tp >> 16), T - 1), bad);
                const uint32_t m2 = act ? ((bad & 1u) | ((bad >> 15) & 2u)) : 0u;      // this lane's two frames
                const int sh = 2 * fp;
                const uint32_t r0 = __reduce_or_sync(0xffffffffu, sh < 32 ? m2 << sh : 0u);
                if (lane == 0 && r0) atomicOr(&misc[MCP_SYND + p * 2], r0);
                if constexpr (WIDE) {
                    const uint32_t r1 = __reduce_or_sync(0xffffffffu, sh >= 32 ? m2 << (sh - 32) : 0u);
                    if (lane == 0 && r1) atomicOr(&misc[MCP_SYND + p * 2 + 1], r1);
                }
            }
            cbar();
            // ================================================================ who stops, who enters
            mask_t sbad, actm, ge1m, attm;
            mload(misc + MCP_SYND + p * 2, sbad); mload(misc + MCP_ACT + p * 2, actm);
            mload(misc + MCP_GE1 + p * 2, ge1m); mload(misc + MCP_ATT + p * 2, attm);
            const mask_t fin = actm & (((P.early_term ? ~sbad : (mask_t)0) & ge1m) | attm);
            const int nfin = mpopc(fin);
            const unsigned long long avail = hi - next;
            const int nnew = avail >= (unsigned long long)nfin ? nfin : (int)avail;
            mask_t enter = fin;                                                        // the lowest nnew of the stopped slots
            for (int k = nfin; k > nnew; --k) enter = mdrop_top(enter);
            const mask_t actn = (actm & ~fin) | enter;
            if (warp < NBW) {
                const bool mine = sl < FB && ((actm >> sl) & 1);
                const bool stop = mine && ((fin >> sl) & 1);
                const uint32_t cw = misc[MCP_CNT + p * 32 + (sl >> 1 & 31)];
                const uint32_t ones = (sl & 1) ? cw >> 16 : cw & 0xffffu;              // of APP_{t-1}
                const bool fbad = (sbad >> sl) & 1;
                if (mine && t >= 1 && ones == 0u) ever = true;                         // D9: right at some iteration
                uint32_t hidx = 0xffffffffu;
                if (stop) {
                    const bool synd_ok = !fbad, one = ones != 0u;
                    acc[0] += 1; acc[1] += one; acc[2] += !ever; acc[3] += ones; acc[4] += (unsigned)t;
                    acc[5] += !synd_ok; acc[6] += synd_ok && one;
                    const bool harvest = P.harvest_mode == 0 ? false
                                         : (P.harvest_mode == 1 ? !ever : (P.harvest_mode == 2 ? one : !synd_ok));
                    if (harvest) {
                        acc[7] += 1;
                        if (P.uncor_count != nullptr) {
                            hidx = atomicAdd(P.uncor_count, 1u);
                            if (hidx >= P.uncor_cap || P.uncor_buf == nullptr) hidx = 0xffffffffu;
                        }
                    }
                    t = 0;
                    ever = false;
                } else if (mine) {
                    ++t;
                }
                __syncwarp();
                const bool on = sl < FB && ((actn >> sl) & 1);
                const uint32_t ge1 = __ballot_sync(0xffffffffu, on && t >= 1);
                const uint32_t att = __ballot_sync(0xffffffffu, on && t >= T);
                const uint32_t hm = __ballot_sync(0xffffffffu, hidx != 0xffffffffu);
                const int tn = __shfl_down_sync(0xffffffffu, t, 1);
                if (sl < FB) misc[MCP_HROW + sl] = hidx;
                if (sl < FB && !(sl & 1)) {
                    misc[MCP_TP + (p ^ 1) * 32 + (sl >> 1)] = (uint32_t)t | ((uint32_t)tn << 16);
                    misc[MCP_CNT + p * 32 + (sl >> 1)] = 0u;                           // both lanes of the pair have read it
                }
                if (lane == 0) {                                                       // this warp's word of every mask
                    misc[MCP_GE1 + (p ^ 1) * 2 + warp] = ge1;
                    misc[MCP_ATT + (p ^ 1) * 2 + warp] = att;
                    misc[MCP_ACT + (p ^ 1) * 2 + warp] = (uint32_t)(actn >> (32 * warp));
                    misc[MCP_SYND + (p ^ 1) * 2 + warp] = 0u;                          // consumed one step ago
                    misc[MCP_HMASK + warp] = hm;
                }
            }
            // harvest: the stopped frames' channel values leave before the generator overwrites them
            if (P.harvest_mode != 0 && fin != 0 && P.uncor_buf != nullptr) {
                cbar();
                mask_t hm;
                mload(misc + MCP_HMASK, hm);
                if (hm != 0) {
                    for (int f = 0; f < FB; ++f) {
                        if (!((hm >> f) & 1)) continue;
                        float *row = P.uncor_buf + (size_t)misc[MCP_HROW + f] * NZ;
                        const __half *xq = reinterpret_cast<const __half *>(nms_smem + P.off_xq) + (f & 1);
                        for (int k = tid; k < NZ; k += NTHR) {
                            const int j = k / G::z, a = k - j * G::z;
                            float v = __half2float(xq[(j * G::LP + a * G::Fp + (f >> 1)) * 2]);
                            if (P.short_s > 0 && k + 1 >= P.short_s && k + 1 <= P.short_e) v = -P.clip;   // as generated
                            row[k] = v;
                        }
                    }
                    cbar();
                }
            }
            empty = (empty | fin) & ~enter;
            first = next;
            next += (unsigned long long)nnew;
            fresh = enter;
            if (actn == 0) break;
        }

        if (warp < NBW && P.counters != nullptr) {
            for (int k = 0; k < 8; ++k) {
                unsigned long long v = sl < FB ? acc[k] : 0ull;
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (lane == 0 && v) atomicAdd(P.counters + k, v);
            }
        }
    }
};

}   // namespace nms
