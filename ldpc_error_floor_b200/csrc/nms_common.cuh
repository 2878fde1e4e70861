// nms_common.cuh -- shared declarations of the sm_100a NMS decode kernels and their launcher.
//
// Layout idea (DESIGN.md "Data layout"): one CTA decodes FB frames at once.  A "slot" packs
// either two frames as fp16x2 (packed kernel, exact for the quantised min-sum grids) or one
// frame as fp32 (float kernel).  Fp slots are interleaved lane-wise, q = a*Fp + fp with a the
// circulant lane, so that the block sees ONE quasi-cyclic code of lifting size L = z*Fp whose
// shift for proto edge e is s_e*Fp: a rotation by s is q -> (q + s*Fp) mod L for every
// interleaved frame at once, and consecutive threads touch consecutive shared-memory words on
// both sides of every rotation (conflict-free up to the single wrap point).
#pragma once
#include <cuda_fp16.h>
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#else
// run-time specialisation (nms_jit.cu) compiles the kernel headers with NVRTC: no host headers there
typedef unsigned char uint8_t;
typedef signed char int8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#endif

#define LDPC_MAX_M 256
#define LDPC_MAX_N 256
#define LDPC_MAX_E 1024
#define LDPC_MAX_FB 64   // frames per CTA
#define LDPC_MAX_T 256
#define NMS_MISC_WORDS 400        // per-CTA bookkeeping words (nms_device.cuh)
#define NMS_WSTAGE_MAX_WORDS 8192 // weights are staged in shared memory when they fit in this many floats

// tables + scalars, passed as ONE __grid_constant__ kernel parameter (constant bank, LDC/ULDC)
struct KParams {
    // geometry
    int M, N, E, z, NZ;
    int Fp;   // slots interleaved per CTA
    int FB;   // frames per CTA (2*Fp packed, Fp float)
    int L;    // z*Fp  (active lanes)
    int LP;   // L rounded up to 32
    int C;    // LP/32 lane chunks
    int R;    // task slots (warps per chunk); CTA = C*R warps
    // arithmetic
    int qms;           // 1: quantised min-sum
    int sp;            // 1: sum-product check update (decoding_type 0, Main_Functions.py:238-245), float kernels only
    float qmagic;      // 1.5*2^23/qk: (x + qmagic) - qmagic rounds x half-to-even to the quantiser step 1/qk
    float qmax;        // Q(x) = clamp(round_to_step(x), +-qmax)            (Main_Functions.py:483-492)
    uint32_t qmax_h2;  // {qmax, qmax} as half2 bits: the packed kernels' clamp operand, converted once on the host
    float clip;        // clip_LLR (main_Base.py:69)
    // message saturation without a mode branch: sat(x) = clamp((x + sat_magic) - sat_magic, +-sat_bound) is Q(x) on the
    // quantised path (qmagic, qmax) and clip(x, +-clip_LLR) on the float path (0, clip)
    float sat_magic, sat_bound;
    int sharing0, sharing1, sharing2;
    int wc, wu, wv;    // weight row widths
    const float *w_all;   // device [T*wc | T*wu | T*wv]
    int w_staged;      // 1: the three blocks are copied to shared memory at off_w (same layout)
    int w_off_cn, w_off_ucn, w_off_vn, w_words;   // float offsets inside w_all / the staged copy
    // packed kernels: weights are ALWAYS staged and indexed branch-free:
    //   w = smem_f(h2w_X + t*h2_wX + (node & h2_mX)), X in {c (CN), u (UCN), v (VN)};
    // "no weight" is a staged row of 1.0f, "no UCN weight" aliases the CN block.
    int h2w_c, h2w_u, h2w_v, h2_wc, h2_wu, h2_wv, h2_mc, h2_mu, h2_mv;
    int T_run, early_term;
    int no_xq;         // packed kernels with VN weights: no xq array, Q(xa) recomputed from xa
    int target_n;      // proto columns that count in the error metrics (systematic: N - M, main_Base.py:83-86; else N)
    // source: llr != nullptr -> global float32 LLRs; llr_q8 != nullptr -> global int8 LLRs in units of q8_step
    // (the compact form of on-grid words); else Philox generator
    const float *llr;
    const signed char *llr_q8;
    float q8_step;
    long long n_frames;
    float sigma, two_over_s2, two_over_s;   // channel: llr = (2/sigma) * n - 2/sigma^2
    unsigned long long seed, frame_offset;
    uint32_t pkeys[20];   // the ten Philox round keys of `seed` (k0, k1 per round): constant-bank operands instead of a key schedule per block
    // two-stage Monte-Carlo (ldpc_mc_run_staged): stage 1 decodes T_run < T iterations with early termination and, instead
    // of counting a frame that has not reached a zero syndrome, appends its global frame index to defer_list; stage 2
    // regenerates exactly those frames (frame_list[k] instead of frame_offset + k: the Philox counter is the global index)
    // and decodes them in full, so the stragglers no longer hold CTAs of converged frames back.  Same counters, bit for bit.
    const unsigned long long *frame_list;
    unsigned long long *defer_list;
    unsigned int *defer_count;
    unsigned int defer_cap;       // entries defer_list holds; frames beyond it are counted in defer_count but not listed
    int punct_s, punct_e, short_s, short_e;
    // outputs (nullable)
    float *app; int app_all; long long app_stride_t;   // elements between iterations (B*NZ)
    uint32_t *hard; int HW;
    int *iters; uint8_t *flags; int *biterr;
    unsigned long long *counters;
    float *uncor_buf; unsigned int *uncor_count; unsigned int uncor_cap; int harvest_mode;
    // shared-memory carve-up (word offsets; the message array always starts at word 0)
    int off_xa, off_xq, off_hb, off_et, off_et2, off_w, off_misc, smem_words;   // off_et (E + 1 words) / off_et2: float kernels only
    int off_ring;      // persistent-slot kernels: the producers' ring of ready frames (NMS_MCP_RING_WORDS)
    // tables
    unsigned short row_ptr[LDPC_MAX_M + 1];   // E(C) edges of proto row i: [row_ptr[i], row_ptr[i+1])
    unsigned short col_ptr[LDPC_MAX_N + 1];   // CSR by proto column into vn_edge
    unsigned short cn_order[LDPC_MAX_M];      // rows sorted by degree (descending); slot s owns positions p % R == s
    unsigned short vn_order[LDPC_MAX_N];      // columns likewise
    int n_cn_cls, n_vn_cls;                   // runs of equal degree inside cn_order / vn_order
    ushort4 cn_cls[32], vn_cls[32];           // {degree, first position, end position, 0}
    // slot-major row list: [slot * ceil(M/R) + n] = {e0*LP*4, dc | row << 16}; dc 0 = none.  R * ceil(M/R) can exceed M by R - 1
    // and the specialised kernels prefetch one entry ahead: hence the slack (the host refuses R > 32)
    uint2 cn_task[LDPC_MAX_M + 34];
    unsigned short e_col[LDPC_MAX_E];         // proto column of E(C) edge e
    unsigned short e_sF[LDPC_MAX_E];          // s_e*Fp: check lane q -> variable lane (q + sF) mod L
    int2 vn_edge[LDPC_MAX_E];                 // column-sorted: {e*LP*4, ((L - s_e*Fp) mod L)*4}: byte offsets
};

struct LaunchGeom {
    int Fp, FB, L, LP, C, R, threads, smem_bytes, ctas_per_sm;
};

// one graph-specialised kernel (generated at build time by csrc/gen_spec.py)
struct NmsSpecEntry {
    const char *name;
    unsigned long long graph_hash;   // FNV-1a over (M, N, z, proto[])
    int M, N, z, E, Fp, R;
    const void *(*func)();
    int noet;   // 1: the variant for launches without early termination (never the default geometry)
};
#ifndef __CUDACC_RTC__
extern "C" const NmsSpecEntry *nms_spec_table(int *count);
extern "C" const NmsSpecEntry *nms_spec_f32_table(int *count);    // float path (decoding_type 1)
extern "C" const NmsSpecEntry *nms_spec_f32q_table(int *count);   // quantised twin (q_bit 6, per-edge weights)
extern "C" const NmsSpecEntry *nms_spec_mcp_table(int *count);    // persistent-slot Monte-Carlo kernels (nms_mcp.cuh)
#endif
#define NMS_MCP_MISC_WORDS(FB) (224 + (FB) * 16)   // their per-CTA state words: masks / per-pair words + 8 uint64 counters per slot
// Producer warps of the persistent-slot kernels (warp specialisation): with NMS_MCP_NP > 0 that many extra warps run the
// sample generator ahead into a ring of FB ready frames in shared memory while the other warps decode; 0 = the decoding warps
// generate in their own loop.  Measured on B200 (profiles/r02_mc_sweep.txt, "producer warps"): one producer warp cannot keep up
// (a single warp sustains ~0.3 IPC on the Philox chain: z72 at 7 dB 23 instead of 37 M frames/s); three are +2..3 % on z72,
// -3..6 % on WiMAX and -10..18 % on MacKay at 4-6 dB (the ring costs shared memory, i.e. resident CTAs, and the all-warp
// generator already runs at full issue rate) -- so the default stays 0.  One constant for host and device.
#ifndef NMS_MCP_NP
#define NMS_MCP_NP 0
#endif
#define NMS_MCP_RING_WORDS(FB, NZ) (NMS_MCP_NP > 0 ? (FB) * ((((NZ) + 3) & ~3) / 2) : 0)   // FB frames of NZ halves, 8-byte rows

// ---- Philox4x32-10 (Salmon et al., SC'11), written out so the host tests can restate it
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// same function with the round keys precomputed (KParams::pkeys)
__device__ __forceinline__ void philox4x32_10_keyed(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&k)[20],
                                                    uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#ifndef __CUDACC_RTC__
cudaError_t nms_launch_generate(const KParams &P, float *out, long long n_frames, cudaStream_t st);
cudaError_t nms_launch_normal_probe(unsigned long long seed, unsigned long long frame_offset, long long n_frames, int nquads,
                                    float *out, unsigned long long *counts, cudaStream_t st);
cudaError_t nms_launch_generate_cw(const KParams &P, const uint32_t *cw, long long cw_stride, float *out, long long n_frames,
                                   cudaStream_t st);
cudaError_t nms_launch_cw_metrics(const float *app, long long B, int T, int NZ, int target_bits, const uint32_t *cw, long long cw_stride,
                                  const int *iters, int early_term, uint8_t *flags, int *biterr, int *biterr_signed,
                                  unsigned long long *counters, cudaStream_t st);
void nms_note_launch();
unsigned long long nms_launch_count();
#endif
