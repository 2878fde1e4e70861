// ldpc_capi.cu -- C-ABI (include/ldpc_b200.h): base-graph compiler, decoder handles, launchers.
//
// Host-side equivalents of Main_Functions.init_parameter / init_connecting_matrix
// (Main_Functions.py:8-150) and weight_init (:387-439): instead of dense (E*z)^2 permutation
// matrices the compiler emits E(C)-ordered circulant shift tables and a column-sorted CSR edge
// list, pre-scaled for the lane-interleaved layout of the kernels (nms_common.cuh).
#include "../../include/ldpc_b200.h"
#include "nms_common.cuh"
#include "nms_train.cuh"
#include "nms_jit.h"
#include "host_pack.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

// kernel address getters, one per compiled degree bucket (nms_h2.cu / nms_f32.cu)
extern "C" {
const void *nms_h2_func_16_8();
const void *nms_h2_func_16_16();
const void *nms_h2_func_32_8();
const void *nms_h2_func_32_16();
const void *nms_h2_func_0_0();
const void *nms_f32_func_16_8();
const void *nms_f32_func_16_16();
const void *nms_f32_func_32_8();
const void *nms_f32_func_32_16();
const void *nms_f32_func_0_0();
}

namespace {

bool env_on(const char *name) {
    const char *v = getenv(name);
    return v && *v && *v != '0';
}

thread_local std::string g_err;
int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return fail(LDPC_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}   // namespace

struct ldpc_graph {
    int M = 0, N = 0, z = 0, E = 0;
    int punct_s = 0, punct_e = 0, short_s = 0, short_e = 0;
    std::vector<int> proto, row, col, shift, row_ptr, col_ptr, col_edge;
    ldpc_graph_info_t info{};
};

constexpr int HOST_SLOTS = 4;   // chunks in flight per ldpc_decode_host call (device buffers, stream, pinned staging each)

struct HostScratch {
    size_t cap_frames = 0;
    bool with_app = false;
    int app_iters = 0;
    float *llr[HOST_SLOTS] = {};      // float32 words of a chunk, or (same buffer) its int8 form
    float *app[HOST_SLOTS] = {};
    uint32_t *hard[HOST_SLOTS] = {};
    int *iters[HOST_SLOTS] = {};
    uint8_t *flags[HOST_SLOTS] = {};
    int *biterr[HOST_SLOTS] = {};
    // pinned host staging of the per-frame outputs, so the device-to-host copies stay asynchronous even when
    // the caller's result arrays are pageable (one blocking copy would serialise the pipeline)
    // (two per slot, used alternately: the results of a slot's previous chunk are handed over AFTER its next chunk has been
    // issued, so that copy is never between the device and its next piece of work)
    uint32_t *h_hard[2 * HOST_SLOTS] = {};
    int *h_iters[2 * HOST_SLOTS] = {};
    uint8_t *h_flags[2 * HOST_SLOTS] = {};
    int *h_biterr[2 * HOST_SLOTS] = {};
    cudaStream_t st[HOST_SLOTS] = {};
    cudaEvent_t ev_h2d[HOST_SLOTS] = {};   // recorded behind the host-to-device copy of the slot's current chunk
    // pinned staging of the INPUT words: the int8 form the host threads pack a float32 chunk into (one byte per value), or
    // the float32 words themselves for callers that pass pageable memory and words that have no int8 form
    char *h_in[HOST_SLOTS] = {};
    size_t h_in_bytes = 0;
    ldpc_host_stats_t stats{};
    unsigned long long *counters = nullptr;
    unsigned int *ucount = nullptr;
    float *ubuf = nullptr;
    size_t ubuf_rows = 0;
};

struct ldpc_decoder {
    ldpc_graph g;
    int sharing[3] = {0, 0, 0};
    int T = 0, decoding_type = 2, q_bit = 5, device = 0, sm_count = 148;
    float clip = 20.0f;
    bool packed = false;
    const void *func = nullptr;
    int dcb = 0, dvb = 0;
    const char *spec_name = nullptr;
    LaunchGeom geom{};
    KParams base{};
    // second geometry of the same graph-specialised kernel family for launches WITHOUT early termination (more frames per
    // CTA: fewer, fatter CTAs win when every frame runs all iterations, but the slowest frame of a CTA holds the others
    // back when they may stop early); nullptr = none
    const void *func_alt = nullptr;
    LaunchGeom geom_alt{};
    KParams base_alt{};
    // persistent-slot Monte-Carlo kernel (nms_mcp.cuh) of the same graph: serves ldpc_mc_run with early termination
    const void *func_mc = nullptr;
    const char *mc_name = nullptr;
    // run-time specialised kernels (nms_jit.cu) are driver-API functions: launched / queried through nms_jit_*
    bool func_cu = false, func_mc_cu = false;
    char jit_name[48] = {0}, jit_mc_name[48] = {0};
    LaunchGeom geom_mc{};
    KParams base_mc{};
    float *d_w = nullptr;
    // ldpc_decode_cw: per-iteration APPs of one chunk of frames + their flags / iteration counts
    float *d_cw_app = nullptr; size_t cw_app_cap = 0;
    int *d_cw_iters = nullptr; uint8_t *d_cw_flags = nullptr; size_t cw_frames_cap = 0;
    // training step (nms_train.cu): the weights as given ([T*wc | T*wu | T*wv], no "effective" rows), graph tables,
    // per-call workspace -- all created on first use
    int wc_raw = 0, wu_raw = 0, wv_raw = 0;
    bool ones_cn = false;
    std::vector<float> w_raw;
    float *d_w_raw = nullptr;
    int *d_tab = nullptr;
    float *d_hist = nullptr; size_t hist_cap = 0;
    float *d_coef = nullptr; float *d_grad = nullptr; double *d_loss = nullptr;
    std::mutex mu;
    HostScratch hs;
};

// ------------------------------------------------------------------------------------ graph
extern "C" const char *ldpc_last_error(void) { return g_err.c_str(); }
extern "C" int ldpc_version(void) { return 100; }

extern "C" int ldpc_graph_create(const int32_t *proto, int32_t M, int32_t N, int32_t z, int32_t ps, int32_t pe,
                                 int32_t ss, int32_t se, ldpc_graph_t **out) {
    if (!proto || !out || M <= 0 || N <= 0 || z <= 0) return fail(LDPC_E_INVALID, "graph_create: bad arguments");
    if (ps < 0 || pe < ps || ss < 0 || se < ss || pe > N * z || se > N * z)
        return fail(LDPC_E_INVALID, "graph_create: bad puncture/shorten range");
    ldpc_graph *g = new (std::nothrow) ldpc_graph();
    if (!g) return fail(LDPC_E_ALLOC, "graph_create: out of memory");
    g->M = M; g->N = N; g->z = z;
    g->punct_s = ps; g->punct_e = pe; g->short_s = ss; g->short_e = se;
    g->proto.assign(proto, proto + (size_t)M * N);
    g->row_ptr.assign(M + 1, 0);
    for (int i = 0; i < M; ++i) {            // E(C) = row-major edge order (Main_Functions.py:69-75)
        g->row_ptr[i] = (int)g->row.size();
        for (int j = 0; j < N; ++j) {
            const int p = proto[(size_t)i * N + j];
            if (p == -1) continue;
            g->row.push_back(i);
            g->col.push_back(j);
            g->shift.push_back(((p % z) + z) % z);   // :72
        }
    }
    g->E = (int)g->row.size();
    g->row_ptr[M] = g->E;
    g->col_ptr.assign(N + 1, 0);
    for (int e = 0; e < g->E; ++e) g->col_ptr[g->col[e] + 1]++;
    for (int j = 0; j < N; ++j) g->col_ptr[j + 1] += g->col_ptr[j];
    g->col_edge.assign(g->E, 0);
    {
        std::vector<int> fill(N, 0);
        for (int e = 0; e < g->E; ++e) g->col_edge[g->col_ptr[g->col[e]] + fill[g->col[e]]++] = e;   // ascending E(C)
    }
    ldpc_graph_info_t &I = g->info;
    I.M = M; I.N = N; I.z = z; I.E = g->E;
    I.max_dc = 0; I.max_dv = 0;
    for (int i = 0; i < M; ++i) I.max_dc = std::max(I.max_dc, g->row_ptr[i + 1] - g->row_ptr[i]);
    for (int j = 0; j < N; ++j) I.max_dv = std::max(I.max_dv, g->col_ptr[j + 1] - g->col_ptr[j]);
    const int punct_ref = pe - ps + 1, short_ref = se - ss + 1;   // "+1" even when 0,0 (:24-25)
    I.n_ref = N * z - punct_ref - short_ref;
    I.k_ref = (N - M) * z - short_ref;
    I.rate_ref = 1.0 * I.k_ref / I.n_ref;                          // :27-29
    const int punct_true = ps > 0 ? punct_ref : 0, short_true = ss > 0 ? short_ref : 0;
    I.n_true = N * z - punct_true - short_true;
    I.k_true = (N - M) * z - short_true;
    I.rate_true = 1.0 * I.k_true / I.n_true;
    *out = g;
    return LDPC_OK;
}

extern "C" int ldpc_graph_destroy(ldpc_graph_t *g) { delete g; return LDPC_OK; }

extern "C" int ldpc_graph_info(const ldpc_graph_t *g, ldpc_graph_info_t *info) {
    if (!g || !info) return fail(LDPC_E_INVALID, "graph_info: null argument");
    *info = g->info;
    return LDPC_OK;
}

extern "C" int ldpc_graph_edges(const ldpc_graph_t *g, int32_t *row, int32_t *col, int32_t *shift) {
    if (!g) return fail(LDPC_E_INVALID, "graph_edges: null graph");
    for (int e = 0; e < g->E; ++e) {
        if (row) row[e] = g->row[e];
        if (col) col[e] = g->col[e];
        if (shift) shift[e] = g->shift[e];
    }
    return LDPC_OK;
}

extern "C" int ldpc_graph_sigma(const ldpc_graph_t *g, const double *snr_db, int32_t n, int32_t use_ref_rate,
                                double *sigma) {
    if (!g || !snr_db || !sigma || n < 0) return fail(LDPC_E_INVALID, "graph_sigma: bad arguments");
    const double R = use_ref_rate ? g->info.rate_ref : g->info.rate_true;
    for (int i = 0; i < n; ++i) sigma[i] = std::sqrt(1.0 / (2.0 * std::pow(10.0, snr_db[i] / 10.0) * R));   // :35-36
    return LDPC_OK;
}

// ---------------------------------------------------------------------------------- decoder
namespace {

int weight_width(int code, int kind, int M, int N, int E) {   // Main_Functions.py:397-405
    if (code == 1) return E;
    if (code == 2) return kind == 2 ? N : M;
    if (code == 3) return 1;
    return 0;
}

// visiting order: sort by degree (descending, stable); returns the runs of equal degree.  Slot s of R owns the
// positions p with p % R == s, so every slot gets the same number of tasks (+-1) of about the same degrees.
int degree_classes(const std::vector<int> &deg, unsigned short *order, ushort4 *cls, int max_cls) {
    const int n = (int)deg.size();
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return deg[a] > deg[b]; });
    int nc = 0;
    for (int p = 0; p < n; ++p) {
        order[p] = (unsigned short)idx[p];
        if (p == 0 || deg[idx[p]] != deg[idx[p - 1]]) {
            if (nc == max_cls) return -1;
            cls[nc] = make_ushort4((unsigned short)deg[idx[p]], (unsigned short)p, (unsigned short)p, 0);
            ++nc;
        }
        cls[nc - 1].z = (unsigned short)(p + 1);
    }
    return nc;
}

const void *pick_kernel(bool packed, int max_dc, int max_dv, int *dcb, int *dvb) {
    int dc = max_dc <= 16 ? 16 : (max_dc <= 32 ? 32 : 0);
    int dv = max_dv <= 8 ? 8 : (max_dv <= 16 ? 16 : 0);
    if (dc == 0 || dv == 0) dc = dv = 0;
    if (env_on("LDPC_B200_FORCE_GENERIC")) dc = dv = 0;
    *dcb = dc; *dvb = dv;
    if (packed) {
        if (dc == 16 && dv == 8) return nms_h2_func_16_8();
        if (dc == 16 && dv == 16) return nms_h2_func_16_16();
        if (dc == 32 && dv == 8) return nms_h2_func_32_8();
        if (dc == 32 && dv == 16) return nms_h2_func_32_16();
        return nms_h2_func_0_0();
    }
    if (dc == 16 && dv == 8) return nms_f32_func_16_8();
    if (dc == 16 && dv == 16) return nms_f32_func_16_16();
    if (dc == 32 && dv == 8) return nms_f32_func_32_8();
    if (dc == 32 && dv == 16) return nms_f32_func_32_16();
    return nms_f32_func_0_0();
}

// spec_f32: graph-specialised float kernel (adds the per-(edge, chunk) syndrome table, nms_f32_spec.cuh)
void fill_smem_layout(KParams *P, bool packed, bool spec_f32 = false) {
    int off = P->E * P->LP;                 // message array at word 0
    off = (off + 3) & ~3;
    P->off_xa = off; off += P->N * P->LP * (packed ? 2 : 1);
    const bool drop_xq = P->no_xq && !env_on("LDPC_B200_KEEP_XQ");   // the env switch keeps the (then unused) array: occupancy A/B
    P->off_xq = off; off += ((packed && !drop_xq) || (!packed && P->qms)) ? P->N * P->LP : 0;
    P->off_hb = off; off += 2 * (packed ? 2 : 1) * P->N * P->C;
    P->off_et = off; off += packed ? 0 : P->E + 1;
    P->off_et2 = off; off += spec_f32 ? P->E * P->C * (P->L != P->LP ? 2 : 1) : 0;
    P->off_w = off; off += P->w_staged ? P->w_words : 0;
    P->off_misc = off; off += NMS_MISC_WORDS;
    P->smem_words = off;
}

// persistent-slot Monte-Carlo kernel: msg | xq (half2) | weights | state -- no float channel array, no ballots
void fill_smem_layout_mc(KParams *P) {
    int off = P->E * P->LP;
    off = (off + 3) & ~3;
    P->off_xa = off; P->off_xq = off; off += P->N * P->LP;
    P->off_hb = off; P->off_et = off; P->off_et2 = off;
    P->off_ring = off; off += (NMS_MCP_RING_WORDS(P->FB, P->N * P->z) + 3) & ~3;   // the producers' ring (8-byte rows)
    P->off_w = off; off += (P->w_words + 3) & ~3;      // the state block holds 64-bit counters
    P->off_misc = off; off += NMS_MCP_MISC_WORDS(P->FB);
    P->smem_words = off;
}

unsigned long long graph_hash(const ldpc_graph &g) {   // FNV-1a over (M, N, z, proto[]) -- same as gen_spec.py
    unsigned long long h = 0xcbf29ce484222325ull;
    auto mix = [&](int v) {
        unsigned u = (unsigned)v;
        for (int k = 0; k < 4; ++k) { h ^= (u >> (8 * k)) & 0xffu; h *= 0x100000001b3ull; }
    };
    mix(g.M); mix(g.N); mix(g.z);
    for (int v : g.proto) mix(v);
    return h;
}

// pick (Fp, R): lane efficiency x task balance x achievable warps/SM (from the real occupancy calculator)
int choose_geometry(const ldpc_graph &g, bool packed, bool qms, int w_words, const void *func, int force_fp,
                    int force_r, int max_warps, LaunchGeom *out, bool spec_f32 = false, bool no_xq = false) {
    const int max_smem = 227 * 1024;
    double best = -1.0;
    int forced_fp = force_fp, forced_r = force_r;
    if (!force_fp) {
        if (const char *s = getenv("LDPC_B200_FP")) forced_fp = atoi(s);
        if (const char *s = getenv("LDPC_B200_R")) forced_r = atoi(s);
    }
    const int fp_max = packed ? LDPC_MAX_FB / 2 : LDPC_MAX_FB;
    CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    for (int Fp = 1; Fp <= fp_max; ++Fp) {
        if (forced_fp && Fp != forced_fp) continue;
        const int L = g.z * Fp, LP = (L + 31) & ~31, C = LP / 32;
        if (C > 16) break;
        KParams tmp{};
        tmp.E = g.E; tmp.N = g.N; tmp.L = L; tmp.LP = LP; tmp.C = C; tmp.qms = qms; tmp.no_xq = no_xq;
        tmp.w_words = w_words; tmp.w_staged = w_words > 0 && w_words <= NMS_WSTAGE_MAX_WORDS;
        fill_smem_layout(&tmp, packed, spec_f32);
        const int smem = tmp.smem_words * 4;
        if (smem > max_smem) break;
        for (int R = 1; R * C <= max_warps; ++R) {
            if (forced_r && R != forced_r) continue;
            const int W = C * R;
            if (W < 2) continue;   // the per-frame bookkeeping uses two warps
            if (R > std::max(g.M, g.N)) break;
            int cps = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, func, W * 32, smem) != cudaSuccess || cps < 1) {
                cudaGetLastError();
                continue;
            }
            const double lane_eff = (double)L / LP;
            const double bal_m = (double)g.M / (((g.M + R - 1) / R) * R), bal_n = (double)g.N / (((g.N + R - 1) / R) * R);
            const double warps = (double)cps * W;
            const double occ = std::min(1.0, warps / 28.0);
            const double score = lane_eff * (0.45 * bal_m + 0.55 * bal_n) * occ + 1e-4 * warps - 1e-5 * smem / 1024.0;
            if (score > best) {
                best = score;
                out->Fp = Fp; out->FB = packed ? 2 * Fp : Fp; out->L = L; out->LP = LP; out->C = C; out->R = R;
                out->threads = W * 32; out->smem_bytes = smem; out->ctas_per_sm = cps;
            }
        }
    }
    if (best < 0) return fail(LDPC_E_LIMIT, "no launch geometry fits this graph in shared memory");
    return LDPC_OK;
}

// does this call go to the persistent-slot Monte-Carlo kernel?  (in-kernel samples, per-frame early termination,
// no per-frame outputs; LDPC_B200_NO_PERSIST=1 keeps the batch kernels: the A/B switch of the tests)
bool takes_mc_kernel(const ldpc_decoder *d, const KParams &P) {
    return d->func_mc != nullptr && P.llr == nullptr && P.llr_q8 == nullptr && P.early_term && P.frame_list == nullptr &&
           P.defer_list == nullptr && P.app == nullptr && P.hard == nullptr && P.iters == nullptr && P.flags == nullptr &&
           P.biterr == nullptr && !env_on("LDPC_B200_NO_PERSIST");
}

int launch(const ldpc_decoder *d, const KParams &Pin, cudaStream_t st) {
    const bool mc = takes_mc_kernel(d, Pin);
    const bool alt = mc || (d->func_alt != nullptr && !Pin.early_term);
    KParams Q;
    if (alt) {   // same call, the other geometry's tables and shared-memory layout
        Q = mc ? d->base_mc : d->base_alt;
        Q.T_run = Pin.T_run; Q.early_term = Pin.early_term;
        Q.llr = Pin.llr; Q.llr_q8 = Pin.llr_q8; Q.q8_step = Pin.q8_step; Q.n_frames = Pin.n_frames;
        Q.sigma = Pin.sigma; Q.two_over_s2 = Pin.two_over_s2; Q.two_over_s = Pin.two_over_s; std::memcpy(Q.pkeys, Pin.pkeys, sizeof Q.pkeys); Q.seed = Pin.seed; Q.frame_offset = Pin.frame_offset;
        Q.app = Pin.app; Q.app_all = Pin.app_all; Q.app_stride_t = Pin.app_stride_t;
        Q.hard = Pin.hard; Q.iters = Pin.iters; Q.flags = Pin.flags; Q.biterr = Pin.biterr; Q.counters = Pin.counters;
        Q.uncor_buf = Pin.uncor_buf; Q.uncor_count = Pin.uncor_count; Q.uncor_cap = Pin.uncor_cap; Q.harvest_mode = Pin.harvest_mode;
        Q.w_all = Pin.w_all;
        Q.frame_list = Pin.frame_list; Q.defer_list = Pin.defer_list; Q.defer_count = Pin.defer_count; Q.defer_cap = Pin.defer_cap;
    }
    const KParams &P = alt ? Q : Pin;
    const LaunchGeom &geo = mc ? d->geom_mc : (alt ? d->geom_alt : d->geom);
    const long long nb = (P.n_frames + P.FB - 1) / P.FB;
    if (nb <= 0) return LDPC_OK;
    const int grid = (int)std::min<long long>(nb, (long long)d->sm_count * geo.ctas_per_sm);
    void *args[] = {(void *)&P};
    const void *func = mc ? d->func_mc : (alt ? d->func_alt : d->func);
    if (mc ? d->func_mc_cu : (!alt && d->func_cu)) {
        if (nms_jit_launch(const_cast<void *>(func), grid, geo.threads, geo.smem_bytes, st, &P) != 0)
            return fail(LDPC_E_CUDA, "launch of the run-time specialised kernel failed");
        nms_note_launch();
        return LDPC_OK;
    }
    CUDA_TRY(cudaLaunchKernel(func, dim3(grid), dim3(geo.threads), args, (size_t)geo.smem_bytes, st));
    nms_note_launch();
    return LDPC_OK;
}

}   // namespace

extern "C" int ldpc_decoder_create(const ldpc_graph_t *g, const int32_t sharing[3], int32_t T, const float *w_cn,
                                   const float *w_ucn, const float *w_vn, int32_t decoding_type, int32_t q_bit,
                                   float clip_llr, int32_t device, ldpc_decoder_t **out) {
    return ldpc_decoder_create2(g, sharing, T, w_cn, w_ucn, w_vn, decoding_type, q_bit, clip_llr, device, 0, out);
}

extern "C" int ldpc_decoder_create2(const ldpc_graph_t *g, const int32_t sharing_in[3], int32_t T, const float *w_cn,
                                    const float *w_ucn, const float *w_vn, int32_t decoding_type, int32_t q_bit,
                                    float clip_llr, int32_t device, int32_t target_node, ldpc_decoder_t **out) {
    if (!g || !sharing_in || !out || T <= 0 || T > LDPC_MAX_T) return fail(LDPC_E_INVALID, "decoder_create: bad arguments");
    if (target_node < 0 || target_node > g->N) return fail(LDPC_E_INVALID, "target_node %d outside 0..%d", target_node, g->N);
    // check_params (Main_Functions.py:507-521)
    for (int i = 0; i < 3; ++i)
        if (sharing_in[i] < 0 || sharing_in[i] > 4)
            return fail(LDPC_E_UNSUPPORTED, "sharing code %d (code 5 has no branch in build_neural_network)", sharing_in[i]);
    if (sharing_in[2] == 1 || sharing_in[2] == 4) return fail(LDPC_E_INVALID, "sharing[2] in [1,4] (Main_Functions.py:515-517)");
    if (sharing_in[1] != 0 && sharing_in[1] != sharing_in[0])
        return fail(LDPC_E_INVALID, "sharing[1] != 0 and sharing[0] != sharing[1] (Main_Functions.py:519-521)");
    // temporal sharing (code 4): the caller passes the T expanded rows (what print_weight writes, Print_Functions.py:
    // 87-94), which makes it per-edge sharing; its UCN twin has no branch in build_neural_network (:299-304)
    int32_t sharing[3] = {sharing_in[0], sharing_in[1], sharing_in[2]};
    if (sharing[0] == 4) { sharing[0] = 1; sharing[1] = 0; }
    if (decoding_type < 0 || decoding_type > 2)
        return fail(LDPC_E_UNSUPPORTED, "decoding_type %d (0 = sum-product, 1 = min-sum, 2 = quantised min-sum)", decoding_type);
    if (decoding_type == 0 && g->info.max_dc > 64)
        return fail(LDPC_E_LIMIT, "row degree > 64 in sum-product mode");
    float qk = 1.f, qmax = 0.f;
    if (decoding_type == 2) {
        switch (q_bit) {   // Main_Functions.py:483-492
        case 5: qk = 2.f; qmax = 7.5f; break;
        case 6: qk = 1.f; qmax = 15.5f; break;
        case -5: qk = 1.f; qmax = 15.f; break;
        case 4: qk = 1.f; qmax = 7.f; break;
        case 3: qk = 0.5f; qmax = 6.f; break;
        default: return fail(LDPC_E_INVALID, "q_bit %d has no quantiser branch", q_bit);
        }
    }
    if ((sharing[0] && !w_cn) || (sharing[1] && !w_ucn) || (sharing[2] && !w_vn))
        return fail(LDPC_E_INVALID, "decoder_create: missing weight block");
    if (!(clip_llr > 0.f)) return fail(LDPC_E_INVALID, "clip_llr must be positive");
    if (g->M > LDPC_MAX_M || g->N > LDPC_MAX_N || g->E > LDPC_MAX_E)
        return fail(LDPC_E_LIMIT, "graph %dx%d with %d edges exceeds the table limits (%d, %d, %d)", g->M, g->N, g->E,
                    LDPC_MAX_M, LDPC_MAX_N, LDPC_MAX_E);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(LDPC_E_CUDA, "no CUDA device: this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(LDPC_E_INVALID, "device %d out of range", device);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", device);

    ldpc_decoder *d = new (std::nothrow) ldpc_decoder();
    if (!d) return fail(LDPC_E_ALLOC, "decoder_create: out of memory");
    d->g = *g;
    std::copy(sharing, sharing + 3, d->sharing);
    d->T = T; d->decoding_type = decoding_type; d->q_bit = q_bit; d->clip = clip_llr; d->device = device;
    const bool qms = decoding_type == 2;
    // packed fp16x2 kernel: uniform quantiser grids closed under addition, no per-edge weights
    d->packed = qms && q_bit != 6 && sharing[0] != 1 && g->info.max_dv <= 30 && !env_on("LDPC_B200_FORCE_F32");
    if (!d->packed && g->info.max_dv > 64) { delete d; return fail(LDPC_E_LIMIT, "column degree > 64 in float mode"); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete d; return fail(LDPC_E_CUDA, "cudaGetDeviceProperties"); }
    d->sm_count = prop.multiProcessorCount;
    int wc = weight_width(sharing[0], 0, g->M, g->N, g->E), wu = weight_width(sharing[1], 1, g->M, g->N, g->E),
        wv = weight_width(sharing[2], 2, g->M, g->N, g->E);
    // host copy of the device weight block [cn | ucn | vn]; the packed kernels get an "effective" block:
    // a row of ones where the reference applies no CN weight (Main_Functions.py:267-268)
    std::vector<float> wh;
    const bool ones_cn = d->packed && sharing[0] == 0;
    if (ones_cn) { wc = 1; wh.assign((size_t)T, 1.0f); }
    else if (wc) wh.assign(w_cn, w_cn + (size_t)T * wc);
    if (wu) wh.insert(wh.end(), w_ucn, w_ucn + (size_t)T * wu);
    if (wv) wh.insert(wh.end(), w_vn, w_vn + (size_t)T * wv);
    if (d->packed && (int)wh.size() > NMS_WSTAGE_MAX_WORDS) d->packed = false;   // weights must fit in shared memory
    if (!d->packed && ones_cn) { wc = 0; wh.erase(wh.begin(), wh.begin() + T); }
    const int w_words = (int)wh.size();
    d->wc_raw = weight_width(sharing[0], 0, g->M, g->N, g->E); d->wu_raw = wu; d->wv_raw = wv;
    d->ones_cn = d->packed && sharing[0] == 0;
    if (d->wc_raw) d->w_raw.assign(w_cn, w_cn + (size_t)T * d->wc_raw);
    if (wu) d->w_raw.insert(d->w_raw.end(), w_ucn, w_ucn + (size_t)T * wu);
    if (wv) d->w_raw.insert(d->w_raw.end(), w_vn, w_vn + (size_t)T * wv);
    // a graph known at build time gets its specialised kernel (gen_spec.py); anything else the generic buckets
    const bool no_xq = sharing[2] != 0;   // packed kernels: see KParams::no_xq
    int rc = LDPC_E_LIMIT;
    d->func = nullptr;
    if (d->packed && !env_on("LDPC_B200_NO_SPEC")) {
        int n = 0;
        const NmsSpecEntry *tab = nms_spec_table(&n);
        const unsigned long long h = graph_hash(d->g);
        const int want_fp = getenv("LDPC_B200_FP") ? atoi(getenv("LDPC_B200_FP")) : 0;
        const int want_r = getenv("LDPC_B200_R") ? atoi(getenv("LDPC_B200_R")) : 0;
        for (int k = 0; k < n; ++k) {
            if (tab[k].graph_hash != h || tab[k].M != g->M || tab[k].N != g->N || tab[k].z != g->z) continue;
            if ((want_fp && tab[k].Fp != want_fp) || (want_r && tab[k].R != want_r)) continue;   // tuning override
            const void *f = tab[k].func();
            LaunchGeom geo{};
            if (choose_geometry(d->g, true, qms, w_words, f, tab[k].Fp, tab[k].R, 32, &geo, false, no_xq) == LDPC_OK) {
                if (d->func == nullptr) {
                    if (tab[k].noet && !want_fp) continue;   // the fixed-iteration variant never is the default
                    d->func = f; d->geom = geo; d->spec_name = tab[k].name; rc = LDPC_OK;
                    if (want_fp || want_r) break;            // tuning override: exactly this geometry, no second one
                } else if (tab[k].noet && d->func_alt == nullptr && !env_on("LDPC_B200_NO_ALT")) {
                    d->func_alt = f; d->geom_alt = geo;
                }
            }
        }
    }
    // persistent-slot Monte-Carlo kernel of the same graph
    if (d->packed && d->spec_name != nullptr && g->N * g->z < 65536 && !env_on("LDPC_B200_NO_SPEC")) {
        int n = 0;
        const NmsSpecEntry *tab = nms_spec_mcp_table(&n);
        const unsigned long long h = graph_hash(d->g);
        const int want_fp = getenv("LDPC_B200_MCP_FP") ? atoi(getenv("LDPC_B200_MCP_FP")) : 0;
        const int want_r = getenv("LDPC_B200_MCP_R") ? atoi(getenv("LDPC_B200_MCP_R")) : 0;
        for (int k = 0; k < n && d->func_mc == nullptr; ++k) {
            if (tab[k].graph_hash != h || tab[k].M != g->M || tab[k].N != g->N || tab[k].z != g->z) continue;
            if ((want_fp && tab[k].Fp != want_fp) || (want_r && tab[k].R != want_r)) continue;
            const void *f = tab[k].func();
            LaunchGeom geo{};
            geo.Fp = tab[k].Fp; geo.FB = 2 * geo.Fp; geo.L = g->z * geo.Fp; geo.LP = (geo.L + 31) & ~31; geo.C = geo.LP / 32;
            geo.R = tab[k].R; geo.threads = geo.C * geo.R * 32 + NMS_MCP_NP * 32;   // decoding warps + producer warps
            KParams tmp{};
            tmp.E = g->E; tmp.N = g->N; tmp.z = g->z; tmp.LP = geo.LP; tmp.w_words = w_words; tmp.FB = geo.FB;
            fill_smem_layout_mc(&tmp);
            geo.smem_bytes = tmp.smem_words * 4;
            if (geo.smem_bytes > 227 * 1024) continue;
            if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.smem_bytes) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&geo.ctas_per_sm, f, geo.threads, geo.smem_bytes) != cudaSuccess ||
                geo.ctas_per_sm < 1) {
                cudaGetLastError();
                continue;
            }
            d->func_mc = f; d->geom_mc = geo; d->mc_name = tab[k].name;
        }
    }
    // a graph the build does not know: specialise it now (NVRTC, cached on disk) -- packed decode kernel and, for lifted
    // graphs, the persistent-slot Monte-Carlo kernel.  Any failure leaves the generic degree-bucketed kernels in charge.
    bool no_empty_node = true;       // the unrolled code has no body for a node without edges
    for (int i = 0; i < g->M; ++i) no_empty_node = no_empty_node && g->row_ptr[i + 1] > g->row_ptr[i];
    for (int j = 0; j < g->N; ++j) no_empty_node = no_empty_node && g->col_ptr[j + 1] > g->col_ptr[j];
    if (d->packed && d->func == nullptr && !env_on("LDPC_B200_NO_SPEC") && !env_on("LDPC_B200_NO_JIT") && g->info.max_dc <= 64 &&
        no_empty_node && nms_jit_available()) {
        char err[512] = {0};
        const unsigned long long h = graph_hash(d->g);
        for (int kind = NMS_JIT_DECODE; kind <= NMS_JIT_MCP; ++kind) {
            if (kind == NMS_JIT_MCP && (g->N * g->z >= 65536 || env_on("LDPC_B200_NO_JIT_MCP"))) continue;
            int Fp = 0, R = 0;
            nms_jit_pick_geometry(g->M, g->N, g->E, g->z, kind, g->info.max_dc, &Fp, &R);
            const char *efp = getenv(kind == NMS_JIT_MCP ? "LDPC_B200_MCP_FP" : "LDPC_B200_FP");
            const char *er = getenv(kind == NMS_JIT_MCP ? "LDPC_B200_MCP_R" : "LDPC_B200_R");
            if (efp && atoi(efp) > 0) Fp = atoi(efp);
            if (er && atoi(er) > 0) R = atoi(er);
            LaunchGeom geo{};
            geo.Fp = Fp; geo.FB = 2 * Fp; geo.L = g->z * Fp; geo.LP = (geo.L + 31) & ~31; geo.C = geo.LP / 32; geo.R = R;
            geo.threads = geo.C * R * 32 + (kind == NMS_JIT_MCP ? NMS_MCP_NP * 32 : 0);
            if (geo.C > 16 || geo.threads > 1024 || geo.FB > LDPC_MAX_FB) continue;
            KParams tmp{};
            tmp.E = g->E; tmp.N = g->N; tmp.z = g->z; tmp.L = geo.L; tmp.LP = geo.LP; tmp.C = geo.C; tmp.qms = 1; tmp.no_xq = no_xq; tmp.FB = geo.FB;
            tmp.w_words = w_words; tmp.w_staged = w_words > 0 && w_words <= NMS_WSTAGE_MAX_WORDS;
            if (kind == NMS_JIT_MCP) fill_smem_layout_mc(&tmp); else fill_smem_layout(&tmp, true, false);
            geo.smem_bytes = tmp.smem_words * 4;
            if (geo.smem_bytes > 227 * 1024) continue;
            void *fn = nullptr;
            if (nms_jit_build(d->g.proto.data(), g->M, g->N, g->z, Fp, R, kind, &fn, err, (int)sizeof err) != 0 ||
                nms_jit_set_smem(fn, geo.smem_bytes) != 0 ||
                nms_jit_occupancy(fn, geo.threads, geo.smem_bytes, &geo.ctas_per_sm) != 0 || geo.ctas_per_sm < 1) {
                if (env_on("LDPC_B200_JIT_VERBOSE")) fprintf(stderr, "ldpc_b200: run-time specialisation failed: %s\n", err);
                cudaGetLastError();
                continue;
            }
            if (kind == NMS_JIT_DECODE) {
                snprintf(d->jit_name, sizeof d->jit_name, "jit%016llx_fp%d_r%d", h, Fp, R);
                d->func = fn; d->func_cu = true; d->geom = geo; d->spec_name = d->jit_name; rc = LDPC_OK;
            } else if (d->func_cu) {
                snprintf(d->jit_mc_name, sizeof d->jit_mc_name, "jit%016llx_fp%d_r%d", h, Fp, R);
                d->func_mc = fn; d->func_mc_cu = true; d->geom_mc = geo; d->mc_name = d->jit_mc_name;
            }
        }
    }
    // graph-specialised float32 kernel (float / quantised twin); it reads its weights from shared memory
    if (!d->packed && w_words <= NMS_WSTAGE_MAX_WORDS && !env_on("LDPC_B200_NO_SPEC")) {
        int n = 0;
        const NmsSpecEntry *tab = qms ? nms_spec_f32q_table(&n) : nms_spec_f32_table(&n);
        const unsigned long long h = graph_hash(d->g);
        const int want_fp = getenv("LDPC_B200_FP") ? atoi(getenv("LDPC_B200_FP")) : 0;
        const int want_r = getenv("LDPC_B200_R") ? atoi(getenv("LDPC_B200_R")) : 0;
        for (int k = 0; k < n; ++k) {
            if (tab[k].graph_hash != h || tab[k].M != g->M || tab[k].N != g->N || tab[k].z != g->z) continue;
            if ((want_fp && tab[k].Fp != want_fp) || (want_r && tab[k].R != want_r)) continue;   // tuning override
            const void *f = tab[k].func();
            LaunchGeom geo{};
            if (choose_geometry(d->g, false, qms, w_words, f, tab[k].Fp, tab[k].R, 32, &geo, true) == LDPC_OK) {
                d->func = f; d->geom = geo; d->spec_name = tab[k].name; rc = LDPC_OK;
                break;
            }
        }
    }
    if (d->func == nullptr) {
        d->func = pick_kernel(d->packed, g->info.max_dc, g->info.max_dv, &d->dcb, &d->dvb);
        rc = choose_geometry(d->g, d->packed, qms, w_words, d->func, 0, 0, 16, &d->geom, false, d->packed && no_xq);   // generic kernels: __launch_bounds__(512)
    }
    if (rc != LDPC_OK) { delete d; return rc; }

    const char *perr = nullptr;
    auto fill = [&](KParams &P, const LaunchGeom &geom, bool mc_layout = false) {
    std::memset(&P, 0, sizeof P);
    P.M = g->M; P.N = g->N; P.E = g->E; P.z = g->z; P.NZ = g->N * g->z;
    P.Fp = geom.Fp; P.FB = geom.FB; P.L = geom.L; P.LP = geom.LP; P.C = geom.C; P.R = geom.R;
    P.qms = qms; P.sp = decoding_type == 0; P.qmagic = 12582912.0f / qk; P.qmax = qmax; P.clip = clip_llr;
    {
        const __half hq = __float2half_rn(qmax);
        unsigned short bits;
        std::memcpy(&bits, &hq, sizeof bits);
        P.qmax_h2 = (uint32_t)bits | ((uint32_t)bits << 16);
    }
    P.sat_magic = qms ? P.qmagic : 0.0f; P.sat_bound = qms ? qmax : clip_llr;
    P.sharing0 = sharing[0]; P.sharing1 = sharing[1]; P.sharing2 = sharing[2];
    P.wc = wc; P.wu = wu; P.wv = wv;
    P.w_words = (int)wh.size(); P.w_staged = P.w_words > 0 && P.w_words <= NMS_WSTAGE_MAX_WORDS;
    P.w_off_cn = 0; P.w_off_ucn = T * wc; P.w_off_vn = T * (wc + wu);
    P.T_run = T;
    P.no_xq = d->packed && no_xq;
    P.target_n = target_node > 0 ? target_node : g->N;
    P.punct_s = g->punct_s; P.punct_e = g->punct_e; P.short_s = g->short_s; P.short_e = g->short_e;
    P.HW = (P.NZ + 31) / 32;
    if (mc_layout) fill_smem_layout_mc(&P);
    else fill_smem_layout(&P, d->packed, !d->packed && d->spec_name != nullptr);
    if (d->packed) {
        P.h2w_c = P.off_w + P.w_off_cn; P.h2_wc = wc; P.h2_mc = wc > 1 ? -1 : 0;
        if (wu) { P.h2w_u = P.off_w + P.w_off_ucn; P.h2_wu = wu; P.h2_mu = wu > 1 ? -1 : 0; }
        else { P.h2w_u = P.h2w_c; P.h2_wu = P.h2_wc; P.h2_mu = P.h2_mc; }
        P.h2w_v = P.off_w + P.w_off_vn; P.h2_wv = wv; P.h2_mv = wv > 1 ? -1 : 0;
    }
    if (P.smem_words * 4 != geom.smem_bytes) { perr = "internal: smem layout mismatch"; return; }
    for (int i = 0; i <= g->M; ++i) P.row_ptr[i] = (unsigned short)g->row_ptr[i];
    for (int j = 0; j <= g->N; ++j) P.col_ptr[j] = (unsigned short)g->col_ptr[j];
    {
        std::vector<int> dc(g->M), dv(g->N);
        for (int i = 0; i < g->M; ++i) dc[i] = g->row_ptr[i + 1] - g->row_ptr[i];
        for (int j = 0; j < g->N; ++j) dv[j] = g->col_ptr[j + 1] - g->col_ptr[j];
        P.n_cn_cls = degree_classes(dc, P.cn_order, P.cn_cls, 32);
        P.n_vn_cls = degree_classes(dv, P.vn_order, P.vn_cls, 32);
        if (P.n_cn_cls < 0 || P.n_vn_cls < 0) { perr = "more than 32 distinct node degrees"; return; }
    }
    {
        const int NT = (g->M + P.R - 1) / P.R;   // rows per task slot (slot s owns positions s, s+R, ... of cn_order)
        if (P.R > 32 || P.R * NT + 1 > (int)(sizeof P.cn_task / sizeof P.cn_task[0])) { perr = "row task table too small"; return; }
        for (int s = 0; s < P.R; ++s)
            for (int n = 0; n < NT; ++n) {
                const int p = s + n * P.R;
                uint2 tk = make_uint2(0u, 0u);
                if (p < g->M) {
                    const int i = P.cn_order[p];
                    tk.x = (unsigned)(g->row_ptr[i] * P.LP * 4);
                    tk.y = (unsigned)(g->row_ptr[i + 1] - g->row_ptr[i]) | ((unsigned)i << 16);
                }
                P.cn_task[s * NT + n] = tk;
            }
    }
    for (int e = 0; e < g->E; ++e) {
        P.e_col[e] = (unsigned short)g->col[e];
        P.e_sF[e] = (unsigned short)(g->shift[e] * P.Fp);
    }
    for (int k = 0; k < g->E; ++k) {
        const int e = g->col_edge[k];
        // variable lane q -> check lane (q - s*Fp) mod L, as byte offsets into shared memory
        P.vn_edge[k].x = e * P.LP * 4;
        P.vn_edge[k].y = ((P.L - g->shift[e] * P.Fp) % P.L) * 4;
    }
    };
    fill(d->base, d->geom);
    if (!perr && d->func_alt) fill(d->base_alt, d->geom_alt);
    if (!perr && d->func_mc) fill(d->base_mc, d->geom_mc, true);
    if (perr) { const std::string msg = perr; delete d; return fail(LDPC_E_LIMIT, "%s", msg.c_str()); }
    KParams &P = d->base;
    if (!wh.empty()) {
        if (cudaMalloc(&d->d_w, wh.size() * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(d->d_w, wh.data(), wh.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            const char *msg = cudaGetErrorString(cudaGetLastError());
            delete d;
            return fail(LDPC_E_CUDA, "uploading weights: %s", msg);
        }
        P.w_all = d->d_w;
        d->base_alt.w_all = d->d_w;
        d->base_mc.w_all = d->d_w;
    }
    *out = d;
    return LDPC_OK;
}

namespace {
void free_scratch(HostScratch &h) {
    for (int i = 0; i < HOST_SLOTS; ++i) {
        cudaFree(h.llr[i]); cudaFree(h.app[i]); cudaFree(h.hard[i]); cudaFree(h.iters[i]);
        cudaFree(h.flags[i]); cudaFree(h.biterr[i]);
        for (int j = i; j < 2 * HOST_SLOTS; j += HOST_SLOTS) {
            cudaFreeHost(h.h_hard[j]); cudaFreeHost(h.h_iters[j]); cudaFreeHost(h.h_flags[j]); cudaFreeHost(h.h_biterr[j]);
        }
        cudaFreeHost(h.h_in[i]);
        if (h.st[i]) cudaStreamDestroy(h.st[i]);
        if (h.ev_h2d[i]) cudaEventDestroy(h.ev_h2d[i]);
    }
    cudaFree(h.counters); cudaFree(h.ucount); cudaFree(h.ubuf);
    h = HostScratch();
}
}   // namespace

extern "C" int ldpc_decoder_destroy(ldpc_decoder_t *d) {
    if (!d) return LDPC_OK;
    {
        DeviceGuard guard(d->device);
        free_scratch(d->hs);
        cudaFree(d->d_cw_app); cudaFree(d->d_cw_iters); cudaFree(d->d_cw_flags);
        cudaFree(d->d_w); cudaFree(d->d_w_raw); cudaFree(d->d_tab); cudaFree(d->d_hist); cudaFree(d->d_coef);
        cudaFree(d->d_grad); cudaFree(d->d_loss);
    }
    delete d;
    return LDPC_OK;
}

extern "C" int ldpc_decoder_uses_packed_kernel(const ldpc_decoder_t *d) { return d && d->packed ? 1 : 0; }

extern "C" const char *ldpc_decoder_kernel_name(const ldpc_decoder_t *d) {
    static thread_local char buf[96];
    if (!d) return "";
    if (d->spec_name) snprintf(buf, sizeof buf, "nms_%s_spec_%s", d->packed ? "h2" : (d->decoding_type == 2 ? "f32q" : "f32"), d->spec_name);
    else snprintf(buf, sizeof buf, "nms_%s_kernel_%d_%d", d->packed ? "h2" : "f32", d->dcb, d->dvb);
    return buf;
}

extern "C" int ldpc_decoder_geometry(const ldpc_decoder_t *d, int32_t *frames_per_cta, int32_t *ctas_per_sm,
                                     int32_t *threads_per_cta, int32_t *smem_bytes) {
    if (!d) return fail(LDPC_E_INVALID, "decoder_geometry: null decoder");
    if (frames_per_cta) *frames_per_cta = d->geom.FB;
    if (ctas_per_sm) *ctas_per_sm = d->geom.ctas_per_sm;
    if (threads_per_cta) *threads_per_cta = d->geom.threads;
    if (smem_bytes) *smem_bytes = d->geom.smem_bytes;
    return LDPC_OK;
}

// the kernel and launch geometry a call with / without early termination uses (they differ when the graph has a second,
// fixed-iteration geometry: ldpc_decoder::func_alt)
extern "C" int ldpc_decoder_launch_info(const ldpc_decoder_t *d, int32_t early_term, int32_t *frames_per_cta, int32_t *ctas_per_sm,
                                        int32_t *threads_per_cta, int32_t *smem_bytes, char *kernel_name, int32_t name_cap) {
    if (!d) return fail(LDPC_E_INVALID, "decoder_launch_info: null decoder");
    const bool alt = d->func_alt != nullptr && !early_term;
    const LaunchGeom &geo = alt ? d->geom_alt : d->geom;
    if (frames_per_cta) *frames_per_cta = geo.FB;
    if (ctas_per_sm) *ctas_per_sm = geo.ctas_per_sm;
    if (threads_per_cta) *threads_per_cta = geo.threads;
    if (smem_bytes) *smem_bytes = geo.smem_bytes;
    if (kernel_name && name_cap > 0) {
        if (alt) snprintf(kernel_name, (size_t)name_cap, "nms_h2_spec_%s_fp%d_r%d", d->spec_name ? std::string(d->spec_name).substr(0, std::string(d->spec_name).rfind("_fp")).c_str() : "", geo.Fp, geo.R);
        else snprintf(kernel_name, (size_t)name_cap, "%s", ldpc_decoder_kernel_name(d));
    }
    return LDPC_OK;
}

// kernel and launch geometry of ldpc_mc_run(early_term = 1): the persistent-slot kernel when the graph has one
extern "C" int ldpc_decoder_mc_info(const ldpc_decoder_t *d, int32_t *persistent, int32_t *frames_per_cta, int32_t *ctas_per_sm,
                                    int32_t *threads_per_cta, int32_t *smem_bytes, char *kernel_name, int32_t name_cap) {
    if (!d) return fail(LDPC_E_INVALID, "decoder_mc_info: null decoder");
    const bool mc = d->func_mc != nullptr && !env_on("LDPC_B200_NO_PERSIST");
    if (!mc) {
        if (persistent) *persistent = 0;
        return ldpc_decoder_launch_info(d, 1, frames_per_cta, ctas_per_sm, threads_per_cta, smem_bytes, kernel_name, name_cap);
    }
    if (persistent) *persistent = 1;
    if (frames_per_cta) *frames_per_cta = d->geom_mc.FB;
    if (ctas_per_sm) *ctas_per_sm = d->geom_mc.ctas_per_sm;
    if (threads_per_cta) *threads_per_cta = d->geom_mc.threads;
    if (smem_bytes) *smem_bytes = d->geom_mc.smem_bytes;
    if (kernel_name && name_cap > 0) snprintf(kernel_name, (size_t)name_cap, "nms_mcp_spec_%s", d->mc_name);
    return LDPC_OK;
}

// ----------------------------------------------------------------------------------- decode
namespace {
float quantiser_step(const ldpc_decoder *d) {   // 1/qk of Main_Functions.py:483-492; 0 for the float decoders
    if (d->decoding_type != 2) return 0.0f;
    return d->q_bit == 5 ? 0.5f : (d->q_bit == 3 ? 2.0f : 1.0f);
}

int decode_dev(const ldpc_decoder *d, const float *llr_dev, const int8_t *llr_q8_dev, float step, int64_t B,
               int32_t iters, int32_t early_term, float *app_dev, int32_t app_all_iters, uint32_t *hard_dev,
               int32_t *iters_dev, uint8_t *flags_dev, int32_t *biterr_dev, uint64_t *counters_dev, void *stream) {
    if (!d || (!llr_dev && !llr_q8_dev && B > 0) || B < 0) return fail(LDPC_E_INVALID, "decode: bad arguments");
    if (iters < 0 || iters > d->T) return fail(LDPC_E_INVALID, "decode: iters %d outside 0..%d", iters, d->T);
    if (llr_q8_dev && !(step > 0.0f)) return fail(LDPC_E_INVALID, "decode_q8: step must be positive for a float decoder");
    if (B == 0) return LDPC_OK;
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    KParams P = d->base;
    P.T_run = iters == 0 ? d->T : iters;
    P.early_term = early_term ? 1 : 0;
    P.llr = llr_dev; P.llr_q8 = (const signed char *)llr_q8_dev; P.q8_step = step; P.n_frames = B;
    P.app = app_dev; P.app_all = app_all_iters ? 1 : 0; P.app_stride_t = (long long)B * P.NZ;
    P.hard = hard_dev; P.iters = iters_dev; P.flags = flags_dev; P.biterr = biterr_dev;
    P.counters = (unsigned long long *)counters_dev;
    return launch(d, P, (cudaStream_t)stream);
}
}   // namespace

extern "C" int ldpc_decode(const ldpc_decoder_t *d, const float *llr_dev, int64_t B, int32_t iters, int32_t early_term,
                           float *app_dev, int32_t app_all_iters, uint32_t *hard_dev, int32_t *iters_dev,
                           uint8_t *flags_dev, int32_t *biterr_dev, void *stream) {
    if (!llr_dev && B > 0) return fail(LDPC_E_INVALID, "decode: bad arguments");
    return decode_dev(d, llr_dev, nullptr, 0.0f, B, iters, early_term, app_dev, app_all_iters, hard_dev, iters_dev,
                      flags_dev, biterr_dev, nullptr, stream);
}

namespace { void fill_channel(KParams &P, double sigma, uint64_t seed, uint64_t frame_offset); }

// ---- non-zero codewords ("next" row N3: Print_Functions.py:40-46, 100-118 with Y != 0)
extern "C" int ldpc_llr_generate_cw(const ldpc_decoder_t *d, double sigma, int64_t n_frames, uint64_t seed, uint64_t frame_offset,
                                    const uint32_t *codeword_bits_dev, int64_t cw_stride_words, float *llr_dev, void *stream) {
    if (!d || !llr_dev || !codeword_bits_dev || n_frames < 0 || cw_stride_words < 0 || !(sigma > 0.0))
        return fail(LDPC_E_INVALID, "llr_generate_cw: bad arguments");
    if (n_frames == 0) return LDPC_OK;
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    KParams P = d->base;
    fill_channel(P, sigma, seed, frame_offset);
    CUDA_TRY(nms_launch_generate_cw(P, codeword_bits_dev, cw_stride_words, llr_dev, n_frames, (cudaStream_t)stream));
    return LDPC_OK;
}

extern "C" int ldpc_decode_cw(const ldpc_decoder_t *dc, const float *llr_dev, const uint32_t *codeword_bits_dev,
                              int64_t cw_stride_words, int64_t B, int32_t iters, int32_t early_term, uint32_t *hard_dev,
                              int32_t *iters_dev, uint8_t *flags_dev, int32_t *biterr_dev, int32_t *biterr_signed_dev,
                              uint64_t *counters_dev, void *stream) {
    ldpc_decoder *d = const_cast<ldpc_decoder *>(dc);
    if (!d || (!llr_dev && B > 0) || !codeword_bits_dev || B < 0 || cw_stride_words < 0)
        return fail(LDPC_E_INVALID, "decode_cw: bad arguments");
    if (iters < 0 || iters > d->T) return fail(LDPC_E_INVALID, "decode_cw: iters %d outside 0..%d", iters, d->T);
    if (B == 0) return LDPC_OK;
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    std::lock_guard<std::mutex> lock(d->mu);
    const KParams &P0 = d->base;
    const int T_run = iters == 0 ? d->T : iters;
    cudaStream_t st = (cudaStream_t)stream;
    // chunks of frames whose [T_run, chunk, N*z] APP tensor stays below 1 GiB
    const size_t per_frame = (size_t)T_run * P0.NZ * sizeof(float);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)((1ull << 30) / per_frame)));
    if (d->cw_app_cap < (size_t)chunk * per_frame) {
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(d->d_cw_app); d->d_cw_app = nullptr; d->cw_app_cap = 0;
        CUDA_TRY(cudaMalloc(&d->d_cw_app, (size_t)chunk * per_frame));
        d->cw_app_cap = (size_t)chunk * per_frame;
    }
    if (d->cw_frames_cap < (size_t)chunk) {
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(d->d_cw_iters); cudaFree(d->d_cw_flags); d->d_cw_iters = nullptr; d->d_cw_flags = nullptr; d->cw_frames_cap = 0;
        CUDA_TRY(cudaMalloc(&d->d_cw_iters, (size_t)chunk * sizeof(int)));
        CUDA_TRY(cudaMalloc(&d->d_cw_flags, (size_t)chunk));
        d->cw_frames_cap = (size_t)chunk;
    }
    for (int64_t off = 0; off < B; off += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, B - off);
        int32_t *it = iters_dev ? iters_dev + off : d->d_cw_iters;
        uint8_t *fl = flags_dev ? flags_dev + off : d->d_cw_flags;
        int rc = decode_dev(d, llr_dev + off * P0.NZ, nullptr, 0.0f, nb, iters, early_term, d->d_cw_app, 1,
                            hard_dev ? hard_dev + off * P0.HW : nullptr, it, fl, nullptr, nullptr, st);
        if (rc != LDPC_OK) return rc;
        CUDA_TRY(nms_launch_cw_metrics(d->d_cw_app, nb, T_run, P0.NZ, P0.target_n * P0.z,
                                       codeword_bits_dev + off * cw_stride_words, cw_stride_words, it, early_term ? 1 : 0, fl,
                                       biterr_dev ? biterr_dev + off : nullptr, biterr_signed_dev ? biterr_signed_dev + off : nullptr,
                                       (unsigned long long *)counters_dev, st));
    }
    return LDPC_OK;
}

extern "C" float ldpc_decoder_q8_step(const ldpc_decoder_t *d) { return d ? quantiser_step(d) : 0.0f; }

extern "C" int ldpc_decode_q8(const ldpc_decoder_t *d, const int8_t *llr_q8_dev, float step, int64_t B, int32_t iters,
                              int32_t early_term, uint32_t *hard_dev, int32_t *iters_dev, uint8_t *flags_dev,
                              int32_t *biterr_dev, uint64_t *counters_dev, void *stream) {
    if (!d || (!llr_q8_dev && B > 0)) return fail(LDPC_E_INVALID, "decode_q8: bad arguments");
    if (step == 0.0f) step = quantiser_step(d);
    return decode_dev(d, nullptr, llr_q8_dev, step, B, iters, early_term, nullptr, 0, hard_dev, iters_dev, flags_dev,
                      biterr_dev, counters_dev, stream);
}

namespace {
int ensure_host_scratch(ldpc_decoder *d, size_t chunk, bool with_app, int app_iters) {
    HostScratch &h = d->hs;
    if (h.cap_frames >= chunk && (!with_app || (h.with_app && h.app_iters >= app_iters))) return LDPC_OK;
    unsigned long long *cnt = h.counters; unsigned int *uc = h.ucount; float *ub = h.ubuf; size_t ur = h.ubuf_rows;
    h.counters = nullptr; h.ucount = nullptr; h.ubuf = nullptr;
    free_scratch(h);
    h.counters = cnt; h.ucount = uc; h.ubuf = ub; h.ubuf_rows = ur;
    const KParams &P = d->base;
    for (int i = 0; i < HOST_SLOTS; ++i) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h.st[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h.ev_h2d[i], cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc(&h.llr[i], chunk * P.NZ * sizeof(float)));
        CUDA_TRY(cudaMalloc(&h.hard[i], chunk * P.HW * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&h.iters[i], chunk * sizeof(int)));
        CUDA_TRY(cudaMalloc(&h.flags[i], chunk));
        CUDA_TRY(cudaMalloc(&h.biterr[i], chunk * sizeof(int)));
        for (int j = i; j < 2 * HOST_SLOTS; j += HOST_SLOTS) {
            CUDA_TRY(cudaMallocHost(&h.h_hard[j], chunk * P.HW * sizeof(uint32_t)));
            CUDA_TRY(cudaMallocHost(&h.h_iters[j], chunk * sizeof(int)));
            CUDA_TRY(cudaMallocHost(&h.h_flags[j], chunk));
            CUDA_TRY(cudaMallocHost(&h.h_biterr[j], chunk * sizeof(int)));
        }
        if (with_app) CUDA_TRY(cudaMalloc(&h.app[i], (size_t)app_iters * chunk * P.NZ * sizeof(float)));
    }
    h.cap_frames = chunk; h.with_app = with_app; h.app_iters = app_iters;
    return LDPC_OK;
}

int ensure_host_staging(HostScratch &h, size_t bytes) {
    if (h.h_in_bytes >= bytes) return LDPC_OK;
    for (int i = 0; i < HOST_SLOTS; ++i) { cudaFreeHost(h.h_in[i]); h.h_in[i] = nullptr; }
    h.h_in_bytes = 0;
    for (int i = 0; i < HOST_SLOTS; ++i) CUDA_TRY(cudaMallocHost(&h.h_in[i], bytes));
    h.h_in_bytes = bytes;
    return LDPC_OK;
}

// Can float32 words of this decoder cross PCIe as int8?  0: no; 1: always (the decoder sees the channel value only through
// Q(x)); 2: when the words are on the quantiser grid already (VN weights also form Q(x * w) from the raw value).
int host_pack_mode(const ldpc_decoder *d, float *qk, float *kmax) {
    if (d->decoding_type != 2 || d->q_bit == 6 || env_on("LDPC_B200_NO_HOST_PACK")) return 0;   // q_bit 6 saturates at 15.5: off its grid
    const float step = d->q_bit == 5 ? 0.5f : (d->q_bit == 3 ? 2.0f : 1.0f);
    *qk = 1.0f / step;
    *kmax = d->base.qmax / step;
    return d->sharing[2] == 0 ? 1 : 2;
}

double now_s() {
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

bool host_ptr_pageable(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return attr.type == cudaMemoryTypeUnregistered;
}

// The host-buffer pipeline.  A call is cut into chunks (a few waves of CTAs, ~32 MiB of float32 words; the first ones
// shorter so the device starts early).  Up to HOST_SLOTS chunks are in flight, each on a slot of its own (device buffers,
// stream, two sets of pinned result buffers).  float32 words of a quantised decoder cross PCIe in one of two forms:
//   * as they are (DMA straight from the caller's pinned memory: no host work, 4 bytes per value), or
//   * packed to int8 by the host threads (host_pack.cpp: 1 byte per value) -- the same decoder input bit for bit.
// PCIe bounds the first form (the decoder is ~2x faster than 55 GB/s of float32 words), the host's cores and memory the
// second, so the two work side by side on ONE list of chunks taken from both ends:
//   * a FEEDER thread claims chunks from the front, packs each (with the pool) into the next buffer of a ring of pinned
//     staging buffers -- free again as soon as the copy that last read it has completed: an event, not a stream -- and
//     publishes it; between chunks it copies finished results into the caller's arrays;
//   * the CALLING thread waits for a free slot, then issues copy + kernel + result copies for the next published chunk
//     or, when none is ready, for a chunk claimed from the BACK of the list, sent as float32.
// The two lanes meet wherever their speeds put them: no tuning, and whichever resource is scarce on the box (PCIe at one
// GPU, cores and host memory when eight GPUs share them) sets the split.  Pageable input has no float32 lane worth having
// (reading it is the cost either way): the feeder packs every chunk, or stages the ones without an int8 form through the
// same pinned buffers.
struct HostFeed {
    enum : int { PENDING = 0, DIRECT, STAGED_F32, STAGED_Q8, FAILED };
    std::mutex claim_mu;
    int front = 0, back = 0;                        // chunks [front, back) are unclaimed
    std::vector<std::atomic<int>> state;            // per front chunk: how the calling thread finds its words
    std::atomic<int> prepared{0};                   // front chunks published so far (chunks 0 .. prepared-1, in order)
    std::atomic<bool> front_done{false};            // the feeder claims no more chunks
    std::atomic<long long> issued[HOST_SLOTS];      // copies issued from staging buffer r so far
    std::atomic<bool> abort{false}, failed{false};
    int chunks_q8 = 0, chunks_unencodable = 0;
    double s_pack = 0.0, s_copy_out = 0.0;
    // results ready for the caller, in issue order: the calling thread queues them, the feeder copies them out with the pool
    struct Results { int64_t off, n; int buf; } rq[32];
    std::atomic<long long> rq_head{0}, rq_tail{0};
    std::atomic<bool> rq_closed{false};
    explicit HostFeed(size_t n) : back((int)n), state(n) {
        for (auto &x : state) x.store(PENDING, std::memory_order_relaxed);
        for (auto &x : issued) x.store(0, std::memory_order_relaxed);
    }
    int claim_front() { std::lock_guard<std::mutex> g(claim_mu); return front < back ? front++ : -1; }
    int claim_back() { std::lock_guard<std::mutex> g(claim_mu); return back > front ? --back : -1; }
};

int decode_host_impl(const ldpc_decoder_t *dc, const void *src_host, bool q8, float step, int64_t B, int32_t iters,
                     int32_t early_term, float *app_host, int32_t app_all_iters, uint32_t *hard_host,
                     int32_t *iters_host, uint8_t *flags_host, int32_t *biterr_host) {
    ldpc_decoder *d = const_cast<ldpc_decoder *>(dc);
    if (!d || (!src_host && B > 0) || B < 0) return fail(LDPC_E_INVALID, "decode_host: bad arguments");
    if (q8 && step == 0.0f) step = quantiser_step(d);
    if (q8 && !(step > 0.0f)) return fail(LDPC_E_INVALID, "decode_q8_host: step must be positive for a float decoder");
    if (iters < 0 || iters > d->T) return fail(LDPC_E_INVALID, "decode_host: iters %d outside 0..%d", iters, d->T);
    if (B == 0) return LDPC_OK;
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    std::lock_guard<std::mutex> lock(d->mu);
    const double t_begin = now_s();
    const KParams &P0 = d->base;
    const size_t NZ = (size_t)P0.NZ;
    const int T_run = iters == 0 ? d->T : iters;
    const int app_iters = app_host ? (app_all_iters ? T_run : 1) : 0;
    const bool use_alt = d->func_alt != nullptr && !early_term;
    const LaunchGeom &geo = use_alt ? d->geom_alt : d->geom;
    const int FBg = use_alt ? d->base_alt.FB : P0.FB;
    const size_t wave = (size_t)d->sm_count * geo.ctas_per_sm * FBg;
    size_t chunk = std::max<size_t>(wave, ((32u << 20) / (NZ * sizeof(float)) / wave) * wave);
    if (app_iters) chunk = wave;
    chunk = std::min<size_t>(chunk, ((size_t)B + FBg - 1) / FBg * FBg);
    int rc = ensure_host_scratch(d, chunk, app_iters > 0, app_iters);
    if (rc != LDPC_OK) return rc;
    HostScratch &h = d->hs;

    float qk = 1.0f, kmax = 0.0f;
    int pack_mode = (q8 || app_iters) ? 0 : host_pack_mode(d, &qk, &kmax);
    const bool pageable = host_ptr_pageable(src_host);
    // with one or two host threads the packing lane adds less than its bookkeeping costs next to a DMA engine that is
    // already busy (measured: 22.7 M frames/s with one thread against 23.0 as float32 only; 28.1 with three)
    if (pack_mode && !pageable && hostpack::pool_threads() < 3) pack_mode = 0;
    const size_t elem = q8 ? 1 : sizeof(float);
    const bool front_lane = pack_mode != 0 || pageable;   // chunks that need the host's hands
    const bool back_lane = !pageable;                     // chunks the DMA engine can fetch from the caller's memory
    if (front_lane) {
        rc = ensure_host_staging(h, chunk * NZ * ((pageable && !q8) ? sizeof(float) : 1));
        if (rc != LDPC_OK) return rc;
    }
    // the chunks of this call
    std::vector<int64_t> c_off, c_n;
    for (int64_t off = 0; off < B;) {
        const int64_t nb = std::min<int64_t>((int64_t)std::min(chunk, wave * (c_off.size() + 1)), B - off);
        c_off.push_back(off); c_n.push_back(nb);
        off += nb;
    }
    const int nchunks = (int)c_off.size();
    ldpc_host_stats_t stats{};
    HostFeed feed((size_t)nchunks);
    stats.threads = front_lane ? hostpack::pool_threads() : 1;

    // results of a chunk: copied by its stream into the pinned staging buffers (slot, parity), from there into the caller's
    // arrays once the stream has drained
    auto copy_results = [&](int64_t off, int64_t nb, int b, bool mt) {
        if (nb <= 0) return;
        if (hard_host) {
            const size_t bytes = (size_t)nb * P0.HW * 4;
            if (mt && bytes >= (256u << 10)) hostpack::memcpy_mt(hard_host + off * P0.HW, h.h_hard[b], bytes);
            else std::memcpy(hard_host + off * P0.HW, h.h_hard[b], bytes);
        }
        if (iters_host) std::memcpy(iters_host + off, h.h_iters[b], (size_t)nb * 4);
        if (flags_host) std::memcpy(flags_host + off, h.h_flags[b], (size_t)nb);
        if (biterr_host) std::memcpy(biterr_host + off, h.h_biterr[b], (size_t)nb * 4);
    };
    auto feeder_copy_out = [&]() {   // feeder thread: everything the calling thread has queued so far
        long long t = feed.rq_tail.load(std::memory_order_relaxed);
        while (t < feed.rq_head.load(std::memory_order_acquire)) {
            const HostFeed::Results r = feed.rq[t % 32];
            const double t0 = now_s();
            copy_results(r.off, r.n, r.buf, true);
            feed.s_copy_out += now_s() - t0;
            feed.rq_tail.store(++t, std::memory_order_release);
        }
    };

    // ---- feeder: front chunk c -> (form, where its words are); between chunks, results -> caller
    auto feeder = [&](bool own_thread) {
        if (cudaSetDevice(d->device) != cudaSuccess) { cudaGetLastError(); }
        int unencodable_run = 0;
        for (;;) {
            if (own_thread) feeder_copy_out();
            // pinned words that keep failing the "has an int8 form" test: leave the rest to the float32 lane
            if (!pageable && (pack_mode == 0 || unencodable_run >= 2)) break;
            const int c = feed.claim_front();
            if (c < 0) break;
            const int r = c % HOST_SLOTS;
            const char *src = (const char *)src_host + (size_t)c_off[c] * NZ * elem;
            const size_t nval = (size_t)c_n[c] * NZ;
            // staging buffer r is free once the copy that last read it (front chunk c - HOST_SLOTS) has completed
            const long long need = c / HOST_SLOTS;
            for (int spin = 0; feed.issued[r].load(std::memory_order_acquire) < need; ++spin) {
                if (feed.abort.load(std::memory_order_relaxed)) return;
                if (own_thread) feeder_copy_out();
                if (spin > 200) std::this_thread::yield();
            }
            if (need > 0 && cudaEventSynchronize(h.ev_h2d[r]) != cudaSuccess) {
                cudaGetLastError();
                feed.failed.store(true, std::memory_order_release);
                feed.state[c].store(HostFeed::FAILED, std::memory_order_release);
                feed.prepared.store(c + 1, std::memory_order_release);
                return;
            }
            const double t0 = now_s();
            int result = HostFeed::DIRECT;
            if (pack_mode && unencodable_run < 2) {
                const int64_t bad = hostpack::pack_q8_mt((const float *)src, (int64_t)nval, qk, kmax, pack_mode == 2,
                                                         (int8_t *)h.h_in[r], true);
                if (bad == 0) { result = HostFeed::STAGED_Q8; unencodable_run = 0; ++feed.chunks_q8; }
                else { ++unencodable_run; ++feed.chunks_unencodable; }
            }
            if (result == HostFeed::DIRECT && pageable) {
                hostpack::memcpy_mt(h.h_in[r], src, nval * elem);
                result = q8 ? HostFeed::STAGED_Q8 : HostFeed::STAGED_F32;
            }
            feed.s_pack += now_s() - t0;
            feed.state[c].store(result, std::memory_order_relaxed);
            feed.prepared.store(c + 1, std::memory_order_release);
        }
        feed.front_done.store(true, std::memory_order_release);
        if (!own_thread) return;
        for (int spin = 0;; ++spin) {   // results, until the calling thread closes the queue
            feeder_copy_out();
            if (feed.rq_closed.load(std::memory_order_acquire) &&
                feed.rq_tail.load(std::memory_order_relaxed) >= feed.rq_head.load(std::memory_order_acquire)) return;
            if (feed.abort.load(std::memory_order_relaxed)) return;
            if (spin > 200) std::this_thread::yield();
        }
    };
    std::thread feeder_thread;
    struct Joiner {
        std::thread &t; HostFeed &f;
        ~Joiner() { f.abort.store(true); if (t.joinable()) t.join(); }
    } joiner{feeder_thread, feed};
    if (front_lane && nchunks > 1) {
        try {
            feeder_thread = std::thread(feeder, true);
        } catch (...) {
            return fail(LDPC_E_ALLOC, "decode_host: cannot start the feeder thread");
        }
    } else if (front_lane) {
        feeder(false);                   // one chunk: prepared right here
    } else {
        feed.front_done.store(true);     // nothing needs preparing: every chunk goes the direct way, in order
    }

    struct Pending { int64_t off = 0, n = 0; int buf = 0; } pend[HOST_SLOTS];
    auto wait_slot = [&](int k) -> int {
        const double t0 = now_s();
        CUDA_TRY(cudaStreamSynchronize(h.st[k]));
        stats.s_wait += now_s() - t0;
        return LDPC_OK;
    };
    const bool feeder_copies = feeder_thread.joinable();
    long long handed = 0;   // results handed over (queued or copied) so far, in issue order
    auto hand_over = [&](const Pending &p) {
        if (p.n <= 0) return;
        if (feeder_copies) {
            feed.rq[handed % 32] = HostFeed::Results{p.off, p.n, p.buf};
            feed.rq_head.store(handed + 1, std::memory_order_release);
        } else {
            const double t1 = now_s();
            copy_results(p.off, p.n, p.buf, false);
            stats.s_copy_out += now_s() - t1;
        }
        ++handed;
    };

    bool slot_back[HOST_SLOTS] = {};   // the slot's chunk in flight came from the back of the list (float32 as it is)
    int pf = 0;          // next published front chunk to issue
    int direct_next = 0; // without a feeder: chunks in order
    for (int i = 0;; ++i) {   // i-th chunk issued by this call
        const int k = i % HOST_SLOTS;
        rc = wait_slot(k);
        if (rc != LDPC_OK) return rc;
        const Pending done = pend[k];          // complete now; handed over below, after the next chunk has been issued
        pend[k] = Pending();
        // ---- which chunk: a published one, else one from the back of the list as it is, else wait for the feeder
        int c = -1, form = HostFeed::DIRECT;
        bool from_front = false;
        const double tf0 = now_s();
        slot_back[k] = false;
        int back_in_flight = 0;
        for (int j = 0; j < HOST_SLOTS; ++j) back_in_flight += slot_back[j] ? 1 : 0;
        for (int spin = 0;; ++spin) {
            if (!front_lane) { c = direct_next < nchunks ? direct_next++ : -1; break; }
            const int prepared = feed.prepared.load(std::memory_order_acquire);
            if (pf < prepared) {
                c = pf++;
                form = feed.state[c].load(std::memory_order_relaxed);
                from_front = true;
                break;
            }
            const bool done_front = feed.front_done.load(std::memory_order_acquire);
            if (done_front && pf >= feed.prepared.load(std::memory_order_acquire)) {
                if (back_lane) c = feed.claim_back();
                break;                                   // c < 0: the list is empty
            }
            // two float32 chunks in flight keep the DMA engine busy (one copying, one queued); more would only queue in
            // front of the packed ones (three in flight: 55.9 instead of 60.7 M frames/s on two GPUs).  Nothing from the
            // back before the feeder's first (short) chunk either: the device would start later.
            if (back_lane && prepared > 0 && back_in_flight < 2 && (c = feed.claim_back()) >= 0) break;
            if (back_in_flight >= 2 && (spin & 63) == 63) {   // still in flight?
                back_in_flight = 0;
                for (int j = 0; j < HOST_SLOTS; ++j) {
                    if (slot_back[j] && cudaStreamQuery(h.st[j]) == cudaSuccess) slot_back[j] = false;
                    back_in_flight += slot_back[j] ? 1 : 0;
                }
                cudaGetLastError();   // cudaErrorNotReady is not an error here
            }
            if (spin > 200) std::this_thread::yield();
        }
        stats.s_wait_feed += now_s() - tf0;
        if (c < 0) { hand_over(done); break; }
        if (form == HostFeed::FAILED) return fail(LDPC_E_CUDA, "decode_host: waiting for a staging buffer failed");
        const int64_t off = c_off[c], nb = c_n[c];
        cudaStream_t st = h.st[k];
        const bool staged = form == HostFeed::STAGED_F32 || form == HostFeed::STAGED_Q8;
        const bool as_q8 = q8 || form == HostFeed::STAGED_Q8;
        const int r = c % HOST_SLOTS;   // staging buffer of a front chunk
        const char *src = staged ? h.h_in[r] : (const char *)src_host + (size_t)off * NZ * elem;
        const size_t bytes = (size_t)nb * NZ * (as_q8 ? 1 : sizeof(float));
        CUDA_TRY(cudaMemcpyAsync(h.llr[k], src, bytes, cudaMemcpyHostToDevice, st));
        if (from_front) {   // its staging buffer (used or not) may go to front chunk c + HOST_SLOTS once this copy is through
            CUDA_TRY(cudaEventRecord(h.ev_h2d[r], st));
            feed.issued[r].fetch_add(1, std::memory_order_release);
        } else {
            slot_back[k] = front_lane;
        }
        stats.h2d_bytes += (int64_t)bytes;
        KParams P = P0;
        P.T_run = T_run; P.early_term = early_term ? 1 : 0;
        P.llr = as_q8 ? nullptr : h.llr[k]; P.llr_q8 = as_q8 ? (const signed char *)h.llr[k] : nullptr;
        P.q8_step = q8 ? step : 1.0f / qk;
        P.n_frames = nb;
        P.app = app_iters ? h.app[k] : nullptr; P.app_all = app_all_iters ? 1 : 0; P.app_stride_t = (long long)nb * P.NZ;
        P.hard = hard_host ? h.hard[k] : nullptr; P.iters = iters_host ? h.iters[k] : nullptr;
        P.flags = flags_host ? h.flags[k] : nullptr; P.biterr = biterr_host ? h.biterr[k] : nullptr;
        rc = launch(d, P, st);
        if (rc != LDPC_OK) return rc;
        const int rb = k + HOST_SLOTS * ((i / HOST_SLOTS) & 1);
        if (feeder_copies && i >= 2 * HOST_SLOTS) {   // the (i - 8)-th chunk last used these result buffers: hand-over job i - 8
            const double t0 = now_s();
            for (int spin = 0; feed.rq_tail.load(std::memory_order_acquire) < (long long)(i - 2 * HOST_SLOTS + 1); ++spin) {
                if (feed.failed.load(std::memory_order_acquire)) return fail(LDPC_E_CUDA, "decode_host: the feeder thread failed");
                if (spin > 200) std::this_thread::yield();
            }
            stats.s_wait_feed += now_s() - t0;
        }
        if (hard_host) CUDA_TRY(cudaMemcpyAsync(h.h_hard[rb], h.hard[k], (size_t)nb * P.HW * 4, cudaMemcpyDeviceToHost, st));
        if (iters_host) CUDA_TRY(cudaMemcpyAsync(h.h_iters[rb], h.iters[k], (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        if (flags_host) CUDA_TRY(cudaMemcpyAsync(h.h_flags[rb], h.flags[k], (size_t)nb, cudaMemcpyDeviceToHost, st));
        if (biterr_host) CUDA_TRY(cudaMemcpyAsync(h.h_biterr[rb], h.biterr[k], (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        stats.d2h_bytes += (int64_t)nb * ((hard_host ? P.HW * 4 : 0) + (iters_host ? 4 : 0) + (flags_host ? 1 : 0) + (biterr_host ? 4 : 0));
        pend[k].off = off; pend[k].n = nb; pend[k].buf = rb;
        if (app_iters) {
            const size_t row = (size_t)nb * P.NZ * sizeof(float);
            CUDA_TRY(cudaMemcpy2DAsync(app_host + off * P.NZ, (size_t)B * P.NZ * sizeof(float), h.app[k], row, row,
                                       (size_t)app_iters, cudaMemcpyDeviceToHost, st));
            stats.d2h_bytes += (int64_t)(row * app_iters);
        }
        hand_over(done);
        ++stats.chunks_total;
    }
    for (int k = 0; k < HOST_SLOTS; ++k) {
        rc = wait_slot(k);
        if (rc != LDPC_OK) return rc;
    }
    // the chunks still in flight when the list ran out, oldest first
    {
        int order[HOST_SLOTS];
        for (int k = 0; k < HOST_SLOTS; ++k) order[k] = k;
        std::sort(order, order + HOST_SLOTS, [&](int x, int y) { return pend[x].off < pend[y].off; });
        for (int j = 0; j < HOST_SLOTS; ++j) { hand_over(pend[order[j]]); pend[order[j]] = Pending(); }
    }
    feed.rq_closed.store(true, std::memory_order_release);
    if (feeder_thread.joinable()) feeder_thread.join();
    if (feed.failed.load()) return fail(LDPC_E_CUDA, "decode_host: the feeder thread failed");
    if (feeder_copies) stats.s_copy_out = feed.s_copy_out;
    stats.chunks_q8 = q8 ? stats.chunks_total : feed.chunks_q8;
    stats.chunks_unencodable = feed.chunks_unencodable;
    stats.chunks_f32 = stats.chunks_total - stats.chunks_q8 - stats.chunks_unencodable;
    stats.s_pack = feed.s_pack;
    stats.float_share = stats.chunks_total ? (double)(stats.chunks_total - stats.chunks_q8) / stats.chunks_total : 0.0;
    stats.s_total = now_s() - t_begin;
    h.stats = stats;
    return LDPC_OK;
}
}   // namespace

extern "C" int ldpc_decode_host_stats(const ldpc_decoder_t *d, ldpc_host_stats_t *out) {
    if (!d || !out) return fail(LDPC_E_INVALID, "decode_host_stats: bad arguments");
    std::lock_guard<std::mutex> lock(const_cast<ldpc_decoder *>(d)->mu);
    *out = d->hs.stats;
    return LDPC_OK;
}

extern "C" int ldpc_pack_q8_values(const float *x, int64_t n, float step, float qmax, int32_t lossless, int8_t *q8,
                                   int64_t *n_unencodable) {
    if (n < 0 || (n > 0 && (!x || !q8)) || !(step > 0.0f) || !(qmax > 0.0f) || qmax / step > 127.0f)
        return fail(LDPC_E_INVALID, "pack_q8_values: bad arguments");
    const int64_t bad = hostpack::pack_q8_mt(x, n, 1.0f / step, qmax / step, lossless ? 1 : 0, q8);
    if (n_unencodable) *n_unencodable = bad;
    return LDPC_OK;
}

extern "C" int ldpc_pack_q8_host(const ldpc_decoder_t *d, const float *llr_host, int64_t B, int8_t *q8_host,
                                 int64_t *n_unencodable) {
    if (!d || B < 0 || (B > 0 && (!llr_host || !q8_host))) return fail(LDPC_E_INVALID, "pack_q8_host: bad arguments");
    float qk = 1.0f, kmax = 0.0f;
    const int mode = host_pack_mode(d, &qk, &kmax);
    if (mode == 0) return fail(LDPC_E_UNSUPPORTED, "pack_q8_host: this decoder has no int8 word form (float decoder, or q_bit 6)");
    const int64_t bad = hostpack::pack_q8_mt(llr_host, B * (int64_t)d->base.NZ, qk, kmax, mode == 2, q8_host);
    if (n_unencodable) *n_unencodable = bad;
    return LDPC_OK;
}

extern "C" int ldpc_decode_host(const ldpc_decoder_t *d, const float *llr_host, int64_t B, int32_t iters,
                                int32_t early_term, float *app_host, int32_t app_all_iters, uint32_t *hard_host,
                                int32_t *iters_host, uint8_t *flags_host, int32_t *biterr_host) {
    return decode_host_impl(d, llr_host, false, 0.0f, B, iters, early_term, app_host, app_all_iters, hard_host,
                            iters_host, flags_host, biterr_host);
}

extern "C" int ldpc_decode_q8_host(const ldpc_decoder_t *d, const int8_t *llr_q8_host, float step, int64_t B,
                                   int32_t iters, int32_t early_term, uint32_t *hard_host, int32_t *iters_host,
                                   uint8_t *flags_host, int32_t *biterr_host) {
    return decode_host_impl(d, llr_q8_host, true, step, B, iters, early_term, nullptr, 0, hard_host, iters_host,
                            flags_host, biterr_host);
}

// ------------------------------------------------------------------- generator / Monte-Carlo
namespace {
void fill_channel(KParams &P, double sigma, uint64_t seed, uint64_t frame_offset) {
    P.sigma = (float)sigma;
    P.two_over_s2 = (float)(2.0 / (sigma * sigma));
    P.two_over_s = (float)(2.0 / sigma);
    for (int r = 0; r < 10; ++r) {
        P.pkeys[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        P.pkeys[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    P.seed = seed; P.frame_offset = frame_offset;
}
}   // namespace

extern "C" int ldpc_llr_generate(const ldpc_decoder_t *d, double sigma, int64_t n_frames, uint64_t seed,
                                 uint64_t frame_offset, float *llr_dev, void *stream) {
    if (!d || !llr_dev || n_frames < 0 || !(sigma > 0.0)) return fail(LDPC_E_INVALID, "llr_generate: bad arguments");
    if (n_frames == 0) return LDPC_OK;
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    KParams P = d->base;
    fill_channel(P, sigma, seed, frame_offset);
    CUDA_TRY(nms_launch_generate(P, llr_dev, n_frames, (cudaStream_t)stream));
    return LDPC_OK;
}

extern "C" int ldpc_normal_probe(int32_t device, uint64_t seed, uint64_t frame_offset, int64_t n_frames, int32_t quads_per_frame,
                                 float *normals_dev, uint64_t *tail_counts_dev, void *stream) {
    if (n_frames < 0 || quads_per_frame <= 0 || (!normals_dev && !tail_counts_dev))
        return fail(LDPC_E_INVALID, "normal_probe: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(LDPC_E_CUDA, "no CUDA device"); }
    if (device < 0 || device >= ndev) return fail(LDPC_E_INVALID, "device %d out of range", device);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", device);
    CUDA_TRY(nms_launch_normal_probe(seed, frame_offset, n_frames, quads_per_frame, normals_dev,
                                     (unsigned long long *)tail_counts_dev, (cudaStream_t)stream));
    return LDPC_OK;
}

extern "C" int ldpc_mc_run(const ldpc_decoder_t *d, double sigma, int64_t n_frames, uint64_t seed, uint64_t frame_offset,
                           int32_t iters, int32_t early_term, int32_t harvest_mode, uint64_t *counters_dev,
                           float *uncor_buf_dev, uint32_t *uncor_count_dev, uint32_t uncor_capacity, void *stream) {
    if (!d || n_frames < 0 || !(sigma > 0.0) || !counters_dev) return fail(LDPC_E_INVALID, "mc_run: bad arguments");
    if (iters < 0 || iters > d->T) return fail(LDPC_E_INVALID, "mc_run: iters %d outside 0..%d", iters, d->T);
    if (harvest_mode < 0 || harvest_mode > 3) return fail(LDPC_E_INVALID, "mc_run: harvest_mode %d", harvest_mode);
    if (n_frames == 0) return LDPC_OK;
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    KParams P = d->base;
    fill_channel(P, sigma, seed, frame_offset);
    P.T_run = iters == 0 ? d->T : iters;
    P.early_term = early_term ? 1 : 0;
    P.llr = nullptr; P.n_frames = n_frames;
    P.counters = (unsigned long long *)counters_dev;
    P.harvest_mode = harvest_mode;
    P.uncor_buf = uncor_buf_dev; P.uncor_count = uncor_count_dev; P.uncor_cap = uncor_capacity;
    return launch(d, P, (cudaStream_t)stream);
}

extern "C" int ldpc_mc_run_staged(const ldpc_decoder_t *d, double sigma, int64_t n_frames, uint64_t seed, uint64_t frame_offset,
                                  int32_t iters, int32_t stage1_iters, int32_t harvest_mode, uint64_t *counters_dev,
                                  float *uncor_buf_dev, uint32_t *uncor_count_dev, uint32_t uncor_capacity,
                                  uint64_t *defer_list_dev, uint32_t *defer_count_dev, uint32_t defer_capacity, void *stream) {
    if (!d || n_frames < 0 || !(sigma > 0.0) || !counters_dev || !defer_list_dev || !defer_count_dev)
        return fail(LDPC_E_INVALID, "mc_run_staged: bad arguments");
    const int T = iters == 0 ? d->T : iters;
    if (iters < 0 || iters > d->T) return fail(LDPC_E_INVALID, "mc_run_staged: iters %d outside 0..%d", iters, d->T);
    if (stage1_iters < 1 || stage1_iters >= T) return fail(LDPC_E_INVALID, "mc_run_staged: stage1_iters %d outside 1..%d", stage1_iters, T - 1);
    if (harvest_mode < 0 || harvest_mode > 3) return fail(LDPC_E_INVALID, "mc_run_staged: harvest_mode %d", harvest_mode);
    if (n_frames == 0) return LDPC_OK;
    // a graph with a persistent-slot kernel needs no staging: its slots are refilled frame by frame (same counters)
    if (defer_capacity == 0) return fail(LDPC_E_INVALID, "mc_run_staged: defer_capacity is 0");
    if (d->func_mc != nullptr && !env_on("LDPC_B200_NO_PERSIST"))
        return ldpc_mc_run(d, sigma, n_frames, seed, frame_offset, iters, 1, harvest_mode, counters_dev, uncor_buf_dev,
                           uncor_count_dev, uncor_capacity, stream);
    if (n_frames > 0xffffffffLL) return fail(LDPC_E_LIMIT, "mc_run_staged: more than 2^32 - 1 frames per call");
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(defer_count_dev, 0, sizeof(uint32_t), st));
    KParams P = d->base;
    fill_channel(P, sigma, seed, frame_offset);
    P.early_term = 1;
    P.llr = nullptr;
    P.counters = (unsigned long long *)counters_dev;
    P.harvest_mode = harvest_mode;
    P.uncor_buf = uncor_buf_dev; P.uncor_count = uncor_count_dev; P.uncor_cap = uncor_capacity;
    // stage 1: every frame, stage1_iters iterations; frames without a zero syndrome by then go to the list
    P.T_run = stage1_iters; P.n_frames = n_frames;
    P.defer_list = (unsigned long long *)defer_list_dev; P.defer_count = defer_count_dev; P.defer_cap = defer_capacity;
    int rc = launch(d, P, st);
    if (rc != LDPC_OK) return rc;
    uint32_t n2 = 0;
    CUDA_TRY(cudaMemcpyAsync(&n2, defer_count_dev, sizeof n2, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (n2 == 0) return LDPC_OK;
    if (n2 > defer_capacity)
        return fail(LDPC_E_LIMIT, "mc_run_staged: %u frames deferred, defer_list holds %u (counters already hold stage 1: "
                                  "zero them and call again with a larger list)", n2, defer_capacity);
    // stage 2: the listed frames regenerated from their global indices, all `iters` iterations (early termination on)
    P.T_run = T; P.n_frames = (long long)n2;
    P.frame_list = (const unsigned long long *)defer_list_dev; P.defer_list = nullptr; P.defer_count = nullptr;
    return launch(d, P, st);
}

extern "C" int ldpc_mc_run_host(const ldpc_decoder_t *dc, double sigma, int64_t n_frames, uint64_t seed,
                                uint64_t frame_offset, int32_t iters, int32_t early_term, int32_t harvest_mode,
                                uint64_t *counters_host, float *uncor_host, uint32_t uncor_capacity,
                                uint32_t *n_uncor_host) {
    ldpc_decoder *d = const_cast<ldpc_decoder *>(dc);
    if (!d || !counters_host) return fail(LDPC_E_INVALID, "mc_run_host: bad arguments");
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    std::lock_guard<std::mutex> lock(d->mu);
    HostScratch &h = d->hs;
    if (!h.counters) CUDA_TRY(cudaMalloc(&h.counters, LDPC_NUM_COUNTERS * sizeof(unsigned long long)));
    if (!h.ucount) CUDA_TRY(cudaMalloc(&h.ucount, sizeof(unsigned int)));
    const bool want_rows = uncor_host && uncor_capacity > 0 && harvest_mode != 0;
    if (want_rows && h.ubuf_rows < uncor_capacity) {
        cudaFree(h.ubuf); h.ubuf = nullptr; h.ubuf_rows = 0;
        CUDA_TRY(cudaMalloc(&h.ubuf, (size_t)uncor_capacity * d->base.NZ * sizeof(float)));
        h.ubuf_rows = uncor_capacity;
    }
    CUDA_TRY(cudaMemsetAsync(h.counters, 0, LDPC_NUM_COUNTERS * sizeof(unsigned long long), 0));
    CUDA_TRY(cudaMemsetAsync(h.ucount, 0, sizeof(unsigned int), 0));
    int rc = ldpc_mc_run(d, sigma, n_frames, seed, frame_offset, iters, early_term, harvest_mode,
                         (uint64_t *)h.counters, want_rows ? h.ubuf : nullptr, h.ucount, want_rows ? uncor_capacity : 0,
                         nullptr);
    if (rc != LDPC_OK) return rc;
    CUDA_TRY(cudaMemcpy(counters_host, h.counters, LDPC_NUM_COUNTERS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned int n = 0;
    CUDA_TRY(cudaMemcpy(&n, h.ucount, sizeof n, cudaMemcpyDeviceToHost));
    n = std::min(n, want_rows ? uncor_capacity : 0u);
    if (n_uncor_host) *n_uncor_host = n;
    if (n > 0) CUDA_TRY(cudaMemcpy(uncor_host, h.ubuf, (size_t)n * d->base.NZ * sizeof(float), cudaMemcpyDeviceToHost));
    return LDPC_OK;
}

extern "C" int ldpc_post_decode(const ldpc_decoder_t *post, const float *uncor_dev, int64_t n_words, int32_t iters,
                                int32_t early_term, uint64_t *counters_dev, uint32_t *hard_dev, int32_t *iters_dev,
                                uint8_t *flags_dev, void *stream) {
    if (!post || (!uncor_dev && n_words > 0) || n_words < 0) return fail(LDPC_E_INVALID, "post_decode: bad arguments");
    if (iters < 0 || iters > post->T) return fail(LDPC_E_INVALID, "post_decode: iters %d outside 0..%d", iters, post->T);
    if (n_words == 0) return LDPC_OK;
    DeviceGuard guard(post->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", post->device);
    KParams P = post->base;
    P.T_run = iters == 0 ? post->T : iters;
    P.early_term = early_term ? 1 : 0;
    P.llr = uncor_dev; P.n_frames = n_words;
    P.counters = (unsigned long long *)counters_dev;
    P.hard = hard_dev; P.iters = iters_dev; P.flags = flags_dev;
    return launch(post, P, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------- training step
extern "C" int ldpc_decoder_set_weights(ldpc_decoder_t *d, const float *w_cn, const float *w_ucn, const float *w_vn) {
    if (!d) return fail(LDPC_E_INVALID, "set_weights: null decoder");
    if ((d->wc_raw && !w_cn) || (d->wu_raw && !w_ucn) || (d->wv_raw && !w_vn))
        return fail(LDPC_E_INVALID, "set_weights: missing weight block");
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    std::lock_guard<std::mutex> lock(d->mu);
    const size_t T = (size_t)d->T;
    std::vector<float> raw, eff;
    if (d->wc_raw) raw.assign(w_cn, w_cn + T * d->wc_raw);
    if (d->wu_raw) raw.insert(raw.end(), w_ucn, w_ucn + T * d->wu_raw);
    if (d->wv_raw) raw.insert(raw.end(), w_vn, w_vn + T * d->wv_raw);
    if (d->ones_cn) eff.assign(T, 1.0f);            // same "effective" layout ldpc_decoder_create2 built
    eff.insert(eff.end(), raw.begin(), raw.end());
    if ((int)eff.size() != d->base.w_words) return fail(LDPC_E_INVALID, "internal: weight block size changed");
    CUDA_TRY(cudaDeviceSynchronize());               // no launch may still be reading the old block
    if (!eff.empty()) CUDA_TRY(cudaMemcpy(d->d_w, eff.data(), eff.size() * sizeof(float), cudaMemcpyHostToDevice));
    d->w_raw = raw;
    if (d->d_w_raw && !raw.empty())
        CUDA_TRY(cudaMemcpy(d->d_w_raw, raw.data(), raw.size() * sizeof(float), cudaMemcpyHostToDevice));
    return LDPC_OK;
}

extern "C" int ldpc_train_grad(const ldpc_decoder_t *dc, const float *llr_dev, int64_t B, int32_t iters, int32_t iter_lo,
                               int32_t loss_type, double etha, double *loss_host, float *g_cn_host, float *g_ucn_host,
                               float *g_vn_host, float *app_dev) {
    ldpc_decoder *d = const_cast<ldpc_decoder *>(dc);
    if (!d || !llr_dev || B <= 0 || !loss_host) return fail(LDPC_E_INVALID, "train_grad: bad arguments");
    const int T = iters == 0 ? d->T : iters;
    if (T < 1 || T > d->T || iter_lo < 0 || iter_lo >= T) return fail(LDPC_E_INVALID, "train_grad: iterations [%d, %d) outside 0..%d", iter_lo, T, d->T);
    if (loss_type < 0 || loss_type > 2) return fail(LDPC_E_INVALID, "train_grad: loss_type %d (0 BCE, 1 soft BER, 2 FER)", loss_type);
    if (d->g.info.max_dc > 64) return fail(LDPC_E_LIMIT, "train_grad: row degree %d > 64", d->g.info.max_dc);
    if (d->decoding_type == 0) return fail(LDPC_E_UNSUPPORTED, "train_grad: sum-product (decoding_type 0) has no training kernel");
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LDPC_E_CUDA, "cudaSetDevice(%d) failed", d->device);
    std::lock_guard<std::mutex> lock(d->mu);
    const ldpc_graph &g = d->g;
    const int E = g.E, M = g.M, N = g.N;
    if (!d->d_tab) {
        std::vector<int> tab;
        tab.insert(tab.end(), g.row.begin(), g.row.end());
        tab.insert(tab.end(), g.col.begin(), g.col.end());
        tab.insert(tab.end(), g.shift.begin(), g.shift.end());
        tab.insert(tab.end(), g.row_ptr.begin(), g.row_ptr.end());
        tab.insert(tab.end(), g.col_ptr.begin(), g.col_ptr.end());
        tab.insert(tab.end(), g.col_edge.begin(), g.col_edge.end());
        CUDA_TRY(cudaMalloc(&d->d_tab, tab.size() * sizeof(int)));
        CUDA_TRY(cudaMemcpy(d->d_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
        const size_t nw = std::max<size_t>(d->w_raw.size(), 1);
        CUDA_TRY(cudaMalloc(&d->d_w_raw, nw * sizeof(float)));
        if (!d->w_raw.empty())
            CUDA_TRY(cudaMemcpy(d->d_w_raw, d->w_raw.data(), d->w_raw.size() * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(&d->d_grad, nw * sizeof(float)));
        CUDA_TRY(cudaMalloc(&d->d_coef, LDPC_MAX_T * sizeof(float)));
        CUDA_TRY(cudaMalloc(&d->d_loss, sizeof(double)));
    }
    TrainParams P{};
    P.M = M; P.N = N; P.E = E; P.z = g.z; P.NZ = N * g.z; P.EZ = E * g.z; P.MZ = M * g.z;
    P.row = d->d_tab; P.col = P.row + E; P.shift = P.col + E; P.row_ptr = P.shift + E; P.col_ptr = P.row_ptr + M + 1;
    P.col_edge = P.col_ptr + N + 1;
    P.qms = d->decoding_type == 2; P.qmagic = d->base.qmagic; P.qmax = d->base.qmax; P.clip = d->clip;
    P.sharing0 = d->sharing[0]; P.sharing1 = d->sharing[1]; P.sharing2 = d->sharing[2];
    P.wc = d->wc_raw; P.wu = d->wu_raw; P.wv = d->wv_raw;
    P.w = d->d_w_raw; P.off_cn = 0; P.off_ucn = d->T * P.wc; P.off_vn = d->T * (P.wc + P.wu);
    P.T = T; P.t_lo = iter_lo; P.loss_type = loss_type; P.target_nz = d->base.target_n * g.z; P.B = (int)B;
    P.llr = llr_dev; P.app_out = app_dev;
    const size_t need = (size_t)B * (T + 1) * P.EZ;
    if (d->hist_cap < need) {
        cudaFree(d->d_hist); d->d_hist = nullptr; d->hist_cap = 0;
        CUDA_TRY(cudaMalloc(&d->d_hist, need * sizeof(float)));
        d->hist_cap = need;
    }
    P.hist = d->d_hist; P.loss = d->d_loss; P.grad = d->d_grad; P.coef = d->d_coef;
    if (nms_train_smem_bytes(P) > 227 * 1024) return fail(LDPC_E_LIMIT, "train_grad: graph needs %zu bytes of shared memory", nms_train_smem_bytes(P));
    // loss coefficients pow(etha, T-1-t) / sum (Main_Functions.py:342-354); pow(0, 0) = 1
    std::vector<float> coef(LDPC_MAX_T, 0.0f);
    double norm = 0.0;
    for (int t = iter_lo; t < T; ++t) norm += std::pow(etha, (double)(T - 1 - t));
    for (int t = iter_lo; t < T; ++t) coef[t] = (float)(std::pow(etha, (double)(T - 1 - t)) / norm);
    CUDA_TRY(cudaMemcpy(d->d_coef, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice));
    const size_t nw = std::max<size_t>(d->w_raw.size(), 1);
    CUDA_TRY(cudaMemset(d->d_grad, 0, nw * sizeof(float)));
    CUDA_TRY(cudaMemset(d->d_loss, 0, sizeof(double)));
    CUDA_TRY(nms_launch_train(P, nullptr));
    CUDA_TRY(cudaMemcpy(loss_host, d->d_loss, sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<float> gh(nw);
    CUDA_TRY(cudaMemcpy(gh.data(), d->d_grad, nw * sizeof(float), cudaMemcpyDeviceToHost));
    const size_t Tw = (size_t)d->T;
    if (g_cn_host && P.wc) std::memcpy(g_cn_host, gh.data(), Tw * P.wc * sizeof(float));
    if (g_ucn_host && P.wu) std::memcpy(g_ucn_host, gh.data() + Tw * P.wc, Tw * P.wu * sizeof(float));
    if (g_vn_host && P.wv) std::memcpy(g_vn_host, gh.data() + Tw * (P.wc + P.wu), Tw * P.wv * sizeof(float));
    return LDPC_OK;
}

// Run-time specialisation without a device: make sure the cubins a decoder for this graph would ask for are in the on-disk
// cache (NVRTC cross-compiles for sm_100a).  Returns the number of kernels now cached, or a negative error code.
extern "C" int ldpc_jit_prebuild(const int32_t *proto, int32_t M, int32_t N, int32_t z) {
    if (!proto || M <= 0 || N <= 0 || z <= 0) return fail(LDPC_E_INVALID, "jit_prebuild: bad arguments");
    if (!nms_jit_available()) return fail(LDPC_E_UNSUPPORTED, "jit_prebuild: libnvrtc not available");
    int E = 0, max_dc = 0;
    for (int i = 0; i < M; ++i) {
        int dc = 0;
        for (int j = 0; j < N; ++j) dc += proto[(size_t)i * N + j] != -1;
        if (dc == 0) return fail(LDPC_E_LIMIT, "jit_prebuild: row %d has no edge", i);
        E += dc; max_dc = std::max(max_dc, dc);
    }
    for (int j = 0; j < N; ++j) {
        int dv = 0;
        for (int i = 0; i < M; ++i) dv += proto[(size_t)i * N + j] != -1;
        if (dv == 0) return fail(LDPC_E_LIMIT, "jit_prebuild: column %d has no edge", j);
    }
    if (max_dc > 64 || M > LDPC_MAX_M || N > LDPC_MAX_N || E > LDPC_MAX_E) return fail(LDPC_E_LIMIT, "jit_prebuild: graph outside the specialised kernels' limits");
    int built = 0;
    char err[512] = {0};
    for (int kind = NMS_JIT_DECODE; kind <= NMS_JIT_MCP; ++kind) {
        if (kind == NMS_JIT_MCP && N * z >= 65536) continue;
        int Fp = 0, R = 0;
        nms_jit_pick_geometry(M, N, E, z, kind, max_dc, &Fp, &R);
        if (nms_jit_build(proto, M, N, z, Fp, R, kind, nullptr, err, (int)sizeof err) != 0) return fail(LDPC_E_UNSUPPORTED, "jit_prebuild: %s", err);
        ++built;
    }
    return built;
}

extern "C" int nms_alu_probe(int device, int kind, double *lane_ops_per_s);

extern "C" int ldpc_alu_peak_probe(int32_t device, int32_t kind, double *lane_ops_per_s) {
    if (!lane_ops_per_s || kind < 0 || kind > 6) return fail(LDPC_E_INVALID, "alu_peak_probe: kind 0..6, non-null result");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(LDPC_E_CUDA, "no CUDA device"); }
    if (device < 0 || device >= ndev) return fail(LDPC_E_INVALID, "device %d out of range", device);
    const int rc = nms_alu_probe(device, kind, lane_ops_per_s);
    if (rc != 0) return fail(LDPC_E_CUDA, "alu_peak_probe: %s", cudaGetErrorString((cudaError_t)rc));
    nms_note_launch(); nms_note_launch();
    return LDPC_OK;
}

extern "C" uint64_t ldpc_launch_count(void) { return nms_launch_count(); }
