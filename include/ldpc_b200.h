/*
 * ldpc_b200.h -- C-ABI of the B200-native neural min-sum (NMS) LDPC decode / Monte-Carlo path.
 *
 * This is the drop-in boundary for the ONE hot path of ghy1228/LDPC_Error_Floor:
 *   sess.run(net_dict["ya_output_all"], feed_dict={xa, ya, ...})     Print_Functions.py:148-151
 *   Print_Functions.compute_results(...)                             Print_Functions.py:130-165
 *   Print_Functions.create_mix_epoch(...)                            Print_Functions.py:29-72
 *   Main_Functions.init_parameter / init_connecting_matrix           Main_Functions.py:8-150
 * The reference has no FFI of its own (pure Python on TensorFlow); each entry point below
 * cites the reference interface it replaces.  INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 or a negative
 * LDPC_E_* code, never throws, never exits; `*_dev` pointers are caller-owned device
 * allocations on the handle's device; `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream); launches are asynchronous w.r.t. the host unless noted;
 * handles are immutable after creation and may be shared across streams.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * LDPC_E_CUDA.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_OK 0
#define LDPC_E_INVALID (-1)   /* bad argument (the reference prints and sys.exit()s, Main_Functions.py:503-521) */
#define LDPC_E_UNSUPPORTED (-2)
#define LDPC_E_CUDA (-3)      /* CUDA runtime error; text via ldpc_last_error() */
#define LDPC_E_ALLOC (-4)
#define LDPC_E_LIMIT (-5)     /* graph exceeds compiled table limits */

typedef struct ldpc_graph ldpc_graph_t;       /* compiled base graph (host tables) */
typedef struct ldpc_decoder ldpc_decoder_t;   /* graph + weights + arithmetic mode on one device */

/* Result of the base-graph compiler.  Replaces the scalars returned by
 * Main_Functions.init_parameter (Main_Functions.py:8-38). */
typedef struct {
    int32_t M, N, z, E;          /* proto rows / cols, lifting size, proto edges */
    int32_t max_dc, max_dv;      /* largest proto row / column degree */
    int32_t n_ref, k_ref;        /* n, k exactly as the reference computes them, incl. the "+1"
                                    when no puncturing/shortening is configured (:24-28) */
    int32_t n_true, k_true;      /* transmitted / information bits without that quirk */
    double rate_ref, rate_true;  /* k/n both ways; the reference's sigma uses rate_ref (:29,:36) */
} ldpc_graph_info_t;

/* Per-SNR-point Monte-Carlo counters (uint64 each).  Replaces the float32 Results[4,nSNR]
 * accumulators of Print_Functions.compute_results (:133, :158-161). */
enum {
    LDPC_CNT_FRAMES = 0,       /* frames decoded */
    LDPC_CNT_FRAME_ERR_LAST,   /* FER_last numerator: last-iteration hard decision != codeword (:115-116) */
    LDPC_CNT_FRAME_ERR_ANY,    /* FER numerator: never correct at ANY executed iteration (:105-111) */
    LDPC_CNT_BIT_ERR_LAST,     /* BER_last numerator (:112-113) */
    LDPC_CNT_ITERS,            /* sum of iterations executed (== frames*T without early termination) */
    LDPC_CNT_SYND_FAIL,        /* final hard decision violates a parity check (detected failure) */
    LDPC_CNT_UNDETECTED,       /* final hard decision is a codeword but not the transmitted one */
    LDPC_CNT_HARVESTED,        /* words that met the harvest criterion (may exceed the buffer capacity) */
    LDPC_NUM_COUNTERS
};

/* flags[] bits written by ldpc_decode */
#define LDPC_FLAG_SYND_OK 1u      /* output hard decision satisfies every check */
#define LDPC_FLAG_UNCOR_ANY 2u    /* never equal to the all-zero codeword at any executed iteration (D9 uncor_flag) */
#define LDPC_FLAG_UNCOR_LAST 4u   /* output hard decision != all-zero codeword */
#define LDPC_FLAG_SYND_OK_EVER 8u /* some iteration's hard decision satisfied every check */

/* harvest criteria for ldpc_mc_run */
#define LDPC_HARVEST_NONE 0
#define LDPC_HARVEST_UNCOR_ANY 1   /* the reference's criterion (Print_Functions.py:155-156) */
#define LDPC_HARVEST_UNCOR_LAST 2
#define LDPC_HARVEST_SYND_FAIL 3

const char *ldpc_last_error(void);   /* thread-local text of the last failure */
int ldpc_version(void);

/* ---- base-graph compiler -------------------------------------------------------------
 * proto: M x N int32, row-major, -1 = no edge else circulant shift (used mod z); z = 1 for
 * non-QC codes (MacKay / BCH / Polar: entries 0 / -1).  punct/short ranges are 1-based
 * inclusive bit indices, 0,0 = none (main_Base.py:31-34).
 * Replaces init_parameter + init_connecting_matrix (Main_Functions.py:8-150): instead of
 * dense (E*z)^2 permutation matrices it emits circulant shift tables in E(C) (row-major)
 * order and a column-sorted CSR edge list. */
int ldpc_graph_create(const int32_t *proto, int32_t M, int32_t N, int32_t z, int32_t punct_start,
                      int32_t punct_end, int32_t short_start, int32_t short_end, ldpc_graph_t **out);
int ldpc_graph_destroy(ldpc_graph_t *g);
int ldpc_graph_info(const ldpc_graph_t *g, ldpc_graph_info_t *info);
/* edge tables, each E int32 in E(C) order (any may be NULL): proto row, proto column, shift */
int ldpc_graph_edges(const ldpc_graph_t *g, int32_t *row, int32_t *col, int32_t *shift);
/* sigma[i] = sqrt(1 / (2 * R * 10^(snr_db[i]/10))), R = rate_ref (use_ref_rate != 0, the
 * reference's value, Main_Functions.py:35-36) or rate_true. */
int ldpc_graph_sigma(const ldpc_graph_t *g, const double *snr_db, int32_t n, int32_t use_ref_rate,
                     double *sigma);

/* ---- decoder handle ------------------------------------------------------------------
 * sharing[3] = {CN, UCN, VN} codes as in main_Base.py:24-25 (0 none, 1 per edge (E(C) order),
 * 2 per proto node, 3 one scalar per iteration); rules of check_params apply
 * (Main_Functions.py:515-521).  w_cn / w_ucn / w_vn: HOST float32 [T, width], width =
 * 1 / M (CN,UCN) or N (VN) / E by code (Main_Functions.py:397-405); NULL when the code is 0.
 * decoding_type 0 = sum-product (tanh / atanh check update, Main_Functions.py:238-245; decode and Monte-Carlo only, no
 * training kernel), 1 = min-sum (float32 messages, clip +-clip_llr), 2 = quantised min-sum
 * (q_bit in {5,6,-5,4,3}, Main_Functions.py:483-492).  Replaces weight_init
 * (Main_Functions.py:387-439) + the graph constants of build_neural_network. */
int ldpc_decoder_create(const ldpc_graph_t *g, const int32_t sharing[3], int32_t T,
                        const float *w_cn, const float *w_ucn, const float *w_vn,
                        int32_t decoding_type, int32_t q_bit, float clip_llr, int32_t device,
                        ldpc_decoder_t **out);
/* Same with the `systematic` switch of main_Base.py:29, 83-86: target_node > 0 restricts the error metrics
 * (LDPC_FLAG_UNCOR_*, biterr, the FER/BER counters) to the first target_node proto columns, which is what
 * ya_output_all / calc_ber_fer see when systematic = 1 (Main_Functions.py:329-335, Print_Functions.py:101-107);
 * 0 = all N columns.  Syndrome flags and hard-decision outputs always cover the whole word.
 * Sharing code 4 (temporal sharing of per-edge CN weights, Main_Functions.py:299-304) is accepted for
 * sharing[0] with the T expanded rows print_weight writes (Print_Functions.py:87-94). */
int ldpc_decoder_create2(const ldpc_graph_t *g, const int32_t sharing[3], int32_t T,
                         const float *w_cn, const float *w_ucn, const float *w_vn,
                         int32_t decoding_type, int32_t q_bit, float clip_llr, int32_t device,
                         int32_t target_node, ldpc_decoder_t **out);
int ldpc_decoder_destroy(ldpc_decoder_t *d);
/* 1 if the packed-fp16x2 kernel serves this decoder, 0 if the float32 kernel does */
int ldpc_decoder_uses_packed_kernel(const ldpc_decoder_t *d);
/* name of the __global__ function that serves this decoder (graph-specialised or generic bucket) */
const char *ldpc_decoder_kernel_name(const ldpc_decoder_t *d);
/* frames decoded per CTA and CTAs per SM the launcher will use (for sizing batches) */
int ldpc_decoder_geometry(const ldpc_decoder_t *d, int32_t *frames_per_cta, int32_t *ctas_per_sm,
                          int32_t *threads_per_cta, int32_t *smem_bytes);
/* Kernel and launch geometry of a call with (early_term != 0) or without early termination: a graph may carry a second
 * geometry for fixed-iteration launches (more frames per CTA).  kernel_name: caller's buffer of name_cap bytes (nullable). */
int ldpc_decoder_launch_info(const ldpc_decoder_t *d, int32_t early_term, int32_t *frames_per_cta, int32_t *ctas_per_sm,
                             int32_t *threads_per_cta, int32_t *smem_bytes, char *kernel_name, int32_t name_cap);

/* ---- run-time graph specialisation ------------------------------------------------------
 * Main_Functions.init_connecting_matrix (:46-150) accepts any proto matrix at run time.  Base graphs the build knows
 * have unrolled kernels compiled in; for any other graph ldpc_decoder_create generates the same kernel source for it,
 * compiles it with NVRTC (sm_100a) and keeps the cubin in an on-disk cache (<library dir>/jit_cache, or
 * $LDPC_B200_JIT_CACHE), so user graphs get the fast path too -- bit-identical to the table-driven generic kernels,
 * which stay the fallback (LDPC_B200_NO_JIT=1, or no libnvrtc).  ldpc_jit_prebuild fills the cache ahead of time and
 * needs no GPU; returns the number of kernels cached (>= 0) or an error code. */
int ldpc_jit_prebuild(const int32_t *proto, int32_t M, int32_t N, int32_t z);

/* ---- decode --------------------------------------------------------------------------
 * Replaces sess.run(ya_output_all / ya_output{t}) (Print_Functions.py:148-151, main_Base.py:160).
 *   llr_dev      f32 [B, N*z]  channel LLRs log(p1/p0), bit index j*z+c (== xa[B,N,z] flattened)
 *   iters        iterations to run, 1..T (0 = T)
 *   early_term   0: run all iterations (reference behaviour); 1: stop a frame at the first
 *                iteration whose hard decision satisfies every check
 *   app_dev      f32, NULL | [B, N*z] (app_all_iters == 0: APP of the output iteration) |
 *                [iters, B, N*z] (app_all_iters != 0: ya_output_all, Main_Functions.py:380-383;
 *                rows of iterations after a frame's early stop are left untouched)
 *   hard_dev     u32 [B, ceil(N*z/32)]  bit k of a frame = (APP[k] >= 0) at word k/32, bit k%32
 *   iters_dev    i32 [B]  iterations until the first zero syndrome (also reported when
 *                early_term == 0), or `iters` if none
 *   flags_dev    u8  [B]  LDPC_FLAG_*
 *   biterr_dev   i32 [B]  ones in the output hard decision (bit errors vs the all-zero word; other codewords:
 *                ldpc_decode_cw)
 * Any output pointer may be NULL. */
int ldpc_decode(const ldpc_decoder_t *d, const float *llr_dev, int64_t B, int32_t iters,
                int32_t early_term, float *app_dev, int32_t app_all_iters, uint32_t *hard_dev,
                int32_t *iters_dev, uint8_t *flags_dev, int32_t *biterr_dev, void *stream);

/* Same with HOST buffers: pinned or pageable host memory in and out, chunked over four
 * streams so copies overlap compute; synchronous.  This is the end-to-end call: what a reference
 * maintainer binds in place of sess.run(..., feed_dict={xa_input: ...}) (Print_Functions.py:148-150).
 * float32 words are the bound of this path (PCIe), so for a quantised decoder the library's host
 * threads pack chunks to int8 (one byte per value) from the front of the call while the DMA engine
 * takes chunks from its back as they are -- the two lanes meet wherever their speeds put them:
 * same decoder input bit for bit -- a quantised decoder sees a channel value only through Q(x)
 * (Main_Functions.py:321-322) and, with VN weights, Q(x * w) (:168-177), so words are packed always
 * without VN weights and, with them, whenever every value of the chunk is on the quantiser grid
 * (always, in the reference's flows).  With fewer than three host threads pinned float32 input is not
 * packed at all (the DMA engine alone is faster).  LDPC_B200_NO_HOST_PACK=1 keeps float32 for every chunk,
 * LDPC_B200_HOST_THREADS sets the pool size (default: hardware threads / LOCAL_WORLD_SIZE, <= 16). */
int ldpc_decode_host(const ldpc_decoder_t *d, const float *llr_host, int64_t B, int32_t iters,
                     int32_t early_term, float *app_host, int32_t app_all_iters,
                     uint32_t *hard_host, int32_t *iters_host, uint8_t *flags_host,
                     int32_t *biterr_host);

/* What the last ldpc_decode_host / ldpc_decode_q8_host call on this handle did. */
typedef struct {
    int32_t threads;            /* host threads that packed / staged */
    int32_t chunks_total;       /* chunks of the call */
    int32_t chunks_q8;          /* chunks that crossed PCIe as int8 */
    int32_t chunks_f32;         /* chunks that crossed as float32 by choice (DMA engine and cores share the load) */
    int32_t chunks_unencodable; /* chunks that crossed as float32 because a value has no int8 form */
    double float_share;         /* share of the chunks that crossed as float32 */
    double s_pack;      /* feeder thread: packing / staging (seconds) */
    double s_wait;      /* calling thread: waiting for the device */
    double s_wait_feed; /* calling thread: waiting for the feeder */
    double s_copy_out;  /* copying results into the caller's arrays (feeder thread + pool; calling thread for short calls) */
    double s_total;     /* the whole call */
    int64_t h2d_bytes, d2h_bytes;               /* bytes actually copied */
} ldpc_host_stats_t;
int ldpc_decode_host_stats(const ldpc_decoder_t *d, ldpc_host_stats_t *out);

/* The packing pass on its own (pack once, decode many times with ldpc_decode_q8_host; no GPU work):
 * q8_host[i] = k with k * ldpc_decoder_q8_step(d) the decoder's view of llr_host[i]; *n_unencodable = values
 * that have no such k (always 0 for decoders without VN weights; off-grid or out-of-range values otherwise).
 * LDPC_E_UNSUPPORTED for float decoders and q_bit 6 (saturates off its grid). */
int ldpc_pack_q8_host(const ldpc_decoder_t *d, const float *llr_host, int64_t B, int8_t *q8_host,
                      int64_t *n_unencodable);
/* The same pass without a handle (step = 1/qk and qmax of Main_Functions.py:483-492; qmax / step <= 127).
 * lossless = 0: q8 = Q(x) / step for every x (NaN -> -qmax, as the kernels' clamp does); lossless = 1: only
 * values that are on the grid and inside +-qmax count as encodable. */
int ldpc_pack_q8_values(const float *x, int64_t n, float step, float qmax, int32_t lossless, int8_t *q8,
                        int64_t *n_unencodable);

/* ---- non-zero codewords ("next" row N3 of SURVEY.md 8f) ----------------------------------------
 * The reference can train / evaluate on random codewords Y = infoWord . code_GM mod 2 (Print_Functions.py:40-46 with
 * is_zeros_word = False; main_Base.py hard-codes the all-zero word) and counts errors against Y (:100-118).
 * codeword_bits_dev: packed bits, bit k of a frame at word k / 32, bit k % 32; cw_stride_words = words between the
 * codewords of consecutive frames (ceil(N*z/32) for one word per frame, 0 = every frame carries the same codeword).
 *
 * ldpc_llr_generate_cw: the samples of ldpc_llr_generate (same Philox stream) around +-1 per bit:
 *   llr = 2 (sigma n + (2 y - 1)) / sigma^2, then quantise / puncture / shorten as there.
 * ldpc_decode_cw: ldpc_decode with the error metrics taken against the codeword -- flags UNCOR_ANY / UNCOR_LAST,
 *   biterr_dev (Hamming distance of the output decision), biterr_signed_dev (sum over bits of decision - y: what
 *   calc_ber_fer sums, :112-113 -- its BER lets 0->1 and 1->0 errors cancel) and counters_dev as in ldpc_mc_run
 *   (accumulated).  Implemented as the decode kernel writing the per-iteration APPs of a chunk of frames plus a
 *   metrics kernel over them (the reference's own order of work, Print_Functions.py:148-154); any output may be NULL. */
int ldpc_llr_generate_cw(const ldpc_decoder_t *d, double sigma, int64_t n_frames, uint64_t seed, uint64_t frame_offset,
                         const uint32_t *codeword_bits_dev, int64_t cw_stride_words, float *llr_dev, void *stream);
int ldpc_decode_cw(const ldpc_decoder_t *d, const float *llr_dev, const uint32_t *codeword_bits_dev,
                   int64_t cw_stride_words, int64_t B, int32_t iters, int32_t early_term, uint32_t *hard_dev,
                   int32_t *iters_dev, uint8_t *flags_dev, int32_t *biterr_dev, int32_t *biterr_signed_dev,
                   uint64_t *counters_dev, void *stream);

/* ---- compact (int8) words ---------------------------------------------------------------
 * On the quantised path every decoder input of Print_Functions.create_mix_epoch (:49-50) and every
 * row of an Inputs/[Uncor]_* file (:120-126, '%.1f' of values on the 0.5 grid) is a small multiple of
 * the quantiser step, so a word fits in N*z int8: llr = q * step.  step = 0 selects the decoder's
 * own quantiser step (ldpc_decoder_q8_step: 0.5 for q_bit 5, 1 for 6/-5/4, 2 for 3; float decoders
 * need an explicit step).  Same outputs as ldpc_decode (no APP), plus optional accumulated
 * counters_dev (u64[LDPC_NUM_COUNTERS]).  A quarter of the bytes of the float32 form over PCIe / HBM;
 * formats.py keeps the matching binary sidecar of the text files ("next" row N2 of SURVEY.md 8f). */
float ldpc_decoder_q8_step(const ldpc_decoder_t *d);
int ldpc_decode_q8(const ldpc_decoder_t *d, const int8_t *llr_q8_dev, float step, int64_t B, int32_t iters,
                   int32_t early_term, uint32_t *hard_dev, int32_t *iters_dev, uint8_t *flags_dev,
                   int32_t *biterr_dev, uint64_t *counters_dev, void *stream);
int ldpc_decode_q8_host(const ldpc_decoder_t *d, const int8_t *llr_q8_host, float step, int64_t B,
                        int32_t iters, int32_t early_term, uint32_t *hard_host, int32_t *iters_host,
                        uint8_t *flags_host, int32_t *biterr_host);

/* ---- channel-sample generator ----------------------------------------------------------
 * Replaces Print_Functions.create_mix_epoch (:29-72) for the all-zero codeword:
 * llr = 2*(sigma*n - 1)/sigma^2, n ~ N(0,1) from Philox4x32-10 (key = seed, counter =
 * (global frame index, bit-quad index)) + Box-Muller; quantised if the decoder is QMS;
 * punctured bits -> 0, shortened bits -> -clip_llr.  Frame f of this call is global frame
 * frame_offset + f, so shards on different ranks draw disjoint, world-size-independent
 * streams.  ldpc_mc_run draws exactly the same samples. */
int ldpc_llr_generate(const ldpc_decoder_t *d, double sigma, int64_t n_frames, uint64_t seed,
                      uint64_t frame_offset, float *llr_dev, void *stream);

/* The generator's N(0,1) stream on its own (same Philox counters: frame f, quad q -> normals 4q..4q+3 of frame
 * frame_offset + f): normals_dev f32 [n_frames * quads_per_frame * 4] (NULL = no copy-out) and/or tail_counts_dev
 * u64[6], ACCUMULATED: samples with |n| > 3, 4, 5, 6, 7 sigma and the total.  Test hook for the distribution the
 * error-floor estimates rest on (the reference draws float64 normals, Print_Functions.py:45). */
int ldpc_normal_probe(int32_t device, uint64_t seed, uint64_t frame_offset, int64_t n_frames, int32_t quads_per_frame,
                      float *normals_dev, uint64_t *tail_counts_dev, void *stream);

/* ---- fused Monte-Carlo ---------------------------------------------------------------------
 * Replaces Print_Functions.compute_results (:130-165) for one SNR point: generate, decode,
 * count, harvest, all on the device.  counters_dev: u64[LDPC_NUM_COUNTERS], ACCUMULATED
 * into (zero them first).  uncor_buf_dev: f32 [uncor_capacity, N*z] decoder-input LLRs of
 * harvested words (NULL/0 = none), uncor_count_dev: u32 running count (accumulated; rows
 * beyond the capacity are counted but dropped). */
int ldpc_mc_run(const ldpc_decoder_t *d, double sigma, int64_t n_frames, uint64_t seed,
                uint64_t frame_offset, int32_t iters, int32_t early_term, int32_t harvest_mode,
                uint64_t *counters_dev, float *uncor_buf_dev, uint32_t *uncor_count_dev,
                uint32_t uncor_capacity, void *stream);

/* Kernel and launch geometry ldpc_mc_run(early_term = 1) uses.  *persistent = 1: the graph has a persistent-slot
 * Monte-Carlo kernel (csrc/nms_mcp.cuh): a CTA owns frames_per_cta frame slots, every slot runs its own frame at its
 * own iteration and is refilled the moment that frame stops -- the device-side replacement of the batch loop
 * `for batch_idx ...` of Print_Functions.compute_results (:136-161).  Same counters and harvested words as the batch
 * kernels, bit for bit (the Philox counter is the global frame index either way). */
int ldpc_decoder_mc_info(const ldpc_decoder_t *d, int32_t *persistent, int32_t *frames_per_cta, int32_t *ctas_per_sm,
                         int32_t *threads_per_cta, int32_t *smem_bytes, char *kernel_name, int32_t name_cap);

/* Host-buffer twin of ldpc_mc_run: counters_host u64[LDPC_NUM_COUNTERS] (overwritten),
 * uncor_host f32 [uncor_capacity, N*z], *n_uncor_host = rows written.  Synchronous. */
/* Two-stage form of ldpc_mc_run with early termination: stage 1 decodes every frame for stage1_iters (< iters) iterations;
 * frames that have not reached a zero syndrome by then are not counted but listed by global frame index in defer_list_dev
 * (caller-owned, uint64[defer_capacity]; defer_count_dev: uint32[1]; n_frames entries always suffice, and when more frames are
 * deferred than the list holds the call fails with LDPC_E_LIMIT instead of writing past it) and decoded in full by stage 2, which regenerates them from
 * the same Philox counters.  Counters and harvested words are those of ldpc_mc_run(early_term = 1), bit for bit; the
 * stragglers no longer keep the CTAs of converged frames busy (a code whose degree-1 parity bits often stay wrong -- 5G NR --
 * otherwise pays all iterations for most CTAs).  Synchronises `stream` once between the stages.  Not in the reference. */
int ldpc_mc_run_staged(const ldpc_decoder_t *d, double sigma, int64_t n_frames, uint64_t seed, uint64_t frame_offset,
                       int32_t iters, int32_t stage1_iters, int32_t harvest_mode, uint64_t *counters_dev,
                       float *uncor_buf_dev, uint32_t *uncor_count_dev, uint32_t uncor_capacity,
                       uint64_t *defer_list_dev, uint32_t *defer_count_dev, uint32_t defer_capacity, void *stream);
int ldpc_mc_run_host(const ldpc_decoder_t *d, double sigma, int64_t n_frames, uint64_t seed,
                     uint64_t frame_offset, int32_t iters, int32_t early_term,
                     int32_t harvest_mode, uint64_t *counters_host, float *uncor_host,
                     uint32_t uncor_capacity, uint32_t *n_uncor_host);

/* ---- post decoder ----------------------------------------------------------------------
 * Replaces main_Post.py's use of the same graph on Inputs/[Uncor]_* words
 * (main_Post.py:25-38; Print_Functions.read_uncor_llr :6-10): `post` is a decoder built
 * from the boosted weight set (base rows followed by post rows); the words are re-decoded
 * from their channel LLRs, which is the reference semantic.  uncor_dev holds
 * n_words rows of N*z decoder-input LLRs, e.g. the buffer ldpc_mc_run filled.  counters_dev
 * as in ldpc_mc_run (accumulated); other outputs as in ldpc_decode. */
int ldpc_post_decode(const ldpc_decoder_t *post, const float *uncor_dev, int64_t n_words,
                     int32_t iters, int32_t early_term, uint64_t *counters_dev, uint32_t *hard_dev,
                     int32_t *iters_dev, uint8_t *flags_dev, void *stream);

/* ---- training step ("next" row N1 of SURVEY.md 8f) ------------------------------------------
 * ldpc_decoder_set_weights replaces the decoder's weights (same shapes as at creation): the one mutable part
 * of a handle; it synchronises the device first, and the caller must not launch from other threads meanwhile.
 * It stands for the variable update of tf.train.AdamOptimizer.minimize (Main_Functions.py:377-378); the
 * optimiser arithmetic itself lives on the host (trainer.py) -- there are at most a few hundred weights.
 *
 * ldpc_train_grad runs the reference's training batch (main_Base.py:160-162) for the all-zero codeword:
 * `iters` iterations forward, loss over iterations [iter_lo, iters) weighted by pow(etha, iters-1-t) and
 * normalised (Main_Functions.py:339-357; loss_type 0 = sigmoid cross entropy, 1 = soft BER, 2 = FER with the
 * straight-through sign), and its gradient with respect to the weights of iterations [iter_lo, iters)
 * (var_list, :360-375; iter_lo = max(training_iter_start - fixed_init, fixed_iter)).  Gradient rules are
 * TensorFlow's for the reference graph: straight-through quantisers (:475-494), tie-split reduce_min, no
 * gradient through signs / comparisons.  Outputs on the HOST: *loss_host; g_*_host f32 [T, width] (T = the
 * decoder's iteration count; rows outside [iter_lo, iters) are zero; NULL for absent blocks); optional
 * app_dev f32 [iters, B, N*z] (= ya_output{t}).  Synchronous; the error metrics' target_node applies. */
int ldpc_decoder_set_weights(ldpc_decoder_t *d, const float *w_cn, const float *w_ucn, const float *w_vn);
int ldpc_train_grad(const ldpc_decoder_t *d, const float *llr_dev, int64_t B, int32_t iters, int32_t iter_lo,
                    int32_t loss_type, double etha, double *loss_host, float *g_cn_host, float *g_ucn_host,
                    float *g_vn_host, float *app_dev);

/* ---- measurement aid ----------------------------------------------------------------------------
 * Instruction-issue micro-benchmark of the SM pipes the decode kernels are bound by (the ALU roofline is
 * to be measured on the box).  kind: 0 FFMA, 6 FADD (FP32 / FMA pipe), 1 FMNMX, 2 LOP3, 3 IADD (integer / logic /
 * min-max pipe), 4 HFMA2, 5 HMNMX2 (packed fp16x2).  *lane_ops_per_s = instructions x 32 lanes per second over the whole
 * GPU (a packed instruction counts once).  Synchronous, a few milliseconds.  No counterpart in the reference. */
int ldpc_alu_peak_probe(int32_t device, int32_t kind, double *lane_ops_per_s);

/* kernels launched by this library since load (for bench.py's gpu_launches) */
uint64_t ldpc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
