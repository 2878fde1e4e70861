// nms_gen.cu -- stand-alone BPSK/AWGN LLR generator kernel (same Philox stream as the fused
// Monte-Carlo path in nms_device.cuh) and the library-wide launch counter.
#include "nms_device.cuh"

#include <algorithm>
#include <atomic>

namespace nms {

__global__ void llr_generate_kernel(const __grid_constant__ KParams P, float *out, long long n_frames) {
    const int nquads = (P.NZ + 3) >> 2;
    const long long total = n_frames * nquads;
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < total;
         it += (long long)gridDim.x * blockDim.x) {
        const long long f = it / nquads;
        const int quad = (int)(it - f * nquads);
        float v[4];
        gen_llr4(P, P.frame_offset + (unsigned long long)f, quad, v);
        for (int k4 = 0; k4 < 4; ++k4) {
            const int k = 4 * quad + k4;
            if (k < P.NZ) out[f * P.NZ + k] = v[k4];
        }
    }
}

// Non-zero codewords (Print_Functions.py:40-46 with is_zeros_word = False): bit 1 -> +1, bit 0 -> -1, llr = 2 (n sigma + s) / sigma^2.
// cw: packed codeword bits (bit k at word k / 32, bit k % 32), one shared word (cw_stride == 0) or one per frame.
__global__ void llr_generate_cw_kernel(const __grid_constant__ KParams P, const uint32_t *cw, long long cw_stride, float *out,
                                       long long n_frames) {
    const int nquads = (P.NZ + 3) >> 2;
    const long long total = n_frames * nquads;
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
        const long long f = it / nquads;
        const int quad = (int)(it - f * nquads);
        float n[4];
        gen_normal4(P, P.frame_offset + (unsigned long long)f, quad, n);
        const uint32_t *w = cw + f * cw_stride;
        for (int k4 = 0; k4 < 4; ++k4) {
            const int k0 = 4 * quad + k4, k = k0 + 1;
            if (k0 >= P.NZ) break;
            const bool one = (__ldg(w + (k0 >> 5)) >> (k0 & 31)) & 1u;
            float llr = fmaf(n[k4], P.two_over_s, one ? P.two_over_s2 : -P.two_over_s2);
            if (P.qms) llr = qf(P, llr);
            if (P.punct_s > 0 && k >= P.punct_s && k <= P.punct_e) llr = P.sp ? 0.001f : 0.0f;
            if (P.short_s > 0 && k >= P.short_s && k <= P.short_e) llr = -P.clip;
            out[f * P.NZ + k0] = llr;
        }
    }
}

// calc_ber_fer (Print_Functions.py:100-118) against a codeword Y: one warp per frame walks the per-iteration APPs
// app[t][b][0 .. target_bits) of the iterations the frame executed and compares (APP >= 0) with Y.
//   flags: LDPC_FLAG_UNCOR_ANY / UNCOR_LAST rewritten (the syndrome bits stay), biterr: Hamming distance of the output decision,
//   biterr_signed: sum over bits of (decision - Y), the quantity the reference sums (its BER lets 0->1 and 1->0 errors cancel),
//   counters: the Monte-Carlo counters of ldpc_mc_run, accumulated.
__global__ void cw_metrics_kernel(const float *app, long long B, int T, int NZ, int target_bits, const uint32_t *cw, long long cw_stride,
                                  const int *iters, int early_term, uint8_t *flags, int *biterr, int *biterr_signed,
                                  unsigned long long *counters) {
    const int lane = threadIdx.x & 31;
    const long long b = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const uint8_t fl = flags[b];
    // iterations executed: with early termination a frame stops at its first zero syndrome (iters[b], flag SYND_OK set then)
    const int exec = (early_term && (fl & 1u)) ? iters[b] : T;
    const uint32_t *y = cw + b * cw_stride;
    bool ever = false;
    int dist = 0, sgn = 0;
    for (int t = 0; t < exec; ++t) {
        const float *a = app + ((long long)t * B + b) * NZ;
        int d = 0, s = 0;
        for (int k = lane; k < target_bits; k += 32) {
            const int h = a[k] >= 0.0f ? 1 : 0, yy = (int)((__ldg(y + (k >> 5)) >> (k & 31)) & 1u);
            d += h != yy;
            s += h - yy;
        }
        d = __reduce_add_sync(0xffffffffu, d);
        s = __reduce_add_sync(0xffffffffu, s);
        if (d == 0) ever = true;
        dist = d; sgn = s;
    }
    if (lane == 0) {
        const bool last = dist != 0, synd_ok = (fl & 1u) != 0;
        flags[b] = (uint8_t)((fl & ~6u) | (ever ? 0u : 2u) | (last ? 4u : 0u));
        if (biterr) biterr[b] = dist;
        if (biterr_signed) biterr_signed[b] = sgn;
        if (counters) {
            atomicAdd(counters + 0, 1ull);
            if (last) atomicAdd(counters + 1, 1ull);
            if (!ever) atomicAdd(counters + 2, 1ull);
            if (dist) atomicAdd(counters + 3, (unsigned long long)dist);
            atomicAdd(counters + 4, (unsigned long long)exec);
            if (!synd_ok) atomicAdd(counters + 5, 1ull);
            if (synd_ok && last) atomicAdd(counters + 6, 1ull);
        }
    }
}

// The generator's normals, bare: optional copy-out and tail counts (|n| > 3, 4, 5, 6, 7 sigma, total) -- what the tests
// hold against erfc and a Kolmogorov-Smirnov bound (tests/test_gpu_mc.py)
__global__ void normal_probe_kernel(unsigned long long seed, unsigned long long frame_offset, long long n_frames, int nquads,
                                    float *out, unsigned long long *counts) {
    const long long total = n_frames * nquads;
    unsigned c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
        const long long f = it / nquads;
        const int quad = (int)(it - f * nquads);
        float n[4];
        gen_normal4(seed, frame_offset + (unsigned long long)f, quad, n);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = fabsf(n[k]);
            c3 += a > 3.0f; c4 += a > 4.0f; c5 += a > 5.0f; c6 += a > 6.0f; c7 += a > 7.0f;
            if (out != nullptr) out[it * 4 + k] = n[k];
        }
    }
    if (counts != nullptr) {
        const unsigned v[5] = {c3, c4, c5, c6, c7};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const unsigned r = __reduce_add_sync(0xffffffffu, v[k]);
            if ((threadIdx.x & 31) == 0 && r) atomicAdd(counts + k, (unsigned long long)r);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(counts + 5, (unsigned long long)total * 4ull);
    }
}

}   // namespace nms

cudaError_t nms_launch_generate_cw(const KParams &P, const uint32_t *cw, long long cw_stride, float *out, long long n_frames,
                                   cudaStream_t st) {
    const long long items = n_frames * ((P.NZ + 3) / 4);
    const int threads = 256;
    const int grid = (int)std::min<long long>((items + threads - 1) / threads, 148LL * 16);
    nms::llr_generate_cw_kernel<<<std::max(grid, 1), threads, 0, st>>>(P, cw, cw_stride, out, n_frames);
    nms_note_launch();
    return cudaGetLastError();
}

cudaError_t nms_launch_cw_metrics(const float *app, long long B, int T, int NZ, int target_bits, const uint32_t *cw, long long cw_stride,
                                  const int *iters, int early_term, uint8_t *flags, int *biterr, int *biterr_signed,
                                  unsigned long long *counters, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    const int threads = 256;
    const long long grid = (B * 32 + threads - 1) / threads;
    nms::cw_metrics_kernel<<<(unsigned)grid, threads, 0, st>>>(app, B, T, NZ, target_bits, cw, cw_stride, iters, early_term, flags,
                                                              biterr, biterr_signed, counters);
    nms_note_launch();
    return cudaGetLastError();
}

cudaError_t nms_launch_normal_probe(unsigned long long seed, unsigned long long frame_offset, long long n_frames, int nquads,
                                    float *out, unsigned long long *counts, cudaStream_t st) {
    const long long items = n_frames * nquads;
    if (items <= 0) return cudaSuccess;
    const int threads = 256;
    const int grid = (int)std::min<long long>((items + threads - 1) / threads, 148LL * 8);
    nms::normal_probe_kernel<<<grid, threads, 0, st>>>(seed, frame_offset, n_frames, nquads, out, counts);
    nms_note_launch();
    return cudaGetLastError();
}

static std::atomic<unsigned long long> g_launches{0};
void nms_note_launch() { ++g_launches; }
unsigned long long nms_launch_count() { return g_launches.load(); }

cudaError_t nms_launch_generate(const KParams &P, float *out, long long n_frames, cudaStream_t st) {
    const long long items = n_frames * ((P.NZ + 3) / 4);
    const int threads = 256;
    const int grid = (int)std::min<long long>((items + threads - 1) / threads, 148LL * 16);
    nms::llr_generate_kernel<<<std::max(grid, 1), threads, 0, st>>>(P, out, n_frames);
    nms_note_launch();
    return cudaGetLastError();
}
