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
