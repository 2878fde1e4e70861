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

}   // namespace nms

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
