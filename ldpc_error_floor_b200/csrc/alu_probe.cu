// alu_probe.cu -- instruction-issue micro-benchmark of the pipes the decode kernels live on (the ALU
// roofline must be MEASURED on the box, MEASURED_PEAKS.json carries no such figure).  Every thread runs 8 independent
// dependency chains of one instruction kind, written as volatile inline PTX so the compiler neither removes nor fuses
// them (each chain's second operand is its neighbour chain, which also keeps ptxas from folding the sequence); with 2048
// resident threads per SM the chains hide the pipe latency and the rate is the pipe's issue rate.
// Result: lane-ops per second (one "lane-op" = one instruction executed by one of the 32 lanes of a warp, i.e. a packed
// half2 instruction counts once).  kinds: 0 FFMA (FP32/FMA pipe), 1 FMNMX, 2 LOP3, 3 IADD (integer/logic pipe),
// 4 HFMA2, 5 HMNMX2 (packed fp16x2), 6 FADD.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// b is the neighbouring chain's value (so ptxas cannot fold repeated operands: it turns a + b + b + ... into one multiply-add
// and min(min(a, b), b) into a three-input minimum), `odd` alternates min / max so the chains do not become idempotent
template <int KIND>
__device__ __forceinline__ void op(uint32_t &a, uint32_t b, uint32_t c, bool odd) {
    if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    else if (KIND == 1) { if (odd) asm volatile("max.f32 %0, %0, %1;" : "+r"(a) : "r"(b)); else asm volatile("min.f32 %0, %0, %1;" : "+r"(a) : "r"(b)); }
    else if (KIND == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    else if (KIND == 3) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
    else if (KIND == 4) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    else if (KIND == 5) { if (odd) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b)); else asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b)); }
    else asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a) : "r"(b));
}

constexpr int CHAINS = 8, UNROLL = 8;

template <int KIND>
__global__ void __launch_bounds__(256) alu_probe_kernel(uint32_t *out, int trips, uint32_t b, uint32_t c) {
    (void)b;
    uint32_t a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) a[k] = 0x3c003c00u + threadIdx.x * 65537u + k;
    for (int i = 0; i < trips; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int k = 0; k < CHAINS; ++k) op<KIND>(a[k], a[(k + 1) % CHAINS], c, (u & 1) != 0);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
cudaError_t run(int sms, int trips, uint32_t *buf, float *ms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8;
    alu_probe_kernel<KIND><<<grid, 256>>>(buf, trips / 8 + 1, 0x3c003c00u, 0x00010001u);   // warm-up
    cudaEventRecord(e0);
    alu_probe_kernel<KIND><<<grid, 256>>>(buf, trips, 0x3c003c00u, 0x00010001u);
    cudaEventRecord(e1);
    cudaError_t rc = cudaEventSynchronize(e1);
    if (rc == cudaSuccess) rc = cudaGetLastError();
    cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

}   // namespace

// lane-ops per second of instruction kind `kind` on `device`; returns a cudaError_t
extern "C" int nms_alu_probe(int device, int kind, double *lane_ops_per_s) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaError_t rc = cudaSetDevice(device);
    if (rc != cudaSuccess) return (int)rc;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int sms = prop.multiProcessorCount, trips = 4096;
    uint32_t *buf = nullptr;
    rc = cudaMalloc(&buf, (size_t)sms * 8 * 256 * sizeof(uint32_t));
    float ms = 0.0f;
    if (rc == cudaSuccess) {
        switch (kind) {
        case 0: rc = run<0>(sms, trips, buf, &ms); break;
        case 1: rc = run<1>(sms, trips, buf, &ms); break;
        case 2: rc = run<2>(sms, trips, buf, &ms); break;
        case 3: rc = run<3>(sms, trips, buf, &ms); break;
        case 4: rc = run<4>(sms, trips, buf, &ms); break;
        case 5: rc = run<5>(sms, trips, buf, &ms); break;
        case 6: rc = run<6>(sms, trips, buf, &ms); break;
        default: rc = cudaErrorInvalidValue;
        }
    }
    cudaFree(buf);
    cudaSetDevice(prev);
    if (rc != cudaSuccess) return (int)rc;
    const double ops = (double)sms * 8 * 256 * (double)trips * UNROLL * CHAINS;
    *lane_ops_per_s = ops / (ms * 1e-3);
    return 0;
}
