#include <cuda_runtime.h>
__global__ void k(float2 *o, const float2 *x, float w, float m) {
    float2 a = x[threadIdx.x];
    float2 p = __fmul2_rn(a, make_float2(w, w));
    float2 q = __fadd2_rn(__fadd2_rn(p, make_float2(m, m)), make_float2(-m, -m));
    o[threadIdx.x] = q;
}
__global__ void k1(float *o, const float *x, float w, float m) {
    float a = x[threadIdx.x];
    float p = __fmul_rn(a, w);
    o[threadIdx.x] = __fadd_rn(__fadd_rn(p, m), -m);
}
