// nms_jit.h -- run-time graph specialisation (nms_jit.cu): NVRTC-compiled graph-specialised kernels for base graphs the
// build does not know, cached on disk, loaded and launched through the driver API.
#pragma once
#include <cuda_runtime.h>

struct KParams;
enum { NMS_JIT_DECODE = 0, NMS_JIT_MCP = 1 };   // packed decode kernel (nms_h2_spec.cuh) / persistent-slot Monte-Carlo kernel (nms_mcp.cuh)

extern "C" int nms_jit_available(void);          // 1 if libnvrtc can be loaded
// (Fp, R) for a graph without a measured launch geometry
void nms_jit_pick_geometry(int M, int N, int E, int z, int kind, int max_dc, int *Fp, int *R);
// Compile (or fetch from the cache) and load: *cufunction receives a CUfunction.  cufunction == nullptr: only make sure the
// cubin is in the cache (needs no device).  0 on success; err receives a one-line reason otherwise.
int nms_jit_build(const int *proto, int M, int N, int z, int Fp, int R, int kind, void **cufunction, char *err, int errcap);
int nms_jit_set_smem(void *cufunction, int bytes);
int nms_jit_occupancy(void *cufunction, int threads, int smem, int *ctas_per_sm);
int nms_jit_launch(void *cufunction, int grid, int threads, int smem, cudaStream_t st, const KParams *P);
