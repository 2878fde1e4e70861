// host_pack.h -- host side of the int8 transport of ldpc_decode_host (host_pack.cpp, plain C++: compiled by g++).
#pragma once
#include <cstddef>
#include <cstdint>

namespace hostpack {

// Writes out[i] = k_i with k_i * (1/qk) the decoder's view of x[i]; returns how many values have NO such k (the caller then
// ships the words as float32).  qk is a power of two (2, 1, 0.5), kmax = qmax * qk <= 127.
//   lossless = 0: k = clamp(rint(clamp(x, +-1e5) * qk), +-kmax) -- exactly Q(x) of Main_Functions.py:475-494 as the kernels
//                 compute it (nms_device.cuh qf), NaN -> -kmax like fminf(fmaxf(NaN, -b), b); always 0 returned.  For decoders
//                 that use the channel value only through Q(x) (no VN weights).
//   lossless = 1: a value is encodable iff it is on the grid and inside +-qmax (x == k / qk exactly; -0.0 counts as 0, which
//                 every use of x maps to the same result).  For decoders that also form Q(x * w) from the raw value.
int64_t pack_q8(const float *x, int64_t n, float qk, float kmax, int lossless, int8_t *out);

// Process-wide pool of host threads for the per-chunk work of the host-buffer entry points (packing, staging copies).
// `fn(ctx, b)` runs once for every block b in [0, nblocks) on the pool's workers and the calling thread; returns when all
// blocks are done.  Calls from different threads are serialised.  Size: LDPC_B200_HOST_THREADS, else hardware threads /
// LOCAL_WORLD_SIZE (torchrun) - 1, at most 16.
void parallel_blocks(int64_t nblocks, void (*fn)(void *ctx, int64_t block), void *ctx);
int pool_threads();

// parallel helpers built on it (blocks of 256 KiB of input)
// stop_on_bad: the return value is then only "zero or not"
int64_t pack_q8_mt(const float *x, int64_t n, float qk, float kmax, int lossless, int8_t *out, bool stop_on_bad = false);
void memcpy_mt(void *dst, const void *src, size_t bytes);

}   // namespace hostpack
