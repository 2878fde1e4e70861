// host_pack.cpp -- host side of the int8 transport of ldpc_decode_host.
//
// The reference hands `xa_input` to its graph as float32 (Print_Functions.py:45-46, 130-165); a quantised min-sum decoder
// sees it only through Q(xa) (Main_Functions.py:321-322, 475-494) and, with VN weights, Q(xa * w) (:168-177).  Q has at most
// 31 levels, so the words can cross PCIe as one byte per value whenever that loses nothing -- a quarter of the float32 bytes,
// which are what bounds the end-to-end path.  This file is the host half: a vectorised float32 -> int8 pass with an exact
// "is this lossless?" verdict, and the small thread pool that runs it (and the staging copies of pageable input) while the
// GPU decodes the previous chunk.  Plain C++ (no CUDA): compiled by g++ so the AVX2 body can use target attributes.
#include "host_pack.h"

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <system_error>
#include <thread>
#include <vector>

#include <unistd.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define HOSTPACK_X86 1
#else
#define HOSTPACK_X86 0
#endif

namespace hostpack {
namespace {

constexpr float XA_BOUND = 1.0e5f;   // nms_device.cuh: inputs of a quantised decoder are clamped here first

int64_t pack_scalar(const float *x, int64_t n, float qk, float kmax, int lossless, int8_t *out) {
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        float v = x[i];
        if (!lossless) {
            v = v > -XA_BOUND ? v : -XA_BOUND;   // NaN -> -bound, as fmaxf(NaN, -b) on the device
            v = v < XA_BOUND ? v : XA_BOUND;
            float k = nearbyintf(v * qk);        // round half to even (default rounding mode); v * qk is exact
            k = k > -kmax ? k : -kmax;
            k = k < kmax ? k : kmax;
            out[i] = (int8_t)k;
        } else {
            const float y = v * qk;
            const bool in = std::fabs(y) <= kmax;   // false for NaN
            const float k = in ? nearbyintf(y) : 0.0f;
            const bool ok = in && k == y;
            bad += !ok;
            out[i] = (int8_t)k;
        }
    }
    return bad;
}

#if HOSTPACK_X86
__attribute__((target("avx2"))) inline __m256i pack32(__m256i a, __m256i b, __m256i c, __m256i d) {
    const __m256i ab = _mm256_packs_epi32(a, b), cd = _mm256_packs_epi32(c, d);
    const __m256i q = _mm256_packs_epi16(ab, cd);
    return _mm256_permutevar8x32_epi32(q, _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7));
}

__attribute__((target("avx2"))) int64_t pack_avx2(const float *x, int64_t n, float qk, float kmax, int lossless, int8_t *out) {
    const __m256 vqk = _mm256_set1_ps(qk), vkmax = _mm256_set1_ps(kmax);
    const __m256 lo = _mm256_set1_ps(-XA_BOUND), hi = _mm256_set1_ps(XA_BOUND);
    const __m256i ikmax = _mm256_set1_epi32((int)kmax), ikmin = _mm256_set1_epi32(-(int)kmax);
    const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
    int64_t bad = 0;
    int64_t i = 0;
    if (!lossless) {
        for (; i + 32 <= n; i += 32) {
            __m256i k[4];
            for (int u = 0; u < 4; ++u) {
                __m256 v = _mm256_loadu_ps(x + i + 8 * u);
                v = _mm256_min_ps(_mm256_max_ps(v, lo), hi);   // max_ps(NaN, lo) = lo: the device's fmaxf(NaN, -b)
                const __m256i r = _mm256_cvtps_epi32(_mm256_mul_ps(v, vqk));
                k[u] = _mm256_min_epi32(_mm256_max_epi32(r, ikmin), ikmax);
            }
            _mm256_storeu_si256((__m256i *)(out + i), pack32(k[0], k[1], k[2], k[3]));
        }
    } else {
        __m256i nbad = _mm256_setzero_si256();
        for (; i + 32 <= n; i += 32) {
            __m256i k[4];
            for (int u = 0; u < 4; ++u) {
                const __m256 y = _mm256_mul_ps(_mm256_loadu_ps(x + i + 8 * u), vqk);
                k[u] = _mm256_cvtps_epi32(y);
                const __m256 ok = _mm256_and_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(k[u]), y, _CMP_EQ_OQ),
                                                _mm256_cmp_ps(_mm256_and_ps(y, absmask), vkmax, _CMP_LE_OQ));
                nbad = _mm256_sub_epi32(nbad, _mm256_andnot_si256(_mm256_castps_si256(ok), _mm256_set1_epi32(-1)));
            }
            _mm256_storeu_si256((__m256i *)(out + i), pack32(k[0], k[1], k[2], k[3]));
        }
        alignas(32) int32_t lanes[8];
        _mm256_store_si256((__m256i *)lanes, nbad);   // <= n / 8 per lane: no overflow below 2^34 values per call
        for (int u = 0; u < 8; ++u) bad += (uint32_t)lanes[u];
    }
    return bad + pack_scalar(x + i, n - i, qk, kmax, lossless, out + i);
}

bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2") && !getenv("LDPC_B200_NO_AVX2");
    return v;
}
#endif

// ------------------------------------------------------------------------------------------------ thread pool
class Pool {
  public:
    Pool() {
        int n = 0;
        if (const char *e = getenv("LDPC_B200_HOST_THREADS")) n = atoi(e);
        if (n <= 0) {
            int hw = (int)std::thread::hardware_concurrency();
            if (hw <= 0) hw = 1;
            int local = 1;
            if (const char *e = getenv("LOCAL_WORLD_SIZE")) local = std::max(1, atoi(e));
            // one hardware thread stays with the thread that issues the CUDA calls and waits for the device (it spins)
            n = std::min(16, std::max(1, hw / local - 1));
        }
        nthreads_ = std::min(n, 64);
        try {
            for (int t = 1; t < nthreads_; ++t) workers_.emplace_back([this] { work(); });
        } catch (...) {   // thread limit of the process: go on with the workers that did start
        }
        nthreads_ = (int)workers_.size() + 1;
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int threads() const { return nthreads_; }

    void run(int64_t nblocks, void (*fn)(void *, int64_t), void *ctx) {
        if (nblocks <= 0) return;
        if (nthreads_ == 1 || nblocks == 1) {
            for (int64_t b = 0; b < nblocks; ++b) fn(ctx, b);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mu_);
        fn_ = fn; ctx_ = ctx; nblocks_ = nblocks;
        next_.store(0, std::memory_order_relaxed);
        pending_.store((int)workers_.size(), std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu_);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        drain_blocks();
        // every worker checks in once per generation, so the job fields are free to change after this
        for (int spin = 0; pending_.load(std::memory_order_acquire) != 0; ++spin) {
            if (spin < 2000) cpu_relax(); else std::this_thread::yield();
        }
    }

  private:
    static void cpu_relax() {
#if HOSTPACK_X86
        _mm_pause();
#endif
    }
    void drain_blocks() {
        for (;;) {
            const int64_t b = next_.fetch_add(1, std::memory_order_relaxed);
            if (b >= nblocks_) break;
            fn_(ctx_, b);
        }
    }
    void work() {
        uint64_t seen = 0;
        for (;;) {
            // chunks of one call arrive every few hundred microseconds: spin for about that long before sleeping
            bool got = false;
            for (int spin = 0; spin < 20000; ++spin) {
                if (gen_.load(std::memory_order_acquire) != seen) { got = true; break; }
                cpu_relax();
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_) return;
            drain_blocks();
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }

    int nthreads_ = 1;
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int64_t> next_{0};
    std::atomic<int> pending_{0};
    bool stop_ = false;
    void (*fn_)(void *, int64_t) = nullptr;
    void *ctx_ = nullptr;
    int64_t nblocks_ = 0;
};

Pool &pool() {
    // never destroyed: worker threads must not be joined from a static destructor at exit.  A forked child (a Python
    // multiprocessing worker) inherits the pointer but none of the threads: it gets a pool of its own.
    static std::mutex mu;
    static Pool *p = nullptr;
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(mu);
    if (p == nullptr || owner != getpid()) {
        p = new Pool();
        owner = getpid();
    }
    return *p;
}

constexpr int64_t BLOCK_BYTES = 256 << 10;

struct PackJob {
    const float *x; int64_t n; float qk, kmax; int lossless; int8_t *out;
    bool stop_on_bad;   // the caller only wants to know WHETHER all values are encodable: skip the rest after a miss
    std::atomic<int64_t> bad{0};
};
void pack_block(void *ctx, int64_t b) {
    PackJob &j = *(PackJob *)ctx;
    if (j.stop_on_bad && j.bad.load(std::memory_order_relaxed) != 0) return;
    const int64_t per = BLOCK_BYTES / 4, lo = b * per, hi = std::min(j.n, lo + per);
    const int64_t bad = pack_q8(j.x + lo, hi - lo, j.qk, j.kmax, j.lossless, j.out + lo);
    if (bad) j.bad.fetch_add(bad, std::memory_order_relaxed);
}
struct CopyJob { char *dst; const char *src; size_t bytes; };
void copy_block(void *ctx, int64_t b) {
    CopyJob &j = *(CopyJob *)ctx;
    const size_t lo = (size_t)b * BLOCK_BYTES, hi = std::min(j.bytes, lo + (size_t)BLOCK_BYTES);
    std::memcpy(j.dst + lo, j.src + lo, hi - lo);
}

}   // namespace

int64_t pack_q8(const float *x, int64_t n, float qk, float kmax, int lossless, int8_t *out) {
#if HOSTPACK_X86
    if (have_avx2()) return pack_avx2(x, n, qk, kmax, lossless, out);
#endif
    return pack_scalar(x, n, qk, kmax, lossless, out);
}

void parallel_blocks(int64_t nblocks, void (*fn)(void *, int64_t), void *ctx) { pool().run(nblocks, fn, ctx); }
int pool_threads() { return pool().threads(); }

int64_t pack_q8_mt(const float *x, int64_t n, float qk, float kmax, int lossless, int8_t *out, bool stop_on_bad) {
    PackJob j;
    j.stop_on_bad = stop_on_bad;
    j.x = x; j.n = n; j.qk = qk; j.kmax = kmax; j.lossless = lossless; j.out = out;
    parallel_blocks((n * 4 + BLOCK_BYTES - 1) / BLOCK_BYTES, pack_block, &j);
    return j.bad.load();
}

void memcpy_mt(void *dst, const void *src, size_t bytes) {
    CopyJob j{(char *)dst, (const char *)src, bytes};
    parallel_blocks((int64_t)((bytes + BLOCK_BYTES - 1) / BLOCK_BYTES), copy_block, &j);
}

}   // namespace hostpack
