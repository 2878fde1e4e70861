// nms_train.cu -- loss and weight gradients of one batch: the training step of the reference
// (loss Main_Functions.py:337-357 on the forward :161-335), "next" row N1 of SURVEY.md 8(f).
//
// One CTA per frame, float32, messages in shared memory.  Forward pass: T iterations, the messages entering
// every iteration are kept in global memory (hist[t], (T+1) x E*z floats per frame).  Backward pass, t = T-1 ..
// t_lo: reload hist[t], recompute that iteration's intermediates, apply the gradient TensorFlow's autodiff
// assigns to the reference graph:
//   * Cal_MSA_Q_TF (:475-494) / clip_by_value: straight-through, passes where |x| <= bound (inclusive);
//   * reduce_min (:249, :350): split evenly among tied minima;  abs -> sign(x);
//   * tf.sign, tf.to_float(x > 0), comparisons: no gradient (sign product, UCN indicator, ReLU gate, zero rules);
//   * sign_through (:457-460): forward sign, gradient of inv_exp(x) = 2 / (1 + exp(-x)) - 1.
// Arithmetic of the forward part follows the reference order (direct extrinsic sums in ascending E(C) order),
// so the quantised path is bit-exact and the trained decoder is the decoder the fast kernels run.
// Checked against torch.autograd of the reference's own code (tests/golden/grad_*.npz).
#include "nms_train.cuh"
#include <cstdlib>

#define TRAIN_MAX_DC 64
// threads per CTA (= per frame) are a launch-time choice: a frame is one CTA, so the per-frame latency -- what bounds a
// training batch of 20-40 frames on 148 SMs -- falls with the thread count until barriers take over
#define TRAIN_MAX_THREADS 1024
#define TRAIN_THREADS ((int)blockDim.x)

namespace {

__device__ __forceinline__ float qsat(const TrainParams &P, float x) {   // Q() or clip
    if (P.qms) {
        const float r = __fsub_rn(__fadd_rn(fminf(fmaxf(x, -1.0e5f), 1.0e5f), P.qmagic), P.qmagic);
        return fminf(fmaxf(r, -P.qmax), P.qmax);
    }
    return fminf(fmaxf(x, -P.clip), P.clip);
}
__device__ __forceinline__ float sat_bound(const TrainParams &P) { return P.qms ? P.qmax : P.clip; }
__device__ __forceinline__ float sgnf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

__device__ __forceinline__ float w_cn(const TrainParams &P, int t, int i, int e) {
    if (P.sharing0 == 0) return 1.0f;
    return P.w[P.off_cn + t * P.wc + (P.sharing0 == 3 ? 0 : (P.sharing0 == 2 ? i : e))];
}
__device__ __forceinline__ float w_ucn(const TrainParams &P, int t, int i, int e) {
    return P.w[P.off_ucn + t * P.wu + (P.sharing1 == 3 ? 0 : (P.sharing1 == 2 ? i : e))];
}
__device__ __forceinline__ int widx(int code, int node, int e) { return code == 3 ? 0 : (code == 2 ? node : e); }

struct Smem {
    float *xa, *xq, *xin, *ga, *c2v, *x2, *gc, *gx, *gw, *red;
    unsigned char *mk_v2c, *mk_xin, *hb;
};

__device__ __forceinline__ float block_reduce(float v, float *red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : v + u;
    }
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int k = 1; k < TRAIN_THREADS / 32; ++k) r = is_max ? fmaxf(r, red[k]) : r + red[k];
    __syncthreads();
    return r;
}

// S1 + S2 of the forward iteration t from the messages in S.c2v: xin (+ its STE mask), the hard decision of the
// previous APP (for the UCN weights), the saturated V->C messages in the check-lane frame (+ STE mask)
__device__ void vn_forward(const TrainParams &P, const Smem &S, int t) {
    const int tid = threadIdx.x;
    for (int v = tid; v < P.NZ; v += TRAIN_THREADS) {
        const int j = v / P.z, c = v - j * P.z;
        float pre = S.xa[v];
        if (P.sharing2 != 0) pre = __fmul_rn(pre, P.w[P.off_vn + t * P.wv + (P.sharing2 == 3 ? 0 : j)]);   // :168-169
        S.mk_xin[v] = (!P.qms || fabsf(pre) <= P.qmax) ? 1 : 0;
        const float xi = P.qms ? qsat(P, pre) : pre;                                                       // :176-177
        S.xin[v] = xi;
        if (P.sharing1 > 0) {                           // source of the UCN indicator (:181-188): bit = (src >= 0)
            float src = xi;
            if (t > 0) {
                float s = 0.0f;
                for (int k = P.col_ptr[j]; k < P.col_ptr[j + 1]; ++k) s = __fadd_rn(s, S.c2v[P.col_edge[k] * P.z + c]);
                src = __fadd_rn(S.xq[v], s);
            }
            S.hb[v] = src >= 0.0f ? 1 : 0;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < P.EZ; idx += TRAIN_THREADS) {
        const int e = idx / P.z, a = idx - e * P.z;
        const int j = P.col[e];
        int c = a + P.shift[e];
        if (c >= P.z) c -= P.z;
        float acc = 0.0f;
        for (int k = P.col_ptr[j]; k < P.col_ptr[j + 1]; ++k) {
            const int e2 = P.col_edge[k];
            if (e2 != e) acc = __fadd_rn(acc, S.c2v[e2 * P.z + c]);                                         // :214
        }
        const float pre = __fadd_rn(S.xin[j * P.z + c], acc);                                              // :215
        S.mk_v2c[idx] = fabsf(pre) <= sat_bound(P) ? 1 : 0;
        float v = qsat(P, pre);                                                                            // :223-226
        if (v == 0.0f) v = 1.0e-4f;                                                                        // :228
        S.x2[idx] = v;
    }
    __syncthreads();
}

// everything the check node (i, a) derives from its saturated inputs
struct RowInfo {
    float m1, m2;      // smallest and second-smallest magnitude (m2 = smallest over the rest when the minimum is unique)
    int c1, c2;        // multiplicity of m1; of m2
    int npos;          // inputs > 0
    int ucn;           // check unsatisfied by the previous hard decision
};

__device__ __forceinline__ RowInfo row_scan(const TrainParams &P, const Smem &S, int i, int a, int e0, int dc, float *val) {
    RowInfo r;
    r.m1 = 3.0e38f; r.m2 = 3.0e38f; r.c1 = 0; r.c2 = 0; r.npos = 0; r.ucn = 0;
    for (int p = 0; p < dc; ++p) {
        const int e = e0 + p;
        const float v = S.x2[e * P.z + a];
        val[p] = v;
        const float av = fabsf(v);
        if (av < r.m1) { r.m2 = r.m1; r.c2 = r.c1; r.m1 = av; r.c1 = 1; }
        else if (av == r.m1) { ++r.c1; }
        else if (av < r.m2) { r.m2 = av; r.c2 = 1; }
        else if (av == r.m2) { ++r.c2; }
        r.npos += v > 0.0f ? 1 : 0;
        if (P.sharing1 > 0) {
            int c = a + P.shift[e];
            if (c >= P.z) c -= P.z;
            r.ucn ^= S.hb[P.col[e] * P.z + c];
        }
    }
    return r;
}

// C->V of edge p of the row (before weighting): minimum over the others with the 1e-4 rule, sign, tie count
__device__ __forceinline__ void edge_min(const RowInfo &r, int dc, float v, float &m_fixed, float &sig, float &m_raw, int &nt) {
    const float av = fabsf(v);
    if (dc == 1) { m_raw = 10000.0f; nt = 1; }                                   // all-masked row (:248)
    else if (av > r.m1) { m_raw = r.m1; nt = r.c1; }
    else if (r.c1 >= 2) { m_raw = r.m1; nt = r.c1 - 1; }
    else { m_raw = r.m2; nt = r.c2; }
    m_fixed = m_raw > 1.0e-4f ? m_raw : m_raw - 1.0e-4f;                         // :250
    const int npos_others = r.npos - (v > 0.0f ? 1 : 0);
    sig = (npos_others & 1) ? 1.0f : -1.0f;                                      // -prod((-1)^[x>0]) (:251-253)
}

__device__ void cn_forward(const TrainParams &P, const Smem &S, int t) {
    float val[TRAIN_MAX_DC];
    for (int idx = threadIdx.x; idx < P.MZ; idx += TRAIN_THREADS) {
        const int i = idx / P.z, a = idx - i * P.z;
        const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
        const RowInfo r = row_scan(P, S, i, a, e0, dc, val);
        for (int p = 0; p < dc; ++p) {
            const int e = e0 + p;
            float m, sig, m_raw;
            int nt;
            edge_min(r, dc, val[p], m, sig, m_raw, nt);
            const float x0 = __fmul_rn(m, sig);                                  // :254
            const float mag = fabsf(x0);
            float x1 = mag;
            if (P.sharing0 != 0) x1 = __fmul_rn(mag, (P.sharing1 == P.sharing0 && r.ucn) ? w_ucn(P, t, i, e) : w_cn(P, t, i, e));
            const float x2o = x1 > 0.0f ? x1 : 0.0f;                             // :308
            int c = a + P.shift[e];
            if (c >= P.z) c -= P.z;
            S.c2v[e * P.z + c] = __fmul_rn(qsat(P, x2o), sgnf(x0));              // :310-316, lifted back (:259-263)
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TRAIN_MAX_THREADS) nms_train_kernel(const TrainParams P) {
    extern __shared__ __align__(16) unsigned char train_smem[];
    Smem S;
    float *f = reinterpret_cast<float *>(train_smem);
    S.xa = f; f += P.NZ; S.xq = f; f += P.NZ; S.xin = f; f += P.NZ; S.ga = f; f += P.NZ;
    S.c2v = f; f += P.EZ; S.x2 = f; f += P.EZ; S.gc = f; f += P.EZ; S.gx = f; f += P.EZ;
    const int gwn = P.wc + P.wu + P.wv;
    S.gw = f; f += gwn > 0 ? gwn : 1; S.red = f; f += 32;
    unsigned char *b = reinterpret_cast<unsigned char *>(f);
    S.mk_v2c = b; b += P.EZ; S.mk_xin = b; b += P.NZ; S.hb = b;

    const int tid = threadIdx.x;
    const long long frame = blockIdx.x;
    float *hist = P.hist + frame * (long long)(P.T + 1) * P.EZ;
    for (int v = tid; v < P.NZ; v += TRAIN_THREADS) {
        const float x = P.llr[frame * P.NZ + v];
        S.xa[v] = x;
        S.xq[v] = P.qms ? qsat(P, x) : x;                                        // :321-322
    }
    for (int idx = tid; idx < P.EZ; idx += TRAIN_THREADS) { S.c2v[idx] = 0.0f; S.gc[idx] = 0.0f; }   // LLRa0 (main_Base.py:126)
    for (int k = tid; k < gwn; k += TRAIN_THREADS) S.gw[k] = 0.0f;
    __syncthreads();

    // ---------------------------------------------------------------- forward, keeping the messages per iteration
    for (int t = 0; t < P.T; ++t) {
        for (int idx = tid; idx < P.EZ; idx += TRAIN_THREADS) hist[(long long)t * P.EZ + idx] = S.c2v[idx];
        vn_forward(P, S, t);
        cn_forward(P, S, t);
        if (P.app_out != nullptr) {
            for (int v = tid; v < P.NZ; v += TRAIN_THREADS) {
                const int j = v / P.z, c = v - j * P.z;
                float s = 0.0f;
                for (int k = P.col_ptr[j]; k < P.col_ptr[j + 1]; ++k) s = __fadd_rn(s, S.c2v[P.col_edge[k] * P.z + c]);
                P.app_out[((long long)t * P.B + frame) * P.NZ + v] = fminf(fmaxf(__fadd_rn(S.xq[v], s), -P.clip), P.clip);
            }
        }
    }
    for (int idx = tid; idx < P.EZ; idx += TRAIN_THREADS) hist[(long long)P.T * P.EZ + idx] = S.c2v[idx];
    __syncthreads();

    // ---------------------------------------------------------------- backward
    double loss_frame = 0.0;
    float val[TRAIN_MAX_DC], gabs[TRAIN_MAX_DC];
    for (int t = P.T - 1; t >= P.t_lo; --t) {
        const float *next = hist + (long long)(t + 1) * P.EZ;     // c2v_{t+1}
        const float coef = P.coef[t];
        // ---- APP_t, its loss term and d loss / d APP (masked by the clip of :324)
        float vmax = -3.0e38f;
        for (int v = tid; v < P.NZ; v += TRAIN_THREADS) {
            const int j = v / P.z, c = v - j * P.z;
            float s = 0.0f;
            for (int k = P.col_ptr[j]; k < P.col_ptr[j + 1]; ++k) s = __fadd_rn(s, next[P.col_edge[k] * P.z + c]);
            const float raw = __fadd_rn(S.xq[v], s);
            const float app = fminf(fmaxf(raw, -P.clip), P.clip);
            S.xin[v] = app;                                        // scratch: xin is recomputed below
            S.mk_xin[v] = fabsf(raw) <= P.clip ? 1 : 0;            // scratch likewise
            if (v < P.target_nz) vmax = fmaxf(vmax, app);
        }
        float nties = 1.0f, dfer = 0.0f;
        if (P.loss_type == 2) {
            vmax = block_reduce(vmax, S.red, true);
            float cnt = 0.0f;
            for (int v = tid; v < P.target_nz; v += TRAIN_THREADS) cnt += S.xin[v] == vmax ? 1.0f : 0.0f;
            nties = block_reduce(cnt, S.red, false);
            const float m = -vmax;                                 // reduce_min(-x) (:350)
            const float em = expf(-m);
            dfer = 0.5f * (2.0f * em / ((1.0f + em) * (1.0f + em)));   // -d/dm [1/2 (1 - inv_exp(m))]
            if (tid == 0 && coef != 0.0f) loss_frame += (double)coef * 0.5 * (1.0 - (double)sgnf(m)) / P.B;
        } else {
            __syncthreads();
        }
        float lsum = 0.0f;
        for (int v = tid; v < P.NZ; v += TRAIN_THREADS) {
            float g = 0.0f;
            if (v < P.target_nz && coef != 0.0f) {
                const float x = S.xin[v];
                if (P.loss_type == 0) {           // softplus(x); derivative sigmoid(x)
                    lsum += fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
                    g = coef / ((float)P.B * (float)P.target_nz) / (1.0f + expf(-x));
                } else if (P.loss_type == 1) {    // sigmoid(x)
                    const float sg = 1.0f / (1.0f + expf(-x));
                    lsum += sg;
                    g = coef / ((float)P.B * (float)P.target_nz) * sg * (1.0f - sg);
                } else if (x == vmax) {
                    g = coef / (float)P.B * dfer / nties;
                }
            }
            S.ga[v] = S.mk_xin[v] ? g : 0.0f;
        }
        if (P.loss_type != 2) {
            lsum = block_reduce(lsum, S.red, false);
            if (tid == 0) loss_frame += (double)coef * (double)lsum / ((double)P.B * (double)P.target_nz);
        }
        __syncthreads();
        // ---- recompute iteration t from the messages that entered it
        for (int idx = tid; idx < P.EZ; idx += TRAIN_THREADS) S.c2v[idx] = hist[(long long)t * P.EZ + idx];
        __syncthreads();
        vn_forward(P, S, t);
        // ---- d loss / d c2v_{t+1} = later iterations (gc) + this iteration's APP
        for (int idx = tid; idx < P.EZ; idx += TRAIN_THREADS) {
            const int e = idx / P.z, c = idx - e * P.z;
            S.gc[idx] += S.ga[P.col[e] * P.z + c];
        }
        __syncthreads();
        // ---- check nodes: weighting / ReLU / saturation backward, weight gradients, tie-split minimum backward
        for (int idx = tid; idx < P.MZ; idx += TRAIN_THREADS) {
            const int i = idx / P.z, a = idx - i * P.z;
            const int e0 = P.row_ptr[i], dc = P.row_ptr[i + 1] - e0;
            const RowInfo r = row_scan(P, S, i, a, e0, dc, val);
            const bool use_ucn = P.sharing1 == P.sharing0 && P.sharing0 != 0 && r.ucn;
            float s_gt = 0.0f, s_m1 = 0.0f, g_single = 0.0f;    // sums of g_m over: inputs above the minimum, the minimum class
            float acc_w = 0.0f;                                  // weight gradient of this row when it has ONE weight
            for (int p = 0; p < dc; ++p) {
                const int e = e0 + p;
                float m, sig, m_raw;
                int nt;
                edge_min(r, dc, val[p], m, sig, m_raw, nt);
                const float x0 = m * sig, mag = fabsf(x0), sx0 = sgnf(x0);
                const float w = P.sharing0 == 0 ? 1.0f : (use_ucn ? w_ucn(P, t, i, e) : w_cn(P, t, i, e));
                const float x1 = P.sharing0 == 0 ? mag : __fmul_rn(mag, w);
                const float gate = x1 > 0.0f ? 1.0f : 0.0f;
                const float omask = (gate ? x1 : 0.0f) <= sat_bound(P) ? 1.0f : 0.0f;
                int c = a + P.shift[e];
                if (c >= P.z) c -= P.z;
                const float g_x1 = S.gc[e * P.z + c] * sx0 * omask * gate;
                if (P.sharing0 == 1) {
                    if (g_x1 != 0.0f) atomicAdd(&S.gw[(use_ucn ? P.wc : 0) + e], g_x1 * mag);
                } else {
                    acc_w += g_x1 * mag;
                }
                const float g_m = g_x1 * w * sx0 * sig;          // |x0| -> x0 -> m (x0 = m * sign(-prod))
                gabs[p] = g_m;                                    // parked; turned into d/d|input| below
                if (dc > 1) {
                    const float av = fabsf(val[p]);
                    if (av > r.m1) s_gt += g_m;
                    else { s_m1 += g_m; g_single = g_m; }
                }
            }
            if (P.sharing0 >= 2 && acc_w != 0.0f)
                atomicAdd(&S.gw[(use_ucn ? P.wc : 0) + widx(P.sharing0, i, 0)], acc_w);
            for (int p = 0; p < dc; ++p) {
                const float av = fabsf(val[p]);
                float ga = 0.0f;
                if (dc > 1) {
                    if (av == r.m1) {
                        ga = s_gt / (float)r.c1;                              // targets above the minimum see all of the class
                        if (r.c1 >= 2) ga += (s_m1 - gabs[p]) / (float)(r.c1 - 1);   // the other members of the class
                    } else if (av == r.m2 && r.c1 == 1) {
                        ga = g_single / (float)r.c2;                          // the unique minimum's own output sees the runner-up class
                    }
                }
                const int e = e0 + p;
                int c = a + P.shift[e];
                if (c >= P.z) c -= P.z;
                S.gx[e * P.z + c] = S.mk_v2c[e * P.z + a] ? ga * sgnf(val[p]) : 0.0f;   // abs, then the STE of :223-226, un-lifted
            }
        }
        __syncthreads();
        // ---- variable nodes: v2c_e = xin + sum of the OTHER incoming messages
        for (int v = tid; v < P.NZ; v += TRAIN_THREADS) {
            const int j = v / P.z, c = v - j * P.z;
            float G = 0.0f;
            for (int k = P.col_ptr[j]; k < P.col_ptr[j + 1]; ++k) G += S.gx[P.col_edge[k] * P.z + c];
            for (int k = P.col_ptr[j]; k < P.col_ptr[j + 1]; ++k) {
                const int e = P.col_edge[k];
                S.gc[e * P.z + c] = G - S.gx[e * P.z + c];        // gradient entering iteration t - 1
            }
            if (P.sharing2 != 0) {
                const float gwv = S.mk_xin[v] ? G * S.xa[v] : 0.0f;            // STE of :176-177
                if (gwv != 0.0f) atomicAdd(&S.gw[P.wc + P.wu + (P.sharing2 == 3 ? 0 : j)], gwv);
            }
        }
        __syncthreads();
        for (int k = tid; k < gwn; k += TRAIN_THREADS) {
            const float gsum = S.gw[k];
            S.gw[k] = 0.0f;
            if (gsum != 0.0f) {
                const int off = k < P.wc ? P.off_cn + t * P.wc + k
                              : (k < P.wc + P.wu ? P.off_ucn + t * P.wu + (k - P.wc) : P.off_vn + t * P.wv + (k - P.wc - P.wu));
                atomicAdd(P.grad + off, gsum);
            }
        }
        __syncthreads();
    }
    if (tid == 0 && loss_frame != 0.0) atomicAdd(P.loss, loss_frame);
}

}   // namespace

size_t nms_train_smem_bytes(const TrainParams &P) {
    const int gwn = P.wc + P.wu + P.wv;
    return (size_t)(4 * P.NZ + 4 * P.EZ + (gwn > 0 ? gwn : 1) + 32) * sizeof(float) + (size_t)P.EZ + 2 * (size_t)P.NZ + 16;
}

cudaError_t nms_launch_train(const TrainParams &P, cudaStream_t st) {
    const size_t smem = nms_train_smem_bytes(P);
    cudaError_t e = cudaFuncSetAttribute(nms_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // measured (tools/train_bench.py, profiles/r02_train_step.txt): while every frame can have an SM of its own, 1024 threads
    // per frame are fastest (z72 batch 40: 2.35 ms at 256, 1.57 at 512, 1.23 at 1024); beyond that 512 (WiMAX batch 200)
    static const int forced = [] {
        const char *e = getenv("LDPC_B200_TRAIN_THREADS");
        const int v = e ? atoi(e) : 0;
        return (v >= 32 && v <= TRAIN_MAX_THREADS && v % 32 == 0) ? v : 0;
    }();
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = forced ? forced : (P.B <= sms ? 1024 : 512);
    nms_train_kernel<<<P.B, threads, smem, st>>>(P);
    nms_note_launch();
    return cudaGetLastError();
}
