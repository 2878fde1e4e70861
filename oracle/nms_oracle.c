/*
 * Oracle B (C twin) -- plain-C restatement of the reference decode path.
 * TEST INFRASTRUCTURE ONLY: linked/called only from tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  The product never loads it.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks this file against the
 * golden vectors minted from the reference's own Main_Functions.build_neural_network
 * (tests/golden/make_golden.py), and against oracle/nms_oracle.py.
 *
 * Restates, in float32 like the TF graph:
 *   Main_Functions.py:161-177   VN weight + input quantiser          (step D2)
 *   Main_Functions.py:180-209   unsatisfied-check indicator          (step D3)
 *   Main_Functions.py:213-221   VN update, direct extrinsic sum      (step D4)
 *   Main_Functions.py:223-230   saturation, zero -> +1e-4            (step D5)
 *   Main_Functions.py:231-254   min-sum check update, 1e-4 rule      (step D6)
 *   Main_Functions.py:259-316   weighting, ReLU, saturation, sign    (step D7)
 *   Main_Functions.py:317-327   APP                                   (step D8)
 *   Main_Functions.py:475-494   quantiser Cal_MSA_Q_TF
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/build.py).
 * No fast-math, no FMA contraction: products are rounded before they are summed,
 * exactly as separate TF ops do.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* Main_Functions.py:483-492; rintf = round-half-to-even = tf.round */
static inline float quant(float x, int q_bit) {
    switch (q_bit) {
    case 6:  return clipf(rintf(x), -15.5f, 15.5f);
    case 5:  return clipf(rintf(x * 2.0f) / 2.0f, -7.5f, 7.5f);
    case -5: return clipf(rintf(x), -15.0f, 15.0f);
    case 4:  return clipf(rintf(x), -7.0f, 7.0f);
    case 3:  return clipf(rintf(x / 2.0f) * 2.0f, -6.0f, 6.0f);
    default: return x;
    }
}

typedef struct {
    int M, N, z, E;
    int *row, *col, *shift;     /* E(C) order: Main_Functions.py:69-75 */
    int *row_ptr;               /* edges of row i are [row_ptr[i], row_ptr[i+1]) */
    int *col_ptr, *col_edge;    /* edges of column j, ascending E(C) index */
} graph_t;

static int graph_build(graph_t *g, const int32_t *proto, int M, int N, int z) {
    int E = 0;
    for (int k = 0; k < M * N; ++k) E += proto[k] != -1;
    g->M = M; g->N = N; g->z = z; g->E = E;
    g->row = malloc(sizeof(int) * E); g->col = malloc(sizeof(int) * E); g->shift = malloc(sizeof(int) * E);
    g->row_ptr = calloc(M + 1, sizeof(int)); g->col_ptr = calloc(N + 1, sizeof(int));
    g->col_edge = malloc(sizeof(int) * E);
    if (!g->row || !g->col || !g->shift || !g->row_ptr || !g->col_ptr || !g->col_edge) return -1;
    int e = 0;
    for (int i = 0; i < M; ++i) {
        g->row_ptr[i] = e;
        for (int j = 0; j < N; ++j)
            if (proto[i * N + j] != -1) {
                g->row[e] = i; g->col[e] = j;
                g->shift[e] = ((proto[i * N + j] % z) + z) % z;   /* :72 */
                ++e;
            }
    }
    g->row_ptr[M] = E;
    for (e = 0; e < E; ++e) g->col_ptr[g->col[e] + 1]++;
    for (int j = 0; j < N; ++j) g->col_ptr[j + 1] += g->col_ptr[j];
    int *fill = calloc(N, sizeof(int));
    for (e = 0; e < E; ++e) { int j = g->col[e]; g->col_edge[g->col_ptr[j] + fill[j]++] = e; }
    free(fill);
    return 0;
}

static void graph_free(graph_t *g) {
    free(g->row); free(g->col); free(g->shift); free(g->row_ptr); free(g->col_ptr); free(g->col_edge);
}

/* width of one weight row for a sharing code (Main_Functions.py:397-405) */
static int wwidth(int code, int kind, int M, int N, int E) {
    if (code == 1 || code == 4) return E;
    if (code == 2 || code == 5) return kind == 2 ? N : M;
    if (code == 3) return 1;
    return 0;
}

static inline float cn_weight(const float *w, int code, int t, int width, int e, int i) {
    if (code == 3) return w[t * width];
    if (code == 2) return w[t * width + i];
    return w[t * width + e];   /* code 1: per edge, E(C) order */
}

/*
 * Decode B frames, T iterations each, no early termination (the reference has none).
 *   proto [M,N] int32, -1 = no edge;  sharing[3] = {CN, UCN, VN} codes
 *   w_cn/w_ucn/w_vn: [T, width] float32 for the non-zero codes (else ignored)
 *   xa [B,N,z] float32 (log p1/p0)
 * Outputs (each may be NULL):
 *   app_all  [T,B,N*z]  = ya_output{t}   (Main_Functions.py:327)
 *   app_last [B,N*z]
 *   synd     [T,B]      1 if any check is unsatisfied by (APP_t >= 0)
 *   c2v_all  [T,B,E,z]  = LLRa{t+1} in the VN lane frame, E(C) order
 * Returns 0, or a negative code on bad arguments / allocation failure.
 */
int nms_oracle_decode(const int32_t *proto, int M, int N, int z, const int *sharing,
                      const float *w_cn, const float *w_ucn, const float *w_vn, int T,
                      int decoding_type, int q_bit, float clip_llr, const float *xa, int B,
                      float *app_all, float *app_last, uint8_t *synd, float *c2v_all, int nthreads) {
    if (decoding_type != 1 && decoding_type != 2) return -2;
    if (sharing[1] != 0 && sharing[1] != sharing[0]) return -3;      /* :519-521 */
    graph_t g;
    if (graph_build(&g, proto, M, N, z) != 0) return -1;
    const int E = g.E, qms = decoding_type == 2;
    const int wc = wwidth(sharing[0], 0, M, N, E), wu = wwidth(sharing[1], 1, M, N, E),
              wv = wwidth(sharing[2], 2, M, N, E);
    int status = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        float *c2v = malloc(sizeof(float) * E * z), *v2c = malloc(sizeof(float) * E * z);
        float *xin = malloc(sizeof(float) * N * z), *xq = malloc(sizeof(float) * N * z);
        float *app = malloc(sizeof(float) * N * z);
        uint8_t *ucn = malloc((size_t)M * z);
        if (!c2v || !v2c || !xin || !xq || !app || !ucn) {
#pragma omp atomic write
            status = -1;
        } else {
#pragma omp for schedule(dynamic, 4)
        for (int b = 0; b < B; ++b) {
            const float *x = xa + (size_t)b * N * z;
            memset(c2v, 0, sizeof(float) * E * z);                          /* LLRa0 = 0 */
            for (int k = 0; k < N * z; ++k) xq[k] = qms ? quant(x[k], q_bit) : x[k];  /* :321-322 */
            for (int t = 0; t < T; ++t) {
                /* D2 */
                for (int j = 0; j < N; ++j) {
                    float w = 1.0f; int use = sharing[2] == 2 || sharing[2] == 3;
                    if (use) w = sharing[2] == 3 ? w_vn[t * wv] : w_vn[t * wv + j];
                    for (int c = 0; c < z; ++c) {
                        float v = use ? x[j * z + c] * w : x[j * z + c];
                        xin[j * z + c] = qms ? quant(v, q_bit) : v;
                    }
                }
                /* D3: parity of the previous hard decision, per lifted check (CN lane frame) */
                if (sharing[1] > 0) {
                    const float *src = t == 0 ? xin : app;
                    for (int i = 0; i < M; ++i)
                        for (int a = 0; a < z; ++a) {
                            int par = 0;
                            for (int e = g.row_ptr[i]; e < g.row_ptr[i + 1]; ++e) {
                                float s = src[g.col[e] * z + (a + g.shift[e]) % z];
                                par ^= !(-s > 0.0f);                         /* :188: bit = (src >= 0) */
                            }
                            ucn[i * z + a] = (uint8_t)par;
                        }
                } else memset(ucn, 0, (size_t)M * z);
                /* D4 + D5: v2c'[e][a] in the CN lane frame */
                for (int e = 0; e < E; ++e) {
                    const int j = g.col[e], s = g.shift[e];
                    for (int a = 0; a < z; ++a) {
                        const int c = (a + s) % z;
                        float acc = 0.0f;
                        for (int q = g.col_ptr[j]; q < g.col_ptr[j + 1]; ++q) {
                            int e2 = g.col_edge[q];
                            if (e2 != e) acc = acc + c2v[e2 * z + c];
                        }
                        float v = xin[j * z + c] + acc;                      /* :215 */
                        v = qms ? quant(v, q_bit) : clipf(v, -clip_llr, clip_llr);   /* :223-226 */
                        v = v + 0.0001f * (1.0f - (fabsf(v) > 0.0f ? 1.0f : 0.0f));  /* :230 */
                        v2c[e * z + a] = v;
                    }
                }
                /* D6 + D7 */
                for (int i = 0; i < M; ++i) {
                    const int e0 = g.row_ptr[i], e1 = g.row_ptr[i + 1];
                    for (int a = 0; a < z; ++a) {
#ifdef NMS_ORACLE_NAIVE
                        for (int e = e0; e < e1; ++e) {
                            float m = 10000.0f; int npos = 0;                /* :248 masked -> 10000 */
                            for (int e2 = e0; e2 < e1; ++e2) if (e2 != e) {
                                float av = fabsf(v2c[e2 * z + a]);
                                if (av < m) m = av;
                                npos += v2c[e2 * z + a] > 0.0f;
                            }
#else
                        float m1 = 10000.0f, m2 = 10000.0f; int npos_all = 0;
                        for (int e = e0; e < e1; ++e) {
                            float av = fabsf(v2c[e * z + a]);
                            if (av < m1) { m2 = m1; m1 = av; } else if (av < m2) m2 = av;
                            npos_all += v2c[e * z + a] > 0.0f;
                        }
                        for (int e = e0; e < e1; ++e) {
                            float av = fabsf(v2c[e * z + a]);
                            float m = av > m1 ? m1 : m2;                     /* min over the others */
                            int npos = npos_all - (v2c[e * z + a] > 0.0f);
#endif
                            m = m + -0.0001f * (1.0f - (fabsf(m) > 0.0001f ? 1.0f : 0.0f));  /* :250 */
                            /* :251-254: prod over others of (v>0 ? -1 : +1), negated, sign() */
                            float sg = (npos & 1) ? 1.0f : -1.0f;
                            float x0 = m * sg;
                            float mag = fabsf(x0), x1;
                            if (sharing[0] == 0) x1 = mag;                   /* :267-268 */
                            else {
                                float w0 = cn_weight(w_cn, sharing[0], t, wc, e, i);
                                if (sharing[1] == sharing[0]) {
                                    float w1 = cn_weight(w_ucn, sharing[1], t, wu, e, i);
                                    float u = ucn[i * z + a] ? 1.0f : 0.0f;
                                    x1 = (mag * w0) * (1.0f - u) + (mag * w1) * u;   /* :275/:285/:295 */
                                } else x1 = mag * w0;
                            }
                            float x2 = x1 * (x1 > 0.0f ? 1.0f : 0.0f);       /* :308 */
                            x2 = qms ? quant(x2, q_bit) : clipf(x2, -clip_llr, clip_llr);   /* :310-313 */
                            float sgn0 = (x0 > 0.0f) - (x0 < 0.0f);
                            c2v[e * z + (a + g.shift[e]) % z] = x2 * sgn0;   /* :316, VN lane frame */
                        }
                    }
                }
                /* D8 */
                int bad = 0;
                for (int j = 0; j < N; ++j)
                    for (int c = 0; c < z; ++c) {
                        float s = 0.0f;
                        for (int q = g.col_ptr[j]; q < g.col_ptr[j + 1]; ++q) s = s + c2v[g.col_edge[q] * z + c];
                        app[j * z + c] = clipf(xq[j * z + c] + s, -clip_llr, clip_llr);
                    }
                if (app_all) memcpy(app_all + ((size_t)t * B + b) * N * z, app, sizeof(float) * N * z);
                if (c2v_all) memcpy(c2v_all + ((size_t)t * B + b) * E * z, c2v, sizeof(float) * E * z);
                if (synd) {
                    for (int i = 0; i < M && !bad; ++i)
                        for (int a = 0; a < z && !bad; ++a) {
                            int par = 0;
                            for (int e = g.row_ptr[i]; e < g.row_ptr[i + 1]; ++e)
                                par ^= app[g.col[e] * z + (a + g.shift[e]) % z] >= 0.0f;
                            bad |= par;
                        }
                    synd[(size_t)t * B + b] = (uint8_t)bad;
                }
            }
            if (app_last) memcpy(app_last + (size_t)b * N * z, app, sizeof(float) * N * z);
        }
        }
        free(c2v); free(v2c); free(xin); free(xq); free(app); free(ucn);
    }
    graph_free(&g);
    return status;
}

int nms_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
