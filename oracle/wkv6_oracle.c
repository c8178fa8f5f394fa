/* CPU oracle (C port) of the reference WKV6 kernels -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain C restatement, fp32 arithmetic like the reference CUDA kernels, one OpenMP task per
 * (batch, head) stream exactly as the reference launches one thread block per (b, h)
 * (cuda/wkv6_cuda.cu:229-242).  Used by tests/ as a fast checker for mid-size shapes and by
 * bench.py as the `cpu_baseline` / `--impl reference` arm ("port").  The product library never
 * links or loads this file.
 *
 *   forward : cuda/wkv6_cuda.cu:7-61 (S0 = 0), cuda/wkv6state_cuda.cu:7-68 (S0 given),
 *             cuda/wkv6infctx_cuda.cu:65-67 (final state written back), cuda/rwkv6.cu:8-71
 *   backward: the quantities of kernel_backward_111 / _222 (cuda/wkv6_cuda.cu:63-227,
 *             cuda/wkv6state_cuda.cu:70-296).  gr, gu, gk, gv are the same sweeps; gw uses the
 *             identity  gl_t = sum_{s>t}(A_s - B_s) - B_t,  A_s = r_s*(S_s gy_s), B_s = k_s*(G_s v_s)
 *             (the reference computes the same sums with its sbbbb[] array).
 *
 * Parity: pinned by tests/test_oracle_c.py against oracle/wkv6_oracle.py (itself pinned against the
 * reference's CPU path through tests/golden/).
 *
 * Layouts: r,k,v,w,gy,y,g* are [B,T,C] fp32, C = H*64; u [H,64]; states [.,H,64(value),64(key)]
 * (the layout of the reference CUDA ops, cuda/wkv6state_cuda.cu:15,24).
 * w_kind: 0 = raw logits (decay = exp(-exp(w))), 1 = log decay l = -exp(w) (the fp32 `ew` of
 * src/model.py:210), 2 = decay itself (src/model_run.py:64).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define N 64

static inline float log_decay(float w, int w_kind) {
    if (w_kind == 0) return -expf(w);
    if (w_kind == 1) return w;
    return logf(w > 1e-38f ? w : 1e-38f);
}

/* n > 0: use n threads from now on (launchers such as torchrun export OMP_NUM_THREADS=1) */
void wkv6_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int wkv6_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* s0: NULL or [s0_batch ? B : 1][H][64][64] (value,key).  sT: NULL or [B][H][64][64]. */
void wkv6_oracle_forward(int B, int T, int H, const float *r, const float *k, const float *v,
                         const float *w, const float *u, const float *s0, int s0_batch,
                         float *y, float *sT, int w_kind) {
    const int C = H * N;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int b = 0; b < B; b++)
        for (int h = 0; h < H; h++) {
            float S[N][N]; /* [value j][key i], as the reference thread j holds state[i] */
            if (s0) memcpy(S, s0 + ((size_t)(s0_batch ? b : 0) * H + h) * N * N, sizeof(S));
            else memset(S, 0, sizeof(S));
            const float *uu = u + h * N;
            for (int t = 0; t < T; t++) {
                const size_t o = ((size_t)b * T + t) * C + h * N;
                float d[N];
                for (int i = 0; i < N; i++) d[i] = expf(log_decay(w[o + i], w_kind));
                for (int j = 0; j < N; j++) {
                    const float vj = v[o + j];
                    float acc = 0.f;
                    for (int i = 0; i < N; i++) {
                        const float x = k[o + i] * vj;
                        acc += r[o + i] * (uu[i] * x + S[j][i]);
                        S[j][i] = S[j][i] * d[i] + x;
                    }
                    y[o + j] = acc;
                }
            }
            if (sT) memcpy(sT + ((size_t)b * H + h) * N * N, S, sizeof(S));
        }
}

/* gu: [B][C] per-batch partials (summed by the caller like src/model.py:232);
 * gs: NULL or [B][H][64][64] (value,key) per-batch partials.
 * zero_gw0: write an exact 0 at t = 0 (plain wkv6, cuda/wkv6_cuda.cu:201). */
void wkv6_oracle_backward(int B, int T, int H, const float *r, const float *k, const float *v,
                          const float *w, const float *u, const float *s0, int s0_batch,
                          const float *gy, float *gr, float *gk, float *gv, float *gw, float *gu,
                          float *gs, int w_kind, int zero_gw0) {
    const int C = H * N;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int b = 0; b < B; b++)
        for (int h = 0; h < H; h++) {
            float S[N][N];  /* [key i][value j] here */
            float G[N][N];  /* dL/dS_{t+1}, [key i][value j] */
            float *A = (float *)malloc(sizeof(float) * (size_t)T * N);
            const float *uu = u + h * N;
            if (s0) {
                const float *sp = s0 + ((size_t)(s0_batch ? b : 0) * H + h) * N * N;
                for (int j = 0; j < N; j++)
                    for (int i = 0; i < N; i++) S[i][j] = sp[j * N + i];
            } else memset(S, 0, sizeof(S));
            float guacc[N];
            memset(guacc, 0, sizeof(guacc));
            /* forward sweep: gr, gu, A */
            for (int t = 0; t < T; t++) {
                const size_t o = ((size_t)b * T + t) * C + h * N;
                float vg = 0.f;
                for (int j = 0; j < N; j++) vg += v[o + j] * gy[o + j];
                for (int i = 0; i < N; i++) {
                    float sg = 0.f;
                    for (int j = 0; j < N; j++) sg += S[i][j] * gy[o + j];
                    gr[o + i] = uu[i] * k[o + i] * vg + sg;
                    A[(size_t)t * N + i] = r[o + i] * sg;
                    guacc[i] += r[o + i] * k[o + i] * vg;
                    const float d = expf(log_decay(w[o + i], w_kind));
                    for (int j = 0; j < N; j++) S[i][j] = S[i][j] * d + k[o + i] * v[o + j];
                }
            }
            for (int i = 0; i < N; i++) gu[(size_t)b * C + h * N + i] = guacc[i];
            /* reverse sweep: gk, gv, gw, gs */
            memset(G, 0, sizeof(G));
            float q[N];
            memset(q, 0, sizeof(q));
            for (int t = T - 1; t >= 0; t--) {
                const size_t o = ((size_t)b * T + t) * C + h * N;
                float gvacc[N];
                memset(gvacc, 0, sizeof(gvacc));
                for (int i = 0; i < N; i++) {
                    const float ur = uu[i] * r[o + i];
                    float gv_dot = 0.f, ugy = 0.f;
                    for (int j = 0; j < N; j++) {
                        gv_dot += G[i][j] * v[o + j];
                        ugy += gy[o + j] * v[o + j];
                        gvacc[j] += k[o + i] * (ur * gy[o + j] + G[i][j]);
                    }
                    gk[o + i] = ur * ugy + gv_dot;
                    const float Bt = k[o + i] * gv_dot;
                    const float l = log_decay(w[o + i], w_kind);
                    const float gl = q[i] - Bt;
                    /* chain rule to the tensor the op differentiates: raw w -> l*gl; for w_kind 1
                       the reference still returns d/d(raw w) = ew * gl (cuda/wkv6_cuda.cu:202,225) */
                    gw[o + i] = (t == T - 1 || (t == 0 && zero_gw0)) ? 0.f : l * gl;
                    q[i] += A[(size_t)t * N + i] - Bt;
                    const float d = expf(l);
                    for (int j = 0; j < N; j++) G[i][j] = r[o + i] * gy[o + j] + d * G[i][j];
                }
                for (int j = 0; j < N; j++) gv[o + j] = gvacc[j];
            }
            if (gs) {
                float *gp = gs + ((size_t)b * H + h) * N * N;
                for (int j = 0; j < N; j++)
                    for (int i = 0; i < N; i++) gp[j * N + i] = G[i][j];
            }
            free(A);
        }
}
