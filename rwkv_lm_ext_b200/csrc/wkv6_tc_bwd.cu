// WKV6 backward, chunked, on tcgen05 tensor cores fed by TMA (reverse sweep over 64-token chunks).
//
// Notation as in wkv6_tc_fwd.cu.  Inputs per chunk: r,k,v,w,gy tiles, S_in = state at the chunk
// start (bf16 checkpoint written by the forward pre-pass, layout [value j][key i]), and the running
// reverse state G = dL/dS at the chunk end, kept in TMEM as fp32 [key i][value j].
//
//   Bm[t,s]   = gy_t . v_s                         (64^3)   -> dA = strict-lower(Bm), bd[t] = Bm[t,t]
//   Bm^T[s,t]                                      (64^3)   -> dA^T
//   A^T[s,t]  (as in the forward)                  4 x N16  -> P^T = strict-upper + diag(sum_i r u k)
//   gv[s,j]   = sum_t P^T[s,t] gy_t[j] + sum_i Kh[s,i] G[i,j]                     2 x 64^3
//   Xr[i,t]   = sum_{s<t} Kt_q[s,i] dA[t,s]   + e^{rho_q} sum_j S_in[i,j] gy_t[j]   (q = block of t)
//   Xk[i,s]   = sum_{t>s} Rp_p[t,i] dA^T[s,t] + e^{Lam-rho_p} sum_j G[i,j] v_s[j]   (p = block of s)
//   gr_t[i]   = E[t,i] Xr[i,t] + u_i k_t[i] bd[t],    E = exp(exc_t - rho_q)
//   gk_s[i]   = F[s,i] Xk[i,s] + u_i r_s[i] bd[s],    F = exp(rho_p - cum_s)
//   gl_t[i]   = e^{Lam_i} <S_in, G>_i + sum_{s<t} Be_s + sum_{t'>t} Ae_t' + sum_{s>t} (Ai_s - Bi_s) - Bi_t
//               Ae = r E e^{rho} (S_in gy),  Be = k F e^{Lam-rho} (G v)      (terms through the states)
//               Ai = Rt_own * Dr,  Bi = Kt_own * Dk                          (intra-chunk pair terms)
//               gw = l * gl.  This is d_t <S_t, G_t>_i expanded so that no two
//               large quantities are subtracted: the references rho_q sit on the INTEGER log2 grid, so
//               all Kt_q / Rp_p versions are exact power-of-two multiples of one another and Ai, Bi are
//               sums of bit-identical pair products r k dA (their difference telescopes exactly, as in
//               the reference's fp32 suffix trick, cuda/wkv6_cuda.cu:161-227).
//   G[i,j]    = e^{Lam_i} G[i,j] + sum_t Rh[t,i] gy_t[j]                           (64^3, in TMEM)
//
// Xr / Xk come out of the tensor cores with the KEY CHANNEL on the TMEM lanes, so the boundary term
// <S_in,G> and the decay of G are per-thread; Dr, Drs, Dk, Dks are handed to the operand-preparation
// threads (which own channel pairs, hold the decay prefix sums and produce coalesced 128-byte output
// rows) through four fp32 staging tiles that reuse the operand tiles once the MMAs are done.
//
// Warp roles (544 threads, 1 CTA per SM): warps 0-7 TMEM side (sub-partition = warp%4, column half =
// warp/4), warps 8-15 operand preparation + output stage, warp 16 TMA + tcgen05.mma issue.
// The kernel handles the non-hazard route only: the forward pre-pass raises a device flag when any
// chunk needed the exact route, this kernel then returns at once and the SIMT backward (enqueued
// right behind it, predicated on the same flag) does the work -- no host round trip.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace wkv6 {
namespace {

using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int L = 64;
constexpr int TM_THREADS = 256, PREP_THREADS = 256, NTHREADS = TM_THREADS + PREP_THREADS + 32;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

// shared memory map
constexpr uint32_t RAW_SET = 40960, RAW_R = 0, RAW_K = 8192, RAW_W = 16384, RAW_V = 24576, RAW_GY = 32768;
constexpr uint32_t OFF_SIN = 81920;                          // 2 x 8 KB
constexpr uint32_t OFF_GB = 98304, OFF_KT = 106496, OFF_RP = 126976;
constexpr uint32_t OFF_RH = 147456, OFF_KH = 155648, OFF_DA = 163840, OFF_DAT = 172032, OFF_PT = 180224;
constexpr uint32_t OFF_TILES_END = 188416;
// fp32 staging [t][i] (256 B rows), valid between the end of the chunk's MMAs and the next operand preparation
constexpr uint32_t ST_DR = OFF_KT, ST_DRS = ST_DR + 16384, ST_DK = ST_DRS + 16384, ST_DKS = ST_DK + 16384;
static_assert(ST_DKS + 16384 <= 172032, "staging must stay inside the dead operand tiles (KT..DA)");
__host__ __device__ constexpr uint32_t kt_off(int q) { return q == 0 ? 18432u : q == 1 ? 14336u : q == 2 ? 8192u : 0u; }
__host__ __device__ constexpr uint32_t rp_off(int p) { return p == 0 ? 0u : p == 1 ? 8192u : p == 2 ? 14336u : 18432u; }

struct Extra {
    alignas(16) float elam[64];
    float diagu[64];
    float bd[64];
    float qend[64];
    float erho[4][64];        // exp(rho_q)
    float elr[4][64];         // exp(Lam - rho_p)
    float2 htot[8][32];
    float4 ftot[8][32];       // per half-block: (sum X) x2, (sum Be) x2
    float gup[8][64];
    uint64_t bar_raw[2], bar_sin[2], bar_prep, bar_p1, bar_e1, bar_p2, bar_e2;
    uint32_t tmem_base;
};
constexpr uint32_t SMEM_BYTES = OFF_TILES_END + sizeof(Extra);
constexpr uint32_t TM_BM = 0, TM_BMT = 64, TM_AT = 128, TM_DRS = 192, TM_DKS = 256, TM_G = 320;
constexpr uint32_t TM_GV = 0, TM_DR = 64, TM_DK = 128, TM_COLS = 512;

struct Params {
    int B, T, H;
    const bf16 *u;
    int has_s0;
    bf16 *gr, *gk, *gv, *gw, *gu, *gs;
    const int *hz_flag;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ float bf_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^n for an integer-valued n <= 0, exact (0 below the normal range)
__device__ __forceinline__ float pow2i(float n) {
    const int e = max((int)n, -127);
    return __int_as_float((e + 127) << 23);
}
__device__ __forceinline__ uint4 pack8f(const float *f) {
    uint4 o;
    o.x = pack2(f[0], f[1]); o.y = pack2(f[2], f[3]); o.z = pack2(f[4], f[5]); o.w = pack2(f[6], f[7]);
    return o;
}

__global__ void __launch_bounds__(NTHREADS, 1)
wkv6_tc_bwd_kernel(const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_ck, Params p) {
    if (p.hz_flag && *p.hz_flag != 0) return;       // the exact (SIMT) route handles this call
    extern __shared__ __align__(1024) uint8_t sm[];
    Extra &ex = *reinterpret_cast<Extra *>(sm + OFF_TILES_END);
    if ((smem_u32(sm) & 1023u) != 0) __trap();
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int T = p.T, C = p.H * 64;
    const int NC = (T + L - 1) / L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&ex.bar_raw[0], 1);
        mbar_init(&ex.bar_raw[1], 1);
        mbar_init(&ex.bar_sin[0], 1);
        mbar_init(&ex.bar_sin[1], 1);
        mbar_init(&ex.bar_prep, PREP_THREADS);
        mbar_init(&ex.bar_p1, 1);
        mbar_init(&ex.bar_e1, TM_THREADS);
        mbar_init(&ex.bar_p2, 1);
        mbar_init(&ex.bar_e2, TM_THREADS);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&ex.tmem_base, TM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = ex.tmem_base;

    if (warp == 16) {
        // =====================================================================================
        // issuer warp
        // =====================================================================================
        auto issue_raw = [&](int c, int set) {
            uint8_t *base = sm + set * RAW_SET;
            mbar_arrive_expect_tx(&ex.bar_raw[set], 5 * 8192);
            tma_load_3d(base + RAW_R, &map_r, &ex.bar_raw[set], h * 64, c * L, b);
            tma_load_3d(base + RAW_K, &map_k, &ex.bar_raw[set], h * 64, c * L, b);
            tma_load_3d(base + RAW_W, &map_w, &ex.bar_raw[set], h * 64, c * L, b);
            tma_load_3d(base + RAW_V, &map_v, &ex.bar_raw[set], h * 64, c * L, b);
            tma_load_3d(base + RAW_GY, &map_gy, &ex.bar_raw[set], h * 64, c * L, b);
        };
        auto issue_sin = [&](int c, int buf) {
            mbar_arrive_expect_tx(&ex.bar_sin[buf], 8192);
            tma_load_3d(sm + OFF_SIN + buf * 8192, &map_ck, &ex.bar_sin[buf], 0, (blockIdx.x * NC + c) * 64, 0);
        };
        if (lane == 0) {
            tma_prefetch_desc(&map_r); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
            tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_gy); tma_prefetch_desc(&map_ck);
            issue_raw(NC - 1, 0);
            issue_sin(NC - 1, 0);
        }
        const uint32_t kt = smem_u32(sm + OFF_KT), rp = smem_u32(sm + OFF_RP), rh = smem_u32(sm + OFF_RH);
        const uint32_t kh = smem_u32(sm + OFF_KH), da = smem_u32(sm + OFF_DA), dat = smem_u32(sm + OFF_DAT);
        const uint32_t pt = smem_u32(sm + OFF_PT), gb = smem_u32(sm + OFF_GB);
        constexpr uint32_t ID16_KK = idesc_bf16(64, 16, 0, 0), ID16_MK = idesc_bf16(64, 16, 1, 0);
        constexpr uint32_t ID_KK = idesc_bf16(64, 64, 0, 0), ID_KM = idesc_bf16(64, 64, 0, 1);
        constexpr uint32_t ID_MK = idesc_bf16(64, 64, 1, 0), ID_MM = idesc_bf16(64, 64, 1, 1);
        for (int it = 0; it < NC; it++) {
            const int c = NC - 1 - it;
            const uint32_t par = it & 1, set = it & 1;
            mbar_wait(&ex.bar_raw[set], (it >> 1) & 1);
            if (lane == 0 && c > 0) {
                issue_raw(c - 1, set ^ 1);
                issue_sin(c - 1, set ^ 1);
            }
            named_bar_sync<1, PREP_THREADS + 32>();
            mbar_wait(&ex.bar_prep, par);
            mbar_wait(&ex.bar_sin[set], (it >> 1) & 1);
            mbar_wait(&ex.bar_e2, par);                       // TMEM free, bf16 copy of G ready
            tc_fence_after();
            const uint32_t raw = smem_u32(sm + set * RAW_SET);
            const uint32_t vv = raw + RAW_V, gy = raw + RAW_GY, sin = smem_u32(sm + OFF_SIN + set * 8192);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; k++)   // Bm[t,s] = GY V^T
                    mma_bf16_ss(tmem + TM_BM, smem_desc_sw128(gy + 32 * k, 8192, 1024), smem_desc_sw128(vv + 32 * k, 8192, 1024), ID_KK, k > 0);
#pragma unroll
                for (int k = 0; k < 4; k++)   // Bm^T[s,t] = V GY^T
                    mma_bf16_ss(tmem + TM_BMT, smem_desc_sw128(vv + 32 * k, 8192, 1024), smem_desc_sw128(gy + 32 * k, 8192, 1024), ID_KK, k > 0);
#pragma unroll
                for (int q = 0; q < 4; q++)   // A^T blocks
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        mma_bf16_ss(tmem + TM_AT + 16 * q, smem_desc_sw128(kt + kt_off(q) + 32 * k, 8192, 1024),
                                    smem_desc_sw128(rp + rp_off(q) + 32 * k, 8192, 1024), ID16_KK, k > 0);
#pragma unroll
                for (int k = 0; k < 4; k++)   // Drs[i,t] = sum_j S_in[i,j] gy_t[j]   (S_in tile is [j][i]: MN-major A)
                    mma_bf16_ss(tmem + TM_DRS, smem_desc_sw128(sin + 2048 * k, 8192, 1024), smem_desc_sw128(gy + 32 * k, 8192, 1024), ID_MK, k > 0);
#pragma unroll
                for (int k = 0; k < 4; k++)   // Dks[i,s] = sum_j G[i,j] v_s[j]
                    mma_bf16_ss(tmem + TM_DKS, smem_desc_sw128(gb + 32 * k, 8192, 1024), smem_desc_sw128(vv + 32 * k, 8192, 1024), ID_KK, k > 0);
                mma_commit(&ex.bar_p1);
            }
            __syncwarp();
            mbar_wait(&ex.bar_e1, par);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; k++)   // gv[s,j] = P^T GY ...
                    mma_bf16_ss(tmem + TM_GV, smem_desc_sw128(pt + 32 * k, 8192, 1024), smem_desc_sw128(gy + 2048 * k, 8192, 1024), ID_KM, k > 0);
#pragma unroll
                for (int k = 0; k < 4; k++)   // ... + Kh G
                    mma_bf16_ss(tmem + TM_GV, smem_desc_sw128(kh + 32 * k, 8192, 1024), smem_desc_sw128(gb + 2048 * k, 8192, 1024), ID_KM, 1);
#pragma unroll
                for (int q = 0; q < 4; q++)   // Dr[i, t in q] = sum_{s in blocks <= q} Kt_q[s,i] dA[t,s]
#pragma unroll
                    for (int ks = 0; ks < 4; ks++)
                        if (ks <= q)
                            mma_bf16_ss(tmem + TM_DR + 16 * q, smem_desc_sw128(kt + kt_off(q) + 2048 * ks, 8192, 1024),
                                        smem_desc_sw128(da + 2048 * q + 32 * ks, 8192, 1024), ID16_MK, ks > 0);
#pragma unroll
                for (int pb = 0; pb < 4; pb++)   // Dk[i, s in p] = sum_{t in blocks >= p} Rp_p[t,i] dA^T[s,t]
#pragma unroll
                    for (int ks = 0; ks < 4; ks++)
                        if (ks < 4 - pb)
                            mma_bf16_ss(tmem + TM_DK + 16 * pb, smem_desc_sw128(rp + rp_off(pb) + 2048 * ks, 8192, 1024),
                                        smem_desc_sw128(dat + 2048 * pb + 32 * (pb + ks), 8192, 1024), ID16_MK, ks > 0);
#pragma unroll
                for (int k = 0; k < 4; k++)   // G[i,j] += sum_t Rh[t,i] gy_t[j]
                    mma_bf16_ss(tmem + TM_G, smem_desc_sw128(rh + 2048 * k, 8192, 1024), smem_desc_sw128(gy + 2048 * k, 8192, 1024), ID_MM, 1);
                mma_commit(&ex.bar_p2);
            }
            __syncwarp();
            mbar_wait(&ex.bar_p2, par);
        }
    } else if (warp >= 8) {
        // =====================================================================================
        // operand preparation + output stage: warp = half-block hb (8 token rows), lane = channel pair
        // =====================================================================================
        const int hb = warp - 8, q = hb >> 1;
        const uint32_t boff = ((uint32_t)(lane >> 2) << 4) | ((uint32_t)(lane & 3) << 2);
        const uint32_t rowbase = 1024u * hb;
        const float u0 = __bfloat162float(p.u[h * 64 + 2 * lane]), u1 = __bfloat162float(p.u[h * 64 + 2 * lane + 1]);
        float gu0 = 0.f, gu1 = 0.f;

        for (int it = 0; it < NC; it++) {
            const int c = NC - 1 - it;
            const uint32_t par = it & 1;
            const int nv = min(L, T - c * L);
            const uint8_t *raw = sm + (it & 1) * RAW_SET;
            named_bar_sync<1, PREP_THREADS + 32>();

            // ---- phase A
            float l0[8], l1[8];
            float hs0 = 0.f, hs1 = 0.f;
#pragma unroll
            for (int n = 0; n < 8; n++) {
                const uint32_t ww = *reinterpret_cast<const uint32_t *>(raw + RAW_W + rowbase + n * 128 + (boff ^ (n << 4)));
                const bool valid = (8 * hb + n) < nv;
                l0[n] = valid ? -fast_ex2(bf_lo(ww) * LOG2E) * LOG2E : 0.f;
                l1[n] = valid ? -fast_ex2(bf_hi(ww) * LOG2E) * LOG2E : 0.f;
                hs0 += l0[n];
                hs1 += l1[n];
            }
            ex.htot[hb][lane] = make_float2(hs0, hs1);
            if (hb == 0) *reinterpret_cast<float2 *>(&ex.qend[2 * lane]) = make_float2(0.f, 0.f);
            named_bar_sync<2, PREP_THREADS>();

            // ---- phase B
            float run0 = 0.f, run1 = 0.f, b0 = 0.f, b1 = 0.f, rho0[4], rho1[4];
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const float2 hv = ex.htot[x][lane];
                if (x == hb) { b0 = run0; b1 = run1; }
                if (x & 1) { rho0[x >> 1] = rintf(run0); rho1[x >> 1] = rintf(run1); }   // block middle, integer log2 grid
                run0 += hv.x;
                run1 += hv.y;
            }
            const float lam0 = run0, lam1 = run1;
            const float rq0 = q == 0 ? rho0[0] : q == 1 ? rho0[1] : q == 2 ? rho0[2] : rho0[3];
            const float rq1 = q == 0 ? rho1[0] : q == 1 ? rho1[1] : q == 2 ? rho1[2] : rho1[3];
            if (hb == 0) {
                *reinterpret_cast<float2 *>(&ex.elam[2 * lane]) = make_float2(fast_ex2(lam0), fast_ex2(lam1));
#pragma unroll
                for (int qq = 0; qq < 4; qq++) {
                    *reinterpret_cast<float2 *>(&ex.erho[qq][2 * lane]) = make_float2(fast_ex2(rho0[qq]), fast_ex2(rho1[qq]));
                    *reinterpret_cast<float2 *>(&ex.elr[qq][2 * lane]) = make_float2(fast_ex2(lam0 - rho0[qq]), fast_ex2(lam1 - rho1[qq]));
                }
            }
            const float er0 = fast_ex2(rq0), er1 = fast_ex2(rq1);
            const float el0 = fast_ex2(lam0 - rq0), el1 = fast_ex2(lam1 - rq1);
            float g0[4], g1[4], f0[4], f1[4];
#pragma unroll
            for (int qq = 0; qq < 4; qq++) {
                g0[qq] = pow2i(fminf(rho0[qq] - rq0, 0.f));   // Kt_{q'} = (k F) * 2^(rho_q' - rho_q),  q' > q  (exact)
                g1[qq] = pow2i(fminf(rho1[qq] - rq1, 0.f));
                f0[qq] = pow2i(fminf(rq0 - rho0[qq], 0.f));   // Rp_{p}  = (r E) * 2^(rho_q - rho_p),   p < q  (exact)
                f1[qq] = pow2i(fminf(rq1 - rho1[qq], 0.f));
            }
            uint32_t rr[8], kk[8], rto[8], kto[8];
            float cum0 = b0, cum1 = b1;
            float du[8];
#pragma unroll
            for (int n = 0; n < 8; n++) {
                const uint32_t off = rowbase + n * 128 + (boff ^ (n << 4));
                rr[n] = *reinterpret_cast<const uint32_t *>(raw + RAW_R + off);
                kk[n] = *reinterpret_cast<const uint32_t *>(raw + RAW_K + off);
                const float r0 = bf_lo(rr[n]), r1 = bf_hi(rr[n]), k0 = bf_lo(kk[n]), k1 = bf_hi(kk[n]);
                const float exc0 = cum0, exc1 = cum1;
                cum0 += l0[n];
                cum1 += l1[n];
                const float rt0 = r0 * fast_ex2(exc0 - rq0), rt1 = r1 * fast_ex2(exc1 - rq1);
                const float kf0 = k0 * fast_ex2(rq0 - cum0), kf1 = k1 * fast_ex2(rq1 - cum1);
                rto[n] = pack2(rt0, rt1);
                kto[n] = pack2(kf0, kf1);
                // Rp version p holds rows t >= 16p at local row (t - 16p): offset shrinks by 2048 per version
#pragma unroll
                for (int pp = 0; pp < 4; pp++) {
                    if (pp == q) *reinterpret_cast<uint32_t *>(sm + OFF_RP + rp_off(pp) + off - 2048u * pp) = rto[n];
                    else if (pp < q) *reinterpret_cast<uint32_t *>(sm + OFF_RP + rp_off(pp) + off - 2048u * pp) = pack2(rt0 * f0[pp], rt1 * f1[pp]);
                }
                *reinterpret_cast<uint32_t *>(sm + OFF_RH + off) = pack2(rt0 * er0, rt1 * er1);
                *reinterpret_cast<uint32_t *>(sm + OFF_KT + kt_off(q) + off) = kto[n];
#pragma unroll
                for (int qq = 1; qq < 4; qq++)
                    if (qq > q) *reinterpret_cast<uint32_t *>(sm + OFF_KT + kt_off(qq) + off) = pack2(kf0 * g0[qq], kf1 * g1[qq]);
                *reinterpret_cast<uint32_t *>(sm + OFF_KH + off) = pack2(kf0 * el0, kf1 * el1);
                du[n] = r0 * u0 * k0 + r1 * u1 * k1;
            }
            {
                const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
                float a4[4], a2[2], a1;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float keep = h16 ? du[4 + j] : du[j], send = h16 ? du[j] : du[4 + j];
                    a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const float keep = h8 ? a4[2 + j] : a4[j], send = h8 ? a4[j] : a4[2 + j];
                    a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                {
                    const float keep = h4 ? a2[1] : a2[0], send = h4 ? a2[0] : a2[1];
                    a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
                a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
                if ((lane & 3) == 0) ex.diagu[8 * hb + (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0)] = a1;
            }
            fence_proxy_async();
            mbar_arrive(&ex.bar_prep);

            // ---- output stage: wait for Xr / Xk staging, bd, qend
            if (hb == 0) mbar_wait(&ex.bar_e2, par ^ 1);        // completion #(it+1)
            named_bar_sync<2, PREP_THREADS>();
            float t10[8], t11[8], xs0[8], xs1[8];
            float tx0 = 0.f, tx1 = 0.f, py0 = 0.f, py1 = 0.f;
            cum0 = b0;
            cum1 = b1;
            const size_t gbase = ((size_t)b * T + (size_t)c * L + 8 * hb) * C + (size_t)h * 64 + 2 * lane;
#pragma unroll
            for (int n = 0; n < 8; n++) {
                const uint32_t so = (uint32_t)(8 * hb + n) * 256u + 8u * lane;
                const float2 dr = *reinterpret_cast<const float2 *>(sm + ST_DR + so);
                const float2 drs = *reinterpret_cast<const float2 *>(sm + ST_DRS + so);
                const float2 dk = *reinterpret_cast<const float2 *>(sm + ST_DK + so);
                const float2 dks = *reinterpret_cast<const float2 *>(sm + ST_DKS + so);
                const float bdt = ex.bd[8 * hb + n];
                const float r0 = bf_lo(rr[n]), r1 = bf_hi(rr[n]), k0 = bf_lo(kk[n]), k1 = bf_hi(kk[n]);
                const float exc0 = cum0, exc1 = cum1;
                cum0 += l0[n];
                cum1 += l1[n];
                const float e0 = fast_ex2(exc0 - rq0), e1 = fast_ex2(exc1 - rq1);
                const float f0_ = fast_ex2(rq0 - cum0), f1_ = fast_ex2(rq1 - cum1);
                const float bi0 = bf_lo(kto[n]) * dk.x, bi1 = bf_hi(kto[n]) * dk.y;
                const float be0 = k0 * f0_ * dks.x, be1 = k1 * f1_ * dks.y;
                const float x0 = fmaf(r0 * e0, drs.x, fmaf(bf_lo(rto[n]), dr.x, -bi0));
                const float x1 = fmaf(r1 * e1, drs.y, fmaf(bf_hi(rto[n]), dr.y, -bi1));
                t10[n] = py0 - bi0;
                t11[n] = py1 - bi1;
                py0 += be0;
                py1 += be1;
                xs0[n] = x0;
                xs1[n] = x1;
                tx0 += x0;
                tx1 += x1;
                gu0 += r0 * k0 * bdt;
                gu1 += r1 * k1 * bdt;
                if (8 * hb + n < nv) {
                    *reinterpret_cast<uint32_t *>(p.gr + gbase + (size_t)n * C) =
                        pack2(fmaf(e0, dr.x + drs.x, u0 * k0 * bdt), fmaf(e1, dr.y + drs.y, u1 * k1 * bdt));
                    *reinterpret_cast<uint32_t *>(p.gk + gbase + (size_t)n * C) =
                        pack2(fmaf(f0_, dk.x + dks.x, u0 * r0 * bdt), fmaf(f1_, dk.y + dks.y, u1 * r1 * bdt));
                }
            }
            ex.ftot[hb][lane] = make_float4(tx0, tx1, py0, py1);
            named_bar_sync<2, PREP_THREADS>();
            float q0 = ex.qend[2 * lane], q1 = ex.qend[2 * lane + 1];
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const float4 fv = ex.ftot[x][lane];
                if (x < hb) { q0 += fv.z; q1 += fv.w; }
                if (x > hb) { q0 += fv.x; q1 += fv.y; }
            }
            float sx0 = 0.f, sx1 = 0.f;
#pragma unroll
            for (int n = 7; n >= 0; n--) {
                float gw0 = l0[n] * LN2 * (q0 + t10[n] + sx0), gw1 = l1[n] * LN2 * (q1 + t11[n] + sx1);
                sx0 += xs0[n];
                sx1 += xs1[n];
                if (c == 0 && hb == 0 && n == 0 && !p.has_s0) { gw0 = 0.f; gw1 = 0.f; }
                if (8 * hb + n < nv) *reinterpret_cast<uint32_t *>(p.gw + gbase + (size_t)n * C) = pack2(gw0, gw1);
            }
        }
        // gu partials of this (b,h): sum over the 8 half-block warps
        ex.gup[hb][2 * lane] = gu0;
        ex.gup[hb][2 * lane + 1] = gu1;
        named_bar_sync<2, PREP_THREADS>();
        if (hb == 0) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int x = 0; x < 8; x++) { s0 += ex.gup[x][2 * lane]; s1 += ex.gup[x][2 * lane + 1]; }
            *reinterpret_cast<uint32_t *>(p.gu + (size_t)b * C + h * 64 + 2 * lane) = pack2(s0, s1);
        }
    } else {
        // =====================================================================================
        // TMEM side: warp w -> sub-partition w%4 (rows 16*(w%4)..+15 in lanes 0-15), column half w/4
        // =====================================================================================
        const int sub = warp & 3, hh = warp >> 2;
        const int row = 16 * sub + (lane & 15);
        const bool act = lane < 16;
        const uint32_t tlane = 32 * sub;
        uint32_t v[32], v2[32];
        float f[32];

        // G = 0, bf16 copy = 0
#pragma unroll
        for (int cc = 0; cc < 32; cc++) v[cc] = 0u;
        tmem_st32(tmem_addr(tmem, tlane, TM_G + 32 * hh), v);
        if (act) {
#pragma unroll
            for (int ch = 0; ch < 4; ch++) *reinterpret_cast<uint4 *>(sm + OFF_GB + sw128(row, 64 * hh + 16 * ch)) = make_uint4(0, 0, 0, 0);
        }
        tmem_wait_st();
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&ex.bar_e2);

        for (int it = 0; it < NC; it++) {
            const int c = NC - 1 - it;
            const uint32_t par = it & 1;
            const int nv = min(L, T - c * L);
            if (warp == 0) {
                mbar_wait(&ex.bar_prep, par);
                mbar_wait(&ex.bar_p1, par);
            }
            named_bar_sync<3, TM_THREADS>();
            tc_fence_after();

            // ---- dA[t][s] = Bm[t,s] for s < t (row t = `row`), bd[t] = Bm[t,t]
            tmem_ld32(tmem_addr(tmem, tlane, TM_BM + 32 * hh), v);
            tmem_wait_ld();
            if (act) {
                float bdv = 0.f;
#pragma unroll
                for (int cc = 0; cc < 32; cc++) {
                    const int s = 32 * hh + cc;
                    const float x = __uint_as_float(v[cc]);
                    f[cc] = (s < row) ? x : 0.f;
                    if (s == row) bdv = x;
                }
                if ((row >> 5) == hh) ex.bd[row] = bdv;
#pragma unroll
                for (int ch = 0; ch < 4; ch++) *reinterpret_cast<uint4 *>(sm + OFF_DA + sw128(row, 64 * hh + 16 * ch)) = pack8f(f + 8 * ch);
            }
            // ---- dA^T[s][t] = Bm^T[s,t] for t > s (row s = `row`)
            tmem_ld32(tmem_addr(tmem, tlane, TM_BMT + 32 * hh), v);
            tmem_wait_ld();
            if (act) {
#pragma unroll
                for (int cc = 0; cc < 32; cc++) f[cc] = (32 * hh + cc > row) ? __uint_as_float(v[cc]) : 0.f;
#pragma unroll
                for (int ch = 0; ch < 4; ch++) *reinterpret_cast<uint4 *>(sm + OFF_DAT + sw128(row, 64 * hh + 16 * ch)) = pack8f(f + 8 * ch);
            }
            // ---- P^T[s][t] = A^T[s,t] for t > s, diag = sum_i r u k
            tmem_ld32(tmem_addr(tmem, tlane, TM_AT + 32 * hh), v);
            tmem_wait_ld();
            if (act) {
                const float dgu = ex.diagu[row];
#pragma unroll
                for (int cc = 0; cc < 32; cc++) {
                    const int t = 32 * hh + cc;
                    f[cc] = (t > row) ? __uint_as_float(v[cc]) : (t == row ? dgu : 0.f);
                }
#pragma unroll
                for (int ch = 0; ch < 4; ch++) *reinterpret_cast<uint4 *>(sm + OFF_PT + sw128(row, 64 * hh + 16 * ch)) = pack8f(f + 8 * ch);
            }
            // ---- G row (key channel i = `row`): boundary term e^{Lam_i} <S_in, G>_i, then decay by exp(Lam_i)
            tmem_ld32(tmem_addr(tmem, tlane, TM_G + 32 * hh), v);
            tmem_wait_ld();
            {
                const float el = ex.elam[row];
                if (act && it > 0) {
                    const uint8_t *send = sm + OFF_SIN + (it & 1) * 8192;   // S_in: checkpoint of this chunk, [j][i]
                    float qs = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 32; cc++) {
                        const int j = 32 * hh + cc;
                        const float sv = __bfloat162float(*reinterpret_cast<const bf16 *>(send + sw128(j, 2 * row)));
                        qs = fmaf(sv, __uint_as_float(v[cc]), qs);
                    }
                    atomicAdd(&ex.qend[row], qs * el);
                }
#pragma unroll
                for (int cc = 0; cc < 32; cc++) v[cc] = __float_as_uint(__uint_as_float(v[cc]) * el);
                tmem_st32(tmem_addr(tmem, tlane, TM_G + 32 * hh), v);
            }
            tmem_wait_st();
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&ex.bar_e1);

            if (warp == 0) mbar_wait(&ex.bar_p2, par);
            named_bar_sync<3, TM_THREADS>();
            tc_fence_after();
            // ---- gv rows (row s = `row`)
            tmem_ld32(tmem_addr(tmem, tlane, TM_GV + 32 * hh), v);
            tmem_wait_ld();
            if (act && row < nv) {
                bf16 *dst = p.gv + ((size_t)b * T + (size_t)c * L + row) * C + (size_t)h * 64 + 32 * hh;
#pragma unroll
                for (int cc = 0; cc < 32; cc++) f[cc] = __uint_as_float(v[cc]);
#pragma unroll
                for (int ch = 0; ch < 4; ch++) *reinterpret_cast<uint4 *>(dst + 8 * ch) = pack8f(f + 8 * ch);
            }
            // ---- Dr, e^{rho_q(t)} Drs, Dk, e^{Lam-rho_p(s)} Dks -> fp32 staging [t][i]   (key channel i = `row`)
            tmem_ld32(tmem_addr(tmem, tlane, TM_DR + 32 * hh), v);
            tmem_ld32(tmem_addr(tmem, tlane, TM_DRS + 32 * hh), v2);
            tmem_wait_ld();
            if (act) {
                const float ea = ex.erho[2 * hh][row], eb = ex.erho[2 * hh + 1][row];
                uint8_t *d0 = sm + ST_DR + (uint32_t)(32 * hh) * 256u + 4u * row;
#pragma unroll
                for (int cc = 0; cc < 32; cc++) {
                    *reinterpret_cast<uint32_t *>(d0 + cc * 256) = v[cc];
                    *reinterpret_cast<float *>(d0 + (ST_DRS - ST_DR) + cc * 256) = (cc < 16 ? ea : eb) * __uint_as_float(v2[cc]);
                }
            }
            tmem_ld32(tmem_addr(tmem, tlane, TM_DK + 32 * hh), v);
            tmem_ld32(tmem_addr(tmem, tlane, TM_DKS + 32 * hh), v2);
            tmem_wait_ld();
            if (act) {
                const float ea = ex.elr[2 * hh][row], eb = ex.elr[2 * hh + 1][row];
                uint8_t *d0 = sm + ST_DK + (uint32_t)(32 * hh) * 256u + 4u * row;
#pragma unroll
                for (int cc = 0; cc < 32; cc++) {
                    *reinterpret_cast<uint32_t *>(d0 + cc * 256) = v[cc];
                    *reinterpret_cast<float *>(d0 + (ST_DKS - ST_DK) + cc * 256) = (cc < 16 ? ea : eb) * __uint_as_float(v2[cc]);
                }
            }
            // ---- new G -> bf16 operand copy [i][j]; after the first chunk it is dL/dS_0
            tmem_ld32(tmem_addr(tmem, tlane, TM_G + 32 * hh), v);
            tmem_wait_ld();
            if (act) {
#pragma unroll
                for (int cc = 0; cc < 32; cc++) f[cc] = __uint_as_float(v[cc]);
#pragma unroll
                for (int ch = 0; ch < 4; ch++) *reinterpret_cast<uint4 *>(sm + OFF_GB + sw128(row, 64 * hh + 16 * ch)) = pack8f(f + 8 * ch);
                if (c == 0 && p.gs) {
#pragma unroll
                    for (int cc = 0; cc < 32; cc++)   // gs[b,h,j,i] = dL/dS_0[i][j]
                        p.gs[(((size_t)b * p.H + h) * 64 + 32 * hh + cc) * 64 + row] = __float2bfloat16_rn(f[cc]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&ex.bar_e2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace

size_t tc_saved_bytes(int B, int T, int H) {
    const size_t NC = (size_t)(T + L - 1) / L;
    return SAVED_HEADER + (size_t)B * H * NC * 8192;
}

size_t tc_backward_workspace_bytes(int B, int T, int H) {
    const size_t NC = (size_t)(T + L - 1) / L;
    return simt_backward_workspace_bytes(B, T, H) + (size_t)B * H * NC * 8192 + 256;
}

bool tc_backward_supported(const Args &a) {
    return a.io_dtype == WKV6_BF16 && a.w_kind == W_RAW_BF16 && a.mask == nullptr && a.T >= 1 && !a.s0_f32 &&
           tc::get_encode_fn() != nullptr;
}

int tc_backward(const Args &a) {
    if (a.B * a.H == 0 || a.T == 0) return WKV6_OK;
    const int C = a.H * 64;
    const size_t NC = (size_t)(a.T + L - 1) / L;
    const size_t simt_ws = simt_backward_workspace_bytes(a.B, a.T, a.H);
    const size_t need = a.saved ? simt_ws : tc_backward_workspace_bytes(a.B, a.T, a.H);
    if (!a.workspace || a.workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes", need);
        return WKV6_EWORKSPACE;
    }
    uint8_t *ws = (uint8_t *)a.workspace;
    bf16 *ckpt;
    int *flag;
    if (a.saved) {
        // training pair: the forward already left the chunk-start states and the hazard flag
        flag = (int *)a.saved;
        ckpt = (bf16 *)((uint8_t *)a.saved + SAVED_HEADER);
    } else {
        ckpt = (bf16 *)(ws + simt_ws);
        flag = (int *)(ws + simt_ws + (size_t)a.B * a.H * NC * 8192);
        WKV6_CUDA_CHECK(cudaMemsetAsync(flag, 0, 256, a.stream));
        // 1. forward pre-pass: state checkpoints at every chunk start + hazard flag (no y)
        Args f = a;
        f.y = nullptr;
        f.sT = nullptr;
        if (int rc = tc_forward_ex(f, ckpt, flag)) return rc;
    }
    // 2. tensor-core reverse sweep (returns at once when the flag is raised)
    CUtensorMap mr, mk, mv, mw, mg, mc;
    const auto dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const bool okm[6] = {tc::make_btc_map(&mr, a.r, a.B, a.T, C, L, dt, 2, 64), tc::make_btc_map(&mk, a.k, a.B, a.T, C, L, dt, 2, 64),
                         tc::make_btc_map(&mv, a.v, a.B, a.T, C, L, dt, 2, 64), tc::make_btc_map(&mw, a.w, a.B, a.T, C, L, dt, 2, 64),
                         tc::make_btc_map(&mg, a.gy, a.B, a.T, C, L, dt, 2, 64),
                         tc::make_btc_map(&mc, ckpt, 1, (int)((size_t)a.B * a.H * NC * 64), 64, 64, dt, 2, 64)};
    for (int i = 0; i < 6; i++)
        if (!okm[i]) {
            const void *ptrs[6] = {a.r, a.k, a.v, a.w, a.gy, ckpt};
            set_error("cuTensorMapEncodeTiled failed for map %d (r,k,v,w,gy,ckpt), pointer %p (must be 16-byte aligned)", i, ptrs[i]);
            return WKV6_ECUDA;
        }
    Params p;
    p.B = a.B; p.T = a.T; p.H = a.H;
    p.u = (const bf16 *)a.u;
    p.has_s0 = a.s0 != nullptr;
    p.gr = (bf16 *)a.gr; p.gk = (bf16 *)a.gk; p.gv = (bf16 *)a.gv; p.gw = (bf16 *)a.gw; p.gu = (bf16 *)a.gu; p.gs = (bf16 *)a.gs;
    p.hz_flag = flag;
    static bool attr_done = false;
    if (!attr_done) {
        WKV6_CUDA_CHECK(cudaFuncSetAttribute(wkv6_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_done = true;
    }
    wkv6_tc_bwd_kernel<<<a.B * a.H, NTHREADS, SMEM_BYTES, a.stream>>>(mr, mk, mv, mw, mg, mc, p);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    // 3. exact route, predicated on the flag
    Args s = a;
    s.run_flag = flag;
    s.run_if = 1;
    s.workspace_bytes = simt_ws;
    return simt_backward(s);
}

}  // namespace wkv6
