// WKV6 forward, chunked, on the 5th-generation tensor cores (tcgen05) fed by TMA.
//
// One CTA owns one (batch, head) stream and walks its T tokens in chunks of L = 64.  Per chunk
// (i = key channel, j = value channel, l_t = -exp(w_t), cum = inclusive
// prefix sum of l inside the chunk, exc = cum - l, Lam = cum at the chunk end):
//
//   (1) A^T[s,t]  = sum_i Kt_q[s,i] * Rt[t,i]          4 MMAs 64x16x64 (one per 16-token target block q)
//        Rt[t]    = r_t * exp(exc_t - rho_q),  Kt_q[s] = k_s * exp(rho_q - cum_s),  rho_q = exc at the
//        middle of block q: every factor is within 8 decay steps of 1, so nothing overflows
//   (3) Y[t,j]    = sum_i Rh[t,i] * S[i,j]             Rh = r * exp(exc)           (state from earlier chunks)
//   (2) Y[t,j]   += sum_s P[t,s] * V[s,j]              P = strict-lower(A) + diag(sum_i r u k)
//   (4) S[j,i]    = exp(Lam_i) * S[j,i] + sum_s V[s,j] * Kh[s,i]      Kh = k * exp(Lam - cum)
//
// All are bf16 x bf16 -> fp32 tcgen05.mma with M = 64; accumulators (A^T, Y, and the fp32 master
// copy of the state S) live in TMEM; operands are 64x64 bf16 tiles in shared memory in the
// 128-byte-swizzled layout that both TMA and the UMMA descriptors use, read K-major or MN-major as
// each product needs, so no tile is ever transposed in memory.  r,k,v,w tiles arrive by TMA
// (3-D tensor map over [B,T,C], rows past T are zero-filled by the hardware); v is consumed in
// place as an MMA operand.  Kh is split into a bf16 high and low part (two accumulating MMAs) so
// that the carried fp32 state keeps ~16 mantissa bits per update instead of 8.
//
// Warp roles (416 threads, 2 CTAs per SM so that one stream's serial phases overlap the other's):
//   warps 0-3  : TMEM side -- mask A^T into P, decay-scale S, write y, refresh the bf16 copy of S
//   warps 4-11 : operand preparation (decay prefix sums, exp2, scaling), 8 rows x 2 channels each
//   warp  12   : issues TMA and every tcgen05.mma, polls the mbarriers
// Waiting is done by ONE polling warp per group (mbarrier.try_wait with a suspend hint) which then
// releases its group through a hardware named barrier: spinning warps would otherwise starve the
// working ones of issue slots.
// Exactness guard: if any channel decays by more than e^-60 within half a block, the chunk takes
// the "hazard" route -- references move to the block starts (all factors <= 1) and the 16x16
// diagonal blocks of A are recomputed pairwise in fp32 on the CUDA cores.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace wkv6 {
namespace {

using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int L = 64;
constexpr int EPI_THREADS = 128, PREP_THREADS = 256, NTHREADS = EPI_THREADS + PREP_THREADS + 32;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float HAZARD2 = 60.0f * LOG2E;   // in log2 units

// shared memory map (bytes from a 1024-aligned base)
constexpr uint32_t OFF_R = 0, OFF_K = 8192, OFF_W = 16384, OFF_V = 24576;      // V: 2 stages
constexpr uint32_t OFF_KT = 40960;                                              // 160 rows, q=3,2,1,0
__host__ __device__ constexpr uint32_t kt_off(int q) { return q == 0 ? 18432u : q == 1 ? 14336u : q == 2 ? 8192u : 0u; }
constexpr uint32_t OFF_RT = 61440, OFF_P = OFF_RT;   // P is written after the MMAs reading Rt are done
constexpr uint32_t OFF_RH = 69632, OFF_KH = 77824, OFF_KL = 86016, OFF_SB = 94208;
constexpr uint32_t OFF_TILES_END = 102400;
struct Extra {
    alignas(16) float elam[64];
    float diagu[64];
    float2 htot[8][32];       // per half-block (8 rows) decay totals, [half-block][channel pair]
    float dg[4][16][16];      // hazard route: exact diagonal blocks of A, [block][s][t]
    uint64_t bar_rkw, bar_v[2], bar_prep, bar_mma_a, bar_p, bar_mma_b, bar_sb;
    uint32_t tmem_base;
    int hz[2];
};
constexpr uint32_t SMEM_BYTES = OFF_TILES_END + sizeof(Extra);

constexpr uint32_t TM_A = 0, TM_Y = 64, TM_S = 128, TM_COLS = 256;

struct Params {
    int B, T, H;
    const bf16 *u;
    const void *s0;
    int s0_f32;
    long long s0_bstride;
    void *sT;
    int sT_f32;
    bf16 *y;          // nullptr: state-only pass (backward pre-pass)
    bf16 *ckpt;       // nullptr or [B*H][NC][64 j][64 i]: bf16 state at the START of every chunk
    int *hz_flag;     // nullptr or a device int set to 1 when any chunk took the hazard route
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
__device__ __forceinline__ uint4 pack8(const uint32_t *v) {
    uint4 o;
    o.x = pack2(__uint_as_float(v[0]), __uint_as_float(v[1]));
    o.y = pack2(__uint_as_float(v[2]), __uint_as_float(v[3]));
    o.z = pack2(__uint_as_float(v[4]), __uint_as_float(v[5]));
    o.w = pack2(__uint_as_float(v[6]), __uint_as_float(v[7]));
    return o;
}

__global__ void __launch_bounds__(NTHREADS, 2)
wkv6_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_w, Params p) {
    // dynamic shared memory is the only shared memory of this kernel, so it starts 1024-aligned
    // (checked below); indexing it directly keeps every access in the shared address space (LDS/STS)
    extern __shared__ __align__(1024) uint8_t sm[];
    Extra &ex = *reinterpret_cast<Extra *>(sm + OFF_TILES_END);
    if ((smem_u32(sm) & 1023u) != 0) __trap();
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int T = p.T, C = p.H * 64;
    const int NC = (T + L - 1) / L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&ex.bar_rkw, 1);
        mbar_init(&ex.bar_v[0], 1);
        mbar_init(&ex.bar_v[1], 1);
        mbar_init(&ex.bar_prep, PREP_THREADS);
        mbar_init(&ex.bar_mma_a, 1);
        mbar_init(&ex.bar_p, EPI_THREADS);
        mbar_init(&ex.bar_mma_b, 1);
        mbar_init(&ex.bar_sb, EPI_THREADS);
        ex.hz[0] = 0;
        ex.hz[1] = 0;
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

    if (warp == 12) {
        // =====================================================================================
        // issuer warp: TMA loads + all tcgen05.mma (lane 0 issues, the whole warp polls)
        // =====================================================================================
        auto issue_loads = [&](int c) {
            mbar_arrive_expect_tx(&ex.bar_rkw, 3 * 8192);
            tma_load_3d(sm + OFF_R, &map_r, &ex.bar_rkw, h * 64, c * L, b);
            tma_load_3d(sm + OFF_K, &map_k, &ex.bar_rkw, h * 64, c * L, b);
            tma_load_3d(sm + OFF_W, &map_w, &ex.bar_rkw, h * 64, c * L, b);
            mbar_arrive_expect_tx(&ex.bar_v[c & 1], 8192);
            tma_load_3d(sm + OFF_V + (c & 1) * 8192, &map_v, &ex.bar_v[c & 1], h * 64, c * L, b);
        };
        if (lane == 0) {
            tma_prefetch_desc(&map_r);
            tma_prefetch_desc(&map_k);
            tma_prefetch_desc(&map_v);
            tma_prefetch_desc(&map_w);
            issue_loads(0);
        }
        const uint32_t kt = smem_u32(sm + OFF_KT), rt = smem_u32(sm + OFF_RT), rh = smem_u32(sm + OFF_RH);
        const uint32_t kh = smem_u32(sm + OFF_KH), kl = smem_u32(sm + OFF_KL), pp = smem_u32(sm + OFF_P);
        const uint32_t sb = smem_u32(sm + OFF_SB);
        constexpr uint32_t ID_A = idesc_bf16(64, 16, 0, 0);
        constexpr uint32_t ID_KK = idesc_bf16(64, 64, 0, 0);
        constexpr uint32_t ID_KM = idesc_bf16(64, 64, 0, 1);
        constexpr uint32_t ID_MM = idesc_bf16(64, 64, 1, 1);
        for (int c = 0; c < NC; c++) {
            const uint32_t par = c & 1;
            mbar_wait(&ex.bar_rkw, par);
            named_bar_sync<1, PREP_THREADS + 32>();          // release the prep warps into chunk c
            mbar_wait(&ex.bar_prep, par);                    // operands written, raw r,k,w consumed
            if (lane == 0 && c + 1 < NC) issue_loads(c + 1);
            mbar_wait(&ex.bar_v[par], (c >> 1) & 1);
            mbar_wait(&ex.bar_sb, par);                      // bf16 S of this chunk ready, TMEM A/Y free
            tc_fence_after();
            const uint32_t vv = smem_u32(sm + OFF_V + par * 8192);
            if (lane == 0) {
#pragma unroll
                for (int qq = 0; qq < 4; qq++)               // (1) A^T blocks
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        mma_bf16_ss(tmem + TM_A + 16 * qq, smem_desc_sw128(kt + kt_off(qq) + 32 * k, 8192, 1024),
                                    smem_desc_sw128(rt + 2048 * qq + 32 * k, 8192, 1024), ID_A, k > 0);
#pragma unroll
                for (int k = 0; k < 4; k++)                  // (3) Y = Rh * S_in
                    mma_bf16_ss(tmem + TM_Y, smem_desc_sw128(rh + 32 * k, 8192, 1024),
                                smem_desc_sw128(sb + 32 * k, 8192, 1024), ID_KK, k > 0);
                mma_commit(&ex.bar_mma_a);
            }
            __syncwarp();
            mbar_wait(&ex.bar_p, par);                       // P written, S decayed
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; k++)                  // (2) Y += P * V
                    mma_bf16_ss(tmem + TM_Y, smem_desc_sw128(pp + 32 * k, 8192, 1024),
                                smem_desc_sw128(vv + 2048 * k, 8192, 1024), ID_KM, 1);
#pragma unroll
                for (int k = 0; k < 4; k++)                  // (4) S += V^T * (Kh_hi + Kh_lo)
                    mma_bf16_ss(tmem + TM_S, smem_desc_sw128(vv + 2048 * k, 8192, 1024),
                                smem_desc_sw128(kh + 2048 * k, 8192, 1024), ID_MM, 1);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    mma_bf16_ss(tmem + TM_S, smem_desc_sw128(vv + 2048 * k, 8192, 1024),
                                smem_desc_sw128(kl + 2048 * k, 8192, 1024), ID_MM, 1);
                mma_commit(&ex.bar_mma_b);
            }
            __syncwarp();
            mbar_wait(&ex.bar_mma_b, par);                   // operand tiles are free again
        }
    } else if (warp >= 4) {
        // =====================================================================================
        // operand preparation: warp pw handles half-block hb = pw (8 token rows), lane = channel pair
        // =====================================================================================
        const int hb = warp - 4, q = hb >> 1;
        const uint32_t boff = ((uint32_t)(lane >> 2) << 4) | ((uint32_t)(lane & 3) << 2);
        const uint32_t rowbase = 1024u * hb;
        const float u0 = __bfloat162float(p.u[h * 64 + 2 * lane]), u1 = __bfloat162float(p.u[h * 64 + 2 * lane + 1]);
        uint8_t *sR = sm + OFF_R, *sK = sm + OFF_K, *sW = sm + OFF_W;

        for (int c = 0; c < NC; c++) {
            const uint32_t par = c & 1;
            const int nv = min(L, T - c * L);
            named_bar_sync<1, PREP_THREADS + 32>();          // raw tiles landed, operand tiles free

            // ---- phase A: log2-decays of my 8 rows x 2 channels, half-block totals, hazard check
            float l0[8], l1[8];
            float hs0 = 0.f, hs1 = 0.f;
#pragma unroll
            for (int n = 0; n < 8; n++) {
                const uint32_t ww = *reinterpret_cast<const uint32_t *>(sW + rowbase + n * 128 + (boff ^ (n << 4)));
                const bool valid = (8 * hb + n) < nv;
                l0[n] = valid ? -fast_ex2(bf_lo(ww) * LOG2E) * LOG2E : 0.f;
                l1[n] = valid ? -fast_ex2(bf_hi(ww) * LOG2E) * LOG2E : 0.f;
                hs0 += l0[n];
                hs1 += l1[n];
            }
            ex.htot[hb][lane] = make_float2(hs0, hs1);
            if (-hs0 > HAZARD2 || -hs1 > HAZARD2) {
                atomicOr(&ex.hz[par], 1);
                if (p.hz_flag) *p.hz_flag = 1;
            }
            named_bar_sync<2, PREP_THREADS>();

            // ---- phase B: references, per-(block, channel) factors
            const bool hazard = ex.hz[par] != 0;
            if (hb == 0 && lane == 0) ex.hz[par ^ 1] = 0;
            float run0 = 0.f, run1 = 0.f, b0 = 0.f, b1 = 0.f, rho0[4], rho1[4];
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const float2 hv = ex.htot[x][lane];
                if (x == hb) { b0 = run0; b1 = run1; }
                if ((x & 1) == 0) { rho0[x >> 1] = run0; rho1[x >> 1] = run1; }          // block start
                else if (!hazard) { rho0[x >> 1] = run0; rho1[x >> 1] = run1; }           // block middle
                run0 += hv.x;
                run1 += hv.y;
            }
            const float lam0 = run0, lam1 = run1;
            if (hb == 0) *reinterpret_cast<float2 *>(&ex.elam[2 * lane]) = make_float2(fast_ex2(lam0), fast_ex2(lam1));
            const float rq0 = q == 0 ? rho0[0] : q == 1 ? rho0[1] : q == 2 ? rho0[2] : rho0[3];
            const float rq1 = q == 0 ? rho1[0] : q == 1 ? rho1[1] : q == 2 ? rho1[2] : rho1[3];
            const float er0 = fast_ex2(rq0), er1 = fast_ex2(rq1);                 // Rh = Rt * exp(rho)
            const float el0 = fast_ex2(lam0 - rq0), el1 = fast_ex2(lam1 - rq1);   // Kh = (k F) * exp(Lam - rho)
            float g0[4], g1[4];                                                   // Kt_{q'} = (k F) * exp(rho_q' - rho_q)
#pragma unroll
            for (int qq = 1; qq < 4; qq++) {
                g0[qq] = fast_ex2(fminf(rho0[qq] - rq0, 0.f));
                g1[qq] = fast_ex2(fminf(rho1[qq] - rq1, 0.f));
            }

            float cum0 = b0, cum1 = b1;
            float du[8];
#pragma unroll
            for (int n = 0; n < 8; n++) {
                const uint32_t off = rowbase + n * 128 + (boff ^ (n << 4));
                const uint32_t rr = *reinterpret_cast<const uint32_t *>(sR + off);
                const uint32_t kk = *reinterpret_cast<const uint32_t *>(sK + off);
                const float r0 = bf_lo(rr), r1 = bf_hi(rr), k0 = bf_lo(kk), k1 = bf_hi(kk);
                const float exc0 = cum0, exc1 = cum1;
                cum0 += l0[n];
                cum1 += l1[n];
                const float rt0 = r0 * fast_ex2(exc0 - rq0), rt1 = r1 * fast_ex2(exc1 - rq1);
                *reinterpret_cast<uint32_t *>(sm + OFF_RT + off) = pack2(rt0, rt1);
                *reinterpret_cast<uint32_t *>(sm + OFF_RH + off) = pack2(rt0 * er0, rt1 * er1);
                float kh0, kh1;
                if (!hazard) {
                    const float kf0 = k0 * fast_ex2(rq0 - cum0), kf1 = k1 * fast_ex2(rq1 - cum1);
                    *reinterpret_cast<uint32_t *>(sm + OFF_KT + kt_off(q) + off) = pack2(kf0, kf1);
#pragma unroll
                    for (int qq = 1; qq < 4; qq++)
                        if (qq > q) *reinterpret_cast<uint32_t *>(sm + OFF_KT + kt_off(qq) + off) = pack2(kf0 * g0[qq], kf1 * g1[qq]);
                    kh0 = kf0 * el0;
                    kh1 = kf1 * el1;
                } else {
                    // block-start references: off-diagonal factors are all <= 1; the diagonal block
                    // (clamped here) is replaced by the exact pairwise values below
                    *reinterpret_cast<uint32_t *>(sm + OFF_KT + kt_off(q) + off) =
                        pack2(k0 * fast_ex2(fminf(rq0 - cum0, 100.f)), k1 * fast_ex2(fminf(rq1 - cum1, 100.f)));
#pragma unroll
                    for (int qq = 1; qq < 4; qq++)
                        if (qq > q)
                            *reinterpret_cast<uint32_t *>(sm + OFF_KT + kt_off(qq) + off) =
                                pack2(k0 * fast_ex2(rho0[qq] - cum0), k1 * fast_ex2(rho1[qq] - cum1));
                    kh0 = k0 * fast_ex2(lam0 - cum0);
                    kh1 = k1 * fast_ex2(lam1 - cum1);
                }
                const uint32_t khi = pack2(kh0, kh1);
                *reinterpret_cast<uint32_t *>(sm + OFF_KH + off) = khi;
                *reinterpret_cast<uint32_t *>(sm + OFF_KL + off) = pack2(kh0 - bf_lo(khi), kh1 - bf_hi(khi));
                du[n] = r0 * u0 * k0 + r1 * u1 * k1;
            }
            // ---- diag(u) term: sum over the 64 channels (= the 32 lanes) of r u k, 8 rows at once
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
            if (hazard) {
                // exact A[t,s], s < t inside my 16-token block, for my 8 target rows:
                //   sum_i r_t[i] k_s[i] exp(sum_{m=s+1}^{t-1} l_m[i])      (one warp = all 64 channels)
                for (int n = 0; n < 8; n++) {
                    const int nt = 8 * (hb & 1) + n;
                    const uint32_t rr = *reinterpret_cast<const uint32_t *>(sR + sw128(16 * q + nt, 4 * lane));
                    const float r0 = bf_lo(rr), r1 = bf_hi(rr);
                    float acc0 = 0.f, acc1 = 0.f;
                    for (int ns = nt - 1; ns >= 0; ns--) {
                        const int ts = 16 * q + ns;
                        const uint32_t kk = *reinterpret_cast<const uint32_t *>(sK + sw128(ts, 4 * lane));
                        float term = r0 * bf_lo(kk) * fast_ex2(acc0) + r1 * bf_hi(kk) * fast_ex2(acc1);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
                        if (lane == 0) ex.dg[q][ns][nt] = term;
                        const uint32_t ww = *reinterpret_cast<const uint32_t *>(sW + sw128(ts, 4 * lane));
                        if (ts < nv) {
                            acc0 -= fast_ex2(bf_lo(ww) * LOG2E) * LOG2E;
                            acc1 -= fast_ex2(bf_hi(ww) * LOG2E) * LOG2E;
                        }
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&ex.bar_prep);
        }
    } else {
        // =====================================================================================
        // TMEM side: warp w owns rows 16w..16w+15 of every 64-row accumulator (lanes 0-15 of its
        // TMEM sub-partition); lanes 16-31 execute the aligned tcgen05 instructions but hold nothing.
        // =====================================================================================
        const int row = 16 * warp + (lane & 15);
        const bool act = lane < 16;
        const uint32_t tlane = 32 * warp;
        uint32_t v[32];

        // initial state -> TMEM (fp32 master) and shared (bf16 operand copy); layout [value j][key i]
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
#pragma unroll
            for (int cc = 0; cc < 32; cc++) {
                float x = 0.f;
                if (p.s0 && act) {
                    const size_t idx = (size_t)b * p.s0_bstride + ((size_t)h * 64 + row) * 64 + hh * 32 + cc;
                    x = p.s0_f32 ? ((const float *)p.s0)[idx] : __bfloat162float(((const bf16 *)p.s0)[idx]);
                }
                v[cc] = __float_as_uint(x);
            }
            tmem_st32(tmem_addr(tmem, tlane, TM_S + 32 * hh), v);
            if (act) {
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    const uint4 o = pack8(v + 8 * ch);
                    *reinterpret_cast<uint4 *>(sm + OFF_SB + sw128(row, 64 * hh + 16 * ch)) = o;
                    if (p.ckpt) *reinterpret_cast<uint4 *>(p.ckpt + ((size_t)blockIdx.x * NC * 64 + row) * 64 + 32 * hh + 8 * ch) = o;
                }
            }
        }
        tmem_wait_st();
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&ex.bar_sb);

        for (int c = 0; c < NC; c++) {
            const uint32_t par = c & 1;
            const int nv = min(L, T - c * L);
            if (warp == 0) {
                mbar_wait(&ex.bar_prep, par);
                mbar_wait(&ex.bar_mma_a, par);
            }
            named_bar_sync<3, EPI_THREADS>();
            tc_fence_after();
            const bool hazard = ex.hz[par] != 0;
            // ---- A^T (lanes = s, columns = t)  ->  P[t][s] bf16, strictly lower + diag(u) term
            //      P[t][s = row] lives at byte t*128 + ((s/8) ^ (t%8))*16 + (s%8)*2 of the swizzled tile
            const uint32_t pcol = ((uint32_t)(row >> 3) << 4), pin = (uint32_t)(row & 7) * 2;
            uint8_t *pbase = sm + OFF_P + pin;
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                tmem_ld32(tmem_addr(tmem, tlane, TM_A + 32 * hh), v);
                tmem_wait_ld();
                if (act) {
#pragma unroll
                    for (int cc = 0; cc < 32; cc++) {
                        const int t = 32 * hh + cc;
                        const float x = (t > row) ? __uint_as_float(v[cc]) : 0.f;
                        *reinterpret_cast<bf16 *>(pbase + t * 128 + (pcol ^ ((uint32_t)(t & 7) << 4))) = __float2bfloat16_rn(x);
                    }
                }
            }
            if (act) {
                *reinterpret_cast<bf16 *>(pbase + row * 128 + (pcol ^ ((uint32_t)(row & 7) << 4))) = __float2bfloat16_rn(ex.diagu[row]);
                if (hazard) {
                    const int qb = row & ~15;
                    for (int t = row + 1; t < qb + 16; t++)
                        *reinterpret_cast<bf16 *>(pbase + t * 128 + (pcol ^ ((uint32_t)(t & 7) << 4))) =
                            __float2bfloat16_rn(ex.dg[row >> 4][row & 15][t & 15]);
                }
            }
            // ---- decay the state: S[j][i] *= exp(Lam_i)
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                tmem_ld32(tmem_addr(tmem, tlane, TM_S + 32 * hh), v);
                tmem_wait_ld();
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++) {
                    const float4 e = *reinterpret_cast<const float4 *>(&ex.elam[32 * hh + 4 * c4]);
                    v[4 * c4 + 0] = __float_as_uint(__uint_as_float(v[4 * c4 + 0]) * e.x);
                    v[4 * c4 + 1] = __float_as_uint(__uint_as_float(v[4 * c4 + 1]) * e.y);
                    v[4 * c4 + 2] = __float_as_uint(__uint_as_float(v[4 * c4 + 2]) * e.z);
                    v[4 * c4 + 3] = __float_as_uint(__uint_as_float(v[4 * c4 + 3]) * e.w);
                }
                tmem_st32(tmem_addr(tmem, tlane, TM_S + 32 * hh), v);
            }
            tmem_wait_st();
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&ex.bar_p);

            if (warp == 0) mbar_wait(&ex.bar_mma_b, par);
            named_bar_sync<3, EPI_THREADS>();
            tc_fence_after();
            // ---- y rows
            {
                bf16 *dst = p.y + ((size_t)b * T + (size_t)c * L + row) * C + (size_t)h * 64;
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    tmem_ld32(tmem_addr(tmem, tlane, TM_Y + 32 * hh), v);
                    tmem_wait_ld();
                    if (act && row < nv && p.y) {
#pragma unroll
                        for (int ch = 0; ch < 4; ch++) *reinterpret_cast<uint4 *>(dst + 32 * hh + 8 * ch) = pack8(v + 8 * ch);
                    }
                }
            }
            // ---- new state -> bf16 operand copy (and the caller's buffer after the last chunk)
            const bool last = (c == NC - 1);
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                tmem_ld32(tmem_addr(tmem, tlane, TM_S + 32 * hh), v);
                tmem_wait_ld();
                if (act) {
#pragma unroll
                    for (int ch = 0; ch < 4; ch++) {
                        const uint4 o = pack8(v + 8 * ch);
                        *reinterpret_cast<uint4 *>(sm + OFF_SB + sw128(row, 64 * hh + 16 * ch)) = o;
                        if (p.ckpt && !last)
                            *reinterpret_cast<uint4 *>(p.ckpt + (((size_t)blockIdx.x * NC + c + 1) * 64 + row) * 64 + 32 * hh + 8 * ch) = o;
                    }
                    if (last && p.sT) {
                        const size_t idx = (((size_t)b * p.H + h) * 64 + row) * 64 + hh * 32;
#pragma unroll
                        for (int cc = 0; cc < 32; cc++) {
                            if (p.sT_f32) ((float *)p.sT)[idx + cc] = __uint_as_float(v[cc]);
                            else ((bf16 *)p.sT)[idx + cc] = __float2bfloat16_rn(__uint_as_float(v[cc]));
                        }
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&ex.bar_sb);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace

bool tc_forward_supported(const Args &a) {
    return a.io_dtype == WKV6_BF16 && a.w_kind == W_RAW_BF16 && a.mask == nullptr && a.T >= 1 &&
           tc::get_encode_fn() != nullptr;
}

int tc_forward_ex(const Args &a, void *ckpt, int *hz_flag);
int tc_forward(const Args &a) { return tc_forward_ex(a, nullptr, nullptr); }

// ckpt: nullptr or bf16 [B*H][ceil(T/64)][64][64] receiving the state at the start of every chunk;
// a.y may be nullptr (state-only pass).
int tc_forward_ex(const Args &a, void *ckpt, int *hz_flag) {
    if (a.B * a.H == 0 || a.T == 0) return WKV6_OK;
    const int C = a.H * 64;
    CUtensorMap mr, mk, mv, mw;
    const auto dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (!tc::make_btc_map(&mr, a.r, a.B, a.T, C, L, dt, 2, 64) || !tc::make_btc_map(&mk, a.k, a.B, a.T, C, L, dt, 2, 64) ||
        !tc::make_btc_map(&mv, a.v, a.B, a.T, C, L, dt, 2, 64) || !tc::make_btc_map(&mw, a.w, a.B, a.T, C, L, dt, 2, 64)) {
        set_error("cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned)");
        return WKV6_ECUDA;
    }
    Params p;
    p.B = a.B; p.T = a.T; p.H = a.H;
    p.u = (const bf16 *)a.u;
    p.s0 = a.s0; p.s0_f32 = a.s0_f32; p.s0_bstride = a.s0_bstride;
    p.sT = a.sT; p.sT_f32 = a.sT_f32;
    p.y = (bf16 *)a.y;
    p.ckpt = (bf16 *)ckpt;
    p.hz_flag = hz_flag;
    static bool attr_done = false;
    if (!attr_done) {
        WKV6_CUDA_CHECK(cudaFuncSetAttribute(wkv6_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        WKV6_CUDA_CHECK(cudaFuncSetAttribute(wkv6_tc_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             cudaSharedmemCarveoutMaxShared));
        attr_done = true;
    }
    wkv6_tc_fwd_kernel<<<a.B * a.H, NTHREADS, SMEM_BYTES, a.stream>>>(mr, mk, mv, mw, p);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

}  // namespace wkv6
