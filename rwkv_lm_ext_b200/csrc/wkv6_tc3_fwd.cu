// WKV6 forward, chunked, on tcgen05 tensor cores fed by TMA -- role-uniform version.
//
// One CTA owns one (batch, head) stream and walks its T tokens in chunks of L = 64; 2 CTAs per SM.
// Per chunk (i = key channel, j = value channel, l_t = -exp(w_t), cum =
// inclusive prefix sum of l inside the chunk, exc = cum - l, Lam = cum at the chunk end):
//
//   M1  A^T[s,t] = sum_i Kt_q[s,i] * Rt[t,i]        4 x (64 x 16 x 64), one per 16-token target block q
//        Rt[t] = r_t * 2^(exc_t - rho_q),  Kt_q[s] = k_s * 2^(rho_q - cum_s),  rho_q = exc at the middle
//        of block q rounded to the INTEGER log2 grid: every factor stays within 8 decay steps of 1,
//        and the versions Kt_q are exact power-of-two multiples of one another (bf16 mul.rn, no exp)
//       Y[t,j]   = sum_i Rh[t,i] * S[i,j]           Rh = r * 2^exc          (state from earlier chunks)
//   T1  P = strict-lower(A) + diag(sum_i r u k)  (bf16),   S[i,j] *= 2^Lam_i  (in TMEM)
//   M2  Y[t,j]  += sum_s P[t,s] * V[s,j];    S[i,j] += sum_s Kh[s,i] * V[s,j]    Kh = k * 2^(Lam - cum),
//        split into bf16 hi + lo so that the fp32 master state keeps ~16 mantissa bits per update
//   T2  y tile -> TMA store;  bf16 copy of S for the next chunk's M1 (and, for training, TMA-stored
//        as the chunk-start checkpoint the backward kernel reads)
//
// 8 compute warps do EVERY stage in the fragment mapping of tc3_common.cuh (a thread keeps its two
// channels from the decay scan to the state rows it rescales); warp 8 issues TMA and tcgen05.mma.
// While one CTA waits for its MMAs the co-resident CTA computes.  Chunks whose decay is too strong
// for the block references (more than e^-60 inside an aligned 16-token span) raise the stream's hazard flag: the
// exact SIMT kernel, enqueued behind this one and predicated per stream on that flag, redoes them.
#include "common.cuh"
#include "tc3_common.cuh"

namespace wkv6 {
namespace {

using namespace tc3;

constexpr uint32_t OFF_R = 0, OFF_K = 8192, OFF_W = 16384, OFF_V = 24576;
constexpr uint32_t OFF_KT = 32768;                       // 160 rows: versions q = 3,2,1,0
__host__ __device__ constexpr uint32_t kt_off(int q) { return q == 0 ? 18432u : q == 1 ? 14336u : q == 2 ? 8192u : 0u; }
constexpr uint32_t OFF_RT = 53248, OFF_P = OFF_RT;       // P is written after the MMAs reading Rt are done
constexpr uint32_t OFF_RH = 61440, OFF_KH = 69632, OFF_KL = 77824, OFF_SB = 86016, OFF_YT = 94208;
constexpr uint32_t OFF_TILES_END = 102400;
struct Extra {
    float gtot[8][64];        // decay total of every 8-token group, per channel (log2 units)
    float pdu[4][64];         // per channel-quarter partial sums of r u k, per token
    uint64_t bar_rkw, bar_v, bar_m1, bar_m2;
    uint32_t tmem_base;
    int hz;
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
    int has_y, has_ckpt;
    int *hz_flags;            // [B*H]: 1 = this stream needs the exact route
};

// 9 warps x 2 CTAs = 5 warps on the fullest SM sub-partition (16384 registers): at most 96 per thread
__global__ void __launch_bounds__(NTHREADS, 2)
wkv6_tc3_fwd_kernel(const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_ck, Params p) {
    extern __shared__ __align__(1024) uint8_t sm[];
    Extra &ex = *reinterpret_cast<Extra *>(sm + OFF_TILES_END);
    if ((smem_u32(sm) & 1023u) != 0) __trap();
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int T = p.T;
    const int NC = (T + L - 1) / L;
    Frag F;
    F.init();
    const int warp = F.warp, lane = F.lane;

    if (threadIdx.x == 0) {
        mbar_init(&ex.bar_rkw, 1);
        mbar_init(&ex.bar_v, 1);
        mbar_init(&ex.bar_m1, 1);
        mbar_init(&ex.bar_m2, 1);
        ex.hz = 0;
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
    const uint32_t sbase = smem_u32(sm);

    if (warp == CWARPS) {
        // =====================================================================================
        // issuer: TMA loads / stores and every tcgen05.mma, one thread
        // =====================================================================================
        auto issue_rkw = [&](int c) {
            mbar_arrive_expect_tx(&ex.bar_rkw, 3 * 8192);
            tma_load_3d(sm + OFF_R, &map_r, &ex.bar_rkw, h * 64, c * L, b);
            tma_load_3d(sm + OFF_K, &map_k, &ex.bar_rkw, h * 64, c * L, b);
            tma_load_3d(sm + OFF_W, &map_w, &ex.bar_rkw, h * 64, c * L, b);
        };
        auto issue_v = [&](int c) {
            mbar_arrive_expect_tx(&ex.bar_v, 8192);
            tma_load_3d(sm + OFF_V, &map_v, &ex.bar_v, h * 64, c * L, b);
        };
        if (lane == 0) {
            tma_prefetch_desc(&map_r);
            tma_prefetch_desc(&map_k);
            tma_prefetch_desc(&map_v);
            tma_prefetch_desc(&map_w);
            issue_rkw(0);
            issue_v(0);
        }
        const uint32_t kt = sbase + OFF_KT, rt = sbase + OFF_RT, rh = sbase + OFF_RH, kh = sbase + OFF_KH;
        const uint32_t kl = sbase + OFF_KL, pp = sbase + OFF_P, sb = sbase + OFF_SB, vv = sbase + OFF_V;
        constexpr uint32_t ID_A = idesc_bf16(64, 16, 0, 0);
        constexpr uint32_t ID_KM = idesc_bf16(64, 64, 0, 1);
        constexpr uint32_t ID_MM = idesc_bf16(64, 64, 1, 1);
        for (int c = 0; c < NC; c++) {
            const uint32_t par = c & 1;
            if (lane == 0) mbar_wait(&ex.bar_rkw, par);
            __syncwarp();
            bar_arrive_all<B_RAW>();                             // raw r,k,w landed
            bar_sync_all<B_PREP>();                              // operands written, raw r,k,w consumed
            if (lane == 0 && c + 1 < NC) issue_rkw(c + 1);
            bar_sync_all<B_T2>();                                // completion #c: bf16 S ready, TMEM A / Y free
            if (lane == 0) {
                if (p.has_ckpt) tma_store_3d(&map_ck, sm + OFF_SB, 0, (blockIdx.x * NC + c) * 64, 0);
                if (c > 0 && p.has_y) tma_store_3d(&map_y, sm + OFF_YT, h * 64, (c - 1) * L, b);
                tma_store_commit();
                tc_fence_after();
#pragma unroll
                for (int qq = 0; qq < 4; qq++)                   // A^T blocks
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        mma_bf16_ss(tmem + TM_A + 16 * qq, smem_desc_sw128(kt + kt_off(qq) + 32 * k, 8192, 1024),
                                    smem_desc_sw128(rt + 2048 * qq + 32 * k, 8192, 1024), ID_A, k > 0);
#pragma unroll
                for (int k = 0; k < 4; k++)                      // Y = Rh * S_in      (S tile is [i][j]: MN-major B)
                    mma_bf16_ss(tmem + TM_Y, smem_desc_sw128(rh + 32 * k, 8192, 1024),
                                smem_desc_sw128(sb + 2048 * k, 8192, 1024), ID_KM, k > 0);
                mma_commit(&ex.bar_m1);
                mbar_wait(&ex.bar_m1, par);
            }
            __syncwarp();
            bar_arrive_all<B_M1>();
            bar_sync_all<B_T1>();                                // P written, S decayed
            if (lane == 0) {
                mbar_wait(&ex.bar_v, par);
                tma_store_wait_read<0>();                        // SB / YT may be rewritten once M2 is done
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; k++)                      // Y += P * V
                    mma_bf16_ss(tmem + TM_Y, smem_desc_sw128(pp + 32 * k, 8192, 1024),
                                smem_desc_sw128(vv + 2048 * k, 8192, 1024), ID_KM, 1);
#pragma unroll
                for (int k = 0; k < 4; k++)                      // S[i,j] += Kh^T V  (hi, then lo)
                    mma_bf16_ss(tmem + TM_S, smem_desc_sw128(kh + 2048 * k, 8192, 1024),
                                smem_desc_sw128(vv + 2048 * k, 8192, 1024), ID_MM, 1);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    mma_bf16_ss(tmem + TM_S, smem_desc_sw128(kl + 2048 * k, 8192, 1024),
                                smem_desc_sw128(vv + 2048 * k, 8192, 1024), ID_MM, 1);
                mma_commit(&ex.bar_m2);
                mbar_wait(&ex.bar_m2, par);                      // V is free again
                if (c + 1 < NC) issue_v(c + 1);
            }
            __syncwarp();
            bar_arrive_all<B_M2>();
        }
        bar_sync_all<B_T2>();                                    // completion #NC
        if (lane == 0) {
            if (p.has_y) tma_store_3d(&map_y, sm + OFF_YT, h * 64, (NC - 1) * L, b);
            tma_store_commit();
            tma_store_wait_all<0>();
        }
    } else {
        // =====================================================================================
        // compute warps
        // =====================================================================================
        const int sp = F.sp, ch = F.ch, q = F.q;
        const float u_h[2] = {__bfloat162float(p.u[h * 64 + F.row(0)]), __bfloat162float(p.u[h * 64 + F.row(1)])};
        const uint32_t tS = tmem_addr(tmem, 32 * sp, TM_S + 32 * ch);
        uint32_t v[16];

        // ---- initial state -> TMEM (fp32 master, [i][j]) and shared (bf16 operand copy)
#pragma unroll
        for (int g = 0; g < 4; g++)
#pragma unroll
            for (int hh = 0; hh < 2; hh++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    float x = 0.f;
                    if (p.s0) {   // caller layout [.,H,64(value j),64(key i)]
                        const size_t idx = (size_t)b * p.s0_bstride + ((size_t)h * 64 + F.col(g, e)) * 64 + F.row(hh);
                        x = p.s0_f32 ? ((const float *)p.s0)[idx] : __bfloat162float(((const bf16 *)p.s0)[idx]);
                    }
                    v[4 * g + 2 * hh + e] = __float_as_uint(x);
                }
        tmem_st_frag(tS, v);
#pragma unroll
        for (int hh = 0; hh < 2; hh++)
            stsm_x4(sbase + OFF_SB + F.rc(hh), pack2(__uint_as_float(v[0 + 2 * hh]), __uint_as_float(v[1 + 2 * hh])),
                    pack2(__uint_as_float(v[4 + 2 * hh]), __uint_as_float(v[5 + 2 * hh])),
                    pack2(__uint_as_float(v[8 + 2 * hh]), __uint_as_float(v[9 + 2 * hh])),
                    pack2(__uint_as_float(v[12 + 2 * hh]), __uint_as_float(v[13 + 2 * hh])));
        tmem_wait_st();
        fence_proxy_async();
        tc_fence_before();
        bar_arrive_all<B_T2>();

        for (int c = 0; c < NC; c++) {
            const int nv = min(L, T - c * L);
            // ================================================================== P: operand preparation
            bar_sync_all<B_RAW>();
            float l[2][4][2], exq[2][4];
            {
                uint32_t wp[2][4];
                ldsm_x4_t(sbase + OFF_W + F.ti(0), wp[0][0], wp[0][1], wp[0][2], wp[0][3]);
                ldsm_x4_t(sbase + OFF_W + F.ti(1), wp[1][0], wp[1][1], wp[1][2], wp[1][3]);
                bool hazard = false;
                float prev = 0.f;
#pragma unroll
                for (int hh = 0; hh < 2; hh++)
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        float l0 = -fast_ex2(bf_lo(wp[hh][g]) * LOG2E) * LOG2E;
                        float l1 = -fast_ex2(bf_hi(wp[hh][g]) * LOG2E) * LOG2E;
                        if (nv < L) {                                   // ragged last chunk: no decay on the padded rows
                            const int t0 = F.col(g, 0);
                            if (t0 >= nv) l0 = 0.f;
                            if (t0 + 1 >= nv) l1 = 0.f;
                        }
                        l[hh][g][0] = l0;
                        l[hh][g][1] = l1;
                        const float ps = l0 + l1;
                        float x = ps, y;
                        y = __shfl_up_sync(0xffffffffu, x, 1, 4);
                        if (q >= 1) x += y;
                        y = __shfl_up_sync(0xffffffffu, x, 2, 4);
                        if (q >= 2) x += y;
                        exq[hh][g] = x - ps;
                        if (q == 3) ex.gtot[4 * ch + g][F.row(hh)] = x;
                        // exactness guard: total decay of every aligned 16-token span (two groups)
                        if (g & 1) hazard |= (-(x + prev) > HAZARD2);
                        prev = x;
                    }
                if (hazard && q == 3) ex.hz = 1;
            }
            named_bar_sync<B_SCAN, CTHREADS>();

            uint32_t rr[2][4], kk[2][4];
            ldsm_x4_t(sbase + OFF_R + F.ti(0), rr[0][0], rr[0][1], rr[0][2], rr[0][3]);
            ldsm_x4_t(sbase + OFF_R + F.ti(1), rr[1][0], rr[1][1], rr[1][2], rr[1][3]);
            ldsm_x4_t(sbase + OFF_K + F.ti(0), kk[0][0], kk[0][1], kk[0][2], kk[0][3]);
            ldsm_x4_t(sbase + OFF_K + F.ti(1), kk[1][0], kk[1][1], kk[1][2], kk[1][3]);
            float elam[2];
            float du[4][2];
#pragma unroll
            for (int g = 0; g < 4; g++) du[g][0] = du[g][1] = 0.f;
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                // prefix over the 8 groups of this channel: start of my groups, block references, total
                float run = 0.f, gb[4], rho[4];
#pragma unroll
                for (int x8 = 0; x8 < 8; x8++) {
                    if (x8 & 1) rho[x8 >> 1] = rintf(run);                 // block middle, integer log2 grid
                    if ((x8 >> 2) == ch) gb[x8 & 3] = run;
                    run += ex.gtot[x8][F.row(hh)];
                }
                const float lam = run;
                elam[hh] = fast_ex2(lam);
                const float rqa = ch ? rho[2] : rho[0], rqb = ch ? rho[3] : rho[1];      // my two blocks
                const int ir0 = (int)rho[0], ir1 = (int)rho[1], ir2 = (int)rho[2], ir3 = (int)rho[3];
                const uint32_t erqa = bfpow2pair(ch ? ir2 : ir0), erqb = bfpow2pair(ch ? ir3 : ir1);
                uint32_t rto[4], kto[4], rhp[4], khp[4], klp[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const float rq = (g < 2) ? rqa : rqb;
                    const float exc0 = gb[g] + exq[hh][g], cum0 = exc0 + l[hh][g][0], cum1 = cum0 + l[hh][g][1];
                    const float r0 = bf_lo(rr[hh][g]), r1 = bf_hi(rr[hh][g]), k0 = bf_lo(kk[hh][g]), k1 = bf_hi(kk[hh][g]);
                    const float rt0 = r0 * fast_ex2(exc0 - rq), rt1 = r1 * fast_ex2(cum0 - rq);
                    const float kf0 = k0 * fast_ex2(rq - cum0), kf1 = k1 * fast_ex2(rq - cum1);
                    rto[g] = pack2(rt0, rt1);
                    kto[g] = pack2(kf0, kf1);
                    rhp[g] = hmul2(rto[g], (g < 2) ? erqa : erqb);                      // Rh = Rt * 2^rho (exact)
                    const float el = fast_ex2(lam - rq);
                    const float kh0 = kf0 * el, kh1 = kf1 * el;
                    khp[g] = pack2(kh0, kh1);
                    klp[g] = pack2(kh0 - bf_lo(khp[g]), kh1 - bf_hi(khp[g]));
                    du[g][0] = fmaf(r0 * u_h[hh], k0, du[g][0]);
                    du[g][1] = fmaf(r1 * u_h[hh], k1, du[g][1]);
                }
                const uint32_t ti = F.ti(hh);
                stsm_x4_t(sbase + OFF_RT + ti, rto[0], rto[1], rto[2], rto[3]);
                stsm_x4_t(sbase + OFF_RH + ti, rhp[0], rhp[1], rhp[2], rhp[3]);
                stsm_x4_t(sbase + OFF_KH + ti, khp[0], khp[1], khp[2], khp[3]);
                stsm_x4_t(sbase + OFF_KL + ti, klp[0], klp[1], klp[2], klp[3]);
                // Kt versions: version qq holds rows s <= 16qq+15 scaled to the reference of block qq
                if (ch == 0) {
                    const uint32_t f10 = bfpow2pair(ir1 - ir0);
                    const uint32_t f20 = bfpow2pair(ir2 - ir0), f21 = bfpow2pair(ir2 - ir1);
                    const uint32_t f30 = bfpow2pair(ir3 - ir0), f31 = bfpow2pair(ir3 - ir1);
                    stsm_x2_t(sbase + OFF_KT + kt_off(0) + ti, kto[0], kto[1]);
                    stsm_x4_t(sbase + OFF_KT + kt_off(1) + ti, hmul2(kto[0], f10), hmul2(kto[1], f10), kto[2], kto[3]);
                    stsm_x4_t(sbase + OFF_KT + kt_off(2) + ti, hmul2(kto[0], f20), hmul2(kto[1], f20), hmul2(kto[2], f21), hmul2(kto[3], f21));
                    stsm_x4_t(sbase + OFF_KT + kt_off(3) + ti, hmul2(kto[0], f30), hmul2(kto[1], f30), hmul2(kto[2], f31), hmul2(kto[3], f31));
                } else {
                    const uint32_t f32_ = bfpow2pair(ir3 - ir2);
                    stsm_x2_t(sbase + OFF_KT + kt_off(2) + ti, kto[0], kto[1]);
                    stsm_x4_t(sbase + OFF_KT + kt_off(3) + ti, hmul2(kto[0], f32_), hmul2(kto[1], f32_), kto[2], kto[3]);
                }
            }
            // ---- diag(u) term: sum over channels of r u k per token; reduce-scatter over the 8 lanes ri
            {
                const bool b2 = lane & 16, b1 = lane & 8, b0 = lane & 4;
                float a4[2][2], a2[2], a1;
#pragma unroll
                for (int gg = 0; gg < 2; gg++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const float keep = b2 ? du[2 + gg][e] : du[gg][e], send = b2 ? du[gg][e] : du[2 + gg][e];
                        a4[gg][e] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const float keep = b1 ? a4[1][e] : a4[0][e], send = b1 ? a4[0][e] : a4[1][e];
                    a2[e] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                {
                    const float keep = b0 ? a2[1] : a2[0], send = b0 ? a2[0] : a2[1];
                    a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                ex.pdu[sp][32 * ch + 8 * ((b2 ? 2 : 0) + (b1 ? 1 : 0)) + 2 * q + (b0 ? 1 : 0)] = a1;
            }
            fence_proxy_async();
            bar_arrive_all<B_PREP>();

            // ================================================================== T1: A^T -> P, decay S
            bar_sync_all<B_M1>();
            tc_fence_after();
            tmem_ld_frag(tmem_addr(tmem, 32 * sp, TM_A + 32 * ch), v);
            tmem_wait_ld();
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                // 8x8 blocks: rows s in block 2sp+hh, columns t in block 4ch+g; only the diagonal block is mixed
                const int sb8 = 2 * sp + hh;
                uint32_t pk[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const int tb8 = 4 * ch + g;
                    const float a0 = __uint_as_float(v[4 * g + 2 * hh]), a1 = __uint_as_float(v[4 * g + 2 * hh + 1]);
                    if (tb8 > sb8) pk[g] = pack2(a0, a1);
                    else if (tb8 < sb8) pk[g] = 0u;
                    else {
                        const int s = F.row(hh);
                        const float dg = ex.pdu[0][s] + ex.pdu[1][s] + ex.pdu[2][s] + ex.pdu[3][s];
                        const int d = 2 * q - F.ri;                       // (t - s) for e = 0
                        pk[g] = pack2(d > 0 ? a0 : (d == 0 ? dg : 0.f), d + 1 > 0 ? a1 : (d + 1 == 0 ? dg : 0.f));
                    }
                }
                stsm_x4_t(sbase + OFF_P + F.ti(hh), pk[0], pk[1], pk[2], pk[3]);       // P[t][s]
            }
            tmem_ld_frag(tS, v);
            tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < 4; g++)
#pragma unroll
                for (int hh = 0; hh < 2; hh++)
#pragma unroll
                    for (int e = 0; e < 2; e++)
                        v[4 * g + 2 * hh + e] = __float_as_uint(__uint_as_float(v[4 * g + 2 * hh + e]) * elam[hh]);
            tmem_st_frag(tS, v);
            tmem_wait_st();
            fence_proxy_async();
            tc_fence_before();
            bar_arrive_all<B_T1>();

            // ================================================================== T2: y tile, new bf16 S
            bar_sync_all<B_M2>();
            tc_fence_after();
            if (p.has_y) {
                tmem_ld_frag(tmem_addr(tmem, 32 * sp, TM_Y + 32 * ch), v);
                tmem_wait_ld();
#pragma unroll
                for (int hh = 0; hh < 2; hh++)
                    stsm_x4(sbase + OFF_YT + F.rc(hh), pack2(__uint_as_float(v[0 + 2 * hh]), __uint_as_float(v[1 + 2 * hh])),
                            pack2(__uint_as_float(v[4 + 2 * hh]), __uint_as_float(v[5 + 2 * hh])),
                            pack2(__uint_as_float(v[8 + 2 * hh]), __uint_as_float(v[9 + 2 * hh])),
                            pack2(__uint_as_float(v[12 + 2 * hh]), __uint_as_float(v[13 + 2 * hh])));
            }
            tmem_ld_frag(tS, v);
            tmem_wait_ld();
#pragma unroll
            for (int hh = 0; hh < 2; hh++)
                stsm_x4(sbase + OFF_SB + F.rc(hh), pack2(__uint_as_float(v[0 + 2 * hh]), __uint_as_float(v[1 + 2 * hh])),
                        pack2(__uint_as_float(v[4 + 2 * hh]), __uint_as_float(v[5 + 2 * hh])),
                        pack2(__uint_as_float(v[8 + 2 * hh]), __uint_as_float(v[9 + 2 * hh])),
                        pack2(__uint_as_float(v[12 + 2 * hh]), __uint_as_float(v[13 + 2 * hh])));
            if (c == NC - 1 && p.sT && !ex.hz) {
#pragma unroll
                for (int g = 0; g < 4; g++)
#pragma unroll
                    for (int hh = 0; hh < 2; hh++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const size_t idx = (((size_t)b * p.H + h) * 64 + F.col(g, e)) * 64 + F.row(hh);
                            const float x = __uint_as_float(v[4 * g + 2 * hh + e]);
                            if (p.sT_f32) ((float *)p.sT)[idx] = x;
                            else ((bf16 *)p.sT)[idx] = __float2bfloat16_rn(x);
                        }
            }
            fence_proxy_async();
            tc_fence_before();
            bar_arrive_all<B_T2>();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && ex.hz) p.hz_flags[blockIdx.x] = 1;
    if (warp == 0) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace

bool tc3_forward_supported(const Args &a) {
    return a.io_dtype == WKV6_BF16 && a.w_kind == W_RAW_BF16 && a.mask == nullptr && a.T >= 1 &&
           tc::get_encode_fn() != nullptr;
}

// ckpt: nullptr or bf16 [B*H][ceil(T/64)][64 i][64 j] receiving the state at the start of every chunk;
// hz_flags: device int [B*H], zeroed by the caller; a.y may be nullptr (state-only pass).
int tc3_forward(const Args &a, void *ckpt, int *hz_flags) {
    if (a.B * a.H == 0 || a.T == 0) return WKV6_OK;
    const int C = a.H * 64;
    const size_t NC = (size_t)(a.T + L - 1) / L;
    CUtensorMap mr, mk, mv, mw, my, mc;
    const auto dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const void *yy = a.y ? a.y : a.r, *cc = ckpt ? ckpt : a.r;    // unused maps still have to encode
    const int ck_rows = ckpt ? (int)((size_t)a.B * a.H * NC * 64) : 64, ck_cols = ckpt ? 64 : C;
    if (!tc::make_btc_map(&mr, a.r, a.B, a.T, C, L, dt, 2, 64) || !tc::make_btc_map(&mk, a.k, a.B, a.T, C, L, dt, 2, 64) ||
        !tc::make_btc_map(&mv, a.v, a.B, a.T, C, L, dt, 2, 64) || !tc::make_btc_map(&mw, a.w, a.B, a.T, C, L, dt, 2, 64) ||
        !tc::make_btc_map(&my, yy, a.B, a.T, C, L, dt, 2, 64) || !tc::make_btc_map(&mc, cc, 1, ck_rows, ck_cols, 64, dt, 2, 64)) {
        set_error("cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned)");
        return WKV6_ECUDA;
    }
    Params p;
    p.B = a.B; p.T = a.T; p.H = a.H;
    p.u = (const bf16 *)a.u;
    p.s0 = a.s0; p.s0_f32 = a.s0_f32; p.s0_bstride = a.s0_bstride;
    p.sT = a.sT; p.sT_f32 = a.sT_f32;
    p.has_y = a.y != nullptr;
    p.has_ckpt = ckpt != nullptr;
    p.hz_flags = hz_flags;
    static bool attr_done = false;
    if (!attr_done) {
        WKV6_CUDA_CHECK(cudaFuncSetAttribute(wkv6_tc3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        WKV6_CUDA_CHECK(cudaFuncSetAttribute(wkv6_tc3_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             cudaSharedmemCarveoutMaxShared));
        attr_done = true;
    }
    wkv6_tc3_fwd_kernel<<<a.B * a.H, NTHREADS, SMEM_BYTES, a.stream>>>(mr, mk, mv, mw, my, mc, p);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

}  // namespace wkv6
