// Self-test of the tcgen05 / TMA / TMEM building blocks in tc_common.cuh: one 64 x N x 64 bf16 MMA
// with every operand layout the WKV6 kernels use.  Called from tests/ on the GPU before trusting
// the real kernels (descriptor mistakes produce silent garbage, not faults).
#include <cuda_bf16.h>

#include <stdio.h>

#include "../../rwkv_lm_ext_b200/csrc/tc_common.cuh"

// TEST CODE: built into tests/libwkv6_b200_selftest.so by tests/test_gpu_tc_selftest.py (plain nvcc), not into the
// product library.  It only shares the header of building blocks with the product kernels.
#define WKV6_OK 0
#define WKV6_EINVAL (-1)
#define WKV6_ECUDA (-2)
static char g_selftest_err[256];
static void set_error(const char *msg) { snprintf(g_selftest_err, sizeof(g_selftest_err), "%s", msg); }
static void count_launch() {}
#define WKV6_CUDA_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error(cudaGetErrorString(_e)); return WKV6_ECUDA; } } while (0)
extern "C" __attribute__((visibility("default"))) const char *wkv6b200_selftest_last_error(void) { return g_selftest_err; }

namespace wkv6 {
namespace {

using namespace tc;

// flags: bit0 A is MN-major ([K rows][M]); bit1 B is MN-major ([K rows][N]);
//        bit2 re-write both tiles through the generic proxy with sw128() before the MMA
//        bit3 use rows 16.. of the B tile (start-address offset of 2048 B; K-major B only)
__global__ void __launch_bounds__(128) selftest_kernel(const __grid_constant__ CUtensorMap mapA,
                                                       const __grid_constant__ CUtensorMap mapB, float *D, int flags, int Nn) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *tA = smem, *tB = smem + 8192, *tA2 = smem + 16384, *tB2 = smem + 24576;
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_tma, 16384);
        tma_load_3d(tA, &mapA, &bar_tma, 0, 0, 0);
        tma_load_3d(tB, &mapB, &bar_tma, 0, 0, 0);
    }
    mbar_wait(&bar_tma, 0);
    const uint8_t *useA = tA, *useB = tB;
    if (flags & 4) {
        // copy element-wise through registers with the swizzle arithmetic (validates sw128 both ways)
        for (int e = threadIdx.x; e < 64 * 32; e += blockDim.x) {
            const int row = e >> 5, w4 = (e & 31) * 4;
            *(uint32_t *)(tA2 + sw128(row, w4)) = *(const uint32_t *)(tA + sw128(row, w4));
            *(uint32_t *)(tB2 + sw128(row, w4)) = *(const uint32_t *)(tB + sw128(row, w4));
        }
        fence_proxy_async();
        useA = tA2;
        useB = tB2;
    }
    __syncthreads();

    const int a_mn = flags & 1, b_mn = (flags >> 1) & 1;
    if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t idesc = idesc_bf16(64, Nn, a_mn, b_mn);
        const uint32_t boff = (flags & 8) ? 2048u : 0u;
        for (int k = 0; k < 4; k++) {
            const uint64_t ad = smem_desc_sw128(smem_u32(useA) + (a_mn ? k * 2048 : k * 32), 8192, 1024);
            const uint64_t bd = smem_desc_sw128(smem_u32(useB) + boff + (b_mn ? k * 2048 : k * 32), 8192, 1024);
            mma_bf16_ss(tmem, ad, bd, idesc, k > 0);
        }
        mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    // M = 64 accumulator: row r lives in TMEM lane (r % 16) + 32 * (r / 16)
    uint32_t v0[32], v1[32];
    tmem_ld32(tmem_addr(tmem, 32 * warp, 0), v0);
    tmem_ld32(tmem_addr(tmem, 32 * warp, 32), v1);
    tmem_wait_ld();
    if (lane < 16) {
        const int row = 16 * warp + lane;
        for (int c = 0; c < 32; c++) {
            if (c < Nn) D[row * 64 + c] = __uint_as_float(v0[c]);
            if (32 + c < Nn) D[row * 64 + 32 + c] = __uint_as_float(v1[c]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// Second self-test: the fragment-layout building blocks of the role-uniform kernels.
//   A: bf16 [64 t][64 i] tile, B: bf16 [64 k][64 n] tile.
//   F[i][t]    <- what each thread believes it holds after ldmatrix.x4.trans of the A tile
//   O[i][t]    <- stmatrix.x4 (no trans) of those fragments into a swizzled tile + TMA store  (= A^T)
//   A2         <- stmatrix.x4.trans of the fragments (= A again), used as the K-major MMA A operand
//   D[m][n]    <- sum_k A2[m][k] * B[k][16*co + n], n < 16: MN-major B, N = 16, column offset inside
//                 the 128-byte swizzle row; read back with tcgen05.ld.16x256b
__global__ void __launch_bounds__(256) selftest2_kernel(const __grid_constant__ CUtensorMap mapA,
                                                        const __grid_constant__ CUtensorMap mapB,
                                                        const __grid_constant__ CUtensorMap mapO, float *F, float *D, int co) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *tA = smem, *tB = smem + 8192, *tA2 = smem + 16384, *tO = smem + 24576;
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp = warp & 3, ch = warp >> 2, ri = lane >> 2, q = lane & 3;

    if (threadIdx.x == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_tma, 16384);
        tma_load_3d(tA, &mapA, &bar_tma, 0, 0, 0);
        tma_load_3d(tB, &mapB, &bar_tma, 0, 0, 0);
    }
    mbar_wait(&bar_tma, 0);

    // fragments: x[h][g] = (channel 16sp + 8h + ri, tokens 32ch + 8g + 2q + {0,1})
    uint32_t x[2][4];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t a = smem_u32(tA) + sw128(32 * ch + 8 * (lane >> 3) + (lane & 7), 32 * sp + 16 * h);
        ldsm_x4_t(a, x[h][0], x[h][1], x[h][2], x[h][3]);
    }
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int i = 16 * sp + 8 * h + ri, t = 32 * ch + 8 * g + 2 * q;
            F[i * 64 + t] = __uint_as_float(x[h][g] << 16);
            F[i * 64 + t + 1] = __uint_as_float(x[h][g] & 0xffff0000u);
        }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        // back to [t][i]: transposing store, rows t = 32ch + 8g + r, 16-byte chunk of channels 16sp + 8h
        const uint32_t a2 = smem_u32(tA2) + sw128(32 * ch + 8 * (lane >> 3) + (lane & 7), 32 * sp + 16 * h);
        stsm_x4_t(a2, x[h][0], x[h][1], x[h][2], x[h][3]);
        // [i][t]: plain store, rows i = 16sp + 8h + r, 16-byte chunks of tokens 32ch + 8g
        const uint32_t ao = smem_u32(tO) + sw128(16 * sp + 8 * h + (lane & 7), 64 * ch + 16 * (lane >> 3));
        stsm_x4(ao, x[h][0], x[h][1], x[h][2], x[h][3]);
    }
    // shadow lanes: an M = 64 accumulator only occupies lanes 0-15 of each 32-lane sub-partition; park a
    // pattern in lanes 16-31 of the accumulator columns and check below that the MMA leaves it alone
    {
        uint32_t pat[16];
#pragma unroll
        for (int x = 0; x < 16; x++) pat[x] = __float_as_uint(1000.f + 64.f * threadIdx.x + x);
        tmem_st_frag(tmem_addr(tmem, 32 * sp + 16, 32 * ch), pat);
        tmem_wait_st();
        tc_fence_before();
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_store_3d(&mapO, tO, 0, 0, 0);
        tma_store_commit();
        tc_fence_after();
        const uint32_t idesc = idesc_bf16(64, 16, 0, 1);
        for (int k = 0; k < 4; k++)
            mma_bf16_ss(tmem, smem_desc_sw128(smem_u32(tA2) + 32 * k, 8192, 1024),
                        smem_desc_sw128(smem_u32(tB) + 32 * co + 2048 * k, 8192, 1024), idesc, k > 0);
        mma_commit(&bar_mma);
        tma_store_wait_all<0>();
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld_frag(tmem_addr(tmem, 32 * sp, 32 * ch), v);
    tmem_wait_ld();
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
        for (int hh = 0; hh < 2; hh++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int row = 16 * sp + 8 * hh + ri, col = 32 * ch + 8 * g + 2 * q + e;
                if (col < 16) D[row * 16 + col] = __uint_as_float(v[4 * g + 2 * hh + e]);
            }
    {
        uint32_t pat[16];
        tmem_ld_frag(tmem_addr(tmem, 32 * sp + 16, 32 * ch), pat);
        tmem_wait_ld();
        int bad = 0;
#pragma unroll
        for (int x = 0; x < 16; x++) bad |= (__uint_as_float(pat[x]) != 1000.f + 64.f * threadIdx.x + x);
        if (bad) D[64 * 16] = 1.f;        // one extra float after the D matrix: shadow lanes were disturbed
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

}  // namespace
}  // namespace wkv6

using namespace wkv6;

extern "C" __attribute__((visibility("default"))) int wkv6b200_tc_selftest2(int co, const void *A, const void *B, void *O,
                                                                              float *F, float *D, void *stream) {
    if (co < 0 || co > 3) { set_error("selftest2: bad column offset"); return WKV6_EINVAL; }
    CUtensorMap mA, mB, mO;
    if (!tc::make_btc_map(&mA, A, 1, 64, 64, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64) ||
        !tc::make_btc_map(&mB, B, 1, 64, 64, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64) ||
        !tc::make_btc_map(&mO, O, 1, 64, 64, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64)) {
        set_error("selftest2: cuTensorMapEncodeTiled failed");
        return WKV6_ECUDA;
    }
    const int smem = 4 * 8192 + 1024;
    WKV6_CUDA_CHECK(cudaFuncSetAttribute(selftest2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    selftest2_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(mA, mB, mO, F, D, co);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

// Not part of the reference-facing ABI (not declared in include/wkv6_b200.h): test hook only.
extern "C" __attribute__((visibility("default"))) int wkv6b200_tc_selftest(int flags, int Nn, const void *A, const void *B,
                                                                             float *D, void *stream) {
    if (Nn < 8 || Nn > 64 || (Nn & 7)) { set_error("selftest: bad N"); return WKV6_EINVAL; }
    CUtensorMap mA, mB;
    if (!tc::make_btc_map(&mA, A, 1, 64, 64, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64) ||
        !tc::make_btc_map(&mB, B, 1, 64, 64, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64)) {
        set_error("selftest: cuTensorMapEncodeTiled failed");
        return WKV6_ECUDA;
    }
    const int smem = 4 * 8192 + 1024;
    WKV6_CUDA_CHECK(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mA, mB, D, flags, Nn);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}
