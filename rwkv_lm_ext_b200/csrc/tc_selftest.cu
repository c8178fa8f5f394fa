// Self-test of the tcgen05 / TMA / TMEM building blocks in tc_common.cuh: one 64 x N x 64 bf16 MMA
// with every operand layout the WKV6 kernels use.  Called from tests/ on the GPU before trusting
// the real kernels (descriptor mistakes produce silent garbage, not faults).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

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

}  // namespace
}  // namespace wkv6

using namespace wkv6;

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
