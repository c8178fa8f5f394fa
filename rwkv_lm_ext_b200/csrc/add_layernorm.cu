// Residual add + LayerNorm in one pass (the LayerNorm / residual glue of `Block.forward`,
// src/model.py:904-933):
//
//     x_new = x + delta                (bf16, rounded like the eager add)
//     y     = LayerNorm(x_new) * w + b (fp32 statistics, bf16 result, like torch's bf16 layer_norm)
//
// The eager chain reads and writes x_new twice more (add: 6 B/element, LayerNorm: 4 B/element); fused it is
// 8 B/element forward (x, delta in; x_new, y out).  Backward: g = g_xnew + dLN(g_y) serves BOTH x and delta (one
// tensor), 8 B/element; the affine-parameter gradients (frozen under LoRA / state tuning: skipped then) come from
// a second, column-wise streaming pass with deterministic fixed-order partial sums.
// One warp per row, 128-bit accesses, the row kept in registers between the statistics and the output (D <= 2560;
// larger rows are re-read from L1/L2).
#include "common.cuh"

namespace wkv6 {
namespace {

typedef __nv_bfloat16 bf16;
// 8 bf16 as one 128-bit access.  (A struct of four __nv_bfloat162 is loaded as four 32-bit words; and a load guarded by
// `if (j < nv)` inside the loop that consumes it is not hoisted over the stores of the iteration before -- the first
// version of these kernels had one 32-bit round trip to memory in flight per lane and ran at 59 % of the HBM peak.)
typedef uint4 bf16x8;
__device__ __forceinline__ bf16x8 ld8(const bf16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ void st8(bf16 *p, const bf16x8 &x) { *reinterpret_cast<uint4 *>(p) = x; }
__device__ __forceinline__ void unpack8(const bf16x8 &x, float *f) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; i++) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ bf16x8 pack8(const float *f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t *>(&t);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

constexpr int KMAX = 10;          // vectors of 8 per lane kept in registers (packed bf16): D <= 32 * 8 * KMAX = 2560
// K (template): register vectors of the instantiation, the smallest of 4 / 8 / 10 that holds the row; 0 = rows wider
// than that (streamed twice, the second time from L1 / L2)

// lane l of the row's warp owns the vectors j*32 + l, j = 0..nv-1 (coalesced 512-byte segments).  REG: every load of
// the row is issued before the first use (unconditionally: the vectors past nv re-read the last valid one), the row
// then stays in registers as packed bf16 between the statistics and the output.
template <bool HAS_DELTA, int K>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(long long rows, int D, float eps, const bf16 *__restrict__ x,
                                                         const bf16 *__restrict__ delta, const bf16 *__restrict__ w,
                                                         const bf16 *__restrict__ b, bf16 *__restrict__ x_new,
                                                         bf16 *__restrict__ y, float2 *__restrict__ stats) {
    const int lane = threadIdx.x & 31, nv = D / 256;      // full vectors per lane; D % 256 == 0
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        const bf16 *xr = x + row * D, *dr = HAS_DELTA ? delta + row * D : nullptr;
        bf16 *xo = HAS_DELTA ? x_new + row * D : nullptr, *yr = y + row * D;
        auto col = [&](int j) { return (min(j, nv - 1) * 32 + lane) * 8; };
        // x_new = bf16(x + delta) (the eager bf16 add) of one vector, its sum added to s
        auto add = [&](const bf16x8 &vx, const bf16x8 &vd, float &s) {
            float f[8], d[8];
            unpack8(vx, f);
            if (HAS_DELTA) {
                unpack8(vd, d);
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = f[i] + d[i];
            }
            const bf16x8 o = HAS_DELTA ? pack8(f) : vx;
            if (HAS_DELTA) unpack8(o, f);
#pragma unroll
            for (int i = 0; i < 8; i++) s += f[i];
            return o;
        };
        auto sq = [&](const bf16x8 &v, float mean, float &q) {
            float f[8];
            unpack8(v, f);
#pragma unroll
            for (int i = 0; i < 8; i++) { const float t = f[i] - mean; q = fmaf(t, t, q); }
        };
        auto emit = [&](int j, const bf16x8 &v, float mean, float rstd) {
            const int c = (j * 32 + lane) * 8;
            float f[8], wf[8], bfv[8], o[8];
            unpack8(v, f);
            unpack8(ld8(w + c), wf);
            unpack8(ld8(b + c), bfv);
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = fmaf((f[i] - mean) * rstd, wf[i], bfv[i]);
            st8(yr + c, pack8(o));
        };
        float s = 0.f, q = 0.f;
        if (K > 0) {
            constexpr int KK = K > 0 ? K : 1;
            bf16x8 px[KK], pd[HAS_DELTA ? KK : 1];
#pragma unroll
            for (int j = 0; j < KK; j++) {
                px[j] = ld8(xr + col(j));
                if (HAS_DELTA) pd[j] = ld8(dr + col(j));
            }
#pragma unroll
            for (int j = 0; j < KK; j++)
                if (j < nv) {
                    px[j] = add(px[j], pd[HAS_DELTA ? j : 0], s);
                    if (HAS_DELTA) st8(xo + col(j), px[j]);
                }
            const float mean = warp_sum(s) / (float)D;
#pragma unroll
            for (int j = 0; j < KK; j++)
                if (j < nv) sq(px[j], mean, q);
            const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
            if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
#pragma unroll
            for (int j = 0; j < KK; j++)
                if (j < nv) emit(j, px[j], mean, rstd);
        } else {
            for (int j = 0; j < nv; j++) {
                const bf16x8 o = add(ld8(xr + col(j)), HAS_DELTA ? ld8(dr + col(j)) : make_uint4(0, 0, 0, 0), s);
                if (HAS_DELTA) st8(xo + col(j), o);
            }
            const float mean = warp_sum(s) / (float)D;
            const bf16 *src = HAS_DELTA ? xo : xr;                         // re-read (L1 / L2): x_new was just written by this lane
            for (int j = 0; j < nv; j++) sq(*reinterpret_cast<const uint4 *>(src + col(j)), mean, q);
            const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
            if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
            for (int j = 0; j < nv; j++) emit(j, *reinterpret_cast<const uint4 *>(src + col(j)), mean, rstd);
        }
    }
}

// g = g_xnew + rstd * (dxh - mean(dxh) - xh * mean(dxh * xh)),  dxh = g_y * w,  xh = (x_new - mean) * rstd
// REG: the row's x_new and g_y are loaded up front and stay in registers as packed bf16 between the two sweeps
template <bool HAS_GX, int K>
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(long long rows, int D, const bf16 *__restrict__ x_new,
                                                         const float2 *__restrict__ stats, const bf16 *__restrict__ w,
                                                         const bf16 *__restrict__ g_y, const bf16 *__restrict__ g_xnew,
                                                         bf16 *__restrict__ g) {
    const int lane = threadIdx.x & 31, nv = D / 256;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        const bf16 *xr = x_new + row * D, *gyr = g_y + row * D, *gxr = HAS_GX ? g_xnew + row * D : nullptr;
        bf16 *gr = g + row * D;
        const float2 st = stats[row];
        auto col = [&](int j) { return (min(j, nv - 1) * 32 + lane) * 8; };
        float s1 = 0.f, s2 = 0.f;
        auto sums = [&](int j, const bf16x8 &vx, const bf16x8 &vg) {
            float fx[8], fd[8], wf[8];
            unpack8(vx, fx);
            unpack8(vg, fd);
            unpack8(ld8(w + col(j)), wf);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float d = fd[i] * wf[i];
                s1 += d;
                s2 = fmaf(d, (fx[i] - st.x) * st.y, s2);
            }
        };
        auto emit = [&](int j, const bf16x8 &vx, const bf16x8 &vg, const bf16x8 &vgx, float m1, float m2) {
            const int c = (j * 32 + lane) * 8;
            float fx[8], fd[8], wf[8], o[8], gx[8];
            unpack8(vx, fx);
            unpack8(vg, fd);
            unpack8(ld8(w + c), wf);
            if (HAS_GX) unpack8(vgx, gx);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float t = st.y * (fd[i] * wf[i] - m1 - (fx[i] - st.x) * st.y * m2);
                o[i] = HAS_GX ? gx[i] + t : t;
            }
            st8(gr + c, pack8(o));
        };
        if (K > 0) {
            constexpr int KK = K > 0 ? K : 1;
            bf16x8 px[KK], pg[KK], pgx[HAS_GX ? KK : 1];
#pragma unroll
            for (int j = 0; j < KK; j++) {
                px[j] = ld8(xr + col(j));
                pg[j] = ld8(gyr + col(j));
            }
            if (HAS_GX) {                                  // behind the two it is needed after: in flight under the first sweep
#pragma unroll
                for (int j = 0; j < KK; j++) pgx[j] = ld8(gxr + col(j));
            }
#pragma unroll
            for (int j = 0; j < KK; j++)
                if (j < nv) sums(j, px[j], pg[j]);
            const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
#pragma unroll
            for (int j = 0; j < KK; j++)
                if (j < nv) emit(j, px[j], pg[j], pgx[HAS_GX ? j : 0], m1, m2);
        } else {
            for (int j = 0; j < nv; j++) sums(j, ld8(xr + col(j)), ld8(gyr + col(j)));
            const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
            for (int j = 0; j < nv; j++)                   // re-read (L1 / L2)
                emit(j, ld8(xr + col(j)), ld8(gyr + col(j)), HAS_GX ? ld8(gxr + col(j)) : make_uint4(0, 0, 0, 0), m1, m2);
        }
    }
}

// gw[c] = sum_rows g_y * xh,  gb[c] = sum_rows g_y: thread = 8 columns, grid.y = row splits, fixed-order partials
__global__ void __launch_bounds__(256) add_ln_param_kernel(long long rows, int D, long long rows_per_split,
                                                           const bf16 *__restrict__ x_new, const float2 *__restrict__ stats,
                                                           const bf16 *__restrict__ g_y, float *__restrict__ partial) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (c >= D) return;
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = min(rows, r0 + rows_per_split);
    float aw[8] = {}, ab[8] = {};
    for (long long row = r0; row < r1; row++) {
        const float2 st = stats[row];
        float fx[8], fg[8];
        unpack8(ld8(x_new + row * D + c), fx);
        unpack8(ld8(g_y + row * D + c), fg);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            aw[i] = fmaf(fg[i], (fx[i] - st.x) * st.y, aw[i]);
            ab[i] += fg[i];
        }
    }
    float *pw = partial + (size_t)blockIdx.y * 2 * D, *pb = pw + D;
#pragma unroll
    for (int i = 0; i < 8; i++) { pw[c + i] = aw[i]; pb[c + i] = ab[i]; }
}

__global__ void add_ln_param_sum_kernel(int S, int D, const float *__restrict__ partial, float *__restrict__ gw,
                                        float *__restrict__ gb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D) return;
    float sw = 0.f, sb = 0.f;
    for (int k = 0; k < S; k++) { sw += partial[(size_t)k * 2 * D + i]; sb += partial[(size_t)k * 2 * D + D + i]; }
    gw[i] = sw;
    gb[i] = sb;
}

int row_grid(long long rows) {
    const long long blocks = (rows + 7) / 8;
    return (int)std::min<long long>(blocks, 148 * 8);          // 8 resident blocks of 8 warps per SM, grid-stride beyond
}
int param_splits(long long rows) { return (int)std::max<long long>(1, std::min<long long>(64, rows / 256)); }

}  // namespace
}  // namespace wkv6

using namespace wkv6;

extern "C" {

int add_layernorm_bf16(long long rows, int D, float eps, const void *x, const void *delta, const void *w, const void *b,
                       void *x_new, void *y, float *stats, void *stream) {
    if (rows < 0 || D <= 0 || (D % 256)) { set_error("add_layernorm_bf16: need D %% 256 == 0"); return WKV6_EINVAL; }
    if (rows == 0) return WKV6_OK;
    if (!x || !w || !b || !y || (delta && !x_new)) { set_error("add_layernorm_bf16: null pointer"); return WKV6_EINVAL; }
    const int nvec = D / 256, kreg = nvec <= 4 ? 4 : nvec <= 8 ? 8 : nvec <= KMAX ? KMAX : 0;
    const int grid = row_grid(rows);
    auto s = (cudaStream_t)stream;
#define LAUNCH(HD, RG) add_ln_fwd_kernel<HD, RG><<<grid, 256, 0, s>>>(rows, D, eps, (const bf16 *)x, (const bf16 *)delta, \
        (const bf16 *)w, (const bf16 *)b, (bf16 *)x_new, (bf16 *)y, (float2 *)stats)
#define PICK(HD) do { if (kreg == 4) LAUNCH(HD, 4); else if (kreg == 8) LAUNCH(HD, 8); else if (kreg == KMAX) LAUNCH(HD, KMAX); else LAUNCH(HD, 0); } while (0)
    if (delta) PICK(true); else PICK(false);
#undef PICK
#undef LAUNCH
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

size_t add_layernorm_backward_workspace_bytes(long long rows, int D) {
    if (rows <= 0 || D <= 0) return 0;
    return (size_t)param_splits(rows) * 2 * D * sizeof(float);
}

int add_layernorm_backward_bf16(long long rows, int D, const void *x_new, const float *stats, const void *w,
                                const void *g_y, const void *g_xnew, void *g, float *gw, float *gb, void *workspace,
                                size_t workspace_bytes, void *stream) {
    if (rows < 0 || D <= 0 || (D % 256)) { set_error("add_layernorm_backward_bf16: need D %% 256 == 0"); return WKV6_EINVAL; }
    if (rows == 0) return WKV6_OK;
    if (!x_new || !stats || !w || !g_y || !g) { set_error("add_layernorm_backward_bf16: null pointer"); return WKV6_EINVAL; }
    const int nvec = D / 256, kreg = nvec <= 4 ? 4 : nvec <= 8 ? 8 : nvec <= KMAX ? KMAX : 0;
    const int grid = row_grid(rows);
    auto s = (cudaStream_t)stream;
#define LAUNCH(HG, RG) add_ln_bwd_kernel<HG, RG><<<grid, 256, 0, s>>>(rows, D, (const bf16 *)x_new, (const float2 *)stats, \
        (const bf16 *)w, (const bf16 *)g_y, (const bf16 *)g_xnew, (bf16 *)g)
#define PICK(HG) do { if (kreg == 4) LAUNCH(HG, 4); else if (kreg == 8) LAUNCH(HG, 8); else if (kreg == KMAX) LAUNCH(HG, KMAX); else LAUNCH(HG, 0); } while (0)
    if (g_xnew) PICK(true); else PICK(false);
#undef PICK
#undef LAUNCH
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    if (gw || gb) {
        if (!gw || !gb) { set_error("add_layernorm_backward_bf16: gw and gb come together"); return WKV6_EINVAL; }
        const int S = param_splits(rows);
        if (!workspace || workspace_bytes < (size_t)S * 2 * D * sizeof(float)) {
            set_error("add_layernorm_backward_bf16: workspace too small");
            return WKV6_EWORKSPACE;
        }
        const long long rps = (rows + S - 1) / S;
        dim3 pg((D / 8 + 255) / 256, S);
        add_ln_param_kernel<<<pg, 256, 0, s>>>(rows, D, rps, (const bf16 *)x_new, (const float2 *)stats, (const bf16 *)g_y,
                                               (float *)workspace);
        add_ln_param_sum_kernel<<<(D + 255) / 256, 256, 0, s>>>(S, D, (const float *)workspace, gw, gb);
        count_launch(2);
        WKV6_CUDA_CHECK(cudaGetLastError());
    }
    return WKV6_OK;
}

}  // extern "C"
