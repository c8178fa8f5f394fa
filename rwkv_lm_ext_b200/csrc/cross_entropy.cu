// Token cross entropy + L2Wrap of the SFT step on bf16 logits (src/model.py:1244-1283 `training_step`, my_qa_mask == 0
// branch, and `L2Wrap` src/model.py:960-974), fp32 arithmetic:
//
//     loss        = mean over the rows with target != ignore of  lse(x_row) - x_row[target]
//     dL/dx[r,j]  = g * (softmax(x_r)_j - [j == target_r]) / n_valid   (0 for an ignored row)
//                   + [j == argmax_j x_r] * max_j x_r * 1e-4 / rows     (L2Wrap: pulls the largest logit towards 0)
//
// The eager chain around the 65536-wide head moves the [rows, V] matrix nine times (fp32 copy, log-softmax forward and
// backward in fp32, the cast back, max, zeros, scatter, add); here the forward reads it once and the backward reads it
// once and writes the gradient once.  One 512-thread block per row; the row stays in registers (packed bf16, 16 x 128
// bit per thread: V <= 65536) between the maximum and the exponential sum.  Row sums are reduced in a fixed order
// (deterministic); ties of the maximum go to the lowest index.
#include "common.cuh"

namespace wkv6 {
namespace {

typedef __nv_bfloat16 bf16;
constexpr int CE_THREADS = 512, CE_MAXV = 16;            // vectors of 8 logits per thread kept in registers
constexpr float LOG2E_F = 1.4426950408889634f, LN2_F = 0.6931471805599453f;

__device__ __forceinline__ float bflo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bfhi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

__device__ __forceinline__ float ex2_ftz(float x) {            // arguments <= 0 here: results in [0, 1], denormals flushed
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct MaxIdx { float v; int i; };
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

// block-wide (max, lowest index of it) and sum, through one shared-memory round each
__device__ __forceinline__ MaxIdx block_max(MaxIdx m, MaxIdx *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxIdx t;
        t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
        t.i = __shfl_xor_sync(0xffffffffu, m.i, o);
        m = better(m, t);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = m;
    __syncthreads();
    MaxIdx r = sh[0];
#pragma unroll
    for (int k = 1; k < CE_THREADS / 32; k++) r = better(r, sh[k]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_sum(float s, float *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = s;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < CE_THREADS / 32; k++) r += sh[k];          // fixed order
    __syncthreads();
    return r;
}

// thread t owns the vectors t, t + 512, ... (coalesced 8 KB segments; indices increase inside a thread)
__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(int V, const bf16 *__restrict__ logits, const int64_t *__restrict__ targets,
                                                            long long ignore, float *__restrict__ row_loss, float *__restrict__ row_lse,
                                                            float *__restrict__ row_max, int *__restrict__ row_argmax) {
    __shared__ MaxIdx sh_m[CE_THREADS / 32];
    __shared__ float sh_s[CE_THREADS / 32];
    const long long row = blockIdx.x;
    const uint4 *x = reinterpret_cast<const uint4 *>(logits + row * V);
    const int nvec = V / 8;
    // all loads first (unconditional code: a guarded load inside the loop that consumes it is not hoisted, and the 16
    // round trips to memory serialise -- measured 184 us for 268 MB); vectors past the row are -inf
    uint4 c[CE_MAXV];
    const uint4 ninf = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
#pragma unroll
    for (int j = 0; j < CE_MAXV; j++) {
        const int vi = j * CE_THREADS + threadIdx.x;
        c[j] = vi < nvec ? x[vi] : ninf;
    }
    // maximum: packed bf16 max over the thread's vectors, then one unpack
    __nv_bfloat162 pm = *reinterpret_cast<const __nv_bfloat162 *>(&ninf.x);
#pragma unroll
    for (int j = 0; j < CE_MAXV; j++) {
        const uint32_t w[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
        for (int e = 0; e < 4; e++) pm = __hmax2(pm, *reinterpret_cast<const __nv_bfloat162 *>(&w[e]));
    }
    const float tmax = fmaxf(__low2float(pm), __high2float(pm));
    // its lowest index: only a thread that holds the row maximum looks for it
    MaxIdx m = {tmax, 0x7fffffff};
    m = block_max(m, sh_m);
    if (tmax == m.v) {
        int first = 0x7fffffff;
#pragma unroll
        for (int j = CE_MAXV - 1; j >= 0; j--) {
            const uint32_t w[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
            const int base = (j * CE_THREADS + threadIdx.x) * 8;
#pragma unroll
            for (int e = 3; e >= 0; e--) {
                if (bfhi(w[e]) == tmax) first = base + 2 * e + 1;
                if (bflo(w[e]) == tmax) first = base + 2 * e;
            }
        }
        m.i = first;
    }
    m = block_max(m, sh_m);
    const float ms = m.v * LOG2E_F;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < CE_MAXV; j++) {
        const uint32_t w[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            s0 += ex2_ftz(fmaf(bflo(w[e]), LOG2E_F, -ms));
            s1 += ex2_ftz(fmaf(bfhi(w[e]), LOG2E_F, -ms));
        }
    }
    float s = s0 + s1;
    s = block_sum(s, sh_s);
    if (threadIdx.x == 0) {
        const float lse = m.v + log2f(s) * LN2_F;
        const long long t = targets[row];
        float loss = 0.f;
        if (t != ignore && t >= 0 && t < V) loss = lse - __bfloat162float(logits[row * V + t]);
        row_loss[row] = loss;
        row_lse[row] = lse;
        row_max[row] = m.v;
        row_argmax[row] = m.i;
    }
}

// out[0] = sum(row_loss) / n_valid, out[1] = n_valid  (one block, fixed order)
__global__ void __launch_bounds__(1024) ce_mean_kernel(long long rows, const float *__restrict__ row_loss,
                                                       const int64_t *__restrict__ targets, long long ignore, float *__restrict__ out) {
    __shared__ float sh_l[32], sh_n[32];
    float l = 0.f, n = 0.f;
    for (long long r = threadIdx.x; r < rows; r += 1024) {
        l += row_loss[r];
        n += targets[r] != ignore ? 1.f : 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l += __shfl_xor_sync(0xffffffffu, l, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0) { sh_l[threadIdx.x >> 5] = l; sh_n[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tl = 0.f, tn = 0.f;
        for (int k = 0; k < 32; k++) { tl += sh_l[k]; tn += sh_n[k]; }
        out[0] = tl / tn;
        out[1] = tn;
    }
}

__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(int V, const bf16 *__restrict__ logits, const int64_t *__restrict__ targets,
                                                            long long ignore, const float *__restrict__ row_lse,
                                                            const float *__restrict__ row_max, const int *__restrict__ row_argmax,
                                                            const float *__restrict__ mean_out, const float *__restrict__ gloss,
                                                            float l2_factor, bf16 *__restrict__ glogits) {
    const long long row = blockIdx.x;
    const uint4 *x = reinterpret_cast<const uint4 *>(logits + row * V);
    uint4 *gx = reinterpret_cast<uint4 *>(glogits + row * V);
    const int nvec = V / 8;
    const long long t = targets[row];
    const bool valid = t != ignore;
    const float scale = valid ? gloss[0] / mean_out[1] : 0.f;
    const float ls = row_lse[row] * LOG2E_F;
    const int tgt = valid ? (int)t : -1, amax = row_argmax[row];
    const float l2 = row_max[row] * l2_factor;
    for (int vi = threadIdx.x; vi < nvec; vi += CE_THREADS) {
        const uint4 c = x[vi];
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
        uint32_t o[4];
        const int j0 = vi * 8;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float a = exp2f(fmaf(bflo(w[e]), LOG2E_F, -ls)) * scale, b = exp2f(fmaf(bfhi(w[e]), LOG2E_F, -ls)) * scale;
            const int ja = j0 + 2 * e, jb = ja + 1;
            if (ja == tgt) a -= scale;
            if (jb == tgt) b -= scale;
            if (ja == amax) a += l2;
            if (jb == amax) b += l2;
            __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
            o[e] = *reinterpret_cast<uint32_t *>(&p);
        }
        gx[vi] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace
}  // namespace wkv6

using namespace wkv6;

extern "C" {

int cross_entropy_l2wrap_bf16(long long rows, int V, const void *logits, const int64_t *targets, long long ignore_index,
                              float *row_stats, int *row_argmax, float *mean_out, void *stream) {
    if (rows < 0 || V <= 0 || (V % 8) || V > CE_THREADS * CE_MAXV * 8) {
        set_error("cross_entropy_l2wrap_bf16: need V %% 8 == 0 and V <= 65536");
        return WKV6_EINVAL;
    }
    if (rows == 0) return WKV6_OK;
    if (!logits || !targets || !row_stats || !row_argmax || !mean_out) { set_error("cross_entropy_l2wrap_bf16: null pointer"); return WKV6_EINVAL; }
    auto s = (cudaStream_t)stream;
    ce_fwd_kernel<<<(unsigned)rows, CE_THREADS, 0, s>>>(V, (const bf16 *)logits, targets, ignore_index, row_stats, row_stats + rows,
                                                         row_stats + 2 * rows, row_argmax);
    ce_mean_kernel<<<1, 1024, 0, s>>>(rows, row_stats, targets, ignore_index, mean_out);
    count_launch(2);
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

int cross_entropy_l2wrap_backward_bf16(long long rows, int V, const void *logits, const int64_t *targets, long long ignore_index,
                                       const float *row_stats, const int *row_argmax, const float *mean_out, const float *gloss,
                                       float l2_factor, void *glogits, void *stream) {
    if (rows < 0 || V <= 0 || (V % 8)) { set_error("cross_entropy_l2wrap_backward_bf16: need V %% 8 == 0"); return WKV6_EINVAL; }
    if (rows == 0) return WKV6_OK;
    if (!logits || !targets || !row_stats || !row_argmax || !mean_out || !gloss || !glogits) {
        set_error("cross_entropy_l2wrap_backward_bf16: null pointer");
        return WKV6_EINVAL;
    }
    ce_bwd_kernel<<<(unsigned)rows, CE_THREADS, 0, (cudaStream_t)stream>>>(V, (const bf16 *)logits, targets, ignore_index, row_stats + rows,
                                                                           row_stats + 2 * rows, row_argmax, mean_out, gloss, l2_factor,
                                                                           (bf16 *)glogits);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

}  // extern "C"
