// Time-axis segmentation for calls with few (b,h) streams.
//
// The chunked kernels map one stream to one CTA, so a call with B*H < 148 streams (infctx training
// with B = 1, inference prefill, small SFT buckets) leaves most SMs idle.  The recurrence is linear in
// the state, so a sequence can be cut into `nseg` segments that run as independent "batch rows":
//     S_end(seg) = diag(exp(Lam_seg)) * S_start(seg) + S_loc(seg)
// with S_loc the final state of the segment run from a ZERO state and Lam_seg the segment's total
// log-decay.  Forward = state-only pass over all segments (S_loc) + the two small kernels below
// (Lam, scan over segments) + the ordinary forward over all segments with S_start as initial states.
// [B,T,C] viewed as [B*nseg, T/nseg, C] is the same memory, so the big kernels run unchanged.
#include "common.cuh"
#include <cstdlib>

namespace wkv6 {
namespace {

typedef __nv_bfloat16 bf16;

// lam[row, c] = sum over the segment's tokens of -exp(w[b, t, c]);  row = b*nseg + seg covers tokens
// [seg*seg_tokens, min(T, (seg+1)*seg_tokens)) of sequence b.
// grid (ceil(C/64), rows), block 256 = 8 column lanes (8 channels each: 128 contiguous bytes of a token row)
// x 32 token lanes; every thread keeps four independent 16-byte loads in flight.  Deterministic.
__global__ void __launch_bounds__(256) seg_decay_kernel(int T, int C, int nseg, int seg_tokens, const bf16 *__restrict__ w,
                                                        float *__restrict__ lam, float lmin) {
    const int cl = threadIdx.x & 7, tl = threadIdx.x >> 3;
    const int c = (blockIdx.x * 8 + cl) * 8;
    const size_t row = blockIdx.y;
    const int b = (int)(row / nseg), t0 = (int)(row % nseg) * seg_tokens;
    const int Tseg = min(seg_tokens, T - t0);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < C) {
        const bf16 *base = w + ((size_t)b * T + t0) * C + c;
        for (int t = tl; t < Tseg; t += 128) {
            uint4 u[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                u[j] = t + 32 * j < Tseg ? *reinterpret_cast<const uint4 *>(base + (size_t)(t + 32 * j) * C) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (t + 32 * j >= Tseg) break;
                const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u[j]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 f = __bfloat1622float2(h[i]);
                    acc[2 * i] += fmaxf(-__expf(f.x), lmin);
                    acc[2 * i + 1] += fmaxf(-__expf(f.y), lmin);
                }
            }
        }
    }
    __shared__ float red[32][64 + 1];
#pragma unroll
    for (int e = 0; e < 8; e++) red[tl][cl * 8 + e] = acc[e];
    __syncthreads();
    if (threadIdx.x < 64) {
        const int cc = blockIdx.x * 64 + threadIdx.x;
        if (cc < C) {
            float s = 0.f;
#pragma unroll 8
            for (int l = 0; l < 32; l++) s += red[l][threadIdx.x];
            lam[row * C + cc] = s;
        }
    }
}

// Scan over the segments of one (b,h) stream, element-wise on the 64x64 state (caller layout
// [value j][key i], the decay acts on the key index = fastest dimension).
//   S_start[b,0] = s0 (or 0);  S_start[b,seg+1] = exp(lam[b,seg,h,i]) * S_start[b,seg] + S_loc[b,seg]
// reverse = 1 walks the segments from the last to the first (the backward's state-gradient chain).
// grid (16, B*H), block 256: one state element per thread; the loads do not depend on the running value and
// are issued four segments ahead.
__global__ void __launch_bounds__(256) seg_scan_kernel(int nseg, int H, const float *__restrict__ lam,
                                                       const float *__restrict__ s_loc, const void *__restrict__ s0,
                                                       int s0_f32, long long s0_bstride, float *__restrict__ s_start,
                                                       void *__restrict__ sT, int sT_f32, int reverse,
                                                       const int *__restrict__ stream_flags) {
    const int stream = blockIdx.y, b = stream / H, h = stream % H;
    // a flagged stream is recomputed by the exact route from the ORIGINAL initial state, which sT may alias
    const bool write_sT = sT != nullptr && !(stream_flags && stream_flags[stream] != 0);
    const size_t C = (size_t)H * 64;
    const int e = blockIdx.x * 256 + threadIdx.x, i = e & 63;
    float S = 0.f;
    if (s0) {
        const size_t idx = (size_t)b * s0_bstride + (size_t)h * 4096 + e;
        S = s0_f32 ? ((const float *)s0)[idx] : __bfloat162float(((const bf16 *)s0)[idx]);
    }
    for (int q0 = 0; q0 < nseg; q0 += 4) {
        float l[4], x[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int q = q0 + j;
            if (q < nseg) {
                const size_t row = (size_t)b * nseg + (reverse ? nseg - 1 - q : q);
                l[j] = lam[row * C + h * 64 + i];
                x[j] = s_loc[(row * H + h) * 4096 + e];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int q = q0 + j;
            if (q < nseg) {
                const size_t row = (size_t)b * nseg + (reverse ? nseg - 1 - q : q);
                s_start[(row * H + h) * 4096 + e] = S;
                S = __expf(l[j]) * S + x[j];
            }
        }
    }
    if (write_sT) {
        const size_t idx = ((size_t)b * H + h) * 4096 + e;
        if (sT_f32) ((float *)sT)[idx] = S;
        else ((bf16 *)sT)[idx] = __float2bfloat16_rn(S);
    }
}

// a stream is flagged when any of its segments is; then all its segments are (so later segmented
// launches skip them and the exact route recomputes the whole stream)
__global__ void seg_flags_kernel(int n, int nseg, int H, int *__restrict__ seg_flags, int *__restrict__ stream_flags) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;   // b*H + h
    if (s >= n) return;
    const int b = s / H, h = s % H;
    int f = stream_flags[s];
    for (int q = 0; q < nseg; q++) f |= seg_flags[((size_t)b * nseg + q) * H + h];
    f = f != 0;
    stream_flags[s] = f;
    for (int q = 0; q < nseg; q++) seg_flags[((size_t)b * nseg + q) * H + h] = f;
}

// Reverse the token order inside every segment of three [B,T,C] tensors at once (the backward's
// state-gradient chain is the forward state recurrence on time-reversed r, gy, w): token t of segment
// [t0, t1) takes the row of token t0 + t1 - 1 - t.
__global__ void __launch_bounds__(256) seg_reverse3_kernel(size_t nvec, int T, int C, int seg_tokens, const bf16 *__restrict__ a,
                                                           const bf16 *__restrict__ b, const bf16 *__restrict__ c,
                                                           bf16 *__restrict__ ra, bf16 *__restrict__ rb_, bf16 *__restrict__ rc) {
    const int cv = C / 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        const size_t tok = i / cv, bb = tok / T;
        const int t = (int)(tok % T);
        const int t0 = t / seg_tokens * seg_tokens, t1 = min(T, t0 + seg_tokens);
        const size_t src = ((bb * T + (size_t)(t0 + t1 - 1 - t)) * cv + i % cv) * 8;
        *reinterpret_cast<uint4 *>(ra + i * 8) = *reinterpret_cast<const uint4 *>(a + src);
        *reinterpret_cast<uint4 *>(rb_ + i * 8) = *reinterpret_cast<const uint4 *>(b + src);
        *reinterpret_cast<uint4 *>(rc + i * 8) = *reinterpret_cast<const uint4 *>(c + src);
    }
}

// gu[b, c] = sum_seg part[b*nseg + seg, c]   (bf16 in, fp32 sum, bf16 out)
__global__ void seg_sum_gu_kernel(int B, int nseg, int C, const bf16 *__restrict__ part, bf16 *__restrict__ gu) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i % C;
    float s = 0.f;
    for (int q = 0; q < nseg; q++) s += __bfloat162float(part[((size_t)b * nseg + q) * C + c]);
    gu[i] = __float2bfloat16_rn(s);
}

}  // namespace

// stream-ordered scratch (cudaMallocAsync) stays cached in the device's default pool instead of going
// back to the OS at every synchronisation
void ensure_pool_keeps_memory() {
    static bool pool_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !pool_set[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool_set[dev] = true;
    }
}

// Segment plan: nseg segments of seg_chunks 64-token chunks each (the last one may be shorter), enough of them
// to fill the 296 CTA slots, none empty.  nseg = 1: do not segment.
static void plan(long long streams, int T, int min_chunks, int *nseg, int *seg_chunks) {
    const int NC = (T + 63) / 64;
    int n = (int)(296 / streams);
    if (n > NC / min_chunks) n = NC / min_chunks;
    if (n < 2) { *nseg = 1; *seg_chunks = NC; return; }
    const int sc = (NC + n - 1) / n;
    *seg_chunks = sc;
    *nseg = (NC + sc - 1) / sc;
    if (*nseg < 2) { *nseg = 1; *seg_chunks = NC; }
}
// forward-only calls: worth the extra launches from T = 8192 on (at least two chunks per segment)
void seg_plan(int B, int T, int H, int *nseg, int *seg_chunks) {
    const long long streams = (long long)B * H;
    static const bool off = getenv("WKV6B200_NO_SEG") != nullptr;     // A/B switch for profiles/bench_few_streams*.py
    *nseg = 1; *seg_chunks = (T + 63) / 64;
    if (off || streams <= 0 || streams >= 148 || T < 4096) return;
    plan(streams, T, 2, nseg, seg_chunks);
}
// training pair: forward and backward are both segmented, worth it from T = 2048 on; at least 4 chunks per segment
void seg_plan_train(int B, int T, int H, int *nseg, int *seg_chunks) {
    const long long streams = (long long)B * H;
    static const bool off = getenv("WKV6B200_NO_SEG") != nullptr;
    *nseg = 1; *seg_chunks = (T + 63) / 64;
    if (off || streams <= 0 || streams > 74 || T < 2048) return;
    plan(streams, T, 4, nseg, seg_chunks);
}

int seg_reverse3(int B, int T, int C, int nseg, int seg_tokens, const void *a, const void *b, const void *c, void *ra, void *rb,
                 void *rc, cudaStream_t stream) {
    (void)nseg;
    const size_t nvec = (size_t)B * T * C / 8;
    size_t g = (nvec + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    seg_reverse3_kernel<<<(int)g, 256, 0, stream>>>(nvec, T, C, seg_tokens, (const bf16 *)a, (const bf16 *)b, (const bf16 *)c, (bf16 *)ra,
                                                    (bf16 *)rb, (bf16 *)rc);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

int seg_sum_gu(int B, int nseg, int C, const void *part, void *gu, cudaStream_t stream) {
    seg_sum_gu_kernel<<<(B * C + 255) / 256, 256, 0, stream>>>(B, nseg, C, (const bf16 *)part, (bf16 *)gu);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

int seg_decay(int B, int T, int C, int nseg, int seg_tokens, const void *w, float *lam, float lmin, cudaStream_t stream) {
    dim3 grid((C + 63) / 64, B * nseg);
    seg_decay_kernel<<<grid, 256, 0, stream>>>(T, C, nseg, seg_tokens, (const bf16 *)w, lam, lmin);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

int seg_scan(int B, int nseg, int H, const float *lam, const float *s_loc, const void *s0, int s0_f32,
             long long s0_bstride, float *s_start, void *sT, int sT_f32, int reverse, const int *stream_flags,
             cudaStream_t stream) {
    seg_scan_kernel<<<dim3(16, B * H), 256, 0, stream>>>(nseg, H, lam, s_loc, s0, s0_f32, s0_bstride, s_start, sT, sT_f32, reverse,
                                               stream_flags);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

int seg_flags_merge(int B, int nseg, int H, int *seg_flags, int *stream_flags, cudaStream_t stream) {
    const int n = B * H;
    if (n == 0) return WKV6_OK;
    // n < 148 by construction: one small block
    seg_flags_kernel<<<(n + 127) / 128, 128, 0, stream>>>(n, nseg, H, seg_flags, stream_flags);
    count_launch();
    WKV6_CUDA_CHECK(cudaGetLastError());
    return WKV6_OK;
}

}  // namespace wkv6
