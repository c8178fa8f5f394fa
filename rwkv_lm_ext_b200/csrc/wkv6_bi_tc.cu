// wkv6_bi forward on the tensor-core kernels (cuda/wkv6_bi_cuda.cu:7-112, semantics 3 of SURVEY.md 2.3).
//
//   y_t = r_t^T (diag(u) kv_t + S_t) + r_t^T S'_t   for t <= p,   0 for t > p        (p = first masked token, or T-1)
//   S'_{t-1} = diag(d_t) S'_t + kv_t,  S'_p = 0                                       (reverse pass: exclusive, no u)
//
// In reversed time tau = p - t the reverse pass IS the causal recurrence with u = 0 (Z_{tau+1} = d Z_tau + kv,
// y'_tau = r^T Z_tau), so the op is two launches of the chunked forward kernel -- one on the tokens as they
// are, one on each row's first p+1 tokens reversed -- between a vectorised reverse-gather of r,k,v,w and a
// combine pass that un-reverses, adds (bf16 accumulation across the two passes, like the reference) and
// zeroes the tail.  Streams the kernels flag (decay too strong for the block references) are recomputed by
// the exact SIMT bidirectional kernels, predicated per stream.
#include "common.cuh"

namespace wkv6 {
namespace {

typedef __nv_bfloat16 bf16;

__global__ void __launch_bounds__(256) bi_last_kernel(int T, const int *__restrict__ mask, int *__restrict__ p) {
    const int *row = mask + (size_t)blockIdx.x * T;
    __shared__ int best;
    if (threadIdx.x == 0) best = T;
    __syncthreads();
    int mine = T;
    for (int t = threadIdx.x; t < T; t += blockDim.x)
        if (row[t] == 0) { mine = t; break; }
    atomicMin(&best, mine);
    __syncthreads();
    if (threadIdx.x == 0) p[blockIdx.x] = best >= T ? T - 1 : best;
}

// one block per (b, t): out_x[b, t, :] = x[b, src, :],  src = t <= p ? p - t : t,  x in {r, k, v, w}
__global__ void __launch_bounds__(256) bi_reverse4_kernel(int T, int C, const int *__restrict__ p, const bf16 *__restrict__ r,
                                                          const bf16 *__restrict__ k, const bf16 *__restrict__ v,
                                                          const bf16 *__restrict__ w, bf16 *__restrict__ ro, bf16 *__restrict__ ko,
                                                          bf16 *__restrict__ vo, bf16 *__restrict__ wo) {
    const int b = blockIdx.x / T, t = blockIdx.x % T;
    const int pb = p[b], src = t <= pb ? pb - t : t;
    const size_t so = ((size_t)b * T + src) * C, doff = (size_t)blockIdx.x * C;
    for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
        const uint4 a0 = *reinterpret_cast<const uint4 *>(r + so + c), a1 = *reinterpret_cast<const uint4 *>(k + so + c);
        const uint4 a2 = *reinterpret_cast<const uint4 *>(v + so + c), a3 = *reinterpret_cast<const uint4 *>(w + so + c);
        *reinterpret_cast<uint4 *>(ro + doff + c) = a0;
        *reinterpret_cast<uint4 *>(ko + doff + c) = a1;
        *reinterpret_cast<uint4 *>(vo + doff + c) = a2;
        *reinterpret_cast<uint4 *>(wo + doff + c) = a3;
    }
}

// y[b,t,:] = t <= p ? bf16(float(y[b,t,:]) + float(y2[b,p-t,:])) : 0
__global__ void __launch_bounds__(256) bi_combine_kernel(int T, int C, const int *__restrict__ p, bf16 *__restrict__ y,
                                                         const bf16 *__restrict__ y2) {
    const int b = blockIdx.x / T, t = blockIdx.x % T;
    const int pb = p[b];
    bf16 *dst = y + (size_t)blockIdx.x * C;
    if (t > pb) {
        for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) *reinterpret_cast<uint4 *>(dst + c) = make_uint4(0, 0, 0, 0);
        return;
    }
    const bf16 *src = y2 + ((size_t)b * T + (pb - t)) * C;
    for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
        uint4 a = *reinterpret_cast<const uint4 *>(dst + c);
        const uint4 bb = *reinterpret_cast<const uint4 *>(src + c);
        __nv_bfloat162 *pa = reinterpret_cast<__nv_bfloat162 *>(&a);
        const __nv_bfloat162 *pc = reinterpret_cast<const __nv_bfloat162 *>(&bb);
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float2 x = __bfloat1622float2(pa[e]), z = __bfloat1622float2(pc[e]);
            pa[e] = __floats2bfloat162_rn(x.x + z.x, x.y + z.y);
        }
        *reinterpret_cast<uint4 *>(dst + c) = a;
    }
}

// masked / reversed output gradient: gm[b,t] = t <= p ? gy[b,t] : 0;  gr[b,t] = t <= p ? gy[b,p-t] : 0
__global__ void __launch_bounds__(256) bi_gy_kernel(int T, int C, const int *__restrict__ p, const bf16 *__restrict__ gy,
                                                    bf16 *__restrict__ gm, bf16 *__restrict__ grev) {
    const int b = blockIdx.x / T, t = blockIdx.x % T;
    const int pb = p[b];
    const size_t doff = (size_t)blockIdx.x * C, so = ((size_t)b * T + (t <= pb ? pb - t : t)) * C;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
        *reinterpret_cast<uint4 *>(gm + doff + c) = t <= pb ? *reinterpret_cast<const uint4 *>(gy + doff + c) : z;
        *reinterpret_cast<uint4 *>(grev + doff + c) = t <= pb ? *reinterpret_cast<const uint4 *>(gy + so + c) : z;
    }
}

// g[b,t,:] += g2[b,p-t,:] for t <= p, for the four gradients at once (bf16 accumulation like the forward)
__global__ void __launch_bounds__(256) bi_combine4_kernel(int T, int C, const int *__restrict__ p, bf16 *__restrict__ g0,
                                                          bf16 *__restrict__ g1, bf16 *__restrict__ g2, bf16 *__restrict__ g3,
                                                          const bf16 *__restrict__ h0, const bf16 *__restrict__ h1,
                                                          const bf16 *__restrict__ h2, const bf16 *__restrict__ h3) {
    const int b = blockIdx.x / T, t = blockIdx.x % T;
    const int pb = p[b];
    if (t > pb) return;
    const size_t doff = (size_t)blockIdx.x * C, so = ((size_t)b * T + (pb - t)) * C;
    bf16 *g[4] = {g0, g1, g2, g3};
    const bf16 *h[4] = {h0, h1, h2, h3};
    for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8)
#pragma unroll
        for (int x = 0; x < 4; x++) {
            uint4 a = *reinterpret_cast<const uint4 *>(g[x] + doff + c);
            const uint4 bb = *reinterpret_cast<const uint4 *>(h[x] + so + c);
            __nv_bfloat162 *pa = reinterpret_cast<__nv_bfloat162 *>(&a);
            const __nv_bfloat162 *pc = reinterpret_cast<const __nv_bfloat162 *>(&bb);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 u = __bfloat1622float2(pa[e]), z = __bfloat1622float2(pc[e]);
                pa[e] = __floats2bfloat162_rn(u.x + z.x, u.y + z.y);
            }
            *reinterpret_cast<uint4 *>(g[x] + doff + c) = a;
        }
}

}  // namespace

bool bi_forward_tc_supported(const Args &a) {
    Args t = a;
    t.mask = nullptr;
    if (t.w_kind == W_LOG_F32) t.w_kind = W_RAW_BF16;
    return a.mask != nullptr && a.s0 == nullptr && a.sT == nullptr && tc3_forward_supported(t);
}

// flags: device int [B*H], zeroed by the caller
int bi_forward_tc(const Args &a, int *flags) {
    if (a.B * a.H == 0 || a.T == 0) return WKV6_OK;
    const int C = a.H * N;
    const size_t n = (size_t)a.B * a.T * C * sizeof(bf16);
    const bool convert = a.w_kind == W_LOG_F32;
    // stream-ordered scratch: reversed r,k,v,w, y of the reverse pass, (raw logits), zero u, p
    const size_t total = (5 + (convert ? 1 : 0)) * n + (size_t)a.H * N * sizeof(bf16) + (size_t)a.B * sizeof(int) + 256;
    uint8_t *sc = nullptr;
    WKV6_CUDA_CHECK(cudaMallocAsync((void **)&sc, total, a.stream));
    bf16 *rr = (bf16 *)sc, *kr = (bf16 *)(sc + n), *vr = (bf16 *)(sc + 2 * n), *wr = (bf16 *)(sc + 3 * n), *y2 = (bf16 *)(sc + 4 * n);
    uint8_t *q = sc + 5 * n;
    const void *w_raw = a.w;
    int rc = WKV6_OK;
    if (convert) {
        rc = ew_to_raw_bf16(a.B, a.T, a.H, (const float *)a.w, q, flags, a.stream);
        w_raw = q;
        q += n;
    }
    bf16 *u0 = (bf16 *)q;
    int *p = (int *)(q + ((size_t)a.H * N * sizeof(bf16) + 255) / 256 * 256);
    if (rc == WKV6_OK && cudaMemsetAsync(u0, 0, (size_t)a.H * N * sizeof(bf16), a.stream) != cudaSuccess) rc = WKV6_ECUDA;
    if (rc == WKV6_OK) {
        bi_last_kernel<<<a.B, 256, 0, a.stream>>>(a.T, a.mask, p);
        bi_reverse4_kernel<<<a.B * a.T, 256, 0, a.stream>>>(a.T, C, p, (const bf16 *)a.r, (const bf16 *)a.k, (const bf16 *)a.v,
                                                            (const bf16 *)w_raw, rr, kr, vr, wr);
        count_launch(2);
        if (cudaGetLastError() != cudaSuccess) { set_error("wkv6_bi gather launch failed"); rc = WKV6_ECUDA; }
    }
    if (rc == WKV6_OK) {
        Args f = a;                       // causal pass on the tokens as they are
        f.mask = nullptr; f.w = w_raw; f.w_kind = W_RAW_BF16;
        rc = tc3_forward(f, nullptr, flags);
        if (rc == WKV6_OK) {              // reverse pass: reversed tokens, u = 0
            f.r = rr; f.k = kr; f.v = vr; f.w = wr; f.u = u0; f.y = y2;
            rc = tc3_forward(f, nullptr, flags);
        }
    }
    if (rc == WKV6_OK) {
        bi_combine_kernel<<<a.B * a.T, 256, 0, a.stream>>>(a.T, C, p, (bf16 *)a.y, y2);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { set_error("wkv6_bi combine launch failed"); rc = WKV6_ECUDA; }
    }
    if (rc == WKV6_OK) {                  // exact route for the flagged streams (both passes), from the caller's own tensors
        Args s = a;
        s.stream_flags = flags;
        rc = simt_forward(s);
    }
    cudaFreeAsync(sc, a.stream);
    return rc;
}

// wkv6_bi backward: the gradient of the two-pass forward above is two runs of the chunked backward kernel
// (pass 1 on the tokens as they are with gy masked beyond p; pass 2 on the reversed tokens with the reversed
// gy and u = 0), the second un-reversed and added.  a.workspace must hold wkv6_backward_workspace_bytes.
int bi_backward_tc(const Args &a) {
    if (a.B * a.H == 0 || a.T == 0) return WKV6_OK;
    const int C = a.H * N;
    const size_t n = (size_t)a.B * a.T * C * sizeof(bf16);
    const bool convert = a.w_kind == W_LOG_F32;
    const size_t base = tc3_backward_workspace_bytes(a.B, a.T, a.H, false);
    if (!a.workspace || a.workspace_bytes < base) { set_error("workspace too small: need %zu bytes", base); return WKV6_EWORKSPACE; }
    const size_t nflag = (size_t)a.B * a.H * sizeof(int);
    const size_t small = ((size_t)a.H * N * sizeof(bf16) + (size_t)a.B * C * sizeof(bf16) + (size_t)a.B * sizeof(int) + nflag + 1023) / 256 * 256;
    const size_t total = (10 + (convert ? 1 : 0)) * n + small;
    uint8_t *sc = nullptr;
    WKV6_CUDA_CHECK(cudaMallocAsync((void **)&sc, total, a.stream));
    bf16 *t[10];
    for (int i = 0; i < 10; i++) t[i] = (bf16 *)(sc + i * n);       // r',k',v',w', gy_m, gy_r, gr2,gk2,gv2,gw2
    uint8_t *q = sc + 10 * n;
    const void *w_raw = a.w;
    if (convert) { w_raw = q; q += n; }
    bf16 *u0 = (bf16 *)q;
    bf16 *gu2 = (bf16 *)(q + (size_t)a.H * N * sizeof(bf16));
    int *p = (int *)((uint8_t *)gu2 + (size_t)a.B * C * sizeof(bf16));
    int *allflags = p + ((a.B + 63) / 64) * 64;
    int *wsflags = (int *)((uint8_t *)a.workspace + simt_backward_workspace_bytes(a.B, a.T, a.H));   // where tc3_backward keeps them
    int rc = WKV6_OK;
    auto ck = [&](cudaError_t e) { if (e != cudaSuccess && rc == WKV6_OK) { set_error("wkv6_bi backward: %s", cudaGetErrorString(e)); rc = WKV6_ECUDA; } };
    ck(cudaMemsetAsync(u0, 0, (size_t)a.H * N * sizeof(bf16), a.stream));
    ck(cudaMemsetAsync(allflags, 0, nflag, a.stream));
    ck(cudaMemsetAsync(wsflags, 0, nflag, a.stream));
    if (rc == WKV6_OK && convert) rc = ew_to_raw_bf16(a.B, a.T, a.H, (const float *)a.w, const_cast<void *>(w_raw), wsflags, a.stream);
    if (rc == WKV6_OK) {
        bi_last_kernel<<<a.B, 256, 0, a.stream>>>(a.T, a.mask, p);
        bi_reverse4_kernel<<<a.B * a.T, 256, 0, a.stream>>>(a.T, C, p, (const bf16 *)a.r, (const bf16 *)a.k, (const bf16 *)a.v,
                                                            (const bf16 *)w_raw, t[0], t[1], t[2], t[3]);
        bi_gy_kernel<<<a.B * a.T, 256, 0, a.stream>>>(a.T, C, p, (const bf16 *)a.gy, t[4], t[5]);
        count_launch(3);
        ck(cudaGetLastError());
    }
    // streams flagged in pass 1 are skipped by pass 2 (same flag words) and redone below by the exact kernels
    if (rc == WKV6_OK) {
        Args b1 = a;
        b1.mask = nullptr; b1.w = w_raw; b1.w_kind = W_RAW_BF16; b1.gy = t[4]; b1.workspace_bytes = base; b1.stream_flags = nullptr;
        rc = tc3_backward(b1, nullptr, true, /*run_fallback=*/false);
    }
    if (rc == WKV6_OK) {
        Args b2 = a;
        b2.mask = nullptr; b2.w_kind = W_RAW_BF16; b2.r = t[0]; b2.k = t[1]; b2.v = t[2]; b2.w = t[3]; b2.u = u0; b2.gy = t[5];
        b2.gr = t[6]; b2.gk = t[7]; b2.gv = t[8]; b2.gw = t[9]; b2.gu = gu2; b2.workspace_bytes = base;
        rc = tc3_backward(b2, nullptr, true, /*run_fallback=*/false);
    }
    if (rc == WKV6_OK) {
        bi_combine4_kernel<<<a.B * a.T, 256, 0, a.stream>>>(a.T, C, p, (bf16 *)a.gr, (bf16 *)a.gk, (bf16 *)a.gv, (bf16 *)a.gw,
                                                            t[6], t[7], t[8], t[9]);
        count_launch();
        ck(cudaGetLastError());
        ck(cudaMemcpyAsync(allflags, wsflags, nflag, cudaMemcpyDeviceToDevice, a.stream));
    }
    if (rc == WKV6_OK) {          // exact bidirectional backward for the flagged streams, from the caller's own tensors
        Args s = a;
        s.stream_flags = allflags;
        s.workspace_bytes = simt_backward_workspace_bytes(a.B, a.T, a.H);
        rc = simt_backward(s);
    }
    cudaFreeAsync(sc, a.stream);
    return rc;
}

}  // namespace wkv6
