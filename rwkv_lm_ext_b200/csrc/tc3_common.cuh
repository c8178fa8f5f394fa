// Shared pieces of the role-uniform tcgen05 kernels (wkv6_tc3_fwd.cu / wkv6_tc3_bwd.cu).
//
// Thread <-> data mapping ("fragment mapping"), used by EVERY stage of both kernels.  A CTA has 8
// compute warps; warp w = (sp = w % 4, ch = w / 4), lane T = (ri = T / 4, q = T % 4).  Of a 64 x 64
// tile the thread owns rows  R_h = 16 sp + 8 h + ri  (h = 0,1)  and columns  c(g,e) = 32 ch + 8 g +
// 2 q + e  (g = 0..3, e = 0,1): 16 elements, register index [4g + 2h + e] as tcgen05.ld.16x256b.x4
// delivers an M = 64 accumulator (sub-partition sp, lanes 0-15).  For the operand preparation the
// same thread owns CHANNELS i = R_h and TOKENS t = c(g,e): ldmatrix.trans / stmatrix.trans move
// its bf16 pairs between that mapping and the [token][channel] shared-memory tiles, so the
// per-channel quantities (decay prefix sums, references, state rows) never change threads between
// the preparation, the TMEM epilogues and the output stage.
#pragma once
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace wkv6 {
namespace tc3 {

using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int L = 64;
constexpr int CWARPS = 8, CTHREADS = 256, NTHREADS = CTHREADS + 32;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr float LOG2_LOG2E = 0.5287663729448977f;   // log2(log2(e)): exp(w) * log2(e) = 2^(w log2(e) + log2(log2(e)))
// Built-in floor of the per-token log2-decay.  A factor 2^-13 = 1.2e-4 per token is below bf16 resolution
// (2^-9), so the floor changes no output beyond the stated tolerance (checked against the UNCLAMPED fp64
// recurrence at w ~ N(0,1) and hotter, tests/test_gpu_parity.py), and it bounds every scaled operand: a 16-token
// block has its reference in the middle, so exponents stay within 8 x 13 = 104 < 127 binary orders.
constexpr float LCLAMP2 = 13.0f;
constexpr int B_SCAN_ID = 1;   // == B_SCAN below (named barrier of the 256 compute threads)

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
// (d == c) ? a : b as one compare + one select (the compiler turns chains of ?: over runtime ints into branches)
__device__ __forceinline__ float sel_eq(int d, int c, float a, float b) {
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\tselp.f32 %0, %3, %4, p;\n\t}" : "=f"(r) : "r"(d), "r"(c), "f"(a), "f"(b));
    return r;
}
// 2^n for an integer-valued n, exact; 0 below the normal range, clamped at 2^0 above
__device__ __forceinline__ float pow2i_le0(float n) {
    const int e = min(max((int)n, -127), 0);
    return __int_as_float((e + 127) << 23);
}
// 2^e for an integer e <= 0, exact; 0 below the normal range
__device__ __forceinline__ float pow2i(int e) { return __int_as_float((max(e, -127) + 127) << 23); }
// packed bf16 pair (x, x) of an exact power of two given as fp32
__device__ __forceinline__ uint32_t bfpair(float x) {
    const uint32_t hi = __float_as_uint(x) & 0xffff0000u;   // a power of two has no low mantissa bits
    return hi | (hi >> 16);
}
// exact scaling of a packed bf16 pair by a packed pair of powers of two
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 operations per issue slot).  A pair lives in an
// aligned 64-bit register; lo = the even element (e = 0), hi = the odd one.
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2pack(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 f2packu(uint32_t lo, uint32_t hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void f2unpack(f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void f2unpacku(f2 v, uint32_t &lo, uint32_t &hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2bcast(float x) { return f2pack(x, x); }
__device__ __forceinline__ f2 f2fma(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 f2mul(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2add(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2sub(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// packed bf16 pair -> packed fp32 pair (exact), and back (round to nearest even)
__device__ __forceinline__ f2 bf2f2(uint32_t x) { return f2packu(x << 16, x & 0xffff0000u); }
__device__ __forceinline__ uint32_t f2tobf(f2 v) {
    float lo, hi;
    f2unpack(v, lo, hi);
    return pack2(lo, hi);
}

struct Frag {
    int warp, lane, sp, ch, ri, q;
    uint32_t ti_off;     // byte offset (within a [token][channel] tile) this lane supplies to ldmatrix/stmatrix .x4:
                         // row 32ch + 8(lane/8) + lane%8, 16-byte chunk of channels 16sp (+ 16 bytes for h = 1)
    uint32_t ti1_off;    // same for .x1 (one 8-token group): row 32ch + lane%8; add 1024 per group, ^16 for h = 1
    uint32_t rc_off;     // byte offset for a plain stmatrix .x4 of an accumulator fragment into a [row][col] tile:
                         // row 16sp + lane%8 (+ 8 rows for h = 1), 16-byte chunk of columns 32ch + 8(lane/8)
    __device__ __forceinline__ void init() {
        warp = threadIdx.x >> 5;
        lane = threadIdx.x & 31;
        sp = warp & 3;
        ch = (warp >> 2) & 1;
        ri = lane >> 2;
        q = lane & 3;
        ti_off = sw128(32 * ch + 8 * (lane >> 3) + (lane & 7), 32 * sp);
        ti1_off = sw128(32 * ch + (lane & 7), 32 * sp);
        rc_off = sw128(16 * sp + (lane & 7), 64 * ch + 16 * (lane >> 3));
        // opaque to the optimiser: otherwise, under register pressure, ptxas re-derives these from SR_TID inside the
        // chunk loop (S2R + ~10 integer instructions per use, 6 % of all instructions executed) instead of keeping them
        asm volatile("" : "+r"(ti_off), "+r"(ti1_off), "+r"(rc_off));
    }
    // h = 1 moves the 16-byte chunk (ti) / the row by 8 (rc); both keep row % 8, so the swizzle XOR is unchanged
    __device__ __forceinline__ uint32_t ti(int h) const { return ti_off ^ (h ? 16u : 0u); }
    __device__ __forceinline__ uint32_t ti1(int g, int h) const { return (ti1_off + 1024u * g) ^ (h ? 16u : 0u); }
    __device__ __forceinline__ uint32_t rc(int h) const { return rc_off + (h ? 1024u : 0u); }
    __device__ __forceinline__ int row(int h) const { return 16 * sp + 8 * h + ri; }
    __device__ __forceinline__ int col(int g, int e) const { return 32 * ch + 8 * g + 2 * q + e; }
};

// ---- bidirectional op (wkv6_bi_tc.cu): the kernels take a mode BI.
//   BI_NONE   the ordinary call
//   BI_CAUSAL the causal direction of wkv6_bi: every batch row has its own length (row_len[b] = p + 1 tokens); tokens
//             past it are treated as absent and the outputs behind it are written as zeros
//   BI_REV    the reverse direction: the SAME tiles, read so that the kernel sees the row's first p + 1 tokens in
//             reversed order.  Chunk c covers reversed positions tau = 64c .. 64c + nv - 1 (nv = 64 but for the last
//             chunk), i.e. the tile that starts at token max(p + 1 - 64(c+1), 0) -- TMA stores do not take negative
//             coordinates -- and inside the tile row x holds tau_local = (nv - 1 - x) mod 64; the rows x >= nv of a
//             short last chunk (tokens the previous chunk already covered) are masked like the padding of a ragged
//             causal chunk.  ldmatrix / stmatrix take one row address per lane, so reading the raw r, k, w tiles and
//             writing the output tiles in that row order costs nothing; the tiles the tensor cores read as they
//             arrive (v, gy) are permuted in place by the compute warps first.  Outputs are ADDED to what the causal
//             pass stored (cp.reduce.async.bulk.tensor; masked rows add zeros), u = 0.
enum : int { BI_NONE = 0, BI_CAUSAL = 1, BI_REV = 2 };
// byte offset inside an 8 KB [64 rows][128 B] swizzled tile -> the same 16-byte chunk of row (r0 - row) mod 64
__device__ __forceinline__ uint32_t flip_rows(uint32_t off, int r0) {
    const uint32_t row = off >> 7, row2 = (uint32_t)(r0 - (int)row) & 63u;
    return (row2 << 7) | ((off & 0x7fu) ^ (((row ^ row2) & 7u) << 4));
}
// Reverse the 64 rows of NT 8 KB swizzled tiles in place (row x <-> 63 - x).  Warp w of the 8 compute warps owns rows
// 4w..4w+3 and their mirror images 60-4w..63-4w: a set closed under the reversal, so no two warps touch the same bytes.
template <int NT>
__device__ __forceinline__ void flip_tiles_full(const uint32_t (&tiles)[NT], int warp, int lane) {
    const int rr = lane & 7, row = rr < 4 ? 4 * warp + rr : 56 - 4 * warp + rr;     // rr 4..7 -> 60-4w .. 63-4w
    uint32_t x[NT][2][4];
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int half = 0; half < 2; half++)
            ldsm_x4(tiles[t] + sw128(row, 16 * ((lane >> 3) + 4 * half)), x[t][half][0], x[t][half][1], x[t][half][2], x[t][half][3]);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int half = 0; half < 2; half++)
            stsm_x4(tiles[t] + sw128(63 - row, 16 * ((lane >> 3) + 4 * half)), x[t][half][0], x[t][half][1], x[t][half][2], x[t][half][3]);
}
// Short last chunk (nv < 64 valid rows): row x -> (nv - 1 - x) mod 64, the rows x >= nv zeroed.  The images of a warp's
// rows belong to other warps, hence the barrier between reading and writing (all 256 compute threads call this).
template <int NT>
__device__ __forceinline__ void flip_tiles_short(const uint32_t (&tiles)[NT], int nv, int warp, int lane) {
    const int row = 8 * warp + (lane & 7), mine = 8 * warp + (lane >> 2);           // address row / row of my fragment
    uint32_t x[NT][2][4];
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int half = 0; half < 2; half++) {
            ldsm_x4(tiles[t] + sw128(row, 16 * ((lane >> 3) + 4 * half)), x[t][half][0], x[t][half][1], x[t][half][2], x[t][half][3]);
            if (mine >= nv) x[t][half][0] = x[t][half][1] = x[t][half][2] = x[t][half][3] = 0u;
        }
    named_bar_sync<B_SCAN_ID, 256>();
    const int dst = (nv - 1 - row) & 63;
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int half = 0; half < 2; half++)
            stsm_x4(tiles[t] + sw128(dst, 16 * ((lane >> 3) + 4 * half)), x[t][half][0], x[t][half][1], x[t][half][2], x[t][half][3]);
}
// zero the rows >= nv of an 8 KB tile (256 threads)
__device__ __forceinline__ void zero_tile_rows(uint8_t *tile, int nv, int tid) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int u = tid + 256 * k, row = u >> 3;
        if (row >= nv) *reinterpret_cast<uint4 *>(tile + u * 16) = make_uint4(0, 0, 0, 0);
    }
}
// packed bf16 pair of tokens (t0, t0 + 1): keep what lies before token nv
__device__ __forceinline__ uint32_t pair_mask(int t0, int nv) { return t0 >= nv ? 0u : t0 + 1 >= nv ? 0xffffu : 0xffffffffu; }

// Hand-offs between the issuer warp and the compute warps go through hardware named barriers
// (producer: bar.arrive, consumer: bar.sync, 288 threads): nobody spins.  Only lane 0 of the issuer
// warp ever polls an mbarrier (TMA / tcgen05.commit completion); a compute warp spinning on
// try_wait would take issue slots from the co-resident CTA that is doing useful work.
enum : int { B_SCAN = 1, B_RAW, B_PREP, B_M1, B_T1, B_M2, B_T2, B_M3, B_T3, B_T1A, B_T2A, B_BM, B_DR, B_FREE, B_VG };
template <int ID>
__device__ __forceinline__ void bar_arrive_all() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory"); }
template <int ID>
__device__ __forceinline__ void bar_sync_all() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory"); }

// 2^d as a packed bf16 pair for an integer d <= 0 (0 when below the normal range)
__device__ __forceinline__ uint32_t bfpow2pair(int d) {
    return (uint32_t)(max(d, -127) + 127) * 0x00800080u;        // both halves with one multiply: (e << 7) | (e << 23), e < 128
}
// the same for -127 <= d <= 0 (no clamp): a single multiply-add
__device__ __forceinline__ uint32_t bfpow2pair_nc(int d) { return (uint32_t)d * 0x00800080u + 127u * 0x00800080u; }
// exact scaling of packed bf16 pairs by 2^d, -254 <= d <= 0, as two factors (2^d alone may leave the bf16
// range although the scaled values do not).  d is the distance of two neighbouring block references: at most 16
// tokens x LCLAMP2 = 208 (+ 1 of rounding), so both halves are >= -127 and need no clamp.
__device__ __forceinline__ void scale1(uint32_t &a, int d) {
    a = hmul2(hmul2(a, bfpow2pair_nc(d >> 1)), bfpow2pair_nc(d - (d >> 1)));
}
__device__ __forceinline__ void scale2(uint32_t &a, uint32_t &b, int d) {
    const uint32_t fa = bfpow2pair_nc(d >> 1), fb = bfpow2pair_nc(d - (d >> 1));
    a = hmul2(hmul2(a, fa), fb);
    b = hmul2(hmul2(b, fa), fb);
}
__device__ __forceinline__ void scale4(uint32_t (&x)[4], int d) {
    const uint32_t fa = bfpow2pair_nc(d >> 1), fb = bfpow2pair_nc(d - (d >> 1));
#pragma unroll
    for (int g = 0; g < 4; g++) x[g] = hmul2(hmul2(x[g], fa), fb);
}

}  // namespace tc3
}  // namespace wkv6
