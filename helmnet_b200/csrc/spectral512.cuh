// spectral512.cuh -- the N = 512 fast path of the spectral residual stage (same mathematics and reference semantics as
// spectral.cuh / spectral256.cuh: helmnet/spectral.py:31-79, helmnet/hybridnet.py:544-556, :295-297).
//
// A line of 512 points is owned by one warp: the two half-warps g = 0, 1 hold the even / odd samples and each runs the
// register-resident 256-point transform of spectral256.cuh; one radix-2 butterfly across the half-warps (shuffle with
// lane ^ 16) completes the 512-point transform:
//   forward (decimation in time):  X[k] = E[k] + w512^k O[k],  X[k + 256] = E[k] - w512^k O[k]
//       half-warp g ends up with the frequencies k + 256 g, k = h + 16 j;
//   the two following transforms take exactly that layout (decimation in frequency):
//       Y[2m] = F256(W[k] + W[k+256])[m],  Y[2m+1] = F256((W[k] - W[k+256]) w512^k)[m]
//       half-warp g ends up with the positions n = 2 (h + 16 j) + g -- the layout the line was loaded in.
// So, as at N = 256, the three transforms of a line need no block-wide barrier.
#pragma once
#include "spectral256.cuh"

namespace hn {
namespace s512 {

constexpr int N = 512;
constexpr int TB = 273;                  // per-unit transpose buffer pitch; 273 = 1 mod 16 keeps the column epilogue's reads
                                         // across the 16 unit buffers (8 columns x even/odd) on distinct banks
constexpr int LINES = 8;                 // lines per CTA: one warp each
constexpr int THREADS = 256;
constexpr int UNITS = 16;                // 256-point transform units per CTA (2 per line)
constexpr int TILE_P = 9;

struct Tab {
    float2 tw[256];                      // tw[16 j + h] = w256^(h j)
    float2 tw2[256];                     // tw2[k] = w512^k
    float2 b[N];
    float mk[N];
    float msq[N];
};

__device__ __forceinline__ void load_tab(Tab& tab, const SpecTables& t) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        if (i < 256) {
            tab.tw[i] = __ldg(t.tw + 2 * ((i & 15) * (i >> 4)));     // t.tw[k] = w512^k
            tab.tw2[i] = __ldg(t.tw + i);
        }
        tab.b[i] = __ldg(t.b + i);
        tab.mk[i] = __ldg(t.mk + i);
        tab.msq[i] = __ldg(t.msq + i);
    }
}

__device__ __forceinline__ float2 shfl_xor16(float2 v) {
    return make_float2(__shfl_xor_sync(0xffffffffu, v.x, 16), __shfl_xor_sync(0xffffffffu, v.y, 16));
}

// in: a[k] = x[2 (h + 16 k) + g];  out: a[j] = X[h + 16 j + 256 g]
__device__ __forceinline__ void fft512_dit(float2 (&a)[16], float2* tb, int h, int g, const Tab& tab) {
    s256::fft256(a, tb, h, tab.tw);
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (g) a[j] = cmul(a[j], tab.tw2[h + 16 * j]);
        const float2 p = shfl_xor16(a[j]);
        a[j] = g ? csub(p, a[j]) : cadd(a[j], p);
    }
}
// in: a[j] = W[h + 16 j + 256 g];  out: a[j] = Y[2 (h + 16 j) + g]
__device__ __forceinline__ void fft512_dif(float2 (&a)[16], float2* tb, int h, int g, const Tab& tab) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const float2 p = shfl_xor16(a[j]);
        a[j] = g ? cmul(csub(p, a[j]), tab.tw2[h + 16 * j]) : cadd(a[j], p);
    }
    s256::fft256(a, tb, h, tab.tw);
}

// One axis of the operator for the line held as X[h + 16 j + 256 g] -> out[j] at positions n = 2 (h + 16 j) + g
__device__ __forceinline__ void axis512(const float2 (&X)[16], float2 (&out)[16], float2* tb, int h, int g, const Tab& tab,
                                        const float2* a_tab, int pml) {
    float2 w[16];
    float2 strip_lo = make_float2(0.f, 0.f), strip_hi = strip_lo;
    const int q = 2 * h + g;              // position within a block of 32: n = q + 32 j
    if (pml > 0) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const float mk = tab.mk[h + 16 * j + 256 * g];
            w[j] = make_float2(-mk * X[j].y, -mk * X[j].x);          // conj( (i k / n) X )
        }
        fft512_dif(w, tb, h, g, tab);
        // strips: n < pml (j = 0, q < pml) and n >= 512 - pml (j = 15, q >= 32 - pml); pml <= 16
        if (q < pml) strip_lo = cmul(__ldg(a_tab + q), cconj(w[0]));
        if (q >= 32 - pml) strip_hi = cmul(__ldg(a_tab + 480 + q), cconj(w[15]));
    }
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const float ms = tab.msq[h + 16 * j + 256 * g];
        w[j] = make_float2(ms * X[j].x, -ms * X[j].y);               // conj( (-k^2 / n) X )
    }
    fft512_dif(w, tb, h, g, tab);
#pragma unroll
    for (int j = 0; j < 16; j++) out[j] = cmul(tab.b[q + 32 * j], cconj(w[j]));
    if (pml > 0) {
        if (q < pml) out[0] = cadd(out[0], strip_lo);
        if (q >= 32 - pml) out[15] = cadd(out[15], strip_hi);
    }
}

struct RowsSmem {
    Tab tab;
    float2 tbuf[UNITS][TB];
};
constexpr size_t ROWS_SMEM_BYTES = sizeof(RowsSmem);

__global__ void __launch_bounds__(THREADS) spectral_rows512_kernel(SpecTables t, const float2* __restrict__ u,
                                                                   float2* __restrict__ rx, int total_rows) {
    HN_DYN_SMEM(unsigned char, smem_raw);
    RowsSmem& sh = *reinterpret_cast<RowsSmem*>(smem_raw);
    load_tab(sh.tab, t);
    __syncthreads();
    pdl_wait();      // operator tables above; fields written by earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane & 15, g = lane >> 4;
    const int row = blockIdx.x * LINES + warp;
    const bool live = row < total_rows;
    const size_t base = (size_t)(live ? row : 0) * N;
    float2* tb = sh.tbuf[warp * 2 + g];
    float2 X[16], o[16];
#pragma unroll
    for (int k = 0; k < 16; k++) X[k] = live ? __ldg(u + base + 2 * (h + 16 * k) + g) : make_float2(0.f, 0.f);
    fft512_dit(X, tb, h, g, sh.tab);
    axis512(X, o, tb, h, g, sh.tab, t.a, t.pml);
    if (live) {
#pragma unroll
        for (int j = 0; j < 16; j++) rx[base + 2 * (h + 16 * j) + g] = o[j];
    }
}

struct ColsSmem {
    Tab tab;
    float2 tbuf[UNITS][TB];
    float2 tile[N * TILE_P];             // u, 8 columns x 512 rows, pitch 9
    float red[THREADS / 32];
};
constexpr size_t COLS_SMEM_BYTES = sizeof(ColsSmem);

__global__ void __launch_bounds__(THREADS) spectral_cols512_kernel(SpecTables t, ColsArgs a) {
    HN_DYN_SMEM(unsigned char, smem_raw);
    ColsSmem& sh = *reinterpret_cast<ColsSmem*>(smem_raw);
    const int b = blockIdx.y, j0 = blockIdx.x * LINES;
    const size_t img = (size_t)b * N * N;
    pdl_wait();      // common.cuh: HN_LAUNCH_PDL
    pdl_trigger();
    for (int it = threadIdx.x; it < N * LINES; it += THREADS) {
        const int i = it >> 3, c = it & 7;
        s256::cp_async8(&sh.tile[i * TILE_P + c], a.u + img + (size_t)i * N + j0 + c);
    }
    s256::cp_async_commit();
#ifndef HN_EMU
    // the epilogue's operands (rx, k_sq) are fetched towards L2 now so that their latency hides behind the transforms
    for (int i = threadIdx.x; i < N; i += THREADS) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rx + img + (size_t)i * N + j0));
        if (a.ksq != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ksq + img + (size_t)i * N + j0));
    }
#endif
    load_tab(sh.tab, t);
    s256::cp_async_wait<0>();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane & 15, g = lane >> 4;
    {
        float2* tb = sh.tbuf[warp * 2 + g];
        float2 X[16], o[16];
#pragma unroll
        for (int k = 0; k < 16; k++) X[k] = sh.tile[(2 * (h + 16 * k) + g) * TILE_P + warp];
        fft512_dit(X, tb, h, g, sh.tab);
        axis512(X, o, tb, h, g, sh.tab, t.a, t.pml);
        // park C(u): unit (column, parity g) holds the positions n = 2 m + g at index m
#pragma unroll
        for (int j = 0; j < 16; j++) tb[pidx(h + 16 * j)] = o[j];
    }
    __syncthreads();
    float part = 0.f, lmax = 0.f;
    constexpr int EPI_CHUNK = 8;
    const float2* srcp = tile_source(a, b, N, j0);
#pragma unroll 1
    for (int it0 = threadIdx.x; it0 < N * LINES; it0 += THREADS * EPI_CHUNK) {
        float2 sv[EPI_CHUNK], rxv[EPI_CHUNK];
        float kq[EPI_CHUNK];
#pragma unroll
        for (int q = 0; q < EPI_CHUNK; q++) {
            const int it = it0 + q * THREADS;
            const size_t off = (size_t)(it >> 3) * N + (it & 7);
            rxv[q] = __ldg(a.rx + img + j0 + off);
            kq[q] = a.ksq != nullptr ? __ldg(a.ksq + img + j0 + off) : 0.f;
            sv[q] = srcp != nullptr ? __ldg(srcp + off) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < EPI_CHUNK; q++) {
            const int it = it0 + q * THREADS;
            const int i = it >> 3, c = it & 7;
            float2 r = cadd(rxv[q], sh.tbuf[2 * c + (i & 1)][pidx(i >> 1)]);
            const float2 uu = sh.tile[i * TILE_P + c];
            r.x = fmaf(kq[q], uu.x, r.x);
            r.y = fmaf(kq[q], uu.y, r.y);
            r.x -= sv[q].x;
            r.y -= sv[q].y;
            a.res[img + (size_t)i * N + j0 + c] = r;
            part = fmaf(r.x, r.x, part);
            part = fmaf(r.y, r.y, part);
            lmax = fmaxf(lmax, fmaxf(fabsf(r.x), fabsf(r.y)));
        }
    }
    publish_amax(a.amax_out, lmax);
    if (a.ssq != nullptr) {
        part = warp_sum(part);
        if (lane == 0) sh.red[warp] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int w = 0; w < THREADS / 32; w++) tot += sh.red[w];
            atomicAdd(a.ssq + (size_t)(*a.slot) * a.B + a.b0 + b, (double)tot);
        }
    }
}

}  // namespace s512
}  // namespace hn
