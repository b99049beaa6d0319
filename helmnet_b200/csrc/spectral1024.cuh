// spectral1024.cuh -- the N = 1024 fast path of the spectral residual stage (same mathematics and reference semantics as
// spectral.cuh / spectral256.cuh / spectral512.cuh: helmnet/spectral.py:31-79, helmnet/hybridnet.py:544-556, :295-297).
//
// A line of 1024 points is owned by two warps s = 0, 1; warp s holds the positions n + 512 s (n < 512), i.e. contiguous
// halves of the line.  One radix-2 butterfly ACROSS the two warps (through the transpose buffers, two block barriers)
// wraps the warp-level 512-point transforms of spectral512.cuh:
//   forward (decimation in frequency):  Y[2m] = F512(x[n] + x[n+512])[m],  Y[2m+1] = F512((x[n] - x[n+512]) w1024^n)[m]
//       warp s ends up with the frequencies 2 m + s, m = h + 16 j + 256 g;
//   the two following transforms take that layout (decimation in time):
//       Z[m'] = E[m'] + w1024^m' O[m'],  Z[m' + 512] = E[m'] - w1024^m' O[m'],   E/O = F512 of the even/odd frequencies
//       warp s ends up with the positions m' + 512 s, m' = 2 (h + 16 j) + g -- the layout the line was loaded in.
// Every warp of a CTA executes the same sequence, so the block barriers inside the transforms are uniform.
#pragma once
#include "spectral512.cuh"

namespace hn {
namespace s1024 {

constexpr int N = 1024;
constexpr int TB = 273;                  // = 1 mod 16: see the unit numbering in the column kernel
constexpr int TILE_P = 9;

struct Tab {
    float2 tw[256];                      // tw[16 j + h] = w256^(h j)
    float2 tw2[256];                     // w512^k
    float2 tw4[512];                     // w1024^n
    float2 b[N];
    float mk[N];
    float msq[N];
};

__device__ __forceinline__ void load_tab(Tab& tab, const SpecTables& t) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        if (i < 256) {
            tab.tw[i] = __ldg(t.tw + 4 * ((i & 15) * (i >> 4)));     // t.tw[k] = w1024^k
            tab.tw2[i] = __ldg(t.tw + 2 * i);
        }
        if (i < 512) tab.tw4[i] = __ldg(t.tw + i);
        tab.b[i] = __ldg(t.b + i);
        tab.mk[i] = __ldg(t.mk + i);
        tab.msq[i] = __ldg(t.msq + i);
    }
}

// the warp-level 512-point transforms on this header's table type
__device__ __forceinline__ void fft512_dit(float2 (&a)[16], float2* tb, int h, int g, const Tab& tab) {
    s256::fft256(a, tb, h, tab.tw);
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (g) a[j] = cmul(a[j], tab.tw2[h + 16 * j]);
        const float2 p = s512::shfl_xor16(a[j]);
        a[j] = g ? csub(p, a[j]) : cadd(a[j], p);
    }
}
__device__ __forceinline__ void fft512_dif(float2 (&a)[16], float2* tb, int h, int g, const Tab& tab) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const float2 p = s512::shfl_xor16(a[j]);
        a[j] = g ? cmul(csub(p, a[j]), tab.tw2[h + 16 * j]) : cadd(a[j], p);
    }
    s256::fft256(a, tb, h, tab.tw);
}

// Exchange the 16 values of every thread with the thread of the partner warp (other s, same g, h): own values go to the
// own unit buffer, the partner's are read from its unit buffer.  Two block barriers (the buffers are reused right after).
__device__ __forceinline__ void exchange(const float2 (&a)[16], float2 (&p)[16], float2* tb_own, const float2* tb_partner, int h) {
#pragma unroll
    for (int j = 0; j < 16; j++) tb_own[pidx(h + 16 * j)] = a[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; j++) p[j] = tb_partner[pidx(h + 16 * j)];
    __syncthreads();
}

// in: a[k] = x[2 (h + 16 k) + g + 512 s];  out: a[j] = X[2 (h + 16 j + 256 g) + s]
__device__ __forceinline__ void fft1024_fwd(float2 (&a)[16], float2* tb, const float2* tbp, int h, int g, int s, const Tab& tab) {
    float2 p[16];
    exchange(a, p, tb, tbp, h);
#pragma unroll
    for (int k = 0; k < 16; k++)
        a[k] = s ? cmul(csub(p[k], a[k]), tab.tw4[2 * (h + 16 * k) + g]) : cadd(a[k], p[k]);
    fft512_dit(a, tb, h, g, tab);
}
// in: a[j] = W[2 (h + 16 j + 256 g) + s];  out: a[j] = Y[2 (h + 16 j) + g + 512 s]
__device__ __forceinline__ void fft1024_inv(float2 (&a)[16], float2* tb, const float2* tbp, int h, int g, int s, const Tab& tab) {
    fft512_dif(a, tb, h, g, tab);
    if (s) {
#pragma unroll
        for (int j = 0; j < 16; j++) a[j] = cmul(a[j], tab.tw4[2 * (h + 16 * j) + g]);
    }
    float2 p[16];
    exchange(a, p, tb, tbp, h);
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = s ? csub(p[j], a[j]) : cadd(a[j], p[j]);
}

// One axis of the operator for the line held as X[2 (h + 16 j + 256 g) + s] -> out[j] at positions 2 (h + 16 j) + g + 512 s
__device__ __forceinline__ void axis1024(const float2 (&X)[16], float2 (&out)[16], float2* tb, const float2* tbp, int h, int g, int s,
                                         const Tab& tab, const float2* a_tab, int pml) {
    float2 w[16];
    float2 strip_lo = make_float2(0.f, 0.f), strip_hi = strip_lo;
    const int q = 2 * h + g;              // position n = q + 32 j + 512 s
    if (pml > 0) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const float mk = tab.mk[2 * (h + 16 * j + 256 * g) + s];
            w[j] = make_float2(-mk * X[j].y, -mk * X[j].x);          // conj( (i k / n) X )
        }
        fft1024_inv(w, tb, tbp, h, g, s, tab);
        // strips: n < pml (s = 0, j = 0, q < pml) and n >= 1024 - pml (s = 1, j = 15, q >= 32 - pml); pml <= 16
        if (s == 0 && q < pml) strip_lo = cmul(__ldg(a_tab + q), cconj(w[0]));
        if (s == 1 && q >= 32 - pml) strip_hi = cmul(__ldg(a_tab + 992 + q), cconj(w[15]));
    }
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const float ms = tab.msq[2 * (h + 16 * j + 256 * g) + s];
        w[j] = make_float2(ms * X[j].x, -ms * X[j].y);               // conj( (-k^2 / n) X )
    }
    fft1024_inv(w, tb, tbp, h, g, s, tab);
#pragma unroll
    for (int j = 0; j < 16; j++) out[j] = cmul(tab.b[q + 32 * j + 512 * s], cconj(w[j]));
    if (pml > 0) {
        if (s == 0 && q < pml) out[0] = cadd(out[0], strip_lo);
        if (s == 1 && q >= 32 - pml) out[15] = cadd(out[15], strip_hi);
    }
}

// ---- rows: 4 lines per CTA of 256 threads ---------------------------------------------------------------------------
constexpr int ROWS_LINES = 4, ROWS_THREADS = 256, ROWS_UNITS = 16;
struct RowsSmem {
    Tab tab;
    float2 tbuf[ROWS_UNITS][TB];
};
constexpr size_t ROWS_SMEM_BYTES = sizeof(RowsSmem);

__global__ void __launch_bounds__(ROWS_THREADS) spectral_rows1024_kernel(SpecTables t, const float2* __restrict__ u,
                                                                         float2* __restrict__ rx, int total_rows) {
    HN_DYN_SMEM(unsigned char, smem_raw);
    RowsSmem& sh = *reinterpret_cast<RowsSmem*>(smem_raw);
    load_tab(sh.tab, t);
    __syncthreads();
    pdl_wait();      // operator tables above; fields written by earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane & 15, g = lane >> 4, s = warp & 1;
    const int row = blockIdx.x * ROWS_LINES + (warp >> 1);
    const bool live = row < total_rows;
    const size_t base = (size_t)(live ? row : 0) * N + 512 * s;
    float2* tb = sh.tbuf[warp * 2 + g];
    const float2* tbp = sh.tbuf[(warp ^ 1) * 2 + g];
    float2 X[16], o[16];
#pragma unroll
    for (int k = 0; k < 16; k++) X[k] = live ? __ldg(u + base + 2 * (h + 16 * k) + g) : make_float2(0.f, 0.f);
    fft1024_fwd(X, tb, tbp, h, g, s, sh.tab);
    axis1024(X, o, tb, tbp, h, g, s, sh.tab, t.a, t.pml);
    if (live) {
#pragma unroll
        for (int j = 0; j < 16; j++) rx[base + 2 * (h + 16 * j) + g] = o[j];
    }
}

// ---- columns: 8 columns per CTA of 512 threads -----------------------------------------------------------------------
constexpr int COLS = 8, COLS_THREADS = 512, COLS_UNITS = 32;
struct ColsSmem {
    Tab tab;
    float2 tbuf[COLS_UNITS][TB];         // unit = (2 s + g) * 8 + column: with TB = 1 mod 16 the epilogue's reads of the 8
                                         // columns x 2 parities of a half-warp fall on 16 distinct bank pairs
    float2 tile[N * TILE_P];             // u, 8 columns x 1024 rows, pitch 9
    float red[COLS_THREADS / 32];
};
constexpr size_t COLS_SMEM_BYTES = sizeof(ColsSmem);

__global__ void __launch_bounds__(COLS_THREADS) spectral_cols1024_kernel(SpecTables t, ColsArgs a) {
    HN_DYN_SMEM(unsigned char, smem_raw);
    ColsSmem& sh = *reinterpret_cast<ColsSmem*>(smem_raw);
    const int b = blockIdx.y, j0 = blockIdx.x * COLS;
    const size_t img = (size_t)b * N * N;
    pdl_wait();      // common.cuh: HN_LAUNCH_PDL
    pdl_trigger();
    for (int it = threadIdx.x; it < N * COLS; it += COLS_THREADS) {
        const int i = it >> 3, c = it & 7;
        s256::cp_async8(&sh.tile[i * TILE_P + c], a.u + img + (size_t)i * N + j0 + c);
    }
    s256::cp_async_commit();
#ifndef HN_EMU
    for (int i = threadIdx.x; i < N; i += COLS_THREADS) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rx + img + (size_t)i * N + j0));
        if (a.ksq != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ksq + img + (size_t)i * N + j0));
    }
#endif
    load_tab(sh.tab, t);
    s256::cp_async_wait<0>();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane & 15, g = lane >> 4, s = warp & 1, col = warp >> 1;
    {
        float2* tb = sh.tbuf[(2 * s + g) * 8 + col];
        const float2* tbp = sh.tbuf[(2 * (s ^ 1) + g) * 8 + col];
        float2 X[16], o[16];
#pragma unroll
        for (int k = 0; k < 16; k++) X[k] = sh.tile[(2 * (h + 16 * k) + g + 512 * s) * TILE_P + col];
        fft1024_fwd(X, tb, tbp, h, g, s, sh.tab);
        axis1024(X, o, tb, tbp, h, g, s, sh.tab, t.a, t.pml);
        // park C(u): unit (s, g, column) holds the positions n = 2 m + g + 512 s at index m
#pragma unroll
        for (int j = 0; j < 16; j++) tb[pidx(h + 16 * j)] = o[j];
    }
    __syncthreads();
    float part = 0.f, lmax = 0.f;
    constexpr int EPI_CHUNK = 8;
    const float2* srcp = tile_source(a, b, N, j0);
#pragma unroll 1
    for (int it0 = threadIdx.x; it0 < N * COLS; it0 += COLS_THREADS * EPI_CHUNK) {
        float2 sv[EPI_CHUNK], rxv[EPI_CHUNK];
        float kq[EPI_CHUNK];
#pragma unroll
        for (int q = 0; q < EPI_CHUNK; q++) {
            const int it = it0 + q * COLS_THREADS;
            const size_t off = (size_t)(it >> 3) * N + (it & 7);
            rxv[q] = __ldg(a.rx + img + j0 + off);
            kq[q] = a.ksq != nullptr ? __ldg(a.ksq + img + j0 + off) : 0.f;
            sv[q] = srcp != nullptr ? __ldg(srcp + off) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < EPI_CHUNK; q++) {
            const int it = it0 + q * COLS_THREADS;
            const int i = it >> 3, c = it & 7;
            const int m = i & 511;
            float2 r = cadd(rxv[q], sh.tbuf[(2 * (i >> 9) + (m & 1)) * 8 + c][pidx(m >> 1)]);
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
            for (int w = 0; w < COLS_THREADS / 32; w++) tot += sh.red[w];
            atomicAdd(a.ssq + (size_t)(*a.slot) * a.B + a.b0 + b, (double)tot);
        }
    }
}

}  // namespace s1024
}  // namespace hn
