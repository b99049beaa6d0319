// spectral256.cuh -- the N = 256 fast path of the spectral residual stage (same mathematics and reference semantics
// as spectral.cuh: helmnet/spectral.py:31-79, helmnet/hybridnet.py:544-556, :295-297).
//
// A line of 256 = 16 x 16 points is owned by 16 threads (half a warp), 16 complex values per thread in registers.
// One 256-point FFT = radix-16 butterfly in registers, a transpose through a small per-line shared-memory buffer
// synchronised with __syncwarp only, second radix-16 butterfly.  The layout closes on itself: after the forward
// transform thread h holds the frequencies h + 16 j, which is exactly the input layout of the (conjugated) inverse
// transforms, so the three transforms of a line (u^ ; first derivative for the PML strips ; second derivative) need
// no block-wide barrier and only three shared-memory transposes.
//   rows kernel: global -> registers directly (16 lanes read 128 contiguous bytes), results registers -> global.
//   cols kernel: 8 columns of one sample staged with cp.async through a padded shared-memory tile (coalesced 64-byte
//                segments), fused with  r = rx + C(u) + k_sq u - source,  sum r^2  and  max |r|; the epilogue's operands
//                are prefetched to L2 at kernel start and loaded 8 points at a time before their first use.
#pragma once
#include "spectral.cuh"

namespace hn {
namespace s256 {

constexpr int N = 256;
constexpr int TB = 274;                  // padded transpose buffer per line (pidx(255) = 270); 274 = 2 mod 16 keeps the
                                         // column epilogue's reads across the 8 line buffers on distinct banks
constexpr int LINES = 8;                 // lines per CTA (4 warps x 2)
constexpr int THREADS = 128;
constexpr int TILE_P = 9;                // column tile pitch (8 columns + 1 pad)

struct Tab {                             // shared-memory copies of the operator tables
    float2 tw[N];                        // tw[16 j + h] = w256^(h j): lane h reads consecutive entries (no bank conflicts)
    float2 b[N];
    float mk[N];
    float msq[N];
};

// natural-order 16-point DFT: a[j] <- sum_k a[k] w16^{jk}
__device__ __forceinline__ void dft16n(float2 (&a)[16]) {
    dft16(a);                            // output j = k2 + 4 k1 sits in a[4 k2 + k1]
    float2 t[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) t[k2 + 4 * k1] = a[4 * k2 + k1];
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = t[j];
}

// in: a[k] = x[h + 16 k];  out: a[j] = X[h + 16 j].   tb: this line's transpose buffer.  Whole warp must call.
__device__ __forceinline__ void fft256(float2 (&a)[16], float2* tb, int h, const float2* tw) {
    dft16n(a);
#pragma unroll
    for (int j = 0; j < 16; j++) tb[pidx(16 * h + j)] = (j == 0) ? a[0] : cmul(a[j], tw[16 * j + h]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = tb[pidx(h + 16 * k)];
    __syncwarp();
    dft16n(a);
}

// One axis of the operator for the line held as X[h + 16 j] (forward spectrum) -> out[j] at positions n = h + 16 j:
//   out_n = b_n * F^-1(-k^2 X)_n + a_n * F^-1(i k X)_n     (a is non-zero only in the PML strips)
__device__ __forceinline__ void axis256(const float2 (&X)[16], float2 (&out)[16], float2* tb, int h, const Tab& tab,
                                        const float2* a_tab, int pml) {
    float2 w[16];
    float2 strip_lo = make_float2(0.f, 0.f), strip_hi = strip_lo;
    if (pml > 0) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const float mk = tab.mk[h + 16 * j];
            w[j] = make_float2(-mk * X[j].y, -mk * X[j].x);          // conj( (i k / n) X )
        }
        fft256(w, tb, h, tab.tw);
        // positions in the strips: n < pml (j = 0, h < pml) and n >= 256 - pml (j = 15, h >= 16 - pml); pml <= 16
        if (h < pml) strip_lo = cmul(__ldg(a_tab + h), cconj(w[0]));
        if (h >= 16 - pml) strip_hi = cmul(__ldg(a_tab + 240 + h), cconj(w[15]));
    }
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const float ms = tab.msq[h + 16 * j];
        w[j] = make_float2(ms * X[j].x, -ms * X[j].y);               // conj( (-k^2 / n) X )
    }
    fft256(w, tb, h, tab.tw);
#pragma unroll
    for (int j = 0; j < 16; j++) out[j] = cmul(tab.b[h + 16 * j], cconj(w[j]));
    if (pml > 0) {
        if (h < pml) out[0] = cadd(out[0], strip_lo);
        if (h >= 16 - pml) out[15] = cadd(out[15], strip_hi);
    }
}

__device__ __forceinline__ void load_tab(Tab& tab, const SpecTables& t) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        tab.tw[i] = __ldg(t.tw + (i & 15) * (i >> 4));
        tab.b[i] = __ldg(t.b + i);
        tab.mk[i] = __ldg(t.mk + i);
        tab.msq[i] = __ldg(t.msq + i);
    }
}

__global__ void __launch_bounds__(THREADS) spectral_rows256_kernel(SpecTables t, const float2* __restrict__ u,
                                                                   float2* __restrict__ rx, int total_rows) {
    __shared__ Tab tab;
    __shared__ float2 tbuf[LINES][TB];
    load_tab(tab, t);
    __syncthreads();
    pdl_wait();      // operator tables above; fields written by earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane & 15, ll = warp * 2 + (lane >> 4);
    const int row = blockIdx.x * LINES + ll;
    const bool live = row < total_rows;
    const size_t base = (size_t)(live ? row : 0) * N;
    float2 X[16], o[16];
#pragma unroll
    for (int k = 0; k < 16; k++) X[k] = live ? __ldg(u + base + h + 16 * k) : make_float2(0.f, 0.f);
    fft256(X, tbuf[ll], h, tab.tw);
    axis256(X, o, tbuf[ll], h, tab, t.a, t.pml);
    if (live) {
#pragma unroll
        for (int j = 0; j < 16; j++) rx[base + h + 16 * j] = o[j];
    }
}

// asynchronous global -> shared copies (LDGSTS): the whole working set of a column tile is put in flight at once
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
#ifdef HN_EMU
    *reinterpret_cast<float2*>(dst) = *reinterpret_cast<const float2*>(src);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
#ifdef HN_EMU
    *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef HN_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
#ifndef HN_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#endif
}

struct ColsSmem {                        // 42 KB -> 5 CTAs per SM (staging rx and k_sq as well: 67 KB, 3 CTAs, measured slower)
    Tab tab;
    float2 tbuf[LINES][TB];
    float2 tile[N * TILE_P];
    float red[THREADS / 32];
};
constexpr size_t COLS_SMEM_BYTES = sizeof(ColsSmem);

__global__ void __launch_bounds__(THREADS) spectral_cols256_kernel(SpecTables t, ColsArgs a) {
    HN_DYN_SMEM(unsigned char, smem_raw);
    ColsSmem& sh = *reinterpret_cast<ColsSmem*>(smem_raw);
    const int b = blockIdx.y, j0 = blockIdx.x * LINES;
    const size_t img = (size_t)b * N * N;
    pdl_wait();      // common.cuh: HN_LAUNCH_PDL
    pdl_trigger();
    for (int it = threadIdx.x; it < N * LINES; it += THREADS) {
        const int i = it >> 3, c = it & 7;
        cp_async8(&sh.tile[i * TILE_P + c], a.u + img + (size_t)i * N + j0 + c);
    }
    cp_async_commit();
#ifndef HN_EMU
    for (int i = threadIdx.x; i < N; i += THREADS) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rx + img + (size_t)i * N + j0));
        if (a.ksq != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ksq + img + (size_t)i * N + j0));
    }
#endif
    load_tab(sh.tab, t);
    cp_async_wait<0>();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane & 15, ll = warp * 2 + (lane >> 4);
    {
        float2 X[16], o[16];
#pragma unroll
        for (int k = 0; k < 16; k++) X[k] = sh.tile[(h + 16 * k) * TILE_P + ll];
        fft256(X, sh.tbuf[ll], h, sh.tab.tw);
        axis256(X, o, sh.tbuf[ll], h, sh.tab, t.a, t.pml);
#pragma unroll
        for (int j = 0; j < 16; j++) sh.tbuf[ll][pidx(h + 16 * j)] = o[j];
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
            float2 r = cadd(rxv[q], sh.tbuf[c][pidx(i)]);
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

}  // namespace s256
}  // namespace hn
