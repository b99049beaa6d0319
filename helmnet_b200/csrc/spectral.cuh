// spectral.cuh -- pseudo-spectral Laplacian with PML + Helmholtz residual + per-sample residual norm.
//
// Reference semantics:
//   helmnet/spectral.py:31-79   fast_laplacian_with_pml:  L u = ax*F^-1(i kx u^) + bx*F^-1(-kx^2 u^)
//                                                             + ay*F^-1(i ky u^) + by*F^-1(-ky^2 u^)
//   helmnet/hybridnet.py:544-556 get_residual:            r = L u + k_sq * u - source
//   helmnet/hybridnet.py:295-297 test_loss_function:      rmse_b = sqrt(mean_{c,h,w} r^2)
//
// The reference evaluates L with one 2-D FFT and four 2-D inverse FFTs.  The 2-D transform is separable
// and each term differentiates along ONE axis, so  F2^-1(i kx F2 u) == F_x^-1(i k F_x u)  exactly:
// L u = R(u) + C(u) with R acting on whole rows and C on whole columns (the operator has
// global support per axis, so a tile is a set of complete lines staged in shared memory -- there is no
// halo).  Two kernels:
//   spectral_rows_kernel : R(u) for L rows per CTA             -> rx
//   spectral_cols_kernel : C(u) for CW columns per CTA, fused with  r = rx + C(u) + k_sq*u - source,
//                          the store of r and the per-sample sum of squares (warp shuffles + one
//                          double atomicAdd per CTA).
// Line transforms are radix-4/2/generic Stockham FFTs in shared memory with twiddles computed in double
// precision on the host; inverse transforms use conj(FFT(conj(.))) with the 1/N folded into the
// spectral multipliers.
#pragma once
#include "common.cuh"

namespace hn {

constexpr int kMaxStages = 16;

struct SpecTables {
    const float2* tw;   // [n]  exp(-2 pi i k / n)
    const float* mk;    // [n]  k_kappa / n              (k cast to float32 first, spectral.py:141)
    const float* msq;   // [n]  -(k_kappa^2) / n         (k^2 in float32, spectral.py:281)
    const float2* a;    // [n]  -gamma' / gamma^3        (spectral.py:334-337), zero outside the PML strips
    const float2* b;    // [n]  1 / gamma^2
    int n, pml, nstages;
    int radix[kMaxStages];
};

// Line buffers are padded by one element every 16 (physical index = i + i/16) so that the stride-16 accesses of the
// radix-16 passes fall on distinct banks.
__host__ __device__ __forceinline__ int pidx(int i) { return i + (i >> 4); }
__host__ __device__ inline int line_pitch(int n) { return n + (n >> 4) + 1; }

// 16-point DFT in registers (4 x 4 Cooley-Tukey), natural-order in and out: a[j] <- sum_k a[k] w16^{jk}
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i t3
    a3 = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i t3
}
__device__ __forceinline__ void dft16(float2 (&a)[16]) {
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
    // columns n1 = 0..3: elements n1, n1+4, n1+8, n1+12 -> y[n1][k2] stored back in place (a[n1 + 4*k2])
#pragma unroll
    for (int n1 = 0; n1 < 4; n1++) dft4(a[n1], a[n1 + 4], a[n1 + 8], a[n1 + 12]);
    // twiddle y[n1][k2] *= w16^{n1*k2}
    a[5] = cmul(a[5], make_float2(C1, -S1));     // n1=1,k2=1
    a[9] = cmul(a[9], make_float2(R2, -R2));     // n1=1,k2=2
    a[13] = cmul(a[13], make_float2(S1, -C1));   // n1=1,k2=3
    a[6] = cmul(a[6], make_float2(R2, -R2));     // n1=2,k2=1
    a[10] = make_float2(a[10].y, -a[10].x);      // n1=2,k2=2: w^4 = -i
    a[14] = cmul(a[14], make_float2(-R2, -R2));  // n1=2,k2=3: w^6
    a[7] = cmul(a[7], make_float2(S1, -C1));     // n1=3,k2=1: w^3
    a[11] = cmul(a[11], make_float2(-R2, -R2));  // n1=3,k2=2: w^6
    a[15] = cmul(a[15], make_float2(-C1, S1));   // n1=3,k2=3: w^9
    // rows k2 = 0..3: over n1 -> z[k1][k2], output index j = k2 + 4*k1 lives in a[4*k2 + k1]
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) dft4(a[4 * k2], a[4 * k2 + 1], a[4 * k2 + 2], a[4 * k2 + 3]);
}

// Forward FFT of `nl` lines of length n stored with pitch lp (padded indexing) in x; y is scratch of the same
// shape.  Stockham autosort, decimation in frequency:  for stage length len = r*m and stride s,
//   y[q + s*(r*p + j)] = ( sum_k x[q + s*(p + m*k)] * w_r^{jk} ) * w_len^{p*j},   p < m, q < s.
// Radix 16 runs in registers (one thread = one 16-point butterfly), radix 4 / 2 / generic for what is left.
// Returns the buffer holding the result.  Ends with a __syncthreads().
__device__ inline float2* fft_lines(float2* x, float2* y, int nl, int lp, const float2* tw, const SpecTables& t) {
    const int n = t.n;
    int s = 1, len = n;
    for (int st = 0; st < t.nstages; st++) {
        const int r = t.radix[st];
        const int m = len / r;
        const int per_line = n / r;
        const int items = nl * per_line;
        for (int it = threadIdx.x; it < items; it += blockDim.x) {
            const int l = it / per_line, idx = it - l * per_line;
            const int p = idx / s, q = idx - p * s;
            const float2* xl = x + l * lp;
            float2* yl = y + l * lp;
            const int i0 = q + s * p, o0 = q + s * r * p;
            const int sm = s * m;
            if (r == 16) {
                float2 a[16];
#pragma unroll
                for (int k = 0; k < 16; k++) a[k] = xl[pidx(i0 + k * sm)];
                dft16(a);
                const int tp = p * s;
                // output j = k2 + 4*k1 is in a[4*k2 + k1]
#pragma unroll
                for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
                    for (int k2 = 0; k2 < 4; k2++) {
                        const int j = k2 + 4 * k1;
                        const float2 v = a[4 * k2 + k1];
                        yl[pidx(o0 + j * s)] = (j == 0) ? v : cmul(v, tw[tp * j]);
                    }
            } else if (r == 4) {
                const float2 a0 = xl[pidx(i0)], a1 = xl[pidx(i0 + sm)], a2 = xl[pidx(i0 + 2 * sm)], a3 = xl[pidx(i0 + 3 * sm)];
                const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
                const float2 b0 = cadd(t0, t2), b2 = csub(t0, t2);
                const float2 b1 = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i t3
                const float2 b3 = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i t3
                const int tp = p * s;
                yl[pidx(o0)] = b0;
                yl[pidx(o0 + s)] = cmul(b1, tw[tp]);
                yl[pidx(o0 + 2 * s)] = cmul(b2, tw[2 * tp]);
                yl[pidx(o0 + 3 * s)] = cmul(b3, tw[3 * tp]);
            } else if (r == 2) {
                const float2 a0 = xl[pidx(i0)], a1 = xl[pidx(i0 + sm)];
                yl[pidx(o0)] = cadd(a0, a1);
                yl[pidx(o0 + s)] = cmul(csub(a0, a1), tw[p * s]);
            } else {
                const int wr = n / r;
                for (int j = 0; j < r; j++) {
                    float2 acc = xl[pidx(i0)];
                    for (int k = 1; k < r; k++) acc = cadd(acc, cmul(xl[pidx(i0 + k * sm)], tw[((j * k) % r) * wr]));
                    yl[pidx(o0 + j * s)] = (j == 0) ? acc : cmul(acc, tw[p * s * j]);
                }
            }
        }
        __syncthreads();
        float2* tmp = x;
        x = y;
        y = tmp;
        len = m;
        s *= r;
    }
    return x;
}

// One axis of the operator applied to `nl` lines held in buffer A (pitch lp):
//   out_j = b_j * F^-1(-k^2 F u)_j + a_j * F^-1(i k F u)_j
// Uses three line buffers A,B,C and a small strip buffer S[nl][2*pml]; returns the buffer holding
// conj(F^-1(-k^2 u^)) (call it E) -- the caller finishes  b_j*conj(E_j) + strip term.
__device__ inline float2* axis_operator(float2* A, float2* B, float2* C, float2* S, int nl, int lp, const float2* tw,
                                        const SpecTables& t) {
    const int n = t.n, pml = t.pml;
    float2* F = fft_lines(A, B, nl, lp, tw, t);   // u^
    float2* O = (F == A) ? B : A;
    // first derivative (only its PML-strip samples are needed: a == 0 elsewhere)
    if (pml > 0) {
        for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
            const int l = it / n, k = it - l * n;
            const float2 v = F[l * lp + pidx(k)];
            const float mk = __ldg(t.mk + k);
            O[l * lp + pidx(k)] = make_float2(-mk * v.y, -mk * v.x);   // conj( (i k / n) * u^ )
        }
        __syncthreads();
        float2* D = fft_lines(O, C, nl, lp, tw, t);
        for (int it = threadIdx.x; it < nl * 2 * pml; it += blockDim.x) {
            const int l = it / (2 * pml), mth = it - l * 2 * pml;
            const int j = (mth < pml) ? mth : n - 2 * pml + mth;
            S[it] = cmul(__ldg(t.a + j), cconj(D[l * lp + pidx(j)]));
        }
        __syncthreads();
    }
    // second derivative
    for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
        const int l = it / n, k = it - l * n;
        const float2 v = F[l * lp + pidx(k)];
        const float ms = __ldg(t.msq + k);
        O[l * lp + pidx(k)] = make_float2(ms * v.x, -ms * v.y);        // conj( (-k^2 / n) * u^ )
    }
    __syncthreads();
    return fft_lines(O, C, nl, lp, tw, t);
}

__device__ __forceinline__ float2 axis_value(const float2* E, const float2* S, int l, int lp, int j, const SpecTables& t) {
    float2 v = cmul(__ldg(t.b + j), cconj(E[l * lp + pidx(j)]));
    const int n = t.n, pml = t.pml;
    if (j < pml) v = cadd(v, S[l * 2 * pml + j]);
    else if (j >= n - pml) v = cadd(v, S[l * 2 * pml + j - (n - 2 * pml)]);
    return v;
}

__host__ __device__ inline size_t spectral_smem_bytes(int n, int lines, int pml) {
    return (size_t)n * 8 + (size_t)3 * lines * line_pitch(n) * 8 + (size_t)lines * 2 * (pml > 0 ? pml : 1) * 8;
}

constexpr int SPEC_THREADS = 256;

__global__ void __launch_bounds__(SPEC_THREADS) spectral_rows_kernel(SpecTables t, const float2* __restrict__ u,
                                                                     float2* __restrict__ rx, int total_rows, int L) {
    HN_DYN_SMEM(float2, smem_sp);
    const int n = t.n, lp = line_pitch(n);
    float2* tw = smem_sp;
    float2* A = tw + n;
    float2* B = A + L * lp;
    float2* C = B + L * lp;
    float2* S = C + L * lp;
    const int row0 = blockIdx.x * L;
    const int nl = min(L, total_rows - row0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = __ldg(t.tw + i);
    pdl_wait();      // operator tables above; fields written by earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
        const int l = it / n, j = it - l * n;
        A[l * lp + pidx(j)] = __ldg(u + (size_t)(row0 + l) * n + j);
    }
    __syncthreads();
    const float2* E = axis_operator(A, B, C, S, nl, lp, tw, t);
    for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
        const int l = it / n, j = it - l * n;
        rx[(size_t)(row0 + l) * n + j] = axis_value(E, S, l, lp, j, t);
    }
}

struct ColsArgs {
    const float2* u;      // [B][n][n]
    const float2* rx;     // [B][n][n]  row-axis part
    const float* ksq;     // [B][n][n] or null
    const float2* src;    // [src_batch][n][n] or null
    const unsigned char* src_nz;   // [src_batch][n] or null: 1 where column j of the source map holds a non-zero (src_colnz_kernel)
    float2* res;          // [B][n][n]
    double* ssq;          // [slots][B] or null
    const int* slot;      // device scalar: which ssq slot (iteration index)
    unsigned* amax_out;   // running max |res| slot (publish_amax) or null
    int src_batch, B, CW;
    int b0;               // first sample of this launch (ssq is indexed by the absolute sample)
};

// Point sources are zero almost everywhere: one flag per (source map, column) lets a column tile skip the source loads
// of the residual epilogue altogether (subtracting an exact zero changes nothing, so results are bit-identical).
__global__ void src_colnz_kernel(const float2* __restrict__ src, unsigned char* __restrict__ nz, int n, int total_cols) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_cols) return;
    const int s = idx / n, j = idx - s * n;
    const float2* p = src + (size_t)s * n * n + j;
    bool any = false;
    for (int i = 0; i < n; i++) {
        const float2 v = p[(size_t)i * n];
        any = any || !(v.x == 0.f && v.y == 0.f);
    }
    nz[idx] = any ? 1 : 0;
}
// source pointer of the 8-column tile starting at column j0 of sample `b` (relative to the launch), or null when the
// whole tile of the source map is zero
__device__ __forceinline__ const float2* tile_source(const ColsArgs& a, int b, int n, int j0) {
    if (a.src == nullptr) return nullptr;
    const size_t sb = a.src_batch > 1 ? (size_t)b : (size_t)0;
    if (a.src_nz != nullptr) {
        const unsigned long long f = *reinterpret_cast<const unsigned long long*>(a.src_nz + sb * n + j0);   // 8 flags, j0 % 8 == 0
        if (f == 0ull) return nullptr;
    }
    return a.src + sb * (size_t)n * n + j0;
}

__global__ void __launch_bounds__(SPEC_THREADS) spectral_cols_kernel(SpecTables t, ColsArgs a) {
    HN_DYN_SMEM(float2, smem_sp);
    __shared__ float red[SPEC_THREADS / 32];
    const int n = t.n, lp = line_pitch(n), CW = a.CW;
    float2* tw = smem_sp;
    float2* A = tw + n;
    float2* B = A + CW * lp;
    float2* C = B + CW * lp;
    float2* S = C + CW * lp;
    const int b = blockIdx.y, j0 = blockIdx.x * CW;
    const int nc = min(CW, n - j0);
    const size_t img = (size_t)b * n * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = __ldg(t.tw + i);
    pdl_wait();      // operator tables above; fields written by earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    for (int it = threadIdx.x; it < n * CW; it += blockDim.x) {
        const int i = it / CW, c = it - i * CW;
        if (c < nc) A[c * lp + pidx(i)] = __ldg(a.u + img + (size_t)i * n + j0 + c);
    }
    __syncthreads();
    const float2* E = axis_operator(A, B, C, S, nc, lp, tw, t);
    float part = 0.f, lmax = 0.f;
    for (int it = threadIdx.x; it < n * CW; it += blockDim.x) {
        const int i = it / CW, c = it - i * CW;
        if (c >= nc) continue;
        const size_t p = img + (size_t)i * n + j0 + c;
        float2 r = cadd(__ldg(a.rx + p), axis_value(E, S, c, lp, i, t));
        if (a.ksq != nullptr) {
            const float kq = __ldg(a.ksq + p);
            const float2 uu = __ldg(a.u + p);
            r.x = fmaf(kq, uu.x, r.x);
            r.y = fmaf(kq, uu.y, r.y);
        }
        if (a.src != nullptr) {
            const float2 sv = __ldg(a.src + (a.src_batch > 1 ? img : (size_t)0) + (size_t)i * n + j0 + c);
            r.x -= sv.x;
            r.y -= sv.y;
        }
        a.res[p] = r;
        part = fmaf(r.x, r.x, part);
        part = fmaf(r.y, r.y, part);
        lmax = fmaxf(lmax, fmaxf(fabsf(r.x), fabsf(r.y)));
    }
    publish_amax(a.amax_out, lmax);
    if (a.ssq != nullptr) {
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int w = 0; w < SPEC_THREADS / 32; w++) tot += red[w];
            atomicAdd(a.ssq + (size_t)(*a.slot) * a.B + a.b0 + b, (double)tot);
        }
    }
}

}  // namespace hn
