// layout.cuh -- boundary kernels: conversion between the reference's tensor layouts (NCHW, possibly
// strided) and the solver's resident layouts (interleaved complex float2 / NHWC8), plus the solve setup
// of IterativeSolver.get_initials (helmnet/hybridnet.py:522-538).  All are off the per-iteration path
// except reset_amax_kernel (64 threads, first kernel of an iteration) and the optional history snapshots.
#pragma once
#include "common.cuh"

namespace hn {

constexpr int LAY_THREADS = 256;

// [B,2,H,W] -> float2 [B,H,W]
__global__ void nchw2_to_c2_kernel(const float* __restrict__ in, float2* __restrict__ out, int hw, size_t total,
                                   unsigned* amax_out) {
    float lmax = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        const float2 v = make_float2(in[(b * 2) * hw + p], in[(b * 2 + 1) * hw + p]);
        out[i] = v;
        lmax = fmaxf(lmax, fmaxf(fabsf(v.x), fabsf(v.y)));
    }
    publish_amax(amax_out, lmax);
}
// float2 [B,H,W] -> [B,2,H,W] with an output batch stride (in floats) so that per-level hidden states can be
// scattered into the flattened [B,2,S] layout of HybridNet.flatten_state (architectures.py:419-423).
__global__ void c2_to_nchw2_kernel(const float2* __restrict__ in, float* __restrict__ out, int hw, size_t total,
                                   size_t out_bstride, size_t out_cstride) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        const float2 v = in[i];
        out[b * out_bstride + p] = v.x;
        out[b * out_bstride + out_cstride + p] = v.y;
    }
}
// gather variant of the above: [B,2,*] with batch/channel strides -> float2 [B,hw]
__global__ void nchw2_strided_to_c2_kernel(const float* __restrict__ in, float2* __restrict__ out, int hw, size_t total,
                                           size_t in_bstride, size_t in_cstride, unsigned* amax_out) {
    float lmax = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        const float2 v = make_float2(in[b * in_bstride + p], in[b * in_bstride + in_cstride + p]);
        out[i] = v;
        lmax = fmaxf(lmax, fmaxf(fabsf(v.x), fabsf(v.y)));
    }
    publish_amax(amax_out, lmax);
}
// arbitrary 4-D strides (source maps arrive as permuted views, hybridnet.py:152,167)
__global__ void src_strided_to_c2_kernel(const float* __restrict__ in, float2* __restrict__ out, int n, size_t total,
                                         long long sb, long long sc, long long sh, long long sw) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)n * n), p = i - b * (size_t)n * n;
        const long long y = (long long)(p / n), x = (long long)(p % n);
        const long long o = (long long)b * sb + y * sh + x * sw;
        out[i] = make_float2(in[o], in[o + sc]);
    }
}
// Monochromatic point sources, one map per location (SourceModule.make_abs_spatial_map + spatial_map, source_module.py:41-116;
// IterativeSolver.set_multiple_sources, hybridnet.py:161-170), written as the NCHW tensor [S,2,n,n] the reference builds.
// The reference forms |ifft2(ifftshift(fftshift(fft2(delta)) * W))| with W = 1 or the outer product of two (periodic)
// Blackman windows.  The shifted window is 0.42 + 0.5 cos(2 pi k/n) + 0.08 cos(4 pi k/n) over the frequency index k, whose
// inverse DFT is the 5-tap kernel g = [0.04, 0.25, 0.42, 0.25, 0.04], so the map is amplitude * g(y - r) * g(x - c) with
// periodic wrap (a delta without smoothing) -- no transform needed.  real = |map| cos(arg), imag = |map| sin(arg).
__global__ void point_sources_kernel(const int* __restrict__ loc, int count, int n, float amplitude, float cos_arg, float sin_arg,
                                     int smooth, float* __restrict__ out) {
    const size_t hw = (size_t)n * n, total = (size_t)count * hw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t s = i / hw, p = i - s * hw;
        const int y = (int)(p / n), x = (int)(p % n);
        int dy = y - loc[2 * s], dx = x - loc[2 * s + 1];
        dy = ((dy % n) + n) % n;
        dx = ((dx % n) + n) % n;
        if (dy > n / 2) dy = n - dy;
        if (dx > n / 2) dx = n - dx;
        float gy, gx;
        if (smooth) {
            gy = dy == 0 ? 0.42f : dy == 1 ? 0.25f : dy == 2 ? 0.04f : 0.f;
            gx = dx == 0 ? 0.42f : dx == 1 ? 0.25f : dx == 2 ? 0.04f : 0.f;
        } else {
            gy = dy == 0 ? 1.f : 0.f;
            gx = dx == 0 ? 1.f : 0.f;
        }
        const float v = fabsf(amplitude) * gy * gx;
        out[(s * 2) * hw + p] = v * cos_arg;
        out[(s * 2 + 1) * hw + p] = v * sin_arg;
    }
}
// get_initials + initial residual: k_sq = (omega/sos)^2, wf = 0, res = L(0) + k_sq*0 - source = -source
__global__ void reset_kernel(const float* __restrict__ sos, float* __restrict__ ksq, float2* __restrict__ wf,
                             float2* __restrict__ res, const float2* __restrict__ src, int src_batch, float omega, int hw,
                             size_t total, unsigned* amax_res) {
    float lmax = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        const float q = omega / sos[i];
        ksq[i] = q * q;
        wf[i] = make_float2(0.f, 0.f);
        const float2 s = src[(src_batch > 1 ? b * hw : 0) + p];
        res[i] = make_float2(0.f - s.x, 0.f - s.y);
        lmax = fmaxf(lmax, fmaxf(fabsf(s.x), fabsf(s.y)));
    }
    publish_amax(amax_res, lmax);
}
// [B,6,H,W] -> NHWC8 (channels 6,7 zero): input of HybridNet.forward when called directly
__global__ void nchw6_to_nhwc8_kernel(const float* __restrict__ in, float* __restrict__ out, int hw, size_t total,
                                      unsigned* amax_out) {
    float lmax = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        float v[8];
#pragma unroll
        for (int c = 0; c < 6; c++) {
            v[c] = in[(b * 6 + c) * hw + p];
            lmax = fmaxf(lmax, fabsf(v[c]));
        }
        v[6] = v[7] = 0.f;
        float4* o = reinterpret_cast<float4*>(out + i * 8);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    publish_amax(amax_out, lmax);
}
// NHWC8 -> [B,8,H,W] (debug taps)
__global__ void nhwc8_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int hw, size_t total) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
#pragma unroll
        for (int c = 0; c < 8; c++) out[(b * 8 + c) * hw + p] = in[i * 8 + c];
    }
}
__global__ void finalize_rmse_kernel(const double* __restrict__ ssq, float* __restrict__ rmse, int count, double inv_count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        rmse[i] = (float)sqrt(ssq[i] * inv_count);
}
// zero the amax slots whose bit is set in `mask` (slots of tensors that are re-produced in this UNet pass)
// First kernel of an iteration: it also advances the iteration slot of the residual-norm history when `it` is given (the slot
// starts at -1 in hn_run), which saves a one-thread kernel at the end of every iteration.
__global__ void reset_amax_kernel(unsigned* slots, unsigned long long mask, int* it) {
    const int i = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (i < 64 && ((mask >> i) & 1ull)) slots[i] = 0u;
    if (i == 0 && it != nullptr) *it += 1;
}

}  // namespace hn
