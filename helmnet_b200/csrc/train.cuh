// train.cuh -- backward kernels of ONE solver step: the training unroll of the reference
// (IterativeSolver.n_steps under autograd, helmnet/hybridnet.py:586-623, driven by training_step :385-410).
//
// SURVEY.md section 8 row f4.  hn_step_backward (train_host.cuh) recomputes the step in fp32 with every pre-activation kept,
// then walks the UNet (helmnet/architectures.py:439-465, 240-252, 63-84) and the residual (hybridnet.py:544-584) backwards.
// Everything here runs on the fp32 CUDA cores, one thread per pixel (data gradients) or per weight (weight gradients):
// these kernels are the correct-first version of the row, not tuned like the inference path.
//
// Tensor layout: NHWC, contiguous, C = the real channel count (2, 6 or 8 floats per pixel) -- complex fields are float2.
#pragma once
#include "common.cuh"
#include "spectral.cuh"

namespace hn {
namespace tr {

constexpr int T_THREADS = 128;

// ---- k x k (k = 1, 3) stride-1 zero-padded correlation over up to two concatenated sources ------------------------
//   transposed = 0:  forward,   out[p][co] = bias[co] + sum_{tap,ci} in[p + tap][ci] * w[co][ci][tap]     (w = [CO][CI][k][k])
//   transposed = 1:  data gradient of the same layer, sources = dL/dz (the layer's C_out channels), CO = the layer's C_in:
//                    out[p][c] = sum_{tap,o} dz[p - tap][o] * w[o][c][tap]                                (w = [CI_src][CO][k][k])
// Output channels [0, c0) go to o0, [c0, CO) to o1 (the two halves of a torch.cat); acc*: add to what is there.
struct ConvArgs {
    const float* a; int ca;
    const float* b; int cb;
    const float* w;
    const float* bias;
    float in_scale;
    int ks, transposed;
    int H, W;
    long long P;
    float* o0; int c0, acc0;
    float* o1; int acc1;
    float slope; float* act;     // act != null: act = PReLU(out) next to the pre-activation in o0 (c0 == CO)
};

// acc[co] += scale * px[ci] * ws[ci][co] over the C channels of one source pixel (C is even; 128-bit loads when C % 4 == 0:
// a pixel of C floats is then 16-byte aligned)
template <int CO>
__device__ __forceinline__ void accumulate(float (&acc)[CO], const float* px, int C, float scale, const float* ws) {
    if ((C & 3) == 0) {
        for (int ci = 0; ci < C; ci += 4) {
            const float4 v = *reinterpret_cast<const float4*>(px + ci);
            const float vv[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int co = 0; co < CO; co++) acc[co] = fmaf(vv[j], ws[(ci + j) * CO + co], acc[co]);
        }
    } else {
        for (int ci = 0; ci < C; ci += 2) {
            const float2 v = *reinterpret_cast<const float2*>(px + ci);
            const float vv[2] = {v.x * scale, v.y * scale};
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int co = 0; co < CO; co++) acc[co] = fmaf(vv[j], ws[(ci + j) * CO + co], acc[co]);
        }
    }
}

template <int CO>
__global__ void __launch_bounds__(T_THREADS) conv_kernel(ConvArgs p) {
    HN_DYN_SMEM(float, wsm);
    const int CI = p.ca + p.cb, KK = p.ks * p.ks, pad = p.ks / 2;
    for (int i = threadIdx.x; i < KK * CI * CO; i += blockDim.x) {
        const int co = i % CO, ci = (i / CO) % CI, tap = i / (CO * CI);
        wsm[i] = p.transposed ? p.w[((size_t)ci * CO + co) * KK + (KK - 1 - tap)] : p.w[((size_t)co * CI + ci) * KK + tap];
    }
    __syncthreads();
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < p.P; pix += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(pix % p.W), y = (int)((pix / p.W) % p.H);
        float acc[CO];
#pragma unroll
        for (int co = 0; co < CO; co++) acc[co] = p.bias ? p.bias[co] : 0.f;
        for (int ky = 0; ky < p.ks; ky++) {
            const int yy = y + ky - pad;
            if (yy < 0 || yy >= p.H) continue;
            for (int kx = 0; kx < p.ks; kx++) {
                const int xx = x + kx - pad;
                if (xx < 0 || xx >= p.W) continue;
                const long long q = pix + (long long)(ky - pad) * p.W + (kx - pad);
                const float* ws = wsm + (size_t)(ky * p.ks + kx) * CI * CO;
                accumulate<CO>(acc, p.a + q * p.ca, p.ca, p.in_scale, ws);
                if (p.cb > 0) accumulate<CO>(acc, p.b + q * p.cb, p.cb, p.in_scale, ws + p.ca * CO);
            }
        }
        const int c1 = CO - p.c0;
#pragma unroll
        for (int co = 0; co < CO; co++) {
            if (co < p.c0) {
                float* d = p.o0 + pix * p.c0 + co;
                const float v = p.acc0 ? *d + acc[co] : acc[co];
                *d = v;
                if (p.act != nullptr) p.act[pix * p.c0 + co] = v >= 0.f ? v : p.slope * v;
            } else if (p.o1 != nullptr) {
                float* d = p.o1 + pix * c1 + (co - p.c0);
                *d = p.acc1 ? *d + acc[co] : acc[co];
            }
        }
    }
}

// ---- weight / bias gradient of the k x k stride-1 layer:  gw[o][c][tap] += sum_p dz[p][o] * in[p + tap][c] -----------
// A CTA walks tiles of 16 x 16 pixels (input tile with a one-pixel halo and the dz tile in shared memory).  A thread owns one
// (input channel, tap) pair -- plus one virtual all-ones channel for the bias -- and ALL output channels: per pixel one
// shared-memory load of the input value, the CO values of dz as a broadcast, CO FMAs.  Several groups of threads split the
// pixels of a tile; the partial sums stay in registers across all tiles of the CTA and reach memory as one atomicAdd per
// weight, group and CTA.
constexpr int WG_T = 16;          // pixels per tile edge
constexpr int WG_THREADS = 256;
constexpr int WG_MAX_THREADS = 512;
struct WgradArgs {
    const float* a; int ca;
    const float* b; int cb;
    const float* dz;
    float dz_scale;
    int ks, H, W, B;
    float* gw;      // [CO][ca + cb][ks][ks]
    float* gb;      // [CO] or null
};
// (input pixels are stored CI + 1 floats apart: with a pitch of 2, 8 or 16 floats the nine taps of one channel would share two banks)
__host__ __device__ inline size_t wgrad_smem_bytes(int ci, int co) {
    const size_t tiles = ((size_t)(WG_T + 2) * (WG_T + 2) * (ci + 1) + (size_t)WG_T * WG_T * co) * sizeof(float);
    const size_t red = (size_t)512 * co * sizeof(float);      // final reduction across the thread groups (WG_MAX_THREADS x CO)
    return tiles > red ? tiles : red;
}
// threads per CTA: as many whole groups of (ci * k^2 + 1) threads as fit into WG_MAX_THREADS, rounded up to whole warps
__host__ __device__ inline int wgrad_threads(int ci, int ks) {
    const int units = ci * ks * ks + 1;
    const int g = WG_MAX_THREADS / units > 0 ? WG_MAX_THREADS / units : 1;
    const int t = (units * g + 31) & ~31;
    return t > WG_MAX_THREADS ? WG_MAX_THREADS : t;
}
template <int CO>
__global__ void __launch_bounds__(WG_MAX_THREADS) wgrad_kernel(WgradArgs p) {
    HN_DYN_SMEM(float, sm);
    const int CI = p.ca + p.cb, KK = p.ks * p.ks, pad = p.ks / 2, TP = WG_T + 2, PS = CI + 1;
    float* in_s = sm;                               // [TP][TP][PS], halo of one pixel
    float* dz_s = sm + (size_t)TP * TP * PS;        // [WG_T][WG_T][CO]
    const int tiles_x = (p.W + WG_T - 1) / WG_T, tiles_y = (p.H + WG_T - 1) / WG_T;
    const int tiles = tiles_x * tiles_y * p.B;
    const int units = CI * KK + 1;                  // the last unit is the bias (input == 1)
    const int groups = max(1, (int)blockDim.x / units);
    const int unit = threadIdx.x % units, grp = threadIdx.x / units;
    const bool active = grp < groups && ((int)threadIdx.x < groups * units) && (unit < units - 1 || p.gb != nullptr);
    const bool is_bias = unit == units - 1;
    const int c = is_bias ? 0 : unit / KK, tap = is_bias ? 0 : unit - c * KK;
    const int oy = tap / p.ks - pad + 1, ox = tap % p.ks - pad + 1;
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; o++) acc[o] = 0.f;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int b = tile / (tiles_x * tiles_y), tt = tile - b * tiles_x * tiles_y;
        const int ty = tt / tiles_x, tx = tt - ty * tiles_x;
        const int y0 = ty * WG_T, x0 = tx * WG_T;
        const size_t img = (size_t)b * p.H * p.W;
        for (int i = threadIdx.x; i < TP * TP * CI; i += blockDim.x) {
            const int cc = i % CI, q = i / CI, px = q % TP, py = q / TP;
            const int y = y0 + py - 1, x = x0 + px - 1;
            float v = 0.f;
            if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
                const size_t g = img + (size_t)y * p.W + x;
                v = cc < p.ca ? p.a[g * p.ca + cc] : p.b[g * p.cb + (cc - p.ca)];
            }
            in_s[q * PS + cc] = v;
        }
        for (int i = threadIdx.x; i < WG_T * WG_T * CO; i += blockDim.x) {
            const int cc = i % CO, q = i / CO, px = q % WG_T, py = q / WG_T;
            const int y = y0 + py, x = x0 + px;
            dz_s[i] = (y < p.H && x < p.W) ? p.dz[(img + (size_t)y * p.W + x) * CO + cc] * p.dz_scale : 0.f;
        }
        __syncthreads();
        if (active) {
            for (int q = grp; q < WG_T * WG_T; q += groups) {
                const int py = q / WG_T, px = q - py * WG_T;
                const float v = is_bias ? 1.f : in_s[((py + oy) * TP + px + ox) * PS + c];
#pragma unroll
                for (int o = 0; o < CO; o++) acc[o] = fmaf(v, dz_s[q * CO + o], acc[o]);
            }
        }
        __syncthreads();
    }
    // the groups' partial sums meet in shared memory (the tiles are done with it): one atomicAdd per weight and CTA
    float* red = sm;                                // [groups][units][CO]
#pragma unroll
    for (int o = 0; o < CO; o++) red[threadIdx.x * CO + o] = active ? acc[o] : 0.f;
    __syncthreads();
    if (grp == 0 && active) {
#pragma unroll
        for (int o = 0; o < CO; o++) {
            float tot = 0.f;
            for (int g = 0; g < groups; g++) tot += red[(g * units + unit) * CO + o];
            if (is_bias) atomicAdd(p.gb + o, tot);
            else atomicAdd(p.gw + ((size_t)o * CI + c) * KK + tap, tot);
        }
    }
}

// ---- PReLU backward (one signed slope per layer, architectures.py:32-33):  dz = da * (z >= 0 ? 1 : slope),
//      gslope += sum da * min(z, 0).  In place (dz may alias da).  The slope gradient is a sum of B * r^2 * C signed terms that
//      largely cancel (it comes out 1e-3 .. 1e-5 of the sum of their magnitudes), so it is accumulated in double -- per thread,
//      per block and across blocks -- and folded into the fp32 gradient blob once at the end (fold_slopes_kernel).
__global__ void __launch_bounds__(256) prelu_bwd_kernel(const float* __restrict__ z, const float* da, float* dz, float slope,
                                                        double* gslope, size_t total) {
    __shared__ double red[256];
    double part = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float zz = z[i], g = da[i];
        dz[i] = zz >= 0.f ? g : slope * g;
        if (zz < 0.f) part += (double)g * (double)zz;
    }
    red[threadIdx.x] = part;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(gslope, red[0]);
}
struct SlopeFold { int n; int off[16]; };
__global__ void fold_slopes_kernel(const double* __restrict__ acc, float* gp, SlopeFold f) {
    const int i = threadIdx.x;
    if (i < f.n) gp[f.off[i]] += (float)acc[i];
}

// ---- the 8 x 8, stride 2, padding 3 pair (enc[d].down: Conv2d, up[d]: ConvTranspose2d; 8 -> 8 channels) -----------
// "gather":  out (small, Hs/2) from src (big, Hs):  out[o][c] = bias[c] + sum_{k,s} src[2o - 3 + k][s] * w[c][s][k]
//            = enc[d].down forward (w = [co][ci][8][8]) and the data gradient of up[d] (w = [ci][co][8][8], src = dL/d(up output)).
// "scatter": out (big, 2 Hs) from src (small, Hs):  out[y][c] = bias[c] + sum_{k: y + 3 - k even,s} src[(y + 3 - k)/2][s] * w[s][c][k]
//            = up[d] forward (w = [ci][co][8][8]) and the data gradient of enc[d].down (w = [co][ci][8][8], src = dL/d(down output)).
struct S2Args {
    const float* src; int Hs, Ws;
    float* out; int Ho, Wo;
    const float* w; const float* bias;
    int acc;
    long long P;     // batch * Ho * Wo
};
__global__ void __launch_bounds__(T_THREADS) s2_gather_kernel(S2Args p) {
    HN_DYN_SMEM(float, wsm);    // [k][s][c]
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
        const int c = i & 7, s = (i >> 3) & 7, k = i >> 6;
        wsm[i] = p.w[(c * 8 + s) * 64 + k];
    }
    __syncthreads();
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < p.P; pix += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(pix % p.Wo), oy = (int)((pix / p.Wo) % p.Ho);
        const long long img = pix / ((long long)p.Wo * p.Ho);
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; c++) acc[c] = p.bias ? p.bias[c] : 0.f;
        for (int ky = 0; ky < 8; ky++) {
            const int y = 2 * oy - 3 + ky;
            if (y < 0 || y >= p.Hs) continue;
            for (int kx = 0; kx < 8; kx++) {
                const int x = 2 * ox - 3 + kx;
                if (x < 0 || x >= p.Ws) continue;
                const float* sp = p.src + ((img * p.Hs + y) * p.Ws + x) * 8;
                const float* ws = wsm + (ky * 8 + kx) * 64;
                const float4 v0 = *reinterpret_cast<const float4*>(sp), v1 = *reinterpret_cast<const float4*>(sp + 4);
                const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int s = 0; s < 8; s++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[c] = fmaf(vv[s], ws[s * 8 + c], acc[c]);
            }
        }
        float* d = p.out + pix * 8;
#pragma unroll
        for (int c = 0; c < 8; c++) d[c] = p.acc ? d[c] + acc[c] : acc[c];
    }
}
__global__ void __launch_bounds__(T_THREADS) s2_scatter_kernel(S2Args p) {
    HN_DYN_SMEM(float, wsm);    // [k][s][c]
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
        const int c = i & 7, s = (i >> 3) & 7, k = i >> 6;
        wsm[i] = p.w[(s * 8 + c) * 64 + k];
    }
    __syncthreads();
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < p.P; pix += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(pix % p.Wo), y = (int)((pix / p.Wo) % p.Ho);
        const long long img = pix / ((long long)p.Wo * p.Ho);
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; c++) acc[c] = p.bias ? p.bias[c] : 0.f;
        for (int ky = (y + 3) & 1; ky < 8; ky += 2) {
            const int sy = (y + 3 - ky) / 2;
            if (y + 3 - ky < 0 || sy >= p.Hs) continue;
            for (int kx = (x + 3) & 1; kx < 8; kx += 2) {
                const int sx = (x + 3 - kx) / 2;
                if (x + 3 - kx < 0 || sx >= p.Ws) continue;
                const float* sp = p.src + ((img * p.Hs + sy) * p.Ws + sx) * 8;
                const float* ws = wsm + (ky * 8 + kx) * 64;
                const float4 v0 = *reinterpret_cast<const float4*>(sp), v1 = *reinterpret_cast<const float4*>(sp + 4);
                const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int s = 0; s < 8; s++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[c] = fmaf(vv[s], ws[s * 8 + c], acc[c]);
            }
        }
        float* d = p.out + pix * 8;
#pragma unroll
        for (int c = 0; c < 8; c++) d[c] = p.acc ? d[c] + acc[c] : acc[c];
    }
}
// weight gradient of both:  gw[a][b][k] += sum_i small[i][a] * big[2i - 3 + k][b]
//   enc[d].down: small = dL/d(output), big = input  -> [co][ci][8][8];   up[d]: small = input, big = dL/d(output) -> [ci][co][8][8]
// A CTA walks tiles of 8 x 8 small pixels (22 x 22 big pixels).  Thread t owns tap k = t % 64 and the channels a = t / 64 and
// a + 4 of the small tensor against all 8 channels of the big one: per small pixel two float4 loads of the big pixel (stored as
// two planes of float4 so that the 8 taps of a row fall on distinct banks), two loads of the small one, 16 FMAs; the 16 partial
// sums stay in registers across the CTA's tiles and reach memory as one atomicAdd per weight and CTA.
constexpr int SG_T = 8;           // small pixels per tile edge
constexpr int SG_BT = 2 * SG_T + 6;
struct S2WgradArgs {
    const float* small_t; const float* big_t;
    int Hs, Ws, B;   // resolution of `small_t`; `big_t` is 2 Hs x 2 Ws
    float* gw;
};
__host__ __device__ inline size_t s2_wgrad_smem_bytes() { return ((size_t)SG_BT * SG_BT * 8 + (size_t)SG_T * SG_T * 8) * sizeof(float); }
__global__ void __launch_bounds__(WG_THREADS) s2_wgrad_kernel(S2WgradArgs p) {
    HN_DYN_SMEM(float4, sm4);
    float4* big_s = sm4;                                                        // [2 planes][SG_BT * SG_BT]
    float* small_s = reinterpret_cast<float*>(sm4 + 2 * SG_BT * SG_BT);         // [SG_T * SG_T][8]
    const int tiles_x = (p.Ws + SG_T - 1) / SG_T, tiles_y = (p.Hs + SG_T - 1) / SG_T;
    const int tiles = tiles_x * tiles_y * p.B;
    const int Hb = 2 * p.Hs, Wb = 2 * p.Ws;
    const int k = threadIdx.x & 63, a0 = threadIdx.x >> 6, ky = k >> 3, kx = k & 7;
    float acc0[8], acc1[8];
#pragma unroll
    for (int b = 0; b < 8; b++) acc0[b] = acc1[b] = 0.f;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int img = tile / (tiles_x * tiles_y), tt = tile - img * tiles_x * tiles_y;
        const int ty = tt / tiles_x, tx = tt - ty * tiles_x;
        const int y0 = ty * SG_T, x0 = tx * SG_T;
        const size_t img_s = (size_t)img * p.Hs * p.Ws, img_b = (size_t)img * Hb * Wb;
        for (int i = threadIdx.x; i < 2 * SG_BT * SG_BT; i += blockDim.x) {
            const int pl = i / (SG_BT * SG_BT), q = i - pl * SG_BT * SG_BT, px = q % SG_BT, py = q / SG_BT;
            const int y = 2 * y0 - 3 + py, x = 2 * x0 - 3 + px;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < Hb && x >= 0 && x < Wb) v = *reinterpret_cast<const float4*>(p.big_t + (img_b + (size_t)y * Wb + x) * 8 + 4 * pl);
            big_s[i] = v;
        }
        for (int i = threadIdx.x; i < SG_T * SG_T * 8; i += blockDim.x) {
            const int cc = i & 7, q = i >> 3, px = q % SG_T, py = q / SG_T;
            const int y = y0 + py, x = x0 + px;
            small_s[i] = (y < p.Hs && x < p.Ws) ? p.small_t[(img_s + (size_t)y * p.Ws + x) * 8 + cc] : 0.f;
        }
        __syncthreads();
        for (int iy = 0; iy < SG_T; iy++)
            for (int ix = 0; ix < SG_T; ix++) {
                const int q = (2 * iy + ky) * SG_BT + 2 * ix + kx;
                const float4 b0 = big_s[q], b1 = big_s[SG_BT * SG_BT + q];
                const float s0 = small_s[(iy * SG_T + ix) * 8 + a0], s1 = small_s[(iy * SG_T + ix) * 8 + a0 + 4];
                acc0[0] = fmaf(s0, b0.x, acc0[0]); acc0[1] = fmaf(s0, b0.y, acc0[1]); acc0[2] = fmaf(s0, b0.z, acc0[2]); acc0[3] = fmaf(s0, b0.w, acc0[3]);
                acc0[4] = fmaf(s0, b1.x, acc0[4]); acc0[5] = fmaf(s0, b1.y, acc0[5]); acc0[6] = fmaf(s0, b1.z, acc0[6]); acc0[7] = fmaf(s0, b1.w, acc0[7]);
                acc1[0] = fmaf(s1, b0.x, acc1[0]); acc1[1] = fmaf(s1, b0.y, acc1[1]); acc1[2] = fmaf(s1, b0.z, acc1[2]); acc1[3] = fmaf(s1, b0.w, acc1[3]);
                acc1[4] = fmaf(s1, b1.x, acc1[4]); acc1[5] = fmaf(s1, b1.y, acc1[5]); acc1[6] = fmaf(s1, b1.z, acc1[6]); acc1[7] = fmaf(s1, b1.w, acc1[7]);
            }
        __syncthreads();
    }
#pragma unroll
    for (int b = 0; b < 8; b++) {
        atomicAdd(p.gw + ((size_t)a0 * 8 + b) * 64 + k, acc0[b]);
        atomicAdd(p.gw + ((size_t)(a0 + 4) * 8 + b) * 64 + k, acc1[b]);
    }
}

// per-channel sum (bias gradient of the stride-2 layers): gb[c] += sum_p t[p][c]
__global__ void __launch_bounds__(256) chan_sum_kernel(const float* __restrict__ t, int C, size_t P, float* gb) {
    __shared__ float red[8][8];
    float part[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x)
        for (int c = 0; c < C; c++) part[c] += t[i * C + c];
    for (int c = 0; c < C; c++) {
        const float s = warp_sum(part[c]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = s;
    }
    __syncthreads();
    if ((int)threadIdx.x < C) {
        float tot = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += red[w][threadIdx.x];
        atomicAdd(gb + threadIdx.x, tot);
    }
}

// ---- network input / output of the step (hybridnet.py:561-570) ------------------------------------------------------
// in6 = cat[wf, 1e3 * residual, sigma_x, sigma_y]  (sigma_x varies along W, sigma_y along H: spectral.py:307-312, hybridnet.py:126-131)
__global__ void make_in6_kernel(const float2* __restrict__ wf, const float2* __restrict__ res, const float* __restrict__ sigma1d,
                                float* __restrict__ in6, int n, size_t total) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % n), y = (int)((i / n) % n);
        const float2 u = wf[i], r = res[i];
        float* d = in6 + i * 6;
        d[0] = u.x; d[1] = u.y; d[2] = 1e3f * r.x; d[3] = 1e3f * r.y; d[4] = sigma1d[x]; d[5] = sigma1d[y];
    }
}
// gradients of the step's inputs, written in the reference's NCHW layout:
//   g_wf = G + g_in6[0:2]   (wf feeds the update wf + out/1e3 and the network),   g_res = 1e3 * g_in6[2:4]
__global__ void step_input_grads_kernel(const float2* __restrict__ G, const float* __restrict__ gin6, float* __restrict__ g_wf,
                                        float* __restrict__ g_res, int hw, size_t total) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        const float2 g = G[i];
        const float* s = gin6 + i * 6;
        if (g_wf != nullptr) {
            g_wf[(b * 2) * hw + p] = g.x + s[0];
            g_wf[(b * 2 + 1) * hw + p] = g.y + s[1];
        }
        if (g_res != nullptr) {
            g_res[(b * 2) * hw + p] = 1e3f * s[2];
            g_res[(b * 2 + 1) * hw + p] = 1e3f * s[3];
        }
    }
}

// ---- adjoint of the spectral operator (spectral.py:31-79):  one axis  M = diag(a) D1 + diag(b) D2  with
//      D1 = F^-1 diag(i k) F,  D2 = F^-1 diag(-k^2) F.  The transpose of a complex-linear map on (re, im) pairs is its
//      Hermitian adjoint:  M^H g = D1^H (conj(a) g) + D2^H (conj(b) g) = F^-1 [ (-i k) F(conj(a) g) + (-k^2) F(conj(b) g) ].
// Lines in A (pitch lp); A, B, C are line buffers; returns the buffer holding conj(M^H g).
__device__ inline float2* axis_adjoint(float2* A, float2* B, float2* C, int nl, int lp, const float2* tw, const SpecTables& t) {
    const int n = t.n, pml = t.pml;
    for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
        const int l = it / n, j = it - l * n;
        B[l * lp + pidx(j)] = cmul(cconj(__ldg(t.b + j)), A[l * lp + pidx(j)]);
    }
    __syncthreads();
    float2* F2 = fft_lines(B, C, nl, lp, tw, t);
    float2* O = (F2 == B) ? C : B;
    if (pml > 0) {
        for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
            const int l = it / n, j = it - l * n;
            const bool strip = j < pml || j >= n - pml;      // a == 0 outside the PML strips
            O[l * lp + pidx(j)] = strip ? cmul(cconj(__ldg(t.a + j)), A[l * lp + pidx(j)]) : make_float2(0.f, 0.f);
        }
        __syncthreads();
        float2* F1 = fft_lines(O, A, nl, lp, tw, t);
        float2* T = (F1 == O) ? A : O;
        for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
            const int l = it / n, k = it - l * n;
            const float2 v1 = F1[l * lp + pidx(k)], v2 = F2[l * lp + pidx(k)];
            const float mk = __ldg(t.mk + k), ms = __ldg(t.msq + k);
            // conj( (-i k / n) v1 + (-k^2 / n) v2 )
            T[l * lp + pidx(k)] = make_float2(fmaf(mk, v1.y, ms * v2.x), fmaf(mk, v1.x, -ms * v2.y));
        }
        __syncthreads();
        return fft_lines(T, F1, nl, lp, tw, t);
    }
    for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
        const int l = it / n, k = it - l * n;
        const float2 v2 = F2[l * lp + pidx(k)];
        const float ms = __ldg(t.msq + k);
        O[l * lp + pidx(k)] = make_float2(ms * v2.x, -ms * v2.y);
    }
    __syncthreads();
    return fft_lines(O, A, nl, lp, tw, t);
}

__global__ void __launch_bounds__(SPEC_THREADS) spectral_rows_adj_kernel(SpecTables t, const float2* __restrict__ g,
                                                                         float2* __restrict__ rx, int total_rows, int L) {
    HN_DYN_SMEM(float2, smem_sp);
    const int n = t.n, lp = line_pitch(n);
    float2* tw = smem_sp;
    float2* A = tw + n;
    float2* B = A + L * lp;
    float2* C = B + L * lp;
    const int row0 = blockIdx.x * L;
    const int nl = min(L, total_rows - row0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = __ldg(t.tw + i);
    for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
        const int l = it / n, j = it - l * n;
        A[l * lp + pidx(j)] = g[(size_t)(row0 + l) * n + j];
    }
    __syncthreads();
    const float2* E = axis_adjoint(A, B, C, nl, lp, tw, t);
    for (int it = threadIdx.x; it < nl * n; it += blockDim.x) {
        const int l = it / n, j = it - l * n;
        rx[(size_t)(row0 + l) * n + j] = cconj(E[l * lp + pidx(j)]);
    }
}
// out = rows part + column part + k_sq * g + add   (the gradient of  r = L u + k_sq u - source  with respect to u, plus `add`)
__global__ void __launch_bounds__(SPEC_THREADS) spectral_cols_adj_kernel(SpecTables t, const float2* __restrict__ g,
                                                                         const float2* __restrict__ rx, const float* __restrict__ ksq,
                                                                         const float2* __restrict__ add, float2* __restrict__ out, int CW) {
    HN_DYN_SMEM(float2, smem_sp);
    const int n = t.n, lp = line_pitch(n);
    float2* tw = smem_sp;
    float2* A = tw + n;
    float2* B = A + CW * lp;
    float2* C = B + CW * lp;
    const int b = blockIdx.y, j0 = blockIdx.x * CW;
    const int nc = min(CW, n - j0);
    const size_t img = (size_t)b * n * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = __ldg(t.tw + i);
    for (int it = threadIdx.x; it < n * CW; it += blockDim.x) {
        const int i = it / CW, c = it - i * CW;
        if (c < nc) A[c * lp + pidx(i)] = g[img + (size_t)i * n + j0 + c];
    }
    __syncthreads();
    const float2* E = axis_adjoint(A, B, C, nc, lp, tw, t);
    for (int it = threadIdx.x; it < n * CW; it += blockDim.x) {
        const int i = it / CW, c = it - i * CW;
        if (c >= nc) continue;
        const size_t p = img + (size_t)i * n + j0 + c;
        float2 r = cadd(rx[p], cconj(E[c * lp + pidx(i)]));
        if (ksq != nullptr) {
            const float kq = ksq[p];
            const float2 gg = g[p];
            r.x = fmaf(kq, gg.x, r.x);
            r.y = fmaf(kq, gg.y, r.y);
        }
        if (add != nullptr) r = cadd(r, add[p]);
        out[p] = r;
    }
}

}  // namespace tr
}  // namespace hn
