// conv_simt.cuh -- fp32 CUDA-core convolution kernels of the HybridNet learned optimizer.
//
// Reference semantics: helmnet/architectures.py
//   DoubleConv            :63-84    conv3x3(p1) -> PReLU(one signed slope) -> conv3x3(p1)
//   EncoderBlock.forward  :240-252  cat[x,state] / cat[out,state] inputs, Conv2d(8,8,k8,s2,p3) down-sampling
//   HybridNet.forward     :439-465  ConvTranspose2d(8,8,k8,s2,p3) up-sampling, cat[up,skip], 1x1 outc
// and the wavefield update of IterativeSolver.single_step (helmnet/hybridnet.py:564-570).
//
// Data layout in HBM: activations are NHWC fp32 with C=8 (32 B per pixel, two float4) or C=2
// (wavefield / residual / hidden state as interleaved float2 = complex).  Concatenations are never
// materialised: a kernel reads its (up to two) sources straight into shared-memory channel planes.
//
// Shared-memory tile layout: plane-major float4, tile[plane][y][x], plane = 4 consecutive input channels
// of the (virtual) concatenated input.  Consecutive lanes read consecutive pixels -> conflict-free
// LDS.128; the weights of one (plane, tap) are 4ci x COUT floats read as warp-uniform broadcasts.
// Each thread owns one column x and RP consecutive rows and keeps RP x COUT accumulators in registers,
// updated with packed FFMA2.
#pragma once
#include "common.cuh"

namespace hn {

enum ConvSrc : int {
    SRC_INC = 0,    // planes: [wf.re wf.im 1e3*r.re 1e3*r.im] [sigma_x sigma_y 0 0]   (hybridnet.py:564-566)
    SRC_A8 = 1,     // A: NHWC8
    SRC_A8_B2 = 2,  // A: NHWC8, B: float2           (cat[x, state], cat[out, state])
    SRC_A8_B8 = 3,  // A: NHWC8, B: NHWC8            (cat[up, skip])
    SRC_A2 = 4      // A: float2                     (second conv of conv_state)
};
enum ConvEpi : int {
    EPI_STORE = 0,  // bias (+PReLU) -> NHWC store
    EPI_OUTC = 1,   // bias -> 1x1 outc (8->2) -> wf += out/1e3   (or raw out when dwf_out != nullptr)
    EPI_STORE2 = 2  // tcgen05 kernels only: C_out = 2 layer padded to 8 columns, store channels 0,1 as float2
};

__host__ __device__ constexpr int src_planes(int src) {
    return src == SRC_INC ? 2 : src == SRC_A8 ? 2 : src == SRC_A8_B2 ? 3 : src == SRC_A8_B8 ? 4 : 1;
}

struct Conv3Args {
    const float* inA;
    const float* inB;
    const float* sigma;   // SRC_INC: 1-D PML sigma profile [W]
    const float* w;       // packed [planes][9 taps][4 ci][COUT]
    const float* bias;    // [COUT]
    const float* bias8;   // bias zero-padded to 8 (tcgen05 C_out = 2 path)
    const float* slope;   // PReLU slope (device scalar) when PRELU
    float* out;           // NHWC COUT
    const float* wo;      // EPI_OUTC: outc weight [2][8]
    const float* bo;      // EPI_OUTC: outc bias [2]
    float* wf;            // EPI_OUTC: wavefield float2 [B][H][W], updated in place
    float* dwf_out;       // EPI_OUTC: when non-null store the raw network output here instead
    unsigned* amax_out;         // running max |out| slot (see publish_amax), or null
    const unsigned* amax_in0;   // tcgen05 engine only: max |x| slots of the two sources
    const unsigned* amax_in1;
    const void* tcr_bmat; // row-streaming tcgen05 kernel: fp16 split-weight image (N = 48)
    const void* tc_bmat;  // tcgen05 engine only: fp16 split-weight image of this layer (unused by the SIMT kernel)
    float tc_inv;         // tcgen05 engine only: 2^-kw
    int H, W;
};

constexpr int C3_TX = 32;            // tile width  (one warp = 32 columns)
constexpr int C3_RP = 8;             // rows per thread
constexpr int C3_WARPS = 4;
constexpr int C3_TY = C3_RP * C3_WARPS;
constexpr int C3_PITCH = C3_TX + 2;  // 34
constexpr int C3_PLANE = C3_PITCH * (C3_TY + 2);  // 1156 float4; 18496 B = 64 (mod 128) -> planes offset banks
constexpr int C3_THREADS = 32 * C3_WARPS;

__host__ __device__ constexpr size_t conv3_smem_bytes(int src, int cout) {
    return (size_t)src_planes(src) * C3_PLANE * 16 + (size_t)src_planes(src) * 9 * 4 * cout * 4;
}

template <int SRC, int COUT, bool PRELU, int EPI>
__global__ void __launch_bounds__(C3_THREADS) conv3x3_kernel(Conv3Args a) {
    constexpr int NPL = src_planes(SRC);
    constexpr int CP = COUT / 2;  // accumulator pairs
    HN_DYN_SMEM(float4, smem_c3);
    float4* tile = smem_c3;
    float* wsm = reinterpret_cast<float*>(smem_c3 + NPL * C3_PLANE);

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * C3_TX, ty0 = blockIdx.y * C3_TY, b = blockIdx.z;
    const int H = a.H, W = a.W;

    // ---- weights -> smem -------------------------------------------------------------------------
    {
        constexpr int NW4 = NPL * 9 * 4 * COUT / 4;
        const float4* wg = reinterpret_cast<const float4*>(a.w);
        float4* ws4 = reinterpret_cast<float4*>(wsm);
        for (int i = tid; i < NW4; i += C3_THREADS) ws4[i] = __ldg(wg + i);
    }
    pdl_wait();      // constant weights above; activations written by earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    // ---- input tile (zero padded) -> smem planes --------------------------------------------------
    const size_t img = (size_t)b * H * W;
    if (SRC == SRC_A8 || SRC == SRC_A8_B2 || SRC == SRC_A8_B8) {
        for (int i = tid; i < C3_PLANE * 2; i += C3_THREADS) {
            const int px = i >> 1, half = i & 1;
            const int y = px / C3_PITCH, x = px - y * C3_PITCH;
            const int gy = ty0 - 1 + y, gx = tx0 - 1 + x;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = ldg4(a.inA + (img + (size_t)gy * W + gx) * 8 + half * 4);
            tile[half * C3_PLANE + px] = v;
        }
    }
    if (SRC == SRC_A8_B8) {
        for (int i = tid; i < C3_PLANE * 2; i += C3_THREADS) {
            const int px = i >> 1, half = i & 1;
            const int y = px / C3_PITCH, x = px - y * C3_PITCH;
            const int gy = ty0 - 1 + y, gx = tx0 - 1 + x;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = ldg4(a.inB + (img + (size_t)gy * W + gx) * 8 + half * 4);
            tile[(2 + half) * C3_PLANE + px] = v;
        }
    }
    if (SRC == SRC_A8_B2 || SRC == SRC_A2) {
        const float* src2 = (SRC == SRC_A2) ? a.inA : a.inB;
        constexpr int PL = (SRC == SRC_A2) ? 0 : 2;
        for (int px = tid; px < C3_PLANE; px += C3_THREADS) {
            const int y = px / C3_PITCH, x = px - y * C3_PITCH;
            const int gy = ty0 - 1 + y, gx = tx0 - 1 + x;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const float2 s = ldg2(src2 + (img + (size_t)gy * W + gx) * 2);
                v.x = s.x;
                v.y = s.y;
            }
            tile[PL * C3_PLANE + px] = v;
        }
    }
    if (SRC == SRC_INC) {
        for (int px = tid; px < C3_PLANE; px += C3_THREADS) {
            const int y = px / C3_PITCH, x = px - y * C3_PITCH;
            const int gy = ty0 - 1 + y, gx = tx0 - 1 + x;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const float2 u = ldg2(a.inA + (img + (size_t)gy * W + gx) * 2);
                const float2 r = ldg2(a.inB + (img + (size_t)gy * W + gx) * 2);
                v0 = make_float4(u.x, u.y, 1e3f * r.x, 1e3f * r.y);   // hybridnet.py:566  1e3 * residual
                v1 = make_float4(__ldg(a.sigma + gx), __ldg(a.sigma + gy), 0.f, 0.f);  // sigma_x[i,j]=s[j], sigma_y[i,j]=s[i]
            }
            tile[px] = v0;
            tile[C3_PLANE + px] = v1;
        }
    }
    __syncthreads();

    // ---- direct convolution from smem ---------------------------------------------------------------
    const int tx = tid & 31, wy = (tid >> 5) * C3_RP;
    float2 acc[C3_RP][CP];
    {
        float2 bv[CP];
#pragma unroll
        for (int c = 0; c < CP; c++) bv[c] = make_float2(__ldg(a.bias + 2 * c), __ldg(a.bias + 2 * c + 1));
#pragma unroll
        for (int r = 0; r < C3_RP; r++)
#pragma unroll
            for (int c = 0; c < CP; c++) acc[r][c] = bv[c];
    }
#pragma unroll 1
    for (int pl = 0; pl < NPL; pl++) {
#pragma unroll 1
        for (int tap = 0; tap < 9; tap++) {
            const int dy = tap / 3, dx = tap - dy * 3;
            float2 wr[4][CP];
            {
                const float2* wp = reinterpret_cast<const float2*>(wsm + ((pl * 9 + tap) * 4) * COUT);
#pragma unroll
                for (int ci = 0; ci < 4; ci++)
#pragma unroll
                    for (int c = 0; c < CP; c++) wr[ci][c] = wp[ci * CP + c];
            }
            const float4* tp = tile + pl * C3_PLANE + (wy + dy) * C3_PITCH + tx + dx;
#pragma unroll
            for (int r = 0; r < C3_RP; r++) {
                const float4 v = tp[r * C3_PITCH];
#pragma unroll
                for (int c = 0; c < CP; c++) ffma2(acc[r][c], v.x, wr[0][c]);
#pragma unroll
                for (int c = 0; c < CP; c++) ffma2(acc[r][c], v.y, wr[1][c]);
#pragma unroll
                for (int c = 0; c < CP; c++) ffma2(acc[r][c], v.z, wr[2][c]);
#pragma unroll
                for (int c = 0; c < CP; c++) ffma2(acc[r][c], v.w, wr[3][c]);
            }
        }
    }

    // ---- epilogue ----------------------------------------------------------------------------------
    const int gx = tx0 + tx;
    const bool colok = gx < W;
    float slope = 0.f;
    if (PRELU) slope = __ldg(a.slope);
    if constexpr (EPI == EPI_STORE) {
        float lmax = 0.f;
#pragma unroll
        for (int r = 0; r < C3_RP; r++) {
            const int gy = ty0 + wy + r;
            if (!colok || gy >= H) continue;
            float o[COUT];
#pragma unroll
            for (int c = 0; c < CP; c++) {
                o[2 * c] = acc[r][c].x;
                o[2 * c + 1] = acc[r][c].y;
            }
            if (PRELU) {
#pragma unroll
                for (int c = 0; c < COUT; c++) o[c] = o[c] >= 0.f ? o[c] : slope * o[c];  // signed slope
            }
#pragma unroll
            for (int c = 0; c < COUT; c++) lmax = fmaxf(lmax, fabsf(o[c]));
            float* dst = a.out + (img + (size_t)gy * W + gx) * COUT;
            if constexpr (COUT == 8) {
                reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
            } else {
                reinterpret_cast<float2*>(dst)[0] = make_float2(o[0], o[1]);
            }
        }
        publish_amax(a.amax_out, lmax);
    } else {  // EPI_OUTC (COUT == 8)
        float wo0[8], wo1[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            wo0[c] = __ldg(a.wo + c);
            wo1[c] = __ldg(a.wo + 8 + c);
        }
        const float bo0 = __ldg(a.bo), bo1 = __ldg(a.bo + 1);
        float lmax = 0.f;
#pragma unroll
        for (int r = 0; r < C3_RP; r++) {
            const int gy = ty0 + wy + r;
            if (!colok || gy >= H) continue;
            float o0 = bo0, o1 = bo1;
#pragma unroll
            for (int c = 0; c < CP; c++) {
                o0 = fmaf(acc[r][c].x, wo0[2 * c], o0);
                o0 = fmaf(acc[r][c].y, wo0[2 * c + 1], o0);
                o1 = fmaf(acc[r][c].x, wo1[2 * c], o1);
                o1 = fmaf(acc[r][c].y, wo1[2 * c + 1], o1);
            }
            const size_t p = img + (size_t)gy * W + gx;
            if (a.dwf_out != nullptr) {
                reinterpret_cast<float2*>(a.dwf_out)[p] = make_float2(o0, o1);
            } else {
                float2* wfp = reinterpret_cast<float2*>(a.wf) + p;
                const float2 u = *wfp;
                const float2 nw = make_float2(o0 / 1e3f + u.x, o1 / 1e3f + u.y);  // hybridnet.py:570  d_wavefield / 1e3 + wavefield
                *wfp = nw;
                lmax = fmaxf(lmax, fmaxf(fabsf(nw.x), fabsf(nw.y)));
            }
        }
        publish_amax(a.amax_out, lmax);
    }
}

// ------------------------------------------------------------------------------------------------------
// Down-sampling: Conv2d(8, 8, kernel 8, stride 2, padding 3)   (architectures.py:209-211)
//   out[o] = b + sum_k in[2o - 3 + k] * w[k],  k = 0..7 per axis.
// Input tile is stored split by x parity so that a warp of 32 consecutive outputs reads consecutive float4.
// ------------------------------------------------------------------------------------------------------
struct DownArgs {
    const float* in;    // NHWC8 [B][H][W]
    const float* w;     // packed [2 planes][8 ky][8 kx][4 ci][8 co]
    const float* bias;  // [8]
    float* out;         // NHWC8 [B][H/2][W/2]
    unsigned* amax_out; // running max |out| slot or null
    int H, W;           // input resolution
};
constexpr int DN_TX = 32, DN_RP = 4, DN_WARPS = 4, DN_TY = DN_RP * DN_WARPS;  // output tile 32 x 16
constexpr int DN_IW = 2 * DN_TX + 6, DN_IH = 2 * DN_TY + 6;                    // input tile 70 x 38
constexpr int DN_XH = DN_IW / 2;                                               // 35
constexpr int DN_PLANE = DN_IH * DN_XH;                                        // per (plane, parity)
constexpr int DN_THREADS = 32 * DN_WARPS;
constexpr size_t DN_SMEM = (size_t)4 * DN_PLANE * 16 + 2 * 64 * 4 * 8 * 4;

__global__ void __launch_bounds__(DN_THREADS) down_kernel(DownArgs a) {
    HN_DYN_SMEM(float4, smem_dn);
    float4* tile = smem_dn;  // [(plane*2 + xpar)][y][xh]
    float* wsm = reinterpret_cast<float*>(smem_dn + 4 * DN_PLANE);
    const int tid = threadIdx.x;
    const int H = a.H, W = a.W, Ho = H >> 1, Wo = W >> 1;
    const int ox0 = blockIdx.x * DN_TX, oy0 = blockIdx.y * DN_TY, b = blockIdx.z;
    const int gx0 = 2 * ox0 - 3, gy0 = 2 * oy0 - 3;
    {
        const float4* wg = reinterpret_cast<const float4*>(a.w);
        float4* ws4 = reinterpret_cast<float4*>(wsm);
        for (int i = tid; i < 2 * 64 * 4 * 8 / 4; i += DN_THREADS) ws4[i] = __ldg(wg + i);
    }
    pdl_wait();      // constant weights above; activations written by earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    const size_t img = (size_t)b * H * W;
    for (int i = tid; i < DN_IW * DN_IH * 2; i += DN_THREADS) {
        const int px = i >> 1, half = i & 1;
        const int ly = px / DN_IW, lx = px - ly * DN_IW;
        const int gy = gy0 + ly, gx = gx0 + lx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = ldg4(a.in + (img + (size_t)gy * W + gx) * 8 + half * 4);
        tile[(half * 2 + (lx & 1)) * DN_PLANE + ly * DN_XH + (lx >> 1)] = v;
    }
    __syncthreads();

    const int tx = tid & 31, wy = (tid >> 5) * DN_RP;
    float2 acc[DN_RP][4];
    {
        float2 bv[4];
#pragma unroll
        for (int c = 0; c < 4; c++) bv[c] = make_float2(__ldg(a.bias + 2 * c), __ldg(a.bias + 2 * c + 1));
#pragma unroll
        for (int r = 0; r < DN_RP; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = bv[c];
    }
#pragma unroll 1
    for (int pl = 0; pl < 2; pl++) {
#pragma unroll 1
        for (int ky = 0; ky < 8; ky++) {
#pragma unroll 2
            for (int kx = 0; kx < 8; kx++) {
                float2 wr[4][4];
                const float2* wp = reinterpret_cast<const float2*>(wsm + (((pl * 8 + ky) * 8 + kx) * 4) * 8);
#pragma unroll
                for (int ci = 0; ci < 4; ci++)
#pragma unroll
                    for (int c = 0; c < 4; c++) wr[ci][c] = wp[ci * 4 + c];
                const float4* tp = tile + (pl * 2 + (kx & 1)) * DN_PLANE + (2 * wy + ky) * DN_XH + tx + (kx >> 1);
#pragma unroll
                for (int r = 0; r < DN_RP; r++) {
                    const float4 v = tp[2 * r * DN_XH];
#pragma unroll
                    for (int c = 0; c < 4; c++) ffma2(acc[r][c], v.x, wr[0][c]);
#pragma unroll
                    for (int c = 0; c < 4; c++) ffma2(acc[r][c], v.y, wr[1][c]);
#pragma unroll
                    for (int c = 0; c < 4; c++) ffma2(acc[r][c], v.z, wr[2][c]);
#pragma unroll
                    for (int c = 0; c < 4; c++) ffma2(acc[r][c], v.w, wr[3][c]);
                }
            }
        }
    }
    const int ox = ox0 + tx;
    float lmax = 0.f;
#pragma unroll
    for (int r = 0; r < DN_RP; r++) {
        const int oy = oy0 + wy + r;
        if (ox >= Wo || oy >= Ho) continue;
        float4* dst = reinterpret_cast<float4*>(a.out + (((size_t)b * Ho + oy) * Wo + ox) * 8);
        dst[0] = make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
        dst[1] = make_float4(acc[r][2].x, acc[r][2].y, acc[r][3].x, acc[r][3].y);
#pragma unroll
        for (int c = 0; c < 4; c++) lmax = fmaxf(lmax, fmaxf(fabsf(acc[r][c].x), fabsf(acc[r][c].y)));
    }
    publish_amax(a.amax_out, lmax);
}

// ------------------------------------------------------------------------------------------------------
// Up-sampling: ConvTranspose2d(8, 8, kernel 8, stride 2, padding 3), weight (in, out, kh, kw)
// (architectures.py:373-385):  o = 2 i - 3 + k.  For output o = 2 o' + p (p = parity) the contributing
// inputs are i = o' + t - 2 + p with kernel tap k = 7 - 2 t - p, t = 0..3 -> 4 x 4 taps per output and four
// parity classes with their own 4x4x8x8 weight sets.  One warp = one parity class (warp-uniform weights).
// ------------------------------------------------------------------------------------------------------
struct UpArgs {
    const float* in;    // NHWC8 [B][Hi][Wi]
    const float* w;     // packed [4 classes][2 planes][4 ty][4 tx][4 ci][8 co]
    const float* bias;  // [8]
    float* out;         // NHWC8 [B][2Hi][2Wi]
    unsigned* amax_out; // running max |out| slot or null
    int Hi, Wi;
};
constexpr int UP_TL = 16;                 // low-res cells per tile side -> 32 x 32 outputs
constexpr int UP_RP = 8;
constexpr int UP_IW = UP_TL + 4;          // 20
constexpr int UP_PLANE = UP_IW * UP_IW;   // 400
constexpr int UP_THREADS = 128;
constexpr size_t UP_SMEM = (size_t)2 * UP_PLANE * 16 + 4 * 2 * 16 * 4 * 8 * 4;

__global__ void __launch_bounds__(UP_THREADS) up_kernel(UpArgs a) {
    HN_DYN_SMEM(float4, smem_up);
    float4* tile = smem_up;  // [plane][y][x]
    float* wsm = reinterpret_cast<float*>(smem_up + 2 * UP_PLANE);
    const int tid = threadIdx.x;
    const int Hi = a.Hi, Wi = a.Wi, Ho = 2 * Hi, Wo = 2 * Wi;
    const int cx0 = blockIdx.x * UP_TL, cy0 = blockIdx.y * UP_TL, b = blockIdx.z;
    {
        const float4* wg = reinterpret_cast<const float4*>(a.w);
        float4* ws4 = reinterpret_cast<float4*>(wsm);
        for (int i = tid; i < 4 * 2 * 16 * 4 * 8 / 4; i += UP_THREADS) ws4[i] = __ldg(wg + i);
    }
    pdl_wait();      // constant weights above; activations written by earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    const size_t img = (size_t)b * Hi * Wi;
    for (int i = tid; i < UP_PLANE * 2; i += UP_THREADS) {
        const int px = i >> 1, half = i & 1;
        const int ly = px / UP_IW, lx = px - ly * UP_IW;
        const int gy = cy0 - 2 + ly, gx = cx0 - 2 + lx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < Hi && gx >= 0 && gx < Wi) v = ldg4(a.in + (img + (size_t)gy * Wi + gx) * 8 + half * 4);
        tile[half * UP_PLANE + px] = v;
    }
    __syncthreads();

    const int cls = tid >> 5, py = cls >> 1, pxp = cls & 1;
    const int lane = tid & 31, cx = lane & 15, cyb = (lane >> 4) * UP_RP;
    float2 acc[UP_RP][4];
    {
        float2 bv[4];
#pragma unroll
        for (int c = 0; c < 4; c++) bv[c] = make_float2(__ldg(a.bias + 2 * c), __ldg(a.bias + 2 * c + 1));
#pragma unroll
        for (int r = 0; r < UP_RP; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = bv[c];
    }
#pragma unroll 1
    for (int pl = 0; pl < 2; pl++) {
#pragma unroll 1
        for (int ty = 0; ty < 4; ty++) {
#pragma unroll 1
            for (int tx = 0; tx < 4; tx++) {
                float2 wr[4][4];
                const float2* wp =
                    reinterpret_cast<const float2*>(wsm + ((((cls * 2 + pl) * 4 + ty) * 4 + tx) * 4) * 8);
#pragma unroll
                for (int ci = 0; ci < 4; ci++)
#pragma unroll
                    for (int c = 0; c < 4; c++) wr[ci][c] = wp[ci * 4 + c];
                const float4* tp = tile + pl * UP_PLANE + (cyb + ty + py) * UP_IW + cx + tx + pxp;
#pragma unroll
                for (int r = 0; r < UP_RP; r++) {
                    const float4 v = tp[r * UP_IW];
#pragma unroll
                    for (int c = 0; c < 4; c++) ffma2(acc[r][c], v.x, wr[0][c]);
#pragma unroll
                    for (int c = 0; c < 4; c++) ffma2(acc[r][c], v.y, wr[1][c]);
#pragma unroll
                    for (int c = 0; c < 4; c++) ffma2(acc[r][c], v.z, wr[2][c]);
#pragma unroll
                    for (int c = 0; c < 4; c++) ffma2(acc[r][c], v.w, wr[3][c]);
                }
            }
        }
    }
    const int ox = 2 * (cx0 + cx) + pxp;
    float lmax = 0.f;
#pragma unroll
    for (int r = 0; r < UP_RP; r++) {
        const int oy = 2 * (cy0 + cyb + r) + py;
        if (ox >= Wo || oy >= Ho) continue;
        float4* dst = reinterpret_cast<float4*>(a.out + (((size_t)b * Ho + oy) * Wo + ox) * 8);
        dst[0] = make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
        dst[1] = make_float4(acc[r][2].x, acc[r][2].y, acc[r][3].x, acc[r][3].y);
#pragma unroll
        for (int c = 0; c < 4; c++) lmax = fmaxf(lmax, fmaxf(fabsf(acc[r][c].x), fabsf(acc[r][c].y)));
    }
    publish_amax(a.amax_out, lmax);
}

// ------------------------------------------------------------------------------------------------------
// conv_state second convolution: Conv2d(2, 2, 3, padding 1) on the 2-channel hidden-state intermediate
// (architectures.py:213-217, second layer of conv_state's DoubleConv).  36 MACs and 16 bytes per pixel: purely
// bandwidth bound, so a lean kernel: float2 tile in shared memory, each thread a 2 x 4 pixel patch from a 4 x 6
// register window (3 shared-memory loads per output pixel instead of 9).
// ------------------------------------------------------------------------------------------------------
struct State2Args {
    const float* in;      // float2 [B][H][W]
    const float* w;       // [co 2][ci 2][3][3] as stored in the checkpoint
    const float* bias;    // [2]
    float* out;           // float2 [B][H][W]
    unsigned* amax_out;
    int H, W;
};
constexpr int S2_TX = 64, S2_TY = 32, S2_THREADS = 256;   // thread (tx 0..15, ty 0..15) -> pixels x = 4 tx.., y = 2 ty..
constexpr int S2_PITCH = S2_TX + 2;

__global__ void __launch_bounds__(S2_THREADS) state2_kernel(State2Args a) {
    __shared__ float2 tile[(S2_TY + 2) * S2_PITCH];
    __shared__ float wsm[36 + 2];
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * S2_TX, ty0 = blockIdx.y * S2_TY, b = blockIdx.z;
    const int H = a.H, W = a.W;
    const size_t img = (size_t)b * H * W;
    if (tid < 36) wsm[tid] = __ldg(a.w + tid);
    if (tid >= 64 && tid < 66) wsm[36 + tid - 64] = __ldg(a.bias + tid - 64);
    pdl_wait();      // constant weights above; activations written by earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    for (int i = tid; i < (S2_TY + 2) * S2_PITCH; i += S2_THREADS) {
        const int y = i / S2_PITCH, x = i - y * S2_PITCH;
        const int gy = ty0 - 1 + y, gx = tx0 - 1 + x;
        float2 v = make_float2(0.f, 0.f);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = ldg2(a.in + (img + (size_t)gy * W + gx) * 2);
        tile[i] = v;
    }
    __syncthreads();
    const int lx = (tid & 15) * 4, ly = (tid >> 4) * 2;
    float2 win[4][6];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) win[r][c] = tile[(ly + r) * S2_PITCH + lx + c];
    float lmax = 0.f;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int gy = ty0 + ly + r;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int gx = tx0 + lx + c;
            float o0 = wsm[36], o1 = wsm[37];
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const float2 v = win[r + dy][c + dx];
                    const int t = dy * 3 + dx;
                    o0 = fmaf(v.x, wsm[t], o0);            // W[co=0][ci=0]
                    o0 = fmaf(v.y, wsm[9 + t], o0);        // W[0][1]
                    o1 = fmaf(v.x, wsm[18 + t], o1);       // W[1][0]
                    o1 = fmaf(v.y, wsm[27 + t], o1);       // W[1][1]
                }
            if (gy < H && gx < W) {
                reinterpret_cast<float2*>(a.out)[img + (size_t)gy * W + gx] = make_float2(o0, o1);
                lmax = fmaxf(lmax, fmaxf(fabsf(o0), fabsf(o1)));
            }
        }
    }
    publish_amax(a.amax_out, lmax);
}

}  // namespace hn
