// conv_tc.cuh -- 3x3 convolutions (C_out = 8) of the HybridNet UNet as implicit GEMMs on the 5th-generation
// tensor cores: tcgen05.mma (kind::f16) issued by one thread, operands in shared memory, fp32 accumulators
// in TMEM, epilogue (bias, PReLU, 1x1 outc + wavefield update, concat-free stores) fused after tcgen05.ld.
//
// Reference semantics are those of conv_simt.cuh (helmnet/architectures.py:63-84, 240-252, 439-465).
//
// Precision.  The parity bar is 1e-5 per iteration against the fp32 reference, which plain TF32/BF16/FP16
// operands miss by 30x.  Operands are therefore SPLIT into two fp16 terms,
//     x * 2^sa = hi + lo * 2^-11        (hi = fp16(x'), lo = fp16((x' - hi) * 2^11); 22 significant bits)
// with a power-of-two block scale 2^sa per input tensor (activations: every producer kernel publishes the
// running max |x| of what it writes, see publish_amax; the 6-channel solver input is reduced per CTA tile
// while it sits in registers) and 2^kw per layer (weights, chosen on the host), so fp16's narrow exponent
// range is never a limit.  Products of fp16 pairs are exact in the fp32 accumulator.  With K-stacked A = [hi | lo]
// (one 32-byte row per pixel and tap) and N-stacked B = [[W_hi, W_lo], [0, W_hi]] a single K=16 MMA per
// (pixel block, tap, channel group) yields   g1 = hi*W_hi   and   g2 = hi*W_lo + lo*W_hi   in 16 TMEM columns;
// the epilogue forms  (g1 + 2^-11 g2) * 2^-(sa+kw) + bias.  Only the lo*lo term (2^-22 relative) is dropped.
// This is the fp16 analogue of 3xTF32 at half the shared-memory operand traffic, which is what bounds an
// N = 8..16 UMMA (A-operand reads: 128 rows x 32 B per instruction).
//
// Implicit GEMM without im2col.  The input tile (32 rows x 34 pixels incl. halo) is stored as planes of
// 16-byte entries, plane[pos] = 8 channels of one pixel, pos = y * 34 + x.  In the K-major SWIZZLE_NONE
// canonical layout a core matrix is 8 rows x 16 B, rows 16 B apart, so "row m of the A operand" can simply be
// "pixel pos0 + m": a filter tap (dy,dx) is the SAME plane with the descriptor start address advanced by
// (dy*34 + dx) * 16 bytes, and the second K chunk (the lo term) is reached through the descriptor's leading
// byte offset = distance between the hi and lo planes.  GEMM rows that fall on the two pad columns of each
// tile row compute garbage that is never stored (32/34 useful).
//
// One CTA = one 32 x 30 output tile = 8 blocks of 128 GEMM rows -> 8 x 16 = 128 TMEM columns, up to 4 CTAs
// per SM.  All 8 warps load/convert the tile, thread 0 issues the 72 (or 144) MMAs and commits one mbarrier
// per block, then the 8 warps drain the blocks (warp w: TMEM lane quadrant w%4, blocks w/4, w/4+2, ...) while
// later blocks are still being multiplied.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_simt.cuh"

namespace hn {
namespace tc {

constexpr int TX = 32, TY = 30, PITCH = 34;
constexpr int NBLK = 8;                       // GEMM row blocks of 128 positions (30*34 = 1020 <= 1024)
constexpr int TILE_IN = (TY + 2) * PITCH;     // 1088 input positions
constexpr int TILE_POS = 1096;                // + slack: block 7 / tap (2,2) reads up to 1023 + 70
constexpr int THREADS = 256;
constexpr int PER_THREAD = 5;                 // ceil(1096 / 256)
constexpr int BBLK_BYTES = 512;               // one (group, tap) B operand: 16 x 16 fp16
constexpr int TMEM_COLS = 128;

__host__ __device__ constexpr int groups_of(int src) { return (src == SRC_A8_B2 || src == SRC_A8_B8) ? 2 : 1; }
__host__ __device__ constexpr size_t smem_bytes(int src) {
    return (size_t)groups_of(src) * 2 * TILE_POS * 16 + (size_t)groups_of(src) * 9 * BBLK_BYTES + 256;
}

struct Args {
    const float* inA;
    const float* inB;
    const float* sigma;
    const __half* bmat;   // [groups][9 taps] x 512 B, canonical K-major SWIZZLE_NONE image (host packed)
    const float* bias;    // [8]
    const float* slope;
    float* out;           // NHWC8
    const float* wo;      // EPI_OUTC
    const float* bo;
    float* wf;
    float* dwf_out;
    const unsigned* amax_in0;   // max |x| slots of the sources (bit patterns, see publish_amax); INC computes its own
    const unsigned* amax_in1;
    unsigned* amax_out;         // running max |out| slot or null
    int* error_flag;      // set to 1 if an MMA completion was not observed (watchdog)
    float w_inv_scale;    // 2^-kw
    int H, W;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// start address, leading byte offset (between the two 16-byte K chunks), stride byte offset (between 8-row groups).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, N = 16, M = 128.
constexpr uint32_t kIdesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void split8(const float (&v)[8], float mult, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float a = v[2 * i] * mult, b = v[2 * i + 1] * mult;
        const __half2 hh = __floats2half2_rn(a, b);            // one F2FP per pair
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn((a - back.x) * 2048.f, (b - back.y) * 2048.f);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int SRC, bool PRELU, int EPI>
__global__ void __launch_bounds__(THREADS, (groups_of(SRC) == 1) ? 4 : 2) conv3x3_tc_kernel(Args a) {
    constexpr int G = groups_of(SRC);
    extern __shared__ __align__(128) uint8_t smem_tc[];
    uint4* planes = reinterpret_cast<uint4*>(smem_tc);                       // [G][2 (hi,lo)][TILE_POS]
    uint8_t* bsm = smem_tc + (size_t)G * 2 * TILE_POS * 16;                  // [G][9][512]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(bsm + G * 9 * BBLK_BYTES);  // [NBLK]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + NBLK);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);                    // [8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY, b = blockIdx.z;
    const int H = a.H, W = a.W;
    const size_t img = (size_t)b * H * W;

    // ---- one-time setup: TMEM, barriers, B operand ------------------------------------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
#pragma unroll
        for (int j = 0; j < NBLK; j++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + j)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    {
        const uint4* bg = reinterpret_cast<const uint4*>(a.bmat);
        uint4* bs = reinterpret_cast<uint4*>(bsm);
        for (int i = tid; i < G * 9 * BBLK_BYTES / 16; i += THREADS) bs[i] = __ldg(bg + i);
    }

    pdl_wait();      // TMEM / barriers / weight images above; activations of earlier kernels below (common.cuh: HN_LAUNCH_PDL)
    pdl_trigger();
    // ---- stage the input tile in registers, find the tile's max |x| ---------------------------------------
    // g0 / g1: the 8-channel K groups of the (virtual) concatenated input, zero padded
    float g0[PER_THREAD][8];
    float g1[PER_THREAD][G == 2 ? 8 : 1];
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < PER_THREAD; i++) {
        const int p = tid + i * THREADS;
        const int y = p / PITCH, x = p - y * PITCH;
        const int gy = ty0 - 1 + y, gx = tx0 - 1 + x;
        const bool in = (p < TILE_IN) && gy >= 0 && gy < H && gx >= 0 && gx < W;
#pragma unroll
        for (int c = 0; c < 8; c++) g0[i][c] = 0.f;
#pragma unroll
        for (int c = 0; c < (G == 2 ? 8 : 1); c++) g1[i][c] = 0.f;
        if (in) {
            const size_t pix = img + (size_t)gy * W + gx;
            if constexpr (SRC == SRC_INC) {
                const float2 u = ldg2(a.inA + pix * 2), r = ldg2(a.inB + pix * 2);
                g0[i][0] = u.x; g0[i][1] = u.y;
                g0[i][2] = 1e3f * r.x; g0[i][3] = 1e3f * r.y;                            // hybridnet.py:566
                g0[i][4] = __ldg(a.sigma + gx); g0[i][5] = __ldg(a.sigma + gy);          // sigma_x[i,j]=s[j], sigma_y[i,j]=s[i]
            } else {
                const float4 q0 = ldg4(a.inA + pix * 8), q1 = ldg4(a.inA + pix * 8 + 4);
                g0[i][0] = q0.x; g0[i][1] = q0.y; g0[i][2] = q0.z; g0[i][3] = q0.w;
                g0[i][4] = q1.x; g0[i][5] = q1.y; g0[i][6] = q1.z; g0[i][7] = q1.w;
                if constexpr (SRC == SRC_A8_B8) {
                    const float4 s0 = ldg4(a.inB + pix * 8), s1 = ldg4(a.inB + pix * 8 + 4);
                    g1[i][0] = s0.x; g1[i][1] = s0.y; g1[i][2] = s0.z; g1[i][3] = s0.w;
                    g1[i][4] = s1.x; g1[i][5] = s1.y; g1[i][6] = s1.z; g1[i][7] = s1.w;
                } else if constexpr (SRC == SRC_A8_B2) {
                    const float2 sv = ldg2(a.inB + pix * 2);
                    g1[i][0] = sv.x; g1[i][1] = sv.y;
                }
            }
            if constexpr (SRC == SRC_INC) {
#pragma unroll
                for (int c = 0; c < 6; c++) amax = fmaxf(amax, fabsf(g0[i][c]));
            }
        }
    }
    if constexpr (SRC == SRC_INC) {
        amax = warp_max(amax);
        if (lane == 0) red[warp] = amax;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    if constexpr (SRC == SRC_INC) {
        amax = red[0];
#pragma unroll
        for (int w = 1; w < THREADS / 32; w++) amax = fmaxf(amax, red[w]);
    } else {
        unsigned mb = ld_fresh(a.amax_in0);
        if (G == 2) mb = max(mb, ld_fresh(a.amax_in1));
        amax = __uint_as_float(mb);
    }
    // block scale: x' = x * 2^sa with max|x'| in [2^13, 2^14); exact powers of two built from exponent bits
    int e = (int)((__float_as_uint(amax) >> 23) & 0xffu);          // biased exponent of the max (NaN/Inf -> 255)
    if (e < 40 || e > 250) e = 127;                                // all-zero / denormal / non-finite tile: scale 1
    const float mult = __uint_as_float((uint32_t)(267 - e) << 23);  // 2^(13 - (e - 127))
    const float out_scale = __uint_as_float((uint32_t)(e - 13) << 23) * a.w_inv_scale;   // 2^-(sa) * 2^-kw

    // ---- split to fp16 hi/lo and store the operand planes ---------------------------------------------------
#pragma unroll
    for (int i = 0; i < PER_THREAD; i++) {
        const int p = tid + i * THREADS;
        if (p < TILE_POS) {
            uint4 hi, lo;
            split8(g0[i], mult, hi, lo);
            planes[p] = hi;
            planes[TILE_POS + p] = lo;
            if constexpr (G == 2) {
                split8(g1[i], mult, hi, lo);
                planes[2 * TILE_POS + p] = hi;
                planes[3 * TILE_POS + p] = lo;
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;");   // make the generic-proxy smem writes visible to the tensor core
    __syncthreads();

    // ---- MMA issue: one thread, 9 taps x G groups per 128-row block, one commit per block -------------------
    // (Measured with tools/tc_bench.cu: an M=128,K=16 kind::f16 MMA costs 44.4 cycles for any N <= 32, chained
    // into one accumulator or not, so blocks are issued one after the other and the epilogue of block j overlaps
    // the MMAs of blocks j+1.. .  Interleaving four accumulators was tried and is slower: 310 vs 287 us.)
    if (tid == 0) {
        const uint32_t plane_bytes = TILE_POS * 16;
        const uint32_t a_base = smem_u32(planes), b_base = smem_u32(bsm);
#pragma unroll 1
        for (int j = 0; j < NBLK; j++) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(j * 16);
            uint32_t acc = 0;
#pragma unroll
            for (int g = 0; g < G; g++) {
#pragma unroll
                for (int tap = 0; tap < 9; tap++) {
                    const int shift = (tap / 3) * PITCH + (tap % 3);
                    const uint64_t da = smem_desc(a_base + (uint32_t)(2 * g) * plane_bytes + (uint32_t)(j * 128 + shift) * 16, plane_bytes, 128);
                    const uint64_t db = smem_desc(b_base + (uint32_t)(g * 9 + tap) * BBLK_BYTES, 128, 256);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                        "l"(da), "l"(db), "r"(kIdesc), "r"(acc));
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar + j)));
        }
    }

    // ---- epilogue: TMEM -> registers -> bias / PReLU / outc -> global ---------------------------------------
    float bias[8], wo0[8], wo1[8];
#pragma unroll
    for (int c = 0; c < 8; c++) bias[c] = __ldg(a.bias + c);
    float slope = 0.f, bo0 = 0.f, bo1 = 0.f;
    if (PRELU) slope = __ldg(a.slope);
    if (EPI == EPI_OUTC) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            wo0[c] = __ldg(a.wo + c);
            wo1[c] = __ldg(a.wo + 8 + c);
        }
        bo0 = __ldg(a.bo);
        bo1 = __ldg(a.bo + 1);
    }
    const int quad = warp & 3;
    float lmax = 0.f;
#pragma unroll 1
    for (int j = warp >> 2; j < NBLK; j += 2) {
        uint32_t done = 0;
#pragma unroll 1
        for (int spin = 0; spin < (1 << 24) && !done; spin++) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(mbar + j)), "r"(0u) : "memory");
        }
        if (!done) {
            if (lane == 0) *a.error_flag = 1;
            break;
        }
        asm volatile("tcgen05.fence::after_thread_sync;");
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int q = j * 128 + quad * 32 + lane;
        const int y = q / PITCH, x = q - y * PITCH;
        const int gy = ty0 + y, gx = tx0 + x;
        if (x < TX && y < TY && gy < H && gx < W) {
            float o[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float s = fmaf(__uint_as_float(v[8 + c]), 1.f / 2048.f, __uint_as_float(v[c]));
                o[c] = fmaf(s, out_scale, bias[c]);
                if (PRELU) o[c] = o[c] >= 0.f ? o[c] : slope * o[c];
                if (EPI == EPI_STORE) lmax = fmaxf(lmax, fabsf(o[c]));
            }
            const size_t pix = img + (size_t)gy * W + gx;
            if (EPI == EPI_STORE) {
                float4* dst = reinterpret_cast<float4*>(a.out + pix * 8);
                st_nhwc8(reinterpret_cast<float*>(dst), o);
            } else {
                float o0 = bo0, o1 = bo1;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    o0 = fmaf(o[c], wo0[c], o0);
                    o1 = fmaf(o[c], wo1[c], o1);
                }
                if (a.dwf_out != nullptr) {
                    reinterpret_cast<float2*>(a.dwf_out)[pix] = make_float2(o0, o1);
                } else {
                    float2* wfp = reinterpret_cast<float2*>(a.wf) + pix;
                    const float2 u = *wfp;
                    const float2 nw = make_float2(o0 / 1e3f + u.x, o1 / 1e3f + u.y);   // hybridnet.py:570
                    *wfp = nw;
                    lmax = fmaxf(lmax, fmaxf(fabsf(nw.x), fabsf(nw.y)));
                }
            }
        }
    }
    publish_amax(a.amax_out, lmax);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

}  // namespace tc
}  // namespace hn
