// conv_tcf.cuh -- FUSED DoubleConv (conv3x3 -> PReLU -> conv3x3, helmnet/architectures.py:63-84) on tcgen05:
// the 8-channel intermediate tensor never leaves the SM.
//
// Two chained row-streaming implicit GEMMs (the mapping of conv_tcr.cuh: M = 128 pixels of one image row,
// vertical taps in N, split-fp16 operands, output-stationary TMEM accumulator rings) inside ONE persistent CTA
// that owns FULL-WIDTH image rows (up to W = 128 * NH pixels, NH = 1 or 2; NH = 0 / -1 stand for W = 64 / 32 with M = 64 MMAs, whose
// accumulator row i sits in TMEM lane 32 (i / 16) + i % 16), so the second convolution finds the left/right
// neighbours of every intermediate pixel in its own shared memory -- no halo exchange and no recomputation
// along x; along y a strip of R output rows recomputes 2 intermediate rows.
//
//   TMA warp      : cp.async.bulk of whole fp32 image rows -> staging ring
//   converters    : staging fp32 -> fp16 hi/lo split -> operand ring A1              (one thread per image column)
//   MMA-1 thread  : A1 x W1 -> accumulator ring 1 (TMEM), tcgen05.commit per row pair
//   epilogue 1    : tcgen05.ld ring 1 -> bias + PReLU -> fp16 hi/lo split -> operand ring A2 (rows outside the
//                   image are written as zeros = the zero padding of the second conv); zero + free the accumulators
//   MMA-2 thread  : A2 x W2 -> accumulator ring 2
//   epilogue 2    : tcgen05.ld ring 2 -> bias (+ 1x1 outc + wavefield update | NHWC8 store | float2 store) -> HBM
// All rings advance in steps of two image rows.  Saves 64 B per point and pair of layers of HBM traffic
// (write + re-read of the intermediate tensor) compared with two conv_tcr launches.
//
// Block scale of the intermediate tensor: it is not known before it is produced, so the split uses the bound
//   |mid| <= max(1,|slope|) * (max|in| * max_co sum|W1[co]| + max|b1|)
// (host supplies the weight norms).  fp16 is a floating format: the split keeps 22 significant bits per ELEMENT
// down to 2^-27 of the block maximum, so a loose bound costs nothing.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_simt.cuh"
#include "conv_tc.cuh"
#include "conv_tcr.cuh"

namespace hn {
namespace tcf {

constexpr int TR = 8;              // accumulator ring depth (image rows) per convolution
constexpr int NC = 16;             // TMEM columns per (row, 128-pixel half): [g1(8) | g2(8)]
constexpr int NPB = TR / 2;        // accumulator-pair barriers
constexpr int NDB = 16;            // commit barriers per convolution
constexpr int BROW_BYTES = 1536;   // one (group, dx) B operand: 48 x 16 fp16

__host__ __device__ constexpr int groups_of(int src) { return (src == SRC_A8_B2 || src == SRC_A8_B8) ? 2 : 1; }
__host__ __device__ constexpr int wpx(int nh) { return nh < 0 ? 32 : nh == 0 ? 64 : 128 * nh; }   // image width (NH = -1: 32)
__host__ __device__ constexpr int halves(int nh) { return nh <= 0 ? 1 : nh; }               // MMAs (M = 128 or 64) per row
__host__ __device__ constexpr int psw(int nh) { return (nh < 0 ? 64 : wpx(nh)) + 8; }       // operand positions per row (an M = 64
                                                                                            // MMA reads 64 of them: zeros beyond W)
__host__ __device__ constexpr int conv_warps(int nh) { return wpx(nh) / 32; }
__host__ __device__ constexpr int epi_warps(int nh) { return 4 * halves(nh); }
__host__ __device__ constexpr int threads(int nh) { return (conv_warps(nh) + 3 + 2 * epi_warps(nh)) * 32; }
__host__ __device__ constexpr int tmem_cols(int nh) { return 2 * TR * NC * halves(nh); }
__host__ __device__ constexpr size_t stage_row_bytes(int src, int nh) {
    return (size_t)wpx(nh) * (src == SRC_INC ? 16 : src == SRC_A8 ? 32 : src == SRC_A8_B2 ? 40 : 64);
}
// ring depths in row PAIRS
// (staging: the narrow variants have shared memory to spare and one strip per CTA, so their step time is the HBM latency divided
//  by the number of row pairs in flight: six pairs (four for the 16-channel source) instead of two to four)
#ifndef HN_TCF_NSP_NARROW
#define HN_TCF_NSP_NARROW 6
#endif
// operand-ring depths of the narrow (64- / 32-pixel) variants: one strip per CTA, so a step is bounded by the loops
// converter(j + SRP1) -> MMA-1(j) and epilogue-1(p + SRP2) -> MMA-2(p); 0 = the depths of the wide variants
// (measured, r2: four operand pairs per ring and a six / four deep staging ring instead of 2 / 2 / 8: 64^2 x 32 0.241 -> 0.228 ms
//  per iteration, 64^2 x 256 0.563 -> 0.555, nothing lost elsewhere; the shared memory still allows two CTAs per SM)
#ifndef HN_TCF_SRP1_NARROW
#define HN_TCF_SRP1_NARROW 4
#endif
#ifndef HN_TCF_SRP2_NARROW
#define HN_TCF_SRP2_NARROW 4
#endif
#ifndef HN_TCF_NSP_NARROW_B8
#define HN_TCF_NSP_NARROW_B8 4
#endif
__host__ __device__ constexpr int nsp(int src, int nh) {
    return (nh <= 0 && HN_TCF_NSP_NARROW > 0) ? (src == SRC_A8_B8 ? HN_TCF_NSP_NARROW_B8 : HN_TCF_NSP_NARROW)
                                              : (src == SRC_A8_B8 ? 2 : (src == SRC_INC ? 4 : 3));
}
__host__ __device__ constexpr int srp1(int src, int nh) {
    return (nh <= 0 && HN_TCF_SRP1_NARROW > 0) ? HN_TCF_SRP1_NARROW : (groups_of(src) == 1 ? 4 : (nh == 2 ? 3 : 2));
}
__host__ __device__ constexpr int srp2(int src, int nh) { return (nh <= 0 && HN_TCF_SRP2_NARROW > 0) ? HN_TCF_SRP2_NARROW : 2; }
__host__ __device__ constexpr size_t a1_row_bytes(int src, int nh) { return (size_t)groups_of(src) * 2 * psw(nh) * 16; }
__host__ __device__ constexpr size_t a2_row_bytes(int nh) { return (size_t)2 * psw(nh) * 16; }
__host__ __device__ constexpr size_t smem_bytes(int src, int nh) {
    return (size_t)nsp(src, nh) * 2 * stage_row_bytes(src, nh) + (size_t)srp1(src, nh) * 2 * a1_row_bytes(src, nh) +
           (size_t)srp2(src, nh) * 2 * a2_row_bytes(nh) + (size_t)(groups_of(src) + 1) * 3 * BROW_BYTES + 1024;
}

struct Args {
    const float* inA;
    const float* inB;
    const float* sigma;
    const __half* bmat1;        // first conv: [groups][3 dx] x 1536 B (pack_tcr)
    const __half* bmat2;        // second conv: [3 dx] x 1536 B
    const __half* bfold1;       // SRC_A8_B2: the 2-channel group of the first conv with its 3 dx taps folded into K (1536 B)
    const __half* bfold2;       // EPI_STORE2: the 2 -> 2 second conv, dx folded into K (1536 B)
    // small per-layer constants travel as kernel parameters: the epilogues read them straight from the constant bank
    // (operand form c[0x0][..] of FFMA) instead of re-loading them from shared memory after every tcgen05.wait
    float bias1[8];             // zero padded
    float bias2[8];             // zero padded
    float wo[16];               // 1x1 outc weights [2][8] (EPI_OUTC)
    float bo[2];
    float slope;
    float* out;
    float* wf;
    float* dwf_out;
    const unsigned* amax_in0;
    const unsigned* amax_in1;
    unsigned* amax_out;
    int* error_flag;
    int pdl_trig;               // PDL: let the next kernel's CTAs become resident as this grid's CTAs exit (hn_ctx::pdl)
    float sigma_max;
    float w_inv1, w_inv2;       // 2^-kw of the two layers
    float mid_l1, mid_bmax;     // max_co sum |W1[co]|, max |b1|
    int H;                      // image rows
    int W;                      // image columns, even, <= wpx(NH): the kernel owns full-width rows; GEMM rows / operand positions
                                // beyond W stay zero (they are the right-hand zero padding of pixel W - 1)
    int rows;                   // output rows per strip (even)
    int spi;                    // strips per image
    int total_strips;
    int bal;                    // != 0: balanced strips (strip_of): the batch size; the grid then is one chunk of rows per CTA
    // Narrow images side by side (the 128-pixel variant, images at most 62 pixels wide; not for SRC_INC / EPI_OUTC): `pack` images of a
    // group share the GEMM rows of one M = 128 MMA.  Image k owns the columns [k S, k S + W) of the kernel's row, S = W + 2 (even, so
    // that every staged segment stays 16-byte aligned); the columns between two images are never written by the converters /
    // epilogue 1, so they keep their zeros and are the x zero padding of both neighbours.  The strip walk runs over groups
    // ("image" b of strip_of = group b).
    int pack, pack_s, batch;
};

// Strip i of this CTA (image b, first output row y0, R rows); false when it has none.
//   bal == 0: strips of `rows` rows, strip st = blockIdx.x + i * gridDim.x of the batch (whole rounds of equal strips).
//   bal != 0: the rows of all images are laid end to end, each image preceded by BAL_PAD virtual rows that stand for the cost of
//             starting a strip (4 recomputed halo rows + pipeline fill), and CTA c takes the c-th of gridDim.x equal chunks of that
//             line: every CTA gets the same rows + BAL_PAD * strips, whatever the ratio of images to SMs (one image boundary inside a
//             chunk costs that CTA BAL_PAD rows of work less).  All boundaries are even.
constexpr int BAL_PAD = 8;
// column x of the kernel's row -> (image k of the group, column xl of that image); false for the columns between / beyond the images
__device__ __forceinline__ bool packed_column(const Args& a, int x, int nimg, int& k, int& xl) {
    if (a.pack == 0) {
        k = 0;
        xl = x;
        return x < a.W;
    }
    k = x / a.pack_s;
    xl = x - k * a.pack_s;
    return k < nimg && xl < a.W;
}
__device__ __forceinline__ int group_images(const Args& a, int b) { return a.pack == 0 ? 1 : min(a.pack, a.batch - b * a.pack); }
__device__ __forceinline__ size_t group_first_pixel(const Args& a, int b) { return (size_t)(a.pack == 0 ? b : b * a.pack) * a.H * a.W; }

__device__ __forceinline__ bool strip_of(const Args& a, int i, int& b, int& y0, int& R) {
    if (a.bal == 0) {
        const int st = (int)blockIdx.x + i * (int)gridDim.x;
        if (st >= a.total_strips) return false;
        b = st / a.spi;
        y0 = (st - b * a.spi) * a.rows;
        R = min(a.rows, a.H - y0);
        return true;
    }
    return balanced_strip(a.bal, a.H, BAL_PAD, i, b, y0, R);
}

using tcr::mbar_arrive;
using tcr::mbar_arrive_expect_tx;
using tcr::mbar_wait;
using tcr::tma_load_1d;

using tcr::elect_one;
using tcr::mma_commit;
using tcr::mma_f16;

// MMAs of ONE operand row: local row k of a strip whose conv has R live output rows; gk = global index of the output
// row with the same index (dy = 0).  The accumulator of output row y lives in unit 7 - (y & 7), so rows k, k-1, k-2
// are ascending adjacent units (split at the ring wrap).  acc_base: TMEM column of (half 0, unit 0) of this conv.
// FOLD: the LAST channel group carries only 2 channels (hidden state / 2-channel intermediate); its three horizontal
// taps are folded into K -- operand entry e holds [pixel e-2 | pixel e-1] (hi2 lo2 each) in plane 0 and [pixel e | 0]
// in plane 1, so ONE MMA at the centre position (start entry m + 1) covers dx = 0, 1, 2 against the B image
// k = dx * 4 + part * 2 + c (pack_tcf_fold): 1 MMA instead of 3 for that group.
template <int G, int NH, bool FOLD>
__device__ __forceinline__ void issue_row(uint32_t acc_base, uint32_t a_lo /* (half 0, group 0, dx 0) of this operand row */, uint32_t a_hi,
                                          uint32_t b_lo /* (group 0, dx 0) */, uint32_t b_hi, int k, int gk, int R) {
    constexpr uint32_t kIdescBase = (1u << 4) | (((NH <= 0 ? 64u : 128u) >> 4) << 24);
    constexpr uint32_t kGroup16 = (uint32_t)(2 * psw(NH));     // operand planes of the next channel group, in 16-byte units
    constexpr int NHALF = halves(NH);
    constexpr uint32_t kB16 = BROW_BYTES / 16;
    if (k >= 2 && k < R && (gk & 7) >= 2) {
        const uint32_t d0 = acc_base + (uint32_t)((7 - (gk & 7)) * NC);
#pragma unroll
        for (int h = 0; h < NHALF; h++)
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (FOLD && g == G - 1) {
                    mma_f16(d0 + (uint32_t)(h * TR * NC), a_lo + (uint32_t)(h * 128 + 1) + g * kGroup16, a_hi,
                            b_lo + (uint32_t)(g * 3) * kB16, b_hi, kIdescBase | (6u << 17));
                } else {
#pragma unroll
                    for (int dx = 0; dx < 3; dx++)
                        mma_f16(d0 + (uint32_t)(h * TR * NC), a_lo + (uint32_t)(h * 128 + dx) + g * kGroup16, a_hi,
                                b_lo + (uint32_t)(g * 3 + dx) * kB16, b_hi, kIdescBase | (6u << 17));
                }
            }
    } else {
        const int dlo = max(0, k - (R - 1)), dhi = min(2, k);
        int dy = dlo;
        while (dy <= dhi) {
            const int u = 7 - ((gk - dy) & 7);
            int len = 1;
            while (dy + len <= dhi && u + len <= 7) len++;
            const uint32_t d0 = acc_base + (uint32_t)(u * NC);
            const uint32_t idesc = kIdescBase | ((uint32_t)(2 * len) << 17);   // N = 16 * len
#pragma unroll
            for (int h = 0; h < NHALF; h++)
#pragma unroll
                for (int g = 0; g < G; g++) {
                    if (FOLD && g == G - 1) {
                        mma_f16(d0 + (uint32_t)(h * TR * NC), a_lo + (uint32_t)(h * 128 + 1) + g * kGroup16, a_hi,
                                b_lo + (uint32_t)(g * 3) * kB16 + (uint32_t)(dy * 32), b_hi, idesc);
                    } else {
#pragma unroll
                        for (int dx = 0; dx < 3; dx++)
                            mma_f16(d0 + (uint32_t)(h * TR * NC), a_lo + (uint32_t)(h * 128 + dx) + g * kGroup16, a_hi,
                                    b_lo + (uint32_t)(g * 3 + dx) * kB16 + (uint32_t)(dy * 32), b_hi, idesc);
                    }
                }
            dy += len;
        }
    }
}

// Two channels -> [hi c0, hi c1, lo c0, lo c1] (8 bytes): the split of tc::split8 for one channel pair.
__device__ __forceinline__ uint2 split2(float v0, float v1, float mult) {
    const float a = v0 * mult, b = v1 * mult;
    const __half2 hh = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn((a - back.x) * 2048.f, (b - back.y) * 2048.f);
    return make_uint2(*reinterpret_cast<const uint32_t*>(&hh), *reinterpret_cast<const uint32_t*>(&ll));
}
// Store pixel x of a folded 2-channel group: `planes` = plane 0 of the group (plane 1 is PSW entries further).
template <int PSW>
__device__ __forceinline__ void store_fold(uint4* planes, int x, uint2 v) {
    uint2* f = reinterpret_cast<uint2*>(planes);
    f[2 * (x + 2)] = v;             // entry x + 2, dx = 0 field: GEMM row x + 1 sees its left neighbour
    f[2 * (x + 1) + 1] = v;         // entry x + 1, dx = 1 field: GEMM row x, centre tap
    f[2 * (PSW + x)] = v;           // plane 1, entry x, dx = 2 field: GEMM row x - 1 sees its right neighbour
}

// Read the accumulators of output rows (gr, gr + 1) (gr even, global row index) of this thread's TMEM lane, then hand
// them back zeroed.  v[t] = row gr + t.
__device__ __forceinline__ void drain_pair(uint32_t tb /* lane/half-adjusted address of unit(gr + 1) */, uint32_t (&v)[2][16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[1][0]), "=r"(v[1][1]), "=r"(v[1][2]), "=r"(v[1][3]), "=r"(v[1][4]), "=r"(v[1][5]), "=r"(v[1][6]), "=r"(v[1][7]),
                   "=r"(v[1][8]), "=r"(v[1][9]), "=r"(v[1][10]), "=r"(v[1][11]), "=r"(v[1][12]), "=r"(v[1][13]), "=r"(v[1][14]), "=r"(v[1][15])
                 : "r"(tb));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0][0]), "=r"(v[0][1]), "=r"(v[0][2]), "=r"(v[0][3]), "=r"(v[0][4]), "=r"(v[0][5]), "=r"(v[0][6]), "=r"(v[0][7]),
                   "=r"(v[0][8]), "=r"(v[0][9]), "=r"(v[0][10]), "=r"(v[0][11]), "=r"(v[0][12]), "=r"(v[0][13]), "=r"(v[0][14]), "=r"(v[0][15])
                 : "r"(tb + NC));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tb), "r"(z));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
}

__device__ __forceinline__ int exp_of(float v) {
    int e = (int)((__float_as_uint(v) >> 23) & 0xffu);
    if (e < 40 || e > 250) e = 127;
    return e;
}

template <int SRC, int NH, int EPI>
__global__ void __launch_bounds__(threads(NH), NH == 2 ? 1 : 2) dconv_tcf_kernel(Args a) {
    static_assert(NH >= -1 && NH <= 2, "NH = -1 (32 px), 0 (64 px), 1 (128 px) or 2 (256 px)");
    constexpr int G = groups_of(SRC);
    constexpr bool FOLD1 = SRC == SRC_A8_B2;     // hidden-state group of the first conv: dx folded into K
    constexpr bool FOLD2 = EPI == EPI_STORE2;    // 2 -> 2 second conv: dx folded into K
    constexpr int W = wpx(NH);
    constexpr int PSW = psw(NH);
    constexpr int NSP = nsp(SRC, NH), SRP1 = srp1(SRC, NH), SRP2 = srp2(SRC, NH);
    constexpr int CONV_WARPS = conv_warps(NH), EPI_WARPS = epi_warps(NH);
    constexpr int MMA1_WARP = CONV_WARPS, MMA2_WARP = CONV_WARPS + 1, TMA_WARP = CONV_WARPS + 2;
    constexpr int EPI1_WARP0 = CONV_WARPS + 3, EPI2_WARP0 = EPI1_WARP0 + EPI_WARPS;
    constexpr int THREADS = threads(NH);
    constexpr int NCT = CONV_WARPS * 32;         // converter threads == W: one per image column
    constexpr int NET = EPI_WARPS * 32;          // threads per epilogue == W
    constexpr size_t SROW = stage_row_bytes(SRC, NH), A1ROW = a1_row_bytes(SRC, NH), A2ROW = a2_row_bytes(NH);
    constexpr uint32_t ACC1 = 0, ACC2 = halves(NH) * TR * NC;   // TMEM column offsets of the two accumulator rings

    extern __shared__ __align__(128) uint8_t smem_tcf[];
    uint8_t* stage = smem_tcf;                                   // [NSP][2][SROW]
    uint8_t* a1 = stage + (size_t)NSP * 2 * SROW;                // [SRP1][2][G][2][PSW] x 16 B
    uint8_t* a2 = a1 + (size_t)SRP1 * 2 * A1ROW;                 // [SRP2][2][2][PSW] x 16 B
    uint8_t* b1 = a2 + (size_t)SRP2 * 2 * A2ROW;                 // [G][3][1536]
    uint8_t* b2 = b1 + (size_t)G * 3 * BROW_BYTES;               // [3][1536]
    uint64_t* bars = reinterpret_cast<uint64_t*>(b2 + 3 * BROW_BYTES);
    uint64_t* stage_full = bars;                 // [NSP]  1 arrival + TMA bytes
    uint64_t* stage_empty = stage_full + NSP;    // [NSP]  NCT converter arrivals
    uint64_t* a1_full = stage_empty + NSP;       // [SRP1] NCT
    uint64_t* c1_done = a1_full + SRP1;          // [NDB]  tcgen05.commit of conv-1 input pair
    uint64_t* acc1_empty = c1_done + NDB;        // [NPB]  NET
    uint64_t* a2_full = acc1_empty + NPB;        // [SRP2] NET
    uint64_t* c2_done = a2_full + SRP2;          // [NDB]  tcgen05.commit of conv-2 input (mid) pair
    uint64_t* acc2_empty = c2_done + NDB;        // [NPB]  NET
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + NPB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, Wa = a.W;

    // ---- setup ------------------------------------------------------------------------------------------
    if (warp == MMA1_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "n"(tmem_cols(NH)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        auto init = [](uint64_t* b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(b)), "r"(cnt)); };
        for (int i = 0; i < NSP; i++) { init(stage_full + i, 1); init(stage_empty + i, NCT); }
        for (int i = 0; i < SRP1; i++) init(a1_full + i, NCT);
        for (int i = 0; i < SRP2; i++) init(a2_full + i, NET);
        for (int i = 0; i < NDB; i++) { init(c1_done + i, 1); init(c2_done + i, 1); }
        for (int i = 0; i < NPB; i++) { init(acc1_empty + i, NET); init(acc2_empty + i, NET); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    {
        // operand rings start zeroed: positions 0 and W+1 of every row (the x zero padding) are never written again
        uint4* z = reinterpret_cast<uint4*>(a1);
        const int nz = (int)(((size_t)SRP1 * 2 * A1ROW + (size_t)SRP2 * 2 * A2ROW) / 16);
        for (int i = tid; i < nz; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        const uint4* bg1 = reinterpret_cast<const uint4*>(a.bmat1);
        const uint4* bg2 = reinterpret_cast<const uint4*>(a.bmat2);
        uint4* bs1 = reinterpret_cast<uint4*>(b1);
        uint4* bs2 = reinterpret_cast<uint4*>(b2);
        constexpr int NB1 = FOLD1 ? 3 : G * 3;       // images taken from bmat1 (FOLD1: group 0 only, then the folded image)
        for (int i = tid; i < NB1 * BROW_BYTES / 16; i += THREADS) bs1[i] = __ldg(bg1 + i);
        if constexpr (FOLD1) {
            const uint4* bf = reinterpret_cast<const uint4*>(a.bfold1);
            for (int i = tid; i < BROW_BYTES / 16; i += THREADS) bs1[3 * BROW_BYTES / 16 + i] = __ldg(bf + i);
        }
        if constexpr (FOLD2) {
            const uint4* bf = reinterpret_cast<const uint4*>(a.bfold2);
            for (int i = tid; i < BROW_BYTES / 16; i += THREADS) bs2[i] = __ldg(bf + i);
        } else {
            for (int i = tid; i < 3 * BROW_BYTES / 16; i += THREADS) bs2[i] = __ldg(bg2 + i);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= EPI1_WARP0) {   // every epilogue warp zeroes its lane quadrant / half of its accumulator ring
        const bool second = warp >= EPI2_WARP0;
        const int we = warp - (second ? EPI2_WARP0 : EPI1_WARP0);
        const uint32_t base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (second ? ACC2 : ACC1) + (uint32_t)((we >> 2) * TR * NC);
        const uint32_t z = 0u;
#pragma unroll 1
        for (int u = 0; u < TR; u++)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(base + (uint32_t)(u * NC)), "r"(z));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    // everything above touched only this CTA's shared memory / TMEM and the constant weight images: from here on the kernel
    // reads what earlier kernels of the iteration wrote (common.cuh: HN_LAUNCH_PDL)
    pdl_wait();
    if (a.pdl_trig) pdl_trigger();
    // ---- block scales ---------------------------------------------------------------------------------------
    float amax_in;
    if constexpr (SRC == SRC_INC) {
        amax_in = fmaxf(fmaxf(__uint_as_float(ld_fresh(a.amax_in0)), 1e3f * __uint_as_float(ld_fresh(a.amax_in1))), a.sigma_max);
    } else {
        unsigned mb = ld_fresh(a.amax_in0);
        if (G == 2) mb = max(mb, ld_fresh(a.amax_in1));
        amax_in = __uint_as_float(mb);
    }
    const int e_in = exp_of(amax_in);
    const float mult1 = __uint_as_float((uint32_t)(267 - e_in) << 23);                      // x' = x * 2^(140 - e)
    const float scale1 = __uint_as_float((uint32_t)(e_in - 13) << 23) * a.w_inv1;           // 2^(e - 140) * 2^-kw
    const float in_hi = __uint_as_float((uint32_t)(e_in + 1) << 23);                        // >= max |in|
    const float slope = a.slope;
    const float bound = fmaxf(1.f, fabsf(slope)) * fmaf(in_hi, a.mid_l1, a.mid_bmax);
    const int e_mid = exp_of(bound);
    const float mult2 = __uint_as_float((uint32_t)(267 - e_mid) << 23);
    const float scale2 = __uint_as_float((uint32_t)(e_mid - 13) << 23) * a.w_inv2;

    bool ok = true;
    if (warp == TMA_WARP) {
        // =============================== TMA issuer ===============================================================
        if (lane == 0) {
            int gj = 0;
#pragma unroll 1
            for (int si = 0, b, y0, R; ok && strip_of(a, si, b, y0, R); si++) {
                const int NP1 = (R + 4) / 2;
                const size_t img = group_first_pixel(a, b);
                const int nimg = group_images(a, b);
                const size_t img_px = (size_t)H * Wa;
                const uint32_t row_tx = (uint32_t)Wa * (uint32_t)(SROW / W) * (uint32_t)nimg;   // bytes one row of the group brings in
#pragma unroll 1
                for (int j = 0; j < NP1; j++, gj++) {
                    const int sidx = gj % NSP;
                    if (!mbar_wait(stage_empty + sidx, ((uint32_t)(gj / NSP) & 1u) ^ 1u)) { ok = false; break; }
                    const int gy0 = y0 - 2 + 2 * j;
                    const bool v0 = gy0 >= 0 && gy0 < H, v1 = gy0 + 1 >= 0 && gy0 + 1 < H;
                    if (!v0 && !v1) {
                        mbar_arrive(stage_full + sidx);
                        continue;
                    }
                    mbar_arrive_expect_tx(stage_full + sidx, ((uint32_t)v0 + (uint32_t)v1) * row_tx);
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        if (!(t == 0 ? v0 : v1)) continue;
                        uint8_t* dst0 = stage + (size_t)(sidx * 2 + t) * SROW;
                        for (int k = 0; k < nimg; k++) {      // (packed: image k of the group lands at column k S of the staged row)
                            const size_t rowpix = img + k * img_px + (size_t)(gy0 + t) * Wa;
                            const int xk = k * a.pack_s;
                            uint8_t* dst = dst0;
                            if constexpr (SRC == SRC_INC) {
                                tma_load_1d(dst, a.inA + rowpix * 2, Wa * 8, stage_full + sidx);
                                tma_load_1d(dst + W * 8, a.inB + rowpix * 2, Wa * 8, stage_full + sidx);
                            } else if constexpr (SRC == SRC_A8) {
                                tma_load_1d(dst + xk * 32, a.inA + rowpix * 8, Wa * 32, stage_full + sidx);
                            } else if constexpr (SRC == SRC_A8_B8) {
                                tma_load_1d(dst + xk * 32, a.inA + rowpix * 8, Wa * 32, stage_full + sidx);
                                tma_load_1d(dst + W * 32 + xk * 32, a.inB + rowpix * 8, Wa * 32, stage_full + sidx);
                            } else {
                                tma_load_1d(dst + xk * 32, a.inA + rowpix * 8, Wa * 32, stage_full + sidx);
                                tma_load_1d(dst + W * 32 + xk * 8, a.inB + rowpix * 2, Wa * 8, stage_full + sidx);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp < CONV_WARPS) {
        // =============================== converters: thread <-> image column x = tid ================================
        const int x = tid;
        bool live = x < Wa;           // columns beyond the image (packed: between / beyond the images of the group) keep zero operands
        const float sig_x = (SRC == SRC_INC && live) ? __ldg(a.sigma + x) : 0.f;
        const int sw16 = ((x >> 2) & 1) * 16;
        int gj = 0;
#pragma unroll 1
        for (int si = 0, b, y0, R; ok && strip_of(a, si, b, y0, R); si++) {
            const int NP1 = (R + 4) / 2;
            if (a.pack > 0) {
                int kk, xl;
                live = packed_column(a, x, group_images(a, b), kk, xl);
            }
#pragma unroll 1
            for (int j = 0; j < NP1; j++, gj++) {
                const int sidx = gj % NSP, s = gj % SRP1;
                if (!mbar_wait(stage_full + sidx, (uint32_t)(gj / NSP) & 1u)) { ok = false; break; }
                // operand pair slot s was last used by pair gj - SRP1: free once that pair's MMAs completed
                if (gj >= SRP1 && !mbar_wait(c1_done + ((gj - SRP1) & (NDB - 1)), (uint32_t)((gj - SRP1) / NDB) & 1u)) { ok = false; break; }
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int gy = y0 - 2 + 2 * j + t;
                    float g0[8];
                    float g1[G == 2 ? 8 : 1];
#pragma unroll
                    for (int c = 0; c < 8; c++) g0[c] = 0.f;
#pragma unroll
                    for (int c = 0; c < (G == 2 ? 8 : 1); c++) g1[c] = 0.f;
                    if (gy >= 0 && gy < H && live) {
                        const uint8_t* src = stage + (size_t)(sidx * 2 + t) * SROW;
                        if constexpr (SRC == SRC_INC) {
                            const float2 u2 = *reinterpret_cast<const float2*>(src + x * 8);
                            const float2 r2 = *reinterpret_cast<const float2*>(src + W * 8 + x * 8);
                            g0[0] = u2.x; g0[1] = u2.y;
                            g0[2] = 1e3f * r2.x; g0[3] = 1e3f * r2.y;                             // hybridnet.py:566
                            g0[4] = sig_x; g0[5] = __ldg(a.sigma + gy);
                        } else {
                            // The two 16-byte halves of a pixel are read in an order that alternates every 4 lanes, so the
                            // 8 lanes of a quarter warp touch 8 distinct 16-byte bank groups (pixel stride is 32 B).
                            const float4 qa = *reinterpret_cast<const float4*>(src + x * 32 + sw16), qb = *reinterpret_cast<const float4*>(src + x * 32 + (sw16 ^ 16));
                            const float4 q0 = sw16 ? qb : qa, q1 = sw16 ? qa : qb;
                            g0[0] = q0.x; g0[1] = q0.y; g0[2] = q0.z; g0[3] = q0.w;
                            g0[4] = q1.x; g0[5] = q1.y; g0[6] = q1.z; g0[7] = q1.w;
                            if constexpr (SRC == SRC_A8_B8) {
                                const float4 sa = *reinterpret_cast<const float4*>(src + W * 32 + x * 32 + sw16);
                                const float4 sb = *reinterpret_cast<const float4*>(src + W * 32 + x * 32 + (sw16 ^ 16));
                                const float4 s0 = sw16 ? sb : sa, s1 = sw16 ? sa : sb;
                                g1[0] = s0.x; g1[1] = s0.y; g1[2] = s0.z; g1[3] = s0.w;
                                g1[4] = s1.x; g1[5] = s1.y; g1[6] = s1.z; g1[7] = s1.w;
                            } else if constexpr (SRC == SRC_A8_B2) {
                                const float2 sv = *reinterpret_cast<const float2*>(src + W * 32 + x * 8);
                                g1[0] = sv.x; g1[1] = sv.y;
                            }
                        }
                    }
                    uint4* slot = reinterpret_cast<uint4*>(a1 + (size_t)(s * 2 + t) * A1ROW);
                    uint4 hi, lo;
                    tc::split8(g0, mult1, hi, lo);
                    slot[x + 1] = hi;
                    slot[PSW + x + 1] = lo;
                    if constexpr (FOLD1) {
                        store_fold<PSW>(slot + 2 * PSW, x, split2(g1[0], g1[1], mult1));
                    } else if constexpr (G == 2) {
                        tc::split8(g1, mult1, hi, lo);
                        slot[2 * PSW + x + 1] = hi;
                        slot[3 * PSW + x + 1] = lo;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;");
                mbar_arrive(a1_full + s);
                mbar_arrive(stage_empty + sidx);
            }
        }
    } else if (warp == MMA1_WARP) {
        // =============================== MMA issuer, first convolution ==============================================
        // The whole warp runs this loop with warp-uniform control flow and values (so descriptors live in uniform
        // registers); one elected lane issues the MMAs and the commit.
        {
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0) + ACC1;
            const uint64_t da = tc::smem_desc(tc::smem_u32(a1), PSW * 16, 128);
            const uint64_t db = tc::smem_desc(tc::smem_u32(b1), 128, 256);
            const uint32_t a_lo = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
            constexpr uint32_t kRow16 = (uint32_t)(A1ROW >> 4);
            int gj = 0, gmp = 0;   // global input pair / mid pair counters at the start of the strip
#pragma unroll 1
            for (int si = 0, b, y0, R; ok && strip_of(a, si, b, y0, R); si++) {
                const int NP1 = (R + 4) / 2, RM = R + 2;
#pragma unroll 1
                for (int j = 0; j < NP1; j++) {
                    const int g_ = gj + j, s = g_ % SRP1;
                    bool w = mbar_wait(a1_full + s, (uint32_t)(g_ / SRP1) & 1u);
                    const int gp = gmp + j;   // mid pair first touched by this input pair
                    if (2 * j < RM) w = mbar_wait(acc1_empty + (gp & (NPB - 1)), ((uint32_t)(gp / NPB) & 1u) ^ 1u) && w;
                    if (!__all_sync(0xffffffffu, w)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < 2; t++) {
                            const int k = 2 * j + t;
                            issue_row<G, NH, FOLD1>(tb, a_lo + (uint32_t)(s * 2 + t) * kRow16, a_hi, b_lo, b_hi, k, 2 * gmp + k, RM);
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(c1_done + (g_ & (NDB - 1)))));
                    }
                    __syncwarp();
                }
                gj += NP1;
                gmp += RM / 2;
            }
        }
    } else if (warp == MMA2_WARP) {
        // =============================== MMA issuer, second convolution =============================================
        {
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0) + ACC2;
            const uint64_t da = tc::smem_desc(tc::smem_u32(a2), PSW * 16, 128);
            const uint64_t db = tc::smem_desc(tc::smem_u32(b2), 128, 256);
            const uint32_t a_lo = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
            constexpr uint32_t kRow16 = (uint32_t)(A2ROW >> 4);
            int gmp = 0, gop = 0;
#pragma unroll 1
            for (int si = 0, b, y0, R; ok && strip_of(a, si, b, y0, R); si++) {
                const int NM = (R + 2) / 2;
#pragma unroll 1
                for (int p = 0; p < NM; p++) {
                    const int g_ = gmp + p, s = g_ % SRP2;
                    bool w = mbar_wait(a2_full + s, (uint32_t)(g_ / SRP2) & 1u);
                    const int go_ = gop + p;
                    if (2 * p < R) w = mbar_wait(acc2_empty + (go_ & (NPB - 1)), ((uint32_t)(go_ / NPB) & 1u) ^ 1u) && w;
                    if (!__all_sync(0xffffffffu, w)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < 2; t++) {
                            const int k = 2 * p + t;
                            issue_row<1, NH, FOLD2>(tb, a_lo + (uint32_t)(s * 2 + t) * kRow16, a_hi, b_lo, b_hi, k, 2 * gop + k, R);
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(c2_done + (g_ & (NDB - 1)))));
                    }
                    __syncwarp();
                }
                gmp += NM;
                gop += R / 2;
            }
        }
    } else if (warp < EPI2_WARP0) {
        // =============================== epilogue 1: accumulators -> PReLU -> operand ring A2 ========================
        const int half = (warp - EPI1_WARP0) >> 2, quad = warp & 3;
        const int x = NH <= 0 ? quad * 16 + (lane & 15) : half * 128 + quad * 32 + lane;
        bool act = (NH > 0 || (lane < 16 && quad * 16 < W)) && x < Wa;   // M = 64: 16 accumulator rows per TMEM lane quadrant
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + ACC1 + (uint32_t)(half * TR * NC);
        int gj = 0, gmp = 0;
#pragma unroll 1
        for (int si = 0, b, y0, R; ok && strip_of(a, si, b, y0, R); si++) {
            const int NP1 = (R + 4) / 2, NM = (R + 2) / 2;
            if (a.pack > 0) {
                int kk, xl;
                act = packed_column(a, x, group_images(a, b), kk, xl);
            }
#pragma unroll 1
            for (int p = 0; p < NM; p++) {
                const int gjd = gj + p + 1;                 // input pair that completes mid rows 2p, 2p+1
                if (!mbar_wait(c1_done + (gjd & (NDB - 1)), (uint32_t)(gjd / NDB) & 1u)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;");
                uint32_t v[2][16];
                const int gp = gmp + p;
                drain_pair(tlane + (uint32_t)((7 - ((2 * gp + 1) & 7)) * NC), v);
                mbar_arrive(acc1_empty + (gp & (NPB - 1)));
                // A2 pair slot was last read by the MMAs of mid pair gp - SRP2
                const int s2 = gp % SRP2;
                if (gp >= SRP2 && !mbar_wait(c2_done + ((gp - SRP2) & (NDB - 1)), (uint32_t)((gp - SRP2) / NDB) & 1u)) { ok = false; break; }
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int gy = y0 - 1 + 2 * p + t;      // image row of this mid row
                    float m[8];
                    const bool inside = gy >= 0 && gy < H;
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        float o = fmaf(fmaf(__uint_as_float(v[t][8 + c]), 1.f / 2048.f, __uint_as_float(v[t][c])), scale1, a.bias1[c]);
                        o = o >= 0.f ? o : slope * o;
                        m[c] = inside ? o : 0.f;
                    }
                    uint4* slot = reinterpret_cast<uint4*>(a2 + (size_t)(s2 * 2 + t) * A2ROW);
                    if constexpr (FOLD2) {
                        if (act) store_fold<PSW>(slot, x, split2(m[0], m[1], mult2));
                    } else {
                        uint4 hi, lo;
                        tc::split8(m, mult2, hi, lo);
                        if (act) {
                            slot[x + 1] = hi;
                            slot[PSW + x + 1] = lo;
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;");
                mbar_arrive(a2_full + s2);
            }
            gj += NP1;
            gmp += NM;
        }
    } else {
        // =============================== epilogue 2: accumulators -> bias / outc / update -> HBM =====================
        const int half = (warp - EPI2_WARP0) >> 2, quad = warp & 3;
        const int x = NH <= 0 ? quad * 16 + (lane & 15) : half * 128 + quad * 32 + lane;
        bool act = (NH > 0 || (lane < 16 && quad * 16 < W)) && x < Wa;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + ACC2 + (uint32_t)(half * TR * NC);
        float lmax = 0.f;
        int gmp = 0, gop = 0;
        const int xcol = x;             // column of the kernel's row; `x` becomes the column inside this thread's image when packed
#pragma unroll 1
        for (int si = 0, b, y0, R; ok && strip_of(a, si, b, y0, R); si++) {
            const int NM = (R + 2) / 2;
            size_t img = group_first_pixel(a, b);
            int x = xcol;
            if (a.pack > 0) {
                int kk;
                act = packed_column(a, xcol, group_images(a, b), kk, x);
                img += (size_t)kk * H * Wa;
            }
#pragma unroll 1
            for (int q = 0; q < R / 2; q++) {
                const int ya = 2 * q;
                float2 wfa = make_float2(0.f, 0.f), wfb = wfa;
                if (EPI == EPI_OUTC && a.dwf_out == nullptr && act) {   // issue the wavefield loads before waiting on the MMAs
                    wfa = reinterpret_cast<const float2*>(a.wf)[img + (size_t)(y0 + ya) * Wa + x];
                    wfb = reinterpret_cast<const float2*>(a.wf)[img + (size_t)(y0 + ya + 1) * Wa + x];
                }
                const int gmd = gmp + q + 1;               // mid pair that completes output rows 2q, 2q+1
                if (!mbar_wait(c2_done + (gmd & (NDB - 1)), (uint32_t)(gmd / NDB) & 1u)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;");
                uint32_t v[2][16];
                const int go_ = gop + q;
                drain_pair(tlane + (uint32_t)((7 - ((2 * go_ + 1) & 7)) * NC), v);
                mbar_arrive(acc2_empty + (go_ & (NPB - 1)));
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    if (!act) break;        // M = 64: idle lanes (structured: the warp reconverges after this loop)
                    float o[8];
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        o[c] = fmaf(fmaf(__uint_as_float(v[t][8 + c]), 1.f / 2048.f, __uint_as_float(v[t][c])), scale2, a.bias2[c]);
                    const size_t pix = img + (size_t)(y0 + ya + t) * Wa + x;
                    if (EPI == EPI_STORE) {
#pragma unroll
                        for (int c = 0; c < 8; c++) lmax = fmaxf(lmax, fabsf(o[c]));
                        float4* dst = reinterpret_cast<float4*>(a.out + pix * 8);
                        st_nhwc8(reinterpret_cast<float*>(dst), o);
                    } else if (EPI == EPI_STORE2) {
                        lmax = fmaxf(lmax, fmaxf(fabsf(o[0]), fabsf(o[1])));
                        reinterpret_cast<float2*>(a.out)[pix] = make_float2(o[0], o[1]);
                    } else {
                        float o0 = a.bo[0], o1 = a.bo[1];
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            o0 = fmaf(o[c], a.wo[c], o0);
                            o1 = fmaf(o[c], a.wo[8 + c], o1);
                        }
                        if (a.dwf_out != nullptr) {
                            reinterpret_cast<float2*>(a.dwf_out)[pix] = make_float2(o0, o1);
                        } else {
                            const float2 u = t == 0 ? wfa : wfb;
                            const float2 nw = make_float2(__fdividef(o0, 1e3f) + u.x, __fdividef(o1, 1e3f) + u.y);   // hybridnet.py:570
                            reinterpret_cast<float2*>(a.wf)[pix] = nw;
                            lmax = fmaxf(lmax, fmaxf(fabsf(nw.x), fabsf(nw.y)));
                        }
                    }
                }
            }
            gmp += NM;
            gop += R / 2;
        }
        publish_amax(a.amax_out, lmax);
    }
    if (!ok) *a.error_flag = 1;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == MMA1_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(tmem_cols(NH)));
}

}  // namespace tcf
}  // namespace hn
