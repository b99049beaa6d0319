// conv_tcr.cuh -- row-streaming tcgen05 implicit-GEMM 3x3 convolution (C_out = 8), the main conv engine.
//
// Same arithmetic as conv_tc.cuh (split-fp16 operands, fp32 TMEM accumulators, fused epilogues; reference
// semantics helmnet/architectures.py:63-84, 240-252, 439-465) with a different GEMM mapping.  tools/tc_bench.cu
// shows that an M=128, K=16 MMA costs a fixed ~44 cycles for any N <= 48 (A-operand fetch), so the way to go
// faster is to do more useful N per A fetch:
//
//   GEMM block = ONE image row segment of 128 pixels (M = 128, every row useful).
//   The three dy taps go into N:  for input row i, one MMA per dx computes
//       D[x, (dy, co)] = sum_ci in[i][x + dx - 1][ci] * W[dy][dx][ci][co]        (N = 3 dy x (8 + 8) = 48 columns)
//   and column group dy belongs to OUTPUT row y = i + 1 - dy.  TMEM holds one 16-column accumulator per output
//   row, laid out so that the accumulators of rows y, y-1, y-2 are adjacent: the 48-column result of input row
//   i lands directly on (and accumulates into) the three output rows it feeds -- same TMEM lane, no cross-lane
//   traffic, no partial sums to add in the epilogue.  Accumulators are zeroed by the epilogue (tcgen05.st) when
//   it frees them, so every MMA runs with accumulate = 1.
//   => 3 MMAs per 128 output pixels and channel group instead of 9, 16 accumulator columns per output row,
//      a 16-deep accumulator ring in 256 TMEM columns.
//
// A CTA streams a strip of 128 columns x up to 32 rows through four warp-specialised roles connected by
// mbarrier rings:
//   TMA warp   (1 thread)          : cp.async.bulk of whole fp32 row segments global -> staging ring (deep prefetch,
//                                    no registers: this is what keeps enough bytes in flight to load HBM)
//   converters (2 teams x 5 warps) : staging fp32 -> fp16 hi/lo split -> operand row ring (canonical UMMA layout)
//   MMA warp   (1 thread issues)   : operand ring -> tcgen05.mma -> 16-deep ring of output-row accumulators, tcgen05.commit
//   epilogue   (4 warps)           : tcgen05.ld of one accumulator -> bias/PReLU/outc/update -> global; zero + free it
// Every ring advances in steps of TWO image rows (one barrier round trip per row pair): the roles are latency bound
// per step (mbarrier wake-ups, tcgen05.ld/st round trips, one serial MMA-issuing thread), not throughput bound,
// so halving the number of steps is what moved the kernel (profiles/r1_ncu_tcr_*.txt).
// The kernel is PERSISTENT: 2 CTAs per SM (2 x 256 TMEM columns) walk over all strips of all samples with running
// ring counters, so TMEM allocation, barrier setup, accumulator zeroing and pipeline fill/drain are paid once per
// CTA instead of once per strip (32-row strips cost 15 % in setup/drain when launched one CTA per strip).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_simt.cuh"
#include "conv_tc.cuh"

namespace hn {
namespace tcr {

constexpr int CW = 128;            // strip width = GEMM M
constexpr int PS = 136;            // positions per shared-memory row (130 used: x0-1 .. x0+128)
constexpr int SS = 8;              // staging ring depth (fp32 rows landed by TMA); 6 for the 16-channel source
constexpr int TR = 16;             // TMEM accumulator ring depth: one 16-column unit per output row
constexpr int NC = 16;             // TMEM columns per output row: [g1(8) | g2(8)]
constexpr int TMEM_COLS = 256;
constexpr int TEAM = 160;          // converter threads per team (136 active: one per position)
constexpr int PROD_WARPS = 10, EPI_WARPS = 4;
constexpr int MMA_WARP = PROD_WARPS;
constexpr int TMA_WARP = PROD_WARPS + 1 + EPI_WARPS;
constexpr int THREADS = (PROD_WARPS + 1 + EPI_WARPS + 1) * 32;   // 512
constexpr int ROWS = 32;           // default output rows per strip (Args::rows is picked per launch by the host, even)
constexpr int ST8 = 4224;          // staging bytes of one 8-channel row segment (130 px x 32 B, padded)
constexpr int ST2 = 1152;          // staging bytes of one 2-channel row segment (132 px x 8 B starting at x0-2, padded)
constexpr int BROW_BYTES = 1536;   // one (group, dx) B operand: 48 x 16 fp16
constexpr int SPIN_LIMIT = 1 << 22;

__host__ __device__ constexpr int groups_of(int src) { return (src == SRC_A8_B2 || src == SRC_A8_B8) ? 2 : 1; }
__host__ __device__ constexpr size_t slot_bytes(int src) { return (size_t)groups_of(src) * 2 * PS * 16; }
__host__ __device__ constexpr int op_ring(int src) { return groups_of(src) == 1 ? 8 : 4; }            // operand ring depth
__host__ __device__ constexpr int stage_ring(int src) { return src == SRC_A8_B8 ? 6 : SS; }
__host__ __device__ constexpr size_t stage_bytes(int src) {
    return src == SRC_INC ? 2 * ST2 : src == SRC_A8 ? ST8 : src == SRC_A8_B2 ? ST8 + ST2 : 2 * ST8;
}
__host__ __device__ constexpr size_t smem_bytes(int src) {
    return (size_t)stage_ring(src) * stage_bytes(src) + (size_t)op_ring(src) * slot_bytes(src) +
           (size_t)groups_of(src) * 3 * BROW_BYTES + 768;
}

struct Args {
    const float* inA;
    const float* inB;
    const float* sigma;
    const __half* bmat;         // [groups][3 dx] x 1536 B canonical K-major SWIZZLE_NONE images (host packed)
    const float* bias;
    const float* slope;
    float* out;
    const float* wo;
    const float* bo;
    float* wf;
    float* dwf_out;
    const unsigned* amax_in0;   // INC: max|wf| slot; others: max|x| of source A
    const unsigned* amax_in1;   // INC: max|res| slot; others: source B
    unsigned* amax_out;         // EPI_STORE: max|out|; EPI_OUTC: max|updated wavefield|
    int* error_flag;
    int pdl_trig;               // PDL: let the next kernel's CTAs become resident as this grid's CTAs exit (hn_ctx::pdl)
    float sigma_max;            // INC: max of the sigma profile
    float w_inv_scale;
    int H, W;
    int rows;                    // output rows per strip (even): short strips for small batches, one per CTA slot
    int nsx, nsy, total_strips;  // strips per row / per column of one sample, and over the whole batch
    int bal;                     // != 0: balanced strips (common.cuh: balanced_strip) over the rows of bal = batch * nsx column strips
};

struct Strip {
    int x0, y0, R, NP;
    size_t img;
};
__device__ __forceinline__ Strip strip_of(int st, const Args& a) {
    Strip g;
    const int sx = st % a.nsx, r = st / a.nsx;
    const int sy = r % a.nsy, b = r / a.nsy;
    g.x0 = sx * CW;
    g.y0 = sy * a.rows;
    g.R = min(a.rows, a.H - g.y0);     // even: H and rows are even
    g.NP = (g.R + 2) / 2;              // input row pairs incl. halo; input row k = 2j + t is image row y0 - 1 + k
    g.img = (size_t)b * a.H * a.W;
    return g;
}
constexpr int BAL_PAD = 6;     // a strip start costs ~6 row steps (2 halo rows + pipeline fill)
// strip i of this CTA; false when it has none
__device__ __forceinline__ bool strip_at(const Args& a, int i, Strip& g) {
    if (a.bal == 0) {
        const int st = (int)blockIdx.x + i * (int)gridDim.x;
        if (st >= a.total_strips) return false;
        g = strip_of(st, a);
        return true;
    }
    int bi, y0, R;
    if (!balanced_strip(a.bal, a.H, BAL_PAD, i, bi, y0, R)) return false;
    const int b = bi / a.nsx, sx = bi - b * a.nsx;
    g.x0 = sx * CW;
    g.y0 = y0;
    g.R = R;
    g.NP = (R + 2) / 2;
    g.img = (size_t)b * a.H * a.W;
    return true;
}

__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
    if (done) return true;
#pragma unroll 1
    for (int spin = 0; spin < SPIN_LIMIT && !done; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
    }
    return done != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}


// tcgen05.mma, kind::f16, accumulate.  Descriptors are passed as (low word, high word): only the 14-bit start-address
// field in the low word changes from MMA to MMA.  The MMA-issuing warp runs its loop with warp-UNIFORM control flow
// and values (whole warp, one elected lane around the asm), so ptxas keeps descriptors in uniform registers and emits
// back-to-back UTCHMMA; a lone `if (lane == 0)` thread makes it wrap every MMA in an R2UR waterfall loop (~20
// dependent instructions, ~200 cycles per MMA: the old serial bottleneck of these kernels).
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(1u));
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(bar)));
}

template <int SRC, bool PRELU, int EPI>
__global__ void __launch_bounds__(THREADS, 2) conv3x3_tcr_kernel(Args a) {
    constexpr int G = groups_of(SRC);
    constexpr int SRP = op_ring(SRC) / 2;      // operand ring depth in row pairs
    constexpr int NSP = stage_ring(SRC) / 2;   // staging ring depth in row pairs
    constexpr int NPB = 8;                     // output-pair barriers: 16 accumulator units = 8 row pairs
    constexpr int NDB = 16;                    // input-pair completion barriers (deeper: halo pairs make input run ahead)
    // f16 x f16 -> f32, both K-major; N field added per MMA.  M = 128 pixels of a row, or M = 64 when the image is no wider
    // than 64 pixels (half the A-operand fetch; the accumulator rows then sit in lanes 0..15 of every 32-lane TMEM
    // quadrant: row i -> lane 32 (i / 16) + i % 16, cute::tmem_frg_1sm "half subpartitions" layout).
    const bool m64 = a.W <= 64;
    const uint32_t kIdescBase = (1u << 4) | ((m64 ? (64u >> 4) : (128u >> 4)) << 24);
    extern __shared__ __align__(128) uint8_t smem_tcr[];
    uint8_t* stage = smem_tcr;                                                    // [NSP][2 rows] fp32 row segments (TMA destination)
    uint8_t* ring = stage + (size_t)NSP * 2 * stage_bytes(SRC);                   // [SRP][2 rows][G][2][PS] x 16 B operand rows
    uint8_t* bsm = ring + (size_t)SRP * 2 * slot_bytes(SRC);                      // [G][3][1536]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + G * 3 * BROW_BYTES);
    uint64_t* smem_full = bars;                  // [SRP]  2 x 136 converter arrivals (operand rows of a pair written)
    uint64_t* pair_done = smem_full + SRP;       // [NDB]  1 (tcgen05.commit): MMAs of input row pair j done
    uint64_t* tmem_empty = pair_done + NDB;      // [NPB]  128 epilogue arrivals: accumulators of output pair read + zeroed
    uint64_t* stage_full = tmem_empty + NPB;     // [NSP]  1 arrival + TMA transaction bytes
    uint64_t* stage_empty = stage_full + NSP;    // [NSP]  2 x 136 converter arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_empty + NSP);
    float* cst = reinterpret_cast<float*>(tmem_slot + 4);   // epilogue constants: bias[8] wo0[8] wo1[8] bo0 bo1 slope

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, W = a.W;

    // ---- setup ------------------------------------------------------------------------------------------
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < SRP; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(smem_full + i)), "r"(2 * PS));
        for (int i = 0; i < NDB; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(pair_done + i)));
        for (int i = 0; i < NPB; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(tmem_empty + i)), "r"(EPI_WARPS * 32));
        for (int i = 0; i < NSP; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(stage_full + i)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(stage_empty + i)), "r"(2 * PS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    {
        const uint4* bg = reinterpret_cast<const uint4*>(a.bmat);
        uint4* bs = reinterpret_cast<uint4*>(bsm);
        for (int i = tid; i < G * 3 * BROW_BYTES / 16; i += THREADS) bs[i] = __ldg(bg + i);
        if (tid < 8) cst[tid] = __ldg(a.bias + tid);
        if (EPI == EPI_OUTC && tid >= 32 && tid < 48) cst[8 + tid - 32] = __ldg(a.wo + tid - 32);
        if (EPI == EPI_OUTC && tid >= 64 && tid < 66) cst[24 + tid - 64] = __ldg(a.bo + tid - 64);
        if (PRELU && tid == 96) cst[26] = __ldg(a.slope);
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    if (warp > MMA_WARP && warp < TMA_WARP) {   // epilogue warps zero their lane quadrant of all accumulators
        const uint32_t z = 0u;
#pragma unroll 1
        for (int u = 0; u < TR; u++) {
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(u * NC);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    // everything above touched only this CTA's shared memory / TMEM and the constant weight images: from here on the kernel
    // reads what earlier kernels of the iteration wrote (common.cuh: HN_LAUNCH_PDL)
    pdl_wait();
    if (a.pdl_trig) pdl_trigger();
    // block scale of the activations (see conv_tc.cuh): x' = x * 2^sa, max|x'| in [2^13, 2^14)
    float amax;
    if constexpr (SRC == SRC_INC) {
        amax = fmaxf(fmaxf(__uint_as_float(ld_fresh(a.amax_in0)), 1e3f * __uint_as_float(ld_fresh(a.amax_in1))), a.sigma_max);
    } else {
        unsigned mb = ld_fresh(a.amax_in0);
        if (G == 2) mb = max(mb, ld_fresh(a.amax_in1));
        amax = __uint_as_float(mb);
    }
    int e = (int)((__float_as_uint(amax) >> 23) & 0xffu);
    if (e < 40 || e > 250) e = 127;
    const float mult = __uint_as_float((uint32_t)(267 - e) << 23);
    const float out_scale = __uint_as_float((uint32_t)(e - 13) << 23) * a.w_inv_scale;

    // Running counters (identical in every role because all roles walk the same strips in the same order):
    //   gj = input row pairs processed so far, go = output rows processed so far.
    bool ok = true;
    if (warp == TMA_WARP) {
        // =============================== TMA issuer ===============================================================
        if (lane == 0) {
            int gj = 0;
#pragma unroll 1
            for (int si = 0; ok; si++) {
                Strip g;
                if (!strip_at(a, si, g)) break;
                const int x0 = g.x0;
                // valid pixel range of a row segment: 8-channel sources start at x0-1, 2-channel ones at x0-2 (16-byte alignment)
                const int lo8 = max(0, x0 - 1), hi8 = min(W, x0 + CW + 1);
                const int lo2 = max(0, x0 - 2), hi2 = min(W, x0 + CW + 2);
                const uint32_t b8 = (uint32_t)(hi8 - lo8) * 32u, b2 = (uint32_t)(hi2 - lo2) * 8u;
                const uint32_t row_bytes = SRC == SRC_INC ? 2 * b2 : SRC == SRC_A8 ? b8 : SRC == SRC_A8_B8 ? 2 * b8 : b8 + b2;
#pragma unroll 1
                for (int j = 0; j < g.NP; j++, gj++) {
                    const int sidx = gj % NSP;
                    if (!mbar_wait(stage_empty + sidx, ((uint32_t)(gj / NSP) & 1u) ^ 1u)) { ok = false; break; }
                    const int gy0 = g.y0 - 1 + 2 * j;
                    const bool v0 = gy0 >= 0 && gy0 < H, v1 = gy0 + 1 >= 0 && gy0 + 1 < H;
                    if (!v0 && !v1) {
                        mbar_arrive(stage_full + sidx);            // zero padding rows only: nothing to copy
                        continue;
                    }
                    mbar_arrive_expect_tx(stage_full + sidx, ((uint32_t)v0 + (uint32_t)v1) * row_bytes);
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        if (!(t == 0 ? v0 : v1)) continue;
                        uint8_t* dst = stage + (size_t)(sidx * 2 + t) * stage_bytes(SRC);
                        const size_t rowpix = g.img + (size_t)(gy0 + t) * W;
                        if constexpr (SRC == SRC_INC) {
                            tma_load_1d(dst + (lo2 - (x0 - 2)) * 8, a.inA + (rowpix + lo2) * 2, b2, stage_full + sidx);
                            tma_load_1d(dst + ST2 + (lo2 - (x0 - 2)) * 8, a.inB + (rowpix + lo2) * 2, b2, stage_full + sidx);
                        } else if constexpr (SRC == SRC_A8) {
                            tma_load_1d(dst + (lo8 - (x0 - 1)) * 32, a.inA + (rowpix + lo8) * 8, b8, stage_full + sidx);
                        } else if constexpr (SRC == SRC_A8_B8) {
                            tma_load_1d(dst + (lo8 - (x0 - 1)) * 32, a.inA + (rowpix + lo8) * 8, b8, stage_full + sidx);
                            tma_load_1d(dst + ST8 + (lo8 - (x0 - 1)) * 32, a.inB + (rowpix + lo8) * 8, b8, stage_full + sidx);
                        } else {   // SRC_A8_B2
                            tma_load_1d(dst + (lo8 - (x0 - 1)) * 32, a.inA + (rowpix + lo8) * 8, b8, stage_full + sidx);
                            tma_load_1d(dst + ST8 + (lo2 - (x0 - 2)) * 8, a.inB + (rowpix + lo2) * 2, b2, stage_full + sidx);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp < PROD_WARPS) {
        // =============================== converters ===============================================================
        const int team = tid / TEAM, p = tid - team * TEAM;     // team t converts row 2j + t; p: position in the row (136 active)
        const int sw16 = ((p >> 2) & 1) * 16;
        if (p < PS) {
            int gj = 0;
#pragma unroll 1
            for (int si = 0; ok; si++) {
                Strip g;
                if (!strip_at(a, si, g)) break;
                const int gx = g.x0 - 1 + p;
                const bool colok = (p < CW + 2) && gx >= 0 && gx < W;
#pragma unroll 1
                for (int j = 0; j < g.NP; j++, gj++) {
                    const int sidx = gj % NSP, s = gj % SRP;
                    const int gy = g.y0 - 1 + 2 * j + team;
                    float g0[8];
                    float g1[G == 2 ? 8 : 1];
#pragma unroll
                    for (int c = 0; c < 8; c++) g0[c] = 0.f;
#pragma unroll
                    for (int c = 0; c < (G == 2 ? 8 : 1); c++) g1[c] = 0.f;
                    if (!mbar_wait(stage_full + sidx, (uint32_t)(gj / NSP) & 1u)) { ok = false; break; }
                    if (colok && gy >= 0 && gy < H) {
                        const uint8_t* src = stage + (size_t)(sidx * 2 + team) * stage_bytes(SRC);
                        if constexpr (SRC == SRC_INC) {
                            const float2 u2 = *reinterpret_cast<const float2*>(src + (p + 1) * 8);
                            const float2 r2 = *reinterpret_cast<const float2*>(src + ST2 + (p + 1) * 8);
                            g0[0] = u2.x; g0[1] = u2.y;
                            g0[2] = 1e3f * r2.x; g0[3] = 1e3f * r2.y;                             // hybridnet.py:566
                            g0[4] = __ldg(a.sigma + gx); g0[5] = __ldg(a.sigma + gy);
                        } else {
                            // halves read in an order that alternates every 4 lanes: conflict-free LDS.128 at a 32-byte stride
                            const float4 qa = *reinterpret_cast<const float4*>(src + p * 32 + sw16), qb = *reinterpret_cast<const float4*>(src + p * 32 + (sw16 ^ 16));
                            const float4 q0 = sw16 ? qb : qa, q1 = sw16 ? qa : qb;
                            g0[0] = q0.x; g0[1] = q0.y; g0[2] = q0.z; g0[3] = q0.w;
                            g0[4] = q1.x; g0[5] = q1.y; g0[6] = q1.z; g0[7] = q1.w;
                            if constexpr (SRC == SRC_A8_B8) {
                                const float4 sa = *reinterpret_cast<const float4*>(src + ST8 + p * 32 + sw16);
                                const float4 sb = *reinterpret_cast<const float4*>(src + ST8 + p * 32 + (sw16 ^ 16));
                                const float4 s0 = sw16 ? sb : sa, s1 = sw16 ? sa : sb;
                                g1[0] = s0.x; g1[1] = s0.y; g1[2] = s0.z; g1[3] = s0.w;
                                g1[4] = s1.x; g1[5] = s1.y; g1[6] = s1.z; g1[7] = s1.w;
                            } else if constexpr (SRC == SRC_A8_B2) {
                                const float2 sv = *reinterpret_cast<const float2*>(src + ST8 + (p + 1) * 8);
                                g1[0] = sv.x; g1[1] = sv.y;
                            }
                        }
                    }
                    // operand pair slot s was last used by pair gj - SRP: free once that pair's MMAs completed
                    if (gj >= SRP && !mbar_wait(pair_done + ((gj - SRP) & (NDB - 1)), (uint32_t)((gj - SRP) / NDB) & 1u)) { ok = false; break; }
                    uint4* slot = reinterpret_cast<uint4*>(ring + (size_t)(s * 2 + team) * slot_bytes(SRC));
                    uint4 hi, lo;
                    tc::split8(g0, mult, hi, lo);
                    slot[p] = hi;
                    slot[PS + p] = lo;
                    if constexpr (G == 2) {
                        tc::split8(g1, mult, hi, lo);
                        slot[2 * PS + p] = hi;
                        slot[3 * PS + p] = lo;
                    }
                    asm volatile("fence.proxy.async.shared::cta;");
                    mbar_arrive(smem_full + s);
                    mbar_arrive(stage_empty + sidx);             // staging rows consumed (their values went through registers)
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // =============================== MMA issuer ===============================================================
        // input row k (image row y0-1+k) feeds output rows y = k - dy, dy = 0..2; the accumulator of output row y lives
        // in unit 15 - (y & 15), so rows k, k-1, k-2 occupy ascending adjacent units (split in two MMAs at the ring wrap).
        // Whole warp, warp-uniform control flow; one elected lane issues (see mma_f16).
        {
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint64_t da = tc::smem_desc(tc::smem_u32(ring), PS * 16, 128);
            const uint64_t db = tc::smem_desc(tc::smem_u32(bsm), 128, 256);
            const uint32_t a_lo0 = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
            constexpr uint32_t kSlot16 = (uint32_t)(slot_bytes(SRC) >> 4);      // operand row pitch in 16-byte units
            constexpr uint32_t kGroup16 = 2 * PS, kB16 = BROW_BYTES / 16;
            const uint32_t kIdesc48 = kIdescBase | (6u << 17);
            int gj = 0, go = 0;
#pragma unroll 1
            for (int si = 0; ok; si++) {
                Strip gs;
                if (!strip_at(a, si, gs)) break;
                const int R = gs.R;
#pragma unroll 1
                for (int j = 0; j < gs.NP; j++, gj++) {
                    const int s = gj % SRP;
                    bool w = mbar_wait(smem_full + s, (uint32_t)(gj / SRP) & 1u);
                    // the accumulators this pair touches first (output rows 2j, 2j+1) must have been drained and zeroed
                    const int gop = (go >> 1) + j;           // global output pair index
                    if (2 * j < R) w = mbar_wait(tmem_empty + (gop & (NPB - 1)), ((uint32_t)(gop / NPB) & 1u) ^ 1u) && w;
                    if (!__all_sync(0xffffffffu, w)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < 2; t++) {
                            const int k = 2 * j + t;
                            const int gk = go + k;               // global index of output row y = k (dy = 0)
                            const uint32_t a_lo = a_lo0 + (uint32_t)(s * 2 + t) * kSlot16;
                            if (k >= 2 && k < R && (gk & 15) >= 2) {
                                // common case: three adjacent accumulators (rows k, k-1, k-2), one N = 48 MMA per (group, dx)
                                const uint32_t d_tmem = tb + (uint32_t)((15 - (gk & 15)) * NC);
#pragma unroll
                                for (int g = 0; g < G; g++)
#pragma unroll
                                    for (int dx = 0; dx < 3; dx++)
                                        mma_f16(d_tmem, a_lo + g * kGroup16 + dx, a_hi, b_lo + (uint32_t)(g * 3 + dx) * kB16, b_hi, kIdesc48);
                            } else {
                                // strip edges (fewer than three live output rows) and the ring wrap: contiguous sub-ranges of dy
                                const int dlo = max(0, k - (R - 1)), dhi = min(2, k);
                                int dy = dlo;
                                while (dy <= dhi) {
                                    const int u = 15 - ((gk - dy) & 15);
                                    int len = 1;
                                    while (dy + len <= dhi && u + len <= 15) len++;
                                    const uint32_t d_tmem = tb + (uint32_t)(u * NC);
                                    const uint32_t idesc = kIdescBase | ((uint32_t)(2 * len) << 17);          // N = 16 * len
#pragma unroll
                                    for (int g = 0; g < G; g++)
#pragma unroll
                                        for (int dx = 0; dx < 3; dx++)
                                            mma_f16(d_tmem, a_lo + g * kGroup16 + dx, a_hi, b_lo + (uint32_t)(g * 3 + dx) * kB16 + (uint32_t)(dy * 32),
                                                    b_hi, idesc);
                                    dy += len;
                                }
                            }
                        }
                        // one commit per pair: frees operand pair slot s AND publishes the accumulators
                        mma_commit(pair_done + (gj & (NDB - 1)));
                    }
                    __syncwarp();
                }
                go += R;
            }
        }
    } else {
        // =============================== epilogue ===============================================================
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
        const float slope = PRELU ? cst[26] : 0.f;
        float lmax = 0.f;
        int gj = 0, go = 0;
#pragma unroll 1
        for (int si = 0; ok; si++) {
            Strip gs;
            if (!strip_at(a, si, gs)) break;
            // M = 64: quadrant q holds pixels 16 q .. 16 q + 15 in its first 16 lanes, the other lanes are idle
            const int gx = m64 ? (lane < 16 ? gs.x0 + quad * 16 + lane : W) : gs.x0 + quad * 32 + lane;
            const int y0 = gs.y0;
            const size_t img = gs.img;
#pragma unroll 1
            for (int jo = 0; jo < gs.R / 2; jo++) {
                const int ya = 2 * jo;                         // output rows ya, ya+1 need input rows up to ya+3 = pair jo+1
                float2 wfa = make_float2(0.f, 0.f), wfb = wfa;
                if (EPI == EPI_OUTC && a.dwf_out == nullptr && gx < W) {   // issue the wavefield loads before waiting on the MMAs
                    wfa = reinterpret_cast<const float2*>(a.wf)[img + (size_t)(y0 + ya) * W + gx];
                    wfb = reinterpret_cast<const float2*>(a.wf)[img + (size_t)(y0 + ya + 1) * W + gx];
                }
                const int gjd = gj + jo + 1;                   // global input pair that completes these rows
                if (!mbar_wait(pair_done + (gjd & (NDB - 1)), (uint32_t)(gjd / NDB) & 1u)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;");
                uint32_t v[2][16];
                // rows ya (even) and ya+1 sit in adjacent units: unit(ya+1) = unit(ya) - 1
                const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((15 - ((go + ya + 1) & 15)) * NC);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[1][0]), "=r"(v[1][1]), "=r"(v[1][2]), "=r"(v[1][3]), "=r"(v[1][4]), "=r"(v[1][5]), "=r"(v[1][6]), "=r"(v[1][7]),
                               "=r"(v[1][8]), "=r"(v[1][9]), "=r"(v[1][10]), "=r"(v[1][11]), "=r"(v[1][12]), "=r"(v[1][13]), "=r"(v[1][14]), "=r"(v[1][15])
                             : "r"(tb));
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0][0]), "=r"(v[0][1]), "=r"(v[0][2]), "=r"(v[0][3]), "=r"(v[0][4]), "=r"(v[0][5]), "=r"(v[0][6]), "=r"(v[0][7]),
                               "=r"(v[0][8]), "=r"(v[0][9]), "=r"(v[0][10]), "=r"(v[0][11]), "=r"(v[0][12]), "=r"(v[0][13]), "=r"(v[0][14]), "=r"(v[0][15])
                             : "r"(tb + NC));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                {
                    const uint32_t z = 0u;   // hand the accumulators back zeroed: the MMAs always accumulate
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tb), "r"(z));
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                const int gop = (go >> 1) + jo;
                mbar_arrive(tmem_empty + (gop & (NPB - 1)));
                if (gx < W) {
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        float o[8];
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            o[c] = fmaf(fmaf(__uint_as_float(v[t][8 + c]), 1.f / 2048.f, __uint_as_float(v[t][c])), out_scale, cst[c]);
                            if (PRELU) o[c] = o[c] >= 0.f ? o[c] : slope * o[c];
                        }
                        const size_t pix = img + (size_t)(y0 + ya + t) * W + gx;
                        if (EPI == EPI_STORE) {
#pragma unroll
                            for (int c = 0; c < 8; c++) lmax = fmaxf(lmax, fabsf(o[c]));
                            float4* dst = reinterpret_cast<float4*>(a.out + pix * 8);
                            st_nhwc8(reinterpret_cast<float*>(dst), o);
                        } else if (EPI == EPI_STORE2) {
                            lmax = fmaxf(lmax, fmaxf(fabsf(o[0]), fabsf(o[1])));
                            reinterpret_cast<float2*>(a.out)[pix] = make_float2(o[0], o[1]);
                        } else {
                            float o0 = cst[24], o1 = cst[25];
#pragma unroll
                            for (int c = 0; c < 8; c++) {
                                o0 = fmaf(o[c], cst[8 + c], o0);
                                o1 = fmaf(o[c], cst[16 + c], o1);
                            }
                            if (a.dwf_out != nullptr) {
                                reinterpret_cast<float2*>(a.dwf_out)[pix] = make_float2(o0, o1);
                            } else {
                                const float2 u = t == 0 ? wfa : wfb;
                                const float2 nw = make_float2(__fdividef(o0, 1e3f) + u.x, __fdividef(o1, 1e3f) + u.y);   // hybridnet.py:570 (d / 1e3 + wf)
                                reinterpret_cast<float2*>(a.wf)[pix] = nw;
                                lmax = fmaxf(lmax, fmaxf(fabsf(nw.x), fabsf(nw.y)));
                            }
                        }
                    }
                }
            }
            gj += gs.NP;
            go += gs.R;
        }
        publish_amax(a.amax_out, lmax);
    }
    if (!ok) *a.error_flag = 1;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

}  // namespace tcr
}  // namespace hn
