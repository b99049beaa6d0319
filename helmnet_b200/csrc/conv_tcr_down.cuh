// conv_tcr_down.cuh -- the encoder's down-sampling Conv2d(8, 8, kernel 8, stride 2, padding 3)
// (helmnet/architectures.py:209-211) as a row-streaming tcgen05 implicit GEMM; same machinery as conv_tcr.cuh.
//
//   out[oy][ox][co] = b[co] + sum_{ky,kx,ci} in[2 oy - 3 + ky][2 ox - 3 + kx][ci] * W[co][ci][ky][kx]
//
// GEMM mapping: M = 128 consecutive output columns ox.  For a fixed kx the inputs 2 ox - 3 + kx of consecutive ox
// are two pixels apart, so each staged input row is stored de-interleaved by x parity: plane[par][h] holds pixel
// xb + 2 h + par (xb = 2 ox0 - 4) and "A row ox, tap kx" is plane[(kx+1)&1][ox + (kx+1)/2] -- again a pure
// start-address shift of a K-major SWIZZLE_NONE operand.  Input row k = 2 j + t of a strip feeds the four output
// rows j, j-1, j-2, j-3 through ky = 2 m + t (m = 0..3), so the ky taps go into N: one MMA per (row, kx) with
// N = 4 x (8 + 8) = 64 accumulator columns that land on the adjacent 16-column accumulators of those four output
// rows (output-stationary ring of 16 units in 256 TMEM columns, zeroed by the epilogue, accumulate always on).
// 16 MMAs (N = 64, ~48 cycles each) per 128 output pixels instead of 64 MMAs in a tap-per-MMA mapping.
//
// Roles, rings and the persistent strip walk are those of conv_tcr.cuh: TMA warp -> fp32 staging ring,
// 2 converter teams (one input row each) -> fp16 hi/lo operand ring, one MMA-issuing thread, 4 epilogue warps.
#pragma once
#include "conv_tcr.cuh"

namespace hn {
namespace tcd {

using tcr::CW;
using tcr::PS;
using tcr::TEAM;
using tcr::PROD_WARPS;
using tcr::EPI_WARPS;
using tcr::MMA_WARP;
using tcr::TMA_WARP;
using tcr::THREADS;
using tcr::TMEM_COLS;
using tcr::mbar_wait;
using tcr::mbar_arrive;
using tcr::mbar_arrive_expect_tx;
using tcr::tma_load_1d;

constexpr int ROWS_O = 32;                 // default output rows per strip (Args::rows_o is picked per launch by the host)
constexpr int NC = 16;                     // accumulator columns per output row
constexpr int SRP = 2;                     // operand ring depth (input row pairs)
constexpr int NSP = 2;                     // staging ring depth (input row pairs) for images wider than 128 pixels
constexpr int NSP_MAX = 4;                 // ... and for narrower ones: half-size rows, twice the depth in the same shared memory
                                           // (the bytes a CTA keeps in flight towards HBM stay the same: at 128 pixels two
                                           // pairs of 4 KB rows per CTA covered half of the bandwidth-latency product)
constexpr int NUB = 16;                    // one barrier per accumulator unit (output row)
constexpr int NDB = 32;                    // input-pair completion barriers
constexpr int ROW_OP_BYTES = 4 * PS * 16;  // one operand row: [par 0: hi, lo][par 1: hi, lo] x PS entries
constexpr int ROW_ST_BYTES = 2 * PS * 32;  // one staged fp32 row: 2*PS pixels x 32 B (264 used)
constexpr int BIMG_BYTES = 2048;           // one (t, kx) B operand: 64 x 16 fp16
constexpr size_t SMEM_BYTES = (size_t)NSP * 2 * ROW_ST_BYTES + (size_t)SRP * 2 * ROW_OP_BYTES + 16 * BIMG_BYTES + 768;

struct Args {
    const float* in;            // NHWC8 [B][H][W]
    const __half* bmat;         // [2 t][8 kx] x 2048 B canonical K-major images (host packed)
    float bias[8];              // launch parameter: read from the constant bank by the epilogue
    float* out;                 // NHWC8 [B][H/2][W/2]
    const unsigned* amax_in;
    unsigned* amax_out;
    int* error_flag;
    int pdl_trig;               // PDL: let the next kernel's CTAs become resident as this grid's CTAs exit (hn_ctx::pdl)
    float w_inv_scale;
    int H, W;                   // input resolution
    int rows_o;                 // output rows per strip: small batches take short strips so that every SM gets one
    int nsx, nsy, total_strips;
    int bal;                    // != 0 (nsx == 1 only): balanced strips over the output rows of `bal` images (common.cuh: balanced_strip)
    // Narrow images side by side (output width <= 62, nsx == 1): `pack` images share ONE M = 128 MMA.  Image k of a group owns the
    // operand entries [k S, k S + S) of each parity plane, S = W/2 + 2: two zero entries (pixels -4 .. -1), then its W/2 data entries;
    // the two zero entries its right-hand taps need are the next image's left-hand ones.  GEMM row m = k S + ox, every tap stays a
    // pure start-address shift.  The strip walk then runs over groups of images ("image" b of strip_of / balanced_strip = group b).
    int pack, pack_s, batch;
};

struct Strip {
    int ox0, oy0, Ro, NP;
    int nimg;                   // images of this strip's group (1 without packing)
    size_t img_in, img_out;     // first pixel of the (first) image
};
__device__ __forceinline__ Strip strip_of(int st, const Args& a) {
    Strip g;
    const int sx = st % a.nsx, r = st / a.nsx;
    const int sy = r % a.nsy, b = r / a.nsy;
    g.ox0 = sx * CW;
    g.oy0 = sy * a.rows_o;
    g.Ro = min(a.rows_o, (a.H >> 1) - g.oy0);
    g.NP = g.Ro + 3;            // input row pairs: rows k = 0 .. 2 Ro + 5, image row 2 oy0 - 3 + k
    const int img0 = a.pack > 0 ? b * a.pack : b;
    g.nimg = a.pack > 0 ? min(a.pack, a.batch - img0) : 1;
    g.img_in = (size_t)img0 * a.H * a.W;
    g.img_out = (size_t)img0 * (a.H >> 1) * (a.W >> 1);
    return g;
}
constexpr int BAL_PAD = 6;     // a strip start costs ~5 row steps (3 extra input row pairs + fill)
// strip i of this CTA; false when it has none
__device__ __forceinline__ bool strip_at(const Args& a, int i, Strip& g) {
    if (a.bal == 0) {
        const int st = (int)blockIdx.x + i * (int)gridDim.x;
        if (st >= a.total_strips) return false;
        g = strip_of(st, a);
        return true;
    }
    int b, oy0, Ro;
    if (!balanced_strip(a.bal, a.H >> 1, BAL_PAD, i, b, oy0, Ro)) return false;
    g.ox0 = 0;
    g.oy0 = oy0;
    g.Ro = Ro;
    g.NP = Ro + 3;
    const int img0 = a.pack > 0 ? b * a.pack : b;
    g.nimg = a.pack > 0 ? min(a.pack, a.batch - img0) : 1;
    g.img_in = (size_t)img0 * a.H * a.W;
    g.img_out = (size_t)img0 * (a.H >> 1) * (a.W >> 1);
    return true;
}

__global__ void __launch_bounds__(THREADS, 2) down_tcr_kernel(Args a) {
    // M = 64 when the output is no wider than 64 pixels (half the A-operand fetch; accumulator row i then sits in lane
    // 32 (i / 16) + i % 16, see conv_tcr.cuh)
    const bool m64 = (a.W >> 1) <= 64 && a.pack == 0;
    const uint32_t kIdescBase = (1u << 4) | ((m64 ? (64u >> 4) : (128u >> 4)) << 24);
    extern __shared__ __align__(128) uint8_t smem_tcd[];
    uint8_t* stage = smem_tcd;                                          // [NSP][2] fp32 rows
    uint8_t* ring = stage + (size_t)NSP * 2 * ROW_ST_BYTES;             // [SRP][2] operand rows
    uint8_t* bsm = ring + (size_t)SRP * 2 * ROW_OP_BYTES;               // [2][8] B images
    uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + 16 * BIMG_BYTES);
    uint64_t* smem_full = bars;                  // [SRP]  2 x 136 converter arrivals
    uint64_t* pair_done = smem_full + SRP;       // [NDB]  tcgen05.commit
    uint64_t* tmem_empty = pair_done + NDB;      // [NUB]  128 epilogue arrivals per output row
    uint64_t* stage_full = tmem_empty + NUB;     // [NSP_MAX]
    uint64_t* stage_empty = stage_full + NSP_MAX;    // [NSP_MAX]  2 x 136
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_empty + NSP_MAX);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, W = a.W, Wo = W >> 1;
    // staging ring geometry: rows of up to 132 staged pixels fit half a slot
    const bool narrow = W <= 128 && a.pack == 0;      // (packed groups fill a whole staging row)
    const int nsp_sh = narrow ? 2 : 1, nsp_mask = (1 << nsp_sh) - 1;       // ring depth 4 or 2 (row pairs)
    const uint32_t row_st = narrow ? ROW_ST_BYTES / 2 : ROW_ST_BYTES;

    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < SRP; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(smem_full + i)), "r"(2 * PS));
        for (int i = 0; i < NDB; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(pair_done + i)));
        for (int i = 0; i < NUB; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(tmem_empty + i)), "r"(EPI_WARPS * 32));
        for (int i = 0; i < NSP_MAX; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(stage_full + i)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(stage_empty + i)), "r"(2 * PS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    {
        const uint4* bg = reinterpret_cast<const uint4*>(a.bmat);
        uint4* bs = reinterpret_cast<uint4*>(bsm);
        for (int i = tid; i < 16 * BIMG_BYTES / 16; i += THREADS) bs[i] = __ldg(bg + i);
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    if (warp > MMA_WARP && warp < TMA_WARP) {
        const uint32_t z = 0u;
#pragma unroll 1
        for (int u = 0; u < 16; u++) {
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
    const float amax = __uint_as_float(ld_fresh(a.amax_in));
    int e = (int)((__float_as_uint(amax) >> 23) & 0xffu);
    if (e < 40 || e > 250) e = 127;
    const float mult = __uint_as_float((uint32_t)(267 - e) << 23);
    const float out_scale = __uint_as_float((uint32_t)(e - 13) << 23) * a.w_inv_scale;

    bool ok = true;
    if (warp == TMA_WARP) {
        // =============================== TMA issuer ===============================================================
        if (lane == 0) {
            int gj = 0;
#pragma unroll 1
            for (int si = 0; ok; si++) {
                Strip g;
                if (!strip_at(a, si, g)) break;
                const int xb = 2 * g.ox0 - 4;                               // first staged pixel (even)
                const int lo = max(0, xb), hi = min(W, xb + 2 * PS - 8);    // 264 pixels cover every tap of 128 outputs
                const uint32_t rb = (uint32_t)(hi - lo) * 32u;
                const size_t img_px = (size_t)H * W;
#pragma unroll 1
                for (int j = 0; j < g.NP; j++, gj++) {
                    const int sidx = gj & nsp_mask;
                    if (!mbar_wait(stage_empty + sidx, ((uint32_t)(gj >> nsp_sh) & 1u) ^ 1u)) { ok = false; break; }
                    const int gy0 = 2 * g.oy0 - 3 + 2 * j;
                    const bool v0 = gy0 >= 0 && gy0 < H, v1 = gy0 + 1 >= 0 && gy0 + 1 < H;
                    if (!v0 && !v1) {
                        mbar_arrive(stage_full + sidx);
                        continue;
                    }
                    mbar_arrive_expect_tx(stage_full + sidx, ((uint32_t)v0 + (uint32_t)v1) * rb * (uint32_t)g.nimg);
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        if (!(t == 0 ? v0 : v1)) continue;
                        uint8_t* dst = stage + (size_t)(sidx * 2 + t) * row_st;
                        // (packed: image k's pixel 0 sits at entry k S + 2 of the parity planes = byte (k S + 2) * 64 of the staged row)
                        for (int k = 0; k < g.nimg; k++)
                            tma_load_1d(dst + (size_t)k * a.pack_s * 64 + (lo - xb) * 32, a.in + (g.img_in + k * img_px + (size_t)(gy0 + t) * W + lo) * 8, rb,
                                        stage_full + sidx);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp < PROD_WARPS) {
        // =============================== converters ===============================================================
        const int team = tid / TEAM, p = tid - team * TEAM;     // p: half-position h; handles pixels xb + 2p and xb + 2p + 1
        const int rot = (p >> 1) & 3;
        const bool r1 = (rot & 1) != 0, r2 = (rot & 2) != 0;
        const int pl_a = r2 ? 2 * PS : 0, pl_b = r2 ? 0 : 2 * PS;
        if (p < PS) {
            int gj = 0;
#pragma unroll 1
            for (int si = 0; ok; si++) {
                Strip g;
                if (!strip_at(a, si, g)) break;
                const int xb = 2 * g.ox0 - 4;
                const int kimg = a.pack > 0 ? p / a.pack_s : 0;                 // packed: entry p belongs to image kimg of the group
                const int gxe = xb + 2 * (p - kimg * a.pack_s), gxo = gxe + 1;
                const bool mine = (p < PS - 4) && kimg < g.nimg;
                const bool oke = mine && gxe >= 0 && gxe < W, oko = mine && gxo >= 0 && gxo < W;
#pragma unroll 1
                for (int j = 0; j < g.NP; j++, gj++) {
                    const int sidx = gj & nsp_mask, s = gj % SRP;
                    const int gy = 2 * g.oy0 - 3 + 2 * j + team;
                    float ge[8], go_[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) { ge[c] = 0.f; go_[c] = 0.f; }
                    if (!mbar_wait(stage_full + sidx, (uint32_t)(gj >> nsp_sh) & 1u)) { ok = false; break; }
                    if (gy >= 0 && gy < H && (oke || oko)) {
                        // A thread owns 64 contiguous bytes (even pixel | odd pixel) at a 64-byte stride.  Reading the four
                        // 16-byte chunks in the rotated order (k + p/2) & 3 makes the 8 lanes of a quarter warp hit 8 distinct
                        // bank groups; the rotation by 2 is undone for free by swapping the parity planes at the store below,
                        // the rotation by 1 with one select per value.
                        const uint8_t* src = stage + (size_t)(sidx * 2 + team) * row_st + (size_t)p * 64;
                        const float4 l0 = *reinterpret_cast<const float4*>(src + ((rot + 0) & 3) * 16);
                        const float4 l1 = *reinterpret_cast<const float4*>(src + ((rot + 1) & 3) * 16);
                        const float4 l2 = *reinterpret_cast<const float4*>(src + ((rot + 2) & 3) * 16);
                        const float4 l3 = *reinterpret_cast<const float4*>(src + ((rot + 3) & 3) * 16);
                        // m_k = chunk (k + 2 * (rot >> 1)) & 3
                        const float4 m0 = r1 ? l3 : l0, m1 = r1 ? l0 : l1, m2 = r1 ? l1 : l2, m3 = r1 ? l2 : l3;
                        const bool ok_a = r2 ? oko : oke, ok_b = r2 ? oke : oko;      // (m0, m1) is the odd pixel when r2
                        if (ok_a) { ge[0] = m0.x; ge[1] = m0.y; ge[2] = m0.z; ge[3] = m0.w; ge[4] = m1.x; ge[5] = m1.y; ge[6] = m1.z; ge[7] = m1.w; }
                        if (ok_b) { go_[0] = m2.x; go_[1] = m2.y; go_[2] = m2.z; go_[3] = m2.w; go_[4] = m3.x; go_[5] = m3.y; go_[6] = m3.z; go_[7] = m3.w; }
                    }
                    if (gj >= SRP && !mbar_wait(pair_done + ((gj - SRP) & (NDB - 1)), (uint32_t)((gj - SRP) / NDB) & 1u)) { ok = false; break; }
                    uint4* slot = reinterpret_cast<uint4*>(ring + (size_t)(s * 2 + team) * ROW_OP_BYTES);
                    uint4 hi4, lo4;
                    tc::split8(ge, mult, hi4, lo4);            // ge: even pixel, or the odd one when r2 (planes swapped)
                    slot[pl_a + p] = hi4;
                    slot[pl_a + PS + p] = lo4;
                    tc::split8(go_, mult, hi4, lo4);
                    slot[pl_b + p] = hi4;
                    slot[pl_b + PS + p] = lo4;
                    asm volatile("fence.proxy.async.shared::cta;");
                    mbar_arrive(smem_full + s);
                    mbar_arrive(stage_empty + sidx);
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // =============================== MMA issuer ===============================================================
        // Whole warp, warp-uniform control flow; one elected lane issues (see tcr::mma_f16).
        {
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            // per-kx A descriptor of operand row 0: parity plane (kx+1)&1, start shift (kx+1)>>1 entries
            const uint64_t da = tc::smem_desc(tc::smem_u32(ring), PS * 16, 128);
            const uint64_t db = tc::smem_desc(tc::smem_u32(bsm), 128, 256);
            const uint32_t a_lo0 = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
            constexpr uint32_t kRow16 = ROW_OP_BYTES >> 4;
            const uint32_t kIdesc64 = kIdescBase | (8u << 17);
            int gj = 0, go = 0;
#pragma unroll 1
            for (int si = 0; ok; si++) {
                Strip gs;
                if (!strip_at(a, si, gs)) break;
                const int Ro = gs.Ro;
#pragma unroll 1
                for (int j = 0; j < gs.NP; j++, gj++) {
                    const int s = gj % SRP;
                    bool w = mbar_wait(smem_full + s, (uint32_t)(gj / SRP) & 1u);
                    const int gr = go + j;                    // global index of output row j (first touched by this pair)
                    if (j < Ro) w = mbar_wait(tmem_empty + (gr & (NUB - 1)), ((uint32_t)(gr / NUB) & 1u) ^ 1u) && w;
                    if (!__all_sync(0xffffffffu, w)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (tcr::elect_one()) {
                        // live output rows of this pair: j - m, m in [mlo, mhi]
                        const int mlo = max(0, j - (Ro - 1)), mhi = min(3, j);
                        int m = mlo;
                        while (m <= mhi) {
                            const int u = 15 - ((gr - m) & 15);
                            int len = 1;
                            while (m + len <= mhi && u + len <= 15) len++;
                            const uint32_t d_tmem = tb + (uint32_t)(u * NC);
                            const uint32_t idesc = (len == 4) ? kIdesc64 : (kIdescBase | ((uint32_t)(2 * len) << 17));
#pragma unroll
                            for (int t = 0; t < 2; t++) {
                                const uint32_t a_lo = a_lo0 + (uint32_t)(s * 2 + t) * kRow16;
#pragma unroll
                                for (int kx = 0; kx < 8; kx++)
                                    tcr::mma_f16(d_tmem, a_lo + (uint32_t)(((kx + 1) & 1) * 2 * PS + ((kx + 1) >> 1)), a_hi,
                                                 b_lo + (uint32_t)((t * 8 + kx) * (BIMG_BYTES >> 4) + m * 32), b_hi, idesc);
                            }
                            m += len;
                        }
                        tcr::mma_commit(pair_done + (gj & (NDB - 1)));
                    }
                    __syncwarp();
                }
                go += Ro;
            }
        }
    } else {
        // =============================== epilogue ===============================================================
        const int quad = warp & 3;
        float lmax = 0.f;
        int gj = 0, go = 0;
#pragma unroll 1
        for (int si = 0; ok; si++) {
            Strip gs;
            if (!strip_at(a, si, gs)) break;
            int ox = m64 ? (lane < 16 ? gs.ox0 + quad * 16 + lane : Wo) : gs.ox0 + quad * 32 + lane;
            size_t img_out = gs.img_out;
            if (a.pack > 0) {            // GEMM row m = k S + ox of image k of the group
                const int m = quad * 32 + lane, k = m / a.pack_s;
                ox = k < gs.nimg ? m - k * a.pack_s : Wo;
                img_out += (size_t)k * (H >> 1) * Wo;
            }
#pragma unroll 1
            for (int oyl = 0; oyl < gs.Ro; oyl++) {
                const int gjd = gj + oyl + 3;                  // input pair that completes output row oyl
                if (!mbar_wait(pair_done + (gjd & (NDB - 1)), (uint32_t)(gjd / NDB) & 1u)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int gr = go + oyl;
                uint32_t v[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((15 - (gr & 15)) * NC);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                {
                    const uint32_t z = 0u;
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z));
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                mbar_arrive(tmem_empty + (gr & (NUB - 1)));
                if (ox < Wo) {
                    float o[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        o[c] = fmaf(fmaf(__uint_as_float(v[8 + c]), 1.f / 2048.f, __uint_as_float(v[c])), out_scale, a.bias[c]);
                        lmax = fmaxf(lmax, fabsf(o[c]));
                    }
                    float4* dst = reinterpret_cast<float4*>(a.out + (img_out + (size_t)(gs.oy0 + oyl) * Wo + ox) * 8);
                    st_nhwc8(reinterpret_cast<float*>(dst), o);
                }
            }
            gj += gs.NP;
            go += gs.Ro;
        }
        publish_amax(a.amax_out, lmax);
    }
    if (!ok) *a.error_flag = 1;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

}  // namespace tcd
}  // namespace hn
