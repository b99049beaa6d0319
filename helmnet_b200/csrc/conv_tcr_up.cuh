// conv_tcr_up.cuh -- the decoder's up-sampling ConvTranspose2d(8, 8, kernel 8, stride 2, padding 3), weight layout
// (in, out, kh, kw) (helmnet/architectures.py:373-385), as a row-streaming tcgen05 implicit GEMM.
//
//   out[oy][ox][co] = b[co] + sum in[iy][ix][ci] * W[ci][co][ky][kx],   oy = 2 iy - 3 + ky,  ox = 2 ix - 3 + kx
//
// GEMM mapping: M = 128 consecutive low-resolution columns ("cells") j'; a cell owns the two output columns
// 2 j' + px.  For a start-address shift s = ix - j' in [-2, 2] the taps are kx = px + 3 - 2 s (when in 0..7), and
// input row iy feeds the EIGHT output rows oy = 2 iy - 3 + ky, ky = 0..7.  All of that goes into N:
//   one MMA per (input row, shift s) with N = 8 ky x 2 px x (8 + 8) = 256 accumulator columns,
// which land on the adjacent 32-column accumulators [px0: g1|g2][px1: g1|g2] of those eight output rows
// (output-stationary ring of 16 rows in all 512 TMEM columns -> one persistent CTA per SM).
// 5 MMAs per input row produce 2 finished output rows of 256 pixels: ~1.3 tensor cycles per output pixel
// versus 1024 MACs per pixel on the CUDA cores.
// Pipeline roles as in conv_tcr.cuh; one step = two input rows = four output rows.
#pragma once
#include "conv_tcr.cuh"

namespace hn {
namespace tcu {

using tcr::CW;
using tcr::PS;
using tcr::TEAM;
using tcr::PROD_WARPS;
using tcr::EPI_WARPS;
using tcr::MMA_WARP;
using tcr::TMA_WARP;
using tcr::THREADS;
using tcr::mbar_wait;
using tcr::mbar_arrive;
using tcr::mbar_arrive_expect_tx;
using tcr::tma_load_1d;

constexpr int ROWS_I = 16;                 // default input rows per strip (Args::rows_i; the host picks it per launch)
constexpr int UC = 32;                     // accumulator columns per output row: 2 px x (g1 8 | g2 8)
constexpr int TMEM_COLS = 512;
constexpr int SRP = 4;                     // operand ring depth (input row pairs)
constexpr int NSP = 4;                     // staging ring depth (input row pairs)
constexpr int NEB = 4;                     // epilogue-step barriers: 16 rows / 4 rows per step
constexpr int NDB = 16;                    // input-pair completion barriers
constexpr int ROW_OP_BYTES = 2 * PS * 16;  // hi plane + lo plane
constexpr int ROW_ST_BYTES = 4224;         // 132 pixels x 32 B
constexpr int BIMG_BYTES = 8192;           // one shift: 256 x 16 fp16
constexpr size_t SMEM_BYTES = 120 * 1024;  // > half an SM's shared memory: exactly one CTA per SM (it owns all of TMEM)
static_assert((size_t)NSP * 2 * ROW_ST_BYTES + (size_t)SRP * 2 * ROW_OP_BYTES + 5 * BIMG_BYTES + 768 <= SMEM_BYTES, "smem");

struct Args {
    const float* in;            // NHWC8 [B][Hi][Wi]
    const __half* bmat;         // [5 shifts] x 8192 B canonical K-major images (host packed)
    float bias[8];              // launch parameter: read from the constant bank by the epilogue
    float* out;                 // NHWC8 [B][2Hi][2Wi]
    const unsigned* amax_in;
    unsigned* amax_out;
    int* error_flag;
    int pdl_trig;               // PDL: let the next kernel's CTAs become resident as this grid's CTAs exit (hn_ctx::pdl)
    float w_inv_scale;
    int Hi, Wi;
    int nsx, nsy, total_strips;
    int rows_i;                 // input rows per strip (even)
    int bal;                    // != 0 (nsx == 1 only): balanced strips over the input rows of `bal` images (common.cuh: balanced_strip)
    // Narrow images side by side (input width <= 62, nsx == 1): `pack` images share ONE M = 128 MMA (see conv_tcr_down.cuh).  Image k
    // of a group owns the operand entries [k S, k S + S), S = Wi + 2: two zero cells, then its Wi cells; GEMM row m = k S + cell.
    int pack, pack_s, batch;
};

struct Strip {
    int x0, iy0, Ri, NP;
    int nimg;                   // images of this strip's group (1 without packing)
    size_t img_in, img_out;     // first pixel of the (first) image
};
__device__ __forceinline__ Strip strip_of(int st, const Args& a) {
    Strip g;
    const int sx = st % a.nsx, r = st / a.nsx;
    const int sy = r % a.nsy, b = r / a.nsy;
    g.x0 = sx * CW;
    g.iy0 = sy * a.rows_i;
    g.Ri = min(a.rows_i, a.Hi - g.iy0);   // even
    g.NP = (g.Ri + 4) / 2;                // input rows k = 0 .. Ri + 3, image row iy0 - 2 + k
    const int img0 = a.pack > 0 ? b * a.pack : b;
    g.nimg = a.pack > 0 ? min(a.pack, a.batch - img0) : 1;
    g.img_in = (size_t)img0 * a.Hi * a.Wi;
    g.img_out = g.img_in * 4;
    return g;
}
constexpr int BAL_PAD = 6;     // a strip start costs ~6 row steps (4 extra input rows + fill)
// strip i of this CTA; false when it has none
__device__ __forceinline__ bool strip_at(const Args& a, int i, Strip& g) {
    if (a.bal == 0) {
        const int st = (int)blockIdx.x + i * (int)gridDim.x;
        if (st >= a.total_strips) return false;
        g = strip_of(st, a);
        return true;
    }
    int b, iy0, Ri;
    if (!balanced_strip(a.bal, a.Hi, BAL_PAD, i, b, iy0, Ri)) return false;
    g.x0 = 0;
    g.iy0 = iy0;
    g.Ri = Ri;
    g.NP = (Ri + 4) / 2;
    const int img0 = a.pack > 0 ? b * a.pack : b;
    g.nimg = a.pack > 0 ? min(a.pack, a.batch - img0) : 1;
    g.img_in = (size_t)img0 * a.Hi * a.Wi;
    g.img_out = g.img_in * 4;
    return true;
}

__global__ void __launch_bounds__(THREADS, 1) up_tcr_kernel(Args a) {
    // M = 64 when the input is no wider than 64 cells (accumulator row i then sits in lane 32 (i / 16) + i % 16, see conv_tcr.cuh)
    const bool m64 = a.Wi <= 64 && a.pack == 0;
    const uint32_t kIdescBase = (1u << 4) | ((m64 ? (64u >> 4) : (128u >> 4)) << 24);
    extern __shared__ __align__(128) uint8_t smem_tcu[];
    uint8_t* stage = smem_tcu;
    uint8_t* ring = stage + (size_t)NSP * 2 * ROW_ST_BYTES;
    uint8_t* bsm = ring + (size_t)SRP * 2 * ROW_OP_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + 5 * BIMG_BYTES);
    uint64_t* smem_full = bars;                  // [SRP]  2 x 136
    uint64_t* pair_done = smem_full + SRP;       // [NDB]  tcgen05.commit
    uint64_t* tmem_empty = pair_done + NDB;      // [NEB]  128 epilogue arrivals per 4 output rows
    uint64_t* stage_full = tmem_empty + NEB;     // [NSP]
    uint64_t* stage_empty = stage_full + NSP;    // [NSP]  2 x 136
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_empty + NSP);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Hi = a.Hi, Wi = a.Wi, Wo = 2 * Wi;

    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < SRP; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(smem_full + i)), "r"(2 * PS));
        for (int i = 0; i < NDB; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(pair_done + i)));
        for (int i = 0; i < NEB; i++)
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
        for (int i = tid; i < 5 * BIMG_BYTES / 16; i += THREADS) bs[i] = __ldg(bg + i);
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
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(u * UC);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z));
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
                const int lo = max(0, g.x0 - 2), hi = min(Wi, g.x0 + CW + 2);
                const uint32_t rb = (uint32_t)(hi - lo) * 32u;
#pragma unroll 1
                for (int j = 0; j < g.NP; j++, gj++) {
                    const int sidx = gj % NSP;
                    if (!mbar_wait(stage_empty + sidx, ((uint32_t)(gj / NSP) & 1u) ^ 1u)) { ok = false; break; }
                    const int gy0 = g.iy0 - 2 + 2 * j;
                    const bool v0 = gy0 >= 0 && gy0 < Hi, v1 = gy0 + 1 >= 0 && gy0 + 1 < Hi;
                    if (!v0 && !v1) {
                        mbar_arrive(stage_full + sidx);
                        continue;
                    }
                    mbar_arrive_expect_tx(stage_full + sidx, ((uint32_t)v0 + (uint32_t)v1) * rb * (uint32_t)g.nimg);
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        if (!(t == 0 ? v0 : v1)) continue;
                        uint8_t* dst = stage + (size_t)(sidx * 2 + t) * ROW_ST_BYTES;
                        // (packed: image k's cell 0 sits at entry k S + 2 of the staged row)
                        for (int k = 0; k < g.nimg; k++)
                            tma_load_1d(dst + (size_t)k * a.pack_s * 32 + (lo - (g.x0 - 2)) * 32,
                                        a.in + (g.img_in + (size_t)k * Hi * Wi + (size_t)(gy0 + t) * Wi + lo) * 8, rb, stage_full + sidx);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp < PROD_WARPS) {
        // =============================== converters ===============================================================
        const int team = tid / TEAM, p = tid - team * TEAM;     // p: position, cell x0 - 2 + p
        const int sw16 = ((p >> 2) & 1) * 16;
        if (p < PS) {
            int gj = 0;
#pragma unroll 1
            for (int si = 0; ok; si++) {
                Strip g;
                if (!strip_at(a, si, g)) break;
                const int kimg = a.pack > 0 ? p / a.pack_s : 0;                 // packed: entry p belongs to image kimg of the group
                const int gx = g.x0 - 2 + (p - kimg * a.pack_s);
                const bool colok = (p < CW + 4) && kimg < g.nimg && gx >= 0 && gx < Wi;
#pragma unroll 1
                for (int j = 0; j < g.NP; j++, gj++) {
                    const int sidx = gj % NSP, s = gj % SRP;
                    const int gy = g.iy0 - 2 + 2 * j + team;
                    float g0[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) g0[c] = 0.f;
                    if (!mbar_wait(stage_full + sidx, (uint32_t)(gj / NSP) & 1u)) { ok = false; break; }
                    if (colok && gy >= 0 && gy < Hi) {
                        const uint8_t* src = stage + (size_t)(sidx * 2 + team) * ROW_ST_BYTES + (size_t)p * 32;
                        // halves read in an order that alternates every 4 lanes: conflict-free LDS.128 at a 32-byte stride
                        const float4 qa = *reinterpret_cast<const float4*>(src + sw16), qb = *reinterpret_cast<const float4*>(src + (sw16 ^ 16));
                        const float4 q0 = sw16 ? qb : qa, q1 = sw16 ? qa : qb;
                        g0[0] = q0.x; g0[1] = q0.y; g0[2] = q0.z; g0[3] = q0.w; g0[4] = q1.x; g0[5] = q1.y; g0[6] = q1.z; g0[7] = q1.w;
                    }
                    if (gj >= SRP && !mbar_wait(pair_done + ((gj - SRP) & (NDB - 1)), (uint32_t)((gj - SRP) / NDB) & 1u)) { ok = false; break; }
                    uint4* slot = reinterpret_cast<uint4*>(ring + (size_t)(s * 2 + team) * ROW_OP_BYTES);
                    uint4 hi4, lo4;
                    tc::split8(g0, mult, hi4, lo4);
                    slot[p] = hi4;
                    slot[PS + p] = lo4;
                    asm volatile("fence.proxy.async.shared::cta;");
                    mbar_arrive(smem_full + s);
                    mbar_arrive(stage_empty + sidx);
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // =============================== MMA issuer ===============================================================
        // input row k (image row iy0-2+k) feeds output rows oyl = 2k - 7 + ky (strip-local), ky = 0..7, ascending units.
        // Whole warp, warp-uniform control flow; one elected lane issues (see tcr::mma_f16).
        {
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint64_t da = tc::smem_desc(tc::smem_u32(ring), PS * 16, 128);     // + si: shift s = si - 2
            const uint64_t db = tc::smem_desc(tc::smem_u32(bsm), 128, 256);
            const uint32_t a_lo0 = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
            constexpr uint32_t kRow16 = ROW_OP_BYTES >> 4;
            int gj = 0, go = 0;   // go: output rows finished in previous strips (multiple of 4)
#pragma unroll 1
            for (int si = 0; ok; si++) {
                Strip gs;
                if (!strip_at(a, si, gs)) break;
                const int Ro = 2 * gs.Ri;
#pragma unroll 1
                for (int j = 0; j < gs.NP; j++, gj++) {
                    const int s = gj % SRP;
                    bool w = mbar_wait(smem_full + s, (uint32_t)(gj / SRP) & 1u);
                    // rows first touched by this pair (4j-1 .. 4j+2) reuse the units of rows 16 earlier: epilogue step G done?
                    const int G = (go >> 2) + j - 4;
                    if (G >= 0) w = mbar_wait(tmem_empty + (G & (NEB - 1)), (uint32_t)(G / NEB) & 1u) && w;
                    if (!__all_sync(0xffffffffu, w)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (tcr::elect_one()) {
#pragma unroll
                        for (int t = 0; t < 2; t++) {
                            const int k = 2 * j + t;
                            const uint32_t a_lo = a_lo0 + (uint32_t)(s * 2 + t) * kRow16;
                            const int klo = max(0, 7 - 2 * k), khi = min(7, Ro + 6 - 2 * k);   // 0 <= 2k-7+ky < Ro
                            int ky = klo;
                            while (ky <= khi) {
                                const int u = (go + 2 * k - 7 + ky) & 15;
                                int len = 1;
                                while (ky + len <= khi && u + len <= 15) len++;
                                const uint32_t d_tmem = tb + (uint32_t)(u * UC);
                                const uint32_t idesc = kIdescBase | ((uint32_t)(4 * len) << 17);      // N = 32 * len
#pragma unroll
                                for (int si = 0; si < 5; si++)
                                    tcr::mma_f16(d_tmem, a_lo + (uint32_t)si, a_hi, b_lo + (uint32_t)(si * (BIMG_BYTES >> 4) + ky * 64), b_hi, idesc);
                                ky += len;
                            }
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
            int cell = m64 ? (lane < 16 ? gs.x0 + quad * 16 + lane : Wi) : gs.x0 + quad * 32 + lane;
            size_t img_out = gs.img_out;
            if (a.pack > 0) {            // GEMM row m = k S + cell of image k of the group
                const int m = quad * 32 + lane, k = m / a.pack_s;
                cell = k < gs.nimg ? m - k * a.pack_s : Wi;
                img_out += (size_t)k * 4 * Hi * Wi;
            }
            const int Ro = 2 * gs.Ri;
#pragma unroll 1
            for (int es = 0; es < Ro / 4; es++) {
                const int gjd = gj + es + 2;                   // input pair that completes output rows 4es .. 4es+3
                if (!mbar_wait(pair_done + (gjd & (NDB - 1)), (uint32_t)(gjd / NDB) & 1u)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;");
                // two output rows per TMEM round trip: both loads are in flight before the one wait, and the two rows'
                // arithmetic and stores interleave (a single epilogue warp per scheduler is latency-bound otherwise)
#pragma unroll 1
                for (int r = 0; r < 4; r += 2) {
                    const int oyl = 4 * es + r;
                    uint32_t v[2][32];
                    const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16);
                    const uint32_t taddr0 = tbase + (uint32_t)(((go + oyl) & 15) * UC), taddr1 = tbase + (uint32_t)(((go + oyl + 1) & 15) * UC);
#pragma unroll
                    for (int t = 0; t < 2; t++)
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                            : "=r"(v[t][0]), "=r"(v[t][1]), "=r"(v[t][2]), "=r"(v[t][3]), "=r"(v[t][4]), "=r"(v[t][5]), "=r"(v[t][6]), "=r"(v[t][7]),
                              "=r"(v[t][8]), "=r"(v[t][9]), "=r"(v[t][10]), "=r"(v[t][11]), "=r"(v[t][12]), "=r"(v[t][13]), "=r"(v[t][14]),
                              "=r"(v[t][15]), "=r"(v[t][16]), "=r"(v[t][17]), "=r"(v[t][18]), "=r"(v[t][19]), "=r"(v[t][20]), "=r"(v[t][21]),
                              "=r"(v[t][22]), "=r"(v[t][23]), "=r"(v[t][24]), "=r"(v[t][25]), "=r"(v[t][26]), "=r"(v[t][27]), "=r"(v[t][28]),
                              "=r"(v[t][29]), "=r"(v[t][30]), "=r"(v[t][31])
                            : "r"(t == 0 ? taddr0 : taddr1));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    {
                        const uint32_t z = 0u;
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr0), "r"(z));
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr1), "r"(z));
                    }
                    if (cell < Wi) {
#pragma unroll
                        for (int t = 0; t < 2; t++) {
                            float4* dst = reinterpret_cast<float4*>(a.out + (img_out + (size_t)(2 * gs.iy0 + oyl + t) * Wo + 2 * cell) * 8);
#pragma unroll
                            for (int px = 0; px < 2; px++) {
                                float o[8];
#pragma unroll
                                for (int c = 0; c < 8; c++) {
                                    o[c] = fmaf(fmaf(__uint_as_float(v[t][px * 16 + 8 + c]), 1.f / 2048.f, __uint_as_float(v[t][px * 16 + c])), out_scale,
                                                a.bias[c]);
                                    lmax = fmaxf(lmax, fabsf(o[c]));
                                }
                                st_nhwc8(reinterpret_cast<float*>(dst + 2 * px), o);
                            }
                        }
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;");
                mbar_arrive(tmem_empty + (((go >> 2) + es) & (NEB - 1)));
            }
            gj += gs.NP;
            go += Ro;
        }
        publish_amax(a.amax_out, lmax);
    }
    if (!ok) *a.error_flag = 1;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

}  // namespace tcu
}  // namespace hn
