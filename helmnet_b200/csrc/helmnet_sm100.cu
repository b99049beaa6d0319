// helmnet_sm100.cu -- C ABI (include/helmnet_sm100.h) and host-side orchestration of the helmnet
// inference inner loop on one B200.
//
// One hn_ctx owns, resident in HBM for the whole solve: wavefield, residual, k_sq, source, hidden states
// (ping-pong), every UNet activation, operator tables, packed weights.  One solver iteration
// (IterativeSolver.single_step, helmnet/hybridnet.py:558-584) is a fixed sequence of kernels that is
// captured once per (batch, state parity) into a CUDA graph and replayed; nothing is allocated, copied
// to the host or synchronised inside hn_run.
#include "../../include/helmnet_sm100.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_simt.cuh"
#include "layout.cuh"
#include "spectral.cuh"
#include "spectral256.cuh"
#include "spectral512.cuh"
#include "spectral1024.cuh"
#include "train.cuh"
#ifndef HN_EMU
#include "conv_tc.cuh"
#include "conv_tcr.cuh"
#include "conv_tcr_down.cuh"
#include "conv_tcr_up.cuh"
#include "conv_tcf.cuh"
#define HN_HAVE_TC 1
#endif

using namespace hn;

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define HN_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(HN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                         std::to_string(__LINE__) + ")");                               \
    } while (0)
#define HN_TRY(expr)             \
    do {                         \
        int rc_ = (expr);        \
        if (rc_ != HN_OK) return rc_; \
    } while (0)

// Every ABI entry point runs with the context's device current and restores the caller's device on exit, so two
// contexts on different GPUs can be driven from one host thread (and PyTorch's current device is left alone).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
#ifndef HN_EMU
        if (device >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
#else
        (void)device;
#endif
    }
    ~DeviceGuard() {
#ifndef HN_EMU
        if (switched) cudaSetDevice(prev);
#endif
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct ConvW {      // offsets (in floats) into the packed device blob
    size_t w = 0, b = 0, slope = 0;
    size_t b8 = 0;            // bias zero-padded to 8 entries (tcgen05 kernels always read 8)
    size_t raw = 0;           // unpacked checkpoint layout [co][ci][3][3] (lean 2->2 state kernel)
    size_t tc = (size_t)-1;   // offset (in halfs) of the tcgen05 B-operand image, C_out = 8 layers only
    size_t tcr = (size_t)-1;  // same for the row-streaming kernel (N = 48 images)
    size_t tcf = (size_t)-1;  // fused DoubleConv kernel: the 2-channel group with its 3 dx taps folded into K (one N = 48 image)
    float tc_inv = 1.f;       // 2^-kw, inverse of the layer's weight block scale
    float l1 = 0.f, bmax = 0.f;   // max_co sum |W[co]| and max |b|: bound of the layer's output (fused DoubleConv mid scale)
};
struct Weights {
    ConvW inc[2], sig[kDepth][2], sta[kDepth][2], down[kDepth], bot[2], up[kDepth], dec[kDepth][2], outc;
};

struct TrainWs;
struct hn_ctx {
    int device = 0, n = 0, max_batch = 0, pml = 0;
    double sigma_max = 0, k0 = 1, omega = 1;
    int r[kDepth + 1] = {0};
    int state_len = 0;
    int engine = 2;            // 2: tcgen05 kernels with fused DoubleConvs (default), 1: tcgen05 one kernel per conv, 0: fp32 CUDA cores
    // resident fields
    float *wf = nullptr, *res = nullptr, *ksq = nullptr, *src = nullptr, *rx = nullptr;
    unsigned char* src_nz = nullptr;   // [max_batch][n]: column j of source map s holds a non-zero (spectral.cuh: tile_source)
    bool src_skip = true;              // HELMNET_SRC_SKIP=0: always load the source in the residual epilogue
    float* state[kDepth][2] = {{nullptr}};
    int cur = 0;
    int src_batch = 0, batch = 0;
    bool weights_set = false, problem_set = false;
    // UNet activations (NHWC8 unless noted)
    float* x[kDepth + 1] = {nullptr};
    float* mid[kDepth + 1] = {nullptr};
    float* mid2[kDepth] = {nullptr};   // float2
    float* skip[kDepth] = {nullptr};
    float* upo[kDepth] = {nullptr};
    float* dec[kDepth] = {nullptr};
    float* bot = nullptr;
    float* in6 = nullptr;   // NHWC8 staging for hn_unet
    float* dwf = nullptr;   // float2 raw network output for hn_unet
    float* tmp2 = nullptr;  // float2 scratch [B][n][n] (hn_residual / hn_laplacian inputs)
    float* tmp2b = nullptr;
    // tables
    SpecTables spec;
    float* sigma1d = nullptr;
    int rows_L = 1, cols_CW = 1;
    bool lean_state2 = true;   // dedicated 2->2 conv kernel for conv_state's second layer (both engines)
    bool spec_fast = true;     // use the register-resident N = 256 spectral kernels when they apply
    int spec_chunk = 0;        // samples per rows/cols kernel pair (0: sized to keep a chunk in L2)
    // weights
    float* wdev = nullptr;
    std::vector<float> whost;   // host copy of the packed fp32 weight blob (offsets of ConvW index both)
    float* wraw = nullptr;      // the 48,160 parameters as loaded (state_dict order): the backward pass reads the checkpoint layouts
    std::vector<float> wraw_host;
    TrainWs* tws = nullptr;     // workspace of hn_step_backward, allocated on first use
    uint16_t* tcw = nullptr;   // fp16 split-weight images for the tcgen05 convolutions
    int* err_flag = nullptr;   // device watchdog flag of the tcgen05 kernels
    unsigned* amax = nullptr;  // [64] running max |x| per activation tensor (publish_amax), feeds the fp16 block scales
    int tc_min_res = 16;       // use the tensor-core kernels for levels with resolution >= this
    int tcr_min_res = 48;      // row-streaming tensor-core kernel for widths >= this (128-wide strips)
    int num_sms = 148;
    // Programmatic dependent launch for the kernels of an iteration (common.cuh: HN_LAUNCH_PDL).  pdl_mode (HELMNET_PDL):
    // 0 off; 1 every kernel triggers its dependents early; 2 the persistent tcgen05 kernels that run several rounds of
    // strips with TWO CTAs per SM do not (pdl_early()); 3 no tcgen05 kernel triggers early.  Unset (pdl_cfg = -1): mode 2
    // for solves of at most kPdlAutoPoints points, off above -- measured on B200 (tools/gpu_ab_pdl.sh, r1): 256^2 x 1 -12 %,
    // 128^2 x 64 -5 %, 96^2 x 32 -3.5 % per iteration, but nothing (+-0.5 %) at 256^2 x 256 / 512^2 x 64 / 1024^2 x 8, where
    // every kernel runs for 100+ us and the iteration is power-capped.  pdl / pdl_mode are the values in effect (pdl_select).
    int pdl_cfg = -1;
    bool pdl = false;
    int pdl_mode = 0;
    bool tcf_any_width = true; // fused DoubleConv kernels for every even width up to 256 (not only 32 / 64 / 128 / 256)
    int tcf_min_width = 6;     // (6: the bottom DoubleConv of the 96^2 training-domain size)
    int dconv_min_rows = 4;    // shortest strip of the fused DoubleConv kernels (small batches: more, shorter strips fill more SMs)
    int pack_penalty = 12;     // per cent a pipeline step of those kernels gets slower per additional packed image (pick_pack)
    int pack_narrow = 1;       // levels at most 62 pixels wide: several images per M = 128 MMA (HELMNET_PACK_NARROW: 0 off, 1 the down- /
                               // up-sampling kernels (default), 2 also the fused DoubleConv kernels -- measured: -0.5 .. -1.2 % per iteration
                               // at 256^2 x 32 / x 64 / 128^2 x 64, +1 % at 256^2 x 256, so it stays opt-in)
    int bal_thresh = 100;      // balanced strips are taken when the model predicts at most this many per cent of the uniform cost (HELMNET_BAL_THRESH)
    int dconv_balance = 2;     // balanced strips where the model predicts a gain (HELMNET_DCONV_BALANCE: 0 off, 1 the fused DoubleConv
                               // kernels only, 2 also the down- / up-sampling kernels)
    bool fuse_bottom = true;   // decode[4] (the 8 -> 8 -> 8 DoubleConv at the bottom of the UNet) through the fused DoubleConv kernel
    // conv_state[d] (the hidden-state update) feeds nothing else in the same iteration: with side_state its kernels run on a
    // second stream / graph branch that forks after conv_signal[d] and joins at the end of the iteration, off the critical path
    // (HELMNET_SIDE_STATE: 0 off, 1 on, unset: on for solves of at most kSideAutoPoints points)
    int side_cfg = -1;
    bool side_state = false;
    bool side_pending = false;
    int tcd_min_res = 4;       // tensor-core down-/up-sampling for output (input) widths >= this (M = 64 MMAs up to 64 pixels): every
                               // level of the 256^2 and 96^2 pyramids (r1: 64 -- the 64- and 32-pixel levels of 256^2 ran on the CUDA
                               // cores, 0.18 ms; 96^2 x 32: 0.290 -> 0.255 ms with the 12- and 6-pixel levels on the tensor cores too)
    Weights W;
    // residual norms
    double* ssq = nullptr;
    size_t ssq_cap = 0;
    double* ssq1 = nullptr;   // [max_batch] for hn_residual
    int* iter_dev = nullptr;  // [0] = running slot, [1] = constant 0
    // launch accounting + graphs
    int64_t launches = 0;
    int kernels_per_iter = 0;
    bool use_graph = true;
#ifndef HN_EMU
    std::map<long long, cudaGraphExec_t> graphs;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork[kDepth] = {nullptr}, ev_join = nullptr;
#endif
    std::vector<void*> allocs;
};

// amax slot ids
enum : int { S_X = 0, S_MID = 5, S_SKIP = 10, S_UPO = 14, S_DEC = 18, S_BOT = 22, S_STATE = 23 /* + 2*d + buf */, S_IN6 = 31, S_DMID = 32, S_IMID = 36, S_WF = 37 /* +buf */, S_RES = 39 /* +buf */, S_COUNT = 41 };

#ifndef HN_EMU
static void drop_graphs(hn_ctx* c) {
    for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
    c->graphs.clear();
}
#endif

static int dalloc(hn_ctx* c, void** p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return fail(HN_ERR_NOMEM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    c->allocs.push_back(*p);
    return HN_OK;
}
template <typename T>
static int dalloc_t(hn_ctx* c, T** p, size_t count) {
    return dalloc(c, reinterpret_cast<void**>(p), count * sizeof(T));
}

static inline int grid1d(size_t total) {
    size_t g = (total + LAY_THREADS - 1) / LAY_THREADS;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (int)g;
}

// ------------------------------------------------------------------------------------------------
// Operator tables (helmnet/spectral.py:122-146, 267-363), computed in double precision on the host.
// ------------------------------------------------------------------------------------------------
static void factorize(int n, int* radix, int* nst) {
    int k = 0;
    while (n % 16 == 0) { radix[k++] = 16; n /= 16; }
    while (n % 4 == 0) { radix[k++] = 4; n /= 4; }
    while (n % 2 == 0) { radix[k++] = 2; n /= 2; }
    for (int p = 3; n > 1; p += 2)
        while (n % p == 0) { radix[k++] = p; n /= p; }
    *nst = k;
}

static int build_tables(hn_ctx* c) {
    const int n = c->n, pml = c->pml;
    std::vector<float> tw(2 * n), mk(n), msq(n), a(2 * n, 0.f), b(2 * n), sig(n, 0.f);
    const double PI = 3.14159265358979323846;
    for (int k = 0; k < n; k++) {
        const double ang = -2.0 * PI * (double)k / (double)n;
        tw[2 * k] = (float)cos(ang);
        tw[2 * k + 1] = (float)sin(ang);
    }
    // spectral.py:126-127: k = 2 pi linspace(-0.5, 0.5, n, endpoint=False), rotated by n/2, cast to float32 (:141)
    for (int i = 0; i < n; i++) {
        const int j = (i + n / 2) % n;                       // position in the un-rotated linspace
        const double kd = 2.0 * PI * (-0.5 + (double)j * (1.0 / (double)n));
        const float k32 = (float)kd;
        const float ksq32 = k32 * k32;                       // spectral.py:281 float32 pow(2)
        mk[i] = (float)((double)k32 / (double)n);
        msq[i] = (float)(-(double)ksq32 / (double)n);
    }
    // spectral.py:306-338
    std::vector<double> sigma(n, 0.0), sprime(n, 0.0);
    for (int i = 0; i < pml; i++) {
        const double so = c->sigma_max * pow(fabs(1.0 - (double)i / (double)pml), 2.0);
        const double sp = -2.0 * c->sigma_max * (1.0 - (double)i / (double)pml) / (double)pml;
        sigma[i] = so;
        sigma[n - 1 - i] = so;
        sprime[i] = sp;
        sprime[n - 1 - i] = -sp;
    }
    for (int i = 0; i < n; i++) {
        // inv_gamma = 1 / (1 + (i/k0) sigma);  gamma' = (i/k0) sigma'
        const double s = sigma[i] / c->k0, d = 1.0 + s * s;
        const double gr = 1.0 / d, gi = -s / d;              // inv_gamma
        const double g2r = gr * gr - gi * gi, g2i = 2.0 * gr * gi;          // inv_gamma^2  = b
        const double g3r = g2r * gr - g2i * gi, g3i = g2r * gi + g2i * gr;  // inv_gamma^3
        const double gpi = sprime[i] / c->k0;                // gamma' = i * gpi
        // a = -gamma' * inv_gamma^3 = -(i gpi)(g3r + i g3i) = gpi*g3i - i gpi*g3r
        a[2 * i] = (float)(gpi * g3i);
        a[2 * i + 1] = (float)(-gpi * g3r);
        b[2 * i] = (float)g2r;
        b[2 * i + 1] = (float)g2i;
        sig[i] = (float)sigma[i];
    }
    float *d_tw, *d_mk, *d_msq, *d_a, *d_b;
    HN_TRY(dalloc_t(c, &d_tw, 2 * n));
    HN_TRY(dalloc_t(c, &d_mk, n));
    HN_TRY(dalloc_t(c, &d_msq, n));
    HN_TRY(dalloc_t(c, &d_a, 2 * n));
    HN_TRY(dalloc_t(c, &d_b, 2 * n));
    HN_TRY(dalloc_t(c, &c->sigma1d, n));
    HN_CUDA(cudaMemcpy(d_tw, tw.data(), 2 * n * 4, cudaMemcpyHostToDevice));
    HN_CUDA(cudaMemcpy(d_mk, mk.data(), n * 4, cudaMemcpyHostToDevice));
    HN_CUDA(cudaMemcpy(d_msq, msq.data(), n * 4, cudaMemcpyHostToDevice));
    HN_CUDA(cudaMemcpy(d_a, a.data(), 2 * n * 4, cudaMemcpyHostToDevice));
    HN_CUDA(cudaMemcpy(d_b, b.data(), 2 * n * 4, cudaMemcpyHostToDevice));
    HN_CUDA(cudaMemcpy(c->sigma1d, sig.data(), n * 4, cudaMemcpyHostToDevice));
    c->spec.tw = reinterpret_cast<const float2*>(d_tw);
    c->spec.mk = d_mk;
    c->spec.msq = d_msq;
    c->spec.a = reinterpret_cast<const float2*>(d_a);
    c->spec.b = reinterpret_cast<const float2*>(d_b);
    c->spec.n = n;
    c->spec.pml = pml;
    factorize(n, c->spec.radix, &c->spec.nstages);
    // lines per CTA: ~48 KB of line buffers for the row pass; the column pass needs >= 32 B segments
    int L = 2048 / n;     // ~50 KB of line buffers: four CTAs per SM
    // small solves: half the lines per CTA, twice the CTAs (96^2 x 32: 192 CTAs with 16 lines; residual stage 0.052 -> 0.048 ms)
    if ((long long)c->max_batch * n / 16 < 4ll * c->num_sms && L > 8) L = 8;
    if (const char* ev = getenv("HELMNET_SPEC_L")) L = atoi(ev);
    if (L < 1) L = 1;
    if (L > 16) L = 16;
    c->rows_L = L;
    int CW = ((long long)c->max_batch * n / 16 < 4ll * c->num_sms) ? 8 : 16;
    if (const char* ev = getenv("HELMNET_SPEC_CW")) CW = atoi(ev);
    if (const char* ev = getenv("HELMNET_SPEC_CHUNK")) c->spec_chunk = atoi(ev);
    if (const char* ev = getenv("HELMNET_SPEC_FAST")) c->spec_fast = atoi(ev) != 0;
    if (const char* ev = getenv("HELMNET_LEAN_STATE2")) c->lean_state2 = atoi(ev) != 0;
    // ~56 KB per CTA (four CTAs per SM hide the tile-load latency; measured best at 256: L = 8, CW = 8), but keep
    // column segments >= 32 B unless the line buffers would not fit at all
    while (CW > 4 && spectral_smem_bytes(n, CW, pml) > 56 * 1024) CW >>= 1;
    while (CW > 1 && spectral_smem_bytes(n, CW, pml) > 200 * 1024) CW >>= 1;
    c->cols_CW = CW;
    if (spectral_smem_bytes(n, CW, pml) > 227 * 1024) return fail(HN_ERR_ARG, "domain size too large for one line per CTA");
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// weights: state_dict order of HybridNet (helmnet/architectures.py:317-388) -> packed device layouts
// ------------------------------------------------------------------------------------------------
struct Packer {
    std::vector<float> blob;
    std::vector<uint16_t> halfs;   // tcgen05 B images
    size_t reserve(size_t nfl) {
        size_t off = (blob.size() + 3) & ~(size_t)3;
        blob.resize(off + nfl, 0.f);
        return off;
    }
};
// W[co][ci][3][3] -> P[pl][tap][ci4][co]
static size_t pack_conv3(Packer& pk, const float* w, int cout, int cin) {
    const int npl = (cin + 3) / 4;
    size_t off = pk.reserve((size_t)npl * 9 * 4 * cout);
    for (int pl = 0; pl < npl; pl++)
        for (int tap = 0; tap < 9; tap++)
            for (int c4 = 0; c4 < 4; c4++)
                for (int co = 0; co < cout; co++) {
                    const int ci = pl * 4 + c4;
                    pk.blob[off + ((pl * 9 + tap) * 4 + c4) * cout + co] = ci < cin ? w[(co * cin + ci) * 9 + tap] : 0.f;
                }
    return off;
}
// W[co][ci][8][8] -> P[pl][ky][kx][ci4][co]
static size_t pack_down(Packer& pk, const float* w) {
    size_t off = pk.reserve(2 * 64 * 4 * 8);
    for (int pl = 0; pl < 2; pl++)
        for (int k = 0; k < 64; k++)
            for (int c4 = 0; c4 < 4; c4++)
                for (int co = 0; co < 8; co++) pk.blob[off + ((pl * 64 + k) * 4 + c4) * 8 + co] = w[(co * 8 + pl * 4 + c4) * 64 + k];
    return off;
}
// ConvTranspose weight W[ci][co][8][8] -> P[cls][pl][ty][tx][ci4][co],  ky = 7 - 2ty - py, kx = 7 - 2tx - px
static size_t pack_up(Packer& pk, const float* w) {
    size_t off = pk.reserve(4 * 2 * 16 * 4 * 8);
    for (int cls = 0; cls < 4; cls++) {
        const int py = cls >> 1, px = cls & 1;
        for (int pl = 0; pl < 2; pl++)
            for (int ty = 0; ty < 4; ty++)
                for (int tx = 0; tx < 4; tx++)
                    for (int c4 = 0; c4 < 4; c4++)
                        for (int co = 0; co < 8; co++) {
                            const int ky = 7 - 2 * ty - py, kx = 7 - 2 * tx - px, ci = pl * 4 + c4;
                            pk.blob[off + ((((cls * 2 + pl) * 4 + ty) * 4 + tx) * 4 + c4) * 8 + co] = w[((ci * 8 + co) * 8 + ky) * 8 + kx];
                        }
    }
    return off;
}
static size_t pack_vec(Packer& pk, const float* v, int nfl) {
    size_t off = pk.reserve(nfl);
    for (int i = 0; i < nfl; i++) pk.blob[off + i] = v[i];
    return off;
}

#ifdef HN_HAVE_TC
// W[co][ci][3][3] (co = 8) -> per (channel group g, tap) a 16 x 16 fp16 B operand in the canonical K-major
// SWIZZLE_NONE layout:  byte(n,k) = (n/8)*256 + (k/8)*128 + (n%8)*16 + (k%8)*2  with
//   n <  8 (g1 = hi*W_hi):            k < 8: W_hi[co=n][ci=k]      k >= 8: 0
//   n >= 8 (g2 = hi*W_lo + lo*W_hi):  k < 8: W_lo[co=n-8][ci=k]    k >= 8: W_hi[co=n-8][ci=k-8]
// W' = W * 2^kw with max|W'| in [2^9, 2^10);  W_hi = fp16(W'),  W_lo = fp16((W' - W_hi) * 2^11).
static void pack_tc(Packer& pk, const float* w, int cin, ConvW& out) {
    const int G = (cin + 7) / 8;
    float mx = 0.f;
    for (int i = 0; i < 8 * cin * 9; i++) mx = fmaxf(mx, fabsf(w[i]));
    int ex = 0;
    if (mx > 0.f) frexpf(mx, &ex);
    const int kw = 10 - ex;
    const float scale = ldexpf(1.f, kw);
    out.tc_inv = ldexpf(1.f, -kw);
    out.tc = pk.halfs.size();
    pk.halfs.resize(out.tc + (size_t)G * 9 * 256, 0);
    for (int g = 0; g < G; g++)
        for (int tap = 0; tap < 9; tap++)
            for (int n = 0; n < 16; n++)
                for (int k = 0; k < 16; k++) {
                    const int co = n & 7, ci = g * 8 + (k & 7);
                    const float wv = ci < cin ? w[(co * cin + ci) * 9 + tap] * scale : 0.f;
                    const __half hi = __float2half_rn(wv);
                    const __half lo = __float2half_rn((wv - __half2float(hi)) * 2048.f);
                    __half val = __float2half_rn(0.f);
                    if (n < 8) { if (k < 8) val = hi; }
                    else val = (k < 8) ? lo : hi;
                    const int byte = (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
                    uint16_t bits;
                    memcpy(&bits, &val, 2);
                    pk.halfs[out.tc + (size_t)(g * 9 + tap) * 256 + byte / 2] = bits;
                }
}
// Row-streaming kernel (conv_tcr.cuh): per (group g, dx) a 48 x 16 fp16 B operand, column n = dy*16 + h*8 + co
// (h = 0: hi*W_hi part, h = 1: hi*W_lo + lo*W_hi part), same canonical layout, 1536 B each.
static void pack_tcr(Packer& pk, const float* w, int cin, ConvW& out, int cout = 8) {
    const int G = (cin + 7) / 8;
    float mx = 0.f;
    for (int i = 0; i < cout * cin * 9; i++) mx = fmaxf(mx, fabsf(w[i]));
    int ex = 0;
    if (mx > 0.f) frexpf(mx, &ex);
    const int kw = 10 - ex;
    const float scale = ldexpf(1.f, kw);
    out.tc_inv = ldexpf(1.f, -kw);
    out.tcr = pk.halfs.size();
    pk.halfs.resize(out.tcr + (size_t)G * 3 * 768, 0);
    for (int g = 0; g < G; g++)
        for (int dx = 0; dx < 3; dx++)
            for (int n = 0; n < 48; n++)
                for (int k = 0; k < 16; k++) {
                    const int dy = n / 16, h = (n / 8) & 1, co = n & 7, ci = g * 8 + (k & 7);
                    const float wv = (ci < cin && co < cout) ? w[(co * cin + ci) * 9 + dy * 3 + dx] * scale : 0.f;
                    const __half hi = __float2half_rn(wv);
                    const __half lo = __float2half_rn((wv - __half2float(hi)) * 2048.f);
                    __half val = __float2half_rn(0.f);
                    if (h == 0) { if (k < 8) val = hi; }
                    else val = (k < 8) ? lo : hi;
                    const int byte = (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
                    uint16_t bits;
                    memcpy(&bits, &val, 2);
                    pk.halfs[out.tcr + (size_t)(g * 3 + dx) * 768 + byte / 2] = bits;
                }
}
// Fused DoubleConv kernel (conv_tcf.cuh), 2-channel group (input channels ci0, ci0 + 1) with the three horizontal taps
// folded into K:  ONE 48 x 16 image, column n = dy*16 + h*8 + co as in pack_tcr, row k = dx*4 + part*2 + c
// (part 0: the activation's hi half, part 1: its lo half; k = 12..15 zero).  Same block scale as pack_tcr of the layer.
static void pack_tcf_fold(Packer& pk, const float* w, int cin, int ci0, ConvW& out, int cout) {
    float mx = 0.f;
    for (int i = 0; i < cout * cin * 9; i++) mx = fmaxf(mx, fabsf(w[i]));
    int ex = 0;
    if (mx > 0.f) frexpf(mx, &ex);
    const int kw = 10 - ex;
    const float scale = ldexpf(1.f, kw);
    out.tcf = pk.halfs.size();
    pk.halfs.resize(out.tcf + 768, 0);
    for (int n = 0; n < 48; n++)
        for (int k = 0; k < 12; k++) {
            const int dy = n / 16, h = (n / 8) & 1, co = n & 7, dx = k / 4, part = (k / 2) & 1, ci = ci0 + (k & 1);
            const float wv = co < cout ? w[(co * cin + ci) * 9 + dy * 3 + dx] * scale : 0.f;
            const __half hi = __float2half_rn(wv);
            const __half lo = __float2half_rn((wv - __half2float(hi)) * 2048.f);
            __half val = __float2half_rn(0.f);
            if (h == 0) { if (part == 0) val = hi; }
            else val = (part == 0) ? lo : hi;
            const int byte = (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
            uint16_t bits;
            memcpy(&bits, &val, 2);
            pk.halfs[out.tcf + byte / 2] = bits;
        }
}
// Down-sampling conv on tensor cores (conv_tcr_down.cuh): W[co][ci][8][8] -> per (t = ky parity, kx) a 64 x 16 fp16
// B operand, column n = m*16 + h*8 + co with ky = 2m + t.
static void pack_tcd(Packer& pk, const float* w, ConvW& out) {
    float mx = 0.f;
    for (int i = 0; i < 8 * 8 * 64; i++) mx = fmaxf(mx, fabsf(w[i]));
    int ex = 0;
    if (mx > 0.f) frexpf(mx, &ex);
    const int kw = 10 - ex;
    const float scale = ldexpf(1.f, kw);
    out.tc_inv = ldexpf(1.f, -kw);
    out.tcr = pk.halfs.size();
    pk.halfs.resize(out.tcr + (size_t)16 * 1024, 0);
    for (int t = 0; t < 2; t++)
        for (int kx = 0; kx < 8; kx++)
            for (int n = 0; n < 64; n++)
                for (int k = 0; k < 16; k++) {
                    const int m = n / 16, h = (n / 8) & 1, co = n & 7, ci = k & 7, ky = 2 * m + t;
                    const float wv = w[((co * 8 + ci) * 8 + ky) * 8 + kx] * scale;
                    const __half hi = __float2half_rn(wv);
                    const __half lo = __float2half_rn((wv - __half2float(hi)) * 2048.f);
                    __half val = __float2half_rn(0.f);
                    if (h == 0) { if (k < 8) val = hi; }
                    else val = (k < 8) ? lo : hi;
                    const int byte = (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
                    uint16_t bits;
                    memcpy(&bits, &val, 2);
                    pk.halfs[out.tcr + (size_t)(t * 8 + kx) * 1024 + byte / 2] = bits;
                }
}
// Up-sampling transposed conv on tensor cores (conv_tcr_up.cuh): W[ci][co][8][8] -> per shift s = si - 2 a 256 x 16
// fp16 B operand, column n = ky*32 + px*16 + h*8 + co with kx = px + 3 - 2 s (zero block when kx is out of range).
static void pack_tcu(Packer& pk, const float* w, ConvW& out) {
    float mx = 0.f;
    for (int i = 0; i < 8 * 8 * 64; i++) mx = fmaxf(mx, fabsf(w[i]));
    int ex = 0;
    if (mx > 0.f) frexpf(mx, &ex);
    const int kw = 10 - ex;
    const float scale = ldexpf(1.f, kw);
    out.tc_inv = ldexpf(1.f, -kw);
    out.tcr = pk.halfs.size();
    pk.halfs.resize(out.tcr + (size_t)5 * 4096, 0);
    for (int si = 0; si < 5; si++)
        for (int n = 0; n < 256; n++)
            for (int k = 0; k < 16; k++) {
                const int ky = n / 32, px = (n / 16) & 1, h = (n / 8) & 1, co = n & 7, ci = k & 7;
                const int kx = px + 3 - 2 * (si - 2);
                const float wv = (kx >= 0 && kx < 8) ? w[((ci * 8 + co) * 8 + ky) * 8 + kx] * scale : 0.f;
                const __half hi = __float2half_rn(wv);
                const __half lo = __float2half_rn((wv - __half2float(hi)) * 2048.f);
                __half val = __float2half_rn(0.f);
                if (h == 0) { if (k < 8) val = hi; }
                else val = (k < 8) ? lo : hi;
                const int byte = (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
                uint16_t bits;
                memcpy(&bits, &val, 2);
                pk.halfs[out.tcr + (size_t)si * 4096 + byte / 2] = bits;
            }
}
#endif

struct Cursor {
    const float* p;
    size_t left;
    bool ok = true;
    const float* take(size_t nfl) {
        if (nfl > left) { ok = false; return p; }
        const float* q = p;
        p += nfl;
        left -= nfl;
        return q;
    }
};
static void pack_double_conv(Packer& pk, Cursor& cur, ConvW out[2], int cin, int cmid, int cout) {
    const float* w0 = cur.take((size_t)cmid * cin * 9);
    const float* b0 = cur.take(cmid);
    const float* sl = cur.take(1);
    const float* w1 = cur.take((size_t)cout * cmid * 9);
    const float* b1 = cur.take(cout);
    if (!cur.ok) return;
    out[0].w = pack_conv3(pk, w0, cmid, cin);
    out[0].b = pack_vec(pk, b0, cmid);
    {
        float b8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < cmid && i < 8; i++) b8[i] = b0[i];
        out[0].b8 = pack_vec(pk, b8, 8);
    }
    out[0].slope = pack_vec(pk, sl, 1);
    out[1].w = pack_conv3(pk, w1, cout, cmid);
    out[1].b = pack_vec(pk, b1, cout);
    {
        float b8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < cout && i < 8; i++) b8[i] = b1[i];
        out[1].b8 = pack_vec(pk, b8, 8);
    }
    for (int co = 0; co < cmid; co++) {
        float l1 = 0.f;
        for (int i = 0; i < cin * 9; i++) l1 += fabsf(w0[(size_t)co * cin * 9 + i]);
        out[0].l1 = fmaxf(out[0].l1, l1);
        out[0].bmax = fmaxf(out[0].bmax, fabsf(b0[co]));
    }
    out[1].raw = pack_vec(pk, w1, cout * cmid * 9);
    out[1].slope = out[0].slope;
#ifdef HN_HAVE_TC
    if (cmid == 8) { pack_tc(pk, w0, cin, out[0]); pack_tcr(pk, w0, cin, out[0]); }
    if (cmid == 2 && cin == 10) pack_tcr(pk, w0, cin, out[0], 2);   // conv_state.0: C_out = 2 padded to 8 accumulator columns
    if (cout == 8 && cmid == 8) { pack_tc(pk, w1, cmid, out[1]); pack_tcr(pk, w1, cmid, out[1]); }
    if (cout == 2 && cmid == 2) pack_tcr(pk, w1, cmid, out[1], 2);  // conv_state.2 for the fused DoubleConv kernel
    if (cin == 10) pack_tcf_fold(pk, w0, cin, 8, out[0], cmid);     // hidden-state channels of cat[x, state]
    if (cout == 2 && cmid == 2) pack_tcf_fold(pk, w1, cmid, 0, out[1], 2);
#endif
}

// ------------------------------------------------------------------------------------------------
// kernel launch helpers
// ------------------------------------------------------------------------------------------------
constexpr long long kPdlAutoPoints = 8ll << 20;      // r2: 256^2 x 128 1.342 -> 1.306 ms with PDL mode 2 + side branch, nothing at x 256
constexpr long long kSideAutoPoints = 8ll << 20;   // side branch for conv_state: on for small solves (measured)
static inline void pdl_select(hn_ctx* c, int B) {
    c->pdl_mode = c->pdl_cfg >= 0 ? c->pdl_cfg : ((long long)B * c->n * c->n <= kPdlAutoPoints ? 2 : 0);
    c->pdl = c->pdl_mode != 0;
#ifndef HN_EMU
    c->side_state = c->side_cfg >= 0 ? c->side_cfg != 0 : ((long long)B * c->n * c->n <= kSideAutoPoints);
#endif
}
// fork / join of the conv_state side branch (no-ops when it is off)
static int side_fork(hn_ctx* c, int d, cudaStream_t st, cudaStream_t* out) {
    *out = st;
#ifndef HN_EMU
    if (c->side_state) {
        HN_CUDA(cudaEventRecord(c->ev_fork[d], st));
        HN_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork[d], 0));
        *out = c->side;
        c->side_pending = true;
    }
#endif
    return HN_OK;
}
static int side_join(hn_ctx* c, cudaStream_t st) {
#ifndef HN_EMU
    if (c->side_pending) {
        HN_CUDA(cudaEventRecord(c->ev_join, c->side));
        HN_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
        c->side_pending = false;
    }
#endif
    return HN_OK;
}
// Should a persistent tcgen05 kernel let its dependents in early (pdl_trigger)?  Mode 2: yes when it runs one CTA per SM or
// when every CTA has at most one strip (a single round); a kernel that walks several rounds with two CTAs per SM keeps the
// default CTA placement, which pairs a long and a short strip list on every SM (measured: down[0] 177 -> 186 us otherwise).
static inline int pdl_early(const hn_ctx* c, int total_strips, int ctas_per_sm) {
    if (c->pdl_mode == 1) return 1;
    if (c->pdl_mode == 2) return (ctas_per_sm == 1 || total_strips <= ctas_per_sm * c->num_sms) ? 1 : 0;
    return 0;
}

#ifdef HN_HAVE_TC
// Output rows per strip of the per-conv row-streaming kernel (conv_tcr.cuh): 32 as in r1 while an SM has 64+ rows to do
// (throughput-bound), otherwise whole rounds of equal strips over two CTA slots per SM; a strip of R rows streams R + 2 rows
// (+ ~2 row steps of pipeline fill).  1024^2 x 1: the 512-pixel level ran on 64 CTAs with the fixed height.
static int tcr_rows_per_strip(const hn_ctx* c, int H, int nsx, int B) {
    if ((long long)B * H * nsx / c->num_sms >= 64) return H < tcr::ROWS ? H : tcr::ROWS;
    int best = tcr::ROWS;
    long long best_cost = -1;
    for (int rows = 2; rows <= 128 && rows <= H; rows += 2) {
        const long long total = (long long)nsx * ((H + rows - 1) / rows) * B;
        const long long g = total < 2 * c->num_sms ? total : 2 * c->num_sms;
        const long long cost = ((total + g - 1) / g) * (rows + 6);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rows; }
    }
    return best;
}
#endif

#ifdef HN_HAVE_TC
// strips of the per-conv row-streaming kernel: uniform rounds (tcr_rows_per_strip) or, where the model predicts 3 %, balanced chunks
// over the rows of all batch * nsx column strips (common.cuh: balanced_strip).  Returns the grid.
static int tcr_plan_strips(const hn_ctx* c, tcr::Args& t, int B) {
    t.nsx = (t.W + tcr::CW - 1) / tcr::CW;
    t.rows = tcr_rows_per_strip(c, t.H, t.nsx, B);
    t.nsy = (t.H + t.rows - 1) / t.rows;
    t.total_strips = t.nsx * t.nsy * B;
    const int cap = 2 * c->num_sms;       // persistent: 2 CTAs per SM
    int tgrid = t.total_strips < cap ? t.total_strips : cap;
    t.pdl_trig = pdl_early(c, t.total_strips, 2);
    t.bal = 0;
    if (c->dconv_balance >= 2) {
        const long long rounds = (t.total_strips + tgrid - 1) / tgrid;
        // (32-row strips at 64+ rows per SM pair a long and a short strip list per SM: what counts there is the work per SM)
        const bool per_sm = (long long)B * t.H * t.nsx / c->num_sms >= 64;
        const long long uniform_sm = per_sm ? ((long long)t.total_strips + c->num_sms - 1) / c->num_sms * (t.rows + 6) : 2 * rounds * (t.rows + 6);
        const long long vt = (long long)B * t.nsx * (t.H + tcr::BAL_PAD);
        const long long balanced = (vt + cap - 1) / cap + tcr::BAL_PAD + 2;
        if (vt / cap >= 12 && 2 * balanced * 100 <= uniform_sm * c->bal_thresh) {
            t.bal = B * t.nsx;
            tgrid = cap;
            t.pdl_trig = c->pdl_mode == 1 || c->pdl_mode == 2;
        }
    }
    return tgrid;
}
#endif

template <int SRC, int COUT, bool PRELU, int EPI>
static int launch_conv3(hn_ctx* c, const Conv3Args& a, int B, cudaStream_t st) {
    const size_t smem = conv3_smem_bytes(SRC, COUT);
#ifndef HN_EMU
    static bool attr_done[16] = {false};
    if (!attr_done[c->device & 15]) {
        HN_CUDA(cudaFuncSetAttribute(conv3x3_kernel<SRC, COUT, PRELU, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[c->device & 15] = true;
    }
#endif
#ifdef HN_HAVE_TC
    if constexpr (COUT == 2 && SRC == SRC_A8_B2 && EPI == EPI_STORE) {
        if (c->engine >= 1 && a.tcr_bmat != nullptr && a.W >= c->tcr_min_res && (a.H % 2) == 0 && a.amax_in0 != nullptr) {
            static bool tcr2_attr_done[16] = {false};
            if (!tcr2_attr_done[c->device & 15]) {
                HN_CUDA(cudaFuncSetAttribute(tcr::conv3x3_tcr_kernel<SRC, PRELU, EPI_STORE2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)tcr::smem_bytes(SRC)));
                tcr2_attr_done[c->device & 15] = true;
            }
            tcr::Args t;
            t.inA = a.inA; t.inB = a.inB; t.sigma = a.sigma;
            t.bmat = reinterpret_cast<const __half*>(a.tcr_bmat);
            t.bias = a.bias8; t.slope = a.slope; t.out = a.out; t.wo = nullptr; t.bo = nullptr; t.wf = nullptr; t.dwf_out = nullptr;
            t.amax_in0 = a.amax_in0; t.amax_in1 = a.amax_in1; t.amax_out = a.amax_out;
            t.error_flag = c->err_flag; t.sigma_max = 0.f; t.w_inv_scale = a.tc_inv;
            t.H = a.H; t.W = a.W;
            const int tgrid = tcr_plan_strips(c, t, B);
            HN_LAUNCH_PDL(c->pdl, (tcr::conv3x3_tcr_kernel<SRC, PRELU, EPI_STORE2>), dim3(tgrid), dim3(tcr::THREADS), tcr::smem_bytes(SRC), st, t);
            c->launches++;
            return HN_OK;
        }
    }
    if constexpr (COUT == 8) {
        if (c->engine >= 1 && a.tcr_bmat != nullptr && a.W >= c->tcr_min_res && (a.H % 2) == 0 && a.amax_in0 != nullptr) {
            static bool tcr_attr_done[16] = {false};
            if (!tcr_attr_done[c->device & 15]) {
                HN_CUDA(cudaFuncSetAttribute(tcr::conv3x3_tcr_kernel<SRC, PRELU, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)tcr::smem_bytes(SRC)));
                tcr_attr_done[c->device & 15] = true;
            }
            tcr::Args t;
            t.inA = a.inA; t.inB = a.inB; t.sigma = a.sigma;
            t.bmat = reinterpret_cast<const __half*>(a.tcr_bmat);
            t.bias = a.bias; t.slope = a.slope; t.out = a.out; t.wo = a.wo; t.bo = a.bo; t.wf = a.wf; t.dwf_out = a.dwf_out;
            t.amax_in0 = a.amax_in0; t.amax_in1 = a.amax_in1; t.amax_out = a.amax_out;
            t.error_flag = c->err_flag; t.sigma_max = c->pml > 0 ? (float)c->sigma_max : 0.f; t.w_inv_scale = a.tc_inv;
            t.H = a.H; t.W = a.W;
            const int tgrid = tcr_plan_strips(c, t, B);
            HN_LAUNCH_PDL(c->pdl, (tcr::conv3x3_tcr_kernel<SRC, PRELU, EPI>), dim3(tgrid), dim3(tcr::THREADS), tcr::smem_bytes(SRC), st, t);
            c->launches++;
            return HN_OK;
        }
        if (c->engine >= 1 && a.tc_bmat != nullptr && a.H >= c->tc_min_res && (SRC == SRC_INC || a.amax_in0 != nullptr)) {
            static bool tc_attr_done[16] = {false};
            if (!tc_attr_done[c->device & 15]) {
                HN_CUDA(cudaFuncSetAttribute(tc::conv3x3_tc_kernel<SRC, PRELU, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)tc::smem_bytes(SRC)));
                tc_attr_done[c->device & 15] = true;
            }
            tc::Args t;
            t.inA = a.inA; t.inB = a.inB; t.sigma = a.sigma;
            t.bmat = reinterpret_cast<const __half*>(a.tc_bmat);
            t.bias = a.bias; t.slope = a.slope; t.out = a.out; t.wo = a.wo; t.bo = a.bo; t.wf = a.wf; t.dwf_out = a.dwf_out;
            t.amax_in0 = a.amax_in0; t.amax_in1 = a.amax_in1; t.amax_out = a.amax_out;
            t.error_flag = c->err_flag; t.w_inv_scale = a.tc_inv; t.H = a.H; t.W = a.W;
            dim3 tgrid((a.W + tc::TX - 1) / tc::TX, (a.H + tc::TY - 1) / tc::TY, B);
            HN_LAUNCH_PDL(c->pdl, (tc::conv3x3_tc_kernel<SRC, PRELU, EPI>), tgrid, dim3(tc::THREADS), tc::smem_bytes(SRC), st, t);
            c->launches++;
            return HN_OK;
        }
    }
#endif
    dim3 grid((a.W + C3_TX - 1) / C3_TX, (a.H + C3_TY - 1) / C3_TY, B);
    HN_LAUNCH_PDL(c->pdl, (conv3x3_kernel<SRC, COUT, PRELU, EPI>), grid, dim3(C3_THREADS), smem, st, a);
    c->launches++;
    return HN_OK;
}

static Conv3Args conv_args(hn_ctx* c, const ConvW& w, const float* inA, const float* inB, float* out, int r, int slot_out = -1,
                           int slot_in0 = -1, int slot_in1 = -1) {
    Conv3Args a;
    memset(&a, 0, sizeof(a));
    a.amax_out = slot_out >= 0 ? c->amax + slot_out : nullptr;
    a.amax_in0 = slot_in0 >= 0 ? c->amax + slot_in0 : nullptr;
    a.amax_in1 = slot_in1 >= 0 ? c->amax + slot_in1 : (slot_in0 >= 0 ? c->amax + slot_in0 : nullptr);
    a.inA = inA;
    a.inB = inB;
    a.sigma = c->sigma1d;
    a.w = c->wdev + w.w;
    a.bias = c->wdev + w.b;
    a.bias8 = c->wdev + w.b8;
    a.slope = c->wdev + w.slope;
    a.out = out;
    a.H = r;
    a.W = r;
    a.tc_bmat = (w.tc != (size_t)-1 && c->tcw) ? (const void*)(c->tcw + w.tc) : nullptr;
    a.tcr_bmat = (w.tcr != (size_t)-1 && c->tcw) ? (const void*)(c->tcw + w.tcr) : nullptr;
    a.tc_inv = w.tc_inv;
    return a;
}

#ifndef HN_EMU
static int set_smem_attrs(hn_ctx* c) {
    HN_CUDA(cudaFuncSetAttribute(down_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DN_SMEM));
    HN_CUDA(cudaFuncSetAttribute(up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UP_SMEM));
    // the attribute is per function and device, not per context: keep the running maximum over all contexts, or a second
    // context with a smaller domain would lower the limit under the first one's launches
    static int rows_max[16] = {0}, cols_max[16] = {0};
    int& rmax = rows_max[c->device & 15];
    int& cmax = cols_max[c->device & 15];
    const int rneed = (int)spectral_smem_bytes(c->n, c->rows_L, c->pml), cneed = (int)spectral_smem_bytes(c->n, c->cols_CW, c->pml);
    if (rneed > rmax) {
        HN_CUDA(cudaFuncSetAttribute(spectral_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rneed));
        rmax = rneed;
    }
    if (cneed > cmax) {
        HN_CUDA(cudaFuncSetAttribute(spectral_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cneed));
        cmax = cneed;
    }
    HN_CUDA(cudaFuncSetAttribute(s256::spectral_cols256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s256::COLS_SMEM_BYTES));
    HN_CUDA(cudaFuncSetAttribute(s512::spectral_rows512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s512::ROWS_SMEM_BYTES));
    HN_CUDA(cudaFuncSetAttribute(s512::spectral_cols512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s512::COLS_SMEM_BYTES));
    HN_CUDA(cudaFuncSetAttribute(s1024::spectral_rows1024_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1024::ROWS_SMEM_BYTES));
    HN_CUDA(cudaFuncSetAttribute(s1024::spectral_cols1024_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1024::COLS_SMEM_BYTES));
    return HN_OK;
}
#endif

#ifdef HN_HAVE_TC
// Images per M = 128 MMA of a narrow down- / up-sampling level (conv_tcr_down.cuh: Args::pack), 0 = one image per MMA as before.
// These launches are bound by the latency of a pipeline step, not by its work: packing g images cuts the steps per CTA only while
// there are more rows than CTA slots, and every packed image makes a step a little heavier (2 more TMA loads, M = 128 instead of
// 64).  Model: (rows per CTA slot + pad) x (1 + penalty (g - 1)) with whole even strips of at least two rows; measured on B200
// (profiles/r2_ab_runs.txt): full packing gains 4-5 % per iteration at 256^2 x 256 and loses 7 % at 64^2 x 32, where a level has
// fewer rows than slots to begin with.
static int pick_pack(const hn_ctx* c, int B, int width, int rows, int cap, int pad) {
    if (!c->pack_narrow || width > 62 || B < 2) return 0;
    const int s_ = width + 2, gmax = (128 - width) / s_ + 1;
    int best = 1;
    long long best_cost = -1;
    for (int g = 1; g <= gmax && g <= B; g++) {
        const long long groups = (B + g - 1) / g;
        long long R = (groups * rows + cap - 1) / cap;
        R = R < 2 ? 2 : ((R + 1) & ~1ll);
        const long long cost = (R + pad) * (100 + (long long)c->pack_penalty * (g - 1));
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = g; }
    }
    return best > 1 ? best : 0;
}
#endif

static int launch_down(hn_ctx* c, int d, int B, cudaStream_t st) {
    const Weights& W = c->W;
    const int r = c->r[d];
    DownArgs dn;
    dn.in = c->skip[d];
    dn.w = c->wdev + W.down[d].w;
    dn.bias = c->wdev + W.down[d].b;
    dn.out = c->x[d + 1];
    dn.amax_out = c->amax + S_X + d + 1;
    dn.H = r;
    dn.W = r;
#ifdef HN_HAVE_TC
    if (c->engine >= 1 && W.down[d].tcr != (size_t)-1 && r / 2 >= c->tcd_min_res) {
        static bool tcd_attr_done[16] = {false};
        if (!tcd_attr_done[c->device & 15]) {
            HN_CUDA(cudaFuncSetAttribute(tcd::down_tcr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcd::SMEM_BYTES));
            tcd_attr_done[c->device & 15] = true;
        }
        tcd::Args t;
        t.in = c->skip[d];
        t.bmat = reinterpret_cast<const __half*>(c->tcw + W.down[d].tcr);
        for (int i = 0; i < 8; i++) t.bias[i] = c->whost[W.down[d].b + i];
        t.out = c->x[d + 1];
        t.amax_in = c->amax + S_SKIP + d;
        t.amax_out = c->amax + S_X + d + 1;
        t.error_flag = c->err_flag;
        t.w_inv_scale = W.down[d].tc_inv;
        t.H = r;
        t.W = r;
        t.nsx = (r / 2 + tcr::CW - 1) / tcr::CW;
        // narrow levels: several images side by side in one M = 128 MMA (conv_tcr_down.cuh: Args::pack); the strip walk below then
        // runs over groups of images
        t.pack = pick_pack(c, B, r / 2, r / 2, 2 * c->num_sms, 5); t.pack_s = r / 2 + 2; t.batch = B;
        if (t.pack > 0) B = (B + t.pack - 1) / t.pack;
        {
            // output rows per strip: whole rounds of equal strips over two CTAs per SM; a strip of R output rows streams
            // R + 3 input row pairs (+ ~2 steps of pipeline fill).  Small batches get short strips (one per CTA slot) instead
            // of the fixed 32 rows of r1 -- 256^2 x 32: down[0..3] 41 / 30 / 29 / 19 us on 128 / 64 / 32 / 32 CTAs before.
            const int Ho = r / 2;
            int best = tcd::ROWS_O;
            long long best_cost = -1;
            const int env_rows = getenv("HELMNET_DOWN_ROWS") ? atoi(getenv("HELMNET_DOWN_ROWS")) : 0;
            for (int rows = 2; rows <= 128 && rows <= (Ho > 2 ? Ho : 2); rows += 2) {
                const long long total = (long long)t.nsx * ((Ho + rows - 1) / rows) * B;
                const long long g = total < 2 * c->num_sms ? total : 2 * c->num_sms;
                const long long cost = ((total + g - 1) / g) * (rows + 5);
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rows; }
            }
            // with 64+ output rows per SM the kernel is throughput-bound and what counts is the work per SM, not per CTA slot:
            // the 32-row strips of r1 balance best there (256^2 x 256: 175 us against 185 us with one 128-row strip per sample)
            if ((long long)B * Ho / c->num_sms >= 64) best = Ho < tcd::ROWS_O ? Ho : tcd::ROWS_O;
            t.rows_o = env_rows >= 1 ? env_rows : best;
        }
        t.nsy = (r / 2 + t.rows_o - 1) / t.rows_o;
        t.total_strips = t.nsx * t.nsy * B;
        int tgrid = t.total_strips < 2 * c->num_sms ? t.total_strips : 2 * c->num_sms;
        t.pdl_trig = pdl_early(c, t.total_strips, 2);
        t.bal = 0;
        if (c->dconv_balance >= 2 && t.nsx == 1) {     // balanced strips (common.cuh: balanced_strip) where the model predicts 3 %
            const int cap = 2 * c->num_sms;
            const long long uniform = (long long)((t.total_strips + tgrid - 1) / tgrid) * (t.rows_o + 5);
            const long long vt = (long long)B * (r / 2 + tcd::BAL_PAD);
            const long long balanced = (vt + cap - 1) / cap + tcd::BAL_PAD + 2;
            // (uniform 32-row strips at 64+ rows per SM pair a long and a short strip list per SM: compare per SM, not per slot)
            const long long uniform_sm = (long long)B * (r / 2) / c->num_sms >= 64 ? ((long long)t.total_strips * (t.rows_o + 5) + c->num_sms - 1) / c->num_sms : 2 * uniform;
            if (vt / cap >= 12 && 2 * balanced * 100 <= uniform_sm * c->bal_thresh) {
                t.bal = B;
                tgrid = cap;
                t.pdl_trig = c->pdl_mode == 1 || c->pdl_mode == 2;
            }
        }
        HN_LAUNCH_PDL(c->pdl, (tcd::down_tcr_kernel), dim3(tgrid), dim3(tcr::THREADS), tcd::SMEM_BYTES, st, t);
        c->launches++;
        return HN_OK;
    }
#endif
    dim3 g((r / 2 + DN_TX - 1) / DN_TX, (r / 2 + DN_TY - 1) / DN_TY, B);
    HN_LAUNCH_PDL(c->pdl, down_kernel, g, dim3(DN_THREADS), DN_SMEM, st, dn);
    c->launches++;
    return HN_OK;
}

static int launch_up(hn_ctx* c, int d, int B, cudaStream_t st) {
    const Weights& W = c->W;
    const int r = c->r[d];
    UpArgs up;
    up.in = (d == kDepth - 1) ? c->bot : c->dec[d + 1];
    up.w = c->wdev + W.up[d].w;
    up.bias = c->wdev + W.up[d].b;
    up.out = c->upo[d];
    up.amax_out = c->amax + S_UPO + d;
    up.Hi = r / 2;
    up.Wi = r / 2;
#ifdef HN_HAVE_TC
    if (c->engine >= 1 && W.up[d].tcr != (size_t)-1 && r / 2 >= c->tcd_min_res && ((r / 2) % 2) == 0) {
        static bool tcu_attr_done[16] = {false};
        if (!tcu_attr_done[c->device & 15]) {
            HN_CUDA(cudaFuncSetAttribute(tcu::up_tcr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcu::SMEM_BYTES));
            tcu_attr_done[c->device & 15] = true;
        }
        tcu::Args t;
        t.in = up.in;
        t.bmat = reinterpret_cast<const __half*>(c->tcw + W.up[d].tcr);
        for (int i = 0; i < 8; i++) t.bias[i] = c->whost[W.up[d].b + i];
        t.out = up.out;
        t.amax_in = c->amax + ((d == kDepth - 1) ? S_BOT : S_DEC + d + 1);
        t.amax_out = up.amax_out;
        t.error_flag = c->err_flag;
        t.w_inv_scale = W.up[d].tc_inv;
        t.Hi = r / 2;
        t.Wi = r / 2;
        t.nsx = (r / 2 + tcr::CW - 1) / tcr::CW;
        // narrow levels: several images side by side in one M = 128 MMA (conv_tcr_up.cuh: Args::pack)
        t.pack = pick_pack(c, B, r / 2, r / 2, c->num_sms, 6); t.pack_s = r / 2 + 2; t.batch = B;
        if (t.pack > 0) B = (B + t.pack - 1) / t.pack;
        {
            // input rows per strip: whole rounds of equal strips over one CTA per SM; a strip of R rows streams R + 4 rows
            // (+ ~2 row steps of pipeline fill)
            int best = tcu::ROWS_I;
            long long best_cost = -1;
            const int env_rows = getenv("HELMNET_UP_ROWS") ? atoi(getenv("HELMNET_UP_ROWS")) : 0;
            for (int rows = 2; rows <= 128 && rows <= t.Hi; rows += 2) {      // (r1 started at 8: too few strips for small batches)
                if (t.Hi % rows != 0) continue;
                const long long total = (long long)t.nsx * (t.Hi / rows) * B;
                const long long g = total < c->num_sms ? total : c->num_sms;
                const long long cost = ((total + g - 1) / g) * (rows + 6);
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rows; }
            }
            t.rows_i = (env_rows >= 2 && env_rows % 2 == 0) ? env_rows : best;
        }
        t.nsy = (t.Hi + t.rows_i - 1) / t.rows_i;
        t.total_strips = t.nsx * t.nsy * B;
        int tgrid = t.total_strips < c->num_sms ? t.total_strips : c->num_sms;   // one CTA per SM (owns all 512 TMEM columns)
        t.pdl_trig = pdl_early(c, t.total_strips, 1);
        t.bal = 0;
        if (c->dconv_balance >= 2 && t.nsx == 1) {     // balanced strips (common.cuh: balanced_strip) where the model predicts 3 %
            const int cap = c->num_sms;
            const long long uniform = (long long)((t.total_strips + tgrid - 1) / tgrid) * (t.rows_i + 6);
            const long long vt = (long long)B * (t.Hi + tcu::BAL_PAD);
            const long long balanced = (vt + cap - 1) / cap + tcu::BAL_PAD + 2;
            if (vt / cap >= 12 && balanced * 100 <= uniform * c->bal_thresh) {
                t.bal = B;
                tgrid = cap;
            }
        }
        HN_LAUNCH_PDL(c->pdl, (tcu::up_tcr_kernel), dim3(tgrid), dim3(tcr::THREADS), tcu::SMEM_BYTES, st, t);
        c->launches++;
    } else
#endif
    {
        dim3 g((r / 2 + UP_TL - 1) / UP_TL, (r / 2 + UP_TL - 1) / UP_TL, B);
        HN_LAUNCH_PDL(c->pdl, up_kernel, g, dim3(UP_THREADS), UP_SMEM, st, up);
        c->launches++;
    }
    return HN_OK;
}

#ifdef HN_HAVE_TC
// Fused DoubleConv (conv_tcf.cuh) when the engine and the level allow it: full-width rows of 8 .. 256 pixels.
// Returns 1 when launched, 0 when the caller has to fall back to two launches, negative on error.
static int dconv_rows_per_strip(int H, int B, int cap, int min_rows) {
    int best = min_rows;
    long long best_cost = -1;
    for (int rows = min_rows; rows <= 128 && rows <= H; rows += 2) {
        const int spi = (H + rows - 1) / rows;
        const long long total = (long long)spi * B;
        const long long g = total < cap ? total : cap;
        const long long rounds = (total + g - 1) / g;
        const long long cost = rounds * (rows + 4 + 3);   // + halo rows + pipeline fill/drain per strip
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rows; }
    }
    return best;
}
template <int SRC, int NH, int EPI>
static int launch_dconv_nh(hn_ctx* c, const tcf::Args& t0, int B, cudaStream_t st) {
    static bool attr_done[16] = {false};
    if (!attr_done[c->device & 15]) {
        HN_CUDA(cudaFuncSetAttribute(tcf::dconv_tcf_kernel<SRC, NH, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tcf::smem_bytes(SRC, NH)));
        attr_done[c->device & 15] = true;
    }
    tcf::Args t = t0;
    const int cap = (NH == 2 ? 1 : 2) * c->num_sms;
    t.rows = dconv_rows_per_strip(t.H, B, cap, c->dconv_min_rows);
    t.spi = (t.H + t.rows - 1) / t.rows;
    t.total_strips = t.spi * B;
    int grid = t.total_strips < cap ? t.total_strips : cap;
    t.pdl_trig = pdl_early(c, t.total_strips, NH == 2 ? 1 : 2);
    // Balanced strips (conv_tcf.cuh: strip_of): one chunk of the images' rows laid end to end per CTA slot instead of whole rounds
    // of equal strips -- when the number of images is no multiple of the SMs (256^2 x 32: 128 strips of 64 rows on 148 SMs, 71 row
    // steps each; balanced: 148 chunks of ~57 steps).  Taken when the model (rows + 7 per strip, as above) predicts a gain of 3 %.
    if (c->dconv_balance) {
        const long long uniform = (long long)((t.total_strips + grid - 1) / grid) * (t.rows + 7);
        const long long vt = (long long)B * (t.H + tcf::BAL_PAD);
        const long long balanced = (vt + cap - 1) / cap + tcf::BAL_PAD + 2;      // chunk + one strip start + even rounding
        if (vt / cap >= 12 && balanced * 100 <= uniform * c->bal_thresh) {
            t.bal = B;
            grid = cap;
            t.pdl_trig = c->pdl_mode == 1 || c->pdl_mode == 2;     // equal work per CTA: placement does not matter
        }
    }
    HN_LAUNCH_PDL(c->pdl, (tcf::dconv_tcf_kernel<SRC, NH, EPI>), dim3(grid), dim3(tcf::threads(NH)), tcf::smem_bytes(SRC, NH), st, t);
    c->launches++;
    return 1;
}
template <int SRC, int EPI>
static int launch_dconv(hn_ctx* c, const ConvW (&w)[2], const float* inA, const float* inB, float* out, int r, int slot_out, int slot_in0,
                        int slot_in1, int B, cudaStream_t st, const ConvW* outc = nullptr, float* wf = nullptr, float* dwf_out = nullptr) {
    // full-width rows of up to 256 pixels (even width and height); widths other than 32 / 64 / 128 / 256 run on the next larger
    // variant with the GEMM rows beyond the image left zero (HELMNET_TCF_ANY_WIDTH=0: only the four exact widths)
    const bool exact = r == 32 || r == 64 || r == 128 || r == 256;
    if (c->engine < 2 || r > 256 || r < c->tcf_min_width || (r & 1) || (!exact && !c->tcf_any_width) || w[0].tcr == (size_t)-1 ||
        w[1].tcr == (size_t)-1 || !c->tcw)
        return 0;
    tcf::Args t;
    memset(&t, 0, sizeof(t));
    t.inA = inA; t.inB = inB; t.sigma = c->sigma1d;
    t.bmat1 = reinterpret_cast<const __half*>(c->tcw + w[0].tcr);
    t.bmat2 = reinterpret_cast<const __half*>(c->tcw + w[1].tcr);
    if (SRC == SRC_A8_B2) {
        if (w[0].tcf == (size_t)-1) return 0;
        t.bfold1 = reinterpret_cast<const __half*>(c->tcw + w[0].tcf);
    }
    if (EPI == EPI_STORE2) {
        if (w[1].tcf == (size_t)-1) return 0;
        t.bfold2 = reinterpret_cast<const __half*>(c->tcw + w[1].tcf);
    }
    const float* hw = c->whost.data();
    for (int i = 0; i < 8; i++) { t.bias1[i] = hw[w[0].b8 + i]; t.bias2[i] = hw[w[1].b8 + i]; }
    t.slope = hw[w[0].slope];
    t.out = out;
    if (outc) {
        for (int i = 0; i < 16; i++) t.wo[i] = hw[outc->w + i];
        t.bo[0] = hw[outc->b]; t.bo[1] = hw[outc->b + 1];
    }
    t.wf = wf; t.dwf_out = dwf_out;
    t.amax_in0 = c->amax + slot_in0;
    t.amax_in1 = c->amax + (slot_in1 >= 0 ? slot_in1 : slot_in0);
    t.amax_out = slot_out >= 0 ? c->amax + slot_out : nullptr;
    t.error_flag = c->err_flag;
    t.sigma_max = c->pml > 0 ? (float)c->sigma_max : 0.f;
    t.w_inv1 = w[0].tc_inv; t.w_inv2 = w[1].tc_inv;
    t.mid_l1 = w[0].l1; t.mid_bmax = w[0].bmax;
    t.H = r;
    t.W = r;
    // narrow levels: several images side by side in the 128-pixel variant (conv_tcf.cuh: Args::pack), picked by the step model
    if constexpr (SRC != SRC_INC && EPI != EPI_OUTC) {
        if (c->pack_narrow >= 2) {
            const int g = pick_pack(c, B, r, r, 2 * c->num_sms, 7);
            if (g > 1) {
                t.pack = g;
                t.pack_s = r + 2;
                t.batch = B;
                return launch_dconv_nh<SRC, 1, EPI>(c, t, (B + g - 1) / g, st);
            }
        }
    }
    if (r > 128) return launch_dconv_nh<SRC, 2, EPI>(c, t, B, st);
    if (r > 64) return launch_dconv_nh<SRC, 1, EPI>(c, t, B, st);
    if (r > 32) return launch_dconv_nh<SRC, 0, EPI>(c, t, B, st);
    return launch_dconv_nh<SRC, -1, EPI>(c, t, B, st);
}
#define HN_TRY_DCONV(var, expr) \
    int var = (expr);           \
    if (var < 0) return var
#else
#define HN_TRY_DCONV(var, expr) int var = 0
#endif

// HybridNet.forward (architectures.py:439-465).  `from_in6`: read the 6-channel input from c->in6 instead of
// building it from (wf, 1e3*res, sigmas); `raw_out`: store the network output to c->dwf instead of updating wf.
static int launch_unet(hn_ctx* c, int B, cudaStream_t st, bool from_in6, bool raw_out, bool defer_join = false, int* advance_slot = nullptr) {
    const Weights& W = c->W;
    pdl_select(c, B);
    const int cur = c->cur, nxt = cur ^ 1;
    // zero the amax slots of everything this pass (re)produces; keep the current hidden-state slots and in6
    {
        unsigned long long mask = 0;
        for (int i = 0; i < S_COUNT; i++) mask |= 1ull << i;
        for (int d = 0; d < kDepth; d++) mask &= ~(1ull << (S_STATE + 2 * d + cur));
        mask &= ~(1ull << S_IN6);
        mask &= ~(1ull << (S_WF + cur));
        mask &= ~(1ull << (S_RES + cur));
        HN_LAUNCH_PDL(c->pdl, reset_amax_kernel, dim3(1), dim3(64), 0, st, c->amax, mask, advance_slot);
        c->launches++;
    }
    // inc
    {
        HN_TRY_DCONV(fused, from_in6 ? 0 : (launch_dconv<SRC_INC, EPI_STORE>(c, W.inc, c->wf, c->res, c->x[0], c->r[0], S_X + 0, S_WF + cur,
                                                                              S_RES + cur, B, st)));
        if (!fused) {
            Conv3Args a = conv_args(c, W.inc[0], from_in6 ? c->in6 : c->wf, c->res, c->mid[0], c->r[0], S_IMID, from_in6 ? S_IN6 : S_WF + cur,
                                    from_in6 ? -1 : S_RES + cur);
            if (from_in6) HN_TRY((launch_conv3<SRC_A8, 8, true, EPI_STORE>(c, a, B, st)));
            else HN_TRY((launch_conv3<SRC_INC, 8, true, EPI_STORE>(c, a, B, st)));
            Conv3Args a2 = conv_args(c, W.inc[1], c->mid[0], nullptr, c->x[0], c->r[0], S_X + 0, S_IMID);
            HN_TRY((launch_conv3<SRC_A8, 8, false, EPI_STORE>(c, a2, B, st)));
        }
    }
    // encoder
    for (int d = 0; d < kDepth; d++) {
        const int r = c->r[d];
        HN_TRY_DCONV(fsig, (launch_dconv<SRC_A8_B2, EPI_STORE>(c, W.sig[d], c->x[d], c->state[d][cur], c->skip[d], r, S_SKIP + d, S_X + d,
                                                                S_STATE + 2 * d + cur, B, st)));
        if (!fsig) {
            Conv3Args s0 = conv_args(c, W.sig[d][0], c->x[d], c->state[d][cur], c->mid[d], r, S_MID + d, S_X + d, S_STATE + 2 * d + cur);
            HN_TRY((launch_conv3<SRC_A8_B2, 8, true, EPI_STORE>(c, s0, B, st)));
            Conv3Args s1 = conv_args(c, W.sig[d][1], c->mid[d], nullptr, c->skip[d], r, S_SKIP + d, S_MID + d);
            HN_TRY((launch_conv3<SRC_A8, 8, false, EPI_STORE>(c, s1, B, st)));
        }
        // hidden-state update: nothing else of this iteration reads it, so it may run on the side branch (no PDL edge there)
        cudaStream_t ss;
        HN_TRY(side_fork(c, d, st, &ss));
        const bool pdl_main = c->pdl;
        if (ss != st) c->pdl = false;
        HN_TRY_DCONV(fsta, (launch_dconv<SRC_A8_B2, EPI_STORE2>(c, W.sta[d], c->skip[d], c->state[d][cur], c->state[d][nxt], r,
                                                                 S_STATE + 2 * d + nxt, S_SKIP + d, S_STATE + 2 * d + cur, B, ss)));
        if (!fsta) {
            Conv3Args t0 = conv_args(c, W.sta[d][0], c->skip[d], c->state[d][cur], c->mid2[d], r, -1, S_SKIP + d, S_STATE + 2 * d + cur);
            HN_TRY((launch_conv3<SRC_A8_B2, 2, true, EPI_STORE>(c, t0, B, ss)));
            if (c->lean_state2) {
                State2Args s2;
                s2.in = c->mid2[d];
                s2.w = c->wdev + W.sta[d][1].raw;
                s2.bias = c->wdev + W.sta[d][1].b;
                s2.out = c->state[d][nxt];
                s2.amax_out = c->amax + S_STATE + 2 * d + nxt;
                s2.H = r;
                s2.W = r;
                HN_LAUNCH_PDL(c->pdl, state2_kernel, dim3((r + S2_TX - 1) / S2_TX, (r + S2_TY - 1) / S2_TY, B), dim3(S2_THREADS), 0, ss, s2);
                c->launches++;
            } else {
                Conv3Args t1 = conv_args(c, W.sta[d][1], c->mid2[d], nullptr, c->state[d][nxt], r, S_STATE + 2 * d + nxt);
                HN_TRY((launch_conv3<SRC_A2, 2, false, EPI_STORE>(c, t1, B, ss)));
            }
        }
        c->pdl = pdl_main;
        HN_TRY(launch_down(c, d, B, st));
    }
    // bottom
    {
        const int r = c->r[kDepth];
        HN_TRY_DCONV(fbot, c->fuse_bottom ? (launch_dconv<SRC_A8, EPI_STORE>(c, W.bot, c->x[kDepth], nullptr, c->bot, r, S_BOT, S_X + kDepth, -1, B, st)) : 0);
        if (!fbot) {
            Conv3Args b0 = conv_args(c, W.bot[0], c->x[kDepth], nullptr, c->mid[kDepth], r, S_MID + kDepth, S_X + kDepth);
            HN_TRY((launch_conv3<SRC_A8, 8, true, EPI_STORE>(c, b0, B, st)));
            Conv3Args b1 = conv_args(c, W.bot[1], c->mid[kDepth], nullptr, c->bot, r, S_BOT, S_MID + kDepth);
            HN_TRY((launch_conv3<SRC_A8, 8, false, EPI_STORE>(c, b1, B, st)));
        }
    }
    // decoder
    for (int d = kDepth - 1; d >= 0; d--) {
        const int r = c->r[d];
        HN_TRY(launch_up(c, d, B, st));
        // mid[d] is reused as scratch by inc / encoder / decoder; each use has its own amax slot
        if (d > 0) {
            HN_TRY_DCONV(fdec, (launch_dconv<SRC_A8_B8, EPI_STORE>(c, W.dec[d], c->upo[d], c->skip[d], c->dec[d], r, S_DEC + d, S_UPO + d,
                                                                    S_SKIP + d, B, st)));
            if (fdec) continue;
        } else {
            HN_TRY_DCONV(fdec, (launch_dconv<SRC_A8_B8, EPI_OUTC>(c, W.dec[d], c->upo[d], c->skip[d], c->dec[d], r, raw_out ? -1 : S_WF + nxt,
                                                                   S_UPO + d, S_SKIP + d, B, st, &W.outc, c->wf, raw_out ? c->dwf : nullptr)));
            if (fdec) continue;
        }
        Conv3Args d0 = conv_args(c, W.dec[d][0], c->upo[d], c->skip[d], c->mid[d], r, S_DMID + d, S_UPO + d, S_SKIP + d);
        HN_TRY((launch_conv3<SRC_A8_B8, 8, true, EPI_STORE>(c, d0, B, st)));
        Conv3Args d1 = conv_args(c, W.dec[d][1], c->mid[d], nullptr, c->dec[d], r, S_DEC + d, S_DMID + d);
        if (d > 0) {
            HN_TRY((launch_conv3<SRC_A8, 8, false, EPI_STORE>(c, d1, B, st)));
        } else {
            d1.wo = c->wdev + W.outc.w;
            d1.bo = c->wdev + W.outc.b;
            d1.wf = c->wf;
            d1.dwf_out = raw_out ? c->dwf : nullptr;
            d1.amax_out = raw_out ? nullptr : c->amax + S_WF + nxt;
            HN_TRY((launch_conv3<SRC_A8, 8, false, EPI_OUTC>(c, d1, B, st)));
        }
    }
    if (!defer_join) HN_TRY(side_join(c, st));
    return HN_OK;
}

// r = L(u) + ksq*u - src  (any of ksq/src/ssq may be null)
static int launch_spectral(hn_ctx* c, int B, cudaStream_t st, const float* u, const float* ksq, const float* src,
                           int src_batch, float* res, double* ssq, const int* slot, unsigned* amax_out = nullptr) {
    const int n = c->n;
    pdl_select(c, B);
    const int L = c->rows_L, CW = c->cols_CW;
    // Row and column passes alternate over chunks of samples small enough that u, rx and k_sq of a chunk stay in the
    // 126 MB L2 between the two kernels (20 B per point), so the column pass re-reads them from L2, not HBM.
    // (Measured at 256^2 x 256: not worth it -- 0.70 ms unchunked vs 0.77 ms in chunks of 64; the kernels are bound by
    // their shared-memory FFT passes, not by HBM, so the default is one pair of launches.)
    int chunk = c->spec_chunk > 0 ? c->spec_chunk : B;
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = (B - b0 < chunk) ? B - b0 : chunk;
        const size_t off = (size_t)b0 * n * n;
        const int total_rows = nb * n;
        const bool fast256 = (n == 256 && c->pml <= 16 && c->spec_fast);
        const bool fast512 = (n == 512 && c->pml <= 16 && c->spec_fast);
        const bool fast1024 = (n == 1024 && c->pml <= 16 && c->spec_fast);
        if (fast256)
            HN_LAUNCH_PDL(c->pdl, s256::spectral_rows256_kernel, dim3((total_rows + s256::LINES - 1) / s256::LINES), dim3(s256::THREADS), 0, st,
                      c->spec, reinterpret_cast<const float2*>(u) + off, reinterpret_cast<float2*>(c->rx) + off, total_rows);
        else if (fast512)
            HN_LAUNCH_PDL(c->pdl, s512::spectral_rows512_kernel, dim3((total_rows + s512::LINES - 1) / s512::LINES), dim3(s512::THREADS),
                      s512::ROWS_SMEM_BYTES, st, c->spec, reinterpret_cast<const float2*>(u) + off,
                      reinterpret_cast<float2*>(c->rx) + off, total_rows);
        else if (fast1024)
            HN_LAUNCH_PDL(c->pdl, s1024::spectral_rows1024_kernel, dim3((total_rows + s1024::ROWS_LINES - 1) / s1024::ROWS_LINES),
                      dim3(s1024::ROWS_THREADS), s1024::ROWS_SMEM_BYTES, st, c->spec, reinterpret_cast<const float2*>(u) + off,
                      reinterpret_cast<float2*>(c->rx) + off, total_rows);
        else
            HN_LAUNCH_PDL(c->pdl, spectral_rows_kernel, dim3((total_rows + L - 1) / L), dim3(SPEC_THREADS), spectral_smem_bytes(n, L, c->pml), st,
                      c->spec, reinterpret_cast<const float2*>(u) + off, reinterpret_cast<float2*>(c->rx) + off, total_rows, L);
        ColsArgs a;
        a.u = reinterpret_cast<const float2*>(u) + off;
        a.rx = reinterpret_cast<const float2*>(c->rx) + off;
        a.ksq = ksq ? ksq + off : nullptr;
        a.src = src ? reinterpret_cast<const float2*>(src) + (src_batch > 1 ? off : 0) : nullptr;
        a.src_nz = (src != nullptr && src == c->src && c->src_skip && (n % 8) == 0) ? c->src_nz + (src_batch > 1 ? (size_t)b0 * n : 0) : nullptr;
        a.res = reinterpret_cast<float2*>(res) + off;
        a.ssq = ssq;
        a.slot = slot;
        a.amax_out = amax_out;
        a.src_batch = src_batch;
        a.B = B;
        a.b0 = b0;
        a.CW = CW;
        if (fast256)
            HN_LAUNCH_PDL(c->pdl, s256::spectral_cols256_kernel, dim3(n / s256::LINES, nb), dim3(s256::THREADS), s256::COLS_SMEM_BYTES, st, c->spec, a);
        else if (fast512)
            HN_LAUNCH_PDL(c->pdl, s512::spectral_cols512_kernel, dim3(n / s512::LINES, nb), dim3(s512::THREADS), s512::COLS_SMEM_BYTES, st, c->spec, a);
        else if (fast1024)
            HN_LAUNCH_PDL(c->pdl, s1024::spectral_cols1024_kernel, dim3(n / s1024::COLS, nb), dim3(s1024::COLS_THREADS), s1024::COLS_SMEM_BYTES, st,
                      c->spec, a);
        else
            HN_LAUNCH_PDL(c->pdl, spectral_cols_kernel, dim3((n + CW - 1) / CW, nb), dim3(SPEC_THREADS), spectral_smem_bytes(n, CW, c->pml), st,
                      c->spec, a);
        c->launches += 2;
    }
    return HN_OK;
}

static int launch_iteration(hn_ctx* c, int B, cudaStream_t st) {
    // the iteration slot (row of the residual-norm history) is advanced by the first kernel of the iteration: hn_run starts it at -1
    HN_TRY(launch_unet(c, B, st, false, false, true, c->iter_dev));
    HN_TRY(launch_spectral(c, B, st, c->wf, c->ksq, c->src, c->src_batch, c->res, c->ssq, c->iter_dev, c->amax + S_RES + (c->cur ^ 1)));
    HN_TRY(side_join(c, st));      // the conv_state branch joins after the residual stage
    return HN_OK;
}

#include "train_host.cuh"

// ------------------------------------------------------------------------------------------------
// ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* hn_version(void) {
#ifdef HN_EMU
    return "helmnet_sm100 0.1.0 EMULATOR (tests only)";
#else
    return "helmnet_sm100 0.1.0 sm_100a";
#endif
}
const char* hn_last_error(void) { return g_err.c_str(); }

int hn_create(hn_ctx** out, int device, int n, int max_batch, int pml_size, double sigma_max, double k0, double omega) {
    if (!out) return fail(HN_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (n <= 0 || n % 16 != 0) return fail(HN_ERR_ARG, "domain size must be a positive multiple of 16");
    if (max_batch <= 0) return fail(HN_ERR_ARG, "max_batch must be positive");
    if (pml_size < 0 || 2 * pml_size > n) return fail(HN_ERR_ARG, "PML does not fit the domain");
    if (k0 == 0.0) return fail(HN_ERR_ARG, "k must be non-zero");
#ifndef HN_EMU
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(HN_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(HN_ERR_ARG, "bad device ordinal");
    DeviceGuard dev_guard(device);
    cudaDeviceProp prop;
    HN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(HN_ERR_CUDA, "libhelmnet_sm100 needs a compute capability 10.x (Blackwell) device");
    const int num_sms_ = prop.multiProcessorCount;
#else
    const int num_sms_ = 148;
#endif
    hn_ctx* c = new hn_ctx();
    c->num_sms = num_sms_;
    c->device = device;
    c->n = n;
    c->max_batch = max_batch;
    c->pml = pml_size;
    c->sigma_max = sigma_max;
    c->k0 = k0;
    c->omega = omega;
    c->state_len = 0;
    for (int d = 0; d <= kDepth; d++) c->r[d] = n >> d;
    for (int d = 0; d < kDepth; d++) c->state_len += c->r[d] * c->r[d];
    const char* ng = getenv("HELMNET_NO_GRAPH");
    c->use_graph = !(ng && ng[0] == '1');
#ifdef HN_EMU
    c->use_graph = false;
#endif
    auto cleanup = [&](int rc) {
        hn_destroy(c);
        return rc;
    };
    int rc = build_tables(c);
    if (rc != HN_OK) return cleanup(rc);
    const size_t B = (size_t)max_batch, hw = (size_t)n * n;
#define A_(ptr, cnt)                                      \
    do {                                                  \
        rc = dalloc_t(c, &(ptr), (cnt));                  \
        if (rc != HN_OK) return cleanup(rc);              \
    } while (0)
    A_(c->wf, B * hw * 2);
    A_(c->res, B * hw * 2);
    A_(c->ksq, B * hw);
    A_(c->src, B * hw * 2);
    A_(c->rx, B * hw * 2);
    A_(c->src_nz, B * (size_t)n + 16);
    A_(c->tmp2, B * hw * 2);
    A_(c->tmp2b, B * hw * 2);
    A_(c->dwf, B * hw * 2);
    A_(c->in6, B * hw * 8);
    for (int d = 0; d <= kDepth; d++) {
        const size_t p = (size_t)c->r[d] * c->r[d];
        A_(c->x[d], B * p * 8);
        A_(c->mid[d], B * p * 8);
        if (d < kDepth) {
            A_(c->state[d][0], B * p * 2);
            A_(c->state[d][1], B * p * 2);
            A_(c->mid2[d], B * p * 2);
            A_(c->skip[d], B * p * 8);
            A_(c->upo[d], B * p * 8);
            A_(c->dec[d], B * p * 8);
        }
    }
    A_(c->bot, B * (size_t)c->r[kDepth] * c->r[kDepth] * 8);
    A_(c->ssq1, B);
    A_(c->iter_dev, 4);
    A_(c->wdev, 65536);
    A_(c->wraw, HN_NUM_WEIGHTS);
    A_(c->tcw, 458752);
    A_(c->err_flag, 4);
    A_(c->amax, 64);
#undef A_
    if (cudaMemset(c->amax, 0, 256) != cudaSuccess) return cleanup(fail(HN_ERR_CUDA, "cudaMemset failed"));
    if (cudaMemset(c->err_flag, 0, 16) != cudaSuccess) return cleanup(fail(HN_ERR_CUDA, "cudaMemset failed"));
    if (const char* mr = getenv("HELMNET_TC_MIN_RES")) c->tc_min_res = atoi(mr);
    if (const char* mr = getenv("HELMNET_TCR_MIN_RES")) c->tcr_min_res = atoi(mr);
    if (const char* mr = getenv("HELMNET_TCD_MIN_RES")) c->tcd_min_res = atoi(mr);
    if (const char* pv = getenv("HELMNET_PDL")) c->pdl_cfg = atoi(pv);
    if (const char* pv = getenv("HELMNET_SRC_SKIP")) c->src_skip = atoi(pv) != 0;
    if (const char* pv = getenv("HELMNET_FUSE_BOTTOM")) c->fuse_bottom = atoi(pv) != 0;
    if (const char* pv = getenv("HELMNET_SIDE_STATE")) c->side_cfg = atoi(pv);
    if (const char* pv = getenv("HELMNET_DCONV_BALANCE")) c->dconv_balance = atoi(pv);
    if (const char* pv = getenv("HELMNET_BAL_THRESH")) c->bal_thresh = atoi(pv);
    if (const char* pv = getenv("HELMNET_PACK_NARROW")) c->pack_narrow = atoi(pv);
    if (const char* pv = getenv("HELMNET_PACK_PENALTY")) c->pack_penalty = atoi(pv);
#ifndef HN_EMU
    if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess) return cleanup(fail(HN_ERR_CUDA, "cudaStreamCreate failed"));
    for (int d = 0; d < kDepth; d++)
        if (cudaEventCreateWithFlags(&c->ev_fork[d], cudaEventDisableTiming) != cudaSuccess) return cleanup(fail(HN_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) return cleanup(fail(HN_ERR_CUDA, "cudaEventCreate failed"));
#endif
    if (const char* pv = getenv("HELMNET_TCF_ANY_WIDTH")) c->tcf_any_width = atoi(pv) != 0;
    if (const char* pv = getenv("HELMNET_TCF_MIN_WIDTH")) c->tcf_min_width = atoi(pv);
    if (const char* pv = getenv("HELMNET_DCONV_MIN_ROWS")) { const int v = atoi(pv); if (v >= 2 && v % 2 == 0) c->dconv_min_rows = v; }
    pdl_select(c, max_batch);
    if (const char* en = getenv("HELMNET_ENGINE")) { const int ev = atoi(en); c->engine = ev < 0 ? 0 : ev > 2 ? 2 : ev; }
#ifndef HN_HAVE_TC
    c->engine = 0;
#endif
    if (cudaMemset(c->iter_dev, 0, 16) != cudaSuccess) return cleanup(fail(HN_ERR_CUDA, "cudaMemset failed"));
    for (int d = 0; d < kDepth; d++)
        for (int k = 0; k < 2; k++)
            if (cudaMemset(c->state[d][k], 0, B * (size_t)c->r[d] * c->r[d] * 8) != cudaSuccess)
                return cleanup(fail(HN_ERR_CUDA, "cudaMemset failed"));
#ifndef HN_EMU
    rc = set_smem_attrs(c);
    if (rc != HN_OK) return cleanup(rc);
#endif
    *out = c;
    return HN_OK;
}

int hn_destroy(hn_ctx* c) {
    if (!c) return HN_OK;
    DeviceGuard dev_guard(c->device);
#ifndef HN_EMU
    cudaDeviceSynchronize();
    for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
    if (c->side) cudaStreamDestroy(c->side);
    for (int d = 0; d < kDepth; d++)
        if (c->ev_fork[d]) cudaEventDestroy(c->ev_fork[d]);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
#endif
    for (void* p : c->allocs) cudaFree(p);
    if (c->ssq) cudaFree(c->ssq);
    if (c->tws) {
        cudaFree(c->tws->gacc);
        cudaFree(c->tws->base);
        delete c->tws;
    }
    delete c;
    return HN_OK;
}

int hn_load_weights(hn_ctx* c, const float* host_blob, size_t n_floats) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c || !host_blob) return fail(HN_ERR_ARG, "NULL argument");
    if (n_floats != HN_NUM_WEIGHTS) return fail(HN_ERR_ARG, "expected 48160 floats (HybridNet features=8 depth=4 state=2)");
    Packer pk;
    Cursor cur{host_blob, n_floats};
    c->W = Weights();      // per-layer norms (ConvW::l1 / bmax) are running maxima: start from zero on every load
    Weights& W = c->W;
    pack_double_conv(pk, cur, W.inc, 6, 8, 8);
    for (int d = 0; d < kDepth; d++) {   // module order inside EncoderBlock: conv_signal, down, conv_state
        pack_double_conv(pk, cur, W.sig[d], 10, 8, 8);
        const float* dw = cur.take(8 * 8 * 64);
        const float* db = cur.take(8);
        if (cur.ok) {
            W.down[d].w = pack_down(pk, dw);
            W.down[d].b = pack_vec(pk, db, 8);
#ifdef HN_HAVE_TC
            pack_tcd(pk, dw, W.down[d]);
#endif
        }
        pack_double_conv(pk, cur, W.sta[d], 10, 2, 2);
    }
    for (int d = 0; d < kDepth; d++) pack_double_conv(pk, cur, W.dec[d], 16, 8, 8);
    pack_double_conv(pk, cur, W.bot, 8, 8, 8);   // decode[depth]
    for (int d = 0; d < kDepth; d++) {
        const float* uw = cur.take(8 * 8 * 64);
        const float* ub = cur.take(8);
        if (cur.ok) {
            W.up[d].w = pack_up(pk, uw);
            W.up[d].b = pack_vec(pk, ub, 8);
#ifdef HN_HAVE_TC
            pack_tcu(pk, uw, W.up[d]);
#endif
        }
    }
    {
        const float* ow = cur.take(16);
        const float* ob = cur.take(2);
        if (cur.ok) {
            W.outc.w = pack_vec(pk, ow, 16);
            W.outc.b = pack_vec(pk, ob, 2);
        }
    }
    if (!cur.ok || cur.left != 0) return fail(HN_ERR_ARG, "weight blob does not match the HybridNet state_dict layout");
    if (pk.blob.size() > 65536) return fail(HN_ERR_STATE, "packed weights exceed the reserved buffer");
#ifndef HN_EMU
    HN_CUDA(cudaDeviceSynchronize());
    // The captured iteration graphs carry per-layer constants BY VALUE (biases, PReLU slopes, block scales, the 1x1 outc
    // weights are kernel parameters read from c->whost at capture time): drop them so the next hn_run re-captures.
    drop_graphs(c);
#endif
    HN_CUDA(cudaMemcpy(c->wdev, pk.blob.data(), pk.blob.size() * 4, cudaMemcpyHostToDevice));
    c->whost = pk.blob;   // host copy: small per-layer constants are passed to the tcgen05 kernels as launch parameters
    HN_CUDA(cudaMemcpy(c->wraw, host_blob, n_floats * sizeof(float), cudaMemcpyHostToDevice));
    c->wraw_host.assign(host_blob, host_blob + n_floats);
    if (pk.halfs.size() > 458752) return fail(HN_ERR_STATE, "tensor-core weight images exceed the reserved buffer");
    if (!pk.halfs.empty()) HN_CUDA(cudaMemcpy(c->tcw, pk.halfs.data(), pk.halfs.size() * 2, cudaMemcpyHostToDevice));
    c->weights_set = true;
    return HN_OK;
}

int hn_set_source(hn_ctx* c, const float* d_src, int src_batch, const int64_t strides[4], void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c || !d_src || !strides) return fail(HN_ERR_ARG, "NULL argument");
    if (src_batch < 1 || src_batch > c->max_batch) return fail(HN_ERR_ARG, "source batch out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t total = (size_t)src_batch * c->n * c->n;
    HN_LAUNCH(src_strided_to_c2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_src,
              reinterpret_cast<float2*>(c->src), c->n, total, (long long)strides[0], (long long)strides[1],
              (long long)strides[2], (long long)strides[3]);
    c->launches++;
    {
        const int cols = src_batch * c->n;
        HN_LAUNCH(src_colnz_kernel, dim3((cols + 127) / 128), dim3(128), 0, st, reinterpret_cast<const float2*>(c->src), c->src_nz, c->n, cols);
        c->launches++;
    }
    HN_CUDA(cudaGetLastError());
    c->src_batch = src_batch;
    return HN_OK;
}

int hn_point_sources(int device, int n, int count, const int32_t* d_locations, double amplitude, double arg, int smooth, float* d_out,
                     void* stream) {
    if (!d_locations || !d_out) return fail(HN_ERR_ARG, "NULL argument");
    if (n <= 0 || count <= 0) return fail(HN_ERR_ARG, "domain size and source count must be positive");
    if (smooth && n < 5) return fail(HN_ERR_ARG, "smoothed sources need a domain of at least 5 points");
#ifndef HN_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(HN_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(HN_ERR_ARG, "bad device ordinal");
#endif
    DeviceGuard dev_guard(device);
    const size_t total = (size_t)count * n * n;
    HN_LAUNCH(point_sources_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, (cudaStream_t)stream, reinterpret_cast<const int*>(d_locations),
              count, n, (float)amplitude, (float)cos(arg), (float)sin(arg), smooth, d_out);
    HN_CUDA(cudaGetLastError());
    return HN_OK;
}

static int check_ready(hn_ctx* c, int batch) {
    if (!c) return fail(HN_ERR_ARG, "ctx is NULL");
    if (batch < 1 || batch > c->max_batch) return fail(HN_ERR_ARG, "batch out of range for this context");
    if (!c->weights_set) return fail(HN_ERR_STATE, "hn_load_weights has not been called");
    if (c->src_batch == 0) return fail(HN_ERR_STATE, "hn_set_source has not been called");
    if (c->src_batch != 1 && c->src_batch != batch) return fail(HN_ERR_ARG, "source batch must be 1 or equal to the batch");
    return HN_OK;
}

int hn_reset(hn_ctx* c, const float* d_sos, int batch, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    HN_TRY(check_ready(c, batch));
    if (!d_sos) return fail(HN_ERR_ARG, "d_sos is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = c->n * c->n;
    const size_t total = (size_t)batch * hw;
    HN_CUDA(cudaMemsetAsync(c->amax + S_WF + c->cur, 0, sizeof(unsigned), st));
    HN_CUDA(cudaMemsetAsync(c->amax + S_RES + c->cur, 0, sizeof(unsigned), st));
    HN_LAUNCH(reset_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_sos, c->ksq, reinterpret_cast<float2*>(c->wf),
              reinterpret_cast<float2*>(c->res), reinterpret_cast<const float2*>(c->src), c->src_batch, (float)c->omega, hw,
              total, c->amax + S_RES + c->cur);
    c->launches++;
    for (int d = 0; d < kDepth; d++) {
        HN_CUDA(cudaMemsetAsync(c->state[d][c->cur], 0, (size_t)batch * c->r[d] * c->r[d] * 8, st));
        HN_CUDA(cudaMemsetAsync(c->amax + S_STATE + 2 * d + c->cur, 0, sizeof(unsigned), st));
    }
    HN_CUDA(cudaGetLastError());
    c->batch = batch;
    c->problem_set = true;
    return HN_OK;
}

int hn_set_state(hn_ctx* c, const float* d_wf, const float* d_res, const float* d_ksq, const float* d_hflat, int batch,
                 void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    HN_TRY(check_ready(c, batch));
    if (!(d_wf && d_res && d_ksq) && (d_wf || d_res || d_ksq) && !(c->problem_set && c->batch == batch))
        return fail(HN_ERR_STATE, "partial field update needs an existing solve state of the same batch");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = c->n * c->n;
    const size_t total = (size_t)batch * hw;
    if (d_wf) {
        HN_CUDA(cudaMemsetAsync(c->amax + S_WF + c->cur, 0, sizeof(unsigned), st));
        HN_LAUNCH(nchw2_to_c2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_wf, reinterpret_cast<float2*>(c->wf), hw, total,
                  c->amax + S_WF + c->cur);
        c->launches++;
    }
    if (d_res) {
        HN_CUDA(cudaMemsetAsync(c->amax + S_RES + c->cur, 0, sizeof(unsigned), st));
        HN_LAUNCH(nchw2_to_c2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_res, reinterpret_cast<float2*>(c->res), hw, total,
                  c->amax + S_RES + c->cur);
        c->launches++;
    }
    if (d_ksq) HN_CUDA(cudaMemcpyAsync(c->ksq, d_ksq, total * 4, cudaMemcpyDeviceToDevice, st));
    if (d_hflat) {
        size_t off = 0;
        for (int d = 0; d < kDepth; d++) {
            const int p = c->r[d] * c->r[d];
            const size_t tot = (size_t)batch * p;
            HN_CUDA(cudaMemsetAsync(c->amax + S_STATE + 2 * d + c->cur, 0, sizeof(unsigned), st));
            HN_LAUNCH(nchw2_strided_to_c2_kernel, dim3(grid1d(tot)), dim3(LAY_THREADS), 0, st, d_hflat + off,
                      reinterpret_cast<float2*>(c->state[d][c->cur]), p, tot, (size_t)2 * c->state_len, (size_t)c->state_len,
                      c->amax + S_STATE + 2 * d + c->cur);
            c->launches++;
            off += p;
        }
    }
    HN_CUDA(cudaGetLastError());
    if (d_wf && d_res && d_ksq) {
        c->batch = batch;
        c->problem_set = true;
    }
    return HN_OK;
}

static int copy_states_out(hn_ctx* c, float* d_hflat, int batch, cudaStream_t st) {
    size_t off = 0;
    for (int d = 0; d < kDepth; d++) {
        const int p = c->r[d] * c->r[d];
        const size_t tot = (size_t)batch * p;
        HN_LAUNCH(c2_to_nchw2_kernel, dim3(grid1d(tot)), dim3(LAY_THREADS), 0, st,
                  reinterpret_cast<const float2*>(c->state[d][c->cur]), d_hflat + off, p, tot, (size_t)2 * c->state_len,
                  (size_t)c->state_len);
        c->launches++;
        off += p;
    }
    return HN_OK;
}

int hn_get(hn_ctx* c, float* d_wf, float* d_res, float* d_hflat, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c) return fail(HN_ERR_ARG, "ctx is NULL");
    if (!c->problem_set) return fail(HN_ERR_STATE, "no solve state: call hn_reset or hn_set_state first");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = c->n * c->n;
    const size_t total = (size_t)c->batch * hw;
    if (d_wf) {
        HN_LAUNCH(c2_to_nchw2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(c->wf), d_wf,
                  hw, total, (size_t)2 * hw, (size_t)hw);
        c->launches++;
    }
    if (d_res) {
        HN_LAUNCH(c2_to_nchw2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(c->res), d_res,
                  hw, total, (size_t)2 * hw, (size_t)hw);
        c->launches++;
    }
    if (d_hflat) HN_TRY(copy_states_out(c, d_hflat, c->batch, st));
    HN_CUDA(cudaGetLastError());
    return HN_OK;
}

int hn_get_states(hn_ctx* c, float* d_hflat, int batch, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c || !d_hflat) return fail(HN_ERR_ARG, "NULL argument");
    if (batch < 1 || batch > c->max_batch) return fail(HN_ERR_ARG, "batch out of range for this context");
    HN_TRY(copy_states_out(c, d_hflat, batch, (cudaStream_t)stream));
    HN_CUDA(cudaGetLastError());
    return HN_OK;
}

#ifndef HN_EMU
static int get_graph(hn_ctx* c, int B, cudaGraphExec_t* out) {
    pdl_select(c, B);     // launch options that depend on the batch (PDL mode, side branch) are part of the graph
    // the source pointer offset / broadcast flag are baked into the spectral kernel's parameters: part of the key
    const long long key = ((long long)B << 10) | ((long long)(c->side_state ? 1 : 0) << 9) | ((long long)(c->src_batch > 1 ? 1 : 0) << 8) | ((long long)c->cur << 4) | (long long)c->engine;
    auto it = c->graphs.find(key);
    if (it != c->graphs.end()) {
        *out = it->second;
        return HN_OK;
    }
    cudaStream_t cs;
    HN_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    const int64_t before = c->launches;
    HN_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = launch_iteration(c, B, cs);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(cs, &g);
    c->kernels_per_iter = (int)(c->launches - before);
    c->launches = before;
    if (rc != HN_OK) {
        if (g) cudaGraphDestroy(g);
        cudaStreamDestroy(cs);
        return rc;
    }
    if (e != cudaSuccess) {
        cudaStreamDestroy(cs);
        return fail(HN_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(e));
    }
    cudaGraphExec_t ex = nullptr;
    e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    cudaStreamDestroy(cs);
    if (e != cudaSuccess) return fail(HN_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    c->graphs[key] = ex;
    *out = ex;
    return HN_OK;
}
#endif

int hn_run(hn_ctx* c, int n_iters, float* d_rmse, float* d_wf_hist, float* d_res_hist, float* d_h_hist, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c) return fail(HN_ERR_ARG, "ctx is NULL");
    if (!c->problem_set) return fail(HN_ERR_STATE, "no solve state: call hn_reset or hn_set_state first");
    if (n_iters < 0) return fail(HN_ERR_ARG, "n_iters < 0");
    if (n_iters == 0) return HN_OK;
    HN_TRY(check_ready(c, c->batch));
    cudaStream_t st = (cudaStream_t)stream;
    const int B = c->batch, hw = c->n * c->n;
    const size_t need = (size_t)n_iters * B;
    if (need > c->ssq_cap) {
#ifndef HN_EMU
        HN_CUDA(cudaStreamSynchronize(st));
        drop_graphs(c);   // graphs hold the old ssq pointer
#endif
        if (c->ssq) cudaFree(c->ssq);
        c->ssq = nullptr;
        size_t cap = need < 4096 ? 4096 : need;
        HN_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->ssq), cap * sizeof(double)));
        c->ssq_cap = cap;
    }
    HN_CUDA(cudaMemsetAsync(c->ssq, 0, need * sizeof(double), st));
    HN_CUDA(cudaMemsetAsync(c->iter_dev, 0xFF, sizeof(int), st));      // slot -1: incremented at the start of every iteration
    const size_t total = (size_t)B * hw;
    pdl_select(c, B);
    for (int it = 0; it < n_iters; it++) {
#ifndef HN_EMU
        if (c->use_graph) {
            cudaGraphExec_t ex;
            HN_TRY(get_graph(c, B, &ex));
            HN_CUDA(cudaGraphLaunch(ex, st));
            c->launches += c->kernels_per_iter;
        } else
#endif
        {
            const int64_t before = c->launches;
            HN_TRY(launch_iteration(c, B, st));
            c->kernels_per_iter = (int)(c->launches - before);
        }
        c->cur ^= 1;
        if (d_wf_hist) {
            HN_LAUNCH(c2_to_nchw2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(c->wf),
                      d_wf_hist + (size_t)it * total * 2, hw, total, (size_t)2 * hw, (size_t)hw);
            c->launches++;
        }
        if (d_res_hist) {
            HN_LAUNCH(c2_to_nchw2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(c->res),
                      d_res_hist + (size_t)it * total * 2, hw, total, (size_t)2 * hw, (size_t)hw);
            c->launches++;
        }
        if (d_h_hist) HN_TRY(copy_states_out(c, d_h_hist + (size_t)it * B * 2 * c->state_len, B, st));
    }
    if (d_rmse) {
        HN_LAUNCH(finalize_rmse_kernel, dim3(grid1d(need)), dim3(LAY_THREADS), 0, st, c->ssq, d_rmse, (int)need,
                  1.0 / (2.0 * (double)hw));
        c->launches++;
    }
    HN_CUDA(cudaGetLastError());
    return HN_OK;
}

int hn_residual(hn_ctx* c, const float* d_x, const float* d_ksq, float* d_out, float* d_rmse, int batch, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    HN_TRY(check_ready(c, batch));
    if (!d_x || !d_out) return fail(HN_ERR_ARG, "NULL argument");
    if (!d_ksq && !(c->problem_set && c->batch == batch)) return fail(HN_ERR_STATE, "no k_sq in the context for this batch");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = c->n * c->n;
    const size_t total = (size_t)batch * hw;
    HN_LAUNCH(nchw2_to_c2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_x, reinterpret_cast<float2*>(c->tmp2), hw, total, (unsigned*)nullptr);
    HN_CUDA(cudaMemsetAsync(c->ssq1, 0, (size_t)batch * sizeof(double), st));
    HN_TRY(launch_spectral(c, batch, st, c->tmp2, d_ksq ? d_ksq : c->ksq, c->src, c->src_batch, c->tmp2b, c->ssq1, c->iter_dev + 1));
    HN_LAUNCH(c2_to_nchw2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(c->tmp2b), d_out, hw,
              total, (size_t)2 * hw, (size_t)hw);
    c->launches += 2;
    if (d_rmse) {
        HN_LAUNCH(finalize_rmse_kernel, dim3(1), dim3(LAY_THREADS), 0, st, c->ssq1, d_rmse, batch, 1.0 / (2.0 * (double)hw));
        c->launches++;
    }
    HN_CUDA(cudaGetLastError());
    return HN_OK;
}

int hn_laplacian(hn_ctx* c, const float* d_x, float* d_out, int batch, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c || !d_x || !d_out) return fail(HN_ERR_ARG, "NULL argument");
    if (batch < 1 || batch > c->max_batch) return fail(HN_ERR_ARG, "batch out of range for this context");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = c->n * c->n;
    const size_t total = (size_t)batch * hw;
    HN_LAUNCH(nchw2_to_c2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_x, reinterpret_cast<float2*>(c->tmp2), hw, total, (unsigned*)nullptr);
    HN_TRY(launch_spectral(c, batch, st, c->tmp2, nullptr, nullptr, 1, c->tmp2b, nullptr, c->iter_dev + 1));
    HN_LAUNCH(c2_to_nchw2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(c->tmp2b), d_out, hw,
              total, (size_t)2 * hw, (size_t)hw);
    c->launches += 2;
    HN_CUDA(cudaGetLastError());
    return HN_OK;
}

int hn_unet(hn_ctx* c, const float* d_in, float* d_out, int batch, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c || !d_in || !d_out) return fail(HN_ERR_ARG, "NULL argument");
    if (batch < 1 || batch > c->max_batch) return fail(HN_ERR_ARG, "batch out of range for this context");
    if (!c->weights_set) return fail(HN_ERR_STATE, "hn_load_weights has not been called");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = c->n * c->n;
    const size_t total = (size_t)batch * hw;
    HN_CUDA(cudaMemsetAsync(c->amax + S_IN6, 0, sizeof(unsigned), st));
    HN_LAUNCH(nchw6_to_nhwc8_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_in, c->in6, hw, total, c->amax + S_IN6);
    c->launches++;
    HN_TRY(launch_unet(c, batch, st, true, true));
    c->cur ^= 1;
    HN_LAUNCH(c2_to_nchw2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(c->dwf), d_out, hw,
              total, (size_t)2 * hw, (size_t)hw);
    c->launches++;
    HN_CUDA(cudaGetLastError());
    return HN_OK;
}

int hn_step_backward(hn_ctx* c, const float* d_wf, const float* d_res, const float* d_ksq, const float* d_hflat, const float* d_g_wf,
                     const float* d_g_res, const float* d_g_hflat, float* d_gwf_in, float* d_gres_in, float* d_ghflat_in, float* d_gparams,
                     int batch, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c) return fail(HN_ERR_ARG, "ctx is NULL");
    if (batch < 1 || batch > c->max_batch) return fail(HN_ERR_ARG, "batch out of range for this context");
    if (!c->weights_set) return fail(HN_ERR_STATE, "hn_load_weights has not been called");
    if (!d_wf || !d_res || !d_ksq || !d_hflat || !d_gparams) return fail(HN_ERR_ARG, "NULL argument");
    return step_backward(c, d_wf, d_res, d_ksq, d_hflat, d_g_wf, d_g_res, d_g_hflat, d_gwf_in, d_gres_in, d_ghflat_in, d_gparams, batch,
                         (cudaStream_t)stream);
}

int hn_state_len(const hn_ctx* c) { return c ? c->state_len : HN_ERR_ARG; }
int64_t hn_launch_count(const hn_ctx* c) { return c ? c->launches : 0; }
int hn_kernels_per_iteration(const hn_ctx* c) { return c ? c->kernels_per_iter : HN_ERR_ARG; }

int hn_debug_tensor(hn_ctx* c, const char* name, float* d_out, int batch, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c || !name || !d_out) return fail(HN_ERR_ARG, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const std::string s(name);
    const float* src = nullptr;
    int r = 0;
    auto lvl = [&](size_t pos) { return (s.size() > pos && s[pos] >= '0' && s[pos] <= '4') ? s[pos] - '0' : -1; };
    if (s == "bot") { src = c->bot; r = c->r[kDepth]; }
    else if (s.rfind("skip", 0) == 0 && lvl(4) >= 0 && lvl(4) < kDepth) { src = c->skip[lvl(4)]; r = c->r[lvl(4)]; }
    else if (s.rfind("mid", 0) == 0 && lvl(3) >= 0) { src = c->mid[lvl(3)]; r = c->r[lvl(3)]; }
    else if (s.rfind("up", 0) == 0 && lvl(2) >= 0 && lvl(2) < kDepth) { src = c->upo[lvl(2)]; r = c->r[lvl(2)]; }
    else if (s.rfind("dec", 0) == 0 && lvl(3) >= 1 && lvl(3) < kDepth) { src = c->dec[lvl(3)]; r = c->r[lvl(3)]; }
    else if (s.rfind("x", 0) == 0 && lvl(1) >= 0) { src = c->x[lvl(1)]; r = c->r[lvl(1)]; }
    else return fail(HN_ERR_ARG, "unknown tensor name");
    const size_t total = (size_t)batch * r * r;
    HN_LAUNCH(nhwc8_to_nchw_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, src, d_out, r * r, total);
    c->launches++;
    HN_CUDA(cudaGetLastError());
    return 8;
}

int hn_set_engine(hn_ctx* c, int engine) {
    if (!c) return fail(HN_ERR_ARG, "ctx is NULL");
#ifdef HN_HAVE_TC
    if (engine < 0 || engine > 2)
        return fail(HN_ERR_ARG, "engine must be 0 (fp32 CUDA cores), 1 (tcgen05 split-fp16) or 2 (tcgen05 with fused DoubleConvs)");
#else
    if (engine != 0) return fail(HN_ERR_ARG, "only engine 0 is available in the emulator build");
#endif
    c->engine = engine;
    return c->engine;
}

int hn_sync_check(hn_ctx* c, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c) return fail(HN_ERR_ARG, "ctx is NULL");
#ifndef HN_EMU
    HN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    int flag = 0;
    HN_CUDA(cudaMemcpy(&flag, c->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag != 0) return fail(HN_ERR_CUDA, "tcgen05 watchdog: an MMA completion barrier was never signalled");
#endif
    return HN_OK;
}

// Average device time (CUDA events on `stream`) of `reps` back-to-back launches of ONE kernel of the iteration on the
// context's current buffers:  which = 0: inc conv #2 (8->8, level 0)   1: decode[0] conv #1 (16->8, level 0)
//   2: enc[0].down   3: up[0]   4: spectral rows   5: spectral cols.   Used by bench.py for the per-kernel roofline.
int hn_profile_layer(hn_ctx* c, int which, int reps, float* out_ms, void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c || !out_ms || reps < 1) return fail(HN_ERR_ARG, "bad argument");
    if (!c->problem_set) return fail(HN_ERR_STATE, "no solve state");
    *out_ms = 0.f;
#ifndef HN_EMU
    cudaStream_t st = (cudaStream_t)stream;
    const Weights& W = c->W;
    const int B = c->batch, cur = c->cur;
    cudaEvent_t e0, e1;
    HN_CUDA(cudaEventCreate(&e0));
    HN_CUDA(cudaEventCreate(&e1));
    int rc = HN_OK;
    for (int i = -1; i < reps && rc == HN_OK; i++) {   // i = -1: untimed warm-up launch
        if (i == 0) cudaEventRecord(e0, st);
        switch (which) {
            case 0: {
                Conv3Args a = conv_args(c, W.inc[1], c->mid[0], nullptr, c->x[0], c->r[0], S_X + 0, S_IMID);
                rc = launch_conv3<SRC_A8, 8, false, EPI_STORE>(c, a, B, st);
            } break;
            case 1: {
                Conv3Args a = conv_args(c, W.dec[0][0], c->upo[0], c->skip[0], c->mid[0], c->r[0], S_DMID + 0, S_UPO + 0, S_SKIP + 0);
                rc = launch_conv3<SRC_A8_B8, 8, true, EPI_STORE>(c, a, B, st);
            } break;
            case 2: rc = launch_down(c, 0, B, st); break;
            case 3: rc = launch_up(c, 0, B, st); break;
            case 4:
            case 5: {
                const int n = c->n;
                const bool fast256 = (n == 256 && c->pml <= 16 && c->spec_fast);
                const bool fast512 = (n == 512 && c->pml <= 16 && c->spec_fast);
                if (which == 4) {
                    const int total_rows = B * n;
                    if (n == 1024 && c->pml <= 16 && c->spec_fast)
                        s1024::spectral_rows1024_kernel<<<dim3((total_rows + s1024::ROWS_LINES - 1) / s1024::ROWS_LINES),
                                                          dim3(s1024::ROWS_THREADS), s1024::ROWS_SMEM_BYTES, st>>>(
                            c->spec, reinterpret_cast<const float2*>(c->wf), reinterpret_cast<float2*>(c->rx), total_rows);
                    else if (fast512)
                        s512::spectral_rows512_kernel<<<dim3((total_rows + s512::LINES - 1) / s512::LINES), dim3(s512::THREADS),
                                                        s512::ROWS_SMEM_BYTES, st>>>(
                            c->spec, reinterpret_cast<const float2*>(c->wf), reinterpret_cast<float2*>(c->rx), total_rows);
                    else if (fast256)
                        s256::spectral_rows256_kernel<<<dim3((total_rows + s256::LINES - 1) / s256::LINES), dim3(s256::THREADS), 0, st>>>(
                            c->spec, reinterpret_cast<const float2*>(c->wf), reinterpret_cast<float2*>(c->rx), total_rows);
                    else
                        spectral_rows_kernel<<<dim3((total_rows + c->rows_L - 1) / c->rows_L), dim3(SPEC_THREADS),
                                               spectral_smem_bytes(n, c->rows_L, c->pml), st>>>(
                            c->spec, reinterpret_cast<const float2*>(c->wf), reinterpret_cast<float2*>(c->rx), total_rows, c->rows_L);
                } else {
                    ColsArgs a;
                    a.u = reinterpret_cast<const float2*>(c->wf);
                    a.rx = reinterpret_cast<const float2*>(c->rx);
                    a.ksq = c->ksq;
                    a.src = reinterpret_cast<const float2*>(c->src);
                    a.src_nz = c->src_skip ? c->src_nz : nullptr;
                    a.res = reinterpret_cast<float2*>(c->tmp2b);
                    a.ssq = nullptr;
                    a.slot = c->iter_dev + 1;
                    a.amax_out = nullptr;
                    a.src_batch = c->src_batch;
                    a.B = B;
                    a.b0 = 0;
                    a.CW = c->cols_CW;
                    if (fast256) s256::spectral_cols256_kernel<<<dim3(n / s256::LINES, B), dim3(s256::THREADS), s256::COLS_SMEM_BYTES, st>>>(c->spec, a);
                    else if (fast512)
                        s512::spectral_cols512_kernel<<<dim3(n / s512::LINES, B), dim3(s512::THREADS), s512::COLS_SMEM_BYTES, st>>>(c->spec, a);
                    else if (n == 1024 && c->pml <= 16 && c->spec_fast)
                        s1024::spectral_cols1024_kernel<<<dim3(n / s1024::COLS, B), dim3(s1024::COLS_THREADS), s1024::COLS_SMEM_BYTES, st>>>(
                            c->spec, a);
                    else
                        spectral_cols_kernel<<<dim3((n + a.CW - 1) / a.CW, B), dim3(SPEC_THREADS), spectral_smem_bytes(n, a.CW, c->pml), st>>>(
                            c->spec, a);
                }
            } break;
#ifdef HN_HAVE_TC
            // fused DoubleConv kernels of level 0 (engine 2); outputs go to the buffers the iteration itself overwrites
            case 6: {
                int f = launch_dconv<SRC_INC, EPI_STORE>(c, W.inc, c->wf, c->res, c->x[0], c->r[0], S_X + 0, S_WF + cur, S_RES + cur, B, st);
                rc = f == 1 ? HN_OK : (f < 0 ? f : fail(HN_ERR_STATE, "fused DoubleConv kernel not available for this level/engine"));
            } break;
            case 7: {
                int f = launch_dconv<SRC_A8_B2, EPI_STORE>(c, W.sig[0], c->x[0], c->state[0][cur], c->skip[0], c->r[0], S_SKIP + 0, S_X + 0,
                                                           S_STATE + cur, B, st);
                rc = f == 1 ? HN_OK : (f < 0 ? f : fail(HN_ERR_STATE, "fused DoubleConv kernel not available for this level/engine"));
            } break;
            case 8: {
                int f = launch_dconv<SRC_A8_B2, EPI_STORE2>(c, W.sta[0], c->skip[0], c->state[0][cur], c->state[0][cur ^ 1], c->r[0],
                                                            S_STATE + (cur ^ 1), S_SKIP + 0, S_STATE + cur, B, st);
                rc = f == 1 ? HN_OK : (f < 0 ? f : fail(HN_ERR_STATE, "fused DoubleConv kernel not available for this level/engine"));
            } break;
            case 9: {   // decode[0] + outc + wavefield update, exactly the launch of the iteration (reads and rewrites wf: the solve
                        // state is disturbed, callers reset afterwards)
                int f = launch_dconv<SRC_A8_B8, EPI_OUTC>(c, W.dec[0], c->upo[0], c->skip[0], c->dec[0], c->r[0], -1, S_UPO + 0, S_SKIP + 0, B, st,
                                                          &W.outc, c->wf, nullptr);
                rc = f == 1 ? HN_OK : (f < 0 ? f : fail(HN_ERR_STATE, "fused DoubleConv kernel not available for this level/engine"));
            } break;
#endif
            default: rc = fail(HN_ERR_ARG, "unknown kernel id");
        }
    }
    (void)cur;
    if (rc == HN_OK) {
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *out_ms = ms / (float)reps;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    HN_TRY(rc);
    HN_CUDA(cudaGetLastError());
#endif
    return HN_OK;
}

int hn_profile_iteration(hn_ctx* c, float out_ms[2], void* stream) {
    DeviceGuard dev_guard(c ? c->device : -1);
    if (!c || !out_ms) return fail(HN_ERR_ARG, "NULL argument");
    if (!c->problem_set) return fail(HN_ERR_STATE, "no solve state");
    out_ms[0] = out_ms[1] = 0.f;
#ifndef HN_EMU
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1, e2;
    HN_CUDA(cudaEventCreate(&e0));
    HN_CUDA(cudaEventCreate(&e1));
    HN_CUDA(cudaEventCreate(&e2));
    HN_CUDA(cudaMemsetAsync(c->ssq1, 0, (size_t)c->batch * sizeof(double), st));
    HN_CUDA(cudaEventRecord(e0, st));
    int rc = launch_unet(c, c->batch, st, false, false);
    if (rc == HN_OK) {
        cudaEventRecord(e1, st);
        rc = launch_spectral(c, c->batch, st, c->wf, c->ksq, c->src, c->src_batch, c->res, c->ssq1, c->iter_dev + 1,
                             c->amax + S_RES + (c->cur ^ 1));
    }
    if (rc == HN_OK) {
        cudaEventRecord(e2, st);
        cudaEventSynchronize(e2);
        cudaEventElapsedTime(&out_ms[0], e0, e1);
        cudaEventElapsedTime(&out_ms[1], e1, e2);
        c->cur ^= 1;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    HN_TRY(rc);
    HN_CUDA(cudaGetLastError());
#endif
    return HN_OK;
}

}  // extern "C"
