// train_host.cuh -- host side of hn_step_backward: the backward pass of ONE IterativeSolver.single_step
// (helmnet/hybridnet.py:558-584) for the training unroll n_steps under autograd (hybridnet.py:586-623, 385-410).
// Included by helmnet_sm100.cu (needs hn_ctx, HN_LAUNCH, grid1d).  Kernels: train.cuh.
#pragma once

// offsets (in floats) of every parameter tensor in the state_dict-order blob of hn_load_weights / the gradient blob
struct RawDC { size_t w0, b0, sl, w1, b1; };
struct RawOff {
    RawDC inc, sig[kDepth], sta[kDepth], dec[kDepth], bot;
    size_t downw[kDepth], downb[kDepth], upw[kDepth], upb[kDepth], outw, outb, total;
};
static RawOff raw_offsets() {
    RawOff o;
    size_t p = 0;
    auto take = [&](size_t nfl) { const size_t q = p; p += nfl; return q; };
    auto dc = [&](RawDC& d, int cin, int cmid, int cout) {
        d.w0 = take((size_t)cmid * cin * 9); d.b0 = take(cmid); d.sl = take(1); d.w1 = take((size_t)cout * cmid * 9); d.b1 = take(cout);
    };
    dc(o.inc, 6, 8, 8);
    for (int d = 0; d < kDepth; d++) {     // module order inside EncoderBlock: conv_signal, down, conv_state
        dc(o.sig[d], 10, 8, 8);
        o.downw[d] = take(8 * 8 * 64); o.downb[d] = take(8);
        dc(o.sta[d], 10, 2, 2);
    }
    for (int d = 0; d < kDepth; d++) dc(o.dec[d], 16, 8, 8);
    dc(o.bot, 8, 8, 8);
    for (int d = 0; d < kDepth; d++) { o.upw[d] = take(8 * 8 * 64); o.upb[d] = take(8); }
    o.outw = take(16); o.outb = take(2);
    o.total = p;
    return o;
}

// workspace of the backward pass: the recomputed forward (every conv input and PReLU pre-activation) and the gradients in flight
struct TrainWs {
    int batch = 0;
    float* base = nullptr;
    double* gacc = nullptr;     // [16] PReLU slope gradients of the current backward pass, accumulated in double
    int n_slopes = 0;           // slots of gacc handed out so far in this pass
    int slope_off[16];          // where each slot goes in the gradient blob
    // step inputs (c2)
    float *wf, *res, *h[kDepth];
    // forward
    float *in6, *inc_z, *inc_a, *x[kDepth + 1], *s_z[kDepth], *s_a[kDepth], *skip[kDepth], *t_z[kDepth], *t_a[kDepth];
    float *bot_z, *bot_a, *y[kDepth + 1], *u[kDepth], *d_z[kDepth], *d_a[kDepth];
    // gradients
    float *G, *gadd, *gr, *rxadj, *gin6;
    float *gx[kDepth + 1], *gy[kDepth + 1], *gz[kDepth + 1], *gskip[kDepth], *gu[kDepth], *gh[kDepth], *gz2[kDepth], *ghn[kDepth];
};

static int train_ws(hn_ctx* c, int B, TrainWs** out) {
    TrainWs* w = c->tws;
    if (w != nullptr && w->batch >= B) { *out = w; return HN_OK; }
    if (w != nullptr) {
#ifndef HN_EMU
        cudaDeviceSynchronize();
#endif
        cudaFree(w->base);
        cudaFree(w->gacc);
        delete w;
        c->tws = nullptr;
    }
    w = new TrainWs();
    std::vector<std::pair<float**, size_t>> req;
    auto R = [&](float*& p, size_t nfl) { req.push_back({&p, (nfl + 63) & ~(size_t)63}); };
    const size_t Bz = (size_t)B;
    size_t P[kDepth + 1];
    for (int d = 0; d <= kDepth; d++) P[d] = Bz * c->r[d] * c->r[d];
    R(w->wf, P[0] * 2); R(w->res, P[0] * 2); R(w->in6, P[0] * 6); R(w->inc_z, P[0] * 8); R(w->inc_a, P[0] * 8);
    R(w->G, P[0] * 2); R(w->gadd, P[0] * 2); R(w->gr, P[0] * 2); R(w->rxadj, P[0] * 2); R(w->gin6, P[0] * 6);
    R(w->bot_z, P[kDepth] * 8); R(w->bot_a, P[kDepth] * 8);
    for (int d = 0; d <= kDepth; d++) { R(w->x[d], P[d] * 8); R(w->y[d], P[d] * 8); R(w->gx[d], P[d] * 8); R(w->gy[d], P[d] * 8); R(w->gz[d], P[d] * 8); }
    for (int d = 0; d < kDepth; d++) {
        R(w->h[d], P[d] * 2); R(w->s_z[d], P[d] * 8); R(w->s_a[d], P[d] * 8); R(w->skip[d], P[d] * 8); R(w->t_z[d], P[d] * 2); R(w->t_a[d], P[d] * 2);
        R(w->u[d], P[d] * 8); R(w->d_z[d], P[d] * 8); R(w->d_a[d], P[d] * 8);
        R(w->gskip[d], P[d] * 8); R(w->gu[d], P[d] * 8); R(w->gh[d], P[d] * 2); R(w->gz2[d], P[d] * 2); R(w->ghn[d], P[d] * 2);
    }
    size_t total = 0;
    for (auto& q : req) total += q.second;
    if (cudaMalloc(reinterpret_cast<void**>(&w->base), total * sizeof(float)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&w->gacc), 16 * sizeof(double)) != cudaSuccess) {
        if (w->base) cudaFree(w->base);
        delete w;
        return fail(HN_ERR_NOMEM, "cudaMalloc of the backward workspace (" + std::to_string(total * 4) + " bytes) failed");
    }
    size_t off = 0;
    for (auto& q : req) { *q.first = w->base + off; off += q.second; }
    w->batch = B;
    c->tws = w;
    *out = w;
    return HN_OK;
}

static inline int tgrid(long long items, int threads) {
    long long g = (items + threads - 1) / threads;
    if (g > 148 * 8) g = 148 * 8;
    if (g < 1) g = 1;
    return (int)g;
}

static int t_conv(hn_ctx* c, cudaStream_t st, const tr::ConvArgs& a, int CO) {
    const size_t smem = (size_t)a.ks * a.ks * (a.ca + a.cb) * CO * sizeof(float);
    const dim3 g(tgrid(a.P, tr::T_THREADS)), b(tr::T_THREADS);
    switch (CO) {
        case 2: HN_LAUNCH(tr::conv_kernel<2>, g, b, smem, st, a); break;
        case 6: HN_LAUNCH(tr::conv_kernel<6>, g, b, smem, st, a); break;
        case 8: HN_LAUNCH(tr::conv_kernel<8>, g, b, smem, st, a); break;
        case 10: HN_LAUNCH(tr::conv_kernel<10>, g, b, smem, st, a); break;
        case 16: HN_LAUNCH(tr::conv_kernel<16>, g, b, smem, st, a); break;
        default: return fail(HN_ERR_ARG, "unsupported channel count in the backward pass");
    }
    c->launches++;
    return HN_OK;
}
// forward of one conv layer: z (pre-activation) and, when act != null, act = PReLU(z)
static int t_fwd(hn_ctx* c, cudaStream_t st, const float* a, int ca, const float* b, int cb, const float* w, const float* bias, int co,
                 int ks, int r, int B, float* z, float slope, float* act) {
    tr::ConvArgs p = {};
    p.a = a; p.ca = ca; p.b = b; p.cb = b ? cb : 0; p.w = w; p.bias = bias; p.in_scale = 1.f; p.ks = ks; p.transposed = 0;
    p.H = r; p.W = r; p.P = (long long)B * r * r; p.o0 = z; p.c0 = co; p.acc0 = 0; p.o1 = nullptr; p.acc1 = 0; p.slope = slope; p.act = act;
    return t_conv(c, st, p, co);
}
// data gradient of one conv layer: dz (co channels) -> the layer's input channels [0, c0) into o0 and [c0, cin) into o1
static int t_bwd(hn_ctx* c, cudaStream_t st, const float* dz, int co, const float* w, int cin, int ks, int r, int B, float* o0, int c0, int acc0,
                 float* o1, int acc1, float in_scale = 1.f) {
    tr::ConvArgs p = {};
    p.a = dz; p.ca = co; p.b = nullptr; p.cb = 0; p.w = w; p.bias = nullptr; p.in_scale = in_scale; p.ks = ks; p.transposed = 1;
    p.H = r; p.W = r; p.P = (long long)B * r * r; p.o0 = o0; p.c0 = c0; p.acc0 = acc0; p.o1 = o1; p.acc1 = acc1; p.slope = 0.f; p.act = nullptr;
    return t_conv(c, st, p, cin);
}
static int t_wgrad(hn_ctx* c, cudaStream_t st, const float* a, int ca, const float* b, int cb, const float* dz, int co, int ks, int r, int B,
                   float* gw, float* gb, float dz_scale = 1.f) {
    tr::WgradArgs p = {};
    p.a = a; p.ca = ca; p.b = b; p.cb = b ? cb : 0; p.dz = dz; p.dz_scale = dz_scale; p.ks = ks; p.H = r; p.W = r; p.B = B; p.gw = gw; p.gb = gb;
    const int tiles = (r + tr::WG_T - 1) / tr::WG_T;
    const int total = tiles * tiles * B, grid = total < 2 * c->num_sms ? total : 2 * c->num_sms;
    const size_t smem = tr::wgrad_smem_bytes(p.ca + p.cb, co);
    const int threads = tr::wgrad_threads(p.ca + p.cb, ks);
    if (co == 8) HN_LAUNCH(tr::wgrad_kernel<8>, dim3(grid), dim3(threads), smem, st, p);
    else if (co == 2) HN_LAUNCH(tr::wgrad_kernel<2>, dim3(grid), dim3(threads), smem, st, p);
    else return fail(HN_ERR_ARG, "unsupported channel count in the backward pass");
    c->launches++;
    return HN_OK;
}
static int t_prelu_bwd(hn_ctx* c, cudaStream_t st, const float* z, float* g, float slope, size_t slope_off, size_t total) {
    TrainWs* w = c->tws;
    if (w->n_slopes >= 16) return fail(HN_ERR_STATE, "more PReLU layers than slope slots");
    w->slope_off[w->n_slopes] = (int)slope_off;
    double* gslope = w->gacc + w->n_slopes++;
    HN_LAUNCH(tr::prelu_bwd_kernel, dim3(tgrid((long long)total, 256)), dim3(256), 0, st, z, g, g, slope, gslope, total);
    c->launches++;
    return HN_OK;
}
static int t_s2(hn_ctx* c, cudaStream_t st, bool gather, const float* src, int rs, float* out, int ro, const float* w, const float* bias,
                int acc, int B) {
    tr::S2Args p = {};
    p.src = src; p.Hs = rs; p.Ws = rs; p.out = out; p.Ho = ro; p.Wo = ro; p.w = w; p.bias = bias; p.acc = acc; p.P = (long long)B * ro * ro;
    if (gather) HN_LAUNCH(tr::s2_gather_kernel, dim3(tgrid(p.P, tr::T_THREADS)), dim3(tr::T_THREADS), 4096 * sizeof(float), st, p);
    else HN_LAUNCH(tr::s2_scatter_kernel, dim3(tgrid(p.P, tr::T_THREADS)), dim3(tr::T_THREADS), 4096 * sizeof(float), st, p);
    c->launches++;
    return HN_OK;
}
static int t_s2_wgrad(hn_ctx* c, cudaStream_t st, const float* small_t, const float* big_t, int rs, int B, float* gw, const float* bias_src,
                      int r_bias, float* gb) {
    tr::S2WgradArgs p = {};
    p.small_t = small_t; p.big_t = big_t; p.Hs = rs; p.Ws = rs; p.B = B; p.gw = gw;
    const int tiles = (rs + tr::SG_T - 1) / tr::SG_T;
    const int total = tiles * tiles * B, grid = total < 2 * c->num_sms ? total : 2 * c->num_sms;
    HN_LAUNCH(tr::s2_wgrad_kernel, dim3(grid), dim3(tr::WG_THREADS), tr::s2_wgrad_smem_bytes(), st, p);
    const size_t P = (size_t)B * r_bias * r_bias;
    HN_LAUNCH(tr::chan_sum_kernel, dim3(tgrid((long long)P, 256)), dim3(256), 0, st, bias_src, 8, P, gb);
    c->launches += 2;
    return HN_OK;
}

// DoubleConv (architectures.py:63-84) forward with everything kept / backward.  `wr` = raw weights on the device, `gp` = gradient blob.
static int t_dc_fwd(hn_ctx* c, cudaStream_t st, const float* wr, const float* wh, const RawDC& o, const float* a, int ca, const float* b, int cb,
                    int cmid, int cout, int r, int B, float* z, float* act, float* out) {
    HN_TRY(t_fwd(c, st, a, ca, b, cb, wr + o.w0, wr + o.b0, cmid, 3, r, B, z, wh[o.sl], act));
    if (out != nullptr) HN_TRY(t_fwd(c, st, act, cmid, nullptr, 0, wr + o.w1, wr + o.b1, cout, 3, r, B, out, 0.f, nullptr));
    return HN_OK;
}
// gout: gradient of the DoubleConv output (cout channels); gmid: scratch (cmid channels); input gradient split [0, c0) -> o0, rest -> o1
static int t_dc_bwd(hn_ctx* c, cudaStream_t st, const float* wr, const float* wh, float* gp, const RawDC& o, const float* a, int ca, const float* b,
                    int cb, int cmid, int cout, int r, int B, const float* z, const float* act, const float* gout, float* gmid, float* o0, int c0,
                    int acc0, float* o1, int acc1) {
    const int cin = ca + (b ? cb : 0);
    HN_TRY(t_bwd(c, st, gout, cout, wr + o.w1, cmid, 3, r, B, gmid, cmid, 0, nullptr, 0));
    HN_TRY(t_wgrad(c, st, act, cmid, nullptr, 0, gout, cout, 3, r, B, gp + o.w1, gp + o.b1));
    HN_TRY(t_prelu_bwd(c, st, z, gmid, wh[o.sl], o.sl, (size_t)B * r * r * cmid));
    HN_TRY(t_bwd(c, st, gmid, cmid, wr + o.w0, cin, 3, r, B, o0, c0, acc0, o1, acc1));
    HN_TRY(t_wgrad(c, st, a, ca, b, cb, gmid, cmid, 3, r, B, gp + o.w0, gp + o.b0));
    return HN_OK;
}

static int launch_spectral_adjoint(hn_ctx* c, cudaStream_t st, int B, const float* g, const float* ksq, const float* add, float* rx, float* out) {
    const int n = c->n;
    int L = 2048 / n;
    if (L < 1) L = 1;
    if (L > 16) L = 16;
    while (L > 1 && spectral_smem_bytes(n, L, c->pml) > 200 * 1024) L >>= 1;
    const size_t smem = spectral_smem_bytes(n, L, c->pml);
#ifndef HN_EMU
    static size_t attr_smem[16] = {0};     // per function AND device; keep the running maximum (contexts of different sizes)
    if (smem > attr_smem[c->device & 15]) {
        HN_CUDA(cudaFuncSetAttribute(tr::spectral_rows_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HN_CUDA(cudaFuncSetAttribute(tr::spectral_cols_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem[c->device & 15] = smem;
    }
#endif
    const int total_rows = B * n;
    HN_LAUNCH(tr::spectral_rows_adj_kernel, dim3((total_rows + L - 1) / L), dim3(SPEC_THREADS), smem, st, c->spec,
              reinterpret_cast<const float2*>(g), reinterpret_cast<float2*>(rx), total_rows, L);
    HN_LAUNCH(tr::spectral_cols_adj_kernel, dim3((n + L - 1) / L, B), dim3(SPEC_THREADS), smem, st, c->spec, reinterpret_cast<const float2*>(g),
              reinterpret_cast<const float2*>(rx), ksq, reinterpret_cast<const float2*>(add), reinterpret_cast<float2*>(out), L);
    c->launches += 2;
    return HN_OK;
}

static int step_backward(hn_ctx* c, const float* d_wf, const float* d_res, const float* d_ksq, const float* d_hflat, const float* d_g_wf,
                         const float* d_g_res, const float* d_g_hflat, float* d_gwf_in, float* d_gres_in, float* d_ghflat_in, float* gp,
                         int B, cudaStream_t st) {
    static const RawOff O = raw_offsets();
    TrainWs* w = nullptr;
    HN_TRY(train_ws(c, B, &w));
    const float* wr = c->wraw;
    const float* wh = c->wraw_host.data();
    const int n = c->n, hw = n * n;
    const size_t total = (size_t)B * hw;
    const bool have_hgrad = d_g_hflat != nullptr;
    w->n_slopes = 0;
    HN_CUDA(cudaMemsetAsync(w->gacc, 0, 16 * sizeof(double), st));
    // ---- step inputs into the kernels' layouts
    HN_LAUNCH(nchw2_to_c2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_wf, reinterpret_cast<float2*>(w->wf), hw, total, (unsigned*)nullptr);
    HN_LAUNCH(nchw2_to_c2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_res, reinterpret_cast<float2*>(w->res), hw, total, (unsigned*)nullptr);
    c->launches += 2;
    {
        size_t off = 0;
        for (int d = 0; d < kDepth; d++) {
            const int p = c->r[d] * c->r[d];
            const size_t tot = (size_t)B * p;
            HN_LAUNCH(nchw2_strided_to_c2_kernel, dim3(grid1d(tot)), dim3(LAY_THREADS), 0, st, d_hflat + off, reinterpret_cast<float2*>(w->h[d]), p,
                      tot, (size_t)2 * c->state_len, (size_t)c->state_len, (unsigned*)nullptr);
            if (have_hgrad)
                HN_LAUNCH(nchw2_strided_to_c2_kernel, dim3(grid1d(tot)), dim3(LAY_THREADS), 0, st, d_g_hflat + off,
                          reinterpret_cast<float2*>(w->ghn[d]), p, tot, (size_t)2 * c->state_len, (size_t)c->state_len, (unsigned*)nullptr);
            c->launches += have_hgrad ? 2 : 1;
            off += p;
        }
    }
    // ---- forward of the UNet with every conv input and pre-activation kept (architectures.py:439-465)
    HN_LAUNCH(tr::make_in6_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(w->wf),
              reinterpret_cast<const float2*>(w->res), c->sigma1d, w->in6, n, total);
    c->launches++;
    HN_TRY(t_dc_fwd(c, st, wr, wh, O.inc, w->in6, 6, nullptr, 0, 8, 8, n, B, w->inc_z, w->inc_a, w->x[0]));
    for (int d = 0; d < kDepth; d++) {
        const int r = c->r[d];
        HN_TRY(t_dc_fwd(c, st, wr, wh, O.sig[d], w->x[d], 8, w->h[d], 2, 8, 8, r, B, w->s_z[d], w->s_a[d], w->skip[d]));
        // conv_state: only its first conv is needed again (its output is the next step's hidden state, nothing here reads it)
        if (have_hgrad) HN_TRY(t_dc_fwd(c, st, wr, wh, O.sta[d], w->skip[d], 8, w->h[d], 2, 2, 2, r, B, w->t_z[d], w->t_a[d], nullptr));
        HN_TRY(t_s2(c, st, true, w->skip[d], r, w->x[d + 1], c->r[d + 1], wr + O.downw[d], wr + O.downb[d], 0, B));
    }
    HN_TRY(t_dc_fwd(c, st, wr, wh, O.bot, w->x[kDepth], 8, nullptr, 0, 8, 8, c->r[kDepth], B, w->bot_z, w->bot_a, w->y[kDepth]));
    for (int d = kDepth - 1; d >= 0; d--) {
        const int r = c->r[d];
        HN_TRY(t_s2(c, st, false, w->y[d + 1], c->r[d + 1], w->u[d], r, wr + O.upw[d], wr + O.upb[d], 0, B));
        HN_TRY(t_dc_fwd(c, st, wr, wh, O.dec[d], w->u[d], 8, w->skip[d], 8, 8, 8, r, B, w->d_z[d], w->d_a[d], w->y[d]));
    }
    // ---- backward.  wf' = wf + outc(y0) / 1e3,  r' = L wf' + k_sq wf' - source  (hybridnet.py:570-584):
    //      G = dLoss/dwf' = g_wf' + L^H g_r' + k_sq g_r'
    const float* gadd = nullptr;
    if (d_g_wf != nullptr) {
        HN_LAUNCH(nchw2_to_c2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_g_wf, reinterpret_cast<float2*>(w->gadd), hw, total, (unsigned*)nullptr);
        c->launches++;
        gadd = w->gadd;
    }
    if (d_g_res != nullptr) {
        HN_LAUNCH(nchw2_to_c2_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, d_g_res, reinterpret_cast<float2*>(w->gr), hw, total, (unsigned*)nullptr);
        c->launches++;
        HN_TRY(launch_spectral_adjoint(c, st, B, w->gr, d_ksq, gadd, w->rxadj, w->G));
    } else if (gadd != nullptr) {
        HN_CUDA(cudaMemcpyAsync(w->G, gadd, total * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
        HN_CUDA(cudaMemsetAsync(w->G, 0, total * 2 * sizeof(float), st));
    }
    // outc (1 x 1, 8 -> 2; architectures.py:47-60) with the 1/1e3 of the update folded in
    HN_TRY(t_bwd(c, st, w->G, 2, wr + O.outw, 8, 1, n, B, w->gy[0], 8, 0, nullptr, 0, 1e-3f));
    HN_TRY(t_wgrad(c, st, w->y[0], 8, nullptr, 0, w->G, 2, 1, n, B, gp + O.outw, gp + O.outb, 1e-3f));
    for (int d = 0; d < kDepth; d++) {       // decoder, backwards
        const int r = c->r[d];
        HN_TRY(t_dc_bwd(c, st, wr, wh, gp, O.dec[d], w->u[d], 8, w->skip[d], 8, 8, 8, r, B, w->d_z[d], w->d_a[d], w->gy[d], w->gz[d], w->gu[d], 8, 0,
                        w->gskip[d], 0));
        // up[d] = ConvTranspose2d(8, 8, 8, 2, 3) (architectures.py:373-385)
        HN_TRY(t_s2(c, st, true, w->gu[d], r, w->gy[d + 1], c->r[d + 1], wr + O.upw[d], nullptr, 0, B));
        HN_TRY(t_s2_wgrad(c, st, w->y[d + 1], w->gu[d], c->r[d + 1], B, gp + O.upw[d], w->gu[d], r, gp + O.upb[d]));
    }
    HN_TRY(t_dc_bwd(c, st, wr, wh, gp, O.bot, w->x[kDepth], 8, nullptr, 0, 8, 8, c->r[kDepth], B, w->bot_z, w->bot_a, w->gy[kDepth], w->gz[kDepth],
                    w->gx[kDepth], 8, 0, nullptr, 0));
    for (int d = kDepth - 1; d >= 0; d--) {  // encoder, backwards (architectures.py:240-252)
        const int r = c->r[d];
        HN_TRY(t_s2(c, st, false, w->gx[d + 1], c->r[d + 1], w->gskip[d], r, wr + O.downw[d], nullptr, 1, B));
        HN_TRY(t_s2_wgrad(c, st, w->gx[d + 1], w->skip[d], c->r[d + 1], B, gp + O.downw[d], w->gx[d + 1], c->r[d + 1], gp + O.downb[d]));
        if (have_hgrad)       // new state = conv_state(cat[output, old state])
            HN_TRY(t_dc_bwd(c, st, wr, wh, gp, O.sta[d], w->skip[d], 8, w->h[d], 2, 2, 2, r, B, w->t_z[d], w->t_a[d], w->ghn[d], w->gz2[d], w->gskip[d], 8,
                            1, w->gh[d], 0));
        HN_TRY(t_dc_bwd(c, st, wr, wh, gp, O.sig[d], w->x[d], 8, w->h[d], 2, 8, 8, r, B, w->s_z[d], w->s_a[d], w->gskip[d], w->gz[d], w->gx[d], 8, 0,
                        w->gh[d], have_hgrad ? 1 : 0));
    }
    HN_TRY(t_dc_bwd(c, st, wr, wh, gp, O.inc, w->in6, 6, nullptr, 0, 8, 8, n, B, w->inc_z, w->inc_a, w->gx[0], w->gz[0], w->gin6, 6, 0, nullptr, 0));
    {
        tr::SlopeFold f;
        f.n = w->n_slopes;
        for (int i = 0; i < 16; i++) f.off[i] = i < w->n_slopes ? w->slope_off[i] : 0;
        HN_LAUNCH(tr::fold_slopes_kernel, dim3(1), dim3(32), 0, st, w->gacc, gp, f);
        c->launches++;
    }
    // ---- gradients of the step inputs, NCHW
    if (d_gwf_in != nullptr || d_gres_in != nullptr) {
        HN_LAUNCH(tr::step_input_grads_kernel, dim3(grid1d(total)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(w->G), w->gin6, d_gwf_in,
                  d_gres_in, hw, total);
        c->launches++;
    }
    if (d_ghflat_in != nullptr) {
        size_t off = 0;
        for (int d = 0; d < kDepth; d++) {
            const int p = c->r[d] * c->r[d];
            const size_t tot = (size_t)B * p;
            HN_LAUNCH(c2_to_nchw2_kernel, dim3(grid1d(tot)), dim3(LAY_THREADS), 0, st, reinterpret_cast<const float2*>(w->gh[d]), d_ghflat_in + off, p,
                      tot, (size_t)2 * c->state_len, (size_t)c->state_len);
            c->launches++;
            off += p;
        }
    }
    HN_CUDA(cudaGetLastError());
    return HN_OK;
}
