/*
 * helmnet_sm100.h -- C ABI of libhelmnet_sm100.so, the B200 (sm_100a) implementation of the
 * helmnet inference inner loop.
 *
 * The reference (ucl-bug/helmnet) has no FFI layer: its boundary is the Python class
 * helmnet.IterativeSolver.  Each entry point below names the reference method it replaces
 * (paths relative to the reference repository).  The Python host side in
 * helmnet_b200/solver.py binds these with ctypes; INTEGRATION.md shows the stub a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative hn_status on failure; nothing throws
 *     across the ABI; hn_last_error() returns a thread-local message for the last failure.
 *   - all `d_*` pointers are DEVICE pointers on the context's device, float32, and use the
 *     reference's own tensor layouts (NCHW, contiguous) unless strides are passed.
 *   - `stream` is a cudaStream_t (NULL = legacy default stream).  Work is enqueued, never
 *     synchronised, unless stated.
 *   - a context is bound to (device, N, max_batch); not thread-safe; one context per GPU when a
 *     batch is sharded over GPUs.
 *   - there is NO CPU fallback: every call fails with HN_ERR_CUDA when no sm_100 device is present.
 */
#ifndef HELMNET_SM100_H
#define HELMNET_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hn_ctx hn_ctx;

typedef enum {
    HN_OK = 0,
    HN_ERR_ARG = -1,     /* bad argument (size not divisible by 16, batch > max_batch, NULL, ...) */
    HN_ERR_CUDA = -2,    /* CUDA runtime error, message in hn_last_error() */
    HN_ERR_STATE = -3,   /* call order violated (weights/source not set, ...) */
    HN_ERR_NOMEM = -4
} hn_status;

/* Number of float32 parameters of HybridNet(features=8, depth=4, state_channels=2, inchannels=6)
 * in state_dict order (helmnet/architectures.py:317-388): 48160. */
#define HN_NUM_WEIGHTS 48160
#define HN_DEPTH 4

/* Library / build identification: "helmnet_sm100 <version> sm_100a". */
const char* hn_version(void);
const char* hn_last_error(void);

/* IterativeSolver.__init__ + set_domain_size + set_laplacian (helmnet/hybridnet.py:20-75, 92-131)
 * and FastLaplacianWithPML.init_variables/get_gamma_functions (helmnet/spectral.py:267-363):
 * builds the k-vector, PML sigma/a/b tables and FFT twiddles for an N x N domain and allocates
 * every per-iteration workspace for up to max_batch samples.  N must be a positive multiple of 16. */
int hn_create(hn_ctx** out, int device, int n, int max_batch, int pml_size, double sigma_max, double k0,
              double omega);
int hn_destroy(hn_ctx* ctx);

/* IterativeSolver.load_state_dict for the `f.*` entries (helmnet/architectures.py:317-388):
 * host_blob holds the HN_NUM_WEIGHTS floats of HybridNet.state_dict() concatenated in
 * state_dict order (each tensor flattened C-contiguously). Synchronous. */
int hn_load_weights(hn_ctx* ctx, const float* host_blob, size_t n_floats);

/* IterativeSolver.set_source_maps (helmnet/hybridnet.py:145-149): source[src_batch, 2, N, N]
 * with element strides (in floats) given explicitly so that the reference's permuted view
 * (strides 2N^2, 1, 2N, 2) is accepted as is.  src_batch is 1 (broadcast) or the batch size. */
int hn_set_source(hn_ctx* ctx, const float* d_src, int src_batch, const int64_t strides[4], void* stream);

/* IterativeSolver.set_multiple_sources / SourceModule.make_abs_spatial_map + spatial_map
 * (helmnet/hybridnet.py:161-170, helmnet/source_module.py:41-116): one monochromatic point source map per location,
 * d_out [count, 2, N, N] (NCHW, contiguous; channel 0 = |map| cos(arg), channel 1 = |map| sin(arg), arg = omega*t + phase).
 * d_locations is int32 [count, 2] = (row, col) on `device`.  smooth != 0 applies the reference's Blackman window in the spatial
 * frequency domain (in closed form: a separable 5-tap kernel around the location, periodic wrap).  Needs no context. */
int hn_point_sources(int device, int n, int count, const int32_t* d_locations, double amplitude, double arg, int smooth,
                     float* d_out, void* stream);

/* IterativeSolver.get_initials + HybridNet.clear_states + the initial get_residual
 * (helmnet/hybridnet.py:522-538, 670-673; helmnet/architectures.py:415-417):
 * k_sq = (omega/sos)^2, wavefield = 0, hidden states = 0, residual = L(0) + k_sq*0 - source.
 * d_sos is [batch, 1, N, N]. */
int hn_reset(hn_ctx* ctx, const float* d_sos, int batch, void* stream);

/* Entry state of IterativeSolver.n_steps / single_step (helmnet/hybridnet.py:558-623):
 * caller supplied wavefield [B,2,N,N], residual [B,2,N,N], k_sq [B,1,N,N] and flattened hidden
 * state [B,2,sum_d (N/2^d)^2] (HybridNet.flatten_state, helmnet/architectures.py:419-423).
 * Any of d_wf/d_res/d_ksq/d_hflat may be NULL to keep the context's current value. */
int hn_set_state(hn_ctx* ctx, const float* d_wf, const float* d_res, const float* d_ksq, const float* d_hflat,
                 int batch, void* stream);

/* The hot loop of IterativeSolver.forward / n_steps (helmnet/hybridnet.py:677-689, 600-612):
 * n_iters x single_step.  Per iteration: UNet update of wavefield and hidden states, then the
 * spectral Laplacian/PML residual and its per-sample RMSE (test_loss_function, hybridnet.py:295-297).
 *   d_rmse     [n_iters, batch]            or NULL
 *   d_wf_hist  [n_iters, batch, 2, N, N]   or NULL  (return_wavefields=True)
 *   d_res_hist [n_iters, batch, 2, N, N]   or NULL  (the reference's `residuals` list)
 *   d_h_hist   [n_iters, batch, 2, S]      or NULL  (return_states=True)
 * No host synchronisation. */
int hn_run(hn_ctx* ctx, int n_iters, float* d_rmse, float* d_wf_hist, float* d_res_hist, float* d_h_hist,
           void* stream);

/* Backward pass of ONE IterativeSolver.single_step (helmnet/hybridnet.py:558-584) for the training unroll
 * IterativeSolver.n_steps under autograd (helmnet/hybridnet.py:586-623, used by training_step :385-410): what
 * loss.backward() does through one step of the reference's graph (37 convolutions, 14 PReLUs, the 1/1e3 update and the
 * spectral residual).  Inputs of the step as given to hn_set_state: d_wf, d_res [B,2,N,N], d_ksq [B,1,N,N], d_hflat [B,2,S].
 * Upstream gradients of the step's outputs (any may be NULL = zero): d_g_wf, d_g_res [B,2,N,N] (new wavefield / residual),
 * d_g_hflat [B,2,S] (new hidden states).  Results: gradients of the inputs d_gwf_in, d_gres_in [B,2,N,N], d_ghflat_in [B,2,S]
 * (any may be NULL) and the parameter gradients ADDED to d_gparams [HN_NUM_WEIGHTS] (state_dict order, as hn_load_weights).
 * The step is recomputed in fp32 on the CUDA cores with every pre-activation kept; k_sq and the source get no gradient. */
int hn_step_backward(hn_ctx* ctx, const float* d_wf, const float* d_res, const float* d_ksq, const float* d_hflat,
                     const float* d_g_wf, const float* d_g_res, const float* d_g_hflat, float* d_gwf_in, float* d_gres_in,
                     float* d_ghflat_in, float* d_gparams, int batch, void* stream);

/* Read back current wavefield / residual / flattened hidden state (NCHW); any may be NULL. */
int hn_get(hn_ctx* ctx, float* d_wf, float* d_res, float* d_hflat, void* stream);

/* HybridNet.get_states(flatten=True) (helmnet/architectures.py:406-413) for the first `batch` samples:
 * d_hflat [batch, 2, S].  Unlike hn_get it needs no solve state (used after hn_unet). */
int hn_get_states(hn_ctx* ctx, float* d_hflat, int batch, void* stream);

/* IterativeSolver.get_residual (helmnet/hybridnet.py:544-556) for a caller supplied field:
 * d_out = L(d_x) + d_ksq * d_x - source.  d_x, d_out [batch,2,N,N]; d_ksq [batch,1,N,N] (NULL: use the
 * context's k_sq).  d_rmse [batch] or NULL.  Does not touch the solver state. */
int hn_residual(hn_ctx* ctx, const float* d_x, const float* d_ksq, float* d_out, float* d_rmse, int batch,
                void* stream);

/* IterativeSolver.apply_laplacian / FastLaplacianWithPML.forward
 * (helmnet/hybridnet.py:540-542, helmnet/spectral.py:251-262, 31-79): d_out = L(d_x), both [batch,2,N,N]. */
int hn_laplacian(hn_ctx* ctx, const float* d_x, float* d_out, int batch, void* stream);

/* HybridNet.forward (helmnet/architectures.py:439-465): d_in [batch,6,N,N] -> d_out [batch,2,N,N];
 * reads and updates the context's hidden states exactly like the module mutates enc[d].state. */
int hn_unet(hn_ctx* ctx, const float* d_in, float* d_out, int batch, void* stream);

/* Introspection used by tests and bench.py. */
int hn_state_len(const hn_ctx* ctx);           /* S = sum_d (N/2^d)^2 */
int64_t hn_launch_count(const hn_ctx* ctx);    /* kernels launched (graph nodes counted per replay) since create */
int hn_kernels_per_iteration(const hn_ctx* ctx);
/* Copies an internal activation (NHWC on device) to d_out as NCHW [batch,C,r,r]; name is e.g.
 * "x0", "skip1", "up2", "dec1", "bot", "x3". Returns channel count (>0) or a negative status. */
int hn_debug_tensor(hn_ctx* ctx, const char* name, float* d_out, int batch, void* stream);
/* Selects the convolution engine: 0 = fp32 CUDA-core kernels, 1 = tcgen05 split-fp16 kernels (one per conv) for the
 * layers that have one, 2 (default) = the same with each DoubleConv (architectures.py:63-84) of a 128- or 256-pixel
 * wide level fused into one kernel.  Returns the engine now in effect or a negative status. */
int hn_set_engine(hn_ctx* ctx, int engine);
/* Synchronises `stream` and reports device-side faults recorded by the kernels (tcgen05 completion
 * watchdog). Returns HN_OK or HN_ERR_CUDA. */
int hn_sync_check(hn_ctx* ctx, void* stream);
/* Average device time in ms of `reps` launches of one kernel of the iteration on the current buffers
 * (which = 0 inc conv #2, 1 decode[0] conv #1 [engine-1 single-conv kernels], 2 enc[0].down, 3 up[0], 4 spectral rows,
 * 5 spectral cols, and the fused DoubleConv kernels of level 0 [engine 2]: 6 inc, 7 enc[0].conv_signal,
 * 8 enc[0].conv_state, 9 decode[0] + outc + wavefield update -- this one rewrites the wavefield: reset afterwards). Synchronous. */
int hn_profile_layer(hn_ctx* ctx, int which, int reps, float* out_ms, void* stream);
/* Per-stage device time of the last hn_profile_iteration() in milliseconds:
 * out[0] = UNet stage, out[1] = spectral residual stage. Synchronous; runs ONE iteration. */
int hn_profile_iteration(hn_ctx* ctx, float out_ms[2], void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HELMNET_SM100_H */
