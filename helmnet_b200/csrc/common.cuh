// common.cuh -- shared definitions for the helmnet sm_100a kernels.
//
// The same kernel sources compile in two modes:
//   * nvcc, -gencode arch=compute_100a,code=sm_100a  -> libhelmnet_sm100.so (the product)
//   * g++ -DHN_EMU (tests/emu/)                       -> a fiber-based functional emulator used ONLY by
//     the CPU test-suite to exercise kernel index logic without a GPU.  It is never loaded by the
//     helmnet_b200 package.
#pragma once

#ifdef HN_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include <stdint.h>

#ifdef HN_EMU
#define HN_LAUNCH(kernel, grid, block, smem, stream, ...) \
    hn_emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#else
#define HN_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// Programmatic dependent launch (PDL): every kernel of a solver iteration is launched with the programmatic stream
// serialization attribute (when hn_ctx::pdl is set), so its CTAs may become resident while the previous kernel of the
// stream / graph is still draining.  Each such kernel calls pdl_wait() before its first access to anything an earlier
// kernel wrote (or still reads) -- only its own prologue (mbarrier init, TMEM allocation, constant weight / table loads
// into shared memory) runs ahead -- and pdl_trigger() right after, which lets the NEXT kernel's CTAs take the SM slots
// this grid frees.  Without the launch attribute both are no-ops.
#ifdef HN_EMU
#define HN_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, ...) HN_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__)
#else
namespace hn {
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);
}
}  // namespace hn
// (used inside functions that return an hn status: a launch error is reported through HN_CUDA)
#define HN_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, ...) \
    HN_CUDA(hn::launch_pdl((pdl), kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__))
#endif

// dynamic shared memory of the running CTA
#ifdef HN_EMU
#define HN_DYN_SMEM(type, name) \
    type* name = reinterpret_cast<type*>((reinterpret_cast<uintptr_t>(hn_emu::g.dyn_smem) + 1023) & ~uintptr_t(1023))
#else
#define HN_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

namespace hn {

constexpr int kDepth = 4;         // encoder levels with hidden state (ckpt: depth=4, state_depth=4)
constexpr int kFeat = 8;          // feature channels

// ---- packed fp32 FMA (Blackwell FFMA2: two fp32 FMAs per issue slot) --------------------------------
// d.{x,y} += a * b.{x,y}.  ptxas folds the {a,a} pack into FFMA2's scalar-broadcast operand form.
__device__ __forceinline__ void ffma2(float2& d, float a, float2 b) {
#if defined(HN_EMU) || defined(HN_NO_FFMA2)
    d.x = fmaf(a, b.x, d.x);
    d.y = fmaf(a, b.y, d.y);
#else
    asm("{ .reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rc; }"
        : "+f"(d.x), "+f"(d.y)
        : "f"(a), "f"(b.x), "f"(b.y));
#endif
}

// see HN_LAUNCH_PDL above
__device__ __forceinline__ void pdl_wait() {
#ifndef HN_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#ifndef HN_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// a 32-bit word another kernel of the same stream may have written while this grid was already resident (PDL): read it
// through L2, never from a (non-coherent) L1 / read-only cache line
__device__ __forceinline__ unsigned ld_fresh(const unsigned* p) {
#ifdef HN_EMU
    return *p;
#else
    return __ldcg(p);
#endif
}

// Balanced strips of the persistent row-streaming kernels (conv_tcf / conv_tcr_down / conv_tcr_up).  The rows of all `batch`
// images (H rows each) are laid end to end, each image preceded by `pad` virtual rows that stand for the cost of starting a strip
// (recomputed halo rows + pipeline fill), and CTA c takes the c-th of gridDim.x equal chunks of that line: every CTA carries the
// same rows + pad * strips whatever the ratio of images to CTA slots (a chunk that crosses an image boundary gets `pad` rows of
// work less).  Returns strip i of this CTA (image b, first row y0, R rows; all even when H and pad are) or false when it has none.
__device__ __forceinline__ bool balanced_strip(int batch, int H, int pad, int i, int& b, int& y0, int& R) {
    const int HV = H + pad;
    const long long VT = (long long)batch * HV;
    const int v0 = (int)((VT * blockIdx.x / gridDim.x) & ~1ll);
    const int v1 = (blockIdx.x + 1 == gridDim.x) ? (int)VT : (int)((VT * (blockIdx.x + 1) / gridDim.x) & ~1ll);
    int found = -1;
    for (int bb = v0 / HV; bb * HV < v1; bb++) {
        const int ys = max(0, v0 - bb * HV - pad), ye = min(H, v1 - bb * HV - pad);
        if (ye > ys && ++found == i) {
            b = bb;
            y0 = ys;
            R = ye - ys;
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// One NHWC8 pixel (32 bytes, 32-byte aligned) as a single 256-bit store: every store instruction fills whole 32-byte
// sectors (two 128-bit stores at a 32-byte thread stride each write half sectors, twice the L2 write requests).
__device__ __forceinline__ void st_nhwc8(float* dst, const float (&o)[8]) {
#ifdef HN_EMU
    reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
#else
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]),
                 "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Running max |x| of a tensor, kept as the bit pattern of a non-negative float (unsigned order == float order).
// Producers publish it so that the tcgen05 convolutions can pick a power-of-two block scale for their fp16
// operand split without a second pass over the data.  Call with the whole warp; lane 0 publishes.
__device__ __forceinline__ void publish_amax(unsigned* slot, float local_max) {
    if (slot == nullptr) return;
    const float m = warp_max(local_max);
    if ((threadIdx.x & 31) == 0) {
        const unsigned bits = __float_as_uint(m);
        if (bits > *reinterpret_cast<volatile unsigned*>(slot)) atomicMax(slot, bits);
    }
}

}  // namespace hn
