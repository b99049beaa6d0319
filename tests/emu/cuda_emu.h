// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
// A tiny functional emulator of the CUDA execution model (grid of CTAs, threads as ucontext fibers,
// __syncthreads, warp shuffles, dynamic/static shared memory, double atomicAdd) so that the *unmodified*
// kernel sources under helmnet_b200/csrc can be compiled with g++ and their index logic exercised by the
// CPU test-suite (tests/test_emu_*.py) against the oracle.  It is built into tests/emu/libhelmnet_emu.so,
// which the helmnet_b200 package never loads: the product has no CPU path.
//
// Shared memory is filled with NaNs before each CTA starts so reads of unwritten smem show up in the
// results; fibers run in ascending order in one pass and descending order in the next (HN_EMU_REVERSE=1)
// so that most missing-barrier bugs change the answer.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
using std::max;
using std::min;

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { return posix_memalign(p, 256, n ? n : 16) == 0 ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }

namespace hn_emu {
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    int wait = 0;  // 0 running, 1 block barrier, 2 warp barrier
    dim3 tid;
};
struct State {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    int current = -1;
    const std::function<void()>* fn = nullptr;
    char* dyn_smem = nullptr;
    uint32_t shfl_buf[2][64][32];  // [phase][warp][lane]
    int shfl_phase[64];
};
extern State g;
extern dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn);
void yield_wait(int kind);
uint32_t shfl_exchange(uint32_t v, int src_lane_of_self_fn_kind, int arg);
}  // namespace hn_emu

#define threadIdx hn_emu::g_threadIdx
#define blockIdx hn_emu::g_blockIdx
#define blockDim hn_emu::g_blockDim
#define gridDim hn_emu::g_gridDim

static inline void __syncthreads() { hn_emu::yield_wait(1); }
static inline void __syncwarp() { hn_emu::shfl_exchange(0u, 0, 0); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline float __shfl_xor_sync(unsigned, float v, int lanemask) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u = hn_emu::shfl_exchange(u, 0, lanemask);
    memcpy(&v, &u, 4);
    return v;
}
