// tc_bench.cu -- development microbenchmark (not part of the product): cycles per tcgen05.mma for the tiny-N
// shapes the conv kernels use, chained vs independent accumulators, and tcgen05.ld drain rate.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)));
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// mode: N of the MMA; nacc: number of accumulators rotated over; nmma: MMAs per measurement; kchunks: K=16 fixed
__global__ void bench(int N, int nacc, int nmma, int shiftmode, long long* out, int nld, int commit_every) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t bar2;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tslot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm) + 40 * 1024;
        uint64_t da[9], db = make_desc(b0, 128, 256);
        for (int t = 0; t < 9; t++) da[t] = make_desc(a0 + (shiftmode ? t * 16 : 0), 17536, 128);
        t0 = clock64();
        if (commit_every > 0) {
            for (int i = 0; i < nmma; i += 3) {
#pragma unroll
                for (int t = 0; t < 3; t++) mma(tb, da[t], db, idesc, 1);
                if ((i / 3) % commit_every == 0) commit(&bar2);
            }
        } else if (nacc == 1) {
            for (int i = 0; i < nmma; i += 9) {
#pragma unroll
                for (int t = 0; t < 9; t++) mma(tb, da[t], db, idesc, 1);
            }
        } else {
            for (int i = 0; i < nmma; i += 36) {
#pragma unroll
                for (int t = 0; t < 9; t++) {
#pragma unroll
                    for (int a = 0; a < 4; a++) mma(tb + a * N, da[t], db, idesc, 1);
                }
            }
        }
        commit(&bar);
        wait(&bar, 0);
        t1 = clock64();
        out[blockIdx.x * 4 + 0] = t1 - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // tcgen05.ld drain: each of the first 4 warps reads nld x16 tiles
    if (warp < 4) {
        long long s0 = clock64();
        uint32_t sum = 0;
        for (int i = 0; i < nld; i++) {
            uint32_t v[16];
            const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)((i * 16) % 512);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            sum += v[0] ^ v[15];
        }
        long long s1 = clock64();
        if ((tid & 31) == 0) out[blockIdx.x * 4 + 1] = (s1 - s0) + (sum == 0x12345 ? 1 : 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * 4 * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    const int nmma = 1152;
    struct { int N, nacc, shift, grid, ce; } cfg[] = {{16, 1, 1, 1, 1}, {16, 1, 1, 1, 2}, {16, 1, 1, 1, 4}, {16, 1, 1, 1, 16}, {48, 1, 1, 1, 1}, {48, 1, 1, 1, 4},{16, 1, 0, 1}, {16, 1, 1, 1}, {16, 4, 1, 1}, {8, 1, 1, 1}, {8, 4, 1, 1}, {32, 1, 1, 1}, {32, 4, 1, 1},
                                                   {64, 4, 0, 1}, {128, 1, 0, 1}, {128, 4, 0, 1}, {256, 1, 0, 1}, {16, 4, 1, 148}, {16, 4, 1, 296}, {16, 4, 1, 592}};
    for (auto& c : cfg) {
        cudaMemset(d, 0, 148 * 32);
        bench<<<c.grid, 128, 48 * 1024>>>(c.N, c.nacc, nmma, c.shift, d, 256, c.ce);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("N=%3d nacc=%d shift=%d grid=%3d commit_every_3mma_x%d : %s  %.1f cyc/MMA   ld.x16: %.1f cyc/ld (per warp, 4 warps)\n", c.N, c.nacc, c.shift, c.grid, c.ce,
               cudaGetErrorString(e), (double)h[0] / nmma, (double)h[1] / 256);
    }
    return 0;
}
