// tc_contend.cu -- development microbenchmark (not part of the product): does background shared-memory traffic (LDS/STS
// from other warps) or TMEM traffic (tcgen05.ld/st) slow down a stream of small-N tcgen05.mma?  One issuing thread runs
// M=128, K=16, N=48 MMAs back to back (dx-shifted A descriptors as in conv_tcf.cuh); `bg` selects what 16 more warps do.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc));
}
__global__ void __launch_bounds__(544) bench(int N, int nmma, int bg, int bgwarps, long long* out, volatile int* stop_g) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tslot;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (tid == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm) + 40 * 1024;
            uint64_t da[3], db = make_desc(b0, 128, 256);
            for (int t = 0; t < 3; t++) da[t] = make_desc(a0 + t * 16, 264 * 16, 128);
            long long t0 = clock64();
            for (int i = 0; i < nmma; i += 3) {
#pragma unroll
                for (int t = 0; t < 3; t++) mma(tb + (uint32_t)((i / 3) % 4) * 48, da[t], db, idesc);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            long long t1 = clock64();
            out[0] = t1 - t0;
            stop = 1;
        }
    } else if (warp <= bgwarps) {
        long long cnt = 0;
        if (bg == 1 || bg == 2 || bg == 3) {
            // LDS.128 and/or STS.128 streaming over a 64 KB region well away from the MMA operands (conflict-free)
            uint4* base = reinterpret_cast<uint4*>(sm + 64 * 1024);
            uint4 acc = make_uint4(0, 0, 0, 0);
            int idx = tid & 1023;
            while (!stop) {
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    if (bg == 1 || bg == 3) { uint4 v = base[(idx + r * 512) & 4095]; acc.x ^= v.x; acc.y += v.y; }
                    if (bg == 2 || bg == 3) base[(idx + r * 512 + 256) & 4095] = acc;
                }
                cnt += 8;
            }
            if (acc.x == 0x1234567) out[3] = acc.y;
        } else if (bg == 4) {
            // TMEM traffic: ld.x16 x2 + st.x32 on columns 256.. of this warp's lane quadrant
            const uint32_t taddr = tb + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)(((warp >> 2) & 3) * 32);
            uint32_t sum = 0;
            while (!stop) {
                uint32_t v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                sum += v[0] ^ v[15];
                const uint32_t z = sum & 1u;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z));
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                cnt++;
            }
            if (sum == 0x1234567) out[3] = sum;
        } else if (bg == 5) {
            // pure ALU/FMA load (issue-slot pressure, no memory)
            float a = (float)tid, b = 1.0001f;
            while (!stop) {
#pragma unroll
                for (int r = 0; r < 32; r++) a = fmaf(a, b, 0.5f);
                cnt += 32;
            }
            if (a == 0.12345f) out[3] = 1;
        }
        if (lane == 0 && warp == 1) out[1] = cnt;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int nmma = 3072;
    const char* names[] = {"none", "LDS.128", "STS.128", "LDS+STS", "TMEM ld/st", "FFMA"};
    for (int bg = 0; bg < 6; bg++)
        for (int w : {4, 8, 16}) {
            if (bg == 0 && w != 4) continue;
            cudaMemset(d, 0, 64);
            bench<<<1, 544, 160 * 1024>>>(48, nmma, bg, w, d, nullptr);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[4];
            cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
            printf("bg=%-10s warps=%2d : %s  %.1f cyc/MMA   bg ops per warp-thread: %lld (%.2f per MMA-cycle)\n", names[bg], w, cudaGetErrorString(e),
                   (double)h[0] / nmma, h[1], (double)h[1] / (double)h[0]);
        }
    return 0;
}
