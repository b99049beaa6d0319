// tc_probe.cu -- development probe (not part of the product): validates the tcgen05 encodings the conv
// kernels rely on, in isolation, against an exact integer reference computed on the host:
//   * kind::f16 MMA, M=128 N=16 K=16, fp32 accumulate in TMEM
//   * K-major SWIZZLE_NONE smem descriptors: 8 rows x 16 B core matrices, SBO = 128 B between 8-row groups,
//     LBO = distance between the two K chunks (here: between the "hi" and "lo" planes)
//   * A start addresses that are only 16-byte aligned (tap shifts of a padded linear pixel layout)
//   * accumulate chain over 9 taps, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b.x16
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define POS 256          // positions per plane
#define PITCH 34

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (sm_100)
    return d;                // layout_type = 0 (SWIZZLE_NONE), base_offset = 0, lbo_mode = 0
}

__global__ void probe(const __half* a_hi, const __half* a_lo, const __half* bmat, float* out, int* status) {
    __shared__ __align__(128) uint4 plane_hi[POS];
    __shared__ __align__(128) uint4 plane_lo[POS];
    __shared__ __align__(128) uint4 bsm[9 * 32];      // 9 taps x 512 B
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < POS; i += blockDim.x) {
        plane_hi[i] = reinterpret_cast<const uint4*>(a_hi)[i];
        plane_lo[i] = reinterpret_cast<const uint4*>(a_lo)[i];
    }
    for (int i = tid; i < 9 * 32; i += blockDim.x) bsm[i] = reinterpret_cast<const uint4*>(bmat)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");   // generic-proxy smem writes -> visible to the tensor core
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t lbo_a = smem_u32(plane_lo) - smem_u32(plane_hi);
        for (int t = 0; t < 9; t++) {
            const int shift = (t / 3) * PITCH + (t % 3);
            const uint64_t da = make_desc(smem_u32(plane_hi) + shift * 16, lbo_a, 128);
            const uint64_t db = make_desc(smem_u32(bsm) + t * 512, 128, 256);
            const uint32_t acc = t > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
    }
    // wait (bounded) for the MMAs
    {
        uint32_t done = 0;
        for (int it = 0; it < (1 << 22) && !done; it++) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u));
        }
        if (!done) { if (tid == 0) *status = 1; }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t v[16];
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int n = 0; n < 16; n++) out[(warp * 32 + lane) * 16 + n] = __uint_as_float(v[n]);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tbase));
}

int main() {
    std::vector<__half> hi(POS * 8), lo(POS * 8), b(9 * 256);
    std::vector<float> hif(POS * 8), lof(POS * 8), bf(9 * 256, 0.f);
    srand(1);
    for (int i = 0; i < POS * 8; i++) {
        hif[i] = (float)(rand() % 17 - 8);
        lof[i] = (float)(rand() % 9 - 4);
        hi[i] = __float2half(hif[i]);
        lo[i] = __float2half(lof[i]);
    }
    // logical B_t[n][k], n<16, k<16, stored as: (n/8)*256 + (k/8)*128 + (n%8)*16 + (k%8)*2 bytes
    std::vector<float> blog(9 * 16 * 16);
    for (int t = 0; t < 9; t++)
        for (int n = 0; n < 16; n++)
            for (int k = 0; k < 16; k++) {
                float val = (float)(rand() % 7 - 3);
                blog[(t * 16 + n) * 16 + k] = val;
                int byte = (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
                b[t * 256 + byte / 2] = __float2half(val);
            }
    __half *dhi, *dlo, *db;
    float* dout;
    int* dstat;
    cudaMalloc(&dhi, POS * 16); cudaMalloc(&dlo, POS * 16); cudaMalloc(&db, 9 * 512); cudaMalloc(&dout, 128 * 16 * 4); cudaMalloc(&dstat, 4);
    cudaMemcpy(dhi, hi.data(), POS * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(dlo, lo.data(), POS * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), 9 * 512, cudaMemcpyHostToDevice);
    cudaMemset(dstat, 0, 4);
    cudaMemset(dout, 0, 128 * 16 * 4);
    probe<<<1, 128>>>(dhi, dlo, db, dout, dstat);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> out(128 * 16);
    int stat = -1;
    cudaMemcpy(out.data(), dout, 128 * 16 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&stat, dstat, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 16; n++) {
            double ref = 0;
            for (int t = 0; t < 9; t++) {
                int p = m + (t / 3) * PITCH + (t % 3);
                for (int k = 0; k < 16; k++) {
                    float a = k < 8 ? hif[p * 8 + k] : lof[p * 8 + k - 8];
                    ref += (double)a * blog[(t * 16 + n) * 16 + k];
                }
            }
            double err = fabs(ref - out[m * 16 + n]);
            if (err > maxerr) maxerr = err;
            if (err > 1e-3 && bad < 8) { printf("mismatch m=%d n=%d got %f want %f\n", m, n, out[m * 16 + n], ref); bad++; }
        }
    printf("status=%d maxerr=%g  %s\n", stat, maxerr, (maxerr < 1e-3 && stat == 0) ? "PROBE_OK" : "PROBE_FAIL");
    return 0;
}
