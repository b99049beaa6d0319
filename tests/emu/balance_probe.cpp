// balance_probe.cpp -- TEST INFRASTRUCTURE ONLY: exposes common.cuh's balanced_strip() (the strip enumeration of the persistent
// tcgen05 kernels, which the emulator cannot run) so that the CPU suite can check its partition for any (batch, H, pad, grid).
#define HN_EMU 1
#include "../../helmnet_b200/csrc/common.cuh"

namespace {
__global__ void balance_probe_kernel(int batch, int H, int pad, int* out, int cap_per_cta) {
    // out[cta][i] = {b, y0, R} for the strips of this CTA, terminated by b = -1
    int* o = out + (size_t)blockIdx.x * cap_per_cta * 3;
    int i = 0, b, y0, R;
    for (; i < cap_per_cta - 1 && hn::balanced_strip(batch, H, pad, i, b, y0, R); i++) {
        o[3 * i] = b;
        o[3 * i + 1] = y0;
        o[3 * i + 2] = R;
    }
    o[3 * i] = -1;
}
}  // namespace

extern "C" int emu_balanced_strips(int batch, int H, int pad, int grid, int* out, int cap_per_cta) {
    HN_LAUNCH(balance_probe_kernel, dim3(grid), dim3(1), 0, nullptr, batch, H, pad, out, cap_per_cta);
    return 0;
}
