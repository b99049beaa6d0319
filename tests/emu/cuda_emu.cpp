// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

namespace hn_emu {
State g;
dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
static const size_t kStack = 256 * 1024;

static void fiber_main() {
    (*g.fn)();
    g.fibers[g.current].done = true;
    swapcontext(&g.fibers[g.current].ctx, &g.sched);
}

void yield_wait(int kind) {
    Fiber& f = g.fibers[g.current];
    f.wait = kind;
    swapcontext(&f.ctx, &g.sched);
}

uint32_t shfl_exchange(uint32_t v, int, int lanemask) {
    const int t = g.current, w = t >> 5, lane = t & 31;
    const int ph = g.shfl_phase[w];
    g.shfl_buf[ph][w][lane] = v;
    yield_wait(2);
    // phase is flipped by the scheduler when the warp is released
    return g.shfl_buf[ph][w][(lane ^ lanemask) & 31];
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const bool reverse = getenv("HN_EMU_REVERSE") && getenv("HN_EMU_REVERSE")[0] == '1';
    g.fn = &fn;
    g_blockDim = block;
    g_gridDim = grid;
    if ((int)g.fibers.size() < nthreads) {
        size_t old = g.fibers.size();
        g.fibers.resize(nthreads);
        for (size_t i = old; i < g.fibers.size(); i++) g.fibers[i].stack = (char*)malloc(kStack);
    }
    char* dyn = nullptr;
    if (posix_memalign((void**)&dyn, 1024, smem + 1024) != 0) abort();
    g.dyn_smem = dyn;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx = dim3(bx, by, bz);
                memset(dyn, 0xFF, smem + 1024);  // NaN pattern
                memset(g.shfl_phase, 0, sizeof(g.shfl_phase));
                for (int t = 0; t < nthreads; t++) {
                    Fiber& f = g.fibers[t];
                    f.done = false;
                    f.wait = 0;
                    f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = kStack;
                    f.ctx.uc_link = &g.sched;
                    makecontext(&f.ctx, fiber_main, 0);
                }
                int alive = nthreads;
                while (alive > 0) {
                    bool progressed = false;
                    for (int k = 0; k < nthreads; k++) {
                        const int t = reverse ? nthreads - 1 - k : k;
                        Fiber& f = g.fibers[t];
                        if (f.done || f.wait != 0) continue;
                        g.current = t;
                        g_threadIdx = f.tid;
                        swapcontext(&g.sched, &f.ctx);
                        progressed = true;
                        if (f.done) alive--;
                    }
                    // release warps whose live lanes all wait on a shuffle
                    for (int w = 0; w * 32 < nthreads; w++) {
                        bool all = true, any = false;
                        for (int t = w * 32; t < std::min(nthreads, w * 32 + 32); t++) {
                            Fiber& f = g.fibers[t];
                            if (f.done) continue;
                            if (f.wait == 2) any = true;
                            else all = false;
                        }
                        if (any && all) {
                            for (int t = w * 32; t < std::min(nthreads, w * 32 + 32); t++)
                                if (!g.fibers[t].done) g.fibers[t].wait = 0;
                            g.shfl_phase[w] ^= 1;
                            progressed = true;
                        }
                    }
                    // release the block barrier when every live thread waits on it
                    bool all = alive > 0, any = false;
                    for (int t = 0; t < nthreads; t++) {
                        Fiber& f = g.fibers[t];
                        if (f.done) continue;
                        if (f.wait == 1) any = true;
                        else all = false;
                    }
                    if (any && all) {
                        for (int t = 0; t < nthreads; t++) g.fibers[t].wait = 0;
                        progressed = true;
                    }
                    if (!progressed && alive > 0) {
                        fprintf(stderr, "hn_emu: deadlock (divergent barrier) in block (%u,%u,%u)\n", bx, by, bz);
                        abort();
                    }
                }
            }
    free(dyn);
    g.dyn_smem = nullptr;
}
}  // namespace hn_emu
