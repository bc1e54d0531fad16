/* cuda_emu.cpp — TEST INFRASTRUCTURE ONLY: fiber scheduler of the lock-step CUDA emulator
 * (see cuda_emu.h). */
#include "cuda_emu.h"

namespace emu {

Cta *g_cta = nullptr;
static const size_t STACK = 256 << 10;
static std::vector<char *> g_stacks; /* fiber stacks are recycled from launch to launch (a CTA of 416 threads would otherwise map and unmap 100 MB each time) */

static void fiber_entry()
{
    Cta &c = *g_cta;
    Fiber *f = c.cur;
    c.body();
    f->done = true;
    recheck();
    swapcontext(&f->ctx, &c.sched);
}

static void run_cta(Cta &c)
{
    g_cta = &c;
    unsigned n = c.block_dim.x;
    c.fibers.assign(n, Fiber());
    c.warps.assign((n + 31) / 32, Warp());
    c.cta_arrived = 0;
    for (unsigned t = 0; t < n; t++) {
        Fiber &f = c.fibers[t];
        f.tid = t;
        if (g_stacks.empty()) f.stack = (char *)malloc(STACK);
        else { f.stack = g_stacks.back(); g_stacks.pop_back(); }
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK;
        f.ctx.uc_link = &c.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    for (;;) {
        bool any_live = false, progressed = false;
        for (unsigned t = 0; t < n; t++) {
            Fiber &f = c.fibers[t];
            if (f.done) continue;
            any_live = true;
            if (f.wait != W_NONE && !f.released) continue;   /* parked */
            c.cur = &f;
            swapcontext(&c.sched, &f.ctx);
            progressed = true;
        }
        if (!any_live) break;
        if (!progressed) {
            fprintf(stderr, "emu: DEADLOCK in block %u — parked threads:\n", c.block_idx.x);
            for (auto &f : c.fibers)
                if (!f.done) fprintf(stderr, "  tid %u waits %s mask %08x\n", f.tid, f.wait == W_CTA ? "__syncthreads" : "warp", f.wait_mask);
            abort();
        }
    }
    for (auto &f : c.fibers) g_stacks.push_back(f.stack);
    g_cta = nullptr;
}

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()> &body)
{
    for (unsigned b = 0; b < grid.x; b++) {
        Cta c;
        c.block_idx = dim3(b, 0, 0);
        c.block_dim = block;
        c.grid_dim = grid;
        c.body = body;
        c.dyn_smem.assign(dyn_smem_bytes + 16, 0xCD);
        run_cta(c);
    }
}

} // namespace emu
