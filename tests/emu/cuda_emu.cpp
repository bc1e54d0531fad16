/* cuda_emu.cpp — TEST INFRASTRUCTURE ONLY: fiber scheduler of the lock-step CUDA emulator
 * (see cuda_emu.h). */
#include "cuda_emu.h"

#ifdef ZSK_EMU_FAST_SWITCH
/* saves rbp, rbx, r12-r15 and the stack pointer of the running fiber, resumes the other one (System V x86-64) */
__asm__(".text\n"
        ".globl zsk_emu_switch\n"
        ".hidden zsk_emu_switch\n"
        ".type zsk_emu_switch,@function\n"
        "zsk_emu_switch:\n"
        "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
        "    movq %rsp, (%rdi)\n"
        "    movq (%rsi), %rsp\n"
        "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
        "    ret\n"
        ".size zsk_emu_switch,.-zsk_emu_switch\n");
#endif

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
    zsk_emu_switch(&f->ctx, &c.sched);
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
#ifdef ZSK_EMU_FAST_SWITCH
        /* first switch: six zeroed registers are popped, then `ret` enters fiber_entry with the stack aligned as after a call */
        uintptr_t top = ((uintptr_t)f.stack + STACK) & ~(uintptr_t)15;
        void **sp = (void **)(top - 8 * sizeof(void *));
        for (int k = 0; k < 6; k++) sp[k] = nullptr;
        sp[6] = (void *)fiber_entry;
        sp[7] = nullptr; /* fiber_entry never returns: it switches back to the scheduler for good */
        f.ctx.sp = sp;
#else
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK;
        f.ctx.uc_link = &c.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
#endif
    }
    for (;;) {
        bool any_live = false, progressed = false;
        for (unsigned t = 0; t < n; t++) {
            Fiber &f = c.fibers[t];
            if (f.done) continue;
            any_live = true;
            if (f.wait != W_NONE && !f.released) continue;   /* parked */
            c.cur = &f;
            zsk_emu_switch(&c.sched, &f.ctx);
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
