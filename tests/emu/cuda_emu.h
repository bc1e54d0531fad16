/*
 * cuda_emu.h — TEST INFRASTRUCTURE ONLY: a tiny lock-step CUDA execution emulator.
 *
 * There is no GPU in the development container, and a gpurun round-trip costs minutes.  To debug the
 * LOGIC of the sm_100a kernels (format handling, warp-collective protocols, barrier placement) before
 * spending GPU time, tests/emu compiles the very same kernel source (libzseek_b200/csrc/zsk_kernels.cuh)
 * with g++ and runs each CUDA thread as a fiber (own stack, cooperative switch):
 *   - one CTA at a time, all of its threads as cooperatively scheduled fibers;
 *   - __syncthreads / __syncwarp / __shfl*_sync / __ballot_sync / __any/__all / __reduce_*_sync are
 *     rendezvous points: a fiber yields until every live lane named in the mask has arrived;
 *   - a rendezvous that can never complete (divergent collective, missing barrier) aborts with a
 *     diagnostic instead of hanging.
 * It is NOT a product path and NOT a CPU fallback: nothing under libzseek_b200/ includes it, the
 * shipped library contains no host decode code, and the emulated kernels are only ever compared
 * against the oracle by tests (tests/test_emu_kernels.py).  It cannot find memory-model races; those
 * are checked on the GPU with compute-sanitizer.
 */
#ifndef ZSK_CUDA_EMU_H
#define ZSK_CUDA_EMU_H

#include <ucontext.h>

/* Context switch between fibers.  swapcontext() saves and restores the signal mask with a system call on every switch
 * (a fifth of the emulator's run time); on x86-64 a switch only has to exchange the callee-saved registers and the stack
 * pointer.  Other targets keep ucontext. */
#if defined(__x86_64__) && !defined(ZSK_EMU_UCONTEXT)
#define ZSK_EMU_FAST_SWITCH 1
struct zsk_emu_ctx { void *sp; };
extern "C" void zsk_emu_switch(zsk_emu_ctx *from, zsk_emu_ctx *to);
#else
typedef ucontext_t zsk_emu_ctx;
static inline void zsk_emu_switch(zsk_emu_ctx *from, zsk_emu_ctx *to) { swapcontext(from, to); }
#endif
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

namespace emu {

enum WaitKind { W_NONE = 0, W_CTA, W_WARP };

struct Fiber {
    zsk_emu_ctx ctx;
    char *stack = nullptr;
    unsigned tid = 0;
    bool done = false;
    WaitKind wait = W_NONE;
    unsigned wait_mask = 0;
    bool released = false;
};

struct Warp {
    uint64_t xch[32];       /* exchange slots for shuffles / votes */
    unsigned arrived = 0;   /* lanes currently parked at (or just arrived to) a warp rendezvous */
};

struct Cta {
    dim3 block_idx, block_dim, grid_dim;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    unsigned cta_arrived = 0;
    int or_acc = 0, or_result = 0;   /* __syncthreads_or */
    zsk_emu_ctx sched;
    Fiber *cur = nullptr;
    std::function<void()> body;
    std::vector<uint8_t> dyn_smem;
};

extern Cta *g_cta;

static inline Fiber &self() { return *g_cta->cur; }
static inline unsigned lane() { return self().tid & 31; }
static inline Warp &mywarp() { return g_cta->warps[self().tid >> 5]; }

static inline unsigned live_mask(unsigned warp)
{
    unsigned m = 0;
    for (unsigned l = 0; l < 32; l++) {
        unsigned t = warp * 32 + l;
        if (t < g_cta->fibers.size() && !g_cta->fibers[t].done) m |= 1u << l;
    }
    return m;
}

static inline void yield_to_sched() { zsk_emu_switch(&self().ctx, &g_cta->sched); }

/* Releases every parked rendezvous of the CTA that has become complete (called on arrival and
 * whenever a thread exits). */
static inline void recheck()
{
    Cta &c = *g_cta;
    for (unsigned wi = 0; wi < c.warps.size(); wi++) {
        Warp &w = c.warps[wi];
        unsigned live = live_mask(wi);
        for (unsigned l = 0; l < 32; l++) {
            unsigned t = wi * 32 + l;
            if (t >= c.fibers.size()) break;
            Fiber &f = c.fibers[t];
            if (f.done || f.wait != W_WARP || f.released) continue;
            unsigned need = f.wait_mask & live;
            if ((w.arrived & need) != need) continue;
            bool same = true;          /* every needed lane must be parked on the same mask */
            for (unsigned k = 0; k < 32; k++)
                if (need >> k & 1) { Fiber &g = c.fibers[wi * 32 + k]; if (g.wait != W_WARP || g.wait_mask != f.wait_mask) same = false; }
            if (!same) continue;
            for (unsigned k = 0; k < 32; k++)
                if (need >> k & 1) c.fibers[wi * 32 + k].released = true;
            w.arrived &= ~need;
        }
    }
    unsigned live_n = 0;
    for (auto &x : c.fibers) live_n += !x.done;
    if (live_n && c.cta_arrived == live_n) {
        for (auto &x : c.fibers) if (!x.done && x.wait == W_CTA) x.released = true;
        c.cta_arrived = 0;
        c.or_result = c.or_acc;
        c.or_acc = 0;
    }
}

/* Warp rendezvous over `mask`: returns once all live lanes in mask have arrived. */
static inline void warp_rendezvous(unsigned mask)
{
    Fiber &f = self();
    Warp &w = mywarp();
    if (!(mask & (1u << (f.tid & 31)))) { fprintf(stderr, "emu: lane %u not in its own sync mask %08x\n", f.tid, mask); abort(); }
    w.arrived |= 1u << (f.tid & 31);
    f.wait = W_WARP; f.wait_mask = mask; f.released = false;
    recheck();
    while (!f.released) yield_to_sched();
    f.wait = W_NONE;
}

static inline void cta_rendezvous()
{
    Fiber &f = self();
    g_cta->cta_arrived++;
    f.wait = W_CTA; f.released = false;
    recheck();
    while (!f.released) yield_to_sched();
    f.wait = W_NONE;
}

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()> &body);

} // namespace emu

/* ---- built-in variables */
#define threadIdx (emu::ThreadIdxProxy{})
#define blockIdx (emu::g_cta->block_idx)
#define blockDim (emu::g_cta->block_dim)
#define gridDim (emu::g_cta->grid_dim)
namespace emu { struct ThreadIdxProxy { struct X { operator unsigned() const { return emu::self().tid; } } x; unsigned y = 0, z = 0; }; }

/* ---- barriers */
static inline void __syncthreads() { emu::cta_rendezvous(); }
static inline int __syncthreads_or(int pred) { if (pred) emu::g_cta->or_acc = 1; emu::cta_rendezvous(); return emu::g_cta->or_result; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_rendezvous(mask); }
static inline void __threadfence_block() {}
static inline void __threadfence() {}
static inline unsigned __activemask() { return emu::live_mask(emu::self().tid >> 5); }

/* ---- warp collectives: deposit, rendezvous, read, rendezvous */
template <typename T> static inline T emu_xchg_read(unsigned mask, T v, int src_lane, bool valid_src)
{
    static_assert(sizeof(T) <= 8, "emu shuffle supports <= 8 byte types");
    emu::Warp &w = emu::mywarp();
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    w.xch[emu::lane()] = raw;
    emu::warp_rendezvous(mask);
    T out = v;
    if (valid_src && (mask >> src_lane & 1)) { uint64_t r = w.xch[src_lane]; memcpy(&out, &r, sizeof(T)); }
    emu::warp_rendezvous(mask);
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{ int l = (int)emu::lane(); int s = (l & ~(width - 1)) | (src & (width - 1)); return emu_xchg_read(mask, v, s, true); }
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{ int l = (int)emu::lane(); int s = l - (int)d; bool ok = s >= (l & ~(width - 1)); return emu_xchg_read(mask, v, ok ? s : l, ok); }
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{ int l = (int)emu::lane(); int s = l + (int)d; bool ok = s <= (l | (width - 1)); return emu_xchg_read(mask, v, ok ? s : l, ok); }
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{ (void)width; int l = (int)emu::lane(); return emu_xchg_read(mask, v, l ^ x, true); }

static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    emu::Warp &w = emu::mywarp();
    w.xch[emu::lane()] = pred ? 1 : 0;
    emu::warp_rendezvous(mask);
    unsigned live = emu::live_mask(emu::self().tid >> 5) & mask, r = 0;
    for (unsigned l = 0; l < 32; l++) if ((live >> l & 1) && w.xch[l]) r |= 1u << l;
    emu::warp_rendezvous(mask);
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { unsigned live = emu::live_mask(emu::self().tid >> 5) & mask; return __ballot_sync(mask, pred) == live; }

template <typename Op> static inline unsigned emu_reduce(unsigned mask, unsigned v, Op op)
{
    emu::Warp &w = emu::mywarp();
    w.xch[emu::lane()] = v;
    emu::warp_rendezvous(mask);
    unsigned live = emu::live_mask(emu::self().tid >> 5) & mask;
    bool first = true; unsigned r = 0;
    for (unsigned l = 0; l < 32; l++) if (live >> l & 1) { unsigned x = (unsigned)w.xch[l]; r = first ? x : op(r, x); first = false; }
    emu::warp_rendezvous(mask);
    return r;
}
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a + b; }); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a | b; }); }
static inline unsigned __reduce_and_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a & b; }); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }

/* ---- integer intrinsics */
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; s = s > 32 ? 32 : s; return (unsigned)(v >> s); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) { unsigned s = (sel >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * s)) & 0xff) << (8 * i); }
    return r;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }

/* ---- atomics (fibers are cooperative: plain read-modify-write is atomic) */
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

static inline uint8_t *zsk_emu_dyn_smem() { return emu::g_cta->dyn_smem.data(); }

#endif /* ZSK_CUDA_EMU_H */
