/*
 * zsk_common.cuh — shared device-side definitions for the sm_100a read-path kernels.
 *
 * Compiled two ways:
 *   - by nvcc for sm_100a into libzseek_b200.so (the product);
 *   - by g++ with -DZSK_EMU against tests/emu/cuda_emu.h, which runs every CUDA thread as a fiber so the
 *     kernel logic can be checked against the oracle without a GPU (test infrastructure only; the
 *     product contains no host decode path).
 */
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "zsk_abi.h"

#ifdef ZSK_EMU
#include "cuda_emu.h"
#define ZSK_LDG(p) (*(p))
#else
#include <cuda_runtime.h>
#define ZSK_LDG(p) __ldg(p)
#endif

#define ZSK_FULL 0xffffffffu

static __device__ __forceinline__ uint32_t zsk_rd16(const uint8_t *p) { return (uint32_t)ZSK_LDG(p) | ((uint32_t)ZSK_LDG(p + 1) << 8); }
static __device__ __forceinline__ uint32_t zsk_rd24(const uint8_t *p) { return zsk_rd16(p) | ((uint32_t)ZSK_LDG(p + 2) << 16); }
static __device__ __forceinline__ uint32_t zsk_rd32(const uint8_t *p) { return zsk_rd16(p) | (zsk_rd16(p + 2) << 16); }
static __device__ __forceinline__ uint64_t zsk_rd64(const uint8_t *p) { return (uint64_t)zsk_rd32(p) | ((uint64_t)zsk_rd32(p + 4) << 32); }

/* Aligned-word reads of an arbitrarily aligned byte stream: 32 bits starting at byte address p. */
static __device__ __forceinline__ uint32_t zsk_ld32_unaligned(const uint8_t *p)
{
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t lo = ZSK_LDG(w);
    if (sh == 0) return lo;
    uint32_t hi = ZSK_LDG(w + 1);
    return __funnelshift_r(lo, hi, sh);
}

static __device__ __forceinline__ uint64_t zsk_ld64_unaligned(const uint8_t *p)
{
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t w0 = ZSK_LDG(w), w1 = ZSK_LDG(w + 1);
    if (sh == 0) return (uint64_t)w0 | ((uint64_t)w1 << 32);
    uint32_t w2 = ZSK_LDG(w + 2);
    return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
}

/* ------------------------------------------------------------------------------------------
 * Group copies.  `rank`/`size` describe the cooperating thread group (a warp: lane/32; a CTA:
 * threadIdx.x/blockDim.x).  Destination head is aligned to 16 B, the body moves 16 B per thread
 * per step with the source re-aligned by funnel shifts, the tail goes bytewise.
 * ------------------------------------------------------------------------------------------ */
static __device__ __forceinline__ void zsk_group_copy(uint8_t *dst, const uint8_t *src, uint32_t n,
                                                      unsigned rank, unsigned size)
{
    if (n < 64) {
        for (uint32_t i = rank; i < n; i += size) dst[i] = src[i];
        return;
    }
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    for (uint32_t i = rank; i < head; i += size) dst[i] = src[i];
    uint32_t body = (n - head) & ~15u;
    const uint8_t *s = src + head;
    uint8_t *d = dst + head;
    unsigned sh = (unsigned)((uintptr_t)s & 3) * 8;
    const uint32_t *sw = (const uint32_t *)((uintptr_t)s & ~(uintptr_t)3);
    for (uint32_t i = rank * 16; i < body; i += size * 16) {
        const uint32_t *w = sw + (i >> 2);
        uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        uint4 v;
        if (sh == 0) {
            v = make_uint4(w0, w1, w2, w3);
        } else {
            uint32_t w4 = w[4];
            v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                           __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
        *(uint4 *)(d + i) = v;
    }
    for (uint32_t i = head + body + rank; i < n; i += size) dst[i] = src[i];
}

static __device__ __forceinline__ void zsk_group_fill(uint8_t *dst, uint8_t v, uint32_t n, unsigned rank, unsigned size)
{
    if (n < 64) {
        for (uint32_t i = rank; i < n; i += size) dst[i] = v;
        return;
    }
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    for (uint32_t i = rank; i < head; i += size) dst[i] = v;
    uint32_t body = (n - head) & ~15u;
    uint32_t w = 0x01010101u * v;
    uint4 vv = make_uint4(w, w, w, w);
    for (uint32_t i = rank * 16; i < body; i += size * 16) *(uint4 *)(dst + head + i) = vv;
    for (uint32_t i = head + body + rank; i < n; i += size) dst[i] = v;
}

/* ------------------------------------------------------------------------------------------
 * LZ77 match copy by a lane group of G lanes (G = 32: a warp, G = 8: a quarter warp):
 * out[op .. op+ml) = out[op-off ..), byte-serial semantics when the regions overlap (off < ml
 * replicates the pattern).  All lanes of the group call with identical arguments; gl = lane index
 * inside the group, gmask = the group's lanes.  The caller has made earlier stores visible
 * (__syncwarp(gmask)) and must sync again afterwards.
 * ------------------------------------------------------------------------------------------ */
template <unsigned G>
static __device__ __forceinline__ void zsk_group_match(uint8_t *out, uint32_t op, uint32_t off, uint32_t ml, unsigned gl, unsigned gmask)
{
    if (off >= G || off >= ml) {
        /* every G-byte slice reads only bytes written before the slice started */
        const bool overlap = off < ml;
        for (uint32_t base = 0; base < ml; base += G) {
            uint32_t i = base + gl;
            if (i < ml) out[op + i] = out[op + i - off];
            if (overlap) __syncwarp(gmask);
        }
    } else {
        /* short period: every byte is a copy of one of the `off` bytes before op */
        const uint8_t *pat = out + op - off;
        uint32_t r = gl % off;
        const uint32_t step = G % off;
        for (uint32_t i = gl; i < ml; i += G) {
            out[op + i] = pat[r];
            r += step;
            if (r >= off) r -= off;
        }
    }
}

static __device__ __forceinline__ void zsk_warp_match(uint8_t *out, uint32_t op, uint32_t off, uint32_t ml, unsigned lane)
{
    zsk_group_match<32>(out, op, off, ml, lane, ZSK_FULL);
}
