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
#define ZSK_STCG(p, v) (*(p) = (v))
#define ZSK_PREFETCH_L1(p) ((void)(p))
#define ZSK_PREFETCH_L2(p) ((void)(p))
#else
#include <cuda_runtime.h>
#define ZSK_LDG(p) __ldg(p)
/* store that is kept out of L1 (scratch streams written once and read by a later kernel must not evict the lines the
 * read streams of the same SM live in) */
#define ZSK_STCG(p, v) __stcg((p), (v))
#define ZSK_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#define ZSK_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#endif

#define ZSK_FULL 0xffffffffu

/* jobs of this launch: the host's count, or what an earlier kernel of the stream left in device memory */
static __device__ __forceinline__ uint32_t zsk_njobs(const zsk_decode_args &a)
{
    if (!a.njobs_dev) return a.njobs;
    const uint32_t n = *a.njobs_dev;
    return n > a.job_base ? min(a.njobs, n - a.job_base) : 0u;
}

/* ---- cp.async (LDGSTS): 16-byte global -> shared copies that occupy no registers and stall nobody */
#ifdef ZSK_EMU
#include <vector>
/* The emulator models the WORST legal timing: the source is read when the copy is issued, the destination is
 * written only when a wait proves the group complete. */
struct zsk_emu_cp { void *dst; uint8_t data[16]; unsigned group; };
struct zsk_emu_cpq { std::vector<zsk_emu_cp> pend; unsigned committed = 0; };
static zsk_emu_cpq zsk_emu_cpqs[1024];
static inline void zsk_cp16(void *sdst, const void *gsrc)
{
    zsk_emu_cpq &q = zsk_emu_cpqs[(unsigned)threadIdx.x];
    zsk_emu_cp c;
    c.dst = sdst;
    memcpy(c.data, gsrc, 16);
    c.group = q.committed;
    q.pend.push_back(c);
}
static inline void zsk_cp16_cg(void *sdst, const void *gsrc) { zsk_cp16(sdst, gsrc); }
static inline void zsk_cp_commit() { zsk_emu_cpqs[(unsigned)threadIdx.x].committed++; }
template <unsigned N> static inline void zsk_cp_wait()
{
    zsk_emu_cpq &q = zsk_emu_cpqs[(unsigned)threadIdx.x];
    size_t k = 0;
    for (auto &c : q.pend) {
        if (c.group + N < q.committed) memcpy(c.dst, c.data, 16);
        else q.pend[k++] = c;
    }
    q.pend.resize(k);
}
static inline void zsk_cp_reset() { zsk_emu_cpqs[(unsigned)threadIdx.x] = zsk_emu_cpq(); }
#else
static __device__ __forceinline__ void zsk_cp16(void *sdst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
/* same, cached in L2 only: streams that are read once must not evict what lives in L1 */
static __device__ __forceinline__ void zsk_cp16_cg(void *sdst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
static __device__ __forceinline__ void zsk_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <unsigned N> static __device__ __forceinline__ void zsk_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
static __device__ __forceinline__ void zsk_cp_reset() {}
#endif

/* eight words to a 32-byte aligned address as ONE 256-bit store (STG.E.256, sm_100); _cg: kept out of L1 */
#ifdef ZSK_EMU
static inline void zsk_st256_cg(uint32_t *p, const uint32_t v[8]) { for (int k = 0; k < 8; k++) p[k] = v[k]; }
static inline void zsk_st256(uint32_t *p, const uint32_t v[8]) { for (int k = 0; k < 8; k++) p[k] = v[k]; }
#else
static __device__ __forceinline__ void zsk_st256(uint32_t *p, const uint32_t v[8])
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                 "r"(v[6]), "r"(v[7])
                 : "memory");
}
static __device__ __forceinline__ void zsk_st256_cg(uint32_t *p, const uint32_t v[8])
{
    asm volatile("st.global.cg.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                 "r"(v[6]), "r"(v[7])
                 : "memory");
}
#endif



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

/* ------------------------------------------------------------------------------------------
 * XXH32 / XXH64 (the checksums of the LZ4 frame format and of zstd frames; public xxHash
 * specification).  liblz4 verifies the frame-header, block and content checksums and libzstd the
 * content checksum whenever a frame carries them, so the reference's zseek_pread fails on a
 * mismatch (probe: "ERROR_headerChecksum_invalid", "ERROR_blockChecksum_invalid",
 * "ERROR_contentChecksum_invalid", "Restored data doesn't match checksum"); the kernels do the same.
 * The four accumulators of a hash are independent chains: the group versions give each of four
 * lanes one accumulator (one coalesced 16- or 32-byte stripe per step), the serial versions run on
 * a single lane.  p may have any alignment.
 * ------------------------------------------------------------------------------------------ */
#define ZSK_X32_P1 2654435761u
#define ZSK_X32_P2 2246822519u
#define ZSK_X32_P3 3266489917u
#define ZSK_X32_P4 668265263u
#define ZSK_X32_P5 374761393u
#define ZSK_X64_P1 11400714785074694791ull
#define ZSK_X64_P2 14029467366897019727ull
#define ZSK_X64_P3 1609587929392839161ull
#define ZSK_X64_P4 9650029242287828579ull
#define ZSK_X64_P5 2870177450012600261ull

static __device__ __forceinline__ uint32_t zsk_rotl32(uint32_t x, unsigned r) { return (x << r) | (x >> (32 - r)); }
static __device__ __forceinline__ uint64_t zsk_rotl64(uint64_t x, unsigned r) { return (x << r) | (x >> (64 - r)); }
static __device__ __forceinline__ uint32_t zsk_x32_round(uint32_t acc, uint32_t w) { return zsk_rotl32(acc + w * ZSK_X32_P2, 13) * ZSK_X32_P1; }
static __device__ __forceinline__ uint64_t zsk_x64_round(uint64_t acc, uint64_t w) { return zsk_rotl64(acc + w * ZSK_X64_P2, 31) * ZSK_X64_P1; }
static __device__ __forceinline__ uint32_t zsk_ldb32(const uint8_t *p) /* plain (coherent) loads: p may be data this kernel wrote */
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* tail of XXH32 after the stripes: h already holds the merged accumulators (or seed + P5) */
static __device__ __forceinline__ uint32_t zsk_xxh32_finish(uint32_t h, const uint8_t *p, uint32_t rest, uint32_t len)
{
    h += len;
    while (rest >= 4) { h = zsk_rotl32(h + zsk_ldb32(p) * ZSK_X32_P3, 17) * ZSK_X32_P4; p += 4; rest -= 4; }
    while (rest) { h = zsk_rotl32(h + (uint32_t)(*p) * ZSK_X32_P5, 11) * ZSK_X32_P1; p++; rest--; }
    h ^= h >> 15; h *= ZSK_X32_P2; h ^= h >> 13; h *= ZSK_X32_P3; h ^= h >> 16;
    return h;
}

static __device__ uint32_t zsk_xxh32_serial(const uint8_t *p, uint32_t len)
{
    uint32_t h, done = 0;
    if (len >= 16) {
        uint32_t v1 = ZSK_X32_P1 + ZSK_X32_P2, v2 = ZSK_X32_P2, v3 = 0, v4 = 0u - ZSK_X32_P1;
        for (; done + 16 <= len; done += 16) {
            v1 = zsk_x32_round(v1, zsk_ldb32(p + done));
            v2 = zsk_x32_round(v2, zsk_ldb32(p + done + 4));
            v3 = zsk_x32_round(v3, zsk_ldb32(p + done + 8));
            v4 = zsk_x32_round(v4, zsk_ldb32(p + done + 12));
        }
        h = zsk_rotl32(v1, 1) + zsk_rotl32(v2, 7) + zsk_rotl32(v3, 12) + zsk_rotl32(v4, 18);
    } else {
        h = ZSK_X32_P5;
    }
    return zsk_xxh32_finish(h, p + done, len - done, len);
}

/* XXH32 by four consecutive lanes (gl = 0..3 inside the group, gmask = their lane mask); every lane gets the hash */
static __device__ uint32_t zsk_xxh32_group4(const uint8_t *p, uint32_t len, unsigned gl, unsigned gmask)
{
    uint32_t h, done = 0;
    if (len >= 16) {
        uint32_t v = gl == 0 ? ZSK_X32_P1 + ZSK_X32_P2 : gl == 1 ? ZSK_X32_P2 : gl == 2 ? 0u : 0u - ZSK_X32_P1;
        for (; done + 16 <= len; done += 16) v = zsk_x32_round(v, zsk_ldb32(p + done + 4 * gl));
        const uint32_t r = zsk_rotl32(v, gl == 0 ? 1 : gl == 1 ? 7 : gl == 2 ? 12 : 18);
        h = r + __shfl_xor_sync(gmask, r, 1);
        h += __shfl_xor_sync(gmask, h, 2);
    } else {
        h = ZSK_X32_P5;
    }
    return zsk_xxh32_finish(h, p + done, len - done, len);
}

/* XXH64 by four consecutive lanes; every lane gets the hash */
static __device__ uint64_t zsk_xxh64_group4(const uint8_t *p, uint64_t len, unsigned gl, unsigned gmask)
{
    uint64_t h, done = 0;
    if (len >= 32) {
        uint64_t v = gl == 0 ? ZSK_X64_P1 + ZSK_X64_P2 : gl == 1 ? ZSK_X64_P2 : gl == 2 ? 0ull : 0ull - ZSK_X64_P1;
        for (; done + 32 <= len; done += 32) {
            const uint8_t *q = p + done + 8 * gl;
            v = zsk_x64_round(v, (uint64_t)zsk_ldb32(q) | ((uint64_t)zsk_ldb32(q + 4) << 32));
        }
        const uint64_t v1 = __shfl_sync(gmask, v, 0, 4), v2 = __shfl_sync(gmask, v, 1, 4), v3 = __shfl_sync(gmask, v, 2, 4),
                       v4 = __shfl_sync(gmask, v, 3, 4);
        h = zsk_rotl64(v1, 1) + zsk_rotl64(v2, 7) + zsk_rotl64(v3, 12) + zsk_rotl64(v4, 18);
        h = (h ^ zsk_x64_round(0, v1)) * ZSK_X64_P1 + ZSK_X64_P4;
        h = (h ^ zsk_x64_round(0, v2)) * ZSK_X64_P1 + ZSK_X64_P4;
        h = (h ^ zsk_x64_round(0, v3)) * ZSK_X64_P1 + ZSK_X64_P4;
        h = (h ^ zsk_x64_round(0, v4)) * ZSK_X64_P1 + ZSK_X64_P4;
    } else {
        h = ZSK_X64_P5;
    }
    h += len;
    const uint8_t *q = p + done;
    uint64_t rest = len - done;
    while (rest >= 8) {
        const uint64_t k = zsk_x64_round(0, (uint64_t)zsk_ldb32(q) | ((uint64_t)zsk_ldb32(q + 4) << 32));
        h = zsk_rotl64(h ^ k, 27) * ZSK_X64_P1 + ZSK_X64_P4;
        q += 8; rest -= 8;
    }
    if (rest >= 4) { h = zsk_rotl64(h ^ ((uint64_t)zsk_ldb32(q) * ZSK_X64_P1), 23) * ZSK_X64_P2 + ZSK_X64_P3; q += 4; rest -= 4; }
    while (rest) { h = zsk_rotl64(h ^ ((uint64_t)(*q) * ZSK_X64_P5), 11) * ZSK_X64_P1; q++; rest--; }
    h ^= h >> 33; h *= ZSK_X64_P2; h ^= h >> 29; h *= ZSK_X64_P3; h ^= h >> 32;
    return h;
}
