/*
 * zsk_seek.cuh — K1 (batched offset -> frame lookup) and K4 (range gather).
 *
 * K1 replaces offset_to_frame_idx (reference src/seek_table.c:187-202) plus the result arithmetic of
 * zseek_pread (reference src/decompress.c:445,556-557) for a whole batch of requests: one thread per
 * request runs the same binary search over the device-resident decompressed prefix array (<= 17 probes
 * into an L2-resident 512 KiB array at N = 65,536) and emits
 *     frame   = largest i in [0,N) with d_off[i] <= offset, or -1 when offset >= d_off[N]   (B2, B4)
 *     inframe = offset - d_off[frame]
 *     nbytes  = MIN(count, d_off[frame+1] - offset)   — a read never crosses a frame boundary (B1, B3)
 * and, per frame, touched[frame] = the largest inframe + nbytes any request of the batch reaches (>= 1 when touched).
 *
 * K4 replaces the memcpy out of the cached frame (reference src/decompress.c:558,788): one warp per
 * request copies nbytes from the decoded frame (HBM frame cache slot) to the caller's buffer with
 * 16-byte stores, source re-aligned by funnel shift.
 */
#pragma once
#include "zsk_common.cuh"

static __device__ __forceinline__ int64_t zsk_offset_to_frame(const uint64_t *__restrict__ d_off, uint32_t nframes, uint64_t offset)
{
    if (offset >= d_off[nframes]) return -1;
    uint32_t lo = 0, hi = nframes;
    while (lo + 1 < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (d_off[mid] <= offset) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) zsk_lookup_kernel(zsk_lookup_args a)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const uint64_t off = a.offsets[i];
        const uint64_t cnt = a.counts ? a.counts[i] : a.fixed_count;
        const int64_t f = a.nframes ? zsk_offset_to_frame(a.d_off, a.nframes, off) : -1;
        uint32_t inframe = 0, nb = 0;
        if (f >= 0) {
            const uint64_t d0 = a.d_off[f], d1 = a.d_off[f + 1];
            inframe = (uint32_t)(off - d0);
            const uint64_t room = d1 - off;
            nb = (uint32_t)(cnt < room ? cnt : room);
            if (a.touched) atomicMax(&a.touched[f], max(1u, inframe + nb)); /* how much of the frame this batch needs */
        }
        a.frame[i] = (int32_t)f;
        a.inframe[i] = inframe;
        a.nbytes[i] = nb;
    }
}

__global__ void __launch_bounds__(256) zsk_gather_kernel(zsk_gather_args a)
{
    const unsigned lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < a.n; i += warps) {
        const int32_t f = a.frame[i];
        const uint32_t nb = a.nbytes[i];
        if (a.results && lane == 0) a.results[i] = (int64_t)nb;
        if (f < 0 || nb == 0) continue;
        const int64_t so = a.frame_src[f];
        if (so < 0) continue;
        const uint8_t *src = a.src_base + so + a.inframe[i];
        uint8_t *dst = a.dst + (a.dst_offs ? a.dst_offs[i] : (uint64_t)i * a.dst_stride);
        zsk_group_copy(dst, src, nb, lane, 32);
    }
}

/* Job list of a stream-ordered batch (zseek_b200_pread_batch_async): see zsk_compact_args. */
__global__ void __launch_bounds__(256) zsk_compact_kernel(zsk_compact_args a)
{
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < a.nframes; f += gridDim.x * blockDim.x) {
        const uint32_t need = a.touched[f];
        int64_t src = -1;
        if (need) {
            if (f < a.shard_lo || f >= a.shard_hi) *a.error = 1u;
            else {
                const uint32_t j = atomicAdd(a.count, 1u);
                if (j >= a.max_jobs) *a.error = 2u;
                else {
                    a.job_ids[j] = f;
                    a.job_offs[j] = (uint64_t)j * a.slot_size;
                    a.job_limits[j] = need;
                    src = (int64_t)((uint64_t)j * a.slot_size);
                }
            }
        }
        a.frame_src[f] = src;
    }
}
