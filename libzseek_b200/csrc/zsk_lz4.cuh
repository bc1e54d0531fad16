/*
 * zsk_lz4.cuh — K2: LZ4 frame decode, ONE WARP PER FRAME (replaces the reference's calls into liblz4,
 * LZ4F_decompress at reference src/decompress.c:631,653,762; format: SURVEY.md Appendix A.1).
 *
 * A warp walks the frame header and the block chain.  Inside a block the warp walks the sequence
 * chain with warp-uniform control flow: token / length-extension / offset bytes are read through
 * the read-only path as warp-broadcast loads (one L1 wavefront each), literal runs and matches are
 * copied cooperatively by the 32 lanes (coalesced byte stores for short runs, 16-byte vector stores
 * with funnel-shift source realignment for long ones).  Linked blocks need nothing special: matches
 * address the frame's own output, which the same warp wrote earlier.
 *
 * Work distribution: persistent warps pull frame jobs from a global atomic counter, so ragged
 * frames (raw blocks, short tails) do not leave lanes of a CTA idle.
 *
 * HBM traffic per frame: cSize bytes read once + dSize bytes written once; match sources are re-read
 * from L1/L2 (the frame's output was just written by this SM).
 */
#pragma once
#include "zsk_common.cuh"

#define ZSK_LZ4_MAGIC 0x184D2204u

/* Decodes one LZ4 block of n bytes at src into out[*pop ...); returns a zsk_status. */
static __device__ __forceinline__ int zsk_lz4_block_warp(const uint8_t *__restrict__ src, uint32_t n, uint8_t *out,
                                                         uint32_t *pop, uint32_t cap, unsigned lane)
{
    uint32_t ip = 0, op = *pop;
    if (n == 0) return ZSK_ST_FORMAT;
    for (;;) {
        if (ip >= n) return ZSK_ST_TRUNC;
        const uint32_t tok = ZSK_LDG(src + ip);
        ip++;
        uint32_t ll = tok >> 4;
        if (ll == 15) {
            uint32_t b;
            do {
                if (ip >= n) return ZSK_ST_TRUNC;
                b = ZSK_LDG(src + ip);
                ip++;
                ll += b;
            } while (b == 255);
        }
        if (ll > n - ip) return ZSK_ST_TRUNC;
        if (ll > cap - op) return ZSK_ST_DST;
        /* literal run */
        if (ll <= 32) {
            if (lane < ll) out[op + lane] = ZSK_LDG(src + ip + lane);
        } else {
            zsk_group_copy(out + op, src + ip, ll, lane, 32);
        }
        ip += ll;
        op += ll;
        if (ip == n) break; /* last sequence carries literals only */
        if (n - ip < 2) return ZSK_ST_TRUNC;
        const uint32_t off = zsk_rd16(src + ip);
        ip += 2;
        uint32_t ml = tok & 15;
        if (ml == 15) {
            uint32_t b;
            do {
                if (ip >= n) return ZSK_ST_TRUNC;
                b = ZSK_LDG(src + ip);
                ip++;
                ml += b;
            } while (b == 255);
        }
        ml += 4;
        if (off == 0 || off > op) return ZSK_ST_OFFSET; /* never before the frame start */
        if (ml > cap - op) return ZSK_ST_DST;
        __syncwarp(); /* literal stores of all lanes are visible to the match loads */
        zsk_warp_match(out, op, off, ml, lane);
        __syncwarp();
        op += ml;
    }
    *pop = op;
    return ZSK_ST_OK;
}

/* Decodes one complete LZ4 frame (header, block chain, EndMark); all lanes pass identical arguments.
 * *produced receives the decoded size. */
static __device__ __forceinline__ int zsk_lz4_frame_warp(const uint8_t *__restrict__ src, uint32_t n, uint8_t *out,
                                                         uint32_t cap, uint32_t *produced, unsigned lane)
{
    if (n < 7) return ZSK_ST_TRUNC;
    if (zsk_rd32(src) != ZSK_LZ4_MAGIC) return ZSK_ST_MAGIC;
    const uint32_t flg = ZSK_LDG(src + 4), bd = ZSK_LDG(src + 5);
    if ((flg >> 6) != 1 || (flg & 0x02) || (bd & 0x8F)) return ZSK_ST_FORMAT;
    const uint32_t bsid = (bd >> 4) & 7;
    if (bsid < 4) return ZSK_ST_FORMAT;
    const uint32_t max_block = 1u << (8 + 2 * bsid);
    const bool block_cksum = (flg >> 4) & 1, has_csize = (flg >> 3) & 1, content_cksum = (flg >> 2) & 1, dict = flg & 1;
    uint32_t ip = 6;
    uint64_t content_size = 0;
    if (has_csize) {
        if (n - ip < 8) return ZSK_ST_TRUNC;
        content_size = zsk_rd64(src + ip);
        ip += 8;
    }
    if (dict) {
        if (n - ip < 4) return ZSK_ST_TRUNC;
        ip += 4;
    }
    if (n - ip < 1) return ZSK_ST_TRUNC;
    ip += 1; /* header checksum byte */
    uint32_t op = 0;
    for (;;) {
        if (n - ip < 4) return ZSK_ST_TRUNC;
        uint32_t bs = zsk_rd32(src + ip);
        ip += 4;
        if (bs == 0) break; /* EndMark */
        const bool raw = bs >> 31;
        bs &= 0x7FFFFFFFu;
        if (bs > max_block) return ZSK_ST_FORMAT;
        if (bs > n - ip) return ZSK_ST_TRUNC;
        if (raw) {
            if (bs > cap - op) return ZSK_ST_DST;
            zsk_group_copy(out + op, src + ip, bs, lane, 32);
            op += bs;
            __syncwarp();
        } else {
            int st = zsk_lz4_block_warp(src + ip, bs, out, &op, cap, lane);
            if (st) return st;
        }
        ip += bs;
        if (block_cksum) {
            if (n - ip < 4) return ZSK_ST_TRUNC;
            ip += 4;
        }
    }
    if (content_cksum && n - ip < 4) return ZSK_ST_TRUNC;
    if (has_csize && content_size != op) return ZSK_ST_FORMAT;
    *produced = op;
    return ZSK_ST_OK;
}

#define ZSK_LZ4_CTA_THREADS 128

__global__ void __launch_bounds__(ZSK_LZ4_CTA_THREADS) zsk_lz4_decode_kernel(zsk_decode_args a)
{
    const unsigned lane = threadIdx.x & 31;
    for (;;) {
        uint32_t job = 0;
        if (lane == 0) job = atomicAdd(a.work_counter, 1u);
        job = __shfl_sync(ZSK_FULL, job, 0);
        if (job >= a.njobs) break;
        const uint32_t f = a.frame_ids ? a.frame_ids[job] : a.first_frame + job;
        const uint64_t c0 = a.c_off[f], c1 = a.c_off[f + 1], d0 = a.d_off[f], d1 = a.d_off[f + 1];
        const uint8_t *src = a.comp + (c0 - a.comp_base);
        uint8_t *out = a.dst + (a.dst_offs ? a.dst_offs[job] : d0 - a.dst_base);
        const uint32_t cap = (uint32_t)(d1 - d0);
        uint32_t produced = 0;
        int st = zsk_lz4_frame_warp(src, (uint32_t)(c1 - c0), out, cap, &produced, lane);
        if (st == ZSK_ST_OK && produced != cap) st = ZSK_ST_SIZE;
        if (lane == 0) a.status[job] = st;
        __syncwarp();
    }
}
