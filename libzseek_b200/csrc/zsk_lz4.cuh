/*
 * zsk_lz4.cuh — K2: LZ4 frame decode, one LANE GROUP per frame (replaces the reference's calls into liblz4,
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

/* Decodes one LZ4 block of n bytes at src into out[*pop ...); returns a zsk_status.
 *
 * G = 8 lanes work on the block; four such groups (four different frames) share a warp and run the
 * SAME instruction stream, so the common sequence shape is handled by one branch-free, predicated
 * fast path (no divergence between the groups of a warp):
 *     literal length <= 14 (token nibble < 15), match length <= 18 (nibble < 15),
 *     no overlap between the match source and anything this sequence writes (off >= ll + ml)
 * which covers ~95 % of the sequences of text-like data.  Everything else (length-extension bytes,
 * overlapping matches, the final literal-only sequence, any bounds violation) leaves through the
 * general path, which is the plain sequential algorithm.
 */
template <unsigned G>
static __device__ __forceinline__ int zsk_lz4_block_group(const uint8_t *__restrict__ src, uint32_t n, uint8_t *out,
                                                          uint32_t *pop, uint32_t cap, unsigned lane, unsigned gmask)
{
    uint32_t ip = 0, op = *pop;
    if (n == 0) return ZSK_ST_FORMAT;
    for (;;) {
        if (ip >= n) return ZSK_ST_TRUNC;
        const uint32_t tok = ZSK_LDG(src + ip);
        uint32_t ll = tok >> 4, ml = (tok & 15) + 4;
        if (G == 8) {
            /* ---- fast path (all quantities group-uniform) */
            const uint32_t after = ip + 1 + ll;            /* offset position */
            const bool shape = ll < 15 && ml < 19 && after + 2 < n && ll + ml <= cap - op;
            if (shape) {
                const uint32_t off = zsk_rd16(src + after);
                if (off >= ll + ml && off <= op + ll) {
                    const uint8_t *lp = src + ip + 1;
                    uint8_t *o = out + op;
                    const bool l0 = lane < ll, l1 = lane + 8 < ll;
                    uint32_t a0 = 0, a1 = 0;
                    if (l0) a0 = ZSK_LDG(lp + lane);
                    if (l1) a1 = ZSK_LDG(lp + lane + 8);
                    const uint8_t *m = o + ll - off;        /* match source; disjoint from [o, o + ll + ml) */
                    const bool m0 = lane < ml, m1 = lane + 8 < ml, m2 = lane + 16 < ml;
                    uint32_t b0 = 0, b1 = 0, b2 = 0;
                    if (m0) b0 = m[lane];
                    if (m1) b1 = m[lane + 8];
                    if (m2) b2 = m[lane + 16];
                    if (l0) o[lane] = (uint8_t)a0;
                    if (l1) o[lane + 8] = (uint8_t)a1;
                    o += ll;
                    if (m0) o[lane] = (uint8_t)b0;
                    if (m1) o[lane + 8] = (uint8_t)b1;
                    if (m2) o[lane + 16] = (uint8_t)b2;
                    ip = after + 2;
                    op += ll + ml;
                    __syncwarp(gmask); /* this sequence's bytes are visible to the group's later match loads */
                    continue;
                }
            }
        }
        /* ---- general path */
        ip++;
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
        if (ll <= G) {
            if (lane < ll) out[op + lane] = ZSK_LDG(src + ip + lane);
        } else {
            zsk_group_copy(out + op, src + ip, ll, lane, G);
        }
        ip += ll;
        op += ll;
        if (ip == n) break; /* last sequence carries literals only */
        if (n - ip < 2) return ZSK_ST_TRUNC;
        const uint32_t off = zsk_rd16(src + ip);
        ip += 2;
        if (ml == 19) {
            uint32_t b;
            do {
                if (ip >= n) return ZSK_ST_TRUNC;
                b = ZSK_LDG(src + ip);
                ip++;
                ml += b;
            } while (b == 255);
        }
        if (off == 0 || off > op) return ZSK_ST_OFFSET; /* never before the frame start */
        if (ml > cap - op) return ZSK_ST_DST;
        __syncwarp(gmask); /* literal stores of all lanes are visible to the match loads */
        zsk_group_match<G>(out, op, off, ml, lane, gmask);
        __syncwarp(gmask);
        op += ml;
    }
    *pop = op;
    return ZSK_ST_OK;
}

/* Decodes one complete LZ4 frame (header, block chain, EndMark); all lanes pass identical arguments.
 * *produced receives the decoded size. */
template <unsigned G>
static __device__ __forceinline__ int zsk_lz4_frame_group(const uint8_t *__restrict__ src, uint32_t n, uint8_t *out,
                                                          uint32_t cap, uint32_t *produced, unsigned lane, unsigned gmask)
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
            zsk_group_copy(out + op, src + ip, bs, lane, G);
            op += bs;
            __syncwarp(gmask);
        } else {
            int st = zsk_lz4_block_group<G>(src + ip, bs, out, &op, cap, lane, gmask);
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
#define ZSK_LZ4_GROUP 8 /* default lanes per frame: 4 frames decode side by side in one warp */

/* Plain variant: every group runs the sequential frame decoder on its own (kept for A/B runs with
 * other group widths, ZSEEK_B200_LZ4_GROUP=4|8|16|32). */
template <unsigned G>
__global__ void __launch_bounds__(ZSK_LZ4_CTA_THREADS) zsk_lz4_decode_kernel(zsk_decode_args a)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned gl = lane & (G - 1);
    const unsigned gmask = (G == 32 ? ZSK_FULL : ((1u << G) - 1u)) << (lane & ~(G - 1));
    for (;;) {
        uint32_t job = 0;
        if (gl == 0) job = atomicAdd(a.work_counter, 1u);
        job = __shfl_sync(gmask, job, 0, G);
        if (job >= a.njobs) break;
        const uint32_t f = a.frame_ids ? a.frame_ids[job] : a.first_frame + job;
        const uint64_t c0 = a.c_off[f], c1 = a.c_off[f + 1], d0 = a.d_off[f], d1 = a.d_off[f + 1];
        const uint8_t *src = a.comp + (c0 - a.comp_base);
        uint8_t *out = a.dst + (a.dst_offs ? a.dst_offs[job] : d0 - a.dst_base);
        const uint32_t cap = (uint32_t)(d1 - d0);
        uint32_t produced = 0;
        int st = zsk_lz4_frame_group<G>(src, (uint32_t)(c1 - c0), out, cap, &produced, gl, gmask);
        if (st == ZSK_ST_OK && produced != cap) st = ZSK_ST_SIZE;
        if (gl == 0) a.status[job] = st;
        __syncwarp(gmask);
    }
}

/*
 * Lock-step variant (the one that ships): four 8-lane groups of a warp decode four frames side by
 * side as ONE instruction stream.  Every trip of the warp loop starts with a full-warp
 * convergence point, then each group advances its own frame by one step of a small state machine
 * (fetch a job + frame header -> block header -> sequences ... -> next job).  Because the groups
 * re-converge every trip, the predicated sequence fast path is issued once for all four frames —
 * without the convergence point the groups drift apart after the first divergent branch and the
 * warp ends up issuing every instruction four times (measured: 2x slower).
 *
 * Fast path (branch-free, predicated): literal length <= 14 (token nibble < 15), match length <= 18
 * (nibble < 15) and no overlap between the match source and anything this sequence writes
 * (off >= ll + ml) — ~95 % of the sequences of text-like data.  Length-extension bytes, overlapping
 * matches, the final literal-only sequence of a block and every bounds violation take the general
 * path (the plain sequential algorithm).
 */
enum { ZSK_LZ4_S_FETCH = 0, ZSK_LZ4_S_BLOCK = 1, ZSK_LZ4_S_SEQ = 2, ZSK_LZ4_S_DONE = 3 };

#ifndef ZSK_LZ4_MIN_CTAS
#define ZSK_LZ4_MIN_CTAS 12
#endif
template <unsigned G>
__global__ void __launch_bounds__(ZSK_LZ4_CTA_THREADS, ZSK_LZ4_MIN_CTAS) zsk_lz4_decode_lockstep_kernel(zsk_decode_args a)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned gl = lane & (G - 1);
    const unsigned gmask = ((1u << G) - 1u) << (lane & ~(G - 1));
    /* per-group state, identical in the 8 lanes of a group */
    int state = ZSK_LZ4_S_FETCH;
    const uint8_t *src = nullptr; /* frame start */
    uint8_t *out = nullptr;
    uint32_t n = 0, ip = 0, bend = 0, op = 0, cap = 0, job = 0, flags = 0, max_block = 0;
    uint64_t content_size = 0;
    for (;;) {
        __syncwarp(); /* per-trip convergence point; orders the previous trip's stores before this trip's loads */
        if (__all_sync(ZSK_FULL, state == ZSK_LZ4_S_DONE)) break;
        int st = ZSK_ST_OK;
        bool frame_end = false;
        /* Every trip each group reads a 16-byte window of its compressed stream with TWO coalesced loads
         * (lane gl takes bytes ip+gl and ip+8+gl): the token, up to 13 literals and the 2-byte offset all come out
         * of these registers by width-8 shuffles, instead of 5 separate loads that each cost one L1 wavefront
         * per group (the kernel is L1-wavefront bound).  Groups that are not in the sequence
         * state this trip read a harmless dummy window so that the shuffles stay warp-uniform. */
        const bool in_seq = state == ZSK_LZ4_S_SEQ;
        const uint8_t *wp = (in_seq ? src + ip : a.comp) + gl;
        const uint32_t w0 = ZSK_LDG(wp), w1 = ZSK_LDG(wp + 8);
        const uint32_t tok = __shfl_sync(ZSK_FULL, w0, 0, G);
        const uint32_t tll = tok >> 4;
        const uint32_t pl = (1 + tll) & 7, ph = (2 + tll) & 7;                     /* offset bytes sit at window positions 1+ll, 2+ll */
        const uint32_t lo0 = __shfl_sync(ZSK_FULL, w0, pl, G), lo1 = __shfl_sync(ZSK_FULL, w1, pl, G);
        const uint32_t hi0 = __shfl_sync(ZSK_FULL, w0, ph, G), hi1 = __shfl_sync(ZSK_FULL, w1, ph, G);
        if (in_seq) {
            uint32_t ll = tll, ml = (tok & 15) + 4;
            const uint32_t after = ip + 1 + ll;
            uint32_t off = ((1 + ll < 8) ? lo0 : lo1) | (((2 + ll < 8) ? hi0 : hi1) << 8);
            const bool fast = ll <= 13 && ml < 19 && after + 2 < bend && ll + ml <= cap - op && off >= ll + ml && off <= op + ll;
            if (fast) {
                /* predicated, branch-free: literals straight out of the window registers, ceil(18/G) match passes;
                 * every load is issued before the first store */
                constexpr unsigned MP = (18 + G - 1) / G;
                uint8_t *o = out + op + gl;
                const uint8_t *m = o + ll - off; /* match source, disjoint from [o, o + ll + ml) */
                uint32_t mv[MP];
#pragma unroll
                for (unsigned k = 0; k < MP; k++) mv[k] = (gl + k * G < ml) ? m[k * G] : 0;
                if (gl >= 1 && gl <= ll) o[-1] = (uint8_t)w0;        /* window byte gl is literal gl-1 */
                if (gl + 7 < ll) o[7] = (uint8_t)w1;                 /* window byte 8+gl is literal 7+gl */
                o += ll;
#pragma unroll
                for (unsigned k = 0; k < MP; k++) if (gl + k * G < ml) o[k * G] = (uint8_t)mv[k];
                ip = after + 2;
                op += ll + ml;
            } else {
                ip++;
                if (ll == 15) {
                    uint32_t b;
                    do {
                        if (ip >= bend) { st = ZSK_ST_TRUNC; break; }
                        b = ZSK_LDG(src + ip);
                        ip++;
                        ll += b;
                    } while (b == 255);
                }
                if (!st && ll > bend - ip) st = ZSK_ST_TRUNC;
                if (!st && ll > cap - op) st = ZSK_ST_DST;
                if (!st) {
                    zsk_group_copy(out + op, src + ip, ll, gl, G);
                    ip += ll;
                    op += ll;
                    if (ip == bend) {
                        state = ZSK_LZ4_S_BLOCK; /* last sequence of the block: literals only */
                    } else if (bend - ip < 2) {
                        st = ZSK_ST_TRUNC;
                    } else {
                        off = zsk_rd16(src + ip);
                        ip += 2;
                        if (ml == 19) {
                            uint32_t b;
                            do {
                                if (ip >= bend) { st = ZSK_ST_TRUNC; break; }
                                b = ZSK_LDG(src + ip);
                                ip++;
                                ml += b;
                            } while (b == 255);
                        }
                        if (!st && (off == 0 || off > op)) st = ZSK_ST_OFFSET;
                        if (!st && ml > cap - op) st = ZSK_ST_DST;
                        if (!st) {
                            __syncwarp(gmask);
                            zsk_group_match<G>(out, op, off, ml, gl, gmask);
                            op += ml;
                            if (ip >= bend) st = ZSK_ST_TRUNC; /* a block never ends with a match */
                        }
                    }
                }
                __syncwarp(gmask);
            }
            /* fast-path stores need no group barrier of their own: every trip begins with a full-warp barrier */
        } else if (state == ZSK_LZ4_S_BLOCK) {
            if (bend && (flags & 16)) { /* block checksum after the compressed block just finished */
                if (n - ip < 4) st = ZSK_ST_TRUNC; else ip += 4;
            }
            bend = 0;
            if (!st && n - ip < 4) st = ZSK_ST_TRUNC;
            if (!st) {
                uint32_t bs = zsk_rd32(src + ip);
                ip += 4;
                if (bs == 0) { /* EndMark */
                    if ((flags & 4) && n - ip < 4) st = ZSK_ST_TRUNC;
                    else if ((flags & 8) && content_size != op) st = ZSK_ST_FORMAT;
                    else if (op != cap) st = ZSK_ST_SIZE;
                    frame_end = true;
                } else {
                    const bool raw = bs >> 31;
                    bs &= 0x7FFFFFFFu;
                    if (bs > max_block) st = ZSK_ST_FORMAT;
                    else if (bs > n - ip) st = ZSK_ST_TRUNC;
                    else if (raw) {
                        if (bs > cap - op) st = ZSK_ST_DST;
                        else {
                            zsk_group_copy(out + op, src + ip, bs, gl, G);
                            op += bs;
                            ip += bs;
                            if (flags & 16) { if (n - ip < 4) st = ZSK_ST_TRUNC; else ip += 4; }
                            __syncwarp(gmask);
                        }
                    } else {
                        bend = ip + bs;
                        state = ZSK_LZ4_S_SEQ;
                    }
                }
            }
        } else if (state == ZSK_LZ4_S_FETCH) {
            if (gl == 0) job = atomicAdd(a.work_counter, 1u);
            job = __shfl_sync(gmask, job, 0, G);
            if (job >= a.njobs) state = ZSK_LZ4_S_DONE;
            else {
                const uint32_t f = a.frame_ids ? a.frame_ids[job] : a.first_frame + job;
                const uint64_t c0 = a.c_off[f], c1 = a.c_off[f + 1], d0 = a.d_off[f], d1 = a.d_off[f + 1];
                src = a.comp + (c0 - a.comp_base);
                out = a.dst + (a.dst_offs ? a.dst_offs[job] : d0 - a.dst_base);
                cap = (uint32_t)(d1 - d0);
                n = (uint32_t)(c1 - c0);
                op = 0;
                bend = 0;
                if (n < 7) st = ZSK_ST_TRUNC;
                else if (zsk_rd32(src) != ZSK_LZ4_MAGIC) st = ZSK_ST_MAGIC;
                else {
                    const uint32_t flg = ZSK_LDG(src + 4), bd = ZSK_LDG(src + 5);
                    const uint32_t bsid = (bd >> 4) & 7;
                    if ((flg >> 6) != 1 || (flg & 0x02) || (bd & 0x8F) || bsid < 4) st = ZSK_ST_FORMAT;
                    else {
                        flags = flg;
                        max_block = 1u << (8 + 2 * bsid);
                        ip = 6;
                        if (flg & 8) {
                            if (n - ip < 8) st = ZSK_ST_TRUNC;
                            else { content_size = zsk_rd64(src + ip); ip += 8; }
                        }
                        if (!st && (flg & 1)) { if (n - ip < 4) st = ZSK_ST_TRUNC; else ip += 4; }
                        if (!st) { if (n - ip < 1) st = ZSK_ST_TRUNC; else ip += 1; }
                        if (!st) state = ZSK_LZ4_S_BLOCK;
                    }
                }
                if (st) frame_end = true;
            }
        }
        if (st || frame_end) {
            if (gl == 0) a.status[job] = st;
            state = ZSK_LZ4_S_FETCH;
        }
    }
}


