/*
 * zsk_lz4_lane.cuh — K2b: LZ4 frame decode, one LANE per frame, software-pipelined (replaces the reference's
 * calls into liblz4, LZ4F_decompress at reference src/decompress.c:631,653,762; format: SURVEY.md Appendix A.1).
 *
 * Why: the warp-per-frame kernels are bound by the instruction issue rate — a sequence moves
 * ~10 bytes and costs 30-50 warp instructions however the 32 lanes share it.  Here a lane owns a frame and
 * runs the plain sequential algorithm, so one warp instruction advances 32 frames.  What is left is the memory
 * system, and the kernel is built around it:
 *
 *   - per lane, shared memory holds an INPUT ring (ZSK_LZ4L_RI bytes of compressed stream), an OUTPUT ring
 *     (ZSK_LZ4L_RO bytes behind the write frontier; word-interleaved across the warp, so lane-private accesses
 *     are bank-conflict free) and ZSK_LZ4L_SLOTS staging slots of 32 bytes for match sources;
 *   - NOTHING in a trip waits for global memory.  A trip has a PARSE half and an EXECUTE half that work on
 *     different sequences: the parse half decodes the token at the read position into a MICRO-OP (<= 8
 *     literal bytes, carried in registers, then <= 16 match bytes) and, when the match source is older than
 *     the output ring, starts cp.async copies of its (at most two) 16-byte chunks from global memory into a
 *     staging slot; the execute half runs the micro-op parsed ZSK_LZ4L_DEPTH trips earlier, whose copies
 *     have landed (cp.async.wait_group), entirely inside shared memory.  The input ring is refilled by
 *     cp.async as well, 16 bytes per lane and trip;
 *   - output leaves through the ring: a lane writes a completed, 32-byte aligned sector with two 16-byte
 *     stores (full sectors reach L2, no byte stores to global memory).  At most ~140 bytes are ever parsed
 *     but not yet flushed, so a match with offset > ZSK_LZ4L_NEAR can safely be fetched from global memory
 *     at parse time, and one with a smaller offset is still in the ring at execute time;
 *   - a common sequence (literal run <= 8, match <= 16, offset >= match length) is exactly one micro-op;
 *     longer literal runs, longer or overlapping matches, length-extension bytes, block and frame headers
 *     are spread over several trips by a small per-lane phase machine, so lanes never loop on their own.
 *     Overlapping matches double their effective offset after every full period (the output is periodic),
 *     so run-length patterns reach 16 bytes per trip after <= 4 trips.
 *
 * No cross-lane communication exists apart from the exit vote.  Persistent lanes pull frame jobs from the
 * global atomic counter.  A corrupt or truncated frame ends in a non-zero status, never in a hang or an
 * out-of-bounds access (ring indices are masked, global reads stay inside [frame output) +- 15 bytes and the
 * padded compressed image).
 */
#pragma once
#include "zsk_common.cuh"

#ifndef ZSK_LZ4_MAGIC
#define ZSK_LZ4_MAGIC 0x184D2204u
#endif

#ifndef ZSK_LZ4L_RI
#define ZSK_LZ4L_RI 64u /* input ring bytes per lane (power of two; >= 64 for DEPTH 2, >= 128 for DEPTH 3) */
#endif
#ifndef ZSK_LZ4L_RO
#define ZSK_LZ4L_RO 256u /* output ring bytes per lane (power of two, >= 256) */
#endif
#define ZSK_LZ4L_NEAR (ZSK_LZ4L_RO - 64u) /* matches with offset <= NEAR are served from the output ring */
#ifndef ZSK_LZ4L_DEPTH
#define ZSK_LZ4L_DEPTH 2u                 /* trips between parse and execute of a micro-op (1..3) */
#endif
#ifndef ZSK_LZ4L_SLOTS
#define ZSK_LZ4L_SLOTS 4u                 /* staging slots (>= DEPTH + 1, power of two) */
#endif
#ifndef ZSK_LZ4L_WARPS
#define ZSK_LZ4L_WARPS 2u
#endif
#define ZSK_LZ4L_THREADS (32u * ZSK_LZ4L_WARPS)
#define ZSK_LZ4L_WORDS_PER_WARP (32u * (ZSK_LZ4L_RI + ZSK_LZ4L_RO + 32u * ZSK_LZ4L_SLOTS) / 4u)
#define ZSK_LZ4L_SMEM (ZSK_LZ4L_WARPS * ZSK_LZ4L_WORDS_PER_WARP * 4u)

enum {
    ZSK_L_FETCH = 0, /* take a job, parse the frame header from global memory */
    ZSK_L_BLOCK,     /* block header / EndMark */
    ZSK_L_TOKEN,     /* sequence token */
    ZSK_L_LLEXT,     /* literal-length extension bytes, one per trip */
    ZSK_L_LIT,       /* rest of a literal run (also: raw blocks), 8 bytes per trip */
    ZSK_L_OFF,       /* match offset after a long literal run */
    ZSK_L_MLEXT,     /* match-length extension bytes, one per trip */
    ZSK_L_MATCH,     /* rest of a match, <= 16 bytes per trip */
    ZSK_L_DRAIN,     /* frame finished by the parser; wait until the executor has run its last micro-op */
    ZSK_L_DONE
};

/* micro-op word: [3:0] literal bytes, [8:4] match bytes, [9] source in staging slot, [10] end of frame,
 * [31:16] near: match offset / staged: source misalignment inside its 16-byte chunk / end: status */
#define ZSK_L_MOP_FAR 0x200u
#define ZSK_L_MOP_END 0x400u
#define ZSK_L_MOP_CK 0x800u /* with END: the literal word carries the frame's content checksum, to be verified */

#define ZSK_L_IW(j) inw[(j) & (ZSK_LZ4L_RI / 4u - 1u)]
#define ZSK_L_OW(j) outr[(((j) & (ZSK_LZ4L_RO / 4u - 1u)) << 5)]

/* Writes the low `len` (<= 4*(NW-1)) bytes of the little-endian string x[0..NW-2] at byte position y of the
 * lane's output ring.  Bytes of the last touched word beyond the string are overwritten with garbage; they
 * belong to the oldest 3 bytes of the ring, which nobody reads (NEAR leaves 64 bytes of slack). */
template <unsigned NW>
static __device__ __forceinline__ void zsk_l_ring_write(uint32_t *outr, uint32_t y, const uint32_t *x, uint32_t len)
{
    const uint32_t a = y & 3u, s = a * 8u, j = y >> 2, nb = a + len;
    if (len) {
        const uint32_t old = ZSK_L_OW(j);
        const uint32_t keep = old & ~(0xffffffffu << s); /* the a bytes below y stay */
        ZSK_L_OW(j) = keep | (x[0] << s);
    }
#pragma unroll
    for (unsigned k = 1; k < NW; k++) {
        if (nb > 4u * k) {
            const uint32_t lo = x[k - 1], hi = (k < NW - 1) ? x[k] : 0u;
            ZSK_L_OW(j + k) = __funnelshift_l(lo, hi, s);
        }
    }
}

__global__ void __launch_bounds__(ZSK_LZ4L_THREADS) zsk_lz4_decode_lane_kernel(zsk_decode_args a)
{
#ifdef ZSK_EMU
    uint32_t *smem = (uint32_t *)zsk_emu_dyn_smem();
#else
    extern __shared__ __align__(16) uint32_t zsk_l_smem[];
    uint32_t *smem = zsk_l_smem;
#endif
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wbase = smem + warp * ZSK_LZ4L_WORDS_PER_WARP;
    uint32_t *inw = wbase + lane * (ZSK_LZ4L_RI / 4u);                   /* my input ring, contiguous (cp.async target) */
    uint32_t *outr = wbase + 32u * (ZSK_LZ4L_RI / 4u) + lane;            /* word j of my output ring: outr[j*32] */
    uint32_t *stg = wbase + 32u * ((ZSK_LZ4L_RI + ZSK_LZ4L_RO) / 4u) + lane * 4u; /* 16-byte chunk c of my staging: stg[c*128 ..+4) */
    zsk_cp_reset();

    /* parser state */
    int phase = ZSK_L_FETCH;
    const uint8_t *src = nullptr; /* frame start (compressed) */
    uint8_t *out = nullptr;       /* frame start (decoded) */
    uint32_t n = 0, ip = 0, bend = 0, opp = 0, cap = 0, job = 0, flags = 0, max_block = 0;
    uint32_t sal = 0, oal = 0;    /* misalignment of src inside its 16-byte chunk / of out inside its 32-byte sector */
    uint32_t fx = 0, fhist = 0;   /* input ring refills issued up to fx (ip + sal coordinates); refill flags of the last DEPTH trips */
    uint32_t lrem = 0, mrem = 0, moff = 0, moffe = 0, mnib = 0, drain = 0, stop = 0xffffffffu, bstart = 0, cksum = 0;
    uint64_t content_size = 0;
    /* executor state */
    uint32_t ope = 0, flushed = 0; /* decoded bytes [0, ope) are in the ring or in global memory, [0, flushed) in global memory */
    /* micro-ops in flight: q*[0] was parsed one trip ago, q*[DEPTH-1] runs this trip */
    uint32_t qm[ZSK_LZ4L_DEPTH], ql0[ZSK_LZ4L_DEPTH], ql1[ZSK_LZ4L_DEPTH];
#pragma unroll
    for (unsigned k = 0; k < ZSK_LZ4L_DEPTH; k++) qm[k] = ql0[k] = ql1[k] = 0u;
    uint32_t slot = 0; /* staging slot of the micro-op parsed this trip */

    for (;;) {
        if (__all_sync(ZSK_FULL, phase == ZSK_L_DONE)) break;
        const bool streaming = phase >= ZSK_L_BLOCK && phase <= ZSK_L_MATCH;
        int st = ZSK_ST_OK;
        bool frame_end = false, verify = false;

        /* ---- 1. input ring refill (lands within DEPTH trips) */
        const uint32_t x = ip + sal, xend = sal + n;
        const uint32_t landed = fx - 16u * (uint32_t)__popc(fhist);
        const bool do_fill = streaming && fx < xend && fx + 16u - (x & ~15u) <= ZSK_LZ4L_RI;
        if (do_fill) {
            zsk_cp16(&ZSK_L_IW(fx >> 2), src - sal + fx);
            fx += 16u;
        }
        fhist = ((fhist << 1) | (do_fill ? 1u : 0u)) & ((1u << ZSK_LZ4L_DEPTH) - 1u);

        /* ---- 2. 12-byte window at ip */
        const bool ready = streaming && (landed >= x + 12u || landed >= xend);
        uint32_t a0, a1, a2;
        {
            const uint32_t j = x >> 2, sh = (x & 3u) * 8u;
            const uint32_t w0 = ZSK_L_IW(j), w1 = ZSK_L_IW(j + 1), w2 = ZSK_L_IW(j + 2), w3 = ZSK_L_IW(j + 3);
            a0 = __funnelshift_r(w0, w1, sh);
            a1 = __funnelshift_r(w1, w2, sh);
            a2 = __funnelshift_r(w2, w3, sh);
        }

        /* ---- 3. phase machine: decide this trip's micro-op */
        uint32_t nlit = 0, lpos = 0, mlen = 0, muse = 0;
        if (streaming && opp >= stop) {
            frame_end = true; /* the caller needs no byte beyond `stop` of this frame: close it as if the EndMark had come */
        } else if (ready) {
            if (phase == ZSK_L_TOKEN) {
                const uint32_t tok = a0 & 0xffu, L = tok >> 4, M = tok & 15u;
                if (ip >= bend) {
                    st = ZSK_ST_TRUNC; /* a block never ends with a match */
                } else if (L <= 8u) {
                    const uint32_t after = ip + 1u + L;
                    lpos = 1;
                    nlit = L;
                    if (after == bend) { /* last sequence of the block: literals only */
                        ip = after;
                        phase = ZSK_L_BLOCK;
                    } else if (after + 2u > bend) {
                        st = ZSK_ST_TRUNC;
                    } else {
                        const uint32_t pos = 1u + L, sh = (pos & 3u) * 8u;      /* offset bytes at window positions 1+L, 2+L (<= 10) */
                        const uint32_t lo = pos < 4u ? a0 : (pos < 8u ? a1 : a2), hi = pos < 4u ? a1 : (pos < 8u ? a2 : 0u);
                        const uint32_t off = __funnelshift_r(lo, hi, sh) & 0xffffu;
                        ip = after + 2u;
                        if (off == 0u || off > opp + L) {
                            st = ZSK_ST_OFFSET;
                        } else if (M == 15u) {
                            mrem = 19u;
                            moff = off;
                            phase = ZSK_L_MLEXT;
                        } else {
                            const uint32_t ml = M + 4u;
                            mlen = min(ml, min(16u, off));
                            muse = off;
                            mrem = ml - mlen;
                            moff = off;
                            moffe = (mlen == off) ? off * 2u : off;
                            phase = mrem ? ZSK_L_MATCH : ZSK_L_TOKEN;
                        }
                    }
                } else if (L < 15u) { /* 9..14 literals: 8 now, the rest next trip */
                    if (bend - ip < 9u) st = ZSK_ST_TRUNC;
                    else {
                        lpos = 1;
                        nlit = 8;
                        lrem = L - 8u;
                        mnib = M;
                        ip += 9u;
                        phase = ZSK_L_LIT;
                    }
                } else {
                    lrem = 15u;
                    mnib = M;
                    ip += 1u;
                    phase = ZSK_L_LLEXT;
                }
            } else if (phase != ZSK_L_BLOCK) {
                /* continuation phases (rest of a match, rest of a literal run, offset after a long literal run, length
                 * extension bytes): one block of select logic, so that lanes in different continuation phases share
                 * one pass of the warp instead of one pass per phase */
                const bool is_m = phase == ZSK_L_MATCH, is_l = phase == ZSK_L_LIT, is_o = phase == ZSK_L_OFF;
                const uint32_t ln = min(lrem, 8u), mn = min(mrem, min(16u, moffe));
                const uint32_t eat = is_m ? 0u : (is_l ? ln : (is_o ? 2u : 1u));
                if (eat > bend - ip) {
                    st = ZSK_ST_TRUNC;
                } else {
                    ip += eat;
                    if (is_m) {
                        mlen = mn;
                        muse = moffe;
                        mrem -= mn;
                        if (mn == moffe) moffe *= 2u; /* a full period was appended: the pattern now repeats with twice the period */
                        if (mrem == 0u) phase = ZSK_L_TOKEN;
                    } else if (is_l) {
                        nlit = ln;
                        lrem -= ln;
                        if (lrem == 0u) phase = (ip == bend) ? ZSK_L_BLOCK : ZSK_L_OFF;
                    } else if (is_o) {
                        const uint32_t off = a0 & 0xffffu;
                        if (off == 0u || off > opp) st = ZSK_ST_OFFSET;
                        moff = off;
                        moffe = off;
                        mrem = mnib == 15u ? 19u : mnib + 4u;
                        phase = mnib == 15u ? ZSK_L_MLEXT : ZSK_L_MATCH;
                    } else {
                        const uint32_t b = a0 & 0xffu;
                        if (phase == ZSK_L_LLEXT) {
                            lrem += b;
                            if (b != 255u) phase = ZSK_L_LIT;
                        } else {
                            mrem += b;
                            moffe = moff;
                            if (b != 255u) phase = ZSK_L_MATCH;
                        }
                    }
                }
            } else { /* ZSK_L_BLOCK */
                const uint32_t skip = (bend && (flags & 16u)) ? 4u : 0u; /* block checksum after the block just finished */
                if (n - ip < skip + 4u) st = ZSK_ST_TRUNC;
                else if (skip && zsk_xxh32_serial(src + bstart, bend - bstart) != a0) st = ZSK_ST_CHECKSUM; /* XXH32 of its compressed bytes */
                bend = 0;
                if (!st) {
                    uint32_t bs = skip ? a1 : a0;
                    ip += skip + 4u;
                    if (bs == 0u) { /* EndMark */
                        if ((flags & 4u) && n - ip < 4u) st = ZSK_ST_TRUNC;
                        else if ((flags & 8u) && content_size != opp) st = ZSK_ST_FORMAT;
                        else if (opp != cap) st = ZSK_ST_SIZE;
                        else if (flags & 4u) { /* content checksum: the executor verifies it once the frame is written */
                            cksum = skip ? a2 : a1;
                            verify = true;
                        }
                        frame_end = true;
                    } else {
                        const bool raw = bs >> 31;
                        bs &= 0x7FFFFFFFu;
                        if (bs > max_block) st = ZSK_ST_FORMAT;
                        else if (bs > n - ip) st = ZSK_ST_TRUNC;
                        else if (raw) {
                            bstart = ip;
                            bend = ip + bs;
                            lrem = bs;
                            phase = bs ? ZSK_L_LIT : ZSK_L_BLOCK;
                            if (bs > cap - opp) st = ZSK_ST_DST;
                        } else {
                            bstart = ip;
                            bend = ip + bs;
                            phase = ZSK_L_TOKEN;
                        }
                    }
                }
            }
            if (!st && nlit + mlen > cap - opp) st = ZSK_ST_DST;
        }

        /* ---- 4. the new micro-op; a match source older than the ring is copied into this trip's staging slot */
        uint32_t nm, nl0, nl1;
        if (st || frame_end) {
            nm = ZSK_L_MOP_END | ((verify && !st) ? ZSK_L_MOP_CK : 0u) | ((uint32_t)st << 16);
            nl0 = cksum;
            nl1 = 0u;
            phase = ZSK_L_DRAIN;
            drain = ZSK_LZ4L_DEPTH;
        } else {
            nl0 = lpos ? __funnelshift_r(a0, a1, 8) : a0;
            nl1 = lpos ? __funnelshift_r(a1, a2, 8) : a1;
            opp += nlit;
            nm = nlit | (mlen << 4) | (muse << 16);
            if (mlen && muse > ZSK_LZ4L_NEAR) {
                /* flushed long ago: [p - 15, p + 31) lies below opp - NEAR + 31, and fewer than 56 + 24*DEPTH + 8 bytes are unflushed */
                const uint8_t *p = out + (opp - muse);
                const uint32_t b = (uint32_t)((uintptr_t)p & 15u);
#ifdef ZSK_EXP_L_NOFAR /* ablation build: what do the far match-source loads cost? (the decoded bytes are then wrong) */
                if (mlen == 77u)
#endif
                {
                zsk_cp16(stg + (slot * 2u) * 128u, p - b);
                if (b + mlen > 16u) zsk_cp16(stg + (slot * 2u + 1u) * 128u, p - b + 16);
                }
                nm = nlit | (mlen << 4) | ZSK_L_MOP_FAR | (b << 16);
            }
            opp += mlen;
        }
        zsk_cp_commit();
        zsk_cp_wait<ZSK_LZ4L_DEPTH>(); /* everything issued DEPTH trips ago (or earlier) has landed */

        /* ---- 5. execute the micro-op parsed DEPTH trips ago */
        {
            const uint32_t m = qm[ZSK_LZ4L_DEPTH - 1];
            const uint32_t xlit = m & 15u, xlen = (m >> 4) & 31u, arg = m >> 16;
            {
                uint32_t lw[2] = {ql0[ZSK_LZ4L_DEPTH - 1], ql1[ZSK_LZ4L_DEPTH - 1]};
                zsk_l_ring_write<3>(outr, ope + oal, lw, xlit);
                ope += xlit;
            }
            if (xlen) {
                /* five source words: from the staging slot (two 16-byte chunks, 128 words apart) or from the ring */
                const bool far = m & ZSK_L_MOP_FAR;
                const uint32_t xslot = (slot + ZSK_LZ4L_SLOTS - ZSK_LZ4L_DEPTH) & (ZSK_LZ4L_SLOTS - 1u);
                const uint32_t ys = ope + oal - arg;
                const uint32_t jb = far ? (arg >> 2) : (ys >> 2), sh = ((far ? arg : ys) & 3u) * 8u;
                const uint32_t *sbase = far ? stg + xslot * 256u : outr;
                uint32_t w[5];
#pragma unroll
                for (unsigned k = 0; k < 5; k++) {
                    const uint32_t j = jb + k;
                    const uint32_t idx = far ? (((j >> 2) << 7) | (j & 3u)) : ((j & (ZSK_LZ4L_RO / 4u - 1u)) << 5);
                    w[k] = sbase[idx];
                }
                uint32_t mw[4];
#pragma unroll
                for (unsigned k = 0; k < 4; k++) mw[k] = __funnelshift_r(w[k], w[k + 1], sh);
                zsk_l_ring_write<5>(outr, ope + oal, mw, xlen);
                ope += xlen;
            }
            /* flush one completed, 32-byte aligned chunk (a full sector: two 16-byte stores); a trip appends at most 24
             * bytes, so one flush per trip keeps up.  The frame's unaligned head goes bytewise. */
            {
                const uint32_t chunk = (flushed == 0u && oal) ? 32u - oal : 32u;
                if (ope - flushed >= chunk) {
                    if (chunk == 32u) {
                        const uint32_t *fp = &ZSK_L_OW((flushed + oal) >> 2); /* 32-byte aligned: the 8 words do not wrap */
#ifdef ZSK_LZ4L_SPLIT_FLUSH
                        const uint4 v0 = make_uint4(fp[0], fp[32], fp[64], fp[96]);
                        const uint4 v1 = make_uint4(fp[128], fp[160], fp[192], fp[224]);
                        uint4 *o = (uint4 *)(out + flushed);
                        o[0] = v0;
                        o[1] = v1;
#else
                        /* ONE 256-bit store per sector: an SM retires scattered sector writes slowly (~5 cycles each, measured
                         * on the zstd FSE stage), and 32 lanes write 32 different sectors here */
                        const uint32_t v[8] = { fp[0], fp[32], fp[64], fp[96], fp[128], fp[160], fp[192], fp[224] };
#ifdef ZSK_EXP_L_NOFLUSH /* ablation build (tools/variant_sweep.py): what do the output stores cost? */
                        if (v[0] == 0x12345678u && v[7] == 0x9abcdef0u)
#endif
                        zsk_st256((uint32_t *)(out + flushed), v);
#endif
                        flushed += 32u;
                    } else {
                        for (uint32_t i = 0; i < chunk; i++) {
                            const uint32_t y = i + oal;
                            out[i] = (uint8_t)(ZSK_L_OW(y >> 2) >> ((y & 3u) * 8u));
                        }
                        flushed = chunk;
                    }
                }
            }
            if (m & ZSK_L_MOP_END) {
                uint32_t fst = arg;
                if (arg == 0u) {
                    for (uint32_t i = flushed; i < ope; i++) {
                        const uint32_t y = i + oal;
                        out[i] = (uint8_t)(ZSK_L_OW(y >> 2) >> ((y & 3u) * 8u));
                    }
                    if ((m & ZSK_L_MOP_CK) && zsk_xxh32_serial(out, ope) != ql0[ZSK_LZ4L_DEPTH - 1]) fst = ZSK_ST_CHECKSUM;
                }
                a.status[job] = (int32_t)fst;
                ope = 0;
                flushed = 0;
            }
#pragma unroll
            for (unsigned k = ZSK_LZ4L_DEPTH - 1; k > 0; k--) {
                qm[k] = qm[k - 1];
                ql0[k] = ql0[k - 1];
                ql1[k] = ql1[k - 1];
            }
            qm[0] = nm;
            ql0[0] = nl0;
            ql1[0] = nl1;
            slot = (slot + 1u) & (ZSK_LZ4L_SLOTS - 1u);
        }

        /* ---- 6. next job once the executor has finished the frame */
        if (phase == ZSK_L_DRAIN) {
            if (drain == 0u) phase = ZSK_L_FETCH;
            else drain--;
        }
        if (phase == ZSK_L_FETCH) {
            job = atomicAdd(a.work_counter, 1u);
            if (job >= zsk_njobs(a)) phase = ZSK_L_DONE;
            else {
                const uint32_t f = a.frame_ids ? a.frame_ids[job] : a.first_frame + job;
                const uint64_t c0 = a.c_off[f], c1 = a.c_off[f + 1], d0 = a.d_off[f], d1 = a.d_off[f + 1];
                src = a.comp + (c0 - a.comp_base);
                out = a.dst + (a.dst_offs ? a.dst_offs[job] : d0 - a.dst_base);
                cap = (uint32_t)(d1 - d0);
                n = (uint32_t)(c1 - c0);
                opp = 0;
                bend = 0;
                stop = a.limits ? a.limits[job] : 0xffffffffu;
                int hs = ZSK_ST_OK;
                if (n < 7) hs = ZSK_ST_TRUNC;
                else if (zsk_rd32(src) != ZSK_LZ4_MAGIC) hs = ZSK_ST_MAGIC;
                else {
                    const uint32_t flg = ZSK_LDG(src + 4), bd = ZSK_LDG(src + 5);
                    const uint32_t bsid = (bd >> 4) & 7;
                    if ((flg >> 6) != 1 || (flg & 0x02) || (bd & 0x8F) || bsid < 4) hs = ZSK_ST_FORMAT;
                    else {
                        flags = flg;
                        max_block = 1u << (8 + 2 * bsid);
                        ip = 6;
                        if (flg & 8) {
                            if (n - ip < 8) hs = ZSK_ST_TRUNC;
                            else { content_size = zsk_rd64(src + ip); ip += 8; }
                        }
                        if (!hs && (flg & 1)) { if (n - ip < 4) hs = ZSK_ST_TRUNC; else ip += 4; }
                        if (!hs) { if (n - ip < 1) hs = ZSK_ST_TRUNC; else if (!zsk_lz4_header_checksum_ok(src, ip)) hs = ZSK_ST_CHECKSUM; else ip += 1; }
                    }
                }
                if (hs) {
                    a.status[job] = hs; /* stays in FETCH: next trip takes the next job */
                } else {
                    sal = (uint32_t)((uintptr_t)src & 15u);
                    oal = (uint32_t)((uintptr_t)out & 31u);
                    fx = (ip + sal) & ~15u;
                    fhist = 0;
                    phase = ZSK_L_BLOCK;
                }
            }
        }
    }
    zsk_cp_wait<0>();
}

#undef ZSK_L_IW
#undef ZSK_L_OW
