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

/* LZ4 frame header checksum: second byte of XXH32 over the descriptor (FLG, BD, optional content size and dictID),
 * i.e. bytes [4, hc_pos) of the frame; liblz4 always verifies it ("ERROR_headerChecksum_invalid"). */
static __device__ __forceinline__ bool zsk_lz4_header_checksum_ok(const uint8_t *src, uint32_t hc_pos)
{
    return ((zsk_xxh32_serial(src + 4, hc_pos - 4) >> 8) & 0xffu) == (uint32_t)ZSK_LDG(src + hc_pos);
}

#define ZSK_LZ4_CTA_THREADS 128

/*
 * Warp-per-frame variant with speculative parallel parsing ("batch" kernel).
 *
 * Motivation: the lock-step kernel keeps 28,416 frames in flight and advances each by one
 * sequence per DRAM round trip; the L2 then holds only the last few KiB of every frame's output, 80 % of the
 * match loads miss to DRAM and the kernel ends up bound by random DRAM sector reads.  This kernel gives a frame
 * a whole warp and advances it by up to EIGHT sequences per trip, so four times fewer frames are in flight at
 * the same occupancy and each frame moves ~30x faster — its 64 KiB match window stays in L2.
 *
 * One trip in the sequence state:
 *   1. two coalesced loads put 64 bytes of the compressed stream into registers (lane i: bytes ip+i, ip+32+i);
 *   2. every lane parses the token that WOULD start at its byte: lengths, offset (bytes fetched from the window
 *      registers by shuffles), position of the following token;
 *   3. the warp walks the true token chain from window position 0 (one shuffle per sequence, <= 8 hops);
 *   4. sequence k goes to the 4-lane group k; a 3-step segmented scan gives every group its output position;
 *   5. if every match source of the batch ends before the batch's first output byte, all literal and match
 *      loads are issued, then all stores (the batch's output is contiguous: coalesced stores);
 *      otherwise only the independent prefix is executed, or — for a token with length-extension bytes, an
 *      overlapping match, the block's last sequence — one sequence is done by the whole warp (general path).
 */
enum { ZSK_LZ4_W_FETCH = 0, ZSK_LZ4_W_BLOCK = 1, ZSK_LZ4_W_SEQ = 2 };

__global__ void __launch_bounds__(ZSK_LZ4_CTA_THREADS) zsk_lz4_decode_batch_kernel(zsk_decode_args a)
{
    __shared__ uint8_t s_pos[ZSK_LZ4_CTA_THREADS / 32][8];
    const unsigned lane = threadIdx.x & 31;
    const unsigned g = lane >> 2, q = lane & 3; /* 4-lane group index, lane inside the group */
    int state = ZSK_LZ4_W_FETCH;
    const uint8_t *src = nullptr;
    uint8_t *out = nullptr;
    uint32_t n = 0, ip = 0, bend = 0, op = 0, cap = 0, job = 0, flags = 0, max_block = 0, stop = 0xffffffffu, bstart = 0;
    uint64_t content_size = 0;
    for (;;) {
        __syncwarp(); /* orders the previous trip's stores before this trip's loads */
        int st = ZSK_ST_OK;
        bool frame_end = false;
        if (state == ZSK_LZ4_W_SEQ) {
            /* ---- 1. window */
            const uint8_t *wp = src + ip + lane;
            const uint32_t b0 = ZSK_LDG(wp), b1 = ZSK_LDG(wp + 32);
            /* ---- 2. speculative parse of the token at window position `lane` */
            const uint32_t cll = b0 >> 4;
            const uint32_t opos = lane + 1 + cll;                 /* window position of the offset (<= 45 when cll <= 13) */
            const uint32_t i0 = opos & 31, i1 = (opos + 1) & 31, i2 = (opos + 2) & 31;
            const uint32_t lo0 = __shfl_sync(ZSK_FULL, b0, i0), lo1 = __shfl_sync(ZSK_FULL, b1, i0);
            const uint32_t hi0 = __shfl_sync(ZSK_FULL, b0, i1), hi1 = __shfl_sync(ZSK_FULL, b1, i1);
            const uint32_t ex0 = __shfl_sync(ZSK_FULL, b0, i2), ex1 = __shfl_sync(ZSK_FULL, b1, i2);
            const uint32_t coff = ((opos < 32) ? lo0 : lo1) | (((opos + 1 < 32) ? hi0 : hi1) << 8);
            const uint32_t ext = (opos + 2 < 32) ? ex0 : ex1;    /* first match-length extension byte, if the nibble is 15 */
            const bool longm = (b0 & 15) == 15;
            const uint32_t cml = (b0 & 15) + 4 + (longm ? ext : 0);
            /* simple = literal run <= 13, match <= 50 (at most one extension byte), not the block's last sequence */
            const uint32_t cnx = opos + 2 + (longm ? 1 : 0);      /* window position of the following token (<= 48) */
            const bool simple = cll <= 13 && cml <= 50 && ip + cnx < bend && coff != 0;
            const uint32_t cnext = simple ? cnx : 255u;
            /* ---- 3. token chain by pointer doubling instead of a serial walk: R = window positions reachable from
             *         position 0 in <= 7 hops (three rounds: one warp-wide OR + one shuffle each).  A non-simple
             *         token has no successor, so the chain stops there; the executable sequences are the simple
             *         tokens in R, numbered by their rank. */
            uint32_t jump = cnext;                                 /* 2^k-th successor of my position */
            uint32_t R = 1u;
#pragma unroll
            for (unsigned k = 0; k < 3; k++) {
                const bool in = (R >> lane) & 1u;
                R |= __reduce_or_sync(ZSK_FULL, (in && jump < 32) ? (1u << jump) : 0u);
                if (k < 2) {
                    const uint32_t j2 = __shfl_sync(ZSK_FULL, jump, jump & 31);
                    jump = (jump < 32) ? j2 : 255u;
                }
            }
            const uint32_t E = R & __ballot_sync(ZSK_FULL, simple); /* token positions of the executable sequences */
            const uint32_t K = (uint32_t)__popc(E);
            const uint32_t rank = (uint32_t)__popc(E & ((1u << lane) - 1u));
            if ((E >> lane) & 1u) s_pos[threadIdx.x >> 5][rank] = (uint8_t)lane;
            __syncwarp();
            const uint32_t myL = s_pos[threadIdx.x >> 5][g];       /* window position of sequence g (garbage if g >= K) */
            const uint32_t endpos = K ? __shfl_sync(ZSK_FULL, cnext, 31 - __clz((int)E)) : 0u; /* token after the last one */
            /* ---- 4. per-group parameters and output positions */
            const uint32_t pk = cll | (cml << 8) | (coff << 16);
            const uint32_t mine = __shfl_sync(ZSK_FULL, pk, myL);
            const uint32_t ll = mine & 0xff, ml = (mine >> 8) & 0xff, off = mine >> 16;
            const bool active = g < K;
            const uint32_t len = active ? ll + ml : 0;
            uint32_t inc = len;
            uint32_t t_;
            t_ = __shfl_up_sync(ZSK_FULL, inc, 4);  if (g >= 1) inc += t_;
            t_ = __shfl_up_sync(ZSK_FULL, inc, 8);  if (g >= 2) inc += t_;
            t_ = __shfl_up_sync(ZSK_FULL, inc, 16); if (g >= 4) inc += t_;
            const uint32_t start = inc - len;                     /* output start of my sequence inside the batch */
            /* ---- 5. independence: my match source must end before the batch's first output byte */
            const bool good = inc <= cap - op && off >= start + len && off <= op + start + ll;
            const uint32_t bad = __ballot_sync(ZSK_FULL, active && !good);
            const uint32_t Kx = bad ? (uint32_t)(__ffs((int)bad) - 1) >> 2 : K; /* sequences executed in parallel this trip */
            if (Kx > 0) {
                /* every output byte j of my sequence comes from the literal run (j < ll) or from `off` bytes back
                 * (j >= ll); pass k handles bytes q + 4k; the pass count is the batch maximum (warp-uniform) */
                const bool run = g < Kx;
                const uint32_t mylen = run ? len : 0;
                const uint32_t npass = (__reduce_max_sync(ZSK_FULL, mylen) + 3) >> 2;   /* <= 16 */
                uint8_t *o = out + op + start + q;
                const uint8_t *lp = src + ip + myL + 1 + q;
                const uint8_t *mp = o - off;
                uint32_t v[6];
#pragma unroll
                for (unsigned k = 0; k < 4; k++) { /* 16 bytes per sequence: always issued */
                    const uint32_t j = q + 4 * k;
                    v[k] = 0;
                    if (j < mylen) v[k] = (j < ll) ? (uint32_t)ZSK_LDG(lp + 4 * k) : (uint32_t)mp[4 * k];
                }
                if (npass > 4) {
#pragma unroll
                    for (unsigned k = 4; k < 6; k++) {
                        const uint32_t j = q + 4 * k;
                        v[k] = 0;
                        if (j < mylen) v[k] = (j < ll) ? (uint32_t)ZSK_LDG(lp + 4 * k) : (uint32_t)mp[4 * k];
                    }
                }
#pragma unroll
                for (unsigned k = 0; k < 4; k++)
                    if (q + 4 * k < mylen) o[4 * k] = (uint8_t)v[k];
                if (npass > 4) {
#pragma unroll
                    for (unsigned k = 4; k < 6; k++)
                        if (q + 4 * k < mylen) o[4 * k] = (uint8_t)v[k];
                    for (uint32_t k = 6; k < npass; k++) { /* rare: a sequence longer than 24 bytes */
                        const uint32_t j = q + 4 * k;
                        if (j < mylen) o[4 * k] = (j < ll) ? ZSK_LDG(lp + 4 * k) : mp[4 * k];
                    }
                }
                /* advance past the executed sequences: output end and window position of the next token */
                const uint32_t last = 4 * (Kx - 1);
                const uint32_t tot = __shfl_sync(ZSK_FULL, inc, last);
                const uint32_t lastL = __shfl_sync(ZSK_FULL, myL, last);
                const uint32_t nxl = __shfl_sync(ZSK_FULL, cnext, lastL);
                const uint32_t adv = (Kx == K) ? endpos : nxl;
                op += tot;
                ip += adv;
            } else {
                /* ---- general path: ONE sequence by the whole warp (length extensions, overlap, last sequence) */
                const uint32_t tok = __shfl_sync(ZSK_FULL, b0, 0);
                uint32_t gll = tok >> 4, gml = (tok & 15) + 4;
                ip++;
                if (gll == 15) {
                    uint32_t b;
                    do {
                        if (ip >= bend) { st = ZSK_ST_TRUNC; break; }
                        b = ZSK_LDG(src + ip);
                        ip++;
                        gll += b;
                    } while (b == 255);
                }
                if (!st && gll > bend - ip) st = ZSK_ST_TRUNC;
                if (!st && gll > cap - op) st = ZSK_ST_DST;
                if (!st) {
                    zsk_group_copy(out + op, src + ip, gll, lane, 32);
                    ip += gll;
                    op += gll;
                    if (ip == bend) {
                        state = ZSK_LZ4_W_BLOCK; /* last sequence of the block: literals only */
                    } else if (bend - ip < 2) {
                        st = ZSK_ST_TRUNC;
                    } else {
                        const uint32_t goff = zsk_rd16(src + ip);
                        ip += 2;
                        if (gml == 19) {
                            uint32_t b;
                            do {
                                if (ip >= bend) { st = ZSK_ST_TRUNC; break; }
                                b = ZSK_LDG(src + ip);
                                ip++;
                                gml += b;
                            } while (b == 255);
                        }
                        if (!st && (goff == 0 || goff > op)) st = ZSK_ST_OFFSET;
                        if (!st && gml > cap - op) st = ZSK_ST_DST;
                        if (!st) {
                            __syncwarp();
                            zsk_group_match<32>(out, op, goff, gml, lane, ZSK_FULL);
                            op += gml;
                            if (ip >= bend) st = ZSK_ST_TRUNC; /* a block never ends with a match */
                        }
                    }
                }
            }
            if (!st && op >= stop) frame_end = true; /* the caller needs no byte beyond `stop` of this frame */
        } else if (state == ZSK_LZ4_W_BLOCK) {
            if (bend && (flags & 16)) { /* block checksum after the compressed block just finished: XXH32 of its compressed bytes */
                if (n - ip < 4) st = ZSK_ST_TRUNC;
                else if (zsk_xxh32_group4(src + bstart, bend - bstart, lane & 3, ZSK_FULL) != zsk_rd32(src + ip)) st = ZSK_ST_CHECKSUM;
                else ip += 4;
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
                    else if (flags & 4) { /* content checksum: XXH32 of the decoded frame (this warp's stores were ordered by the trip barrier) */
                        if (zsk_xxh32_group4(out, op, lane & 3, ZSK_FULL) != zsk_rd32(src + ip)) st = ZSK_ST_CHECKSUM;
                    }
                    frame_end = true;
                } else {
                    const bool raw = bs >> 31;
                    bs &= 0x7FFFFFFFu;
                    if (bs > max_block) st = ZSK_ST_FORMAT;
                    else if (bs > n - ip) st = ZSK_ST_TRUNC;
                    else if (raw) {
                        if (bs > cap - op) st = ZSK_ST_DST;
                        else {
                            zsk_group_copy(out + op, src + ip, bs, lane, 32);
                            op += bs;
                            ip += bs;
                            if (flags & 16) {
                                if (n - ip < 4) st = ZSK_ST_TRUNC;
                                else if (zsk_xxh32_group4(src + ip - bs, bs, lane & 3, ZSK_FULL) != zsk_rd32(src + ip)) st = ZSK_ST_CHECKSUM;
                                else ip += 4;
                            }
                            if (!st && op >= stop) frame_end = true;
                        }
                    } else if (bs == 0) {
                        st = ZSK_ST_FORMAT;
                    } else {
                        bstart = ip;
                        bend = ip + bs;
                        state = ZSK_LZ4_W_SEQ;
                    }
                }
            }
        } else { /* ZSK_LZ4_W_FETCH */
            if (lane == 0) job = atomicAdd(a.work_counter, 1u);
            job = __shfl_sync(ZSK_FULL, job, 0);
            if (job >= zsk_njobs(a)) break;
            const uint32_t f = a.frame_ids ? a.frame_ids[job] : a.first_frame + job;
            const uint64_t c0 = a.c_off[f], c1 = a.c_off[f + 1], d0 = a.d_off[f], d1 = a.d_off[f + 1];
            src = a.comp + (c0 - a.comp_base);
            out = a.dst + (a.dst_offs ? a.dst_offs[job] : d0 - a.dst_base);
            cap = (uint32_t)(d1 - d0);
            n = (uint32_t)(c1 - c0);
            op = 0;
            bend = 0;
            stop = a.limits ? a.limits[job] : 0xffffffffu;
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
                    if (!st) { if (n - ip < 1) st = ZSK_ST_TRUNC; else if (!zsk_lz4_header_checksum_ok(src, ip)) st = ZSK_ST_CHECKSUM; else ip += 1; }
                    if (!st) state = ZSK_LZ4_W_BLOCK;
                }
            }
            if (st) frame_end = true;
        }
        if (st || frame_end) {
            if (lane == 0) a.status[job] = st;
            state = ZSK_LZ4_W_FETCH;
        }
    }
}
