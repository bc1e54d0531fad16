/*
 * zsk_zstd.cuh — K3: zstd frame decode, ONE CTA PER FRAME (replaces the reference's calls into libzstd:
 * ZSTD_decompressDCtx at reference src/decompress.c:537 and ZSTD_decompressStream at :434,:448;
 * format: RFC 8878 as restated in SURVEY.md Appendix A.2).
 *
 * CTA = 2 warps with fixed roles, so the two serial chains of a compressed block overlap:
 *
 *   warp 0 ("literals + execute")              warp 1 ("sequences")
 *   ------------------------------------       ---------------------------------------------
 *   Huffman tree description -> weights         FSE table descriptions -> LL/OF/ML decode tables
 *   parallel fill of the 2^maxBits table        (Predefined / RLE / FSE_Compressed / Repeat)
 *   lanes 0..3 decode the 4 Huffman streams
 *   ------------------------------ __syncthreads ------------------------------
 *   executes sequence chunk c-1, FOUR           lane 0 decodes sequence chunk c from the backward
 *   sequences per trip (8 lanes each) when      bitstream (3 interleaved FSE states, one 8-byte
 *   they are independent, else one by one       table load per state), resolves repeat-offsets,
 *   with 32 lanes; warp-level sync only         writes {LL, ML, offset} to smem
 *   ------------------------------ __syncthreads (per chunk) ------------------
 *
 * All tables live in shared memory (Huffman 4 KiB, FSE 5 KiB, two sequence chunks 2.25 KiB); Huffman
 * output goes to a per-CTA literal scratch in HBM (L2-resident, <= 128 KiB), the frame's output is
 * written straight to its final place and re-read from L1/L2 for match copies (offsets reach up to
 * the whole frame, far beyond what shared memory could hold for 256 KiB - 1 MiB frames).
 * Cross-block state (repeat offsets, previous Huffman table, previous FSE tables) stays in shared
 * memory for the life of the frame.  Persistent CTAs pull frames from a global atomic counter.
 */
#pragma once
#include "zsk_common.cuh"

#define ZSK_ZSTD_MAGIC 0xFD2FB528u
#define ZSK_ZSTD_CTA_THREADS 64
#ifndef ZSK_SEQ_CHUNK
#define ZSK_SEQ_CHUNK 96 /* sequences per hand-over between the two warps; 96 keeps the CTA at 18.0 KB of shared memory = 12 CTAs per SM */
#endif
#define ZSK_BLOCK_MAX (128u << 10)

struct zsk_zstd_smem {
    uint16_t huf[2048];                   /* sym | nbBits << 8 */
    uint32_t fse_ll[512];                 /* [9:0] next-state base, [15:10] nbBits, [22:16] extra bits of the code, [28:23] the code */
    uint32_t fse_ml[512];
    uint32_t fse_of[256];
    uint32_t wtab[64];                    /* FSE table of the Huffman weights: sym | nbBits << 8 | base << 16 */
    uint32_t seq[2][ZSK_SEQ_CHUNK][3];    /* literal length, match length, offset; doubles as table-build scratch in stage 1 */
    uint16_t huf_start[256];
    uint8_t weights[256];
    int16_t probs[64];                    /* sequence-table build scratch (warp 1) */
    uint16_t next[64];
    int16_t probs_w[16];                  /* Huffman-weight table build scratch (warp 0) */
    uint16_t next_w[16];
    uint32_t rep[3];
    int32_t log_ll, log_ml, log_of, log_huf; /* -1 = table not valid yet */
    int32_t err;                          /* first error of the frame; read only after __syncthreads_or */
    uint32_t bs_start;                    /* start of the sequence bitstream inside the block */
    uint32_t op;                          /* block output position handed from warp 0 to the CTA */
    uint32_t job;
};

static_assert(sizeof(((zsk_zstd_smem *)0)->seq) >= 512 * sizeof(uint32_t), "the sequence ring doubles as the 512-entry FSE table-build scratch: ZSK_SEQ_CHUNK >= 86");

/* CTA-uniform error exchange: every thread passes its own status; if any is non-zero all threads
 * get the same non-zero status back (the barrier also orders shared-memory traffic). */
static __device__ __forceinline__ int zsk_cta_status(zsk_zstd_smem &S, int st)
{
    if (st) S.err = st;
    return __syncthreads_or(st) ? S.err : ZSK_ST_OK;
}

/* ---- backward bitstream reader (see SURVEY.md Appendix A.2 "Backward bitstreams") ---- */
struct zsk_bits {
    const uint8_t *base;
    int32_t pos;   /* number of unread bits; goes negative on over-read (reads then yield zeros) */
    int32_t wbase; /* stream bit index of win bit 0 */
    uint64_t win;
};

static __device__ __forceinline__ void zsk_bits_refill(zsk_bits &b)
{
    int32_t byteoff = ((b.pos + 7) >> 3) - 8;
    b.wbase = byteoff * 8;
    if (byteoff >= 0) b.win = zsk_ld64_unaligned(b.base + byteoff);
    else if (byteoff <= -8) b.win = 0;
    else b.win = zsk_ld64_unaligned(b.base) << (unsigned)(-byteoff * 8);
}

static __device__ __forceinline__ int zsk_bits_init(zsk_bits &b, const uint8_t *p, uint32_t n)
{
    if (n == 0) return ZSK_ST_BITSTREAM;
    uint32_t last = ZSK_LDG(p + n - 1);
    if (last == 0) return ZSK_ST_BITSTREAM;
    b.base = p;
    b.pos = (int32_t)(n - 1) * 8 + (31 - __clz((int)last));
    zsk_bits_refill(b);
    return ZSK_ST_OK;
}

static __device__ __forceinline__ uint32_t zsk_bits_read(zsk_bits &b, uint32_t n) /* n <= 32 */
{
    if (b.pos - (int32_t)n < b.wbase) zsk_bits_refill(b);
    b.pos -= (int32_t)n;
    uint64_t v = b.win >> (unsigned)(b.pos - b.wbase);
    return (uint32_t)v & (uint32_t)((1ull << n) - 1);
}

/* unchecked read; the caller has called zsk_bits_ensure for the sum of the following reads */
static __device__ __forceinline__ uint32_t zsk_bits_take(zsk_bits &b, uint32_t n) /* n <= 31 */
{
    b.pos -= (int32_t)n;
    return (uint32_t)(b.win >> (unsigned)(b.pos - b.wbase)) & ((1u << n) - 1u);
}

static __device__ __forceinline__ void zsk_bits_ensure(zsk_bits &b, int32_t need) /* need <= 57 */
{
    if (b.pos - b.wbase < need) zsk_bits_refill(b);
}

static __device__ __forceinline__ int zsk_bitlen(uint32_t v) { return 32 - __clz((int)v); }

/* ---- FSE table description (forward bitstream). Single thread. Returns bytes consumed (>0) or -status. */
static __device__ int zsk_fse_read_ncount(const uint8_t *p, uint32_t n, int max_log, int max_sym, int16_t *probs,
                                          int *nsym, int *log_out)
{
    if (n == 0) return -ZSK_ST_TRUNC;
    uint32_t bitpos = 0;
    int al = 5 + (int)(zsk_ld32_unaligned(p) & 15);
    bitpos = 4;
    if (al > max_log) return -ZSK_ST_TABLE;
    int remaining = 1 << al, s = 0;
    while (remaining > 0) {
        if (s > max_sym) return -ZSK_ST_TABLE;
        int bits = zsk_bitlen((uint32_t)(remaining + 1));
        uint32_t val = (zsk_ld32_unaligned(p + (bitpos >> 3)) >> (bitpos & 7)) & ((1u << bits) - 1);
        uint32_t low = (1u << (bits - 1)) - 1;
        uint32_t thr = (1u << bits) - 1 - (uint32_t)(remaining + 1);
        if ((val & low) < thr) { val &= low; bitpos += (uint32_t)bits - 1; }
        else { if (val > low) val -= thr; bitpos += (uint32_t)bits; }
        int prob = (int)val - 1;
        probs[s++] = (int16_t)prob;
        remaining -= prob < 0 ? -prob : prob;
        if (prob == 0) {
            for (;;) {
                uint32_t rep = (zsk_ld32_unaligned(p + (bitpos >> 3)) >> (bitpos & 7)) & 3;
                bitpos += 2;
                for (uint32_t i = 0; i < rep; i++) { if (s > max_sym) return -ZSK_ST_TABLE; probs[s++] = 0; }
                if (rep != 3) break;
                if ((bitpos + 7) / 8 > n) return -ZSK_ST_TRUNC;
            }
        }
        if ((bitpos + 7) / 8 > n) return -ZSK_ST_TRUNC;
    }
    if (remaining != 0) return -ZSK_ST_TABLE;
    *nsym = s;
    *log_out = al;
    return (int)((bitpos + 7) / 8);
}

/* ---- FSE decode-table build. Single thread. tab entries: sym | nb << 8 | base << 16. */
static __device__ int zsk_fse_build(uint32_t *tab, const int16_t *probs, int nsym, int log, uint16_t *next)
{
    const int size = 1 << log;
    int high = size - 1;
    for (int s = 0; s < nsym; s++) {
        if (probs[s] == -1) { tab[high--] = (uint32_t)s; next[s] = 1; }
        else next[s] = (uint16_t)probs[s];
    }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < nsym; s++)
        for (int i = 0; i < probs[s]; i++) {
            tab[pos] = (uint32_t)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    if (pos != 0) return ZSK_ST_TABLE;
    for (int i = 0; i < size; i++) {
        uint32_t s = tab[i];
        uint32_t d = next[s]++;
        uint32_t nb = (uint32_t)(log - (zsk_bitlen(d) - 1));
        tab[i] = s | (nb << 8) | ((((d << nb) - (uint32_t)size) & 0xffffu) << 16);
    }
    return ZSK_ST_OK;
}

/* predefined distributions, RFC 8878 3.1.1.3.2.2 */
static __device__ const int16_t ZSK_LL_DEF[36] = { 4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1 };
static __device__ const int16_t ZSK_ML_DEF[53] = { 1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1 };
static __device__ const int16_t ZSK_OF_DEF[29] = { 1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1 };
static __device__ const uint32_t ZSK_LL_BASE[36] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,128,256,512,1024,2048,4096,8192,16384,32768,65536 };
static __device__ const uint8_t ZSK_LL_BITS[36] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16 };
static __device__ const uint32_t ZSK_ML_BASE[53] = { 3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,35,37,39,41,43,47,51,59,67,83,99,131,259,515,1027,2051,4099,8195,16387,32771,65539 };
static __device__ const uint8_t ZSK_ML_BITS[53] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16 };

/* Re-packs a built table (sym | nb << 8 | base << 16) into 4-byte decode entries that also carry the code and its
 * extra-bit count: the state walk needs next-state base, nbBits and the extra-bit count (one 4-byte shared-memory load per
 * state; the three fields of three entries add up without carries into each other), the extra-bit fetch looks the code's
 * base value up (zsk_code_base).  kind: 0 LL, 1 OF, 2 ML.  Executed by all 32 lanes of the building warp. */
static __device__ __forceinline__ void zsk_fse_expand(uint32_t *dst, const uint32_t *tmp, int log, int kind, unsigned lane)
{
    const uint32_t size = 1u << log;
    for (uint32_t i = lane; i < size; i += 32) {
        const uint32_t e = tmp[i], sym = e & 0xff, nb = (e >> 8) & 0xff, base = e >> 16;
        const uint32_t xb = kind == 0 ? ZSK_LL_BITS[sym] : kind == 2 ? ZSK_ML_BITS[sym] : sym;
        dst[i] = (base & 0x3ffu) | (nb << 10) | (xb << 16) | (sym << 23);
    }
}

static __device__ __forceinline__ uint32_t zsk_entry_nb(uint32_t e) { return (e >> 10) & 0x3fu; }
static __device__ __forceinline__ uint32_t zsk_entry_xb(uint32_t e) { return (e >> 16) & 0x7fu; }
static __device__ __forceinline__ uint32_t zsk_code_base(uint32_t e, int kind)
{
    const uint32_t sym = e >> 23;
    return kind == 0 ? ZSK_LL_BASE[sym] : kind == 2 ? ZSK_ML_BASE[sym] : 1u << sym;
}

/* One of the three sequence tables.  Single thread.  Advances *ip past the description. */
static __device__ int zsk_seq_table(uint32_t *tab, int32_t *log_io, int mode, const uint8_t *p, uint32_t n, uint32_t *ip,
                                    const int16_t *def, int def_n, int def_log, int max_log, int max_sym,
                                    int16_t *probs, uint16_t *next)
{
    if (mode == 0) {
        for (int i = 0; i < def_n; i++) probs[i] = def[i];
        int r = zsk_fse_build(tab, probs, def_n, def_log, next);
        if (r) return r;
        *log_io = def_log;
        return ZSK_ST_OK;
    }
    if (mode == 1) {
        if (*ip >= n) return ZSK_ST_TRUNC;
        uint32_t sym = ZSK_LDG(p + *ip);
        if ((int)sym > max_sym) return ZSK_ST_TABLE;
        *ip += 1;
        tab[0] = sym; /* 0 bits, base 0 */
        *log_io = 0;
        return ZSK_ST_OK;
    }
    if (mode == 2) {
        int nsym, log;
        int used = zsk_fse_read_ncount(p + *ip, n - *ip, max_log, max_sym, probs, &nsym, &log);
        if (used < 0) return -used;
        int r = zsk_fse_build(tab, probs, nsym, log, next);
        if (r) return r;
        *ip += (uint32_t)used;
        *log_io = log;
        return ZSK_ST_OK;
    }
    return *log_io >= 0 ? ZSK_ST_OK : ZSK_ST_TABLE; /* Repeat */
}

/* ---- Huffman tree description -> weights[] (incl. the implicit last one).  Single thread.
 * Returns bytes consumed (>0) or -status; sets *nw and *max_bits. */
static __device__ int zsk_huf_read_weights(uint8_t *w, uint16_t *huf_start, uint32_t *wtab, int16_t *probs_w, uint16_t *next_w,
                                           const uint8_t *p, uint32_t n, int *nw_out, int *max_bits_out)
{
    if (n < 1) return -ZSK_ST_TRUNC;
    int nw = 0;
    const uint32_t hb = ZSK_LDG(p);
    uint32_t used;
    if (hb >= 128) {
        nw = (int)hb - 127;
        const uint32_t bytes = ((uint32_t)nw + 1) / 2;
        if (1 + bytes > n) return -ZSK_ST_TRUNC;
        for (int i = 0; i < nw; i++) {
            uint32_t b = ZSK_LDG(p + 1 + i / 2);
            w[i] = (uint8_t)((i & 1) ? (b & 15) : (b >> 4));
        }
        used = 1 + bytes;
    } else {
        if (hb == 0 || 1 + hb > n) return -ZSK_ST_TRUNC;
        int nsym, log;
        int hdr = zsk_fse_read_ncount(p + 1, hb, 6, 12, probs_w, &nsym, &log);
        if (hdr < 0) return hdr;
        int r = zsk_fse_build(wtab, probs_w, nsym, log, next_w);
        if (r) return -r;
        if ((uint32_t)hdr >= hb) return -ZSK_ST_TRUNC;
        zsk_bits b;
        r = zsk_bits_init(b, p + 1 + hdr, hb - (uint32_t)hdr);
        if (r) return -r;
        uint32_t s1 = zsk_bits_read(b, (uint32_t)log), s2 = zsk_bits_read(b, (uint32_t)log);
        for (;;) {
            if (nw > 253) return -ZSK_ST_TABLE;
            uint32_t e1 = wtab[s1];
            w[nw++] = (uint8_t)e1;
            s1 = (e1 >> 16) + zsk_bits_read(b, (e1 >> 8) & 0xff);
            if (b.pos < 0) { w[nw++] = (uint8_t)wtab[s2]; break; }
            uint32_t e2 = wtab[s2];
            w[nw++] = (uint8_t)e2;
            s2 = (e2 >> 16) + zsk_bits_read(b, (e2 >> 8) & 0xff);
            if (b.pos < 0) { w[nw++] = (uint8_t)wtab[s1]; break; }
        }
        used = 1 + hb;
    }
    uint32_t total = 0;
    for (int i = 0; i < nw; i++) {
        if (w[i] > 12) return -ZSK_ST_TABLE;
        if (w[i]) total += 1u << (w[i] - 1);
    }
    if (total == 0) return -ZSK_ST_TABLE;
    const int max_bits = zsk_bitlen(total);
    if (max_bits > 11) return -ZSK_ST_TABLE;
    const uint32_t rest = (1u << max_bits) - total;
    if (rest == 0 || (rest & (rest - 1))) return -ZSK_ST_TABLE;
    w[nw++] = (uint8_t)zsk_bitlen(rest);
    /* table start of every symbol: codes of decreasing length first, symbols in increasing order */
    uint32_t cnt[13], idx[13];
    for (int i = 0; i < 13; i++) cnt[i] = 0;
    for (int i = 0; i < nw; i++) if (w[i]) cnt[max_bits + 1 - w[i]]++;
    idx[max_bits] = 0;
    for (int L = max_bits; L >= 1; L--) idx[L - 1] = idx[L] + cnt[L] * (1u << (max_bits - L));
    for (int s = 0; s < nw; s++) {
        if (!w[s]) continue;
        int nbits = max_bits + 1 - w[s];
        huf_start[s] = (uint16_t)idx[nbits];
        idx[nbits] += 1u << (max_bits - nbits);
    }
    *nw_out = nw;
    *max_bits_out = max_bits;
    return (int)used;
}

/* One Huffman stream, one lane: regenerates nout bytes into out. */
static __device__ int zsk_huf_stream(const uint16_t *tab, int log, const uint8_t *p, uint32_t n, uint8_t *out, uint32_t nout)
{
    zsk_bits b;
    int r = zsk_bits_init(b, p, n);
    if (r) return r;
    const uint32_t mask = (1u << log) - 1;
    uint32_t state = zsk_bits_read(b, (uint32_t)log);
    for (uint32_t i = 0; i < nout; i++) {
        uint32_t e = tab[state];
        out[i] = (uint8_t)e;
        uint32_t nb = e >> 8;
        state = ((state << nb) | zsk_bits_read(b, nb)) & mask;
    }
    return b.pos == -log ? ZSK_ST_OK : ZSK_ST_BITSTREAM;
}

/* Literal source of a block, as the executing warp sees it. */
struct zsk_lits {
    const uint8_t *ptr; /* raw bytes (compressed stream or Huffman scratch); NULL for RLE */
    uint32_t rle;
    uint32_t size;
};

static __device__ __forceinline__ void zsk_warp_literals(uint8_t *dst, const zsk_lits &L, uint32_t lpos, uint32_t n, unsigned lane)
{
    if (L.ptr) {
        if (n <= 32) { if (lane < n) dst[lane] = L.ptr[lpos + lane]; }
        else zsk_group_copy(dst, L.ptr + lpos, n, lane, 32);
    } else {
        zsk_group_fill(dst, (uint8_t)L.rle, n, lane, 32);
    }
}

/*
 * One compressed block.  p[0..n) is the block content.  All threads of the CTA call this with
 * identical arguments; returns the (CTA-uniform) status.  *pop is advanced by the block's output.
 */
static __device__ int zsk_zstd_block(zsk_zstd_smem &S, const uint8_t *__restrict__ p, uint32_t n, uint8_t *out, uint32_t *pop,
                                     uint32_t cap, uint8_t *lit_scratch, uint32_t stop)
{
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    /* ---------------- literals section header (uniform) */
    if (n < 1) return ZSK_ST_TRUNC;
    const uint32_t b0 = ZSK_LDG(p), ltype = b0 & 3, sf = (b0 >> 2) & 3;
    uint32_t regen, comp = 0, hsize, streams = 1;
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { regen = b0 >> 3; hsize = 1; }
        else if (sf == 1) { if (n < 2) return ZSK_ST_TRUNC; regen = (b0 >> 4) + ((uint32_t)ZSK_LDG(p + 1) << 4); hsize = 2; }
        else { if (n < 3) return ZSK_ST_TRUNC; regen = (b0 >> 4) + ((uint32_t)ZSK_LDG(p + 1) << 4) + ((uint32_t)ZSK_LDG(p + 2) << 12); hsize = 3; }
    } else {
        uint32_t bits;
        if (sf == 0) { hsize = 3; bits = 10; }
        else if (sf == 1) { hsize = 3; bits = 10; streams = 4; }
        else if (sf == 2) { hsize = 4; bits = 14; streams = 4; }
        else { hsize = 5; bits = 18; streams = 4; }
        if (n < hsize) return ZSK_ST_TRUNC;
        uint64_t v = 0;
        for (uint32_t i = 0; i < hsize; i++) v |= (uint64_t)ZSK_LDG(p + i) << (8 * i);
        regen = (uint32_t)(v >> 4) & ((1u << bits) - 1);
        comp = (uint32_t)(v >> (4 + bits));
    }
    if (regen > ZSK_BLOCK_MAX) return ZSK_ST_FORMAT;
    uint32_t ip = hsize; /* start of literal payload */
    zsk_lits L;
    L.size = regen; L.rle = 0; L.ptr = nullptr;
    uint32_t seq_ip; /* start of the sequences section */
    if (ltype == 0) { if (regen > n - ip) return ZSK_ST_TRUNC; L.ptr = p + ip; seq_ip = ip + regen; }
    else if (ltype == 1) { if (n - ip < 1) return ZSK_ST_TRUNC; L.rle = ZSK_LDG(p + ip); seq_ip = ip + 1; }
    else { if (comp > n - ip) return ZSK_ST_TRUNC; L.ptr = lit_scratch; seq_ip = ip + comp; }

    /* ---------------- sequences section header (uniform) */
    if (seq_ip >= n) return ZSK_ST_TRUNC;
    uint32_t sp = seq_ip;
    uint32_t nseq = ZSK_LDG(p + sp);
    sp++;
    if (nseq >= 128) {
        if (nseq == 255) { if (n - sp < 2) return ZSK_ST_TRUNC; nseq = zsk_rd16(p + sp) + 0x7F00; sp += 2; }
        else { if (n - sp < 1) return ZSK_ST_TRUNC; nseq = ((nseq - 128) << 8) + ZSK_LDG(p + sp); sp += 1; }
    }
    uint32_t modes = 0;
    if (nseq) {
        if (sp >= n) return ZSK_ST_TRUNC;
        modes = ZSK_LDG(p + sp);
        sp++;
        if (modes & 3) return ZSK_ST_FORMAT;
    }

    /* ---------------- stage 1: warp 0 regenerates literals, warp 1 builds the sequence tables */
    int st = ZSK_ST_OK;
    if (warp == 0) {
        if (ltype >= 2) {
            const uint8_t *q = p + ip;
            uint32_t qn = comp;
            int nw = 0, max_bits = 0;
            if (ltype == 2) {
                int used = 0;
                if (lane == 0) {
                    used = zsk_huf_read_weights(S.weights, S.huf_start, S.wtab, S.probs_w, S.next_w, q, qn, &nw, &max_bits);
                    if (used < 0) st = -used; else S.log_huf = max_bits;
                }
                __syncwarp(); /* weights[], huf_start[], log_huf visible to the warp */
                st = __shfl_sync(ZSK_FULL, st, 0);
                used = __shfl_sync(ZSK_FULL, used, 0);
                nw = __shfl_sync(ZSK_FULL, nw, 0);
                max_bits = __shfl_sync(ZSK_FULL, max_bits, 0);
                if (st == ZSK_ST_OK) {
                    for (int s = (int)lane; s < nw; s += 32) {
                        uint32_t w = S.weights[s];
                        if (!w) continue;
                        uint32_t nbits = (uint32_t)max_bits + 1 - w, run = 1u << (w - 1), start = S.huf_start[s];
                        uint16_t e = (uint16_t)((uint32_t)s | (nbits << 8));
                        for (uint32_t i = 0; i < run; i++) S.huf[start + i] = e;
                    }
                    q += used;
                    qn -= (uint32_t)used;
                }
                __syncwarp();
            } else if (S.log_huf < 0) st = ZSK_ST_TABLE;
            if (st == ZSK_ST_OK) {
                const int log = S.log_huf;
                if (streams == 1) {
                    if (lane == 0) st = zsk_huf_stream(S.huf, log, q, qn, lit_scratch, regen);
                } else {
                    if (qn < 6) st = ZSK_ST_TRUNC;
                    else {
                        const uint32_t s1 = zsk_rd16(q), s2 = zsk_rd16(q + 2), s3 = zsk_rd16(q + 4);
                        const uint32_t per = (regen + 3) / 4;
                        if (6 + s1 + s2 + s3 > qn) st = ZSK_ST_TRUNC;
                        else if (3 * per > regen) st = ZSK_ST_FORMAT;
                        else if (lane < 4) {
                            const uint32_t s4 = qn - 6 - s1 - s2 - s3;
                            const uint32_t soff = lane == 0 ? 0 : lane == 1 ? s1 : lane == 2 ? s1 + s2 : s1 + s2 + s3;
                            const uint32_t slen = lane == 0 ? s1 : lane == 1 ? s2 : lane == 2 ? s3 : s4;
                            const uint32_t nout = lane < 3 ? per : regen - 3 * per;
                            st = zsk_huf_stream(S.huf, log, q + 6 + soff, slen, lit_scratch + lane * per, nout);
                        }
                    }
                }
            }
        }
    } else if (nseq) {
        /* warp 1: lane 0 parses + builds each table into scratch, then all 32 lanes expand it */
        uint32_t *tmp = &S.seq[0][0][0];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int mode = (modes >> (6 - 2 * t)) & 3;
            int32_t *logp = t == 0 ? &S.log_ll : t == 1 ? &S.log_of : &S.log_ml;
            uint32_t *dst = t == 0 ? S.fse_ll : t == 1 ? S.fse_of : S.fse_ml;
            if (lane == 0 && !st) {
                if (t == 0) st = zsk_seq_table(tmp, logp, mode, p, n, &sp, ZSK_LL_DEF, 36, 6, 9, 35, S.probs, S.next);
                else if (t == 1) st = zsk_seq_table(tmp, logp, mode, p, n, &sp, ZSK_OF_DEF, 29, 5, 8, 31, S.probs, S.next);
                else st = zsk_seq_table(tmp, logp, mode, p, n, &sp, ZSK_ML_DEF, 53, 6, 9, 52, S.probs, S.next);
            }
            __syncwarp();
            st = __shfl_sync(ZSK_FULL, st, 0);
            if (!st && mode != 3) zsk_fse_expand(dst, tmp, *logp, t, lane);
            __syncwarp();
        }
        if (lane == 0) S.bs_start = sp;
    }
    if ((st = zsk_cta_status(S, st))) return st;

    /* ---------------- stage 2: sequence decode (warp 1 lane 0) overlapped with execution (warp 0) */
    uint32_t op = *pop, lpos = 0;
    if (nseq) {
        zsk_bits b;
        uint32_t sl = 0, so = 0, sm = 0;
        if (tid == 32) {
            sp = S.bs_start;
            st = sp < n ? zsk_bits_init(b, p + sp, n - sp) : ZSK_ST_TRUNC;
            if (!st) {
                sl = zsk_bits_read(b, (uint32_t)S.log_ll);
                so = zsk_bits_read(b, (uint32_t)S.log_of);
                sm = zsk_bits_read(b, (uint32_t)S.log_ml);
            }
        }
        const uint32_t nchunks = (nseq + ZSK_SEQ_CHUNK - 1) / ZSK_SEQ_CHUNK;
        for (uint32_t c = 0; c <= nchunks; c++) {
            if (warp == 1) {
                if (c < nchunks) {
                    /* Chunk c in three passes, so that the serial chain (one lane) carries only what is serial:
                     *   A  lane 0 walks the three FSE states: per sequence it records the states and the bit position,
                     *      SKIPS the extra bits (their widths come with the table entries) and reads only the state bits;
                     *   B  all 32 lanes, four sequences each: re-read the entries, fetch the extra bits at the recorded
                     *      position -> literal length, match length, offset value;
                     *   C  (warp 0, before it executes the chunk) resolves the repeat-offset codes: a 3-entry history, serial by
                     *      definition, done by the warp that would otherwise wait for this one at the chunk barrier. */
                    uint32_t (*dstq)[3] = S.seq[c & 1];
                    const uint32_t first = c * ZSK_SEQ_CHUNK;
                    const uint32_t cnt = min(nseq - first, (uint32_t)ZSK_SEQ_CHUNK);
                    if (lane == 0 && !st) {
                        /* the block's very last sequence reads no state bits: walk all but that one with the update, then it */
                        const uint32_t nwalk = (first + cnt == nseq) ? cnt - 1u : cnt;
                        uint32_t i = 0;
                        for (; i < nwalk; i++) {
                            const uint32_t el = S.fse_ll[sl], eo = S.fse_of[so], em = S.fse_ml[sm];
                            dstq[i][0] = sl | (so << 10) | (sm << 20);
                            dstq[i][1] = (uint32_t)b.pos;
                            /* e >> 10 = nbBits (6 bits) | extra bits (7 bits) << 6 | code: the sums of the three nbBits (<= 26) and of
                             * the three extra-bit counts (<= 63) stay inside their fields */
                            const uint32_t pk = (el >> 10) + (eo >> 10) + (em >> 10);
                            b.pos -= (int32_t)((pk >> 6) & 0x7fu);
                            /* the state bits of LL, ML, OF are adjacent (LL highest): one field of <= 26 bits, then split */
                            const uint32_t nm = zsk_entry_nb(em), no = zsk_entry_nb(eo), nb = pk & 0x3fu;
                            zsk_bits_ensure(b, (int32_t)nb);
                            const uint32_t v = zsk_bits_take(b, nb);
                            so = (eo & 0x3ffu) + (v & ((1u << no) - 1u));
                            sm = (em & 0x3ffu) + ((v >> no) & ((1u << nm) - 1u));
                            sl = (el & 0x3ffu) + (v >> (no + nm));
                        }
                        if (i < cnt) {
                            const uint32_t el = S.fse_ll[sl], eo = S.fse_of[so], em = S.fse_ml[sm];
                            dstq[i][0] = sl | (so << 10) | (sm << 20);
                            dstq[i][1] = (uint32_t)b.pos;
                            b.pos -= (int32_t)((((el >> 10) + (eo >> 10) + (em >> 10)) >> 6) & 0x7fu);
                        }
                        /* an over-read leaves pos negative for good (reads past the start yield zeros, states stay inside their
                         * tables), so one check per chunk is enough */
                        if (b.pos < 0) st = ZSK_ST_BITSTREAM;
                        if (!st && c + 1 == nchunks && b.pos != 0) st = ZSK_ST_BITSTREAM;
                    }
                    const int st_a = __shfl_sync(ZSK_FULL, st, 0);
                    if (!st_a) {
                        const uint8_t *bits_base = p + S.bs_start;
                        for (uint32_t i = lane; i < cnt; i += 32) {
                            const uint32_t ss = dstq[i][0];
                            const uint32_t el = S.fse_ll[ss & 1023u], eo = S.fse_of[(ss >> 10) & 1023u], em = S.fse_ml[ss >> 20];
                            zsk_bits bb;
                            bb.base = bits_base;
                            bb.pos = (int32_t)dstq[i][1];
                            zsk_bits_refill(bb);                       /* >= 57 bits below pos */
                            const uint32_t xo = zsk_entry_xb(eo), xm = zsk_entry_xb(em), xl = zsk_entry_xb(el);
                            const uint32_t ov = zsk_code_base(eo, 1) + zsk_bits_take(bb, xo);
                            zsk_bits_ensure(bb, (int32_t)(xm + xl));
                            const uint32_t mlen = zsk_code_base(em, 2) + zsk_bits_take(bb, xm);
                            const uint32_t llen = zsk_code_base(el, 0) + zsk_bits_take(bb, xl);
                            dstq[i][0] = llen; dstq[i][1] = mlen; dstq[i][2] = ov;
                        }
                    }
                }
            } else if (c > 0) {
                /* warp 0 executes chunk c-1 four sequences per trip: lane group g (8 lanes) takes sequence 4t+g.
                 * Fast trip (all four): literal run <= 16, match <= 32 and every match source ends before the
                 * first byte this trip writes -> all loads of the four sequences are in flight together, then
                 * all stores.  Otherwise the four are executed one after another by the whole warp. */
                uint32_t (*q)[3] = S.seq[(c - 1) & 1];
                const uint32_t first = (c - 1) * ZSK_SEQ_CHUNK;
                const uint32_t cnt = min(nseq - first, (uint32_t)ZSK_SEQ_CHUNK);
                const unsigned g = lane >> 3, gl = lane & 7;
                if (lane == 0) { /* pass C: offset value -> offset (repeat-offset history) */
                    uint32_t r0 = S.rep[0], r1 = S.rep[1], r2 = S.rep[2];
                    for (uint32_t i = 0; i < cnt; i++) {
                        const uint32_t ov = q[i][2];
                        uint32_t offset;
                        if (ov > 3) { offset = ov - 3; r2 = r1; r1 = r0; r0 = offset; }
                        else {
                            const uint32_t idx = ov - 1 + (q[i][0] == 0);
                            if (idx == 0) offset = r0;
                            else {
                                offset = idx == 1 ? r1 : idx == 2 ? r2 : r0 - 1;
                                if (offset == 0) { st = ZSK_ST_OFFSET; break; }
                                if (idx > 1) r2 = r1;
                                r1 = r0;
                                r0 = offset;
                            }
                        }
                        q[i][2] = offset;
                    }
                    S.rep[0] = r0; S.rep[1] = r1; S.rep[2] = r2;
                }
                {
                    const int st_c = __shfl_sync(ZSK_FULL, st, 0);
                    if (st_c) st = st_c;
                }
                for (uint32_t t = 0; t < cnt && !st; t += 4) {
                    const bool act = t + g < cnt;
                    uint32_t llen = 0, mlen = 0, offset = 1;
                    if (act) { llen = q[t + g][0]; mlen = q[t + g][1]; offset = q[t + g][2]; }
                    const uint32_t len = llen + mlen;
                    const uint32_t n0 = __shfl_sync(ZSK_FULL, len, 0), n1 = __shfl_sync(ZSK_FULL, len, 8), n2 = __shfl_sync(ZSK_FULL, len, 16),
                                   n3 = __shfl_sync(ZSK_FULL, len, 24);
                    const uint32_t l0 = __shfl_sync(ZSK_FULL, llen, 0), l1 = __shfl_sync(ZSK_FULL, llen, 8), l2 = __shfl_sync(ZSK_FULL, llen, 16),
                                   l3 = __shfl_sync(ZSK_FULL, llen, 24);
                    const uint32_t rel = (g > 0 ? n0 : 0) + (g > 1 ? n1 : 0) + (g > 2 ? n2 : 0);      /* output start inside the trip */
                    const uint32_t lrel = (g > 0 ? l0 : 0) + (g > 1 ? l1 : 0) + (g > 2 ? l2 : 0);    /* literal start inside the trip */
                    const uint32_t tot = n0 + n1 + n2 + n3, ltot = l0 + l1 + l2 + l3;
                    bool ok = llen <= 16 && mlen <= 32 && offset >= rel + len && offset <= op + rel + llen;
                    ok = ok && ltot <= regen - lpos && tot <= cap - op;
                    if (__all_sync(ZSK_FULL, ok)) {
                        uint8_t *o = out + op + rel + gl;
                        const uint8_t *m = o + llen - offset;
                        uint32_t lv[2], mv[4];
                        if (L.ptr) {
                            const uint8_t *lp = L.ptr + lpos + lrel + gl;
#pragma unroll
                            for (unsigned k = 0; k < 2; k++) lv[k] = (gl + 8 * k < llen) ? lp[8 * k] : 0;
                        } else {
                            lv[0] = lv[1] = L.rle;
                        }
#pragma unroll
                        for (unsigned k = 0; k < 4; k++) mv[k] = (gl + 8 * k < mlen) ? m[8 * k] : 0;
#pragma unroll
                        for (unsigned k = 0; k < 2; k++) if (gl + 8 * k < llen) o[8 * k] = (uint8_t)lv[k];
                        o += llen;
#pragma unroll
                        for (unsigned k = 0; k < 4; k++) if (gl + 8 * k < mlen) o[8 * k] = (uint8_t)mv[k];
                        op += tot;
                        lpos += ltot;
                        __syncwarp();
                    } else {
                        const uint32_t k_end = min(4u, cnt - t);
                        for (uint32_t k = 0; k < k_end; k++) {
                            const uint32_t ll1 = __shfl_sync(ZSK_FULL, llen, 8 * k), ml1 = __shfl_sync(ZSK_FULL, mlen, 8 * k),
                                           of1 = __shfl_sync(ZSK_FULL, offset, 8 * k);
                            if (ll1 > regen - lpos) { st = ZSK_ST_FORMAT; break; }
                            if (ll1 > cap - op || ml1 > cap - op - ll1) { st = ZSK_ST_DST; break; }
                            if (ll1) zsk_warp_literals(out + op, L, lpos, ll1, lane);
                            op += ll1; lpos += ll1;
                            if (of1 > op) { st = ZSK_ST_OFFSET; break; }
                            __syncwarp();
                            zsk_warp_match(out, op, of1, ml1, lane);
                            __syncwarp();
                            op += ml1;
                        }
                    }
                }
                if (!st && op >= stop) st = ZSK_ST_STOPPED; /* the caller needs no byte beyond `stop` of this frame */
            }
            if ((st = zsk_cta_status(S, st))) return st;
        }
    }
    /* ---------------- literals after the last sequence (warp 0; op/lpos are only tracked there) */
    if (warp == 0) {
        const uint32_t rest = regen - lpos;
        if (rest > cap - op) st = ZSK_ST_DST;
        else if (rest) zsk_warp_literals(out + op, L, lpos, rest, lane);
        if (lane == 0) S.op = op + rest;
    }
    if ((st = zsk_cta_status(S, st))) return st;
    op = S.op;
    if (op - *pop > ZSK_BLOCK_MAX) return ZSK_ST_FORMAT;
    *pop = op;
    return ZSK_ST_OK;
}

/* One complete zstd frame; all CTA threads call with identical arguments. */
static __device__ int zsk_zstd_frame(zsk_zstd_smem &S, const uint8_t *__restrict__ src, uint32_t n, uint8_t *out, uint32_t cap,
                                     uint32_t *produced, uint8_t *lit_scratch, uint32_t stop)
{
    const unsigned tid = threadIdx.x;
    if (n < 6) return ZSK_ST_TRUNC;
    if (zsk_rd32(src) != ZSK_ZSTD_MAGIC) return ZSK_ST_MAGIC;
    const uint32_t fhd = ZSK_LDG(src + 4);
    const uint32_t fcs_flag = fhd >> 6, ss = (fhd >> 5) & 1, cksum = (fhd >> 2) & 1, did = fhd & 3;
    if (fhd & 0x08) return ZSK_ST_FORMAT;
    uint32_t ip = 5;
    if (!ss) ip += 1; /* window descriptor: the frame's own output is the window */
    const uint32_t did_sz = did == 3 ? 4 : did;
    if (did_sz) {
        if (n - ip < did_sz) return ZSK_ST_TRUNC;
        uint32_t id = 0;
        for (uint32_t i = 0; i < did_sz; i++) id |= (uint32_t)ZSK_LDG(src + ip + i) << (8 * i);
        if (id) return ZSK_ST_UNSUPPORTED;
        ip += did_sz;
    }
    const uint32_t fcs_sz = fcs_flag == 0 ? ss : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
    if (n < ip || n - ip < fcs_sz) return ZSK_ST_TRUNC;
    uint64_t fcs = 0;
    for (uint32_t i = 0; i < fcs_sz; i++) fcs |= (uint64_t)ZSK_LDG(src + ip + i) << (8 * i);
    if (fcs_sz == 2) fcs += 256;
    ip += fcs_sz;

    if (tid == 0) {
        S.rep[0] = 1; S.rep[1] = 4; S.rep[2] = 8;
        S.log_ll = S.log_ml = S.log_of = S.log_huf = -1;
        S.err = 0;
    }
    __syncthreads();
    uint32_t op = 0;
    for (;;) {
        if (n - ip < 3) return ZSK_ST_TRUNC;
        const uint32_t bh = zsk_rd24(src + ip);
        ip += 3;
        const uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
        if (type == 0) {
            if (bsize > n - ip) return ZSK_ST_TRUNC;
            if (bsize > cap - op) return ZSK_ST_DST;
            zsk_group_copy(out + op, src + ip, bsize, tid, blockDim.x);
            op += bsize; ip += bsize;
        } else if (type == 1) {
            if (n - ip < 1) return ZSK_ST_TRUNC;
            if (bsize > cap - op) return ZSK_ST_DST;
            zsk_group_fill(out + op, ZSK_LDG(src + ip), bsize, tid, blockDim.x);
            op += bsize; ip += 1;
        } else if (type == 2) {
            if (bsize > n - ip) return ZSK_ST_TRUNC;
            if (bsize > ZSK_BLOCK_MAX) return ZSK_ST_FORMAT;
            int st = zsk_zstd_block(S, src + ip, bsize, out, &op, cap, lit_scratch, stop);
            if (st) return st;
            ip += bsize;
        } else return ZSK_ST_FORMAT;
        __syncthreads(); /* this block's output is visible to whoever reads it as match source next */
        if (last) break;
        if (op >= stop) return ZSK_ST_STOPPED;
    }
    if (cksum && n - ip < 4) return ZSK_ST_TRUNC;
    if (fcs_sz && fcs != op) return ZSK_ST_FORMAT;
    *produced = op;
    if (cksum) { /* content checksum: low 32 bits of XXH64 over the decoded frame (the last block ended with a CTA barrier) */
        int st = ZSK_ST_OK;
        if (tid < 4 && (uint32_t)zsk_xxh64_group4(out, op, tid, 0xFu) != zsk_rd32(src + ip)) st = ZSK_ST_CHECKSUM;
        return zsk_cta_status(S, st);
    }
    return ZSK_ST_OK;
}

#ifndef ZSK_ZSTD_MIN_CTAS
#define ZSK_ZSTD_MIN_CTAS 16 /* 13.6 KB of shared memory per CTA allow 16 CTAs per SM; this sizes the register allocation for them (64 per thread) */
#endif
__global__ void __launch_bounds__(ZSK_ZSTD_CTA_THREADS, ZSK_ZSTD_MIN_CTAS) zsk_zstd_decode_kernel(zsk_decode_args a)
{
    __shared__ zsk_zstd_smem S;
    uint8_t *lit_scratch = a.scratch + (size_t)blockIdx.x * ZSK_LIT_SCRATCH + 16;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) S.job = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        uint32_t job = S.job;
        if (a.job_list) { /* the pipeline's deferred frames only */
            if (job >= (uint32_t)*a.job_list_count) break;
            job = a.job_list[job];
        } else if (job >= zsk_njobs(a)) break;
        const uint32_t f = a.frame_ids ? a.frame_ids[job] : a.first_frame + job;
        const uint64_t c0 = a.c_off[f], c1 = a.c_off[f + 1], d0 = a.d_off[f], d1 = a.d_off[f + 1];
        const uint8_t *src = a.comp + (c0 - a.comp_base);
        uint8_t *out = a.dst + (a.dst_offs ? a.dst_offs[job] : d0 - a.dst_base);
        const uint32_t cap = (uint32_t)(d1 - d0);
        uint32_t produced = 0;
        const uint32_t stop = a.limits ? a.limits[job] : 0xffffffffu;
        int st = zsk_zstd_frame(S, src, (uint32_t)(c1 - c0), out, cap, &produced, lit_scratch, stop);
        if (st == ZSK_ST_STOPPED) st = ZSK_ST_OK;
        else if (st == ZSK_ST_OK && produced != cap) st = ZSK_ST_SIZE;
        if (threadIdx.x == 0) a.status[job] = st;
    }
}
