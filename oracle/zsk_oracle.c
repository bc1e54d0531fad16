/*
 * zsk_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle, "port" kind).
 *
 * A plain-C restatement of the algorithm on libzseek's seekable-format READ path.  It is the checker
 * for the CUDA kernels: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product library never links, loads or calls it.
 *
 * What it restates, and from where:
 *   - seek-table parse + prefix sums ........ reference src/seek_table.c:62-176
 *   - offset -> frame binary search ......... reference src/seek_table.c:187-202
 *   - zseek_pread result semantics (B1-B4) .. reference src/decompress.c:470-574, 685-804
 *   - LZ4 frame/block decode ................ THIRD PARTY: liblz4 (>= 1.8.3 per reference
 *       meson.build:11; 1.9.4 installed), reached by the reference at src/decompress.c:631,653,762
 *       (LZ4F_decompress).  Not vendored under /root/reference, so the public LZ4 frame + block
 *       format is restated here.
 *   - zstd frame decode ..................... THIRD PARTY: libzstd (>= 1.4.9 per reference
 *       meson.build:10; 1.5.5 installed), reached at src/decompress.c:434,448,537
 *       (ZSTD_decompressStream / ZSTD_decompressDCtx).  Restated from RFC 8878 (SURVEY.md
 *       Appendix A.2).
 *
 * PINNING: parity is pinned.  tests/test_oracle.py checks this file (a) against the committed golden
 * fixtures in tests/golden/ — files written by the reference writer plus the byte ranges the
 * reference's own zseek_pread returned for them — and (b), when oracle/_ref/libzseek_ref.so is
 * present, live against the reference reader on freshly written files of every writer mode.
 * The reference's only in-tree result check for this path is the round-trip memcmp of
 * test/example.c:82-86; that property (decode == original input) is asserted too.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#define EXPORT __attribute__((visibility("default")))

enum {
    ZO_OK = 0,
    ZO_ERR_TRUNC = -1,      /* input ends early */
    ZO_ERR_MAGIC = -2,
    ZO_ERR_FORMAT = -3,     /* reserved bits / impossible field */
    ZO_ERR_DST = -4,        /* output would exceed dst capacity */
    ZO_ERR_OFFSET = -5,     /* match offset reaches before the frame start */
    ZO_ERR_BITSTREAM = -6,  /* backward bitstream not exactly consumed / corrupt */
    ZO_ERR_TABLE = -7,      /* bad FSE / Huffman description */
    ZO_ERR_UNSUPPORTED = -8, /* dictionary id etc. */
    ZO_ERR_CHECKSUM = -9    /* header / block / content checksum mismatch (liblz4 and libzstd verify them, so the reference fails) */
};

static uint32_t rd_le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t rd_le64(const uint8_t *p) { return (uint64_t)rd_le32(p) | ((uint64_t)rd_le32(p + 4) << 32); }

/* XXH32 / XXH64, seed 0 (public xxHash specification): the checksums of the LZ4 frame format (header, block, content)
 * and of zstd frames (content: low 32 bits of XXH64).  liblz4 / libzstd verify them inside LZ4F_decompress /
 * ZSTD_decompressDCtx (call sites reference src/decompress.c:537,762), so a mismatch fails the reference's zseek_pread. */
static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint32_t xxh32(const uint8_t *p, size_t len)
{
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    const uint8_t *end = p + len;
    uint32_t h;
    if (len >= 16) {
        uint32_t v[4] = { P1 + P2, P2, 0, 0u - P1 };
        for (; p + 16 <= end; p += 16)
            for (int k = 0; k < 4; k++) v[k] = rotl32(v[k] + rd_le32(p + 4 * k) * P2, 13) * P1;
        h = rotl32(v[0], 1) + rotl32(v[1], 7) + rotl32(v[2], 12) + rotl32(v[3], 18);
    } else h = P5;
    h += (uint32_t)len;
    for (; p + 4 <= end; p += 4) h = rotl32(h + rd_le32(p) * P3, 17) * P4;
    for (; p < end; p++) h = rotl32(h + *p * P5, 11) * P1;
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}
static uint64_t xxh64_round(uint64_t acc, uint64_t w) { return rotl64(acc + w * 14029467366897019727ull, 31) * 11400714785074694791ull; }
static uint64_t xxh64(const uint8_t *p, size_t len)
{
    const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                   P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v[4] = { P1 + P2, P2, 0, 0ull - P1 };
        for (; p + 32 <= end; p += 32)
            for (int k = 0; k < 4; k++) v[k] = xxh64_round(v[k], rd_le64(p + 8 * k));
        h = rotl64(v[0], 1) + rotl64(v[1], 7) + rotl64(v[2], 12) + rotl64(v[3], 18);
        for (int k = 0; k < 4; k++) h = (h ^ xxh64_round(0, v[k])) * P1 + P4;
    } else h = P5;
    h += (uint64_t)len;
    for (; p + 8 <= end; p += 8) h = rotl64(h ^ xxh64_round(0, rd_le64(p)), 27) * P1 + P4;
    if (p + 4 <= end) { h = rotl64(h ^ ((uint64_t)rd_le32(p) * P1), 23) * P2 + P3; p += 4; }
    for (; p < end; p++) h = rotl64(h ^ (*p * P5), 11) * P1;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* =====================================================================================
 * Seek table — reference src/seek_table.c:112-176 (read_seek_table), :62-110 (read_st_entries)
 * ===================================================================================== */
typedef struct {
    uint64_t n;      /* frames */
    uint64_t *c_off; /* n+1 compressed prefix offsets  */
    uint64_t *d_off; /* n+1 decompressed prefix offsets */
    int checksum_flag;
} zo_seek_table;

EXPORT void zo_seek_table_free(zo_seek_table *st)
{
    if (!st) return;
    free(st->c_off);
    free(st->d_off);
    free(st);
}

EXPORT zo_seek_table *zo_seek_table_parse(const uint8_t *img, size_t size)
{
    if (size < 9) return NULL;                                   /* footer pread comes back short */
    const uint8_t *footer = img + size - 9;
    if (rd_le32(footer + 5) != 0x8F92EAB1u) return NULL;          /* Seekable_Magic_Number */
    uint8_t desc = footer[4];
    if (desc & 0x7c) return NULL;                                 /* reserved bits */
    int checksum = (desc & 0x80) != 0;
    uint64_t n = rd_le32(footer);
    uint64_t es = checksum ? 12 : 8;
    uint64_t frame_size = 8 + n * es + 9;
    if (frame_size > size) return NULL;                           /* header pread fails */
    const uint8_t *hdr = img + size - frame_size;
    if (rd_le32(hdr) != 0x184D2A5Eu) return NULL;                 /* skippable magic | 0xE */
    if (rd_le32(hdr + 4) != (uint32_t)(frame_size - 8)) return NULL;
    zo_seek_table *st = calloc(1, sizeof(*st));
    st->n = n;
    st->checksum_flag = checksum;
    st->c_off = malloc((n + 1) * sizeof(uint64_t));
    st->d_off = malloc((n + 1) * sizeof(uint64_t));
    uint64_t c = 0, d = 0;
    const uint8_t *e = hdr + 8;
    for (uint64_t i = 0; i < n; i++, e += es) {
        st->c_off[i] = c;
        st->d_off[i] = d;
        c += rd_le32(e);
        d += rd_le32(e + 4);
    }
    st->c_off[n] = c;
    st->d_off[n] = d;
    return st;
}

EXPORT uint64_t zo_seek_table_frames(const zo_seek_table *st) { return st->n; }
EXPORT const uint64_t *zo_seek_table_coff(const zo_seek_table *st) { return st->c_off; }
EXPORT const uint64_t *zo_seek_table_doff(const zo_seek_table *st) { return st->d_off; }

/* reference src/seek_table.c:187-202: largest i in [0,n) with d_off[i] <= offset; -1 past the end */
EXPORT int64_t zo_offset_to_frame(const zo_seek_table *st, uint64_t offset)
{
    if (offset >= st->d_off[st->n]) return -1;
    uint64_t lo = 0, hi = st->n;
    while (lo + 1 < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (st->d_off[mid] <= offset) lo = mid; else hi = mid;
    }
    return (int64_t)lo;
}

/* =====================================================================================
 * LZ4 frame — public LZ4 frame format 1.6.x + block format
 * ===================================================================================== */
static int lz4_block(const uint8_t *src, size_t n, uint8_t *dst_base, size_t *dpos, size_t cap)
{
    size_t ip = 0, op = *dpos;
    if (n == 0) return ZO_ERR_FORMAT;
    for (;;) {
        if (ip >= n) return ZO_ERR_TRUNC;
        unsigned tok = src[ip++];
        size_t ll = tok >> 4;
        if (ll == 15) {
            unsigned b;
            do { if (ip >= n) return ZO_ERR_TRUNC; b = src[ip++]; ll += b; } while (b == 255);
        }
        if (ip + ll > n) return ZO_ERR_TRUNC;
        if (op + ll > cap) return ZO_ERR_DST;
        memcpy(dst_base + op, src + ip, ll);
        ip += ll; op += ll;
        if (ip == n) break;                   /* last sequence: literals only */
        if (ip + 2 > n) return ZO_ERR_TRUNC;
        size_t off = src[ip] | ((size_t)src[ip + 1] << 8);
        ip += 2;
        size_t ml = tok & 15;
        if (ml == 15) {
            unsigned b;
            do { if (ip >= n) return ZO_ERR_TRUNC; b = src[ip++]; ml += b; } while (b == 255);
        }
        ml += 4;
        if (off == 0 || off > op) return ZO_ERR_OFFSET;   /* never before the frame start */
        if (op + ml > cap) return ZO_ERR_DST;
        for (size_t i = 0; i < ml; i++)       /* byte-serial: offset < ml replicates the pattern */
            dst_base[op + i] = dst_base[op - off + i];
        op += ml;
    }
    *dpos = op;
    return ZO_OK;
}

/* Decodes one complete LZ4 frame; returns bytes produced or a negative ZO_ERR_*. */
EXPORT ssize_t zo_lz4_frame_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
    if (n < 7) return ZO_ERR_TRUNC;
    if (rd_le32(src) != 0x184D2204u) return ZO_ERR_MAGIC;
    unsigned flg = src[4], bd = src[5];
    if ((flg >> 6) != 1) return ZO_ERR_FORMAT;            /* version */
    if (flg & 0x02) return ZO_ERR_FORMAT;                 /* reserved */
    if (bd & 0x8F) return ZO_ERR_FORMAT;                  /* reserved */
    unsigned bsid = (bd >> 4) & 7;
    if (bsid < 4) return ZO_ERR_FORMAT;
    size_t max_block = (size_t)1 << (8 + 2 * bsid);       /* 4:64K 5:256K 6:1M 7:4M */
    int block_cksum = (flg >> 4) & 1, has_csize = (flg >> 3) & 1, content_cksum = (flg >> 2) & 1, dict = flg & 1;
    size_t ip = 6;
    uint64_t content_size = 0;
    if (has_csize) { if (ip + 8 > n) return ZO_ERR_TRUNC; content_size = rd_le64(src + ip); ip += 8; }
    if (dict) { if (ip + 4 > n) return ZO_ERR_TRUNC; ip += 4; }
    if (ip + 1 > n) return ZO_ERR_TRUNC;
    if (((xxh32(src + 4, ip - 4) >> 8) & 0xff) != src[ip]) return ZO_ERR_CHECKSUM; /* HC: second byte of XXH32(descriptor) */
    ip += 1;
    size_t op = 0;
    for (;;) {
        if (ip + 4 > n) return ZO_ERR_TRUNC;
        uint32_t bs = rd_le32(src + ip);
        ip += 4;
        if (bs == 0) break;                               /* EndMark */
        int raw = (bs >> 31) & 1;
        bs &= 0x7FFFFFFFu;
        if (bs > max_block) return ZO_ERR_FORMAT;
        if (ip + bs > n) return ZO_ERR_TRUNC;
        if (raw) {
            if (op + bs > cap) return ZO_ERR_DST;
            memcpy(dst + op, src + ip, bs);
            op += bs;
        } else {
            int r = lz4_block(src + ip, bs, dst, &op, cap);
            if (r) return r;
        }
        ip += bs;
        if (block_cksum) {                                   /* XXH32 of the block's compressed bytes */
            if (ip + 4 > n) return ZO_ERR_TRUNC;
            if (xxh32(src + ip - bs, bs) != rd_le32(src + ip)) return ZO_ERR_CHECKSUM;
            ip += 4;
        }
    }
    if (has_csize && content_size != op) return ZO_ERR_FORMAT;
    if (content_cksum) {                                     /* XXH32 of the decoded frame */
        if (ip + 4 > n) return ZO_ERR_TRUNC;
        if (xxh32(dst, op) != rd_le32(src + ip)) return ZO_ERR_CHECKSUM;
        ip += 4;
    }
    return (ssize_t)op;
}

/* =====================================================================================
 * zstd — RFC 8878
 * ===================================================================================== */

/* ---- backward bitstream: bit k of the stream is bit (k & 7) of byte (k >> 3); reading proceeds
 *      from high k to low k; each read(n) returns bits [pos-n, pos) with the higher index more
 *      significant; bits below index 0 read as zero. */
typedef struct { const uint8_t *p; int64_t pos; } bbits;

static int bb_init(bbits *b, const uint8_t *p, size_t n)
{
    if (n == 0 || p[n - 1] == 0) return ZO_ERR_BITSTREAM;
    int hb = 31 - __builtin_clz((unsigned)p[n - 1]);
    b->p = p;
    b->pos = (int64_t)(n - 1) * 8 + hb;
    return ZO_OK;
}

static uint32_t bb_read(bbits *b, unsigned n) /* n <= 31 */
{
    if (n == 0) return 0;
    int64_t start = b->pos - (int64_t)n; /* may be negative */
    b->pos = start;
    uint64_t v = 0;
    /* gather bits [start, start+n) */
    for (unsigned i = 0; i < n;) {
        int64_t k = start + i;
        if (k < 0) { unsigned skip = (unsigned)(-k) < n - i ? (unsigned)(-k) : n - i; i += skip; continue; }
        unsigned bit_in_byte = (unsigned)(k & 7);
        unsigned take = 8 - bit_in_byte;
        if (take > n - i) take = n - i;
        uint64_t bits = ((uint64_t)b->p[k >> 3] >> bit_in_byte) & ((1u << take) - 1);
        v |= bits << i;
        i += take;
    }
    return (uint32_t)v;
}

/* ---- forward bitstream (FSE table description): LSB-first */
typedef struct { const uint8_t *p; size_t n; size_t bitpos; } fbits;

static uint32_t fb_peek(const fbits *f, unsigned n)
{
    uint64_t v = 0;
    size_t byte = f->bitpos >> 3;
    for (unsigned i = 0; i < 5; i++)
        if (byte + i < f->n) v |= (uint64_t)f->p[byte + i] << (8 * i);
    return (uint32_t)((v >> (f->bitpos & 7)) & (((uint64_t)1 << n) - 1));
}

/* ---- FSE decoding table */
#define FSE_MAX_LOG 9
typedef struct {
    int log;                       /* accuracy log; 0 for an RLE table (single cell, 0 bits) */
    uint8_t sym[1 << FSE_MAX_LOG];
    uint8_t nb[1 << FSE_MAX_LOG];
    uint16_t base[1 << FSE_MAX_LOG];
} fse_table;

static int bit_length(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

/* Reads the normalised counts; returns bytes consumed (>0) or error.  probs[] gets -1 for
 * "less than one". */
static int fse_read_ncount(const uint8_t *p, size_t n, int max_log, int max_sym, int16_t *probs, int *nsym, int *log_out)
{
    fbits f = { p, n, 0 };
    if (n == 0) return ZO_ERR_TRUNC;
    int al = 5 + (int)fb_peek(&f, 4);
    f.bitpos += 4;
    if (al > max_log) return ZO_ERR_TABLE;
    int remaining = 1 << al;
    int s = 0;
    while (remaining > 0) {
        if (s > max_sym) return ZO_ERR_TABLE;
        int bits = bit_length((uint32_t)(remaining + 1));
        uint32_t val = fb_peek(&f, (unsigned)bits);
        uint32_t low = (1u << (bits - 1)) - 1;
        uint32_t thr = (1u << bits) - 1 - (uint32_t)(remaining + 1);
        if ((val & low) < thr) { val &= low; f.bitpos += (size_t)bits - 1; }
        else { if (val > low) val -= thr; f.bitpos += (size_t)bits; }
        int prob = (int)val - 1;
        probs[s++] = (int16_t)prob;
        remaining -= prob < 0 ? -prob : prob;
        if (prob == 0) {
            for (;;) {
                uint32_t rep = fb_peek(&f, 2);
                f.bitpos += 2;
                for (uint32_t i = 0; i < rep; i++) { if (s > max_sym) return ZO_ERR_TABLE; probs[s++] = 0; }
                if (rep != 3) break;
            }
        }
        if ((f.bitpos + 7) / 8 > n) return ZO_ERR_TRUNC;
    }
    if (remaining != 0) return ZO_ERR_TABLE;
    *nsym = s;
    *log_out = al;
    return (int)((f.bitpos + 7) / 8);
}

static int fse_build(fse_table *t, const int16_t *probs, int nsym, int log)
{
    int size = 1 << log;
    int high = size - 1;
    uint16_t next[256];
    t->log = log;
    for (int s = 0; s < nsym; s++) {
        if (probs[s] == -1) { t->sym[high--] = (uint8_t)s; next[s] = 1; }
        else next[s] = (uint16_t)probs[s];
    }
    int step = (size >> 1) + (size >> 3) + 3, mask = size - 1, pos = 0;
    for (int s = 0; s < nsym; s++) {
        for (int i = 0; i < probs[s]; i++) {
            t->sym[pos] = (uint8_t)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    }
    if (pos != 0) return ZO_ERR_TABLE;
    for (int i = 0; i < size; i++) {
        int s = t->sym[i];
        uint32_t d = next[s]++;
        int nb = log - (bit_length(d) - 1);
        t->nb[i] = (uint8_t)nb;
        t->base[i] = (uint16_t)((d << nb) - (uint32_t)size);
    }
    return ZO_OK;
}

static void fse_build_rle(fse_table *t, uint8_t sym)
{
    t->log = 0; t->sym[0] = sym; t->nb[0] = 0; t->base[0] = 0;
}

/* ---- predefined distributions (RFC 8878 §3.1.1.3.2.2) */
static const int16_t LL_DEFAULT[36] = { 4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1 };
static const int16_t ML_DEFAULT[53] = { 1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1 };
static const int16_t OF_DEFAULT[29] = { 1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1 };

static const uint32_t LL_BASE[36] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,128,256,512,1024,2048,4096,8192,16384,32768,65536 };
static const uint8_t LL_BITS[36] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16 };
static const uint32_t ML_BASE[53] = { 3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,35,37,39,41,43,47,51,59,67,83,99,131,259,515,1027,2051,4099,8195,16387,32771,65539 };
static const uint8_t ML_BITS[53] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16 };

/* ---- Huffman */
#define HUF_MAX_LOG 11
typedef struct { int log; uint8_t sym[1 << HUF_MAX_LOG]; uint8_t nb[1 << HUF_MAX_LOG]; } huf_table;

/* Parses the tree description at p; fills the decode table; returns bytes consumed or error. */
static int huf_read_table(huf_table *h, const uint8_t *p, size_t n)
{
    uint8_t w[256];
    int nw = 0;
    if (n < 1) return ZO_ERR_TRUNC;
    unsigned hb = p[0];
    size_t used;
    if (hb >= 128) {                      /* direct 4-bit weights, high nibble first */
        nw = (int)hb - 127;
        size_t bytes = ((size_t)nw + 1) / 2;
        if (1 + bytes > n) return ZO_ERR_TRUNC;
        for (int i = 0; i < nw; i++) {
            unsigned b = p[1 + i / 2];
            w[i] = (uint8_t)((i & 1) ? (b & 15) : (b >> 4));
        }
        used = 1 + bytes;
    } else {                              /* FSE-compressed weights, two interleaved states */
        if (hb == 0 || 1 + (size_t)hb > n) return ZO_ERR_TRUNC;
        int16_t probs[16];
        int nsym, log;
        int hdr = fse_read_ncount(p + 1, hb, 6, 12, probs, &nsym, &log);
        if (hdr < 0) return hdr;
        fse_table *t = malloc(sizeof(*t));
        int r = fse_build(t, probs, nsym, log);
        if (r) { free(t); return r; }
        bbits b;
        if ((size_t)hdr >= hb) { free(t); return ZO_ERR_TRUNC; }
        r = bb_init(&b, p + 1 + hdr, (size_t)hb - (size_t)hdr);
        if (r) { free(t); return r; }
        uint32_t s1 = bb_read(&b, (unsigned)log), s2 = bb_read(&b, (unsigned)log);
        for (;;) {
            if (nw > 253) { free(t); return ZO_ERR_TABLE; }
            w[nw++] = t->sym[s1];
            s1 = t->base[s1] + bb_read(&b, t->nb[s1]);
            if (b.pos < 0) { w[nw++] = t->sym[s2]; break; }
            w[nw++] = t->sym[s2];
            s2 = t->base[s2] + bb_read(&b, t->nb[s2]);
            if (b.pos < 0) { w[nw++] = t->sym[s1]; break; }
        }
        free(t);
        used = 1 + (size_t)hb;
    }
    /* implicit last weight */
    uint32_t total = 0;
    for (int i = 0; i < nw; i++) {
        if (w[i] > HUF_MAX_LOG + 1) return ZO_ERR_TABLE;
        if (w[i]) total += 1u << (w[i] - 1);
    }
    if (total == 0) return ZO_ERR_TABLE;
    int max_bits = bit_length(total);
    if (max_bits > HUF_MAX_LOG) return ZO_ERR_TABLE;
    uint32_t rest = (1u << max_bits) - total;
    if (rest == 0 || (rest & (rest - 1))) return ZO_ERR_TABLE;
    w[nw++] = (uint8_t)(bit_length(rest));  /* log2(rest) + 1 */
    if (nw > 256) return ZO_ERR_TABLE;
    /* table fill: decreasing code length, symbols in increasing order inside a length */
    uint32_t count[HUF_MAX_LOG + 2] = {0}, idx[HUF_MAX_LOG + 2];
    for (int i = 0; i < nw; i++) if (w[i]) count[max_bits + 1 - w[i]]++;      /* count by nbBits */
    idx[max_bits] = 0;
    for (int L = max_bits; L >= 1; L--) idx[L - 1] = idx[L] + count[L] * (1u << (max_bits - L));
    for (int s = 0; s < nw; s++) {
        if (!w[s]) continue;
        int nbits = max_bits + 1 - w[s];
        uint32_t run = 1u << (max_bits - nbits);
        for (uint32_t i = 0; i < run; i++) { h->sym[idx[nbits] + i] = (uint8_t)s; h->nb[idx[nbits] + i] = (uint8_t)nbits; }
        idx[nbits] += run;
    }
    h->log = max_bits;
    return (int)used;
}

static int huf_decode_stream(const huf_table *h, const uint8_t *p, size_t n, uint8_t *out, size_t nout)
{
    bbits b;
    int r = bb_init(&b, p, n);
    if (r) return r;
    uint32_t mask = (1u << h->log) - 1;
    uint32_t state = bb_read(&b, (unsigned)h->log);
    for (size_t i = 0; i < nout; i++) {
        out[i] = h->sym[state];
        unsigned nb = h->nb[state];
        state = ((state << nb) | bb_read(&b, nb)) & mask;
    }
    if (b.pos != -(int64_t)h->log) return ZO_ERR_BITSTREAM;
    return ZO_OK;
}

/* ---- per-frame decoder state carried across blocks */
typedef struct {
    huf_table huf; int huf_valid;
    fse_table ll, of, ml; int ll_valid, of_valid, ml_valid;
    uint32_t rep[3];
    uint8_t *lit;     /* literal buffer (<= 128 KiB) */
} zstd_state;

static int seq_table(fse_table *t, int *valid, int mode, const uint8_t *p, size_t n, size_t *ip,
                     const int16_t *def, int def_n, int def_log, int max_log, int max_sym)
{
    switch (mode) {
    case 0: { int r = fse_build(t, def, def_n, def_log); if (r) return r; *valid = 1; return ZO_OK; }
    case 1: if (*ip + 1 > n) return ZO_ERR_TRUNC; if (p[*ip] > max_sym) return ZO_ERR_TABLE; fse_build_rle(t, p[*ip]); *ip += 1; *valid = 1; return ZO_OK;
    case 2: {
        int16_t probs[64]; int nsym, log;
        int used = fse_read_ncount(p + *ip, n - *ip, max_log, max_sym, probs, &nsym, &log);
        if (used < 0) return used;
        int r = fse_build(t, probs, nsym, log);
        if (r) return r;
        *ip += (size_t)used; *valid = 1; return ZO_OK;
    }
    default: return *valid ? ZO_OK : ZO_ERR_TABLE;   /* Repeat */
    }
}

static int zstd_block(zstd_state *st, const uint8_t *p, size_t n, uint8_t *dst, size_t *dpos, size_t cap)
{
    size_t ip = 0, op = *dpos;
    /* ---------- literals section */
    if (n < 1) return ZO_ERR_TRUNC;
    unsigned b0 = p[0], ltype = b0 & 3, sf = (b0 >> 2) & 3;
    size_t regen, comp = 0, hsize;
    int streams = 1;
    const uint8_t *lit;
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { regen = b0 >> 3; hsize = 1; }
        else if (sf == 1) { if (n < 2) return ZO_ERR_TRUNC; regen = (b0 >> 4) + ((size_t)p[1] << 4); hsize = 2; }
        else { if (n < 3) return ZO_ERR_TRUNC; regen = (b0 >> 4) + ((size_t)p[1] << 4) + ((size_t)p[2] << 12); hsize = 3; }
        if (regen > (128u << 10)) return ZO_ERR_FORMAT;
        ip = hsize;
        if (ltype == 0) { if (ip + regen > n) return ZO_ERR_TRUNC; lit = p + ip; ip += regen; }
        else { if (ip + 1 > n) return ZO_ERR_TRUNC; memset(st->lit, p[ip], regen); lit = st->lit; ip += 1; }
    } else {
        int bits;
        if (sf == 0) { hsize = 3; bits = 10; streams = 1; }
        else if (sf == 1) { hsize = 3; bits = 10; streams = 4; }
        else if (sf == 2) { hsize = 4; bits = 14; streams = 4; }
        else { hsize = 5; bits = 18; streams = 4; }
        if (n < hsize) return ZO_ERR_TRUNC;
        uint64_t v = 0;
        for (size_t i = 0; i < hsize; i++) v |= (uint64_t)p[i] << (8 * i);
        regen = (size_t)((v >> 4) & ((1u << bits) - 1));
        comp = (size_t)(v >> (4 + bits));
        if (regen > (128u << 10)) return ZO_ERR_FORMAT;
        ip = hsize;
        if (ip + comp > n) return ZO_ERR_TRUNC;
        const uint8_t *q = p + ip;
        size_t qn = comp;
        if (ltype == 2) {
            int used = huf_read_table(&st->huf, q, qn);
            if (used < 0) return used;
            st->huf_valid = 1;
            q += used; qn -= (size_t)used;
        } else if (!st->huf_valid) return ZO_ERR_TABLE;
        if (streams == 1) {
            int r = huf_decode_stream(&st->huf, q, qn, st->lit, regen);
            if (r) return r;
        } else {
            if (qn < 6) return ZO_ERR_TRUNC;
            size_t s1 = q[0] | ((size_t)q[1] << 8), s2 = q[2] | ((size_t)q[3] << 8), s3 = q[4] | ((size_t)q[5] << 8);
            if (6 + s1 + s2 + s3 > qn) return ZO_ERR_TRUNC;
            size_t s4 = qn - 6 - s1 - s2 - s3;
            size_t per = (regen + 3) / 4;
            if (3 * per > regen) return ZO_ERR_FORMAT;
            const uint8_t *sp = q + 6;
            int r;
            if ((r = huf_decode_stream(&st->huf, sp, s1, st->lit, per))) return r;
            if ((r = huf_decode_stream(&st->huf, sp + s1, s2, st->lit + per, per))) return r;
            if ((r = huf_decode_stream(&st->huf, sp + s1 + s2, s3, st->lit + 2 * per, per))) return r;
            if ((r = huf_decode_stream(&st->huf, sp + s1 + s2 + s3, s4, st->lit + 3 * per, regen - 3 * per))) return r;
        }
        lit = st->lit;
        ip += comp;
    }
    /* ---------- sequences section */
    if (ip + 1 > n) return ZO_ERR_TRUNC;
    size_t nseq = p[ip++];
    if (nseq >= 128) {
        if (nseq == 255) { if (ip + 2 > n) return ZO_ERR_TRUNC; nseq = p[ip] + ((size_t)p[ip + 1] << 8) + 0x7F00; ip += 2; }
        else { if (ip + 1 > n) return ZO_ERR_TRUNC; nseq = ((nseq - 128) << 8) + p[ip]; ip += 1; }
    }
    size_t lpos = 0;
    if (nseq > 0) {
        if (ip + 1 > n) return ZO_ERR_TRUNC;
        unsigned modes = p[ip++];
        if (modes & 3) return ZO_ERR_FORMAT;
        int r;
        if ((r = seq_table(&st->ll, &st->ll_valid, (modes >> 6) & 3, p, n, &ip, LL_DEFAULT, 36, 6, 9, 35))) return r;
        if ((r = seq_table(&st->of, &st->of_valid, (modes >> 4) & 3, p, n, &ip, OF_DEFAULT, 29, 5, 8, 31))) return r;
        if ((r = seq_table(&st->ml, &st->ml_valid, (modes >> 2) & 3, p, n, &ip, ML_DEFAULT, 53, 6, 9, 52))) return r;
        bbits b;
        if (ip >= n) return ZO_ERR_TRUNC;
        if ((r = bb_init(&b, p + ip, n - ip))) return r;
        uint32_t sl = bb_read(&b, (unsigned)st->ll.log), so = bb_read(&b, (unsigned)st->of.log), sm = bb_read(&b, (unsigned)st->ml.log);
        for (size_t i = 0; i < nseq; i++) {
            unsigned oc = st->of.sym[so], mc = st->ml.sym[sm], lc = st->ll.sym[sl];
            if (oc > 31 || mc > 52 || lc > 35) return ZO_ERR_TABLE;
            uint32_t ov = (1u << oc) + bb_read(&b, oc);
            uint32_t mlen = ML_BASE[mc] + bb_read(&b, ML_BITS[mc]);
            uint32_t llen = LL_BASE[lc] + bb_read(&b, LL_BITS[lc]);
            if (i + 1 < nseq) {
                sl = st->ll.base[sl] + bb_read(&b, st->ll.nb[sl]);
                sm = st->ml.base[sm] + bb_read(&b, st->ml.nb[sm]);
                so = st->of.base[so] + bb_read(&b, st->of.nb[so]);
            }
            if (b.pos < 0) return ZO_ERR_BITSTREAM;
            uint32_t offset;
            if (ov > 3) {
                offset = ov - 3;
                st->rep[2] = st->rep[1]; st->rep[1] = st->rep[0]; st->rep[0] = offset;
            } else {
                unsigned idx = ov - 1 + (llen == 0);
                if (idx == 0) offset = st->rep[0];
                else {
                    offset = idx < 3 ? st->rep[idx] : st->rep[0] - 1;
                    if (offset == 0) return ZO_ERR_OFFSET;
                    if (idx > 1) st->rep[2] = st->rep[1];
                    st->rep[1] = st->rep[0];
                    st->rep[0] = offset;
                }
            }
            if (lpos + llen > regen) return ZO_ERR_FORMAT;
            if (op + llen + mlen > cap) return ZO_ERR_DST;
            memcpy(dst + op, lit + lpos, llen);
            op += llen; lpos += llen;
            if (offset > op) return ZO_ERR_OFFSET;
            for (uint32_t k = 0; k < mlen; k++) dst[op + k] = dst[op - offset + k];
            op += mlen;
        }
        if (b.pos != 0) return ZO_ERR_BITSTREAM;
    }
    size_t rest = regen - lpos;
    if (op + rest > cap) return ZO_ERR_DST;
    memcpy(dst + op, lit + lpos, rest);
    op += rest;
    if (op - *dpos > (128u << 10)) return ZO_ERR_FORMAT;
    *dpos = op;
    return ZO_OK;
}

/* Decodes one complete zstd frame; returns bytes produced or a negative ZO_ERR_*. */
EXPORT ssize_t zo_zstd_frame_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
    if (n < 6) return ZO_ERR_TRUNC;
    if (rd_le32(src) != 0xFD2FB528u) return ZO_ERR_MAGIC;
    unsigned fhd = src[4];
    unsigned fcs_flag = fhd >> 6, ss = (fhd >> 5) & 1, cksum = (fhd >> 2) & 1, did = fhd & 3;
    if (fhd & 0x08) return ZO_ERR_FORMAT;
    size_t ip = 5;
    if (!ss) { if (ip + 1 > n) return ZO_ERR_TRUNC; ip += 1; }               /* window descriptor */
    static const unsigned did_sz[4] = { 0, 1, 2, 4 };
    if (did) {
        if (ip + did_sz[did] > n) return ZO_ERR_TRUNC;
        uint32_t id = 0;
        for (unsigned i = 0; i < did_sz[did]; i++) id |= (uint32_t)src[ip + i] << (8 * i);
        if (id) return ZO_ERR_UNSUPPORTED;
        ip += did_sz[did];
    }
    unsigned fcs_sz = fcs_flag == 0 ? (ss ? 1 : 0) : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
    uint64_t fcs = 0; int has_fcs = fcs_sz > 0;
    if (ip + fcs_sz > n) return ZO_ERR_TRUNC;
    for (unsigned i = 0; i < fcs_sz; i++) fcs |= (uint64_t)src[ip + i] << (8 * i);
    if (fcs_sz == 2) fcs += 256;
    ip += fcs_sz;

    zstd_state *st = calloc(1, sizeof(*st));
    st->lit = malloc(128u << 10);
    st->rep[0] = 1; st->rep[1] = 4; st->rep[2] = 8;
    size_t op = 0;
    int r = ZO_OK;
    for (;;) {
        if (ip + 3 > n) { r = ZO_ERR_TRUNC; break; }
        uint32_t bh = src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
        ip += 3;
        int last = bh & 1, type = (bh >> 1) & 3;
        size_t bsize = bh >> 3;
        if (type == 0) {
            if (ip + bsize > n) { r = ZO_ERR_TRUNC; break; }
            if (op + bsize > cap) { r = ZO_ERR_DST; break; }
            memcpy(dst + op, src + ip, bsize); op += bsize; ip += bsize;
        } else if (type == 1) {
            if (ip + 1 > n) { r = ZO_ERR_TRUNC; break; }
            if (op + bsize > cap) { r = ZO_ERR_DST; break; }
            memset(dst + op, src[ip], bsize); op += bsize; ip += 1;
        } else if (type == 2) {
            if (ip + bsize > n) { r = ZO_ERR_TRUNC; break; }
            if (bsize > (128u << 10)) { r = ZO_ERR_FORMAT; break; }
            r = zstd_block(st, src + ip, bsize, dst, &op, cap);
            if (r) break;
            ip += bsize;
        } else { r = ZO_ERR_FORMAT; break; }
        if (last) break;
    }
    free(st->lit);
    free(st);
    if (r) return r;
    if (has_fcs && fcs != op) return ZO_ERR_FORMAT;
    if (cksum) {                                             /* low 32 bits of XXH64 of the decoded frame */
        if (ip + 4 > n) return ZO_ERR_TRUNC;
        if ((uint32_t)xxh64(dst, op) != rd_le32(src + ip)) return ZO_ERR_CHECKSUM;
        ip += 4;
    }
    return (ssize_t)op;
}

/* =====================================================================================
 * zseek_pread result semantics — reference src/decompress.c:470-574 (zstd) / :685-804 (lz4)
 *   B1 never crosses a frame boundary: returns MIN(count, frame_end - offset)
 *   B2 offset >= total -> 0 ; B3 count == 0 -> 0 ; B4 lookup = last frame starting <= offset
 *   B6 the seek table's dSize is the frame length that is trusted
 * ===================================================================================== */
typedef struct {
    const uint8_t *img; size_t size;
    zo_seek_table *st;
    int type; /* 0 zstd, 1 lz4 */
    int64_t cached_idx; uint8_t *cached; size_t cached_len;
} zo_reader;

EXPORT zo_reader *zo_reader_open(const uint8_t *img, size_t size)
{
    if (size < 4) return NULL;                      /* "unexpected EOF" */
    uint32_t magic = rd_le32(img);
    int type;
    if (magic == 0xFD2FB528u) type = 0;
    else if (magic == 0x184D2204u) type = 1;
    else return NULL;                               /* "unrecognized file format" */
    zo_seek_table *st = zo_seek_table_parse(img, size);
    if (!st) return NULL;                           /* "read_seek_table failed" */
    zo_reader *r = calloc(1, sizeof(*r));
    r->img = img; r->size = size; r->st = st; r->type = type; r->cached_idx = -1;
    return r;
}

EXPORT void zo_reader_close(zo_reader *r)
{
    if (!r) return;
    zo_seek_table_free(r->st);
    free(r->cached);
    free(r);
}

EXPORT int zo_reader_type(const zo_reader *r) { return r->type; }
EXPORT const zo_seek_table *zo_reader_seek_table(const zo_reader *r) { return r->st; }

/* Decodes frame idx into dst (capacity must be >= its dSize); returns bytes produced or <0. */
EXPORT ssize_t zo_reader_decode_frame(zo_reader *r, uint64_t idx, uint8_t *dst, size_t cap)
{
    if (idx >= r->st->n) return ZO_ERR_FORMAT;
    uint64_t c0 = r->st->c_off[idx], c1 = r->st->c_off[idx + 1];
    if (c1 > r->size) return ZO_ERR_TRUNC;         /* "unexpected EOF" */
    return r->type == 0 ? zo_zstd_frame_decode(r->img + c0, (size_t)(c1 - c0), dst, cap)
                        : zo_lz4_frame_decode(r->img + c0, (size_t)(c1 - c0), dst, cap);
}

EXPORT ssize_t zo_pread(zo_reader *r, void *buf, size_t count, size_t offset)
{
    int64_t idx = zo_offset_to_frame(r->st, offset);
    if (idx < 0) return 0;
    size_t dsize = (size_t)(r->st->d_off[idx + 1] - r->st->d_off[idx]);
    if (r->cached_idx != idx) {
        uint8_t *d = malloc(dsize ? dsize : 1);
        ssize_t got = zo_reader_decode_frame(r, (uint64_t)idx, d, dsize);
        if (got < 0 || (size_t)got != dsize) { free(d); return -1; }
        free(r->cached);
        r->cached = d; r->cached_idx = idx; r->cached_len = dsize;
    }
    size_t in_frame = offset - (size_t)r->st->d_off[idx];
    size_t n = count < dsize - in_frame ? count : dsize - in_frame;
    memcpy(buf, r->cached + in_frame, n);
    return (ssize_t)n;
}

/* Whole-file decode into dst (capacity >= total decompressed size). Returns bytes or <0. */
EXPORT ssize_t zo_decode_all(zo_reader *r, uint8_t *dst, size_t cap)
{
    size_t total = (size_t)r->st->d_off[r->st->n];
    if (cap < total) return ZO_ERR_DST;
    for (uint64_t i = 0; i < r->st->n; i++) {
        size_t dsize = (size_t)(r->st->d_off[i + 1] - r->st->d_off[i]);
        ssize_t got = zo_reader_decode_frame(r, i, dst + r->st->d_off[i], dsize);
        if (got < 0) return got;
        if ((size_t)got != dsize) return ZO_ERR_FORMAT;
    }
    return (ssize_t)total;
}
