/*
 * reader.c — host side of the B200 read path: the reference's reader API (include/zseek.h) plus the
 * additive entry points of include/zseek_b200.h, implemented in C over the C-ABI launch layer
 * (zsk_cuda.h).  All decode arithmetic runs in the sm_100a kernels; this file contains NO decoder.
 *
 * Reference components re-designed here (reference paths relative to /root/reference):
 *   src/decompress.c:261-295  zseek_reader_open_full/_open  -> magic sniff, seek-table parse, device context
 *   src/seek_table.c:62-176   read_seek_table               -> st_load(): same checks, N+1 u64 prefix arrays,
 *                                                             uploaded once to HBM (g_coff/g_doff)
 *   src/seek_table.c:187-202  offset_to_frame_idx           -> st_lookup() for single calls, K1 for batches
 *   src/decompress.c:806-824  zseek_pread                   -> lookup, HBM cache / pinned mirror hit, else
 *                                                             decode a read-ahead window of frames in ONE launch
 *   src/cache.c               LRU of malloc'd frames        -> HBM slab of nslots decoded frames with a real
 *                                                             O(1) LRU (the reference list is buggy)
 *   src/buffer.c              growable staging buffers      -> pinned ingest staging (h_stage) + pinned decoded
 *                                                             windows (h_mir[]) + device compressed image
 *   src/common.c              set_error                     -> set_error (same message strings)
 */
#define _GNU_SOURCE
#include <errno.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdatomic.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/zseek_b200.h"
#include "zsk_cuda.h"

#define ZSTD_MAGIC 0xFD2FB528u
#define LZ4_MAGIC 0x184D2204u
#define SEEKABLE_MAGIC 0x8F92EAB1u
#define SKIPPABLE_MAGIC 0x184D2A5Eu
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))

struct zseek_reader {
    zseek_read_file_t user_file;
    int codec;
    pthread_mutex_t lock;
    size_t pos;

    /* default FILE* I/O (zseek_reader_open): the descriptor behind the FILE*, read with pread(2) from a small worker pool
     * when a large range is pulled into the pinned staging; -1 = caller-supplied callbacks */
    int file_fd;
    struct io_pool *io;

    /* memory-image mode (zseek_b200_reader_open_mem) */
    const uint8_t *mem_image;
    size_t mem_size;

    /* seek table, host */
    uint64_t nframes;
    uint64_t *c_off, *d_off; /* [N+1] */
    uint32_t max_csize, max_dsize;

    /* device */
    zsk_cuda_ctx *cx;
    uint64_t *g_coff, *g_doff;
    size_t parked_bytes; /* device memory this reader held when it was parked */
    size_t n_cap;      /* entries g_coff / g_doff / g_frame_src hold */
    size_t slab_bytes; /* bytes of g_slab */

    /* shard */
    uint64_t shard_lo, shard_hi;

    /* compressed image residency: frames [res_lo, res_hi) */
    uint8_t *g_comp; /* allocation; image byte c_off[res_lo] sits at g_comp + ZSK_PAD_FRONT */
    size_t g_comp_cap;
    uint64_t res_lo, res_hi;

    /* pinned host staging */
    uint8_t *h_stage; /* two halves of stage_half bytes (grown on demand up to stage_half_max) */
    size_t stage_half, stage_half_max;
    int stage_next, stage_inflight;
    uint8_t *h_mir[2]; /* two pinned windows of mir_cap[] bytes; window mir_cur holds the decoded bytes of frames [mir_lo, mir_hi) */
    size_t mir_cap[2];
    uint64_t mir_lo, mir_hi;
    int mir_cur;
    /* asynchronous read-ahead of a sequential host scan: frames [pf_lo, pf_hi) are being decoded into the OTHER half
     * (everything is queued on the streams; the host has not waited for it yet) */
    bool pf_active, pf_resident, pf_ok;
    uint64_t pf_lo, pf_hi;

    /* HBM decoded-frame cache */
    size_t user_cache_size;
    uint32_t nslots;
    size_t slot_size;
    uint8_t *g_slab; /* allocation; slot s at g_slab + ZSK_PAD_FRONT + s * slot_size */
    int32_t *slot_frame, *lru_prev, *lru_next;
    int32_t lru_head, lru_tail; /* head = MRU */
    int32_t *frame_slot;        /* [N] */
    uint32_t *valid_len;        /* [N] decoded bytes of a cached frame that are valid (a batch may decode only a prefix) */
    uint32_t cached;
    int64_t *g_frame_src;       /* device [N]: byte offset of frame in slab data area, -1 absent */
    int64_t *h_frame_src;       /* host copy being edited */
    bool frame_src_dirty;

    /* job buffers (device + pinned host), capacity job_cap */
    uint32_t job_cap;
    uint32_t *g_job_ids, *h_job_ids;
    uint64_t *g_job_offs, *h_job_offs;
    uint32_t *g_job_limits, *h_job_limits;
    int32_t *g_job_status, *h_job_status;

    /* batch scratch (device), capacity batch_cap requests */
    size_t batch_cap;
    uint64_t *g_b_offsets, *g_b_counts, *g_b_dstoffs;
    int32_t *g_b_frame;
    uint32_t *g_b_inframe, *g_b_nbytes;
    uint32_t *g_touched;
    uint32_t *h_touched;
    uint8_t *g_out; /* device staging for host destinations */
    size_t g_out_cap;

    /* stream-ordered batches (zseek_b200_pread_batch_async): slab of bs_cap decoded-frame slots, job arrays filled on the
     * device by zsk_compact_kernel, per-frame slot map, {job count, error flag} */
    uint8_t *g_bslab;
    uint32_t bs_cap;
    uint32_t *g_bjob_ids, *g_bjob_limits, *g_bctl;
    uint64_t *g_bjob_offs;
    int32_t *g_bjob_status;
    int64_t *g_bsrc;
    int async_stream;        /* ZSK_STREAM_COMPUTE or ZSK_STREAM_USER of the batch in flight; -1 = none */
    uint32_t async_jobs_max;

    /* residency for random host readers: after resident_after isolated misses the whole shard is decoded once into the
     * pinned window (if it is at most resident_max bytes); every later host read is a memcpy */
    uint32_t random_misses, resident_after;
    size_t resident_max;
    bool resident_tried, resident;
    size_t resident_bytes; /* what this reader added to g_resident_bytes */
    size_t window_cap; /* most bytes an ordinary read-ahead window holds */
    bool counted;      /* this reader is in g_live_readers */

    /* read-ahead */
    uint64_t ra_next;
    uint32_t ra_window, ra_max;
    uint32_t ra_cap; /* largest read-ahead window (frames) this reader uses; 0 = not decided yet (ra_limit) */
    size_t chunk_bytes; /* decoded bytes per pipeline stage of host-destination range reads */
    size_t ramp_bytes;  /* size of the first pipeline stage; stages double until they reach chunk_bytes */
    /* launches of at least sort_min LZ4 frames (sort_min_zstd zstd frames) hand the kernel a job list ordered by
     * compressed size, largest first: the lane-per-frame LZ4 kernel then runs frames of similar length side by
     * side in a warp, warps retire as a whole, and for both codecs the longest frames start first.  0 = never. */
    size_t sort_min, sort_min_zstd;
    bool partial_decode; /* batches decode a missing frame only up to the last byte they need of it (ZSEEK_B200_PARTIAL=0: whole frames) */
    uint64_t sorted_lo, sorted_hi; /* g_job_ids currently holds the ordered list of [sorted_lo, sorted_hi) */

    /* host/device classification of caller buffers, cached per 2 MiB virtual-address block so that the hot
     * zseek_pread path does not enter the CUDA driver (a global lock) on every call */
    _Atomic uint64_t ptr_cache[8]; /* ((address >> 21) + 1) << 1 | is_device; 0 = empty.  Atomic: the resident fast path reads it without the lock */

    /* Shared readers: once the whole shard sits decoded in the pinned window (resident), host reads are memcpys out of
     * memory that no longer changes; they take res_lock for reading instead of the reader mutex, so the threads of a pool
     * that shares one reader copy in parallel (the reference serialises them on its lock, src/decompress.c:387,499). */
    pthread_rwlock_t res_lock;
    atomic_bool resident_fast;

    /* residency in HBM: when the pinned budget of the process is spent (one reader per thread over the same file), the shard
     * is decoded once into device memory instead and every read becomes a small device-to-host copy */
    uint8_t *g_res;
    uint8_t *h_bounce; /* pinned, ZSK_BOUNCE bytes: small reads land here first (a copy to pageable memory goes through the driver's own staging and takes ~10x longer when many threads do it at once) */
    size_t hbm_resident_bytes;
    bool hbm_resident;
};

/* Process-wide bookkeeping for callers that keep one reader per thread: pinned host memory is the scarce resource (it is
 * slow to create and every reader wants two decoded windows of it), so the size of a read-ahead window follows the number
 * of live readers, and whole-shard residency draws from one process-wide budget. */
static atomic_uint g_live_readers;
static atomic_size_t g_resident_bytes;
static atomic_size_t g_hbm_resident_bytes;
#define ZSK_BOUNCE ((size_t)64 << 10)

/* Closed readers are PARKED, not destroyed: the next zseek_reader_open* of the process takes over the device context (streams,
 * events, the zstd scratch pools), the device buffers and the pinned windows of a parked one.  A caller that opens one
 * reader per thread (or per request) otherwise pays ~20 allocations, each under the driver's process-wide lock, on the
 * first read of every reader: measured 12-19 ms for a one-frame miss with 16 threads, against 1.3 ms alone. */
#define ZSK_MAX_PARKED 64
static pthread_mutex_t g_park_mu = PTHREAD_MUTEX_INITIALIZER;
static struct zseek_reader *g_parked[ZSK_MAX_PARKED];
static int g_nparked;
static size_t g_parked_bytes; /* device memory the parked readers hold; at most ZSEEK_B200_PARK_MB (default 8192) */

/* ------------------------------------------------------------------ errors (reference src/common.c:45-54) */
static bool stream_frames_finish(zseek_reader_t *r, uint64_t lo, uint64_t hi, bool resident, bool ok, char *errbuf);
static void prefetch_drop(zseek_reader_t *r);
static void prefetch_start(zseek_reader_t *r, void *call_data);
static uint8_t *mirror_half(zseek_reader_t *r, int which);
static size_t env_size(const char *name, size_t dflt);
static bool stream_frames_to_host(zseek_reader_t *r, uint64_t lo, uint64_t hi, uint8_t *dst, void *call_data, char *errbuf);

/* A stream-ordered batch queued on a CALLER stream is invisible to the reader's own streams.  Before any other entry point
 * touches what that batch uses (compressed image, batch scratch, the context's zstd scratch pools), the reader's streams
 * are made to wait for it on the device; the host does not block and the batch's verdict stays pending for
 * zseek_b200_batch_wait. */
static void async_fence(zseek_reader_t *r)
{
    if (r->async_stream != ZSK_STREAM_USER)
        return;
    zsk_cuda_stream_wait(r->cx, ZSK_STREAM_COMPUTE, ZSK_STREAM_USER);
    zsk_cuda_stream_wait(r->cx, ZSK_STREAM_H2D, ZSK_STREAM_USER);
}

static void set_error(char errbuf[ZSEEK_ERRBUF_SIZE], const char *fmt, ...)
{
    if (!errbuf)
        return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errbuf, ZSEEK_ERRBUF_SIZE, fmt, ap);
    va_end(ap);
}

static const char *status_name(int st)
{
    switch (st) {
    case ZSK_ST_TRUNC: return "compressed frame truncated";
    case ZSK_ST_MAGIC: return "bad frame magic";
    case ZSK_ST_FORMAT: return "corrupted frame";
    case ZSK_ST_DST: return "frame larger than seek table entry";
    case ZSK_ST_OFFSET: return "match offset out of range";
    case ZSK_ST_BITSTREAM: return "corrupted bitstream";
    case ZSK_ST_TABLE: return "corrupted entropy table";
    case ZSK_ST_UNSUPPORTED: return "dictionary not supported";
    case ZSK_ST_SIZE: return "frame smaller than seek table entry";
    case ZSK_ST_CHECKSUM: return "checksum mismatch";
    default: return "unknown error";
    }
}

static bool cuda_fail(zseek_reader_t *r, char *errbuf, const char *what)
{
    set_error(errbuf, "%s: %s", what, zsk_cuda_error(r->cx));
    return false;
}

static uint32_t rd_le32(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* ------------------------------------------------------------------ default FILE* I/O (reference src/decompress.c:47-98) */
static ssize_t file_pread(void *data, size_t size, size_t offset, void *user_data, void *call_data)
{
    (void)call_data;
    FILE *f = user_data;
    long saved = ftell(f);
    if (saved == -1)
        return -1;
    if (fseek(f, (long)offset, SEEK_SET) == -1)
        return -1;
    size_t got = fread(data, 1, size, f);
    if (got != size && ferror(f))
        return -1;
    if (fseek(f, saved, SEEK_SET) == -1) /* the caller's file position is left untouched */
        return -1;
    return (ssize_t)got;
}

static ssize_t file_fsize(void *user_data, void *call_data)
{
    (void)call_data;
    int fd = fileno((FILE *)user_data);
    struct stat st;
    if (fd == -1 || fstat(fd, &st) == -1)
        return -1;
    return st.st_size;
}

/* memory image I/O (zseek_b200_reader_open_mem) */
static ssize_t mem_pread(void *data, size_t size, size_t offset, void *user_data, void *call_data)
{
    (void)call_data;
    zseek_reader_t *r = user_data;
    if (offset >= r->mem_size)
        return 0;
    size_t n = MIN(size, r->mem_size - offset);
    memcpy(data, r->mem_image + offset, n);
    return (ssize_t)n;
}

static ssize_t mem_fsize(void *user_data, void *call_data)
{
    (void)call_data;
    return (ssize_t)((zseek_reader_t *)user_data)->mem_size;
}

/* ------------------------------------------------------------------ seek table (reference src/seek_table.c:62-176) */
static bool st_load(zseek_reader_t *r, void *call_data)
{
    zseek_read_file_t uf = r->user_file;
    ssize_t fsize = uf.fsize(uf.user_data, call_data);
    if (fsize < 0)
        return false;
    uint8_t footer[9];
    if (uf.pread(footer, 9, (size_t)fsize - 9, uf.user_data, call_data) != 9)
        return false;
    if (rd_le32(footer + 5) != SEEKABLE_MAGIC)
        return false;
    uint8_t desc = footer[4];
    if (desc & 0x7c) /* reserved descriptor bits */
        return false;
    size_t esz = (desc & 0x80) ? 12 : 8; /* per-entry checksums are parsed over but never verified (B7) */
    uint64_t n = rd_le32(footer);
    uint64_t table_bytes = 8 + n * esz + 9;
    uint8_t header[8];
    if (uf.pread(header, 8, (size_t)fsize - table_bytes, uf.user_data, call_data) != 8)
        return false;
    if (rd_le32(header) != SKIPPABLE_MAGIC || rd_le32(header + 4) != (uint32_t)(table_bytes - 8))
        return false;
    r->c_off = malloc((n + 1) * sizeof(uint64_t));
    r->d_off = malloc((n + 1) * sizeof(uint64_t));
    size_t chunk_entries = (1u << 20) / esz;
    uint8_t *buf = malloc(chunk_entries * esz);
    if (!r->c_off || !r->d_off || !buf) {
        free(buf);
        return false;
    }
    uint64_t c = 0, d = 0, e = 0;
    size_t file_off = (size_t)fsize - table_bytes + 8;
    uint32_t max_c = 0, max_d = 0;
    while (e < n) {
        size_t cnt = MIN(chunk_entries, n - e);
        if (uf.pread(buf, cnt * esz, file_off, uf.user_data, call_data) != (ssize_t)(cnt * esz)) {
            free(buf);
            return false;
        }
        file_off += cnt * esz;
        for (size_t i = 0; i < cnt; i++, e++) {
            uint32_t cs = rd_le32(buf + i * esz), ds = rd_le32(buf + i * esz + 4);
            r->c_off[e] = c;
            r->d_off[e] = d;
            c += cs;
            d += ds;
            max_c = MAX(max_c, cs);
            max_d = MAX(max_d, ds);
        }
    }
    free(buf);
    r->c_off[n] = c;
    r->d_off[n] = d;
    r->nframes = n;
    r->max_csize = max_c;
    r->max_dsize = max_d;
    return true;
}

/* reference src/seek_table.c:187-202: last frame whose start is <= offset; -1 at/after EOF */
static int64_t st_lookup(const zseek_reader_t *r, uint64_t offset)
{
    if (offset >= r->d_off[r->nframes])
        return -1;
    uint64_t lo = 0, hi = r->nframes;
    while (lo + 1 < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (r->d_off[mid] <= offset)
            lo = mid;
        else
            hi = mid;
    }
    return (int64_t)lo;
}

/* ------------------------------------------------------------------ HBM frame cache: O(1) LRU over slots */
static void lru_unlink(zseek_reader_t *r, int32_t s)
{
    int32_t p = r->lru_prev[s], n = r->lru_next[s];
    if (p >= 0) r->lru_next[p] = n; else r->lru_head = n;
    if (n >= 0) r->lru_prev[n] = p; else r->lru_tail = p;
}

static void lru_push_front(zseek_reader_t *r, int32_t s)
{
    r->lru_prev[s] = -1;
    r->lru_next[s] = r->lru_head;
    if (r->lru_head >= 0) r->lru_prev[r->lru_head] = s; else r->lru_tail = s;
    r->lru_head = s;
}

static uint8_t *slot_ptr(zseek_reader_t *r, int32_t s) { return r->g_slab + ZSK_PAD_FRONT + (size_t)s * r->slot_size; }

static void cache_drop_slot(zseek_reader_t *r, int32_t s);

/* slot of frame f if at least its first `need` bytes are decoded there (a batch may have decoded only a prefix);
 * a shorter entry is dropped and reported as a miss */
static int32_t cache_find_prefix(zseek_reader_t *r, uint64_t f, uint32_t need)
{
    int32_t s = r->frame_slot[f];
    if (s >= 0 && r->valid_len[f] < need) {
        cache_drop_slot(r, s);
        return -1;
    }
    if (s >= 0 && r->lru_head != s) { /* promote to MRU */
        lru_unlink(r, s);
        lru_push_front(r, s);
    }
    return s;
}

/* slot of the completely decoded frame f, or -1 */
static int32_t cache_find(zseek_reader_t *r, uint64_t f)
{
    return cache_find_prefix(r, f, (uint32_t)(r->d_off[f + 1] - r->d_off[f]));
}

static void cache_drop_slot(zseek_reader_t *r, int32_t s)
{
    int32_t f = r->slot_frame[s];
    if (f >= 0) {
        r->frame_slot[f] = -1;
        r->valid_len[f] = 0;
        r->h_frame_src[f] = -1;
        r->slot_frame[s] = -1;
        r->cached--;
        r->frame_src_dirty = true;
    }
}

/* takes the LRU slot for frame f (evicting its occupant) and makes it MRU */
static int32_t cache_take(zseek_reader_t *r, uint64_t f)
{
    int32_t s = r->lru_tail;
    cache_drop_slot(r, s);
    lru_unlink(r, s);
    lru_push_front(r, s);
    r->slot_frame[s] = (int32_t)f;
    r->frame_slot[f] = s;
    r->h_frame_src[f] = (int64_t)((size_t)s * r->slot_size);
    r->cached++;
    r->frame_src_dirty = true;
    return s;
}

static void cache_clear(zseek_reader_t *r)
{
    for (uint32_t s = 0; s < r->nslots; s++)
        cache_drop_slot(r, (int32_t)s);
    r->mir_lo = r->mir_hi = 0;
}

/* ------------------------------------------------------------------ parallel file ingest (default FILE* I/O only) */
/* The callback interface (reference src/zseek.h:88-116) hands the library one buffer per call, and the default callbacks
 * (reference src/decompress.c:47-98) seek a shared FILE*: one thread at a time.  When the reader was opened over a FILE*,
 * the library knows the descriptor and may read it with pread(2), which needs no file position: a few worker threads
 * fill one staging half in parallel (page cache -> pinned memory runs at memcpy speed per thread), the caller's thread
 * queues the DMA of the half that is complete. */
typedef struct io_pool {
    pthread_t th[16];
    int n;
    pthread_mutex_t mu;
    pthread_cond_t cv_work, cv_done;
    int fd;
    uint8_t *dst;
    size_t off, len, piece, next;
    int pending, err;
    bool stop;
} io_pool;

static void *io_worker(void *arg)
{
    io_pool *p = arg;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (!p->stop && p->next >= p->len)
            pthread_cond_wait(&p->cv_work, &p->mu);
        if (p->stop)
            break;
        const size_t o = p->next, n = MIN(p->piece, p->len - o);
        p->next += n;
        const int fd = p->fd;
        uint8_t *dst = p->dst + o;
        const size_t foff = p->off + o;
        pthread_mutex_unlock(&p->mu);
        size_t done = 0;
        int err = 0;
        while (done < n) {
            ssize_t k = pread(fd, dst + done, n - done, (off_t)(foff + done));
            if (k < 0 && errno == EINTR)
                continue;
            if (k <= 0) {
                err = k < 0 ? 1 : 2; /* 2: unexpected EOF */
                break;
            }
            done += (size_t)k;
        }
        pthread_mutex_lock(&p->mu);
        if (err && !p->err)
            p->err = err;
        if (--p->pending == 0)
            pthread_cond_signal(&p->cv_done);
    }
    pthread_mutex_unlock(&p->mu);
    return NULL;
}

static io_pool *io_pool_create(int threads)
{
    io_pool *p = calloc(1, sizeof(*p));
    if (!p)
        return NULL;
    pthread_mutex_init(&p->mu, NULL);
    pthread_cond_init(&p->cv_work, NULL);
    pthread_cond_init(&p->cv_done, NULL);
    for (p->n = 0; p->n < threads && p->n < 16; p->n++)
        if (pthread_create(&p->th[p->n], NULL, io_worker, p))
            break;
    if (p->n == 0) {
        free(p);
        return NULL;
    }
    return p;
}

static void io_pool_destroy(io_pool *p)
{
    if (!p)
        return;
    pthread_mutex_lock(&p->mu);
    p->stop = true;
    pthread_cond_broadcast(&p->cv_work);
    pthread_mutex_unlock(&p->mu);
    for (int i = 0; i < p->n; i++)
        pthread_join(p->th[i], NULL);
    pthread_mutex_destroy(&p->mu);
    pthread_cond_destroy(&p->cv_work);
    pthread_cond_destroy(&p->cv_done);
    free(p);
}

/* file bytes [off, off + len) -> dst, all workers; 0 ok, 1 read error, 2 unexpected EOF */
static int io_pool_read(io_pool *p, int fd, uint8_t *dst, size_t off, size_t len)
{
    pthread_mutex_lock(&p->mu);
    p->fd = fd;
    p->dst = dst;
    p->off = off;
    p->len = len;
    p->next = 0;
    p->piece = MAX((size_t)1 << 20, (len / (size_t)p->n + 4095) & ~(size_t)4095);
    p->pending = (int)((len + p->piece - 1) / p->piece);
    p->err = 0;
    pthread_cond_broadcast(&p->cv_work);
    while (p->pending)
        pthread_cond_wait(&p->cv_done, &p->mu);
    const int err = p->err;
    p->len = p->next = 0;
    pthread_mutex_unlock(&p->mu);
    return err;
}

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

/* ------------------------------------------------------------------ pinned host buffers, allocated on first use */
/* two pinned staging halves, each large enough for min(bytes, stage_half_max): a reader that only ever pulls small windows
 * through its callback never pins the full 2 x 32 MiB */
static bool ensure_stage(zseek_reader_t *r, size_t bytes, char *errbuf)
{
    size_t want = MIN(r->stage_half_max, MAX(MAX(bytes, (size_t)r->max_csize), (size_t)1 << 20));
    if (r->h_stage && want <= r->stage_half)
        return true;
    want = MIN(r->stage_half_max, MAX(want, 2 * r->stage_half));
    if (r->h_stage) {
        if (zsk_cuda_stream_sync(r->cx, ZSK_STREAM_H2D)) return cuda_fail(r, errbuf, "synchronize");
        r->stage_inflight = 0;
        zsk_cuda_free_host(r->cx, r->h_stage);
        r->h_stage = NULL;
    }
    r->stage_half = 0;
    if (zsk_cuda_malloc_host(r->cx, (void **)&r->h_stage, 2 * want))
        return cuda_fail(r, errbuf, "allocate pinned staging");
    r->stage_half = want;
    r->stage_next = 0;
    return true;
}

/* most frames an ordinary read-ahead window may hold right now: the reader's own maximum, cut down so that the two pinned
 * windows of every live reader of the process fit ZSEEK_B200_WINDOW_BUDGET_MB (default 2048) together */
static uint32_t ra_limit(zseek_reader_t *r)
{
    static size_t budget;
    if (!budget)
        budget = MAX(env_size("ZSEEK_B200_WINDOW_BUDGET_MB", 2048), 1) << 20;
    if (!r->ra_cap) {
        /* decided once per reader, when its scan first grows a window, and rounded down to a power of two: the pinned
         * windows of the readers of a thread pool then come in one size and are recycled from scan to scan */
        const unsigned live = MAX(atomic_load(&g_live_readers), 1u);
        size_t per_window = (size_t)8 << 20;
        while (per_window * 2 <= budget / (2 * (size_t)live))
            per_window *= 2;
        const size_t frames = per_window / MAX((size_t)r->max_dsize, 1);
        r->ra_cap = (uint32_t)MAX(1, MIN((size_t)r->ra_max, frames));
    }
    return r->ra_cap;
}

/* pinned window `which` of at least `bytes` (grown geometrically up to window_cap: a reader that only ever sees small
 * windows never pins the 2 x 128 MiB of the largest one).  The window must not be the one reads are served from. */
static bool ensure_window(zseek_reader_t *r, int which, size_t bytes, char *errbuf)
{
    if (r->h_mir[which] && bytes <= r->mir_cap[which])
        return true;
    zsk_cuda_stream_sync(r->cx, ZSK_STREAM_D2H);
    /* two sizes only — 4 MiB for the first windows of a scan (and for random readers), then the largest window this reader
     * may use right now — so that a scan pins each of its windows at most twice */
    size_t cap = MAX(bytes, (size_t)4 << 20);
    if (bytes > ((size_t)4 << 20))
        cap = MAX(bytes, MIN(r->window_cap, (size_t)ra_limit(r) * r->max_dsize));
    zsk_cuda_free_host(r->cx, r->h_mir[which]);
    r->h_mir[which] = NULL;
    r->mir_cap[which] = 0;
    if (zsk_cuda_malloc_host(r->cx, (void **)&r->h_mir[which], cap))
        return cuda_fail(r, errbuf, "allocate pinned window");
    r->mir_cap[which] = cap;
    return true;
}

/* ------------------------------------------------------------------ compressed image residency */
/* Queues the copy of file bytes [file_off, file_off + bytes) to device memory `dst` on the H2D
 * stream: one DMA straight from the memory image, or pread-callback -> pinned staging halves -> DMA. */
static bool h2d_range(zseek_reader_t *r, size_t file_off, size_t bytes, uint8_t *dst, void *call_data, char *errbuf)
{
    if (r->mem_image) {
        if (file_off + bytes > r->mem_size) {
            set_error(errbuf, "unexpected EOF");
            return false;
        }
        if (zsk_cuda_memcpy_async(r->cx, dst, r->mem_image + file_off, bytes, ZSK_H2D, ZSK_STREAM_H2D))
            return cuda_fail(r, errbuf, "copy image to device");
        return true;
    }
    if (!ensure_stage(r, bytes, errbuf))
        return false;
    size_t done = 0;
    while (done < bytes) {
        size_t n = MIN(r->stage_half, bytes - done);
        if (r->stage_inflight == 2) { /* the half about to be refilled may still be in flight */
            if (zsk_cuda_stream_sync(r->cx, ZSK_STREAM_H2D)) return cuda_fail(r, errbuf, "synchronize");
            r->stage_inflight = 0;
        }
        uint8_t *st = r->h_stage + (size_t)r->stage_next * r->stage_half;
        ssize_t got;
        if (r->file_fd >= 0 && n >= ((size_t)4 << 20) && (r->io || (r->io = io_pool_create((int)env_size("ZSEEK_B200_IO_THREADS", 8))))) {
            const int e = io_pool_read(r->io, r->file_fd, st, file_off + done, n);
            got = e == 0 ? (ssize_t)n : e == 2 ? 0 : -1;
        } else
            got = r->user_file.pread(st, n, file_off + done, r->user_file.user_data, call_data);
        if (got != (ssize_t)n) {
            zsk_cuda_stream_sync(r->cx, ZSK_STREAM_H2D);
            r->stage_inflight = 0;
            set_error(errbuf, got >= 0 ? "unexpected EOF" : "read file failed");
            return false;
        }
        if (zsk_cuda_memcpy_async(r->cx, dst + done, st, n, ZSK_H2D, ZSK_STREAM_H2D))
            return cuda_fail(r, errbuf, "copy frames to device");
        done += n;
        r->stage_next ^= 1;
        r->stage_inflight++;
    }
    return true;
}

/* (re)allocates the device image for frames [lo, hi) without filling it; residency is cleared */
static bool alloc_image(zseek_reader_t *r, uint64_t lo, uint64_t hi, char *errbuf)
{
    size_t need = (size_t)(r->c_off[hi] - r->c_off[lo]) + ZSK_PAD_FRONT + ZSK_PAD_BACK;
    /* everything queued against the old image must have finished before it is replaced */
    if (zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE)) return cuda_fail(r, errbuf, "synchronize");
    r->res_lo = r->res_hi = 0;
    if (need > r->g_comp_cap) {
        /* grow geometrically (bounded by the whole payload): a growing read-ahead window must not reallocate every time */
        size_t whole = (size_t)r->c_off[r->nframes] + ZSK_PAD_FRONT + ZSK_PAD_BACK;
        size_t want = MIN(MAX(need, 2 * r->g_comp_cap), MAX(whole, need));
        zsk_cuda_free(r->cx, r->g_comp);
        r->g_comp = NULL;
        r->g_comp_cap = 0;
        if (zsk_cuda_malloc(r->cx, (void **)&r->g_comp, want)) return cuda_fail(r, errbuf, "allocate device image");
        r->g_comp_cap = want;
    }
    return true;
}

static bool ensure_resident(zseek_reader_t *r, uint64_t lo, uint64_t hi, void *call_data, char *errbuf)
{
    if (lo >= hi || (lo >= r->res_lo && hi <= r->res_hi))
        return true;
    if (!alloc_image(r, lo, hi, errbuf))
        return false;
    if (!h2d_range(r, (size_t)r->c_off[lo], (size_t)(r->c_off[hi] - r->c_off[lo]), r->g_comp + ZSK_PAD_FRONT, call_data, errbuf))
        return false;
    /* kernels queued later on the compute stream see the complete image */
    if (zsk_cuda_stream_wait(r->cx, ZSK_STREAM_COMPUTE, ZSK_STREAM_H2D)) return cuda_fail(r, errbuf, "order streams");
    if (!r->mem_image) {
        if (zsk_cuda_stream_sync(r->cx, ZSK_STREAM_H2D)) return cuda_fail(r, errbuf, "synchronize");
        r->stage_inflight = 0;
    }
    r->res_lo = lo;
    r->res_hi = hi;
    return true;
}

static void fill_decode_args(zseek_reader_t *r, zsk_decode_args *a)
{
    memset(a, 0, sizeof(*a));
    a->c_off = r->g_coff;
    a->d_off = r->g_doff;
    a->comp = r->g_comp + ZSK_PAD_FRONT;
    a->comp_base = r->c_off[r->res_lo];
}

static bool ensure_jobs(zseek_reader_t *r, uint32_t n, char *errbuf)
{
    if (n <= r->job_cap)
        return true;
    uint32_t cap = MAX(n, 1024u);
    zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
    zsk_cuda_free(r->cx, r->g_job_ids); zsk_cuda_free(r->cx, r->g_job_offs); zsk_cuda_free(r->cx, r->g_job_status);
    zsk_cuda_free(r->cx, r->g_job_limits);
    zsk_cuda_free_host(r->cx, r->h_job_ids); zsk_cuda_free_host(r->cx, r->h_job_offs); zsk_cuda_free_host(r->cx, r->h_job_status);
    zsk_cuda_free_host(r->cx, r->h_job_limits);
    r->g_job_ids = NULL; r->g_job_offs = NULL; r->g_job_status = NULL; r->g_job_limits = NULL;
    r->h_job_ids = NULL; r->h_job_offs = NULL; r->h_job_status = NULL; r->h_job_limits = NULL;
    r->job_cap = 0;
    r->sorted_lo = r->sorted_hi = 0;
    if (zsk_cuda_malloc(r->cx, (void **)&r->g_job_ids, cap * sizeof(uint32_t)) ||
        zsk_cuda_malloc(r->cx, (void **)&r->g_job_offs, cap * sizeof(uint64_t)) ||
        zsk_cuda_malloc(r->cx, (void **)&r->g_job_status, cap * sizeof(int32_t)) ||
        zsk_cuda_malloc(r->cx, (void **)&r->g_job_limits, cap * sizeof(uint32_t)) ||
        zsk_cuda_malloc_host(r->cx, (void **)&r->h_job_limits, cap * sizeof(uint32_t)) ||
        zsk_cuda_malloc_host(r->cx, (void **)&r->h_job_ids, cap * sizeof(uint32_t)) ||
        zsk_cuda_malloc_host(r->cx, (void **)&r->h_job_offs, cap * sizeof(uint64_t)) ||
        zsk_cuda_malloc_host(r->cx, (void **)&r->h_job_status, cap * sizeof(int32_t)))
        return cuda_fail(r, errbuf, "allocate job buffers");
    r->job_cap = cap;
    return true;
}

/* waits for the compute stream and turns the first non-zero frame status into an error */
static bool finish_decode(zseek_reader_t *r, uint32_t njobs, char *errbuf)
{
    if (zsk_cuda_memcpy_async(r->cx, r->h_job_status, r->g_job_status, njobs * sizeof(int32_t), ZSK_D2H, ZSK_STREAM_COMPUTE) ||
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE))
        return cuda_fail(r, errbuf, "decompress frame");
    for (uint32_t i = 0; i < njobs; i++)
        if (r->h_job_status[i] != ZSK_ST_OK) {
            if (getenv("ZSEEK_B200_DEBUG")) {
                uint32_t bad = 0;
                for (uint32_t k = 0; k < njobs; k++) bad += r->h_job_status[k] != ZSK_ST_OK;
                fprintf(stderr, "[zsk] job %u of %u failed with status %d (%u failed jobs in this launch)\n", i, njobs, r->h_job_status[i], bad);
            }
            set_error(errbuf, "decompress frame: %s", status_name(r->h_job_status[i]));
            return false;
        }
    return true;
}

/* Decodes the frames listed in h_job_ids[0..n) (not cached yet) into freshly taken cache slots.  need != NULL: frame f
 * is only needed up to byte need[f]; the kernels may stop there (the slot then holds a valid prefix). */
static bool decode_into_cache(zseek_reader_t *r, uint32_t n, const uint32_t *need, char *errbuf)
{
    if (n == 0)
        return true;
    r->sorted_lo = r->sorted_hi = 0; /* the job list is about to be overwritten */
    uint64_t dsum = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t f = r->h_job_ids[i];
        int32_t s = cache_take(r, f);
        r->h_job_offs[i] = (uint64_t)((size_t)s * r->slot_size);
        r->h_job_limits[i] = need ? need[f] : 0xffffffffu;
        dsum += r->d_off[f + 1] - r->d_off[f];
    }
    zsk_decode_args a;
    fill_decode_args(r, &a);
    a.dsize_sum = dsum;
    a.frame_ids = r->g_job_ids;
    a.dst_offs = r->g_job_offs;
    a.dst = r->g_slab + ZSK_PAD_FRONT;
    a.njobs = n;
    a.status = r->g_job_status;
    a.limits = need ? r->g_job_limits : NULL;
    bool ok = !zsk_cuda_memcpy_async(r->cx, r->g_job_ids, r->h_job_ids, n * sizeof(uint32_t), ZSK_H2D, ZSK_STREAM_COMPUTE) &&
              !zsk_cuda_memcpy_async(r->cx, r->g_job_offs, r->h_job_offs, n * sizeof(uint64_t), ZSK_H2D, ZSK_STREAM_COMPUTE) &&
              (!need || !zsk_cuda_memcpy_async(r->cx, r->g_job_limits, r->h_job_limits, n * sizeof(uint32_t), ZSK_H2D, ZSK_STREAM_COMPUTE)) &&
              !zsk_cuda_launch_decode(r->cx, r->codec, &a, ZSK_STREAM_COMPUTE);
    if (!ok)
        cuda_fail(r, errbuf, "decompress frame");
    else
        ok = finish_decode(r, n, errbuf);
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t f = r->h_job_ids[i];
        if (!ok) /* never leave undecoded slots marked valid */
            cache_drop_slot(r, r->frame_slot[f]);
        else
            r->valid_len[f] = (uint32_t)MIN((uint64_t)r->h_job_limits[i], r->d_off[f + 1] - r->d_off[f]);
    }
    return ok;
}

/* Makes frames [lo, hi) (hi - lo <= nslots) resident in the decoded-frame cache; with `mirror` also
 * copies their bytes into the pinned host window so later host reads are plain memcpy. */
static bool fill_window(zseek_reader_t *r, uint64_t lo, uint64_t hi, bool mirror, void *call_data, char *errbuf)
{
    uint32_t n = 0;
    if (!ensure_jobs(r, (uint32_t)(hi - lo), errbuf))
        return false;
    for (uint64_t f = lo; f < hi; f++)
        if (cache_find(r, f) < 0) /* also promotes the hits so that cache_take cannot evict them */
            r->h_job_ids[n++] = (uint32_t)f;
    if (n) {
        uint64_t mlo = r->h_job_ids[0], mhi = (uint64_t)r->h_job_ids[n - 1] + 1;
        if (!ensure_resident(r, mlo, mhi, call_data, errbuf))
            return false;
        if (!decode_into_cache(r, n, NULL, errbuf))
            return false;
    }
    if (mirror) {
        r->mir_lo = r->mir_hi = 0;
        for (uint64_t f = lo; f < hi; f++) {
            size_t dsz = (size_t)(r->d_off[f + 1] - r->d_off[f]);
            if (zsk_cuda_memcpy_async(r->cx, mirror_half(r, r->mir_cur) + (r->d_off[f] - r->d_off[lo]), slot_ptr(r, r->frame_slot[f]), dsz,
                                      ZSK_D2H, ZSK_STREAM_COMPUTE))
                return cuda_fail(r, errbuf, "copy frame to host");
        }
        if (zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE))
            return cuda_fail(r, errbuf, "copy frame to host");
        r->mir_lo = lo;
        r->mir_hi = hi;
    }
    return true;
}

static int buf_on_device(zseek_reader_t *r, const void *buf)
{
    const uint64_t block = ((uint64_t)(uintptr_t)buf >> 21) + 1; /* +1: 0 marks an empty cache entry */
    const unsigned i = (unsigned)(block & 7);
    const uint64_t e = atomic_load_explicit(&r->ptr_cache[i], memory_order_relaxed);
    if ((e >> 1) == block)
        return (int)(e & 1);
    int d = zsk_cuda_pointer_is_device(r->cx, buf);
    atomic_store_explicit(&r->ptr_cache[i], (block << 1) | (uint64_t)(d > 0), memory_order_relaxed);
    return d > 0;
}

static bool in_shard(zseek_reader_t *r, uint64_t f, char *errbuf)
{
    if (f >= r->shard_lo && f < r->shard_hi)
        return true;
    set_error(errbuf, "frame outside this reader's shard");
    return false;
}

/* ------------------------------------------------------------------ open / close */
/* waits for everything the reader queued and takes it out of the process-wide counts */
static void drop_resident(zseek_reader_t *r);
static void reader_quiesce(zseek_reader_t *r)
{
    if (r->cx) {
        prefetch_drop(r);
        drop_resident(r);
    }
    if (r->counted)
        atomic_fetch_sub(&g_live_readers, 1u);
    r->counted = false;
    if (r->resident_bytes)
        atomic_fetch_sub(&g_resident_bytes, r->resident_bytes);
    r->resident_bytes = 0;
    io_pool_destroy(r->io);
    r->io = NULL;
    if (r->cx && r->async_stream >= 0)
        zsk_cuda_stream_sync(r->cx, r->async_stream);
    r->async_stream = -1;
    if (r->cx) {
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_H2D);
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_D2H);
    }
}

static void free_file_state(zseek_reader_t *r)
{
    free(r->c_off); free(r->d_off);
    free(r->slot_frame); free(r->lru_prev); free(r->lru_next); free(r->frame_slot); free(r->valid_len);
    free(r->h_frame_src); free(r->h_touched);
    r->c_off = r->d_off = NULL;
    r->slot_frame = r->lru_prev = r->lru_next = r->frame_slot = NULL;
    r->valid_len = NULL;
    r->h_frame_src = NULL;
    r->h_touched = NULL;
}

/* Keeps the device side of a closed reader for the next open: context, seek-table / slab / image / staging / job buffers
 * and the pinned windows with their capacities; everything that describes the file is dropped.  Large buffers are given
 * back so that a parked reader holds at most a few hundred MiB.  false: not parkable (the caller destroys it). */
static bool reader_park(zseek_reader_t *r)
{
    static int enabled = -1;
    if (enabled < 0)
        enabled = env_size("ZSEEK_B200_PARK", 1) != 0;
    if (!enabled || !r->cx)
        return false;
    reader_quiesce(r);
    free_file_state(r);
    const size_t keep_max = ((size_t)512 << 20) + 4096;
    /* buffers whose size depends on the file's frame geometry in ways the next file need not share, and the big ones */
    void *drop[] = { r->g_touched, r->g_bslab, r->g_bjob_ids, r->g_bjob_limits, r->g_bctl, r->g_bjob_offs, r->g_bjob_status, r->g_bsrc,
                     r->g_comp_cap > keep_max ? r->g_comp : NULL, r->g_out_cap > keep_max ? r->g_out : NULL,
                     r->slab_bytes > keep_max ? r->g_slab : NULL };
    for (size_t i = 0; i < sizeof(drop) / sizeof(drop[0]); i++)
        zsk_cuda_free(r->cx, drop[i]);
    zseek_reader_t k;
    memset(&k, 0, sizeof(k));
    k.cx = r->cx;
    k.g_coff = r->g_coff; k.g_doff = r->g_doff; k.g_frame_src = r->g_frame_src; k.n_cap = r->n_cap;
    if (r->slab_bytes <= keep_max) { k.g_slab = r->g_slab; k.slab_bytes = r->slab_bytes; }
    if (r->g_comp_cap <= keep_max) { k.g_comp = r->g_comp; k.g_comp_cap = r->g_comp_cap; }
    if (r->g_out_cap <= keep_max) { k.g_out = r->g_out; k.g_out_cap = r->g_out_cap; }
    k.h_stage = r->h_stage; k.stage_half = r->stage_half;
    k.h_bounce = r->h_bounce;
    for (int w = 0; w < 2; w++) {
        if (r->mir_cap[w] <= keep_max) { k.h_mir[w] = r->h_mir[w]; k.mir_cap[w] = r->mir_cap[w]; }
        else zsk_cuda_free_host(r->cx, r->h_mir[w]); /* a whole-shard window: back to the pinned cache */
    }
    k.job_cap = r->job_cap;
    k.g_job_ids = r->g_job_ids; k.g_job_offs = r->g_job_offs; k.g_job_limits = r->g_job_limits; k.g_job_status = r->g_job_status;
    k.h_job_ids = r->h_job_ids; k.h_job_offs = r->h_job_offs; k.h_job_limits = r->h_job_limits; k.h_job_status = r->h_job_status;
    k.batch_cap = r->batch_cap;
    k.g_b_offsets = r->g_b_offsets; k.g_b_counts = r->g_b_counts; k.g_b_dstoffs = r->g_b_dstoffs;
    k.g_b_frame = r->g_b_frame; k.g_b_inframe = r->g_b_inframe; k.g_b_nbytes = r->g_b_nbytes;
    zsk_cuda_ctx_trim(k.cx, keep_max);
    k.parked_bytes = k.g_comp_cap + k.g_out_cap + k.slab_bytes + 24 * k.n_cap + zsk_cuda_ctx_held(k.cx);
    pthread_mutex_destroy(&r->lock);
    pthread_rwlock_destroy(&r->res_lock);
    *r = k;
    static size_t park_max;
    if (!park_max)
        park_max = MAX(env_size("ZSEEK_B200_PARK_MB", 8192), 1) << 20;
    pthread_mutex_lock(&g_park_mu);
    const bool parked = g_nparked < ZSK_MAX_PARKED && g_parked_bytes + r->parked_bytes <= park_max;
    if (parked) {
        g_parked[g_nparked++] = r;
        g_parked_bytes += r->parked_bytes;
    }
    pthread_mutex_unlock(&g_park_mu);
    if (!parked) { /* reader_free destroys them */
        pthread_mutex_init(&r->lock, NULL);
        pthread_rwlock_init(&r->res_lock, NULL);
    }
    return parked;
}

static void reader_free(zseek_reader_t *r);

/* a zeroed reader, or a parked one of the device the next context would use */
static zseek_reader_t *reader_new(void)
{
    zseek_reader_t *r = NULL;
    pthread_mutex_lock(&g_park_mu);
    const int dev = g_nparked ? zsk_cuda_pick_device() : -1; /* no device call before the file has been looked at otherwise */
    for (int i = g_nparked - 1; i >= 0 && dev >= 0; i--)
        if (zsk_cuda_device(g_parked[i]->cx) == dev) {
            r = g_parked[i];
            g_parked[i] = g_parked[--g_nparked];
            g_parked_bytes -= r->parked_bytes;
            break;
        }
    pthread_mutex_unlock(&g_park_mu);
    if (r && zsk_cuda_ctx_reuse(r->cx)) { /* the parked context is unusable: start afresh */
        pthread_mutex_init(&r->lock, NULL);
        pthread_rwlock_init(&r->res_lock, NULL);
        reader_free(r);
        r = NULL;
    }
    if (!r)
        r = calloc(1, sizeof(*r));
    if (r) {
        pthread_mutex_init(&r->lock, NULL);
        pthread_rwlock_init(&r->res_lock, NULL);
    }
    return r;
}

static void reader_free(zseek_reader_t *r)
{
    if (!r)
        return;
    reader_quiesce(r);
    if (r->cx) {
        void *dev[] = { r->g_coff, r->g_doff, r->g_comp, r->g_slab, r->g_frame_src, r->g_job_ids, r->g_job_offs, r->g_job_status, r->g_job_limits,
                        r->g_b_offsets, r->g_b_counts, r->g_b_dstoffs, r->g_b_frame, r->g_b_inframe, r->g_b_nbytes, r->g_touched,
                        r->g_out, r->g_bslab, r->g_bjob_ids, r->g_bjob_limits, r->g_bctl, r->g_bjob_offs, r->g_bjob_status, r->g_bsrc };
        for (size_t i = 0; i < sizeof(dev) / sizeof(dev[0]); i++)
            zsk_cuda_free(r->cx, dev[i]);
        void *pin[] = { r->h_stage, r->h_mir[0], r->h_mir[1], r->h_job_ids, r->h_job_offs, r->h_job_status, r->h_job_limits, r->h_bounce };
        for (size_t i = 0; i < sizeof(pin) / sizeof(pin[0]); i++)
            zsk_cuda_free_host(r->cx, pin[i]);
        zsk_cuda_ctx_destroy(r->cx);
    }
    free_file_state(r);
    pthread_mutex_destroy(&r->lock);
    pthread_rwlock_destroy(&r->res_lock);
    free(r);
}

static size_t env_size(const char *name, size_t dflt)
{
    const char *s = getenv(name);
    if (!s || !*s)
        return dflt;
    return (size_t)strtoull(s, NULL, 10);
}

static zseek_reader_t *reader_open_common(zseek_reader_t *r, size_t cache_size, void *call_data, char *errbuf)
{
    /* magic sniff, reference src/decompress.c:264-287 */
    uint8_t magic_le[4];
    ssize_t got = r->user_file.pread(magic_le, 4, 0, r->user_file.user_data, call_data);
    if (got != 4) {
        set_error(errbuf, got >= 0 ? "unexpected EOF" : "read file failed");
        goto fail;
    }
    uint32_t magic = rd_le32(magic_le);
    if (magic == ZSTD_MAGIC)
        r->codec = ZSK_CODEC_ZSTD;
    else if (magic == LZ4_MAGIC)
        r->codec = ZSK_CODEC_LZ4;
    else {
        set_error(errbuf, "unrecognized file format");
        goto fail;
    }
    if (!st_load(r, call_data)) {
        set_error(errbuf, "read_seek_table failed");
        goto fail;
    }
    if (r->nframes > 0x7fffffffu) {
        set_error(errbuf, "too many frames");
        goto fail;
    }
    char derr[128];
    if (!r->cx && zsk_cuda_ctx_create(-1, &r->cx, derr, sizeof(derr))) { /* no CPU fallback: fail loudly (a parked reader brings its context) */
        set_error(errbuf, "context creation failed: %s", derr);
        goto fail;
    }
    uint64_t N = r->nframes;
    r->async_stream = -1;
    r->shard_lo = 0;
    r->shard_hi = N;
    r->user_cache_size = cache_size;
    /* read-ahead window of a sequential scan: up to ~128 MiB of decoded frames (grows x4 per sequential miss) */
    size_t ra = r->max_dsize ? (128u << 20) / r->max_dsize : 1;
    ra = MAX(1, MIN(ra, 8192));
    ra = env_size("ZSEEK_B200_READAHEAD", ra);
    r->ra_max = (uint32_t)MAX(1, MIN(ra, 65536));
    if (N && r->ra_max > N)
        r->ra_max = (uint32_t)N;
    r->ra_window = 1;
    r->ra_next = UINT64_MAX;
    r->resident_after = (uint32_t)env_size("ZSEEK_B200_RESIDENT_AFTER", 32);
    r->resident_max = env_size("ZSEEK_B200_RESIDENT_MB", 4096) << 20;
    /* decoded-frame cache in HBM: what the caller asked for, and never fewer than 64 slots for device-side readers */
    r->nslots = (uint32_t)MAX(MAX(cache_size, 64), 1);
    if (N && r->nslots > N)
        r->nslots = (uint32_t)N;
    r->slot_size = ((size_t)r->max_dsize + 255) & ~(size_t)255;
    if (r->slot_size == 0)
        r->slot_size = 256;
    r->stage_half_max = env_size("ZSEEK_B200_STAGE_MB", 64) * (1u << 20) / 2;
    if (r->stage_half_max < r->max_csize)
        r->stage_half_max = r->max_csize;
    r->window_cap = (size_t)r->ra_max * r->max_dsize;
    r->chunk_bytes = env_size("ZSEEK_B200_CHUNK_MB", 512) << 20;
    r->ramp_bytes = env_size("ZSEEK_B200_RAMP_MB", 16) << 20;
    r->sort_min = env_size("ZSEEK_B200_SORT_MIN", 40960);
    r->partial_decode = env_size("ZSEEK_B200_PARTIAL", 1) != 0;
    r->sort_min_zstd = env_size("ZSEEK_B200_SORT_MIN_ZSTD", 2048); /* zstd: one CTA per frame, largest frames first trims the last wave */

    r->slot_frame = malloc(r->nslots * sizeof(int32_t));
    r->lru_prev = malloc(r->nslots * sizeof(int32_t));
    r->lru_next = malloc(r->nslots * sizeof(int32_t));
    r->frame_slot = malloc((N + 1) * sizeof(int32_t));
    r->valid_len = calloc(N + 1, sizeof(uint32_t));
    r->h_frame_src = malloc((N + 1) * sizeof(int64_t));
    r->h_touched = malloc((N + 1) * sizeof(uint32_t));
    if (!r->slot_frame || !r->lru_prev || !r->lru_next || !r->frame_slot || !r->valid_len || !r->h_frame_src || !r->h_touched) {
        set_error(errbuf, "cache creation failed");
        goto fail;
    }
    r->lru_head = r->lru_tail = -1;
    for (uint32_t s = 0; s < r->nslots; s++) {
        r->slot_frame[s] = -1;
        lru_push_front(r, (int32_t)s);
    }
    for (uint64_t f = 0; f < N; f++) {
        r->frame_slot[f] = -1;
        r->h_frame_src[f] = -1;
    }
    r->frame_src_dirty = true;

    /* the pinned ingest staging (h_stage) and the pinned decoded windows (h_mir[]) are allocated on first use: readers
     * that only serve device buffers or batches never pay for them.  A parked reader's buffers are kept when they are
     * large enough. */
    if (r->n_cap < N + 1) {
        zsk_cuda_free(r->cx, r->g_coff); zsk_cuda_free(r->cx, r->g_doff); zsk_cuda_free(r->cx, r->g_frame_src);
        r->g_coff = r->g_doff = NULL;
        r->g_frame_src = NULL;
        r->n_cap = 0;
        const size_t cap = MAX(N + 1, 4096);
        if (zsk_cuda_malloc(r->cx, (void **)&r->g_coff, cap * sizeof(uint64_t)) ||
            zsk_cuda_malloc(r->cx, (void **)&r->g_doff, cap * sizeof(uint64_t)) ||
            zsk_cuda_malloc(r->cx, (void **)&r->g_frame_src, cap * sizeof(int64_t))) {
            set_error(errbuf, "buffer creation failed: %s", zsk_cuda_error(r->cx));
            goto fail;
        }
        r->n_cap = cap;
    }
    const size_t slab_need = (size_t)r->nslots * r->slot_size + ZSK_PAD_FRONT + ZSK_PAD_BACK;
    if (r->slab_bytes < slab_need) {
        zsk_cuda_free(r->cx, r->g_slab);
        r->g_slab = NULL;
        r->slab_bytes = 0;
        if (zsk_cuda_malloc(r->cx, (void **)&r->g_slab, slab_need)) {
            set_error(errbuf, "buffer creation failed: %s", zsk_cuda_error(r->cx));
            goto fail;
        }
        r->slab_bytes = slab_need;
    }
    if (zsk_cuda_memcpy_async(r->cx, r->g_coff, r->c_off, (N + 1) * sizeof(uint64_t), ZSK_H2D, ZSK_STREAM_COMPUTE) ||
        zsk_cuda_memcpy_async(r->cx, r->g_doff, r->d_off, (N + 1) * sizeof(uint64_t), ZSK_H2D, ZSK_STREAM_COMPUTE) ||
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE)) {
        set_error(errbuf, "seek table upload failed: %s", zsk_cuda_error(r->cx));
        goto fail;
    }
    atomic_fetch_add(&g_live_readers, 1u);
    r->counted = true;
    return r;
fail:
    reader_free(r);
    return NULL;
}

zseek_reader_t *zseek_reader_open_full(zseek_read_file_t user_file, size_t cache_size, void *call_data,
                                       char errbuf[ZSEEK_ERRBUF_SIZE])
{
    zseek_reader_t *r = reader_new();
    if (!r) {
        set_error(errbuf, "allocate reader: %s", strerror(errno));
        return NULL;
    }
    r->user_file = user_file;
    r->file_fd = -1;
    return reader_open_common(r, cache_size, call_data, errbuf);
}

zseek_reader_t *zseek_reader_open(FILE *cfile, size_t cache_size, void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    zseek_reader_t *r = reader_new();
    if (!r) {
        set_error(errbuf, "allocate reader: %s", strerror(errno));
        return NULL;
    }
    r->user_file = (zseek_read_file_t){ cfile, file_pread, file_fsize };
    r->file_fd = cfile ? fileno(cfile) : -1; /* large ingests read the descriptor with pread(2) from worker threads */
    return reader_open_common(r, cache_size, call_data, errbuf);
}

zseek_reader_t *zseek_b200_reader_open_mem(const void *image, size_t size, size_t cache_size, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    zseek_reader_t *r = reader_new();
    if (!r) {
        set_error(errbuf, "allocate reader: %s", strerror(errno));
        return NULL;
    }
    r->mem_image = image;
    r->mem_size = size;
    r->file_fd = -1;
    r->user_file = (zseek_read_file_t){ r, mem_pread, mem_fsize };
    return reader_open_common(r, cache_size, NULL, errbuf);
}

bool zseek_reader_close(zseek_reader_t *reader, void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    (void)call_data;
    (void)errbuf;
    if (!reader)
        return true;
    if (!reader_park(reader))
        reader_free(reader);
    return true;
}

static bool stream_frames_to_host(zseek_reader_t *r, uint64_t lo, uint64_t hi, uint8_t *dst, void *call_data, char *errbuf);

/* Decodes the whole shard into one pinned window; false (and the ordinary paths take over) when memory is short or a
 * frame of the shard is bad. */
static bool go_resident(zseek_reader_t *r, void *call_data)
{
    char scratch[ZSEEK_ERRBUF_SIZE];
    const size_t bytes = (size_t)(r->d_off[r->shard_hi] - r->d_off[r->shard_lo]);
    if (bytes == 0)
        return false;
    /* the budget (ZSEEK_B200_RESIDENT_MB) is shared by the readers of the process: sixteen readers of one file must not
     * pin sixteen decoded copies of it */
    if (atomic_fetch_add(&g_resident_bytes, bytes) + bytes > r->resident_max) {
        atomic_fetch_sub(&g_resident_bytes, bytes);
        return false;
    }
    r->resident_bytes = bytes;
    prefetch_drop(r);
    zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
    zsk_cuda_stream_sync(r->cx, ZSK_STREAM_D2H);
    uint8_t *win = NULL;
    if (zsk_cuda_malloc_host(r->cx, (void **)&win, bytes) ||
        !stream_frames_to_host(r, r->shard_lo, r->shard_hi, win, call_data, scratch)) {
        zsk_cuda_free_host(r->cx, win);
        atomic_fetch_sub(&g_resident_bytes, r->resident_bytes);
        r->resident_bytes = 0;
        return false;
    }
    zsk_cuda_free_host(r->cx, r->h_mir[0]);
    zsk_cuda_free_host(r->cx, r->h_mir[1]);
    r->h_mir[0] = win; /* with the whole shard in the window no read-ahead is ever started */
    r->h_mir[1] = NULL;
    r->mir_cap[0] = bytes;
    r->mir_cap[1] = 0;
    r->mir_cur = 0;
    r->mir_lo = r->shard_lo;
    r->mir_hi = r->shard_hi;
    pthread_rwlock_wrlock(&r->res_lock);
    r->resident = true;
    pthread_rwlock_unlock(&r->res_lock);
    atomic_store(&r->resident_fast, true);
    return true;
}

/* Decodes the whole shard into device memory (one launch per 4 GiB of compressed input); afterwards the compressed image
 * and the scratch pools are given back: nothing is decoded again until the shard changes. */
static bool read_range_device(zseek_reader_t *r, uint8_t *dst, size_t count, size_t offset, void *call_data, char *errbuf);
static bool go_resident_hbm(zseek_reader_t *r, void *call_data)
{
    static size_t budget;
    if (!budget)
        budget = MAX(env_size("ZSEEK_B200_RESIDENT_HBM_MB", 32768), 1) << 20;
    char scratch[ZSEEK_ERRBUF_SIZE];
    const size_t bytes = (size_t)(r->d_off[r->shard_hi] - r->d_off[r->shard_lo]);
    if (bytes == 0)
        return false;
    if (atomic_fetch_add(&g_hbm_resident_bytes, bytes) + bytes > budget) {
        atomic_fetch_sub(&g_hbm_resident_bytes, bytes);
        return false;
    }
    r->hbm_resident_bytes = bytes;
    prefetch_drop(r);
    /* room for the decoded shard, its compressed image and the scratch of one decode wave */
    const size_t need = bytes + (size_t)(r->c_off[r->shard_hi] - r->c_off[r->shard_lo]) + 4 * MIN(bytes, (size_t)4 << 30) + ((size_t)1 << 30);
    if (zsk_cuda_free_memory(r->cx) < need || (!r->h_bounce && zsk_cuda_malloc_host(r->cx, (void **)&r->h_bounce, ZSK_BOUNCE)) ||
        zsk_cuda_malloc(r->cx, (void **)&r->g_res, bytes + ZSK_PAD_BACK) ||
        !read_range_device(r, r->g_res, bytes, (size_t)r->d_off[r->shard_lo], call_data, scratch)) {
        zsk_cuda_free(r->cx, r->g_res);
        r->g_res = NULL;
        atomic_fetch_sub(&g_hbm_resident_bytes, r->hbm_resident_bytes);
        r->hbm_resident_bytes = 0;
        return false;
    }
    r->hbm_resident = true;
    zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
    zsk_cuda_free(r->cx, r->g_comp);
    r->g_comp = NULL;
    r->g_comp_cap = 0;
    r->res_lo = r->res_hi = 0;
    zsk_cuda_ctx_trim(r->cx, 0);
    return true;
}

/* back to the ordinary two-half read-ahead window (allocated again on first use) */
static void drop_resident(zseek_reader_t *r)
{
    if (r->hbm_resident) {
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
        zsk_cuda_free(r->cx, r->g_res);
        r->g_res = NULL;
        r->hbm_resident = false;
        atomic_fetch_sub(&g_hbm_resident_bytes, r->hbm_resident_bytes);
        r->hbm_resident_bytes = 0;
        r->resident_tried = false;
        r->random_misses = 0;
    }
    if (!r->resident)
        return;
    atomic_store(&r->resident_fast, false);
    pthread_rwlock_wrlock(&r->res_lock); /* readers on the fast path are through with the window */
    r->resident = false;
    pthread_rwlock_unlock(&r->res_lock);
    zsk_cuda_free_host(r->cx, r->h_mir[0]);
    r->h_mir[0] = NULL;
    r->mir_cap[0] = 0;
    r->mir_lo = r->mir_hi = 0;
    r->mir_cur = 0;
    r->resident = r->resident_tried = false;
    r->random_misses = 0;
    atomic_fetch_sub(&g_resident_bytes, r->resident_bytes);
    r->resident_bytes = 0;
}

/* ------------------------------------------------------------------ zseek_pread (reference src/decompress.c:806-824) */
static ssize_t pread_locked(zseek_reader_t *r, void *buf, size_t count, size_t offset, void *call_data, char *errbuf)
{
    int64_t fi = st_lookup(r, offset);
    if (fi < 0)
        return 0; /* B2: at/after EOF, buf untouched */
    uint64_t f = (uint64_t)fi;
    size_t in_frame = offset - (size_t)r->d_off[f];
    size_t n = MIN(count, (size_t)(r->d_off[f + 1] - r->d_off[f]) - in_frame); /* B1: never crosses a frame */
    if (n == 0)
        return 0; /* B3 */
    if (!in_shard(r, f, errbuf))
        return -1;
    async_fence(r);
    int on_device = buf_on_device(r, buf);
    if (!on_device && f >= r->mir_lo && f < r->mir_hi) { /* pinned window hit: plain memcpy */
        memcpy(buf, mirror_half(r, r->mir_cur) + (offset - r->d_off[r->mir_lo]), n);
        return (ssize_t)n;
    }
    if (r->hbm_resident) { /* the shard sits decoded in HBM: one small copy */
        const uint8_t *from = r->g_res + (offset - (size_t)r->d_off[r->shard_lo]);
        const bool bounce = !on_device && n <= ZSK_BOUNCE;
        if (zsk_cuda_memcpy_async(r->cx, bounce ? (void *)r->h_bounce : buf, from, n, on_device ? ZSK_D2D : ZSK_D2H, ZSK_STREAM_COMPUTE) ||
            zsk_cuda_stream_sync_spin(r->cx, ZSK_STREAM_COMPUTE)) {
            cuda_fail(r, errbuf, "copy frame");
            return -1;
        }
        if (bounce)
            memcpy(buf, r->h_bounce, n);
        return (ssize_t)n;
    }
    if (!on_device && r->pf_active && f == r->pf_lo) {
        /* the scan reached the window that was being decoded behind its back: wait for it, swap halves, and queue
         * the window after it before copying this call's bytes */
        r->pf_active = false;
        const bool dbg = getenv("ZSEEK_B200_DEBUG") != NULL;
        const double t_a = dbg ? now_ms() : 0;
        if (stream_frames_finish(r, r->pf_lo, r->pf_hi, r->pf_resident, r->pf_ok, errbuf)) {
            const double t_b = dbg ? now_ms() : 0;
            r->mir_cur ^= 1;
            r->mir_lo = r->pf_lo;
            r->mir_hi = r->pf_hi;
            r->ra_next = r->pf_hi;
            prefetch_start(r, call_data);
            if (dbg)
                fprintf(stderr, "[zsk %p] read-ahead window [%llu, %llu): waited %.2f ms, queued the next one in %.2f ms\n", (void *)r,
                        (unsigned long long)r->mir_lo, (unsigned long long)r->mir_hi, t_b - t_a, now_ms() - t_b);
            memcpy(buf, mirror_half(r, r->mir_cur) + (offset - r->d_off[r->mir_lo]), n);
            return (ssize_t)n;
        }
        /* fall through: some frame of the read-ahead window is bad; decode frame f alone and report only its verdict */
        r->ra_window = 1;
        r->ra_next = UINT64_MAX;
    }
    prefetch_drop(r);
    int32_t s = cache_find(r, f);
    if (s < 0) {
        /* miss: decode a window of frames in one launch; the window grows while the access pattern
         * stays sequential and collapses to a single frame on a random access */
        if (f == r->ra_next)
            r->ra_window = MIN(r->ra_window * 8, ra_limit(r)); /* a launch costs about the same whatever its size: grow fast */
        else
            r->ra_window = 1;
        if (!on_device && r->ra_window == 1 && !r->resident_tried && r->resident_after && ++r->random_misses >= r->resident_after &&
            r->d_off[r->shard_hi] - r->d_off[r->shard_lo] <= r->resident_max) {
            /* a host reader that keeps missing at random places: decode the whole shard once (every frame is going to be
             * wanted sooner or later, and one launch over thousands of frames is what the GPU is good at) */
            r->resident_tried = true;
            if (go_resident(r, call_data)) {
                memcpy(buf, mirror_half(r, r->mir_cur) + (offset - r->d_off[r->mir_lo]), n);
                return (ssize_t)n;
            }
            if (go_resident_hbm(r, call_data)) /* the pinned budget of the process is spent: keep the decoded shard in HBM */
                return pread_locked(r, buf, count, offset, call_data, errbuf);
        }
        uint64_t hi = MIN(f + (on_device ? MIN(r->ra_window, r->nslots) : r->ra_window), r->shard_hi);
        const bool dbg = getenv("ZSEEK_B200_DEBUG") != NULL;
        const double t_a = dbg ? now_ms() : 0;
        if (!on_device) { /* the window reads are served from is about to be replaced anyway */
            r->mir_lo = r->mir_hi = 0;
            if (!ensure_window(r, r->mir_cur, (size_t)(r->d_off[hi] - r->d_off[f]), errbuf))
                return -1;
        }
        const double t_b = dbg ? now_ms() : 0;
        bool ok;
        if (!on_device && hi - f > 1) {
            /* sequential host reader: decode the window through the H2D / decode / D2H pipeline straight into
             * the pinned mirror (one contiguous copy per chunk); later reads of the window are memcpys */
            r->mir_lo = r->mir_hi = 0;
            ok = stream_frames_to_host(r, f, hi, mirror_half(r, r->mir_cur), call_data, errbuf);
            if (ok) {
                r->mir_lo = f;
                r->mir_hi = hi;
                r->ra_next = hi;
                prefetch_start(r, call_data);
            }
        } else
            ok = fill_window(r, f, hi, !on_device, call_data, errbuf);
        if (dbg)
            fprintf(stderr, "[zsk %p] miss: frames [%llu, %llu) window alloc %.2f ms, decode+copy %.2f ms\n", (void *)r, (unsigned long long)f,
                    (unsigned long long)hi, t_b - t_a, now_ms() - t_b);
        if (!ok) {
            /* the reference decodes only the frame a call asks for (src/decompress.c:700-790): a bad frame further
             * ahead in the window must not fail this read.  Forget the window and decode frame f alone. */
            if (hi - f <= 1)
                return -1;
            r->ra_window = 1;
            hi = f + 1;
            if (!fill_window(r, f, hi, !on_device, call_data, errbuf)) {
                r->ra_next = UINT64_MAX;
                return -1;
            }
            r->ra_next = UINT64_MAX; /* the next call starts over with a window of one frame */
        } else
            r->ra_next = hi;
        if (!on_device) {
            memcpy(buf, mirror_half(r, r->mir_cur) + (offset - r->d_off[r->mir_lo]), n);
            return (ssize_t)n;
        }
        s = r->frame_slot[f];
    }
    if (zsk_cuda_memcpy_async(r->cx, buf, slot_ptr(r, s) + in_frame, n, on_device ? ZSK_D2D : ZSK_D2H, ZSK_STREAM_COMPUTE) ||
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE)) {
        cuda_fail(r, errbuf, "copy frame");
        return -1;
    }
    return (ssize_t)n;
}

ssize_t zseek_pread(zseek_reader_t *reader, void *buf, size_t count, size_t offset, void *call_data,
                    char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!reader) {
        set_error(errbuf, "invalid reader");
        return 0; /* sic: the reference returns `false` here (src/decompress.c:809-812) */
    }
    if (atomic_load_explicit(&reader->resident_fast, memory_order_acquire)) {
        /* resident shard, host buffer: a memcpy under the read lock; everything else takes the ordinary path */
        ssize_t fast = -2;
        pthread_rwlock_rdlock(&reader->res_lock);
        if (reader->resident) {
            const uint64_t block = ((uint64_t)(uintptr_t)buf >> 21) + 1;
            const uint64_t e = atomic_load_explicit(&reader->ptr_cache[block & 7], memory_order_relaxed);
            if ((e >> 1) == block && !(e & 1)) { /* known host memory */
                const int64_t fi = st_lookup(reader, offset);
                if (fi < 0)
                    fast = 0;
                else if ((uint64_t)fi >= reader->mir_lo && (uint64_t)fi < reader->mir_hi) {
                    const size_t in_frame = offset - (size_t)reader->d_off[fi];
                    const size_t n = MIN(count, (size_t)(reader->d_off[fi + 1] - reader->d_off[fi]) - in_frame);
                    memcpy(buf, reader->h_mir[0] + (offset - reader->d_off[reader->mir_lo]), n);
                    fast = (ssize_t)n;
                }
            }
        }
        pthread_rwlock_unlock(&reader->res_lock);
        if (fast != -2)
            return fast;
    }
    pthread_mutex_lock(&reader->lock);
    ssize_t ret = pread_locked(reader, buf, count, offset, call_data, errbuf);
    pthread_mutex_unlock(&reader->lock);
    return ret;
}

ssize_t zseek_read(zseek_reader_t *reader, void *buf, size_t count, void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!reader) {
        set_error(errbuf, "invalid reader");
        return 0;
    }
    pthread_mutex_lock(&reader->lock);
    ssize_t ret = pread_locked(reader, buf, count, reader->pos, call_data, errbuf);
    if (ret > 0)
        reader->pos += (size_t)ret;
    pthread_mutex_unlock(&reader->lock);
    return ret;
}

bool zseek_reader_stats(zseek_reader_t *reader, zseek_reader_stats_t *stats, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!reader) {
        set_error(errbuf, "invalid reader");
        return false;
    }
    if (!stats) {
        set_error(errbuf, "invalid stats pointer");
        return false;
    }
    pthread_mutex_lock(&reader->lock);
    stats->seek_table_memory = 24 + 24 * (size_t)reader->nframes; /* reference src/seek_table.c:228-231 */
    stats->frames = (size_t)reader->nframes;
    stats->decompressed_size = (size_t)reader->d_off[reader->nframes];
    stats->cache_memory = (size_t)reader->cached * reader->slot_size;
    stats->cached_frames = reader->cached;
    stats->buffer_size = reader->g_comp_cap + (reader->h_stage ? 2 * reader->stage_half : 0) + reader->mir_cap[0] + reader->mir_cap[1];
    pthread_mutex_unlock(&reader->lock);
    return true;
}

/* ------------------------------------------------------------------ additive entry points (include/zseek_b200.h) */
bool zseek_b200_set_shard(zseek_reader_t *r, unsigned rank, unsigned world, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!r || world == 0 || rank >= world) {
        set_error(errbuf, "invalid shard");
        return false;
    }
    pthread_mutex_lock(&r->lock);
    async_fence(r);
    prefetch_drop(r);
    drop_resident(r);
    r->shard_lo = r->nframes * rank / world;
    r->shard_hi = r->nframes * (rank + 1) / world;
    pthread_mutex_unlock(&r->lock);
    return true;
}

bool zseek_b200_get_shard(zseek_reader_t *r, size_t *lo, size_t *hi)
{
    if (!r)
        return false;
    if (lo) *lo = (size_t)r->shard_lo;
    if (hi) *hi = (size_t)r->shard_hi;
    return true;
}

bool zseek_b200_seek_table(zseek_reader_t *r, size_t *n, const uint64_t **c_off, const uint64_t **d_off, int *codec)
{
    if (!r)
        return false;
    if (n) *n = (size_t)r->nframes;
    if (c_off) *c_off = r->c_off;
    if (d_off) *d_off = r->d_off;
    if (codec) *codec = r->codec;
    return true;
}

bool zseek_b200_load(zseek_reader_t *r, size_t lo, size_t hi, void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!r || lo > hi || hi > r->nframes) {
        set_error(errbuf, "invalid frame range");
        return false;
    }
    pthread_mutex_lock(&r->lock);
    async_fence(r);
    prefetch_drop(r);
    bool ok = ensure_resident(r, lo, hi, call_data, errbuf) && (zsk_cuda_stream_sync(r->cx, ZSK_STREAM_H2D) == 0);
    pthread_mutex_unlock(&r->lock);
    return ok;
}

/* whole frames [lo, hi) -> device memory, frame f at dst + d_off[f] - d_off[lo]; asynchronous part */
/* h_job_ids[0 .. hi-lo) = frames [lo, hi) ordered by compressed size, largest first (counting sort over up to 4096
 * size classes; stable, so equal classes stay in file order) */
static void order_jobs_by_size(zseek_reader_t *r, uint64_t lo, uint64_t hi)
{
    enum { CLASSES = 4096 };
    static _Thread_local uint32_t start[CLASSES + 1];
    memset(start, 0, sizeof(start));
    unsigned shift = 0; /* the finest power-of-two class width that still maps the largest frame below CLASSES */
    while (((uint64_t)r->max_csize >> shift) >= CLASSES) shift++;
#define SIZE_CLASS(f) ((uint32_t)MIN((uint64_t)(CLASSES - 1), (r->c_off[(f) + 1] - r->c_off[(f)]) >> shift))
    for (uint64_t f = lo; f < hi; f++) start[CLASSES - 1 - SIZE_CLASS(f) + 1]++;
    for (uint32_t c = 0; c < CLASSES; c++) start[c + 1] += start[c];
    for (uint64_t f = lo; f < hi; f++) r->h_job_ids[start[CLASSES - 1 - SIZE_CLASS(f)]++] = (uint32_t)f;
#undef SIZE_CLASS
}

static bool decode_range_device(zseek_reader_t *r, uint64_t lo, uint64_t hi, uint8_t *dst, void *call_data, char *errbuf)
{
    if (lo >= hi)
        return true;
    if (!ensure_resident(r, lo, hi, call_data, errbuf))
        return false;
    if (!ensure_jobs(r, (uint32_t)(hi - lo), errbuf))
        return false;
    zsk_decode_args a;
    fill_decode_args(r, &a);
    a.dst = dst;
    a.dst_base = r->d_off[lo];
    a.first_frame = (uint32_t)lo;
    a.njobs = (uint32_t)(hi - lo);
    a.status = r->g_job_status;
    a.dsize_sum = r->d_off[hi] - r->d_off[lo];
    if (r->sort_min && hi - lo >= (r->codec == ZSK_CODEC_LZ4 ? r->sort_min : r->sort_min_zstd)) {
        if (r->sorted_lo != lo || r->sorted_hi != hi) {
            order_jobs_by_size(r, lo, hi);
            if (zsk_cuda_memcpy_async(r->cx, r->g_job_ids, r->h_job_ids, (size_t)(hi - lo) * sizeof(uint32_t), ZSK_H2D, ZSK_STREAM_COMPUTE))
                return cuda_fail(r, errbuf, "copy job list");
            r->sorted_lo = lo;
            r->sorted_hi = hi;
        }
        a.frame_ids = r->g_job_ids; /* job i = frame g_job_ids[i]; its bytes still land at d_off[frame] - dst_base */
    }
    if (zsk_cuda_launch_decode(r->cx, r->codec, &a, ZSK_STREAM_COMPUTE))
        return cuda_fail(r, errbuf, "decompress frame");
    return true;
}

ssize_t zseek_b200_decode_frames(zseek_reader_t *r, size_t lo, size_t hi, void *dev_dst, void *call_data,
                                 char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!r || lo > hi || hi > r->nframes) {
        set_error(errbuf, "invalid frame range");
        return -1;
    }
    pthread_mutex_lock(&r->lock);
    async_fence(r);
    prefetch_drop(r);
    ssize_t ret = -1;
    if (lo == hi)
        ret = 0;
    else if (in_shard(r, lo, errbuf) && in_shard(r, hi - 1, errbuf) &&
             decode_range_device(r, lo, hi, dev_dst, call_data, errbuf) && finish_decode(r, (uint32_t)(hi - lo), errbuf))
        ret = (ssize_t)(r->d_off[hi] - r->d_off[lo]);
    pthread_mutex_unlock(&r->lock);
    return ret;
}

/* copies the piece [offset, offset + n) of frame f (decoded through the cache) to device memory dst */
static bool partial_to_device(zseek_reader_t *r, uint64_t f, size_t in_frame, size_t n, uint8_t *dst, void *call_data, char *errbuf)
{
    if (cache_find(r, f) < 0 && !fill_window(r, f, f + 1, false, call_data, errbuf))
        return false;
    if (zsk_cuda_memcpy_async(r->cx, dst, slot_ptr(r, r->frame_slot[f]) + in_frame, n, ZSK_D2D, ZSK_STREAM_COMPUTE))
        return cuda_fail(r, errbuf, "copy frame");
    return true;
}

/* [offset, offset+count) -> device memory dst; count already clipped to EOF and > 0 */
static bool read_range_device(zseek_reader_t *r, uint8_t *dst, size_t count, size_t offset, void *call_data, char *errbuf)
{
    uint64_t f0 = (uint64_t)st_lookup(r, offset), f1 = (uint64_t)st_lookup(r, offset + count - 1);
    if (!in_shard(r, f0, errbuf) || !in_shard(r, f1, errbuf))
        return false;
    size_t end = offset + count;
    uint64_t full_lo = f0, full_hi = f1 + 1;
    if (offset != r->d_off[f0]) { /* partial head frame */
        size_t n = MIN(end, (size_t)r->d_off[f0 + 1]) - offset;
        if (!partial_to_device(r, f0, offset - (size_t)r->d_off[f0], n, dst, call_data, errbuf))
            return false;
        full_lo = f0 + 1;
    }
    if (full_hi > full_lo && end != r->d_off[f1 + 1]) { /* partial tail frame (distinct from the head) */
        size_t start = (size_t)r->d_off[f1];
        if (!partial_to_device(r, f1, 0, end - start, dst + (start - offset), call_data, errbuf))
            return false;
        full_hi = f1;
    }
    if (full_hi > full_lo) {
        /* bound the resident compressed window when the range is larger than what is resident */
        uint64_t lo = full_lo;
        while (lo < full_hi) {
            uint64_t hi = full_hi;
            if (!(lo >= r->res_lo && hi <= r->res_hi)) {
                size_t budget = (size_t)4 << 30;
                hi = lo + 1;
                while (hi < full_hi && r->c_off[hi + 1] - r->c_off[lo] <= budget)
                    hi++;
            }
            if (!decode_range_device(r, lo, hi, dst + (r->d_off[lo] - offset), call_data, errbuf) ||
                !finish_decode(r, (uint32_t)(hi - lo), errbuf))
                return false;
            lo = hi;
        }
    }
    if (zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE))
        return cuda_fail(r, errbuf, "synchronize");
    return true;
}

static bool ensure_out(zseek_reader_t *r, size_t n, char *errbuf)
{
    if (n <= r->g_out_cap)
        return true;
    zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
    zsk_cuda_stream_sync(r->cx, ZSK_STREAM_D2H);
    const size_t old_cap = r->g_out_cap;
    zsk_cuda_free(r->cx, r->g_out);
    r->g_out = NULL;
    r->g_out_cap = 0;
    size_t want = MAX(n, 2 * old_cap);
    if (zsk_cuda_malloc(r->cx, (void **)&r->g_out, want + ZSK_PAD_BACK)) {
        want = n;
        if (zsk_cuda_malloc(r->cx, (void **)&r->g_out, want + ZSK_PAD_BACK))
            return cuda_fail(r, errbuf, "allocate output staging");
    }
    r->g_out_cap = want;
    return true;
}

/* events of the host-destination pipeline */
enum { EV_H2D0 = 0, EV_H2D1 = 1, EV_DEC0 = 2, EV_DEC1 = 3, EV_D2H0 = 4, EV_D2H1 = 5 };

/*
 * Whole frames [lo, hi) -> HOST memory `dst` as a three-stage pipeline over chunks of ~chunk_bytes:
 *     H2D stream     : compressed bytes of chunk k+1 (pinned image / staged callback reads)
 *     compute stream : decode kernel of chunk k into device staging half k&1
 *     D2H stream     : decoded bytes of chunk k-1 to the caller's buffer
 * The host only queues work; chunks are ordered by events, so PCIe runs in both directions while
 * the kernel of the next chunk executes.
 */
/* Queues the whole pipeline; returns without waiting for the device.  *was_resident tells stream_frames_finish
 * whether the compressed image was already in HBM. */
static bool stream_frames_begin(zseek_reader_t *r, uint64_t lo, uint64_t hi, uint8_t *dst, void *call_data, char *errbuf,
                                bool *was_resident)
{
    const uint64_t nfr = hi - lo;
    const bool resident = lo >= r->res_lo && hi <= r->res_hi;
    size_t chunk = MAX(r->chunk_bytes, (size_t)r->max_dsize);
    const size_t span = (size_t)(r->d_off[hi] - r->d_off[lo]);
    if (span < chunk) /* small windows do not need the full-size staging halves */
        chunk = MAX(span, (size_t)r->max_dsize);
    if (!ensure_jobs(r, (uint32_t)nfr, errbuf) || !ensure_out(r, 2 * chunk, errbuf))
        return false;
    if (!resident && !alloc_image(r, lo, hi, errbuf))
        return false;
    const uint64_t img_lo = resident ? r->res_lo : lo;
    uint8_t *img = r->g_comp + ZSK_PAD_FRONT;
    bool ok = true;
    uint64_t a = lo, next_b = lo;
    unsigned k = 0;
    /* chunk k = frames [a, b): as many frames as fit `chunk` decoded bytes */
    /* the first chunks are small and double up to `chunk`: the first D2H starts after ~1 ms instead of after the
     * H2D + decode of a full-size chunk, which nothing overlaps */
    size_t lim = MIN(chunk, MAX(r->ramp_bytes, (size_t)r->max_dsize));
#define CHUNK_END(from, to)                                                                             \
    do {                                                                                                \
        (to) = (from) + 1;                                                                              \
        while ((to) < hi && r->d_off[(to) + 1] - r->d_off[(from)] <= lim) (to)++;                       \
        lim = MIN(chunk, lim * 2);                                                                      \
    } while (0)
    uint64_t b;
    CHUNK_END(a, b);
    zsk_cuda_trace_reset(r->cx);
    zsk_cuda_trace_mark(r->cx, ZSK_STREAM_H2D, "start", 0);
    if (!resident) {
        ok = h2d_range(r, (size_t)r->c_off[a], (size_t)(r->c_off[b] - r->c_off[a]), img + (r->c_off[a] - r->c_off[img_lo]), call_data, errbuf) &&
             (zsk_cuda_event_record(r->cx, EV_H2D0, ZSK_STREAM_H2D) == 0 || cuda_fail(r, errbuf, "order streams"));
    }
    while (ok && a < hi) {
        const int half = (int)(k & 1);
        if (b < hi && !resident) { /* prefetch the compressed bytes of chunk k+1 */
            CHUNK_END(b, next_b);
            ok = h2d_range(r, (size_t)r->c_off[b], (size_t)(r->c_off[next_b] - r->c_off[b]), img + (r->c_off[b] - r->c_off[img_lo]), call_data, errbuf) &&
                 (zsk_cuda_event_record(r->cx, EV_H2D0 + (half ^ 1), ZSK_STREAM_H2D) == 0 || cuda_fail(r, errbuf, "order streams"));
            if (!ok) break;
        } else if (b < hi) {
            CHUNK_END(b, next_b);
        }
        if (!resident && zsk_cuda_stream_wait_event(r->cx, ZSK_STREAM_COMPUTE, EV_H2D0 + half)) { ok = cuda_fail(r, errbuf, "order streams"); break; }
        if (k >= 2 && zsk_cuda_stream_wait_event(r->cx, ZSK_STREAM_COMPUTE, EV_D2H0 + half)) { ok = cuda_fail(r, errbuf, "order streams"); break; }
        uint8_t *stage = r->g_out + (size_t)half * chunk;
        zsk_decode_args da;
        memset(&da, 0, sizeof(da));
        da.c_off = r->g_coff;
        da.d_off = r->g_doff;
        da.comp = img;
        da.comp_base = r->c_off[img_lo];
        da.dst = stage;
        da.dst_base = r->d_off[a];
        da.first_frame = (uint32_t)a;
        da.njobs = (uint32_t)(b - a);
        da.status = r->g_job_status + (a - lo);
        const size_t nbytes = (size_t)(r->d_off[b] - r->d_off[a]);
        da.dsize_sum = nbytes;
        zsk_cuda_trace_mark(r->cx, ZSK_STREAM_H2D, "h2d queued up to chunk", k + 1);
        zsk_cuda_trace_mark(r->cx, ZSK_STREAM_COMPUTE, "decode begin", k);
        if (zsk_cuda_launch_decode(r->cx, r->codec, &da, ZSK_STREAM_COMPUTE) ||
            zsk_cuda_event_record(r->cx, EV_DEC0 + half, ZSK_STREAM_COMPUTE) ||
            zsk_cuda_stream_wait_event(r->cx, ZSK_STREAM_D2H, EV_DEC0 + half)) {
            ok = cuda_fail(r, errbuf, "decompress frame");
            break;
        }
        zsk_cuda_trace_mark(r->cx, ZSK_STREAM_COMPUTE, "decode end", k);
        zsk_cuda_trace_mark(r->cx, ZSK_STREAM_D2H, "d2h begin", k);
        if (zsk_cuda_memcpy_async(r->cx, dst + (r->d_off[a] - r->d_off[lo]), stage, nbytes, ZSK_D2H, ZSK_STREAM_D2H) ||
            zsk_cuda_event_record(r->cx, EV_D2H0 + half, ZSK_STREAM_D2H)) {
            ok = cuda_fail(r, errbuf, "decompress frame");
            break;
        }
        zsk_cuda_trace_mark(r->cx, ZSK_STREAM_D2H, "d2h end", k);
        a = b;
        b = next_b;
        k++;
    }
#undef CHUNK_END
    *was_resident = resident;
    return ok;
}

/* Waits for a pipeline queued by stream_frames_begin (ok = what begin returned) and checks the frame statuses. */
static bool stream_frames_finish(zseek_reader_t *r, uint64_t lo, uint64_t hi, bool resident, bool ok, char *errbuf)
{
    if (ok)
        ok = finish_decode(r, (uint32_t)(hi - lo), errbuf);
    else
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
    if (zsk_cuda_stream_sync(r->cx, ZSK_STREAM_D2H) && ok)
        ok = cuda_fail(r, errbuf, "copy to host");
    zsk_cuda_trace_dump(r->cx);
    if (!r->mem_image) {
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_H2D);
        r->stage_inflight = 0;
    }
    if (ok && !resident) {
        r->res_lo = lo;
        r->res_hi = hi;
    }
    return ok;
}

static bool stream_frames_to_host(zseek_reader_t *r, uint64_t lo, uint64_t hi, uint8_t *dst, void *call_data, char *errbuf)
{
    bool resident = false;
    bool ok = stream_frames_begin(r, lo, hi, dst, call_data, errbuf, &resident);
    return stream_frames_finish(r, lo, hi, resident, ok, errbuf);
}

/* ---- asynchronous read-ahead of the plain zseek_pread path */
static uint8_t *mirror_half(zseek_reader_t *r, int which) { return r->h_mir[which]; }

/* Waits for the queued read-ahead (if any) and forgets it: every path that uses the streams, the staging buffers or the
 * job buffers calls this first. */
static void prefetch_drop(zseek_reader_t *r)
{
    if (!r->pf_active)
        return;
    char scratch[ZSEEK_ERRBUF_SIZE];
    stream_frames_finish(r, r->pf_lo, r->pf_hi, r->pf_resident, r->pf_ok, scratch);
    r->pf_active = false;
}

/* Queues the decode of the window after the current one into the other mirror half. */
static void prefetch_start(zseek_reader_t *r, void *call_data)
{
    if (r->pf_active || r->mir_hi <= r->mir_lo || r->mir_hi >= r->shard_hi || r->mir_hi - r->mir_lo < 4)
        return;
    r->ra_window = MIN(r->ra_window * 8, ra_limit(r));
    uint64_t lo = r->mir_hi, hi = MIN(lo + r->ra_window, r->shard_hi);
    char scratch[ZSEEK_ERRBUF_SIZE]; /* a failing read-ahead is not an error of this call: the window is retried synchronously */
    if (!ensure_window(r, r->mir_cur ^ 1, (size_t)(r->d_off[hi] - r->d_off[lo]), scratch))
        return;
    r->pf_ok = stream_frames_begin(r, lo, hi, mirror_half(r, r->mir_cur ^ 1), call_data, scratch, &r->pf_resident);
    r->pf_lo = lo;
    r->pf_hi = hi;
    r->pf_active = true;
}

/* one partial frame piece -> host memory (through the decoded-frame cache) */
static bool partial_to_host(zseek_reader_t *r, uint64_t f, size_t in_frame, size_t n, uint8_t *dst, void *call_data, char *errbuf)
{
    if (cache_find(r, f) < 0 && !fill_window(r, f, f + 1, false, call_data, errbuf))
        return false;
    if (zsk_cuda_memcpy_async(r->cx, dst, slot_ptr(r, r->frame_slot[f]) + in_frame, n, ZSK_D2H, ZSK_STREAM_COMPUTE) ||
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE))
        return cuda_fail(r, errbuf, "copy frame");
    return true;
}

/* [offset, offset+count) -> host memory; count already clipped to EOF and > 0 */
static bool read_range_host(zseek_reader_t *r, uint8_t *dst, size_t count, size_t offset, void *call_data, char *errbuf)
{
    uint64_t f0 = (uint64_t)st_lookup(r, offset), f1 = (uint64_t)st_lookup(r, offset + count - 1);
    if (!in_shard(r, f0, errbuf) || !in_shard(r, f1, errbuf))
        return false;
    size_t end = offset + count;
    uint64_t full_lo = f0, full_hi = f1 + 1;
    if (offset != r->d_off[f0]) {
        size_t n = MIN(end, (size_t)r->d_off[f0 + 1]) - offset;
        if (!partial_to_host(r, f0, offset - (size_t)r->d_off[f0], n, dst, call_data, errbuf))
            return false;
        full_lo = f0 + 1;
    }
    if (full_hi > full_lo && end != r->d_off[f1 + 1]) {
        size_t start = (size_t)r->d_off[f1];
        if (!partial_to_host(r, f1, 0, end - start, dst + (start - offset), call_data, errbuf))
            return false;
        full_hi = f1;
    }
    /* a window of compressed bytes larger than this is streamed piecewise instead of being made resident at once */
    const size_t budget = (size_t)16 << 30;
    while (full_lo < full_hi) {
        uint64_t hi = full_hi;
        if (!(full_lo >= r->res_lo && hi <= r->res_hi) && r->c_off[hi] - r->c_off[full_lo] > budget) {
            hi = full_lo + 1;
            while (hi < full_hi && r->c_off[hi + 1] - r->c_off[full_lo] <= budget)
                hi++;
        }
        if (!stream_frames_to_host(r, full_lo, hi, dst + (r->d_off[full_lo] - offset), call_data, errbuf))
            return false;
        full_lo = hi;
    }
    return true;
}

ssize_t zseek_b200_read_range(zseek_reader_t *r, void *buf, size_t count, size_t offset, void *call_data,
                              char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!r) {
        set_error(errbuf, "invalid reader");
        return -1;
    }
    pthread_mutex_lock(&r->lock);
    async_fence(r);
    prefetch_drop(r);
    ssize_t ret = -1;
    size_t total = (size_t)r->d_off[r->nframes];
    if (offset >= total || count == 0) {
        ret = 0;
        goto out;
    }
    count = MIN(count, total - offset);
    if (buf_on_device(r, buf)) {
        if (read_range_device(r, buf, count, offset, call_data, errbuf))
            ret = (ssize_t)count;
        goto out;
    }
    if (read_range_host(r, buf, count, offset, call_data, errbuf))
        ret = (ssize_t)count;
out:
    pthread_mutex_unlock(&r->lock);
    return ret;
}

static bool ensure_batch(zseek_reader_t *r, size_t n, char *errbuf)
{
    uint64_t N = r->nframes;
    if (!r->g_touched) {
        if (zsk_cuda_malloc(r->cx, (void **)&r->g_touched, (N + 1) * sizeof(uint32_t)))
            return cuda_fail(r, errbuf, "allocate batch buffers");
    }
    if (n <= r->batch_cap)
        return true;
    zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
    void *old[] = { r->g_b_offsets, r->g_b_counts, r->g_b_dstoffs, r->g_b_frame, r->g_b_inframe, r->g_b_nbytes };
    for (size_t i = 0; i < 6; i++)
        zsk_cuda_free(r->cx, old[i]);
    r->g_b_offsets = r->g_b_counts = r->g_b_dstoffs = NULL;
    r->g_b_frame = NULL;
    r->g_b_inframe = r->g_b_nbytes = NULL;
    r->batch_cap = 0;
    size_t cap = MAX(n, 4096);
    if (zsk_cuda_malloc(r->cx, (void **)&r->g_b_offsets, cap * 8) || zsk_cuda_malloc(r->cx, (void **)&r->g_b_counts, cap * 8) ||
        zsk_cuda_malloc(r->cx, (void **)&r->g_b_dstoffs, cap * 8) || zsk_cuda_malloc(r->cx, (void **)&r->g_b_frame, cap * 4) ||
        zsk_cuda_malloc(r->cx, (void **)&r->g_b_inframe, cap * 4) || zsk_cuda_malloc(r->cx, (void **)&r->g_b_nbytes, cap * 4))
        return cuda_fail(r, errbuf, "allocate batch buffers");
    r->batch_cap = cap;
    return true;
}

static bool push_frame_src(zseek_reader_t *r, char *errbuf)
{
    if (!r->frame_src_dirty)
        return true;
    if (zsk_cuda_memcpy_async(r->cx, r->g_frame_src, r->h_frame_src, r->nframes * sizeof(int64_t), ZSK_H2D, ZSK_STREAM_COMPUTE) ||
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE)) /* h_frame_src is pageable and edited right after */
        return cuda_fail(r, errbuf, "upload cache map");
    r->frame_src_dirty = false;
    return true;
}

ssize_t zseek_b200_pread_batch(zseek_reader_t *r, size_t n, const uint64_t *offsets, const uint64_t *counts,
                               uint64_t fixed_count, void *dst, const uint64_t *dst_offs, uint64_t dst_stride,
                               int64_t *results, void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!r) {
        set_error(errbuf, "invalid reader");
        return -1;
    }
    if (n == 0)
        return 0;
    if (n > 0x7fffffffu) {
        set_error(errbuf, "batch too large");
        return -1;
    }
    pthread_mutex_lock(&r->lock);
    async_fence(r);
    prefetch_drop(r);
    ssize_t ret = -1;
    uint64_t N = r->nframes;
    int on_device = buf_on_device(r, dst);
    uint8_t *gdst = dst;
    size_t extent = 0;
    uint32_t *nb = NULL; /* clipped byte count of every request (K1), host copy */
    if (!ensure_batch(r, n, errbuf))
        goto out;
    /* K1: lookup on the device */
    if (zsk_cuda_memcpy_async(r->cx, r->g_b_offsets, offsets, n * 8, ZSK_H2D, ZSK_STREAM_COMPUTE) ||
        (counts && zsk_cuda_memcpy_async(r->cx, r->g_b_counts, counts, n * 8, ZSK_H2D, ZSK_STREAM_COMPUTE)) ||
        (dst_offs && zsk_cuda_memcpy_async(r->cx, r->g_b_dstoffs, dst_offs, n * 8, ZSK_H2D, ZSK_STREAM_COMPUTE)) ||
        zsk_cuda_memset_async(r->cx, r->g_touched, 0, (N + 1) * sizeof(uint32_t), ZSK_STREAM_COMPUTE)) {
        cuda_fail(r, errbuf, "upload batch");
        goto out;
    }
    zsk_lookup_args la = { r->g_doff, (uint32_t)N, r->g_b_offsets, counts ? r->g_b_counts : NULL, fixed_count, (uint32_t)n,
                           r->g_b_frame, r->g_b_inframe, r->g_b_nbytes, r->g_touched };
    const bool want_nb = results || !on_device;
    if (want_nb && !(nb = malloc(n * sizeof(uint32_t)))) {
        set_error(errbuf, "allocate results");
        goto out;
    }
    if (zsk_cuda_launch_lookup(r->cx, &la, ZSK_STREAM_COMPUTE) ||
        zsk_cuda_memcpy_async(r->cx, r->h_touched, r->g_touched, N * sizeof(uint32_t), ZSK_D2H, ZSK_STREAM_COMPUTE) ||
        (want_nb && zsk_cuda_memcpy_async(r->cx, nb, r->g_b_nbytes, n * sizeof(uint32_t), ZSK_D2H, ZSK_STREAM_COMPUTE)) ||
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE)) {
        cuda_fail(r, errbuf, "lookup");
        goto out;
    }
    if (!on_device) {
        /* host destination: gather into device staging laid out like the caller's buffer, then copy back ONLY the
         * bytes the requests produced (the reference leaves everything else untouched: stride gaps, the tail of a
         * read that ends at a frame boundary, requests at or after EOF) */
        for (size_t i = 0; i < n; i++) {
            if (!nb[i])
                continue;
            const uint64_t o = dst_offs ? dst_offs[i] : (uint64_t)i * dst_stride;
            if (o + nb[i] < o || o + nb[i] > (uint64_t)SIZE_MAX / 2) {
                set_error(errbuf, "batch destination out of range");
                goto out;
            }
            extent = MAX(extent, (size_t)(o + nb[i]));
        }
        if (!ensure_out(r, extent, errbuf))
            goto out;
        gdst = r->g_out;
    }
    /* touched frames, ascending; processed in groups that fit the decoded-frame cache */
    zsk_gather_args ga = { r->g_b_frame, r->g_b_inframe, r->g_b_nbytes, r->g_frame_src, r->g_slab + ZSK_PAD_FRONT, gdst,
                           dst_offs ? r->g_b_dstoffs : NULL, dst_stride, (uint32_t)n, NULL };
    uint64_t f = 0;
    while (f < N) {
        uint32_t in_group = 0, nmiss = 0;
        uint64_t g_lo = N, g_hi = 0;
        if (!ensure_jobs(r, r->nslots, errbuf))
            goto out;
        for (; f < N && in_group < r->nslots; f++) {
            if (!r->h_touched[f])
                continue;
            if (f < r->shard_lo || f >= r->shard_hi) {
                set_error(errbuf, "frame outside this reader's shard");
                goto out;
            }
            in_group++;
            if (cache_find_prefix(r, f, r->h_touched[f]) < 0) { /* hits are promoted to MRU so the takes below cannot evict them */
                r->h_job_ids[nmiss++] = (uint32_t)f;
                g_lo = MIN(g_lo, f);
                g_hi = MAX(g_hi, f + 1);
            }
        }
        if (in_group == 0)
            break;
        if (nmiss) {
            if (!ensure_resident(r, g_lo, g_hi, call_data, errbuf) || !decode_into_cache(r, nmiss, r->partial_decode ? r->h_touched : NULL, errbuf))
                goto out;
        }
        if (!push_frame_src(r, errbuf))
            goto out;
        /* K4: serves every request whose frame is resident now; requests of other groups are either
         * skipped (frame absent) or harmlessly served early/again with identical bytes */
        if (zsk_cuda_launch_gather(r->cx, &ga, ZSK_STREAM_COMPUTE) || zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE)) {
            cuda_fail(r, errbuf, "gather");
            goto out;
        }
    }
    if (results)
        for (size_t i = 0; i < n; i++)
            results[i] = nb[i];
    if (!on_device && extent) {
        /* produced ranges, merged while they touch: full reads at a regular stride leave as one copy */
        size_t run_lo = 0, run_hi = 0;
        bool ok = true;
        for (size_t i = 0; i <= n && ok; i++) {
            const size_t o = i < n ? (size_t)(dst_offs ? dst_offs[i] : (uint64_t)i * dst_stride) : 0;
            if (i < n && !nb[i])
                continue;
            if (i < n && run_hi > run_lo && o == run_hi) {
                run_hi += nb[i];
                continue;
            }
            if (run_hi > run_lo)
                ok = !zsk_cuda_memcpy_async(r->cx, (uint8_t *)dst + run_lo, gdst + run_lo, run_hi - run_lo, ZSK_D2H, ZSK_STREAM_COMPUTE);
            if (i < n) {
                run_lo = o;
                run_hi = o + nb[i];
            }
        }
        if (!ok || zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE)) {
            cuda_fail(r, errbuf, "copy to host");
            goto out;
        }
    }
    ret = (ssize_t)n;
out:
    free(nb);
    pthread_mutex_unlock(&r->lock);
    return ret;
}

/* ------------------------------------------------------------------ stream-ordered batch */
static int batch_wait_locked(zseek_reader_t *r, char *errbuf)
{
    if (r->async_stream < 0)
        return 0;
    const int sidx = r->async_stream;
    r->async_stream = -1;
    uint32_t ctl[2] = { 0, 0 };
    if (zsk_cuda_memcpy_async(r->cx, ctl, r->g_bctl, sizeof(ctl), ZSK_D2H, sidx) || zsk_cuda_stream_sync(r->cx, sidx)) {
        cuda_fail(r, errbuf, "batch");
        return -1;
    }
    if (ctl[1]) {
        set_error(errbuf, ctl[1] == 1 ? "frame outside this reader's shard" : "batch touches more frames than the slab holds");
        return -1;
    }
    const uint32_t njobs = MIN(ctl[0], r->async_jobs_max);
    if (njobs == 0)
        return 0;
    int32_t *st = malloc(njobs * sizeof(int32_t));
    if (!st) {
        set_error(errbuf, "allocate statuses");
        return -1;
    }
    int rc = 0;
    if (zsk_cuda_memcpy_async(r->cx, st, r->g_bjob_status, njobs * sizeof(int32_t), ZSK_D2H, sidx) || zsk_cuda_stream_sync(r->cx, sidx)) {
        cuda_fail(r, errbuf, "batch");
        rc = -1;
    } else
        for (uint32_t i = 0; i < njobs; i++)
            if (st[i] != ZSK_ST_OK) {
                set_error(errbuf, "decompress frame: %s", status_name(st[i]));
                rc = -1;
                break;
            }
    free(st);
    return rc;
}

int zseek_b200_batch_wait(zseek_reader_t *r, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!r) {
        set_error(errbuf, "invalid reader");
        return -1;
    }
    pthread_mutex_lock(&r->lock);
    int rc = batch_wait_locked(r, errbuf);
    pthread_mutex_unlock(&r->lock);
    return rc;
}

ssize_t zseek_b200_pread_batch_async(zseek_reader_t *r, size_t n, const uint64_t *dev_offsets, const uint64_t *dev_counts,
                                     uint64_t fixed_count, void *dev_dst, const uint64_t *dev_dst_offs, uint64_t dst_stride,
                                     int64_t *dev_results, void *stream, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!r) {
        set_error(errbuf, "invalid reader");
        return -1;
    }
    if (n == 0)
        return 0;
    if (n > 0x7fffffffu) {
        set_error(errbuf, "batch too large");
        return -1;
    }
    pthread_mutex_lock(&r->lock);
    ssize_t ret = -1;
    char scratch[ZSEEK_ERRBUF_SIZE];
    prefetch_drop(r);
    batch_wait_locked(r, scratch); /* one async batch in flight: its buffers are about to be reused */
    const uint64_t N = r->nframes, nshard = r->shard_hi - r->shard_lo;
    if (!(r->shard_lo >= r->res_lo && r->shard_hi <= r->res_hi) && nshard) {
        set_error(errbuf, "compressed image not resident: call zseek_b200_load first");
        goto out;
    }
    if (!ensure_batch(r, n, errbuf))
        goto out;
    const uint32_t need = (uint32_t)MIN((uint64_t)n, nshard);
    if (!r->g_bsrc &&
        (zsk_cuda_malloc(r->cx, (void **)&r->g_bsrc, (N + 1) * sizeof(int64_t)) || zsk_cuda_malloc(r->cx, (void **)&r->g_bctl, 2 * sizeof(uint32_t)))) {
        cuda_fail(r, errbuf, "allocate batch buffers");
        goto out;
    }
    if (need > r->bs_cap) {
        zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
        void *old[] = { r->g_bslab, r->g_bjob_ids, r->g_bjob_limits, r->g_bjob_offs, r->g_bjob_status };
        for (size_t i = 0; i < 5; i++)
            zsk_cuda_free(r->cx, old[i]);
        r->g_bslab = NULL; r->g_bjob_ids = r->g_bjob_limits = NULL; r->g_bjob_offs = NULL; r->g_bjob_status = NULL;
        r->bs_cap = 0;
        const uint32_t cap = (uint32_t)MIN(nshard, MAX((uint64_t)need, (uint64_t)1024));
        if (zsk_cuda_malloc(r->cx, (void **)&r->g_bslab, (size_t)cap * r->slot_size + ZSK_PAD_FRONT + ZSK_PAD_BACK) ||
            zsk_cuda_malloc(r->cx, (void **)&r->g_bjob_ids, cap * sizeof(uint32_t)) ||
            zsk_cuda_malloc(r->cx, (void **)&r->g_bjob_limits, cap * sizeof(uint32_t)) ||
            zsk_cuda_malloc(r->cx, (void **)&r->g_bjob_offs, cap * sizeof(uint64_t)) ||
            zsk_cuda_malloc(r->cx, (void **)&r->g_bjob_status, cap * sizeof(int32_t))) {
            cuda_fail(r, errbuf, "allocate batch slab");
            goto out;
        }
        r->bs_cap = cap;
    }
    const int sidx = stream ? ZSK_STREAM_USER : ZSK_STREAM_COMPUTE;
    if (stream) {
        zsk_cuda_set_user_stream(r->cx, stream);
        /* the caller's stream has never seen what the reader's own streams did (image upload, earlier launches that
         * use the same scratch): order it after them */
        if (zsk_cuda_stream_wait(r->cx, ZSK_STREAM_USER, ZSK_STREAM_H2D) || zsk_cuda_stream_wait(r->cx, ZSK_STREAM_USER, ZSK_STREAM_COMPUTE)) {
            cuda_fail(r, errbuf, "order streams");
            goto out;
        }
    }
    zsk_lookup_args la = { r->g_doff, (uint32_t)N, dev_offsets, dev_counts, fixed_count, (uint32_t)n,
                           r->g_b_frame, r->g_b_inframe, r->g_b_nbytes, r->g_touched };
    zsk_compact_args ca = { r->g_touched, (uint32_t)N, (uint32_t)r->shard_lo, (uint32_t)r->shard_hi, r->slot_size, r->bs_cap,
                            r->g_bjob_ids, r->g_bjob_offs, r->g_bjob_limits, r->g_bsrc, r->g_bctl, r->g_bctl + 1 };
    zsk_decode_args da;
    fill_decode_args(r, &da);
    da.frame_ids = r->g_bjob_ids;
    da.dst_offs = r->g_bjob_offs;
    da.dst = r->g_bslab + ZSK_PAD_FRONT;
    da.njobs = need;
    da.njobs_dev = r->g_bctl;
    da.status = r->g_bjob_status;
    da.limits = r->partial_decode ? r->g_bjob_limits : NULL;
    da.dsize_sum = (uint64_t)need * r->max_dsize;
    zsk_gather_args ga = { r->g_b_frame, r->g_b_inframe, r->g_b_nbytes, r->g_bsrc, r->g_bslab + ZSK_PAD_FRONT, dev_dst,
                           dev_dst_offs, dst_stride, (uint32_t)n, dev_results };
    if (zsk_cuda_memset_async(r->cx, r->g_touched, 0, (N + 1) * sizeof(uint32_t), sidx) ||
        zsk_cuda_memset_async(r->cx, r->g_bctl, 0, 2 * sizeof(uint32_t), sidx) ||
        zsk_cuda_launch_lookup(r->cx, &la, sidx) || zsk_cuda_launch_compact(r->cx, &ca, sidx) ||
        (need && zsk_cuda_launch_decode(r->cx, r->codec, &da, sidx)) || zsk_cuda_launch_gather(r->cx, &ga, sidx)) {
        cuda_fail(r, errbuf, "queue batch");
        zsk_cuda_stream_sync(r->cx, sidx);
        goto out;
    }
    r->async_stream = sidx;
    r->async_jobs_max = need;
    ret = (ssize_t)n;
out:
    pthread_mutex_unlock(&r->lock);
    return ret;
}

void zseek_b200_cache_clear(zseek_reader_t *r)
{
    if (!r)
        return;
    pthread_mutex_lock(&r->lock);
    async_fence(r);
    prefetch_drop(r);
    drop_resident(r);
    cache_clear(r);
    pthread_mutex_unlock(&r->lock);
}

void zseek_b200_unload(zseek_reader_t *r)
{
    if (!r)
        return;
    pthread_mutex_lock(&r->lock);
    async_fence(r);
    prefetch_drop(r);
    zsk_cuda_stream_sync(r->cx, ZSK_STREAM_COMPUTE);
    r->res_lo = r->res_hi = 0;
    pthread_mutex_unlock(&r->lock);
}

bool zseek_b200_timer_start(zseek_reader_t *r) { return r && zsk_cuda_timer_start(r->cx, ZSK_STREAM_COMPUTE) == 0; }

double zseek_b200_timer_stop(zseek_reader_t *r)
{
    float ms = -1.0f;
    if (!r || zsk_cuda_timer_stop(r->cx, ZSK_STREAM_COMPUTE, &ms))
        return -1.0;
    return ms;
}

unsigned long long zseek_b200_launch_count(zseek_reader_t *r) { return r ? zsk_cuda_launch_count(r->cx) : 0; }

const char *zseek_b200_last_decode_kernel(zseek_reader_t *r) { return r ? zsk_cuda_last_decode_kernel(r->cx) : ""; }

double zseek_b200_last_decode_ms(zseek_reader_t *r)
{
    float ms = -1.0f;
    if (!r || zsk_cuda_last_decode_ms(r->cx, &ms))
        return -1.0;
    return ms;
}

int zseek_b200_device(zseek_reader_t *r) { return r ? zsk_cuda_device(r->cx) : -1; }
