/*
 * refdrive.c — TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by, or called from the product
 * library (libzseek_b200/); only tests/, __graft_entry__.smoke() and bench.py (input production,
 * the cpu_baseline leg and --impl reference) load it.
 *
 * A thin driver around the UNMODIFIED reference build oracle/_ref/libzseek_ref.so (compiled by
 * oracle/Makefile from the .c files under /root/reference/src).  It dlopen()s that library with RTLD_LOCAL so the
 * reference's zseek_* symbols never collide with the product's, and offers:
 *
 *   - refdrive_compress      : run the reference WRITER (zseek_writer_open_full / zseek_write /
 *                              zseek_writer_close, reference src/compress.c:247,815,578) over a
 *                              memory buffer with a constant chunk size, collecting the file image
 *                              in memory.  This is how every input file is produced (north_star:
 *                              "the write path stays the reference CPU writer").
 *   - refdrive_reader_*      : open the reference READER over a memory image (memcpy pread
 *                              callback, like the reference benchmark's "load whole file to memory",
 *                              reference README.md:42-43) and forward zseek_pread/stats/close.
 *   - refdrive_scan / _random: the CPU-baseline timing harness of BASELINE.md §2 — T pinned threads,
 *                              ONE READER PER THREAD (a shared reader serialises on its write lock,
 *                              reference src/decompress.c:387,499), cache_size 0, CLOCK_MONOTONIC
 *                              from first thread start to last join.
 *
 * The writer-side parameter structs below restate the public layout of reference src/zseek.h:121-159.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <sched.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../include/zseek.h"

/* ---- writer-side public types of the reference (layout restated, src/zseek.h:52-159) ---- */
typedef bool (*ref_write_t)(const void *data, size_t size, void *user_data, void *call_data);
typedef struct { void *user_data; ref_write_t write; } ref_write_file_t;
typedef struct {
    int nb_workers; size_t cpusetsize; cpu_set_t *cpuset; int compression_level; int strategy;
} ref_zstd_param_t;
typedef struct { int compression_level; } ref_lz4_param_t;
typedef struct {
    int type; /* 0 = zstd, 1 = lz4 */
    union { ref_zstd_param_t zstd_params; ref_lz4_param_t lz4_params; } params;
} ref_compression_param_t;

typedef void *(*fn_writer_open_full)(ref_write_file_t, ref_compression_param_t *, size_t, void *, char *);
typedef bool (*fn_write)(void *, const void *, size_t, void *, char *);
typedef bool (*fn_writer_close)(void *, void *, char *);
typedef void *(*fn_reader_open_full)(zseek_read_file_t, size_t, void *, char *);
typedef ssize_t (*fn_pread)(void *, void *, size_t, size_t, void *, char *);
typedef ssize_t (*fn_read)(void *, void *, size_t, void *, char *);
typedef bool (*fn_reader_stats)(void *, zseek_reader_stats_t *, char *);
typedef bool (*fn_reader_close)(void *, void *, char *);

static struct {
    void *dl;
    fn_writer_open_full writer_open_full;
    fn_write write;
    fn_writer_close writer_close;
    fn_reader_open_full reader_open_full;
    fn_pread pread;
    fn_read read;
    fn_reader_stats reader_stats;
    fn_reader_close reader_close;
} R;

#define EXPORT __attribute__((visibility("default")))

EXPORT int refdrive_init(const char *so_path)
{
    if (R.dl)
        return 0;
    void *dl = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) {
        fprintf(stderr, "refdrive: dlopen(%s): %s\n", so_path, dlerror());
        return -1;
    }
    R.writer_open_full = (fn_writer_open_full)dlsym(dl, "zseek_writer_open_full");
    R.write = (fn_write)dlsym(dl, "zseek_write");
    R.writer_close = (fn_writer_close)dlsym(dl, "zseek_writer_close");
    R.reader_open_full = (fn_reader_open_full)dlsym(dl, "zseek_reader_open_full");
    R.pread = (fn_pread)dlsym(dl, "zseek_pread");
    R.read = (fn_read)dlsym(dl, "zseek_read");
    R.reader_stats = (fn_reader_stats)dlsym(dl, "zseek_reader_stats");
    R.reader_close = (fn_reader_close)dlsym(dl, "zseek_reader_close");
    /* the writer symbols are optional so that the same scan harness can be pointed at a reader-only library */
    if (!R.reader_open_full || !R.pread || !R.read || !R.reader_stats || !R.reader_close) {
        dlclose(dl);
        return -2;
    }
    R.dl = dl;
    return 0;
}

/* ------------------------------------------------------------------ writer over memory ---- */
typedef struct { uint8_t *data; size_t len, cap; } membuf_t;

static bool membuf_write(const void *data, size_t size, void *user_data, void *call_data)
{
    (void)call_data;
    membuf_t *m = user_data;
    if (m->len + size > m->cap) {
        size_t ncap = m->cap ? m->cap : (1u << 20);
        while (ncap < m->len + size)
            ncap *= 2;
        uint8_t *nd = realloc(m->data, ncap);
        if (!nd)
            return false;
        m->data = nd;
        m->cap = ncap;
    }
    memcpy(m->data + m->len, data, size);
    m->len += size;
    return true;
}

/*
 * type: 0 zstd, 1 lz4.  chunk = constant zseek_write size (always feed the writer a
 * constant chunk size).  On success returns 0 and a malloc'd image in *out (free with refdrive_free).
 */
EXPORT int refdrive_compress(const void *src, size_t n, int type, int level, int strategy,
                             int nb_workers, size_t min_frame_size, size_t chunk, uint8_t **out,
                             size_t *out_len, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!R.dl || !R.writer_open_full || !R.write || !R.writer_close)
        return -1;
    membuf_t m = {0};
    ref_write_file_t wf = { &m, membuf_write };
    ref_compression_param_t p;
    memset(&p, 0, sizeof(p));
    p.type = type;
    if (type == 0) {
        p.params.zstd_params.nb_workers = nb_workers;
        p.params.zstd_params.compression_level = level;
        p.params.zstd_params.strategy = strategy;
    } else {
        p.params.lz4_params.compression_level = level;
    }
    void *w = R.writer_open_full(wf, &p, min_frame_size, NULL, errbuf);
    if (!w)
        return -2;
    const uint8_t *s = src;
    for (size_t off = 0; off < n; off += chunk) {
        size_t len = n - off < chunk ? n - off : chunk;
        if (!R.write(w, s + off, len, NULL, errbuf)) {
            R.writer_close(w, NULL, NULL);
            free(m.data);
            return -3;
        }
    }
    if (!R.writer_close(w, NULL, errbuf)) {
        free(m.data);
        return -4;
    }
    *out = m.data;
    *out_len = m.len;
    return 0;
}

EXPORT void refdrive_free(void *p) { free(p); }

/* ------------------------------------------------------------------ reader over memory ---- */
typedef struct { const uint8_t *data; size_t size; } memimg_t;

static ssize_t memimg_pread(void *data, size_t size, size_t offset, void *user_data, void *call_data)
{
    (void)call_data;
    memimg_t *m = user_data;
    if (offset >= m->size)
        return 0;
    size_t n = m->size - offset < size ? m->size - offset : size;
    memcpy(data, m->data + offset, n);
    return (ssize_t)n;
}

static ssize_t memimg_fsize(void *user_data, void *call_data)
{
    (void)call_data;
    return (ssize_t)((memimg_t *)user_data)->size;
}

typedef struct { memimg_t img; void *reader; } refreader_t;

EXPORT void *refdrive_reader_open(const void *image, size_t size, size_t cache_size,
                                  char errbuf[ZSEEK_ERRBUF_SIZE])
{
    if (!R.dl)
        return NULL;
    refreader_t *rr = calloc(1, sizeof(*rr));
    if (!rr)
        return NULL;
    rr->img.data = image;
    rr->img.size = size;
    zseek_read_file_t uf = { &rr->img, memimg_pread, memimg_fsize };
    rr->reader = R.reader_open_full(uf, cache_size, NULL, errbuf);
    if (!rr->reader) {
        free(rr);
        return NULL;
    }
    return rr;
}

EXPORT ssize_t refdrive_pread(void *h, void *buf, size_t count, size_t offset,
                              char errbuf[ZSEEK_ERRBUF_SIZE])
{
    refreader_t *rr = h;
    return R.pread(rr ? rr->reader : NULL, buf, count, offset, NULL, errbuf);
}

EXPORT ssize_t refdrive_read(void *h, void *buf, size_t count, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    refreader_t *rr = h;
    return R.read(rr->reader, buf, count, NULL, errbuf);
}

EXPORT int refdrive_stats(void *h, zseek_reader_stats_t *st, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    refreader_t *rr = h;
    return R.reader_stats(rr ? rr->reader : NULL, st, errbuf) ? 1 : 0;
}

EXPORT int refdrive_reader_close(void *h, char errbuf[ZSEEK_ERRBUF_SIZE])
{
    refreader_t *rr = h;
    if (!rr)
        return 1;
    bool ok = R.reader_close(rr->reader, NULL, errbuf);
    free(rr);
    return ok ? 1 : 0;
}

/* Full read of [offset, offset+count) looping on the short reads zseek_pread returns at frame
 * boundaries (reference test/example.c:64-80).  Returns bytes read or -1. */
EXPORT ssize_t refdrive_pread_full(void *h, void *buf, size_t count, size_t offset,
                                   char errbuf[ZSEEK_ERRBUF_SIZE])
{
    refreader_t *rr = h;
    size_t done = 0;
    while (done < count) {
        ssize_t r = R.pread(rr->reader, (uint8_t *)buf + done, count - done, offset + done, NULL, errbuf);
        if (r < 0)
            return -1;
        if (r == 0)
            break;
        done += (size_t)r;
    }
    return (ssize_t)done;
}

/* ------------------------------------------------------------------ CPU timing harness ---- */
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

typedef struct {
    const uint8_t *image; size_t size;
    int tid, threads, pin;
    size_t cache_size;
    /* scan */
    size_t lo, hi, req;
    uint8_t *dst; /* optional: where decoded bytes of [lo,hi) go (dst + off); else a scratch buffer */
    /* random */
    const uint64_t *offsets; size_t n_lo, n_hi, count;
    uint64_t fold; /* xor-fold of first bytes so the work cannot be elided */
    size_t bytes; size_t ops; int err;
    pthread_barrier_t *bar;
    double t_start, t_end;
    void *shared_rd; /* refdrive_set_shared(1): every thread calls into this one reader instead of opening its own */
} job_t;

static int G_shared;
EXPORT void refdrive_set_shared(int on) { G_shared = on; }

static void pin_to(int cpu)
{
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(cpu, &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
}

static void *scan_thread(void *arg)
{
    job_t *j = arg;
    char errbuf[ZSEEK_ERRBUF_SIZE];
    if (j->pin)
        pin_to(j->tid % (int)sysconf(_SC_NPROCESSORS_ONLN));
    memimg_t img = { j->image, j->size };
    zseek_read_file_t uf = { &img, memimg_pread, memimg_fsize };
    void *rd = j->shared_rd ? j->shared_rd : R.reader_open_full(uf, j->cache_size, NULL, errbuf);
    uint8_t *scratch = j->dst ? NULL : malloc(j->req);
    if (!rd || (!j->dst && !scratch)) {
        j->err = 1;
        pthread_barrier_wait(j->bar);
        return NULL;
    }
    pthread_barrier_wait(j->bar);
    j->t_start = now_s();
    size_t off = j->lo;
    while (off < j->hi) {
        size_t want = j->hi - off < j->req ? j->hi - off : j->req;
        uint8_t *out = j->dst ? j->dst + off : scratch;
        size_t got = 0;
        while (got < want) { /* loop on short reads at frame boundaries */
            ssize_t r = R.pread(rd, out + got, want - got, off + got, NULL, errbuf);
            if (r <= 0) { j->err = 2; break; }
            got += (size_t)r;
        }
        if (j->err)
            break;
        j->fold ^= out[0];
        off += got;
        j->bytes += got;
    }
    j->t_end = now_s();
    if (!j->shared_rd)
        R.reader_close(rd, NULL, errbuf);
    free(scratch);
    return NULL;
}

/*
 * Sequential baseline: threads T, thread t scans the contiguous 1/T slice of decompressed range
 * [0, dsize) with `req`-byte requests.  Returns wall seconds (first start → last end), or <0.
 */
EXPORT double refdrive_scan(const void *image, size_t size, size_t dsize, int threads, size_t req,
                            size_t cache_size, int pin, uint8_t *dst, uint64_t *bytes_out)
{
    if (!R.dl || threads < 1)
        return -1.0;
    job_t *jobs = calloc((size_t)threads, sizeof(job_t));
    pthread_t *th = calloc((size_t)threads, sizeof(pthread_t));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)threads);
    char serr[ZSEEK_ERRBUF_SIZE];
    memimg_t simg = { image, size };
    zseek_read_file_t suf = { &simg, memimg_pread, memimg_fsize };
    void *shared = G_shared ? R.reader_open_full(suf, cache_size, NULL, serr) : NULL;
    if (G_shared && !shared)
        return -9.0;
    for (int t = 0; t < threads; t++) {
        job_t *j = &jobs[t];
        j->shared_rd = shared;
        j->image = image; j->size = size; j->tid = t; j->threads = threads; j->pin = pin;
        j->cache_size = cache_size; j->req = req; j->dst = dst; j->bar = &bar;
        j->lo = (size_t)((unsigned __int128)dsize * (unsigned)t / (unsigned)threads);
        j->hi = (size_t)((unsigned __int128)dsize * (unsigned)(t + 1) / (unsigned)threads);
        pthread_create(&th[t], NULL, scan_thread, j);
    }
    double t0 = 1e300, t1 = 0;
    uint64_t bytes = 0;
    int err = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].err) err = jobs[t].err;
        if (jobs[t].t_start < t0) t0 = jobs[t].t_start;
        if (jobs[t].t_end > t1) t1 = jobs[t].t_end;
        bytes += jobs[t].bytes;
    }
    pthread_barrier_destroy(&bar);
    free(jobs); free(th);
    if (shared) R.reader_close(shared, NULL, serr);
    if (bytes_out) *bytes_out = bytes;
    return err ? -(double)err : t1 - t0;
}

static void *random_thread(void *arg)
{
    job_t *j = arg;
    char errbuf[ZSEEK_ERRBUF_SIZE];
    if (j->pin)
        pin_to(j->tid % (int)sysconf(_SC_NPROCESSORS_ONLN));
    memimg_t img = { j->image, j->size };
    zseek_read_file_t uf = { &img, memimg_pread, memimg_fsize };
    void *rd = j->shared_rd ? j->shared_rd : R.reader_open_full(uf, j->cache_size, NULL, errbuf);
    uint8_t *scratch = malloc(j->count ? j->count : 1);
    if (!rd || !scratch) {
        j->err = 1;
        pthread_barrier_wait(j->bar);
        return NULL;
    }
    pthread_barrier_wait(j->bar);
    j->t_start = now_s();
    for (size_t i = j->n_lo; i < j->n_hi; i++) {
        size_t off = (size_t)j->offsets[i], got = 0;
        uint8_t *out = j->dst ? j->dst + i * j->count : scratch;
        while (got < j->count) { /* a request straddling a frame boundary is completed by re-issuing */
            ssize_t r = R.pread(rd, out + got, j->count - got, off + got, NULL, errbuf);
            if (r < 0) { j->err = 2; break; }
            if (r == 0) break;
            got += (size_t)r;
        }
        if (j->err)
            break;
        j->fold ^= out[0];
        j->bytes += got;
        j->ops++;
    }
    j->t_end = now_s();
    if (!j->shared_rd)
        R.reader_close(rd, NULL, errbuf);
    free(scratch);
    return NULL;
}

/* Random baseline: n requests (offsets[i], count) split evenly over T threads, one reader each. */
EXPORT double refdrive_random(const void *image, size_t size, const uint64_t *offsets, size_t n,
                              size_t count, int threads, size_t cache_size, int pin, uint8_t *dst,
                              uint64_t *ops_out)
{
    if (!R.dl || threads < 1)
        return -1.0;
    job_t *jobs = calloc((size_t)threads, sizeof(job_t));
    pthread_t *th = calloc((size_t)threads, sizeof(pthread_t));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)threads);
    char serr[ZSEEK_ERRBUF_SIZE];
    memimg_t simg = { image, size };
    zseek_read_file_t suf = { &simg, memimg_pread, memimg_fsize };
    void *shared = G_shared ? R.reader_open_full(suf, cache_size, NULL, serr) : NULL;
    if (G_shared && !shared)
        return -9.0;
    for (int t = 0; t < threads; t++) {
        job_t *j = &jobs[t];
        j->shared_rd = shared;
        j->image = image; j->size = size; j->tid = t; j->threads = threads; j->pin = pin;
        j->cache_size = cache_size; j->offsets = offsets; j->count = count; j->dst = dst; j->bar = &bar;
        j->n_lo = n * (size_t)t / (size_t)threads;
        j->n_hi = n * (size_t)(t + 1) / (size_t)threads;
        pthread_create(&th[t], NULL, random_thread, j);
    }
    double t0 = 1e300, t1 = 0;
    uint64_t ops = 0;
    int err = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].err) err = jobs[t].err;
        if (jobs[t].t_start < t0) t0 = jobs[t].t_start;
        if (jobs[t].t_end > t1) t1 = jobs[t].t_end;
        ops += jobs[t].ops;
    }
    pthread_barrier_destroy(&bar);
    free(jobs); free(th);
    if (shared) R.reader_close(shared, NULL, serr);
    if (ops_out) *ops_out = ops;
    return err ? -(double)err : t1 - t0;
}
