/*
 * zseek.h — reader half of the libzseek public API, served by the B200-native implementation.
 *
 * This header is the DROP-IN BOUNDARY.  Every type below is layout-identical to, and every function
 * has the same name, argument meaning, return convention and error text as, the reference header
 * /root/reference/src/zseek.h (v3.0.2):
 *
 *   zseek_pread_t / zseek_fsize_t / zseek_read_file_t      reference src/zseek.h:88-116
 *   zseek_reader_stats_t                                    reference src/zseek.h:190-203
 *   zseek_reader_open_full                                  reference src/zseek.h:335-336  (impl src/decompress.c:261-288)
 *   zseek_reader_open                                       reference src/zseek.h:355-356  (impl src/decompress.c:290-295)
 *   zseek_reader_close                                      reference src/zseek.h:374-375  (impl src/decompress.c:359-375)
 *   zseek_pread                                             reference src/zseek.h:398-399  (impl src/decompress.c:806-824)
 *   zseek_read                                              reference src/zseek.h:422-423  (impl src/decompress.c:826-835)
 *   zseek_reader_stats                                      reference src/zseek.h:442-443  (impl src/decompress.c:837-891)
 *
 * A caller that today links -lzseek and uses only the reader functions can link libzseek_b200.so
 * instead without changing a line.  The writer half of the reference API
 * (zseek_writer_*, zseek_write) is deliberately NOT provided: files are produced by the reference
 * CPU writer.  GPU-specific additions (batched reads, device-pointer destinations, frame-range
 * sharding) live in zseek_b200.h and never alter the six signatures below.
 *
 * There is no CPU decode path behind these functions: if no CUDA device is usable,
 * zseek_reader_open* fails with an error message.
 */
#ifndef ZSEEK_H
#define ZSEEK_H

#include <stddef.h>
#include <stdbool.h>
#include <stdio.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZSEEK_EXPORT __attribute__((visibility("default")))

/** Size of the caller-provided error message buffer (reference src/zseek.h:36). */
#define ZSEEK_ERRBUF_SIZE 80

/**
 * Read callback: store up to @p size bytes found at file offset @p offset into @p data.
 * Returns the number of bytes stored (short only at EOF) or <0 on error.
 */
typedef ssize_t (*zseek_pread_t)(void *data, size_t size, size_t offset, void *user_data,
                                 void *call_data);

/** File-size callback: returns the compressed file size in bytes, or <0 on error. */
typedef ssize_t (*zseek_fsize_t)(void *user_data, void *call_data);

/** User-defined readable file (passed BY VALUE to zseek_reader_open_full). */
typedef struct {
    void *user_data;
    zseek_pread_t pread;
    zseek_fsize_t fsize;
} zseek_read_file_t;

/** Opaque reader handle. */
typedef struct zseek_reader zseek_reader_t;

/** Reader statistics: six size_t, same order as the reference. */
typedef struct {
    size_t seek_table_memory; /* 24 + 24*frames, exactly as the reference reports it */
    size_t frames;
    size_t decompressed_size;
    size_t cache_memory;      /* bytes of HBM held by the decoded-frame cache */
    size_t cached_frames;     /* frames currently resident in the HBM cache */
    size_t buffer_size;       /* pinned-host staging + device compressed image, bytes */
} zseek_reader_stats_t;

ZSEEK_EXPORT zseek_reader_t *zseek_reader_open_full(zseek_read_file_t user_file, size_t cache_size,
                                                    void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE]);

ZSEEK_EXPORT zseek_reader_t *zseek_reader_open(FILE *cfile, size_t cache_size, void *call_data,
                                               char errbuf[ZSEEK_ERRBUF_SIZE]);

ZSEEK_EXPORT bool zseek_reader_close(zseek_reader_t *reader, void *call_data,
                                     char errbuf[ZSEEK_ERRBUF_SIZE]);

ZSEEK_EXPORT ssize_t zseek_pread(zseek_reader_t *reader, void *buf, size_t count, size_t offset,
                                 void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE]);

ZSEEK_EXPORT ssize_t zseek_read(zseek_reader_t *reader, void *buf, size_t count, void *call_data,
                                char errbuf[ZSEEK_ERRBUF_SIZE]);

ZSEEK_EXPORT bool zseek_reader_stats(zseek_reader_t *reader, zseek_reader_stats_t *stats,
                                     char errbuf[ZSEEK_ERRBUF_SIZE]);

#ifdef __cplusplus
}
#endif
#endif /* ZSEEK_H */
