/*
 * zseek_b200.h — additive GPU entry points of libzseek_b200.so.
 *
 * None of these exist in the reference; they are what SURVEY.md §8(b) calls "additive extensions
 * needed for the metric" and they never change the six reference signatures in zseek.h.  Each one is
 * DEFINED in terms of the reference API so that parity stays checkable:
 *
 *   zseek_b200_pread_batch  result[i] / bytes == what zseek_pread(reader, dst_i, count_i, offset_i)
 *                           returns / stores (reference src/decompress.c:806-824, semantics B1-B4)
 *   zseek_b200_read_range   == looping zseek_pread over the short reads at frame boundaries, exactly
 *                           like reference test/example.c:64-80 (the reference's own TODO at
 *                           src/decompress.c:473-474: "return as much as possible (multiple frames)")
 *   zseek_b200_decode_frames == read_range over whole frames [lo, hi), device destination
 *
 * `buf`/`dst` arguments of these functions AND of zseek_pread/zseek_read may be device pointers
 * (detected with cudaPointerGetAttributes); then no byte crosses PCIe.
 *
 * Environment (read at open): ZSEEK_B200_DEVICE (ordinal; default LOCAL_RANK, else current device),
 * ZSEEK_B200_READAHEAD (max frames decoded ahead of a sequential scan; 0 disables),
 * ZSEEK_B200_STAGE_MB (pinned ingest staging, default 64).  INTEGRATION.md lists every knob (read-ahead ramp,
 * residency of random host readers, parked readers, pinned-memory budgets, pread(2) worker threads).
 */
#ifndef ZSEEK_B200_H
#define ZSEEK_B200_H
#include "zseek.h"
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Opens a reader over a compressed file image that already sits in host memory (pageable or
 * pinned).  Same results as zseek_reader_open_full with a memcpy pread callback; compressed bytes
 * are DMA'd to HBM straight from `image` (no staging copy).  The image must outlive the reader. */
ZSEEK_EXPORT zseek_reader_t *zseek_b200_reader_open_mem(const void *image, size_t size, size_t cache_size,
                                                        char errbuf[ZSEEK_ERRBUF_SIZE]);

/* Frame-range sharding: this reader serves frames [N*rank/world, N*(rank+1)/world).
 * Reads that resolve to a frame outside the shard fail with "frame outside this reader's shard". */
ZSEEK_EXPORT bool zseek_b200_set_shard(zseek_reader_t *reader, unsigned rank, unsigned world,
                                       char errbuf[ZSEEK_ERRBUF_SIZE]);
ZSEEK_EXPORT bool zseek_b200_get_shard(zseek_reader_t *reader, size_t *frame_lo, size_t *frame_hi);

/* Host view of the parsed seek table: *n frames, prefix arrays of n+1 entries, valid until close. */
ZSEEK_EXPORT bool zseek_b200_seek_table(zseek_reader_t *reader, size_t *n, const uint64_t **c_off,
                                        const uint64_t **d_off, int *codec);

/* Makes the compressed bytes of frames [frame_lo, frame_hi) resident in HBM (pulled through the
 * reader's pread callback via pinned staging, or DMA'd from the memory image). */
ZSEEK_EXPORT bool zseek_b200_load(zseek_reader_t *reader, size_t frame_lo, size_t frame_hi, void *call_data,
                                  char errbuf[ZSEEK_ERRBUF_SIZE]);

/* Decodes whole frames [frame_lo, frame_hi) into the DEVICE buffer dev_dst (frame f lands at
 * dev_dst + d_off[f] - d_off[frame_lo]).  One kernel launch when the compressed range is resident.
 * Returns the number of bytes produced or -1. */
ZSEEK_EXPORT ssize_t zseek_b200_decode_frames(zseek_reader_t *reader, size_t frame_lo, size_t frame_hi,
                                              void *dev_dst, void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE]);

/* Multi-frame read: up to `count` bytes starting at decompressed `offset`, crossing frame boundaries;
 * buf may be host or device memory.  Returns bytes read (short only at EOF) or -1. */
ZSEEK_EXPORT ssize_t zseek_b200_read_range(zseek_reader_t *reader, void *buf, size_t count, size_t offset,
                                           void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE]);

/* n independent reads.  Request i reads counts[i] (or fixed_count when counts == NULL) bytes at
 * offsets[i] into dst + (dst_offs ? dst_offs[i] : i * dst_stride); results[i] (optional) receives what
 * zseek_pread would have returned (never crosses a frame boundary; 0 at/after EOF).  offsets, counts,
 * dst_offs and results are host arrays; dst may be host or device memory.  Returns n or -1. */
ZSEEK_EXPORT ssize_t zseek_b200_pread_batch(zseek_reader_t *reader, size_t n, const uint64_t *offsets,
                                            const uint64_t *counts, uint64_t fixed_count, void *dst,
                                            const uint64_t *dst_offs, uint64_t dst_stride, int64_t *results,
                                            void *call_data, char errbuf[ZSEEK_ERRBUF_SIZE]);

/* Stream-ordered batch: the same n reads, but every array lives in DEVICE memory (dev_offsets,
 * dev_counts or NULL + fixed_count, dev_dst with dev_dst_offs or dst_stride, optional dev_results) and nothing is
 * copied to or from the host: lookup, the list of touched frames, their decode (each only up to the last byte the
 * batch needs of it) and the gather are queued on `stream` — a cudaStream_t passed as void*, NULL = the reader's own
 * stream — and the call returns without waiting.  Work queued on `stream` afterwards sees dev_dst / dev_results
 * complete.  The compressed bytes of the reader's shard must be resident (zseek_b200_load); frames are decoded into a
 * batch slab of the reader, not into the LRU cache.  Results are defined exactly like zseek_b200_pread_batch.
 * zseek_b200_batch_wait waits for the most recent async batch and returns 0, or -1 with the first frame error
 * ("decompress frame: ...", "frame outside this reader's shard"); dev_dst of requests in failed frames is undefined.
 * One async batch per reader may be in flight (the next call waits for the previous one).  Every other entry point of
 * the reader may be called while a batch is pending: its device work is ordered after the batch (an event wait on the
 * reader's streams, the host does not block).  The caller's stream must stay alive until zseek_b200_batch_wait. */
ZSEEK_EXPORT ssize_t zseek_b200_pread_batch_async(zseek_reader_t *reader, size_t n, const uint64_t *dev_offsets,
                                                  const uint64_t *dev_counts, uint64_t fixed_count, void *dev_dst,
                                                  const uint64_t *dev_dst_offs, uint64_t dst_stride, int64_t *dev_results,
                                                  void *stream, char errbuf[ZSEEK_ERRBUF_SIZE]);
ZSEEK_EXPORT int zseek_b200_batch_wait(zseek_reader_t *reader, char errbuf[ZSEEK_ERRBUF_SIZE]);

/* Forgets the HBM-resident compressed image (the next read pulls its frames again). */
ZSEEK_EXPORT void zseek_b200_unload(zseek_reader_t *reader);

/* Device-side stopwatch on the stream the kernels are launched on (CUDA events): start records an
 * event, stop records a second one, waits for it and returns the milliseconds in between (<0 on error). */
ZSEEK_EXPORT bool zseek_b200_timer_start(zseek_reader_t *reader);
ZSEEK_EXPORT double zseek_b200_timer_stop(zseek_reader_t *reader);

/* Drops every decoded frame from the HBM cache (benchmarks use it to time cold batches). */
ZSEEK_EXPORT void zseek_b200_cache_clear(zseek_reader_t *reader);

/* Instrumentation: kernels launched so far by this reader; device time of its most recent decode
 * kernel in milliseconds (CUDA events on the launching stream; <0 if none). */
ZSEEK_EXPORT unsigned long long zseek_b200_launch_count(zseek_reader_t *reader);
ZSEEK_EXPORT double zseek_b200_last_decode_ms(zseek_reader_t *reader);
/* Name of the decode kernel of that launch (the LZ4 path picks a kernel by the number of frames in the
 * launch); a static string, "" before the first launch. */
ZSEEK_EXPORT const char *zseek_b200_last_decode_kernel(zseek_reader_t *reader);
ZSEEK_EXPORT int zseek_b200_device(zseek_reader_t *reader);

#ifdef __cplusplus
}
#endif
#endif
