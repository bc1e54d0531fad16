/*
 * zsk_cuda.h — the thin C-ABI launch layer between the C host reader (reader.c) and CUDA
 * (zsk_cuda.cu).  Plain pointers and sizes only; every function returns 0 on success or a non-zero
 * code with a message retrievable through zsk_cuda_error().  Internal to libzseek_b200.so (hidden
 * visibility); the public surface is include/zseek.h + include/zseek_b200.h.
 */
#ifndef ZSK_CUDA_H
#define ZSK_CUDA_H
#include "zsk_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zsk_cuda_ctx zsk_cuda_ctx;

/* streams of a context */
enum { ZSK_STREAM_COMPUTE = 0, ZSK_STREAM_H2D = 1, ZSK_STREAM_D2H = 2, ZSK_NSTREAMS = 3, ZSK_STREAM_USER = 3 };
/* copy kinds */
enum { ZSK_H2D = 1, ZSK_D2H = 2, ZSK_D2D = 3 };

/* device < 0: take ZSEEK_B200_DEVICE, else LOCAL_RANK, else the current CUDA device */
int zsk_cuda_ctx_create(int device, zsk_cuda_ctx **out, char *err, size_t errlen);
void zsk_cuda_ctx_destroy(zsk_cuda_ctx *cx);
/* contexts of closed readers are parked and handed to later readers of the same device (reader.c) */
int zsk_cuda_pick_device(void);                               /* the device zsk_cuda_ctx_create(-1, ...) would take; -1: none */
int zsk_cuda_ctx_reuse(zsk_cuda_ctx *cx);                     /* re-reads the tuning knobs, forgets per-reader statistics */
void zsk_cuda_ctx_trim(zsk_cuda_ctx *cx, size_t max_bytes);   /* frees the scratch pools when they hold more than max_bytes */
size_t zsk_cuda_ctx_held(const zsk_cuda_ctx *cx);            /* device bytes of the scratch pools */
const char *zsk_cuda_error(zsk_cuda_ctx *cx);
int zsk_cuda_device(const zsk_cuda_ctx *cx);
int zsk_cuda_sm_count(const zsk_cuda_ctx *cx);
unsigned long long zsk_cuda_launch_count(const zsk_cuda_ctx *cx);
size_t zsk_cuda_free_memory(zsk_cuda_ctx *cx);

int zsk_cuda_malloc(zsk_cuda_ctx *cx, void **p, size_t n);
int zsk_cuda_free(zsk_cuda_ctx *cx, void *p);
int zsk_cuda_malloc_host(zsk_cuda_ctx *cx, void **p, size_t n);   /* pinned */
int zsk_cuda_free_host(zsk_cuda_ctx *cx, void *p);
int zsk_cuda_memset_async(zsk_cuda_ctx *cx, void *p, int v, size_t n, int stream);
int zsk_cuda_memcpy_async(zsk_cuda_ctx *cx, void *dst, const void *src, size_t n, int kind, int stream);
int zsk_cuda_stream_sync(zsk_cuda_ctx *cx, int stream);       /* the caller sleeps (blocking event) */
int zsk_cuda_stream_sync_spin(zsk_cuda_ctx *cx, int stream);  /* the caller spins: lowest latency, for tiny copies */
/* ZSK_STREAM_USER: a caller-owned cudaStream_t (NULL = the legacy default stream) that stream-ordered calls enqueue on */
void zsk_cuda_set_user_stream(zsk_cuda_ctx *cx, void *stream);
int zsk_cuda_stream_wait(zsk_cuda_ctx *cx, int waiter, int signaler); /* waiter waits for work queued on signaler so far */
/* a small pool of user events: record on a stream, block the host until it has completed */
#define ZSK_NEVENTS 8
int zsk_cuda_event_record(zsk_cuda_ctx *cx, int ev, int stream);
int zsk_cuda_event_sync(zsk_cuda_ctx *cx, int ev);
int zsk_cuda_stream_wait_event(zsk_cuda_ctx *cx, int stream, int ev); /* stream waits for the event's latest record */
/* 1 = device memory, 0 = host (pageable or pinned), <0 = error */
int zsk_cuda_pointer_is_device(zsk_cuda_ctx *cx, const void *p);

/* kernels (asynchronous on `stream`) */
int zsk_cuda_launch_decode(zsk_cuda_ctx *cx, int codec, const zsk_decode_args *args, int stream);
int zsk_cuda_launch_lookup(zsk_cuda_ctx *cx, const zsk_lookup_args *args, int stream);
int zsk_cuda_launch_gather(zsk_cuda_ctx *cx, const zsk_gather_args *args, int stream);
int zsk_cuda_launch_compact(zsk_cuda_ctx *cx, const zsk_compact_args *args, int stream);

/* device-side timing of what is queued on `stream` between start and stop (CUDA events) */
int zsk_cuda_timer_start(zsk_cuda_ctx *cx, int stream);
int zsk_cuda_timer_stop(zsk_cuda_ctx *cx, int stream, float *ms); /* synchronises the stream */
/* duration of the most recent decode kernel launched through this context (events around the launch,
 * valid after the stream was synchronised) */
int zsk_cuda_last_decode_ms(zsk_cuda_ctx *cx, float *ms);
const char *zsk_cuda_last_decode_kernel(const zsk_cuda_ctx *cx); /* "" before the first launch */

/* diagnostic timeline (ZSEEK_B200_TRACE=1): timestamps of up to ZSK_NTRACE points queued on the streams */
#define ZSK_NTRACE 192
int zsk_cuda_trace_enabled(const zsk_cuda_ctx *cx);
void zsk_cuda_trace_reset(zsk_cuda_ctx *cx);
void zsk_cuda_trace_mark(zsk_cuda_ctx *cx, int stream, const char *what, unsigned k);
void zsk_cuda_trace_dump(zsk_cuda_ctx *cx); /* synchronises the device; prints to stderr */

#ifdef __cplusplus
}
#endif
#endif
