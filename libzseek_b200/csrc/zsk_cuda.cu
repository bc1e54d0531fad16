/*
 * zsk_cuda.cu — C-ABI launch layer: device selection, memory, streams, events and the kernel launches.
 * Compiled by nvcc for sm_100a only (-gencode arch=compute_100a,code=sm_100a).  There is no host
 * decode path in this file or anywhere else in the library: if no CUDA device is usable,
 * zsk_cuda_ctx_create fails and zseek_reader_open* reports the error.
 */
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "zsk_cuda.h"
#include "zsk_lz4.cuh"
#include "zsk_lz4_lane.cuh"
#include "zsk_zstd_pipe.cuh"
#include "zsk_seek.cuh"

#define ZSK_NCOUNTERS 64

#include <time.h>
/* ZSEEK_B200_DEBUG=1: slow allocations are reported on stderr */
static int zsk_dbg(void)
{
    static int v = -1;
    if (v < 0) v = getenv("ZSEEK_B200_DEBUG") != NULL;
    return v;
}
static double zsk_now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

struct zsk_cuda_ctx {
    int device;
    int sm_count;
    cudaStream_t alloc_stream;            /* cudaMallocAsync / cudaFreeAsync */
    cudaStream_t streams[ZSK_NSTREAMS + 1];   /* + ZSK_STREAM_USER: caller-owned, never created or destroyed here */
    cudaEvent_t sync_ev;                 /* cross-stream dependencies */
    cudaEvent_t block_ev[ZSK_NSTREAMS + 1]; /* host waits: blocking events, so that a waiting caller thread sleeps instead of spinning */
    cudaEvent_t user_ev[ZSK_NEVENTS];
    cudaEvent_t t0, t1;                  /* zsk_cuda_timer_* */
    cudaEvent_t k0, k1;                  /* around the most recent decode kernel */
    int k_valid;
    const char *k_name;                  /* name of the most recent decode kernel */
    uint32_t *counters;                  /* ZSK_NCOUNTERS work counters, used round-robin */
    unsigned counter_next;
    uint8_t *scratch;                    /* literal scratch of the one-CTA-per-frame zstd kernel, ZSK_LIT_SCRATCH bytes for each of scratch_ctas CTAs; lazily allocated */
    size_t scratch_ctas;
    int zstd_ctas, lz4_ctas;
    /* zstd pipeline (zsk_zstd_pipe.cuh): scratch pools grown on demand, per-launch counter blocks used round-robin */
    zsk_zframe *zframes;
    uint32_t *zdeferred;
    size_t zjobs_cap;
    zsk_zblock *zblocks;
    uint32_t *zbprog;                    /* 2 words per block descriptor (launches with limits) */
    size_t zbprog_cap;
    uint32_t *zseqs;
    uint8_t *zlits;
    size_t zblocks_cap, zseqs_cap, zlits_cap;
    unsigned long long *zctr;            /* ZSK_NCOUNTERS blocks of ZSK_ZC_N counters */
    unsigned zctr_next;
    int zfse_ctas, zhuf_ctas, zexec_ctas;
    uint64_t zwave_bytes;                /* ZSEEK_B200_ZSTD_WAVE_MB (default 4096) */
    int zstd_legacy;                     /* ZSEEK_B200_ZSTD_LEGACY=1: every frame goes to the one-CTA-per-frame kernel (A/B runs) */
    int lz4_lane_ctas;                   /* resident CTAs of the lane-per-frame kernel */
    unsigned lz4_lane_min;               /* launches with at least this many frames use the lane-per-frame kernel */
    unsigned long long launches;
    int trace, trace_ready;              /* ZSEEK_B200_TRACE=1: timeline of the host-destination pipeline */
    cudaEvent_t trace_ev[ZSK_NTRACE];
    const char *trace_what[ZSK_NTRACE];
    unsigned trace_k[ZSK_NTRACE], trace_n;
    char err[160];
};

#define CK(cx, call)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            snprintf((cx)->err, sizeof((cx)->err), "%s: %s", #call, cudaGetErrorString(e_));           \
            return (int)e_ ? (int)e_ : -1;                                                             \
        }                                                                                              \
    } while (0)

extern "C" {

/* the device a context created now would use: ZSEEK_B200_DEVICE, else LOCAL_RANK, else the current CUDA device; -1 without one */
int zsk_cuda_pick_device(void)
{
    int ndev = 0, device = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return -1;
    }
    const char *s = getenv("ZSEEK_B200_DEVICE");
    if (!s) s = getenv("LOCAL_RANK");
    if (s) device = atoi(s) % ndev;
    else if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    return device;
}

/* launch geometry and tuning knobs (environment); run when a context is created and again when a parked one is reused.
 * The occupancy figures are properties of the kernels and the device: asked once per device. */
static int ctx_configure(zsk_cuda_ctx *cx, char *err, size_t errlen)
{
    enum { K_ZSTD = 0, K_FSE, K_HUF, K_EXEC, K_LZ4, K_LANE, K_N };
    static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    static int occ[64][K_N], have[64];
    int per_sm = 0;
#define CKC(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e_));                            \
            pthread_mutex_unlock(&mu);                                                                 \
            return -1;                                                                                 \
        }                                                                                              \
    } while (0)
    pthread_mutex_lock(&mu);
    const int di = cx->device & 63;
    if (!have[di]) {
        CKC(cudaFuncSetAttribute(zsk_zstd_fse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZSK_ZFSE_SMEM));
        CKC(cudaFuncSetAttribute(zsk_zstd_huf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZSK_ZHUF_SMEM));
        CKC(cudaFuncSetAttribute(zsk_zstd_exec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZSK_ZX_SMEM));
#ifdef ZSK_EXP_FSE_CARVEOUT
        CKC(cudaFuncSetAttribute(zsk_zstd_fse_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, ZSK_EXP_FSE_CARVEOUT));
#endif
        CKC(cudaFuncSetAttribute(zsk_zstd_huf_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CKC(cudaFuncSetAttribute(zsk_zstd_exec_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CKC(cudaFuncSetAttribute(zsk_lz4_decode_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZSK_LZ4L_SMEM));
        CKC(cudaFuncSetAttribute(zsk_lz4_decode_lane_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[di][K_ZSTD], zsk_zstd_decode_kernel, ZSK_ZSTD_CTA_THREADS, 0));
        CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[di][K_FSE], zsk_zstd_fse_kernel, ZSK_ZFSE_THREADS, ZSK_ZFSE_SMEM));
        CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[di][K_HUF], zsk_zstd_huf_kernel, 32, ZSK_ZHUF_SMEM));
        CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[di][K_EXEC], zsk_zstd_exec_kernel, 32, ZSK_ZX_SMEM));
        CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[di][K_LZ4], zsk_lz4_decode_batch_kernel, ZSK_LZ4_CTA_THREADS, 0));
        CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[di][K_LANE], zsk_lz4_decode_lane_kernel, ZSK_LZ4L_THREADS, ZSK_LZ4L_SMEM));
        for (int k = 0; k < K_N; k++)
            if (occ[di][k] < 1) occ[di][k] = 1;
        have[di] = 1;
    }
    int o[K_N];
    memcpy(o, occ[di], sizeof(o));
    pthread_mutex_unlock(&mu);
#undef CKC
    per_sm = o[K_ZSTD];
    if (const char *g = getenv("ZSEEK_B200_ZSTD_CTAS_PER_SM")) { /* tuning knob: resident zstd CTAs per SM */
        int v = atoi(g);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    cx->zstd_ctas = per_sm * cx->sm_count;
    cx->zstd_legacy = 0;
    if (const char *g = getenv("ZSEEK_B200_ZSTD_LEGACY")) cx->zstd_legacy = atoi(g);
    cx->zwave_bytes = (uint64_t)4096 << 20;
    if (const char *g = getenv("ZSEEK_B200_ZSTD_WAVE_MB")) {
        const unsigned long long v = strtoull(g, NULL, 10);
        if (v >= 1) cx->zwave_bytes = (uint64_t)v << 20;
    }
    cx->zfse_ctas = o[K_FSE] * cx->sm_count;
    cx->zhuf_ctas = o[K_HUF] * cx->sm_count;
    per_sm = o[K_EXEC];
    if (const char *g = getenv("ZSEEK_B200_ZEXEC_CTAS_PER_SM")) { /* tuning knob: resident executor warps per SM */
        int v = atoi(g);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    cx->zexec_ctas = per_sm * cx->sm_count;
    per_sm = o[K_LZ4];
    if (const char *g = getenv("ZSEEK_B200_LZ4_CTAS_PER_SM")) { /* tuning knob: resident LZ4 CTAs per SM */
        int v = atoi(g);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    cx->lz4_ctas = per_sm * cx->sm_count;
    /* lane-per-frame kernel: wins once there are enough frames to fill the lanes of the GPU (one frame per lane
     * runs ~10x slower than one frame per warp, but 32x more of them run at once) */
    per_sm = o[K_LANE];
    if (const char *g = getenv("ZSEEK_B200_LZ4_LANE_CTAS_PER_SM")) {
        int v = atoi(g);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    cx->lz4_lane_ctas = per_sm * cx->sm_count;
    cx->lz4_lane_min = 40960;             /* measured crossover: 32,768 frames 11.6 ms (warp per frame) vs 14.1 ms (lane per frame), 49,152 frames 18.6 vs 15.9 ms */
    if (const char *g = getenv("ZSEEK_B200_LZ4_LANE_MIN")) cx->lz4_lane_min = (unsigned)strtoul(g, NULL, 10); /* 0 = always, huge = never */
    const int trace = getenv("ZSEEK_B200_TRACE") ? atoi(getenv("ZSEEK_B200_TRACE")) : 0;
    if (trace && !cx->trace_ready) {
        for (int i = 0; i < ZSK_NTRACE; i++)
            if (cudaEventCreate(&cx->trace_ev[i]) != cudaSuccess) {
                snprintf(err, errlen, "cudaEventCreate failed");
                return -1;
            }
        cx->trace_ready = 1;
    }
    cx->trace = trace;
    return 0;
}

/* a parked context goes to a new reader: knobs re-read, per-reader statistics and the caller's stream forgotten */
int zsk_cuda_ctx_reuse(zsk_cuda_ctx *cx)
{
    cx->streams[ZSK_STREAM_USER] = NULL;
    cx->launches = 0;
    cx->k_valid = 0;
    cx->k_name = NULL;
    cx->trace_n = 0;
    CK(cx, cudaSetDevice(cx->device));
    return ctx_configure(cx, cx->err, sizeof(cx->err));
}

/* device bytes of the scratch pools a context holds */
size_t zsk_cuda_ctx_held(const zsk_cuda_ctx *cx)
{
    return cx->zblocks_cap * sizeof(zsk_zblock) + cx->zseqs_cap * 3 * sizeof(uint32_t) + cx->zlits_cap +
           cx->zjobs_cap * (sizeof(zsk_zframe) + sizeof(uint32_t)) + cx->zbprog_cap * 2 * sizeof(uint32_t) +
           cx->scratch_ctas * ZSK_LIT_SCRATCH;
}

/* gives back the scratch pools of a context that is being parked when they hold more than max_bytes */
void zsk_cuda_ctx_trim(zsk_cuda_ctx *cx, size_t max_bytes)
{
    cudaSetDevice(cx->device);
    if (zsk_cuda_ctx_held(cx) <= max_bytes) return;
    void *pools[] = { cx->zbprog, cx->zframes, cx->zdeferred, cx->zblocks, cx->zseqs, cx->zlits };
    for (size_t i = 0; i < sizeof(pools) / sizeof(pools[0]); i++)
        if (pools[i]) cudaFreeAsync(pools[i], cx->alloc_stream);
    cx->zbprog = NULL; cx->zframes = NULL; cx->zdeferred = NULL; cx->zblocks = NULL; cx->zseqs = NULL; cx->zlits = NULL;
    cx->zjobs_cap = cx->zbprog_cap = cx->zblocks_cap = cx->zseqs_cap = cx->zlits_cap = 0;
}

int zsk_cuda_ctx_create(int device, zsk_cuda_ctx **out, char *err, size_t errlen)
{
    /* Every reader brings three streams, and callers keep one reader per thread: with the default 8 hardware work queues
     * the streams of different readers share queues and wait for each other's events.  Only effective when this library
     * initialises CUDA in the process; an explicit setting of the caller is left alone. */
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        snprintf(err, errlen, "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "count is 0");
        return -1;
    }
    if (device < 0) device = zsk_cuda_pick_device();
    if (device >= ndev) {
        snprintf(err, errlen, "CUDA device %d out of range (%d present)", device, ndev);
        return -1;
    }
    zsk_cuda_ctx *cx = (zsk_cuda_ctx *)calloc(1, sizeof(*cx));
    if (!cx) {
        snprintf(err, errlen, "allocate device context");
        return -1;
    }
    cx->device = device;
#define CK0(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e_));                            \
            free(cx);                                                                                  \
            return -1;                                                                                 \
        }                                                                                              \
    } while (0)
    CK0(cudaSetDevice(device));
    int cc_major = 0, cc_minor = 0;
    CK0(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    CK0(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
    if (cc_major < 10) {
        snprintf(err, errlen, "device %d is sm_%d%d; this library ships sm_100a code only", device, cc_major, cc_minor);
        free(cx);
        return -1;
    }
    CK0(cudaDeviceGetAttribute(&cx->sm_count, cudaDevAttrMultiProcessorCount, device));
    for (int i = 0; i < ZSK_NSTREAMS; i++) CK0(cudaStreamCreateWithFlags(&cx->streams[i], cudaStreamNonBlocking));
    CK0(cudaStreamCreateWithFlags(&cx->alloc_stream, cudaStreamNonBlocking));
    {
        cudaMemPool_t pool;
        CK0(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull; /* freed buffers stay in the pool */
        CK0(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CK0(cudaEventCreateWithFlags(&cx->sync_ev, cudaEventDisableTiming));
    /* every host wait goes through a BLOCKING event: with one reader per caller thread and the threads pinned to all cores,
     * spinning waits starved the driver's own threads (measured: 100-2,000 ms for a one-frame miss at 16 threads) */
    for (int i = 0; i <= ZSK_NSTREAMS; i++) CK0(cudaEventCreateWithFlags(&cx->block_ev[i], cudaEventDisableTiming | cudaEventBlockingSync));
    for (int i = 0; i < ZSK_NEVENTS; i++) CK0(cudaEventCreateWithFlags(&cx->user_ev[i], cudaEventDisableTiming | cudaEventBlockingSync));
    CK0(cudaEventCreateWithFlags(&cx->t0, cudaEventBlockingSync));
    CK0(cudaEventCreateWithFlags(&cx->t1, cudaEventBlockingSync));
    CK0(cudaEventCreateWithFlags(&cx->k0, cudaEventBlockingSync));
    CK0(cudaEventCreateWithFlags(&cx->k1, cudaEventBlockingSync));
    CK0(cudaMallocAsync((void **)&cx->counters, ZSK_NCOUNTERS * sizeof(uint32_t), cx->alloc_stream));
    CK0(cudaMallocAsync((void **)&cx->zctr, ZSK_NCOUNTERS * ZSK_ZC_N * sizeof(unsigned long long), cx->alloc_stream));
    CK0(cudaStreamSynchronize(cx->alloc_stream));
    if (ctx_configure(cx, err, errlen)) {
        free(cx);
        return -1;
    }
#undef CK0
    *out = cx;
    return 0;
}

void zsk_cuda_ctx_destroy(zsk_cuda_ctx *cx)
{
    if (!cx) return;
    cudaSetDevice(cx->device);
    for (int i = 0; i < ZSK_NSTREAMS; i++) {
        cudaStreamSynchronize(cx->streams[i]);
        cudaStreamDestroy(cx->streams[i]);
    }
    cudaEventDestroy(cx->sync_ev);
    for (int i = 0; i <= ZSK_NSTREAMS; i++) cudaEventDestroy(cx->block_ev[i]);
    for (int i = 0; i < ZSK_NEVENTS; i++) cudaEventDestroy(cx->user_ev[i]);
    cudaEventDestroy(cx->t0);
    cudaEventDestroy(cx->t1);
    cudaEventDestroy(cx->k0);
    cudaEventDestroy(cx->k1);
    void *pools[] = { cx->counters, cx->scratch, cx->zctr, cx->zbprog, cx->zframes, cx->zdeferred, cx->zblocks, cx->zseqs, cx->zlits };
    for (size_t i = 0; i < sizeof(pools) / sizeof(pools[0]); i++)
        if (pools[i]) cudaFreeAsync(pools[i], cx->alloc_stream);
    cudaStreamSynchronize(cx->alloc_stream);
    cudaStreamDestroy(cx->alloc_stream);
    free(cx);
}

int zsk_cuda_trace_enabled(const zsk_cuda_ctx *cx) { return cx->trace; }
void zsk_cuda_trace_reset(zsk_cuda_ctx *cx) { cx->trace_n = 0; }
void zsk_cuda_trace_mark(zsk_cuda_ctx *cx, int stream, const char *what, unsigned k)
{
    if (!cx->trace || cx->trace_n >= ZSK_NTRACE) return;
    cx->trace_what[cx->trace_n] = what;
    cx->trace_k[cx->trace_n] = k;
    cudaEventRecord(cx->trace_ev[cx->trace_n++], cx->streams[stream]);
}
void zsk_cuda_trace_dump(zsk_cuda_ctx *cx)
{
    if (!cx->trace || !cx->trace_n) return;
    cudaDeviceSynchronize();
    for (unsigned i = 0; i < cx->trace_n; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, cx->trace_ev[0], cx->trace_ev[i]);
        fprintf(stderr, "[zsk trace] %8.3f ms  %s %u\n", ms, cx->trace_what[i], cx->trace_k[i]);
    }
}

const char *zsk_cuda_error(zsk_cuda_ctx *cx) { return cx ? cx->err : "no device context"; }
void zsk_cuda_set_user_stream(zsk_cuda_ctx *cx, void *stream) { cx->streams[ZSK_STREAM_USER] = (cudaStream_t)stream; }
int zsk_cuda_device(const zsk_cuda_ctx *cx) { return cx->device; }
int zsk_cuda_sm_count(const zsk_cuda_ctx *cx) { return cx->sm_count; }
unsigned long long zsk_cuda_launch_count(const zsk_cuda_ctx *cx) { return cx->launches; }

size_t zsk_cuda_free_memory(zsk_cuda_ctx *cx)
{
    size_t fr = 0, tot = 0;
    cudaSetDevice(cx->device);
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return 0;
    return fr;
}

/*
 * Device memory comes from the device's default stream-ordered pool (cudaMallocAsync) with the release threshold
 * lifted: a freed buffer goes back to the pool instead of the driver, and neither call synchronises the DEVICE the way
 * cudaFree does — with one reader per caller thread, sixteen readers growing their windows used to stop each other's
 * kernels hundreds of times per scan.  The allocation is made usable from every stream of the context by waiting for
 * the (otherwise idle) allocation stream; callers free a buffer only after the streams that used it were waited for.
 */
int zsk_cuda_malloc(zsk_cuda_ctx *cx, void **p, size_t n)
{
    const double t0 = zsk_dbg() ? zsk_now_ms() : 0;
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaMallocAsync(p, n ? n : 1, cx->alloc_stream));
    CK(cx, cudaStreamSynchronize(cx->alloc_stream));
    if (zsk_dbg() && zsk_now_ms() - t0 > 0.5) fprintf(stderr, "[zsk cx %p] device alloc %zu bytes: %.2f ms\n", (void *)cx, n, zsk_now_ms() - t0);
    return 0;
}

int zsk_cuda_free(zsk_cuda_ctx *cx, void *p)
{
    if (!p) return 0;
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaFreeAsync(p, cx->alloc_stream));
    return 0;
}

/*
 * Pinned host memory is expensive to create (page pinning under a process-wide driver lock: ~0.6-1.5 GB/s on the boxes
 * measured, and every other CUDA call of the process waits meanwhile), and readers come and go: buffers are rounded up to
 * a size class (powers of two and 1.5 x powers of two) and recycled through a process-wide free list; at most
 * ZSEEK_B200_PIN_CACHE_MB (default 6144) stay cached.  The class of a live buffer is kept in a side table, not in a header
 * in front of the bytes (a header pushed every power-of-two request into the next class: 64 MiB asked, 128 MiB pinned).
 */
#define ZSK_PIN_CLASSES 80
#define ZSK_PIN_LIVE 4096
struct zsk_pin_node { zsk_pin_node *next; };
static pthread_mutex_t g_pin_mu = PTHREAD_MUTEX_INITIALIZER;
static zsk_pin_node *g_pin_free[ZSK_PIN_CLASSES];
static size_t g_pin_cached, g_pin_cache_max;
static struct { void *p; int c; } g_pin_live[ZSK_PIN_LIVE];

static size_t pin_class_bytes(int c) { return ((size_t)2 + (size_t)(c & 1)) << (11 + c / 2); } /* 4K, 6K, 8K, 12K, 16K ... */

static int pin_class(size_t n)
{
    int c = 0;
    while (c < ZSK_PIN_CLASSES - 1 && pin_class_bytes(c) < n) c++;
    return c;
}

int zsk_cuda_malloc_host(zsk_cuda_ctx *cx, void **p, size_t n)
{
    const int c = pin_class(n ? n : 1);
    const size_t bytes = pin_class_bytes(c);
    if (bytes < n) {
        snprintf(cx->err, sizeof(cx->err), "pinned allocation too large");
        return -1;
    }
    pthread_mutex_lock(&g_pin_mu);
    zsk_pin_node *nd = g_pin_free[c];
    if (nd) {
        g_pin_free[c] = nd->next;
        g_pin_cached -= bytes;
    }
    pthread_mutex_unlock(&g_pin_mu);
    if (!nd) {
        const double t0 = zsk_dbg() ? zsk_now_ms() : 0;
        CK(cx, cudaSetDevice(cx->device));
        void *raw = NULL;
        CK(cx, cudaHostAlloc(&raw, bytes, cudaHostAllocPortable));
        nd = (zsk_pin_node *)raw;
        if (zsk_dbg()) fprintf(stderr, "[zsk cx %p] pinned alloc %zu bytes (asked %zu): %.2f ms\n", (void *)cx, bytes, n, zsk_now_ms() - t0);
    }
    pthread_mutex_lock(&g_pin_mu);
    int slot = -1;
    for (int i = 0; i < ZSK_PIN_LIVE; i++)
        if (!g_pin_live[i].p) { slot = i; break; }
    if (slot >= 0) { g_pin_live[slot].p = nd; g_pin_live[slot].c = c; }
    pthread_mutex_unlock(&g_pin_mu);
    if (slot < 0) { /* more live pinned buffers than the table holds: not recyclable, still usable */
        cudaFreeHost(nd);
        snprintf(cx->err, sizeof(cx->err), "too many pinned buffers");
        return -1;
    }
    *p = nd;
    return 0;
}

int zsk_cuda_free_host(zsk_cuda_ctx *cx, void *p)
{
    if (!p) return 0;
    zsk_pin_node *nd = (zsk_pin_node *)p;
    pthread_mutex_lock(&g_pin_mu);
    if (!g_pin_cache_max) {
        const char *g = getenv("ZSEEK_B200_PIN_CACHE_MB");
        g_pin_cache_max = (g && *g ? (size_t)strtoull(g, NULL, 10) : (size_t)6144) << 20;
        if (!g_pin_cache_max) g_pin_cache_max = 1;
    }
    int c = -1;
    for (int i = 0; i < ZSK_PIN_LIVE; i++)
        if (g_pin_live[i].p == p) { c = g_pin_live[i].c; g_pin_live[i].p = NULL; break; }
    const size_t bytes = c >= 0 ? pin_class_bytes(c) : 0;
    const bool keep = c >= 0 && g_pin_cached + bytes <= g_pin_cache_max;
    if (keep) {
        nd->next = g_pin_free[c];
        g_pin_free[c] = nd;
        g_pin_cached += bytes;
    }
    pthread_mutex_unlock(&g_pin_mu);
    if (!keep) CK(cx, cudaFreeHost(nd));
    return 0;
}

int zsk_cuda_memset_async(zsk_cuda_ctx *cx, void *p, int v, size_t n, int stream)
{
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaMemsetAsync(p, v, n, cx->streams[stream]));
    return 0;
}

int zsk_cuda_memcpy_async(zsk_cuda_ctx *cx, void *dst, const void *src, size_t n, int kind, int stream)
{
    if (n == 0) return 0;
    cudaMemcpyKind k = kind == ZSK_H2D ? cudaMemcpyHostToDevice : kind == ZSK_D2H ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaMemcpyAsync(dst, src, n, k, cx->streams[stream]));
    return 0;
}

int zsk_cuda_stream_sync(zsk_cuda_ctx *cx, int stream)
{
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaEventRecord(cx->block_ev[stream], cx->streams[stream]));
    CK(cx, cudaEventSynchronize(cx->block_ev[stream]));
    return 0;
}

/* spinning wait: for copies of a few KiB, where the ~30 us wake-up of a blocking wait would be most of the call */
int zsk_cuda_stream_sync_spin(zsk_cuda_ctx *cx, int stream)
{
    CK(cx, cudaStreamSynchronize(cx->streams[stream]));
    return 0;
}

int zsk_cuda_stream_wait(zsk_cuda_ctx *cx, int waiter, int signaler)
{
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaEventRecord(cx->sync_ev, cx->streams[signaler]));
    CK(cx, cudaStreamWaitEvent(cx->streams[waiter], cx->sync_ev, 0));
    return 0;
}

int zsk_cuda_event_record(zsk_cuda_ctx *cx, int ev, int stream)
{
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaEventRecord(cx->user_ev[ev], cx->streams[stream]));
    return 0;
}

int zsk_cuda_stream_wait_event(zsk_cuda_ctx *cx, int stream, int ev)
{
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaStreamWaitEvent(cx->streams[stream], cx->user_ev[ev], 0));
    return 0;
}

int zsk_cuda_event_sync(zsk_cuda_ctx *cx, int ev)
{
    CK(cx, cudaEventSynchronize(cx->user_ev[ev]));
    return 0;
}

int zsk_cuda_pointer_is_device(zsk_cuda_ctx *cx, const void *p)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError(); /* unregistered host memory reports an error on old runtimes */
        return 0;
    }
    (void)cx;
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? 1 : 0;
}

static int next_counter(zsk_cuda_ctx *cx, int stream, uint32_t **out)
{
    uint32_t *c = cx->counters + (cx->counter_next++ % ZSK_NCOUNTERS);
    CK(cx, cudaMemsetAsync(c, 0, sizeof(uint32_t), cx->streams[stream]));
    *out = c;
    return 0;
}

/* (re)allocates *p to hold `need` bytes; growth is geometric */
static int grow_pool(zsk_cuda_ctx *cx, void **p, size_t *cap_bytes, size_t need)
{
    if (need <= *cap_bytes) return 0;
    const double t0 = zsk_dbg() ? zsk_now_ms() : 0;
    size_t want = need > 2 * *cap_bytes ? need : 2 * *cap_bytes;
    /* the pools are shared by every launch of this context: the launches in flight must be through with them */
    for (int i = 0; i < ZSK_NSTREAMS; i++) CK(cx, cudaStreamSynchronize(cx->streams[i]));
    if (cx->streams[ZSK_STREAM_USER]) CK(cx, cudaStreamSynchronize(cx->streams[ZSK_STREAM_USER]));
    if (*p) CK(cx, cudaFreeAsync(*p, cx->alloc_stream));
    *p = NULL;
    *cap_bytes = 0;
    if (cudaMallocAsync(p, want, cx->alloc_stream) != cudaSuccess) {
        cudaGetLastError();
        want = need;
        CK(cx, cudaMallocAsync(p, want, cx->alloc_stream));
    }
    CK(cx, cudaStreamSynchronize(cx->alloc_stream));
    *cap_bytes = want;
    if (zsk_dbg()) fprintf(stderr, "[zsk cx %p] scratch pool grown to %zu bytes: %.2f ms\n", (void *)cx, want, zsk_now_ms() - t0);
    return 0;
}

/* literal scratch of the one-CTA-per-frame zstd kernel for launches of up to `ctas` CTAs.  A context keeps what it has; it
 * grows only when a parked context that ran the pipeline (one wave for deferred frames) is taken over by a reader opened with
 * ZSEEK_B200_ZSTD_LEGACY=1, whose launches use more CTAs. */
static int ensure_lit_scratch(zsk_cuda_ctx *cx, size_t ctas)
{
    if (cx->scratch && cx->scratch_ctas >= ctas) return 0;
    if (cx->scratch) {
        CK(cx, cudaDeviceSynchronize());
        CK(cx, cudaFreeAsync(cx->scratch, cx->alloc_stream));
        cx->scratch = NULL;
        cx->scratch_ctas = 0;
    }
    CK(cx, cudaMallocAsync((void **)&cx->scratch, ctas * ZSK_LIT_SCRATCH + ZSK_PAD_BACK, cx->alloc_stream));
    CK(cx, cudaStreamSynchronize(cx->alloc_stream));
    cx->scratch_ctas = ctas;
    return 0;
}

/* zstd: index -> FSE sequences + Huffman literals -> execution -> (deferred frames) one-CTA-per-frame kernel */
static int launch_zstd_pipeline(zsk_cuda_ctx *cx, zsk_decode_args a, cudaStream_t s, int stream)
{
    const size_t njobs = a.njobs;
    const uint64_t dsum = a.dsize_sum ? a.dsize_sum : (uint64_t)njobs << 20;
    int rc;
    if ((rc = ensure_lit_scratch(cx, (size_t)cx->sm_count))) return rc; /* deferred frames are rare: one wave of the old kernel is plenty */
    if (njobs > cx->zjobs_cap) {
        size_t cap_f = cx->zjobs_cap * sizeof(zsk_zframe), cap_d = cx->zjobs_cap * sizeof(uint32_t);
        if ((rc = grow_pool(cx, (void **)&cx->zframes, &cap_f, njobs * sizeof(zsk_zframe)))) return rc;
        if ((rc = grow_pool(cx, (void **)&cx->zdeferred, &cap_d, njobs * sizeof(uint32_t)))) return rc;
        cx->zjobs_cap = cap_f / sizeof(zsk_zframe) < cap_d / sizeof(uint32_t) ? cap_f / sizeof(zsk_zframe) : cap_d / sizeof(uint32_t);
    }
    /* pool sizes for typical data: one block per 2 KiB of output, one sequence (12 bytes of records) per 5 bytes, 3/4 of the output as
     * Huffman literals; frames beyond that are deferred to the old kernel by P0 */
    const size_t want_blocks = (size_t)(dsum / 2048) + 4 * njobs + 64;
    const size_t want_seqs = (size_t)(dsum / 5) + 64 * njobs + 8 * want_blocks + 1024; /* a block's arrays are padded to 8 sequences */
    const size_t want_lits = (size_t)(dsum / 4 * 3) + 64 * njobs + 4096;
    size_t cap_b = cx->zblocks_cap * sizeof(zsk_zblock), cap_s = cx->zseqs_cap * 3 * sizeof(uint32_t), cap_l = cx->zlits_cap ? cx->zlits_cap + ZSK_PAD_BACK : 0;
    if ((rc = grow_pool(cx, (void **)&cx->zblocks, &cap_b, want_blocks * sizeof(zsk_zblock)))) return rc;
    if ((rc = grow_pool(cx, (void **)&cx->zseqs, &cap_s, want_seqs * 3 * sizeof(uint32_t)))) return rc;
    if ((rc = grow_pool(cx, (void **)&cx->zlits, &cap_l, want_lits + ZSK_PAD_BACK))) return rc;
    cx->zblocks_cap = cap_b / sizeof(zsk_zblock);
    if (a.limits) {
        size_t cap_p = cx->zbprog_cap * 2 * sizeof(uint32_t);
        if ((rc = grow_pool(cx, (void **)&cx->zbprog, &cap_p, cx->zblocks_cap * 2 * sizeof(uint32_t)))) return rc;
        cx->zbprog_cap = cap_p / (2 * sizeof(uint32_t));
    }
    cx->zseqs_cap = cap_s / (3 * sizeof(uint32_t));
    cx->zlits_cap = cap_l - ZSK_PAD_BACK;

    zsk_zpipe_args z;
    z.a = a;
    z.frames = cx->zframes; z.blocks = cx->zblocks; z.seqs = cx->zseqs; z.lits = cx->zlits;
    z.blocks_cap = cx->zblocks_cap; z.seqs_cap = cx->zseqs_cap; z.lits_cap = cx->zlits_cap;
    z.ctr = cx->zctr + (size_t)(cx->zctr_next++ % ZSK_NCOUNTERS) * ZSK_ZC_N;
    z.deferred = cx->zdeferred;
    z.bprog = cx->zbprog;
    CK(cx, cudaMemsetAsync(z.ctr, 0, ZSK_ZC_N * sizeof(unsigned long long), s));
    /* grids follow the work of the launch (an estimate: one block per 32 KiB of output), so that the small launches of
     * several readers run side by side instead of each filling the GPU with CTAs that find nothing to do */
    const size_t est_blocks = (size_t)(dsum >> 15) + njobs;
    const size_t fse_units = a.limits ? njobs : est_blocks; /* with limits a trio owns a frame, else a block */
    unsigned fse_ctas = (unsigned)((fse_units + ZSK_ZFSE_WARPS * ZSK_ZFSE_TRIOS - 1) / (ZSK_ZFSE_WARPS * ZSK_ZFSE_TRIOS));
    if (fse_ctas > (unsigned)cx->zfse_ctas) fse_ctas = (unsigned)cx->zfse_ctas;
    unsigned huf_ctas = (unsigned)((est_blocks + ZSK_ZHUF_SLOTS - 1) / ZSK_ZHUF_SLOTS);
    if (huf_ctas > (unsigned)cx->zhuf_ctas) huf_ctas = (unsigned)cx->zhuf_ctas;
    zsk_zstd_index_kernel<<<(unsigned)((njobs + 127) / 128), 128, 0, s>>>(z);
    zsk_zstd_fse_kernel<<<fse_ctas, ZSK_ZFSE_THREADS, ZSK_ZFSE_SMEM, s>>>(z);
    zsk_zstd_huf_kernel<<<huf_ctas, 32, ZSK_ZHUF_SMEM, s>>>(z);
    unsigned xctas = njobs < (size_t)cx->zexec_ctas ? (unsigned)njobs : (unsigned)cx->zexec_ctas;
    zsk_zstd_exec_kernel<<<xctas, 32, ZSK_ZX_SMEM, s>>>(z);
    /* frames P0 / P1a handed over (normally none: the CTAs find an empty list and leave) */
    zsk_decode_args d = a;
    d.job_list = z.deferred;
    d.job_list_count = z.ctr + ZSK_ZC_DEFERRED;
    d.scratch = cx->scratch;
    if ((rc = next_counter(cx, stream, &d.work_counter))) return rc;
    unsigned dctas = njobs < (size_t)cx->sm_count ? (unsigned)njobs : (unsigned)cx->sm_count;
    zsk_zstd_decode_kernel<<<dctas, ZSK_ZSTD_CTA_THREADS, 0, s>>>(d);
    cx->launches += 4;
    return 0;
}

int zsk_cuda_launch_decode(zsk_cuda_ctx *cx, int codec, const zsk_decode_args *args, int stream)
{
    if (args->njobs == 0) return 0;
    CK(cx, cudaSetDevice(cx->device));
    zsk_decode_args a = *args;
    int rc = next_counter(cx, stream, &a.work_counter);
    if (rc) return rc;
    a.job_list = NULL;
    a.job_list_count = NULL;
    if (!a.njobs_dev) a.job_base = 0;
    cudaStream_t s = cx->streams[stream];
    CK(cx, cudaEventRecord(cx->k0, s));
    if (codec == ZSK_CODEC_LZ4 && a.njobs >= cx->lz4_lane_min) {
        unsigned ctas = (a.njobs + ZSK_LZ4L_THREADS - 1) / ZSK_LZ4L_THREADS;
        if (ctas > (unsigned)cx->lz4_lane_ctas) ctas = (unsigned)cx->lz4_lane_ctas;
        zsk_lz4_decode_lane_kernel<<<ctas, ZSK_LZ4L_THREADS, ZSK_LZ4L_SMEM, s>>>(a);
        cx->k_name = "zsk_lz4_decode_lane_kernel";
    } else if (codec == ZSK_CODEC_LZ4) {
        const unsigned frames_per_cta = ZSK_LZ4_CTA_THREADS / 32;
        unsigned ctas = (a.njobs + frames_per_cta - 1) / frames_per_cta;
        if (ctas > (unsigned)cx->lz4_ctas) ctas = (unsigned)cx->lz4_ctas;
        zsk_lz4_decode_batch_kernel<<<ctas, ZSK_LZ4_CTA_THREADS, 0, s>>>(a);
        cx->k_name = "zsk_lz4_decode_batch_kernel";
    } else if (codec == ZSK_CODEC_ZSTD && cx->zstd_legacy) {
        if (int rc = ensure_lit_scratch(cx, (size_t)cx->zstd_ctas)) return rc;
        a.scratch = cx->scratch;
        unsigned ctas = a.njobs < (unsigned)cx->zstd_ctas ? a.njobs : (unsigned)cx->zstd_ctas;
        zsk_zstd_decode_kernel<<<ctas, ZSK_ZSTD_CTA_THREADS, 0, s>>>(a);
        cx->k_name = "zsk_zstd_decode_kernel";
    } else if (codec == ZSK_CODEC_ZSTD) {
        /* waves of at most zwave_bytes of output share one set of scratch pools (2.75 bytes of scratch per output byte) */
        const uint64_t dsum = a.dsize_sum ? a.dsize_sum : (uint64_t)a.njobs << 20;
        const uint32_t total = a.njobs;
        uint32_t per_wave = total;
        if (dsum > cx->zwave_bytes) per_wave = (uint32_t)((double)total * (double)cx->zwave_bytes / (double)dsum) + 1u;
        for (uint32_t j0 = 0; j0 < total; j0 += per_wave) {
            zsk_decode_args w = a;
            w.njobs = total - j0 < per_wave ? total - j0 : per_wave;
            w.first_frame = a.first_frame + j0;
            if (a.frame_ids) w.frame_ids = a.frame_ids + j0;
            if (a.dst_offs) w.dst_offs = a.dst_offs + j0;
            if (a.limits) w.limits = a.limits + j0;
            w.status = a.status + j0;
            w.job_base = a.job_base + j0;
            w.dsize_sum = (uint64_t)((double)dsum * (double)w.njobs / (double)total) + 1u;
            rc = launch_zstd_pipeline(cx, w, s, stream);
            if (rc) return rc;
        }
        cx->k_name = "zsk_zstd_exec_kernel";
    } else {
        snprintf(cx->err, sizeof(cx->err), "unknown codec %d", codec);
        return -1;
    }
    CK(cx, cudaGetLastError());
    CK(cx, cudaEventRecord(cx->k1, s));
    cx->k_valid = 1;
    cx->launches++;
    return 0;
}

static unsigned grid_for(zsk_cuda_ctx *cx, uint64_t threads_wanted, unsigned block)
{
    uint64_t ctas = (threads_wanted + block - 1) / block;
    uint64_t cap = (uint64_t)cx->sm_count * 8;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (unsigned)ctas;
}

int zsk_cuda_launch_lookup(zsk_cuda_ctx *cx, const zsk_lookup_args *args, int stream)
{
    if (args->n == 0) return 0;
    CK(cx, cudaSetDevice(cx->device));
    zsk_lookup_kernel<<<grid_for(cx, args->n, 256), 256, 0, cx->streams[stream]>>>(*args);
    CK(cx, cudaGetLastError());
    cx->launches++;
    return 0;
}

int zsk_cuda_launch_gather(zsk_cuda_ctx *cx, const zsk_gather_args *args, int stream)
{
    if (args->n == 0) return 0;
    CK(cx, cudaSetDevice(cx->device));
    zsk_gather_kernel<<<grid_for(cx, (uint64_t)args->n * 32, 256), 256, 0, cx->streams[stream]>>>(*args);
    CK(cx, cudaGetLastError());
    cx->launches++;
    return 0;
}

int zsk_cuda_launch_compact(zsk_cuda_ctx *cx, const zsk_compact_args *args, int stream)
{
    if (args->nframes == 0) return 0;
    CK(cx, cudaSetDevice(cx->device));
    zsk_compact_kernel<<<grid_for(cx, args->nframes, 256), 256, 0, cx->streams[stream]>>>(*args);
    CK(cx, cudaGetLastError());
    cx->launches++;
    return 0;
}

int zsk_cuda_timer_start(zsk_cuda_ctx *cx, int stream)
{
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaEventRecord(cx->t0, cx->streams[stream]));
    return 0;
}

int zsk_cuda_timer_stop(zsk_cuda_ctx *cx, int stream, float *ms)
{
    CK(cx, cudaSetDevice(cx->device));
    CK(cx, cudaEventRecord(cx->t1, cx->streams[stream]));
    CK(cx, cudaEventSynchronize(cx->t1));
    CK(cx, cudaEventElapsedTime(ms, cx->t0, cx->t1));
    return 0;
}

const char *zsk_cuda_last_decode_kernel(const zsk_cuda_ctx *cx) { return cx->k_name ? cx->k_name : ""; }

int zsk_cuda_last_decode_ms(zsk_cuda_ctx *cx, float *ms)
{
    if (!cx->k_valid) {
        snprintf(cx->err, sizeof(cx->err), "no decode kernel launched yet");
        return -1;
    }
    CK(cx, cudaEventSynchronize(cx->k1));
    CK(cx, cudaEventElapsedTime(ms, cx->k0, cx->k1));
    return 0;
}

} /* extern "C" */
