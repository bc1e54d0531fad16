/* hostemu.cpp — TEST INFRASTRUCTURE ONLY: a stand-in for the C-ABI launch layer (libzseek_b200/csrc/zsk_cuda.h) that keeps
 * "device" memory in host memory and runs the product's kernel sources on the lock-step emulator (cuda_emu.h).
 *
 * Linked with the UNMODIFIED libzseek_b200/csrc/reader.c it gives tests/emu/libzsk_hostemu.so, which exports the same
 * zseek_* symbols as the product.  Its only purpose is to exercise the HOST logic of reader.c — seek table, LRU cache,
 * read-ahead windows, batch bookkeeping, residency, parking, error paths — against the oracle in a container without a
 * GPU (tests/test_host_emulation.py).  Every queued operation runs at once, in issue order: one legal schedule of the
 * stream semantics, so functional logic is checked, missing synchronisation is not.
 *
 * It is NOT a product path and NOT a CPU fallback: nothing under libzseek_b200/ or include/ names it, the shipped library
 * is never built from it, and it is orders of magnitude slower than any decoder (every CUDA thread is a fiber). */
#include <algorithm>
#include <map>
#include <mutex>
#include <string>

#include "emu_kernels.cpp" /* after the standard headers: cuda_emu.h defines __noinline__ and friends as macros */
#include "../../libzseek_b200/csrc/zsk_cuda.h"

struct zsk_cuda_ctx {
    char err[256];
    unsigned long long launches;
    const char *k_name;
    size_t lane_min;
    bool legacy;
};

static std::recursive_mutex g_mu;                 /* the emulator has one global CTA: launches are serialised */
struct DevAlloc { size_t span, bytes; };
static std::map<uintptr_t, DevAlloc> g_dev;       /* "device" allocations: base -> extent */
static size_t g_dev_bytes, g_dev_peak, g_dev_allocs;

static size_t env_num(const char *name, size_t dflt)
{
    const char *s = getenv(name);
    return s && *s ? (size_t)strtoull(s, nullptr, 10) : dflt;
}

static void configure(zsk_cuda_ctx *cx)
{
    cx->lane_min = env_num("ZSEEK_B200_LZ4_LANE_MIN", 40960);
    cx->legacy = env_num("ZSEEK_B200_ZSTD_LEGACY", 0) != 0;
}

extern "C" {

int zsk_cuda_ctx_create(int device, zsk_cuda_ctx **out, char *err, size_t errlen)
{
    (void)device;
    if (getenv("ZSK_HOSTEMU_NO_DEVICE")) {
        snprintf(err, errlen, "no CUDA device: emulated absence");
        return -1;
    }
    zsk_cuda_ctx *cx = (zsk_cuda_ctx *)calloc(1, sizeof(*cx));
    if (!cx) return -1;
    configure(cx);
    *out = cx;
    return 0;
}
void zsk_cuda_ctx_destroy(zsk_cuda_ctx *cx) { free(cx); }
int zsk_cuda_pick_device(void) { return 0; }
int zsk_cuda_ctx_reuse(zsk_cuda_ctx *cx)
{
    cx->launches = 0;
    cx->k_name = nullptr;
    configure(cx);
    return 0;
}
void zsk_cuda_ctx_trim(zsk_cuda_ctx *, size_t) {}
size_t zsk_cuda_ctx_held(const zsk_cuda_ctx *) { return 0; }
const char *zsk_cuda_error(zsk_cuda_ctx *cx) { return cx ? cx->err : "no device context"; }
int zsk_cuda_device(const zsk_cuda_ctx *) { return 0; }
int zsk_cuda_sm_count(const zsk_cuda_ctx *) { return 2; }
unsigned long long zsk_cuda_launch_count(const zsk_cuda_ctx *cx) { return cx->launches; }
size_t zsk_cuda_free_memory(zsk_cuda_ctx *) { return (size_t)64 << 30; }

int zsk_cuda_malloc(zsk_cuda_ctx *cx, void **p, size_t n)
{
    std::lock_guard<std::recursive_mutex> g(g_mu);
    const size_t cap = env_num("ZSK_HOSTEMU_DEVICE_MB", 4096) << 20;
    if (g_dev_bytes + n > cap) {
        snprintf(cx->err, sizeof(cx->err), "out of (emulated) device memory");
        *p = nullptr;
        return 2;
    }
    uint8_t *q = (uint8_t *)malloc(n ? n : 1);
    if (!q) return 2;
    memset(q, 0xA5, n); /* cudaMalloc does not clear memory either */
    g_dev[(uintptr_t)q] = DevAlloc{n ? n : 1, n};
    g_dev_bytes += n;
    g_dev_allocs++;
    if (g_dev_bytes > g_dev_peak) g_dev_peak = g_dev_bytes;
    *p = q;
    return 0;
}
int zsk_cuda_free(zsk_cuda_ctx *, void *p)
{
    if (!p) return 0;
    std::lock_guard<std::recursive_mutex> g(g_mu);
    auto it = g_dev.find((uintptr_t)p);
    if (it == g_dev.end()) {
        fprintf(stderr, "hostemu: free of a pointer that is not a live device allocation\n");
        abort();
    }
    g_dev_bytes -= it->second.bytes;
    g_dev.erase(it);
    free(p);
    return 0;
}
int zsk_cuda_malloc_host(zsk_cuda_ctx *, void **p, size_t n)
{
    *p = malloc(n ? n : 1);
    if (*p) memset(*p, 0x5C, n);
    return *p ? 0 : 2;
}
int zsk_cuda_free_host(zsk_cuda_ctx *, void *p)
{
    free(p);
    return 0;
}
int zsk_cuda_memset_async(zsk_cuda_ctx *, void *p, int v, size_t n, int) { memset(p, v, n); return 0; }
int zsk_cuda_memcpy_async(zsk_cuda_ctx *, void *dst, const void *src, size_t n, int, int) { memmove(dst, src, n); return 0; }
int zsk_cuda_stream_sync(zsk_cuda_ctx *, int) { return 0; }
int zsk_cuda_stream_sync_spin(zsk_cuda_ctx *, int) { return 0; }
void zsk_cuda_set_user_stream(zsk_cuda_ctx *, void *) {}
int zsk_cuda_stream_wait(zsk_cuda_ctx *, int, int) { return 0; }
int zsk_cuda_event_record(zsk_cuda_ctx *, int, int) { return 0; }
int zsk_cuda_event_sync(zsk_cuda_ctx *, int) { return 0; }
int zsk_cuda_stream_wait_event(zsk_cuda_ctx *, int, int) { return 0; }

int zsk_cuda_pointer_is_device(zsk_cuda_ctx *, const void *p)
{
    std::lock_guard<std::recursive_mutex> g(g_mu);
    auto it = g_dev.upper_bound((uintptr_t)p);
    if (it == g_dev.begin()) return 0;
    --it;
    return (uintptr_t)p < it->first + it->second.span ? 1 : 0;
}

int zsk_cuda_launch_decode(zsk_cuda_ctx *cx, int codec, const zsk_decode_args *args, int)
{
    if (args->njobs == 0) return 0;
    std::lock_guard<std::recursive_mutex> g(g_mu);
    const uint32_t ctas = 2;
    uint32_t counter = 0;
    std::vector<uint8_t> scratch((size_t)ctas * ZSK_LIT_SCRATCH + 64);
    zsk_decode_args a = *args;
    a.work_counter = &counter;
    a.scratch = scratch.data();
    a.job_list = nullptr;
    a.job_list_count = nullptr;
    if (!a.njobs_dev) a.job_base = 0;
    if (codec == ZSK_CODEC_LZ4 && a.njobs >= cx->lane_min) {
        emu::launch(dim3(ctas), dim3(ZSK_LZ4L_THREADS), ZSK_LZ4L_SMEM, [&] { zsk_lz4_decode_lane_kernel(a); });
        cx->k_name = "zsk_lz4_decode_lane_kernel";
    } else if (codec == ZSK_CODEC_LZ4) {
        emu::launch(dim3(ctas), dim3(ZSK_LZ4_CTA_THREADS), 0, [&] { zsk_lz4_decode_batch_kernel(a); });
        cx->k_name = "zsk_lz4_decode_batch_kernel";
    } else if (codec == ZSK_CODEC_ZSTD && cx->legacy) {
        emu::launch(dim3(ctas), dim3(ZSK_ZSTD_CTA_THREADS), 0, [&] { zsk_zstd_decode_kernel(a); });
        cx->k_name = "zsk_zstd_decode_kernel";
    } else if (codec == ZSK_CODEC_ZSTD) {
        /* pools are sized from the frames the launch may run: with a device-resident job count, njobs is an upper bound */
        uint32_t live = a.njobs;
        if (a.njobs_dev) live = *a.njobs_dev > a.job_base ? std::min(a.njobs, *a.njobs_dev - a.job_base) : 0u;
        std::vector<uint32_t> ids(a.njobs);
        for (uint32_t j = 0; j < a.njobs; j++) {
            const uint32_t jj = j < live ? j : 0;
            ids[j] = a.frame_ids ? (live ? a.frame_ids[jj] : 0u) : a.first_frame + jj;
        }
        emu_zstd_pipeline(a, a.njobs, ctas, a.d_off, a.first_frame, ids.data(), false);
        cx->k_name = "zsk_zstd_exec_kernel";
    } else {
        snprintf(cx->err, sizeof(cx->err), "unknown codec %d", codec);
        return -1;
    }
    cx->launches++;
    return 0;
}
int zsk_cuda_launch_lookup(zsk_cuda_ctx *cx, const zsk_lookup_args *args, int)
{
    if (args->n == 0) return 0;
    std::lock_guard<std::recursive_mutex> g(g_mu);
    zsk_lookup_args a = *args;
    emu::launch(dim3(2), dim3(256), 0, [&] { zsk_lookup_kernel(a); });
    cx->launches++;
    return 0;
}
int zsk_cuda_launch_gather(zsk_cuda_ctx *cx, const zsk_gather_args *args, int)
{
    if (args->n == 0) return 0;
    std::lock_guard<std::recursive_mutex> g(g_mu);
    zsk_gather_args a = *args;
    emu::launch(dim3(2), dim3(256), 0, [&] { zsk_gather_kernel(a); });
    cx->launches++;
    return 0;
}
int zsk_cuda_launch_compact(zsk_cuda_ctx *cx, const zsk_compact_args *args, int)
{
    if (args->nframes == 0) return 0;
    std::lock_guard<std::recursive_mutex> g(g_mu);
    zsk_compact_args a = *args;
    emu::launch(dim3(2), dim3(256), 0, [&] { zsk_compact_kernel(a); });
    cx->launches++;
    return 0;
}

int zsk_cuda_timer_start(zsk_cuda_ctx *, int) { return 0; }
int zsk_cuda_timer_stop(zsk_cuda_ctx *, int, float *ms) { *ms = 0.0f; return 0; }
int zsk_cuda_last_decode_ms(zsk_cuda_ctx *cx, float *ms)
{
    if (!cx->k_name) return -1;
    *ms = 0.0f;
    return 0;
}
const char *zsk_cuda_last_decode_kernel(const zsk_cuda_ctx *cx) { return cx->k_name ? cx->k_name : ""; }
int zsk_cuda_trace_enabled(const zsk_cuda_ctx *) { return 0; }
void zsk_cuda_trace_reset(zsk_cuda_ctx *) {}
void zsk_cuda_trace_mark(zsk_cuda_ctx *, int, const char *, unsigned) {}
void zsk_cuda_trace_dump(zsk_cuda_ctx *) {}

/* test hooks: "device" buffers for destinations, and allocation statistics */
__attribute__((visibility("default"))) void *hostemu_device_alloc(size_t n)
{
    zsk_cuda_ctx tmp{};
    void *p = nullptr;
    return zsk_cuda_malloc(&tmp, &p, n) ? nullptr : p;
}
__attribute__((visibility("default"))) void hostemu_device_free(void *p) { zsk_cuda_free(nullptr, p); }
__attribute__((visibility("default"))) size_t hostemu_device_bytes(void) { return g_dev_bytes; }
__attribute__((visibility("default"))) size_t hostemu_device_allocs(void) { return g_dev_allocs; }
}
