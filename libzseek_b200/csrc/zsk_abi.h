/*
 * zsk_abi.h — plain-C types shared by the host reader (reader.c), the C-ABI launch layer
 * (zsk_cuda.cu) and the kernels (zsk_*.cuh).  No CUDA or C++ types appear here.
 */
#ifndef ZSK_ABI_H
#define ZSK_ABI_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ZSK_CODEC_ZSTD = 0, ZSK_CODEC_LZ4 = 1 }; /* same numbering as zseek_compression_type_t, reference src/zseek.h:121-124 */

/* Per-frame status words written by the decode kernels (0 = ok).  A corrupt or truncated frame makes
 * zseek_pread return -1 instead of hanging. */
enum zsk_status {
    ZSK_ST_OK = 0,
    ZSK_ST_TRUNC = 1,       /* compressed frame ends early */
    ZSK_ST_MAGIC = 2,       /* frame does not start with the codec magic */
    ZSK_ST_FORMAT = 3,      /* reserved bits set / impossible field */
    ZSK_ST_DST = 4,         /* frame decodes to more bytes than the seek table's dSize */
    ZSK_ST_OFFSET = 5,      /* match offset reaches before the frame start */
    ZSK_ST_BITSTREAM = 6,   /* backward bitstream over/under-run */
    ZSK_ST_TABLE = 7,       /* bad FSE / Huffman description */
    ZSK_ST_UNSUPPORTED = 8, /* dictionary id */
    ZSK_ST_SIZE = 9,        /* frame decodes to fewer bytes than the seek table's dSize */
    ZSK_ST_CHECKSUM = 10,   /* header / block / content checksum of the frame does not match */
    ZSK_ST_STOPPED = 100,   /* kernel-internal: the job's limit was reached (reported as ZSK_ST_OK) */
    ZSK_ST_DEFERRED = 101   /* kernel-internal: the zstd pipeline handed the frame to the one-CTA-per-frame kernel that runs
                             * last in the same launch (scratch pools too small for it, or an offset beyond 2^28) */
};

/* Device buffers handed to the kernels must be readable ZSK_PAD_FRONT bytes before and ZSK_PAD_BACK
 * bytes after their logical extent (aligned-word reads of unaligned byte streams). */
#define ZSK_PAD_FRONT 16
#define ZSK_PAD_BACK 64

/* zstd: per-CTA literal scratch in HBM (a block regenerates <= 128 KiB of literals) */
#define ZSK_LIT_SCRATCH (128u * 1024u + 64u)

/* One decode launch works through `njobs` frames.  Job i is frame
 *     f = frame_ids ? frame_ids[i] : first_frame + i
 * whose compressed bytes are comp[c_off[f] - comp_base .. c_off[f+1] - comp_base) and whose dSize
 * bytes go to  dst + (dst_offs ? dst_offs[i] : d_off[f] - dst_base).  All pointers are device pointers. */
typedef struct zsk_decode_args {
    const uint64_t *c_off;      /* [N+1] compressed prefix offsets */
    const uint64_t *d_off;      /* [N+1] decompressed prefix offsets */
    const uint8_t *comp;        /* device-resident compressed image (slice) */
    uint64_t comp_base;         /* file offset of comp[0] */
    const uint32_t *frame_ids;  /* optional [njobs] */
    const uint64_t *dst_offs;   /* optional [njobs] */
    uint8_t *dst;
    uint64_t dst_base;          /* decompressed offset that maps to dst[0] when dst_offs == NULL */
    uint32_t first_frame;
    uint32_t njobs;
    int32_t *status;            /* [njobs] */
    uint32_t *work_counter;     /* zero-initialised; dynamic job distribution (filled in by the launch layer) */
    uint8_t *scratch;           /* zstd literal scratch, ZSK_LIT_SCRATCH bytes per CTA (filled in by the launch layer) */
    const uint32_t *limits;     /* optional [njobs]: job i may stop once limits[i] bytes of its frame are decoded (a batch of
                                 * small reads needs only the prefix of a frame up to its last requested byte, like the
                                 * reference's streaming no-cache path, src/decompress.c:419-454); NULL = whole frames */
    uint64_t dsize_sum;         /* sum of the jobs' decompressed sizes (sizes the zstd pipeline's scratch pools); 0 = unknown */
    const uint32_t *job_list;   /* launch-layer internal: when set, only jobs job_list[0 .. *job_list_count) are run */
    const unsigned long long *job_list_count;
    const uint32_t *njobs_dev;  /* optional: the job count lives in device memory (written by an earlier kernel of the same
                                 * stream); njobs is then an upper bound that sizes grids and scratch, and the launch runs
                                 * jobs [0, min(njobs, *njobs_dev - job_base)) */
    uint32_t job_base;
} zsk_decode_args;

/* K1: batched offset -> frame lookup (semantics of reference src/seek_table.c:187-202 + decompress.c:445) */
typedef struct zsk_lookup_args {
    const uint64_t *d_off;   /* [N+1] */
    uint32_t nframes;
    const uint64_t *offsets; /* [n] */
    const uint64_t *counts;  /* [n], or NULL with fixed_count */
    uint64_t fixed_count;
    uint32_t n;
    int32_t *frame;          /* [n] out: frame index or -1 (EOF) */
    uint32_t *inframe;       /* [n] out */
    uint32_t *nbytes;        /* [n] out: MIN(count, frame_end - offset) */
    uint32_t *touched;       /* optional [N], zero-initialised: for every frame some request needs, the largest in-frame end
                              * offset (inframe + nbytes, at least 1) over the batch */
} zsk_lookup_args;

/* K4: per-request range copy out of decoded frames */
typedef struct zsk_gather_args {
    const int32_t *frame;        /* [n] from K1 */
    const uint32_t *inframe;     /* [n] */
    const uint32_t *nbytes;      /* [n] */
    const int64_t *frame_src;    /* [N] byte offset of each decoded frame inside src_base, -1 = absent */
    const uint8_t *src_base;     /* decoded-frame store (HBM cache slab or a decoded range) */
    uint8_t *dst;                /* request i lands at dst + (dst_offs ? dst_offs[i] : i * dst_stride) */
    const uint64_t *dst_offs;
    uint64_t dst_stride;
    uint32_t n;
    int64_t *results;            /* optional [n]: nbytes of every request, as zseek_pread would return it */
} zsk_gather_args;

/* Turns the per-frame marks of K1 into the job list of a stream-ordered batch: every touched frame gets a job (frame id,
 * limit = how much of it the batch needs, a slot of the batch slab) and its slot offset in frame_src; untouched frames
 * get frame_src = -1. */
typedef struct zsk_compact_args {
    const uint32_t *touched;   /* [N] from K1 */
    uint32_t nframes;
    uint32_t shard_lo, shard_hi; /* frames outside [shard_lo, shard_hi) set *error */
    uint64_t slot_size;
    uint32_t max_jobs;         /* capacity of the job arrays / slab slots */
    uint32_t *job_ids;         /* [max_jobs] out */
    uint64_t *job_offs;        /* [max_jobs] out: slot offset inside the slab */
    uint32_t *job_limits;      /* [max_jobs] out */
    int64_t *frame_src;        /* [N] out */
    uint32_t *count;           /* zero-initialised; out: number of jobs */
    uint32_t *error;           /* zero-initialised; out: 1 = frame outside the shard, 2 = more touched frames than max_jobs */
} zsk_compact_args;

#ifdef __cplusplus
}
#endif
#endif
