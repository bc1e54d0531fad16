/* TEST INFRASTRUCTURE ONLY — declaration-only ABI shim (no code).
 *
 * The reference (libzseek, /root/reference/src) includes <zstd.h>, but this image ships only the
 * libzstd *runtime* (libzstd.so.1, v1.5.5) and no headers.  This file declares exactly the stable
 * libzstd ABI the reference sources use so that oracle/Makefile can compile
 * the untouched reference sources into oracle/_ref/libzseek_ref.so.  Nothing in the product
 * (libzseek_b200/) includes it.
 */
#ifndef ZSK_SHIM_ZSTD_H
#define ZSK_SHIM_ZSTD_H
#include <stddef.h>

#define ZSTD_MAGIC_SKIPPABLE_START 0x184D2A50
#define ZSTD_CLEVEL_DEFAULT 3

typedef struct ZSTD_CCtx_s ZSTD_CCtx;
typedef struct ZSTD_DCtx_s ZSTD_DCtx;
typedef ZSTD_DCtx ZSTD_DStream;

typedef struct ZSTD_inBuffer_s  { const void *src; size_t size; size_t pos; } ZSTD_inBuffer;
typedef struct ZSTD_outBuffer_s { void *dst;       size_t size; size_t pos; } ZSTD_outBuffer;

typedef enum {
    ZSTD_fast = 1, ZSTD_dfast = 2, ZSTD_greedy = 3, ZSTD_lazy = 4, ZSTD_lazy2 = 5,
    ZSTD_btlazy2 = 6, ZSTD_btopt = 7, ZSTD_btultra = 8, ZSTD_btultra2 = 9
} ZSTD_strategy;

typedef enum {
    ZSTD_c_compressionLevel = 100, ZSTD_c_windowLog = 101, ZSTD_c_strategy = 107,
    ZSTD_c_contentSizeFlag = 200, ZSTD_c_checksumFlag = 201, ZSTD_c_nbWorkers = 400
} ZSTD_cParameter;

typedef enum { ZSTD_e_continue = 0, ZSTD_e_flush = 1, ZSTD_e_end = 2 } ZSTD_EndDirective;

unsigned    ZSTD_isError(size_t code);
const char *ZSTD_getErrorName(size_t code);

ZSTD_CCtx *ZSTD_createCCtx(void);
size_t     ZSTD_freeCCtx(ZSTD_CCtx *cctx);
size_t     ZSTD_sizeof_CCtx(const ZSTD_CCtx *cctx);
size_t     ZSTD_CCtx_setParameter(ZSTD_CCtx *cctx, ZSTD_cParameter param, int value);
size_t     ZSTD_compress2(ZSTD_CCtx *cctx, void *dst, size_t dstCapacity, const void *src, size_t srcSize);
size_t     ZSTD_compressStream2(ZSTD_CCtx *cctx, ZSTD_outBuffer *output, ZSTD_inBuffer *input,
                                ZSTD_EndDirective endOp);
size_t     ZSTD_compressBound(size_t srcSize);
size_t     ZSTD_CStreamOutSize(void);

ZSTD_DCtx    *ZSTD_createDCtx(void);
size_t        ZSTD_freeDCtx(ZSTD_DCtx *dctx);
size_t        ZSTD_sizeof_DCtx(const ZSTD_DCtx *dctx);
ZSTD_DStream *ZSTD_createDStream(void);
size_t        ZSTD_freeDStream(ZSTD_DStream *zds);
size_t        ZSTD_sizeof_DStream(const ZSTD_DStream *zds);
size_t        ZSTD_initDStream(ZSTD_DStream *zds);
size_t        ZSTD_decompressStream(ZSTD_DStream *zds, ZSTD_outBuffer *output, ZSTD_inBuffer *input);
size_t        ZSTD_decompressDCtx(ZSTD_DCtx *dctx, void *dst, size_t dstCapacity, const void *src,
                                  size_t srcSize);
#endif
