/* TEST INFRASTRUCTURE ONLY — declaration-only ABI shim, see zstd.h in this directory. */
#ifndef ZSK_SHIM_ZSTD_ERRORS_H
#define ZSK_SHIM_ZSTD_ERRORS_H
typedef enum {
    ZSTD_error_no_error = 0,
    ZSTD_error_GENERIC = 1,
    ZSTD_error_memory_allocation = 64,
    ZSTD_error_frameIndex_tooLarge = 100
} ZSTD_ErrorCode;
#endif
