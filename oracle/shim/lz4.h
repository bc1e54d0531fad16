/* TEST INFRASTRUCTURE ONLY — the reference only #includes <lz4.h>; it uses nothing from it. */
#ifndef ZSK_SHIM_LZ4_H
#define ZSK_SHIM_LZ4_H
#endif
