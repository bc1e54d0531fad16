/* emu_kernels.cpp — TEST INFRASTRUCTURE ONLY: runs the product's kernel sources
 * (libzseek_b200/csrc/zsk_*.cuh) on the lock-step CPU emulator of cuda_emu.h so tests can compare the
 * kernel LOGIC with the oracle without a GPU.  Not part of libzseek_b200.so. */
#define ZSK_EMU 1
#include "cuda_emu.h"
#include "../../libzseek_b200/csrc/zsk_lz4.cuh"
#include "../../libzseek_b200/csrc/zsk_lz4_lane.cuh"
#include "../../libzseek_b200/csrc/zsk_zstd_pipe.cuh"
#include "../../libzseek_b200/csrc/zsk_seek.cuh"

static uint32_t g_emu_deferred = 0;

/* The shipped zstd path: the four pipeline kernels, then the one-CTA-per-frame kernel over the deferred list.
 * tiny_pools: pools far too small, so that P0 defers most frames (exercises the hand-over). */
static void emu_zstd_pipeline(zsk_decode_args a, uint32_t njobs, uint32_t ctas, const uint64_t *d_off, uint32_t first_frame,
                              const uint32_t *frame_ids, bool tiny_pools)
{
    uint64_t dsum = 0;
    for (uint32_t j = 0; j < njobs; j++) {
        const uint32_t f = frame_ids ? frame_ids[j] : first_frame + j;
        dsum += d_off[f + 1] - d_off[f];
    }
    zsk_zpipe_args z;
    z.a = a;
    z.blocks_cap = tiny_pools ? 3 : dsum / 2048 + 4 * njobs + 64;
    z.seqs_cap = tiny_pools ? 2000 : dsum / 4 + 64 * njobs + 1024;
    z.lits_cap = tiny_pools ? 20000 : dsum / 4 * 3 + 64 * njobs + 4096;
    std::vector<zsk_zframe> frames(njobs + 1);
    std::vector<zsk_zblock> blocks(z.blocks_cap + 1);
    std::vector<uint32_t> seqs(3 * z.seqs_cap + 3, 0xEEEEEEEEu);
    std::vector<uint8_t> lits(z.lits_cap + ZSK_PAD_BACK, 0xEE);
    std::vector<uint32_t> deferred(njobs + 1);
    std::vector<uint32_t> bprog(2 * z.blocks_cap + 2, 0xEEEEEEEEu);
    z.bprog = bprog.data();
    unsigned long long ctr[ZSK_ZC_N] = { 0 };
    z.frames = frames.data(); z.blocks = blocks.data(); z.seqs = seqs.data(); z.lits = lits.data();
    z.ctr = ctr; z.deferred = deferred.data();
    emu::launch(dim3((njobs + 127) / 128), dim3(128), 0, [&] { zsk_zstd_index_kernel(z); });
    emu::launch(dim3(ctas), dim3(ZSK_ZFSE_THREADS), ZSK_ZFSE_SMEM, [&] { zsk_zstd_fse_kernel(z); });
    emu::launch(dim3(ctas), dim3(32), ZSK_ZHUF_SMEM, [&] { zsk_zstd_huf_kernel(z); });
    emu::launch(dim3(ctas), dim3(32), ZSK_ZX_SMEM, [&] { zsk_zstd_exec_kernel(z); });
    uint32_t counter = 0;
    zsk_decode_args d = a;
    d.job_list = z.deferred;
    d.job_list_count = z.ctr + ZSK_ZC_DEFERRED;
    d.work_counter = &counter;
    emu::launch(dim3(ctas), dim3(ZSK_ZSTD_CTA_THREADS), 0, [&] { zsk_zstd_decode_kernel(d); });
    g_emu_deferred = (uint32_t)ctr[ZSK_ZC_DEFERRED];
}

extern "C" {
__attribute__((visibility("default"))) uint32_t emu_last_deferred() { return g_emu_deferred; }
__attribute__((visibility("default"))) void emu_set_addr_bias(uint32_t bias) { zsk_emu_addr_bias = bias & ~15u; }

/* comp points at the byte with file offset comp_base; the caller guarantees >= 16 readable bytes
 * before it and >= 64 after the last frame (same contract as the device buffers). */
__attribute__((visibility("default")))
void emu_decode(int codec, const uint8_t *comp, uint64_t comp_base, const uint64_t *c_off, const uint64_t *d_off,
                const uint32_t *frame_ids, const uint64_t *dst_offs, uint8_t *dst, uint64_t dst_base,
                uint32_t first_frame, uint32_t njobs, int32_t *status, uint32_t ctas, const uint32_t *limits)
{
    uint32_t counter = 0;
    std::vector<uint8_t> scratch((size_t)ctas * ZSK_LIT_SCRATCH + 64);
    zsk_decode_args a{};
    a.c_off = c_off; a.d_off = d_off; a.comp = comp; a.comp_base = comp_base; a.frame_ids = frame_ids;
    a.dst_offs = dst_offs; a.dst = dst; a.dst_base = dst_base; a.first_frame = first_frame; a.njobs = njobs;
    a.status = status; a.work_counter = &counter; a.scratch = scratch.data(); a.limits = limits;
    a.dsize_sum = 0; a.job_list = nullptr; a.job_list_count = nullptr;
    if (codec == 1) emu::launch(dim3(ctas), dim3(ZSK_LZ4_CTA_THREADS), 0, [&] { zsk_lz4_decode_batch_kernel(a); });           /* shipped default */
    else if (codec == 102) emu::launch(dim3(ctas), dim3(ZSK_LZ4L_THREADS), ZSK_LZ4L_SMEM, [&] { zsk_lz4_decode_lane_kernel(a); });  /* many-frame launches */
    else if (codec == 200) emu::launch(dim3(ctas), dim3(ZSK_ZSTD_CTA_THREADS), 0, [&] { zsk_zstd_decode_kernel(a); }); /* one CTA per frame: only deferred frames in the product */
    else emu_zstd_pipeline(a, njobs, ctas, d_off, first_frame, frame_ids, codec == 201);
}

/* The device side of a stream-ordered batch (reader.c zseek_b200_pread_batch_async), kernel after kernel as the product
 * queues them: K1 lookup -> zsk_compact_kernel (touched frames -> job list, frame -> slot map, job COUNT in "device"
 * memory) -> decode launch that reads the job count through njobs_dev (njobs = upper bound) with per-job limits ->
 * K4 gather that also writes the results.  ctl[0] = job count, ctl[1] = error flag of the compaction. */
__attribute__((visibility("default")))
void emu_batch(int codec, const uint8_t *comp, const uint64_t *c_off, const uint64_t *d_off, uint32_t nframes,
               uint32_t shard_lo, uint32_t shard_hi, const uint64_t *offsets, const uint64_t *counts, uint64_t fixed_count,
               uint32_t n, uint64_t slot_size, uint32_t max_jobs, uint8_t *slab, uint8_t *dst, uint64_t dst_stride,
               int64_t *results, int32_t *job_status, uint32_t *ctl, int use_limits, uint32_t ctas)
{
    std::vector<int32_t> frame(n + 1);
    std::vector<uint32_t> inframe(n + 1), nbytes(n + 1), touched(nframes + 1, 0u);
    std::vector<uint32_t> job_ids(max_jobs + 1, 0xEEEEEEEEu), job_limits(max_jobs + 1, 0xEEEEEEEEu);
    std::vector<uint64_t> job_offs(max_jobs + 1, ~0ull);
    std::vector<int64_t> frame_src(nframes + 1, -7);
    ctl[0] = ctl[1] = 0;
    zsk_lookup_args la{d_off, nframes, offsets, counts, fixed_count, n, frame.data(), inframe.data(), nbytes.data(), touched.data()};
    emu::launch(dim3(2), dim3(256), 0, [&] { zsk_lookup_kernel(la); });
    zsk_compact_args ca{touched.data(), nframes, shard_lo, shard_hi, slot_size, max_jobs, job_ids.data(), job_offs.data(),
                        job_limits.data(), frame_src.data(), ctl, ctl + 1};
    emu::launch(dim3(2), dim3(256), 0, [&] { zsk_compact_kernel(ca); });
    if (max_jobs) {
        uint32_t counter = 0;
        std::vector<uint8_t> scratch((size_t)ctas * ZSK_LIT_SCRATCH + 64);
        zsk_decode_args a{};
        a.c_off = c_off; a.d_off = d_off; a.comp = comp; a.comp_base = 0; a.frame_ids = job_ids.data();
        a.dst_offs = job_offs.data(); a.dst = slab; a.dst_base = 0; a.first_frame = 0; a.njobs = max_jobs;
        a.status = job_status; a.work_counter = &counter; a.scratch = scratch.data();
        a.limits = use_limits ? job_limits.data() : nullptr;
        a.njobs_dev = ctl; a.job_base = 0;
        if (codec == 1) emu::launch(dim3(ctas), dim3(ZSK_LZ4_CTA_THREADS), 0, [&] { zsk_lz4_decode_batch_kernel(a); });
        else if (codec == 102) emu::launch(dim3(ctas), dim3(ZSK_LZ4L_THREADS), ZSK_LZ4L_SMEM, [&] { zsk_lz4_decode_lane_kernel(a); });
        else {
            /* pools sized like the launch layer does it: from the upper bound of the job count */
            std::vector<uint32_t> ids(max_jobs);
            const uint32_t live = ctl[0] < max_jobs ? ctl[0] : max_jobs;
            for (uint32_t j = 0; j < max_jobs; j++) ids[j] = j < live ? job_ids[j] : (live ? job_ids[0] : 0u);
            emu_zstd_pipeline(a, max_jobs, ctas, d_off, 0, ids.data(), false);
        }
    }
    zsk_gather_args ga{frame.data(), inframe.data(), nbytes.data(), frame_src.data(), slab, dst, nullptr, dst_stride, n, results};
    emu::launch(dim3(2), dim3(256), 0, [&] { zsk_gather_kernel(ga); });
}

__attribute__((visibility("default")))
void emu_lookup(const uint64_t *d_off, uint32_t nframes, const uint64_t *offsets, const uint64_t *counts,
                uint64_t fixed_count, uint32_t n, int32_t *frame, uint32_t *inframe, uint32_t *nbytes, uint32_t *touched)
{
    zsk_lookup_args a{d_off, nframes, offsets, counts, fixed_count, n, frame, inframe, nbytes, touched};
    emu::launch(dim3(2), dim3(256), 0, [&] { zsk_lookup_kernel(a); });
}

__attribute__((visibility("default")))
void emu_gather(const int32_t *frame, const uint32_t *inframe, const uint32_t *nbytes, const int64_t *frame_src,
                const uint8_t *src_base, uint8_t *dst, const uint64_t *dst_offs, uint64_t dst_stride, uint32_t n)
{
    zsk_gather_args a{frame, inframe, nbytes, frame_src, src_base, dst, dst_offs, dst_stride, n};
    emu::launch(dim3(2), dim3(256), 0, [&] { zsk_gather_kernel(a); });
}
}
