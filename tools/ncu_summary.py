"""Summarise an .ncu-rep: key raw metrics + stall breakdown + hottest SASS lines (reads ncu CSV pages)."""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__grid_size', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_static',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = page(rep, 'raw')
    hdr, units, vals = rows[0], rows[1], rows[2]
    print('== raw metrics')
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            print(f'{h:70s} {v} {u}')
    rows = page(rep, 'source')
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stall_cols = [h for h in hdr if h.startswith('stall') and 'Not Issued' not in h]
    agg = {h: sum(float(r[ix[h]] or 0) for r in data) for h in stall_cols}
    tot = sum(agg.values()) or 1
    print('== warp stall sampling (all samples)')
    for h, v in sorted(agg.items(), key=lambda x: -x[1])[:9]:
        print(f'{h:28s} {100 * v / tot:5.1f}%')
    tot_inst = sum(float(r[ix['Instructions Executed']] or 0) for r in data)
    print(f'== total warp instructions {tot_inst:.3e}; hottest lines by samples')
    for r in sorted(data, key=lambda r: -float(r[ix['# Samples']] or 0))[:top]:
        st = {h: float(r[ix[h]] or 0) for h in stall_cols}
        s = sorted(st.items(), key=lambda x: -x[1])[:2]
        print(r[ix['Address']][-5:], f"{r[ix['Source']][:64]:64s}", 'samp', r[ix['# Samples']], 'inst', r[ix['Instructions Executed']],
              'thr/inst', r[ix['Avg. Threads Executed']], s)


if __name__ == '__main__':
    main()
