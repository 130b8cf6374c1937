"""Print the handful of ncu metrics the roofline discussion needs from a .ncu-rep (run where ncu is installed)."""
import csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active']


def main(path, stalls=True):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print('==', name[:110])
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f'  {h:75s} {r[i]:>16s} {units[i]}')
        if stalls:
            st = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
                  if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h]
            for v, h in sorted(st, reverse=True)[:7]:
                print(f'  stall {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:28s} {v:8.3f}')


if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
