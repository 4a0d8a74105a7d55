"""Summarise an Nsight Compute report (read here, on the CPU box) into profiles/<name>.md + .json.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_local_step_fast64 ["free-form note"]
Extracts the per-launch metrics the roofline in bench.py is compared against (duration, DRAM traffic, pipe
utilisation, issue-slot use, stall breakdown, shared-memory wavefronts, registers / occupancy).
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ('duration_ms', 'gpu__time_duration.sum'),
    ('dram_read_bytes', 'dram__bytes_read.sum'),
    ('dram_write_bytes', 'dram__bytes_write.sum'),
    ('dram_throughput_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
    ('sm_throughput_pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
    ('issue_active_pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
    ('pipe_fma_inst_pct', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
    ('pipe_fma_cycles_pct', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
    ('pipe_alu_inst_pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
    ('pipe_lsu_inst_pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
    ('pipe_tensor_pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
    ('smem_wavefronts', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
    ('smem_wavefronts_pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
    ('smem_bank_conflicts', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
    ('warps_active_pct', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
    ('registers_per_thread', 'launch__registers_per_thread'),
    ('grid', 'launch__grid_size'),
    ('block', 'launch__block_size'),
    ('inst_executed', 'smsp__inst_executed.sum'),
    ('sm_clock_ghz', 'sm__cycles_elapsed.max.per_second'),
]
STALLS = ['barrier', 'branch_resolving', 'dispatch_stall', 'long_scoreboard', 'math_pipe_throttle', 'mio_throttle',
          'no_instruction', 'not_selected', 'selected', 'short_scoreboard', 'wait', 'lg_throttle', 'tex_throttle']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        rec = {'kernel': d.get('Kernel Name', '')}
        for k, m in KEYS:
            if m in d:
                try:
                    rec[k] = float(d[m].replace(',', ''))
                except ValueError:
                    rec[k] = d[m]
                rec[k + '_unit'] = units[hdr.index(m)]
        rec['stalls_per_issue'] = {}
        for s in STALLS:
            m = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s
            if m in d:
                try:
                    rec['stalls_per_issue'][s] = float(d[m])
                except ValueError:
                    pass
        launches.append(rec)
    with open(out + '.json', 'w') as f:
        json.dump({'report': rep, 'note': note, 'launches': launches}, f, indent=1)
    with open(out + '.md', 'w') as f:
        f.write('# ncu summary: %s\n\n%s\n\n' % (rep, note))
        for r in launches:
            f.write('## %s\n\n' % r['kernel'])
            for k, _ in KEYS:
                if k in r:
                    f.write('- %s: %s %s\n' % (k, r[k], r.get(k + '_unit', '')))
            if 'dram_read_bytes' in r and 'dram_write_bytes' in r:
                f.write('- dram traffic per launch (read+write, as reported): %s + %s %s\n'
                        % (r['dram_read_bytes'], r['dram_write_bytes'], r.get('dram_read_bytes_unit', '')))
            f.write('- warp stall cycles per issued instruction: %s\n\n'
                    % ', '.join('%s %.2f' % kv for kv in sorted(r['stalls_per_issue'].items(), key=lambda kv: -kv[1])))
    print('wrote', out + '.md', out + '.json')


if __name__ == '__main__':
    main()
