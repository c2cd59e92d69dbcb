"""Condenses an .ncu-rep (ncu --set full) into a small per-kernel JSON summary
for profiles/.  Usage: python tools/ncu_summary.py in.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    'gpu__time_duration.sum': 'duration_us',
    'dram__bytes_read.sum': 'dram_read_bytes',
    'dram__bytes_write.sum': 'dram_write_bytes',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct_of_peak',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct_of_peak',
    'sm__issue_active.avg.pct_of_peak_sustained_elapsed': 'issue_active_pct',
    'smsp__inst_executed.sum': 'warp_instructions',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed': 'smem_wavefront_pct',
    'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'achieved_occupancy_pct',
    'launch__registers_per_thread': 'registers',
    'launch__shared_mem_per_block_dynamic': 'dyn_smem_bytes' ,
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct',
    'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_cycles_active_pct',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'fma_pipe_pct',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'alu_pipe_pct',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active': 'lsu_pipe_pct',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'xu_pipe_pct',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'stall_long_scoreboard',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'stall_short_scoreboard',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'stall_barrier',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'stall_math_throttle',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio': 'stall_mio_throttle',
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = dict(kernel=d.get('Kernel Name', '')[:100])
        for src, dst in KEYS.items():
            if src in d and d[src] != '':
                try:
                    v = float(d[src].replace(',', ''))
                except ValueError:
                    continue
                u = units[hdr.index(src)]
                if dst == 'duration_us':
                    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
                if u == 'Kbyte' or u == 'Kbyte/block':
                    v *= 1e3
                if u == 'Mbyte':
                    v *= 1e6
                if u == 'Gbyte':
                    v *= 1e9
                k[dst] = v
        res.append(k)
    json.dump(dict(source=rep.split('/')[-1], note='ncu --set full --clock-control none; durations under ncu are '
                   'cold-cache/serialised and never reported as bench values', kernels=res), open(out, 'w'), indent=1)
    for k in res:
        print(json.dumps(k))


if __name__ == '__main__':
    main()
