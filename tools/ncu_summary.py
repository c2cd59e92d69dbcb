"""Summarises an .ncu-rep (ncu --set full) into a small JSON for profiles/:

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--note "..."] > profiles/xxx.json

One entry per captured launch: duration, registers, shared memory, occupancy
limits, issue / pipe utilisation, DRAM bytes, L2 hit rate, shared-memory
wavefronts and the warp-stall breakdown (warps stalled per issue)."""
import argparse
import csv
import json
import subprocess
import sys

KEEP = {
    'gpu__time_duration.sum': 'duration',
    'launch__grid_size': 'grid', 'launch__block_size': 'block',
    'launch__registers_per_thread': 'registers',
    'launch__shared_mem_per_block_dynamic': 'smem_dynamic',
    'launch__occupancy_limit_shared_mem': 'occupancy_limit_smem_blocks',
    'launch__occupancy_limit_registers': 'occupancy_limit_regs_blocks',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'achieved_occupancy_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'smsp__inst_executed.sum': 'warp_instructions',
    'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
    'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_throughput_pct',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed': 'smem_wavefront_pct',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active': 'lsu_pipe_pct',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'fma_pipe_pct',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'alu_pipe_pct',
    'smsp__average_warp_latency_per_inst_issued.ratio': 'warp_latency_per_inst',
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--note', default='')
    args = ap.parse_args()
    out = subprocess.run(['ncu', '-i', args.report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        e = {'kernel': r[hdr.index('Kernel Name')][:100]}
        stalls = {}
        for i, h in enumerate(hdr):
            if h in KEEP and r[i] != '':
                try:
                    v = float(r[i].replace(',', ''))
                except ValueError:
                    v = r[i]
                e[KEEP[h]] = v if not units[i] else (('%g %s' % (v, units[i])) if isinstance(v, float) else v)
            elif h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.2:
                    stalls[h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = round(v, 2)
        e['warps_stalled_per_issue'] = stalls
        launches.append(e)
    print('{"report": %s, "note": %s, "launches": [' % (json.dumps(args.report), json.dumps(args.note)))
    print(',\n'.join(' ' + json.dumps(e) for e in launches))
    print(']}')


if __name__ == '__main__':
    main()
