#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS source page) into the handful of numbers DESIGN.md/profiles cite."""
import collections, csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
rep = sys.argv[1]
traffic_record = None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
print('kernel:', vals[hdr.index('Kernel Name')])
if '--traffic' in sys.argv:      # record DRAM bytes per launch for bench.py's roofline.traffic
    import json, os
    kname = vals[hdr.index('Kernel Name')]
    cls = 'v_pass_stream' if 'iwb_v2' in kname or 'vstream' in kname else ('h_pass_epilogue' if 'iwb_h2' in kname or 'hstream' in kname else kname)
    def num(m):
        i = hdr.index(m); v = float(vals[i].replace(',', '')); u = units[i]
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    tp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'traffic.json')
    d = json.load(open(tp)) if os.path.exists(tp) else {}
    # the capture is only valid for the kernels it was taken on: stamp it with bench.py's fingerprint of their sources
    # (bench.py prints null when HEAD's differ)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    sha = bench.kernel_sources_sha()
    if d.get('kernel_sources_sha') != sha:
        d = {}
    d['kernel_sources_sha'] = sha
    d['commit'] = subprocess.run(['git', '-C', root, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    d[cls] = num('dram__bytes_read.sum') + num('dram__bytes_write.sum')
    traffic_record = (tp, d, cls)
    sys.argv.remove('--traffic')
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f'  {w} = {vals[i]} {units[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]; data = rows[2:]
iex, isrc, ist = h.index('Instructions Executed'), h.index('Source'), h.index('Warp Stall Sampling (All Samples)')
tot = sum(int(r[iex]) for r in data)
ops = collections.Counter(); st = collections.Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[op] += int(r[iex]); st[op] += int(r[ist])
if traffic_record:               # warp instructions per launch, all and FP64 (DADD / DMUL / DFMA / DSETP): bench.py's issue-port model
    import json
    tp, d, cls = traffic_record
    d[cls + '_warp_instr'] = tot
    d[cls + '_warp_instr_fp64'] = sum(v for k, v in ops.items() if k in ('DADD', 'DMUL', 'DFMA', 'DSETP', 'DMNMX'))
    json.dump(d, open(tp, 'w'), indent=1)
print('  SASS instructions executed:', tot)
for k, v in ops.most_common(14):
    print(f'    {k:9s} {100 * v / tot:5.1f}%   stall samples {st[k]}')
if len(sys.argv) > 2:
    print('  hottest stall lines:')
    for r in sorted(data, key=lambda r: -int(r[ist]))[:int(sys.argv[2])]:
        print('   ', r[ist], r[iex], r[isrc][:100])
