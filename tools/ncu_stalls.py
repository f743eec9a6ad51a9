"""Warp-state (stall reason) shares and pipe instruction counts of every kernel in an ncu report (dev tool).
usage: python tools/ncu_stalls.py report.ncu-rep [kernel substring]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if sub not in d['Kernel Name']: continue
    print('----', d['Kernel Name'][:100])
    st = []
    for k, v in d.items():
        if k.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in k:
            try: st.append((float(v.replace(',', '')), k.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError: pass
    tot = sum(x for x, _ in st) or 1
    print('  stall samples: ' + '  '.join(f"{k} {100*x/tot:.1f}%" for x, k in sorted(st, reverse=True)[:10]))
    for k in ('smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
              'smsp__inst_executed.sum', 'smsp__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_alu.sum',
              'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_xu.sum', 'smsp__inst_executed_pipe_cbu.sum', 'smsp__inst_executed_pipe_adu.sum',
              'smsp__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
              'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu.sum'):
        if k in d: print(f"  {k:70s} {d[k]}")
