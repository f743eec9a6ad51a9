"""Print the last round's launches from an `ncu --csv` launch list (dev tool). usage: python tools/launch_table.py launches.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[h]
recs = collections.OrderedDict()
for r in rows[h + 1:]:
    d = dict(zip(hdr, r))
    recs.setdefault(d['ID'], {'k': d['Kernel Name'][:46]})[d['Metric Name']] = d['Metric Value']
items = list(recs.items())
idx = [i for i, (k, v) in enumerate(items) if 'photon_deposit' in v['k']]
tot = 0
for k, v in items[idx[-2] + 1: idx[-1] + 1]:
    t = float(v.get('gpu__time_duration.sum')); tot += t
    g = lambda m: float(v.get(m, 'nan'))
    print(f"{v['k']:46s} {t/1e6:8.3f} ms  thr/inst {g('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f}  inst {g('smsp__inst_executed.sum')/1e6:8.1f} M  "
          f"warps {g('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} %  issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} %  "
          f"dram_r {g('dram__bytes_read.sum')/1e6:8.1f} MB  dram_w {g('dram__bytes_write.sum')/1e6:8.1f} MB")
print(f"sum {tot/1e6:.3f} ms (serialised under ncu)")
