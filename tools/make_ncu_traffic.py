"""profiles/ncu_traffic.json from an `ncu --csv` launch list of `python bench.py --steps 1 --warmup 1 ...` (dev tool): DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) of the LAST round's launches, summed per kernel family the way bench.py's `kernels` are.
usage: python tools/make_ncu_traffic.py <launches.csv> <bench.json of the same code> <source label>"""
import collections, csv, json, sys
csvf, benchf, label = sys.argv[1:4]
rows = list(csv.reader(open(csvf)))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[h]
recs = collections.OrderedDict()
for r in rows[h + 1:]:
    d = dict(zip(hdr, r))
    recs.setdefault(d['ID'], {'k': d['Kernel Name']})[d['Metric Name']] = d['Metric Value']
items = list(recs.values())
dep = [i for i, v in enumerate(items) if 'photon_deposit' in v['k']]
last = items[dep[-2] + 1: dep[-1] + 1]   # everything after the previous round's deposit up to and including this round's
fam = collections.Counter(); ms = collections.Counter()
for v in last:
    b = float(v.get('dram__bytes_read.sum', 0)) + float(v.get('dram__bytes_write.sum', 0))
    t = float(v.get('gpu__time_duration.sum', 0)) / 1e6
    k = v['k']
    name = ('photon_trace_family' if ('photon_trace_kernel' in k or 'photon_traverse_kernel' in k or 'photon_bezier_kernel' in k) else
            'photon_deposit_kernel' if 'photon_deposit' in k else 'bin_scan+bin_scatter_kernel' if k.startswith('bin_') or 'bin_s' in k else
            'round_update_kernel' if 'round_update' in k else 'other')
    fam[name] += b; ms[name] += t
b = json.load(open(benchf))
out = {"source": label, "config": {"workload": b["config"]["workload"], "photons": b["config"]["photons_per_gpu_per_step"],
                                   "accum": 1 if "f32" in b["dtype"] else 0},
       "bytes_per_round": {k: v for k, v in fam.items() if k != 'other'}, "ncu_ms_per_round": dict(ms),
       "note": "per-launch times under ncu are cold-cache and serialised; only the bytes are used by bench.py"}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
