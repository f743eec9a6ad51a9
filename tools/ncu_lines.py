"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by source line (dev tool).
usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:X | python tools/ncu_lines.py [topN]"""
import csv, sys, collections
rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cur_file = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or r[0] == '': continue
    try: line = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr[4:], r[4:]))
    def f(k):
        try: return float(d.get(k, '0').replace(',', ''))
        except ValueError: return 0.0
    key = (cur_file, line)
    a = agg.setdefault(key, dict(src=r[1].strip()[:90], samples=0, inst=0, tinst=0, lsb=0, math=0, lg=0, wait=0, br=0))
    a['samples'] += f('# Samples'); a['inst'] += f('Instructions Executed'); a['tinst'] += f('Thread Instructions Executed')
    a['lsb'] += f('stall_long_sb'); a['math'] += f('stall_math'); a['lg'] += f('stall_lg'); a['wait'] += f('stall_wait'); a['br'] += f('stall_branch_resolving')
tot = sum(a['samples'] for a in agg.values()) or 1
toti = sum(a['inst'] for a in agg.values()) or 1
print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
print(f"{'samp%':>6} {'inst%':>6} {'thr/inst':>8} {'long_sb':>7} {'math':>6}  where")
for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    eff = a['tinst'] / a['inst'] if a['inst'] else 0
    print(f"{100*a['samples']/tot:6.2f} {100*a['inst']/toti:6.2f} {eff:8.1f} {a['lsb']:7.0f} {a['math']:6.0f}  {fn}:{ln}  {a['src']}")
