"""Write profiles/<tag>_ncu_summary.md from a launch list csv and an ncu --set full report (dev tool).
usage: python tools/make_ncu_summary.py <tag> <launches.csv> <report.ncu-rep> <bench.json>"""
import json, subprocess, sys, os
tag, csvf, rep, benchf = sys.argv[1:5]
here = os.path.dirname(os.path.abspath(__file__))
run = lambda *a: subprocess.run(a, capture_output=True, text=True).stdout
b = json.load(open(benchf))
k = b["kernels"]
sp = k["photon_trace_family"]["split_ms"]
out = []
out.append(f"# {tag} - ncu evidence, one B200, {b['config']['workload']} {b['config']['width']}x{b['config']['height']}, "
           f"{b['config']['photons_per_gpu_per_step']} photons per round\n")
out.append("Command under ncu: `python bench.py --steps 1 --warmup 1 --cpu-photons 0 --e2e-rounds 0` (the same command exited 0 without ncu immediately "
           "before each pass), `--clock-control none`. Numbers printed by runs under ncu are not bench values; the bench line of the same code is "
           f"`profiles/{tag}_bench.json` (un-profiled run: {b['value']/1e6:.1f} M photons/s, {b['ms_per_step']:.2f} ms per round).\n")
out.append(f"## 1. Launch list of one round (profiles/{tag}_launches.csv)\n")
out.append("CUDA-event split of the un-profiled bench run for comparison: " + ", ".join(f"{a} {v:.2f} ms" for a, v in sp.items()) +
           f"; photon_deposit_kernel {k['photon_deposit_kernel']['seconds']*1e3:.2f} ms; counting sort {k['bin_scan+bin_scatter_kernel']['seconds']*1e3:.2f} ms; "
           f"round_update_kernel {k['round_update_kernel']['seconds']*1e3:.2f} ms. ncu serialises launches and runs them cold, so the shares agree, not the absolutes.\n")
tbl = run(sys.executable, os.path.join(here, "launch_table.py"), csvf).splitlines()
# the last round of the run: from the last round_update before the final deposit back to the deposit
last_dep = max(i for i, l in enumerate(tbl) if "photon_deposit_kernel" in l)
start = max(i for i, l in enumerate(tbl[:last_dep]) if "photon_trace_kernel<1>" in l)
out.append("```\n" + "\n".join(tbl[start:last_dep + 1]) + "\n```\n")
out.append(f"## 2. Per-kernel metrics (`ncu --set full`, the 12 photon kernels of one round)\n")
out.append("```\n" + run(sys.executable, os.path.join(here, "ncu_raw.py"), rep) + "```\n")
out.append("## 3. Warp states (pc sampling) per kernel\n")
out.append("```\n" + run(sys.executable, os.path.join(here, "ncu_stalls.py"), rep) + "```\n")
out.append("## 4. Hot source lines (ncu --page source aggregated by tools/ncu_lines.py; first launch of each kernel)\n")
for name in ("photon_deposit", "photon_traverse", "photon_trace"):
    src = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass --kernel-name regex:{name} -c 1 2>/dev/null | {sys.executable} {here}/ncu_lines.py 16",
                         shell=True, capture_output=True, text=True).stdout
    out.append(f"### {name}\n```\n{src}```\n")
out.append("""## 5. Reading

* No kernel is bound by HBM (DRAM throughput <= 35 % of the measured copy peak) or by launch overhead (16 launches per ~19 ms round).
* photon_trace_kernel (emission and continuation) is bound by dependent fp64 latency: 4 warps per scheduler (114 registers), ~0.48 issue
  slots per cycle, warp states `wait` + `short_scoreboard` + `long_scoreboard` ~ 65 %. Removing instructions does not move it (r02: axis-aligned
  plane quotients without the dot products and a division-free texel index, both bit-exact: 4.78 vs 4.77 ms); neither do register caps of
  96/80/64 or block sizes 64/96/128 (r01).
* photon_traverse_kernel walks a 4-wide tree (r02): 7-9 of 32 lanes live per instruction (one ray per lane, path lengths from 1 to hundreds of
  boxes), `long_scoreboard` ~ 50 %: dependent node and triangle fetches at 8 warps per scheduler. It is sensitive to every L1 transaction
  (three unconditional stack stores per node: +11 %) and to occupancy (80 registers: +35 %), not to instruction count (fused slabs: flat).
* photon_deposit_kernel is bound by the LSU data pipe (`l1tex__data_pipe_lsu_wavefronts` ~ 92 % of peak: half shared-memory loads — the
  broadcast prefilter scan costs four wavefronts per staged candidate —, the rest record gathers, staging loads and one `red` per deposit).
  r02 removed per-group work (a cell's staged list is reused by the following batches; filter records only for culled survivors), a third of
  the record bytes (64-byte records: four loads instead of six) and, with float accumulators, the pair queue and the shared-memory copy of the hits
  (every lane walks its own surviving pairs from registers).
""")

open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out))
print("wrote", f"profiles/{tag}_ncu_summary.md")
