"""profiles/r02_workloads.md from the bench lines saved under profiles/ (dev tool)."""
import json
L = lambda f: json.loads(open(f).read().strip().splitlines()[-1])
P = "profiles/r02_bench"
names = {"c1_spheres_bezier": "c1 512², spheres + Bezier vase, 1 Mi photons/round", "c2_bunny_chess": "c2 1024², bunny + chess floor, 4 Mi",
         "c3_dragon_glass": "c3 1024², glass dragon, 16 Mi (headline)", "c4_bump_dof": "c4 1920×1080, bump floor + DOF ×4 samples, 16 Mi",
         "c5_dragon_4096": "c5 4096², dragon, 128 Mi per step on 1 GPU"}
out = ["# r02 — every BASELINE config on one B200, and the multi-GPU points\n",
       "`python bench.py --workload <name>` (`--steps 3 --warmup 3`; c3 is the full default run `r02_final_bench.json`; c5 `--photons 134217728 --steps 2 --warmup 1`). "
       "Full lines: `profiles/r02_bench_<name>.json`. `photons/s` is device-timed over whole rounds (trace + sort + deposit + update); `e2e` is a whole "
       "`render()` of the config's own round count (10 / 20 / 50 / 20) from host arrays to the host image through the C ABI; `CPU` is the oracle port on "
       "the box's host threads; `deposit frac` is the deposit kernel's compulsory bytes / time / 6,552 GB/s (bench.py `roofline`), `step frac` all algorithmic bytes of a round over the round.\n",
       "| config | hitpoints | photons/s | ms/step | trace / sort / deposit / update (ms) | eye rays/s | trace frac | deposit frac | step frac | e2e photons/s | CPU photons/s |", "|---|---|---|---|---|---|---|---|---|---|---|"]
one = {}
for w in names:
    d = L("profiles/r02_final_bench.json" if w == "c3_dragon_glass" else f"{P}_{w}.json")
    one[w] = d
    k, r = d["kernels"], d["roofline"]
    fr = {r["kernel"]: r["frac"], **{a: b["frac"] for a, b in r["other_kernel"].items()}}
    out.append(f"| {names[w]} | {d['config']['hitpoints']:,} | {d['value']/1e6:.1f} M | {d['ms_per_step']:.2f} | {k['photon_trace_family']['seconds']*1e3:.2f} / "
               f"{k['bin_scan+bin_scatter_kernel']['seconds']*1e3:.2f} / {k['photon_deposit_kernel']['seconds']*1e3:.2f} / {k['round_update_kernel']['seconds']*1e3:.2f} | "
               f"{d['eye_rays_per_s']/1e6:.0f} M | {fr['photon_trace_family']:.2f} | {fr['photon_deposit_kernel']:.2f} | {r['whole_step']['frac']:.2f} | "
               f"{(str(round(d['e2e']['value']/1e6, 1)) + ' M') if d['e2e'] else '—'} | {d['cpu_baseline']['value']/1e6:.2f} M |")
out += ["\nr01 → r02 on one GPU: c1 220 → 222, c2 940 → 1030, c3 862 → 967, c4 173 → 184, c5 620 → 718 M photons/s (4-wide BVH, staged-candidate reuse, two-phase staging, 64-byte deposit records, 2^23 sort bins for the large hash table).\n",
        "## Multi-GPU (one box, torchrun, one process per GPU; r02 code: the 2-GPU c3 rows with the final code, the 8-GPU rows and the 2-GPU c2 rows one change earlier, before the 64-byte deposit records — their `vs 1 GPU` is against the 1-GPU rate of that code, 892 M for c3)\n",
        "| GPUs | config | scaling | collective | photons/s | ms/step | e2e photons/s | vs 1 GPU |", "|---|---|---|---|---|---|---|---|"]
c3 = one["c3_dragon_glass"]
out.append(f"| 1 | c3 | — | — | {c3['value']/1e6:.1f} M | {c3['ms_per_step']:.2f} | {c3['e2e']['value']/1e6:.1f} M | 1.00 |")
def row(n, f, what, scaling, base):
    d = L(f)
    e = d.get("e2e")
    out.append(f"| {n} | {what} | {scaling} | {d['config'].get('collective')} | {d['value']/1e6:.1f} M | {d['ms_per_step']:.2f} | {(str(round(e['value']/1e6, 1)) + ' M') if e else '—'} | "
               f"{d['value']/base:.2f} ({100*d['value']/base/n:.0f} % of linear) |")
row(2, f"{P}_2gpu.json", "c3, 16 Mi photons per GPU per round", "weak", c3["value"])
C3_THEN = 892.3e6  # 1-GPU rate of the code the 8-GPU rows were measured with
row(8, f"{P}_8gpu.json", "c3, 16 Mi photons per GPU per round", "weak", C3_THEN)
row(8, f"{P}_8gpu_strong.json", "c3, 16 Mi photons per round split over the GPUs (2 Mi each)", "strong", C3_THEN)
for c, w, what in (("c1", "c1_spheres_bezier", "1 Mi photons per GPU per round"), ("c2", "c2_bunny_chess", "4 Mi photons per GPU per round")):
    row(8, f"{P}_8gpu_{c}.json", f"{c}, {what}, all-reduce on a side stream", "weak", one[w]["value"])
    row(8, f"{P}_8gpu_{c}_torch.json", f"{c}, {what}, all-reduce in stream order", "weak", one[w]["value"])
row(2, f"{P}_2gpu_c2_nccl.json", "c2, ncclAllReduce in stream order (the library's default)", "weak", one["c2_bunny_chess"]["value"])
row(2, f"{P}_2gpu_c2_peer.json", "c2, exchange over peer memory fused with the update", "weak", one["c2_bunny_chess"]["value"])
row(2, f"{P}_2gpu_peer.json", "c3, exchange over peer memory fused with the update", "weak", c3["value"])
row(8, f"{P}_8gpu_c2_nccl.json", "c2, ncclAllReduce in stream order (the library's default)", "weak", one["c2_bunny_chess"]["value"])
row(8, f"{P}_8gpu_nccl2.json", "c3, ncclAllReduce in stream order (the library's default)", "weak", C3_THEN)
row(8, f"{P}_8gpu_c1_peer.json", "c1, peer exchange (remote loads one after the other)", "weak", one["c1_spheres_bezier"]["value"])
row(8, f"{P}_8gpu_c2_peer.json", "c2, peer exchange (remote loads one after the other)", "weak", one["c2_bunny_chess"]["value"])
row(8, f"{P}_8gpu_peer.json", "c3, peer exchange (remote loads one after the other)", "weak", C3_THEN)
row(8, f"{P}_8gpu_c1_peer2.json", "c1, peer exchange, remote loads issued together, side stream at the highest priority", "weak", one["c1_spheres_bezier"]["value"])
row(8, f"{P}_8gpu_c2_peer2.json", "c2, peer exchange, remote loads issued together, side stream at the highest priority", "weak", one["c2_bunny_chess"]["value"])
row(8, f"{P}_8gpu_peer2.json", "c3, peer exchange, remote loads issued together, side stream at the highest priority", "weak", C3_THEN)
row(8, f"{P}_8gpu_c4.json", "c4, 16 Mi photons per GPU per round", "weak", one["c4_bump_dof"]["value"])
row(8, f"{P}_8gpu_c5.json", "c5, 1 Gi photons per round split over the GPUs", "strong", one["c5_dragon_4096"]["value"])
r = L(f"{P}_8gpu_reference_arm.json")
out.append(f"\nThe rows marked `native` ran the all-reduce inside the library on a side stream under the next round's trace launches; that was measured SLOWER than the "
           f"in-stream collective (`torch` rows: same NCCL call, in stream order) on the short rounds of c1/c2 — NCCL's blocks cannot become resident next to the "
           f"persistent emission kernel — and the library now issues it in stream order. Reference arm under torchrun at N=8: {r['value']/1e6:.2f} M photons/s on "
           f"{r['cpu_baseline']['cores']} host threads with OMP_NUM_THREADS={r['cpu_baseline']['omp_num_threads_env']} in the environment (the thread count comes from the affinity mask).")
out.append("\nThe exchange over peer memory (`cgrt_peer_*`: flags after the deposit kernel, one kernel on a side stream that reads every rank's accumulators over NVLink, "
           "applies the update and clears the next buffer; `--collective peer`) is the faster one at 2 GPUs (c2: 4.12 vs 4.39 ms per round, 101 % of linear) and on "
           "c1 at 8 GPUs (5.25 vs 5.54 ms), but loses to the in-stream ncclAllReduce on c2 and c3 at 8 GPUs (5.17 vs 4.71, 19.66 vs 19.12 ms): its reduce kernel "
           "only becomes resident as the next round's trace kernels drain, and reads 8 x 17 MB per rank where the all-reduce moves 2 x 7/8 x 17 MB. Issuing the eight "
           "remote loads of a hitpoint together helped c3 (19.24); giving the side stream the highest priority made c2 worse (8.06) and was taken out again. "
           "`native` (ncclAllReduce inside `cgrt_round_update`, in stream order) is therefore the default of the library, of `bench.py` and of the C++ `render()`.")
out.append("\nStrong scaling of c3 at 8 GPUs (2 Mi photons per GPU and round) is bound by the fixed part of a round: 11 trace launches whose late passes hold a few "
           "thousand rays, the 18 MB all-reduce and the update over all 1.13 M hitpoints (replicated) — 4.2 ms per round against 18.8 / 8 = 2.35 ms.")
open("profiles/r02_workloads.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
