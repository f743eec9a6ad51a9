set -x
CMD="python bench.py --workload c5_dragon_4096 --photons 16777216 --steps 1 --warmup 1 --cpu-photons 0 --e2e-rounds 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"photon_deposit" -s 2 -c 1 -f -o gpurun_out/prof_c5_dep $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
