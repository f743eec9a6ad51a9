set -x
CMD="python bench.py --steps 1 --warmup 1 --cpu-photons 0 --e2e-rounds 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"photon_" -s 24 -c 3 -f -o gpurun_out/prof_v12_trace $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
