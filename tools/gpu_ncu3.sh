# ncu --set full of one launch each: first traversal launch, emission kernel, deposit kernel of the timed round. usage: gpu_ncu3.sh <tag>
TAG=${1:-dev}
CMD="python bench.py --steps 1 --warmup 1 --cpu-photons 0 --e2e-rounds 0 --f64-too 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:photon_traverse_kernel -s 5 -c 1 -f -o gpurun_out/${TAG}_trav $CMD > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:photon_trace_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_emit $CMD > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:photon_deposit_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_dep $CMD > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full1.log gpurun_out/ncu_full2.log gpurun_out/ncu_full3.log | cut -c1-200
