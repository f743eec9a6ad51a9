# one kernel of the timed round under ncu --set full. usage: gpu_ncu1.sh <kernel regex> <skip> <out name>
set -x
CMD="python bench.py --steps 1 --warmup 1 --cpu-photons 0 --e2e-rounds 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c 1 -f -o gpurun_out/$3 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
