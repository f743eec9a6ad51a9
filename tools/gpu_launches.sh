# launch list (time, lanes per instruction, instructions, occupancy, issue rate, DRAM bytes) of one bench run. usage: gpu_launches.sh <out name>
set -x
CMD="python bench.py --steps 1 --warmup 1 --cpu-photons 0 --e2e-rounds 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${1:-dev}.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
