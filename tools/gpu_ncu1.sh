set -x
CMD="python bench.py --steps 1 --warmup 1 --photons 4194304 --cpu-photons 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:photon_ -s 12 -c 4 -o gpurun_out/prof_r01_photon $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
