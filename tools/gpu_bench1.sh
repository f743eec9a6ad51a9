set -x
python bench.py --steps 2 --warmup 1 --photons 4194304 --cpu-photons 0 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
