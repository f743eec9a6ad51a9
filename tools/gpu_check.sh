set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --cpu-photons 0 --e2e-rounds 4 > gpurun_out/bench_v17_f32.json 2> gpurun_out/bench_v17_f32.err; tail -3 gpurun_out/bench_v17_f32.err; cat gpurun_out/bench_v17_f32.json
python bench.py --steps 3 --warmup 2 --cpu-photons 0 --e2e-rounds 0 --accum 0 > gpurun_out/bench_v17_f64.json 2> gpurun_out/bench_v17_f64.err; tail -3 gpurun_out/bench_v17_f64.err; cat gpurun_out/bench_v17_f64.json
