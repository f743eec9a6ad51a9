python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 4 --warmup 2 --cpu-photons 0 --e2e-rounds 0 > gpurun_out/bench_q2.json 2> gpurun_out/bench_q.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_q2.json'))
print('value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],2), d['kernels']['photon_trace_kernel']['split_ms'], round(d['kernels']['photon_deposit_kernel']['seconds']*1e3,2))
PY
