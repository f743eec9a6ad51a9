set -x
for mb in 0 32 64 96 200; do
  CGRT_L2_PERSIST_MB=$mb python bench.py --steps 3 --warmup 2 --cpu-photons 0 --e2e-rounds 0 > gpurun_out/bench_l2_$mb.json 2> gpurun_out/bench_l2_$mb.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_l2_$mb.json'))
print('L2 persist MB', $mb, 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],2), {k:round(v['seconds']*1e3,2) for k,v in d['kernels'].items()})
PY
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
