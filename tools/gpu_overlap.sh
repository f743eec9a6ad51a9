for ov in 1 0 1 0; do
  python bench.py --steps 10 --warmup 3 --cpu-photons 0 --e2e-rounds 0 --overlap $ov > gpurun_out/bench_ov.json 2> gpurun_out/bench_ov.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ov.json'))
print('overlap $ov', 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],2), {k:round(v['seconds']*1e3,2) for k,v in d['kernels'].items()}, d['clocks'])
PY
done
