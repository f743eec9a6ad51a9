for v in C A B; do
  if [ $v = base ]; then unset CGRT_LIB; else export CGRT_LIB=$PWD/gpurun_variants_$v.so; fi
  python bench.py --steps 3 --warmup 2 --cpu-photons 0 --e2e-rounds 0 > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_var_$v.json'))
print('variant $v', 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],2), d['kernels']['photon_trace_kernel']['split_ms'], round(d['kernels']['photon_deposit_kernel']['seconds']*1e3,2))
PY
done
