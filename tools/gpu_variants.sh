# A/B of library builds in one call: gpurun_variants_<v>.so, optional per-variant environment in VENV_<v>
for v in ${VARIANTS:-A B}; do
  export CGRT_LIB=$PWD/gpurun_variants_$v.so
  ev="VENV_$v"; 
  env ${!ev} python bench.py --steps 3 --warmup 2 --cpu-photons 0 --e2e-rounds 0 > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_var_$v.json'))
print('variant $v ${!ev}', 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],2), {a.split('<')[-1][:8]: round(b,2) for a,b in d['kernels']['photon_trace_family']['split_ms'].items()}, 'deposit', round(d['kernels']['photon_deposit_kernel']['seconds']*1e3,2), 'sort', round(d['kernels']['bin_scan+bin_scatter_kernel']['seconds']*1e3,3))
PY
done
