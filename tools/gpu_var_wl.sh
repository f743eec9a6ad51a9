# variants on several workloads. usage: VARIANTS="A B" WLS="c3_dragon_glass c4_bump_dof" bash tools/gpu_var_wl.sh
for w in ${WLS:-c3_dragon_glass}; do
for v in ${VARIANTS:-A B}; do
  export CGRT_LIB=$PWD/gpurun_variants_$v.so
  extra="--steps 3 --warmup 2"
  if [ $w = c5_dragon_4096 ]; then extra="--photons 134217728 --steps 2 --warmup 1"; fi
  python bench.py --workload $w $extra --cpu-photons 0 --e2e-rounds 0 --f64-too 0 > gpurun_out/bench_var_${w}_$v.json 2> gpurun_out/bench_var_${w}_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_var_${w}_$v.json'))
print('$w variant $v', 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],2), {a.split('<')[-1][:8]: round(b,2) for a,b in d['kernels']['photon_trace_family']['split_ms'].items()}, 'deposit', round(d['kernels']['photon_deposit_kernel']['seconds']*1e3,2), 'sort', round(d['kernels']['bin_scan+bin_scatter_kernel']['seconds']*1e3,3))
PY
done; done
