for cfg in "0 0" "3 1" "2 2" "3 2" "2 1" "4 1"; do
  set -- $cfg
  if [ $1 = 0 ]; then unset CGRT_TRACE_BPS CGRT_DEPOSIT_BPS; else export CGRT_TRACE_BPS=$1 CGRT_DEPOSIT_BPS=$2; fi
  python bench.py --steps 6 --warmup 3 --cpu-photons 0 --e2e-rounds 0 > gpurun_out/bench_share.json 2> gpurun_out/bench_share.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_share.json'))
print('trace_bps $1 deposit_bps $2', 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],2), {k:round(v['seconds']*1e3,2) for k,v in d['kernels'].items()})
PY
done
