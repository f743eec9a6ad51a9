# One bench line per BASELINE config (short runs; the headline c3 run is separate). Results: gpurun_out/wl_<name>.json
for w in c1_spheres_bezier c2_bunny_chess c4_bump_dof c5_dragon_4096; do
  extra=""
  if [ $w = c5_dragon_4096 ]; then extra="--photons 134217728 --steps 2 --warmup 1 --e2e-rounds 0"; else extra="--steps 3 --warmup 3"; fi
  timeout 600 python bench.py --workload $w $extra --cpu-photons 100000 --f64-too 0 --shipped-photons 0 > gpurun_out/wl_$w.json 2> gpurun_out/wl_$w.err || { echo "FAILED $w"; tail -5 gpurun_out/wl_$w.err; }
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/wl_$w.json'))
    print('$w', 'photons/s', round(d['value']/1e6,1), 'M  ms/step', round(d['ms_per_step'],2), 'eye rays/s', round(d['eye_rays_per_s']/1e6,1), 'M  hitpoints', d['config']['hitpoints'], {k:round(v['seconds']*1e3,2) for k,v in d['kernels'].items()}, 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']/1e6,3), 'e2e', d['e2e'] and round(d['e2e']['value']/1e6,1))
except Exception as e: print('$w', 'no result', e)
PY
done
