python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1
for w in c3_dragon_glass c1_spheres_bezier; do
python bench.py --workload $w --steps 3 --warmup 2 --cpu-photons 0 --e2e-rounds 0 > gpurun_out/chk_$w.json 2> gpurun_out/chk_$w.err || tail -5 gpurun_out/chk_$w.err
python - <<PY
import json
d=json.load(open('gpurun_out/chk_$w.json'))
print('$w', 'photons/s', round(d['value']/1e6,1), 'M  ms/step', round(d['ms_per_step'],2), 'eye rays/s', round(d['eye_rays_per_s']/1e6,1), {k:round(v['seconds']*1e3,2) for k,v in d['kernels'].items()})
PY
done
