# GPU tests, then one short bench line with the kernel split (dev loop)
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --cpu-photons 0 --e2e-rounds ${1:-0} > gpurun_out/b.json 2>gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/b.json'))
k = d['kernels']
print('photons/s %.1f M  ms/round %.2f  split %s  deposit %.2f  sort %.2f  e2e %s  eye %.2f ms' % (d['value'] / 1e6, d['ms_per_step'],
      {a.split('<')[-1].split('>')[0].split(' ')[0][:12]: round(b, 2) for a, b in k['photon_trace_kernel']['split_ms'].items()},
      k['photon_deposit_kernel']['seconds'] * 1e3, k['bin_scan+bin_scatter_kernel']['seconds'] * 1e3,
      d['e2e'] and round(d['e2e']['value'] / 1e6, 1), d['setup']['eye_ms']))
PY
