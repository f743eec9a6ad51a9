# bench.py at N GPUs of one box (launched the way the driver does). usage: gpu_scale.sh N [extra bench args] ; result gpurun_out/scale_<N><suffix>.json
N=$1; shift
SUF=${SUFFIX:-}
if [ "$N" = 1 ]; then
  python bench.py --gpus 1 "$@" > gpurun_out/scale_${N}${SUF}.json 2> gpurun_out/scale_${N}${SUF}.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/scale_${N}${SUF}.json 2> gpurun_out/scale_${N}${SUF}.err
fi
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/scale_${N}${SUF}.json').read().strip().splitlines()[-1])
    print('N=$N', d['config']['workload'], d['scaling'], 'photons/s %.1f M' % (d['value'] / 1e6), 'ms/step %.2f' % d['ms_per_step'], 'launches', d['gpu_launches'])
except Exception as e:
    print('N=$N failed', e); print(open('gpurun_out/scale_${N}${SUF}.err').read()[-1500:])
PY
