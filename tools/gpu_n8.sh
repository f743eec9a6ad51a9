# eight GPUs: the driver's scaling point for the headline config, the short-round configs, c5 strong scaling, the reference arm under torchrun
N=${1:-8}
bash tools/gpu_scale.sh $N
SUFFIX=_c1 bash tools/gpu_scale.sh $N --workload c1_spheres_bezier --e2e-rounds 0
SUFFIX=_c2 bash tools/gpu_scale.sh $N --workload c2_bunny_chess --e2e-rounds 0
SUFFIX=_c1_torch bash tools/gpu_scale.sh $N --workload c1_spheres_bezier --e2e-rounds 0 --collective torch
SUFFIX=_c2_torch bash tools/gpu_scale.sh $N --workload c2_bunny_chess --e2e-rounds 0 --collective torch
SUFFIX=_c4 bash tools/gpu_scale.sh $N --workload c4_bump_dof --e2e-rounds 0 --steps 3 --warmup 2
SUFFIX=_c5 bash tools/gpu_scale.sh $N --workload c5_dragon_4096 --e2e-rounds 0 --steps 2 --warmup 1
SUFFIX=_strong bash tools/gpu_scale.sh $N --photons $((16777216 / N)) --e2e-rounds 0
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/scale_${N}_ref.json 2> gpurun_out/scale_${N}_ref.err; python -c "
import json; d=json.loads(open('gpurun_out/scale_${N}_ref.json').read().strip().splitlines()[-1]); print('ref arm N=$N', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['omp_num_threads_env'])"
