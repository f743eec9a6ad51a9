# two GPUs: the 2-GPU C++ host test, bench at N=2 with both collective paths, the reference arm under torchrun, c1/c2 short rounds
python -m pytest tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_scale.sh 2
SUFFIX=_torch bash tools/gpu_scale.sh 2 --collective torch --e2e-rounds 0 --f64-too 0
SUFFIX=_c1 bash tools/gpu_scale.sh 2 --workload c1_spheres_bezier --e2e-rounds 0
SUFFIX=_c2 bash tools/gpu_scale.sh 2 --workload c2_bunny_chess --e2e-rounds 0
SUFFIX=_c1_torch bash tools/gpu_scale.sh 2 --workload c1_spheres_bezier --e2e-rounds 0 --collective torch
SUFFIX=_c2_torch bash tools/gpu_scale.sh 2 --workload c2_bunny_chess --e2e-rounds 0 --collective torch
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/scale_2_ref.json 2> gpurun_out/scale_2_ref.err; python -c "
import json; d=json.loads(open('gpurun_out/scale_2_ref.json').read().strip().splitlines()[-1]); print('ref arm N=2', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['omp_num_threads_env'], d['cpu_baseline']['as_shipped']['value'] if d['cpu_baseline']['as_shipped'] else None)"
